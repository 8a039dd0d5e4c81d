// gsr_record.cuh — the 48-byte splat record staged by TMA into the tile renderers.
#pragma once
#include "gsr_common.cuh"
#include "gsr_math.cuh"

namespace gsr {

__device__ __forceinline__ void make_record(float u, float v, float c0, float c1, float c2, float opa,
                                            float r, float g, float b, float* __restrict__ rec) {
    // a, b, c, det: src/render.cu:117-127 (fp32 branch: +0.25 dilation)
    const float a = __fadd_rn(c0, 0.25f);
    const float c = __fadd_rn(c2, 0.25f);
    const float bh = __fmul_rn(c1, 0.5f);
    const float det = __fmaf_rn(a, c, -__fmul_rn(bh, bh));
    // refined reciprocal used by the 3-FMA exact division in the forward kernel
    float r0;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(det));
    const float e = __fmaf_rn(-det, r0, 1.0f);
    const float r1 = __fmaf_rn(r0, e, r0);
    rec[R_U] = u;
    rec[R_V] = v;
    rec[R_A] = a;
    rec[R_B2] = __fadd_rn(bh, bh);
    rec[R_C] = c;
    rec[R_DET] = det;
    rec[R_RCP] = r1;
    rec[R_RDET] = (float)(1.0 / (double)det);  // src/render_backward.cu:153
    rec[R_OPA] = opa;
    rec[R_CR] = __fmul_rn(r, GSR_SH0);  // sh_to_rgb with N_SH == 1: Y0 * rgb
    rec[R_CG] = __fmul_rn(g, GSR_SH0);
    rec[R_CB] = __fmul_rn(b, GSR_SH0);
}

}  // namespace gsr
