// gsr_common.cuh — shared declarations, TMA (cp.async.bulk) + mbarrier helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gsr_b200.h"

namespace gsr {

constexpr int TILE = 16;            // tile edge in pixels (splat_py/structs.py:4)
constexpr int TILE_PIXELS = 256;
constexpr int REC = GSR_REC_FLOATS; // floats per splat record

// record slots: three 16-byte lanes
//   q0 = (u, v, tau', opacity)     tau' = inflated Mahalanobis threshold beyond which alpha < 1/255
//   q1 = (a, 2b, c, det)           a = conic0 + 0.25, 2b = conic1, c = conic2 + 0.25, det = a*c - b*b
//   q2 = (rcp, colR, colG, colB)   rcp = correctly rounded 1/det,
//                                  col = SH_0 * rgb
enum RecSlot {
    R_U = 0, R_V = 1, R_R2 = 2, R_OPA = 3,
    R_A = 4, R_B2 = 5, R_C = 6, R_DET = 7,
    R_RCP = 8, R_CR = 9, R_CG = 10, R_CB = 11
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier.
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// 1-D TMA bulk copy shared -> global (bulk async-group completion).  Generic-proxy writes to the source
// must be ordered before it with fence_proxy_async_smem() + a barrier.
__device__ __forceinline__ void tma_store_1d(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void tma_store_commit_and_wait() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// float thresholds equivalent to the reference's float-vs-double-literal compares
// (src/render.cu:106,145,169; src/render_backward.cu:167-174); derivation in DESIGN.md.
//   (double)x >  0.9999        <=>  x >  0x3f7ff972 (0.9999f)
//   (double)x <  0.00392156862 <=>  x <= 0x3b808080
//   (double)x <  0.999         <=>  x <  0x3f7fbe77 (0.999f)
//   (double)x >  0.001         <=>  x >= 0x3a83126f (0.001f)
#define GSR_SAT_THRESH __uint_as_float(0x3f7ff972u)
#define GSR_ALPHA_SKIP_MAX __uint_as_float(0x3b808080u)
#define GSR_BG_THRESH __uint_as_float(0x3f7fbe77u)
#define GSR_BGW_MIN __uint_as_float(0x3a83126fu)
#define GSR_ALPHA_CLAMP __uint_as_float(0x3f7ff972u) /* fl32(0.9999) */

}  // namespace gsr
