// gsr_densify.cu — clone / split / delete of the reference's adaptive density control applied to the FLAT
// parameter buffer and both Adam moments in one pass.
//
// Reference: splat_py/trainer.py:114-206 (delete_gaussians, clone_gaussians, split_gaussians) together with the
// optimizer surgery of splat_py/optimizer_manager.py:74-172.  There every one of the 6 parameter tensors, their 12
// Adam moment tensors and 3 statistics tensors goes through a boolean-mask gather (nonzero + host sync + gather)
// for the delete, a torch.cat for the clone, another mask gather and another cat for the split, and each tensor is
// re-wrapped as a new nn.Parameter: ~60 full passes over the parameter set, each with its own allocation.
//
// Here the densification PLAN (which old row every new row comes from, and what is done to it) is a set of small
// per-row index arrays, and this kernel applies it: every output element of the new flat parameter buffer and of
// both moment buffers is produced exactly once, from one read of the old buffers — 3 x 236 B read and written
// per surviving Gaussian, HBM-bound, no intermediate copies.
//
//   out row r  <-  in row src[r]
//   clone_row[r] >= 0 : a clone — xyz -= xyz_sub[clone_row[r]]        (trainer.py:126: `cloned_xyz -= avg * 0.01`)
//   split_row[r] >= 0 : a split sample — xyz += xyz_add[split_row[r]]  (trainer.py:191)
//                       quaternion = q_set[...]  (normalised, :184), scale = scale_set[...]  (:194)
//   a row with clone_row >= 0 or split_row >= 0 is NEW: its Adam moments are zero (optimizer_manager.py:112-160);
//   every other row keeps the moments of its source row (:74-86).
// A cloned Gaussian that is split in the same pass has both (the reference splits the already-moved clone).
// The arithmetic is one correctly rounded fp32 subtract / add, i.e. the values torch's `-=` / `+=` produce.
#include <cstdint>

#include "gsr_common.cuh"

namespace gsr {

constexpr int DENSIFY_SECTIONS = 6;  // xyz 3 | quaternion 4 | scale 3 | opacity 1 | rgb 3 | sh 3K

struct DensifyLayout {
    long long in_off[DENSIFY_SECTIONS + 1];   // element offset of every section in the old flat buffer
    long long out_off[DENSIFY_SECTIONS + 1];  // ... in the new one (section starts are 16-byte aligned)
    int width[DENSIFY_SECTIONS];
};

__global__ void __launch_bounds__(256)
    k_densify_apply(int n_out, DensifyLayout L, const float* __restrict__ p_in, const float* __restrict__ m_in,
                    const float* __restrict__ v_in, const int32_t* __restrict__ src,
                    const int32_t* __restrict__ clone_row, const int32_t* __restrict__ split_row,
                    const float* __restrict__ xyz_sub, const float* __restrict__ xyz_add,
                    const float* __restrict__ q_set, const float* __restrict__ scale_set, float* __restrict__ p_out,
                    float* __restrict__ m_out, float* __restrict__ v_out) {
    const int s = blockIdx.y;  // section
    const int w = L.width[s];
    if (w == 0) return;
    const long long n_el = (long long)n_out * w;
    const long long padded = L.out_off[s + 1] - L.out_off[s];  // includes the alignment padding (< 4 floats)
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < padded; e += stride) {
        const long long o = L.out_off[s] + e;
        if (e >= n_el) {  // padding between sections
            p_out[o] = 0.0f;
            if (m_out != nullptr) { m_out[o] = 0.0f; v_out[o] = 0.0f; }
            continue;
        }
        const int r = (int)(e / w), k = (int)(e - (long long)r * w);
        const long long i = L.in_off[s] + (long long)src[r] * w + k;
        const int cr = clone_row != nullptr ? clone_row[r] : -1;
        const int sr = split_row != nullptr ? split_row[r] : -1;
        float val = p_in[i];
        if (s == 0) {
            if (cr >= 0) val = __fsub_rn(val, xyz_sub[(size_t)cr * 3 + k]);
            if (sr >= 0) val = __fadd_rn(val, xyz_add[(size_t)sr * 3 + k]);
        } else if (s == 1) {
            if (sr >= 0) val = q_set[(size_t)sr * 4 + k];
        } else if (s == 2) {
            if (sr >= 0) val = scale_set[(size_t)sr * 3 + k];
        }
        p_out[o] = val;
        if (m_out != nullptr) {
            const bool fresh = (cr >= 0) | (sr >= 0);
            m_out[o] = fresh ? 0.0f : m_in[i];
            v_out[o] = fresh ? 0.0f : v_in[i];
        }
    }
}

static void densify_layout(long long n, int n_sh_rest, long long* off, int* width) {
    const int w[DENSIFY_SECTIONS] = {3, 4, 3, 1, 3, 3 * n_sh_rest};
    off[0] = 0;
    for (int i = 0; i < DENSIFY_SECTIONS; ++i) {
        off[i + 1] = (off[i] + n * w[i] + 3) & ~3ll;
        if (width != nullptr) width[i] = w[i];
    }
}

}  // namespace gsr

using namespace gsr;

extern "C" {

int64_t gsr_flat_numel(int64_t n_gaussians, int n_sh_rest) {
    long long off[DENSIFY_SECTIONS + 1];
    densify_layout(n_gaussians, n_sh_rest, off, nullptr);
    return off[DENSIFY_SECTIONS];
}

int gsr_densify_apply(int n_in, int n_out, int n_sh_rest, const float* p_in, const float* m_in, const float* v_in,
                      const int32_t* src, const int32_t* clone_row, const int32_t* split_row, const float* xyz_sub,
                      const float* xyz_add, const float* q_set, const float* scale_set, float* p_out, float* m_out,
                      float* v_out, void* stream) {
    if (n_in < 0 || n_out < 0 || n_sh_rest < 0) return GSR_ERR_BAD_ARG;
    if ((m_in == nullptr) != (v_in == nullptr) || (m_out == nullptr) != (v_out == nullptr) ||
        (m_in == nullptr) != (m_out == nullptr))
        return GSR_ERR_BAD_ARG;
    if (clone_row != nullptr && xyz_sub == nullptr) return GSR_ERR_BAD_ARG;
    if (split_row != nullptr && (xyz_add == nullptr || q_set == nullptr || scale_set == nullptr)) return GSR_ERR_BAD_ARG;
    if (n_out == 0) return GSR_OK;
    DensifyLayout L;
    densify_layout(n_in, n_sh_rest, L.in_off, L.width);
    densify_layout(n_out, n_sh_rest, L.out_off, nullptr);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long widest = (long long)n_out * (n_sh_rest > 0 ? 3 * n_sh_rest : 4);
    long long bx = (widest + 255) / 256;
    const long long cap = (long long)sms * 8;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    const dim3 grid((unsigned)bx, DENSIFY_SECTIONS);
    k_densify_apply<<<grid, 256, 0, (cudaStream_t)stream>>>(n_out, L, p_in, m_in, v_in, src, clone_row, split_row,
                                                            xyz_sub, xyz_add, q_set, scale_set, p_out, m_out, v_out);
    return (int)cudaGetLastError();
}

}  // extern "C"
