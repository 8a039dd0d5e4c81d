// gsr_adam.cu — Adam update of the flat Gaussian-parameter buffer.
//
// Contract (values): torch.optim.Adam as the reference's OptimizerManager configures it
// (splat_py/optimizer_manager.py:13-44: one parameter group per field with its own learning rate, betas
// (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad).  Per element, in float, in the operation order of
// torch's CUDA kernels (torch/optim/adam.py `_multi_tensor_adam`: lerp_, mul_, addcmul_, sqrt, div by the
// bias-correction scalar (a multiplication by its float reciprocal on CUDA), add eps, addcdiv_):
//     m <- fma(1-b1, g - m, m)
//     v <- fma((1-b2) * g, g, v * b2)
//     d <- sqrt(v) * (1 / sqrt(1 - b2^t)) + eps
//     p <- fma(-lr / (1 - b1^t), m / d, p)
//
// Layout: ONE flat fp32 buffer [xyz 3N | quaternion 4N | scale 3N | opacity N | rgb 3N | sh 45N], every section
// starting on a 16-byte boundary — the layout gsr_preprocess_backward's gradients already have — so the whole
// optimizer step is one streaming kernel: 16 B read + 12 B written per parameter (p, g, m, v in; p, m, v out),
// float4 accesses, persistent grid of a multiple of the SM count.  HBM-bound by construction.
//
// Sharded form (one process per GPU, view-parallel training): rank r owns elements [lo, hi).  It reads the
// gradient of its range from EVERY rank's gradient buffer over NVLink (peer loads; summed in rank order and
// scaled by 1/world), updates its own shard of m and v (optimizer state is sharded: 1/world of the memory),
// and stores the new parameters into EVERY rank's parameter buffer (peer stores).  That is reduce-scatter +
// Adam + all-gather in one kernel: each gradient / parameter byte crosses NVLink once, the optimizer math
// rides along, and every replica receives bit-identical parameters because each element is computed once.
// The caller brackets the launch with cross-rank barriers (gradients complete before, parameters visible after).
#include <cmath>
#include <cstdint>

#include "gsr_common.cuh"

namespace gsr {

constexpr int ADAM_MAX_SECTIONS = 8;
constexpr int ADAM_MAX_PEERS = 16;

struct AdamSections {
    int n;
    long long end4[ADAM_MAX_SECTIONS];  // exclusive end of each section, in float4 units
    float neg_step[ADAM_MAX_SECTIONS];  // -lr / (1 - beta1^t)
};
struct AdamPeers {
    int world;
    const float4* grad[ADAM_MAX_PEERS];
    float4* param[ADAM_MAX_PEERS];
};

__device__ __forceinline__ float adam_one(float p, float g, float& m, float& v, float w1, float beta2, float w2,
                                          float inv_bias2_sqrt, float eps, float neg_step) {
    m = __fmaf_rn(w1, __fsub_rn(g, m), m);
    v = __fmaf_rn(__fmul_rn(w2, g), g, __fmul_rn(v, beta2));
    const float d = __fadd_rn(__fmul_rn(__fsqrt_rn(v), inv_bias2_sqrt), eps);
    return __fmaf_rn(neg_step, __fdiv_rn(m, d), p);
}

__device__ __forceinline__ float4 ld_peer(const float4* p) {  // peer memory changes between steps: never L1
    float4 v;
    asm volatile("ld.global.cv.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// i4 indexes float4 elements of the flat buffer; m / v are indexed relative to lo4 (the shard).
template <bool SHARDED>
__global__ void __launch_bounds__(256)
    k_adam(long long lo4, long long hi4, float4* __restrict__ p, const float4* __restrict__ g,
           float4* __restrict__ m, float4* __restrict__ v, AdamSections sec, AdamPeers peers, float w1, float beta2,
           float w2, float inv_bias2_sqrt, float eps, float inv_world) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = lo4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi4; i += stride) {
        int s = 0;
#pragma unroll
        for (int k = 0; k < ADAM_MAX_SECTIONS - 1; ++k) s += (k < sec.n - 1 && i >= sec.end4[k]) ? 1 : 0;
        const float ns = sec.neg_step[s];
        float4 gg;
        if (SHARDED) {
            gg = ld_peer(peers.grad[0] + i);
            for (int q = 1; q < peers.world; ++q) {
                const float4 t = ld_peer(peers.grad[q] + i);
                gg.x += t.x; gg.y += t.y; gg.z += t.z; gg.w += t.w;
            }
            gg.x *= inv_world; gg.y *= inv_world; gg.z *= inv_world; gg.w *= inv_world;
        } else {
            gg = g[i];
        }
        float4 pp = p[i], mm = m[i - lo4], vv = v[i - lo4];
        pp.x = adam_one(pp.x, gg.x, mm.x, vv.x, w1, beta2, w2, inv_bias2_sqrt, eps, ns);
        pp.y = adam_one(pp.y, gg.y, mm.y, vv.y, w1, beta2, w2, inv_bias2_sqrt, eps, ns);
        pp.z = adam_one(pp.z, gg.z, mm.z, vv.z, w1, beta2, w2, inv_bias2_sqrt, eps, ns);
        pp.w = adam_one(pp.w, gg.w, mm.w, vv.w, w1, beta2, w2, inv_bias2_sqrt, eps, ns);
        m[i - lo4] = mm;
        v[i - lo4] = vv;
        if (SHARDED) {
            for (int q = 0; q < peers.world; ++q) peers.param[q][i] = pp;
        } else {
            p[i] = pp;
        }
    }
}

static int fill_sections(AdamSections& sec, int n_sections, const int64_t* section_end, const double* section_lr,
                         double bias1) {
    if (n_sections < 1 || n_sections > ADAM_MAX_SECTIONS) return GSR_ERR_BAD_ARG;
    sec.n = n_sections;
    for (int k = 0; k < n_sections; ++k) {
        if (section_end[k] % 4 != 0) return GSR_ERR_BAD_ARG;  // sections are 16-byte aligned
        sec.end4[k] = section_end[k] / 4;
        sec.neg_step[k] = (float)(-section_lr[k] / bias1);
    }
    for (int k = n_sections; k < ADAM_MAX_SECTIONS; ++k) {
        sec.end4[k] = sec.end4[n_sections - 1];
        sec.neg_step[k] = 0.0f;
    }
    return 0;
}

static int adam_grid(long long n4) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long want = (n4 + 255) / 256;
    const long long cap = (long long)sms * 8;  // 8 resident CTAs of 256 threads per SM
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

// Per-step densification statistics (splat_py/trainer.py:376-385), one pass instead of two boolean-mask
// scatters (each with a nonzero() host sync) and four elementwise kernels:
//   uv_grad_accum[i]  += |uv_grad[r] * (fx, fy)|   for the r-th visible gaussian i = vis_idx[r]
//   grad_accum_count[i] += 1                        for the same i
//   xyz_grad_accum[i] += |xyz_grad[i]|              for every gaussian
// The scaling by the focal lengths is applied to uv_grad IN PLACE, as the reference does (:379-381).
__global__ void __launch_bounds__(256)
    k_densify_accumulate(int N, int M, const int32_t* __restrict__ vis_idx, float* __restrict__ uv_grad,
                         const float* __restrict__ xyz_grad, const float* __restrict__ K,
                         float* __restrict__ uv_acc, float* __restrict__ xyz_acc, int32_t* __restrict__ count) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < M) {
        const float fx = K[0], fy = K[4];
        const int i = vis_idx[t];
        const float gu = __fmul_rn(uv_grad[2 * t], fx), gv = __fmul_rn(uv_grad[2 * t + 1], fy);
        uv_grad[2 * t] = gu;
        uv_grad[2 * t + 1] = gv;
        uv_acc[2 * i] = __fadd_rn(uv_acc[2 * i], fabsf(gu));
        uv_acc[2 * i + 1] = __fadd_rn(uv_acc[2 * i + 1], fabsf(gv));
        count[i] += 1;
    }
    for (int e = t; e < 3 * N; e += gridDim.x * blockDim.x) xyz_acc[e] = __fadd_rn(xyz_acc[e], fabsf(xyz_grad[e]));
}

}  // namespace gsr

using namespace gsr;

extern "C" {

int gsr_densify_accumulate(int N, int M, const int32_t* vis_idx, float* uv_grad, const float* xyz_grad, const float* K,
                           float* uv_grad_accum, float* xyz_grad_accum, int32_t* grad_accum_count, void* stream) {
    if (N < 0 || M < 0 || M > N) return GSR_ERR_BAD_ARG;
    if (N == 0) return 0;
    const int blocks = (int)(((long long)3 * N + 255) / 256);
    k_densify_accumulate<<<blocks, 256, 0, (cudaStream_t)stream>>>(N, M, vis_idx, uv_grad, xyz_grad, K, uv_grad_accum,
                                                                   xyz_grad_accum, grad_accum_count);
    return (int)cudaGetLastError();
}

int gsr_adam_step(int64_t n, float* p, const float* g, float* m, float* v, int n_sections,
                  const int64_t* section_end, const double* section_lr, double beta1, double beta2, double eps, int step,
                  void* stream) {
    if (n < 0 || n % 4 != 0 || step < 1) return GSR_ERR_BAD_ARG;
    if (n == 0) return 0;
    const double bias1 = 1.0 - pow(beta1, (double)step), bias2 = 1.0 - pow(beta2, (double)step);
    AdamSections sec;
    if (int rc = fill_sections(sec, n_sections, section_end, section_lr, bias1)) return rc;
    AdamPeers peers;
    peers.world = 1;
    const float inv_b2s = 1.0f / (float)sqrt(bias2);
    k_adam<false><<<adam_grid(n / 4), 256, 0, (cudaStream_t)stream>>>(
        0, n / 4, reinterpret_cast<float4*>(p), reinterpret_cast<const float4*>(g), reinterpret_cast<float4*>(m),
        reinterpret_cast<float4*>(v), sec, peers, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2),
        inv_b2s, (float)eps, 1.0f);
    return (int)cudaGetLastError();
}

int gsr_adam_step_sharded(int64_t lo, int64_t hi, int world, const float* const* peer_grads,
                          float* const* peer_params, int self_rank, float* m_shard, float* v_shard, int n_sections,
                          const int64_t* section_end, const double* section_lr, double beta1, double beta2, double eps,
                          int step, void* stream) {
    if (lo < 0 || hi < lo || lo % 4 != 0 || hi % 4 != 0 || step < 1 || world < 1 || world > ADAM_MAX_PEERS ||
        self_rank < 0 || self_rank >= world)
        return GSR_ERR_BAD_ARG;
    if (hi == lo) return 0;
    const double bias1 = 1.0 - pow(beta1, (double)step), bias2 = 1.0 - pow(beta2, (double)step);
    AdamSections sec;
    if (int rc = fill_sections(sec, n_sections, section_end, section_lr, bias1)) return rc;
    AdamPeers peers;
    peers.world = world;
    for (int q = 0; q < world; ++q) {
        peers.grad[q] = reinterpret_cast<const float4*>(peer_grads[q]);
        peers.param[q] = reinterpret_cast<float4*>(peer_params[q]);
    }
    const float inv_b2s = 1.0f / (float)sqrt(bias2);
    k_adam<true><<<adam_grid((hi - lo) / 4), 256, 0, (cudaStream_t)stream>>>(
        lo / 4, hi / 4, reinterpret_cast<float4*>(peer_params[self_rank]), nullptr, reinterpret_cast<float4*>(m_shard),
        reinterpret_cast<float4*>(v_shard), sec, peers, (float)(1.0 - beta1), (float)beta2,
        (float)(1.0 - beta2), inv_b2s, (float)eps, 1.0f / (float)world);
    return (int)cudaGetLastError();
}

}  // extern "C"
