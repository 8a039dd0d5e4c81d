// gsr_pergaussian.cu — the reference's eight projection operators and the SH
// precompute pair as stand-alone kernels (C ABI in include/gsr_b200.h).
//
// These exist so the reference's splat_py/cuda_autograd_functions.py runs
// unmodified on this library (fp32 and fp64).  The training path does not use
// them: it goes through the fused kernel in gsr_preprocess.cu, which calls the
// same device functions (gsr_math.cuh) and therefore produces the same bits.
#include "gsr_common.cuh"
#include "gsr_math.cuh"
#include "gsr_math_bwd.cuh"

namespace gsr {

constexpr int PG_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(PG_THREADS) k_camera_projection(int N, const T* __restrict__ xyz,
                                                                  const T* __restrict__ K,
                                                                  T* __restrict__ uv) {
    const int i = blockIdx.x * PG_THREADS + threadIdx.x;
    if (i >= N) return;
    T u, v;
    project_uv<T>(xyz[i * 3 + 0], xyz[i * 3 + 1], xyz[i * 3 + 2], K[0], K[2], K[4], K[5], u, v);
    uv[i * 2 + 0] = u;
    uv[i * 2 + 1] = v;
}

template <typename T>
__global__ void __launch_bounds__(PG_THREADS)
    k_camera_projection_bwd(int N, const T* __restrict__ xyz, const T* __restrict__ K,
                            const T* __restrict__ guv, T* __restrict__ gxyz) {
    const int i = blockIdx.x * PG_THREADS + threadIdx.x;
    if (i >= N) return;
    T g[3];
    if (project_uv_bwd<T>(xyz[i * 3 + 0], xyz[i * 3 + 1], xyz[i * 3 + 2], K[0], K[4], guv[i * 2 + 0],
                          guv[i * 2 + 1], g)) {
        gxyz[i * 3 + 0] = g[0];
        gxyz[i * 3 + 1] = g[1];
        gxyz[i * 3 + 2] = g[2];
    }
}

template <typename T>
__global__ void __launch_bounds__(PG_THREADS) k_sigma_world(int N, const T* __restrict__ q,
                                                            const T* __restrict__ s,
                                                            T* __restrict__ out) {
    const int i = blockIdx.x * PG_THREADS + threadIdx.x;
    if (i >= N) return;
    T S6[6], S9[9];
    sigma_world<T>(q[i * 4 + 0], q[i * 4 + 1], q[i * 4 + 2], q[i * 4 + 3], s[i * 3 + 0], s[i * 3 + 1],
                   s[i * 3 + 2], S6);
    sym6_to_full(S6, S9);
#pragma unroll
    for (int k = 0; k < 9; ++k) out[i * 9 + k] = S9[k];
}

template <typename T>
__global__ void __launch_bounds__(PG_THREADS)
    k_sigma_world_bwd(int N, const T* __restrict__ q, const T* __restrict__ s, const T* __restrict__ G,
                      T* __restrict__ gq, T* __restrict__ gs) {
    const int i = blockIdx.x * PG_THREADS + threadIdx.x;
    if (i >= N) return;
    T Gl[9], oq[4], os[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) Gl[k] = G[i * 9 + k];
    sigma_world_bwd<T>(q[i * 4 + 0], q[i * 4 + 1], q[i * 4 + 2], q[i * 4 + 3], s[i * 3 + 0],
                       s[i * 3 + 1], s[i * 3 + 2], Gl, oq, os);
#pragma unroll
    for (int k = 0; k < 4; ++k) gq[i * 4 + k] = oq[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) gs[i * 3 + k] = os[k];
}

template <typename T>
__global__ void __launch_bounds__(PG_THREADS) k_jacobian(int N, const T* __restrict__ xyz,
                                                         const T* __restrict__ K, T* __restrict__ J) {
    const int i = blockIdx.x * PG_THREADS + threadIdx.x;
    if (i >= N) return;
    T Jl[6];
    proj_jacobian<T>(xyz[i * 3 + 0], xyz[i * 3 + 1], xyz[i * 3 + 2], K[0], K[4], Jl);
#pragma unroll
    for (int k = 0; k < 6; ++k) J[i * 6 + k] = Jl[k];
}

template <typename T>
__global__ void __launch_bounds__(PG_THREADS)
    k_jacobian_bwd(int N, const T* __restrict__ xyz, const T* __restrict__ K, const T* __restrict__ gJ,
                   T* __restrict__ gxyz) {
    const int i = blockIdx.x * PG_THREADS + threadIdx.x;
    if (i >= N) return;
    T gl[6], g[3];
#pragma unroll
    for (int k = 0; k < 6; ++k) gl[k] = gJ[i * 6 + k];
    proj_jacobian_bwd<T>(xyz[i * 3 + 0], xyz[i * 3 + 1], xyz[i * 3 + 2], K[0], K[4], gl, g);
#pragma unroll
    for (int k = 0; k < 3; ++k) gxyz[i * 3 + k] = g[k];
}

template <typename T>
__device__ __forceinline__ void load_W(const T* __restrict__ M, T* __restrict__ W) {
    W[0] = M[0]; W[1] = M[1]; W[2] = M[2];
    W[3] = M[4]; W[4] = M[5]; W[5] = M[6];
    W[6] = M[8]; W[7] = M[9]; W[8] = M[10];
}

template <typename T>
__global__ void __launch_bounds__(PG_THREADS) k_conic(int N, const T* __restrict__ S,
                                                      const T* __restrict__ J, const T* __restrict__ M,
                                                      T* __restrict__ conic) {
    const int i = blockIdx.x * PG_THREADS + threadIdx.x;
    if (i >= N) return;
    T W[9], Sl[9], Jl[6], c[3];
    load_W<T>(M, W);
#pragma unroll
    for (int k = 0; k < 9; ++k) Sl[k] = S[i * 9 + k];
#pragma unroll
    for (int k = 0; k < 6; ++k) Jl[k] = J[i * 6 + k];
    conic_from<T>(Sl, Jl, W, c, nullptr);
#pragma unroll
    for (int k = 0; k < 3; ++k) conic[i * 3 + k] = c[k];
}

template <typename T>
__global__ void __launch_bounds__(PG_THREADS)
    k_conic_bwd(int N, const T* __restrict__ S, const T* __restrict__ J, const T* __restrict__ M,
                const T* __restrict__ gc, T* __restrict__ gS, T* __restrict__ gJ) {
    const int i = blockIdx.x * PG_THREADS + threadIdx.x;
    if (i >= N) return;
    T W[9], Sl[9], Jl[6], gcl[3], oS[9], oJ[6];
    load_W<T>(M, W);
#pragma unroll
    for (int k = 0; k < 9; ++k) Sl[k] = S[i * 9 + k];
#pragma unroll
    for (int k = 0; k < 6; ++k) Jl[k] = J[i * 6 + k];
#pragma unroll
    for (int k = 0; k < 3; ++k) gcl[k] = gc[i * 3 + k];
    conic_bwd<T>(Sl, Jl, W, gcl, oS, oJ);
#pragma unroll
    for (int k = 0; k < 9; ++k) gS[i * 9 + k] = oS[k];
#pragma unroll
    for (int k = 0; k < 6; ++k) gJ[i * 6 + k] = oJ[k];
}

// SH -> RGB for one gaussian.  Reference: src/precompute_sh.cu:22-56
template <typename T, int N_SH>
__global__ void __launch_bounds__(PG_THREADS)
    k_sh_to_rgb(int N, const T* __restrict__ xyz, const T* __restrict__ sh, const T* __restrict__ M,
                T* __restrict__ rgb) {
    const int i = blockIdx.x * PG_THREADS + threadIdx.x;
    if (i >= N) return;
    if (N_SH == 1) {
#pragma unroll
        for (int c = 0; c < 3; ++c) rgb[i * 3 + c] = sh[i * 3 + c];
        return;
    }
    T dx, dy, dz, Y[N_SH];
    view_dir<T>(xyz[i * 3 + 0], xyz[i * 3 + 1], xyz[i * 3 + 2], M[3], M[7], M[11], dx, dy, dz);
    sh_basis<T, N_SH>(dx, dy, dz, Y);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        T acc = T(0);
#pragma unroll
        for (int k = 0; k < N_SH; ++k) acc += Y[k] * sh[(size_t)i * 3 * N_SH + c * N_SH + k];
        rgb[i * 3 + c] = acc * T(GSR_RSH0);
    }
}

// Reference: src/precompute_sh.cu:75-109
template <typename T, int N_SH>
__global__ void __launch_bounds__(PG_THREADS)
    k_sh_to_rgb_bwd(int N, const T* __restrict__ xyz, const T* __restrict__ M,
                    const T* __restrict__ grgb, T* __restrict__ gsh) {
    const int i = blockIdx.x * PG_THREADS + threadIdx.x;
    if (i >= N) return;
    if (N_SH == 1) {
#pragma unroll
        for (int c = 0; c < 3; ++c) gsh[i * 3 + c] = grgb[i * 3 + c];
        return;
    }
    T dx, dy, dz, Y[N_SH];
    view_dir<T>(xyz[i * 3 + 0], xyz[i * 3 + 1], xyz[i * 3 + 2], M[3], M[7], M[11], dx, dy, dz);
    sh_basis<T, N_SH>(dx, dy, dz, Y);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const T g = grgb[i * 3 + c] * T(GSR_RSH0);
#pragma unroll
        for (int k = 0; k < N_SH; ++k) gsh[(size_t)i * 3 * N_SH + c * N_SH + k] = g * Y[k];
    }
}

}  // namespace gsr

using namespace gsr;

#define GSR_GRID(N) dim3(((N) + PG_THREADS - 1) / PG_THREADS), dim3(PG_THREADS), 0, (cudaStream_t)stream
#define GSR_DISPATCH_DTYPE(CALL_F, CALL_D)   \
    if (N <= 0) return GSR_OK;               \
    if (dtype == GSR_F32) { CALL_F; }        \
    else if (dtype == GSR_F64) { CALL_D; }   \
    else return GSR_ERR_BAD_ARG;             \
    return (int)cudaGetLastError();

extern "C" {

int gsr_camera_projection(int dtype, int N, const void* xyz, const void* K, void* uv, void* stream) {
    GSR_DISPATCH_DTYPE(
        (k_camera_projection<float><<<GSR_GRID(N)>>>(N, (const float*)xyz, (const float*)K, (float*)uv)),
        (k_camera_projection<double><<<GSR_GRID(N)>>>(N, (const double*)xyz, (const double*)K, (double*)uv)))
}

int gsr_camera_projection_backward(int dtype, int N, const void* xyz, const void* K,
                                   const void* g, void* out, void* stream) {
    GSR_DISPATCH_DTYPE(
        (k_camera_projection_bwd<float><<<GSR_GRID(N)>>>(N, (const float*)xyz, (const float*)K, (const float*)g, (float*)out)),
        (k_camera_projection_bwd<double><<<GSR_GRID(N)>>>(N, (const double*)xyz, (const double*)K, (const double*)g, (double*)out)))
}

int gsr_compute_sigma_world(int dtype, int N, const void* q, const void* s, void* out, void* stream) {
    GSR_DISPATCH_DTYPE(
        (k_sigma_world<float><<<GSR_GRID(N)>>>(N, (const float*)q, (const float*)s, (float*)out)),
        (k_sigma_world<double><<<GSR_GRID(N)>>>(N, (const double*)q, (const double*)s, (double*)out)))
}

int gsr_compute_sigma_world_backward(int dtype, int N, const void* q, const void* s, const void* G,
                                     void* gq, void* gs, void* stream) {
    GSR_DISPATCH_DTYPE(
        (k_sigma_world_bwd<float><<<GSR_GRID(N)>>>(N, (const float*)q, (const float*)s, (const float*)G, (float*)gq, (float*)gs)),
        (k_sigma_world_bwd<double><<<GSR_GRID(N)>>>(N, (const double*)q, (const double*)s, (const double*)G, (double*)gq, (double*)gs)))
}

int gsr_compute_projection_jacobian(int dtype, int N, const void* xyz, const void* K, void* J,
                                    void* stream) {
    GSR_DISPATCH_DTYPE(
        (k_jacobian<float><<<GSR_GRID(N)>>>(N, (const float*)xyz, (const float*)K, (float*)J)),
        (k_jacobian<double><<<GSR_GRID(N)>>>(N, (const double*)xyz, (const double*)K, (double*)J)))
}

int gsr_compute_projection_jacobian_backward(int dtype, int N, const void* xyz, const void* K,
                                             const void* gJ, void* out, void* stream) {
    GSR_DISPATCH_DTYPE(
        (k_jacobian_bwd<float><<<GSR_GRID(N)>>>(N, (const float*)xyz, (const float*)K, (const float*)gJ, (float*)out)),
        (k_jacobian_bwd<double><<<GSR_GRID(N)>>>(N, (const double*)xyz, (const double*)K, (const double*)gJ, (double*)out)))
}

int gsr_compute_conic(int dtype, int N, const void* S, const void* J, const void* M, void* conic,
                      void* stream) {
    GSR_DISPATCH_DTYPE(
        (k_conic<float><<<GSR_GRID(N)>>>(N, (const float*)S, (const float*)J, (const float*)M, (float*)conic)),
        (k_conic<double><<<GSR_GRID(N)>>>(N, (const double*)S, (const double*)J, (const double*)M, (double*)conic)))
}

int gsr_compute_conic_backward(int dtype, int N, const void* S, const void* J, const void* M,
                               const void* gc, void* gS, void* gJ, void* stream) {
    GSR_DISPATCH_DTYPE(
        (k_conic_bwd<float><<<GSR_GRID(N)>>>(N, (const float*)S, (const float*)J, (const float*)M, (const float*)gc, (float*)gS, (float*)gJ)),
        (k_conic_bwd<double><<<GSR_GRID(N)>>>(N, (const double*)S, (const double*)J, (const double*)M, (const double*)gc, (double*)gS, (double*)gJ)))
}

#define GSR_SH_CASES(KERNEL, T, ...)                                                       \
    switch (n_sh) {                                                                        \
        case 1: KERNEL<T, 1><<<GSR_GRID(N)>>>(__VA_ARGS__); break;                         \
        case 4: KERNEL<T, 4><<<GSR_GRID(N)>>>(__VA_ARGS__); break;                         \
        case 9: KERNEL<T, 9><<<GSR_GRID(N)>>>(__VA_ARGS__); break;                         \
        case 16: KERNEL<T, 16><<<GSR_GRID(N)>>>(__VA_ARGS__); break;                       \
        default: return GSR_ERR_UNSUPPORTED;                                               \
    }

int gsr_precompute_rgb_from_sh(int dtype, int N, int n_sh, const void* xyz, const void* sh,
                               const void* M, void* rgb, void* stream) {
    GSR_DISPATCH_DTYPE(
        GSR_SH_CASES(k_sh_to_rgb, float, N, (const float*)xyz, (const float*)sh, (const float*)M, (float*)rgb),
        GSR_SH_CASES(k_sh_to_rgb, double, N, (const double*)xyz, (const double*)sh, (const double*)M, (double*)rgb))
}

int gsr_precompute_rgb_from_sh_backward(int dtype, int N, int n_sh, const void* xyz, const void* M,
                                        const void* grgb, void* gsh, void* stream) {
    GSR_DISPATCH_DTYPE(
        GSR_SH_CASES(k_sh_to_rgb_bwd, float, N, (const float*)xyz, (const float*)M, (const float*)grgb, (float*)gsh),
        GSR_SH_CASES(k_sh_to_rgb_bwd, double, N, (const double*)xyz, (const double*)M, (const double*)grgb, (double*)gsh))
}

}  // extern "C"
