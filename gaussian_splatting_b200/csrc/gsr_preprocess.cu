// gsr_preprocess.cu — the fused per-Gaussian stage behind rasterize().
//
// One kernel replaces, for every Gaussian of the scene, the reference's chain
//   transform_points_torch        splat_py/utils.py:60-72
//   CameraPointProjection         src/projection.cu:8-19
//   frustum cull + 8 mask gathers splat_py/rasterize.py:33-75
//   sigmoid(opacity)              splat_py/rasterize.py:60-62
//   ComputeSigmaWorld             src/projection.cu:56-109
//   ComputeProjectionJacobian     src/projection.cu:154-175
//   ComputeConic                  src/projection.cu:213-257
//   cat(rgb, sh) + PrecomputeRGBFromSH   splat_py/rasterize.py:89-93, src/precompute_sh.cu:7-58
//   compute_num_splats_kernel     src/tile_culling.cu:124-177
// and writes one 48-byte splat record, a depth key, a visibility flag and the
// tiles-touched count per Gaussian.  No Sigma_world / J intermediates reach HBM.
// The backward kernel is the fused VJP of the same chain.
#include <cub/cub.cuh>

#include "gsr_common.cuh"
#include "gsr_math.cuh"
#include "gsr_math_bwd.cuh"
#include "gsr_record.cuh"

namespace gsr {

constexpr int PRE_THREADS = 128;  // gaussians per CTA; every per-gaussian array slice of a CTA is one contiguous,
                                  // 16-byte aligned run in HBM, moved by TMA bulk copies

struct ViewConsts {
    float T[16];
    float K[9];
    float cam[3];  // camera centre in world frame = inverse(camera_T_world)[:3, 3]
};

// Last column of inverse(T) with the bits torch.inverse / torch.linalg.inv_ex produce on this platform for one 4x4
// fp32 matrix (the reference's op, splat_py/rasterize.py:91-93).  torch runs cuSOLVER getrf + cuBLAS trsm x2 — 15
// micro-kernels; their arithmetic was pinned on a B200 by tools/dev/inverse_probe2.py (600 of 600 random rigid and
// general matrices bit-identical, LU factors and solution separately):
//   LU, partial pivoting (first maximum):  l_ik = a_ik * (1 / a_kk)   [IEEE reciprocal, then a multiply]
//                                          a_ij = fma(-l_ik, a_kj, a_ij)
//   L y = P e_3 forward, fma accumulation;  U x = y by columns from the last one (right-looking: the terms of a row
//   are subtracted in DESCENDING column order), fma accumulation, then one IEEE division by the diagonal.
// One thread; ~100 flops.
__device__ void camera_centre(const float* __restrict__ T, float* __restrict__ cam) {
    float A[4][4];
    int perm[4] = {0, 1, 2, 3};
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) A[r][c] = T[r * 4 + c];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int piv = k;
        float best = fabsf(A[k][k]);
#pragma unroll
        for (int r = k + 1; r < 4; ++r) {
            const float v = fabsf(A[r][k]);
            if (v > best) { best = v; piv = r; }
        }
#pragma unroll
        for (int r = k + 1; r < 4; ++r) {  // swap rows k and piv without dynamic indexing
            if (r == piv) {
#pragma unroll
                for (int c = 0; c < 4; ++c) { const float t = A[k][c]; A[k][c] = A[r][c]; A[r][c] = t; }
                const int tp = perm[k]; perm[k] = perm[r]; perm[r] = tp;
            }
        }
        const float rp = __frcp_rn(A[k][k]);
#pragma unroll
        for (int i = k + 1; i < 4; ++i) {
            const float l = __fmul_rn(A[i][k], rp);
            A[i][k] = l;
#pragma unroll
            for (int c = k + 1; c < 4; ++c) A[i][c] = __fmaf_rn(-l, A[k][c], A[i][c]);
        }
    }
    float y[4], x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float acc = (perm[i] == 3) ? 1.0f : 0.0f;
#pragma unroll
        for (int c = 0; c < i; ++c) acc = __fmaf_rn(-A[i][c], y[c], acc);
        y[i] = acc;
    }
#pragma unroll
    for (int i = 3; i >= 0; --i) {
        float acc = y[i];
#pragma unroll
        for (int c = 3; c > i; --c) acc = __fmaf_rn(-A[i][c], x[c], acc);
        x[i] = __fdiv_rn(acc, A[i][i]);
    }
    cam[0] = x[0];
    cam[1] = x[1];
    cam[2] = x[2];
}

__global__ void k_camera_centre(const float* __restrict__ T, float* __restrict__ cam) {
    if (threadIdx.x == 0 && blockIdx.x == 0) camera_centre(T, cam);
}

// Whole CTA: threads 0..27 fetch one constant each (one memory latency instead of 28 dependent-issue loads by a
// single thread); ends with a CTA barrier.  The camera centre is taken from `cam` when the caller supplies it,
// else derived from T by thread 0.
__device__ __forceinline__ void load_view_cta(const float* __restrict__ T, const float* __restrict__ K,
                                              const float* __restrict__ cam, ViewConsts& vc, bool need_centre) {
    const int t = threadIdx.x;
    if (t < 16) vc.T[t] = T[t];
    else if (t < 25) vc.K[t - 16] = K[t - 16];
    else if (t < 28 && cam != nullptr) vc.cam[t - 25] = cam[t - 25];
    __syncthreads();
    if (need_centre && cam == nullptr) {
        if (t == 0) camera_centre(vc.T, vc.cam);
        __syncthreads();
    }
}

// SH -> RGB exactly as src/precompute_sh.cu:29-56 evaluates it on cat(rgb_dc, sh_rest)
template <int N_SH>
__device__ __forceinline__ void sh_rgb_fused(const float* __restrict__ dc, const float* __restrict__ rest,
                                             float px, float py, float pz, const float* cam,
                                             float* __restrict__ out) {
    if (N_SH == 1) {
        out[0] = dc[0]; out[1] = dc[1]; out[2] = dc[2];
        return;
    }
    float dx, dy, dz, Y[N_SH];
    view_dir<float>(px, py, pz, cam[0], cam[1], cam[2], dx, dy, dz);
    sh_basis<float, N_SH>(dx, dy, dz, Y);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float acc = __fmaf_rn(Y[0], dc[c], 0.0f);
#pragma unroll
        for (int k = 1; k < N_SH; ++k) acc = __fmaf_rn(Y[k], rest[c * (N_SH - 1) + (k - 1)], acc);
        out[c] = __fmul_rn(acc, GSR_RSH0);
    }
}

// cooperative, coalesced smem <-> global copies for the last (partial) block, where TMA's 16-byte
// granularity does not hold
__device__ __forceinline__ void coop_copy(float* __restrict__ dst, const float* __restrict__ src, int n_floats) {
    for (int k = threadIdx.x; k < n_floats; k += PRE_THREADS) dst[k] = src[k];
}

template <int N_SH, bool HAS_SH>
__global__ void __launch_bounds__(PRE_THREADS, 7)
    k_preprocess_fwd(int N, const float* __restrict__ xyz, const float* __restrict__ xyz_cam, int cam_first,
                     const float* __restrict__ quat,
                     const float* __restrict__ scale, const float* __restrict__ opa_logit,
                     const float* __restrict__ rgb_dc, const float* __restrict__ sh_rest,
                     const float* __restrict__ Tdev, const float* __restrict__ Kdev,
                     const float* __restrict__ camdev, float width, float height, float near_t, float far_t,
                     float pad, float mh, int ntx, int nty, uint32_t depth_base,
                     float* __restrict__ records, uint32_t* __restrict__ zkey,
                     uint8_t* __restrict__ visible, uint64_t* __restrict__ packed, uint64_t* __restrict__ tile_mask,
                     uint32_t* __restrict__ tile_win, int use_tma) {
    constexpr int NR3 = HAS_SH ? 3 * (N_SH - 1) : 1;
    __shared__ ViewConsts vc;
    __shared__ __align__(128) float s_sh[PRE_THREADS * NR3];       // SH coefficients of the CTA's gaussians
    __shared__ __align__(128) float s_rec[PRE_THREADS * REC];      // output records
    __shared__ __align__(8) uint64_t s_bar;
    const int tid = threadIdx.x;
    const int i0 = blockIdx.x * PRE_THREADS;
    const int cnt = min(PRE_THREADS, N - i0);
    const bool tma = use_tma && cnt == PRE_THREADS;
    // Latencies overlap instead of queueing up: the SH slice's bulk copy is issued first, every thread then
    // fetches its own gaussian's parameters, and only then the view constants are fetched (cooperatively).
    if (tid == 0 && HAS_SH && tma) {
        mbar_init(&s_bar, 1);
        fence_mbar_init();
        const uint32_t bytes = (uint32_t)(PRE_THREADS * NR3 * sizeof(float));
        mbar_arrive_expect_tx(&s_bar, bytes);
        tma_load_1d(s_sh, sh_rest + (size_t)i0 * NR3, bytes, &s_bar);
    }
    if (HAS_SH && !tma) coop_copy(s_sh, sh_rest + (size_t)i0 * NR3, cnt * NR3);
    const int i = i0 + tid;
    float x = 0.f, y = 0.f, z = 0.f, q0 = 1.f, q1 = 0.f, q2 = 0.f, q3 = 0.f, sc0 = 0.f, sc1 = 0.f, sc2 = 0.f,
          opl = 0.f, dc0 = 0.f, dc1 = 0.f, dc2 = 0.f;
    if (i < N) {
        x = xyz[i * 3 + 0]; y = xyz[i * 3 + 1]; z = xyz[i * 3 + 2];
        q0 = quat[i * 4 + 0]; q1 = quat[i * 4 + 1]; q2 = quat[i * 4 + 2]; q3 = quat[i * 4 + 3];
        sc0 = scale[i * 3 + 0]; sc1 = scale[i * 3 + 1]; sc2 = scale[i * 3 + 2];
        opl = opa_logit[i];
        dc0 = rgb_dc[i * 3 + 0]; dc1 = rgb_dc[i * 3 + 1]; dc2 = rgb_dc[i * 3 + 2];
    }
    load_view_cta(Tdev, Kdev, camdev, vc, HAS_SH);
    float rec[REC];
#pragma unroll
    for (int k = 0; k < REC; ++k) rec[k] = 0.0f;
    if (i < N) {
        float px, py, pz, u, v;
        if (xyz_cam != nullptr && i >= cam_first) {
            const float* pc = xyz_cam + (size_t)(i - cam_first) * 3;
            px = pc[0]; py = pc[1]; pz = pc[2];
        } else {
            transform_point<float>(vc.T, x, y, z, px, py, pz);
        }
        project_uv<float>(px, py, pz, vc.K[0], vc.K[2], vc.K[4], vc.K[5], u, v);
        // splat_py/rasterize.py:38-49 (strict compares, fp32)
        const bool culled = (pz < near_t) | (pz > far_t) | (u < -pad) | (u > width + pad) | (v < -pad) |
                            (v > height + pad);
        // visible => pz >= near > 0: float bits are monotone, and relative to bits(near) only a few low
        // bits are significant (shorter radix sort)
        zkey[i] = depth_base ? (__float_as_uint(pz) - depth_base) : depth_key(pz);
        // NaN coordinates compare false everywhere and would survive; the reference aborts on them
        // (splat_py/tile_culling.py:15-18) — treat as culled instead.
        const bool vis = !culled && (pz == pz) && (u == u) && (v == v);
        visible[i] = vis ? 1 : 0;
        uint64_t pk = 0ull;
        if (vis) {
            float S6[6], S9[9], J[6], W[9], conic[3];
            sigma_world<float>(q0, q1, q2, q3, sc0, sc1, sc2, S6);
            sym6_to_full(S6, S9);
            proj_jacobian<float>(px, py, pz, vc.K[0], vc.K[4], J);
            W[0] = vc.T[0]; W[1] = vc.T[1]; W[2] = vc.T[2];
            W[3] = vc.T[4]; W[4] = vc.T[5]; W[5] = vc.T[6];
            W[6] = vc.T[8]; W[7] = vc.T[9]; W[8] = vc.T[10];
            conic_from<float>(S9, J, W, conic, nullptr);

            const float opa = sigmoid_torch(opl);
            float dc[3] = {dc0, dc1, dc2};
            float col[3];
            if (HAS_SH) {
                if (tma) mbar_wait(&s_bar, 0);
                sh_rgb_fused<N_SH>(dc, s_sh + tid * NR3, x, y, z, vc.cam, col);  // stride 3*(N_SH-1): odd, no conflicts
            } else {
                col[0] = dc[0]; col[1] = dc[1]; col[2] = dc[2];
            }
            make_record(u, v, conic[0], conic[1], conic[2], opa, col[0], col[1], col[2], rec);

            // tiles touched (src/tile_culling.cu:139-176)
            Obb ob;
            compute_obb(u, v, rec[R_A], __fmul_rn(rec[R_B2], 0.5f), rec[R_C], mh, ob);
            int x0, x1, y0, y1, c = 0;
            tile_window(u, v, ob.radius_tiles, ntx, nty, x0, x1, y0, y1);
            // The tiles hit are remembered as a bit mask over the window (bit = (tx - x0) * height + (ty - y0), the
            // enumeration order) when the window has at most 64 tiles — footprints up to 3 sigma = 48 px — so that
            // the pair emission does not repeat the OBB / separating-axis tests (gsr_binning.cu); larger windows
            // are flagged and re-tested there.
            const int wx = x1 - x0, wy = y1 - y0;
            const bool small = (wx * wy <= 64) & (x0 < 256) & (y0 < 256);
            uint64_t hits = 0ull;
            for (int tx = x0; tx < x1; ++tx) {
                const float left = __fmul_rn(__int2float_rn(tx), 16.0f);
                const float right = __fmul_rn(__int2float_rn(tx + 1), 16.0f);
                for (int ty = y0; ty < y1; ++ty) {
                    const float top = __fmul_rn(__int2float_rn(ty), 16.0f);
                    const float bottom = __fmul_rn(__int2float_rn(ty + 1), 16.0f);
                    if (obb_hits_tile(ob, left, right, top, bottom)) {
                        ++c;
                        if (small) hits |= 1ull << ((tx - x0) * wy + (ty - y0));
                    }
                }
            }
            if (tile_mask != nullptr) {
                tile_mask[i] = hits;
                tile_win[i] = small ? ((uint32_t)x0 | ((uint32_t)y0 << 8) | ((uint32_t)wx << 16) | ((uint32_t)wy << 24))
                                    : 0xffffffffu;
            }
            pk = (1ull << 32) | (uint64_t)(uint32_t)c;
        }
        packed[i] = pk;
    }
    // records leave through shared memory: one 6 KB bulk store per CTA instead of 48-byte strided stores
    float4* sr = reinterpret_cast<float4*>(s_rec + tid * REC);
    sr[0] = make_float4(rec[0], rec[1], rec[2], rec[3]);
    sr[1] = make_float4(rec[4], rec[5], rec[6], rec[7]);
    sr[2] = make_float4(rec[8], rec[9], rec[10], rec[11]);
    if (tma) {
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            tma_store_1d(records + (size_t)i0 * REC, s_rec, (uint32_t)(PRE_THREADS * REC * sizeof(float)));
            tma_store_commit_and_wait();
        }
    } else {
        __syncthreads();
        coop_copy(records + (size_t)i0 * REC, s_rec, cnt * REC);
    }
    if (HAS_SH && tma && tid == 0) mbar_wait(&s_bar, 0);  // never exit with the SH load still in flight
}

// Fused VJP.  g_rgb/g_opa/g_uv/g_conic are the per-Gaussian sums produced by the render backward
// (indexed by original gaussian).  Order of the chain: SURVEY.md Appendix A "Per-Gaussian backward".
#ifndef GSR_PRE_BWD_MINB
#define GSR_PRE_BWD_MINB 6
#endif
template <int N_SH, bool HAS_SH>
__global__ void __launch_bounds__(PRE_THREADS, GSR_PRE_BWD_MINB)
    k_preprocess_bwd(int N, const float* __restrict__ xyz, const float* __restrict__ quat,
                     const float* __restrict__ scale, const float* __restrict__ opa_logit,
                     const float* __restrict__ Tdev, const float* __restrict__ Kdev,
                     const float* __restrict__ camdev, const uint8_t* __restrict__ visible,
                     const float* __restrict__ g_rgb,
                     const float* __restrict__ g_opa, const float* __restrict__ g_uv,
                     const float* __restrict__ g_conic, const float* __restrict__ g_rows, int rows_uv,
                     const float* __restrict__ g_uv_compact,
                     const uint64_t* __restrict__ scan, float* __restrict__ o_xyz,
                     float* __restrict__ o_quat, float* __restrict__ o_scale, float* __restrict__ o_opa,
                     float* __restrict__ o_dc, float* __restrict__ o_sh, int use_tma) {
    constexpr int NR = N_SH - 1;
    constexpr int NR3 = HAS_SH ? 3 * NR : 1;
    // all outputs of the CTA's 128 gaussians are staged in shared memory and leave as six bulk stores
    __shared__ ViewConsts vc;
    __shared__ __align__(128) float s_sh[PRE_THREADS * NR3];
    __shared__ __align__(128) float s_xyz[PRE_THREADS * 3];
    __shared__ __align__(128) float s_quat[PRE_THREADS * 4];
    __shared__ __align__(128) float s_scale[PRE_THREADS * 3];
    __shared__ __align__(128) float s_opa[PRE_THREADS];
    __shared__ __align__(128) float s_dc[PRE_THREADS * 3];
    const int tid = threadIdx.x;
    const int i0 = blockIdx.x * PRE_THREADS;
    const int cnt = min(PRE_THREADS, N - i0);
    const bool tma = use_tma && cnt == PRE_THREADS;
    const int i = i0 + tid;
    float gx[3] = {0.f, 0.f, 0.f}, gq[4] = {0.f, 0.f, 0.f, 0.f}, gs[3] = {0.f, 0.f, 0.f}, go = 0.f,
          gdc[3] = {0.f, 0.f, 0.f};
    float* my_sh = s_sh + tid * NR3;
    bool vis = false;
    if (i < N) vis = visible[i] != 0;
    // every thread fetches its gaussian's inputs before the view constants are fetched (cooperatively): the
    // memory latencies overlap instead of queueing up behind one thread's 28 loads
    float x = 0.f, y = 0.f, z = 0.f, qw = 1.f, qx = 0.f, qy = 0.f, qz = 0.f, s0 = 0.f, s1 = 0.f, s2 = 0.f, opl = 0.f,
          gcv[3] = {0.f, 0.f, 0.f}, guv[2] = {0.f, 0.f}, gov = 0.f, grv[3] = {0.f, 0.f, 0.f};
    if (vis) {
        x = xyz[i * 3 + 0]; y = xyz[i * 3 + 1]; z = xyz[i * 3 + 2];
        qw = quat[i * 4 + 0]; qx = quat[i * 4 + 1]; qy = quat[i * 4 + 2]; qz = quat[i * 4 + 3];
        s0 = scale[i * 3 + 0]; s1 = scale[i * 3 + 1]; s2 = scale[i * 3 + 2];
        opl = opa_logit[i];
        if (g_rows != nullptr) {  // interleaved rows: rgb3 opa | uv2 conic0 conic1 | conic2 pad3
            const float4* row = reinterpret_cast<const float4*>(g_rows + (size_t)i * GSR_GRAD_ROW_FLOATS);
            const float4 r0 = row[0], r1 = row[1];
            grv[0] = r0.x; grv[1] = r0.y; grv[2] = r0.z; gov = r0.w;
            if (rows_uv) { guv[0] = r1.x; guv[1] = r1.y; }
            gcv[0] = r1.z; gcv[1] = r1.w; gcv[2] = g_rows[(size_t)i * GSR_GRAD_ROW_FLOATS + 8];
        } else {
        gcv[0] = g_conic[i * 3 + 0]; gcv[1] = g_conic[i * 3 + 1]; gcv[2] = g_conic[i * 3 + 2];
        // gradient on the projected mean = the render backward's sum (g_uv, by gaussian; may be absent) + whatever
        // arrives on the COMPACT uv rasterize returned (g_uv_compact [M,2]; row = rank among the visible gaussians,
        // read off the forward's packed inclusive scan)
        if (g_uv != nullptr) { guv[0] = g_uv[i * 2 + 0]; guv[1] = g_uv[i * 2 + 1]; }
        gov = g_opa[i];
        grv[0] = g_rgb[i * 3 + 0]; grv[1] = g_rgb[i * 3 + 1]; grv[2] = g_rgb[i * 3 + 2];
        }
        if (g_uv_compact != nullptr) {
            const size_t r = (size_t)(scan[i] >> 32) - 1;
            const float2 t = *reinterpret_cast<const float2*>(g_uv_compact + r * 2);
            guv[0] = __fadd_rn(guv[0], t.x);
            guv[1] = __fadd_rn(guv[1], t.y);
        }
    }
    load_view_cta(Tdev, Kdev, camdev, vc, HAS_SH);
    if (vis) {
        float px, py, pz;
        transform_point<float>(vc.T, x, y, z, px, py, pz);
        float S6[6], S9[9], J[6], W[9];
        sigma_world<float>(qw, qx, qy, qz, s0, s1, s2, S6);
        sym6_to_full(S6, S9);
        proj_jacobian<float>(px, py, pz, vc.K[0], vc.K[4], J);
        W[0] = vc.T[0]; W[1] = vc.T[1]; W[2] = vc.T[2];
        W[3] = vc.T[4]; W[4] = vc.T[5]; W[5] = vc.T[6];
        W[6] = vc.T[8]; W[7] = vc.T[9]; W[8] = vc.T[10];

        const float gc[3] = {gcv[0], gcv[1], gcv[2]};
        float gS[9], gJ[6];
        conic_bwd<float>(S9, J, W, gc, gS, gJ);
        sigma_world_bwd<float>(qw, qx, qy, qz, s0, s1, s2, gS, gq, gs);
        float gp_j[3], gp_uv[3] = {0.f, 0.f, 0.f};
        proj_jacobian_bwd<float>(px, py, pz, vc.K[0], vc.K[4], gJ, gp_j);
        project_uv_bwd<float>(px, py, pz, vc.K[0], vc.K[4], guv[0], guv[1], gp_uv);
        const float gp[3] = {gp_j[0] + gp_uv[0], gp_j[1] + gp_uv[1], gp_j[2] + gp_uv[2]};
        // xyz_cam = W xyz + t  =>  grad_xyz = W^T grad_xyz_cam
#pragma unroll
        for (int k = 0; k < 3; ++k) gx[k] = W[0 + k] * gp[0] + W[3 + k] * gp[1] + W[6 + k] * gp[2];
        const float sg = sigmoid_torch(opl);
        go = gov * ((1.0f - sg) * sg);  // torch sigmoid_backward: grad * (1 - y) * y

        const float gr[3] = {grv[0], grv[1], grv[2]};
        if (HAS_SH) {
            // src/precompute_sh.cu:96-109, split into the DC column and the rest
            float dx, dy, dz, Y[N_SH];
            view_dir<float>(x, y, z, vc.cam[0], vc.cam[1], vc.cam[2], dx, dy, dz);
            sh_basis<float, N_SH>(dx, dy, dz, Y);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float g = gr[c] * GSR_RSH0;
                gdc[c] = g * Y[0];
#pragma unroll
                for (int k = 1; k < N_SH; ++k) my_sh[c * NR + (k - 1)] = g * Y[k];
            }
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) gdc[c] = gr[c];
        }
    } else if (HAS_SH) {
#pragma unroll
        for (int k = 0; k < NR3; ++k) my_sh[k] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) { s_xyz[tid * 3 + k] = gx[k]; s_scale[tid * 3 + k] = gs[k]; s_dc[tid * 3 + k] = gdc[k]; }
    *reinterpret_cast<float4*>(s_quat + tid * 4) = make_float4(gq[0], gq[1], gq[2], gq[3]);
    s_opa[tid] = go;
    if (tma) {
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            tma_store_1d(o_xyz + (size_t)i0 * 3, s_xyz, PRE_THREADS * 3 * 4);
            tma_store_1d(o_quat + (size_t)i0 * 4, s_quat, PRE_THREADS * 4 * 4);
            tma_store_1d(o_scale + (size_t)i0 * 3, s_scale, PRE_THREADS * 3 * 4);
            tma_store_1d(o_opa + (size_t)i0, s_opa, PRE_THREADS * 4);
            tma_store_1d(o_dc + (size_t)i0 * 3, s_dc, PRE_THREADS * 3 * 4);
            if (HAS_SH) tma_store_1d(o_sh + (size_t)i0 * NR3, s_sh, PRE_THREADS * NR3 * 4);
            tma_store_commit_and_wait();
        }
    } else {
        __syncthreads();
        coop_copy(o_xyz + (size_t)i0 * 3, s_xyz, cnt * 3);
        coop_copy(o_quat + (size_t)i0 * 4, s_quat, cnt * 4);
        coop_copy(o_scale + (size_t)i0 * 3, s_scale, cnt * 3);
        coop_copy(o_opa + (size_t)i0, s_opa, cnt);
        coop_copy(o_dc + (size_t)i0 * 3, s_dc, cnt * 3);
        if (HAS_SH) coop_copy(o_sh + (size_t)i0 * NR3, s_sh, cnt * NR3);
    }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace gsr

using namespace gsr;

extern "C" {

int gsr_camera_centre(const float* camera_T_world, float* centre, void* stream) {
    if (camera_T_world == nullptr || centre == nullptr) return GSR_ERR_BAD_ARG;
    k_camera_centre<<<1, 32, 0, (cudaStream_t)stream>>>(camera_T_world, centre);
    return (int)cudaGetLastError();
}

size_t gsr_preprocess_temp_bytes(int N) {
    size_t b = 0;
    cub::DeviceScan::InclusiveSum((void*)nullptr, b, (uint64_t*)nullptr, (uint64_t*)nullptr, N > 0 ? N : 1);
    return align256(b) + align256(sizeof(uint64_t) * (size_t)(N > 0 ? N : 1));
}

int gsr_preprocess_forward(int N, int n_sh_rest, const float* xyz, const float* xyz_camera_frame,
                           int cam_first, const float* quaternion,
                           const float* scale, const float* opacity_logit, const float* rgb_dc,
                           const float* sh_rest, const float* camera_T_world, const float* K,
                           const float* camera_centre, int H, int W, float near_thresh, float far_thresh,
                           float cull_mask_padding, float mh_dist, uint32_t depth_base, float* records,
                           uint32_t* depth_key,
                           uint8_t* visible, uint64_t* scan, uint64_t* tile_mask, uint32_t* tile_win, void* temp,
                           size_t temp_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (N <= 0) return GSR_OK;
    if (temp_bytes < gsr_preprocess_temp_bytes(N)) return GSR_ERR_BAD_ARG;
    if ((tile_mask == nullptr) != (tile_win == nullptr)) return GSR_ERR_BAD_ARG;
    size_t scan_bytes = 0;
    cub::DeviceScan::InclusiveSum((void*)nullptr, scan_bytes, (uint64_t*)nullptr, (uint64_t*)nullptr, N);
    uint64_t* packed = reinterpret_cast<uint64_t*>((char*)temp + align256(scan_bytes));
    const int ntx = (W + TILE - 1) / TILE, nty = (H + TILE - 1) / TILE;
    const dim3 grid((N + PRE_THREADS - 1) / PRE_THREADS), block(PRE_THREADS);
    const int use_tma = (aligned16(records) && (sh_rest == nullptr || aligned16(sh_rest))) ? 1 : 0;
#define GSR_PRE_ARGS                                                                              \
    N, xyz, xyz_camera_frame, cam_first, quaternion, scale, opacity_logit, rgb_dc, sh_rest, camera_T_world, K, \
        camera_centre, (float)W, (float)H,                                                                 \
        near_thresh, far_thresh, cull_mask_padding, mh_dist, ntx, nty, depth_base, records, depth_key, visible, \
        packed, tile_mask, tile_win, use_tma
    switch (n_sh_rest) {
        case 0: k_preprocess_fwd<1, false><<<grid, block, 0, st>>>(GSR_PRE_ARGS); break;
        case 3: k_preprocess_fwd<4, true><<<grid, block, 0, st>>>(GSR_PRE_ARGS); break;
        case 8: k_preprocess_fwd<9, true><<<grid, block, 0, st>>>(GSR_PRE_ARGS); break;
        case 15: k_preprocess_fwd<16, true><<<grid, block, 0, st>>>(GSR_PRE_ARGS); break;
        default: return GSR_ERR_UNSUPPORTED;
    }
#undef GSR_PRE_ARGS
    cudaError_t e = cub::DeviceScan::InclusiveSum(temp, scan_bytes, packed, scan, N, st);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaGetLastError();
}

int gsr_preprocess_backward(int N, int n_sh_rest, const float* xyz, const float* quaternion,
                            const float* scale, const float* opacity_logit, const float* camera_T_world,
                            const float* K, const float* camera_centre, const uint8_t* visible,
                            const float* grad_rgb,
                            const float* grad_opacity, const float* grad_uv, const float* grad_conic,
                            const float* grad_rows, int use_rows_uv,
                            const float* grad_uv_compact, const uint64_t* scan, float* g_xyz, float* g_quaternion, float* g_scale, float* g_opacity_logit,
                            float* g_rgb_dc, float* g_sh_rest, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (N <= 0) return GSR_OK;
    if (grad_uv_compact != nullptr && scan == nullptr) return GSR_ERR_BAD_ARG;
    const dim3 grid((N + PRE_THREADS - 1) / PRE_THREADS), block(PRE_THREADS);
    const int use_tma = (aligned16(g_xyz) && aligned16(g_quaternion) && aligned16(g_scale) &&
                         aligned16(g_opacity_logit) && aligned16(g_rgb_dc) &&
                         (g_sh_rest == nullptr || aligned16(g_sh_rest)))
                            ? 1
                            : 0;
#define GSR_PRE_ARGS                                                                                   \
    N, xyz, quaternion, scale, opacity_logit, camera_T_world, K, camera_centre, visible, grad_rgb,         \
        grad_opacity, grad_uv,                                                                             \
        grad_conic, grad_rows, use_rows_uv, grad_uv_compact, scan, g_xyz, g_quaternion, g_scale, g_opacity_logit, g_rgb_dc, g_sh_rest, use_tma
    switch (n_sh_rest) {
        case 0: k_preprocess_bwd<1, false><<<grid, block, 0, st>>>(GSR_PRE_ARGS); break;
        case 3: k_preprocess_bwd<4, true><<<grid, block, 0, st>>>(GSR_PRE_ARGS); break;
        case 8: k_preprocess_bwd<9, true><<<grid, block, 0, st>>>(GSR_PRE_ARGS); break;
        case 15: k_preprocess_bwd<16, true><<<grid, block, 0, st>>>(GSR_PRE_ARGS); break;
        default: return GSR_ERR_UNSUPPORTED;
    }
#undef GSR_PRE_ARGS
    return (int)cudaGetLastError();
}

}  // extern "C"
