/* gsr_oracle.c — CPU oracle for the Gaussian-splat rasterizer hot path.
 *
 * TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu baseline
 * through oracle/cpu_oracle.py.  The product (gaussian_splatting_b200/) never links or calls this.
 *
 * Plain-C restatement of the reference's algorithm (joeyan/gaussian_splatting @ ae0d717):
 *   per-gaussian chain   splat_py/rasterize.py:29-93, src/projection.cu, src/precompute_sh.cu
 *   tile lists           src/tile_culling.cu:8-340
 *   tile renderer        src/render.cu:101-188, src/render_backward.cu:120-284
 *   per-gaussian VJP     src/projection_backward.cu
 * Pinning: tests/test_oracle_golden.py checks it against the reference's own known-answer tests
 * (test/test_projection.py, test_tile_culling.py, test_rasterize.py) and against fixtures produced by
 * the compiled reference on a B200 (tests/golden/, generator: tools/make_golden.py).
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -mfma -shared -fPIC gsr_oracle.c -o _build/libgsr_oracle.so -lm
 * (-ffp-contract=off: only the explicit fma() calls fuse, like the pinned CUDA code.)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SH0F 0.28209479177387814f
#define RSH0F 3.544907701811032f
#define SH1F 0.4886025119029199f
#define SH20F 1.0925484305920792f
#define SH22F 0.31539156525252005f
#define SH24F 0.5462742152960396f
#define SH30F 0.5900435899266435f
#define SH31F 2.890611442640554f
#define SH32F 0.4570457994644658f
#define SH33F 0.263875515352797f
#define SH35F 1.445305721320277f

/* ---------------- fp32 instantiation (production branch of the reference) ---------------- */
#define REAL float
#define SUF(x) x##_f32
#define FAST 1
#define FMA(a, b, c) fmaf((a), (b), (c))
#define SQRT(a) sqrtf(a)
#define EXP(a) expf(a)
/* __expf(x) = ex2.approx(x * log2(e)) (src/render.cu:138) */
#define FASTEXP(a) exp2f((a) * 1.4426950216293334961f)
#include "gsr_oracle_body.inc"
#undef REAL
#undef SUF
#undef FAST
#undef FMA
#undef SQRT
#undef EXP
#undef FASTEXP

/* ---------------- fp64 instantiation (the reference's gradcheck branch) ---------------- */
#define REAL double
#define SUF(x) x##_f64
#define FAST 0
#define FMA(a, b, c) fma((a), (b), (c))
#define SQRT(a) sqrt(a)
#define EXP(a) exp(a)
#define FASTEXP(a) exp(a)
#include "gsr_oracle_body.inc"
#undef REAL
#undef SUF
#undef FAST
#undef FMA
#undef SQRT
#undef EXP
#undef FASTEXP

/* ---------------- tile lists (fp32 only, like the reference) ---------------- */
typedef struct {
    float c[8];
    int radius_tiles;
} obb_t;

/* src/tile_culling.cu:69-122 */
static void compute_obb(float u, float v, float c0, float c1, float c2, float mh, obb_t* o) {
    const float a = c0 + 0.25f, b = c1 * 0.5f, c = c2 + 0.25f;
    const float d = a - c;
    const float right = sqrtf(fmaf(b, b, (d * d) * 0.25f));
    const float s = a + c;
    const float l1 = fmaf(s, 0.5f, right), l2 = fmaf(s, 0.5f, -right);
    const float r_major = sqrtf(l1) * mh, r_minor = sqrtf(l2) * mh;
    float theta;
    if ((double)fabsf(b) < 1e-16) theta = (a >= c) ? 0.0f : (float)(M_PI / 2);
    else theta = atan2f(l1 - a, b);
    const float ct = cosf(theta), st = sinf(theta);
    const float t_mc = ct * r_minor, t_ms = r_minor * st;
    o->c[0] = u + fmaf(ct, -r_major, t_ms);
    o->c[1] = v + fmaf(-r_major, st, -t_mc);
    o->c[2] = u + fmaf(ct, r_major, t_ms);
    o->c[3] = v + fmaf(r_major, st, -t_mc);
    o->c[4] = u + fmaf(ct, -r_major, -t_ms);
    o->c[5] = v + fmaf(-r_major, st, t_mc);
    o->c[6] = u + fmaf(ct, r_major, -t_ms);
    o->c[7] = v + fmaf(r_major, st, t_mc);
    o->radius_tiles = (int)(ceilf(r_major * 0.0625f) + 1.0f);
}

static float min4(float a, float b, float c, float d) { return fminf(fminf(a, b), fminf(c, d)); }
static float max4(float a, float b, float c, float d) { return fmaxf(fmaxf(a, b), fmaxf(c, d)); }

/* src/tile_culling.cu:8-66; tb = [left, right, top, bottom] */
static int sat_overlap(const float* obb, const float* tb) {
    if (min4(obb[0], obb[2], obb[4], obb[6]) > tb[1] || max4(obb[0], obb[2], obb[4], obb[6]) < tb[0]) return 0;
    if (min4(obb[1], obb[3], obb[5], obb[7]) > tb[3] || max4(obb[1], obb[3], obb[5], obb[7]) < tb[2]) return 0;
    for (int axis = 0; axis < 2; ++axis) {
        const float ax = axis == 0 ? obb[2] - obb[0] : obb[2] - obb[6];
        const float ay = axis == 0 ? obb[3] - obb[1] : obb[3] - obb[7];
        const float xl = ax * tb[0], xr = ax * tb[1];
        const float tl = fmaf(ay, tb[2], xl), tr = fmaf(ay, tb[2], xr);
        const float bl = fmaf(ay, tb[3], xl), br = fmaf(ay, tb[3], xr);
        const float p0 = fmaf(obb[2], ax, obb[3] * ay);
        const float p1 = axis == 0 ? fmaf(obb[0], ax, obb[1] * ay) : fmaf(obb[6], ax, obb[7] * ay);
        if (min4(tl, tr, bl, br) > fmaxf(p0, p1) || max4(tl, tr, bl, br) < fminf(p0, p1)) return 0;
    }
    return 1;
}

typedef struct {
    int tile;
    float z;
    int idx;
} pair_t;

static int pair_cmp(const void* pa, const void* pb) {
    const pair_t* a = (const pair_t*)pa;
    const pair_t* b = (const pair_t*)pb;
    if (a->tile != b->tile) return a->tile < b->tile ? -1 : 1;
    if (a->z != b->z) return a->z < b->z ? -1 : 1;
    return a->idx < b->idx ? -1 : (a->idx > b->idx ? 1 : 0);
}

/* src/tile_culling.cu:124-340: visits the tile window of every gaussian, keeps (gaussian, tile) pairs
 * that pass the SAT test, orders them by (tile, camera z, gaussian index).
 * Call with sorted_idx == NULL to get the pair count; then with a buffer of that size.
 * ranges: [ntx*nty + 1].  Returns P. */
int orc_tile_lists(int N, const float* uvs, const float* xyz_cam, const float* conic, int ntx, int nty, float mh,
                   int* sorted_idx, int* ranges) {
    size_t cap = 1024, P = 0;
    pair_t* pairs = (pair_t*)malloc(cap * sizeof(pair_t));
    for (int i = 0; i < N; ++i) {
        const float u = uvs[i * 2], v = uvs[i * 2 + 1];
        obb_t o;
        compute_obb(u, v, conic[i * 3], conic[i * 3 + 1], conic[i * 3 + 2], mh, &o);
        const int ptx = (int)floorf(u * 0.0625f), pty = (int)floorf(v * 0.0625f);
        const int x0 = (int)fmaxf(0.0f, (float)(ptx - o.radius_tiles));
        const int x1 = (int)fminf((float)ntx, (float)(ptx + o.radius_tiles));
        const int y0 = (int)fmaxf(0.0f, (float)(pty - o.radius_tiles));
        const int y1 = (int)fminf((float)nty, (float)(pty + o.radius_tiles));
        for (int tx = x0; tx < x1; ++tx)
            for (int ty = y0; ty < y1; ++ty) {
                const float tb[4] = {(float)tx * 16.0f, (float)(tx + 1) * 16.0f, (float)ty * 16.0f,
                                     (float)(ty + 1) * 16.0f};
                if (!sat_overlap(o.c, tb)) continue;
                if (P == cap) {
                    cap *= 2;
                    pairs = (pair_t*)realloc(pairs, cap * sizeof(pair_t));
                }
                pairs[P].tile = ty * ntx + tx;
                pairs[P].z = xyz_cam[i * 3 + 2];
                pairs[P].idx = i;
                ++P;
            }
    }
    if (sorted_idx != NULL) {
        qsort(pairs, P, sizeof(pair_t), pair_cmp);
        const int n_tiles = ntx * nty;
        memset(ranges, 0, sizeof(int) * (size_t)(n_tiles + 1));
        for (size_t p = 0; p < P; ++p) {
            sorted_idx[p] = pairs[p].idx;
            ranges[pairs[p].tile + 1] += 1;
        }
        for (int t = 0; t < n_tiles; ++t) ranges[t + 1] += ranges[t];
    }
    free(pairs);
    return (int)P;
}

/* camera centre = inverse(T)[:3,3] by Gauss-Jordan in double */
void orc_camera_centre(const double* T, double* cam) {
    double A[4][5];
    for (int r = 0; r < 4; ++r) {
        for (int c = 0; c < 4; ++c) A[r][c] = T[r * 4 + c];
        A[r][4] = (r == 3) ? 1.0 : 0.0;
    }
    for (int col = 0; col < 4; ++col) {
        int piv = col;
        for (int r = col + 1; r < 4; ++r)
            if (fabs(A[r][col]) > fabs(A[piv][col])) piv = r;
        for (int c = 0; c < 5; ++c) { const double t = A[col][c]; A[col][c] = A[piv][c]; A[piv][c] = t; }
        const double inv = 1.0 / A[col][col];
        for (int c = 0; c < 5; ++c) A[col][c] *= inv;
        for (int r = 0; r < 4; ++r) {
            if (r == col) continue;
            const double f = A[r][col];
            for (int c = 0; c < 5; ++c) A[r][c] -= f * A[col][c];
        }
    }
    cam[0] = A[0][4]; cam[1] = A[1][4]; cam[2] = A[2][4];
}
