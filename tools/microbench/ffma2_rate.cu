// Issue-rate probe for packed fp32 (fma.rn.f32x2 -> FFMA2) on sm_100a, and a check that explicit-rounding
// packed mul + add are NOT contracted into one fused operation.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu && ./ffma2_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// explicit-FMA forms that cannot be contracted: a*b = fma(a,b,-0), a+b = fma(a,1,b) (both exact identities)
__device__ __forceinline__ u64 mul2_safe(u64 a, u64 b) { return fma2(a, b, pack(-0.0f, -0.0f)); }
__device__ __forceinline__ u64 add2_safe(u64 a, u64 b) { return fma2(a, pack(1.0f, 1.0f), b); }

constexpr int CH = 8;  // independent chains per thread

template <int MODE>
__global__ void __launch_bounds__(256) k_rate(float* out, int iters, float m, float c) {
    float x[CH * 2];
    uint32_t z[CH];
#pragma unroll
    for (int i = 0; i < CH * 2; ++i) x[i] = (float)(threadIdx.x + i) * 1e-3f;
#pragma unroll
    for (int i = 0; i < CH; ++i) z[i] = threadIdx.x * 2654435761u + i;
    u64 p[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) p[i] = pack(x[2 * i], x[2 * i + 1]);
    const u64 pm = pack(m, m), pc = pack(c, c);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 2) {  // scalar: 2*CH FFMA
#pragma unroll
            for (int i = 0; i < CH * 2; ++i) x[i] = __fmaf_rn(x[i], m, c);
        }
        if (MODE == 5) {  // scalar: CH FFMA (same instruction count as packed)
#pragma unroll
            for (int i = 0; i < CH; ++i) x[i] = __fmaf_rn(x[i], m, c);
        }
        if (MODE == 1 || MODE == 3) {  // packed: CH FFMA2 (same flops as mode 0)
#pragma unroll
            for (int i = 0; i < CH; ++i) p[i] = fma2(p[i], pm, pc);
        }
        if (MODE == 2 || MODE == 3 || MODE == 4) {  // + CH integer ALU ops (LOP3)
#pragma unroll
            for (int i = 0; i < CH; ++i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(z[i]) : "r"(z[(i + 1) % CH]), "r"((uint32_t)it));
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < CH * 2; ++i) acc += x[i];
#pragma unroll
    for (int i = 0; i < CH; ++i) { float a, b; unpack(p[i], a, b); acc += a + b + __uint_as_float(z[i] & 0x3fffffffu); }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__global__ void k_round(const float* a, const float* b, const float* c, float* packed, float* scalar, float* fused, float* safe, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (2 * i + 1 >= n) return;
    const u64 pa = pack(a[2 * i], a[2 * i + 1]), pb = pack(b[2 * i], b[2 * i + 1]), pcc = pack(c[2 * i], c[2 * i + 1]);
    float r0, r1;
    unpack(add2(mul2(pa, pb), pcc), r0, r1);
    packed[2 * i] = r0; packed[2 * i + 1] = r1;
    unpack(add2_safe(mul2_safe(pa, pb), pcc), r0, r1);
    safe[2 * i] = r0; safe[2 * i + 1] = r1;
    scalar[2 * i] = __fadd_rn(__fmul_rn(a[2 * i], b[2 * i]), c[2 * i]);
    scalar[2 * i + 1] = __fadd_rn(__fmul_rn(a[2 * i + 1], b[2 * i + 1]), c[2 * i + 1]);
    fused[2 * i] = __fmaf_rn(a[2 * i], b[2 * i], c[2 * i]);
    fused[2 * i + 1] = __fmaf_rn(a[2 * i + 1], b[2 * i + 1], c[2 * i + 1]);
}

template <int MODE>
float run(float* out, int iters, int blocks) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_rate<MODE><<<blocks, 256>>>(out, iters, 0.999f, 1e-3f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k_rate<MODE><<<blocks, 256>>>(out, iters, 0.999f, 1e-3f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int blocks = prop.multiProcessorCount * 8, iters = 20000;
    float* out; cudaMalloc(&out, blocks * 256 * sizeof(float));
    const char* names[] = {"16 FFMA", "8 FFMA2 (same flops)", "16 FFMA + 8 LOP3", "8 FFMA2 + 8 LOP3", "8 LOP3", "8 FFMA"};
    float ms[6] = {run<0>(out, iters, blocks), run<1>(out, iters, blocks), run<2>(out, iters, blocks), run<3>(out, iters, blocks), run<4>(out, iters, blocks), run<5>(out, iters, blocks)};
    const double warps = (double)blocks * 8, clk = prop.clockRate * 1e3;
    printf("{\"device\": \"%s\", \"sms\": %d, \"blocks\": %d, \"iters\": %d,\n \"rows\": [\n", prop.name, prop.multiProcessorCount, blocks, iters);
    for (int i = 0; i < 6; ++i) {
        // issue cycles per loop iteration per SM sub-partition, assuming the nominal boost clock
        const double cyc = ms[i] * 1e-3 * clk / ((double)iters * warps / (prop.multiProcessorCount * 4));
        printf("  {\"loop_body\": \"%s\", \"ms\": %.3f, \"smsp_cycles_per_warp_iteration\": %.2f}%s\n", names[i], ms[i], cyc, i < 5 ? "," : "");
    }
    // rounding check
    const int n = 1 << 20;
    float *a, *b, *c, *pk, *sc, *fu, *sf;
    cudaMallocManaged(&a, n * 4); cudaMallocManaged(&b, n * 4); cudaMallocManaged(&c, n * 4);
    cudaMallocManaged(&pk, n * 4); cudaMallocManaged(&sc, n * 4); cudaMallocManaged(&fu, n * 4); cudaMallocManaged(&sf, n * 4);
    uint32_t s = 12345u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)((s >> 8) & 0xffffff) / 16777216.0f * 4.0f - 2.0f; };
    for (int i = 0; i < n; ++i) { a[i] = rnd(); b[i] = rnd(); c[i] = -a[i] * b[i] * (1.0f + rnd() * 1e-6f); }
    k_round<<<n / 2 / 256, 256>>>(a, b, c, pk, sc, fu, sf, n);
    cudaDeviceSynchronize();
    int ne_scalar = 0, ne_fused = 0, fused_differs = 0, ne_safe = 0;
    for (int i = 0; i < n; ++i) {
        ne_scalar += (pk[i] != sc[i]); ne_fused += (pk[i] != fu[i]); fused_differs += (sc[i] != fu[i]); ne_safe += (sf[i] != sc[i]);
    }
    printf(" ],\n \"rounding\": {\"n\": %d, \"packed_mul_add_ne_scalar_mul_add\": %d, \"packed_mul_add_ne_fma\": %d, \"scalar_mul_add_ne_fma\": %d, \"explicit_fma_forms_ne_scalar_mul_add\": %d}}\n",
           n, ne_scalar, ne_fused, fused_differs, ne_safe);
    return 0;
}
