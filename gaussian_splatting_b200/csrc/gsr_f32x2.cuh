// gsr_f32x2.cuh — packed fp32 pairs (Blackwell FFMA2 / FMUL2 / FADD2).
//
// sm_100 executes `fma/mul/add.rn.f32x2` on a 64-bit register pair as ONE issued instruction that performs
// two independent IEEE-754 round-to-nearest operations (measured on B200, profiles/r01_microbench_ffma2.json:
// an FFMA2 occupies the FMA pipe for the same two cycles as two FFMAs but takes one issue slot, and a scalar
// register can be broadcast to both halves for free).  The tile renderers are issue-bound, so a lane carries
// TWO pixels and the per-pixel arithmetic is issued once per pair; every half is still the correctly rounded
// scalar operation, which is what the bitwise parity with the reference rests on.
//
// CONTRACTION HAZARD (ptxas 12.9): `mul.rn.f32x2` followed by `add.rn.f32x2` on its result IS fused into one
// FFMA2 — unlike the scalar `.rn` forms — and `-fmad=false` does not stop it.  Wherever the rounding contract
// needs an unfused product followed by a sum, the sum goes through add2_unfused() (two scalar add.rn.f32).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gsr {

struct F2 {
    unsigned long long v;  // lo = bits 0..31 (first pixel), hi = bits 32..63 (second pixel)
};

__device__ __forceinline__ F2 pk(float lo, float hi) {
    F2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ F2 bc(float x) { return pk(x, x); }
__device__ __forceinline__ float lo(F2 a) {
    float l, h;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(l), "=f"(h) : "l"(a.v));
    (void)h;
    return l;
}
__device__ __forceinline__ float hi(F2 a) {
    float l, h;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(l), "=f"(h) : "l"(a.v));
    (void)l;
    return h;
}
__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c) {
    F2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return r;
}
__device__ __forceinline__ F2 mul2(F2 a, F2 b) {
    F2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
// a + b where neither operand is the result of a mul2 (see the contraction hazard above)
__device__ __forceinline__ F2 add2(F2 a, F2 b) {
    F2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ F2 add2_unfused(F2 a, F2 b) {
    return pk(__fadd_rn(lo(a), lo(b)), __fadd_rn(hi(a), hi(b)));
}
// b - a, one rounding per half: a * (-1) is exact
__device__ __forceinline__ F2 rsub2(F2 a, F2 b) { return fma2(a, bc(-1.0f), b); }

}  // namespace gsr
