// cpab_f32x2.cuh -- Blackwell packed-FP32 arithmetic (FFMA2 / FADD2 / FMUL2, sm_100+).
//
// One instruction performs the same IEEE-754 binary32 operation on both halves of a 64-bit
// register pair; a 32-bit operand is broadcast for free (`R.F32` operand form).  The integration
// kernels of this library are bound by instruction issue, not by the FMA pipe, and a 2-D point is a
// natural pair, so the per-axis arithmetic of the cell search and the affine updates issue in half
// the slots.  Per lane the results are bit-identical to the scalar instructions.
//
// NOTE: ptxas contracts `mul.rn.f32x2` + `add.rn.f32x2` into one FFMA2 even though both carry an
// explicit rounding mode (CUDA 12.9, also with -fmad=false).  Code that needs separately rounded
// products and sums (the bit-exact "strict" forward) must not feed a packed product into a packed
// add.
#pragma once
#if defined(__CUDACC__)

namespace cpab {

struct F2 { unsigned long long v; };

__device__ __forceinline__ F2 pk(float lo, float hi)
{
    F2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ F2 bc(float x) { return pk(x, x); }            // becomes an R.F32 operand
__device__ __forceinline__ void unpk(F2 a, float& lo, float& hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v));
}
__device__ __forceinline__ float lo(F2 a) { float l, h; unpk(a, l, h); return l; }
__device__ __forceinline__ float hi(F2 a) { float l, h; unpk(a, l, h); return h; }

__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c)
{
    F2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return r;
}
__device__ __forceinline__ F2 fma2_rm(F2 a, F2 b, F2 c)                    // both lanes rounded down
{
    F2 r;
    asm("fma.rm.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return r;
}
__device__ __forceinline__ F2 add2(F2 a, F2 b)
{
    F2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ F2 sub2(F2 a, F2 b)
{
    F2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ F2 mul2(F2 a, F2 b)
{
    F2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}

}  // namespace cpab
#endif
