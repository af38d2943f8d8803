// cpab_sample.cuh -- device helpers of the linear / bilinear / trilinear sampler, shared by the
// stand-alone interpolate kernels (cpab_interp.cu) and the fused transform_data kernels
// (cpab_integrate.cu).  Arithmetic follows libcpab/pytorch/interpolation.py:18-172 operation by
// operation (every product and sum rounded separately), so that results are bit-identical to the
// reference's float32 output for the same grid.
#pragma once

#include "cpab_common.cuh"
#include "cpab_f32x2.cuh"

namespace cpab {

template <typename T> struct R;
template <> struct R<float> {
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float flo(float a) { return floorf(a); }
};
template <> struct R<double> {
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double flo(double a) { return floor(a); }
};

struct Shape {
    int N, C;
    int S[3];   // input spatial sizes  (W, H, D)
    int O[3];   // output spatial sizes (W_o, H_o, D_o)
};

// scale, floor, +1, clamp, weight (interpolation.py:29-47): returns clamped taps and xd
template <typename T>
__device__ __forceinline__ void taps(T gcoord, int size, int& t0, int& t1, T& wgt)
{
    const T hi = (T)(size - 1);
    const T x = R<T>::mul(gcoord, hi);
    const T f = R<T>::flo(x);
    const T f0 = fmin(fmax(f, (T)0), hi);
    const T f1 = fmin(fmax(f + (T)1, (T)0), hi);
    t0 = (int)f0;
    t1 = (int)f1;
    wgt = R<T>::sub(x, f0);
}

// float32: the same values without a conversion-pipe instruction (FRND + 2 F2I per coordinate run
// at 16 lanes/clk/SM on a B200, profiles/r02_pipe_probe.txt).  For x in [-1, 2^22) the sum
// x + 1.5*2^23 rounded DOWN is exactly 1.5*2^23 + floor(x) (ulp 1 in [2^23, 2^24)), so the integer
// floor is a difference of bit patterns; x below -1 (and NaN: fmaxf returns the other operand) is
// lifted to -1, which has the same clamped taps (0, 0); x >= 2^22 yields an integer >= 2^22 > size-1,
// clamped to size-1 like the reference's floor.  The clamped lower tap goes back to float exactly,
// and the weight is formed from the unclamped x as the reference does.
// Requires size <= 2^22 (kMaxInterpExtentF32; the launchers refuse larger float32 extents).
// `t1 - t0` is 0 or 1.
template <>
__device__ __forceinline__ void taps<float>(float gcoord, int size, int& t0, int& t1, float& wgt)
{
    const int ihi = size - 1;
    const float hi = (float)ihi;
    const float x = __fmul_rn(gcoord, hi);
    const float magic = 12582912.0f;                        // 1.5 * 2^23 = 0x4B400000
    const float t = __fadd_rd(fmaxf(x, -1.0f), magic);
    const int i = __float_as_int(t) - 0x4B400000;           // floor(x), any int when |x| is huge
    t0 = min(max(i, 0), ihi);
    t1 = t0 + ((unsigned)i < (unsigned)ihi ? 1 : 0);        // clamp(i + 1) differs from t0 iff 0 <= i < size-1
    wgt = __fsub_rn(x, __int2float_rn(t0));                 // (0 <= t0 < 2^22: exact; I2FP runs on the ALU pipe)
}

// multilinear blend of 2^NDIM corner values, x first (bit 0), then y, then z
template <int NDIM, typename T>
__device__ __forceinline__ T blend(const T* v, const T* w)
{
    T a[1 << NDIM];
#pragma unroll
    for (int i = 0; i < (1 << NDIM); ++i) a[i] = v[i];
#pragma unroll
    for (int j = 0; j < NDIM; ++j) {
        const T om = R<T>::sub((T)1, w[j]);
#pragma unroll
        for (int m = 0; m < (1 << (NDIM - 1 - j)); ++m)
            a[m] = R<T>::add(R<T>::mul(a[2 * m], om), R<T>::mul(a[2 * m + 1], w[j]));
    }
    return a[0];
}

// float32, NDIM >= 2: the same separately rounded products and sums with the products packed
// (FMUL2, cpab_f32x2.cuh): two outputs of a level share their weights, so (a[2m], a[2m+2]) * (1-w)
// and (a[2m+1], a[2m+3]) * w are one instruction each; the last level multiplies (a0, a1) by
// (1-w, w).  Sums stay scalar (ptxas would contract a packed product feeding a packed add into an
// FFMA2).  Bit-identical to the scalar form; 8 instead of 11 instructions in 2-D, 17 instead of 24
// in 3-D.
template <int NDIM>
__device__ __forceinline__ float blend_packed(const float* v, const float* w)
{
    static_assert(NDIM >= 2, "packed blend needs at least two levels");
    float a[1 << NDIM];
#pragma unroll
    for (int i = 0; i < (1 << NDIM); ++i) a[i] = v[i];
#pragma unroll
    for (int j = 0; j < NDIM - 1; ++j) {
        const float om = __fsub_rn(1.0f, w[j]);
#pragma unroll
        for (int m = 0; m < (1 << (NDIM - 1 - j)); m += 2) {
            const F2 P = mul2(pk(a[2 * m], a[2 * m + 2]), bc(om));
            const F2 Q = mul2(pk(a[2 * m + 1], a[2 * m + 3]), bc(w[j]));
            a[m] = __fadd_rn(lo(P), lo(Q));
            a[m + 1] = __fadd_rn(hi(P), hi(Q));
        }
    }
    const F2 Rr = mul2(pk(a[0], a[1]), pk(__fsub_rn(1.0f, w[NDIM - 1]), w[NDIM - 1]));
    return __fadd_rn(lo(Rr), hi(Rr));
}
template <> __device__ __forceinline__ float blend<2, float>(const float* v, const float* w) { return blend_packed<2>(v, w); }
template <> __device__ __forceinline__ float blend<3, float>(const float* v, const float* w) { return blend_packed<3>(v, w); }

// reverse of blend: corner weights gv[] (d out / d v) and dw[] (d out / d w_j), scaled by g
template <int NDIM, typename T>
__device__ __forceinline__ void blend_vjp(const T* v, const T* w, T g, T* gv, T* dw)
{
    // forward levels
    T lev[NDIM + 1][1 << NDIM];
#pragma unroll
    for (int i = 0; i < (1 << NDIM); ++i) lev[0][i] = v[i];
#pragma unroll
    for (int j = 0; j < NDIM; ++j)
#pragma unroll
        for (int m = 0; m < (1 << (NDIM - 1 - j)); ++m)
            lev[j + 1][m] = lev[j][2 * m] * ((T)1 - w[j]) + lev[j][2 * m + 1] * w[j];
    T gl[1 << NDIM];
    gl[0] = g;
#pragma unroll
    for (int j = NDIM - 1; j >= 0; --j) {
        T acc = 0;
#pragma unroll
        for (int m = (1 << (NDIM - 1 - j)) - 1; m >= 0; --m) {
            const T gm = gl[m];
            acc += gm * (lev[j][2 * m + 1] - lev[j][2 * m]);
            gl[2 * m + 1] = gm * w[j];
            gl[2 * m] = gm * ((T)1 - w[j]);
        }
        dw[j] = acc;
    }
#pragma unroll
    for (int i = 0; i < (1 << NDIM); ++i) gv[i] = gl[i];
}

constexpr int kMaxInterpExtentF32 = 1 << 22;
constexpr int TILE = 32;
constexpr int REPS = TILE / 8;

// Per-point sampling state shared by forward and backward: for every corner of the leading
// NDIM-1 dimensions the 32-bit element offset of the tap pair along the LAST (memory-fastest)
// dimension, whether that pair really is two texels (it collapses to one at a clamped border),
// and the interpolation weights.
template <typename T, int NDIM> struct Taps {
    int base[1 << (NDIM - 1)];   // offset of (x?,y?,.., last = t0)
    bool two;                    // t1 != t0 along the last dimension
    T w[NDIM];
};

template <typename T, int NDIM>
__device__ __forceinline__ Taps<T, NDIM> make_taps(const T* gcoord, const Shape& s)
{
    Taps<T, NDIM> tp;
    int t0[NDIM], t1[NDIM];
#pragma unroll
    for (int j = 0; j < NDIM; ++j) taps(gcoord[j], s.S[j], t0[j], t1[j], tp.w[j]);
    tp.two = t1[NDIM - 1] != t0[NDIM - 1];
    // t1 - t0 is 0 or 1 per dimension: neighbouring rows / slabs are an optional stride away
    if (NDIM == 1) {
        tp.base[0] = t0[0];
    } else if (NDIM == 2) {
        tp.base[0] = t0[0] * s.S[1] + t0[1];
        tp.base[1] = tp.base[0] + (t1[0] != t0[0] ? s.S[1] : 0);
    } else {
        const int r = (t0[0] * s.S[1] + t0[1]) * s.S[2] + t0[2];
        const int sx = t1[0] != t0[0] ? s.S[1] * s.S[2] : 0;
        const int sy = t1[1] != t0[1] ? s.S[2] : 0;
        tp.base[0] = r;
        tp.base[1] = r + sx;
        tp.base[2] = r + sy;
        tp.base[3] = r + sx + sy;
    }
    return tp;
}

// gather the 2^NDIM corner values of one channel plane (corner bit j <-> dimension j)
template <typename T, int NDIM>
__device__ __forceinline__ void gather(const T* __restrict__ dp, const Taps<T, NDIM>& tp, T* v)
{
    constexpr int H = 1 << (NDIM - 1);
#pragma unroll
    for (int u = 0; u < H; ++u) {
        const T* q = dp + tp.base[u];
        const T lo = __ldg(q);
        v[u] = lo;
        v[u + H] = tp.two ? __ldg(q + 1) : lo;
    }
}


}  // namespace cpab
