// cpab_interp.cu -- linear / bilinear / trilinear resampling and its backward (sm_100a).
//
// Replaces libcpab/pytorch/interpolation.py:18-172 of the reference (2^ndim advanced-index
// gathers that materialise ~10 [N*nP, C] temporaries, a CPU-built arange per call, and an
// autograd-derived backward) with one fused kernel per direction.
//
// Layout problem and how it is solved: grid points are ordered with the FIRST coordinate fastest
// (p = ix + W_o*(iy + H_o*iz), libcpab/pytorch/functions.py:102-108) while both the input
// [N,C,W,H(,D)] and the output [N,C,W_o,H_o(,D_o)] have the LAST spatial index fastest
// (interpolation.py:104-105,170-171).  A CTA therefore owns a 32 x 32 tile spanning the first and
// the last output index: the grid tile is loaded unit-stride along ix into shared memory, the
// threads then re-map so that a warp runs along the last index -- texel gathers and output stores
// become unit-stride for near-identity warps -- and for the backward the d/dgrid tile goes back
// through shared memory to be stored unit-stride along ix.  HBM traffic per output point is the
// algorithmic 4*ndim (grid) + 4C (texels, each fetched once per tile through L1/L2) + 4C (store).
//
// Arithmetic: every product and sum is rounded separately, in the reference's order
// (x*(size-1); floor; +1; clamp; xd = x - x0; c00*(1-xd) + c10*xd; ...), so the forward output is
// bit-identical to the reference's float32 result for the same grid.
#include "cpab_common.cuh"

namespace cpab {

namespace {

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

// offset of texel (t[0],t[1],t[2]) inside one [S0,S1,S2] channel plane
template <int NDIM>
__device__ __forceinline__ size_t texel(const int* t, const Shape& s)
{
    size_t o = t[0];
    if (NDIM >= 2) o = o * s.S[1] + t[1];
    if (NDIM >= 3) o = o * s.S[2] + t[2];
    return o;
}

constexpr int TILE = 32;

// ---------------------------------------------------------------------------------------------
// forward.  NDIM >= 2: CTA = 256 threads, tile 32 (first index) x 32 (last index);
// blockIdx.x -> tile along the first index, blockIdx.y -> tile along the last index,
// blockIdx.z -> n * (middle extent) + middle index.
// ---------------------------------------------------------------------------------------------
template <typename T, int NDIM>
__global__ void __launch_bounds__(256)
k_interp_fwd(const T* __restrict__ data, const T* __restrict__ grid, T* __restrict__ out, Shape s)
{
    __shared__ T sg[NDIM][TILE][TILE + 1];
    const int mid = NDIM == 3 ? s.O[1] : 1;
    const int n = blockIdx.z / mid, im = blockIdx.z - n * mid;
    const int a0 = blockIdx.x * TILE, f0 = blockIdx.y * TILE;
    const int OF = s.O[NDIM - 1];
    const long nP = (long)s.O[0] * (NDIM >= 2 ? s.O[1] : 1) * (NDIM >= 3 ? s.O[2] : 1);
    const T* gn = grid + (size_t)n * NDIM * nP;

    {   // phase A: unit stride along the first index
        const int a = threadIdx.x & 31, b = threadIdx.x >> 5;
#pragma unroll
        for (int rep = 0; rep < TILE / 8; ++rep) {
            const int fo = b + 8 * rep;
            const int i0 = a0 + a, iF = f0 + fo;
            if (i0 < s.O[0] && iF < OF) {
                const long p = NDIM == 2 ? (long)i0 + (long)s.O[0] * iF
                                         : (long)i0 + (long)s.O[0] * (im + (long)s.O[1] * iF);
#pragma unroll
                for (int j = 0; j < NDIM; ++j) sg[j][fo][a] = gn[(size_t)j * nP + p];
            }
        }
    }
    __syncthreads();
    {   // phase B: a warp runs along the last index
        const int fo = threadIdx.x & 31, b = threadIdx.x >> 5;
        const int iF = f0 + fo;
        const size_t plane = (size_t)s.S[0] * (NDIM >= 2 ? s.S[1] : 1) * (NDIM >= 3 ? s.S[2] : 1);
        const size_t oplane = (size_t)nP;
#pragma unroll
        for (int rep = 0; rep < TILE / 8; ++rep) {
            const int a = b + 8 * rep;
            const int i0 = a0 + a;
            if (i0 >= s.O[0] || iF >= OF) continue;
            int t0[NDIM], t1[NDIM];
            T w[NDIM];
#pragma unroll
            for (int j = 0; j < NDIM; ++j) taps(sg[j][fo][a], s.S[j], t0[j], t1[j], w[j]);
            size_t off[1 << NDIM];
#pragma unroll
            for (int cn = 0; cn < (1 << NDIM); ++cn) {
                int t[3] = {0, 0, 0};
#pragma unroll
                for (int j = 0; j < NDIM; ++j) t[j] = ((cn >> j) & 1) ? t1[j] : t0[j];
                off[cn] = texel<NDIM>(t, s);
            }
            const size_t oidx = NDIM == 2 ? (size_t)i0 * s.O[1] + iF
                                          : ((size_t)i0 * s.O[1] + im) * s.O[2] + iF;
            for (int c = 0; c < s.C; ++c) {
                const T* dp = data + ((size_t)n * s.C + c) * plane;
                T v[1 << NDIM];
#pragma unroll
                for (int cn = 0; cn < (1 << NDIM); ++cn) v[cn] = __ldg(dp + off[cn]);
                out[((size_t)n * s.C + c) * oplane + oidx] = blend<NDIM>(v, w);
            }
        }
    }
}

// 1-D: no transposition needed
template <typename T>
__global__ void __launch_bounds__(256)
k_interp_fwd_1d(const T* __restrict__ data, const T* __restrict__ grid, T* __restrict__ out, Shape s)
{
    const int n = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.O[0]) return;
    int t0, t1;
    T w[1];
    taps(grid[(size_t)n * s.O[0] + i], s.S[0], t0, t1, w[0]);
    for (int c = 0; c < s.C; ++c) {
        const T* dp = data + ((size_t)n * s.C + c) * s.S[0];
        T v[2] = {__ldg(dp + t0), __ldg(dp + t1)};
        out[((size_t)n * s.C + c) * s.O[0] + i] = blend<1>(v, w);
    }
}

// ---------------------------------------------------------------------------------------------
// backward: dgrid [N,NDIM,nP] (optional) and ddata [N,C,S...] (optional, accumulated atomically
// into a zero-initialised buffer).
// ---------------------------------------------------------------------------------------------
template <typename T, int NDIM>
__global__ void __launch_bounds__(256)
k_interp_bwd(const T* __restrict__ data, const T* __restrict__ grid, const T* __restrict__ gout,
             T* __restrict__ dgrid, T* __restrict__ ddata, Shape s)
{
    __shared__ T sg[NDIM][TILE][TILE + 1];
    const int mid = NDIM == 3 ? s.O[1] : 1;
    const int n = blockIdx.z / mid, im = blockIdx.z - n * mid;
    const int a0 = blockIdx.x * TILE, f0 = blockIdx.y * TILE;
    const int OF = s.O[NDIM - 1];
    const long nP = (long)s.O[0] * (NDIM >= 2 ? s.O[1] : 1) * (NDIM >= 3 ? s.O[2] : 1);
    const T* gn = grid + (size_t)n * NDIM * nP;
    const int a_ld = threadIdx.x & 31, b_ld = threadIdx.x >> 5;

#pragma unroll
    for (int rep = 0; rep < TILE / 8; ++rep) {
        const int fo = b_ld + 8 * rep;
        const int i0 = a0 + a_ld, iF = f0 + fo;
        if (i0 < s.O[0] && iF < OF) {
            const long p = NDIM == 2 ? (long)i0 + (long)s.O[0] * iF
                                     : (long)i0 + (long)s.O[0] * (im + (long)s.O[1] * iF);
#pragma unroll
            for (int j = 0; j < NDIM; ++j) sg[j][fo][a_ld] = gn[(size_t)j * nP + p];
        }
    }
    __syncthreads();
    {
        const int fo = threadIdx.x & 31, b = threadIdx.x >> 5;
        const int iF = f0 + fo;
        const size_t plane = (size_t)s.S[0] * (NDIM >= 2 ? s.S[1] : 1) * (NDIM >= 3 ? s.S[2] : 1);
        const size_t oplane = (size_t)nP;
#pragma unroll
        for (int rep = 0; rep < TILE / 8; ++rep) {
            const int a = b + 8 * rep;
            const int i0 = a0 + a;
            if (i0 >= s.O[0] || iF >= OF) continue;
            int t0[NDIM], t1[NDIM];
            T w[NDIM];
#pragma unroll
            for (int j = 0; j < NDIM; ++j) taps(sg[j][fo][a], s.S[j], t0[j], t1[j], w[j]);
            size_t off[1 << NDIM];
#pragma unroll
            for (int cn = 0; cn < (1 << NDIM); ++cn) {
                int t[3] = {0, 0, 0};
#pragma unroll
                for (int j = 0; j < NDIM; ++j) t[j] = ((cn >> j) & 1) ? t1[j] : t0[j];
                off[cn] = texel<NDIM>(t, s);
            }
            const size_t oidx = NDIM == 2 ? (size_t)i0 * s.O[1] + iF
                                          : ((size_t)i0 * s.O[1] + im) * s.O[2] + iF;
            T dg[NDIM];
#pragma unroll
            for (int j = 0; j < NDIM; ++j) dg[j] = 0;
            for (int c = 0; c < s.C; ++c) {
                const size_t ch = (size_t)n * s.C + c;
                const T* dp = data + ch * plane;
                const T g = gout[ch * oplane + oidx];
                T v[1 << NDIM], gv[1 << NDIM], dw[NDIM];
#pragma unroll
                for (int cn = 0; cn < (1 << NDIM); ++cn) v[cn] = __ldg(dp + off[cn]);
                blend_vjp<NDIM>(v, w, g, gv, dw);
#pragma unroll
                for (int j = 0; j < NDIM; ++j) dg[j] += dw[j];
                if (ddata != nullptr) {
#pragma unroll
                    for (int cn = 0; cn < (1 << NDIM); ++cn) atomicAdd(ddata + ch * plane + off[cn], gv[cn]);
                }
            }
            // xd = x - x0 with x = g*(size-1): d/dg = size-1
#pragma unroll
            for (int j = 0; j < NDIM; ++j) sg[j][fo][a] = dg[j] * (T)(s.S[j] - 1);
        }
    }
    if (dgrid == nullptr) return;
    __syncthreads();
    T* dn = dgrid + (size_t)n * NDIM * nP;
#pragma unroll
    for (int rep = 0; rep < TILE / 8; ++rep) {
        const int fo = b_ld + 8 * rep;
        const int i0 = a0 + a_ld, iF = f0 + fo;
        if (i0 < s.O[0] && iF < OF) {
            const long p = NDIM == 2 ? (long)i0 + (long)s.O[0] * iF
                                     : (long)i0 + (long)s.O[0] * (im + (long)s.O[1] * iF);
#pragma unroll
            for (int j = 0; j < NDIM; ++j) dn[(size_t)j * nP + p] = sg[j][fo][a_ld];
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
k_interp_bwd_1d(const T* __restrict__ data, const T* __restrict__ grid, const T* __restrict__ gout,
                T* __restrict__ dgrid, T* __restrict__ ddata, Shape s)
{
    const int n = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.O[0]) return;
    int t0, t1;
    T w[1];
    taps(grid[(size_t)n * s.O[0] + i], s.S[0], t0, t1, w[0]);
    T dg = 0;
    for (int c = 0; c < s.C; ++c) {
        const size_t ch = (size_t)n * s.C + c;
        const T* dp = data + ch * s.S[0];
        const T g = gout[ch * s.O[0] + i];
        T v[2] = {__ldg(dp + t0), __ldg(dp + t1)}, gv[2], dw[1];
        blend_vjp<1>(v, w, g, gv, dw);
        dg += dw[0];
        if (ddata != nullptr) {
            atomicAdd(ddata + ch * s.S[0] + t0, gv[0]);
            atomicAdd(ddata + ch * s.S[0] + t1, gv[1]);
        }
    }
    if (dgrid != nullptr) dgrid[(size_t)n * s.O[0] + i] = dg * (T)(s.S[0] - 1);
}

template <typename T>
int interp_t(bool backward, int ndim, const Shape& s, const void* data, const void* grid,
             const void* gout, void* out_or_dgrid, void* ddata, cudaStream_t st)
{
    if (s.N == 0 || s.C == 0) return kOk;
    for (int j = 0; j < ndim; ++j) if (s.O[j] == 0) return kOk;
    if (ndim == 1) {
        dim3 g((unsigned)((s.O[0] + 255) / 256), (unsigned)s.N);
        if (s.N > 65535) {   // grid.y limit: slab the batch
            for (int n0 = 0; n0 < s.N; n0 += 65535) {
                Shape sub = s;
                sub.N = s.N - n0 < 65535 ? s.N - n0 : 65535;
                const size_t din = (size_t)n0 * s.C * s.S[0], dout = (size_t)n0 * s.C * s.O[0];
                const size_t dgr = (size_t)n0 * s.O[0];
                int rc = interp_t<T>(backward, ndim, sub, (const T*)data + din, (const T*)grid + dgr,
                                     gout ? (const T*)gout + dout : nullptr,
                                     out_or_dgrid ? (T*)out_or_dgrid + (backward ? dgr : dout) : nullptr,
                                     ddata ? (T*)ddata + din : nullptr, st);
                if (rc != kOk) return rc;
            }
            return kOk;
        }
        prof_begin(backward ? kProfInterpBwd : kProfInterpFwd, st);
        if (!backward) k_interp_fwd_1d<T><<<g, 256, 0, st>>>((const T*)data, (const T*)grid, (T*)out_or_dgrid, s);
        else k_interp_bwd_1d<T><<<g, 256, 0, st>>>((const T*)data, (const T*)grid, (const T*)gout, (T*)out_or_dgrid, (T*)ddata, s);
        prof_end(backward ? kProfInterpBwd : kProfInterpFwd, st);
        count_launch();
        CPAB_CUDA_OK(cudaGetLastError());
        return kOk;
    }
    const long z = (long)s.N * (ndim == 3 ? s.O[1] : 1);
    const unsigned gy = (unsigned)((s.O[ndim - 1] + TILE - 1) / TILE);
    if (z > 65535 || gy > 65535) { set_error("interpolate: batch x middle extent %ld exceeds 65535", z); return kErrUnsupported; }
    dim3 g((unsigned)((s.O[0] + TILE - 1) / TILE), gy, (unsigned)z);
    prof_begin(backward ? kProfInterpBwd : kProfInterpFwd, st);
    if (ndim == 2) {
        if (!backward) k_interp_fwd<T, 2><<<g, 256, 0, st>>>((const T*)data, (const T*)grid, (T*)out_or_dgrid, s);
        else k_interp_bwd<T, 2><<<g, 256, 0, st>>>((const T*)data, (const T*)grid, (const T*)gout, (T*)out_or_dgrid, (T*)ddata, s);
    } else {
        if (!backward) k_interp_fwd<T, 3><<<g, 256, 0, st>>>((const T*)data, (const T*)grid, (T*)out_or_dgrid, s);
        else k_interp_bwd<T, 3><<<g, 256, 0, st>>>((const T*)data, (const T*)grid, (const T*)gout, (T*)out_or_dgrid, (T*)ddata, s);
    }
    prof_end(backward ? kProfInterpBwd : kProfInterpFwd, st);
    count_launch();
    CPAB_CUDA_OK(cudaGetLastError());
    return kOk;
}

Shape make_shape(int ndim, int N, int C, const int* in_size, const int* out_size)
{
    Shape s;
    s.N = N; s.C = C;
    for (int j = 0; j < 3; ++j) { s.S[j] = j < ndim ? in_size[j] : 1; s.O[j] = j < ndim ? out_size[j] : 1; }
    return s;
}

}  // namespace

int launch_interp_forward(int dtype, int ndim, int N, int C, const int* in_size, const int* out_size,
                          const void* data, const void* grid, void* out, cudaStream_t st)
{
    const Shape s = make_shape(ndim, N, C, in_size, out_size);
    return dtype == kF32 ? interp_t<float>(false, ndim, s, data, grid, nullptr, out, nullptr, st)
                         : interp_t<double>(false, ndim, s, data, grid, nullptr, out, nullptr, st);
}

int launch_interp_backward(int dtype, int ndim, int N, int C, const int* in_size, const int* out_size,
                           const void* data, const void* grid, const void* grad_out, void* dgrid,
                           void* ddata, cudaStream_t st)
{
    const Shape s = make_shape(ndim, N, C, in_size, out_size);
    if (ddata != nullptr) {
        size_t elems = (size_t)N * C;
        for (int j = 0; j < ndim; ++j) elems *= in_size[j];
        CPAB_CUDA_OK(cudaMemsetAsync(ddata, 0, elems * (dtype == kF32 ? 4 : 8), st));
    }
    return dtype == kF32 ? interp_t<float>(true, ndim, s, data, grid, grad_out, dgrid, ddata, st)
                         : interp_t<double>(true, ndim, s, data, grid, grad_out, dgrid, ddata, st);
}

}  // namespace cpab
