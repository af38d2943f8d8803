// cpab_integrate.cu -- integration of the CPA velocity field and its theta-gradient (sm_100a).
//
// Kernels in this file and the reference code each one replaces (paths under /root/reference):
//
//   k_findcellidx     libcpab/core/cpab_ops.cu:14-227 (device findcellidx) -- exposed for tests and
//                     for Cpab.visualize_tesselation-style callers
//   k_forward         libcpab/core/cpab_ops.cu:268-388 + launch libcpab/pytorch/transformer_cuda.cu:18-64
//                     nstepsolver x { cell search ; p <- Trels[cell] [p;1] }
//   k_jacobian        libcpab/core/cpab_ops.cu:390-697 + launch transformer_cuda.cu:66-119
//                     the reference-layout [d,n_theta,ndim,nP] RK2 Jacobian (kept for drop-in and
//                     op-level parity; d-fold redundant by construction)
//   k_backward        the same RK2 discretisation restated as an adjoint sweep: one pass per
//                     (point,theta) instead of one per (point,theta,k); accumulates
//                     R[theta][cell] and never materialises the Jacobian.  Together with
//                     k_prepare_backward (per-cell RK2 step records), k_r_to_g and
//                     k_grad_epilogue (dtheta = G . B) it replaces cpab_ops.cu:390-697 AND the
//                     contraction libcpab/pytorch/transformer.py:201.
//
// Design notes (B200): one thread owns one (point,theta) trajectory; a CTA works on one theta at a
// time so that theta's per-cell records are staged in shared memory and every step's gather is a
// shared-memory read.  k_forward: static grid of n_theta x 1024-point units; its step loop holds
// no call (the warp votes and leaves the loop when a lane needs the complete cell search).
// k_backward: persistent grid, units drawn from a counter (WorkPlan).  2-D arithmetic is packed
// FP32 (cpab_f32x2.cuh).  Loads/stores of the planar [ndim,nP] point arrays are unit-stride per
// coordinate.  The loop trip count is fixed (nstepsolver), so divergence is confined to the rare
// exact paths of the cell search and to the flushes of the per-thread gradient accumulators.
#include <atomic>
#include <string_view>
#include <type_traits>

#include "cpab_common.cuh"
#include "cpab_sample.cuh"

// -DCPAB_FAST_BUILD instantiates only the float32 2-D kernels (SASS experiments; never shipped).
#ifdef CPAB_FAST_BUILD
#ifndef CPAB_FAST_DIM
#define CPAB_FAST_DIM 2
#endif
#if CPAB_FAST_DIM == 1
#define CPAB_DISPATCH(T1D, T2D, T3D) (T1D)
#elif CPAB_FAST_DIM == 2
#define CPAB_DISPATCH(T1D, T2D, T3D) (T2D)
#else
#define CPAB_DISPATCH(T1D, T2D, T3D) (T3D)
#endif
#define CPAB_DTYPE(F, D) (F)
#else
#define CPAB_DISPATCH(T1D, T2D, T3D) (g.ndim == 1 ? (T1D) : g.ndim == 2 ? (T2D) : (T3D))
#define CPAB_DTYPE(F, D) (dtype == kF32 ? (F) : (D))
#endif


namespace cpab {

// ---- rounding-controlled scalar ops ---------------------------------------------------------------
template <typename T> struct Num;
template <> struct Num<float> {
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float fma(float a, float b, float c) { return fmaf(a, b, c); }
    static __device__ __forceinline__ float atomic_add(float* p, float v) { return atomicAdd(p, v); }
};
template <> struct Num<double> {
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double fma(double a, double b, double c) { return ::fma(a, b, c); }
    static __device__ __forceinline__ double atomic_add(double* p, double v) { return atomicAdd(p, v); }
};

// Fire-and-forget reduction of one cell's COUNT accumulators into G.  float32 uses the vector forms
// (REDG.E.ADD.F32x2 / F32x4, sm_90+): a 2-D cell is 3 instructions instead of 6, a 3-D cell 3 instead
// of 12 -- this code runs divergently (lanes leave cells at different steps), so its length is paid
// per occurrence by the whole warp.  Needs addr 8-byte (COUNT % 4 != 0) / 16-byte aligned, which the
// [n_theta][nC][ndim][ndim+1] layout of G gives for a 16-byte aligned workspace.
template <int COUNT> __device__ __forceinline__ void red_cell(float* addr, const float* v)
{
    if (COUNT % 4 == 0) {
#pragma unroll
        for (int e = 0; e < COUNT; e += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                         :: "l"(addr + e), "f"(v[e]), "f"(v[e + 1]), "f"(v[e + 2]), "f"(v[e + 3]) : "memory");
    } else {
#pragma unroll
        for (int e = 0; e < COUNT; e += 2)
            asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" :: "l"(addr + e), "f"(v[e]), "f"(v[e + 1]) : "memory");
    }
}
template <int COUNT> __device__ __forceinline__ void red_cell(double* addr, const double* v)
{
#pragma unroll
    for (int e = 0; e < COUNT; ++e) atomicAdd(addr + e, v[e]);
}

// ---- per-cell matrix fetch (vectorised; the row-major [n][n+1] block is 8/16-byte aligned) --------
template <int NDIM> __device__ __forceinline__ void load_affine(const float* M, float* a)
{
    if (NDIM == 3) {
        const float4* v = reinterpret_cast<const float4*>(M);
#pragma unroll
        for (int i = 0; i < 3; ++i) { const float4 t = v[i]; a[4*i] = t.x; a[4*i+1] = t.y; a[4*i+2] = t.z; a[4*i+3] = t.w; }
    } else {
        const float2* v = reinterpret_cast<const float2*>(M);
#pragma unroll
        for (int i = 0; i < Dim<NDIM>::kPpc / 2; ++i) { const float2 t = v[i]; a[2*i] = t.x; a[2*i+1] = t.y; }
    }
}
template <int NDIM> __device__ __forceinline__ void load_affine(const double* M, double* a)
{
    const double2* v = reinterpret_cast<const double2*>(M);
#pragma unroll
    for (int i = 0; i < Dim<NDIM>::kPpc / 2; ++i) { const double2 t = v[i]; a[2*i] = t.x; a[2*i+1] = t.y; }
}

// Per-theta table of per-cell matrices: shared memory (32-bit shared-window address, explicit
// ld.shared so that the address arithmetic is one IMAD per step) or, for tessellations too large
// to stage, global memory through the read-only path.
__device__ __forceinline__ void lds_vec(uint32_t addr, float* a, int n4, int n2)
{
    for (int i = 0; i < n4; ++i)
        asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
            : "=f"(a[4 * i]), "=f"(a[4 * i + 1]), "=f"(a[4 * i + 2]), "=f"(a[4 * i + 3]) : "r"(addr + 16 * i));
    for (int i = 0; i < n2; ++i)
        asm("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(a[2 * i]), "=f"(a[2 * i + 1]) : "r"(addr + 8 * i));
}
__device__ __forceinline__ void lds_vec(uint32_t addr, double* a, int n4, int n2)
{
    (void)n4;
    for (int i = 0; i < n2; ++i)
        asm("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(a[2 * i]), "=d"(a[2 * i + 1]) : "r"(addr + 16 * i));
}

// 16-byte global loads of COUNT consecutive elements (COUNT * sizeof(T) a multiple of 16, aligned)
template <typename T, int COUNT> __device__ __forceinline__ void load_vec16(const T* M, T* a)
{
    constexpr int PER = 16 / sizeof(T);
    const int4* v = reinterpret_cast<const int4*>(M);
#pragma unroll
    for (int i = 0; i < COUNT / PER; ++i) {
        const int4 t = __ldg(v + i);
        memcpy(a + PER * i, &t, 16);
    }
}

// STRIDE = elements per cell: the affine block itself (ndim (ndim+1)) for A / Trels tables, the
// padded RK2 step record (StepRec) for the backward sweep.
template <typename T, int NDIM, bool SMEM, int STRIDE = Dim<NDIM>::kPpc> struct CellTable {
    const T* gptr;
    uint32_t saddr;
    __device__ __forceinline__ void load(int c, T* a) const
    {
        if (SMEM) {
            const uint32_t addr = saddr + (uint32_t)c * (uint32_t)(STRIDE * sizeof(T));
            if ((STRIDE * sizeof(T)) % 16 == 0) {
                if (sizeof(T) == 4) lds_vec(addr, a, STRIDE / 4, 0);
                else lds_vec(addr, a, 0, STRIDE / 2);
            } else {
                lds_vec(addr, a, 0, STRIDE / 2);
            }
        } else if (STRIDE == Dim<NDIM>::kPpc) {
            load_affine<NDIM>(gptr + (size_t)c * STRIDE, a);
        } else {
            load_vec16<T, STRIDE>(gptr + (size_t)c * STRIDE, a);
        }
    }
};

// Shared-memory record of one cell's Trels block in k_forward.  2-D float32: the 2x3 block is
// re-ordered to [a00 a01 a10 a11 | a02 a12 . .] (32 bytes): one LDS.128 + one LDS.64 instead of
// three LDS.64, and the four products of T [p;1] are two packed multiplies (a_r0, a_r1) * (p0, p1).
template <typename T, int NDIM, bool SMEM> struct FwdRec {
    static constexpr bool kPacked = SMEM && NDIM == 2 && sizeof(T) == 4;
    static constexpr int kStride = kPacked ? 8 : Dim<NDIM>::kPpc;
};

// out = A [v;1] in the reference's left-to-right order with every product and sum rounded
// (cpab_ops.cpp:192-206) -- bit-identical to the CPU reference.
template <int NDIM, typename T>
__device__ __forceinline__ void affine_strict(const T* A, const T* v, T* out)
{
#pragma unroll
    for (int r = 0; r < NDIM; ++r) {
        T acc = Num<T>::mul(A[r * (NDIM + 1)], v[0]);
#pragma unroll
        for (int c = 1; c < NDIM; ++c) acc = Num<T>::add(acc, Num<T>::mul(A[r * (NDIM + 1) + c], v[c]));
        out[r] = Num<T>::add(acc, A[r * (NDIM + 1) + NDIM]);
    }
}
// same map as a nest of FMAs (fast-math mode, and everywhere inside the gradient)
template <int NDIM, typename T>
__device__ __forceinline__ void affine_fma(const T* A, const T* v, T* out)
{
#pragma unroll
    for (int r = 0; r < NDIM; ++r) {
        T acc = A[r * (NDIM + 1) + NDIM];
#pragma unroll
        for (int c = NDIM - 1; c >= 0; --c) acc = Num<T>::fma(A[r * (NDIM + 1) + c], v[c], acc);
        out[r] = acc;
    }
}
// linear part only, strict order (cpab_ops.cpp:208-222)
template <int NDIM, typename T>
__device__ __forceinline__ void linear_strict(const T* A, const T* v, T* out)
{
#pragma unroll
    for (int r = 0; r < NDIM; ++r) {
        T acc = Num<T>::mul(A[r * (NDIM + 1)], v[0]);
#pragma unroll
        for (int c = 1; c < NDIM; ++c) acc = Num<T>::add(acc, Num<T>::mul(A[r * (NDIM + 1) + c], v[c]));
        out[r] = acc;
    }
}

// stage one theta's [nC][ppc] block into shared memory with 16-byte copies
template <typename T>
__device__ __forceinline__ void stage_block(T* dst, const T* __restrict__ src, int count)
{
    const int vec = 16 / sizeof(T);
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (count % vec) == 0) {
        const int4* s4 = reinterpret_cast<const int4*>(src);
        int4* d4 = reinterpret_cast<int4*>(dst);
        for (int i = threadIdx.x; i < count / vec; i += blockDim.x) d4[i] = __ldg(s4 + i);
    } else {
        for (int i = threadIdx.x; i < count; i += blockDim.x) dst[i] = __ldg(src + i);
    }
}

// =====================================================================================================
// fused transform_data: sampling epilogue of the forward, sampling-VJP prologue of the adjoint.
// The integration kernels are issue-bound with idle memory bandwidth, the stand-alone sampling
// kernels are latency-bound; one gather per trajectory at either end of a 50-step loop costs ~2 %
// and removes two launches and the d/dgrid round trip.  Same arithmetic as cpab_interp.cu
// (shared helpers), hence identical results.
// =====================================================================================================
template <int NDIM>
__device__ __forceinline__ int image_index(long p, const Shape& s)
{
    // grid point p = i0 + O0 (i1 + O1 i2)  ->  offset in a [O0,O1(,O2)] image, last index fastest
    const int O0 = s.O[0];
    if (NDIM == 1) return (int)p;
    const int q = (int)(p / O0), i0 = (int)(p - (long)q * O0);
    if (NDIM == 2) return i0 * s.O[1] + q;
    const int i2 = q / s.O[1], i1 = q - i2 * s.O[1];
    return (i0 * s.O[1] + i1) * s.O[2] + i2;
}

template <typename T, int NDIM>
__device__ __forceinline__ void sample_store(const T* pt, int n, long p, const T* __restrict__ data,
                                             T* __restrict__ img, const Shape& s)
{
    const Taps<T, NDIM> tp = make_taps<T, NDIM>(pt, s);
    const int plane = s.S[0] * (NDIM >= 2 ? s.S[1] : 1) * (NDIM >= 3 ? s.S[2] : 1);
    const int nPo = s.O[0] * (NDIM >= 2 ? s.O[1] : 1) * (NDIM >= 3 ? s.O[2] : 1);
    const T* dp = data + (size_t)n * s.C * plane;
    T* op = img + (size_t)n * s.C * nPo + image_index<NDIM>(p, s);
#pragma unroll 1
    for (int c = 0; c < s.C; ++c, dp += plane, op += nPo) {
        T v[1 << NDIM];
        gather<T, NDIM>(dp, tp, v);
        *op = blend<NDIM>(v, tp.w);
    }
}

template <typename T, int NDIM>
__device__ __forceinline__ void sample_vjp(const T* pt, int n, long p, const T* __restrict__ data,
                                           const T* __restrict__ gimg, const Shape& s, T* lam)
{
    const Taps<T, NDIM> tp = make_taps<T, NDIM>(pt, s);
    const int plane = s.S[0] * (NDIM >= 2 ? s.S[1] : 1) * (NDIM >= 3 ? s.S[2] : 1);
    const int nPo = s.O[0] * (NDIM >= 2 ? s.O[1] : 1) * (NDIM >= 3 ? s.O[2] : 1);
    const T* dp = data + (size_t)n * s.C * plane;
    const T* gp = gimg + (size_t)n * s.C * nPo + image_index<NDIM>(p, s);
#pragma unroll
    for (int j = 0; j < NDIM; ++j) lam[j] = 0;
#pragma unroll 1
    for (int c = 0; c < s.C; ++c, dp += plane, gp += nPo) {
        T v[1 << NDIM], gv[1 << NDIM], dw[NDIM];
        gather<T, NDIM>(dp, tp, v);
        blend_vjp<NDIM>(v, tp.w, *gp, gv, dw);
#pragma unroll
        for (int j = 0; j < NDIM; ++j) lam[j] += dw[j];
    }
#pragma unroll
    for (int j = 0; j < NDIM; ++j) lam[j] *= (T)(s.S[j] - 1);
}

// =====================================================================================================
// Work distribution of the adjoint kernel.
//
// The grid is persistent (one CTA per resident slot); CTAs draw work units from a global counter.
// A unit is a range of points of one theta: `bulk_pts` points while plenty of work remains, and
// `small_pts` for the last ~2 units per slot, so that every SM runs until the end -- with a static
// grid of equal CTAs the SMs of a B200 finished up to 14 % apart on BASELINE configs[1]
// (sm__cycles_active min/max 1169k/1354k; profiles/r01b_*), although every CTA does the same work.
// (k_forward keeps a static grid: there the per-unit barrier and bookkeeping cost more than the
// balance gained -- 2.21 vs 2.02 ms on 128 thetas x 512^2.)
// Units are numbered theta-major: resident CTAs spread over many thetas, which keeps the G
// reductions of one theta from colliding in L2.
//
// Counters live in a small ring of self-resetting slots in module memory (no host memset, no
// caller-provided buffer, usable under graph capture): the last CTA to leave resets the slot.
// =====================================================================================================
struct WorkPlan {
    unsigned total_bulk, total;      // units of the bulk phase / of both phases
    int bulk_per_theta, small_per_theta;
    int bulk_pts, small_pts;
    long nP_bulk;                    // points [0, nP_bulk) of every theta are bulk units, the rest small ones
    unsigned slot;                   // index into g_work_ring
};
constexpr int kWorkRing = 256;
__device__ unsigned int g_work_ring[kWorkRing][2];

struct WorkUnit { int theta; long begin, end; };

// Unit number -> (theta, point range): bulk units of all thetas first, then the small ones.
__device__ __forceinline__ void unit_of(const WorkPlan& wp, unsigned w, long nP, WorkUnit& u)
{
    if (w < wp.total_bulk) {
        u.theta = (int)(w / (unsigned)wp.bulk_per_theta);
        const int c = (int)(w - (unsigned)u.theta * (unsigned)wp.bulk_per_theta);
        u.begin = (long)c * wp.bulk_pts;
        u.end = u.begin + wp.bulk_pts < wp.nP_bulk ? u.begin + wp.bulk_pts : wp.nP_bulk;
    } else {
        const unsigned w2 = w - wp.total_bulk;
        u.theta = (int)(w2 / (unsigned)wp.small_per_theta);
        const int c = (int)(w2 - (unsigned)u.theta * (unsigned)wp.small_per_theta);
        u.begin = wp.nP_bulk + (long)c * wp.small_pts;
        u.end = u.begin + wp.small_pts < nP ? u.begin + wp.small_pts : nP;
    }
}

// All threads of the CTA must call this; returns false when the work is exhausted.
__device__ __forceinline__ bool next_unit(const WorkPlan& wp, long nP, unsigned* s_work, WorkUnit& u)
{
    __syncthreads();                                   // everyone is done with the previous unit
    if (threadIdx.x == 0) *s_work = atomicAdd(&g_work_ring[wp.slot][0], 1u);
    __syncthreads();
    const unsigned w = *s_work;
    if (w >= wp.total) return false;
    unit_of(wp, w, nP, u);
    return true;
}

__device__ __forceinline__ void leave_grid(const WorkPlan& wp)
{
    if (threadIdx.x == 0) {
        const unsigned done = atomicAdd(&g_work_ring[wp.slot][1], 1u);
        if (done == gridDim.x - 1) {                   // every CTA has stopped drawing: recycle the slot
            g_work_ring[wp.slot][0] = 0;
            g_work_ring[wp.slot][1] = 0;
            __threadfence();
        }
    }
}

// =====================================================================================================
// findcellidx
// =====================================================================================================
template <typename T, int NDIM>
__global__ void __launch_bounds__(256) k_findcellidx(const T* __restrict__ pts, long nP,
                                                      int* __restrict__ out, const __grid_constant__ Geom g)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nP) return;
    T p[3] = {0, 0, 0};
#pragma unroll
    for (int j = 0; j < NDIM; ++j) p[j] = pts[i + (long)j * nP];
    out[i] = find_cell<NDIM>(p, g);
}

// =====================================================================================================
// forward: nsteps x { c = cell(p) ; p = Trels[theta][c] [p;1] }
// =====================================================================================================
template <typename T, int NDIM, bool STRICT, bool SMEM, int PPT, bool SAMPLE>
__global__ void __launch_bounds__(256, (sizeof(T) == 4 ? 4 : 1))
k_forward(const T* __restrict__ points, const T* __restrict__ trels, T* __restrict__ out, long nP,
          int broadcast, int nsteps, const __grid_constant__ Geom g, const __grid_constant__ WorkPlan wp,
          const T* __restrict__ data, T* __restrict__ img, const __grid_constant__ Shape sh)
{
    constexpr int PPC = Dim<NDIM>::kPpc;
    constexpr bool kPacked = FwdRec<T, NDIM, SMEM>::kPacked;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // static grid, one CTA per work unit (1024 points of one theta; 256 when the whole problem
    // would not fill the chip otherwise)
    WorkUnit wu;
    unit_of(wp, blockIdx.x, nP, wu);
    const int theta = wu.theta;
    const int tsize = g.n_cells * PPC;
    CellTable<T, NDIM, SMEM, FwdRec<T, NDIM, SMEM>::kStride> tab;
    tab.gptr = trels + (size_t)theta * tsize;
    tab.saddr = 0;
    if (SMEM) {
        T* sT = reinterpret_cast<T*>(smem_raw);
        if (kPacked) {
            const float2* src2 = reinterpret_cast<const float2*>(tab.gptr);     // 24-byte blocks, 8-byte aligned
            float4* dst4 = reinterpret_cast<float4*>(smem_raw);
            for (int c = threadIdx.x; c < g.n_cells; c += blockDim.x) {
                const float2 r0 = __ldg(src2 + 3 * c), r1 = __ldg(src2 + 3 * c + 1), r2 = __ldg(src2 + 3 * c + 2);
                dst4[2 * c] = make_float4(r0.x, r0.y, r1.y, r2.x);               // a00 a01 a10 a11
                dst4[2 * c + 1] = make_float4(r1.x, r2.y, 0.0f, 0.0f);            // a02 a12
            }
        } else {
            stage_block(sT, tab.gptr, tsize);
        }
        __syncthreads();
        // CTA-local offset (no cluster launch: rank bits are 0).  `nsteps >> 30` is a zero ptxas
        // does not know: without it the base is rebuilt in the loop (S2R, MOV, LEA, LOP3 per step)
        // wherever the mask does not fold to a constant.
        tab.saddr = ((uint32_t)__cvta_generic_to_shared(sT) & 0xffffffu) + (uint32_t)(nsteps >> 30);
    }
    const T* src = points + (broadcast ? (size_t)theta * NDIM * nP : 0);
    T* dst = out + (size_t)theta * NDIM * nP;
    const long begin = wu.begin, end = wu.end;

    const float magic = 12582912.0f;     // 1.5 * 2^23, rounding constant of the 2-D cell search
    for (long b0 = begin; b0 < end; b0 += (long)blockDim.x * PPT) {      // warp-uniform trip count
        const long base = b0 + threadIdx.x;
        T p[PPT][NDIM];
#pragma unroll
        for (int u = 0; u < PPT; ++u) {
            const long i = base + (long)u * blockDim.x;
#pragma unroll
            for (int j = 0; j < NDIM; ++j) p[u][j] = i < end ? src[i + (long)j * nP] : (T)0.25;
        }
        auto advance = [&](int u, int c) {
            if constexpr (kPacked) {
                const uint32_t addr = tab.saddr + (uint32_t)c * 32u;
                float a[4], t[2];
                lds_vec(addr, a, 1, 0);
                asm("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(t[0]), "=f"(t[1]) : "r"(addr + 16));
                if (STRICT) {       // products packed, sums scalar and separately rounded (cpab_f32x2.cuh)
                    const F2 P = pk(p[u][0], p[u][1]);
                    const F2 m0 = mul2(pk(a[0], a[1]), P), m1 = mul2(pk(a[2], a[3]), P);
                    p[u][0] = __fadd_rn(__fadd_rn(lo(m0), hi(m0)), t[0]);
                    p[u][1] = __fadd_rn(__fadd_rn(lo(m1), hi(m1)), t[1]);
                } else {
                    const float x = p[u][0], y = p[u][1];
                    p[u][0] = fmaf(a[0], x, fmaf(a[1], y, t[0]));
                    p[u][1] = fmaf(a[2], x, fmaf(a[3], y, t[1]));
                }
            } else if constexpr (NDIM == 3 && sizeof(T) == 4 && STRICT) {
                // row r = (a_r0 a_r1 | a_r2 a_r3): the first two products as one packed multiply with
                // (p0 p1), sums scalar and separately rounded, in the reference's order
                float a[PPC];
                tab.load(c, a);
                const F2 P01 = pk(p[u][0], p[u][1]);
                float q[3];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const F2 m0 = mul2(pk(a[4 * r], a[4 * r + 1]), P01);
                    const float m2 = __fmul_rn(a[4 * r + 2], p[u][2]);
                    q[r] = __fadd_rn(__fadd_rn(__fadd_rn(lo(m0), hi(m0)), m2), a[4 * r + 3]);
                }
#pragma unroll
                for (int j = 0; j < 3; ++j) p[u][j] = q[j];
            } else {
                T a[PPC], q[NDIM];
                tab.load(c, a);
                if (STRICT) affine_strict<NDIM>(a, p[u], q); else affine_fma<NDIM>(a, p[u], q);
#pragma unroll
                for (int j = 0; j < NDIM; ++j) p[u][j] = q[j];
            }
        };
        // The inner loop runs fast-path steps until some lane of the warp needs the complete cell
        // search (a point on a diagonal, a corner outside the domain...; ~1e-5 per point and
        // step); that one step is then done for the whole warp below and the loop resumes.
        constexpr bool kHasRarePath = NDIM >= 2 && sizeof(T) == 4;   // see find_cell_try
        int s = 0;
        if (!kHasRarePath) {
            for (; s < nsteps; ++s) {
#pragma unroll
                for (int u = 0; u < PPT; ++u) advance(u, find_cell<NDIM>(p[u], g));
            }
        }
        while (kHasRarePath) {
            int c[PPT];
            bool rare[PPT];
            CellEst est[PPT];
#pragma unroll 2
            for (; s < nsteps; ++s) {
                bool any = false;
#pragma unroll
                for (int u = 0; u < PPT; ++u) any |= (rare[u] = find_cell_try<NDIM>(p[u], g, magic, c[u], est[u]));
                if (__any_sync(0xffffffffu, any)) break;
#pragma unroll
                for (int u = 0; u < PPT; ++u) advance(u, c[u]);
            }
            if (s >= nsteps) break;
            // slow step: the lanes that asked for it finish their search from the estimates
#pragma unroll
            for (int u = 0; u < PPT; ++u) advance(u, rare[u] ? find_cell_finish<NDIM>(p[u], g, est[u]) : c[u]);
            ++s;
        }
#pragma unroll
        for (int u = 0; u < PPT; ++u) {
            const long i = base + (long)u * blockDim.x;
            if (i < end) {
#pragma unroll
                for (int j = 0; j < NDIM; ++j) dst[i + (long)j * nP] = p[u][j];
                if (SAMPLE) sample_store<T, NDIM>(p[u], theta, i, data, img, sh);
            }
        }
    }
}

// =====================================================================================================
// reference-layout Jacobian: per (point, theta, k), RK2 with the reference's `double h`
// (cpab_ops.cpp:289-366).  Every float expression is evaluated in the reference's order.
// =====================================================================================================
template <typename T, int NDIM>
__global__ void __launch_bounds__(128)
k_jacobian(const T* __restrict__ points, const T* __restrict__ As, const T* __restrict__ Bs,
           T* __restrict__ jac, long nP, int n_theta, int d, int broadcast, int nsteps, const __grid_constant__ Geom g)
{
    constexpr int PPC = Dim<NDIM>::kPpc;
    const int tk = blockIdx.y;                      // theta * d + k  (host keeps n_theta*d <= 65535,
    const int theta = tk / d, k = tk - theta * d;   //  otherwise loops over slabs)
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nP) return;
    const size_t tsize = (size_t)g.n_cells * PPC;
    const T* At = As + (size_t)theta * tsize;
    const T* Bk = Bs + (size_t)k * tsize;
    const T* src = points + (broadcast ? (size_t)theta * NDIM * nP : 0);
    const double h = 1.0 / nsteps;

    T p[NDIM], q[NDIM];
#pragma unroll
    for (int j = 0; j < NDIM; ++j) { p[j] = src[i + (long)j * nP]; q[j] = 0; }
    for (int s = 0; s < nsteps; ++s) {
        const int c = find_cell<NDIM>(p, g);
        T A[PPC], B[PPC];
        load_affine<NDIM>(At + (size_t)c * PPC, A);
        load_affine<NDIM>(Bk + (size_t)c * PPC, B);
        T v[NDIM], pm[NDIM], vm[NDIM], bt[NDIM], aq[NDIM], u[NDIM], qm[NDIM], um[NDIM];
        affine_strict<NDIM>(A, p, v);
#pragma unroll
        for (int j = 0; j < NDIM; ++j)   // p + h*v/2.0 in double, rounded on store
            pm[j] = (T)__dadd_rn((double)p[j], __dmul_rn(__dmul_rn(h, (double)v[j]), 0.5));
        affine_strict<NDIM>(A, pm, vm);
        affine_strict<NDIM>(B, p, bt);
        linear_strict<NDIM>(A, q, aq);
#pragma unroll
        for (int j = 0; j < NDIM; ++j) u[j] = Num<T>::add(bt[j], aq[j]);
#pragma unroll
        for (int j = 0; j < NDIM; ++j)
            qm[j] = (T)__dadd_rn((double)q[j], __dmul_rn(__dmul_rn(h, (double)u[j]), 0.5));
        affine_strict<NDIM>(B, pm, bt);
        linear_strict<NDIM>(A, qm, aq);
#pragma unroll
        for (int j = 0; j < NDIM; ++j) um[j] = Num<T>::add(bt[j], aq[j]);
#pragma unroll
        for (int j = 0; j < NDIM; ++j) q[j] = (T)__dadd_rn((double)q[j], __dmul_rn((double)um[j], h));
#pragma unroll
        for (int j = 0; j < NDIM; ++j) p[j] = (T)__dadd_rn((double)p[j], __dmul_rn((double)vm[j], h));
    }
    T* dst = jac + ((size_t)k * n_theta + theta) * NDIM * nP;
#pragma unroll
    for (int j = 0; j < NDIM; ++j) dst[i + (long)j * nP] = q[j];
}

// =====================================================================================================
// adjoint backward.
//
// Per step (cell c, A = A_c, h = 1/nsteps) the reference's RK2 recursion for the sensitivity
// q_k = dp/dtheta_k is   q+ = M q + h B_kc [pMid;1] + (h^2/2) A_lin B_kc [p;1],
// M = I + h A_lin + (h^2/2) A_lin^2, which is linear in the entries of B_k restricted to cell c.
// With lambda_N = dL/dp_N and lambda_n = M_n^T lambda_{n+1},
//     dL/dtheta_k = sum_c < B_kc , G_c >,   G_c += h lambda_{n+1} [pMid;1]^T + (h^2/2)(A_lin^T lambda_{n+1}) [p;1]^T
// so one reverse sweep per (point,theta) yields G[theta] (nC x ndim x (ndim+1)) and the epilogue
// dtheta = G . B finishes the job; lambda_0 is dL/dpoints for free.
//
// The reverse sweep needs p_n.  Trajectories are checkpointed every SEG steps in shared memory
// during a first forward pass and recomputed segment by segment into registers.
// =====================================================================================================
// One RK2 (midpoint) step inside a cell is itself an affine map of the point:
//     p+ = p + h (L pMid + t),  pMid = p + (h/2)(L p + t)   =>   p+ = p + (D p + s),
//     D = h L + (h^2/2) L^2,    s = h t + (h^2/2) L t,
// and the sensitivity recursion's M = I + h L + (h^2/2) L^2 is I + D.  k_prepare_backward
// evaluates one record per (theta, cell) (in double, rounded once); the sweeps then cost one affine
// map per step instead of two, and lambda_n = lambda_{n+1} + D^T lambda_{n+1}.
//   * The point is advanced by an *increment*, so each step rounds like the reference's own
//     `p += vMid * h` (no systematic error from storing 1 + D_ii in float).
//   * D p and s cancel (zero-boundary fields: |L p|, |t| >> |v|), which would amplify the
//     rounding of the stored D and s; the record therefore holds the map about an origin o inside
//     the cell:  inc = D (p - o) + s',  s' = s + D o  -- |p - o| is at most a cell, s' is the
//     increment at o itself, nothing cancels.
// Record layout (StepRec<NDIM>::kStride elements, 16-byte multiple):
//     D [n][n] COLUMN-major | s' [n] | o [n] | padding
// (column-major so that in 2-D a column, s' and o are register pairs straight out of two LDS.128:
// the step is FADD2, FFMA2, FFMA2, FADD2 -- packed FP32, cpab_f32x2.cuh.)
// 1-D keeps the plain pair (D, s) about the global origin: the loop there is bound by the
// shared-memory data pipe, a 16-byte record costs twice the wavefronts of an 8-byte one (measured
// 18 % on 8192 x 1024), and |L p|, |t| stay within a small multiple of |v| for 1-D tessellations.
template <int NDIM> struct StepRec {
    static constexpr bool kLocal = NDIM > 1;
    static constexpr int kStride = NDIM == 1 ? 2 : NDIM == 2 ? 8 : 16;
    static constexpr int kS = NDIM * NDIM;          // offset of s'
    static constexpr int kO = NDIM * NDIM + NDIM;   // offset of o
    static __host__ __device__ constexpr int d(int r, int c) { return c * NDIM + r; }   // D[r][c]
};

template <int NDIM, typename T> struct UsePacked { static constexpr bool value = false; };
template <> struct UsePacked<2, float> { static constexpr bool value = true; };

template <int NDIM, typename T>
__device__ __forceinline__ void step_inc(const T* W, T* p)
{
    if constexpr (UsePacked<NDIM, T>::value) {
        F2 P = pk(p[0], p[1]);
        const F2 q = sub2(P, pk(W[6], W[7]));
        F2 inc = fma2(pk(W[2], W[3]), bc(hi(q)), pk(W[4], W[5]));
        inc = fma2(pk(W[0], W[1]), bc(lo(q)), inc);
        P = add2(P, inc);
        unpk(P, p[0], p[1]);
    } else {
        T q[NDIM], inc[NDIM];
#pragma unroll
        for (int j = 0; j < NDIM; ++j) q[j] = StepRec<NDIM>::kLocal ? p[j] - W[StepRec<NDIM>::kO + j] : p[j];
#pragma unroll
        for (int r = 0; r < NDIM; ++r) {
            T acc = W[StepRec<NDIM>::kS + r];
#pragma unroll
            for (int c = NDIM - 1; c >= 0; --c) acc = Num<T>::fma(W[StepRec<NDIM>::d(r, c)], q[c], acc);
            inc[r] = acc;
        }
#pragma unroll
        for (int j = 0; j < NDIM; ++j) p[j] += inc[j];
    }
}

// R_c += lambda [p;1]^T.  The accumulators are held COLUMN-major, acc[c * n + r] = R[r][c] (so is
// the R / G scratch until k_r_to_g): in 2-D a column is a register pair updated by one FFMA2.
template <int NDIM, typename T>
__device__ __forceinline__ void accumulate_outer(T* acc, const T* lam, const T* pn)
{
    if constexpr (UsePacked<NDIM, T>::value) {
        const F2 L = pk(lam[0], lam[1]);
        F2 a0 = fma2(L, bc(pn[0]), pk(acc[0], acc[1]));
        F2 a1 = fma2(L, bc(pn[1]), pk(acc[2], acc[3]));
        F2 a2 = add2(pk(acc[4], acc[5]), L);
        unpk(a0, acc[0], acc[1]);
        unpk(a1, acc[2], acc[3]);
        unpk(a2, acc[4], acc[5]);
    } else {
#pragma unroll
        for (int r = 0; r < NDIM; ++r) {
#pragma unroll
            for (int cc = 0; cc < NDIM; ++cc) acc[cc * NDIM + r] = Num<T>::fma(lam[r], pn[cc], acc[cc * NDIM + r]);
            acc[NDIM * NDIM + r] += lam[r];
        }
    }
}

// Final flush of the per-thread accumulators: the lanes of a warp are neighbouring points, so they
// mostly end in the same one to three cells.  Runs of equal cell index are summed with a segmented
// shuffle scan and only the last lane of each run issues the (native, fire-and-forget) global
// reductions.  Must be called by all 32 lanes; lanes without a trajectory pass key = -1.
template <typename T, int PPC>
__device__ __forceinline__ void flush_runs(T* __restrict__ Gg, int key, T* acc)
{
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int prev = __shfl_up_sync(full, key, 1);
    const unsigned heads = __ballot_sync(full, lane == 0 || prev != key);
    const unsigned upto = heads & (full >> (31 - lane));           // heads at or below this lane
    const int start = 31 - __clz(upto);
    const bool tail = lane == 31 || ((heads >> (lane + 1)) & 1u);  // next lane starts a new run
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
#pragma unroll
        for (int e = 0; e < PPC; ++e) {
            const T t = __shfl_up_sync(full, acc[e], off);
            if (lane - off >= start) acc[e] += t;
        }
    }
    if (tail && key >= 0) red_cell<PPC>(Gg + (size_t)key * PPC, acc);
}

#ifndef CPAB_BWD_REGS
#define CPAB_BWD_REGS 80
#endif
// resident CTAs per SM the register allocation is tuned for (float: 80 regs in 1-D/2-D, 128 in 3-D -- fewer registers spill)
template <typename T, int NDIM, int SEG, int BLOCK> struct BwdOcc {
    static constexpr int kRegs = NDIM == 3 ? (SEG <= 3 ? 102 : 128) : CPAB_BWD_REGS;
    static constexpr int kMinBlocks = sizeof(T) == 8 ? 1 : 65536 / (BLOCK * kRegs);
};

// SAMPLE: `gout` holds the transformed grid (output of the forward) and the upstream gradient is
// that of the sampled image, `gimg`; lambda_N is formed in the prologue (fused transform_data).
template <typename T, int NDIM, int SEG, bool SMEM, int BLOCK, bool SAMPLE>
__global__ void __launch_bounds__(BLOCK, (BwdOcc<T, NDIM, SEG, BLOCK>::kMinBlocks))
k_backward(const T* __restrict__ points, const T* __restrict__ Ws, const T* __restrict__ gout,
           T* __restrict__ G, T* __restrict__ dpoints, long nP, int broadcast, int nsteps,
           const __grid_constant__ Geom g, const __grid_constant__ WorkPlan wp,
           const T* __restrict__ data, const T* __restrict__ gimg, const __grid_constant__ Shape sh)
{
    constexpr int PPC = Dim<NDIM>::kPpc;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned s_work;
    constexpr int WS = StepRec<NDIM>::kStride;
    const int tsize = g.n_cells * PPC;
    const int wsize = g.n_cells * WS;
    const int nseg = (nsteps + SEG - 1) / SEG;

    // shared layout: [step records] (if SMEM), checkpoints [nseg][NDIM][BLOCK], cell trace
    // [nsteps][BLOCK] (16-bit; 32-bit only for tessellations of >= 65536 simplices, which never
    // fit the staged path)
    T* sW = reinterpret_cast<T*>(smem_raw);
    T* ck = sW + (SMEM ? wsize : 0);
    unsigned short* ct16 = reinterpret_cast<unsigned short*>(ck + (size_t)nseg * NDIM * BLOCK);
    int* ct32 = reinterpret_cast<int*>(ct16);
    const bool wide = !SMEM && g.n_cells > 65535;
    CellTable<T, NDIM, SMEM, WS> tab;
    tab.saddr = SMEM ? (uint32_t)__cvta_generic_to_shared(sW) & 0xffffffu : 0;   // CTA-local offset (no cluster launch: rank bits are 0)
    int staged = -1;
    WorkUnit wu;
    while (next_unit(wp, nP, &s_work, wu)) {
    const int theta = wu.theta;
    const long begin = wu.begin, end = wu.end;
    tab.gptr = Ws + (size_t)theta * wsize;
    if (SMEM && theta != staged) {           // (next_unit synchronised: nobody reads the old table any more)
        stage_block(sW, tab.gptr, wsize);
        staged = theta;
        __syncthreads();
    }
    T* Gg = G + (size_t)theta * tsize;
    // keep the base in registers: the flush blocks run divergently, often, and would otherwise
    // rebuild it from the kernel parameters (11 uniform-datapath instructions per occurrence)
    asm volatile("" : "+l"(Gg));
    const T* src = points + (broadcast ? (size_t)theta * NDIM * nP : 0);
    const T* gsrc = gout + (size_t)theta * NDIM * nP;

    for (long base = begin; base < end; base += BLOCK) {      // warp-uniform trip count
        const long i = base + threadIdx.x;
        const bool valid = i < end;
        T acc[PPC];
        int cur = -1;
#pragma unroll
        for (int e = 0; e < PPC; ++e) acc[e] = 0;
        if (valid) {
            T p[NDIM], lam[NDIM];
#pragma unroll
            for (int j = 0; j < NDIM; ++j) { p[j] = src[i + (long)j * nP]; lam[j] = gsrc[i + (long)j * nP]; }
            if (SAMPLE) {       // gsrc is the transformed grid: turn it into dL/d(grid_t)
                T pt[NDIM];
#pragma unroll
                for (int j = 0; j < NDIM; ++j) pt[j] = lam[j];
                sample_vjp<T, NDIM>(pt, theta, i, data, gimg, sh, lam);
            }

            // ---- pass 1: the RK2 trajectory.  Records the cell of every step and a checkpoint
            //      of p at the start of every segment; the only pass that searches cells.
            //      (FULL = a whole segment that is not the last one: no bounds checks.)
            auto pass1 = [&](int sg, auto full_tag) {
                constexpr bool FULL = decltype(full_tag)::value;
#pragma unroll
                for (int j = 0; j < NDIM; ++j) ck[(sg * NDIM + j) * BLOCK + threadIdx.x] = p[j];
#pragma unroll
                for (int s = 0; s < SEG; ++s) {
                    const int n = sg * SEG + s;
                    if (FULL || n < nsteps) {
                        const int c = find_cell<NDIM>(p, g);
                        if (wide) ct32[n * BLOCK + threadIdx.x] = c;
                        else ct16[n * BLOCK + threadIdx.x] = (unsigned short)c;
                        if (FULL || n + 1 < nsteps) {
                            T w[WS];
                            tab.load(c, w);
                            step_inc<NDIM>(w, p);
                        }
                    }
                }
            };
            for (int sg = 0; sg + 1 < nseg; ++sg) pass1(sg, std::true_type{});
            pass1(nseg - 1, std::false_type{});

            // ---- pass 2: segments in reverse; replay p into registers (no search), sweep back
            auto pass2 = [&](int sg, auto full_tag) {
                constexpr bool FULL = decltype(full_tag)::value;
                const int len = FULL ? SEG : nsteps - sg * SEG;
                T ps[SEG][NDIM];
                T w[WS];
                int cs[SEG];
#pragma unroll
                for (int j = 0; j < NDIM; ++j) p[j] = ck[(sg * NDIM + j) * BLOCK + threadIdx.x];
#pragma unroll
                for (int s = 0; s < SEG; ++s) {
                    if (FULL || s < len) {
                        const int n = sg * SEG + s;
                        cs[s] = wide ? ct32[n * BLOCK + threadIdx.x] : (int)ct16[n * BLOCK + threadIdx.x];
#pragma unroll
                        for (int j = 0; j < NDIM; ++j) ps[s][j] = p[j];
                        if (s + 1 < SEG && (FULL || s + 1 < len)) {
                            tab.load(cs[s], w);       // (the compiler keeps these for the sweep below)
                            step_inc<NDIM>(w, p);
                        }
                    }
                }
#pragma unroll
                for (int s = SEG - 1; s >= 0; --s) {
                    if (FULL || s < len) {
                        const int c = cs[s];
                        tab.load(c, w);
                        if (c != cur) {                 // left a cell: hand its sum to R[theta]
                            if (cur >= 0) red_cell<PPC>(Gg + (size_t)cur * PPC, acc);
#pragma unroll
                            for (int e = 0; e < PPC; ++e) acc[e] = 0;
                            cur = c;
                        }
                        // R_c += lambda_{n+1} [p_n;1]^T
                        accumulate_outer<NDIM>(acc, lam, ps[s]);
                        // lambda_n = M^T lambda_{n+1} = lambda_{n+1} + D^T lambda_{n+1}
                        T nl[NDIM];
#pragma unroll
                        for (int r = 0; r < NDIM; ++r) {
                            T t = lam[r];
#pragma unroll
                            for (int j = 0; j < NDIM; ++j) t = Num<T>::fma(w[StepRec<NDIM>::d(j, r)], lam[j], t);
                            nl[r] = t;
                        }
#pragma unroll
                        for (int r = 0; r < NDIM; ++r) lam[r] = nl[r];
                    }
                }
            };
            pass2(nseg - 1, std::false_type{});
            for (int sg = nseg - 2; sg >= 0; --sg) pass2(sg, std::true_type{});
            if (dpoints != nullptr) {
                T* dp = dpoints + (size_t)theta * NDIM * nP;
#pragma unroll
                for (int j = 0; j < NDIM; ++j) dp[i + (long)j * nP] = lam[j];
            }
        }
        flush_runs<T, PPC>(Gg, cur, acc);
    }
    }   // work units
    leave_grid(wp);
}

// Per (theta, cell): the RK2 step record (see step_inc) from A_c = [L | t], and a zeroed R_c block.
// Evaluated in double and rounded once.  The origin is the centre of the cell's square / cube (any
// float near the cell serves: s' is formed from the rounded value).
template <typename T, int NDIM>
__global__ void __launch_bounds__(256)
k_prepare_backward(const T* __restrict__ As, T* __restrict__ Ws, T* __restrict__ R, long n_blocks, int nsteps,
                   const __grid_constant__ Geom g)
{
    constexpr int PPC = Dim<NDIM>::kPpc;
    constexpr int M = NDIM + 1;
    constexpr int WS = StepRec<NDIM>::kStride;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_blocks) return;
    double A[PPC];
#pragma unroll
    for (int e = 0; e < PPC; ++e) { A[e] = (double)As[i * PPC + e]; R[i * PPC + e] = 0; }
    int box = (int)(i % g.n_cells) / (NDIM == 1 ? 1 : NDIM == 2 ? 4 : 5);
    T o[NDIM];
#pragma unroll
    for (int j = 0; j < NDIM; ++j) {
        const int k = box % g.nc[j];
        box /= g.nc[j];
        o[j] = StepRec<NDIM>::kLocal ? (T)((k + 0.5) / g.nc[j]) : (T)0;
    }
    const double h = 1.0 / nsteps, h2 = 0.5 / nsteps / nsteps;
    T* W = Ws + i * WS;
#pragma unroll
    for (int r = 0; r < NDIM; ++r) {
        double sp = 0;
#pragma unroll
        for (int cc = 0; cc < M; ++cc) {                       // (L Atilde)[r][cc] = sum_k L[r][k] A[k][cc]
            double t = 0;
#pragma unroll
            for (int k = 0; k < NDIM; ++k) t = ::fma(A[r * M + k], A[k * M + cc], t);
            const double v = ::fma(h2, t, h * A[r * M + cc]);
            if (cc < NDIM) { W[StepRec<NDIM>::d(r, cc)] = (T)v; sp = ::fma(v, (double)o[cc], sp); }
            else W[StepRec<NDIM>::kS + r] = (T)(v + sp);
        }
        if (StepRec<NDIM>::kLocal) W[StepRec<NDIM>::kO + r] = o[r];
    }
    if (StepRec<NDIM>::kLocal) {
#pragma unroll
        for (int e = StepRec<NDIM>::kO + NDIM; e < WS; ++e) W[e] = 0;
    }
}

// R -> G, per (theta, cell), in place:  G_c = h R_c + (h^2/2) (R_c Atilde^T + L^T R_c),
// Atilde = [[L, t], [0, 0]].  (sum over the steps spent in cell c of
// h lambda [pMid;1]^T + (h^2/2) (L^T lambda) [p;1]^T with [pMid;1] = (I + (h/2) Atilde) [p;1].)
template <typename T, int NDIM>
__global__ void __launch_bounds__(256)
k_r_to_g(T* __restrict__ RG, const T* __restrict__ As, long n_blocks, int nsteps)
{
    constexpr int PPC = Dim<NDIM>::kPpc;
    constexpr int M = NDIM + 1;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_blocks) return;
    T R[PPC], A[PPC], Gc[PPC];      // R arrives column-major (accumulate_outer), G leaves row-major
#pragma unroll
    for (int r = 0; r < NDIM; ++r) {
#pragma unroll
        for (int cc = 0; cc < M; ++cc) R[r * M + cc] = RG[i * PPC + cc * NDIM + r];
    }
#pragma unroll
    for (int e = 0; e < PPC; ++e) A[e] = As[i * PPC + e];
    const T h = (T)(1.0 / nsteps), h2 = (T)(0.5 / nsteps / nsteps);
#pragma unroll
    for (int r = 0; r < NDIM; ++r) {
#pragma unroll
        for (int cc = 0; cc < M; ++cc) {
            T t = 0;
            if (cc < NDIM) {                                   // (R Atilde^T)[r][cc] = sum_k R[r][k] A[cc][k]
#pragma unroll
                for (int k = 0; k < M; ++k) t = Num<T>::fma(R[r * M + k], A[cc * M + k], t);
            }
#pragma unroll
            for (int j = 0; j < NDIM; ++j) t = Num<T>::fma(A[j * M + r], R[j * M + cc], t);   // (L^T R)[r][cc]
            Gc[r * M + cc] = Num<T>::fma(h2, t, h * R[r * M + cc]);
        }
    }
#pragma unroll
    for (int e = 0; e < PPC; ++e) RG[i * PPC + e] = Gc[e];
}

// dtheta[t][k] = sum_e G[t][e] * B[e][k]      (G [n_theta,D], B [D,d] row-major, dtheta [n_theta,d])
// A skinny GEMM (d ~ 50-250, D up to a few thousand, n_theta from 16 to 65536): grid = theta tiles
// of TT x chunks of `ek` rows of B.  With few thetas the work is split along e (split-K) so that
// the chip is filled -- 16 thetas x D = 3840 ran on 4 CTAs for 0.57 ms before -- and the partial
// sums are reduced into a zeroed dtheta with native float atomics.  Four independent loads of B
// are in flight per thread; G tiles are read through shared memory.
template <typename T, int TT>
__global__ void __launch_bounds__(128)
k_grad_epilogue(const T* __restrict__ G, const T* __restrict__ B, T* __restrict__ dtheta,
                int n_theta, int D, int d, int ek, int split)
{
    constexpr int TILE = 256;
    __shared__ T sG[TT][TILE];
    const int t0 = blockIdx.x * TT;
    const int e0 = blockIdx.y * ek;
    const int n = D - e0 < ek ? D - e0 : ek;
    for (int x = threadIdx.x; x < TT * TILE; x += blockDim.x) {
        const int u = x / TILE, e = x - u * TILE;
        sG[u][e] = (t0 + u < n_theta && e < n) ? G[(size_t)(t0 + u) * D + e0 + e] : (T)0;
    }
    __syncthreads();
    const T* Bp = B + (size_t)e0 * d;
    for (int k = threadIdx.x; k < d; k += blockDim.x) {
        T acc[TT];
#pragma unroll
        for (int u = 0; u < TT; ++u) acc[u] = 0;
        int e = 0;
        for (; e + 4 <= n; e += 4) {
            T bv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = __ldg(Bp + (size_t)(e + j) * d + k);
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int u = 0; u < TT; ++u) acc[u] = Num<T>::fma(sG[u][e + j], bv[j], acc[u]);
        }
        for (; e < n; ++e) {
            const T bv = __ldg(Bp + (size_t)e * d + k);
#pragma unroll
            for (int u = 0; u < TT; ++u) acc[u] = Num<T>::fma(sG[u][e], bv, acc[u]);
        }
#pragma unroll
        for (int u = 0; u < TT; ++u) {
            if (t0 + u < n_theta) {
                T* dst = dtheta + (size_t)(t0 + u) * d + k;
                if (split) Num<T>::atomic_add(dst, acc[u]); else *dst = acc[u];
            }
        }
    }
}

// =====================================================================================================
// host launchers
// =====================================================================================================
static int g_tune_fwd_ppt = 1;        // points advanced concurrently per thread in k_forward
static int g_tune_chunk_auto = 1;     // 1: cut chunks finer when the grid would not fill the chip
static int g_tune_chunk_pts = 1024;   // points of one theta handled by one CTA
static int g_tune_bwd_seg = 0;        // 0 = auto (5 in 1-D/2-D, 3 in 3-D: fits 96 registers -> 5 CTAs/SM)       // checkpoint spacing of k_backward
static int g_tune_bwd_block = 128;
static int g_tune_bwd_stage = -1;     // 1: stage A[theta] in shared memory, 0: read it through L1, -1 = auto (0 in 3-D)

int set_tuning(const char* key, int value)
{
    const std::string_view k(key);
    if (k == "fwd_ppt" && (value == 1 || value == 2)) { g_tune_fwd_ppt = value; return kOk; }
    if (k == "chunk_pts" && value >= 256 && value % 256 == 0) { g_tune_chunk_pts = value; g_tune_chunk_auto = 0; return kOk; }
    if (k == "chunk_auto" && (value == 0 || value == 1)) { g_tune_chunk_auto = value; return kOk; }
    if (k == "bwd_seg" && (value == 0 || value == 3 || value == 5 || value == 10)) { g_tune_bwd_seg = value; return kOk; }
    if (k == "bwd_stage" && value >= -1 && value <= 1) { g_tune_bwd_stage = value; return kOk; }
    if (k == "bwd_block" && (value == 64 || value == 128 || value == 256)) { g_tune_bwd_block = value; return kOk; }
    if (k == "interp_variant" && value >= 0 && value <= 4) { set_interp_variant(value); return kOk; }
    set_error("unknown tuning key/value %s=%d", key, value);
    return kErrArgument;
}

template <typename T, int NDIM>
static int findcellidx_t(const Geom& g, const void* points, long nP, int* out, cudaStream_t st)
{
    if (nP == 0) return kOk;
    const unsigned blocks = (unsigned)((nP + 255) / 256);
    k_findcellidx<T, NDIM><<<blocks, 256, 0, st>>>((const T*)points, nP, out, g);
    count_launch();
    CPAB_CUDA_OK(cudaGetLastError());
    return kOk;
}

int launch_findcellidx(int dtype, const Geom& g, const void* points, long nP, int* out, cudaStream_t st)
{
#define GO(T) CPAB_DISPATCH((findcellidx_t<T, 1>(g, points, nP, out, st)), (findcellidx_t<T, 2>(g, points, nP, out, st)), (findcellidx_t<T, 3>(g, points, nP, out, st)))
    return CPAB_DTYPE(GO(float), GO(double));
#undef GO
}

static int sm_count()
{
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    return sms;
}

// Work plan of one launch (see WorkPlan).  Bulk units of `chunk_pts` points (1024 by default:
// within 1 % of the best size on every BASELINE shape, profiles/r01b_chunk_sweep.txt -- smaller
// units concentrate the resident CTAs on few thetas and the G reductions of one theta collide in
// L2, larger ones leave a longer tail), then about two small units per resident CTA.  Problems
// that cannot fill the chip with bulk units are cut into small ones altogether.
static std::atomic<unsigned> g_plan_seq{0};
static WorkPlan plan_work(long nP, int n_theta, int block, int ctas_per_sm, unsigned& grid, bool counter = true)
{
    WorkPlan wp;
    const int unit = block > 256 ? block : 256;
    const long slots = (long)sm_count() * (ctas_per_sm > 0 ? ctas_per_sm : 1);
    wp.bulk_pts = g_tune_chunk_pts;
    wp.small_pts = unit;
    long small_per_theta = 0;
    if (g_tune_chunk_auto) {
        const long all_small = (nP + unit - 1) / unit;
        // the tail pays for itself only when units are drawn from the counter (measured: +6 % for
        // k_backward on configs[1], nothing for the static grid of k_forward), and not when there
        // are so many thetas that one small unit each is already a large share of the work
        if (counter && (long)n_theta <= 2 * slots) small_per_theta = (2 * slots + n_theta - 1) / n_theta;
        if (small_per_theta > all_small || (long)n_theta * ((nP + wp.bulk_pts - 1) / wp.bulk_pts) < slots)
            small_per_theta = all_small;
    }
    long tail_pts = small_per_theta * unit;
    if (tail_pts > nP) tail_pts = nP;
    wp.nP_bulk = (nP - tail_pts) / unit * unit;
    wp.small_per_theta = (int)((nP - wp.nP_bulk + unit - 1) / unit);
    wp.bulk_per_theta = (int)((wp.nP_bulk + wp.bulk_pts - 1) / wp.bulk_pts);
    wp.total_bulk = (unsigned)((long)n_theta * wp.bulk_per_theta);
    wp.total = wp.total_bulk + (unsigned)((long)n_theta * wp.small_per_theta);
    wp.slot = counter ? g_plan_seq.fetch_add(1, std::memory_order_relaxed) % kWorkRing : 0;
    grid = (unsigned)((long)wp.total < slots ? (long)wp.total : slots);
    return wp;
}

struct SampleArgs {       // fused transform_data: images to sample from / to, their geometry
    const void* data = nullptr;
    void* img = nullptr;          // forward: sampled output image
    const void* gimg = nullptr;   // backward: upstream gradient of the sampled image
    Shape sh{};
};

template <typename T, int NDIM, bool STRICT, bool SMEM, int PPT, bool SAMPLE = false>
static int forward_launch(const Geom& g, int nsteps, int n_theta, long nP, int broadcast,
                          const void* points, const void* trels, void* out, cudaStream_t st,
                          const SampleArgs& sa = SampleArgs())
{
    const size_t smem = SMEM ? (size_t)g.n_cells * FwdRec<T, NDIM, SMEM>::kStride * sizeof(T) : 0;
    auto kern = k_forward<T, NDIM, STRICT, SMEM, PPT, SAMPLE>;
    if (smem > 48 * 1024)
        CPAB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem);
    if ((long long)n_theta * ((nP + 255) / 256) > 0x7fffffffLL) { set_error("grid too large"); return kErrUnsupported; }
    unsigned slots_grid = 0;
    const WorkPlan wp = plan_work(nP, n_theta, 256, per_sm, slots_grid, false);
    prof_begin(kProfForward, st);
    kern<<<wp.total, 256, smem, st>>>((const T*)points, (const T*)trels, (T*)out, nP,
                                      broadcast, nsteps, g, wp,
                                      (const T*)sa.data, (T*)sa.img, sa.sh);
    prof_end(kProfForward, st);
    count_launch();
    CPAB_CUDA_OK(cudaGetLastError());
    return kOk;
}

template <typename T, int NDIM>
static int forward_t(int flags, const Geom& g, int nsteps, int n_theta, long nP, int broadcast,
                     const void* points, const void* trels, void* out, cudaStream_t st,
                     const SampleArgs* sa = nullptr)
{
    const bool smem = (size_t)g.n_cells * FwdRec<T, NDIM, true>::kStride * sizeof(T) <= 160 * 1024;
    const bool strict = !(flags & kFlagFastMath);
    const int ppt = g_tune_fwd_ppt;
    if (sa != nullptr) {       // fused sampling epilogue
#define SARGS g, nsteps, n_theta, nP, broadcast, points, trels, out, st, *sa
        if (smem) return strict ? forward_launch<T, NDIM, true, true, 1, true>(SARGS) : forward_launch<T, NDIM, false, true, 1, true>(SARGS);
        return strict ? forward_launch<T, NDIM, true, false, 1, true>(SARGS) : forward_launch<T, NDIM, false, false, 1, true>(SARGS);
#undef SARGS
    }
#define ARGS g, nsteps, n_theta, nP, broadcast, points, trels, out, st
    if (smem) {
        if (strict) return ppt == 2 ? forward_launch<T, NDIM, true, true, 2>(ARGS) : forward_launch<T, NDIM, true, true, 1>(ARGS);
        return ppt == 2 ? forward_launch<T, NDIM, false, true, 2>(ARGS) : forward_launch<T, NDIM, false, true, 1>(ARGS);
    }
    if (strict) return forward_launch<T, NDIM, true, false, 1>(ARGS);
    return forward_launch<T, NDIM, false, false, 1>(ARGS);
#undef ARGS
}

int launch_forward(int dtype, int flags, const Geom& g, int nsteps, int n_theta, long nP,
                   int broadcast, const void* points, const void* trels, void* out, cudaStream_t st)
{
    if (n_theta == 0 || nP == 0) return kOk;
#define GO(T) CPAB_DISPATCH((forward_t<T, 1>(flags, g, nsteps, n_theta, nP, broadcast, points, trels, out, st)), \
                            (forward_t<T, 2>(flags, g, nsteps, n_theta, nP, broadcast, points, trels, out, st)), \
                            (forward_t<T, 3>(flags, g, nsteps, n_theta, nP, broadcast, points, trels, out, st)))
    return CPAB_DTYPE(GO(float), GO(double));
#undef GO
}

template <typename T, int NDIM>
static int jacobian_t(const Geom& g, int nsteps, int n_theta, int d, long nP, int broadcast,
                      const void* points, const void* As, const void* Bs, void* jac, cudaStream_t st)
{
    const long long tk = (long long)n_theta * d;
    if (tk > 65535) { set_error("jacobian: n_theta*d = %lld exceeds 65535 (use backward_theta)", tk); return kErrUnsupported; }
    dim3 grid((unsigned)((nP + 127) / 128), (unsigned)tk);
    k_jacobian<T, NDIM><<<grid, 128, 0, st>>>((const T*)points, (const T*)As, (const T*)Bs, (T*)jac,
                                               nP, n_theta, d, broadcast, nsteps, g);
    count_launch();
    CPAB_CUDA_OK(cudaGetLastError());
    return kOk;
}

int launch_jacobian(int dtype, const Geom& g, int nsteps, int n_theta, int d, long nP, int broadcast,
                    const void* points, const void* As, const void* Bs, void* jac, cudaStream_t st)
{
    if (n_theta == 0 || nP == 0 || d == 0) return kOk;
#define GO(T) CPAB_DISPATCH((jacobian_t<T, 1>(g, nsteps, n_theta, d, nP, broadcast, points, As, Bs, jac, st)), \
                            (jacobian_t<T, 2>(g, nsteps, n_theta, d, nP, broadcast, points, As, Bs, jac, st)), \
                            (jacobian_t<T, 3>(g, nsteps, n_theta, d, nP, broadcast, points, As, Bs, jac, st)))
    return CPAB_DTYPE(GO(float), GO(double));
#undef GO
}

// G [n_theta, D] (accumulated as R, converted in place) followed by the RK2 step table W [n_theta, D]
size_t backward_g_bytes(int dtype, const Geom& g, int n_theta)
{
    const size_t elt = dtype == kF32 ? 4 : 8;
    return (size_t)n_theta * g.n_cells * g.ndim * (g.ndim + 1) * elt;
}
size_t backward_workspace_bytes(int dtype, const Geom& g, int n_theta)
{
    const size_t elt = dtype == kF32 ? 4 : 8;
    const int stride = g.ndim == 1 ? StepRec<1>::kStride : g.ndim == 2 ? StepRec<2>::kStride : StepRec<3>::kStride;
    return ((backward_g_bytes(dtype, g, n_theta) + 15) & ~(size_t)15) + (size_t)n_theta * g.n_cells * stride * elt;
}

template <typename T, int NDIM, int SEG, bool SMEM, int BLOCK>
static int backward_launch(const Geom& g, int nsteps, int n_theta, long nP, int broadcast,
                           const void* points, const void* Ws, const void* gout, void* G,
                           void* dpoints, cudaStream_t st, bool& fits, const SampleArgs* sa)
{
    const int nseg = (nsteps + SEG - 1) / SEG;
    const size_t tbytes = (size_t)g.n_cells * StepRec<NDIM>::kStride * sizeof(T);
    const size_t smem = (SMEM ? tbytes : 0) + (size_t)nseg * NDIM * BLOCK * sizeof(T) +
                        (size_t)nsteps * BLOCK * (g.n_cells > 65535 ? 4 : 2);
    fits = smem <= kMaxSmemBytes;
    if (!fits) return kOk;
    auto launch = [&](auto kern, const SampleArgs& a) -> int {
        if (smem > 48 * 1024)
            CPAB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BLOCK, smem);
        if ((long long)n_theta * ((nP + 255) / 256) > 0x7fffffffLL) { set_error("grid too large"); return kErrUnsupported; }
        unsigned blocks = 0;
        const WorkPlan wp = plan_work(nP, n_theta, BLOCK, per_sm, blocks);
        prof_begin(kProfBackward, st);
        kern<<<blocks, BLOCK, smem, st>>>((const T*)points, (const T*)Ws, (const T*)gout, (T*)G,
                                          (T*)dpoints, nP, broadcast, nsteps, g, wp,
                                          (const T*)a.data, (const T*)a.gimg, a.sh);
        prof_end(kProfBackward, st);
        count_launch();
        return kOk;
    };
    const int rc = sa != nullptr ? launch(k_backward<T, NDIM, SEG, SMEM, BLOCK, true>, *sa)
                                 : launch(k_backward<T, NDIM, SEG, SMEM, BLOCK, false>, SampleArgs());
    if (rc != kOk) return rc;
    CPAB_CUDA_OK(cudaGetLastError());
    return kOk;
}

template <typename T, int NDIM>
static int backward_t(const Geom& g, int nsteps, int n_theta, int d, long nP, int broadcast,
                      const void* points, const void* As, const void* basis, const void* gout,
                      void* dtheta, void* dpoints, void* ws, cudaStream_t st, const SampleArgs* sa = nullptr)
{
    const int D = g.n_cells * Dim<NDIM>::kPpc;
    const long n_blocks = (long)n_theta * g.n_cells;
    // second part of the workspace (16-byte aligned): RK2 step records
    T* Ws = reinterpret_cast<T*>(reinterpret_cast<char*>(ws) + (((size_t)n_theta * D * sizeof(T) + 15) & ~(size_t)15));
    k_prepare_backward<T, NDIM><<<(unsigned)((n_blocks + 255) / 256), 256, 0, st>>>((const T*)As, Ws, (T*)ws, n_blocks, nsteps, g);
    CPAB_CUDA_OK(cudaGetLastError());
    count_launch();
    bool fits = nP == 0;      // nothing to integrate: G stays zero, the epilogue writes dtheta = 0
    int rc = kOk;
#define TRY(SEG, SMEM, BLOCK)                                                                      \
    if (!fits && rc == kOk)                                                                        \
        rc = backward_launch<T, NDIM, SEG, SMEM, BLOCK>(g, nsteps, n_theta, nP, broadcast, points, \
                                                        Ws, gout, ws, dpoints, st, fits, sa)
    // preferred configuration first, then progressively smaller shared-memory footprints
    // measured (profiles/): 3-D runs best with 3-step segments (96 registers, 5 CTAs/SM) and the
    // per-theta matrices read through L1 instead of staged (shared memory then holds only the
    // checkpoints and the cell trace); 1-D/2-D with 5-step segments and staged matrices
    const int seg = g_tune_bwd_seg != 0 ? g_tune_bwd_seg : (NDIM == 3 ? 3 : 5);
    const bool stage = g_tune_bwd_stage >= 0 ? g_tune_bwd_stage != 0 : NDIM != 3;
    if (seg == 3) {
        if (stage) { if (g_tune_bwd_block == 256) TRY(3, true, 256); TRY(3, true, 128); }
        else { if (g_tune_bwd_block == 256) TRY(3, false, 256); TRY(3, false, 128); }
    } else if (seg == 5) {
        if (stage) {
            if (g_tune_bwd_block == 256) TRY(5, true, 256);
            if (g_tune_bwd_block == 64) TRY(5, true, 64);
            TRY(5, true, 128);
        } else {
            if (g_tune_bwd_block == 256) TRY(5, false, 256);
            TRY(5, false, 128);
        }
    } else {
        if (g_tune_bwd_block == 256) TRY(10, true, 256);
        if (g_tune_bwd_block == 64) TRY(10, true, 64);
        TRY(10, true, 128);
    }
    TRY(10, true, 64);
    TRY(10, false, 128);
    TRY(10, false, 64);
#undef TRY
    if (rc != kOk) return rc;
    if (!fits) {
        set_error("backward: nstepsolver=%d needs more checkpoint memory than one CTA has", nsteps);
        return kErrUnsupported;
    }
    {
        k_r_to_g<T, NDIM><<<(unsigned)((n_blocks + 255) / 256), 256, 0, st>>>((T*)ws, (const T*)As, n_blocks, nsteps);
        CPAB_CUDA_OK(cudaGetLastError());
        count_launch();
    }
    return launch_grad_epilogue(sizeof(T) == 4 ? kF32 : kF64, ws, basis, dtheta, n_theta, D, d, st);
}

template <typename T>
static int grad_epilogue_t(const void* G, const void* basis, void* dtheta, int n_theta, int D, int d, cudaStream_t st)
{
    // theta tile and e-chunk: as large as possible (B is re-read once per theta tile) while the
    // grid still covers the chip about twice
    const long want = 2L * sm_count();
    int tt = 8, ek = 256;
    auto ctas = [&]() { return (long)((n_theta + tt - 1) / tt) * ((D + ek - 1) / ek); };
    while (tt > 1 && ctas() < want) tt /= 2;
    while (ek > 32 && ctas() < want) ek /= 2;
    const int split = (D + ek - 1) / ek > 1;
    if (split) CPAB_CUDA_OK(cudaMemsetAsync(dtheta, 0, (size_t)n_theta * d * sizeof(T), st));
    dim3 grid((unsigned)((n_theta + tt - 1) / tt), (unsigned)((D + ek - 1) / ek));
    prof_begin(kProfEpilogue, st);
#define EPI(TT) k_grad_epilogue<T, TT><<<grid, 128, 0, st>>>((const T*)G, (const T*)basis, (T*)dtheta, n_theta, D, d, ek, split)
    if (tt == 8) EPI(8); else if (tt == 4) EPI(4); else if (tt == 2) EPI(2); else EPI(1);
#undef EPI
    prof_end(kProfEpilogue, st);
    count_launch();
    CPAB_CUDA_OK(cudaGetLastError());
    return kOk;
}

int launch_grad_epilogue(int dtype, const void* G, const void* basis, void* dtheta, int n_theta, int D,
                         int d, cudaStream_t st)
{
    if (n_theta == 0 || d == 0) return kOk;
    if ((long)((n_theta + 7) / 8) > 0x7fffffffL || (D + 31) / 32 > 65535) { set_error("gradient epilogue: grid too large"); return kErrUnsupported; }
    return dtype == kF32 ? grad_epilogue_t<float>(G, basis, dtheta, n_theta, D, d, st)
                         : grad_epilogue_t<double>(G, basis, dtheta, n_theta, D, d, st);
}

int launch_backward(int dtype, int flags, const Geom& g, int nsteps, int n_theta, int d, long nP,
                    int broadcast, const void* points, const void* As, const void* basis,
                    const void* grad_out, void* dtheta, void* dpoints, void* workspace,
                    size_t workspace_bytes, cudaStream_t st)
{
    (void)flags;
    if (n_theta == 0 || d == 0) return kOk;
    if (workspace_bytes < backward_workspace_bytes(dtype, g, n_theta)) {
        set_error("backward: workspace has %zu bytes, needs %zu", workspace_bytes,
                  backward_workspace_bytes(dtype, g, n_theta));
        return kErrWorkspace;
    }
    if (reinterpret_cast<uintptr_t>(workspace) & 15) { set_error("backward: workspace must be 16-byte aligned"); return kErrArgument; }
#define GO(T) CPAB_DISPATCH((backward_t<T, 1>(g, nsteps, n_theta, d, nP, broadcast, points, As, basis, grad_out, dtheta, dpoints, workspace, st)), \
                            (backward_t<T, 2>(g, nsteps, n_theta, d, nP, broadcast, points, As, basis, grad_out, dtheta, dpoints, workspace, st)), \
                            (backward_t<T, 3>(g, nsteps, n_theta, d, nP, broadcast, points, As, basis, grad_out, dtheta, dpoints, workspace, st)))
    return CPAB_DTYPE(GO(float), GO(double));
#undef GO
}

// ---- fused transform_data --------------------------------------------------------------------------
static bool make_sample_shape(int ndim, int N, int C, const int* in_size, const int* out_size, Shape& sh, long& nP)
{
    sh.N = N; sh.C = C;
    long long gridpts = 1, inpts = C;
    for (int j = 0; j < 3; ++j) {
        sh.S[j] = j < ndim ? in_size[j] : 1;
        sh.O[j] = j < ndim ? out_size[j] : 1;
        if (j < ndim) { gridpts *= out_size[j]; inpts *= in_size[j]; }
    }
    nP = (long)gridpts;
    if (gridpts * ndim >= (1LL << 31) || inpts >= (1LL << 31) || gridpts * C >= (1LL << 31)) {
        set_error("transform_data: one sample exceeds 2^31 elements");
        return false;
    }
    return true;
}

int launch_transform_data_forward(int dtype, int flags, const Geom& g, int nsteps, int n_theta, int C,
                                  const int* in_size, const int* out_size, const void* points,
                                  const void* trels, const void* data, void* grid_t, void* img,
                                  cudaStream_t st)
{
    SampleArgs sa;
    long nP = 0;
    if (!make_sample_shape(g.ndim, n_theta, C, in_size, out_size, sa.sh, nP)) return kErrUnsupported;
    if (n_theta == 0 || nP == 0) return kOk;
    sa.data = data;
    sa.img = img;
#define GO(T) CPAB_DISPATCH((forward_t<T, 1>(flags, g, nsteps, n_theta, nP, 0, points, trels, grid_t, st, &sa)), \
                            (forward_t<T, 2>(flags, g, nsteps, n_theta, nP, 0, points, trels, grid_t, st, &sa)), \
                            (forward_t<T, 3>(flags, g, nsteps, n_theta, nP, 0, points, trels, grid_t, st, &sa)))
    return CPAB_DTYPE(GO(float), GO(double));
#undef GO
}

int launch_transform_data_backward(int dtype, const Geom& g, int nsteps, int n_theta, int d, int C,
                                   const int* in_size, const int* out_size, const void* points,
                                   const void* As, const void* basis, const void* data,
                                   const void* grid_t, const void* gimg, void* dtheta, void* workspace,
                                   size_t workspace_bytes, cudaStream_t st)
{
    SampleArgs sa;
    long nP = 0;
    if (!make_sample_shape(g.ndim, n_theta, C, in_size, out_size, sa.sh, nP)) return kErrUnsupported;
    if (n_theta == 0 || d == 0) return kOk;
    if (workspace_bytes < backward_workspace_bytes(dtype, g, n_theta)) {
        set_error("backward: workspace has %zu bytes, needs %zu", workspace_bytes, backward_workspace_bytes(dtype, g, n_theta));
        return kErrWorkspace;
    }
    if (reinterpret_cast<uintptr_t>(workspace) & 15) { set_error("backward: workspace must be 16-byte aligned"); return kErrArgument; }
    sa.data = data;
    sa.gimg = gimg;
#define GO(T) CPAB_DISPATCH((backward_t<T, 1>(g, nsteps, n_theta, d, nP, 0, points, As, basis, grid_t, dtheta, nullptr, workspace, st, &sa)), \
                            (backward_t<T, 2>(g, nsteps, n_theta, d, nP, 0, points, As, basis, grid_t, dtheta, nullptr, workspace, st, &sa)), \
                            (backward_t<T, 3>(g, nsteps, n_theta, d, nP, 0, points, As, basis, grid_t, dtheta, nullptr, workspace, st, &sa)))
    return CPAB_DTYPE(GO(float), GO(double));
#undef GO
}

}  // namespace cpab
