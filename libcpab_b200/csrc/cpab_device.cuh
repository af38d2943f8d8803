// cpab_device.cuh -- device helpers shared by the integration kernels (cpab_integrate.cu: forward,
// Jacobian, gradient epilogue) and the adjoint kernels (cpab_adjoint.cuh, instantiated per
// dimension in cpab_adjoint_{1,2,3}d.cu): rounding-controlled scalar ops, vector reductions,
// per-cell table fetches, the strict affine maps of the reference (libcpab/core/cpab_ops.cpp:192-222),
// the fused sampling epilogue / VJP prologue, and the work-unit plan of one launch.
#pragma once

#include <type_traits>

#include "cpab_common.cuh"
#include "cpab_sample.cuh"

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
// Work distribution.
//
// A unit is a range of points of one theta: `bulk_pts` points while plenty of work remains, and
// `small_pts` for the last ~2 units per resident CTA, so that every SM runs until the end -- with a
// static grid of equal CTAs the SMs of a B200 finished up to 14 % apart on BASELINE configs[1]
// (sm__cycles_active min/max 1169k/1354k; profiles/r01b_*), although every CTA does the same work.
// Units are numbered theta-major: resident CTAs spread over many thetas, which keeps the G
// reductions of one theta from colliding in L2.
// k_forward maps one CTA to one unit (static grid: there the per-unit barrier and bookkeeping cost
// more than the balance gained -- 2.21 vs 2.02 ms on 128 thetas x 512^2); k_backward runs a
// persistent grid whose CTAs draw units from a counter that lives in the CALLER's workspace
// (zeroed by k_prepare_backward of the same launch sequence: nothing is shared between launches,
// streams or captured graphs).
// =====================================================================================================
struct WorkPlan {
    unsigned total_bulk, total;      // units of the bulk phase / of both phases
    int bulk_per_theta, small_per_theta;
    int bulk_pts, small_pts;
    long nP_bulk;                    // points [0, nP_bulk) of every theta are bulk units, the rest small ones
};

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

// Per-thread tuning knobs (cpab_b200_set_tuning affects the calling thread only).
struct Tuning {
    int fwd_ppt = 1;         // points advanced concurrently per thread in k_forward
    int chunk_auto = 1;      // 1: cut chunks finer when the grid would not fill the chip
    int chunk_pts = 1024;    // points of one theta handled by one work unit
    int bwd_seg = 0;         // checkpoint spacing of k_backward; 0 = auto (5 in 1-D/2-D, 3 in 3-D)
    int bwd_block = 128;
    int bwd_stage = -1;      // 1: stage the step records in shared memory, 0: read them through L1, -1 = auto (0 in 3-D)
};
Tuning& tuning();            // cpab_integrate.cu (thread_local)
int sm_count();              // of the current device
WorkPlan plan_work(long nP, int n_theta, int block, int ctas_per_sm, unsigned& grid, bool counter);

struct SampleArgs {       // fused transform_data: images to sample from / to, their geometry
    const void* data = nullptr;
    void* img = nullptr;          // forward: sampled output image
    const void* gimg = nullptr;   // backward: upstream gradient of the sampled image
    Shape sh{};
};

}  // namespace cpab
