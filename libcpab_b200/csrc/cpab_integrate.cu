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

#include "cpab_device.cuh"


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


#ifndef CPAB_FWD_UNROLL
#define CPAB_FWD_UNROLL 2      // steps per trip of k_forward's inner loop (measured: profiles/r02_forward_unroll.txt)
#endif

namespace cpab {

constexpr int kFwdUnroll = CPAB_FWD_UNROLL;

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
#pragma unroll kFwdUnroll
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
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nP) return;
    const size_t tsize = (size_t)g.n_cells * PPC;
    const double h = 1.0 / nsteps;
    // blockIdx.y walks over (theta, k) pairs; more than 65535 of them are taken in strides
    for (long tk = blockIdx.y; tk < (long)n_theta * d; tk += gridDim.y) {
    const int theta = (int)(tk / d), k = (int)(tk - (long)theta * d);
    const T* At = As + (size_t)theta * tsize;
    const T* Bk = Bs + (size_t)k * tsize;
    const T* src = points + (broadcast ? (size_t)theta * NDIM * nP : 0);

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
Tuning& tuning()
{
    static thread_local Tuning t;
    return t;
}

int set_tuning(const char* key, int value)
{
    const std::string_view k(key);
    Tuning& t = tuning();
    if (k == "fwd_ppt" && (value == 1 || value == 2)) { t.fwd_ppt = value; return kOk; }
    if (k == "chunk_pts" && value >= 256 && value <= 32768 && value % 256 == 0) { t.chunk_pts = value; t.chunk_auto = 0; return kOk; }
    if (k == "chunk_auto" && (value == 0 || value == 1)) { t.chunk_auto = value; return kOk; }
    if (k == "bwd_seg" && (value == 0 || value == 3 || value == 5 || value == 10)) { t.bwd_seg = value; return kOk; }
    if (k == "bwd_stage" && value >= -1 && value <= 1) { t.bwd_stage = value; return kOk; }
    if (k == "bwd_block" && (value == 64 || value == 128 || value == 256)) { t.bwd_block = value; return kOk; }
    if (k == "interp_variant" && value >= 0 && value <= 15) { set_interp_variant(value); return kOk; }
    if (k == "interp_max_ctas" && value >= 0) { set_interp_max_ctas(value); return kOk; }
    if (k == "closed_refill" && (value == 0 || value == 1)) { set_closed_refill(value); return kOk; }
    if (k == "closed_stage" && (value == 0 || value == 1)) { set_closed_stage(value); return kOk; }
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

int sm_count()
{
    // per device: a process may drive several GPUs
    static std::atomic<int> cache[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    int sms = cache[dev].load(std::memory_order_relaxed);
    if (sms == 0) {
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
        cache[dev].store(sms, std::memory_order_relaxed);
    }
    return sms;
}

// Work plan of one launch (see WorkPlan).  Bulk units of `chunk_pts` points (1024 by default:
// within 1 % of the best size on every BASELINE shape, profiles/r01b_chunk_sweep.txt -- smaller
// units concentrate the resident CTAs on few thetas and the G reductions of one theta collide in
// L2, larger ones leave a longer tail), then about two small units per resident CTA.  Problems
// that cannot fill the chip with bulk units are cut into small ones altogether.
WorkPlan plan_work(long nP, int n_theta, int block, int ctas_per_sm, unsigned& grid, bool counter)
{
    WorkPlan wp;
    const Tuning& tn = tuning();
    const int unit = block > 256 ? block : 256;
    const long slots = (long)sm_count() * (ctas_per_sm > 0 ? ctas_per_sm : 1);
    wp.bulk_pts = tn.chunk_pts;
    wp.small_pts = unit;
    long small_per_theta = 0;
    if (tn.chunk_auto) {
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
    grid = (unsigned)((long)wp.total < slots ? (long)wp.total : slots);
    return wp;
}

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
    const int ppt = tuning().fwd_ppt;
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
    dim3 grid((unsigned)((nP + 127) / 128), (unsigned)(tk < 65535 ? tk : 65535));
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

// ---- adjoint gradient: per-dimension translation units (cpab_adjoint_{1,2,3}d.cu) ------------------
#define CPAB_DECL_DIM(N)                                                                             \
    int backward_dim_##N(int dtype, int flags, const Geom& g, int nsteps, int n_theta, int d, long nP, \
                         int broadcast, const void* points, const void* As, const void* basis,      \
                         const void* gout, void* dtheta, void* dpoints, void* ws, int* flagged,     \
                         cudaStream_t st, const SampleArgs* sa);                                    \
    int rk2_trace_dim_##N(const Geom& g, int nsteps, int n_theta, long nP, int broadcast, int mode, \
                          const void* points, const void* As, void* ws, int* cells,                 \
                          unsigned char* failed, cudaStream_t st);                                  \
    size_t backward_workspace_bytes_##N(size_t elt, const Geom& g, int n_theta, long nP);
CPAB_DECL_DIM(1)
CPAB_DECL_DIM(2)
CPAB_DECL_DIM(3)
#undef CPAB_DECL_DIM

size_t backward_g_bytes(int dtype, const Geom& g, int n_theta)
{
    const size_t elt = dtype == kF32 ? 4 : 8;
    return (size_t)n_theta * g.n_cells * g.ndim * (g.ndim + 1) * elt;
}
size_t backward_workspace_bytes(int dtype, const Geom& g, int n_theta, long nP)
{
    const size_t elt = dtype == kF32 ? 4 : 8;
    return CPAB_DISPATCH(backward_workspace_bytes_1(elt, g, n_theta, nP), backward_workspace_bytes_2(elt, g, n_theta, nP),
                         backward_workspace_bytes_3(elt, g, n_theta, nP));
}

static int backward_any(int dtype, int flags, const Geom& g, int nsteps, int n_theta, int d, long nP,
                        int broadcast, const void* points, const void* As, const void* basis,
                        const void* gout, void* dtheta, void* dpoints, void* ws, size_t ws_bytes,
                        int* flagged, cudaStream_t st, const SampleArgs* sa)
{
    if (ws_bytes < backward_workspace_bytes(dtype, g, n_theta, nP)) {
        set_error("backward: workspace has %zu bytes, needs %zu", ws_bytes, backward_workspace_bytes(dtype, g, n_theta, nP));
        return kErrWorkspace;
    }
    if (reinterpret_cast<uintptr_t>(ws) & 15) { set_error("backward: workspace must be 16-byte aligned"); return kErrArgument; }
#define ARGS dtype, flags, g, nsteps, n_theta, d, nP, broadcast, points, As, basis, gout, dtheta, dpoints, ws, flagged, st, sa
    return CPAB_DISPATCH(backward_dim_1(ARGS), backward_dim_2(ARGS), backward_dim_3(ARGS));
#undef ARGS
}

int launch_backward(int dtype, int flags, const Geom& g, int nsteps, int n_theta, int d, long nP,
                    int broadcast, const void* points, const void* As, const void* basis,
                    const void* grad_out, void* dtheta, void* dpoints, void* workspace,
                    size_t workspace_bytes, int* flagged, cudaStream_t st)
{
    if (n_theta == 0 || d == 0) return kOk;
    return backward_any(dtype, flags, g, nsteps, n_theta, d, nP, broadcast, points, As, basis, grad_out,
                        dtheta, dpoints, workspace, workspace_bytes, flagged, st, nullptr);
}

int launch_rk2_trace(const Geom& g, int nsteps, int n_theta, long nP, int broadcast, int mode,
                     const void* points, const void* As, void* workspace, size_t workspace_bytes,
                     int* cells, unsigned char* failed, cudaStream_t st)
{
    if (n_theta == 0 || nP == 0) return kOk;
    if (workspace_bytes < backward_workspace_bytes(kF32, g, n_theta, nP)) { set_error("rk2_trace: workspace too small"); return kErrWorkspace; }
    if (reinterpret_cast<uintptr_t>(workspace) & 15) { set_error("rk2_trace: workspace must be 16-byte aligned"); return kErrArgument; }
#define ARGS g, nsteps, n_theta, nP, broadcast, mode, points, As, workspace, cells, failed, st
    return CPAB_DISPATCH(rk2_trace_dim_1(ARGS), rk2_trace_dim_2(ARGS), rk2_trace_dim_3(ARGS));
#undef ARGS
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
    for (int j = 0; j < ndim; ++j)
        if (in_size[j] > kMaxInterpExtentF32) { set_error("transform_data: input extent %d exceeds %d", in_size[j], kMaxInterpExtentF32); return false; }
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

int launch_transform_data_backward(int dtype, int flags, const Geom& g, int nsteps, int n_theta, int d, int C,
                                   const int* in_size, const int* out_size, const void* points,
                                   const void* As, const void* basis, const void* data,
                                   const void* grid_t, const void* gimg, void* dtheta, void* workspace,
                                   size_t workspace_bytes, cudaStream_t st)
{
    SampleArgs sa;
    long nP = 0;
    if (!make_sample_shape(g.ndim, n_theta, C, in_size, out_size, sa.sh, nP)) return kErrUnsupported;
    if (n_theta == 0 || d == 0) return kOk;
    sa.data = data;
    sa.gimg = gimg;
    return backward_any(dtype, flags, g, nsteps, n_theta, d, nP, 0, points, As, basis, grid_t, dtheta,
                        nullptr, workspace, workspace_bytes, nullptr, st, &sa);
}

}  // namespace cpab
