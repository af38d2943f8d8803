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
#include "cpab_sample.cuh"

namespace cpab {

namespace {

thread_local int g_interp_max_ctas = 0;  // (per thread, like the other tuning knobs) tests: cap the persistent grid so that every CTA walks over many tiles ("interp_max_ctas")
thread_local int g_interp_variant = 9;   // 0-4: one tile per CTA (0: 4 points in flight, 1 CTA/SM target; 1: 2 / 6; 2: 1 / 8); 5-8: persistent kernels, grid tiles prefetched; 9 (default): the same, backward with its loads software-pipelined in registers (single channel); 10-11: forward software-pipelined as well; 12-15: measurement probes (cpab_b200_set_tuning "interp_variant")

// ---------------------------------------------------------------------------------------------
// forward.  NDIM >= 2: CTA = 256 threads, tile 32 (first index) x 32 (last index);
// blockIdx.x -> tile along the first index, blockIdx.y -> tile along the last index,
// blockIdx.z -> n * (middle extent) + middle index.  All intra-sample offsets are 32-bit (the
// host checks that one sample's grid, input and output each have < 2^31 elements); a thread owns
// 4 points and issues all their gathers for a channel before blending.
// ---------------------------------------------------------------------------------------------
template <typename T, int NDIM, bool FULL, int BATCH, bool ONECH>
__device__ __forceinline__ void interp_fwd_tile(const T* __restrict__ data, const T* __restrict__ grid,
                                                T* __restrict__ out, const Shape& s,
                                                T (&sg)[NDIM][TILE][TILE + 1])
{
    constexpr int NC = 1 << NDIM;
    const int mid = NDIM == 3 ? s.O[1] : 1;
    const int n = blockIdx.z / mid, im = blockIdx.z - n * mid;
    const int a0 = blockIdx.x * TILE, f0 = blockIdx.y * TILE;
    const int O0 = s.O[0], OF = s.O[NDIM - 1];
    const int nP = O0 * (NDIM >= 2 ? s.O[1] : 1) * (NDIM >= 3 ? s.O[2] : 1);
    const int pstride = NDIM == 2 ? O0 : O0 * s.O[1];            // grid-point stride of the last index
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;

    {   // phase A: unit stride along the first index
        const T* gp = grid + (size_t)n * NDIM * nP + ((a0 + lane) + (NDIM == 3 ? O0 * im : 0) + pstride * (f0 + wrp));
        T tmp[REPS][NDIM];
#pragma unroll
        for (int rep = 0; rep < REPS; ++rep) {
            const bool ok = FULL || ((a0 + lane < O0) && (f0 + wrp + 8 * rep < OF));
#pragma unroll
            for (int j = 0; j < NDIM; ++j) tmp[rep][j] = ok ? gp[j * nP + rep * 8 * pstride] : (T)0;
        }
#pragma unroll
        for (int rep = 0; rep < REPS; ++rep)
#pragma unroll
            for (int j = 0; j < NDIM; ++j) sg[j][wrp + 8 * rep][lane] = tmp[rep][j];
    }
    __syncthreads();

    // phase B: a warp runs along the last index; BATCH points per thread are in flight at a time.
    // Pointers advance by one channel plane per iteration; everything else is a 32-bit offset.
    const int iF = f0 + lane;
    const int plane = s.S[0] * (NDIM >= 2 ? s.S[1] : 1) * (NDIM >= 3 ? s.S[2] : 1);
    const T* dn = data + (size_t)n * s.C * plane;
    T* on = out + (size_t)n * s.C * nP + (NDIM == 2 ? (a0 + wrp) * s.O[1] + iF
                                                      : ((a0 + wrp) * s.O[1] + im) * s.O[2] + iF);
    const int ostride = 8 * (NDIM == 2 ? s.O[1] : s.O[1] * s.O[2]);     // output stride of one rep
    const int nch = ONECH ? 1 : s.C;       // single-channel images: no channel loop at all
#pragma unroll
    for (int r0 = 0; r0 < REPS; r0 += BATCH) {
        Taps<T, NDIM> tp[BATCH];
        bool ok[BATCH];
#pragma unroll
        for (int b = 0; b < BATCH; ++b) {
            const int a = wrp + 8 * (r0 + b);
            ok[b] = FULL || (a0 + a < O0 && iF < OF);
            T gc[NDIM];
#pragma unroll
            for (int j = 0; j < NDIM; ++j) gc[j] = sg[j][lane][a];
            tp[b] = make_taps<T, NDIM>(gc, s);
        }
        const T* dp = dn;
        T* op = on + r0 * ostride;
#pragma unroll 1
        for (int c = 0; c < nch; ++c, dp += plane, op += nP) {
            T v[BATCH][NC];
#pragma unroll
            for (int b = 0; b < BATCH; ++b)
                if (FULL || ok[b]) gather<T, NDIM>(dp, tp[b], v[b]);
#pragma unroll
            for (int b = 0; b < BATCH; ++b)
                if (FULL || ok[b]) op[b * ostride] = blend<NDIM>(v[b], tp[b].w);
        }
    }
}

// BATCH = points per thread whose gathers are in flight together; MINB = resident CTAs per SM
// the register allocation targets (more CTAs hide the two dependent memory round trips per tile)
template <typename T, int NDIM, int BATCH, int MINB, bool ONECH = false>
__global__ void __launch_bounds__(256, MINB)
k_interp_fwd(const T* __restrict__ data, const T* __restrict__ grid, T* __restrict__ out, Shape s)
{
    __shared__ T sg[NDIM][TILE][TILE + 1];
    const bool full = (blockIdx.x * TILE + TILE <= s.O[0]) && (blockIdx.y * TILE + TILE <= s.O[NDIM - 1]);
    if (full) interp_fwd_tile<T, NDIM, true, BATCH, ONECH>(data, grid, out, s, sg);
    else interp_fwd_tile<T, NDIM, false, BATCH, ONECH>(data, grid, out, s, sg);
}

// 1-D: no transposition needed.  A thread owns PT points (strided by the CTA width so that every
// access stays unit-stride) and issues all their loads before blending.
constexpr int PT1D = 4;

template <typename T>
__global__ void __launch_bounds__(256)
k_interp_fwd_1d(const T* __restrict__ data, const T* __restrict__ grid, T* __restrict__ out, Shape s)
{
    const int n = blockIdx.y;
    const int i0 = blockIdx.x * (256 * PT1D) + threadIdx.x;
    const T* gn = grid + (size_t)n * s.O[0];
    T gc[PT1D];
#pragma unroll
    for (int u = 0; u < PT1D; ++u) gc[u] = (i0 + 256 * u < s.O[0]) ? gn[i0 + 256 * u] : (T)0;
    int t0[PT1D], t1[PT1D];
    T w[PT1D];
#pragma unroll
    for (int u = 0; u < PT1D; ++u) taps(gc[u], s.S[0], t0[u], t1[u], w[u]);
    for (int c = 0; c < s.C; ++c) {
        const T* dp = data + ((size_t)n * s.C + c) * s.S[0];
        T v[PT1D][2];
#pragma unroll
        for (int u = 0; u < PT1D; ++u) { v[u][0] = __ldg(dp + t0[u]); v[u][1] = __ldg(dp + t1[u]); }
        T* op = out + ((size_t)n * s.C + c) * s.O[0];
#pragma unroll
        for (int u = 0; u < PT1D; ++u)
            if (i0 + 256 * u < s.O[0]) op[i0 + 256 * u] = blend<1>(v[u], &w[u]);
    }
}

// ---------------------------------------------------------------------------------------------
// backward: dgrid [N,NDIM,nP] (optional) and ddata [N,C,S...] (optional, accumulated atomically
// into a zero-initialised buffer).  Same tiling as the forward; the d/dgrid tile returns through
// shared memory so that it is stored unit-stride along the first index.
// ---------------------------------------------------------------------------------------------
template <typename T, int NDIM, bool FULL, bool DDATA, int BATCH>
__device__ __forceinline__ void interp_bwd_tile(const T* __restrict__ data, const T* __restrict__ grid,
                                                const T* __restrict__ gout, T* __restrict__ dgrid,
                                                T* __restrict__ ddata, const Shape& s,
                                                T (&sg)[NDIM][TILE][TILE + 1])
{
    constexpr int NC = 1 << NDIM;
    constexpr int H = 1 << (NDIM - 1);
    const int mid = NDIM == 3 ? s.O[1] : 1;
    const int n = blockIdx.z / mid, im = blockIdx.z - n * mid;
    const int a0 = blockIdx.x * TILE, f0 = blockIdx.y * TILE;
    const int O0 = s.O[0], OF = s.O[NDIM - 1];
    const int nP = O0 * (NDIM >= 2 ? s.O[1] : 1) * (NDIM >= 3 ? s.O[2] : 1);
    const int pstride = NDIM == 2 ? O0 : O0 * s.O[1];
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const size_t goff = (size_t)n * NDIM * nP + ((a0 + lane) + (NDIM == 3 ? O0 * im : 0) + pstride * (f0 + wrp));

    {
        const T* gp = grid + goff;
        T tmp[REPS][NDIM];
#pragma unroll
        for (int rep = 0; rep < REPS; ++rep) {
            const bool ok = FULL || ((a0 + lane < O0) && (f0 + wrp + 8 * rep < OF));
#pragma unroll
            for (int j = 0; j < NDIM; ++j) tmp[rep][j] = ok ? gp[j * nP + rep * 8 * pstride] : (T)0;
        }
#pragma unroll
        for (int rep = 0; rep < REPS; ++rep)
#pragma unroll
            for (int j = 0; j < NDIM; ++j) sg[j][wrp + 8 * rep][lane] = tmp[rep][j];
    }
    __syncthreads();

    const int iF = f0 + lane;
    const int plane = s.S[0] * (NDIM >= 2 ? s.S[1] : 1) * (NDIM >= 3 ? s.S[2] : 1);
    const T* dn = data + (size_t)n * s.C * plane;
    T* ddn = DDATA ? ddata + (size_t)n * s.C * plane : nullptr;
    const T* gon = gout + (size_t)n * s.C * nP + (NDIM == 2 ? (a0 + wrp) * s.O[1] + iF
                                                             : ((a0 + wrp) * s.O[1] + im) * s.O[2] + iF);
    const int ostride = 8 * (NDIM == 2 ? s.O[1] : s.O[1] * s.O[2]);
    const int nch = s.C;
    T dg[REPS][NDIM];
#pragma unroll
    for (int r0 = 0; r0 < REPS; r0 += BATCH) {
        Taps<T, NDIM> tp[BATCH];
        bool ok[BATCH];
#pragma unroll
        for (int b = 0; b < BATCH; ++b) {
            const int a = wrp + 8 * (r0 + b);
            ok[b] = FULL || (a0 + a < O0 && iF < OF);
            T gc[NDIM];
#pragma unroll
            for (int j = 0; j < NDIM; ++j) { gc[j] = sg[j][lane][a]; dg[r0 + b][j] = 0; }
            tp[b] = make_taps<T, NDIM>(gc, s);
        }
        const T* dp = dn;
        const T* gp = gon + r0 * ostride;
        T* qd = ddn;
#pragma unroll 1
        for (int c = 0; c < nch; ++c, dp += plane, gp += nP) {
            T v[BATCH][NC], g[BATCH];
#pragma unroll
            for (int b = 0; b < BATCH; ++b) {
                if (FULL || ok[b]) {
                    gather<T, NDIM>(dp, tp[b], v[b]);
                    g[b] = gp[b * ostride];
                }
            }
#pragma unroll
            for (int b = 0; b < BATCH; ++b) {
                if (FULL || ok[b]) {
                    T gv[NC], dw[NDIM];
                    blend_vjp<NDIM>(v[b], tp[b].w, g[b], gv, dw);
#pragma unroll
                    for (int j = 0; j < NDIM; ++j) dg[r0 + b][j] += dw[j];
                    if (DDATA) {
#pragma unroll
                        for (int u = 0; u < H; ++u) {
                            atomicAdd(qd + tp[b].base[u], gv[u]);
                            atomicAdd(qd + tp[b].base[u] + (tp[b].two ? 1 : 0), gv[u + H]);
                        }
                    }
                }
            }
            if (DDATA) qd += plane;
        }
    }
    if (dgrid == nullptr) return;
    // xd = x - x0 with x = g*(size-1): d/dg = size-1
#pragma unroll
    for (int rep = 0; rep < REPS; ++rep)
#pragma unroll
        for (int j = 0; j < NDIM; ++j) sg[j][lane][wrp + 8 * rep] = dg[rep][j] * (T)(s.S[j] - 1);
    __syncthreads();
    T* dp_out = dgrid + goff;
#pragma unroll
    for (int rep = 0; rep < REPS; ++rep) {
        const bool okA = FULL || ((a0 + lane < O0) && (f0 + wrp + 8 * rep < OF));
        if (okA) {
#pragma unroll
            for (int j = 0; j < NDIM; ++j) dp_out[j * nP + rep * 8 * pstride] = sg[j][wrp + 8 * rep][lane];
        }
    }
}

template <typename T, int NDIM, int BATCH, int MINB>
__global__ void __launch_bounds__(256, MINB)
k_interp_bwd(const T* __restrict__ data, const T* __restrict__ grid, const T* __restrict__ gout,
             T* __restrict__ dgrid, T* __restrict__ ddata, Shape s)
{
    __shared__ T sg[NDIM][TILE][TILE + 1];
    const bool full = (blockIdx.x * TILE + TILE <= s.O[0]) && (blockIdx.y * TILE + TILE <= s.O[NDIM - 1]);
    if (ddata != nullptr) {
        if (full) interp_bwd_tile<T, NDIM, true, true, BATCH>(data, grid, gout, dgrid, ddata, s, sg);
        else interp_bwd_tile<T, NDIM, false, true, BATCH>(data, grid, gout, dgrid, ddata, s, sg);
    } else {
        if (full) interp_bwd_tile<T, NDIM, true, false, BATCH>(data, grid, gout, dgrid, ddata, s, sg);
        else interp_bwd_tile<T, NDIM, false, false, BATCH>(data, grid, gout, dgrid, ddata, s, sg);
    }
}

// ---------------------------------------------------------------------------------------------
// Persistent, software-pipelined variants (float32, NDIM >= 2, first output extent a multiple of 4)
// -- the default.
//
// The one-tile-per-CTA kernels above pay two dependent HBM round trips per tile (grid tile, then
// texels), one CTA launch per 1024 points and ~24 instructions per point for the staging alone.
// Here a CTA walks over tiles (grid = resident CTAs only) and keeps the grid tiles of the next two
// iterations in flight with 16-byte cp.async (LDGSTS: global -> shared without passing through
// registers) in a 3-stage ring: a tile's coordinates are already in shared memory when its turn
// comes and the only exposed latency per tile is the texel gather, which all four points of a
// thread issue together.  One __syncthreads per tile: it publishes the landed stage and, at the
// same time, retires the stage that is about to be refilled (read during the previous iteration).
//
// Stage layout: [NDIM][TILE rows (last output index)][PITCH = 36 floats].  Copy mapping: thread t
// moves the 16-byte chunk (row t/8, chunk t%8) of every coordinate plane -- 8 threads cover 128
// contiguous bytes of global memory.  Compute mapping: lane = row (runs along the last output index:
// texel gathers and stores are unit-stride for near-identity warps), warp w owns the FOUR first-index
// points 4w..4w+3 of its row and reads their coordinates with ONE LDS.128 per plane; with the
// 144-byte pitch the eight lanes of a quarter-warp hit disjoint bank groups (conflict-free).
// ---------------------------------------------------------------------------------------------
constexpr int kStages = 3;
constexpr int PITCH = TILE + 4;

__device__ __forceinline__ void cp_async16(uint32_t dst, const float* src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// shared memory through 32-bit window addresses (a generic pointer into the ring makes ptxas rebuild
// the window base -- S2UR SR_CgaCtaId, ULEA -- in every iteration)
__device__ __forceinline__ void lds_f32x4(uint32_t a, float* v)
{
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(a) : "memory");
}
__device__ __forceinline__ void sts_f32x4(uint32_t a, const float* v)
{
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
}

// Tiles are numbered first-index-fastest: t = ((n * mid + im) * tiles_f + tf) * tiles_a + ta.  A CTA
// visits t = blockIdx.x + k * gridDim.x; the position is kept as mixed-radix digits and advanced by
// the digits of gridDim.x (host-computed) with carries -- no division in the tile loop.
struct TileGeom {
    int tiles_a, tiles_f, mid;      // tiles along the first / last output index, middle extent (3-D)
    int da, df, dm, dn;             // gridDim.x in the same radix
};

struct TilePos { int n, im, a0, f0; };

struct TileCursor {
    int ta, tf, im, n;
    __device__ __forceinline__ void init(unsigned t, const TileGeom& tg)
    {
        unsigned q = t / (unsigned)tg.tiles_a;
        ta = (int)(t - q * tg.tiles_a);
        unsigned q2 = q / (unsigned)tg.tiles_f;
        tf = (int)(q - q2 * tg.tiles_f);
        n = (int)(q2 / (unsigned)tg.mid);
        im = (int)(q2 - (unsigned)n * tg.mid);
    }
    __device__ __forceinline__ void advance(const TileGeom& tg)
    {
        ta += tg.da;
        tf += tg.df;
        if (ta >= tg.tiles_a) { ta -= tg.tiles_a; ++tf; }
        im += tg.dm;
        if (tf >= tg.tiles_f) { tf -= tg.tiles_f; ++im; }
        n += tg.dn;
        if (im >= tg.mid) { im -= tg.mid; ++n; }
    }
    __device__ __forceinline__ TilePos pos() const { return TilePos{n, im, ta * TILE, tf * TILE}; }
};

// element offset of this thread's 16-byte chunk of the tile at `tp` within ONE SAMPLE's first
// coordinate plane (32-bit: the host checks that a sample's grid has < 2^31 elements).  Rows /
// chunks beyond the image are clamped to its last row / chunk: those entries hold valid but unused
// values, and no copy needs a predicate.
template <int NDIM>
__device__ __forceinline__ unsigned chunk_offset32(const Shape& s, const TilePos& tp)
{
    const int row = threadIdx.x >> 3, ch = threadIdx.x & 7;
    const int O0 = s.O[0];
    const int pstride = NDIM == 2 ? O0 : O0 * s.O[1];
    const int ix = min(tp.a0 + 4 * ch, O0 - 4);
    const int f = min(tp.f0 + row, s.O[NDIM - 1] - 1);
    return (unsigned)(pstride * f + (NDIM == 3 ? O0 * tp.im : 0) + ix);
}
// start of sample n's grid (uniform)
template <int NDIM>
__device__ __forceinline__ size_t sample_offset(const Shape& s, int n)
{
    const unsigned nP = (unsigned)(s.O[0] * s.O[1] * (NDIM >= 3 ? s.O[2] : 1));
    return (size_t)n * NDIM * nP;
}
template <int NDIM>
__device__ __forceinline__ size_t chunk_offset(const Shape& s, const TilePos& tp)
{
    return sample_offset<NDIM>(s, tp.n) + chunk_offset32<NDIM>(s, tp);
}

template <int NDIM>
__device__ __forceinline__ void prefetch_grid_tile(const float* __restrict__ grid, const Shape& s,
                                                   const TileCursor& cur, uint32_t sg)
{
    if (cur.n < s.N) {
        const unsigned nP = (unsigned)(s.O[0] * s.O[1] * (NDIM >= 3 ? s.O[2] : 1));
        const float* gn = grid + sample_offset<NDIM>(s, cur.n);               // uniform
        const float* src = gn + chunk_offset32<NDIM>(s, cur.pos());           // + one 32-bit thread offset
        const uint32_t dst = sg + 4u * (uint32_t)((threadIdx.x >> 3) * PITCH + 4 * (threadIdx.x & 7));
#pragma unroll
        for (int j = 0; j < NDIM; ++j) cp_async16(dst + 4u * (uint32_t)(j * TILE * PITCH), src + (size_t)j * nP);
    }
    cp_async_commit();      // (an empty group when there is no tile: keeps the group count uniform)
}

// PROBE != 0: measurement-only variants that keep the traffic and drop parts of the work (tools/interp_variants.py):
// 1 = no texel access at all (out = gx + gy), 2 = one aligned texel load at the identity position,
// 3 = four texel loads around the identity position (no tap arithmetic).  Never used by the product.
template <int NDIM, bool FULL, bool ONECH, int PROBE = 0>
__device__ __forceinline__ void interp_fwd_compute(const float* __restrict__ data, float* __restrict__ out, const Shape& s,
                                                   const TilePos& tp, uint32_t sg)
{
    constexpr int NC = 1 << NDIM;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int O0 = s.O[0], OF = s.O[NDIM - 1];
    const int nP = O0 * s.O[1] * (NDIM >= 3 ? s.O[2] : 1);
    const int iF = tp.f0 + lane, iA = tp.a0 + 4 * wrp;
    const int plane = s.S[0] * s.S[1] * (NDIM >= 3 ? s.S[2] : 1);
    const float* dn = data + (size_t)tp.n * s.C * plane;
    asm volatile("" : "+l"(dn));       // one base pointer per image: every gather address is then one IMAD.WIDE
    const int ostride = NDIM == 2 ? s.O[1] : s.O[1] * s.O[2];            // output stride of the first index
    float* on = out + (size_t)tp.n * s.C * nP + (NDIM == 2 ? iA * s.O[1] + iF : (iA * s.O[1] + tp.im) * s.O[2] + iF);
    const int nch = ONECH ? 1 : s.C;
    float gc[NDIM][4];
#pragma unroll
    for (int j = 0; j < NDIM; ++j) lds_f32x4(sg + 4u * (uint32_t)((j * TILE + lane) * PITCH + 4 * wrp), gc[j]);
    if constexpr (PROBE != 0 && NDIM == 2) {
        float r[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            r[b] = gc[0][b] + gc[1][b];
            const int ia = min(iA + b, s.S[0] - 2), jf = min(iF, s.S[1] - 2);
            const float* q = dn + ia * s.S[1] + jf;
            if (PROBE == 2) r[b] += __ldg(q);
            if (PROBE == 3) r[b] += (__ldg(q) + __ldg(q + 1)) + (__ldg(q + s.S[1]) + __ldg(q + s.S[1] + 1));
        }
#pragma unroll
        for (int b = 0; b < 4; ++b)
            if (FULL || (iA + b < O0 && iF < OF)) on[b * ostride] = r[b];
    } else {
    Taps<float, NDIM> tq[4];
    bool ok[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        ok[b] = FULL || (iA + b < O0 && iF < OF);
        float g1[NDIM];
#pragma unroll
        for (int j = 0; j < NDIM; ++j) g1[j] = gc[j][b];
        tq[b] = make_taps<float, NDIM>(g1, s);
    }
    const float* dp = dn;
    float* op = on;
#pragma unroll 1
    for (int c = 0; c < nch; ++c, dp += plane, op += nP) {
        float v[4][NC];
#pragma unroll
        for (int b = 0; b < 4; ++b)
            if (FULL || ok[b]) gather<float, NDIM>(dp, tq[b], v[b]);
#pragma unroll
        for (int b = 0; b < 4; ++b)
            if (FULL || ok[b]) op[b * ostride] = blend<NDIM>(v[b], tq[b].w);
    }
    }
}

template <int NDIM, int MINB, bool ONECH, int PROBE = 0>
__global__ void __launch_bounds__(256, MINB)
k_interp_fwd_pipe(const float* __restrict__ data, const float* __restrict__ grid, float* __restrict__ out,
                  const __grid_constant__ Shape s, const __grid_constant__ TileGeom tg)
{
    extern __shared__ __align__(16) float sg_ring[];       // [kStages][NDIM][TILE][PITCH]
    constexpr int STAGE = NDIM * TILE * PITCH;
    const uint32_t ring = (uint32_t)__cvta_generic_to_shared(sg_ring);
    TileCursor cur, pf;                 // tile being computed / tile being prefetched (two ahead)
    cur.init(blockIdx.x, tg);
    pf = cur;
    prefetch_grid_tile<NDIM>(grid, s, pf, ring);
    pf.advance(tg);
    prefetch_grid_tile<NDIM>(grid, s, pf, ring + 4u * STAGE);
    int stage = 0;
    for (; cur.n < s.N; cur.advance(tg)) {
        cp_async_wait<1>();
        if (PROBE == 4) __syncwarp(); else __syncthreads();       // (probe 4: timing only, results are wrong)
        const int nxt = stage >= 1 ? stage - 1 : kStages - 1;      // (stage + 2) % 3
        pf.advance(tg);
        prefetch_grid_tile<NDIM>(grid, s, pf, ring + 4u * (uint32_t)(nxt * STAGE));
        const TilePos tp = cur.pos();
        const bool full = (tp.a0 + TILE <= s.O[0]) && (tp.f0 + TILE <= s.O[NDIM - 1]);
        const uint32_t sg = ring + 4u * (uint32_t)(stage * STAGE);
        if (full) interp_fwd_compute<NDIM, true, ONECH, PROBE == 4 ? 0 : PROBE>(data, out, s, tp, sg);
        else interp_fwd_compute<NDIM, false, ONECH, PROBE == 4 ? 0 : PROBE>(data, out, s, tp, sg);
        stage = stage + 1 == kStages ? 0 : stage + 1;
    }
    cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------
// Forward, single channel, software-pipelined in REGISTERS (variants 10-11; NOT the default): the
// texel loads of tile i+1 are issued before tile i is blended, so a thread has up to 2 x 16 gathers
// in flight and a tile's gather latency overlaps the previous tile's arithmetic and stores.  Two
// register sets alternate (the loop is unrolled by two, no moves).  Loads are unconditional
// (clamped grid coordinates give valid addresses everywhere), only the stores of edge tiles are
// predicated.  Measured on 128 x 512^2 (profiles/r02_interp_variants.txt): 3.74 TB/s against 4.03 for
// k_interp_fwd_pipe -- the forward is limited by instruction issue (the skeleton alone, no texel
// access, streams at 5.6 TB/s; one aligned texel load per point: 5.2; four: 4.3-4.9), not by exposed
// latency, and this version executes ~12 % more instructions.  The same pipeline DOES pay in the
// backward (k_interp_bwd_sw: 4.14 against 3.69 TB/s).  Gathers through cp.async into a second shared
// ring -- more loads in flight at no register cost -- were slower still (3.3 TB/s) and were removed.
// ---------------------------------------------------------------------------------------------
template <int NDIM, int MINB>
__global__ void __launch_bounds__(256, MINB)
k_interp_fwd_sw(const float* __restrict__ data, const float* __restrict__ grid, float* __restrict__ out,
                const __grid_constant__ Shape s, const __grid_constant__ TileGeom tg)
{
    constexpr int NC = 1 << NDIM;
    constexpr int GSTAGE = NDIM * TILE * PITCH;
    extern __shared__ __align__(16) float sg_ring[];
    const uint32_t ring = (uint32_t)__cvta_generic_to_shared(sg_ring);
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int O0 = s.O[0], OF = s.O[NDIM - 1];
    const int nP = O0 * s.O[1] * (NDIM >= 3 ? s.O[2] : 1);
    const int plane = s.S[0] * s.S[1] * (NDIM >= 3 ? s.S[2] : 1);
    const int ostride = NDIM == 2 ? s.O[1] : s.O[1] * s.O[2];
    const uint32_t mine = ring + 4u * (uint32_t)(lane * PITCH + 4 * wrp);

    TileCursor cur, nx, pf;
    cur.init(blockIdx.x, tg);
    pf = cur;
    int gstage = 0;

    auto load_tile = [&](const TileCursor& c, int gs, float (&v)[4][NC], float (&w)[4][NDIM]) {
        if (c.n >= s.N) return;
        const float* dn = data + (size_t)c.n * plane;
        asm volatile("" : "+l"(dn));
        float gc[NDIM][4];
#pragma unroll
        for (int j = 0; j < NDIM; ++j) lds_f32x4(mine + 4u * (uint32_t)(gs * GSTAGE + j * TILE * PITCH), gc[j]);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            float g1[NDIM];
#pragma unroll
            for (int j = 0; j < NDIM; ++j) g1[j] = gc[j][b];
            const Taps<float, NDIM> tq = make_taps<float, NDIM>(g1, s);
#pragma unroll
            for (int j = 0; j < NDIM; ++j) w[b][j] = tq.w[j];
            gather<float, NDIM>(dn, tq, v[b]);
        }
    };
    auto blend_tile = [&](const TileCursor& c, const float (&v)[4][NC], const float (&w)[4][NDIM]) {
        const TilePos tp = c.pos();
        const int iF = tp.f0 + lane, iA = tp.a0 + 4 * wrp;
        float* on = out + (size_t)tp.n * nP + (NDIM == 2 ? iA * s.O[1] + iF : (iA * s.O[1] + tp.im) * s.O[2] + iF);
        const bool full = (tp.a0 + TILE <= O0) && (tp.f0 + TILE <= OF);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const float r = blend<NDIM>(v[b], w[b]);
            if (full || (iA + b < O0 && iF < OF)) on[b * ostride] = r;
        }
    };
    // one pipeline step: the grid of tile `nx` has landed -> issue its texel loads into (vn, wn);
    // refill the grid ring; blend tile `cur` from (vc, wc)
    auto step = [&](float (&vc)[4][NC], float (&wc)[4][NDIM], float (&vn)[4][NC], float (&wn)[4][NDIM]) {
        cp_async_wait<1>();
        __syncthreads();
        const int g1s = gstage + 1 == kStages ? 0 : gstage + 1;
        load_tile(nx, g1s, vn, wn);
        pf.advance(tg);
        prefetch_grid_tile<NDIM>(grid, s, pf, ring + 4u * (uint32_t)(gstage * GSTAGE));
        blend_tile(cur, vc, wc);
        gstage = g1s;
        cur = nx;
        nx.advance(tg);
    };

    float vA[4][NC], wA[4][NDIM], vB[4][NC], wB[4][NDIM];
    prefetch_grid_tile<NDIM>(grid, s, pf, ring);
    pf.advance(tg);
    prefetch_grid_tile<NDIM>(grid, s, pf, ring + 4u * GSTAGE);
    pf.advance(tg);
    cp_async_wait<1>();
    __syncthreads();
    load_tile(cur, 0, vA, wA);
    prefetch_grid_tile<NDIM>(grid, s, pf, ring + 4u * 2 * GSTAGE);
    nx = cur;
    nx.advance(tg);
    while (cur.n < s.N) {
        step(vA, wA, vB, wB);
        if (cur.n >= s.N) break;
        step(vB, wB, vA, wA);
    }
    cp_async_wait<0>();
}

// backward: d/dgrid goes back through the tile's own stage (every thread overwrites exactly the
// entries it read) and leaves in the copy mapping, 16 bytes per thread and plane.
template <int NDIM, bool FULL, bool DDATA, int BATCH>
__device__ __forceinline__ void interp_bwd_compute(const float* __restrict__ data, const float* __restrict__ gout,
                                                   float* __restrict__ dgrid, float* __restrict__ ddata,
                                                   const Shape& s, const TilePos& tp, uint32_t sg)
{
    constexpr int NC = 1 << NDIM;
    constexpr int H = 1 << (NDIM - 1);
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int O0 = s.O[0], OF = s.O[NDIM - 1];
    const int nP = O0 * s.O[1] * (NDIM >= 3 ? s.O[2] : 1);
    const int iF = tp.f0 + lane, iA = tp.a0 + 4 * wrp;
    const int plane = s.S[0] * s.S[1] * (NDIM >= 3 ? s.S[2] : 1);
    const float* dn = data + (size_t)tp.n * s.C * plane;
    asm volatile("" : "+l"(dn));
    float* ddn = DDATA ? ddata + (size_t)tp.n * s.C * plane : nullptr;
    const int ostride = NDIM == 2 ? s.O[1] : s.O[1] * s.O[2];
    const float* gon = gout + (size_t)tp.n * s.C * nP + (NDIM == 2 ? iA * s.O[1] + iF : (iA * s.O[1] + tp.im) * s.O[2] + iF);
    const int nch = s.C;
    const uint32_t mine = sg + 4u * (uint32_t)(lane * PITCH + 4 * wrp);
    float gc[NDIM][4], dg[NDIM][4];
#pragma unroll
    for (int j = 0; j < NDIM; ++j) lds_f32x4(mine + 4u * (uint32_t)(j * TILE * PITCH), gc[j]);
#pragma unroll
    for (int r0 = 0; r0 < 4; r0 += BATCH) {
        Taps<float, NDIM> tq[BATCH];
        bool ok[BATCH];
#pragma unroll
        for (int b = 0; b < BATCH; ++b) {
            ok[b] = FULL || (iA + r0 + b < O0 && iF < OF);
            float g1[NDIM];
#pragma unroll
            for (int j = 0; j < NDIM; ++j) { g1[j] = gc[j][r0 + b]; dg[j][r0 + b] = 0; }
            tq[b] = make_taps<float, NDIM>(g1, s);
        }
        const float* dp = dn;
        const float* gp = gon + r0 * ostride;
        float* qd = ddn;
#pragma unroll 1
        for (int c = 0; c < nch; ++c, dp += plane, gp += nP) {
            float v[BATCH][NC], g[BATCH];
#pragma unroll
            for (int b = 0; b < BATCH; ++b) {
                if (FULL || ok[b]) {
                    gather<float, NDIM>(dp, tq[b], v[b]);
                    g[b] = __ldg(gp + b * ostride);
                }
            }
#pragma unroll
            for (int b = 0; b < BATCH; ++b) {
                if (FULL || ok[b]) {
                    float gv[NC], dw[NDIM];
                    blend_vjp<NDIM>(v[b], tq[b].w, g[b], gv, dw);
#pragma unroll
                    for (int j = 0; j < NDIM; ++j) dg[j][r0 + b] += dw[j];
                    if (DDATA) {
#pragma unroll
                        for (int u = 0; u < H; ++u) {
                            atomicAdd(qd + tq[b].base[u], gv[u]);
                            atomicAdd(qd + tq[b].base[u] + (tq[b].two ? 1 : 0), gv[u + H]);
                        }
                    }
                }
            }
            if (DDATA) qd += plane;
        }
    }
    if (dgrid == nullptr) return;
    // xd = x - x0 with x = g*(size-1): d/dg = size-1
#pragma unroll
    for (int j = 0; j < NDIM; ++j) {
        const float sc = (float)(s.S[j] - 1);
#pragma unroll
        for (int b = 0; b < 4; ++b) dg[j][b] *= sc;
        sts_f32x4(mine + 4u * (uint32_t)(j * TILE * PITCH), dg[j]);
    }
    __syncthreads();
    const int row = threadIdx.x >> 3, ch = threadIdx.x & 7;
    if (FULL || (tp.a0 + 4 * ch < O0 && tp.f0 + row < OF)) {
        float* dst = dgrid + chunk_offset<NDIM>(s, tp);
#pragma unroll
        for (int j = 0; j < NDIM; ++j) {
            float v[4];
            lds_f32x4(sg + 4u * (uint32_t)((j * TILE + row) * PITCH + 4 * ch), v);
            *reinterpret_cast<float4*>(dst + (size_t)j * nP) = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
}

template <int NDIM, int BATCH, int MINB>
__global__ void __launch_bounds__(256, MINB)
k_interp_bwd_pipe(const float* __restrict__ data, const float* __restrict__ grid, const float* __restrict__ gout,
                  float* __restrict__ dgrid, float* __restrict__ ddata, const __grid_constant__ Shape s,
                  const __grid_constant__ TileGeom tg)
{
    extern __shared__ __align__(16) float sg_ring[];
    constexpr int STAGE = NDIM * TILE * PITCH;
    const uint32_t ring = (uint32_t)__cvta_generic_to_shared(sg_ring);
    TileCursor cur, pf;
    cur.init(blockIdx.x, tg);
    pf = cur;
    prefetch_grid_tile<NDIM>(grid, s, pf, ring);
    pf.advance(tg);
    prefetch_grid_tile<NDIM>(grid, s, pf, ring + 4u * STAGE);
    int stage = 0;
    for (; cur.n < s.N; cur.advance(tg)) {
        cp_async_wait<1>();
        __syncthreads();
        const int nxt = stage >= 1 ? stage - 1 : kStages - 1;
        pf.advance(tg);
        prefetch_grid_tile<NDIM>(grid, s, pf, ring + 4u * (uint32_t)(nxt * STAGE));
        const TilePos tp = cur.pos();
        const bool full = (tp.a0 + TILE <= s.O[0]) && (tp.f0 + TILE <= s.O[NDIM - 1]);
        const uint32_t sg = ring + 4u * (uint32_t)(stage * STAGE);
        if (ddata != nullptr) {
            if (full) interp_bwd_compute<NDIM, true, true, BATCH>(data, gout, dgrid, ddata, s, tp, sg);
            else interp_bwd_compute<NDIM, false, true, BATCH>(data, gout, dgrid, ddata, s, tp, sg);
        } else {
            if (full) interp_bwd_compute<NDIM, true, false, BATCH>(data, gout, dgrid, ddata, s, tp, sg);
            else interp_bwd_compute<NDIM, false, false, BATCH>(data, gout, dgrid, ddata, s, tp, sg);
        }
        stage = stage + 1 == kStages ? 0 : stage + 1;
    }
    cp_async_wait<0>();
}

// Backward w.r.t. the grid, single channel, no d/d(data): the same register software pipeline.  The
// d/dgrid tile leaves through a stage of its own (the grid ring is being refilled meanwhile).
template <int NDIM, int MINB>
__global__ void __launch_bounds__(256, MINB)
k_interp_bwd_sw(const float* __restrict__ data, const float* __restrict__ grid, const float* __restrict__ gout,
                float* __restrict__ dgrid, const __grid_constant__ Shape s, const __grid_constant__ TileGeom tg)
{
    constexpr int NC = 1 << NDIM;
    constexpr int GSTAGE = NDIM * TILE * PITCH;
    extern __shared__ __align__(16) float sg_ring[];       // [kStages][GSTAGE] grid ring | [GSTAGE] d/dgrid staging
    const uint32_t ring = (uint32_t)__cvta_generic_to_shared(sg_ring);
    const uint32_t ostage = ring + 4u * (uint32_t)(kStages * GSTAGE);
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int O0 = s.O[0], OF = s.O[NDIM - 1];
    const int nP = O0 * s.O[1] * (NDIM >= 3 ? s.O[2] : 1);
    const int plane = s.S[0] * s.S[1] * (NDIM >= 3 ? s.S[2] : 1);
    const int ostride = NDIM == 2 ? s.O[1] : s.O[1] * s.O[2];
    const uint32_t mine = 4u * (uint32_t)(lane * PITCH + 4 * wrp);
    float scale[NDIM];                  // xd = x - x0 with x = g*(size-1): d/dg = size-1
#pragma unroll
    for (int j = 0; j < NDIM; ++j) scale[j] = (float)(s.S[j] - 1);

    TileCursor cur, nx, pf;
    cur.init(blockIdx.x, tg);
    pf = cur;
    int gstage = 0;

    auto load_tile = [&](const TileCursor& c, int gs, float (&v)[4][NC], float (&w)[4][NDIM], float (&g)[4]) {
        if (c.n >= s.N) return;
        const TilePos tp = c.pos();
        const float* dn = data + (size_t)c.n * plane;
        asm volatile("" : "+l"(dn));
        // upstream gradient: clamped like the grid chunks (edge tiles read valid, unused values)
        const int iF = min(tp.f0 + lane, OF - 1), iA = tp.a0 + 4 * wrp;
        const float* gp = gout + (size_t)tp.n * nP + (NDIM == 2 ? iF : tp.im * s.O[2] + iF);
        float gc[NDIM][4];
#pragma unroll
        for (int j = 0; j < NDIM; ++j) lds_f32x4(ring + mine + 4u * (uint32_t)(gs * GSTAGE + j * TILE * PITCH), gc[j]);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            float g1[NDIM];
#pragma unroll
            for (int j = 0; j < NDIM; ++j) g1[j] = gc[j][b];
            const Taps<float, NDIM> tq = make_taps<float, NDIM>(g1, s);
#pragma unroll
            for (int j = 0; j < NDIM; ++j) w[b][j] = tq.w[j];
            gather<float, NDIM>(dn, tq, v[b]);
            g[b] = __ldg(gp + (size_t)min(iA + b, O0 - 1) * ostride);
        }
    };
    auto finish_tile = [&](const TileCursor& c, const float (&v)[4][NC], const float (&w)[4][NDIM], const float (&g)[4]) {
        const TilePos tp = c.pos();
        float dg[NDIM][4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            float gv[NC], dw[NDIM];
            blend_vjp<NDIM>(v[b], w[b], g[b], gv, dw);
#pragma unroll
            for (int j = 0; j < NDIM; ++j) dg[j][b] = dw[j] * scale[j];
        }
#pragma unroll
        for (int j = 0; j < NDIM; ++j) sts_f32x4(ostage + mine + 4u * (uint32_t)(j * TILE * PITCH), dg[j]);
        __syncthreads();
        const int row = threadIdx.x >> 3, ch = threadIdx.x & 7;
        if (tp.a0 + 4 * ch < O0 && tp.f0 + row < OF) {
            float* dst = dgrid + chunk_offset<NDIM>(s, tp);
#pragma unroll
            for (int j = 0; j < NDIM; ++j) {
                float t[4];
                lds_f32x4(ostage + 4u * (uint32_t)((j * TILE + row) * PITCH + 4 * ch), t);
                *reinterpret_cast<float4*>(dst + (size_t)j * nP) = make_float4(t[0], t[1], t[2], t[3]);
            }
        }
    };
    auto step = [&](float (&vc)[4][NC], float (&wc)[4][NDIM], float (&gc_)[4],
                    float (&vn)[4][NC], float (&wn)[4][NDIM], float (&gn)[4]) {
        cp_async_wait<1>();
        __syncthreads();
        const int g1s = gstage + 1 == kStages ? 0 : gstage + 1;
        load_tile(nx, g1s, vn, wn, gn);
        pf.advance(tg);
        prefetch_grid_tile<NDIM>(grid, s, pf, ring + 4u * (uint32_t)(gstage * GSTAGE));
        finish_tile(cur, vc, wc, gc_);
        gstage = g1s;
        cur = nx;
        nx.advance(tg);
    };

    float vA[4][NC], wA[4][NDIM], gA[4], vB[4][NC], wB[4][NDIM], gB[4];
    prefetch_grid_tile<NDIM>(grid, s, pf, ring);
    pf.advance(tg);
    prefetch_grid_tile<NDIM>(grid, s, pf, ring + 4u * GSTAGE);
    pf.advance(tg);
    cp_async_wait<1>();
    __syncthreads();
    load_tile(cur, 0, vA, wA, gA);
    prefetch_grid_tile<NDIM>(grid, s, pf, ring + 4u * 2 * GSTAGE);
    nx = cur;
    nx.advance(tg);
    while (cur.n < s.N) {
        step(vA, wA, gA, vB, wB, gB);
        if (cur.n >= s.N) break;
        step(vB, wB, gB, vA, wA, gA);
    }
    cp_async_wait<0>();
}

template <typename KERN, typename... Args>
static int launch_pipe(KERN kern, int ndim, const Shape& s, cudaStream_t st, int slot, size_t extra_smem, Args... args)
{
    TileGeom tg;
    tg.tiles_a = (s.O[0] + TILE - 1) / TILE;
    tg.tiles_f = (s.O[ndim - 1] + TILE - 1) / TILE;
    tg.mid = ndim == 3 ? s.O[1] : 1;
    const long long total = (long long)s.N * tg.mid * tg.tiles_f * tg.tiles_a;
    if (total >= (1LL << 31)) { set_error("interpolate: %lld tiles exceed 2^31", total); return kErrUnsupported; }
    const size_t smem = (size_t)kStages * ndim * TILE * PITCH * sizeof(float) + extra_smem;
    if (smem > 48 * 1024) CPAB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0, dev = 0, sms = 0;
    CPAB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem));
    CPAB_CUDA_OK(cudaGetDevice(&dev));
    CPAB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long blocks = (long long)sms * (per_sm > 0 ? per_sm : 1);
    if (g_interp_max_ctas > 0 && blocks > g_interp_max_ctas) blocks = g_interp_max_ctas;
    if (blocks > total) blocks = total;
    {   // gridDim.x in the radix (tiles_a, tiles_f, mid)
        long long q = blocks;
        tg.da = (int)(q % tg.tiles_a); q /= tg.tiles_a;
        tg.df = (int)(q % tg.tiles_f); q /= tg.tiles_f;
        tg.dm = (int)(q % tg.mid);     q /= tg.mid;
        tg.dn = (int)q;
    }
    prof_begin(slot, st);
    kern<<<(unsigned)blocks, 256, smem, st>>>(args..., s, tg);
    prof_end(slot, st);
    count_launch();
    CPAB_CUDA_OK(cudaGetLastError());
    return kOk;
}

template <typename T>
__global__ void __launch_bounds__(256)
k_interp_bwd_1d(const T* __restrict__ data, const T* __restrict__ grid, const T* __restrict__ gout,
                T* __restrict__ dgrid, T* __restrict__ ddata, Shape s)
{
    const int n = blockIdx.y;
    const int i0 = blockIdx.x * (256 * PT1D) + threadIdx.x;
    const T* gn = grid + (size_t)n * s.O[0];
    T gc[PT1D];
#pragma unroll
    for (int u = 0; u < PT1D; ++u) gc[u] = (i0 + 256 * u < s.O[0]) ? gn[i0 + 256 * u] : (T)0;
    int t0[PT1D], t1[PT1D];
    T w[PT1D], dg[PT1D];
#pragma unroll
    for (int u = 0; u < PT1D; ++u) { taps(gc[u], s.S[0], t0[u], t1[u], w[u]); dg[u] = 0; }
    for (int c = 0; c < s.C; ++c) {
        const size_t ch = (size_t)n * s.C + c;
        const T* dp = data + ch * s.S[0];
        const T* gp = gout + ch * s.O[0];
        T v[PT1D][2], g[PT1D];
#pragma unroll
        for (int u = 0; u < PT1D; ++u) {
            v[u][0] = __ldg(dp + t0[u]);
            v[u][1] = __ldg(dp + t1[u]);
            g[u] = (i0 + 256 * u < s.O[0]) ? gp[i0 + 256 * u] : (T)0;
        }
#pragma unroll
        for (int u = 0; u < PT1D; ++u) {
            T gv[2], dw[1];
            blend_vjp<1>(v[u], &w[u], g[u], gv, dw);
            dg[u] += dw[0];
            if (ddata != nullptr && i0 + 256 * u < s.O[0]) {
                atomicAdd(ddata + ch * s.S[0] + t0[u], gv[0]);
                atomicAdd(ddata + ch * s.S[0] + t1[u], gv[1]);
            }
        }
    }
    if (dgrid != nullptr) {
#pragma unroll
        for (int u = 0; u < PT1D; ++u)
            if (i0 + 256 * u < s.O[0]) dgrid[(size_t)n * s.O[0] + i0 + 256 * u] = dg[u] * (T)(s.S[0] - 1);
    }
}

template <typename T>
int interp_t(bool backward, int ndim, const Shape& s, const void* data, const void* grid,
             const void* gout, void* out_or_dgrid, void* ddata, cudaStream_t st)
{
    if (s.N == 0 || s.C == 0) return kOk;
    for (int j = 0; j < ndim; ++j) if (s.O[j] == 0) return kOk;
    if (sizeof(T) == 4) {
        for (int j = 0; j < ndim; ++j)
            if (s.S[j] > kMaxInterpExtentF32) { set_error("interpolate: float32 input extent %d exceeds %d", s.S[j], kMaxInterpExtentF32); return kErrUnsupported; }
    }
    if (ndim == 1) {
        dim3 g((unsigned)((s.O[0] + 256 * PT1D - 1) / (256 * PT1D)), (unsigned)s.N);
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
    {   // the kernels index one sample with 32-bit offsets
        long long gridpts = 1, inpts = s.C;
        for (int j = 0; j < ndim; ++j) { gridpts *= s.O[j]; inpts *= s.S[j]; }
        if (gridpts * ndim >= (1LL << 31) || inpts >= (1LL << 31) || gridpts * s.C >= (1LL << 31)) {
            set_error("interpolate: one sample exceeds 2^31 elements");
            return kErrUnsupported;
        }
    }
    if constexpr (sizeof(T) == 4) {
        // persistent, pipelined kernels (variants >= 5); they move 16-byte chunks of the grid, which
        // needs 16-byte aligned planes -- anything else takes the one-tile-per-CTA kernels
        const int var = g_interp_variant;
        const float* d = (const float*)data; const float* gr = (const float*)grid; const float* go = (const float*)gout;
        float* o = (float*)out_or_dgrid; float* dd = (float*)ddata;
        const bool aligned = s.O[0] % 4 == 0 && (reinterpret_cast<uintptr_t>(grid) & 15) == 0 &&
                             (!backward || out_or_dgrid == nullptr || (reinterpret_cast<uintptr_t>(out_or_dgrid) & 15) == 0);
        if (var >= 5 && aligned) {
#define FWD(ND, M) (s.C == 1 ? launch_pipe(k_interp_fwd_pipe<ND, M, true>, ndim, s, st, kProfInterpFwd, 0, d, gr, o) \
                             : launch_pipe(k_interp_fwd_pipe<ND, M, false>, ndim, s, st, kProfInterpFwd, 0, d, gr, o))
#define BWD(ND, B, M) launch_pipe(k_interp_bwd_pipe<ND, B, M>, ndim, s, st, kProfInterpBwd, 0, d, gr, go, o, dd)
            if (!backward && s.C == 1 && ndim == 2 && var >= 12) {      // measurement probes (see interp_fwd_compute)
                if (var == 12) return launch_pipe(k_interp_fwd_pipe<2, 4, true, 1>, ndim, s, st, kProfInterpFwd, 0, d, gr, o);
                if (var == 13) return launch_pipe(k_interp_fwd_pipe<2, 4, true, 2>, ndim, s, st, kProfInterpFwd, 0, d, gr, o);
                if (var == 15) return launch_pipe(k_interp_fwd_pipe<2, 4, true, 4>, ndim, s, st, kProfInterpFwd, 0, d, gr, o);
                return launch_pipe(k_interp_fwd_pipe<2, 4, true, 3>, ndim, s, st, kProfInterpFwd, 0, d, gr, o);
            }
            if (!backward && s.C == 1 && ndim == 2 && var >= 12) {      // measurement probes (see interp_fwd_compute)
                if (var == 12) return launch_pipe(k_interp_fwd_pipe<2, 4, true, 1>, ndim, s, st, kProfInterpFwd, 0, d, gr, o);
                if (var == 13) return launch_pipe(k_interp_fwd_pipe<2, 4, true, 2>, ndim, s, st, kProfInterpFwd, 0, d, gr, o);
                if (var == 15) return launch_pipe(k_interp_fwd_pipe<2, 4, true, 4>, ndim, s, st, kProfInterpFwd, 0, d, gr, o);
                return launch_pipe(k_interp_fwd_pipe<2, 4, true, 3>, ndim, s, st, kProfInterpFwd, 0, d, gr, o);
            }
            // 9-11: backward (single channel, d/dgrid only) with the register software pipeline; forward:
            // 9 = the pipe kernel (measured best, profiles/r02_interp_variants.txt), 10-11 = software pipeline
            if (backward && s.C == 1 && dd == nullptr && o != nullptr && var >= 9) {
#define BWDS(ND, M) launch_pipe(k_interp_bwd_sw<ND, M>, ndim, s, st, kProfInterpBwd, (size_t)ND * TILE * PITCH * sizeof(float), d, gr, go, o)
                if (ndim == 2) return var == 9 ? BWDS(2, 3) : var == 10 ? BWDS(2, 2) : BWDS(2, 4);
                return var == 9 ? BWDS(3, 2) : BWDS(3, 1);
#undef BWDS
            }
            if (!backward && s.C == 1 && var >= 10) {
#define FWDS(ND, M) launch_pipe(k_interp_fwd_sw<ND, M>, ndim, s, st, kProfInterpFwd, 0, d, gr, o)
                if (ndim == 2) return var == 10 ? FWDS(2, 3) : FWDS(2, 4);
                return var == 10 ? FWDS(3, 2) : FWDS(3, 1);
#undef FWDS
            }
            if (ndim == 2) {
                if (!backward) return (var == 5 || var >= 9) ? FWD(2, 4) : var == 6 ? FWD(2, 5) : var == 7 ? FWD(2, 6) : FWD(2, 3);
                return (var == 5 || var >= 9) ? BWD(2, 2, 4) : var == 6 ? BWD(2, 4, 3) : var == 7 ? BWD(2, 1, 5) : BWD(2, 4, 2);
            }
            if (!backward) return (var == 5 || var >= 9) ? FWD(3, 3) : var == 6 ? FWD(3, 4) : var == 7 ? FWD(3, 2) : FWD(3, 5);
            return (var == 5 || var >= 9) ? BWD(3, 1, 3) : var == 6 ? BWD(3, 2, 2) : var == 7 ? BWD(3, 1, 4) : BWD(3, 4, 1);
#undef FWD
#undef BWD
        }
    }
    const long z = (long)s.N * (ndim == 3 ? s.O[1] : 1);
    const unsigned gy = (unsigned)((s.O[ndim - 1] + TILE - 1) / TILE);
    const int mid1 = ndim == 3 ? s.O[1] : 1;
    if (gy > 65535 || mid1 > 65535) { set_error("interpolate: output extent exceeds the grid limits of the tile kernels"); return kErrUnsupported; }
    if (z > 65535) {     // gridDim.z limit of the one-tile-per-CTA kernels: slab the batch (as the 1-D path does)
        const int per = 65535 / mid1;
        size_t in_plane = (size_t)s.C, out_plane = (size_t)s.C, pts = 1;
        for (int j = 0; j < ndim; ++j) { in_plane *= s.S[j]; out_plane *= s.O[j]; pts *= s.O[j]; }
        for (int n0 = 0; n0 < s.N; n0 += per) {
            Shape sub = s;
            sub.N = s.N - n0 < per ? s.N - n0 : per;
            const size_t din = (size_t)n0 * in_plane, dout = (size_t)n0 * out_plane, dgr = (size_t)n0 * ndim * pts;
            const int rc = interp_t<T>(backward, ndim, sub, (const T*)data + din, (const T*)grid + dgr,
                                       gout ? (const T*)gout + dout : nullptr,
                                       out_or_dgrid ? (T*)out_or_dgrid + (backward ? dgr : dout) : nullptr,
                                       ddata ? (T*)ddata + din : nullptr, st);
            if (rc != kOk) return rc;
        }
        return kOk;
    }
    dim3 g((unsigned)((s.O[0] + TILE - 1) / TILE), gy, (unsigned)z);
    prof_begin(backward ? kProfInterpBwd : kProfInterpFwd, st);
    if (ndim == 2) {
        if (!backward) {
            // (points in flight per thread, resident CTAs targeted); float32 only -- the check mode
            // keeps the plain configuration
            const int var = sizeof(T) == 4 ? g_interp_variant : 0;
            if (var == 1) k_interp_fwd<T, 2, 1, 6><<<g, 256, 0, st>>>((const T*)data, (const T*)grid, (T*)out_or_dgrid, s);
            else if (var == 2 && s.C == 1) k_interp_fwd<T, 2, 1, 8, true><<<g, 256, 0, st>>>((const T*)data, (const T*)grid, (T*)out_or_dgrid, s);
            else if (var == 2) k_interp_fwd<T, 2, 1, 8><<<g, 256, 0, st>>>((const T*)data, (const T*)grid, (T*)out_or_dgrid, s);
            else if (var == 3) k_interp_fwd<T, 2, 2, 5><<<g, 256, 0, st>>>((const T*)data, (const T*)grid, (T*)out_or_dgrid, s);
            else if (var == 4) k_interp_fwd<T, 2, 2, 4><<<g, 256, 0, st>>>((const T*)data, (const T*)grid, (T*)out_or_dgrid, s);
            else k_interp_fwd<T, 2, 4, 1><<<g, 256, 0, st>>>((const T*)data, (const T*)grid, (T*)out_or_dgrid, s);
        }
        else if (sizeof(T) == 4 && g_interp_variant == 0) k_interp_bwd<T, 2, 2, 3><<<g, 256, 0, st>>>((const T*)data, (const T*)grid, (const T*)gout, (T*)out_or_dgrid, (T*)ddata, s);
        else if (sizeof(T) == 4) k_interp_bwd<T, 2, 1, 5><<<g, 256, 0, st>>>((const T*)data, (const T*)grid, (const T*)gout, (T*)out_or_dgrid, (T*)ddata, s);
        else k_interp_bwd<T, 2, 1, 1><<<g, 256, 0, st>>>((const T*)data, (const T*)grid, (const T*)gout, (T*)out_or_dgrid, (T*)ddata, s);
    } else {
        if (!backward) {
            if (sizeof(T) == 4 && g_interp_variant >= 1) k_interp_fwd<T, 3, 1, 4><<<g, 256, 0, st>>>((const T*)data, (const T*)grid, (T*)out_or_dgrid, s);
            else k_interp_fwd<T, 3, 2, 1><<<g, 256, 0, st>>>((const T*)data, (const T*)grid, (T*)out_or_dgrid, s);
        }
        else if (sizeof(T) == 4) k_interp_bwd<T, 3, 1, 3><<<g, 256, 0, st>>>((const T*)data, (const T*)grid, (const T*)gout, (T*)out_or_dgrid, (T*)ddata, s);
        else k_interp_bwd<T, 3, 1, 1><<<g, 256, 0, st>>>((const T*)data, (const T*)grid, (const T*)gout, (T*)out_or_dgrid, (T*)ddata, s);
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

void set_interp_variant(int v) { g_interp_variant = v; }
void set_interp_max_ctas(int v) { g_interp_max_ctas = v; }

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
