// cpab_adjoint.cuh -- the theta-gradient as an adjoint sweep (sm_100a).  Instantiated per dimension
// in cpab_adjoint_{1,2,3}d.cu (three translation units, so that they compile in parallel).
//
// Replaces libcpab/core/cpab_ops.cu:390-697 (per (point, theta, k) RK2 sensitivity kernels) AND the
// contraction libcpab/pytorch/transformer.py:201-202; CPU semantics: libcpab/core/cpab_ops.cpp:262-372.
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
//
// ---- RK2 step records ---------------------------------------------------------------------------
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
//
// ---- cell-sequence certificate (float32 default; CPAB_FLAG_FAST_GRAD switches it off) ------------
// The discretised flow is piecewise affine in the point, so the gradient jumps when an iterate is
// assigned to a neighbouring simplex.  The record arithmetic above is not the reference's
// (`p += vMid*h` with a double h, cpab_ops.cpp:289-366): its iterates differ from the reference's
// by ulps, and an iterate within that distance of a face can land in the other simplex -- about
// once per 1e6 (point, step) events, moving that theta's gradient by 1e-5..1e-3 relative.
// Reproducing the reference's double-rounded updates for every trajectory would cost ten
// F2F.F64<->F32 conversions per 2-D step, and those run at 16 lanes/clk/SM on a B200
// (profiles/r02_pipe_probe.txt) -- 80 SMSP-cycles per warp-step against ~95 for the whole sweep.
// Instead, pass 1 integrates with the records and carries a certificate:
//     Let f_n be the record iterate, r_n the reference iterate (f_0 = r_0), both inside the unit
//     box.  While both have visited the same cells c_0..c_{n-1},
//         |f_{n+1} - r_{n+1}|_inf <= ||I + D_c||_inf |f_n - r_n|_inf + eta,
//     ||I + D_c||_inf = 1 + k_c,  k_c = max_i (D_ii + sum_{j != i} |D_ij|)   (|D_ii| < 1),
//     eta = 2^-24 (1 + h (15 a + 9 tau)(1 + h a / 2)): half an ulp for each scheme's final
//     rounding plus the rounding of the increments; a bounds ||L_c||_inf over the cells of this
//     theta, tau bounds |t_c|_inf, h = 1/nsteps (derivation: DESIGN.md 2).  The kernel carries
//     the bound along the trajectory with the record it has in registers anyway,
//         M_{n+1} = (1 + k_{c_n}) M_n + eta * cert_scale,   M_0 = cert_floor,
//     in the local units of find_cell_near(): that function returns, with the fast-path cell of
//     f_n, a lower bound `dist` of the distance of f_n from the nearest face of that simplex;
//     a perturbation of |dp|_inf moves it by at most |dp| * cert_scale, and cert_floor covers the
//     search's own rounding.  If dist >= M_n at every step n >= 1 (step 0 uses the complete
//     search on the bit-identical input), then by induction r_n lies in the same simplex as f_n
//     at every step: the recorded cell sequence IS the reference's.
// A trajectory that fails the test at some step (a few per cent: iterates that land next to a
// face while crossing it, points parked on the domain boundary) is marked in a bit mask and
// re-integrated by a second kernel (k_backward_redo) with the reference's own arithmetic (float A p~ in the reference's
// order, double-rounded updates, the complete cell search) -- bit-identical iterates, hence the
// reference's cell sequence again.  Pass 2 is shared.  The gradient then differs from the
// reference's only by summation order and record rounding, never by a different cell.
#pragma once

#include "cpab_device.cuh"

namespace cpab {

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

// One RK2 step with the reference's own arithmetic (cpab_ops.cpp:304-313,362-365): float A p~ in
// the reference's order, `pMid = p + h*v/2.0` and `p += vMid*h` evaluated in double (h is a
// double there) and rounded on store.  (h*v)/2.0 == (h/2)*v bit for bit: halving is exact in binary
// arithmetic (no underflow at these magnitudes), so it commutes with the rounding of the product.
template <int NDIM>
__device__ __forceinline__ void step_reference(const float* A, double h, float* p)
{
    float v[NDIM], pm[NDIM], vm[NDIM];
    affine_strict<NDIM>(A, p, v);
#pragma unroll
    for (int j = 0; j < NDIM; ++j)
        pm[j] = (float)__dadd_rn((double)p[j], __dmul_rn(0.5 * h, (double)v[j]));
    affine_strict<NDIM>(A, pm, vm);
#pragma unroll
    for (int j = 0; j < NDIM; ++j) p[j] = (float)__dadd_rn((double)p[j], __dmul_rn((double)vm[j], h));
}
template <int NDIM>
__device__ __forceinline__ void step_reference(const double* A, double h, double* p)
{
    double v[NDIM], pm[NDIM], vm[NDIM];
    affine_strict<NDIM>(A, p, v);
#pragma unroll
    for (int j = 0; j < NDIM; ++j) pm[j] = __dadd_rn(p[j], __dmul_rn(0.5 * h, v[j]));
    affine_strict<NDIM>(A, pm, vm);
#pragma unroll
    for (int j = 0; j < NDIM; ++j) p[j] = __dadd_rn(p[j], __dmul_rn(vm[j], h));
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

// Certificate inputs of one launch (float32 only).
struct CertArgs {
    const void* As;        // [n_theta,nC,n,n+1] velocity fields, for the reference-arithmetic re-integration
    const float* stats;    // [n_theta][2]: max_c ||L_c||_inf and max_c |t_c|_inf (k_prepare_backward)
    float scale, floor;    // cert_scale(g), cert_floor(g)
    unsigned* mask;        // [n_theta][words] one bit per trajectory: certificate failed, re-integrate
    long words;            // 32-bit words per theta: ceil(nP / 32)
    int* flagged;          // optional [n_theta] counters of re-integrated trajectories (diagnostics), may be NULL
};

// eta * cert_scale of one theta: the per-step additive term of the certificate bound (see the header)
__device__ __forceinline__ float cert_eta(const CertArgs& ca, int theta, int nsteps)
{
    const float a = ca.stats[2 * theta], tau = ca.stats[2 * theta + 1];
    const float h = 1.0f / (float)nsteps;
    // |D_ii| < 1 is what makes ||I + D||_inf = 1 + k; a theta this wild (or NaN) certifies nothing
    if (!(h * a < 0.5f)) return __int_as_float(0x7f800000);
    // + 6e-8: k_c is evaluated from the rounded record (|error| < 1e-7) and multiplies a bound that is
    // below 1/2 for every trajectory still certified (no point is further than that from a face)
    return 1.001f * 5.9604645e-08f * (1.0f + h * (15.0f * a + 9.0f * tau) * (1.0f + 0.5f * h * a)) * ca.scale + 6e-8f;
}
// k_c of a step record held in registers: max_i (D_ii + sum_{j != i} |D_ij|), D column-major
template <int NDIM>
__device__ __forceinline__ float cert_gain(const float* W)
{
    if (NDIM == 1) return W[0];
    if (NDIM == 2) return fmaxf(W[0] + fabsf(W[2]), W[3] + fabsf(W[1]));
    return fmaxf(fmaxf(W[0] + (fabsf(W[3]) + fabsf(W[6])), W[4] + (fabsf(W[1]) + fabsf(W[7]))),
                 W[8] + (fabsf(W[2]) + fabsf(W[5])));
}
template <int NDIM> __device__ __forceinline__ float cert_gain(const double*) { return 0.0f; }

// =====================================================================================================
// The two passes over one trajectory, shared by k_backward (all trajectories, step records) and
// k_backward_redo (the trajectories whose certificate failed, reference arithmetic).
// A thread owns one `slot` of the checkpoint / cell-trace arrays in shared memory (STRIDE slots).
// =====================================================================================================
template <typename T, int NDIM, int SEG>
struct Sweep {
    static constexpr int PPC = Dim<NDIM>::kPpc;
    static constexpr int WS = StepRec<NDIM>::kStride;
    T* ck;                 // checkpoints [nseg][NDIM][stride]
    unsigned short* ct16;  // cell trace  [nsteps][stride]  (ct32 when `wide`)
    int* ct32;
    bool wide;
    int stride, slot;
    int nsteps, nseg;

    // ---- pass 1: the RK2 trajectory from p.  Records the cell of every step and a checkpoint of p
    //      at the start of every segment; the only pass that searches cells.
    //      MODE 0: records + complete search (no certificate: float64, CPAB_FLAG_FAST_GRAD)
    //      MODE 1: records + certified fast search; returns true (and stops early) when it fails
    //      MODE 2: the reference's arithmetic + complete search
    template <int MODE, typename Table>
    __device__ __forceinline__ bool pass1(const Geom& g, const Table& table, const T* Ag, float m0, float eta_s,
                                          float magic, T* p) const
    {
        bool failed = false;
        float m = m0;                  // certificate bound M_n of the current step (local units)
        int c_first = 0;
        // MODE 2 reads its velocity fields through L1 (every lane its own theta): the block of the
        // current cell stays in registers and is re-fetched only when the trajectory changes cell
        // (1-2 % of the steps) -- the dependent global load per step was the kernel's main stall
        T a_held[PPC];
        int c_held = -1;
        if (MODE == 1) c_first = find_cell<NDIM>(p, g);
        auto segment = [&](int sg, auto full_tag) {     // FULL = a whole segment that is not the last one: no bounds checks
            constexpr bool FULL = decltype(full_tag)::value;
#pragma unroll
            for (int j = 0; j < NDIM; ++j) ck[(sg * NDIM + j) * stride + slot] = p[j];
#pragma unroll
            for (int s = 0; s < SEG; ++s) {
                const int n = sg * SEG + s;
                if (FULL || n < nsteps) {
                    int c;
                    if constexpr (MODE == 1) {
                        float dist;
                        c = find_cell_near<NDIM>(p, g, magic, dist);
                        if (s == 0 && sg == 0) { c = c_first; dist = 1.0f; }   // step 0: identical input, complete search
                        failed |= dist < m;
                    } else {
                        c = find_cell<NDIM>(p, g);
                    }
                    if (wide) ct32[n * stride + slot] = c;
                    else ct16[n * stride + slot] = (unsigned short)c;
                    if (FULL || n + 1 < nsteps) {
                        if constexpr (MODE == 2) {
                            if (c != c_held) { load_affine<NDIM>(Ag + (size_t)c * PPC, a_held); c_held = c; }
                            step_reference<NDIM>(a_held, 1.0 / nsteps, p);
                        } else {
                            T w[WS];
                            table.load(c, w);
                            step_inc<NDIM>(w, p);
                            if constexpr (MODE == 1) m = fmaf(m, cert_gain<NDIM>(w), m + eta_s);
                        }
                    }
                }
            }
        };
        int sg = 0;
        for (; sg + 1 < nseg && !failed; ++sg) segment(sg, std::true_type{});
        if (!failed) segment(nseg - 1, std::false_type{});
        return failed;
    }

    // ---- pass 2: segments in reverse; replay p into registers (no search), sweep lambda back,
    //      accumulate R_c += lambda_{n+1} [p_n;1]^T per thread and hand a cell's sum to R[theta]
    //      (Gt) when the trajectory leaves the cell.  Leaves lambda_0 in lam.
    //      KEEP: the record of the current cell is held in registers across steps and re-fetched only
    //      on a cell change (tables read through L1, k_backward_redo); staged tables re-read it.
    template <bool KEEP = false, typename Table>
    __device__ __forceinline__ void pass2(const Table& table, T* Gt, T* lam, T* acc, int& cur) const
    {
        T p[NDIM];
        T wk[WS];
        int wc = -1;
        auto fetch = [&](int c, T* w) {
            if constexpr (KEEP) {
                if (c != wc) { table.load(c, wk); wc = c; }
#pragma unroll
                for (int e = 0; e < WS; ++e) w[e] = wk[e];
            } else {
                table.load(c, w);
            }
        };
        auto segment = [&](int sg, auto full_tag) {
            constexpr bool FULL = decltype(full_tag)::value;
            const int len = FULL ? SEG : nsteps - sg * SEG;
            T ps[SEG][NDIM];
            T w[WS];
            int cs[SEG];
#pragma unroll
            for (int j = 0; j < NDIM; ++j) p[j] = ck[(sg * NDIM + j) * stride + slot];
#pragma unroll
            for (int s = 0; s < SEG; ++s) {
                if (FULL || s < len) {
                    const int n = sg * SEG + s;
                    cs[s] = wide ? ct32[n * stride + slot] : (int)ct16[n * stride + slot];
#pragma unroll
                    for (int j = 0; j < NDIM; ++j) ps[s][j] = p[j];
                    if (s + 1 < SEG && (FULL || s + 1 < len)) {
                        fetch(cs[s], w);            // (the compiler keeps these for the sweep below)
                        step_inc<NDIM>(w, p);
                    }
                }
            }
#pragma unroll
            for (int s = SEG - 1; s >= 0; --s) {
                if (FULL || s < len) {
                    const int c = cs[s];
                    fetch(c, w);
                    if (c != cur) {                 // left a cell: hand its sum to R[theta]
                        if (cur >= 0) red_cell<PPC>(Gt + (size_t)cur * PPC, acc);
#pragma unroll
                        for (int e = 0; e < PPC; ++e) acc[e] = 0;
                        cur = c;
                    }
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
        segment(nseg - 1, std::false_type{});
        for (int sg = nseg - 2; sg >= 0; --sg) segment(sg, std::true_type{});
    }
};

// the trajectory's start point and lambda_N = dL/dp_N.  SAMPLE: `gout` holds the transformed grid
// (output of the forward) and the upstream gradient is that of the sampled image, `gimg` (fused
// transform_data): lambda_N is formed here.
template <typename T, int NDIM, bool SAMPLE>
__device__ __forceinline__ void load_trajectory(const T* __restrict__ points, const T* __restrict__ gout, long nP,
                                                int broadcast, int theta, long i, const T* __restrict__ data,
                                                const T* __restrict__ gimg, const Shape& sh, T* p, T* lam)
{
    const T* src = points + (broadcast ? (size_t)theta * NDIM * nP : 0);
    const T* gsrc = gout + (size_t)theta * NDIM * nP;
#pragma unroll
    for (int j = 0; j < NDIM; ++j) { p[j] = src[i + (long)j * nP]; lam[j] = gsrc[i + (long)j * nP]; }
    if (SAMPLE) {
        T pt[NDIM];
#pragma unroll
        for (int j = 0; j < NDIM; ++j) pt[j] = lam[j];
        sample_vjp<T, NDIM>(pt, theta, i, data, gimg, sh, lam);
    }
}

#ifndef CPAB_BWD_REGS
#define CPAB_BWD_REGS 80
#endif
// resident CTAs per SM the register allocation is tuned for (float: 80 regs in 1-D/2-D, 128 in 3-D -- fewer registers spill)
template <typename T, int NDIM, int SEG, int BLOCK> struct BwdOcc {
    static constexpr int kRegs = NDIM == 3 ? (SEG <= 3 ? 102 : 128) : CPAB_BWD_REGS;
    static constexpr int kMinBlocks = sizeof(T) == 8 ? 1 : 65536 / (BLOCK * kRegs);
};

// CERT: pass 1 carries the cell-sequence certificate; a trajectory that fails it is skipped here
// and marked in a bit mask (one 32-bit word per warp and iteration: the ballot), from which
// k_backward_redo re-integrates it with the reference's arithmetic (float32 only; see the header
// of this file).  Keeping that cold, conversion-bound code out of this kernel keeps its register
// budget and its unit barriers as they are without the certificate.
template <typename T, int NDIM, int SEG, bool SMEM, int BLOCK, bool SAMPLE, bool CERT>
__global__ void __launch_bounds__(BLOCK, (BwdOcc<T, NDIM, SEG, BLOCK>::kMinBlocks))
k_backward(const T* __restrict__ points, const T* __restrict__ Ws, const T* __restrict__ gout,
           T* __restrict__ G, T* __restrict__ dpoints, long nP, int broadcast, int nsteps,
           const __grid_constant__ Geom g, const __grid_constant__ WorkPlan wp, unsigned* __restrict__ counter,
           const T* __restrict__ data, const T* __restrict__ gimg, const __grid_constant__ Shape sh,
           const __grid_constant__ CertArgs ca)
{
    static_assert(!CERT || sizeof(T) == 4, "the certificate is a float32 device");
    constexpr int PPC = Dim<NDIM>::kPpc;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned s_work;
    __shared__ float s_eta;
    constexpr int WS = StepRec<NDIM>::kStride;
    const int tsize = g.n_cells * PPC;
    const int wsize = g.n_cells * WS;

    // shared layout: [step records] (if SMEM), checkpoints [nseg][NDIM][BLOCK], cell trace
    // [nsteps][BLOCK] (16-bit; 32-bit only for tessellations of >= 65536 simplices, which never
    // fit the staged path)
    T* sW = reinterpret_cast<T*>(smem_raw);
    Sweep<T, NDIM, SEG> sw;
    sw.nsteps = nsteps;
    sw.nseg = (nsteps + SEG - 1) / SEG;
    sw.ck = sW + (SMEM ? wsize : 0);
    sw.ct16 = reinterpret_cast<unsigned short*>(sw.ck + (size_t)sw.nseg * NDIM * BLOCK);
    sw.ct32 = reinterpret_cast<int*>(sw.ct16);
    sw.wide = !SMEM && g.n_cells > 65535;
    sw.stride = BLOCK;
    sw.slot = threadIdx.x;
    CellTable<T, NDIM, SMEM, WS> tab;
    tab.saddr = SMEM ? (uint32_t)__cvta_generic_to_shared(sW) & 0xffffffu : 0;   // CTA-local offset (no cluster launch: rank bits are 0)
    float magic = 12582912.0f;        // 1.5 * 2^23, rounding constant of the cell search
    asm volatile("" : "+f"(magic));
    int staged = -1;
    WorkUnit wu;
    for (;;) {
    // ---- next work unit (all threads): drawn from the counter in the caller's workspace
    __syncthreads();                                   // everyone is done with the previous unit
    if (threadIdx.x == 0) s_work = atomicAdd(counter, 1u);
    __syncthreads();
    if (s_work >= wp.total) break;
    unit_of(wp, s_work, nP, wu);
    const int theta = wu.theta;
    const long begin = wu.begin, end = wu.end;
    tab.gptr = Ws + (size_t)theta * wsize;
    if (SMEM && theta != staged) {           // (synchronised above: nobody reads the old table any more)
        stage_block(sW, tab.gptr, wsize);
        staged = theta;
    }
    if (CERT && threadIdx.x == 0) s_eta = cert_eta(ca, theta, nsteps);
    if (SMEM || CERT) __syncthreads();
    const float eta_s = CERT ? s_eta : 0.0f;
    T* Gg = G + (size_t)theta * tsize;
    // keep the base in registers: the flush blocks run divergently, often, and would otherwise
    // rebuild it from the kernel parameters (11 uniform-datapath instructions per occurrence)
    asm volatile("" : "+l"(Gg));

    for (long base = begin; base < end; base += BLOCK) {      // warp-uniform trip count
        const long i = base + threadIdx.x;
        const bool valid = i < end;
        T acc[PPC];
        int cur = -1;
        bool failed = false;
#pragma unroll
        for (int e = 0; e < PPC; ++e) acc[e] = 0;
        if (valid) {
            T p[NDIM], lam[NDIM];
            load_trajectory<T, NDIM, SAMPLE>(points, gout, nP, broadcast, theta, i, data, gimg, sh, p, lam);
            failed = sw.template pass1<CERT ? 1 : 0>(g, tab, nullptr, ca.floor, eta_s, magic, p);
            if (!failed) {
                sw.pass2(tab, Gg, lam, acc, cur);
                if (dpoints != nullptr) {
                    T* dp = dpoints + (size_t)theta * NDIM * nP;
#pragma unroll
                    for (int j = 0; j < NDIM; ++j) dp[i + (long)j * nP] = lam[j];
                }
            }
        }
        flush_runs<T, PPC>(Gg, cur, acc);
        if (CERT) {
            // units begin at multiples of 256 points, so a warp's 32 points are exactly one mask
            // word; a warp wholly beyond `end` owns none
            const unsigned word = __ballot_sync(0xffffffffu, failed);
            if ((threadIdx.x & 31) == 0 && i < end) ca.mask[(size_t)theta * ca.words + (i >> 5)] = word;
        }
    }
    }   // work units
}

// Re-integration of the trajectories whose certificate failed, with the reference's arithmetic.
// Every warp is an independent worker: it draws chunks of 8 mask words (256 trajectories of one
// theta) from a counter, expands the set bits into a list in shared memory and, whenever 32 are
// waiting, integrates them as one full-width iteration -- pass 1 in mode 2 (the reference's float
// A p~ and double-rounded updates, complete cell search), pass 2 as in k_backward but with the step
// records of each lane's own theta read through L1.  The loop is bound by the latency of the
// F2F.F64<->F32 conversions (16 lanes/clk/SM); it touches a few per cent of the trajectories.
// 6 CTAs x 128 threads per SM (80 registers): 4 CTAs (109 registers, no spills) run as fast, 8 (64
// registers) spill and run 30 % slower; chunks of 4 mask words: 1 or 2 cost more draws than they
// save in tail (0.69 / 0.77 / 1.01 ms for 4 / 2 / 1 words on 128 thetas x 512^2)
#ifndef CPAB_REDO_CTAS
#define CPAB_REDO_CTAS 4
#endif
#ifndef CPAB_REDO_CHUNK
#define CPAB_REDO_CHUNK 4
#endif
constexpr int kRedoChunkWords = CPAB_REDO_CHUNK;
template <int NDIM, int SEG, bool SAMPLE>
__global__ void __launch_bounds__(128, CPAB_REDO_CTAS)
k_backward_redo(const float* __restrict__ points, const float* __restrict__ Ws, const float* __restrict__ gout,
                float* __restrict__ G, float* __restrict__ dpoints, long nP, int broadcast, int nsteps,
                int n_theta, const __grid_constant__ Geom g, unsigned* __restrict__ counter,
                const float* __restrict__ data, const float* __restrict__ gimg, const __grid_constant__ Shape sh,
                const __grid_constant__ CertArgs ca)
{
    using T = float;
    constexpr int PPC = Dim<NDIM>::kPpc;
    constexpr int WS = StepRec<NDIM>::kStride;
    constexpr int kChunkWords = kRedoChunkWords;
    constexpr int kListCap = 31 + 32 * kChunkWords;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tsize = g.n_cells * PPC;
    const int wsize = g.n_cells * WS;
    Sweep<T, NDIM, SEG> sw;
    sw.nsteps = nsteps;
    sw.nseg = (nsteps + SEG - 1) / SEG;
    sw.wide = g.n_cells > 65535;
    sw.stride = 32;
    sw.slot = lane;
    // per warp: list [kListCap] x (theta, point) | checkpoints [nseg][NDIM][32] | cell trace [nsteps][32]
    const size_t per_warp = (((size_t)kListCap * 8 + (size_t)sw.nseg * NDIM * 32 * 4 +
                              (size_t)nsteps * 32 * (sw.wide ? 4 : 2)) + 15) & ~(size_t)15;
    unsigned char* mine = smem_raw + per_warp * warp;
    uint2* list = reinterpret_cast<uint2*>(mine);
    sw.ck = reinterpret_cast<T*>(mine + (size_t)kListCap * 8);
    sw.ct16 = reinterpret_cast<unsigned short*>(sw.ck + (size_t)sw.nseg * NDIM * 32);
    sw.ct32 = reinterpret_cast<int*>(sw.ct16);
    const long chunks_per_theta = (ca.words + kChunkWords - 1) / kChunkWords;
    const long total = chunks_per_theta * n_theta;
    int n_list = 0;                   // warp-uniform

    auto redo = [&](int first, int n) {
        if (lane < n) {
            const uint2 e = list[first + lane];
            const int theta = (int)e.x;
            const long i = (long)e.y;
            T p[NDIM], lam[NDIM], acc[PPC];
            int cur = -1;
#pragma unroll
            for (int k = 0; k < PPC; ++k) acc[k] = 0;
            load_trajectory<T, NDIM, SAMPLE>(points, gout, nP, broadcast, theta, i, data, gimg, sh, p, lam);
            CellTable<T, NDIM, false, WS> gtab;      // this lane's theta: records through L1
            gtab.gptr = Ws + (size_t)theta * wsize;
            gtab.saddr = 0;
            const T* Ag = reinterpret_cast<const T*>(ca.As) + (size_t)theta * tsize;
            sw.template pass1<2>(g, gtab, Ag, 0.0f, 0.0f, 0.0f, p);
            T* Gt = G + (size_t)theta * tsize;
            sw.template pass2<true>(gtab, Gt, lam, acc, cur);
            if (dpoints != nullptr) {
                T* dp = dpoints + (size_t)theta * NDIM * nP;
#pragma unroll
                for (int j = 0; j < NDIM; ++j) dp[i + (long)j * nP] = lam[j];
            }
            if (cur >= 0) red_cell<PPC>(Gt + (size_t)cur * PPC, acc);
            if (ca.flagged != nullptr) atomicAdd(ca.flagged + theta, 1);
        }
        __syncwarp();
    };

    for (;;) {
        long chunk = 0;
        if (lane == 0) chunk = (long)atomicAdd(counter, 1u);
        chunk = __shfl_sync(0xffffffffu, chunk, 0);
        if (chunk >= total) break;
        const int theta = (int)(chunk / chunks_per_theta);
        const long w0 = (chunk - (long)theta * chunks_per_theta) * kChunkWords;
        unsigned word = 0;
        if (lane < kChunkWords && w0 + lane < ca.words) word = ca.mask[(size_t)theta * ca.words + w0 + lane];
        // exclusive prefix of the bit counts over the (8) loading lanes
        const int cnt = __popc(word);
        int pre = cnt;
#pragma unroll
        for (int off = 1; off < kChunkWords; off <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, pre, off);
            if (lane >= off) pre += t;
        }
        const int all = __shfl_sync(0xffffffffu, pre, kChunkWords - 1);
        if (all == 0) continue;
        int at = n_list + pre - cnt;
        while (word != 0) {
            const int b = __ffs(word) - 1;
            word &= word - 1;
            list[at++] = make_uint2((unsigned)theta, (unsigned)((w0 + lane) * 32 + b));
        }
        n_list += all;
        __syncwarp();
        while (n_list >= 32) { n_list -= 32; redo(n_list, 32); }
    }
    if (n_list > 0) redo(0, n_list);
}

// Per (theta, cell): the RK2 step record (see step_inc) from A_c = [L | t], and a zeroed R_c block.
// Evaluated in double and rounded once.  The origin is the centre of the cell's square / cube (any
// float near the cell serves: s' is formed from the rounded value).  Also: the work counter of
// the launch sequence is zeroed, and (float32) the per-theta maxima the certificate needs are
// gathered: stats[theta] = (max_c ||L_c||_inf, max_c |t_c|_inf), zeroed by a memset node before.
template <typename T, int NDIM>
__global__ void __launch_bounds__(256)
k_prepare_backward(const T* __restrict__ As, T* __restrict__ Ws, T* __restrict__ R, long n_blocks, int nsteps,
                   const __grid_constant__ Geom g, unsigned* __restrict__ counter, float* __restrict__ stats)
{
    constexpr int PPC = Dim<NDIM>::kPpc;
    constexpr int M = NDIM + 1;
    constexpr int WS = StepRec<NDIM>::kStride;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { counter[0] = 0; counter[1] = 0; }
    if (i >= n_blocks) return;
    double A[PPC];
#pragma unroll
    for (int e = 0; e < PPC; ++e) { A[e] = (double)As[i * PPC + e]; R[i * PPC + e] = 0; }
    if (stats != nullptr) {
        float a = 0.0f, tau = 0.0f;
#pragma unroll
        for (int r = 0; r < NDIM; ++r) {
            float row = 0.0f;
#pragma unroll
            for (int k = 0; k < NDIM; ++k) row += fabsf((float)A[r * M + k]);
            a = fmaxf(a, row * 1.000001f);          // (the float sum rounds: keep it an upper bound)
            tau = fmaxf(tau, fabsf((float)A[r * M + NDIM]));
        }
        // non-negative floats order like their bit patterns; NaN/inf patterns sort above every finite value
        const long theta = i / g.n_cells;
        atomicMax(reinterpret_cast<unsigned*>(stats) + 2 * theta, __float_as_uint(a));
        atomicMax(reinterpret_cast<unsigned*>(stats) + 2 * theta + 1, __float_as_uint(tau));
    }
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

// ---- reference cell trace (tests, diagnostics) -----------------------------------------------------
// The cell sequence of the RK2 trajectory each of the three pass-1 modes records, one launch per
// mode: cells [n_theta][nsteps][nP] (int32), failed [n_theta][nP] (uint8, mode 1 only).  Lets the
// tests check on the device that (a) the reference-arithmetic mode reproduces the oracle's
// trajectory cells bit for bit and (b) every trajectory the certificate passes has those cells.
template <int NDIM>
__global__ void __launch_bounds__(128)
k_rk2_trace(const float* __restrict__ points, const float* __restrict__ As, const float* __restrict__ Ws,
            const float* __restrict__ stats, int* __restrict__ cells, unsigned char* __restrict__ failed_out,
            long nP, int n_theta, int broadcast, int nsteps, int mode, const __grid_constant__ Geom g)
{
    constexpr int PPC = Dim<NDIM>::kPpc;
    constexpr int WS = StepRec<NDIM>::kStride;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int theta = blockIdx.y;
    if (i >= nP) return;
    const float* src = points + (broadcast ? (size_t)theta * NDIM * nP : 0);
    float p[NDIM];
#pragma unroll
    for (int j = 0; j < NDIM; ++j) p[j] = src[i + (long)j * nP];
    CertArgs ca{As, stats, cert_scale(g), cert_floor(g), nullptr, 0, nullptr};
    const float eta_s = mode == 1 ? cert_eta(ca, theta, nsteps) : 0.0f;
    float m = ca.floor;
    bool failed = false;
    const float magic = 12582912.0f;
    int* out = cells + (size_t)theta * nsteps * nP + i;
    for (int n = 0; n < nsteps; ++n) {
        int c;
        if (mode == 1 && n > 0) {
            float dist;
            c = find_cell_near<NDIM>(p, g, magic, dist);
            failed |= dist < m;
        } else {
            c = find_cell<NDIM>(p, g);
        }
        out[(size_t)n * nP] = c;
        if (mode == 2) {
            float a[PPC];
            load_affine<NDIM>(As + ((size_t)theta * g.n_cells + c) * PPC, a);
            step_reference<NDIM>(a, 1.0 / nsteps, p);
        } else {
            float w[WS];
#pragma unroll
            for (int e = 0; e < WS; ++e) w[e] = Ws[((size_t)theta * g.n_cells + c) * WS + e];
            step_inc<NDIM>(w, p);
            m = fmaf(m, cert_gain<NDIM>(w), m + eta_s);
        }
    }
    if (failed_out != nullptr) failed_out[(size_t)theta * nP + i] = failed ? 1 : 0;
}

// =====================================================================================================
// host side
// =====================================================================================================
// workspace layout (all offsets 16-byte aligned):
//   G [n_theta, D] (accumulated as R, converted in place) | step records W [n_theta, nC, stride] |
//   certificate stats [n_theta][2] float | work counters [4] unsigned |
//   (float32) failed-certificate mask [n_theta][ceil(nP/32)] unsigned
struct BwdLayout {
    size_t off_w, off_stats, off_counter, off_mask, total;
    long words;
};
template <int NDIM>
inline BwdLayout backward_layout(size_t elt, const Geom& g, int n_theta, long nP)
{
    auto up = [](size_t x) { return (x + 15) & ~(size_t)15; };
    BwdLayout l;
    l.off_w = up((size_t)n_theta * g.n_cells * Dim<NDIM>::kPpc * elt);
    l.off_stats = up(l.off_w + (size_t)n_theta * g.n_cells * StepRec<NDIM>::kStride * elt);
    l.off_counter = up(l.off_stats + (size_t)n_theta * 2 * sizeof(float));
    l.off_mask = l.off_counter + 16;
    l.words = (nP + 31) / 32;
    l.total = up(l.off_mask + (elt == 4 ? (size_t)n_theta * l.words * 4 : 0));
    return l;
}

template <typename T, int NDIM, int SEG, bool SMEM, int BLOCK, bool CERT>
static int backward_launch(const Geom& g, int nsteps, int n_theta, long nP, int broadcast,
                           const void* points, const void* Ws, const void* gout, void* G,
                           void* dpoints, unsigned* counter, const CertArgs& ca, cudaStream_t st,
                           bool& fits, const SampleArgs* sa)
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
        if ((long long)n_theta * ((nP + 255) / 256) > 0x7fffffffLL || nP >= (1L << 32)) { set_error("grid too large"); return kErrUnsupported; }
        unsigned blocks = 0;
        const WorkPlan wp = plan_work(nP, n_theta, BLOCK, per_sm, blocks, true);
        prof_begin(kProfBackward, st);
        kern<<<blocks, BLOCK, smem, st>>>((const T*)points, (const T*)Ws, (const T*)gout, (T*)G,
                                          (T*)dpoints, nP, broadcast, nsteps, g, wp, counter,
                                          (const T*)a.data, (const T*)a.gimg, a.sh, ca);
        prof_end(kProfBackward, st);
        count_launch();
        return kOk;
    };
    const int rc = sa != nullptr ? launch(k_backward<T, NDIM, SEG, SMEM, BLOCK, true, CERT>, *sa)
                                 : launch(k_backward<T, NDIM, SEG, SMEM, BLOCK, false, CERT>, SampleArgs());
    if (rc != kOk) return rc;
    CPAB_CUDA_OK(cudaGetLastError());
    return kOk;
}

// second kernel of the certified mode: the marked trajectories, reference arithmetic
template <int NDIM, int SEG>
static int redo_launch(const Geom& g, int nsteps, int n_theta, long nP, int broadcast, const void* points,
                       const void* Ws, const void* gout, void* G, void* dpoints, unsigned* counter,
                       const CertArgs& ca, cudaStream_t st, bool& fits, const SampleArgs* sa)
{
    const int nseg = (nsteps + SEG - 1) / SEG;
    const size_t per_warp = (((size_t)(31 + 32 * kRedoChunkWords) * 8 + (size_t)nseg * NDIM * 32 * 4 +
                              (size_t)nsteps * 32 * (g.n_cells > 65535 ? 4 : 2)) + 15) & ~(size_t)15;
    int warps = 4;
    while (warps > 1 && per_warp * warps > kMaxSmemBytes) warps /= 2;
    fits = per_warp * warps <= kMaxSmemBytes;
    if (!fits) return kOk;
    const size_t smem = per_warp * warps;
    auto launch = [&](auto kern, const SampleArgs& a) -> int {
        if (smem > 48 * 1024)
            CPAB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * warps, smem);
        const long chunks = ((ca.words + kRedoChunkWords - 1) / kRedoChunkWords) * n_theta;
        long blocks = (long)sm_count() * (per_sm > 0 ? per_sm : 1);
        if (blocks * warps > chunks) blocks = (chunks + warps - 1) / warps;
        prof_begin(kProfBackwardRedo, st);
        kern<<<(unsigned)blocks, 32 * warps, smem, st>>>((const float*)points, (const float*)Ws, (const float*)gout,
                                                         (float*)G, (float*)dpoints, nP, broadcast, nsteps, n_theta, g,
                                                         counter, (const float*)a.data, (const float*)a.gimg, a.sh, ca);
        prof_end(kProfBackwardRedo, st);
        count_launch();
        return kOk;
    };
    const int rc = sa != nullptr ? launch(k_backward_redo<NDIM, SEG, true>, *sa)
                                 : launch(k_backward_redo<NDIM, SEG, false>, SampleArgs());
    if (rc != kOk) return rc;
    CPAB_CUDA_OK(cudaGetLastError());
    return kOk;
}

template <typename T, int NDIM, bool CERT>
static int backward_t(const Geom& g, int nsteps, int n_theta, int d, long nP, int broadcast,
                      const void* points, const void* As, const void* basis, const void* gout,
                      void* dtheta, void* dpoints, void* ws, int* flagged, cudaStream_t st, const SampleArgs* sa)
{
    const int D = g.n_cells * Dim<NDIM>::kPpc;
    const long n_blocks = (long)n_theta * g.n_cells;
    const BwdLayout lay = backward_layout<NDIM>(sizeof(T), g, n_theta, nP);
    char* base = reinterpret_cast<char*>(ws);
    T* Ws = reinterpret_cast<T*>(base + lay.off_w);
    float* stats = CERT ? reinterpret_cast<float*>(base + lay.off_stats) : nullptr;
    unsigned* counter = reinterpret_cast<unsigned*>(base + lay.off_counter);
    if (CERT) CPAB_CUDA_OK(cudaMemsetAsync(stats, 0, (size_t)n_theta * 2 * sizeof(float), st));
    k_prepare_backward<T, NDIM><<<(unsigned)((n_blocks + 255) / 256), 256, 0, st>>>((const T*)As, Ws, (T*)ws, n_blocks, nsteps, g, counter, stats);
    CPAB_CUDA_OK(cudaGetLastError());
    count_launch();
    CertArgs ca{As, stats, cert_scale(g), cert_floor(g), reinterpret_cast<unsigned*>(base + lay.off_mask), lay.words, flagged};
    bool fits = nP == 0;      // nothing to integrate: G stays zero, the epilogue writes dtheta = 0
    int rc = kOk;
#define TRY(SEG, SMEM, BLOCK)                                                                      \
    if (!fits && rc == kOk)                                                                        \
        rc = backward_launch<T, NDIM, SEG, SMEM, BLOCK, CERT>(g, nsteps, n_theta, nP, broadcast, points, \
                                                        Ws, gout, ws, dpoints, counter, ca, st, fits, sa)
    // preferred configuration first, then progressively smaller shared-memory footprints
    // measured (profiles/): 3-D runs best with 3-step segments (96 registers, 5 CTAs/SM) and the
    // per-theta records read through L1 instead of staged (shared memory then holds only the
    // checkpoints and the cell trace); 1-D/2-D with 5-step segments and staged records
    const Tuning& tn = tuning();
    const int seg = tn.bwd_seg != 0 ? tn.bwd_seg : (NDIM == 3 ? 3 : 5);
    const bool stage = tn.bwd_stage >= 0 ? tn.bwd_stage != 0 : NDIM != 3;
    if (seg == 3) {
        if (stage) { if (tn.bwd_block == 256) TRY(3, true, 256); TRY(3, true, 128); }
        else { if (tn.bwd_block == 256) TRY(3, false, 256); TRY(3, false, 128); }
    } else if (seg == 5) {
        if (stage) {
            if (tn.bwd_block == 256) TRY(5, true, 256);
            if (tn.bwd_block == 64) TRY(5, true, 64);
            TRY(5, true, 128);
        } else {
            if (tn.bwd_block == 256) TRY(5, false, 256);
            TRY(5, false, 128);
        }
    } else {
        if (tn.bwd_block == 256) TRY(10, true, 256);
        if (tn.bwd_block == 64) TRY(10, true, 64);
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
    if constexpr (CERT) {
        if (nP > 0) {
            bool rfits = false;
            rc = redo_launch<NDIM, 5>(g, nsteps, n_theta, nP, broadcast, points, Ws, gout, ws, dpoints, counter + 1, ca, st, rfits, sa);
            if (rc == kOk && !rfits)
                rc = redo_launch<NDIM, 10>(g, nsteps, n_theta, nP, broadcast, points, Ws, gout, ws, dpoints, counter + 1, ca, st, rfits, sa);
            if (rc != kOk) return rc;
            if (!rfits) {
                set_error("backward: nstepsolver=%d needs more checkpoint memory than one CTA has", nsteps);
                return kErrUnsupported;
            }
        }
    }
    {
        k_r_to_g<T, NDIM><<<(unsigned)((n_blocks + 255) / 256), 256, 0, st>>>((T*)ws, (const T*)As, n_blocks, nsteps);
        CPAB_CUDA_OK(cudaGetLastError());
        count_launch();
    }
    return launch_grad_epilogue(sizeof(T) == 4 ? kF32 : kF64, ws, basis, dtheta, n_theta, D, d, st);
}

// entry point of one dimension (defined by the including translation unit through this template)
template <int NDIM>
int backward_dim_impl(int dtype, int flags, const Geom& g, int nsteps, int n_theta, int d, long nP,
                      int broadcast, const void* points, const void* As, const void* basis,
                      const void* gout, void* dtheta, void* dpoints, void* ws, int* flagged,
                      cudaStream_t st, const SampleArgs* sa)
{
#define ARGS g, nsteps, n_theta, d, nP, broadcast, points, As, basis, gout, dtheta, dpoints, ws, flagged, st, sa
#ifndef CPAB_FAST_BUILD
    if (dtype == kF64) return backward_t<double, NDIM, false>(ARGS);
#endif
    if (flags & kFlagFastGrad) return backward_t<float, NDIM, false>(ARGS);
    return backward_t<float, NDIM, true>(ARGS);
#undef ARGS
}

template <int NDIM>
int rk2_trace_dim_impl(const Geom& g, int nsteps, int n_theta, long nP, int broadcast, int mode,
                       const void* points, const void* As, void* ws, int* cells, unsigned char* failed,
                       cudaStream_t st)
{
    const long n_blocks = (long)n_theta * g.n_cells;
    const BwdLayout lay = backward_layout<NDIM>(sizeof(float), g, n_theta, nP);
    char* base = reinterpret_cast<char*>(ws);
    float* Ws = reinterpret_cast<float*>(base + lay.off_w);
    float* stats = reinterpret_cast<float*>(base + lay.off_stats);
    unsigned* counter = reinterpret_cast<unsigned*>(base + lay.off_counter);
    CPAB_CUDA_OK(cudaMemsetAsync(stats, 0, (size_t)n_theta * 2 * sizeof(float), st));
    k_prepare_backward<float, NDIM><<<(unsigned)((n_blocks + 255) / 256), 256, 0, st>>>((const float*)As, Ws, (float*)ws, n_blocks, nsteps, g, counter, stats);
    count_launch();
    if (n_theta > 65535) { set_error("rk2_trace: n_theta exceeds 65535"); return kErrUnsupported; }
    dim3 grid((unsigned)((nP + 127) / 128), (unsigned)n_theta);
    k_rk2_trace<NDIM><<<grid, 128, 0, st>>>((const float*)points, (const float*)As, Ws, stats, cells, failed,
                                            nP, n_theta, broadcast, nsteps, mode, g);
    count_launch();
    CPAB_CUDA_OK(cudaGetLastError());
    return kOk;
}

}  // namespace cpab
