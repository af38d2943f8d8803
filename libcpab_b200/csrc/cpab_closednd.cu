// cpab_closednd.cu -- hit-time ("closed-form") integration of 2-D / 3-D CPA velocity fields (sm_100a).
//
// NOT IN THE REFERENCE (SURVEY.md section 0.2, row a11 / f1 of section 8): libcpab integrates with
// nstepsolver fixed steps in every backend.  This is the algorithm BASELINE.json's north_star
// describes -- locate the simplex, follow the analytic in-cell flow x~(t) = expm(t [[L,b],[0,0]]) x~
// to the face it reaches first, compute the hit time, cross into the neighbouring simplex, repeat
// until t = 1 -- as an OPT-IN mode (Cpab.params.closed_form = True).  Checked against the float64
// scipy checker oracle.closed_form_nd, which is anchored to the reference's semantics by
// convergence of its float64 RK2 flow (tests/test_closed_form_oracle.py).
//
// In one simplex of the square / cube (i, j[, k]) everything is done in that cube's local
// coordinates u = x n - idx in [0,1]^n:  du/dt = L' u + b',  L'_rc = L_rc n_r / n_c,
// b'_r = n_r (b_r + sum_c L_rc idx_c / n_c).
//   * In-cell flow: the Taylor polynomial u(t) = sum_k c_k t^k, c_0 = u, c_1 = L' u + b',
//     c_{k+1} = L' c_k / (k+1), on sub-steps tau <= rho / ||L'||_inf (rho = 1/2: the series is then
//     truncated below the rounding of T after K = 8 (float) / 14 (double) terms).  A sub-step that
//     reaches no face just continues in the same simplex; the composition is still the exact flow.
//   * Hit time: every face f (3 per triangle, 4 per tetrahedron; tables below, inward positive) is
//     the polynomial g_f(t) = n_f . u(t) + d_f + eps; the exit is the first time it turns negative (a
//     point within eps of a face -- it has just come in through it, it started on it, or it slides
//     along it -- counts as inside).  Phase 1, the same code for every lane: faces that cannot be
//     reached within the sub-step are excluded by a bound on the polynomial (most are).  Phase 2,
//     per lane over its reachable faces, nearest first: the polynomial at 4 nodes of the sub-step
//     (plus a search for a minimum between two nodes where its derivative changes sign: a
//     trajectory that dips through the face and returns), the first sign change refined by
//     safeguarded Newton.  The hit time has no closed form in more than one dimension (a sum of
//     exponentials); the polynomial is its series.
//   * Crossing: the walk advances to a point 2 eps behind the face (first order; the fields of the
//     two simplices agree on the face) and classifies it in the neighbouring cube's local
//     coordinates by the reference's own inequalities (cpab_ops.cpp:94-103, :160-184).  A
//     trajectory that only grazes the face classifies back into its own simplex and goes on.  Outer
//     faces of the domain are never crossed: outside the unit box (tessellations without zero
//     boundary) the boundary cubes' planes are simply continued.
// The loop runs once per sub-step; its trip count varies per trajectory (crossings + ||L|| / rho).
// Lanes that finish refill themselves with the next point of their warp's range ("closed_refill"
// tuning key, default on), so a warp keeps 32 trajectories in flight instead of waiting for its
// slowest lane; cpab_b200_closed_form_lane_stats measures the lane utilisation with and without.
//
// Gradient.  CPA fields are continuous across faces, so the sensitivity dx/dtheta has no jump at a
// crossing (the jump would be (v- - v+) dt*/dtheta = 0): the exact gradient is the in-cell
// variational equation integrated piecewise, in adjoint form
//     dL/dtheta = sum_c <B_c, G_c>,   G_c = int_{t: x(t) in c} lambda(t) [x(t); 1]^T dt,   lambda' = -L_c^T lambda.
// The backward kernel takes x(1) from the caller (autograd has it; a kernel that walks forward to it first
// has lanes in both phases at once and pays for both code paths: 2.4x slower) and walks the REVERSED field
// (-A) back from there -- the same hit-time code, the same cells in reverse order up to rounding, no trajectory
// storage -- carrying lambda'(r) = expm(r L'^T) lambda' as a second Taylor polynomial and adding
// int lambda u^T dr per sub-step by Gauss-Legendre quadrature (3 / 8 nodes).  G goes through the same
// G.B epilogue as the fixed-step adjoint; lambda at t = 0 is dL/dpoints.
#include "cpab_device.cuh"

namespace cpab {

namespace {

// faces of the simplices in local coordinates, rows (normal..., offset), inward positive.
// 2-D: triangle types of cpab_ops.cpp:94-103.  3-D: tetrahedra of :160-184 in the coordinates of an
// even cube; cubes of odd i+j+k use (x, y) <- (y, 1 - x) (:170-174).
__constant__ float c_faces2[4][3][3] = {
    {{1, -1, 0}, {-1, -1, 1}, {0, 1, 0}},
    {{1, -1, 0}, {1, 1, -1}, {-1, 0, 1}},
    {{-1, 1, 0}, {1, 1, -1}, {0, -1, 1}},
    {{-1, 1, 0}, {-1, -1, 1}, {1, 0, 0}},
};
__constant__ float c_faces3[5][4][4] = {
    {{1, 1, -1, 0}, {-1, -1, -1, 2}, {1, -1, 1, 0}, {-1, 1, 1, 0}},
    {{-1, -1, 1, 0}, {1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, -1, 1}},
    {{1, 1, 1, -2}, {-1, 0, 0, 1}, {0, -1, 0, 1}, {0, 0, -1, 1}},
    {{-1, 1, -1, 0}, {1, 0, 0, 0}, {0, -1, 0, 1}, {0, 0, 1, 0}},
    {{1, -1, -1, 0}, {-1, 0, 0, 1}, {0, 1, 0, 0}, {0, 0, 1, 0}},
};

template <typename T> struct Cf;
template <> struct Cf<float> {
    static constexpr int K = 8, Q = 3, kNewton = 4, kScanNodes = 4;     // (2 nodes: 1e-4 errors on 16 x 256^2 trajectories, measured)
    static constexpr float kEps = 1e-6f;        // x max(nc): |g| below this is "on the face"
    // 3-point Gauss-Legendre on [0, 1]: error 5e-7 (2 rho)^6 of a sub-step's contribution, below float's own rounding
    static __device__ __forceinline__ void gl(int q, float& x, float& w)
    {
        const float xs[3] = {0.112701665379258f, 0.5f, 0.887298334620742f};
        const float ws[3] = {0.277777777777778f, 0.444444444444444f, 0.277777777777778f};
        x = xs[q]; w = ws[q];
    }
};
template <> struct Cf<double> {
    static constexpr int K = 14, Q = 8, kNewton = 6, kScanNodes = 4;
    static constexpr double kEps = 1e-13;
    static __device__ __forceinline__ void gl(int q, double& x, double& w)
    {
        const double xs[8] = {0.019855071751231884, 0.10166676129318664, 0.2372337950418355, 0.4082826787521751,
                              0.5917173212478249, 0.7627662049581645, 0.8983332387068134, 0.9801449282487681};
        const double ws[8] = {0.05061426814518813, 0.11119051722668724, 0.15685332293894364, 0.18134189168918100,
                              0.18134189168918100, 0.15685332293894364, 0.11119051722668724, 0.05061426814518813};
        x = xs[q]; w = ws[q];
    }
};

constexpr int kRing = 128;              // points a warp keeps prefetched in shared memory
constexpr int kMaxSubsteps = 1 << 14;     // guard: after this many sub-steps the faces are no longer tested

template <typename T, int NDIM> struct Walker {
    T x[NDIM];          // current point, global coordinates
    T trem;             // time left
    int idx[NDIM];      // square / cube
    int typ;            // simplex within it
    int steps;
    unsigned closed;    // faces not tested (found outside one at time 0 but classified into this simplex), until time advances
    // adjoint (reverse walk only)
    T lam[NDIM];                    // dL/dx, global
    T iu[NDIM][NDIM], i1[NDIM];     // int lambda' u^T dr, int lambda' dr over the stay in the current simplex
};

// cell width 1 / n (the division costs a dozen instructions a time)
__device__ __forceinline__ float width_of(const Geom& g, int j, float) { return g.w[j]; }
__device__ __forceinline__ double width_of(const Geom& g, int j, double) { return g.wd[j]; }

template <int NDIM> __device__ __forceinline__ int parity_of(const int* idx)
{
    return NDIM == 3 ? ((idx[0] + idx[1] + idx[2]) & 1) : 0;
}

// simplex of a local point; the inequalities of cpab_ops.cpp:94-103 / :160-184 (continued outside the cube)
template <typename T, int NDIM> __device__ __forceinline__ int simplex_type(const T* u, int parity)
{
    if (NDIM == 2) {
        const T x = u[0], y = u[1];
        if (x < y) return ((T)1 - x < y) ? 2 : 3;
        return ((T)1 - x < y) ? 1 : 0;
    }
    T x = u[0], y = u[1];
    const T z = u[NDIM - 1];
    if (parity) { const T t = x; x = y; y = (T)1 - t; }
    if (-x - y + z >= (T)0) return 1;
    if (x + y + z - (T)2 >= (T)0) return 2;
    if (-x + y - z >= (T)0) return 3;
    if (x - y - z >= (T)0) return 4;
    return 0;
}

// face f of simplex `typ`: normal and offset in the cube's own (unrotated) local coordinates
// (`tbl`: the kernel's shared-memory copy of the table -- lanes sit in different simplices, and a constant-bank
//  load with 32 different addresses is replayed per address)
template <typename T, int NDIM> __device__ __forceinline__ void face_of(const float* tbl, int typ, int f, int parity, T* n, T& d)
{
    if (NDIM == 2) {
        const float* r = tbl + (typ * 3 + f) * 3;
        n[0] = (T)r[0]; n[1] = (T)r[1]; d = (T)r[2];
    } else {
        const float* r = tbl + (typ * 4 + f) * 4;
        const T a = (T)r[0], b = (T)r[1], c = (T)r[2];
        d = (T)r[3];
        if (parity) { n[0] = -b; n[1] = a; d += b; }     // a x' + b y' + c z + d,  x' = y, y' = 1 - x
        else { n[0] = a; n[1] = b; }
        n[NDIM - 1] = c;
    }
}

template <typename T, int K> __device__ __forceinline__ T horner(const T* a, T t)
{
    T r = a[K];
#pragma unroll
    for (int k = K - 1; k >= 0; --k) r = Num<T>::fma(r, t, a[k]);
    return r;
}
template <typename T, int K> __device__ __forceinline__ void horner2(const T* a, T t, T& f, T& df)
{
    f = a[K]; df = (T)0;
#pragma unroll
    for (int k = K - 1; k >= 0; --k) { df = Num<T>::fma(df, t, f); f = Num<T>::fma(f, t, a[k]); }
}

template <typename T, int NDIM> struct CfTable {      // velocity matrices of one theta: shared memory or global
    const T* A;
    const float* faces;     // shared-memory copy of c_faces2 / c_faces3
    __device__ __forceinline__ void load(int c, T* a) const
    {
#pragma unroll
        for (int e = 0; e < Dim<NDIM>::kPpc; ++e) a[e] = A[(size_t)c * Dim<NDIM>::kPpc + e];
    }
};

template <typename T, int NDIM>
__device__ __forceinline__ int cell_of(const Geom& g, const int* idx, int typ)
{
    constexpr int SPC = NDIM == 2 ? 4 : 5;
    int s = idx[0] + idx[1] * g.nc[0];
    if (NDIM == 3) s += idx[2] * g.nc[0] * g.nc[1];
    return SPC * s + typ;
}

template <typename T, int NDIM>
__device__ __forceinline__ void flush_cell(const Geom& g, Walker<T, NDIM>& w, T* Gt, int c)
{
    constexpr int M = NDIM + 1;
    T v[NDIM * M];
#pragma unroll
    for (int r = 0; r < NDIM; ++r) {
        const T nr = (T)g.nc[r];
#pragma unroll
        for (int cc = 0; cc < NDIM; ++cc)
            v[r * M + cc] = nr * (w.iu[r][cc] + (T)w.idx[cc] * w.i1[r]) * width_of(g, cc, (T)0);
        v[r * M + NDIM] = nr * w.i1[r];
    }
    red_cell<NDIM * M>(Gt + (size_t)c * (NDIM * M), v);
#pragma unroll
    for (int r = 0; r < NDIM; ++r) {
        w.i1[r] = (T)0;
#pragma unroll
        for (int cc = 0; cc < NDIM; ++cc) w.iu[r][cc] = (T)0;
    }
}

// One sub-step of the walk along sgn * field.  Returns true when the unit time is used up.
// ADJ (reverse walk, sgn = -1): also carries lambda and the per-simplex integrals, and hands a
// simplex's integrals to G when the trajectory leaves it.
template <typename T, int NDIM, bool ADJ>
__device__ __noinline__ bool substep(const Geom& g, const CfTable<T, NDIM>& tab, Walker<T, NDIM>& w, T sgn, T eps, T* Gt)
{
    constexpr int K = Cf<T>::K;
    constexpr int M = NDIM + 1;
    constexpr int NF = NDIM + 1;
    const int c = cell_of<T, NDIM>(g, w.idx, w.typ);
    T A[NDIM * M];
    tab.load(c, A);
    // local field (signed)
    T Lp[NDIM][NDIM], bp[NDIM], u[NDIM];
    T norm = (T)0;
#pragma unroll
    for (int r = 0; r < NDIM; ++r) {
        const T nr = sgn * (T)g.nc[r];
        T acc = A[r * M + NDIM], row = (T)0;
#pragma unroll
        for (int cc = 0; cc < NDIM; ++cc) {
            const T wc = width_of(g, cc, (T)0);
            acc = Num<T>::fma(A[r * M + cc], (T)w.idx[cc] * wc, acc);
            Lp[r][cc] = nr * A[r * M + cc] * wc;
            row += fabs(Lp[r][cc]);
        }
        bp[r] = nr * acc;
        norm = fmax(norm, row);
        u[r] = Num<T>::fma(w.x[r], (T)g.nc[r], -(T)w.idx[r]);
    }
    T tau = w.trem;
    if (norm * tau > (T)0.5 && norm < (T)1e30) tau = (T)0.5 / norm;      // (a non-finite field takes one step and ends)
    if (w.steps > 4 * kMaxSubsteps) w.trem = tau;                          // hard bound on the trip count
    // Taylor coefficients of u(t)
    T ck[K + 1][NDIM];
#pragma unroll
    for (int j = 0; j < NDIM; ++j) ck[0][j] = u[j];
#pragma unroll
    for (int r = 0; r < NDIM; ++r) {
        T acc = bp[r];
#pragma unroll
        for (int cc = 0; cc < NDIM; ++cc) acc = Num<T>::fma(Lp[r][cc], u[cc], acc);
        ck[1][r] = acc;
    }
#pragma unroll
    for (int k = 1; k < K; ++k) {
        const T inv = (T)1 / (T)(k + 1);
#pragma unroll
        for (int r = 0; r < NDIM; ++r) {
            T acc = (T)0;
#pragma unroll
            for (int cc = 0; cc < NDIM; ++cc) acc = Num<T>::fma(Lp[r][cc], ck[k][cc], acc);
            ck[k + 1][r] = acc * inv;
        }
    }
    // first face reached within (0, tau]
    const int parity = parity_of<NDIM>(w.idx);
    T best = tau;
    T probe = tau;          // a time just after the crossing: where the trajectory is classified
    int hit = -1;
    if (w.steps < kMaxSubsteps) {
        // Phase 1, the same code for every lane: which faces can be reached at all within the sub-step
        // (g(t) >= a0 + min(0, a1 t) - sum_{k>=2} |a_k| t^k on [0, tau]).  Most cannot; a lane is usually left
        // with none or one.  An outer face of the domain is never crossed.
        unsigned pend = 0u;
        T est[NF];      // linear estimate of the hit time: the order in which phase 2 takes the faces
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            T n[NDIM], d, a[K + 1];
            face_of<T, NDIM>(tab.faces, w.typ, f, parity, n, d);
            int axis = -1, nz = 0;
#pragma unroll
            for (int j = 0; j < NDIM; ++j) if (n[j] != (T)0) { axis = j; ++nz; }
            const bool outer = nz == 1 && ((n[axis] > (T)0 && w.idx[axis] == 0) || (n[axis] < (T)0 && w.idx[axis] == g.nc[axis] - 1));
#pragma unroll
            for (int k = 0; k <= K; ++k) {
                T acc = k == 0 ? d + eps : (T)0;
#pragma unroll
                for (int j = 0; j < NDIM; ++j) acc = Num<T>::fma(n[j], ck[k][j], acc);
                a[k] = acc;
            }
            T r = fabs(a[K]);
#pragma unroll
            for (int k = K - 1; k >= 2; --k) r = Num<T>::fma(r, tau, fabs(a[k]));
            const bool reach = !(a[0] + fmin((T)0, a[1] * tau) - r * tau * tau > (T)0);
            if (reach && !outer && !((w.closed >> f) & 1u)) pend |= 1u << f;
            est[f] = a[1] < (T)0 ? a[0] / -a[1] : tau;
        }
        // Phase 2: every lane scans ITS reachable faces -- different lanes different faces at the same time, so the
        // warp iterates as often as its busiest lane has faces (1-2), not once per face of the simplex.
#pragma unroll 1
        while (pend != 0u) {
            // nearest first: once its exit is known, the bound on (0, best] usually disposes of the others
            int f = 0;
            T e = (T)INFINITY;
#pragma unroll
            for (int k = 0; k < NF; ++k)
                if (((pend >> k) & 1u) && est[k] < e) { e = est[k]; f = k; }
            pend &= ~(1u << f);
            T n[NDIM], d;
            face_of<T, NDIM>(tab.faces, w.typ, f, parity, n, d);
            T a[K + 1];
#pragma unroll
            for (int k = 0; k <= K; ++k) {
                T acc = k == 0 ? d : (T)0;
#pragma unroll
                for (int j = 0; j < NDIM; ++j) acc = Num<T>::fma(n[j], ck[k][j], acc);
                a[k] = acc;
            }
            // exit = the first time g(t) = f(t) + eps turns negative: a point on the face (|f| <= eps: it has
            // just come in through it, started on it, or slides along it) is still inside
            a[0] += eps;
            const T lim = best;
            {   // the same exclusion on (0, best]: an earlier face may have shortened the interval
                T r = fabs(a[K]);
#pragma unroll
                for (int k = K - 1; k >= 2; --k) r = Num<T>::fma(r, lim, fabs(a[k]));
                if (a[0] + fmin((T)0, a[1] * lim) - r * lim * lim > (T)0) continue;
            }
            T tprev = (T)0, fprev = a[0], dprev = a[1];
            bool found = a[0] < (T)0;       // already outside (a crossing classified into the wrong simplex): leave at once
            T cand = (T)0, after = (T)0;
#pragma unroll 1
            for (int m = 1; m <= Cf<T>::kScanNodes && !found; ++m) {
                T t = m == Cf<T>::kScanNodes ? lim : lim * ((T)m / (T)Cf<T>::kScanNodes);
                T fm, dm;
                horner2<T, K>(a, t, fm, dm);
                if (fm >= (T)0 && dprev < (T)0 && dm > (T)0) {
                    // a minimum inside (tprev, t): the trajectory may dip through the face and come back
                    T lo = tprev, hi = t;
#pragma unroll 1
                    for (int it = 0; it < 2 * Cf<T>::kNewton; ++it) {
                        const T mid = (T)0.5 * (lo + hi);
                        T fv, dv;
                        horner2<T, K>(a, mid, fv, dv);
                        if (dv < (T)0) lo = mid; else hi = mid;
                    }
                    T fv, dv;
                    horner2<T, K>(a, lo, fv, dv);
                    if (fv < (T)0) { t = lo; fm = fv; dm = dv; }      // it does: the exit is bracketed by [tprev, lo]
                }
                if (fm < (T)0) {     // bracket [tprev, t]: safeguarded Newton on the polynomial
                    found = true;
                    T lo = tprev, hi = t;
                    T tt = lo + (hi - lo) * fprev / (fprev - fm);
                    T dv = dm;
#pragma unroll 1
                    for (int it = 0; it < Cf<T>::kNewton; ++it) {
                        T fv;
                        horner2<T, K>(a, tt, fv, dv);
                        if (fv > (T)0) lo = tt; else hi = tt;
                        T tn = tt - fv / dv;
                        if (!(tn >= lo && tn <= hi)) tn = (T)0.5 * (lo + hi);
                        const bool conv = fabs(tn - tt) <= (T)(sizeof(T) == 4 ? 2e-7 : 1e-15) * hi;
                        tt = tn;
                        if (conv) break;           // (also: a strict interval test would bisect away from a root it sits on)
                    }
                    cand = tt;
                    // go one more eps out (first order; 2 eps behind the face in all: ten times the rounding of the
                    // classification, and as little of the next simplex's time as possible spent with this one's field),
                    // but not later than the node that saw it outside
                    const T dt = eps / fmax(fabs(dv), (T)1e-30);
                    after = cand + dt < t ? cand + dt : t;
                }
                tprev = t; fprev = fm; dprev = dm;
            }
            if (found && (cand < best || hit < 0)) {
                best = fmin(cand, best);
                probe = after;
                hit = f;
            }
        }
    }
    // advance -- on a crossing to the probe time, a few eps behind the face: the point is then strictly inside the
    // simplex it is classified into below (also when that is this one again: a trajectory that grazes a face
    // and comes back just goes on, and a later, real crossing of the same face is seen by the next sub-step)
    const bool at_once = hit >= 0 && !(probe > (T)0);
    if (hit >= 0) best = probe;
    T un[NDIM];
#pragma unroll
    for (int j = 0; j < NDIM; ++j) {
        T col[K + 1];
#pragma unroll
        for (int k = 0; k <= K; ++k) col[k] = ck[k][j];
        un[j] = horner<T, K>(col, best);
    }
    if (ADJ) {
        // lambda'(r) = expm(r L'^T) lambda' along the reversed walk: l_{j+1} = -(sgn L')^T l_j / (j+1)
        T lk[K + 1][NDIM];
#pragma unroll
        for (int j = 0; j < NDIM; ++j) lk[0][j] = w.lam[j] * width_of(g, j, (T)0);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const T inv = -(T)1 / (T)(k + 1);
#pragma unroll
            for (int cc = 0; cc < NDIM; ++cc) {
                T acc = (T)0;
#pragma unroll
                for (int r = 0; r < NDIM; ++r) acc = Num<T>::fma(Lp[r][cc], lk[k][r], acc);
                lk[k + 1][cc] = acc * inv;
            }
        }
#pragma unroll 1
        for (int q = 0; q < Cf<T>::Q; ++q) {
            T xi, wq;
            Cf<T>::gl(q, xi, wq);
            const T r = best * xi;
            wq *= best;
            T uq[NDIM], lq[NDIM];
#pragma unroll
            for (int j = 0; j < NDIM; ++j) {
                T cu[K + 1], cl[K + 1];
#pragma unroll
                for (int k = 0; k <= K; ++k) { cu[k] = ck[k][j]; cl[k] = lk[k][j]; }
                uq[j] = horner<T, K>(cu, r);
                lq[j] = horner<T, K>(cl, r) * wq;
            }
#pragma unroll
            for (int rr = 0; rr < NDIM; ++rr) {
                w.i1[rr] += lq[rr];
#pragma unroll
                for (int cc = 0; cc < NDIM; ++cc) w.iu[rr][cc] = Num<T>::fma(lq[rr], uq[cc], w.iu[rr][cc]);
            }
        }
#pragma unroll
        for (int j = 0; j < NDIM; ++j) {
            T cl[K + 1];
#pragma unroll
            for (int k = 0; k <= K; ++k) cl[k] = lk[k][j];
            w.lam[j] = horner<T, K>(cl, best) * (T)g.nc[j];
        }
    }
#pragma unroll
    for (int j = 0; j < NDIM; ++j) w.x[j] = ((T)w.idx[j] + un[j]) * width_of(g, j, (T)0);
    w.trem -= best;
    w.steps += 1;
    const bool done = !(w.trem > (T)0);
    if (best > (T)0) w.closed = 0u;
    if (hit >= 0) {
        // the simplex the trajectory is in now: the reference's inequalities in the (neighbouring) cube
        int idx2[NDIM];
        T v[NDIM];
        bool moved = false;
#pragma unroll
        for (int j = 0; j < NDIM; ++j) {
            v[j] = un[j];
            idx2[j] = w.idx[j];
            if (v[j] < (T)0 && idx2[j] > 0) { idx2[j] -= 1; v[j] += (T)1; moved = true; }
            else if (v[j] > (T)1 && idx2[j] < g.nc[j] - 1) { idx2[j] += 1; v[j] -= (T)1; moved = true; }
        }
        const int typ2 = simplex_type<T, NDIM>(v, parity_of<NDIM>(idx2));
        if (moved || typ2 != w.typ) {
            if (ADJ) flush_cell<T, NDIM>(g, w, Gt, c);
#pragma unroll
            for (int j = 0; j < NDIM; ++j) w.idx[j] = idx2[j];
            w.typ = typ2;
            w.closed = 0u;
        } else if (at_once) {
            // found outside a face at the start of the sub-step, yet classified into this simplex (rounding): no
            // time has passed, so go on without this face until it has
            w.closed |= 1u << hit;
        }
    }
    if (ADJ && done) flush_cell<T, NDIM>(g, w, Gt, cell_of<T, NDIM>(g, w.idx, w.typ));
    return done;
}

// Start simplex by the walk's own geometry (floor of the local coordinate, clamped to the tessellation,
// then the inequalities of simplex_type): inside the box this is findcellidx up to ties on faces -- which
// do not matter, the fields agree there -- and without the float32 quirks of the reference's search
// (cpab_ops.cpp:139-148: a coordinate of exactly 1 lands in the wrong simplex; the walk would leave it
// at once anyway); outside the box it is the continuation the walk itself uses.
template <typename T, int NDIM>
__device__ __forceinline__ void start_walk(const Geom& g, Walker<T, NDIM>& w)
{
    T u[NDIM];
#pragma unroll
    for (int j = 0; j < NDIM; ++j) {
        const T s = w.x[j] * (T)g.nc[j];
        int k = (int)floor(s);
        k = k < 0 ? 0 : (k > g.nc[j] - 1 ? g.nc[j] - 1 : k);
        w.idx[j] = k;
        u[j] = s - (T)k;
    }
    w.typ = simplex_type<T, NDIM>(u, parity_of<NDIM>(w.idx));
    w.trem = (T)1;
    w.steps = 0;
    w.closed = 0u;
}

// One CTA = one theta x one chunk of points; every warp walks a contiguous sub-range with lane refill.
// stats (optional, device): [0] += sub-steps executed by lanes, [1] += 32 x loop iterations of warps.
template <typename T, int NDIM, bool BACKWARD>
__global__ void __launch_bounds__(128)
k_closednd(const T* __restrict__ points, const T* __restrict__ As, const T* __restrict__ gout,
           T* __restrict__ out, T* __restrict__ G, T* __restrict__ dpoints, long nP, int broadcast,
           int chunks, int chunk_pts, int staged, int refill, unsigned long long* __restrict__ stats,
           const __grid_constant__ Geom g)
{
    constexpr int PPC = Dim<NDIM>::kPpc;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int theta = blockIdx.x / chunks;
    const int chunk = blockIdx.x - theta * chunks;
    const T* Ag = As + (size_t)theta * g.n_cells * PPC;
    CfTable<T, NDIM> tab;
    __shared__ float s_faces[NDIM == 2 ? 36 : 80];
    for (int i = threadIdx.x; i < (NDIM == 2 ? 36 : 80); i += blockDim.x)
        s_faces[i] = NDIM == 2 ? (&c_faces2[0][0][0])[i] : (&c_faces3[0][0][0])[i];
    tab.faces = s_faces;
    if (!staged) __syncthreads();
    constexpr int NV = BACKWARD ? 2 * NDIM : NDIM;
    if (staged) {
        T* sA = reinterpret_cast<T*>(smem_raw) + (size_t)4 * kRing * NV;
        for (int i = threadIdx.x; i < g.n_cells * PPC; i += blockDim.x) sA[i] = Ag[i];
        __syncthreads();
        tab.A = sA;
    } else {
        tab.A = Ag;
    }
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int per_warp = chunk_pts / nwarps;
    const long cbegin = (long)chunk * chunk_pts;
    long next = cbegin + (long)warp * per_warp;
    long wend = next + per_warp;
    if (wend > nP) wend = nP;
    const T* src = points + (broadcast ? (size_t)theta * NDIM * nP : 0);
    T* Gt = BACKWARD ? G + (size_t)theta * g.n_cells * PPC : nullptr;
    int nmax = g.nc[0] > g.nc[1] ? g.nc[0] : g.nc[1];
    if (NDIM == 3 && g.nc[2] > nmax) nmax = g.nc[2];
    const T eps = Cf<T>::kEps * (T)nmax;

    // ---- prefetch ring of this warp: slot (i & 127) of value v at ringw[v * kRing + slot]
    T* ringw = reinterpret_cast<T*>(smem_raw) + (size_t)warp * kRing * NV;
    const bool from_x1 = BACKWARD && out != nullptr;     // backward with the forward's output at hand (`out`, an input then)
    long loaded = next;                                   // points [next, loaded) are in the ring (or on their way)
    auto issue_group = [&]() {
        const long pi = loaded + lane;
        if (pi < wend) {
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const T* gp = v < NDIM ? (from_x1 ? out + (size_t)theta * NDIM * nP + pi + (long)v * nP : src + pi + (long)v * nP)
                                       : gout + (size_t)theta * NDIM * nP + pi + (long)(v - NDIM) * nP;
                const uint32_t sp = (uint32_t)__cvta_generic_to_shared(ringw + v * kRing + (int)(pi & (kRing - 1)));
                if (sizeof(T) == 4) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(sp), "l"(gp) : "memory");
                else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(sp), "l"(gp) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        loaded += 32;
    };
    issue_group(); issue_group(); issue_group(); issue_group();      // (beyond wend: empty groups)

    Walker<T, NDIM> w;
    T lam0[NDIM];           // upstream gradient of the lane's trajectory (backward), taken from the ring at the refill
    bool active = false, reverse = false;
    long i = 0;
    auto begin_reverse = [&]() {
        w.trem = (T)1;
        w.steps = 0;
        w.closed = 0u;
#pragma unroll
        for (int r = 0; r < NDIM; ++r) {
            w.lam[r] = BACKWARD ? lam0[r] : (T)0;
            w.i1[r] = (T)0;
#pragma unroll
            for (int cc = 0; cc < NDIM; ++cc) w.iu[r][cc] = (T)0;
        }
    };
    unsigned long long lane_steps = 0, warp_iters = 0;
    for (;;) {
        // keep the ring at least 64 points ahead (one group per iteration is as fast as the lanes can consume):
        // what a refill reads then always belongs to a group older than the newest one
        bool issued = false;
        if (loaded < wend && loaded - next <= kRing - 32) { issue_group(); issued = true; }
        if (refill || !__any_sync(full, active)) {
            const unsigned m = __ballot_sync(full, !active);
            const long ni = next + __popc(m & ((1u << lane) - 1u));
            if (m != 0u) {
                if (issued) asm volatile("cp.async.wait_group 1;" ::: "memory");
                else asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncwarp();          // the copies of every lane are visible to every lane
            }
            if (!active && ni < wend) {
                i = ni;
                const int slot = (int)(i & (kRing - 1));
#pragma unroll
                for (int j = 0; j < NDIM; ++j) {
                    w.x[j] = ringw[j * kRing + slot];
                    if (BACKWARD) lam0[j] = ringw[(NDIM + j) * kRing + slot];
                }
                start_walk<T, NDIM>(g, w);
                reverse = from_x1;
                if (from_x1) begin_reverse();
                active = true;
            }
            next += __popc(m);
        }
        if (!__any_sync(full, active)) break;
        ++warp_iters;
        if (active) {
            ++lane_steps;
            bool fin;
            if (BACKWARD && reverse) fin = substep<T, NDIM, true>(g, tab, w, (T)-1, eps, Gt);
            else fin = substep<T, NDIM, false>(g, tab, w, (T)1, eps, nullptr);
            if (fin) {
                if (BACKWARD && !reverse) {      // x(1) reached: walk the reversed field back with the adjoint
                    reverse = true;
                    begin_reverse();
                } else {
                    if (!BACKWARD) {
#pragma unroll
                        for (int j = 0; j < NDIM; ++j) out[(size_t)theta * NDIM * nP + i + (long)j * nP] = w.x[j];
                    } else if (dpoints != nullptr) {
#pragma unroll
                        for (int j = 0; j < NDIM; ++j) dpoints[(size_t)theta * NDIM * nP + i + (long)j * nP] = w.lam[j];
                    }
                    active = false;
                }
            }
        }
    }
    if (stats != nullptr) {
        atomicAdd(stats, lane_steps);
        if (lane == 0) atomicAdd(stats + 1, 32ull * warp_iters);
    }
}

int& closed_refill_flag()
{
    thread_local int v = 1;
    return v;
}
int& closed_stage_flag()
{
    thread_local int v = 1;
    return v;
}


template <typename T, int NDIM, bool BACKWARD>
int closednd_launch(const Geom& g, int n_theta, long nP, int broadcast, const void* points, const void* As,
                    const void* gout, void* out, void* G, void* dpoints, int refill, unsigned long long* stats,
                    cudaStream_t st)
{
    const size_t table = (size_t)g.n_cells * Dim<NDIM>::kPpc * sizeof(T);
    const int staged = table <= 96 * 1024 && closed_stage_flag() != 0;      // else: the matrices through L1 (large tessellations)
    // per warp: a ring of the next 128 points' inputs (start point; backward: the upstream gradient too), filled by
    // cp.async three groups ahead of the lane refills that consume it -- a refill that loads from global memory
    // stalls its whole warp on every loop iteration in which some lane finishes, i.e. on nearly every one
    const size_t ring = (size_t)4 * kRing * (BACKWARD ? 2 * NDIM : NDIM) * sizeof(T);
    const size_t smem = ring + (staged ? table : 0);
    // points per CTA (a quarter of it per warp): the longer a warp's range, the shorter the share of its refill loop
    // spent with idle lanes at the end (lane utilisation 0.91 / 0.96 / 0.98 at 1024 / 2048 / 4096).  Measured, 2-D / 3-D:
    // forward 1.67 / 6.53 ms at 1024, 1.61 / 6.61 at 2048, 1.69 / 6.75 at 4096; backward 2.39 / 13.2 at 1024,
    // 2.20 / 9.9 at 4096, 2.51 / 10.3 at 8192
    int chunk_pts = BACKWARD ? 4096 : 2048;
    // few thetas: cut finer so that the grid fills the chip
    while (chunk_pts > 128 && (long long)n_theta * ((nP + chunk_pts - 1) / chunk_pts) < 4LL * sm_count()) chunk_pts /= 2;
    const long chunks = (nP + chunk_pts - 1) / chunk_pts;
    const long long blocks = (long long)n_theta * chunks;
    if (blocks > 0x7fffffffLL) { set_error("grid too large"); return kErrUnsupported; }
    auto kern = k_closednd<T, NDIM, BACKWARD>;
    if (smem > 48 * 1024) CPAB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int slot = BACKWARD ? kProfBackward : kProfForward;
    prof_begin(slot, st);
    kern<<<(unsigned)blocks, 128, smem, st>>>((const T*)points, (const T*)As, (const T*)gout, (T*)out, (T*)G, (T*)dpoints,
                                              nP, broadcast, (int)chunks, chunk_pts, staged, refill, stats, g);
    prof_end(slot, st);
    count_launch();
    CPAB_CUDA_OK(cudaGetLastError());
    return kOk;
}

}  // namespace

void set_closed_refill(int v) { closed_refill_flag() = v; }
void set_closed_stage(int v) { closed_stage_flag() = v; }

int launch_closednd_forward(int dtype, const Geom& g, int n_theta, long nP, int broadcast, const void* points,
                            const void* As, void* out, unsigned long long* stats, cudaStream_t st)
{
    if (n_theta == 0 || nP == 0) return kOk;
    const int rf = closed_refill_flag();
#define GO(T, N) closednd_launch<T, N, false>(g, n_theta, nP, broadcast, points, As, nullptr, out, nullptr, nullptr, rf, stats, st)
    if (g.ndim == 2) return dtype == kF32 ? GO(float, 2) : GO(double, 2);
    return dtype == kF32 ? GO(float, 3) : GO(double, 3);
#undef GO
}

// G [n_theta, D] must be zero-initialised by the caller
// newpoints: the forward's output if the caller has it (the reverse walk then starts from it), else NULL
int launch_closednd_backward(int dtype, const Geom& g, int n_theta, long nP, int broadcast, const void* points,
                             const void* As, const void* gout, const void* newpoints, void* G, void* dpoints,
                             cudaStream_t st)
{
    if (n_theta == 0 || nP == 0) return kOk;
    const int rf = closed_refill_flag();
#define GO(T, N) closednd_launch<T, N, true>(g, n_theta, nP, broadcast, points, As, gout, const_cast<void*>(newpoints), G, dpoints, rf, nullptr, st)
    if (g.ndim == 2) return dtype == kF32 ? GO(float, 2) : GO(double, 2);
    return dtype == kF32 ? GO(float, 3) : GO(double, 3);
#undef GO
}

}  // namespace cpab
