// cpab_cell.cuh -- point -> simplex index ("findcellidx") for 1-D / 2-D / 3-D CPAB tessellations.
//
// Replaces libcpab/core/cpab_ops.cpp:26-190 (CPU) and libcpab/core/cpab_ops.cu:14-227 (GPU) of the
// reference.  The contract is BIT-EXACT agreement with the reference's CPU function on every float
// input, but the arithmetic is redesigned for the GPU:
//
//   * The reference works in double with 6 (2-D) / 9 (3-D) fp64 divisions and an fmod per axis.
//     fmod is exact, so the reference's column index is floor_exact(p / w) and its remainder is
//     p - k*w, both of which a single FP32 FMA delivers exactly: estimate k with one FFMA against
//     a rounding constant, form r = fma(-k, w, p) (exact: r is a multiple of ulp(w) below 2^24
//     ulps), fix the estimate by the sign of r.  No division, no conversion instruction.  (2-D uses
//     a one-sided estimate -- FFMA rounded down with a multiplier >= 1/w -- whose rare miss joins
//     the guard-band branch, and runs both axes as packed FP32; see divmod_up.)
//   * The triangle / tetrahedron tests compare local coordinates r/w.  They are evaluated in FP32
//     with a guard band; only a point within the band of a diagonal (or in a corner region outside
//     the domain) drops to an exact path.  In 2-D the exact path is still division-free (products
//     of two floats are exact in double); the last resort replays the reference's own double
//     expression sequence.
//   * All per-tessellation constants the reference recomputes per call (cell widths rounded to
//     float, the `n*inc - 1e-9` clamp, the row/column a clamped coordinate falls in) are evaluated
//     once on the host by make_geom(), literally as the reference writes them.
//
// The header compiles for host and device.  The host build is used ONLY by the CPU test harness
// (tests/harness), which sweeps these functions against the oracle; the product has no CPU path.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#include "cpab_f32x2.cuh"

#if defined(__CUDACC__)
#define CPAB_HD __host__ __device__ __forceinline__
#define CPAB_HD_NOINLINE inline __host__ __device__ __noinline__
#else
#define CPAB_HD inline
#define CPAB_HD_NOINLINE inline
#endif

namespace cpab {

// Tessellation geometry plus host-evaluated constants.  Passed to kernels by value.
struct Geom {
    int ndim;
    int nc[3];
    int n_cells;        // simplices: nx | 4 nx ny | 5 nx ny nz
    // ---- float32 path -----------------------------------------------------------------------
    float nf[3];        // (float)n
    float nm1[3];       // (float)(n-1)
    float w[3];         // cell width rounded to float: `const float inc = 1.0 / n`
    float span[3];      // (float)(n * inc): the upper domain bound the reference compares with
    float khi[3];       // 2-D: column/row (already min'ed with n-1) of a coordinate clamped high
    float spanm[3];     // 2-D: largest float below span
    float nup[3];       // 2-D: smallest float >= 1/w (multiplier of the one-sided column estimate)
    float hi2[3];       // 2-D: fast-path clamp (<= spanm, estimate there is exact)
    float band2;        // 2-D: guard band of the diagonal tests
    float hi3[3];       // 3-D: clamp bound (float)(n*inc - 1e-8); z uses inc_x (cpab_ops.cpp:141)
    float top3[3];      // 3-D lean path: largest coordinate it accepts (x, y: min(1, hi3); z: hi3)
    float ctop3[3];     // 3-D lean path: largest coordinate at which the one-sided estimate is still cell n-1
    float band3;        // 3-D: guard band of the separating-plane tests
    int clamp3;         // 3-D lean path: ctop3 < top3 on some axis (the estimate needs the clamp)
    int zpush3;         // 3-D lean path: the reference's z bound nz*inc_x - 1e-8 exceeds 1 (push != clamp for z)
    // ---- float64 check mode -------------------------------------------------------------------
    double wd[3];       // 1.0 / n
    double spand[3];    // n * wd
    double hi3d[3];     // n*wd - 1e-8, z with wd[0]
};

CPAB_HD float  abs_t(float x)  { return fabsf(x); }
CPAB_HD double abs_t(double x) { return fabs(x); }
CPAB_HD float  sign_t(float m, float s)   { return copysignf(m, s); }
CPAB_HD double sign_t(double m, double s) { return copysign(m, s); }

// the reference's mymin (cpab_ops.cpp:22-24): `!(b<a) ? a : round(b)`
CPAB_HD int pick_min(int a, double b) { return (b < (double)a) ? (int)round(b) : a; }

inline Geom make_geom(int ndim, const int* nc)
{
    Geom g;
    memset(&g, 0, sizeof(g));
    g.ndim = ndim;
    long cells = ndim == 1 ? 1 : (ndim == 2 ? 4 : 5);
    for (int j = 0; j < 3; ++j) {
        const int n = j < ndim ? nc[j] : 1;
        g.nc[j] = n;
        if (j < ndim) cells *= n;
        g.nf[j] = (float)n;
        g.nm1[j] = (float)(n - 1);
        g.w[j] = (float)(1.0 / n);
        g.span[j] = (float)n * g.w[j];                       // int*float -> float multiply
        g.wd[j] = 1.0 / n;
        g.spand[j] = n * g.wd[j];
        // 2-D, coordinate clamped to span - 1e-9 (cpab_ops.cpp:44-45,53-54)
        const double top = (double)g.span[j] - 0.000000001;
        const double rem = fmod(top, (double)g.w[j]);
        g.khi[j] = (float)pick_min(n - 1, (top - rem) / (double)g.w[j]);
        g.spanm[j] = nextafterf(g.span[j], 0.0f);
    }
    g.n_cells = (int)cells;
    // 2-D one-sided column estimate (divmod_up): multiplier nup >= 1/w, and the largest clamp value
    // at which the estimate is still the exact column n-1 (so that points on or beyond the upper
    // domain edge -- where zero-boundary flows park them -- stay on the fast path)
    float edge = 0.0f;
    for (int j = 0; j < 3; ++j) {
        const double inv = 1.0 / (double)g.w[j];
        float nup = (float)inv;
        if ((double)nup < inv) nup = nextafterf(nup, 3.0e38f);
        nup = nextafterf(nup, 3.0e38f);      // one ulp of slack: 1/w above is itself rounded
        g.nup[j] = nup;
        float hi = g.spanm[j];
        for (int it = 0; it < 64; ++it) {
            const double kest = floor((double)hi * (double)nup);
            if (kest <= (double)(g.nc[j] - 1) && (double)hi - kest * (double)g.w[j] >= 0.0) break;
            hi = nextafterf(hi, 0.0f);
        }
        g.hi2[j] = hi;
        const float gap = (g.span[j] - hi) * g.nf[j];
        if (j < 2 && gap > edge) edge = gap;
    }
    // |error| of the float local coordinates (3e-7) plus, for a coordinate clamped to hi2, the
    // distance of its local coordinate from 1: (span - hi2)/w (twice: it enters d1 and d2)
    g.band2 = 2e-6f + 2.4f * edge;
    for (int j = 0; j < 3; ++j) {
        const int wsel = (j == 2) ? 0 : j;                   // sic: nz * inc_x
        g.hi3[j] = (float)((double)((float)g.nc[j] * g.w[wsel]) - 1e-8);
        g.hi3d[j] = g.nc[j] * g.wd[wsel] - 1e-8;
    }
    // 3-D lean path (find_cell_3d_lean): accepted coordinate range and, where the cell width is
    // rounded up (n*w > 1: floor(1/w) = n-1 but floor(1*nup) = n), the clamp below which the
    // one-sided estimate stays in the last cell -- zero-boundary flows park points on the upper
    // faces, they must not take the rare path on every step.  The clamp moves a local coordinate
    // by at most (top - ctop)*n; the guard band of the plane tests grows by three times that.
    float edge3 = 0.0f;
    g.clamp3 = 0;
    for (int j = 0; j < 3; ++j) {
        g.top3[j] = j < 2 ? fminf(1.0f, g.hi3[j]) : g.hi3[j];
        float c = g.top3[j];
        const bool up = (double)g.nc[j] * (double)g.w[j] > (double)g.top3[j];      // true floor(top/w) is n-1
        if (up) {
            for (int it = 0; it < 64; ++it) {
                const double kest = floor((double)c * (double)g.nup[j]);
                if (kest <= (double)(g.nc[j] - 1) && (double)c - kest * (double)g.w[j] >= 0.0) break;
                c = nextafterf(c, 0.0f);
            }
        }
        g.ctop3[j] = c;
        if (c < g.top3[j]) g.clamp3 = 1;
        const float gap = (g.top3[j] - c) * g.nf[j];
        if (gap > edge3) edge3 = gap;
    }
    // + the (q - 0.5) + 0.5 round trip of the reference's push: up to 2^-25 per coordinate, three
    // coordinates per plane, in local units
    float nmax = 1.0f;
    for (int j = 0; j < 3; ++j) nmax = fmaxf(nmax, g.nf[j]);
    g.band3 = 4e-6f + 3.0f * edge3 + 3.0f * 2.9802322e-08f * nmax;
    g.zpush3 = g.hi3[2] > 1.0f ? 1 : 0;
    return g;
}

// -------------------------------------------------------------------------------------------------
// float32 building block: exact floor(p / w) and exact remainder p - floor(p/w)*w for p >= 0.
//
// Requires n <= 2^20 (enforced by the ABI) and p/w < 2^21.  With n*w = 1 + delta, |delta| <= 2^-24,
// the exact product p*n equals Q(1+delta) for the exact quotient Q = p/w, i.e. it is off by less
// than 1/16; the FFMA against 1.5*2^23 rounds it to the nearest integer, so the estimate is
// floor(Q) or floor(Q)+1 and never anything else.  r = fma(-k, w, p) is exact for both (a
// multiple of ulp(w) of magnitude <= w), its sign tells which, and r + w is exact as well.
// kf is returned as a float holding the integer.
CPAB_HD void divmod_exact(float p, float nf, float w, float& kf, float& r)
{
    float kMagic = 12582912.0f;                             // 1.5 * 2^23
#if defined(__CUDA_ARCH__)
    asm("" : "+f"(kMagic));   // keep it in a register: FFMA takes one non-register operand, let that be n
#endif
    const float t = fmaf(p, nf, kMagic);
    kf = t - kMagic;
    r = fmaf(-kf, w, p);
    if (r < 0.0f) { kf -= 1.0f; r += w; }
}

// ------------------------------------------------------------------------------------------- 1-D
CPAB_HD int find_cell_1d(float p0, const Geom& g)
{
#if defined(__CUDA_ARCH__)
    // the reference's float multiply, unfused, then floor-and-convert in ONE conversion-pipe
    // instruction (cvt.rmi saturates, NaN -> 0) and an integer clamp: the 1-D loop is otherwise
    // bound by the conversion pipe (FRND + F2I per step)
    const int c = __float2int_rd(__fmul_rn(p0, g.nf[0]));
    return min(max(c, 0), g.nc[0] - 1);
#else
    const float s = p0 * g.nf[0];
    const float c = fminf(fmaxf(floorf(s), 0.0f), g.nf[0] - 1.0f);
    return (int)c;
#endif
}

CPAB_HD int find_cell_1d(double p0, const Geom& g)
{
    int c = (int)floor(p0 * g.nc[0]);
    c = c > g.nc[0] - 1 ? g.nc[0] - 1 : c;
    return c < 0 ? 0 : c;
}

// ------------------------------------------------------------------------------------------- 2-D
// Literal double replay of cpab_ops.cpp:33-104 for T = float (point widened, widths float) and
// T = double (everything double).  Used as the last-resort path and as the fp64 check mode.
template <typename T>
CPAB_HD int find_cell_2d_replay_body(T p0, T p1, const Geom& g)
{
    const bool f32 = sizeof(T) == 4;
    const double qx = p0, qy = p1;
    const double wx = f32 ? (double)g.w[0] : g.wd[0];
    const double wy = f32 ? (double)g.w[1] : g.wd[1];
    const double sx = f32 ? (double)g.span[0] : g.spand[0];
    const double sy = f32 ? (double)g.span[1] : g.spand[1];
    const int nx = g.nc[0], ny = g.nc[1];
    double cx = qx > 0.0 ? qx : 0.0;
    double cy = qy > 0.0 ? qy : 0.0;
    if (sx - 0.000000001 < cx) cx = sx - 0.000000001;
    if (sy - 0.000000001 < cy) cy = sy - 0.000000001;
    const double rx = fmod(cx, wx), ry = fmod(cy, wy);
    const double lx = rx / wx, ly = ry / wy;
    int base = 4 * (pick_min(nx - 1, (cx - rx) / wx) + pick_min(ny - 1, (cy - ry) / wy) * nx);
    if (qx <= 0) {
        if (qy <= 0 && qy / wy < qx / wx) return base;
        if (qy >= sy && qy / wy - ny > -qx / wx) return base + 2;
        return base + 3;
    }
    if (qx >= sx) {
        if (qy <= 0 && -qy / wy > qx / wx - nx) return base;
        if (qy >= sy && qy / wy - ny > qx / wx - nx) return base + 2;
        return base + 1;
    }
    if (qy <= 0) return base;
    if (qy >= sy) return base + 2;
    if (lx < ly) return (1 - lx < ly) ? base + 2 : base + 3;
    if (1 - lx < ly) return base + 1;
    return base;
}

template <typename T>
CPAB_HD_NOINLINE int find_cell_2d_replay(T p0, T p1, const Geom& g) { return find_cell_2d_replay_body<T>(p0, p1, g); }

// In-domain triangle choice from exact remainders, without division where possible.
//   x = RN(rx/wx), y = RN(ry/wy) in the reference.  rx*wy and ry*wx are products of two floats and
//   therefore exact in double; distinct values differ by >= 2^-48 relative, which survives the
//   rounding of the quotients, so  x<y  <=>  rx*wy < ry*wx  exactly.
//   1-x<y is decided by the same products when |x+y-1| is clearly non-zero; inside a 2^-40 band the
//   reference's own divisions are replayed.
CPAB_HD int triangle_2d_exact_body(float rx, float ry, float wx, float wy)
{
    const double a = (double)rx * (double)wy;               // exact
    const double b = (double)ry * (double)wx;               // exact
    const double P = (double)wx * (double)wy;               // exact
    const bool x_lt_y = a < b;
    const double S = a + b;
    bool anti;                                              // 1 - x < y
    const double band = P * 9.094947017729282e-13;          // 2^-40
    if (S > P + band) anti = true;
    else if (S < P - band) anti = false;
    else {
        const double lx = (double)rx / (double)wx, ly = (double)ry / (double)wy;
        anti = (1 - lx) < ly;
    }
    return x_lt_y ? (anti ? 2 : 3) : (anti ? 1 : 0);
}

CPAB_HD_NOINLINE int triangle_2d_exact(float rx, float ry, float wx, float wy) { return triangle_2d_exact_body(rx, ry, wx, wy); }

// One-sided variant used by the 2-D search: with a multiplier nup >= 1/w and the FFMA rounding
// *down*, the estimate is floor(p * nup) >= floor(p / w) and exceeds it by at most one (only when
// p/w lies within n 2^-22 below an integer).  r = fma(-k, w, p) is exact for both, and negative
// exactly when the estimate is one too large -- a rare event that the caller folds into its
// guard-band branch instead of paying a compare and two predicated adds per axis and step.
// `magic` is 1.5 * 2^23; kernels pass it in a register they made opaque to ptxas (an FFMA takes
// one non-register operand: with the constant as an immediate the multiplier would be re-loaded
// from the constant bank every step).
CPAB_HD void divmod_up(float p, float nup, float w, float magic, float& kf, float& r)
{
#if defined(__CUDA_ARCH__)
    kf = __fmaf_rd(p, nup, magic) - magic;
#else
    (void)magic;
    kf = (float)floor((double)p * (double)nup);             // p * nup is exact in double
#endif
    r = fmaf(-kf, w, p);
}

// Rare path of find_cell_2d: a column/row estimate one too large, a point on or near a diagonal,
// or a point outside the domain near a corner.
// INLINE selects how the (rarer still) exact sub-paths are reached: by call from code that is itself
// out of line, inlined where the caller must stay free of calls (see find_cell_try).
template <bool INLINE>
CPAB_HD int find_cell_2d_rare_body(float p0, float p1, float kx, float rx, float ky, float ry, const Geom& g)
{
    if (rx < 0.0f) { kx -= 1.0f; rx += g.w[0]; }
    if (ry < 0.0f) { ky -= 1.0f; ry += g.w[1]; }
    const float xf = rx * g.nf[0], yf = ry * g.nf[1];
    const float d1 = xf - yf, d2 = (1.0f - xf) - yf;
    int tri;
    if (fminf(fabsf(d1), fabsf(d2)) < g.band2) {
        // outside the domain, or inside but clamped for the estimate: the reference's own sequence
        if (!(p0 > 0.0f) | (p0 >= g.span[0]) | !(p1 > 0.0f) | (p1 >= g.span[1]) | (p0 > g.hi2[0]) | (p1 > g.hi2[1]))
            return INLINE ? find_cell_2d_replay_body<float>(p0, p1, g) : find_cell_2d_replay<float>(p0, p1, g);
        if (rx == ry && g.w[0] == g.w[1]) {
            // exactly on the main diagonal of a square cell (a uniform_meshgrid commensurate with
            // the tessellation puts ~2 % of its points there at t = 0): the reference divides the
            // same operands twice, so x == y and `x < y` is false; 1 - x < x <=> x > 1/2 <=> 2 rx > w,
            // exact in float32
            tri = (rx + rx > g.w[0]) ? 1 : 0;
        } else {
            tri = INLINE ? triangle_2d_exact_body(rx, ry, g.w[0], g.w[1]) : triangle_2d_exact(rx, ry, g.w[0], g.w[1]);
        }
    } else {
        tri = (d1 < 0.0f ? 3 : 0) ^ (d2 < 0.0f ? 1 : 0);
    }
    return 4 * (int)fmaf(ky, g.nf[0], kx) + tri;
}
CPAB_HD_NOINLINE int find_cell_2d_rare(float p0, float p1, float kx, float rx, float ky, float ry, const Geom& g)
{
    return find_cell_2d_rare_body<false>(p0, p1, kx, rx, ky, ry, g);
}

// Fast path.  Coordinates are clamped to [0, hi2] (hi2 = a float just below the span): a
// coordinate at or below 0 gets column 0 and local coordinate 0, one at or above the span gets
// the last column and a local coordinate within band2/2.4 of 1.  With those values the in-domain
// diagonal tests reproduce the reference's out-of-bound branches (left -> 3, right -> 1,
// above -> 0, below -> 2) whenever a single axis is outside and the point is not within the
// guard band of a diagonal; every corner region (both axes outside) lands on a diagonal, i.e. in
// the band, from where the reference's own expression sequence is replayed.
// Returns true when the point needs find_cell_2d_rare (then `cell` is not set); the estimates are
// handed back so that the caller can continue there.
CPAB_HD bool find_cell_2d_fast(float p0, float p1, const Geom& g, float magic, int& cell,
                               float& kx, float& rx, float& ky, float& ry)
{
    const float c0 = fminf(fmaxf(p0, 0.0f), g.hi2[0]), c1 = fminf(fmaxf(p1, 0.0f), g.hi2[1]);
    float xf, yf;
#if defined(__CUDA_ARCH__)
    // both axes at once (packed FP32, cpab_f32x2.cuh); per lane identical to divmod_up
    const F2 pc = pk(c0, c1), mg = bc(magic);
    const F2 kf = sub2(fma2_rm(pc, pk(g.nup[0], g.nup[1]), mg), mg);
    const F2 r = fma2(kf, pk(-g.w[0], -g.w[1]), pc);
    const F2 xy = mul2(r, pk(g.nf[0], g.nf[1]));
    unpk(kf, kx, ky);
    unpk(r, rx, ry);
    unpk(xy, xf, yf);
#else
    divmod_up(c0, g.nup[0], g.w[0], magic, kx, rx);
    divmod_up(c1, g.nup[1], g.w[1], magic, ky, ry);
    xf = rx * g.nf[0];
    yf = ry * g.nf[1];
#endif
    // xf, yf: approximate local coordinates (|error| < 3e-7); the two diagonal tests
    const float d1 = xf - yf;                               // < 0  <=>  x < y
    const float d2 = (1.0f - xf) - yf;                      // < 0  <=>  1 - x < y
    // (x<y, 1-x<y) -> (0,0):0 (0,1):1 (1,1):2 (1,0):3  ==  (x<y ? 3 : 0) ^ (1-x<y ? 1 : 0);
    // outside the band d1, d2 are non-zero, so their sign bits are the comparisons
#if defined(__CUDA_ARCH__)
    const int tri = ((__float_as_int(d1) >> 31) & 3) ^ (int)((unsigned)__float_as_int(d2) >> 31);
#else
    const int tri = (d1 < 0.0f ? 3 : 0) ^ (d2 < 0.0f ? 1 : 0);
#endif
    cell = 4 * (int)fmaf(ky, g.nf[0], kx) + tri;
    return (fminf(fabsf(d1), fabsf(d2)) < g.band2) | (rx < 0.0f) | (ry < 0.0f);
}

CPAB_HD int find_cell_2d(float p0, float p1, const Geom& g, float magic = 12582912.0f)
{
    float kx, rx, ky, ry;
    int cell;
    if (find_cell_2d_fast(p0, p1, g, magic, cell, kx, rx, ky, ry))
        return find_cell_2d_rare(p0, p1, kx, rx, ky, ry, g);
    return cell;
}

CPAB_HD int find_cell_2d(double p0, double p1, const Geom& g)
{
    return find_cell_2d_replay<double>(p0, p1, g);
}

// ------------------------------------------------------------------------------------------- 3-D
// The reference's outside-the-box push (cpab_ops.cpp:119-137); all float/double ops as written.
template <typename T>
CPAB_HD void push_inside_3d(T& q0, T& q1, T& q2, T w0, T w1, T w2)
{
    const T half = (T)0.5;
    q0 -= half; q1 -= half; q2 -= half;
    const T ax = abs_t(q0), ay = abs_t(q1), az = abs_t(q2);
    const T shx = (ax < ay && ax < az) ? half * w0 : (T)0;
    const T shy = (ay < ax && ax < az) ? half * w1 : (T)0;
    const T shz = (az < ax && ax < ay) ? half * w2 : (T)0;
    if (ax > half) q0 = sign_t(half - shx, q0);
    if (ay > half) q1 = sign_t(half - shy, q1);
    if (az > half) q2 = sign_t(half - shz, q2);
    q0 += half; q1 += half; q2 += half;
}

// Literal replay of cpab_ops.cpp:138-184 from the (already pushed) point; T = float or double.
template <typename T>
CPAB_HD_NOINLINE int find_cell_3d_replay(T q0, T q1, T q2, const Geom& g)
{
    const bool f32 = sizeof(T) == 4;
    const T w0 = f32 ? (T)g.w[0] : (T)g.wd[0];
    const T w1 = f32 ? (T)g.w[1] : (T)g.wd[1];
    const T w2 = f32 ? (T)g.w[2] : (T)g.wd[2];
    const T h0 = f32 ? (T)g.hi3[0] : (T)g.hi3d[0];
    const T h1 = f32 ? (T)g.hi3[1] : (T)g.hi3d[1];
    const T h2 = f32 ? (T)g.hi3[2] : (T)g.hi3d[2];
    const T zero = (T)0;
    T c0 = q0 > zero ? q0 : zero; if (h0 < c0) c0 = h0;
    T c1 = q1 > zero ? q1 : zero; if (h1 < c1) c1 = h1;
    T c2 = q2 > zero ? q2 : zero; if (h2 < c2) c2 = h2;
    const double r0 = fmod((double)c0, (double)w0);
    const double r1 = fmod((double)c1, (double)w1);
    const double r2 = fmod((double)c2, (double)w2);
    const int i = pick_min(g.nc[0] - 1, ((double)c0 - r0) / (double)w0);
    const int j = pick_min(g.nc[1] - 1, ((double)c1 - r1) / (double)w1);
    const int k = pick_min(g.nc[2] - 1, ((double)c2 - r2) / (double)w2);
    int cell = 5 * (i + j * g.nc[0] + k * g.nc[0] * g.nc[1]);
    double x = r0 / (double)w0, y = r1 / (double)w1, z = r2 / (double)w2;
    if ((i + j + k) & 1) { const double t = x; x = y; y = 1 - t; }
    if (-x - y + z >= 0) cell += 1;
    else if (x + y + z - 2 >= 0) cell += 2;
    else if (-x + y - z >= 0) cell += 3;
    else if (x - y - z >= 0) cell += 4;
    return cell;
}

// Fast path; returns true when the point lies within the guard band of a separating plane and
// needs the exact replay (`cell` is then not meaningful).  The outside-the-box push stays inline:
// zero-boundary flows park points on the faces, where they wobble an ulp outside on many steps.
template <bool NEAR>
CPAB_HD bool find_cell_3d_fast_t(float p0, float p1, float p2, const Geom& g, int& cell, float* q, float& dist)
{
    float q0 = p0, q1 = p1, q2 = p2;
    const bool outside = q0 < 0.0f || q0 > 1.0f || q1 < 0.0f || q1 > 1.0f;   // sic: z is not tested (:119)
    if (outside) push_inside_3d<float>(q0, q1, q2, g.w[0], g.w[1], g.w[2]);
    q[0] = q0; q[1] = q1; q[2] = q2;
    float kx, ky, kz, rx, ry, rz;
    const float c0 = fminf(g.hi3[0], fmaxf(0.0f, q0)), c1 = fminf(g.hi3[1], fmaxf(0.0f, q1));
#if defined(__CUDA_ARCH__)
    {   // x and y as one packed sequence (cpab_f32x2.cuh); per lane identical to divmod_exact
        const F2 pc = pk(c0, c1), mg = bc(12582912.0f);
        const F2 kf = sub2(fma2(pc, pk(g.nf[0], g.nf[1]), mg), mg);
        const F2 r = fma2(kf, pk(-g.w[0], -g.w[1]), pc);
        unpk(kf, kx, ky);
        unpk(r, rx, ry);
        if (rx < 0.0f) { kx -= 1.0f; rx += g.w[0]; }
        if (ry < 0.0f) { ky -= 1.0f; ry += g.w[1]; }
    }
#else
    divmod_exact(c0, g.nf[0], g.w[0], kx, rx);
    divmod_exact(c1, g.nf[1], g.w[1], ky, ry);
#endif
    divmod_exact(fminf(g.hi3[2], fmaxf(0.0f, q2)), g.nf[2], g.w[2], kz, rz);
    kx = fminf(kx, g.nm1[0]);
    ky = fminf(ky, g.nm1[1]);
    kz = fminf(kz, g.nm1[2]);
    const int cube = (int)fmaf(fmaf(kz, g.nf[1], ky), g.nf[0], kx);
    const int par = (int)(kx + ky + kz);
    float x = rx * g.nf[0], y = ry * g.nf[1];
    const float z = rz * g.nf[2];
    if (par & 1) { const float t = x; x = y; y = 1.0f - t; }
    const float s = x + y, u = y - x;
    const float t1 = z - s, t2 = (s + z) - 2.0f, t3 = u - z, t4 = -u - z;
    const float nearest = fminf(fminf(fabsf(t1), fabsf(t2)), fminf(fabsf(t3), fabsf(t4)));
    const int tet = (t1 >= 0.0f) ? 1 : (t2 >= 0.0f) ? 2 : (t3 >= 0.0f) ? 3 : (t4 >= 0.0f) ? 4 : 0;
    cell = 5 * cube + tet;
    if (NEAR) {     // distance (local units) from the nearest face of the tetrahedron, cube faces included
        const float lo3 = fminf(fminf(x, y), z), hi3 = fmaxf(fmaxf(x, y), z);
        dist = fminf(nearest, fminf(lo3, 1.0f - hi3));
        if (outside) dist = -1.0f;      // the push is discontinuous in the point: never certified
    }
    return nearest < 4e-6f;
}
CPAB_HD bool find_cell_3d_fast(float p0, float p1, float p2, const Geom& g, int& cell, float* q)
{
    float dist;
    return find_cell_3d_fast_t<false>(p0, p1, p2, g, cell, q, dist);
}

// Complete search, out of line: the previous fast path (handles the outside-the-box push and the
// clamps inline) with the exact replay behind it.  The integration loops reach it only through the
// warp vote of find_cell_3d_lean.
CPAB_HD_NOINLINE int find_cell_3d_full(float p0, float p1, float p2, const Geom& g)
{
    int cell;
    float q[3];
    if (find_cell_3d_fast(p0, p1, p2, g, cell, q)) return find_cell_3d_replay<float>(q[0], q[1], q[2], g);
    return cell;
}

// Lean 3-D fast path (the 2-D scheme on three axes).
//   * Coordinates are clamped to [0, ctop].  This IS the reference's treatment of a point with at
//     most ONE coordinate outside [0, 1] (cpab_ops.cpp:119-141): if that coordinate is x or y, its
//     outside-the-box push only moves it onto the face (the `half*inc` shifts need two coordinates
//     outside; a z outside as well would be pushed too, to a different place than the later clamp
//     with the reference's z bound nz*inc_x - 1e-8 puts it); if it is z, there is no push and the
//     clamp to [0, n*inc - 1e-8] acts alone.  The push also sends the other coordinates through
//     (q - 0.5) + 0.5, which can move them by 2^-25: the guard band of the plane tests covers that
//     (band3).  Zero-boundary flows park points on the faces, where they wobble an ulp outside on
//     many steps -- edge and corner points on two or three faces at once -- and all of this has
//     to stay on the fast path.  Only a point with all three coordinates outside (a `half*inc`
//     shift may apply), or -- in tessellations whose z bound exceeds 1 -- with x or y outside and
//     z outside, leaves.
//   * The reference's `mymin(n-1, .)` is min(k, n-1): a coordinate of exactly 1.0 keeps local
//     coordinate 0 in the last cube (the reference's own quirk).
//   * One-sided column estimates as in divmod_up (x, y packed, z scalar): floor(c * nup) is the exact
//     column or one more, r = fma(-k, w, c) is exact and negative exactly in the second case (leaves).
//   * Parity swap, then the four separating planes in the reference's order; the corner tetrahedra
//     are disjoint, so outside the guard band at most one of t1..t4 is non-negative and the index
//     is 10 - sum_i i * signbit(t_i).
// Returns true when the point needs find_cell_3d_full instead.  ~55 instructions against ~100 of
// find_cell_3d_fast_t.  The cell index is in range even then.  NEAR: also the certificate's `dist`
// (see find_cell_near); -1 where the lean path does not apply.
template <bool NEAR>
CPAB_HD bool find_cell_3d_lean(float q0, float q1, float q2, const Geom& g, float magic, int& cell, float& dist)
{
    // the reference's ax, ay, az (same rounding); a coordinate is pushed when |q - 0.5| > 0.5
    const float dx = fabsf(q0 - 0.5f), dy = fabsf(q1 - 0.5f), dz = fabsf(q2 - 0.5f);
    const float c0 = fminf(fmaxf(q0, 0.0f), g.ctop3[0]), c1 = fminf(fmaxf(q1, 0.0f), g.ctop3[1]);
    const float c2 = fminf(fmaxf(q2, 0.0f), g.ctop3[2]);
    // When does the push differ from these clamps?  (1) All three coordinates outside: only then a
    // `half*inc` shift applies (each shift needs its coordinate to be the smallest of three |q - 0.5|
    // AND above 0.5) -- the eight corner regions.  (2) Tessellations whose z bound nz*inc_x - 1e-8
    // exceeds 1 (g.zpush3): the push, ENTERED on the exact tests x, y < 0 or > 1, moves a z with
    // dz > 0.5 onto its face where the clamp alone would leave it in (1, bound].  `c != q` is that
    // exact test (or a superset where ctop < 1).
    const bool out2 = (fminf(fminf(dx, dy), dz) > 0.5f) |
                      ((g.zpush3 != 0) & ((c0 != q0) | (c1 != q1)) & (dz > 0.5f));
    float kx, ky, kz, rx, ry, rz, x, y, z;
#if defined(__CUDA_ARCH__)
    {
        const F2 pc = pk(c0, c1), mg = bc(magic);
        const F2 kf = sub2(fma2_rm(pc, pk(g.nup[0], g.nup[1]), mg), mg);
        const F2 r = fma2(kf, pk(-g.w[0], -g.w[1]), pc);
        const F2 xy = mul2(r, pk(g.nf[0], g.nf[1]));
        unpk(kf, kx, ky);
        unpk(r, rx, ry);
        unpk(xy, x, y);
    }
#else
    divmod_up(c0, g.nup[0], g.w[0], magic, kx, rx);
    divmod_up(c1, g.nup[1], g.w[1], magic, ky, ry);
    x = rx * g.nf[0];
    y = ry * g.nf[1];
#endif
    divmod_up(c2, g.nup[2], g.w[2], magic, kz, rz);
    z = rz * g.nf[2];
    const float rmin = fminf(fminf(rx, ry), rz);
    kx = fminf(kx, g.nm1[0]);
    ky = fminf(ky, g.nm1[1]);
    kz = fminf(kz, g.nm1[2]);
    const int cube = (int)fmaf(fmaf(kz, g.nf[1], ky), g.nf[0], kx);
    const int par = (int)((kx + ky) + kz);
    if (par & 1) { const float t = x; x = y; y = 1.0f - t; }
    const float s = x + y, u = y - x;
    const float t1 = z - s, t2 = (s + z) - 2.0f, t3 = u - z, t4 = -u - z;
    const float nearest = fminf(fminf(fabsf(t1), fabsf(t2)), fminf(fabsf(t3), fabsf(t4)));
#if defined(__CUDA_ARCH__)
    const int neg = (int)((unsigned)__float_as_int(t1) >> 31) + 2 * (int)((unsigned)__float_as_int(t2) >> 31) +
                    3 * (int)((unsigned)__float_as_int(t3) >> 31) + 4 * (int)((unsigned)__float_as_int(t4) >> 31);
    const int tet = 10 - neg;
#else
    const int tet = (t1 >= 0.0f ? 1 : 0) + (t2 >= 0.0f ? 2 : 0) + (t3 >= 0.0f ? 3 : 0) + (t4 >= 0.0f ? 4 : 0);
#endif
    cell = 5 * cube + tet;
    const bool rare = out2 | !(rmin >= 0.0f) | !(nearest >= g.band3);
    if (NEAR) {
        const float lo3 = fminf(fminf(x, y), z), hi3 = fmaxf(fmaxf(x, y), z);
        // a clamped coordinate sits on a face (local coordinate 0 or next to 1): dist ~ 0, never certified
        dist = rare ? -1.0f : fminf(nearest, fminf(lo3, 1.0f - hi3));
    }
    return rare;
}

// Slow step of the integration loops (out of line, behind the warp vote).  The common reason to be
// here is a point within the guard band of a separating plane -- a uniform_meshgrid commensurate
// with the tessellation puts ~3 % of its points EXACTLY on such planes at t = 0 -- with everything
// else in order: for those the exact remainders of the one-sided estimate are the reference's
// fmod results, and only its double divisions and plane tests (cpab_ops.cpp:143-182) remain to be
// replayed (~150 instructions instead of ~500 for the complete search).
CPAB_HD_NOINLINE int find_cell_3d_slow(float q0, float q1, float q2, const Geom& g, float magic)
{
    const float dx = fabsf(q0 - 0.5f), dy = fabsf(q1 - 0.5f), dz = fabsf(q2 - 0.5f);
    const float c0 = fminf(fmaxf(q0, 0.0f), g.top3[0]), c1 = fminf(fmaxf(q1, 0.0f), g.top3[1]);
    const float c2 = fminf(fmaxf(q2, 0.0f), g.top3[2]);
    const bool out2 = (fminf(fminf(dx, dy), dz) > 0.5f) |
                      ((g.zpush3 != 0) & ((c0 != q0) | (c1 != q1)) & (dz > 0.5f));
    // (clamped to top3 here, not ctop3: the reference's own value; beyond ctop3 the estimate may be off)
    const bool beyond = (c0 > g.ctop3[0]) | (c1 > g.ctop3[1]) | (c2 > g.ctop3[2]);
    float kx, ky, kz, rx, ry, rz;
    divmod_up(c0, g.nup[0], g.w[0], magic, kx, rx);
    divmod_up(c1, g.nup[1], g.w[1], magic, ky, ry);
    divmod_up(c2, g.nup[2], g.w[2], magic, kz, rz);
    // (a point that enters the reference's push -- x or y outside, exact tests -- has all three
    //  coordinates sent through (q - 0.5) + 0.5, which matters inside the band: complete search)
    const bool entered = (q0 < 0.0f) | (q0 > 1.0f) | (q1 < 0.0f) | (q1 > 1.0f);
    if (out2 | beyond | entered | !(fminf(fminf(rx, ry), rz) >= 0.0f)) return find_cell_3d_full(q0, q1, q2, g);
    const int i = (int)fminf(kx, g.nm1[0]), j = (int)fminf(ky, g.nm1[1]), k = (int)fminf(kz, g.nm1[2]);
    int cell = 5 * (i + j * g.nc[0] + k * g.nc[0] * g.nc[1]);
    double x = (double)rx / (double)g.w[0], y = (double)ry / (double)g.w[1];
    const double z = (double)rz / (double)g.w[2];
    if ((i + j + k) & 1) { const double t = x; x = y; y = 1 - t; }
    if (-x - y + z >= 0) cell += 1;
    else if (x + y + z - 2 >= 0) cell += 2;
    else if (-x + y - z >= 0) cell += 3;
    else if (x - y - z >= 0) cell += 4;
    return cell;
}

CPAB_HD int find_cell_3d(float p0, float p1, float p2, const Geom& g, float magic = 12582912.0f)
{
    int cell;
    float dist;
    if (find_cell_3d_lean<false>(p0, p1, p2, g, magic, cell, dist)) return find_cell_3d_slow(p0, p1, p2, g, magic);
    return cell;
}

CPAB_HD int find_cell_3d(double p0, double p1, double p2, const Geom& g)
{
    double q0 = p0, q1 = p1, q2 = p2;
    if (q0 < 0.0 || q0 > 1.0 || q1 < 0.0 || q1 > 1.0)
        push_inside_3d<double>(q0, q1, q2, g.wd[0], g.wd[1], g.wd[2]);
    return find_cell_3d_replay<double>(q0, q1, q2, g);
}

// ------------------------------------------------------------------ certified search (float32)
// find_cell_near: the fast-path cell of a point together with `dist`, a lower bound -- in local
// cell units (fractions of a cell width) -- of the distance of the point from the nearest face of
// its simplex, domain boundary included.  The strict adjoint uses it as a certificate: a fast RK2
// iterate that is further than the accumulated rounding bound from every face lies in the same
// simplex as the reference's iterate, so the recorded cell sequence is the reference's
// (cpab_integrate.cu, "cell-sequence certificate").  No exact path here: whenever the fast search
// would need one (point within the guard band of a diagonal, column estimate one too large,
// coordinate clamped or pushed from outside the domain) `dist` is below every margin the
// certificate uses (>= band2 in 2-D, 4e-6 in 3-D), and the caller re-integrates that trajectory
// with the complete search instead.  A moved perturbation of |dp| (max norm, absolute) changes
// `dist` by at most |dp| * cert_scale(g).
CPAB_HD int find_cell_1d_near(float p0, const Geom& g, float magic, float& dist)
{
    // t = p * n is the reference's own rounded product (cpab_ops.cpp:28); rint and floor of it
    // come from one magic add (|t| < 2^22; beyond, `dist` is 0 or NaN-free garbage >= 0 and the
    // integer clamp below keeps the index in range)
    const float t = p0 * g.nf[0];
    const float tk = t + magic;
    const float kf = tk - magic;                             // rint(t)
    const float diff = t - kf;
    dist = fabsf(diff);
#if defined(__CUDA_ARCH__)
    int c = (__float_as_int(tk) - __float_as_int(magic)) + (__float_as_int(diff) >> 31);   // floor(t) when diff != 0
    return min(max(c, 0), g.nc[0] - 1);
#else
    const float c = fminf(fmaxf(floorf(t), 0.0f), g.nf[0] - 1.0f);
    return (int)c;
#endif
}

CPAB_HD int find_cell_2d_near(float p0, float p1, const Geom& g, float magic, float& dist)
{
    const float c0 = fminf(fmaxf(p0, 0.0f), g.hi2[0]), c1 = fminf(fmaxf(p1, 0.0f), g.hi2[1]);
    float kx, ky, xf, yf, u, v, d1, d2;
#if defined(__CUDA_ARCH__)
    const F2 pc = pk(c0, c1), mg = bc(magic);
    const F2 kf = sub2(fma2_rm(pc, pk(g.nup[0], g.nup[1]), mg), mg);
    const F2 r = fma2(kf, pk(-g.w[0], -g.w[1]), pc);
    const F2 xy = mul2(r, pk(g.nf[0], g.nf[1]));             // local coordinates (negative: estimate one too large)
    const F2 uv = sub2(bc(1.0f), xy);
    unpk(kf, kx, ky);
    unpk(xy, xf, yf);
    unpk(uv, u, v);
    const F2 dd = sub2(pk(xf, u), bc(yf));                   // (x - y, 1 - x - y)
    unpk(dd, d1, d2);
    const int tri = ((__float_as_int(d1) >> 31) & 3) ^ (int)((unsigned)__float_as_int(d2) >> 31);
#else
    float rx, ry;
    divmod_up(c0, g.nup[0], g.w[0], magic, kx, rx);
    divmod_up(c1, g.nup[1], g.w[1], magic, ky, ry);
    xf = rx * g.nf[0]; yf = ry * g.nf[1];
    u = 1.0f - xf; v = 1.0f - yf;
    d1 = xf - yf; d2 = u - yf;
    const int tri = (d1 < 0.0f ? 3 : 0) ^ (d2 < 0.0f ? 1 : 0);
#endif
    dist = fminf(fminf(fminf(xf, yf), fminf(u, v)), fminf(fabsf(d1), fabsf(d2)));
    return 4 * (int)fmaf(ky, g.nf[0], kx) + tri;
}

// bound of d(dist)/d|p|_inf: axis faces move by n_j, the diagonal planes by the sum over the axes
CPAB_HD float cert_scale(const Geom& g)
{
    float s = 0.0f;
    for (int j = 0; j < g.ndim; ++j) s += g.nf[j];
    return s;
}
// M_0 of the certificate: the guard band of the fast search itself, plus one ulp of the unit box
// in local units (an iterate whose pre-rounding value reaches the domain boundary is then flagged)
CPAB_HD float cert_floor(const Geom& g)
{
    const float band = g.ndim == 1 ? 4.0f * 5.9604645e-08f * g.nf[0] : g.ndim == 2 ? g.band2 : g.band3;
    return band + 1.1920929e-07f * cert_scale(g);
}
CPAB_HD int find_cell_3d_near(float p0, float p1, float p2, const Geom& g, float magic, float& dist)
{
    int cell;
    find_cell_3d_lean<true>(p0, p1, p2, g, magic, cell, dist);     // dist = -1 wherever the lean path does not apply
    return cell;
}
template <int NDIM> CPAB_HD int find_cell_near(const float* p, const Geom& g, float magic, float& dist)
{
    if (NDIM == 1) return find_cell_1d_near(p[0], g, magic, dist);
    if (NDIM == 2) return find_cell_2d_near(p[0], p[1], g, magic, dist);
    return find_cell_3d_near(p[0], p[1], p[2], g, magic, dist);
}
template <int NDIM> CPAB_HD int find_cell_near(const double* p, const Geom& g, float, float& dist);   // float32 only

// ------------------------------------------------------------------------------- generic front end
template <int NDIM, typename T>
CPAB_HD int find_cell(const T* p, const Geom& g)
{
    if (NDIM == 1) return find_cell_1d(p[0], g);
    if (NDIM == 2) return find_cell_2d(p[0], p[1], g);
    return find_cell_3d(p[0], p[1], p[2], g);
}

// Loop form for the integration kernels.  find_cell_try computes the cell on the fast path and
// returns true if the point needs the complete search instead (find_cell); the kernels vote on
// that flag and leave their inner loop warp-uniformly, so that the loop body contains no call
// (a call site pins live values to callee-saved registers and makes ptxas rebuild loop invariants
// after it every step).  Only the float32 2-D search has a separate rare path.
template <int NDIM> CPAB_HD bool find_cell_try(const float* p, const Geom& g, float magic, int& cell)
{
    if (NDIM == 2) {
        float kx, rx, ky, ry;
        return find_cell_2d_fast(p[0], p[1], g, magic, cell, kx, rx, ky, ry);
    }
    cell = find_cell<NDIM, float>(p, g);
    return false;
}
// Second half for the kernels' slow step: `est` carries the estimates of find_cell_try.
struct CellEst { float kx, rx, ky, ry; };
template <int NDIM> CPAB_HD bool find_cell_try(const float* p, const Geom& g, float magic, int& cell, CellEst& est)
{
    if (NDIM == 2) return find_cell_2d_fast(p[0], p[1], g, magic, cell, est.kx, est.rx, est.ky, est.ry);
    if (NDIM == 3) { float dist; return find_cell_3d_lean<false>(p[0], p[1], p[2], g, magic, cell, dist); }
    cell = find_cell<NDIM, float>(p, g);
    return false;
}
template <int NDIM> CPAB_HD int find_cell_finish(const float* p, const Geom& g, const CellEst& est)
{
    if (NDIM == 2) return find_cell_2d_rare(p[0], p[1], est.kx, est.rx, est.ky, est.ry, g);
    if (NDIM == 3) return find_cell_3d_slow(p[0], p[1], p[2], g, 12582912.0f);
    return find_cell<NDIM, float>(p, g);
}
template <int NDIM> CPAB_HD bool find_cell_try(const double* p, const Geom& g, float, int& cell, CellEst&)
{
    cell = find_cell<NDIM, double>(p, g);
    return false;
}
template <int NDIM> CPAB_HD int find_cell_finish(const double* p, const Geom& g, const CellEst&) { return find_cell<NDIM, double>(p, g); }

template <int NDIM> CPAB_HD bool find_cell_try(const double* p, const Geom& g, float, int& cell)
{
    cell = find_cell<NDIM, double>(p, g);
    return false;
}

}  // namespace cpab
