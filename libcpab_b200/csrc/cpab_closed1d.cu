// cpab_closed1d.cu -- closed-form ("hit-time") integration of a 1-D CPA velocity field.
//
// NOT IN THE REFERENCE (SURVEY.md section 0.2, row a11 of section 8): libcpab integrates with 50
// fixed steps in every backend.  This is the algorithm BASELINE.json's north_star describes --
// locate the cell, take the analytic flow to the cell boundary, compute the hit time, cross into
// the next cell, repeat until t = 1 -- following Freifeld, Hauberg, Batmanghelich, Fisher,
// "Transformations based on continuous piecewise-affine velocity fields" (TPAMI 2017), where
// the construction is closed-form in one dimension.  It is an OPT-IN mode
// (Cpab.params.closed_form = True; this file is the 1-D case): parity with the reference is defined against the
// fixed-step kernels; this mode converges to the reference's semantics as nstepsolver -> inf and
// is validated that way (tests/test_gpu_closed_form.py).  In 2-D/3-D the hit time is the root of
// a sum of exponentials and has no closed form; cpab_closednd.cu works with its series there.
//
// Within cell c, v(x) = a x + b:
//     psi(x0, t) = x0 e^{at} + b t phi1(at),             phi1(z) = (e^z - 1)/z
//     t_hit      = (Delta / v_s) L(a Delta / v_s),       L(z) = log1p(z)/z,
//                  Delta = x_b - x_s, v_s = v(x_s); no hit if 1 + a Delta / v_s <= 0.
// Gradient (same G[theta][cell] + G.B epilogue as the fixed-step adjoint): with m crossings,
// T = 1 - sum_i t_i the time spent in the last cell and x_f = psi_m(x_m, T),
//     d x_f = psi_a da_m + psi_b db_m - v(x_f) sum_{i<m} (t_{i,a} da_i + t_{i,b} db_i)
//     psi_b = t phi1(at),  psi_a = x0 t e^{at} + b t^2 phi1'(at)
//     t_b = -Delta / (v_b v_s),  t_a = (b Delta / (v_b v_s) - t) / a   (series for small a Delta / v_s)
//     d x_f / d x_0 = v(x_f) / v(x_0).
// The loop trip count varies per trajectory (1 + number of crossings, mean ~3 for theta ~ N(0,I)),
// so lanes of a warp diverge only in how many times they iterate; the body is branch-light.
#include "cpab_common.cuh"

namespace cpab {

namespace {

template <typename T> __device__ __forceinline__ T phi1(T z)      // (e^z - 1)/z
{
    return fabs(z) < (T)1e-4 ? (T)1 + z * ((T)0.5 + z * (T)(1.0 / 6.0)) : expm1(z) / z;
}
template <typename T> __device__ __forceinline__ T dphi1(T z)     // d/dz (e^z - 1)/z
{
    if (fabs(z) < (T)(sizeof(T) == 4 ? 2e-2 : 1e-3))
        return (T)0.5 + z * ((T)(1.0 / 3.0) + z * ((T)0.125 + z * ((T)(1.0 / 30.0) + z * (T)(1.0 / 144.0))));
    const T e = exp(z);
    return (z * e - (e - (T)1)) / (z * z);
}
template <typename T> __device__ __forceinline__ T log1p_over(T z)   // log1p(z)/z
{
    return fabs(z) < (T)1e-4 ? (T)1 - z * ((T)0.5 - z * (T)(1.0 / 3.0)) : log1p(z) / z;
}

// One segment: from x (inside or on the edge of cell c, velocity v = a x + b != 0) towards the
// boundary in the direction of motion.  Returns the hit time (inf if never) and the boundary.
template <typename T>
__device__ __forceinline__ T hit_time(T x, T a, T v, int c, int nc, T& xb, int& cnext)
{
    const T inf = (T)INFINITY;
    const bool right = v > (T)0;
    cnext = right ? c + 1 : c - 1;
    if (cnext < 0 || cnext >= nc) return inf;               // last cell: its affine map extends outside
    xb = (T)(right ? c + 1 : c) / (T)nc;
    const T delta = xb - x;
    const T z = a * delta / v;
    if (!(z > (T)-1)) return inf;                           // velocity vanishes before the boundary
    return (delta / v) * log1p_over(z);
}

template <typename T, bool BACKWARD>
__global__ void __launch_bounds__(256)
k_closed1d(const T* __restrict__ points, const T* __restrict__ As, const T* __restrict__ gout,
           T* __restrict__ out, T* __restrict__ G, T* __restrict__ dpoints, long nP, int broadcast,
           int nc, int chunks, int chunk_pts, const __grid_constant__ Geom g)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* sA = reinterpret_cast<T*>(smem_raw);
    const int theta = blockIdx.x / chunks;
    const int chunk = blockIdx.x - theta * chunks;
    for (int i = threadIdx.x; i < 2 * nc; i += blockDim.x) sA[i] = As[(size_t)theta * 2 * nc + i];
    __syncthreads();
    const T* src = points + (broadcast ? (size_t)theta * nP : 0);
    const long begin = (long)chunk * chunk_pts;
    const long end = begin + chunk_pts < nP ? begin + chunk_pts : nP;
    for (long i = begin + threadIdx.x; i < end; i += blockDim.x) {
        const T x0 = src[i];
        // ---- forward: walk the cells
        T x = x0, t = (T)1;
        bool fixed = false;
        int c = find_cell_1d(x, g);
        for (int it = 0; it <= nc; ++it) {
            const T a = sA[2 * c], b = sA[2 * c + 1];
            const T v = a * x + b;
            // fixed point of the field: stay in this cell for the REMAINING time (psi(x, t) = x there, and
            // psi_a = x t e^{at}, psi_b = t phi1(at) below are the correct sensitivities -- zeroing t
            // would make dL/dtheta vanish at theta = 0, the usual identity initialisation)
            if (v == (T)0) { fixed = true; break; }
            T xb = x;
            int cn = c;
            const T th = hit_time(x, a, v, c, nc, xb, cn);
            if (!(th < t)) break;
            x = xb; t -= th; c = cn;
        }
        const T am = sA[2 * c], bm = sA[2 * c + 1];
        const T zm = am * t;
        const T em = exp(zm);
        const T xf = fixed ? x : x * em + bm * t * phi1(zm);      // (a fixed point stays put exactly, not up to the cancellation in psi)
        if (!BACKWARD) {
            out[(size_t)theta * nP + i] = xf;
            continue;
        }
        // ---- backward: contributions per visited cell (replay the walk; x_f and v(x_f) known)
        const T gup = gout[(size_t)theta * nP + i];
        const T vf = am * xf + bm;
        T* Gt = G + (size_t)theta * 2 * nc;
        // last cell: psi_a, psi_b at (x, t)
        atomicAdd(Gt + 2 * c, gup * (x * t * em + bm * t * t * dphi1(zm)));
        atomicAdd(Gt + 2 * c + 1, gup * (t * phi1(zm)));
        const T v0 = sA[2 * find_cell_1d(x0, g)] * x0 + sA[2 * find_cell_1d(x0, g) + 1];
        if (dpoints != nullptr) dpoints[(size_t)theta * nP + i] = gup * (v0 != (T)0 ? vf / v0 : em);
        // crossed cells
        T xs = x0, tr = (T)1;
        int cs = find_cell_1d(x0, g);
        for (int it = 0; it <= nc; ++it) {
            const T a = sA[2 * cs], b = sA[2 * cs + 1];
            const T v = a * xs + b;
            if (v == (T)0) break;
            T xb = xs;
            int cn = cs;
            const T th = hit_time(xs, a, v, cs, nc, xb, cn);
            if (!(th < tr)) break;
            const T delta = xb - xs;
            const T vb = a * xb + b;
            const T tb = -delta / (vb * v);
            const T z = a * delta / v;
            T ta;
            if (fabs(z) < (T)(sizeof(T) == 4 ? 3e-2 : 1e-3)) {
                // t = (delta/v) L(z), z = a delta / v, v = a xs + b:
                //   dt/da = -(delta xs / v^2) L(z) + (delta/v) L'(z) dz/da,  dz/da = delta b / v^2
                //   L'(z) = -1/2 + 2z/3 - 3z^2/4 + 4z^3/5 - ...
                const T Lp = (T)-0.5 + z * ((T)(2.0 / 3.0) + z * ((T)-0.75 + z * ((T)0.8 + z * (T)(-5.0 / 6.0))));
                ta = -(delta * xs / (v * v)) * log1p_over(z) + (delta / v) * Lp * (delta * b / (v * v));
            } else {
                ta = (b * delta / (vb * v) - th) / a;
            }
            atomicAdd(Gt + 2 * cs, -gup * vf * ta);
            atomicAdd(Gt + 2 * cs + 1, -gup * vf * tb);
            xs = xb; tr -= th; cs = cn;
        }
    }
}

template <typename T>
int closed1d_t(bool backward, const Geom& g, int n_theta, long nP, int broadcast, const void* points,
               const void* As, const void* gout, void* out, void* G, void* dpoints, cudaStream_t st)
{
    const int nc = g.nc[0];
    const size_t smem = (size_t)2 * nc * sizeof(T);
    if (smem > kMaxSmemBytes) { set_error("closed form: tessellation too large for shared memory"); return kErrUnsupported; }
    int chunk_pts = 2048;
    if (nP < chunk_pts) chunk_pts = (int)((nP + 255) / 256 * 256);
    const int chunks = (int)((nP + chunk_pts - 1) / chunk_pts);
    const long long blocks = (long long)n_theta * chunks;
    if (blocks > 0x7fffffffLL) { set_error("grid too large"); return kErrUnsupported; }
    if (backward) {
        auto kern = k_closed1d<T, true>;
        if (smem > 48 * 1024) CPAB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        prof_begin(kProfBackward, st);
        kern<<<(unsigned)blocks, 256, smem, st>>>((const T*)points, (const T*)As, (const T*)gout, nullptr, (T*)G,
                                                  (T*)dpoints, nP, broadcast, nc, chunks, chunk_pts, g);
        prof_end(kProfBackward, st);
    } else {
        auto kern = k_closed1d<T, false>;
        if (smem > 48 * 1024) CPAB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        prof_begin(kProfForward, st);
        kern<<<(unsigned)blocks, 256, smem, st>>>((const T*)points, (const T*)As, nullptr, (T*)out, nullptr,
                                                  nullptr, nP, broadcast, nc, chunks, chunk_pts, g);
        prof_end(kProfForward, st);
    }
    count_launch();
    CPAB_CUDA_OK(cudaGetLastError());
    return kOk;
}

}  // namespace

int launch_closed1d_forward(int dtype, const Geom& g, int n_theta, long nP, int broadcast,
                            const void* points, const void* As, void* out, cudaStream_t st)
{
    if (n_theta == 0 || nP == 0) return kOk;
    return dtype == kF32 ? closed1d_t<float>(false, g, n_theta, nP, broadcast, points, As, nullptr, out, nullptr, nullptr, st)
                         : closed1d_t<double>(false, g, n_theta, nP, broadcast, points, As, nullptr, out, nullptr, nullptr, st);
}

// G [n_theta, 2 nc] must be zero-initialised by the caller (launch_closed1d_backward_theta does)
int launch_closed1d_backward(int dtype, const Geom& g, int n_theta, long nP, int broadcast,
                             const void* points, const void* As, const void* gout, void* G,
                             void* dpoints, cudaStream_t st)
{
    if (n_theta == 0 || nP == 0) return kOk;
    return dtype == kF32 ? closed1d_t<float>(true, g, n_theta, nP, broadcast, points, As, gout, nullptr, G, dpoints, st)
                         : closed1d_t<double>(true, g, n_theta, nP, broadcast, points, As, gout, nullptr, G, dpoints, st);
}

}  // namespace cpab
