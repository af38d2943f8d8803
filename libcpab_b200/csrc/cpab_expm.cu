// cpab_expm.cu -- theta -> per-cell affine velocity matrices -> per-step transition matrices.
//
// Replaces, in one kernel and without any host round trip (paths under /root/reference):
//   libcpab/pytorch/transformer.py:146-150   B uploaded every call, Avees = B @ theta^T, reshape,
//                                            zero-row padding
//   libcpab/pytorch/transformer.py:153-155   Trels = expm(dT * AsSquare)[:, :ndim, :]
//   libcpab/pytorch/expm.py:11-54            batched Pade-13 scaling-and-squaring (20 torch ops,
//                                            LU solve, float32)
//
// One thread owns one (theta, cell): it contracts the cell's ndim*(ndim+1) basis rows with theta
// (basis stored transposed, [d, D], so the read is unit-stride across the cells of a warp; theta
// is broadcast from shared memory), writes A, then evaluates the same Pade-13 / squaring scheme on
// the (ndim+1)x(ndim+1) matrix dT*[A;0] in double and rounds once.  The float32 product dT*A is
// formed first, as the reference does, so the exponential sees the same argument.
//
// The matrices are tiny (2x2 / 3x3 / 4x4) and there are n_theta*nC of them (BASELINE configs:
// 2e5 .. 6.5e6), so this is register-resident SIMT work; the contraction is a [D x d].[d x n_theta]
// GEMM of 0.24 .. 2.6 GFLOP in total -- three orders of magnitude below the integration that
// follows -- and is left on the FP32 pipes.
#include "cpab_common.cuh"

namespace cpab {

namespace {

__device__ __constant__ double kPade13[14] = {
    64764752532480000., 32382376266240000., 7771770303897600., 1187353796428800.,
    129060195264000., 10559470521600., 670442572800., 33522128640., 1323241920.,
    40840800., 960960., 16380., 182., 1.};

template <int M> struct Mat { double a[M][M]; };

template <int M> __device__ __forceinline__ Mat<M> mm(const Mat<M>& x, const Mat<M>& y)
{
    Mat<M> r;
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < M; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < M; ++k) s = fma(x.a[i][k], y.a[k][j], s);
            r.a[i][j] = s;
        }
    return r;
}

// c1*x + c2*y + c3*z (+ c0*I)
template <int M>
__device__ __forceinline__ Mat<M> comb(double c1, const Mat<M>& x, double c2, const Mat<M>& y,
                                       double c3, const Mat<M>& z, double c0)
{
    Mat<M> r;
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < M; ++j)
            r.a[i][j] = c1 * x.a[i][j] + c2 * y.a[i][j] + c3 * z.a[i][j] + (i == j ? c0 : 0.0);
    return r;
}

// Solve Q X = P in place (X returned in P): Gaussian elimination, partial pivoting by
// conditional row exchange so that every index stays compile-time.
template <int M> __device__ __forceinline__ void solve(Mat<M>& Q, Mat<M>& P)
{
#pragma unroll
    for (int k = 0; k < M; ++k) {
#pragma unroll
        for (int r = k + 1; r < M; ++r) {
            if (fabs(Q.a[r][k]) > fabs(Q.a[k][k])) {
#pragma unroll
                for (int c = 0; c < M; ++c) {
                    double t = Q.a[r][c]; Q.a[r][c] = Q.a[k][c]; Q.a[k][c] = t;
                    t = P.a[r][c]; P.a[r][c] = P.a[k][c]; P.a[k][c] = t;
                }
            }
        }
        const double inv = 1.0 / Q.a[k][k];
#pragma unroll
        for (int r = k + 1; r < M; ++r) {
            const double f = Q.a[r][k] * inv;
#pragma unroll
            for (int c = k; c < M; ++c) Q.a[r][c] = fma(-f, Q.a[k][c], Q.a[r][c]);
#pragma unroll
            for (int c = 0; c < M; ++c) P.a[r][c] = fma(-f, P.a[k][c], P.a[r][c]);
        }
    }
#pragma unroll
    for (int k = M - 1; k >= 0; --k) {
        const double inv = 1.0 / Q.a[k][k];
#pragma unroll
        for (int c = 0; c < M; ++c) {
            double s = P.a[k][c];
#pragma unroll
            for (int r = k + 1; r < M; ++r) s = fma(-Q.a[k][r], P.a[r][c], s);
            P.a[k][c] = s * inv;
        }
    }
}

// expm by the reference's scheme (pytorch/expm.py:11-36): Frobenius norm, squarings =
// max(0, ceil(log2(norm / 5.3719...))), Pade-13, solve (V-U) R = (V+U), square back.
template <int M> __device__ __forceinline__ Mat<M> expm_pade13(Mat<M> A)
{
    double fro = 0.0;
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < M; ++j) fro = fma(A.a[i][j], A.a[i][j], fro);
    fro = sqrt(fro);
    int nsq = 0;
    if (fro > 5.371920351148152) nsq = (int)ceil(log2(fro / 5.371920351148152));
    // a diverged theta (an entry of +-inf, norm = +inf) would saturate the conversion to INT_MAX and
    // run ~2^31 squarings per thread; beyond 2^1100 the scaling underflows anyway.  The result is
    // then inf/NaN (0 * inf in the scaled matrix), which is what torch's expm gives as well.
    if (nsq > 1100) nsq = 1100;
    if (nsq > 0) {
        const double sc = ldexp(1.0, -nsq);
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
            for (int j = 0; j < M; ++j) A.a[i][j] *= sc;
    }
    const double* b = kPade13;
    const Mat<M> A2 = mm(A, A), A4 = mm(A2, A2), A6 = mm(A4, A2);
    Mat<M> U = mm(A, comb(b[7], A6, b[5], A4, b[3], A2, b[1]));
    {
        const Mat<M> hi = mm(A, mm(A6, comb(b[13], A6, b[11], A4, b[9], A2, 0.0)));
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
            for (int j = 0; j < M; ++j) U.a[i][j] += hi.a[i][j];
    }
    Mat<M> V = comb(b[6], A6, b[4], A4, b[2], A2, b[0]);
    {
        const Mat<M> hi = mm(A6, comb(b[12], A6, b[10], A4, b[8], A2, 0.0));
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
            for (int j = 0; j < M; ++j) V.a[i][j] += hi.a[i][j];
    }
    Mat<M> P, Q;
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < M; ++j) { P.a[i][j] = V.a[i][j] + U.a[i][j]; Q.a[i][j] = V.a[i][j] - U.a[i][j]; }
    solve(Q, P);
    for (int s = 0; s < nsq; ++s) P = mm(P, P);
    return P;
}

// ------------------------------------------------------------------------------------------------
// One thread owns one cell for TT consecutive thetas: every basis row fetched from L2 is reused
// TT times from registers (the contraction is a skinny GEMM; for 65536 thetas of a 1-D [100]
// tessellation the un-tiled version re-read the 79 KB basis 65536 times).
template <typename T, int NDIM, int TT>
__global__ void __launch_bounds__(128)
k_theta_to_trels(const T* __restrict__ basis_t, const T* __restrict__ theta, T* __restrict__ As,
                 T* __restrict__ trels, int n_cells, int d, int nsteps, int n_theta)
{
    constexpr int PPC = Dim<NDIM>::kPpc;
    constexpr int M = NDIM + 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* sth = reinterpret_cast<T*>(smem_raw);                    // [TT][d]
    const int t0 = blockIdx.y * TT;
    const int D = n_cells * PPC;
    for (int x = threadIdx.x; x < TT * d; x += blockDim.x) {
        const int u = x / d, j = x - u * d;
        sth[x] = (t0 + u < n_theta) ? theta[(size_t)(t0 + u) * d + j] : (T)0;
    }
    __syncthreads();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;

    T acc[TT][PPC];
#pragma unroll
    for (int u = 0; u < TT; ++u)
#pragma unroll
        for (int e = 0; e < PPC; ++e) acc[u][e] = 0;
    const T* col = basis_t + (size_t)c * PPC;
    for (int j = 0; j < d; ++j) {
        const T* row = col + (size_t)j * D;
        T b[PPC];
#pragma unroll
        for (int e = 0; e < PPC; ++e) b[e] = __ldg(row + e);
#pragma unroll
        for (int u = 0; u < TT; ++u) {
            const T th = sth[u * d + j];
#pragma unroll
            for (int e = 0; e < PPC; ++e) acc[u][e] = fma(b[e], th, acc[u][e]);
        }
    }
    const T dT = (T)(1.0 / nsteps);
#pragma unroll 1
    for (int u = 0; u < TT; ++u) {
        if (t0 + u >= n_theta) break;
        T* Aout = As + ((size_t)(t0 + u) * n_cells + c) * PPC;
        // dT * A in the working precision (as the reference: `dT*AsSquare` on a float32 tensor),
        // then the exponential in double
        Mat<M> X;
#pragma unroll
        for (int e = 0; e < PPC; ++e) {
            T v = acc[0][e];
#pragma unroll
            for (int q = 1; q < TT; ++q) v = (u == q) ? acc[q][e] : v;      // keeps acc in registers
            Aout[e] = v;
            X.a[e / M][e % M] = (double)(T)(dT * v);
        }
#pragma unroll
        for (int j = 0; j < M; ++j) X.a[NDIM][j] = 0.0;
        const Mat<M> E = expm_pade13<M>(X);
        T* Tout = trels + ((size_t)(t0 + u) * n_cells + c) * PPC;
#pragma unroll
        for (int i = 0; i < NDIM; ++i)
#pragma unroll
            for (int j = 0; j < M; ++j) Tout[i * M + j] = (T)E.a[i][j];
    }
}

// standalone batched expm of [n, M, M] (tests; and the reference's expm() as an op)
template <typename T, int M>
__global__ void __launch_bounds__(128) k_expm(const T* __restrict__ A, T* __restrict__ E, long n)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Mat<M> X;
#pragma unroll
    for (int r = 0; r < M; ++r)
#pragma unroll
        for (int c = 0; c < M; ++c) X.a[r][c] = (double)A[(i * M + r) * M + c];
    const Mat<M> R = expm_pade13<M>(X);
#pragma unroll
    for (int r = 0; r < M; ++r)
#pragma unroll
        for (int c = 0; c < M; ++c) E[(i * M + r) * M + c] = (T)R.a[r][c];
}

template <typename T, int NDIM>
int theta_to_trels_t(const Geom& g, int nsteps, int n_theta, int d, const void* basis_t,
                     const void* theta, void* As, void* trels, cudaStream_t st)
{
    constexpr int TT = NDIM == 1 ? 16 : (NDIM == 2 ? 8 : 4);
    if (n_theta < 1024) {
        // small batches: one (theta, cell) per thread keeps more SMs busy than reusing basis rows
        dim3 grid1((unsigned)((g.n_cells + 127) / 128), (unsigned)n_theta);
        prof_begin(kProfThetaToTrels, st);
        k_theta_to_trels<T, NDIM, 1><<<grid1, 128, (size_t)d * sizeof(T), st>>>(
            (const T*)basis_t, (const T*)theta, (T*)As, (T*)trels, g.n_cells, d, nsteps, n_theta);
        prof_end(kProfThetaToTrels, st);
        count_launch();
        CPAB_CUDA_OK(cudaGetLastError());
        return kOk;
    }
    const long ty = (n_theta + TT - 1) / TT;
    if (ty > 65535) {
        // grid.y limit: process in slabs
        const int slab = 65535 * TT;
        for (int t0 = 0; t0 < n_theta; t0 += slab) {
            const int n = n_theta - t0 < slab ? n_theta - t0 : slab;
            const size_t off = (size_t)t0 * g.n_cells * Dim<NDIM>::kPpc;
            int rc = theta_to_trels_t<T, NDIM>(g, nsteps, n, d, basis_t, (const T*)theta + (size_t)t0 * d,
                                               (T*)As + off, (T*)trels + off, st);
            if (rc != kOk) return rc;
        }
        return kOk;
    }
    dim3 grid((unsigned)((g.n_cells + 127) / 128), (unsigned)ty);
    const size_t smem = (size_t)TT * d * sizeof(T);
    if (smem > 48 * 1024) { set_error("theta dimension %d too large", d); return kErrUnsupported; }
    prof_begin(kProfThetaToTrels, st);
    k_theta_to_trels<T, NDIM, TT><<<grid, 128, smem, st>>>((const T*)basis_t, (const T*)theta, (T*)As,
                                                            (T*)trels, g.n_cells, d, nsteps, n_theta);
    prof_end(kProfThetaToTrels, st);
    count_launch();
    CPAB_CUDA_OK(cudaGetLastError());
    return kOk;
}

template <typename T, int M>
int expm_t(long n, const void* A, void* E, cudaStream_t st)
{
    k_expm<T, M><<<(unsigned)((n + 127) / 128), 128, 0, st>>>((const T*)A, (T*)E, n);
    count_launch();
    CPAB_CUDA_OK(cudaGetLastError());
    return kOk;
}

}  // namespace

int launch_theta_to_trels(int dtype, const Geom& g, int nsteps, int n_theta, int d,
                          const void* basis_t, const void* theta, void* As, void* trels,
                          cudaStream_t st)
{
    if (n_theta == 0) return kOk;
#define GO(T) (g.ndim == 1 ? theta_to_trels_t<T, 1>(g, nsteps, n_theta, d, basis_t, theta, As, trels, st) \
             : g.ndim == 2 ? theta_to_trels_t<T, 2>(g, nsteps, n_theta, d, basis_t, theta, As, trels, st) \
                           : theta_to_trels_t<T, 3>(g, nsteps, n_theta, d, basis_t, theta, As, trels, st))
    return dtype == kF32 ? GO(float) : GO(double);
#undef GO
}

int launch_expm(int dtype, int m, long n, const void* A, void* E, cudaStream_t st)
{
    if (n == 0) return kOk;
    if (m < 2 || m > 4) { set_error("expm: matrix size %d not in 2..4", m); return kErrArgument; }
#define GO(T) (m == 2 ? expm_t<T, 2>(n, A, E, st) : m == 3 ? expm_t<T, 3>(n, A, E, st) : expm_t<T, 4>(n, A, E, st))
    return dtype == kF32 ? GO(float) : GO(double);
#undef GO
}

}  // namespace cpab
