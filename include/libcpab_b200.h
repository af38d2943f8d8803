/*
 * libcpab_b200.h -- C ABI of the B200-native CPAB transformation library (libcpab_b200.so).
 *
 * This is the drop-in boundary for the hot path of SkafteNicki/libcpab when
 * backend='pytorch', device='gpu'.  Each entry point names the reference interface it replaces
 * (paths relative to the reference repository root).  The reference binds its native code through
 * a pybind11/torch extension (libcpab/pytorch/transformer_cuda.cpp:73-76: forward, backward);
 * here the same operations are plain C functions over device pointers, callable from ctypes
 * (see INTEGRATION.md for the reference-side stub).
 *
 * Conventions
 *   - All data pointers are DEVICE pointers to contiguous row-major arrays of `dtype`
 *     (CPAB_F32 = float, CPAB_F64 = double "check mode") on the current device.
 *   - Every entry point is safe under CUDA-graph stream capture (no host synchronisation
 *     or allocation).  The work counters of the persistent kernels and the scratch of the gradient's
 *     cell-sequence certificate live in the CALLER's workspace (zeroed by the first kernel of the
 *     same launch sequence): nothing is shared between launches, streams or captured graphs.
 *   - The caller owns every buffer (outputs and workspace included); the library allocates no
 *     persistent memory and keeps no state besides the per-thread error string, the per-thread
 *     tuning knobs set through cpab_b200_set_tuning and the (off by default) profiling accumulators.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all work is
 *     enqueued on it and the call returns without synchronising.
 *   - `nc` points to `ndim` host ints (tessellation size per dimension), ndim in {1,2,3}.
 *   - Layouts: points/grids are planar [ndim, nP] or
 *     [n_theta, ndim, nP]; affine parameters are [n_theta, nC, ndim, ndim+1] with
 *     nC = nx | 4*nx*ny | 5*nx*ny*nz simplices; the basis is [D, d], D = nC*ndim*(ndim+1).
 *   - Return value: CPAB_OK (0) or a negative CPAB_ERR_* code; cpab_b200_last_error() gives the
 *     message for the calling thread.  The library never calls exit() (the reference does:
 *     libcpab/pytorch/transformer_cuda.cu:8-16).
 */
#ifndef LIBCPAB_B200_H
#define LIBCPAB_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CPAB_OK               0
#define CPAB_ERR_ARGUMENT    (-1)
#define CPAB_ERR_CUDA        (-2)
#define CPAB_ERR_UNSUPPORTED (-3)
#define CPAB_ERR_WORKSPACE   (-4)

#define CPAB_F32 0
#define CPAB_F64 1

/* flags */
#define CPAB_FLAG_FAST_MATH 1   /* forward: contract a*b+c into FMAs (default: every operation
                                   rounded as the CPU reference does -> bit-identical results) */
#define CPAB_FLAG_FAST_GRAD 2   /* backward: skip the cell-sequence certificate.  Default: every
                                   trajectory provably follows the cells the reference's float32 RK2
                                   iterates visit (libcpab/core/cpab_ops.cpp:289-366), so the gradient
                                   differs from the reference's by rounding only; with this flag an
                                   iterate within rounding of a cell face may follow the neighbouring
                                   cell (~1e-6 per point and step; moves a theta's gradient by up to
                                   ~1e-3 relative). */

/* ABI revision of this header; bumped on any signature change. */
int cpab_b200_abi_version(void);

/* Message of the last failing call on this thread ("" if none). */
const char* cpab_b200_last_error(void);

/* Compile-time facts: "sm_100a;cuda=12.9;..." */
const char* cpab_b200_build_info(void);

/* Experiment knobs, per calling thread ("fwd_ppt", "chunk_pts", "chunk_auto", "bwd_seg", "bwd_stage",
 * "bwd_block", "interp_variant" 0-11, "interp_max_ctas", "closed_refill", "closed_stage"); results never depend on them beyond
 * floating-point summation order in the gradient.  ("interp_variant" 12-15 are measurement probes of
 * the interpolate forward that deliberately skip work: tools/interp_variants.py only.) */
int cpab_b200_set_tuning(const char* key, int value);

/* Number of kernels this library has launched in this process (bench.py's `gpu_launches`). */
long long cpab_b200_launch_count(void);

/* Per-kernel device timing.  enable(1) clears the accumulators and makes the library bracket its
 * dominant kernels with CUDA events on the stream they are launched on; read() synchronises the
 * recorded events and returns the accumulated milliseconds and launch count of one slot:
 * 0 forward, 1 adjoint backward, 2 interpolate forward, 3 interpolate backward,
 * 4 theta_to_trels, 5 gradient epilogue.  Off by default (no events are created). */
int cpab_b200_profile_enable(int on);
int cpab_b200_profile_read(int slot, double* total_ms, long long* launches);

/* Diagnostic: `blocks` CTAs of 256 threads each run `iters` x 64 dependent-chain FP32 FMAs
 * (8 independent chains per thread).  bench.py times it to obtain the measured FP32 peak that
 * the integration roofline is quoted against.  `out` is one device float (never written). */
int cpab_b200_fp32_fma_probe(int blocks, int iters, void* out, void* stream);

/*
 * Cell index of every point.  Replaces findcellidx, libcpab/core/cpab_ops.cpp:26-190 /
 * libcpab/core/cpab_ops.cu:14-227 (the index the integrators use), bit-exact with the CPU one.
 *   points [ndim, nP] -> out_idx int32 [nP]
 */
int cpab_b200_findcellidx(int dtype, int ndim, const int* nc, const void* points, long nP,
                          int* out_idx, void* stream);

/*
 * theta -> velocity-field matrices and per-step transition matrices.  Replaces
 * libcpab/pytorch/transformer.py:146-155 (B @ theta^T, reshape, zero-row pad, expm(dT*A)) and
 * libcpab/pytorch/expm.py:11-54.
 *   basis_t [d, D] (the basis TRANSPOSED, i.e. the reference's `Bs` tensor, transformer.py:174)
 *   theta   [n_theta, d]
 *   As      [n_theta, nC, ndim, ndim+1]  out
 *   trels   [n_theta, nC, ndim, ndim+1]  out, expm(A/nsteps) top rows
 */
int cpab_b200_theta_to_trels(int dtype, int ndim, const int* nc, int nsteps, int n_theta, int d,
                             const void* basis_t, const void* theta, void* As, void* trels,
                             void* stream);

/* Batched expm of n (m x m) matrices, m in {2,3,4}; libcpab/pytorch/expm.py:11-36 as an op. */
int cpab_b200_expm(int dtype, int m, long n, const void* A, void* E, void* stream);

/*
 * Forward integration.  Replaces cpab_gpu.forward(points, trels, nstepsolver, nc),
 * libcpab/pytorch/transformer_cuda.cpp:17-41 -> transformer_cuda.cu:18-64 -> core/cpab_ops.cu:268-388.
 *   points    [ndim, nP] (broadcast=0) or [n_theta, ndim, nP] (broadcast=1)
 *   trels     [n_theta, nC, ndim, ndim+1]
 *   newpoints [n_theta, ndim, nP]  out
 */
int cpab_b200_forward(int dtype, int flags, int ndim, const int* nc, int nsteps, int n_theta,
                      long nP, int broadcast, const void* points, const void* trels,
                      void* newpoints, void* stream);

/*
 * Reference-layout theta-Jacobian.  Replaces cpab_gpu.backward(points, As, Bs, nstepsolver, nc),
 * libcpab/pytorch/transformer_cuda.cpp:43-70 -> transformer_cuda.cu:66-119 -> core/cpab_ops.cu:390-697.
 *   Bs  [d, nC, ndim, ndim+1] (= basis_t)      jac [d, n_theta, ndim, nP]  out
 * Kept for drop-in completeness and op-level parity; it is d-fold redundant by construction.
 * Training code should use cpab_b200_backward_theta.
 */
int cpab_b200_backward_jacobian(int dtype, int ndim, const int* nc, int nsteps, int n_theta, int d,
                                long nP, int broadcast, const void* points, const void* As,
                                const void* Bs, void* jac, void* stream);

/* Bytes of scratch cpab_b200_backward_theta needs for n_theta thetas x nP points: the per-cell
 * gradient G [n_theta, D], the per-cell RK2 step records [n_theta, nC, 2|8|16], the per-theta
 * certificate bounds, the work counters of the launch and (float32) one bit per trajectory for
 * the certificate's verdict.  The workspace must be 16-byte aligned; it is written by the call
 * only (no state survives it, concurrent calls need distinct workspaces). */
size_t cpab_b200_backward_workspace_bytes(int dtype, int ndim, const int* nc, int n_theta, long nP);

/*
 * dL/dtheta (and optionally dL/dpoints) in one adjoint sweep.  Replaces the pair
 * cpab_gpu.backward (above) + `gradient.mul_(grad).sum(dim=(2,3))`,
 * libcpab/pytorch/transformer.py:187-202, without materialising [d, n_theta, ndim, nP].
 *   points   as in forward          As       [n_theta, nC, ndim, ndim+1]
 *   basis    [D, d] (row-major, as params.basis)
 *   grad_out [n_theta, ndim, nP]    dtheta   [n_theta, d]  out
 *   dpoints  [n_theta, ndim, nP]    out, may be NULL (the reference returns None for it)
 *   flags    0 (certified cell sequences, see CPAB_FLAG_FAST_GRAD) or CPAB_FLAG_FAST_GRAD
 */
int cpab_b200_backward_theta(int dtype, int flags, int ndim, const int* nc, int nsteps,
                             int n_theta, int d, long nP, int broadcast, const void* points,
                             const void* As, const void* basis, const void* grad_out,
                             void* dtheta, void* dpoints, void* workspace, size_t workspace_bytes,
                             void* stream);

/* As cpab_b200_backward_theta, plus diagnostics: redo_count [n_theta] int32 (zeroed by the caller)
 * receives the number of trajectories per theta whose certificate failed and which were
 * re-integrated with the reference's arithmetic.  May be NULL. */
int cpab_b200_backward_theta_diag(int dtype, int flags, int ndim, const int* nc, int nsteps,
                                  int n_theta, int d, long nP, int broadcast, const void* points,
                                  const void* As, const void* basis, const void* grad_out,
                                  void* dtheta, void* dpoints, void* workspace, size_t workspace_bytes,
                                  int* redo_count, void* stream);

/*
 * Test / diagnostic door: the cell index the adjoint's first pass records at every RK2 step
 * (float32).  mode 0: step records + complete search (what CPAB_FLAG_FAST_GRAD integrates),
 * mode 1: step records + certified fast search (failed[t,i] = 1 where the certificate fails),
 * mode 2: the reference's arithmetic (libcpab/core/cpab_ops.cpp:289-366, p-recursion only).
 *   cells  [n_theta, nsteps, nP] int32 out     failed [n_theta, nP] uint8 out, may be NULL
 *   workspace as for cpab_b200_backward_theta (CPAB_F32).
 */
int cpab_b200_rk2_cell_trace(int ndim, const int* nc, int nsteps, int n_theta, long nP, int broadcast,
                             int mode, const void* points, const void* As, void* workspace,
                             size_t workspace_bytes, int* cells, unsigned char* failed, void* stream);

/*
 * Closed-form ("hit-time") integration; OPT-IN extension, no counterpart in the reference (which
 * integrates with nstepsolver fixed steps in every backend, libcpab/cpab.py:77,
 * libcpab/core/cpab_ops.cpp:240-253).  The algorithm BASELINE.json's north_star names (Freifeld et
 * al., TPAMI 2017): analytic in-cell flow to the cell boundary, hit time, cross, repeat to t = 1.
 * 1-D: closed-form hit time (cpab_closed1d.cu).  2-D / 3-D: the in-cell flow and the face
 * functions as Taylor polynomials on sub-steps of ||L|| tau <= 1/2, hit time by bracketing + Newton
 * (cpab_closednd.cu).  Takes the velocity matrices `As` (not Trels) and no step count.
 */
int cpab_b200_forward_closed_form(int dtype, int ndim, const int* nc, int n_theta, long nP,
                                  int broadcast, const void* points, const void* As, void* newpoints,
                                  void* stream);

/* Exact gradient of the above w.r.t. theta (and optionally the points); same workspace as
 * cpab_b200_backward_theta. */
int cpab_b200_backward_theta_closed_form(int dtype, int ndim, const int* nc, int n_theta, int d,
                                         long nP, int broadcast, const void* points, const void* As,
                                         const void* basis, const void* grad_out, void* dtheta,
                                         void* dpoints, void* workspace, size_t workspace_bytes,
                                         void* stream);

/* The same with the forward's output at hand (`newpoints` [n_theta, ndim, nP], may be NULL): in 2-D / 3-D the
 * adjoint walks the reversed field back from x(1), and takes x(1) from here instead of walking forward to it
 * first (a quarter of the kernel's time).  1-D ignores it. */
int cpab_b200_backward_theta_closed_form_from(int dtype, int ndim, const int* nc, int n_theta, int d,
                                              long nP, int broadcast, const void* points, const void* As,
                                              const void* basis, const void* grad_out, const void* newpoints,
                                              void* dtheta, void* dpoints, void* workspace,
                                              size_t workspace_bytes, void* stream);

/*
 * Diagnostic door (2-D / 3-D): the forward walk of cpab_b200_forward_closed_form, which also counts
 * how well the variable-trip-count loop fills its warps.  counts (device, 2 x uint64, zeroed here):
 *   counts[0] = sub-steps executed by lanes, counts[1] = 32 x loop iterations executed by warps;
 * counts[0] / counts[1] is the lane utilisation.  The tuning key "closed_refill" (1: a lane that
 * finishes takes the warp's next point, default; 0: the warp waits for its slowest lane) selects
 * the mitigation that is being measured.
 */
int cpab_b200_closed_form_lane_stats(int dtype, int ndim, const int* nc, int n_theta, long nP, int broadcast,
                                     const void* points, const void* As, void* newpoints,
                                     unsigned long long* counts, void* stream);

/*
 * Linear / bilinear / trilinear sampling.  Replaces interpolate(ndim, data, grid, outsize),
 * libcpab/pytorch/interpolation.py:12-172.
 *   data [N, C, in_size...]   grid [N, ndim, prod(out_size)]   out [N, C, out_size...]
 */
int cpab_b200_interpolate_forward(int dtype, int ndim, int N, int C, const int* in_size,
                                  const int* out_size, const void* data, const void* grid,
                                  void* out, void* stream);

/*
 * Backward of the above (what autograd derives for the reference).  dgrid [N, ndim, nP] and
 * ddata [N, C, in_size...] are outputs; either may be NULL.  ddata is zeroed by the call.
 */
int cpab_b200_interpolate_backward(int dtype, int ndim, int N, int C, const int* in_size,
                                   const int* out_size, const void* data, const void* grid,
                                   const void* grad_out, void* dgrid, void* ddata, void* stream);

/*
 * Fused Cpab.transform_data (libcpab/cpab.py:304-331: uniform_meshgrid -> transform_grid ->
 * interpolate) for a meshgrid `points` [ndim, prod(out_size)] shared by all thetas: the forward
 * kernel samples `data` at the end of every trajectory, the adjoint kernel forms dL/d(grid_t) from
 * the image gradient in its prologue.  Results are identical to cpab_b200_forward followed by
 * cpab_b200_interpolate_forward (same arithmetic); two launches and the d/dgrid round trip less.
 *   data [n_theta, C, in_size...]   grid_t [n_theta, ndim, nP] out (kept for the backward)
 *   out  [n_theta, C, out_size...]
 */
int cpab_b200_transform_data_forward(int dtype, int flags, int ndim, const int* nc, int nsteps,
                                     int n_theta, int C, const int* in_size, const int* out_size,
                                     const void* points, const void* trels, const void* data,
                                     void* grid_t, void* out, void* stream);

/* dL/dtheta of the above from grad_out [n_theta, C, out_size...]; workspace and flags as for
 * cpab_b200_backward_theta.  (dL/ddata, if wanted, is cpab_b200_interpolate_backward's.) */
int cpab_b200_transform_data_backward(int dtype, int flags, int ndim, const int* nc, int nsteps, int n_theta,
                                      int d, int C, const int* in_size, const int* out_size,
                                      const void* points, const void* As, const void* basis,
                                      const void* data, const void* grid_t, const void* grad_out,
                                      void* dtheta, void* workspace, size_t workspace_bytes,
                                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LIBCPAB_B200_H */
