// cpab_common.cuh -- shared declarations of the libcpab_b200 CUDA library (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "cpab_cell.cuh"

namespace cpab {

// ---- status codes (mirrored in include/libcpab_b200.h) ------------------------------------------
enum Status : int {
    kOk = 0,
    kErrArgument = -1,      // bad shape / null pointer / unsupported ndim
    kErrCuda = -2,          // a CUDA runtime call or launch failed; see cpab_b200_last_error()
    kErrUnsupported = -3,   // valid request this build cannot serve (e.g. tessellation too large)
    kErrWorkspace = -4,     // caller's workspace is smaller than cpab_b200_backward_workspace_bytes
};

enum DType : int { kF32 = 0, kF64 = 1 };

enum Flags : int {
    kFlagFastMath = 1,      // forward: allow FMA contraction (not bit-exact with the CPU reference)
    kFlagFastGrad = 2,      // backward: skip the cell-sequence certificate (cpab_adjoint.cuh); a trajectory
                            // that lands within rounding of a cell face may then follow a neighbouring cell
};

#define CPAB_STR2(x) #x
#define CPAB_STR(x) CPAB_STR2(x)

void set_error(const char* fmt, ...);

// ---- instrumentation (cpab_abi.cu) -----------------------------------------------------------------
// Every kernel launch of this library goes through count_launch(); bench.py reports the total as
// `gpu_launches`.  When profiling is switched on (cpab_b200_profile_enable), the dominant kernels
// are bracketed by CUDA events on their own stream and the elapsed time is accumulated per slot.
enum ProfSlot : int { kProfForward = 0, kProfBackward = 1, kProfInterpFwd = 2, kProfInterpBwd = 3,
                      kProfThetaToTrels = 4, kProfEpilogue = 5, kProfBackwardRedo = 6, kProfSlots = 7 };
void count_launch(int n = 1);
bool prof_begin(int slot, cudaStream_t st);   // true if an event was recorded
void prof_end(int slot, cudaStream_t st);

int set_tuning(const char* key, int value);
void set_interp_variant(int v);   // cpab_interp.cu
void set_interp_max_ctas(int v);
const char* get_error();

#define CPAB_CUDA_OK(expr)                                                                   \
    do {                                                                                     \
        cudaError_t err__ = (expr);                                                          \
        if (err__ != cudaSuccess) {                                                          \
            ::cpab::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__),     \
                              __FILE__, __LINE__);                                           \
            return ::cpab::kErrCuda;                                                         \
        }                                                                                    \
    } while (0)

template <int NDIM> struct Dim {
    static constexpr int kPpc = NDIM * (NDIM + 1);   // affine parameters per simplex: 2 / 6 / 12
};

// ---- launchers implemented in the .cu files --------------------------------------------------------
// cpab_integrate.cu
int launch_findcellidx(int dtype, const Geom& g, const void* points, long nP, int* out,
                       cudaStream_t st);
int launch_forward(int dtype, int flags, const Geom& g, int nsteps, int n_theta, long nP,
                   int broadcast, const void* points, const void* trels, void* out,
                   cudaStream_t st);
int launch_jacobian(int dtype, const Geom& g, int nsteps, int n_theta, int d, long nP,
                    int broadcast, const void* points, const void* As, const void* Bs, void* jac,
                    cudaStream_t st);
size_t backward_g_bytes(int dtype, const Geom& g, int n_theta);           // G [n_theta, D]
size_t backward_workspace_bytes(int dtype, const Geom& g, int n_theta, long nP);   // G + RK2 step table + certificate scratch
int launch_backward(int dtype, int flags, const Geom& g, int nsteps, int n_theta, int d, long nP,
                    int broadcast, const void* points, const void* As, const void* basis,
                    const void* grad_out, void* dtheta, void* dpoints, void* workspace,
                    size_t workspace_bytes, int* flagged, cudaStream_t st);
int launch_rk2_trace(const Geom& g, int nsteps, int n_theta, long nP, int broadcast, int mode,
                     const void* points, const void* As, void* workspace, size_t workspace_bytes,
                     int* cells, unsigned char* failed, cudaStream_t st);
int launch_transform_data_forward(int dtype, int flags, const Geom& g, int nsteps, int n_theta, int C,
                                  const int* in_size, const int* out_size, const void* points,
                                  const void* trels, const void* data, void* grid_t, void* img,
                                  cudaStream_t st);
int launch_transform_data_backward(int dtype, int flags, const Geom& g, int nsteps, int n_theta, int d, int C,
                                   const int* in_size, const int* out_size, const void* points,
                                   const void* As, const void* basis, const void* data,
                                   const void* grid_t, const void* gimg, void* dtheta, void* workspace,
                                   size_t workspace_bytes, cudaStream_t st);
int launch_grad_epilogue(int dtype, const void* G, const void* basis, void* dtheta, int n_theta, int D,
                         int d, cudaStream_t st);
// cpab_closed1d.cu
int launch_closed1d_forward(int dtype, const Geom& g, int n_theta, long nP, int broadcast,
                            const void* points, const void* As, void* out, cudaStream_t st);
int launch_closed1d_backward(int dtype, const Geom& g, int n_theta, long nP, int broadcast,
                             const void* points, const void* As, const void* gout, void* G,
                             void* dpoints, cudaStream_t st);
// cpab_closednd.cu
int launch_closednd_forward(int dtype, const Geom& g, int n_theta, long nP, int broadcast, const void* points,
                            const void* As, void* out, unsigned long long* stats, cudaStream_t st);
int launch_closednd_backward(int dtype, const Geom& g, int n_theta, long nP, int broadcast, const void* points,
                             const void* As, const void* gout, const void* newpoints, void* G, void* dpoints,
                             cudaStream_t st);
void set_closed_refill(int v);
void set_closed_stage(int v);
// cpab_expm.cu
int launch_theta_to_trels(int dtype, const Geom& g, int nsteps, int n_theta, int d,
                          const void* basis_t, const void* theta, void* As, void* trels,
                          cudaStream_t st);
int launch_expm(int dtype, int m, long n, const void* A, void* E, cudaStream_t st);
// cpab_interp.cu
int launch_interp_forward(int dtype, int ndim, int N, int C, const int* in_size,
                          const int* out_size, const void* data, const void* grid, void* out,
                          cudaStream_t st);
int launch_interp_backward(int dtype, int ndim, int N, int C, const int* in_size,
                           const int* out_size, const void* data, const void* grid,
                           const void* grad_out, void* dgrid, void* ddata, cudaStream_t st);

// cpab_probe.cu
int launch_fma_probe(int blocks, int iters, float* out, cudaStream_t st);

// Largest dynamic shared memory a kernel of this library asks for (B200: 227 KB per CTA).
constexpr size_t kMaxSmemBytes = 227 * 1024;

}  // namespace cpab
