// cpab_abi.cu -- extern "C" surface of libcpab_b200.so (declared in include/libcpab_b200.h).
// Argument validation and error reporting live here; kernels live in the other .cu files.
#include <stdarg.h>
#include <stdio.h>

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/libcpab_b200.h"
#include "cpab_common.cuh"

namespace cpab {

static thread_local std::string t_error;

void set_error(const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    t_error = buf;
}
const char* get_error() { return t_error.c_str(); }

int set_tuning(const char* key, int value);   // cpab_integrate.cu

// ---- instrumentation ---------------------------------------------------------------------------------
static std::atomic<long long> g_launches{0};
static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;
struct ProfPair { cudaEvent_t a, b; int slot; };
static std::vector<ProfPair> g_prof_pending;     // recorded, not yet read back
static double g_prof_ms[kProfSlots];
static long long g_prof_n[kProfSlots];

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

bool prof_begin(int slot, cudaStream_t st)
{
    if (!g_prof_on.load(std::memory_order_relaxed)) return false;
    ProfPair p;
    p.slot = slot;
    if (cudaEventCreate(&p.a) != cudaSuccess || cudaEventCreate(&p.b) != cudaSuccess) return false;
    cudaEventRecord(p.a, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_pending.push_back(p);
    return true;
}

void prof_end(int slot, cudaStream_t st)
{
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto it = g_prof_pending.rbegin(); it != g_prof_pending.rend(); ++it)
        if (it->slot == slot) { cudaEventRecord(it->b, st); return; }
}

static void prof_collect()
{
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto& p : g_prof_pending) {
        float ms = 0.f;
        if (cudaEventSynchronize(p.b) == cudaSuccess && cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
            g_prof_ms[p.slot] += ms;
            g_prof_n[p.slot] += 1;
        }
        cudaEventDestroy(p.a);
        cudaEventDestroy(p.b);
    }
    g_prof_pending.clear();
}

static bool check_geom(int dtype, int ndim, const int* nc)
{
    if (dtype != kF32 && dtype != kF64) { set_error("dtype must be CPAB_F32 or CPAB_F64, got %d", dtype); return false; }
    if (ndim < 1 || ndim > 3) { set_error("ndim must be 1, 2 or 3, got %d", ndim); return false; }
    if (nc == nullptr) { set_error("nc is NULL"); return false; }
    long cells = ndim == 1 ? 1 : (ndim == 2 ? 4 : 5);
    for (int j = 0; j < ndim; ++j) {
        if (nc[j] <= 0) { set_error("nc[%d] = %d must be positive", j, nc[j]); return false; }
        if (nc[j] > (1 << 20)) { set_error("nc[%d] = %d exceeds 2^20", j, nc[j]); return false; }
        cells *= nc[j];
    }
    if (cells * ndim * (ndim + 1) > (1L << 30)) { set_error("tessellation too large (%ld cells)", cells); return false; }
    return true;
}

#define REQUIRE(cond, ...)                                  \
    do {                                                    \
        if (!(cond)) { set_error(__VA_ARGS__); return kErrArgument; } \
    } while (0)

}  // namespace cpab

using namespace cpab;

extern "C" {

int cpab_b200_abi_version(void) { return 2; }

const char* cpab_b200_last_error(void) { return get_error(); }

const char* cpab_b200_build_info(void)
{
    return "libcpab_b200;arch=sm_100a;cuda=" CPAB_STR(CUDART_VERSION) ";abi=2";
}

int cpab_b200_set_tuning(const char* key, int value)
{
    if (key == nullptr) { set_error("key is NULL"); return kErrArgument; }
    return set_tuning(key, value);
}

long long cpab_b200_launch_count(void) { return g_launches.load(); }

int cpab_b200_profile_enable(int on)
{
    prof_collect();
    for (int i = 0; i < kProfSlots; ++i) { g_prof_ms[i] = 0.0; g_prof_n[i] = 0; }
    g_prof_on.store(on ? 1 : 0);
    return kOk;
}

int cpab_b200_profile_read(int slot, double* total_ms, long long* launches)
{
    REQUIRE(slot >= 0 && slot < kProfSlots && total_ms && launches, "profile_read: bad arguments");
    prof_collect();
    *total_ms = g_prof_ms[slot];
    *launches = g_prof_n[slot];
    return kOk;
}

int cpab_b200_fp32_fma_probe(int blocks, int iters, void* out, void* stream)
{
    REQUIRE(blocks > 0 && iters > 0 && out, "fp32_fma_probe: bad arguments");
    return launch_fma_probe(blocks, iters, (float*)out, (cudaStream_t)stream);
}

int cpab_b200_findcellidx(int dtype, int ndim, const int* nc, const void* points, long nP,
                          int* out_idx, void* stream)
{
    if (!check_geom(dtype, ndim, nc)) return kErrArgument;
    REQUIRE(nP >= 0, "nP = %ld is negative", nP);
    REQUIRE(nP == 0 || (points && out_idx), "points/out_idx is NULL");
    return launch_findcellidx(dtype, make_geom(ndim, nc), points, nP, out_idx, (cudaStream_t)stream);
}

int cpab_b200_theta_to_trels(int dtype, int ndim, const int* nc, int nsteps, int n_theta, int d,
                             const void* basis_t, const void* theta, void* As, void* trels,
                             void* stream)
{
    if (!check_geom(dtype, ndim, nc)) return kErrArgument;
    REQUIRE(nsteps > 0, "nstepsolver = %d must be positive", nsteps);
    REQUIRE(n_theta >= 0 && d >= 0, "negative size (n_theta=%d, d=%d)", n_theta, d);
    REQUIRE(n_theta == 0 || (theta && As && trels && (d == 0 || basis_t)), "NULL pointer argument");
    return launch_theta_to_trels(dtype, make_geom(ndim, nc), nsteps, n_theta, d, basis_t, theta, As,
                                 trels, (cudaStream_t)stream);
}

int cpab_b200_expm(int dtype, int m, long n, const void* A, void* E, void* stream)
{
    REQUIRE(dtype == kF32 || dtype == kF64, "bad dtype %d", dtype);
    REQUIRE(n >= 0, "n = %ld is negative", n);
    REQUIRE(n == 0 || (A && E), "NULL pointer argument");
    return launch_expm(dtype, m, n, A, E, (cudaStream_t)stream);
}

int cpab_b200_forward(int dtype, int flags, int ndim, const int* nc, int nsteps, int n_theta,
                      long nP, int broadcast, const void* points, const void* trels,
                      void* newpoints, void* stream)
{
    if (!check_geom(dtype, ndim, nc)) return kErrArgument;
    REQUIRE(nsteps > 0, "nstepsolver = %d must be positive", nsteps);
    REQUIRE(n_theta >= 0 && nP >= 0, "negative size (n_theta=%d, nP=%ld)", n_theta, nP);
    REQUIRE(broadcast == 0 || broadcast == 1, "broadcast must be 0 or 1");
    REQUIRE(n_theta == 0 || nP == 0 || (points && trels && newpoints), "NULL pointer argument");
    return launch_forward(dtype, flags, make_geom(ndim, nc), nsteps, n_theta, nP, broadcast, points,
                          trels, newpoints, (cudaStream_t)stream);
}

int cpab_b200_backward_jacobian(int dtype, int ndim, const int* nc, int nsteps, int n_theta, int d,
                                long nP, int broadcast, const void* points, const void* As,
                                const void* Bs, void* jac, void* stream)
{
    if (!check_geom(dtype, ndim, nc)) return kErrArgument;
    REQUIRE(nsteps > 0, "nstepsolver = %d must be positive", nsteps);
    REQUIRE(n_theta >= 0 && nP >= 0 && d >= 0, "negative size");
    REQUIRE(broadcast == 0 || broadcast == 1, "broadcast must be 0 or 1");
    REQUIRE(n_theta == 0 || nP == 0 || d == 0 || (points && As && Bs && jac), "NULL pointer argument");
    return launch_jacobian(dtype, make_geom(ndim, nc), nsteps, n_theta, d, nP, broadcast, points, As,
                           Bs, jac, (cudaStream_t)stream);
}

size_t cpab_b200_backward_workspace_bytes(int dtype, int ndim, const int* nc, int n_theta, long nP)
{
    if (!check_geom(dtype, ndim, nc) || n_theta < 0 || nP < 0) return 0;
    return backward_workspace_bytes(dtype, make_geom(ndim, nc), n_theta, nP);
}

int cpab_b200_backward_theta(int dtype, int flags, int ndim, const int* nc, int nsteps,
                             int n_theta, int d, long nP, int broadcast, const void* points,
                             const void* As, const void* basis, const void* grad_out,
                             void* dtheta, void* dpoints, void* workspace, size_t workspace_bytes,
                             void* stream)
{
    if (!check_geom(dtype, ndim, nc)) return kErrArgument;
    REQUIRE(nsteps > 0, "nstepsolver = %d must be positive", nsteps);
    REQUIRE(n_theta >= 0 && nP >= 0 && d >= 0, "negative size");
    REQUIRE(broadcast == 0 || broadcast == 1, "broadcast must be 0 or 1");
    REQUIRE(n_theta == 0 || d == 0 || (As && basis && dtheta && workspace), "NULL pointer argument");
    REQUIRE(n_theta == 0 || nP == 0 || (points && grad_out), "NULL pointer argument");
    return launch_backward(dtype, flags, make_geom(ndim, nc), nsteps, n_theta, d, nP, broadcast,
                           points, As, basis, grad_out, dtheta, dpoints, workspace,
                           workspace_bytes, nullptr, (cudaStream_t)stream);
}

int cpab_b200_backward_theta_diag(int dtype, int flags, int ndim, const int* nc, int nsteps,
                                  int n_theta, int d, long nP, int broadcast, const void* points,
                                  const void* As, const void* basis, const void* grad_out,
                                  void* dtheta, void* dpoints, void* workspace, size_t workspace_bytes,
                                  int* redo_count, void* stream)
{
    if (!check_geom(dtype, ndim, nc)) return kErrArgument;
    REQUIRE(nsteps > 0, "nstepsolver = %d must be positive", nsteps);
    REQUIRE(n_theta >= 0 && nP >= 0 && d >= 0, "negative size");
    REQUIRE(broadcast == 0 || broadcast == 1, "broadcast must be 0 or 1");
    REQUIRE(n_theta == 0 || d == 0 || (As && basis && dtheta && workspace), "NULL pointer argument");
    REQUIRE(n_theta == 0 || nP == 0 || (points && grad_out), "NULL pointer argument");
    return launch_backward(dtype, flags, make_geom(ndim, nc), nsteps, n_theta, d, nP, broadcast,
                           points, As, basis, grad_out, dtheta, dpoints, workspace,
                           workspace_bytes, redo_count, (cudaStream_t)stream);
}

int cpab_b200_rk2_cell_trace(int ndim, const int* nc, int nsteps, int n_theta, long nP, int broadcast,
                             int mode, const void* points, const void* As, void* workspace,
                             size_t workspace_bytes, int* cells, unsigned char* failed, void* stream)
{
    if (!check_geom(kF32, ndim, nc)) return kErrArgument;
    REQUIRE(nsteps > 0, "nstepsolver = %d must be positive", nsteps);
    REQUIRE(n_theta >= 0 && nP >= 0, "negative size");
    REQUIRE(mode >= 0 && mode <= 2, "mode must be 0, 1 or 2");
    REQUIRE(broadcast == 0 || broadcast == 1, "broadcast must be 0 or 1");
    REQUIRE(n_theta == 0 || nP == 0 || (points && As && workspace && cells), "NULL pointer argument");
    return launch_rk2_trace(make_geom(ndim, nc), nsteps, n_theta, nP, broadcast, mode, points, As,
                            workspace, workspace_bytes, cells, failed, (cudaStream_t)stream);
}

int cpab_b200_forward_closed_form(int dtype, int ndim, const int* nc, int n_theta, long nP,
                                  int broadcast, const void* points, const void* As, void* newpoints,
                                  void* stream)
{
    if (!check_geom(dtype, ndim, nc)) return kErrArgument;
    REQUIRE(n_theta >= 0 && nP >= 0, "negative size");
    REQUIRE(broadcast == 0 || broadcast == 1, "broadcast must be 0 or 1");
    REQUIRE(n_theta == 0 || nP == 0 || (points && As && newpoints), "NULL pointer argument");
    const Geom g = make_geom(ndim, nc);
    if (ndim == 1) return launch_closed1d_forward(dtype, g, n_theta, nP, broadcast, points, As, newpoints, (cudaStream_t)stream);
    return launch_closednd_forward(dtype, g, n_theta, nP, broadcast, points, As, newpoints, nullptr, (cudaStream_t)stream);
}

int cpab_b200_backward_theta_closed_form(int dtype, int ndim, const int* nc, int n_theta, int d,
                                         long nP, int broadcast, const void* points, const void* As,
                                         const void* basis, const void* grad_out, void* dtheta,
                                         void* dpoints, void* workspace, size_t workspace_bytes,
                                         void* stream)
{
    return cpab_b200_backward_theta_closed_form_from(dtype, ndim, nc, n_theta, d, nP, broadcast, points, As, basis,
                                                     grad_out, nullptr, dtheta, dpoints, workspace, workspace_bytes, stream);
}

int cpab_b200_backward_theta_closed_form_from(int dtype, int ndim, const int* nc, int n_theta, int d,
                                              long nP, int broadcast, const void* points, const void* As,
                                              const void* basis, const void* grad_out, const void* newpoints,
                                              void* dtheta, void* dpoints, void* workspace,
                                              size_t workspace_bytes, void* stream)
{
    if (!check_geom(dtype, ndim, nc)) return kErrArgument;
    REQUIRE(n_theta >= 0 && nP >= 0 && d >= 0, "negative size");
    REQUIRE(broadcast == 0 || broadcast == 1, "broadcast must be 0 or 1");
    REQUIRE(n_theta == 0 || d == 0 || (As && basis && dtheta && workspace), "NULL pointer argument");
    REQUIRE(n_theta == 0 || nP == 0 || (points && grad_out), "NULL pointer argument");
    const Geom g = make_geom(ndim, nc);
    const size_t need = backward_workspace_bytes(dtype, g, n_theta, nP);
    if (workspace_bytes < need) { set_error("backward: workspace has %zu bytes, needs %zu", workspace_bytes, need); return kErrWorkspace; }
    if (n_theta == 0 || d == 0) return kOk;
    cudaStream_t st = (cudaStream_t)stream;
    CPAB_CUDA_OK(cudaMemsetAsync(workspace, 0, backward_g_bytes(dtype, g, n_theta), st));
    int rc = ndim == 1 ? launch_closed1d_backward(dtype, g, n_theta, nP, broadcast, points, As, grad_out, workspace, dpoints, st)
                       : launch_closednd_backward(dtype, g, n_theta, nP, broadcast, points, As, grad_out, newpoints, workspace, dpoints, st);
    if (rc != kOk) return rc;
    return launch_grad_epilogue(dtype, workspace, basis, dtheta, n_theta, g.n_cells * ndim * (ndim + 1), d, st);
}

int cpab_b200_closed_form_lane_stats(int dtype, int ndim, const int* nc, int n_theta, long nP, int broadcast,
                                     const void* points, const void* As, void* newpoints,
                                     unsigned long long* counts, void* stream)
{
    if (!check_geom(dtype, ndim, nc)) return kErrArgument;
    REQUIRE(ndim == 2 || ndim == 3, "lane statistics exist for the 2-D / 3-D walk, got ndim = %d", ndim);
    REQUIRE(n_theta >= 0 && nP >= 0, "negative size");
    REQUIRE(broadcast == 0 || broadcast == 1, "broadcast must be 0 or 1");
    REQUIRE(counts != nullptr, "NULL pointer argument");
    REQUIRE(n_theta == 0 || nP == 0 || (points && As && newpoints), "NULL pointer argument");
    CPAB_CUDA_OK(cudaMemsetAsync(counts, 0, 2 * sizeof(unsigned long long), (cudaStream_t)stream));
    return launch_closednd_forward(dtype, make_geom(ndim, nc), n_theta, nP, broadcast, points, As, newpoints,
                                   counts, (cudaStream_t)stream);
}

static int check_interp(int dtype, int ndim, int N, int C, const int* in_size, const int* out_size)
{
    REQUIRE(dtype == kF32 || dtype == kF64, "bad dtype %d", dtype);
    REQUIRE(ndim >= 1 && ndim <= 3, "ndim must be 1, 2 or 3, got %d", ndim);
    REQUIRE(N >= 0 && C >= 0, "negative batch/channel count");
    REQUIRE(in_size && out_size, "in_size/out_size is NULL");
    for (int j = 0; j < ndim; ++j) {
        REQUIRE(in_size[j] > 0, "in_size[%d] = %d must be positive", j, in_size[j]);
        REQUIRE(out_size[j] >= 0, "out_size[%d] = %d is negative", j, out_size[j]);
    }
    return kOk;
}

int cpab_b200_interpolate_forward(int dtype, int ndim, int N, int C, const int* in_size,
                                  const int* out_size, const void* data, const void* grid,
                                  void* out, void* stream)
{
    const int rc = check_interp(dtype, ndim, N, C, in_size, out_size);
    if (rc != kOk) return rc;
    REQUIRE(N == 0 || C == 0 || (data && grid && out), "NULL pointer argument");
    return launch_interp_forward(dtype, ndim, N, C, in_size, out_size, data, grid, out,
                                 (cudaStream_t)stream);
}

int cpab_b200_transform_data_forward(int dtype, int flags, int ndim, const int* nc, int nsteps,
                                     int n_theta, int C, const int* in_size, const int* out_size,
                                     const void* points, const void* trels, const void* data,
                                     void* grid_t, void* out, void* stream)
{
    if (!check_geom(dtype, ndim, nc)) return kErrArgument;
    const int rc = check_interp(dtype, ndim, n_theta, C, in_size, out_size);
    if (rc != kOk) return rc;
    REQUIRE(nsteps > 0, "nstepsolver = %d must be positive", nsteps);
    REQUIRE(n_theta == 0 || C == 0 || (points && trels && data && grid_t && out), "NULL pointer argument");
    if (C == 0) return kOk;
    return launch_transform_data_forward(dtype, flags, make_geom(ndim, nc), nsteps, n_theta, C, in_size,
                                         out_size, points, trels, data, grid_t, out, (cudaStream_t)stream);
}

int cpab_b200_transform_data_backward(int dtype, int flags, int ndim, const int* nc, int nsteps, int n_theta,
                                      int d, int C, const int* in_size, const int* out_size,
                                      const void* points, const void* As, const void* basis,
                                      const void* data, const void* grid_t, const void* grad_out,
                                      void* dtheta, void* workspace, size_t workspace_bytes,
                                      void* stream)
{
    if (!check_geom(dtype, ndim, nc)) return kErrArgument;
    const int rc = check_interp(dtype, ndim, n_theta, C, in_size, out_size);
    if (rc != kOk) return rc;
    REQUIRE(nsteps > 0, "nstepsolver = %d must be positive", nsteps);
    REQUIRE(d >= 0, "negative size");
    REQUIRE(n_theta == 0 || d == 0 || (points && As && basis && data && grid_t && grad_out && dtheta && workspace),
            "NULL pointer argument");
    return launch_transform_data_backward(dtype, flags, make_geom(ndim, nc), nsteps, n_theta, d, C, in_size,
                                          out_size, points, As, basis, data, grid_t, grad_out, dtheta,
                                          workspace, workspace_bytes, (cudaStream_t)stream);
}

int cpab_b200_interpolate_backward(int dtype, int ndim, int N, int C, const int* in_size,
                                   const int* out_size, const void* data, const void* grid,
                                   const void* grad_out, void* dgrid, void* ddata, void* stream)
{
    const int rc = check_interp(dtype, ndim, N, C, in_size, out_size);
    if (rc != kOk) return rc;
    REQUIRE(N == 0 || C == 0 || (data && grid && grad_out), "NULL pointer argument");
    return launch_interp_backward(dtype, ndim, N, C, in_size, out_size, data, grid, grad_out, dgrid,
                                  ddata, (cudaStream_t)stream);
}

}  // extern "C"
