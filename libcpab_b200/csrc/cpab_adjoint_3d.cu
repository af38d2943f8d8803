// cpab_adjoint_3d.cu -- 3-D instantiation of the adjoint gradient kernels (cpab_adjoint.cuh).
#include "cpab_adjoint.cuh"

namespace cpab {

#if !defined(CPAB_FAST_BUILD) || CPAB_FAST_DIM == 3
int backward_dim_3(int dtype, int flags, const Geom& g, int nsteps, int n_theta, int d, long nP,
                   int broadcast, const void* points, const void* As, const void* basis,
                   const void* gout, void* dtheta, void* dpoints, void* ws, int* flagged,
                   cudaStream_t st, const SampleArgs* sa)
{
    return backward_dim_impl<3>(dtype, flags, g, nsteps, n_theta, d, nP, broadcast, points, As, basis,
                                gout, dtheta, dpoints, ws, flagged, st, sa);
}

int rk2_trace_dim_3(const Geom& g, int nsteps, int n_theta, long nP, int broadcast, int mode,
                    const void* points, const void* As, void* ws, int* cells, unsigned char* failed,
                    cudaStream_t st)
{
    return rk2_trace_dim_impl<3>(g, nsteps, n_theta, nP, broadcast, mode, points, As, ws, cells, failed, st);
}

size_t backward_workspace_bytes_3(size_t elt, const Geom& g, int n_theta, long nP)
{
    return backward_layout<3>(elt, g, n_theta, nP).total;
}
#endif

}  // namespace cpab
