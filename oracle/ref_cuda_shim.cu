// ref_cuda_shim.cu -- TEST / BENCH INFRASTRUCTURE ONLY (see oracle/README.md).
//
// extern "C" doors onto the reference's own CUDA kernels so that bench.py can time them on the
// same B200 as this library (the on-GPU comparator).  The reference's kernel source
// /root/reference/libcpab/core/cpab_ops.cu is compiled UNMODIFIED, where it lies (included below
// through -I$(REF)/libcpab/core; never copied into this repo), for sm_100a into
// oracle/_ref/libcpab_ref_cuda.so by oracle/Makefile.  The launch configurations restate the
// reference's torch binding, which cannot be built without ATen:
//   forward   libcpab/pytorch/transformer_cuda.cu:30-31   grid (ceil(nP/256), n_theta), block (256,1)
//   backward  libcpab/pytorch/transformer_cuda.cu:82-84   block (min(nP,128), min(n_theta,4), 1),
//                                                          grid (ceil(nP/bx), ceil(n_theta/by), d)
// The reference launches on the legacy default stream; the shim takes the caller's stream so that
// CUDA events on torch's current stream bracket the kernels.
#include "cpab_ops.cu"

#include <algorithm>

extern "C" {

int cpab_refcuda_abi(void) { return 1; }

// points [ndim,nP] or [n_theta,ndim,nP]; trels [n_theta,nC,ndim,ndim+1]; out [n_theta,ndim,nP];
// nstep_dev int[1] and nc_dev int[ndim] are DEVICE pointers, as in the reference binding.
int cpab_refcuda_forward(float* out, const float* points, const float* trels, const int* nstep_dev,
                         const int* nc_dev, int ndim, int nP, int n_theta, int broadcast, void* stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    dim3 bc((int)ceil(nP / 256.0), n_theta);
    dim3 tpb(256, 1);
    if (ndim == 1) cpab_cuda_kernel_forward_1D<<<bc, tpb, 0, st>>>(nP, n_theta, out, points, trels, nstep_dev, nc_dev, broadcast);
    if (ndim == 2) cpab_cuda_kernel_forward_2D<<<bc, tpb, 0, st>>>(nP, n_theta, out, points, trels, nstep_dev, nc_dev, broadcast);
    if (ndim == 3) cpab_cuda_kernel_forward_3D<<<bc, tpb, 0, st>>>(nP, n_theta, out, points, trels, nstep_dev, nc_dev, broadcast);
    return (int)cudaPeekAtLastError();
}

// grad [d,n_theta,ndim,nP] must be zeroed by the caller (the binding allocates it with torch::zeros,
// libcpab/pytorch/transformer_cuda.cpp:65; the kernels read q back from it every step).
int cpab_refcuda_backward(float* grad, const float* points, const float* As, const float* Bs,
                          const int* nstep_dev, const int* nc_dev, int ndim, int nP, int n_theta, int d,
                          int nC, int broadcast, void* stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    dim3 tpb(std::min(nP, 128), std::min(n_theta, 4), std::min(d, 1));
    dim3 bc((nP + tpb.x - 1) / tpb.x, (n_theta + tpb.y - 1) / tpb.y, (d + tpb.z - 1) / tpb.z);
    dim3 vtc(nP, n_theta, d);
    if (ndim == 1) cpab_cuda_kernel_backward_1D<<<bc, tpb, 0, st>>>(vtc, n_theta, d, nP, nC, grad, points, As, Bs, nstep_dev, nc_dev, broadcast);
    if (ndim == 2) cpab_cuda_kernel_backward_2D<<<bc, tpb, 0, st>>>(vtc, n_theta, d, nP, nC, grad, points, As, Bs, nstep_dev, nc_dev, broadcast);
    if (ndim == 3) cpab_cuda_kernel_backward_3D<<<bc, tpb, 0, st>>>(vtc, n_theta, d, nP, nC, grad, points, As, Bs, nstep_dev, nc_dev, broadcast);
    return (int)cudaPeekAtLastError();
}

}  // extern "C"
