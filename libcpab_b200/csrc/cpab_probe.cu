// cpab_probe.cu -- FP32 FMA throughput probe: the measured denominator of the FP32 roofline.
//
// MEASURED_PEAKS.json carries an HBM copy bandwidth and a bf16 tensor peak but no FP32 CUDA-core
// figure, and the integration kernels are FP32-pipe work.  bench.py times this kernel (8
// independent FFMA chains per thread, every SM filled to 2048 threads) and uses the result as
// the FP32 peak "of measured"; nominal is 148 SMs x 128 lanes x 2 flop x 1.965 GHz = 74.4 TFLOP/s.
#include "cpab_common.cuh"

namespace cpab {

__global__ void __launch_bounds__(256) k_fma_probe(float* out, int iters, float a, float b)
{
    float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f;
    float x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    const float s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456f) out[0] = s;      // never true; keeps the chains alive
}

// flops issued = blocks * 256 threads * iters * 64 FMAs * 2
int launch_fma_probe(int blocks, int iters, float* out, cudaStream_t st)
{
    k_fma_probe<<<blocks, 256, 0, st>>>(out, iters, 0.999f, 0.001f);
    count_launch();
    CPAB_CUDA_OK(cudaGetLastError());
    return kOk;
}

}  // namespace cpab
