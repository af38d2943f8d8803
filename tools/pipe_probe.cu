// pipe_probe.cu -- issue/pipe throughput of the instructions the strict RK2 trajectory needs
// (development tool: F2F.F64.F32, F2F.F32.F64, DMUL, DADD, DFMA against FFMA on one B200).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/pipe_probe tools/pipe_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP> __global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b, double da, double db)
{
    float x[8];
    double y[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { x[u] = threadIdx.x + u; y[u] = threadIdx.x + u + 0.5; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (OP == 0) x[u] = fmaf(x[u], a, b);
                if (OP == 1) y[u] = fma(y[u], da, db);
                if (OP == 2) y[u] = __dadd_rn(y[u], db);
                if (OP == 3) y[u] = __dmul_rn(y[u], da);
                if (OP == 4) { y[u] = __dadd_rn((double)x[u], db); x[u] = __int_as_float(__double2loint(y[u]) | 0x3f000000); }   // cvt up + dadd + 2 int ops
                if (OP == 5) { x[u] = (float)y[u]; y[u] = __hiloint2double(__float_as_int(x[u]) | 0x3ff00000, __double2loint(y[u])); }  // cvt down + int op
                if (OP == 6) {     // the strict update: p = (float)((double)p + (double)v * h)
                    x[u] = (float)__dadd_rn((double)x[u], __dmul_rn((double)a, da));
                }
                if (OP == 7) {     // same with a chain through v so nothing hoists
                    const float v = __fadd_rn(__fmul_rn(x[u], a), b);
                    x[u] = (float)__dadd_rn((double)x[u], __dmul_rn((double)v, da));
                }
            }
        }
    }
    float s = 0; double t = 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) { s += x[u]; t += y[u]; }
    if (s == 123.456f || t == 123.456) out[0] = s + (float)t;
}

template <int OP> void run(const char* name, int ops_per_inner)
{
    float* out; cudaMalloc(&out, 4);
    const int blocks = 148 * 8, iters = 2048;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        k<OP><<<blocks, 256>>>(out, iters, 0.999f, 0.001f, 0.999, 0.001);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double inner = (double)blocks * 256 * iters * 32;
    // per SM per clock at 1.965 GHz
    const double per_clk_sm = inner / (best * 1e-3) / 148 / 1.965e9;
    printf("%-28s %8.3f ms  %7.2f inner-iterations/clk/SM  (x%d instr = %7.2f lane-instr/clk/SM)\n", name, best, per_clk_sm,
           ops_per_inner, per_clk_sm * ops_per_inner);
    cudaFree(out);
}

int main()
{
    run<0>("FFMA", 1);
    run<1>("DFMA", 1);
    run<2>("DADD", 1);
    run<3>("DMUL", 1);
    run<4>("F2F.F64.F32 + DADD + LOP", 3);
    run<5>("F2F.F32.F64 + LOP", 2);
    run<6>("strict update (hoistable v)", 4);
    run<7>("strict update + fmul,fadd", 7);
    return 0;
}
