// FP64 roofline denominators measured on the device in use (MEASURED_PEAKS.json has no FP64 figure):
//   kind 0: dependent-chain-free DFMA throughput on the CUDA cores (the pipe the QP kernel runs on)
//   kind 1: DMMA (mma.sync.m8n8k4.f64) throughput, for the "tensor cores or not" decision in DESIGN.md
#include <cuda_runtime.h>

#include "vsmpc.h"

namespace
{
constexpr int ACC = 8;

__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b)
{
    double acc[ACC];
#pragma unroll
    for (int i = 0; i < ACC; ++i)
        acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < ACC; ++i)
                acc[i] = fma(acc[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < ACC; ++i)
        s += acc[i];
    if (s == 123.456)
        out[0] = s;
}

__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters, double a, double b)
{
    double c0[4], c1[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
        c0[i] = threadIdx.x * 1e-9 + i;
        c1[i] = i;
    }
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(c0[i]), "+d"(c1[i])
                             : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        s += c0[i] + c1[i];
    if (s == 123.456)
        out[0] = s;
}
} // namespace

extern "C" int vsmpc_microbench_fp64(int device, int kind, double* tflops)
{
    if (!tflops)
        return VSMPC_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess)
        return VSMPC_ERR_CUDA;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    double* d = nullptr;
    cudaMalloc(&d, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int grid = prop.multiProcessorCount * 8, block = 256, iters = 4000;
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep)
    {
        cudaEventRecord(e0);
        if (kind == 0)
            dfma_kernel<<<grid, block>>>(d, iters, 0.999999, 1e-7);
        else
            dmma_kernel<<<grid, block>>>(d, iters, 0.999999, 1e-7);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess)
            break;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        // flops: DFMA = 2 per thread-op ; DMMA m8n8k4 = 2*8*8*4 per warp-op
        const double ops = kind == 0 ? (double)grid * block * iters * 8.0 * ACC * 2.0
                                     : (double)grid * (block / 32) * iters * 8.0 * 4.0 * 512.0;
        const double tf = ops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best)
            best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    const cudaError_t e = cudaGetLastError();
    *tflops = best;
    return e == cudaSuccess ? VSMPC_OK : VSMPC_ERR_CUDA;
}
