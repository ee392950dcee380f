// Development probe: FP64 FMA throughput of one SM as a function of the number of resident warps per
// scheduler and of independent dependency chains per warp (ILP). Answers "how much ILP does a kernel that
// can only hold 3 warps per scheduler (168 registers) need to fill the FP64 pipe?".
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_ilp_probe dfma_ilp_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k(double* out, int iters, double a, double b)
{
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
double run(int warps_per_sm, int sms, double* d)
{
    const int iters = 4096 / ILP * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<ILP><<<sms, warps_per_sm * 32>>>(d, 16, 0.999, 1e-3);
    cudaEventRecord(e0);
    k<ILP><<<sms, warps_per_sm * 32>>>(d, iters, 0.999, 1e-3);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fma = (double)sms * warps_per_sm * 32 * iters * 16.0 * ILP;
    return 2.0 * fma / (ms * 1e-3) / 1e12;
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* d; cudaMalloc(&d, sizeof(double) * p.multiProcessorCount * 1024);
    printf("TFLOP/s (FP64 FMA), %d SMs; rows: warps per SM (per scheduler), cols: independent chains per warp\n", p.multiProcessorCount);
    printf("%12s %8s %8s %8s %8s %8s\n", "warps/SM", "ILP1", "ILP2", "ILP4", "ILP8", "ILP16");
    for (int w : {4, 8, 12, 16, 32}) {
        printf("%6d (%2d) ", w, w / 4);
        printf("%8.2f ", run<1>(w, p.multiProcessorCount, d));
        printf("%8.2f ", run<2>(w, p.multiProcessorCount, d));
        printf("%8.2f ", run<4>(w, p.multiProcessorCount, d));
        printf("%8.2f ", run<8>(w, p.multiProcessorCount, d));
        printf("%8.2f\n", run<16>(w, p.multiProcessorCount, d));
    }
    return 0;
}
