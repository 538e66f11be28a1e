#include <cstdio>
__global__ void k_dp(double* out, long long* cyc, int iters, int chains)
{
    double x[8];
    for (int c = 0; c < 8; c++) x[c] = 1.0 + threadIdx.x * 1e-9 + c;
    const double a = 1.0000001, b = 1e-9;
    __syncthreads();
    const long long t0 = clock64();
    if (chains == 1) for (int i = 0; i < iters; i++) x[0] = __fma_rn(x[0], a, b);
    else if (chains == 2) for (int i = 0; i < iters; i++) { x[0] = __fma_rn(x[0], a, b); x[1] = __fma_rn(x[1], a, b); }
    else if (chains == 4) for (int i = 0; i < iters; i++) { x[0] = __fma_rn(x[0], a, b); x[1] = __fma_rn(x[1], a, b); x[2] = __fma_rn(x[2], a, b); x[3] = __fma_rn(x[3], a, b); }
    else for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < 8; c++) x[c] = __fma_rn(x[c], a, b);
    }
    const long long t1 = clock64();
    double s = 0; for (int c = 0; c < 8; c++) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main()
{
    double* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    for (int threads : {32, 128, 512, 1024}) for (int chains : {1, 2, 4, 8}) {
        k_dp<<<1, threads>>>(out, cyc, iters, chains); cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("threads %4d chains %d: %.1f cycles per iteration (%d DFMA per thread) -> %.2f DFMA lanes/clk/SM\n", threads, chains, (double)h / iters, chains,
               (double)threads * chains * iters / h);
    }
    return 0;
}
