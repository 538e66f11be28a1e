// Development probe: cost of waiting on an mbarrier phase that has already completed (sm_100a), one warp.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mbar_probe mbar_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity, uint32_t hint)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity), "r"(hint) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool try_wait_nohint(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool test_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__global__ void k(long long* out)
{
    __shared__ uint64_t bar[64];
    const int lane = threadIdx.x;
    if (lane == 0) for (int i = 0; i < 64; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar[i])));
    __syncwarp();
    // complete phase 0 of every barrier
    if (lane == 0) for (int i = 0; i < 64; i++) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(&bar[i])) : "memory");
    __syncwarp();
    long long t[8];
    uint32_t acc = 0;
    for (int mode = 0; mode < 6; mode++) {
        __syncwarp();
        const long long t0 = clock64();
        for (int i = 0; i < 64; i++) {
            const uint32_t b = smem_u32(&bar[i]);
            switch (mode) {
            case 0: while (!try_wait(b, 0, 0x989680u)) {} break;                       // all lanes, 10 ms hint (the kernel's wait)
            case 1: while (!try_wait(b, 0, 1000u)) {} break;                           // all lanes, 1 us hint
            case 2: while (!try_wait_nohint(b, 0)) {} break;                           // all lanes, no hint
            case 3: while (!test_wait(b, 0)) {} break;                                 // all lanes, test_wait
            case 4: if (lane == 0) { while (!try_wait_nohint(b, 0)) {} } __syncwarp(); break;   // lane 0 + syncwarp
            case 5: if (lane == 0) { while (!test_wait(b, 0)) {} } __syncwarp(); break;
            }
            acc += i;
        }
        t[mode] = clock64() - t0;
    }
    if (lane == 0) { for (int m = 0; m < 6; m++) out[m] = t[m]; out[7] = acc; }
}
int main()
{
    long long* d; long long h[8];
    cudaMalloc(&d, 64);
    k<<<1, 32>>>(d); k<<<1, 32>>>(d);
    cudaDeviceSynchronize();
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    const char* names[] = {"try_wait, all lanes, 10 ms hint", "try_wait, all lanes, 1 us hint", "try_wait, all lanes, no hint", "test_wait, all lanes",
                           "try_wait, lane 0 + syncwarp", "test_wait, lane 0 + syncwarp"};
    for (int m = 0; m < 6; m++) printf("%-36s: %.1f cycles per wait on a completed phase\n", names[m], h[m] / 64.0);
    return 0;
}
