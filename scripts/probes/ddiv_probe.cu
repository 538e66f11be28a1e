// ddiv_probe -- ddiv_fast (tg_detmath.cuh) against __ddiv_rn, bit for bit, on the operand ranges of the PUCT scores (value
// sums / visit counts, prior * sqrt(n) / (visits + 1)) and on random bit patterns.  Wherever ddiv_fast does not raise `slow`
// the two must agree; in the PUCT ranges `slow` must be raised only for numerators below 2^-120.
// build + run: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -I tamago_b200/csrc -o /tmp/ddiv_probe scripts/probes/ddiv_probe.cu && /tmp/ddiv_probe
#include <cstdio>
#include <cstdint>
#include "tg_detmath.cuh"
using namespace tg;

__device__ unsigned long long g_bad, g_slow, g_slow_in_range, g_n;

__global__ void k_check(int mode, unsigned long long seed, int per_thread)
{
    const unsigned long long tid = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    unsigned long long bad = 0, slow = 0, sir = 0;
    for (int r = 0; r < per_thread; r++) {
        const u64 h1 = mix64(seed + tid * 1000003ull + (u64)r * 2ull), h2 = mix64(h1 ^ 0x1234567ull);
        double a, b;
        if (mode == 0) {                       // q: fp32 value sum over an integer visit count
            b = (double)(1 + (int)(h2 % 4000));
            a = (double)((float)((h1 >> 11) * (1.0 / 9007199254740992.0)) * (float)b);
        } else if (mode == 1) {                // u: prior (fp32 or float64, down to 1e-30) * sqrt(n) over visits + 1
            const double p = (h1 & 1) ? (double)(float)((h1 >> 11) * (1.0 / 9007199254740992.0)) : (h1 >> 11) * (1.0 / 9007199254740992.0);
            const double scale = __longlong_as_double((long long)(1023 - (int)((h2 >> 40) % 100)) << 52);
            a = __dmul_rn(__dmul_rn(p, scale), sqrt((double)(1 + (int)((h2 >> 8) % 4000))));
            b = (double)(1 + (int)(h2 % 4000));
        } else {                               // random bit patterns (finite, any exponent)
            a = __longlong_as_double((long long)h1); b = __longlong_as_double((long long)h2);
        }
        bool sl;
        const double f = ddiv_fast(a, b, sl);
        const double ref = __ddiv_rn(a, b);
        if (sl) { slow++; if (mode < 2 && fabs(a) >= 7.52316384526264e-37) sir++; }
        else if (__double_as_longlong(f) != __double_as_longlong(ref) && !(f != f && ref != ref)) bad++;
    }
    atomicAdd(&g_bad, bad); atomicAdd(&g_slow, slow); atomicAdd(&g_slow_in_range, sir); atomicAdd(&g_n, (unsigned long long)per_thread);
}

int main()
{
    int rc = 0;
    for (int mode = 0; mode < 3; mode++) {
        unsigned long long z = 0;
        cudaMemcpyToSymbol(g_bad, &z, 8); cudaMemcpyToSymbol(g_slow, &z, 8); cudaMemcpyToSymbol(g_slow_in_range, &z, 8); cudaMemcpyToSymbol(g_n, &z, 8);
        k_check<<<1184, 256>>>(mode, 0x9e3779b9ull * (mode + 1), 1024);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed\n"); return 2; }
        unsigned long long bad, slow, sir, n;
        cudaMemcpyFromSymbol(&bad, g_bad, 8); cudaMemcpyFromSymbol(&slow, g_slow, 8); cudaMemcpyFromSymbol(&sir, g_slow_in_range, 8); cudaMemcpyFromSymbol(&n, g_n, 8);
        printf("mode %d: %llu pairs, %llu mismatches, %llu flagged slow, %llu flagged slow with a numerator >= 2^-120\n", mode, n, bad, slow, sir);
        if (bad || sir) rc = 1;
    }
    printf(rc ? "FAIL\n" : "OK\n");
    return rc;
}
