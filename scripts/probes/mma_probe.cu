// Development probe: tcgen05.mma issue patterns on one SM (cycles per instruction), sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe mma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t@p bra WD;\n\tbra WL;\n\tWD:\n\t}"
                 :: "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar)
{ asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory"); }
__device__ __forceinline__ void mma(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
                 :: "r"(d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(acc) : "memory");
}
constexpr uint32_t idesc_n(uint32_t n) { return (1u << 4) | ((n >> 3) << 17) | ((128u >> 4) << 24); }
constexpr int PLANE = 8704;

// pattern ids
// 0: chain N=128 on one accumulator          1: N=128 round-robin over 4 accumulators
// 2: chain N=64                               3: N=64 round-robin over 4 accumulators
// 4: product order: per tile 4 x (N128 @D, N64 @D+64), tiles 0..3
// 5: interleaved: per ks: N128 over 4 tiles, then N64 over 4 tiles
// 6: chain N=256                              7: N=256 round-robin over 2 accumulators
// 8: product order, N64 into a third accumulator range (no column overlap with the N128 of the same tile)
// 9: chain N=128 with a commit after every 8     10: N=192 chain (x_hi.[whi|wlo] + nothing) reference
// 11: per tile: 4 x N128 then 4 x N64            12: pairs over 2 tiles interleaved at the instruction level
__global__ void __launch_bounds__(640, 1) k_probe(int pattern, int reps, long long* out, int interf, int fill)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar, bar2, bar3;
    __shared__ uint32_t tmem_slot;
    __shared__ volatile int done;
    for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x) {
        uint32_t v = fill ? (0x3c003c00u ^ ((i * 2654435761u) & 0x83ff83ffu)) : 0u;      // fp16 values around +-1
        reinterpret_cast<uint4*>(smem)[i] = make_uint4(v, v * 3u & 0xbfffbfffu, v, v);
    }
    if (threadIdx.x == 0) done = 0;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&bar2), 1 << 20); mbar_init(smem_u32(&bar3), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t a0 = ((smem_u32(smem) >> 4) & 0x3FFFu) | ((uint32_t)(PLANE >> 4) << 16);
        const uint32_t al0 = ((smem_u32(smem + 8 * PLANE) >> 4) & 0x3FFFu) | ((uint32_t)(PLANE >> 4) << 16);
        const uint32_t w0 = ((smem_u32(smem + 16 * PLANE) >> 4) & 0x3FFFu) | ((2048u >> 4) << 16);
        const uint32_t w256 = ((smem_u32(smem + 16 * PLANE) >> 4) & 0x3FFFu) | ((4096u >> 4) << 16);
        constexpr uint32_t DH = (128u >> 4) | (1u << 14);
        constexpr uint32_t KS = 2 * PLANE / 16;
        uint32_t par = 0;
        for (int rep = 0; rep < 3; rep++) {                  // rep 0,1 warm up
            long long n = 0;
            const long long t0 = clock64();
            for (int r = 0; r < reps; r++) {
                const uint32_t shift = 16 + (r % 9);          // row shift like a tap
                switch (pattern) {
                case 0: for (int i = 0; i < 32; i++, n++) mma(tm, a0 + shift + (i & 3) * KS, w0 + (i & 3) * 256, DH, idesc_n(128), 1); break;
                case 1: for (int i = 0; i < 32; i++, n++) mma(tm + (i & 3) * 128, a0 + shift + (i & 3) * 128, w0, DH, idesc_n(128), 1); break;
                case 2: for (int i = 0; i < 32; i++, n++) mma(tm, a0 + shift + (i & 3) * KS, w0 + (i & 3) * 256, DH, idesc_n(64), 1); break;
                case 3: for (int i = 0; i < 32; i++, n++) mma(tm + (i & 3) * 128, a0 + shift + (i & 3) * 128, w0, DH, idesc_n(64), 1); break;
                case 4:
                    for (int t = 0; t < 4; t++)
                        for (int ks = 0; ks < 4; ks++, n += 2) {
                            mma(tm + t * 128, a0 + shift + t * 128 + ks * KS, w0 + ks * 256, DH, idesc_n(128), 1);
                            mma(tm + t * 128 + 64, al0 + shift + t * 128 + ks * KS, w0 + ks * 256, DH, idesc_n(64), 1);
                        }
                    break;
                case 5:
                    for (int ks = 0; ks < 4; ks++) {
                        for (int t = 0; t < 4; t++, n++) mma(tm + t * 128, a0 + shift + t * 128 + ks * KS, w0 + ks * 256, DH, idesc_n(128), 1);
                        for (int t = 0; t < 4; t++, n++) mma(tm + t * 128 + 64, al0 + shift + t * 128 + ks * KS, w0 + ks * 256, DH, idesc_n(64), 1);
                    }
                    break;
                case 6: for (int i = 0; i < 32; i++, n++) mma(tm, a0 + shift + (i & 3) * KS, w256, DH, idesc_n(256), 1); break;
                case 7: for (int i = 0; i < 32; i++, n++) mma(tm + (i & 1) * 256, a0 + shift + (i & 1) * 128, w256, DH, idesc_n(256), 1); break;
                case 8:
                    for (int t = 0; t < 2; t++)
                        for (int ks = 0; ks < 4; ks++, n += 2) {
                            mma(tm + t * 128, a0 + shift + t * 128 + ks * KS, w0 + ks * 256, DH, idesc_n(128), 1);
                            mma(tm + 256 + t * 64, al0 + shift + t * 128 + ks * KS, w0 + ks * 256, DH, idesc_n(64), 1);
                        }
                    break;
                case 9:
                    for (int i = 0; i < 32; i++, n++) {
                        mma(tm + (i & 3) * 128, a0 + shift + (i & 3) * 128, w0, DH, idesc_n(128), 1);
                        if ((i & 7) == 7) tc_commit(smem_u32(&bar) + 0 * 0), mbar_wait(smem_u32(&bar), par), par ^= 1;
                    }
                    break;
                case 10: for (int i = 0; i < 32; i++, n++) mma(tm, a0 + shift + (i & 3) * KS, w256, DH, idesc_n(192), 1); break;
                case 11:
                    for (int t = 0; t < 4; t++) {
                        for (int ks = 0; ks < 4; ks++, n++) mma(tm + t * 128, a0 + shift + t * 128 + ks * KS, w0 + ks * 256, DH, idesc_n(128), 1);
                        for (int ks = 0; ks < 4; ks++, n++) mma(tm + t * 128 + 64, al0 + shift + t * 128 + ks * KS, w0 + ks * 256, DH, idesc_n(64), 1);
                    }
                    break;
                case 13: case 14: case 15: case 16:
                    for (int t = 0; t < 4; t++) {
                        for (int ks = 0; ks < 4; ks++, n += 2) {
                            mma(tm + t * 128, a0 + shift + t * 128 + ks * KS, w0 + ks * 256, DH, idesc_n(128), 1);
                            mma(tm + t * 128 + 64, al0 + shift + t * 128 + ks * KS, w0 + ks * 256, DH, idesc_n(64), 1);
                        }
                        if (t & 1) {
                            if (pattern == 13) tc_commit(smem_u32(&bar2));
                            if (pattern == 14) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                            if (pattern == 15) mbar_wait(smem_u32(&bar3), 1);        // phase 0 not complete -> parity 1 passes immediately
                            if (pattern == 16) { tc_commit(smem_u32(&bar2)); mbar_wait(smem_u32(&bar3), 1); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
                        }
                    }
                    break;
                case 17: for (int i = 0; i < 32; i++, n++) mma(tm, a0 + shift + (i & 3) * KS, w0 + (i & 3) * 256, DH, idesc_n(32), 1); break;
                case 18: for (int i = 0; i < 32; i++, n++) mma(tm, a0 + shift + (i & 3) * KS, w0 + (i & 3) * 256, DH, idesc_n(16), 1); break;
                case 12:
                    for (int tp = 0; tp < 2; tp++)
                        for (int ks = 0; ks < 4; ks++, n += 4) {
                            mma(tm + (2 * tp) * 128, a0 + shift + (2 * tp) * 128 + ks * KS, w0 + ks * 256, DH, idesc_n(128), 1);
                            mma(tm + (2 * tp + 1) * 128, a0 + shift + (2 * tp + 1) * 128 + ks * KS, w0 + ks * 256, DH, idesc_n(128), 1);
                            mma(tm + (2 * tp) * 128 + 64, al0 + shift + (2 * tp) * 128 + ks * KS, w0 + ks * 256, DH, idesc_n(64), 1);
                            mma(tm + (2 * tp + 1) * 128 + 64, al0 + shift + (2 * tp + 1) * 128 + ks * KS, w0 + ks * 256, DH, idesc_n(64), 1);
                        }
                    break;
                }
            }
            const long long t1 = clock64();
            tc_commit(smem_u32(&bar));
            mbar_wait(smem_u32(&bar), par); par ^= 1;
            const long long t2 = clock64();
            if (rep == 2) { out[0] = t1 - t0; out[1] = t2 - t0; out[2] = n; }
        }
        done = 1;
    } else if (threadIdx.x >= 128) {
        // interference warps (16 of them, 4 per TMEM lane quarter)
        const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
        uint32_t acc = 0;
        unsigned char* scratch = smem + 190 * 1024 + (w - 4) * 512;
        while (!done) {
            if (interf == 1) {              // TMEM loads
                uint32_t v[16];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(tm + ((uint32_t)((w & 3) * 32) << 16) + ((w >> 2) & 3) * 128));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc += v[0] + v[15];
            } else if (interf == 2) {       // shared stores, 16 B per lane
                asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" :: "r"(smem_u32(scratch + lane * 16)), "r"(acc) : "memory");
                acc++;
            } else if (interf == 3) {       // fp32 math
#pragma unroll
                for (int i = 0; i < 64; i++) acc = acc * 1664525u + 1013904223u;
            } else if (interf == 4) {       // sleeping poll
                __nanosleep(64);
            } else if (interf == 5) {       // global (L2) loads
                acc += *reinterpret_cast<volatile long long*>(out + 8 + ((threadIdx.x * 16 + acc) & 1023));
            }
            // interf == 0: tight poll of the flag
        }
        if (acc == 0x12345678u) out[3] = acc;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tm), "r"(512));
    }
}

// two issuing threads (warps 0 and 1), each walking "product order" over its own two tiles with the per-stage overheads
__global__ void __launch_bounds__(128, 1) k_dual(int issuers, int reps, int overhead, long long* out)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar[4], bar2, bar3;
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x3c003c00u, 0x3c00bc00u, 0x38003c00u, 0xbc003c00u);
    if (threadIdx.x == 0) { for (int i = 0; i < 4; i++) mbar_init(smem_u32(&bar[i]), 1); mbar_init(smem_u32(&bar2), 1 << 20); mbar_init(smem_u32(&bar3), 1);
                            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_slot;
    const int w = threadIdx.x >> 5;
    const long long t0 = clock64();
    if ((threadIdx.x & 31) == 0 && w < issuers) {
        const uint32_t a0 = ((smem_u32(smem) >> 4) & 0x3FFFu) | ((uint32_t)(PLANE >> 4) << 16);
        const uint32_t al0 = ((smem_u32(smem + 8 * PLANE) >> 4) & 0x3FFFu) | ((uint32_t)(PLANE >> 4) << 16);
        const uint32_t w0 = ((smem_u32(smem + 16 * PLANE) >> 4) & 0x3FFFu) | ((2048u >> 4) << 16);
        constexpr uint32_t DH = (128u >> 4) | (1u << 14);
        constexpr uint32_t KS = 2 * PLANE / 16;
        const int tiles_per = 4 / issuers;
        long long n = 0;
        for (int r = 0; r < reps; r++) {
            const uint32_t shift = 16 + (r % 9);
            for (int tt = 0; tt < tiles_per; tt++) {
                const int t = w * tiles_per + tt;
                for (int ks = 0; ks < 4; ks++, n += 2) {
                    mma(tm + t * 128, a0 + shift + t * 128 + ks * KS, w0 + ks * 256, DH, idesc_n(128), 1);
                    mma(tm + t * 128 + 64, al0 + shift + t * 128 + ks * KS, w0 + ks * 256, DH, idesc_n(64), 1);
                }
                if (overhead && ((tt & 1) || tiles_per == 1 || overhead == 2)) {
                    tc_commit(smem_u32(&bar2)); mbar_wait(smem_u32(&bar3), 1); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
            }
        }
        tc_commit(smem_u32(&bar[w]));
        mbar_wait(smem_u32(&bar[w]), 0);
        out[4 + w] = clock64() - t0;
        out[8 + w] = n;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tm), "r"(512));
    }
}

// ---- CTA pair (cta_group::2): M = 256 (128 rows per CTA), each CTA supplies half of the N rows of B ----
__device__ __forceinline__ void mma2(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}"
                 :: "r"(d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(acc) : "memory");
}
constexpr uint32_t idesc2_n(uint32_t n) { return (1u << 4) | ((n >> 3) << 17) | ((256u >> 4) << 24); }
__device__ __forceinline__ void cluster_sync()
{ asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k_pair(int issuers, int reps, long long* out)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar[4];
    __shared__ uint32_t tmem_slot;
    uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x3c003c00u, 0x3c00bc00u, 0x38003c00u, 0xbc003c00u);
    if (threadIdx.x == 0) { for (int i = 0; i < 4; i++) mbar_init(smem_u32(&bar[i]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_slot;
    const int w = threadIdx.x >> 5;
    const long long t0 = clock64();
    if (rank == 0 && (threadIdx.x & 31) == 0 && w < issuers) {
        const uint32_t a0 = ((smem_u32(smem) >> 4) & 0x3FFFu) | ((uint32_t)(PLANE >> 4) << 16);
        const uint32_t al0 = ((smem_u32(smem + 8 * PLANE) >> 4) & 0x3FFFu) | ((uint32_t)(PLANE >> 4) << 16);
        const uint32_t w1 = ((smem_u32(smem + 16 * PLANE) >> 4) & 0x3FFFu) | ((1024u >> 4) << 16);          // 64 rows per CTA, chunks 1 KB apart
        const uint32_t w2 = ((smem_u32(smem + 16 * PLANE + 8192) >> 4) & 0x3FFFu) | ((512u >> 4) << 16);    // 32 rows per CTA, chunks 512 B apart
        constexpr uint32_t DH = (128u >> 4) | (1u << 14);
        constexpr uint32_t KS = 2 * PLANE / 16;
        const int tiles_per = 4 / issuers;
        long long n = 0;
        for (int r = 0; r < reps; r++) {
            const uint32_t shift = 16 + (r % 9);
            for (int tt = 0; tt < tiles_per; tt++) {
                const int t = w * tiles_per + tt;
                for (int ks = 0; ks < 4; ks++, n += 2) {
                    mma2(tm + t * 128, a0 + shift + t * 128 + ks * KS, w1 + ks * 128, DH, idesc2_n(128), 1);
                    mma2(tm + t * 128 + 64, al0 + shift + t * 128 + ks * KS, w2 + ks * 64, DH, idesc2_n(64), 1);
                }
            }
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar[w])) : "memory");
        mbar_wait(smem_u32(&bar[w]), 0);
        out[4 + w] = clock64() - t0;
        out[8 + w] = n;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tm), "r"(512));
    }
}

// two issuers sharing ONE tile: issuer w takes every other "tap" (8 MMAs: 4 x (N128, N64)) into its own accumulator pair,
// with the kernel's per-tap bookkeeping (commit + wait + fence) after each tap
__global__ void __launch_bounds__(128, 1) k_coop(int issuers, int reps, long long* out)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar[4], bar2, bar3;
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x3c003c00u, 0x3c00bc00u, 0x38003c00u, 0xbc003c00u);
    if (threadIdx.x == 0) { for (int i = 0; i < 4; i++) mbar_init(smem_u32(&bar[i]), 1); mbar_init(smem_u32(&bar2), 1 << 20); mbar_init(smem_u32(&bar3), 1);
                            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_slot;
    const int w = threadIdx.x >> 5;
    const long long t0 = clock64();
    if ((threadIdx.x & 31) == 0 && w < issuers) {
        const uint32_t a0 = ((smem_u32(smem) >> 4) & 0x3FFFu) | ((uint32_t)(PLANE >> 4) << 16);
        const uint32_t al0 = ((smem_u32(smem + 8 * PLANE) >> 4) & 0x3FFFu) | ((uint32_t)(PLANE >> 4) << 16);
        const uint32_t w0 = ((smem_u32(smem + 16 * PLANE) >> 4) & 0x3FFFu) | ((2048u >> 4) << 16);
        constexpr uint32_t DH = (128u >> 4) | (1u << 14);
        constexpr uint32_t KS = 2 * PLANE / 16;
        long long n = 0;
        for (int r = 0; r < reps; r++) {
            if ((r % issuers) != w) continue;
            const uint32_t shift = 16 + (r % 9);
            for (int ks = 0; ks < 4; ks++, n += 2) {
                mma(tm + w * 128, a0 + shift + ks * KS, w0 + ks * 256, DH, idesc_n(128), 1);
                mma(tm + w * 128 + 64, al0 + shift + ks * KS, w0 + ks * 256, DH, idesc_n(64), 1);
            }
            tc_commit(smem_u32(&bar2)); mbar_wait(smem_u32(&bar3), 1); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        tc_commit(smem_u32(&bar[w]));
        mbar_wait(smem_u32(&bar[w]), 0);
        out[4 + w] = clock64() - t0;
        out[8 + w] = n;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tm), "r"(512));
    }
}

int main()
{
    long long* d; long long h[3];
    cudaMalloc(&d, 16384); cudaMemset(d, 0, 16384);
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const char* names[] = {"chain N128", "N128 rr4", "chain N64", "N64 rr4", "product order (tile: 4x(N128,N64))", "interleaved (ks: 4xN128, 4xN64)",
                           "chain N256", "N256 rr2", "product order, N64 into separate columns", "N128 rr4 + commit/wait every 8", "chain N192",
                           "per tile 4xN128 then 4xN64", "pairs over 2 tiles interleaved",
        "product + commit every 16", "product + fence::after every 16", "product + passing try_wait every 16", "product + commit, wait, fence every 16", "chain N32", "chain N16"};
    cudaFuncSetAttribute(k_dual, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_coop, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int issuers = 1; issuers <= 4; issuers *= 2) {
        long long hh[16];
        cudaMemset(d, 0, 16384);
        k_coop<<<1, 128, 200 * 1024>>>(issuers, 1024, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("coop %d: %s\n", issuers, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hh, d, sizeof hh, cudaMemcpyDeviceToHost);
        long long tot = 0, tmax = 0;
        for (int w = 0; w < issuers; w++) { tot += hh[8 + w]; tmax = hh[4 + w] > tmax ? hh[4 + w] : tmax; }
        printf("one tile shared by %d issuer(s), taps alternate, bookkeeping per tap: %.1f cyc per (N128, N64) pair\n", issuers, 2.0 * tmax / tot);
    }
    cudaFuncSetAttribute(k_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int grid = 2; grid <= 148; grid *= 74)
    for (int issuers = 1; issuers <= 4; issuers *= 2) {
        long long hh[16];
        cudaMemset(d, 0, 16384);
        k_pair<<<grid, 128, 200 * 1024>>>(issuers, 256, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("pair %d: %s\n", issuers, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hh, d, sizeof hh, cudaMemcpyDeviceToHost);
        long long tot = 0, tmax = 0;
        for (int w = 0; w < issuers; w++) { tot += hh[8 + w]; tmax = hh[4 + w] > tmax ? hh[4 + w] : tmax; }
        printf("CTA pair (cta_group::2, M=256), grid %d: %d issuer thread(s): %.1f cyc per (N128, N64) instruction pair\n", grid, issuers, 2.0 * tmax / tot);
    }
    for (int issuers = 1; issuers <= 4; issuers *= 2)
        for (int ov = 0; ov <= 2; ov++) {
            long long hh[16];
            k_dual<<<1, 128, 200 * 1024>>>(issuers, 256, ov, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("dual %d: %s\n", issuers, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(hh, d, sizeof hh, cudaMemcpyDeviceToHost);
            long long tot = 0, tmax = 0;
            for (int w = 0; w < issuers; w++) { tot += hh[8 + w]; tmax = hh[4 + w] > tmax ? hh[4 + w] : tmax; }
            printf("dual: %d issuer thread(s), overhead %s: %.1f cyc/MMA\n", issuers, ov == 0 ? "none" : ov == 1 ? "per 16 MMAs" : "per 8 MMAs", (double)tmax / tot);
        }
    // sustained (power-capped) executed throughput: ~1.5 s of back-to-back launches, single CTAs with two issuers vs CTA pairs
    for (int mode = 0; mode < 2; mode++) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        const int reps = 65536, launches = 12;
        if (mode == 0) k_dual<<<148, 128, 200 * 1024>>>(2, 1024, 0, d); else k_pair<<<148, 128, 200 * 1024>>>(2, 1024, d);
        cudaEventRecord(e0);
        for (int i = 0; i < launches; i++) {
            if (mode == 0) k_dual<<<148, 128, 200 * 1024>>>(2, reps, 0, d); else k_pair<<<148, 128, 200 * 1024>>>(2, reps, d);
        }
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("sustained %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        // per CTA and rep: 4 tiles x 4 k-steps x (N128 + N64) on 128 rows
        const double flop = (double)launches * 148 * reps * 16.0 * 2.0 * 128 * 16 * 192;
        printf("sustained %s: %.1f ms, %.1f TFLOP/s executed\n", mode == 0 ? "single CTAs, 2 issuers" : "CTA pairs, 2 issuers on the leader", ms, flop / ms / 1e9);
    }
    const char* inames[] = {"tight flag poll", "TMEM loads", "shared stores", "integer math", "sleeping poll", "L2 loads", "no extra warps"};
    for (int fill = 0; fill < 2; fill++)
    for (int interf = 0; interf <= 6; interf++) {
        k_probe<<<1, interf == 6 ? 128 : 640, 200 * 1024>>>(4, 64, d, interf, fill);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("interf %d: %s\n", interf, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
        printf("product order, data %s, 16 warps doing %-16s: %.1f cyc/MMA\n", fill ? "random" : "zero", inames[interf], (double)h[1] / h[2]);
    }
    for (int fill = 0; fill < 2; fill++)
    for (int grid = 1; grid <= 148; grid = grid == 1 ? 37 : grid * 2) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k_probe<<<grid, 128, 200 * 1024>>>(4, 4096, d, 6, fill);     // warm
        cudaEventRecord(e0);
        k_probe<<<grid, 128, 200 * 1024>>>(4, 16384, d, 6, fill);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("grid %d: %s\n", grid, cudaGetErrorString(e)); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
        // executed flop: per pair of MMAs 2*128*16*(128+64)
        const double flop = (double)grid * 3.0 * h[2] / 2 * 2.0 * 128 * 16 * 192;
        printf("grid %3d data %s: %.1f cyc/MMA, kernel %.2f ms (3 reps) -> %.0f MHz effective, %.1f TFLOP/s executed\n", grid, fill ? "random" : "zero",
               (double)h[1] / h[2], ms, 3.0 * h[1] / ms / 1e3, flop / ms / 1e9);
    }
    for (int p = 0; p <= 18; p++) {
        k_probe<<<1, 128, 200 * 1024>>>(p, 64, d, 6, 1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("pattern %d: %s\n", p, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
        printf("pattern %2d %-45s: %lld MMAs, issue %.1f cyc/MMA, complete %.1f cyc/MMA\n", p, names[p], h[2], (double)h[0] / h[2], (double)h[1] / h[2]);
    }
    return 0;
}
