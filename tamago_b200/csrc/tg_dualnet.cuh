// tg_dualnet.cuh -- DualNet forward pass (nn/network/dual_net.py:41-106, res_block.py:27-39,
// policy_head.py:25-40, value_head.py:26-40) for sm_100a.
//
// k_dualnet_tc: the whole network for a group of boards in ONE persistent CTA per SM.
//   * every 3x3 convolution is an implicit GEMM on the 5th-gen tensor cores: tcgen05.mma (kind::f16,
//     M=128 board points x N=64 output channels x K=16 input channels per instruction), accumulators in TMEM;
//   * activations never leave the SM: they sit in shared memory as fp16 (hi, lo) pairs in the no-swizzle
//     K-major "interleaved" layout [8-channel chunk][row][8 ch], where a row is one point of the flattened,
//     halo-padded boards.  A 3x3 tap is then nothing but a row shift of the operand's start address, so the nine
//     taps are nine shared-memory descriptors over the same buffer (no im2col, no copies);
//   * fp32-grade accuracy from fp16 operands: x*w ~= x_hi*w_hi + x_lo*w_hi + x_hi*w_lo, three MMAs accumulating
//     into the same fp32 TMEM tile (the reference net is fp32, tolerance 1e-4).  Two details matter for the
//     last digit: x_lo and w_lo are stored scaled by 2^11 so that they never become fp16 subnormals (the epilogue
//     multiplies the correction sum by 2^-11), and the big x_hi*w_hi terms and the two small correction terms accumulate in SEPARATE
//     TMEM accumulators that are only added in the epilogue -- the tensor core truncates when it adds into the
//     accumulator, and that error scales with the accumulator's magnitude;
//   * shared-memory operand traffic bounds the MMA rate (a 128x16 A tile is 4 KB per instruction), so x_hi is
//     multiplied against [w_hi | w_lo] stacked to N=128 in ONE instruction (columns 0..63 -> main accumulator,
//     64..127 -> small accumulator); only x_lo needs a second, N=64 instruction;
//   * BatchNorm (eval mode) is folded into the weights/bias on the host; the residual skip goes through a
//     per-CTA fp32 scratch that stays L2 resident (128 KB per CTA) and is added in conv2's epilogue;
//   * weights stream from L2 through a cp.async.bulk/mbarrier ring of two-tap stages (16 KB per tap: [w_hi | w_lo * 2^11];
//     the x_lo instruction reads the w_hi rows of the same tile, the small accumulator carries a factor 2^11);
//   * a layer is issued as two halves (tiles 0,1 then tiles 2,3 -- or tile 2 alone at 13x13 / 19x19, where three tiles
//     cover the rows -- each over all nine taps, weights streamed once per half): the epilogue of one half runs under the MMAs of the other, so the tensor pipe never waits for it;
//   * warp roles: 0..15 = epilogue (TMEM -> bias/ReLU/split -> smem; warp w owns rows 32 w .. 32 w + 31),
//     16 = weight producer, 17 and 20 = MMA issuers (one tile of the half each), 18..19 = heads (FC layers + softmax
//     of the previous group, overlapped with the next group's convolutions).
//
// k_conv3x3_simt / k_heads_simt: plain fp32 CUDA-core implementation of the same network, used as the on-device
// numerical reference for the tensor-core kernel and for board sizes without a tensor-core instantiation.
#pragma once
#include <cstdio>
#include <cuda_fp16.h>
#include "tg_common.cuh"

namespace tg {

constexpr int NET_F = 64;                 // filters (dual_net.py:25)
constexpr int W_TAP_BYTES = 8 * 128 * 16;           // one tap: [w_hi | w_lo * 2^11] as a 128-row x 64-channel tile (16 KB)
constexpr int W_STEM_TAP_BYTES = 2 * 128 * 16;      // stem tap: the same tile with 16 input channels (4 KB)
constexpr int W_LAYER_HALVES = 9 * (W_TAP_BYTES / 2);        // fp16 elements per 64->64 layer
constexpr int SKIP_FLOATS_PER_CTA = 16 * 512 * 4;
constexpr float LO_SCALE = 2048.0f;                 // x_lo and w_lo are stored scaled by 2^11; the small accumulator is 2^11 x

struct NetDev {
    int blocks;                  // residual blocks (dual_net.py:26)
    // tensor-core operands
    const __half* w_stem;        // [9 taps][2 chunks][128 rows: w_hi oc 0..63, w_lo*2^11 oc 0..63][8 ic]
    const __half* w_conv;        // per layer [9 taps][8 chunks][128 rows: w_hi oc 0..63, w_lo*2^11 oc 0..63][8 ic]
    const float* bias;           // [1+2*blocks][64]   BN-folded bias
    const float* scale;          // [1+2*blocks]       power-of-two weight scale of the layer
    // heads (fp32)
    const float* head_w;         // [3][64]  policy conv (2) + value conv (1), BN folded
    const float* head_b;         // [3]
    const float* pfc_t;          // [2*NN][A rounded up to 4]  policy FC, transposed (rows are float4 aligned)
    const float* pfc_b;          // [A]
    const float* vfc_w;          // [3][NN]
    const float* vfc_b;          // [3]
    // fp32 CUDA-core path
    const float* w32_stem;       // [6 ic][9][64 oc]
    const float* w32_conv;       // [2*blocks][64 ic][9][64 oc]
    float* skip;                 // [CTAs][16 channel quads][512 rows][4] fp32 residual scratch (L2 resident)
    int* overflow;               // set when an activation left the fp16 operand range (|x| > 6e4) and was clamped
    long long* dbg;              // optional timing probe (development): [64] clock64 stamps of CTA 0, first group
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{ asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
    return ok != 0;
}
// A protocol error must surface as a launch failure, never as a hung GPU: every wait gives up after ~2^28 polls.
// mbar_wait: tight poll, for the MMA issuer only (any delay there is a bubble in the tensor pipe).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 28)) {
            printf("k_dualnet_tc: mbarrier wait timed out (block %d thread %d barrier %u parity %u)\n",
                   (int)blockIdx.x, (int)threadIdx.x, bar, parity);
            __trap();
        }
    }
}
// mbar_wait_relaxed: for every other role.  The pollers share their SM sub-partition's issue slots with the MMA
// issuer, so they back off between polls.
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity)
{
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(64);
        if (++spins > (1u << 26)) {
            printf("k_dualnet_tc: mbarrier wait timed out (block %d thread %d barrier %u parity %u)\n",
                   (int)blockIdx.x, (int)threadIdx.x, bar, parity);
            __trap();
        }
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar)
{ asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory"); }
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
// Same instruction with the descriptors given as their low words (start address | LBO) plus the shared high
// word (SBO | version): the issuing thread only does 32-bit adds between MMAs.
__device__ __forceinline__ void tc_mma_f16_w(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
        :: "r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate) : "memory");
}
// no-swizzle K-major operand: rows 16 B apart inside an 8-row core matrix, SBO between 8-row groups,
// LBO between the two 8-element K chunks of one K=16 step (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16)
         | ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor: D=F32, A=B=F16, both K-major, M=128, N=64 (cute::UMMA::InstrDescriptor)
constexpr uint32_t IDESC_F16_M128_N64 = (1u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t IDESC_F16_M128_N128 = (1u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

#define TG_TMEM_LD32(taddr, v) \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, " \
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" \
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), \
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), \
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), \
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) \
        : "r"(taddr) : "memory")
#define TG_TMEM_LD16(taddr, v) \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 " \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];" \
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), \
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) \
        : "r"(taddr) : "memory")
#define TG_TMEM_ST32(taddr, v) \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], " \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, " \
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31};" \
        :: "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), \
           "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), \
           "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), \
           "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]), \
           "r"(taddr) : "memory")

// ------------------------------------------------------------------------------------------------
// Geometry of the flattened, halo-padded board group held by one CTA.
//   point (y, x) of board b  ->  row  b*BR + y*PITCH + x      (PITCH = N+1: column N of every line is the halo column
//   shared by two lines, line N of every board is the halo line shared by two boards; the halo above the first board
//   is the L0 leading zero rows); tap (dy, dx) = row shift dy*PITCH + dx.
// ------------------------------------------------------------------------------------------------
template <int N, int G> struct NetGeo {
    static constexpr int NN = N * N, A = NN + 1;
    static constexpr int PITCH = N + 1;
    static constexpr int BR = PITCH * PITCH;                    // rows per board
    static constexpr int MROWS = G * BR;
    // rows that have to be COMPUTED end with the last point of the last board: its trailing halo line only has to exist
    // (as zeros) for the downward taps.  13x13 (two boards) and 19x19: three 128-row tiles instead of four.
    static constexpr int OUT_ROWS = (G - 1) * BR + (N - 1) * PITCH + N;
    static constexpr int TILES = (OUT_ROWS + 127) / 128;
    static constexpr int L0 = ((PITCH + 1 + 7) / 8) * 8;         // leading zero rows
    static constexpr int R = ((L0 + TILES * 128 + PITCH + 1 + 7) / 8) * 8;   // rows allocated per chunk plane
    static constexpr int PLANE_BYTES = R * 16;                   // one 8-channel chunk plane
    static constexpr int ACT_BYTES = 8 * PLANE_BYTES;            // one fp16 copy (hi or lo) of the activations
    // TMEM: tile t owns columns [128 t, 128 t + 128): main accumulator (x_hi w_hi) | small accumulator (corrections * 2^11)
    static_assert(TILES == 3 || TILES == 4, "the two-half layer schedule and the epilogue warp mapping assume three or four 128-row tiles");
    static constexpr int AP4 = (A + 3) & ~3;                     // policy FC outputs are produced four at a time
    // weight ring: a stage holds TWO conv taps (32 KB), which halves the issuers' stage hand-overs (their per-stage
    // bookkeeping -- wait, fence, commits next to busy epilogue warps -- is what limits the MMA issue rate, measured).
    // Two stages in the four-tile geometry, three in the three-tile one (its activation rows are 32 KB smaller).
    static constexpr int TPS = 2;                                 // conv taps per stage
    static constexpr int W_STAGES = TILES == 3 ? 3 : 2;
    static constexpr int STAGE_BYTES = TPS * W_TAP_BYTES;
    static constexpr int STEM_TPS = STAGE_BYTES / W_STEM_TAP_BYTES;   // stem taps (4 KB each) per stage
    static constexpr int JP = N >= 19 ? 5 : 3;                   // policy FC: input range split into JP partial sums
    // shared memory carve-up
    static constexpr int OFF_HI = 0;
    static constexpr int OFF_LO = OFF_HI + ACT_BYTES;
    static constexpr int OFF_W = OFF_LO + ACT_BYTES;
    static constexpr int OFF_BIAS = OFF_W + W_STAGES * STAGE_BYTES;            // [32 layers][64] fp32
    static constexpr int OFF_HEADW = OFF_BIAS + 32 * 64 * 4;                   // [3][64] + [4]
    static constexpr int OFF_PACT = OFF_HEADW + (3 * 64 + 4) * 4;              // [G][2*NN] policy-head activations
    static constexpr int OFF_VACT = OFF_PACT + G * 2 * NN * 4;                 // [G][NN]
    static constexpr int OFF_LOGIT = OFF_VACT + G * NN * 4;                    // [G][A]
    static constexpr int OFF_PART = (OFF_LOGIT + G * A * 4 + 15) & ~15;                     // [JP][G][AP4] policy FC partial sums
    static constexpr int OFF_BAR = (OFF_PART + JP * G * AP4 * 4 + 15) & ~15;     // mbarriers + tmem address
    static constexpr int SMEM_BYTES = OFF_BAR + 192;
    static_assert(SMEM_BYTES <= 232448, "shared memory budget of one SM");
};

constexpr int EPI_WARPS = 16;             // warp w owns rows 32 w .. 32 w + 31: tile w / 4, TMEM lane quarter w % 4
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int HEAD_WARPS = 2;
constexpr int HEAD_THREADS = HEAD_WARPS * 32;
constexpr int WARP_PRODUCER = EPI_WARPS, WARP_MMA = EPI_WARPS + 1, WARP_HEAD0 = EPI_WARPS + 2, WARP_MMA2 = WARP_HEAD0 + HEAD_WARPS;
constexpr int TC_THREADS = (EPI_WARPS + 3 + HEAD_WARPS) * 32;
constexpr int BND_WARP = 7;               // rows 224..255: the tail of tile 1, still read by tile 2's first taps

template <int N, int G>
__global__ void __launch_bounds__(TC_THREADS, 1)       // 21 warps -> 6 on one scheduler -> 80 registers per thread
k_dualnet_tc(NetDev P, const float* __restrict__ planes, const int* __restrict__ n_slots_ptr, int n_direct, int use_logit,
             float* __restrict__ policy, float* __restrict__ value,
             const uint8_t* __restrict__ snap, const int* __restrict__ slot_src, int snap_bytes)
{
    using NG = NetGeo<N, G>;
    extern __shared__ __align__(1024) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_slots = n_slots_ptr ? *n_slots_ptr : n_direct;     // device-side count (search) or launch argument (tg_forward*)
    const int ngroups = (n_slots + G - 1) / G;
    const int L = 1 + 2 * P.blocks;

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NG::OFF_BAR);
    const uint32_t bar_wfull = smem_u32(bars + 0);           // [NG::W_STAGES]
    const uint32_t bar_wempty = smem_u32(bars + NG::W_STAGES);   // [NG::W_STAGES]
    const uint32_t bar_accfull = smem_u32(bars + 2 * NG::W_STAGES);          // [4] per tile: accumulators of the layer complete
    const uint32_t bar_actready = smem_u32(bars + 2 * NG::W_STAGES + 4);     // [4] per tile: next layer's input rows written
    const uint32_t bar_bnd = smem_u32(bars + 2 * NG::W_STAGES + 8);          // tile 2's upward taps have read the tail of tile 1
    const uint32_t bar_headin = smem_u32(bars + 2 * NG::W_STAGES + 9);       // head activations of a group are in shared memory
    const uint32_t bar_headfree = smem_u32(bars + 2 * NG::W_STAGES + 10);    // ... and have been consumed by the head warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NG::W_STAGES + 11);
    static_assert((2 * NG::W_STAGES + 11) * 8 + 4 <= 192, "barrier block");

    // ---- one-time setup -------------------------------------------------------------------------
    {   // zero both activation copies once (halo rows stay zero for the lifetime of the CTA)
        uint4* z = reinterpret_cast<uint4*>(smem + NG::OFF_HI);
        for (int i = threadIdx.x; i < 2 * NG::ACT_BYTES / 16; i += TC_THREADS) z[i] = make_uint4(0, 0, 0, 0);
        float* bs = reinterpret_cast<float*>(smem + NG::OFF_BIAS);
        for (int i = threadIdx.x; i < L * 64; i += TC_THREADS) bs[i] = P.bias[i];
        float* hw = reinterpret_cast<float*>(smem + NG::OFF_HEADW);
        for (int i = threadIdx.x; i < 3 * 64; i += TC_THREADS) hw[i] = P.head_w[i];
        if (threadIdx.x < 3) hw[3 * 64 + threadIdx.x] = P.head_b[threadIdx.x];
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < NG::W_STAGES; s++) { mbar_init(bar_wfull + 8 * s, 1); mbar_init(bar_wempty + 8 * s, 2); }
        // bar_accfull[t]: the owner's commit + the commit of the other tile of the same half, if there is one
        for (int t = 0; t < 4; t++) { mbar_init(bar_accfull + 8 * t, (t ^ 1) < NG::TILES ? 2 : 1); mbar_init(bar_actready + 8 * t, 128); }
        mbar_init(bar_bnd, 1);
        mbar_init(bar_headin, NG::TILES * 128);
        mbar_init(bar_headfree, HEAD_THREADS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == WARP_PRODUCER) {
        // ===== weight producer: one tap per stage; a layer's weights are streamed once per half =====
        if (lane == 0) {
            uint32_t wc = 0;
            for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
                for (int l = 0; l < L; l++) {
                    // a stage holds TPS conv taps (16 KB each) or STEM_TPS stem taps (4 KB each)
                    const int tps = l == 0 ? NG::STEM_TPS : NG::TPS;
                    const int nxfer = (9 + tps - 1) / tps;
                    for (int h = 0; h < 2; h++) {
                        for (int x = 0; x < nxfer; x++, wc++) {
                            const uint32_t st = wc % NG::W_STAGES, par = (wc / NG::W_STAGES) & 1;
                            mbar_wait(bar_wempty + 8 * st, par ^ 1);
                            const uint32_t tap_bytes = l == 0 ? (uint32_t)W_STEM_TAP_BYTES : (uint32_t)W_TAP_BYTES;
                            const uint32_t bytes = (uint32_t)min(tps, 9 - x * tps) * tap_bytes;
                            const __half* src = (l == 0 ? P.w_stem : P.w_conv + (size_t)(l - 1) * W_LAYER_HALVES) + (size_t)x * tps * (tap_bytes / 2);
                            mbar_arrive_expect_tx(bar_wfull + 8 * st, bytes);
                            bulk_g2s(smem_u32(smem + NG::OFF_W + st * NG::STAGE_BYTES), src, bytes, bar_wfull + 8 * st);
                        }
                    }
                }
            }
        }
    } else if (warp == WARP_MMA || warp == WARP_MMA2) {
        // ===== MMA issuers: two warps, each owns one tile of the current half (warp A: tiles 0 and 2, warp B: tiles 1
        // and 3).  A single issuing thread leaves bubbles in the tensor pipe whenever it waits for a weight stage or
        // commits (measured, scripts/probes/mma_probe.cu: 151 -> 112 cycles per (N=128, N=64) instruction pair with two
        // issuers); with two, one thread's MMAs cover the other's bookkeeping.  Each warp walks the pipeline together,
        // one elected lane issues.
        const int tt = warp == WARP_MMA ? 0 : 1;
        uint32_t wc = 0, lc = 0;
        // descriptor low words: (address >> 4) | (LBO >> 4) << 16; high word: (SBO >> 4) | version 1
        const uint32_t a_hi0 = ((smem_u32(smem + NG::OFF_HI) >> 4) & 0x3FFFu) | ((uint32_t)(NG::PLANE_BYTES >> 4) << 16);
        const uint32_t a_lo0 = ((smem_u32(smem + NG::OFF_LO) >> 4) & 0x3FFFu) | ((uint32_t)(NG::PLANE_BYTES >> 4) << 16);
        const uint32_t w_addr16 = (smem_u32(smem + NG::OFF_W) >> 4) & 0x3FFFu;
        constexpr uint32_t LBO_W = (2048u >> 4) << 16;           // 128-row weight tile: K chunks 2 KB apart
        constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);
        constexpr uint32_t KS16 = 2 * NG::PLANE_BYTES / 16;      // A advance per K=16 step, in 16-byte units
        for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
            for (int l = 0; l < L; l++, lc++) {
                // A layer runs as two halves -- tiles 0,1 over all nine taps, then tiles 2,3 -- so that the epilogue of one
                // half (accumulators -> next layer's input rows, in place) runs under the MMAs of the other.  In-place
                // hazards: a tile's rows are read by its own taps, by the upward taps (0..3) of the next tile (its last
                // PITCH+1 rows) and by the downward taps (5..8) of the previous tile (its first PITCH+1 rows).
                // (a) inside a half: bar_accfull[t] takes two commits, one from each issuer -- the owner's after its last
                //     tap, the neighbour's after its last tap that reads tile t (tap 3 for the upward reader, tap 8 for
                //     the downward reader) -- so the epilogue of a tile starts when every reader in the half is done;
                // (b) across the half boundary: tile 2's upward taps read the tail of tile 1: bar_bnd releases those rows
                //     to the epilogue warp that owns them; tile 1's downward taps read the first rows of tile 2 and wait
                //     for tile 2's epilogue of the previous layer (bar_actready[2]) before tap 5.
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int t = 2 * h + tt;
                    if (t >= NG::TILES) {
                        // three-tile geometry: the second half has one tile and issuer A takes it alone.  (It runs at ~900
                        // cycles per tap, bound by A's per-tap bookkeeping next to busy epilogue warps, not by the weight stream;
                        // sharing the tile between the issuers makes the half faster but then the epilogue of tiles 0 and 1 is
                        // the critical path -- measured, profiles/r01_mma_probe.md.)  This issuer only keeps the ring in step:
                        // it takes every stage of the half and hands it straight back.
                        const int tps_idle = l == 0 ? NG::STEM_TPS : NG::TPS;
                        const int nstage = (9 + tps_idle - 1) / tps_idle;
                        for (int x = 0; x < nstage; x++, wc++) {
                            mbar_wait_relaxed(bar_wfull + 8 * (wc % NG::W_STAGES), (wc / NG::W_STAGES) & 1);
                            if (lane == 0) mbar_arrive(bar_wempty + 8 * (wc % NG::W_STAGES));
                            __syncwarp();
                        }
                        continue;
                    }
                    // input rows of both tiles of the half (own rows + the neighbour rows the taps reach)
                    mbar_wait(bar_actready + 8 * (2 * h), lc & 1);
                    if (2 * h + 1 < NG::TILES) mbar_wait(bar_actready + 8 * (2 * h + 1), lc & 1);
                    tc_fence_after();
                    if (P.dbg && blockIdx.x == 0 && lc < 14 && lane == 0 && t == 0) P.dbg[lc * 4 + 0] = clock64();
                    uint32_t st = 0;
                    const int tps = l == 0 ? NG::STEM_TPS : NG::TPS;      // taps per weight stage
                    for (int tap = 0; tap < 9; tap++) {
                        const bool new_stage = (tap % tps) == 0;
                        if (new_stage) {
                            st = wc % NG::W_STAGES;
                            mbar_wait(bar_wfull + 8 * st, (wc / NG::W_STAGES) & 1);
                            tc_fence_after();
                            wc++;
                        }
                        const bool stage_done = (tap % tps) == tps - 1 || tap == 8;
                        const uint32_t row16 = (uint32_t)(NG::L0 + (tap / 3 - 1) * NG::PITCH + (tap % 3 - 1));
                        const uint32_t ah = a_hi0 + row16, al = a_lo0 + row16;
                        const uint32_t wst = (w_addr16 + st * (NG::STAGE_BYTES / 16)
                                              + (uint32_t)(tap % tps) * ((l == 0 ? W_STEM_TAP_BYTES : W_TAP_BYTES) / 16)) | LBO_W;
                        if (h == 0 && tt == 1 && tap == 5) {                 // tile 1's downward taps reach into tile 2
                            mbar_wait(bar_actready + 8 * 2, lc & 1);
                            tc_fence_after();
                        }
                        if (elect_one()) {
                            if (l == 0) {
                                // K = 16: 6 input planes + zero padding, exact in fp16 -> only the weight is split
                                tc_mma_f16_w(tmem_base + t * 128, ah + t * 128, wst, DESC_HI, IDESC_F16_M128_N128, tap > 0);
                            } else {
#pragma unroll
                                for (int ks = 0; ks < 4; ks++) {
                                    // x_hi . [w_hi | w_lo 2^11] -> main | small  (initialises both at the first step)
                                    tc_mma_f16_w(tmem_base + t * 128, ah + t * 128 + ks * KS16, wst + ks * 256, DESC_HI, IDESC_F16_M128_N128,
                                                 (tap > 0 || ks > 0) ? 1u : 0u);
                                    // (x_lo 2^11) . w_hi -> small  (the first 64 rows of the same weight tile)
                                    tc_mma_f16_w(tmem_base + t * 128 + 64, al + t * 128 + ks * KS16, wst + ks * 256, DESC_HI, IDESC_F16_M128_N64, 1u);
                                }
                            }
                            if (tap == 8) {                                             // this issuer is done with the half
                                tc_commit(bar_accfull + 8 * t);
                                if (tt == 0 && t + 1 < NG::TILES) tc_commit(bar_accfull + 8 * (t + 1));   // ... and no longer reads the next tile's head
                            }
                            if (tap == 3 && tt == 1) tc_commit(bar_accfull + 8 * (t - 1));   // the previous tile's tail has been read
                            if (t == 2 && tap == 3) tc_commit(bar_bnd);               // tile 1's tail rows may be rewritten
                            if (stage_done) tc_commit(bar_wempty + 8 * st);           // stage free once both issuers' MMAs have read it
                        }
                        __syncwarp();
                    }
                }
                if (P.dbg && blockIdx.x == 0 && lc < 14 && lane == 0 && tt == 0) P.dbg[lc * 4 + 1] = clock64();
            }
        }
    } else if (warp < NG::TILES * 4) {
        // ===== epilogue warps: thread et owns row et (tile et / 128, TMEM lane et % 128) =====
        const int et = threadIdx.x;                              // 0..EPI_THREADS-1
        const int quarter = warp & 3;                            // TMEM lane quarter this warp may access
        const int tile0 = warp >> 2;
        const float* bias_s = reinterpret_cast<const float*>(smem + NG::OFF_BIAS);
        const float* headw_s = reinterpret_cast<const float*>(smem + NG::OFF_HEADW);
        float* pact = reinterpret_cast<float*>(smem + NG::OFF_PACT);
        float* vact = reinterpret_cast<float*>(smem + NG::OFF_VACT);
        uint32_t lc = 0, gi = 0;
        static_assert(NG::TILES * 128 <= EPI_THREADS, "one epilogue thread per computed row");
        const int r = et;
        const int b = r / NG::BR, q = r - b * NG::BR;
        const int y = q / NG::PITCH, x = q - y * NG::PITCH;
        const bool interior = (b < G) && (y < N) && (x < N);
        const float cap = interior ? 60000.0f : 0.0f;                        // halo rows stay zero; fp16 range guard
        float* skip_row = P.skip + (size_t)blockIdx.x * SKIP_FLOATS_PER_CTA + (size_t)r * 4;   // [quad][row][4]
        // input planes of a group -> fp16 rows (channels 0..5; 6..15 zero)
        auto load_planes = [&](int grp_) {
            const int s0 = grp_ * G;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (interior && s0 + b < n_slots) {
                float p0, p1, p2, p3, p4, p5;
                if (snap) {
                    // fused feature planes (nn/feature.py:10-57, the arithmetic of k_planes): the search kernels leave one
                    // leaf snapshot per slot (stone colours + previous move + colour to move, 16 + N^2 bytes); reading it
                    // here instead of fp32 planes removes the k_planes launch and 95 % of the evaluator's input bytes
                    const uint8_t* sn = snap + (size_t)slot_src[s0 + b] * snap_bytes;
                    const int idx = y * N + x, color = sn[3];
                    int d = sn[16 + idx];
                    if (color == 2 && d != 0) d = 3 - d;                         // :24-25 colours swap for white to move
                    p0 = d == 0 ? 1.0f : 0.0f; p1 = d == 1 ? 1.0f : 0.0f; p2 = d == 2 ? 1.0f : 0.0f;     // :31
                    p3 = (*reinterpret_cast<const int16_t*>(sn) == idx) ? 1.0f : 0.0f;                     // :43-46
                    p4 = sn[2] ? 1.0f : 0.0f;                                    // :39-41
                    p5 = color == 2 ? -1.0f : 1.0f;                              // :50-52
                } else {
                    const float* pl = planes + (size_t)(s0 + b) * 6 * NG::NN + y * N + x;
                    p0 = pl[0]; p1 = pl[NG::NN]; p2 = pl[2 * NG::NN]; p3 = pl[3 * NG::NN]; p4 = pl[4 * NG::NN]; p5 = pl[5 * NG::NN];
                }
                __half2 h01 = __floats2half2_rn(p0, p1);
                __half2 h23 = __floats2half2_rn(p2, p3);
                __half2 h45 = __floats2half2_rn(p4, p5);
                v.x = *reinterpret_cast<uint32_t*>(&h01); v.y = *reinterpret_cast<uint32_t*>(&h23);
                v.z = *reinterpret_cast<uint32_t*>(&h45);
            }
            *reinterpret_cast<uint4*>(smem + NG::OFF_HI + (NG::L0 + r) * 16) = v;
            *reinterpret_cast<uint4*>(smem + NG::OFF_HI + NG::PLANE_BYTES + (NG::L0 + r) * 16) = make_uint4(0, 0, 0, 0);
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(bar_actready + 8 * tile0);
        };
        if ((int)blockIdx.x < ngroups) load_planes(blockIdx.x);
        for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x, gi++) {
            for (int l = 0; l < L; l++, lc++) {
                const bool is_conv1 = (l >= 1) && ((l - 1) % 2 == 0);
                const bool is_conv2 = (l >= 2) && !is_conv1;
                const bool last = (l == L - 1);
                const float inv_scale = 1.0f / P.scale[l];
                const float inv_small = inv_scale / LO_SCALE;
                float hp0 = 0.f, hp1 = 0.f, hv = 0.f;
                // conv2: the first quarter of the skip row is fetched (L2) before the accumulators are waited for
                float4 sk0[4];
                if (is_conv2) {
#pragma unroll
                    for (int qd = 0; qd < 4; qd++) sk0[qd] = *reinterpret_cast<const float4*>(skip_row + (size_t)qd * 2048);
                }
                // one lane per warp polls its tile's MMA-completion barrier (and, for the rows at the end of tile 1, the
                // barrier that says tile 2's upward taps no longer read them)
                if (lane == 0) {
                    mbar_wait_relaxed(bar_accfull + 8 * tile0, lc & 1);
                    if (warp == BND_WARP) mbar_wait_relaxed(bar_bnd, lc & 1);
                    if (last && gi > 0) mbar_wait_relaxed(bar_headfree, (gi - 1) & 1);        // the previous group's head inputs were consumed
                }
                __syncwarp();
                tc_fence_after();
                if (P.dbg && blockIdx.x == 0 && lc < 14 && et == 0) P.dbg[lc * 4 + 2] = clock64();
#pragma unroll
                for (int c0 = 0; c0 < 64; c0 += 16) {
                    uint32_t v[16], w[16];
                    float4 sk[4];
                    if (is_conv2) {                                                 // issue the skip loads early (L2)
#pragma unroll
                        for (int qd = 0; qd < 4; qd++)
                            sk[qd] = c0 == 0 ? sk0[qd] : *reinterpret_cast<const float4*>(skip_row + (size_t)(c0 / 4 + qd) * 2048);
                    }
                    const uint32_t lane_col = tmem_base + ((uint32_t)(quarter * 32) << 16) + tile0 * 128 + c0;
                    TG_TMEM_LD16(lane_col, v);                                      // main accumulator: x_hi . w_hi
                    TG_TMEM_LD16(lane_col + 64, w);                                 // small accumulator: corrections * 2^11
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    float o[16];
#pragma unroll
                    for (int j = 0; j < 16; j++)
                        o[j] = fmaf(__uint_as_float(v[j]), inv_scale, fmaf(__uint_as_float(w[j]), inv_small, bias_s[l * 64 + c0 + j]));
                    if (is_conv2) {                                                 // res_block.py:39: relu(input + hidden_2)
#pragma unroll
                        for (int qd = 0; qd < 4; qd++) {
                            o[qd * 4 + 0] += sk[qd].x; o[qd * 4 + 1] += sk[qd].y; o[qd * 4 + 2] += sk[qd].z; o[qd * 4 + 3] += sk[qd].w;
                        }
                    }
                    if (interior) {                                                 // fp16 operand range: never silent
                        bool big = false;
#pragma unroll
                        for (int j = 0; j < 16; j++) big |= !(o[j] <= 60000.0f);          // (also catches NaN)
                        if (big) atomicOr(P.overflow, 1);
                    }
#pragma unroll
                    for (int j = 0; j < 16; j++) o[j] = fminf(fmaxf(o[j], 0.0f), cap);   // ReLU
                    if (last) {
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            hp0 = fmaf(o[j], headw_s[c0 + j], hp0);
                            hp1 = fmaf(o[j], headw_s[64 + c0 + j], hp1);
                            hv = fmaf(o[j], headw_s[128 + c0 + j], hv);
                        }
                    } else {
#pragma unroll
                        for (int kk = 0; kk < 2; kk++) {
                            uint32_t ph[4], pl[4];
#pragma unroll
                            for (int e = 0; e < 4; e++) {
                                const float f0 = o[kk * 8 + 2 * e], f1 = o[kk * 8 + 2 * e + 1];
                                const __half h0 = __float2half_rn(f0), h1 = __float2half_rn(f1);
                                const __half2 hh = __halves2half2(h0, h1);
                                const __half2 ll = __floats2half2_rn((f0 - __half2float(h0)) * LO_SCALE, (f1 - __half2float(h1)) * LO_SCALE);
                                ph[e] = *reinterpret_cast<const uint32_t*>(&hh);
                                pl[e] = *reinterpret_cast<const uint32_t*>(&ll);
                            }
                            const int off = (c0 / 8 + kk) * NG::PLANE_BYTES + (NG::L0 + r) * 16;
                            *reinterpret_cast<uint4*>(smem + NG::OFF_HI + off) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                            *reinterpret_cast<uint4*>(smem + NG::OFF_LO + off) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                        }
                        if (!is_conv1) {                                            // park the block input for the next skip
#pragma unroll
                            for (int qd = 0; qd < 4; qd++)
                                *reinterpret_cast<float4*>(skip_row + (size_t)(c0 / 4 + qd) * 2048) =
                                    make_float4(o[qd * 4], o[qd * 4 + 1], o[qd * 4 + 2], o[qd * 4 + 3]);
                        }
                    }
                }
                if (!last) {
                    fence_proxy_async();
                    tc_fence_before();
                    mbar_arrive(bar_actready + 8 * tile0);
                } else {
                    if (interior) {                                                 // policy_head.py:34-36, value_head.py:35-37
                        const int idx = y * N + x;
                        pact[b * 2 * NG::NN + idx] = fmaxf(hp0 + headw_s[192], 0.0f);
                        pact[b * 2 * NG::NN + NG::NN + idx] = fmaxf(hp1 + headw_s[193], 0.0f);
                        vact[b * NG::NN + idx] = fmaxf(hv + headw_s[194], 0.0f);
                    }
                    mbar_arrive(bar_headin);                                        // hand the group to the head warps
                    // the last layer wrote nothing to the activation rows and its accumulators have been read: the next
                    // group's planes go in now, so that its stem MMAs overlap the rest of this layer and the heads
                    if (grp + (int)gridDim.x < ngroups) load_planes(grp + gridDim.x);
                }
                if (P.dbg && blockIdx.x == 0 && lc < 14 && et == (NG::TILES - 1) * 128) P.dbg[lc * 4 + 3] = clock64();
            }
        }
    } else if (warp >= WARP_HEAD0 && warp < WARP_HEAD0 + HEAD_WARPS) {
        // ===== head warps: FC layers + softmax (policy_head.py:37-40, value_head.py:38-40, dual_net.py:81-106) =====
        const int ht = threadIdx.x - WARP_HEAD0 * 32, hw = warp - WARP_HEAD0;
        const float* pact = reinterpret_cast<const float*>(smem + NG::OFF_PACT);
        const float* vact = reinterpret_cast<const float*>(smem + NG::OFF_VACT);
        float* logit_s = reinterpret_cast<float*>(smem + NG::OFF_LOGIT);
        float* part = reinterpret_cast<float*>(smem + NG::OFF_PART);
        uint32_t gi = 0;
        for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x, gi++) {
            const int slot0 = grp * G;
            if (lane == 0) mbar_wait_relaxed(bar_headin, gi & 1);
            __syncwarp();
            {   // policy FC: work item = (four consecutive outputs, one JP-th of the inputs).  The weights come from L2 with
                // 16-byte loads, U of them in flight per thread (two warps have to stream the whole matrix -- 1 MB at 19x19 --
                // inside one group's convolution time); partial sums are added in a fixed order.
                constexpr int JP = NG::JP, Q = NG::AP4 / 4, JN = (2 * NG::NN + JP - 1) / JP, U = G >= 4 ? 4 : 8;
                const float4* wq = reinterpret_cast<const float4*>(P.pfc_t);
                for (int wi = ht; wi < Q * JP; wi += HEAD_THREADS) {
                    const int oq = wi % Q, jp = wi / Q;
                    const int j0 = jp * JN, j1 = min(2 * NG::NN, j0 + JN);
                    float4 acc[G];
#pragma unroll
                    for (int bb = 0; bb < G; bb++) acc[bb] = make_float4(0.f, 0.f, 0.f, 0.f);
                    int j = j0;
                    for (; j + U <= j1; j += U) {
                        float4 wv[U];
#pragma unroll
                        for (int u = 0; u < U; u++) wv[u] = __ldg(wq + (size_t)(j + u) * Q + oq);
#pragma unroll
                        for (int u = 0; u < U; u++)
#pragma unroll
                            for (int bb = 0; bb < G; bb++) {
                                const float a = pact[bb * 2 * NG::NN + j + u];
                                acc[bb].x = fmaf(wv[u].x, a, acc[bb].x); acc[bb].y = fmaf(wv[u].y, a, acc[bb].y);
                                acc[bb].z = fmaf(wv[u].z, a, acc[bb].z); acc[bb].w = fmaf(wv[u].w, a, acc[bb].w);
                            }
                    }
                    for (; j < j1; j++) {
                        const float4 wv = __ldg(wq + (size_t)j * Q + oq);
#pragma unroll
                        for (int bb = 0; bb < G; bb++) {
                            const float a = pact[bb * 2 * NG::NN + j];
                            acc[bb].x = fmaf(wv.x, a, acc[bb].x); acc[bb].y = fmaf(wv.y, a, acc[bb].y);
                            acc[bb].z = fmaf(wv.z, a, acc[bb].z); acc[bb].w = fmaf(wv.w, a, acc[bb].w);
                        }
                    }
#pragma unroll
                    for (int bb = 0; bb < G; bb++) *reinterpret_cast<float4*>(part + (jp * G + bb) * NG::AP4 + oq * 4) = acc[bb];
                }
                asm volatile("bar.sync 1, 64;" ::: "memory");
                for (int wi = ht; wi < NG::A * G; wi += HEAD_THREADS) {
                    const int o = wi % NG::A, bb = wi / NG::A;
                    float sum = P.pfc_b[o];
#pragma unroll
                    for (int jp = 0; jp < JP; jp++) sum += part[(jp * G + bb) * NG::AP4 + o];
                    logit_s[bb * NG::A + o] = sum;
                }
            }
            asm volatile("bar.sync 1, 64;" ::: "memory");
            for (int bb = hw; bb < G; bb += HEAD_WARPS) {
                const int slot = slot0 + bb;
                if (slot >= n_slots) continue;
                // value head: 3 logits + softmax
                float z[3];
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    float s = 0.f;
                    for (int j = lane; j < NG::NN; j += 32) s = fmaf(P.vfc_w[k * NG::NN + j], vact[bb * NG::NN + j], s);
#pragma unroll
                    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                    z[k] = s + P.vfc_b[k];
                }
                const float zm = fmaxf(z[0], fmaxf(z[1], z[2]));
                const float e0 = expf(z[0] - zm), e1 = expf(z[1] - zm), e2 = expf(z[2] - zm);
                const float es = e0 + e1 + e2;
                if (lane == 0) { value[(size_t)slot * 3] = e0 / es; value[(size_t)slot * 3 + 1] = e1 / es; value[(size_t)slot * 3 + 2] = e2 / es; }
                // policy
                if (use_logit) {
                    for (int o = lane; o < NG::A; o += 32) policy[(size_t)slot * NG::A + o] = logit_s[bb * NG::A + o];
                } else {
                    float mx = -3.0e38f;
                    for (int o = lane; o < NG::A; o += 32) mx = fmaxf(mx, logit_s[bb * NG::A + o]);
#pragma unroll
                    for (int o = 16; o >= 1; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                    float sum = 0.f;
                    for (int o = lane; o < NG::A; o += 32) sum += expf(logit_s[bb * NG::A + o] - mx);
#pragma unroll
                    for (int o = 16; o >= 1; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                    for (int o = lane; o < NG::A; o += 32) policy[(size_t)slot * NG::A + o] = expf(logit_s[bb * NG::A + o] - mx) / sum;
                }
            }
            mbar_arrive(bar_headfree);                  // pact / vact may be overwritten by the next group
            asm volatile("bar.sync 1, 64;" ::: "memory");   // logit_s / part are reused by the next group
        }
    }
    // ---- teardown ----
    tc_fence_before();
    __syncthreads();
    if (warp == WARP_MMA) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512));
    }
}

// ------------------------------------------------------------------------------------------------
// fp32 CUDA-core reference path
// ------------------------------------------------------------------------------------------------
// out[slot][oc][p] = act( sum_{ic,tap} w[ic][tap][oc] * in[slot][ic][p+tap] + bias[oc] (+ skip) ), one block per slot.
template <int N>
__global__ void __launch_bounds__(256) k_conv3x3_simt(const float* __restrict__ in, int cin, const float* __restrict__ w,
                                                      const float* __restrict__ bias, const float* __restrict__ skip,
                                                      float* __restrict__ out, int n_slots)
{
    constexpr int NN = N * N, W = N + 2, CELLS = W * W;
    extern __shared__ float s_in[];                    // [cin][CELLS] with a zero halo
    const int slot = blockIdx.x;
    if (slot >= n_slots) return;
    for (int i = threadIdx.x; i < cin * CELLS; i += 256) s_in[i] = 0.0f;
    __syncthreads();
    const float* src = in + (size_t)slot * cin * NN;
    for (int i = threadIdx.x; i < cin * NN; i += 256) {
        const int c = i / NN, p = i - c * NN;
        s_in[c * CELLS + (p / N + 1) * W + (p % N + 1)] = src[i];
    }
    __syncthreads();
    const int oc = threadIdx.x & 63, pg = threadIdx.x >> 6;
    constexpr int PT = 8;
    for (int p0 = pg * PT; p0 < NN; p0 += 4 * PT) {
        float acc[PT];
        int base[PT];
#pragma unroll
        for (int i = 0; i < PT; i++) {
            const int p = min(p0 + i, NN - 1);
            acc[i] = 0.0f; base[i] = (p / N) * W + (p % N);          // top-left of the 3x3 window in the padded board
        }
        for (int ic = 0; ic < cin; ic++) {
            const float* si = s_in + ic * CELLS;
#pragma unroll
            for (int tap = 0; tap < 9; tap++) {
                const float wv = __ldg(w + ((size_t)ic * 9 + tap) * 64 + oc);
                const int off = (tap / 3) * W + (tap % 3);
#pragma unroll
                for (int i = 0; i < PT; i++) acc[i] = fmaf(wv, si[base[i] + off], acc[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < PT; i++) {
            const int p = p0 + i;
            if (p < NN) {
                const size_t o = ((size_t)slot * 64 + oc) * NN + p;
                float v = acc[i] + bias[oc];
                if (skip) v += skip[o];
                out[o] = fmaxf(v, 0.0f);
            }
        }
    }
}

// heads on the CUDA cores: act [slot][64][NN] -> policy [slot][A], value [slot][3]; one block per slot
template <int N>
__global__ void __launch_bounds__(256) k_heads_simt(NetDev P, const float* __restrict__ act, int n_slots, int use_logit,
                                                    float* __restrict__ policy, float* __restrict__ value)
{
    constexpr int NN = N * N, A = NN + 1;
    __shared__ float pact[2 * NN];
    __shared__ float vact[NN];
    __shared__ float logit[A];
    __shared__ float red[8];
    const int slot = blockIdx.x;
    if (slot >= n_slots) return;
    const float* a = act + (size_t)slot * 64 * NN;
    for (int p = threadIdx.x; p < NN; p += 256) {
        float h0 = 0.f, h1 = 0.f, hv = 0.f;
        for (int c = 0; c < 64; c++) {
            const float v = a[c * NN + p];
            h0 = fmaf(v, P.head_w[c], h0); h1 = fmaf(v, P.head_w[64 + c], h1); hv = fmaf(v, P.head_w[128 + c], hv);
        }
        pact[p] = fmaxf(h0 + P.head_b[0], 0.f); pact[NN + p] = fmaxf(h1 + P.head_b[1], 0.f); vact[p] = fmaxf(hv + P.head_b[2], 0.f);
    }
    __syncthreads();
    for (int o = threadIdx.x; o < A; o += 256) {
        float s = P.pfc_b[o];
        for (int j = 0; j < 2 * NN; j++) s = fmaf(P.pfc_t[(size_t)j * ((A + 3) & ~3) + o], pact[j], s);
        logit[o] = s;
    }
    if (threadIdx.x < 3) {
        float s = P.vfc_b[threadIdx.x];
        for (int j = 0; j < NN; j++) s = fmaf(P.vfc_w[threadIdx.x * NN + j], vact[j], s);
        red[threadIdx.x] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const float zm = fmaxf(red[0], fmaxf(red[1], red[2]));
        const float e0 = expf(red[0] - zm), e1 = expf(red[1] - zm), e2 = expf(red[2] - zm), es = e0 + e1 + e2;
        value[(size_t)slot * 3] = e0 / es; value[(size_t)slot * 3 + 1] = e1 / es; value[(size_t)slot * 3 + 2] = e2 / es;
        float mx = -3.0e38f;
        for (int o = 0; o < A; o++) mx = fmaxf(mx, logit[o]);
        float sum = 0.f;
        for (int o = 0; o < A; o++) sum += expf(logit[o] - mx);
        red[4] = mx; red[5] = sum;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < A; o += 256)
        policy[(size_t)slot * A + o] = use_logit ? logit[o] : expf(logit[o] - red[4]) / red[5];
}

}  // namespace tg
