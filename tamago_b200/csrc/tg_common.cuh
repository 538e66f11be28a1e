// tg_common.cuh -- shared constants and small device helpers (sm_100a).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned long long u64;

namespace tg {

enum : int { EMPTY = 0, BLACK = 1, WHITE = 2, OB = 3 };
enum : int { PASS = 0, RESIGN = -1, NOT_EXPANDED = -1 };

// reference constants: mcts/constant.py:5-41
constexpr int    C_VISIT = 50;
constexpr double C_SCALE = 1.0;
constexpr int    MAX_CONSIDERED = 16;
constexpr int    PLAYOUTS = 100;

constexpr int BLOOM_WORDS = 128;          // 4096-bit filter over the position history (super-ko pre-check)

// Board geometry, all derived from the board size like board/constant.py:4-31.
template <int N> struct Geo {
    static constexpr int W = N + 2;
    static constexpr int CELLS = W * W;
    static constexpr int CP = (CELLS + 3) & ~3;       // padded cell count
    static constexpr int NN = N * N;
    static constexpr int A = NN + 1;                  // MAX_ACTIONS (mcts/node.py:15)
    static constexpr int AP = (A + 31) & ~31;         // child-array stride
    static constexpr int MAXREC = 3 * NN;             // MAX_RECORDS (constant.py:31)
    static constexpr int PLANES = 6 * NN;
};

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int opp(int c) { return 3 - c; }         // only used with BLACK/WHITE

__device__ __forceinline__ u64 warp_xor64(u64 v)
{
    unsigned lo = (unsigned)v, hi = (unsigned)(v >> 32);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        lo ^= __shfl_xor_sync(0xffffffffu, lo, o);
        hi ^= __shfl_xor_sync(0xffffffffu, hi, o);
    }
    return ((u64)hi << 32) | lo;
}
__device__ __forceinline__ int warp_sum_i(int v)
{
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_max_i(int v)
{
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double shfl_xor_d(double v, int o)
{
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(0xffffffffu, lo, o);
    hi = __shfl_xor_sync(0xffffffffu, hi, o);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_d(double v, int src)
{
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_sync(0xffffffffu, lo, src);
    hi = __shfl_sync(0xffffffffu, hi, src);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ u64 shfl_u64(u64 v, int src)
{
    unsigned lo = (unsigned)v, hi = (unsigned)(v >> 32);
    lo = __shfl_sync(0xffffffffu, lo, src);
    hi = __shfl_sync(0xffffffffu, hi, src);
    return ((u64)hi << 32) | lo;
}

// argmax with numpy semantics (first index wins ties): lanes hold (value, index), index = INT_MAX when empty.
//   Doubles are mapped to order-preserving 64-bit integer keys, so the reduction is three REDUX instructions (max of the
//   high words, max of the low words among the holders of that maximum, min of the indices among the holders of both)
//   instead of a five-round shuffle butterfly on (double, index) pairs.  Scores are never NaN; -0.0 is folded into +0.0.
__device__ __forceinline__ u64 order_key(double v)
{
    const u64 b = (u64)__double_as_longlong(v == 0.0 ? 0.0 : v);
    return b ^ ((u64)((long long)b >> 63) | 0x8000000000000000ull);
}
__device__ __forceinline__ void warp_argmax_key(u64& key, int& idx)
{   // key = 0 marks an empty lane (real keys are never 0)
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
    const unsigned mi = __reduce_min_sync(0xffffffffu, (key != 0ull && hi == mhi && lo == mlo) ? (unsigned)idx : 0x7fffffffu);
    key = ((u64)mhi << 32) | mlo; idx = (int)mi;
}
__device__ __forceinline__ void warp_argmax_d(double& v, int& idx)
{
    u64 key = idx != 0x7fffffff ? order_key(v) : 0ull;
    warp_argmax_key(key, idx);
    // (callers only use idx; v is left as this lane's own value)
}

}  // namespace tg
