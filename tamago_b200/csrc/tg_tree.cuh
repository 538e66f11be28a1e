// tg_tree.cuh -- flat MCTS node pool and warp-cooperative selection rules.
//
// Reference: mcts/node.py (MCTSNode 18-39, select_next_action 141-157,
// calculate_completed_q_value 281-305, calculate_improved_policy 308-321,
// select_move_by_sequential_halving_for_root 324-346, ..._for_node 349-361),
// mcts/pucb/pucb.py:8-29, mcts/sequential_halving.py:7-60.
//
// One warp owns one game's tree.  Child statistics are rows of AP entries per
// node (structure of arrays); a selection sweeps a row 32 children at a time
// and finishes with a warp-shuffle argmax that keeps numpy's first-index tie
// break.  All float64 arithmetic uses single correctly rounded operations
// (tg_detmath.cuh) so decisions are bit-identical to the reference's numpy math.
#pragma once
#include "tg_common.cuh"
#include "tg_detmath.cuh"

namespace tg {

enum : int { H_K = 0, H_NV = 1, H_VL = 2, H_VSUM = 3, H_RAW = 4, H_STRIDE = 8 };

// Pointers to one game's slice of the node pool.
struct Tree {
    int*      hdr;     // [max_nodes][H_STRIDE]
    int16_t*  action;  // [max_nodes][AP]      node.py:30
    int*      cidx;    // children_index        node.py:31
    float*    cval;    // children_value        node.py:32
    int*      cvis;    // children_visits       node.py:33
    double*   cpol;    // children_policy       node.py:34
    int*      cvl;     // children_virtual_loss node.py:35
    float*    cvsum;   // children_value_sum    node.py:36 (fp32-accumulated: the addend is a torch scalar)
    double*   noise;   // [AP] root noise       node.py:37
};

struct TreePool {
    int*      hdr;  int16_t* action;  int* cidx;  float* cval;  int* cvis;  double* cpol;  int* cvl;  float* cvsum;
    double*   noise;
    int       max_nodes;
};

template <int AP> __device__ __forceinline__ Tree tree_of(const TreePool& p, int g)
{
    Tree t;
    const size_t nb = (size_t)g * p.max_nodes;
    t.hdr = p.hdr + nb * H_STRIDE;
    t.action = p.action + nb * AP;  t.cidx = p.cidx + nb * AP;  t.cval = p.cval + nb * AP;  t.cvis = p.cvis + nb * AP;
    t.cpol = p.cpol + nb * AP;      t.cvl = p.cvl + nb * AP;    t.cvsum = p.cvsum + nb * AP;
    t.noise = p.noise + (size_t)g * AP;
    return t;
}

// Shared-memory staging of one node's child rows for the PUCT selection ("child visit/value stats staged in shared memory").
// Everything a ply reads from the node -- visits, virtual losses, value sums, priors, child indices, actions and the node
// header -- is fetched with ONE burst of 16-byte asynchronous copies; the descent issues the burst for the next node as
// soon as it knows it (right after the selection, under put_stone), so the L2 latency of a ply is hidden.
struct SelStage { double* pol; int* vis; int* vl; float* vsum; int* cidx; int16_t* action; int* hdr; };

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}

template <int AP>
__device__ __forceinline__ void stage_node_warp(const SelStage& st, const Tree& t, int node, int lane)
{
    const size_t row = (size_t)node * AP;
    for (int c = lane; c < AP / 4; c += 32) {
        cp_async16(st.vis + 4 * c, t.cvis + row + 4 * c);
        cp_async16(st.vl + 4 * c, t.cvl + row + 4 * c);
        cp_async16(st.vsum + 4 * c, t.cvsum + row + 4 * c);
        cp_async16(st.cidx + 4 * c, t.cidx + row + 4 * c);
    }
    for (int c = lane; c < AP / 2; c += 32) cp_async16(st.pol + 2 * c, t.cpol + row + 2 * c);
    for (int c = lane; c < AP / 8; c += 32) cp_async16(st.action + 8 * c, t.action + row + 8 * c);
    if (lane < H_STRIDE / 4) cp_async16(st.hdr + 4 * lane, t.hdr + (size_t)node * H_STRIDE + 4 * lane);
}

// node.py:141-157 + pucb.py:8-29 on a staged node.  Returns the child index (warp-uniform); the float64 arithmetic runs in
// the reference's order (lowest index first on ties).
// sqtab (optional): sqrt((double)n) for n < nsqrt in shared memory -- a float64 square root is a ~450-cycle dependent chain
// that every lane would repeat at every ply.
__device__ inline int select_puct(const SelStage& st, bool cgos, int lane, const double* sqtab = nullptr, int nsqrt = 0)
{
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    const int k = st.hdr[H_K];
    const int nsq = st.hdr[H_NV] + st.hdr[H_VL] + 1;
    const double sq = nsq < nsqrt ? sqtab[nsq] : sqrt((double)nsq);
    double bv = 0.0; int bi = 0x7fffffff;
    for (int i = lane; i < k; i += 32) {          // (unrolling for overlapping divisions was measured slower: 7.4 k -> 10.6 k cycles)
        const int cv = st.vis[i] + st.vl[i];
        // an unvisited child (n + vl = 0) divides by one and has q = 0: x / 1.0 = x and 0.0 + x = x exactly, so the two
        // correctly rounded divisions are only executed by lanes whose child has been visited (most sweeps of a deep,
        // wide node skip them altogether)
        const double num = dmul(dmul(1.0, st.pol[i]), sq);
        double v = num;
        if (cv != 0) {
            // zero numerators (virtual losses without a finished visit; an underflowed prior) would take the division's slow
            // path; 0 / x = +0 exactly
            const float vs = st.vsum[i];
            const double q = vs == 0.0f ? 0.0 : ddiv((double)vs, (double)cv);
            const double u = num == 0.0 ? 0.0 : ddiv(num, (double)(cv + 1));
            v = dadd(q, u);
        }
        if (cgos && i == k - 1) v = dsub(v, 0.1);
        if (bi == 0x7fffffff || v > bv) { bv = v; bi = i; }
    }
    __syncwarp();
    warp_argmax_d(bv, bi);
    return bi;
}

// node.py:324-346
template <int AP>
__device__ inline int select_sh_root(const Tree& t, int node, int thr, int lane)
{
    const int k = t.hdr[(size_t)node * H_STRIDE + H_K];
    const size_t row = (size_t)node * AP;
    int mx = 0;
    for (int i = lane; i < k; i += 32) mx = max(mx, t.cvis[row + i]);
    mx = warp_max_i(mx);
    const double sigma = dmul((double)(C_VISIT + mx), C_SCALE);
    double bv = 0.0; int bi = 0x7fffffff;
    for (int i = lane; i < k; i += 32) {
        const int vis = t.cvis[row + i];
        const int cnt = vis + t.cvl[row + i];
        const double q = vis > 0 ? ddiv((double)t.cvsum[row + i], (double)vis) : 0.0;
        const double v = cnt >= thr ? -10000.0 : dadd(dadd(t.cpol[row + i], t.noise[i]), dmul(sigma, q));
        if (bi == 0x7fffffff || v > bv) { bv = v; bi = i; }
    }
    warp_argmax_d(bv, bi);
    return bi;
}

// nn/utility.py:125-136 on a shared-memory vector (in place allowed: out may alias in)
__device__ inline void softmax_smem(const double* in, double* out, int k, int lane)
{
    double mx = -1.0e300;
    for (int i = lane; i < k; i += 32) mx = fmax(mx, in[i]);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) mx = fmax(mx, shfl_xor_d(mx, o));
    for (int i = lane; i < k; i += 32) out[i] = det_exp(dsub(in[i], mx));
    __syncwarp();
    const double s = np_sum(out, k, lane);
    __syncwarp();
    for (int i = lane; i < k; i += 32) out[i] = ddiv(out[i], s);
    __syncwarp();
}

// node.py:281-321.  s0, s1: two shared-memory vectors of >= k doubles; the result is left in s0.
template <int AP>
__device__ inline void improved_policy(const Tree& t, int node, double* s0, double* s1, int lane)
{
    const int* h = t.hdr + (size_t)node * H_STRIDE;
    const int k = h[H_K], nv = h[H_NV];
    const float raw = __int_as_float(h[H_RAW]);
    const size_t row = (size_t)node * AP;
    int mx = 0;
    for (int i = lane; i < k; i += 32) mx = max(mx, t.cvis[row + i]);
    mx = warp_max_i(mx);
    const double sigma = dmul((double)(C_VISIT + mx), C_SCALE);
    for (int i = lane; i < k; i += 32) s1[i] = t.cpol[row + i];
    __syncwarp();
    softmax_smem(s1, s0, k, lane);                                   // s0 = softmax(prior)
    const double sum_prob = np_sum(s0, k, lane);
    for (int i = lane; i < k; i += 32) {
        const int vis = t.cvis[row + i];
        const double q = vis > 0 ? ddiv((double)t.cvsum[row + i], (double)vis) : 0.0;
        s1[i] = dmul(s0[i], q);
    }
    __syncwarp();
    const double v_pi = np_sum(s1, k, lane);
    const double vmix = ddiv(dadd(dmul((double)raw, 1.0), ddiv(dmul((double)nv, v_pi), sum_prob)), dadd((double)nv, 1.0));
    __syncwarp();
    for (int i = lane; i < k; i += 32) {
        const int vis = t.cvis[row + i];
        const double cq = vis > 0 ? ddiv((double)t.cvsum[row + i], (double)vis) : vmix;
        s1[i] = dadd(t.cpol[row + i], dmul(sigma, cq));
    }
    __syncwarp();
    softmax_smem(s1, s0, k, lane);
}

// node.py:349-361
template <int AP>
__device__ inline int select_sh_node(const Tree& t, int node, double* s0, double* s1, int lane)
{
    improved_policy<AP>(t, node, s0, s1, lane);
    const int* h = t.hdr + (size_t)node * H_STRIDE;
    const int k = h[H_K];
    const double den = dadd(1.0, (double)h[H_NV]);
    const size_t row = (size_t)node * AP;
    double bv = 0.0; int bi = 0x7fffffff;
    for (int i = lane; i < k; i += 32) {
        const double v = dsub(s0[i], ddiv((double)t.cvis[row + i], den));
        if (bi == 0x7fffffff || v > bv) { bv = v; bi = i; }
    }
    warp_argmax_d(bv, bi);
    __syncwarp();
    return bi;
}

// node.py:169-175 (first index of the maximum visit count)
template <int AP>
__device__ inline int best_visit_child(const Tree& t, int node, int lane)
{
    const int k = t.hdr[(size_t)node * H_STRIDE + H_K];
    const size_t row = (size_t)node * AP;
    double bv = 0.0; int bi = 0x7fffffff;
    for (int i = lane; i < k; i += 32) {
        const double v = (double)t.cvis[row + i];
        if (bi == 0x7fffffff || v > bv) { bv = v; bi = i; }
    }
    warp_argmax_d(bv, bi);
    return bi;
}

// node.py:364-375
template <int AP>
__device__ __forceinline__ double value_evaluation(const Tree& t, int node, int child)
{
    const size_t row = (size_t)node * AP;
    const int vis = t.cvis[row + child];
    return vis == 0 ? 0.5 : ddiv((double)t.cvsum[row + child], (double)vis);
}

// mcts/sequential_halving.py:7-60: {num_considered -> rounds}, insertion ordered.  Scalar code (lane 0).
__device__ inline int sh_schedule(int m, int visits, int* considered, int* counts, int cap)
{
    int np = 0;
    if (m <= 1) { considered[0] = 1; counts[0] = visits; return visits > 0 ? 1 : 0; }
    const int log2max = m <= 2 ? 1 : (m <= 4 ? 2 : (m <= 8 ? 3 : (m <= 16 ? 4 : 5)));
    int total = 0, nc = m;
    while (total < visits) {
        int extra = visits / (log2max * nc);
        if (extra < 1) extra = 1;
        for (int e = 0; e < extra && total < visits; e++) {
            const int c = min(nc, visits - total);
            total += c;
            int f = -1;
            for (int j = 0; j < np; j++) if (considered[j] == c) f = j;
            if (f >= 0) counts[f]++;
            else if (np < cap) { considered[np] = c; counts[np] = 1; np++; }
        }
        nc = max(2, nc / 2);
    }
    return np;
}

}  // namespace tg
