// tg_block.cuh -- block-per-game PUCT search: one CTA of NT threads walks one game's tree.
//
// Reference path: MCTSTree.search / search_mcts (mcts/tree.py:130-174, 199-244), select_next_action
// (mcts/node.py:141-157, mcts/pucb/pucb.py:8-29), expand_node (tree.py:247-270), process_mini_batch (tree.py:273-315),
// TimeManager.is_move_decided (mcts/time_manager.py:146-163), GoBoard.put_stone / is_legal (board/go_board.py:131-304).
//
// Why a block: descents of one PUCT batch are sequentially dependent (each one changes the virtual losses the next
// selection reads), so the only parallelism inside a game is INSIDE a ply -- the 362-child selection sweep, the
// board sweeps of put_stone, the 361-point legality analysis of an expansion, the 362-prior Dirichlet draw.  The
// warp-per-game kernels (tg_search.cuh) run those 32 wide; with few games per GPU (BASELINE configs[3]: 1024 games,
// configs[4]: ONE game) almost every warp slot of the machine is idle and a move is latency bound.  Here the same
// steps run NT wide (NT = 256: 1024 games are one resident wave of 8 CTAs per SM) with block barriers in place of warp
// barriers.  Results are bit-identical to the warp kernels: same float64 operations per child, same first-index
// tie break, same summation orders (the Dirichlet normaliser keeps the "warp shape" of expand_node in tg_search.cuh).
//
// The backup of a batch is split the same way: priors of all evaluated nodes are scattered by the whole block (they
// are independent), then one warp applies the values leaf by leaf in queue order (fp32 sums are order dependent).
#pragma once
#include "tg_search.cuh"

namespace tg {

template <int N, int NT> struct BlkSmem {
    using G = Geo<N>;
    static constexpr int NW = NT / 32;
    static constexpr int CH = (G::NN + NT - 1) / NT;      // analysis points per thread
    // child rows of a node staged from the HBM/L2 node pool; two buffers: the rows of the node a descent moves to are
    // fetched (cp.async) while put_stone runs, the root's rows for the next descent while the leaf is expanded
    struct NodeStage {
        alignas(16) double pol[G::AP];
        alignas(16) int vis[G::AP]; int vl[G::AP]; float vsum[G::AP]; int cidx[G::AP];
        alignas(16) int16_t action[G::AP];
        alignas(16) int hdr[H_STRIDE];
    } st[2];
    alignas(16) uint32_t eye2[4096];                       // eye table, two bits per 3x3 code (pattern.py:53-98)
    alignas(16) u64 zob[4 * G::CP];                        // Zobrist keys (zobrist_hash.py:9-10): every put_stone reads one
    WBoard<N> root;
    WBoard<N> scratch;
    alignas(16) WAnalysis<N> an;                           // expansion scratch
    alignas(16) double s0[G::AP];                          // expansion: super-ko hit hashes, then -log u
    double s1[G::AP];                                      // expansion: points of the super-ko hits (int16)
    // reductions (ping-pong: one barrier per reduction)
    u64 r_x[2][NW]; int r_i[2][NW]; int r_j[2][NW]; double r_d[2][NW];
    int wcnt[NW];
    int nhit;
    double bc_d;
    uint8_t flag[(G::NN + 15) & ~15];                      // expansion: candidate flags / super-ko verdicts per point
};

template <int NT> struct Blk {
    int tid, lane, warp, ph;
    __device__ __forceinline__ Blk() : tid(threadIdx.x), lane(threadIdx.x & 31), warp(threadIdx.x >> 5), ph(0) {}
    __device__ __forceinline__ void sync() const { __syncthreads(); }
};

// ---- block reductions: REDUX / shuffle inside a warp, one shared slot per warp, then EVERY warp folds the NW slots with
// its lanes in parallel (lane l reads slot l, one more REDUX) -- no second barrier, and no 16-step serial fold per thread
// (with 512 threads the redundant serial folds of 16 warps on 4 schedulers were the larger part of a selection).
template <int N, int NT>
__device__ __forceinline__ void blk_xor_sum(BlkSmem<N, NT>& sm, Blk<NT>& k, u64& x, int& cnt)
{
    constexpr int NW = NT / 32;
    x = warp_xor64(x); cnt = warp_sum_i(cnt);
    const int p = k.ph; k.ph ^= 1;
    if (k.lane == 0) { sm.r_x[p][k.warp] = x; sm.r_i[p][k.warp] = cnt; }
    k.sync();
    const u64 xs = k.lane < NW ? sm.r_x[p][k.lane] : 0ull;
    const int cs = k.lane < NW ? sm.r_i[p][k.lane] : 0;
    x = ((u64)__reduce_xor_sync(0xffffffffu, (unsigned)(xs >> 32)) << 32) | __reduce_xor_sync(0xffffffffu, (unsigned)xs);
    cnt = __reduce_add_sync(0xffffffffu, cs);
}
template <int N, int NT>
__device__ __forceinline__ int blk_max_i(BlkSmem<N, NT>& sm, Blk<NT>& k, int v)
{
    constexpr int NW = NT / 32;
    v = __reduce_max_sync(0xffffffffu, v);
    const int p = k.ph; k.ph ^= 1;
    if (k.lane == 0) sm.r_i[p][k.warp] = v;
    k.sync();
    return __reduce_max_sync(0xffffffffu, k.lane < NW ? sm.r_i[p][k.lane] : (int)0x80000000);
}
template <int N, int NT>
__device__ __forceinline__ void blk_sum_max(BlkSmem<N, NT>& sm, Blk<NT>& k, int& s, int& mx)
{
    constexpr int NW = NT / 32;
    s = __reduce_add_sync(0xffffffffu, s); mx = __reduce_max_sync(0xffffffffu, mx);
    const int p = k.ph; k.ph ^= 1;
    if (k.lane == 0) { sm.r_i[p][k.warp] = s; sm.r_j[p][k.warp] = mx; }
    k.sync();
    s = __reduce_add_sync(0xffffffffu, k.lane < NW ? sm.r_i[p][k.lane] : 0);
    mx = __reduce_max_sync(0xffffffffu, k.lane < NW ? sm.r_j[p][k.lane] : (int)0x80000000);
}
// argmax with numpy semantics (first index wins ties); idx = INT_MAX marks an empty thread.  Warp stage: three REDUX
// instructions on order-preserving integer keys (tg_common.cuh warp_argmax_key); block stage: the same on the NW slots.
template <int N, int NT>
__device__ __forceinline__ int blk_argmax_d(BlkSmem<N, NT>& sm, Blk<NT>& k, double v, int idx)
{
    constexpr int NW = NT / 32;
    u64 key = idx != 0x7fffffff ? order_key(v) : 0ull;
    warp_argmax_key(key, idx);
    const int p = k.ph; k.ph ^= 1;
    if (k.lane == 0) { sm.r_x[p][k.warp] = key; sm.r_i[p][k.warp] = idx; }
    k.sync();
    u64 k2 = k.lane < NW ? sm.r_x[p][k.lane] : 0ull;
    int i2 = k.lane < NW ? sm.r_i[p][k.lane] : 0x7fffffff;
    warp_argmax_key(k2, i2);
    return i2;
}
template <int N, int NT>
__device__ __forceinline__ int blk_min_i(BlkSmem<N, NT>& sm, Blk<NT>& k, int v)
{
    constexpr int NW = NT / 32;
    v = __reduce_min_sync(0xffffffffu, v);
    const int p = k.ph; k.ph ^= 1;
    if (k.lane == 0) sm.r_i[p][k.warp] = v;
    k.sync();
    return __reduce_min_sync(0xffffffffu, k.lane < NW ? sm.r_i[p][k.lane] : 0x7fffffff);
}

// ---- board ------------------------------------------------------------------------------------------------------
template <int N, int NT> __device__ inline void bb_recount(WBoard<N>& b, const Blk<NT>& k)
{
    using G = Geo<N>;
    for (int c = k.tid; c < G::CP; c += NT) b.ls[c] = 0;
    k.sync();
    for (int c = k.tid; c < G::CELLS; c += NT) {
        const int col = b.color[c];
        if (col == BLACK || col == WHITE) atomicAdd(&b.ls[b.chain[c]], 1u);
        else if (col == EMPTY) {
            const int q[4] = { c - G::W, c - 1, c + 1, c + G::W };
            int seen[4], ns = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int cc = b.color[q[i]];
                if (cc != BLACK && cc != WHITE) continue;
                const int l = b.chain[q[i]];
                bool dup = false;
                for (int j = 0; j < ns; j++) dup |= (seen[j] == l);
                if (!dup) { seen[ns++] = l; atomicAdd(&b.ls[l], 1u << 16); }
            }
        }
    }
    k.sync();
}

template <int N, int NT>
__device__ inline void bb_load(WBoard<N>& b, BScal& s, const BoardPool<N>& pool, int g, const Blk<NT>& k)
{
    using G = Geo<N>;
    const uint32_t* c4 = reinterpret_cast<const uint32_t*>(pool.color + (size_t)g * G::CP);
    uint32_t* d4 = reinterpret_cast<uint32_t*>(b.color);
    for (int i = k.tid; i < G::CP / 4; i += NT) d4[i] = c4[i];
    const uint32_t* h2 = reinterpret_cast<const uint32_t*>(pool.chain + (size_t)g * G::CP);
    uint32_t* e2 = reinterpret_cast<uint32_t*>(b.chain);
    for (int i = k.tid; i < G::CP / 2; i += NT) e2[i] = h2[i];
    for (int i = k.tid; i < BLOOM_WORDS; i += NT) b.bloom[i] = pool.bloom[(size_t)g * BLOOM_WORDS + i];
    const int* sc = pool.scal + (size_t)g * 8;
    s.hash = pool.hash[g];
    s.moves = sc[0]; s.ko_pos = sc[1]; s.ko_move = sc[2]; s.pris0 = sc[3]; s.pris1 = sc[4];
    k.sync();
    bb_recount<N, NT>(b, k);
}

template <int N, int NT> __device__ inline void bb_copy(WBoard<N>& dst, const WBoard<N>& src, const Blk<NT>& k)
{
    static_assert(sizeof(WBoard<N>) % 4 == 0, "word copy");
    const uint32_t* a = reinterpret_cast<const uint32_t*>(&src);
    uint32_t* d = reinterpret_cast<uint32_t*>(&dst);
    for (int i = k.tid; i < (int)(sizeof(WBoard<N>) / 4); i += NT) d[i] = a[i];
    k.sync();
}

// GoBoard.put_stone (go_board.py:131-185); same steps as wb_put_stone, sweeps NT wide.  The common case (nothing
// captured, at most one own string extended) costs two barriers: every thread reads the neighbourhood, thread 0 applies
// the move (stone, label, local liberty update, history, Bloom bit), every thread re-reads what the ko rule needs.
template <int N, int NT>
__device__ inline void bb_put_stone(BlkSmem<N, NT>& sm, WBoard<N>& b, BScal& s, int pos, int color, const u64* __restrict__ zob,
                                    u64* hist_hash, int16_t* hist_pos, Blk<NT>& k)
{
    using G = Geo<N>;
    if (pos == PASS) {                                       // :138-141
        if (k.tid == 0 && s.moves < G::MAXREC) { hist_hash[s.moves] = s.hash; hist_pos[s.moves] = 0; }
        s.moves++;
        __threadfence_block();
        k.sync();
        return;
    }
    const int other = opp(color);
    const int q[4] = { pos - G::W, pos - 1, pos + 1, pos + G::W };
    int cap[4], ncap = 0, own[4], nown = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int cc = b.color[q[i]];
        if (cc != color && cc != other) continue;
        const int l = b.chain[q[i]];
        if (cc == color) {
            bool dup = false;
            for (int j = 0; j < nown; j++) dup |= (own[j] == l);
            if (!dup) own[nown++] = l;
        } else if ((b.ls[l] >> 16) == 1u) {
            bool dup = false;
            for (int j = 0; j < ncap; j++) dup |= (cap[j] == l);
            if (!dup) cap[ncap++] = l;
        }
    }
    const int label = nown > 0 ? own[0] : pos;
    s.hash ^= zob[color * G::CELLS + pos];
    k.sync();                                                // every thread has read the neighbourhood
    int prisoner = 0;
    if (ncap == 0 && nown <= 1) {
        // local liberty update (see wb_put_stone), one lane of warp 0 per neighbour: every distinct adjacent enemy string
        // loses the liberty pos; the stone's string loses pos, gains the empty neighbours that were not its liberties yet
        // and grows by one
        if (k.warp == 0) {
            bool gain = false;
            if (k.lane < 4) {
                const int qi = q[k.lane];
                const int cc = b.color[qi];
                if (cc == other) {
                    const int l = b.chain[qi];
                    bool dup = false;
                    for (int j = 0; j < k.lane; j++) dup |= (b.color[q[j]] == other && b.chain[q[j]] == l);
                    if (!dup) b.ls[l] -= 1u << 16;           // distinct labels per lane: no two lanes touch the same word
                } else if (cc == EMPTY) {
                    bool already = false;
                    if (nown == 1) {
                        const int r[4] = { qi - G::W, qi - 1, qi + 1, qi + G::W };
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            already |= (r[j] != pos && b.color[r[j]] == color && b.chain[r[j]] == label);
                    }
                    gain = !already;
                }
            }
            const unsigned gained = (unsigned)__popc(__ballot_sync(0xffffffffu, gain));
            if (k.lane == 0) {
                b.color[pos] = (uint8_t)color; b.chain[pos] = (uint16_t)label;
                if (nown == 1) b.ls[label] += (gained << 16) - (1u << 16) + 1u;
                else b.ls[label] = (gained << 16) | 1u;
                if (s.moves < G::MAXREC) { hist_hash[s.moves] = s.hash; hist_pos[s.moves] = (int16_t)pos; }
                const unsigned bit = bloom_bit(s.hash);
                b.bloom[bit >> 5] |= 1u << (bit & 31);
                __threadfence_block();
            }
        }
        k.sync();
        s.moves++;
        return;                                              // no prisoners: the ko rule (:173-177) cannot apply
    }
    if (k.tid == 0) { b.color[pos] = (uint8_t)color; b.chain[pos] = (uint16_t)label; }
    if (ncap == 0) {
        // merge without capture: one sweep relabels the absorbed strings and recounts the merged string's liberties
        // (see wb_put_stone); sizes add up; adjacent enemy strings lose pos
        unsigned size = 1;
        for (int j = 0; j < nown; j++) size += b.ls[own[j]] & 0xffffu;
        u64 none = 0; int cnt = 0;
        for (int c = k.tid; c < G::CELLS; c += NT) {
            const int cc = c == pos ? color : b.color[c];    // (thread 0's store of the new stone may not have landed yet)
            if (cc == color && c != pos) {
                const int l = b.chain[c];
                bool hit = false;
                for (int j = 1; j < nown; j++) hit |= (own[j] == l);
                if (hit) b.chain[c] = (uint16_t)label;
            } else if (cc == EMPTY) {
                const int r[4] = { c - G::W, c - 1, c + 1, c + G::W };
                bool adj = false;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (r[j] == pos) { adj = true; continue; }
                    if (b.color[r[j]] != color) continue;
                    // (another thread may be relabelling this stone right now: compute-sanitizer racecheck reports the pair.
                    //  Benign by construction: the 16-bit label is read either as the absorbed label or as `label`, and both
                    //  are members of own[].)
                    const int l = b.chain[r[j]];
                    for (int m = 0; m < nown; m++) adj |= (own[m] == l);
                }
                cnt += adj;
            }
        }
        blk_xor_sum<N, NT>(sm, k, none, cnt);                // (contains a barrier: all reads of ls[own[]] are done)
        if (k.tid == 0) {
            b.ls[label] = ((unsigned)cnt << 16) | size;
            int el[4], ne = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if (b.color[q[i]] != other) continue;
                const int l = b.chain[q[i]];
                bool dup = false;
                for (int j = 0; j < ne; j++) dup |= (el[j] == l);
                if (!dup) { el[ne++] = l; b.ls[l] -= 1u << 16; }
            }
            if (s.moves < G::MAXREC) { hist_hash[s.moves] = s.hash; hist_pos[s.moves] = (int16_t)pos; }
            const unsigned bit = bloom_bit(s.hash);
            b.bloom[bit >> 5] |= 1u << (bit & 31);
            __threadfence_block();
        }
        s.moves++;
        k.sync();
        return;                                              // no prisoners: the ko rule (:173-177) cannot apply
    }
    {
        u64 hx = 0; int cnt = 0;
        for (int c = k.tid; c < G::CELLS; c += NT) {
            const int cc = b.color[c];
            if (cc == other) {
                const int l = b.chain[c];
                bool hit = false;
                for (int j = 0; j < ncap; j++) hit |= (cap[j] == l);
                if (hit) { b.color[c] = EMPTY; hx ^= zob[other * G::CELLS + c]; cnt++; }
            } else if (cc == color && c != pos) {
                const int l = b.chain[c];
                bool hit = false;
                for (int j = 1; j < nown; j++) hit |= (own[j] == l);
                if (hit) b.chain[c] = (uint16_t)label;
            }
        }
        blk_xor_sum<N, NT>(sm, k, hx, cnt);                  // (contains a barrier)
        s.hash ^= hx;
        prisoner = cnt;
    }
    if (color == BLACK) s.pris0 += prisoner; else s.pris1 += prisoner;
    bb_recount<N, NT>(b, k);                                 // its first barrier orders the sweep above before the recount
    if (nown == 0 && prisoner == 1 && (b.ls[label] >> 16) == 1u) {       // :173-177
        s.ko_move = s.moves;
#pragma unroll
        for (int i = 0; i < 4; i++) if (b.color[q[i]] == EMPTY) s.ko_pos = q[i];
    }
    if (k.tid == 0) {
        if (s.moves < G::MAXREC) { hist_hash[s.moves] = s.hash; hist_pos[s.moves] = (int16_t)pos; }
        const unsigned bit = bloom_bit(s.hash);
        b.bloom[bit >> 5] |= 1u << (bit & 31);
        __threadfence_block();
    }
    s.moves++;
    k.sync();
}

// ---- node staging: one burst of 16-byte asynchronous copies per node ---------------------------------------------------
template <int N, int NT>
__device__ __forceinline__ void stage_node(typename BlkSmem<N, NT>::NodeStage& st, const Tree& t, int node, const Blk<NT>& k)
{
    constexpr int AP = Geo<N>::AP;
    const size_t row = (size_t)node * AP;
    for (int c = k.tid; c < AP / 4; c += NT) {
        cp_async16(st.vis + 4 * c, t.cvis + row + 4 * c);
        cp_async16(st.vl + 4 * c, t.cvl + row + 4 * c);
        cp_async16(st.vsum + 4 * c, t.cvsum + row + 4 * c);
        cp_async16(st.cidx + 4 * c, t.cidx + row + 4 * c);
    }
    for (int c = k.tid; c < AP / 2; c += NT) cp_async16(st.pol + 2 * c, t.cpol + row + 2 * c);
    for (int c = k.tid; c < AP / 8; c += NT) cp_async16(st.action + 8 * c, t.action + row + 8 * c);
    if (k.tid < H_STRIDE / 4) cp_async16(st.hdr + 4 * k.tid, t.hdr + (size_t)node * H_STRIDE + 4 * k.tid);
}
__device__ __forceinline__ void stage_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---- selection (node.py:141-157 + pucb.py:8-29) on a staged node: one child per thread -----------------------------------
template <int N, int NT>
__device__ inline int select_puct_blk(BlkSmem<N, NT>& sm, const typename BlkSmem<N, NT>::NodeStage& st, bool cgos, Blk<NT>& k, long long* prof = nullptr)
{
    long long c0 = prof ? clock64() : 0;
    const int nk = st.hdr[H_K];
    const double sq = sqrt((double)(st.hdr[H_NV] + st.hdr[H_VL] + 1));
    double bv = 0.0; int bi = 0x7fffffff;
    for (int i = k.tid; i < nk; i += NT) {
        const int cv = st.vis[i] + st.vl[i];
        const double num = dmul(dmul(1.0, st.pol[i]), sq);
        double v = num;
        if (cv != 0) {
            // a zero numerator (an edge with virtual losses but no finished visit yet: value sum 0; a prior that underflowed)
            // sends the correctly rounded division down its slow path (~10x the cycles); 0 / x = +0 exactly, so skip it
            const float vs = st.vsum[i];
            const double q = vs == 0.0f ? 0.0 : ddiv((double)vs, (double)cv);
            const double u = num == 0.0 ? 0.0 : ddiv(num, (double)(cv + 1));
            v = dadd(q, u);
        }
        if (cgos && i == nk - 1) v = dsub(v, 0.1);
        if (bi == 0x7fffffff || v > bv) { bv = v; bi = i; }
    }
    if (prof) { const long long c = clock64(); prof[8] += c - c0; c0 = c; }
    const int r = blk_argmax_d<N, NT>(sm, k, bv, bi);
    if (prof) { const long long c = clock64(); prof[9] += c - c0; }
    return r;
}

// Position-hash history seen by an expansion: entries below `split` are the game's own record (global), the rest the
// moves of the descent.  The inline descent keeps both in the game's record (split = 0); the deferred expansion of
// k_expand_leaves_blk keeps the descent's part private, because several CTAs replay leaves of the same game at once.
struct HistView {
    const u64* lo; const u64* hi; int split;
    __device__ __forceinline__ u64 operator[](int e) const { return e < split ? lo[e] : hi[e - split]; }
};

// ---- expansion (tree.py:247-270 + node.py:41-72) ------------------------------------------------------------------
// given_idx < 0: allocate the next node of the pool; otherwise fill the rows of a node the tree walk already allocated.
template <int N, int NT>
__device__ inline int expand_node_blk(BlkSmem<N, NT>& sm, const Dev& D, const Tree& t, int g, int* gs, const WBoard<N>& b, const BScal& s,
                                      int color, const HistView hist_hash, unsigned move_key, Blk<NT>& k, int given_idx = -1)
{
    using G = Geo<N>;
    constexpr int CH = BlkSmem<N, NT>::CH, NW = NT / 32;
    const int idx = given_idx >= 0 ? given_idx : gs[GS_NNODES];
    if (idx >= D.tree.max_nodes) { k.sync(); if (k.tid == 0) gs[GS_ERROR] |= ERR_NODES; k.sync(); return -1; }
    const size_t row = (size_t)idx * G::AP;
    const bool superko = D.superko != 0;
    // per-string liberty extremes and key XORs (wb_prepare_analysis, NT wide)
    WAnalysis<N>& an = sm.an;
    for (int c = k.tid; c < G::CP; c += NT) { an.lmin[c] = 0xffffu; an.lmax[c] = 0; an.cx[c] = 0; }
    if (k.tid == 0) sm.nhit = 0;
    k.sync();
    {
        const int other = opp(color);
        for (int c = k.tid; c < G::CELLS; c += NT) {
            const int col = b.color[c];
            if (col == EMPTY) {
                const int q[4] = { c - G::W, c - 1, c + 1, c + G::W };
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int cc = b.color[q[i]];
                    if (cc == BLACK || cc == WHITE) { const int l = b.chain[q[i]]; atomicMin(&an.lmin[l], (unsigned)c); atomicMax(&an.lmax[l], (unsigned)c); }
                }
            } else if (superko && (col == BLACK || col == WHITE)) {
                const int l = b.chain[c];
                if ((b.ls[l] >> 16) == 1u) atomicXor(&an.cx[l], sm.zob[other * G::CELLS + c]);
            }
        }
    }
    k.sync();
    // lane-local status of every point; super-ko candidates whose hash hits the Bloom filter go to a hit list
    u64* hit_h = reinterpret_cast<u64*>(sm.s0);
    int16_t* hit_pt = reinterpret_cast<int16_t*>(sm.s1);
    bool cand[CH];
#pragma unroll
    for (int c = 0; c < CH; c++) {
        const int pi = c * NT + k.tid;
        cand[c] = false;
        if (pi < G::NN) {
            const PointStatus st = wb_point_status<N>(b, an, s, onboard_pos<N>(pi), color, superko, sm.zob, EyeLutPacked{sm.eye2});
            cand[c] = st.legal_pre && st.satari < 7 && !st.eye;                     // tree.py:261-263 (legality completed below)
            sm.flag[pi] = 0;
            if (cand[c] && st.need_scan) { const int hI = atomicAdd(&sm.nhit, 1); hit_h[hI] = st.h; hit_pt[hI] = (int16_t)pi; }
        }
    }
    k.sync();
    const int nh = sm.nhit;
    if (nh > 0) {                                            // exact history scan (record.py:54-63), all (hit, entry) pairs in parallel
        const int lim = (s.moves < G::MAXREC ? s.moves : G::MAXREC) - 1;             // live entries 1 .. lim
        for (int j = k.tid; j < nh * lim; j += NT) {
            const int hI = j / lim, e = 1 + (j - hI * lim);
            if (hist_hash[e] == hit_h[hI]) sm.flag[hit_pt[hI]] = 1;
        }
        k.sync();
    }
    // ordered compaction of the candidates (raster order), PASS last (tree.py:264)
    int total = 0;
#pragma unroll
    for (int c = 0; c < CH; c++) {
        const int pi = c * NT + k.tid;
        const bool ok = cand[c] && pi < G::NN && !sm.flag[pi];
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (k.lane == 0) sm.wcnt[k.warp] = __popc(m);
        k.sync();
        const int wv = k.lane < NW ? sm.wcnt[k.lane] : 0;                           // lane-parallel fold of the warp counts
        const int all = __reduce_add_sync(0xffffffffu, wv);
        const int before = __reduce_add_sync(0xffffffffu, k.lane < k.warp ? wv : 0);
        if (ok) t.action[row + total + before + __popc(m & ((1u << k.lane) - 1))] = (int16_t)onboard_pos<N>(pi);
        total += all;
        k.sync();
    }
    const int nk = total + 1;
    if (k.tid == 0) t.action[row + total] = PASS;
    // get_tentative_policy (tree.py:509-519): Dirichlet(1,..,1) from the counter-based stream; the normaliser is summed
    // in the warp shape of expand_node (lane l adds entries l, l+32, ... in order, then an xor butterfly)
    const u64 gid = D.game_id[g];
    for (int i = k.tid; i < nk; i += NT) sm.s0[i] = dsub(0.0, det_log(noise_u(D.seed, gid, move_key, (unsigned)idx, 0u, (unsigned)i)));
    k.sync();
    if (k.warp == 0) {
        double part = 0.0;
        for (int i = k.lane; i < nk; i += 32) part = dadd(part, sm.s0[i]);
        const double sum = warp_shape_sum(part);
        if (k.lane == 0) sm.bc_d = sum;
    }
    k.sync();
    const double sum = sm.bc_d;
    for (int i = k.tid; i < G::AP; i += NT) {
        if (i < nk) t.cpol[row + i] = ddiv(sm.s0[i], sum); else { t.cpol[row + i] = 0.0; t.action[row + i] = 0; }
        t.cidx[row + i] = NOT_EXPANDED; t.cval[row + i] = 0.0f; t.cvis[row + i] = 0; t.cvl[row + i] = 0; t.cvsum[row + i] = 0.0f;
    }
    if (k.tid < H_STRIDE) t.hdr[(size_t)idx * H_STRIDE + k.tid] = (k.tid == H_K) ? nk : 0;
    if (k.tid == 0 && given_idx < 0) gs[GS_NNODES] = idx + 1;
    __threadfence_block();
    k.sync();
    return idx;
}

// ---- leaf queue (mcts/batch_data.py:18-27) ------------------------------------------------------------------------
template <int N, int NT>
__device__ inline void push_leaf_blk(BlkSmem<N, NT>& sm, const Dev& D, int g, int* gs, const WBoard<N>& b, const BScal& s, int color,
                                     const unsigned* cur_path, int plen, int node_index, Blk<NT>& k)
{
    const int i = gs[GS_NLEAF];
    if (i >= D.cap) { k.sync(); if (k.tid == 0) gs[GS_ERROR] |= ERR_QUEUE; k.sync(); return; }
    const size_t q = (size_t)g * D.cap;
    int dup_of = -1;
    if (D.dedup) {                                           // identical path among the earlier entries of this batch
        int first = 0x7fffffff;
        for (int j = k.tid; j < i; j += NT) {
            if (D.path_len[q + j] != plen) continue;
            const unsigned* pj = D.path + (q + j) * D.max_depth;
            bool same = true;
            for (int d = plen - 1; d >= 0 && same; d--) same = (pj[d] == cur_path[d]);   // paths differ near the leaf first
            if (same) { first = j; break; }
        }
        first = blk_min_i<N, NT>(sm, k, first);
        if (first != 0x7fffffff) dup_of = first;
    }
    int nu = gs[GS_NUNIQ], slot;
    if (dup_of >= 0) slot = D.leaf_slot[q + dup_of];
    else {
        slot = nu++;
        uint8_t* dst = D.snap + (q + slot) * Snap<N>::BYTES;
        using G = Geo<N>;
        const int prev = (s.moves - 1 < G::MAXREC) ? D.hist_pos[(size_t)g * G::MAXREC + s.moves - 1] : 0;
        if (k.tid == 0) {
            const int pidx = (prev == PASS) ? -1 : ((prev % G::W) - 1) + ((prev / G::W) - 1) * N;
            *reinterpret_cast<int16_t*>(dst) = (int16_t)pidx;
            dst[2] = (s.moves > 1 && prev == PASS) ? 1 : 0;
            dst[3] = (uint8_t)color;
        }
        for (int idx = k.tid; idx < G::NN; idx += NT) dst[Snap<N>::HDR + idx] = b.color[onboard_pos<N>(idx)];
    }
    k.sync();
    if (k.tid == 0) {
        D.path_len[q + i] = plen; D.leaf_node[q + i] = node_index; D.leaf_slot[q + i] = slot;
        gs[GS_NLEAF] = i + 1; gs[GS_NUNIQ] = nu;
    }
    __threadfence_block();
    k.sync();
}

// ---- up to `batch` descents of search_mcts for one game ----------------------------------------------------------
template <int N, int NT>
__global__ void __launch_bounds__(NT) k_descend_puct_blk(Dev D, const uint32_t* __restrict__ eye2, int visits, int batch, int strict)
{
    using G = Geo<N>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BlkSmem<N, NT>& sm = *reinterpret_cast<BlkSmem<N, NT>*>(smem_raw);
    Blk<NT> k;
    const int g = blockIdx.x;
    int* gs = D.gs + (size_t)g * GS_STRIDE;
    const bool idle = !gs[GS_ACTIVE] || gs[GS_FINISHED] || gs[GS_ERROR] || gs[GS_DONE];
    k.sync();
    if (k.tid == 0) { gs[GS_NLEAF] = 0; gs[GS_NUNIQ] = 0; }
    __threadfence_block();
    k.sync();
    if (idle) return;
    const Tree t = tree_of<G::AP>(D.tree, g);
    stage_node<N, NT>(sm.st[0], t, 0, k);                    // the root's rows arrive while the board is loaded
    for (int i = k.tid; i < 4096 / 4; i += NT) reinterpret_cast<uint4*>(sm.eye2)[i] = reinterpret_cast<const uint4*>(eye2)[i];
    for (int i = k.tid; i < 4 * G::CELLS; i += NT) sm.zob[i] = D.zob[i];
    BScal rs;
    bb_load<N, NT>(sm.root, rs, pool_of<N>(D), g, k);
    const int root_color = gs[GS_COLOR];
    u64* hh = D.hist_hash + (size_t)g * G::MAXREC;
    int16_t* hp = D.hist_pos + (size_t)g * G::MAXREC;
    const unsigned move_key = (unsigned)rs.moves;
    const bool prof = D.prof && g == 0 && k.tid == 0;
    for (int bi = 0; bi < batch; bi++) {
        const int desc = gs[GS_DESC];
        if (desc >= visits) { k.sync(); if (k.tid == 0) gs[GS_DONE] = 1; break; }
        stage_wait();
        k.sync();                                            // root rows (buffer 0) are in shared memory
        if (desc > 0) {                                      // is_move_decided (time_manager.py:146-163)
            const int nk = sm.st[0].hdr[H_K];
            int top1 = 0;
            for (int i = k.tid; i < nk; i += NT) top1 = max(top1, sm.st[0].vis[i]);
            top1 = blk_max_i<N, NT>(sm, k, top1);
            int nmax = 0, top2 = 0;
            for (int i = k.tid; i < nk; i += NT) { const int v = sm.st[0].vis[i]; if (v == top1) nmax++; else top2 = max(top2, v); }
            blk_sum_max<N, NT>(sm, k, nmax, top2);
            if (nmax >= 2) top2 = top1;
            const int remaining = visits - sm.st[0].hdr[H_NV];
            const int cutoff = strict ? 0 : top1 - top2;
            if (remaining < cutoff) { k.sync(); if (k.tid == 0) gs[GS_DONE] = 1; break; }
        }
        long long pt0 = prof ? clock64() : 0;
        bb_copy<N, NT>(sm.scratch, sm.root, k);              // tree.py:147
        if (prof) { const long long c = clock64(); D.prof[0] += c - pt0; pt0 = c; }
        BScal s = rs;
        int color = root_color, cur = 0, plen = 0, buf = 0;
        unsigned* path = D.path + ((size_t)g * D.cap + gs[GS_NLEAF]) * D.max_depth;
        bool fail = false;
        for (;;) {
            const typename BlkSmem<N, NT>::NodeStage& st = sm.st[buf];
            const int next = select_puct_blk<N, NT>(sm, st, D.cgos != 0, k, prof ? D.prof : nullptr);         // :213
            if (prof) { const long long c = clock64(); D.prof[1] += c - pt0; pt0 = c; D.prof[5]++; }
            const size_t row = (size_t)cur * G::AP;
            const int mv = st.action[next];
            const int cv_before = st.vis[next] + st.vl[next];
            int ci = st.cidx[next];
            // the child's rows are needed next unless this edge ends the descent: fetch them under put_stone
            const bool spec = ci != NOT_EXPANDED && cv_before >= 1;
            if (spec) stage_node<N, NT>(sm.st[buf ^ 1], t, ci, k);
            if (k.tid == 0) {
                path[plen] = ((unsigned)cur << PATH_NODE_SHIFT) | (unsigned)next;
                t.hdr[(size_t)cur * H_STRIDE + H_VL] = st.hdr[H_VL] + 1; t.cvl[row + next] = st.vl[next] + 1;    // :221 add_virtual_loss
            }
            plen++;
            bb_put_stone<N, NT>(sm, sm.scratch, s, mv, color, sm.zob, hh, hp, k);     // :217
            if (prof) { const long long c = clock64(); D.prof[2] += c - pt0; pt0 = c; }
            color = opp(color);
            int expand_threshold = 1;
            if (s.moves > 2) {                               // :224-229
                if (s.moves - 1 >= G::MAXREC) { k.sync(); if (k.tid == 0) gs[GS_ERROR] |= ERR_HISTORY; fail = true; break; }
                if (hp[s.moves - 1] == PASS && hp[s.moves - 2] == PASS) expand_threshold = 10000000;
            }
            if (cv_before + 1 < expand_threshold + 1) {      // :231-241 (children_visits + children_virtual_loss after add_virtual_loss)
                if (spec) stage_wait();                      // (two-pass rule ended the descent: drain the unused fetch)
                k.sync();
                // root rows for the next descent are fetched under the expansion -- unless the expansion changes the root's
                // own child-index row (a depth-1 leaf): then the fetch has to follow that write
                const bool root_edge = cur == 0 && ci == NOT_EXPANDED;
                if (!root_edge) stage_node<N, NT>(sm.st[0], t, 0, k);
                if (ci == NOT_EXPANDED) {
                    if (prof) pt0 = clock64();
                    ci = expand_node_blk<N, NT>(sm, D, t, g, gs, sm.scratch, s, color, HistView{hh, hh, 0}, move_key, k);
                    if (prof) { const long long c = clock64(); D.prof[3] += c - pt0; pt0 = c; D.prof[6]++; }
                    if (ci < 0) { fail = true; break; }
                    if (k.tid == 0) { t.cidx[row + next] = ci; __threadfence_block(); }
                    if (root_edge) { k.sync(); stage_node<N, NT>(sm.st[0], t, 0, k); }
                }
                if (prof) pt0 = clock64();
                push_leaf_blk<N, NT>(sm, D, g, gs, sm.scratch, s, color, path, plen, ci, k);
                if (prof) { const long long c = clock64(); D.prof[4] += c - pt0; pt0 = c; D.prof[7]++; }
                break;
            }
            cur = ci; buf ^= 1;
            if (plen >= D.max_depth) { k.sync(); if (k.tid == 0) gs[GS_ERROR] |= ERR_DEPTH; fail = true; break; }
            if (prof) pt0 = clock64();
            stage_wait();
            k.sync();                                        // the child's rows are in shared memory
            if (prof) { const long long c = clock64(); D.prof[10] += c - pt0; pt0 = c; }
        }
        __threadfence_block();
        k.sync();
        if (fail) break;
        if (k.tid == 0) gs[GS_DESC] = desc + 1;
        __threadfence_block();
        k.sync();
    }
    stage_wait();
}

// =====================================================================================================================
// Deferred expansion (PUCT batches of more than one descent).
//
// Within a batch only the SELECTIONS are sequentially dependent (each one changes the virtual losses the next one reads).
// What a descent does with the board -- replaying its moves, the legality analysis of the new node, the leaf snapshot for
// the evaluator -- depends on the path alone, and nothing later in the batch reads the new node's rows unless a later
// descent walks INTO that node.  So the batch is split in two kernels:
//   k_walk_puct_blk      one CTA per game: selections only.  A leaf edge allocates the node index, links it into the
//                        parent row and queues (path, node, flags); no board is touched.  The rows of the nodes a launch
//                        visits stay in a small shared-memory cache kept coherent by write-through (everything a walk
//                        changes -- virtual losses, child links -- is written by this CTA), so hot nodes are fetched from
//                        L2 once per launch instead of once per visit.  If a descent does enter a node allocated earlier
//                        in the same launch, that node is materialised on the spot (replay + expansion, as the inline
//                        kernel does for every leaf).
//   k_expand_leaves_blk  one CTA per queued leaf, all leaves of all games at once: replay the path from the root board,
//                        fill the node's rows, write the evaluator snapshot.  With one game and 256-leaf batches this is
//                        where 147 otherwise idle SMs do the board work of the batch in the time of one leaf.
// Results are bit-identical to the inline kernel (same node numbering, same rows, same queue order).
// =====================================================================================================================
constexpr int WALK_MAX_BATCH = 1024;
enum : int { LEAF_EXPAND = 1, LEAF_SNAP = 2 };

template <int N, int NT> struct WalkSmem {
    BlkSmem<N, NT> b;
    int tag[16];                                           // node held by every cache slot (-1: none); slot 0 = root
    int16_t mv[Geo<N>::MAXREC + 2];                        // moves of the current descent
    int16_t newleaf[WALK_MAX_BATCH];                       // queue index of the n-th node allocated by this launch
    int16_t lslot[WALK_MAX_BATCH];                         // unique slot of every queued leaf
    uint8_t lflag[WALK_MAX_BATCH];                         // LEAF_* flags of every queued leaf
    unsigned mat[WALK_MAX_BATCH / 32];                     // nodes of this launch already materialised on demand
    // followed by (nc - 2) more NodeStage slots in dynamic shared memory
};

template <int N, int NT>
__device__ __forceinline__ void write_snap_blk(uint8_t* dst, const WBoard<N>& b, int prev, int moves, int color, const Blk<NT>& k)
{
    using G = Geo<N>;
    if (k.tid == 0) {
        const int pidx = (prev == PASS) ? -1 : ((prev % G::W) - 1) + ((prev / G::W) - 1) * N;
        *reinterpret_cast<int16_t*>(dst) = (int16_t)pidx;
        dst[2] = (moves > 1 && prev == PASS) ? 1 : 0;
        dst[3] = (uint8_t)color;
    }
    for (int idx = k.tid; idx < G::NN; idx += NT) dst[Snap<N>::HDR + idx] = b.color[onboard_pos<N>(idx)];
}

template <int N, int NT>
__global__ void __launch_bounds__(NT) k_walk_puct_blk(Dev D, const uint32_t* __restrict__ eye2, int visits, int batch, int strict, int nc)
{
    using G = Geo<N>;
    using Stage = typename BlkSmem<N, NT>::NodeStage;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WalkSmem<N, NT>& ws = *reinterpret_cast<WalkSmem<N, NT>*>(smem_raw);
    BlkSmem<N, NT>& sm = ws.b;
    Stage* extra = reinterpret_cast<Stage*>(smem_raw + ((sizeof(WalkSmem<N, NT>) + 15) & ~(size_t)15));
    auto slot_ptr = [&](int sl) -> Stage& { return sl < 2 ? sm.st[sl] : extra[sl - 2]; };
    Blk<NT> k;
    const int g = blockIdx.x;
    int* gs = D.gs + (size_t)g * GS_STRIDE;
    const bool idle = !gs[GS_ACTIVE] || gs[GS_FINISHED] || gs[GS_ERROR] || gs[GS_DONE];
    k.sync();
    if (k.tid == 0) { gs[GS_NLEAF] = 0; gs[GS_NUNIQ] = 0; }
    __threadfence_block();
    k.sync();
    if (idle) return;
    const Tree t = tree_of<G::AP>(D.tree, g);
    stage_node<N, NT>(sm.st[0], t, 0, k);                    // the root's rows arrive while the board is loaded
    for (int i = k.tid; i < 4096 / 4; i += NT) reinterpret_cast<uint4*>(sm.eye2)[i] = reinterpret_cast<const uint4*>(eye2)[i];
    for (int i = k.tid; i < 4 * G::CELLS; i += NT) sm.zob[i] = D.zob[i];
    if (k.tid < 16) ws.tag[k.tid] = k.tid == 0 ? 0 : -1;
    for (int i = k.tid; i < WALK_MAX_BATCH / 32; i += NT) ws.mat[i] = 0u;
    BScal rs;
    bb_load<N, NT>(sm.root, rs, pool_of<N>(D), g, k);
    const int root_color = gs[GS_COLOR];
    u64* hh = D.hist_hash + (size_t)g * G::MAXREC;
    int16_t* hp = D.hist_pos + (size_t)g * G::MAXREC;
    const unsigned move_key = (unsigned)rs.moves;
    const int root_last = (rs.moves >= 1 && rs.moves - 1 < G::MAXREC) ? hp[rs.moves - 1] : -1;
    const size_t q = (size_t)g * D.cap;
    const int nn0 = gs[GS_NNODES];
    int nnodes = nn0, nleaf = 0, nuniq = 0, desc = gs[GS_DESC], rr = 1;
    const bool prof = D.prof && g == 0 && k.tid == 0;
    stage_wait();
    k.sync();                                                // root rows (slot 0) are in shared memory
    bool decided = false;                                    // is_move_decided (time_manager.py:146-163): visit counts only
    {                                                        // change in the backup, so one evaluation serves the launch
        const int nk = sm.st[0].hdr[H_K];
        int top1 = 0;
        for (int i = k.tid; i < nk; i += NT) top1 = max(top1, sm.st[0].vis[i]);
        top1 = blk_max_i<N, NT>(sm, k, top1);
        int nmax = 0, top2 = 0;
        for (int i = k.tid; i < nk; i += NT) { const int v = sm.st[0].vis[i]; if (v == top1) nmax++; else top2 = max(top2, v); }
        blk_sum_max<N, NT>(sm, k, nmax, top2);
        if (nmax >= 2) top2 = top1;
        const int remaining = visits - sm.st[0].hdr[H_NV];
        const int cutoff = strict ? 0 : top1 - top2;
        decided = remaining < cutoff;
    }
    bool fail = false;
    for (int bi = 0; bi < batch && bi < WALK_MAX_BATCH; bi++) {
        if (desc >= visits || (desc > 0 && decided)) { k.sync(); if (k.tid == 0) gs[GS_DONE] = 1; break; }
        long long pt0 = prof ? clock64() : 0;
        int color = root_color, cur = 0, plen = 0, slot = 0, m1 = root_last;
        unsigned* path = D.path + (q + nleaf) * D.max_depth;
        for (;;) {
            Stage& st = slot_ptr(slot);
            const int next = select_puct_blk<N, NT>(sm, st, D.cgos != 0, k, prof ? D.prof : nullptr);         // tree.py:213
            if (prof) { const long long c = clock64(); D.prof[1] += c - pt0; pt0 = c; D.prof[5]++; }
            const size_t row = (size_t)cur * G::AP;
            const int mv = st.action[next];
            const int vl = st.vl[next], hvl = st.hdr[H_VL];
            const int cv_before = st.vis[next] + vl;
            int ci = st.cidx[next];
            k.sync();                                        // every thread has read the edge before it changes
            const bool coherent = slot == 0 || nc > 2;       // (with two slots, slot 1 is scratch: refetched on every use)
            if (k.tid == 0) {
                path[plen] = ((unsigned)cur << PATH_NODE_SHIFT) | (unsigned)next;
                t.hdr[(size_t)cur * H_STRIDE + H_VL] = hvl + 1; t.cvl[row + next] = vl + 1;                  // :221 add_virtual_loss
                if (coherent) { st.hdr[H_VL] = hvl + 1; st.vl[next] = vl + 1; }
                ws.mv[plen] = (int16_t)mv;
            }
            plen++;
            color = opp(color);
            const int moves = rs.moves + plen;               // GoBoard.moves after this ply
            int expand_threshold = 1;
            if (moves > 2) {                                 // :224-229
                if (moves - 1 >= G::MAXREC) { k.sync(); if (k.tid == 0) gs[GS_ERROR] |= ERR_HISTORY; fail = true; break; }
                if (mv == PASS && m1 == PASS) expand_threshold = 10000000;
            }
            m1 = mv;
            if (cv_before + 1 < expand_threshold + 1) {      // :231-241: this edge ends the descent
                int flags = 0;
                if (ci == NOT_EXPANDED) {                    // the node gets its index and its link now, its rows later
                    if (nnodes >= D.tree.max_nodes) { k.sync(); if (k.tid == 0) gs[GS_ERROR] |= ERR_NODES; fail = true; break; }
                    ci = nnodes++;
                    if (k.tid == 0) {
                        t.cidx[row + next] = ci;
                        if (coherent) st.cidx[next] = ci;
                        ws.newleaf[ci - nn0] = (int16_t)nleaf;
                    }
                    flags |= LEAF_EXPAND;
                }
                if (nleaf >= D.cap) { k.sync(); if (k.tid == 0) gs[GS_ERROR] |= ERR_QUEUE; fail = true; break; }
                int dup_of = -1;
                if (D.dedup) {                               // identical path among the earlier entries of this batch
                    __threadfence_block();
                    k.sync();
                    int first = 0x7fffffff;
                    for (int j = k.tid; j < nleaf; j += NT) {
                        if (D.path_len[q + j] != plen) continue;
                        const unsigned* pj = D.path + (q + j) * D.max_depth;
                        bool same = true;
                        for (int d = plen - 1; d >= 0 && same; d--) same = (pj[d] == path[d]);
                        if (same) { first = j; break; }
                    }
                    first = blk_min_i<N, NT>(sm, k, first);
                    if (first != 0x7fffffff) dup_of = first;
                }
                int lsl;
                if (dup_of >= 0) lsl = ws.lslot[dup_of]; else { lsl = nuniq++; flags |= LEAF_SNAP; }
                if (k.tid == 0) {
                    D.path_len[q + nleaf] = plen; D.leaf_node[q + nleaf] = ci; D.leaf_slot[q + nleaf] = lsl;
                    D.leaf_flag[q + nleaf] = (uint8_t)flags;
                    ws.lslot[nleaf] = (int16_t)lsl; ws.lflag[nleaf] = (uint8_t)flags;
                }
                nleaf++;
                if (nc == 2 && slot == 1 && k.tid == 0) ws.tag[1] = -1;
                if (prof) D.prof[7]++;
                break;
            }
            if (plen >= D.max_depth) { k.sync(); if (k.tid == 0) gs[GS_ERROR] |= ERR_DEPTH; fail = true; break; }
            __threadfence_block();
            k.sync();                                        // thread 0's updates (rows, ws.mv, tags) are visible
            if (ci >= nn0 && !((ws.mat[(ci - nn0) >> 5] >> ((ci - nn0) & 31)) & 1u)) {
                // a descent walks into a node allocated earlier in this launch: materialise it now (what the inline kernel
                // does at every leaf): replay the path on the scratch board, fill the rows, write the snapshot
                if (prof) pt0 = clock64();
                bb_copy<N, NT>(sm.scratch, sm.root, k);
                BScal s = rs;
                int c = root_color;
                for (int d = 0; d < plen; d++) { bb_put_stone<N, NT>(sm, sm.scratch, s, ws.mv[d], c, sm.zob, hh, hp, k); c = opp(c); }
                expand_node_blk<N, NT>(sm, D, t, g, gs, sm.scratch, s, c, HistView{hh, hh, 0}, move_key, k, ci);
                const int li = ws.newleaf[ci - nn0];
                if (ws.lflag[li] & LEAF_SNAP) {
                    const int prev = (s.moves - 1 < G::MAXREC) ? hp[s.moves - 1] : 0;
                    write_snap_blk<N, NT>(D.snap + (q + ws.lslot[li]) * Snap<N>::BYTES, sm.scratch, prev, s.moves, c, k);
                }
                k.sync();
                if (k.tid == 0) {
                    ws.lflag[li] = 0; D.leaf_flag[q + li] = 0;
                    ws.mat[(ci - nn0) >> 5] |= 1u << ((ci - nn0) & 31);
                }
                __threadfence_block();
                k.sync();
                if (prof) { const long long cc = clock64(); D.prof[3] += cc - pt0; pt0 = cc; D.prof[6]++; }
            }
            // the child's rows: cache hit, or one burst of asynchronous copies into the next victim slot
            int hit = -1;
            if (nc > 2) for (int sl = 1; sl < nc; sl++) if (ws.tag[sl] == ci) hit = sl;
            if (hit < 0) {
                if (prof) pt0 = clock64();
                int victim = 1;
                if (nc > 2) { victim = rr; if (victim == slot) victim = victim + 1 < nc ? victim + 1 : 1; rr = victim + 1 < nc ? victim + 1 : 1; }
                stage_node<N, NT>(slot_ptr(victim), t, ci, k);
                stage_wait();
                k.sync();
                if (k.tid == 0) ws.tag[victim] = ci;         // (after the barrier: slower threads may still be in the lookup above;
                hit = victim;                                //  the next lookup is behind the barriers of the next selection)
                if (prof) { const long long cc = clock64(); D.prof[10] += cc - pt0; pt0 = cc; }
            }
            cur = ci; slot = hit;
        }
        __threadfence_block();
        k.sync();
        if (fail) break;
        desc++;
    }
    if (k.tid == 0) { gs[GS_DESC] = desc; gs[GS_NLEAF] = nleaf; gs[GS_NUNIQ] = nuniq; gs[GS_NNODES] = nnodes; }
}

// ---------------------------------------------------------------------------------------------------------------------
// Wavefront tree walk: the descents of one batch as a software pipeline.
//
// Descent d+1 depends on descent d only through the nodes both visit, and it visits them one ply later: d changes a node
// (virtual loss, child link) at the step it selects there and then moves DOWN.  So if every descent in flight advances one
// ply per step and a new descent enters at the root each step, all descents in flight are at different depths -- different
// nodes -- and every node still sees its visitors in descent order: the result is the sequential one, with up to NG plies
// (and, more to the point, NG row fetches from L2) in flight at once.  A CTA of NT threads runs NG = NT / GT thread groups,
// one descent each; three block barriers per step separate scoring, the group leaders' updates, and the row fetches.
//
// Two things are numbered in descent order but finish out of order (a shallow leaf of a later descent ends before a deep
// earlier one): queue slots (= the descent's ordinal, known up front) and node indices of new leaves.  New nodes are
// therefore linked with a provisional code (-2 - ordinal) and numbered after the walk, by a prefix count over the ordinals.
// A descent that selects an edge carrying a provisional code has to enter a node that does not exist yet: it waits until
// it is the oldest descent in flight (every older ordinal has ended, so the node's index is final), the younger descents
// freeze behind it, and the whole CTA materialises the node (replay + expansion, as k_walk_puct_blk does).
// ---------------------------------------------------------------------------------------------------------------------
template <int N, int NT, int GT> struct WaveSmem {
    static constexpr int NG = NT / GT, GW = GT / 32;
    BlkSmem<N, NT> b;
    struct GState { int active, ord, cur, plen, color, m1, wait_ci, pad; } gst[NG];
    u64 px[NG][GW]; int pi[NG][GW];                        // per-warp argmax partials of every group
    int dec[NG];                                           // node whose rows the group fetches this step (-1: none)
    int err;
    int16_t mv[Geo<N>::MAXREC + 2];                        // moves of a path being materialised
    int final_idx[WALK_MAX_BATCH];                         // node index of the leaf of every ordinal once it is final (-1 before)
    uint8_t alloc[WALK_MAX_BATCH];                         // the descent of this ordinal allocated a node
    // followed by (NG - 1) more NodeStage buffers in dynamic shared memory (group g > 0 uses buffer g - 1; group 0 b.st[1])
};

template <int N, int NT, int GT>
__global__ void __launch_bounds__(NT) k_wave_puct_blk(Dev D, const uint32_t* __restrict__ eye2, int visits, int batch, int strict, int nsqrt)
{
    using G = Geo<N>;
    using WS = WaveSmem<N, NT, GT>;
    using Stage = typename BlkSmem<N, NT>::NodeStage;
    constexpr int NG = WS::NG, GW = WS::GW, INF = 0x7fffffff;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WS& ws = *reinterpret_cast<WS*>(smem_raw);
    BlkSmem<N, NT>& sm = ws.b;
    Stage* extra = reinterpret_cast<Stage*>(smem_raw + ((sizeof(WS) + 15) & ~(size_t)15));
    double* sqtab = reinterpret_cast<double*>(extra + (NG - 1));      // sqrt(n), n < nsqrt: one table per launch instead of a
    Blk<NT> k;                                                        // ~450-cycle float64 square root per thread and ply
    for (int i = threadIdx.x; i < nsqrt; i += NT) sqtab[i] = sqrt((double)i);
    const int gi = k.tid / GT, gt = k.tid % GT, gw = gt >> 5;
    Stage& gbuf = gi == 0 ? sm.st[1] : extra[gi - 1];
    const int g = blockIdx.x;
    int* gs = D.gs + (size_t)g * GS_STRIDE;
    const bool idle = !gs[GS_ACTIVE] || gs[GS_FINISHED] || gs[GS_ERROR] || gs[GS_DONE];
    k.sync();
    if (k.tid == 0) { gs[GS_NLEAF] = 0; gs[GS_NUNIQ] = 0; }
    __threadfence_block();
    k.sync();
    if (idle) return;
    const bool prof = D.prof && g == 0 && k.tid == 0;
    const long long pk0 = prof ? clock64() : 0;
    const Tree t = tree_of<G::AP>(D.tree, g);
    stage_node<N, NT>(sm.st[0], t, 0, k);
    for (int i = k.tid; i < WALK_MAX_BATCH; i += NT) { ws.final_idx[i] = -1; ws.alloc[i] = 0; }
    if (k.tid < NG) { ws.gst[k.tid].active = 0; ws.gst[k.tid].wait_ci = -1; ws.gst[k.tid].ord = INF; ws.dec[k.tid] = -1; }
    if (k.tid == 0) ws.err = 0;
    BScal rs;                                                // the walk needs the root's scalars; its board, the eye table and the
    {                                                        // Zobrist keys are loaded by the first on-demand materialisation, if any
        const int* sc = D.b_scal + (size_t)g * 8;
        rs.hash = D.b_hash[g];
        rs.moves = sc[0]; rs.ko_pos = sc[1]; rs.ko_move = sc[2]; rs.pris0 = sc[3]; rs.pris1 = sc[4];
    }
    bool board_loaded = false;
    const int root_color = gs[GS_COLOR];
    u64* hh = D.hist_hash + (size_t)g * G::MAXREC;
    int16_t* hp = D.hist_pos + (size_t)g * G::MAXREC;
    const unsigned move_key = (unsigned)rs.moves;
    const int root_last = (rs.moves >= 1 && rs.moves - 1 < G::MAXREC) ? hp[rs.moves - 1] : -1;
    const size_t q = (size_t)g * D.cap;
    const int nn0 = gs[GS_NNODES], desc0 = gs[GS_DESC];
    stage_wait();
    k.sync();
    int nd;                                                  // descents of this launch (tree.py:130-174 + is_move_decided)
    {
        const int nk = sm.st[0].hdr[H_K];
        int top1 = 0;
        for (int i = k.tid; i < nk; i += NT) top1 = max(top1, sm.st[0].vis[i]);
        top1 = blk_max_i<N, NT>(sm, k, top1);
        int nmax = 0, top2 = 0;
        for (int i = k.tid; i < nk; i += NT) { const int v = sm.st[0].vis[i]; if (v == top1) nmax++; else top2 = max(top2, v); }
        blk_sum_max<N, NT>(sm, k, nmax, top2);
        if (nmax >= 2) top2 = top1;
        const int remaining = visits - sm.st[0].hdr[H_NV];
        const int cutoff = strict ? 0 : top1 - top2;
        const bool decided = remaining < cutoff;
        const int room = max(0, visits - desc0);
        nd = min(min(batch, WALK_MAX_BATCH), room);
        if (decided) nd = min(nd, desc0 == 0 ? 1 : 0);
        nd = min(nd, D.cap);
    }
    const bool stops_early = nd < batch;                     // the sequential loop would hit its stop test inside this launch
    // Group states live in shared memory, written only by the group's own leader in S2 (and by thread 0 inside the
    // materialisation, behind barriers); what every thread has to agree on from one step to the next -- which group a new
    // descent starts in, the stalled ordinal, the number of descents in flight -- is carried in registers.
    int next_ord = nd > 0 ? 1 : 0, start_g = nd > 0 ? 0 : -1, stall_ord = INF, nact = nd > 0 ? 1 : 0;
    k.sync();
    long long pr[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // clock64 accumulators of thread 0 (registers; flushed at the end)
    if (prof) pr[0] = clock64() - pk0;
    const bool cgos = D.cgos != 0;
    for (;;) {
        // ---- S1: every running group scores the children of its node
        if (nact == 0 && next_ord >= nd) break;
        long long pt0 = prof ? clock64() : 0;
        typename WS::GState me = ws.gst[gi];
        const bool fresh = gi == start_g;
        if (fresh) { me.active = 1; me.ord = next_ord - 1; me.cur = 0; me.plen = 0; me.color = root_color; me.m1 = root_last; me.wait_ci = -1; }
        const bool run = me.active && me.wait_ci == -1 && me.ord < stall_ord;
        const Stage& st = me.cur == 0 ? sm.st[0] : gbuf;
        if (run) {
            const int nk = st.hdr[H_K];
            const int nsq = st.hdr[H_NV] + st.hdr[H_VL] + 1;                  // pucb.py:16
            const double sq = nsq < nsqrt ? sqtab[nsq] : sqrt((double)nsq);
            // node.py:141-157 + pucb.py:8-29, the arithmetic of select_puct_blk per child.  The walk is bound by the SM's float64
            // issue rate (measured: ~1 warp instruction per cycle per SM), so float64 work is what is trimmed: the square root
            // comes from the group leader (ws.sq), 1.0 * prior is the prior, and a warp none of whose 32 children has been
            // visited skips both correctly rounded divisions (~24 float64 instructions) -- below the root that is most warps.
            constexpr int CHG = (G::A + GT - 1) / GT;
            double vv[CHG];
            int cvs[CHG];
            bool redo = false;                               // redo: an operand outside the fast division's range (a prior below 2^-120)
#pragma unroll
            for (int c = 0; c < CHG; c++) {
                const int i = gt + c * GT;
                const bool valid = i < nk;
                cvs[c] = valid ? st.vis[i] + st.vl[i] : 0;
                vv[c] = dmul(valid ? st.pol[i] : 0.0, sq);   // (1.0 * prior) * sqrt(...): the first product is exact
            }
            if (prof) { const long long c = clock64(); pr[2] += c - pt0; pr[6]++; }
            // divisions only where a warp has a visited child (below the root most warps have none): the walk is bound by
            // instruction issue (16 warps on 4 schedulers), not by the latency of the division chains
#pragma unroll
            for (int c = 0; c < CHG; c++) {
                const bool has = cvs[c] != 0;
                if (!__any_sync(0xffffffffu, has)) continue;
                const int i = gt + c * GT;
                const float vs = has ? st.vsum[i] : 0.f;
                const double num = vv[c];
                const bool hq = has && vs != 0.0f;           // 0 / x = +0 exactly: no division
                const bool hu = has && num != 0.0;
                bool sq_, su_;
                const double qd = ddiv_fast(hq ? (double)vs : 1.0, has ? (double)cvs[c] : 1.0, sq_);
                const double ud = ddiv_fast(hu ? num : 1.0, (double)(cvs[c] + 1), su_);
                if (has) vv[c] = dadd(hq ? qd : 0.0, hu ? ud : 0.0);
                redo |= (hq && sq_) || (hu && su_);
            }
            if (cgos && nk >= 1 && (nk - 1) % GT == gt) {
#pragma unroll
                for (int c = 0; c < CHG; c++) if (gt + c * GT == nk - 1) vv[c] = dsub(vv[c], 0.1);
            }
            if (prof) pr[3] += clock64() - pt0;
            if (redo) {                                      // rare: the library division for this thread's children
                for (int c = 0; c < CHG; c++) {
                    const int i = gt + c * GT;
                    if (i >= nk) continue;
                    const int cv = st.vis[i] + st.vl[i];
                    const double num = dmul(st.pol[i], sq);
                    double v = num;
                    if (cv != 0) {
                        const float vs = st.vsum[i];
                        const double qv = vs == 0.0f ? 0.0 : ddiv((double)vs, (double)cv);
                        const double u = num == 0.0 ? 0.0 : ddiv(num, (double)(cv + 1));
                        v = dadd(qv, u);
                    }
                    if (cgos && i == nk - 1) v = dsub(v, 0.1);
                    vv[c] = v;
                }
            }
            double bv = 0.0; int bi = INF;
#pragma unroll
            for (int c = 0; c < CHG; c++) {
                const int i = gt + c * GT;
                if (i < nk && (bi == INF || vv[c] > bv)) { bv = vv[c]; bi = i; }
            }
            u64 key = bi != INF ? order_key(bv) : 0ull;
            warp_argmax_key(key, bi);
            if (k.lane == 0) { ws.px[gi][gw] = key; ws.pi[gi][gw] = bi; }
            if (prof) pr[4] += clock64() - pt0;
        }
        k.sync();                                            // B1
        if (prof) { const long long c = clock64(); pr[8] += c - pt0; pt0 = c; pr[5]++; }
        // ---- S2: the group leader applies the ply
        if (run) {
            if (gt == 0) {
                // only the leader needs the group's choice: it folds the GW per-warp partials itself (a handful of compares
                // instead of three dependent warp reductions in every warp of the group)
                u64 k2 = ws.px[gi][0];
                int next = ws.pi[gi][0];
#pragma unroll
                for (int w = 1; w < GW; w++) {
                    const u64 kw = ws.px[gi][w];
                    const int iw = ws.pi[gi][w];
                    if (kw > k2 || (kw == k2 && kw != 0ull && iw < next)) { k2 = kw; next = iw; }
                }
                if (prof) pr[7] += clock64() - pt0;
                typename WS::GState& ms = ws.gst[gi];
                if (fresh) ms = me;
                const int cur = me.cur;
                const size_t row = (size_t)cur * G::AP;
                const int mv = st.action[next];
                const int vl = st.vl[next], hvl = st.hdr[H_VL];
                const int cv_before = st.vis[next] + vl;
                int ci = st.cidx[next];
                if (ci < -1 && ws.final_idx[-ci - 2] >= 0) ci = ws.final_idx[-ci - 2];   // (a copy fetched before the node was materialised)
                unsigned* path = D.path + (q + me.ord) * D.max_depth;
                path[me.plen] = ((unsigned)cur << PATH_NODE_SHIFT) | (unsigned)next;
                t.hdr[(size_t)cur * H_STRIDE + H_VL] = hvl + 1; t.cvl[row + next] = vl + 1;                  // :221 add_virtual_loss
                if (cur == 0) { sm.st[0].hdr[H_VL] = hvl + 1; sm.st[0].vl[next] = vl + 1; }
                const int plen = me.plen + 1;
                const int moves = rs.moves + plen;
                int expand_threshold = 1, dec = -1;
                bool bad = false;
                if (moves > 2) {                             // :224-229
                    if (moves - 1 >= G::MAXREC) { atomicOr(&ws.err, ERR_HISTORY); bad = true; }
                    if (mv == PASS && me.m1 == PASS) expand_threshold = 10000000;
                }
                if (bad) ms.active = 0;
                else if (cv_before + 1 < expand_threshold + 1) {                  // :231-241: the edge ends the descent
                    int flags = LEAF_SNAP;
                    if (ci == NOT_EXPANDED) {
                        ci = -2 - me.ord;                    // provisional: numbered after the walk
                        t.cidx[row + next] = ci;
                        if (cur == 0) sm.st[0].cidx[next] = ci;
                        ws.alloc[me.ord] = 1;
                        flags |= LEAF_EXPAND;
                    }
                    D.path_len[q + me.ord] = plen; D.leaf_node[q + me.ord] = ci; D.leaf_slot[q + me.ord] = me.ord;
                    D.leaf_flag[q + me.ord] = (uint8_t)flags;
                    ms.active = 0; ms.ord = INF;
                } else if (plen >= D.max_depth) { atomicOr(&ws.err, ERR_DEPTH); ms.active = 0; }
                else {
                    ms.plen = plen; ms.color = opp(me.color); ms.m1 = mv;
                    if (ci < -1) ms.wait_ci = ci;            // the child is a leaf of an earlier descent of this launch
                    else { ms.cur = ci; dec = ci; }
                }
                if (prof) pr[11] += clock64() - pt0;
                ws.dec[gi] = dec;                            // (no fence: the block barrier that follows orders these writes, global
            }                                                //  ones included, before every later access of the CTA)
        }
        k.sync();                                            // B2
        if (prof) { const long long c = clock64(); pr[9] += c - pt0; pt0 = c; }
        // ---- S3: row fetches, on-demand materialisation, the next descent enters
        if (ws.err) break;
        if (run && ws.dec[gi] >= 0) {
            const int node = ws.dec[gi];
            const size_t row = (size_t)node * G::AP;
            for (int c = gt; c < G::AP / 4; c += GT) {
                cp_async16(gbuf.vis + 4 * c, t.cvis + row + 4 * c);
                cp_async16(gbuf.vl + 4 * c, t.cvl + row + 4 * c);
                cp_async16(gbuf.vsum + 4 * c, t.cvsum + row + 4 * c);
                cp_async16(gbuf.cidx + 4 * c, t.cidx + row + 4 * c);
            }
            for (int c = gt; c < G::AP / 2; c += GT) cp_async16(gbuf.pol + 2 * c, t.cpol + row + 2 * c);
            for (int c = gt; c < G::AP / 8; c += GT) cp_async16(gbuf.action + 8 * c, t.action + row + 8 * c);
            if (gt < H_STRIDE / 4) cp_async16(gbuf.hdr + 4 * gt, t.hdr + (size_t)node * H_STRIDE + 4 * gt);
        }
        stall_ord = INF; nact = 0; start_g = -1;
        int first_free = -1, mg = -1, min_ord = INF;
#pragma unroll
        for (int j = NG - 1; j >= 0; j--) {
            if (!ws.gst[j].active) { first_free = j; continue; }
            nact++; min_ord = min(min_ord, ws.gst[j].ord);
            if (ws.gst[j].wait_ci != -1 && ws.gst[j].ord < stall_ord) { stall_ord = ws.gst[j].ord; mg = j; }
        }
        if (stall_ord != INF && min_ord == stall_ord) {
            // the waiting descent is the oldest in flight: the node it wants to enter gets its final index and its rows now
            const int L = -ws.gst[mg].wait_ci - 2, ord = ws.gst[mg].ord, plen = ws.gst[mg].plen, c_at = ws.gst[mg].color;
            int cnt = 0, dummy = 0;
            for (int j = k.tid; j < L; j += NT) cnt += ws.alloc[j];
            blk_sum_max<N, NT>(sm, k, cnt, dummy);
            const int idx = nn0 + cnt;
            const unsigned* path = D.path + (q + ord) * D.max_depth;
            for (int d = k.tid; d < plen; d += NT) {
                const unsigned e = path[d];
                ws.mv[d] = t.action[(size_t)(e >> PATH_NODE_SHIFT) * G::AP + (e & ((1u << PATH_NODE_SHIFT) - 1))];
            }
            k.sync();
            if (idx >= D.tree.max_nodes) { if (k.tid == 0) ws.err = ERR_NODES; k.sync(); break; }
            if (!board_loaded) {
                for (int i = k.tid; i < 4096 / 4; i += NT) reinterpret_cast<uint4*>(sm.eye2)[i] = reinterpret_cast<const uint4*>(eye2)[i];
                for (int i = k.tid; i < 4 * G::CELLS; i += NT) sm.zob[i] = D.zob[i];
                BScal tmp;
                bb_load<N, NT>(sm.root, tmp, pool_of<N>(D), g, k);
                board_loaded = true;
            }
            bb_copy<N, NT>(sm.scratch, sm.root, k);
            BScal s = rs;
            int c = root_color;
            for (int d = 0; d < plen; d++) { bb_put_stone<N, NT>(sm, sm.scratch, s, ws.mv[d], c, sm.zob, hh, hp, k); c = opp(c); }
            expand_node_blk<N, NT>(sm, D, t, g, gs, sm.scratch, s, c_at, HistView{hh, hh, 0}, move_key, k, idx);
            {
                const int prev = (s.moves - 1 < G::MAXREC) ? hp[s.moves - 1] : 0;
                write_snap_blk<N, NT>(D.snap + (q + L) * Snap<N>::BYTES, sm.scratch, prev, s.moves, c_at, k);
            }
            if (k.tid == 0) {
                const unsigned e = path[plen - 1];
                const int pnode = (int)(e >> PATH_NODE_SHIFT), pchild = (int)(e & ((1u << PATH_NODE_SHIFT) - 1));
                t.cidx[(size_t)pnode * G::AP + pchild] = idx;
                if (pnode == 0) sm.st[0].cidx[pchild] = idx;
                D.leaf_node[q + L] = idx; D.leaf_flag[q + L] = 0;
                ws.final_idx[L] = idx;
                ws.gst[mg].cur = idx; ws.gst[mg].wait_ci = -1;
                __threadfence_block();
            }
            k.sync();
            if (gi == mg) {
                const size_t row = (size_t)idx * G::AP;
                for (int cc = gt; cc < G::AP / 4; cc += GT) {
                    cp_async16(gbuf.vis + 4 * cc, t.cvis + row + 4 * cc);
                    cp_async16(gbuf.vl + 4 * cc, t.cvl + row + 4 * cc);
                    cp_async16(gbuf.vsum + 4 * cc, t.cvsum + row + 4 * cc);
                    cp_async16(gbuf.cidx + 4 * cc, t.cidx + row + 4 * cc);
                }
                for (int cc = gt; cc < G::AP / 2; cc += GT) cp_async16(gbuf.pol + 2 * cc, t.cpol + row + 2 * cc);
                for (int cc = gt; cc < G::AP / 8; cc += GT) cp_async16(gbuf.action + 8 * cc, t.action + row + 8 * cc);
                if (gt < H_STRIDE / 4) cp_async16(gbuf.hdr + 4 * gt, t.hdr + (size_t)idx * H_STRIDE + 4 * gt);
            }
            stall_ord = INF;
            for (int j = 0; j < NG; j++) if (j != mg && ws.gst[j].active && ws.gst[j].wait_ci != -1) stall_ord = min(stall_ord, ws.gst[j].ord);
        }
        if (stall_ord == INF && next_ord < nd && first_free >= 0) { start_g = first_free; next_ord++; nact++; }   // (its leader writes the state in S2)
        stage_wait();
        k.sync();                                            // B3
        if (prof) { const long long c = clock64(); pr[10] += c - pt0; pr[1] += nact; }
    }
    stage_wait();
    k.sync();
    if (ws.err) {                                            // the game is dead: nothing of this launch is evaluated or backed up
        if (k.tid == 0) { gs[GS_ERROR] |= ws.err; gs[GS_NLEAF] = 0; gs[GS_NUNIQ] = 0; }
        return;
    }
    // ---- node numbering in descent order, links of the new leaves, queue entries that refer to them
    const int nleaf = next_ord;
    int total = 0, dummy = 0;
    for (int j = k.tid; j < nleaf; j += NT) total += ws.alloc[j];
    blk_sum_max<N, NT>(sm, k, total, dummy);
    if (nn0 + total > D.tree.max_nodes) {
        if (k.tid == 0) { gs[GS_ERROR] |= ERR_NODES; gs[GS_NLEAF] = 0; gs[GS_NUNIQ] = 0; }
        return;
    }
    for (int j = k.tid; j < nleaf; j += NT) {
        if (!ws.alloc[j] || ws.final_idx[j] >= 0) continue;
        int cnt = 0;
        for (int i = 0; i < j; i++) cnt += ws.alloc[i];
        const int idx = nn0 + cnt;
        ws.final_idx[j] = idx;
        const int plen = D.path_len[q + j];
        const unsigned e = D.path[(q + j) * D.max_depth + plen - 1];
        t.cidx[(size_t)(e >> PATH_NODE_SHIFT) * G::AP + (e & ((1u << PATH_NODE_SHIFT) - 1))] = idx;
    }
    __threadfence_block();
    k.sync();
    for (int j = k.tid; j < nleaf; j += NT) {
        const int ln = D.leaf_node[q + j];
        if (ln < -1) D.leaf_node[q + j] = ws.final_idx[-ln - 2];
    }
    if (k.tid == 0) {
        gs[GS_DESC] = desc0 + nleaf; gs[GS_NLEAF] = nleaf; gs[GS_NUNIQ] = nleaf; gs[GS_NNODES] = nn0 + total;
        if (stops_early) gs[GS_DONE] = 1;
    }
    if (prof) for (int i = 0; i < 12; i++) D.prof[i] += pr[i];
}

// The board half of a batch: one CTA per queued leaf (blockIdx.x) of every game (blockIdx.y).
template <int N, int NT> struct ExpandSmem {
    BlkSmem<N, NT> b;
    u64 ph[Geo<N>::MAXREC];                                // position hashes of the descent's own moves
    int16_t pp[Geo<N>::MAXREC];                            // ... and their points
    int16_t mv[Geo<N>::MAXREC + 2];
};

template <int N, int NT>
__global__ void __launch_bounds__(NT) k_expand_leaves_blk(Dev D, const uint32_t* __restrict__ eye2)
{
    using G = Geo<N>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ExpandSmem<N, NT>& es = *reinterpret_cast<ExpandSmem<N, NT>*>(smem_raw);
    BlkSmem<N, NT>& sm = es.b;
    Blk<NT> k;
    const int g = blockIdx.y, i = blockIdx.x;
    int* gs = D.gs + (size_t)g * GS_STRIDE;
    if (i >= gs[GS_NLEAF]) return;
    const size_t q = (size_t)g * D.cap;
    const int flags = D.leaf_flag[q + i];
    if (!flags) return;
    const Tree t = tree_of<G::AP>(D.tree, g);
    const int plen = D.path_len[q + i];
    const unsigned* path = D.path + (q + i) * D.max_depth;
    for (int d = k.tid; d < plen; d += NT) {
        const unsigned e = path[d];
        es.mv[d] = t.action[(size_t)(e >> PATH_NODE_SHIFT) * G::AP + (e & ((1u << PATH_NODE_SHIFT) - 1))];
    }
    if (flags & LEAF_EXPAND)
        for (int j = k.tid; j < 4096 / 4; j += NT) reinterpret_cast<uint4*>(sm.eye2)[j] = reinterpret_cast<const uint4*>(eye2)[j];
    for (int j = k.tid; j < 4 * G::CELLS; j += NT) sm.zob[j] = D.zob[j];
    BScal s;
    bb_load<N, NT>(sm.root, s, pool_of<N>(D), g, k);
    const int root_moves = s.moves;
    int color = gs[GS_COLOR];
    // the descent's record entries land in the private arrays: index `moves` of the game's record is ph[moves - root_moves]
    u64* hh = es.ph - root_moves;
    int16_t* hp = es.pp - root_moves;
    for (int d = 0; d < plen; d++) { bb_put_stone<N, NT>(sm, sm.root, s, es.mv[d], color, sm.zob, hh, hp, k); color = opp(color); }
    if (flags & LEAF_EXPAND)
        expand_node_blk<N, NT>(sm, D, t, g, gs, sm.root, s, color, HistView{D.hist_hash + (size_t)g * G::MAXREC, es.ph, root_moves},
                               (unsigned)root_moves, k, D.leaf_node[q + i]);
    if (flags & LEAF_SNAP) {
        const int prev = (s.moves - 1 < G::MAXREC) ? (s.moves - 1 >= root_moves ? hp[s.moves - 1] : D.hist_pos[(size_t)g * G::MAXREC + s.moves - 1]) : 0;
        write_snap_blk<N, NT>(D.snap + (q + D.leaf_slot[q + i]) * Snap<N>::BYTES, sm.root, prev, s.moves, color, k);
    }
}

// ---- process_mini_batch after the forward pass (tree.py:287-315) ----------------------------------------------------
// (a) priors / raw value of every evaluated node: all (leaf, child) pairs are independent and are spread over the
//     whole block with several loads in flight per thread; (b) values: one warp walks the leaves in queue order (fp32
//     sums depend on it) with the next leaf's path entries already in flight.
// Part (a) alone, one CTA per evaluated leaf (blockIdx.x) of every game (blockIdx.y): with a handful of games the priors of
// a 256-leaf batch are the larger half of the backup, and every (leaf, child) pair is independent.
template <int N>
__global__ void __launch_bounds__(128) k_backup_priors_blk(Dev D, int use_logit)
{
    using G = Geo<N>;
    const int g = blockIdx.y, i = blockIdx.x;
    const int* gs = D.gs + (size_t)g * GS_STRIDE;
    if (i >= gs[GS_NLEAF]) return;
    const size_t q = (size_t)g * D.cap;
    const int slot = gs[GS_SLOT0] + D.leaf_slot[q + i];
    const int ni = D.leaf_node[q + i];
    if (slot >= D.slot_cap || ni < 0) return;                                    // (k_backup_blk reports the overflow)
    const Tree t = tree_of<G::AP>(D.tree, g);
    const float* v = D.value + (size_t)slot * 3;
    const float* pol = D.policy + (size_t)slot * G::A;
    if (threadIdx.x == 0) t.hdr[(size_t)ni * H_STRIDE + H_RAW] = __float_as_int(__fadd_rn(__fmul_rn(v[1], 0.5f), v[2]));  // tree.py:300
    const int nk = t.hdr[(size_t)ni * H_STRIDE + H_K];
    for (int c = threadIdx.x; c < nk; c += 128) {                                // node.py:86-93, tree.py:287-299
        const int a = t.action[(size_t)ni * G::AP + c];
        float p = a == PASS ? pol[G::NN] : pol[(a / G::W - 1) * N + (a % G::W - 1)];
        if (a == PASS && use_logit) p = __fsub_rn(p, 0.5f);                                                           // tree.py:292-294
        t.cpol[(size_t)ni * G::AP + c] = (double)p;
    }
}

template <int N, int NT>
__global__ void __launch_bounds__(NT) k_backup_blk(Dev D, int use_logit, int priors_done)
{
    using G = Geo<N>;
    const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int* gs = D.gs + (size_t)g * GS_STRIDE;
    const int nl = gs[GS_NLEAF];
    if (nl == 0) return;
    const Tree t = tree_of<G::AP>(D.tree, g);
    const int slot0 = gs[GS_SLOT0];
    const size_t q = (size_t)g * D.cap;
    __shared__ int bad;
    if (tid == 0) bad = 0;
    __syncthreads();
    for (int i = tid; i < nl; i += NT) {
        const int slot = slot0 + D.leaf_slot[q + i];
        if (slot >= D.slot_cap) { bad = 1; continue; }
        const int ni = D.leaf_node[q + i];
        if (ni >= 0 && !priors_done) {
            const float* v = D.value + (size_t)slot * 3;
            t.hdr[(size_t)ni * H_STRIDE + H_RAW] = __float_as_int(__fadd_rn(__fmul_rn(v[1], 0.5f), v[2]));           // tree.py:300
        }
    }
    __syncthreads();
    if (bad) { if (tid == 0) gs[GS_ERROR] |= ERR_QUEUE; return; }
    constexpr int U = 4;
    for (int p0 = tid; p0 < nl * G::AP && !priors_done; p0 += NT * U) {          // node.py:86-93, tree.py:287-299
        int ni[U], c[U], a[U]; const float* pol[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int p = p0 + u * NT;
            const int i = p / G::AP;
            c[u] = p - i * G::AP; ni[u] = -1; a[u] = 0; pol[u] = nullptr;
            if (p < nl * G::AP) {
                ni[u] = D.leaf_node[q + i];
                pol[u] = D.policy + (size_t)(slot0 + D.leaf_slot[q + i]) * G::A;
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++)
            if (ni[u] >= 0) { if (c[u] < t.hdr[(size_t)ni[u] * H_STRIDE + H_K]) a[u] = t.action[(size_t)ni[u] * G::AP + c[u]]; else ni[u] = -1; }
        float pv[U];
#pragma unroll
        for (int u = 0; u < U; u++)
            if (ni[u] >= 0) pv[u] = a[u] == PASS ? pol[u][G::NN] : pol[u][(a[u] / G::W - 1) * N + (a[u] % G::W - 1)];
#pragma unroll
        for (int u = 0; u < U; u++)
            if (ni[u] >= 0) {
                float p = pv[u];
                if (a[u] == PASS && use_logit) p = __fsub_rn(p, 0.5f);                                                // tree.py:292-294
                t.cpol[(size_t)ni[u] * G::AP + c[u]] = (double)p;
            }
    }
    // (b) tree.py:301-313 + node.py:118-138: every leaf adds its value to the edges and nodes of its path, leaf by leaf in
    //     queue order -- the sums are fp32, so the ORDER of the additions into one accumulator is part of the result, but
    //     different accumulators are independent.  The path entries of the batch are laid out leaf-major in shared memory;
    //     the thread that owns the first entry of an edge (of a node) adds up all later entries of that edge (node) in order
    //     and writes the accumulator once.  One thread per accumulator instead of one warp walking 256 leaves in sequence.
    extern __shared__ __align__(16) unsigned char bk_smem[];
    constexpr int BK_MAX = 256, BK_ENT = 3072;
    float* s_val = reinterpret_cast<float*>(bk_smem);                            // [BK_MAX]
    int* s_off = reinterpret_cast<int*>(bk_smem + BK_MAX * 4);                   // [BK_MAX + 1]
    unsigned* e_key = reinterpret_cast<unsigned*>(bk_smem + BK_MAX * 8 + 16);    // [BK_ENT]  node << PATH_NODE_SHIFT | child
    unsigned* e_who = e_key + BK_ENT;                                            // [BK_ENT]  leaf << 16 | distance from the leaf
    __shared__ int s_total;
    if (warp == 0) {                                                             // exclusive scan of the path lengths
        int run = 0;
        for (int i0 = 0; i0 < nl && i0 < BK_MAX; i0 += 32) {
            const int i = i0 + lane;
            const int pl = i < nl ? D.path_len[q + i] : 0;
            int inc = pl;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
            if (i < nl) s_off[i] = run + inc - pl;
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) { s_total = run; if (nl <= BK_MAX) s_off[nl] = run; }
    }
    __syncthreads();
    const int E = s_total;
    if (nl <= BK_MAX && E <= BK_ENT) {
        for (int i = tid; i < nl; i += NT) {
            const float* v = D.value + (size_t)(slot0 + D.leaf_slot[q + i]) * 3;
            s_val[i] = __fadd_rn(v[0], __fmul_rn(v[1], 0.5f));                                                        // :303
        }
        for (int i = warp; i < nl; i += NT / 32) {
            const int o = s_off[i], plen = s_off[i + 1] - o;
            const unsigned* path = D.path + (q + i) * D.max_depth;
            for (int d = lane; d < plen; d += 32) { e_key[o + d] = path[plen - 1 - d]; e_who[o + d] = ((unsigned)i << 16) | (unsigned)d; }
        }
        __syncthreads();
        // A node sits at ONE depth of the tree, so in every leaf's segment it can only be at one position (its depth counted
        // from the root end of the segment): both scans visit one entry per leaf instead of every entry of the batch.
        for (int e = tid; e < E; e += NT) {
            const unsigned key = e_key[e], node = key >> PATH_NODE_SHIFT;
            const int i0 = (int)(e_who[e] >> 16);
            const int depth = (s_off[i0 + 1] - s_off[i0]) - 1 - (int)(e_who[e] & 0xffffu);       // plies between the root and this edge
            bool seen_edge = false, seen_node = false;
            for (int i = 0; i < i0 && !(seen_edge && seen_node); i++) {
                const int o = s_off[i], pl = s_off[i + 1] - o;
                if (depth >= pl) continue;
                const unsigned kj = e_key[o + pl - 1 - depth];
                seen_node |= (kj >> PATH_NODE_SHIFT) == node;
                seen_edge |= kj == key;
            }
            if (seen_edge && seen_node) continue;
            const size_t row = (size_t)node * G::AP;
            const int c = (int)(key & ((1u << PATH_NODE_SHIFT) - 1));
            int* h = t.hdr + (size_t)node * H_STRIDE;
            float es = seen_edge ? 0.f : t.cvsum[row + c];
            float hs = seen_node ? 0.f : __int_as_float(h[H_VSUM]);
            int ec = 0, hc = 0; bool has0 = false; float last0 = 0.f;
            for (int i = i0; i < nl; i++) {
                const int o = s_off[i], pl = s_off[i + 1] - o;
                if (depth >= pl) continue;
                const int d = pl - 1 - depth;
                const unsigned kj = e_key[o + d];
                if ((kj >> PATH_NODE_SHIFT) != node) continue;
                const float v0 = s_val[i];
                const float v1 = __fsub_rn(1.0f, v0);
                const float val = d == 0 ? v0 : ((d & 1) ? v1 : __fsub_rn(1.0f, v1));                                 // value = 1 - value per ply (:313)
                if (!seen_node) { hs = __fadd_rn(hs, val); hc++; }
                if (!seen_edge && kj == key) { es = __fadd_rn(es, val); ec++; if (d == 0) { has0 = true; last0 = v0; } }
            }
            if (!seen_edge) {
                t.cvsum[row + c] = es; t.cvis[row + c] += ec; t.cvl[row + c] -= ec;
                if (has0) t.cval[row + c] = last0;                                                                    // :308
            }
            if (!seen_node) { h[H_VSUM] = __float_as_int(hs); h[H_NV] += hc; h[H_VL] -= hc; }
        }
        if (tid == 0) { gs[GS_EVALS] += nl; gs[GS_UEVALS] += gs[GS_NUNIQ]; }
        return;
    }
    // fallback (very long paths): one warp walks the leaves in queue order
    if (warp != 0) return;
    for (int i = 0; i < nl; i++) {
        const float* v = D.value + (size_t)(slot0 + D.leaf_slot[q + i]) * 3;
        const float val0 = __fadd_rn(v[0], __fmul_rn(v[1], 0.5f));
        const int plen = D.path_len[q + i];
        if (plen > 0) {
            const unsigned* path = D.path + (q + i) * D.max_depth;
            const float val1 = __fsub_rn(1.0f, val0), val2 = __fsub_rn(1.0f, val1);
            for (int d0 = 0; d0 < plen; d0 += 32) {
                const int d = d0 + lane;                                         // distance from the leaf
                if (d < plen) {
                    const unsigned e = path[plen - 1 - d];
                    const int node = (int)(e >> PATH_NODE_SHIFT), c = (int)(e & ((1u << PATH_NODE_SHIFT) - 1));
                    const float val = d == 0 ? val0 : ((d & 1) ? val1 : val2);   // value = 1 - value per ply (:313)
                    const size_t row = (size_t)node * G::AP;
                    if (d == 0) t.cval[row + c] = val0;                          // :308
                    t.cvsum[row + c] = __fadd_rn(t.cvsum[row + c], val);
                    t.cvis[row + c] += 1; t.cvl[row + c] -= 1;
                    int* h = t.hdr + (size_t)node * H_STRIDE;
                    h[H_VSUM] = __float_as_int(__fadd_rn(__int_as_float(h[H_VSUM]), val));
                    h[H_NV] += 1; h[H_VL] -= 1;
                }
            }
        }
        __syncwarp();
    }
    if (lane == 0) { gs[GS_EVALS] += nl; gs[GS_UEVALS] += gs[GS_NUNIQ]; }
}

}  // namespace tg
