// tg_search.cuh -- device side of the search: one warp walks one game's tree.
//
// Reference path: mcts/tree.py (expand_node 247-270, process_mini_batch 273-315,
// generate_move_with_sequential_halving 318-356, search_by_sequential_halving 359-384,
// search_sequential_halving 387-422, search_best_move 57-105, search 130-174,
// search_mcts 199-244), mcts/node.py, selfplay/worker.py:46-90.
//
// Kernel sequence for one move of every game in the pool (host side: tg_engine.cu):
//   k_root_begin   expand the root, enqueue it as leaf 0
//   [k_scan, k_planes, <evaluator>, k_backup]           root evaluation
//   k_root_post    Gumbel noise + sequential-halving schedule (SH) / single-child shortcut (PUCT)
//   repeat: k_descend_sh | k_descend_puct, k_scan, k_planes, <evaluator>, k_backup
//   k_move_end     final choice, resign test, improved policy, play the move, scoring
#pragma once
#include "tg_board.cuh"
#include "tg_tree.cuh"

namespace tg {

enum : int {
    GS_COLOR = 0, GS_NNODES = 1, GS_PHASE = 2, GS_NPHASES = 3, GS_CONS = 4, GS_CNTS = 12,
    GS_NLEAF = 20, GS_NUNIQ = 21, GS_SLOT0 = 22, GS_ERROR = 23, GS_DONE = 24, GS_DESC = 25,
    GS_PASSCNT = 26, GS_NMOVES = 27, GS_FINISHED = 28, GS_NEVER_RESIGN = 29, GS_LAST_MOVE = 30,
    GS_ROOT_K = 31, GS_ACTIVE = 32, GS_WINNER = 33, GS_RESIGNED = 34, GS_SCORE = 35, GS_ROOTPASS = 36,
    GS_LAST_COLOR = 37, GS_EVALS = 38, GS_UEVALS = 39, GS_PREVLEN = 40, GS_SNAPLV = 41, GS_STRIDE = 48
};
enum : int { ERR_DEPTH = 1, ERR_HISTORY = 2, ERR_NODES = 4, ERR_QUEUE = 8 };
enum : int { MODE_SH = 0, MODE_PUCT = 1 };

constexpr int SEARCH_WARPS = 4;           // warps (= games) per block in the search kernels
constexpr int PATH_NODE_SHIFT = 12;       // path entry: node << 12 | child

// Everything the search kernels need, passed by value.
struct Dev {
    int games, cap, max_depth, superko, cgos, dedup;
    int scoring;             // 0: GoBoard.count_score (the reference), 1: Tromp-Taylor area score (tg_config.scoring)
    u64 seed;
    // root boards
    uint8_t* b_color; uint16_t* b_chain; unsigned* b_bloom; u64* b_hash; int* b_scal; u64* hist_hash; int16_t* hist_pos;
    TreePool tree;
    int* gs;                 // [games][GS_STRIDE]
    u64* game_id;            // [games]
    // leaf queue
    unsigned* path;          // [games][cap][max_depth]
    int* path_len;           // [games][cap]
    int* leaf_node;          // [games][cap]  node whose priors the evaluation fills (-1: none)
    int* leaf_slot;          // [games][cap]  slot of the leaf relative to the game's slot base
    uint8_t* leaf_flag;      // [games][cap]  deferred expansion (tg_block.cuh): what k_expand_leaves_blk still owes the leaf
    uint8_t* snap;           // [games][cap][Snap::BYTES]  one per unique slot
    // evaluator batch
    float* planes; float* policy; float* value; int* n_slots;
    int* slot_src;           // [slot_cap] snapshot index (game * cap + unique leaf) of every evaluator slot (fused plane load)
    int slot_cap;
    // per-move outputs (read back by the host)
    int16_t* out_action;     // [games][AP]
    double* out_improved;    // [games][AP]
    int* out_visits;         // [games][AP]
    // per-game record ring (sgf/selfplay_record.py:45-64 save_record): one row per root move of the running game
    int rec_moves;           // rows per game (2 N^2, the move limit of selfplay/worker.py:44); 0 = no ring
    int16_t* rec_move;       // [games][rec_moves]
    uint8_t* rec_color;      // [games][rec_moves]
    int16_t* rec_k;          // [games][rec_moves]      root.get_num_children()
    int16_t* rec_action;     // [games][rec_moves][AP]  root.action
    double* rec_improved;    // [games][rec_moves][AP]  root.calculate_improved_policy()
    // training samples emitted straight from the record ring (nn/data_generator.py:89-149)
    float* smp_input;        // [sample_cap][6][N*N]
    double* smp_policy;      // [sample_cap][A]
    int* smp_value;          // [sample_cap]
    // constants
    const u64* zob; const uint8_t* eye;
    long long* prof;         // optional clock64 accumulators of game 0 (development: TG_PROF=1)
    // board snapshots along the previous descent's path (k_descend_puct_snap): level l = the board after (l + 1) * SNAP_K plies
    uint32_t* snapb;         // [games][snap_levels][snap_words]
    int snap_levels, snap_words;
};

template <int N> struct WarpSmem {
    WBoard<N> root;
    WBoard<N> scratch;
    alignas(16) WAnalysis<N> an;     // expansion scratch; PUCT selection stages the child rows here (never live at the same time)
    alignas(16) double s0[Geo<N>::AP];
    double s1[Geo<N>::AP];
    int16_t memo[Geo<N>::AP];        // sequential halving: root child -> first leaf of this phase that went through it
    alignas(16) int sthdr[H_STRIDE]; // PUCT: staged node header
};

template <int N> __device__ __forceinline__ BoardPool<N> pool_of(const Dev& D)
{
    BoardPool<N> p;
    p.color = D.b_color; p.chain = D.b_chain; p.bloom = D.b_bloom; p.hash = D.b_hash; p.scal = D.b_scal;
    p.hist_hash = D.hist_hash; p.hist_pos = D.hist_pos;
    return p;
}

// mcts/tree.py:247-270 + node.py:41-72: allocate the next node, list candidates (PASS last), tentative priors.
template <int N>
__device__ inline int expand_node(const Dev& D, const Tree& t, int g, int* gs, const WBoard<N>& b, WAnalysis<N>& an, const BScal& s,
                                  int color, const u64* hist_hash, unsigned move_key, int lane)
{
    using G = Geo<N>;
    const int idx = gs[GS_NNODES];
    if (idx >= D.tree.max_nodes) { if (lane == 0) gs[GS_ERROR] |= ERR_NODES; __syncwarp(); return -1; }
    const size_t row = (size_t)idx * G::AP;
    int k = 0;
    wb_analyze<N>(b, an, s, color, D.superko != 0, D.zob, D.eye, hist_hash, lane,
        [&](int, int, int pos, bool legal, int satari, bool eye) {
            const bool cand = legal && satari < 7 && !eye;                          // tree.py:261-263
            const unsigned m = __ballot_sync(0xffffffffu, cand);
            if (cand) t.action[row + k + __popc(m & ((1u << lane) - 1))] = (int16_t)pos;
            k += __popc(m);
        });
    if (lane == 0) t.action[row + k] = PASS;                                        // tree.py:264
    k++;
    // get_tentative_policy (tree.py:509-519): Dirichlet(1,..,1) draw from the counter-based stream
    double part = 0.0;
    const u64 gid = D.game_id[g];
    for (int i = lane; i < k; i += 32) {
        const double e = dsub(0.0, det_log(noise_u(D.seed, gid, move_key, (unsigned)idx, 0u, (unsigned)i)));
        t.cpol[row + i] = e;
        part = dadd(part, e);
    }
    const double sum = warp_shape_sum(part);
    for (int i = lane; i < G::AP; i += 32) {
        if (i < k) t.cpol[row + i] = ddiv(t.cpol[row + i], sum); else { t.cpol[row + i] = 0.0; t.action[row + i] = 0; }
        t.cidx[row + i] = NOT_EXPANDED; t.cval[row + i] = 0.0f; t.cvis[row + i] = 0; t.cvl[row + i] = 0; t.cvsum[row + i] = 0.0f;
    }
    if (lane < H_STRIDE) t.hdr[(size_t)idx * H_STRIDE + lane] = (lane == H_K) ? k : 0;
    if (lane == 0) gs[GS_NNODES] = idx + 1;
    __syncwarp();
    return idx;
}

// mcts/batch_data.py:18-27: append a leaf to the game's queue.
//   dup_of >= 0: the leaf is the same position as queue entry dup_of (same path).  With dedup it shares that entry's
//   evaluator slot (result-preserving), otherwise it gets its own slot with a copy of the snapshot.
//   dup_of == -1 && search_dups: look for an identical path among the earlier entries (PUCT batches).
template <int N>
__device__ inline void push_leaf(const Dev& D, int g, int* gs, const WBoard<N>& b, const BScal& s, int color,
                                 const unsigned* cur_path, int plen, int node_index, int lane, int dup_of = -1, bool search_dups = false)
{
    const int i = gs[GS_NLEAF];
    if (i >= D.cap) { if (lane == 0) gs[GS_ERROR] |= ERR_QUEUE; __syncwarp(); return; }
    const size_t q = (size_t)g * D.cap;
    if (dup_of < 0 && search_dups && D.dedup) {
        for (int j = 0; j < i && dup_of < 0; j++) {
            if (D.path_len[q + j] != plen) continue;
            const unsigned* pj = D.path + (q + j) * D.max_depth;
            bool diff = false;
            for (int d = lane; d < plen; d += 32) diff |= (pj[d] != cur_path[d]);
            if (!__any_sync(0xffffffffu, diff)) dup_of = j;
        }
    }
    int nu = gs[GS_NUNIQ], slot;
    if (dup_of >= 0 && D.dedup) slot = D.leaf_slot[q + dup_of];
    else {
        slot = nu++;
        uint8_t* dst = D.snap + (q + slot) * Snap<N>::BYTES;
        if (dup_of >= 0) {
            const uint4* src = reinterpret_cast<const uint4*>(D.snap + (q + D.leaf_slot[q + dup_of]) * Snap<N>::BYTES);
            for (int k = lane; k < Snap<N>::BYTES / 16; k += 32) reinterpret_cast<uint4*>(dst)[k] = src[k];
        } else wb_snapshot<N>(b, s, color, D.hist_pos + (size_t)g * Geo<N>::MAXREC, dst, lane);
    }
    if (lane == 0) {
        D.path_len[q + i] = plen; D.leaf_node[q + i] = node_index; D.leaf_slot[q + i] = slot;
        gs[GS_NLEAF] = i + 1; gs[GS_NUNIQ] = nu;
    }
    __syncwarp();
}

template <int N>
__device__ __forceinline__ WarpSmem<N>& warp_smem()
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    return reinterpret_cast<WarpSmem<N>*>(smem_raw)[threadIdx.x >> 5];
}

// ---------------------------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(SEARCH_WARPS * 32) k_root_begin(Dev D)
{
    using G = Geo<N>;
    const int g = blockIdx.x * SEARCH_WARPS + (threadIdx.x >> 5), lane = lane_id();
    if (g >= D.games) return;
    int* gs = D.gs + (size_t)g * GS_STRIDE;
    if (lane == 0) { gs[GS_NLEAF] = 0; gs[GS_NUNIQ] = 0; }
    __syncwarp();
    if (!gs[GS_ACTIVE] || gs[GS_FINISHED]) return;
    WarpSmem<N>& sm = warp_smem<N>();
    BScal s;
    wb_load<N>(sm.root, s, pool_of<N>(D), g, lane);
    const Tree t = tree_of<G::AP>(D.tree, g);
    if (lane == 0) {
        gs[GS_NNODES] = 0; gs[GS_PHASE] = 0; gs[GS_NPHASES] = 0; gs[GS_DONE] = 0; gs[GS_DESC] = 0; gs[GS_ROOTPASS] = 0;
        gs[GS_ERROR] = 0; gs[GS_EVALS] = 0; gs[GS_UEVALS] = 0;
        gs[GS_PREVLEN] = 0; gs[GS_SNAPLV] = 0;               // board snapshots of k_descend_puct_snap belong to the old root
    }
    __syncwarp();
    const int color = gs[GS_COLOR];
    const u64* hh = D.hist_hash + (size_t)g * G::MAXREC;
    const int root = expand_node<N>(D, t, g, gs, sm.root, sm.an, s, color, hh, (unsigned)s.moves, lane);   // tree.py:332 / 52
    if (root < 0) return;
    for (int i = lane; i < G::AP; i += 32) t.noise[i] = 0.0;                      // node.py:57
    push_leaf<N>(D, g, gs, sm.root, s, color, nullptr, 0, root, lane);            // tree.py:333-334 / 53-54
}

// After the root evaluation: Gumbel noise and the halving schedule (tree.py:336, 370-373), or the
// single-candidate shortcut of search_best_move (tree.py:76-77).
template <int N>
__global__ void __launch_bounds__(SEARCH_WARPS * 32) k_root_post(Dev D, int mode, int visits)
{
    using G = Geo<N>;
    const int g = blockIdx.x * SEARCH_WARPS + (threadIdx.x >> 5), lane = lane_id();
    if (g >= D.games) return;
    int* gs = D.gs + (size_t)g * GS_STRIDE;
    if (!gs[GS_ACTIVE] || gs[GS_FINISHED] || gs[GS_ERROR]) return;
    const Tree t = tree_of<G::AP>(D.tree, g);
    const int k = t.hdr[H_K];
    if (lane == 0) gs[GS_ROOT_K] = k;
    if (mode == MODE_SH) {
        const u64 gid = D.game_id[g];
        const unsigned mv = (unsigned)D.b_scal[(size_t)g * 8 + 0];
        for (int i = lane; i < G::A; i += 32) {                                   // node.py:275-278: gumbel(size=MAX_ACTIONS)
            const double e = dsub(0.0, det_log(noise_u(D.seed, gid, mv, 0u, 1u, (unsigned)i)));
            t.noise[i] = dsub(0.0, det_log(e));
        }
        if (lane == 0) {
            int cons[8], cnts[8];
            const int np = sh_schedule(min(k, MAX_CONSIDERED), visits, cons, cnts, 8);
            for (int i = 0; i < 8; i++) { gs[GS_CONS + i] = i < np ? cons[i] : 0; gs[GS_CNTS + i] = i < np ? cnts[i] : 0; }
            gs[GS_NPHASES] = np; gs[GS_PHASE] = 0;
        }
    } else if (lane == 0 && k == 1) { gs[GS_DONE] = 1; gs[GS_ROOTPASS] = 1; }
}

// ---------------------------------------------------------------------------------------------
// One phase of search_by_sequential_halving (tree.py:375-384): cnt rounds x cons descents, all enqueued.
template <int N>
__global__ void __launch_bounds__(SEARCH_WARPS * 32) k_descend_sh(Dev D)
{
    using G = Geo<N>;
    const int g = blockIdx.x * SEARCH_WARPS + (threadIdx.x >> 5), lane = lane_id();
    if (g >= D.games) return;
    int* gs = D.gs + (size_t)g * GS_STRIDE;
    if (lane == 0) { gs[GS_NLEAF] = 0; gs[GS_NUNIQ] = 0; }
    __syncwarp();
    if (!gs[GS_ACTIVE] || gs[GS_FINISHED] || gs[GS_ERROR]) return;
    const int phase = gs[GS_PHASE];
    if (phase >= gs[GS_NPHASES]) return;
    const int cons = gs[GS_CONS + phase], cnt = gs[GS_CNTS + phase];
    WarpSmem<N>& sm = warp_smem<N>();
    BScal rs;
    wb_load<N>(sm.root, rs, pool_of<N>(D), g, lane);
    const Tree t = tree_of<G::AP>(D.tree, g);
    const int root_color = gs[GS_COLOR];
    u64* hh = D.hist_hash + (size_t)g * G::MAXREC;
    int16_t* hp = D.hist_pos + (size_t)g * G::MAXREC;
    const unsigned move_key = (unsigned)rs.moves;
    // Inside one phase nothing below the root changes (backups happen at the end of the phase, tree.py:384, and
    // the non-root rule reads visit counts and value sums only, node.py:349-361), so every descent through a given
    // root child follows the same path to the same leaf.  The first descent through a child walks the board; later
    // ones replay its path: same virtual-loss updates, same queue entry (tree.py:387-422), no board work.
    for (int i = lane; i < G::AP; i += 32) sm.memo[i] = -1;
    __syncwarp();
    for (int thr = 1; thr <= cnt; thr++) {
        for (int j = 0; j < cons; j++) {
            const int li = gs[GS_NLEAF];
            if (li >= D.cap) { if (lane == 0) gs[GS_ERROR] |= ERR_QUEUE; __syncwarp(); return; }
            unsigned* path = D.path + ((size_t)g * D.cap + li) * D.max_depth;
            const int first = select_sh_root<G::AP>(t, 0, thr, lane);
            const int m = sm.memo[first];
            if (m >= 0) {
                const size_t qm = (size_t)g * D.cap + m;
                const int plen = D.path_len[qm];
                const unsigned* src = D.path + qm * D.max_depth;
                for (int d = lane; d < plen; d += 32) {
                    const unsigned e = src[d];
                    path[d] = e;
                    const int node = (int)(e >> PATH_NODE_SHIFT), c = (int)(e & ((1u << PATH_NODE_SHIFT) - 1));
                    t.hdr[(size_t)node * H_STRIDE + H_VL] += 1; t.cvl[(size_t)node * G::AP + c] += 1;   // node.py:76-83
                }
                __syncwarp();
                push_leaf<N>(D, g, gs, sm.scratch, rs, root_color, path, plen, D.leaf_node[qm], lane, m);
                continue;
            }
            wb_copy<N>(sm.scratch, sm.root, lane);                               // tree.py:378
            BScal s = rs;
            int color = root_color, cur = 0, plen = 0;
            for (;;) {                                                           // tree.py:387-422
                const int next = (cur == 0) ? first : select_sh_node<G::AP>(t, cur, sm.s0, sm.s1, lane);
                const size_t row = (size_t)cur * G::AP;
                const int mv = t.action[row + next];
                if (lane == 0) path[plen] = ((unsigned)cur << PATH_NODE_SHIFT) | (unsigned)next;
                plen++;
                wb_put_stone<N>(sm.scratch, s, mv, color, D.zob, hh, hp, lane);  // :407
                color = opp(color);
                if (lane == 0) { t.hdr[(size_t)cur * H_STRIDE + H_VL] += 1; t.cvl[row + next] += 1; }   // node.py:76-83
                __syncwarp();
                if (t.cvis[row + next] < 1) {                                    // :412-416
                    push_leaf<N>(D, g, gs, sm.scratch, s, color, path, plen, t.cidx[row + next], lane);
                    if (lane == 0) sm.memo[first] = (int16_t)li;
                    __syncwarp();
                    break;
                }
                int ci = t.cidx[row + next];
                if (ci == NOT_EXPANDED) {                                        // :418-420
                    ci = expand_node<N>(D, t, g, gs, sm.scratch, sm.an, s, color, hh, move_key, lane);
                    if (ci < 0) return;
                    if (lane == 0) t.cidx[row + next] = ci;
                    __syncwarp();
                }
                cur = ci;
                if (plen >= D.max_depth) { if (lane == 0) gs[GS_ERROR] |= ERR_DEPTH; __syncwarp(); return; }
            }
        }
    }
    if (lane == 0) gs[GS_PHASE] = phase + 1;
}

// Up to `batch` descents of search_mcts (tree.py:199-244) with the early stop of TimeManager.is_move_decided
// (time_manager.py:146-163) checked after every descent as in MCTSTree.search (tree.py:146-152).
//   Every ply reads its node from shared memory: the rows are staged with one burst of asynchronous copies that is issued
//   as soon as the node is known -- the root's at kernel start (under the board load) and at the end of a descent, a
//   child's right after the selection that leads to it (under put_stone).
template <int N>
__global__ void __launch_bounds__(SEARCH_WARPS * 32) k_descend_puct(Dev D, int visits, int batch, int strict, int nsqrt)
{
    using G = Geo<N>;
    const int g = blockIdx.x * SEARCH_WARPS + (threadIdx.x >> 5), lane = lane_id();
    // the Zobrist keys (one read per put_stone, ~100 plies per descent) sit behind the per-warp scratch, shared by the CTA,
    // and behind them a table of sqrt(n) for the selections
    extern __shared__ __align__(16) unsigned char smem_raw_[];
    u64* zs = reinterpret_cast<u64*>(smem_raw_ + sizeof(WarpSmem<N>) * SEARCH_WARPS);
    double* sqtab = reinterpret_cast<double*>(zs + 4 * G::CP);
    for (int i = threadIdx.x; i < 4 * G::CELLS; i += SEARCH_WARPS * 32) zs[i] = D.zob[i];
    for (int i = threadIdx.x; i < nsqrt; i += SEARCH_WARPS * 32) sqtab[i] = sqrt((double)i);
    __syncthreads();
    if (g >= D.games) return;
    int* gs = D.gs + (size_t)g * GS_STRIDE;
    if (lane == 0) { gs[GS_NLEAF] = 0; gs[GS_NUNIQ] = 0; }
    __syncwarp();
    if (!gs[GS_ACTIVE] || gs[GS_FINISHED] || gs[GS_ERROR] || gs[GS_DONE]) return;
    WarpSmem<N>& sm = warp_smem<N>();
    static_assert(sizeof(WAnalysis<N>) >= 4 * G::AP * 4 + G::AP * 2, "child-row staging aliases the analysis scratch");
    // staging lives in the expansion scratch (never live at the same time: a descent expands only at its last ply)
    const SelStage stage = { sm.s0, reinterpret_cast<int*>(&sm.an), reinterpret_cast<int*>(&sm.an) + G::AP,
                             reinterpret_cast<float*>(&sm.an) + 2 * G::AP, reinterpret_cast<int*>(&sm.an) + 3 * G::AP,
                             reinterpret_cast<int16_t*>(reinterpret_cast<int*>(&sm.an) + 4 * G::AP), sm.sthdr };
    const Tree t = tree_of<G::AP>(D.tree, g);
    stage_node_warp<G::AP>(stage, t, 0, lane);                                    // root rows arrive under the board load
    BScal rs;
    wb_load<N>(sm.root, rs, pool_of<N>(D), g, lane);
    const int root_color = gs[GS_COLOR];
    u64* hh = D.hist_hash + (size_t)g * G::MAXREC;
    int16_t* hp = D.hist_pos + (size_t)g * G::MAXREC;
    const unsigned move_key = (unsigned)rs.moves;
    for (int b = 0; b < batch; b++) {
        int desc = gs[GS_DESC];
        if (desc >= visits) { if (lane == 0) gs[GS_DONE] = 1; __syncwarp(); break; }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
        if (desc > 0) {
            // is_move_decided (time_manager.py:146-163), evaluated after the previous descent and after the
            // mini-batch flush that descent may have triggered (tree.py:149-152, 240-241):
            // sorted(children_visits)[-1] - [-2] against the remaining budget
            const int k = stage.hdr[H_K];
            int top1 = 0;
            for (int i = lane; i < k; i += 32) top1 = max(top1, stage.vis[i]);
            top1 = warp_max_i(top1);
            int nmax = 0, top2 = 0;
            for (int i = lane; i < k; i += 32) { const int v = stage.vis[i]; if (v == top1) nmax++; else top2 = max(top2, v); }
            nmax = warp_sum_i(nmax); top2 = warp_max_i(top2);
            if (nmax >= 2) top2 = top1;
            const int remaining = visits - stage.hdr[H_NV];
            const int cutoff = strict ? 0 : top1 - top2;
            if (remaining < cutoff) { if (lane == 0) gs[GS_DONE] = 1; __syncwarp(); break; }
        }
        const bool prof = D.prof && g == 0 && lane == 0;
        long long pt0 = prof ? clock64() : 0;
        wb_copy<N>(sm.scratch, sm.root, lane);                                   // tree.py:147
        if (prof) { const long long c = clock64(); D.prof[0] += c - pt0; pt0 = c; }
        BScal s = rs;
        int color = root_color, cur = 0, plen = 0;
        unsigned* path = D.path + ((size_t)g * D.cap + gs[GS_NLEAF]) * D.max_depth;
        bool fail = false;
        for (;;) {
            const int next = select_puct(stage, D.cgos != 0, lane, sqtab, nsqrt);       // :213
            if (prof) { const long long c = clock64(); D.prof[1] += c - pt0; pt0 = c; D.prof[5]++; }
            const size_t row = (size_t)cur * G::AP;
            const int mv = stage.action[next];
            const int cv_before = stage.vis[next] + stage.vl[next];
            int ci = stage.cidx[next];
            const int vl_node = stage.hdr[H_VL], vl_edge = stage.vl[next];
            __syncwarp();                                                        // every lane has read the staged node
            // the child's rows are needed next unless this edge ends the descent: fetch them under put_stone
            const bool spec = ci != NOT_EXPANDED && cv_before >= 1;
            if (spec) stage_node_warp<G::AP>(stage, t, ci, lane);
            if (lane == 0) {
                path[plen] = ((unsigned)cur << PATH_NODE_SHIFT) | (unsigned)next;
                t.hdr[(size_t)cur * H_STRIDE + H_VL] = vl_node + 1; t.cvl[row + next] = vl_edge + 1;       // :221 add_virtual_loss
            }
            plen++;
            const int pris_before = s.pris0 + s.pris1;
            wb_put_stone<N>(sm.scratch, s, mv, color, zs, hh, hp, lane);         // :217
            if (prof) { const long long c = clock64(); D.prof[2] += c - pt0;
                        if (s.pris0 + s.pris1 != pris_before) { D.prof[12]++; D.prof[13] += c - pt0; } else if (mv == PASS) { D.prof[14]++; D.prof[15] += c - pt0; }
                        pt0 = c; }
            color = opp(color);
            int expand_threshold = 1;
            if (s.moves > 2) {                                                   // :224-229
                if (s.moves - 1 >= G::MAXREC) { if (lane == 0) gs[GS_ERROR] |= ERR_HISTORY; fail = true; break; }
                if (hp[s.moves - 1] == PASS && hp[s.moves - 2] == PASS) expand_threshold = 10000000;
            }
            if (cv_before + 1 < expand_threshold + 1) {                          // :231-241
                if (spec) { asm volatile("cp.async.wait_all;" ::: "memory"); __syncwarp(); }   // two-pass rule: drain the unused fetch
                if (ci == NOT_EXPANDED) {
                    if (prof) pt0 = clock64();
                    ci = expand_node<N>(D, t, g, gs, sm.scratch, sm.an, s, color, hh, move_key, lane);
                    if (prof) { const long long c = clock64(); D.prof[3] += c - pt0; pt0 = c; D.prof[6]++; }
                    if (ci < 0) { fail = true; break; }
                    if (lane == 0) t.cidx[row + next] = ci;
                    __threadfence_block();
                    __syncwarp();
                }
                // the expansion scratch is free again: the root's rows for the next descent arrive under push_leaf
                if (b + 1 < batch) stage_node_warp<G::AP>(stage, t, 0, lane);
                if (prof) pt0 = clock64();
                push_leaf<N>(D, g, gs, sm.scratch, s, color, path, plen, ci, lane, -1, true);
                if (prof) { const long long c = clock64(); D.prof[4] += c - pt0; pt0 = c; D.prof[7]++; }
                break;
            }
            cur = ci;
            if (plen >= D.max_depth) { if (lane == 0) gs[GS_ERROR] |= ERR_DEPTH; fail = true; break; }
        }
        __syncwarp();
        if (fail) break;
        if (lane == 0) gs[GS_DESC] = desc + 1;
        __syncwarp();
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
}

// One descent per launch (PUCT with batch 1, BASELINE configs[3]) WITHOUT replaying the whole path on the board.
//   With one evaluation per descent nothing carries a virtual loss between descents, so consecutive descents of a game
//   share almost their entire path (with a random-init net the tree is a ~100-ply chain: the previous path plus one ply),
//   and k_descend_puct spends more than half of its time re-playing those same ~100 stones (put_stone: 4 k cycles per ply
//   on a crowded 19x19 board).  Here the descent first walks the tree alone -- selections need no board -- and compares its
//   path with the previous one; then it restores the deepest board snapshot that lies on the shared prefix (one snapshot
//   every SNAP_K plies of the previous path, kept in HBM/L2), replays only the plies behind it (saving the snapshots it
//   passes), and expands the leaf.  Same tree, same boards, same record entries as k_descend_puct (bit-identical results).
constexpr int SNAP_K = 8;

template <int N> __device__ __forceinline__ void snap_store(uint32_t* dst, const WBoard<N>& b, const BScal& s, int lane)
{
    constexpr int W = (int)(sizeof(WBoard<N>) / 4);
    const uint32_t* a = reinterpret_cast<const uint32_t*>(&b);
    for (int i = lane; i < W; i += 32) dst[i] = a[i];
    if (lane == 0) {
        dst[W] = (uint32_t)s.hash; dst[W + 1] = (uint32_t)(s.hash >> 32); dst[W + 2] = (uint32_t)s.moves; dst[W + 3] = (uint32_t)s.ko_pos;
        dst[W + 4] = (uint32_t)s.ko_move; dst[W + 5] = (uint32_t)s.pris0; dst[W + 6] = (uint32_t)s.pris1;
    }
}
template <int N> __device__ __forceinline__ void snap_load(WBoard<N>& b, BScal& s, const uint32_t* src, int lane)
{
    constexpr int W = (int)(sizeof(WBoard<N>) / 4);
    uint32_t* d = reinterpret_cast<uint32_t*>(&b);
    for (int i = lane; i < W; i += 32) d[i] = src[i];
    s.hash = (u64)src[W] | ((u64)src[W + 1] << 32); s.moves = (int)src[W + 2]; s.ko_pos = (int)src[W + 3];
    s.ko_move = (int)src[W + 4]; s.pris0 = (int)src[W + 5]; s.pris1 = (int)src[W + 6];
    __syncwarp();
}

template <int N>
__global__ void __launch_bounds__(SEARCH_WARPS * 32) k_descend_puct_snap(Dev D, int visits, int strict, int nsqrt)
{
    using G = Geo<N>;
    static_assert(sizeof(WBoard<N>) == 7 * G::CP + 4 * BLOOM_WORDS, "snapshot stride (tg_engine.cu: snap_words)");
    constexpr int PMAX = G::AP * 2;                          // path entries kept in shared memory (the s1 scratch, 8 bytes per child)
    const int g = blockIdx.x * SEARCH_WARPS + (threadIdx.x >> 5), lane = lane_id();
    extern __shared__ __align__(16) unsigned char smem_raw_[];
    u64* zs = reinterpret_cast<u64*>(smem_raw_ + sizeof(WarpSmem<N>) * SEARCH_WARPS);
    double* sqtab = reinterpret_cast<double*>(zs + 4 * G::CP);
    for (int i = threadIdx.x; i < 4 * G::CELLS; i += SEARCH_WARPS * 32) zs[i] = D.zob[i];
    for (int i = threadIdx.x; i < nsqrt; i += SEARCH_WARPS * 32) sqtab[i] = sqrt((double)i);
    __syncthreads();
    if (g >= D.games) return;
    int* gs = D.gs + (size_t)g * GS_STRIDE;
    if (lane == 0) { gs[GS_NLEAF] = 0; gs[GS_NUNIQ] = 0; }
    __syncwarp();
    if (!gs[GS_ACTIVE] || gs[GS_FINISHED] || gs[GS_ERROR] || gs[GS_DONE]) return;
    WarpSmem<N>& sm = warp_smem<N>();
    const SelStage stage = { sm.s0, reinterpret_cast<int*>(&sm.an), reinterpret_cast<int*>(&sm.an) + G::AP,
                             reinterpret_cast<float*>(&sm.an) + 2 * G::AP, reinterpret_cast<int*>(&sm.an) + 3 * G::AP,
                             reinterpret_cast<int16_t*>(reinterpret_cast<int*>(&sm.an) + 4 * G::AP), sm.sthdr };
    unsigned* spath = reinterpret_cast<unsigned*>(sm.s1);    // previous path, overwritten ply by ply with the new one
    const Tree t = tree_of<G::AP>(D.tree, g);
    stage_node_warp<G::AP>(stage, t, 0, lane);
    BScal rs;
    {                                                        // (the root board itself is only loaded if no snapshot can be used)
        const int* sc = D.b_scal + (size_t)g * 8;
        rs.hash = D.b_hash[g];
        rs.moves = sc[0]; rs.ko_pos = sc[1]; rs.ko_move = sc[2]; rs.pris0 = sc[3]; rs.pris1 = sc[4];
    }
    const int root_color = gs[GS_COLOR];
    u64* hh = D.hist_hash + (size_t)g * G::MAXREC;
    int16_t* hp = D.hist_pos + (size_t)g * G::MAXREC;
    const unsigned move_key = (unsigned)rs.moves;
    unsigned* path = D.path + (size_t)g * D.cap * D.max_depth;                   // (queue entry 0: one leaf per launch)
    const int prevlen = min(gs[GS_PREVLEN], PMAX), snaplv = gs[GS_SNAPLV];
    for (int i = lane; i < prevlen; i += 32) spath[i] = path[i];
    const int desc = gs[GS_DESC];
    if (desc >= visits) { if (lane == 0) gs[GS_DONE] = 1; __syncwarp(); asm volatile("cp.async.wait_all;" ::: "memory"); return; }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    if (desc > 0) {                                          // is_move_decided (time_manager.py:146-163), as k_descend_puct
        const int k = stage.hdr[H_K];
        int top1 = 0;
        for (int i = lane; i < k; i += 32) top1 = max(top1, stage.vis[i]);
        top1 = warp_max_i(top1);
        int nmax = 0, top2 = 0;
        for (int i = lane; i < k; i += 32) { const int v = stage.vis[i]; if (v == top1) nmax++; else top2 = max(top2, v); }
        nmax = warp_sum_i(nmax); top2 = warp_max_i(top2);
        if (nmax >= 2) top2 = top1;
        const int remaining = visits - stage.hdr[H_NV];
        const int cutoff = strict ? 0 : top1 - top2;
        if (remaining < cutoff) { if (lane == 0) gs[GS_DONE] = 1; __syncwarp(); return; }
    }
    const bool prof = D.prof && g == 0 && lane == 0;
    long long pt0 = prof ? clock64() : 0;
    // ---- the walk (tree.py:213-241 without the board).  The rows of the next node are fetched after the selection that leads
    // to it and waited for (≈ 3 k cycles per ply at 19x19: what put_stone used to hide).  Fetching the node the previous path
    // visited at the next depth DURING the selection was measured slower (select 3.0 k -> 7.0 k cycles per ply, step 315 -> 343 ms).
    int cur = 0, plen = 0, shared = 0, m1 = (rs.moves >= 1 && rs.moves - 1 < G::MAXREC) ? hp[rs.moves - 1] : -1;
    int ci = NOT_EXPANDED, next = 0;
    size_t row = 0;
    for (;;) {
        next = select_puct(stage, D.cgos != 0, lane, sqtab, nsqrt);
        if (prof) { const long long c = clock64(); D.prof[1] += c - pt0; pt0 = c; D.prof[5]++; }
        row = (size_t)cur * G::AP;
        const int mv = stage.action[next];
        const int cv_before = stage.vis[next] + stage.vl[next];
        ci = stage.cidx[next];
        const int vl_node = stage.hdr[H_VL], vl_edge = stage.vl[next];
        const unsigned entry = ((unsigned)cur << PATH_NODE_SHIFT) | (unsigned)next;
        const bool same = shared == plen && plen < prevlen && spath[plen] == entry;
        __syncwarp();                                        // every lane has read the staged node and the old path entry
        if (same) shared++;
        const bool spec = ci != NOT_EXPANDED && cv_before >= 1;
        if (spec) stage_node_warp<G::AP>(stage, t, ci, lane);
        if (lane == 0) {
            path[plen] = entry;
            if (plen < PMAX) spath[plen] = entry;
            t.hdr[(size_t)cur * H_STRIDE + H_VL] = vl_node + 1; t.cvl[row + next] = vl_edge + 1;           // :221 add_virtual_loss
        }
        plen++;
        const int moves = rs.moves + plen;
        int expand_threshold = 1;
        if (moves > 2) {                                     // :224-229
            if (moves - 1 >= G::MAXREC) { asm volatile("cp.async.wait_all;" ::: "memory"); if (lane == 0) gs[GS_ERROR] |= ERR_HISTORY; __syncwarp(); return; }
            if (mv == PASS && m1 == PASS) expand_threshold = 10000000;
        }
        m1 = mv;
        if (cv_before + 1 < expand_threshold + 1) {          // :231-241: this edge ends the descent
            if (spec) { asm volatile("cp.async.wait_all;" ::: "memory"); }
            break;
        }
        cur = ci;
        if (plen >= D.max_depth) { asm volatile("cp.async.wait_all;" ::: "memory"); if (lane == 0) gs[GS_ERROR] |= ERR_DEPTH; __syncwarp(); return; }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
        if (prof) { const long long c = clock64(); D.prof[10] += c - pt0; pt0 = c; }
    }
    __syncwarp();
    // ---- the board at the leaf: deepest snapshot on the shared prefix, then the plies behind it
    const int want = min(min(shared / SNAP_K, snaplv), D.snap_levels);
    uint32_t* snaps = D.snapb + (size_t)g * D.snap_levels * D.snap_words;
    BScal s = rs;
    if (want == 0) wb_load<N>(sm.scratch, s, pool_of<N>(D), g, lane);            // tree.py:147 (straight from the root pool)
    else snap_load<N>(sm.scratch, s, snaps + (size_t)(want - 1) * D.snap_words, lane);
    if (prof) { const long long c = clock64(); D.prof[0] += c - pt0; pt0 = c; }
    for (int d0 = want * SNAP_K; d0 < plen; d0 += 32) {
        const int dl = d0 + lane;
        int mvl = 0;
        if (dl < plen) {
            const unsigned e = dl < PMAX ? spath[dl] : path[dl];
            mvl = t.action[(size_t)(e >> PATH_NODE_SHIFT) * G::AP + (e & ((1u << PATH_NODE_SHIFT) - 1))];
        }
        const int n = min(32, plen - d0);
        for (int i = 0; i < n; i++) {
            const int d = d0 + i;
            const int mv = __shfl_sync(0xffffffffu, mvl, i);
            wb_put_stone<N>(sm.scratch, s, mv, (d & 1) ? opp(root_color) : root_color, zs, hh, hp, lane);   // :217
            const int lv = (d + 1) / SNAP_K;
            if ((d + 1) % SNAP_K == 0 && lv <= D.snap_levels) { __syncwarp(); snap_store<N>(snaps + (size_t)(lv - 1) * D.snap_words, sm.scratch, s, lane); }
        }
    }
    if (prof) { const long long c = clock64(); D.prof[2] += c - pt0; pt0 = c; }
    const int color = (plen & 1) ? opp(root_color) : root_color;                 // to move at the leaf
    if (ci == NOT_EXPANDED) {
        ci = expand_node<N>(D, t, g, gs, sm.scratch, sm.an, s, color, hh, move_key, lane);
        if (prof) { const long long c = clock64(); D.prof[3] += c - pt0; pt0 = c; D.prof[6]++; }
        if (ci < 0) return;
        if (lane == 0) t.cidx[row + next] = ci;
        __threadfence_block();
        __syncwarp();
    }
    push_leaf<N>(D, g, gs, sm.scratch, s, color, path, plen, ci, lane, -1, true);
    if (prof) { const long long c = clock64(); D.prof[4] += c - pt0; D.prof[7]++; }
    __syncwarp();
    if (lane == 0) {
        gs[GS_DESC] = desc + 1;
        gs[GS_PREVLEN] = plen;
        gs[GS_SNAPLV] = min(D.snap_levels, plen / SNAP_K);
    }
}

// ---------------------------------------------------------------------------------------------
// Exclusive scan of the per-game unique-leaf counts -> slot bases; one block.
__global__ void __launch_bounds__(1024) k_scan(Dev D)
{
    __shared__ int warp_tot[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < D.games; base += 1024) {
        const int g = base + threadIdx.x;
        const int v = g < D.games ? D.gs[(size_t)g * GS_STRIDE + GS_NUNIQ] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
        if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = warp_tot[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, o); if (threadIdx.x >= o) w += y; }
            warp_tot[threadIdx.x] = w;
        }
        __syncthreads();
        const int prefix = carry + (threadIdx.x >= 32 ? warp_tot[(threadIdx.x >> 5) - 1] : 0) + x - v;
        if (g < D.games) D.gs[(size_t)g * GS_STRIDE + GS_SLOT0] = prefix;
        __syncthreads();
        if (threadIdx.x == 1023) carry = prefix + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *D.n_slots = min(carry, D.slot_cap);
}

// Slot -> snapshot map for the tensor-core evaluator, which builds the feature planes itself from the leaf snapshots
// (tg_dualnet.cuh, fused plane load): 4 bytes per slot instead of the 6 N^2 fp32 planes k_planes writes.
__global__ void __launch_bounds__(256) k_slotmap(Dev D)
{
    const int g = blockIdx.x;
    const int* gs = D.gs + (size_t)g * GS_STRIDE;
    const int nu = gs[GS_NUNIQ], slot0 = gs[GS_SLOT0];
    for (int u = threadIdx.x; u < nu; u += 256)
        if (slot0 + u < D.slot_cap) D.slot_src[slot0 + u] = g * D.cap + u;
}

// K3 feature planes (nn/feature.py:10-57): one block per game expands its leaf snapshots into fp32 planes
// [slot][6][N*N], written with coalesced 4-byte stores over the game's contiguous slot range.
template <int N>
__global__ void __launch_bounds__(256) k_planes(Dev D)
{
    using G = Geo<N>;
    const int g = blockIdx.x;
    const int* gs = D.gs + (size_t)g * GS_STRIDE;
    const int nu = gs[GS_NUNIQ], slot0 = gs[GS_SLOT0];
    if (nu == 0) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int CH = 16;                                   // snapshots staged per pass
    for (int u0 = 0; u0 < nu; u0 += CH) {
        const int nc = min(CH, nu - u0);
        const uint4* src = reinterpret_cast<const uint4*>(D.snap + ((size_t)g * D.cap + u0) * Snap<N>::BYTES);
        uint4* dst = reinterpret_cast<uint4*>(smem_raw);
        for (int i = threadIdx.x; i < nc * (Snap<N>::BYTES / 16); i += blockDim.x) dst[i] = src[i];
        __syncthreads();
        if (slot0 + u0 + nc <= D.slot_cap) {
            float* out = D.planes + (size_t)(slot0 + u0) * G::PLANES;
            // one warp per (snapshot, plane) row of N*N floats: index arithmetic once per row, coalesced stores
            const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
            for (int row = warp; row < nc * 6; row += 8) {
                const int u = row / 6, p = row - u * 6;
                const uint8_t* sn = smem_raw + u * Snap<N>::BYTES;
                const int color = sn[3];
                float* o = out + (size_t)row * G::NN;
                if (p >= 4) {                                                      // feature.py:39-41, 50-52: constant planes
                    const float v = p == 5 ? (color == WHITE ? -1.0f : 1.0f) : (sn[2] ? 1.0f : 0.0f);
                    for (int idx = lane; idx < G::NN; idx += 32) o[idx] = v;
                } else if (p == 3) {                                               // :43-46 previous move
                    const int pidx = *reinterpret_cast<const int16_t*>(sn);
                    for (int idx = lane; idx < G::NN; idx += 32) o[idx] = idx == pidx ? 1.0f : 0.0f;
                } else {                                                           // :24-31 empty / own / opponent
                    const int want = (color == WHITE && p != 0) ? 3 - p : p;
                    for (int idx = lane; idx < G::NN; idx += 32) o[idx] = sn[Snap<N>::HDR + idx] == want ? 1.0f : 0.0f;
                }
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// process_mini_batch after the forward pass (tree.py:287-315): priors/raw value of the evaluated node, then the
// value walks the path back to the root.  Leaves of a game are applied in queue order (fp32 sums depend on it);
// the plies of one path touch distinct nodes and are updated by different lanes.
template <int N>
__global__ void __launch_bounds__(SEARCH_WARPS * 32) k_backup(Dev D, int use_logit)
{
    using G = Geo<N>;
    const int g = blockIdx.x * SEARCH_WARPS + (threadIdx.x >> 5), lane = lane_id();
    if (g >= D.games) return;
    int* gs = D.gs + (size_t)g * GS_STRIDE;
    const int nl = gs[GS_NLEAF];
    if (nl == 0) return;
    const Tree t = tree_of<G::AP>(D.tree, g);
    const int slot0 = gs[GS_SLOT0];
    const size_t q = (size_t)g * D.cap;
    for (int i = 0; i < nl; i++) {
        const int slot = slot0 + D.leaf_slot[q + i];
        if (slot >= D.slot_cap) { if (lane == 0) gs[GS_ERROR] |= ERR_QUEUE; break; }
        const float* pol = D.policy + (size_t)slot * G::A;
        const float* v = D.value + (size_t)slot * 3;
        const float v0 = v[0], v1 = v[1], v2 = v[2];
        const int ni = D.leaf_node[q + i];
        if (ni >= 0) {                                                           // node.py:86-93, tree.py:287-300
            const size_t row = (size_t)ni * G::AP;
            const int k = t.hdr[(size_t)ni * H_STRIDE + H_K];
            for (int c = lane; c < k; c += 32) {
                const int a = t.action[row + c];
                float p;
                if (a == PASS) { p = pol[G::NN]; if (use_logit) p = __fsub_rn(p, 0.5f); }              // :292-294
                else p = pol[(a / G::W - 1) * N + (a % G::W - 1)];
                t.cpol[row + c] = (double)p;
            }
            if (lane == 0) t.hdr[(size_t)ni * H_STRIDE + H_RAW] = __float_as_int(__fadd_rn(__fmul_rn(v1, 0.5f), v2));   // :300
        }
        const int plen = D.path_len[q + i];
        if (plen > 0) {
            const unsigned* path = D.path + (q + i) * D.max_depth;
            const float val0 = __fadd_rn(v0, __fmul_rn(v1, 0.5f));               // :303
            const float val1 = __fsub_rn(1.0f, val0), val2 = __fsub_rn(1.0f, val1);
            for (int d0 = 0; d0 < plen; d0 += 32) {
                const int d = d0 + lane;                                         // distance from the leaf
                if (d < plen) {
                    const unsigned e = path[plen - 1 - d];
                    const int node = (int)(e >> PATH_NODE_SHIFT), c = (int)(e & ((1u << PATH_NODE_SHIFT) - 1));
                    const float val = d == 0 ? val0 : ((d & 1) ? val1 : val2);   // value = 1 - value per ply (:313)
                    const size_t row = (size_t)node * G::AP;
                    if (d == 0) t.cval[row + c] = val0;                          // :308
                    t.cvsum[row + c] = __fadd_rn(t.cvsum[row + c], val);         // node.py:118-138
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

// ---------------------------------------------------------------------------------------------
// End of a move: final choice (tree.py:344-356 / 86-105), record (selfplay_record.py:45-64) and, when `play`,
// the game loop body of selfplay/worker.py:58-87 (play the move, pass/resign/terminal handling, scoring).
template <int N>
__global__ void __launch_bounds__(SEARCH_WARPS * 32) k_move_end(Dev D, int mode, int play, float komi, int max_moves)
{
    using G = Geo<N>;
    const int g = blockIdx.x * SEARCH_WARPS + (threadIdx.x >> 5), lane = lane_id();
    if (g >= D.games) return;
    int* gs = D.gs + (size_t)g * GS_STRIDE;
    if (!gs[GS_ACTIVE] || gs[GS_FINISHED]) { if (lane == 0) gs[GS_LAST_MOVE] = -2; return; }
    if (gs[GS_ERROR]) { if (lane == 0) { gs[GS_LAST_MOVE] = -2; gs[GS_FINISHED] = 1; } return; }
    WarpSmem<N>& sm = warp_smem<N>();
    const Tree t = tree_of<G::AP>(D.tree, g);
    const int k = t.hdr[H_K];
    int next, move; bool resign = false;
    if (mode == MODE_SH) {
        next = select_sh_root<G::AP>(t, 0, PLAYOUTS, lane);                      // tree.py:344
        const double value = value_evaluation<G::AP>(t, 0, next);
        resign = !gs[GS_NEVER_RESIGN] && value < 0.05;                           // :353
        move = t.action[next];
    } else if (gs[GS_ROOTPASS]) { next = 0; move = PASS; }                      // tree.py:76-77
    else {
        next = best_visit_child<G::AP>(t, 0, lane);                              // :86-87
        resign = value_evaluation<G::AP>(t, 0, next) < 0.05;                     // :100-103
        move = t.action[next];
    }
    // record what selfplay_record.save_record reads from the root
    improved_policy<G::AP>(t, 0, sm.s0, sm.s1, lane);
    for (int i = lane; i < G::AP; i += 32) {
        D.out_action[(size_t)g * G::AP + i] = i < k ? t.action[i] : (int16_t)0;
        D.out_improved[(size_t)g * G::AP + i] = i < k ? sm.s0[i] : 0.0;
        D.out_visits[(size_t)g * G::AP + i] = i < k ? t.cvis[i] : 0;
    }
    const int color = gs[GS_COLOR];
    if (lane == 0) { gs[GS_LAST_MOVE] = resign ? RESIGN : move; gs[GS_LAST_COLOR] = color; }
    if (!play) return;
    if (resign) {                                                                // worker.py:60-63
        if (lane == 0) { gs[GS_WINNER] = opp(color); gs[GS_RESIGNED] = 1; gs[GS_FINISHED] = 1; gs[GS_SCORE] = __float_as_int(0.0f); }
        return;
    }
    if (D.rec_moves > 0 && gs[GS_NMOVES] < D.rec_moves) {                        // worker.py:72 record.save_record(root, pos, color)
        const size_t r = (size_t)g * D.rec_moves + gs[GS_NMOVES];
        for (int i = lane; i < G::AP; i += 32) {
            D.rec_action[r * G::AP + i] = i < k ? t.action[i] : (int16_t)0;
            D.rec_improved[r * G::AP + i] = i < k ? sm.s0[i] : 0.0;
        }
        if (lane == 0) { D.rec_move[r] = (int16_t)move; D.rec_color[r] = (uint8_t)color; D.rec_k[r] = (int16_t)k; }
    }
    BScal s;
    wb_load<N>(sm.root, s, pool_of<N>(D), g, lane);
    wb_put_stone<N>(sm.root, s, move, color, D.zob, D.hist_hash + (size_t)g * G::MAXREC, D.hist_pos + (size_t)g * G::MAXREC, lane);
    wb_store<N>(sm.root, s, pool_of<N>(D), g, lane);
    const int pass_count = move == PASS ? gs[GS_PASSCNT] + 1 : 0;                // worker.py:67-70
    const int nmoves = gs[GS_NMOVES] + 1;
    __syncwarp();
    if (lane == 0) { gs[GS_PASSCNT] = pass_count; gs[GS_NMOVES] = nmoves; gs[GS_COLOR] = opp(color); }
    if (pass_count == 2) {                                                       // :76-87
        uint8_t* tmp = reinterpret_cast<uint8_t*>(sm.scratch.color);
        const int sc = D.scoring == 1 ? wb_tromp_taylor<N>(sm.root, sm.scratch.chain, sm.scratch.ls, lane)
                                      : wb_count_score<N>(sm.root, tmp, lane);                   // worker.py:81
        const float score = (float)sc - komi;
        if (lane == 0) {
            gs[GS_SCORE] = __float_as_int(score);
            gs[GS_WINNER] = score > 0.1f ? BLACK : (score < -0.1f ? WHITE : OB);
            gs[GS_RESIGNED] = 0; gs[GS_FINISHED] = 1;
        }
    } else if (nmoves >= max_moves && lane == 0) {                               // :56 loop bound: winner stays EMPTY
        gs[GS_WINNER] = EMPTY; gs[GS_RESIGNED] = 0; gs[GS_FINISHED] = 1; gs[GS_SCORE] = __float_as_int(0.0f);
    }
}

// ---------------------------------------------------------------------------------------------
// Board-only kernels behind the C ABI (parity tests and the position setup of genmove).

// Reset the games flagged in `mask` (nullptr: all) to an empty board and a fresh game id.
template <int N>
__global__ void __launch_bounds__(SEARCH_WARPS * 32) k_reset(Dev D, const uint8_t* mask, const u64* new_ids, const uint8_t* never_resign)
{
    using G = Geo<N>;
    const int g = blockIdx.x * SEARCH_WARPS + (threadIdx.x >> 5), lane = lane_id();
    if (g >= D.games) return;
    if (mask && !mask[g]) return;
    WarpSmem<N>& sm = warp_smem<N>();
    BScal s;
    wb_clear<N>(sm.root, s, lane);
    wb_store<N>(sm.root, s, pool_of<N>(D), g, lane);
    int* gs = D.gs + (size_t)g * GS_STRIDE;
    for (int i = lane; i < GS_STRIDE; i += 32) gs[i] = 0;
    __syncwarp();
    if (lane == 0) {
        gs[GS_COLOR] = BLACK; gs[GS_ACTIVE] = 1;
        gs[GS_NEVER_RESIGN] = never_resign ? never_resign[g] : 0;
        if (new_ids) D.game_id[g] = new_ids[g];
        D.hist_hash[(size_t)g * G::MAXREC] = 0; D.hist_pos[(size_t)g * G::MAXREC] = 0;
    }
}

// Play moves[g][0..count[g]) on each root board, colours alternating from the stored colour to move unless
// `colors` is given.  Optionally dumps the per-ply state for the parity tests.
struct PlyDump {
    uint8_t* color;      // [games][plies][CELLS]
    int16_t* libs;       // [games][plies][CELLS]   liberties of the string on the point (0 when empty)
    int16_t* size;       // [games][plies][CELLS]
    int* scal;           // [games][plies][5]       moves, ko_pos, ko_move, prisoner[0], prisoner[1]
    u64* hash;           // [games][plies]
    uint8_t* legal;      // [games][plies][2][NN]
    int16_t* satari;     // [games][plies][2][NN]
    uint8_t* eye;        // [games][plies][2][NN]
    uint8_t* cand;       // [games][plies][2][NN]
    int* score;          // [games][plies]
    int* tt_score;       // [games][plies]   Tromp-Taylor area score (optional)
    int stride;          // plies allocated per game
};

template <int N>
__global__ void __launch_bounds__(SEARCH_WARPS * 32) k_play(Dev D, const int16_t* moves, const uint8_t* colors, const int* count,
                                                            int stride, PlyDump dump, int do_dump)
{
    using G = Geo<N>;
    const int g = blockIdx.x * SEARCH_WARPS + (threadIdx.x >> 5), lane = lane_id();
    if (g >= D.games) return;
    const int n = count[g];
    if (n == 0) return;
    WarpSmem<N>& sm = warp_smem<N>();
    int* gs = D.gs + (size_t)g * GS_STRIDE;
    BScal s;
    wb_load<N>(sm.root, s, pool_of<N>(D), g, lane);
    u64* hh = D.hist_hash + (size_t)g * G::MAXREC;
    int16_t* hp = D.hist_pos + (size_t)g * G::MAXREC;
    int color = gs[GS_COLOR];
    for (int i = 0; i < n; i++) {
        const int mv = moves[(size_t)g * stride + i];
        if (colors) color = colors[(size_t)g * stride + i];
        wb_put_stone<N>(sm.root, s, mv, color, D.zob, hh, hp, lane);
        color = opp(color);
        if (do_dump) {
            const size_t pi = (size_t)g * dump.stride + i;
            for (int c = lane; c < G::CELLS; c += 32) {
                const int col = sm.root.color[c];
                const bool stone = (col == BLACK || col == WHITE);
                const unsigned ls = stone ? sm.root.ls[sm.root.chain[c]] : 0u;
                dump.color[pi * G::CELLS + c] = (uint8_t)col;
                dump.libs[pi * G::CELLS + c] = (int16_t)(ls >> 16);
                dump.size[pi * G::CELLS + c] = (int16_t)(ls & 0xffffu);
            }
            if (lane == 0) {
                int* sc = dump.scal + pi * 5;
                sc[0] = s.moves; sc[1] = s.ko_pos; sc[2] = s.ko_move; sc[3] = s.pris0; sc[4] = s.pris1;
                dump.hash[pi] = s.hash;
            }
            for (int ci = 0; ci < 2; ci++) {
                const int col = ci == 0 ? BLACK : WHITE;
                const size_t ob = (pi * 2 + ci) * G::NN;
                wb_analyze<N>(sm.root, sm.an, s, col, D.superko != 0, D.zob, D.eye, hh, lane,
                    [&](int, int idx, int, bool legal, int satari, bool eye) {
                        if (idx < G::NN) {
                            dump.legal[ob + idx] = legal ? 1 : 0;
                            dump.satari[ob + idx] = legal ? (int16_t)satari : (int16_t)0;
                            dump.eye[ob + idx] = (legal && eye) ? 1 : 0;
                            dump.cand[ob + idx] = (legal && satari < 7 && !eye) ? 1 : 0;
                        }
                    });
            }
            uint8_t* tmp = reinterpret_cast<uint8_t*>(sm.scratch.color);
            const int score = wb_count_score<N>(sm.root, tmp, lane);
            if (lane == 0) dump.score[pi] = score;
            if (dump.tt_score) {
                const int tts = wb_tromp_taylor<N>(sm.root, sm.scratch.chain, sm.scratch.ls, lane);
                if (lane == 0) dump.tt_score[pi] = tts;
            }
        }
    }
    wb_store<N>(sm.root, s, pool_of<N>(D), g, lane);
    if (lane == 0) gs[GS_COLOR] = color;
}

// Training samples of finished games straight from the device record ring: what
// generate_reinforcement_learning_data (nn/data_generator.py:89-149) produces by re-reading the SGF -- the position before
// each sampled ply as input planes under one of the 8 symmetries (nn/feature.py:10-57, go_board.py:74-104 sym_map), the
// improved policy of that move's search as the target under the same symmetry with 1e-18 elsewhere (feature.py:80-102), and
// the game result seen from the side to move (data_generator.py:121-137) -- without SGF text, parsing or a host replay.
// One warp replays one game on a shared-memory board; which plies / symmetries are used is the caller's draw (the
// reference takes them from numpy's global stream), passed as plies[8] (ascending, -1 padded) and syms[8].
template <int N> __device__ __forceinline__ int sym_source(int idx, int sym)
{   // raster index of the point that supplies output point idx (go_board.py:86-103)
    const int x = idx % N, y = idx / N, m = N - 1;
    int sx, sy;
    switch (sym) {
        case 1: sx = m - x; sy = y; break;
        case 2: sx = x; sy = m - y; break;
        case 3: sx = m - x; sy = m - y; break;
        case 4: sx = y; sy = x; break;
        case 5: sx = y; sy = m - x; break;
        case 6: sx = m - y; sy = x; break;
        case 7: sx = m - y; sy = m - x; break;
        default: sx = x; sy = y; break;
    }
    return sy * N + sx;
}

template <int N>
__global__ void __launch_bounds__(SEARCH_WARPS * 32) k_emit_samples(Dev D, const int* games, int n, const int* plies, const int* syms,
                                                                    const int* out_base)
{
    using G = Geo<N>;
    const int j = blockIdx.x * SEARCH_WARPS + (threadIdx.x >> 5), lane = lane_id();
    if (j >= n) return;
    const int g = games[j];
    WarpSmem<N>& sm = warp_smem<N>();
    const int* gs = D.gs + (size_t)g * GS_STRIDE;
    const int n_moves = min(gs[GS_NMOVES], D.rec_moves);
    const int winner = gs[GS_WINNER];
    int vlabel = winner == BLACK ? 2 : (winner == WHITE ? 0 : 1);             // sgf/reader.py:345-358 get_value_label
    BScal s;
    wb_clear<N>(sm.root, s, lane);
    // the slot's history rows are free scratch: the game is over and tg_reset rewrites them for the next one
    u64* hh = D.hist_hash + (size_t)g * G::MAXREC;
    int16_t* hp = D.hist_pos + (size_t)g * G::MAXREC;
    if (lane == 0) { hh[0] = 0; hp[0] = 0; }
    __syncwarp();
    const size_t r0 = (size_t)g * D.rec_moves;
    int color = BLACK, next = 0, out = out_base[j];
    for (int i = 0; i < n_moves && next < 8; i++) {
        if (plies[j * 8 + next] == i) {
            const int sym = syms[j * 8 + next];
            float* pl = D.smp_input + (size_t)out * G::PLANES;
            const int prev = hp[s.moves - 1];                                 // feature.py:34 record.get(moves - 1)
            const bool prev_pass = s.moves > 1 && prev == PASS;               // :39
            for (int idx = lane; idx < G::NN; idx += 32) {
                const int src = sym_source<N>(idx, sym);
                const int spos = onboard_pos<N>(src);
                int d = sm.root.color[spos];
                if (color == WHITE && d != 0) d = 3 - d;                      // :24-25
                pl[idx] = d == 0 ? 1.0f : 0.0f; pl[G::NN + idx] = d == 1 ? 1.0f : 0.0f; pl[2 * G::NN + idx] = d == 2 ? 1.0f : 0.0f;
                pl[3 * G::NN + idx] = (!prev_pass && prev == spos) ? 1.0f : 0.0f;   // :43-46
                pl[4 * G::NN + idx] = prev_pass ? 1.0f : 0.0f;
                pl[5 * G::NN + idx] = color == WHITE ? -1.0f : 1.0f;          // :50-52
            }
            // policy target: feature.py:91-100
            const size_t rr = (r0 + i) * G::AP;
            const int k = D.rec_k[r0 + i];
            double p_pass = 1e-18;
            for (int idx = lane; idx < G::NN; idx += 32) sm.s0[idx] = 1e-18;
            __syncwarp();
            for (int c = lane; c < k; c += 32) {
                const int a = D.rec_action[rr + c];
                if (a != PASS) sm.s0[(a / G::W - 1) * N + (a % G::W - 1)] = D.rec_improved[rr + c];
            }
            __syncwarp();
            if (k > 0 && D.rec_action[rr + k - 1] == PASS) p_pass = D.rec_improved[rr + k - 1];     // PASS is the last child (tree.py:264)
            double* po = D.smp_policy + (size_t)out * G::A;
            for (int idx = lane; idx < G::NN; idx += 32) po[idx] = sm.s0[sym_source<N>(idx, sym)];
            if (lane == 0) { po[G::NN] = p_pass; D.smp_value[out] = vlabel; }
            __syncwarp();
            out++; next++;
        }
        wb_put_stone<N>(sm.root, s, D.rec_move[r0 + i], color, D.zob, hh, hp, lane);   // data_generator.py:131
        color = opp(color);
        vlabel = 2 - vlabel;                                                  // :133
    }
}

// Snapshot the root boards as leaf 0 of every game (used by the stand-alone feature-plane entry point).
template <int N>
__global__ void __launch_bounds__(SEARCH_WARPS * 32) k_snapshot_roots(Dev D)
{
    const int g = blockIdx.x * SEARCH_WARPS + (threadIdx.x >> 5), lane = lane_id();
    if (g >= D.games) return;
    int* gs = D.gs + (size_t)g * GS_STRIDE;
    WarpSmem<N>& sm = warp_smem<N>();
    BScal s;
    wb_load<N>(sm.root, s, pool_of<N>(D), g, lane);
    if (lane == 0) { gs[GS_NLEAF] = 0; gs[GS_NUNIQ] = 0; }
    __syncwarp();
    push_leaf<N>(D, g, gs, sm.root, s, gs[GS_COLOR], nullptr, 0, -1, lane);
}

// Test evaluators, one warp per slot: variant 0 = the hash "network" of oracle.hashnet (dyadic fp32 outputs: every sum of
// them is exact in any order), variant 1 = oracle.hashnet2 (values k/1000, logits raw/1000 - 4: fp32 sums round, so the
// queue-order fp32 accumulation of the backup is observable).
template <int N>
__global__ void __launch_bounds__(128) k_hashnet(const float* planes, const int* n_slots, int n_direct, int use_logit, float* policy, float* value, int variant)
{
    using G = Geo<N>;
    const int slot = blockIdx.x * 4 + (threadIdx.x >> 5), lane = lane_id();
    if (slot >= (n_slots ? *n_slots : n_direct)) return;
    const float* pl = planes + (size_t)slot * G::PLANES;
    u64 h = 0;
    for (int j = lane; j < G::PLANES; j += 32) {
        const u64 code = (u64)(long long)(pl[j] + 1.0f);
        h += mix64(3ull * (u64)j + code);
    }
    {   // 64-bit wrap-around sum across the warp
        unsigned lo = (unsigned)h, hi = (unsigned)(h >> 32);
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            const unsigned olo = __shfl_xor_sync(0xffffffffu, lo, o), ohi = __shfl_xor_sync(0xffffffffu, hi, o);
            const u64 a = ((u64)hi << 32) | lo, b = ((u64)ohi << 32) | olo, c = a + b;
            lo = (unsigned)c; hi = (unsigned)(c >> 32);
        }
        h = ((u64)hi << 32) | lo;
    }
    for (int i = lane; i < G::A; i += 32) {
        const u64 r = mix64(h + (u64)i);
        const float raw = (float)((r >> 40) & 0xFFFFull);
        if (variant == 0) policy[(size_t)slot * G::A + i] = use_logit ? __fsub_rn(__fdiv_rn(raw, 8192.0f), 4.0f) : __fdiv_rn(raw, 1048576.0f);
        else policy[(size_t)slot * G::A + i] = use_logit ? __fsub_rn(__fdiv_rn(raw, 1000.0f), 4.0f) : __fdiv_rn(raw, 1000000.0f);
    }
    if (lane == 0) {
        const u64 ra = mix64(h + 1000ull), rb = mix64(h + 1001ull);
        const float va = variant == 0 ? (float)(ra & 0xFFull) : (float)(ra % 500ull);
        const float vb = variant == 0 ? (float)(rb & 0xFFull) : (float)(rb % 500ull);
        const float den = variant == 0 ? 512.0f : 1000.0f;
        const float v0 = __fdiv_rn(va, den), v1 = __fdiv_rn(vb, den);
        value[(size_t)slot * 3 + 0] = v0; value[(size_t)slot * 3 + 1] = v1;
        value[(size_t)slot * 3 + 2] = __fsub_rn(__fsub_rn(1.0f, v0), v1);
    }
}

}  // namespace tg
