// tg_engine.cu -- host side of the engine and the C ABI (include/tamago_b200.h).
//
// Owns the structure-of-arrays pools in HBM (root boards, node pool, leaf queues, evaluator batch), folds and
// packs the DualNet parameters, and queues the kernel sequence of one move of every game on one CUDA stream.
#include "../../include/tamago_b200.h"
#include "tg_search.cuh"
#include "tg_block.cuh"
#include "tg_dualnet.cuh"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <atomic>
#include <thread>

using namespace tg;

// host-side fan-out for record formatting / file output (a few worker threads; the GPU runs the next step meanwhile)
template <class F> static void parallel_for(int n, F&& f)
{
    if (n <= 0) return;
    const int hw = (int)std::thread::hardware_concurrency();
    const int nt = std::max(1, std::min(std::min(n, 16), hw > 1 ? hw - 1 : 1));
    if (nt == 1) { for (int i = 0; i < n; i++) f(i); return; }
    std::atomic<int> next{0};
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back([&] { for (int i; (i = next.fetch_add(1)) < n;) f(i); });
    for (auto& t : th) t.join();
}

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return fail(TG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

struct tg_engine {
    tg_config cfg{};
    int N = 0, NN = 0, A = 0, AP = 0, CELLS = 0, CP = 0, MAXREC = 0, SNAP = 0, PLANES = 0;
    int cap_sh = 0, cap_puct = 0, depth_sh = 64, depth_puct = 0, cap_max = 0;
    size_t path_words = 0;
    int slot_cap = 0;
    cudaStream_t stream = nullptr;
    std::vector<cudaEvent_t> events;
    Dev D{};
    NetDev net{};
    std::vector<void*> allocs;
    std::vector<void*> net_allocs;
    // fp32 path scratch
    float* act[3] = {nullptr, nullptr, nullptr};
    int simt_chunk = 0;
    // host staging (pinned)
    int* h_gs = nullptr; int16_t* h_action = nullptr; double* h_improved = nullptr; int* h_visits = nullptr;
    // tg_reset staging: pinned host block + device block, sized once (nothing is allocated on the hot path)
    unsigned char* h_reset = nullptr; unsigned char* d_reset = nullptr; cudaEvent_t ev_reset = nullptr; bool reset_pending = false;
    // asynchronous step state (tg_genmove_async -> tg_collect)
    bool step_pending = false, step_arrays = false; cudaEvent_t ev_step = nullptr;
    // finished-game records fetched from the device ring (tg_fetch_records -> tg_format_records / tg_write_records)
    struct Fetched { int game, n_moves, winner, resigned; float score; size_t off; };
    std::vector<Fetched> fetched; unsigned char* h_rec = nullptr; size_t h_rec_cap = 0; cudaEvent_t ev_rec = nullptr; bool rec_pending = false;
    int rec_row_bytes = 0;
    cudaEvent_t ev_x = nullptr;                  // cross-stream ordering with a caller's stream
    int64_t sample_cap = 0, sample_count = 0;    // training samples emitted from the record ring
    int* d_emit = nullptr; size_t d_emit_cap = 0;
    bool have_weights = false, have_zobrist = false;
    int64_t launches = 0;
    int n_eval_events = 2;
    float last_ms = 0.f, last_eval_ms = 0.f;
    int64_t last_eval_slots = 0;
    int sms = 148;
    bool puct_warp = false, unfused_planes = false;   // TG_UNFUSED_PLANES=1: k_planes + fp32 planes for the tensor-core evaluator too (A/B)
    int puct_nt = 256;                           // threads per game of the block-per-game PUCT kernels (128, 256 or 512)
    bool puct_defer = false;                     // block-per-game batches: selections in one kernel, board work of all leaves in another
    int walk_slots = 2;                          // node-row cache slots of k_walk_puct_blk
    int snap_plan = 0;                           // board-snapshot levels per game k_descend_puct_snap will get (0: none)
    bool puct_wave = false;                      // deferred mode: the tree walk runs as a wavefront (k_wave_puct_blk)
    int wave_gt = 128;                           // threads per descent of the wavefront walk (512-thread CTAs: 128 measured faster than 64)
    const uint32_t* eye2 = nullptr;              // eye table packed to two bits per code (block-per-game kernels keep it in shared memory)
};

template <class T> static int dalloc(tg_engine* e, T** p, size_t n, bool zero = true)
{
    void* q = nullptr;
    const size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
    CK(cudaMalloc(&q, bytes));
    if (zero) CK(cudaMemsetAsync(q, 0, bytes, e->stream));
    e->allocs.push_back(q);
    *p = reinterpret_cast<T*>(q);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// eye table (board/pattern.py:53-98): 3x3 neighbourhood codes whose centre is an eye of a colour.
// Built geometrically: the 8 neighbours carry (dy, dx); symmetries permute coordinates.
// ---------------------------------------------------------------------------------------------
static const int kDy[8] = {-1, -1, -1, 0, 0, 1, 1, 1}, kDx[8] = {-1, 0, 1, -1, 1, -1, 0, 1};
static int cell_of(int dy, int dx) { for (int i = 0; i < 8; i++) if (kDy[i] == dy && kDx[i] == dx) return i; return -1; }
template <class F> static unsigned pat_map(unsigned p, F f)
{   // new neighbour (dy,dx) takes the value of the old neighbour f(dy,dx)
    unsigned r = 0;
    for (int i = 0; i < 8; i++) { int sy, sx; f(kDy[i], kDx[i], sy, sx); r |= ((p >> (2 * cell_of(sy, sx))) & 3u) << (2 * i); }
    return r;
}
static unsigned pat_swap(unsigned p)
{   // exchange black and white stones (pattern.py:227-236)
    unsigned r = 0;
    for (int i = 0; i < 8; i++) { unsigned v = (p >> (2 * i)) & 3u; v = ((v >> 1) & 1u) | ((v & 1u) << 1); r |= v << (2 * i); }
    return r;
}
static void build_eye_table(std::vector<uint8_t>& eye)
{
    static const unsigned seeds[20] = {
        0x5554, 0x5556, 0x5544, 0x5546, 0x1554, 0x1556, 0x1544, 0x1546, 0x1564, 0x1146,
        0xFD54, 0xFD55, 0xFF74, 0xFF75, 0x5566, 0xFD66, 0x5965, 0x9955, 0xFD56, 0xFF76 };
    eye.assign(65536, EMPTY);
    auto put = [&](unsigned p) { eye[p & 0xffff] = BLACK; eye[pat_swap(p) & 0xffff] = WHITE; };
    put(0x5555); put(0x1144);
    auto flip_ud = [](int dy, int dx, int& sy, int& sx) { sy = -dy; sx = dx; };      // pat3_vertical_mirror
    auto flip_lr = [](int dy, int dx, int& sy, int& sx) { sy = dy; sx = -dx; };      // pat3_horizontal_mirror
    auto rot = [](int dy, int dx, int& sy, int& sx) { sy = dx; sx = -dy; };          // pat3_rotate_90
    for (unsigned s : seeds) {
        unsigned sym[8];
        sym[0] = s; sym[1] = pat_map(s, flip_ud); sym[2] = pat_map(s, flip_lr); sym[3] = pat_map(sym[2], flip_ud);
        for (int i = 0; i < 4; i++) sym[4 + i] = pat_map(sym[i], rot);
        for (unsigned p : sym) put(p);
    }
}

// ---------------------------------------------------------------------------------------------
constexpr int PUCT_SQRT_MAX = 1024;             // entries of the warp PUCT kernels' square-root table (larger arguments are computed)
static size_t search_smem(int N)
{
    switch (N) { case 9: return sizeof(WarpSmem<9>) * SEARCH_WARPS; case 13: return sizeof(WarpSmem<13>) * SEARCH_WARPS;
                 default: return sizeof(WarpSmem<19>) * SEARCH_WARPS; }
}

#define DISPATCH_N(e, ...) do { switch ((e)->N) { \
    case 9:  { constexpr int BN = 9;  __VA_ARGS__; } break; \
    case 13: { constexpr int BN = 13; __VA_ARGS__; } break; \
    case 19: { constexpr int BN = 19; __VA_ARGS__; } break; } } while (0)

template <int BN> static int setup_kernel_attrs()
{
    const int bytes = (int)(sizeof(WarpSmem<BN>) * SEARCH_WARPS);
    CK(cudaFuncSetAttribute(k_root_begin<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    CK(cudaFuncSetAttribute(k_descend_sh<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    CK(cudaFuncSetAttribute(k_descend_puct<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes + 4 * Geo<BN>::CP * 8 + PUCT_SQRT_MAX * 8));
    CK(cudaFuncSetAttribute(k_descend_puct_snap<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes + 4 * Geo<BN>::CP * 8 + PUCT_SQRT_MAX * 8));
    CK(cudaFuncSetAttribute(k_move_end<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    CK(cudaFuncSetAttribute(k_reset<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    CK(cudaFuncSetAttribute(k_play<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    CK(cudaFuncSetAttribute(k_snapshot_roots<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    CK(cudaFuncSetAttribute(k_emit_samples<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    CK(cudaFuncSetAttribute(k_conv3x3_simt<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * (BN + 2) * (BN + 2) * 4));
    // the search kernels keep whole boards in shared memory: prefer the largest carve-out so that more games are resident
    CK(cudaFuncSetAttribute(k_root_begin<BN>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CK(cudaFuncSetAttribute(k_descend_sh<BN>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CK(cudaFuncSetAttribute(k_descend_puct<BN>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CK(cudaFuncSetAttribute(k_descend_puct_snap<BN>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    return 0;
}
constexpr int WAVE_SQRT_MAX = 4096;             // entries of the wavefront walk's square-root table (larger arguments are computed)
template <int BN, int NT, int GT> static size_t wave_smem(int nsqrt = WAVE_SQRT_MAX)
{
    return ((sizeof(WaveSmem<BN, NT, GT>) + 15) & ~(size_t)15) + (size_t)(NT / GT - 1) * sizeof(typename BlkSmem<BN, NT>::NodeStage) + (size_t)nsqrt * 8;
}
template <int BN, int NT, int GT> static int setup_wave_attr()
{
    static_assert(sizeof(WaveSmem<BN, NT, GT>) + (NT / GT) * sizeof(typename BlkSmem<BN, NT>::NodeStage) + WAVE_SQRT_MAX * 8 <= 227 * 1024, "wavefront walk scratch");
    CK(cudaFuncSetAttribute(k_wave_puct_blk<BN, NT, GT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wave_smem<BN, NT, GT>()));
    return 0;
}
template <int BN> static int setup_blk_attr()
{
    static_assert(sizeof(BlkSmem<BN, 512>) <= 96 * 1024, "block-per-game scratch");
    CK(cudaFuncSetAttribute(k_descend_puct_blk<BN, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BlkSmem<BN, 128>)));
    CK(cudaFuncSetAttribute(k_descend_puct_blk<BN, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BlkSmem<BN, 256>)));
    CK(cudaFuncSetAttribute(k_descend_puct_blk<BN, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BlkSmem<BN, 512>)));
    CK(cudaFuncSetAttribute(k_walk_puct_blk<BN, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(k_walk_puct_blk<BN, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(k_walk_puct_blk<BN, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    if ( setup_wave_attr<BN, 512, 64>() || setup_wave_attr<BN, 512, 128>() || setup_wave_attr<BN, 256, 64>() || setup_wave_attr<BN, 128, 32>()) return -1;
    CK(cudaFuncSetAttribute(k_expand_leaves_blk<BN, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ExpandSmem<BN, 128>)));
    CK(cudaFuncSetAttribute(k_expand_leaves_blk<BN, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ExpandSmem<BN, 256>)));
    CK(cudaFuncSetAttribute(k_expand_leaves_blk<BN, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ExpandSmem<BN, 512>)));
    return 0;
}
// shared memory of the tree-walk kernel with `slots` cached node rows
template <int BN, int NT> static size_t walk_smem(int slots)
{
    return ((sizeof(WalkSmem<BN, NT>) + 15) & ~(size_t)15) + (size_t)(slots - 2) * sizeof(typename BlkSmem<BN, NT>::NodeStage);
}
template <int BN, int G> static int setup_tc_attr()
{
    CK(cudaFuncSetAttribute(k_dualnet_tc<BN, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, NetGeo<BN, G>::SMEM_BYTES));
    return 0;
}
template <int BN> struct TcGroup { static constexpr int G = 1; };
template <> struct TcGroup<9> { static constexpr int G = 5; };
template <> struct TcGroup<13> { static constexpr int G = 2; };

static int search_grid(const tg_engine* e) { return (e->cfg.games + SEARCH_WARPS - 1) / SEARCH_WARPS; }

// mcts/sequential_halving.py:36-60 on the host: the longest schedule over m = min(#children, 16) sizes the phase loop
static int sh_phases_host(int m, int visits, int* max_leaves)
{
    std::vector<int> cons, cnts;
    if (m <= 1) { cons.push_back(1); cnts.push_back(visits); }
    else {
        const int log2max = (int)std::ceil(std::log2((double)m));
        int total = 0, nc = m;
        while (total < visits) {
            int extra = std::max(1, visits / (log2max * nc));
            for (int x = 0; x < extra && total < visits; x++) {
                const int c = std::min(nc, visits - total);
                total += c;
                size_t f = 0;
                for (; f < cons.size(); f++) if (cons[f] == c) break;
                if (f < cons.size()) cnts[f]++; else { cons.push_back(c); cnts.push_back(1); }
            }
            nc = std::max(2, nc / 2);
        }
    }
    int ml = 0;
    for (size_t i = 0; i < cons.size(); i++) ml = std::max(ml, cons[i] * cnts[i]);
    if (max_leaves) *max_leaves = ml;
    return (int)cons.size();
}

// ---------------------------------------------------------------------------------------------
extern "C" const char* tg_last_error(void) { return g_err.c_str(); }
extern "C" int tg_action_stride(int n) { return (n * n + 1 + 31) & ~31; }

extern "C" int tg_engine_create(const tg_config* cfg, tg_engine** out)
{
    if (!cfg || !out) return fail(TG_ERR_ARG, "null argument");
    if (cfg->board_size != 9 && cfg->board_size != 13 && cfg->board_size != 19)
        return fail(TG_ERR_ARG, "board_size must be 9, 13 or 19");
    if (cfg->games < 1 || cfg->max_visits < 1) return fail(TG_ERR_ARG, "games and max_visits must be positive");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(TG_ERR_CUDA, "no CUDA device: tamago_b200 has no CPU path");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(TG_ERR_ARG, "bad device ordinal");
    CK(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major < 10) return fail(TG_ERR_CUDA, "tamago_b200 kernels are built for sm_100a only");

    tg_engine* e = new tg_engine();
    e->cfg = *cfg;
    if (e->cfg.batch_size < 1) e->cfg.batch_size = 1;
    if (e->cfg.net_blocks <= 0) e->cfg.net_blocks = 6;
    if (e->cfg.net_blocks > 15) { delete e; return fail(TG_ERR_ARG, "net_blocks must be <= 15"); }
    const int N = cfg->board_size;
    e->N = N; e->NN = N * N; e->A = e->NN + 1; e->AP = tg_action_stride(N);
    e->CELLS = (N + 2) * (N + 2); e->CP = (e->CELLS + 3) & ~3; e->MAXREC = 3 * e->NN; e->PLANES = 6 * e->NN;
    e->SNAP = (16 + e->NN + 15) & ~15;
    e->sms = prop.multiProcessorCount;
    const int games = cfg->games, V = cfg->max_visits;
    const int max_nodes = cfg->max_nodes > 0 ? cfg->max_nodes : V + 2;
    if (max_nodes >= (1 << (32 - PATH_NODE_SHIFT))) { delete e; return fail(TG_ERR_ARG, "max_nodes too large"); }

    // leaf queue geometry: sequential halving enqueues a whole phase (<= V leaves, depth <= 64);
    // PUCT enqueues batch_size leaves whose paths may be as long as the move history allows
    int sh_leaves = 1;
    for (int m = 1; m <= 16; m++) { int ml; sh_phases_host(m, V, &ml); sh_leaves = std::max(sh_leaves, ml); }
    e->cap_sh = sh_leaves; e->depth_sh = 64;
    e->cap_puct = e->cfg.batch_size; e->depth_puct = e->MAXREC + 2;
    e->cap_max = std::max(e->cap_sh, e->cap_puct);
    e->path_words = std::max((size_t)e->cap_sh * e->depth_sh, (size_t)e->cap_puct * e->depth_puct);
    const size_t want_slots = (size_t)games * e->cap_max;
    e->slot_cap = (int)std::min<size_t>(want_slots, (size_t)1 << 22);

    auto bail = [&](int rc) { tg_engine_destroy(e); return rc; };
    if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(fail(TG_ERR_CUDA, "stream"));
    e->events.resize(64);
    for (auto& ev : e->events) if (cudaEventCreate(&ev) != cudaSuccess) return bail(fail(TG_ERR_CUDA, "event"));

    Dev& D = e->D;
    D.scoring = cfg->scoring;
    D.games = games; D.superko = cfg->superko; D.cgos = cfg->cgos_mode; D.dedup = cfg->dedup; D.seed = cfg->seed;
    D.cap = e->cap_sh; D.max_depth = e->depth_sh; D.slot_cap = e->slot_cap;
    D.tree.max_nodes = max_nodes;
    int rc = 0;
    const size_t gn = (size_t)games * max_nodes;
#define DA(p, n) if ((rc = dalloc(e, &(p), (n))) != 0) return bail(rc)
    DA(D.b_color, (size_t)games * e->CP); DA(D.b_chain, (size_t)games * e->CP); DA(D.b_bloom, (size_t)games * BLOOM_WORDS);
    DA(D.b_hash, games); DA(D.b_scal, (size_t)games * 8);
    DA(D.hist_hash, (size_t)games * e->MAXREC); DA(D.hist_pos, (size_t)games * e->MAXREC);
    DA(D.tree.hdr, gn * H_STRIDE); DA(D.tree.action, gn * e->AP); DA(D.tree.cidx, gn * e->AP); DA(D.tree.cval, gn * e->AP);
    DA(D.tree.cvis, gn * e->AP); DA(D.tree.cpol, gn * e->AP); DA(D.tree.cvl, gn * e->AP); DA(D.tree.cvsum, gn * e->AP);
    DA(D.tree.noise, (size_t)games * e->AP);
    DA(D.gs, (size_t)games * GS_STRIDE); DA(D.game_id, games);
    DA(D.path, (size_t)games * e->path_words); DA(D.path_len, (size_t)games * e->cap_max);
    DA(D.leaf_node, (size_t)games * e->cap_max); DA(D.leaf_slot, (size_t)games * e->cap_max);
    DA(D.leaf_flag, (size_t)games * e->cap_max);
    DA(D.snap, (size_t)games * e->cap_max * e->SNAP);
    DA(D.planes, (size_t)e->slot_cap * e->PLANES); DA(D.policy, (size_t)e->slot_cap * e->A); DA(D.value, (size_t)e->slot_cap * 3);
    DA(D.n_slots, 1); DA(D.slot_src, (size_t)e->slot_cap);
    DA(D.out_action, (size_t)games * e->AP); DA(D.out_improved, (size_t)games * e->AP); DA(D.out_visits, (size_t)games * e->AP);
    if (cfg->record_ring) {                       // per-game record of the running game (SelfPlayRecord), rows = move limit
        const size_t rows = (size_t)games * 2 * e->NN;
        D.rec_moves = 2 * e->NN;
        DA(D.rec_move, rows); DA(D.rec_color, rows); DA(D.rec_k, rows); DA(D.rec_action, rows * e->AP); DA(D.rec_improved, rows * e->AP);
    }
    if (cfg->sample_cap > 0) {
        if (!cfg->record_ring) return bail(fail(TG_ERR_ARG, "sample_cap needs record_ring"));
        e->sample_cap = cfg->sample_cap;
        DA(D.smp_input, (size_t)cfg->sample_cap * e->PLANES); DA(D.smp_policy, (size_t)cfg->sample_cap * e->A); DA(D.smp_value, (size_t)cfg->sample_cap);
    }
    {
        void* q = nullptr;
        if (cudaMalloc(&q, (size_t)games * 10) != cudaSuccess) return bail(fail(TG_ERR_CUDA, "reset staging"));
        e->allocs.push_back(q); e->d_reset = reinterpret_cast<unsigned char*>(q);
    }
    {
        u64* z; uint8_t* eye;
        DA(z, (size_t)4 * e->CELLS); DA(eye, 65536);
        std::vector<uint8_t> tab; build_eye_table(tab);
        if (cudaMemcpyAsync(eye, tab.data(), 65536, cudaMemcpyHostToDevice, e->stream) != cudaSuccess) return bail(fail(TG_ERR_CUDA, "eye table upload"));
        if (cudaStreamSynchronize(e->stream) != cudaSuccess) return bail(fail(TG_ERR_CUDA, "sync"));
        D.zob = z; D.eye = eye;
        D.prof = nullptr;
        if (getenv("TG_PROF")) {
            void* q = nullptr;
            if (cudaMalloc(&q, 64 * sizeof(long long)) == cudaSuccess) { cudaMemset(q, 0, 64 * sizeof(long long)); e->allocs.push_back(q); D.prof = reinterpret_cast<long long*>(q); }
        }
    }
    if (cfg->evaluator == TG_EVAL_DUALNET_FP32) {
        e->simt_chunk = std::min(e->slot_cap, 8192);
        for (int i = 0; i < 3; i++) DA(e->act[i], (size_t)e->simt_chunk * 64 * e->NN);
    }
#undef DA
    if (cudaMallocHost(&e->h_gs, (size_t)games * GS_STRIDE * sizeof(int)) != cudaSuccess ||
        cudaMallocHost(&e->h_action, (size_t)games * e->AP * sizeof(int16_t)) != cudaSuccess ||
        cudaMallocHost(&e->h_improved, (size_t)games * e->AP * sizeof(double)) != cudaSuccess ||
        cudaMallocHost(&e->h_visits, (size_t)games * e->AP * sizeof(int)) != cudaSuccess ||
        cudaMallocHost(&e->h_reset, (size_t)games * 10) != cudaSuccess)
        return bail(fail(TG_ERR_CUDA, "pinned host allocation failed"));
    if (cudaEventCreateWithFlags(&e->ev_reset, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&e->ev_step, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&e->ev_rec, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&e->ev_x, cudaEventDisableTiming) != cudaSuccess)
        return bail(fail(TG_ERR_CUDA, "event"));

    DISPATCH_N(e, rc = setup_kernel_attrs<BN>());
    if (rc) return bail(rc);
    DISPATCH_N(e, (rc = setup_tc_attr<BN, TcGroup<BN>::G>()));
    if (rc) return bail(rc);
    DISPATCH_N(e, rc = setup_blk_attr<BN>());
    if (rc) return bail(rc);
    // PUCT kernels: one warp per game when the pool fills the machine with warps (throughput), one CTA per game when it
    // does not (latency): BASELINE configs[3] (1024 games) runs warp-per-game, configs[4] (one game) block-per-game
    e->puct_warp = getenv("TG_PUCT_WARP") != nullptr || (games > 3 * e->sms && getenv("TG_PUCT_BLOCK") == nullptr);
    e->unfused_planes = getenv("TG_UNFUSED_PLANES") != nullptr;
    // board snapshots along the previous path for the one-descent-per-launch PUCT kernel (warp-per-game pools with batch 1:
    // BASELINE configs[3]); up to 32 levels (256 plies) per game within 512 MB, none otherwise
    // (allocated by the first PUCT move: a pool that only ever runs sequential halving does not pay for them)
    D.snap_levels = 0; D.snap_words = e->CP * 7 / 4 + BLOOM_WORDS + 8; e->snap_plan = 0;
    if (e->puct_warp && e->cfg.batch_size == 1 && getenv("TG_PUCT_NOSNAP") == nullptr)
        e->snap_plan = (int)std::min<size_t>(32, ((size_t)512 << 20) / ((size_t)games * D.snap_words * 4));
    e->puct_nt = games <= e->sms ? 512 : 256;
    if (const char* nt = getenv("TG_PUCT_NT")) e->puct_nt = atoi(nt);
    // few games and real batches: the board work of a batch's leaves is spread over the idle SMs (tg_block.cuh, deferred
    // expansion); TG_PUCT_DEFER=0/1 forces the choice (A/B measurements, tests)
    e->puct_defer = games <= e->sms && e->cfg.batch_size >= 2;
    if (const char* df = getenv("TG_PUCT_DEFER")) e->puct_defer = atoi(df) != 0;
    if (e->cfg.batch_size > WALK_MAX_BATCH) e->puct_defer = false;
    // ... and the walk itself is pipelined over the descents of the batch (leaf deduplication needs the queue in order:
    // sequential walk); TG_PUCT_WAVE=0/1, TG_WAVE_GT=64/128 (512-thread CTAs) for A/B measurements
    e->puct_wave = e->puct_defer && !e->cfg.dedup;
    if (const char* wv = getenv("TG_PUCT_WAVE")) e->puct_wave = e->puct_defer && !e->cfg.dedup && atoi(wv) != 0;
    if (const char* gt = getenv("TG_WAVE_GT")) e->wave_gt = atoi(gt) == 64 ? 64 : 128;
    e->walk_slots = games <= e->sms ? 12 : 2;
    if (const char* ws = getenv("TG_WALK_SLOTS")) e->walk_slots = std::min(16, std::max(2, atoi(ws)));
    {
        std::vector<uint8_t> tab; build_eye_table(tab);
        std::vector<uint32_t> packed(4096, 0u);
        for (int i = 0; i < 65536; i++) packed[i >> 4] |= (uint32_t)(tab[i] & 3u) << ((i & 15) * 2);
        uint32_t* q = nullptr;
        if ((rc = dalloc(e, &q, 4096, false)) != 0) return bail(rc);
        if (cudaMemcpyAsync(q, packed.data(), 4096 * 4, cudaMemcpyHostToDevice, e->stream) != cudaSuccess || cudaStreamSynchronize(e->stream) != cudaSuccess)
            return bail(fail(TG_ERR_CUDA, "eye table upload"));
        e->eye2 = q;
    }

    // default Zobrist table (splitmix64 stream); tg_set_zobrist replaces it
    {
        std::vector<u64> z((size_t)4 * e->CELLS);
        for (size_t i = 0; i < z.size(); i++) z[i] = mix64(mix64(0x7A6Dull) + (u64)i);
        if (cudaMemcpyAsync(const_cast<u64*>(D.zob), z.data(), z.size() * 8, cudaMemcpyHostToDevice, e->stream) != cudaSuccess)
            return bail(fail(TG_ERR_CUDA, "zobrist upload"));
        if (cudaStreamSynchronize(e->stream) != cudaSuccess) return bail(fail(TG_ERR_CUDA, "sync"));
    }
    *out = e;
    rc = tg_reset(e, nullptr, nullptr, nullptr);
    if (rc) { *out = nullptr; return bail(rc); }
    return TG_OK;
}

// DualNet activations outside the fp16 operand range are clamped by the kernel AND reported: every entry point that returns
// evaluator results to the host checks the flag (the stream is idle at these points) and fails loudly.
static int check_net_overflow(tg_engine* e)
{
    if (!e->have_weights || !e->net.overflow || e->cfg.evaluator != TG_EVAL_DUALNET_TC) return 0;
    int flag = 0;
    CK(cudaMemcpy(&flag, e->net.overflow, 4, cudaMemcpyDeviceToHost));
    if (!flag) return 0;
    CK(cudaMemset(e->net.overflow, 0, 4));
    return fail(TG_ERR_SEARCH, "DualNet activation overflow: a feature map left the fp16 operand range (|x| > 6e4) and was clamped; "
                               "the tensor-core evaluator is specified for networks whose activations stay below that "
                               "(use TG_EVAL_DUALNET_FP32 for others)");
}

static void free_net(tg_engine* e) { for (void* p : e->net_allocs) cudaFree(p); e->net_allocs.clear(); e->have_weights = false; }

extern "C" void tg_engine_destroy(tg_engine* e)
{
    if (!e) return;
    cudaSetDevice(e->cfg.device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    free_net(e);
    for (void* p : e->allocs) cudaFree(p);
    for (auto ev : e->events) if (ev) cudaEventDestroy(ev);
    if (e->h_gs) cudaFreeHost(e->h_gs);
    if (e->h_action) cudaFreeHost(e->h_action);
    if (e->h_improved) cudaFreeHost(e->h_improved);
    if (e->h_visits) cudaFreeHost(e->h_visits);
    if (e->h_reset) cudaFreeHost(e->h_reset);
    if (e->h_rec) cudaFreeHost(e->h_rec);
    if (e->d_emit) cudaFree(e->d_emit);
    for (cudaEvent_t ev : {e->ev_reset, e->ev_step, e->ev_rec, e->ev_x}) if (ev) cudaEventDestroy(ev);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

extern "C" int tg_set_zobrist(tg_engine* e, const uint64_t* table)
{
    if (!e || !table) return fail(TG_ERR_ARG, "null argument");
    CK(cudaSetDevice(e->cfg.device));
    CK(cudaMemcpyAsync(const_cast<u64*>(e->D.zob), table, (size_t)4 * e->CELLS * 8, cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->have_zobrist = true;
    return TG_OK;
}

// ---------------------------------------------------------------------------------------------
// DualNet parameters: fold eval-mode BatchNorm into the convolutions, split the tensor-core operands into
// fp16 (hi, lo) pairs under a per-layer power-of-two scale, and lay them out as UMMA B tiles.
// ---------------------------------------------------------------------------------------------
template <class T> static int upload(tg_engine* e, const std::vector<T>& h, const T** dptr)
{
    void* q = nullptr;
    CK(cudaMalloc(&q, std::max<size_t>(h.size(), 1) * sizeof(T)));
    e->net_allocs.push_back(q);
    CK(cudaMemcpyAsync(q, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, e->stream));
    *dptr = reinterpret_cast<const T*>(q);
    return 0;
}

extern "C" int tg_load_weights(tg_engine* e, const tg_weights* w)
{
    if (!e || !w) return fail(TG_ERR_ARG, "null argument");
    CK(cudaSetDevice(e->cfg.device));
    CK(cudaStreamSynchronize(e->stream));
    free_net(e);
    const int blocks = e->cfg.net_blocks, L = 1 + 2 * blocks, NN = e->NN, A = e->A;
    // folded fp64 weights per layer: wf[l][oc][ic][tap], bias[l][oc]
    std::vector<std::vector<double>> wf(L);
    std::vector<float> bias((size_t)L * 64), scale(L);
    auto fold = [&](int l, const float* cw, int cin, const float* bn, float eps) {
        wf[l].assign((size_t)64 * cin * 9, 0.0);
        for (int oc = 0; oc < 64; oc++) {
            const double g = (double)bn[oc] / std::sqrt((double)bn[3 * 64 + oc] + (double)eps);   // weight / sqrt(running_var + eps)
            for (int i = 0; i < cin * 9; i++) wf[l][(size_t)oc * cin * 9 + i] = (double)cw[(size_t)oc * cin * 9 + i] * g;
            bias[(size_t)l * 64 + oc] = (float)((double)bn[64 + oc] - (double)bn[2 * 64 + oc] * g);
        }
    };
    fold(0, w->conv_w, 6, w->bn, w->bn_eps);
    for (int b = 0; b < blocks; b++)
        for (int c = 0; c < 2; c++)
            fold(1 + 2 * b + c, w->block_conv_w + ((size_t)b * 2 + c) * 64 * 64 * 9, 64, w->block_bn + ((size_t)b * 2 + c) * 4 * 64, w->block_bn_eps);

    const size_t stem_tap = (size_t)W_STEM_TAP_BYTES / 2, conv_tap = (size_t)W_TAP_BYTES / 2;
    std::vector<__half> w_stem(9 * stem_tap), w_conv((size_t)(L - 1) * W_LAYER_HALVES);
    std::vector<float> w32_stem((size_t)6 * 9 * 64), w32_conv((size_t)(L - 1) * 64 * 9 * 64);
    for (int l = 0; l < L; l++) {
        const int cin = l == 0 ? 6 : 64, chunks = l == 0 ? 2 : 8;
        double mx = 0.0;
        for (double v : wf[l]) mx = std::max(mx, std::fabs(v));
        int ex = 0;
        if (mx > 0.0) { std::frexp(mx, &ex); }                        // mx = f * 2^ex, f in [0.5, 1)
        const double s = std::ldexp(1.0, 14 - ex);                    // scaled maximum in [2^13, 2^14)
        scale[l] = (float)s;
        __half* lw = l == 0 ? w_stem.data() : w_conv.data() + (size_t)(l - 1) * W_LAYER_HALVES;
        float* d32 = l == 0 ? w32_stem.data() : w32_conv.data() + (size_t)(l - 1) * 64 * 9 * 64;
        for (int tap = 0; tap < 9; tap++)
            for (int oc = 0; oc < 64; oc++)
                for (int ic = 0; ic < chunks * 8; ic++) {
                    const double v = ic < cin ? wf[l][((size_t)oc * cin + ic) * 9 + tap] : 0.0;
                    const double vs = v * s;
                    const __half hi = __float2half_rn((float)vs);
                    // w_lo is stored * 2^11 like x_lo: the small accumulator then holds 2^11 (x_hi w_lo + x_lo w_hi)
                    const __half lo = __float2half_rn((float)((vs - (double)__half2float(hi)) * (double)LO_SCALE));
                    // [chunk][row][8]: the stacked tile has 128 rows per chunk (w_hi rows 0..63, w_lo rows 64..127)
                    const size_t hl_hi = ((size_t)(ic / 8) * 128 + oc) * 8 + (ic % 8), hl_lo = hl_hi + 64 * 8;
                    if (l == 0) {
                        lw[(size_t)tap * stem_tap + hl_hi] = hi;
                        lw[(size_t)tap * stem_tap + hl_lo] = lo;
                    } else {
                        lw[(size_t)tap * conv_tap + hl_hi] = hi;
                        lw[(size_t)tap * conv_tap + hl_lo] = lo;
                    }
                    if (ic < cin) d32[((size_t)ic * 9 + tap) * 64 + oc] = (float)v;
                }
    }
    // heads
    std::vector<float> head_w(3 * 64), head_b(3), pfc_t((size_t)2 * NN * ((A + 3) & ~3), 0.0f), pfc_b(A), vfc_w((size_t)3 * NN), vfc_b(3);
    for (int k = 0; k < 3; k++) {
        const float* cw = k < 2 ? w->policy_conv_w + (size_t)k * 64 : w->value_conv_w;
        const float* bn = k < 2 ? w->policy_bn : w->value_bn;
        const int C = k < 2 ? 2 : 1, c = k < 2 ? k : 0;
        const double g = (double)bn[c] / std::sqrt((double)bn[3 * C + c] + (double)w->head_bn_eps);
        for (int i = 0; i < 64; i++) head_w[(size_t)k * 64 + i] = (float)((double)cw[i] * g);
        head_b[k] = (float)((double)bn[C + c] - (double)bn[2 * C + c] * g);
    }
    for (int o = 0; o < A; o++) {
        pfc_b[o] = w->policy_fc_b[o];
        for (int j = 0; j < 2 * NN; j++) pfc_t[(size_t)j * ((A + 3) & ~3) + o] = w->policy_fc_w[(size_t)o * 2 * NN + j];
    }
    for (int i = 0; i < 3 * NN; i++) vfc_w[i] = w->value_fc_w[i];
    for (int i = 0; i < 3; i++) vfc_b[i] = w->value_fc_b[i];

    NetDev& n = e->net;
    n.blocks = blocks;
    int rc;
    if ((rc = upload(e, w_stem, &n.w_stem)) || (rc = upload(e, w_conv, &n.w_conv)) || (rc = upload(e, bias, &n.bias)) ||
        (rc = upload(e, scale, &n.scale)) || (rc = upload(e, head_w, &n.head_w)) || (rc = upload(e, head_b, &n.head_b)) ||
        (rc = upload(e, pfc_t, &n.pfc_t)) || (rc = upload(e, pfc_b, &n.pfc_b)) || (rc = upload(e, vfc_w, &n.vfc_w)) ||
        (rc = upload(e, vfc_b, &n.vfc_b)) || (rc = upload(e, w32_stem, &n.w32_stem)) || (rc = upload(e, w32_conv, &n.w32_conv)))
        return rc;
    {
        void* q = nullptr;
        CK(cudaMalloc(&q, (size_t)e->sms * SKIP_FLOATS_PER_CTA * sizeof(float)));
        CK(cudaMemsetAsync(q, 0, (size_t)e->sms * SKIP_FLOATS_PER_CTA * sizeof(float), e->stream));
        e->net_allocs.push_back(q);
        n.skip = reinterpret_cast<float*>(q);
        n.dbg = nullptr;
        void* f = nullptr;
        CK(cudaMalloc(&f, 4));
        CK(cudaMemsetAsync(f, 0, 4, e->stream));
        e->net_allocs.push_back(f);
        n.overflow = reinterpret_cast<int*>(f);
    }
    CK(cudaStreamSynchronize(e->stream));
    e->have_weights = true;
    return TG_OK;
}

// ---------------------------------------------------------------------------------------------
// The same fold / split / pack on the device, from parameters that already live there (the trainer's tensors):
// tg_load_weights_device.  Every operation mirrors tg_load_weights (float64 BatchNorm fold, per-layer power-of-two
// scale from the largest folded weight, fp16 (hi, lo * 2^11) split, UMMA tile layout), so both loaders produce the
// same bits (tests/test_gpu_dualnet.py::test_device_weight_loader_matches_host_loader).
// ---------------------------------------------------------------------------------------------
struct FoldOut {
    __half* w_stem; __half* w_conv; float* bias; float* scale; float* w32_stem; float* w32_conv;
    float* head_w; float* head_b; float* pfc_t; float* pfc_b; float* vfc_w; float* vfc_b;
};

__global__ void __launch_bounds__(256) k_fold_conv(tg_weights w, FoldOut o, int blocks)
{
    const int l = blockIdx.x, tid = threadIdx.x;                 // one CTA per convolution layer
    const int cin = l == 0 ? 6 : 64, chunks = l == 0 ? 2 : 8;
    const float* cw = l == 0 ? w.conv_w : w.block_conv_w + (size_t)(l - 1) * 64 * 64 * 9;
    const float* bn = l == 0 ? w.bn : w.block_bn + (size_t)(l - 1) * 4 * 64;
    const double eps = (double)(l == 0 ? w.bn_eps : w.block_bn_eps);
    __shared__ double g[64];
    __shared__ double red[256];
    if (tid < 64) {
        g[tid] = (double)bn[tid] / sqrt((double)bn[3 * 64 + tid] + eps);
        o.bias[(size_t)l * 64 + tid] = (float)((double)bn[64 + tid] - (double)bn[2 * 64 + tid] * g[tid]);
    }
    __syncthreads();
    double mx = 0.0;
    for (int i = tid; i < 64 * cin * 9; i += 256) mx = fmax(mx, fabs((double)cw[i] * g[i / (cin * 9)]));
    red[tid] = mx;
    __syncthreads();
    for (int st = 128; st >= 1; st >>= 1) { if (tid < st) red[tid] = fmax(red[tid], red[tid + st]); __syncthreads(); }
    mx = red[0];
    int ex = 0;
    if (mx > 0.0) ex = (int)(((unsigned long long)__double_as_longlong(mx) >> 52) & 0x7ffull) - 1022;    // frexp: mx = f 2^ex, f in [0.5, 1)
    const double s = __longlong_as_double((long long)((unsigned long long)(14 - ex + 1023) << 52));     // ldexp(1.0, 14 - ex)
    if (tid == 0) o.scale[l] = (float)s;
    const size_t stem_tap = (size_t)W_STEM_TAP_BYTES / 2, conv_tap = (size_t)W_TAP_BYTES / 2;
    __half* lw = l == 0 ? o.w_stem : o.w_conv + (size_t)(l - 1) * W_LAYER_HALVES;
    float* d32 = l == 0 ? o.w32_stem : o.w32_conv + (size_t)(l - 1) * 64 * 9 * 64;
    const int per_tap = 64 * chunks * 8;
    for (int i = tid; i < 9 * per_tap; i += 256) {
        const int tap = i / per_tap, r = i - tap * per_tap, oc = r / (chunks * 8), ic = r - oc * (chunks * 8);
        const double v = ic < cin ? (double)cw[((size_t)oc * cin + ic) * 9 + tap] * g[oc] : 0.0;
        const double vs = v * s;
        const __half hi = __float2half_rn((float)vs);
        const __half lo = __float2half_rn((float)((vs - (double)__half2float(hi)) * (double)LO_SCALE));
        const size_t hl_hi = ((size_t)(ic / 8) * 128 + oc) * 8 + (ic % 8), hl_lo = hl_hi + 64 * 8;
        const size_t base = (size_t)tap * (l == 0 ? stem_tap : conv_tap);
        lw[base + hl_hi] = hi; lw[base + hl_lo] = lo;
        if (ic < cin) d32[((size_t)ic * 9 + tap) * 64 + oc] = (float)v;
    }
}

__global__ void __launch_bounds__(256) k_fold_heads(tg_weights w, FoldOut o, int NN, int A)
{
    const int tid = blockIdx.x * 256 + threadIdx.x, nthreads = gridDim.x * 256;
    if (tid < 3) {
        const int k = tid;
        const float* bn = k < 2 ? w.policy_bn : w.value_bn;
        const int C = k < 2 ? 2 : 1, c = k < 2 ? k : 0;
        const double g = (double)bn[c] / sqrt((double)bn[3 * C + c] + (double)w.head_bn_eps);
        const float* cw = k < 2 ? w.policy_conv_w + (size_t)k * 64 : w.value_conv_w;
        for (int i = 0; i < 64; i++) o.head_w[(size_t)k * 64 + i] = (float)((double)cw[i] * g);
        o.head_b[k] = (float)((double)bn[C + c] - (double)bn[2 * C + c] * g);
        o.vfc_b[k] = w.value_fc_b[k];
    }
    const int A4 = (A + 3) & ~3;
    for (int i = tid; i < A * 2 * NN; i += nthreads) { const int oo = i / (2 * NN), j = i - oo * 2 * NN; o.pfc_t[(size_t)j * A4 + oo] = w.policy_fc_w[i]; }
    for (int i = tid; i < A; i += nthreads) o.pfc_b[i] = w.policy_fc_b[i];
    for (int i = tid; i < 3 * NN; i += nthreads) o.vfc_w[i] = w.value_fc_w[i];
}

extern "C" int tg_load_weights_device(tg_engine* e, const tg_weights* w)
{
    if (!e || !w) return fail(TG_ERR_ARG, "null argument");
    CK(cudaSetDevice(e->cfg.device));
    CK(cudaStreamSynchronize(e->stream));
    free_net(e);
    const int blocks = e->cfg.net_blocks, L = 1 + 2 * blocks, NN = e->NN, A = e->A, A4 = (A + 3) & ~3;
    FoldOut o{};
    auto dz = [&](void** p, size_t bytes) {
        if (cudaMalloc(p, std::max<size_t>(bytes, 1)) != cudaSuccess) return 1;
        e->net_allocs.push_back(*p);
        return cudaMemsetAsync(*p, 0, bytes, e->stream) != cudaSuccess ? 1 : 0;
    };
    int bad = 0;
    bad |= dz((void**)&o.w_stem, 9 * (size_t)W_STEM_TAP_BYTES); bad |= dz((void**)&o.w_conv, (size_t)(L - 1) * W_LAYER_HALVES * 2);
    bad |= dz((void**)&o.bias, (size_t)L * 64 * 4); bad |= dz((void**)&o.scale, (size_t)L * 4);
    bad |= dz((void**)&o.w32_stem, (size_t)6 * 9 * 64 * 4); bad |= dz((void**)&o.w32_conv, (size_t)(L - 1) * 64 * 9 * 64 * 4);
    bad |= dz((void**)&o.head_w, 3 * 64 * 4); bad |= dz((void**)&o.head_b, 3 * 4); bad |= dz((void**)&o.pfc_t, (size_t)2 * NN * A4 * 4);
    bad |= dz((void**)&o.pfc_b, (size_t)A * 4); bad |= dz((void**)&o.vfc_w, (size_t)3 * NN * 4); bad |= dz((void**)&o.vfc_b, 3 * 4);
    void* skip = nullptr; void* ovf = nullptr;
    bad |= dz(&skip, (size_t)e->sms * SKIP_FLOATS_PER_CTA * sizeof(float)); bad |= dz(&ovf, 4);
    if (bad) { free_net(e); return fail(TG_ERR_CUDA, "weight buffers"); }
    k_fold_conv<<<L, 256, 0, e->stream>>>(*w, o, blocks);
    k_fold_heads<<<64, 256, 0, e->stream>>>(*w, o, NN, A);
    e->launches += 2;
    CK(cudaGetLastError());
    NetDev& n = e->net;
    n.blocks = blocks;
    n.w_stem = o.w_stem; n.w_conv = o.w_conv; n.bias = o.bias; n.scale = o.scale; n.head_w = o.head_w; n.head_b = o.head_b;
    n.pfc_t = o.pfc_t; n.pfc_b = o.pfc_b; n.vfc_w = o.vfc_w; n.vfc_b = o.vfc_b; n.w32_stem = o.w32_stem; n.w32_conv = o.w32_conv;
    n.skip = reinterpret_cast<float*>(skip); n.dbg = nullptr; n.overflow = reinterpret_cast<int*>(ovf);
    CK(cudaStreamSynchronize(e->stream));
    e->have_weights = true;
    return TG_OK;
}

// ---------------------------------------------------------------------------------------------
extern "C" int tg_reset(tg_engine* e, const uint8_t* mask, const uint64_t* game_ids, const uint8_t* never_resign)
{
    if (!e) return fail(TG_ERR_ARG, "null engine");
    CK(cudaSetDevice(e->cfg.device));
    const int games = e->cfg.games;
    // staging block [ids u64 x games | mask u8 x games | never_resign u8 x games]: the previous reset's upload must have
    // left the pinned block before it is rewritten (one event wait, normally long past)
    if (e->reset_pending) { CK(cudaEventSynchronize(e->ev_reset)); e->reset_pending = false; }
    u64* h_ids = reinterpret_cast<u64*>(e->h_reset);
    uint8_t* h_mask = e->h_reset + (size_t)games * 8; uint8_t* h_nr = h_mask + games;
    for (int g = 0; g < games; g++) h_ids[g] = game_ids ? game_ids[g] : (u64)g;
    if (mask) memcpy(h_mask, mask, games);
    if (never_resign) memcpy(h_nr, never_resign, games);
    CK(cudaMemcpyAsync(e->d_reset, e->h_reset, (size_t)games * 10, cudaMemcpyHostToDevice, e->stream));
    CK(cudaEventRecord(e->ev_reset, e->stream));
    e->reset_pending = true;
    const u64* d_ids = reinterpret_cast<const u64*>(e->d_reset);
    const uint8_t* d_mask = mask ? e->d_reset + (size_t)games * 8 : nullptr;
    const uint8_t* d_nr = never_resign ? e->d_reset + (size_t)games * 9 : nullptr;
    DISPATCH_N(e, (k_reset<BN><<<search_grid(e), SEARCH_WARPS * 32, search_smem(BN), e->stream>>>(e->D, d_mask, d_ids, d_nr)));
    e->launches++;
    CK(cudaGetLastError());
    return TG_OK;                                   // asynchronous: ordered before everything queued on the engine's stream later
}

extern "C" int tg_set_to_move(tg_engine* e, const int32_t* colors)
{
    if (!e || !colors) return fail(TG_ERR_ARG, "null argument");
    CK(cudaSetDevice(e->cfg.device));
    CK(cudaMemcpy2DAsync(e->D.gs + GS_COLOR, GS_STRIDE * sizeof(int), colors, sizeof(int), sizeof(int), e->cfg.games,
                         cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return TG_OK;
}

extern "C" int tg_play(tg_engine* e, const int16_t* moves, const uint8_t* colors, const int32_t* counts, int32_t stride, tg_ply_dump* dump)
{
    if (!e || !moves || !counts || stride < 1) return fail(TG_ERR_ARG, "bad argument");
    const int games = e->cfg.games;
    const int plies = dump ? dump->plies : 0;
    if (dump && plies < 1) return fail(TG_ERR_ARG, "dump->plies must be positive");
    for (int g = 0; g < games; g++) {
        if (counts[g] < 0 || counts[g] > stride) return fail(TG_ERR_ARG, "counts[g] outside [0, stride]");
        if (dump && counts[g] > plies) return fail(TG_ERR_ARG, "dump->plies smaller than counts[g]");
    }
    CK(cudaSetDevice(e->cfg.device));
    const size_t nm = (size_t)games * stride;
    std::vector<void*> tmp;                         // every temporary is released on every exit path
    struct Guard { std::vector<void*>& v; ~Guard() { for (void* p : v) cudaFree(p); } } guard{tmp};
    auto da = [&](void** p, size_t bytes) { if (cudaMalloc(p, std::max<size_t>(bytes, 1)) != cudaSuccess) return 1; tmp.push_back(*p); return 0; };
    int16_t* d_moves = nullptr; uint8_t* d_colors = nullptr; int* d_counts = nullptr;
    int bad = da((void**)&d_moves, nm * 2) | da((void**)&d_counts, (size_t)games * 4);
    if (colors) bad |= da((void**)&d_colors, nm);
    PlyDump pd{};
    const size_t np = (size_t)games * std::max(plies, 1);
    if (dump) {
        bad |= da((void**)&pd.color, np * e->CELLS); bad |= da((void**)&pd.libs, np * e->CELLS * 2); bad |= da((void**)&pd.size, np * e->CELLS * 2);
        bad |= da((void**)&pd.scal, np * 5 * 4); bad |= da((void**)&pd.hash, np * 8);
        bad |= da((void**)&pd.legal, np * 2 * e->NN); bad |= da((void**)&pd.satari, np * 2 * e->NN * 2);
        bad |= da((void**)&pd.eye, np * 2 * e->NN); bad |= da((void**)&pd.cand, np * 2 * e->NN); bad |= da((void**)&pd.score, np * 4);
        if (dump->tt_score) bad |= da((void**)&pd.tt_score, np * 4);
        pd.stride = plies;
    }
    if (bad) return fail(TG_ERR_CUDA, "temporary allocation failed");
    CK(cudaMemcpyAsync(d_moves, moves, nm * 2, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(d_counts, counts, (size_t)games * 4, cudaMemcpyHostToDevice, e->stream));
    if (colors) CK(cudaMemcpyAsync(d_colors, colors, nm, cudaMemcpyHostToDevice, e->stream));
    DISPATCH_N(e, (k_play<BN><<<search_grid(e), SEARCH_WARPS * 32, search_smem(BN), e->stream>>>(e->D, d_moves, d_colors, d_counts, stride, pd, dump ? 1 : 0)));
    e->launches++;
    CK(cudaGetLastError());
    if (dump) {
#define CP_OUT(field, bytes) if (dump->field) CK(cudaMemcpyAsync(dump->field, pd.field, (bytes), cudaMemcpyDeviceToHost, e->stream))
        CP_OUT(color, np * e->CELLS); CP_OUT(libs, np * e->CELLS * 2); CP_OUT(size, np * e->CELLS * 2); CP_OUT(scal, np * 5 * 4);
        CP_OUT(hash, np * 8); CP_OUT(legal, np * 2 * e->NN); CP_OUT(satari, np * 2 * e->NN * 2); CP_OUT(eye, np * 2 * e->NN);
        CP_OUT(cand, np * 2 * e->NN); CP_OUT(score, np * 4); CP_OUT(tt_score, np * 4);
#undef CP_OUT
    }
    CK(cudaStreamSynchronize(e->stream));
    return TG_OK;
}

// ---------------------------------------------------------------------------------------------
// evaluator: slot bases, feature planes, network (device-side slot count; no host round trip)
// ---------------------------------------------------------------------------------------------
// n_direct < 0: the slot count is read from device memory (written by k_scan); otherwise it is a launch argument, so that
// back-to-back tg_forward_device calls cannot race on a staged count.
template <int BN> static int launch_net(tg_engine* e, int use_logit, int max_slots, int n_direct = -1, bool from_snapshots = false)
{
    const Dev& D = e->D;
    const int* n_ptr = n_direct < 0 ? D.n_slots : nullptr;
    if (e->cfg.evaluator == TG_EVAL_HASHNET || e->cfg.evaluator == TG_EVAL_HASHNET2) {
        k_hashnet<BN><<<(max_slots + 3) / 4, 128, 0, e->stream>>>(D.planes, n_ptr, n_direct, use_logit, D.policy, D.value,
                                                                  e->cfg.evaluator == TG_EVAL_HASHNET2 ? 1 : 0);
        e->launches++;
    } else if (e->cfg.evaluator == TG_EVAL_DUALNET_TC) {
        if (!e->have_weights) return fail(TG_ERR_STATE, "tg_load_weights has not been called");
        constexpr int G = TcGroup<BN>::G;
        const int groups = (max_slots + G - 1) / G;
        const int grid = std::max(1, std::min(e->sms, groups));
        k_dualnet_tc<BN, G><<<grid, TC_THREADS, NetGeo<BN, G>::SMEM_BYTES, e->stream>>>(e->net, D.planes, n_ptr, n_direct, use_logit, D.policy, D.value,
            from_snapshots ? D.snap : nullptr, D.slot_src, (int)Snap<BN>::BYTES);
        e->launches++;
    } else {
        if (!e->have_weights) return fail(TG_ERR_STATE, "tg_load_weights has not been called");
        // the CUDA-core path needs the slot count on the host (reference path only)
        int n = n_direct;
        if (n_direct < 0) {
            CK(cudaMemcpyAsync(&n, D.n_slots, 4, cudaMemcpyDeviceToHost, e->stream));
            CK(cudaStreamSynchronize(e->stream));
        }
        const int L = 1 + 2 * e->net.blocks;
        const int smem = 64 * (BN + 2) * (BN + 2) * 4;
        for (int s0 = 0; s0 < n; s0 += e->simt_chunk) {
            const int ns = std::min(e->simt_chunk, n - s0);
            const float* in = D.planes + (size_t)s0 * 6 * BN * BN;
            k_conv3x3_simt<BN><<<ns, 256, smem, e->stream>>>(in, 6, e->net.w32_stem, e->net.bias, nullptr, e->act[0], ns);
            int cur = 0;
            for (int l = 1; l < L; l += 2) {
                const int t1 = (cur + 1) % 3, t2 = (cur + 2) % 3;
                k_conv3x3_simt<BN><<<ns, 256, smem, e->stream>>>(e->act[cur], 64, e->net.w32_conv + (size_t)(l - 1) * 64 * 9 * 64,
                                                                 e->net.bias + (size_t)l * 64, nullptr, e->act[t1], ns);
                k_conv3x3_simt<BN><<<ns, 256, smem, e->stream>>>(e->act[t1], 64, e->net.w32_conv + (size_t)l * 64 * 9 * 64,
                                                                 e->net.bias + (size_t)(l + 1) * 64, e->act[cur], e->act[t2], ns);
                cur = t2;
                e->launches += 2;
            }
            k_heads_simt<BN><<<ns, 256, 0, e->stream>>>(e->net, e->act[cur], ns, use_logit, D.policy + (size_t)s0 * (BN * BN + 1), D.value + (size_t)s0 * 3);
            e->launches += 2;
        }
    }
    CK(cudaGetLastError());
    return 0;
}

template <int BN> static int launch_eval(tg_engine* e, int use_logit, int max_slots, int* ev_idx)
{
    const Dev& D = e->D;
    k_scan<<<1, 1024, 0, e->stream>>>(D);
    // the tensor-core evaluator reads the leaf snapshots itself (fused feature planes); the others take fp32 planes
    const bool fused = e->cfg.evaluator == TG_EVAL_DUALNET_TC && !e->unfused_planes;
    if (fused) k_slotmap<<<D.games, 256, 0, e->stream>>>(D);
    else k_planes<BN><<<D.games, 256, 16 * Snap<BN>::BYTES, e->stream>>>(D);
    e->launches += 2;
    const bool timed = ev_idx && *ev_idx + 2 <= (int)e->events.size();
    if (timed) CK(cudaEventRecord(e->events[(*ev_idx)++], e->stream));
    const int rc = launch_net<BN>(e, use_logit, std::min(max_slots, e->slot_cap), -1, fused);
    if (rc) return rc;
    if (timed) CK(cudaEventRecord(e->events[(*ev_idx)++], e->stream));
    return 0;
}

// one PUCT batch with the block-per-game kernels: descents, evaluation, backup
template <int BN, int NT> static int puct_iter_blk(tg_engine* e, int visits, int batch, int strict, int max_slots, int* ev)
{
    const Dev& D = e->D;
    if (e->puct_defer && e->puct_wave) {
        const int nsq = std::min(WAVE_SQRT_MAX, visits + batch + 2);              // sqrt table: visits + virtual losses + 1 of any node
        if (NT == 512 && e->wave_gt == 128)
            k_wave_puct_blk<BN, NT, (NT == 512 ? 128 : NT / 4)><<<D.games, NT, wave_smem<BN, NT, (NT == 512 ? 128 : NT / 4)>(nsq), e->stream>>>(D, e->eye2, visits, batch, strict, nsq);
        else
            k_wave_puct_blk<BN, NT, (NT == 512 ? 64 : NT / 4)><<<D.games, NT, wave_smem<BN, NT, (NT == 512 ? 64 : NT / 4)>(nsq), e->stream>>>(D, e->eye2, visits, batch, strict, nsq);
        k_expand_leaves_blk<BN, NT><<<dim3(std::min(batch, D.cap), D.games), NT, sizeof(ExpandSmem<BN, NT>), e->stream>>>(D, e->eye2);
        e->launches++;
    } else if (e->puct_defer) {
        int slots = e->walk_slots;
        while (slots > 2 && walk_smem<BN, NT>(slots) > 227 * 1024) slots--;
        k_walk_puct_blk<BN, NT><<<D.games, NT, walk_smem<BN, NT>(slots), e->stream>>>(D, e->eye2, visits, batch, strict, slots);
        k_expand_leaves_blk<BN, NT><<<dim3(std::min(batch, D.cap), D.games), NT, sizeof(ExpandSmem<BN, NT>), e->stream>>>(D, e->eye2);
        e->launches++;
    } else
        k_descend_puct_blk<BN, NT><<<D.games, NT, sizeof(BlkSmem<BN, NT>), e->stream>>>(D, e->eye2, visits, batch, strict);
    const int rc = launch_eval<BN>(e, 0, max_slots, ev);
    if (e->puct_defer) { k_backup_priors_blk<BN><<<dim3(std::min(batch, D.cap), D.games), 128, 0, e->stream>>>(D, 0); e->launches++; }
    k_backup_blk<BN, NT><<<D.games, NT, 256 * 136, e->stream>>>(D, 0, e->puct_defer ? 1 : 0);
    return rc;
}

// Queue one move of every game on the engine's stream (no host synchronisation); tg_collect waits for it.
extern "C" int tg_genmove_async(tg_engine* e, int32_t mode, int32_t visits, int32_t strict, int32_t play, int32_t root_arrays)
{
    if (!e) return fail(TG_ERR_ARG, "null engine");
    if (e->step_pending) return fail(TG_ERR_STATE, "a step is already in flight: call tg_collect first");
    if (visits < 1 || visits > e->cfg.max_visits) return fail(TG_ERR_ARG, "visits outside [1, max_visits]");
    if (mode != TG_MODE_SH && mode != TG_MODE_PUCT) return fail(TG_ERR_ARG, "bad mode");
    CK(cudaSetDevice(e->cfg.device));
    const int games = e->cfg.games, grid = search_grid(e), thr = SEARCH_WARPS * 32;
    const int use_logit = mode == TG_MODE_SH ? 1 : 0;
    int rc = 0, ev = 2;
    Dev& D = e->D;
    if (mode == TG_MODE_SH) { D.cap = e->cap_sh; D.max_depth = e->depth_sh; } else { D.cap = e->cap_puct; D.max_depth = e->depth_puct; }
    if (mode == TG_MODE_PUCT && e->snap_plan > 0 && D.snap_levels == 0) {        // first PUCT move of a warp-per-game pool
        const int rc0 = dalloc(e, &D.snapb, (size_t)games * e->snap_plan * D.snap_words, false);
        if (rc0) return rc0;
        D.snap_levels = e->snap_plan;
    }
    const int max_moves = 2 * e->NN;
    CK(cudaEventRecord(e->events[0], e->stream));
    DISPATCH_N(e, {
        const size_t sm = search_smem(BN);
        k_root_begin<BN><<<grid, thr, sm, e->stream>>>(D);
        e->launches++;
        rc = launch_eval<BN>(e, use_logit, games, &ev);
        if (!rc) {
            k_backup<BN><<<grid, thr, 0, e->stream>>>(D, use_logit);
            k_root_post<BN><<<grid, thr, 0, e->stream>>>(D, mode, visits);
            e->launches += 2;
            if (mode == TG_MODE_SH) {
                int phases = 1;
                for (int m = 1; m <= 16; m++) phases = std::max(phases, sh_phases_host(m, visits, nullptr));
                for (int p = 0; p < phases && !rc; p++) {
                    k_descend_sh<BN><<<grid, thr, sm, e->stream>>>(D);
                    e->launches++;
                    rc = launch_eval<BN>(e, 1, (int)std::min<size_t>((size_t)games * e->cap_sh, (size_t)e->slot_cap), &ev);
                    k_backup<BN><<<grid, thr, 0, e->stream>>>(D, 1);
                    e->launches++;
                }
            } else {
                const int batch = e->cfg.batch_size;
                const int iters = (visits + batch - 1) / batch + 1;
                for (int it = 0; it < iters && !rc; it++) {
                    if (e->puct_warp) {                  // warp-per-game kernels (TG_PUCT_WARP=1: A/B measurements)
                        const int nsq = std::min(PUCT_SQRT_MAX, visits + batch + 2);   // sqrt(visits + virtual losses + 1) table
                        if (batch == 1 && D.snap_levels > 0)
                            k_descend_puct_snap<BN><<<grid, thr, sm + 4 * Geo<BN>::CP * 8 + nsq * 8, e->stream>>>(D, visits, strict, nsq);
                        else
                            k_descend_puct<BN><<<grid, thr, sm + 4 * Geo<BN>::CP * 8 + nsq * 8, e->stream>>>(D, visits, batch, strict, nsq);
                        rc = launch_eval<BN>(e, 0, (int)std::min<size_t>((size_t)games * batch, (size_t)e->slot_cap), iters <= 24 ? &ev : nullptr);
                        k_backup<BN><<<grid, thr, 0, e->stream>>>(D, 0);
                    } else {                             // block-per-game (tg_block.cuh): a ply runs puct_nt threads wide
                        const int ms_ = (int)std::min<size_t>((size_t)games * batch, (size_t)e->slot_cap);
                        int* evp = iters <= 24 ? &ev : nullptr;
                        if (e->puct_nt == 512) rc = puct_iter_blk<BN, 512>(e, visits, batch, strict, ms_, evp);
                        else if (e->puct_nt == 128) rc = puct_iter_blk<BN, 128>(e, visits, batch, strict, ms_, evp);
                        else rc = puct_iter_blk<BN, 256>(e, visits, batch, strict, ms_, evp);
                    }
                    e->launches += 2;
                }
            }
            if (!rc) {
                k_move_end<BN><<<grid, thr, sm, e->stream>>>(D, mode, play, e->cfg.komi, max_moves);
                e->launches++;
            }
        }
    });
    if (rc) return rc;
    CK(cudaGetLastError());
    CK(cudaEventRecord(e->events[1], e->stream));
    e->n_eval_events = ev;
    // results -> pinned staging (the caller's buffers are filled by tg_collect)
    CK(cudaMemcpyAsync(e->h_gs, D.gs, (size_t)games * GS_STRIDE * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    if (root_arrays) {
        CK(cudaMemcpyAsync(e->h_action, D.out_action, (size_t)games * e->AP * 2, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaMemcpyAsync(e->h_improved, D.out_improved, (size_t)games * e->AP * 8, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaMemcpyAsync(e->h_visits, D.out_visits, (size_t)games * e->AP * 4, cudaMemcpyDeviceToHost, e->stream));
    }
    CK(cudaEventRecord(e->ev_step, e->stream));
    e->step_pending = true; e->step_arrays = root_arrays != 0;
    return TG_OK;
}

extern "C" int tg_collect(tg_engine* e, tg_step_result* out)
{
    if (!e) return fail(TG_ERR_ARG, "null engine");
    if (!e->step_pending) return fail(TG_ERR_STATE, "no step in flight");
    CK(cudaSetDevice(e->cfg.device));
    CK(cudaEventSynchronize(e->ev_step));
    e->step_pending = false;
    const int games = e->cfg.games;
    Dev& D = e->D;
    if (D.prof) {
        long long h[16];
        CK(cudaMemcpy(h, D.prof, sizeof h, cudaMemcpyDeviceToHost));
        CK(cudaMemset(D.prof, 0, 64 * sizeof(long long)));
        if (e->puct_wave) {
            const double st = (double)std::max(1ll, h[5]), g0 = (double)std::max(1ll, h[6]);
            fprintf(stderr, "wavefront walk (game 0): prologue %lld cycles, %lld steps, descents in flight per step %.2f, cycles per step: scoring %.0f  leader %.0f  fetch %.0f"
                            "  [group 0 when it scores (%lld steps): operands %.0f, + divisions %.0f, + warp argmax %.0f; after B1: fold %.0f, + leader updates %.0f]\n",
                    h[0], h[5], (double)h[1] / st, (double)h[8] / st, (double)h[9] / st, (double)h[10] / st, h[6], (double)h[2] / g0, (double)h[3] / g0, (double)h[4] / g0,
                    (double)h[7] / g0, (double)h[11] / g0);
        }
        fprintf(stderr, "descend profile (game 0, cycles): copy %lld  select %lld (%lld)  put_stone %lld  expand %lld (%lld)  push_leaf %lld (%lld)"
                        "  [block kernels, select: scores %lld  argmax %lld  row wait %lld]  [warp put_stone: captures %lld (%lld cycles), passes %lld (%lld)]\n",
                h[0], h[1], h[5], h[2], h[3], h[6], h[4], h[7], h[8], h[9], h[10], h[12], h[13], h[14], h[15]);
    }
    { const int orc = check_net_overflow(e); if (orc) return orc; }
    CK(cudaEventElapsedTime(&e->last_ms, e->events[0], e->events[1]));
    e->last_eval_ms = 0.f;
    for (int i = 2; i + 1 < e->n_eval_events; i += 2) { float ms = 0.f; CK(cudaEventElapsedTime(&ms, e->events[i], e->events[i + 1])); e->last_eval_ms += ms; }
    int64_t evals = 0, uevals = 0; int any_err = 0;
    for (int g = 0; g < games; g++) {
        const int* gs = e->h_gs + (size_t)g * GS_STRIDE;
        evals += gs[GS_EVALS]; uevals += gs[GS_UEVALS];
        any_err |= gs[GS_ERROR];
        if (!out) continue;
        if (out->move) out->move[g] = gs[GS_LAST_MOVE];
        if (out->color) out->color[g] = gs[GS_LAST_COLOR];
        if (out->num_children) out->num_children[g] = gs[GS_ROOT_K];
        if (out->finished) out->finished[g] = gs[GS_FINISHED];
        if (out->winner) out->winner[g] = gs[GS_WINNER];
        if (out->resigned) out->resigned[g] = gs[GS_RESIGNED];
        if (out->score) { float f; memcpy(&f, &gs[GS_SCORE], 4); out->score[g] = f; }
        if (out->error) out->error[g] = gs[GS_ERROR];
        if (out->n_moves) out->n_moves[g] = gs[GS_NMOVES];
    }
    e->last_eval_slots = uevals;
    if (out) {
        if ((out->action || out->improved || out->visits) && !e->step_arrays)
            return fail(TG_ERR_STATE, "root arrays were not requested from tg_genmove_async");
        if (out->action) memcpy(out->action, e->h_action, (size_t)games * e->AP * 2);
        if (out->improved) memcpy(out->improved, e->h_improved, (size_t)games * e->AP * 8);
        if (out->visits) memcpy(out->visits, e->h_visits, (size_t)games * e->AP * 4);
        if (out->evals) { out->evals[0] = evals; out->evals[1] = uevals; }
    }
    if (any_err && !(out && out->error)) return fail(TG_ERR_SEARCH, "search error flags set (history/depth/node/queue overflow); pass tg_step_result.error to inspect");
    return TG_OK;
}

extern "C" int tg_genmove(tg_engine* e, int32_t mode, int32_t visits, int32_t strict, int32_t play, tg_step_result* out)
{
    const int arrays = out && (out->action || out->improved || out->visits);
    const int rc = tg_genmove_async(e, mode, visits, strict, play, arrays);
    return rc ? rc : tg_collect(e, out);
}

// ---------------------------------------------------------------------------------------------
// Finished-game records: rows of the device ring -> pinned staging -> SGF text / files (C++ writer, tg_record.cpp).
//   tg_fetch_records queues the copies on the engine's stream, i.e. BEFORE a later tg_reset / tg_genmove_async can reuse
//   the rows; tg_format_records / tg_write_records wait for them and run on the host while the GPU already works on
//   the next step.
// ---------------------------------------------------------------------------------------------
extern "C" int tg_fetch_records(tg_engine* e, const int32_t* games_list, int32_t n)
{
    if (!e || n < 0 || (n > 0 && !games_list)) return fail(TG_ERR_ARG, "bad argument");
    if (!e->D.rec_moves) return fail(TG_ERR_STATE, "engine was created without record_ring");
    if (e->step_pending) return fail(TG_ERR_STATE, "collect the step in flight first");
    CK(cudaSetDevice(e->cfg.device));
    if (e->rec_pending) { CK(cudaEventSynchronize(e->ev_rec)); e->rec_pending = false; }
    const Dev& D = e->D;
    const int AP = e->AP, MM = D.rec_moves;
    e->fetched.clear();
    size_t total = 0;
    for (int i = 0; i < n; i++) {
        const int g = games_list[i];
        if (g < 0 || g >= e->cfg.games) return fail(TG_ERR_ARG, "game index out of range");
        const int* gs = e->h_gs + (size_t)g * GS_STRIDE;                   // state block of the last collected step
        tg_engine::Fetched f;
        f.game = g; f.n_moves = std::min(gs[GS_NMOVES], MM); f.winner = gs[GS_WINNER]; f.resigned = gs[GS_RESIGNED];
        memcpy(&f.score, &gs[GS_SCORE], 4);
        f.off = total;
        total += ((size_t)f.n_moves * (2 + 1 + 2 + (size_t)AP * 10) + 15) & ~(size_t)15;
        e->fetched.push_back(f);
    }
    if (total > e->h_rec_cap) {
        if (e->h_rec) cudaFreeHost(e->h_rec);
    if (e->d_emit) cudaFree(e->d_emit);
        e->h_rec = nullptr; e->h_rec_cap = 0;
        // pinned allocations cost tens of milliseconds: size generously once (steady state: games / game length finish per
        // step), double on the rare overflow (a pool whose games all end in the same step)
        const size_t want = std::max<size_t>(2 * total, (size_t)32 << 20);
        CK(cudaMallocHost(&e->h_rec, want));
        e->h_rec_cap = want;
    }
    for (const auto& f : e->fetched) {
        if (f.n_moves == 0) continue;
        const size_t m = (size_t)f.n_moves, r0 = (size_t)f.game * MM;
        unsigned char* p = e->h_rec + f.off;                               // [improved f64 m*AP | action i16 m*AP | move i16 m | k i16 m | color u8 m]
        CK(cudaMemcpyAsync(p, D.rec_improved + r0 * AP, m * AP * 8, cudaMemcpyDeviceToHost, e->stream)); p += m * AP * 8;
        CK(cudaMemcpyAsync(p, D.rec_action + r0 * AP, m * AP * 2, cudaMemcpyDeviceToHost, e->stream)); p += m * AP * 2;
        CK(cudaMemcpyAsync(p, D.rec_move + r0, m * 2, cudaMemcpyDeviceToHost, e->stream)); p += m * 2;
        CK(cudaMemcpyAsync(p, D.rec_k + r0, m * 2, cudaMemcpyDeviceToHost, e->stream)); p += m * 2;
        CK(cudaMemcpyAsync(p, D.rec_color + r0, m, cudaMemcpyDeviceToHost, e->stream));
    }
    CK(cudaEventRecord(e->ev_rec, e->stream));
    e->rec_pending = true;
    return TG_OK;
}

int tg_record_text(int n, int n_moves, const int16_t* moves, const uint8_t* colors, const int16_t* ks, const int16_t* action,
                   const double* improved, int stride, int winner, int resigned, double score, double komi, std::string& out);

static int wait_records(tg_engine* e)
{
    if (e->rec_pending) { CK(cudaEventSynchronize(e->ev_rec)); e->rec_pending = false; }
    return 0;
}

static void record_text_of(const tg_engine* e, const tg_engine::Fetched& f, std::string& out)
{
    const size_t m = (size_t)f.n_moves, AP = (size_t)e->AP;
    const unsigned char* p = e->h_rec + f.off;
    const double* improved = reinterpret_cast<const double*>(p); p += m * AP * 8;
    const int16_t* action = reinterpret_cast<const int16_t*>(p); p += m * AP * 2;
    const int16_t* mv = reinterpret_cast<const int16_t*>(p); p += m * 2;
    const int16_t* ks = reinterpret_cast<const int16_t*>(p); p += m * 2;
    const uint8_t* col = p;
    tg_record_text(e->N, f.n_moves, mv, col, ks, action, improved, e->AP, f.winner, f.resigned, (double)f.score, (double)e->cfg.komi, out);
}

// raw arrays of fetched record i (tests, training-data emitters on the host)
extern "C" int tg_fetched_record(tg_engine* e, int32_t i, int32_t* n_moves, int16_t* move, uint8_t* color, int16_t* num_children,
                                 int16_t* action, double* improved)
{
    if (!e || i < 0 || i >= (int)e->fetched.size()) return fail(TG_ERR_ARG, "bad argument");
    CK(cudaSetDevice(e->cfg.device));
    int rc = wait_records(e); if (rc) return rc;
    const auto& f = e->fetched[i];
    const size_t m = (size_t)f.n_moves, AP = (size_t)e->AP;
    const unsigned char* p = e->h_rec + f.off;
    if (n_moves) *n_moves = f.n_moves;
    if (improved) memcpy(improved, p, m * AP * 8);
    p += m * AP * 8;
    if (action) memcpy(action, p, m * AP * 2);
    p += m * AP * 2;
    if (move) memcpy(move, p, m * 2);
    p += m * 2;
    if (num_children) memcpy(num_children, p, m * 2);
    p += m * 2;
    if (color) memcpy(color, p, m);
    return TG_OK;
}

// SGF text of every fetched game, concatenated; offsets[i]..offsets[i+1] delimit game i.  Returns the total size, or
// the size needed (as a negative number minus one is never used: the call fails with TG_ERR_ARG when cap is too small
// and tg_last_error names the size).
extern "C" int64_t tg_format_records(tg_engine* e, char* buf, int64_t cap, int64_t* offsets)
{
    if (!e || !buf || !offsets) return fail(TG_ERR_ARG, "null argument");
    cudaSetDevice(e->cfg.device);
    if (wait_records(e)) return TG_ERR_CUDA;
    const int n = (int)e->fetched.size();
    std::vector<std::string> texts(n);
    parallel_for(n, [&](int i) { record_text_of(e, e->fetched[i], texts[i]); });
    int64_t total = 0;
    for (int i = 0; i < n; i++) { offsets[i] = total; total += (int64_t)texts[i].size(); }
    offsets[n] = total;
    if (total > cap) return fail(TG_ERR_ARG, "buffer too small: need " + std::to_string(total) + " bytes");
    for (int i = 0; i < n; i++) memcpy(buf + offsets[i], texts[i].data(), texts[i].size());
    return total;
}

// sgf/selfplay_record.py:104-108: write <dir>/<index[i]>.sgf for every fetched game (formatting and file output on
// host worker threads).  Returns the number of root moves recorded in the files.
extern "C" int64_t tg_write_records(tg_engine* e, const char* dir, const int64_t* index)
{
    if (!e || !dir || !index) return fail(TG_ERR_ARG, "null argument");
    cudaSetDevice(e->cfg.device);
    if (wait_records(e)) return TG_ERR_CUDA;
    const int n = (int)e->fetched.size();
    std::vector<int> bad(n, 0);
    parallel_for(n, [&](int i) {
        std::string text;
        record_text_of(e, e->fetched[i], text);
        const std::string path = std::string(dir) + "/" + std::to_string((long long)index[i]) + ".sgf";
        FILE* fp = fopen(path.c_str(), "wb");
        if (!fp) { bad[i] = 1; return; }
        if (fwrite(text.data(), 1, text.size(), fp) != text.size()) bad[i] = 1;
        if (fclose(fp) != 0) bad[i] = 1;
    });
    int64_t moves = 0;
    for (int i = 0; i < n; i++) { if (bad[i]) return fail(TG_ERR_ARG, "could not write a record file under " + std::string(dir)); moves += e->fetched[i].n_moves; }
    return moves;
}

// ---------------------------------------------------------------------------------------------
extern "C" int tg_planes(tg_engine* e, float* outp)
{
    if (!e || !outp) return fail(TG_ERR_ARG, "null argument");
    CK(cudaSetDevice(e->cfg.device));
    Dev& D = e->D;
    D.cap = e->cap_max; D.max_depth = 1;
    if (e->slot_cap < e->cfg.games) return fail(TG_ERR_STATE, "evaluator batch smaller than the game pool");
    DISPATCH_N(e, {
        k_snapshot_roots<BN><<<search_grid(e), SEARCH_WARPS * 32, search_smem(BN), e->stream>>>(D);
        k_scan<<<1, 1024, 0, e->stream>>>(D);
        k_planes<BN><<<D.games, 256, 16 * Snap<BN>::BYTES, e->stream>>>(D);
    });
    e->launches += 3;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(outp, D.planes, (size_t)e->cfg.games * e->PLANES * 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return TG_OK;
}

extern "C" int tg_eval_buffers(tg_engine* e, float** planes, float** policy, float** value, int32_t* slot_cap)
{
    if (!e) return fail(TG_ERR_ARG, "null engine");
    if (planes) *planes = e->D.planes;
    if (policy) *policy = e->D.policy;
    if (value) *value = e->D.value;
    if (slot_cap) *slot_cap = e->slot_cap;
    return TG_OK;
}

extern "C" int tg_forward_device(tg_engine* e, int32_t n, int32_t use_logit)
{
    if (!e || n < 0 || n > e->slot_cap) return fail(TG_ERR_ARG, "n outside [0, slot_cap]");
    CK(cudaSetDevice(e->cfg.device));
    // the slot count travels as a launch argument: back-to-back calls need no synchronisation between them
    int rc = 0;
    DISPATCH_N(e, rc = launch_net<BN>(e, use_logit, n, n));
    return rc;
}

// Cross-stream ordering for the zero-copy buffers: the engine's stream is non-blocking, so work a caller queued on ITS
// stream (e.g. torch filling `planes`) is ordered with the engine's only through these two calls.
extern "C" int tg_stream_wait(tg_engine* e, void* caller_stream)      // engine work queued after this call waits for the caller's stream
{
    if (!e) return fail(TG_ERR_ARG, "null engine");
    CK(cudaSetDevice(e->cfg.device));
    CK(cudaEventRecord(e->ev_x, (cudaStream_t)caller_stream));
    CK(cudaStreamWaitEvent(e->stream, e->ev_x, 0));
    return TG_OK;
}
extern "C" int tg_stream_signal(tg_engine* e, void* caller_stream)    // caller work queued after this call waits for the engine's stream
{
    if (!e) return fail(TG_ERR_ARG, "null engine");
    CK(cudaSetDevice(e->cfg.device));
    CK(cudaEventRecord(e->ev_x, e->stream));
    CK(cudaStreamWaitEvent((cudaStream_t)caller_stream, e->ev_x, 0));
    return TG_OK;
}

extern "C" void* tg_stream(tg_engine* e) { return e ? (void*)e->stream : nullptr; }

extern "C" int tg_sync(tg_engine* e)
{
    if (!e) return fail(TG_ERR_ARG, "null engine");
    CK(cudaSetDevice(e->cfg.device));
    CK(cudaStreamSynchronize(e->stream));
    return check_net_overflow(e);
}

extern "C" int tg_forward(tg_engine* e, const float* planes, int32_t n, int32_t use_logit, float* policy, float* value)
{
    if (!e || !planes || !policy || !value || n < 0) return fail(TG_ERR_ARG, "bad argument");
    CK(cudaSetDevice(e->cfg.device));
    const Dev& D = e->D;
    CK(cudaEventRecord(e->events[0], e->stream));
    for (int s0 = 0; s0 < n; s0 += e->slot_cap) {
        const int ns = std::min(e->slot_cap, n - s0);
        CK(cudaMemcpyAsync(D.planes, planes + (size_t)s0 * e->PLANES, (size_t)ns * e->PLANES * 4, cudaMemcpyHostToDevice, e->stream));
        int rc = 0;
        DISPATCH_N(e, rc = launch_net<BN>(e, use_logit, ns, ns));
        if (rc) return rc;
        CK(cudaMemcpyAsync(policy + (size_t)s0 * e->A, D.policy, (size_t)ns * e->A * 4, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaMemcpyAsync(value + (size_t)s0 * 3, D.value, (size_t)ns * 3 * 4, cudaMemcpyDeviceToHost, e->stream));
    }
    CK(cudaEventRecord(e->events[1], e->stream));
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaEventElapsedTime(&e->last_ms, e->events[0], e->events[1]));
    return check_net_overflow(e);
}

extern "C" int tg_tree_size(tg_engine* e, int32_t game, int32_t* num_nodes)
{
    if (!e || !num_nodes || game < 0 || game >= e->cfg.games) return fail(TG_ERR_ARG, "bad argument");
    CK(cudaSetDevice(e->cfg.device));
    CK(cudaMemcpyAsync(num_nodes, e->D.gs + (size_t)game * GS_STRIDE + GS_NNODES, 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return TG_OK;
}

extern "C" int tg_read_node(tg_engine* e, int32_t game, int32_t index, tg_node_view* o)
{
    if (!e || !o || game < 0 || game >= e->cfg.games || index < 0 || index >= e->D.tree.max_nodes) return fail(TG_ERR_ARG, "bad argument");
    CK(cudaSetDevice(e->cfg.device));
    const TreePool& t = e->D.tree;
    const size_t nb = (size_t)game * t.max_nodes + index, row = nb * e->AP;
    int hdr[H_STRIDE];
    CK(cudaMemcpyAsync(hdr, t.hdr + nb * H_STRIDE, sizeof hdr, cudaMemcpyDeviceToHost, e->stream));
#define RD(dst, src, T) if (o->dst) CK(cudaMemcpyAsync(o->dst, src + row, (size_t)e->AP * sizeof(T), cudaMemcpyDeviceToHost, e->stream))
    RD(action, t.action, int16_t); RD(children_index, t.cidx, int); RD(children_value, t.cval, float); RD(children_visits, t.cvis, int);
    RD(children_policy, t.cpol, double); RD(children_virtual_loss, t.cvl, int); RD(children_value_sum, t.cvsum, float);
#undef RD
    if (o->noise) {                                             // MCTSNode.noise: the Gumbel draw lives on the root only (node.py:57, 275-278)
        if (index == 0) CK(cudaMemcpyAsync(o->noise, t.noise + (size_t)game * e->AP, (size_t)e->AP * 8, cudaMemcpyDeviceToHost, e->stream));
        else memset(o->noise, 0, (size_t)e->AP * 8);
    }
    CK(cudaStreamSynchronize(e->stream));
    o->num_children = hdr[H_K]; o->node_visits = hdr[H_NV]; o->virtual_loss = hdr[H_VL];
    memcpy(&o->node_value_sum, &hdr[H_VSUM], 4); memcpy(&o->raw_value, &hdr[H_RAW], 4);
    return TG_OK;
}

// ---------------------------------------------------------------------------------------------
// Training samples from the record ring
// ---------------------------------------------------------------------------------------------
extern "C" int64_t tg_emit_samples(tg_engine* e, const int32_t* games_list, int32_t n, const int32_t* plies, const int32_t* syms)
{
    if (!e || n < 0 || (n > 0 && (!games_list || !plies || !syms))) return fail(TG_ERR_ARG, "bad argument");
    if (!e->sample_cap) return fail(TG_ERR_STATE, "engine was created without sample_cap");
    if (e->step_pending) return fail(TG_ERR_STATE, "collect the step in flight first");
    if (n == 0) return e->sample_count;
    CK(cudaSetDevice(e->cfg.device));
    std::vector<int> host((size_t)n * 18);                       // [games n | out_base n | plies 8n | syms 8n]
    int64_t total = e->sample_count;
    for (int i = 0; i < n; i++) {
        const int g = games_list[i];
        if (g < 0 || g >= e->cfg.games) return fail(TG_ERR_ARG, "game index out of range");
        const int* gs = e->h_gs + (size_t)g * GS_STRIDE;         // state block of the last collected step
        if (!gs[GS_FINISHED]) return fail(TG_ERR_STATE, "tg_emit_samples needs finished games (their slots are used as replay scratch)");
        host[i] = g; host[(size_t)n + i] = (int)total;
        int prev = -1, cnt = 0;
        for (int j = 0; j < 8; j++) {
            const int p = plies[(size_t)i * 8 + j], sy = syms[(size_t)i * 8 + j];
            host[(size_t)2 * n + (size_t)i * 8 + j] = p; host[(size_t)10 * n + (size_t)i * 8 + j] = sy;
            if (p < 0) continue;
            if (p <= prev || p >= gs[GS_NMOVES] || cnt != j || sy < 0 || sy > 7) return fail(TG_ERR_ARG, "plies must be ascending move indices of the game, -1 padded at the end; syms in 0..7");
            prev = p; cnt++;
        }
        total += cnt;
    }
    if (total > e->sample_cap) return fail(TG_ERR_STATE, "sample buffer full: read and clear it (tg_samples_read / tg_samples_clear)");
    if (host.size() > e->d_emit_cap) {
        CK(cudaStreamSynchronize(e->stream));
        if (e->d_emit) cudaFree(e->d_emit);
        e->d_emit = nullptr; e->d_emit_cap = 0;
        CK(cudaMalloc(&e->d_emit, host.size() * 2 * sizeof(int)));
        e->d_emit_cap = host.size() * 2;
    }
    CK(cudaMemcpyAsync(e->d_emit, host.data(), host.size() * sizeof(int), cudaMemcpyHostToDevice, e->stream));   // pageable source: staged before the call returns
    const int* d = e->d_emit;
    DISPATCH_N(e, (k_emit_samples<BN><<<(n + SEARCH_WARPS - 1) / SEARCH_WARPS, SEARCH_WARPS * 32, search_smem(BN), e->stream>>>(
        e->D, d, n, d + (size_t)2 * n, d + (size_t)10 * n, d + n)));
    e->launches++;
    CK(cudaGetLastError());
    e->sample_count = total;
    return total;
}

extern "C" int tg_sample_buffers(tg_engine* e, float** input, double** policy, int32_t** value, int64_t* count, int64_t* cap)
{
    if (!e) return fail(TG_ERR_ARG, "null engine");
    if (input) *input = e->D.smp_input;
    if (policy) *policy = e->D.smp_policy;
    if (value) *value = e->D.smp_value;
    if (count) *count = e->sample_count;
    if (cap) *cap = e->sample_cap;
    return TG_OK;
}

double tg_round_3e(double x);                                    // tg_record.cpp: float(f"{x:.3e}")

extern "C" int tg_samples_read(tg_engine* e, int64_t first, int64_t n, float* input, double* policy, int32_t* value, int32_t round_like_sgf)
{
    if (!e || first < 0 || n < 0 || first + n > e->sample_count) return fail(TG_ERR_ARG, "sample range outside [0, count)");
    CK(cudaSetDevice(e->cfg.device));
    if (input) CK(cudaMemcpyAsync(input, e->D.smp_input + (size_t)first * e->PLANES, (size_t)n * e->PLANES * 4, cudaMemcpyDeviceToHost, e->stream));
    if (policy) CK(cudaMemcpyAsync(policy, e->D.smp_policy + (size_t)first * e->A, (size_t)n * e->A * 8, cudaMemcpyDeviceToHost, e->stream));
    if (value) CK(cudaMemcpyAsync(value, e->D.smp_value + first, (size_t)n * 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    if (policy && round_like_sgf) {
        // the reference's targets went through the SGF comment: four significant digits ("%.3e", selfplay_record.py:61) read
        // back with float() (feature.py:96); unlisted points hold the literal 1e-18
        const int64_t tot = n * e->A;
        const int chunks = (int)std::min<int64_t>(64, std::max<int64_t>(1, tot / 4096));
        parallel_for(chunks, [&](int c) {
            const int64_t a = tot * c / chunks, b = tot * (c + 1) / chunks;
            for (int64_t i = a; i < b; i++) if (policy[i] != 1e-18) policy[i] = tg_round_3e(policy[i]);
        });
    }
    return TG_OK;
}

extern "C" int tg_samples_clear(tg_engine* e)
{
    if (!e) return fail(TG_ERR_ARG, "null engine");
    e->sample_count = 0;
    return TG_OK;
}

extern "C" int64_t tg_launch_count(tg_engine* e) { return e ? e->launches : 0; }
extern "C" float tg_last_device_ms(tg_engine* e) { return e ? e->last_ms : 0.f; }

// Time one kernel class on the evaluator batch as it stands (bench.py roofline legs).
extern "C" int tg_bench_kernel(tg_engine* e, const char* name, int32_t slots, int32_t iters, float* ms_out)
{
    if (!e || !name || !ms_out || iters < 1) return fail(TG_ERR_ARG, "bad argument");
    CK(cudaSetDevice(e->cfg.device));
    const std::string k(name);
    slots = std::min(slots, e->slot_cap);
    if (k == "eval_ms") { *ms_out = e->last_eval_ms; return TG_OK; }
    if (k == "eval_slots") { *ms_out = (float)e->last_eval_slots; return TG_OK; }
    if (k == "dualnet") {
        int rc = 0;
        DISPATCH_N(e, rc = launch_net<BN>(e, 1, slots, slots));     // warm-up
        if (rc) return rc;
        CK(cudaEventRecord(e->events[0], e->stream));
        for (int i = 0; i < iters && !rc; i++) DISPATCH_N(e, rc = launch_net<BN>(e, 1, slots, slots));
        if (rc) return rc;
        CK(cudaEventRecord(e->events[1], e->stream));
        CK(cudaStreamSynchronize(e->stream));
        CK(cudaEventElapsedTime(ms_out, e->events[0], e->events[1]));
        *ms_out /= (float)iters;
        return TG_OK;
    }
    if (k == "dualnet_dbg") {
        // development probe: per-layer clock64 stamps of CTA 0 (MMA start, MMA issued, epilogue start, epilogue end)
        long long* d = nullptr;
        CK(cudaMalloc(&d, 64 * 8));
        CK(cudaMemset(d, 0, 64 * 8));
        NetDev saved = e->net;
        e->net.dbg = d;
        int rc = 0;
        DISPATCH_N(e, rc = launch_net<BN>(e, 1, slots, slots));
        e->net = saved;
        if (rc) return rc;
        CK(cudaStreamSynchronize(e->stream));
        long long h[64];
        CK(cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost));
        cudaFree(d);
        for (int l = 0; l < 14; l++)
            fprintf(stderr, "layer %2d: mma_issue %7lld  mma_start->epi0_start %7lld  epi0_start->epi3_end %7lld  epi3_end->next_mma_start %7lld\n", l,
                    h[l * 4 + 1] - h[l * 4 + 0], h[l * 4 + 2] - h[l * 4 + 0], h[l * 4 + 3] - h[l * 4 + 2],
                    l < 13 ? h[(l + 1) * 4 + 0] - h[l * 4 + 3] : 0LL);
        *ms_out = 0.f;
        return TG_OK;
    }
    if (k == "planes") {
        // every game re-expands its current leaf snapshots (whatever the last search phase left in the queue)
        CK(cudaEventRecord(e->events[0], e->stream));
        for (int i = 0; i < iters; i++) DISPATCH_N(e, (k_planes<BN><<<e->D.games, 256, 16 * Snap<BN>::BYTES, e->stream>>>(e->D)));
        e->launches += iters;
        CK(cudaEventRecord(e->events[1], e->stream));
        CK(cudaStreamSynchronize(e->stream));
        CK(cudaEventElapsedTime(ms_out, e->events[0], e->events[1]));
        *ms_out /= (float)iters;
        return TG_OK;
    }
    return fail(TG_ERR_ARG, "unknown kernel name");
}
