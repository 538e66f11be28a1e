/*
 * tamago_b200.h -- C ABI of the B200-native TamaGo self-play / search engine.
 *
 * The reference (kobanium/TamaGo) is pure Python and has no FFI; its seams for this path are Python
 * classes.  Each entry point below names the reference interface it stands in for (file:line relative to
 * the reference checkout).  A maintainer binds the library with ctypes.
 *
 * Conventions
 *   - plain C types only; every function returns 0 on success or a negative tg_status, and
 *     tg_last_error() gives the message of the last failure on the calling thread;
 *   - one engine = one GPU = one host thread at a time; work is queued on the engine's CUDA stream and the
 *     functions that return data to host buffers synchronise before returning;
 *   - positions are the reference's padded indices pos = x + y*(N+2), x,y in [1,N]; PASS = 0, RESIGN = -1
 *     (board/constant.py:24-26); colours are Stone values 1 = black, 2 = white (board/stone.py:5);
 *   - per-child arrays have the stride tg_action_stride(N) (= N*N+1 rounded up to 32).
 */
#ifndef TAMAGO_B200_H
#define TAMAGO_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tg_engine tg_engine;

enum tg_status { TG_OK = 0, TG_ERR_ARG = -1, TG_ERR_CUDA = -2, TG_ERR_STATE = -3, TG_ERR_SEARCH = -4 };
enum tg_mode { TG_MODE_SH = 0,      /* Gumbel sequential halving: MCTSTree.generate_move_with_sequential_halving, mcts/tree.py:318 */
               TG_MODE_PUCT = 1 };  /* PUCT: MCTSTree.search_best_move, mcts/tree.py:57 */
enum tg_evaluator { TG_EVAL_DUALNET_TC = 0,   /* tcgen05 tensor-core DualNet (product path) */
                    TG_EVAL_DUALNET_FP32 = 1, /* CUDA-core fp32 DualNet (on-device numerical reference) */
                    TG_EVAL_HASHNET = 2,      /* deterministic hash evaluator with dyadic fp32 outputs (tree parity tests) */
                    TG_EVAL_HASHNET2 = 3 };   /* same with non-dyadic outputs (k/1000): fp32 accumulation order is observable */

typedef struct tg_config {
    int32_t board_size;       /* N: 9, 13 or 19 (board/constant.py:4 BOARD_SIZE) */
    float   komi;             /* GoBoard(komi=...), board/go_board.py:20 */
    int32_t superko;          /* GoBoard(check_superko=...) */
    int32_t games;            /* concurrent games (board + tree pools) on this GPU */
    int32_t max_visits;       /* largest visit budget that will be requested (sizes the node pool and leaf queue) */
    int32_t batch_size;       /* PUCT leaves per game per evaluation: MCTSTree(batch_size=...), mcts/tree.py:29 */
    int32_t max_nodes;        /* nodes per game; 0 = max_visits + 2 (MCTSTree(tree_size=...)) */
    int32_t device;           /* CUDA device ordinal */
    int32_t evaluator;        /* tg_evaluator */
    int32_t dedup;            /* 1: evaluate identical leaves of one batch once (result-preserving) */
    int32_t cgos_mode;        /* MCTSTree(cgos_mode=...) */
    int32_t net_blocks;       /* DualNet residual blocks (nn/network/dual_net.py:26), default 6 */
    uint64_t seed;            /* seed of the counter-based Dirichlet/Gumbel stream */
    int32_t record_ring;      /* 1: keep the SelfPlayRecord of every running game on the device (sgf/selfplay_record.py:45-64) */
    int32_t sample_cap;       /* > 0: room for that many training samples emitted from the record ring (tg_emit_samples) */
    int32_t scoring;          /* final score of finished games: 0 = GoBoard.count_score (go_board.py:561-608, the reference),
                                 1 = Tromp-Taylor area scoring (the adjudication get_final_status.py:15-64 asks GNU Go for) */
} tg_config;

/* DualNet parameters: host fp32 arrays named as in DualNet.state_dict() (nn/network/dual_net.py:14-39).
 * bn arrays are [4][C]: weight, bias, running_mean, running_var. */
typedef struct tg_weights {
    const float* conv_w;            /* conv_layer.weight        [64][6][3][3]  */
    const float* bn;                /* bn_layer.*               [4][64]        */
    float        bn_eps;            /* 1e-5 (dual_net.py:32)                   */
    const float* block_conv_w;      /* blocks.i.conv{1,2}.weight [blocks][2][64][64][3][3] */
    const float* block_bn;          /* blocks.i.bn{1,2}.*        [blocks][2][4][64]        */
    float        block_bn_eps;      /* 2e-5 (res_block.py:23-24)               */
    const float* policy_conv_w;     /* policy_head.conv_layer.weight [2][64]   */
    const float* policy_bn;         /* policy_head.bn_layer.*   [4][2]         */
    const float* policy_fc_w;       /* policy_head.fc_layer.weight [N*N+1][2*N*N] */
    const float* policy_fc_b;       /* policy_head.fc_layer.bias [N*N+1]       */
    const float* value_conv_w;      /* value_head.conv_layer.weight [1][64]    */
    const float* value_bn;          /* value_head.bn_layer.*    [4][1]         */
    const float* value_fc_w;        /* value_head.fc_layer.weight [3][N*N]     */
    const float* value_fc_b;        /* value_head.fc_layer.bias [3]            */
    float        head_bn_eps;       /* 2e-5 (head/policy_head.py:21, head/value_head.py:22) */
} tg_weights;

/* Per-ply state dump of tg_play (parity tests): what GoBoard / StringData expose after put_stone. */
typedef struct tg_ply_dump {
    uint8_t*  color;    /* [games][plies][(N+2)^2]  GoBoard.board                               */
    int16_t*  libs;     /* [games][plies][(N+2)^2]  String.get_num_liberties() of the stone's string */
    int16_t*  size;     /* [games][plies][(N+2)^2]  String.get_size()                            */
    int32_t*  scal;     /* [games][plies][5]        moves, ko_pos, ko_move, prisoner[0], prisoner[1] */
    uint64_t* hash;     /* [games][plies]           positional_hash                              */
    uint8_t*  legal;    /* [games][plies][2][N*N]   is_legal for black, white (go_board.py:260)  */
    int16_t*  satari;   /* [games][plies][2][N*N]   check_self_atari_stone (go_board.py:327), 0 where illegal */
    uint8_t*  eye;      /* [games][plies][2][N*N]   is_complete_eye (go_board.py:367), 0 where illegal */
    uint8_t*  cand;     /* [games][plies][2][N*N]   expansion candidates (mcts/tree.py:260-263)  */
    int32_t*  score;    /* [games][plies]           count_score (go_board.py:561)                */
    int32_t   plies;    /* plies allocated per game */
    int32_t*  tt_score; /* [games][plies]           Tromp-Taylor area score, Black - White (tg_config.scoring = 1); may be NULL */
} tg_ply_dump;

/* Result of one move of every game (host buffers, caller-owned; any pointer may be NULL). */
typedef struct tg_step_result {
    int32_t* move;          /* [games] chosen pos, PASS 0, RESIGN -1, -2 = slot idle                      */
    int32_t* color;         /* [games] colour that moved                                                   */
    int32_t* num_children;  /* [games] root.get_num_children()                                             */
    int16_t* action;        /* [games][stride] root.action (sgf/selfplay_record.py:56-61)                  */
    double*  improved;      /* [games][stride] root.calculate_improved_policy() (mcts/node.py:308)         */
    int32_t* visits;        /* [games][stride] root.children_visits                                        */
    int32_t* finished;      /* [games] 1 when the game ended with this move                                */
    int32_t* winner;        /* [games] Stone value: 1, 2, 3 = draw (OUT_OF_BOARD), 0 = undecided (worker.py:57-87) */
    int32_t* resigned;      /* [games]                                                                     */
    float*   score;         /* [games] count_score() - komi (worker.py:81)                                 */
    int32_t* error;         /* [games] search error bits (0 = ok)                                          */
    int64_t* evals;         /* [2] leaf evaluations requested by the search, evaluations executed (differ with dedup) */
    int32_t* n_moves;       /* [games] root moves played so far in the running game (SGFReader.get_n_moves of its record) */
} tg_step_result;

/* One node of a game's tree (MCTSNode, mcts/node.py:18-39), for tree-parity tests and get_root(). */
typedef struct tg_node_view {
    int32_t num_children, node_visits, virtual_loss;
    float   node_value_sum, raw_value;
    int16_t* action;            /* [stride] */
    int32_t* children_index;    /* [stride] */
    float*   children_value;    /* [stride] */
    int32_t* children_visits;   /* [stride] */
    double*  children_policy;   /* [stride] */
    int32_t* children_virtual_loss; /* [stride] */
    float*   children_value_sum;    /* [stride] */
    double*  noise;             /* [stride] root noise of the game (only filled for node 0) */
} tg_node_view;

const char* tg_last_error(void);
int  tg_action_stride(int board_size);

/* GoBoard pool + MCTSTree pool + DualNet on one GPU. */
int  tg_engine_create(const tg_config* cfg, tg_engine** out);
void tg_engine_destroy(tg_engine* e);

/* board/zobrist_hash.py:9-10 hash_bit_mask [4][(N+2)^2] (drawn unseeded at import by the reference, hence an input) */
int  tg_set_zobrist(tg_engine* e, const uint64_t* table);
/* nn/utility.py:139-159 load_network: parameters of DualNet */
int  tg_load_weights(tg_engine* e, const tg_weights* w);
/* the same from DEVICE pointers (fp32 arrays on the engine's GPU, e.g. the trainer's parameters): the BatchNorm fold and
 * the operand packing run on the device, no host round trip; results are bit-identical to tg_load_weights */
int  tg_load_weights_device(tg_engine* e, const tg_weights* w);

/* GoBoard.clear (go_board.py:111) for the games flagged in mask (NULL = all); game_ids key the noise stream.
 * Asynchronous: the arguments are copied before the call returns, the reset is ordered on the engine's stream. */
int  tg_reset(tg_engine* e, const uint8_t* mask, const uint64_t* game_ids, const uint8_t* never_resign);
/* GoBoard.put_stone (go_board.py:131) for counts[g] moves of every game; colors NULL = alternate from the side to move */
int  tg_play(tg_engine* e, const int16_t* moves, const uint8_t* colors, const int32_t* counts, int32_t stride, tg_ply_dump* dump);
/* side to move of every game (the colour passed to the search) */
int  tg_set_to_move(tg_engine* e, const int32_t* colors);

/* nn/feature.py:10 generate_input_planes(board, color, sym=0) of every root position -> [games][6][N][N] fp32 */
int  tg_planes(tg_engine* e, float* out);
/* DualNet.inference / inference_with_policy_logits (dual_net.py:81-106): planes [n][6][N][N] -> policy [n][N*N+1], value [n][3] */
int  tg_forward(tg_engine* e, const float* planes, int32_t n, int32_t use_logit, float* policy, float* value);

/* Device-resident evaluation (zero copy): the evaluator batch lives in device buffers owned by the engine -- planes
 * [slot_cap][6][N][N], policy [slot_cap][N*N+1], value [slot_cap][3], all fp32 -- which a caller may alias (e.g. as torch
 * tensors through __cuda_array_interface__ / DLPack).  tg_forward_device runs DualNet.inference / inference_with_policy_logits
 * (dual_net.py:81-106) on the first n slots, asynchronously on the engine's stream (tg_stream, a cudaStream_t);
 * tg_sync waits for everything queued on it. */
int   tg_eval_buffers(tg_engine* e, float** planes, float** policy, float** value, int32_t* slot_cap);
int   tg_forward_device(tg_engine* e, int32_t n, int32_t use_logit);
void* tg_stream(tg_engine* e);
int   tg_sync(tg_engine* e);
/* ordering between the engine's (non-blocking) stream and a caller's cudaStream_t, e.g. the torch stream that fills
 * `planes`: tg_stream_wait makes engine work queued afterwards wait for what the caller has queued so far,
 * tg_stream_signal makes caller work queued afterwards wait for the engine */
int   tg_stream_wait(tg_engine* e, void* caller_stream);
int   tg_stream_signal(tg_engine* e, void* caller_stream);

/* MCTSTree.generate_move_with_sequential_halving (tree.py:318) / search_best_move (tree.py:57) for every game.
 * play = 1 additionally runs the body of selfplay_worker's move loop (worker.py:58-87). */
int  tg_genmove(tg_engine* e, int32_t mode, int32_t visits, int32_t strict, int32_t play, tg_step_result* out);
/* The same in two halves: tg_genmove_async queues the move of every game and returns at once (root_arrays = 1 also
 * stages root.action / improved policy / visits for tg_collect); tg_collect waits for it and fills the result.  The
 * host is free in between (format records, write files, prepare the next reset) while the GPU searches. */
int  tg_genmove_async(tg_engine* e, int32_t mode, int32_t visits, int32_t strict, int32_t play, int32_t root_arrays);
int  tg_collect(tg_engine* e, tg_step_result* out);

/* SelfPlayRecord of finished games (sgf/selfplay_record.py:45-110), kept per game on the device (tg_config.record_ring):
 * tg_fetch_records queues the device->host copies of the listed games' records (call it after tg_collect and before the
 * tg_reset that recycles their slots); tg_format_records returns their SGF texts back to back (offsets[n+1]);
 * tg_write_records writes <dir>/<index[i]>.sgf for each (write_record, :67-110) and returns the number of moves written;
 * tg_fetched_record exposes the raw arrays of one fetched game ([n_moves] and [n_moves][stride]). */
int     tg_fetch_records(tg_engine* e, const int32_t* games, int32_t n);
int64_t tg_format_records(tg_engine* e, char* buf, int64_t cap, int64_t* offsets);
int64_t tg_write_records(tg_engine* e, const char* dir, const int64_t* index);
int     tg_fetched_record(tg_engine* e, int32_t i, int32_t* n_moves, int16_t* move, uint8_t* color, int16_t* num_children,
                          int16_t* action, double* improved);

/* MCTSTree.node[index] of one game */
int  tg_tree_size(tg_engine* e, int32_t game, int32_t* num_nodes);
int  tg_read_node(tg_engine* e, int32_t game, int32_t index, tg_node_view* out);

/* sgf/selfplay_record.py:67-110 write_record: SGF text of one finished game from the per-move records.
 * Returns the number of bytes written (excluding the terminator) or a negative status. */
int  tg_format_sgf(int32_t board_size, int32_t n_moves, const int32_t* moves, const int32_t* colors,
                   const int32_t* num_children, const int16_t* action, const double* improved, int32_t stride,
                   int32_t winner, int32_t resigned, double score, double komi, char* out, int32_t out_cap);

/* Training samples straight from the record ring (nn/data_generator.py:89-149 without the SGF round trip): for each listed
 * FINISHED game (after tg_collect, before the tg_reset that recycles its slot) the positions before the plies plies[i][0..8)
 * (ascending, -1 padded) under the symmetries syms[i][0..8) (go_board.py:74-104) are appended to device-resident arrays
 * input [count][6][N][N] f32 (nn/feature.py:10-57), policy [count][N*N+1] f64 (feature.py:80-102: improved policy, 1e-18
 * elsewhere) and value [count] i32 (2 = side to move won, 1 = draw, 0 = lost).  Which plies / symmetries: the caller's draw
 * (the reference uses np.random.permutation, data_generator.py:123-124).  Returns the new sample count.
 * tg_sample_buffers exposes the device arrays (zero copy: train on them or gather them over NCCL), tg_samples_read copies a
 * range to the host -- round_like_sgf = 1 applies float(f"{p:.3e}") to the policy entries, which makes the result bit-equal
 * to what the reference's generator reads back from the SGF comments. */
int64_t tg_emit_samples(tg_engine* e, const int32_t* games, int32_t n, const int32_t* plies, const int32_t* syms);
int     tg_sample_buffers(tg_engine* e, float** input, double** policy, int32_t** value, int64_t* count, int64_t* cap);
int     tg_samples_read(tg_engine* e, int64_t first, int64_t n, float* input, double* policy, int32_t* value, int32_t round_like_sgf);
int     tg_samples_clear(tg_engine* e);
/* float(f"{p:.3e}") in place (sgf/selfplay_record.py:61 -> nn/feature.py:96), entries equal to 1e-18 untouched */
void    tg_round_policy(double* p, int64_t n);

/* instrumentation: kernels launched by this engine so far, and device milliseconds of the last tg_genmove */
int64_t tg_launch_count(tg_engine* e);
float   tg_last_device_ms(tg_engine* e);
/* device pointers of the evaluator batch (planes in, policy/value out) for benchmarks that time single kernels */
int  tg_bench_kernel(tg_engine* e, const char* name, int32_t slots, int32_t iters, float* ms_out);

#ifdef __cplusplus
}
#endif
#endif
