/*
 * tg_oracle.h -- CPU oracle for the TamaGo self-play hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference
 * algorithm (kobanium/TamaGo, Python) used as the checker in tests/, in
 * __graft_entry__.smoke() and as bench.py's cpu_baseline / --impl reference
 * arm.  Nothing under tamago_b200/ may include, link or call it.
 *
 * Parity pin: the reference ships no tests or golden vectors,
 * so the oracle is pinned against outputs of the reference itself, generated
 * in the build container by tests/golden/make_golden.py (imports
 * /root/reference) and committed under tests/golden/.
 *
 * Each function cites the reference file:line it follows.
 */
#ifndef TG_ORACLE_H
#define TG_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TGO_MAX_N        19
#define TGO_MAX_CELLS    441          /* (19+2)^2 */
#define TGO_MAX_ACTIONS  362          /* 19*19+1  */
#define TGO_MAX_RECORDS  1083         /* 3*19*19  */

enum { TGO_EMPTY = 0, TGO_BLACK = 1, TGO_WHITE = 2, TGO_OB = 3 };
#define TGO_PASS     0
#define TGO_RESIGN (-1)
#define TGO_NOT_EXPANDED (-1)

/* ---- board (board/go_board.py, board/string.py, board/pattern.py,
 *      board/record.py, board/zobrist_hash.py) ------------------------- */
typedef struct TgoBoard {
    int n, w, cells, max_records;
    int superko;
    double komi;
    uint8_t  color[TGO_MAX_CELLS];
    int      moves;                    /* 1-based number of the next move  */
    int      ko_pos, ko_move;
    int      prisoner[2];
    uint64_t hash;
    uint64_t hist_hash[TGO_MAX_RECORDS];
    int16_t  hist_pos[TGO_MAX_RECORDS];
    uint8_t  hist_color[TGO_MAX_RECORDS];
    /* derived, rebuilt after every stone placement */
    int16_t  chain[TGO_MAX_CELLS];     /* chain label (min pos) or -1      */
    int16_t  libs[TGO_MAX_CELLS];      /* by label                         */
    int16_t  size[TGO_MAX_CELLS];      /* by label                         */
    const uint64_t *zob;               /* [4][cells], caller-owned         */
} TgoBoard;

void     tgo_default_zobrist(int n, uint64_t seed, uint64_t *out /*[4*cells]*/);
void     tgo_board_init(TgoBoard *b, int n, double komi, int superko, const uint64_t *zob);
void     tgo_board_clear(TgoBoard *b);
void     tgo_board_copy(TgoBoard *dst, const TgoBoard *src);
void     tgo_put_stone(TgoBoard *b, int pos, int color);
int      tgo_is_legal(const TgoBoard *b, int pos, int color);
int      tgo_self_atari(const TgoBoard *b, int pos, int color);
int      tgo_complete_eye(const TgoBoard *b, int pos, int color);
int      tgo_eye_color(const TgoBoard *b, int pos);
int      tgo_candidates(const TgoBoard *b, int color, int16_t *out);   /* PASS last; returns k */
void     tgo_analyze(const TgoBoard *b, int color, uint8_t *legal, int16_t *satari,
                     uint8_t *eye, uint8_t *cand);
void     tgo_planes(const TgoBoard *b, int color, float *out /*[6*n*n]*/);
int      tgo_count_score(const TgoBoard *b);
int      tgo_onboard_pos(const TgoBoard *b, int idx);
/* flat state export for tests: libs/size per point (0 for empty) */
void     tgo_export_state(const TgoBoard *b, uint8_t *color, int16_t *libs_pt,
                          int16_t *size_pt, int32_t *scalars /*[6]: moves,ko_pos,ko_move,pr0,pr1,0*/,
                          uint64_t *hash);
const uint8_t *tgo_eye_table(void);     /* 65536 entries */

/* ---- deterministic noise + math (ours; see DESIGN.md "noise") -------- */
double   tgo_det_log(double x);
double   tgo_det_exp(double x);
double   tgo_noise_u(uint64_t seed, uint64_t game, uint32_t move, uint32_t node,
                     uint32_t tag, uint32_t idx);
void     tgo_dirichlet(uint64_t seed, uint64_t game, uint32_t move, uint32_t node,
                       int k, double *out);
void     tgo_gumbel(uint64_t seed, uint64_t game, uint32_t move, int count, double *out);
double   tgo_np_sum(const double *a, int n);      /* numpy pairwise order */
void     tgo_softmax(const double *logits, int n, double *out, int use_libm);

/* ---- search (mcts/tree.py, mcts/node.py, mcts/pucb/pucb.py,
 *      mcts/sequential_halving.py) ------------------------------------- */
typedef struct TgoNode {
    int    num_children;
    int    node_visits, virtual_loss;
    float  node_value_sum;             /* fp32-accumulated (torch scalar)  */
    float  raw_value;
    int16_t action[TGO_MAX_ACTIONS];
    int32_t children_index[TGO_MAX_ACTIONS];
    float   children_value[TGO_MAX_ACTIONS];
    int32_t children_visits[TGO_MAX_ACTIONS];
    double  children_policy[TGO_MAX_ACTIONS];
    int32_t children_virtual_loss[TGO_MAX_ACTIONS];
    float   children_value_sum[TGO_MAX_ACTIONS];   /* fp32-accumulated     */
    double  noise[TGO_MAX_ACTIONS];
} TgoNode;

/* evaluator: planes [nb][6][n][n] fp32 -> policy [nb][n*n+1] (logits when
 * use_logit else softmax probabilities), value [nb][3] softmax probs. */
typedef void (*tgo_eval_fn)(void *ctx, const float *planes, int nb, int use_logit,
                            float *policy, float *value);

typedef struct TgoTree {
    int n, max_actions;
    TgoNode *node; int tree_size, num_nodes, current_root;
    int batch_size, cgos_mode;
    tgo_eval_fn eval; void *eval_ctx;
    /* noise keys */
    uint64_t seed, game; uint32_t move;
    int use_libm;                       /* 1: libm exp (closest to numpy)  */
    /* queue */
    float *q_planes; int *q_path_node; int *q_path_child; int *q_path_len; int *q_node_index;
    int q_len, q_cap, max_depth;
    long evals;                         /* NN evaluations executed          */
    int error;
} TgoTree;

TgoTree *tgo_tree_new(int n, int tree_size, int batch_size, int cgos_mode,
                      tgo_eval_fn eval, void *ctx);
void     tgo_tree_free(TgoTree *t);
void     tgo_tree_set_noise_key(TgoTree *t, uint64_t seed, uint64_t game, uint32_t move);
int      tgo_sh_schedule(int m, int visits, int *considered, int *counts, int cap);
int      tgo_genmove_sh(TgoTree *t, const TgoBoard *b, int color, int visits, int never_resign);
int      tgo_genmove_puct(TgoTree *t, const TgoBoard *b, int color, int visits, int strict);
void     tgo_improved_policy(const TgoTree *t, int node_index, double *out);
void     tgo_tree_set_use_libm(TgoTree *t, int v);
long     tgo_tree_evals(const TgoTree *t);
int      tgo_tree_error(const TgoTree *t);
const TgoNode *tgo_tree_node(const TgoTree *t, int idx);
int      tgo_tree_num_nodes(const TgoTree *t);
int      tgo_sizeof_node(void);
int      tgo_sizeof_board(void);
int      tgo_sizeof_record(void);

/* ---- self-play game loop (selfplay/worker.py:46-90) ------------------- */
typedef struct TgoGameRecord {
    int n_moves;
    int winner;                 /* TGO_BLACK / TGO_WHITE / TGO_OB(draw) / TGO_EMPTY(unset) */
    int is_resign;
    double score;
    int16_t pos[2 * 19 * 19 + 2];
    uint8_t color[2 * 19 * 19 + 2];
    int16_t num_children[2 * 19 * 19 + 2];
} TgoGameRecord;

/* improved: caller buffer [max_moves][max_actions] doubles, actions likewise int16 */
int      tgo_selfplay_game(TgoTree *t, int n, double komi, const uint64_t *zob,
                           uint64_t seed, uint64_t game, int visits, int never_resign,
                           int use_puct, TgoGameRecord *rec, double *improved, int16_t *actions);

/* ---- native hash evaluators, bulk corpus digests, Tromp-Taylor (ours) -- */
typedef struct TgoHashNetCtx { int n, variant; } TgoHashNetCtx;   /* variant 0: oracle.hashnet, 1: oracle.hashnet2 */
void     tgo_hashnet_eval(void *ctx, const float *planes, int nb, int use_logit, float *policy, float *value);
tgo_eval_fn tgo_hashnet_fn(void);
uint64_t tgo_ply_digest(const TgoBoard *b);
int      tgo_random_game(int n, const uint64_t *zob, int superko, uint64_t seed, uint64_t game, int max_plies,
                         double p_pass, double p_any_legal, int16_t *moves_out, uint64_t *digest_out);
int      tgo_tromp_taylor(const TgoBoard *b);

#ifdef __cplusplus
}
#endif
#endif
