/*
 * tg_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY, see tg_oracle.h).
 *
 * Plain-C restatement of the TamaGo self-play hot path.  All file:line
 * citations are relative to the reference checkout (kobanium/TamaGo @af99376).
 * Compile with -ffp-contract=off: the float64/float32 arithmetic below must
 * round exactly like numpy / torch scalar arithmetic in the reference.
 */
#include "tg_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>

/* ------------------------------------------------------------------ */
/* geometry helpers (board/go_board.py:32-58)                          */
/* ------------------------------------------------------------------ */
static inline void nbr4(const TgoBoard *b, int pos, int out[4])
{   /* order up, left, right, down: go_board.py:53 */
    out[0] = pos - b->w; out[1] = pos - 1; out[2] = pos + 1; out[3] = pos + b->w;
}
static inline int opp(int c) { return c == TGO_BLACK ? TGO_WHITE : (c == TGO_WHITE ? TGO_BLACK : c); }

int tgo_onboard_pos(const TgoBoard *b, int idx)
{   /* go_board.py:82-86: raster order, x fastest */
    int x = idx % b->n, y = idx / b->n;
    return (x + 1) + (y + 1) * b->w;
}

/* ------------------------------------------------------------------ */
/* Zobrist (board/zobrist_hash.py:9-10 draws an unseeded table; the     */
/* table is therefore an input.  This default generator is ours.)       */
/* ------------------------------------------------------------------ */
static inline uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
void tgo_default_zobrist(int n, uint64_t seed, uint64_t *out)
{
    int cells = (n + 2) * (n + 2);
    for (int i = 0; i < 4 * cells; i++) out[i] = mix64(mix64(seed) + (uint64_t)i);
}

/* ------------------------------------------------------------------ */
/* eye pattern table (board/pattern.py:53-98, 212-296)                  */
/* ------------------------------------------------------------------ */
static uint8_t g_eye[65536];
static int g_eye_ready = 0;

static unsigned p3_rev(unsigned b)  { return (b >> 2) | ((b & 0x3) << 2); }
static unsigned p3_rev3(unsigned b) { return (b >> 4) | (b & 0xC) | ((b & 0x3) << 4); }
static unsigned p3_swap(unsigned p) { return ((p >> 1) & 0x5555) | ((p & 0x5555) << 1); }
static unsigned p3_vmirror(unsigned p)
{ return ((p & 0xfc00) >> 10) | (p & 0x03c0) | ((p & 0x003f) << 10); }
static unsigned p3_hmirror(unsigned p)
{ return (p3_rev3((p & 0xfc00) >> 10) << 10) | (p3_rev((p & 0x03c0) >> 6) << 6) | p3_rev3(p & 0x003f); }
static unsigned p3_rot90(unsigned p)
{
    return ((p & 0x0003) << 10) | ((p & 0x0c0c) << 4) | ((p & 0x3030) >> 4)
         | ((p & 0x00c0) << 6) | ((p & 0x0300) >> 6) | ((p & 0xc000) >> 10);
}
static void build_eye(void)
{
    static const unsigned seeds[20] = {        /* pattern.py:53-78 */
        0x5554, 0x5556, 0x5544, 0x5546, 0x1554, 0x1556, 0x1544, 0x1546,
        0x1564, 0x1146, 0xFD54, 0xFD55, 0xFF74, 0xFF75, 0x5566, 0xFD66,
        0x5965, 0x9955, 0xFD56, 0xFF76 };
    memset(g_eye, TGO_EMPTY, sizeof g_eye);
    g_eye[0x5555] = TGO_BLACK; g_eye[p3_swap(0x5555)] = TGO_WHITE;      /* :85-86 */
    g_eye[0x1144] = TGO_BLACK; g_eye[p3_swap(0x1144)] = TGO_WHITE;      /* :91-92 */
    for (int s = 0; s < 20; s++) {
        unsigned sym[8];
        sym[0] = seeds[s];               sym[1] = p3_vmirror(sym[0]);
        sym[2] = p3_hmirror(sym[0]);     sym[3] = p3_vmirror(sym[2]);
        sym[4] = p3_rot90(sym[0]);       sym[5] = p3_rot90(sym[1]);
        sym[6] = p3_rot90(sym[2]);       sym[7] = p3_rot90(sym[3]);
        for (int k = 0; k < 8; k++) {    /* :94-98, in this exact order */
            g_eye[sym[k] & 0xffff] = TGO_BLACK;
            g_eye[p3_swap(sym[k]) & 0xffff] = TGO_WHITE;
        }
    }
    g_eye_ready = 1;
}
const uint8_t *tgo_eye_table(void) { if (!g_eye_ready) build_eye(); return g_eye; }

/* pat3 of a point rebuilt from the colours of its 8 neighbours
 * (pattern.py:47-50 update_pos order UL,U,UR,L,R,DL,D,DR; 2 bits each). */
static unsigned pat3_at(const TgoBoard *b, int pos)
{
    const int w = b->w;
    const int d[8] = { -w - 1, -w, -w + 1, -1, 1, w - 1, w, w + 1 };
    unsigned p = 0;
    for (int i = 0; i < 8; i++) p |= (unsigned)b->color[pos + d[i]] << (2 * i);
    return p;
}
int tgo_eye_color(const TgoBoard *b, int pos)
{   /* pattern.py:153-162 */
    return tgo_eye_table()[pat3_at(b, pos)];
}
static int nb4_empty(const TgoBoard *b, int pos)
{   /* pattern.py:21-30, 142-151 */
    int n4[4], c = 0; nbr4(b, pos, n4);
    for (int i = 0; i < 4; i++) c += (b->color[n4[i]] == TGO_EMPTY);
    return c;
}

/* ------------------------------------------------------------------ */
/* strings: the reference keeps sorted intrusive lists (string.py);    */
/* every observable (liberty count, size, membership) is a function of  */
/* the stones, so the oracle rebuilds them by flood fill.               */
/* ------------------------------------------------------------------ */
static void rebuild_chains(TgoBoard *b)
{
    int stack[TGO_MAX_CELLS];
    uint8_t libmark[TGO_MAX_CELLS];
    for (int i = 0; i < b->cells; i++) { b->chain[i] = -1; b->libs[i] = 0; b->size[i] = 0; }
    for (int p = 0; p < b->cells; p++) {
        int c = b->color[p];
        if ((c != TGO_BLACK && c != TGO_WHITE) || b->chain[p] >= 0) continue;
        memset(libmark, 0, (size_t)b->cells);
        int sp = 0, size = 0, libs = 0;
        stack[sp++] = p; b->chain[p] = (int16_t)p;
        while (sp) {
            int q = stack[--sp], n4[4]; size++;
            nbr4(b, q, n4);
            for (int i = 0; i < 4; i++) {
                int r = n4[i];
                if (b->color[r] == TGO_EMPTY) { if (!libmark[r]) { libmark[r] = 1; libs++; } }
                else if (b->color[r] == c && b->chain[r] < 0) { b->chain[r] = (int16_t)p; stack[sp++] = r; }
            }
        }
        b->libs[p] = (int16_t)libs; b->size[p] = (int16_t)size;
    }
}
static inline int libs_at(const TgoBoard *b, int pos)
{   /* string.py:355-365; empty / off-board points map to the dummy string 0 (libs 0) */
    int l = b->chain[pos];
    return l < 0 ? 0 : b->libs[l];
}

void tgo_board_clear(TgoBoard *b)
{   /* go_board.py:111-129 */
    b->moves = 1; b->ko_move = 0; b->ko_pos = 0; b->prisoner[0] = b->prisoner[1] = 0; b->hash = 0;
    for (int i = 0; i < b->cells; i++) b->color[i] = TGO_OB;
    for (int y = 1; y <= b->n; y++) for (int x = 1; x <= b->n; x++) b->color[x + y * b->w] = TGO_EMPTY;
    memset(b->hist_hash, 0, sizeof b->hist_hash);
    memset(b->hist_pos, 0, sizeof b->hist_pos);
    memset(b->hist_color, 0, sizeof b->hist_color);
    rebuild_chains(b);
}
void tgo_board_init(TgoBoard *b, int n, double komi, int superko, const uint64_t *zob)
{   /* go_board.py:20-109; MAX_RECORDS = 3*BOARD_SIZE^2 (constant.py:31) */
    memset(b, 0, sizeof *b);
    b->n = n; b->w = n + 2; b->cells = b->w * b->w; b->max_records = 3 * n * n;
    b->komi = komi; b->superko = superko; b->zob = zob;
    tgo_board_clear(b);
}
void tgo_board_copy(TgoBoard *dst, const TgoBoard *src) { memcpy(dst, src, sizeof *dst); } /* go_board.py:611-626 */

static void record_save(TgoBoard *b, int color, int pos)
{   /* record.py:30-44: moves >= MAX_RECORDS are dropped */
    if (b->moves < b->max_records) {
        b->hist_color[b->moves] = (uint8_t)color;
        b->hist_pos[b->moves] = (int16_t)pos;
        b->hist_hash[b->moves] = b->hash;
    }
}

void tgo_put_stone(TgoBoard *b, int pos, int color)
{   /* go_board.py:131-185 */
    if (pos == TGO_PASS) { record_save(b, color, pos); b->moves++; return; }
    const int other = opp(color);
    const uint64_t *zob = b->zob;
    int n4[4]; nbr4(b, pos, n4);

    /* own neighbours are looked up before captures, like `connection` (:155-157) */
    int connection = 0;
    for (int i = 0; i < 4; i++) if (b->color[n4[i]] == color) connection++;

    b->color[pos] = (uint8_t)color;                          /* :145 */
    b->hash ^= zob[color * b->cells + pos];                  /* :147 */
    rebuild_chains(b);

    /* captures: enemy neighbour strings left without liberties (:158-166) */
    int prisoner = 0;
    for (int i = 0; i < 4; i++) {
        int q = n4[i];
        if (b->color[q] != other) continue;
        int l = b->chain[q];
        if (b->libs[l] != 0) continue;
        for (int p = 0; p < b->cells; p++)
            if (b->chain[p] == l && b->color[p] == other) {
                b->color[p] = TGO_EMPTY; b->chain[p] = -1;
                b->hash ^= zob[other * b->cells + p];        /* :165-166 */
                prisoner++;
            }
    }
    if (color == TGO_BLACK) b->prisoner[0] += prisoner;      /* :168-171 */
    else if (color == TGO_WHITE) b->prisoner[1] += prisoner;
    rebuild_chains(b);

    if (connection == 0) {                                   /* :173-177 */
        if (prisoner == 1 && b->libs[b->chain[pos]] == 1) {
            b->ko_move = b->moves;
            for (int i = 0; i < 4; i++) if (b->color[n4[i]] == TGO_EMPTY) b->ko_pos = n4[i];
        }
    }
    record_save(b, color, pos);                              /* :184 */
    b->moves++;
}

static int is_suicide(const TgoBoard *b, int pos, int color)
{   /* go_board.py:237-258 */
    int other = opp(color), n4[4]; nbr4(b, pos, n4);
    for (int i = 0; i < 4; i++) {
        int q = n4[i];
        if (b->color[q] == other && libs_at(b, q) == 1) return 0;
        if (b->color[q] == color && libs_at(b, q) > 1) return 0;
    }
    return 1;
}

int tgo_is_legal(const TgoBoard *b, int pos, int color)
{   /* go_board.py:260-304 */
    if (b->color[pos] != TGO_EMPTY) return 0;
    if (nb4_empty(b, pos) == 0 && is_suicide(b, pos, color)) return 0;
    if (b->ko_pos == pos && b->ko_move == b->moves - 1) return 0;
    if (b->superko && pos != TGO_PASS) {
        int other = opp(color), n4[4]; nbr4(b, pos, n4);
        uint64_t h = b->hash;
        int seen[4], ns = 0;
        for (int i = 0; i < 4; i++) {
            int l = b->chain[n4[i]];
            if (l < 0) continue;                 /* id 0: dummy string, libs 0 (:285-291) */
            int dup = 0; for (int k = 0; k < ns; k++) dup |= (seen[k] == l);
            if (dup) continue;
            seen[ns++] = l;
            if (b->libs[l] == 1)                 /* any colour, XOR-ed with the opponent's keys (:290-292) */
                for (int p = 0; p < b->cells; p++)
                    if (b->chain[p] == l) h ^= b->zob[other * b->cells + p];
        }
        h ^= b->zob[color * b->cells + pos];     /* :294 */
        for (int i = 0; i < b->max_records; i++) /* record.py:54-63 scans every slot */
            if (b->hist_hash[i] == h) return 0;
    }
    return 1;
}

int tgo_self_atari(const TgoBoard *b, int pos, int color)
{   /* go_board.py:327-365 */
    int n4[4]; nbr4(b, pos, n4);
    uint8_t mark[TGO_MAX_CELLS]; memset(mark, 0, (size_t)b->cells);
    int nlib = 0;
    for (int i = 0; i < 4; i++)
        if (b->color[n4[i]] == TGO_EMPTY && !mark[n4[i]]) { mark[n4[i]] = 1; nlib++; }
    if (nlib > 1) return 0;
    int checked[4], nc = 0, size = 0, other = opp(color);
    for (int i = 0; i < 4; i++) {
        int q = n4[i];
        if (b->color[q] == color) {
            int l = b->chain[q], dup = 0;
            for (int k = 0; k < nc; k++) dup |= (checked[k] == l);
            if (dup) continue;
            for (int p = 0; p < b->cells; p++) {       /* liberties of string l */
                if (b->color[p] != TGO_EMPTY || mark[p]) continue;
                int m4[4]; nbr4(b, p, m4);
                for (int k = 0; k < 4; k++)
                    if (b->chain[m4[k]] == l && b->color[m4[k]] == color) { mark[p] = 1; nlib++; break; }
            }
            if (nlib >= 3) return 0;
            size += b->size[l];
            checked[nc++] = l;
        } else if (b->color[q] == other) {
            if (libs_at(b, q) == 1) return 0;
        }
    }
    return size + 1;
}

int tgo_complete_eye(const TgoBoard *b, int pos, int color)
{   /* go_board.py:367-397 */
    if (tgo_eye_color(b, pos) != color) return 0;
    const int w = b->w;
    const int cross[4] = { pos - w - 1, pos - w + 1, pos + w - 1, pos + w + 1 };
    int cnt = 0, edge = 0;
    for (int i = 0; i < 4; i++) {
        int c = b->color[cross[i]];
        if (c == color || c == TGO_OB) cnt++;
        else if (c == TGO_EMPTY && tgo_eye_color(b, cross[i]) == color) cnt++;
        if (c == TGO_OB) edge = 1;
    }
    return (edge && cnt == 4) || (!edge && cnt >= 3);
}

int tgo_candidates(const TgoBoard *b, int color, int16_t *out)
{   /* mcts/tree.py:260-264 */
    int k = 0;
    for (int i = 0; i < b->n * b->n; i++) {
        int pos = tgo_onboard_pos(b, i);
        if (!tgo_is_legal(b, pos, color)) continue;
        if (tgo_self_atari(b, pos, color) >= 7) continue;
        if (tgo_complete_eye(b, pos, color)) continue;
        out[k++] = (int16_t)pos;
    }
    out[k++] = TGO_PASS;
    return k;
}

void tgo_analyze(const TgoBoard *b, int color, uint8_t *legal, int16_t *satari, uint8_t *eye, uint8_t *cand)
{   /* batch form of is_legal / check_self_atari_stone / is_complete_eye over onboard_pos (tests) */
    for (int i = 0; i < b->n * b->n; i++) {
        int pos = tgo_onboard_pos(b, i);
        legal[i] = (uint8_t)tgo_is_legal(b, pos, color);
        satari[i] = legal[i] ? (int16_t)tgo_self_atari(b, pos, color) : 0;
        eye[i] = legal[i] ? (uint8_t)tgo_complete_eye(b, pos, color) : 0;
        cand[i] = (uint8_t)(legal[i] && satari[i] < 7 && !eye[i]);
    }
}

void tgo_planes(const TgoBoard *b, int color, float *out)
{   /* nn/feature.py:10-57 (sym = 0) */
    const int nn = b->n * b->n;
    memset(out, 0, sizeof(float) * 6 * (size_t)nn);
    int prev = b->hist_pos[b->moves - 1];          /* record.get(moves-1): slot 0 is PASS */
    int prev_pass = (b->moves > 1 && prev == TGO_PASS);
    for (int i = 0; i < nn; i++) {
        int pos = tgo_onboard_pos(b, i);
        int d = b->color[pos];
        if (color == TGO_WHITE && d != 0) d = 3 - d;           /* :24-25 */
        out[d * nn + i] = 1.0f;                                /* :31 */
        if (!prev_pass && prev == pos) out[3 * nn + i] = 1.0f; /* :43-46 */
        if (prev_pass) out[4 * nn + i] = 1.0f;                 /* :39-41 */
        out[5 * nn + i] = (color == TGO_WHITE) ? -1.0f : 1.0f; /* :50-52 */
    }
}

int tgo_count_score(const TgoBoard *b)
{   /* go_board.py:561-608, including its non-flood-fill behaviour */
    uint8_t bd[TGO_MAX_CELLS];
    memcpy(bd, b->color, (size_t)b->cells);
    const int nn = b->n * b->n;
    for (int i = 0; i < nn; i++) {                 /* :570-573 */
        int pos = tgo_onboard_pos(b, i);
        if ((b->color[pos] == TGO_BLACK || b->color[pos] == TGO_WHITE) && libs_at(b, pos) == 1)
            bd[pos] = TGO_EMPTY;
    }
    for (int i = 0; i < nn; i++) {                 /* :579-603 */
        int pos = tgo_onboard_pos(b, i);
        if (bd[pos] != TGO_EMPTY) continue;
        int n4[4], col = TGO_EMPTY; nbr4(b, pos, n4);
        for (int k = 0; k < 4; k++) {
            int c = bd[n4[k]];
            if (c == TGO_BLACK || c == TGO_WHITE) {
                if (col == TGO_EMPTY) col = c;
                else if (col != c) col = TGO_OB;
            }
        }
        bd[pos] = (uint8_t)col;                    /* :601 writes board[pos] */
    }
    int black = 0, white = 0;
    for (int p = 0; p < b->cells; p++) { black += (bd[p] == TGO_BLACK); white += (bd[p] == TGO_WHITE); }
    return black - white;
}

void tgo_export_state(const TgoBoard *b, uint8_t *color, int16_t *libs_pt, int16_t *size_pt,
                      int32_t *sc, uint64_t *hash)
{
    for (int p = 0; p < b->cells; p++) {
        color[p] = b->color[p];
        int l = b->chain[p];
        libs_pt[p] = l < 0 ? 0 : b->libs[l];
        size_pt[p] = l < 0 ? 0 : b->size[l];
    }
    sc[0] = b->moves; sc[1] = b->ko_pos; sc[2] = b->ko_move;
    sc[3] = b->prisoner[0]; sc[4] = b->prisoner[1]; sc[5] = 0;
    *hash = b->hash;
}

/* ------------------------------------------------------------------ */
/* deterministic math + noise (ours: numpy's MT19937 Dirichlet/Gumbel   */
/* streams cannot be matched on a device).  Every      */
/* operation is a single IEEE-754 binary64 op, so the CUDA side can     */
/* reproduce it bit for bit with __dadd_rn/__dmul_rn/__ddiv_rn.         */
/* ------------------------------------------------------------------ */
static inline double u64_as_double(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
static inline uint64_t double_as_u64(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }

#define TG_LN2_HI 6.93147180369123816490e-01
#define TG_LN2_LO 1.90821492927058770002e-10
#define TG_INV_LN2 1.44269504088896338700e+00
#define TG_SQRT2  1.41421356237309514547e+00

double tgo_det_log(double x)
{
    uint64_t bits = double_as_u64(x);
    int k = (int)((bits >> 52) & 0x7ff) - 1023;
    double m = u64_as_double((bits & 0x000FFFFFFFFFFFFFULL) | 0x3FF0000000000000ULL);
    if (m > TG_SQRT2) { m = m * 0.5; k += 1; }
    double f = m - 1.0;
    double s = f / (2.0 + f);
    double z = s * s;
    double p = 1.0 / 23.0;
    p = p * z + 1.0 / 21.0; p = p * z + 1.0 / 19.0; p = p * z + 1.0 / 17.0;
    p = p * z + 1.0 / 15.0; p = p * z + 1.0 / 13.0; p = p * z + 1.0 / 11.0;
    p = p * z + 1.0 / 9.0;  p = p * z + 1.0 / 7.0;  p = p * z + 1.0 / 5.0;
    p = p * z + 1.0 / 3.0;  p = p * z + 1.0;
    double logm = (2.0 * s) * p;
    double dk = (double)k;
    return dk * TG_LN2_HI + (dk * TG_LN2_LO + logm);
}

double tgo_det_exp(double x)
{
    if (x < -708.0) return 0.0;
    if (x > 709.0) return u64_as_double(0x7FF0000000000000ULL);
    double t = x * TG_INV_LN2;
    int n = (int)(t < 0.0 ? t - 0.5 : t + 0.5);
    double dn = (double)n;
    double r = x - dn * TG_LN2_HI;
    r = r - dn * TG_LN2_LO;
    double p = 1.0 / 6227020800.0;
    p = p * r + 1.0 / 479001600.0; p = p * r + 1.0 / 39916800.0; p = p * r + 1.0 / 3628800.0;
    p = p * r + 1.0 / 362880.0;    p = p * r + 1.0 / 40320.0;    p = p * r + 1.0 / 5040.0;
    p = p * r + 1.0 / 720.0;       p = p * r + 1.0 / 120.0;      p = p * r + 1.0 / 24.0;
    p = p * r + 1.0 / 6.0;         p = p * r + 0.5;              p = p * r + 1.0;
    p = p * r + 1.0;
    return p * u64_as_double((uint64_t)(n + 1023) << 52);
}

double tgo_noise_u(uint64_t seed, uint64_t game, uint32_t move, uint32_t node, uint32_t tag, uint32_t idx)
{
    uint64_t h = mix64(seed + game);
    h = mix64(h + move);
    h = mix64(h + ((uint64_t)node * 4u + tag));
    h = mix64(h + idx);
    return (double)(((h >> 12) << 1) | 1ULL) * 1.1102230246251565404e-16;   /* (2m+1) * 2^-53 */
}

/* warp-shaped sum: 32 strided partials, then an xor-butterfly (16,8,4,2,1) */
static double warp_sum(const double *a, int n)
{
    double part[32], tmp[32];
    for (int l = 0; l < 32; l++) { double s = 0.0; for (int i = l; i < n; i += 32) s = s + a[i]; part[l] = s; }
    for (int off = 16; off >= 1; off >>= 1) {
        for (int l = 0; l < 32; l++) tmp[l] = part[l] + part[l ^ off];
        memcpy(part, tmp, sizeof part);
    }
    return part[0];
}

void tgo_dirichlet(uint64_t seed, uint64_t game, uint32_t move, uint32_t node, int k, double *out)
{   /* stands in for np.random.dirichlet(ones(k)), mcts/tree.py:518 */
    for (int i = 0; i < k; i++) out[i] = 0.0 - tgo_det_log(tgo_noise_u(seed, game, move, node, 0, (uint32_t)i));
    double s = warp_sum(out, k);
    for (int i = 0; i < k; i++) out[i] = out[i] / s;
}
void tgo_gumbel(uint64_t seed, uint64_t game, uint32_t move, int count, double *out)
{   /* stands in for np.random.gumbel(size=MAX_ACTIONS), mcts/node.py:278 */
    for (int i = 0; i < count; i++) {
        double e = 0.0 - tgo_det_log(tgo_noise_u(seed, game, move, 0, 1, (uint32_t)i));
        out[i] = 0.0 - tgo_det_log(e);
    }
}

double tgo_np_sum(const double *a, int n)
{   /* numpy pairwise_sum (float64 add.reduce), verified against np.sum in tests */
    if (n < 8) { double r = 0.0; for (int i = 0; i < n; i++) r = r + a[i]; return r; }
    if (n <= 128) {
        double r[8]; int i;
        for (int j = 0; j < 8; j++) r[j] = a[j];
        for (i = 8; i < n - (n % 8); i += 8) for (int j = 0; j < 8; j++) r[j] = r[j] + a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res = res + a[i];
        return res;
    }
    int n2 = n / 2; n2 -= n2 % 8;
    return tgo_np_sum(a, n2) + tgo_np_sum(a + n2, n - n2);
}

void tgo_softmax(const double *logits, int n, double *out, int use_libm)
{   /* nn/utility.py:125-136 */
    double mx = logits[0];
    for (int i = 1; i < n; i++) if (logits[i] > mx) mx = logits[i];
    for (int i = 0; i < n; i++) { double d = logits[i] - mx; out[i] = use_libm ? exp(d) : tgo_det_exp(d); }
    double s = tgo_np_sum(out, n);
    for (int i = 0; i < n; i++) out[i] = out[i] / s;
}

/* ------------------------------------------------------------------ */
/* search                                                               */
/* ------------------------------------------------------------------ */
#define C_VISIT 50
#define C_SCALE 1.0
#define MAX_CONSIDERED 16
#define PLAYOUTS 100
#define RESIGN_THRESHOLD 0.05

int tgo_sizeof_node(void) { return (int)sizeof(TgoNode); }
int tgo_sizeof_board(void) { return (int)sizeof(TgoBoard); }
int tgo_sizeof_record(void) { return (int)sizeof(TgoGameRecord); }

TgoTree *tgo_tree_new(int n, int tree_size, int batch_size, int cgos_mode, tgo_eval_fn eval, void *ctx)
{   /* mcts/tree.py:29-46 */
    TgoTree *t = (TgoTree *)calloc(1, sizeof *t);
    t->n = n; t->max_actions = n * n + 1;
    t->tree_size = tree_size;
    t->node = (TgoNode *)calloc((size_t)tree_size, sizeof(TgoNode));
    t->batch_size = batch_size < 1 ? 1 : batch_size; t->cgos_mode = cgos_mode;
    t->eval = eval; t->eval_ctx = ctx;
    t->max_depth = 3 * n * n + 8;
    t->q_cap = 0; t->use_libm = 0;
    return t;
}
void tgo_tree_free(TgoTree *t)
{
    if (!t) return;
    free(t->node); free(t->q_planes); free(t->q_path_node); free(t->q_path_child);
    free(t->q_path_len); free(t->q_node_index); free(t);
}
void tgo_tree_set_noise_key(TgoTree *t, uint64_t seed, uint64_t game, uint32_t move)
{ t->seed = seed; t->game = game; t->move = move; }
void tgo_tree_set_use_libm(TgoTree *t, int v) { t->use_libm = v; }
long tgo_tree_evals(const TgoTree *t) { return t->evals; }
int tgo_tree_error(const TgoTree *t) { return t->error; }
const TgoNode *tgo_tree_node(const TgoTree *t, int idx) { return &t->node[idx]; }
int tgo_tree_num_nodes(const TgoTree *t) { return t->num_nodes; }

static void q_reserve(TgoTree *t, int cap)
{
    if (cap <= t->q_cap) return;
    int pl = 6 * t->n * t->n;
    t->q_planes = (float *)realloc(t->q_planes, sizeof(float) * (size_t)cap * pl);
    t->q_path_node = (int *)realloc(t->q_path_node, sizeof(int) * (size_t)cap * t->max_depth);
    t->q_path_child = (int *)realloc(t->q_path_child, sizeof(int) * (size_t)cap * t->max_depth);
    t->q_path_len = (int *)realloc(t->q_path_len, sizeof(int) * (size_t)cap);
    t->q_node_index = (int *)realloc(t->q_node_index, sizeof(int) * (size_t)cap);
    t->q_cap = cap;
}
static void q_push(TgoTree *t, const TgoBoard *b, int color, const int *pn, const int *pc, int plen, int node_index)
{   /* mcts/batch_data.py:18-27 */
    if (t->q_len >= t->q_cap) q_reserve(t, t->q_cap ? t->q_cap * 2 : 128);
    int pl = 6 * t->n * t->n, i = t->q_len++;
    tgo_planes(b, color, t->q_planes + (size_t)i * pl);
    for (int d = 0; d < plen; d++) { t->q_path_node[i * t->max_depth + d] = pn[d]; t->q_path_child[i * t->max_depth + d] = pc[d]; }
    t->q_path_len[i] = plen; t->q_node_index[i] = node_index;
}

static int expand_node(TgoTree *t, const TgoBoard *b, int color)
{   /* mcts/tree.py:247-270 + node.py:41-72 */
    int idx = t->num_nodes;
    if (idx >= t->tree_size) {                  /* :254-258 doubles the array */
        t->node = (TgoNode *)realloc(t->node, sizeof(TgoNode) * (size_t)t->tree_size * 2);
        memset(t->node + t->tree_size, 0, sizeof(TgoNode) * (size_t)t->tree_size);
        t->tree_size *= 2;
    }
    TgoNode *nd = &t->node[idx];
    memset(nd, 0, sizeof *nd);
    int k = tgo_candidates(b, color, nd->action);
    nd->num_children = k;
    for (int i = 0; i < t->max_actions; i++) nd->children_index[i] = TGO_NOT_EXPANDED;
    tgo_dirichlet(t->seed, t->game, t->move, (uint32_t)idx, k, nd->children_policy);   /* :266, 509-519 */
    t->num_nodes++;
    return idx;
}

static void process_mini_batch(TgoTree *t, const TgoBoard *b, int use_logit)
{   /* mcts/tree.py:273-315 */
    if (t->q_len == 0) return;
    const int nn = t->n * t->n, A = nn + 1;
    float *policy = (float *)malloc(sizeof(float) * (size_t)t->q_len * A);
    float *value = (float *)malloc(sizeof(float) * (size_t)t->q_len * 3);
    t->eval(t->eval_ctx, t->q_planes, t->q_len, use_logit, policy, value);
    t->evals += t->q_len;
    for (int i = 0; i < t->q_len; i++) {
        const float *pol = policy + (size_t)i * A, *v = value + (size_t)i * 3;
        int ni = t->q_node_index[i];
        if (ni >= 0) {
            /* ni == -1 (every sequential-halving leaf, tree.py:412-416) lands on
             * self.node[-1], a never-expanded node: no observable effect (Q1). */
            TgoNode *nd = &t->node[ni];
            for (int c = 0; c < nd->num_children; c++) {          /* node.py:86-93 */
                int a = nd->action[c];
                float p;
                if (a == TGO_PASS) { p = pol[nn]; if (use_logit) p = p - 0.5f; }   /* :292-294, fp32 */
                else { int x = a % b->w - 1, y = a / b->w - 1; p = pol[y * t->n + x]; }
                nd->children_policy[c] = (double)p;
            }
            nd->raw_value = v[1] * 0.5f + v[2];                   /* :300, fp32 tensor math */
        }
        int plen = t->q_path_len[i];
        if (plen > 0) {
            float val = v[0] + v[1] * 0.5f;                        /* :303 */
            const int *pn = t->q_path_node + i * t->max_depth, *pc = t->q_path_child + i * t->max_depth;
            t->node[pn[plen - 1]].children_value[pc[plen - 1]] = val;            /* :308 */
            for (int d = plen - 1; d >= 0; d--) {                  /* :310-313 */
                TgoNode *nd = &t->node[pn[d]]; int c = pc[d];
                nd->children_value_sum[c] = nd->children_value_sum[c] + val;      /* fp32 accumulate */
                nd->children_visits[c] += 1; nd->children_virtual_loss[c] -= 1;
                nd->node_value_sum = nd->node_value_sum + val;
                nd->node_visits += 1; nd->virtual_loss -= 1;
                val = 1.0f - val;                                  /* fp32 */
            }
        }
    }
    free(policy); free(value);
    t->q_len = 0;
}

/* ---- selection rules (mcts/node.py) ---- */
static int select_puct(const TgoTree *t, const TgoNode *nd)
{   /* node.py:141-157 + pucb/pucb.py:8-29 */
    int best = 0; double bestv = 0.0;
    double sq = sqrt((double)(nd->node_visits + nd->virtual_loss + 1));
    for (int i = 0; i < nd->num_children; i++) {
        int cv = nd->children_visits[i] + nd->children_virtual_loss[i];
        double q = cv != 0 ? (double)nd->children_value_sum[i] / (double)cv : 0.0;
        double u = (1.0 * (nd->children_policy[i] + nd->noise[i])) * sq / (double)(cv + 1);
        double v = q + u;
        if (t->cgos_mode && i == nd->num_children - 1) v = v - 0.1;
        if (i == 0 || v > bestv) { best = i; bestv = v; }
    }
    return best;
}
static int select_sh_root(const TgoNode *nd, int thr)
{   /* node.py:324-346 */
    int mx = 0;
    for (int i = 0; i < nd->num_children; i++) if (nd->children_visits[i] > mx) mx = nd->children_visits[i];
    double sigma = (double)(C_VISIT + mx) * C_SCALE;
    int best = 0; double bestv = 0.0;
    for (int i = 0; i < nd->num_children; i++) {
        int cnt = nd->children_visits[i] + nd->children_virtual_loss[i];
        double q = nd->children_visits[i] > 0 ? (double)nd->children_value_sum[i] / (double)nd->children_visits[i] : 0.0;
        double v = cnt >= thr ? -10000.0 : nd->children_policy[i] + nd->noise[i] + sigma * q;
        if (i == 0 || v > bestv) { best = i; bestv = v; }
    }
    return best;
}
static void improved_policy(const TgoTree *t, const TgoNode *nd, double *out)
{   /* node.py:281-321 */
    const int k = nd->num_children;
    double pol[TGO_MAX_ACTIONS], q[TGO_MAX_ACTIONS], pq[TGO_MAX_ACTIONS], lg[TGO_MAX_ACTIONS];
    int mx = 0;
    for (int i = 0; i < t->max_actions; i++) if (nd->children_visits[i] > mx) mx = nd->children_visits[i];
    double sigma = (double)(C_VISIT + mx) * C_SCALE;
    tgo_softmax(nd->children_policy, k, pol, t->use_libm);
    for (int i = 0; i < k; i++) {
        q[i] = nd->children_visits[i] > 0 ? (double)nd->children_value_sum[i] / (double)nd->children_visits[i] : 0.0;
        pq[i] = pol[i] * q[i];
    }
    double sum_prob = tgo_np_sum(pol, k), v_pi = tgo_np_sum(pq, k);
    double vmix = ((double)nd->raw_value * 1.0 + (double)nd->node_visits * v_pi / sum_prob) / ((double)nd->node_visits + 1.0);
    for (int i = 0; i < k; i++) {
        double cq = nd->children_visits[i] > 0 ? q[i] : vmix;
        lg[i] = nd->children_policy[i] + sigma * cq;
    }
    tgo_softmax(lg, k, out, t->use_libm);
}
void tgo_improved_policy(const TgoTree *t, int node_index, double *out) { improved_policy(t, &t->node[node_index], out); }

static int select_sh_node(const TgoTree *t, const TgoNode *nd)
{   /* node.py:349-361 */
    double ip[TGO_MAX_ACTIONS];
    improved_policy(t, nd, ip);
    int best = 0; double bestv = 0.0;
    for (int i = 0; i < nd->num_children; i++) {
        double v = ip[i] - (double)nd->children_visits[i] / (1.0 + (double)nd->node_visits);
        if (i == 0 || v > bestv) { best = i; bestv = v; }
    }
    return best;
}

int tgo_sh_schedule(int m, int visits, int *considered, int *counts, int cap)
{   /* mcts/sequential_halving.py:7-60 */
    int *seq = (int *)malloc(sizeof(int) * (size_t)(visits + 2 * (m > 16 ? m : 16) * (visits + 1)));
    int len = 0;
    if (m <= 1) { for (int i = 0; i < visits; i++) seq[len++] = i; }
    else {
        int log2max = (int)ceil(log2((double)m));
        int *vis = (int *)calloc((size_t)m, sizeof(int));
        int nc = m;
        while (len < visits) {
            int extra = (int)((double)visits / (double)(log2max * nc));
            if (extra < 1) extra = 1;
            for (int e = 0; e < extra && len < visits + m; e++) {
                for (int i = 0; i < nc; i++) seq[len++] = vis[i];
                for (int i = 0; i < nc; i++) vis[i]++;
            }
            nc = nc / 2 > 2 ? nc / 2 : 2;
        }
        if (len > visits) len = visits;
        free(vis);
    }
    int mxv = 0; for (int i = 0; i < len; i++) if (seq[i] > mxv) mxv = seq[i];
    int *cl = (int *)calloc((size_t)mxv + 1, sizeof(int));
    for (int i = 0; i < len; i++) cl[seq[i]]++;
    int np = 0;
    for (int v = 0; v <= mxv; v++) {            /* dict keyed by count, insertion-ordered */
        int f = -1;
        for (int j = 0; j < np; j++) if (considered[j] == cl[v]) f = j;
        if (f >= 0) counts[f]++;
        else if (np < cap) { considered[np] = cl[v]; counts[np] = 1; np++; }
    }
    free(cl); free(seq);
    return np;
}

static void descend_sh(TgoTree *t, TgoBoard *sb, int color, int thr)
{   /* mcts/tree.py:387-422 (recursion unrolled) */
    int pn[64], pc[64], plen = 0, cur = t->current_root;
    for (;;) {
        TgoNode *nd = &t->node[cur];
        int next = (cur == t->current_root) ? select_sh_root(nd, thr) : select_sh_node(t, nd);
        int mv = nd->action[next];
        pn[plen] = cur; pc[plen] = next; plen++;
        tgo_put_stone(sb, mv, color);
        color = opp(color);
        nd->virtual_loss += 1; nd->children_virtual_loss[next] += 1;     /* node.py:76-83 */
        if (nd->children_visits[next] < 1) {
            q_push(t, sb, color, pn, pc, plen, nd->children_index[next]);
            return;
        }
        if (nd->children_index[next] == TGO_NOT_EXPANDED) {
            int ci = expand_node(t, sb, color);
            nd = &t->node[cur];                  /* array may have moved */
            nd->children_index[next] = ci;
        }
        cur = nd->children_index[next];
        if (plen >= 64) { t->error = 1; return; }
    }
}

static double value_evaluation(const TgoNode *nd, int idx)
{   /* node.py:364-375 */
    if (nd->children_visits[idx] == 0) return 0.5;
    return (double)nd->children_value_sum[idx] / (double)nd->children_visits[idx];
}

int tgo_genmove_sh(TgoTree *t, const TgoBoard *b, int color, int visits, int never_resign)
{   /* mcts/tree.py:318-356, 359-384 */
    t->num_nodes = 0; t->q_len = 0;
    t->current_root = expand_node(t, b, color);
    q_push(t, b, color, NULL, NULL, 0, t->current_root);
    process_mini_batch(t, b, 1);
    TgoNode *root = &t->node[t->current_root];
    tgo_gumbel(t->seed, t->game, t->move, t->max_actions, root->noise);     /* node.py:275-278 */

    int k = root->num_children;
    int m = k < MAX_CONSIDERED ? k : MAX_CONSIDERED;
    int cons[64], cnts[64];
    int np = tgo_sh_schedule(m, visits, cons, cnts, 64);
    TgoBoard sb;
    for (int p = 0; p < np; p++) {
        for (int thr = 0; thr < cnts[p]; thr++)
            for (int j = 0; j < cons[p]; j++) {
                tgo_board_copy(&sb, b);
                descend_sh(t, &sb, color, thr + 1);
            }
        process_mini_batch(t, b, 1);
    }
    root = &t->node[t->current_root];
    int next = select_sh_root(root, PLAYOUTS);                              /* :344 */
    double value = value_evaluation(root, next);
    if (!never_resign && value < 0.05) return TGO_RESIGN;                   /* :353 */
    return root->action[next];
}

static void descend_puct(TgoTree *t, TgoBoard *sb, int color, const TgoBoard *rootb)
{   /* mcts/tree.py:199-244 */
    int *pn = (int *)malloc(sizeof(int) * (size_t)t->max_depth), *pc = (int *)malloc(sizeof(int) * (size_t)t->max_depth);
    int plen = 0, cur = t->current_root;
    for (;;) {
        TgoNode *nd = &t->node[cur];
        int next = select_puct(t, nd);
        int mv = nd->action[next];
        pn[plen] = cur; pc[plen] = next; plen++;
        tgo_put_stone(sb, mv, color);
        color = opp(color);
        nd->virtual_loss += 1; nd->children_virtual_loss[next] += 1;
        int expand_threshold = 1;
        if (sb->moves > 2) {                                               /* :224-229 */
            if (sb->moves - 1 >= sb->max_records) { t->error = 2; break; } /* record.get would raise (Q12) */
            if (sb->hist_pos[sb->moves - 1] == TGO_PASS && sb->hist_pos[sb->moves - 2] == TGO_PASS)
                expand_threshold = 10000000;
        }
        if (nd->children_visits[next] + nd->children_virtual_loss[next] < expand_threshold + 1) {
            int ci;
            if (nd->children_index[next] == TGO_NOT_EXPANDED) {
                ci = expand_node(t, sb, color);
                nd = &t->node[cur];
                nd->children_index[next] = ci;
            } else ci = nd->children_index[next];
            q_push(t, sb, color, pn, pc, plen, ci);
            if (t->q_len >= t->batch_size) process_mini_batch(t, rootb, 0);
            break;
        }
        cur = nd->children_index[next];
        if (plen >= t->max_depth) { t->error = 1; break; }
    }
    free(pn); free(pc);
}

int tgo_genmove_puct(TgoTree *t, const TgoBoard *b, int color, int visits, int strict)
{   /* mcts/tree.py:57-105, 130-152; time_manager.py:146-163 */
    t->num_nodes = 0; t->q_len = 0;
    t->current_root = expand_node(t, b, color);                             /* :49-54 */
    q_push(t, b, color, NULL, NULL, 0, t->current_root);
    process_mini_batch(t, b, 0);
    if (t->node[t->current_root].num_children == 1) return TGO_PASS;        /* :76-77 */
    TgoBoard sb;
    for (int counter = 0; counter < visits; counter++) {
        tgo_board_copy(&sb, b);
        descend_puct(t, &sb, color, b);
        if (t->error) break;
        const TgoNode *root = &t->node[t->current_root];
        int top1 = 0, top2 = 0;                  /* sorted(children_visits)[-1], [-2] over MAX_ACTIONS */
        for (int i = 0; i < t->max_actions; i++) {
            int v = root->children_visits[i];
            if (v > top1) { top2 = top1; top1 = v; } else if (v > top2) top2 = v;
        }
        int remaining = visits - root->node_visits, cutoff = strict ? 0 : top1 - top2;
        if (remaining < cutoff) break;
    }
    if (t->q_len > 0) process_mini_batch(t, b, 0);                          /* :82-83 */
    const TgoNode *root = &t->node[t->current_root];
    int best = 0;                                                           /* node.py:169-175 */
    for (int i = 1; i < root->num_children; i++) if (root->children_visits[i] > root->children_visits[best]) best = i;
    if (value_evaluation(root, best) < RESIGN_THRESHOLD) return TGO_RESIGN; /* :100-103 */
    return root->action[best];
}

int tgo_selfplay_game(TgoTree *t, int n, double komi, const uint64_t *zob, uint64_t seed, uint64_t game,
                      int visits, int never_resign, int use_puct, TgoGameRecord *rec,
                      double *improved, int16_t *actions)
{   /* selfplay/worker.py:46-90 + sgf/selfplay_record.py:45-64 */
    TgoBoard b;
    tgo_board_init(&b, n, komi, 1, zob);
    const int max_moves = 2 * n * n, A = n * n + 1;
    int color = TGO_BLACK, pass_count = 0;
    memset(rec, 0, sizeof *rec);
    rec->winner = TGO_EMPTY;
    for (int mv = 0; mv < max_moves; mv++) {
        tgo_tree_set_noise_key(t, seed, game, (uint32_t)b.moves);
        int pos = use_puct ? tgo_genmove_puct(t, &b, color, visits, 0)
                           : tgo_genmove_sh(t, &b, color, visits, never_resign);
        if (t->error) return -1;
        if (pos == TGO_RESIGN) { rec->winner = opp(color); rec->is_resign = 1; break; }
        tgo_put_stone(&b, pos, color);
        pass_count = (pos == TGO_PASS) ? pass_count + 1 : 0;
        int i = rec->n_moves++;
        rec->pos[i] = (int16_t)pos; rec->color[i] = (uint8_t)color;
        const TgoNode *root = &t->node[t->current_root];
        rec->num_children[i] = (int16_t)root->num_children;
        if (improved) improved_policy(t, root, improved + (size_t)i * A);
        if (actions) memcpy(actions + (size_t)i * A, root->action, sizeof(int16_t) * (size_t)A);
        color = opp(color);
        if (pass_count == 2) { rec->winner = TGO_EMPTY; break; }
    }
    if (pass_count == 2) {                                                  /* :80-87 */
        rec->score = (double)tgo_count_score(&b) - komi;
        if (rec->score > 0.1) rec->winner = TGO_BLACK;
        else if (rec->score < -0.1) rec->winner = TGO_WHITE;
        else rec->winner = TGO_OB;
    }
    return rec->n_moves;
}

/* ------------------------------------------------------------------ */
/* native hash evaluators (ours; numpy twins: oracle.py hashnet / hashnet2) */
/* ------------------------------------------------------------------ */
void tgo_hashnet_eval(void *ctx, const float *planes, int nb, int use_logit, float *policy, float *value)
{
    const TgoHashNetCtx *c = (const TgoHashNetCtx *)ctx;
    const int nn = c->n * c->n, npl = 6 * nn, A = nn + 1;
    for (int s = 0; s < nb; s++) {
        const float *pl = planes + (size_t)s * npl;
        uint64_t h = 0;
        for (int j = 0; j < npl; j++) h += mix64(3ull * (uint64_t)j + (uint64_t)(long long)(pl[j] + 1.0f));
        for (int i = 0; i < A; i++) {
            const float raw = (float)((mix64(h + (uint64_t)i) >> 40) & 0xFFFFull);
            if (c->variant == 0) policy[(size_t)s * A + i] = use_logit ? raw / 8192.0f - 4.0f : raw / 1048576.0f;
            else policy[(size_t)s * A + i] = use_logit ? raw / 1000.0f - 4.0f : raw / 1000000.0f;
        }
        const uint64_t ra = mix64(h + 1000ull), rb = mix64(h + 1001ull);
        const float va = c->variant == 0 ? (float)(ra & 0xFFull) : (float)(ra % 500ull);
        const float vb = c->variant == 0 ? (float)(rb & 0xFFull) : (float)(rb % 500ull);
        const float den = c->variant == 0 ? 512.0f : 1000.0f;
        const float v0 = va / den, v1 = vb / den;
        value[s * 3 + 0] = v0; value[s * 3 + 1] = v1; value[s * 3 + 2] = (1.0f - v0) - v1;
    }
}
tgo_eval_fn tgo_hashnet_fn(void) { return tgo_hashnet_eval; }

/* ------------------------------------------------------------------ */
/* bulk differential corpus (ours): random games with one 64-bit digest */
/* of the whole observable board state per ply.  The numpy twin of the  */
/* digest (tests/gpu_util.py ply_digest) is applied to the engine's dump.*/
/* ------------------------------------------------------------------ */
static inline uint64_t dw(uint64_t salt, uint64_t i) { return mix64(salt * 0x100000001B3ull + i) | 1ull; }

uint64_t tgo_ply_digest(const TgoBoard *b)
{
    const int nn = b->n * b->n;
    uint64_t acc = b->hash * dw(5, 0);
    for (int c = 0; c < b->cells; c++) {
        int col = b->color[c], libs = 0, size = 0;
        if (col == TGO_BLACK || col == TGO_WHITE) { libs = b->libs[b->chain[c]]; size = b->size[b->chain[c]]; }
        acc += (uint64_t)(col + 4 * libs + 4096 * size) * dw(1, (uint64_t)c);
    }
    uint8_t legal[TGO_MAX_ACTIONS], eye[TGO_MAX_ACTIONS], cand[TGO_MAX_ACTIONS]; int16_t sa[TGO_MAX_ACTIONS];
    for (int ci = 0; ci < 2; ci++) {
        tgo_analyze(b, ci == 0 ? TGO_BLACK : TGO_WHITE, legal, sa, eye, cand);
        for (int i = 0; i < nn; i++)
            acc += (uint64_t)(legal[i] + 2 * eye[i] + 4 * cand[i] + 8 * (int)sa[i]) * dw(2, (uint64_t)(ci * nn + i));
    }
    const int sc[5] = { b->moves, b->ko_pos, b->ko_move, b->prisoner[0], b->prisoner[1] };
    for (int i = 0; i < 5; i++) acc += (uint64_t)(int64_t)sc[i] * dw(3, (uint64_t)i);
    acc += (uint64_t)(int64_t)tgo_count_score(b) * dw(4, 0);
    return acc;
}

int tgo_random_game(int n, const uint64_t *zob, int superko, uint64_t seed, uint64_t game, int max_plies,
                    double p_pass, double p_any_legal, int16_t *moves_out, uint64_t *digest_out)
{
    TgoBoard b;
    tgo_board_init(&b, n, 7.0, superko, zob);
    const int nn = n * n;
    int color = TGO_BLACK, plies = 0, passes = 0;
    uint8_t legal[TGO_MAX_ACTIONS], eye[TGO_MAX_ACTIONS], cand[TGO_MAX_ACTIONS]; int16_t sa[TGO_MAX_ACTIONS];
    for (; plies < max_plies; plies++) {
        uint64_t r = mix64(mix64(seed + game) + (uint64_t)plies);
        const double u0 = (double)(r >> 11) * (1.0 / 9007199254740992.0);
        r = mix64(r);
        const double u1 = (double)(r >> 11) * (1.0 / 9007199254740992.0);
        r = mix64(r);
        tgo_analyze(&b, color, legal, sa, eye, cand);
        const uint8_t *pool = u1 < p_any_legal ? legal : cand;
        int cnt = 0;
        for (int i = 0; i < nn; i++) cnt += pool[i];
        int pos = TGO_PASS;
        if (cnt > 0 && u0 >= p_pass) {
            int pick = (int)(r % (uint64_t)cnt);
            for (int i = 0; i < nn; i++) if (pool[i] && pick-- == 0) { pos = tgo_onboard_pos(&b, i); break; }
        }
        tgo_put_stone(&b, pos, color);
        moves_out[plies] = (int16_t)pos;
        digest_out[plies] = tgo_ply_digest(&b);
        color = opp(color);
        passes = pos == TGO_PASS ? passes + 1 : 0;
        if (passes >= 2 && plies > 20) { plies++; break; }
    }
    return plies;
}

/* ------------------------------------------------------------------ */
/* Tromp-Taylor area score (ours; the adjudication the reference gets   */
/* from GNU Go, get_final_status.py:15-64, restated as a flood fill):    */
/* stones + empty regions that reach only one colour.  Black - White.    */
/* ------------------------------------------------------------------ */
int tgo_tromp_taylor(const TgoBoard *b)
{
    uint8_t seen[TGO_MAX_CELLS];
    int stack[TGO_MAX_CELLS];
    memset(seen, 0, sizeof seen);
    int score = 0;
    for (int i = 0; i < b->n * b->n; i++) {
        const int p0 = tgo_onboard_pos(b, i);
        if (b->color[p0] == TGO_BLACK) { score++; continue; }
        if (b->color[p0] == TGO_WHITE) { score--; continue; }
        if (seen[p0]) continue;
        int sp = 0, cnt = 0, reach = 0;
        stack[sp++] = p0; seen[p0] = 1;
        while (sp > 0) {
            const int p = stack[--sp];
            cnt++;
            int q[4]; nbr4(b, p, q);
            for (int k = 0; k < 4; k++) {
                const int c = b->color[q[k]];
                if (c == TGO_BLACK) reach |= 1;
                else if (c == TGO_WHITE) reach |= 2;
                else if (c == TGO_EMPTY && !seen[q[k]]) { seen[q[k]] = 1; stack[sp++] = q[k]; }
            }
        }
        if (reach == 1) score += cnt; else if (reach == 2) score -= cnt;
    }
    return score;
}
