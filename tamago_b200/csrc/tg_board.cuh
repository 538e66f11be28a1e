// tg_board.cuh -- warp-cooperative Go board living in shared memory.
//
// One warp owns one board.  Every function here is warp-collective: all 32
// lanes call it with identical arguments.  Scalars (hash, move number, ko,
// prisoners) are kept replicated in registers (BScal); per-point arrays live in
// shared memory (WBoard) and are swept 32 points at a time.
//
// The reference keeps strings as sorted intrusive linked lists
// (board/string.py:9-247) that are updated stone by stone.  What is observable
// from them -- liberty count, string size, membership -- is a pure function of
// the stones, so the device keeps one label per stone (`chain`, the position of
// some stone of the string) and rebuilds liberty counts / sizes for all strings
// with one atomic sweep after each move.  Results are bit-identical to the
// reference (tests/test_board_gpu.py) while every step is a coalesced sweep.
#pragma once
#include "tg_common.cuh"

namespace tg {

template <int N> struct WBoard {
    using G = Geo<N>;
    unsigned ls[G::CP];        // by label: liberties << 16 | size
    unsigned bloom[BLOOM_WORDS];
    uint16_t chain[G::CP];     // label of the string owning a stone
    uint8_t  color[G::CP];
};

// Scratch for the expansion-time analysis of one board (one per warp).
template <int N> struct WAnalysis {
    using G = Geo<N>;
    u64      cx[G::CP];        // by label: XOR of mover-opponent Zobrist keys of the string (super-ko)
    unsigned lmin[G::CP];      // by label: smallest liberty position
    unsigned lmax[G::CP];      // by label: largest liberty position
};

struct BScal {                 // replicated across the warp
    u64 hash;
    int moves, ko_pos, ko_move, pris0, pris1;
};

// Root-board pool in HBM: one row per game and field (structure of arrays).
template <int N> struct BoardPool {
    using G = Geo<N>;
    uint8_t*  color;      // [games][CP]
    uint16_t* chain;      // [games][CP]
    unsigned* bloom;      // [games][BLOOM_WORDS]
    u64*      hash;       // [games]
    int*      scal;       // [games][8]: moves, ko_pos, ko_move, pris0, pris1, -, -, -
    u64*      hist_hash;  // [games][MAXREC]   record.py:19
    int16_t*  hist_pos;   // [games][MAXREC]   record.py:18
};

template <int N> __device__ __forceinline__ bool on_board(int pos)
{
    const int x = pos % Geo<N>::W, y = pos / Geo<N>::W;
    return x >= 1 && x <= N && y >= 1 && y <= N;
}
template <int N> __device__ __forceinline__ int onboard_pos(int idx)   // go_board.py:82-86
{
    return (idx % N + 1) + (idx / N + 1) * Geo<N>::W;
}

// go_board.py:111-129
template <int N> __device__ inline void wb_clear(WBoard<N>& b, BScal& s, int lane)
{
    using G = Geo<N>;
    for (int c = lane; c < G::CP; c += 32) {
        b.color[c] = (c < G::CELLS && on_board<N>(c)) ? EMPTY : OB;
        b.chain[c] = 0; b.ls[c] = 0;
    }
    for (int i = lane; i < BLOOM_WORDS; i += 32) b.bloom[i] = 0;
    s.hash = 0; s.moves = 1; s.ko_pos = 0; s.ko_move = 0; s.pris0 = 0; s.pris1 = 0;
    __syncwarp();
}

// Rebuild liberties and sizes of every string from colours + labels.
template <int N> __device__ inline void wb_recount(WBoard<N>& b, int lane)
{
    using G = Geo<N>;
    for (int c = lane; c < G::CP; c += 32) b.ls[c] = 0;
    __syncwarp();
    for (int c = lane; c < G::CELLS; c += 32) {
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
                for (int k = 0; k < ns; k++) dup |= (seen[k] == l);
                if (!dup) { seen[ns++] = l; atomicAdd(&b.ls[l], 1u << 16); }
            }
        }
    }
    __syncwarp();
}

template <int N> __device__ __forceinline__ int wb_libs(const WBoard<N>& b, int pos) { return (int)(b.ls[b.chain[pos]] >> 16); }

template <int N>
__device__ inline void wb_load(WBoard<N>& b, BScal& s, const BoardPool<N>& pool, int g, int lane)
{
    using G = Geo<N>;
    const uint32_t* c4 = reinterpret_cast<const uint32_t*>(pool.color + (size_t)g * G::CP);
    uint32_t* d4 = reinterpret_cast<uint32_t*>(b.color);
    for (int i = lane; i < G::CP / 4; i += 32) d4[i] = c4[i];
    const uint32_t* h2 = reinterpret_cast<const uint32_t*>(pool.chain + (size_t)g * G::CP);
    uint32_t* e2 = reinterpret_cast<uint32_t*>(b.chain);
    for (int i = lane; i < G::CP / 2; i += 32) e2[i] = h2[i];
    for (int i = lane; i < BLOOM_WORDS; i += 32) b.bloom[i] = pool.bloom[(size_t)g * BLOOM_WORDS + i];
    const int* sc = pool.scal + (size_t)g * 8;
    s.hash = pool.hash[g];
    s.moves = sc[0]; s.ko_pos = sc[1]; s.ko_move = sc[2]; s.pris0 = sc[3]; s.pris1 = sc[4];
    __syncwarp();
    wb_recount(b, lane);
}

template <int N>
__device__ inline void wb_store(const WBoard<N>& b, const BScal& s, const BoardPool<N>& pool, int g, int lane)
{
    using G = Geo<N>;
    __syncwarp();
    uint32_t* c4 = reinterpret_cast<uint32_t*>(pool.color + (size_t)g * G::CP);
    const uint32_t* d4 = reinterpret_cast<const uint32_t*>(b.color);
    for (int i = lane; i < G::CP / 4; i += 32) c4[i] = d4[i];
    uint32_t* h2 = reinterpret_cast<uint32_t*>(pool.chain + (size_t)g * G::CP);
    const uint32_t* e2 = reinterpret_cast<const uint32_t*>(b.chain);
    for (int i = lane; i < G::CP / 2; i += 32) h2[i] = e2[i];
    for (int i = lane; i < BLOOM_WORDS; i += 32) pool.bloom[(size_t)g * BLOOM_WORDS + i] = b.bloom[i];
    if (lane == 0) {
        int* sc = pool.scal + (size_t)g * 8;
        pool.hash[g] = s.hash;
        sc[0] = s.moves; sc[1] = s.ko_pos; sc[2] = s.ko_move; sc[3] = s.pris0; sc[4] = s.pris1;
    }
}

// scratch <- root, shared to shared (go_board.py:611-626; the record is not copied: history rows are
// shared with the root and only entries >= root moves are written by a descent)
template <int N> __device__ inline void wb_copy(WBoard<N>& dst, const WBoard<N>& src, int lane)
{
    using G = Geo<N>;
    const uint32_t* a = reinterpret_cast<const uint32_t*>(src.color);
    uint32_t* d = reinterpret_cast<uint32_t*>(dst.color);
    for (int i = lane; i < G::CP / 4; i += 32) d[i] = a[i];
    const uint32_t* a2 = reinterpret_cast<const uint32_t*>(src.chain);
    uint32_t* d2 = reinterpret_cast<uint32_t*>(dst.chain);
    for (int i = lane; i < G::CP / 2; i += 32) d2[i] = a2[i];
    for (int i = lane; i < G::CP; i += 32) dst.ls[i] = src.ls[i];
    for (int i = lane; i < BLOOM_WORDS; i += 32) dst.bloom[i] = src.bloom[i];
    __syncwarp();
}

__device__ __forceinline__ unsigned bloom_bit(u64 h) { return (unsigned)(h & (BLOOM_WORDS * 32 - 1)); }

// GoBoard.put_stone (go_board.py:131-185): place, capture, merge, ko, record.
template <int N>
__device__ inline void wb_put_stone(WBoard<N>& b, BScal& s, int pos, int color, const u64* __restrict__ zob,
                                    u64* hist_hash, int16_t* hist_pos, int lane)
{
    using G = Geo<N>;
    if (pos == PASS) {                                       // :138-141
        if (lane == 0 && s.moves < G::MAXREC) { hist_hash[s.moves] = s.hash; hist_pos[s.moves] = 0; }
        s.moves++;
        __syncwarp();
        return;
    }
    const int other = opp(color);
    const int q[4] = { pos - G::W, pos - 1, pos + 1, pos + G::W };   // up, left, right, down (:53)
    int cap[4], ncap = 0, own[4], nown = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int cc = b.color[q[i]];
        if (cc != color && cc != other) continue;
        const int l = b.chain[q[i]];
        if (cc == color) {                                   // :155-157 `connection`
            bool dup = false;
            for (int k = 0; k < nown; k++) dup |= (own[k] == l);
            if (!dup) own[nown++] = l;
        } else if ((b.ls[l] >> 16) == 1u) {                  // :158-166 pos was its last liberty
            bool dup = false;
            for (int k = 0; k < ncap; k++) dup |= (cap[k] == l);
            if (!dup) cap[ncap++] = l;
        }
    }
    __syncwarp();
    const int label = nown > 0 ? own[0] : pos;
    if (lane == 0) { b.color[pos] = (uint8_t)color; b.chain[pos] = (uint16_t)label; }
    s.hash ^= zob[color * G::CELLS + pos];                   // :147
    int prisoner = 0;
    if (ncap == 0 && nown > 1) {
        // Merge without capture (string.py:443-545): ONE sweep relabels the absorbed strings and recounts the liberties of
        // the merged string -- the empty points next to any stone of the merging strings or to the new stone -- instead of
        // the relabel sweep plus the two sweeps of the full recount.  Sizes add up; adjacent enemy strings lose pos.
        __syncwarp();
        unsigned size = 1;
        for (int k = 0; k < nown; k++) size += b.ls[own[k]] & 0xffffu;
        int cnt = 0;
        for (int c = lane; c < G::CELLS; c += 32) {
            const int cc = b.color[c];
            if (cc == color && c != pos) {
                const int l = b.chain[c];
                bool hit = false;
                for (int k = 1; k < nown; k++) hit |= (own[k] == l);
                if (hit) b.chain[c] = (uint16_t)label;
            } else if (cc == EMPTY) {
                const int r[4] = { c - G::W, c - 1, c + 1, c + G::W };
                bool adj = false;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (r[j] == pos) { adj = true; continue; }
                    if (b.color[r[j]] != color) continue;
                    const int l = b.chain[r[j]];                 // an absorbed label or already `label` (another lane may be
                                                                 // relabelling it right now): both are in own[], same verdict
                    for (int k = 0; k < nown; k++) adj |= (own[k] == l);
                }
                cnt += adj;
            }
        }
        const unsigned libs = (unsigned)warp_sum_i(cnt);
        if (lane == 0) {
            b.ls[label] = (libs << 16) | size;
            int el[4], ne = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if (b.color[q[i]] != other) continue;
                const int l = b.chain[q[i]];
                bool dup = false;
                for (int k = 0; k < ne; k++) dup |= (el[k] == l);
                if (!dup) { el[ne++] = l; b.ls[l] -= 1u << 16; }
            }
            if (s.moves < G::MAXREC) { hist_hash[s.moves] = s.hash; hist_pos[s.moves] = (int16_t)pos; }   // record.py:30-44
            const unsigned bit = bloom_bit(s.hash);
            b.bloom[bit >> 5] |= 1u << (bit & 31);
        }
        s.moves++;
        __syncwarp();
        return;                                              // no prisoners: the ko rule (:173-177) cannot apply
    }
    if (ncap > 0 || nown > 1) {
        __syncwarp();
        u64 hx = 0; int cnt = 0;
        for (int c = lane; c < G::CELLS; c += 32) {
            const int cc = b.color[c];
            if (cc == other) {
                const int l = b.chain[c];
                bool hit = false;
                for (int k = 0; k < ncap; k++) hit |= (cap[k] == l);
                if (hit) { b.color[c] = EMPTY; hx ^= zob[other * G::CELLS + c]; cnt++; }   // string.py:286-325
            } else if (cc == color && c != pos) {
                const int l = b.chain[c];
                bool hit = false;
                for (int k = 1; k < nown; k++) hit |= (own[k] == l);
                if (hit) b.chain[c] = (uint16_t)label;       // string.py:443-545 merge
            }
        }
        s.hash ^= warp_xor64(hx);                            // :165-166
        prisoner = warp_sum_i(cnt);
    }
    if (color == BLACK) s.pris0 += prisoner; else s.pris1 += prisoner;   // :168-171
    __syncwarp();
    if (ncap == 0 && nown <= 1) {
        // Nothing was captured and at most one own string is extended: liberties and sizes change only around pos,
        // so the full recount sweep is replaced by a local update with the same result.  Every distinct adjacent
        // enemy string loses the liberty pos; the stone's string loses pos, gains the empty neighbours of pos that
        // were not its liberties yet (string.py:411-441 add_stone / 371-409 make_string) and grows by one.
        // One lane per neighbour (the four checks are independent), combined with a ballot.
        bool gain = false;
        if (lane < 4) {
            const int qi = q[lane];
            const int cc = b.color[qi];
            if (cc == other) {
                const int l = b.chain[qi];
                bool dup = false;
                for (int k = 0; k < lane; k++) dup |= (b.color[q[k]] == other && b.chain[q[k]] == l);
                if (!dup) b.ls[l] -= 1u << 16;               // distinct labels per lane: no two lanes touch the same word
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
        if (lane == 0) {
            if (nown == 1) b.ls[label] += (gained << 16) - (1u << 16) + 1u;
            else b.ls[label] = (gained << 16) | 1u;
        }
        __syncwarp();
    } else {
        wb_recount(b, lane);
    }
    if (nown == 0 && prisoner == 1 && (b.ls[label] >> 16) == 1u) {       // :173-177
        s.ko_move = s.moves;
#pragma unroll
        for (int i = 0; i < 4; i++) if (b.color[q[i]] == EMPTY) s.ko_pos = q[i];
    }
    if (lane == 0) {
        if (s.moves < G::MAXREC) { hist_hash[s.moves] = s.hash; hist_pos[s.moves] = (int16_t)pos; }   // record.py:30-44
        const unsigned bit = bloom_bit(s.hash);
        b.bloom[bit >> 5] |= 1u << (bit & 31);
    }
    s.moves++;
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// Expansion-time analysis: legality (incl. positional super-ko), self-atari and complete-eye filters.
// ---------------------------------------------------------------------------------------------
template <int N> __device__ __forceinline__ unsigned pat3_at(const WBoard<N>& b, int pos)
{   // pattern.py:47-50: UL,U,UR,L,R,DL,D,DR, two bits each from bit 0
    constexpr int W = Geo<N>::W;
    return (unsigned)b.color[pos - W - 1] | ((unsigned)b.color[pos - W] << 2) | ((unsigned)b.color[pos - W + 1] << 4)
         | ((unsigned)b.color[pos - 1] << 6) | ((unsigned)b.color[pos + 1] << 8)
         | ((unsigned)b.color[pos + W - 1] << 10) | ((unsigned)b.color[pos + W] << 12) | ((unsigned)b.color[pos + W + 1] << 14);
}

// per-string liberty extremes and (for super-ko) string key XORs; call before wb_point_status
template <int N>
__device__ inline void wb_prepare_analysis(const WBoard<N>& b, WAnalysis<N>& an, int mover, bool superko, const u64* __restrict__ zob, int lane)
{
    using G = Geo<N>;
    for (int c = lane; c < G::CP; c += 32) { an.lmin[c] = 0xffffu; an.lmax[c] = 0; an.cx[c] = 0; }
    __syncwarp();
    const int other = opp(mover);
    for (int c = lane; c < G::CELLS; c += 32) {
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
            if ((b.ls[l] >> 16) == 1u) atomicXor(&an.cx[l], zob[other * G::CELLS + c]);   // go_board.py:290-292: always the opponent's keys
        }
    }
    __syncwarp();
}

struct PointStatus { bool legal_pre; bool need_scan; u64 h; int satari; bool eye; };

// eye colour of a 3x3 code (board/pattern.py:53-98): byte table in global memory, or the same table packed to two bits per
// code (16 KB) so that it fits in shared memory next to the boards
struct EyeLutGlobal { const uint8_t* __restrict__ t; __device__ __forceinline__ int operator()(unsigned p) const { return t[p]; } };
struct EyeLutPacked { const uint32_t* t; __device__ __forceinline__ int operator()(unsigned p) const { return (int)((t[p >> 4] >> ((p & 15u) * 2)) & 3u); } };

// Lane-local part of is_legal / check_self_atari_stone / is_complete_eye for one point.
template <int N, class EyeFn>
__device__ inline PointStatus wb_point_status(const WBoard<N>& b, const WAnalysis<N>& an, const BScal& s, int pos, int color, bool superko,
                                              const u64* __restrict__ zob, const EyeFn& eye_lut)
{
    using G = Geo<N>;
    PointStatus r; r.legal_pre = false; r.need_scan = false; r.h = 0; r.satari = 0; r.eye = false;
    if (b.color[pos] != EMPTY) return r;                                    // go_board.py:271-272
    const int other = opp(color);
    const int q[4] = { pos - G::W, pos - 1, pos + 1, pos + G::W };
    int ncol[4], nl[4], nlibs[4], nempty = 0;
    bool suicide = true;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        ncol[i] = b.color[q[i]];
        nl[i] = -1; nlibs[i] = 0;
        if (ncol[i] == EMPTY) nempty++;
        else if (ncol[i] == BLACK || ncol[i] == WHITE) {
            nl[i] = b.chain[q[i]]; nlibs[i] = (int)(b.ls[nl[i]] >> 16);
            if (ncol[i] == other && nlibs[i] == 1) suicide = false;         // :252-253
            if (ncol[i] == color && nlibs[i] > 1) suicide = false;          // :254-255
        }
    }
    if (nempty == 0 && suicide) return r;                                   // :275-277
    if (s.ko_pos == pos && s.ko_move == s.moves - 1) return r;              // :280-281
    r.legal_pre = true;
    if (superko) {                                                          // :284-301
        u64 h = s.hash ^ zob[color * G::CELLS + pos];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (nl[i] < 0 || nlibs[i] != 1) continue;
            bool dup = false;
            for (int k = 0; k < i; k++) dup |= (nl[k] == nl[i]);
            if (!dup) h ^= an.cx[nl[i]];
        }
        r.h = h;
        if (h == 0) r.legal_pre = false;            // record.py:63 also matches the unused (zero) slots
        else { const unsigned bit = bloom_bit(h); r.need_scan = (b.bloom[bit >> 5] >> (bit & 31)) & 1u; }
    }
    // check_self_atari_stone (go_board.py:327-365); order independent
    {
        int sa;
        if (nempty > 1) sa = 0;
        else {
            int lib[12], nlib = 0, size = 0; bool zero = false;
#pragma unroll
            for (int i = 0; i < 4; i++) if (ncol[i] == EMPTY) lib[nlib++] = q[i];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if (ncol[i] == other) { if (nlibs[i] == 1) zero = true; continue; }
                if (ncol[i] != color) continue;
                bool dup = false;
                for (int k = 0; k < i; k++) dup |= (ncol[k] == color && nl[k] == nl[i]);
                if (dup) continue;
                if (nlibs[i] >= 3) { zero = true; continue; }
                const int a = (int)an.lmin[nl[i]], c2 = (int)an.lmax[nl[i]];
                bool da = false, dc = false;
                for (int k = 0; k < nlib; k++) { da |= (lib[k] == a); dc |= (lib[k] == c2); }
                if (!da) lib[nlib++] = a;
                if (!dc && c2 != a) lib[nlib++] = c2;
                size += (int)(b.ls[nl[i]] & 0xffffu);
            }
            sa = (zero || nlib >= 3) ? 0 : size + 1;
        }
        r.satari = sa;
    }
    // is_complete_eye (go_board.py:367-397)
    {
        bool eye = false;
        if (eye_lut(pat3_at(b, pos)) == color) {
            const int x4[4] = { pos - G::W - 1, pos - G::W + 1, pos + G::W - 1, pos + G::W + 1 };
            int cnt = 0; bool edge = false;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int cc = b.color[x4[i]];
                if (cc == color || cc == OB) cnt++;
                else if (cc == EMPTY && eye_lut(pat3_at(b, x4[i])) == color) cnt++;
                if (cc == OB) edge = true;
            }
            eye = (edge && cnt == 4) || (!edge && cnt >= 3);
        }
        r.eye = eye;
    }
    return r;
}

// Warp-wide history scan for one hash (record.py:54-63).  Entries [1, moves) are live; slot 0 and the
// unused tail are zero in the reference and are covered by the h == 0 test in wb_point_status.
template <int N>
__device__ inline bool wb_hash_in_history(u64 h, const u64* hist_hash, int moves, int lane)
{
    const int lim = moves < Geo<N>::MAXREC ? moves : Geo<N>::MAXREC;
    bool f = false;
    for (int i = 1 + lane; i < lim; i += 32) f |= (hist_hash[i] == h);
    return __any_sync(0xffffffffu, f);
}

// Full analysis of all points for `color`, 32 points (raster order) per step.  After super-ko has been
// resolved for the block, every lane calls block_done(base, idx, pos, legal, satari, eye) for its own point
// (idx = base + lane; idx >= N*N lanes pass legal = false), so the callback may use warp ballots.
template <int N, class BlockFn>
__device__ inline void wb_analyze(const WBoard<N>& b, WAnalysis<N>& an, const BScal& s, int color, bool superko, const u64* __restrict__ zob,
                                  const uint8_t* __restrict__ eye_lut, const u64* hist_hash, int lane, BlockFn&& block_done)
{
    using G = Geo<N>;
    wb_prepare_analysis(b, an, color, superko, zob, lane);
    for (int base = 0; base < G::NN; base += 32) {
        const int idx = base + lane;
        PointStatus st; st.legal_pre = false; st.need_scan = false; st.h = 0; st.satari = 0; st.eye = false;
        int pos = 0;
        if (idx < G::NN) { pos = onboard_pos<N>(idx); st = wb_point_status<N>(b, an, s, pos, color, superko, zob, EyeLutGlobal{eye_lut}); }
        bool legal = st.legal_pre;
        unsigned m = __ballot_sync(0xffffffffu, legal && st.need_scan);
        while (m) {
            const int src = __ffs(m) - 1; m &= m - 1;
            const u64 hq = shfl_u64(st.h, src);
            const bool found = wb_hash_in_history<N>(hq, hist_hash, s.moves, lane);
            if (lane == src && found) legal = false;
        }
        block_done(base, idx, pos, legal, st.satari, st.eye);
    }
}

// Leaf snapshot: everything nn/feature.py:10-57 reads from a board, packed for the feature-plane kernel.
//   bytes [0,2) previous move as raster index (-1: none or pass), [2] previous move was a pass,
//   [3] colour to move, [4,16) reserved, [16,16+NN) stone colours in raster order.
template <int N> struct Snap {
    static constexpr int HDR = 16;
    static constexpr int BYTES = (HDR + Geo<N>::NN + 15) & ~15;
};
template <int N>
__device__ inline void wb_snapshot(const WBoard<N>& b, const BScal& s, int color, const int16_t* hist_pos, uint8_t* out, int lane)
{
    using G = Geo<N>;
    const int prev = (s.moves - 1 < G::MAXREC) ? hist_pos[s.moves - 1] : 0;   // record.get(moves-1); slot 0 = PASS
    const bool prev_pass = (s.moves > 1 && prev == PASS);                      // feature.py:39-41
    if (lane == 0) {
        const int pidx = (prev == PASS) ? -1 : ((prev % G::W) - 1) + ((prev / G::W) - 1) * N;
        *reinterpret_cast<int16_t*>(out) = (int16_t)pidx;
        out[2] = prev_pass ? 1 : 0;
        out[3] = (uint8_t)color;
    }
    for (int idx = lane; idx < G::NN; idx += 32) out[Snap<N>::HDR + idx] = b.color[onboard_pos<N>(idx)];
}

// nn/feature.py:10-57 (sym 0): value of plane p at raster index idx for a snapshot.
template <int N>
__device__ __forceinline__ float snap_plane_value(const uint8_t* snap, int p, int idx)
{
    const int color = snap[3];
    if (p == 5) return color == WHITE ? -1.0f : 1.0f;                           // :50-52
    if (p == 4) return snap[2] ? 1.0f : 0.0f;                                   // :39-41
    if (p == 3) return (*reinterpret_cast<const int16_t*>(snap) == idx) ? 1.0f : 0.0f;   // :43-46
    int d = snap[Snap<N>::HDR + idx];
    if (color == WHITE && d != 0) d = 3 - d;                                    // :24-25
    return d == p ? 1.0f : 0.0f;                                                // :31
}

// GoBoard.count_score (go_board.py:561-608) with its raster-order, non-flood-fill colouring.
// Sequential by construction (each point sees earlier results), so lane 0 walks it; called once per game.
template <int N>
__device__ inline int wb_count_score(const WBoard<N>& b, uint8_t* tmp /*CP bytes of shared scratch*/, int lane)
{
    using G = Geo<N>;
    for (int c = lane; c < G::CELLS; c += 32) {
        int col = b.color[c];
        if ((col == BLACK || col == WHITE) && (b.ls[b.chain[c]] >> 16) == 1u) col = EMPTY;   // :570-573
        tmp[c] = (uint8_t)col;
    }
    __syncwarp();
    int score = 0;
    if (lane == 0) {
        for (int idx = 0; idx < G::NN; idx++) {
            const int pos = onboard_pos<N>(idx);
            if (tmp[pos] != EMPTY) continue;
            const int q[4] = { pos - G::W, pos - 1, pos + 1, pos + G::W };
            int col = EMPTY;
            for (int i = 0; i < 4; i++) {
                const int cc = tmp[q[i]];
                if (cc == BLACK || cc == WHITE) { if (col == EMPTY) col = cc; else if (col != cc) col = OB; }
            }
            tmp[pos] = (uint8_t)col;
        }
        for (int idx = 0; idx < G::NN; idx++) {
            const int cc = tmp[onboard_pos<N>(idx)];
            score += (cc == BLACK) - (cc == WHITE);
        }
    }
    __syncwarp();
    return __shfl_sync(0xffffffffu, score, 0);
}

// Tromp-Taylor area score, Black - White: stones plus the empty regions that reach one colour only --
// the adjudication the reference's get_final_status.py:15-64 obtains from GNU Go over a pipe, as a warp flood fill on the
// shared-memory board.  Regions are labelled by min-label propagation (every sweep lowers each empty point to the
// smallest label among its empty neighbours until nothing changes), then each region ORs the colours it touches.
//   lab: CP uint16 of scratch, reach: CP unsigned of scratch.
template <int N>
__device__ inline int wb_tromp_taylor(const WBoard<N>& b, uint16_t* lab, unsigned* reach, int lane)
{
    using G = Geo<N>;
    for (int c = lane; c < G::CP; c += 32) { lab[c] = (c < G::CELLS && b.color[c] == EMPTY) ? (uint16_t)c : (uint16_t)0xffff; reach[c] = 0; }
    __syncwarp();
    for (;;) {
        bool changed = false;
        for (int c = lane; c < G::CELLS; c += 32) {
            if (b.color[c] != EMPTY) continue;
            const int q[4] = { c - G::W, c - 1, c + 1, c + G::W };
            unsigned m = lab[c];
#pragma unroll
            for (int i = 0; i < 4; i++) if (b.color[q[i]] == EMPTY) m = min(m, (unsigned)lab[q[i]]);
            if (m < lab[c]) { lab[c] = (uint16_t)m; changed = true; }
        }
        __syncwarp();
        if (!__any_sync(0xffffffffu, changed)) break;
    }
    for (int c = lane; c < G::CELLS; c += 32) {
        if (b.color[c] != EMPTY) continue;
        const int q[4] = { c - G::W, c - 1, c + 1, c + G::W };
        unsigned r = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) { const int cc = b.color[q[i]]; if (cc == BLACK) r |= 1u; else if (cc == WHITE) r |= 2u; }
        if (r) atomicOr(&reach[lab[c]], r);
    }
    __syncwarp();
    int score = 0;
    for (int c = lane; c < G::CELLS; c += 32) {
        const int cc = b.color[c];
        if (cc == BLACK) score++;
        else if (cc == WHITE) score--;
        else if (cc == EMPTY) { const unsigned r = reach[lab[c]]; score += (r == 1u) - (r == 2u); }
    }
    __syncwarp();
    return warp_sum_i(score);
}

}  // namespace tg
