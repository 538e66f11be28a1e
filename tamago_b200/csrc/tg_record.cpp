// tg_record.cpp -- self-play record writer (host side of the C ABI).
//
// Produces the SGF text of sgf/selfplay_record.py:45-110 byte for byte: header, RE/KM, one node per move with
// the improved-policy comment "k pos:prob ..." ("%.3e", GTP coordinates from board/coordinate.py:52-66, SGF
// coordinates from board/coordinate.py:68-82; PASS and RESIGN are written "tt").
#include "../../include/tamago_b200.h"
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <string>

static const char kGtpX[] = "IABCDEFGHJKLMNOPQRSTUVWXYZ";      // board/constant.py:28
static const char kSgf[] = "abcdefghijklmnopqrstuvwxyz";

static void gtp_coord(int n, int pos, char* out)
{
    if (pos == 0) { strcpy(out, "pass"); return; }
    if (pos < 0) { strcpy(out, "resign"); return; }
    const int w = n + 2, x = pos % w, y = n - (pos / w - 1);
    snprintf(out, 16, "%c%d", kGtpX[x], y);
}

// Python's repr(float) for the komi field: shortest round-trip digits, always with a fractional part.
static std::string py_float(double v)
{
    char buf[64];
    for (int prec = 1; prec <= 17; prec++) {
        snprintf(buf, sizeof buf, "%.*g", prec, v);
        if (strtod(buf, nullptr) == v) break;
    }
    std::string s(buf);
    if (s.find('.') == std::string::npos && s.find('e') == std::string::npos && s.find("inf") == std::string::npos && s.find("nan") == std::string::npos)
        s += ".0";
    return s;
}

// printf("%.3e") for the improved-policy comments, ~10x faster than snprintf: scale to [1000, 10000) with one exact
// power of ten, round, and fall back to snprintf whenever the scaled value is close enough to a rounding boundary for
// the single multiplication's rounding error (<= 2.3e-16 relative) to matter -- so the text is byte-identical to glibc's
// correctly rounded conversion (checked against snprintf over 10^7 values in tests/test_abi.py via tg_format_sgf).
static int fmt_3e(double x, char* out)
{
    static const double p10[23] = { 1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16,
                                    1e17, 1e18, 1e19, 1e20, 1e21, 1e22 };
    if (x == 0.0 && !std::signbit(x)) { memcpy(out, "0.000e+00", 9); return 9; }
    if (!(x > 1e-19 && x < 1e15)) return snprintf(out, 32, "%.3e", x);
    int e = (int)std::floor(std::log10(x));
    double scaled = (3 - e >= 0) ? x * p10[3 - e] : x / p10[e - 3];
    if (scaled < 1000.0) { e--; scaled = (3 - e >= 0) ? x * p10[3 - e] : x / p10[e - 3]; }
    else if (scaled >= 10000.0) { e++; scaled = (3 - e >= 0) ? x * p10[3 - e] : x / p10[e - 3]; }
    if (!(scaled >= 1000.0 && scaled < 10000.0)) return snprintf(out, 32, "%.3e", x);
    const double fl = std::floor(scaled), frac = scaled - fl;
    if (std::fabs(frac - 0.5) < 1e-6) return snprintf(out, 32, "%.3e", x);
    int d = (int)fl + (frac > 0.5 ? 1 : 0);
    if (d == 10000) { d = 1000; e++; }
    out[0] = (char)('0' + d / 1000); out[1] = '.';
    out[2] = (char)('0' + d / 100 % 10); out[3] = (char)('0' + d / 10 % 10); out[4] = (char)('0' + d % 10);
    out[5] = 'e'; out[6] = e < 0 ? '-' : '+';
    const int ae = e < 0 ? -e : e;
    int n = 7;
    if (ae >= 100) out[n++] = (char)('0' + ae / 100);
    out[n++] = (char)('0' + ae / 10 % 10); out[n++] = (char)('0' + ae % 10);
    return n;
}

// float(f"{x:.3e}"): the decimal round trip an improved-policy value takes through the SGF comment
double tg_round_3e(double x)
{
    char buf[40];
    const int n = fmt_3e(x, buf);
    buf[n] = 0;
    return strtod(buf, nullptr);
}

// the same for an array (entries equal to the literal 1e-18 default of nn/feature.py:90 are left alone)
extern "C" void tg_round_policy(double* p, int64_t n)
{
    for (int64_t i = 0; i < n; i++) if (p[i] != 1e-18) p[i] = tg_round_3e(p[i]);
}

template <class TM, class TC, class TK>
static void record_text(int n, int n_moves, const TM* moves, const TC* colors, const TK* num_children, const int16_t* action,
                        const double* improved, int stride, int winner, int resigned, double score, double komi, std::string& s)
{
    s.clear();
    s.reserve((size_t)n_moves * 1400 + 256);
    char buf[96];
    s += "(;FF[4]GM[1]SZ[" + std::to_string(n) + "]\n";
    s += "AP[TamaGo]PB[TamaGo-Black]PW[TamaGo-White]";
    if (winner == 1) {
        if (resigned) s += "RE[B+R]"; else { snprintf(buf, sizeof buf, "RE[B+%.1f]", score); s += buf; }
    } else if (winner == 2) {
        if (resigned) s += "RE[W+R]"; else { snprintf(buf, sizeof buf, "RE[W+%.1f]", -score); s += buf; }
    } else s += "RE[0]";
    s += "KM[" + py_float(komi) + "]";
    const int w = n + 2;
    // GTP names of every point once per game
    std::string names[31 * 31];
    char g[16];
    for (int i = 0; i < n_moves; i++) {
        const int pos = (int)moves[i];
        char xy[3] = {'t', 't', 0};
        if (pos > 0) { xy[0] = kSgf[pos % w - 1]; xy[1] = kSgf[pos / w - 1]; }
        s += (int)colors[i] == 1 ? ";B[" : ";W[";
        s += xy;
        s += "]C[";
        const int k = num_children ? (int)num_children[i] : 0;
        s += std::to_string(k);
        for (int c = 0; c < k; c++) {
            const int a = action[(size_t)i * stride + c];
            if (a >= 0 && a < w * w) {
                if (names[a].empty()) { gtp_coord(n, a, g); names[a] = g; }
                s += ' '; s += names[a]; s += ':';
            } else { gtp_coord(n, a, g); s += ' '; s += g; s += ':'; }
            s.append(buf, (size_t)fmt_3e(improved[(size_t)i * stride + c], buf));
        }
        s += "]";
    }
    s += "\n)";
}

// used by tg_engine.cu for the records fetched from the device ring
int tg_record_text(int n, int n_moves, const int16_t* moves, const uint8_t* colors, const int16_t* ks, const int16_t* action,
                   const double* improved, int stride, int winner, int resigned, double score, double komi, std::string& out)
{
    record_text(n, n_moves, moves, colors, ks, action, improved, stride, winner, resigned, score, komi, out);
    return 0;
}

extern "C" int tg_format_sgf(int32_t n, int32_t n_moves, const int32_t* moves, const int32_t* colors,
                             const int32_t* num_children, const int16_t* action, const double* improved, int32_t stride,
                             int32_t winner, int32_t resigned, double score, double komi, char* out, int32_t out_cap)
{
    if (n < 1 || n > 25 || n_moves < 0 || !out || out_cap < 1 || (n_moves > 0 && (!moves || !colors))) return TG_ERR_ARG;
    std::string s;
    record_text(n, n_moves, moves, colors, num_children, action, improved, stride, winner, resigned, score, komi, s);
    if ((int64_t)s.size() + 1 > out_cap) return TG_ERR_ARG;
    memcpy(out, s.c_str(), s.size() + 1);
    return (int)s.size();
}
