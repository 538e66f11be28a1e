// tg_record.cpp -- self-play record writer (host side of the C ABI).
//
// Produces the SGF text of sgf/selfplay_record.py:45-110 byte for byte: header, RE/KM, one node per move with
// the improved-policy comment "k pos:prob ..." ("%.3e", GTP coordinates from board/coordinate.py:52-66, SGF
// coordinates from board/coordinate.py:68-82; PASS and RESIGN are written "tt").
#include "../../include/tamago_b200.h"
#include <cstdio>
#include <cstdlib>
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

extern "C" int tg_format_sgf(int32_t n, int32_t n_moves, const int32_t* moves, const int32_t* colors,
                             const int32_t* num_children, const int16_t* action, const double* improved, int32_t stride,
                             int32_t winner, int32_t resigned, double score, double komi, char* out, int32_t out_cap)
{
    if (n < 1 || n > 25 || n_moves < 0 || !out || out_cap < 1 || (n_moves > 0 && (!moves || !colors))) return TG_ERR_ARG;
    std::string s;
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
    for (int i = 0; i < n_moves; i++) {
        const int pos = moves[i];
        char xy[3] = {'t', 't', 0};
        if (pos > 0) { xy[0] = kSgf[pos % w - 1]; xy[1] = kSgf[pos / w - 1]; }
        s += colors[i] == 1 ? ";B[" : ";W[";
        s += xy;
        s += "]C[";
        const int k = num_children ? num_children[i] : 0;
        s += std::to_string(k);
        for (int c = 0; c < k; c++) {
            char g[16];
            gtp_coord(n, action[(size_t)i * stride + c], g);
            snprintf(buf, sizeof buf, " %s:%.3e", g, improved[(size_t)i * stride + c]);
            s += buf;
        }
        s += "]";
    }
    s += "\n)";
    if ((int64_t)s.size() + 1 > out_cap) return TG_ERR_ARG;
    memcpy(out, s.c_str(), s.size() + 1);
    return (int)s.size();
}
