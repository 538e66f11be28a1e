"""The C-ABI library loads and exports every symbol include/tamago_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "tamago_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    from tamago_b200 import build, _lib
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/tamago_b200.h but not exported"
    assert set(_lib.EXPORTS) == set(names), set(_lib.EXPORTS) ^ set(names)


def test_no_cpu_fallback():
    """Without a CUDA device the product path refuses to run instead of falling back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import tamago_b200
    with pytest.raises(RuntimeError, match="no CUDA device"):
        tamago_b200.Engine(board_size=9, games=1, max_visits=8)


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "tamago_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "tg_oracle" not in src and "from oracle" not in src and "import oracle" not in src, f


def test_sgf_writer_matches_reference_text(golden_dir):
    """tg_format_sgf against the SGF files the reference wrote (records re-parsed from the golden text)."""
    import numpy as np
    import tamago_b200
    g = np.load(os.path.join(golden_dir, "selfplay_9.npz"))
    n = int(g["size"])
    gtp_x = "IABCDEFGHJKLMNOPQRSTUVWXYZ"
    for text in g["sgf"]:
        text = str(text)
        re_field = re.search(r"RE\[([^\]]*)\]", text).group(1)
        komi = float(re.search(r"KM\[([^\]]*)\]", text).group(1))
        moves, colors, ks, acts, imps = [], [], [], [], []
        for col, xy, comment in re.findall(r";([BW])\[([a-z]{2})\]C\[([^\]]*)\]", text):
            pos = 0 if xy == "tt" else (ord(xy[0]) - 96) + (ord(xy[1]) - 96) * (n + 2)
            toks = comment.split()
            k = int(toks[0])
            a, p = np.zeros(96, np.int16), np.zeros(96, np.float64)
            for i, t in enumerate(toks[1:]):
                c, v = t.split(":")
                a[i] = 0 if c == "pass" else gtp_x.index(c[0]) + (n - int(c[1:]) + 1) * (n + 2)
                p[i] = float(v)
            moves.append(pos); colors.append(1 if col == "B" else 2); ks.append(k); acts.append(a); imps.append(p)
        if re_field == "0":
            winner, resigned, score = 3, 0, 0.0
        else:
            winner = 1 if re_field[0] == "B" else 2
            resigned = int(re_field.endswith("+R"))
            score = 0.0 if resigned else float(re_field[2:]) * (1 if winner == 1 else -1)
        out = tamago_b200.format_sgf(n, moves, colors, ks, np.array(acts), np.array(imps), winner, resigned, score, komi)
        assert out == text


def test_fast_policy_formatting_equals_printf():
    """The record writer formats ~10^7 improved-policy values per pool turn-over with its own "%.3e" (tg_record.cpp
    fmt_3e).  It must be byte-identical to the reference's f"{p:.3e}" (sgf/selfplay_record.py:61): random values over
    every magnitude the softmax produces, values next to rounding boundaries, exact ties, zeros and denormals."""
    import numpy as np
    import tamago_b200
    rs = np.random.RandomState(12)
    vals = [rs.uniform(0, 1, 60000), 10.0 ** rs.uniform(-30, 0, 60000), 10.0 ** rs.uniform(-320, -290, 2000),
            np.array([0.0, 1.0, 0.5, 0.25, 0.015625, 0.0009765625, 1e-18, 9.9995e-5, 9.99949999e-5, 0.99995, 0.999949999999,
                      1.2345e-3, 1.2355e-3, 5e-324, 1e-19, 1.0000000000000002e-19, 123456.0, 12345.0, 1e15, 2e15])]
    # values constructed to sit within a few ulp of a 4-digit rounding boundary
    d = rs.randint(1000, 10000, 40000) + 0.5
    e = rs.randint(-25, 0, 40000)
    near = d * 10.0 ** (e - 3.0)
    vals += [near, np.nextafter(near, 0), np.nextafter(near, 1), near * (1 + 1e-9), near * (1 - 1e-9)]
    x = np.concatenate(vals)
    per = 90
    rows = (len(x) + per - 1) // per
    pad = np.zeros(rows * per)
    pad[:len(x)] = x
    imp = np.zeros((rows, 96))
    imp[:, :per] = pad.reshape(rows, per)
    act = np.zeros((rows, 96), np.int16)
    text = tamago_b200.format_sgf(9, [0] * rows, [1] * rows, [per] * rows, act, imp, 3, 0, 0.0, 7.0)
    got = re.findall(r"pass:([^ \]]+)", text)
    assert len(got) == rows * per
    want = [f"{v:.3e}" for v in pad]
    bad = [(g, w) for g, w in zip(got, want) if g != w]
    assert not bad, bad[:5]
