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
