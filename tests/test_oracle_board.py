"""T0: the C oracle against golden board states recorded from the reference.

Golden files come from tests/golden/make_golden.py (runs the unmodified
reference: board/go_board.py, board/string.py, board/pattern.py, nn/feature.py).
"""
import os

import numpy as np
import pytest

from oracle import oracle as orc


def _replay(path, check):
    g = dict(np.load(path))
    n = int(g["size"])
    zob = g["zobrist"]
    legal = np.unpackbits(g["legal"], axis=-1)[..., : n * n]
    cand = np.unpackbits(g["cand"], axis=-1)[..., : n * n]
    eye = np.unpackbits(g["eye"], axis=-1)[..., : n * n]
    b, cur = None, -1
    for i in range(len(g["pos"])):
        if g["game"][i] != cur:
            cur = g["game"][i]
            b = orc.OracleBoard(n, 7.0, True, zob)
        b.put_stone(int(g["pos"][i]), int(g["mover"][i]))
        check(b, g, i, legal[i], cand[i], eye[i])
    return len(g["pos"])


@pytest.mark.parametrize("size", [9, 19])
def test_board_state_matches_reference(golden_dir, size):
    path = os.path.join(golden_dir, f"board_{size}.npz")
    if not os.path.isfile(path):
        pytest.skip("golden file not generated")

    def check(b, g, i, legal, cand, eye):
        s = b.state()
        assert np.array_equal(s["color"], g["color"][i]), f"ply {i} colours"
        assert np.array_equal(s["libs"], g["libs"][i]), f"ply {i} liberties"
        assert np.array_equal(s["size"], g["size_pt"][i]), f"ply {i} string sizes"
        assert [s["moves"], s["ko_pos"], s["ko_move"], *s["prisoner"]] == list(g["scal"][i]), f"ply {i} scalars"
        assert s["hash"] == int(g["hash"][i]), f"ply {i} hash"
        assert b.count_score() == int(g["score"][i]), f"ply {i} count_score"
        for ci, col in enumerate((orc.BLACK, orc.WHITE)):
            lm, sa, ey, cm = b.analyze(col)
            assert np.array_equal(lm, legal[ci]), f"ply {i} legal colour {col}"
            assert np.array_equal(cm, cand[ci]), f"ply {i} candidates colour {col}"
            assert np.array_equal(sa, g["satari"][i][ci]), f"ply {i} self-atari colour {col}"
            assert np.array_equal(ey, eye[ci]), f"ply {i} complete-eye colour {col}"
            if i % 16 == 0:
                cl = b.candidates(col)
                onb = np.array(b.onboard_pos)
                assert cl[-1] == 0 and np.array_equal(cl[:-1], onb[cm.astype(bool)]), f"ply {i} candidate list"
        to_move = 3 - int(g["mover"][i])
        pl = b.planes(to_move)
        w = np.arange(1, pl.size + 1, dtype=np.float64)
        assert float((pl.reshape(-1).astype(np.float64) * w).sum()) == g["planes_sum"][i], f"ply {i} planes"

    assert _replay(path, check) > 100


@pytest.mark.parametrize("size", [9, 19])
def test_planes_exact(golden_dir, size):
    path = os.path.join(golden_dir, f"board_{size}.npz")
    if not os.path.isfile(path):
        pytest.skip("golden file not generated")
    g = dict(np.load(path))
    want = {int(p): g["planes"][k] for k, p in enumerate(g["planes_ply"])}
    seen = []

    def check(b, g, i, *_):
        if i in want:
            assert np.array_equal(b.planes(3 - int(g["mover"][i])), want[i])
            seen.append(i)

    _replay(path, check)
    assert len(seen) == len(want) > 0


def test_eye_table(golden_dir):
    eye = np.load(os.path.join(golden_dir, "eye_table.npz"))["eye"]
    mine = np.ctypeslib.as_array(orc.lib().tgo_eye_table(), shape=(65536,))
    assert np.array_equal(mine, eye)
    assert int((eye != 0).sum()) > 100
