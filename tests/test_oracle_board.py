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


@pytest.mark.parametrize("size", [9, 13, 19])
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


@pytest.mark.parametrize("size", [9, 13, 19])
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


def test_ply_digest_numpy_twin_and_bulk_corpus():
    """The bulk differential corpus (oracle.random_games): the C digest of a ply equals its numpy twin applied to the
    exported state, so the GPU test can digest the engine's dump with numpy and compare 10^4 games ply by ply."""
    n = 9
    zob = orc.default_zobrist(n)
    mv, cnt, dig = orc.random_games(n, zob, seed=5, games=6, max_plies=160, p_any_legal=0.5)
    assert cnt.min() > 20
    for g in (0, 5):
        b = orc.OracleBoard(n, 7.0, True, zob)
        color = orc.BLACK
        keys = {k: [] for k in ("color", "libs", "size", "scal", "hash", "legal", "satari", "eye", "cand", "score")}
        for i in range(cnt[g]):
            b.put_stone(int(mv[g, i]), color)
            color = 3 - color
            s = b.state()
            keys["color"].append(s["color"]); keys["libs"].append(s["libs"]); keys["size"].append(s["size"])
            keys["scal"].append([s["moves"], s["ko_pos"], s["ko_move"], *s["prisoner"]]); keys["hash"].append(np.uint64(s["hash"]))
            a = [b.analyze(c) for c in (orc.BLACK, orc.WHITE)]
            for j, k in enumerate(("legal", "satari", "eye", "cand")):
                keys[k].append([x[j] for x in a])
            keys["score"].append(b.count_score())
            assert b.ply_digest() == int(dig[g, i])
        d = {k: np.array(v) for k, v in keys.items()}
        assert np.array_equal(orc.ply_digest_np(d, n), dig[g, :cnt[g]])
    # the digest is sensitive to every field
    d2 = {k: v.copy() for k, v in d.items()}
    d2["libs"][3, 40] += 1
    assert orc.ply_digest_np(d2, n)[3] != dig[g, 3]


def _tromp_taylor_py(color, n):
    """independent 20-line flood fill: area score Black - White"""
    w = n + 2
    seen, score = set(), 0
    for y in range(1, n + 1):
        for x in range(1, n + 1):
            p = x + y * w
            if color[p] == 1:
                score += 1
            elif color[p] == 2:
                score -= 1
            elif p not in seen:
                region, reach, todo = 0, set(), [p]
                seen.add(p)
                while todo:
                    q = todo.pop()
                    region += 1
                    for r in (q - w, q - 1, q + 1, q + w):
                        if color[r] in (1, 2):
                            reach.add(int(color[r]))
                        elif color[r] == 0 and r not in seen:
                            seen.add(r); todo.append(r)
                if reach == {1}:
                    score += region
                elif reach == {2}:
                    score -= region
    return score


@pytest.mark.parametrize("size", [9, 19])
def test_tromp_taylor_oracle(golden_dir, size):
    """the oracle's area scorer against an independent Python flood fill on every ply of the golden games,
    plus how often it agrees with the reference's count_score (go_board.py:561-608, not a flood fill: SURVEY A.3 Q9)."""
    agree = total = 0

    def check(b, g, i, *_):
        nonlocal agree, total
        if i % 3:
            return
        tt = b.tromp_taylor()
        assert tt == _tromp_taylor_py(b.state()["color"], size), f"ply {i}"
        agree += tt == int(g["score"][i]); total += 1

    _replay(os.path.join(golden_dir, f"board_{size}.npz"), check)
    assert total > 50 and 0 < agree < total      # the two scorers differ on open positions and agree on settled ones
