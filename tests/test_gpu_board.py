"""T0/T1 on the B200: board kernels through the C ABI vs golden states of the reference and vs the oracle.

Covers GoBoard.put_stone / is_legal (incl. super-ko) / check_self_atari_stone / is_complete_eye /
count_score (board/go_board.py), string bookkeeping (board/string.py), Zobrist hashing, and
generate_input_planes (nn/feature.py).  Everything is compared bit for bit.
"""
import os

import numpy as np
import pytest

from gpu_util import pack_games

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tb():
    import tamago_b200
    return tamago_b200


@pytest.mark.parametrize("size", [9, 13, 19])
def test_board_states_match_reference_golden(golden_dir, tb, size):
    g = dict(np.load(os.path.join(golden_dir, f"board_{size}.npz")))
    moves, colors, counts, start = pack_games(g)
    ng = len(counts)
    nn = size * size
    e = tb.Engine(board_size=size, games=ng, max_visits=8, superko=True, evaluator=tb.EVAL_HASHNET)
    e.set_zobrist(g["zobrist"])
    d = e.play(moves, counts, colors, dump=True)
    legal = np.unpackbits(g["legal"], axis=-1)[..., :nn]
    cand = np.unpackbits(g["cand"], axis=-1)[..., :nn]
    eye = np.unpackbits(g["eye"], axis=-1)[..., :nn]
    checked = 0
    for k in range(ng):
        s, c = start[k], counts[k]
        sl = slice(s, s + c)
        for name, got, want in (("color", d["color"][k, :c], g["color"][sl]), ("libs", d["libs"][k, :c], g["libs"][sl]),
                                ("size", d["size"][k, :c], g["size_pt"][sl]), ("scal", d["scal"][k, :c], g["scal"][sl]),
                                ("hash", d["hash"][k, :c], g["hash"][sl]), ("legal", d["legal"][k, :c], legal[sl]),
                                ("satari", d["satari"][k, :c], g["satari"][sl]), ("eye", d["eye"][k, :c], eye[sl]),
                                ("cand", d["cand"][k, :c], cand[sl]), ("score", d["score"][k, :c], g["score"][sl])):
            if not np.array_equal(got, want):
                bad = np.argwhere(np.asarray(got) != np.asarray(want))[0]
                raise AssertionError(f"size {size} game {k}: {name} differs first at ply/idx {bad.tolist()}: "
                                     f"got {np.asarray(got)[tuple(bad)]}, want {np.asarray(want)[tuple(bad)]}")
        checked += c
    assert checked == len(g["pos"]) > 100
    e.close()


@pytest.mark.parametrize("size", [9, 13, 19])
def test_planes_match_reference_golden(golden_dir, tb, size):
    """Planes after a prefix of each golden game, against the sampled reference planes and the oracle for every game."""
    from oracle import oracle as orc
    g = dict(np.load(os.path.join(golden_dir, f"board_{size}.npz")))
    moves, colors, counts, start = pack_games(g)
    ng = len(counts)
    want = {int(p): g["planes"][i] for i, p in enumerate(g["planes_ply"])}
    # choose, per game, a sampled ply if it has one, else the middle of the game
    cut = np.zeros(ng, np.int32)
    for k in range(ng):
        plies = [p for p in want if start[k] <= p < start[k] + counts[k]]
        cut[k] = (plies[0] - start[k] + 1) if plies else max(1, counts[k] // 2)
    e = tb.Engine(board_size=size, games=ng, max_visits=8, superko=True, evaluator=tb.EVAL_HASHNET)
    e.set_zobrist(g["zobrist"])
    e.play(moves, cut, colors)
    to_move = np.array([3 - colors[k, cut[k] - 1] for k in range(ng)], np.int32)
    e.set_to_move(to_move)
    pl = e.planes()
    hits = 0
    for k in range(ng):
        ob = orc.OracleBoard(size, 7.0, True, g["zobrist"])
        for i in range(cut[k]):
            ob.put_stone(int(moves[k, i]), int(colors[k, i]))
        assert np.array_equal(pl[k], ob.planes(int(to_move[k]))), f"game {k} vs oracle"
        gi = int(start[k] + cut[k] - 1)
        if gi in want:
            assert np.array_equal(pl[k], want[gi]), f"game {k} vs reference golden"
            hits += 1
    assert hits >= 1
    e.close()


def test_random_games_vs_oracle(tb):
    """Many random legal games on the device vs the oracle: final state and per-ply hashes (9x9, super-ko on)."""
    from oracle import oracle as orc
    rs = np.random.RandomState(11)
    size, ng, plies = 9, 64, 150
    zob = orc.default_zobrist(size)
    moves = np.zeros((ng, plies), np.int16)
    boards = []
    for k in range(ng):
        b = orc.OracleBoard(size, 7.0, True, zob)
        color = orc.BLACK
        for i in range(plies):
            legal, sa, ey, cand = b.analyze(color)
            idx = np.flatnonzero(cand if rs.rand() > 0.2 else legal)
            pos = 0 if len(idx) == 0 or rs.rand() < 0.02 else b.onboard_pos[int(rs.choice(idx))]
            b.put_stone(pos, color)
            moves[k, i] = pos
            color = 3 - color
        boards.append(b)
    e = tb.Engine(board_size=size, games=ng, max_visits=8, superko=True, evaluator=tb.EVAL_HASHNET)
    e.set_zobrist(zob)
    d = e.play(moves, dump=True)
    for k, b in enumerate(boards):
        s = b.state()
        assert np.array_equal(d["color"][k, -1], s["color"])
        assert np.array_equal(d["libs"][k, -1], s["libs"]) and np.array_equal(d["size"][k, -1], s["size"])
        assert int(d["hash"][k, -1]) == s["hash"]
        assert list(d["scal"][k, -1]) == [s["moves"], s["ko_pos"], s["ko_move"], *s["prisoner"]]
        for ci, col in enumerate((orc.BLACK, orc.WHITE)):
            lm, sa, ey, cm = b.analyze(col)
            assert np.array_equal(d["legal"][k, -1, ci], lm) and np.array_equal(d["cand"][k, -1, ci], cm)
            assert np.array_equal(d["satari"][k, -1, ci], sa) and np.array_equal(d["eye"][k, -1, ci], ey)
        assert int(d["score"][k, -1]) == b.count_score()
    e.close()


def _corpus_chunk(args):
    from oracle import oracle as orc
    size, seed, first, games, plies = args
    return orc.random_games(size, orc.default_zobrist(size), seed, games, plies, p_pass=0.02, p_any_legal=0.35, first_game=first)


@pytest.mark.parametrize("size,games,chunk", [(9, 10240, 2048), (19, 1024, 256)])
def test_bulk_random_game_corpus_vs_oracle(tb, size, games, chunk):
    """SURVEY 7 step 3 at scale: >= 10^4 9x9 and >= 10^3 19x19 random games, EVERY ply compared through a 64-bit digest of
    the whole observable state (colours, liberties, sizes, ko, prisoners, hash, legality / self-atari / eye / candidate
    masks of both colours, count_score).  The oracle plays the games (C, one process per host core) and digests each
    ply; the engine replays the moves and its per-ply dump is digested with the numpy twin."""
    import multiprocessing as mp
    from oracle import oracle as orc
    orc.build()
    plies = 2 * size * size
    procs = max(1, min(os.cpu_count() or 1, 32))
    per = max(1, chunk // procs)
    e = tb.Engine(board_size=size, games=chunk, max_visits=4, superko=True, evaluator=tb.EVAL_HASHNET)
    e.set_zobrist(orc.default_zobrist(size))
    total_plies = 0
    with mp.get_context("fork").Pool(procs) as pool:
        for c0 in range(0, games, chunk):
            jobs = [(size, 77, c0 + j, min(per, chunk - j), plies) for j in range(0, chunk, per)]
            parts = pool.map(_corpus_chunk, jobs)
            moves = np.concatenate([p[0] for p in parts]); counts = np.concatenate([p[1] for p in parts])
            dig = np.concatenate([p[2] for p in parts])
            assert len(moves) == chunk
            e.reset()
            d = e.play(moves, counts, dump=True)
            got = orc.ply_digest_np(d, size)
            valid = np.arange(plies)[None, :] < counts[:, None]
            bad = np.argwhere((got != dig) & valid)
            assert len(bad) == 0, f"size {size}: first mismatch at game {c0 + bad[0][0]} ply {bad[0][1]} of {len(bad)}"
            total_plies += int(counts.sum())
    assert total_plies > games * size * 4
    e.close()
