"""BASELINE.json's full-size configurations on the B200, checked through size-independent properties of the
search (the oracle cannot replay 4096 x 400 visits in seconds): visit conservation, schedule profiles, legality of the
chosen moves against the oracle's rules, determinism, and dedup == no-dedup."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _net(size, seed=0):
    from tamago_b200.nn.utility import random_init_state_dict
    return random_init_state_dict(size, seed)


def test_c2_9x9_4096_games_sh400_properties():
    import tamago_b200 as tb
    from oracle import oracle as orc
    games, visits = 4096, 400
    outs = []
    for dedup in (False, True):
        e = tb.Engine(board_size=9, games=games, max_visits=visits, evaluator=tb.EVAL_DUALNET_TC, dedup=dedup, seed=11)
        e.load_state_dict(_net(9))
        moves = []
        for step in range(3):
            r = e.genmove(mode=tb.MODE_SH, visits=visits, play=True)
            assert (r["error"] == 0).all()
            assert (r["visits"].sum(axis=1) == visits).all()                      # every simulation lands on one root child
            assert (r["evals"][0] == games * (visits + 1))                        # V+1 leaf evaluations per move
            if not dedup:
                assert r["evals"][1] == r["evals"][0]
            else:
                assert r["evals"][1] < r["evals"][0] // 4
            np.testing.assert_allclose(r["improved"].sum(axis=1), 1.0, rtol=1e-9)
            moves.append(r["move"].copy())
            if step == 0:
                # empty board, 82 children: the halving schedule {16:6, 8:12, 4:25, 2:54} gives this exact visit profile
                prof = np.sort(r["visits"][0][r["visits"][0] > 0])[::-1]
                assert list(prof[:2]) == [54, 54] and prof.sum() == 400 and (r["num_children"] == 82).all()
        outs.append(np.array(moves))
        # the moves played are legal for the oracle's rules
        b = [orc.OracleBoard(9, 7.0, True, orc.default_zobrist(9)) for _ in range(8)]
        for g in range(8):
            color = 1
            for step in range(3):
                mv = int(outs[-1][step, g])
                assert mv == 0 or b[g].is_legal(mv, color)
                b[g].put_stone(mv, color); color = 3 - color
        e.close()
    assert np.array_equal(outs[0], outs[1]), "dedup changed the moves"


def test_c4_19x19_1024_games_puct400_properties():
    import tamago_b200 as tb
    games, visits = 1024, 400
    e = tb.Engine(board_size=19, games=games, max_visits=visits, superko=True, evaluator=tb.EVAL_DUALNET_TC, seed=3)
    e.load_state_dict(_net(19))
    for step in range(2):
        r = e.genmove(mode=tb.MODE_PUCT, visits=visits, strict=False, play=True)
        assert (r["error"] == 0).all()
        tot = r["visits"].sum(axis=1)
        assert (tot <= visits).all() and (tot >= visits // 4).all()               # early stop (time_manager.py:146-163) may cut the budget
        best = r["visits"].argmax(axis=1)
        chosen = r["action"][np.arange(games), best]
        ok = (r["move"] == chosen) | (r["move"] == -1)
        assert ok.all()                                                           # node.py:169-175: first index of the max visit count
        for g in range(0, games, 97):
            root = e.node(g, 0)
            assert root["node_visits"] == root["children_visits"].sum() and root["virtual_loss"] == 0
            assert (root["children_virtual_loss"] == 0).all()
            assert e.tree_size(g) <= visits + 2
            psum = root["children_policy"].sum()     # PUCT roots carry the softmax policy of their candidates
            assert 0.9 < psum <= 1.0 + 1e-4          # (non-candidate points are dropped, so the sum is <= 1)
    e.close()


def test_c5_19x19_genmove_1600_visits_batch256():
    import tamago_b200 as tb
    e = tb.Engine(board_size=19, games=1, max_visits=1600, batch_size=256, superko=False, evaluator=tb.EVAL_DUALNET_TC, seed=5)
    e.load_state_dict(_net(19))
    r1 = e.genmove(mode=tb.MODE_PUCT, visits=1600, strict=True, play=False)
    assert r1["error"][0] == 0 and r1["visits"][0].sum() == 1600
    root = e.node(0, 0)
    assert root["node_visits"] == 1600 and (root["children_virtual_loss"] == 0).all()
    r2 = e.genmove(mode=tb.MODE_PUCT, visits=1600, strict=True, play=False)       # same position, same seed: deterministic
    assert np.array_equal(r1["visits"], r2["visits"]) and r1["move"][0] == r2["move"][0]
    e.close()


def test_c3_9x9_sh50_many_games():
    import tamago_b200 as tb
    games, visits = 16384, 50
    e = tb.Engine(board_size=9, games=games, max_visits=visits, evaluator=tb.EVAL_DUALNET_TC, dedup=True, seed=2)
    e.load_state_dict(_net(9))
    r = e.genmove(mode=tb.MODE_SH, visits=visits, play=True)
    assert (r["error"] == 0).all() and (r["visits"].sum(axis=1) == visits).all()
    prof = np.sort(r["visits"][5][r["visits"][5] > 0])[::-1]
    assert list(prof[:2]) == [7, 7]                                              # {16:1, 8:1, 4:3, 2:7}
    e.close()


@pytest.mark.parametrize("size,mode,visits,batch", [(9, 0, 50, 1), (19, 1, 48, 8), (13, 0, 16, 1)])
def test_fused_feature_planes_equal_the_plane_kernel(monkeypatch, size, mode, visits, batch):
    """The tensor-core evaluator builds its input planes from the leaf snapshots (fused plane load, tg_dualnet.cuh); with
    TG_UNFUSED_PLANES=1 it reads the fp32 planes k_planes writes (the kernel pinned to nn/feature.py by the goldens).
    Same planes => bit-identical searches: moves, root visits, improved policies, after several moves into the game
    (white to move, previous-move and pass planes in use)."""
    import tamago_b200 as tb
    outs = []
    for unfused in (True, False):
        if unfused:
            monkeypatch.setenv("TG_UNFUSED_PLANES", "1")
        else:
            monkeypatch.delenv("TG_UNFUSED_PLANES", raising=False)
        e = tb.Engine(board_size=size, games=24, max_visits=visits, batch_size=batch, evaluator=tb.EVAL_DUALNET_TC, dedup=False, seed=9)
        e.load_state_dict(_net(size, 3))
        e.reset(never_resign=np.ones(24, np.uint8))
        hist = []
        for step in range(7):
            r = e.genmove(mode=mode, visits=visits, play=True)
            assert (r["error"] == 0).all()
            hist.append((r["move"].copy(), r["visits"].copy(), r["improved"].copy()))
        outs.append(hist)
        e.close()
    for a, b in zip(*outs):
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
