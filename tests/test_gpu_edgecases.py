"""Edge cases of the path on the B200, differential against the oracle (which is pinned to the reference goldens):
ko / super-ko / multi-stone captures / suicide / history limit / ragged and empty move lists on the board side;
one-candidate roots, late-game positions with few candidates, resignation and pass-pass endings on the search side."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _random_games(orc, size, ng, plies, seed, p_pass=0.02, p_any_legal=0.3, superko=True, zob=None):
    rs = np.random.RandomState(seed)
    zob = orc.default_zobrist(size) if zob is None else zob
    moves = np.zeros((ng, plies), np.int16)
    boards, stats = [], dict(ko=0, multi=0, superko_only=0, suicide=0)
    for k in range(ng):
        b = orc.OracleBoard(size, 7.0, superko, zob)
        color = orc.BLACK
        for i in range(plies):
            legal, sa, ey, cand = b.analyze(color)
            pool = np.flatnonzero(legal if rs.rand() < p_any_legal else cand)
            pos = 0 if len(pool) == 0 or rs.rand() < p_pass else b.onboard_pos[int(rs.choice(pool))]
            pris = b.state()["prisoner"]
            b.put_stone(pos, color)
            s = b.state()
            cap = sum(s["prisoner"]) - sum(pris)
            stats["multi"] += cap >= 2
            stats["ko"] += (s["ko_move"] == s["moves"] - 1 and pos != 0)
            moves[k, i] = pos
            color = 3 - color
        boards.append(b)
    return zob, moves, boards, stats


def _compare_final(e, d, boards, orc):
    for k, b in enumerate(boards):
        s = b.state()
        assert np.array_equal(d["color"][k, -1], s["color"]), k
        assert np.array_equal(d["libs"][k, -1], s["libs"]) and np.array_equal(d["size"][k, -1], s["size"]), k
        assert int(d["hash"][k, -1]) == s["hash"], k
        assert list(d["scal"][k, -1]) == [s["moves"], s["ko_pos"], s["ko_move"], *s["prisoner"]], k
        for ci, col in enumerate((orc.BLACK, orc.WHITE)):
            lm, sa, ey, cm = b.analyze(col)
            assert np.array_equal(d["legal"][k, -1, ci], lm), (k, col)
            assert np.array_equal(d["satari"][k, -1, ci], sa) and np.array_equal(d["eye"][k, -1, ci], ey), (k, col)
            assert np.array_equal(d["cand"][k, -1, ci], cm), (k, col)
        assert int(d["score"][k, -1]) == b.count_score(), k


def test_ko_superko_capture_events_small_board():
    """9x9 games that allow any legal move (eye fills, self-atari): many ko fights, multi-stone captures and positions
    where only the positional super-ko rule forbids a move.  Every ply is compared (hash, ko, prisoners, legality)."""
    import tamago_b200 as tb
    from oracle import oracle as orc
    size, ng, plies = 9, 48, 200
    zob, moves, boards, stats = _random_games(orc, size, ng, plies, seed=21, p_any_legal=0.6)
    assert stats["ko"] >= 10 and stats["multi"] >= 20
    e = tb.Engine(board_size=size, games=ng, max_visits=4, superko=True, evaluator=tb.EVAL_HASHNET)
    e.set_zobrist(zob)
    d = e.play(moves, dump=True)
    _compare_final(e, d, boards, orc)
    # per-ply replay of a few games on the oracle: hashes, ko state and both legality masks at every ply
    superko_only = 0
    for k in range(6):
        b = orc.OracleBoard(size, 7.0, True, zob)
        nb = orc.OracleBoard(size, 7.0, False, zob)          # same rules without super-ko, to count super-ko-only refusals
        color = 1
        for i in range(plies):
            b.put_stone(int(moves[k, i]), color); nb.put_stone(int(moves[k, i]), color); color = 3 - color
            s = b.state()
            assert int(d["hash"][k, i]) == s["hash"] and list(d["scal"][k, i][:3]) == [s["moves"], s["ko_pos"], s["ko_move"]]
            for ci, col in enumerate((1, 2)):
                lm = b.analyze(col)[0]
                assert np.array_equal(d["legal"][k, i, ci], lm), (k, i, col)
                superko_only += int((nb.analyze(col)[0] != lm).sum())
    assert superko_only > 0, "corpus never exercised the positional super-ko rule"
    e.close()


def test_history_limit_and_pass_heavy_games():
    """Games longer than MAX_RECORDS = 3 N^2 plies (record.py:39-44 drops the overflow) with many passes."""
    import tamago_b200 as tb
    from oracle import oracle as orc
    size, ng, plies = 9, 16, 3 * 81 + 20
    zob, moves, boards, _ = _random_games(orc, size, ng, plies, seed=5, p_pass=0.45)
    e = tb.Engine(board_size=size, games=ng, max_visits=4, superko=True, evaluator=tb.EVAL_HASHNET)
    e.set_zobrist(zob)
    d = e.play(moves, dump=True)
    _compare_final(e, d, boards, orc)
    e.close()


def test_ragged_and_empty_move_lists_19x19():
    import tamago_b200 as tb
    from oracle import oracle as orc
    size, ng, plies = 19, 6, 420
    zob, moves, boards, stats = _random_games(orc, size, ng, plies, seed=9, p_any_legal=0.5)
    counts = np.array([0, 1, 57, 200, 419, 420], np.int32)
    e = tb.Engine(board_size=size, games=ng, max_visits=4, superko=True, evaluator=tb.EVAL_HASHNET)
    e.set_zobrist(zob)
    d = e.play(moves, counts, dump=True)
    for k in range(ng):
        b = orc.OracleBoard(size, 7.0, True, zob)
        color = 1
        for i in range(counts[k]):
            b.put_stone(int(moves[k, i]), color); color = 3 - color
        if counts[k] == 0:
            continue
        s = b.state()
        i = counts[k] - 1
        assert np.array_equal(d["color"][k, i], s["color"]) and int(d["hash"][k, i]) == s["hash"]
        assert np.array_equal(d["legal"][k, i, color - 1], b.analyze(color)[0])
    # the untouched game is still the empty board: all points legal, planes of an empty position
    pl = e.planes()
    assert pl[0, 0].min() == 1.0 and pl[0, 1:5].max() == 0.0 and (pl[0, 5] == 1.0).all()
    e.close()


def test_late_game_searches_with_few_candidates_and_resignation():
    """Roots with 1..16 candidates exercise every sequential-halving schedule shape and the single-child
    PUCT shortcut (tree.py:76-77); never_resign = False exercises the resign rule (tree.py:347-354)."""
    import tamago_b200 as tb
    from oracle import oracle as orc
    size, ng = 9, 40
    rs = np.random.RandomState(17)
    zob = orc.default_zobrist(size)
    boards, mls = [], []
    for k in range(ng):
        b = orc.OracleBoard(size, 7.0, True, zob)
        color, ml = 1, []
        target = int(rs.randint(90, 150))
        for _ in range(target):
            cand = b.candidates(color)
            pos = int(rs.choice(cand[:-1])) if len(cand) > 1 else 0
            b.put_stone(pos, color); ml.append(pos); color = 3 - color
        boards.append((b, color)); mls.append(ml)
    ks = sorted(len(b.candidates(c)) for b, c in boards)
    assert ks[0] <= 3 and ks[-1] >= 8, ks
    mp = max(len(m) for m in mls)
    moves = np.zeros((ng, mp), np.int16)
    for k, m in enumerate(mls):
        moves[k, :len(m)] = m
    resigned = 0
    cases = [(0, 50, 123), (0, 16, 123), (1, 60, 123), (0, 5, 7), (1, 3, 8)]
    for kind, visits, seed in cases:
        e = tb.Engine(board_size=size, games=ng, max_visits=visits, superko=True, evaluator=tb.EVAL_HASHNET, seed=seed, dedup=True)
        e.set_zobrist(zob)
        e.play(moves, np.array([len(m) for m in mls], np.int32))
        res = e.genmove(mode=kind, visits=visits, play=False)
        for k, (b, color) in enumerate(boards):
            t = orc.OracleTree(size, orc.hashnet, tree_size=4096)
            t.set_noise_key(seed, k, b.moves)
            mv = t.genmove_sh(b, color, visits, False) if kind == 0 else t.genmove_puct(b, color, visits, False)
            assert res["error"][k] == 0
            assert res["move"][k] == mv, (kind, visits, k, len(b.candidates(color)))
            resigned += mv == -1               # (the hash evaluator rarely produces values below the resign threshold;
                                               #  the resign rule itself is exercised in test_resignation_rule)
            root, nd = t.node(0), e.node(k, 0)
            assert np.array_equal(nd["children_visits"], root["children_visits"]), (kind, visits, k)
            assert np.array_equal(nd["children_value_sum"], root["children_value_sum"]), (kind, visits, k)
            assert e.tree_size(k) == t.num_nodes
        e.close()
    assert resigned >= 0


def test_resignation_rule():
    """tree.py:347-354 / 100-103 and worker.py:60-63: a network that sees every position as lost for the side that just
    moved (value head biased to "side to move wins") makes the search resign unless never_resign is set."""
    import tamago_b200 as tb
    from tamago_b200.nn.utility import random_init_state_dict
    sd = random_init_state_dict(9, 0)
    sd["value_head.fc_layer.weight"] = np.zeros_like(sd["value_head.fc_layer.weight"])
    sd["value_head.fc_layer.bias"] = np.array([-10.0, 0.0, 10.0], np.float32)
    ng = 8
    never = np.array([0, 1] * (ng // 2), np.uint8)
    for mode, visits in ((tb.MODE_SH, 16), (tb.MODE_PUCT, 20)):
        e = tb.Engine(board_size=9, games=ng, max_visits=visits, evaluator=tb.EVAL_DUALNET_TC, seed=1)
        e.load_state_dict(sd)
        e.reset(never_resign=never)
        r = e.genmove(mode=mode, visits=visits, play=True)
        assert (r["error"] == 0).all()
        for g in range(ng):
            if mode == tb.MODE_SH and never[g]:
                assert r["move"][g] >= 0 and not r["finished"][g]          # never_resign honoured (tree.py:353)
            else:                                                          # PUCT ignores never_resign (tree.py:100-103)
                assert r["move"][g] == -1 and r["finished"][g] and r["resigned"][g] and r["winner"][g] == 2
        e.close()


def test_pass_pass_and_move_limit_endings():
    """worker.py:56-87 endings: two passes -> count_score - komi decides; 2 N^2 moves -> undecided."""
    import tamago_b200 as tb
    from oracle import oracle as orc
    size, ng, visits = 9, 6, 16
    e = tb.Engine(board_size=size, games=ng, max_visits=visits, superko=True, evaluator=tb.EVAL_HASHNET, seed=4)
    zob = orc.default_zobrist(size)
    e.set_zobrist(zob)
    e.reset(never_resign=np.ones(ng, np.uint8))
    boards = [orc.OracleBoard(size, 7.0, True, zob) for _ in range(ng)]
    done = np.zeros(ng, bool)
    for step in range(2 * size * size):
        r = e.genmove(mode=tb.MODE_SH, visits=visits, play=True)
        for k in range(ng):
            if done[k]:
                assert r["move"][k] == -2
                continue
            boards[k].put_stone(int(r["move"][k]), int(r["color"][k]))
            if r["finished"][k]:
                done[k] = True
                if boards[k].b.hist_pos[boards[k].moves - 1] == 0 and boards[k].b.hist_pos[boards[k].moves - 2] == 0:
                    score = boards[k].count_score() - 7.0
                    assert abs(float(r["score"][k]) - score) < 1e-6
                    assert int(r["winner"][k]) == (1 if score > 0.1 else 2 if score < -0.1 else 3)
                else:
                    assert step == 2 * size * size - 1 and int(r["winner"][k]) == 0
    assert done.all()
    e.close()


def test_tromp_taylor_scorer_on_device():
    """the device flood-fill area scorer (tg_config.scoring = 1; default stays count_score)
    against the oracle's flood fill on every ply of random games, and as the terminal score of finished self-play games."""
    import tamago_b200 as tb
    from oracle import oracle as orc
    for size, ng, plies in ((9, 32, 140), (19, 6, 500)):
        zob, moves, boards, _ = _random_games(orc, size, ng, plies, seed=31 + size, p_any_legal=0.4)
        e = tb.Engine(board_size=size, games=ng, max_visits=4, superko=True, evaluator=tb.EVAL_HASHNET)
        e.set_zobrist(zob)
        d = e.play(moves, dump=True)
        differ = 0
        for k in range(ng):
            b = orc.OracleBoard(size, 7.0, True, zob)
            color = 1
            for i in range(plies):
                b.put_stone(int(moves[k, i]), color); color = 3 - color
                if i % 5 == 4 or i == plies - 1:
                    assert int(d["tt_score"][k, i]) == b.tromp_taylor(), (size, k, i)
                    differ += int(d["tt_score"][k, i]) != int(d["score"][k, i])
        assert differ > 0                       # the two scorers are different functions of the position
        e.close()
    # terminal scoring of self-play games under both settings
    size, ng, visits = 9, 6, 16
    zob = orc.default_zobrist(size)
    for scoring in (0, 1):
        e = tb.Engine(board_size=size, games=ng, max_visits=visits, superko=True, evaluator=tb.EVAL_HASHNET, seed=4, scoring=scoring)
        e.set_zobrist(zob)
        e.reset(never_resign=np.ones(ng, np.uint8))
        boards = [orc.OracleBoard(size, 7.0, True, zob) for _ in range(ng)]
        done = np.zeros(ng, bool)
        for step in range(2 * size * size):
            r = e.genmove(mode=tb.MODE_SH, visits=visits, play=True)
            for k in range(ng):
                if done[k]:
                    continue
                boards[k].put_stone(int(r["move"][k]), int(r["color"][k]))
                if r["finished"][k]:
                    done[k] = True
                    if int(r["winner"][k]) != 0:
                        want = (boards[k].tromp_taylor() if scoring else boards[k].count_score()) - 7.0
                        assert abs(float(r["score"][k]) - want) < 1e-6, (scoring, k)
        assert done.all()
        e.close()
