"""T3 at the full BASELINE budgets on the B200: trees of the CUDA search against the digest goldens recorded from the
reference MCTSTree (tests/golden/search2_<N>.npz) -- non-dyadic evaluator (fp32 queue-order accumulation observable),
13x13, 19x19 PUCT-400 batch 1 / 8 with super-ko (configs[3]), SH-400, PUCT-1600 with 256-leaf batches (configs[4]) --
and whole 19x19 PUCT-400 games replayed move for move against the oracle."""
import os
from collections import defaultdict

import numpy as np
import pytest

from golden_util import DigestGolden, compare_digest_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("size", [9, 13, 19])
@pytest.mark.parametrize("dedup", [False, True])
def test_trees_match_digest_goldens(golden_dir, size, dedup):
    import tamago_b200 as tb
    dg = DigestGolden(os.path.join(golden_dir, f"search2_{size}.npz"))
    groups = defaultdict(list)
    for i in range(dg.ncases):
        meta, ref = dg.case(i)
        groups[(meta["kind"], meta["visits"], meta["batch"], meta["evaluator"], meta["strict"])].append((meta, ref))
    for (kind, visits, batch, ev, strict), cases in groups.items():
        if dedup and kind == 1 and visits >= 1600:
            continue                                     # same code path as the non-dedup run; keep the GPU time for the rest
        ng = len(cases)
        e = tb.Engine(board_size=size, games=ng, max_visits=visits, superko=True, batch_size=batch,
                      evaluator=tb.EVAL_HASHNET2 if ev else tb.EVAL_HASHNET, dedup=dedup, seed=dg.seed)
        e.set_zobrist(dg.zobrist)
        e.reset(game_ids=[c[0]["pos_index"] for c in cases])
        mls = [dg.movelist(c[0]["pos_index"]) for c in cases]
        mp = max(1, max(len(m) for m in mls))
        moves = np.zeros((ng, mp), np.int16)
        for k, m in enumerate(mls):
            moves[k, :len(m)] = m
        e.play(moves, np.array([len(m) for m in mls], np.int32))
        res = e.genmove(mode=kind, visits=visits, strict=bool(strict), play=False)
        for k, (meta, ref) in enumerate(cases):
            tag = f"size={size} kind={kind} pos={meta['pos_index']} visits={visits} batch={batch} ev={ev} strict={strict} dedup={dedup}"
            assert res["error"][k] == 0, tag
            compare_digest_case(lambda ni: e.node(k, ni), e.tree_size(k), int(res["move"][k]), meta, ref, tag)
            if kind == 0:
                kk = len(ref["improved"])
                np.testing.assert_allclose(res["improved"][k, :kk], ref["improved"], rtol=1e-12, atol=1e-300, err_msg=tag)
        e.close()


def _oracle_game(args):
    from oracle import oracle as orc
    size, seed, game, visits, batch = args
    t = orc.OracleTree(size, orc.hashnet2, tree_size=4096, batch_size=batch)
    r = t.selfplay_game(7.0, orc.default_zobrist(size), seed=seed, game=game, visits=visits, never_resign=False, use_puct=True)
    return r["pos"], r["winner"], r["is_resign"], r["score"]


@pytest.mark.parametrize("batch,warp", [(1, False), (8, False), (1, True)])
def test_c4_full_games_19x19_puct400_vs_oracle(batch, warp, monkeypatch):
    """BASELINE configs[3] played to the end: 8 whole 19x19 games, 400-visit PUCT + super-ko with the early stop of
    is_move_decided, every move / resignation / final score equal to the oracle's game (non-dyadic evaluator)."""
    import multiprocessing as mp
    import tamago_b200 as tb
    from oracle import oracle as orc
    orc.build()
    size, ng, visits, seed = 19, 8, 400, 41
    if warp:                                             # the kernels configs[3] itself runs (1024 games): one warp per game, board snapshots
        monkeypatch.setenv("TG_PUCT_WARP", "1")
    with mp.get_context("fork").Pool(min(ng, os.cpu_count() or 1)) as pool:
        fut = pool.map_async(_oracle_game, [(size, seed, g, visits, batch) for g in range(ng)])
        e = tb.Engine(board_size=size, games=ng, max_visits=visits, superko=True, batch_size=batch, evaluator=tb.EVAL_HASHNET2, seed=seed)
        e.set_zobrist(orc.default_zobrist(size))
        e.reset(game_ids=np.arange(ng))
        played = [[] for _ in range(ng)]
        final = [None] * ng
        for _ in range(2 * size * size):
            r = e.genmove(mode=tb.MODE_PUCT, visits=visits, strict=False, play=True)
            for g in range(ng):
                if final[g] is not None:
                    continue
                assert r["error"][g] == 0
                if r["move"][g] >= 0:
                    played[g].append(int(r["move"][g]))
                if r["finished"][g]:
                    final[g] = (int(r["winner"][g]), int(r["resigned"][g]), float(r["score"][g]))
            if all(f is not None for f in final):
                break
        e.close()
        want = fut.get(timeout=1200)
    total = 0
    for g in range(ng):
        pos, winner, is_resign, score = want[g]
        assert played[g] == [int(p) for p in pos], f"game {g}: first difference at move " \
            f"{next((i for i, (a, b) in enumerate(zip(played[g], pos)) if a != b), min(len(played[g]), len(pos)))}"
        assert final[g][0] == winner and final[g][1] == int(is_resign) and abs(final[g][2] - score) < 1e-6, g
        total += len(pos)
    assert total > ng * 60


def test_puct_kernel_variants_agree(monkeypatch):
    """The PUCT descent/backup kernels (warp per game; CTA per game with 256 or 512 threads, inline or deferred expansion) are bit-identical:
    160 positions (-> 256-thread CTAs by default), 100 (-> 512-thread CTAs) and the warp kernels forced by TG_PUCT_WARP,
    batch 1 and batch 16 (tentative priors, duplicate leaves), compared on every root statistic; a sample against the oracle."""
    import tamago_b200 as tb
    from oracle import oracle as orc
    size = 13
    rs = np.random.RandomState(23)
    zob = orc.default_zobrist(size)
    boards, mls = [], []
    for k in range(160):
        b = orc.OracleBoard(size, 7.0, True, zob)
        color, ml = orc.BLACK, []
        for _ in range(int(rs.randint(0, 150))):
            cand = b.candidates(color)
            pos = int(rs.choice(cand[:-1])) if len(cand) > 1 and rs.rand() > 0.03 else 0
            b.put_stone(pos, color); ml.append(pos); color = 3 - color
        boards.append((b, color)); mls.append(ml)

    def run(ng, batch, visits, dedup, warp, env=()):
        for key in ("TG_PUCT_WARP", "TG_PUCT_DEFER", "TG_WALK_SLOTS", "TG_PUCT_WAVE", "TG_WAVE_GT", "TG_PUCT_NOSNAP"):
            monkeypatch.delenv(key, raising=False)
        if warp:
            monkeypatch.setenv("TG_PUCT_WARP", "1")
        for key, val in env:
            monkeypatch.setenv(key, val)
        e = tb.Engine(board_size=size, games=ng, max_visits=visits, superko=True, batch_size=batch, evaluator=tb.EVAL_HASHNET2,
                      dedup=dedup, seed=3)
        e.set_zobrist(zob)
        mp = max(1, max(len(m) for m in mls[:ng]))
        moves = np.zeros((ng, mp), np.int16)
        for k, m in enumerate(mls[:ng]):
            moves[k, :len(m)] = m
        e.play(moves, np.array([len(m) for m in mls[:ng]], np.int32))
        res = e.genmove(mode=tb.MODE_PUCT, visits=visits, play=False)
        assert (res["error"] == 0).all()
        roots = [e.node(k, 0) for k in range(0, ng, 7)]
        sizes = [e.tree_size(k) for k in range(ng)]
        e.close()
        return res, roots, sizes

    for batch, visits, dedup in ((1, 48, False), (16, 90, True), (16, 90, False)):
        ref = run(160, batch, visits, dedup, warp=True)
        if batch == 1:                                   # warp kernels: board snapshots along the previous path (default) or full replays
            alt = run(160, batch, visits, dedup, warp=True, env=(("TG_PUCT_NOSNAP", "1"),))
            assert np.array_equal(alt[0]["move"], ref[0]["move"]) and np.array_equal(alt[0]["visits"], ref[0]["visits"]) and alt[2] == ref[2]
            for a, b in zip(alt[1], ref[1]):
                for key in ("children_visits", "children_value_sum", "children_policy", "children_index", "children_virtual_loss"):
                    assert np.array_equal(a[key], b[key]), ("snapshots", key)
        # block kernels: 256 / 512 threads; batches > 1 on <= 148 games split into tree walk + deferred expansion by default,
        # forced on / off here, with the full node-row cache, a three-slot cache (evictions) and none
        # and the walk either pipelined over the descents (wavefront; not with leaf deduplication) or sequential
        variants = [(160, ()), (100, ())]
        if batch > 1:
            variants += [(160, (("TG_PUCT_DEFER", "1"),)), (100, (("TG_PUCT_DEFER", "0"),)), (100, (("TG_WAVE_GT", "64"),)),
                         (100, (("TG_PUCT_WAVE", "0"),)), (100, (("TG_PUCT_WAVE", "0"), ("TG_WALK_SLOTS", "3"))),
                         (100, (("TG_PUCT_WAVE", "0"), ("TG_WALK_SLOTS", "2")))]
        for ng, env in variants:
            got = run(ng, batch, visits, dedup, warp=False, env=env)
            assert np.array_equal(got[0]["move"], ref[0]["move"][:ng]) and np.array_equal(got[0]["visits"], ref[0]["visits"][:ng])
            assert got[2] == ref[2][:ng]
            for a, b in zip(got[1], ref[1]):
                for key in ("children_visits", "children_value_sum", "children_policy", "children_index", "children_virtual_loss"):
                    assert np.array_equal(a[key], b[key]), (ng, batch, key, env)
        for k in range(0, 160, 40):
            b, color = boards[k]
            t = orc.OracleTree(size, orc.hashnet2, tree_size=4096, batch_size=batch)
            t.set_noise_key(3, k, b.moves)
            mv = t.genmove_puct(b, color, visits, False)
            assert ref[0]["move"][k] in (mv, -1) and ref[2][k] == t.num_nodes
            assert np.array_equal(ref[0]["visits"][k, :t.node(0)["num_children"]], t.node(0)["children_visits"])
