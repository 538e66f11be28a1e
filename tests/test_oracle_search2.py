"""T3 at the full BASELINE budgets: oracle trees against the digest goldens recorded from the reference MCTSTree
(tests/golden/make_golden.py gen_search2).  Covers what search_<N>.npz does not: a NON-dyadic evaluator (fp32 sums round,
so the reference's fp32 queue-order value accumulation is observable), 13x13, and 19x19 at PUCT-400 batch 1 / 8 with
super-ko (configs[3]), SH-400, PUCT-1600 with 256-leaf batches and Dirichlet tentative priors (configs[4])."""
import os

import numpy as np
import pytest

from oracle import oracle as orc
from golden_util import DigestGolden, compare_digest_case

EVALS = {0: orc.hashnet, 1: orc.hashnet2}


def run_oracle_case(dg, i):
    meta, ref = dg.case(i)
    n = dg.size
    b = orc.OracleBoard(n, 7.0, True, dg.zobrist)
    color = orc.BLACK
    for p in dg.movelist(meta["pos_index"]):
        b.put_stone(int(p), color)
        color = 3 - color
    assert color == meta["color"]
    t = orc.OracleTree(n, EVALS[meta["evaluator"]], tree_size=4096, batch_size=meta["batch"], use_libm=True)
    t.set_noise_key(dg.seed, meta["pos_index"], b.moves)
    if meta["kind"] == 0:
        mv = t.genmove_sh(b, color, meta["visits"], True)
    else:
        mv = t.genmove_puct(b, color, meta["visits"], bool(meta["strict"]))
    return meta, ref, t, mv


@pytest.mark.parametrize("size", [9, 13, 19])
def test_oracle_matches_digest_goldens(golden_dir, size):
    dg = DigestGolden(os.path.join(golden_dir, f"search2_{size}.npz"))
    assert dg.ncases > 0
    seen = set()
    for i in range(dg.ncases):
        meta, ref, t, mv = run_oracle_case(dg, i)
        tag = f"size={size} kind={meta['kind']} pos={meta['pos_index']} visits={meta['visits']} batch={meta['batch']} ev={meta['evaluator']}"
        compare_digest_case(t.node, t.num_nodes, mv, meta, ref, tag)
        if meta["kind"] == 0:
            np.testing.assert_allclose(t.improved_policy(0), ref["improved"], rtol=1e-12, atol=1e-300, err_msg=tag)
        seen.add((meta["kind"], meta["visits"], meta["batch"], meta["evaluator"]))
    if size == 19:
        assert {(1, 400, 1, 1), (1, 400, 8, 1), (0, 400, 1, 1), (1, 1600, 256, 1)} <= seen


def test_nondyadic_sums_really_round(golden_dir):
    """The point of hashnet2: its value sums are NOT exact in fp32 (they differ from a float64 accumulation), so a wrong
    accumulation precision or order cannot hide."""
    dg = DigestGolden(os.path.join(golden_dir, "search2_9.npz"))
    differ = 0
    for i in range(dg.ncases):
        meta, ref = dg.case(i)
        if meta["evaluator"] != 1:
            continue
        vs, n = ref["root"]["vsum"].astype(np.float64), ref["root"]["visits"]
        # k/1000 values: a float64 sum of fp32-rounded terms would carry bits below the fp32 ulp
        differ += int(np.any(np.abs(vs * 1000.0 - np.round(vs * 1000.0)) > 1e-9) and n.sum() > 0)
    assert differ > 0


def test_native_hash_evaluators_equal_numpy_twins():
    """OracleTree runs hashnet / hashnet2 through their C twins (no Python in the search loop): same bits as numpy."""
    import ctypes as C
    rs = np.random.RandomState(0)
    for n in (9, 19):
        cls = rs.randint(0, 3, (7, n, n))
        x = np.zeros((7, 6, n, n), np.float32)
        for c in range(3):
            x[:, c] = cls == c
        x[:, 3, 2, 3] = 1.0
        x[:, 5] = np.where(rs.rand(7) < 0.5, 1.0, -1.0)[:, None, None]
        for var, fn in enumerate((orc.hashnet, orc.hashnet2)):
            for use_logit in (False, True):
                ctx = orc.HashNetCtx(n, var)
                pol = np.zeros((7, n * n + 1), np.float32)
                val = np.zeros((7, 3), np.float32)
                f = C.cast(orc.lib().tgo_hashnet_fn(), orc.EVAL_FN)
                f(C.cast(C.pointer(ctx), C.c_void_p), x.ctypes.data_as(C.POINTER(C.c_float)), 7, int(use_logit),
                  pol.ctypes.data_as(C.POINTER(C.c_float)), val.ctypes.data_as(C.POINTER(C.c_float)))
                rp, rv = fn(x, use_logit)
                assert np.array_equal(pol, rp) and np.array_equal(val, rv), (n, var, use_logit)
