"""T3: oracle search trees against trees recorded from the reference MCTSTree.

Both sides use the hash evaluator (oracle.hashnet, exactly representable fp32
outputs) and the counter-based Dirichlet/Gumbel noise, so the whole tree must
agree: integers and fp32-accumulated sums bit for bit, float64 priors exactly,
softmax-derived quantities to 1e-12 (np.exp vs libm exp may differ by an ulp).
"""
import os

import numpy as np
import pytest

from oracle import oracle as orc
from golden_util import SearchGolden


def _run_case(sg, i, use_libm=True):
    meta, nodes, improved = sg.case(i)
    n = sg.size
    b = orc.OracleBoard(n, 7.0, True, sg.zobrist)
    color = orc.BLACK
    for p in sg.movelist(meta["pos_index"]):
        b.put_stone(int(p), color)
        color = 3 - color
    assert color == meta["color"]
    t = orc.OracleTree(n, orc.hashnet, tree_size=4096, batch_size=meta["batch"], use_libm=use_libm)
    t.set_noise_key(sg.seed, meta["pos_index"], b.moves)
    if meta["kind"] == 0:
        mv = t.genmove_sh(b, color, meta["visits"], True)
    else:
        mv = t.genmove_puct(b, color, meta["visits"], False)
    return meta, nodes, improved, t, mv


def _compare(meta, nodes, improved, t, mv):
    tag = f"case kind={meta['kind']} pos={meta['pos_index']} visits={meta['visits']} batch={meta['batch']}"
    assert mv == meta["move"], tag
    assert t.num_nodes == len(nodes), tag
    for ni, ref in enumerate(nodes):
        nd = t.node(ni)
        assert nd["num_children"] == ref["k"], f"{tag} node {ni}"
        assert nd["node_visits"] == ref["node_visits"] and nd["virtual_loss"] == ref["virtual_loss"], f"{tag} node {ni}"
        assert np.array_equal(nd["action"], ref["action"]), f"{tag} node {ni} actions"
        assert np.array_equal(nd["children_index"], ref["cidx"]), f"{tag} node {ni} child index"
        assert np.array_equal(nd["children_visits"], ref["visits"]), f"{tag} node {ni} visits"
        assert np.array_equal(nd["children_virtual_loss"], ref["vl"]), f"{tag} node {ni} virtual loss"
        assert np.array_equal(nd["children_value_sum"].astype(np.float64), ref["vsum"]), f"{tag} node {ni} value sums"
        assert np.array_equal(nd["children_value"].astype(np.float64), ref["value"]), f"{tag} node {ni} leaf values"
        assert np.array_equal(nd["children_policy"], ref["policy"]), f"{tag} node {ni} policy"
        assert nd["node_value_sum"] == ref["node_value_sum"], f"{tag} node {ni} node_value_sum"
        assert nd["raw_value"] == ref["raw_value"], f"{tag} node {ni} raw_value"
    if meta["kind"] == 0:
        ip = t.improved_policy(0)
        np.testing.assert_allclose(ip, improved, rtol=1e-12, atol=1e-300, err_msg=tag)


@pytest.mark.parametrize("size", [9, 19])
def test_search_trees_match_reference(golden_dir, size):
    path = os.path.join(golden_dir, f"search_{size}.npz")
    if not os.path.isfile(path):
        pytest.skip("golden file not generated")
    sg = SearchGolden(path)
    assert sg.ncases > 0
    for i in range(sg.ncases):
        _compare(*_run_case(sg, i))


def test_det_exp_mode_keeps_decisions(golden_dir):
    """The deterministic exp (shared with the CUDA side) changes softmax values by <= 2 ulp, never the tree."""
    sg = SearchGolden(os.path.join(golden_dir, "search_9.npz"))
    for i in range(sg.ncases):
        _compare(*_run_case(sg, i, use_libm=False))


def test_sh_schedule_table():
    # SURVEY.md B.3 (probe of mcts/sequential_halving.py)
    assert orc.sh_schedule(16, 400) == [(16, 6), (8, 12), (4, 25), (2, 54)]
    assert orc.sh_schedule(16, 50) == [(16, 1), (8, 1), (4, 3), (2, 7)]
    assert orc.sh_schedule(16, 1600) == [(16, 25), (8, 50), (4, 100), (2, 200)]
    assert orc.sh_schedule(8, 16) == [(8, 1), (4, 1), (2, 2)]
    assert orc.sh_schedule(5, 16) == [(5, 1), (2, 5), (1, 1)]
    assert orc.sh_schedule(5, 50) == [(5, 3), (2, 17), (1, 1)]
    assert orc.sh_schedule(3, 400) == [(3, 66), (2, 101)]
    assert orc.sh_schedule(2, 100) == [(2, 50)]
    assert orc.sh_schedule(1, 7) == [(1, 7)]
    for m, v in orc.sh_schedule(16, 100):
        assert m * v > 0
    assert sum(m * v for m, v in orc.sh_schedule(16, 100)) == 100


def test_np_sum_matches_numpy():
    import ctypes as C
    rs = np.random.RandomState(3)
    for n in list(range(1, 140)) + [176, 186, 200, 255, 256, 300, 361, 362]:
        for _ in range(5):
            a = rs.standard_exponential(n) * rs.choice([1e-3, 1.0, 1e3], n)
            got = orc.lib().tgo_np_sum(a.ctypes.data_as(C.POINTER(C.c_double)), n)
            assert got == np.sum(a), n


def test_det_math_accuracy():
    rs = np.random.RandomState(4)
    L = orc.lib()
    for x in np.concatenate([rs.uniform(1e-300, 1, 200), rs.uniform(1, 1e6, 100), [1.0, 0.5, 2.0, 1e-16]]):
        assert abs(L.tgo_det_log(float(x)) - np.log(x)) <= 4e-16 * max(1.0, abs(np.log(x)))
    for x in np.concatenate([rs.uniform(-700, 0, 300), rs.uniform(0, 50, 50), [0.0, -1e-9]]):
        assert abs(L.tgo_det_exp(float(x)) - np.exp(x)) <= 4e-16 * np.exp(x)
    assert L.tgo_det_exp(-800.0) == 0.0
