"""The Python surfaces that stay drop-in on the B200: MCTSTree, selfplay_worker, GoBoard, DualNet."""
import os

import numpy as np
import pytest

from golden_util import SearchGolden

pytestmark = pytest.mark.gpu


class _HashNet:
    """stands in for the network object: selects the engine's hash evaluator (tree parity with the goldens)"""
    def __init__(self):
        import tamago_b200 as tb
        self.evaluator = tb.EVAL_HASHNET
        self.state_dict_np = None


def test_mctstree_dropin_matches_reference_golden(golden_dir):
    from tamago_b200.board.go_board import GoBoard
    from tamago_b200.board.stone import Stone
    from tamago_b200.mcts.tree import MCTSTree
    from tamago_b200.mcts.time_manager import TimeManager, TimeControl
    sg = SearchGolden(os.path.join(golden_dir, "search_9.npz"))
    checked = 0
    for i in range(sg.ncases):
        meta, nodes, improved = sg.case(i)
        if meta["batch"] != 1:
            continue
        board = GoBoard(9, 7.0, True)
        board.zobrist_table = sg.zobrist
        color = Stone.BLACK
        for p in sg.movelist(meta["pos_index"]):
            board.put_stone(int(p), color)
            color = Stone.get_opponent_color(color)
        tree = MCTSTree(_HashNet(), tree_size=4096, seed=sg.seed)
        tree._game_counter = meta["pos_index"] - 1          # noise key (seed, game = pos_index, move = board.moves)
        tm = TimeManager(TimeControl.CONSTANT_PLAYOUT, constant_visits=meta["visits"])
        before = board.get_move_history()
        if meta["kind"] == 0:
            mv = tree.generate_move_with_sequential_halving(board, color, tm, True)
        else:
            mv = tree.search_best_move(board, color, tm, {})
        assert board.get_move_history() == before            # the caller's board is not modified
        assert mv == meta["move"], (i, meta)
        root = tree.get_root()
        assert root.get_num_children() == nodes[0]["k"]
        assert np.array_equal(root.children_visits, nodes[0]["visits"])
        assert [root.get_child_move(j) for j in range(root.get_num_children())] == list(nodes[0]["action"])
        assert tree.num_nodes == len(nodes)
        if meta["kind"] == 0:
            np.testing.assert_allclose(root.calculate_improved_policy(), improved, rtol=1e-12)
        checked += 1
    assert checked >= 10


def test_selfplay_worker_writes_reference_sgf(golden_dir, tmp_path):
    import tamago_b200 as tb
    from tamago_b200.selfplay.worker import selfplay_worker
    g = np.load(os.path.join(golden_dir, "selfplay_9.npz"))
    ng = len(g["sgf"])
    for dedup in (False, True):
        d = tmp_path / f"dedup{int(dedup)}"
        d.mkdir()
        n = selfplay_worker(str(d), "/nonexistent/model.bin", list(range(ng)), int(g["size"]), int(g["visits"]), True,
                            pool_size=ng, dedup=dedup, seed=int(g["seed"]), network=_HashNet(), evaluator=tb.EVAL_HASHNET,
                            zobrist=g["zobrist"], never_resign_fn=lambda idx: idx % 2 == 0)
        assert n == len(g["moves"])
        for k in range(ng):
            assert (d / f"{k}.sgf").read_text(encoding="utf-8") == str(g["sgf"][k])
        # resume rule of worker.py:47-48: existing files are skipped
        assert selfplay_worker(str(d), "/nonexistent/model.bin", list(range(ng)), 9, 16, True, network=_HashNet(),
                               evaluator=tb.EVAL_HASHNET) == 0


def test_selfplay_worker_pool_smaller_than_index_list(tmp_path):
    """More games than pool slots: finished games are replaced by the next index; all files appear, all parse."""
    import tamago_b200 as tb
    from tamago_b200.selfplay.worker import selfplay_worker
    from tamago_b200.nn.network import DualNet
    net = DualNet(board_size=9, seed=3)
    n = selfplay_worker(str(tmp_path), "", list(range(1, 13)), 9, 16, True, pool_size=5, seed=1, network=net)
    files = sorted(os.listdir(tmp_path))
    assert len(files) == 12 and n > 12 * 10
    for f in files:
        text = (tmp_path / f).read_text()
        assert text.startswith("(;FF[4]GM[1]SZ[9]\nAP[TamaGo]") and text.endswith("\n)") and ";B[" in text


def test_goboard_queries_match_oracle():
    from oracle import oracle as orc
    from tamago_b200.board.go_board import GoBoard
    import tamago_b200.board.go_board as gb
    rs = np.random.RandomState(8)
    zob = orc.default_zobrist(9)
    b = GoBoard(9, 7.0, True)
    ob = orc.OracleBoard(9, 7.0, True, zob)
    gb._rule_engine(9, True).set_zobrist(zob)
    color = 1
    for i in range(60):
        cand = ob.candidates(color)
        pos = int(rs.choice(cand[:-1])) if len(cand) > 1 else 0
        b.put_stone(pos, color); ob.put_stone(pos, color); color = 3 - color
        if i % 10 == 9:
            legal = [p for p in ob.onboard_pos if ob.is_legal(p, color)]
            assert b.get_all_legal_pos(color) == legal
            assert b.get_candidates(color) == list(ob.candidates(color))
            assert b.count_score() == ob.count_score() and b.get_hash() == ob.hash
            assert b.get_board_data() == [int(ob.state()["color"][p]) for p in ob.onboard_pos]


def test_dualnet_handle_inference_matches_oracle():
    import torch
    from oracle.dualnet_ref import DualNetRef
    from tamago_b200.nn.network import DualNet
    net = DualNet(board_size=9, seed=5)
    ref = DualNetRef(net.state_dict_np, 9)
    rs = np.random.RandomState(0)
    x = np.zeros((7, 6, 9, 9), np.float32)
    cls = rs.randint(0, 3, (7, 9, 9))
    for c in range(3):
        x[:, c] = (cls == c)
    x[:, 5] = 1.0
    pol, val = net.inference(torch.from_numpy(x))
    rp, rv = ref.evaluator()(x, False)
    assert isinstance(pol, torch.Tensor) and np.abs(pol.numpy() - rp).max() <= 1e-4 and np.abs(val.numpy() - rv).max() <= 1e-4
    lg, _ = net.inference_with_policy_logits(torch.from_numpy(x))
    assert np.abs(lg.numpy() - ref.evaluator()(x, True)[0]).max() <= 1e-4


def test_c1_one_game_puct100_through_the_dropin_loop():
    """BASELINE.json configs[0] plumbing: one 9x9 game played move by move through GoBoard + MCTSTree.search_best_move
    (100-visit PUCT, batch 1) exactly like the reference's worker / GTP loop, against the oracle's game with the same
    evaluator and noise keys."""
    from oracle import oracle as orc
    from tamago_b200.board.go_board import GoBoard
    from tamago_b200.board.stone import Stone
    from tamago_b200.mcts.tree import MCTSTree
    from tamago_b200.mcts.time_manager import TimeManager, TimeControl
    zob = orc.default_zobrist(9)
    ot = orc.OracleTree(9, orc.hashnet, tree_size=4096)
    want = ot.selfplay_game(7.0, zob, seed=31, game=1, visits=100, never_resign=False, use_puct=True)
    board = GoBoard(9, 7.0, True)
    board.zobrist_table = zob
    tree = MCTSTree(_HashNet(), batch_size=1, seed=31)
    tm = TimeManager(TimeControl.CONSTANT_PLAYOUT, constant_visits=100)
    color, moves, passes = Stone.BLACK, [], 0
    for _ in range(2 * 81):
        tree._game_counter = 0                                   # noise key: game id 1 for every move of this game
        pos = tree.search_best_move(board, color, tm, {})
        if pos == -1:
            break
        board.put_stone(pos, color)
        moves.append(pos)
        passes = passes + 1 if pos == 0 else 0
        color = Stone.get_opponent_color(color)
        if passes == 2:
            break
    assert moves == [int(p) for p in want["pos"]]
    assert len(moves) > 20


def test_async_step_and_record_ring(golden_dir):
    """tg_genmove_async / tg_collect / tg_fetch_records / tg_format_records: the golden self-play games again, driven
    through the asynchronous API with the records taken from the device ring instead of per-step root arrays."""
    import tamago_b200 as tb
    g = np.load(os.path.join(golden_dir, "selfplay_9.npz"))
    size, seed, visits = int(g["size"]), int(g["seed"]), int(g["visits"])
    ng = len(g["sgf"])
    e = tb.Engine(board_size=size, games=ng, max_visits=visits, superko=True, evaluator=tb.EVAL_HASHNET, seed=seed, record_ring=True)
    e.set_zobrist(g["zobrist"])
    e.reset(game_ids=np.arange(ng), never_resign=g["never_resign"])
    texts = {}
    e.genmove_async(mode=tb.MODE_SH, visits=visits, play=True)
    with pytest.raises(Exception):
        e.genmove_async(mode=tb.MODE_SH, visits=visits, play=True)        # one step in flight at a time
    for _ in range(2 * size * size + 2):
        r = e.collect()
        fin = [k for k in range(ng) if r["finished"][k] and k not in texts]
        if fin:
            e.fetch_records(fin)
        if len(texts) + len(fin) < ng:
            e.genmove_async(mode=tb.MODE_SH, visits=visits, play=True)    # next step runs while the records are formatted
        if fin:
            for k, t in zip(fin, e.format_records()):
                texts[k] = t
            rec = e.fetched_record(0)
            want = g["moves"][g["moves_off"][fin[0]]:g["moves_off"][fin[0] + 1]]
            assert np.array_equal(rec["move"], want) and rec["action"].shape == (len(want), e.stride)
        if len(texts) == ng:
            break
    assert len(texts) == ng
    for k in range(ng):
        assert texts[k] == str(g["sgf"][k]), f"game {k}"
    e.close()


def test_forward_device_back_to_back_and_stream_ordering(golden_dir):
    """two tg_forward_device calls without a sync in between must not race on the slot count, and the engine's
    stream must be ordered with the torch stream that fills the aliased planes (tg_stream_wait / tg_stream_signal)."""
    import torch
    import tamago_b200 as tb
    from tamago_b200.nn.utility import random_init_state_dict
    g = np.load(os.path.join(golden_dir, "dualnet_9.npz"))
    x = g["planes"]
    e = tb.Engine(board_size=9, games=64, max_visits=32, evaluator=tb.EVAL_DUALNET_TC)
    e.load_state_dict(random_init_state_dict(9, 1))
    ref_p, ref_v = e.forward(x, use_logit=True)
    planes, policy, value = e.eval_tensors()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for rep in range(20):
            n1, n2 = len(x), 3
            policy.zero_(); value.zero_()
            planes[:n1].copy_(torch.from_numpy(x).cuda(non_blocking=True), non_blocking=True)
            e.forward_device(n1, use_logit=True, torch_stream=st)          # no host synchronisation anywhere
            p1 = policy[:n1].clone()
            e.forward_device(n2, use_logit=True, torch_stream=st)          # a smaller count right behind the first call
            p2 = policy[:n1].clone()
        st.synchronize()
    assert np.array_equal(p1.cpu().numpy(), ref_p) and np.array_equal(p2.cpu().numpy(), ref_p)
    e.close()
