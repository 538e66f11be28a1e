"""GTP analysis surface on the device search: lz-analyze / cgos-analyze text, PV lists and the
tamago-dump_tree JSON of PUCT searches are compared with what the reference wrote for the same positions
(tests/golden/analysis_9.npz: hash evaluator and counter-based noise injected into the reference)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


class _HashNet:
    def __init__(self):
        import tamago_b200 as tb
        self.evaluator = tb.EVAL_HASHNET
        self.state_dict_np = None


def test_analysis_output_matches_reference(golden_dir, capsys):
    from tamago_b200.board.go_board import GoBoard
    from tamago_b200.board.stone import Stone
    from tamago_b200.mcts.tree import MCTSTree
    from tamago_b200.mcts.time_manager import TimeManager, TimeControl
    g = np.load(os.path.join(golden_dir, "analysis_9.npz"))
    size, seed = int(g["size"]), int(g["seed"])
    off = g["movelist_off"]
    for i, (pi, visits, batch, move, color_v, num_nodes) in enumerate(g["meta"]):
        board = GoBoard(size, 7.0, True)
        board.zobrist_table = g["zobrist"]
        color = Stone.BLACK
        for p in g["movelist"][off[pi]:off[pi + 1]]:
            board.put_stone(int(p), color)
            color = Stone.get_opponent_color(color)
        assert color.value == color_v
        tree = MCTSTree(_HashNet(), tree_size=4096, batch_size=int(batch), seed=seed)
        tree._game_counter = int(pi) - 1                     # noise key (seed, game = pos_index, move = board.moves)
        tm = TimeManager(TimeControl.CONSTANT_PLAYOUT, constant_visits=int(visits))
        capsys.readouterr()
        mv = tree.search_best_move(board, color, tm, {"mode": "lz", "interval": 0})
        out = capsys.readouterr().out
        assert mv == move and tree.num_nodes == num_nodes, (i, mv, tree.num_nodes)
        root = tree.get_root()
        if batch == 1:      # with batch > 1 the reference reports before its last partial batch is evaluated (tree.py:170-174 vs 81-82)
            assert out == str(g["lz_stdout"][i])
        assert root.get_analysis(board, "lz", tree.get_pv_lists) == str(g["lz"][i])
        assert root.get_analysis(board, "cgos", tree.get_pv_lists) == str(g["cgos"][i])
        assert json.dumps(tree.get_pv_lists(root, board.coordinate)) == str(g["pv"][i])
        mine, ref = json.loads(tree.dump_to_json(board, True)), json.loads(str(g["dump"][i]))
        assert mine.keys() == ref.keys()
        for key in ref:
            if key != "tree":
                assert mine[key] == ref[key], (i, key)
        for key in ref["tree"]:
            if key != "node":
                assert mine["tree"][key] == ref["tree"][key], (i, key)
        for a, b in zip(mine["tree"]["node"], ref["tree"]["node"][:num_nodes]):
            for key in b:
                assert a[key] == b[key], (i, key)


def test_search_and_ponder_report(golden_dir, capsys):
    """MCTSTree.search (tree.py:130) and ponder (tree.py:108) write well-formed lz / cgos reports for the DualNet evaluator."""
    from tamago_b200.board.go_board import GoBoard
    from tamago_b200.board.stone import Stone
    from tamago_b200.mcts.tree import MCTSTree
    from tamago_b200.mcts.time_manager import TimeManager, TimeControl
    from tamago_b200.nn.network import DualNet
    net = DualNet(9, seed=0)
    board = GoBoard(9, 7.0, True)
    for p, c in ((60, Stone.BLACK), (62, Stone.WHITE)):
        board.put_stone(p, c)
    tree = MCTSTree(net, batch_size=4)
    tm = TimeManager(TimeControl.STRICT_PLAYOUT, constant_visits=200)
    capsys.readouterr()
    tree.search(board, Stone.BLACK, tm, {"mode": "cgos", "interval": 0})
    rep = json.loads(capsys.readouterr().out)
    assert rep["visits"] == tree.get_root().node_visits and rep["visits"] >= 199
    assert sum(m["visits"] for m in rep["moves"]) == rep["visits"] and rep["moves"][0]["order"] == 0
    assert all(m["pv"].split(" ")[0] == m["move"] for m in rep["moves"])
    # without the "ponder" flag (no stdin polling) the analysis is ONE bounded round (it used to keep doubling)
    tree.ponder(board, Stone.BLACK, {"mode": "lz", "interval": 100, "visits": 300})
    lines = capsys.readouterr().out.strip().split("\n")
    assert len(lines) == 1 and lines[0].startswith("info move ") and tree.get_root().node_visits == 300
    # with it, rounds double until stdin has input or max_visits is reached (pytest's captured stdin ends it after a round)
    tree.ponder(board, Stone.BLACK, {"mode": "lz", "interval": 100, "ponder": True, "max_visits": 512})
    lines = capsys.readouterr().out.strip().split("\n")
    assert 1 <= len(lines) <= 2 and all(l.startswith("info move ") for l in lines)
    with pytest.raises(NotImplementedError):
        tree.search_with_callback(board, Stone.BLACK, lambda path: True)
