"""GTP analysis surface -- host formatting checked against the reference's own strings without a GPU:
the tree of every golden case is rebuilt from the reference's tamago-dump_tree JSON and fed to the host mirror."""
import json
import os

import numpy as np

from tamago_b200.board.coordinate import Coordinate
from tamago_b200.mcts.node import MCTSNodeView
from tamago_b200.mcts.tree import MCTSTree


class _Board:
    def __init__(self, size):
        self.coordinate = Coordinate(size)


class _DumpTree(MCTSTree):
    """MCTSTree whose nodes come from a reference dump instead of the device"""
    def __init__(self, tree_dict):
        super().__init__(network=None, batch_size=tree_dict["batch_size"], cgos_mode=tree_dict["cgos_mode"])
        self._nodes = tree_dict["node"]
        self.num_nodes = tree_dict["num_nodes"]
        self.to_move = 1 if tree_dict["to_move"] == "black" else 2

    def _node_at(self, index):
        d = self._nodes[index]
        k = d["num_children"]
        f = {key: (np.asarray(v[:k]) if isinstance(v, list) and key != "noise" else v) for key, v in d.items()}
        f["noise"] = np.asarray(d["noise"])
        f["children_value_sum"] = f["children_value_sum"].astype(np.float32)      # the device keeps fp32 sums
        return MCTSNodeView(f, max_actions=len(d["action"]))


def test_analysis_strings_and_pv_from_reference_dumps(golden_dir):
    g = np.load(os.path.join(golden_dir, "analysis_9.npz"))
    board = _Board(int(g["size"]))
    for i in range(len(g["meta"])):
        dump = json.loads(str(g["dump"][i]))
        tree = _DumpTree(dump["tree"])
        root = tree._node_at(0)
        assert json.dumps(tree.get_pv_lists(root, board.coordinate)) == str(g["pv"][i])
        assert root.get_analysis(board, "lz", tree.get_pv_lists) == str(g["lz"][i])
        assert root.get_analysis(board, "cgos", tree.get_pv_lists) == str(g["cgos"][i])
        # to_dict round trip: the host mirror reproduces the reference's node dictionaries
        mine = tree.to_dict()
        assert mine["num_nodes"] == dump["tree"]["num_nodes"] and mine["to_move"] == dump["tree"]["to_move"]
        for a, b in zip(mine["node"], dump["tree"]["node"][:mine["num_nodes"]]):
            assert a.keys() == b.keys()
            for key in a:
                assert a[key] == b[key], (i, key)
