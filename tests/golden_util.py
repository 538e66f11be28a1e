"""Helpers to read the ragged golden search trees written by tests/golden/make_golden.py."""
import numpy as np


class SearchGolden:
    def __init__(self, path):
        g = dict(np.load(path))
        self.g = g
        self.size = int(g["size"])
        self.seed = int(g["seed"])
        self.zobrist = g["zobrist"]
        self.ncases = len(g["case_kind"])

    def movelist(self, pos_index):
        o = self.g["movelist_off"]
        return self.g["movelist"][o[pos_index]:o[pos_index + 1]]

    def case(self, i):
        g = self.g
        meta = {k[5:]: int(g[k][i]) for k in g if k.startswith("case_")}
        nodes = []
        for ni in range(g["node_off"][i], g["node_off"][i + 1]):
            a, b = g["child_off"][ni], g["child_off"][ni + 1]
            nd = dict(k=int(g["node_scal"][ni][0]), node_visits=int(g["node_scal"][ni][1]),
                      virtual_loss=int(g["node_scal"][ni][2]), node_value_sum=float(g["node_fsum"][ni][0]),
                      raw_value=float(g["node_fsum"][ni][1]))
            for k in ("action", "cidx", "value", "visits", "policy", "vl", "vsum"):
                nd[k] = g["ch_" + k][a:b]
            nodes.append(nd)
        io = g["improved_off"]
        improved = g["improved"][io[i]:io[i + 1]]
        return meta, nodes, improved
