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


# ----------------------------------------------------------------------------------------------------------
# Digest goldens (search2_<N>.npz): big trees (19x19 PUCT-400 / PUCT-1600 batch 256) are pinned by a 64-bit digest per
# node over the exact bytes of its child arrays plus the full root node, instead of megabytes of child rows.
# ----------------------------------------------------------------------------------------------------------
def node_digest(action, cidx, visits, vl, vsum, value, policy):
    """blake2b-64 over the child arrays of one node: int16 actions, int32 child indices / visits / virtual losses,
    float32 value sums and leaf values (the reference holds fp32-representable numbers in float64 slots: asserted when
    the golden is made), float64 priors."""
    import hashlib
    h = hashlib.blake2b(digest_size=8)
    for a, dt in ((action, np.int16), (cidx, np.int32), (visits, np.int32), (vl, np.int32), (vsum, np.float32),
                  (value, np.float32), (policy, np.float64)):
        h.update(np.ascontiguousarray(np.asarray(a).astype(dt)).tobytes())
    return np.frombuffer(h.digest(), np.uint64)[0]


class DigestGolden:
    """search2_<N>.npz: per case meta, per node scalars + digest, the root node in full, improved policy (SH)."""
    def __init__(self, path):
        g = dict(np.load(path))
        self.g = g
        self.size, self.seed, self.zobrist = int(g["size"]), int(g["seed"]), g["zobrist"]
        self.ncases = len(g["case_kind"])

    def movelist(self, pos_index):
        o = self.g["movelist_off"]
        return self.g["movelist"][o[pos_index]:o[pos_index + 1]]

    def case(self, i):
        g = self.g
        meta = {k[5:]: int(g[k][i]) for k in g if k.startswith("case_")}
        a, b = g["node_off"][i], g["node_off"][i + 1]
        ra, rb = g["root_off"][i], g["root_off"][i + 1]
        root = {k: g["root_" + k][ra:rb] for k in ("action", "cidx", "value", "visits", "policy", "vl", "vsum")}
        io = g["improved_off"]
        return meta, dict(scal=g["node_scal"][a:b], fsum=g["node_fsum"][a:b], digest=g["node_digest"][a:b], root=root,
                          improved=g["improved"][io[i]:io[i + 1]])


def compare_digest_case(read_node, num_nodes, move, meta, ref, tag):
    """read_node(i) -> dict with the engine's / oracle's node arrays (keys as in Engine.node / OracleTree.node)."""
    assert move == meta["move"], f"{tag}: move {move} != {meta['move']}"
    assert num_nodes == len(ref["scal"]), f"{tag}: {num_nodes} nodes != {len(ref['scal'])}"
    for ni in range(num_nodes):
        nd = read_node(ni)
        t = f"{tag} node {ni}"
        assert [nd["num_children"], nd["node_visits"], nd["virtual_loss"]] == list(ref["scal"][ni]), t + " scalars"
        assert np.float32(nd["node_value_sum"]) == np.float32(ref["fsum"][ni][0]), t + " node_value_sum"
        assert np.float32(nd["raw_value"]) == np.float32(ref["fsum"][ni][1]), t + " raw_value"
        if ni == 0:
            r = ref["root"]
            assert np.array_equal(nd["action"], r["action"]), t + " actions"
            assert np.array_equal(nd["children_index"], r["cidx"]), t + " child index"
            assert np.array_equal(nd["children_visits"], r["visits"]), t + " visits"
            assert np.array_equal(nd["children_virtual_loss"], r["vl"]), t + " virtual loss"
            assert np.array_equal(nd["children_value_sum"].astype(np.float32), r["vsum"].astype(np.float32)), t + " value sums"
            assert np.array_equal(nd["children_value"].astype(np.float32), r["value"].astype(np.float32)), t + " leaf values"
            assert np.array_equal(nd["children_policy"], r["policy"]), t + " policy"
        d = node_digest(nd["action"], nd["children_index"], nd["children_visits"], nd["children_virtual_loss"],
                        nd["children_value_sum"], nd["children_value"], nd["children_policy"])
        assert d == ref["digest"][ni], t + " digest of the child arrays"
