#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference (kobanium/TamaGo).

Runs only in the build container (needs /root/reference).  The reference is
pure Python, has no tests and cannot travel to the GPU box, so everything the
parity tests need from it is recorded here and committed as small .npz files:

  board_<N>.npz    per-ply board state of random legal games (T0 parity)
  search_<N>.npz   MCTS trees (sequential halving + PUCT) with a hash
                   evaluator and counter-based noise injected on both sides (T3)
  selfplay_9.npz   full self-play games incl. the SGF text (T4)
  dualnet_<N>.npz  DualNet outputs for seeded numpy weights on real planes (T2)
  analysis_9.npz   lz-analyze / cgos-analyze strings, PV lists and tree dumps of PUCT searches
  eye_table.npz    the 65 536-entry eye LUT of board/pattern.py
  search2_<N>.npz  digest goldens at the full BASELINE budgets (19x19 PUCT-400 batch 1/8 + super-ko, SH-400, PUCT-1600
                   batch 256), 13x13, and a NON-dyadic hash evaluator (fp32 queue-order value accumulation observable)
  model_ref_9.bin  a model.bin written by the reference's save_model and read back by its load_network (+ .npz outputs)

Usage:  python tests/golden/make_golden.py --size 9   (and --size 19)
        19x19 needs BOARD_SIZE patched in a private copy of the reference; the
        script makes that copy under /tmp and never writes to /root/reference.
"""
import argparse
import os
import random
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.dont_write_bytecode = True


def prepare_reference(size):
    src = "/root/reference"
    if size == 9:
        return src
    dst = f"/tmp/tamago_ref{size}"
    if not os.path.isdir(dst):
        shutil.copytree(src, dst)
        os.system(f"chmod -R u+w {dst}")
        p = os.path.join(dst, "board", "constant.py")
        s = open(p).read().replace("BOARD_SIZE = 9", f"BOARD_SIZE = {size}")
        open(p, "w").write(s)
    return dst


def board_state(b, Stone):
    cells = len(b.board)
    color = np.array([s.value for s in b.board], np.uint8)
    libs = np.zeros(cells, np.int16)
    size = np.zeros(cells, np.int16)
    for p in range(cells):
        if b.board[p] in (Stone.BLACK, Stone.WHITE):
            st = b.strings.string[b.strings.get_id(p)]
            libs[p] = st.get_num_liberties()
            size[p] = st.get_size()
    return color, libs, size


def expand_candidates(b, color):
    """mcts/tree.py:260-264 verbatim call sequence."""
    cand = b.get_all_legal_pos(color)
    cand = [c for c in cand if b.check_self_atari_stone(c, color) < 7 and not b.is_complete_eye(c, color)]
    return cand


def gen_board(size, n_games, seed, out):
    from board.go_board import GoBoard
    from board.stone import Stone
    from board.zobrist_hash import hash_bit_mask
    from nn.feature import generate_input_planes
    rng = random.Random(seed)
    rec = {k: [] for k in ("game", "pos", "mover", "color", "libs", "size", "scal", "hash", "legal", "cand",
                           "satari", "eye", "score", "planes_sum")}
    planes_samples, planes_keys = [], []
    for g in range(n_games):
        b = GoBoard(board_size=size, komi=7.0, check_superko=True)
        color = Stone.BLACK
        passes = 0
        max_plies = min(2 * size * size, 3 * size * size - 4)
        for ply in range(max_plies):
            legal = b.get_all_legal_pos(color)
            cand = expand_candidates(b, color)
            r = rng.random()
            # mostly candidate moves, sometimes any legal move (eye fills, self atari), rarely pass
            if r < 0.03 or not legal:
                pos = 0
            elif r < 0.25 or not cand:
                pos = rng.choice(legal)
            else:
                pos = rng.choice(cand)
            b.put_stone(pos, color)
            passes = passes + 1 if pos == 0 else 0
            color = Stone.get_opponent_color(color)
            c, l, s = board_state(b, Stone)
            rec["game"].append(g); rec["pos"].append(pos); rec["mover"].append(3 - color.value)
            rec["color"].append(c); rec["libs"].append(l); rec["size"].append(s)
            rec["scal"].append([b.moves, b.ko_pos, b.ko_move, b.prisoner[0], b.prisoner[1]])
            rec["hash"].append(int(b.positional_hash[0]))
            lg = np.zeros((2, size * size), np.uint8); cd = np.zeros((2, size * size), np.uint8)
            sa = np.zeros((2, size * size), np.int16); ey = np.zeros((2, size * size), np.uint8)
            for ci, col in enumerate((Stone.BLACK, Stone.WHITE)):
                cset = set(expand_candidates(b, col))
                for i, p in enumerate(b.onboard_pos):
                    if b.is_legal(p, col):
                        lg[ci, i] = 1
                        sa[ci, i] = b.check_self_atari_stone(p, col)
                        ey[ci, i] = int(b.is_complete_eye(p, col))
                    cd[ci, i] = int(p in cset)
            rec["legal"].append(lg); rec["cand"].append(cd); rec["satari"].append(sa); rec["eye"].append(ey)
            rec["score"].append(b.count_score())
            pl = generate_input_planes(b, color, 0)
            w = np.arange(1, pl.size + 1, dtype=np.float64)
            rec["planes_sum"].append(float((pl.reshape(-1).astype(np.float64) * w).sum()))
            if rng.random() < 0.02:
                planes_samples.append(pl); planes_keys.append(len(rec["pos"]) - 1)
            if passes >= 2 and ply > 20:
                break
    np.savez_compressed(
        out, size=size, zobrist=np.asarray(hash_bit_mask, np.uint64),
        game=np.array(rec["game"], np.int32), pos=np.array(rec["pos"], np.int16),
        mover=np.array(rec["mover"], np.uint8), color=np.array(rec["color"], np.uint8),
        libs=np.array(rec["libs"], np.int16), size_pt=np.array(rec["size"], np.int16),
        scal=np.array(rec["scal"], np.int32), hash=np.array(rec["hash"], np.uint64),
        legal=np.packbits(np.array(rec["legal"], np.uint8), axis=-1),
        cand=np.packbits(np.array(rec["cand"], np.uint8), axis=-1),
        satari=np.array(rec["satari"], np.int16), eye=np.packbits(np.array(rec["eye"], np.uint8), axis=-1),
        score=np.array(rec["score"], np.int32), planes_sum=np.array(rec["planes_sum"], np.float64),
        planes=np.array(planes_samples, np.float32), planes_ply=np.array(planes_keys, np.int32))
    print(f"board golden: {len(rec['pos'])} plies, {n_games} games -> {out}")


class HashNet:
    """Stands in for DualNet on both sides (oracle.hashnet: dyadic fp32 outputs; variant 1 = oracle.hashnet2: non-dyadic)."""
    def __init__(self, variant=0):
        import torch
        from oracle import oracle as orc
        self.torch = torch
        self.calls = 0
        self.fn = orc.hashnet2 if variant else orc.hashnet

    def inference(self, x):
        pol, val = self.fn(x.numpy(), False)
        self.calls += x.shape[0]
        return self.torch.from_numpy(pol), self.torch.from_numpy(val)

    def inference_with_policy_logits(self, x):
        pol, val = self.fn(x.numpy(), True)
        self.calls += x.shape[0]
        return self.torch.from_numpy(pol), self.torch.from_numpy(val)


class NoisePatch:
    """Replace the reference's np.random draws by the oracle's counter-based noise."""
    def __init__(self, seed):
        import mcts.tree as mtree
        import mcts.node as mnode
        from oracle import oracle as orc
        self.seed, self.game, self.move = seed, 0, 0
        self.tree = None
        patch = self

        orig_expand = mtree.MCTSTree.expand_node

        def expand_node(tree_self, board, color):
            patch.node_index = tree_self.num_nodes
            return orig_expand(tree_self, board, color)

        def tentative(candidates):
            import ctypes as C
            k = len(candidates)
            out = np.zeros(k, np.float64)
            orc.lib().tgo_dirichlet(patch.seed, patch.game, patch.move, patch.node_index, k,
                                    out.ctypes.data_as(C.POINTER(C.c_double)))
            return dict(zip(candidates, out))

        def gumbel(node_self):
            import ctypes as C
            out = np.zeros(node_self.noise.size, np.float64)
            orc.lib().tgo_gumbel(patch.seed, patch.game, patch.move, out.size,
                                 out.ctypes.data_as(C.POINTER(C.c_double)))
            node_self.noise = out

        mtree.MCTSTree.expand_node = expand_node
        mtree.get_tentative_policy = tentative
        mnode.MCTSNode.set_gumbel_noise = gumbel

    def key(self, game, move):
        self.game, self.move = game, move


def dump_tree(tree):
    nodes = []
    for i in range(tree.num_nodes):
        nd = tree.node[i]
        k = nd.num_children
        nodes.append(dict(
            k=k, node_visits=int(nd.node_visits), virtual_loss=int(nd.virtual_loss),
            node_value_sum=float(nd.node_value_sum), raw_value=float(nd.raw_value),
            action=np.array(nd.action[:k], np.int16), cidx=np.array(nd.children_index[:k], np.int32),
            value=np.array(nd.children_value[:k], np.float64), visits=np.array(nd.children_visits[:k], np.int32),
            policy=np.array(nd.children_policy[:k], np.float64), vl=np.array(nd.children_virtual_loss[:k], np.int32),
            vsum=np.array(nd.children_value_sum[:k], np.float64)))
    return nodes


def pack_trees(cases):
    """Flatten a list of (meta, nodes) into ragged arrays for npz."""
    out = {}
    meta_keys = list(cases[0][0].keys())
    for k in meta_keys:
        out["case_" + k] = np.array([c[0][k] for c in cases])
    node_off = [0]
    child_off = [0]
    scal, arrs = [], {k: [] for k in ("action", "cidx", "value", "visits", "policy", "vl", "vsum")}
    for _, nodes in cases:
        for nd in nodes:
            scal.append([nd["k"], nd["node_visits"], nd["virtual_loss"]])
            for k in arrs:
                arrs[k].append(nd[k])
            child_off.append(child_off[-1] + nd["k"])
        node_off.append(node_off[-1] + len(nodes))
    out["node_off"] = np.array(node_off, np.int64)
    out["child_off"] = np.array(child_off, np.int64)
    out["node_scal"] = np.array(scal, np.int32)
    out["node_fsum"] = np.array([[nd["node_value_sum"], nd["raw_value"]] for _, ns in cases for nd in ns], np.float64)
    for k, v in arrs.items():
        out["ch_" + k] = np.concatenate(v) if v else np.zeros(0)
    return out


def positions_for_search(size, seed, count):
    """A few positions (move lists) reached by random candidate play."""
    from board.go_board import GoBoard
    from board.stone import Stone
    rng = random.Random(seed)
    plist = [[]]
    for i in range(count - 1):
        b = GoBoard(board_size=size, komi=7.0, check_superko=True)
        color = Stone.BLACK
        moves = []
        target = rng.randint(4, size * size + size)
        for _ in range(target):
            cand = expand_candidates(b, color)
            pos = rng.choice(cand) if cand and rng.random() > 0.02 else 0
            b.put_stone(pos, color)
            moves.append(pos)
            color = Stone.get_opponent_color(color)
        plist.append(moves)
    return plist


def gen_search(size, seed, out, sh_visits, puct_visits):
    from board.go_board import GoBoard
    from board.stone import Stone
    from mcts.tree import MCTSTree
    from mcts.time_manager import TimeManager, TimeControl
    net = HashNet()
    patch = NoisePatch(seed=seed)
    cases = []
    movelists = positions_for_search(size, seed + 1, 6 if size == 9 else 3)
    ml_flat, ml_off = [], [0]
    for ml in movelists:
        ml_flat += ml; ml_off.append(len(ml_flat))
    for pi, ml in enumerate(movelists):
        b = GoBoard(board_size=size, komi=7.0, check_superko=True)
        color = Stone.BLACK
        for p in ml:
            b.put_stone(p, color); color = Stone.get_opponent_color(color)
        for visits in sh_visits:
            tree = MCTSTree(net, tree_size=4096)
            patch.key(pi, b.moves)
            tm = TimeManager(TimeControl.CONSTANT_PLAYOUT, constant_visits=visits)
            mv = tree.generate_move_with_sequential_halving(b, color, tm, True)
            ip = tree.get_root().calculate_improved_policy()
            cases.append((dict(kind=0, pos_index=pi, visits=visits, batch=1, move=mv, color=color.value), dump_tree(tree)))
            cases[-1][1][0]["improved"] = np.asarray(ip, np.float64)
        for visits, batch in puct_visits:
            tree = MCTSTree(net, tree_size=4096, batch_size=batch)
            patch.key(pi, b.moves)
            tm = TimeManager(TimeControl.CONSTANT_PLAYOUT, constant_visits=visits)
            mv = tree.search_best_move(b, color, tm, {})
            cases.append((dict(kind=1, pos_index=pi, visits=visits, batch=batch, move=mv, color=color.value), dump_tree(tree)))
    packed = pack_trees(cases)
    improved = [c[1][0].get("improved", np.zeros(0)) for c in cases]
    packed["improved"] = np.concatenate(improved)
    packed["improved_off"] = np.cumsum([0] + [len(x) for x in improved])
    from board.zobrist_hash import hash_bit_mask
    np.savez_compressed(out, size=size, seed=seed, zobrist=np.asarray(hash_bit_mask, np.uint64),
                        movelist=np.array(ml_flat, np.int16), movelist_off=np.array(ml_off, np.int64), **packed)
    print(f"search golden: {len(cases)} cases -> {out}")


def gen_search2(size, seed, out, n_positions, sh_cases, puct_cases, fixed_positions=None):
    """Digest goldens (tests/golden_util.py DigestGolden): trees of searches at full BASELINE budgets.
    sh_cases: [(visits, evaluator)], puct_cases: [(visits, batch, evaluator, strict)]; evaluator 0 = hashnet, 1 = hashnet2."""
    import time
    from board.go_board import GoBoard
    from board.stone import Stone
    from mcts.tree import MCTSTree
    from mcts.time_manager import TimeManager, TimeControl
    from golden_util import node_digest
    patch = NoisePatch(seed=seed)
    movelists = fixed_positions if fixed_positions is not None else positions_for_search(size, seed + 1, n_positions)
    ml_flat, ml_off = [], [0]
    for ml in movelists:
        ml_flat += ml; ml_off.append(len(ml_flat))
    metas, scal, fsum, digest, roots, improved = [], [], [], [], [], []
    node_off = [0]

    def f32_exact(a, what):
        a = np.asarray(a, np.float64)
        assert np.array_equal(a.astype(np.float32).astype(np.float64), a), f"{what} is not fp32-representable in the reference"
        return a.astype(np.float32)

    def record(meta, tree, ip):
        nodes = dump_tree(tree)
        for nd in nodes:
            scal.append([nd["k"], nd["node_visits"], nd["virtual_loss"]])
            fsum.append([f32_exact(nd["node_value_sum"], "node_value_sum"), f32_exact(nd["raw_value"], "raw_value")])
            digest.append(node_digest(nd["action"], nd["cidx"], nd["visits"], nd["vl"], f32_exact(nd["vsum"], "children_value_sum"),
                                      f32_exact(nd["value"], "children_value"), nd["policy"]))
        node_off.append(node_off[-1] + len(nodes))
        roots.append(nodes[0]); improved.append(np.asarray(ip, np.float64)); metas.append(meta)

    for pi, ml in enumerate(movelists):
        b = GoBoard(board_size=size, komi=7.0, check_superko=True)
        color = Stone.BLACK
        for p in ml:
            b.put_stone(p, color); color = Stone.get_opponent_color(color)
        for visits, ev in sh_cases:
            t0 = time.time()
            tree = MCTSTree(HashNet(ev), tree_size=4096)
            patch.key(pi, b.moves)
            tm = TimeManager(TimeControl.CONSTANT_PLAYOUT, constant_visits=visits)
            mv = tree.generate_move_with_sequential_halving(b, color, tm, True)
            record(dict(kind=0, pos_index=pi, visits=visits, batch=1, move=mv, color=color.value, evaluator=ev, strict=0),
                   tree, tree.get_root().calculate_improved_policy())
            print(f"  pos {pi} SH {visits} ev {ev}: {tree.num_nodes} nodes, {time.time() - t0:.1f} s")
        for visits, batch, ev, strict in puct_cases:
            t0 = time.time()
            tree = MCTSTree(HashNet(ev), tree_size=4096, batch_size=batch)
            patch.key(pi, b.moves)
            tm = TimeManager(TimeControl.STRICT_PLAYOUT if strict else TimeControl.CONSTANT_PLAYOUT, constant_visits=visits)
            mv = tree.search_best_move(b, color, tm, {})
            record(dict(kind=1, pos_index=pi, visits=visits, batch=batch, move=mv, color=color.value, evaluator=ev, strict=int(strict)),
                   tree, np.zeros(0))
            print(f"  pos {pi} PUCT {visits} batch {batch} ev {ev} strict {strict}: {tree.num_nodes} nodes, root visits "
                  f"{tree.get_root().node_visits}, {time.time() - t0:.1f} s")
    out_d = {"case_" + k: np.array([m[k] for m in metas]) for k in metas[0]}
    root_off = np.cumsum([0] + [r["k"] for r in roots])
    for k in ("action", "cidx", "value", "visits", "policy", "vl", "vsum"):
        out_d["root_" + k] = np.concatenate([r[k] for r in roots])
    from board.zobrist_hash import hash_bit_mask
    np.savez_compressed(out, size=size, seed=seed, zobrist=np.asarray(hash_bit_mask, np.uint64),
                        movelist=np.array(ml_flat, np.int16), movelist_off=np.array(ml_off, np.int64),
                        node_off=np.array(node_off, np.int64), node_scal=np.array(scal, np.int32),
                        node_fsum=np.array(fsum, np.float32), node_digest=np.array(digest, np.uint64),
                        root_off=root_off.astype(np.int64), improved=np.concatenate(improved),
                        improved_off=np.cumsum([0] + [len(x) for x in improved]), **out_d)
    print(f"search2 golden: {len(metas)} cases, {node_off[-1]} nodes -> {out}")


def gen_modelbin(size, seed, board_npz, out_bin, out_npz):
    """model.bin written by the reference's own save_model (nn/utility.py:80-87) from a seeded DualNet whose BatchNorm
    statistics are non-trivial, re-read through the reference's load_network (nn/utility.py:139-159), and the logits /
    softmax outputs of that loaded network on real positions."""
    import torch
    from nn.network.dual_net import DualNet
    from nn.utility import save_model, load_network
    torch.manual_seed(seed)
    net = DualNet(torch.device("cpu"), board_size=size)
    rs = np.random.RandomState(seed)
    with torch.no_grad():
        for name, buf in net.named_buffers():
            if name.endswith("running_mean"):
                buf.copy_(torch.from_numpy((rs.standard_normal(buf.shape) * 0.1).astype(np.float32)))
            elif name.endswith("running_var"):
                buf.copy_(torch.from_numpy(rs.uniform(0.5, 1.5, buf.shape).astype(np.float32)))
            elif name.endswith("num_batches_tracked"):
                buf.fill_(17)
        for name, p in net.named_parameters():
            if ".bn" in name or name.startswith("bn_layer"):
                p.copy_(torch.from_numpy((rs.uniform(0.5, 1.5, p.shape) if name.endswith("weight")
                                          else rs.standard_normal(p.shape) * 0.1).astype(np.float32)))
        # keep both heads alive (a negative BatchNorm shift would zero them behind the ReLU and the test would only see the FC bias)
        net.policy_head.bn_layer.bias.copy_(torch.tensor([0.6, 0.4]))
        net.value_head.bn_layer.bias.copy_(torch.tensor([0.5]))
    save_model(net, out_bin)
    loaded = load_network(out_bin, False)
    planes = np.load(board_npz)["planes"][:16]
    x = torch.from_numpy(planes)
    logits, _ = loaded.inference_with_policy_logits(x)
    pol, val = loaded.inference(x)
    assert logits.numpy().std(axis=0).min() > 1e-3 and val.numpy().std(axis=0).min() > 1e-4, "dead head: outputs do not depend on the position"
    np.savez_compressed(out_npz, size=size, seed=seed, planes=planes, logits=logits.numpy(), policy_softmax=pol.numpy(),
                        value_softmax=val.numpy(), names=np.array(list(loaded.state_dict().keys())))
    print(f"model.bin golden: {os.path.getsize(out_bin)} bytes, {planes.shape[0]} positions -> {out_bin}, {out_npz}")


def gen_analysis(size, seed, out):
    """GTP analysis surface: lz-analyze / cgos-analyze strings (mcts/node.py:399-482), PV lists
    (mcts/tree.py:432-473) and tamago-dump_tree JSON (mcts/dump.py:10) of PUCT searches with injected net/noise."""
    import contextlib
    import io
    import json
    from board.go_board import GoBoard
    from board.stone import Stone
    from mcts.tree import MCTSTree
    from mcts.time_manager import TimeManager, TimeControl
    net = HashNet()
    patch = NoisePatch(seed=seed)
    movelists = positions_for_search(size, seed + 1, 5)
    ml_flat, ml_off = [], [0]
    for ml in movelists:
        ml_flat += ml; ml_off.append(len(ml_flat))
    meta, lz_stdout, lz, cgos, pv, dump = [], [], [], [], [], []
    for pi, ml in enumerate(movelists):
        b = GoBoard(board_size=size, komi=7.0, check_superko=True)
        color = Stone.BLACK
        for p in ml:
            b.put_stone(p, color); color = Stone.get_opponent_color(color)
        for visits, batch in ((60, 1), (90, 8)):
            tree = MCTSTree(net, tree_size=4096, batch_size=batch)
            patch.key(pi, b.moves)
            tm = TimeManager(TimeControl.CONSTANT_PLAYOUT, constant_visits=visits)
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf), contextlib.redirect_stderr(io.StringIO()):
                mv = tree.search_best_move(b, color, tm, {"mode": "lz", "interval": 0})
            root = tree.get_root()
            meta.append([pi, visits, batch, mv, color.value, tree.num_nodes])
            lz_stdout.append(buf.getvalue())
            lz.append(root.get_analysis(b, "lz", tree.get_pv_lists))
            cgos.append(root.get_analysis(b, "cgos", tree.get_pv_lists))
            pv.append(json.dumps(tree.get_pv_lists(root, b.coordinate)))
            dump.append(tree.dump_to_json(b, True))
    from board.zobrist_hash import hash_bit_mask
    np.savez_compressed(out, size=size, seed=seed, zobrist=np.asarray(hash_bit_mask, np.uint64),
                        movelist=np.array(ml_flat, np.int16), movelist_off=np.array(ml_off, np.int64),
                        meta=np.array(meta, np.int64), lz_stdout=np.array(lz_stdout), lz=np.array(lz), cgos=np.array(cgos),
                        pv=np.array(pv), dump=np.array(dump))
    print(f"analysis golden: {len(meta)} cases -> {out}")


def gen_selfplay(size, seed, out, visits, n_games):
    """selfplay/worker.py loop with injected net/noise; keeps the SGF text."""
    import tempfile
    from board.go_board import GoBoard, copy_board
    from board.stone import Stone
    from board.constant import PASS, RESIGN
    from mcts.tree import MCTSTree
    from mcts.time_manager import TimeManager, TimeControl
    from sgf.selfplay_record import SelfPlayRecord
    from board.zobrist_hash import hash_bit_mask
    net = HashNet()
    patch = NoisePatch(seed=seed)
    texts, movesl, nres = [], [], []
    tmp = tempfile.mkdtemp()
    for g in range(n_games):
        board = GoBoard(board_size=size, komi=7.0, check_superko=True)
        record = SelfPlayRecord(tmp, board.coordinate)
        mcts = MCTSTree(net, tree_size=160)
        tm = TimeManager(TimeControl.CONSTANT_PLAYOUT, constant_visits=visits)
        color = Stone.BLACK
        never_resign = (g % 2 == 0)
        pass_count, is_resign, score, winner = 0, False, 0.0, Stone.EMPTY
        moves = []
        for _ in range(size * size * 2):
            patch.key(g, board.moves)
            pos = mcts.generate_move_with_sequential_halving(board=board, color=color, time_manager=tm,
                                                             never_resign=never_resign)
            if pos == RESIGN:
                winner = Stone.get_opponent_color(color); is_resign = True
                break
            board.put_stone(pos, color)
            moves.append(pos)
            pass_count = pass_count + 1 if pos == PASS else 0
            record.save_record(mcts.get_root(), pos, color)
            color = Stone.get_opponent_color(color)
            if pass_count == 2:
                winner = Stone.EMPTY
                break
        if pass_count == 2:
            score = board.count_score() - board.get_komi()
            winner = Stone.BLACK if score > 0.1 else (Stone.WHITE if score < -0.1 else Stone.OUT_OF_BOARD)
        record.set_index(g)
        record.write_record(winner, board.get_komi(), is_resign, score)
        texts.append(open(os.path.join(tmp, f"{g}.sgf"), encoding="utf-8").read())
        movesl.append(moves); nres.append(int(never_resign))
        print(f"  selfplay game {g}: {len(moves)} moves, winner {winner}, score {score}")
    shutil.rmtree(tmp)
    np.savez_compressed(out, size=size, seed=seed, visits=visits, zobrist=np.asarray(hash_bit_mask, np.uint64),
                        sgf=np.array(texts), never_resign=np.array(nres),
                        moves=np.concatenate([np.array(m, np.int16) for m in movesl]),
                        moves_off=np.cumsum([0] + [len(m) for m in movesl]))
    print(f"selfplay golden -> {out}")


def numpy_weights(size, seed):
    """Seeded numpy state_dict for DualNet (names: SURVEY.md A.2)."""
    rs = np.random.RandomState(seed)
    sd = {}

    def conv(name, o, i, k):
        sd[name] = (rs.standard_normal((o, i, k, k)) * np.sqrt(2.0 / (i * k * k))).astype(np.float32)

    def bn(prefix, c):
        sd[prefix + ".weight"] = rs.uniform(0.5, 1.5, c).astype(np.float32)
        sd[prefix + ".bias"] = (rs.standard_normal(c) * 0.1).astype(np.float32)
        sd[prefix + ".running_mean"] = (rs.standard_normal(c) * 0.1).astype(np.float32)
        sd[prefix + ".running_var"] = rs.uniform(0.5, 1.5, c).astype(np.float32)
        sd[prefix + ".num_batches_tracked"] = np.array(0, np.int64)

    conv("conv_layer.weight", 64, 6, 3); bn("bn_layer", 64)
    for b in range(6):
        conv(f"blocks.{b}.conv1.weight", 64, 64, 3); conv(f"blocks.{b}.conv2.weight", 64, 64, 3)
        bn(f"blocks.{b}.bn1", 64); bn(f"blocks.{b}.bn2", 64)
    nn_ = size * size
    conv("policy_head.conv_layer.weight", 2, 64, 1); bn("policy_head.bn_layer", 2)
    sd["policy_head.fc_layer.weight"] = (rs.standard_normal((nn_ + 1, 2 * nn_)) * np.sqrt(1.0 / (2 * nn_))).astype(np.float32)
    sd["policy_head.fc_layer.bias"] = (rs.standard_normal(nn_ + 1) * 0.1).astype(np.float32)
    conv("value_head.conv_layer.weight", 1, 64, 1); bn("value_head.bn_layer", 1)
    sd["value_head.fc_layer.weight"] = (rs.standard_normal((3, nn_)) * np.sqrt(1.0 / nn_)).astype(np.float32)
    sd["value_head.fc_layer.bias"] = (rs.standard_normal(3) * 0.1).astype(np.float32)
    return sd


def train_data_set(size, seed, samples):
    """Seeded synthetic rl_data npz (layout of nn/data_generator.py:16-33): one-hot stone planes, a last-move plane, colour
    plane, improved-policy-like targets (softmax of Gumbel noise over a random support, 1e-18 elsewhere) and value labels."""
    rs = np.random.RandomState(seed)
    nn_ = size * size
    cls = rs.randint(0, 3, (samples, size, size))
    x = np.zeros((samples, 6, size, size), np.float32)
    for c in range(3):
        x[:, c] = cls == c
    last = rs.randint(0, nn_, samples)
    x.reshape(samples, 6, nn_)[np.arange(samples), 3, last] = 1.0
    x[:, 5] = np.where(rs.rand(samples) < 0.5, 1.0, -1.0)[:, None, None]
    logits = rs.gumbel(size=(samples, nn_ + 1)) * 2.0
    support = rs.rand(samples, nn_ + 1) < 0.4
    support[:, -1] = True
    p = np.where(support, np.exp(logits), 0.0)
    p = p / p.sum(axis=1, keepdims=True)
    policy = np.where(support, p, 1e-18).astype(np.float64)
    value = rs.randint(0, 3, samples).astype(np.int32)
    return dict(input=x, policy=policy, value=value, kifu_count=np.array(samples // 8))


def gen_train(size, out, weight_seed=515, data_seed=616, perm_seed=717, batch=64, steps=4):
    """K mini-batches of the reference's own train_with_gumbel_alphazero_on_cpu (nn/learn.py:234-315) from a seeded model
    and data set; records the per-step losses and, per tensor of the saved rl-model.bin, 64 sampled entries + the L1 norm."""
    import tempfile
    import torch
    import nn.learn as learn
    torch.set_num_threads(2)
    tmp = tempfile.mkdtemp()
    os.makedirs(os.path.join(tmp, "data")); os.makedirs(os.path.join(tmp, "model"))
    samples = batch * steps
    np.savez_compressed(os.path.join(tmp, "data", "rl_data_0.npz"), **train_data_set(size, data_seed, samples))
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in numpy_weights(size, weight_seed).items()}
    torch.save(sd, os.path.join(tmp, "model", "rl-model.bin"))
    losses = []
    orig_kld, orig_val = learn.calculate_policy_kld_loss, learn.calculate_value_loss
    state = {}

    def kld(o, t):
        state["p"] = orig_kld(o, t)
        return state["p"]

    def val(o, t):
        v = orig_val(o, t)
        losses.append([float((state["p"] + v).mean()), float(state["p"]), float(v.mean())])
        return v
    learn.calculate_policy_kld_loss, learn.calculate_value_loss = kld, val          # pass-through probes (values only)
    np.random.seed(perm_seed)
    torch.set_grad_enabled(True)
    learn.train_with_gumbel_alphazero_on_cpu(tmp, size, batch)
    learn.calculate_policy_kld_loss, learn.calculate_value_loss = orig_kld, orig_val
    final = torch.load(os.path.join(tmp, "model", "rl-model.bin"))
    rs = np.random.RandomState(1)
    names, idxs, vals, sums = [], [], [], []
    for k, v in final.items():
        t = v.double().reshape(-1).numpy()
        idx = rs.randint(0, 1 << 30, 64)
        sel = idx[:min(len(t), 64)] % len(t)
        pad = np.zeros(64); pad[:len(sel)] = t[sel]
        names.append(k); idxs.append(idx); vals.append(pad); sums.append(np.abs(t).sum())
    np.savez_compressed(out, size=size, weight_seed=weight_seed, data_seed=data_seed, perm_seed=perm_seed, batch=batch, steps=steps,
                        samples=samples, losses=np.array(losses), names=np.array(names), sample_idx=np.array(idxs),
                        sample_val=np.array(vals), abs_sum=np.array(sums))
    shutil.rmtree(tmp)
    print(f"train golden: {steps} steps, losses {np.array(losses)[:, 0]} -> {out}")


def gen_dualnet(size, seed, board_npz, out):
    import torch
    from nn.network.dual_net import DualNet
    torch.set_grad_enabled(False)
    net = DualNet(torch.device("cpu"), board_size=size)
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in numpy_weights(size, seed).items()}
    net.load_state_dict(sd)
    net.eval()
    planes = np.load(board_npz)["planes"][:24]
    x = torch.from_numpy(planes)
    logits, vlogit = net.forward(x)
    pol_sm, val_sm = net.inference(x)
    np.savez_compressed(out, size=size, weight_seed=seed, planes=planes, logits=logits.numpy(),
                        value_logits=vlogit.numpy(), policy_softmax=pol_sm.numpy(), value_softmax=val_sm.numpy())
    print(f"dualnet golden: {planes.shape[0]} positions -> {out}")


def gen_rldata(size, selfplay_npz, out):
    """nn/data_generator.py:89-149 on the golden self-play SGFs (BATCH_SIZE patched to 8 so that 24 samples are written)."""
    import tempfile
    import nn.data_generator as dg
    g = np.load(selfplay_npz)
    tmp = tempfile.mkdtemp()
    kdir = os.path.join(tmp, "kifu"); os.makedirs(kdir); os.makedirs(os.path.join(tmp, "data"))
    for i, text in enumerate(g["sgf"]):
        open(os.path.join(kdir, f"{i}.sgf"), "w", encoding="utf-8").write(str(text))
    order = []
    orig_shuffle = random.shuffle

    def shuffle(lst):
        lst.sort()
        orig_shuffle(lst)
        order.extend(int(os.path.basename(p).split(".")[0]) for p in lst)
    dg.random.shuffle = shuffle
    dg.BATCH_SIZE = 8
    random.seed(4242); np.random.seed(4242)
    dg.generate_reinforcement_learning_data(tmp, [kdir], size)
    dg.random.shuffle = orig_shuffle
    d = np.load(os.path.join(tmp, "data", "rl_data_0.npz"))
    np.savez_compressed(out, size=size, seed=4242, order=np.array(order), input=d["input"], policy=d["policy"],
                        value=d["value"], kifu_count=d["kifu_count"])
    shutil.rmtree(tmp)
    print(f"rl data golden: {len(d['value'])} samples, order {order} -> {out}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=9)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    ref = prepare_reference(a.size)
    np.random.seed(20261017)          # BEFORE importing board.* (Zobrist table is drawn at import)
    random.seed(20261017)
    sys.path.insert(0, ref)
    import torch
    torch.manual_seed(0)
    N = a.size
    want = lambda k: (not a.only) or k in a.only.split(",")
    if want("eye") and N == 9:
        from board.go_board import GoBoard
        b = GoBoard(board_size=9)
        eye = np.array([s.value for s in b.pattern.eye], np.uint8)
        np.savez_compressed(os.path.join(HERE, "eye_table.npz"), eye=eye)
    if want("board"):
        gen_board(N, 24 if N == 9 else 3, 1234, os.path.join(HERE, f"board_{N}.npz"))
    if want("search"):
        if N == 9:
            gen_search(N, 77, os.path.join(HERE, f"search_{N}.npz"), [16, 50, 400], [(100, 1), (120, 8)])
        else:
            gen_search(N, 77, os.path.join(HERE, f"search_{N}.npz"), [50], [(40, 1)])
    if want("analysis") and N == 9:
        gen_analysis(N, 91, os.path.join(HERE, "analysis_9.npz"))
    if want("selfplay") and N == 9:
        gen_selfplay(N, 5, os.path.join(HERE, "selfplay_9.npz"), visits=16, n_games=3)
    if want("rldata") and N == 9:
        gen_rldata(N, os.path.join(HERE, "selfplay_9.npz"), os.path.join(HERE, "rldata_9.npz"))
    if want("dualnet"):
        gen_dualnet(N, 31337, os.path.join(HERE, f"board_{N}.npz"), os.path.join(HERE, f"dualnet_{N}.npz"))
    if want("search2"):
        out = os.path.join(HERE, f"search2_{N}.npz")
        if N == 9:      # non-dyadic evaluator at every budget the 9x9 goldens use (fp32 queue-order accumulation)
            gen_search2(N, 78, out, 6, [(16, 1), (50, 1), (400, 1)], [(100, 1, 1, 0), (120, 8, 1, 0), (400, 8, 1, 1)])
        elif N == 13:   # 13x13 had no reference pin at all
            gen_search2(N, 79, out, 3, [(50, 0), (400, 1)], [(100, 1, 1, 0), (120, 8, 0, 0)])
        else:           # BASELINE configs[3] (PUCT-400 + super-ko, batch 1 and 8), SH-400, configs[4] (PUCT-1600, batch 256)
            ml = positions_for_search(N, 78, 2)
            mid = positions_for_search(N, 80, 4)
            fixed = [ml[0], ml[1], next(m for m in mid[1:] if len(m) >= 60)[:60]]
            gen_search2(N, 80, out, 0, [(400, 1)], [(400, 1, 1, 0), (400, 8, 1, 0), (1600, 256, 1, 0), (1600, 256, 0, 1)],
                        fixed_positions=fixed)
    if want("train") and N == 9:
        gen_train(N, os.path.join(HERE, "train_9.npz"))
    if want("modelbin") and N == 9:
        gen_modelbin(N, 2026, os.path.join(HERE, "board_9.npz"), os.path.join(HERE, "model_ref_9.bin"),
                     os.path.join(HERE, "model_ref_9.npz"))


if __name__ == "__main__":
    main()
