"""ctypes binding of the C oracle (oracle/tg_oracle.c) -- TEST INFRASTRUCTURE ONLY.

The oracle restates the reference algorithm (kobanium/TamaGo) on the CPU; it is
pinned against golden vectors produced by the reference itself
(tests/golden/make_golden.py).  Nothing under tamago_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libtg_oracle.so")

MAX_CELLS, MAX_ACTIONS, MAX_RECORDS = 441, 362, 1083
EMPTY, BLACK, WHITE, OB = 0, 1, 2, 3
PASS, RESIGN = 0, -1


def build(force=False):
    """Compile the oracle with the Makefile next to it (gcc only)."""
    src = [os.path.join(_HERE, f) for f in ("tg_oracle.c", "tg_oracle.h", "Makefile")]
    if force or not os.path.isfile(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "libtg_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


class Board(C.Structure):
    _fields_ = [
        ("n", C.c_int), ("w", C.c_int), ("cells", C.c_int), ("max_records", C.c_int),
        ("superko", C.c_int), ("komi", C.c_double),
        ("color", C.c_uint8 * MAX_CELLS),
        ("moves", C.c_int), ("ko_pos", C.c_int), ("ko_move", C.c_int),
        ("prisoner", C.c_int * 2),
        ("hash", C.c_uint64),
        ("hist_hash", C.c_uint64 * MAX_RECORDS),
        ("hist_pos", C.c_int16 * MAX_RECORDS),
        ("hist_color", C.c_uint8 * MAX_RECORDS),
        ("chain", C.c_int16 * MAX_CELLS),
        ("libs", C.c_int16 * MAX_CELLS),
        ("size", C.c_int16 * MAX_CELLS),
        ("zob", C.c_void_p),
    ]


class Node(C.Structure):
    _fields_ = [
        ("num_children", C.c_int), ("node_visits", C.c_int), ("virtual_loss", C.c_int),
        ("node_value_sum", C.c_float), ("raw_value", C.c_float),
        ("action", C.c_int16 * MAX_ACTIONS),
        ("children_index", C.c_int32 * MAX_ACTIONS),
        ("children_value", C.c_float * MAX_ACTIONS),
        ("children_visits", C.c_int32 * MAX_ACTIONS),
        ("children_policy", C.c_double * MAX_ACTIONS),
        ("children_virtual_loss", C.c_int32 * MAX_ACTIONS),
        ("children_value_sum", C.c_float * MAX_ACTIONS),
        ("noise", C.c_double * MAX_ACTIONS),
    ]


class GameRecord(C.Structure):
    _fields_ = [
        ("n_moves", C.c_int), ("winner", C.c_int), ("is_resign", C.c_int), ("score", C.c_double),
        ("pos", C.c_int16 * (2 * 361 + 2)), ("color", C.c_uint8 * (2 * 361 + 2)),
        ("num_children", C.c_int16 * (2 * 361 + 2)),
    ]


class HashNetCtx(C.Structure):
    _fields_ = [("n", C.c_int), ("variant", C.c_int)]


EVAL_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_float), C.c_int, C.c_int,
                      C.POINTER(C.c_float), C.POINTER(C.c_float))

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    u64p, i16p, f32p, f64p = C.POINTER(C.c_uint64), C.POINTER(C.c_int16), C.POINTER(C.c_float), C.POINTER(C.c_double)
    BP, NP = C.POINTER(Board), C.POINTER(Node)
    sig = {
        "tgo_default_zobrist": (None, [C.c_int, C.c_uint64, u64p]),
        "tgo_board_init": (None, [BP, C.c_int, C.c_double, C.c_int, u64p]),
        "tgo_board_clear": (None, [BP]),
        "tgo_board_copy": (None, [BP, BP]),
        "tgo_put_stone": (None, [BP, C.c_int, C.c_int]),
        "tgo_is_legal": (C.c_int, [BP, C.c_int, C.c_int]),
        "tgo_self_atari": (C.c_int, [BP, C.c_int, C.c_int]),
        "tgo_complete_eye": (C.c_int, [BP, C.c_int, C.c_int]),
        "tgo_eye_color": (C.c_int, [BP, C.c_int]),
        "tgo_candidates": (C.c_int, [BP, C.c_int, i16p]),
        "tgo_planes": (None, [BP, C.c_int, f32p]),
        "tgo_analyze": (None, [BP, C.c_int, C.POINTER(C.c_uint8), i16p, C.POINTER(C.c_uint8), C.POINTER(C.c_uint8)]),
        "tgo_count_score": (C.c_int, [BP]),
        "tgo_onboard_pos": (C.c_int, [BP, C.c_int]),
        "tgo_export_state": (None, [BP, C.POINTER(C.c_uint8), i16p, i16p, C.POINTER(C.c_int32), u64p]),
        "tgo_eye_table": (C.POINTER(C.c_uint8), []),
        "tgo_det_log": (C.c_double, [C.c_double]),
        "tgo_det_exp": (C.c_double, [C.c_double]),
        "tgo_noise_u": (C.c_double, [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]),
        "tgo_dirichlet": (None, [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, f64p]),
        "tgo_gumbel": (None, [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, f64p]),
        "tgo_np_sum": (C.c_double, [f64p, C.c_int]),
        "tgo_softmax": (None, [f64p, C.c_int, f64p, C.c_int]),
        "tgo_tree_new": (C.c_void_p, [C.c_int, C.c_int, C.c_int, C.c_int, EVAL_FN, C.c_void_p]),
        "tgo_tree_free": (None, [C.c_void_p]),
        "tgo_tree_set_noise_key": (None, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32]),
        "tgo_sh_schedule": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int]),
        "tgo_genmove_sh": (C.c_int, [C.c_void_p, BP, C.c_int, C.c_int, C.c_int]),
        "tgo_genmove_puct": (C.c_int, [C.c_void_p, BP, C.c_int, C.c_int, C.c_int]),
        "tgo_improved_policy": (None, [C.c_void_p, C.c_int, f64p]),
        "tgo_tree_node": (NP, [C.c_void_p, C.c_int]),
        "tgo_tree_set_use_libm": (None, [C.c_void_p, C.c_int]),
        "tgo_tree_evals": (C.c_long, [C.c_void_p]),
        "tgo_tree_error": (C.c_int, [C.c_void_p]),
        "tgo_tree_num_nodes": (C.c_int, [C.c_void_p]),
        "tgo_sizeof_node": (C.c_int, []),
        "tgo_sizeof_board": (C.c_int, []),
        "tgo_sizeof_record": (C.c_int, []),
        "tgo_selfplay_game": (C.c_int, [C.c_void_p, C.c_int, C.c_double, u64p, C.c_uint64, C.c_uint64,
                                        C.c_int, C.c_int, C.c_int, C.POINTER(GameRecord), f64p, i16p]),
        "tgo_hashnet_fn": (C.c_void_p, []),
        "tgo_ply_digest": (C.c_uint64, [BP]),
        "tgo_random_game": (C.c_int, [C.c_int, u64p, C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.c_double, C.c_double,
                                      i16p, u64p]),
        "tgo_tromp_taylor": (C.c_int, [BP]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    assert L.tgo_sizeof_node() == C.sizeof(Node), (L.tgo_sizeof_node(), C.sizeof(Node))
    assert L.tgo_sizeof_board() == C.sizeof(Board), (L.tgo_sizeof_board(), C.sizeof(Board))
    assert L.tgo_sizeof_record() == C.sizeof(GameRecord)
    _lib = L
    return L


def _p(arr, ctype):
    return arr.ctypes.data_as(C.POINTER(ctype))


def default_zobrist(n, seed=0x7A6D):
    out = np.zeros(4 * (n + 2) ** 2, dtype=np.uint64)
    lib().tgo_default_zobrist(n, seed, _p(out, C.c_uint64))
    return out.reshape(4, (n + 2) ** 2)


class OracleBoard:
    """Thin object wrapper: mirrors the GoBoard methods on the hot path."""

    def __init__(self, n, komi=7.0, superko=True, zobrist=None):
        self.n, self.w = n, n + 2
        self.cells = self.w * self.w
        self.zob = np.ascontiguousarray(default_zobrist(n) if zobrist is None else zobrist, dtype=np.uint64)
        assert self.zob.size == 4 * self.cells
        self.b = Board()
        lib().tgo_board_init(C.byref(self.b), n, komi, int(superko), _p(self.zob, C.c_uint64))
        self.onboard_pos = [lib().tgo_onboard_pos(C.byref(self.b), i) for i in range(n * n)]

    def put_stone(self, pos, color):
        lib().tgo_put_stone(C.byref(self.b), int(pos), int(color))

    def is_legal(self, pos, color):
        return bool(lib().tgo_is_legal(C.byref(self.b), int(pos), int(color)))

    def legal_mask(self, color):
        return self.analyze(color)[0]

    def analyze(self, color):
        """(legal, self_atari, complete_eye, candidate) per on-board point, raster order."""
        nn = self.n * self.n
        legal, eye, cand = np.zeros(nn, np.uint8), np.zeros(nn, np.uint8), np.zeros(nn, np.uint8)
        sa = np.zeros(nn, np.int16)
        lib().tgo_analyze(C.byref(self.b), int(color), _p(legal, C.c_uint8), _p(sa, C.c_int16),
                          _p(eye, C.c_uint8), _p(cand, C.c_uint8))
        return legal, sa, eye, cand

    def self_atari(self, pos, color):
        return lib().tgo_self_atari(C.byref(self.b), int(pos), int(color))

    def complete_eye(self, pos, color):
        return bool(lib().tgo_complete_eye(C.byref(self.b), int(pos), int(color)))

    def candidates(self, color):
        out = np.zeros(MAX_ACTIONS, dtype=np.int16)
        k = lib().tgo_candidates(C.byref(self.b), int(color), _p(out, C.c_int16))
        return out[:k].copy()

    def planes(self, color):
        out = np.zeros(6 * self.n * self.n, dtype=np.float32)
        lib().tgo_planes(C.byref(self.b), int(color), _p(out, C.c_float))
        return out.reshape(6, self.n, self.n)

    def count_score(self):
        return lib().tgo_count_score(C.byref(self.b))

    def tromp_taylor(self):
        """Area score Black - White by flood fill (the adjudication get_final_status.py:15-64 gets from GNU Go)."""
        return lib().tgo_tromp_taylor(C.byref(self.b))

    def ply_digest(self):
        return int(lib().tgo_ply_digest(C.byref(self.b)))

    def state(self):
        color = np.zeros(self.cells, np.uint8)
        libs = np.zeros(self.cells, np.int16)
        size = np.zeros(self.cells, np.int16)
        sc = np.zeros(6, np.int32)
        h = np.zeros(1, np.uint64)
        lib().tgo_export_state(C.byref(self.b), _p(color, C.c_uint8), _p(libs, C.c_int16), _p(size, C.c_int16),
                               _p(sc, C.c_int32), _p(h, C.c_uint64))
        return dict(color=color, libs=libs, size=size, moves=int(sc[0]), ko_pos=int(sc[1]), ko_move=int(sc[2]),
                    prisoner=(int(sc[3]), int(sc[4])), hash=int(h[0]))

    def copy(self):
        o = OracleBoard.__new__(OracleBoard)
        o.n, o.w, o.cells, o.zob, o.onboard_pos = self.n, self.w, self.cells, self.zob, self.onboard_pos
        o.b = Board()
        lib().tgo_board_copy(C.byref(o.b), C.byref(self.b))
        return o

    @property
    def moves(self):
        return self.b.moves

    @property
    def hash(self):
        return int(self.b.hash)


class OracleTree:
    """MCTSTree restatement with an injected evaluator.

    evaluator(planes[nb,6,n,n] f32, use_logit) -> (policy[nb,n*n+1] f32, value[nb,3] f32)
    """

    def __init__(self, n, evaluator, tree_size=4096, batch_size=1, cgos_mode=False, use_libm=False):
        self.n, self.A = n, n * n + 1
        self.evaluator = evaluator
        self.eval_calls = []
        if evaluator is hashnet or evaluator is hashnet2:
            # the C twin of the numpy hash evaluators (same bits; tests/test_oracle_search2.py checks it): no Python in the loop
            self._ctx = HashNetCtx(n, 1 if evaluator is hashnet2 else 0)
            self._cb = C.cast(lib().tgo_hashnet_fn(), EVAL_FN)
            self.t = lib().tgo_tree_new(n, tree_size, batch_size, int(cgos_mode), self._cb, C.cast(C.pointer(self._ctx), C.c_void_p))
            lib().tgo_tree_set_use_libm(self.t, int(use_libm))
            return

        def _cb(_ctx, planes, nb, use_logit, policy, value):
            x = np.ctypeslib.as_array(planes, shape=(nb, 6, n, n))
            pol, val = evaluator(x.copy(), bool(use_logit))
            np.ctypeslib.as_array(policy, shape=(nb, self.A))[:] = np.asarray(pol, dtype=np.float32)
            np.ctypeslib.as_array(value, shape=(nb, 3))[:] = np.asarray(val, dtype=np.float32)
            self.eval_calls.append(nb)

        self._cb = EVAL_FN(_cb)
        self.t = lib().tgo_tree_new(n, tree_size, batch_size, int(cgos_mode), self._cb, None)
        lib().tgo_tree_set_use_libm(self.t, int(use_libm))

    @property
    def evals(self):
        return lib().tgo_tree_evals(self.t)

    def __del__(self):
        if getattr(self, "t", None):
            lib().tgo_tree_free(self.t)
            self.t = None

    def set_noise_key(self, seed, game, move):
        lib().tgo_tree_set_noise_key(self.t, seed, game, move)

    def genmove_sh(self, board, color, visits, never_resign=True):
        return lib().tgo_genmove_sh(self.t, C.byref(board.b), color, visits, int(never_resign))

    def genmove_puct(self, board, color, visits, strict=False):
        return lib().tgo_genmove_puct(self.t, C.byref(board.b), color, visits, int(strict))

    @property
    def num_nodes(self):
        return lib().tgo_tree_num_nodes(self.t)

    def node(self, idx):
        nd = lib().tgo_tree_node(self.t, idx).contents
        k = nd.num_children
        f = lambda a, dt: np.frombuffer(a, dtype=dt)[:k].copy()
        return dict(num_children=k, node_visits=nd.node_visits, virtual_loss=nd.virtual_loss,
                    node_value_sum=float(nd.node_value_sum), raw_value=float(nd.raw_value),
                    action=f(nd.action, np.int16), children_index=f(nd.children_index, np.int32),
                    children_value=f(nd.children_value, np.float32), children_visits=f(nd.children_visits, np.int32),
                    children_policy=f(nd.children_policy, np.float64),
                    children_virtual_loss=f(nd.children_virtual_loss, np.int32),
                    children_value_sum=f(nd.children_value_sum, np.float32),
                    noise=np.frombuffer(nd.noise, dtype=np.float64)[:self.A].copy())

    def improved_policy(self, idx=0):
        k = self.node(idx)["num_children"]
        out = np.zeros(MAX_ACTIONS, np.float64)
        lib().tgo_improved_policy(self.t, idx, _p(out, C.c_double))
        return out[:k].copy()

    def selfplay_game(self, komi, zobrist, seed, game, visits, never_resign=False, use_puct=False):
        n = self.n
        rec = GameRecord()
        mm = 2 * n * n
        improved = np.zeros((mm, self.A), np.float64)
        actions = np.zeros((mm, self.A), np.int16)
        zob = np.ascontiguousarray(zobrist, dtype=np.uint64)
        r = lib().tgo_selfplay_game(self.t, n, komi, _p(zob, C.c_uint64), seed, game, visits, int(never_resign),
                                    int(use_puct), C.byref(rec), _p(improved, C.c_double), _p(actions, C.c_int16))
        if r < 0:
            raise RuntimeError("oracle self-play left the parity domain (history overflow)")
        m = rec.n_moves
        return dict(n_moves=m, winner=rec.winner, is_resign=bool(rec.is_resign), score=rec.score,
                    pos=np.array(rec.pos[:m]), color=np.array(rec.color[:m]),
                    num_children=np.array(rec.num_children[:m]), improved=improved[:m], actions=actions[:m])


def random_games(n, zobrist, seed, games, max_plies, p_pass=0.02, p_any_legal=0.3, superko=True, first_game=0):
    """Bulk differential corpus: `games` random games -> (moves [games, max_plies] int16, counts, digests [games, max_plies]).
    One digest covers everything observable after a ply (colours, liberties, sizes, ko, prisoners, hash, legality /
    self-atari / eye / candidate masks of both colours, count_score); numpy twin: ply_digest_np below."""
    zob = np.ascontiguousarray(zobrist, dtype=np.uint64)
    moves = np.zeros((games, max_plies), np.int16)
    dig = np.zeros((games, max_plies), np.uint64)
    counts = np.zeros(games, np.int32)
    L = lib()
    for g in range(games):
        counts[g] = L.tgo_random_game(n, _p(zob, C.c_uint64), int(superko), seed, first_game + g, max_plies, p_pass, p_any_legal,
                                      moves[g].ctypes.data_as(C.POINTER(C.c_int16)), dig[g].ctypes.data_as(C.POINTER(C.c_uint64)))
    return moves, counts, dig


def _dw(salt, count):
    with np.errstate(over="ignore"):
        return _mix64_np(np.uint64(salt) * np.uint64(0x100000001B3) + np.arange(count, dtype=np.uint64)) | np.uint64(1)


def ply_digest_np(d, n):
    """numpy twin of tgo_ply_digest over a per-ply state dump with the keys of tamago_b200.Engine.play(dump=True):
    color/libs/size [..., cells], scal [..., 5], hash [...], legal/satari/eye/cand [..., 2, n*n], score [...]."""
    cells, nn = (n + 2) ** 2, n * n
    with np.errstate(over="ignore"):
        acc = d["hash"].astype(np.uint64) * _dw(5, 1)[0]
        v = d["color"].astype(np.uint64) + np.uint64(4) * d["libs"].astype(np.uint64) + np.uint64(4096) * d["size"].astype(np.uint64)
        acc = acc + (v * _dw(1, cells)).sum(axis=-1, dtype=np.uint64)
        m = (d["legal"].astype(np.uint64) + np.uint64(2) * d["eye"].astype(np.uint64) + np.uint64(4) * d["cand"].astype(np.uint64)
             + np.uint64(8) * d["satari"].astype(np.int64).astype(np.uint64))
        acc = acc + (m.reshape(m.shape[:-2] + (2 * nn,)) * _dw(2, 2 * nn)).sum(axis=-1, dtype=np.uint64)
        acc = acc + (d["scal"].astype(np.int64).astype(np.uint64) * _dw(3, 5)).sum(axis=-1, dtype=np.uint64)
        acc = acc + d["score"].astype(np.int64).astype(np.uint64) * _dw(4, 1)[0]
    return acc


def sh_schedule(m, visits):
    cons = (C.c_int * 64)()
    cnts = (C.c_int * 64)()
    k = lib().tgo_sh_schedule(m, visits, cons, cnts, 64)
    return [(cons[i], cnts[i]) for i in range(k)]


def hashnet(planes, use_logit):
    """Deterministic stand-in network with exactly representable fp32 outputs.

    Used to compare search trees bit for bit across the reference (golden
    generation), this oracle and the CUDA engine without any NN arithmetic in
    the loop.  h = sum_j mix64(3*j + (v_j + 1)) over the 6*n*n plane values.
    """
    planes = np.asarray(planes, dtype=np.float32)
    nb = planes.shape[0]
    flat = planes.reshape(nb, -1)
    npl = flat.shape[1]
    nn = npl // 6
    j = np.arange(npl, dtype=np.uint64)
    code = (flat + 1.0).astype(np.uint64)
    with np.errstate(over="ignore"):
        h = _mix64_np(np.uint64(3) * j[None, :] + code).sum(axis=1, dtype=np.uint64)
        A = nn + 1
        idx = np.arange(A, dtype=np.uint64)
        r = _mix64_np(h[:, None] + idx[None, :])
        raw = ((r >> np.uint64(40)) & np.uint64(0xFFFF)).astype(np.float32)
        va = (_mix64_np(h + np.uint64(1000)) & np.uint64(0xFF)).astype(np.float32)
        vb = (_mix64_np(h + np.uint64(1001)) & np.uint64(0xFF)).astype(np.float32)
    if use_logit:
        pol = raw / np.float32(8192.0) - np.float32(4.0)
    else:
        pol = raw / np.float32(1048576.0)
    v0 = va / np.float32(512.0)
    v1 = vb / np.float32(512.0)
    v2 = np.float32(1.0) - v0 - v1
    return pol.astype(np.float32), np.stack([v0, v1, v2], axis=1).astype(np.float32)


def hashnet2(planes, use_logit):
    """hashnet with NON-dyadic fp32 outputs (values k/1000, logits raw/1000 - 4).

    Sums of hashnet's outputs are exact in fp32 and fp64 and in any order; these are not, so trees built with this
    evaluator pin the reference's fp32, queue-order value accumulation (mcts/tree.py:303-313, mcts/node.py:118-138).
    """
    planes = np.asarray(planes, dtype=np.float32)
    nb = planes.shape[0]
    flat = planes.reshape(nb, -1)
    npl = flat.shape[1]
    nn = npl // 6
    j = np.arange(npl, dtype=np.uint64)
    code = (flat + 1.0).astype(np.uint64)
    with np.errstate(over="ignore"):
        h = _mix64_np(np.uint64(3) * j[None, :] + code).sum(axis=1, dtype=np.uint64)
        idx = np.arange(nn + 1, dtype=np.uint64)
        r = _mix64_np(h[:, None] + idx[None, :])
        raw = ((r >> np.uint64(40)) & np.uint64(0xFFFF)).astype(np.float32)
        va = (_mix64_np(h + np.uint64(1000)) % np.uint64(500)).astype(np.float32)
        vb = (_mix64_np(h + np.uint64(1001)) % np.uint64(500)).astype(np.float32)
    if use_logit:
        pol = raw / np.float32(1000.0) - np.float32(4.0)
    else:
        pol = raw / np.float32(1000000.0)
    v0 = va / np.float32(1000.0)
    v1 = vb / np.float32(1000.0)
    v2 = np.float32(1.0) - v0 - v1
    return pol.astype(np.float32), np.stack([v0, v1, v2], axis=1).astype(np.float32)


def _mix64_np(z):
    z = z.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))
