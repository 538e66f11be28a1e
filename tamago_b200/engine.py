"""Engine: numpy-facing wrapper of the C ABI (one engine = one GPU's board pool, tree pool and DualNet)."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check

MODE_SH, MODE_PUCT = 0, 1
EVAL_DUALNET_TC, EVAL_DUALNET_FP32, EVAL_HASHNET, EVAL_HASHNET2 = 0, 1, 2, 3
PASS, RESIGN = 0, -1
BLACK, WHITE = 1, 2


def _ptr(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


class Engine:
    def __init__(self, board_size=9, games=1, max_visits=400, komi=7.0, superko=True, batch_size=1, max_nodes=0,
                 device=0, evaluator=EVAL_DUALNET_TC, dedup=False, cgos_mode=False, net_blocks=6, seed=0,
                 record_ring=False, scoring=0, sample_cap=0):
        self.lib = _lib.load()
        self.n, self.games = board_size, games
        self.nn, self.A = board_size * board_size, board_size * board_size + 1
        self.cells = (board_size + 2) ** 2
        self.stride = self.lib.tg_action_stride(board_size)
        self.komi, self.net_blocks, self.device = komi, net_blocks, device
        cfg = _lib.Config(board_size, komi, int(superko), games, max_visits, batch_size, max_nodes, device,
                          evaluator, int(dedup), int(cgos_mode), net_blocks, seed, int(record_ring), int(sample_cap), int(scoring))
        h = C.c_void_p()
        check(self.lib.tg_engine_create(C.byref(cfg), C.byref(h)))
        self.h = h
        self._keep = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.tg_engine_destroy(self.h)
            self.h = None

    __del__ = close

    # -- setup ------------------------------------------------------------------------------------
    def set_zobrist(self, table):
        t = np.ascontiguousarray(table, dtype=np.uint64).reshape(-1)
        assert t.size == 4 * self.cells
        check(self.lib.tg_set_zobrist(self.h, _ptr(t, C.c_uint64)))

    def load_state_dict(self, sd):
        """sd: mapping of DualNet.state_dict() names to arrays (numpy or torch tensors)."""
        def get(name):
            v = sd[name]
            if hasattr(v, "detach"):
                v = v.detach().cpu().numpy()
            return np.ascontiguousarray(np.asarray(v, dtype=np.float32))

        def bn(prefix):
            return np.ascontiguousarray(np.stack([get(prefix + ".weight"), get(prefix + ".bias"),
                                                  get(prefix + ".running_mean"), get(prefix + ".running_var")]))
        B = self.net_blocks
        keep = dict(
            conv_w=get("conv_layer.weight"), bn=bn("bn_layer"),
            block_conv_w=np.ascontiguousarray(np.stack([np.stack([get(f"blocks.{b}.conv1.weight"), get(f"blocks.{b}.conv2.weight")])
                                                        for b in range(B)])),
            block_bn=np.ascontiguousarray(np.stack([np.stack([bn(f"blocks.{b}.bn1"), bn(f"blocks.{b}.bn2")]) for b in range(B)])),
            policy_conv_w=get("policy_head.conv_layer.weight").reshape(2, 64), policy_bn=bn("policy_head.bn_layer"),
            policy_fc_w=get("policy_head.fc_layer.weight"), policy_fc_b=get("policy_head.fc_layer.bias"),
            value_conv_w=get("value_head.conv_layer.weight").reshape(1, 64), value_bn=bn("value_head.bn_layer"),
            value_fc_w=get("value_head.fc_layer.weight"), value_fc_b=get("value_head.fc_layer.bias"))
        assert keep["conv_w"].shape == (64, 6, 3, 3) and keep["block_conv_w"].shape == (B, 2, 64, 64, 3, 3)
        assert keep["policy_fc_w"].shape == (self.A, 2 * self.nn) and keep["value_fc_w"].shape == (3, self.nn)
        keep = {k: np.ascontiguousarray(v) for k, v in keep.items()}
        w = _lib.Weights()
        for k, v in keep.items():
            setattr(w, k, _ptr(v, C.c_float))
        w.bn_eps, w.block_bn_eps, w.head_bn_eps = 1e-5, 2e-5, 2e-5     # dual_net.py:32, res_block.py:23-24, head/*.py
        check(self.lib.tg_load_weights(self.h, C.byref(w)))

    def load_state_dict_device(self, sd):
        """The same from torch CUDA tensors on this engine's device (e.g. TrainDualNet.state_dict() right after a training
        step): the fold and packing run on the device (tg_load_weights_device), nothing crosses PCIe."""
        import torch
        dev = torch.device("cuda", self.device)

        def get(name):
            return sd[name].detach().to(device=dev, dtype=torch.float32).contiguous()

        def bn(prefix):
            return torch.stack([get(prefix + ".weight"), get(prefix + ".bias"), get(prefix + ".running_mean"), get(prefix + ".running_var")]).contiguous()
        B = self.net_blocks
        keep = dict(
            conv_w=get("conv_layer.weight"), bn=bn("bn_layer"),
            block_conv_w=torch.stack([torch.stack([get(f"blocks.{b}.conv1.weight"), get(f"blocks.{b}.conv2.weight")]) for b in range(B)]).contiguous(),
            block_bn=torch.stack([torch.stack([bn(f"blocks.{b}.bn1"), bn(f"blocks.{b}.bn2")]) for b in range(B)]).contiguous(),
            policy_conv_w=get("policy_head.conv_layer.weight").reshape(2, 64).contiguous(), policy_bn=bn("policy_head.bn_layer"),
            policy_fc_w=get("policy_head.fc_layer.weight"), policy_fc_b=get("policy_head.fc_layer.bias"),
            value_conv_w=get("value_head.conv_layer.weight").reshape(1, 64).contiguous(), value_bn=bn("value_head.bn_layer"),
            value_fc_w=get("value_head.fc_layer.weight"), value_fc_b=get("value_head.fc_layer.bias"))
        assert tuple(keep["policy_fc_w"].shape) == (self.A, 2 * self.nn) and tuple(keep["block_conv_w"].shape) == (B, 2, 64, 64, 3, 3)
        w = _lib.Weights()
        for k, v in keep.items():
            setattr(w, k, C.cast(C.c_void_p(v.data_ptr()), C.POINTER(C.c_float)))
        w.bn_eps, w.block_bn_eps, w.head_bn_eps = 1e-5, 2e-5, 2e-5
        torch.cuda.synchronize(dev)                      # the stacks above were queued on torch's stream
        check(self.lib.tg_load_weights_device(self.h, C.byref(w)))

    # -- boards -----------------------------------------------------------------------------------
    def reset(self, mask=None, game_ids=None, never_resign=None):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        ids = None if game_ids is None else np.ascontiguousarray(game_ids, dtype=np.uint64)
        nr = None if never_resign is None else np.ascontiguousarray(never_resign, dtype=np.uint8)
        check(self.lib.tg_reset(self.h, None if m is None else _ptr(m, C.c_uint8),
                                None if ids is None else _ptr(ids, C.c_uint64),
                                None if nr is None else _ptr(nr, C.c_uint8)))

    def play(self, moves, counts=None, colors=None, dump=False):
        """moves: [games, stride] int16 positions; counts: moves per game.  Returns the per-ply dump if asked."""
        mv = np.ascontiguousarray(moves, dtype=np.int16).reshape(self.games, -1)
        stride = mv.shape[1]
        cnt = np.full(self.games, stride, np.int32) if counts is None else np.ascontiguousarray(counts, dtype=np.int32)
        col = None if colors is None else np.ascontiguousarray(colors, dtype=np.uint8).reshape(self.games, stride)
        d, out = None, None
        if dump:
            g, p, n2, c = self.games, stride, self.nn, self.cells
            out = dict(color=np.zeros((g, p, c), np.uint8), libs=np.zeros((g, p, c), np.int16), size=np.zeros((g, p, c), np.int16),
                       scal=np.zeros((g, p, 5), np.int32), hash=np.zeros((g, p), np.uint64), legal=np.zeros((g, p, 2, n2), np.uint8),
                       satari=np.zeros((g, p, 2, n2), np.int16), eye=np.zeros((g, p, 2, n2), np.uint8),
                       cand=np.zeros((g, p, 2, n2), np.uint8), score=np.zeros((g, p), np.int32), tt_score=np.zeros((g, p), np.int32))
            d = _lib.PlyDump(_ptr(out["color"], C.c_uint8), _ptr(out["libs"], C.c_int16), _ptr(out["size"], C.c_int16),
                             _ptr(out["scal"], C.c_int32), _ptr(out["hash"], C.c_uint64), _ptr(out["legal"], C.c_uint8),
                             _ptr(out["satari"], C.c_int16), _ptr(out["eye"], C.c_uint8), _ptr(out["cand"], C.c_uint8),
                             _ptr(out["score"], C.c_int32), p, _ptr(out["tt_score"], C.c_int32))
        check(self.lib.tg_play(self.h, _ptr(mv, C.c_int16), None if col is None else _ptr(col, C.c_uint8),
                               _ptr(cnt, C.c_int32), stride, None if d is None else C.byref(d)))
        return out

    def set_to_move(self, colors):
        c = np.ascontiguousarray(np.broadcast_to(np.asarray(colors, np.int32), (self.games,)))
        check(self.lib.tg_set_to_move(self.h, _ptr(c, C.c_int32)))

    def planes(self):
        out = np.zeros((self.games, 6, self.n, self.n), np.float32)
        check(self.lib.tg_planes(self.h, _ptr(out, C.c_float)))
        return out

    def forward(self, planes, use_logit=True):
        x = np.ascontiguousarray(planes, dtype=np.float32).reshape(-1, 6, self.n, self.n)
        pol = np.zeros((x.shape[0], self.A), np.float32)
        val = np.zeros((x.shape[0], 3), np.float32)
        check(self.lib.tg_forward(self.h, _ptr(x, C.c_float), x.shape[0], int(use_logit), _ptr(pol, C.c_float), _ptr(val, C.c_float)))
        return pol, val

    # -- device-resident evaluation (zero copy) ------------------------------------------------------
    def eval_tensors(self):
        """torch views of the engine's evaluator batch on the device: planes [slot_cap, 6, N, N], policy [slot_cap, N*N+1],
        value [slot_cap, 3] (fp32).  They alias engine memory: fill `planes[:n]`, call forward_device(n), read the others."""
        import torch
        p, q, v, cap = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int32()
        check(self.lib.tg_eval_buffers(self.h, C.byref(p), C.byref(q), C.byref(v), C.byref(cap)))

        class _View:
            def __init__(self, ptr, shape):
                self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}
        dev = torch.device("cuda", self.device)
        mk = lambda ptr, shape: torch.as_tensor(_View(ptr.value, shape), device=dev)
        return (mk(p, (cap.value, 6, self.n, self.n)), mk(q, (cap.value, self.A)), mk(v, (cap.value, 3)))

    def forward_device(self, n, use_logit=True, torch_stream=None):
        """DualNet on the first n slots of the device batch, asynchronous on the engine's stream (see sync / stream).
        With torch_stream (a torch.cuda.Stream) the call is ordered after what that stream has queued (the fill of
        `planes`) and the stream's later work waits for the outputs: no host synchronisation on either side."""
        cs = None if torch_stream is None else C.c_void_p(torch_stream.cuda_stream)
        if cs is not None:
            check(self.lib.tg_stream_wait(self.h, cs))
        check(self.lib.tg_forward_device(self.h, int(n), int(use_logit)))
        if cs is not None:
            check(self.lib.tg_stream_signal(self.h, cs))

    @property
    def stream(self):
        return self.lib.tg_stream(self.h)

    def sync(self):
        check(self.lib.tg_sync(self.h))

    # -- search -----------------------------------------------------------------------------------
    def genmove_async(self, mode=MODE_SH, visits=400, strict=False, play=False, full=False):
        """Queue one move of every game and return at once; collect() waits for it."""
        self._pending_full = bool(full)
        check(self.lib.tg_genmove_async(self.h, mode, visits, int(strict), int(play), int(full)))

    def collect(self):
        g, s = self.games, self.stride
        full = getattr(self, "_pending_full", False)
        r = dict(move=np.zeros(g, np.int32), color=np.zeros(g, np.int32), num_children=np.zeros(g, np.int32),
                 finished=np.zeros(g, np.int32), winner=np.zeros(g, np.int32), resigned=np.zeros(g, np.int32),
                 score=np.zeros(g, np.float32), error=np.zeros(g, np.int32), evals=np.zeros(2, np.int64),
                 n_moves=np.zeros(g, np.int32))
        if full:
            r.update(action=np.zeros((g, s), np.int16), improved=np.zeros((g, s), np.float64), visits=np.zeros((g, s), np.int32))
        sr = _lib.StepResult()
        ct = dict(move=C.c_int32, color=C.c_int32, num_children=C.c_int32, action=C.c_int16, improved=C.c_double,
                  visits=C.c_int32, finished=C.c_int32, winner=C.c_int32, resigned=C.c_int32, score=C.c_float,
                  error=C.c_int32, evals=C.c_int64, n_moves=C.c_int32)
        for k, v in r.items():
            setattr(sr, k, _ptr(v, ct[k]))
        check(self.lib.tg_collect(self.h, C.byref(sr)))
        return r

    # -- training samples straight from the record ring (nn/data_generator.py:89-149) -----------------------
    def emit_samples(self, games, plies, syms):
        """games [n], plies [n, 8] (ascending move indices, -1 padded), syms [n, 8] -> new sample count"""
        gl = np.ascontiguousarray(games, dtype=np.int32)
        pl = np.ascontiguousarray(plies, dtype=np.int32).reshape(len(gl), 8)
        sy = np.ascontiguousarray(syms, dtype=np.int32).reshape(len(gl), 8)
        return int(check(self.lib.tg_emit_samples(self.h, _ptr(gl, C.c_int32), len(gl), _ptr(pl, C.c_int32), _ptr(sy, C.c_int32))))

    @property
    def sample_count(self):
        n, cap = C.c_int64(), C.c_int64()
        check(self.lib.tg_sample_buffers(self.h, None, None, None, C.byref(n), C.byref(cap)))
        return n.value

    def read_samples(self, first=0, n=None, round_like_sgf=True):
        """host copies: input [n, 6, N, N] f32, policy [n, N*N+1] f64, value [n] i32 (npz layout of data_generator.py:16-33)"""
        n = self.sample_count - first if n is None else n
        inp = np.zeros((n, 6, self.n, self.n), np.float32)
        pol = np.zeros((n, self.A), np.float64)
        val = np.zeros(n, np.int32)
        check(self.lib.tg_samples_read(self.h, first, n, _ptr(inp, C.c_float), _ptr(pol, C.c_double), _ptr(val, C.c_int32), int(round_like_sgf)))
        return inp, pol, val

    def sample_tensors(self):
        """torch views (zero copy) of the device-resident samples: input, policy (f64), value (i32), valid up to sample_count"""
        import torch
        pi, pp, pv, n, cap = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int64(), C.c_int64()
        check(self.lib.tg_sample_buffers(self.h, C.byref(pi), C.byref(pp), C.byref(pv), C.byref(n), C.byref(cap)))

        class _View:
            def __init__(self, ptr, shape, typestr):
                self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 3, "strides": None}
        dev = torch.device("cuda", self.device)
        k = max(1, n.value)
        mk = lambda ptr, shape, ts: torch.as_tensor(_View(ptr.value, shape, ts), device=dev)
        return (mk(pi, (k, 6, self.n, self.n), "<f4")[:n.value], mk(pp, (k, self.A), "<f8")[:n.value], mk(pv, (k,), "<i4")[:n.value])

    def clear_samples(self):
        check(self.lib.tg_samples_clear(self.h))

    # -- records of finished games (device ring -> SGF) ---------------------------------------------
    def fetch_records(self, games):
        gl = np.ascontiguousarray(games, dtype=np.int32)
        self._n_fetched = len(gl)
        check(self.lib.tg_fetch_records(self.h, _ptr(gl, C.c_int32), len(gl)))

    def format_records(self):
        """SGF text of every fetched game (sgf/selfplay_record.py:67-110)."""
        n = getattr(self, "_n_fetched", 0)
        off = np.zeros(n + 1, np.int64)
        cap = 1 << 16
        while True:
            buf = C.create_string_buffer(cap)
            rc = self.lib.tg_format_records(self.h, buf, cap, _ptr(off, C.c_int64))
            if rc >= 0:
                break
            if int(off[n]) > cap:
                cap = int(off[n]) + 16
                continue
            check(int(rc))
        return [buf.raw[off[i]:off[i + 1]].decode("utf-8") for i in range(n)]

    def write_records(self, save_dir, index):
        """<save_dir>/<index[i]>.sgf for every fetched game; returns the number of moves written."""
        idx = np.ascontiguousarray(index, dtype=np.int64)
        assert len(idx) == getattr(self, "_n_fetched", 0)
        return int(check(self.lib.tg_write_records(self.h, str(save_dir).encode(), _ptr(idx, C.c_int64))))

    def fetched_record(self, i):
        n = C.c_int32()
        check(self.lib.tg_fetched_record(self.h, i, C.byref(n), None, None, None, None, None))
        m, s = n.value, self.stride
        mv, col, k = np.zeros(m, np.int16), np.zeros(m, np.uint8), np.zeros(m, np.int16)
        act, imp = np.zeros((m, s), np.int16), np.zeros((m, s), np.float64)
        check(self.lib.tg_fetched_record(self.h, i, C.byref(n), _ptr(mv, C.c_int16), _ptr(col, C.c_uint8), _ptr(k, C.c_int16),
                                         _ptr(act, C.c_int16), _ptr(imp, C.c_double)))
        return dict(move=mv, color=col, num_children=k, action=act, improved=imp)

    def genmove(self, mode=MODE_SH, visits=400, strict=False, play=False, full=True):
        g, s = self.games, self.stride
        r = dict(move=np.zeros(g, np.int32), color=np.zeros(g, np.int32), num_children=np.zeros(g, np.int32),
                 finished=np.zeros(g, np.int32), winner=np.zeros(g, np.int32), resigned=np.zeros(g, np.int32),
                 score=np.zeros(g, np.float32), error=np.zeros(g, np.int32), evals=np.zeros(2, np.int64),
                 n_moves=np.zeros(g, np.int32))
        if full:
            r.update(action=np.zeros((g, s), np.int16), improved=np.zeros((g, s), np.float64), visits=np.zeros((g, s), np.int32))
        sr = _lib.StepResult()
        ct = dict(move=C.c_int32, color=C.c_int32, num_children=C.c_int32, action=C.c_int16, improved=C.c_double,
                  visits=C.c_int32, finished=C.c_int32, winner=C.c_int32, resigned=C.c_int32, score=C.c_float,
                  error=C.c_int32, evals=C.c_int64, n_moves=C.c_int32)
        for k, v in r.items():
            setattr(sr, k, _ptr(v, ct[k]))
        check(self.lib.tg_genmove(self.h, mode, visits, int(strict), int(play), C.byref(sr)))
        return r

    def tree_size(self, game):
        n = C.c_int32()
        check(self.lib.tg_tree_size(self.h, game, C.byref(n)))
        return n.value

    def node(self, game, index):
        s = self.stride
        a = dict(action=np.zeros(s, np.int16), children_index=np.zeros(s, np.int32), children_value=np.zeros(s, np.float32),
                 children_visits=np.zeros(s, np.int32), children_policy=np.zeros(s, np.float64),
                 children_virtual_loss=np.zeros(s, np.int32), children_value_sum=np.zeros(s, np.float32), noise=np.zeros(s, np.float64))
        nv = _lib.NodeView()
        ct = dict(action=C.c_int16, children_index=C.c_int32, children_value=C.c_float, children_visits=C.c_int32,
                  children_policy=C.c_double, children_virtual_loss=C.c_int32, children_value_sum=C.c_float, noise=C.c_double)
        for k, v in a.items():
            setattr(nv, k, _ptr(v, ct[k]))
        check(self.lib.tg_read_node(self.h, game, index, C.byref(nv)))
        k = nv.num_children
        out = {key: v[:k].copy() for key, v in a.items() if key != "noise"}
        out.update(num_children=k, node_visits=nv.node_visits, virtual_loss=nv.virtual_loss,
                   node_value_sum=float(nv.node_value_sum), raw_value=float(nv.raw_value), noise=a["noise"][:self.A].copy())
        return out

    # -- instrumentation --------------------------------------------------------------------------
    @property
    def launches(self):
        return int(self.lib.tg_launch_count(self.h))

    @property
    def last_device_ms(self):
        return float(self.lib.tg_last_device_ms(self.h))

    def bench_kernel(self, name, slots=0, iters=1):
        ms = C.c_float()
        check(self.lib.tg_bench_kernel(self.h, name.encode(), slots, iters, C.byref(ms)))
        return ms.value


def format_sgf(board_size, moves, colors, num_children, action, improved, winner, resigned, score, komi):
    """sgf/selfplay_record.py:67-110 text of one finished game."""
    lib = _lib.load()
    m = np.ascontiguousarray(moves, dtype=np.int32)
    c = np.ascontiguousarray(colors, dtype=np.int32)
    k = np.ascontiguousarray(num_children, dtype=np.int32)
    a = np.ascontiguousarray(action, dtype=np.int16).reshape(len(m), -1) if len(m) else np.zeros((0, 1), np.int16)
    p = np.ascontiguousarray(improved, dtype=np.float64).reshape(len(m), -1) if len(m) else np.zeros((0, 1), np.float64)
    cap = 512 + len(m) * (64 + 24 * (a.shape[1] + 1))
    buf = C.create_string_buffer(cap)
    n = check(lib.tg_format_sgf(board_size, len(m), _ptr(m, C.c_int32), _ptr(c, C.c_int32), _ptr(k, C.c_int32),
                                _ptr(a, C.c_int16), _ptr(p, C.c_double), a.shape[1], int(winner), int(resigned),
                                float(score), float(komi), buf, cap))
    return buf.raw[:n].decode("utf-8")
