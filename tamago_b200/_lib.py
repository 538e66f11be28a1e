"""ctypes binding of libtamago_b200.so (the C ABI in include/tamago_b200.h).

The library is built in-tree by tamago_b200/build.py.  There is no CPU implementation behind this
package: if the shared library or a CUDA device is missing, loading / engine creation raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libtamago_b200.so")

u8p, i16p, i32p, i64p = C.POINTER(C.c_uint8), C.POINTER(C.c_int16), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
u64p, f32p, f64p = C.POINTER(C.c_uint64), C.POINTER(C.c_float), C.POINTER(C.c_double)


class Config(C.Structure):
    _fields_ = [("board_size", C.c_int32), ("komi", C.c_float), ("superko", C.c_int32), ("games", C.c_int32),
                ("max_visits", C.c_int32), ("batch_size", C.c_int32), ("max_nodes", C.c_int32), ("device", C.c_int32),
                ("evaluator", C.c_int32), ("dedup", C.c_int32), ("cgos_mode", C.c_int32), ("net_blocks", C.c_int32),
                ("seed", C.c_uint64), ("record_ring", C.c_int32), ("sample_cap", C.c_int32), ("scoring", C.c_int32)]


class Weights(C.Structure):
    _fields_ = [("conv_w", f32p), ("bn", f32p), ("bn_eps", C.c_float),
                ("block_conv_w", f32p), ("block_bn", f32p), ("block_bn_eps", C.c_float),
                ("policy_conv_w", f32p), ("policy_bn", f32p), ("policy_fc_w", f32p), ("policy_fc_b", f32p),
                ("value_conv_w", f32p), ("value_bn", f32p), ("value_fc_w", f32p), ("value_fc_b", f32p),
                ("head_bn_eps", C.c_float)]


class PlyDump(C.Structure):
    _fields_ = [("color", u8p), ("libs", i16p), ("size", i16p), ("scal", i32p), ("hash", u64p), ("legal", u8p),
                ("satari", i16p), ("eye", u8p), ("cand", u8p), ("score", i32p), ("plies", C.c_int32), ("tt_score", i32p)]


class StepResult(C.Structure):
    _fields_ = [("move", i32p), ("color", i32p), ("num_children", i32p), ("action", i16p), ("improved", f64p),
                ("visits", i32p), ("finished", i32p), ("winner", i32p), ("resigned", i32p), ("score", f32p),
                ("error", i32p), ("evals", i64p), ("n_moves", i32p)]


class NodeView(C.Structure):
    _fields_ = [("num_children", C.c_int32), ("node_visits", C.c_int32), ("virtual_loss", C.c_int32),
                ("node_value_sum", C.c_float), ("raw_value", C.c_float),
                ("action", i16p), ("children_index", i32p), ("children_value", f32p), ("children_visits", i32p),
                ("children_policy", f64p), ("children_virtual_loss", i32p), ("children_value_sum", f32p), ("noise", f64p)]


EXPORTS = {
    "tg_last_error": (C.c_char_p, []),
    "tg_action_stride": (C.c_int, [C.c_int]),
    "tg_engine_create": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "tg_engine_destroy": (None, [C.c_void_p]),
    "tg_set_zobrist": (C.c_int, [C.c_void_p, u64p]),
    "tg_load_weights": (C.c_int, [C.c_void_p, C.POINTER(Weights)]),
    "tg_load_weights_device": (C.c_int, [C.c_void_p, C.POINTER(Weights)]),
    "tg_reset": (C.c_int, [C.c_void_p, u8p, u64p, u8p]),
    "tg_play": (C.c_int, [C.c_void_p, i16p, u8p, i32p, C.c_int32, C.POINTER(PlyDump)]),
    "tg_set_to_move": (C.c_int, [C.c_void_p, i32p]),
    "tg_planes": (C.c_int, [C.c_void_p, f32p]),
    "tg_forward": (C.c_int, [C.c_void_p, f32p, C.c_int32, C.c_int32, f32p, f32p]),
    "tg_eval_buffers": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), i32p]),
    "tg_forward_device": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    "tg_stream": (C.c_void_p, [C.c_void_p]),
    "tg_sync": (C.c_int, [C.c_void_p]),
    "tg_genmove": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(StepResult)]),
    "tg_genmove_async": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "tg_collect": (C.c_int, [C.c_void_p, C.POINTER(StepResult)]),
    "tg_fetch_records": (C.c_int, [C.c_void_p, i32p, C.c_int32]),
    "tg_format_records": (C.c_int64, [C.c_void_p, C.c_char_p, C.c_int64, i64p]),
    "tg_write_records": (C.c_int64, [C.c_void_p, C.c_char_p, i64p]),
    "tg_fetched_record": (C.c_int, [C.c_void_p, C.c_int32, i32p, i16p, u8p, i16p, i16p, f64p]),
    "tg_emit_samples": (C.c_int64, [C.c_void_p, i32p, C.c_int32, i32p, i32p]),
    "tg_sample_buffers": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), i64p, i64p]),
    "tg_samples_read": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, f32p, f64p, i32p, C.c_int32]),
    "tg_samples_clear": (C.c_int, [C.c_void_p]),
    "tg_round_policy": (None, [f64p, C.c_int64]),
    "tg_stream_wait": (C.c_int, [C.c_void_p, C.c_void_p]),
    "tg_stream_signal": (C.c_int, [C.c_void_p, C.c_void_p]),
    "tg_tree_size": (C.c_int, [C.c_void_p, C.c_int32, i32p]),
    "tg_read_node": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(NodeView)]),
    "tg_format_sgf": (C.c_int, [C.c_int32, C.c_int32, i32p, i32p, i32p, i16p, f64p, C.c_int32, C.c_int32, C.c_int32,
                                C.c_double, C.c_double, C.c_char_p, C.c_int32]),
    "tg_launch_count": (C.c_int64, [C.c_void_p]),
    "tg_last_device_ms": (C.c_float, [C.c_void_p]),
    "tg_bench_kernel": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int32, C.c_int32, f32p]),
}

_lib = None


def load():
    """Load libtamago_b200.so; raises if it has not been built (python -m tamago_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m tamago_b200.build` "
                               "(tamago_b200 has no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


class EngineError(RuntimeError):
    pass


def check(rc):
    if rc < 0:
        raise EngineError(f"tamago_b200 error {rc}: {load().tg_last_error().decode()}")
    return rc
