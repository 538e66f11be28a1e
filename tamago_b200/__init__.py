"""tamago_b200: B200-native self-play / search engine behind TamaGo's Python surface.

The package is a thin Python host over libtamago_b200.so (hand-written sm_100a CUDA behind a C ABI,
include/tamago_b200.h).  There is no CPU implementation; importing works anywhere, creating an Engine
needs the built library and a B200.
"""
from .engine import Engine, format_sgf, MODE_SH, MODE_PUCT, EVAL_DUALNET_TC, EVAL_DUALNET_FP32, EVAL_HASHNET, EVAL_HASHNET2  # noqa: F401
