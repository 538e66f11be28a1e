"""DualNet kernel probe (development aid): throughput at two batch sizes + per-layer clock64 stamps of CTA 0."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tamago_b200 as tb
from tamago_b200.nn.utility import random_init_state_dict
size = int(sys.argv[1]) if len(sys.argv) > 1 else 9
flop = {9: 72281646, 13: 0, 19: 322548446}[size]
e = tb.Engine(board_size=size, games=256, max_visits=256 if size == 9 else 64, evaluator=tb.EVAL_DUALNET_TC)
e.load_state_dict(random_init_state_dict(size, 0))
for b in ((7400, 65536) if size == 9 else (1480, 16384)):
    ms = e.bench_kernel("dualnet", b, 3)
    print(f"size {size} B={b}: {ms:.3f} ms  {b / ms * 1e3 / 1e6:.3f} M evals/s  {b * flop / ms / 1e9:.1f} TFLOP/s algorithmic", flush=True)
e.bench_kernel("dualnet_dbg", 7400 if size == 9 else 1480, 1)
