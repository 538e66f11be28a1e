"""Quick GPU probe: DualNet kernel time and one self-play step (not the bench; development aid)."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import tamago_b200 as tb
from make_golden import numpy_weights

size = int(sys.argv[1]) if len(sys.argv) > 1 else 9
games = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
visits = int(sys.argv[3]) if len(sys.argv) > 3 else 400
dedup = int(sys.argv[4]) if len(sys.argv) > 4 else 0
e = tb.Engine(board_size=size, games=games, max_visits=visits, evaluator=tb.EVAL_DUALNET_TC, dedup=bool(dedup), seed=1)
e.load_state_dict(numpy_weights(size, 1))
flop = 72281646 if size == 9 else 322548446
for slots in (740, 7400, 74000, 296000):
    if slots > games * visits: break
    ms = e.bench_kernel("dualnet", slots, 3)
    print(f"dualnet_tc {slots} slots: {ms:.3f} ms  -> {slots/ms*1e3:.0f} pos/s, {slots*flop/ms/1e9:.1f} TFLOP/s algorithmic")
for step in range(4):
    t0 = time.time()
    r = e.genmove(mode=tb.MODE_SH, visits=visits, play=True, full=True)
    dt = time.time() - t0
    ev_ms = e.bench_kernel("eval_ms"); ev_n = e.bench_kernel("eval_slots")
    print(f"step {step}: wall {dt*1e3:.1f} ms, device {e.last_device_ms:.1f} ms, eval {ev_ms:.1f} ms for {ev_n:.0f} evals, "
          f"errors {int((r['error']!=0).sum())}, moves {int((r['move']>=0).sum())}, launches {e.launches}")
print("planes kernel:", e.bench_kernel("planes", 0, 5), "ms")
