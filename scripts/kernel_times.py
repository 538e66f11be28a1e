"""Real (not serialised, warm) per-kernel durations of one configuration through torch.profiler / CUPTI (development aid).
usage: kernel_times.py c5 [steps]"""
import collections
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tamago_b200 as tb
from tamago_b200.nn.utility import random_init_state_dict

CFG = {"c4": (19, 1024, 400, tb.MODE_PUCT, 1, False, True, True), "c5": (19, 1, 1600, tb.MODE_PUCT, 256, True, False, False),
       "c2": (9, 4096, 400, tb.MODE_SH, 1, False, True, True)}
name = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
size, games, visits, mode, batch, strict, play, superko = CFG[name]
e = tb.Engine(board_size=size, games=games, max_visits=visits, batch_size=batch, superko=superko, evaluator=tb.EVAL_DUALNET_TC, seed=1)
e.load_state_dict(random_init_state_dict(size, 0))
e.reset(never_resign=np.ones(games, np.uint8))
e.genmove(mode=mode, visits=visits, strict=strict, play=play, full=False)
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        e.genmove(mode=mode, visits=visits, strict=strict, play=play, full=False)
    torch.cuda.synchronize()
agg = collections.OrderedDict()
t0, t1 = None, None
for ev in prof.events():
    if ev.device_type != torch.autograd.DeviceType.CUDA:
        continue
    k = ev.name.split("(")[0].replace("void ", "").replace("tg::", "")
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += ev.device_time
    t0 = ev.time_range.start if t0 is None else min(t0, ev.time_range.start)
    t1 = ev.time_range.end if t1 is None else max(t1, ev.time_range.end)
tot = sum(a[1] for a in agg.values())
for k, (n, us) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:44s} {n / steps:8.1f} launches/step {us / steps / 1e3:9.3f} ms/step  avg {us / n:8.1f} us")
print(f"kernel time {tot / steps / 1e3:.3f} ms/step, span {(t1 - t0) / steps / 1e3:.3f} ms/step, last_device_ms {e.last_device_ms:.3f}")
