"""Micro-benchmarks of SURVEY.md 8d on a B200 (development aid; results are summarised in profiles/):
K1 board step (tg_play replay of self-play move corpora), K3 feature planes, K4 DualNet forward at several batch sizes."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tamago_b200 as tb
from tamago_b200.nn.utility import random_init_state_dict

out = {}
for size, games in ((9, 16384), (19, 1024)):
    # corpus: moves of hash-evaluator self-play (2 visits) -- legal, varied, with captures and ko
    e = tb.Engine(board_size=size, games=games, max_visits=2, evaluator=tb.EVAL_HASHNET, seed=1234)
    e.reset(never_resign=np.ones(games, np.uint8))
    plies = 2 * size * size
    moves = np.zeros((games, plies), np.int16)
    counts = np.zeros(games, np.int32)
    for i in range(plies):
        r = e.genmove(mode=tb.MODE_SH, visits=2, play=True, full=False)
        live = r["move"] >= 0
        moves[live, counts[live]] = r["move"][live]
        counts[live] += 1
        if not live.any():
            break
    e.reset()
    t0 = time.perf_counter()
    e.play(moves, counts)
    dt = time.perf_counter() - t0
    out[f"K1_put_stone_{size}x{size}"] = {"games": games, "plies": int(counts.sum()), "seconds_incl_h2d": dt, "plies_per_s": float(counts.sum() / dt)}
    ms = e.bench_kernel("planes", 0, 1)
    e.close()
    en = tb.Engine(board_size=size, games=256, max_visits=1024 if size == 9 else 512, evaluator=tb.EVAL_DUALNET_TC)
    en.load_state_dict(random_init_state_dict(size, 0))
    flop = 72281646 if size == 9 else 322548446
    for b in (256, 1024, 65536, 262144):
        if b > 256 * (1024 if size == 9 else 512):
            continue
        ms = en.bench_kernel("dualnet", b, 3)
        out[f"K4_dualnet_{size}x{size}_B{b}"] = {"ms": ms, "evals_per_s": b / ms * 1e3, "tflops_algorithmic": b * flop / ms / 1e9}
    en.close()
print(json.dumps(out, indent=1))
