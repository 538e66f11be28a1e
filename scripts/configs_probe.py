"""Throughput of the BASELINE.json configurations other than the bench line (development aid; numbers go to DESIGN.md)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tamago_b200 as tb
from tamago_b200.nn.utility import random_init_state_dict


def run(name, size, games, visits, mode, batch=1, dedup=False, steps=3, superko=True, strict=False, play=True):
    e = tb.Engine(board_size=size, games=games, max_visits=visits, batch_size=batch, superko=superko,
                  evaluator=tb.EVAL_DUALNET_TC, dedup=dedup, seed=1)
    e.load_state_dict(random_init_state_dict(size, 0))
    e.reset(never_resign=np.ones(games, np.uint8))
    e.genmove(mode=mode, visits=visits, strict=strict, play=play, full=False)
    ms, moves, evals, wall = 0.0, 0, 0, time.perf_counter()
    for _ in range(steps):
        r = e.genmove(mode=mode, visits=visits, strict=strict, play=play, full=False)
        assert (r["error"] == 0).all()
        ms += e.last_device_ms; moves += int((r["move"] >= 0).sum()); evals += int(r["evals"][1])
    wall = time.perf_counter() - wall
    e.close()
    return {"config": name, "ms_per_step": ms / steps, "moves_per_s": moves / (ms * 1e-3), "e2e_moves_per_s": moves / wall,
            "evals_per_step": evals / steps, "visits_per_s": moves * visits / (ms * 1e-3)}


out = [
    run("C3 9x9, 16384 games, 50-visit Gumbel SH (faithful)", 9, 16384, 50, tb.MODE_SH),
    run("C3 9x9, 16384 games, 50-visit Gumbel SH (dedup)", 9, 16384, 50, tb.MODE_SH, dedup=True),
    run("C4 19x19, 1024 games, 400-visit PUCT + super-ko, batch 1", 19, 1024, 400, tb.MODE_PUCT, steps=2),
    run("C4 19x19, 1024 games, 400-visit PUCT + super-ko, batch 8", 19, 1024, 400, tb.MODE_PUCT, batch=8, steps=2),
    run("19x19, 1024 games, 400-visit Gumbel SH (faithful)", 19, 1024, 400, tb.MODE_SH, steps=2),
    run("C5 19x19 genmove, 1600 visits, batch 256 (strict)", 19, 1, 1600, tb.MODE_PUCT, batch=256, strict=True, play=False, superko=False),
    run("C1-like 9x9, 1 game, 100-visit PUCT, batch 1", 9, 1, 100, tb.MODE_PUCT, steps=5),
]
print(json.dumps(out, indent=1))
