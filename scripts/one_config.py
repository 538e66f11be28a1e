"""Runs a few moves of one BASELINE.json configuration (development aid, e.g. under ncu for a launch list).
usage: one_config.py c2|c3|c4|c4b8|c5|sh19 [steps]"""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tamago_b200 as tb
from tamago_b200.nn.utility import random_init_state_dict

CFG = {   # size, games, visits, mode, batch, strict, play, superko
    "c2": (9, 4096, 400, tb.MODE_SH, 1, False, True, True),
    "c3": (9, 16384, 50, tb.MODE_SH, 1, False, True, True),
    "c4": (19, 1024, 400, tb.MODE_PUCT, 1, False, True, True),
    "c4b8": (19, 1024, 400, tb.MODE_PUCT, 8, False, True, True),
    "c5": (19, 1, 1600, tb.MODE_PUCT, 256, True, False, False),
    "sh19": (19, 1024, 400, tb.MODE_SH, 1, False, True, True),
}
name = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dedup = len(sys.argv) > 3 and sys.argv[3] == "dedup"
size, games, visits, mode, batch, strict, play, superko = CFG[name]
games = int(os.environ.get('TG_GAMES', games)); superko = bool(int(os.environ.get('TG_SUPERKO', superko))); play = bool(int(os.environ.get('TG_PLAY', play)))
batch = int(os.environ.get('TG_BATCH', batch)); visits = int(os.environ.get('TG_VISITS', visits))
e = tb.Engine(board_size=size, games=games, max_visits=visits, batch_size=batch, superko=superko, evaluator=tb.EVAL_DUALNET_TC,
              dedup=dedup, seed=1)
e.load_state_dict(random_init_state_dict(size, 0))
e.reset(never_resign=np.ones(games, np.uint8))
ms = moves = 0
for i in range(steps + 1):
    r = e.genmove(mode=mode, visits=visits, strict=strict, play=play, full=False)
    assert (r["error"] == 0).all()
    if i:
        ms += e.last_device_ms; moves += int((r["move"] >= 0).sum())
print(f"{name}{' dedup' if dedup else ''} games={games} superko={superko} play={play} batch={batch}: {ms / steps:.2f} ms/step, {moves / (ms * 1e-3):.1f} moves/s, evals/step {int(r['evals'][1])}, "
      f"eval kernel {e.bench_kernel('eval_ms'):.2f} ms of the last step, launches {e.launches}")
