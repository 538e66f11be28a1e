"""Small searches that touch every search kernel variant (for compute-sanitizer memcheck / racecheck runs)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import tamago_b200 as tb
rs = np.random.RandomState(0)
for size, games, mode, visits, batch, env in ((9, 3, tb.MODE_SH, 50, 1, {}), (9, 3, tb.MODE_PUCT, 40, 8, {}), (19, 2, tb.MODE_PUCT, 64, 16, {}),
                                              (13, 200, tb.MODE_PUCT, 24, 4, {}), (9, 5, tb.MODE_PUCT, 30, 4, {"TG_PUCT_WARP": "1"})):
    for k in ("TG_PUCT_WARP",):
        os.environ.pop(k, None)
    os.environ.update(env)
    e = tb.Engine(board_size=size, games=games, max_visits=visits, batch_size=batch, superko=True, evaluator=tb.EVAL_HASHNET2, dedup=True,
                  seed=3, record_ring=True, sample_cap=64)
    for step in range(6):
        r = e.genmove(mode=mode, visits=visits, play=True)
        assert (r["error"] == 0).all()
    e.close()
    print("ok", size, games, mode, visits, batch, env, flush=True)
