"""Small searches that touch every search kernel variant (for compute-sanitizer memcheck / racecheck runs)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import tamago_b200 as tb
rs = np.random.RandomState(0)
# (size, games, mode, visits, batch, dedup, env): sequential halving; block PUCT with the wavefront walk + deferred expansion (dedup off),
# with the sequential walk (dedup on / TG_PUCT_WAVE=0, two cache slots), inline expansion (200 games), warp kernels (batch 4: full replay,
# batch 1: board snapshots along the previous path)
for size, games, mode, visits, batch, dedup, env in ((9, 3, tb.MODE_SH, 50, 1, True, {}), (9, 3, tb.MODE_PUCT, 40, 8, False, {}),
                                                     (9, 2, tb.MODE_PUCT, 120, 64, False, {}), (19, 2, tb.MODE_PUCT, 64, 16, False, {"TG_WAVE_GT": "64"}),
                                                     (19, 2, tb.MODE_PUCT, 64, 16, True, {}), (9, 3, tb.MODE_PUCT, 40, 8, False, {"TG_PUCT_WAVE": "0", "TG_WALK_SLOTS": "2"}),
                                                     (13, 200, tb.MODE_PUCT, 24, 4, True, {}), (9, 5, tb.MODE_PUCT, 30, 4, True, {"TG_PUCT_WARP": "1"}),
                                                     (19, 5, tb.MODE_PUCT, 60, 1, False, {"TG_PUCT_WARP": "1"})):   # batch 1: board snapshots
    for k in ("TG_PUCT_WARP", "TG_PUCT_WAVE", "TG_WALK_SLOTS", "TG_WAVE_GT", "TG_PUCT_DEFER"):
        os.environ.pop(k, None)
    os.environ.update(env)
    e = tb.Engine(board_size=size, games=games, max_visits=visits, batch_size=batch, superko=True, evaluator=tb.EVAL_HASHNET2, dedup=dedup,
                  seed=3, record_ring=True, sample_cap=64)
    for step in range(6):
        r = e.genmove(mode=mode, visits=visits, play=True)
        assert (r["error"] == 0).all()
    e.close()
    print("ok", size, games, mode, visits, batch, dedup, env, flush=True)
