import os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
import tamago_b200 as tb
from tamago_b200.nn.utility import random_init_state_dict
e = tb.Engine(board_size=19, games=1024, max_visits=400, batch_size=1, superko=True, evaluator=tb.EVAL_DUALNET_TC, seed=1)
e.load_state_dict(random_init_state_dict(19, 0))
e.reset(never_resign=np.ones(1024, np.uint8))
for step in range(3):
    r = e.genmove(mode=tb.MODE_PUCT, visits=400, strict=False, play=True, full=True)
    v = r["visits"].sum(axis=1)
    print("step", step, "ms", round(e.last_device_ms, 1), "descents per game: min", v.min(), "median", int(np.median(v)), "mean", round(v.mean(), 1), "max", v.max(),
          "games still searching after 210/250/300/350:", int((v > 210).sum()), int((v > 250).sum()), int((v > 300).sum()), int((v > 350).sum()))
