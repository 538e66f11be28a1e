import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import tamago_b200 as tb
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
size, ng, visits, seed = 19, 8, 400, 41
e = tb.Engine(board_size=size, games=ng, max_visits=visits, superko=True, batch_size=batch, evaluator=tb.EVAL_HASHNET2, seed=seed)
e.reset(game_ids=np.arange(ng))
for step in range(2 * size * size):
    r = e.genmove(mode=tb.MODE_PUCT, visits=visits, strict=False, play=True)
    print(step, r["move"].tolist(), r["error"].tolist(), flush=True)
    if (r["finished"] != 0).all():
        break
