import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tamago_b200 as tb
from tamago_b200.nn.utility import random_init_state_dict
size = int(sys.argv[1]) if len(sys.argv) > 1 else 9
e = tb.Engine(board_size=size, games=256, max_visits=64, evaluator=tb.EVAL_DUALNET_TC)
e.load_state_dict(random_init_state_dict(size, 0))
e.bench_kernel("dualnet", 7400, 2)
e.bench_kernel("dualnet_dbg", 7400, 1)
