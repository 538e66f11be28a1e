"""One CPU worker process of the reference arm: runs the reference's OWN code from baseline/_ref/ for a bounded time.

    python baseline/ref_worker.py --ref <dir> --mode sh|puct --size N --visits V --seconds T --index K [--batch B]

mode sh   : calls the reference's selfplay.worker.selfplay_worker(save_dir, <missing model>, [K], N, V, use_gpu=False)
            unmodified (random-init net: load_network prints "Failed to load" and continues, nn/utility.py:152-155).
            Moves are counted by a pass-through wrapper around SelfPlayRecord.save_record (called once per root move,
            selfplay/worker.py:72); a watchdog ends the process after T seconds (a 400-visit game takes minutes).
mode puct : the same game loop around MCTSTree.search_best_move (mcts/tree.py:57) with CONSTANT_PLAYOUT, i.e. what
            gtp/client.py:215-219 runs per genmove -- the reference has no PUCT self-play entry point.
Prints one JSON line: {"moves": n, "first": t_first_move_done, "last": t_last_move_done, "start": t_start[, "stamps": [...]]}.
One torch thread per worker (OMP_NUM_THREADS=1): the fan-out is one process per core like selfplay_main.py:58.
"""
import argparse
import json
import os
import sys
import tempfile
import threading
import time


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", required=True)
    ap.add_argument("--mode", default="sh")
    ap.add_argument("--size", type=int, default=9)
    ap.add_argument("--visits", type=int, default=400)
    ap.add_argument("--seconds", type=float, default=20.0)
    ap.add_argument("--index", type=int, default=1)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--stamps", type=int, default=0, help="1: also print the wall-clock stamp of every root move")
    a = ap.parse_args()
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    sys.dont_write_bytecode = True
    a.ref = os.path.abspath(a.ref)
    sys.path.insert(0, a.ref)
    os.chdir(a.ref)
    import random
    import numpy as np
    import torch
    torch.set_num_threads(1)
    random.seed(a.index); np.random.seed(a.index); torch.manual_seed(0)
    stamps = []
    t_start = time.time()

    def finish():
        d = {"moves": len(stamps), "first": stamps[0] if stamps else None, "last": stamps[-1] if stamps else None, "start": t_start}
        if a.stamps:
            d["stamps"] = list(stamps)
        line = json.dumps(d) + "\n"
        os.write(1, line.encode())                                  # fd 1 directly: sys.stdout is muted while the reference runs
        os._exit(0)

    threading.Timer(a.seconds, finish).start()
    devnull = open(os.devnull, "w")
    if a.mode == "sh":
        import sgf.selfplay_record as spr
        from selfplay.worker import selfplay_worker
        orig = spr.SelfPlayRecord.save_record

        def counted(self, *args, **kw):
            stamps.append(time.time())
            return orig(self, *args, **kw)
        spr.SelfPlayRecord.save_record = counted
        save_dir = tempfile.mkdtemp(prefix="tamago_ref_")
        sys.stdout, real = devnull, sys.stdout                      # "Failed to load ..." banner
        try:
            selfplay_worker(save_dir, os.path.join(save_dir, "missing-model.bin"), list(range(a.index * 1000, a.index * 1000 + 999)),
                            a.size, a.visits, False)
        finally:
            sys.stdout = real
    else:
        from board.constant import PASS, RESIGN
        from board.go_board import GoBoard
        from board.stone import Stone
        from mcts.tree import MCTSTree
        from mcts.time_manager import TimeManager, TimeControl
        from nn.utility import load_network
        sys.stdout, real = devnull, sys.stdout
        sys.stderr, real_err = devnull, sys.stderr                  # per-move search tables (node.py:254-272)
        net = load_network("missing-model.bin", False)
        tree = MCTSTree(net, tree_size=max(4096, 4 * a.visits), batch_size=a.batch)
        tm = TimeManager(TimeControl.CONSTANT_PLAYOUT, constant_visits=a.visits)
        while True:
            board = GoBoard(board_size=a.size, komi=7.0, check_superko=True)
            color, passes = Stone.BLACK, 0
            for _ in range(2 * a.size * a.size):
                pos = tree.search_best_move(board, color, tm, {})
                if pos == RESIGN:
                    break
                board.put_stone(pos, color)
                stamps.append(time.time())
                passes = passes + 1 if pos == PASS else 0
                color = Stone.get_opponent_color(color)
                if passes == 2:
                    break
    finish()


if __name__ == "__main__":
    main()
