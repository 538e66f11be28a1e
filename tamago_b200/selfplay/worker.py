"""selfplay_worker with the reference's signature (selfplay/worker.py:21-90), running a pool of games per GPU.

The reference plays its index list one game after another in one OS process; here every index is a slot of the
device pool and all slots advance one move per engine step (tg_genmove with play=1), finished games are replaced by
the next unplayed index, and each finished game is written as `<save_dir>/<index>.sgf` in the reference's format.
"""
import os
import random

import numpy as np

from ..engine import Engine, MODE_SH, EVAL_DUALNET_TC
from ..nn.network import load_network
from ..sgf.selfplay_record import SelfPlayRecord


def selfplay_worker(save_dir, model_file_path, index_list, size, visits, use_gpu, pool_size=4096, device_index=0,
                    dedup=True, seed=None, network=None, evaluator=EVAL_DUALNET_TC, zobrist=None, never_resign_fn=None):
    if not use_gpu:
        raise RuntimeError("tamago_b200 has no CPU path: use_gpu must be True")
    todo = [i for i in index_list if not os.path.isfile(os.path.join(save_dir, f"{i}.sgf"))]     # worker.py:47-48
    if not todo:
        return 0
    net = network if network is not None else load_network(model_file_path, True, board_size=size, device_index=device_index)
    games = min(pool_size, len(todo))
    rng = random.Random(seed if seed is not None else random.choice(index_list))                  # worker.py:39
    nr = never_resign_fn or (lambda index: rng.randint(1, 10) == 1)                                  # worker.py:53
    eng = Engine(board_size=size, games=games, max_visits=visits, komi=7.0, superko=True, device=device_index,
                 evaluator=evaluator, dedup=dedup, seed=seed or 0)
    if evaluator == EVAL_DUALNET_TC or getattr(net, "state_dict_np", None) is not None:
        eng.load_state_dict(net.state_dict_np)
    if zobrist is not None:
        eng.set_zobrist(zobrist)
    queue = list(todo)
    slot_index = [queue.pop(0) for _ in range(games)]
    active = np.ones(games, bool)
    eng.reset(game_ids=np.array(slot_index, np.uint64), never_resign=np.array([nr(i) for i in slot_index], np.uint8))
    records = [SelfPlayRecord(save_dir, size) for _ in range(games)]
    moves_played = 0
    while active.any():
        r = eng.genmove(mode=MODE_SH, visits=visits, play=True, full=True)
        if (r["error"][active] != 0).any():
            raise RuntimeError("device search reported an error (node pool / history overflow)")
        reset_mask = np.zeros(games, np.uint8)
        ids = np.array(slot_index, np.uint64)
        nrf = np.zeros(games, np.uint8)
        for g in np.flatnonzero(active):
            mv = int(r["move"][g])
            if mv >= 0:                                                                           # worker.py:65-72
                k = int(r["num_children"][g])
                records[g].save_record(mv, int(r["color"][g]), k, r["action"][g, :k], r["improved"][g, :k])
                moves_played += 1
            if r["finished"][g]:                                                                  # worker.py:76-90
                records[g].write_record(slot_index[g], int(r["winner"][g]), bool(r["resigned"][g]), float(r["score"][g]))
                records[g].clear()
                if queue:
                    slot_index[g] = queue.pop(0)
                    ids[g] = slot_index[g]; nrf[g] = nr(slot_index[g]); reset_mask[g] = 1
                else:
                    active[g] = False
        if reset_mask.any():
            eng.reset(mask=reset_mask, game_ids=ids, never_resign=nrf)
    eng.close()
    return moves_played
