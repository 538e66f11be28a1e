"""selfplay_worker with the reference's signature (selfplay/worker.py:21-90), running a pool of games per GPU.

The reference plays its index list one game after another in one OS process; here every index is a slot of the
device pool and all slots advance one move per engine step, finished games are replaced by the next unplayed index,
and each finished game is written as `<save_dir>/<index>.sgf` in the reference's format.

Host/device overlap: the per-move records (move, colour, root actions, improved policy) stay on the device in the
engine's record ring; a step is queued with tg_genmove_async, and while the GPU searches step i+1 the host formats and
writes the SGF files of the games that ended in step i (C++ writer threads behind tg_write_records).  No Python loop
over games runs per step.
"""
import os
import random
import sys
import time

import numpy as np

from ..engine import Engine, MODE_SH, EVAL_DUALNET_TC
from ..nn.network import load_network


def selfplay_worker(save_dir, model_file_path, index_list, size, visits, use_gpu, pool_size=4096, device_index=0,
                    dedup=True, seed=None, network=None, evaluator=EVAL_DUALNET_TC, zobrist=None, never_resign_fn=None,
                    max_steps=None, stats=None, scoring=0):
    """Returns the number of root moves played.  Extensions over the reference's six arguments are keyword-only in
    spirit: pool_size (games resident on the GPU), dedup (evaluate identical leaves of a phase once: same results),
    seed (None = a fresh 64-bit seed per call, like the reference's unseeded `random`, worker.py:39), max_steps (stop
    after that many engine steps; unfinished games are not written, exactly like a killed reference worker, whose
    resume rule worker.py:47-48 this function also follows), stats (dict filled with timing counters)."""
    if not use_gpu:
        raise RuntimeError("tamago_b200 has no CPU path: use_gpu must be True")
    todo = [i for i in index_list if not os.path.isfile(os.path.join(save_dir, f"{i}.sgf"))]     # worker.py:47-48
    if not todo:
        return 0
    net = network if network is not None else load_network(model_file_path, True, board_size=size, device_index=device_index)
    games = min(pool_size, len(todo))
    if seed is None:
        # worker.py:39 seeds numpy from an UNSEEDED `random`: every run plays different games.  The device noise stream
        # is keyed by (seed, game index, move, node), so a fixed default seed would replay identical games whenever a
        # pipeline iteration reuses indices 1..N with unchanged weights.
        seed = int.from_bytes(os.urandom(8), "little")
        print(f"selfplay_worker: noise seed {seed}", file=sys.stderr)
    rng = random.Random(seed)
    nr = never_resign_fn or (lambda index: rng.randint(1, 10) == 1)                               # worker.py:53
    todo = np.array(todo, np.int64)
    nr_all = np.array([nr(int(i)) for i in todo], np.uint8)
    eng = Engine(board_size=size, games=games, max_visits=visits, komi=7.0, superko=True, device=device_index,
                 evaluator=evaluator, dedup=dedup, seed=seed & 0xFFFFFFFFFFFFFFFF, record_ring=True, scoring=scoring)
    if evaluator == EVAL_DUALNET_TC or getattr(net, "state_dict_np", None) is not None:
        eng.load_state_dict(net.state_dict_np)
    if zobrist is not None:
        eng.set_zobrist(zobrist)
    slot_index = todo[:games].copy()
    slot_nr = nr_all[:games].copy()
    next_todo = games
    active = np.ones(games, bool)
    eng.reset(game_ids=slot_index.astype(np.uint64), never_resign=slot_nr)
    moves_played = files = file_moves = steps = 0
    t0 = time.perf_counter()
    eng.genmove_async(mode=MODE_SH, visits=visits, play=True, full=False)
    while True:
        r = eng.collect()
        if (r["error"][active] != 0).any():
            raise RuntimeError("device search reported an error (node pool / history overflow)")
        moves_played += int(((r["move"] >= 0) & active).sum())                                    # worker.py:65-72
        fin_slots = np.flatnonzero(active & (r["finished"] != 0))                                 # worker.py:76-90
        fin_index = slot_index[fin_slots].copy()
        if len(fin_slots):
            eng.fetch_records(fin_slots)                     # copies are ordered before the reset below recycles the slots
            nre = min(len(fin_slots), len(todo) - next_todo)
            re_slots = fin_slots[:nre]
            slot_index[re_slots] = todo[next_todo:next_todo + nre]
            slot_nr[re_slots] = nr_all[next_todo:next_todo + nre]
            next_todo += nre
            active[fin_slots[nre:]] = False
            if nre:
                mask = np.zeros(games, np.uint8)
                mask[re_slots] = 1
                eng.reset(mask=mask, game_ids=slot_index.astype(np.uint64), never_resign=slot_nr)
        steps += 1
        more = bool(active.any()) and (max_steps is None or steps < max_steps)
        if more:
            eng.genmove_async(mode=MODE_SH, visits=visits, play=True, full=False)
        if len(fin_slots):                                   # the GPU is already searching the next step
            file_moves += eng.write_records(save_dir, fin_index)
            files += len(fin_slots)
        if not more:
            break
    if stats is not None:
        stats.update(seconds=time.perf_counter() - t0, steps=steps, files=files, file_moves=file_moves, moves=moves_played,
                     seed=seed, launches=eng.launches)
    eng.close()
    return moves_played
