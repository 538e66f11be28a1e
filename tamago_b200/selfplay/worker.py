"""selfplay_worker with the reference's signature (selfplay/worker.py:21-90), running a pool of games per GPU.

The reference plays its index list one game after another in one OS process; here every index is a slot of the
device pool and all slots advance one move per engine step, finished games are replaced by the next unplayed index,
and each finished game is written as `<save_dir>/<index>.sgf` in the reference's format.

Host/device overlap: the per-move records (move, colour, root actions, improved policy) stay on the device in the
engine's record ring; a step is queued with tg_genmove_async, and while the GPU searches step i+1 the host formats and
writes the SGF files of the games that ended in step i (C++ writer threads behind tg_write_records).  No Python loop
over games runs per step.
"""
import itertools
import os
import random
import sys
import time

import numpy as np

from ..engine import Engine, MODE_SH, EVAL_DUALNET_TC
from ..nn.network import load_network


class SelfPlayPool:
    """The game loop of selfplay/worker.py:46-90 for `games` concurrent slots on one GPU.

    indices: iterable of game indices (file names); the pool pulls from it whenever a slot frees up.  step() returns
    after it has queued the NEXT engine step and written the files of the games that ended in the step it collected."""

    def __init__(self, save_dir, size, visits, games, indices, state_dict=None, device_index=0, dedup=True, seed=0,
                 evaluator=EVAL_DUALNET_TC, zobrist=None, never_resign_fn=None, scoring=0, sample_cap=0, write_sgf=True):
        self.save_dir, self.size, self.visits, self.games = save_dir, size, visits, games
        self.indices = iter(indices)
        rng = random.Random(seed)
        self.nr = never_resign_fn or (lambda index: rng.randint(1, 10) == 1)                      # worker.py:53
        self.eng = Engine(board_size=size, games=games, max_visits=visits, komi=7.0, superko=True, device=device_index,
                          evaluator=evaluator, dedup=dedup, seed=seed & 0xFFFFFFFFFFFFFFFF, record_ring=True, scoring=scoring,
                          sample_cap=sample_cap)
        self.sample_cap, self.write_sgf, self.samples = sample_cap, write_sgf, 0
        if state_dict is not None:
            self.eng.load_state_dict(state_dict)
        if zobrist is not None:
            self.eng.set_zobrist(zobrist)
        self.slot_index = np.zeros(games, np.int64)
        self.slot_nr = np.zeros(games, np.uint8)
        self.active = np.zeros(games, bool)
        self.moves_played = self.files = self.file_moves = self.steps = 0
        self.in_flight = False
        self.timing = dict(queue=0.0, wait=0.0, host_between_steps=0.0, write_files=0.0, fetch_records=0.0, emit_samples=0.0,
                           refill_reset=0.0)             # host seconds by phase

    def _refill(self, slots):
        """next unplayed indices into these slots (worker.py:46-55); returns the slots that got a game"""
        got = list(itertools.islice(self.indices, len(slots)))
        re = slots[:len(got)]
        self.slot_index[re] = got
        self.slot_nr[re] = [self.nr(int(i)) for i in got]
        self.active[re] = True
        self.active[slots[len(got):]] = False
        if len(re):
            mask = np.zeros(self.games, np.uint8)
            mask[re] = 1
            self.eng.reset(mask=mask, game_ids=self.slot_index.astype(np.uint64), never_resign=self.slot_nr)
        return re

    def start(self):
        self._refill(np.arange(self.games))
        return bool(self.active.any())

    def preage(self, max_age, seed=0):
        """Untimed set-up for steady-state measurements: brings the slots to a uniform spread of game ages in [0, max_age)
        with cheap 2-visit moves, so that games end (and records are written) at a steady rate afterwards instead of all
        at once ~2 N^2 steps later.  Restarted slots keep their index."""
        age = np.random.RandomState(seed).randint(0, max_age, self.games)
        for s in range(max_age, 0, -1):
            r = self.eng.genmove(mode=MODE_SH, visits=2, play=True, full=False)
            restart = ((age == s - 1) | (r["finished"] != 0)) & self.active
            if restart.any():
                self.eng.reset(mask=restart.astype(np.uint8), game_ids=self.slot_index.astype(np.uint64), never_resign=self.slot_nr)

    def queue(self):
        self.eng.genmove_async(mode=MODE_SH, visits=self.visits, play=True, full=False)
        self.in_flight = True

    def step(self, queue_next=True):
        """Collect the step in flight (queue one first if there is none); returns (moves, games finished)."""
        tm = self.timing
        t0 = time.perf_counter()
        if not self.in_flight:
            self.queue()
        eng, active = self.eng, self.active
        t1 = time.perf_counter()
        r = eng.collect()
        t2 = time.perf_counter()
        self.in_flight = False
        self.last = r
        if (r["error"][active] != 0).any():
            raise RuntimeError("device search reported an error (node pool / history overflow)")
        moves = int(((r["move"] >= 0) & active).sum())                                            # worker.py:65-72
        fin_slots = np.flatnonzero(active & (r["finished"] != 0))                                 # worker.py:76-90
        fin_index = self.slot_index[fin_slots].copy()
        if len(fin_slots):
            ta = time.perf_counter()
            if self.write_sgf:
                eng.fetch_records(fin_slots)                 # copies are ordered before the reset below recycles the slots
            tb_ = time.perf_counter()
            if self.sample_cap:                              # training samples straight from the device ring
                from ..nn.data_generator import emit_from_ring
                self.samples = emit_from_ring(eng, fin_slots, r["n_moves"][fin_slots])
            tc = time.perf_counter()
            self._refill(fin_slots)
            tm["fetch_records"] += tb_ - ta; tm["emit_samples"] += tc - tb_; tm["refill_reset"] += time.perf_counter() - tc
        self.steps += 1
        self.moves_played += moves
        t3 = time.perf_counter()
        if queue_next and active.any():
            self.queue()
        t4 = time.perf_counter()
        if len(fin_slots) and self.write_sgf:                # the GPU is already searching the next step
            self.file_moves += eng.write_records(self.save_dir, fin_index)
            self.files += len(fin_slots)
        t5 = time.perf_counter()
        tm["queue"] += (t1 - t0) + (t4 - t3); tm["wait"] += t2 - t1; tm["host_between_steps"] += t3 - t2; tm["write_files"] += t5 - t4
        return moves, len(fin_slots)

    def close(self):
        if self.in_flight:
            self.eng.collect()
            self.in_flight = False
        self.eng.close()


def selfplay_worker(save_dir, model_file_path, index_list, size, visits, use_gpu, pool_size=4096, device_index=0,
                    dedup=True, seed=None, network=None, evaluator=EVAL_DUALNET_TC, zobrist=None, never_resign_fn=None,
                    max_steps=None, stats=None, scoring=0):
    """Returns the number of root moves played.  Extensions over the reference's six arguments: pool_size (games resident
    on the GPU), dedup (evaluate identical leaves of a phase once: same results), seed (None = a fresh 64-bit seed per
    call, like the reference's unseeded `random`, worker.py:39), max_steps (stop after that many engine steps; unfinished
    games are not written, exactly like a killed reference worker, whose resume rule worker.py:47-48 this function also
    follows), stats (dict filled with timing counters), scoring (tg_config.scoring)."""
    if not use_gpu:
        raise RuntimeError("tamago_b200 has no CPU path: use_gpu must be True")
    todo = [i for i in index_list if not os.path.isfile(os.path.join(save_dir, f"{i}.sgf"))]     # worker.py:47-48
    if not todo:
        return 0
    net = network if network is not None else load_network(model_file_path, True, board_size=size, device_index=device_index)
    if seed is None:
        # worker.py:39 seeds numpy from an UNSEEDED `random`: every run plays different games.  The device noise stream
        # is keyed by (seed, game index, move, node), so a fixed default seed would replay identical games whenever a
        # pipeline iteration reuses indices 1..N with unchanged weights.
        seed = int.from_bytes(os.urandom(8), "little")
        print(f"selfplay_worker: noise seed {seed}", file=sys.stderr)
    sd = net.state_dict_np if (evaluator == EVAL_DUALNET_TC or getattr(net, "state_dict_np", None) is not None) else None
    pool = SelfPlayPool(save_dir, size, visits, min(pool_size, len(todo)), todo, state_dict=sd, device_index=device_index,
                        dedup=dedup, seed=seed, evaluator=evaluator, zobrist=zobrist, never_resign_fn=never_resign_fn, scoring=scoring)
    pool.start()
    t0 = time.perf_counter()
    while pool.active.any() and (max_steps is None or pool.steps < max_steps):
        pool.step(queue_next=max_steps is None or pool.steps + 1 < max_steps)
    if stats is not None:
        stats.update(seconds=time.perf_counter() - t0, steps=pool.steps, files=pool.files, file_moves=pool.file_moves,
                     moves=pool.moves_played, seed=seed, launches=pool.eng.launches)
    moves = pool.moves_played
    pool.close()
    return moves
