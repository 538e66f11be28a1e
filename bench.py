#!/usr/bin/env python
"""bench.py -- self-play moves/sec of the B200 engine (BASELINE.json metric) and of the CPU reference.

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's OWN code from baseline/_ref on the host cores

A "step" is one move of every game in the pool (configs[1]: 9x9, 4096 games/GPU, 400-visit Gumbel sequential
halving, random-init DualNet, komi 7, super-ko on).  The timed loop is the shipped self-play pipeline
(tamago_b200.selfplay.worker.SelfPlayPool, the body of selfplay_worker): asynchronous engine steps, finished games'
records fetched from the device ring, formatted and written as SGF files while the next step runs, finished slots
refilled.  The pool is pre-aged (untimed, 2-visit moves) to a uniform spread of game ages so that games end -- and files
are written -- at a steady rate inside the timed region.  `value` = moves / device time (CUDA events around each step's
kernel sequence), `e2e` = moves / wall time of the same steps (host<->device copies, record files, resets included).
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_EVAL = {9: 72281646, 13: 150798686, 19: 322548446}      # SURVEY.md B.4 (2 x MAC); 13x13 by the same count
METRIC = "self-play moves/sec @400 visits (9x9 & 19x19), 1/2/4/8 B200 vs CPU ref"
DRAM_BYTES_PER_EVAL = {9: 480.2}         # k_dualnet_tc, ncu --set full capture (profiles/r02_dualnet_tc.md): 196.69 MB / 409 600 evaluations


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows if len(r) > 2 + i)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


# ----------------------------------------------------------------------------------------------------------
# CPU arms
# ----------------------------------------------------------------------------------------------------------
def host_cores():
    cores = os.cpu_count() or 1
    try:                                    # one torch-importing process per core: stay inside the host's free memory
        import psutil
        cores = max(1, min(cores, int(psutil.virtual_memory().available // (1 << 30))))
    except Exception:
        pass
    return cores


def reference_dir(size):
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    try:
        import install_ref
        return install_ref.ref_dir(size)
    finally:
        sys.path.pop(0)


def reference_run(mode, size, visits, procs, warm_s, steps, step_s, batch=1):
    """The reference's own code (baseline/_ref, unmodified) on `procs` host processes, launched ONCE for the whole run:
    every worker (baseline/ref_worker.py) plays for warm_s + steps * step_s seconds and reports the wall-clock stamp of
    every root move; the moves are then counted per step window.  Returns (moves/s, per-step moves, seconds per step)."""
    ref = reference_dir(size)
    if ref is None:
        return None
    total = warm_s + steps * step_s
    env = dict(os.environ, OMP_NUM_THREADS="1", MKL_NUM_THREADS="1", PYTHONDONTWRITEBYTECODE="1")
    t_launch = time.time()
    ps = [subprocess.Popen([sys.executable, os.path.join(ROOT, "baseline", "ref_worker.py"), "--ref", ref, "--mode", mode, "--size", str(size),
                            "--visits", str(visits), "--seconds", f"{total:.1f}", "--index", str(i + 1), "--batch", str(batch), "--stamps", "1"],
                           stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, env=env) for i in range(procs)]
    stamps, starts = [], []
    for p in ps:
        out, _ = p.communicate(timeout=total + 180)
        try:
            d = json.loads(out.decode().strip().splitlines()[-1])
        except Exception:
            continue
        stamps += d.get("stamps", []); starts.append(d["start"])
    if not starts:
        return None
    # workers start their clocks after the imports; the window opens warm_s after the LAST worker started
    t0 = max(starts) + warm_s
    st = np.array(sorted(stamps))
    t_end = min(t0 + steps * step_s, min(starts) + total - 0.2)      # every worker is still playing until then
    eff_step = (t_end - t0) / steps
    per_step = [int(((st >= t0 + i * eff_step) & (st < t0 + (i + 1) * eff_step)).sum()) for i in range(steps)]
    moves = sum(per_step)
    return moves / max(1e-9, steps * eff_step), per_step, eff_step, time.time() - t_launch


def _cpu_port_worker(args):
    size, visits, moves, seed, game = args
    import torch
    torch.set_num_threads(1)
    from oracle import oracle as orc
    from oracle.dualnet_ref import DualNetRef
    from tamago_b200.nn.utility import random_init_state_dict
    net = DualNetRef(random_init_state_dict(size, 0), size)
    zob = orc.default_zobrist(size)
    b = orc.OracleBoard(size, 7.0, True, zob)
    t = orc.OracleTree(size, net.evaluator(), tree_size=4 * visits + 16, batch_size=1)
    color, played, passes = orc.BLACK, 0, 0
    t0 = time.perf_counter()
    for _ in range(moves):
        t.set_noise_key(seed, game, b.moves)
        pos = t.genmove_sh(b, color, visits, True)
        b.put_stone(pos, color)
        color = 3 - color
        played += 1
        passes = passes + 1 if pos == 0 else 0
        if passes == 2 or b.moves >= 2 * size * size:
            b = orc.OracleBoard(size, 7.0, True, zob)
            color, passes = orc.BLACK, 0
    return played, time.perf_counter() - t0, t.evals


def cpu_port_run(size, visits, procs, moves_per_proc, seed=0):
    """moves/sec of the oracle port (C search + torch fp32 DualNet) on `procs` host threads."""
    import multiprocessing as mp
    from oracle import oracle as orc
    orc.build()
    jobs = [(size, visits, moves_per_proc, seed, g) for g in range(procs)]
    t0 = time.perf_counter()
    if procs == 1:
        res = [_cpu_port_worker(jobs[0])]
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_cpu_port_worker, jobs)
    wall = time.perf_counter() - t0
    moves = sum(r[0] for r in res)
    return moves / wall, wall, moves


def workload_config(a):
    named = a.size == 9 and a.games == 4096 and a.visits == 400
    return {"workload": f"{a.size}x{a.size}, {a.games} parallel games/GPU, {a.visits}-visit Gumbel sequential halving"
                        + (" (BASELINE.json configs[1])" if named else ""),
            "board_size": a.size, "games_per_gpu": a.games, "visits": a.visits, "search": "gumbel-sequential-halving",
            "net": "DualNet 6x64 random init", "komi": 7.0, "superko": True, "dedup_identical_leaves": bool(a.dedup),
            "pipeline": "selfplay_worker pool: async steps, device record ring, SGF files written for finished games",
            "l2": "working set per step (leaf planes + node pool, > 1 GB) exceeds the 126 MB L2; no flush needed"}


def run_reference(a, out=sys.stdout):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    # bounded sample: the whole run (start-up + warm + K windows) stays within ~3 minutes whatever K the driver asks for
    step_s = a.ref_step_seconds if a.ref_step_seconds > 0 else float(min(20.0, max(4.0, 150.0 / max(1, a.steps))))
    warm_s = 8.0 + 4.0 * min(a.warmup, 2)
    res = reference_run("sh", a.size, a.visits, cores, warm_s, a.steps, step_s) if not a.ref_port else None
    if res is not None:
        value, per_step, eff_step, wall = res
        kind = "reference"
        sample = (f"{cores} processes of the unmodified reference (baseline/_ref, selfplay.worker.selfplay_worker, use_gpu False, random-init net, "
                  f"1 torch thread each) playing from the empty board; root moves counted in {a.steps} windows of {eff_step:.1f} s after {warm_s:.0f} s of warm-up")
        ms_per_step = eff_step * 1e3
        extra = {"moves_per_step_window": per_step, "run_seconds": wall}
    else:
        mpp = max(1, a.ref_moves)
        vals, walls = [], []
        cpu_port_run(a.size, a.visits, cores, 1)
        for _ in range(a.steps):
            v, w, m = cpu_port_run(a.size, a.visits, cores, mpp)
            vals.append(v); walls.append(w)
        value, kind, ms_per_step = float(np.mean(vals)), "port", float(np.mean(walls)) * 1e3
        sample = f"{cores} processes x {mpp} moves from the empty board per step, oracle C search + torch fp32 DualNet (1 thread each); baseline/_ref absent"
        extra = {}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "moves/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(a),
        "cpu_baseline": {"value": value, "unit": "moves/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "moves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    line.update(extra)
    print(json.dumps(line), file=out, flush=True)


# ----------------------------------------------------------------------------------------------------------
def run_ours(a, out=sys.stdout):
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line (no NCCL version banner)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import tamago_b200 as tb
    from tamago_b200.nn.utility import random_init_state_dict
    from tamago_b200.selfplay.worker import SelfPlayPool
    from tamago_b200.selfplay.shard import shard_offset

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(vals_max, vals_sum):
        """max over ranks of the times, sum over ranks of the counts (one all-reduce each, counters only)"""
        if world == 1:
            return list(vals_max), list(vals_sum)
        mx = torch.tensor(vals_max, dtype=torch.float64, device="cuda"); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = torch.tensor(vals_sum, dtype=torch.float64, device="cuda"); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        return mx.tolist(), sm.tolist()

    peaks = measured_peaks()
    peak = peaks["bf16_tflops_sustained"]
    tmp_root = tempfile.mkdtemp(prefix=f"tamago_bench_r{rank}_")
    W = max(a.warmup, 3)

    def run_pool(size, games, visits, dedup, steps, warm, tag, preage=True):
        """K timed steps of the shipped self-play pipeline on this rank; returns the per-rank counters"""
        save_dir = os.path.join(tmp_root, tag)
        os.makedirs(save_dir, exist_ok=True)
        first = shard_offset(rank) + 1                       # disjoint index ranges per rank (selfplay_main.py:44-47 split)
        pool = SelfPlayPool(save_dir, size, visits, games, iter(range(first, first + 10_000_000)),
                            state_dict=random_init_state_dict(size, 0), device_index=local, dedup=dedup, seed=1234 + rank)
        pool.start()
        if preage:
            pool.preage(max_age=int(1.4 * size * size), seed=rank)
        for i in range(warm):
            pool.step(queue_next=i + 1 < warm)               # nothing is in flight when the clock starts
        barrier()
        l0, f0, fm0 = pool.eng.launches, pool.files, pool.file_moves
        pool.timing = {k: 0.0 for k in pool.timing}
        dev_ms = eval_ms = 0.0
        moves = evals = 0
        t0 = time.perf_counter()
        for i in range(steps):
            # step(): queue a step if none is in flight, wait for it, fetch finished records, reset + refill, queue the NEXT step,
            # write the SGF files while it runs.  The last timed step queues nothing: every step's work lies inside [t0, t1].
            m, nf = pool.step(queue_next=i + 1 < steps)
            moves += m
            dev_ms += pool.eng.last_device_ms
            eval_ms += pool.eng.bench_kernel("eval_ms")
            evals += int(pool.last["evals"][1])
        barrier()
        wall = time.perf_counter() - t0
        res = dict(dev_ms=dev_ms, eval_ms=eval_ms, wall=wall, moves=moves, evals=evals, launches=pool.eng.launches - l0,
                   files=pool.files - f0, file_moves=pool.file_moves - fm0, stride=pool.eng.stride,
                   host_ms_per_step={k: v / steps * 1e3 for k, v in pool.timing.items()})
        pool.close()
        shutil.rmtree(save_dir, ignore_errors=True)
        return res

    def summarise(r, size, steps):
        (dev_ms, wall, eval_ms), (moves, evals, launches, files, file_moves) = reduce(
            [r["dev_ms"], r["wall"], r["eval_ms"]], [r["moves"], r["evals"], r["launches"], r["files"], r["file_moves"]])
        tf = evals * FLOP_PER_EVAL[size] / (eval_ms * 1e-3) / 1e12 / world if eval_ms > 0 else 0.0     # per GPU
        return dict(value=moves / (dev_ms * 1e-3), e2e=moves / wall, ms_per_step=dev_ms / steps, wall_ms_per_step=wall / steps * 1e3,
                    evals_per_step=evals / steps / world, kernel_tflops=tf, roofline_frac=tf / peak, kernel_ms_per_step=eval_ms / steps,
                    kernel_share=eval_ms / dev_ms if dev_ms else None, launches=int(launches), files=int(files), file_moves=int(file_moves),
                    moves=moves)

    n, games, visits = a.size, a.games, a.visits
    sampler = ClockSampler(local)
    sampler.start()
    main_raw = run_pool(n, games, visits, bool(a.dedup), a.steps, W, "main")
    sampler.stop_flag = True
    main = summarise(main_raw, n, a.steps)
    # per-step bytes across PCIe: the game-state block of every slot down (tg_collect), reset staging up, plus the record
    # rows of the games that finished in the step (device ring -> pinned staging)
    rec_bytes = main_raw["file_moves"] * (main_raw["stride"] * 10 + 5) / max(1, a.steps)
    d2h = games * 48 * 4 + rec_bytes
    h2d = games * 10 * (main_raw["files"] > 0)

    extras = {}
    xs = min(a.steps, 10)
    if a.also_dedup and not a.dedup:
        d = summarise(run_pool(n, games, visits, True, xs, 3, "dedup"), n, xs)
        extras["result_preserving_dedup"] = {
            "value": d["value"], "unit": "moves/s", "e2e": d["e2e"], "ms_per_step": d["ms_per_step"], "evals_per_step": d["evals_per_step"],
            "kernel_tflops": d["kernel_tflops"], "files_written": d["files"],
            "note": "same workload and results; leaves of one phase reached by the same path are evaluated once (engine dedup=1, the worker's default)"}
    if a.extras:
        # BASELINE.json configs[2]: 9x9, 16384 games/GPU, 50-visit Gumbel (pipeline.sh RL setting)
        c3 = summarise(run_pool(9, 16384, 50, False, xs, 3, "c3"), 9, xs)
        extras["c3_sh50"] = {"workload": "9x9, 16384 parallel games/GPU, 50-visit Gumbel sequential halving (BASELINE.json configs[2])",
                             "value": c3["value"], "unit": "moves/s", "e2e": c3["e2e"], "ms_per_step": c3["ms_per_step"],
                             "evals_per_step": c3["evals_per_step"], "kernel_tflops": c3["kernel_tflops"], "roofline_frac": c3["roofline_frac"],
                             "files_written": c3["files"]}
        # 19x19 half of the metric, Gumbel SH
        s19 = summarise(run_pool(19, 1024, visits, False, xs, 2, "sh19"), 19, xs)
        extras["board_19x19"] = {"workload": f"19x19, 1024 parallel games/GPU, {visits}-visit Gumbel sequential halving, super-ko on",
                                 "value": s19["value"], "unit": "moves/s", "e2e": s19["e2e"], "ms_per_step": s19["ms_per_step"],
                                 "evals_per_step": s19["evals_per_step"], "kernel_tflops": s19["kernel_tflops"], "roofline_frac": s19["roofline_frac"],
                                 "flop_per_eval": FLOP_PER_EVAL[19], "steps": xs}

        def puct_run(size, games_, visits_, batch, strict, steps, play, tag):
            eng = tb.Engine(board_size=size, games=games_, max_visits=visits_, batch_size=batch, komi=7.0, superko=True, device=local,
                            evaluator=tb.EVAL_DUALNET_TC, seed=77 + rank)
            eng.load_state_dict(random_init_state_dict(size, 0))
            eng.reset(game_ids=np.arange(games_, dtype=np.uint64) + np.uint64(shard_offset(rank)), never_resign=np.ones(games_, np.uint8))
            for _ in range(2):
                eng.genmove(mode=tb.MODE_PUCT, visits=visits_, strict=strict, play=play, full=False)
            barrier()
            dev = ev_ms = 0.0
            moves = evals = 0
            t0 = time.perf_counter()
            for _ in range(steps):
                r = eng.genmove(mode=tb.MODE_PUCT, visits=visits_, strict=strict, play=play, full=False)
                if int((r["error"] != 0).sum()):
                    raise SystemExit(f"bench.py: {tag} reported search errors")
                moves += int((r["move"] >= 0).sum()); dev += eng.last_device_ms; ev_ms += eng.bench_kernel("eval_ms"); evals += int(r["evals"][1])
            wall = time.perf_counter() - t0
            barrier()
            eng.close()
            (dev_m, wall_m, ev_m), (mv, evs) = reduce([dev, wall, ev_ms], [moves, evals])
            return dict(dev_ms=dev_m, wall=wall_m, eval_ms=ev_m, moves=mv, evals=evs)
        # BASELINE.json configs[3]: 19x19, 1024 games/GPU, 400-visit PUCB + super-ko (batch 1 = the reference's NN_BATCH_SIZE,
        # CONSTANT_PLAYOUT: the early stop of is_move_decided is honoured, so fewer than 401 evaluations per move are needed)
        c4 = puct_run(19, 1024, 400, 1, False, xs, True, "c4")
        tf4 = c4["evals"] * FLOP_PER_EVAL[19] / max(1e-9, c4["dev_ms"] * 1e-3) / 1e12 / world
        extras["c4_puct_19x19"] = {"workload": "19x19, 1024 parallel games/GPU, 400-visit PUCT + super-ko, batch 1, CONSTANT_PLAYOUT (not STRICT) "
                                               "(BASELINE.json configs[3])",
                                   "value": c4["moves"] / (c4["dev_ms"] * 1e-3), "unit": "moves/s", "e2e": c4["moves"] / c4["wall"],
                                   "ms_per_step": c4["dev_ms"] / xs, "evals_per_step": c4["evals"] / xs / world, "step_tflops": tf4,
                                   "roofline_frac": tf4 / peak, "dualnet_share_of_step": None,
                                   "note": "roofline_frac = algorithmic DualNet FLOP of the step / whole step time / bf16 sustained peak "
                                           "(401 dependent search iterations per move, each a 1024-position evaluation)"}
        if rank == 0:
            # BASELINE.json configs[4]: 19x19 GTP genmove, 1600 visits, 256-leaf NN batch, one game (engine/analyze path)
            eng = tb.Engine(board_size=19, games=1, max_visits=1600, batch_size=256, komi=7.0, superko=False, device=local,
                            evaluator=tb.EVAL_DUALNET_TC, seed=5)
            eng.load_state_dict(random_init_state_dict(19, 0))
            for _ in range(3):
                eng.genmove(mode=tb.MODE_PUCT, visits=1600, strict=True, play=False, full=False)
            ms, t0 = 0.0, time.perf_counter()
            for _ in range(xs):
                eng.genmove(mode=tb.MODE_PUCT, visits=1600, strict=True, play=False, full=True)
                ms += eng.last_device_ms
            wall = time.perf_counter() - t0
            eng.close()
            tf5 = 1601 * FLOP_PER_EVAL[19] / (ms / xs * 1e-3) / 1e12
            extras["c5_genmove"] = {"workload": "19x19 genmove, one game, 1600-visit PUCT (STRICT), 256-leaf evaluator batches (BASELINE.json configs[4])",
                                    "ms_per_genmove": ms / xs, "e2e_ms_per_genmove": wall / xs * 1e3, "visits_per_s": 1600 / (ms / xs * 1e-3),
                                    "roofline_frac": tf5 / peak, "n_gpus": 1}
        barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        shutil.rmtree(tmp_root, ignore_errors=True)
        return
    flop = FLOP_PER_EVAL[n]
    achieved = main["kernel_tflops"]
    line = {
        "metric": METRIC, "value": main["value"], "unit": "moves/s", "n_gpus": world, "steps": a.steps,
        "warmup": W, "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f16x3 split operands, f32 accumulate (1e-4 parity with the fp32 reference net)",
        "data": "synthetic", "config": workload_config(a),
        "e2e": {"value": main["e2e"], "unit": "moves/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": main["wall_ms_per_step"], "sgf_files_written": main["files"], "moves_in_files": main["file_moves"],
                "api": "tamago_b200.selfplay.worker.SelfPlayPool.step (the loop body of selfplay_worker)",
                "host_ms_per_step": main_raw["host_ms_per_step"]},
        "gpu_launches": main["launches"],
        "clocks": sampler.summary(),
        "roofline": {"bound": "tensor", "kernel": "k_dualnet_tc", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved / peak,
                     "traffic": DRAM_BYTES_PER_EVAL.get(n, 0.0) * (main["evals_per_step"] - games) / 4 if n in DRAM_BYTES_PER_EVAL else None,
                     "traffic_note": "ESTIMATE for one phase launch of this run (a step = 1 root launch of `games` evaluations + 4 phase launches): "
                                     "dram__bytes_read+write per evaluation of the ncu --set full capture in profiles/r02_dualnet_tc.md "
                                     "(196.7 MB for the 409 600-evaluation launch of phase 3 = 480 B/eval; algorithmic 441 B/eval: 97 B leaf snapshot + 4 B slot map in, 340 B policy/value out) x evaluations of the launch; "
                                     "not measured inside this run",
                     "peak_source": f"{peaks['src']} bf16 sustained (kernel timed inside a long step)",
                     "evals_per_step": main["evals_per_step"], "flop_per_eval": flop, "kernel_ms_per_step": main["kernel_ms_per_step"],
                     "kernel_share_of_step": main["kernel_share"],
                     "note": "algorithmic FLOP (72.28 MFLOP/eval at 9x9); the kernel executes 3 fp16 MMAs per product for fp32-grade accuracy"},
    }
    line.update(extras)
    if a.cpu_baseline and world == 1:
        cores = host_cores()
        cb = {}
        res = reference_run("sh", n, visits, cores, 10.0, 1, a.cpu_seconds)
        if res is not None:
            cb = {"value": res[0], "unit": "moves/s", "cores": cores, "kind": "reference",
                  "sample": f"{cores} processes of the unmodified reference (baseline/_ref, selfplay_worker, use_gpu False, 1 torch thread each), "
                            f"root moves counted over {res[2]:.1f} s after 10 s of warm-up, games from the empty board"}
            if a.extras:
                r19 = reference_run("puct", 19, 400, cores, 14.0, 1, a.cpu_seconds + 10.0)
                if r19 is not None:
                    cb["c4_puct_19x19"] = {"value": r19[0], "unit": "moves/s", "cores": cores, "kind": "reference",
                                           "sample": f"{cores} processes looping MCTSTree.search_best_move (400 visits, batch 1, CONSTANT_PLAYOUT) of the "
                                                     f"BOARD_SIZE=19 copy of the reference over {r19[2]:.1f} s"}
        v, w, m = cpu_port_run(n, visits, 1, a.cpu_moves)
        port = {"value": v, "unit": "moves/s", "cores": 1, "kind": "port",
                "sample": f"{m} moves of one {n}x{n} game from the empty board at {visits} visits, oracle C search + torch fp32 DualNet, 1 thread ({w:.1f} s)"}
        if cb:
            cb["port"] = port
        else:
            cb = port
        line["cpu_baseline"] = cb
    print(json.dumps(line), file=out, flush=True)
    if world > 1:
        dist.destroy_process_group()
    shutil.rmtree(tmp_root, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=9)
    ap.add_argument("--games", type=int, default=4096)
    ap.add_argument("--visits", type=int, default=400)
    ap.add_argument("--dedup", type=int, default=0)
    ap.add_argument("--also-dedup", type=int, default=1, help="also report the result-preserving dedup mode as an extra object")
    ap.add_argument("--extras", type=int, default=1, help="also report configs[2] (c3_sh50), 19x19 SH, configs[3] (c4_puct_19x19), configs[4] (c5_genmove)")
    ap.add_argument("--cpu-baseline", type=int, default=1)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="window of the bounded cpu_baseline sample of the reference")
    ap.add_argument("--cpu-moves", type=int, default=12, help="moves of the bounded cpu_baseline sample of the port")
    ap.add_argument("--ref-moves", type=int, default=6, help="moves per process and step of the port fallback in the --impl reference arm")
    ap.add_argument("--ref-step-seconds", type=float, default=0.0, help="--impl reference: seconds per step window (0 = sized so that the run takes ~3 minutes)")
    ap.add_argument("--ref-port", type=int, default=0, help="--impl reference: time the oracle port even when baseline/_ref exists")
    a = ap.parse_args()
    # stdout carries exactly one JSON line: anything libraries write to fd 1 meanwhile (e.g. the NCCL version banner) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    real_stdout = os.fdopen(saved_stdout, "w")
    try:
        if a.impl == "reference":
            run_reference(a, real_stdout)
        else:
            run_ours(a, real_stdout)
    finally:
        real_stdout.flush()


if __name__ == "__main__":
    main()
