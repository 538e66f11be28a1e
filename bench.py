#!/usr/bin/env python
"""bench.py -- self-play moves/sec of the B200 engine (BASELINE.json metric) and of the CPU reference port.

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port) on the host cores

A "step" is one move of every game in the pool (configs[1]: 9x9, 4096 games/GPU, 400-visit Gumbel sequential
halving, random-init DualNet, komi 7, super-ko on).  Games that end are restarted in place, so every step plays
`games` moves (minus resignations).  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_EVAL = {9: 72281646, 13: 150798686, 19: 322548446}      # SURVEY.md B.4 (2 x MAC); 13x13 by the same count
METRIC = "self-play moves/sec @400 visits (9x9 & 19x19), 1/2/4/8 B200 vs CPU ref"


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
# CPU arm: the oracle port (C search + torch fp32 DualNet on the CPU), one process per host thread
# ----------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
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
    """moves/sec of the oracle port on `procs` host threads; every process plays `moves_per_proc` moves."""
    from oracle import oracle as orc
    orc.build()
    jobs = [(size, visits, moves_per_proc, seed, g) for g in range(procs)]
    t0 = time.perf_counter()
    if procs == 1:
        res = [_cpu_worker(jobs[0])]
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    moves = sum(r[0] for r in res)
    return moves / wall, wall, moves


# ----------------------------------------------------------------------------------------------------------
def run_reference(a, out=sys.stdout):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    try:                                    # one torch-importing process per core: stay inside the host's free memory
        import psutil
        cores = max(1, min(cores, int(psutil.virtual_memory().available // (1 << 30))))
    except Exception:
        pass
    mpp = max(1, a.ref_moves)
    vals, walls = [], []
    for _ in range(a.warmup if a.warmup < 2 else 1):
        cpu_port_run(a.size, a.visits, cores, 1)
    for _ in range(a.steps):
        v, w, m = cpu_port_run(a.size, a.visits, cores, mpp)
        vals.append(v); walls.append(w)
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "moves/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": float(np.mean(walls)) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(a),
        "cpu_baseline": {"value": value, "unit": "moves/s", "cores": cores, "kind": "port",
                         "sample": f"{cores} processes x {mpp} moves from the empty board per step, oracle C search + torch fp32 DualNet (1 thread each)"},
        "e2e": {"value": value, "unit": "moves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), file=out, flush=True)


def workload_config(a):
    return {"workload": f"{a.size}x{a.size}, {a.games} parallel games/GPU, {a.visits}-visit Gumbel sequential halving (BASELINE.json configs[1])"
            if a.size == 9 and a.games == 4096 and a.visits == 400 else
            f"{a.size}x{a.size}, {a.games} parallel games/GPU, {a.visits}-visit Gumbel sequential halving",
            "board_size": a.size, "games_per_gpu": a.games, "visits": a.visits, "search": "gumbel-sequential-halving",
            "net": "DualNet 6x64 random init", "komi": 7.0, "superko": True, "dedup_identical_leaves": bool(a.dedup),
            "l2": "working set per step (leaf planes + node pool, > 1 GB) exceeds the 126 MB L2; no flush needed"}


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

    n, games, visits = a.size, a.games, a.visits
    eng = tb.Engine(board_size=n, games=games, max_visits=visits, komi=7.0, superko=True, device=local,
                    evaluator=tb.EVAL_DUALNET_TC, dedup=bool(a.dedup), seed=1234 + rank)
    eng.load_state_dict(random_init_state_dict(n, 0))
    # games shard across ranks exactly like selfplay_main.py:44-47 splits its index list: contiguous slices
    next_id = np.uint64(rank * 10_000_000)
    ids = np.arange(games, dtype=np.uint64) + next_id
    next_id += np.uint64(games)
    rs = np.random.RandomState(99 + rank)
    eng.reset(game_ids=ids, never_resign=(rs.rand(games) < 0.1).astype(np.uint8))       # worker.py:53

    def step():
        nonlocal next_id, ids
        r = eng.genmove(mode=tb.MODE_SH, visits=visits, play=True, full=True)
        fin = r["finished"] != 0
        nf = int(fin.sum())
        if nf:                                      # worker.py:46-55: next game of the index list in the same slot
            ids = ids.copy()
            ids[fin] = np.arange(nf, dtype=np.uint64) + next_id
            next_id += np.uint64(nf)
            eng.reset(mask=fin.astype(np.uint8), game_ids=ids, never_resign=(rs.rand(games) < 0.1).astype(np.uint8))
        bad = int((r["error"] != 0).sum())
        if bad:
            raise SystemExit(f"bench.py: {bad} games reported search errors")
        return int((r["move"] >= 0).sum()), nf, r

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(a.warmup, 3)):
        step()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    l0 = eng.launches
    dev_ms = eval_ms = 0.0
    moves = evals = 0
    t0 = time.perf_counter()
    for _ in range(a.steps):
        m, nf, r = step()
        moves += m
        dev_ms += eng.last_device_ms
        eval_ms += eng.bench_kernel("eval_ms")
        evals += int(r["evals"][1])
    barrier()
    wall = time.perf_counter() - t0
    sampler.stop_flag = True
    launches = eng.launches - l0
    # per-step bytes across PCIe: results (game state block, root actions / improved policy / visits) down, reset masks up
    d2h = games * (48 * 4 + eng.stride * (2 + 8 + 4))
    h2d = games * (1 + 8 + 1)

    # second measurement, same workload: identical leaves of a phase evaluated once (same search results, tested)
    extra = None
    if a.also_dedup and not a.dedup:
        eng.close()
        eng = tb.Engine(board_size=n, games=games, max_visits=visits, komi=7.0, superko=True, device=local,
                        evaluator=tb.EVAL_DUALNET_TC, dedup=True, seed=1234 + rank)
        eng.load_state_dict(random_init_state_dict(n, 0))
        eng.reset(game_ids=ids, never_resign=(rs.rand(games) < 0.1).astype(np.uint8))
        for _ in range(3):
            step()
        barrier()
        d_ms = d_ev_ms = 0.0
        d_moves = d_evals = 0
        t1 = time.perf_counter()
        for _ in range(a.steps):
            m, nf, r = step()
            d_moves += m; d_ms += eng.last_device_ms; d_ev_ms += eng.bench_kernel("eval_ms"); d_evals += int(r["evals"][1])
        barrier()
        d_wall = time.perf_counter() - t1
        dstats = torch.tensor([d_ms, d_wall, float(d_moves)], dtype=torch.float64, device="cuda")
        if world > 1:
            dmx = dstats.clone(); dist.all_reduce(dmx, op=dist.ReduceOp.MAX)
            dsm = dstats.clone(); dist.all_reduce(dsm, op=dist.ReduceOp.SUM)
            d_ms_max, d_wall_max, d_moves_all = dmx[0].item(), dmx[1].item(), dsm[2].item()
        else:
            d_ms_max, d_wall_max, d_moves_all = d_ms, d_wall, float(d_moves)
        extra = {"value": d_moves_all / (d_ms_max * 1e-3), "unit": "moves/s", "e2e": d_moves_all / d_wall_max,
                 "ms_per_step": d_ms_max / a.steps, "evals_per_step": d_evals / a.steps,
                 "kernel_tflops": d_evals * FLOP_PER_EVAL[n] / (d_ev_ms * 1e-3) / 1e12 if d_ev_ms > 0 else None,
                 "note": "same workload and results; leaves of one phase reached by the same path are evaluated once (engine dedup=1)"}

    # third measurement: the 19x19 half of the metric (BASELINE.json configs[3] geometry: 1024 games/GPU, 400 visits), as an extra
    extra19 = None
    if a.also_19 and n != 19:
        eng.close()
        g19 = 1024
        eng = tb.Engine(board_size=19, games=g19, max_visits=visits, komi=7.0, superko=True, device=local,
                        evaluator=tb.EVAL_DUALNET_TC, dedup=False, seed=4321 + rank)
        eng.load_state_dict(random_init_state_dict(19, 0))
        eng.reset(game_ids=np.arange(g19, dtype=np.uint64) + np.uint64(rank * 10_000_000), never_resign=np.ones(g19, np.uint8))
        for _ in range(2):
            eng.genmove(mode=tb.MODE_SH, visits=visits, play=True, full=True)
        barrier()
        x_ms = x_ev_ms = 0.0
        x_moves = x_evals = 0
        t2 = time.perf_counter()
        for _ in range(min(a.steps, 4)):
            r = eng.genmove(mode=tb.MODE_SH, visits=visits, play=True, full=True)
            if int((r["error"] != 0).sum()):
                raise SystemExit("bench.py: 19x19 games reported search errors")
            x_moves += int((r["move"] >= 0).sum()); x_ms += eng.last_device_ms
            x_ev_ms += eng.bench_kernel("eval_ms"); x_evals += int(r["evals"][1])
        barrier()
        x_wall = time.perf_counter() - t2
        xs = torch.tensor([x_ms, x_wall, float(x_moves)], dtype=torch.float64, device="cuda")
        if world > 1:
            xmx = xs.clone(); dist.all_reduce(xmx, op=dist.ReduceOp.MAX)
            xsm = xs.clone(); dist.all_reduce(xsm, op=dist.ReduceOp.SUM)
            x_ms_max, x_wall_max, x_moves_all = xmx[0].item(), xmx[1].item(), xsm[2].item()
        else:
            x_ms_max, x_wall_max, x_moves_all = x_ms, x_wall, float(x_moves)
        x_tf = x_evals * FLOP_PER_EVAL[19] / (x_ev_ms * 1e-3) / 1e12 if x_ev_ms > 0 else 0.0
        extra19 = {"workload": f"19x19, {g19} parallel games/GPU, {visits}-visit Gumbel sequential halving, super-ko on",
                   "value": x_moves_all / (x_ms_max * 1e-3), "unit": "moves/s", "e2e": x_moves_all / x_wall_max,
                   "ms_per_step": x_ms_max / min(a.steps, 4), "evals_per_step": x_evals / min(a.steps, 4),
                   "kernel_tflops": x_tf, "roofline_frac": x_tf / measured_peaks()["bf16_tflops_sustained"],
                   "flop_per_eval": FLOP_PER_EVAL[19]}

    stats = torch.tensor([dev_ms, wall, float(moves), eval_ms, float(evals), float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dev_ms_max, wall_max, moves_all = mx[0].item(), mx[1].item(), sm[2].item()
        launches_all = int(sm[5].item())
    else:
        dev_ms_max, wall_max, moves_all, launches_all = dev_ms, wall, float(moves), launches
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = measured_peaks()
    flop = FLOP_PER_EVAL[n]
    achieved = evals * flop / (eval_ms * 1e-3) / 1e12 if eval_ms > 0 else 0.0
    peak = peaks["bf16_tflops_sustained"]
    line = {
        "metric": METRIC, "value": moves_all / (dev_ms_max * 1e-3), "unit": "moves/s", "n_gpus": world, "steps": a.steps,
        "warmup": max(a.warmup, 3), "ms_per_step": dev_ms_max / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f16x3 split operands, f32 accumulate (1e-4 parity with the fp32 reference net)",
        "data": "synthetic", "config": workload_config(a),
        "e2e": {"value": moves_all / wall_max, "unit": "moves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": wall_max / a.steps * 1e3},
        "gpu_launches": launches_all,
        "clocks": sampler.summary(),
        "roofline": {"bound": "tensor", "kernel": "k_dualnet_tc", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved / peak, "traffic": 2455.0 * evals / max(1, a.steps * 5),
                     "traffic_source": "profiles/r01_dualnet_tc.md: 241.3 MB DRAM read+write for a 98.3 k-evaluation launch = 2455 B per evaluation (algorithmic 2284 B: planes in, policy/value out), scaled to this run's mean evaluations per launch (5 launches per step)",
                     "peak_source": f"{peaks['src']} bf16 sustained (kernel timed inside a long step)",
                     "evals_per_step": evals / a.steps, "flop_per_eval": flop, "kernel_ms_per_step": eval_ms / a.steps,
                     "kernel_share_of_step": eval_ms / dev_ms if dev_ms else None,
                     "note": "algorithmic FLOP (72.28 MFLOP/eval at 9x9); the kernel executes 3 fp16 MMAs per product for fp32-grade accuracy"},
    }
    if extra is not None:
        line["result_preserving_dedup"] = extra
    if extra19 is not None:
        line["board_19x19"] = extra19
    if a.cpu_baseline and world == 1:
        v, w, m = cpu_port_run(n, visits, 1, a.cpu_moves)
        line["cpu_baseline"] = {"value": v, "unit": "moves/s", "cores": 1, "kind": "port",
                                "sample": f"{m} moves of one {n}x{n} game from the empty board at {visits} visits, oracle C search + torch fp32 DualNet, 1 thread ({w:.1f} s)"}
    print(json.dumps(line), file=out, flush=True)
    if world > 1:
        dist.destroy_process_group()


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
    ap.add_argument("--also-19", type=int, default=1, help="also report the 19x19 half of the metric (1024 games/GPU) as an extra object")
    ap.add_argument("--cpu-baseline", type=int, default=1)
    ap.add_argument("--cpu-moves", type=int, default=16, help="moves of the bounded cpu_baseline sample")
    ap.add_argument("--ref-moves", type=int, default=6, help="moves per process and step in the --impl reference arm")
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
