"""Multi-GPU plumbing on real hardware (run under torchrun on 2+ GPUs; tests/test_gpu_multi.py drives it):
one RL iteration -- sharded self-play, NCCL gather of the device-resident training samples, data-parallel training step
with one flat NCCL all-reduce per step, reload of the written model.bin -- with the cross-rank invariants checked."""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    import tamago_b200 as tb
    from tamago_b200.pipeline import run_iteration
    from tamago_b200.selfplay.shard import gather_sample_tensors
    from tamago_b200.selfplay.worker import SelfPlayPool
    from tamago_b200.nn.utility import random_init_state_dict
    out = {"world": world}
    # (1) gather of device-resident samples over NCCL: every rank must receive every rank's rows, in rank order
    tmp = tempfile.mkdtemp(prefix=f"tg_multi_{rank}_")
    np.random.seed(100 + rank)
    games = 24 + 8 * rank                                               # ragged counts across ranks
    pool = SelfPlayPool(tmp, 9, 16, games, iter(range(rank * 1000 + 1, rank * 1000 + 1 + games)), state_dict=random_init_state_dict(9, 0),
                        device_index=local, seed=7 + rank, sample_cap=8 * games)
    pool.start()
    while pool.active.any():
        pool.step()
    mine = pool.eng.sample_tensors()
    t0 = time.perf_counter()
    (inp, pol, val), counts = gather_sample_tensors(mine)
    torch.cuda.synchronize()
    out["gather_ms"] = (time.perf_counter() - t0) * 1e3
    assert counts == [8 * (24 + 8 * r) for r in range(world)], counts
    off = sum(counts[:rank])
    assert torch.equal(inp[off:off + counts[rank]], mine[0]) and torch.equal(pol[off:off + counts[rank]], mine[1]) and torch.equal(val[off:off + counts[rank]], mine[2])
    chk = torch.stack([inp.double().sum(), pol.sum(), val.double().sum()])
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi), "ranks disagree on the gathered arrays"
    out["gathered_samples"] = int(sum(counts)); out["gathered_bytes"] = int(sum(counts)) * (6 * 81 * 4 + 82 * 8 + 4)
    pool.close()
    # (2) one whole iteration: self-play shards -> gather -> npz -> data-parallel training -> model.bin
    prog = [tempfile.mkdtemp(prefix="tg_multi_prog_") if rank == 0 else None]
    dist.broadcast_object_list(prog, 0)
    r = run_iteration(prog[0], size=9, visits=16, num_data=64 * world, batch_size=64, pool_size=64, seed=11, amp=False, device_index=local)
    assert r["num_trained_batches"] >= 1 and r["allreduce_bytes_per_step"] == 461298 * 4
    sd = torch.load(os.path.join(prog[0], "model", "rl-model.bin"))
    assert len(sd) == 94
    # every rank holds the same weights after the all-reduced steps
    w = torch.cat([p.detach().reshape(-1).double() for p in r.pop("net").parameters()]).sum()
    lo, hi = w.clone(), w.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert lo.item() == hi.item(), "ranks diverged"
    out["iteration"] = {k: v for k, v in r.items() if k != "net"}
    # (3) the next iteration's engine reads the file the trainer wrote
    from tamago_b200.nn.network import load_network
    net = load_network(os.path.join(prog[0], "model", "rl-model.bin"), True, board_size=9, device_index=local)
    e = tb.Engine(board_size=9, games=4, max_visits=16, device=local)
    e.load_state_dict(net.state_dict_np)
    rr = e.genmove(mode=tb.MODE_SH, visits=16, play=True)
    assert (rr["error"] == 0).all()
    e.close()
    dist.barrier()
    if rank == 0:
        print("MULTI_GPU_CHECK " + json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
