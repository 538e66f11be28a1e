"""One iteration of the reference's RL loop (pipeline.sh:1-5: selfplay_main.py -> [get_final_status.py] -> train.py --rl true)
on the GPUs of one node, one process per GPU (torchrun):

  1. self-play: every rank plays its contiguous slice of the game indices (selfplay_main.py:44-47) on its own engine; SGF
     files go to <program_dir>/archive/<k>/ like the reference's, training samples go straight from the device record ring
     to device tensors (tg_emit_samples);
  2. data: the samples of all ranks are collected with one NCCL all_gather per array (the reference's workers meet in a
     shared directory instead) and rank 0 writes <program_dir>/data/rl_data_0.npz in the reference's layout;
     (data="sgf" instead re-reads the SGF files of the last `window_size` games like train.py:43-57 does);
  3. training: train_with_gumbel_alphazero_on_gpu, data parallel over the same ranks, writes model/rl-model.bin;
  4. the next iteration's engines load that file (load_network), closing the loop without leaving the box.

    torchrun --nproc-per-node N -m tamago_b200.pipeline --program-dir DIR --size 9 --visits 16 --num-data 10000 --iterations 3
"""
import argparse
import glob
import json
import os
import time

import numpy as np


def _dist():
    import torch.distributed as dist
    return (dist.get_rank(), dist.get_world_size()) if dist.is_available() and dist.is_initialized() else (0, 1)


def run_iteration(program_dir, size=9, visits=16, num_data=10000, batch_size=256, pool_size=4096, data="direct", window_size=300000,
                  dedup=True, scoring=0, seed=None, amp=True, max_train_steps=None, device_index=None, net_on_device=None):
    """net_on_device: the TrainDualNet the previous iteration returned (out["net"]).  Its parameters are handed to the
    self-play engine on the device (tg_load_weights_device: fold + packing on the GPU, no host round trip); without it the
    engine reads model/rl-model.bin like the reference's workers do."""
    import torch
    import torch.distributed as dist
    from .selfplay.shard import shard_for_rank, gather_sample_tensors
    from .selfplay.worker import SelfPlayPool
    from .nn.network import load_network
    from .nn.learn import train_with_gumbel_alphazero_on_gpu
    from .nn.data_generator import generate_reinforcement_learning_data
    rank, world = _dist()
    dev = int(os.environ.get("LOCAL_RANK", "0")) if device_index is None else device_index
    archive = os.path.join(program_dir, "archive")
    if rank == 0:                                                       # selfplay_main.py:48-54: next archive/<k>
        os.makedirs(archive, exist_ok=True)
        ks = [int(os.path.basename(d)) for d in glob.glob(os.path.join(archive, "*")) if os.path.basename(d).isdigit()] + [0]
        k = max(ks) + 1
        os.makedirs(os.path.join(archive, str(k)))
        for sub in ("data", "model"):
            os.makedirs(os.path.join(program_dir, sub), exist_ok=True)
    if world > 1:
        kk = torch.tensor([k if rank == 0 else 0], device=f"cuda:{dev}")
        dist.broadcast(kk, 0)
        k = int(kk.item())
    save_dir = os.path.join(archive, str(k))
    out = {"iteration_dir": save_dir, "world": world}
    # 1. self-play
    t0 = time.perf_counter()
    mine = shard_for_rank(num_data, rank, world)
    if seed is None:
        seed = int.from_bytes(os.urandom(8), "little")
    sd_host = None
    if net_on_device is None:
        sd_host = load_network(os.path.join(program_dir, "model", "rl-model.bin"), True, board_size=size, device_index=dev).state_dict_np
    pool = SelfPlayPool(save_dir, size, visits, max(1, min(pool_size, len(mine))), mine, state_dict=sd_host, device_index=dev,
                        dedup=dedup, seed=seed + rank, scoring=scoring, sample_cap=8 * len(mine) if data == "direct" else 0)
    if net_on_device is not None:
        pool.eng.load_state_dict_device(net_on_device.state_dict())
    out["weights_from"] = "device" if net_on_device is not None else "model.bin"
    pool.start()
    while pool.active.any():
        pool.step()
    moves = pool.moves_played
    torch.cuda.synchronize(dev)
    out["selfplay_seconds"] = time.perf_counter() - t0
    # 2. data
    t1 = time.perf_counter()
    if data == "direct":
        (inp, pol, val), counts = gather_sample_tensors(pool.eng.sample_tensors())
        if rank == 0:
            from .csrc_round import round_policy_like_sgf
            for f in glob.glob(os.path.join(program_dir, "data", "rl_data_*.npz")):
                os.remove(f)
            n = (len(val) // batch_size) * batch_size                    # data_generator.py:143-148: whole batches only
            np.savez_compressed(os.path.join(program_dir, "data", "rl_data_0"), input=inp[:n].cpu().numpy(),
                                policy=round_policy_like_sgf(pol[:n].cpu().numpy()), value=val[:n].cpu().numpy().astype(np.int32),
                                kifu_count=np.array(num_data))
        out["samples"] = int(sum(counts)); out["gathered_bytes"] = int(sum(counts)) * (6 * size * size * 4 + (size * size + 1) * 8 + 4)
    pool.close()
    if data != "direct" and rank == 0:
        dirs, n = [], 0
        for idx in sorted((int(os.path.basename(d)) for d in glob.glob(os.path.join(archive, "*")) if os.path.basename(d).isdigit()), reverse=True):
            d = os.path.join(archive, str(idx)); dirs.append(d); n += len(glob.glob(os.path.join(d, "*.sgf")))
            if n >= window_size:
                break
        for f in glob.glob(os.path.join(program_dir, "data", "rl_data_*.npz")):
            os.remove(f)
        generate_reinforcement_learning_data(program_dir, dirs, size, device=dev)
    if world > 1:
        dist.barrier()
    out["data_seconds"] = time.perf_counter() - t1
    # 3. training
    t2 = time.perf_counter()
    np.random.seed(seed % (1 << 31))                                    # the same data permutation on every rank
    tr = train_with_gumbel_alphazero_on_gpu(program_dir, size, batch_size, device=torch.device("cuda", dev), amp=amp, max_steps=max_train_steps)
    torch.cuda.synchronize(dev)
    out["train_seconds"] = time.perf_counter() - t2
    out.update(moves=moves, games=len(mine), num_trained_batches=tr["num_trained_batches"], allreduce_bytes_per_step=tr["allreduce_bytes_per_step"],
               net=tr["net"])
    return out


def main():
    import torch
    import torch.distributed as dist
    ap = argparse.ArgumentParser()
    ap.add_argument("--program-dir", required=True)
    ap.add_argument("--size", type=int, default=9)
    ap.add_argument("--visits", type=int, default=16)                  # learning_param.py:40 SELF_PLAY_VISITS
    ap.add_argument("--num-data", type=int, default=10000)             # learning_param.py:46 NUM_SELF_PLAY_GAMES
    ap.add_argument("--iterations", type=int, default=1)
    ap.add_argument("--pool-size", type=int, default=4096)
    ap.add_argument("--data", default="direct", choices=["direct", "sgf"])
    ap.add_argument("--scoring", type=int, default=0)
    a = ap.parse_args()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    net = None
    for it in range(a.iterations):
        r = run_iteration(a.program_dir, a.size, a.visits, a.num_data, pool_size=a.pool_size, data=a.data, scoring=a.scoring, net_on_device=net)
        net = r["net"]
        if _dist()[0] == 0:
            print(json.dumps({k: v for k, v in r.items() if k != "net"}))
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
