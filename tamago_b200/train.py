"""Entry point with the reference's train.py options for the RL half (train.py:14-75): optional data generation from the
SGF archive (last `window-size` games), then the Gumbel-AlphaZero training step on the GPUs.

    python -m tamago_b200.train --kifu-dir archive --size 9 --rl true [--window-size 300000] [--program-dir .]
    torchrun --nproc-per-node N -m tamago_b200.train ...          # data parallel, NCCL all-reduce
Supervised learning (train.py --rl false) is outside the self-play path and not provided.
"""
import argparse
import glob
import os


def train_main(kifu_dir, size, use_gpu, rl, window_size, program_dir=".", batch_size=256):
    import torch
    import torch.distributed as dist
    from .nn.data_generator import generate_reinforcement_learning_data
    from .nn.learn import train_with_gumbel_alphazero_on_gpu
    if not use_gpu:
        raise RuntimeError("tamago_b200 has no CPU path: --use-gpu must be true")
    if not rl:
        raise RuntimeError("only the reinforcement-learning step (--rl true) is part of the self-play path")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl")
    rank = dist.get_rank() if dist.is_initialized() else 0
    if kifu_dir is not None and rank == 0:                              # train.py:43-57
        idx = sorted((int(os.path.split(d)[-1]) for d in glob.glob(os.path.join(kifu_dir, "*")) if os.path.split(d)[-1].isdigit()), reverse=True)
        num, dirs = 0, []
        for i in idx:
            d = os.path.join(kifu_dir, str(i))
            num += len(glob.glob(os.path.join(d, "*.sgf")))
            dirs.append(d)
            if num >= window_size:
                break
        generate_reinforcement_learning_data(program_dir=program_dir, kifu_dir_list=dirs, board_size=size, device=local)
    if dist.is_initialized():
        dist.barrier()
    return train_with_gumbel_alphazero_on_gpu(program_dir=program_dir, board_size=size, batch_size=batch_size, device=torch.device("cuda", local))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kifu-dir", default=None)
    ap.add_argument("--size", type=int, default=9)
    ap.add_argument("--use-gpu", type=lambda s: s.lower() in ("1", "true", "yes"), default=True)
    ap.add_argument("--rl", type=lambda s: s.lower() in ("1", "true", "yes"), default=False)
    ap.add_argument("--window-size", type=int, default=300000)
    ap.add_argument("--program-dir", default=".")
    a = ap.parse_args()
    train_main(a.kifu_dir, a.size, a.use_gpu, a.rl, a.window_size, a.program_dir)


if __name__ == "__main__":
    main()
