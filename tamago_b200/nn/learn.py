"""Gumbel-AlphaZero training step on the self-play GPUs.

Drop-in for the reference's trainer entry point and files:
  train_with_gumbel_alphazero_on_gpu(program_dir, board_size, batch_size)           nn/learn.py:318-403
  losses: KL-divergence policy loss (batchmean) + cross-entropy value loss          nn/loss.py:33-55
  optimiser: SGD lr 0.01, momentum 0.9, nesterov, weight decay 1e-4                 learning_param.py, learn.py:335-339
  files: <program_dir>/data/rl_data_*.npz in, model/rl-model.bin + model/rl-state.ckpt out (reference layouts)

What is different is where it runs: one process per GPU (torch.distributed, NCCL over NVLink), the global mini-batch of
`batch_size` positions split evenly over the ranks.  Two things keep an N-rank step equal to the reference's one-process
step on the same batch: BatchNorm statistics are reduced over all ranks before they are used (the per-channel sums travel
through a differentiable all-reduce), and the gradients are summed over ranks in ONE flat all-reduce per step (461 298
parameters = 1.85 MB at 9x9).  With one rank nothing is reduced and the arithmetic is the reference's.
The forward/backward arithmetic is stock torch ("stock torch is adequate" for training); the hand-written
engine stays the self-play hot path and reloads model/rl-model.bin for the next iteration.
"""
import glob
import os
import time

import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch import nn

RL_LEARNING_RATE, MOMENTUM, WEIGHT_DECAY, RL_VALUE_WEIGHT = 0.01, 0.9, 1e-4, 1.0      # learning_param.py:5-34


def _world():
    return (dist.get_rank(), dist.get_world_size()) if dist.is_available() and dist.is_initialized() else (0, 1)


class _AllReduceSum(torch.autograd.Function):
    """y = sum over ranks of x; dL/dx = sum over ranks of dL/dy (each rank's loss term depends on every rank's x)."""
    @staticmethod
    def forward(ctx, x):
        y = x.clone()
        dist.all_reduce(y, op=dist.ReduceOp.SUM)
        return y

    @staticmethod
    def backward(ctx, g):
        g = g.clone()
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
        return g


class GlobalBatchNorm2d(nn.BatchNorm2d):
    """nn.BatchNorm2d whose training statistics cover the GLOBAL mini-batch (all ranks), so that splitting a batch over
    GPUs does not change the result.  Same parameters / buffers / state_dict keys as nn.BatchNorm2d; running statistics
    follow torch's update rule (momentum m: r <- (1 - m) r + m s, unbiased variance for running_var)."""

    def forward(self, x):
        rank, world = _world()
        if not self.training or world == 1:
            return super().forward(x)
        n_local = x.numel() // x.shape[1]
        n = float(n_local * world)
        # two passes like torch's own kernel (mean first, then the centred second moment: no E[x^2] - mean^2 cancellation)
        mean = _AllReduceSum.apply(x.sum(dim=(0, 2, 3))) / n
        xc = x - mean[None, :, None, None]
        var = _AllReduceSum.apply((xc * xc).sum(dim=(0, 2, 3))) / n   # biased variance of the global batch
        with torch.no_grad():
            m = self.momentum
            self.running_mean.mul_(1 - m).add_(mean.detach(), alpha=m)
            self.running_var.mul_(1 - m).add_(var.detach() * (n / (n - 1.0)), alpha=m)
            self.num_batches_tracked += 1
        xh = xc * torch.rsqrt(var[None, :, None, None] + self.eps)
        return xh * self.weight[None, :, None, None] + self.bias[None, :, None, None]


class _Block(nn.Module):                                           # nn/network/res_block.py:8-39
    def __init__(self, c):
        super().__init__()
        self.conv1 = nn.Conv2d(c, c, 3, padding=1, bias=False)
        self.conv2 = nn.Conv2d(c, c, 3, padding=1, bias=False)
        self.bn1 = GlobalBatchNorm2d(c, eps=2e-5, momentum=0.01)
        self.bn2 = GlobalBatchNorm2d(c, eps=2e-5, momentum=0.01)

    def forward(self, x):
        h = F.relu(self.bn1(self.conv1(x)))
        return F.relu(x + self.bn2(self.conv2(h)))


class _Head(nn.Module):                                            # nn/network/head/policy_head.py:7-40, value_head.py:7-40
    def __init__(self, n, c, planes, outputs):
        super().__init__()
        self.conv_layer = nn.Conv2d(c, planes, 1, bias=False)
        self.bn_layer = GlobalBatchNorm2d(planes, eps=2e-5, momentum=0.01)
        self.fc_layer = nn.Linear(planes * n * n, outputs)

    def forward(self, x):
        return self.fc_layer(F.relu(self.bn_layer(self.conv_layer(x))).flatten(1))


class TrainDualNet(nn.Module):
    """DualNet for training with the reference's state_dict layout (nn/network/dual_net.py:14-52; 94 tensors),
    so that model/rl-model.bin files are interchangeable with the reference and with the engine's loader."""

    def __init__(self, board_size=9, filters=64, blocks=6):
        super().__init__()
        self.conv_layer = nn.Conv2d(6, filters, 3, padding=1, bias=False)
        self.bn_layer = GlobalBatchNorm2d(filters)                 # default eps 1e-5, momentum 0.1 (dual_net.py:32)
        self.blocks = nn.Sequential(*[_Block(filters) for _ in range(blocks)])
        self.policy_head = _Head(board_size, filters, 2, board_size ** 2 + 1)
        self.value_head = _Head(board_size, filters, 1, 3)

    def forward(self, x):
        h = self.blocks(F.relu(self.bn_layer(self.conv_layer(x))))
        return self.policy_head(h), self.value_head(h)


def rl_losses(policy_logits, value_logits, policy_target, value_target, global_batch):
    """nn/loss.py:33-55 + learn.py:367-371 for this rank's slice of the global batch.  Returns (loss term of this rank,
    policy-loss term, value-loss term); the sum of each over the ranks is the reference's number:
        policy_loss = KLDivLoss(batchmean)(log_softmax(out), target) = sum_b sum_a t (log t - log p) / B
        value_loss  = CrossEntropy(reduction none) per sample;  loss = (policy_loss + 1.0 * value_loss).mean()."""
    logp = F.log_softmax(policy_logits, -1)
    kl = torch.where(policy_target > 0, policy_target * (policy_target.log() - logp), torch.zeros_like(logp)).sum()
    policy_term = kl / global_batch
    value_term = F.cross_entropy(value_logits, value_target, reduction="sum") / global_batch
    return policy_term + RL_VALUE_WEIGHT * value_term, policy_term, value_term


def load_data_set(path):
    """nn/utility.py:90-103 (same np.random call: one permutation per file)"""
    data = np.load(path)
    perm = np.random.permutation(len(data["value"]))
    return data["input"][perm], data["policy"][perm].astype(np.float32), data["value"][perm].astype(np.int64)


class GradientAllReduce:
    """One flat SUM all-reduce of all gradients per step (NCCL on GPUs, gloo in the CPU tests)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        self.numel = sum(p.numel() for p in self.params)
        self.flat = None
        self.bytes_per_step = self.numel * 4

    def __call__(self):
        rank, world = _world()
        if world == 1:
            return
        if self.flat is None:
            self.flat = torch.empty(self.numel, dtype=torch.float32, device=self.params[0].device)
        o = 0
        for p in self.params:
            n = p.numel()
            self.flat[o:o + n].copy_(p.grad.reshape(-1)); o += n
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        o = 0
        for p in self.params:
            n = p.numel()
            p.grad.copy_(self.flat[o:o + n].view_as(p.grad)); o += n


def train_steps(net, optimizer, plane_data, policy_data, value_data, batch_size, device, amp=False, scaler=None, max_steps=None, log=None):
    """The inner loop of learn.py:361-384 over one data set: global batches of `batch_size`, this rank's slice of each."""
    rank, world = _world()
    assert batch_size % world == 0, "the global batch must split evenly over the ranks"
    b_local = batch_size // world
    reducer = GradientAllReduce(net.parameters())
    sums = {"loss": 0.0, "policy": 0.0, "value": 0.0}
    iteration = 0
    net.train()
    for i in range(0, len(value_data) - batch_size + 1, batch_size):
        lo = i + rank * b_local
        plane = torch.from_numpy(np.ascontiguousarray(plane_data[lo:lo + b_local])).to(device)
        policy = torch.from_numpy(np.ascontiguousarray(policy_data[lo:lo + b_local])).to(device)
        value = torch.from_numpy(np.ascontiguousarray(value_data[lo:lo + b_local])).to(device)
        with torch.autocast(device_type=device.type, enabled=amp):
            pp, vp = net(plane)
            loss, pl, vl = rl_losses(pp.float(), vp.float(), policy, value, batch_size)
        optimizer.zero_grad(set_to_none=False)
        if scaler is not None:
            scaler.scale(loss).backward()
            scaler.unscale_(optimizer)
            reducer()
            scaler.step(optimizer)
            scaler.update()
        else:
            loss.backward()
            reducer()
            optimizer.step()
        stats = torch.stack([loss.detach(), pl.detach(), vl.detach()])
        if world > 1:
            dist.all_reduce(stats, op=dist.ReduceOp.SUM)
        stats = stats.tolist()
        sums["loss"] += stats[0]; sums["policy"] += stats[1]; sums["value"] += stats[2]
        if log is not None:
            log.append(stats)
        iteration += 1
        if max_steps is not None and iteration >= max_steps:
            break
    return sums, iteration, reducer.bytes_per_step


def train_with_gumbel_alphazero_on_gpu(program_dir, board_size, batch_size, device=None, amp=True, max_steps=None, log=None):
    """nn/learn.py:318-403.  device: torch.device of this rank (default cuda:LOCAL_RANK); amp=True is the reference's
    autocast + GradScaler path, amp=False trains in fp32 (bit-comparable with the reference's CPU trainer, learn.py:234-315)."""
    rank, world = _world()
    if device is None:
        device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    data_set = sorted(glob.glob(os.path.join(program_dir, "data", "rl_data_*.npz")))
    net = TrainDualNet(board_size).to(device)
    optimizer = torch.optim.SGD(net.parameters(), lr=RL_LEARNING_RATE, momentum=MOMENTUM, weight_decay=WEIGHT_DECAY, nesterov=True)
    use_scaler = amp and device.type == "cuda"
    scaler = torch.amp.GradScaler(device.type) if use_scaler else None
    num_trained_batches = 0
    model_file_path = os.path.join(program_dir, "model", "rl-model.bin")
    if os.path.exists(model_file_path):
        if rank == 0:
            print(f"load {model_file_path}")
        net.load_state_dict(torch.load(model_file_path, map_location=device))
    if world > 1:                                                   # every rank trains the same network: rank 0's initialisation
        for t in list(net.parameters()) + list(net.buffers()):
            dist.broadcast(t.data, 0)
    state_file_path = os.path.join(program_dir, "model", "rl-state.ckpt")
    if os.path.exists(state_file_path):
        if rank == 0:
            print(f"load {state_file_path}")
        ck = torch.load(state_file_path, map_location=device)
        optimizer.load_state_dict(ck["optimizer_state_dict"])
        if scaler is not None and "scaler_state_dict" in ck:
            scaler.load_state_dict(ck["scaler_state_dict"])
        num_trained_batches = ck["num_trained_batches"]
        for group in optimizer.param_groups:
            group["lr"] = RL_LEARNING_RATE
    allreduce_bytes = 0
    for data_index, path in enumerate(data_set):
        plane_data, policy_data, value_data = load_data_set(path)
        t0 = time.time()
        sums, iteration, allreduce_bytes = train_steps(net, optimizer, plane_data, policy_data, value_data, batch_size, device,
                                                       amp=amp and device.type == "cuda", scaler=scaler, max_steps=max_steps, log=log)
        num_trained_batches += iteration
        if rank == 0 and iteration:
            print(f"epoch 0, data-{data_index} : loss = {sums['loss'] / iteration:6f}, time = {time.time() - t0:3f} seconds.\n"
                  f"\tpolicy loss : {sums['policy'] / iteration:6f}\n\tvalue loss  : {sums['value'] / iteration:6f}")
    if rank == 0:                                                   # nn/utility.py:80-87 save_model + learn.py:398-403
        os.makedirs(os.path.join(program_dir, "model"), exist_ok=True)
        torch.save({k: v.detach().to("cpu") for k, v in net.state_dict().items()}, model_file_path)
        state = {"num_trained_batches": num_trained_batches, "optimizer_state_dict": optimizer.state_dict()}
        if scaler is not None:
            state["scaler_state_dict"] = scaler.state_dict()
        torch.save(state, state_file_path)
    if world > 1:
        dist.barrier()
    return {"num_trained_batches": num_trained_batches, "allreduce_bytes_per_step": allreduce_bytes if world > 1 else 0, "world": world, "net": net}


def train_with_gumbel_alphazero_on_cpu(program_dir, board_size, batch_size, **kw):
    """The product has no CPU path; the name exists so that train.py's dispatch fails loudly instead of silently."""
    raise RuntimeError("tamago_b200 has no CPU training path: use train_with_gumbel_alphazero_on_gpu")
