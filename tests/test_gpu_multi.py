"""Multi-GPU paths on real hardware (needs >= 2 GPUs; skipped on a one-GPU box): the NCCL gather of device-resident training
samples, the data-parallel training step (one flat NCCL all-reduce per step) and a whole pipeline iteration.  The N > 1 host
logic is also covered without GPUs by tests/test_shard_gloo.py and tests/test_learn.py (gloo, world_size 2)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_gather_train_pipeline():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    world = 2
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", "29617", os.path.join(ROOT, "scripts", "multi_gpu_check.py")],
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "MULTI_GPU_CHECK" in p.stdout


def test_pipeline_iteration_single_gpu(tmp_path):
    """pipeline.sh on one box: self-play -> training data (device ring) -> training step -> model.bin -> next self-play."""
    import numpy as np
    import torch
    from tamago_b200.pipeline import run_iteration
    net = None
    for it in range(2):
        r = run_iteration(str(tmp_path), size=9, visits=16, num_data=96, batch_size=64, pool_size=48, seed=5 + it, amp=bool(it), net_on_device=net)
        net = r["net"]                                          # iteration 2 takes the weights on the device, no file read
        assert r["weights_from"] == ("device" if it else "model.bin")
        assert r["samples"] == 96 * 8 and r["num_trained_batches"] == 12 * (it + 1) and r["allreduce_bytes_per_step"] == 0
        assert len(os.listdir(r["iteration_dir"])) == 96
    z = np.load(tmp_path / "data" / "rl_data_0.npz")
    assert z["input"].shape == (768, 6, 9, 9) and z["policy"].shape == (768, 82) and z["value"].dtype == np.int32
    sd = torch.load(tmp_path / "model" / "rl-model.bin")
    assert len(sd) == 94 and all(torch.isfinite(v.float()).all() for v in sd.values())
    assert sorted(os.listdir(tmp_path / "archive")) == ["1", "2"]
