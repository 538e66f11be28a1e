"""the Gumbel-AlphaZero training step (tamago_b200/nn/learn.py) against the reference's own trainer.

tests/golden/train_9.npz was produced by running the reference's train_with_gumbel_alphazero_on_cpu (nn/learn.py:234-315)
for four mini-batches from a seeded model and a seeded data set (make_golden.py gen_train); both are regenerated here from
the same seeds, the same number of steps is run, and losses / weights / BatchNorm statistics must agree to 1e-5 (both
sides are fp32 torch on the CPU; the tolerance covers thread-count dependent summation order).  A 2-rank gloo run must
reproduce the 1-rank result (global-batch BatchNorm + one flat gradient all-reduce)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def _setup(tmp_path, g):
    from make_golden import numpy_weights, train_data_set
    size = int(g["size"])
    (tmp_path / "data").mkdir(); (tmp_path / "model").mkdir()
    d = train_data_set(size, int(g["data_seed"]), int(g["samples"]))
    np.savez_compressed(tmp_path / "data" / "rl_data_0.npz", **d)
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in numpy_weights(size, int(g["weight_seed"])).items()}
    torch.save(sd, tmp_path / "model" / "rl-model.bin")
    return size


def _check(g, program_dir, log, tol=1e-5):
    want = g["losses"]
    got = np.array(log)
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=tol, atol=tol)
    sd = torch.load(os.path.join(program_dir, "model", "rl-model.bin"))
    assert [str(k) for k in g["names"]] == list(sd.keys())              # the reference's 94-tensor layout, in its order
    for i, name in enumerate(g["names"]):
        t = sd[str(name)].double().reshape(-1).numpy()
        idx = g["sample_idx"][i][:min(len(t), 64)] % len(t)
        np.testing.assert_allclose(t[idx], g["sample_val"][i][:len(idx)], rtol=10 * tol, atol=tol, err_msg=str(name))
        np.testing.assert_allclose(np.abs(t).sum(), g["abs_sum"][i], rtol=10 * tol, err_msg=str(name))
    ck = torch.load(os.path.join(program_dir, "model", "rl-state.ckpt"))
    assert ck["num_trained_batches"] == int(g["steps"]) and "optimizer_state_dict" in ck


def test_training_steps_match_the_reference_trainer(golden_dir, tmp_path):
    from tamago_b200.nn.learn import train_with_gumbel_alphazero_on_gpu
    torch.set_num_threads(2)
    g = np.load(os.path.join(golden_dir, "train_9.npz"))
    size = _setup(tmp_path, g)
    np.random.seed(int(g["perm_seed"]))
    log = []
    out = train_with_gumbel_alphazero_on_gpu(str(tmp_path), size, int(g["batch"]), device=torch.device("cpu"), amp=False, log=log)
    assert out["num_trained_batches"] == int(g["steps"]) and out["allreduce_bytes_per_step"] == 0
    _check(g, str(tmp_path), log)
    # the engine-side loader reads what the trainer wrote (same file the next self-play iteration loads)
    from tamago_b200.nn.utility import load_state_dict_file, state_dict_names
    assert set(load_state_dict_file(str(tmp_path / "model" / "rl-model.bin"))) == set(state_dict_names())


def _rank_main(rank, world, port, program_dir, size, batch, perm_seed, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from tamago_b200.nn.learn import train_with_gumbel_alphazero_on_gpu
    np.random.seed(perm_seed)
    log = []
    out = train_with_gumbel_alphazero_on_gpu(program_dir, size, batch, device=torch.device("cpu"), amp=False, log=log)
    if rank == 0:
        q.put((log, out["allreduce_bytes_per_step"]))
    dist.destroy_process_group()


def test_two_ranks_reproduce_one_rank(golden_dir, tmp_path):
    """world_size 2 over gloo: each rank trains on half of every global batch; BatchNorm statistics and gradients are
    reduced over the ranks, so losses and final weights equal the reference's single-process numbers."""
    import torch.multiprocessing as mp
    g = np.load(os.path.join(golden_dir, "train_9.npz"))
    size = _setup(tmp_path, g)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_rank_main, args=(r, 2, port, str(tmp_path), size, int(g["batch"]), int(g["perm_seed"]), q)) for r in range(2)]
    for p in ps:
        p.start()
    log, nbytes = q.get(timeout=600)
    for p in ps:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert nbytes == 461298 * 4                                        # one flat all-reduce of every parameter gradient
    _check(g, str(tmp_path), log, tol=2e-5)
