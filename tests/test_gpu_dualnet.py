"""T2 on the B200: DualNet forward (tcgen05 tensor-core kernel and the CUDA-core fp32 kernel) vs the reference.

Tolerance 1e-4 on policy logits, softmax policy and softmax value (BASELINE.json north_star), against outputs of
the reference's torch DualNet on CPU for the same seeded weights and real positions (tests/golden/dualnet_<N>.npz).
"""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _weights(size, seed):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_golden import numpy_weights
    return numpy_weights(size, seed)


@pytest.mark.parametrize("size", [9, 13, 19])
@pytest.mark.parametrize("evaluator", ["tc", "fp32"])
def test_dualnet_matches_reference_golden(golden_dir, size, evaluator):
    import tamago_b200 as tb
    g = np.load(os.path.join(golden_dir, f"dualnet_{size}.npz"))
    ev = tb.EVAL_DUALNET_TC if evaluator == "tc" else tb.EVAL_DUALNET_FP32
    e = tb.Engine(board_size=size, games=32, max_visits=8, evaluator=ev)
    e.load_state_dict(_weights(size, int(g["weight_seed"])))
    logits, val = e.forward(g["planes"], use_logit=True)
    d_logit = np.abs(logits - g["logits"]).max()
    d_val = np.abs(val - g["value_softmax"]).max()
    pol, _ = e.forward(g["planes"], use_logit=False)
    d_pol = np.abs(pol - g["policy_softmax"]).max()
    print(f"size {size} {evaluator}: max |dlogit| {d_logit:.3e}, |dvalue| {d_val:.3e}, |dpolicy| {d_pol:.3e}")
    assert d_logit <= TOL and d_val <= TOL and d_pol <= TOL
    e.close()


def _torch_dualnet_f64(sd, x, size):
    """fp64 restatement of DualNet.forward (dual_net.py:41-52, res_block.py:27-39, head/*.py) with torch ops."""
    import torch
    import torch.nn.functional as F
    dev = "cuda"
    t = {k: torch.from_numpy(np.asarray(v)).to(dev).double() for k, v in sd.items() if not k.endswith("num_batches_tracked")}

    def bn(h, p, eps):
        return F.batch_norm(h, t[p + ".running_mean"], t[p + ".running_var"], t[p + ".weight"], t[p + ".bias"], False, 0.0, eps)
    h = torch.from_numpy(x).to(dev).double()
    h = F.relu(bn(F.conv2d(h, t["conv_layer.weight"], padding=1), "bn_layer", 1e-5))
    b = 0
    while f"blocks.{b}.conv1.weight" in t:
        h1 = F.relu(bn(F.conv2d(h, t[f"blocks.{b}.conv1.weight"], padding=1), f"blocks.{b}.bn1", 2e-5))
        h2 = bn(F.conv2d(h1, t[f"blocks.{b}.conv2.weight"], padding=1), f"blocks.{b}.bn2", 2e-5)
        h = F.relu(h + h2)
        b += 1
    p = F.relu(bn(F.conv2d(h, t["policy_head.conv_layer.weight"]), "policy_head.bn_layer", 2e-5)).flatten(1)
    logits = p @ t["policy_head.fc_layer.weight"].T + t["policy_head.fc_layer.bias"]
    v = F.relu(bn(F.conv2d(h, t["value_head.conv_layer.weight"]), "value_head.bn_layer", 2e-5)).flatten(1)
    vlog = v @ t["value_head.fc_layer.weight"].T + t["value_head.fc_layer.bias"]
    return logits.cpu().numpy(), torch.softmax(vlog, 1).cpu().numpy()


@pytest.mark.parametrize("size", [9, 19])
def test_kernels_match_fp64_reference_on_large_ragged_batch(golden_dir, size):
    """Ragged batch sizes (group remainders, several waves of CTAs): both kernels against an fp64 evaluation of the
    same network, 1e-4 on logits and values.  The policy FC of the seeded weights is scaled so that logits stay in the
    range of trained networks (|logit| < ~10); the error of any fp32-accumulating implementation scales with it."""
    import tamago_b200 as tb
    g = np.load(os.path.join(golden_dir, f"dualnet_{size}.npz"))
    rs = np.random.RandomState(3)
    n = 1003 if size == 9 else 301
    base = g["planes"]
    x = base[rs.randint(0, len(base), n)].copy()
    # perturb stones so that positions differ (keep one-hot structure of planes 0..2)
    flip = rs.rand(n, size, size) < 0.15
    cls = rs.randint(0, 3, (n, size, size))
    for c in range(3):
        x[:, c][flip] = (cls[flip] == c).astype(np.float32)
    sd = _weights(size, 4242)
    sd["policy_head.fc_layer.weight"] = (sd["policy_head.fc_layer.weight"] * 0.3).astype(np.float32)
    ref_logits, ref_val = _torch_dualnet_f64(sd, x, size)
    print(f"size {size}: |logit| max {np.abs(ref_logits).max():.2f}")
    for name, ev in (("tc", tb.EVAL_DUALNET_TC), ("fp32", tb.EVAL_DUALNET_FP32)):
        e = tb.Engine(board_size=size, games=64, max_visits=32, evaluator=ev)
        e.load_state_dict(sd)
        logits, val = e.forward(x, use_logit=True)
        e.close()
        dl, dv = np.abs(logits - ref_logits).max(), np.abs(val - ref_val).max()
        print(f"size {size} {name}: max |dlogit| {dl:.3e}, |dvalue| {dv:.3e} vs fp64")
        assert np.isfinite(logits).all() and dl <= TOL and dv <= TOL


def test_13x13_tensor_core_kernel_matches_fp64():
    """13x13 (two boards per CTA group): no reference golden exists for this size, so the check is the fp64 torch evaluation."""
    import tamago_b200 as tb
    size, n = 13, 157
    rs = np.random.RandomState(13)
    cls = rs.randint(0, 3, (n, size, size))
    x = np.zeros((n, 6, size, size), np.float32)
    for c in range(3):
        x[:, c] = (cls == c)
    x[:, 5] = np.where(rs.rand(n) < 0.5, 1.0, -1.0)[:, None, None]
    sd = _weights(size, 77)
    sd["policy_head.fc_layer.weight"] = (sd["policy_head.fc_layer.weight"] * 0.3).astype(np.float32)
    ref_logits, ref_val = _torch_dualnet_f64(sd, x, size)
    e = tb.Engine(board_size=size, games=8, max_visits=32, evaluator=tb.EVAL_DUALNET_TC)
    e.load_state_dict(sd)
    logits, val = e.forward(x, use_logit=True)
    e.close()
    assert np.abs(logits - ref_logits).max() <= TOL and np.abs(val - ref_val).max() <= TOL


@pytest.mark.parametrize("size", [9, 13, 19])
def test_tensor_core_kernel_is_batch_invariant(golden_dir, size):
    """A board's policy / value must not depend on its slot, its group, its tile or the batch size: the same planes are
    evaluated in batches of 1, G-1, G, G+1, one full wave + 1 and in reversed order, and every result must be
    BIT-identical (the kernel has no cross-board reduction and a fixed summation order)."""
    import tamago_b200 as tb
    g = np.load(os.path.join(golden_dir, f"dualnet_{9 if size == 13 else size}.npz"))
    rs = np.random.RandomState(11)
    G = {9: 5, 13: 2, 19: 1}[size]
    n = 148 * G + 1
    if size == 13:
        x = np.zeros((n, 6, 13, 13), np.float32)
        stones = rs.randint(0, 3, (n, 13, 13))
        for c in range(3):
            x[:, c] = (stones == c)
        x[:, 5] = np.where(rs.rand(n) < 0.5, 1.0, -1.0)[:, None, None]
    else:
        x = g["planes"][rs.randint(0, len(g["planes"]), n)].copy()
    e = tb.Engine(board_size=size, games=64, max_visits=64, evaluator=tb.EVAL_DUALNET_TC)
    e.load_state_dict(_weights(size, 99))
    full_p, full_v = e.forward(x, use_logit=True)
    rev_p, rev_v = e.forward(x[::-1].copy(), use_logit=True)
    assert np.array_equal(rev_p[::-1], full_p) and np.array_equal(rev_v[::-1], full_v)
    for m in sorted({1, max(1, G - 1), G, G + 1, 2 * G + 1}):
        p, v = e.forward(x[:m], use_logit=True)
        assert np.array_equal(p, full_p[:m]) and np.array_equal(v, full_v[:m]), m
        p, v = e.forward(x[n - m:], use_logit=False)              # softmax output path, boards in other slots
        sp, sv = e.forward(x, use_logit=False)
        assert np.array_equal(p, sp[n - m:]) and np.array_equal(v, sv[n - m:]), m
    e.close()


def test_device_resident_forward_aliases_engine_buffers(golden_dir):
    """tg_eval_buffers / tg_forward_device / tg_sync: torch tensors alias the engine's device batch (zero copy) and give the
    same bits as the host-buffer call."""
    import torch
    import tamago_b200 as tb
    g = np.load(os.path.join(golden_dir, "dualnet_9.npz"))
    x = g["planes"][:37]
    e = tb.Engine(board_size=9, games=16, max_visits=32, evaluator=tb.EVAL_DUALNET_TC)
    e.load_state_dict(_weights(9, 5))
    ref_p, ref_v = e.forward(x, use_logit=True)
    planes, policy, value = e.eval_tensors()
    assert planes.is_cuda and planes.shape[1:] == (6, 9, 9) and policy.shape[1] == 82 and value.shape[1] == 3
    policy.zero_(); value.zero_()
    planes[:len(x)].copy_(torch.from_numpy(x).cuda())
    torch.cuda.synchronize()
    e.forward_device(len(x), use_logit=True)
    e.sync()
    assert np.array_equal(policy[:len(x)].cpu().numpy(), ref_p) and np.array_equal(value[:len(x)].cpu().numpy(), ref_v)
    with pytest.raises(Exception):
        e.forward_device(planes.shape[0] + 1)
    e.close()


def test_model_bin_of_the_reference_through_load_network(golden_dir):
    """north_star: "the model.bin checkpoint loader stays drop-in".  model_ref_9.bin was written by the reference's
    save_model (nn/utility.py:80-87); load_network (139-159) -> inference / inference_with_policy_logits on the device
    must give the outputs the reference's own loaded network produced (1e-4), and a missing file keeps a random init."""
    import torch
    from tamago_b200.nn.network import load_network
    g = np.load(os.path.join(golden_dir, "model_ref_9.npz"))
    net = load_network(os.path.join(golden_dir, "model_ref_9.bin"), True, board_size=9)
    x = torch.from_numpy(g["planes"])
    logits, val = net.inference_with_policy_logits(x)
    pol, val2 = net.inference(x)
    dl = np.abs(logits.numpy() - g["logits"]).max()
    print(f"model.bin: max |dlogit| {dl:.3e}")
    assert dl <= TOL and np.abs(val.numpy() - g["value_softmax"]).max() <= TOL
    assert np.abs(pol.numpy() - g["policy_softmax"]).max() <= TOL and np.array_equal(val.numpy(), val2.numpy())
    rnd = load_network("/nonexistent/model.bin", True, board_size=9)       # utility.py:152-155: failure is swallowed
    assert np.abs(rnd.inference_with_policy_logits(x)[0].numpy() - g["logits"]).max() > 1e-2


def test_activation_overflow_is_reported_not_clamped_silently(golden_dir):
    """activations above the fp16 operand range (6e4) were clamped silently.  Now the kernel still clamps (no
    inf/NaN can reach the tensor cores) but every host-facing entry point fails with a message; the fp32 evaluator of the
    same engine build handles the same weights."""
    import tamago_b200 as tb
    g = np.load(os.path.join(golden_dir, "dualnet_9.npz"))
    sd = _weights(9, 4242)
    sd["conv_layer.weight"] = (sd["conv_layer.weight"] * 3.0e5).astype(np.float32)       # stem activations ~1e5
    e = tb.Engine(board_size=9, games=4, max_visits=8, evaluator=tb.EVAL_DUALNET_TC)
    e.load_state_dict(sd)
    with pytest.raises(Exception, match="overflow"):
        e.forward(g["planes"][:4], use_logit=True)
    e.load_state_dict(_weights(9, 4242))                                                  # the flag does not stick
    logits, _ = e.forward(g["planes"][:4], use_logit=True)
    assert np.isfinite(logits).all()
    e.close()
    e = tb.Engine(board_size=9, games=4, max_visits=8, evaluator=tb.EVAL_DUALNET_FP32)
    e.load_state_dict(sd)
    logits, _ = e.forward(g["planes"][:4], use_logit=True)
    assert np.isfinite(logits).all()
    e.close()


@pytest.mark.parametrize("size", [9, 19])
def test_unscaled_seeded_weights_error_scales_with_logit_magnitude(golden_dir, size):
    """The 1e-4 contract is stated for logits of the magnitude trained networks produce (|logit| <~ 16; the golden nets reach
    18).  With the UNSCALED seeded policy FC the logits are several times larger; any fp32-accumulating implementation's
    error grows with them.  Pinned here: error <= 1e-4 * max(1, max|logit| / 16) against an fp64 evaluation."""
    import tamago_b200 as tb
    g = np.load(os.path.join(golden_dir, f"dualnet_{size}.npz"))
    sd = _weights(size, 4242)
    x = g["planes"][:24]
    ref_logits, ref_val = _torch_dualnet_f64(sd, x, size)
    bound = 1e-4 * max(1.0, np.abs(ref_logits).max() / 16.0)
    for ev in (tb.EVAL_DUALNET_TC, tb.EVAL_DUALNET_FP32):
        e = tb.Engine(board_size=size, games=8, max_visits=8, evaluator=ev)
        e.load_state_dict(sd)
        logits, val = e.forward(x, use_logit=True)
        e.close()
        dl = np.abs(logits - ref_logits).max()
        print(f"size {size} evaluator {ev}: max|logit| {np.abs(ref_logits).max():.1f}, max |dlogit| {dl:.3e} (bound {bound:.3e})")
        assert dl <= bound and np.abs(val - ref_val).max() <= 1e-4


@pytest.mark.parametrize("size", [9, 19])
def test_device_weight_loader_matches_host_loader(golden_dir, size):
    """tg_load_weights_device (BatchNorm fold, power-of-two layer scale, fp16 hi/lo split and UMMA tile packing done by two
    kernels from torch CUDA tensors) produces the same operand bits as the host loader: outputs of both evaluators are
    BIT-identical for the same state_dict."""
    import torch
    import tamago_b200 as tb
    g = np.load(os.path.join(golden_dir, f"dualnet_{size}.npz"))
    sd = _weights(size, 777)
    x = g["planes"][:24]
    for ev in (tb.EVAL_DUALNET_TC, tb.EVAL_DUALNET_FP32):
        e = tb.Engine(board_size=size, games=8, max_visits=8, evaluator=ev)
        e.load_state_dict(sd)
        want = e.forward(x, use_logit=True)
        e.load_state_dict_device({k: torch.from_numpy(np.asarray(v)).cuda() for k, v in sd.items() if not k.endswith("num_batches_tracked")})
        got = e.forward(x, use_logit=True)
        e.close()
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]), ev
