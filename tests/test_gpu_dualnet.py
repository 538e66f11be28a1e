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


@pytest.mark.parametrize("size", [9, 19])
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


@pytest.mark.parametrize("size", [9, 19])
def test_tensor_core_kernel_matches_fp32_kernel_on_large_batch(golden_dir, size):
    """Ragged batch sizes (group remainders, several waves of CTAs): tcgen05 path vs the CUDA-core fp32 path."""
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
    out = {}
    for name, ev in (("tc", tb.EVAL_DUALNET_TC), ("fp32", tb.EVAL_DUALNET_FP32)):
        e = tb.Engine(board_size=size, games=64, max_visits=32, evaluator=ev)
        e.load_state_dict(sd)
        out[name] = e.forward(x, use_logit=True)
        e.close()
    assert np.abs(out["tc"][0] - out["fp32"][0]).max() <= TOL
    assert np.abs(out["tc"][1] - out["fp32"][1]).max() <= TOL
    assert np.isfinite(out["tc"][0]).all()
