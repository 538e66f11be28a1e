"""T2 pin of the oracle's DualNet restatement against the reference's own DualNet outputs (goldens)."""
import os
import sys

import numpy as np
import pytest


@pytest.mark.parametrize("size", [9, 13, 19])
def test_dualnet_ref_matches_reference(golden_dir, size):
    sys.path.insert(0, os.path.join(golden_dir))
    from make_golden import numpy_weights
    from oracle.dualnet_ref import DualNetRef
    import torch
    torch.set_num_threads(1)
    g = np.load(os.path.join(golden_dir, f"dualnet_{size}.npz"))
    net = DualNetRef(numpy_weights(size, int(g["weight_seed"])), size)
    logits, vlog = net.forward(g["planes"])
    assert np.abs(logits.numpy() - g["logits"]).max() <= 2e-5
    assert np.abs(vlog.numpy() - g["value_logits"]).max() <= 2e-5
    pol, val = net.evaluator()(g["planes"], False)
    assert np.abs(pol - g["policy_softmax"]).max() <= 1e-6 and np.abs(val - g["value_softmax"]).max() <= 1e-6


def test_random_init_has_reference_layout():
    from tamago_b200.nn.utility import random_init_state_dict, state_dict_names
    sd = random_init_state_dict(9, 0)
    assert set(sd) == set(state_dict_names())
    assert sd["policy_head.fc_layer.weight"].shape == (82, 162) and sd["blocks.5.conv2.weight"].shape == (64, 64, 3, 3)


def test_model_bin_written_by_the_reference(golden_dir):
    """tests/golden/model_ref_9.bin was written by the reference's save_model (nn/utility.py:80-87) and the golden
    outputs come from the reference's load_network (139-159) on that file.  The host loader must read the file (94 tensors,
    counters dropped) and the oracle network built from it must reproduce the reference's outputs."""
    import torch
    from oracle.dualnet_ref import DualNetRef
    from tamago_b200.nn.utility import load_state_dict_file, state_dict_names
    torch.set_num_threads(1)
    g = np.load(os.path.join(golden_dir, "model_ref_9.npz"))
    sd = load_state_dict_file(os.path.join(golden_dir, "model_ref_9.bin"))
    assert len(g["names"]) == 94 and set(sd) == set(state_dict_names()) == {str(k) for k in g["names"] if not str(k).endswith("num_batches_tracked")}
    assert sd["bn_layer.running_var"].std() > 0.1            # non-trivial BatchNorm statistics (folding is exercised)
    net = DualNetRef(sd, 9)
    logits, _ = net.forward(g["planes"])
    assert np.abs(logits.numpy() - g["logits"]).max() <= 2e-5
    pol, val = net.evaluator()(g["planes"], False)
    assert np.abs(pol - g["policy_softmax"]).max() <= 1e-6 and np.abs(val - g["value_softmax"]).max() <= 1e-6
