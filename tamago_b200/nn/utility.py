"""DualNet parameters on the host: state_dict plumbing of nn/utility.py (load_network, 139-159) and the
layer inventory of nn/network/dual_net.py:14-39 (94 tensors).  No arithmetic happens here:
the forward pass runs in the CUDA engine (tamago_b200.Engine.forward)."""
import numpy as np


def state_dict_names(blocks=6):
    names = ["conv_layer.weight"] + [f"bn_layer.{s}" for s in ("weight", "bias", "running_mean", "running_var")]
    for b in range(blocks):
        for c in (1, 2):
            names.append(f"blocks.{b}.conv{c}.weight")
            names += [f"blocks.{b}.bn{c}.{s}" for s in ("weight", "bias", "running_mean", "running_var")]
    for head in ("policy_head", "value_head"):
        names.append(f"{head}.conv_layer.weight")
        names += [f"{head}.bn_layer.{s}" for s in ("weight", "bias", "running_mean", "running_var")]
        names += [f"{head}.fc_layer.weight", f"{head}.fc_layer.bias"]
    return names


def random_init_state_dict(board_size, seed=0, blocks=6, filters=64):
    """Random initialisation with the distributions torch gives DualNet when no model file loads
    (nn/utility.py:152-155): Conv2d / Linear weights ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (kaiming_uniform, a=sqrt(5)),
    Linear bias ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)), BatchNorm weight 1, bias 0, running_mean 0, running_var 1."""
    rs = np.random.RandomState(seed)
    nn_ = board_size * board_size
    sd = {}

    def uni(shape, fan_in):
        b = 1.0 / np.sqrt(fan_in)
        return rs.uniform(-b, b, shape).astype(np.float32)

    def bn(prefix, c):
        sd[prefix + ".weight"] = np.ones(c, np.float32)
        sd[prefix + ".bias"] = np.zeros(c, np.float32)
        sd[prefix + ".running_mean"] = np.zeros(c, np.float32)
        sd[prefix + ".running_var"] = np.ones(c, np.float32)

    sd["conv_layer.weight"] = uni((filters, 6, 3, 3), 6 * 9)
    bn("bn_layer", filters)
    for b in range(blocks):
        for c in (1, 2):
            sd[f"blocks.{b}.conv{c}.weight"] = uni((filters, filters, 3, 3), filters * 9)
            bn(f"blocks.{b}.bn{c}", filters)
    sd["policy_head.conv_layer.weight"] = uni((2, filters, 1, 1), filters)
    bn("policy_head.bn_layer", 2)
    sd["policy_head.fc_layer.weight"] = uni((nn_ + 1, 2 * nn_), 2 * nn_)
    sd["policy_head.fc_layer.bias"] = uni((nn_ + 1,), 2 * nn_)
    sd["value_head.conv_layer.weight"] = uni((1, filters, 1, 1), filters)
    bn("value_head.bn_layer", 1)
    sd["value_head.fc_layer.weight"] = uni((3, nn_), nn_)
    sd["value_head.fc_layer.bias"] = uni((3,), nn_)
    return sd


def load_state_dict_file(model_file_path):
    """torch.load of a model.bin written by nn/utility.py:80-87 (a DualNet.state_dict())."""
    import torch
    sd = torch.load(model_file_path, map_location="cpu")
    return {k: v.detach().cpu().numpy() for k, v in sd.items() if not k.endswith("num_batches_tracked")}
