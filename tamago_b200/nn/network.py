"""DualNet handle with the reference's inference surface (nn/network/dual_net.py:81-106, nn/utility.py:139-159).

It holds the parameters on the host (numpy state_dict) and evaluates on the device through the engine's tcgen05
forward pass; MCTSTree / selfplay take the parameters from it and keep everything on the GPU.
"""
import numpy as np

from ..engine import Engine, EVAL_DUALNET_TC
from .utility import random_init_state_dict, load_state_dict_file


class DualNet:
    def __init__(self, board_size=9, device_index=0, blocks=6, seed=0):
        self.board_size, self.device_index, self.blocks = board_size, device_index, blocks
        self.evaluator = EVAL_DUALNET_TC
        self.state_dict_np = random_init_state_dict(board_size, seed, blocks)
        self.training = False
        self._engine = None

    def load_state_dict(self, sd):
        self.state_dict_np = {k: np.asarray(v.detach().cpu().numpy() if hasattr(v, "detach") else v)
                              for k, v in sd.items() if not k.endswith("num_batches_tracked")}
        self._engine = None

    def state_dict(self):
        return dict(self.state_dict_np)

    def eval(self):
        return self

    def _eng(self):
        if self._engine is None:
            self._engine = Engine(board_size=self.board_size, games=1, max_visits=2, evaluator=self.evaluator,
                                  device=self.device_index, net_blocks=self.blocks)
            self._engine.load_state_dict(self.state_dict_np)
        return self._engine

    def _run(self, input_plane, use_logit):
        x = input_plane.detach().cpu().numpy() if hasattr(input_plane, "detach") else np.asarray(input_plane)
        pol, val = self._eng().forward(x, use_logit=use_logit)
        if hasattr(input_plane, "detach"):
            import torch
            return torch.from_numpy(pol), torch.from_numpy(val)
        return pol, val

    def inference(self, input_plane):
        """softmax policy, softmax value (dual_net.py:81-91)"""
        return self._run(input_plane, False)

    def inference_with_policy_logits(self, input_plane):
        """policy logits, softmax value (dual_net.py:94-106)"""
        return self._run(input_plane, True)


def load_network(model_file_path, use_gpu=True, board_size=9, device_index=0):
    """nn/utility.py:139-159: load model.bin; on failure keep the random initialisation and say so."""
    if not use_gpu:
        raise RuntimeError("tamago_b200 has no CPU path: use_gpu must be True")
    net = DualNet(board_size=board_size, device_index=device_index)
    try:
        net.load_state_dict(load_state_dict_file(model_file_path))
    except Exception:  # the reference swallows every failure here (utility.py:152-155)
        print(f"Failed to load {model_file_path}.")
    return net
