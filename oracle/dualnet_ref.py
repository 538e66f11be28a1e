"""CPU restatement of the reference DualNet forward pass -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows nn/network/dual_net.py:41-106 (forward / inference / inference_with_policy_logits),
nn/network/res_block.py:27-39, nn/network/head/policy_head.py:25-40, nn/network/head/value_head.py:26-40 with
plain torch fp32 ops on the CPU.  Pinned by tests/test_oracle_dualnet.py against outputs of the reference's own
DualNet (tests/golden/dualnet_<N>.npz).  Used as the evaluator of the C oracle's search in bench.py's
cpu_baseline / --impl reference legs and in smoke().
"""
import numpy as np
import torch
import torch.nn.functional as F


class DualNetRef:
    def __init__(self, state_dict, board_size):
        self.n = board_size
        self.t = {k: torch.from_numpy(np.ascontiguousarray(np.asarray(v))).float()
                  for k, v in state_dict.items() if not k.endswith("num_batches_tracked")}
        self.blocks = 0
        while f"blocks.{self.blocks}.conv1.weight" in self.t:
            self.blocks += 1

    def _bn(self, h, p, eps):
        t = self.t
        return F.batch_norm(h, t[p + ".running_mean"], t[p + ".running_var"], t[p + ".weight"], t[p + ".bias"], False, 0.0, eps)

    @torch.no_grad()
    def forward(self, x):
        t = self.t
        h = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).reshape(-1, 6, self.n, self.n)
        h = F.relu(self._bn(F.conv2d(h, t["conv_layer.weight"], padding=1), "bn_layer", 1e-5))        # dual_net.py:50
        for b in range(self.blocks):                                                                # res_block.py:36-39
            h1 = F.relu(self._bn(F.conv2d(h, t[f"blocks.{b}.conv1.weight"], padding=1), f"blocks.{b}.bn1", 2e-5))
            h2 = self._bn(F.conv2d(h1, t[f"blocks.{b}.conv2.weight"], padding=1), f"blocks.{b}.bn2", 2e-5)
            h = F.relu(h + h2)
        p = F.relu(self._bn(F.conv2d(h, t["policy_head.conv_layer.weight"]), "policy_head.bn_layer", 2e-5)).flatten(1)
        logits = F.linear(p, t["policy_head.fc_layer.weight"], t["policy_head.fc_layer.bias"])      # policy_head.py:37-40
        v = F.relu(self._bn(F.conv2d(h, t["value_head.conv_layer.weight"]), "value_head.bn_layer", 2e-5)).flatten(1)
        vlog = F.linear(v, t["value_head.fc_layer.weight"], t["value_head.fc_layer.bias"])          # value_head.py:38-40
        return logits, vlog

    def evaluator(self):
        """Callable for oracle.OracleTree: (planes, use_logit) -> (policy, value) like DualNet.inference*."""
        def ev(planes, use_logit):
            logits, vlog = self.forward(planes)
            pol = logits if use_logit else torch.softmax(logits, 1)                                 # dual_net.py:81-106
            return pol.numpy(), torch.softmax(vlog, 1).numpy()
        return ev
