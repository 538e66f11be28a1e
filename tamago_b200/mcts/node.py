"""Read-only view of one device tree node with the attribute names of MCTSNode (mcts/node.py:18-39)."""
import json

import numpy as np

from .constant import NOT_EXPANDED


class MCTSNodeView:
    def __init__(self, fields, improved=None, max_actions=None):
        self.num_children = fields["num_children"]
        self.node_visits = fields["node_visits"]
        self.virtual_loss = fields["virtual_loss"]
        self.node_value_sum = fields["node_value_sum"]
        self.raw_value = fields["raw_value"]
        self.action = fields["action"]
        self.children_index = fields["children_index"]
        self.children_value = fields["children_value"]
        self.children_visits = fields["children_visits"]
        self.children_policy = fields["children_policy"]
        self.children_virtual_loss = fields["children_virtual_loss"]
        self.children_value_sum = fields["children_value_sum"]
        self.noise = fields["noise"]
        self._improved = improved
        self._max_actions = max_actions if max_actions is not None else len(self.action)

    def get_num_children(self):
        return self.num_children

    def get_child_move(self, index):
        return int(self.action[index])

    def get_child_index(self, index):
        return int(self.children_index[index])

    def get_best_move_index(self):
        return int(np.argmax(self.children_visits))

    def get_best_move(self):
        return int(self.action[self.get_best_move_index()])

    def calculate_improved_policy(self):
        """mcts/node.py:308-321, computed on the device at the end of the search."""
        return self._improved

    def calculate_value_evaluation(self, index):
        if self.children_visits[index] == 0:
            return 0.5
        return float(self.children_value_sum[index]) / float(self.children_visits[index])

    # -- GTP analysis surface (mcts/node.py:399-482) -----------------------------------------------
    def get_analysis(self, board, mode, pv_lists_func):
        """mcts/node.py:399-413: the lz-analyze ("lz") or cgos-analyze ("cgos") response text of this node."""
        return self.get_analysis_from_status_list(mode, self.get_analysis_status_list(board, pv_lists_func))

    def get_analysis_status_list(self, board, pv_lists_func):
        """mcts/node.py:416-449: children with visits, most visited first (ties: higher child index first)."""
        order_list = sorted(((int(self.children_visits[i]), i) for i in range(self.num_children)), reverse=True)
        coordinate = board.coordinate
        pv_lists = pv_lists_func(self, coordinate)
        status, order = [], 0
        for visits, i in order_list:
            if visits == 0:
                continue
            move = coordinate.convert_to_gtp_format(int(self.action[i]))
            winrate = float(self.children_value_sum[i]) / visits
            status.append({"move": move, "visits": int(visits), "winrate": float(winrate), "prior": float(self.children_policy[i]),
                           "lcb": float(winrate), "order": int(order), "pv": " ".join(f"{p}" for p in pv_lists[move])})
            order += 1
        return status

    def get_analysis_from_status_list(self, mode, children_status_list):
        """mcts/node.py:452-482."""
        if mode == "cgos":
            cgos = {"winrate": float(self.node_value_sum) / self.node_visits, "visits": int(self.node_visits),
                    "moves": list(children_status_list)}
            return json.dumps(cgos, indent=None, separators=(",", ":")) + "\n"
        out = ""
        if mode == "lz":
            for st in children_status_list:
                out += (f"info move {st['move']} visits {st['visits']} winrate {int(10000 * st['winrate'])} "
                        f"prior {int(10000 * st['prior'])} lcb {int(10000 * st['lcb'])} order {st['order']} pv {st['pv']} ")
        return out[:-1] + "\n"

    def to_dict(self):
        """mcts/node.py:221-243: every per-child array has MAX_ACTIONS entries (unused ones as MCTSNode.expand leaves them)."""
        n, k = self._max_actions, self.num_children

        def pad(a, fill, cast):
            return [cast(v) for v in a[:k]] + [fill] * (n - k)
        noise = [float(v) for v in self.noise[:n]] + [0.0] * max(0, n - len(self.noise))
        return {"node_visits": int(self.node_visits), "virtual_loss": int(self.virtual_loss),
                "node_value_sum": float(self.node_value_sum), "raw_value": float(self.raw_value),
                "action": pad(self.action, 0, int), "children_index": pad(self.children_index, NOT_EXPANDED, int),
                "children_value": pad(self.children_value, 0.0, float), "children_visits": pad(self.children_visits, 0, int),
                "children_policy": pad(self.children_policy, 0.0, float),
                "children_virtual_loss": pad(self.children_virtual_loss, 0, int),
                "children_value_sum": pad(self.children_value_sum, 0.0, float), "noise": noise, "num_children": int(k)}
