"""Read-only view of one device tree node with the attribute names of MCTSNode (mcts/node.py:18-39)."""
import numpy as np


class MCTSNodeView:
    def __init__(self, fields, improved=None):
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
