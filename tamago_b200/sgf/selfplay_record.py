"""Self-play record (sgf/selfplay_record.py:13-110): per-move improved-policy comments, SGF text written by the
library's C++ writer (tg_format_sgf)."""
import os

import numpy as np

from ..engine import format_sgf


class SelfPlayRecord:
    def __init__(self, save_dir, board_size, komi=7.0):
        self.save_dir, self.board_size, self.komi = save_dir, board_size, komi
        self.clear()

    def clear(self):
        self.moves, self.colors, self.k, self.action, self.improved = [], [], [], [], []

    def save_record(self, pos, color, num_children, action, improved):
        self.moves.append(int(pos)); self.colors.append(int(color)); self.k.append(int(num_children))
        self.action.append(np.array(action, np.int16)); self.improved.append(np.array(improved, np.float64))

    def text(self, winner, is_resign, score):
        width = max([len(a) for a in self.action] + [1])
        action = np.zeros((len(self.moves), width), np.int16)
        improved = np.zeros((len(self.moves), width), np.float64)
        for i, (a, p) in enumerate(zip(self.action, self.improved)):
            action[i, :len(a)] = a
            improved[i, :len(p)] = p
        return format_sgf(self.board_size, self.moves, self.colors, self.k, action, improved, winner, is_resign, score, self.komi)

    def write_record(self, index, winner, is_resign, score):
        path = os.path.join(self.save_dir, f"{index}.sgf")
        with open(path, mode="w", encoding="utf-8") as f:
            f.write(self.text(winner, is_resign, score))
        return path
