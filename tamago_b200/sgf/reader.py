"""Minimal SGF reader for self-play records (the subset of sgf/reader.py:33-380 that nn/data_generator.py uses):
board size, result, moves with colours and the per-move comment."""
import re

from ..board.constant import PASS

_NODE = re.compile(r";\s*([BW])\[([a-zA-Z]{0,2})\](?:\s*C\[((?:[^\]\\]|\\.)*)\])?")


class SGFReader:
    def __init__(self, kifu_path_or_text, board_size, literal=False):
        text = kifu_path_or_text if literal else open(kifu_path_or_text, encoding="utf-8").read()
        self.board_size = board_size
        m = re.search(r"SZ\[(\d+)\]", text)
        self.size = int(m.group(1)) if m else board_size
        m = re.search(r"RE\[([^\]]*)\]", text)
        re_field = m.group(1).strip().upper() if m else "0"
        self.result = "B" if re_field.startswith("B") else ("W" if re_field.startswith("W") else "D")   # reader.py:241-263
        m = re.search(r"KM\[([^\]]*)\]", text)
        self.komi = float(m.group(1)) if m else 7.0
        self.moves, self.colors, self.comments = [], [], []
        w = self.size + 2
        for col, xy, comment in _NODE.findall(text):
            if xy in ("", "tt") and self.size <= 19:
                pos = PASS
            else:
                pos = (ord(xy[0].lower()) - 96) + (ord(xy[1].lower()) - 96) * w
            self.moves.append(pos); self.colors.append(1 if col == "B" else 2); self.comments.append(comment or "")

    def get_n_moves(self):
        return len(self.moves)

    def get_moves(self):
        return list(self.moves)

    def get_comment(self, index):
        return self.comments[index]

    def get_value_label(self):
        """reader.py:345-358: black win 2, white win 0, draw 1."""
        return 2 if self.result == "B" else (0 if self.result == "W" else 1)
