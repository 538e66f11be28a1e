"""Stone colours (board/stone.py:5)."""
from enum import Enum


class Stone(Enum):
    EMPTY = 0
    BLACK = 1
    WHITE = 2
    OUT_OF_BOARD = 3

    @classmethod
    def get_opponent_color(cls, color):
        if color == Stone.BLACK:
            return Stone.WHITE
        if color == Stone.WHITE:
            return Stone.BLACK
        return color


def color_value(color):
    """Stone (ours or the reference's) or int -> 1 / 2."""
    return int(getattr(color, "value", color))
