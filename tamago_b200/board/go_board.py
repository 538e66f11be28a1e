"""Host-side GoBoard with the reference's call surface (board/go_board.py:17-626).

The rules live on the GPU: this class only records the move history; every rule query replays that history on a
one-game device board through the C ABI (tg_play with a state dump) and reads the answer back.  It exists so that
host code written against the reference (GTP front-end, self-play loop) keeps working; the search itself takes the
history and never calls back into Python.
"""
import numpy as np

from .constant import PASS, OB_SIZE
from .coordinate import Coordinate
from .stone import Stone, color_value

_ENGINES = {}


def _rule_engine(size, superko, zobrist=None):
    from ..engine import Engine, EVAL_HASHNET
    key = (size, bool(superko))
    if key not in _ENGINES:
        _ENGINES[key] = Engine(board_size=size, games=1, max_visits=2, superko=superko, evaluator=EVAL_HASHNET)
    return _ENGINES[key]


class GoBoard:
    def __init__(self, board_size, komi=7.0, check_superko=False):
        self.board_size = board_size
        self.board_size_with_ob = board_size + OB_SIZE * 2
        self.komi = komi
        self.check_superko = check_superko
        self.coordinate = Coordinate(board_size)
        self.onboard_pos = [(x + OB_SIZE) + (y + OB_SIZE) * self.board_size_with_ob
                            for y in range(board_size) for x in range(board_size)]
        self.clear()

    # -- state ------------------------------------------------------------------------------------
    def clear(self):
        self.history = []            # (Stone, pos)
        self.handicap = []
        self._cache = None

    @property
    def moves(self):
        return len(self.history) + 1

    def put_stone(self, pos, color):
        self.history.append((Stone(color_value(color)), int(pos)))
        self._cache = None

    def get_board_size(self):
        return self.board_size

    def get_komi(self):
        return self.komi

    def set_komi(self, komi):
        self.komi = komi

    def get_to_move(self):
        return Stone.BLACK if not self.history else Stone.get_opponent_color(self.history[-1][0])

    def get_move_history(self):
        return [(c, p, None) for c, p in self.history]

    def get_handicap_history(self):
        return self.handicap[:]

    # -- rule queries (device) --------------------------------------------------------------------
    def _state(self):
        if self._cache is None:
            e = _rule_engine(self.board_size, self.check_superko)
            e.reset()
            if not self.history:
                self._cache = None, e
                return self._cache
            mv = np.array([[p for _, p in self.history]], np.int16)
            col = np.array([[c.value for c, _ in self.history]], np.uint8)
            d = e.play(mv, colors=col, dump=True)
            self._cache = {k: v[0, -1] for k, v in d.items()}, e
        return self._cache

    def _col_index(self, color):
        return 0 if color_value(color) == 1 else 1

    def is_legal(self, pos, color):
        d, _ = self._state()
        if pos == PASS:
            return False
        idx = self.onboard_pos.index(pos)
        if d is None:
            return True
        return bool(d["legal"][self._col_index(color)][idx])

    def get_all_legal_pos(self, color):
        d, _ = self._state()
        if d is None:
            return list(self.onboard_pos)
        m = d["legal"][self._col_index(color)]
        return [p for p, ok in zip(self.onboard_pos, m) if ok]

    def get_candidates(self, color):
        """Expansion candidates of mcts/tree.py:260-264 (legal, self-atari < 7, not a complete eye), PASS last."""
        d, _ = self._state()
        if d is None:
            return list(self.onboard_pos) + [PASS]
        m = d["cand"][self._col_index(color)]
        return [p for p, ok in zip(self.onboard_pos, m) if ok] + [PASS]

    def count_score(self):
        d, _ = self._state()
        return 0 if d is None else int(d["score"])

    def get_board_data(self):
        d, _ = self._state()
        if d is None:
            return [0] * (self.board_size ** 2)
        return [int(d["color"][p]) for p in self.onboard_pos]

    def get_hash(self):
        d, _ = self._state()
        return 0 if d is None else int(d["hash"])


def copy_board(dst, src):
    """board/go_board.py:611-626."""
    dst.history = src.history[:]
    dst.handicap = src.handicap[:]
    dst.komi = src.komi
    dst._cache = None
