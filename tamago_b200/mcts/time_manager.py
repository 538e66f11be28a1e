"""Visit budget of a search (mcts/time_manager.py:12-83).  The device search runs a fixed visit budget; the
time-based modes convert a time allowance into visits with the measured search speed exactly as the reference does."""
import time
from enum import Enum

from ..board.stone import color_value
from .constant import CONST_VISITS, CONST_TIME, REMAINING_TIME, VISITS_PER_SEC


class TimeControl(Enum):
    CONSTANT_PLAYOUT = 0
    CONSTANT_TIME = 1
    TIME_CONTROL = 2
    STRICT_PLAYOUT = 3


class TimeManager:
    def __init__(self, mode, constant_visits=CONST_VISITS, constant_time=CONST_TIME, remaining_time=REMAINING_TIME):
        self.mode = mode
        self.constant_visits = constant_visits
        self.constant_time = constant_time
        self.default_time = remaining_time
        self.search_speed = VISITS_PER_SEC
        self.remaining_time = [remaining_time] * 2
        self.time_limit = 0
        self.start_time = 0

    def initialize(self):
        self.remaining_time = [self.default_time] * 2

    def set_search_speed(self, visits, consumption_time):
        self.search_speed = visits / consumption_time if visits > 0 and consumption_time > 0 else VISITS_PER_SEC

    def get_num_visits_threshold(self, color):
        name = getattr(self.mode, "name", str(self.mode))
        if name in ("CONSTANT_PLAYOUT", "STRICT_PLAYOUT"):
            self.time_limit = 10000.0
            return int(self.constant_visits)
        if name == "CONSTANT_TIME":
            self.time_limit = self.constant_time
            return max(1, int(self.search_speed * self.constant_time))
        if name == "TIME_CONTROL":
            remaining = self.remaining_time[0] if color_value(color) == 1 else self.remaining_time[1]
            self.time_limit = remaining / 10.0
            return max(1, int(self.search_speed * self.time_limit))
        return int(self.constant_visits)

    def is_strict(self):
        return getattr(self.mode, "name", "") == "STRICT_PLAYOUT"

    def set_remaining_time(self, color, remaining_time):
        self.remaining_time[0 if color_value(color) == 1 else 1] = remaining_time

    def substract_consumption_time(self, color, consumption_time):
        self.remaining_time[0 if color_value(color) == 1 else 1] -= consumption_time

    def set_mode(self, mode):
        self.mode = mode

    def start_timer(self):
        self.start_time = time.time()

    def calculate_consumption_time(self):
        return time.time() - self.start_time
