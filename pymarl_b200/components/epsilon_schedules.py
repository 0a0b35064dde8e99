"""Epsilon schedule used by the action selector (reference: components/epsilon_schedules.py:4-25)."""
import math


class DecayThenFlatSchedule:
    """Decay from `start` to `finish` over `time_length` env steps, then stay flat.
    decay = "linear": max(finish, start - (start - finish) / time_length * T)
    decay = "exp":    min(start, max(finish, exp(-T / scaling)))"""

    def __init__(self, start, finish, time_length, decay="exp"):
        self.start, self.finish, self.time_length, self.decay = start, finish, time_length, decay
        self.delta = (self.start - self.finish) / self.time_length
        if self.decay == "exp":
            self.exp_scaling = (-1) * self.time_length / math.log(self.finish) if self.finish > 0 else 1

    def eval(self, T):
        if self.decay == "linear":
            return max(self.finish, self.start - self.delta * T)
        if self.decay == "exp":
            return min(self.start, max(self.finish, math.exp(-T / self.exp_scaling)))
        return None
