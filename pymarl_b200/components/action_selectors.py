"""Action selectors (reference: components/action_selectors.py:35-65).

EpsilonGreedyActionSelector.select_action runs the fused CUDA kernel (pmb_epsilon_greedy).
Two draw modes:
  rng="torch"  (default) the uniform and Exp(1) draws come from torch's generator in the
               reference's order (rand_like first, then the exponential_ that
               Categorical.sample()/multinomial consumes), so with the same seed on the same
               device the selected actions are bit-identical to the reference's;
  rng="philox" the kernel draws from its own Philox4x32-10 stream (no extra launches).
"""
import ctypes as C

import torch as th

from .epsilon_schedules import DecayThenFlatSchedule
from .. import _lib

REGISTRY = {}


class MultinomialActionSelector:
    """reference: components/action_selectors.py:9-33 (COMA).  select_action samples from the MAC's (already epsilon-
    floored) policy with pmb_multinomial: Categorical(masked probs).sample() is arg-max(p / Exp(1)); rng="torch" draws the
    exponentials from torch's generator in the reference's order (bit-identical actions for the same seed on the same
    device), rng="philox" lets the kernel generate them."""

    def __init__(self, args):
        self.args = args
        self.schedule = DecayThenFlatSchedule(args.epsilon_start, args.epsilon_finish, args.epsilon_anneal_time,
                                              decay="linear")
        self.epsilon = self.schedule.eval(0)
        self.test_greedy = getattr(args, "test_greedy", True)
        self.rng = getattr(args, "action_rng", "torch")
        self._offset = 0

    def draw(self, agent_inputs):
        if self.rng != "torch":
            return None
        b, n, a = agent_inputs.shape
        return th.empty(b * n, a, dtype=th.float32, device=agent_inputs.device).exponential_()

    def select_action(self, agent_inputs, avail_actions, t_env, test_mode=False):
        self.epsilon = self.schedule.eval(t_env)
        _lib.require_cuda(agent_inputs, "agent_inputs")
        b, n, a = agent_inputs.shape
        probs = agent_inputs.detach().to(th.float32).contiguous()
        avail = avail_actions.to(device=probs.device, dtype=th.int32).contiguous()
        greedy = bool(test_mode and self.test_greedy)
        expo = None if greedy else self.draw(agent_inputs)
        self._offset += 1
        out = th.empty(b, n, dtype=th.int64, device=probs.device)
        _lib.check(_lib.lib().pmb_multinomial(b * n, a, _lib.ptr(probs), _lib.ptr(avail), _lib.ptr(expo), int(greedy),
                                              th.initial_seed() & 0xFFFFFFFFFFFFFFFF, self._offset, _lib.ptr(out),
                                              _lib.stream_ptr(probs.device)), "pmb_multinomial")
        return out


REGISTRY["multinomial"] = MultinomialActionSelector


class EpsilonGreedyActionSelector:

    def __init__(self, args):
        self.args = args
        self.schedule = DecayThenFlatSchedule(args.epsilon_start, args.epsilon_finish, args.epsilon_anneal_time,
                                              decay="linear")
        self.epsilon = self.schedule.eval(0)
        self.rng = getattr(args, "action_rng", "torch")
        self._philox_offset = 0

    def draw(self, agent_inputs):
        """(u, expo) in the reference's generator order, or (None, None) in philox mode."""
        if self.rng != "torch":
            return None, None
        b, n, a = agent_inputs.shape
        u = th.rand_like(agent_inputs[:, :, 0]).contiguous()
        expo = th.empty(b * n, a, dtype=th.float32, device=agent_inputs.device).exponential_()
        return u, expo

    def next_philox(self):
        self._philox_offset += 1
        return th.initial_seed() & 0xFFFFFFFFFFFFFFFF, self._philox_offset

    def select_action(self, agent_inputs, avail_actions, t_env, test_mode=False):
        """agent_inputs [b, N, A] Q-values, avail_actions [b, N, A] -> LongTensor [b, N]."""
        self.epsilon = self.schedule.eval(t_env)
        if test_mode:
            self.epsilon = 0.0
        _lib.require_cuda(agent_inputs, "agent_inputs")
        b, n, a = agent_inputs.shape
        q = agent_inputs.detach().to(th.float32).contiguous()
        avail = avail_actions.to(device=q.device, dtype=th.int32).contiguous()
        u, expo = self.draw(agent_inputs)
        seed, offset = (0, 0) if u is not None else self.next_philox()
        out = th.empty(b, n, dtype=th.int64, device=q.device)
        _lib.check(_lib.lib().pmb_epsilon_greedy(b * n, a, _lib.ptr(q), _lib.ptr(avail), C.c_float(self.epsilon),
                                                 _lib.ptr(u), _lib.ptr(expo), seed, offset, _lib.ptr(out),
                                                 _lib.stream_ptr(q.device)), "pmb_epsilon_greedy")
        return out


REGISTRY["epsilon_greedy"] = EpsilonGreedyActionSelector
