"""Harness around the STAGED reference sources (oracle/_ref/src, see stage_reference.py)  --  TEST INFRASTRUCTURE.

Imports the reference's own modules (unmodified) and drives them on the CPU:

* `time_learner` / `time_select_actions`: the reference's `QLearner.train(use_cuda=False)` and
  `BasicMAC.select_actions` on SMAC-shaped synthetic batches - the `cpu_baseline` / `--impl reference` legs of
  bench.py (`kind: "reference"`);
* `run_sequential_args` + `register_synthetic_env`: what the reference's `run.run_sequential` needs to train
  end to end on a synthetic MultiAgentEnv (StarCraft II is not available) - used by the integration test that
  swaps this package's classes into the reference registries.

Only tests/, __graft_entry__ and bench.py's CPU legs import this module (SURVEY.md appendix A is the recipe).
"""
import os
import sys
import time
import types
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, "_ref", "src")


def locate():
    """Path of an importable reference `src` directory: the staged copy, else $PYMARL_REF_SRC, else
    /root/reference/src (build container only); None when there is none."""
    for p in (STAGED, os.environ.get("PYMARL_REF_SRC", ""), "/root/reference/src"):
        if p and os.path.isfile(os.path.join(p, "learners", "q_learner.py")):
            return p
    return None


def available():
    return locate() is not None


_SC2_MOCKS = ["pygame", "pysc2", "pysc2.maps", "pysc2.run_configs", "pysc2.lib", "pysc2.lib.protocol", "pysc2.lib.units",
              "s2clientprotocol", "s2clientprotocol.common_pb2", "s2clientprotocol.sc2api_pb2", "s2clientprotocol.raw_pb2",
              "s2clientprotocol.debug_pb2", "s2clientprotocol.query_pb2"]


def activate(mock_sc2=False):
    """Put the reference `src` first on sys.path (its modules use top-level absolute imports: `components`,
    `controllers`, `learners`, `modules`, ...).  mock_sc2=True additionally pre-seeds sys.modules with stand-ins for
    pygame / pysc2 / s2clientprotocol so that `import run`, `import runners`, `import envs` work without StarCraft II."""
    src = locate()
    if src is None:
        raise RuntimeError("no reference sources: run `python oracle/stage_reference.py` in the build container")
    if src in sys.path:
        sys.path.remove(src)
    sys.path.insert(0, src)
    if mock_sc2:
        from unittest.mock import MagicMock
        for name in _SC2_MOCKS:
            if name not in sys.modules:
                sys.modules[name] = MagicMock()
            if "." in name:                                  # `from pysc2 import maps` reads the parent's attribute
                parent, child = name.rsplit(".", 1)
                setattr(sys.modules[parent], child, sys.modules[name])
        if "pysc2.maps.melee" not in sys.modules:
            melee = types.ModuleType("pysc2.maps.melee")
            melee.Melee = type("Melee", (), {})              # subclassed via type(...) at envs/starcraft2/__init__.py:6-7
            sys.modules["pysc2.maps.melee"] = melee
            sys.modules["pysc2.maps"].melee = melee
            sys.modules["pysc2.maps"].lib = MagicMock()
    return src


class Logger:
    """utils/logging.py Logger's surface without sacred / tensorboard."""

    def __init__(self):
        self.stats, self.infos = {}, []
        self.console_logger = SimpleNamespace(info=lambda msg, *a: self.infos.append(msg))

    def log_stat(self, key, value, t, to_sacred=True):
        self.stats.setdefault(key, []).append((t, float(value)))

    def print_recent_stats(self):
        pass


def scheme_of(shape):
    """run.py:122-135."""
    import torch as th
    from components.transforms import OneHot
    scheme = {
        "state": {"vshape": shape.state_dim},
        "obs": {"vshape": shape.obs_dim, "group": "agents", "vshape_decoded": shape.obs_dim},
        "actions": {"vshape": (1,), "group": "agents", "dtype": th.long},
        "avail_actions": {"vshape": (shape.n_actions,), "group": "agents", "dtype": th.int},
        "reward": {"vshape": (1,)},
        "terminated": {"vshape": (1,), "dtype": th.uint8},
    }
    groups = {"agents": shape.n_agents}
    preprocess = {"actions": ("actions_onehot", [OneHot(out_dim=shape.n_actions)])}
    return scheme, groups, preprocess


def make_batch(shape, fields):
    """The reference's EpisodeBatch holding the given numpy fields (CPU)."""
    import torch as th
    from components.episode_buffer import EpisodeBatch
    B, T = fields["obs"].shape[:2]
    scheme, groups, preprocess = scheme_of(shape)
    batch = EpisodeBatch(scheme, groups, B, T, preprocess=preprocess, device="cpu")
    for k, v in fields.items():
        t = th.from_numpy(np.ascontiguousarray(v))
        assert batch.data.transition_data[k].shape == t.shape, (k, tuple(t.shape))
        batch.data.transition_data[k] = t
    return batch


def make_learner(shape, args, seed=7):
    """The reference's BasicMAC + QLearner on the CPU (run.py:142-148)."""
    import torch as th
    from controllers import REGISTRY as mac_REGISTRY
    from learners import REGISTRY as le_REGISTRY
    th.manual_seed(seed)
    scheme, groups, _ = scheme_of(shape)
    scheme["actions_onehot"] = {"vshape": (shape.n_actions,), "dtype": th.float32, "group": "agents"}
    mac = mac_REGISTRY[args.mac](scheme, groups, args)
    logger = Logger()
    learner = le_REGISTRY[args.learner](mac, scheme, logger, args)
    return learner, logger


def _threads():
    import torch as th
    n = os.cpu_count() or 1
    th.set_num_threads(n)
    return n


def time_learner(shape, args, batch_size, T, steps, warmup, seed=0):
    """Seconds per reference QLearner.train step (use_cuda=False) on `batch_size` full-length synthetic episodes.
    Returns (episodes/s, ms per step, threads)."""
    from pymarl_b200.synthetic import numpy_episode_fields
    activate()
    cores = _threads()
    args.device, args.use_cuda = "cpu", False
    learner, _ = make_learner(shape, args)
    batch = make_batch(shape, numpy_episode_fields(shape, batch_size, T, seed=seed, ragged=False))
    for i in range(warmup):
        learner.train(batch, i, 0)
    times = []
    for i in range(steps):
        t0 = time.perf_counter()
        learner.train(batch, warmup + i, 0)
        times.append(time.perf_counter() - t0)
    mean = sum(times) / len(times)
    return batch_size / mean, mean * 1e3, cores


def time_coma_learner(shape, args, batch_size, T, steps, warmup, seed=0):
    """Seconds per reference COMALearner.train step (use_cuda=False).  Returns (episodes/s, ms per step, threads)."""
    import copy
    import torch as th
    from pymarl_b200.synthetic import numpy_episode_fields
    activate()
    from controllers import REGISTRY as mac_REGISTRY
    from learners import REGISTRY as le_REGISTRY
    cores = _threads()
    args.device, args.use_cuda = "cpu", False
    th.manual_seed(7)
    scheme, groups, _ = scheme_of(shape)
    scheme["actions_onehot"] = {"vshape": (shape.n_actions,), "dtype": th.float32, "group": "agents"}
    # the fork's BasicMAC mutates scheme["obs"]["vshape"] in place, which COMACritic cannot digest: separate copies
    mac = mac_REGISTRY[args.mac](copy.deepcopy(scheme), groups, args)
    learner = le_REGISTRY["coma_learner"](mac, scheme, Logger(), args)
    batch = make_batch(shape, numpy_episode_fields(shape, batch_size, T, seed=seed, ragged=False))
    for i in range(warmup):
        learner.train(batch, i, 0)
    times = []
    for i in range(steps):
        t0 = time.perf_counter()
        learner.train(batch, warmup + i, 0)
        times.append(time.perf_counter() - t0)
    mean = sum(times) / len(times)
    return batch_size / mean, mean * 1e3, cores


def time_select_actions(shape, args, envs, steps, warmup, seed=0):
    """Reference BasicMAC.select_actions (epsilon-greedy at t_env = 0) over `envs` synthetic envs.
    Returns (agent-steps/s, ms per step, threads)."""
    from pymarl_b200.synthetic import numpy_episode_fields
    activate()
    cores = _threads()
    args.device, args.use_cuda = "cpu", False
    learner, _ = make_learner(shape, args)
    mac = learner.mac
    batch = make_batch(shape, numpy_episode_fields(shape, envs, 4, seed=seed, ragged=False))
    mac.init_hidden(envs)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        mac.select_actions(batch, t_ep=1 + i % 3, t_env=25000)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    mean = sum(times) / len(times)
    return envs * shape.n_agents / mean, mean * 1e3, cores


# ---------------------------------------------------------------------------------------------------------------
# end-to-end driver pieces (run.run_sequential on a synthetic environment)
# ---------------------------------------------------------------------------------------------------------------
def register_synthetic_env(name="synthetic"):
    """A MultiAgentEnv (envs/multiagentenv.py) with SMAC-shaped random observations: fixed-length or randomly
    terminating episodes, random availability masks with action 0 always legal, reward = fraction of agents that chose
    action 1 when legal.  Registered in the reference's envs.REGISTRY; returns the class."""
    activate(mock_sc2=True)
    import envs as ref_envs
    from envs.multiagentenv import MultiAgentEnv

    class SyntheticEnv(MultiAgentEnv):
        def __init__(self, n_agents=3, obs_dim=30, state_dim=48, n_actions=9, episode_limit=12, seed=0, p_end=0.05, **kw):
            self.n_agents, self.obs_dim, self.state_dim, self.n_actions = n_agents, obs_dim, state_dim, n_actions
            self.episode_limit = episode_limit
            self.p_end = p_end
            self.rng = np.random.default_rng(seed)
            self.t = 0
            self._roll()

        def _roll(self):
            self._obs = self.rng.standard_normal((self.n_agents, self.obs_dim)).astype(np.float32)
            self._state = self.rng.standard_normal(self.state_dim).astype(np.float32)
            av = (self.rng.random((self.n_agents, self.n_actions)) < 0.6).astype(np.int32)
            av[:, 0] = 1
            self._avail = av

        def reset(self):
            self.t = 0
            self._roll()
            return self.get_obs(), self.get_state()

        def step(self, actions):
            acts = np.asarray([int(a) for a in actions])
            assert all(self._avail[i, a] == 1 for i, a in enumerate(acts)), "illegal action selected"
            reward = float((acts == 1).mean())
            self.t += 1
            self._roll()
            info = {}
            terminated = False
            if self.t >= self.episode_limit:
                terminated, info = True, {"episode_limit": True}
            elif self.rng.random() < self.p_end:
                terminated = True
            return reward, terminated, info

        def get_obs(self):
            return [self._obs[i] for i in range(self.n_agents)]

        def get_obs_agent(self, agent_id):
            return self._obs[agent_id]

        def get_obs_size(self):
            return self.obs_dim

        def get_state(self):
            return self._state

        def get_state_size(self):
            return self.state_dim

        def get_avail_actions(self):
            return [list(self._avail[i]) for i in range(self.n_agents)]

        def get_avail_agent_actions(self, agent_id):
            return list(self._avail[agent_id])

        def get_total_actions(self):
            return self.n_actions

        def get_stats(self):
            return {}

        def close(self):
            pass

        def save_replay(self):
            pass

        def get_env_info(self):
            info = super().get_env_info()
            info.update(obs_decoder=None, avail_actions_encoder_grid=None)       # consumed at run.py:116-117
            return info

    ref_envs.REGISTRY[name] = lambda **kw: SyntheticEnv(**kw)
    return SyntheticEnv


def run_sequential_args(shape, t_max=60, **over):
    """The args namespace run.run_sequential reads (config/default.yaml is missing in the fork: SURVEY.md section 5)."""
    from pymarl_b200.synthetic import default_args
    a = default_args(shape, mixer="qmix")
    extra = dict(runner="episode", env="synthetic",
                 env_args=dict(n_agents=shape.n_agents, obs_dim=shape.obs_dim, state_dim=shape.state_dim,
                               n_actions=shape.n_actions, episode_limit=shape.max_seq_length - 1, seed=0),
                 batch_size_run=1, test_nepisode=2, test_interval=10 ** 9, test_greedy=True, log_interval=10 ** 9,
                 runner_log_interval=10 ** 9, learner_log_interval=1, t_max=t_max, save_model=False,
                 save_model_interval=10 ** 9, checkpoint_path="", load_step=0, evaluate=False, save_replay=False,
                 local_results_path="/tmp/pymarl_b200_results", unique_token="synthetic", buffer_size=16, batch_size=4,
                 buffer_cpu_only=True, use_tensorboard=False, name="qmix", label="test", seed=0, meta=None)
    for k, v in extra.items():
        setattr(a, k, v)
    for k, v in over.items():
        setattr(a, k, v)
    a.device = "cuda" if a.use_cuda else "cpu"
    return a
