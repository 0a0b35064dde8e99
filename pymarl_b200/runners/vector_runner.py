"""VectorRunner: the ParallelRunner loop (reference: runners/parallel_runner.py:60-204) for environments that step as ONE
batched object on the device.

The reference runs `batch_size_run` env processes and moves every observation / action through pipes and pickle each
timestep, then scatters them into the EpisodeBatch with per-field indexed assignments.  Here the env batch lives on the
GPU (`SyntheticVectorEnv`: SMAC-shaped random dynamics, StarCraft II is not available), one timestep is

    mac.select_actions        -> the fused rollout kernels (fc1 -> GRU -> fc2 -> avail mask -> epsilon-greedy)
    batch.update(actions)     -> ONE pmb_batch_update launch (actions + fused OneHot)
    env.step                  -> device ops
    batch.update(reward, terminated), batch.update(state, avail_actions, obs; mark_filled)   -> one launch each

with NO host synchronisation per timestep: the live-env masks stay on the device and the updates are masked launches
(`EpisodeBatch.update_masked`); the loop checks "all terminated" every 16 steps and reads the step count once per run.  Episode semantics follow the reference
loop exactly: actions are also selected and stored in an env's final state, `terminated` is stored as 0 when the episode
ended on the time limit (parallel_runner.py:150-156), `filled` covers t = 0 .. L, `t_env` counts env steps of training
runs only."""
from functools import partial

import torch as th

from ..components.episode_buffer import EpisodeBatch
from ..components.transforms import OneHot


class SyntheticVectorEnv:
    """B independent SMAC-shaped synthetic envs as one device-resident object (same statistics as
    pymarl_b200.synthetic: N(0,1) obs / state, Bernoulli(0.6) availability with action 0 always legal, episodes end at
    `episode_limit` or earlier with probability `p_end` per step; reward = fraction of agents that picked action 1)."""

    def __init__(self, n_envs, n_agents, obs_dim, state_dim, n_actions, episode_limit, seed=0, p_end=0.02, device="cuda"):
        self.n_envs, self.n_agents, self.obs_dim, self.state_dim, self.n_actions = n_envs, n_agents, obs_dim, state_dim, n_actions
        self.episode_limit, self.p_end, self.device = episode_limit, p_end, th.device(device)
        self.gen = th.Generator(device=self.device)
        self.gen.manual_seed(seed)
        self.t = None
        self.reset()

    def _roll(self):
        B, N, d = self.n_envs, self.n_agents, self.device
        self.obs = th.randn(B, N, self.obs_dim, generator=self.gen, device=d)
        self.state = th.randn(B, self.state_dim, generator=self.gen, device=d)
        av = (th.rand(B, N, self.n_actions, generator=self.gen, device=d) < 0.6).to(th.int32)
        av[..., 0] = 1
        self.avail = av

    def reset(self):
        self.t = th.zeros(self.n_envs, dtype=th.long, device=self.device)
        self._roll()

    def step(self, actions, alive):
        """actions [B, N] int64; `alive` [B] bool: only those envs advance.  -> (reward [B], terminated [B] bool,
        episode_limit [B] bool: terminated because the time limit was reached)."""
        reward = (actions == 1).float().mean(1)
        self.t = self.t + alive.long()
        limit = self.t >= self.episode_limit
        ended = th.rand(self.n_envs, generator=self.gen, device=self.device) < self.p_end
        self._roll()
        return reward, (limit | ended) & alive, limit & alive

    def get_obs(self):
        return self.obs

    def get_state(self):
        return self.state

    def get_avail_actions(self):
        return self.avail

    def get_env_info(self):
        return {"state_shape": self.state_dim, "obs_shape": self.obs_dim, "n_actions": self.n_actions, "n_agents": self.n_agents,
                "episode_limit": self.episode_limit, "obs_decoder": None, "avail_actions_encoder_grid": None}

    def get_stats(self):
        return {}

    def close(self):
        pass

    def save_replay(self):
        pass


class VectorRunner:
    def __init__(self, args, env=None, logger=None):
        self.args, self.logger = args, logger
        self.batch_size = args.batch_size_run
        if env is None:
            env = SyntheticVectorEnv(self.batch_size, **args.env_args)
        assert env.n_envs == self.batch_size
        self.env = env
        self.episode_limit = env.episode_limit
        self.t = 0
        self.t_env = 0
        self.train_returns, self.test_returns = [], []
        self.log_train_stats_t = -100000

    def setup(self, mac, scheme=None, groups=None, preprocess=None):
        """run.py:145: runner.setup(scheme=..., groups=..., preprocess=..., mac=mac); all but `mac` default to the scheme
        run.py:122-135 builds from the env info."""
        info = self.env.get_env_info()
        if scheme is None:
            scheme = {"state": {"vshape": info["state_shape"]},
                      "obs": {"vshape": info["obs_shape"], "group": "agents", "vshape_decoded": info["obs_shape"]},
                      "actions": {"vshape": (1,), "group": "agents", "dtype": th.long},
                      "avail_actions": {"vshape": (info["n_actions"],), "group": "agents", "dtype": th.int},
                      "reward": {"vshape": (1,)}, "terminated": {"vshape": (1,), "dtype": th.uint8}}
            groups = {"agents": info["n_agents"]}
            preprocess = {"actions": ("actions_onehot", [OneHot(out_dim=info["n_actions"])])}
        self.new_batch = partial(EpisodeBatch, scheme, groups, self.batch_size, self.episode_limit + 1, preprocess=preprocess,
                                 device=str(self.env.device))
        self.mac = mac

    def get_env_info(self):
        return self.env.get_env_info()

    def save_replay(self):
        self.env.save_replay()

    def close_env(self):
        self.env.close()

    def reset(self):
        self.batch = self.new_batch()
        self.env.reset()
        self.batch.update({"state": self.env.get_state(), "avail_actions": self.env.get_avail_actions(),
                           "obs": self.env.get_obs()}, ts=0)
        self.t = 0
        self.env_steps_this_run = 0

    def run(self, test_mode=False):
        """One batch of episodes.  No host synchronisation inside the loop except an all-terminated check every 16
        timesteps: live-env masks stay on the device and every batch.update is a masked single launch."""
        self.reset()
        B, dev = self.batch_size, self.env.device
        self.mac.init_hidden(batch_size=B)
        alive = th.ones(B, dtype=th.bool, device=dev)       # envs that still step
        store = alive.clone()                               # envs whose action at this t is stored (alive one step ago)
        returns = th.zeros(B, device=dev)
        steps = th.zeros((), dtype=th.long, device=dev)
        while True:
            actions = self.mac.select_actions(self.batch, t_ep=self.t, t_env=self.t_env, test_mode=test_mode)     # [B, N]
            self.batch.update_masked({"actions": actions.unsqueeze(-1)}, store, self.t, mark_filled=False)
            if self.t >= self.episode_limit or (self.t % 16 == 15 and not bool(alive.any())):
                break                                       # at t = limit every env has terminated (time limit)
            reward, term, limit = self.env.step(actions, alive)
            returns += reward * alive
            steps += alive.sum()
            self.batch.update_masked({"reward": reward.unsqueeze(-1), "terminated": (term & ~limit).unsqueeze(-1)},
                                     alive, self.t, mark_filled=False)
            self.t += 1
            self.batch.update_masked({"state": self.env.get_state(), "avail_actions": self.env.get_avail_actions(),
                                      "obs": self.env.get_obs()}, alive, self.t, mark_filled=True)
            store = alive
            alive = alive & ~term
        if not test_mode:
            self.env_steps_this_run = int(steps)            # the one device -> host read of the run
        if not test_mode:
            self.t_env += self.env_steps_this_run
        (self.test_returns if test_mode else self.train_returns).extend(returns.tolist())
        if self.logger is not None and not test_mode and self.t_env - self.log_train_stats_t >= getattr(self.args, "runner_log_interval", 0):
            rs = self.train_returns
            self.logger.log_stat("return_mean", sum(rs) / max(1, len(rs)), self.t_env)
            if hasattr(self.mac.action_selector, "epsilon"):
                self.logger.log_stat("epsilon", self.mac.action_selector.epsilon, self.t_env)
            rs.clear()
            self.log_train_stats_t = self.t_env
        return self.batch
