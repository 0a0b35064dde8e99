"""SMAC-shaped synthetic episode data (StarCraft II is not available).

The generator produces the EpisodeBatch fields QLearner.train reads, in the
reference's layout and dtypes (reference: src/components/episode_buffer.py:58-86,
scheme in src/run.py:122-135):

    state          [B, T, S]     float32
    obs            [B, T, N, O]  float32
    actions        [B, T, N, 1]  int64
    avail_actions  [B, T, N, A]  int32
    actions_onehot [B, T, N, A]  float32   (OneHot preprocess, transforms.py:12-21)
    reward         [B, T, 1]     float32
    terminated     [B, T, 1]     uint8
    filled         [B, T, 1]     int64

Shapes follow SURVEY.md section 8 (map limits from src/envs/starcraft2/maps/map_params.py).
Two generators share the same episode structure:

* ``numpy_episode_fields``  - host, numpy ``default_rng(seed)``; bit-reproducible on every
  machine.  Used by the tests (oracle and CUDA path see the same arrays) and the golden
  fixtures.
* ``torch_episode_fields``  - builds the fields directly on a torch device (used by
  bench.py for the 32 GB 27m_vs_30m batch, where a host generator would take minutes).
"""
from collections import namedtuple

import numpy as np

SmacShape = namedtuple("SmacShape", "name n_agents obs_dim state_dim n_actions max_seq_length")

# T = episode_limit + 1 (src/run.py:137).
SMAC_SHAPES = {
    "3m": SmacShape("3m", 3, 30, 48, 9, 61),
    "2s3z": SmacShape("2s3z", 5, 80, 120, 11, 121),
    "MMM2": SmacShape("MMM2", 10, 176, 322, 18, 181),
    "27m_vs_30m": SmacShape("27m_vs_30m", 27, 285, 1170, 36, 181),
}

# BASELINE.json quotes T = 60/120/180/180 and these batch sizes.
BASELINE_CONFIGS = {
    "3m": dict(shape="3m", T=60, batch=32, mixer="qmix"),
    "2s3z": dict(shape="2s3z", T=120, batch=1024, mixer="qmix"),
    "MMM2_vdn": dict(shape="MMM2", T=180, batch=2048, mixer="vdn"),
    "MMM2_iql": dict(shape="MMM2", T=180, batch=2048, mixer=None),
    "27m_vs_30m": dict(shape="27m_vs_30m", T=180, batch=4096, mixer="qmix"),
}


def get_shape(shape):
    if isinstance(shape, SmacShape):
        return shape
    if isinstance(shape, str):
        return SMAC_SHAPES[shape]
    return SmacShape(*shape)


def make_scheme(shape):
    """scheme / groups exactly as src/run.py:122-135 builds them (torch dtypes)."""
    import torch as th
    s = get_shape(shape)
    scheme = {
        "state": {"vshape": s.state_dim},
        "obs": {"vshape": s.obs_dim, "group": "agents", "vshape_decoded": s.obs_dim},
        "actions": {"vshape": (1,), "group": "agents", "dtype": th.long},
        "avail_actions": {"vshape": (s.n_actions,), "group": "agents", "dtype": th.int},
        "reward": {"vshape": (1,)},
        "terminated": {"vshape": (1,), "dtype": th.uint8},
    }
    groups = {"agents": s.n_agents}
    return scheme, groups


def numpy_episode_fields(shape, batch_size, T=None, seed=0, ragged=True, avail_p=0.6):
    """Host generator.  Episode b has L_b transitions, L_b ~ U{T//2 .. T-1} (a quarter of the
    episodes, and all of them when not ragged, run to the time limit L_b = T-1).
    filled = 1 for t <= L_b (the runner also stores the final observation).  Episodes
    shorter than the limit ended in the environment, so terminated[b, L_b-1] = 1; episodes
    that reach the limit store terminated = 1 for about half (env ended on the last step)
    and 0 for the rest (time-limit ends store 0, src/runners/episode_runner.py:75).
    Every field is zero beyond L_b, avail has at least one legal action on filled steps and
    none on padded steps, actions are uniformly random legal actions."""
    s = get_shape(shape)
    T = s.max_seq_length if T is None else int(T)
    B, N, O, S, A = batch_size, s.n_agents, s.obs_dim, s.state_dim, s.n_actions
    rng = np.random.default_rng(seed)
    if ragged:
        L = rng.integers(max(1, T // 2), T, size=B)       # in [T//2, T-1]
        L = np.where(rng.random(B) < 0.25, T - 1, L)
    else:
        L = np.full(B, T - 1, dtype=np.int64)
    t_idx = np.arange(T)[None, :]
    filled = (t_idx <= L[:, None])                             # [B, T]
    has_transition = (t_idx < L[:, None])
    env_terminated = (rng.random(B) < 0.5) | (L < T - 1)
    terminated = np.zeros((B, T), dtype=np.uint8)
    terminated[np.arange(B), L - 1] = env_terminated.astype(np.uint8)

    state = rng.standard_normal((B, T, S), dtype=np.float32) * filled[:, :, None]
    obs = rng.standard_normal((B, T, N, O), dtype=np.float32) * filled[:, :, None, None]
    reward = rng.standard_normal((B, T, 1), dtype=np.float32) * has_transition[:, :, None]

    avail = (rng.random((B, T, N, A)) < avail_p)
    avail[..., 0] = True
    avail &= filled[:, :, None, None]
    # uniformly random legal action: argmax of iid keys restricted to legal actions
    keys = rng.random((B, T, N, A)) * avail
    actions = keys.argmax(-1).astype(np.int64)
    actions *= filled[:, :, None]
    onehot = np.zeros((B, T, N, A), dtype=np.float32)
    np.put_along_axis(onehot, actions[..., None], 1.0, axis=-1)
    # the OneHot preprocess also fires on padded steps of an inserted batch (they hold
    # action 0); the reference zero-initialises and only writes filled steps, so keep
    # padded one-hots zero.
    onehot *= filled[:, :, None, None]
    return {
        "state": state.astype(np.float32),
        "obs": obs.astype(np.float32),
        "actions": actions[..., None],
        "avail_actions": avail.astype(np.int32),
        "actions_onehot": onehot,
        "reward": reward.astype(np.float32),
        "terminated": terminated[..., None],
        "filled": filled.astype(np.int64)[..., None],
    }


def torch_episode_fields(shape, batch_size, T=None, seed=0, ragged=False, avail_p=0.6,
                         device="cuda", with_onehot=True, chunk=256):
    """Device generator with the same episode structure (different random stream)."""
    import torch as th
    s = get_shape(shape)
    T = s.max_seq_length if T is None else int(T)
    B, N, O, S, A = batch_size, s.n_agents, s.obs_dim, s.state_dim, s.n_actions
    g = th.Generator(device=device)
    g.manual_seed(seed)
    if ragged:
        L = th.randint(max(1, T // 2), T, (B,), generator=g, device=device)
        L = th.where(th.rand(B, generator=g, device=device) < 0.25, th.full_like(L, T - 1), L)
    else:
        L = th.full((B,), T - 1, dtype=th.long, device=device)
    t_idx = th.arange(T, device=device)[None, :]
    filled = t_idx <= L[:, None]
    has_tr = t_idx < L[:, None]
    env_term = (th.rand(B, generator=g, device=device) < 0.5) | (L < T - 1)
    terminated = th.zeros(B, T, dtype=th.uint8, device=device)
    terminated[th.arange(B, device=device), L - 1] = env_term.to(th.uint8)

    out = {
        "state": th.empty(B, T, S, dtype=th.float32, device=device),
        "obs": th.empty(B, T, N, O, dtype=th.float32, device=device),
        "actions": th.empty(B, T, N, 1, dtype=th.long, device=device),
        "avail_actions": th.empty(B, T, N, A, dtype=th.int32, device=device),
        "reward": (th.randn(B, T, 1, generator=g, device=device) * has_tr[:, :, None]),
        "terminated": terminated[..., None],
        "filled": filled.long()[..., None],
    }
    if with_onehot:
        out["actions_onehot"] = th.zeros(B, T, N, A, dtype=th.float32, device=device)
    for b0 in range(0, B, chunk):           # bounded temporaries
        b1 = min(B, b0 + chunk)
        f = filled[b0:b1]
        out["state"][b0:b1] = th.randn(b1 - b0, T, S, generator=g, device=device) * f[:, :, None]
        out["obs"][b0:b1] = th.randn(b1 - b0, T, N, O, generator=g, device=device) * f[:, :, None, None]
        av = th.rand(b1 - b0, T, N, A, generator=g, device=device) < avail_p
        av[..., 0] = True
        av &= f[:, :, None, None]
        keys = th.rand(b1 - b0, T, N, A, generator=g, device=device) * av
        act = keys.argmax(-1) * f[:, :, None]
        out["avail_actions"][b0:b1] = av.to(th.int32)
        out["actions"][b0:b1] = act[..., None]
        if with_onehot:
            oh = out["actions_onehot"][b0:b1]
            oh.scatter_(-1, act[..., None], 1.0)
            oh *= f[:, :, None, None]
    return out


def default_args(shape, mixer="qmix", **over):
    """SimpleNamespace with the key set the hot path reads (SURVEY.md section 5-config;
    values: reference alg yamls where they exist, else upstream PyMARL defaults)."""
    from types import SimpleNamespace
    s = get_shape(shape)
    d = dict(
        mac="basic_mac", agent="rnn", rnn_hidden_dim=64, obs_agent_id=True, obs_last_action=True,
        action_input_representation=None, agent_output_type="q", action_selector="epsilon_greedy",
        epsilon_start=1.0, epsilon_finish=0.05, epsilon_anneal_time=50000, learner="q_learner",
        double_q=True, mixer=mixer, mixing_embed_dim=32, target_update_interval=200,
        buffer_size=5000, batch_size=32, lr=5e-4, optim_alpha=0.99, optim_eps=1e-5, gamma=0.99,
        grad_norm_clip=10, learner_log_interval=10000, use_cuda=False, buffer_cpu_only=True,
        batch_size_run=1, device="cpu", n_agents=s.n_agents, n_actions=s.n_actions,
        state_shape=s.state_dim, obs_decoder=None, avail_actions_encoder=None, meta=None,
    )
    d.update(over)
    return SimpleNamespace(**d)
