"""CPU oracle for the PyMARL Q-learner hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy restatement (forward, hand-derived backward, clip, RMSprop, target sync, action
selection) of the reference algorithm.  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this module; the product
(pymarl_b200/) never does and has no CPU fallback.

Parity pin: the reference ships no tests or golden vectors for this path (SURVEY.md
section 4), so the oracle is pinned against the reference itself: tests/golden/make_golden.py
imports /root/reference/src in the build container, runs the unmodified torch modules
(QLearner.train, BasicMAC.forward/select_actions, QMixer, VDNMixer, EpsilonGreedy...) on
seeded synthetic batches in fp32 and fp64 and commits the inputs/outputs as fixtures under
tests/golden/; tests/test_oracle_golden.py checks every function below against them.

Every function cites the reference file:line it follows (paths relative to
/root/reference/src).  dtype follows the parameters' dtype (float32 or float64).
"""
import numpy as np

AGENT_PARAM_NAMES = ["fc1.weight", "fc1.bias", "rnn.weight_ih", "rnn.weight_hh",
                     "rnn.bias_ih", "rnn.bias_hh", "fc2.weight", "fc2.bias"]
QMIX_PARAM_NAMES = ["hyper_w_1.weight", "hyper_w_1.bias", "hyper_w_final.weight",
                    "hyper_w_final.bias", "hyper_b_1.weight", "hyper_b_1.bias",
                    "V.0.weight", "V.0.bias", "V.2.weight", "V.2.bias"]
MASK_VALUE = -9999999.0           # learners/q_learner.py:68,74


# ----------------------------------------------------------------------------------------
# parameter construction (shapes of modules/agents/rnn_agent.py:19-21, modules/mixers/qmix.py:17-26)
# ----------------------------------------------------------------------------------------
def agent_param_shapes(d_in, H, A):
    return {"fc1.weight": (H, d_in), "fc1.bias": (H,), "rnn.weight_ih": (3 * H, H),
            "rnn.weight_hh": (3 * H, H), "rnn.bias_ih": (3 * H,), "rnn.bias_hh": (3 * H,),
            "fc2.weight": (A, H), "fc2.bias": (A,)}


def qmix_param_shapes(S, N, E):
    return {"hyper_w_1.weight": (N * E, S), "hyper_w_1.bias": (N * E,),
            "hyper_w_final.weight": (E, S), "hyper_w_final.bias": (E,),
            "hyper_b_1.weight": (E, S), "hyper_b_1.bias": (E,),
            "V.0.weight": (E, S), "V.0.bias": (E,), "V.2.weight": (1, E), "V.2.bias": (1,)}


def init_params(shapes, rng, dtype=np.float32):
    """U(-1/sqrt(fan_in), +1/sqrt(fan_in)) like nn.Linear / nn.GRUCell (values are never
    compared against torch's init; parity runs copy the reference state_dict)."""
    out = {}
    for k, shp in shapes.items():
        base = k.rsplit(".", 1)[0]
        wshape = shapes[base + ".weight"] if (base + ".weight") in shapes else shapes[base + ".weight_ih"]
        bound = 1.0 / np.sqrt(wshape[0] // 3 if base == "rnn" else wshape[1])
        if base == "rnn":
            bound = 1.0 / np.sqrt(wshape[1])
        out[k] = rng.uniform(-bound, bound, size=shp).astype(dtype)
    return out


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


# ----------------------------------------------------------------------------------------
# agent forward
# ----------------------------------------------------------------------------------------
def build_inputs(batch, t, obs_last_action=True, obs_agent_id=True):
    """controllers/basic_controller.py:100-135 (obs_decoder None, action_input_representation None):
    cat(obs[:, t], onehot[:, t-1] (zeros at t == 0), eye(N)) reshaped to [B*N, D_in]."""
    obs = batch["obs"][:, t]
    B, N = obs.shape[:2]
    parts = [obs]
    if obs_last_action:
        oh = batch["actions_onehot"]
        parts.append(np.zeros_like(oh[:, t]) if t == 0 else oh[:, t - 1])
    if obs_agent_id:
        parts.append(np.broadcast_to(np.eye(N, dtype=obs.dtype)[None], (B, N, N)))
    x = np.concatenate([p.astype(obs.dtype) for p in parts], axis=-1)
    return x.reshape(B * N, -1)


def rnn_agent_forward(p, inputs, h_in, cache=None):
    """modules/agents/rnn_agent.py:27-36.  GRUCell gate order r, z, n and
    h' = n + z * (h - n) (torch's form).  Returns (q [R, A], h' [R, H])."""
    H = p["fc1.weight"].shape[0]
    pre1 = inputs @ p["fc1.weight"].T + p["fc1.bias"]
    x = np.maximum(pre1, 0)
    gi = x @ p["rnn.weight_ih"].T + p["rnn.bias_ih"]
    gh = h_in @ p["rnn.weight_hh"].T + p["rnn.bias_hh"]
    r = _sigmoid(gi[:, :H] + gh[:, :H])
    z = _sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
    n = np.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
    h = n + z * (h_in - n)
    q = h @ p["fc2.weight"].T + p["fc2.bias"]
    if cache is not None:
        cache.append(dict(inputs=inputs, x=x, h_in=h_in, r=r, z=z, n=n, ghn=gh[:, 2 * H:], h=h))
    return q, h


def mac_unroll(p, batch, obs_last_action=True, obs_agent_id=True, keep_cache=False):
    """learners/q_learner.py:47-52 (and 58-62 for the target net): init_hidden (zeros,
    basic_controller.py:77-81) then forward for t in 0..T-1, stacked to [B, T, N, A]."""
    B, T, N = batch["obs"].shape[:3]
    H = p["fc1.weight"].shape[0]
    dt = p["fc1.weight"].dtype
    h = np.zeros((B * N, H), dtype=dt)
    outs, cache = [], ([] if keep_cache else None)
    for t in range(T):
        q, h = rnn_agent_forward(p, build_inputs(batch, t, obs_last_action, obs_agent_id).astype(dt),
                                 h, cache)
        outs.append(q.reshape(B, N, -1))
    return np.stack(outs, axis=1), cache


# ----------------------------------------------------------------------------------------
# chosen-action gather, target masking, double-Q  (learners/q_learner.py:55-78)
# ----------------------------------------------------------------------------------------
def target_select(mac_out, target_mac_out_full, avail, actions, double_q=True, cur_max_override=None):
    """mac_out, target_mac_out_full: [B, T, N, A]; avail [B, T, N, A]; actions [B, T, N, 1].
    Returns chosen [B, T-1, N], target_max [B, T-1, N], cur_max_actions [B, T-1, N] (int64;
    argmax ties -> lowest index like torch CPU).  The target net's t == 0 output is dropped
    (:65); unavailable actions get -9999999 (:68, :74)."""
    chosen = np.take_along_axis(mac_out[:, :-1], actions[:, :-1], axis=3)[..., 0]
    tq = target_mac_out_full[:, 1:].copy()
    tq[avail[:, 1:] == 0] = MASK_VALUE
    if double_q:
        qd = mac_out.copy()
        qd[avail == 0] = MASK_VALUE
        cur_max = qd[:, 1:].argmax(axis=3)
        if cur_max_override is not None:          # referee runs: the arg-max decisions of another (lower precision) run
            cur_max = np.asarray(cur_max_override).astype(cur_max.dtype)
        tmax = np.take_along_axis(tq, cur_max[..., None], axis=3)[..., 0]
    else:
        cur_max = tq.argmax(axis=3)
        tmax = tq.max(axis=3)
    return chosen, tmax, cur_max.astype(np.int64)


# ----------------------------------------------------------------------------------------
# mixers
# ----------------------------------------------------------------------------------------
def qmixer_forward(mp, agent_qs, states, cache=None):
    """modules/mixers/qmix.py:28-47.  agent_qs [B, T', N], states [B, T', S] -> [B, T', 1]."""
    B = agent_qs.shape[0]
    N = agent_qs.shape[-1]
    E = mp["hyper_b_1.weight"].shape[0]
    s = states.reshape(-1, states.shape[-1])
    q = agent_qs.reshape(-1, N)
    raw_w1 = s @ mp["hyper_w_1.weight"].T + mp["hyper_w_1.bias"]
    w1 = np.abs(raw_w1).reshape(-1, N, E)
    b1 = s @ mp["hyper_b_1.weight"].T + mp["hyper_b_1.bias"]
    pre = np.einsum("mn,mne->me", q, w1) + b1
    hidden = np.where(pre > 0, pre, np.expm1(np.minimum(pre, 0)))      # F.elu, alpha = 1
    raw_wf = s @ mp["hyper_w_final.weight"].T + mp["hyper_w_final.bias"]
    wf = np.abs(raw_wf)
    v0pre = s @ mp["V.0.weight"].T + mp["V.0.bias"]
    v0 = np.maximum(v0pre, 0)
    v = v0 @ mp["V.2.weight"].T + mp["V.2.bias"]
    y = (hidden * wf).sum(-1, keepdims=True) + v
    if cache is not None:
        cache.update(s=s, q=q, raw_w1=raw_w1, w1=w1, pre=pre, hidden=hidden, raw_wf=raw_wf,
                     wf=wf, v0pre=v0pre, v0=v0)
    return y.reshape(B, -1, 1)


def qmixer_backward(mp, cache, g, decisions=None):
    """Gradient of qmixer_forward w.r.t. mixer parameters and agent_qs, given g = dL/dq_tot
    [M, 1] (what autograd does at learners/q_learner.py:101: d|x| = sign(x), dELU = 1 if
    pre > 0 else exp(pre), dReLU = (x > 0); no gradient flows to the state).
    `decisions` (referee runs only): the discontinuous derivative choices of another run - sign_w1 [M, N*E],
    sign_wf [M, E], v0_mask [M, E] - used instead of this run's own."""
    c = cache
    dec = decisions or {}
    N, E = c["w1"].shape[1:]
    g = g.reshape(-1, 1)
    d_hidden = g * c["wf"]
    d_raw_wf = dec.get("sign_wf", np.sign(c["raw_wf"])) * (g * c["hidden"])
    d_v0pre = dec.get("v0_mask", c["v0pre"] > 0) * (g * mp["V.2.weight"])
    d_pre = d_hidden * np.where(c["pre"] > 0, 1.0, np.exp(np.minimum(c["pre"], 0))).astype(g.dtype)
    d_raw_w1 = (dec.get("sign_w1", np.sign(c["raw_w1"])).reshape(-1, N, E) * c["q"][:, :, None] * d_pre[:, None, :]).reshape(-1, N * E)
    d_q = np.einsum("mne,me->mn", c["w1"], d_pre)
    s = c["s"]
    grads = {
        "hyper_w_1.weight": d_raw_w1.T @ s, "hyper_w_1.bias": d_raw_w1.sum(0),
        "hyper_w_final.weight": d_raw_wf.T @ s, "hyper_w_final.bias": d_raw_wf.sum(0),
        "hyper_b_1.weight": d_pre.T @ s, "hyper_b_1.bias": d_pre.sum(0),
        "V.0.weight": d_v0pre.T @ s, "V.0.bias": d_v0pre.sum(0),
        "V.2.weight": (g * c["v0"]).sum(0, keepdims=True), "V.2.bias": g.sum(0),
    }
    return grads, d_q


def vdn_forward(agent_qs):
    """modules/mixers/vdn.py:9-10."""
    return agent_qs.sum(axis=2, keepdims=True)


# ----------------------------------------------------------------------------------------
# agent backward (BPTT) - what loss.backward() does through learners/q_learner.py:47-55
# ----------------------------------------------------------------------------------------
def agent_bptt(p, cache, actions, d_chosen, B, N, relu_mask=None):
    """d_chosen [B, T-1, N] = dL/d chosen_action_qvals.  Gradient reaches mac_out only at
    the taken action for t < T-1 (gather, :55); the double-Q argmax path is detached (:73).
    Returns grads for the 8 agent tensors.  `relu_mask` [T, B*N, H] (referee runs only): the ReLU
    decisions (x > 0) of another run, used instead of this run's own."""
    T = len(cache)
    H = p["fc1.weight"].shape[0]
    A = p["fc2.weight"].shape[0]
    dt = p["fc1.weight"].dtype
    g = {k: np.zeros_like(v) for k, v in p.items()}
    dh_next = np.zeros((B * N, H), dtype=dt)
    for t in range(T - 1, -1, -1):
        c = cache[t]
        dq = np.zeros((B * N, A), dtype=dt)
        if t < T - 1:
            a = actions[:, t].reshape(B * N)
            dq[np.arange(B * N), a] = d_chosen[:, t].reshape(B * N)
        g["fc2.weight"] += dq.T @ c["h"]
        g["fc2.bias"] += dq.sum(0)
        dh = dh_next + dq @ p["fc2.weight"]
        r, z, n, ghn, h_in = c["r"], c["z"], c["n"], c["ghn"], c["h_in"]
        dn = dh * (1 - z)
        dz = dh * (h_in - n)
        da_n = dn * (1 - n * n)
        da_r = (da_n * ghn) * r * (1 - r)
        da_z = dz * z * (1 - z)
        dgi = np.concatenate([da_r, da_z, da_n], axis=1)
        dgh = np.concatenate([da_r, da_z, da_n * r], axis=1)
        g["rnn.weight_ih"] += dgi.T @ c["x"]
        g["rnn.bias_ih"] += dgi.sum(0)
        g["rnn.weight_hh"] += dgh.T @ h_in
        g["rnn.bias_hh"] += dgh.sum(0)
        dx = dgi @ p["rnn.weight_ih"]
        dh_next = dh * z + dgh @ p["rnn.weight_hh"]
        dpre1 = dx * ((c["x"] > 0) if relu_mask is None else relu_mask[t])
        g["fc1.weight"] += dpre1.T @ c["inputs"]
        g["fc1.bias"] += dpre1.sum(0)
    return g


# ----------------------------------------------------------------------------------------
# optimiser
# ----------------------------------------------------------------------------------------
def clip_grad_norm(grads, names, max_norm):
    """torch.nn.utils.clip_grad_norm_ (learners/q_learner.py:102): total = || [||g_i||] ||_2,
    coef = min(1, max_norm / (total + 1e-6)), g_i *= coef.  Returns total (pre-clip)."""
    dt = grads[names[0]].dtype
    norms = np.array([np.sqrt((grads[k].astype(dt) ** 2).sum(dtype=dt)) for k in names], dtype=dt)
    total = np.sqrt((norms ** 2).sum(dtype=dt))
    coef = dt.type(max_norm) / (total + dt.type(1e-6))
    coef = min(dt.type(1.0), coef)
    for k in names:
        grads[k] *= coef
    return total


def rmsprop_step(params, grads, square_avg, names, lr, alpha, eps):
    """torch.optim.RMSprop defaults (momentum 0, centered False, weight_decay 0), q_learner.py:30,103:
    v = alpha v + (1 - alpha) g^2 ;  p -= lr * g / (sqrt(v) + eps)."""
    for k in names:
        dt = params[k].dtype.type
        gk = grads[k]
        square_avg[k] *= dt(alpha)
        square_avg[k] += dt(1 - alpha) * gk * gk
        params[k] -= dt(lr) * gk / (np.sqrt(square_avg[k]) + dt(eps))


# ----------------------------------------------------------------------------------------
# the learner
# ----------------------------------------------------------------------------------------
class OracleQLearner:
    """learners/q_learner.py:9-122 with explicit state.  ``mixer`` in {"qmix", "vdn", None}."""

    def __init__(self, agent_params, mixer_params, args):
        self.args = args
        self.mixer = args.mixer
        if self.mixer not in ("qmix", "vdn", None):
            raise ValueError("Mixer {} not recognised.".format(self.mixer))     # :26
        self.agent = {k: v.copy() for k, v in agent_params.items()}
        self.mixer_p = {k: v.copy() for k, v in (mixer_params or {}).items()} if self.mixer == "qmix" else {}
        self.target_agent = {k: v.copy() for k, v in self.agent.items()}          # deepcopy(mac) :33
        self.target_mixer_p = {k: v.copy() for k, v in self.mixer_p.items()}      # :27
        self.sq_agent = {k: np.zeros_like(v) for k, v in self.agent.items()}
        self.sq_mixer = {k: np.zeros_like(v) for k, v in self.mixer_p.items()}
        self.last_target_update_episode = 0
        self.log_stats_t = -args.learner_log_interval - 1
        self.stats = {}
        self.n_target_updates = 0

    def forward_loss(self, batch, keep_cache=True, cur_max_override=None):
        """q_learner.py:39-97.  Returns a dict with every intermediate the tests compare."""
        a = self.args
        dt = self.agent["fc1.weight"].dtype
        rewards = batch["reward"][:, :-1].astype(dt)
        actions = batch["actions"][:, :-1]
        terminated = batch["terminated"][:, :-1].astype(dt)
        mask = batch["filled"][:, :-1].astype(dt)
        mask[:, 1:] = mask[:, 1:] * (1 - terminated[:, :-1])
        avail = batch["avail_actions"]
        ola, oid = getattr(a, "obs_last_action", True), getattr(a, "obs_agent_id", True)

        mac_out, cache = mac_unroll(self.agent, batch, ola, oid, keep_cache)
        target_out, _ = mac_unroll(self.target_agent, batch, ola, oid, False)
        chosen, tmax, cur_max = target_select(mac_out, target_out, avail, batch["actions"], a.double_q, cur_max_override)

        mcache = {}
        if self.mixer == "qmix":
            q_tot = qmixer_forward(self.mixer_p, chosen, batch["state"][:, :-1].astype(dt), mcache)
            t_tot = qmixer_forward(self.target_mixer_p, tmax, batch["state"][:, 1:].astype(dt))
        elif self.mixer == "vdn":
            q_tot, t_tot = vdn_forward(chosen), vdn_forward(tmax)
        else:
            q_tot, t_tot = chosen, tmax
        # gamma * (1 - terminated) is a python float times a float32 tensor in the reference
        # (terminated = ....float(), :41) -> gamma is rounded to float32 whatever the net dtype
        targets = rewards + (np.float32(a.gamma) * (1 - terminated.astype(np.float32))).astype(dt) * t_tot
        td = q_tot - targets
        mask_e = np.broadcast_to(mask, td.shape)
        masked_td = td * mask_e
        mask_sum = mask_e.sum(dtype=dt)
        loss = (masked_td ** 2).sum(dtype=dt) / mask_sum
        return dict(mac_out=mac_out, target_mac_out=target_out, chosen=chosen, target_max=tmax,
                    cur_max_actions=cur_max, q_tot=q_tot, target_tot=t_tot, targets=targets,
                    td=td, mask=mask_e, masked_td=masked_td, mask_sum=mask_sum, loss=loss,
                    cache=cache, mcache=mcache, actions=actions)

    def backward(self, fw, batch, decisions=None):
        """loss.backward() (q_learner.py:100-101) by hand.  `decisions`: see qmixer_backward / agent_bptt."""
        dec = decisions or {}
        B, T, N = batch["obs"].shape[:3]
        g_tot = 2.0 * fw["masked_td"] * fw["mask"] / fw["mask_sum"]        # dL/dq_tot
        grads = {}
        if self.mixer == "qmix":
            mg, d_q = qmixer_backward(self.mixer_p, fw["mcache"], g_tot.reshape(-1, 1), dec)
            grads.update({"mixer." + k: v for k, v in mg.items()})
            d_chosen = d_q.reshape(B, T - 1, N)
        elif self.mixer == "vdn":
            d_chosen = np.broadcast_to(g_tot, (B, T - 1, N)).copy()
        else:
            d_chosen = g_tot
        ag = agent_bptt(self.agent, fw["cache"], batch["actions"], d_chosen.astype(g_tot.dtype), B, N, dec.get("relu_mask"))
        grads.update({"agent." + k: v for k, v in ag.items()})
        return grads

    def train(self, batch, t_env, episode_num, decisions=None):
        """q_learner.py:37-116.  Returns the 5 logged scalars (always, not only when the
        log interval fires) plus the gradients for inspection.  `decisions` (referee runs only): the
        discontinuous choices of another run (cur_max, relu_mask, sign_w1, sign_wf, v0_mask)."""
        a = self.args
        dec = decisions or {}
        fw = self.forward_loss(batch, cur_max_override=dec.get("cur_max"))
        grads = self.backward(fw, batch, dec)
        names = ["agent." + k for k in AGENT_PARAM_NAMES]
        if self.mixer == "qmix":
            names += ["mixer." + k for k in QMIX_PARAM_NAMES]
        raw_grads = {k: v.copy() for k, v in grads.items()}
        grad_norm = clip_grad_norm(grads, names, a.grad_norm_clip)
        params = {"agent." + k: v for k, v in self.agent.items()}
        params.update({"mixer." + k: v for k, v in self.mixer_p.items()})
        sq = {"agent." + k: v for k, v in self.sq_agent.items()}
        sq.update({"mixer." + k: v for k, v in self.sq_mixer.items()})
        rmsprop_step(params, grads, sq, names, a.lr, a.optim_alpha, a.optim_eps)

        if (episode_num - self.last_target_update_episode) / a.target_update_interval >= 1.0:   # :105
            self._update_targets()
            self.last_target_update_episode = episode_num
        mask_elems = float(fw["mask_sum"])
        n_agents = batch["obs"].shape[2]
        self.stats = dict(
            loss=float(fw["loss"]), grad_norm=float(grad_norm),
            td_error_abs=float(np.abs(fw["masked_td"]).sum()) / mask_elems,
            q_taken_mean=float((fw["q_tot"] * fw["mask"]).sum()) / (mask_elems * n_agents),
            target_mean=float((fw["targets"] * fw["mask"]).sum()) / (mask_elems * n_agents))
        if t_env - self.log_stats_t >= a.learner_log_interval:
            self.log_stats_t = t_env
        return self.stats, raw_grads, fw

    def _update_targets(self):
        """q_learner.py:118-122."""
        for k, v in self.agent.items():
            self.target_agent[k][...] = v
        for k, v in self.mixer_p.items():
            self.target_mixer_p[k][...] = v
        self.n_target_updates += 1


# ----------------------------------------------------------------------------------------
# rollout side
# ----------------------------------------------------------------------------------------
def epsilon_schedule(t_env, start=1.0, finish=0.05, anneal_time=50000):
    """components/epsilon_schedules.py:12-23 (decay='linear')."""
    delta = (start - finish) / anneal_time
    return max(finish, start - delta * t_env)


def select_action(q, avail, epsilon, u, expo):
    """components/action_selectors.py:44-62 with the random draws injected.
    u [b, N] are the th.rand_like draws (:57); expo [b, N, A] the Exp(1) draws that
    Categorical(avail.float()).sample() consumes (torch multinomial, 1 sample:
    argmax(probs / expo), probs = avail / avail.sum(-1)).  float32 arithmetic throughout so
    the integer result is bit-exact.  Returns int64 [b, N]."""
    q = np.asarray(q, dtype=np.float32)
    masked = np.where(avail == 0, -np.inf, q).astype(np.float32)
    greedy = masked.argmax(axis=2)
    af = avail.astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        probs = af / af.sum(-1, keepdims=True, dtype=np.float32)
        ratio = (probs / np.asarray(expo, dtype=np.float32)).astype(np.float32)
    random_actions = ratio.argmax(axis=2)
    pick_random = (np.asarray(u, dtype=np.float32) < np.float32(epsilon)).astype(np.int64)
    return pick_random * random_actions + (1 - pick_random) * greedy


def mac_select_actions(p, batch, t_ep, h, epsilon, u, expo, obs_last_action=True, obs_agent_id=True):
    """controllers/basic_controller.py:30-38: one agent step for every env, then epsilon-greedy.
    Returns (actions [B, N] int64, q [B, N, A], h')."""
    B, N = batch["obs"].shape[0], batch["obs"].shape[2]
    dt = p["fc1.weight"].dtype
    q, h2 = rnn_agent_forward(p, build_inputs(batch, t_ep, obs_last_action, obs_agent_id).astype(dt), h)
    q = q.reshape(B, N, -1)
    return select_action(q, batch["avail_actions"][:, t_ep], epsilon, u, expo), q, h2


def replay_sample_ids(episodes_in_buffer, batch_size, seed):
    """components/episode_buffer.py:291-298: np.random.choice(n, k, replace=False) on the
    legacy global RandomState (seeded here explicitly)."""
    if episodes_in_buffer == batch_size:
        return np.arange(batch_size)
    rs = np.random.RandomState(seed)
    return rs.choice(episodes_in_buffer, batch_size, replace=False)


def max_t_filled(filled):
    """components/episode_buffer.py:255-256."""
    return int(filled.sum(axis=1).max())
