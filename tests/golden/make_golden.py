"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(/root/reference/src, torch CPU) on seeded synthetic batches.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Each ``<case>.npz`` holds the inputs (batch fields, initial online/target parameters), and
the reference's outputs in fp32 (`f32/...`) and fp64 (`f64/...`, the noise referee):
mac_out, target_mac_out, chosen, target_max, cur_max_actions, q_tot, targets, loss, raw
gradients, grad_norm, the 5 logged scalars, and the parameters / RMSprop square_avg /
target parameters after every one of K train steps (the schedule includes a target sync).
``select_actions.npz`` holds epsilon-greedy cases with the torch generator draws.
"""
import copy
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch as th

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
REF_SRC = os.environ.get("PYMARL_REF_SRC", "/root/reference/src")
sys.path.insert(0, REF_SRC)

from pymarl_b200.synthetic import SmacShape, numpy_episode_fields, default_args  # noqa: E402

# reference modules (top-level absolute imports)
from components.episode_buffer import EpisodeBatch as RefEpisodeBatch  # noqa: E402
from components.transforms import OneHot as RefOneHot  # noqa: E402
from components.action_selectors import REGISTRY as ref_action_REGISTRY  # noqa: E402
from controllers import REGISTRY as ref_mac_REGISTRY  # noqa: E402
from learners import REGISTRY as ref_le_REGISTRY  # noqa: E402


class _Logger:
    def __init__(self):
        self.stats = {}
        self.console_logger = SimpleNamespace(info=lambda *a, **k: None)

    def log_stat(self, key, value, t):
        self.stats.setdefault(key, []).append((t, float(value)))


def ref_scheme(shape, th_float=th.float32):
    scheme = {
        "state": {"vshape": shape.state_dim, "dtype": th_float},
        "obs": {"vshape": shape.obs_dim, "group": "agents", "vshape_decoded": shape.obs_dim, "dtype": th_float},
        "actions": {"vshape": (1,), "group": "agents", "dtype": th.long},
        "avail_actions": {"vshape": (shape.n_actions,), "group": "agents", "dtype": th.int},
        "reward": {"vshape": (1,), "dtype": th_float},
        "terminated": {"vshape": (1,), "dtype": th.uint8},
    }
    groups = {"agents": shape.n_agents}
    preprocess = {"actions": ("actions_onehot", [RefOneHot(out_dim=shape.n_actions)])}
    return scheme, groups, preprocess


def ref_batch(shape, fields, th_float=th.float32):
    B, T = fields["obs"].shape[:2]
    scheme, groups, preprocess = ref_scheme(shape, th_float)
    batch = RefEpisodeBatch(scheme, groups, B, T, preprocess=preprocess, device="cpu")
    for k, v in fields.items():
        t = th.from_numpy(np.ascontiguousarray(v))
        if t.is_floating_point():
            t = t.to(th_float)
        assert batch.data.transition_data[k].shape == t.shape, (k, t.shape)
        batch.data.transition_data[k] = t
    return batch, scheme, groups


def build_ref_learner(shape, args, th_float, seed):
    th.manual_seed(seed)
    scheme, groups, _ = ref_scheme(shape, th_float)
    # BasicMAC reads the post-_setup_data scheme (with actions_onehot)
    scheme = dict(scheme)
    scheme["actions_onehot"] = {"vshape": (shape.n_actions,), "dtype": th.float32, "group": "agents"}
    mac = ref_mac_REGISTRY[args.mac](scheme, groups, args)
    logger = _Logger()
    learner = ref_le_REGISTRY[args.learner](mac, scheme, logger, args)
    # make the target nets differ from the online nets so target handling is exercised
    g = th.Generator().manual_seed(seed + 1)
    with th.no_grad():
        for p in learner.target_mac.parameters():
            p.add_(0.1 * th.randn(p.shape, generator=g))
        if learner.mixer is not None and hasattr(learner, "target_mixer"):
            for p in learner.target_mixer.parameters():
                p.add_(0.05 * th.randn(p.shape, generator=g))
    return learner, logger


def state_np(module):
    return {k: v.detach().cpu().numpy().copy() for k, v in module.state_dict().items()}


def instrumented_forward(learner, batch):
    """Re-run the forward part of QLearner.train (q_learner.py:39-97) with the reference's
    own modules to expose intermediates (the train() call itself is still the unmodified
    method; this only reads)."""
    a = learner.args
    with th.no_grad():
        rewards = batch["reward"][:, :-1]
        actions = batch["actions"][:, :-1]
        terminated = batch["terminated"][:, :-1].float()
        mask = batch["filled"][:, :-1].float()
        mask[:, 1:] = mask[:, 1:] * (1 - terminated[:, :-1])
        avail = batch["avail_actions"]
        mac_out = []
        learner.mac.init_hidden(batch.batch_size)
        for t in range(batch.max_seq_length):
            mac_out.append(learner.mac.forward(batch, t=t))
        mac_out = th.stack(mac_out, dim=1)
        chosen = th.gather(mac_out[:, :-1], dim=3, index=actions).squeeze(3)
        tout = []
        learner.target_mac.init_hidden(batch.batch_size)
        for t in range(batch.max_seq_length):
            tout.append(learner.target_mac.forward(batch, t=t))
        target_full = th.stack(tout, dim=1)
        tmo = target_full[:, 1:].clone()
        tmo[avail[:, 1:] == 0] = -9999999
        if a.double_q:
            mod = mac_out.clone()
            mod[avail == 0] = -9999999
            cur_max = mod[:, 1:].max(dim=3, keepdim=True)[1]
            tmax = th.gather(tmo, 3, cur_max).squeeze(3)
            cur_max = cur_max.squeeze(3)
        else:
            tmax, cur_max = tmo.max(dim=3)
        if learner.mixer is not None:
            q_tot = learner.mixer(chosen, batch["state"][:, :-1])
            t_tot = learner.target_mixer(tmax, batch["state"][:, 1:])
        else:
            q_tot, t_tot = chosen, tmax
        targets = rewards + a.gamma * (1 - terminated) * t_tot
    return dict(mac_out=mac_out, target_mac_out=target_full, chosen=chosen, target_max=tmax,
                cur_max_actions=cur_max, q_tot=q_tot, target_tot=t_tot, targets=targets)


def run_case(name, shape, B, T, mixer, double_q, seed, n_steps=3, sync_step=1, **over):
    out = {}
    fields = numpy_episode_fields(shape, B, T, seed=seed, ragged=True)
    for k, v in fields.items():
        out["in/" + k] = v
    meta = dict(name=name, shape=tuple(shape), B=B, T=T, mixer=mixer, double_q=double_q,
                seed=seed, n_steps=n_steps, sync_step=sync_step, over=over)
    for tag, th_float in (("f32", th.float32), ("f64", th.float64)):
        args = default_args(shape, mixer=mixer, double_q=double_q, learner_log_interval=0,
                            target_update_interval=200, **over)
        learner, logger = build_ref_learner(shape, args, th.float32, seed)
        if th_float == th.float64:
            learner.mac.agent.double()
            learner.target_mac.agent.double()
            if learner.mixer is not None:
                learner.mixer.double()
                learner.target_mixer.double()
        batch, _, _ = ref_batch(shape, fields, th_float)
        if tag == "f32":
            for k, v in state_np(learner.mac.agent).items():
                out["init/agent/" + k] = v
            for k, v in state_np(learner.target_mac.agent).items():
                out["init/target_agent/" + k] = v
            if mixer == "qmix":
                for k, v in state_np(learner.mixer).items():
                    out["init/mixer/" + k] = v
                for k, v in state_np(learner.target_mixer).items():
                    out["init/target_mixer/" + k] = v
        fw = instrumented_forward(learner, batch)
        for k, v in fw.items():
            out["%s/fw/%s" % (tag, k)] = v.numpy()
        # raw gradients of step 0: run train on a deep copy with a huge clip so .grad is unclipped
        probe = copy.deepcopy(learner)
        probe.args = copy.copy(args)
        probe.args.grad_norm_clip = 1e30
        probe.logger = _Logger()
        probe.train(batch, 0, 0)
        names = [k for k, _ in probe.mac.agent.named_parameters()]
        for k, p in probe.mac.agent.named_parameters():
            out["%s/grad/agent/%s" % (tag, k)] = p.grad.numpy().copy()
        if mixer == "qmix":
            for k, p in probe.mixer.named_parameters():
                out["%s/grad/mixer/%s" % (tag, k)] = p.grad.numpy().copy()
        del names
        # K real train steps; episode_num crosses target_update_interval at `sync_step`
        for step in range(n_steps):
            episode_num = 200 if step == sync_step else (201 if step > sync_step else 0)
            learner.train(batch, step, episode_num)
            # parameters: fp32 after the first and the last step, fp64 (referee) after the last;
            # target nets right after the sync step and after the last step
            keep_online = (step == n_steps - 1) or (tag == "f32" and step == 0)
            keep_target = (step == n_steps - 1) or (tag == "f32" and step == sync_step)
            if keep_online:
                for k, v in state_np(learner.mac.agent).items():
                    out["%s/step%d/agent/%s" % (tag, step, k)] = v
                if mixer == "qmix":
                    for k, v in state_np(learner.mixer).items():
                        out["%s/step%d/mixer/%s" % (tag, step, k)] = v
                sq = [learner.optimiser.state[p]["square_avg"].numpy().copy() for p in learner.params]
                out["%s/step%d/square_avg_flat" % (tag, step)] = np.concatenate([s.ravel() for s in sq])
            if keep_target:
                for k, v in state_np(learner.target_mac.agent).items():
                    out["%s/step%d/target_agent/%s" % (tag, step, k)] = v
                if mixer == "qmix":
                    for k, v in state_np(learner.target_mixer).items():
                        out["%s/step%d/target_mixer/%s" % (tag, step, k)] = v
            for key in ("loss", "grad_norm", "td_error_abs", "q_taken_mean", "target_mean"):
                out["%s/step%d/stat/%s" % (tag, step, key)] = np.float64(logger.stats[key][-1][1])
    out["meta"] = np.array(repr(meta))
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


def run_select_actions():
    """EpsilonGreedyActionSelector.select_action (action_selectors.py:44-62) and
    BasicMAC.select_actions (basic_controller.py:30-38) with the torch CPU generator: the
    draws are th.rand_like first, then the exponential_ inside multinomial."""
    out = {}
    shape = SmacShape("tiny", 4, 17, 23, 7, 9)
    rng = np.random.default_rng(5)
    cases = []
    for ci, (b, eps_t_env, test_mode) in enumerate([(6, 0, False), (6, 30000, False), (6, 10 ** 7, False),
                                                    (6, 0, True), (33, 25000, False)]):
        N, A = shape.n_agents, shape.n_actions
        q = rng.standard_normal((b, N, A)).astype(np.float32)
        q[0, 0, :] = 0.5                                    # exact ties -> first index
        avail = (rng.random((b, N, A)) < 0.5).astype(np.int32)
        avail[..., rng.integers(0, A)] = 1
        args = default_args(shape)
        sel = ref_action_REGISTRY["epsilon_greedy"](args)
        th.manual_seed(100 + ci)
        acts = sel.select_action(th.from_numpy(q), th.from_numpy(avail), eps_t_env, test_mode=test_mode)
        th.manual_seed(100 + ci)
        u = th.rand_like(th.from_numpy(q)[:, :, 0])
        expo = th.empty(b * N, A).exponential_()
        out["sel%d/q" % ci] = q
        out["sel%d/avail" % ci] = avail
        out["sel%d/u" % ci] = u.numpy()
        out["sel%d/expo" % ci] = expo.numpy().reshape(b, N, A)
        out["sel%d/epsilon" % ci] = np.float64(sel.epsilon)
        out["sel%d/t_env" % ci] = np.int64(eps_t_env)
        out["sel%d/test_mode" % ci] = np.bool_(test_mode)
        out["sel%d/actions" % ci] = acts.numpy()
        cases.append(ci)
    out["n_sel"] = np.int64(len(cases))

    # full MAC step through the reference BasicMAC, two consecutive timesteps
    B, T = 5, 4
    fields = numpy_episode_fields(shape, B, T, seed=11, ragged=False)
    args = default_args(shape)
    scheme, groups, _ = ref_scheme(shape)
    scheme = dict(scheme)
    scheme["actions_onehot"] = {"vshape": (shape.n_actions,), "dtype": th.float32, "group": "agents"}
    th.manual_seed(3)
    mac = ref_mac_REGISTRY["basic_mac"](scheme, groups, args)
    batch, _, _ = ref_batch(shape, fields)
    for k, v in fields.items():
        out["mac/in/" + k] = v
    for k, v in state_np(mac.agent).items():
        out["mac/agent/" + k] = v
    mac.init_hidden(B)
    for t in range(3):
        th.manual_seed(200 + t)
        acts = mac.select_actions(batch, t_ep=t, t_env=20000, bs=slice(None), test_mode=False)
        th.manual_seed(200 + t)
        u = th.rand(B, shape.n_agents)
        expo = th.empty(B * shape.n_agents, shape.n_actions).exponential_()
        out["mac/t%d/actions" % t] = acts.numpy()
        out["mac/t%d/u" % t] = u.numpy()
        out["mac/t%d/expo" % t] = expo.numpy().reshape(B, shape.n_agents, shape.n_actions)
        out["mac/t%d/hidden" % t] = mac.hidden_states.detach().numpy().reshape(B * shape.n_agents, -1).copy()
        out["mac/t%d/epsilon" % t] = np.float64(mac.action_selector.epsilon)
    path = os.path.join(HERE, "select_actions.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


def run_replay_sample():
    """ReplayBuffer.sample ids (episode_buffer.py:291-298) under np.random.seed."""
    from components.episode_buffer import ReplayBuffer as RefReplayBuffer
    shape = SmacShape("tiny", 2, 5, 6, 4, 5)
    scheme, groups, preprocess = ref_scheme(shape)
    buf = RefReplayBuffer(scheme, groups, 16, shape.max_seq_length, preprocess=preprocess, device="cpu")
    fields = numpy_episode_fields(shape, 12, shape.max_seq_length, seed=2, ragged=True)
    eb, _, _ = ref_batch(shape, fields)
    buf.insert_episode_batch(eb)
    out = {"in/" + k: v for k, v in fields.items()}
    for seed in (0, 1, 7):
        np.random.seed(seed)
        s = buf.sample(5)
        out["seed%d/obs" % seed] = s["obs"].numpy()
        out["seed%d/filled" % seed] = s["filled"].numpy()
        out["seed%d/max_t_filled" % seed] = np.int64(int(s.max_t_filled()))
    path = os.path.join(HERE, "replay_sample.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))



def run_checkpoint():
    """Checkpoint files written by the REFERENCE (learner.save_models -> agent.th / mixer.th / opt.th,
    q_learner.py:124-135) after two train steps, plus what a second reference learner holds after
    load_models (q_learner.py:137-143: the target MAC loads the online weights, the target mixer is left
    alone) and after one more train step."""
    tiny = SmacShape("tiny", 3, 10, 14, 5, 8)
    args = default_args(tiny, mixer="qmix", double_q=True, learner_log_interval=0, rnn_hidden_dim=16, mixing_embed_dim=8)
    fields = numpy_episode_fields(tiny, 4, 8, seed=21, ragged=True)
    batch, _, _ = ref_batch(tiny, fields)
    out = {"in/" + k: v for k, v in fields.items()}
    writer, _ = build_ref_learner(tiny, args, th.float32, seed=31)
    writer.train(batch, 0, 0)
    writer.train(batch, 1, 0)
    ckpt = os.path.join(HERE, "ckpt_ref")
    os.makedirs(ckpt, exist_ok=True)
    writer.save_models(ckpt)
    reader, logger = build_ref_learner(tiny, copy.copy(args), th.float32, seed=32)
    for tag, mod in (("agent", reader.mac.agent), ("target_agent", reader.target_mac.agent), ("mixer", reader.mixer),
                     ("target_mixer", reader.target_mixer)):
        for k, v in state_np(mod).items():
            out["init/%s/%s" % (tag, k)] = v
    reader.load_models(ckpt)
    for tag, mod in (("agent", reader.mac.agent), ("target_agent", reader.target_mac.agent), ("mixer", reader.mixer),
                     ("target_mixer", reader.target_mixer)):
        for k, v in state_np(mod).items():
            out["loaded/%s/%s" % (tag, k)] = v
    sq = [reader.optimiser.state[p]["square_avg"].numpy().copy() for p in reader.params]
    out["loaded/square_avg_flat"] = np.concatenate([s_.ravel() for s_ in sq])
    reader.train(batch, 2, 0)
    for tag, mod in (("agent", reader.mac.agent), ("mixer", reader.mixer)):
        for k, v in state_np(mod).items():
            out["after/%s/%s" % (tag, k)] = v
    for key in ("loss", "grad_norm", "td_error_abs", "q_taken_mean", "target_mean"):
        out["after/stat/" + key] = np.float64(logger.stats[key][-1][1])
    sq = [reader.optimiser.state[p]["square_avg"].numpy().copy() for p in reader.params]
    out["after/square_avg_flat"] = np.concatenate([s_.ravel() for s_ in sq])
    out["meta"] = np.array(repr(dict(name="checkpoint", shape=tuple(tiny), mixer="qmix", double_q=True,
                                     over=dict(rnn_hidden_dim=16, mixing_embed_dim=8))))
    path = os.path.join(HERE, "checkpoint.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024), "+", sorted(os.listdir(ckpt)))


def run_episode_update():
    """EpisodeBatch.update / ReplayBuffer.insert_episode_batch (episode_buffer.py:98-154, 271-286) with the OneHot
    preprocess (transforms.py:12-21), driven the way the runners drive them (parallel_runner.py:100-204): per timestep a
    pre-transition update of state / avail_actions / obs for the live envs, then actions / reward / terminated with
    mark_filled=False; list-valued `bs`, scalar `ts`; finally two inserts into a ring buffer that wrap around."""
    from components.episode_buffer import ReplayBuffer as RefReplayBuffer
    shape = SmacShape("tiny", 3, 6, 7, 5, 6)
    scheme, groups, preprocess = ref_scheme(shape)
    B, T = 5, shape.max_seq_length
    rng = np.random.default_rng(17)
    eb = RefEpisodeBatch(scheme, groups, B, T, preprocess=preprocess, device="cpu")
    out, calls = {}, []

    def do(data, bs, ts, mark_filled):
        i = len(calls)
        for k, v in data.items():
            out["call%d/%s" % (i, k)] = np.asarray(v)
        out["call%d/bs" % i] = np.asarray(bs, dtype=np.int64)
        calls.append(dict(ts=int(ts), mark_filled=bool(mark_filled), keys=sorted(data)))
        eb.update({k: np.asarray(v) for k, v in data.items()}, bs=list(bs), ts=ts, mark_filled=mark_filled)

    live = list(range(B))
    for t in range(T - 1):
        n = len(live)
        avail = (rng.random((n, shape.n_agents, shape.n_actions)) < 0.6).astype(np.int32)
        avail[..., 0] = 1
        do({"state": rng.standard_normal((n, shape.state_dim)).astype(np.float32), "avail_actions": avail,
            "obs": rng.standard_normal((n, shape.n_agents, shape.obs_dim)).astype(np.float32)}, live, t, True)
        acts = (rng.random((n, shape.n_agents, shape.n_actions)) * avail).argmax(-1).astype(np.int64)
        term = (rng.random(n) < 0.25)
        do({"actions": acts[:, :, None], "reward": rng.standard_normal((n, 1)).astype(np.float32),
            "terminated": term[:, None].astype(np.uint8)}, live, t, False)
        live = [e for e, d in zip(live, term) if not d]
        if not live:
            break
    for k, v in eb.data.transition_data.items():
        out["final/" + k] = v.numpy()
    buf = RefReplayBuffer(scheme, groups, 8, T, preprocess=preprocess, device="cpu")
    buf.insert_episode_batch(eb)
    buf.insert_episode_batch(eb)                       # 5 + 5 into 8 slots: wraps
    for k, v in buf.data.transition_data.items():
        out["buffer/" + k] = v.numpy()
    out["buffer/index"] = np.int64(buf.buffer_index)
    out["buffer/episodes"] = np.int64(buf.episodes_in_buffer)
    out["meta"] = np.array(repr(dict(shape=tuple(shape), B=B, T=T, calls=calls)))
    path = os.path.join(HERE, "episode_update.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


def run_coma(name="coma_tiny", seed=41, n_steps=2, **over):
    """The reference COMALearner (learners/coma_learner.py) + COMACritic + BasicMAC with agent_output_type "pi_logits" +
    MultinomialActionSelector on a tiny seeded batch: inputs, initial parameters, the pieces that can be read without
    touching train() (target-critic Q, td-lambda targets, the MAC's policy output per timestep, sampled actions with the
    generator draws), and after each of `n_steps` train() calls the logged statistics and every parameter / RMSprop state."""
    from utils.rl_utils import build_td_lambda_targets
    tiny = SmacShape("tiny", 3, 10, 14, 5, 8)
    kw = dict(rnn_hidden_dim=16, agent_output_type="pi_logits", action_selector="multinomial", learner="coma_learner",
              critic_lr=5e-4, td_lambda=0.8, mask_before_softmax=True, epsilon_start=0.5, epsilon_finish=0.01,
              epsilon_anneal_time=100000, target_update_interval=3, learner_log_interval=0, test_greedy=True)
    kw.update(over)
    args = default_args(tiny, mixer=None, **kw)
    B, T = 4, 8
    fields = numpy_episode_fields(tiny, B, T, seed=seed, ragged=True)
    batch, scheme, groups = ref_batch(tiny, fields)
    out = {"in/" + k: v for k, v in fields.items()}
    th.manual_seed(seed)
    # the fork's BasicMAC turns scheme["obs"]["vshape"] into a tuple in place (basic_controller.py:139-144), which
    # COMACritic._get_input_shape (coma.py:52-59) cannot add to an int: give each its own copy of the scheme
    sch = copy.deepcopy(dict(batch.scheme))
    mac = ref_mac_REGISTRY[args.mac](copy.deepcopy(sch), groups, args)
    logger = _Logger()
    learner = ref_le_REGISTRY["coma_learner"](mac, sch, logger, args)
    g = th.Generator().manual_seed(seed + 1)
    with th.no_grad():
        for p_ in learner.target_critic.parameters():
            p_.add_(0.05 * th.randn(p_.shape, generator=g))
    for tag, mod in (("agent", mac.agent), ("critic", learner.critic), ("target_critic", learner.target_critic)):
        for k, v in state_np(mod).items():
            out["init/%s/%s" % (tag, k)] = v
    eps = float(mac.action_selector.epsilon)
    out["epsilon"] = np.float64(eps)
    with th.no_grad():
        tq = learner.target_critic(batch)
        out["fw/target_q"] = tq.numpy()
        out["fw/critic_inputs"] = learner.critic._build_inputs(batch).numpy()
        rewards = batch["reward"][:, :-1]
        terminated = batch["terminated"][:, :-1].float()
        mask = batch["filled"][:, :-1].float()
        mask[:, 1:] = mask[:, 1:] * (1 - terminated[:, :-1])
        taken = th.gather(tq, 3, batch["actions"]).squeeze(3)
        out["fw/td_lambda_targets"] = build_td_lambda_targets(rewards, terminated, mask, taken, tiny.n_agents, args.gamma,
                                                               args.td_lambda).numpy()
        mac.init_hidden(B)
        pis = [mac.forward(batch, t=t) for t in range(T - 1)]
        out["fw/pi"] = th.stack(pis, 1).numpy()
        mac.init_hidden(B)
        out["fw/pi_test_t0"] = mac.forward(batch, t=0, test_mode=True).numpy()
        # action selection: Categorical(masked policies).sample() consumes one exponential_ of the probs' shape
        mac.init_hidden(B)
        th.manual_seed(300)
        acts = mac.select_actions(batch, t_ep=0, t_env=1234, bs=slice(None), test_mode=False)
        th.manual_seed(300)
        expo = th.empty(B * tiny.n_agents, tiny.n_actions).exponential_()
        out["sel/actions"] = acts.numpy()
        out["sel/expo"] = expo.numpy().reshape(B, tiny.n_agents, tiny.n_actions)
        out["sel/epsilon"] = np.float64(mac.action_selector.epsilon)
        mac.init_hidden(B)
        out["sel/greedy"] = mac.select_actions(batch, t_ep=0, t_env=1234, test_mode=True).numpy()
    mac.action_selector.epsilon = eps                   # select_action moved it; train() reads the attribute
    for step in range(n_steps):
        learner.train(batch, step, 0)
        for key in ("critic_loss", "critic_grad_norm", "td_error_abs", "q_taken_mean", "target_mean", "advantage_mean",
                    "coma_loss", "agent_grad_norm", "pi_max"):
            out["step%d/stat/%s" % (step, key)] = np.float64(logger.stats[key][-1][1])
        for tag, mod in (("agent", mac.agent), ("critic", learner.critic), ("target_critic", learner.target_critic)):
            for k, v in state_np(mod).items():
                out["step%d/%s/%s" % (step, tag, k)] = v
        out["step%d/critic_training_steps" % step] = np.int64(learner.critic_training_steps)
        for tag, opt, ps in (("agent", learner.agent_optimiser, learner.agent_params),
                             ("critic", learner.critic_optimiser, learner.critic_params)):
            out["step%d/sq/%s" % (step, tag)] = np.concatenate([opt.state[p_]["square_avg"].numpy().ravel() for p_ in ps])
    out["meta"] = np.array(repr(dict(name=name, shape=tuple(tiny), B=B, T=T, n_steps=n_steps, over=kw)))
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))

if __name__ == "__main__":
    th.set_num_threads(1)
    only = set(sys.argv[1:])            # e.g. `make_golden.py checkpoint episode_update`: regenerate just those
    if only:
        for name in sorted(only):
            {"checkpoint": run_checkpoint, "episode_update": run_episode_update, "select_actions": run_select_actions,
             "replay_sample": run_replay_sample, "coma": run_coma}[name]()
        sys.exit(0)
    tiny = SmacShape("tiny", 3, 10, 14, 5, 8)
    small = dict(rnn_hidden_dim=16, mixing_embed_dim=8)
    run_case("qmix_tiny", tiny, B=4, T=8, mixer="qmix", double_q=True, seed=1, **small)
    run_case("vdn_tiny", tiny, B=4, T=8, mixer="vdn", double_q=True, seed=2, **small)
    run_case("iql_tiny", tiny, B=4, T=8, mixer=None, double_q=True, seed=3, **small)
    run_case("qmix_nodouble_tiny", tiny, B=3, T=6, mixer="qmix", double_q=False, seed=4, **small)
    run_case("qmix_noid_tiny", tiny, B=3, T=6, mixer="qmix", double_q=True, seed=6,
             obs_agent_id=False, obs_last_action=False, **small)
    # real 3m shapes with the default widths (H = 64, E = 32), short T
    m3 = SmacShape("3m", 3, 30, 48, 9, 61)
    run_case("qmix_3m", m3, B=5, T=12, mixer="qmix", double_q=True, seed=5, n_steps=2)
    run_select_actions()
    run_replay_sample()
    run_checkpoint()
    run_episode_update()
    run_coma()
