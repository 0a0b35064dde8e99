"""GPU parity tests of the COMA path through the Python mirror and the C ABI: golden fixture
produced by the reference's COMALearner (tests/golden/coma_tiny.npz) and the numpy oracle (oracle/coma_oracle.py, pinned to
the same fixture) on SMAC-shaped batches.  fp32 tier: statistics, Q, targets, policy and post-update parameters within
1e-5; sampled actions bit-exact given the generator's draws."""
import copy

import numpy as np
import pytest
import torch as th

from golden_utils import Golden, rel_err
from oracle import coma_oracle as co, qlearner_oracle as orc
from pymarl_b200.synthetic import SMAC_SHAPES, default_args, numpy_episode_fields, make_scheme, get_shape

pytestmark = pytest.mark.gpu
TOL = 1e-5
COMA_KW = dict(agent_output_type="pi_logits", action_selector="multinomial", learner="coma_learner", critic_lr=5e-4,
               td_lambda=0.8, mask_before_softmax=True, epsilon_start=0.5, epsilon_finish=0.01, epsilon_anneal_time=100000,
               learner_log_interval=0, test_greedy=True)


def build_coma(shape, args, agent=None, critic=None, target_critic=None):
    from cuda_utils import Logger
    from pymarl_b200 import le_REGISTRY, mac_REGISTRY
    shape = get_shape(shape)
    scheme, groups = make_scheme(shape)
    scheme["actions_onehot"] = {"vshape": (shape.n_actions,), "dtype": th.float32, "group": "agents"}
    args.device, args.use_cuda = "cuda", True
    mac = mac_REGISTRY[args.mac](copy.deepcopy(scheme), groups, args)
    logger = Logger()
    learner = le_REGISTRY["coma_learner"](mac, scheme, logger, args)
    learner.cuda()
    ld = lambda m, p: m.load_state_dict({k: th.from_numpy(np.asarray(v, np.float32)) for k, v in p.items()}) if p else None
    ld(mac.agent, agent)
    ld(learner.critic, critic)
    ld(learner.target_critic, target_critic if target_critic is not None else critic)
    return learner, logger


def state_np(m):
    return {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}


def test_coma_golden_forward_pieces_and_action_selection():
    from cuda_utils import to_batch
    g = Golden("coma_tiny")
    args = default_args(g.shape, mixer=None, **g.meta["over"])
    learner, _ = build_coma(g.shape, args, g.group("init/agent"), g.group("init/critic"), g.group("init/target_critic"))
    f = g.batch_fields()
    batch = to_batch(g.shape, f)
    B, T = batch.batch_size, batch.max_seq_length
    # COMACritic.forward: all timesteps and a single one
    tq = learner.target_critic(batch)
    assert rel_err(tq.cpu().numpy(), g["fw/target_q"]) < TOL
    q3 = learner.target_critic(batch, t=3)
    assert q3.shape == (B, 1, g.shape.n_agents, g.shape.n_actions)
    assert rel_err(q3.cpu().numpy()[:, 0], g["fw/target_q"][:, 3]) < TOL
    # BasicMAC.forward with agent_output_type = pi_logits: masked softmax + epsilon floor; plain softmax in test mode
    mac = learner.mac
    assert mac.action_selector.epsilon == float(g["epsilon"])
    mac.init_hidden(B)
    for t in range(T - 1):
        pi = mac.forward(batch, t)
        assert rel_err(pi.cpu().numpy(), g["fw/pi"][:, t]) < TOL, t
    mac.init_hidden(B)
    assert rel_err(mac.forward(batch, 0, test_mode=True).cpu().numpy(), g["fw/pi_test_t0"]) < TOL
    # MultinomialActionSelector: the generator's exponential draws injected -> bit-exact actions; greedy in test mode
    expo = th.from_numpy(g["sel/expo"]).cuda().reshape(-1, g.shape.n_actions).contiguous()
    mac.action_selector.draw = lambda x: expo
    mac.init_hidden(B)
    acts = mac.select_actions(batch, t_ep=0, t_env=1234)
    np.testing.assert_array_equal(acts.cpu().numpy(), g["sel/actions"])
    assert mac.action_selector.epsilon == float(g["sel/epsilon"])
    mac.init_hidden(B)
    np.testing.assert_array_equal(mac.select_actions(batch, t_ep=0, t_env=1234, test_mode=True).cpu().numpy(), g["sel/greedy"])
    # the kernel's own generator: only available actions, roughly the policy's frequencies
    del mac.action_selector.draw
    mac.action_selector.rng = "philox"
    probs = th.tensor([[0.1, 0.0, 0.6, 0.3, 0.0]], device="cuda").repeat(20000, 1).view(20000, 1, 5)
    avail = (probs > 0).int()
    a = mac.action_selector.select_action(probs, avail, 0)
    freq = th.bincount(a.flatten(), minlength=5).float() / 20000
    assert freq[1] == 0 and freq[4] == 0 and abs(freq[2] - 0.6) < 0.02 and abs(freq[0] - 0.1) < 0.02


def test_coma_train_steps_match_reference_golden():
    from cuda_utils import to_batch
    g = Golden("coma_tiny")
    args = default_args(g.shape, mixer=None, **g.meta["over"])
    learner, logger = build_coma(g.shape, args, g.group("init/agent"), g.group("init/critic"), g.group("init/target_critic"))
    batch = to_batch(g.shape, g.batch_fields())
    learner.mac.action_selector.epsilon = float(g["epsilon"])
    for step in range(g.meta["n_steps"]):
        learner.train(batch, step, 0)
        if step == 0:
            ws = learner.workspace_views()
            assert rel_err(ws["targets"].cpu().numpy(), g["fw/td_lambda_targets"]) < TOL
        for k in ("critic_loss", "critic_grad_norm", "td_error_abs", "q_taken_mean", "target_mean", "advantage_mean",
                  "coma_loss", "agent_grad_norm", "pi_max"):
            r = float(g["step%d/stat/%s" % (step, k)])
            assert abs(logger.stats[k][-1][1] - r) <= 3 * TOL * max(1.0, abs(r)), (step, k, logger.stats[k][-1][1], r)
        for tag, mod in (("agent", learner.mac.agent), ("critic", learner.critic), ("target_critic", learner.target_critic)):
            for k, v in g.group("step%d/%s" % (step, tag)).items():
                assert rel_err(state_np(mod)[k], v) < 5 * TOL, (step, tag, k)
        assert learner.critic_training_steps == int(g["step%d/critic_training_steps" % step])
        for tag, opt in (("agent", learner.agent_optimiser), ("critic", learner.critic_optimiser)):
            sq = np.concatenate([s.cpu().numpy().ravel() for s in opt.square_avg])
            assert rel_err(sq, g["step%d/sq/%s" % (step, tag)]) < 1e-4, (step, tag)
    assert logger.infos.count("Updated target network") >= 1


@pytest.mark.parametrize("shape_name,B,T", [("3m", 16, 20), ("2s3z", 8, 12), ("MMM2", 6, 8)])
def test_coma_train_step_matches_oracle(shape_name, B, T):
    """SMAC-shaped batches (ragged episodes, all-padding timesteps at the end -> skipped critic steps) against the oracle:
    q_vals, targets, the renormalised policy, every statistic and the post-update parameters of agent and critic."""
    from cuda_utils import to_batch
    shape = SMAC_SHAPES[shape_name]
    args = default_args(shape, mixer=None, **COMA_KW)
    rng = np.random.default_rng(5)
    d_in = shape.obs_dim + shape.n_actions + shape.n_agents
    agent = orc.init_params(orc.agent_param_shapes(d_in, 64, shape.n_actions), rng)
    crit = co.init_critic(co.critic_param_shapes(shape.state_dim, shape.obs_dim, shape.n_agents, shape.n_actions), rng)
    tcrit = {k: (v + 0.05 * rng.standard_normal(v.shape)).astype(np.float32) for k, v in crit.items()}
    fields = numpy_episode_fields(shape, B, T, seed=17, ragged=True)
    L = fields["filled"][:, :, 0].sum(1).max()
    for k in fields:                                       # make the last timesteps padding for every episode
        fields[k][:, T - 3:] = 0
    olr = co.OracleCOMALearner(agent, crit, copy.copy(args))
    olr.target_critic = {k: v.copy() for k, v in tcrit.items()}
    learner, logger = build_coma(shape, args, agent, crit, tcrit)
    eps = 0.37
    learner.mac.action_selector.epsilon = eps
    st, _, fw = olr.train(fields, 0, 0, eps)
    learner.train(to_batch(shape, fields), 0, 0)
    ws = learner.workspace_views()
    mask = fields["filled"][:, :-1].astype(np.float32)
    mask[:, 1:] *= 1 - fields["terminated"][:, :-2].astype(np.float32)
    live_t = mask[:, :, 0].sum(0) > 0                      # q_vals of skipped critic steps stay 0 in the reference
    assert rel_err(ws["targets"].cpu().numpy(), fw["targets"]) < TOL
    assert rel_err(ws["q_vals"].cpu().numpy()[:, live_t], fw["q_vals"][:, live_t]) < TOL
    pi = ws["pi"].view(T - 1, B, shape.n_agents, -1).permute(1, 0, 2, 3).cpu().numpy()
    assert rel_err(pi, fw["pi"]) < TOL
    for k in ("critic_loss", "critic_grad_norm", "td_error_abs", "q_taken_mean", "target_mean", "advantage_mean",
              "coma_loss", "agent_grad_norm", "pi_max"):
        got = logger.stats[k][-1][1]
        assert abs(got - st[k]) <= 3 * TOL * max(1.0, abs(st[k])), (k, got, st[k])
    assert learner.critic_training_steps == olr.critic_training_steps < T - 1
    for k, v in olr.agent.items():
        assert rel_err(state_np(learner.mac.agent)[k], v) < 5 * TOL, k
    for k, v in olr.critic.items():
        assert rel_err(state_np(learner.critic)[k], v) < 5 * TOL, k
    del L


def test_coma_checkpoint_round_trip(tmp_path):
    from cuda_utils import to_batch
    shape = SMAC_SHAPES["3m"]
    args = default_args(shape, mixer=None, **COMA_KW)
    fields = numpy_episode_fields(shape, 8, 10, seed=2, ragged=True)
    batch = to_batch(shape, fields)
    a, _ = build_coma(shape, copy.copy(args))
    a.train(batch, 0, 0)
    a.save_models(str(tmp_path))
    import os
    assert sorted(os.listdir(str(tmp_path))) == ["agent.th", "agent_opt.th", "critic.th", "critic_opt.th"]
    b, _ = build_coma(shape, copy.copy(args))
    b.load_models(str(tmp_path))
    a.target_critic.load_state_dict(a.critic.state_dict())           # what load_models does on the other side
    a.train(batch, 1, 0)
    b.train(batch, 1, 0)
    for m1, m2 in ((a.mac.agent, b.mac.agent), (a.critic, b.critic)):
        for k, v in m1.state_dict().items():
            assert th.equal(v, m2.state_dict()[k]), k


def test_coma_cuda_graph_step_is_bit_identical():
    """args.cuda_graph: the ~14 launches per timestep of a COMA step replayed as one graph give the bits of the eager step."""
    from cuda_utils import to_batch
    shape = SMAC_SHAPES["3m"]
    fields = numpy_episode_fields(shape, 8, 16, seed=6, ragged=True)
    rng = np.random.default_rng(9)
    agent = orc.init_params(orc.agent_param_shapes(42, 64, 9), rng)
    crit = co.init_critic(co.critic_param_shapes(48, 30, 3, 9), rng)
    outs = []
    for graph in (False, True):
        args = default_args(shape, mixer=None, cuda_graph=graph, **COMA_KW)
        lr, _ = build_coma(shape, args, agent, crit, crit)
        batch = to_batch(shape, fields)
        for i in range(4):
            lr.train(batch, i, 0)
        if graph:
            assert any(isinstance(v, tuple) for v in lr._graphs.values())
        outs.append((lr._flat["ap"].clone(), lr._flat["cp"].clone(), lr._flat["tp"].clone(), lr.last_stats.clone()))
    for x, y in zip(*outs):
        assert th.equal(x, y)
