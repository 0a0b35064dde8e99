"""Pin the numpy oracle (oracle/qlearner_oracle.py) to the reference: every function is
checked against fixtures produced by the unmodified reference (tests/golden/make_golden.py).
fp64 oracle vs fp64 reference must agree to ~1e-12; fp32 vs fp32 to the fp32 noise floor."""
import numpy as np
import pytest

from golden_utils import Golden, LEARNER_CASES, rel_err
from oracle import qlearner_oracle as orc
from pymarl_b200.synthetic import default_args


def _learner(g, dtype):
    agent = {k: v.astype(dtype) for k, v in g.group("init/agent").items()}
    mixer = {k: v.astype(dtype) for k, v in g.group("init/mixer").items()}
    lr = orc.OracleQLearner(agent, mixer, g.args())
    for k, v in g.group("init/target_agent").items():
        lr.target_agent[k] = v.astype(dtype)
    for k, v in g.group("init/target_mixer").items():
        lr.target_mixer_p[k] = v.astype(dtype)
    return lr


@pytest.mark.parametrize("case", LEARNER_CASES)
@pytest.mark.parametrize("tag,dtype,tol", [("f64", np.float64, 1e-11), ("f32", np.float32, 2e-5)])
def test_forward_matches_reference(case, tag, dtype, tol):
    g = Golden(case)
    lr = _learner(g, dtype)
    fw = lr.forward_loss(g.batch_fields())
    ref = g.group(tag + "/fw")
    for k in ("mac_out", "target_mac_out", "chosen", "target_max", "q_tot", "target_tot", "targets"):
        assert rel_err(fw[k], ref[k]) < tol, k
    # integer output: bit-exact (fp64 has no near-ties; fp32 uses its own q values -> compare
    # against the argmax of the reference's fp32 mac_out as well)
    if tag == "f64":
        np.testing.assert_array_equal(fw["cur_max_actions"], ref["cur_max_actions"])


@pytest.mark.parametrize("case", LEARNER_CASES)
def test_target_select_bit_exact_on_reference_q(case):
    g = Golden(case)
    ref = g.group("f32/fw")
    f = g.batch_fields()
    chosen, tmax, cur = orc.target_select(ref["mac_out"], ref["target_mac_out"], f["avail_actions"],
                                          f["actions"], g.meta["double_q"])
    np.testing.assert_array_equal(cur, ref["cur_max_actions"])
    np.testing.assert_array_equal(chosen, ref["chosen"])
    np.testing.assert_array_equal(tmax, ref["target_max"])


@pytest.mark.parametrize("case", LEARNER_CASES)
@pytest.mark.parametrize("tag,dtype,tol", [("f64", np.float64, 1e-10), ("f32", np.float32, 3e-5)])
def test_gradients_match_autograd(case, tag, dtype, tol):
    g = Golden(case)
    lr = _learner(g, dtype)
    f = g.batch_fields()
    fw = lr.forward_loss(f)
    grads = lr.backward(fw, f)
    for k, v in g.group(tag + "/grad/agent").items():
        assert rel_err(grads["agent." + k], v) < tol, k
    for k, v in g.group(tag + "/grad/mixer").items():
        assert rel_err(grads["mixer." + k], v) < tol, k


@pytest.mark.parametrize("case", LEARNER_CASES)
@pytest.mark.parametrize("tag,dtype,tol", [("f64", np.float64, 1e-9), ("f32", np.float32, 2e-5)])
def test_train_steps_match_reference(case, tag, dtype, tol):
    g = Golden(case)
    lr = _learner(g, dtype)
    f = g.batch_fields()
    names = ["agent." + k for k in orc.AGENT_PARAM_NAMES]
    if g.meta["mixer"] == "qmix":
        names += ["mixer." + k for k in orc.QMIX_PARAM_NAMES]
    for step, (t_env, ep) in enumerate(g.episode_schedule()):
        stats, _, _ = lr.train(f, t_env, ep)
        for key in ("loss", "grad_norm", "td_error_abs", "q_taken_mean", "target_mean"):
            ref = float(g["%s/step%d/stat/%s" % (tag, step, key)])
            assert abs(stats[key] - ref) <= tol * max(1.0, abs(ref)), (step, key, stats[key], ref)
        pre = "%s/step%d/" % (tag, step)
        for k, v in g.group(pre + "agent").items():
            assert rel_err(lr.agent[k], v) < tol, (step, k)
        for k, v in g.group(pre + "mixer").items():
            assert rel_err(lr.mixer_p[k], v) < tol, (step, k)
        for k, v in g.group(pre + "target_agent").items():
            assert rel_err(lr.target_agent[k], v) < tol, (step, k)
        for k, v in g.group(pre + "target_mixer").items():
            assert rel_err(lr.target_mixer_p[k], v) < tol, (step, k)
        if g.has(pre + "square_avg_flat"):
            sq = {"agent." + k: v for k, v in lr.sq_agent.items()}
            sq.update({"mixer." + k: v for k, v in lr.sq_mixer.items()})
            flat = np.concatenate([sq[n].ravel() for n in names])
            assert rel_err(flat, g[pre + "square_avg_flat"]) < max(tol, 1e-4 if tag == "f32" else 0), step
    assert lr.n_target_updates == 1


def test_unknown_mixer_raises():
    g = Golden("qmix_tiny")
    with pytest.raises(ValueError, match="not recognised"):
        orc.OracleQLearner(g.group("init/agent"), {}, g.args(mixer="bogus"))


def test_select_action_bit_exact():
    g = Golden("select_actions")
    for ci in range(int(g["n_sel"])):
        p = "sel%d/" % ci
        eps = orc.epsilon_schedule(int(g[p + "t_env"])) if not bool(g[p + "test_mode"]) else 0.0
        assert eps == float(g[p + "epsilon"])
        acts = orc.select_action(g[p + "q"], g[p + "avail"], eps, g[p + "u"], g[p + "expo"])
        np.testing.assert_array_equal(acts, g[p + "actions"])


def test_mac_select_actions_bit_exact():
    g = Golden("select_actions")
    p = g.group("mac/agent")
    fields = g.group("mac/in")
    B, N = fields["obs"].shape[0], fields["obs"].shape[2]
    h = np.zeros((B * N, p["fc1.weight"].shape[0]), np.float32)
    for t in range(3):
        eps = float(g["mac/t%d/epsilon" % t])
        assert eps == orc.epsilon_schedule(20000)
        acts, q, h = orc.mac_select_actions(p, fields, t, h, eps, g["mac/t%d/u" % t], g["mac/t%d/expo" % t])
        np.testing.assert_array_equal(acts, g["mac/t%d/actions" % t])
        assert rel_err(h, g["mac/t%d/hidden" % t]) < 1e-5


def test_replay_sample_ids_and_max_t():
    g = Golden("replay_sample")
    fields = g.group("in")
    for seed in (0, 1, 7):
        ids = orc.replay_sample_ids(12, 5, seed)
        np.testing.assert_array_equal(fields["obs"][ids], g["seed%d/obs" % seed])
        assert orc.max_t_filled(fields["filled"][ids]) == int(g["seed%d/max_t_filled" % seed])


# ------------------------------------------------------------------------------------------------------------------
# COMA: oracle/coma_oracle.py against the reference's COMALearner / COMACritic / BasicMAC
# (pi_logits) / MultinomialActionSelector, fixture tests/golden/coma_tiny.npz
# ------------------------------------------------------------------------------------------------------------------
def _coma_setup():
    from oracle import coma_oracle as co
    g = Golden("coma_tiny")
    args = default_args(g.shape, mixer=None, **g.meta["over"])
    lr = co.OracleCOMALearner(g.group("init/agent"), g.group("init/critic"), args)
    lr.target_critic = {k: v.copy() for k, v in g.group("init/target_critic").items()}
    return co, g, args, lr


def test_coma_critic_inputs_targets_and_policy_match_reference():
    co, g, args, lr = _coma_setup()
    f = g.batch_fields()
    B, T, N = f["obs"].shape[:3]
    np.testing.assert_array_equal(co.critic_inputs(f), g["fw/critic_inputs"])
    tq = co.critic_forward(lr.target_critic, co.critic_inputs(f).reshape(B * T * N, -1)).reshape(B, T, N, -1)
    assert rel_err(tq, g["fw/target_q"]) < 2e-6
    rewards, term = f["reward"][:, :-1], f["terminated"][:, :-1].astype(np.float32)
    mask = f["filled"][:, :-1].astype(np.float32)
    mask[:, 1:] = mask[:, 1:] * (1 - term[:, :-1])
    taken = np.take_along_axis(g["fw/target_q"], f["actions"], axis=3)[..., 0]
    assert rel_err(co.td_lambda_targets(rewards, term, mask, taken, args.gamma, args.td_lambda), g["fw/td_lambda_targets"]) < 1e-6
    logits, _ = orc.mac_unroll(lr.agent, {k: v[:, :-1] for k, v in f.items()}, True, True, False)
    A = logits.shape[-1]
    pi = co.policy_head(logits.reshape(-1, A), f["avail_actions"][:, :-1].reshape(-1, A), np.float32(g["epsilon"]))
    assert rel_err(pi.reshape(B, T - 1, N, A), g["fw/pi"]) < 1e-6
    pt = co.policy_head(logits[:, 0].reshape(-1, A), f["avail_actions"][:, 0].reshape(-1, A), 0, test_mode=True)
    assert rel_err(pt.reshape(B, N, A), g["fw/pi_test_t0"]) < 1e-6
    # MultinomialActionSelector with the generator's Exp(1) draws: bit-exact actions; greedy in test mode
    acts = co.multinomial_select(g["fw/pi"][:, 0], f["avail_actions"][:, 0], g["sel/expo"])
    np.testing.assert_array_equal(acts, g["sel/actions"])
    np.testing.assert_array_equal(co.multinomial_select(g["fw/pi_test_t0"], f["avail_actions"][:, 0], None, True), g["sel/greedy"])


def test_coma_train_steps_match_reference():
    co, g, args, lr = _coma_setup()
    f = g.batch_fields()
    for step in range(g.meta["n_steps"]):
        st, _, _ = lr.train(f, step, 0, float(g["epsilon"]))
        for k in ("critic_loss", "critic_grad_norm", "td_error_abs", "q_taken_mean", "target_mean", "advantage_mean",
                  "coma_loss", "agent_grad_norm", "pi_max"):
            r = float(g["step%d/stat/%s" % (step, k)])
            assert abs(st[k] - r) <= 2e-6 * max(1.0, abs(r)), (step, k, st[k], r)
        for tag, d in (("agent", lr.agent), ("critic", lr.critic), ("target_critic", lr.target_critic)):
            for k, v in g.group("step%d/%s" % (step, tag)).items():
                assert rel_err(d[k], v) < 5e-6, (step, tag, k)
        assert lr.critic_training_steps == int(g["step%d/critic_training_steps" % step])
    assert lr.n_target_updates >= 1                       # the fixture's schedule crosses target_update_interval
