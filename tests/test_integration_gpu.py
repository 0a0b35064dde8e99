"""GPU integration tests around the hot path: device-side EpisodeBatch.update / insert, checkpoints
(q_learner.py:124-143), the CUDA-graphed step, the reference's own driver (run.run_sequential) running this package's
classes, the vectorised rollout loop, and the data-parallel step on 2 GPUs against the 1-GPU step."""
import copy
import os
import subprocess
import sys

import numpy as np
import pytest
import torch as th

from golden_utils import Golden, rel_err
from oracle import qlearner_oracle as orc
from pymarl_b200.synthetic import SMAC_SHAPES, SmacShape, numpy_episode_fields, default_args

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-5


def test_device_episode_batch_update_is_bit_exact():
    """EpisodeBatch.update / ReplayBuffer.insert_episode_batch with the batch in HBM: ONE pmb_batch_update launch per call
    (all fields + filled + fused OneHot) reproduces, bit for bit, what the reference's classes produced for the same
    runner-style call sequence (list-valued bs, scalar ts, mark_filled on/off, a ring buffer that wraps)."""
    from test_host_logic import replay_update_golden
    from pymarl_b200 import _lib
    n0 = _lib.lib().pmb_launch_count()
    eb, buf, g = replay_update_golden("cuda")
    assert _lib.lib().pmb_launch_count() > n0                       # the kernel path ran, not the torch fallback
    for k, v in eb.data.transition_data.items():
        assert v.is_cuda
        np.testing.assert_array_equal(v.cpu().numpy(), g["final/" + k], err_msg=k)
    for k, v in buf.data.transition_data.items():
        np.testing.assert_array_equal(v.cpu().numpy(), g["buffer/" + k], err_msg=k)
    assert buf.buffer_index == int(g["buffer/index"]) and buf.episodes_in_buffer == int(g["buffer/episodes"])
    # slice-valued bs / ts with tensors already on the device, and the error behaviour of the reference
    shape = SmacShape("tiny", 3, 6, 7, 5, 6)
    from pymarl_b200 import EpisodeBatch
    from pymarl_b200.components.transforms import OneHot
    from pymarl_b200.synthetic import make_scheme
    scheme, groups = make_scheme(shape)
    pre = {"actions": ("actions_onehot", [OneHot(out_dim=shape.n_actions)])}
    a, b = (EpisodeBatch(scheme, groups, 6, 6, preprocess=pre, device=d) for d in ("cpu", "cuda"))
    rng = np.random.default_rng(0)
    obs = rng.standard_normal((3, 2, 3, 6)).astype(np.float32)
    acts = rng.integers(0, 5, (3, 2, 3, 1))
    for eb_, dev in ((a, "cpu"), (b, "cuda")):
        eb_.update({"obs": th.from_numpy(obs).to(dev), "actions": th.from_numpy(acts).to(dev)}, bs=slice(1, 6, 2), ts=slice(2, 4))
    for k in a.data.transition_data:
        np.testing.assert_array_equal(a[k].numpy(), b[k].cpu().numpy(), err_msg=k)
    with pytest.raises(ValueError):
        b.update({"reward": np.zeros((6, 6, 3), np.float32)})
    with pytest.raises(KeyError):
        b.update({"nope": np.zeros(1)})


def test_reference_checkpoint_loads_and_trains_identically():
    """A checkpoint WRITTEN BY THE REFERENCE (tests/golden/ckpt_ref: agent.th, mixer.th, opt.th) loaded through
    QLearner.load_models, then one train step: loaded state bit-exact (incl. the reference quirk that the target MAC takes
    the online weights and the target mixer is left alone), the step's statistics / parameters / RMSprop state as the
    reference's own continuation."""
    from cuda_utils import learner_from_golden, to_batch, state_np
    g = Golden("checkpoint")
    learner, _ = learner_from_golden(g)
    learner.load_models(os.path.join(REPO, "tests", "golden", "ckpt_ref"))
    for tag, mod in (("agent", learner.mac.agent), ("target_agent", learner.target_mac.agent), ("mixer", learner.mixer),
                     ("target_mixer", learner.target_mixer)):
        for k, v in g.group("loaded/" + tag).items():
            np.testing.assert_array_equal(state_np(mod)[k], v, err_msg=tag + "." + k)
    learner.train(to_batch(g.shape, g.batch_fields()), 2, 0)
    st = learner.stats()
    for key in ("loss", "grad_norm", "td_error_abs", "q_taken_mean", "target_mean"):
        r = float(g["after/stat/" + key])
        assert abs(st[key] - r) <= 3 * TOL * max(1.0, abs(r)), (key, st[key], r)
    for tag, mod in (("agent", learner.mac.agent), ("mixer", learner.mixer)):
        for k, v in g.group("after/" + tag).items():
            assert rel_err(state_np(mod)[k], v) < 5 * TOL, (tag, k)
    sd = learner.optimiser.state_dict()
    flat = np.concatenate([sd["state"][i]["square_avg"].cpu().numpy().ravel() for i in range(len(learner.params))])
    assert rel_err(flat, g["after/square_avg_flat"]) < 1e-4
    assert int(sd["state"][0]["step"]) == 3


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_save_load_round_trip_on_gpu(tmp_path, precision):
    """save_models -> a NEW learner -> load_models: the next step is bit-identical to the original learner's next step
    (the target mixer copied by hand: the reference's load_models leaves it alone, q_learner.py:137-143), and a restored
    learning rate is honoured."""
    from cuda_utils import build_learner, to_batch
    shape = SMAC_SHAPES["2s3z"]
    fields = numpy_episode_fields(shape, 16, 20, seed=3, ragged=True)
    args = default_args(shape, mixer="qmix", learner_log_interval=0, precision=precision)
    rng = np.random.default_rng(1)
    d_in = shape.obs_dim + shape.n_actions + shape.n_agents
    agent = orc.init_params(orc.agent_param_shapes(d_in, 64, shape.n_actions), rng)
    mixer = orc.init_params(orc.qmix_param_shapes(shape.state_dim, shape.n_agents, 32), rng)
    a, _ = build_learner(shape, copy.copy(args), agent, agent, mixer, mixer)
    batch = to_batch(shape, fields)
    a.train(batch, 0, 0)
    a.train(batch, 1, 0)
    a.save_models(str(tmp_path))
    assert sorted(os.listdir(str(tmp_path))) == ["agent.th", "mixer.th", "opt.th"]
    b, _ = build_learner(shape, copy.copy(args))
    b.load_models(str(tmp_path))
    b.target_mixer.load_state_dict(a.target_mixer.state_dict())
    # the reference quirk: after load_models the target MAC holds the ONLINE weights; mirror it on `a` for the comparison
    a.target_mac.load_state(a.mac)
    assert b.optimiser.step_count == 2
    a.train(batch, 2, 0)
    b.train(batch, 2, 0)
    assert th.equal(a._flat["p"], b._flat["p"]) and th.equal(a._flat["sq"], b._flat["sq"])
    assert a.stats() == b.stats()
    # hyper-parameters travel with opt.th like torch's param_groups: a checkpoint with another lr changes the step
    sd = a.optimiser.state_dict()
    sd["param_groups"][0]["lr"] = 0.0
    b.optimiser.load_state_dict(sd)
    before = b._flat["p"].clone()
    b.train(batch, 3, 0)
    assert th.equal(before, b._flat["p"])


@pytest.mark.parametrize("mixer,precision", [("qmix", "bf16"), ("qmix", "fp32"), ("vdn", "bf16"), (None, "bf16")])
def test_cuda_graph_step_is_bit_identical(mixer, precision):
    """args.cuda_graph: the whole step captured once and replayed gives the same bits as the eager launches, over steps
    that include a target sync (a second graph) and a changed batch at the same address."""
    from cuda_utils import build_learner, to_batch
    shape = SMAC_SHAPES["3m"]
    fields = numpy_episode_fields(shape, 32, 60, seed=8, ragged=True)
    fields2 = numpy_episode_fields(shape, 32, 60, seed=9, ragged=True)
    args = default_args(shape, mixer=mixer, learner_log_interval=0, precision=precision)
    rng = np.random.default_rng(2)
    agent = orc.init_params(orc.agent_param_shapes(42, 64, 9), rng)
    mix = orc.init_params(orc.qmix_param_shapes(48, 3, 32), rng)
    outs = []
    for graph in (False, True):
        a = copy.copy(args)
        a.cuda_graph = graph
        lr, _ = build_learner(shape, a, agent, agent, mix, mix)
        batch = to_batch(shape, fields)
        sched = [(0, 0), (1, 0), (2, 0), (3, 200), (4, 201), (5, 201), (6, 201)]
        for i, (t_env, ep) in enumerate(sched):
            if i == 5:                                   # new episodes written into the SAME tensors
                for k, v in fields2.items():
                    batch.data.transition_data[k].copy_(th.from_numpy(np.ascontiguousarray(v)))
            lr.train(batch, t_env, ep)
        if graph:
            assert sum(isinstance(v, tuple) for v in lr._graphs.values()) >= 1
        outs.append((lr._flat["p"].clone(), lr._flat["target"].clone(), lr.stats()))
    assert th.equal(outs[0][0], outs[1][0]) and th.equal(outs[0][1], outs[1][1])
    assert outs[0][2] == outs[1][2]


def test_vdn_mixer_module_runs_the_kernel():
    from pymarl_b200 import VDNMixer
    qs = th.randn(7, 11, 5, device="cuda")
    y = VDNMixer()(qs, None)
    assert y.shape == (7, 11, 1)
    assert th.allclose(y, qs.sum(2, keepdim=True), atol=1e-6)


def test_reference_run_sequential_drives_this_package():
    """SURVEY.md section 4 tier 3: the reference's UNMODIFIED run.run_sequential (run.py:107-256) with this package's
    classes swapped into its registries (install_into_reference) trains QMIX on a synthetic MultiAgentEnv: runner ->
    BasicMAC.select_actions (CUDA) -> reference ReplayBuffer -> QLearner.train (CUDA), logging the five learner stats."""
    from oracle import ref_harness as rh
    if not rh.available():
        pytest.skip("reference sources not staged (python oracle/stage_reference.py in the build container)")
    code = r'''
import sys
sys.path.insert(0, %r)
from oracle import ref_harness as rh
rh.register_synthetic_env()
import pymarl_b200
pymarl_b200.install_into_reference()
import run as ref_run
import learners, controllers
assert learners.REGISTRY["q_learner"].__module__.startswith("pymarl_b200")
assert controllers.REGISTRY["basic_mac"].__module__.startswith("pymarl_b200")
from pymarl_b200.synthetic import SmacShape
from pymarl_b200 import _lib
shape = SmacShape("tiny", 3, 30, 48, 9, 13)
for mixer, prec in (("qmix", "fp32"), ("vdn", "bf16")):
    args = rh.run_sequential_args(shape, t_max=70, use_cuda=True, mixer=mixer, precision=prec)
    lg = rh.Logger()
    n0 = _lib.lib().pmb_launch_count()
    ref_run.run_sequential(args, lg)
    assert _lib.lib().pmb_launch_count() - n0 > 100
    for k in ("loss", "grad_norm", "td_error_abs", "q_taken_mean", "target_mean"):
        assert len(lg.stats[k]) >= 3, (k, lg.stats.keys())
        assert all(v == v and abs(v) < 1e6 for _, v in lg.stats[k]), (k, lg.stats[k])
    assert "Finished Training" in lg.infos
print("RUN_SEQUENTIAL_OK")
''' % REPO
    r = subprocess.run([sys.executable, "-W", "ignore", "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "RUN_SEQUENTIAL_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_vectorised_rollout_fills_a_device_batch():
    """the vectorised rollout loop (pymarl_b200.runners.VectorRunner) - B synthetic envs stepped on
    the device, select_actions + batch.update per timestep without host round trips - produces an EpisodeBatch that obeys
    the reference's layout contract (filled prefix, terminated once, one-hot of the stored actions, legal actions only)
    and trains."""
    from cuda_utils import build_learner
    from pymarl_b200.runners import VectorRunner, SyntheticVectorEnv
    shape = SMAC_SHAPES["3m"]
    args = default_args(shape, mixer="qmix", learner_log_interval=0, precision="bf16", device="cuda", batch_size_run=64,
                        action_rng="philox")
    learner, _ = build_learner(shape, args)
    env = SyntheticVectorEnv(64, shape.n_agents, shape.obs_dim, shape.state_dim, shape.n_actions, episode_limit=20, seed=3)
    runner = VectorRunner(args, env)
    runner.setup(learner.mac)
    batch = runner.run(test_mode=False)
    assert batch.batch_size == 64 and batch.max_seq_length == 21
    f = batch["filled"][:, :, 0]
    L = f.sum(1)
    assert bool((f == (th.arange(21, device="cuda")[None] < L[:, None]).long()).all())       # filled is a prefix
    term = batch["terminated"][:, :, 0].long()
    assert bool((term.sum(1) <= 1).all())
    acts, avail, oh = batch["actions"], batch["avail_actions"], batch["actions_onehot"]
    live = f.bool()
    assert bool((avail.gather(3, acts)[live] == 1).all())
    assert bool((oh.argmax(-1, keepdim=True)[live] == acts[live]).all()) and bool((oh.sum(-1)[live] == 1).all())
    assert bool((oh[~live] == 0).all())
    assert runner.t_env == int((L - 1).sum())
    learner.train(batch, runner.t_env, 64)
    st = learner.stats()
    assert np.isfinite(st["loss"]) and st["mask_sum"] > 0


@pytest.mark.skipif(not th.cuda.is_available() or th.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_data_parallel_step_equals_one_gpu():
    """SURVEY.md section 4 tier 4: two ranks (NCCL), each training on its shard of ONE batch through QLearner.train (which
    shards, exchanges [grads | loss sums] in one all-reduce and applies the replicated update), against one GPU on the
    whole batch: bench.py's dp_self_check on 2 ranks."""
    port = 29600 + os.getpid() % 1000
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), os.path.join(REPO, "tools", "dp_check.py")],
                       capture_output=True, text=True, timeout=900)
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("DP_CHECK ")]
    assert r.returncode == 0 and line, r.stdout[-2000:] + r.stderr[-3000:]
    import json
    out = json.loads(line[0][len("DP_CHECK "):])
    assert out["ok"], out


def test_rollout_weight_image_cache_follows_parameter_updates():
    """The rollout step keeps its packed bf16 weight images in the scratch between steps (pmb_dims.reserved bit 0) and must
    re-pack when the learner's kernels rewrite the parameters in place, when load_state / load_state_dict replace them, and
    when torch ops touch them: after each kind of change the cached MAC gives the bits of a freshly built MAC."""
    from cuda_utils import build_learner, to_batch
    from pymarl_b200 import mac_REGISTRY
    from pymarl_b200.synthetic import make_scheme
    shape = SMAC_SHAPES["2s3z"]
    args = default_args(shape, mixer="qmix", learner_log_interval=0, precision="bf16", device="cuda")
    learner, _ = build_learner(shape, args)
    fields = numpy_episode_fields(shape, 48, 12, seed=4, ragged=False)
    batch = to_batch(shape, fields)

    def fresh_q(t):
        scheme, groups = make_scheme(shape)
        m = mac_REGISTRY["basic_mac"](scheme, groups, copy.copy(args))
        m.cuda()
        m.agent.load_state_dict(learner.mac.agent.state_dict())
        m.init_hidden(48)
        return m.forward(batch, t)

    mac = learner.mac
    for t, change in enumerate(("none", "none", "train", "load_state", "torch_op")):
        if change == "train":
            learner.train(batch, 0, 0)
        elif change == "load_state":
            mac.load_state(learner.target_mac)
        elif change == "torch_op":
            with th.no_grad():
                mac.agent.fc2.weight.mul_(1.5)
        mac.init_hidden(48)
        q = mac.forward(batch, t)
        assert th.equal(q, fresh_q(t)), (t, change)


@pytest.mark.parametrize("mixer,precision", [("qmix", "fp32"), ("qmix", "bf16"), ("vdn", "bf16"), (None, "fp32")])
def test_streamed_host_batch_equals_device_batch(mixer, precision):
    """A host-resident (pinned) batch is trained in chunks of args.host_stream_chunk episodes that cross PCIe on a copy
    stream while the previous chunk computes (un-normalised gradients and loss sums added up, ONE update): over three
    steps - the second with a target sync - the loss, the grad norm and the post-update parameters equal the whole batch
    trained from device memory in one piece up to the summation order (1e-6 relative; loss sums are doubles), a ragged
    last chunk included, and two streamed runs give identical bits."""
    from cuda_utils import build_learner, to_batch
    shape = SMAC_SHAPES["2s3z"]
    B, T, chunk = 38, 14, 8                                   # chunks of 8, 8, 8, 8, 6 episodes
    fields = numpy_episode_fields(shape, B, T, seed=11, ragged=True)
    args = default_args(shape, mixer=mixer, learner_log_interval=0, precision=precision, target_update_interval=1)
    rng = np.random.default_rng(5)
    d_in = shape.obs_dim + shape.n_actions + shape.n_agents
    agent = orc.init_params(orc.agent_param_shapes(d_in, 64, shape.n_actions), rng)
    mix = orc.init_params(orc.qmix_param_shapes(shape.state_dim, shape.n_agents, 32), rng) if mixer == "qmix" else None
    dev_batch = to_batch(shape, fields)
    host_batch = to_batch(shape, fields, device="cpu")
    for k, v in host_batch.data.transition_data.items():
        host_batch.data.transition_data[k] = v.pin_memory()

    def run(batch, **over):
        lr, _ = build_learner(shape, copy.copy(args), agent, agent, mix, mix)
        for k, v in over.items():
            setattr(lr.args, k, v)
        lr._ensure_flat()
        lr._flat["sq"].fill_(1e-2)
        out = []
        for step, ep in ((0, 0), (1, 1), (2, 1)):             # the second step syncs the targets
            lr.train(batch, step, ep)
            out.append((lr.stats(), lr._flat["p"].clone(), lr._flat["target"].clone()))
        return lr, out

    _, ref = run(dev_batch)
    lr_s, got = run(host_batch, host_stream_chunk=chunk)
    assert lr_s._hs is not None and lr_s._hs["bufs"][0]["obs"].shape[0] == chunk        # the streamed path ran
    _, again = run(host_batch, host_stream_chunk=chunk)
    tol = 2e-6 if precision == "fp32" else 2e-5               # bf16 tier: per-tile partial sums regroup with the chunks
    for (s0, p0, t0), (s1, p1, t1), (s2, p2, t2) in zip(ref, got, again):
        for k in ("loss", "grad_norm", "td_error_abs", "q_taken_mean", "target_mean"):
            assert abs(s0[k] - s1[k]) <= tol * max(1.0, abs(s0[k])), (k, s0[k], s1[k])
        assert s0["mask_sum"] == s1["mask_sum"]
        err = float((p0 - p1).abs().max() / p0.abs().max())
        assert err <= 10 * tol, err
        assert float((t0 - t1).abs().max() / t0.abs().max()) <= 10 * tol
        assert th.equal(p1, p2) and s1 == s2
    # a host batch smaller than two chunks takes the one-piece path
    lr_o, _ = run(host_batch, host_stream_chunk=32)
    assert lr_o._hs is None
