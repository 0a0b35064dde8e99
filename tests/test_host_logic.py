"""CPU-only tests: the C-ABI library loads and exports what include/pymarl_b200.h declares,
host-side logic (layout, schedules, EpisodeBatch/ReplayBuffer, checkpoint format), and that
the product path fails loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch as th

from golden_utils import Golden
from pymarl_b200 import _lib, EpisodeBatch, ReplayBuffer, le_REGISTRY, mac_REGISTRY
from pymarl_b200.components.epsilon_schedules import DecayThenFlatSchedule
from pymarl_b200.components.transforms import OneHot
from pymarl_b200.synthetic import SmacShape, make_scheme, default_args, numpy_episode_fields

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(REPO, "include", "pymarl_b200.h")).read()
    declared = set(re.findall(r"^(?:int|int64_t|const char\*)\s+(pmb_[a-z0-9_]+)\s*\(", hdr, re.M))
    assert len(declared) >= 20
    cdll = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(cdll, name), "libpymarl_b200.so does not export %s" % name
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    assert _lib.lib().pmb_version() >= 100


def test_flat_layout_matches_parameter_shapes():
    d = _lib.make_dims(B=4, T=8, N=3, O=30, S=48, A=9, H=64, E=32, mixer="qmix")
    L = _lib.flat_layout(d)
    d_in = 30 + 9 + 3
    expect = [64 * d_in, 64, 192 * 64, 192 * 64, 192, 192, 9 * 64, 9,
              3 * 32 * 48, 32 * 48, 32 * 48, 32 * 48, 96, 32, 32, 32, 32, 1]
    assert list(L.numel) == expect
    assert list(L.offset) == list(np.cumsum([0] + expect[:-1]))
    assert L.n_agent == sum(expect[:8]) == 28297
    assert L.n_total - L.n_agent == 9441
    d2 = _lib.make_dims(B=4, T=8, N=3, O=30, S=48, A=9, H=64, E=32, mixer="vdn")
    assert _lib.flat_layout(d2).n_total == 28297


def test_invalid_dims_are_rejected_with_a_message():
    d = _lib.make_dims(B=4, T=8, N=3, O=30, S=48, A=9, H=48, E=32)
    L = _lib.Layout()
    rc = _lib.lib().pmb_flat_layout(ctypes.byref(d), ctypes.byref(L))
    assert rc == 1 and b"rnn_hidden_dim" in _lib.lib().pmb_last_error()
    with pytest.raises(ValueError, match="not recognised"):
        _lib.make_dims(4, 8, 3, 30, 48, 9, 64, 32, mixer="bogus")


def test_epsilon_schedule():
    s = DecayThenFlatSchedule(1.0, 0.05, 50000, decay="linear")
    assert s.eval(0) == 1.0 and s.eval(10 ** 7) == 0.05
    assert abs(s.eval(25000) - 0.525) < 1e-12
    g = Golden("select_actions")
    assert s.eval(20000) == float(g["mac/t0/epsilon"])


def _buffer(shape, size, device="cpu"):
    scheme, groups = make_scheme(shape)
    preprocess = {"actions": ("actions_onehot", [OneHot(out_dim=shape.n_actions)])}
    return scheme, groups, preprocess, ReplayBuffer(scheme, groups, size, shape.max_seq_length,
                                                    preprocess=preprocess, device=device)


def test_replay_buffer_matches_reference_sampling():
    g = Golden("replay_sample")
    shape = SmacShape("tiny", 2, 5, 6, 4, 5)
    scheme, groups, preprocess, buf = _buffer(shape, 16)
    fields = g.group("in")
    eb = EpisodeBatch(scheme, groups, 12, shape.max_seq_length, preprocess=preprocess)
    eb.update({k: fields[k] for k in ("state", "obs", "actions", "avail_actions", "reward", "terminated")},
              mark_filled=False)
    eb.data.transition_data["filled"] = th.from_numpy(fields["filled"])
    assert (eb["actions_onehot"].argmax(-1) == th.from_numpy(fields["actions"])[..., 0]).all()
    buf.insert_episode_batch(eb)
    assert buf.episodes_in_buffer == 12 and buf.can_sample(5) and not buf.can_sample(13)
    for seed in (0, 1, 7):
        np.random.seed(seed)
        s = buf.sample(5)
        np.testing.assert_array_equal(s["obs"].numpy(), g["seed%d/obs" % seed])
        np.testing.assert_array_equal(s["filled"].numpy(), g["seed%d/filled" % seed])
        assert int(s.max_t_filled()) == int(g["seed%d/max_t_filled" % seed])
        t = int(s.max_t_filled())
        cut = s[:, :t]
        assert cut.max_seq_length == t and cut["obs"].shape[1] == t
        assert cut["obs"].stride(0) == s["obs"].stride(0)              # view: full-T batch stride kept


def test_replay_buffer_ring_wraps():
    shape = SmacShape("tiny", 2, 5, 6, 4, 5)
    scheme, groups, preprocess, buf = _buffer(shape, 5)
    for i in range(3):
        eb = EpisodeBatch(scheme, groups, 2, shape.max_seq_length, preprocess=preprocess)
        eb.update({"reward": np.full((2, shape.max_seq_length, 1), float(i + 1), np.float32)})
        buf.insert_episode_batch(eb)
    assert buf.episodes_in_buffer == 5 and buf.buffer_index == 1
    assert buf["reward"][:, 0, 0].tolist() == [3.0, 1.0, 2.0, 2.0, 3.0]
    with pytest.raises(ValueError):
        eb.update({"reward": np.zeros((2, shape.max_seq_length, 3), np.float32)})
    with pytest.raises(KeyError):
        eb.update({"nope": np.zeros(1)})


def test_learner_constructs_on_cpu_but_refuses_to_run_without_cuda():
    shape = SmacShape("3m", 3, 30, 48, 9, 61)
    scheme, groups = make_scheme(shape)
    args = default_args(shape, mixer="qmix")
    mac = mac_REGISTRY["basic_mac"](scheme, groups, args)
    learner = le_REGISTRY["q_learner"](mac, scheme, None, args)
    assert [tuple(p.shape) for p in learner.params][:3] == [(64, 42), (64,), (192, 64)]
    assert set(learner.mixer.state_dict()) == {"hyper_w_1.weight", "hyper_w_1.bias", "hyper_w_final.weight",
                                               "hyper_w_final.bias", "hyper_b_1.weight", "hyper_b_1.bias",
                                               "V.0.weight", "V.0.bias", "V.2.weight", "V.2.bias"}
    assert set(mac.agent.state_dict()) == {"fc1.weight", "fc1.bias", "rnn.weight_ih", "rnn.weight_hh",
                                           "rnn.bias_ih", "rnn.bias_hh", "fc2.weight", "fc2.bias"}
    with pytest.raises(ValueError, match="Mixer bogus not recognised"):
        le_REGISTRY["q_learner"](mac, scheme, None, default_args(shape, mixer="bogus"))
    if not th.cuda.is_available():
        fields = numpy_episode_fields(shape, 2, 4, seed=0)
        batch = {k: th.from_numpy(v) for k, v in fields.items()}
        with pytest.raises(_lib.PmbError, match="no CPU path"):
            learner.train(batch, 0, 0)
        mac.init_hidden(2)
        with pytest.raises(_lib.PmbError, match="no CPU path"):
            mac.forward(type("B", (), {"batch_size": 2, "__getitem__": lambda s, k: batch[k]})(), 0)


def test_optimizer_state_dict_is_torch_rmsprop_compatible():
    from pymarl_b200.learners.q_learner import FusedRMSprop
    ps = [th.nn.Parameter(th.randn(3, 2)), th.nn.Parameter(th.randn(4))]
    ref = th.optim.RMSprop(ps, lr=5e-4, alpha=0.99, eps=1e-5)
    for p in ps:
        p.grad = th.randn_like(p)
    ref.step()
    mine = FusedRMSprop(ps, 5e-4, 0.99, 1e-5)
    mine.load_state_dict(ref.state_dict())
    sd = mine.state_dict()
    for i in range(2):
        assert th.equal(sd["state"][i]["square_avg"], ref.state_dict()["state"][i]["square_avg"])
    assert mine.step_count == 1
    ref2 = th.optim.RMSprop(ps, lr=5e-4, alpha=0.99, eps=1e-5)
    ref2.load_state_dict(sd)                                   # the reference can load our opt.th
    assert th.equal(ref2.state_dict()["state"][1]["square_avg"], sd["state"][1]["square_avg"])


def replay_update_golden(device):
    """Replays the update / insert calls recorded from the reference (tests/golden/make_golden.py:run_episode_update) on
    this package's EpisodeBatch / ReplayBuffer living on `device`; returns (batch, buffer, golden)."""
    import ast
    g = np.load(os.path.join(REPO, "tests", "golden", "episode_update.npz"))
    meta = ast.literal_eval(str(g["meta"]))
    shape = SmacShape(*meta["shape"])
    scheme, groups = make_scheme(shape)
    preprocess = {"actions": ("actions_onehot", [OneHot(out_dim=shape.n_actions)])}
    eb = EpisodeBatch(scheme, groups, meta["B"], meta["T"], preprocess=preprocess, device=device)
    for i, call in enumerate(meta["calls"]):
        data = {k: g["call%d/%s" % (i, k)] for k in call["keys"]}
        eb.update(data, bs=[int(b) for b in g["call%d/bs" % i]], ts=call["ts"], mark_filled=call["mark_filled"])
    buf = ReplayBuffer(scheme, groups, 8, meta["T"], preprocess=preprocess, device=device)
    buf.insert_episode_batch(eb)
    buf.insert_episode_batch(eb)
    return eb, buf, g


def test_episode_batch_update_matches_reference_golden():
    """EpisodeBatch.update / insert_episode_batch / OneHot (episode_buffer.py:98-154, 271-286; transforms.py:12-21): the
    host path against what the reference's classes produced for the same calls - every field bit-exact."""
    eb, buf, g = replay_update_golden("cpu")
    for k, v in eb.data.transition_data.items():
        np.testing.assert_array_equal(v.numpy(), g["final/" + k], err_msg=k)
    for k, v in buf.data.transition_data.items():
        np.testing.assert_array_equal(v.numpy(), g["buffer/" + k], err_msg=k)
    assert buf.buffer_index == int(g["buffer/index"]) and buf.episodes_in_buffer == int(g["buffer/episodes"])


def test_reference_checkpoint_loads_on_cpu_modules():
    """agent.th / mixer.th / opt.th written by the REFERENCE's save_models (tests/golden/ckpt_ref) load into this package's
    modules by name, the reference quirks included (target MAC <- online weights, target mixer untouched)."""
    g = Golden("checkpoint")
    shape = g.shape
    scheme, groups = make_scheme(shape)
    args = g.args()
    mac = mac_REGISTRY["basic_mac"](scheme, groups, args)
    learner = le_REGISTRY["q_learner"](mac, scheme, None, args)
    learner.target_mixer.load_state_dict({k: th.from_numpy(v) for k, v in g.group("init/target_mixer").items()})
    learner.load_models(os.path.join(REPO, "tests", "golden", "ckpt_ref"))
    for tag, mod in (("agent", learner.mac.agent), ("target_agent", learner.target_mac.agent), ("mixer", learner.mixer),
                     ("target_mixer", learner.target_mixer)):
        for k, v in g.group("loaded/" + tag).items():
            np.testing.assert_array_equal(mod.state_dict()[k].numpy(), v, err_msg=tag + "." + k)
    flat = np.concatenate([sq.numpy().ravel() for sq in learner.optimiser.square_avg])
    np.testing.assert_array_equal(flat, g["loaded/square_avg_flat"])
    assert learner.optimiser.step_count == 2


def test_host_stream_plan_decides_from_batch_location_and_size():
    """QLearner._host_stream_plan: only a host-resident batch of at least two chunks is streamed; device batches, zero-copy
    replay samples, graph replay and (unless opted in) data-parallel runs take the one-piece path."""
    shape = SmacShape("3m", 3, 30, 48, 9, 61)
    scheme, groups = make_scheme(shape)
    args = default_args(shape, mixer="qmix")
    mac = mac_REGISTRY["basic_mac"](scheme, groups, args)
    learner = le_REGISTRY["q_learner"](mac, scheme, None, args)

    class HostBatch(dict):
        pass

    def batch(n):
        b = HostBatch(obs=th.zeros(n, 2, 3, 30))
        b.batch_size = n
        return b

    assert learner._host_stream_plan(batch(1024), False) == (0, 1024, 512)
    assert learner._host_stream_plan(batch(1023), False) is None
    args.host_stream_chunk = 8
    assert learner._host_stream_plan(batch(38), False) == (0, 38, 8)
    assert learner._host_stream_plan(batch(15), False) is None
    assert learner._host_stream_plan(batch(38), True) is None              # data parallel: opt-in only
    args.cuda_graph = True
    assert learner._host_stream_plan(batch(38), False) is None
    args.cuda_graph = False
    args.host_stream_chunk = 0
    assert learner._host_stream_plan(batch(38), False) is None
    indexed = batch(38)
    indexed.ep_ids, indexed.buffer = th.arange(38), object()
    args.host_stream_chunk = 8
    assert learner._host_stream_plan(indexed, False) is None
