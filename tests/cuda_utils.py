"""Helpers for the -m gpu parity tests: build the CUDA-backed learner / MAC from a golden
fixture or from oracle parameters, and wrap numpy episode fields in an EpisodeBatch."""
from types import SimpleNamespace

import numpy as np
import torch as th

from pymarl_b200 import le_REGISTRY, mac_REGISTRY, EpisodeBatch
from pymarl_b200.components.transforms import OneHot
from pymarl_b200.synthetic import make_scheme, get_shape


class Logger:
    def __init__(self):
        self.stats, self.infos = {}, []
        self.console_logger = SimpleNamespace(info=lambda msg, *a: self.infos.append(msg))

    def log_stat(self, key, value, t):
        self.stats.setdefault(key, []).append((t, float(value)))


def to_batch(shape, fields, device="cuda"):
    shape = get_shape(shape)
    B, T = fields["obs"].shape[:2]
    scheme, groups = make_scheme(shape)
    preprocess = {"actions": ("actions_onehot", [OneHot(out_dim=shape.n_actions)])}
    batch = EpisodeBatch(scheme, groups, B, T, preprocess=preprocess, device=device)
    for k, v in fields.items():
        t = th.from_numpy(np.ascontiguousarray(v)).to(device)
        assert batch.data.transition_data[k].shape == t.shape, (k, t.shape)
        assert batch.data.transition_data[k].dtype == t.dtype, (k, t.dtype)
        batch.data.transition_data[k] = t
    return batch


def build_learner(shape, args, agent=None, target_agent=None, mixer=None, target_mixer=None):
    """CUDA learner with the given numpy parameter dicts loaded (reference state_dict names)."""
    shape = get_shape(shape)
    scheme, groups = make_scheme(shape)
    scheme["actions_onehot"] = {"vshape": (shape.n_actions,), "dtype": th.float32, "group": "agents"}
    args.device = "cuda"
    args.use_cuda = True
    mac = mac_REGISTRY[args.mac](scheme, groups, args)
    logger = Logger()
    learner = le_REGISTRY[args.learner](mac, scheme, logger, args)
    learner.cuda()
    def load(module, params):
        if params:
            module.load_state_dict({k: th.from_numpy(np.asarray(v, dtype=np.float32)) for k, v in params.items()})
    load(learner.mac.agent, agent)
    load(learner.target_mac.agent, target_agent if target_agent is not None else agent)
    if args.mixer == "qmix":
        load(learner.mixer, mixer)
        load(learner.target_mixer, target_mixer if target_mixer is not None else mixer)
    return learner, logger


def learner_from_golden(g, **over):
    return build_learner(g.shape, g.args(**over), g.group("init/agent"), g.group("init/target_agent"),
                         g.group("init/mixer"), g.group("init/target_mixer"))


def state_np(module):
    return {k: v.detach().cpu().numpy() for k, v in module.state_dict().items()}
