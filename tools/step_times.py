"""Per-step CUDA-event times of the learner step (27m_vs_30m / 4096 by default) + per-phase times of several steps."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch as th
from bench import _DictBatch
from cuda_utils import Logger
from pymarl_b200 import le_REGISTRY, mac_REGISTRY, _lib
from pymarl_b200.synthetic import BASELINE_CONFIGS, SMAC_SHAPES, default_args, make_scheme, torch_episode_fields
cfg = BASELINE_CONFIGS["27m_vs_30m"]; shape = SMAC_SHAPES[cfg["shape"]]
B = int(sys.argv[1]) if len(sys.argv) > 1 else cfg["batch"]; T = cfg["T"]
args = default_args(shape, mixer="qmix", device="cuda", use_cuda=True, learner_log_interval=10 ** 12, precision="bf16")
th.manual_seed(7)
scheme, groups = make_scheme(shape)
scheme["actions_onehot"] = {"vshape": (shape.n_actions,), "dtype": th.float32, "group": "agents"}
mac = mac_REGISTRY["basic_mac"](scheme, groups, args)
learner = le_REGISTRY["q_learner"](mac, scheme, Logger(), args); learner.cuda()
fields = torch_episode_fields(shape, B, T, seed=1000, ragged=False, device="cuda", with_onehot=False)
batch = _DictBatch(fields, B, T)
for i in range(3): learner.train(batch, i, 0)
th.cuda.synchronize()
K = 12
evs = [th.cuda.Event(enable_timing=True) for _ in range(K + 1)]
evs[0].record()
for i in range(K):
    learner.train(batch, i, 0); evs[i + 1].record()
th.cuda.synchronize()
print("per-step ms:", ["%.2f" % evs[i].elapsed_time(evs[i + 1]) for i in range(K)])
tot = {}
for rep in range(4):
    _lib.profile_begin(); learner.train(batch, 0, 0); ph = _lib.profile_end(); th.cuda.synchronize()
    print("profiled step sum %.2f" % sum(p for n, p in ph if n != "end"))
