"""Where does the time between the kernels of a step go?  Times N steps one by one (CUDA events), then prints the
per-launch phases of one profiled step with the begin->end span next to their sum.  Run on the GPU box."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch as th
import bench
from cuda_utils import Logger
from pymarl_b200 import le_REGISTRY, mac_REGISTRY, _lib
from pymarl_b200.synthetic import make_scheme, torch_episode_fields


shape = bench.SMAC_SHAPES["27m_vs_30m"]
B, T = 4096, 180
dev = th.device("cuda", 0)
args = bench.default_args(shape, mixer="qmix", device="cuda", use_cuda=True, learner_log_interval=10 ** 12,
                          precision="bf16")
th.manual_seed(7)
scheme, groups = make_scheme(shape)
scheme["actions_onehot"] = {"vshape": (shape.n_actions,), "dtype": th.float32, "group": "agents"}
mac = mac_REGISTRY["basic_mac"](scheme, groups, args)
learner = le_REGISTRY["q_learner"](mac, scheme, Logger(), args)
learner.cuda()
fields = torch_episode_fields(shape, B, T, seed=1000, ragged=False, device=dev, with_onehot=False)
batch = bench._DictBatch(fields, B, T)
for i in range(3):
    learner.train(batch, i, 0)
th.cuda.synchronize()
n = 12
evs = [th.cuda.Event(enable_timing=True) for _ in range(n + 1)]
host = []
evs[0].record()
for i in range(n):
    t0 = time.perf_counter()
    learner.train(batch, 3 + i, 0)
    host.append((time.perf_counter() - t0) * 1e3)
    evs[i + 1].record()
th.cuda.synchronize()
print("per-step device ms:", [round(evs[i].elapsed_time(evs[i + 1]), 2) for i in range(n)])
print("per-step host ms  :", [round(h, 2) for h in host])
_lib.profile_begin()
learner.train(batch, 0, 0)
ph = _lib.profile_end()
th.cuda.synchronize()
tot = 0.0
for name, ms in ph:
    print(f"  {name:34s} {ms:8.3f}")
    tot += ms
print("sum of phases", round(tot, 3))
