"""Per-kernel device times and inter-kernel gaps of the QMIX learner step (torch.profiler / CUPTI, no ncu replay)."""
import sys, os
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, "tests"))
import torch as th
from torch.profiler import profile, ProfilerActivity
import bench
from cuda_utils import Logger
from pymarl_b200 import le_REGISTRY, mac_REGISTRY
from pymarl_b200.synthetic import make_scheme, torch_episode_fields
name = sys.argv[1] if len(sys.argv) > 1 else "27m_vs_30m"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
shape = bench.SMAC_SHAPES[name]
T = shape.episode_limit + 1 if hasattr(shape, "episode_limit") else 180
dev = th.device("cuda", 0)
args = bench.default_args(shape, mixer="qmix", device="cuda", use_cuda=True, learner_log_interval=10 ** 12, precision="bf16")
th.manual_seed(7)
scheme, groups = make_scheme(shape)
scheme["actions_onehot"] = {"vshape": (shape.n_actions,), "dtype": th.float32, "group": "agents"}
mac = mac_REGISTRY["basic_mac"](scheme, groups, args)
learner = le_REGISTRY["q_learner"](mac, scheme, Logger(), args)
learner.cuda()
T = int(sys.argv[3]) if len(sys.argv) > 3 else 180
fields = torch_episode_fields(shape, B, T, seed=1000, ragged=False, device=dev, with_onehot=False)
batch = bench._DictBatch(fields, B, T)
for i in range(4):
    learner.train(batch, i, 0)
th.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(6):
        learner.train(batch, 10 + i, 0)
    th.cuda.synchronize()
evs = sorted([e for e in prof.events() if e.device_type == th.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
n = len(evs) // 6
evs = evs[2 * n:]
tot, gaps, order = {}, {}, []
for a, b in zip(evs[:-1], evs[1:]):
    k = a.name[:60]
    if k not in tot: order.append(k)
    tot.setdefault(k, []).append(a.time_range.end - a.time_range.start)
    gaps.setdefault(k, []).append(b.time_range.start - a.time_range.end)
steps = 4
span = (evs[-1].time_range.end - evs[0].time_range.start) / steps
print("events per step", n, " span per step %.1f us" % span)
gsum = 0.0
for k in order:
    v, g = tot[k], gaps[k]
    gsum += sum(g) / steps
    print("%-62s n/step %4.1f  mean %9.1f us   gap after: mean %6.1f us" % (k, len(v) / steps, sum(v) / len(v), sum(g) / len(g)))
print("sum of gaps per step %.1f us" % gsum)
