"""Per-kernel device times and inter-kernel gaps of the select_actions step (torch.profiler / CUPTI, no ncu replay)."""
import sys, os
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, "tests"))
import torch as th
from torch.profiler import profile, ProfilerActivity
import bench
from pymarl_b200 import mac_REGISTRY
from pymarl_b200.synthetic import make_scheme, torch_episode_fields
shape = bench.SMAC_SHAPES["27m_vs_30m"]
envs = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
dev = th.device("cuda", 0)
args = bench.default_args(shape, mixer="qmix", device="cuda", use_cuda=True, precision="bf16", action_rng="philox")
th.manual_seed(7)
scheme, groups = make_scheme(shape)
scheme["actions_onehot"] = {"vshape": (shape.n_actions,), "dtype": th.float32, "group": "agents"}
mac = mac_REGISTRY["basic_mac"](scheme, groups, args)
mac.cuda()
fields = torch_episode_fields(shape, envs, 4, seed=1000, ragged=False, device=dev, with_onehot=False)
batch = bench._DictBatch(fields, envs, 4)
mac.init_hidden(envs)
for i in range(30):
    mac.select_actions(batch, 1 + i % 3, 1000 * i)
th.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(40):
        mac.select_actions(batch, 1 + i % 3, 1000 * i)
    th.cuda.synchronize()
evs = sorted([e for e in prof.events() if e.device_type == th.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
evs = evs[len(evs) // 4:]
tot = {}
gaps = {}
for a, b in zip(evs[:-1], evs[1:]):
    tot.setdefault(a.name[:40], []).append(a.time_range.end - a.time_range.start)
    gaps.setdefault(a.name[:24] + " -> " + b.name[:24], []).append(b.time_range.start - a.time_range.end)
for k, v in tot.items():
    print("kernel %-42s n=%3d  mean %.1f us  min %.1f" % (k, len(v), sum(v) / len(v), min(v)))
for k, v in gaps.items():
    print("gap    %-56s n=%3d  mean %.1f us" % (k, len(v), sum(v) / len(v)))
