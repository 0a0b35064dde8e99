"""Is the select_actions bench host-bound?  Times 200 steps three ways: host enqueue time (no sync), device time (events), and
device time of the same steps replayed as ONE CUDA graph (no host work between the launches)."""
import sys, os, time
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, "tests"))
import torch as th
import bench
from pymarl_b200 import mac_REGISTRY
from pymarl_b200.synthetic import make_scheme, torch_episode_fields
shape = bench.SMAC_SHAPES["27m_vs_30m"]
envs = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = 200
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
for i in range(20):
    mac.select_actions(batch, 1 + i % 3, 1000 * i)
th.cuda.synchronize()
ev0, ev1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
ev0.record()
for i in range(steps):
    mac.select_actions(batch, 1 + i % 3, 1000 * i)
ev1.record()
t1 = time.perf_counter()
th.cuda.synchronize()
print("host enqueue us/step %.1f   device us/step %.1f" % ((t1 - t0) / steps * 1e6, ev0.elapsed_time(ev1) / steps * 1e3))
# graph replay of 30 steps
side = th.cuda.Stream()
side.wait_stream(th.cuda.current_stream())
with th.cuda.stream(side):
    for i in range(3):
        mac.select_actions(batch, 1 + i % 3, 1000 * i)
    g = th.cuda.CUDAGraph()
    with th.cuda.graph(g, stream=side):
        for i in range(30):
            mac.select_actions(batch, 1 + i % 3, 1000 * i)
th.cuda.current_stream().wait_stream(side)
g.replay()
th.cuda.synchronize()
ev0.record()
for _ in range(5):
    g.replay()
ev1.record()
th.cuda.synchronize()
print("graph replay device us/step %.1f" % (ev0.elapsed_time(ev1) / 150 * 1e3))
