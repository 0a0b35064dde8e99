"""N select_actions steps at the bench shape (16384 envs x 27 agents): a short target for ncu."""
import sys, os
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, "tests"))
import torch as th
import bench
from pymarl_b200 import mac_REGISTRY
from pymarl_b200.synthetic import make_scheme, torch_episode_fields
shape = bench.SMAC_SHAPES["27m_vs_30m"]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
envs = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
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
for i in range(n):
    a = mac.select_actions(batch, 1 + i % 3, 1000 * i)
th.cuda.synchronize()
print("ok", int(a.sum()))
