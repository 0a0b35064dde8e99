"""Per-role wait profile of fc1_stream_kernel.  Needs a library built with
PMB_EXTRA_NVCC_FLAGS=-DPMB_FC1_PROFILE python -c 'from pymarl_b200.build import build; build(force=True)'."""
import sys, os, ctypes as C
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, "tests"))
import torch as th
import bench
from cuda_utils import Logger
from pymarl_b200 import le_REGISTRY, mac_REGISTRY, _lib
from pymarl_b200.synthetic import make_scheme, torch_episode_fields
shape = bench.SMAC_SHAPES["27m_vs_30m"]
B, T = 2048, 180
dev = th.device("cuda", 0)
args = bench.default_args(shape, mixer="qmix", device="cuda", use_cuda=True, learner_log_interval=10 ** 12, precision="bf16")
th.manual_seed(7)
scheme, groups = make_scheme(shape)
scheme["actions_onehot"] = {"vshape": (shape.n_actions,), "dtype": th.float32, "group": "agents"}
mac = mac_REGISTRY["basic_mac"](scheme, groups, args)
learner = le_REGISTRY["q_learner"](mac, scheme, Logger(), args)
learner.cuda()
fields = torch_episode_fields(shape, B, T, seed=1000, ragged=False, device=dev, with_onehot=False)
batch = bench._DictBatch(fields, B, T)
lib = _lib.lib()
out = (C.c_ulonglong * 16)()
for i in range(2):
    learner.train(batch, i, 0)
lib.pmb_debug_fc1_prof(out, 1)
learner.train(batch, 3, 0)
lib.pmb_debug_fc1_prof(out, 1)
v = list(out)
n_cta = v[9]
tot = v[8] / n_cta
names = ["producer: st_empty (x3 warps)", "converter: st_full (x8 warps)", "converter: a_free (x8)", "MMA: a_full", "MMA: tempty",
         "epilogue: tfull (x4)", "converter: load + pack phase (x8)", "converter: store phase incl. a_free (x8)"]
mult = [3, 8, 8, 1, 1, 4, 8, 8]
print("CTAs", n_cta, "cycles per CTA", round(tot))
for i, (nm, m) in enumerate(zip(names, mult)):
    print(f"{nm:34s} {v[i] / n_cta / m / tot * 100:6.1f} % of the kernel time per warp")
