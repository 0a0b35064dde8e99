"""Per-role wait / phase profile of the two rollout kernels (fc1_stream in rollout mode, gru_rollout_kernel).  Needs a
library built with PMB_EXTRA_NVCC_FLAGS="-DPMB_RO_PROFILE -DPMB_FC1_PROFILE" (python -m pymarl_b200.build --force)."""
import sys, os, ctypes as C
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, "tests"))
import torch as th
import bench
from pymarl_b200 import mac_REGISTRY, _lib
from pymarl_b200.synthetic import make_scheme, torch_episode_fields
shape = bench.SMAC_SHAPES["27m_vs_30m"]
envs = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = 20
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
lib = _lib.lib()
for i in range(5):
    mac.select_actions(batch, 1 + i % 3, 1000 * i)
ro = (C.c_ulonglong * 16)()
fc = (C.c_ulonglong * 16)()
lib.pmb_debug_ro_prof(ro, 1)
lib.pmb_debug_fc1_prof(fc, 1)
ev0, ev1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
ev0.record()
for i in range(steps):
    mac.select_actions(batch, 1 + i % 3, 1000 * i)
ev1.record()
th.cuda.synchronize()
print("ms per step (instrumented build)", ev0.elapsed_time(ev1) / steps)
lib.pmb_debug_ro_prof(ro, 1)
lib.pmb_debug_fc1_prof(fc, 1)
v = list(ro)
n_cta = v[12]
tot = v[11] / n_cta
print("gru_rollout: CTA launches", n_cta, "cycles per CTA", round(tot), "tiles per CTA", envs * shape.n_agents / 128 / (n_cta / steps))
tiles = envs * shape.n_agents / 128 / (n_cta / steps)
rows = [(0, "loader: xh_free", 1), (1, "avail loader: av_free", 1), (2, "MMA thread: issuing (rest = polling)", 1),
        (3, "gate group: in_full (x8)", 8), (4, "gate group: gates_full (x8)", 8),
        (8, "gate group phase: gate math + new h staged (x8)", 8), (6, "selection group: av_full (x8)", 8),
        (10, "selection group phase: availability bits, draws (x8)", 8), (15, "selection group: in_full of the next tile (x8)", 8),
        (7, "selection group phase: h_0 -> bf16 operand of the next tile (x8)", 8), (5, "selection group: q_full (x8)", 8),
        (9, "selection group phase: q, arg-max, hand-over, pick (x8)", 8)]
for i, nm, m in rows:
    print(f"{nm:58s} {v[i] / n_cta / m / tot * 100:6.1f} % of the kernel time per warp   ({v[i] / n_cta / m / tiles:8.0f} cycles per tile)")
print("prologue (kernel entry -> past the first __syncthreads): %.0f cycles; entry -> first tile landed: %.0f cycles" % (v[13] / n_cta, v[14] / n_cta))
v = list(fc)
n_cta = v[9]
tot = v[8] / n_cta
print("fc1_stream (rollout mode): CTA launches", n_cta, "cycles per CTA", round(tot))
names = ["producer: st_empty (x3 warps)", "converter: st_full (x8 warps)", "converter: a_free (x8)", "MMA: a_full", "MMA: tempty",
         "epilogue: tfull (x4)", "converter: load + pack phase (x8)", "converter: store phase incl. a_free (x8)"]
mult = [3, 8, 8, 1, 1, 4, 8, 8]
for i, (nm, m) in enumerate(zip(names, mult)):
    print(f"{nm:48s} {v[i] / n_cta / m / tot * 100:6.1f} % of the kernel time per warp")
for i, nm in ((10, "producer phase: group descriptors (x3)"), (11, "producer phase: tables, copies, arrival (x3)")):
    print(f"{nm:48s} {v[i] / n_cta / 3 / tot * 100:6.1f} % of the kernel time per warp   ({v[i] / n_cta / (8 * tiles):6.0f} cycles per group)")
