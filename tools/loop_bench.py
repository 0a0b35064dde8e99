"""sample() + train() per iteration with the replay buffer in HBM (27m_vs_30m shapes): gathered copy vs zero-copy ids."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch as th
from cuda_utils import Logger
from pymarl_b200 import le_REGISTRY, mac_REGISTRY, ReplayBuffer
from pymarl_b200.components.transforms import OneHot
from pymarl_b200.synthetic import SMAC_SHAPES, default_args, make_scheme, torch_episode_fields
shape = SMAC_SHAPES["27m_vs_30m"]
n_buf, B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 5000, int(sys.argv[2]) if len(sys.argv) > 2 else 4096, 180
args = default_args(shape, mixer="qmix", device="cuda", use_cuda=True, learner_log_interval=10 ** 12, precision="bf16")
th.manual_seed(7)
scheme, groups = make_scheme(shape)
buf = ReplayBuffer(scheme, groups, n_buf, T, preprocess=None, device="cuda")
for b0 in range(0, n_buf, 500):
    f = torch_episode_fields(shape, min(500, n_buf - b0), T, seed=b0, ragged=True, device="cuda", with_onehot=False)
    for k, v in f.items():
        buf.data.transition_data[k][b0:b0 + v.shape[0]] = v
    del f
buf.buffer_index, buf.episodes_in_buffer = 0, n_buf
scheme2 = dict(scheme); scheme2["actions_onehot"] = {"vshape": (shape.n_actions,), "dtype": th.float32, "group": "agents"}
mac = mac_REGISTRY["basic_mac"](scheme2, groups, args)
learner = le_REGISTRY["q_learner"](mac, scheme2, Logger(), args); learner.cuda()
for zero_copy in (False, True):
    buf.zero_copy = zero_copy
    def it(i):
        batch = buf.sample(B)
        max_t = int(batch.max_t_filled())            # D2H sync, as in the reference loop (run.py:211)
        batch = batch[:, :max_t]
        learner.train(batch, i, 0)
    for i in range(3): it(i)
    th.cuda.synchronize(); t0 = time.perf_counter()
    K = 8
    for i in range(K): it(i)
    th.cuda.synchronize(); dt = (time.perf_counter() - t0) / K
    print("zero_copy=%s: %.2f ms per sample+train iteration (%d of %d episodes) -> %.0f episodes/s"
          % (zero_copy, dt * 1e3, B, n_buf, B / dt))
