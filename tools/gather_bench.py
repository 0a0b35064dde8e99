"""Throughput of the replay gather (pmb_gather_episodes) at 27m_vs_30m shapes: sample 1024 of 2048 episodes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch as th
from pymarl_b200 import _lib
from pymarl_b200.synthetic import SMAC_SHAPES, torch_episode_fields
shape = SMAC_SHAPES["27m_vs_30m"]
n_buf, n = 2048, 1024
f = torch_episode_fields(shape, n_buf, 180, seed=0, device="cuda", with_onehot=False)
ids = np.random.default_rng(0).choice(n_buf, n, replace=False)
out = _lib.gather_episodes(f, ids, n_buf)
out2 = _lib.gather_episodes(f, ids, n_buf)      # two output sets alive -> the allocator caches blocks for both
del out2
th.cuda.synchronize()
ev0, ev1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(5):
    out = _lib.gather_episodes(f, ids, n_buf)
ev1.record(); th.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / 5
byts = sum(v[0].numel() * v.element_size() for v in f.values()) * n
ref = {k: v[th.as_tensor(ids, device="cuda")] for k, v in f.items()}
ok = all(th.equal(out[k], ref[k]) for k in f)
print("gather %d episodes: %.3f ms, %.1f GB moved (read+write), %.0f GB/s, equal=%s" % (n, ms, 2 * byts / 1e9, 2 * byts / ms / 1e6, ok))
idt = th.as_tensor(ids, device="cuda")
for name, fn in (("torch advanced indexing", lambda: {k: v[idt] for k, v in f.items()}),
                 ("torch contiguous copy of the first n", lambda: {k: v[:n].clone() for k, v in f.items()})):
    fn(); th.cuda.synchronize()
    ev0.record()
    for _ in range(5): r = fn()
    ev1.record(); th.cuda.synchronize()
    ms2 = ev0.elapsed_time(ev1) / 5
    print("%s: %.3f ms, %.0f GB/s" % (name, ms2, 2 * byts / ms2 / 1e6))
