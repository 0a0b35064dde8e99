"""Helpers shared by the tests: load the golden fixtures written by tests/golden/make_golden.py."""
import ast
import os
from types import SimpleNamespace

import numpy as np

from pymarl_b200.synthetic import SmacShape, default_args

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LEARNER_CASES = ["qmix_tiny", "vdn_tiny", "iql_tiny", "qmix_nodouble_tiny", "qmix_noid_tiny", "qmix_3m"]


class Golden:
    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.meta = ast.literal_eval(str(self.z["meta"])) if "meta" in self.z.files else {}

    def __getitem__(self, k):
        return self.z[k]

    def has(self, k):
        return k in self.z.files

    def group(self, prefix):
        prefix = prefix.rstrip("/") + "/"
        return {k[len(prefix):]: self.z[k] for k in self.z.files if k.startswith(prefix)}

    @property
    def shape(self):
        return SmacShape(*self.meta["shape"])

    def args(self, **over):
        m = self.meta
        kw = dict(mixer=m["mixer"], double_q=m["double_q"], learner_log_interval=0)
        kw.update(m.get("over", {}))
        kw.update(over)
        return default_args(self.shape, **kw)

    def batch_fields(self):
        return self.group("in")

    def episode_schedule(self):
        """(t_env, episode_num) per train step, as make_golden.run_case used."""
        n, s = self.meta["n_steps"], self.meta["sync_step"]
        return [(step, 200 if step == s else (201 if step > s else 0)) for step in range(n)]


def rel_err(a, b):
    """norm-wise relative error  max|a-b| / max|b|  ('÷ tensor max')."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = np.abs(b).max()
    if denom == 0:
        return float(np.abs(a - b).max())
    return float(np.abs(a - b).max() / denom)
