"""Smallest end-to-end case for compute-sanitizer: one bf16-tier learner step (QMIX, ragged) + one rollout step."""
import os, sys, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch as th
from cuda_utils import build_learner, to_batch
from oracle import qlearner_oracle as orc
from pymarl_b200.synthetic import SMAC_SHAPES, numpy_episode_fields, default_args
name = sys.argv[1] if len(sys.argv) > 1 else "2s3z"
shape = SMAC_SHAPES[name]
args = default_args(shape, mixer="qmix", learner_log_interval=0, precision="bf16", action_rng="philox")
rng = np.random.default_rng(0)
d_in = shape.obs_dim + shape.n_actions + shape.n_agents
agent = orc.init_params(orc.agent_param_shapes(d_in, 64, shape.n_actions), rng)
mixer = orc.init_params(orc.qmix_param_shapes(shape.state_dim, shape.n_agents, 32), rng)
learner, _ = build_learner(shape, args, agent, agent, mixer, mixer)
fields = numpy_episode_fields(shape, 7, 6, seed=1, ragged=True)
batch = to_batch(shape, fields)
learner.train(batch, 0, 0)
th.cuda.synchronize()
print("train ok", learner.stats()["loss"])
learner.mac.init_hidden(7)
a = learner.mac.select_actions(batch, 1, 0)
th.cuda.synchronize()
print("rollout ok", a.shape)
