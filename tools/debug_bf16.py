import copy, sys, os
import numpy as np, torch as th
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from cuda_utils import build_learner, to_batch, state_np
from test_cuda_parity import _oracle_learner
from pymarl_b200.synthetic import SMAC_SHAPES, numpy_episode_fields, default_args
shape = SMAC_SHAPES["27m_vs_30m"]; B, T = 8, 12
args = default_args(shape, mixer="qmix", learner_log_interval=0, precision="bf16", grad_norm_clip=1e30)
fields = numpy_episode_fields(shape, B, T, seed=21, ragged=True)
olr = _oracle_learner(shape, copy.copy(args), seed=8)
learner, _ = build_learner(shape, args, olr.agent, olr.target_agent, olr.mixer_p, olr.target_mixer_p)
for sq in list(olr.sq_agent.values()) + list(olr.sq_mixer.values()): sq[...] = 1e-2
learner._flat["sq"].fill_(1e-2)
p0 = {k: v.copy() for k, v in olr.agent.items()}
stats, raw_grads, fw = olr.train(fields, 0, 0)
learner.train(to_batch(shape, fields), 0, 0)
g = dict(learner.mac.agent.named_parameters())["fc1.weight"].grad.cpu().numpy()
ref = raw_grads["agent.fc1.weight"]
err = np.abs(g - ref)
print("grad max", np.abs(ref).max(), "err max", err.max(), "at", np.unravel_index(err.argmax(), err.shape))
O = shape.obs_dim
for name, sl in (("obs<256", slice(0, 256)), ("obs>=256", slice(256, O)), ("act", slice(O, O + 36)), ("id", slice(O + 36, O + 63))):
    print(name, "ref max %.4g  err max %.4g  l2 %.4g" % (np.abs(ref[:, sl]).max(), err[:, sl].max(),
          np.linalg.norm(g[:, sl] - ref[:, sl]) / np.linalg.norm(ref[:, sl])))
pn = state_np(learner.mac.agent)["fc1.weight"]; po = olr.agent["fc1.weight"]
perr = np.abs(pn - po)
print("param max", np.abs(po).max(), "err max", perr.max(), "at", np.unravel_index(perr.argmax(), perr.shape))
i, j = np.unravel_index(perr.argmax(), perr.shape)
print("p0 %.6f  oracle %.6f  mine %.6f  g_ref %.6g g_mine %.6g" % (p0["fc1.weight"][i, j], po[i, j], pn[i, j], ref[i, j], g[i, j]))
print("stats", learner.stats(), stats)
