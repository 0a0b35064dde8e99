"""Randomised shape sweep of the learner step against the oracle (both tiers): catches shape-dependent bugs
(tails, alignment, fold / no-fold, generic vs lean kernels).  usage: python tools/shape_sweep.py [n_cases] [seed]"""
import os, sys, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch as th
from cuda_utils import build_learner, to_batch
from oracle import qlearner_oracle as orc
from pymarl_b200.synthetic import SmacShape, numpy_episode_fields, default_args

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 12
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = 0
EDGE = [(1, 64, 63, 2, 2, 1, "qmix"), (29, 320, 127, 64, 3, 5, "qmix"), (30, 290, 64, 33, 4, 9, "qmix"),
        (64, 200, 191, 5, 3, 3, "qmix"), (2, 3, 2, 2, 2, 130, "vdn"), (27, 285, 1170, 36, 3, 10, "qmix"),
        (13, 128, 128, 16, 5, 20, None), (5, 321, 40, 7, 4, 6, "qmix")]
for case in range(n_cases + 2 * len(EDGE)):
    if case < 2 * len(EDGE):
        N, O, S, A, T, B, mixer = EDGE[case // 2]
    else:
        N = int(rng.integers(1, 40)); O = int(rng.integers(3, 321)); S = int(rng.integers(2, 400)); A = int(rng.integers(2, 41))
        T = int(rng.integers(2, 14)); B = int(rng.integers(1, 40))
        mixer = [None, "vdn", "qmix"][int(rng.integers(0, 3))]
    prec = ["fp32", "bf16"][case % 2]
    shape = SmacShape("rnd", N, O, S, A, T)
    args = default_args(shape, mixer=mixer, learner_log_interval=0, precision=prec)
    prm = np.random.default_rng(case)
    agent = orc.init_params(orc.agent_param_shapes(O + A + N, 64, A), prm)
    mix = orc.init_params(orc.qmix_param_shapes(S, N, 32), prm) if mixer == "qmix" else {}
    olr = orc.OracleQLearner(agent, mix, copy.copy(args))
    learner, _ = build_learner(shape, args, agent, agent, mix, mix)
    for sq in list(olr.sq_agent.values()) + list(olr.sq_mixer.values()):
        sq[...] = 1e-2
    learner._flat["sq"].fill_(1e-2)
    fields = numpy_episode_fields(shape, B, T, seed=case, ragged=True)
    ostats, _, _ = olr.train(fields, 0, 0)
    learner.train(to_batch(shape, fields), 0, 0)
    st = learner.stats()
    tol = 1e-2 if prec == "bf16" else 2e-5
    errs = {k: abs(st[k] - ostats[k]) / max(1.0, abs(ostats[k])) for k in ("loss", "grad_norm", "td_error_abs", "q_taken_mean", "target_mean")}
    perr = 0.0
    for kind, mod, ref in (("agent", learner.mac.agent, olr.agent), ("mixer", learner.mixer, olr.mixer_p)):
        if mod is None:
            continue
        for name, p in mod.named_parameters():
            r = ref[name]
            perr = max(perr, float(np.abs(p.detach().cpu().numpy() - r).max() / max(np.abs(r).max(), 1e-12)))
    ok = all(np.isfinite(v) and v <= (2 * tol) for v in errs.values()) and perr <= (tol if prec == "bf16" else 1e-4)
    bad += 0 if ok else 1
    print("%s %-4s N=%-2d O=%-3d S=%-3d A=%-2d T=%-2d B=%-2d mixer=%-4s  stats %.1e  params %.1e" %
          ("ok  " if ok else "FAIL", prec, N, O, S, A, T, B, mixer, max(errs.values()), perr), flush=True)
print("failures:", bad)
sys.exit(1 if bad else 0)
