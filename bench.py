#!/usr/bin/env python
"""bench.py - QMIX learner episodes/sec on SMAC-shaped synthetic episodes (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 27m_vs_30m] [--impl reference]

One "step" is one QLearner.train call over one batch (both agent unrolls, double-Q targets,
both mixers, masked TD loss, backward, clip, RMSprop, target sync bookkeeping).  At N GPUs
every rank trains on its own `batch` episodes (weak scaling) and the gradients are all-reduced
once per step over NCCL.  Prints ONE JSON line (rank 0).

value        whole-job episodes/s with the batch resident in HBM when the timed region starts
e2e          the same metric through the public API with the batch in pinned HOST memory:
             H2D copy of every field the step reads + D2H read of the loss inside the region
roofline     dominant kernel of the step: algorithmic FLOPs (or bytes) / its CUDA-event time,
             against MEASURED_PEAKS.json
cpu_baseline the numpy oracle (port of the reference learner) on this box's host cores, on a
             bounded sample of the same workload
--impl reference   times that CPU port alone (the reference is Python/torch and does not
             travel to the GPU box; the oracle is its validated restatement)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

from pymarl_b200.synthetic import BASELINE_CONFIGS, SMAC_SHAPES, default_args  # noqa: E402

METRIC = "qmix_learner_episodes_per_sec"
UNIT = "episodes/s"
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def load_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        d["source"] = "measured"
        return d
    d = dict(FALLBACK_PEAKS)
    d["source"] = "fallback"
    return d


# ------------------------------------------------------------------------------------------
# algorithmic work per phase: (FLOPs, HBM bytes) of one launch
# ------------------------------------------------------------------------------------------
def phase_work(name, B, T, N, O, S, A, H, E, mixer):
    """(FLOPs, HBM bytes) one launch of a phase has to do.  bf16-tier phases end in _tc:
    their activations are bf16 tile images (128 B per row of 64 columns)."""
    rows = B * T * N
    M = B * (T - 1)
    BT = B * T
    C = (N + 3) * E
    f4 = 4
    ko = -(-O // 64) * 64                      # obs image width (padded to 64-column chunks)
    ks = -(-(S + 1) // 64) * 64                # state image width
    kc = -(-(N + 3) // 2) * 64                 # raw (hypernet output) image width
    if name == "fc1_fwd_both_tc":              # obs fp32 in; obs image, x of both nets, relu mask out
        return 2.0 * rows * O * 2 * H, rows * (O * f4 + ko * 2 + 2 * 128 + 8)
    if name == "gru_unroll_fwd_online_tc":     # x in; h + 4 gate tiles out
        return 2.0 * rows * H * 6 * H, rows * (128 + 128 + 512)
    if name == "gru_unroll_fwd_target_tc":
        return 2.0 * rows * H * 6 * H, rows * (128 + 128)
    if name == "q_select_tc":                  # h of both nets, avail, actions in; chosen / tmax out
        return 2.0 * rows * H * 2 * A, rows * (2 * 128 + A * f4 + 8 + 8)
    if name == "state_to_images":
        return 0.0, BT * (S * f4 + ks * 2)
    if name == "mixer_fwd_target_tc":
        return 2.0 * BT * ks * C, BT * (ks * 2 + N * f4 + f4)
    if name == "mixer_fwd_online_tc":
        return 2.0 * BT * ks * C, BT * (ks * 2 + kc * 2 + N * f4 + f4)
    if name == "mixer_bwd" and mixer == "qmix":   # mix backward on the raw images + weight-gradient GEMM over both images
        return 2.0 * BT * ks * C, BT * (2 * kc * 2 + kc * 2 + ks * 2 + 2 * N * f4)
    if name == "gru_unroll_bwd_tc":            # 4 gate tiles + h + x in; dpre1 out; all rnn.* gradients in TMEM
        return 2.0 * rows * H * (6 * H + 8 * H), rows * (640 + 128 + 128 + 8)
    if name == "dW_rnn_tc":                    # reduction of the per-tile partials
        return 0.0, -(-(B * N) // 128) * (2 * 192 * 64 + 256) * f4
    if name == "dW_fc1_fc2_tc":
        return 2.0 * rows * H * (ko + H + 2 * H), rows * (128 + 128 + ko * 2 + 16)
    if name.startswith("fc1_fwd"):
        return 2.0 * rows * O * H, rows * (O + H) * f4
    if name == "gru_unroll_fwd_online":
        return 2.0 * rows * H * (6 * H + A), rows * (H + H + 4 * H + A) * f4
    if name == "gru_unroll_fwd_target":
        return 2.0 * rows * H * (6 * H + A), rows * (H + A) * f4
    if name == "target_select":
        return 0.0, M * N * (3 * A * f4 + 8 + 8)
    if name.startswith("mixer_fwd"):
        if mixer != "qmix":
            return 0.0, M * (N + 1) * f4
        return 2.0 * M * S * C, M * (S + 2 * C + N + 1) * f4
    if name == "td_loss":
        return 0.0, M * 5 * f4
    if name == "mixer_bwd":
        if mixer != "qmix":
            return 0.0, M * (N + 1) * f4
        return 2.0 * M * S * C, M * (S + 3 * C + 2 * N) * f4
    if name == "gru_unroll_bwd":
        return 2.0 * rows * H * 6 * H, rows * (4 * H * 2 + 3 * H + H) * f4
    if name == "dW_rnn_gemm_atb":
        return 2.0 * rows * H * 6 * H, rows * (2 * 4 * H + 2 * H) * f4
    if name == "dW_fc1_gemm_atb":
        return 2.0 * rows * H * O, rows * (H + O) * f4
    if name == "agent_scatter_grads":
        return 0.0, rows * (2 * H) * f4
    return 0.0, 0.0


def measured_traffic(name, rows):
    """DRAM bytes per launch of a phase's dominant kernel from the committed ncu capture (profiles/traffic.json:
    dram__bytes_read.sum + dram__bytes_write.sum per (b, t, n) row), scaled to this run's rows; None if not captured."""
    p = os.path.join(REPO, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p)).get("bytes_per_row", {})
    return d[name] * rows if name in d else None


def step_roofline(B, T, N, O, S, A, H, E, mixer, peaks):
    """algorithmic FLOPs and HBM bytes of one whole step."""
    d_in = O + A + N
    f_row = 2 * H * (d_in + 6 * H + A)
    f_agent = T * N * (4 * f_row - 2 * d_in * H)
    f_mix = (T - 1) * (3 * 2 * S * E * (N + 3) + 4 * (2 * E + 2 * N * E + 2 * E)) if mixer == "qmix" else 0
    flops = B * (f_agent + f_mix)
    byts = B * (2 * T * N * O * 4 + (2 * T * S * 4 if mixer == "qmix" else 0) + T * N * A * 4 + T * N * 8 + T * 13
                + 2 * T * N * H * 4)
    t_hbm = byts / (peaks["hbm_gbs"] * 1e9)
    t_tensor = flops / (peaks["bf16_tflops_sustained"] * 1e12)
    return flops, byts, max(t_hbm, t_tensor), ("hbm" if t_hbm >= t_tensor else "tensor")


# ------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.lines, self.proc, self.thread = [], None, None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power)}


# ------------------------------------------------------------------------------------------
def run_cpu_port(cfg, batch, steps, warmup):
    """Time the numpy oracle's train step (port of learners/q_learner.py:37-116) on the host."""
    import copy
    import numpy as np
    from oracle import qlearner_oracle as orc
    from pymarl_b200.synthetic import numpy_episode_fields
    shape = SMAC_SHAPES[cfg["shape"]]
    args = default_args(shape, mixer=cfg["mixer"])
    rng = np.random.default_rng(7)
    d_in = shape.obs_dim + shape.n_actions + shape.n_agents
    agent = orc.init_params(orc.agent_param_shapes(d_in, 64, shape.n_actions), rng)
    mixer = orc.init_params(orc.qmix_param_shapes(shape.state_dim, shape.n_agents, 32), rng) if cfg["mixer"] == "qmix" else {}
    olr = orc.OracleQLearner(agent, mixer, copy.copy(args))
    fields = numpy_episode_fields(shape, batch, cfg["T"], seed=0, ragged=False)
    for _ in range(warmup):
        olr.train(fields, 0, 0)
    times = []
    for i in range(steps):
        t0 = time.perf_counter()
        olr.train(fields, i, 0)
        times.append(time.perf_counter() - t0)
    mean = sum(times) / len(times)
    return batch / mean, mean * 1e3


def reference_arm(a, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    shape = SMAC_SHAPES[cfg["shape"]]
    batch = a.cpu_batch or (32 if shape.n_agents > 5 else cfg["batch"] if cfg["batch"] <= 64 else 64)
    steps, warmup = max(1, min(a.steps, 10)), max(1, min(a.warmup, 2))
    val, ms = run_cpu_port(cfg, batch, steps, warmup)
    cores = os.cpu_count()
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": a.config, "shape": cfg["shape"], "T": cfg["T"], "mixer": cfg["mixer"], "batch": batch,
                   "note": "reference algorithm on host cores; bounded sample of the workload"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d episodes x T=%d per step, %d steps (numpy oracle, OpenBLAS threads)" % (batch, cfg["T"], steps)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="27m_vs_30m", choices=sorted(BASELINE_CONFIGS) + ["select_actions"],
                    help="select_actions = BASELINE config 5: BasicMAC.select_actions over 16384 envs x 27 agents")
    ap.add_argument("--batch", type=int, default=0, help="episodes per GPU (default: BASELINE batch)")
    ap.add_argument("--ragged", action="store_true", help="variable-length episodes instead of full-length")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-batch", type=int, default=0)
    ap.add_argument("--action-rng", default="philox", choices=["philox", "torch"],
                    help="select_actions config: where the epsilon-greedy draws come from")
    ap.add_argument("--precision", default="bf16", choices=["fp32", "bf16"],
                    help="bf16 = tcgen05 tensor-core tier (fp32 accumulate, parity 1e-2); fp32 = CUDA-core tier (parity 1e-5)")
    a = ap.parse_args()
    if a.config == "select_actions":
        return rollout_bench(a)
    cfg = dict(BASELINE_CONFIGS[a.config])
    if a.batch:
        cfg["batch"] = a.batch
    if a.impl == "reference":
        return reference_arm(a, cfg)

    import torch as th
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(REPO, "tests"))
    from cuda_utils import Logger
    from pymarl_b200 import le_REGISTRY, mac_REGISTRY, _lib
    from pymarl_b200.synthetic import make_scheme, torch_episode_fields

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    th.cuda.set_device(local_rank)
    dev = th.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()

    shape = SMAC_SHAPES[cfg["shape"]]
    B, T = cfg["batch"], cfg["T"]
    N, O, S, A = shape.n_agents, shape.obs_dim, shape.state_dim, shape.n_actions
    H, E = 64, 32
    args = default_args(shape, mixer=cfg["mixer"], device="cuda", use_cuda=True, learner_log_interval=10 ** 12,
                        precision=a.precision)
    th.manual_seed(7)                                   # identical parameters on every rank
    scheme, groups = make_scheme(shape)
    scheme["actions_onehot"] = {"vshape": (A,), "dtype": th.float32, "group": "agents"}
    mac = mac_REGISTRY["basic_mac"](scheme, groups, args)
    learner = le_REGISTRY["q_learner"](mac, scheme, Logger(), args)
    learner.cuda()

    fields = torch_episode_fields(shape, B, T, seed=1000 + rank, ragged=a.ragged, device=dev, with_onehot=False)
    batch = _DictBatch(fields, B, T)
    input_bytes = sum(v.numel() * v.element_size() for v in fields.values())

    def barrier():
        if world > 1:
            dist.barrier()
        th.cuda.synchronize()

    for i in range(a.warmup):
        learner.train(batch, i, 0)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = _lib.lib().pmb_launch_count()
    ev0, ev1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(a.steps):
        learner.train(batch, a.warmup + i, 0)
    ev1.record()
    barrier()
    clocks = sampler.stop()
    launches = (_lib.lib().pmb_launch_count() - launches0) // max(1, a.steps)
    ms = ev0.elapsed_time(ev1) / a.steps
    t = th.tensor([ms], dtype=th.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B / (ms * 1e-3)

    # per-kernel times of one more step (CUDA events between the launches, same stream)
    _lib.profile_begin()
    learner.train(batch, 0, 0)
    phases = _lib.profile_end()
    th.cuda.synchronize()
    phase_ms = {}
    for name, pms in phases:
        phase_ms[name] = phase_ms.get(name, 0.0) + pms
    phase_ms.pop("end", None)
    top = max(phase_ms, key=phase_ms.get) if phase_ms else None
    roofline = None
    if top:
        fl, by = phase_work(top, B, T, N, O, S, A, H, E, cfg["mixer"])
        tsec = phase_ms[top] * 1e-3
        t_h = by / (peaks["hbm_gbs"] * 1e9)
        t_t = fl / (peaks["bf16_tflops_sustained"] * 1e12)
        if t_t >= t_h:
            ach, peak, unit, bound = fl / tsec / 1e12, peaks["bf16_tflops_sustained"], "TFLOP/s", "tensor"
        else:
            ach, peak, unit, bound = by / tsec / 1e9, peaks["hbm_gbs"], "GB/s", "hbm"
        roofline = {"kernel": top, "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
                    "traffic": measured_traffic(top, B * T * N), "algorithmic_bytes": by, "algorithmic_flops": fl,
                    "peak_source": peaks["source"], "kernel_ms": phase_ms[top],
                    "share_of_step": phase_ms[top] / sum(phase_ms.values())}
    fl_s, by_s, t_roof, bound_s = step_roofline(B, T, N, O, S, A, H, E, cfg["mixer"], peaks)

    # ---- e2e: host-resident batch through the public API ------------------------------------
    e2e = None
    if not a.no_e2e:
        try:
            # every rank pins its whole batch: refuse (instead of risking the box) when the host cannot hold it
            import psutil
            need = world * input_bytes
            avail = psutil.virtual_memory().available
            if avail < 1.25 * need:
                raise MemoryError("host has %.0f GB available, the pinned batches of %d ranks need %.0f GB"
                                  % (avail / 1e9, world, need / 1e9))
            host = {k: th.empty(v.shape, dtype=v.dtype, pin_memory=True).copy_(v) for k, v in fields.items()}
            hb = _DictBatch(host, B, T)
            del fields, batch
            th.cuda.empty_cache()
            learner.train(hb, 0, 0)                          # warm-up (allocates the device copies)
            barrier()
            t0 = time.perf_counter()
            for i in range(a.e2e_steps):
                learner.train(hb, i, 0)
                loss = learner.last_stats[6].item()          # D2H read of the step's loss
            barrier()
            dt = (time.perf_counter() - t0) / a.e2e_steps
            tt = th.tensor([dt], dtype=th.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            h2d = sum(v.numel() * v.element_size() for k, v in host.items()
                      if k in ("obs", "state", "actions", "avail_actions", "reward", "terminated", "filled")
                      and (k != "state" or cfg["mixer"] == "qmix"))
            e2e = {"value": world * B / float(tt.item()), "unit": UNIT, "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": 8, "ms_per_step": float(tt.item()) * 1e3, "steps": a.e2e_steps, "loss": loss}
        except Exception as ex:                              # e.g. the pod cannot pin 30 GB
            e2e = {"value": None, "unit": UNIT, "error": repr(ex)[:200]}

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cb = a.cpu_batch or (32 if N > 5 else min(B, 64))
        val, cms = run_cpu_port(cfg, cb, 5, 1)
        cpu = {"value": val, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
               "sample": "%d episodes x T=%d per step, 5 steps (numpy oracle, OpenBLAS threads)" % (cb, T),
               "ms_per_step": cms}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if a.precision == "fp32" else "bf16 (fp32 accumulate)", "data": "synthetic",
            "config": {"workload": "QMIX learner step, %s shapes" % cfg["shape"] if cfg["mixer"] == "qmix" else
                       "%s learner step, %s shapes" % (cfg["mixer"] or "iql", cfg["shape"]),
                       "name": a.config, "n_agents": N, "obs": O, "state": S, "n_actions": A, "T": T,
                       "batch_per_gpu": B, "global_batch": world * B, "mixer": cfg["mixer"], "episodes": "ragged" if a.ragged else "full-length",
                       "parallelism": "dp%d" % world, "l2_policy": "inputs (%.1f GB) exceed L2" % (input_bytes / 1e9)},
            "roofline": roofline,
            "step_roofline": {"flops": fl_s, "bytes": by_s, "t_roof_ms": t_roof * 1e3, "bound": bound_s,
                              "frac": t_roof * 1e3 / ms},
            "phases_ms": phase_ms,
            "cpu_baseline": cpu, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def rollout_bench(a):
    """BASELINE config 5: one BasicMAC.select_actions step (fc1 -> GRUCell -> fc2 -> avail mask -> epsilon-greedy) over
    `--batch` (default 16384) synthetic envs x 27 agents; metric = agent-steps/s.  The obs of 4 timesteps live in HBM and
    the step index cycles over them (2 GB, larger than L2)."""
    import numpy as np
    shape = SMAC_SHAPES["27m_vs_30m"]
    N, O, A, H = shape.n_agents, shape.obs_dim, shape.n_actions, 64
    envs = a.batch or 16384
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    metric, unit = "mac_agent_steps_per_sec", "agent-steps/s"
    if a.impl == "reference":
        if rank != 0:
            return
        import copy
        from oracle import qlearner_oracle as orc
        from pymarl_b200.synthetic import numpy_episode_fields
        cb = a.cpu_batch or 256
        rng = np.random.default_rng(7)
        agent = orc.init_params(orc.agent_param_shapes(O + A + N, H, A), rng)
        fields = numpy_episode_fields(shape, cb, 4, seed=0, ragged=False)
        if "actions_onehot" not in fields:
            fields["actions_onehot"] = np.eye(A, dtype=np.float32)[fields["actions"][..., 0]]
        hstate = np.zeros((cb * N, H), np.float32)
        steps, warm = max(1, min(a.steps, 20)), max(1, min(a.warmup, 3))
        times = []
        for i in range(warm + steps):
            t0 = time.perf_counter()
            t = 1 + i % 3
            u = rng.random((cb, N), dtype=np.float32)
            e = rng.exponential(size=(cb, N, A)).astype(np.float32)
            _, _, hstate = orc.mac_select_actions(agent, fields, t, hstate, 0.5, u, e)
            if i >= warm:
                times.append(time.perf_counter() - t0)
        ms = 1e3 * sum(times) / len(times)
        val = cb * N / (ms * 1e-3)
        line = {"impl": "reference", "metric": metric, "value": val, "unit": unit, "n_gpus": a.gpus, "steps": steps,
                "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": "BasicMAC.select_actions step, 27m_vs_30m shapes", "envs": cb, "n_agents": N},
                "cpu_baseline": {"value": val, "unit": unit, "cores": os.cpu_count(), "kind": "port",
                                 "sample": "%d envs x %d agents per step, %d steps (numpy oracle)" % (cb, N, steps)},
                "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return

    import torch as th
    import torch.distributed as dist
    from pymarl_b200 import mac_REGISTRY, _lib
    from pymarl_b200.synthetic import make_scheme, torch_episode_fields
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    th.cuda.set_device(local_rank)
    dev = th.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    # action_rng="philox": the epsilon-greedy draws come from the kernel's own Philox stream (one fused call per step);
    # the default "torch" mode replays the reference's generator order for bit-identical actions and costs two torch
    # RNG launches + 190 MB of extra traffic per step
    args = default_args(shape, mixer="qmix", device="cuda", use_cuda=True, precision=a.precision, action_rng=a.action_rng)
    th.manual_seed(7)
    scheme, groups = make_scheme(shape)
    scheme["actions_onehot"] = {"vshape": (A,), "dtype": th.float32, "group": "agents"}
    mac = mac_REGISTRY["basic_mac"](scheme, groups, args)
    mac.cuda()
    Tb = 4
    fields = torch_episode_fields(shape, envs, Tb, seed=1000 + rank, ragged=False, device=dev, with_onehot=False)
    batch = _DictBatch(fields, envs, Tb)
    mac.init_hidden(envs)

    def barrier():
        if world > 1:
            dist.barrier()
        th.cuda.synchronize()

    for i in range(a.warmup):
        mac.select_actions(batch, 1 + i % 3, 1000 * i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = _lib.lib().pmb_launch_count()
    ev0, ev1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(a.steps):
        mac.select_actions(batch, 1 + i % 3, 1000 * i)
    ev1.record()
    barrier()
    clocks = sampler.stop()
    launches = (_lib.lib().pmb_launch_count() - launches0) // max(1, a.steps)
    ms = ev0.elapsed_time(ev1) / a.steps
    t = th.tensor([ms], dtype=th.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * envs * N / (ms * 1e-3)
    rows = envs * N
    algo_bytes = rows * (O * 4 + 2 * H * 4 + A * 4 + 8 + 8)
    ach = algo_bytes / (ms * 1e-3) / 1e9
    # e2e: obs / avail of the step start in pinned host memory, the chosen actions are read back (what a runner does)
    e2e = None
    if not a.no_e2e:
        host = {k: th.empty(v.shape, dtype=v.dtype, pin_memory=True).copy_(v) for k, v in fields.items()}
        hb = _DictBatch(host, envs, Tb)
        out = th.empty(envs, N, dtype=th.int64, pin_memory=True)
        mac.select_actions(hb, 1, 0)
        barrier()
        t0 = time.perf_counter()
        for i in range(a.e2e_steps):
            acts = mac.select_actions(hb, 1 + i % 3, 1000 * i)
            out.copy_(acts, non_blocking=True)
            th.cuda.synchronize()
        dt = (time.perf_counter() - t0) / a.e2e_steps
        # obs / avail_actions: the step itself; actions / filled: the step and the one before (last-action input)
        h2d = sum(host[k][:, :1].numel() * host[k].element_size() for k in ("obs", "avail_actions")) + \
            sum(host[k][:, :2].numel() * host[k].element_size() for k in ("actions", "filled"))
        e2e = {"value": world * envs * N / dt, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": envs * N * 8,
               "ms_per_step": dt * 1e3, "steps": a.e2e_steps}
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        sub = subprocess.run([sys.executable, os.path.abspath(__file__), "--config", "select_actions", "--impl", "reference",
                              "--steps", "10", "--warmup", "2"], capture_output=True, text=True)
        try:
            cpu = json.loads(sub.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception:
            cpu = {"value": None, "error": (sub.stderr or "")[-200:]}
    if rank == 0:
        line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32" if a.precision == "fp32" else "bf16 (fp32 accumulate)", "data": "synthetic",
                "config": {"workload": "BasicMAC.select_actions step, 27m_vs_30m shapes", "envs_per_gpu": envs, "n_agents": N,
                           "obs": O, "n_actions": A, "parallelism": "dp%d" % world, "action_rng": a.action_rng,
                           "l2_policy": "obs of 4 timesteps (%.1f GB) cycled, larger than L2" % (fields["obs"].numel() * 4 / 1e9)},
                "roofline": {"kernel": "select_actions_step (all launches)", "bound": "hbm", "achieved": ach,
                             "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "traffic": None,
                             "algorithmic_bytes": algo_bytes, "peak_source": peaks["source"]},
                "cpu_baseline": cpu, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


class _DictBatch:
    """Minimal EpisodeBatch stand-in (QLearner.train only reads batch[key])."""

    def __init__(self, fields, batch_size, max_seq_length):
        self.fields, self.batch_size, self.max_seq_length = fields, batch_size, max_seq_length
        self.device = next(iter(fields.values())).device

    def __getitem__(self, k):
        return self.fields[k]


if __name__ == "__main__":
    main()
