"""torchrun target: time the two data-parallel exchanges in isolation (27m_vs_30m QMIX gradient, 1,173,829 floats).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29611 tools/dp_exchange_bench.py

fused: pmb_dp_fused_allreduce_update (one kernel: all-reduce over NVLink peer memory + grad norm + clip + RMSprop)
nccl : pmb_dp_pack -> dist.all_reduce -> pmb_dp_unpack -> pmb_clip_rmsprop_update
Each call is preceded by a device-side barrier-equivalent (a tiny all-reduce) so the ranks start together; CUDA events,
max over ranks.  Prints one JSON line on rank 0."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
import torch.distributed as dist

from pymarl_b200 import _lib, data_parallel

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
th.cuda.set_device(local)
dev = th.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 1173829
L = _lib.lib()
px = data_parallel.PeerExchange(n, dev)
p, sq, tg = (th.zeros(n, device=dev) for _ in range(3))
stats = th.zeros(16, dtype=th.float64, device=dev)
g_nccl = th.zeros(n + _lib.DP_TAIL_FLOATS, device=dev)
scratch = th.empty(4096, device=dev)
tiny = th.zeros(1, device=dev)
s = _lib.stream_ptr(dev)


def fill(buf):
    buf[:n].normal_()
    stats[:5] = th.tensor([1000.0, 5.0, 3.0, 2.0, 1.0], dtype=th.float64)


def fused():
    px.fused_update(p, sq, tg, 0, stats, 5e-4, 0.99, 1e-5, 10.0, s)


def nccl():
    _lib.check(L.pmb_dp_pack(n, _lib.ptr(g_nccl), _lib.ptr(stats), s))
    dist.all_reduce(g_nccl)
    _lib.check(L.pmb_dp_unpack(n, _lib.ptr(g_nccl), _lib.ptr(stats), s))
    _lib.check(L.pmb_clip_rmsprop_update(n, _lib.ptr(p), _lib.ptr(g_nccl), _lib.ptr(sq), _lib.ptr(tg), 0, _lib.ptr(stats), 5e-4,
                                         0.99, 1e-5, 10.0, _lib.ptr(scratch), s))


out = {"world": world, "n_floats": n, "fused_available": px.ok}
for name, fn, buf in (("fused", fused, px.grad_view()), ("nccl", nccl, g_nccl)):
    if name == "fused" and not px.ok:
        continue
    times = []
    for i in range(60):
        fill(buf)
        dist.all_reduce(tiny)                      # align the ranks
        th.cuda.synchronize()
        e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        th.cuda.synchronize()
        if i >= 10:
            times.append(e0.elapsed_time(e1) * 1e3)
    t = th.tensor([sorted(times)[len(times) // 2]], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out[name + "_us_median"] = float(t.item())
if px.ok:
    out["error_word"] = px.error_word()
if rank == 0:
    print("DP_EXCHANGE " + json.dumps(out), flush=True)
dist.destroy_process_group()
