"""World-size-2 gloo test (CPU) of the data-parallel scheme: shard the episodes, compute each shard's
UN-normalised gradient and loss sums (with the oracle as the compute), all-reduce through
pymarl_b200.data_parallel, normalise, clip, RMSprop - and compare with the single-process result."""
import copy
import os
import sys

import numpy as np
import torch as th
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _flat(grads, names):
    return np.concatenate([grads[k].ravel() for k in names]).astype(np.float64)


def _run(rank, world, port, out_dir):
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import qlearner_oracle as orc
    from pymarl_b200 import data_parallel as dp
    from pymarl_b200.synthetic import SMAC_SHAPES, numpy_episode_fields, default_args
    shape = SMAC_SHAPES["3m"]
    args = default_args(shape, mixer="qmix")
    rng = np.random.default_rng(0)
    agent = orc.init_params(orc.agent_param_shapes(42, 64, 9), rng, np.float64)
    mixer = orc.init_params(orc.qmix_param_shapes(48, 3, 32), rng, np.float64)
    fields = numpy_episode_fields(shape, 7, 10, seed=3)               # 7 episodes: uneven shards (4 + 3)
    names = ["agent." + k for k in orc.AGENT_PARAM_NAMES] + ["mixer." + k for k in orc.QMIX_PARAM_NAMES]

    assert dp.is_active() and dp.rank() == rank and dp.world_size() == world
    lo, hi = dp.shard_slice(7, rank, world)
    mine = dp.shard_fields({k: th.from_numpy(v) for k, v in fields.items()}, rank, world)
    mine = {k: v.numpy() for k, v in mine.items()}
    assert mine["obs"].shape[0] == hi - lo == (4 if rank == 0 else 3)
    lr = orc.OracleQLearner(agent, mixer, copy.copy(args))
    fw = lr.forward_loss(mine)
    g = lr.backward(fw, mine)                                         # normalised by the LOCAL mask sum
    flat = th.from_numpy(_flat(g, names) * float(fw["mask_sum"]))      # -> un-normalised
    sums = th.tensor([float(fw["mask_sum"]), float((fw["masked_td"] ** 2).sum()), float(np.abs(fw["masked_td"]).sum()),
                      float((fw["q_tot"] * fw["mask"]).sum()), float((fw["targets"] * fw["mask"]).sum())], dtype=th.float64)
    buf = th.cat([flat, sums])                                        # ONE buffer, ONE collective per step
    dp.allreduce_step(buf)
    flat, sums = buf[:-5], buf[-5:]
    flat = flat.numpy() / float(sums[0])

    full = orc.OracleQLearner(agent, mixer, copy.copy(args))
    ffw = full.forward_loss(fields)
    fg = _flat(full.backward(ffw, fields), names)
    err = np.abs(flat - fg).max() / np.abs(fg).max()
    loss_err = abs(float(sums[1] / sums[0]) - float(ffw["loss"]))
    np.save(os.path.join(out_dir, "r%d.npy" % rank), np.array([err, loss_err]))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_step_equals_full_batch(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_run, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        err, loss_err = np.load(os.path.join(str(tmp_path), "r%d.npy" % r))
        assert err < 1e-10, err
        assert loss_err < 1e-10, loss_err


def test_shard_slices_cover_the_batch():
    from pymarl_b200.data_parallel import shard_slice
    for B in (1, 7, 8, 4096):
        for world in (1, 2, 3, 8):
            spans = [shard_slice(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
