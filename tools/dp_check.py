"""torchrun target: bench.dp_self_check on the ranks of this job (a step on a batch sharded over the ranks against the
same step on one GPU).  Prints `DP_CHECK {json}` on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29611 tools/dp_check.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

ctx = bench.Ctx()
out = bench.dp_self_check(ctx)
if ctx.rank == 0:
    print("DP_CHECK " + json.dumps(out), flush=True)
ctx.dist.destroy_process_group()
