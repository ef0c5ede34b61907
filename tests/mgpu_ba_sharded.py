"""Multi-GPU check (NOT collected by pytest; run under torchrun on >= 2 GPUs):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_ba_sharded.py
Point-sharded bundle adjustment with the NCCL all-reduce of the reduced camera system must reproduce the
single-GPU solve and the oracle (final cost within 1e-6 relative, same iteration count)."""
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as dist

import oracle
import pmv_b200
from pmv_b200 import sharding
from harness import synth


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = pmv_b200.Context(local)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid = torch.from_numpy(np.frombuffer(ctx.comm_unique_id(), np.uint8).copy()).cuda()
    dist.broadcast(uid, 0)
    ctx.comm_init(world, rank, bytes(uid.cpu().numpy().tobytes()))
    for (nposes, npts, iters) in [(12, 900, 5), (40, 4000, 4)]:
        w = synth.ba_large(7, n_poses=nposes, n_points=npts, views=5, span=min(20, nposes))
        pl, ol, cl, ptl, (lo, hi), sel = sharding.shard_points(w["points"], w["obs"], w["cam_idx"], w["pt_idx"], rank, world)
        prob = ctx.ba_problem(w["poses"], pl, ol, cl, ptl, w["K"], 1.0, rank=rank, nranks=world)
        prob.solve(iters)
        p, x, s = prob.download()
        po, xo, so = oracle.ba_solve(w["poses"], w["points"], w["obs"], w["cam_idx"], w["pt_idx"], w["K"], 1.0, iters)
        rel = abs(s[0]["final_cost"] - so["final_cost"]) / so["final_cost"]
        assert s[0]["iterations"] == so["iterations"], (s, so["iterations"])
        assert rel <= 1e-6, rel
        assert np.abs(p[0] - po).max() < 1e-5
        assert np.abs(x[0] - xo[lo:hi]).max() < 1e-4
        # every rank holds bit-identical poses (replicated camera update)
        t = torch.from_numpy(p[0]).cuda(); mx = t.clone(); mn = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        assert bool((mx == mn).all())
        if rank == 0:
            print(f"sharded BA ok: Nc={nposes} Np={npts} world={world} iters={s[0]['iterations']} "
                  f"final_cost rel diff vs oracle {rel:.2e}")
        prob.close()
    ctx.comm_destroy()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
