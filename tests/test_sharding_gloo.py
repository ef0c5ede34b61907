"""CPU, world_size 2, gloo: the host logic of the multi-GPU paths (no GPU kernels involved).
 * the point shards partition the observations exactly once,
 * the per-rank partial reduced camera systems (computed here with numpy from the oracle's Jacobians)
   all-reduce to the full system -- the identity the sharded solver relies on,
 * the NCCL unique-id hand-off pattern (rank 0 creates, broadcast as bytes) works over gloo,
 * window / stream partitioning covers every unit exactly once."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _partial_system(poses, points, obs, cam, pt, K, Nc):
    import oracle
    r, Jc, Jp, _ = oracle.ba_eval(poses, points, obs, cam, pt, K, 1.0)
    n = 6 * Nc
    S = np.zeros((n, n)); g = np.zeros(n)
    for p in np.unique(pt):
        idx = np.nonzero(pt == p)[0]
        V = sum(Jp[i].T @ Jp[i] for i in idx) + 1e-3 * np.eye(3)
        gp = sum(Jp[i].T @ r[i] for i in idx)
        Vi = np.linalg.inv(V)
        for i in idx:
            Wi = Jc[i].T @ Jp[i]
            g[6 * cam[i]:6 * cam[i] + 6] += Jc[i].T @ r[i] - Wi @ Vi @ gp
            S[6 * cam[i]:6 * cam[i] + 6, 6 * cam[i]:6 * cam[i] + 6] += Jc[i].T @ Jc[i]
            for k in idx:
                Wk = Jc[k].T @ Jp[k]
                S[6 * cam[i]:6 * cam[i] + 6, 6 * cam[k]:6 * cam[k] + 6] -= Wi @ Vi @ Wk.T
    return S, g


def _worker(rank, world, port, out):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import pmv_b200
    from pmv_b200 import sharding
    from harness import synth
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = synth.ba_window(5, n_poses=4, n_points=37)
        Nc = len(w["poses"])
        pl, ol, cl, ptl, (lo, hi), sel = sharding.shard_points(w["points"], w["obs"], w["cam_idx"], w["pt_idx"], rank, world)
        # 1. shards partition the observations
        cnt = torch.zeros(len(w["obs"]), dtype=torch.int64); cnt[torch.from_numpy(sel)] = 1
        dist.all_reduce(cnt)
        assert bool((cnt == 1).all())
        # 2. partial reduced systems sum to the full one
        S, g = _partial_system(w["poses"], pl, ol, cl, ptl, w["K"], Nc)
        t = torch.from_numpy(np.concatenate([S.ravel(), g]))
        dist.all_reduce(t)
        Sf, gf = _partial_system(w["poses"], w["points"], w["obs"], w["cam_idx"], w["pt_idx"], w["K"], Nc)
        assert np.allclose(t.numpy(), np.concatenate([Sf.ravel(), gf]), rtol=1e-10, atol=1e-8)
        # 3. unique-id hand-off: rank 0 makes 128 bytes, everyone receives the same
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            uid = torch.from_numpy(np.frombuffer(os.urandom(128), np.uint8).copy())
        dist.broadcast(uid, 0)
        chk = uid.to(torch.int64).sum().reshape(1).clone(); mx = chk.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        assert int(mx) == int(chk)
        # 4. window / stream partition
        lo_w, hi_w = sharding.shard_windows(4096 + 3, rank, world)
        c = torch.zeros(4096 + 3, dtype=torch.int64); c[lo_w:hi_w] = 1
        dist.all_reduce(c)
        assert bool((c == 1).all())
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_split_range_properties():
    sys.path.insert(0, str(ROOT))
    import pmv_b200
    from pmv_b200 import sharding
    for n in (0, 1, 7, 4096, 1_000_001):
        for R in (1, 2, 3, 8):
            spans = [sharding.split_range(n, r, R) for r in range(R)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
