"""CPU suite: the N>1 path's host logic (point-range sharding + the 96-byte partial all-gather +
combine) under world_size-2 gloo.  Per-rank partial sums come from the C oracle here (the CUDA
pipeline takes its place on GPUs); the sharding/gather/combine code is the product's msm_dist.py."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import bn254 as o
import helpers as h


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in ("oracle", "gpu-acceleration_b200", "tests"):
        sys.path.insert(0, os.path.join(root, p))
    import cpu_msm
    import msm_dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pts = o.random_points(n, 501)
    sc = o.random_scalars(n, 502)
    bases, scal = h.pack_bases(pts), h.pack_scalars(sc)
    b, e = msm_dist.shard_range(n, world, rank)
    if e > b:
        part, _ = cpu_msm.msm(bases[b:e], scal[b:e], 1)
    else:
        part = np.zeros(12, dtype=np.uint64)
        part[0:4] = part[4:8] = h.words(o.R_MOD_P)  # identity (R, R, 0)

    def sum_fn(gathered, cnt):
        g = gathered.numpy().view(np.uint64).reshape(cnt, 12)
        acc = g[0].copy()
        for k in range(1, cnt):
            acc = cpu_msm.jac_add(acc, g[k])
        return torch.from_numpy(acc.view(np.uint8))

    out = msm_dist.combine(torch.from_numpy(part.view(np.uint8).copy()), sum_fn)
    got = o.jac_to_affine(o.decode_jacobian(out.numpy().view(np.uint64)))
    want = o.jac_to_affine(o.msm_pippenger(pts, sc, 6))
    q.put((rank, got == want, (b, e)))
    dist.destroy_process_group()


def _run(world, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    return sorted(res)


def test_two_rank_partial_combine():
    res = _run(2, 101)
    assert all(ok for _, ok, _ in res)
    assert [r[2] for r in res] == [(0, 51), (51, 101)]


def test_shard_ranges_cover_and_tolerate_empty_ranks():
    import msm_dist
    for n, w in ((1, 2), (7, 8), (16, 8), (1 << 24, 8), (5, 3)):
        rs = [msm_dist.shard_range(n, w, r) for r in range(w)]
        assert rs[0][0] == 0 and rs[-1][1] == n
        assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
    res = _run(2, 1)  # rank 1 gets an empty shard and contributes the identity
    assert all(ok for _, ok, _ in res)
