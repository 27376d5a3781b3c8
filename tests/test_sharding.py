"""Multi-GPU host logic on CPU: 2 ranks over gloo.  Each rank runs the hot path (here: the CPU oracle,
the GPU path has the same structure) on its shard + halo of a clip; gathered on rank 0 the result must
equal the unsharded run -- i.e. sharding with a radius*delta halo is exact and needs no data-path
collective."""
import os
import socket
import sys

import numpy as np
import pytest

import sharding


def test_shard_ranges_cover_and_are_disjoint():
    for n in (1, 7, 64, 513):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                f, c = sharding.shard_range(n, world, r)
                got.extend(range(f, f + c))
                lo, hi = sharding.with_halo(f, c, n, 3)
                assert lo <= f and hi >= f + c and lo >= 0 and hi <= n
                assert f - lo == min(3, f) and hi - (f + c) == min(3, n - f - c)
            assert got == list(range(n))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_q):
    import torch.distributed as dist
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    import conftest  # noqa: F401  (sys.path)
    import clipgen
    import mvtools_b200 as mb
    import oracle
    import pipeline
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    W, H, N, radius = 64, 48, 11, 2
    frames = clipgen.make_clip(W, H, N, bits=8, scene_cut=False)

    def run(lo, hi):  # the hot path over local frames [lo, hi): returns {global n: [plane hashes]}
        loc = frames[lo:hi]
        n_loc = hi - lo
        sp = mb.super_params(W, H, 8, 1, 1, pel=2)
        aps = [mb.analyse_params(sp, n_loc, isb=isb, delta=d, blksize=8, overlap=4) for d in range(1, radius + 1) for isb in (1, 0)]
        dp = mb.degrain_params(sp, [a.ad for a in aps], W, H, radius)
        sups = [oracle.super_frame(sp, f) for f in loc]
        out = {}
        for n in range(n_loc):
            refs, vecs = [], []
            for ap in aps:
                r = pipeline.nref_of(ap, n, n_loc)
                refs.append(sups[r] if r is not None else None)
                vecs.append(oracle.analyse_frame(ap, sp, sups[n], sups[r] if r is not None else None))
            out[lo + n] = [pipeline.sha(p) for p in oracle.degrain_frame(dp, sp, loc[n], refs, vecs)]
        return out

    plan = sharding.local_plan(N, world, rank, reach=radius)
    mine = run(plan["lo"], plan["hi"])
    mine = {n: h for n, h in mine.items() if plan["first"] <= n < plan["first"] + plan["count"]}
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)  # the job's only communication: the final gather
    if rank == 0:
        full = run(0, N)
        merged = {}
        for g in gathered:
            merged.update(g)
        out_q.put((merged == full, sorted(merged) == list(range(N))))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_run_equals_unsharded(built):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    same, complete = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert complete, "the gathered shards do not cover the clip"
    assert same, "sharded (halo) result differs from the unsharded result"
