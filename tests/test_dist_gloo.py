"""CPU, world_size 2 over gloo: candidate sharding + record all-gather (the N>1 host logic)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from catgrasp_b200.dist import all_gather_records, pack_records, shard_range, unpack_records


def _worker(rank, world, n, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    probs = torch.rand((n, 10), generator=g)
    status = torch.randint(0, 4, (n,), generator=g).to(torch.uint8)
    offset = torch.randint(-1, 5, (n,), generator=g).to(torch.int8)
    lo, hi = shard_range(n, rank, world)
    rec = pack_records(probs[lo:hi], status[lo:hi], offset[lo:hi])   # "score" only the local shard
    full = all_gather_records(rec, n)
    p, s, o = unpack_records(full)
    ok = bool(torch.equal(p, probs) and torch.equal(s, status) and torch.equal(o, offset))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 8, 4096, 4097):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans[:-1], spans[1:]):
                assert a[1] == b[0]


def test_allgather_world2_ragged():
    ctx = mp.get_context("spawn")
    for n in (5, 64):
        q = ctx.Queue()
        port = 29500 + (os.getpid() + n) % 2000
        procs = [ctx.Process(target=_worker, args=(r, 2, n, port, q)) for r in range(2)]
        [p.start() for p in procs]
        res = [q.get(timeout=120) for _ in range(2)]
        [p.join(timeout=60) for p in procs]
        assert all(ok for _, ok in res), res


def _draw_worker(rank, world, B, M, n_pts, port, q):
    """Host half of dist.sharded_predict_batch: every rank walks numpy's stream for the WHOLE candidate list but only
    materialises its own shard (skip / draw / skip), then the shards are all-gathered."""
    import numpy as np
    from catgrasp_b200.predicter import _LegacyDraw, draw_subsample_ids_numpy
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    np.random.seed(11)
    ref = draw_subsample_ids_numpy(M, n_pts, B)
    ref_next = np.random.rand(3)
    np.random.seed(11)
    lo, hi = shard_range(B, rank, world)
    d = _LegacyDraw()
    d.skip(M, n_pts, lo)
    mine = d.draw(M, n_pts, hi - lo, nthreads=2)
    d.skip(M, n_pts, B - hi)
    d.commit()
    same_stream = bool(np.array_equal(np.random.rand(3), ref_next))
    per = (B + world - 1) // world
    pad = torch.zeros((per, n_pts), dtype=torch.int32)
    pad[: hi - lo] = torch.from_numpy(mine)
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    full = torch.cat(out)[:B].numpy()
    q.put((rank, bool(np.array_equal(full, ref)) and same_stream))
    dist.destroy_process_group()


def test_sharded_draw_world2_equals_single_process():
    ctx = mp.get_context("spawn")
    for B, M, n_pts in ((13, 3000, 256), (9, 300, 512)):
        q = ctx.Queue()
        port = 31500 + (os.getpid() + B) % 2000
        procs = [ctx.Process(target=_draw_worker, args=(r, 2, B, M, n_pts, port, q)) for r in range(2)]
        [p.start() for p in procs]
        res = [q.get(timeout=120) for _ in range(2)]
        [p.join(timeout=60) for p in procs]
        assert all(ok for _, ok in res), res
