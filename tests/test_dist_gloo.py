"""N>1 path on CPU: world_size-2 gloo run of the sharding + the one-collective table-arena gather (msc_geom/dist.py)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from msc_geom.dist import TableGather, pack_tables_host, rank_layout, shard_range, split_arena


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _fake_host_tables(lo, hi, n_cams):
    rng = np.random.default_rng(100)
    nb_all = rng.integers(0, 5, 64)
    nb = nb_all[lo:hi]
    off = np.zeros(len(nb) + 1, np.int32); off[1:] = np.cumsum(nb)
    B = int(off[-1])
    ids = np.concatenate([np.full(n, lo + i) for i, n in enumerate(nb)]) if B else np.zeros(0)
    return {"sample_box_off": off, "box_count": (ids * 7 % 13).astype(np.uint32), "box_nearest": ids.astype(np.float32) + 0.5,
            "box_centroid": np.repeat(ids[:, None], 3, 1).astype(np.float32), "proj_visible": np.ones((B, n_cams), np.uint8),
            "proj_extent": np.zeros((B, n_cams, 4), np.float32),
            "stats": np.tile(np.arange(lo, hi, dtype=np.uint32)[:, None], (1, 16))}, nb_all


def _worker(rank, world, port, n_samples, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # every rank knows every shard's shape from the (host-side) box offsets, like bench.py: ragged shards, one arena size for all
    _, nb_all = _fake_host_tables(0, n_samples, 6)
    shards = [shard_range(n_samples, r, world) for r in range(world)]
    layouts = [rank_layout(hi - lo, int(nb_all[lo:hi].sum()), 6) for lo, hi in shards]
    arena_bytes = max(l["bytes"] for l in layouts)
    lo, hi = shards[rank]
    host, _ = _fake_host_tables(lo, hi, 6)
    tg = TableGather(arena_bytes)                      # gloo: host tensors, same code path as the NCCL one minus the side stream
    ok = True
    for step in range(3):                              # double-buffered outputs: consecutive steps land in different buffers
        host["stats"][:, 1] = step
        buf = tg.launch(pack_tables_host(host, layouts[rank], arena_bytes))
        tg.wait()
        tabs = split_arena(buf, layouts)
        ok &= buf.shape == (world, arena_bytes)
        for r, (a, b) in enumerate(shards):
            ref, _ = _fake_host_tables(a, b, 6)
            ok &= bool((tabs[r]["stats"][:, 0] == torch.arange(a, b, dtype=torch.int32)).all()) and bool((tabs[r]["stats"][:, 1] == step).all())
            ok &= np.array_equal(tabs[r]["box_nearest"].numpy(), ref["box_nearest"]) and np.array_equal(tabs[r]["box_centroid"].numpy(), ref["box_centroid"])
            ok &= np.array_equal(tabs[r]["box_count"].numpy().view(np.uint32), ref["box_count"]) and tabs[r]["proj_visible"].shape == ref["proj_visible"].shape
    t = torch.tensor([1.0 + rank]); dist.all_reduce(t, op=dist.ReduceOp.MAX)  # bench.py's max-over-ranks timing reduction
    ok &= float(t) == float(world)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_two_rank_shard_and_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 37, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in procs)
    [p.join(60) for p in procs]
    assert res == [(0, True), (1, True)]
