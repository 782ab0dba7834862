"""N>1 path on CPU: world_size-2 gloo run of the sharding + table gather (msc_geom/dist.py)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from msc_geom.dist import gather_tables, pad_tables, shard_range


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
    lo, hi = shard_range(n_samples, rank, world)
    per = (n_samples + world - 1) // world
    host, nb_all = _fake_host_tables(lo, hi, 6)
    g = gather_tables(pad_tables(host, hi - lo, 6, 8, per))
    ok = g["stats"].shape == (world * per, 16)
    ok &= bool((g["stats"][:n_samples, 0] == torch.arange(n_samples, dtype=torch.int32)).all())
    ok &= bool((g["n_boxes"][:n_samples] == torch.from_numpy(nb_all[:n_samples].astype(np.int32))).all())
    for i in range(n_samples):
        n = int(g["n_boxes"][i])
        ok &= bool((g["box_nearest"][i, :n] == i + 0.5).all()) and bool(torch.isinf(g["box_nearest"][i, n:]).all())
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
