"""Multi-GPU check of the fused table gather (run under torchrun with >= 2 GPUs): every rank runs its own shard through
msc_fused_evidence_batch_replicated; afterwards row r of EVERY rank's gathered buffer must equal rank r's local tables, which must equal
the tables of a plain (non-replicated) call.  usage: torchrun --nproc-per-node N tools/check_fused_gather.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-scene-captioning_b200")); sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from msc_geom.dist import FusedTableGather, PeerTableGather
from msc_geom.engine import GeometryEngine
from msc_geom.layout import pack_batch
from msc_geom.synthetic import make_sample

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = GeometryEngine(local)
hb = pack_batch([make_sample(1000 * rank + i, n_sweeps=2 + (i % 2), n_boxes=10 + 7 * rank + i) for i in range(3 + rank)])  # ragged shards
db = eng.upload(hb)
_, size = eng.table_layout(hb.n_samples, hb.n_boxes, eng.params.n_cams)
t = torch.tensor([size], device="cuda", dtype=torch.int64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
arena_bytes = int(t.item())
plain = eng.run_fused(db, eng.alloc_result(hb, arena_bytes=arena_bytes))
torch.cuda.synchronize()
gat = FusedTableGather(arena_bytes, eng.device)
ok = True
for slot in (0, 1, 0):
    out = gat.result(eng, hb, slot)
    out.table_arena.zero_()
    eng.run_fused(db, out, replicas=gat.replicas(eng, hb, slot))
    gat.wait()
    mine = out.table_arena.clone()
    ok &= bool(torch.equal(mine[:size], plain.table_arena[:size]))            # same tables as the plain call
    rows = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(rows, mine)                                             # reference gather (NCCL) of the local tables
    g = gat.gathered(slot)
    for r in range(world):
        ok &= bool(torch.equal(g[r], rows[r]))                               # every row of MY gathered buffer is rank r's tables
    dist.barrier()
# the copy-engine variant must produce the same gathered buffer
pg = PeerTableGather(arena_bytes, eng.device)
g2 = pg.launch(plain.table_arena)
pg.wait()
rows = [torch.empty_like(plain.table_arena) for _ in range(world)]
dist.all_gather(rows, plain.table_arena)
for r in range(world):
    ok &= bool(torch.equal(g2[r], rows[r]))
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("fused gather check:", "OK" if int(flag.item()) else "MISMATCH", "world", world, "arena bytes", arena_bytes)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
