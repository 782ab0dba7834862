"""Multi-GPU plumbing: samples are independent (the reference's only many-sample loop touches one sample per
iteration, src/evaluation_framework.py:534-556), so a batch is split into contiguous blocks, one per rank, each
rank runs the fused path on its shard with no data-path collective, and ONE all_gather per step moves the shard's
table arena (the six small result tables in one buffer, padded to the largest shard) on a side stream.  BEV grids
stay sharded.  Works with NCCL (device tensors) and gloo (host tensors, used by the CPU tests)."""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_samples: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of the sample index range for `rank`: ceil(n/world) per rank, last ranks may be short."""
    per = (n_samples + world - 1) // world
    lo = min(rank * per, n_samples)
    return lo, min(lo + per, n_samples)


def table_shapes(n_samples: int, n_boxes: int, n_cams: int) -> dict:
    """dtype / shape of the six small result tables of a shard (the views BatchResult holds into its table arena)."""
    return {"box_count": (torch.int32, (n_boxes,)), "box_nearest": (torch.float32, (n_boxes,)), "box_centroid": (torch.float32, (n_boxes, 3)),
            "proj_visible": (torch.uint8, (n_boxes, n_cams)), "proj_extent": (torch.float32, (n_boxes, n_cams, 4)),
            "stats": (torch.int32, (n_samples, 16))}


def rank_layout(n_samples: int, n_boxes: int, n_cams: int) -> dict:
    """Arena layout of one rank's shard (offsets from GeometryEngine.table_layout + shapes), the argument of split_arena()."""
    from .engine import GeometryEngine
    lay, size = GeometryEngine.table_layout(n_samples, n_boxes, n_cams)
    lay = dict(lay)
    lay["shapes"] = table_shapes(n_samples, n_boxes, n_cams)
    lay["bytes"] = size
    return lay


def pack_tables_host(tables: Dict[str, np.ndarray], layout: dict, arena_bytes: int) -> torch.Tensor:
    """Host-side twin of the device arena (CPU tests of the gather path): the ragged per-box arrays of one shard at the arena offsets."""
    arena = torch.zeros(arena_bytes, dtype=torch.uint8)
    for name, (dtype, shape) in layout["shapes"].items():
        o, n = layout[name]
        if n:
            src = np.ascontiguousarray(tables[name])
            arena[o:o + n] = torch.from_numpy(src.view(np.uint8).reshape(-1))
    return arena


class TableGather:
    """ONE collective per step, off the critical path: the shard's table arena (GeometryEngine.alloc_result: the six small tables in
    one buffer, padded to the largest shard) is all-gathered on a side stream into one of two output buffers, so the gather of step k
    overlaps the streaming kernel of step k + 1.  `launch(arena)` orders the gather after everything queued on the current stream;
    `wait()` makes the current stream wait for every gather in flight.  NCCL for device tensors, gloo for host tensors (CPU tests)."""

    def __init__(self, arena_bytes: int, device=None, depth: int = 2):
        self.world = dist.get_world_size()
        self.device = device
        self.cuda = device is not None and torch.device(device).type == "cuda"
        self.arena_bytes = arena_bytes
        self.bufs = [torch.empty(self.world * arena_bytes, dtype=torch.uint8, device=device) for _ in range(depth)]
        self.k = 0
        if self.cuda:
            self.side = torch.cuda.Stream(device=device)
            self.done = [None] * depth

    def launch(self, arena: torch.Tensor) -> torch.Tensor:
        slot = self.k % len(self.bufs)
        self.k += 1
        out = self.bufs[slot]
        if not self.cuda:
            dist.all_gather_into_tensor(out, arena.contiguous())
            return out.view(self.world, self.arena_bytes)
        cur = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(cur)
        self.side.wait_event(ready)          # the tables of this step are complete
        with torch.cuda.stream(self.side):
            dist.all_gather_into_tensor(out, arena)
            ev = torch.cuda.Event()
            ev.record(self.side)
        self.done[slot] = ev
        return out.view(self.world, self.arena_bytes)

    def wait(self) -> None:
        if self.cuda:
            cur = torch.cuda.current_stream(self.device)
            for ev in self.done:
                if ev is not None:
                    cur.wait_event(ev)


class PeerTableGather:
    """The same gather without a collective KERNEL: every rank PUSHES its arena into its row of every peer's gathered buffer with
    copy-engine peer copies (symmetric memory over NVLink / NVSwitch; `Tensor.copy_` between peer-mapped buffers is a cudaMemcpyAsync
    peer copy) on a side stream.  No SM is involved, so the copies of step k run under the streaming kernel of step k + 1 -- that
    kernel holds every SM's shared memory, and an NCCL all-gather kernel queued beside it only starts in its tail (measured: 0.16 ms of
    a 2.13 ms step at 8 GPUs).  `wait()` = own pushes done + a cross-rank barrier on the symmetric-memory signal pads: after it every
    row of the most recent buffers is complete on every rank."""

    def __init__(self, arena_bytes: int, device, depth: int = 2):
        import torch.distributed._symmetric_memory as symm_mem
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.device, self.arena_bytes, self.depth = device, arena_bytes, depth
        self.local = symm_mem.empty(depth * self.world * arena_bytes, dtype=torch.uint8, device=device)
        self.hdl = symm_mem.rendezvous(self.local, dist.group.WORLD)
        self.peer = [self.hdl.get_buffer(p, (depth, self.world, arena_bytes), torch.uint8) for p in range(self.world)]
        self.side = torch.cuda.Stream(device=device)
        self.done = [None] * depth
        self.k = 0

    def launch(self, arena: torch.Tensor) -> torch.Tensor:
        slot = self.k % self.depth
        self.k += 1
        cur = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(cur)
        self.side.wait_event(ready)          # the tables of this step are complete
        with torch.cuda.stream(self.side):
            for j in range(self.world):      # staggered targets: at any moment every rank writes to a different peer
                p = (self.rank + j) % self.world
                self.peer[p][slot, self.rank].copy_(arena, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.side)
        self.done[slot] = ev
        return self.peer[self.rank][slot]

    def wait(self) -> None:
        cur = torch.cuda.current_stream(self.device)
        for ev in self.done:
            if ev is not None:
                cur.wait_event(ev)
        self.hdl.barrier()                   # every peer's pushes into this rank's buffers are done as well


class FusedTableGather:
    """No gather step at all: the result tables of a shard are WRITTEN into every GPU's gathered buffer by the kernels that produce
    them (msc_fused_evidence_batch_replicated: P2P stores through peer-mapped symmetric memory over NVLink / NVSwitch).  This object
    only owns the buffers: `result(eng, hb, slot)` is a BatchResult whose table arena is this rank's row of its OWN gathered buffer,
    `replicas(eng, hb, slot)` the C structs of the same row on every peer; `wait()` is the cross-GPU barrier after which every row of
    every rank's buffers is complete."""

    def __init__(self, arena_bytes: int, device, depth: int = 2):
        import torch.distributed._symmetric_memory as symm_mem
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.device, self.arena_bytes, self.depth = device, arena_bytes, depth
        self.local = symm_mem.empty(depth * self.world * arena_bytes, dtype=torch.uint8, device=device)
        self.local.zero_()
        self.hdl = symm_mem.rendezvous(self.local, dist.group.WORLD)
        self.peer = [self.hdl.get_buffer(p, (depth, self.world, arena_bytes), torch.uint8) for p in range(self.world)]
        self.done = [None] * depth           # (interface of the other gathers: nothing of this one is ever in flight on a side stream)

    def result(self, eng, hb, slot: int, params=None):
        return eng.alloc_result(hb, params, arena=self.peer[self.rank][slot, self.rank])

    def replicas(self, eng, hb, slot: int, params=None):
        return eng.replica_structs(hb, [self.peer[p][slot, self.rank] for p in range(self.world) if p != self.rank], params)

    def gathered(self, slot: int) -> torch.Tensor:
        return self.peer[self.rank][slot]

    def launch(self, arena: torch.Tensor) -> None:  # the kernels already did it
        return None

    def wait(self) -> None:
        torch.cuda.current_stream(self.device).synchronize()
        self.hdl.barrier()


def make_table_gather(arena_bytes: int, device, prefer_peer: bool = True):
    """(gather object, description): the copy-engine gather where symmetric memory is available, else the NCCL / gloo collective."""
    if prefer_peer and device is not None and torch.device(device).type == "cuda":
        try:
            return PeerTableGather(arena_bytes, device), "copy-engine pushes into peer (symmetric) memory, no collective kernel"
        except Exception as e:  # noqa: BLE001 -- no symmetric memory on this platform / build
            why = "%s: %s" % (type(e).__name__, str(e).splitlines()[0][:120] if str(e) else "")
            return TableGather(arena_bytes, device=device), "one all_gather_into_tensor on a side stream (symmetric memory unavailable: %s)" % why
    return TableGather(arena_bytes, device=device), "one all_gather_into_tensor on a side stream"


def split_arena(buf: torch.Tensor, rank_layouts: Sequence[dict]) -> List[Dict[str, torch.Tensor]]:
    """Views of a gathered [world, arena_bytes] buffer as per-rank tables; rank_layouts[r] = GeometryEngine.table_layout(...)[0] plus
    a "shapes" entry {name: (dtype, shape)} for rank r's shard."""
    out = []
    for r, lay in enumerate(rank_layouts):
        row, tabs = buf[r], {}
        for name, (dtype, shape) in lay["shapes"].items():
            o, n = lay[name]
            tabs[name] = row[o:o + n].view(dtype).view(shape)
        out.append(tabs)
    return out


def gpu_numa_node(index: int) -> int:
    """NUMA node of CUDA device `index` from sysfs (-1 when the platform does not say)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        path = "/sys/bus/pci/devices/%s:%s/numa_node" % (dom[-4:].lower(), rest.lower())
        return int(open(path).read().strip())
    except Exception:
        return -1


def bind_to_gpu_numa(index: int) -> dict:
    """Pin this process (and therefore the pinned buffers it first-touches and its copy-issuing threads) to the cores of the NUMA node the
    GPU hangs off, so eight ranks do not stage through one socket.  Returns what was done, for the bench record."""
    import os
    node = gpu_numa_node(index)
    info = {"numa_node": node, "bound": False}
    if node < 0:
        return info
    try:
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            info.update(bound=True, cpus=len(allowed))
    except Exception as e:  # noqa: BLE001
        info["error"] = str(e)
    return info
