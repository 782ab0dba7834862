"""GeometryEngine -- host driver of the CUDA path.  PyTorch only owns device memory and streams; every
computation goes through the C-ABI (include/msc_geom.h).  There is no CPU fallback anywhere in this file.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _capi
from ._capi import MscBatchIn, MscBatchOut, MscParams
from .layout import GeomParams, HostBatch, pack_batch

_IN_FIELDS = ("points", "sample_sweep_off", "sweep_start", "sweep_count", "sweep_pose", "sample_box_off", "boxes", "ego_pose",
              "lidar_calib", "cam_ego_pose", "cam_calib", "cam_K")

_NP2TORCH = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64, np.dtype(np.int32): torch.int32,
             np.dtype(np.uint32): torch.int32, np.dtype(np.uint8): torch.uint8}


def _as_torch_cpu(a: np.ndarray) -> torch.Tensor:
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint32:
        a = a.view(np.int32)
    return torch.from_numpy(a)


def make_params(p: GeomParams) -> MscParams:
    s_lo, s_hi = p.thresholds()
    return MscParams(p.remove_close_radius, p.range_min, p.range_max, p.z_min, p.z_max, p.ground_z, p.bev_range, p.bev_res, p.image_w,
                     p.image_h, p.n_cams, p.fov_keep_mask, p.centroid_shift, p.intensity_shift, s_lo, s_hi)


@dataclass
class DeviceBatch:
    """A HostBatch resident in HBM."""
    host: HostBatch
    tensors: Dict[str, torch.Tensor]

    def check_layout(self) -> None:
        """The layout rules of msc_batch_in that the library cannot check from the host (its arrays live on the device): every sweep
        starts at a multiple of 4 rows (16-byte aligned bulk copies) and the points buffer extends at least 16 bytes past the last
        row of every sweep.  A violation would be a misaligned or out-of-bounds cp.async.bulk inside the kernel; here it is an error
        before the first launch (checked once per batch, on the host copy of the sweep table)."""
        hb = self.host
        if hb.sweep_start.size == 0:
            return
        start, count = hb.sweep_start.astype(np.int64), hb.sweep_count.astype(np.int64)
        if (start % 4 != 0).any():
            raise _capi.MscError("msc_batch_in: every sweep must start at a multiple of 4 rows (sweep %d starts at row %d)"
                                 % (int(np.flatnonzero(start % 4)[0]), int(start[np.flatnonzero(start % 4)[0]])))
        rows = int(self.tensors["points"].shape[0])
        end = int((start + count).max())
        if end * 20 + 16 > rows * 20:
            raise _capi.MscError("msc_batch_in: the points buffer (%d rows) must extend 16 bytes past the last sweep row (%d)" % (rows, end))
        if str(self.tensors["points"].dtype) != "torch.float32" or self.tensors["points"].shape[1] != 5 or not self.tensors["points"].is_contiguous():
            raise _capi.MscError("msc_batch_in: points must be a contiguous (rows, 5) float32 tensor")

    def struct(self) -> MscBatchIn:
        """The C struct of this batch (built once: the tensors of a DeviceBatch never move)."""
        st = self.__dict__.get("_struct")
        if st is None or self.__dict__.get("_struct_host") is not self.host:
            t = self.tensors
            hb = self.host
            hint = int(hb.sweep_count.sum() // max(hb.n_samples, 1)) if hb.sweep_count.size else 0
            self.check_layout()
            st = MscBatchIn(hb.n_samples, hb.max_boxes_per_sample, hb.n_boxes, min(hint, 2 ** 31 - 1), *[t[k].data_ptr() for k in _IN_FIELDS])
            self.__dict__["_struct"], self.__dict__["_struct_host"] = st, hb
        return st


@dataclass
class BatchResult:
    """Result tables of msc_fused_evidence_batch, still on the device."""
    n_samples: int
    n_cams: int
    res: int
    intensity_shift: int
    sample_box_off: np.ndarray
    box_count: torch.Tensor
    box_nearest: torch.Tensor
    box_centroid: torch.Tensor
    proj_visible: torch.Tensor
    proj_extent: torch.Tensor
    bev_ci: torch.Tensor
    bev_height: torch.Tensor
    stats: torch.Tensor
    table_arena: Optional[torch.Tensor] = None   # the six small tables above are views into this one buffer (one copy / one collective)

    def struct(self) -> MscBatchOut:
        st = self.__dict__.get("_struct")
        if st is None:
            st = self.__dict__["_struct"] = self._make_struct()
        return st

    def _make_struct(self) -> MscBatchOut:
        return MscBatchOut(self.box_count.data_ptr(), self.box_nearest.data_ptr(), self.box_centroid.data_ptr(),
                           self.proj_visible.data_ptr(), self.proj_extent.data_ptr(), self.bev_ci.data_ptr(), self.bev_height.data_ptr(),
                           self.stats.data_ptr())

    def table_tensors(self):
        """Small per-box / per-sample tables (what NCCL gathers; BEV grids stay sharded)."""
        return [self.box_count, self.box_nearest, self.box_centroid, self.proj_visible, self.proj_extent, self.stats]

    def to_host(self, with_bev: bool = True) -> Dict[str, np.ndarray]:
        out = {
            "box_count": self.box_count.cpu().numpy().view(np.uint32),
            "box_nearest": self.box_nearest.cpu().numpy(),
            "box_centroid": self.box_centroid.cpu().numpy(),
            "proj_visible": self.proj_visible.cpu().numpy(),
            "proj_extent": self.proj_extent.cpu().numpy(),
            "stats": self.stats.cpu().numpy().view(np.uint32),
            "sample_box_off": self.sample_box_off,
        }
        if with_bev:
            ci = self.bev_ci.cpu().numpy().view(np.uint32)
            out["bev_count"] = np.ascontiguousarray(ci[..., 0])
            out["bev_isum_q"] = np.ascontiguousarray(ci[..., 1])
            out["bev_intensity_sum"] = (ci[..., 1].astype(np.float64) / float(1 << self.intensity_shift)).astype(np.float32)
            out["bev_height"] = self.bev_height.cpu().numpy()
        return out


class InputArena:
    """One pinned host buffer and one device buffer holding a batch's twelve input arrays at fixed offsets (see
    GeometryEngine.input_arena).  `load(hb)` copies a batch of the same shape into the pinned side (the step a real loader replaces by
    reading files straight into `host_view("points")`, msc_geom.io.stage_batch); `upload()` is one async copy."""

    def __init__(self, eng: "GeometryEngine", hb: HostBatch):
        self.eng = eng
        al = lambda v: (v + 255) & ~255
        off, self.layout = 0, {}
        for k in _IN_FIELDS:
            a = getattr(hb, k)
            self.layout[k] = (off, int(a.nbytes), a.dtype, a.shape)
            off = al(off + int(a.nbytes))
        self.nbytes = max(off, 256)
        self.host = torch.empty(self.nbytes, dtype=torch.uint8).pin_memory()
        self.dev = torch.empty(self.nbytes, dtype=torch.uint8, device=eng.device)
        self.meta = hb
        self._host_np = self.host.numpy()
        self.tensors = {}
        for k, (o, n, dt, shape) in self.layout.items():
            tdt = _NP2TORCH[np.dtype(dt)]
            self.tensors[k] = self.dev[o:o + n].view(tdt).view(shape) if n else torch.empty(shape, dtype=tdt, device=eng.device)

    def host_view(self, k: str) -> np.ndarray:
        o, n, dt, shape = self.layout[k]
        return self._host_np[o:o + n].view(dt).reshape(shape)

    def load(self, hb: HostBatch) -> None:
        for k, (o, n, dt, shape) in self.layout.items():
            a = getattr(hb, k)
            if a.shape != shape:
                raise _capi.MscError(f"InputArena: {k} has shape {a.shape}, the arena was sized for {shape}")
            if n:
                self.host_view(k)[...] = a
        self.meta = hb

    def upload(self, non_blocking: bool = True) -> DeviceBatch:
        self.dev.copy_(self.host, non_blocking=non_blocking)
        return DeviceBatch(self.meta, self.tensors)


class GeometryEngine:
    """Owns the device, the loaded library and reusable output buffers."""

    def __init__(self, device: Optional[int] = None, params: Optional[GeomParams] = None, own_context: bool = False):
        if not torch.cuda.is_available():
            raise _capi.MscError("GeometryEngine needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _capi.load()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        torch.cuda.set_device(self.device)
        # options / side stream / timing ring of the fused path; own_context=True gives this engine private ones (one per host thread)
        self.ctx = _capi.FusedContext() if own_context else _capi.default_context()
        self.params = params or GeomParams()
        sm, smem, maj, mnr = (C.c_int32() for _ in range(4))
        _capi.check(self.lib.msc_device_info(C.byref(sm), C.byref(smem), C.byref(maj), C.byref(mnr)), "msc_device_info")
        self.sm_count, self.smem_optin, self.cc = sm.value, smem.value, (maj.value, mnr.value)
        self._workspaces: Dict[int, torch.Tensor] = {}  # one work counter per stream (kernels on different streams overlap)
        self._ws_need: Dict[tuple, int] = {}
        self._params_cache: Dict[tuple, tuple] = {}
        self.kernel_launches = 0

    # ------------------------------------------------------------------ transfers
    def upload(self, hb: HostBatch, pinned: Optional[Dict[str, torch.Tensor]] = None, non_blocking: bool = False) -> DeviceBatch:
        tensors = {}
        for k in _IN_FIELDS:
            src = pinned[k] if pinned is not None else _as_torch_cpu(getattr(hb, k))
            tensors[k] = src.to(self.device, non_blocking=non_blocking)
        return DeviceBatch(hb, tensors)

    def upload_tiled(self, hb: HostBatch, reps: int, n_samples: Optional[int] = None) -> DeviceBatch:
        """Upload `hb` once and replicate it `reps` times ON THE DEVICE (distinct memory, identical content), optionally keeping only
        the first `n_samples` samples.  Host memory stays at one copy, so shards of tens of gigabytes can be built from a small pool of
        distinct samples; the returned DeviceBatch's host side carries the metadata only (its `points` array is empty)."""
        from .layout import tile_batch, truncate_batch
        npad = hb.points.shape[0] - 4
        meta_src = HostBatch(hb.n_samples, np.zeros((npad + 4, 0), np.float32), hb.sample_sweep_off, hb.sweep_start, hb.sweep_count, hb.sweep_pose,
                             hb.sweep_time_lag, hb.sample_box_off, hb.boxes, hb.ego_pose, hb.lidar_calib, hb.cam_ego_pose, hb.cam_calib, hb.cam_K,
                             hb.n_cams, hb.max_boxes_per_sample)
        meta = tile_batch(meta_src, reps)
        if n_samples is not None and n_samples < meta.n_samples:
            meta = truncate_batch(meta, n_samples)
        ns = meta.sweep_start.shape[0]
        rows = int(meta.sweep_start[ns - 1] + ((int(meta.sweep_count[ns - 1]) + 3) & ~3)) if ns else 0
        src = _as_torch_cpu(hb.points).to(self.device)
        pts = torch.empty((rows + 4, 5), dtype=torch.float32, device=self.device)
        pts[rows:] = float("nan")
        for r in range((rows + npad - 1) // npad if npad else 0):
            a = r * npad
            b = min(a + npad, rows)
            pts[a:b] = src[: b - a]
        tensors = {"points": pts}
        for k in _IN_FIELDS:
            if k != "points":
                tensors[k] = _as_torch_cpu(getattr(meta, k)).to(self.device)
        meta.points = np.zeros((0, 5), np.float32)
        return DeviceBatch(meta, tensors)

    @staticmethod
    def pin(hb: HostBatch) -> Dict[str, torch.Tensor]:
        return {k: _as_torch_cpu(getattr(hb, k)).pin_memory() for k in _IN_FIELDS}

    def input_arena(self, hb: HostBatch) -> "InputArena":
        """A reusable staging slot sized for batches shaped like `hb`: one pinned host buffer + one device buffer with the twelve input
        arrays at fixed 256-byte-aligned offsets, so a batch goes up in ONE cudaMemcpyAsync and nothing is allocated per batch."""
        return InputArena(self, hb)

    @staticmethod
    def table_layout(n_samples: int, n_boxes: int, n_cams: int):
        """Byte offsets of the six small result tables inside one arena (256-byte aligned) and its size: what one D2H copy or one
        all_gather moves per shard (fixed-stride tables, one collective)."""
        al = lambda v: (v + 255) & ~255
        off, lay = 0, {}
        for name, nbytes in (("box_count", 4 * n_boxes), ("box_nearest", 4 * n_boxes), ("box_centroid", 12 * n_boxes),
                             ("proj_visible", n_cams * n_boxes), ("proj_extent", 16 * n_cams * n_boxes),
                             ("stats", 4 * _capi.MSC_STATS_STRIDE * n_samples)):
            lay[name] = (off, nbytes)
            off = al(off + nbytes)
        return lay, max(off, 256)

    def alloc_result(self, hb: HostBatch, params: Optional[GeomParams] = None, arena_bytes: Optional[int] = None,
                     arena: Optional[torch.Tensor] = None) -> BatchResult:
        """Output buffers for `hb`.  The small tables live in ONE arena (`arena_bytes` pads it, e.g. to the largest shard of a job,
        so every rank gathers the same size; `arena` places it in caller-owned memory, e.g. this rank's row of a symmetric gathered
        buffer); the BEV layers, which stay sharded, are separate."""
        p = params or self.params
        S, B, Cn, R = hb.n_samples, hb.n_boxes, p.n_cams, p.bev_res
        d = self.device
        lay, size = self.table_layout(S, B, Cn)
        if arena is None:
            arena = torch.zeros(max(size, arena_bytes or 0), dtype=torch.uint8, device=d)
        elif arena.numel() < size or arena.dtype != torch.uint8:
            raise _capi.MscError("alloc_result: the caller's arena is smaller than the tables (%d < %d bytes)" % (arena.numel(), size))

        def view(name, dtype, shape):
            o, n = lay[name]
            return arena[o:o + n].view(dtype).view(shape)
        return BatchResult(
            S, Cn, R, p.intensity_shift, hb.sample_box_off.copy(),
            view("box_count", torch.int32, (B,)), view("box_nearest", torch.float32, (B,)), view("box_centroid", torch.float32, (B, 3)),
            view("proj_visible", torch.uint8, (B, Cn)), view("proj_extent", torch.float32, (B, Cn, 4)),
            torch.empty((S, R, R, 2), dtype=torch.int32, device=d), torch.empty((S, R, R), dtype=torch.float32, device=d),
            view("stats", torch.int32, (S, _capi.MSC_STATS_STRIDE)), arena)

    # ------------------------------------------------------------------ the hot path
    def replica_structs(self, hb: HostBatch, arenas: Sequence[torch.Tensor], params: Optional[GeomParams] = None):
        """A C array of msc_batch_out for msc_fused_evidence_batch_replicated: entry r points at the six small tables laid out (like
        alloc_result) in arenas[r] -- this shard's row of peer r's gathered buffer, peer-mapped memory.  BEV members stay null."""
        p = params or self.params
        lay, size = self.table_layout(hb.n_samples, hb.n_boxes, p.n_cams)
        arr = (MscBatchOut * max(len(arenas), 1))()
        for r, a in enumerate(arenas):
            if a.numel() < size:
                raise _capi.MscError("replica arena %d is smaller than the tables (%d < %d bytes)" % (r, a.numel(), size))
            base = a.data_ptr()
            arr[r] = MscBatchOut(base + lay["box_count"][0], base + lay["box_nearest"][0], base + lay["box_centroid"][0],
                                 base + lay["proj_visible"][0], base + lay["proj_extent"][0], 0, 0, base + lay["stats"][0])
        return arr, len(arenas)

    def run_fused(self, db: DeviceBatch, out: Optional[BatchResult] = None, params: Optional[GeomParams] = None, replicas=None) -> BatchResult:
        """One pass of the hot path over a device-resident batch.  `replicas` = replica_structs(...): the kernels also store every
        small-table entry on those peers (the gather of a sharded batch without a collective)."""
        p = params or self.params
        if p.n_cams != db.host.n_cams:
            raise _capi.MscError(f"params.n_cams={p.n_cams} but the batch was packed for {db.host.n_cams} cameras")
        if out is None:
            out = self.alloc_result(db.host, p)
        key = tuple(p.__dict__.values())  # GeomParams is mutable: the cache is keyed by value
        mp = self._params_cache.get(key)
        if mp is None:
            mp = self._params_cache[key] = make_params(p)
        bi, bo = db.struct(), out.struct()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        wkey = (db.host.n_samples, db.host.n_boxes, p.bev_res, p.bev_range, self.ctx.get_option("cull_shift"))
        ws = self._workspaces.get(stream)
        if self._ws_need.get(wkey) is None:
            self._ws_need[wkey] = int(self.lib.msc_fused_workspace_bytes(self.ctx.handle, C.byref(mp), db.host.n_samples, db.host.n_boxes))
        need = self._ws_need[wkey]
        if ws is None or ws.numel() < need:
            ws = self._workspaces[stream] = torch.zeros(need, dtype=torch.uint8, device=self.device)
        if replicas is not None and replicas[1] > 0:
            _capi.check(self.lib.msc_fused_evidence_batch_replicated(self.ctx.handle, C.byref(mp), C.byref(bi), C.byref(bo), replicas[1], replicas[0],
                                                                     ws.data_ptr(), ws.numel(), C.c_void_p(stream)), "msc_fused_evidence_batch_replicated")
        else:
            _capi.check(self.lib.msc_fused_evidence_batch(self.ctx.handle, C.byref(mp), C.byref(bi), C.byref(bo), ws.data_ptr(), ws.numel(),
                                                          C.c_void_p(stream)), "msc_fused_evidence_batch")
        self.kernel_launches += self.ctx.get_option("last_launches")  # table kernels + the streaming kernel, counted by the library
        return out

    # ------------------------------------------------------------------ batched pairwise relation tables ([EXT] e6)
    def alloc_relations(self, hb: HostBatch):
        """Ragged per-sample n_s x n_s tables in flat device arrays; returns (tensors dict, pair offsets on host)."""
        nb = np.diff(hb.sample_box_off).astype(np.int64)
        pair_off = np.zeros(hb.n_samples, np.int64)
        pair_off[1:] = np.cumsum(nb[:-1] ** 2)
        total = int((nb ** 2).sum())
        d = self.device
        return {"rect": torch.empty((hb.n_boxes, 6), dtype=torch.float64, device=d), "dist": torch.empty(total, dtype=torch.float32, device=d),
                "bearing": torch.empty(total, dtype=torch.float32, device=d), "category": torch.empty(total, dtype=torch.uint8, device=d),
                "overlap": torch.empty(total, dtype=torch.uint8, device=d), "pair_off": torch.from_numpy(pair_off).to(d)}, pair_off

    def run_relations(self, db: DeviceBatch, rel: Dict[str, torch.Tensor]) -> None:
        """Footprints of every box (loader's global frame, like the reference's annotation path) + one batched relation launch."""
        hb = db.host
        if hb.n_boxes == 0:
            return
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _capi.check(self.lib.msc_box_footprints(hb.n_boxes, db.tensors["boxes"].data_ptr(), None, rel["rect"].data_ptr(), stream), "msc_box_footprints")
        _capi.check(self.lib.msc_relation_table_batch(hb.n_samples, hb.max_boxes_per_sample, db.tensors["sample_box_off"].data_ptr(),
                                                      rel["pair_off"].data_ptr(), rel["rect"].data_ptr(), rel["dist"].data_ptr(), rel["bearing"].data_ptr(),
                                                      rel["category"].data_ptr(), rel["overlap"].data_ptr(), stream), "msc_relation_table_batch")
        self.kernel_launches += 2

    def process_samples(self, samples: Sequence[dict], params: Optional[GeomParams] = None) -> Dict[str, np.ndarray]:
        """Host-in / host-out convenience: pack, upload, run, download."""
        p = params or self.params
        hb = pack_batch(list(samples), n_cams=p.n_cams)
        res = self.run_fused(self.upload(hb), params=p)
        torch.cuda.synchronize(self.device)
        return res.to_host()
