"""Host-side logic: batch layout, pose algebra, exact thresholds, sharding."""
import numpy as np
import pytest

from msc_geom import geometry as G
from msc_geom.dist import shard_range
from msc_geom.layout import GeomParams, pack_batch, tile_batch
from msc_geom.synthetic import make_sample


def test_sqrt_thresholds_are_exact():
    s_lo, s_hi = G.sqrt_thresholds(1.0, 50.0)
    for thr, base in ((np.float32(1.0), np.float32(s_lo)), (np.float32(50.0), np.float32(s_hi))):
        s = base
        for _ in range(2000):
            s = np.nextafter(s, np.float32(-np.inf))
        for _ in range(4000):
            d = np.sqrt(s)
            if thr == 1.0:
                assert (d > thr) == (s >= np.float32(s_lo))
            else:
                assert (d < thr) == (s <= np.float32(s_hi))
            s = np.nextafter(s, np.float32(np.inf))
    rng = np.random.default_rng(0)
    s = rng.uniform(0, 3000, 200000).astype(np.float32)
    d = np.sqrt(s)
    assert np.array_equal((d > 1.0) & (d < 50.0), (s >= np.float32(s_lo)) & (s <= np.float32(s_hi)))


def test_ref_from_sweep_identity_and_inverse():
    s = make_sample(3, n_sweeps=4)
    M0 = s["lidar_sweeps"][0]["ref_from_sensor"]
    assert np.allclose(M0, np.eye(4)[:3], atol=1e-12)
    sw = s["lidar_sweeps"][2]
    M = np.vstack([sw["ref_from_sensor"], [0, 0, 0, 1]])
    back = np.vstack([G.ref_from_sweep(sw["ego_pose"], sw["calib"], s["ego_pose"], s["lidar_calib"]), [0, 0, 0, 1]])
    assert np.allclose(M @ back, np.eye(4), atol=1e-9)
    R = G.quat_to_rot(G.rot_to_quat(M[:3, :3]))
    assert np.allclose(R, M[:3, :3], atol=1e-12)


def test_pack_batch_layout():
    samples = [make_sample(i, n_sweeps=3 if i else 2, n_boxes=5 + i) for i in range(3)]
    samples[1]["lidar_sweeps"][1]["points_raw"] = samples[1]["lidar_sweeps"][1]["points_raw"][:1001]  # ragged sweep
    hb = pack_batch(samples)
    assert hb.n_samples == 3 and hb.sample_sweep_off.tolist() == [0, 2, 5, 8]
    assert (hb.sweep_start % 4 == 0).all(), "16-byte aligned sweep starts"
    assert hb.points.shape[0] >= int(hb.sweep_start[-1] + hb.sweep_count[-1]) + 4
    k = 0
    for s in samples:
        for sw in s["lidar_sweeps"]:
            a, n = int(hb.sweep_start[k]), int(hb.sweep_count[k])
            assert np.array_equal(hb.points[a:a + n], sw["points_raw"])
            assert np.array_equal(hb.sweep_pose[k].reshape(3, 4), sw["ref_from_sensor"])
            k += 1
    pad = hb.points[int(hb.sweep_start[3] + hb.sweep_count[3]): int(hb.sweep_start[4])]
    assert pad.shape[0] == 3 and np.isnan(pad).all(), "padding rows are NaN so they fail every compare"
    assert hb.sample_box_off.tolist() == [0, 5, 11, 18] and hb.max_boxes_per_sample == 7
    assert np.array_equal(hb.boxes[5, :3], samples[1]["annotations"][0]["translation"])
    t = tile_batch(hb, 3)
    assert t.n_samples == 9 and t.n_points == 3 * hb.n_points and t.boxes.shape[0] == 3 * hb.boxes.shape[0]
    a, n = int(t.sweep_start[8 + 2]), int(t.sweep_count[8 + 2])
    assert np.array_equal(t.points[a:a + n], samples[1]["lidar_sweeps"][0]["points_raw"])


def test_plain_loader_sample_becomes_identity_sweep():
    pc = np.random.default_rng(1).normal(size=(100, 4)).astype(np.float32)
    hb = pack_batch([{"point_cloud": pc, "annotations": []}])
    assert hb.sweep_count.tolist() == [100] and np.array_equal(hb.points[:100, :4], pc)
    assert np.array_equal(hb.sweep_pose[0].reshape(3, 4), np.eye(4)[:3])


def test_params_defaults_follow_reference():
    p = GeomParams()
    assert (p.range_min, p.range_max, p.z_min, p.z_max, p.ground_z, p.bev_range) == (1.0, 50.0, -3.0, 5.0, -1.4, 50.0)
    assert p.centroid_shift == 17 and GeomParams(range_max=20.0).centroid_shift == 17 and p.intensity_shift == 8


@pytest.mark.parametrize("lo,hi", [(1.0, 50.0), (0.0, 30.0), (-1.0, float("inf")), (2.5, 64.0), (1e-3, 1e3), (5.0, 5.0), (7.0, 3.0)])
def test_sqrt_thresholds_general(lo, hi):
    s_lo, s_hi = G.sqrt_thresholds(lo, hi)
    s = np.concatenate([np.float32([0.0, 1e-30, lo * lo, hi * hi if np.isfinite(hi) else 1e30, 3e38, np.inf, np.nan]),
                        np.random.default_rng(2).uniform(0, 5000, 50000).astype(np.float32)])
    for v in (lo * lo, hi * hi):
        if np.isfinite(v) and v > 0:
            b = np.float32(v)
            ring = [b]
            for _ in range(64):
                ring += [np.nextafter(ring[-1], np.float32(np.inf))]
            ring2 = [b]
            for _ in range(64):
                ring2 += [np.nextafter(ring2[-1], np.float32(0))]
            s = np.concatenate([s, np.float32(ring + ring2)])
    with np.errstate(invalid="ignore"):
        d = np.sqrt(s)
        want = (d > np.float32(lo)) & (d < np.float32(hi))
        got = (s >= np.float32(s_lo)) & (s <= np.float32(s_hi))
    assert np.array_equal(want, got)


@pytest.mark.parametrize("n,world", [(404, 8), (34149, 8), (5, 8), (0, 2), (592, 1)])
def test_shard_range_partitions(n, world):
    spans = [shard_range(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert max(hi - lo for lo, hi in spans) == (n + world - 1) // world if n else True


def test_stage_batch_from_files_equals_pack_batch(tmp_path):
    """.pcd.bin sweeps read in place into the batch layout give exactly what pack_batch builds in memory."""
    from msc_geom import io as mio
    samples = [make_sample(i, n_sweeps=3, n_boxes=4 + i) for i in range(3)]
    samples[2]["lidar_sweeps"][1]["points_raw"] = samples[2]["lidar_sweeps"][1]["points_raw"][:1001]
    file_samples = []
    for i, s in enumerate(samples):
        fs = dict(s)
        fs["lidar_sweeps"] = []
        for k, sw in enumerate(s["lidar_sweeps"]):
            path = str(tmp_path / f"s{i}_{k}.pcd.bin")
            mio.write_pcd_bin(path, sw["points_raw"])
            assert mio.pcd_bin_points(path) == sw["points_raw"].shape[0]
            fs["lidar_sweeps"].append({"path": path, "ref_from_sensor": sw["ref_from_sensor"], "time_lag": sw["time_lag"]})
        file_samples.append(fs)
    file_samples[1]["lidar_sweeps"][0] = samples[1]["lidar_sweeps"][0]          # in-memory sweeps can be mixed in
    a, b = pack_batch(samples), mio.stage_batch(file_samples, threads=4, pinned=False)
    for k in ("sample_sweep_off", "sweep_start", "sweep_count", "sweep_pose", "sweep_time_lag", "sample_box_off", "boxes", "ego_pose", "lidar_calib",
              "cam_ego_pose", "cam_calib", "cam_K"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    assert np.array_equal(a.points, b.points, equal_nan=True) and b.max_boxes_per_sample == a.max_boxes_per_sample
    with open(tmp_path / "bad.pcd.bin", "wb") as f:
        f.write(b"\0" * 30)
    with pytest.raises(ValueError):
        mio.pcd_bin_points(str(tmp_path / "bad.pcd.bin"))


def test_token_only_scene_scan():
    from msc_geom.nuscenes_loader import SyntheticNuScenesLoader, create_loader
    ld = SyntheticNuScenesLoader(n_scenes=3, samples_per_scene=4, n_sweeps=1)
    toks = [t for sc in ld.get_scene_list() for t in ld.scene_sample_tokens(sc["token"])]
    assert len(toks) == 12 and toks[5] == "synth_sample_000005"
    assert ld.load_sample(toks[5])["sample_token"] == toks[5]
    assert len(create_loader(None, use_mock=True).scene_sample_tokens("mock_scene_001")) == 5


def test_device_batch_layout_rules_are_checked_on_the_host():
    """msc_batch_in's alignment / tail-padding rules (include/msc_geom.h) are enforced before the first launch."""
    import pytest
    import torch
    from msc_geom import _capi
    from msc_geom.engine import DeviceBatch
    from msc_geom.layout import pack_batch
    from msc_geom.synthetic import make_sample
    import dataclasses
    hb = pack_batch([make_sample(3, n_sweeps=2, n_boxes=2)])
    pts = torch.from_numpy(hb.points)
    DeviceBatch(hb, {"points": pts}).check_layout()
    bad = dataclasses.replace(hb, sweep_start=hb.sweep_start + np.uint32(2))
    with pytest.raises(_capi.MscError, match="multiple of 4"):
        DeviceBatch(bad, {"points": pts}).check_layout()
    end = int((hb.sweep_start.astype(np.int64) + hb.sweep_count).max())
    with pytest.raises(_capi.MscError, match="16 bytes past"):
        DeviceBatch(hb, {"points": pts[:end]}).check_layout()


def test_jpeg_header_parser_and_entropy_decoder_on_the_host():
    """The host half of the camera-image decode (no GPU involved): headers of every flavour PIL writes, coefficient counts, DC terms of a
    flat image, refusal of progressive files and garbage."""
    import ctypes as C
    import io
    import pytest
    from PIL import Image
    from msc_geom import _capi
    lib = _capi.load()

    def enc(img, **kw):
        b = io.BytesIO(); Image.fromarray(img).save(b, "JPEG", **kw); return np.frombuffer(b.getvalue(), dtype=np.uint8)
    flat = np.full((20, 35, 3), 200, np.uint8)
    for sub, (hmax, vmax) in ((0, (1, 1)), (1, (2, 1)), (2, (2, 2))):
        data = enc(flat, quality=90, subsampling=sub)
        d = _capi.MscJpegDesc()
        assert lib.msc_jpeg_info(data.ctypes.data, data.size, C.byref(d)) == 0
        assert (d.width, d.height, d.n_comp, d.hmax, d.vmax) == (35, 20, 3, hmax, vmax)
        assert d.mcus_x == -(-35 // (8 * hmax)) and d.mcus_y == -(-20 // (8 * vmax))
        assert d.coef_elems == sum(d.comp[c].blocks_x * d.comp[c].blocks_y * 64 for c in range(3))
        coef = np.zeros(d.coef_elems, np.int16)
        assert lib.msc_jpeg_entropy_decode_host(data.ctypes.data, data.size, C.byref(d), coef.ctypes.data) == 0
        y = coef[:d.comp[0].blocks_x * d.comp[0].blocks_y * 64].reshape(-1, 64)
        assert (y[:, 1:] == 0).all() and (y[0, 0] * d.comp[0].qt[0] > 0)      # a flat image has DC terms only (+ level shift)
    prog = enc(flat, progressive=True)
    d = _capi.MscJpegDesc()
    assert lib.msc_jpeg_info(prog.ctypes.data, prog.size, C.byref(d)) == _capi_status("MSC_ERR_UNSUPPORTED")
    junk = np.frombuffer(b"\x00\x01garbage", dtype=np.uint8)
    assert lib.msc_jpeg_info(junk.ctypes.data, junk.size, C.byref(d)) < 0


def _capi_status(name):
    return {"MSC_ERR_BAD_ARGUMENT": -1, "MSC_ERR_LAUNCH": -2, "MSC_ERR_UNSUPPORTED": -3, "MSC_ERR_NO_DEVICE": -4}[name]


def test_loader_only_detours_to_pil_for_unsupported_jpeg_flavours(tmp_path, monkeypatch):
    """NuScenesLoader._load_cameras with an engine: a file flavour the decoder reports as MSC_ERR_UNSUPPORTED goes to the reference's
    decoder (PIL); any other failure of the CUDA path -- missing library, corrupt stream, failed launch -- is raised, never hidden."""
    from PIL import Image
    from msc_geom import _capi, ops
    from msc_geom.nuscenes_loader import NuScenesLoader
    img = np.random.default_rng(1).integers(0, 255, (24, 40, 3), dtype=np.uint8)
    path = tmp_path / "cam.jpg"
    Image.fromarray(img).save(path, "JPEG", quality=95)
    loader = NuScenesLoader.__new__(NuScenesLoader)  # (no devkit tables needed for this method)
    loader.engine = object()

    def refuse(eng, blobs, **kw):
        raise _capi.MscError("msc_jpeg_info failed with status -3: progressive", _capi.MSC_ERR_UNSUPPORTED)
    monkeypatch.setattr(ops, "decode_jpeg_batch", refuse)
    got = loader._load_cameras([path])
    assert np.array_equal(got[0], np.array(Image.open(path)))
    for status in (-1, -2, -4, 0):
        def fail(eng, blobs, status=status, **kw):
            raise _capi.MscError("failed", status)
        monkeypatch.setattr(ops, "decode_jpeg_batch", fail)
        with pytest.raises(_capi.MscError):
            loader._load_cameras([path])


def test_static_partition_arithmetic_matches_brute_force():
    """The closed forms stream4.cu uses for its static partition of the batch's warp tiles over G CTAs (CTA c owns global tiles
    [c * total // G, (c + 1) * total // G)): the owner of a tile, and the number of CTAs that own at least one tile of a sample --
    the count the last-ticket finalisation waits for.  Restated here and checked against enumeration, including totals below G, where
    most CTAs own nothing (the case a contiguous-owners formula gets wrong)."""
    rng = np.random.default_rng(11)

    def owner(r, total, G):  # stream4.cu: cta_of
        return ((r + 1) * G + total - 1) // total - 1

    def n_parts(off, n, total, G):  # stream4.cu: parts of the sample whose tiles are [off, off + n)
        return owner(off + n - 1, total, G) - owner(off, total, G) + 1 if total >= G else n

    for _ in range(300):
        G = int(rng.integers(1, 200))
        sizes = rng.integers(0, 40, int(rng.integers(1, 12)))
        if rng.random() < 0.3:
            sizes = rng.integers(0, 3000, len(sizes))
        total = int(sizes.sum())
        if total == 0:
            continue
        lo = [c * total // G for c in range(G)]
        hi = [(c + 1) * total // G for c in range(G)]
        own = np.empty(total, np.int64)
        for c in range(G):
            own[lo[c]:hi[c]] = c
        assert all(owner(r, total, G) == own[r] for r in rng.integers(0, total, 50))
        off = 0
        for n in map(int, sizes):
            if n:
                assert n_parts(off, n, total, G) == len(set(own[off:off + n].tolist())), (G, total, off, n)
            off += n
