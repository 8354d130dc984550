"""Size-independent properties of the CUDA path at BASELINE.json's full sizes, where the oracle is too slow
to run whole frames: determinism, degenerate-parameter identities, band-split == whole frame (the multi-GPU
claim, checked with two band contexts on one GPU), and an oracle spot check of a row band at 1080p."""
import json
import os
import subprocess

import numpy as np
import pytest

import parity_harness as ph

pytestmark = pytest.mark.gpu
capi, fixtures = ph.capi, ph.fixtures
bands = __import__("restir_vulkan_b200.bands", fromlist=["bands"])


def _torch():
    import torch

    assert torch.cuda.is_available()
    return torch


def _scene_full():
    if fixtures.baked_available("sponza"):
        return fixtures.load_baked("sponza"), ((3.0, 4.0, 5.0), (0.0, 0.0, 0.0))
    return fixtures.make_procedural(seed=7, grid=60, boxes=600, lights="random"), ((3.0, 3.5, 4.2), (0.0, -1.0, 0.0))


class DeviceFrames:
    """Full-size frame sequences with G-buffers rendered on the device by the fixture tool."""

    def __init__(self, scene, cam_pos, look, w, h, band=None, halo=31):
        torch = _torch()
        self.torch, self.scene, self.w, self.h = torch, scene, w, h
        self.ctx = ph.make_context(scene)
        if band is None:
            self.ctx.resize(w, h)
        else:
            self.ctx.resize_band(w, h, band[0], band[1], halo)
        self.rb, self.re, self.a0, self.a1 = self.ctx.band()
        rows = self.a1 - self.a0
        self.cams = [capi.make_camera(position=(cam_pos[0] + 0.05 * k, cam_pos[1], cam_pos[2]), look_at=look, aspect=w / h) for k in range(2)]
        tm = torch.from_numpy(np.ascontiguousarray(scene.tri_material)).cuda()
        mt = torch.from_numpy(scene.material_table().view(np.int32)).cuda()
        self.gb = []
        for c in self.cams:
            planes = [torch.zeros((rows, w, 4), dtype=torch.uint8, device="cuda"), torch.zeros((rows, w, 4), dtype=torch.int16, device="cuda"),
                      torch.zeros((rows, w, 2), dtype=torch.int16, device="cuda"), torch.zeros((rows, w, 4), dtype=torch.float32, device="cuda"),
                      torch.zeros((rows, w), dtype=torch.float32, device="cuda")]
            torch.cuda.synchronize()          # torch fills on its own stream; the context's stream is not ordered against it
            self.ctx.raycast_gbuffer(c, tm, mt, *planes)
            self.gb.append(planes)
        self.ctx.synchronize()
        for s in (0, 1):
            self.ctx.bind_gbuffer(s, *self.gb[s])
        self.image = torch.zeros((rows, w, 4), dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()

    def uniforms(self, f, **over):
        cam, prev = self.cams[f & 1], self.cams[(f & 1) ^ 1] if f > 0 else self.cams[0]
        kw = dict(prevFrameProjectionViewMatrix=capi.camera_matrix(prev), cameraPos=(cam.position[0], cam.position[1], cam.position[2], 1.0),
                  screenSize=(self.w, self.h), frame=f + 1, initialLightSampleCount=32, temporalSampleCountMultiplier=20,
                  spatialPosThreshold=0.1, spatialNormalThreshold=25.0, spatialNeighbors=4, spatialRadius=30.0, flags=3)
        kw.update(over)
        lu = capi.make_lighting_uniforms(prevFrameProjectionViewMatrix=kw["prevFrameProjectionViewMatrix"], cameraPos=kw["cameraPos"],
                                         bufferSize=(self.w, self.h), debugMode=0, gamma=1.0)
        return capi.make_uniforms(**kw), lu

    def set(self, f, **over):
        u, lu = self.uniforms(f, **over)
        self.ctx.set_uniforms(u)
        self.ctx.set_lighting_uniforms(lu)

    def owned(self, reservoirs):
        """Rows [row_begin,row_end) of a download that covers [alloc_begin,alloc_end)."""
        return reservoirs[(self.rb - self.a0) * self.w:(self.re - self.a0) * self.w]


@pytest.mark.parametrize("unbiased", [True, False])
def test_full_size_determinism_and_counters(unbiased):
    """Sponza 1080p (C2/C3): two independent runs of a 3-frame sequence give identical bits; ray counts are
    exactly 1/pixel (biased) or within [1, 1+N+1]/pixel (unbiased)."""
    scene, (pos, look) = _scene_full()
    w, h = 1920, 1080
    finals = []
    for _ in range(2):
        d = DeviceFrames(scene, pos, look, w, h)
        d.ctx.counters(reset=True)
        for f in range(3):
            d.set(f)
            d.ctx.frame(f & 1, unbiased, 1)
        c = d.ctx.counters()
        finals.append(d.ctx.download_reservoirs(0))
        assert c["stack_overflows"] == 0 and c["halo_misses"] == 0
        if unbiased:
            assert 3 * w * h * 2 <= c["shadow_rays"] <= 3 * w * h * 5
        else:
            assert c["shadow_rays"] == 3 * w * h
        d.ctx.close()
    assert np.array_equal(finals[0].view(np.uint8), finals[1].view(np.uint8))
    r = finals[0]
    assert (r["w"] > 0).mean() > 0.3 and np.isfinite(r["w"]).all()


@pytest.mark.parametrize("unbiased,iterations,fmt", [(True, 1, "rgba8"), (True, 1, "f32"), (False, 1, "rgba8"), (False, 2, "f32"), (False, 0, "f32")])
def test_frame_lit_equals_frame_then_lighting(unbiased, iterations, fmt):
    """restir_frame_lit (lighting fused into the kernel that produces the final reservoir) against restir_frame followed by
    restir_pass_lighting: same reservoirs, same image, bit for bit, over a 3-frame sequence with temporal history."""
    torch = _torch()
    scene, (pos, look) = _scene_full()
    w, h = 640, 360
    dtype, out_format = (torch.uint8, capi.RESTIR_OUT_RGBA8_SRGB) if fmt == "rgba8" else (torch.float32, capi.RESTIR_OUT_RGBA32F)
    images, finals, launches = [], [], []
    for fused in (False, True):
        d = DeviceFrames(scene, pos, look, w, h)
        d.ctx.set_unbiased_neighbors(5)
        img = torch.zeros((h, w, 4), dtype=dtype, device="cuda")
        torch.cuda.synchronize()
        d.ctx.counters(reset=True)
        for f in range(3):
            d.set(f)
            if fused:
                d.ctx.frame_lit(f & 1, unbiased, iterations, img, out_format)
            else:
                d.ctx.frame(f & 1, unbiased, iterations)
                d.ctx.pass_lighting(f & 1, f & 1, img, out_format)
        d.ctx.synchronize()
        launches.append(d.ctx.counters()["kernel_launches"])
        images.append(img.cpu().numpy().copy())
        finals.append(d.ctx.download_reservoirs(0))
        d.ctx.close()
    assert np.array_equal(finals[0].view(np.uint8), finals[1].view(np.uint8))
    assert np.array_equal(images[0].view(np.uint8), images[1].view(np.uint8))
    assert images[0].view(np.uint8).any()
    if unbiased or iterations > 0:
        assert launches[1] == launches[0] - 3    # one kernel per frame less


@pytest.mark.parametrize("size,neighbors,radius", [((640, 360), 4, 30.0), ((333, 97), 5, 30.0), ((640, 360), 4, 12.5), ((64, 40), 16, 30.0)])
def test_staged_spatial_pass_is_bit_identical(size, neighbors, radius):
    """restir_set_spatial_staging: the biased spatial pass with its gate data staged in shared memory (tile + 31-pixel apron)
    against the direct kernel — same reservoirs, same image — on frames whose tiles hang over every screen edge, and as two
    bands (the staged rectangle reaches into the halo rows)."""
    torch = _torch()
    scene, (pos, look) = _scene_full()
    w, h = size
    results = []
    for staged in (False, True):
        d = DeviceFrames(scene, pos, look, w, h)
        d.ctx.set_spatial_staging(staged)
        img = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        for f in range(3):
            d.set(f, spatialNeighbors=neighbors, spatialRadius=radius)
            d.ctx.frame_lit(f & 1, False, 1, img, capi.RESTIR_OUT_RGBA32F)
        d.ctx.synchronize()
        results.append((d.ctx.download_reservoirs(0), img.cpu().numpy().copy()))
        d.ctx.close()
    assert np.array_equal(results[0][0].view(np.uint8), results[1][0].view(np.uint8))
    assert np.array_equal(results[0][1].view(np.uint8), results[1][1].view(np.uint8))
    if h >= 97:
        halo = 40
        bounds = [0, h // 2 + 3, h]
        parts = [DeviceFrames(scene, pos, look, w, h, band=(bounds[r], bounds[r + 1]), halo=halo) for r in range(2)]
        for r, part in enumerate(parts):
            part.ctx.set_spatial_staging(True)
            part.ctx.band_connect(0, parts[0].ctx.band_local_peer() if r == 1 else None)
            part.ctx.band_connect(1, parts[1].ctx.band_local_peer() if r == 0 else None)
        for f in range(3):
            for part in parts:
                part.set(f, spatialNeighbors=neighbors, spatialRadius=radius)
                part.ctx.frame(f & 1, False, 1)
        got = np.concatenate([part.owned(part.ctx.download_reservoirs(0)) for part in parts])
        for part in parts:
            c = part.ctx.counters()
            assert c["halo_misses"] == 0
            part.ctx.close()
        assert np.array_equal(got.view(np.uint8), results[0][0].view(np.uint8))


def test_degenerate_parameters_are_identities():
    """spatialNeighbors = 0 makes the spatial pass a copy; no flags => no rays and temporal off equals a
    first frame; M after pass 1 equals the candidate count on every non-background pixel."""
    scene, (pos, look) = _scene_full()
    w, h = 1920, 1080
    d = DeviceFrames(scene, pos, look, w, h)
    d.set(0, flags=0, spatialNeighbors=0)
    d.ctx.counters(reset=True)
    d.ctx.pass_restir(0, 0, 1)
    a = d.ctx.download_reservoirs(0)
    d.ctx.pass_spatial(0, 0, 1, 0)
    b = d.ctx.download_reservoirs(1)
    assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
    assert d.ctx.counters()["shadow_rays"] == 0
    albedo_a = d.gb[0][0].cpu().numpy()[..., 3].reshape(-1)
    normal = d.gb[0][1].cpu().numpy()[..., :3].reshape(-1, 3)
    surface = (normal != 0).any(axis=1)
    assert (a["M"][surface] == 32).all() and (a["M"][~surface] == 0).all()
    assert (albedo_a[~surface] == 255).all()
    # temporal flag with an all-zero history: identical to temporal off, except M (zero history adds M = 0)
    d.set(0, flags=2, spatialNeighbors=0)
    d.ctx.upload_reservoirs(1, np.zeros(w * h, capi.RESERVOIR_DTYPE))
    d.ctx.pass_restir(0, 0, 1)
    c = d.ctx.download_reservoirs(0)
    same_sample = (c["lightIndex"] == a["lightIndex"]) & (c["M"] == a["M"])
    assert same_sample.all()
    d.ctx.close()


@pytest.mark.parametrize("unbiased,bounds", [(True, None), (False, None), (True, [0, 100, 290, 432]), (False, [0, 150, 250, 432])])
def test_band_split_is_bit_identical_to_whole_frame(unbiased, bounds):
    """The multi-GPU claim on one GPU: band contexts (two of equal height, or three of unequal height as
    bands.balanced_bounds cuts them) with halo copies between the passes reproduce the single-context frame exactly
    (RNG is keyed on global pixel coordinates).  The halo is sized like bench.py does: spatial reach, or the rows
    temporal reprojection reaches if that is more (bands.temporal_row_reach)."""
    torch = _torch()
    scene, (pos, look) = _scene_full()
    w, h, halo, frames = 1920, 432, 31, 3
    whole = DeviceFrames(scene, pos, look, w, h)
    for f in range(frames):
        whole.set(f)
        whole.ctx.frame(f & 1, unbiased, 1)
    want = whole.ctx.download_reservoirs((frames - 1) & 1)
    whole.ctx.close()

    world = 2 if bounds is None else len(bounds) - 1
    parts = [DeviceFrames(scene, pos, look, w, h, band=bands.band_rows(h, world, r, bounds), halo=halo) for r in range(world)]
    reach = 0
    for part in parts:
        for cur, prv in ((0, 1), (1, 0)):
            reach = max(reach, bands.temporal_row_reach(part.gb[cur][3], part.gb[cur][1], capi.camera_matrix(part.cams[prv]), w, h,
                                                        part.a0, part.rb, part.re, torch))
    if reach > halo:
        halo = reach
        for part in parts:
            part.ctx.close()
        parts = [DeviceFrames(scene, pos, look, w, h, band=bands.band_rows(h, world, r, bounds), halo=halo) for r in range(world)]
    plans = [bands.halo_plan(h, world, r, halo, bounds) for r in range(world)]

    def exchange(buffer):
        views = [bands.reservoir_rows_tensor(p.ctx, buffer, torch) for p in parts]
        for p in parts:
            p.ctx.synchronize()
        for r, plan in enumerate(plans):
            for peer, (s0, s1), _ in plan:   # deliver what rank r sends into the peer's halo rows
                views[peer][s0 - parts[peer].a0: s1 - parts[peer].a0].copy_(views[r][s0 - parts[r].a0: s1 - parts[r].a0])
        torch.cuda.synchronize()

    for f in range(frames):
        i, p_ = f & 1, (f & 1) ^ 1
        for part in parts:
            part.set(f)
        if unbiased:
            for part in parts:
                part.ctx.pass_restir(i, capi.RESTIR_BUF_TEMP, p_)
            exchange(capi.RESTIR_BUF_TEMP)
            for part in parts:
                part.ctx.pass_unbiased(i, capi.RESTIR_BUF_TEMP, i)
        else:
            for part in parts:
                part.ctx.pass_restir(i, i, p_)
            exchange(i)
            for part in parts:
                part.ctx.pass_spatial(i, i, p_, 0)
            exchange(p_)
            for part in parts:
                part.ctx.pass_spatial(i, p_, i, 1)
        exchange(i)
    got = np.concatenate([part.owned(part.ctx.download_reservoirs((frames - 1) & 1)) for part in parts])
    for part in parts:
        c = part.ctx.counters()
        assert c["halo_misses"] == 0
        part.ctx.close()
    assert ph.compare_reservoirs(got, want, "band split") == 0


@pytest.mark.parametrize("unbiased,bounds", [(True, [0, 216, 432]), (False, [0, 216, 432]), (True, [0, 100, 290, 432]), (False, [0, 150, 250, 432])])
def test_connected_bands_exchange_halos_themselves(unbiased, bounds):
    """restir_band_connect: band contexts wired to their neighbours push their boundary rows into the neighbour's
    buffers with the library's own kernels and wait for the neighbour's rows on the device — no copies by the caller,
    restir_frame works on a band — and reproduce the single-context frame bit for bit."""
    torch = _torch()
    scene, (pos, look) = _scene_full()
    w, h, halo, frames = 1920, 432, 31, 3
    whole = DeviceFrames(scene, pos, look, w, h)
    for f in range(frames):
        whole.set(f)
        whole.ctx.frame(f & 1, unbiased, 1)
    want = whole.ctx.download_reservoirs((frames - 1) & 1)
    whole.ctx.close()

    world = len(bounds) - 1
    parts = [DeviceFrames(scene, pos, look, w, h, band=bands.band_rows(h, world, r, bounds), halo=halo) for r in range(world)]
    reach = 0
    for part in parts:
        for cur, prv in ((0, 1), (1, 0)):
            reach = max(reach, bands.temporal_row_reach(part.gb[cur][3], part.gb[cur][1], capi.camera_matrix(part.cams[prv]), w, h,
                                                        part.a0, part.rb, part.re, torch))
    if reach > halo:
        halo = reach
        for part in parts:
            part.ctx.close()
        parts = [DeviceFrames(scene, pos, look, w, h, band=bands.band_rows(h, world, r, bounds), halo=halo) for r in range(world)]
    for r, part in enumerate(parts):       # same process: the neighbours' raw pointers
        part.ctx.band_connect(0, parts[r - 1].ctx.band_local_peer() if r > 0 else None)
        part.ctx.band_connect(1, parts[r + 1].ctx.band_local_peer() if r + 1 < world else None)
    with pytest.raises(capi.RestirError):  # a peer that does not touch this band's edge is refused
        parts[0].ctx.band_connect(0, parts[-1].ctx.band_local_peer())
    parts[0].ctx.band_connect(0, None)
    for f in range(frames):
        for part in parts:
            part.set(f)
            part.ctx.frame(f & 1, unbiased, 1)   # asynchronous: a band waiting for its neighbour does not block the host
    got = np.concatenate([part.owned(part.ctx.download_reservoirs((frames - 1) & 1)) for part in parts])
    for part in parts:
        c = part.ctx.counters()
        assert c["halo_misses"] == 0 and c["halo_wait_timeouts"] == 0
        part.ctx.close()
    assert ph.compare_reservoirs(got, want, "connected bands") == 0


@pytest.mark.parametrize("unbiased,world", [(True, 2), (False, 3)])
def test_band_processes_over_cuda_ipc(tmp_path, unbiased, world):
    """The path `torchrun bench.py --gpus N` times: one PROCESS per band, neighbours mapped through CUDA IPC handles
    (restir_band_export_ipc / _open_ipc / _connect), halos pushed and awaited by the library's kernels.  The processes share
    this one GPU (tests/band_ipc_worker.py); each compares the reservoirs and the RGBA8 rows it owns with a whole-screen
    context, bit for bit."""
    import socket
    import sys

    _torch()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "band_ipc_worker.py")
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK="0", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, worker, str(tmp_path / f"rank{r}.json"), "1" if unbiased else "0"], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o[-3000:]
    res = [json.load(open(tmp_path / f"rank{r}.json")) for r in range(world)]
    assert len({r["pid"] for r in res}) == world and sum(r["pixels"] for r in res) == 1920 * 432
    for r in res:
        assert r["mismatching_reservoirs"] == 0 and r["mismatching_pixels"] == 0, res
        assert r["halo_misses"] == 0 and r["halo_wait_timeouts"] == 0, res


def test_band_connect_rejects_a_halo_taller_than_the_neighbour():
    """A neighbour pushes only rows it shades itself: when this band's halo reaches past the neighbouring band (rows of a
    band two hops away), those rows would stay empty without any halo_miss — restir_band_connect refuses the connection."""
    scene = fixtures.make_procedural(seed=4, grid=8, boxes=10, lights="point", n_point_lights=9)
    ctxs = [ph.make_context(scene) for _ in range(3)]
    w, h, halo = 64, 96, 40
    for r, (b, e) in enumerate(((0, 32), (32, 64), (64, 96))):     # 32-row bands under a 40-row halo
        ctxs[r].resize_band(w, h, b, e, halo)
    with pytest.raises(capi.RestirError, match="halo"):
        ctxs[0].band_connect(1, ctxs[1].band_local_peer())
    with pytest.raises(capi.RestirError, match="halo"):
        ctxs[2].band_connect(0, ctxs[1].band_local_peer())
    for r, (b, e) in enumerate(((0, 32), (32, 64), (64, 96))):     # a 32-row halo fits
        ctxs[r].resize_band(w, h, b, e, 32)
    ctxs[0].band_connect(1, ctxs[1].band_local_peer())
    ctxs[1].band_connect(0, ctxs[0].band_local_peer())
    for c in ctxs:
        c.close()


def test_oracle_spot_check_at_1080p():
    """A 24-row band of the full-size Sponza frame 2 (temporal history from the GPU's own frame 1) through
    the oracle, compared bit for bit — parity at the BASELINE size without running the oracle on 2 Mpx."""
    scene, (pos, look) = _scene_full()
    po = ph.oracle()
    w, h = 1920, 1080
    d = DeviceFrames(scene, pos, look, w, h)
    d.set(0)
    d.ctx.frame(0, True, 1)
    prev = d.ctx.download_reservoirs(0)
    d.set(1)
    d.ctx.pass_restir(1, capi.RESTIR_BUF_TEMP, 0)
    gpu_initial = d.ctx.download_reservoirs(capi.RESTIR_BUF_TEMP)
    d.ctx.pass_unbiased(1, capi.RESTIR_BUF_TEMP, 1)
    gpu_final = d.ctx.download_reservoirs(1)
    y0, y1 = 528, 552
    a0, a1 = y0 - 31, y1 + 31
    types = (np.uint8, np.int16, np.uint16, np.float32, np.float32)
    g = [po.GBuffer(w, h, *[p.cpu().numpy().view(t) for p, t in zip(planes, types)]) for planes in d.gb]
    u, _ = d.uniforms(1)
    sc = ph.oracle_scene(scene)
    initial, _ = po.restir_pass(sc, u.astype(po.UNIFORMS_DTYPE), g[1], g[0], prev.astype(po.RESERVOIR_DTYPE), (a0, a1))
    final, _ = po.unbiased_pass(sc, u.astype(po.UNIFORMS_DTYPE), g[1], initial, 3, (y0, y1))
    assert ph.compare_reservoirs(gpu_initial[a0 * w: a1 * w], initial[a0 * w: a1 * w], "1080p initial") == 0
    assert ph.compare_reservoirs(gpu_final[y0 * w: y1 * w], final[y0 * w: y1 * w], "1080p final") == 0
    d.ctx.close()


def test_cpp_host_driver_matches_python_driven_run(tmp_path):
    """The C++ driver (host/restir_driver.cpp: passes.hpp over the C ABI) produces the same reservoirs and image
    as the same sequence driven through ctypes."""
    torch = _torch()
    scene = fixtures.make_procedural(seed=4, grid=8, boxes=10, lights="tri")
    d = str(tmp_path)
    scene.triangles.tofile(os.path.join(d, "triangles.bin"))
    scene.tri_material.tofile(os.path.join(d, "tri_material.i32"))
    scene.materials.tofile(os.path.join(d, "materials.f32"))
    scene.dims.tofile(os.path.join(d, "dims.f32"))
    scene.material_table().tofile(os.path.join(d, "material_table.u32"))
    pos, look = (3.0, 3.5, 4.2), (0.0, -1.0, 0.0)
    np.array(pos + look, np.float32).tofile(os.path.join(d, "camera.f32"))
    w, h, frames = 200, 120, 3
    exe = os.path.join(ph.ROOT, "restir-vulkan_b200", "restir_driver")
    out = subprocess.run([exe, d, str(w), str(h), str(frames), "1", "3"], check=True, capture_output=True, text=True).stdout
    res = json.loads(out.strip().splitlines()[-1])

    df = DeviceFrames(scene, pos, look, w, h)
    img = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    df.ctx.counters(reset=True)
    for f in range(frames):
        df.set(f)
        df.ctx.frame(f & 1, True, 1)
        df.ctx.pass_lighting(f & 1, f & 1, img, capi.RESTIR_OUT_RGBA8_SRGB)
    df.ctx.synchronize()
    reservoirs = df.ctx.download_reservoirs((frames - 1) & 1)

    def fnv(b):
        hsh = 1469598103934665603
        for x in np.frombuffer(b, np.uint8).tolist():
            hsh = ((hsh ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        return f"{hsh:016x}"

    assert res["reservoir_fnv1a"] == fnv(reservoirs.tobytes())
    assert res["image_fnv1a"] == fnv(img.cpu().numpy().tobytes())
    assert res["shadow_rays"] == df.ctx.counters()["shadow_rays"]
    df.ctx.close()

    # the same frames as row bands, entirely from C++ (restir::BandSet over restir_band_connect): three band contexts on
    # this GPU exchange their halos themselves and must reproduce the single-context checksums, biased and unbiased
    for unbiased in ("1", "0"):
        one = json.loads(subprocess.run([exe, d, str(w), str(h), str(frames), unbiased, "3"], check=True, capture_output=True,
                                        text=True).stdout.strip().splitlines()[-1])
        banded = json.loads(subprocess.run([exe, d, str(w), str(h), str(frames), unbiased, "3", "--bands", "3", "--halo", "36"], check=True,
                                           capture_output=True, text=True).stdout.strip().splitlines()[-1])
        assert banded["bands"] == 3 and banded["halo_misses"] == 0 and banded["halo_wait_timeouts"] == 0
        assert banded["reservoir_fnv1a"] == one["reservoir_fnv1a"] and banded["image_fnv1a"] == one["image_fnv1a"]
        assert banded["shadow_rays"] == one["shadow_rays"]


@pytest.mark.parametrize("unbiased", [True, False])
def test_cpp_driver_replays_a_capture(tmp_path, unbiased):
    """include/restir_capture.h end to end: a capture written from the oracle's frames (scene buffers, uniforms, G-buffers and
    the outputs of every pass) is replayed by the C++ driver through the C ABI; nothing may differ.  A capture dumped from
    the real Vulkan application is checked with the same command."""
    _torch()
    scene = fixtures.make_procedural(seed=11, grid=8, boxes=10, lights="tri" if unbiased else "point", n_point_lights=9)
    w, h = 72, 40
    cams = ph.moving_cameras(3, (3.0, 3.5, 4.2), (0.0, -1.0, 0.0), w / h)
    case = ph.Case(scene, w, h, cams, candidates=16, unbiased=unbiased, unbiased_neighbors=5 if unbiased else 3, spatial_iterations=1)
    want = ph.run_oracle(case)
    path = str(tmp_path / "frames.rsc")
    ph.make_capture(case, want).write(path)
    exe = os.path.join(ph.ROOT, "restir-vulkan_b200", "restir_driver")
    run = subprocess.run([exe, "--replay", path], capture_output=True, text=True)
    assert run.returncode == 0, run.stdout + run.stderr
    res = json.loads(run.stdout.strip().splitlines()[-1])
    assert res["frames"] == 3 and res["expected"] == 7
    assert res["mismatching_initial_reservoirs"] == 0 and res["mismatching_final_reservoirs"] == 0 and res["mismatching_pixels"] == 0
    assert res["shadow_rays"] == sum(f["rays"] for f in want)
    # and a capture whose expected outputs were tampered with is reported, not accepted
    want[1]["reservoirs"]["M"][5] += 1
    ph.make_capture(case, want).write(path)
    run = subprocess.run([exe, "--replay", path], capture_output=True, text=True)
    assert run.returncode == 5 and json.loads(run.stdout.strip().splitlines()[-1])["mismatching_final_reservoirs"] == 1
