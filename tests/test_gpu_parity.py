"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same inputs.

Bit-exact for visibility bits, integer and float reservoir fields (NaN == NaN); final RGB within
rel 1e-3 / PSNR >= 60 dB (see parity_harness.py).  Sizes are chosen so the oracle finishes in seconds;
full-size behaviour is covered by the property tests in test_gpu_properties.py.
"""
import numpy as np
import pytest

import parity_harness as ph

pytestmark = pytest.mark.gpu
capi, fixtures = ph.capi, ph.fixtures


def _torch():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def _scene(name):
    if name.startswith("procedural"):
        kind = name.split(":")[1]
        return fixtures.make_procedural(seed=21, grid=10, boxes=20, lights=kind, n_point_lights=24)
    if not fixtures.baked_available(name):
        pytest.skip(f"scenes/_baked/{name} not present")
    return fixtures.load_baked(name, rebuild=True)


CAMERAS = {
    # position, lookAt — SURVEY.md §8d: camera.h defaults unless the scene needs another viewpoint
    "procedural": ((3.0, 3.5, 4.2), (0.0, -1.0, 0.0)),
    "cornellBox": ((3.0, 4.0, 5.0), (0.0, 0.0, 0.0)),
    "sponza": ((3.0, 4.0, 5.0), (0.0, 0.0, 0.0)),
    "office": ((3.0, 1.7, 0.5), (3.0, 1.5, -5.0)),
}


def _cams(name, n, w, h):
    pos, look = CAMERAS[name.split(":")[0]]
    return ph.moving_cameras(n, pos, look, w / h)


# ---- minimum slice: testVisibility ---------------------------------------------------------------------

@pytest.mark.parametrize("name,n_rays", [("procedural:point", 400_000), ("cornellBox", 600_000), ("sponza", 1_500_000), ("office", 600_000)])
def test_shadow_rays_bit_exact(name, n_rays):
    torch = _torch()
    scene = _scene(name)
    po = ph.oracle()
    rng = np.random.default_rng(1234)
    # half the segments go from visible surface points to light-like positions (what the passes trace),
    # half are uniform inside the scene bounds (long, grazing, degenerate directions)
    w, h = 160, 90
    cam = _cams(name, 1, w, h)[0]
    g = po.raycast_gbuffer(ph.oracle_scene(scene), scene.tri_material, scene.material_table(), cam, w, h)
    surf = g.world_pos.reshape(-1, 4)[:, :3]
    lo, hi = scene.dims[:3], scene.dims[3:]
    n1 = n_rays // 2
    p1a = surf[rng.integers(0, surf.shape[0], n1)]
    p2a = rng.uniform(lo, hi, (n1, 3)).astype(np.float32)
    p1b = rng.uniform(lo, hi, (n_rays - n1, 3)).astype(np.float32)
    p2b = rng.uniform(lo, hi, (n_rays - n1, 3)).astype(np.float32)
    p2b[: 1000, 0] = p1b[: 1000, 0]          # axis-aligned: zero direction components (0 * inf, 0/0 paths)
    p2b[1000: 2000, 1:] = p1b[1000: 2000, 1:]
    p2b[2000: 2100] = p1b[2000: 2100]        # zero-length segments
    p1 = np.ascontiguousarray(np.concatenate([p1a, p1b]), np.float32)
    p2 = np.ascontiguousarray(np.concatenate([p2a, p2b]), np.float32)

    want, margin, overflow = po.trace_segments(ph.oracle_scene(scene), p1, p2, want_margin=True)
    ctx = ph.make_context(scene)
    d1, d2 = torch.from_numpy(p1).cuda(), torch.from_numpy(p2).cuda()
    out = torch.zeros(n_rays, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()   # torch fills on its own stream; the context's stream is not ordered against it
    ctx.counters(reset=True)
    ctx.trace_segments(d1, d2, n_rays, out)
    ctx.synchronize()
    c = ctx.counters()
    got = out.cpu().numpy()
    mism = np.flatnonzero(got != want)
    edge = margin < 1e-5
    print(f"{name}: {n_rays} rays, shadowed {want.mean():.3f}, edge-flagged {edge.mean():.4f}, mismatches {mism.size}")
    assert mism.size == 0, f"{mism.size} visibility bits differ (of which edge-flagged: {int(edge[mism].sum())})"
    assert c["shadow_rays"] == n_rays and c["stack_overflows"] == int(overflow.sum()) == 0
    assert 0.02 < want.mean() < 0.999    # the sample exercises both outcomes
    ctx.close()


def test_traversal_modes_agree_and_report():
    """The 4-wide quantised image, the 64-byte re-stride and the literal 80-byte walk of the same uploaded tree give the same bits; a tree
    whose child indices are out of range is rejected at upload instead of being traversed."""
    torch = _torch()
    scene = _scene("procedural:tri")
    rng = np.random.default_rng(5)
    n = 200_000
    lo, hi = scene.dims[:3], scene.dims[3:]
    p1 = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    p2 = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    d1, d2 = torch.from_numpy(p1).cuda(), torch.from_numpy(p2).cuda()
    outs = []
    for mode in (capi.RESTIR_TRAVERSAL_AUTO, capi.RESTIR_TRAVERSAL_IMAGE, capi.RESTIR_TRAVERSAL_REFERENCE_ORDER):
        ctx = capi.RestirContext(0)
        ctx.set_traversal(mode)
        ctx.upload_bvh(scene.nodes, scene.triangles)
        info = ctx.bvh_info()
        if mode == capi.RESTIR_TRAVERSAL_AUTO:
            assert info["traversal"] == capi.RESTIR_TRAVERSAL_WIDE and 0 < info["wide_nodes"] < scene.nodes.shape[0]
            assert info["reachable_nodes"] == scene.nodes.shape[0] and info["reference_stack_bound"] <= 32
        elif mode == capi.RESTIR_TRAVERSAL_IMAGE:
            assert info["traversal"] == capi.RESTIR_TRAVERSAL_IMAGE and info["wide_nodes"] == 0
        else:
            assert info["traversal"] == capi.RESTIR_TRAVERSAL_REFERENCE_ORDER
        out = torch.zeros(n, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        ctx.trace_segments(d1, d2, n, out)
        ctx.synchronize()
        outs.append(out.cpu().numpy())
        ctx.close()
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    want = ph.oracle().trace_segments(ph.oracle_scene(scene), p1, p2)
    assert np.array_equal(outs[0], want)
    bad = scene.nodes.copy()
    bad.view(np.int32).reshape(-1, 20)[0, 16] = bad.shape[0] + 7
    ctx = capi.RestirContext(0)
    with pytest.raises(capi.RestirError, match="out of range"):
        ctx.upload_bvh(bad, scene.triangles)
    ctx.close()


# ---- fixture tool ---------------------------------------------------------------------------------------

@pytest.mark.parametrize("name", ["procedural:tri", "cornellBox", "sponza"])
def test_gbuffer_fixture_tool_matches_oracle(name):
    torch = _torch()
    scene = _scene(name)
    w, h = 256, 144
    cam = _cams(name, 1, w, h)[0]
    po = ph.oracle()
    want = po.raycast_gbuffer(ph.oracle_scene(scene), scene.tri_material, scene.material_table(), cam, w, h)
    ctx = ph.make_context(scene)
    ctx.resize(w, h)
    tm = torch.from_numpy(np.ascontiguousarray(scene.tri_material)).cuda()
    mt = torch.from_numpy(scene.material_table().view(np.int32)).cuda()
    planes = [torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda"), torch.zeros((h, w, 4), dtype=torch.int16, device="cuda"),
              torch.zeros((h, w, 2), dtype=torch.int16, device="cuda"), torch.zeros((h, w, 4), dtype=torch.float32, device="cuda"),
              torch.zeros((h, w), dtype=torch.float32, device="cuda")]
    torch.cuda.synchronize()
    ctx.raycast_gbuffer(ph.to_capi_camera(cam), tm, mt, *planes)
    ctx.synchronize()
    for got, ref, what in zip(planes, want.planes(), ["albedo", "normal", "material", "worldPos", "depth"]):
        a = got.cpu().numpy().view(np.uint8).reshape(-1)
        b = ref.view(np.uint8).reshape(-1)
        assert np.array_equal(a, b), f"{what}: {(a != b).sum()} bytes differ"
    assert (want.albedo[..., 3] == 0).mean() > 0.3   # the view actually sees geometry
    ctx.close()


# ---- the four passes, frame sequences ---------------------------------------------------------------

FRAME_CASES = [
    # name, (w, h), frames, kwargs
    ("procedural:point", (192, 108), 3, dict(unbiased=False, spatial_iterations=1)),
    ("procedural:point", (192, 108), 3, dict(unbiased=True)),
    ("procedural:tri", (192, 108), 3, dict(unbiased=False, spatial_iterations=2)),
    ("procedural:tri", (192, 108), 3, dict(unbiased=True, unbiased_neighbors=5)),
    ("procedural:random", (160, 90), 2, dict(unbiased=True, candidates=64)),
    ("procedural:point", (160, 90), 2, dict(unbiased=False, flags=0)),                       # no visibility, no temporal
    ("procedural:point", (160, 90), 3, dict(unbiased=True, flags=2)),                        # temporal only
    ("procedural:point", (97, 61), 3, dict(unbiased=False, neighbors=5, gamma=2.2)),         # ragged tile edges, gamma != 1
    ("cornellBox", (320, 180), 3, dict(unbiased=False, spatial_iterations=1)),               # C1 (reduced size)
    ("sponza", (320, 180), 3, dict(unbiased=False, spatial_iterations=1, neighbors=4)),      # C2
    ("sponza", (320, 180), 3, dict(unbiased=True, unbiased_neighbors=3)),                    # C3 (reference: 3 neighbours)
    ("sponza", (256, 144), 2, dict(unbiased=True, unbiased_neighbors=5)),                    # C3 (north-star: 5)
    ("office", (256, 144), 2, dict(unbiased=True)),                                          # C4 (reduced size)
]


@pytest.mark.parametrize("name,size,frames,kw", FRAME_CASES, ids=[f"{c[0]}-{c[1][0]}x{c[1][1]}-{i}" for i, c in enumerate(FRAME_CASES)])
def test_frames_match_oracle(name, size, frames, kw):
    _torch()
    scene = _scene(name)
    w, h = size
    case = ph.Case(scene, w, h, _cams(name, frames, w, h), **kw)
    got = ph.run_cuda(case)
    want = ph.run_oracle(case)
    ph.assert_frames_match(got, want, name)
    last = want[-1]["reservoirs"]
    lit = (last["w"] > 0).mean()
    print(f"{name} {w}x{h}: {frames} frames identical; w>0 on {lit:.2%} of pixels; rays/frame {want[-1]['rays']}")
    assert lit > 0.05, "degenerate case: almost nothing is lit"


@pytest.mark.parametrize("label,name,size,frames,kw", ph.EDGE_CASES, ids=[c[0] for c in ph.EDGE_CASES])
def test_edge_cases_match_oracle(label, name, size, frames, kw):
    """Boundary values of the parameter block and of the screen (parity_harness.EDGE_CASES): 1x1 and one-tile screens,
    ragged tile edges, one candidate, radius 0 and radius >> screen, M cap 0 and 1, thresholds that reject every
    neighbour, frame numbers that wrap the 32-bit seed arithmetic, each visibility / temporal flag alone."""
    _torch()
    scene = _scene(name)
    w, h = size
    case = ph.Case(scene, w, h, _cams(name, frames, max(w, 2), max(h, 2)), **kw)
    ph.assert_frames_match(ph.run_cuda(case), ph.run_oracle(case), label)


@pytest.mark.parametrize("unbiased", [True, False])
@pytest.mark.parametrize("kind", ph.DEGENERATE_LIGHTS)
def test_degenerate_lights_match_oracle(kind, unbiased):
    """NaN and infinity through the reservoirs (parity_harness.degenerate_light_case: a light exactly on a visible surface
    point, zero-probability lights, luminances that overflow p-hat): the kernels follow the oracle bit for bit, any NaN
    equal to any NaN — the same cases hold between the oracle and the reference's shader text on the CPU."""
    _torch()
    case = ph.degenerate_light_case(kind, unbiased=unbiased)
    got, want = ph.run_cuda(case), ph.run_oracle(case)
    for f, (c, o) in enumerate(zip(got, want)):
        assert ph.compare_reservoirs(c["initial"], o["initial"], f"{kind} frame {f} after restirOmni") == 0
        assert ph.compare_reservoirs(c["reservoirs"], o["reservoirs"], f"{kind} frame {f} final") == 0
        assert c["rays"] == o["rays"]
        same = ph.bits_equal(c["rgba"][..., :3], o["rgba"][..., :3])
        assert same.all(), f"{kind} frame {f}: {(~same).sum()} colour values differ"


def test_zero_candidates_leave_empty_reservoirs():
    """initialLightSampleCount = 0: the candidate loop does not run (restirOmni.glsl:108), reservoirs stay as newReservoir
    left them (oracle definition: zero), yet the passes run to the end and match the oracle."""
    _torch()
    scene = _scene("procedural:point")
    w, h = 48, 27
    case = ph.Case(scene, w, h, _cams("procedural:point", 2, w, h), unbiased=True, candidates=0)
    got, want = ph.run_cuda(case), ph.run_oracle(case)
    ph.assert_frames_match(got, want, "zero candidates")
    assert (got[-1]["reservoirs"]["M"] == 0).all() and (got[-1]["reservoirs"]["w"] == 0).all()


def test_many_lights_gather_matches_oracle():
    """C5-like: 100k random point lights, 64 candidates (light tables no longer fit L1)."""
    _torch()
    base = _scene("procedural:random")
    scene = fixtures.with_random_point_lights(base, 100_000)
    w, h = 160, 90
    case = ph.Case(scene, w, h, _cams("procedural", 2, w, h), candidates=64, unbiased=True)
    ph.assert_frames_match(ph.run_cuda(case), ph.run_oracle(case), "many-lights")


@pytest.mark.parametrize("name", ["procedural:point", "procedural:tri", "sponza"])
def test_ray_elision_is_exact(name):
    """The unbiased pass answers some neighbour rays without walking the tree (the pixel's own ray is shadowed; the
    segment is bit-identical to the neighbour's own ray).  With the shortcut off every ray is walked: both settings
    must give the oracle's bits and the oracle's testVisibility count, and only the number of walks may differ."""
    _torch()
    scene = _scene(name)
    w, h = 224, 126
    case = ph.Case(scene, w, h, _cams(name, 3, w, h), unbiased=True, unbiased_neighbors=5)
    want = ph.run_oracle(case)
    walked = {}
    for enable in (1, 0):
        ctx = ph.make_context(scene)
        ctx.set_ray_elision(enable)
        ctx.set_occluder_cache(enable)                            # the other way of answering a ray without a walk (below)
        got = ph.run_cuda(case, ctx)
        ctx.close()
        ph.assert_frames_match(got, want, f"{name} elision={enable}")
        walked[enable] = sum(f["counters"]["shadow_rays_traced"] for f in got)
        asked = sum(f["counters"]["shadow_rays"] for f in got)
        assert walked[enable] <= asked
        if not enable:
            assert walked[enable] == asked and sum(f["counters"]["shadow_rays_cached"] for f in got) == 0
    print(f"{name}: {walked[0]} rays asked for, {walked[1]} walked with the exact shortcuts on")
    assert walked[1] < walked[0]
    # the experiment on top (restir_set_ray_elision(2)): one walk per distinct neighbour segment
    ctx = ph.make_context(scene)
    ctx.set_occluder_cache(0)
    got = ph.run_cuda(case, ctx)
    ctx.close()
    walked_no_table = sum(f["counters"]["shadow_rays_traced"] for f in got)
    ctx = ph.make_context(scene)
    ctx.set_ray_elision(2)
    ctx.set_occluder_cache(0)
    got = ph.run_cuda(case, ctx)
    ctx.close()
    ph.assert_frames_match(got, want, f"{name} elision=2 (segment table)")
    walked_table = sum(f["counters"]["shadow_rays_traced"] for f in got)
    print(f"{name}: {walked_no_table} walks without the segment table, {walked_table} with it")
    assert walked_table < walked_no_table < walked[0]


@pytest.mark.parametrize("name", ["procedural:point", "procedural:tri", "sponza", "office"])
def test_occluder_cache_is_exact(name):
    """The trace kernel tests the last triangle that occluded a ray of the same screen region at the same light before it queues
    a ray for a walk (restir_trace.cu).  The table only proposes witnesses: with it on, off, and on with entries left behind by a
    DIFFERENT view of the scene (stale witnesses), every reservoir and the testVisibility count are the oracle's."""
    _torch()
    scene = _scene(name)
    w, h = 224, 126
    case = ph.Case(scene, w, h, _cams(name, 3, w, h), unbiased=True, unbiased_neighbors=5)
    want = ph.run_oracle(case)
    ctx = ph.make_context(scene)
    got = ph.run_cuda(case, ctx)
    ph.assert_frames_match(got, want, f"{name} cache on")
    cached = sum(f["counters"]["shadow_rays_cached"] for f in got)
    walked_on = sum(f["counters"]["shadow_rays_traced"] for f in got)
    # the same context again: the table now holds the witnesses of the last frame of the first run (another camera position)
    got = ph.run_cuda(case, ctx)
    ph.assert_frames_match(got, want, f"{name} cache on, stale entries")
    ctx.set_occluder_cache(3)                                     # entries chosen by the segment's direction (the many-light default)
    got = ph.run_cuda(case, ctx)
    got = ph.run_cuda(case, ctx)
    ph.assert_frames_match(got, want, f"{name} cache by direction")
    assert sum(f["counters"]["shadow_rays_cached"] for f in got) > 0
    ctx.set_occluder_cache(0)
    got = ph.run_cuda(case, ctx)
    ph.assert_frames_match(got, want, f"{name} cache off")
    assert sum(f["counters"]["shadow_rays_cached"] for f in got) == 0
    walked_off = sum(f["counters"]["shadow_rays_traced"] for f in got)
    ctx.close()
    print(f"{name}: {walked_off} walks without the cache, {walked_on} with it ({cached} rays answered by a cached witness)")
    # (the segment table's probe limit makes the number of aliases depend, marginally, on the order items arrive in)
    assert cached > 0 and abs(walked_on + cached - walked_off) <= 0.002 * walked_off


# ---- boundary behaviour ---------------------------------------------------------------------------------

def test_reservoir_upload_download_round_trip():
    _torch()
    scene = _scene("procedural:tri")
    w, h = 128, 72
    case = ph.Case(scene, w, h, _cams("procedural", 2, w, h), unbiased=False)
    want = ph.run_oracle(case)[-1]["reservoirs"]
    ctx = ph.make_context(scene)
    ctx.resize(w, h)
    ctx.upload_reservoirs(capi.RESTIR_BUF_FRAME1, want)
    got = ctx.download_reservoirs(capi.RESTIR_BUF_FRAME1)
    assert ph.compare_reservoirs(got, want, "round trip") == 0
    # freshly resized buffers are zero-filled (app.h:276,283)
    zero = ctx.download_reservoirs(capi.RESTIR_BUF_TEMP)
    assert not zero.view(np.uint8).any()
    ctx.close()


def test_errors_are_returned_not_aborted():
    _torch()
    scene = _scene("procedural:point")
    ctx = capi.RestirContext(0)
    with pytest.raises(capi.RestirError, match="restir_resize"):
        ctx.pass_restir(0, 0, 1)
    ctx.resize(64, 32)
    with pytest.raises(capi.RestirError, match="not bound"):
        ctx.pass_restir(0, 0, 1)
    g = ph.oracle().GBuffer(64, 32)
    ctx.upload_gbuffer(0, *g.planes())
    with pytest.raises(capi.RestirError, match="upload_bvh"):
        ctx.pass_restir(0, 0, 1)
    ctx.upload_bvh(scene.nodes, scene.triangles)
    ctx.upload_lights(scene.point_blob, scene.tri_blob, scene.alias_blob)
    with pytest.raises(capi.RestirError, match="set_uniforms"):
        ctx.pass_restir(0, 0, 1)
    ctx.set_uniforms(capi.make_uniforms(screenSize=(32, 32)))
    with pytest.raises(capi.RestirError, match="does not match"):
        ctx.pass_restir(0, 0, 1)
    ctx.set_uniforms(capi.make_uniforms(screenSize=(64, 32), initialLightSampleCount=4, flags=3, frame=1))
    with pytest.raises(capi.RestirError, match="must differ"):
        ctx.pass_restir(0, 0, 0)
    with pytest.raises(capi.RestirError, match="debugMode"):
        ctx.set_lighting_uniforms(capi.make_lighting_uniforms(bufferSize=(64, 32), debugMode=3, gamma=1.0))
    ctx.pass_restir(0, 0, 1)      # and the context still works afterwards
    ctx.synchronize()
    bad_alias = scene.alias_blob.copy()
    bad_alias[:4] = np.array([3], np.int32).view(np.uint8)
    with pytest.raises(capi.RestirError, match="alias table"):
        ctx.upload_lights(scene.point_blob, scene.tri_blob, bad_alias[: 16 + 48])
    ctx.close()


def test_external_memory_import_error_path():
    """restir_import_external_memory with descriptors that are not exported allocations: an error code and a message, no abort, and
    the context stays usable (the import of a real Vulkan allocation cannot be exercised in this image: no Vulkan)."""
    import os

    _torch()
    scene = _scene("procedural:point")
    ctx = ph.make_context(scene)
    with pytest.raises(capi.RestirError):
        ctx.import_external_memory(-1, 4096)
    r, w = os.pipe()
    try:
        with pytest.raises(capi.RestirError, match="cudaImportExternalMemory|cudaExternalMemoryGetMappedBuffer"):
            ctx.import_external_memory(os.dup(r), 1 << 20)
    finally:
        os.close(r)
        os.close(w)
    with pytest.raises(capi.RestirError):
        ctx.release_external_memory(12345)
    ctx.resize(32, 16)                      # not sticky: the context still works
    assert ctx.reservoir_bytes() == 64
    ctx.close()
