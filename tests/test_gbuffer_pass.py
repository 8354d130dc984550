"""The G-buffer pass (restir_upload_geometry / restir_upload_materials / restir_pass_gbuffer).

CPU: the oracle twin (oracle_gbuffer_pass) against the fixture producer and against hand-made cases of gBuffer.frag's
branches.  GPU (-m gpu): the CUDA pass against the oracle twin, every plane bit for bit — procedural inputs that reach every
branch, Sponza (textures, normal maps, alpha-masked foliage) and office (specular-glossiness materials) — and a whole frame
sequence whose G-buffers never exist on the host.
"""
import numpy as np
import pytest

import parity_harness as ph

capi, fixtures = ph.capi, ph.fixtures

CAMERAS = {"sponza": ((3.0, 4.0, 5.0), (0.0, 0.0, 0.0)), "office": ((3.0, 1.7, 0.5), (3.0, 1.5, -5.0)), "cornellBox": ((3.0, 4.0, 5.0), (0.0, 0.0, 0.0)),
           "sponza_plants": ((-7.0, 1.2, 0.6), (-12.0, 1.0, -1.0))}


def _oracle_inputs(gi):
    po = ph.oracle()
    return po.GBufferInputs(gi.vertices, gi.indices, gi.draws, gi.matrices, gi.uniforms, gi.bindings, gi.textures)


def _procedural():
    scene = fixtures.make_procedural(seed=5, grid=8, boxes=14, lights="tri")
    return scene, fixtures.procedural_gbuffer_inputs(scene, seed=3)


def test_oracle_gbuffer_pass_agrees_with_the_fixture_producer_where_they_must():
    """cornellBox has no textures: world position and depth of the two CPU producers are the same arithmetic, the albedo /
    material codes are the same conversion (float32 here, float64 in fixtures.material_table: codes within 1), the
    normal is the interpolated vertex normal under the default normal texel instead of the geometric one."""
    if not fixtures.gbuffer_inputs_available("cornellBox"):
        pytest.skip("scenes/_baked/cornellBox has no G-buffer pass inputs")
    po = ph.oracle()
    scene = fixtures.load_baked("cornellBox")
    w, h = 160, 90
    cam = po.make_camera(position=CAMERAS["cornellBox"][0], look_at=CAMERAS["cornellBox"][1], aspect=w / h)
    sc = ph.oracle_scene(scene)
    a = po.raycast_gbuffer(sc, scene.tri_material, scene.material_table(), cam, w, h)
    b = po.gbuffer_pass(sc, _oracle_inputs(fixtures.load_gbuffer_inputs("cornellBox")), cam, w, h)
    assert np.array_equal(a.world_pos.view(np.uint32), b.world_pos.view(np.uint32)) and np.array_equal(a.depth.view(np.uint32), b.depth.view(np.uint32))
    assert np.abs(a.albedo.astype(int) - b.albedo.astype(int)).max() <= 1 and np.array_equal(a.albedo[..., 3], b.albedo[..., 3])
    assert np.abs(a.material.astype(int) - b.material.astype(int)).max() <= 1
    hit = a.depth < 1.0
    # triangles with a zero tangent (MikkTSpace on cornellBox's degenerate texture coordinates) get a NaN normal, stored as 0
    valid = hit & (b.normal[..., :3] != 0).any(axis=-1)
    assert valid.sum() > 0.5 * hit.sum()
    na, nb = a.normal[valid][:, :3].astype(np.float64) / 32767, b.normal[valid][:, :3].astype(np.float64) / 32767
    assert (np.sum(na * nb, axis=1) > 0.95).mean() > 0.99
    assert (b.normal[~hit][:, :3] == 0).all() and (b.albedo[~hit] == (0, 0, 0, 255)).all()      # the pass's clears


def test_oracle_gbuffer_pass_branches():
    """One quad facing the camera, four ways: textured albedo x colour factor through the sRGB attachment, a normal map tilting
    the normal along the tangent, specular-glossiness conversion (gBuffer.frag:51-67) against its closed form, alpha mask
    discard, emissive override."""
    po = ph.oracle()
    tri = np.array([[[-1, -1, 0, 1], [1, -1, 0, 1], [1, 1, 0, 1]], [[-1, -1, 0, 1], [1, 1, 0, 1], [-1, 1, 0, 1]]], np.float32)
    tris48 = tri.reshape(2, 12).view(np.uint8).reshape(2, 48)
    nodes = capi.build_aabb_tree(tris48)
    blob0 = capi.make_blob(np.zeros((0, 32), np.uint8), 32)
    sc = po.Scene(nodes, tris48, blob0, capi.make_blob(np.zeros((0, 80), np.uint8), 80), capi.make_blob(np.zeros((0, 16), np.uint8), 16))
    verts = np.zeros((4, 20), np.float32)
    verts[:, 0:3] = [(-1, -1, 0), (1, -1, 0), (1, 1, 0), (-1, 1, 0)]
    verts[:, 4:7] = (0, 0, 1)
    verts[:, 8:12] = (1, 0, 0, 1)
    verts[:, 16:18] = [(0, 0), (1, 0), (1, 1), (0, 1)]
    indices = np.array([0, 1, 2, 0, 2, 3], np.uint32)
    eye = np.concatenate([np.eye(4, dtype=np.float32).reshape(-1)] * 2)
    cam = po.make_camera(position=(0.0, 0.0, 3.0), look_at=(0.0, 0.0, 0.0), aspect=1.0)
    w = h = 33

    def run(uniform, binding, textures):
        u = np.zeros(16, np.float32)
        u[0:4] = (1, 1, 1, 1)
        u[15] = 1.0
        for k, v in uniform.items():
            if isinstance(k, int):
                u[k] = v
        ui = u.view(np.int32)
        ui[12], ui[13] = uniform.get("model", 0), uniform.get("alpha", 0)
        gi = po.GBufferInputs(verts.view(np.uint8), indices, np.array([[0, 6, 0, 0]], np.uint32), eye.view(np.uint8), u.view(np.uint8),
                              np.array([binding], np.int32), textures)
        return po.gbuffer_pass(sc, gi, cam, w, h)

    c = (h // 2, w // 2)
    flat = np.full((4, 4, 4), (128, 64, 255, 255), np.uint8)
    g = run({0: 0.5, 5: 0.25, 6: 0.75}, (0, -1, -1, -1), [flat])
    expect = [128 / 255 * 0.5, 64 / 255, 1.0]
    srgb = lambda v: int(round((12.92 * v if v <= 0.0031308 else 1.055 * v ** (1 / 2.4) - 0.055) * 255))
    assert list(g.albedo[c][:3]) == [srgb(v) for v in expect] and g.albedo[c][3] == 0
    assert list(g.material[c]) == [round(0.25 * 65535), round(0.75 * 65535)]          # default white material texture x factors
    assert abs(g.normal[c][2] / 32767 - 1.0) < 1e-3                                  # default normal texel: almost the vertex normal
    tilt = np.full((2, 2, 4), (255, 127, 127, 255), np.uint8)                        # normalTex = (1, ~0, ~0): along the tangent
    g = run({}, (-1, 0, -1, -1), [tilt])
    assert g.normal[c][0] / 32767 > 0.99 and abs(g.normal[c][2] / 32767) < 0.01
    # specular-glossiness: diffuse 0.5, specular 0.3, glossiness 0.8
    g = run({0: 0.5, 1: 0.5, 2: 0.5, 4: 0.3, 5: 0.3, 6: 0.3, 7: 0.8, "model": 1}, (-1, -1, -1, -1), [])
    avg = 0.5 * (0.5 + 0.3)
    root = np.sqrt(avg * avg - 0.04 * 0.5)
    assert g.material[c][0] == round((1 - np.float32(0.8)) * 65535) and abs(g.material[c][1] - min(1.0, 25 * avg - root) * 65535) <= 1
    assert g.albedo[c][0] == srgb(avg + root)
    # alpha mask: the texture's alpha below the cutoff makes a hole, above it not
    holes = np.zeros((2, 2, 4), np.uint8)
    holes[..., :3] = 200
    holes[:, 1, 3] = 255                                                               # right half opaque
    g = run({14: 0.5, "alpha": 1}, (0, -1, -1, -1), [holes])
    row = g.depth[h // 2]                                                              # the quad spans the middle third of the screen
    assert row[w // 2 - 3] == 1.0 and row[w // 2 + 3] < 1.0                            # u = 0.25: discarded down to the clear value; u = 0.75: kept
    # emissive: colour factor x emissive factor x emissive texture, alpha flag set
    g = run({0: 0.5, 8: 2.0, 9: 1.0, 10: 0.0}, (-1, -1, -1, 0), [flat])
    assert g.albedo[c][3] == 255 and list(g.albedo[c][:3]) == [srgb(min(1.0, 0.5 * 2.0 * 128 / 255)), srgb(1.0 * 1.0 * 64 / 255), 0]


# ---- GPU ---------------------------------------------------------------------------------------------------------

def _compare_planes(ctx, slot, want, rows, w, label):
    import torch

    ptrs = ctx.gbuffer_device_planes(slot)
    shapes = [((rows, w, 4), np.uint8), ((rows, w, 4), np.int16), ((rows, w, 2), np.uint16), ((rows, w, 4), np.float32), ((rows, w), np.float32)]
    ctx.synchronize()
    bad = {}
    for ptr, (shape, dt), plane, name in zip(ptrs, shapes, want, ("albedo", "normal", "material", "worldPos", "depth")):
        nbytes = int(np.prod(shape)) * np.dtype(dt).itemsize
        got = torch.as_tensor(fixtures_bytes(ptr, nbytes), device="cuda").cpu().numpy().view(dt).reshape(shape)
        diff = (got.view(np.uint8).reshape(rows * w, -1) != np.ascontiguousarray(plane).view(np.uint8).reshape(rows * w, -1)).any(axis=1)
        if diff.any():
            bad[name] = int(diff.sum())
    assert not bad, f"{label}: planes differ from the oracle twin: {bad}"


def fixtures_bytes(ptr, nbytes):
    bands = __import__("restir_vulkan_b200.bands", fromlist=["bands"])
    return bands._DeviceBytes(ptr, nbytes)


@pytest.mark.gpu
@pytest.mark.parametrize("name,cam_key,size", [("procedural", None, (200, 120)), ("sponza", "sponza", (320, 180)), ("sponza", "sponza_plants", (320, 180)),
                                               ("office", "office", (320, 180)), ("cornellBox", "cornellBox", (256, 144))])
def test_gbuffer_pass_matches_oracle_twin(name, cam_key, size):
    po = ph.oracle()
    if name == "procedural":
        scene, gi = _procedural()
        pos, look = (3.0, 3.5, 4.2), (0.0, -1.0, 0.0)
    else:
        if not fixtures.gbuffer_inputs_available(name):
            pytest.skip(f"scenes/_baked/{name} has no G-buffer pass inputs")
        scene, gi = fixtures.load_baked(name), fixtures.load_gbuffer_inputs(name)
        pos, look = CAMERAS[cam_key]
    w, h = size
    ctx = ph.make_context(scene)
    gi.upload(ctx)
    ctx.resize(w, h)
    oi = _oracle_inputs(gi)
    for k, (p, near, far) in enumerate(((pos, 0.01, 1000.0), ((pos[0] + 0.4, pos[1], pos[2] - 0.3), 0.5, 9.0))):   # the second one clips near and far
        cam = capi.make_camera(position=p, look_at=look, aspect=w / h, z_near=near, z_far=far)
        ocam = po.make_camera(position=p, look_at=look, aspect=w / h, z_near=near, z_far=far)
        ctx.pass_gbuffer(k, cam)
        want = po.gbuffer_pass(ph.oracle_scene(scene), oi, ocam, w, h)
        _compare_planes(ctx, k, want.planes(), h, w, f"{name} camera {k}")
        assert (want.depth < 1.0).mean() > 0.2
    # the pass walks the 64-byte image of the tree and the shadow rays' triangle records; the literal walk of the uploaded 80-byte
    # nodes (trees without an image; restir_set_traversal(REFERENCE_ORDER)) must produce the same planes
    ctx.set_traversal(capi.RESTIR_TRAVERSAL_REFERENCE_ORDER)
    ctx.pass_gbuffer(1, cam)
    _compare_planes(ctx, 1, want.planes(), h, w, f"{name} camera 1, literal walk")
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("unbiased", [True, False])
def test_frames_from_device_made_gbuffers_match_oracle(unbiased):
    """The whole frame without a G-buffer on the host: restir_pass_gbuffer -> restir_frame -> lighting, against the oracle's
    passes run on the oracle twin's G-buffers (three frames, moving camera, temporal reuse on)."""
    import torch

    po = ph.oracle()
    scene, gi = _procedural()
    w, h = 96, 64
    cams = ph.moving_cameras(3, (3.0, 3.5, 4.2), (0.0, -1.0, 0.0), w / h)
    case = ph.Case(scene, w, h, cams, candidates=8, unbiased=unbiased, unbiased_neighbors=3, spatial_iterations=1)
    oi = _oracle_inputs(gi)
    case._gbuffers = [po.gbuffer_pass(ph.oracle_scene(scene), oi, c, w, h) for c in cams]   # the oracle's frames use the twin's planes
    want = ph.run_oracle(case)
    ctx = ph.make_context(scene)
    gi.upload(ctx)
    ctx.resize(w, h)
    ctx.set_unbiased_neighbors(3)
    img = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    got = []
    for f, c in enumerate(cams):
        i = f & 1
        u, lu = case.uniforms(f)
        ctx.set_uniforms(u)
        ctx.set_lighting_uniforms(lu)
        ctx.pass_gbuffer(i, ph.to_capi_camera(c))
        ctx.frame_lit(i, unbiased, 1, img, capi.RESTIR_OUT_RGBA32F)
        ctx.synchronize()
        got.append(dict(reservoirs=ctx.download_reservoirs(i), rgba=img.cpu().numpy().copy()))
    ctx.close()
    for f in range(len(cams)):
        assert ph.compare_reservoirs(got[f]["reservoirs"], want[f]["reservoirs"], f"frame {f}") == 0
        ph.compare_rgb(got[f]["rgba"], want[f]["rgba"])


@pytest.mark.gpu
def test_gbuffer_pass_on_a_band_covers_its_halo_rows():
    scene, gi = _procedural()
    w, h = 128, 96
    cam = capi.make_camera(position=(3.0, 3.5, 4.2), look_at=(0.0, -1.0, 0.0), aspect=w / h)
    po = ph.oracle()
    want = po.gbuffer_pass(ph.oracle_scene(scene), _oracle_inputs(gi), po.make_camera(position=(3.0, 3.5, 4.2), look_at=(0.0, -1.0, 0.0), aspect=w / h), w, h)
    ctx = ph.make_context(scene)
    gi.upload(ctx)
    ctx.resize_band(w, h, 40, 64, 10)
    ctx.pass_gbuffer(0, cam)
    _compare_planes(ctx, 0, [p[30:74] for p in want.planes()], 44, w, "band rows 30..74")
    with pytest.raises(capi.RestirError):
        ctx.gbuffer_device_planes(1)
    ctx.close()


@pytest.mark.gpu
def test_gbuffer_pass_rejects_inconsistent_uploads():
    scene, gi = _procedural()
    ctx = ph.make_context(scene)
    with pytest.raises(capi.RestirError, match="upload"):
        ctx.resize(32, 32)
        ctx.pass_gbuffer(0, capi.make_camera())
    bad = gi.draws.copy()
    bad[0, 1] -= 3                                          # one triangle short of the tree's
    with pytest.raises(capi.RestirError, match="triangles"):
        ctx.upload_geometry(gi.vertices, gi.indices, bad, gi.matrices)
    bad = gi.indices.copy()
    bad[5] = 10 ** 9
    with pytest.raises(capi.RestirError, match="vertex"):
        ctx.upload_geometry(gi.vertices, bad, gi.draws, gi.matrices)
    ctx.close()
