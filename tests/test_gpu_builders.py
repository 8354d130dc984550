"""AabbTree::build on the device (restir_build_bvh_device) against the host builder and the reference's own
output: byte for byte, including inputs that take the median fallback, signed zeros, duplicates and tiny trees."""
import time

import numpy as np
import pytest

import parity_harness as ph

pytestmark = pytest.mark.gpu
capi, fixtures = ph.capi, ph.fixtures


def _soup(name):
    rng = np.random.default_rng(17)
    if name == "two":
        t = rng.uniform(-1, 1, (2, 9))
    elif name == "three":
        t = rng.uniform(-1, 1, (3, 9))
    elif name == "coincident-centroids":                 # every centroid in one point: 0 / 0 bins, median fallback all the way down
        t = np.stack([np.concatenate([v, -v, v * 0]) for v in rng.uniform(0.1, 1, (500, 3))])             # p1 = v, p2 = -v, p3 = 0: box centre 0
    elif name == "signed-zeros":                         # +0 and -0 both occur as the extreme coordinate
        t = rng.uniform(0.0, 1.0, (4000, 9))
        t[::3, 0] = 0.0
        t[1::3, 0] = -0.0
        t[::5, 4] = -0.0
        t[2::5, 4] = 0.0
    elif name == "duplicates":
        base = rng.uniform(-2, 2, (50, 9))
        t = base[rng.integers(0, 50, 5000)]
    elif name == "grid-aligned":                         # many equal centroids per axis: ties in the binning and empty bins
        c = rng.integers(0, 6, (20000, 1, 3)).astype(np.float64)
        t = (c + rng.integers(0, 2, (20000, 3, 3)) * 0.25).reshape(-1, 9)
    else:
        raise KeyError(name)
    return fixtures.soup_to_triangles48(np.asarray(t, np.float32))


@pytest.mark.parametrize("name", ["two", "three", "coincident-centroids", "signed-zeros", "duplicates", "grid-aligned"])
def test_device_tree_equals_host_tree_on_hard_inputs(name):
    tris = _soup(name)
    want = capi.build_aabb_tree(tris)
    ctx = capi.RestirContext(0)
    got = ctx.build_bvh_device(tris)
    ctx.close()
    diff = np.flatnonzero((got != want).any(axis=1))
    assert diff.size == 0, f"{name}: {diff.size} of {want.shape[0]} nodes differ, first {diff[:5]}"


@pytest.mark.parametrize("name", ["cornellBox", "sponza", "office", "procedural"])
def test_device_tree_equals_reference_tree(name):
    """The scenes' trees: the device build against the reference's own AabbTree::build output (scenes/_baked/*/ref_nodes.bin,
    dumped by oracle/_ref/scene_baker) and against what restir_check_aabb_tree reports for it; then shadow rays through the
    installed tree against the oracle."""
    import torch

    if name == "procedural":
        scene = fixtures.make_procedural(seed=9, grid=40, boxes=300, lights="random")
        ref_nodes = scene.nodes
    else:
        if not fixtures.baked_available(name):
            pytest.skip(f"scenes/_baked/{name} not present")
        scene = fixtures.load_baked(name, rebuild=False)          # .nodes = the reference's
        ref_nodes = scene.nodes
    ctx = capi.RestirContext(0)
    ctx.build_bvh_device(scene.triangles)                             # warm (module load, allocations)
    torch.cuda.synchronize()
    ctx.profile_begin()
    t0 = time.perf_counter()
    got = ctx.build_bvh_device(scene.triangles, want_nodes=False)
    ctx.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    prof = ctx.profile_end()
    dev_ms, wide_ms = prof["bvh_build (all kernels)"][1], prof.get("wide_build (all kernels)", (0, 0.0))[1]
    got = ctx.build_bvh_device(scene.triangles)
    assert np.array_equal(got, ref_nodes), f"{name}: {(got != ref_nodes).any(axis=1).sum()} nodes differ from the reference's tree"
    rc, info, _ = capi.check_aabb_tree(ref_nodes, scene.n_triangles)
    mine = ctx.bvh_info()
    assert rc == 0 and mine["depth"] == info["depth"] and mine["nodes"] == info["nodes"] and mine["traversal"] == info["traversal"] == capi.RESTIR_TRAVERSAL_WIDE
    # the 4-wide image derived on the device (restir_wide_build.cu) is the one the host derives from the same nodes (wide_image.cpp)
    assert (mine["wide_nodes"], mine["wide_depth"]) == (info["wide_nodes"], info["wide_depth"])
    dev_nodes, dev_order = ctx.wide_image()
    up = capi.RestirContext(0)
    up.upload_bvh(ref_nodes, scene.triangles)
    host_nodes, host_order = up.wide_image()
    up.close()
    assert np.array_equal(dev_nodes, host_nodes), f"{name}: {(dev_nodes != host_nodes).any(axis=1).sum()} wide nodes differ"
    assert np.array_equal(dev_order, host_order)
    assert mine["reference_stack_bound"] >= info["reference_stack_bound"]
    print(f"{name}: {scene.n_triangles} triangles, device build {dev_ms:.2f} ms on the stream (CUDA events around all of its kernels and per-level "
          f"read-backs) + {wide_ms:.2f} ms for its 4-wide image, {ms:.2f} ms wall with upload and install, depth {mine['depth']}, {mine['wide_nodes']} wide nodes in "
          f"{mine['wide_depth']} levels")
    # the installed tree traces like the uploaded one
    po = ph.oracle()
    rng = np.random.default_rng(3)
    lo, hi = scene.dims[:3], scene.dims[3:]
    p1 = rng.uniform(lo, hi, (200_000, 3)).astype(np.float32)
    p2 = rng.uniform(lo, hi, (200_000, 3)).astype(np.float32)
    want = po.trace_segments(ph.oracle_scene(scene), p1, p2)
    out = torch.zeros(p1.shape[0], dtype=torch.uint8, device="cuda")
    ctx.trace_segments(torch.from_numpy(p1).cuda(), torch.from_numpy(p2).cuda(), p1.shape[0], out)
    ctx.synchronize()
    assert np.array_equal(out.cpu().numpy(), np.asarray(want, np.uint8))
    ctx.close()


def test_device_build_rejects_non_finite_coordinates():
    tris = _soup("duplicates").copy()
    tris.view(np.float32).reshape(-1, 12)[7, 1] = np.nan
    ctx = capi.RestirContext(0)
    with pytest.raises(capi.RestirError, match="non-finite"):
        ctx.build_bvh_device(tris)
    ctx.close()
