"""The 4-wide quantised image of the uploaded tree (restir-vulkan_b200/csrc/wide_image.{h,cpp}, walked by
restir_wide.cuh) on the CPU: restir_check_wide_walk runs the device walk's own operations on the host (the box
arithmetic is one header shared by host and device) and must give the oracle's visibility bits — the reference's
softwareRaytracing.glsl:39-85 on the reference's tree — for every segment it does not hand to the binary image.
No GPU needed; the GPU twin is tests/test_gpu_parity.py::test_wide_walk_*.
"""
import numpy as np
import pytest

import parity_harness as ph

capi, fixtures = ph.capi, ph.fixtures


def _scene(name):
    if name.startswith("procedural"):
        return fixtures.make_procedural(seed=11, grid=12, boxes=40, lights="point", n_point_lights=8)
    if not fixtures.baked_available(name):
        pytest.skip(f"scenes/_baked/{name} not present")
    return fixtures.load_baked(name, rebuild=False)


def adversarial_segments(scene, n, seed):
    """Segments that sit on the edges of the slab arithmetic: axis-parallel directions (a zero component: 1/0 = inf, the
    wide walk must refuse them), end points exactly on box planes and triangle vertices, tiny and huge directions, origins far
    outside the scene, and plain random ones."""
    rng = np.random.default_rng(seed)
    lo, hi = scene.dims[:3].astype(np.float64), scene.dims[3:].astype(np.float64)
    p1 = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    p2 = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    nodes = np.ascontiguousarray(scene.nodes).view(np.float32).reshape(-1, 20)
    tris = np.ascontiguousarray(scene.triangles).view(np.float32).reshape(-1, 12)
    k = n // 8
    # 1: axis-parallel (one or two components of the direction exactly zero)
    ax = rng.integers(0, 3, k)
    p2[np.arange(k), ax] = p1[np.arange(k), ax]
    ax2 = rng.integers(0, 3, k // 2)
    p2[np.arange(k // 2), ax2] = p1[np.arange(k // 2), ax2]
    # 2: origins exactly on box planes of random nodes
    pick = rng.integers(0, nodes.shape[0], k)
    a = rng.integers(0, 3, k)
    p1[k:2 * k][np.arange(k), a] = nodes[pick, a]                     # leftAabbMin plane
    # 3: end points on triangle vertices, start points on others (rays along edges and through vertices)
    t1, t2 = rng.integers(0, tris.shape[0], k), rng.integers(0, tris.shape[0], k)
    p1[2 * k:3 * k] = tris[t1, 0:3]
    p2[2 * k:3 * k] = tris[t2, 4:7]
    # 4: very short segments (direction ~ 1e-3: the tMin offsets dominate) and coincident points (0 / 0)
    p2[3 * k:4 * k] = p1[3 * k:4 * k] + rng.uniform(-2e-3, 2e-3, (k, 3)).astype(np.float32)
    p2[3 * k:3 * k + 8] = p1[3 * k:3 * k + 8]
    # 5: origins far outside the scene, some absurdly far, some non-finite
    far = (hi - lo) * rng.uniform(2, 1000, (k, 1))
    p1[4 * k:5 * k] = (p1[4 * k:5 * k].astype(np.float64) + far * rng.choice([-1, 1], (k, 3))).astype(np.float32)
    p1[4 * k:4 * k + 4] *= np.float32(1e12)
    p1[4 * k + 4, 0] = np.inf
    p1[4 * k + 5, 1] = np.nan
    p2[4 * k + 6, 2] = -np.inf
    # 6: nearly axis-parallel (tiny but non-zero components: huge reciprocal directions)
    tiny = rng.uniform(-1, 1, k).astype(np.float32) * np.float32(10.0) ** rng.integers(-30, -3, k).astype(np.float32)
    a = rng.integers(0, 3, k)
    p2[5 * k:6 * k][np.arange(k), a] = p1[5 * k:6 * k][np.arange(k), a] + tiny
    return p1, p2


@pytest.mark.parametrize("name", ["procedural", "cornellBox", "sponza", "office"])
def test_wide_walk_host_emulation_matches_the_oracle(name):
    scene = _scene(name)
    n = 160_000 if name != "procedural" else 240_000
    p1, p2 = adversarial_segments(scene, n, seed=3)
    rc, shadowed, walked, visits, msg = capi.check_wide_walk(scene.nodes, scene.triangles, p1, p2)
    assert rc == 0, msg
    with np.errstate(all="ignore"):
        want = ph.oracle().trace_segments(ph.oracle_scene(scene), p1, p2)
    took = walked != 0
    assert took.sum() > 0.7 * n                                   # most segments are the wide walk's
    assert np.array_equal(shadowed[took], want[took]), f"{int((shadowed[took] != want[took]).sum())} of {int(took.sum())} bits differ"
    # what it refuses: a zero / non-finite / out-of-range component somewhere (never a plain ray)
    with np.errstate(all="ignore"):
        d = p2.astype(np.float64) - p1.astype(np.float64)
    plain = np.isfinite(d).all(axis=1) & (np.abs(d) > 1e-2).all(axis=1) & (np.abs(p1) < 1e4).all(axis=1)
    assert took[plain].all()
    assert 0 < shadowed[took].mean() < 1 and visits > 0


def test_wide_image_is_reported_and_refused_where_its_argument_does_not_hold():
    scene = _scene("procedural")
    rc, info, msg = capi.check_aabb_tree(scene.nodes, scene.n_triangles)
    assert rc == 0 and info["traversal"] == capi.RESTIR_TRAVERSAL_WIDE and msg == ""
    assert 0 < info["wide_nodes"] < scene.nodes.shape[0] and 0 < info["wide_depth"] <= info["depth"] and info["wide_stack_bound"] <= 32
    f = lambda nodes: np.ascontiguousarray(nodes).view(np.float32).reshape(-1, 20)
    # a child box that sticks out of the box its parent stores for it: not nested => the binary image is walked
    nodes = scene.nodes.copy()
    ints = nodes.view(np.int32).reshape(-1, 20)
    inner = next(i for i in range(ints.shape[0]) if ints[i, 16] >= 0)
    f(nodes)[ints[inner, 16], 4] += 1000.0                       # leftAabbMax.x of the left child node
    rc, info, msg = capi.check_aabb_tree(nodes, scene.n_triangles)
    assert rc == 0 and info["traversal"] == capi.RESTIR_TRAVERSAL_IMAGE and "nested" in msg
    # a non-finite plane
    nodes = scene.nodes.copy()
    f(nodes)[0, 0] = -np.inf
    rc, info, msg = capi.check_aabb_tree(nodes, scene.n_triangles)
    assert rc == 0 and info["traversal"] == capi.RESTIR_TRAVERSAL_IMAGE and "finite" in msg
    # one triangle under two leaves
    nodes = scene.nodes.copy()
    ints = nodes.view(np.int32).reshape(-1, 20)
    leaves = [(i, s) for i in range(ints.shape[0]) for s in (16, 17) if ints[i, s] < 0]
    ints[leaves[1][0], leaves[1][1]] = ints[leaves[0][0], leaves[0][1]]
    rc, info, msg = capi.check_aabb_tree(nodes, scene.n_triangles)
    assert rc == 0 and info["traversal"] == capi.RESTIR_TRAVERSAL_IMAGE and "more than one leaf" in msg
    p = np.zeros((1, 3), np.float32)
    rc, *_, msg = capi.check_wide_walk(nodes, scene.triangles, p, p + 1)
    assert rc != 0 and "more than one leaf" in msg


def test_wide_margin_never_loses_a_box_the_reference_slab_test_passes():
    """The theorem behind the walk, sampled: whenever the reference's slab test passes for a box, the wide test passes for
    every grid box that encloses it.  Exercised through whole walks above; here directly on boxes that just touch the segment,
    where a missing ulp would show: a thin slab of triangles at grazing incidence, 200 000 rays that skim it."""
    rng = np.random.default_rng(9)
    scene = _scene("procedural")
    tris = np.ascontiguousarray(scene.triangles).view(np.float32).reshape(-1, 12)
    n = 200_000
    t = rng.integers(0, tris.shape[0], n)
    w = rng.dirichlet((1, 1, 1), n).astype(np.float32)
    on = w[:, :1] * tris[t, 0:3] + w[:, 1:2] * tris[t, 4:7] + w[:, 2:3] * tris[t, 8:11]     # a point on a triangle
    mix = rng.uniform(0.1, 0.9, (n, 1)).astype(np.float32)                               # a direction in the triangle's plane
    e = (tris[t, 4:7] - tris[t, 0:3]) * mix + (tris[t, 8:11] - tris[t, 0:3]) * (1 - mix)
    e /= np.maximum(np.linalg.norm(e, axis=1, keepdims=True), 1e-20)
    lift = rng.uniform(-1e-5, 1e-5, (n, 1)).astype(np.float32)
    nrm = np.cross(tris[t, 4:7] - tris[t, 0:3], tris[t, 8:11] - tris[t, 0:3])
    nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-20)
    length = rng.uniform(0.05, 3.0, (n, 1)).astype(np.float32)
    p1 = (on - e * length + nrm * lift).astype(np.float32)                                # skims the triangle's plane
    p2 = (on + e * length - nrm * lift).astype(np.float32)
    rc, shadowed, walked, _, msg = capi.check_wide_walk(scene.nodes, scene.triangles, p1, p2)
    assert rc == 0, msg
    want = ph.oracle().trace_segments(ph.oracle_scene(scene), p1, p2)
    took = walked != 0
    assert took.mean() > 0.5 and np.array_equal(shadowed[took], want[took])


def test_fast_div_formula_is_exact():
    """csrc/restir_kernels.h FastDiv: n / d as the high half of n * ceil(2^64 / d) — the trace kernels divide item and pixel
    numbers by the screen width, the tiles per row and the rays per pixel this way.  The formula, restated with Python's integers,
    against n // d on the edges and on random values of the whole 32-bit range."""
    rng = np.random.default_rng(17)
    ds = [2, 3, 5, 7, 240, 1920, 1921, 3840, 7680, 65535, 65536, 65537, (1 << 20), (1 << 31) - 1, (1 << 31), (1 << 32) - 1] + \
        [int(x) for x in rng.integers(2, 1 << 32, 200)]
    for d in ds:
        m = ((1 << 64) - 1) // d + 1
        assert m < (1 << 64)
        ns = [0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, (1 << 32) - 1, (1 << 32) - d, ((1 << 32) - 1) // d * d, ((1 << 32) - 1) // d * d - 1] + \
            [int(x) for x in rng.integers(0, 1 << 32, 400)]
        for n in ns:
            if 0 <= n < (1 << 32):
                assert (n * m) >> 64 == n // d, (n, d)


def test_saturated_strict_box_test_equals_the_clamped_one():
    """restir_wide.cuh gets the clamp of the box test to the segment's range from saturating the x axis' multiply-adds, and reads
    "hit" off the sign of entry - exit (strict).  The claim in wide_image.h (wide_box_hit): with n' = max(sat(entry_x), entry_y,
    entry_z) and f' = min(sat(exit_x), exit_y, exit_z),   n' < f'  <=>  max(entry, 0) < min(exit, 1)   for ANY six parameters — the
    saturated test visits exactly the boxes the clamped one does (boxes that only the clamp separates from the segment collapse onto
    1 - 1 or 0 - 0 and are missed by both).  Checked on random parameters concentrated around the two clamp values, with ties."""
    rng = np.random.default_rng(11)
    n = 400_000
    pool = np.array([-1e30, -2.0, -1.0, -1e-7, -0.0, 0.0, 1e-7, 0.25, 0.5, 0.75, 1.0 - 6e-8, 1.0, 1.0 + 1.2e-7, 2.0, 1e30], dtype=np.float32)
    t = np.where(rng.random((n, 6)) < 0.5, pool[rng.integers(0, len(pool), (n, 6))], rng.normal(0.5, 1.0, (n, 6)).astype(np.float32)).astype(np.float32)
    nx, ny, nz, fx, fy, fz = (t[:, i] for i in range(6))
    sat = lambda v: np.minimum(np.maximum(v, np.float32(0.0)), np.float32(1.0))
    near_sat = np.maximum(np.maximum(sat(nx), ny), nz)
    far_sat = np.minimum(np.minimum(sat(fx), fy), fz)
    near = np.maximum(np.maximum(np.maximum(nx, ny), nz), np.float32(0.0))
    far = np.minimum(np.minimum(np.minimum(fx, fy), fz), np.float32(1.0))
    hit_sat = (near_sat - far_sat) < 0          # the kernel: sign bit of entry - exit
    hit_clamped = near < far
    assert np.array_equal(hit_sat, hit_clamped)
    assert 0.02 < hit_sat.mean() < 0.98          # both outcomes are exercised
