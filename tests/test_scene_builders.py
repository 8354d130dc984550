"""Host builders (AABB tree, triangle lights, random lights, alias table) against output of the REFERENCE's
own C++: committed goldens made by tests/make_golden_scene.py, and — when
scenes/_baked exists — the three reference scenes, byte for byte."""
import ast
import os

import numpy as np
import pytest

import parity_harness as ph

capi, fixtures = ph.capi, ph.fixtures
GOLDEN = os.path.join(ph.ROOT, "tests", "golden")


def _assert_same_bytes(got, want, what):
    got = np.ascontiguousarray(got).view(np.uint8).reshape(-1)
    want = np.ascontiguousarray(want).view(np.uint8).reshape(-1)
    assert got.size == want.size, f"{what}: {got.size} bytes vs {want.size}"
    bad = np.flatnonzero(got != want)
    assert bad.size == 0, f"{what}: {bad.size} bytes differ, first at {bad[:4]}"


@pytest.mark.parametrize("name", ["procedural_point", "procedural_tri", "procedural_random"])
def test_builders_match_reference_on_procedural_goldens(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    kw = ast.literal_eval(str(g["kwargs"]))
    tris9, tri_material, materials, file_lights = fixtures.procedural_scene(**kw)
    tris48 = fixtures.soup_to_triangles48(tris9)
    _assert_same_bytes(tris48, g["triangles"], "triangle order")   # the baker kept the soup's order
    assert np.array_equal(tri_material, g["tri_material"])
    scene = fixtures.assemble_scene(name, tris48, tri_material, materials, file_lights, g["dims"])
    _assert_same_bytes(scene.nodes, g["ref_nodes"], "AABB tree nodes")
    _assert_same_bytes(scene.point_blob, g["ref_point_blob"], "point-light blob")
    _assert_same_bytes(scene.tri_blob, g["ref_tri_blob"], "triangle-light blob")
    _assert_same_bytes(scene.alias_blob, g["ref_alias_blob"], "alias-table blob")


def test_random_point_lights_match_libstdcxx_reference():
    g = np.load(os.path.join(GOLDEN, "random_lights_5000.npz"))
    got = capi.generate_random_point_lights(5000, g["lo"], g["hi"])
    _assert_same_bytes(capi.make_blob(got, 32), g["blob"], "generateRandomPointLights(5000)")
    # first light of the 200-light Sponza fallback
    first = got[0].view(np.float32)
    assert np.allclose(first[:3], [7.12441158, 0.625711799, -9.46231461], rtol=0, atol=1e-4)  # bounds above are the 6-digit roundings
    assert abs(first[7] - 0.460700393) < 1e-6


@pytest.mark.parametrize("name,tris,nodes,point,tri", [("cornellBox", 16732, 16731, 0, 12), ("office", 56520, 56519, 0, 54198),
                                                       ("sponza", 262267, 262266, 200, 0)])
def test_builders_match_reference_on_baked_scenes(name, tris, nodes, point, tri):
    if not fixtures.baked_available(name):
        pytest.skip("scenes/_baked not present (built from /root/reference by `make -C oracle ref`)")
    scene = fixtures.load_baked(name, rebuild=True)
    assert scene.n_triangles == tris and scene.nodes.shape[0] == nodes
    assert scene.light_counts() == (point, tri)
    _assert_same_bytes(scene.nodes, scene.ref["nodes"], f"{name} AABB tree nodes")
    _assert_same_bytes(scene.point_blob, scene.ref["point_blob"], f"{name} point lights")
    _assert_same_bytes(scene.tri_blob, scene.ref["tri_blob"], f"{name} triangle lights")
    _assert_same_bytes(scene.alias_blob, scene.ref["alias_blob"], f"{name} alias table")


def test_alias_table_is_a_distribution():
    scene = fixtures.make_procedural(seed=5, grid=3, boxes=5, lights="tri")
    cols = scene.alias_blob[16:].view(np.dtype([("prob", "f4"), ("alias", "i4"), ("ori", "f4"), ("aori", "f4")]))
    n = cols.shape[0]
    assert n == scene.light_counts()[1] and n > 0
    assert abs(cols["ori"].astype(np.float64).sum() - 1.0) < 1e-4
    assert ((cols["prob"] >= 0) & (cols["prob"] <= 1.0 + 1e-5)).all() and ((cols["alias"] >= 0) & (cols["alias"] < n)).all()
    # the table must reproduce oriProb: P(i) = (prob_i + sum_{j: alias_j = i} (1 - prob_j)) / n
    p = cols["prob"].astype(np.float64).copy()
    np.add.at(p, cols["alias"], 1.0 - cols["prob"].astype(np.float64))
    assert np.allclose(p / n, cols["ori"], atol=1e-5)


def test_tree_is_well_formed():
    scene = fixtures.make_procedural(seed=2, grid=6, boxes=9, lights="random")
    nodes = scene.nodes.view(np.int32).reshape(-1, 20)
    left, right = nodes[:, 16], nodes[:, 17]
    kids = np.concatenate([left, right])
    inner = np.sort(kids[kids >= 0])
    leaves = np.sort(~kids[kids < 0])
    assert np.array_equal(inner, np.arange(1, scene.nodes.shape[0]))     # every node but the root has one parent
    assert np.array_equal(leaves, np.arange(scene.n_triangles))         # every triangle is exactly one leaf
    assert scene.light_counts() == (200, 0)                              # the reference's fallback (sceneBuffers.h:80-82)


def test_builder_rejects_degenerate_input():
    assert capi.load_library().restir_build_aabb_tree(None, 0, None) != 0


@pytest.mark.parametrize("name", ["procedural", "cornellBox", "sponza", "office"])
def test_level_parallel_builder_gives_the_same_bytes(name):
    """restir_build_aabb_tree_mt (one breadth-first level at a time, steps of a level on several threads) against the
    sequential builder — which the tests above hold to the reference's own bytes — for 1, 2, 3 and all threads."""
    import time

    if name == "procedural":
        scene = fixtures.make_procedural(seed=5, grid=30, boxes=200, lights="point")
    else:
        if not fixtures.baked_available(name):
            pytest.skip(f"scenes/_baked/{name} not present")
        scene = fixtures.load_baked(name, rebuild=False)
    tris = scene.triangles
    t0 = time.perf_counter()
    want = capi.build_aabb_tree(tris)
    t1 = time.perf_counter()
    for threads in (1, 2, 3, 0):
        t2 = time.perf_counter()
        got = capi.build_aabb_tree_mt(tris, threads)
        t3 = time.perf_counter()
        assert np.array_equal(got, want), f"{name}: {threads} threads: {(got != want).any(axis=1).sum()} nodes differ"
    print(f"{name}: {tris.shape[0]} triangles, sequential {1e3 * (t1 - t0):.1f} ms, all threads {1e3 * (t3 - t2):.1f} ms")
    # degenerate input is refused the same way
    with pytest.raises(capi.RestirError):
        capi.build_aabb_tree_mt(tris[:1], 0)
