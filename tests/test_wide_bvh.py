"""CPU checks of the 4-wide traversal image (restir-vulkan_b200/csrc/wide_bvh.{h,cpp}).

The product derives, at restir_upload_bvh, a 4-wide re-layout of the uploaded reference tree and claims the
visibility bits are unchanged (wide_bvh.h).  Here the product's builder runs on the host and a scalar
restatement of the GPU's wide traversal (tests/helpers/wide_host_trace.cpp — test infrastructure) is compared
with the oracle's reference-order traversal (softwareRaytracing.glsl:39-85) on the same segments.  The GPU
tests then compare the real kernel with the same oracle.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import parity_harness as ph

fixtures = ph.fixtures
HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "helpers", "libwide_host_trace.so")


def _lib():
    src = [os.path.join(HERE, "helpers", "wide_host_trace.cpp"), os.path.join(ph.ROOT, "restir-vulkan_b200", "csrc", "wide_bvh.cpp")]
    hdr = os.path.join(ph.ROOT, "restir-vulkan_b200", "csrc", "wide_bvh.h")
    if not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in src + [hdr]):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC", "-shared", *src, "-o", SO])
    lib = C.CDLL(SO)
    lib.wide_host_free.restype = None
    lib.wide_host_trace.restype = None
    return lib


def _build(lib, nodes, n_tris):
    nodes = np.ascontiguousarray(nodes).view(np.uint8).reshape(-1, 80)
    handle, info, why = C.c_void_p(), (C.c_uint32 * 6)(), C.create_string_buffer(256)
    rc = lib.wide_host_build(C.c_void_p(nodes.ctypes.data), C.c_uint32(nodes.shape[0]), C.c_uint32(n_tris), C.byref(handle), info, why, 256)
    return rc, handle, list(info), why.value.decode()


def _segments(scene, n, seed):
    rng = np.random.default_rng(seed)
    lo, hi = scene.dims[:3], scene.dims[3:]
    tri = scene.triangles.view(np.float32).reshape(-1, 3, 4)[:, :, :3]
    # from points on triangles (what the passes trace) and from anywhere in the bounds
    pick = rng.integers(0, tri.shape[0], n // 2)
    bc = rng.dirichlet((1, 1, 1), n // 2).astype(np.float32)
    p1a = np.einsum("nk,nkd->nd", bc, tri[pick]).astype(np.float32)
    p1b = rng.uniform(lo, hi, (n - n // 2, 3)).astype(np.float32)
    p1 = np.concatenate([p1a, p1b])
    p2 = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    p2[:500, 0] = p1[:500, 0]            # zero direction components
    p2[500:1000, 1:] = p1[500:1000, 1:]
    p2[1000:1050] = p1[1000:1050]        # zero-length
    return np.ascontiguousarray(p1), np.ascontiguousarray(p2)


def _scenes():
    yield fixtures.make_procedural(seed=5, grid=10, boxes=20, lights="point", n_point_lights=8)
    for name in ("cornellBox", "sponza", "office"):
        if fixtures.baked_available(name):
            yield fixtures.load_baked(name, rebuild=False)


def test_wide_traversal_matches_reference_order():
    lib = _lib()
    po = ph.oracle()
    for scene in _scenes():
        rc, handle, info, why = _build(lib, scene.nodes, scene.n_triangles)
        assert rc == 0, f"{scene.name}: {why}"
        wide_nodes, depth, folded, unfolded, ref_bound, wide_bound = info
        assert unfolded == 0, f"{scene.name}: the reference builder's boxes always nest"
        assert ref_bound <= 32
        assert wide_nodes + folded == scene.nodes.shape[0]   # every binary node is either a wide node's root or folded
        n = 300_000 if scene.n_triangles > 50_000 else 120_000
        p1, p2 = _segments(scene, n, 99)
        want, _, overflow = po.trace_segments(ph.oracle_scene(scene), p1, p2, want_margin=True)
        assert int(overflow.sum()) == 0
        got = np.zeros(n, np.uint8)
        deepest = C.c_int32()
        tris = np.ascontiguousarray(scene.triangles).view(np.float32)
        lib.wide_host_trace(handle, C.c_void_p(tris.ctypes.data), C.c_int64(n), C.c_void_p(p1.ctypes.data), C.c_void_p(p2.ctypes.data),
                            C.c_void_p(got.ctypes.data), C.byref(deepest))
        lib.wide_host_free(handle)
        wide = got != 2                      # rays the product hands to the reference-order traversal
        assert wide.sum() > 0.98 * n
        assert (~wide).sum() >= 1000, "the zero-component segments must take the reference-order path"
        bad = np.flatnonzero(wide & (got != want))
        assert bad.size == 0, f"{scene.name}: {bad.size} of {n} visibility bits differ, first {bad[:5]}"
        assert deepest.value <= wide_bound
        print(f"{scene.name}: {scene.nodes.shape[0]} nodes -> {wide_nodes} wide (depth {depth}), stack bounds ref {ref_bound} / wide {wide_bound}, "
              f"deepest seen {deepest.value}, {int(want.sum())} of {n} shadowed")


def _two_level_tree():
    """4 triangles, 3 nodes; returns (nodes (3,20) f32 view, tris)."""
    scene = fixtures.make_procedural(seed=2, grid=1, boxes=0, lights="point", n_point_lights=1)
    return scene


def test_rejects_out_of_range_children():
    lib = _lib()
    scene = _two_level_tree()
    nodes = scene.nodes.copy()
    n_tris = scene.n_triangles
    ints = nodes.view(np.int32).reshape(-1, 20)
    ints[0, 16] = nodes.shape[0] + 5         # leftChild out of range
    rc, _, _, why = _build(lib, nodes, n_tris)
    assert rc == 1 and "out of range" in why
    nodes = scene.nodes.copy()
    ints = nodes.view(np.int32).reshape(-1, 20)
    ints[0, 17] = ~np.int32(n_tris + 1)      # triangle index out of range
    rc, _, _, why = _build(lib, nodes, n_tris)
    assert rc == 1 and "out of range" in why
    nodes = scene.nodes.copy()
    ints = nodes.view(np.int32).reshape(-1, 20)
    inner = [i for i in range(ints.shape[0]) if ints[i, 16] >= 0]
    ints[inner[-1], 17] = ints[inner[-1], 16]  # same child twice: not a tree
    rc, _, _, why = _build(lib, nodes, n_tris)
    assert rc == 1 and "more than once" in why


def test_non_nested_boxes_are_not_folded():
    """A child box poking out of its parent's box must keep its own test: that node stays a wide node's root."""
    lib = _lib()
    po = ph.oracle()
    scene = fixtures.make_procedural(seed=8, grid=6, boxes=10, lights="point", n_point_lights=4)
    nodes = scene.nodes.copy()
    f = nodes.view(np.float32).reshape(-1, 20)
    ints = nodes.view(np.int32).reshape(-1, 20)
    # shrink the root's left box so that the left child's own boxes stick out of it
    assert ints[0, 16] >= 0
    f[0, 4:7] = f[0, 0:3] + (f[0, 4:7] - f[0, 0:3]) * np.float32(0.25)
    rc, handle, info, why = _build(lib, nodes, scene.n_triangles)
    assert rc == 0, why
    assert info[3] >= 1, "the tampered node must be kept unfolded"
    tampered = fixtures.SceneData(scene.name, scene.triangles, scene.tri_material, scene.materials, nodes, scene.point_blob, scene.tri_blob,
                                  scene.alias_blob, scene.dims)
    p1, p2 = _segments(scene, 60_000, 7)
    want, _, _ = po.trace_segments(ph.oracle_scene(tampered), p1, p2, want_margin=True)
    got = np.zeros(p1.shape[0], np.uint8)
    deepest = C.c_int32()
    tris = np.ascontiguousarray(scene.triangles).view(np.float32)
    lib.wide_host_trace(handle, C.c_void_p(tris.ctypes.data), C.c_int64(p1.shape[0]), C.c_void_p(p1.ctypes.data), C.c_void_p(p2.ctypes.data),
                        C.c_void_p(got.ctypes.data), C.byref(deepest))
    lib.wide_host_free(handle)
    wide = got != 2
    assert np.array_equal(got[wide], want[wide])
    # and the tampering is visible: the untampered tree gives different bits for some rays
    ref, _, _ = po.trace_segments(ph.oracle_scene(scene), p1, p2, want_margin=True)
    assert (ref != want).any()


def test_deep_chain_falls_back_to_reference_order():
    """A 40-deep right-leaning chain can hold more than 32 entries on the reference's stack only if the left
    children are inner nodes; a left-leaning comb does: the builder must refuse the wide path for it."""
    lib = _lib()
    depth = 40
    # comb: node i has left = node i+1 (inner), right = inner node too (a 2-leaf node), so each level leaves one entry pending
    n_nodes = 2 * depth + 1
    nodes = np.zeros((n_nodes, 20), np.float32)
    ints = nodes.view(np.int32)
    box = np.array([0, 0, 0, 0, 1, 1, 1, 0], np.float32)
    tri = 0
    for i in range(depth):
        nodes[i, 0:8] = box
        nodes[i, 8:16] = box
        ints[i, 16] = depth + 1 + i          # left: a 2-leaf node (pushed first, popped last => stays on the stack)
        ints[i, 17] = i + 1                  # right: next comb node (popped first)
    for i in range(depth, n_nodes):
        nodes[i, 0:8] = box
        nodes[i, 8:16] = box
        ints[i, 16] = ~np.int32(tri)
        ints[i, 17] = ~np.int32(tri + 1)
        tri += 2
    rc, _, info, why = _build(lib, nodes.view(np.uint8).reshape(-1, 80), tri)
    assert rc == 2 and "overflow" in why and info[4] > 32
