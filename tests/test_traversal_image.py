"""Host-side checks restir_upload_bvh runs on an uploaded AABB tree (restir-vulkan_b200/csrc/traversal_image.cpp),
through restir_check_aabb_tree — no GPU needed.  Known answers: SURVEY.md Appendix D (measured from the
reference's own builder): worst-case stack 15 / 22 / 17 and deepest leaf 20 / 30 / 26 for cornellBox / Sponza / office.
"""
import numpy as np
import pytest

import parity_harness as ph

capi, fixtures = ph.capi, ph.fixtures


@pytest.mark.parametrize("name,stack,depth", [("cornellBox", 15, 20), ("sponza", 22, 30), ("office", 17, 26)])
def test_shipped_scenes_known_answers(name, stack, depth):
    if not fixtures.baked_available(name):
        pytest.skip(f"scenes/_baked/{name} not present")
    scene = fixtures.load_baked(name, rebuild=False)          # the reference builder's own nodes
    rc, info, msg = capi.check_aabb_tree(scene.nodes, scene.n_triangles)
    assert rc == 0 and msg == ""
    assert info["traversal"] == capi.RESTIR_TRAVERSAL_WIDE      # the reference's trees are nested and finite: wide_image.h applies
    assert info["wide_depth"] <= (depth + 1) // 2 + 4 and info["wide_stack_bound"] <= 32 and info["wide_nodes"] < scene.nodes.shape[0]
    assert info["reachable_nodes"] == scene.nodes.shape[0] == scene.n_triangles - 1
    assert info["reference_stack_bound"] == stack
    assert info["depth"] == depth


def _small():
    return fixtures.make_procedural(seed=2, grid=2, boxes=2, lights="point", n_point_lights=1)


def test_rejects_what_the_reference_would_read_out_of_bounds():
    scene = _small()
    n_tris = scene.n_triangles
    nodes = scene.nodes.copy()
    ints = nodes.view(np.int32).reshape(-1, 20)
    ints[0, 16] = nodes.shape[0] + 5                          # leftChild past the node array
    rc, _, msg = capi.check_aabb_tree(nodes, n_tris)
    assert rc != 0 and "out of range" in msg
    nodes = scene.nodes.copy()
    ints = nodes.view(np.int32).reshape(-1, 20)
    ints[0, 17] = ~np.int32(n_tris + 1)                       # triangle index past the triangle array
    rc, _, msg = capi.check_aabb_tree(nodes, n_tris)
    assert rc != 0 and "out of range" in msg
    nodes = scene.nodes.copy()
    ints = nodes.view(np.int32).reshape(-1, 20)
    inner = [i for i in range(ints.shape[0]) if ints[i, 16] >= 0 and ints[i, 17] >= 0]
    ints[inner[0], 17] = ints[inner[0], 16]                   # the same node twice: not a tree
    rc, _, msg = capi.check_aabb_tree(nodes, n_tris)
    assert rc != 0 and "more than once" in msg


def test_tree_that_can_overflow_the_reference_stack_keeps_the_literal_walk():
    """A comb whose every level leaves one entry pending holds `depth` entries at its deepest point."""
    depth = 40
    n_nodes = 2 * depth + 1
    nodes = np.zeros((n_nodes, 20), np.float32)
    ints = nodes.view(np.int32)
    box = np.array([0, 0, 0, 0, 1, 1, 1, 0], np.float32)
    tri = 0
    for i in range(depth):
        nodes[i, 0:8] = box
        nodes[i, 8:16] = box
        ints[i, 16] = depth + 1 + i      # left: a two-leaf node, pushed first => stays on the stack
        ints[i, 17] = i + 1              # right: the next comb node, popped first
    for i in range(depth, n_nodes):
        nodes[i, 0:8] = box
        nodes[i, 8:16] = box
        ints[i, 16] = ~np.int32(tri)
        ints[i, 17] = ~np.int32(tri + 1)
        tri += 2
    rc, info, msg = capi.check_aabb_tree(nodes.view(np.uint8).reshape(-1, 80), tri)
    assert rc == 0
    assert info["traversal"] == capi.RESTIR_TRAVERSAL_REFERENCE_ORDER and info["reference_stack_bound"] > 32 and "overflow" in msg
