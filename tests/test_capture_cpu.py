"""Frame captures (include/restir_capture.h): the Python writer / reader round-trips every byte, the C++
reader (host/capture.hpp) sees the same sections, and a capture replayed through the oracle reproduces the outputs it
carries — so a dump of the real Vulkan application can be checked the same way."""
import json
import os
import subprocess

import numpy as np
import pytest

import parity_harness as ph

fixtures = ph.fixtures
capture = __import__("restir_vulkan_b200.capture", fromlist=["capture"])


def _case(unbiased):
    scene = fixtures.make_procedural(seed=11, grid=8, boxes=10, lights="tri" if unbiased else "point", n_point_lights=9)
    w, h = 40, 24
    cams = ph.moving_cameras(3, (3.0, 3.5, 4.2), (0.0, -1.0, 0.0), w / h)
    return ph.Case(scene, w, h, cams, candidates=8, unbiased=unbiased, unbiased_neighbors=5 if unbiased else 3, spatial_iterations=1)


def _fnv(chunks):
    h = 1469598103934665603
    for c in chunks:
        for x in np.frombuffer(np.ascontiguousarray(c).tobytes(), np.uint8).tolist():
            h = ((h ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"


@pytest.mark.parametrize("unbiased", [True, False])
def test_capture_round_trip_and_oracle_replay(tmp_path, unbiased):
    case = _case(unbiased)
    want = ph.run_oracle(case)
    path = str(tmp_path / "frames.rsc")
    ph.make_capture(case, want).write(path)
    assert os.path.getsize(path) > 72
    cap = capture.Capture.read(path)
    assert (cap.width, cap.height, len(cap.frames), cap.unbiased, cap.unbiased_neighbors) == (case.w, case.h, 3, unbiased, case.unbiased_neighbors)
    assert cap.expected_bits() == 7
    assert np.array_equal(cap.nodes, case.scene.nodes.reshape(-1, 80)) and np.array_equal(cap.alias_blob, case.scene.alias_blob)
    # replay the read-back capture through the oracle: only what the file holds is used
    po = ph.oracle()
    sc = po.Scene(cap.nodes, cap.triangles, cap.point_blob, cap.tri_blob, cap.alias_blob)
    n = cap.width * cap.height
    bufs = [np.zeros(n, po.RESERVOIR_DTYPE), np.zeros(n, po.RESERVOIR_DTYPE)]
    prev = None
    for f, fr in enumerate(cap.frames):
        i, p = f & 1, (f & 1) ^ 1
        g = po.GBuffer(cap.width, cap.height, *[np.array(x) for x in fr["planes"]])
        u = np.array(fr["uniforms"]).astype(po.UNIFORMS_DTYPE)
        lu = np.array(fr["lighting_uniforms"]).astype(po.LIGHTING_UNIFORMS_DTYPE)
        initial, _ = po.restir_pass(sc, u, g, prev, bufs[p], None)
        assert ph.compare_reservoirs(initial, fr["initial"], f"replay frame {f} initial") == 0
        if cap.unbiased:
            bufs[i], _ = po.unbiased_pass(sc, u, g, initial, cap.unbiased_neighbors, None)
        else:
            bufs[i] = initial
            for j in range(cap.spatial_iterations):
                bufs[p] = po.spatial_pass(u, g, bufs[i], 2 * j, None)
                bufs[i] = po.spatial_pass(u, g, bufs[p], 2 * j + 1, None)
        assert ph.compare_reservoirs(bufs[i], fr["final"], f"replay frame {f} final") == 0
        rgba = po.lighting_pass(sc, lu, g, bufs[i], None)
        assert ph.bits_equal(rgba[..., :3], fr["rgba"][..., :3]).all()
        prev = g


def test_cpp_reader_sees_what_python_wrote(tmp_path):
    case = _case(True)
    want = ph.run_oracle(case)
    path = str(tmp_path / "frames.rsc")
    cap = ph.make_capture(case, want)
    cap.write(path)
    exe = str(tmp_path / "capture_info")
    host = os.path.join(ph.ROOT, "restir-vulkan_b200", "host")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", os.path.join(host, "capture_info.cpp"), "-o", exe], check=True)
    info = json.loads(subprocess.run([exe, path], check=True, capture_output=True, text=True).stdout)
    assert (info["width"], info["height"], info["frames"], info["unbiased"], info["unbiased_neighbors"], info["expected"]) == (40, 24, 3, 1, 5, 7)
    assert info["n_nodes"] == cap.nodes.shape[0] and info["n_triangles"] == cap.triangles.shape[0]
    assert info["scene_fnv1a"] == _fnv([cap.nodes, cap.triangles, cap.point_blob, cap.tri_blob, cap.alias_blob])
    chunks, exp = [], []
    for fr in capture.Capture.read(path).frames:
        chunks += [np.array(fr["uniforms"]), np.array(fr["lighting_uniforms"])] + [np.array(p) for p in fr["planes"]]
        exp += [fr["initial"], fr["final"], fr["rgba"]]
    assert info["inputs_fnv1a"] == _fnv(chunks)
    assert info["expected_fnv1a"] == _fnv(exp)
    # a truncated file is refused, by both readers
    data = open(path, "rb").read()
    open(path, "wb").write(data[:-100])
    assert subprocess.run([exe, path], capture_output=True).returncode == 1
    with pytest.raises(ValueError):
        capture.Capture.read(path)
