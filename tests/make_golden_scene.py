"""Generates tests/golden/procedural_*.npz: small procedural triangle soups pushed through the REFERENCE's
own AabbTree::build / light collection / createAliasTable (oracle/_ref/scene_baker, built from
/root/reference by oracle/ref_build/Makefile).  Run here (where /root/reference exists); the fixtures are
committed so the builder tests also run where the reference is absent.

    python tests/make_golden_scene.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import parity_harness as ph  # noqa: E402

BAKER = os.path.join(ph.ROOT, "oracle", "_ref", "scene_baker")
GOLDEN = os.path.join(ph.ROOT, "tests", "golden")

CASES = {
    "procedural_point": dict(seed=11, grid=5, boxes=10, lights="point", n_point_lights=7),
    "procedural_tri": dict(seed=12, grid=4, boxes=8, lights="tri"),
    "procedural_random": dict(seed=13, grid=3, boxes=6, lights="random"),
}


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    for name, kw in CASES.items():
        tris9, tri_material, materials, file_lights = ph.fixtures.procedural_scene(**kw)
        with tempfile.TemporaryDirectory() as d:
            soup = os.path.join(d, "in.soup")
            ph.fixtures.write_soup(soup, tris9, tri_material, materials, file_lights)
            subprocess.check_call([BAKER, "soup", soup, d], stdout=subprocess.DEVNULL)
            f = lambda n, dt: np.fromfile(os.path.join(d, n), dtype=dt)
            np.savez_compressed(
                os.path.join(GOLDEN, name + ".npz"),
                kwargs=np.array(repr(kw)),
                triangles=f("triangles.bin", np.uint8), tri_material=f("tri_material.i32", np.int32),
                dims=f("dims.f32", np.float32), ref_nodes=f("ref_nodes.bin", np.uint8),
                ref_point_blob=f("ref_point_lights.bin", np.uint8), ref_tri_blob=f("ref_tri_lights.bin", np.uint8),
                ref_alias_blob=f("ref_alias.bin", np.uint8))
        print("wrote", name)
    # the fallback light generator alone, 5000 lights in Sponza's bounds
    lo, hi = (-15.3676, -1.01154, -9.46246), (14.3993, 11.4355, 8.84341)
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "rl.bin")
        subprocess.check_call([BAKER, "randlights", "5000"] + [repr(float(np.float32(v))) for v in lo + hi] + [out])
        np.savez_compressed(os.path.join(GOLDEN, "random_lights_5000.npz"), lo=np.float32(lo), hi=np.float32(hi),
                            blob=np.fromfile(out, np.uint8))
    print("wrote random_lights_5000")


if __name__ == "__main__":
    main()
