"""Generates tests/golden/frames_*.npz: whole frame sequences computed by the REFERENCE's own shader sources
(oracle/_ref/libglslref.so = restirOmni.glsl, spatialReuse.comp, unbiasedReuse.glsl, lighting.frag compiled from
/root/reference by oracle/ref_build/Makefile).  Run here, where /root/reference exists; the fixtures are committed
so that the oracle (CPU suite) and the CUDA path (GPU suite) are checked against reference-made outputs even where
the reference and oracle/_ref are absent.

    python tests/make_golden_frames.py

Each fixture stores the case's parameters, and per frame: the reservoirs after restirOmni (`initial`), the final
reservoirs and the linear RGBA output.  Inputs (procedural scene, cameras, G-buffers) are regenerated from the
parameters by the same deterministic code the tests use; a checksum of the G-buffer guards that.
"""
import os
import sys
import zlib

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import parity_harness as ph  # noqa: E402

GOLDEN = os.path.join(ph.ROOT, "tests", "golden")

CASES = {
    "frames_point_biased": dict(scene=dict(seed=31, grid=6, boxes=10, lights="point", n_point_lights=9), size=(64, 40), frames=3,
                                case=dict(unbiased=False, spatial_iterations=1, neighbors=4)),
    "frames_point_unbiased3": dict(scene=dict(seed=32, grid=6, boxes=10, lights="point", n_point_lights=9), size=(64, 40), frames=3,
                                   case=dict(unbiased=True, unbiased_neighbors=3)),
    "frames_tri_unbiased5": dict(scene=dict(seed=33, grid=5, boxes=8, lights="tri"), size=(64, 40), frames=3,
                                 case=dict(unbiased=True, unbiased_neighbors=5)),
    "frames_tri_biased2": dict(scene=dict(seed=34, grid=5, boxes=8, lights="tri"), size=(61, 37), frames=2,
                               case=dict(unbiased=False, spatial_iterations=2, neighbors=5, candidates=16)),
}
CAMERA = ((3.0, 3.5, 4.2), (0.0, -1.0, 0.0))


def build_case(spec):
    scene = ph.fixtures.make_procedural(**spec["scene"])
    w, h = spec["size"]
    cams = ph.moving_cameras(spec["frames"], CAMERA[0], CAMERA[1], w / h)
    return ph.Case(scene, w, h, cams, **spec["case"])


def gbuffer_checksum(case):
    c = 0
    for g in case.gbuffers():
        for p in g.planes():
            c = zlib.crc32(np.ascontiguousarray(p).view(np.uint8).reshape(-1).tobytes(), c)
    return c


def main():
    gl = ph.glsl_reference()
    if gl is None:
        raise SystemExit("oracle/_ref/libglslref.so missing: run `make -C oracle` where /root/reference exists")
    os.makedirs(GOLDEN, exist_ok=True)
    for name, spec in CASES.items():
        case = build_case(spec)
        frames = ph.run_oracle(case, passes=gl)
        arrays = {"spec": np.array(repr(spec)), "gbuffer_crc32": np.array(gbuffer_checksum(case), np.uint32)}
        for f, fr in enumerate(frames):
            arrays[f"initial_{f}"] = fr["initial"]
            arrays[f"final_{f}"] = fr["reservoirs"]
            arrays[f"rgba_{f}"] = fr["rgba"]
        path = os.path.join(GOLDEN, name + ".npz")
        np.savez_compressed(path, **arrays)
        lit = (frames[-1]["reservoirs"]["w"] > 0).mean()
        print(f"{name}: {len(frames)} frames of {case.w}x{case.h}, w>0 on {lit:.1%}, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
