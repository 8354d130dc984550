"""Pins the CPU oracle against the reference's OWN shader sources.

oracle/_ref/libglslref.so is the reference's restirOmni.glsl, spatialReuse.comp, unbiasedReuse.glsl, lighting.frag
and their includes, compiled by g++ from where they lie under /root/reference after a token-level transliteration
(oracle/ref_build/glsl2cpp.py: float-literal suffixes, `inout` -> reference, interface blocks -> globals — the
authors' arithmetic, order of operations and control flow are untouched) with the GLSL built-ins of
oracle/ref_build/glsl_shim.h (DESIGN.md's arithmetic policy P1-P12).  These tests demand that the oracle's hand
restatement reproduces that build BIT FOR BIT: every reservoir field of every pixel of every frame, every
visibility bit, every output colour.  The CUDA kernels are in turn held bit-identical to the oracle by the GPU
parity tests, which closes the chain  reference source == oracle == kernels.

The library is built wherever /root/reference exists (this container; `make -C oracle`) and travels to the GPU
box as a prebuilt file; without it these tests skip, and say so.
"""
import numpy as np
import pytest

import parity_harness as ph

fixtures = ph.fixtures
po = ph.oracle()
gl = ph.glsl_reference()

pytestmark = pytest.mark.skipif(gl is None, reason="oracle/_ref/libglslref.so not built (needs /root/reference)")

CAMERAS = {
    "procedural": ((3.0, 3.5, 4.2), (0.0, -1.0, 0.0)),
    "cornellBox": ((3.0, 4.0, 5.0), (0.0, 0.0, 0.0)),
    "sponza": ((3.0, 4.0, 5.0), (0.0, 0.0, 0.0)),
    "office": ((3.0, 1.7, 0.5), (3.0, 1.5, -5.0)),
}


def _scene(name):
    if name.startswith("procedural"):
        kind = name.split(":")[1]
        return fixtures.make_procedural(seed=21, grid=10, boxes=20, lights=kind, n_point_lights=24)
    if not fixtures.baked_available(name):
        pytest.skip(f"scenes/_baked/{name} not present")
    return fixtures.load_baked(name, rebuild=True)


def _cams(name, n, w, h):
    pos, look = CAMERAS[name.split(":")[0]]
    return ph.moving_cameras(n, pos, look, w / h)


# ---- the small pieces --------------------------------------------------------------------------------------------

def test_pcg32_and_rand_float():
    for seed, seq in [(42, 54), (1, 0), (1, 3 * 10007 + 5), (2 ** 32 - 1, 1079 * 10007 + 1919), (7 * 17, 99)]:
        assert np.array_equal(po.pcg32(seed, seq, 512), gl.pcg32(seed, seq, 512))
        assert np.array_equal(po.rand_floats(seed, seq, 512).view(np.uint32), gl.rand_floats(seed, seq, 512).view(np.uint32))
    # the canonical pcg32 demo stream, through the reference's own rand.glsl
    assert np.array_equal(gl.pcg32(42, 54, 3), np.array([0xA15C02B7, 0x7B47F409, 0xBA1D3330], np.uint32))


def test_evaluate_phat_bitwise():
    rng = np.random.default_rng(5)
    n = 200_000
    args = np.zeros((n, 16), np.float32)
    args[:, 0:3] = rng.uniform(-5, 5, (n, 3))             # worldPos
    args[:, 3:6] = rng.uniform(-8, 8, (n, 3))             # lightPos
    args[:, 6:9] = rng.uniform(-6, 6, (n, 3))             # camPos
    nrm = rng.normal(size=(n, 3))
    args[:, 9:12] = nrm / np.linalg.norm(nrm, axis=1, keepdims=True)
    ln = rng.normal(size=(n, 3))
    args[:, 12:15] = ln / np.linalg.norm(ln, axis=1, keepdims=True)
    args[:, 15] = rng.integers(0, 2, n)
    args[::97, 9:12] = 0.0                                # background normal
    for alb, lum, rough, metal in [(0.6, 3.0, 0.5, 0.0), (0.18, 40.0, 0.02, 1.0), (0.9, 0.5, 1.0, 0.3), (0.0, 1.0, 0.0, 0.0)]:
        a = po.evaluate_phat(args, alb, lum, rough, metal)
        b = gl.evaluate_phat(args, alb, lum, rough, metal)
        assert ph.bits_equal(a, b).all(), f"{(~ph.bits_equal(a, b)).sum()} of {n} differ"


@pytest.mark.parametrize("name,n_rays", [("procedural:point", 100_000), ("cornellBox", 100_000), ("sponza", 150_000), ("office", 100_000)])
def test_visibility_bits(name, n_rays):
    scene = _scene(name)
    sc = ph.oracle_scene(scene)
    rng = np.random.default_rng(99)
    w, h = 128, 72
    g = po.raycast_gbuffer(sc, scene.tri_material, scene.material_table(), _cams(name, 1, w, h)[0], w, h)
    surf = g.world_pos.reshape(-1, 4)[:, :3]
    lo, hi = scene.dims[:3], scene.dims[3:]
    half = n_rays // 2
    p1 = np.concatenate([surf[rng.integers(0, surf.shape[0], half)], rng.uniform(lo, hi, (n_rays - half, 3))]).astype(np.float32)
    p2 = rng.uniform(lo, hi, (n_rays, 3)).astype(np.float32)
    p2[::50] = p1[::50]                                   # zero-length segments: NaN directions
    p2[1::50, 0] = p1[1::50, 0]                           # axis-parallel: a zero direction component
    a = po.trace_segments(sc, p1, p2)
    b = gl.trace_segments(sc, p1, p2)
    assert np.array_equal(a, b), f"{(a != b).sum()} of {n_rays} visibility bits differ"
    assert 0.05 < a.mean() < 0.98


# ---- whole frame sequences: every pass, every pixel ---------------------------------------------------------------

FRAME_CASES = [
    ("procedural:point", (96, 54), 3, dict(unbiased=False, spatial_iterations=1)),
    ("procedural:point", (96, 54), 3, dict(unbiased=True)),
    ("procedural:tri", (96, 54), 3, dict(unbiased=False, spatial_iterations=2)),
    ("procedural:tri", (96, 54), 3, dict(unbiased=True, unbiased_neighbors=5)),
    ("procedural:random", (96, 54), 2, dict(unbiased=True, candidates=64)),
    ("procedural:point", (96, 54), 2, dict(unbiased=False, flags=0)),
    ("procedural:point", (96, 54), 3, dict(unbiased=True, flags=2)),
    ("procedural:point", (97, 61), 3, dict(unbiased=False, neighbors=5, gamma=2.2)),
    ("cornellBox", (160, 90), 3, dict(unbiased=False, spatial_iterations=1)),
    ("sponza", (160, 90), 3, dict(unbiased=False, spatial_iterations=1, neighbors=4)),
    ("sponza", (160, 90), 3, dict(unbiased=True, unbiased_neighbors=3)),
    ("sponza", (128, 72), 2, dict(unbiased=True, unbiased_neighbors=5)),
    ("office", (128, 72), 2, dict(unbiased=True)),
]


@pytest.mark.parametrize("name,size,frames,kw", FRAME_CASES, ids=[f"{c[0]}-{c[1][0]}x{c[1][1]}-{i}" for i, c in enumerate(FRAME_CASES)])
def test_frames_oracle_equals_reference_source(name, size, frames, kw):
    scene = _scene(name)
    w, h = size
    case = ph.Case(scene, w, h, _cams(name, frames, w, h), **kw)
    want = ph.run_oracle(case, passes=gl)     # the reference's shader text
    got = ph.run_oracle(case)                 # the restatement
    for f, (o, r) in enumerate(zip(got, want)):
        n = ph.compare_reservoirs(o["initial"], r["initial"], f"{name} frame {f} after restirOmni (oracle vs reference source)")
        assert n == 0, f"{name} frame {f}: {n} reservoirs differ after restirOmni"
        n = ph.compare_reservoirs(o["reservoirs"], r["reservoirs"], f"{name} frame {f} final (oracle vs reference source)")
        assert n == 0, f"{name} frame {f}: {n} final reservoirs differ"
        a, b = o["rgba"][..., :3], r["rgba"][..., :3]
        if kw.get("gamma", 1.0) == 1.0:
            assert ph.bits_equal(a, b).all(), f"{name} frame {f}: {(~ph.bits_equal(a, b)).sum()} colour values differ"
        else:                                  # P11: powf on both sides, still the same libm here
            rel, psnr = ph.compare_rgb(o["rgba"], r["rgba"])
            assert rel <= ph.RGB_REL_TOL and psnr >= ph.PSNR_MIN_DB
    lit = (want[-1]["reservoirs"]["w"] > 0).mean()
    assert lit > 0.05, "degenerate case: almost nothing is lit"


@pytest.mark.parametrize("label,name,size,frames,kw", ph.EDGE_CASES, ids=[c[0] for c in ph.EDGE_CASES])
def test_edge_cases_oracle_equals_reference_source(label, name, size, frames, kw):
    """Boundary values of the parameter block and of the screen (parity_harness.EDGE_CASES) through the reference's own
    shader text and through the oracle: bit for bit."""
    scene = _scene(name)
    w, h = size
    aspect_cams = _cams(name, frames, max(w, 2), max(h, 2))
    case = ph.Case(scene, w, h, aspect_cams, **kw)
    want = ph.run_oracle(case, passes=gl)
    got = ph.run_oracle(case)
    for f, (o, r) in enumerate(zip(got, want)):
        assert ph.compare_reservoirs(o["initial"], r["initial"], f"{label} frame {f} after restirOmni") == 0
        assert ph.compare_reservoirs(o["reservoirs"], r["reservoirs"], f"{label} frame {f} final") == 0
        assert ph.bits_equal(o["rgba"][..., :3], r["rgba"][..., :3]).all()


@pytest.mark.parametrize("unbiased", [True, False])
@pytest.mark.parametrize("kind", ph.DEGENERATE_LIGHTS)
def test_degenerate_lights_oracle_equals_reference_source(kind, unbiased):
    """NaN and infinity through the reservoirs (parity_harness.degenerate_light_case): the reference's shader text and the
    oracle agree bit for bit, any NaN equal to any NaN."""
    case = ph.degenerate_light_case(kind, unbiased=unbiased)
    want = ph.run_oracle(case, passes=gl)
    got = ph.run_oracle(case)
    for f, (o, r) in enumerate(zip(got, want)):
        assert ph.compare_reservoirs(o["initial"], r["initial"], f"{kind} frame {f} after restirOmni") == 0
        assert ph.compare_reservoirs(o["reservoirs"], r["reservoirs"], f"{kind} frame {f} final") == 0
        assert ph.bits_equal(o["rgba"][..., :3], r["rgba"][..., :3]).all()
