"""Pins the oracle against the known-answer vectors that exist for this path:
the canonical PCG32 demo stream, the reference's seeding (Appendix D), struct layouts (Appendix A)."""
import ctypes as C

import numpy as np

import parity_harness as ph

po = ph.oracle()


def test_pcg32_canonical_demo_stream():
    # pcg32_srandom_r(42, 54) from pcg-random.org's pcg32-demo; rand.glsl:6 cites that algorithm and
    # seedRand (rand.glsl:20-28) is exactly pcg32_srandom_r
    want = np.array([0xA15C02B7, 0x7B47F409, 0xBA1D3330, 0x83D2F293, 0xBFA4784B, 0xCBED606E], np.uint32)
    assert np.array_equal(po.pcg32(42, 54, 6), want)


def test_reference_seeding_appendix_d():
    # seedRand(frame=1, seq=y*10007+x), SURVEY.md Appendix D
    assert np.array_equal(po.pcg32(1, 0, 3), np.array([0xE2393051, 0x01112F35, 0xD3509D35], np.uint32))
    assert np.array_equal(po.pcg32(1, 3 * 10007 + 5, 3), np.array([0x50C61936, 0xB0F08DE9, 0x08CCDC6F], np.uint32))


def test_rand_float_is_uint_over_2_32_and_reaches_one():
    u = po.pcg32(7, 99, 4096)
    f = po.rand_floats(7, 99, 4096)
    want = (u.astype(np.float32) / np.float32(4294967296.0)).astype(np.float32)   # rand.glsl:31
    assert np.array_equal(f.view(np.uint32), want.view(np.uint32))
    # Appendix B.4: float(u)/2^32 == 1.0f exactly for u >= 0xFFFFFF80
    edge = np.array([0xFFFFFF7F, 0xFFFFFF80, 0xFFFFFFFF], np.uint32).astype(np.float32) / np.float32(4294967296.0)
    assert edge[0] < 1.0 and edge[1] == 1.0 and edge[2] == 1.0


def test_sincos_policy_accuracy():
    a = np.linspace(0.0, 2.0 * np.pi, 200001).astype(np.float32)
    s, c = po.sincos(a)
    assert np.max(np.abs(s - np.sin(a.astype(np.float64)))) < 3e-7
    assert np.max(np.abs(c - np.cos(a.astype(np.float64)))) < 3e-7


def test_struct_layouts_appendix_a():
    assert po.RESERVOIR_DTYPE.itemsize == 64
    assert po.RESERVOIR_DTYPE.fields["lightIndex"][1] == 32 and po.RESERVOIR_DTYPE.fields["M"][1] == 48
    assert po.UNIFORMS_DTYPE.itemsize == 128
    for name, off in [("cameraPos", 64), ("screenSize", 80), ("frame", 88), ("initialLightSampleCount", 92),
                      ("temporalSampleCountMultiplier", 96), ("spatialPosThreshold", 100), ("spatialNormalThreshold", 104),
                      ("spatialNeighbors", 108), ("spatialRadius", 112), ("flags", 116)]:
        assert po.UNIFORMS_DTYPE.fields[name][1] == off, name
    assert po.LIGHTING_UNIFORMS_DTYPE.itemsize == 96
    assert po.LIGHTING_UNIFORMS_DTYPE.fields["debugMode"][1] == 88 and po.LIGHTING_UNIFORMS_DTYPE.fields["gamma"][1] == 92
    assert ph.capi.RESERVOIR_DTYPE == po.RESERVOIR_DTYPE and ph.capi.UNIFORMS_DTYPE == po.UNIFORMS_DTYPE


def test_phat_closed_form_head_on():
    # light straight above a diffuse surface seen from straight above: cosIn = cosOut = cosHalf = cosInHalf = 1
    # => diffuse = albedoLum*(1-metallic)/pi, Fresnel terms 0; specular from GTR2/smithG at normal incidence
    args = np.array([[0, 0, 0, 0, 2, 0, 0, 5, 0, 0, 1, 0, 0, 0, 0, 0]], np.float32)
    rough, metal, alb, lum = 0.5, 0.0, 0.6, 3.0
    got = float(po.evaluate_phat(args, alb, lum, rough, metal)[0])
    a = max(0.001, rough * rough)
    ds = a * a / (np.pi * (1 + (a * a - 1)) ** 2)
    g = 1.0 / (1.0 + max(np.sqrt(a * a + 1 - a * a), 1e-4))
    want = lum * (alb / np.pi + 0.04 * g * g * ds) * (1.0 / 4.0)
    assert abs(got - want) / want < 1e-5
    # behind the surface => 0 (restirUtils.glsl:8-10)
    args[0, 4] = -2
    assert po.evaluate_phat(args, alb, lum, rough, metal)[0] == 0.0


def test_normalisation_divisions_without_the_divider(tmp_path):
    """csrc/restir_pixel.cuh decodes SNORM16 / UNORM16 / UNORM8 texels as q = x c, q + (x - D q) c instead of x / D (a zero numerator
    takes the divider's slow path, and G-buffers are full of zeros): tools/const_div_check.c compares the sequence with the division
    for every integer the three formats hold."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "const_div_check")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-o", exe, os.path.join(root, "tools", "const_div_check.c"), "-lm"])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.split() == ["0", "0", "0"], out.stdout
