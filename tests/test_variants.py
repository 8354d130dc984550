"""The reference's compile-time variants — RESERVOIR_SIZE > 1 and UNBIASED_MIS (restirStructs.glsl:16-17) — through the parity
chain: the reference's shader sources compiled with those defines (oracle/_ref/libglslref_rs<N>[_mis].so) == the oracle's
templated restatement (oracle_*_variant) bit for bit [CPU]; the oracle == the CUDA path's generic kernels [-m gpu]."""
import numpy as np
import pytest

import parity_harness as ph

fixtures, capi = ph.fixtures, ph.capi
po = ph.oracle()
gl = ph.glsl_reference()

VARIANTS = [(1, False), (2, False), (4, False), (1, True), (2, True)]
CASES = {
    # name: scene kind, size, frames, Case kwargs
    "point-biased": ("procedural:point", (64, 40), 3, dict(unbiased=False, candidates=8, neighbors=4, spatial_iterations=1)),
    "tri-unbiased3": ("procedural:tri", (56, 36), 3, dict(unbiased=True, candidates=6, unbiased_neighbors=3)),
    "point-unbiased5": ("procedural:point", (48, 32), 3, dict(unbiased=True, candidates=8, unbiased_neighbors=5)),
    "tri-biased2x": ("procedural:tri", (40, 28), 2, dict(unbiased=False, candidates=4, neighbors=5, spatial_iterations=2)),
}


def _case(name):
    kind, (w, h), frames, kw = CASES[name]
    scene = fixtures.make_procedural(seed=21, grid=8, boxes=14, lights=kind.split(":")[1], n_point_lights=12)
    cams = ph.moving_cameras(frames, (3.0, 3.5, 4.2), (0.0, -1.0, 0.0), w / h)
    return ph.Case(scene, w, h, cams, **kw)


def _same(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))


def test_variant_one_sample_no_mis_is_the_shipped_configuration():
    """(RESERVOIR_SIZE 1, UNBIASED_MIS off) through the templated restatement == the oracle proper, every byte."""
    for name in CASES:
        case = _case(name)
        a, b = ph.run_oracle(case), ph.run_oracle(case, passes=po.Variant(1, False))
        for fa, fb in zip(a, b):
            assert _same(fa["reservoirs"], fb["reservoirs"]) and _same(fa["initial"], fb["initial"]) and _same(fa["rgba"], fb["rgba"])
            assert fa["rays"] == fb["rays"]


@pytest.mark.skipif(gl is None, reason="oracle/_ref/libglslref*.so not built (needs /root/reference)")
@pytest.mark.parametrize("n,mis", VARIANTS[1:])
@pytest.mark.parametrize("name", list(CASES))
def test_oracle_variants_equal_the_reference_sources_compiled_with_the_same_defines(name, n, mis):
    ref = gl.Variant(n, mis)
    if not ref.available():
        pytest.skip(f"{ref.so} not built")
    case = _case(name)
    want = ph.run_oracle(case, passes=ref)
    got = ph.run_oracle(case, passes=po.Variant(n, mis))
    for f, (fw, fg) in enumerate(zip(want, got)):
        for key in ("initial", "reservoirs", "rgba"):
            assert _same(fw[key], fg[key]), f"{name} RESERVOIR_SIZE={n} MIS={mis} frame {f}: {key} differs"
    last = got[-1]["reservoirs"]
    assert (last["samples"]["w"] > 0).any() and (last["M"] > 0).any()
    if n > 1:   # the samples of one reservoir are drawn independently: they must not all coincide
        assert (last["samples"]["lightIndex"][:, 0] != last["samples"]["lightIndex"][:, 1]).any()
    if mis:
        assert (last["samples"]["sumPHat"] > 0).any()


def test_variant_ray_counts():
    """RESERVOIR_SIZE rays per pixel in restirOmni (:149-160), RESERVOIR_SIZE x (neighbours + 1) at most in unbiasedReuse (:132-166)."""
    case = _case("point-unbiased5")
    px = case.w * case.h
    for n, mis in VARIANTS:
        frames = ph.run_oracle(case, passes=po.Variant(n, mis))
        for f in frames:
            assert n * px * 2 <= f["rays"] <= n * px * (1 + 5 + 1)


# ---- GPU: the generic kernels (csrc/restir_generic.cu) against the oracle's variants ---------------------------------------

def _run_cuda(case, n, mis, fused):
    import torch

    ctx = ph.make_context(case.scene)
    ctx.set_reservoir_variant(n, mis, fused)
    ctx.resize(case.w, case.h)
    ctx.set_unbiased_neighbors(case.unbiased_neighbors)
    gb = case.gbuffers()
    img = torch.zeros((case.h, case.w, 4), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    out = []
    ctx.counters(reset=True)
    for f in range(len(case.cameras)):
        i = f & 1
        u, lu = case.uniforms(f)
        ctx.upload_gbuffer(i, *gb[f].planes())
        ctx.set_uniforms(u)
        ctx.set_lighting_uniforms(lu)
        if case.unbiased:
            ctx.pass_restir(i, capi.RESTIR_BUF_TEMP, i ^ 1)
            initial = ctx.download_reservoirs(capi.RESTIR_BUF_TEMP)
            ctx.pass_unbiased(i, capi.RESTIR_BUF_TEMP, i)
        else:
            ctx.pass_restir(i, i, i ^ 1)
            initial = ctx.download_reservoirs(i)
            for j in range(case.iterations):
                ctx.pass_spatial(i, i, i ^ 1, 2 * j)
                ctx.pass_spatial(i, i ^ 1, i, 2 * j + 1)
        ctx.pass_lighting(i, i, img, capi.RESTIR_OUT_RGBA32F)
        ctx.synchronize()
        c = ctx.counters(reset=True)
        out.append(dict(reservoirs=ctx.download_reservoirs(i), initial=initial, rgba=img.cpu().numpy().copy(), rays=c["shadow_rays"]))
    ctx.close()
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("n,mis", VARIANTS)
@pytest.mark.parametrize("name", list(CASES))
def test_generic_kernels_match_oracle_variants(name, n, mis):
    case = _case(name)
    want = ph.run_oracle(case, passes=po.Variant(n, mis))
    got = _run_cuda(case, n, mis, fused=True)
    for f, (fw, fg) in enumerate(zip(want, got)):
        for key in ("initial", "reservoirs"):
            assert _same(fw[key], fg[key]), f"{name} RESERVOIR_SIZE={n} MIS={mis} frame {f}: {key} differs"
        ph.compare_rgb(fg["rgba"], fw["rgba"])
        assert fg["rays"] == fw["rays"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_fused_passes_equal_the_tuned_path(name):
    """(1, off) through the single-kernel passes == the tuned path (cut shaders, packed reservoirs, sorted lockstep rays): every
    byte of every reservoir and of the image, and the same number of testVisibility calls."""
    case = _case(name)
    a, b = ph.run_cuda(case), _run_cuda(case, 1, False, fused=True)
    for fa, fb in zip(a, b):
        assert _same(fa["reservoirs"], fb["reservoirs"]) and _same(fa["initial"], fb["initial"]) and _same(fa["rgba"], fb["rgba"])
        assert fa["rays"] == fb["rays"]


@pytest.mark.gpu
def test_variant_switch_reallocates_and_rejects_what_is_not_built():
    scene = fixtures.make_procedural(seed=4, grid=8, boxes=10, lights="point", n_point_lights=9)
    ctx = ph.make_context(scene)
    ctx.resize(64, 32)
    assert ctx.reservoir_bytes() == 64
    ctx.set_reservoir_variant(2, True)
    assert ctx.reservoir_bytes() == 144 and ctx.download_reservoirs(0).dtype.itemsize == 144 and not ctx.download_reservoirs(0).view(np.uint8).any()
    ctx.set_reservoir_variant(4, False)
    assert ctx.reservoir_bytes() == 208
    with pytest.raises(capi.RestirError, match="not built"):
        ctx.set_reservoir_variant(3, False)
    with pytest.raises(capi.RestirError):
        ctx.band_local_peer()
    ctx.set_reservoir_variant(1, False)
    assert ctx.reservoir_bytes() == 64
    ctx.band_local_peer()
    ctx.close()
