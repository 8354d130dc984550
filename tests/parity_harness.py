"""Shared machinery of the parity tests: builds a test case (scene, cameras, G-buffers, uniforms), runs the
same frame sequence through the CPU oracle and through the CUDA path (via the C ABI), and compares.

Parity criteria (made concrete):
  * shadow-ray visibility bits: identical for every ray;
  * reservoir integer fields (lightIndex, M): identical;
  * reservoir float fields (position, pHat, sumWeights, w, and the normal / emissionLum the 64-byte
    layout carries): BIT-identical (the oracle and the kernels implement the same arithmetic policy);
    NaN equals NaN;
  * final linear RGB: max relative error <= 1e-3 (abs floor 1e-6) and PSNR >= 60 dB — it is bit-exact in
    practice when gamma == 1, the tolerance covers powf when gamma != 1.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
capi, fixtures = pkg.capi, pkg.fixtures

RGB_REL_TOL = 1e-3
RGB_ABS_FLOOR = 1e-6
PSNR_MIN_DB = 60.0


def oracle():
    return graft.load_oracle()


def glsl_reference():
    """TEST INFRASTRUCTURE: the reference's own shader sources compiled as C++ (oracle/pyglslref.py); None where
    oracle/_ref/libglslref.so was never built (it is built wherever /root/reference exists and travels from there)."""
    graft.load_oracle()
    import importlib.util

    if "pyglslref" not in sys.modules:
        spec = importlib.util.spec_from_file_location("pyglslref", os.path.join(ROOT, "oracle", "pyglslref.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules["pyglslref"] = mod
        spec.loader.exec_module(mod)
    mod = sys.modules["pyglslref"]
    return mod if mod.available() else None


def oracle_scene(scene):
    po = oracle()
    return po.Scene(scene.nodes, scene.triangles, scene.point_blob, scene.tri_blob, scene.alias_blob)


class Case:
    """A frame sequence: per frame a camera; G-buffers rendered by the oracle's fixture generator."""

    def __init__(self, scene, width, height, cameras, candidates=32, unbiased=False, spatial_iterations=1, neighbors=4,
                 unbiased_neighbors=3, flags=3, multiplier=20, radius=30.0, pos_thr=0.1, nor_thr=25.0, gamma=1.0, frame_offset=0):
        self.scene, self.w, self.h = scene, width, height
        self.cameras = cameras
        self.candidates, self.unbiased, self.iterations = candidates, unbiased, spatial_iterations
        self.neighbors, self.unbiased_neighbors = neighbors, unbiased_neighbors
        self.flags, self.multiplier, self.radius = flags, multiplier, radius
        self.pos_thr, self.nor_thr, self.gamma = pos_thr, nor_thr, gamma
        self.frame_offset = frame_offset          # first frame number = 1 + frame_offset (mod 2^32: the seeds wrap, rand.glsl:20-28)
        self._gbuffers = None

    def gbuffers(self):
        if self._gbuffers is None:
            po = oracle()
            sc = oracle_scene(self.scene)
            table = self.scene.material_table()
            self._gbuffers = [po.raycast_gbuffer(sc, self.scene.tri_material, table, cam, self.w, self.h) for cam in self.cameras]
        return self._gbuffers

    def uniforms(self, frame_index):
        """frame_index 0 is the reference's first rendered frame (frame == 1, app.cpp:776)."""
        po = oracle()
        cam = self.cameras[frame_index]
        prev_cam = self.cameras[max(frame_index - 1, 0)]
        u = capi.make_uniforms(
            prevFrameProjectionViewMatrix=po.camera_matrix(prev_cam),
            cameraPos=(cam.position[0], cam.position[1], cam.position[2], 1.0),
            screenSize=(self.w, self.h), frame=(frame_index + 1 + self.frame_offset) & 0xFFFFFFFF, initialLightSampleCount=self.candidates,
            temporalSampleCountMultiplier=self.multiplier, spatialPosThreshold=self.pos_thr,
            spatialNormalThreshold=self.nor_thr, spatialNeighbors=self.neighbors, spatialRadius=self.radius, flags=self.flags)
        lu = capi.make_lighting_uniforms(
            prevFrameProjectionViewMatrix=po.camera_matrix(prev_cam),
            cameraPos=(cam.position[0], cam.position[1], cam.position[2], 1.0), bufferSize=(self.w, self.h), debugMode=0,
            gamma=self.gamma)
        return u, lu


def moving_cameras(n, position, look_at, aspect, step=(0.05, 0.0, 0.0), fov_y=None):
    po = oracle()
    cams = []
    for i in range(n):
        p = tuple(np.float32(position[k]) + np.float32(i) * np.float32(step[k]) for k in range(3))
        cams.append(po.make_camera(position=p, look_at=look_at, aspect=aspect, fov_y=fov_y))
    return cams


def to_capi_camera(cam):
    return capi.make_camera(position=tuple(cam.position), look_at=tuple(cam.lookAt), up=tuple(cam.worldUp), z_near=cam.zNear,
                            z_far=cam.zFar, fov_y=cam.fovYRadians, aspect=cam.aspectRatio)


# Boundary cases of the parameter block and of the screen, run through every link of the parity chain
# (reference source == oracle in tests/test_oracle_vs_glsl.py, oracle == kernels in tests/test_gpu_parity.py).
EDGE_CASES = [
    # id, scene, (w, h), frames, kwargs
    ("one-candidate-biased", "procedural:point", (64, 36), 2, dict(unbiased=False, candidates=1)),
    ("one-candidate-unbiased", "procedural:tri", (64, 36), 2, dict(unbiased=True, candidates=1)),
    ("screen-1x1", "procedural:point", (1, 1), 3, dict(unbiased=True)),
    ("screen-8x4-one-warp-tile", "procedural:point", (8, 4), 3, dict(unbiased=False)),
    ("screen-33x9-ragged", "procedural:tri", (33, 9), 3, dict(unbiased=True, unbiased_neighbors=5)),
    ("radius-0-self-neighbour", "procedural:point", (64, 36), 2, dict(unbiased=False, radius=0.0)),
    ("radius-0-unbiased", "procedural:point", (64, 36), 2, dict(unbiased=True, radius=0.0)),
    ("radius-1000-clamped-to-screen", "procedural:point", (64, 36), 2, dict(unbiased=True, radius=1000.0)),
    ("m-cap-0", "procedural:point", (64, 36), 3, dict(unbiased=False, multiplier=0)),
    ("m-cap-1", "procedural:tri", (64, 36), 3, dict(unbiased=True, multiplier=1)),
    ("thresholds-0-reject-all", "procedural:point", (64, 36), 2, dict(unbiased=False, pos_thr=0.0, nor_thr=0.0)),
    ("frame-number-wraps", "procedural:point", (64, 36), 3, dict(unbiased=False, frame_offset=2 ** 32 - 3)),
    ("frame-number-wraps-unbiased", "procedural:tri", (64, 36), 3, dict(unbiased=True, frame_offset=2 ** 32 - 3)),
    ("visibility-off-temporal-on", "procedural:tri", (64, 36), 3, dict(unbiased=True, flags=2)),
    ("visibility-on-temporal-off", "procedural:point", (64, 36), 2, dict(unbiased=True, flags=1)),
]


def degenerate_light_case(kind, w=48, h=27, unbiased=True):
    """Lights that drive the arithmetic to its edges: `on-surface` puts a light exactly at a visible surface point (zero
    distance: 0 * inf = NaN in normalize), `zero-luminance` adds lights whose alias probability is 0 (p-hat / 0),
    `huge` lets p-hat overflow to infinity.  NaN and infinity must flow through the reservoirs exactly as in the reference."""
    base = fixtures.make_procedural(seed=21, grid=10, boxes=20, lights="point", n_point_lights=6)
    cams = moving_cameras(2, (3.0, 3.5, 4.2), (0.0, -1.0, 0.0), w / h)
    probe = Case(base, w, h, cams[:1])
    world = probe.gbuffers()[0].world_pos.reshape(-1, 4)
    normal = probe.gbuffers()[0].normal.reshape(-1, 4)
    surface = np.flatnonzero((normal[:, :3] != 0).any(axis=1))
    rng = np.random.default_rng(3)
    pos = rng.uniform(-3, 3, (5, 3)).astype(np.float32)
    pos[:, 1] = np.abs(pos[:, 1]) + 1.0
    col = rng.uniform(0.2, 1.0, (5, 3)).astype(np.float32)
    if kind == "on-surface":
        pick = surface[[len(surface) // 3, len(surface) // 2, (2 * len(surface)) // 3]]
        pos[:3] = world[pick, :3]
    elif kind == "zero-luminance":
        col[1] = 0.0
        col[3] = 0.0
    elif kind == "huge":
        col[0] = np.float32(3e38)
        col[2] = np.float32(1e30)
    else:
        raise ValueError(kind)
    scene = fixtures.with_point_lights(base, pos, col)
    return Case(scene, w, h, cams, candidates=16, unbiased=unbiased, unbiased_neighbors=3)


DEGENERATE_LIGHTS = ["on-surface", "zero-luminance", "huge"]


# ---- running both sides -----------------------------------------------------------------------------

def run_oracle(case, rows=None, passes=None):
    """Returns per frame: dict(reservoirs=final (N,) RESERVOIR_DTYPE, initial=after restirOmni, rgba, rays).
    passes: the module providing the four passes — the oracle (default) or glsl_reference()."""
    po = oracle()
    sc = oracle_scene(case.scene)
    data_po, po = po, (passes or po)
    gb = case.gbuffers()
    n = case.w * case.h
    frame_bufs = [np.zeros(n, po.RESERVOIR_DTYPE), np.zeros(n, po.RESERVOIR_DTYPE)]  # app.h:264-284 zero-filled
    out = []
    for f in range(len(case.cameras)):
        i, p = f & 1, (f & 1) ^ 1
        u, lu = case.uniforms(f)
        u = u.astype(po.UNIFORMS_DTYPE)
        lu = lu.astype(po.LIGHTING_UNIFORMS_DTYPE)
        cur = gb[f]
        prev = gb[f - 1] if f > 0 else None
        rays = 0
        initial, r = po.restir_pass(sc, u, cur, prev, frame_bufs[p], rows)
        rays += r or 0
        if case.unbiased:  # app.h:298-314
            final, r = po.unbiased_pass(sc, u, cur, initial, case.unbiased_neighbors, rows)
            rays += r or 0
            frame_bufs[i] = final
        else:              # app.h:316-332: the previous frame's buffer is the scratch of the spatial passes
            frame_bufs[i] = initial
            for j in range(case.iterations):
                frame_bufs[p] = po.spatial_pass(u, cur, frame_bufs[i], 2 * j, rows)
                frame_bufs[i] = po.spatial_pass(u, cur, frame_bufs[p], 2 * j + 1, rows)
        rgba = po.lighting_pass(sc, lu, cur, frame_bufs[i], rows)
        out.append(dict(reservoirs=frame_bufs[i].copy(), initial=initial, rgba=rgba, rays=rays))
    return out


def make_context(scene, device=0):
    ctx = capi.RestirContext(device)
    ctx.upload_bvh(scene.nodes, scene.triangles)
    ctx.upload_lights(scene.point_blob, scene.tri_blob, scene.alias_blob)
    return ctx


def run_cuda(case, ctx=None):
    import torch

    own = ctx is None
    if own:
        ctx = make_context(case.scene)
    ctx.resize(case.w, case.h)
    ctx.set_unbiased_neighbors(case.unbiased_neighbors)
    gb = case.gbuffers()
    out_img = torch.zeros((case.h, case.w, 4), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()                  # torch fills on its own stream; the context's stream is not ordered against it
    out = []
    ctx.counters(reset=True)
    for f in range(len(case.cameras)):
        i = f & 1
        u, lu = case.uniforms(f)
        ctx.upload_gbuffer(i, *gb[f].planes())
        ctx.set_uniforms(u)
        ctx.set_lighting_uniforms(lu)
        # same order as restir_frame, split so the initial reservoirs can be inspected
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
        ctx.pass_lighting(i, i, out_img, capi.RESTIR_OUT_RGBA32F)
        ctx.synchronize()
        c = ctx.counters(reset=True)
        out.append(dict(reservoirs=ctx.download_reservoirs(i), initial=initial, rgba=out_img.cpu().numpy().copy(), rays=c["shadow_rays"],
                        counters=c))
    if own:
        ctx.close()
    return out


# ---- frame captures (include/restir_capture.h) -----------------------------------------------------------

def make_capture(case, frames_out=None):
    """The capture of a Case: its scene buffers, per-frame uniforms and G-buffers, and — when `frames_out` (the result of
    run_oracle / run_cuda) is given — the outputs to compare a replay with."""
    capture = __import__("restir_vulkan_b200.capture", fromlist=["capture"])
    frames = []
    for f, g in enumerate(case.gbuffers()):
        u, lu = case.uniforms(f)
        fr = dict(uniforms=u, lighting_uniforms=lu, planes=list(g.planes()))
        if frames_out is not None:
            fr.update(initial=frames_out[f]["initial"], final=frames_out[f]["reservoirs"], rgba=frames_out[f]["rgba"])
        frames.append(fr)
    sc = case.scene
    return capture.Capture(case.w, case.h, case.unbiased, case.unbiased_neighbors, case.iterations, sc.nodes, sc.triangles, sc.point_blob,
                           sc.tri_blob, sc.alias_blob, frames)


# ---- comparison -------------------------------------------------------------------------------------

FLOAT_FIELDS = ("position_emissionLum", "normal", "pHat", "sumWeights", "w")
INT_FIELDS = ("lightIndex", "M")


def bits_equal(a, b):
    """Bitwise equality of float arrays, except that any NaN equals any NaN."""
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    same = a.view(np.uint32) == b.view(np.uint32)
    return same | (np.isnan(a) & np.isnan(b))


def compare_reservoirs(got, want, label=""):
    """Returns the number of mismatching reservoirs; prints the first few."""
    bad = np.zeros(got.shape[0], bool)
    for f in INT_FIELDS:
        bad |= got[f] != want[f]
    for f in FLOAT_FIELDS:
        eq = bits_equal(got[f], want[f])
        bad |= ~(eq.reshape(got.shape[0], -1).all(axis=1))
    n = int(bad.sum())
    if n:
        idx = np.flatnonzero(bad)[:5]
        for k in idx:
            print(f"[{label}] reservoir {k} differs:\n  cuda   {got[k]}\n  oracle {want[k]}")
    return n


def compare_rgb(got, want):
    """(max relative error over compared pixels, PSNR in dB).  Non-finite pixels must match as such."""
    g = got[..., :3].astype(np.float64)
    w = want[..., :3].astype(np.float64)
    finite = np.isfinite(g) & np.isfinite(w)
    assert np.array_equal(np.isfinite(g), np.isfinite(w)), "non-finite pixels differ"
    diff = np.abs(np.where(finite, g - w, 0.0))
    rel = diff / np.maximum(np.abs(np.where(finite, w, 0.0)), RGB_ABS_FLOOR)
    rel = np.where(diff <= RGB_ABS_FLOOR, 0.0, rel)
    mse = float(np.mean(diff ** 2))
    peak = max(float(np.max(np.abs(np.where(finite, w, 0.0)))), 1e-12)
    psnr = float("inf") if mse == 0 else 10.0 * np.log10(peak * peak / mse)
    return float(rel.max()), psnr


def assert_frames_match(cuda_frames, oracle_frames, label=""):
    for f, (c, o) in enumerate(zip(cuda_frames, oracle_frames)):
        assert c["counters"]["stack_overflows"] == 0, f"{label} frame {f}: traversal stack overflow"
        assert c["counters"]["halo_misses"] == 0, f"{label} frame {f}: halo miss"
        n_init = compare_reservoirs(c["initial"], o["initial"], f"{label} frame {f} initial")
        assert n_init == 0, f"{label} frame {f}: {n_init} reservoirs differ after restirOmni"
        n_fin = compare_reservoirs(c["reservoirs"], o["reservoirs"], f"{label} frame {f} final")
        assert n_fin == 0, f"{label} frame {f}: {n_fin} final reservoirs differ"
        assert c["rays"] == o["rays"], f"{label} frame {f}: ray count {c['rays']} != {o['rays']}"
        rel, psnr = compare_rgb(c["rgba"], o["rgba"])
        assert rel <= RGB_REL_TOL and psnr >= PSNR_MIN_DB, f"{label} frame {f}: rgb rel {rel} psnr {psnr}"


# ---- smoke (called by __graft_entry__.smoke on the GPU box) -------------------------------------------

def smoke_case():
    scene = fixtures.make_procedural(seed=3, grid=8, boxes=12, lights="point", n_point_lights=12)
    cams = moving_cameras(2, (3.0, 3.5, 4.2), (0.0, -1.0, 0.0), 96 / 64)
    return Case(scene, 96, 64, cams, candidates=8, unbiased=True)


def run_smoke():
    case = smoke_case()
    got = run_cuda(case)
    want = run_oracle(case)
    assert_frames_match(got, want, "smoke")
    print(f"smoke ok: {len(got)} unbiased frames of {case.w}x{case.h}, {got[-1]['rays']} shadow rays in the last frame, "
          f"{got[-1]['counters']['kernel_launches']} kernel launches, bit-identical to the oracle")
