#!/usr/bin/env python
"""bench.py — ReSTIR ms/frame + shadow Mrays/s on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config NAME] [--no-suite]

A "step" is one frame of the hot path: restirOmni -> (unbiased reuse | 2 x spatial reuse) -> lighting, on
synthetic inputs of the configuration's shape (G-buffers ray-cast from the scene by the fixture tool, two
camera positions alternating so temporal reprojection does real work).

The JSON line (one, on stdout, rank 0):
  * headline (`value`, `ms_per_step`, `e2e`, `roofline`, `cpu_baseline`, ...): the configuration BASELINE.json's
    metric is quoted on — Sponza 1920x1080, 32 candidates, 200 random point lights, unbiased reuse with the
    north-star's 5 spatial neighbours.  N > 1 (torchrun, one rank per GPU): every rank a 1080-row band of a
    1920 x (1080 N) frame (weak scaling), halo rows exchanged by the library's own kernels over NVLink peer memory.
  * `configs`: every other BASELINE configuration, measured in the same run — at N = 1 cornellBox 720p biased,
    Sponza 1080p biased with 4 and 5 neighbours, Sponza unbiased with the reference's 3 neighbours, office 2160p with
    triangle lights; at EVERY N `strong_8k`: the 7680x4320 frame with 1 M point lights and 64 candidates split into
    N row bands (strong scaling).  Each with ms/frame, walked Mrays/s, HBM fraction, clocks, and its parity record.
  * parity records: N = 1 `parity_sample` (>= 64 full-width rows of frame 2 against the CPU oracle, bit for bit);
    N > 1 `band_parity` (three frames through the connected bands against a whole-screen context, every owned
    reservoir and RGBA8 pixel bit for bit).  A run whose band_parity, halo counters or stack counters are not clean
    prints `value` 0 with the reason in `invalid` and exits 1.

`value` = shadow Mrays/s = the reference's testVisibility calls answered per second over the whole frame time (the
unit of work the `--impl reference` arm shares); `value_walked` = the rays that needed a walk of the tree.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

METRIC = "shadow_mrays_per_s"
UNIT = "Mrays/s"

CONFIGS = {
    # name: scene, (w, h), unbiased, neighbours, candidates, point-light override
    "sponza_1080p_unbiased5": dict(scene="sponza", size=(1920, 1080), unbiased=True, neighbors=5, candidates=32, lights=None),
    "sponza_1080p_unbiased3": dict(scene="sponza", size=(1920, 1080), unbiased=True, neighbors=3, candidates=32, lights=None),
    "sponza_1080p_biased4": dict(scene="sponza", size=(1920, 1080), unbiased=False, neighbors=4, candidates=32, lights=None),
    "sponza_1080p_biased5": dict(scene="sponza", size=(1920, 1080), unbiased=False, neighbors=5, candidates=32, lights=None),
    "cornell_720p_biased4": dict(scene="cornellBox", size=(1280, 720), unbiased=False, neighbors=4, candidates=32, lights=None),
    "office_2160p_unbiased3": dict(scene="office", size=(3840, 2160), unbiased=True, neighbors=3, candidates=32, lights=None),
    "sponza_8k_1m_lights": dict(scene="sponza", size=(7680, 4320), unbiased=True, neighbors=5, candidates=64, lights=1_000_000),
    # small frames for the CPU test suite (tests/test_bench_inputs.py runs both kinds of the reference arm on them)
    "cornell_tiny_biased4": dict(scene="cornellBox", size=(160, 90), unbiased=False, neighbors=4, candidates=8, lights=None),
    "cornell_tiny_unbiased3": dict(scene="cornellBox", size=(160, 96), unbiased=True, neighbors=3, candidates=8, lights=None),
}
HEADLINE = "sponza_1080p_unbiased5"
# BASELINE.json configs[0..3] beside the headline (N = 1), and configs[4] at every N (strong scaling)
SUITE_N1 = ["cornell_720p_biased4", "sponza_1080p_biased4", "sponza_1080p_biased5", "sponza_1080p_unbiased3", "office_2160p_unbiased3"]
STRONG = "sponza_8k_1m_lights"
CAMERAS = {
    "sponza": ((3.0, 4.0, 5.0), (0.0, 0.0, 0.0)),       # src/camera.h:7-13 defaults
    "cornellBox": ((3.0, 4.0, 5.0), (0.0, 0.0, 0.0)),
    "office": ((3.0, 1.7, 0.5), (3.0, 1.5, -5.0)),      # inside the room
    "procedural": ((3.0, 3.5, 4.2), (0.0, -1.0, 0.0)),
}
HALO = 31  # ceil(spatialRadius = 30) + 1


def load_scene(fixtures, cfg):
    name = cfg["scene"]
    if fixtures.baked_available(name):
        scene = fixtures.load_baked(name, rebuild=True)
        label = name
    else:
        # the reference scenes are baked from /root/reference where it exists and travel with the repo
        # snapshot; without them fall back to a procedural room of comparable triangle count, and say so
        scene = fixtures.make_procedural(seed=7, grid=150, boxes=4000, lights="random")
        label = f"procedural-fallback({scene.n_triangles} tris; scenes/_baked/{name} missing)"
        name = "procedural"
    if cfg["lights"]:
        scene = fixtures.with_random_point_lights(scene, cfg["lights"])
    return scene, name, label


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe), one line every 20 ms."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines, self.first, self.last = index, None, [], 0, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def wait_alive(self, seconds=3.0):
        """nvidia-smi needs a moment for its first line: the timed region does not start before the sampler answers."""
        t0 = time.time()
        while self.proc is not None and not self.lines and time.time() - t0 < seconds:
            time.sleep(0.01)

    def mark(self):
        """The timed region starts here (the sampler was started during the warm-up)."""
        self.first = len(self.lines)

    def mark_end(self):
        time.sleep(0.03)                                   # the line that covers the end of the region
        self.last = len(self.lines)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines[self.first:self.last]:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# the CPU arm: the oracle (a restatement of the reference's shaders; Vulkan/lavapipe does not exist in
# this image) on all host cores

def oracle_rows(po, sc, cfg, frame_inputs, windows, passes=None):
    """Run the CPU passes of frame index 1 (temporal active) on `windows` = [(y0, y1), ...] full-width row ranges.

    passes: the module whose passes are timed — the oracle port (default) or oracle/pyglslref.py, the reference's
    own shader sources compiled for the CPU (then the rays are counted by the oracle afterwards, untimed: the two
    are bit-identical, tests/test_oracle_vs_glsl.py).  Every pass runs on exactly the rows it needs: the first pass
    also on the HALO-row aprons the reuse pass gathers from.
    Returns seconds (as run), seconds per pass extrapolated to the full frame by rows covered, rays, reservoirs.
    """
    w, h = frame_inputs["size"]
    u, lu = frame_inputs["uniforms"], frame_inputs["lighting_uniforms"]
    pm = passes or po
    total = dict(restir=0.0, reuse=0.0, lighting=0.0)
    rows_done = dict(restir=0, reuse=0, lighting=0)
    rays = rays_scaled = 0.0
    final = initial = None
    for (y0, y1) in windows:
        two = 1 if cfg["unbiased"] else 2
        a0, a1 = max(0, y0 - two * HALO), min(h, y1 + two * HALO)       # biased: two gather hops
        b0, b1 = max(0, y0 - HALO), min(h, y1 + HALO)
        t0 = time.perf_counter()
        ini, rays_a = pm.restir_pass(sc, u, frame_inputs["g_cur"], frame_inputs["g_prev"], frame_inputs["prev_reservoirs"], (a0, a1))
        t1 = time.perf_counter()
        if cfg["unbiased"]:
            fin, rays_b = pm.unbiased_pass(sc, u, frame_inputs["g_cur"], ini, cfg["neighbors"], (y0, y1))
            reuse_rows = y1 - y0
        else:
            mid = pm.spatial_pass(u, frame_inputs["g_cur"], ini, 0, (b0, b1))
            fin = pm.spatial_pass(u, frame_inputs["g_cur"], mid, 1, (y0, y1))
            rays_b, reuse_rows = 0, (b1 - b0) + (y1 - y0)
        t2 = time.perf_counter()
        pm.lighting_pass(sc, lu, frame_inputs["g_cur"], fin, (y0, y1))
        t3 = time.perf_counter()
        if pm is not po:                                   # count the rays of the same rows, untimed
            _, rays_a = po.restir_pass(sc, u, frame_inputs["g_cur"], frame_inputs["g_prev"], frame_inputs["prev_reservoirs"], (a0, a1))
            if cfg["unbiased"]:
                _, rays_b = po.unbiased_pass(sc, u, frame_inputs["g_cur"], ini, cfg["neighbors"], (y0, y1))
        total["restir"] += t1 - t0
        total["reuse"] += t2 - t1
        total["lighting"] += t3 - t2
        rows_done["restir"] += a1 - a0
        rows_done["reuse"] += reuse_rows
        rows_done["lighting"] += y1 - y0
        rays += rays_a + rays_b
        rays_scaled += rays_a * (y1 - y0) / (a1 - a0) + rays_b          # rays of the window's own rows
        if final is None:
            final, initial = fin.copy(), ini.copy()
        else:
            final[y0 * w: y1 * w] = fin[y0 * w: y1 * w]
            initial[a0 * w: a1 * w] = ini[a0 * w: a1 * w]
    own_rows = sum(y1 - y0 for y0, y1 in windows)
    reuse_full = h * (1 if cfg["unbiased"] else 2)
    frame_seconds = (total["restir"] * h / rows_done["restir"] + total["reuse"] * reuse_full / rows_done["reuse"]
                     + total["lighting"] * h / rows_done["lighting"])
    return dict(seconds=sum(total.values()), frame_seconds=frame_seconds, rays=rays, rays_frame=rays_scaled * h / own_rows,
                own_rows=own_rows, final=final, initial=initial,
                per_pass_ms=dict(restir=total["restir"] * h / rows_done["restir"] * 1e3, reuse=total["reuse"] * reuse_full / rows_done["reuse"] * 1e3,
                                 lighting=total["lighting"] * h / rows_done["lighting"] * 1e3))


def make_uniform_blocks(mk, matrix, cfg, w, h, cams, f):
    """mk: module with make_uniforms / make_lighting_uniforms (the product's capi or the oracle's binding)."""
    cam, prev_cam = cams[f & 1], cams[(f & 1) ^ 1] if f > 0 else cams[0]
    pv_prev = matrix(prev_cam)
    u = mk.make_uniforms(prevFrameProjectionViewMatrix=pv_prev, cameraPos=(cam.position[0], cam.position[1], cam.position[2], 1.0),
                         screenSize=(w, h), frame=f + 1, initialLightSampleCount=cfg["candidates"], temporalSampleCountMultiplier=20,
                         spatialPosThreshold=0.1, spatialNormalThreshold=25.0, spatialNeighbors=cfg["neighbors"], spatialRadius=30.0,
                         flags=3)   # src/app.h:150-174, app.cpp:414-434 defaults
    lu = mk.make_lighting_uniforms(prevFrameProjectionViewMatrix=pv_prev, cameraPos=(cam.position[0], cam.position[1], cam.position[2], 1.0),
                                   bufferSize=(w, h), debugMode=0, gamma=1.0)
    return u, lu


def _load_by_path(name, path):
    import importlib.util
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_glsl_reference(cfg):
    """oracle/_ref/libglslref.so (the reference's own shader sources compiled for the CPU, oracle/ref_build), or None
    where it was never built or was not compiled for this neighbour count."""
    graft.load_oracle()
    mod = _load_by_path("pyglslref", os.path.join(ROOT, "oracle", "pyglslref.py"))
    if not mod.available() or (cfg["unbiased"] and cfg["neighbors"] not in (3, 5)):
        return None
    return mod


def run_reference_arm(args):
    """--impl reference: the reference's CPU arm alone.  Nothing of restir-vulkan_b200/ is imported or loaded: the scene is
    the reference's own AabbTree / light / alias blobs (oracle/ref_inputs.py), the passes are the reference's own shader
    sources compiled for the host (oracle/_ref/libglslref.so) when that was built, else the oracle port; G-buffers by the
    oracle's ray caster.  N = 1: every step is one FULL frame (no sampling, no scaling).  N > 1: the GPU arm's frame is
    1920 x 1080 N; a step covers three row windows (top, middle, bottom) of 1/12 of that frame each — 25 % of it —
    and `value` is the throughput of exactly that work (rays of the sample / seconds of the sample), `ms_per_step` the
    time the step really took."""
    po = graft.load_oracle()
    ri = _load_by_path("ref_inputs", os.path.join(ROOT, "oracle", "ref_inputs.py"))
    po.set_num_threads(os.cpu_count() or 1)               # torchrun exports OMP_NUM_THREADS=1; this arm uses every host core
    cfg = CONFIGS[args.config]
    if cfg["lights"] or not ri.available(cfg["scene"]):
        print(json.dumps({"impl": "reference", "unavailable": f"no reference-baked inputs for {args.config} (scenes/_baked/{cfg['scene']}/ref_*.bin, "
                                                                "light override needs the product's host builders)"}))
        return
    gl = load_glsl_reference(cfg)
    scene = ri.ReferenceScene(cfg["scene"])
    world = max(1, args.gpus)
    w, band_h = cfg["size"]
    h = band_h * world if args.scaling == "weak" else band_h
    pos, look = CAMERAS[cfg["scene"]]
    cams = [po.make_camera(position=pos, look_at=look, aspect=w / h),
            po.make_camera(position=(pos[0] + 0.05, pos[1], pos[2]), look_at=look, aspect=w / h)]
    sc = po.Scene(scene.nodes, scene.triangles, scene.point_blob, scene.tri_blob, scene.alias_blob)
    if world == 1:
        windows, sample = [(0, h)], f"every step is the full {w}x{h} frame"
    else:
        rows = h // 12
        windows = [(0, rows), (h // 2 - rows // 2, h // 2 - rows // 2 + rows), (h - rows, h)]
        sample = (f"rows {windows} of the GPU arm's {w}x{h} frame (3 windows, 25 % of it; the first pass also on the {HALO}-row aprons); "
                  "value = rays of the sample / seconds of the sample")
    need = np.zeros(h, bool)
    for y0, y1 in windows:
        need[max(0, y0 - 3 * HALO): min(h, y1 + 3 * HALO)] = True
    spans, y = [], 0
    while y < h:
        if need[y]:
            e = y
            while e < h and need[e]:
                e += 1
            spans.append((y, e))
            y = e
        else:
            y += 1
    table = scene.material_table()
    gbufs = []
    for c in cams:
        g = None
        for span in spans:                                  # only the rows the windows read
            part = po.raycast_gbuffer(sc, scene.tri_material, table, c, w, h, span)
            if g is None:
                g = part
            else:
                for a, b in zip(g.planes(), part.planes()):
                    a[span[0]: span[1]] = b[span[0]: span[1]]
        gbufs.append(g)
    prev = np.zeros(w * h, po.RESERVOIR_DTYPE)
    secs, rays = [], []
    for step in range(args.warmup + args.steps):
        f = step
        u, lu = make_uniform_blocks(po, po.camera_matrix, cfg, w, h, cams, f)
        inputs = dict(size=(w, h), g_cur=gbufs[f & 1], g_prev=gbufs[(f & 1) ^ 1] if f > 0 else None, prev_reservoirs=prev,
                      uniforms=u, lighting_uniforms=lu)
        r = oracle_rows(po, sc, cfg, inputs, windows, passes=gl)
        prev = r["final"]    # valid on the windows, which is where the next step reprojects to
        if step >= args.warmup:
            secs.append(r["seconds"])
            rays.append(r["rays"])
    ms = float(np.mean(secs)) * 1e3
    mrays = float(np.sum(rays)) / float(np.sum(secs)) / 1e6
    cores = po.num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": mrays, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": f"{args.config}: {scene.name} {w}x{h}, {sample}", "note": (
            "the reference's own shader sources (restirOmni.glsl, unbiasedReuse.glsl / spatialReuse.comp, lighting.frag) compiled for "
            "the host by g++ with OpenMP (oracle/_ref/libglslref.so)" if gl else "CPU restatement of the compute-shader path "
            "(oracle/restir_oracle.cpp, OpenMP)") + " on the reference's own AabbTree / light / alias blobs — not lavapipe: no Vulkan in this image"},
        "rays_per_step": float(np.mean(rays)),
        "ms_per_frame": ms if world == 1 else None,
        # evidence that this arm ran without the product: what of this repository the process has loaded
        "native_so_loaded": sorted({ln.split()[-1][len(ROOT) + 1:] for ln in open("/proc/self/maps") if ln.rstrip().endswith(".so") and ROOT in ln}),
        "product_modules_loaded": sorted(m for m in sys.modules if m.startswith("restir_vulkan_b200")),
        "cpu_baseline": {"value": mrays, "unit": UNIT, "cores": cores, "kind": "reference" if gl else "port", "sample": sample},
        "e2e": {"value": mrays, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
# the GPU arm

class Env:
    pass


def run_gpu_config(env, name, scaling, steps, warmup, full):
    """Measure one configuration on all ranks.  full: the headline treatment (per-pass brackets, end-to-end leg, timed CPU
    baseline); otherwise device-resident timing, per-kernel times and the parity record.  Returns the record on rank 0."""
    torch, dist, capi, fixtures, bands, args = env.torch, env.dist, env.capi, env.fixtures, env.bands, env.args
    rank, world, local_rank, stream, dev = env.rank, env.world, env.local_rank, env.stream, env.dev
    cfg = CONFIGS[name]
    scene, cam_key, label = load_scene(fixtures, cfg)
    w, band_h = cfg["size"]
    if scaling == "strong":                               # the configuration's own frame, split into N row bands
        h, band_h = band_h, (band_h + world - 1) // world
    else:
        h = band_h * world                                # weak scaling: one config-sized band per rank
    pos, look = CAMERAS[cam_key]
    cams = [capi.make_camera(position=pos, look_at=look, aspect=w / h),
            capi.make_camera(position=(pos[0] + 0.05, pos[1], pos[2]), look_at=look, aspect=w / h)]
    k = cfg["neighbors"]
    variant = (args.reservoir_size, args.unbiased_mis, args.fused_passes)
    if variant != (1, False, False) and world > 1:
        raise SystemExit("bench.py: reservoir variants run on one GPU (the halo kernels move packed reservoirs)")

    def new_context():
        c = capi.RestirContext(local_rank, stream.cuda_stream)
        if args.traversal != "auto":
            c.set_traversal({"reference-order": capi.RESTIR_TRAVERSAL_REFERENCE_ORDER, "image": capi.RESTIR_TRAVERSAL_IMAGE}[args.traversal])
        c.upload_bvh(scene.nodes, scene.triangles)
        c.upload_lights(scene.point_blob, scene.tri_blob, scene.alias_blob)
        c.set_unbiased_neighbors(k if cfg["unbiased"] else 3)
        c.set_spatial_staging(args.spatial_staging == "on")
        c.set_occluder_cache({"on": 1, "off": 0, "light": 2, "direction": 3}[args.occluder_cache])
        c.set_ray_elision({"on": 1, "off": 0, "dedupe": 2}[args.ray_elision])
        if variant != (1, False, False):
            c.set_reservoir_variant(*variant)
        return c

    ctx = new_context()
    tm = torch.from_numpy(np.ascontiguousarray(scene.tri_material)).to(dev)
    mt = torch.from_numpy(scene.material_table().view(np.int32)).to(dev)

    def render_gbuffers(c, rows_):
        gb_ = []
        for cam in cams:
            planes = [torch.zeros((rows_, w, 4), dtype=torch.uint8, device=dev), torch.zeros((rows_, w, 4), dtype=torch.int16, device=dev),
                      torch.zeros((rows_, w, 2), dtype=torch.int16, device=dev), torch.zeros((rows_, w, 4), dtype=torch.float32, device=dev),
                      torch.zeros((rows_, w), dtype=torch.float32, device=dev)]
            c.raycast_gbuffer(cam, tm, mt, *planes)
            gb_.append(planes)
        c.synchronize()
        return gb_

    def setup_band(rows, halo_rows):
        """Allocates the band `rows` with `halo_rows` of halo and renders the two synthetic G-buffers (one per camera)
        on the device with the fixture tool."""
        if world > 1:
            ctx.resize_band(w, h, rows[0], rows[1], halo_rows)
        else:
            ctx.resize(w, h)
        _, _, a0_, a1_ = ctx.band()
        return a0_, a1_, a1_ - a0_, render_gbuffers(ctx, a1_ - a0_)

    peer = world > 1 and args.halo == "peer"

    def build_bands(bounds, halo_rows):
        """Band of this rank under `bounds` with a halo that also covers the rows temporal reprojection reaches
        (a host-computed bound; the two cameras alternate, so both directions count; every rank uses the
        largest reach)."""
        rows = bands.band_rows(h, world, rank, bounds)
        a0_, a1_, rows_alloc_, gb_ = setup_band(rows, halo_rows)
        if world > 1:
            reach = 0
            for cur, prv in ((0, 1), (1, 0)):
                reach = max(reach, bands.temporal_row_reach(gb_[cur][3], gb_[cur][1], capi.camera_matrix(cams[prv]), w, h, a0_, rows[0], rows[1], torch))
            t_reach = torch.tensor([reach], dtype=torch.int64, device=dev)
            dist.all_reduce(t_reach, op=dist.ReduceOp.MAX)
            if int(t_reach.item()) > halo_rows:
                halo_rows = int(t_reach.item())
                del gb_
                a0_, a1_, rows_alloc_, gb_ = setup_band(rows, halo_rows)
        if peer:   # the context exchanges halos itself: own kernels over NVLink peer memory (restir_band_connect)
            bands.connect_neighbours(ctx, world, rank, dist, torch)
        r_ = bands.BandRenderer(ctx, h, world, rank, halo_rows, torch, dist if world > 1 else None, bounds=bounds, connected=peer)
        for s_ in (0, 1):
            ctx.bind_gbuffer(s_, *gb_[s_])
        return rows, a0_, a1_, rows_alloc_, gb_, halo_rows, r_

    def set_frame(f, c=None):
        u, lu = make_uniform_blocks(capi, capi.camera_matrix, cfg, w, h, cams, f)
        (c or ctx).set_uniforms(u)
        (c or ctx).set_lighting_uniforms(lu)

    bounds = [bands.band_rows(h, world, r)[0] for r in range(world)] + [h]
    (row_begin, row_end), a0, a1, rows_alloc, gb, halo, renderer = build_bands(bounds, HALO)
    fused = world == 1 or peer                            # restir_frame_lit: the lighting runs inside the last reuse kernel

    def render(f, image, c=None, r=None):
        """One frame into `image` (RGBA8)."""
        i = f & 1
        set_frame(f, c)
        if fused:
            (c or ctx).frame_lit(i, cfg["unbiased"], 1, image, capi.RESTIR_OUT_RGBA8_SRGB)
        else:
            (r or renderer).frame(i, cfg["unbiased"], 1)
            (c or ctx).pass_lighting(i, i, image, capi.RESTIR_OUT_RGBA8_SRGB)

    balance_note = "equal heights"
    if world > 1 and not args.no_balance:
        # bands of equal cost instead of equal height: every rank times its own kernels over two frames (events inside the
        # library, no peer waits in them), the times are gathered and the boundaries re-cut; twice, since the cost inside a
        # band is only known as its average
        scratch = torch.zeros((rows_alloc, w, 4), dtype=torch.uint8, device=dev)
        for _ in range(2):
            for f in range(2):
                render(f, scratch)
            torch.cuda.synchronize()
            ctx.profile_begin()
            for f in range(2, 4):
                render(f, scratch)
            mine = sum(ms for n_, (_, ms) in ctx.profile_end().items() if n_ != "halo_wait_kernel")   # waiting for a neighbour is not this band's cost
            gathered = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
            dist.all_gather(gathered, torch.tensor([mine], dtype=torch.float64, device=dev))
            secs = [float(g.item()) for g in gathered]
            new_bounds = bands.balanced_bounds(h, world, bounds, secs, halo)
            if new_bounds == bounds:
                break
            bounds = new_bounds
            del gb, renderer, scratch
            (row_begin, row_end), a0, a1, rows_alloc, gb, halo, renderer = build_bands(bounds, halo)
            scratch = torch.zeros((rows_alloc, w, 4), dtype=torch.uint8, device=dev)
        del scratch
        balance_note = f"equal cost, rows per band {[bounds[r + 1] - bounds[r] for r in range(world)]}"
    own_pixels = (row_end - row_begin) * w
    out_rgba8 = torch.zeros((rows_alloc, w, 4), dtype=torch.uint8, device=dev)
    flush = env.flush

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: K steps, each bracketed by events on the launching stream, L2 flushed
    # between steps (outside the timed brackets) ---------------------------------------------------------
    frame_no = 0
    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    t_w = time.perf_counter()
    for _ in range(warmup):
        render(frame_no, out_rgba8)
        frame_no += 1
    barrier()
    if not full:
        # sub-configurations: as many steps as make the timed region last ~0.6 s (so that the clock sampler sees it), >= `steps`
        est = max((time.perf_counter() - t_w) / warmup, 1e-4)
        t_steps = torch.tensor([max(steps, min(400, int(math.ceil(0.6 / est))))], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(t_steps, op=dist.ReduceOp.MAX)
        steps = int(t_steps.item())
    ctx.counters(reset=True)
    if sampler:
        sampler.wait_alive()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    if sampler:
        sampler.mark()
    for s in range(steps):
        flush.zero_()
        ev[s][0].record(stream)
        render(frame_no, out_rgba8)
        ev[s][1].record(stream)
        frame_no += 1
    barrier()
    if sampler:
        sampler.mark_end()
    clocks = sampler.stop() if sampler else None
    counters = ctx.counters(reset=True, check=False)
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([dev_ms, float(counters["shadow_rays"]), float(counters["kernel_launches"]), float(counters["halo_misses"]),
                      float(counters["stack_overflows"]), float(counters["halo_wait_timeouts"]), float(counters["shadow_rays_traced"]),
                      float(counters["shadow_rays_cached"])],
                     dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dev_ms = float(tmax[0])
        t = tsum
    rays_total, launches, halo_misses, overflows, halo_timeouts, walked_total, cached_total = (float(t[j]) for j in range(1, 8))
    ms_per_frame = dev_ms / steps
    mrays = rays_total / (dev_ms * 1e-3) / 1e6
    mrays_walked = walked_total / (dev_ms * 1e-3) / 1e6

    # ---- per-kernel breakdown (CUDA events on the launching stream inside the library, restir_profile_begin/_end),
    # per-pass brackets for the headline, roofline of the dominant kernel ------------------------------------------
    reps = max(3, min(steps, 10))
    pass_ms, pass_names = None, None
    if full:
        pass_names = ["restir pass", "unbiased pass" if cfg["unbiased"] else "spatial pass (x2)", "lighting pass"]
        pass_ms = np.zeros(3)
        for _ in range(reps):
            i = frame_no & 1
            set_frame(frame_no)
            flush.zero_()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            e[0].record(stream)
            if cfg["unbiased"]:
                ctx.pass_restir(i, capi.RESTIR_BUF_TEMP, i ^ 1)
            else:
                ctx.pass_restir(i, i, i ^ 1)
            e[1].record(stream)
            if world > 1:
                renderer._exchange(capi.RESTIR_BUF_TEMP if cfg["unbiased"] else i)
            if cfg["unbiased"]:
                ctx.pass_unbiased(i, capi.RESTIR_BUF_TEMP, i)
            else:
                ctx.pass_spatial(i, i, i ^ 1, 0)
                if world > 1:
                    renderer._exchange(i ^ 1)
                ctx.pass_spatial(i, i ^ 1, i, 1)
            e[2].record(stream)
            if world > 1:
                renderer._exchange(i)
            ctx.pass_lighting(i, i, out_rgba8, capi.RESTIR_OUT_RGBA8_SRGB)
            e[3].record(stream)
            torch.cuda.synchronize()
            pass_ms += [e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3])]
            frame_no += 1
        pass_ms /= reps
    kernel_ms, kernel_launches = {}, {}
    barrier()
    ctx.counters(reset=True)
    for _ in range(reps):
        flush.zero_()
        torch.cuda.synchronize()
        ctx.profile_begin()
        render(frame_no, out_rgba8)
        for kn, (n, ms) in ctx.profile_end().items():
            kernel_ms[kn] = kernel_ms.get(kn, 0.0) + ms / reps
            kernel_launches[kn] = kernel_launches.get(kn, 0) + n / reps
        frame_no += 1
    barrier()
    c_prof = ctx.counters(reset=True, check=False)
    rays_prof, traced_prof = c_prof["shadow_rays"] / reps, c_prof["shadow_rays_traced"] / reps
    peak, peak_src = peaks()
    info = ctx.bvh_info()
    if info["traversal"] == capi.RESTIR_TRAVERSAL_WIDE:        # 64-byte 4-wide nodes + 64-byte triangle records (csrc/wide_image.h)
        bvh_bytes = info["wide_nodes"] * 64 + info["triangles"] * 64
    elif info["traversal"] == capi.RESTIR_TRAVERSAL_IMAGE:
        bvh_bytes = info["nodes"] * 64 + info["triangles"] * 64
    else:
        bvh_bytes = scene.nodes.size + scene.triangles.size
    light_bytes = scene.point_blob.size + scene.tri_blob.size + scene.alias_blob.size
    # algorithmic bytes per launch of each kernel: every input read once, every output written once
    finalize = 16 + 16 + 4 * k + 4 * (k + 1) + (k + 1)
    alg = {
        "omni_candidates_kernel": own_pixels * (32 + 32) + light_bytes,
        "omni_temporal_kernel": own_pixels * (32 + 1 + 32 + 28 + 32 + 32) + light_bytes,
        "spatial_reuse_kernel": own_pixels * 100 + light_bytes,
        "spatial_reuse_kernel+lighting": own_pixels * (100 + 4) + light_bytes,
        "unbiased_merge_kernel": own_pixels * (32 + 32 + 32 + 4 * k + 4 * (k + 1)) + light_bytes,
        "unbiased_finalize_kernel": own_pixels * finalize,
        "unbiased_finalize_kernel+lighting": own_pixels * (finalize + 16 + 32 + 4) + light_bytes,
        "lighting_kernel": own_pixels * 68 + light_bytes,
    }
    if variant != (1, False, False):
        rb = args.reservoir_size * (64 if args.unbiased_mis else 48) + 16          # the reference's std430 record, as resident in HBM on this path
        alg.update({"generic_restir_kernel": own_pixels * (32 + 28 + 2 * rb) + light_bytes + bvh_bytes,
                    "generic_spatial_kernel": own_pixels * (36 + 2 * rb) + light_bytes,
                    "generic_unbiased_kernel": own_pixels * (36 + 2 * rb) + light_bytes + bvh_bytes,
                    "generic_lighting_kernel": own_pixels * (32 + rb + 4) + light_bytes})
    # one ray = neighbour/own position 16 B + sample position 16 B [+ neighbour index 4 B] + visibility byte; + the tree once per launch
    if "trace_kernel<pixel>" in kernel_ms:
        alg["trace_kernel<pixel>"] = own_pixels * 33 + bvh_bytes
    if "trace_kernel<own>" in kernel_ms:
        alg["trace_kernel<own>"] = own_pixels * 33 + bvh_bytes
    if "trace_kernel<neighbours>" in kernel_ms:
        # every neighbour slot is looked at (index 4 B + own visibility byte), the traced ones read two positions and write a byte
        alg["trace_kernel<neighbours>"] = own_pixels * k * 5 + max(traced_prof - 2 * own_pixels, 0) * 33 + bvh_bytes
    work = {n: v for n, v in kernel_ms.items() if not n.startswith("halo_")}
    top = max(work, key=work.get)
    top_ms = kernel_ms[top] / max(kernel_launches[top], 1)
    achieved = alg[top] / (top_ms * 1e-3) / 1e9
    traffic, warp_inst = None, {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        tdoc = json.load(open(tpath))
        traffic = tdoc.get(name, {}).get(top)
        warp_inst = tdoc.get(name + ":warp_instructions", {})
    # issue-slot utilisation per kernel: warp instructions of one launch (smsp__inst_executed.sum from the committed ncu
    # capture of this configuration) / this run's event-timed duration / (SMs x 4 schedulers x SM clock)
    sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
    sm_hz = ((clocks or {}).get("sm_mhz") or 1965.0) * 1e6
    issue_frac = {n: warp_inst[n] / (kernel_ms[n] / max(kernel_launches[n], 1) * 1e-3) / (sm_count * 4 * sm_hz)
                  for n in kernel_ms if n in warp_inst and world == 1 and scaling == "weak"}
    roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg[top], "kernel_ms": top_ms,
                "launches_per_frame": kernel_launches[top],
                "issue_frac": issue_frac.get(top),
                "note": "the trace kernel is bound by instruction issue and by the ALU pipe (66-71 % of the issue slots, ALU pipe 47-61 % against 22 % "
                        "for the FMA pipe, at 4 CTAs / SM and 12-23 of 32 lanes, ncu captures P / Q), "
                        "not by HBM: the 4-wide image and the triangle records (25 MB) are L2/L1-resident and DRAM sits below 5 % (profiles/); "
                        "its yardsticks are Mrays/s and lanes per instruction; the streaming kernels' "
                        "HBM fractions are in kernel_hbm_frac; issue_frac / kernel_issue_frac = warp instructions per launch "
                        "(profiles/traffic.json, from the ncu capture) over this run's kernel time and the SMs' issue rate (N = 1 only)"}
    kernel_hbm_frac = {n: (alg[n] * kernel_launches[n] / (kernel_ms[n] * 1e-3) / 1e9 / peak) for n in kernel_ms if alg.get(n)}
    frame_bytes = sum(alg[n] * kernel_launches[n] for n in kernel_ms if alg.get(n))
    t_fb = torch.tensor([float(frame_bytes)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_fb, op=dist.ReduceOp.SUM)
    frame_bytes_all = float(t_fb.item())
    hbm_frame_frac = frame_bytes_all / (ms_per_frame * 1e-3) / 1e9 / (peak * world)
    trace_ms = sum(v for k_, v in kernel_ms.items() if k_.startswith("trace_kernel"))
    trace_mrays_walked = traced_prof / (trace_ms * 1e-3) / 1e6 if trace_ms else None

    # ---- end to end through the C ABI: host G-buffers in (pinned), 8-bit image out, every step ---------------
    e2e = None
    if full and not args.no_e2e:
        host_gb = [[p.cpu().pin_memory() for p in planes] for planes in gb]
        # two images in flight: the device->host read of frame f runs on a side stream under frame f+1's passes, as the
        # upload of frame f+1's G-buffer (the library's copy stream) runs under frame f's reuse pass
        out_imgs = [out_rgba8, torch.zeros_like(out_rgba8)]
        host_outs = [torch.empty((rows_alloc, w, 4), dtype=torch.uint8).pin_memory() for _ in range(2)]
        side = torch.cuda.Stream()
        copy_done = [None, None]
        h2d = sum(p.numel() * p.element_size() for p in host_gb[0])
        d2h = host_outs[0].numel()

        def e2e_step(f):
            i = f & 1
            ctx.upload_gbuffer(i, *host_gb[i])          # cudaMemcpyAsync x5 on the context's copy stream
            if copy_done[i] is not None:
                stream.wait_event(copy_done[i])           # image i's previous read-back is done before it is overwritten
            render(f, out_imgs[i])
            lit = torch.cuda.Event()
            lit.record(stream)
            side.wait_event(lit)
            with torch.cuda.stream(side):
                host_outs[i].copy_(out_imgs[i], non_blocking=True)
                copy_done[i] = torch.cuda.Event()
                copy_done[i].record(side)

        def e2e_drain():
            for e_ in copy_done:
                if e_ is not None:
                    stream.wait_event(e_)

        for _ in range(3):
            e2e_step(frame_no)
            frame_no += 1
        barrier()
        ctx.counters(reset=True)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        a.record(stream)
        for _ in range(steps):
            e2e_step(frame_no)
            frame_no += 1
        e2e_drain()                                       # the last image has reached the host inside the timed region
        b.record(stream)
        barrier()
        c2 = ctx.counters(reset=True, check=False)
        tt = torch.tensor([a.elapsed_time(b), float(c2["shadow_rays"])], dtype=torch.float64, device=dev)
        if world > 1:
            mx = tt.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm_ = tt.clone()
            dist.all_reduce(sm_, op=dist.ReduceOp.SUM)
            e2e_ms, e2e_rays = float(mx[0]), float(sm_[1])
        else:
            e2e_ms, e2e_rays = float(tt[0]), float(tt[1])
        e2e = {"value": e2e_rays / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_frame": e2e_ms / steps,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)}
        # the same with the G-buffer made on the device (restir_pass_gbuffer: gBuffer.vert / gBuffer.frag by ray casting, textures
        # included): what crosses PCIe per step is the camera and the two uniform blocks in, the image out
        if fixtures.gbuffer_inputs_available(cfg["scene"]) and cam_key == cfg["scene"] and not cfg["lights"]:
            fixtures.load_gbuffer_inputs(cfg["scene"]).upload(ctx)

            def e2e_step_device(f):
                i = f & 1
                if copy_done[i] is not None:
                    stream.wait_event(copy_done[i])
                ctx.pass_gbuffer(i, cams[i])
                render(f, out_imgs[i])
                lit = torch.cuda.Event()
                lit.record(stream)
                side.wait_event(lit)
                with torch.cuda.stream(side):
                    host_outs[i].copy_(out_imgs[i], non_blocking=True)
                    copy_done[i] = torch.cuda.Event()
                    copy_done[i].record(side)

            for _ in range(3):
                e2e_step_device(frame_no)
                frame_no += 1
            barrier()
            ctx.counters(reset=True)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            a.record(stream)
            for _ in range(steps):
                e2e_step_device(frame_no)
                frame_no += 1
            e2e_drain()
            b.record(stream)
            barrier()
            c3 = ctx.counters(reset=True, check=False)
            tt = torch.tensor([a.elapsed_time(b), float(c3["shadow_rays"])], dtype=torch.float64, device=dev)
            if world > 1:
                mx = tt.clone()
                dist.all_reduce(mx, op=dist.ReduceOp.MAX)
                sm_ = tt.clone()
                dist.all_reduce(sm_, op=dist.ReduceOp.SUM)
                tt = torch.stack([mx[0], sm_[1]])
            e2e["device_gbuffer"] = {"value": float(tt[1]) / (float(tt[0]) * 1e-3) / 1e6, "unit": UNIT, "ms_per_frame": float(tt[0]) / steps,
                                     "h2d_bytes_per_step": 128 + 96 + 44, "d2h_bytes_per_step": int(d2h),
                                     "note": "every step: restir_pass_gbuffer (gBuffer.vert / gBuffer.frag by ray casting on the device, the scene's "
                                             "textures reduced to <= 256 texels a side) -> restir_frame_lit -> image to the host; no G-buffer crosses PCIe; "
                                             "the textured G-buffer is a different input from the factor-only fixture of the other legs"}
        for s in (0, 1):
            ctx.bind_gbuffer(s, *gb[s])
        del host_gb, host_outs

    # ---- N = 1: full-size parity of >= 64 rows against the CPU oracle (+ the timed CPU baseline for the headline) -----
    cpu_baseline, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and variant[:2] == (1, False):
        po = graft.load_oracle()
        # inputs of frame index 1 exactly as the GPU sees them: frame 0's final reservoirs as history
        ctx.resize(w, h)
        for s in (0, 1):
            ctx.bind_gbuffer(s, *gb[s])
        render(0, out_rgba8)
        prev64 = ctx.download_reservoirs(0)
        render(1, out_rgba8)
        gpu_final = ctx.download_reservoirs(1)
        g_host = [po.GBuffer(w, h, *[p.cpu().numpy().view(dt) for p, dt in zip(planes, (np.uint8, np.int16, np.uint16, np.float32, np.float32))])
                  for planes in gb]
        u, lu = make_uniform_blocks(po, capi.camera_matrix, cfg, w, h, cams, 1)
        inputs = dict(size=(w, h), g_cur=g_host[1], g_prev=g_host[0], prev_reservoirs=prev64.astype(po.RESERVOIR_DTYPE),
                      uniforms=u, lighting_uniforms=lu)
        sc = po.Scene(scene.nodes, scene.triangles, scene.point_blob, scene.tri_blob, scene.alias_blob)
        rows = 64
        budget = args.cpu_seconds if full else 0.0
        while True:
            y0 = max(0, h // 2 - rows // 2)
            y1 = min(h, y0 + rows)
            r = oracle_rows(po, sc, cfg, inputs, [(y0, y1)])
            if r["seconds"] >= budget * 0.5 or (y1 - y0) >= h:
                break
            rows = min(h, max(rows * 2, int(rows * budget / max(r["seconds"], 1e-3))))
        if full:
            cpu_baseline = {"value": r["rays_frame"] / r["frame_seconds"] / 1e6, "unit": UNIT, "cores": po.num_threads(), "kind": "port",
                            "sample": f"frame 2 of the same sequence, rows [{y0},{y1}) of {h} (first pass also on {HALO}-row aprons), "
                                      f"{r['seconds']:.1f} s of CPU time, each pass scaled by the rows it covered",
                            "ms_per_frame": r["frame_seconds"] * 1e3, "per_pass_ms": r["per_pass_ms"],
                            "note": "CPU restatement of the compute-shader path (OpenMP oracle) — not lavapipe"}
        # the sample is a full-size parity check of those rows
        sl = slice(y0 * w, y1 * w)
        a_, b_ = gpu_final[sl], r["final"][sl]
        same = np.ones(a_.shape[0], bool)
        for fld in ("lightIndex", "M"):
            same &= a_[fld] == b_[fld]
        for fld in ("position_emissionLum", "normal", "pHat", "sumWeights", "w"):
            x, y = np.ascontiguousarray(a_[fld]).view(np.uint32), np.ascontiguousarray(b_[fld]).view(np.uint32)
            eq = (x == y) | (np.isnan(a_[fld]) & np.isnan(b_[fld]))
            same &= eq.reshape(a_.shape[0], -1).all(axis=1)
        parity = {"rows": [int(y0), int(y1)], "pixels": int(a_.shape[0]), "mismatching_reservoirs": int((~same).sum()),
                  "against": "oracle/restir_oracle.cpp (CPU), frame 2 with the GPU's frame 1 as history, every field bit for bit"}

    # ---- N > 1: the connected bands against a whole-screen context, bit for bit ---------------------------------
    band_parity = None
    if world > 1 and peer:
        frames = 3
        barrier()
        for b in range(3):                                  # a common starting point: empty history everywhere, halo rows included
            bands.reservoir_rows_tensor(ctx, b, torch).zero_()
        barrier()
        ctx.counters(reset=True, check=False)
        for f in range(frames):
            render(f, out_rgba8)
        ctx.synchronize()
        cb = ctx.counters(reset=True, check=False)
        whole = new_context()
        whole.resize(w, h)
        wgb = render_gbuffers(whole, h)
        for s in (0, 1):
            whole.bind_gbuffer(s, *wgb[s])
        whole_img = torch.zeros((h, w, 4), dtype=torch.uint8, device=dev)
        for f in range(frames):
            i = f & 1
            set_frame(f, whole)
            whole.frame_lit(i, cfg["unbiased"], 1, whole_img, capi.RESTIR_OUT_RGBA8_SRGB)
        whole.synchronize()
        last = (frames - 1) & 1
        bad = bands.mismatching_owned_reservoirs(ctx, whole, last, torch)
        bad_px = int((out_rgba8[row_begin - a0: row_end - a0] != whole_img[row_begin: row_end]).any(dim=-1).sum().item())
        tb = torch.tensor([float(own_pixels), float(bad), float(bad_px), float(cb["halo_misses"]), float(cb["halo_wait_timeouts"])],
                          dtype=torch.float64, device=dev)
        dist.all_reduce(tb, op=dist.ReduceOp.SUM)
        band_parity = {"frames": frames, "pixels": int(tb[0]), "mismatching_reservoirs": int(tb[1]), "mismatching_pixels_rgba8": int(tb[2]),
                       "halo_misses": int(tb[3]), "halo_wait_timeouts": int(tb[4]),
                       "against": "a single context rendering the whole screen on every rank's own GPU; each rank compares the packed "
                                  "reservoirs and RGBA8 pixels of the rows it owns on the device"}
        whole.close()
        del wgb, whole_img
        barrier()

    invalid = []
    if halo_misses:
        invalid.append(f"{int(halo_misses)} halo misses")
    if halo_timeouts:
        invalid.append(f"{int(halo_timeouts)} halo wait timeouts")
    if overflows:
        invalid.append(f"{int(overflows)} traversal stack overflows")
    if band_parity and (band_parity["mismatching_reservoirs"] or band_parity["mismatching_pixels_rgba8"] or band_parity["halo_misses"]
                        or band_parity["halo_wait_timeouts"]):
        invalid.append(f"band parity: {band_parity['mismatching_reservoirs']} reservoirs / {band_parity['mismatching_pixels_rgba8']} pixels differ")
    if parity and parity["mismatching_reservoirs"]:
        invalid.append(f"parity sample: {parity['mismatching_reservoirs']} reservoirs differ from the oracle")

    barrier()
    ctx.close()
    del gb, renderer, out_rgba8
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    reuse = f"unbiased reuse, {k} neighbours" if cfg["unbiased"] else f"biased reuse 2 passes x {k} neighbours"
    rec = {
        "value": mrays, "value_walked": mrays_walked, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_per_frame, "ms_per_frame": ms_per_frame, "scaling": scaling,
        "config": {"workload": f"{name}: {label}, {w}x{h} frame ({w}x{band_h} band per GPU), {cfg['candidates']} candidates, {reuse}, "
                               f"temporal reuse on, software shadow rays, lights: {scene.light_counts()}",
                   "l2": "L2 flushed (256 MiB memset) between timed steps, outside the event brackets",
                   "reservoir_layout": "32-byte packed", "gbuffer_bytes_per_pixel": 36,
                   "frame": "restir_frame_lit (lighting fused into the last reuse kernel)" if fused else "restir passes + lighting pass",
                   "variant": {"reservoir_size": variant[0], "unbiased_mis": variant[1], "fused_passes": variant[2]},
                   "parallelism": f"row-bands x{world} ({balance_note}), halo exchange: {'own kernels over NVLink peer memory' if args.halo == 'peer' else 'NCCL send/recv'}, "
                                  f"halo {halo} rows (spatial reach {HALO}, temporal reprojection reach measured on the host)"},
        "rays_per_frame": rays_total / steps, "rays_walked_per_frame": walked_total / steps,
        "rays_answered_by_cached_occluder_per_frame": cached_total / steps,
        "kernel_ms": kernel_ms, "kernel_launches_per_frame": kernel_launches, "kernel_hbm_frac": kernel_hbm_frac,
        "trace_kernel_mrays_walked_per_s": trace_mrays_walked,
        "hbm_frame_frac": hbm_frame_frac, "frame_algorithmic_bytes": frame_bytes_all,
        "gpu_launches": int(launches), "halo_misses": int(halo_misses), "halo_wait_timeouts": int(halo_timeouts), "stack_overflows": int(overflows),
        "clocks": clocks, "parity_sample": parity, "band_parity": band_parity, "invalid": invalid,
    }
    if full:
        rec.update({"pass_ms": dict(zip(pass_names, [float(x) for x in pass_ms])), "kernel_issue_frac": issue_frac, "bvh": info,
                    "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu_baseline})
    else:
        rec["roofline"] = {k_: roofline[k_] for k_ in ("kernel", "achieved", "peak", "unit", "frac", "kernel_ms")}
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default=HEADLINE, choices=sorted(CONFIGS))
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--traversal", default="auto", choices=["auto", "image", "reference-order"],
                    help="reference-order: walk the 80-byte nodes literally (A/B against the 64-byte re-stride)")
    ap.add_argument("--ray-elision", default="on", choices=["on", "off", "dedupe"],
                    help="unbiased pass: answer neighbour rays without a walk where that is exact (restir_set_ray_elision; A/B)")
    ap.add_argument("--occluder-cache", default="on", choices=["on", "off", "light", "direction"],
                    help="trace kernel: test the cached occluder of (screen region, light) before queueing a ray for a walk (exact; A/B)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default, the driver's contract): one config-sized band per GPU; strong: the config's frame split into N bands")
    ap.add_argument("--halo", default="peer", choices=["peer", "nccl"],
                    help="N > 1: halo rows pushed by the library's own kernels into the neighbours' memory (default), or NCCL send/recv "
                         "between the passes (bands.exchange_halo)")
    ap.add_argument("--spatial-staging", default="off", choices=["off", "on"],
                    help="biased spatial pass: gate data staged in shared memory (restir_set_spatial_staging) or read directly (A/B)")
    ap.add_argument("--reservoir-size", type=int, default=1, choices=[1, 2, 4], help="RESERVOIR_SIZE (restir_set_reservoir_variant); N = 1 only")
    ap.add_argument("--unbiased-mis", action="store_true", help="UNBIASED_MIS (restir_set_reservoir_variant); N = 1 only")
    ap.add_argument("--fused-passes", action="store_true",
                    help="one kernel per reference shader, rays traced inline (restir_set_reservoir_variant(..., fused = 1)): the A/B of the tuned "
                         "path's cut passes; N = 1 only")
    ap.add_argument("--no-balance", action="store_true", help="N > 1: keep bands of equal height instead of equal measured cost")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skips the CPU baseline and the oracle parity samples")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-suite", action="store_true", help="only the headline configuration: no `configs` block")
    ap.add_argument("--suite-steps", type=int, default=20, help="timed steps (at least) of every configuration in the `configs` block")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            run_reference_arm(args)
        return 0

    # fd 1 carries exactly one JSON line: everything libraries print meanwhile (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    json_out = os.fdopen(json_fd, "w")

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the ReSTIR passes have no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))

    pkg = graft.load_package()
    env = Env()
    env.torch, env.dist, env.capi, env.fixtures = torch, dist, pkg.capi, pkg.fixtures
    env.bands = __import__("restir_vulkan_b200.bands", fromlist=["bands"])
    env.rank, env.world, env.local_rank, env.args = rank, world, local_rank, args
    env.dev = f"cuda:{local_rank}"
    env.stream = torch.cuda.Stream()
    torch.cuda.set_stream(env.stream)                      # NCCL p2p ops order against the current stream
    env.flush = torch.empty(256 << 20, dtype=torch.uint8, device=env.dev)   # > 126 MB L2

    head = run_gpu_config(env, args.config, args.scaling, args.steps, args.warmup, full=True)
    suite = {}
    if not args.no_suite and args.config == HEADLINE and args.scaling == "weak":
        if world == 1:
            for name in SUITE_N1:
                suite[name] = run_gpu_config(env, name, "weak", args.suite_steps, 3, full=False)
        suite["strong_8k"] = run_gpu_config(env, STRONG, "strong", args.suite_steps, 3, full=False)

    rc = 0
    if rank == 0:
        invalid = list(head["invalid"]) + [f"{n}: {why}" for n, r in suite.items() for why in r["invalid"]]
        line = {"metric": METRIC, "higher_is_better": True, "vs_baseline": None, "dtype": "f32", "data": "synthetic"}
        line.update(head)
        line["metric_note"] = ("value = the reference's testVisibility calls answered per second over the whole frame time (reference-equivalent "
                               "visibility queries, the unit of work the --impl reference arm shares); value_walked = the rays that needed a walk "
                               "of the tree, the rest are answered exactly without one (DESIGN.md §4)")
        line["configs"] = suite
        line["invalid"] = invalid
        if invalid:                                         # never a full-credit number on a frame that is not the reference's frame
            line["value_measured_but_rejected"] = line["value"]
            line["value"] = 0.0
            if line.get("e2e"):
                line["e2e"]["value"] = 0.0
            rc = 1
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
    if world > 1:
        t_rc = torch.tensor([rc], dtype=torch.int64, device=env.dev)
        dist.broadcast(t_rc, 0)
        rc = int(t_rc.item())
        dist.destroy_process_group()
    return rc


if __name__ == "__main__":
    sys.exit(main())
