#!/usr/bin/env python
"""bench.py — ReSTIR ms/frame + shadow Mrays/s on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config NAME]

A "step" is one frame of the hot path: restirOmni -> (unbiased reuse | 2 x spatial reuse) -> lighting, on
synthetic inputs of the configuration's shape (G-buffers ray-cast from the scene by the fixture tool, two
camera positions alternating so temporal reprojection does real work).  N = 1 runs the configuration
BASELINE.json's metric is quoted on: Sponza 1920x1080, 32 candidates, 200 random point lights, unbiased
reuse with the north-star's 5 spatial neighbours.  N > 1 (torchrun, one rank per GPU) gives every rank a
1080-row band of a 1920 x (1080 N) frame (weak scaling) with halo exchange between the passes.

One JSON line on stdout (rank 0).  `value` = shadow Mrays/s over the whole frame time with inputs resident
in HBM; `ms_per_step` = ms/frame; `e2e` = the same through the C ABI with host G-buffers in and the 8-bit
image out; `roofline` for the dominant kernel; `cpu_baseline` = the CPU oracle on a bounded row sample.
"""
import argparse
import json
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
    "cornell_720p_biased4": dict(scene="cornellBox", size=(1280, 720), unbiased=False, neighbors=4, candidates=32, lights=None),
    "office_2160p_unbiased3": dict(scene="office", size=(3840, 2160), unbiased=True, neighbors=3, candidates=32, lights=None),
    "sponza_8k_1m_lights": dict(scene="sponza", size=(7680, 4320), unbiased=True, neighbors=5, candidates=64, lights=1_000_000),
}
CAMERAS = {
    "sponza": ((3.0, 4.0, 5.0), (0.0, 0.0, 0.0)),       # src/camera.h:7-13 defaults
    "cornellBox": ((3.0, 4.0, 5.0), (0.0, 0.0, 0.0)),
    "office": ((3.0, 1.7, 0.5), (3.0, 1.5, -5.0)),      # inside the room
    "procedural": ((3.0, 3.5, 4.2), (0.0, -1.0, 0.0)),
}
HALO = 31  # ceil(spatialRadius = 30) + 1

# algorithmic bytes per pixel with the layout actually resident in HBM (32-byte packed reservoirs,
# 36-byte G-buffer) — DESIGN.md §Kernels, SURVEY.md §8d
BYTES_PER_PIXEL = {"spatial": 100, "lighting_rgba8": 68, "lighting_rgba32f": 80}


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
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines, self.first = index, None, [], 0

    def mark(self):
        """The timed region starts here: nvidia-smi needs a moment to produce its first line, so it is started during
        the warm-up and only the samples from this point on are used."""
        self.first = len(self.lines)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in (self.lines[self.first:] or self.lines):
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
# this image) on all host cores, on a bounded row sample of the same workload

def oracle_rows_sample(po, scene, cfg, cams, frame_inputs, target_seconds, rows_hint=None, passes=None):
    """Time the CPU passes on a centred band of rows of frame index 1 (temporal active).

    passes: the module whose passes are timed — the oracle port (default) or oracle/pyglslref.py, the reference's
    own shader sources compiled for the CPU (then the rays are counted by the oracle afterwards, untimed: the two
    are bit-identical, tests/test_oracle_vs_glsl.py).

    frame_inputs: dict(g_cur, g_prev, prev_reservoirs (64-byte, full frame), uniforms, lighting_uniforms).
    Returns dict with per-pass seconds scaled to the full frame, rays, rows, and the sampled reservoirs.
    """
    w, h = cfg["size"]
    sc = po.Scene(scene.nodes, scene.triangles, scene.point_blob, scene.tri_blob, scene.alias_blob)
    u, lu = frame_inputs["uniforms"], frame_inputs["lighting_uniforms"]
    rows = rows_hint or 8
    result = None
    while True:
        y0 = max(0, h // 2 - rows // 2)
        y1 = min(h, y0 + rows)
        a0, a1 = max(0, y0 - HALO), min(h, y1 + HALO)
        pm = passes or po
        t0 = time.perf_counter()
        initial, rays_a = pm.restir_pass(sc, u, frame_inputs["g_cur"], frame_inputs["g_prev"], frame_inputs["prev_reservoirs"], (a0, a1))
        t1 = time.perf_counter()
        if cfg["unbiased"]:
            final, rays_b = pm.unbiased_pass(sc, u, frame_inputs["g_cur"], initial, cfg["neighbors"], (y0, y1))
        else:
            mid = pm.spatial_pass(u, frame_inputs["g_cur"], initial, 0, (a0, a1))
            final = pm.spatial_pass(u, frame_inputs["g_cur"], mid, 1, (y0, y1))
            rays_b = 0
        t2 = time.perf_counter()
        pm.lighting_pass(sc, lu, frame_inputs["g_cur"], final, (y0, y1))
        t3 = time.perf_counter()
        if pm is not po:                                   # count the rays of the same rows, untimed
            _, rays_a = po.restir_pass(sc, u, frame_inputs["g_cur"], frame_inputs["g_prev"], frame_inputs["prev_reservoirs"], (a0, a1))
            if cfg["unbiased"]:
                _, rays_b = po.unbiased_pass(sc, u, frame_inputs["g_cur"], initial, cfg["neighbors"], (y0, y1))
        # scale each pass by the rows it actually covered
        t_restir = (t1 - t0) * h / (a1 - a0)
        if cfg["unbiased"]:
            t_reuse = (t2 - t1) * h / (y1 - y0)
        else:
            t_reuse = (t2 - t1) * h / ((a1 - a0) + (y1 - y0)) * 2.0
        t_light = (t3 - t2) * h / (y1 - y0)
        rays_frame = rays_a * h / (a1 - a0) + rays_b * h / (y1 - y0)
        result = dict(seconds_sample=t3 - t0, frame_seconds=t_restir + t_reuse + t_light, rays_frame=rays_frame, rows=(y0, y1),
                      apron_rows=(a0, a1), final=final, initial=initial,
                      per_pass_ms=dict(restir=t_restir * 1e3, reuse=t_reuse * 1e3, lighting=t_light * 1e3))
        if (t3 - t0) >= target_seconds * 0.5 or (y1 - y0) >= h or rows_hint:
            return result
        rows = min(h, max(rows * 2, int(rows * target_seconds / max(t3 - t0, 1e-3))))


def make_uniform_blocks(capi, po_or_capi_matrix, cfg, w, h, cams, f):
    cam, prev_cam = cams[f & 1], cams[(f & 1) ^ 1] if f > 0 else cams[0]
    pv_prev = po_or_capi_matrix(prev_cam)
    u = capi.make_uniforms(prevFrameProjectionViewMatrix=pv_prev, cameraPos=(cam.position[0], cam.position[1], cam.position[2], 1.0),
                           screenSize=(w, h), frame=f + 1, initialLightSampleCount=cfg["candidates"], temporalSampleCountMultiplier=20,
                           spatialPosThreshold=0.1, spatialNormalThreshold=25.0, spatialNeighbors=cfg["neighbors"], spatialRadius=30.0,
                           flags=3)   # src/app.h:150-174, app.cpp:414-434 defaults
    lu = capi.make_lighting_uniforms(prevFrameProjectionViewMatrix=pv_prev, cameraPos=(cam.position[0], cam.position[1], cam.position[2], 1.0),
                                     bufferSize=(w, h), debugMode=0, gamma=1.0)
    return u, lu


def load_glsl_reference(cfg):
    """oracle/_ref/libglslref.so (the reference's own shader sources compiled for the CPU, oracle/ref_build), or None
    where it was never built or was not compiled for this neighbour count."""
    import importlib.util
    graft.load_oracle()
    spec = importlib.util.spec_from_file_location("pyglslref", os.path.join(ROOT, "oracle", "pyglslref.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules.setdefault("pyglslref", mod)
    spec.loader.exec_module(mod)
    if not mod.available() or (cfg["unbiased"] and cfg["neighbors"] not in (3, 5)):
        return None
    return mod


def run_reference_arm(args):
    """--impl reference: the reference's CPU arm alone, no CUDA code on the path — its own shader sources compiled
    for the host (oracle/_ref/libglslref.so) when that was built, else the oracle port."""
    pkg = graft.load_package()
    capi, fixtures = pkg.capi, pkg.fixtures
    po = graft.load_oracle()
    po.set_num_threads(os.cpu_count() or 1)               # torchrun exports OMP_NUM_THREADS=1; this arm uses every host core
    gl = load_glsl_reference(CONFIGS[args.config])
    cfg = CONFIGS[args.config]
    scene, cam_key, label = load_scene(fixtures, cfg)
    w, h = cfg["size"]
    pos, look = CAMERAS[cam_key]
    cams = [po.make_camera(position=pos, look_at=look, aspect=w / h),
            po.make_camera(position=(pos[0] + 0.05, pos[1], pos[2]), look_at=look, aspect=w / h)]
    sc = po.Scene(scene.nodes, scene.triangles, scene.point_blob, scene.tri_blob, scene.alias_blob)
    rows = args.reference_rows
    y0 = max(0, h // 2 - rows // 2)
    a0, a1 = max(0, y0 - 2 * HALO), min(h, y0 + rows + 2 * HALO)
    table = scene.material_table()
    gbufs = [po.raycast_gbuffer(sc, scene.tri_material, table, c, w, h, (a0, a1)) for c in cams]
    prev = np.zeros(w * h, po.RESERVOIR_DTYPE)
    times, rays_total = [], []
    for step in range(args.warmup + args.steps):
        f = step
        u, lu = make_uniform_blocks(capi, po.camera_matrix, cfg, w, h, cams, f)
        inputs = dict(g_cur=gbufs[f & 1], g_prev=gbufs[(f & 1) ^ 1] if f > 0 else None, prev_reservoirs=prev,
                      uniforms=u.astype(po.UNIFORMS_DTYPE), lighting_uniforms=lu.astype(po.LIGHTING_UNIFORMS_DTYPE))
        r = oracle_rows_sample(po, scene, cfg, cams, inputs, 0.0, rows_hint=rows, passes=gl)
        prev = r["final"]    # valid on the sampled rows, which is where the next step reprojects to
        if step >= args.warmup:
            times.append(r["frame_seconds"])
            rays_total.append(r["rays_frame"])
    ms = float(np.mean(times)) * 1e3
    mrays = float(np.mean(rays_total)) / (ms * 1e-3) / 1e6
    cores = po.num_threads()
    sample = f"rows [{y0},{y0 + rows}) of {h} (+{HALO}-row aprons for the first pass), each pass scaled by rows covered"
    line = {
        "impl": "reference", "metric": METRIC, "value": mrays, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "ms_per_frame": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": f"{args.config}: {label} {w}x{h}", "note": (
            "the reference's own shader sources (restirOmni.glsl, unbiasedReuse.glsl / spatialReuse.comp, lighting.frag) compiled for "
            "the host by g++ with OpenMP (oracle/_ref/libglslref.so)" if gl else "CPU restatement of the compute-shader path "
            "(oracle/restir_oracle.cpp, OpenMP)") + " — not lavapipe: no Vulkan in this image"},
        "cpu_baseline": {"value": mrays, "unit": UNIT, "cores": cores, "kind": "reference" if gl else "port", "sample": sample, "ms_per_frame": ms},
        "e2e": {"value": mrays, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="sponza_1080p_unbiased5", choices=sorted(CONFIGS))
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--reference-rows", type=int, default=48, help="rows per step of the --impl reference arm")
    ap.add_argument("--traversal", default="auto", choices=["auto", "reference-order"],
                    help="reference-order: walk the 80-byte nodes literally (A/B against the 64-byte re-stride)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default, the driver's contract): one config-sized band per GPU; strong: the config's frame split into N bands")
    ap.add_argument("--halo", default="peer", choices=["peer", "nccl"],
                    help="N > 1: halo rows pushed by the library's own kernels into the neighbours' memory (default), or NCCL send/recv "
                         "between the passes (bands.exchange_halo)")
    ap.add_argument("--no-balance", action="store_true", help="N > 1: keep bands of equal height instead of equal measured cost")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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
    capi, fixtures, bands = pkg.capi, pkg.fixtures, __import__("restir_vulkan_b200.bands", fromlist=["bands"])
    cfg = CONFIGS[args.config]
    scene, cam_key, label = load_scene(fixtures, cfg)
    w, band_h = cfg["size"]
    if args.scaling == "strong":                          # the configuration's own frame, split into N row bands
        h, band_h = band_h, (band_h + world - 1) // world
    else:
        h = band_h * world                                # weak scaling: one config-sized band per rank
    row_begin, row_end = bands.band_rows(h, world, rank)
    pos, look = CAMERAS[cam_key]
    cams = [capi.make_camera(position=pos, look_at=look, aspect=w / h),
            capi.make_camera(position=(pos[0] + 0.05, pos[1], pos[2]), look_at=look, aspect=w / h)]

    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)                          # NCCL p2p ops order against the current stream
    ctx = capi.RestirContext(local_rank, stream.cuda_stream)
    if args.traversal == "reference-order":
        ctx.set_traversal(capi.RESTIR_TRAVERSAL_REFERENCE_ORDER)
    ctx.upload_bvh(scene.nodes, scene.triangles)
    ctx.upload_lights(scene.point_blob, scene.tri_blob, scene.alias_blob)
    ctx.set_unbiased_neighbors(cfg["neighbors"] if cfg["unbiased"] else 3)
    dev = f"cuda:{local_rank}"
    tm = torch.from_numpy(np.ascontiguousarray(scene.tri_material)).to(dev)
    mt = torch.from_numpy(scene.material_table().view(np.int32)).to(dev)

    def setup_band(rows, halo_rows):
        """Allocates the band `rows` with `halo_rows` of halo and renders the two synthetic G-buffers (one per camera)
        on the device with the fixture tool."""
        if world > 1:
            ctx.resize_band(w, h, rows[0], rows[1], halo_rows)
        else:
            ctx.resize(w, h)
        _, _, a0_, a1_ = ctx.band()
        rows_ = a1_ - a0_
        gb_ = []
        for c in cams:
            planes = [torch.zeros((rows_, w, 4), dtype=torch.uint8, device=dev), torch.zeros((rows_, w, 4), dtype=torch.int16, device=dev),
                      torch.zeros((rows_, w, 2), dtype=torch.int16, device=dev), torch.zeros((rows_, w, 4), dtype=torch.float32, device=dev),
                      torch.zeros((rows_, w), dtype=torch.float32, device=dev)]
            ctx.raycast_gbuffer(c, tm, mt, *planes)
            gb_.append(planes)
        ctx.synchronize()
        return a0_, a1_, rows_, gb_

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
        peer = world > 1 and args.halo == "peer"
        if peer:   # the context exchanges halos itself: own kernels over NVLink peer memory (restir_band_connect)
            bands.connect_neighbours(ctx, world, rank, dist, torch)
        r_ = bands.BandRenderer(ctx, h, world, rank, halo_rows, torch, dist if world > 1 else None, bounds=bounds, connected=peer)
        for s_ in (0, 1):
            ctx.bind_gbuffer(s_, *gb_[s_])
        return rows, a0_, a1_, rows_alloc_, gb_, halo_rows, r_

    def set_frame(f):
        u, lu = make_uniform_blocks(capi, capi.camera_matrix, cfg, w, h, cams, f)
        ctx.set_uniforms(u)
        ctx.set_lighting_uniforms(lu)

    bounds = [bands.band_rows(h, world, r)[0] for r in range(world)] + [h]
    (row_begin, row_end), a0, a1, rows_alloc, gb, halo, renderer = build_bands(bounds, HALO)
    balance_note = "equal heights"
    if world > 1 and not args.no_balance:
        # bands of equal cost instead of equal height: every rank times its own kernels over two frames (events inside the
        # library, no peer waits in them), the times are gathered and the boundaries re-cut; twice, since the cost inside a
        # band is only known as its average
        for _ in range(2):
            for f in range(2):
                set_frame(f)
                renderer.frame(f & 1, cfg["unbiased"], 1)
            torch.cuda.synchronize()
            ctx.profile_begin()
            for f in range(2, 4):
                set_frame(f)
                renderer.frame(f & 1, cfg["unbiased"], 1)
            mine = sum(ms for name, (_, ms) in ctx.profile_end().items() if name != "halo_wait_kernel")   # waiting for a neighbour is not this band's cost
            gathered = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
            dist.all_gather(gathered, torch.tensor([mine], dtype=torch.float64, device=dev))
            secs = [float(g.item()) for g in gathered]
            new_bounds = bands.balanced_bounds(h, world, bounds, secs, halo)
            if new_bounds == bounds:
                break
            bounds = new_bounds
            del gb, renderer
            (row_begin, row_end), a0, a1, rows_alloc, gb, halo, renderer = build_bands(bounds, halo)
        balance_note = f"equal cost, rows per band {[bounds[r + 1] - bounds[r] for r in range(world)]}"
    own_pixels = (row_end - row_begin) * w
    out_rgba8 = torch.zeros((rows_alloc, w, 4), dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def device_step(f):
        i = f & 1
        set_frame(f)
        renderer.frame(i, cfg["unbiased"], 1)
        ctx.pass_lighting(i, i, out_rgba8, capi.RESTIR_OUT_RGBA8_SRGB)

    for s in (0, 1):
        ctx.bind_gbuffer(s, *gb[s])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: K steps, each bracketed by events on the launching stream, L2 flushed
    # between steps (outside the timed brackets) ---------------------------------------------------------
    frame_no = 0
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        device_step(frame_no)
        frame_no += 1
    barrier()
    ctx.counters(reset=True)
    sampler.mark()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for k in range(args.steps):
        flush.zero_()
        ev[k][0].record(stream)
        device_step(frame_no)
        ev[k][1].record(stream)
        frame_no += 1
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    counters = ctx.counters(reset=True)
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([dev_ms, float(counters["shadow_rays"]), float(counters["kernel_launches"]), float(counters["halo_misses"]),
                      float(counters["stack_overflows"]), float(counters["halo_wait_timeouts"])], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dev_ms, rays_total, launches = float(tmax[0]), float(tsum[1]), float(tsum[2])
        halo_misses, overflows, halo_timeouts = float(tsum[3]), float(tsum[4]), float(tsum[5])
    else:
        rays_total, launches, halo_misses, overflows, halo_timeouts = float(t[1]), float(t[2]), float(t[3]), float(t[4]), float(t[5])
    ms_per_frame = dev_ms / args.steps
    mrays = rays_total / (dev_ms * 1e-3) / 1e6

    # ---- per-pass and per-kernel breakdown (CUDA events on the launching stream: pass brackets here, kernel
    # brackets inside the library via restir_profile_begin/_end) + roofline of the dominant kernel ------------
    names = ["restir pass", "unbiased pass" if cfg["unbiased"] else "spatial pass (x2)", "lighting pass"]
    pass_ms = np.zeros(3)
    reps = max(3, min(args.steps, 10))
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
    ctx.counters(reset=True)
    for _ in range(reps):
        flush.zero_()
        torch.cuda.synchronize()
        ctx.profile_begin()
        device_step(frame_no)
        for name, (n, ms) in ctx.profile_end().items():
            kernel_ms[name] = kernel_ms.get(name, 0.0) + ms / reps
            kernel_launches[name] = kernel_launches.get(name, 0) + n / reps
        frame_no += 1
    c_prof = ctx.counters(reset=True)
    rays_prof, traced_prof = c_prof["shadow_rays"] / reps, c_prof["shadow_rays_traced"] / reps
    peak, peak_src = peaks()
    info = ctx.bvh_info()
    bvh_bytes = (info["nodes"] * 64 if info["traversal"] == capi.RESTIR_TRAVERSAL_IMAGE else scene.nodes.size) + scene.triangles.size
    light_bytes = scene.point_blob.size + scene.tri_blob.size + scene.alias_blob.size
    k = cfg["neighbors"]
    # algorithmic bytes per launch of each kernel: every input read once, every output written once
    alg = {
        "omni_candidates_kernel": own_pixels * (32 + 32) + light_bytes,
        "omni_temporal_kernel": own_pixels * (32 + 1 + 32 + 28 + 32 + 32) + light_bytes,
        "spatial_reuse_kernel": own_pixels * BYTES_PER_PIXEL["spatial"] + light_bytes,
        "unbiased_merge_kernel": own_pixels * (32 + 32 + 32 + 4 * k) + light_bytes,
        "unbiased_finalize_kernel": own_pixels * (16 + 16 + 16 + 4 * k + (k + 1)),
        "lighting_kernel": own_pixels * BYTES_PER_PIXEL["lighting_rgba8"] + light_bytes,
    }
    # one ray = neighbour/own position 16 B + sample position 16 B [+ neighbour index 4 B] + visibility byte; + the tree once per launch
    rays_pixel = own_pixels if "trace_kernel<pixel>" in kernel_ms else 0
    if "trace_kernel<pixel>" in kernel_ms:
        alg["trace_kernel<pixel>"] = rays_pixel * 33 + bvh_bytes
    if "trace_kernel<own>" in kernel_ms:
        alg["trace_kernel<own>"] = own_pixels * 33 + bvh_bytes
    if "trace_kernel<neighbours>" in kernel_ms:
        # every neighbour slot is looked at (index 4 B + own visibility byte), the traced ones read two positions and write a byte
        alg["trace_kernel<neighbours>"] = own_pixels * k * 5 + max(traced_prof - rays_pixel - own_pixels, 0) * 33 + bvh_bytes
    top = max(kernel_ms, key=kernel_ms.get)
    top_ms = kernel_ms[top] / max(kernel_launches[top], 1)
    achieved = alg[top] / (top_ms * 1e-3) / 1e9
    traffic, warp_inst = None, {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        tdoc = json.load(open(tpath))
        traffic = tdoc.get(args.config, {}).get(top)
        warp_inst = tdoc.get(args.config + ":warp_instructions", {})
    # issue-slot utilisation per kernel: warp instructions of one launch (smsp__inst_executed.sum from the committed ncu
    # capture of this configuration) / this run's event-timed duration / (SMs x 4 schedulers x SM clock)
    sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
    sm_hz = ((clocks or {}).get("sm_mhz") or 1965.0) * 1e6
    issue_frac = {n: warp_inst[n] / (kernel_ms[n] / max(kernel_launches[n], 1) * 1e-3) / (sm_count * 4 * sm_hz)
                  for n in kernel_ms if n in warp_inst and world == 1}
    roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg[top], "kernel_ms": top_ms,
                "launches_per_frame": kernel_launches[top],
                "issue_frac": issue_frac.get(top),
                "note": "the trace kernel is bound by L1 tag lookups (83 % l1tex) and instruction issue (71 % issue-active) together, not by HBM: "
                        "the tree is L2/L1-resident and DRAM sits below 2 % (profiles/r1_m_summary.md); its yardsticks are Mrays/s and lanes per "
                        "instruction; the streaming kernels' HBM fractions are in kernel_hbm_frac; issue_frac / kernel_issue_frac = warp "
                        "instructions per launch (profiles/traffic.json, from the ncu capture) over this run's kernel time and the SMs' "
                        "issue rate (N = 1 only)"}
    kernel_hbm_frac = {n: (alg[n] * kernel_launches[n] / (kernel_ms[n] * 1e-3) / 1e9 / peak) for n in kernel_ms if alg.get(n)}
    frame_bytes = sum(alg[n] * kernel_launches[n] for n in kernel_ms if alg.get(n))
    hbm_frame_frac = frame_bytes / (ms_per_frame * 1e-3) / 1e9 / peak
    trace_ms = sum(v for k_, v in kernel_ms.items() if k_.startswith("trace_kernel"))
    trace_mrays = rays_prof / (trace_ms * 1e-3) / 1e6 if trace_ms else None
    trace_mrays_walked = traced_prof / (trace_ms * 1e-3) / 1e6 if trace_ms else None

    # ---- end to end through the C ABI: host G-buffers in (pinned), 8-bit image out, every step ---------------
    e2e = None
    if not args.no_e2e:
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
            set_frame(f)
            renderer.frame(i, cfg["unbiased"], 1)
            if copy_done[i] is not None:
                stream.wait_event(copy_done[i])           # image i's previous read-back is done before it is overwritten
            ctx.pass_lighting(i, i, out_imgs[i], capi.RESTIR_OUT_RGBA8_SRGB)
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
        for _ in range(args.steps):
            e2e_step(frame_no)
            frame_no += 1
        e2e_drain()                                       # the last image has reached the host inside the timed region
        b.record(stream)
        barrier()
        c2 = ctx.counters(reset=True)
        tt = torch.tensor([a.elapsed_time(b), float(c2["shadow_rays"])], dtype=torch.float64, device=dev)
        if world > 1:
            mx = tt.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm_ = tt.clone()
            dist.all_reduce(sm_, op=dist.ReduceOp.SUM)
            e2e_ms, e2e_rays = float(mx[0]), float(sm_[1])
        else:
            e2e_ms, e2e_rays = float(tt[0]), float(tt[1])
        e2e = {"value": e2e_rays / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_frame": e2e_ms / args.steps,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)}
        for s in (0, 1):
            ctx.bind_gbuffer(s, *gb[s])

    # ---- CPU baseline (rank 0, N = 1): the oracle on a bounded row sample of the same frame ------------------
    cpu_baseline, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        po = graft.load_oracle()
        f = 1
        # inputs of frame index 1 exactly as the GPU sees them: frame 0's final reservoirs as history
        ctx.resize(w, h)
        for s in (0, 1):
            ctx.bind_gbuffer(s, *gb[s])
        set_frame(0)
        renderer.frame(0, cfg["unbiased"], 1)
        prev64 = ctx.download_reservoirs(0)
        set_frame(1)
        renderer.frame(1, cfg["unbiased"], 1)
        gpu_final = ctx.download_reservoirs(1)
        g_host = [po.GBuffer(w, h, *[p.cpu().numpy().view(dt) for p, dt in zip(planes, (np.uint8, np.int16, np.uint16, np.float32, np.float32))])
                  for planes in gb]
        u, lu = make_uniform_blocks(capi, capi.camera_matrix, cfg, w, h, cams, f)
        inputs = dict(g_cur=g_host[1], g_prev=g_host[0], prev_reservoirs=prev64.astype(po.RESERVOIR_DTYPE),
                      uniforms=u.astype(po.UNIFORMS_DTYPE), lighting_uniforms=lu.astype(po.LIGHTING_UNIFORMS_DTYPE))
        r = oracle_rows_sample(po, scene, cfg, cams, inputs, args.cpu_seconds)
        y0, y1 = r["rows"]
        cpu_ms = r["frame_seconds"] * 1e3
        cpu_baseline = {"value": r["rays_frame"] / r["frame_seconds"] / 1e6, "unit": UNIT, "cores": po.num_threads(), "kind": "port",
                        "sample": f"frame 2 of the same sequence, rows [{y0},{y1}) of {h} (first pass also on {HALO}-row aprons), "
                                  f"{r['seconds_sample']:.1f} s of CPU time, each pass scaled by the rows it covered",
                        "ms_per_frame": cpu_ms, "per_pass_ms": r["per_pass_ms"],
                        "note": "CPU restatement of the compute-shader path (OpenMP oracle) — not lavapipe"}
        # the sample doubles as a full-size parity check of those rows
        sl = slice(y0 * w, y1 * w)
        a_, b_ = gpu_final[sl], r["final"][sl]
        same = np.ones(a_.shape[0], bool)
        for fld in ("lightIndex", "M"):
            same &= a_[fld] == b_[fld]
        for fld in ("position_emissionLum", "normal", "pHat", "sumWeights", "w"):
            x, y = np.ascontiguousarray(a_[fld]).view(np.uint32), np.ascontiguousarray(b_[fld]).view(np.uint32)
            eq = (x == y) | (np.isnan(a_[fld]) & np.isnan(b_[fld]))
            same &= eq.reshape(a_.shape[0], -1).all(axis=1)
        parity = {"rows": [int(y0), int(y1)], "pixels": int(a_.shape[0]), "mismatching_reservoirs": int((~same).sum())}

    if rank == 0:
        line = {
            "metric": METRIC, "value": mrays, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_frame, "ms_per_frame": ms_per_frame, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.config}: {label}, {w}x{h} frame ({w}x{band_h} band per GPU), {cfg['candidates']} candidates, "
                                   f"{'unbiased reuse, ' + str(cfg['neighbors']) + ' neighbours' if cfg['unbiased'] else 'biased reuse 2 passes x ' + str(cfg['neighbors']) + ' neighbours'}, "
                                   f"temporal reuse on, software shadow rays, lights: {scene.light_counts()}",
                       "l2": "L2 flushed (256 MiB memset) between timed steps, outside the event brackets",
                       "reservoir_layout": "32-byte packed", "gbuffer_bytes_per_pixel": 36, "parallelism": f"row-bands x{world} ({balance_note}), halo exchange: {'own kernels over NVLink peer memory' if args.halo == 'peer' else 'NCCL send/recv'}, halo {halo} rows (spatial reach {HALO}, temporal reprojection reach measured on the host)"},
            "rays_per_frame": rays_total / args.steps, "rays_walked_per_frame": traced_prof,
            "rays_note": "value counts the reference's testVisibility calls answered per second (the same unit of work as the --impl reference "
                         "arm); rays_walked_per_frame of them needed a walk of the tree, the rest are answered exactly without one",
            "pass_ms": dict(zip(names, [float(x) for x in pass_ms])),
            "kernel_ms": kernel_ms, "kernel_launches_per_frame": kernel_launches, "kernel_hbm_frac": kernel_hbm_frac, "kernel_issue_frac": issue_frac,
            "trace_kernel_mrays_per_s": trace_mrays, "trace_kernel_mrays_walked_per_s": trace_mrays_walked, "bvh": info,
            "hbm_frame_frac": hbm_frame_frac, "frame_algorithmic_bytes": frame_bytes,
            "gpu_launches": int(launches), "halo_misses": int(halo_misses), "halo_wait_timeouts": int(halo_timeouts), "stack_overflows": int(overflows),
            "clocks": clocks, "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu_baseline, "parity_sample": parity,
        }
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
