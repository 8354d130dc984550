"""One rank of the multi-process row-band test (tests/test_gpu_properties.py::test_band_processes_over_cuda_ipc).

    RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT in the environment (torchrun's contract), argv: out.json unbiased

Every rank is its own process with its own CUDA context — as under `torchrun bench.py --gpus N` — but all of them sit on
cuda:0 (CUDA IPC works between processes on one device), so the path restir_band_export_ipc -> restir_band_open_ipc ->
restir_band_connect -> halo_push_kernel / halo_wait_kernel across PROCESSES is covered by the one-GPU test tier.
torch.distributed (gloo) only carries the IPC handles.  Each rank renders its band of three frames through the
connected contexts, renders the whole screen in a second, single context, and compares the reservoirs it owns bit
for bit on the device.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    out_path, unbiased = sys.argv[1], sys.argv[2] == "1"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    import torch
    import torch.distributed as dist

    import parity_harness as ph
    from test_gpu_properties import DeviceFrames, _scene_full

    capi = ph.capi
    bands = __import__("restir_vulkan_b200.bands", fromlist=["bands"])
    dist.init_process_group("gloo")
    scene, (pos, look) = _scene_full()
    w, h, halo, frames = 1920, 432, 64, 3
    bounds = [0, 150, 290, 432][: world] + [432] if world == 3 else None
    band = bands.band_rows(h, world, rank, bounds)
    part = DeviceFrames(scene, pos, look, w, h, band=band, halo=halo)
    bands.connect_neighbours(part.ctx, world, rank, dist, torch)
    image = torch.zeros((part.a1 - part.a0, w, 4), dtype=torch.uint8, device="cuda")
    for f in range(frames):
        part.set(f)
        part.ctx.frame(f & 1, unbiased, 1)
        part.ctx.pass_lighting(f & 1, f & 1, image, capi.RESTIR_OUT_RGBA8_SRGB)
    part.ctx.synchronize()
    counters = part.ctx.counters(check=False)

    whole = DeviceFrames(scene, pos, look, w, h)
    whole_image = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    for f in range(frames):
        whole.set(f)
        whole.ctx.frame(f & 1, unbiased, 1)
        whole.ctx.pass_lighting(f & 1, f & 1, whole_image, capi.RESTIR_OUT_RGBA8_SRGB)
    whole.ctx.synchronize()
    last = (frames - 1) & 1
    bad = bands.mismatching_owned_reservoirs(part.ctx, whole.ctx, last, torch)
    bad_px = int((image[part.rb - part.a0: part.re - part.a0] != whole_image[part.rb: part.re]).any(dim=-1).sum().item())
    dist.barrier()   # nobody unmaps a neighbour's buffers while it may still be pushed to
    json.dump({"rank": rank, "rows": [part.rb, part.re], "pixels": (part.re - part.rb) * w, "mismatching_reservoirs": bad,
               "mismatching_pixels": bad_px, "halo_misses": int(counters["halo_misses"]),
               "halo_wait_timeouts": int(counters["halo_wait_timeouts"]), "pid": os.getpid()}, open(out_path, "w"))
    part.ctx.close()
    whole.ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
