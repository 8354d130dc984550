"""Row-band data parallelism for the ReSTIR passes: one process per GPU, contiguous bands of rows, scene
replicated, nearest-neighbour halo exchange of reservoir rows.

The reference is single-GPU; this is the one subsystem the B200 build adds.  Pixels are independent except
for (i) the spatial / unbiased neighbour gathers within spatialRadius (+1 for round()) rows and (ii) the
temporal gather at the reprojected pixel, so the only communication is: before every pass that reads
another pixel's reservoir, each rank sends the `halo` rows adjacent to each band edge of that pass's INPUT
buffer to the rank on the other side of the edge.  RNG streams are keyed on global pixel coordinates, so
the result is bit-identical to a single-GPU frame.

`torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is plumbing for the point-to-point copies; there
is no collective on the data path.
"""
import math


def band_rows(height, world, rank, bounds=None):
    """Contiguous row bands: rank r owns [begin, end).  Near-equal heights, or `bounds` (world + 1 ascending row
    indices from 0 to height, e.g. from balanced_bounds)."""
    if bounds is not None:
        assert len(bounds) == world + 1 and bounds[0] == 0 and bounds[-1] == height
        return int(bounds[rank]), int(bounds[rank + 1])
    base, extra = divmod(height, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def balanced_bounds(height, world, bounds, seconds, min_rows):
    """Band boundaries of equal COST instead of equal height.

    `seconds[r]` is what rank r needed for its band [bounds[r], bounds[r+1]) (the frame time is the slowest band's:
    sky rows make no candidates and no rays, floor rows trace the longest ones).  Cost is taken as uniform inside each
    measured band (a piecewise-constant density over the rows) and the new boundaries cut its integral into `world`
    equal parts, every band at least `min_rows` high (the halo must fit into a neighbour's band, halo_plan).
    Deterministic: every rank computes the same boundaries from the same gathered times.
    """
    assert len(bounds) == world + 1 and len(seconds) == world
    density = []
    for r in range(world):
        rows = bounds[r + 1] - bounds[r]
        density.extend([max(float(seconds[r]), 1e-9) / rows] * rows)
    total = sum(density)
    new, acc, y = [0], 0.0, 0
    for r in range(1, world):
        target = total * r / world
        while y < height and acc + density[y] <= target:
            acc += density[y]
            y += 1
        lo = new[-1] + min_rows
        hi = height - (world - r) * min_rows
        new.append(min(max(y, lo), hi))
    new.append(height)
    return new


def halo_rows_for(spatial_radius):
    """Rows a neighbour gather can reach: floor/round of radius*sin() is at most ceil(radius), +1 for round-to-even slack."""
    return int(math.ceil(spatial_radius)) + 1


def temporal_row_reach(world_pos, normal, prev_pv, width, height, alloc_begin, row_begin, row_end, torch):
    """How many rows the temporal reprojection of restirOmni.glsl:163-171 moves this band's pixels, at most.

    world_pos: (rows, W, 4) float32 and normal: (rows, W, 4) int16 device planes covering rows [alloc_begin, ...);
    prev_pv: the 16 floats of prevFrameProjectionViewMatrix (column-major).  Only the band's own rows and only
    surface pixels count (a background pixel never passes the normal gate, restir_kernels.cu omni_temporal_kernel);
    pixels that reproject outside the screen are not looked up by the shader.  The host sizes the halo from the
    largest reach over all ranks (+1 row of slack for the float evaluation order, which is not the kernel's).
    """
    wp = world_pos[row_begin - alloc_begin: row_end - alloc_begin, :, :3].to(torch.float32)
    surf = (normal[row_begin - alloc_begin: row_end - alloc_begin, :, :3] != 0).any(dim=-1)
    m = [float(v) for v in prev_pv]
    x, y, z = wp[..., 0], wp[..., 1], wp[..., 2]
    px = m[0] * x + m[4] * y + m[8] * z + m[12]
    py = m[1] * x + m[5] * y + m[9] * z + m[13]
    pw = m[3] * x + m[7] * y + m[11] * z + m[15]
    sx = (px / pw + 1.0) * 0.5 * width
    sy = (py / pw + 1.0) * 0.5 * height
    inside = (sx > 0) & (sy > 0) & (sx < width) & (sy < height) & surf
    rows = torch.arange(row_begin, row_end, device=wp.device, dtype=torch.float32)[:, None]
    reach = torch.where(inside, (torch.trunc(sy) - rows).abs(), torch.zeros_like(sy))
    return int(reach.max().item()) + 1 if reach.numel() else 0


def halo_plan(height, world, rank, halo, bounds=None):
    """Messages of one exchange for `rank`: list of (peer, send_rows, recv_rows) with rows as global [lo, hi).

    A rank sends the first/last `halo` rows it OWNS and receives the `halo` rows just outside its band, each
    clipped to the peer's band (bands thinner than the halo would need a second hop: rejected).
    """
    begin, end = band_rows(height, world, rank, bounds)
    plan = []
    for peer in (rank - 1, rank + 1):
        if peer < 0 or peer >= world:
            continue
        pb, pe = band_rows(height, world, peer, bounds)
        if pe - pb < halo or end - begin < halo:
            raise ValueError(f"band of {min(pe - pb, end - begin)} rows is thinner than the {halo}-row halo")
        if peer < rank:
            plan.append((peer, (begin, begin + halo), (begin - halo, begin)))
        else:
            plan.append((peer, (end - halo, end), (end, end + halo)))
    return plan


def exchange_halo(rows_tensor, alloc_begin, plan, dist, group=None):
    """Run one halo exchange on `rows_tensor` (first dim = rows [alloc_begin, ...), any trailing dims).

    Uses batched isend/irecv; works for CPU tensors over gloo and CUDA tensors over NCCL.  Returns after the
    transfers are complete with respect to the caller's stream (NCCL) / the host (gloo).
    """
    ops, recvs = [], []
    for peer, (s0, s1), (r0, r1) in plan:
        send = rows_tensor[s0 - alloc_begin: s1 - alloc_begin].contiguous()
        recv = rows_tensor[r0 - alloc_begin: r1 - alloc_begin]
        tmp = recv if recv.is_contiguous() else recv.contiguous()
        ops.append(dist.P2POp(dist.isend, send, peer, group))
        ops.append(dist.P2POp(dist.irecv, tmp, peer, group))
        recvs.append((recv, tmp))
    if not ops:
        return
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    for recv, tmp in recvs:
        if tmp is not recv:
            recv.copy_(tmp)


def connect_neighbours(ctx, world, rank, dist, torch, group=None):
    """One process per GPU: hands every rank's reservoir buffers to its two neighbours as CUDA IPC mappings and
    connects them (restir_band_export_ipc / _open_ipc / _connect).  From then on the context exchanges halos itself
    with its own kernels over NVLink peer memory and BandRenderer issues no exchange.  torch.distributed only carries
    the 272-byte handles, once."""
    everyone = [None] * world
    dist.all_gather_object(everyone, ctx.band_export_ipc(), group=group)   # any backend: NCCL between GPUs, gloo in the one-GPU test
    for side, peer_rank in ((0, rank - 1), (1, rank + 1)):
        if 0 <= peer_rank < world:
            ctx.band_connect(side, ctx.band_open_ipc(everyone[peer_rank]))
        else:
            ctx.band_connect(side, None)
    dist.barrier(group=group)   # nobody starts pushing before everybody has connected


class _DeviceBytes:
    """Zero-copy view of raw device memory for torch.as_tensor (CUDA array interface v2)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def reservoir_rows_tensor(ctx, buffer, torch):
    """The context's packed reservoir buffer `buffer` as a (rows, pitch_bytes) uint8 CUDA tensor (no copy)."""
    ptr, pitch = ctx.reservoir_device_ptr(buffer)
    rows = ctx.alloc_rows()
    t = torch.as_tensor(_DeviceBytes(ptr, rows * pitch), device=f"cuda:{ctx.device}")
    return t.view(rows, pitch)


class BandRenderer:
    """Drives one band context through App's pass order (src/app.h:212-262) with the halo exchanges between
    the passes.  With world == 1 it degenerates to restir_frame + lighting."""

    def __init__(self, ctx, height, world, rank, halo, torch, dist=None, group=None, bounds=None, connected=False):
        """connected: the context's neighbours are connected (connect_neighbours): it pushes and waits for halos itself."""
        self.ctx, self.world, self.rank, self.torch, self.dist, self.group = ctx, world, rank, torch, dist, group
        self.height, self.halo = height, halo
        self.plan = halo_plan(height, world, rank, halo, bounds) if world > 1 and not connected else []
        if world > 1:
            halo_plan(height, world, rank, halo, bounds)   # still validates that the halo fits into the neighbours' bands
        self.alloc_begin = ctx.band()[2]
        self._views = {}

    def _exchange(self, buffer):
        if not self.plan:
            return
        if buffer not in self._views:
            self._views[buffer] = reservoir_rows_tensor(self.ctx, buffer, self.torch)
        exchange_halo(self._views[buffer], self.alloc_begin, self.plan, self.dist, self.group)

    def frame(self, i, unbiased, spatial_iterations=1):
        """One frame's resampling passes.  Leaves the final reservoirs (halo included) in FRAME[i]."""
        from .capi import RESTIR_BUF_TEMP

        ctx, p = self.ctx, i ^ 1
        if unbiased:
            ctx.pass_restir(i, RESTIR_BUF_TEMP, p)
            self._exchange(RESTIR_BUF_TEMP)
            ctx.pass_unbiased(i, RESTIR_BUF_TEMP, i)
        else:
            ctx.pass_restir(i, i, p)
            for j in range(spatial_iterations):
                self._exchange(i)
                ctx.pass_spatial(i, i, p, 2 * j)
                self._exchange(p)
                ctx.pass_spatial(i, p, i, 2 * j + 1)
        self._exchange(i)   # next frame's temporal reprojection reads FRAME[i] within the halo


def mismatching_owned_reservoirs(band_ctx, whole_ctx, buffer, torch):
    """How many of the reservoirs this band OWNS in `buffer` differ, in any bit of their packed 32 bytes, from the same
    pixels of a context that rendered the whole screen (same device).  Compared on the device."""
    row_begin, row_end, alloc_begin, _ = band_ctx.band()
    mine = reservoir_rows_tensor(band_ctx, buffer, torch)[row_begin - alloc_begin: row_end - alloc_begin]
    whole = reservoir_rows_tensor(whole_ctx, buffer, torch)[row_begin: row_end]
    differs = (mine.view(mine.shape[0], -1, 32) != whole.view(whole.shape[0], -1, 32)).any(dim=-1)
    return int(differs.sum().item())

