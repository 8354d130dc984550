"""Frame captures (include/restir_capture.h): write and read the hot path's inputs and, optionally, the outputs a
reference run produced.  numpy only; used by tests and by tools that prepare replays for host/restir_driver --replay."""
import struct

import numpy as np

from .capi import LIGHTING_UNIFORMS_DTYPE, RESERVOIR_DTYPE, UNIFORMS_DTYPE

MAGIC = b"RSTRCAP1"
HAS_INITIAL, HAS_FINAL, HAS_RGBA = 1, 2, 4
_HEADER = struct.Struct("<8s8I2I3Q")   # magic, version, w, h, frames, unbiased, neighbours, iterations, expected, [pad-free] n_nodes, n_tris, 3 x u64
PLANES = (("albedo", np.uint8, 4), ("normal", np.int16, 4), ("material", np.uint16, 2), ("world_pos", np.float32, 4), ("depth", np.float32, 1))


class Capture:
    """scene: object with nodes (n x 80 uint8), triangles (n x 48 uint8), point_blob, tri_blob, alias_blob (uint8 arrays).
    frames: list of dicts with uniforms, lighting_uniforms, planes (5 arrays) and optionally initial, final (RESERVOIR_DTYPE),
    rgba (h x w x 4 float32)."""

    def __init__(self, width, height, unbiased, unbiased_neighbors, spatial_iterations, nodes, triangles, point_blob, tri_blob, alias_blob, frames):
        self.width, self.height = int(width), int(height)
        self.unbiased, self.unbiased_neighbors, self.spatial_iterations = bool(unbiased), int(unbiased_neighbors), int(spatial_iterations)
        self.nodes = np.ascontiguousarray(nodes, np.uint8).reshape(-1, 80)
        self.triangles = np.ascontiguousarray(triangles, np.uint8).reshape(-1, 48)
        self.point_blob, self.tri_blob, self.alias_blob = (np.ascontiguousarray(b, np.uint8).reshape(-1) for b in (point_blob, tri_blob, alias_blob))
        self.frames = frames

    def expected_bits(self):
        f = self.frames[0]
        return (HAS_INITIAL if f.get("initial") is not None else 0) | (HAS_FINAL if f.get("final") is not None else 0) | \
               (HAS_RGBA if f.get("rgba") is not None else 0)

    def write(self, path):
        bits = self.expected_bits()
        n = self.width * self.height
        with open(path, "wb") as out:
            out.write(_HEADER.pack(MAGIC, 1, self.width, self.height, len(self.frames), int(self.unbiased), self.unbiased_neighbors,
                                   self.spatial_iterations, bits, self.nodes.shape[0], self.triangles.shape[0], self.point_blob.size,
                                   self.tri_blob.size, self.alias_blob.size))
            for a in (self.nodes, self.triangles, self.point_blob, self.tri_blob, self.alias_blob):
                out.write(a.tobytes())
            for f in self.frames:
                out.write(np.ascontiguousarray(f["uniforms"]).astype(UNIFORMS_DTYPE).tobytes())
                out.write(np.ascontiguousarray(f["lighting_uniforms"]).astype(LIGHTING_UNIFORMS_DTYPE).tobytes())
                for (name, dt, ch), plane in zip(PLANES, f["planes"]):
                    p = np.ascontiguousarray(plane).view(dt).reshape(-1)
                    assert p.size == n * ch, f"{name}: {p.size} values for {n} pixels x {ch}"
                    out.write(p.tobytes())
                for key, bit in (("initial", HAS_INITIAL), ("final", HAS_FINAL)):
                    if bits & bit:
                        r = np.ascontiguousarray(f[key]).astype(RESERVOIR_DTYPE)
                        assert r.size == n
                        out.write(r.tobytes())
                if bits & HAS_RGBA:
                    out.write(np.ascontiguousarray(f["rgba"], np.float32).reshape(n, 4).tobytes())

    @staticmethod
    def read(path):
        with open(path, "rb") as src:
            data = src.read()
        magic, version, w, h, frames, unbiased, neighbours, iterations, bits, n_nodes, n_tris, pb, tb, ab = _HEADER.unpack_from(data, 0)
        if magic != MAGIC or version != 1:
            raise ValueError(f"{path}: not a RSTRCAP1 version-1 capture")
        off = _HEADER.size

        def take(count, dtype):
            nonlocal off
            a = np.frombuffer(data, dtype, count, off).copy()   # own, aligned, writable memory (consumers assume 16-byte alignment)
            off += a.nbytes
            return a

        nodes = take(n_nodes * 80, np.uint8).reshape(-1, 80)
        tris = take(n_tris * 48, np.uint8).reshape(-1, 48)
        point, tri, alias = take(pb, np.uint8), take(tb, np.uint8), take(ab, np.uint8)
        n = w * h
        out = []
        for _ in range(frames):
            f = {"uniforms": take(1, UNIFORMS_DTYPE)[0], "lighting_uniforms": take(1, LIGHTING_UNIFORMS_DTYPE)[0]}
            planes = []
            for name, dt, ch in PLANES:
                p = take(n * ch, dt)
                planes.append(p.reshape(h, w, ch) if ch > 1 else p.reshape(h, w))
            f["planes"] = planes
            f["initial"] = take(n, RESERVOIR_DTYPE) if bits & HAS_INITIAL else None
            f["final"] = take(n, RESERVOIR_DTYPE) if bits & HAS_FINAL else None
            f["rgba"] = take(n * 4, np.float32).reshape(h, w, 4) if bits & HAS_RGBA else None
            out.append(f)
        if off != len(data):
            raise ValueError(f"{path}: {len(data) - off} trailing bytes")
        return Capture(w, h, unbiased, neighbours, iterations, nodes, tris, point, tri, alias, out)
