"""ctypes binding of librestir_b200.so — the C ABI declared in include/restir_b200.h.

This is the thin host-side mirror the tests and bench.py drive; every compute call goes to a CUDA
kernel in the shared library.  There is no CPU fallback: if the library is missing or no GPU is
visible, construction fails loudly.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RESTIR_B200_LIB") or os.path.join(_HERE, "librestir_b200.so")  # the env override is for A/B experiment builds

RESTIR_BUF_FRAME0, RESTIR_BUF_FRAME1, RESTIR_BUF_TEMP = 0, 1, 2
RESTIR_OUT_RGBA32F, RESTIR_OUT_RGBA8_SRGB = 0, 1
RESTIR_VISIBILITY_REUSE_FLAG, RESTIR_TEMPORAL_REUSE_FLAG = 1, 2
RESTIR_TRAVERSAL_AUTO, RESTIR_TRAVERSAL_REFERENCE_ORDER, RESTIR_TRAVERSAL_IMAGE, RESTIR_TRAVERSAL_WIDE = 0, 1, 2, 3
RESTIR_E_INVALID, RESTIR_E_CUDA, RESTIR_E_NOMEM, RESTIR_E_UNSUPPORTED, RESTIR_E_HALO = -1, -2, -3, -4, -5

RESERVOIR_DTYPE = np.dtype(
    [
        ("position_emissionLum", "<f4", (4,)),
        ("normal", "<f4", (4,)),
        ("lightIndex", "<i4"),
        ("pHat", "<f4"),
        ("sumWeights", "<f4"),
        ("w", "<f4"),
        ("M", "<u4"),
        ("_pad", "<u4", (3,)),
    ]
)
UNIFORMS_DTYPE = np.dtype(
    [
        ("prevFrameProjectionViewMatrix", "<f4", (16,)),
        ("cameraPos", "<f4", (4,)),
        ("screenSize", "<u4", (2,)),
        ("frame", "<u4"),
        ("initialLightSampleCount", "<u4"),
        ("temporalSampleCountMultiplier", "<u4"),
        ("spatialPosThreshold", "<f4"),
        ("spatialNormalThreshold", "<f4"),
        ("spatialNeighbors", "<u4"),
        ("spatialRadius", "<f4"),
        ("flags", "<i4"),
        ("_pad", "<u4", (2,)),
    ]
)
LIGHTING_UNIFORMS_DTYPE = np.dtype(
    [
        ("prevFrameProjectionViewMatrix", "<f4", (16,)),
        ("cameraPos", "<f4", (4,)),
        ("bufferSize", "<u4", (2,)),
        ("debugMode", "<i4"),
        ("gamma", "<f4"),
    ]
)
assert RESERVOIR_DTYPE.itemsize == 64 and UNIFORMS_DTYPE.itemsize == 128 and LIGHTING_UNIFORMS_DTYPE.itemsize == 96

# every symbol include/restir_b200.h declares (tests check the library exports all of them)
EXPORTS = [
    "restir_create", "restir_destroy", "restir_last_error", "restir_synchronize", "restir_upload_bvh", "restir_build_bvh_device",
    "restir_upload_lights", "restir_resize", "restir_resize_band", "restir_get_band", "restir_bind_gbuffer",
    "restir_upload_gbuffer", "restir_upload_geometry", "restir_upload_materials", "restir_pass_gbuffer",
    "restir_gbuffer_device_planes", "restir_import_external_memory", "restir_release_external_memory", "restir_set_uniforms", "restir_set_lighting_uniforms", "restir_set_unbiased_neighbors", "restir_set_reservoir_variant",
    "restir_get_reservoir_bytes",
    "restir_set_traversal", "restir_set_occluder_cache", "restir_set_ray_elision", "restir_set_spatial_staging", "restir_get_bvh_info", "restir_check_aabb_tree", "restir_check_wide_walk", "restir_get_wide_image", "restir_profile_begin", "restir_profile_end",
    "restir_pass_restir", "restir_pass_spatial", "restir_pass_unbiased", "restir_pass_lighting", "restir_frame", "restir_frame_lit",
    "restir_download_reservoirs", "restir_upload_reservoirs", "restir_reservoir_device_ptr", "restir_trace_segments",
    "restir_get_counters", "restir_build_aabb_tree", "restir_build_aabb_tree_mt", "restir_collect_triangle_lights",
    "restir_generate_random_point_lights", "restir_create_alias_table", "restir_camera_matrix",
    "restir_tools_raycast_gbuffer", "restir_band_local_peer", "restir_band_export_ipc",
    "restir_band_open_ipc", "restir_band_connect", "restir_band_balanced_bounds",
]


class BandPeer(C.Structure):
    """restir_band_peer: a neighbour's reservoir buffers and counter block, addressable from this process."""
    _fields_ = [("reservoirs", C.c_void_p * 3), ("flags", C.c_void_p), ("alloc_begin", C.c_uint32), ("alloc_end", C.c_uint32),
                ("row_begin", C.c_uint32), ("row_end", C.c_uint32)]


class BandIpc(C.Structure):
    """restir_band_ipc: the same as CUDA IPC handles, to be shipped to the neighbour's process."""
    _fields_ = [("reservoirs", (C.c_ubyte * 64) * 3), ("flags", C.c_ubyte * 64), ("alloc_begin", C.c_uint32), ("alloc_end", C.c_uint32),
                ("row_begin", C.c_uint32), ("row_end", C.c_uint32)]


class GBufferPlanes(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("albedo", "normal", "material", "worldPos", "depth")]


class Texture(C.Structure):
    _fields_ = [("rgba8", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32)]


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("shadow_rays", "stack_overflows", "halo_misses", "kernel_launches", "halo_wait_timeouts",
                                            "shadow_rays_traced", "shadow_rays_cached")]


class BvhInfo(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("nodes", "triangles", "reachable_nodes", "depth", "reference_stack_bound")] + [("traversal", C.c_int32)] + [
        (n, C.c_uint32) for n in ("wide_nodes", "wide_depth", "wide_stack_bound")]


class KernelTime(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("launches", C.c_uint32), ("total_ms", C.c_float)]


class Camera(C.Structure):
    _fields_ = [
        ("position", C.c_float * 3),
        ("lookAt", C.c_float * 3),
        ("worldUp", C.c_float * 3),
        ("zNear", C.c_float),
        ("zFar", C.c_float),
        ("fovYRadians", C.c_float),
        ("aspectRatio", C.c_float),
    ]


def make_camera(position=(3.0, 4.0, 5.0), look_at=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), z_near=0.01, z_far=1000.0,
                fov_y=None, aspect=1.0):
    """Defaults = reference src/camera.h:7-13."""
    cam = Camera()
    cam.position[:] = position
    cam.lookAt[:] = look_at
    cam.worldUp[:] = up
    cam.zNear, cam.zFar = z_near, z_far
    cam.fovYRadians = float(np.float32(0.5) * np.float32(np.pi)) if fov_y is None else fov_y
    cam.aspectRatio = aspect
    return cam


_lib = None


def load_library():
    """Load librestir_b200.so; raises if it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`; "
                               "there is no CPU fallback for the ReSTIR passes")
        lib = C.CDLL(LIB_PATH)
        lib.restir_last_error.restype = C.c_char_p
        lib.restir_last_error.argtypes = [C.c_void_p]
        lib.restir_destroy.restype = None
        lib.restir_destroy.argtypes = [C.c_void_p]
        lib.restir_collect_triangle_lights.restype = C.c_int64
        _lib = lib
    return _lib


def _hp(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _dp(t):
    """Device pointer of a torch tensor / int / None."""
    if t is None:
        return None
    if isinstance(t, int):
        return C.c_void_p(t)
    return C.c_void_p(t.data_ptr())


class RestirError(RuntimeError):
    pass


# ---- host-side scene builders (no GPU needed) ----------------------------------------------------

def build_aabb_tree(triangles):
    """AabbTree::build (src/aabbTreeBuilder.cpp:52-214): triangles (T,48)u8 / (T,12)f32 -> nodes (T-1,80)u8."""
    tris = np.ascontiguousarray(triangles).view(np.uint8).reshape(-1, 48)
    nodes = np.zeros((tris.shape[0] - 1, 80), np.uint8)
    rc = load_library().restir_build_aabb_tree(_hp(tris), C.c_uint32(tris.shape[0]), _hp(nodes))
    if rc != 0:
        raise RestirError(f"restir_build_aabb_tree failed ({rc})")
    return nodes


def build_aabb_tree_mt(triangles, threads=0):
    """restir_build_aabb_tree_mt: the same bytes as build_aabb_tree, one breadth-first level at a time on `threads` host threads."""
    tris = np.ascontiguousarray(triangles).view(np.uint8).reshape(-1, 48)
    nodes = np.zeros((tris.shape[0] - 1, 80), np.uint8)
    rc = load_library().restir_build_aabb_tree_mt(_hp(tris), C.c_uint32(tris.shape[0]), _hp(nodes), C.c_uint32(threads))
    if rc != 0:
        raise RestirError(f"restir_build_aabb_tree_mt failed ({rc})")
    return nodes


def band_balanced_bounds(height, bounds, seconds, min_rows):
    """restir_band_balanced_bounds (host): the C++ twin of bands.balanced_bounds."""
    n = len(bounds) - 1
    b_in = (C.c_uint32 * (n + 1))(*[int(b) for b in bounds])
    secs = (C.c_double * n)(*[float(x) for x in seconds])
    out = (C.c_uint32 * (n + 1))()
    rc = load_library().restir_band_balanced_bounds(C.c_uint32(height), C.c_uint32(n), b_in, secs, C.c_uint32(min_rows), out)
    if rc != 0:
        raise RestirError(f"restir_band_balanced_bounds failed ({rc})")
    return [int(v) for v in out]


def check_aabb_tree(nodes, n_triangles):
    """restir_check_aabb_tree: (rc, info dict, message).  Host only."""
    nodes = np.ascontiguousarray(nodes).view(np.uint8).reshape(-1, 80)
    info, msg = BvhInfo(), C.create_string_buffer(256)
    rc = load_library().restir_check_aabb_tree(_hp(nodes), C.c_uint32(nodes.shape[0]), C.c_uint32(n_triangles), C.byref(info), msg, C.c_size_t(256))
    return rc, {n: getattr(info, n) for n, _ in BvhInfo._fields_}, msg.value.decode()


def check_wide_walk(nodes, triangles, p1, p2):
    """restir_check_wide_walk: (rc, shadowed, walked_wide, visits, message).  Host only."""
    nodes = np.ascontiguousarray(nodes).view(np.uint8).reshape(-1, 80)
    tris = np.ascontiguousarray(triangles).view(np.uint8).reshape(-1, 48)
    p1 = np.ascontiguousarray(p1, np.float32).reshape(-1, 3)
    p2 = np.ascontiguousarray(p2, np.float32).reshape(-1, 3)
    n = p1.shape[0]
    shadowed, walked = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
    visits, msg = C.c_uint64(0), C.create_string_buffer(256)
    rc = load_library().restir_check_wide_walk(_hp(nodes), C.c_uint32(nodes.shape[0]), _hp(tris), C.c_uint32(tris.shape[0]), _hp(p1), _hp(p2), C.c_uint64(n),
                                               _hp(shadowed), _hp(walked), C.byref(visits), msg, C.c_size_t(256))
    return rc, shadowed, walked, int(visits.value), msg.value.decode()


def collect_triangle_lights(triangles, tri_material, material_emissive):
    tris = np.ascontiguousarray(triangles).view(np.uint8).reshape(-1, 48)
    tri_material = np.ascontiguousarray(tri_material, np.int32)
    em = np.ascontiguousarray(material_emissive, np.float32).reshape(-1, 3)
    out = np.zeros((tris.shape[0], 80), np.uint8)
    n = load_library().restir_collect_triangle_lights(_hp(tris), _hp(tri_material), C.c_uint32(tris.shape[0]), _hp(em),
                                                      C.c_uint32(em.shape[0]), _hp(out))
    if n < 0:
        raise RestirError(f"restir_collect_triangle_lights failed ({n})")
    return out[:n].copy()


def generate_random_point_lights(count, lo, hi):
    lo = np.ascontiguousarray(lo, np.float32)
    hi = np.ascontiguousarray(hi, np.float32)
    out = np.zeros((count, 32), np.uint8)
    rc = load_library().restir_generate_random_point_lights(C.c_uint64(count), _hp(lo), _hp(hi), _hp(out))
    if rc != 0:
        raise RestirError(f"restir_generate_random_point_lights failed ({rc})")
    return out


def create_alias_table(point_lights, tri_lights):
    pl = np.ascontiguousarray(point_lights).view(np.uint8).reshape(-1, 32)
    tl = np.ascontiguousarray(tri_lights).view(np.uint8).reshape(-1, 80)
    n = pl.shape[0] if pl.shape[0] else tl.shape[0]
    out = np.zeros((n, 16), np.uint8)
    rc = load_library().restir_create_alias_table(_hp(pl) if pl.size else None, C.c_uint64(pl.shape[0]),
                                                  _hp(tl) if tl.size else None, C.c_uint64(tl.shape[0]), _hp(out))
    if rc != 0:
        raise RestirError(f"restir_create_alias_table failed ({rc})")
    return out


def make_blob(items, stride):
    """{int32 count; pad to 16; array} — src/sceneBuffers.h:100-124, 241-270."""
    items = np.ascontiguousarray(items).view(np.uint8).reshape(-1, stride) if np.size(items) else np.zeros((0, stride), np.uint8)
    blob = np.zeros(16 + items.size, np.uint8)
    blob[:4] = np.array([items.shape[0]], np.int32).view(np.uint8)
    blob[16:] = items.reshape(-1)
    return blob


def camera_matrix(cam):
    out = np.zeros(16, np.float32)
    rc = load_library().restir_camera_matrix(C.byref(cam), _hp(out))
    if rc != 0:
        raise RestirError("restir_camera_matrix failed")
    return out


def make_uniforms(**kw):
    u = np.zeros((), UNIFORMS_DTYPE)
    for k, v in kw.items():
        u[k] = v
    return u


def make_lighting_uniforms(**kw):
    u = np.zeros((), LIGHTING_UNIFORMS_DTYPE)
    for k, v in kw.items():
        u[k] = v
    return u


# ---- the context -----------------------------------------------------------------------------------

class RestirContext:
    """Owns one restir_context on one GPU.  Mirrors the resources App binds to the four passes."""

    def __init__(self, device=0, stream=None):
        self.lib = load_library()
        self._ctx = C.c_void_p()
        rc = self.lib.restir_create(C.byref(self._ctx), C.c_int(device), _dp(stream))
        if rc != 0:
            self._ctx = None
            raise RestirError(f"restir_create(device={device}) failed ({rc}): a CUDA GPU is required, there is no CPU fallback")
        self.device = device
        self.width = self.height = 0

    def close(self):
        if getattr(self, "_ctx", None):
            self.lib.restir_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise RestirError(f"restir error {rc}: {self.lib.restir_last_error(self._ctx).decode()}")

    def synchronize(self):
        self._check(self.lib.restir_synchronize(self._ctx))

    def upload_bvh(self, nodes, triangles):
        nodes = np.ascontiguousarray(nodes).view(np.uint8).reshape(-1, 80)
        tris = np.ascontiguousarray(triangles).view(np.uint8).reshape(-1, 48)
        self._check(self.lib.restir_upload_bvh(self._ctx, _hp(nodes), C.c_uint32(nodes.shape[0]), _hp(tris), C.c_uint32(tris.shape[0])))

    def build_bvh_device(self, triangles, want_nodes=True):
        """restir_build_bvh_device: AabbTree::build on the GPU; returns the (T-1,80)u8 nodes when want_nodes."""
        tris = np.ascontiguousarray(triangles).view(np.uint8).reshape(-1, 48)
        nodes = np.zeros((tris.shape[0] - 1, 80), np.uint8) if want_nodes else None
        self._check(self.lib.restir_build_bvh_device(self._ctx, _hp(tris), C.c_uint32(tris.shape[0]), _hp(nodes) if want_nodes else None))
        return nodes

    def upload_lights(self, point_blob, tri_blob, alias_blob):
        pb, tb, ab = (np.ascontiguousarray(b, np.uint8) for b in (point_blob, tri_blob, alias_blob))
        self._check(self.lib.restir_upload_lights(self._ctx, _hp(pb), C.c_size_t(pb.size), _hp(tb), C.c_size_t(tb.size), _hp(ab),
                                                  C.c_size_t(ab.size)))

    def resize(self, width, height):
        self._check(self.lib.restir_resize(self._ctx, C.c_uint32(width), C.c_uint32(height)))
        self.width, self.height = width, height

    def resize_band(self, width, height, row_begin, row_end, halo):
        self._check(self.lib.restir_resize_band(self._ctx, C.c_uint32(width), C.c_uint32(height), C.c_uint32(row_begin),
                                                C.c_uint32(row_end), C.c_uint32(halo)))
        self.width, self.height = width, height

    def band(self):
        v = [C.c_uint32() for _ in range(4)]
        self._check(self.lib.restir_get_band(self._ctx, *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)

    def alloc_rows(self):
        _, _, a0, a1 = self.band()
        return a1 - a0

    def bind_gbuffer(self, slot, albedo, normal, material, world_pos, depth):
        """Device pointers (torch CUDA tensors or ints)."""
        pl = GBufferPlanes(_dp(albedo), _dp(normal), _dp(material), _dp(world_pos), _dp(depth))
        self._check(self.lib.restir_bind_gbuffer(self._ctx, C.c_int(slot), C.c_int(0), C.byref(pl)))

    def upload_gbuffer(self, slot, albedo, normal, material, world_pos, depth):
        """Host arrays (numpy, or pinned torch CPU tensors) covering rows [alloc_begin, alloc_end)."""
        ptrs = []
        for a in (albedo, normal, material, world_pos, depth):
            ptrs.append(_hp(a) if isinstance(a, np.ndarray) else _dp(a))
        pl = GBufferPlanes(*ptrs)
        self._check(self.lib.restir_upload_gbuffer(self._ctx, C.c_int(slot), C.c_int(0), C.byref(pl)))

    # ---- the G-buffer pass (include/restir_b200.h) ----
    def upload_geometry(self, vertices, indices, draws, matrices):
        """vertices (V,80)u8, indices (I,)u32, draws (D,4)u32 {firstIndex, indexCount, vertexOffset, materialIndex}, matrices (D,128)u8."""
        v = np.ascontiguousarray(vertices).view(np.uint8).reshape(-1, 80)
        i = np.ascontiguousarray(indices, np.uint32).reshape(-1)
        d = np.ascontiguousarray(draws).view(np.uint32).reshape(-1, 4)
        m = np.ascontiguousarray(matrices).view(np.uint8).reshape(-1, 128)
        assert d.shape[0] == m.shape[0]
        self._check(self.lib.restir_upload_geometry(self._ctx, _hp(v), C.c_uint64(v.shape[0]), _hp(i), C.c_uint64(i.shape[0]), _hp(d), _hp(m),
                                                    C.c_uint32(d.shape[0])))

    def upload_materials(self, uniforms, bindings, textures):
        """uniforms (M,64)u8, bindings (M,4)i32, textures: list of (h,w,4)u8 arrays (R8G8B8A8_UNORM, level 0)."""
        u = np.ascontiguousarray(uniforms).view(np.uint8).reshape(-1, 64)
        b = np.ascontiguousarray(bindings, np.int32).reshape(-1, 4)
        keep = [np.ascontiguousarray(t, np.uint8) for t in textures]
        arr = (Texture * max(len(keep), 1))()
        for k, t in enumerate(keep):
            assert t.ndim == 3 and t.shape[2] == 4
            arr[k] = Texture(t.ctypes.data, t.shape[1], t.shape[0])
        self._check(self.lib.restir_upload_materials(self._ctx, _hp(u), _hp(b), C.c_uint32(u.shape[0]), arr if keep else None,
                                                     C.c_uint32(len(keep))))

    def pass_gbuffer(self, slot, cam):
        self._check(self.lib.restir_pass_gbuffer(self._ctx, C.c_int(slot), C.byref(cam)))

    def gbuffer_device_planes(self, slot):
        """Device pointers (ints) of the context-owned planes of `slot`: albedo, normal, material, worldPos, depth."""
        pl = GBufferPlanes()
        self._check(self.lib.restir_gbuffer_device_planes(self._ctx, C.c_int(slot), C.byref(pl)))
        return [pl.albedo, pl.normal, pl.material, pl.worldPos, pl.depth]

    def import_external_memory(self, fd, size):
        """Device pointer (int) of an allocation another API exported as an opaque POSIX fd (restir_import_external_memory)."""
        ptr = C.c_void_p()
        self._check(self.lib.restir_import_external_memory(self._ctx, C.c_int(fd), C.c_uint64(size), C.byref(ptr)))
        return ptr.value

    def release_external_memory(self, ptr):
        self._check(self.lib.restir_release_external_memory(self._ctx, C.c_void_p(ptr)))

    def set_uniforms(self, uniforms):
        u = np.ascontiguousarray(uniforms)
        assert u.dtype == UNIFORMS_DTYPE
        self._check(self.lib.restir_set_uniforms(self._ctx, _hp(u)))

    def set_lighting_uniforms(self, uniforms):
        u = np.ascontiguousarray(uniforms)
        assert u.dtype == LIGHTING_UNIFORMS_DTYPE
        self._check(self.lib.restir_set_lighting_uniforms(self._ctx, _hp(u)))

    def set_reservoir_variant(self, reservoir_size=1, unbiased_mis=False, fused_passes=False):
        """RESERVOIR_SIZE / UNBIASED_MIS / one kernel per shader (restir_set_reservoir_variant)."""
        self._check(self.lib.restir_set_reservoir_variant(self._ctx, C.c_uint32(reservoir_size), C.c_int(1 if unbiased_mis else 0),
                                                          C.c_int(1 if fused_passes else 0)))
        self._variant = (int(reservoir_size), bool(unbiased_mis), bool(fused_passes))

    def reservoir_bytes(self):
        n = C.c_size_t()
        self._check(self.lib.restir_get_reservoir_bytes(self._ctx, C.byref(n)))
        return n.value

    def reservoir_dtype(self):
        """numpy dtype of what download_reservoirs returns: the reference's std430 Reservoir under the context's variant."""
        n, mis, fused = getattr(self, "_variant", (1, False, False))
        if (n, mis) == (1, False):
            return RESERVOIR_DTYPE
        sample = [("position_emissionLum", "<f4", (4,)), ("normal", "<f4", (4,)), ("lightIndex", "<i4"), ("pHat", "<f4"), ("sumWeights", "<f4"),
                  ("w", "<f4")]
        if mis:
            sample += [("sumPHat", "<f4"), ("_pad", "<u4", (3,))]
        return np.dtype([("samples", np.dtype(sample), (n,)), ("M", "<u4"), ("_pad", "<u4", (3,))])

    def set_unbiased_neighbors(self, n):
        self._check(self.lib.restir_set_unbiased_neighbors(self._ctx, C.c_uint32(n)))

    def set_ray_elision(self, enable):
        """0: every ray walked; 1 (default): the exact shortcuts of restir_trace.cu item_resolve; 2: plus one walk per distinct neighbour segment (experiment)."""
        self._check(self.lib.restir_set_ray_elision(self._ctx, C.c_int(int(enable))))

    def set_spatial_staging(self, enable):
        self._check(self.lib.restir_set_spatial_staging(self._ctx, C.c_int(1 if enable else 0)))

    def wide_image(self):
        """(nodes (n, 16) uint32, tri_order (n_triangles,) uint32) of the 4-wide image the context walks."""
        n = C.c_uint32(0)
        info = self.bvh_info()
        nodes = np.zeros((max(info["wide_nodes"], 1), 16), np.uint32)
        order = np.zeros(info["triangles"], np.uint32)
        self._check(self.lib.restir_get_wide_image(self._ctx, _hp(nodes), C.c_uint32(nodes.shape[0]), C.byref(n), _hp(order)))
        return nodes[:n.value], order

    def set_occluder_cache(self, enable):
        """0 off, 1 / True default (by light index up to 4 096 lights, by direction beyond), 2 by light index, 3 by direction."""
        self._check(self.lib.restir_set_occluder_cache(self._ctx, C.c_int(int(enable))))

    def set_traversal(self, mode):
        """Takes effect at the next upload_bvh."""
        self._check(self.lib.restir_set_traversal(self._ctx, C.c_int(mode)))

    def bvh_info(self):
        b = BvhInfo()
        self._check(self.lib.restir_get_bvh_info(self._ctx, C.byref(b)))
        return {n: getattr(b, n) for n, _ in BvhInfo._fields_}

    def profile_begin(self):
        self._check(self.lib.restir_profile_begin(self._ctx))

    def profile_end(self):
        """{kernel name: (launches, total ms)} of everything launched since profile_begin (CUDA events on the context's stream)."""
        arr, n = (KernelTime * 32)(), C.c_uint32()
        self._check(self.lib.restir_profile_end(self._ctx, arr, C.c_uint32(32), C.byref(n)))
        return {arr[i].name.decode(): (arr[i].launches, float(arr[i].total_ms)) for i in range(n.value)}

    def pass_restir(self, gbuffer, out_buffer, prev_buffer):
        self._check(self.lib.restir_pass_restir(self._ctx, C.c_int(gbuffer), C.c_int(out_buffer), C.c_int(prev_buffer)))

    def pass_spatial(self, gbuffer, in_buffer, out_buffer, iteration):
        self._check(self.lib.restir_pass_spatial(self._ctx, C.c_int(gbuffer), C.c_int(in_buffer), C.c_int(out_buffer), C.c_int(iteration)))

    def pass_unbiased(self, gbuffer, in_buffer, out_buffer):
        self._check(self.lib.restir_pass_unbiased(self._ctx, C.c_int(gbuffer), C.c_int(in_buffer), C.c_int(out_buffer)))

    def pass_lighting(self, gbuffer, buffer, out, out_format=RESTIR_OUT_RGBA32F):
        self._check(self.lib.restir_pass_lighting(self._ctx, C.c_int(gbuffer), C.c_int(buffer), _dp(out), C.c_int(out_format)))

    def frame(self, i, unbiased, spatial_iterations=1):
        self._check(self.lib.restir_frame(self._ctx, C.c_int(i), C.c_int(1 if unbiased else 0), C.c_int(spatial_iterations)))

    def frame_lit(self, i, unbiased, spatial_iterations, out, out_format=RESTIR_OUT_RGBA32F):
        """restir_frame + restir_pass_lighting(i, FRAME[i]) with the lighting fused into the last reuse kernel."""
        self._check(self.lib.restir_frame_lit(self._ctx, C.c_int(i), C.c_int(1 if unbiased else 0), C.c_int(spatial_iterations), _dp(out),
                                              C.c_int(out_format)))

    def download_reservoirs(self, buffer):
        dt = self.reservoir_dtype()
        assert dt.itemsize == self.reservoir_bytes()
        out = np.zeros(self.alloc_rows() * self.width, dt)
        self._check(self.lib.restir_download_reservoirs(self._ctx, C.c_int(buffer), _hp(out)))
        return out

    def upload_reservoirs(self, buffer, reservoirs):
        r = np.ascontiguousarray(reservoirs)
        assert r.dtype.itemsize == self.reservoir_bytes() and r.size == self.alloc_rows() * self.width
        self._check(self.lib.restir_upload_reservoirs(self._ctx, C.c_int(buffer), _hp(r)))

    def reservoir_device_ptr(self, buffer):
        ptr, pitch = C.c_void_p(), C.c_size_t()
        self._check(self.lib.restir_reservoir_device_ptr(self._ctx, C.c_int(buffer), C.byref(ptr), C.byref(pitch)))
        return ptr.value, pitch.value

    # ---- row-band neighbours over peer memory (include/restir_b200.h) ----
    def band_local_peer(self):
        peer = BandPeer()
        self._check(self.lib.restir_band_local_peer(self._ctx, C.byref(peer)))
        return peer

    def band_export_ipc(self):
        """The 272 bytes a neighbour process needs (bytes object)."""
        ipc = BandIpc()
        self._check(self.lib.restir_band_export_ipc(self._ctx, C.byref(ipc)))
        return bytes(ipc)

    def band_open_ipc(self, blob):
        ipc = BandIpc.from_buffer_copy(blob)
        peer = BandPeer()
        self._check(self.lib.restir_band_open_ipc(self._ctx, C.byref(ipc), C.byref(peer)))
        return peer

    def band_connect(self, side, peer):
        """side 0: the neighbour that owns the rows above, 1: below; peer None: no neighbour there."""
        self._check(self.lib.restir_band_connect(self._ctx, C.c_int(side), C.byref(peer) if peer is not None else None))

    def trace_segments(self, p1, p2, n, shadowed):
        self._check(self.lib.restir_trace_segments(self._ctx, _dp(p1), _dp(p2), C.c_uint64(n), _dp(shadowed)))

    def counters(self, reset=False, check=True):
        """check=False: return the counters even when halo_wait_timeouts != 0 (the call then reports RESTIR_E_HALO)."""
        c = Counters()
        rc = self.lib.restir_get_counters(self._ctx, C.byref(c), C.c_int(1 if reset else 0))
        if rc != RESTIR_E_HALO or check:
            self._check(rc)
        return {"shadow_rays": c.shadow_rays, "stack_overflows": c.stack_overflows, "halo_misses": c.halo_misses,
                "kernel_launches": c.kernel_launches, "shadow_rays_traced": c.shadow_rays_traced,
                "halo_wait_timeouts": c.halo_wait_timeouts, "shadow_rays_cached": c.shadow_rays_cached}

    def raycast_gbuffer(self, cam, tri_material, material_table, albedo, normal, material, world_pos, depth):
        self._check(self.lib.restir_tools_raycast_gbuffer(self._ctx, C.byref(cam), _dp(tri_material), _dp(material_table), _dp(albedo),
                                                          _dp(normal), _dp(material), _dp(world_pos), _dp(depth)))
