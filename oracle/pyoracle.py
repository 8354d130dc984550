"""ctypes binding of oracle/liboracle.so — the CPU restatement of the reference's ReSTIR passes.

TEST INFRASTRUCTURE.  Import only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

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
assert RESERVOIR_DTYPE.itemsize == 64

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
assert UNIFORMS_DTYPE.itemsize == 128

LIGHTING_UNIFORMS_DTYPE = np.dtype(
    [
        ("prevFrameProjectionViewMatrix", "<f4", (16,)),
        ("cameraPos", "<f4", (4,)),
        ("bufferSize", "<u4", (2,)),
        ("debugMode", "<i4"),
        ("gamma", "<f4"),
    ]
)
assert LIGHTING_UNIFORMS_DTYPE.itemsize == 96


class OracleScene(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("nodes", "tris", "pointBlob", "triBlob", "aliasBlob")]


class OracleGBuffer(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("albedo", "normal", "material", "worldPos", "depth")]


class OracleCamera(C.Structure):
    _fields_ = [
        ("position", C.c_float * 3),
        ("lookAt", C.c_float * 3),
        ("worldUp", C.c_float * 3),
        ("zNear", C.c_float),
        ("zFar", C.c_float),
        ("fovYRadians", C.c_float),
        ("aspectRatio", C.c_float),
    ]


def build(force=False):
    """Compile oracle/liboracle.so (and, where /root/reference exists, oracle/_ref)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "restir_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.oracle_num_threads.restype = C.c_int
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _chk(a, dtype, name):
    assert a.flags["C_CONTIGUOUS"], name
    assert a.dtype == np.dtype(dtype), (name, a.dtype, dtype)
    return a


class Scene:
    """nodes (N,80)u8, tris (T,48)u8 and the three light blobs (u8) in the reference's byte layouts."""

    def __init__(self, nodes, tris, point_blob, tri_blob, alias_blob):
        self.nodes = np.ascontiguousarray(nodes).view(np.uint8).reshape(-1)
        self.tris = np.ascontiguousarray(tris).view(np.uint8).reshape(-1)
        self.point_blob = np.ascontiguousarray(point_blob, dtype=np.uint8)
        self.tri_blob = np.ascontiguousarray(tri_blob, dtype=np.uint8)
        self.alias_blob = np.ascontiguousarray(alias_blob, dtype=np.uint8)
        self.c = OracleScene(_p(self.nodes), _p(self.tris), _p(self.point_blob), _p(self.tri_blob), _p(self.alias_blob))


class GBuffer:
    """The five planes in the NVIDIA-default formats (include/restir_layouts.h)."""

    def __init__(self, w, h, albedo=None, normal=None, material=None, world_pos=None, depth=None):
        self.w, self.h = w, h
        self.albedo = np.zeros((h, w, 4), np.uint8) if albedo is None else _chk(albedo, np.uint8, "albedo")
        self.normal = np.zeros((h, w, 4), np.int16) if normal is None else _chk(normal, np.int16, "normal")
        self.material = np.zeros((h, w, 2), np.uint16) if material is None else _chk(material, np.uint16, "material")
        self.world_pos = np.zeros((h, w, 4), np.float32) if world_pos is None else _chk(world_pos, np.float32, "worldPos")
        self.depth = np.zeros((h, w), np.float32) if depth is None else _chk(depth, np.float32, "depth")
        self.c = OracleGBuffer(_p(self.albedo), _p(self.normal), _p(self.material), _p(self.world_pos), _p(self.depth))

    def planes(self):
        return self.albedo, self.normal, self.material, self.world_pos, self.depth

    def nbytes(self):
        return sum(p.nbytes for p in self.planes())


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


def make_camera(position=(3.0, 4.0, 5.0), look_at=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), z_near=0.01, z_far=1000.0,
                fov_y=None, aspect=1.0):
    """Defaults = reference src/camera.h:7-13."""
    cam = OracleCamera()
    cam.position[:] = position
    cam.lookAt[:] = look_at
    cam.worldUp[:] = up
    cam.zNear, cam.zFar = z_near, z_far
    cam.fovYRadians = float(np.float32(0.5) * np.float32(np.pi)) if fov_y is None else fov_y
    cam.aspectRatio = aspect
    return cam


def camera_matrix(cam):
    out = np.zeros(16, np.float32)
    lib().oracle_camera_matrix(C.byref(cam), _p(out))
    return out


def pcg32(seed, seq, n):
    out = np.zeros(n, np.uint32)
    lib().oracle_pcg32(C.c_uint64(seed), C.c_uint64(seq), C.c_int(n), _p(out))
    return out


def rand_floats(seed, seq, n):
    out = np.zeros(n, np.float32)
    lib().oracle_rand_floats(C.c_uint64(seed), C.c_uint64(seq), C.c_int(n), _p(out))
    return out


def sincos(a):
    a = np.ascontiguousarray(a, np.float32)
    s, c = np.zeros_like(a), np.zeros_like(a)
    lib().oracle_sincos(_p(a), C.c_int(a.size), _p(s), _p(c))
    return s, c


def evaluate_phat(args, albedo_lum, emission_lum, roughness, metallic):
    args = np.ascontiguousarray(args, np.float32).reshape(-1, 16)
    out = np.zeros(args.shape[0], np.float32)
    lib().oracle_evaluate_phat(_p(args), C.c_int(args.shape[0]), C.c_float(albedo_lum), C.c_float(emission_lum),
                               C.c_float(roughness), C.c_float(metallic), _p(out))
    return out


def trace_segments(scene, p1, p2, want_margin=False):
    p1 = np.ascontiguousarray(p1, np.float32).reshape(-1, 3)
    p2 = np.ascontiguousarray(p2, np.float32).reshape(-1, 3)
    n = p1.shape[0]
    shadowed = np.zeros(n, np.uint8)
    margin = np.zeros(n, np.float32) if want_margin else None
    overflow = np.zeros(n, np.int32) if want_margin else None
    lib().oracle_trace_segments(C.byref(scene.c), C.c_int64(n), _p(p1), _p(p2), _p(shadowed), _p(margin), _p(overflow))
    return (shadowed, margin, overflow) if want_margin else shadowed


def restir_pass(scene, uniforms, cur, prev, prev_reservoirs, rows=None):
    w, h = int(uniforms["screenSize"][0]), int(uniforms["screenSize"][1])
    y0, y1 = rows or (0, h)
    out = np.zeros(w * h, RESERVOIR_DTYPE)
    rays = C.c_uint64(0)
    u = np.ascontiguousarray(uniforms)
    prev_reservoirs = np.ascontiguousarray(prev_reservoirs)
    assert prev_reservoirs.dtype == RESERVOIR_DTYPE and prev_reservoirs.size == w * h
    lib().oracle_restir_pass(C.byref(scene.c), _p(u), C.byref(cur.c), C.byref(prev.c) if prev is not None else None,
                             _p(prev_reservoirs), _p(out), C.c_int(y0), C.c_int(y1), C.byref(rays))
    return out, rays.value


def spatial_pass(uniforms, cur, reservoirs, iteration, rows=None):
    w, h = int(uniforms["screenSize"][0]), int(uniforms["screenSize"][1])
    y0, y1 = rows or (0, h)
    out = np.zeros(w * h, RESERVOIR_DTYPE)
    u = np.ascontiguousarray(uniforms)
    reservoirs = np.ascontiguousarray(reservoirs)
    lib().oracle_spatial_pass(_p(u), C.byref(cur.c), _p(reservoirs), _p(out), C.c_int(iteration), C.c_int(y0), C.c_int(y1))
    return out


def unbiased_pass(scene, uniforms, cur, reservoirs, num_neighbors=3, rows=None):
    w, h = int(uniforms["screenSize"][0]), int(uniforms["screenSize"][1])
    y0, y1 = rows or (0, h)
    out = np.zeros(w * h, RESERVOIR_DTYPE)
    rays = C.c_uint64(0)
    u = np.ascontiguousarray(uniforms)
    reservoirs = np.ascontiguousarray(reservoirs)
    lib().oracle_unbiased_pass(C.byref(scene.c), _p(u), C.byref(cur.c), _p(reservoirs), _p(out), C.c_int(num_neighbors),
                               C.c_int(y0), C.c_int(y1), C.byref(rays))
    return out, rays.value


def lighting_pass(scene, lighting_uniforms, cur, reservoirs, rows=None):
    w, h = int(lighting_uniforms["bufferSize"][0]), int(lighting_uniforms["bufferSize"][1])
    y0, y1 = rows or (0, h)
    out = np.zeros((h, w, 4), np.float32)
    u = np.ascontiguousarray(lighting_uniforms)
    reservoirs = np.ascontiguousarray(reservoirs)
    lib().oracle_lighting_pass(C.byref(scene.c), _p(u), C.byref(cur.c), _p(reservoirs), _p(out), C.c_int(y0), C.c_int(y1))
    return out


def raycast_gbuffer(scene, tri_material, material_table, cam, w, h, rows=None):
    y0, y1 = rows or (0, h)
    g = GBuffer(w, h)
    tri_material = np.ascontiguousarray(tri_material, np.int32)
    material_table = np.ascontiguousarray(material_table, np.uint32)
    lib().oracle_raycast_gbuffer(C.byref(scene.c), _p(tri_material), _p(material_table), C.byref(cam), C.c_int(w), C.c_int(h),
                                 C.c_int(y0), C.c_int(y1), _p(g.albedo), _p(g.normal), _p(g.material), _p(g.world_pos),
                                 _p(g.depth))
    return g


class GBufferInputs:
    """What the G-buffer pass binds, as numpy arrays in the reference's layouts (include/restir_layouts.h):
    vertices (V,80)u8, indices (I,)u32, draws (D,4)u32, matrices (D,128)u8, uniforms (M,64)u8, bindings (M,4)i32,
    textures: list of (h,w,4)u8."""

    def __init__(self, vertices, indices, draws, matrices, uniforms, bindings, textures):
        self.vertices = np.ascontiguousarray(vertices).view(np.uint8).reshape(-1, 80)
        self.indices = np.ascontiguousarray(indices, np.uint32).reshape(-1)
        self.draws = np.ascontiguousarray(draws).view(np.uint32).reshape(-1, 4)
        self.matrices = np.ascontiguousarray(matrices).view(np.uint8).reshape(-1, 128)
        self.uniforms = np.ascontiguousarray(uniforms).view(np.uint8).reshape(-1, 64)
        self.bindings = np.ascontiguousarray(bindings, np.int32).reshape(-1, 4)
        self.textures = [np.ascontiguousarray(t, np.uint8) for t in textures]
        n_tris = int(self.draws[:, 1].sum()) // 3
        self.attrs = np.zeros((n_tris, 32), np.float32)
        self.tri_material = np.zeros(n_tris, np.int32)
        lib().oracle_vertex_stage(_p(self.vertices), _p(self.indices), _p(self.draws), _p(self.matrices), C.c_uint32(self.draws.shape[0]),
                                  _p(self.attrs), _p(self.tri_material))
        table, first = [], 0
        for t in self.textures:
            table.append((first, t.shape[1], t.shape[0], 0))
            first += t.shape[0] * t.shape[1]
        self.texture_table = np.asarray(table, np.uint32).reshape(-1, 4)
        self.texels = np.concatenate([t.reshape(-1) for t in self.textures]) if self.textures else np.zeros(4, np.uint8)


def gbuffer_pass(scene, inputs, cam, w, h, rows=None):
    """CPU twin of restir_pass_gbuffer (gBuffer.vert / gBuffer.frag by ray casting, oracle_gbuffer_pass)."""
    y0, y1 = rows or (0, h)
    g = GBuffer(w, h)
    lib().oracle_gbuffer_pass(C.byref(scene.c), _p(inputs.attrs), _p(inputs.tri_material), _p(inputs.uniforms), _p(inputs.bindings),
                              C.c_int(inputs.uniforms.shape[0]), _p(inputs.texels), _p(inputs.texture_table), C.c_int(len(inputs.textures)),
                              C.byref(cam), C.c_int(w), C.c_int(h), C.c_int(y0), C.c_int(y1), _p(g.albedo), _p(g.normal), _p(g.material),
                              _p(g.world_pos), _p(g.depth))
    return g


def num_threads():
    return lib().oracle_num_threads()


def set_num_threads(n):
    lib().oracle_set_num_threads(C.c_int(int(n)))


# ---- the reference's compile-time variants: RESERVOIR_SIZE > 1, UNBIASED_MIS (restirStructs.glsl:16-17) -----------------

def variant_reservoir_dtype(reservoir_size, unbiased_mis):
    """std430 Reservoir under the two switches: RESERVOIR_SIZE samples of 48 (64 with sumPHat) bytes + numStreamSamples padded to 16."""
    sample = [("position_emissionLum", "<f4", (4,)), ("normal", "<f4", (4,)), ("lightIndex", "<i4"), ("pHat", "<f4"), ("sumWeights", "<f4"),
              ("w", "<f4")]
    if unbiased_mis:
        sample += [("sumPHat", "<f4"), ("_pad", "<u4", (3,))]
    dt = np.dtype([("samples", np.dtype(sample), (reservoir_size,)), ("M", "<u4"), ("_pad", "<u4", (3,))])
    assert dt.itemsize == reservoir_size * (64 if unbiased_mis else 48) + 16
    return dt


class Variant:
    """The oracle's passes for one (RESERVOIR_SIZE, UNBIASED_MIS): same call shapes as the module-level functions."""

    def __init__(self, reservoir_size=1, unbiased_mis=False):
        self.n, self.mis = int(reservoir_size), bool(unbiased_mis)
        self.dtype = self.RESERVOIR_DTYPE = variant_reservoir_dtype(self.n, self.mis)
        self.UNIFORMS_DTYPE, self.LIGHTING_UNIFORMS_DTYPE = UNIFORMS_DTYPE, LIGHTING_UNIFORMS_DTYPE
        assert lib().oracle_variant_reservoir_bytes(C.c_int(self.n), C.c_int(self.mis)) == self.dtype.itemsize

    def _v(self):
        return C.c_int(self.n), C.c_int(1 if self.mis else 0)

    def restir_pass(self, scene, uniforms, cur, prev, prev_reservoirs, rows=None):
        w, h = int(uniforms["screenSize"][0]), int(uniforms["screenSize"][1])
        y0, y1 = rows or (0, h)
        out = np.zeros(w * h, self.dtype)
        rays = C.c_uint64(0)
        u = np.ascontiguousarray(uniforms)
        prev_reservoirs = np.ascontiguousarray(prev_reservoirs)
        assert prev_reservoirs.dtype == self.dtype and prev_reservoirs.size == w * h
        rc = lib().oracle_restir_pass_variant(*self._v(), C.byref(scene.c), _p(u), C.byref(cur.c), C.byref(prev.c) if prev is not None else None,
                                              _p(prev_reservoirs), _p(out), C.c_int(y0), C.c_int(y1), C.byref(rays))
        assert rc == 0
        return out, rays.value

    def spatial_pass(self, uniforms, cur, reservoirs, iteration, rows=None):
        w, h = int(uniforms["screenSize"][0]), int(uniforms["screenSize"][1])
        y0, y1 = rows or (0, h)
        out = np.zeros(w * h, self.dtype)
        u = np.ascontiguousarray(uniforms)
        reservoirs = np.ascontiguousarray(reservoirs)
        assert reservoirs.dtype == self.dtype
        rc = lib().oracle_spatial_pass_variant(*self._v(), _p(u), C.byref(cur.c), _p(reservoirs), _p(out), C.c_int(iteration), C.c_int(y0), C.c_int(y1))
        assert rc == 0
        return out

    def unbiased_pass(self, scene, uniforms, cur, reservoirs, num_neighbors=3, rows=None):
        w, h = int(uniforms["screenSize"][0]), int(uniforms["screenSize"][1])
        y0, y1 = rows or (0, h)
        out = np.zeros(w * h, self.dtype)
        rays = C.c_uint64(0)
        u = np.ascontiguousarray(uniforms)
        reservoirs = np.ascontiguousarray(reservoirs)
        assert reservoirs.dtype == self.dtype
        rc = lib().oracle_unbiased_pass_variant(*self._v(), C.byref(scene.c), _p(u), C.byref(cur.c), _p(reservoirs), _p(out), C.c_int(num_neighbors),
                                                C.c_int(y0), C.c_int(y1), C.byref(rays))
        assert rc == 0
        return out, rays.value

    def lighting_pass(self, scene, lighting_uniforms, cur, reservoirs, rows=None):
        w, h = int(lighting_uniforms["bufferSize"][0]), int(lighting_uniforms["bufferSize"][1])
        y0, y1 = rows or (0, h)
        out = np.zeros((h, w, 4), np.float32)
        u = np.ascontiguousarray(lighting_uniforms)
        reservoirs = np.ascontiguousarray(reservoirs)
        assert reservoirs.dtype == self.dtype
        rc = lib().oracle_lighting_pass_variant(*self._v(), C.byref(scene.c), _p(u), C.byref(cur.c), _p(reservoirs), _p(out), C.c_int(y0), C.c_int(y1))
        assert rc == 0
        return out
