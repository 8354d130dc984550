"""ctypes binding of oracle/_ref/libglslref.so — the reference's OWN shader sources compiled as C++
(oracle/ref_build/glsl_ref.cpp, glsl_shim.h, glsl2cpp.py).

TEST INFRASTRUCTURE.  Used only to pin the CPU oracle (tests/test_oracle_vs_glsl.py): same entry points and
argument meaning as oracle/pyoracle.py, so the two can be run side by side on the same inputs.  The library
is built where /root/reference exists (this container) and travels to the GPU box as a prebuilt file.
"""
import ctypes as C
import os

import numpy as np

import pyoracle as po

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libglslref.so")
_LIB = None

RESERVOIR_DTYPE = po.RESERVOIR_DTYPE
UNIFORMS_DTYPE = po.UNIFORMS_DTYPE
LIGHTING_UNIFORMS_DTYPE = po.LIGHTING_UNIFORMS_DTYPE


def available():
    return os.path.exists(SO)


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(SO)
        _LIB.glslref_unbiased_pass.restype = C.c_int
        _LIB.glslref_set_num_threads(C.c_int(os.cpu_count() or 1))
    return _LIB


_p = po._p


def pcg32(seed, seq, n):
    out = np.zeros(n, np.uint32)
    lib().glslref_pcg32(C.c_uint64(seed), C.c_uint64(seq), C.c_int(n), _p(out), None)
    return out


def rand_floats(seed, seq, n):
    out = np.zeros(n, np.float32)
    lib().glslref_pcg32(C.c_uint64(seed), C.c_uint64(seq), C.c_int(n), None, _p(out))
    return out


def evaluate_phat(args, albedo_lum, emission_lum, roughness, metallic):
    args = np.ascontiguousarray(args, np.float32).reshape(-1, 16)
    out = np.zeros(args.shape[0], np.float32)
    lib().glslref_evaluate_phat(_p(args), C.c_int(args.shape[0]), C.c_float(albedo_lum), C.c_float(emission_lum),
                                C.c_float(roughness), C.c_float(metallic), _p(out))
    return out


def trace_segments(scene, p1, p2):
    p1 = np.ascontiguousarray(p1, np.float32).reshape(-1, 3)
    p2 = np.ascontiguousarray(p2, np.float32).reshape(-1, 3)
    shadowed = np.zeros(p1.shape[0], np.uint8)
    lib().glslref_trace_segments(C.byref(scene.c), C.c_longlong(p1.shape[0]), _p(p1), _p(p2), _p(shadowed))
    return shadowed


def _size(u):
    return int(u["screenSize"][0]), int(u["screenSize"][1])


def restir_pass(scene, uniforms, cur, prev, prev_reservoirs, rows=None):
    w, h = _size(uniforms)
    y0, y1 = rows or (0, h)
    out = np.zeros(w * h, RESERVOIR_DTYPE)
    u = np.ascontiguousarray(uniforms)
    prev_reservoirs = np.ascontiguousarray(prev_reservoirs)
    lib().glslref_restir_pass(C.byref(scene.c), _p(u), C.byref(cur.c), C.byref(prev.c) if prev is not None else None,
                              _p(prev_reservoirs), _p(out), C.c_int(y0), C.c_int(y1))
    return out, None


def spatial_pass(uniforms, cur, reservoirs, iteration, rows=None):
    w, h = _size(uniforms)
    y0, y1 = rows or (0, h)
    out = np.zeros(w * h, RESERVOIR_DTYPE)
    u = np.ascontiguousarray(uniforms)
    reservoirs = np.ascontiguousarray(reservoirs)
    lib().glslref_spatial_pass(_p(u), C.byref(cur.c), _p(reservoirs), _p(out), C.c_int(iteration), C.c_int(y0), C.c_int(y1))
    return out


def unbiased_pass(scene, uniforms, cur, reservoirs, num_neighbors=3, rows=None):
    w, h = _size(uniforms)
    y0, y1 = rows or (0, h)
    out = np.zeros(w * h, RESERVOIR_DTYPE)
    u = np.ascontiguousarray(uniforms)
    reservoirs = np.ascontiguousarray(reservoirs)
    rc = lib().glslref_unbiased_pass(C.byref(scene.c), _p(u), C.byref(cur.c), _p(reservoirs), _p(out), C.c_int(num_neighbors),
                                     C.c_int(y0), C.c_int(y1))
    if rc != 0:
        raise ValueError(f"libglslref is compiled for NUM_NEIGHBORS 3 (as shipped) and 5, not {num_neighbors}")
    return out, None


def lighting_pass(scene, lighting_uniforms, cur, reservoirs, rows=None):
    w, h = int(lighting_uniforms["bufferSize"][0]), int(lighting_uniforms["bufferSize"][1])
    y0, y1 = rows or (0, h)
    out = np.zeros((h, w, 4), np.float32)
    u = np.ascontiguousarray(lighting_uniforms)
    reservoirs = np.ascontiguousarray(reservoirs)
    lib().glslref_lighting_pass(C.byref(scene.c), _p(u), C.byref(cur.c), _p(reservoirs), _p(out), C.c_int(y0), C.c_int(y1))
    return out


# ---- the same sources under the switches the authors ship: RESERVOIR_SIZE 2 / 4, UNBIASED_MIS (oracle/ref_build/Makefile) -----

class Variant:
    """libglslref_<name>.so: the reference's shaders transliterated with `#define RESERVOIR_SIZE n` and / or `#define UNBIASED_MIS`."""

    def __init__(self, reservoir_size=1, unbiased_mis=False):
        self.n, self.mis = int(reservoir_size), bool(unbiased_mis)
        name = f"rs{self.n}" + ("_mis" if self.mis else "")
        self.so = SO if (self.n, self.mis) == (1, False) else os.path.join(_HERE, "_ref", f"libglslref_{name}.so")
        self.dtype = self.RESERVOIR_DTYPE = po.variant_reservoir_dtype(self.n, self.mis)
        self.UNIFORMS_DTYPE, self.LIGHTING_UNIFORMS_DTYPE = UNIFORMS_DTYPE, LIGHTING_UNIFORMS_DTYPE
        self._lib = None

    def available(self):
        return os.path.exists(self.so)

    def lib(self):
        if self._lib is None:
            self._lib = C.CDLL(self.so)
            self._lib.glslref_unbiased_pass.restype = C.c_int
            self._lib.glslref_set_num_threads(C.c_int(os.cpu_count() or 1))
            assert self._lib.glslref_reservoir_bytes() == self.dtype.itemsize
        return self._lib

    def restir_pass(self, scene, uniforms, cur, prev, prev_reservoirs, rows=None):
        w, h = _size(uniforms)
        y0, y1 = rows or (0, h)
        out = np.zeros(w * h, self.dtype)
        u = np.ascontiguousarray(uniforms)
        prev_reservoirs = np.ascontiguousarray(prev_reservoirs)
        assert prev_reservoirs.dtype == self.dtype
        self.lib().glslref_restir_pass(C.byref(scene.c), _p(u), C.byref(cur.c), C.byref(prev.c) if prev is not None else None,
                                       _p(prev_reservoirs), _p(out), C.c_int(y0), C.c_int(y1))
        return out, None

    def spatial_pass(self, uniforms, cur, reservoirs, iteration, rows=None):
        w, h = _size(uniforms)
        y0, y1 = rows or (0, h)
        out = np.zeros(w * h, self.dtype)
        u = np.ascontiguousarray(uniforms)
        reservoirs = np.ascontiguousarray(reservoirs)
        self.lib().glslref_spatial_pass(_p(u), C.byref(cur.c), _p(reservoirs), _p(out), C.c_int(iteration), C.c_int(y0), C.c_int(y1))
        return out

    def unbiased_pass(self, scene, uniforms, cur, reservoirs, num_neighbors=3, rows=None):
        w, h = _size(uniforms)
        y0, y1 = rows or (0, h)
        out = np.zeros(w * h, self.dtype)
        u = np.ascontiguousarray(uniforms)
        reservoirs = np.ascontiguousarray(reservoirs)
        rc = self.lib().glslref_unbiased_pass(C.byref(scene.c), _p(u), C.byref(cur.c), _p(reservoirs), _p(out), C.c_int(num_neighbors),
                                              C.c_int(y0), C.c_int(y1))
        if rc != 0:
            raise ValueError(f"libglslref is compiled for NUM_NEIGHBORS 3 (as shipped) and 5, not {num_neighbors}")
        return out, None

    def lighting_pass(self, scene, lighting_uniforms, cur, reservoirs, rows=None):
        w, h = int(lighting_uniforms["bufferSize"][0]), int(lighting_uniforms["bufferSize"][1])
        y0, y1 = rows or (0, h)
        out = np.zeros((h, w, 4), np.float32)
        u = np.ascontiguousarray(lighting_uniforms)
        reservoirs = np.ascontiguousarray(reservoirs)
        self.lib().glslref_lighting_pass(C.byref(scene.c), _p(u), C.byref(cur.c), _p(reservoirs), _p(out), C.c_int(y0), C.c_int(y1))
        return out

