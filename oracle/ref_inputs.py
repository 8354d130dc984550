"""TEST / BENCH INFRASTRUCTURE — inputs of the reference arm (`bench.py --impl reference`) without the product.

The CPU arm must not import, load or execute anything of `restir-vulkan_b200/`: its scene comes from the blobs the
reference's OWN code produced (`oracle/_ref/scene_baker`, built from /root/reference by oracle/ref_build/Makefile and
run into scenes/_baked/<name>/ref_*.bin): AabbTree::build's nodes and triangles (src/aabbTreeBuilder.cpp:52-214), the
light lists and the alias table (src/misc.cpp:343-497).  numpy only.
"""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BAKED_DIR = os.path.join(ROOT, "scenes", "_baked")


class ReferenceScene:
    """The reference's own blobs of one baked scene."""

    def __init__(self, name):
        d = os.path.join(BAKED_DIR, name)
        if not os.path.exists(os.path.join(d, "ref_nodes.bin")):
            raise FileNotFoundError(f"{d}/ref_nodes.bin missing: run `make -C oracle ref` where /root/reference exists")
        f = lambda n, dt: np.fromfile(os.path.join(d, n), dtype=dt)
        self.name = name
        self.triangles = f("triangles.bin", np.uint8).reshape(-1, 48)
        self.tri_material = f("tri_material.i32", np.int32)
        self.materials = f("materials.f32", np.float32).reshape(-1, 16)
        self.dims = f("dims.f32", np.float32)
        self.nodes = f("ref_nodes.bin", np.uint8).reshape(-1, 80)
        self.point_blob = f("ref_point_lights.bin", np.uint8)
        self.tri_blob = f("ref_tri_lights.bin", np.uint8)
        self.alias_blob = f("ref_alias.bin", np.uint8)
        self.manifest = json.load(open(os.path.join(d, "manifest.json")))

    def light_counts(self):
        return int(self.point_blob[:4].view(np.int32)[0]), int(self.tri_blob[:4].view(np.int32)[0])

    def material_table(self):
        return material_table(self.materials)


def available(name):
    return os.path.exists(os.path.join(BAKED_DIR, name, "ref_nodes.bin"))


def _srgb8(c):
    c = np.clip(np.asarray(c, np.float64), 0.0, 1.0)
    return np.rint(np.where(c <= 0.0031308, 12.92 * c, 1.055 * np.power(c, 1.0 / 2.4) - 0.055) * 255.0).astype(np.uint32)


def material_table(materials):
    """src/shaders/gBuffer.frag:27-79 with every texture = 1 (factor-only fixtures): per material
    {albedo RGBA8 with sRGB-encoded rgb and alpha = emissive flag, (roughness, metallic) RG16, bit 0 = discarded by
    ALPHA_MODE_MASK, 0}.  Stated here independently of restir-vulkan_b200/fixtures.py; tests/test_bench_inputs.py holds
    the two equal."""
    rows = np.asarray(materials, np.float64).reshape(-1, 16)
    table = np.zeros((rows.shape[0], 4), np.uint32)
    for i, m in enumerate(rows):
        base, param, emissive = m[0:4], m[4:8], m[8:11]
        if int(m[11]) == 0:                                   # metallic-roughness, gBuffer.frag:47-50
            albedo, roughness, metallic = base[:3], param[1], param[2]
        else:                                                 # specular-glossiness, :51-67
            avg = 0.5 * (base[:3] + param[:3])
            root = np.sqrt(np.maximum(avg * avg - 0.04 * base[:3], 0.0))
            albedo, roughness, metallic = avg + root, 1.0 - param[3], float(np.mean(25.0 * avg - root))
        emits = float(np.sqrt(np.dot(emissive, emissive))) > 0.0  # :73-79
        rgb = _srgb8(base[:3] * emissive if emits else albedo)
        rg = np.rint(np.clip([roughness, metallic], 0.0, 1.0) * 65535.0).astype(np.uint32)
        table[i, 0] = rgb[0] | (rgb[1] << 8) | (rgb[2] << 16) | ((255 if emits else 0) << 24)
        table[i, 1] = rg[0] | (rg[1] << 16)
        table[i, 2] = 1 if (int(m[12]) == 1 and base[3] < m[13]) else 0   # :30-34
    return table
