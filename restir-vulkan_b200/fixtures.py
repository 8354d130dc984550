"""Scene inputs for tests and bench: baked reference scenes, procedural scenes, material tables.

Host-side plumbing only (numpy); no compute of the hot path happens here.
"""
import json
import os

import numpy as np

from . import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BAKED_DIR = os.path.join(ROOT, "scenes", "_baked")


class SceneData:
    """Everything the passes bind for one scene, in the reference's byte layouts."""

    def __init__(self, name, triangles, tri_material, materials, nodes, point_blob, tri_blob, alias_blob, dims, ref=None):
        self.name = name
        self.triangles = triangles          # (T,48) u8
        self.tri_material = tri_material    # (T,) i32
        self.materials = materials          # (M,16) f32: colorParam4, materialParam4, emissive3, shadingModel, alphaMode, alphaCutoff, 0, 0
        self.nodes = nodes                  # (T-1,80) u8
        self.point_blob, self.tri_blob, self.alias_blob = point_blob, tri_blob, alias_blob
        self.dims = dims                    # (6,) f32 min xyz, max xyz (GltfScene::m_dimensions)
        self.ref = ref or {}                # reference-produced blobs, when baked

    @property
    def n_triangles(self):
        return self.triangles.shape[0]

    def light_counts(self):
        return int(self.point_blob[:4].view(np.int32)[0]), int(self.tri_blob[:4].view(np.int32)[0])

    def material_table(self):
        return material_table(self.materials)


def baked_available(name):
    return os.path.exists(os.path.join(BAKED_DIR, name, "triangles.bin"))


def load_baked(name, rebuild=True):
    """Load scenes/_baked/<name> (written by oracle/_ref/scene_baker from the reference's own code).

    With rebuild=True the BVH, lights and alias table are REBUILT by this repo's host builders from the
    triangle list (the reference-produced ones are kept in .ref for comparison); with rebuild=False the
    reference's blobs are used directly.
    """
    d = os.path.join(BAKED_DIR, name)
    f = lambda n, dt: np.fromfile(os.path.join(d, n), dtype=dt)
    tris = f("triangles.bin", np.uint8).reshape(-1, 48)
    tri_material = f("tri_material.i32", np.int32)
    materials = f("materials.f32", np.float32).reshape(-1, 16)
    dims = f("dims.f32", np.float32)
    ref = {
        "nodes": f("ref_nodes.bin", np.uint8).reshape(-1, 80),
        "point_blob": f("ref_point_lights.bin", np.uint8),
        "tri_blob": f("ref_tri_lights.bin", np.uint8),
        "alias_blob": f("ref_alias.bin", np.uint8),
        "gltf_point_blob": f("gltf_point_lights.bin", np.uint8),
        "manifest": json.load(open(os.path.join(d, "manifest.json"))),
    }
    if not rebuild:
        return SceneData(name, tris, tri_material, materials, ref["nodes"], ref["point_blob"], ref["tri_blob"], ref["alias_blob"], dims, ref)
    file_lights = ref["gltf_point_blob"][16:].reshape(-1, 32)
    return assemble_scene(name, tris, tri_material, materials, file_lights, dims, ref)


def assemble_scene(name, tris, tri_material, materials, file_point_lights, dims, ref=None, random_lights=200):
    """What App does at start-up around the builders (src/sceneBuffers.h:78-84, src/app.cpp:358-360)."""
    nodes = capi.build_aabb_tree(tris)
    point = np.ascontiguousarray(file_point_lights, np.uint8).reshape(-1, 32)
    tri = capi.collect_triangle_lights(tris, tri_material, materials[:, 8:11])
    if point.shape[0] == 0 and tri.shape[0] == 0:
        point = capi.generate_random_point_lights(random_lights, dims[:3], dims[3:])
    alias = capi.create_alias_table(point, tri)
    return SceneData(name, tris, tri_material, materials, nodes, capi.make_blob(point, 32), capi.make_blob(tri, 80),
                     capi.make_blob(alias, 16), dims, ref)


def with_random_point_lights(scene, count):
    """The north-star's many-light configs: generateRandomPointLights(count, scene dims)."""
    point = capi.generate_random_point_lights(count, scene.dims[:3], scene.dims[3:])
    alias = capi.create_alias_table(point, np.zeros((0, 80), np.uint8))
    return SceneData(f"{scene.name}+{count}pl", scene.triangles, scene.tri_material, scene.materials, scene.nodes,
                     capi.make_blob(point, 32), capi.make_blob(np.zeros((0, 80), np.uint8), 80), capi.make_blob(alias, 16), scene.dims,
                     scene.ref)


def with_point_lights(scene, positions, colours):
    """The scene lit by exactly these point lights (pointLight: vec4 pos, w = 1; vec4 colour, w = luminance —
    src/misc.cpp:343-356).  Degenerate lights are allowed: tests use zero and huge luminances and lights on surfaces."""
    positions, colours = np.asarray(positions, np.float32).reshape(-1, 3), np.asarray(colours, np.float32).reshape(-1, 3)
    rec = np.zeros((positions.shape[0], 8), np.float32)
    rec[:, 0:3], rec[:, 3] = positions, 1.0
    rec[:, 4:7] = colours
    rec[:, 7] = (np.float32(0.2126) * colours[:, 0] + np.float32(0.7152) * colours[:, 1]) + np.float32(0.0722) * colours[:, 2]  # common.glsl:7-9
    point = rec.view(np.uint8).reshape(-1, 32)
    alias = capi.create_alias_table(point, np.zeros((0, 80), np.uint8))
    return SceneData(f"{scene.name}+{positions.shape[0]}pl", scene.triangles, scene.tri_material, scene.materials, scene.nodes,
                     capi.make_blob(point, 32), capi.make_blob(np.zeros((0, 80), np.uint8), 80), capi.make_blob(alias, 16), scene.dims,
                     scene.ref)


# ---- what the G-buffer pass binds ------------------------------------------------------------------------

class GBufferPassInputs:
    """Vertices, indices, draws, matrices, material uniforms, texture bindings and textures in the reference's layouts
    (include/restir_layouts.h): the arguments of restir_upload_geometry / restir_upload_materials."""

    def __init__(self, vertices, indices, draws, matrices, uniforms, bindings, textures):
        self.vertices, self.indices, self.draws, self.matrices = vertices, indices, draws, matrices
        self.uniforms, self.bindings, self.textures = uniforms, bindings, textures

    def upload(self, ctx):
        ctx.upload_geometry(self.vertices, self.indices, self.draws, self.matrices)
        ctx.upload_materials(self.uniforms, self.bindings, self.textures)


def gbuffer_inputs_available(name):
    return os.path.exists(os.path.join(BAKED_DIR, name, "vertices.bin"))


def load_gbuffer_inputs(name):
    """scenes/_baked/<name>: what SceneBuffers / GBufferPass bind, dumped by oracle/_ref/scene_baker from the reference's own
    loader (textures reduced to <= 256 texels a side by the baker, to keep the snapshot small)."""
    d = os.path.join(BAKED_DIR, name)
    f = lambda n, dt: np.fromfile(os.path.join(d, n), dtype=dt)
    idx = f("textures.idx", np.uint32)
    texels = f("textures.rgba8", np.uint8)
    textures, first = [], 0
    for k in range(int(idx[0])):
        w, h = int(idx[1 + 2 * k]), int(idx[2 + 2 * k])
        textures.append(texels[first: first + w * h * 4].reshape(h, w, 4))
        first += w * h * 4
    return GBufferPassInputs(f("vertices.bin", np.uint8).reshape(-1, 80), f("indices.u32", np.uint32), f("draws.u32", np.uint32).reshape(-1, 4),
                             f("matrices.bin", np.uint8).reshape(-1, 128), f("material_uniforms.bin", np.uint8).reshape(-1, 64),
                             f("material_textures.i32", np.int32).reshape(-1, 4), textures)


def procedural_gbuffer_inputs(scene, seed=1, texture_size=32):
    """G-buffer pass inputs for a procedural scene (make_procedural): unshared vertices with smooth-ish normals, tangents and
    texture coordinates, one draw per run of equal material ids under a non-trivial model matrix pair, random textures, one
    alpha-masked material and one specular-glossiness material — every branch of gBuffer.frag gets pixels."""
    rng = np.random.default_rng(seed)
    tris = scene.triangles.view(np.float32).reshape(-1, 3, 4)[:, :, :3]
    n_tris = tris.shape[0]
    e1, e2 = tris[:, 1] - tris[:, 0], tris[:, 2] - tris[:, 0]
    fn = np.cross(e1, e2)
    fn /= np.maximum(np.linalg.norm(fn, axis=1, keepdims=True), 1e-20)
    verts = np.zeros((n_tris * 3, 20), np.float32)
    pos = tris.reshape(-1, 3)
    verts[:, 0:3] = pos
    verts[:, 3] = 1.0
    nrm = np.repeat(fn, 3, axis=0) + rng.uniform(-0.15, 0.15, (n_tris * 3, 3)).astype(np.float32)   # perturbed: interpolation matters
    verts[:, 4:7] = nrm / np.linalg.norm(nrm, axis=1, keepdims=True)
    tang = np.repeat(e1 / np.maximum(np.linalg.norm(e1, axis=1, keepdims=True), 1e-20), 3, axis=0)
    verts[:, 8:11] = tang
    verts[:, 11] = np.where(rng.uniform(size=n_tris * 3) < 0.5, -1.0, 1.0)
    verts[:, 12:16] = 1.0
    verts[:, 16:18] = pos[:, [0, 2]] * 0.37 + pos[:, [1]] * 0.21                                      # world-space planar mapping
    # draws: runs of equal material; the model matrix is the identity for positions' sake (the tree holds world-space
    # triangles) but the normal matrix is exercised with a non-uniform scale pair that cancels for normals: M = S, MIT = S^-T = S^-1
    mats = scene.tri_material
    draws, matrices = [], []
    t = 0
    while t < n_tris:
        e = t
        while e < n_tris and mats[e] == mats[t]:
            e += 1
        draws.append((t * 3, (e - t) * 3, 0, int(mats[t])))
        m = np.zeros((2, 4, 4), np.float32)
        m[0] = np.eye(4, dtype=np.float32)
        m[1] = np.eye(4, dtype=np.float32)
        matrices.append(m.reshape(-1))
        t = e
    indices = np.arange(n_tris * 3, dtype=np.uint32)
    n_mat = scene.materials.shape[0]
    uniforms = np.zeros((n_mat, 16), np.float32)
    for i, row in enumerate(scene.materials):
        uniforms[i, 0:4] = row[0:4]
        uniforms[i, 4:8] = row[4:8]
        uniforms[i, 8:11] = row[8:11]
        uniforms[i, 15] = 1.0 + 0.5 * (i % 3)                                                         # normalTextureScale
    uniforms_i = uniforms.view(np.int32)
    for i, row in enumerate(scene.materials):
        uniforms_i[i, 12] = int(row[11])
        uniforms_i[i, 13] = int(row[12])
        uniforms[i, 14] = row[13]
    if n_mat > 6:
        uniforms_i[5, 13] = 1                                                                         # ALPHA_MODE_MASK, cutoff 0.5: holes from the texture's alpha
        uniforms[5, 14] = 0.5
        uniforms_i[6, 12] = 1                                                                         # specular-glossiness
        uniforms[6, 4:8] = (0.6, 0.5, 0.4, 0.7)
    textures = []
    for k in range(6):
        t_ = rng.integers(0, 256, (texture_size, texture_size * (1 + k % 2), 4), dtype=np.uint8)
        if k == 1:
            t_[..., 0:2] = rng.integers(96, 160, t_[..., 0:2].shape, dtype=np.uint8)                  # a plausible normal map
            t_[..., 2] = 255
        textures.append(t_)
    bindings = np.full((n_mat, 4), -1, np.int32)
    for i in range(n_mat):
        if i % 2 == 0:
            bindings[i] = (0 + (i % 3) * 2 % 6, 1, 2, 3)
    if n_mat > 5:
        bindings[5] = (4, -1, -1, -1)
    return GBufferPassInputs(verts.view(np.uint8).reshape(-1, 80), indices, np.asarray(draws, np.uint32), np.asarray(matrices, np.float32).view(np.uint8).reshape(-1, 128),
                             uniforms.view(np.uint8).reshape(-1, 64), bindings, textures)


# ---- material -> G-buffer codes (src/shaders/gBuffer.frag:27-79 with all textures = 1) -------------

def _srgb_encode8(c):
    c = np.clip(np.asarray(c, np.float64), 0.0, 1.0)
    s = np.where(c <= 0.0031308, 12.92 * c, 1.055 * np.power(c, 1.0 / 2.4) - 0.055)
    return np.rint(s * 255.0).astype(np.uint32)


def material_table(materials):
    """Per material {albedo RGBA8 (rgb sRGB-encoded), material RG16, flags, 0} as uint32[M][4]."""
    m = np.asarray(materials, np.float64).reshape(-1, 16)
    out = np.zeros((m.shape[0], 4), np.uint32)
    for i, row in enumerate(m):
        color, param, emissive = row[0:4], row[4:8], row[8:11]
        shading, alpha_mode, cutoff = int(row[11]), int(row[12]), row[13]
        albedo = color[:3].copy()
        if shading == 0:                      # SHADING_MODEL_METALLIC_ROUGHNESS, gBuffer.frag:47-50
            roughness, metallic = param[1], param[2]
        else:                                 # SHADING_MODEL_SPECULAR_GLOSSINESS, :51-67
            roughness = 1.0 - param[3]
            average = 0.5 * (albedo + param[:3])
            sqrt_term = np.sqrt(np.maximum(average * average - 0.04 * albedo, 0.0))
            metallic = float(np.mean(25.0 * average - sqrt_term))
            albedo = average + sqrt_term
        if np.linalg.norm(emissive) > 0.0:    # :73-79
            albedo = color[:3] * emissive
            a = 255
        else:
            a = 0
        rgb = _srgb_encode8(albedo)
        rg = np.rint(np.clip([roughness, metallic], 0.0, 1.0) * 65535.0).astype(np.uint32)
        out[i, 0] = rgb[0] | (rgb[1] << 8) | (rgb[2] << 16) | (a << 24)
        out[i, 1] = rg[0] | (rg[1] << 16)
        out[i, 2] = 1 if (alpha_mode == 1 and color[3] < cutoff) else 0   # ALPHA_MODE_MASK discard, :30-34
    return out


# ---- procedural scenes -----------------------------------------------------------------------------

def _quad(p, u, v):
    """Two CCW triangles of the parallelogram p, p+u, p+u+v, p+v (normal = u x v)."""
    a, b, c, d = p, p + u, p + u + v, p + v
    return [np.concatenate([a, b, c]), np.concatenate([a, c, d])]


def _box(lo, hi):
    """12 triangles, outward normals."""
    lo, hi = np.asarray(lo, np.float32), np.asarray(hi, np.float32)
    s = hi - lo
    ex, ey, ez = np.array([s[0], 0, 0], np.float32), np.array([0, s[1], 0], np.float32), np.array([0, 0, s[2]], np.float32)
    t = []
    t += _quad(lo, ey, ex)                 # -z
    t += _quad(lo + ez, ex, ey)            # +z
    t += _quad(lo, ez, ey)                 # -x
    t += _quad(lo + ex, ey, ez)            # +x
    t += _quad(lo, ex, ez)                 # -y
    t += _quad(lo + ey, ez, ex)            # +y
    return t


def _grid(p, u, v, nu, nv, rng=None, jitter=0.0, normal=None):
    """nu x nv quads tiling the parallelogram; optional jitter along `normal` (keeps it watertight)."""
    pts = np.zeros((nu + 1, nv + 1, 3), np.float32)
    for i in range(nu + 1):
        for j in range(nv + 1):
            q = p + u * (i / nu) + v * (j / nv)
            if rng is not None and jitter > 0 and 0 < i < nu and 0 < j < nv:
                q = q + normal * np.float32(rng.uniform(-jitter, jitter))
            pts[i, j] = q
    t = []
    for i in range(nu):
        for j in range(nv):
            a, b, c, d = pts[i, j], pts[i + 1, j], pts[i + 1, j + 1], pts[i, j + 1]
            t.append(np.concatenate([a, b, c]))
            t.append(np.concatenate([a, c, d]))
    return t


def procedural_scene(seed=1, grid=12, boxes=24, lights="point", n_point_lights=16, name=None):
    """A closed room (inward-facing jittered walls) with random boxes inside.

    lights = "point": n_point_lights file lights; "tri": emissive ceiling panels (triangle lights);
    "random": neither, so the 200-random-light fallback of the reference applies.
    Returns (tris (T,9) f32, tri_material (T,) i32, materials (M,16) f32, file_point_lights (L,32) u8).
    """
    rng = np.random.default_rng(seed)
    L = np.float32(5.0)
    X, Y, Z = np.eye(3, dtype=np.float32)
    groups = []  # (material id, [triangles])
    # room: normals point inwards
    groups.append((0, _grid(np.array([-L, -L, -L], np.float32), 2 * L * Z, 2 * L * X, grid, grid, rng, 0.08, Y)))       # floor  (+y)
    groups.append((1, _grid(np.array([-L, L, -L], np.float32), 2 * L * X, 2 * L * Z, grid, grid, rng, 0.08, Y)))        # ceiling (-y)
    groups.append((2, _grid(np.array([-L, -L, -L], np.float32), 2 * L * Y, 2 * L * Z, grid, grid, rng, 0.08, X)))       # -x wall (+x)
    groups.append((3, _grid(np.array([L, -L, -L], np.float32), 2 * L * Z, 2 * L * Y, grid, grid, rng, 0.08, X)))        # +x wall (-x)
    groups.append((0, _grid(np.array([-L, -L, -L], np.float32), 2 * L * X, 2 * L * Y, grid, grid, rng, 0.08, Z)))       # -z wall (+z)
    groups.append((4, _grid(np.array([-L, -L, L], np.float32), 2 * L * Y, 2 * L * X, grid, grid, rng, 0.08, Z)))        # +z wall (-z)
    for b in range(boxes):
        c = rng.uniform(-3.8, 3.8, 3).astype(np.float32)
        c[1] = np.float32(rng.uniform(-4.6, 0.5))
        h = rng.uniform(0.25, 0.9, 3).astype(np.float32)
        groups.append((5 + (b % 3), _box(c - h, c + h)))
    n_materials = 9
    if lights == "tri":
        for k in range(6):
            cx, cz = np.float32(-3.0 + 3.0 * (k % 3)), np.float32(-2.0 + 4.0 * (k // 3))
            p = np.array([cx - 0.6, L - 0.3, cz - 0.6], np.float32)
            groups.append((8, _quad(p, np.float32(1.2) * X, np.float32(1.2) * Z)))   # facing down (-y)
    tris, tri_material = [], []
    for mat, ts in groups:
        tris += ts
        tri_material += [mat] * len(ts)
    tris = np.asarray(tris, np.float32).reshape(-1, 9)
    tri_material = np.asarray(tri_material, np.int32)

    materials = np.zeros((n_materials, 16), np.float32)
    base = [(0.8, 0.8, 0.8), (0.9, 0.9, 0.85), (0.8, 0.1, 0.1), (0.1, 0.7, 0.15), (0.2, 0.3, 0.8), (0.9, 0.6, 0.2), (0.6, 0.6, 0.65),
            (0.95, 0.93, 0.88), (1.0, 1.0, 1.0)]
    rough = [0.9, 1.0, 0.7, 0.7, 0.5, 0.35, 0.15, 0.6, 1.0]
    metal = [0.0, 0.0, 0.0, 0.0, 0.1, 0.0, 1.0, 0.5, 0.0]
    for i in range(n_materials):
        materials[i, 0:3] = base[i]
        materials[i, 3] = 1.0
        materials[i, 5] = rough[i]
        materials[i, 6] = metal[i]
        materials[i, 13] = 0.5
    if lights == "tri":
        materials[8, 8:11] = (8.0, 7.0, 6.0)
    file_lights = np.zeros((0, 32), np.uint8)
    if lights == "point":
        pl = np.zeros((n_point_lights, 8), np.float32)
        pl[:, 0:3] = rng.uniform(-4.2, 4.2, (n_point_lights, 3))
        pl[:, 3] = 1.0
        pl[:, 4:7] = rng.uniform(0.2, 6.0, (n_point_lights, 3))
        pl[:, 7] = 0.2126 * pl[:, 4] + 0.7152 * pl[:, 5] + 0.0722 * pl[:, 6]   # recomputed in fp32 below
        pl[:, 7] = (np.float32(0.2126) * pl[:, 4] + np.float32(0.7152) * pl[:, 5]) + np.float32(0.0722) * pl[:, 6]
        file_lights = pl.view(np.uint8).reshape(-1, 32)
    return tris, tri_material, materials, file_lights


def soup_to_triangles48(tris9):
    """(T,9) f32 -> (T,48) u8 reference Triangle records (w = 1)."""
    t = np.ones((tris9.shape[0], 3, 4), np.float32)
    t[:, :, :3] = tris9.reshape(-1, 3, 3)
    return t.reshape(-1, 12).view(np.uint8).reshape(-1, 48)


def scene_dims(tris9):
    p = tris9.reshape(-1, 3)
    return np.concatenate([p.min(0), p.max(0)]).astype(np.float32)


def make_procedural(seed=1, grid=12, boxes=24, lights="point", n_point_lights=16):
    tris9, tri_material, materials, file_lights = procedural_scene(seed, grid, boxes, lights, n_point_lights)
    return assemble_scene(f"procedural(seed={seed},grid={grid},boxes={boxes},lights={lights})", soup_to_triangles48(tris9), tri_material,
                          materials, file_lights, scene_dims(tris9))


def write_soup(path, tris9, tri_material, materials, file_lights):
    """Input format of oracle/_ref/scene_baker's `soup` mode."""
    lights = np.ascontiguousarray(file_lights).view(np.float32).reshape(-1, 8)
    with open(path, "wb") as f:
        f.write(np.array([0x50554F53, tris9.shape[0], materials.shape[0], lights.shape[0]], np.uint32).tobytes())
        f.write(np.ascontiguousarray(tris9, np.float32).tobytes())
        f.write(np.ascontiguousarray(tri_material, np.int32).tobytes())
        f.write(np.ascontiguousarray(materials, np.float32).tobytes())
        f.write(lights.tobytes())
