// scene_baker — runs the REFERENCE's own CPU scene code (never a restatement) and dumps what it
// produces, so the repo's builders and the oracle can be pinned against real reference output.
//
// TEST/BENCH INFRASTRUCTURE.  Built only where /root/reference exists, by oracle/ref_build/Makefile,
// into the git-ignored oracle/_ref/.  It links, unmodified and from where they lie:
//   thirdparty/gltf/gltfscene.cpp, mikktWrapper.cpp, thirdparty/MikkTSpace/mikktspace.c, tiny_gltf.h
//   src/aabbTreeBuilder.cpp          (AabbTree::build, reference src/aabbTreeBuilder.cpp:52-214)
//   src/misc.cpp:343-497             (collectPointLightsFromScene, generateRandomPointLights,
//                                     collectTriangleLightsFromScene, createAliasTable)
// This file only does what the reference's App does around them:
//   loadScene                         src/misc.cpp:309-341 (texture decode skipped: not needed)
//   light selection + alias table     src/sceneBuffers.h:78-84
//   SSBO blob packing                 src/sceneBuffers.h:100-124, 241-270 (count @0, array @16)
//   material parameter packing        src/sceneBuffers.h:205-222
//   G-buffer pass inputs              src/sceneBuffers.h:126-233 (vertices, indices, matrices, material uniforms, textures),
//                                     src/passes/gBufferPass.cpp:132-152 (draws), :193-247 (which texture a material binds)
//
// usage: scene_baker gltf <file.gltf> <outdir> [texture limit: also dump the G-buffer pass inputs, textures reduced to <= limit texels a side]
//        scene_baker soup <file.soup> <outdir>      (procedural triangle soup, see tests/scenes.py)
//        scene_baker randlights <n> <minx miny minz maxx maxy maxz> <out.bin>

#define TINYGLTF_IMPLEMENTATION
#define STB_IMAGE_IMPLEMENTATION
#define STB_IMAGE_WRITE_IMPLEMENTATION
#include <tiny_gltf.h>
#undef TINYGLTF_IMPLEMENTATION
#undef STB_IMAGE_IMPLEMENTATION
#undef STB_IMAGE_WRITE_IMPLEMENTATION

#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <queue>
#include <random>
#include <string>

#include <gltfscene.h>
#include "aabbTreeBuilder.h"

// declarations as in reference src/misc.h:132-142
[[nodiscard]] std::vector<shader::pointLight> collectPointLightsFromScene(const nvh::GltfScene&);
[[nodiscard]] std::vector<shader::pointLight> generateRandomPointLights(
	std::size_t count, nvmath::vec3 min, nvmath::vec3 max,
	std::uniform_real_distribution<float> distR = std::uniform_real_distribution<float>(0.0f, 1.0f),
	std::uniform_real_distribution<float> distG = std::uniform_real_distribution<float>(0.0f, 1.0f),
	std::uniform_real_distribution<float> distB = std::uniform_real_distribution<float>(0.0f, 1.0f)
);
[[nodiscard]] std::vector<shader::triLight> collectTriangleLightsFromScene(const nvh::GltfScene&);
[[nodiscard]] std::vector<shader::aliasTableColumn> createAliasTable(
	std::vector<shader::pointLight>& ptLights, std::vector<shader::triLight>& triLights);

static void writeFile(const std::string &path, const void *data, std::size_t bytes) {
	std::ofstream f(path, std::ios::binary);
	f.write(static_cast<const char*>(data), static_cast<std::streamsize>(bytes));
	if (!f) {
		std::cerr << "scene_baker: cannot write " << path << "\n";
		std::exit(2);
	}
}

template <typename T> static void writeBlob(const std::string &path, const std::vector<T> &items) {
	// {int32 count; pad to 16; T[count]}  — sceneBuffers.h:100-124 sizes, :241-270 contents
	std::vector<unsigned char> blob(16 + sizeof(T) * items.size(), 0);
	int32_t count = static_cast<int32_t>(items.size());
	std::memcpy(blob.data(), &count, 4);
	if (!items.empty()) {
		std::memcpy(blob.data() + 16, items.data(), sizeof(T) * items.size());
	}
	writeFile(path, blob.data(), blob.size());
}

static bool skipImage(tinygltf::Image*, const int, std::string*, std::string*, int, int,
                      const unsigned char*, int, void*) {
	return true; // scenes baked without their textures
}

// ---- what GBufferPass binds (src/passes/gBufferPass.cpp:116-157), for the G-buffer producer and its oracle twin -------
struct BakedVertex { // src/vertex.h
	nvmath::vec4 position, normal, tangent, color;
	nvmath::vec2 uv;
};
static_assert(sizeof(BakedVertex) == 80, "Vertex is 80 bytes");

// Textures larger than `limit` texels on a side are halved (2x2 box, rounded) until they fit: the snapshot that travels to
// the GPU box stays small.  The producer samples whatever level it is given as level 0.
static void dumpGBufferInputs(const nvh::GltfScene &scene, const std::string &out, int textureLimit) {
	// vertices, sceneBuffers.h:173-193
	std::vector<BakedVertex> vertices(scene.m_positions.size());
	std::memset(vertices.data(), 0, vertices.size() * sizeof(BakedVertex));
	for (std::size_t i = 0; i < scene.m_positions.size(); ++i) {
		BakedVertex &v = vertices[i];
		v.position = scene.m_positions[i];
		if (i < scene.m_normals.size()) {
			v.normal = scene.m_normals[i];
		}
		if (i < scene.m_colors0.size()) {
			v.color = scene.m_colors0[i];
		} else {
			v.color = nvmath::vec4(1.0f, 0.0f, 1.0f, 1.0f);
		}
		if (i < scene.m_texcoords0.size()) {
			v.uv = scene.m_texcoords0[i];
		}
		if (i < scene.m_tangents.size()) {
			v.tangent = scene.m_tangents[i];
		}
	}
	writeFile(out + "/vertices.bin", vertices.data(), vertices.size() * sizeof(BakedVertex));
	writeFile(out + "/indices.u32", scene.m_indices.data(), scene.m_indices.size() * 4);

	// draws (gBufferPass.cpp:132-152) and matrices (sceneBuffers.h:225-230)
	std::vector<uint32_t> draws;
	std::vector<shader::ModelMatrices> matrices(scene.m_nodes.size());
	for (std::size_t i = 0; i < scene.m_nodes.size(); ++i) {
		const nvh::GltfPrimMesh &mesh = scene.m_primMeshes[scene.m_nodes[i].primMesh];
		draws.push_back(mesh.firstIndex);
		draws.push_back(mesh.indexCount);
		draws.push_back(mesh.vertexOffset);
		draws.push_back(static_cast<uint32_t>(mesh.materialIndex));
		matrices[i].transform = scene.m_nodes[i].worldMatrix;
		matrices[i].transformInverseTransposed = nvmath::transpose(nvmath::invert(matrices[i].transform));
	}
	static_assert(sizeof(shader::ModelMatrices) == 128 && sizeof(shader::MaterialUniforms) == 64);
	writeFile(out + "/draws.u32", draws.data(), draws.size() * 4);
	writeFile(out + "/matrices.bin", matrices.data(), matrices.size() * sizeof(shader::ModelMatrices));

	// material uniforms (sceneBuffers.h:205-222; the mapped memory the reference leaves untouched reads as zero here) and
	// the texture each binding of a material's descriptor set gets (gBufferPass.cpp:200-247; -1 = default white / default normal)
	std::vector<shader::MaterialUniforms> uniforms(scene.m_materials.size());
	std::memset(uniforms.data(), 0, uniforms.size() * sizeof(shader::MaterialUniforms));
	std::vector<int32_t> bindings;
	for (std::size_t i = 0; i < scene.m_materials.size(); ++i) {
		const nvh::GltfMaterial &mat = scene.m_materials[i];
		shader::MaterialUniforms &outMat = uniforms[i];
		outMat.emissiveFactor = mat.emissiveFactor;
		outMat.shadingModel = mat.shadingModel;
		outMat.alphaMode = mat.alphaMode;
		outMat.alphaCutoff = mat.alphaCutoff;
		outMat.normalTextureScale = mat.normalTextureScale;
		switch (outMat.shadingModel) {
		case SHADING_MODEL_METALLIC_ROUGHNESS:
			outMat.colorParam = mat.pbrBaseColorFactor;
			outMat.materialParam.y = mat.pbrRoughnessFactor;
			outMat.materialParam.z = mat.pbrMetallicFactor;
			bindings.push_back(mat.pbrBaseColorTexture);
			bindings.push_back(mat.normalTexture);
			bindings.push_back(mat.pbrMetallicRoughnessTexture);
			break;
		case SHADING_MODEL_SPECULAR_GLOSSINESS:
			outMat.colorParam = mat.khrDiffuseFactor;
			outMat.materialParam = mat.khrSpecularFactor;
			outMat.materialParam.w = mat.khrGlossinessFactor;
			bindings.push_back(mat.khrDiffuseTexture);
			bindings.push_back(mat.normalTexture);
			bindings.push_back(mat.khrSpecularGlossinessTexture);
			break;
		default:
			bindings.insert(bindings.end(), {-1, -1, -1});
		}
		bindings.push_back(mat.emissiveTexture);
	}
	writeFile(out + "/material_uniforms.bin", uniforms.data(), uniforms.size() * sizeof(shader::MaterialUniforms));
	writeFile(out + "/material_textures.i32", bindings.data(), bindings.size() * 4);

	// textures: RGBA8 as the reference's loader decodes them (tinygltf + stb_image, 4 components), reduced to <= textureLimit
	std::vector<uint32_t> index;
	std::vector<unsigned char> pixels;
	index.push_back(static_cast<uint32_t>(scene.m_textures.size()));
	for (const tinygltf::Image &img : scene.m_textures) {
		int w = img.width, h = img.height;
		std::vector<unsigned char> cur(img.image.begin(), img.image.end());
		if (img.component != 4 || img.bits != 8 || cur.size() != std::size_t(w) * h * 4) {
			std::cerr << "scene_baker: texture " << img.uri << " is not RGBA8 after decoding\n";
			std::exit(3);
		}
		while (std::max(w, h) > textureLimit && w % 2 == 0 && h % 2 == 0) {
			std::vector<unsigned char> half(std::size_t(w / 2) * (h / 2) * 4);
			for (int y = 0; y < h / 2; ++y) {
				for (int x = 0; x < w / 2; ++x) {
					for (int c = 0; c < 4; ++c) {
						unsigned s = cur[(std::size_t(2 * y) * w + 2 * x) * 4 + c] + cur[(std::size_t(2 * y) * w + 2 * x + 1) * 4 + c] +
						             cur[(std::size_t(2 * y + 1) * w + 2 * x) * 4 + c] + cur[(std::size_t(2 * y + 1) * w + 2 * x + 1) * 4 + c];
						half[(std::size_t(y) * (w / 2) + x) * 4 + c] = static_cast<unsigned char>((s + 2) / 4);
					}
				}
			}
			cur.swap(half);
			w /= 2;
			h /= 2;
		}
		index.push_back(static_cast<uint32_t>(w));
		index.push_back(static_cast<uint32_t>(h));
		pixels.insert(pixels.end(), cur.begin(), cur.end());
	}
	writeFile(out + "/textures.idx", index.data(), index.size() * 4);
	writeFile(out + "/textures.rgba8", pixels.data(), pixels.size());
}

static void dumpScene(nvh::GltfScene &scene, const std::string &out) {
	AabbTree tree = AabbTree::build(scene);

	std::vector<shader::pointLight> pointLights = collectPointLightsFromScene(scene);
	std::vector<shader::triLight> triangleLights = collectTriangleLightsFromScene(scene);
	std::vector<shader::pointLight> fileLights = pointLights;
	if (pointLights.empty() && triangleLights.empty()) {
		pointLights = generateRandomPointLights(200, scene.m_dimensions.min, scene.m_dimensions.max);
	}
	std::vector<shader::aliasTableColumn> aliasTable = createAliasTable(pointLights, triangleLights);

	// inputs for the repo's own builders: triangles in the reference's order + material ids
	std::vector<int32_t> triMaterial;
	triMaterial.reserve(tree.triangles.size());
	for (const nvh::GltfNode &node : scene.m_nodes) {
		const nvh::GltfPrimMesh &mesh = scene.m_primMeshes[node.primMesh];
		for (uint32_t i = 0; i < mesh.indexCount; i += 3) {
			triMaterial.push_back(mesh.materialIndex);
		}
	}
	std::vector<float> materials;
	for (const nvh::GltfMaterial &mat : scene.m_materials) {
		float m[16] = {};
		if (mat.shadingModel == 0) {
			std::memcpy(m, &mat.pbrBaseColorFactor.x, 16);
			m[5] = mat.pbrRoughnessFactor;
			m[6] = mat.pbrMetallicFactor;
		} else {
			std::memcpy(m, &mat.khrDiffuseFactor.x, 16);
			m[4] = mat.khrSpecularFactor.x;
			m[5] = mat.khrSpecularFactor.y;
			m[6] = mat.khrSpecularFactor.z;
			m[7] = mat.khrGlossinessFactor;
		}
		m[8] = mat.emissiveFactor.x;
		m[9] = mat.emissiveFactor.y;
		m[10] = mat.emissiveFactor.z;
		m[11] = static_cast<float>(mat.shadingModel);
		m[12] = static_cast<float>(mat.alphaMode);
		m[13] = mat.alphaCutoff;
		materials.insert(materials.end(), m, m + 16);
	}
	float dims[6] = {
		scene.m_dimensions.min.x, scene.m_dimensions.min.y, scene.m_dimensions.min.z,
		scene.m_dimensions.max.x, scene.m_dimensions.max.y, scene.m_dimensions.max.z
	};

	static_assert(sizeof(shader::Triangle) == 48 && sizeof(shader::AabbTreeNode) == 80);
	static_assert(sizeof(shader::pointLight) == 32 && sizeof(shader::triLight) == 80);
	static_assert(sizeof(shader::aliasTableColumn) == 16);

	// padding bytes of AabbTreeNode (72..79) are indeterminate in the reference; zero them in the dump
	std::vector<unsigned char> nodeBytes(tree.nodes.size() * 80, 0);
	for (std::size_t i = 0; i < tree.nodes.size(); ++i) {
		std::memcpy(nodeBytes.data() + i * 80, &tree.nodes[i], 72);
	}

	writeFile(out + "/triangles.bin", tree.triangles.data(), tree.triangles.size() * 48);
	writeFile(out + "/tri_material.i32", triMaterial.data(), triMaterial.size() * 4);
	writeFile(out + "/materials.f32", materials.data(), materials.size() * 4);
	writeFile(out + "/dims.f32", dims, sizeof(dims));
	writeFile(out + "/ref_nodes.bin", nodeBytes.data(), nodeBytes.size());
	writeBlob(out + "/gltf_point_lights.bin", fileLights);
	writeBlob(out + "/ref_point_lights.bin", pointLights);
	writeBlob(out + "/ref_tri_lights.bin", triangleLights);
	writeBlob(out + "/ref_alias.bin", aliasTable);

	std::printf(
		"{\"triangles\": %zu, \"nodes\": %zu, \"materials\": %zu, \"point_lights\": %zu, "
		"\"tri_lights\": %zu, \"alias\": %zu, \"drawable_nodes\": %zu}\n",
		tree.triangles.size(), tree.nodes.size(), scene.m_materials.size(), pointLights.size(),
		triangleLights.size(), aliasTable.size(), scene.m_nodes.size()
	);
}

static int bakeGltf(const std::string &filename, const std::string &out, int textureLimit) {
	tinygltf::Model tmodel;
	tinygltf::TinyGLTF tcontext;
	std::string warn, error;
	if (textureLimit <= 0) {
		tcontext.SetImageLoader(skipImage, nullptr);
	}
	if (!tcontext.LoadASCIIFromFile(&tmodel, &error, &warn, filename)) {
		std::cerr << "scene_baker: cannot load " << filename << ": " << error << "\n";
		return 1;
	}
	nvh::GltfScene scene;
	// same attribute mask and call order as loadScene (src/misc.cpp:317-318)
	scene.importDrawableNodes(
		tmodel,
		nvh::GltfAttributes::Normal | nvh::GltfAttributes::Texcoord_0 |
		nvh::GltfAttributes::Color_0 | nvh::GltfAttributes::Tangent
	);
	scene.importMaterials(tmodel);
	if (textureLimit > 0) {
		scene.importTexutureImages(tmodel); // src/misc.cpp:319
	}
	dumpScene(scene, out);
	if (textureLimit > 0) {
		dumpGBufferInputs(scene, out, textureLimit);
	}
	return 0;
}

// soup file: u32 'SOUP', u32 nTris, u32 nMaterials, u32 nPointLights,
//            f32 tris[nTris][9], i32 triMaterial[nTris], f32 materials[nMaterials][16], f32 lights[n][8]
static int bakeSoup(const std::string &filename, const std::string &out) {
	std::ifstream f(filename, std::ios::binary);
	uint32_t hdr[4];
	f.read(reinterpret_cast<char*>(hdr), 16);
	if (!f || hdr[0] != 0x50554F53u) {
		std::cerr << "scene_baker: bad soup file\n";
		return 1;
	}
	uint32_t nTris = hdr[1], nMat = hdr[2], nLights = hdr[3];
	std::vector<float> tris(std::size_t(nTris) * 9), mats(std::size_t(nMat) * 16), lights(std::size_t(nLights) * 8);
	std::vector<int32_t> triMat(nTris);
	f.read(reinterpret_cast<char*>(tris.data()), tris.size() * 4);
	f.read(reinterpret_cast<char*>(triMat.data()), triMat.size() * 4);
	f.read(reinterpret_cast<char*>(mats.data()), mats.size() * 4);
	f.read(reinterpret_cast<char*>(lights.data()), lights.size() * 4);
	if (!f) {
		std::cerr << "scene_baker: truncated soup file\n";
		return 1;
	}

	nvh::GltfScene scene;
	for (uint32_t m = 0; m < nMat; ++m) {
		const float *s = &mats[std::size_t(m) * 16];
		nvh::GltfMaterial mat;
		mat.shadingModel = static_cast<int>(s[11]);
		if (mat.shadingModel == 0) {
			mat.pbrBaseColorFactor = nvmath::vec4(s[0], s[1], s[2], s[3]);
			mat.pbrRoughnessFactor = s[5];
			mat.pbrMetallicFactor = s[6];
		} else {
			mat.khrDiffuseFactor = nvmath::vec4(s[0], s[1], s[2], s[3]);
			mat.khrSpecularFactor = nvmath::vec3(s[4], s[5], s[6]);
			mat.khrGlossinessFactor = s[7];
		}
		mat.emissiveFactor = nvmath::vec3(s[8], s[9], s[10]);
		mat.alphaMode = static_cast<int>(s[12]);
		mat.alphaCutoff = s[13];
		scene.m_materials.push_back(mat);
	}
	// one drawable node per run of equal material ids, identity transform, unshared vertices
	for (uint32_t t = 0; t < nTris; ) {
		uint32_t e = t;
		while (e < nTris && triMat[e] == triMat[t]) {
			++e;
		}
		nvh::GltfPrimMesh mesh;
		mesh.firstIndex = static_cast<uint32_t>(scene.m_indices.size());
		mesh.vertexOffset = static_cast<uint32_t>(scene.m_positions.size());
		mesh.indexCount = (e - t) * 3;
		mesh.vertexCount = (e - t) * 3;
		mesh.materialIndex = triMat[t];
		nvmath::vec3f mn(std::numeric_limits<float>::max()), mx(-std::numeric_limits<float>::max());
		for (uint32_t i = t; i < e; ++i) {
			for (int v = 0; v < 3; ++v) {
				const float *p = &tris[std::size_t(i) * 9 + v * 3];
				nvmath::vec3f pos(p[0], p[1], p[2]);
				scene.m_indices.push_back(static_cast<uint32_t>((i - t) * 3 + v));
				scene.m_positions.push_back(pos);
				mn = nvmath::nv_min(mn, pos);
				mx = nvmath::nv_max(mx, pos);
			}
		}
		mesh.posMin = mn;
		mesh.posMax = mx;
		scene.m_primMeshes.push_back(mesh);
		nvh::GltfNode node;
		node.primMesh = static_cast<int>(scene.m_primMeshes.size() - 1);
		scene.m_nodes.push_back(node);
		t = e;
	}
	for (uint32_t l = 0; l < nLights; ++l) {
		const float *s = &lights[std::size_t(l) * 8];
		nvh::GltfLight light;
		light.worldMatrix.as_translation(nvmath::vec3f(s[0], s[1], s[2]));
		light.light.color = { s[4], s[5], s[6] };
		scene.m_lights.push_back(light);
	}
	scene.computeSceneDimensions();
	dumpScene(scene, out);
	return 0;
}

int main(int argc, char **argv) {
	if ((argc == 4 || argc == 5) && std::string(argv[1]) == "gltf") {
		return bakeGltf(argv[2], argv[3], argc == 5 ? std::stoi(argv[4]) : 0);
	}
	if (argc == 4 && std::string(argv[1]) == "soup") {
		return bakeSoup(argv[2], argv[3]);
	}
	if (argc == 10 && std::string(argv[1]) == "randlights") {
		std::size_t n = std::stoull(argv[2]);
		nvmath::vec3 mn(std::stof(argv[3]), std::stof(argv[4]), std::stof(argv[5]));
		nvmath::vec3 mx(std::stof(argv[6]), std::stof(argv[7]), std::stof(argv[8]));
		writeBlob(argv[9], generateRandomPointLights(n, mn, mx));
		return 0;
	}
	std::cerr << "usage: scene_baker gltf <file.gltf> <outdir> [texture limit] | soup <file.soup> <outdir> | "
	             "randlights <n> <min xyz> <max xyz> <out.bin>\n";
	return 64;
}
