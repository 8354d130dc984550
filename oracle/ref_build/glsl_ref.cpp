// glsl_ref.cpp — runs the reference's OWN shader sources on the CPU (TEST INFRASTRUCTURE; built into the
// git-ignored oracle/_ref/libglslref.so only where /root/reference exists, see Makefile).
//
// The four shaders of the path are compiled from where they lie — restirOmniSoftware.comp (-> restirOmni.glsl),
// spatialReuse.comp, unbiasedReuseSoftware.comp (-> unbiasedReuse.glsl), lighting.frag and everything they
// #include — after glsl2cpp.py's token-level transliteration into oracle/_ref/glsl/ (float-literal suffixes,
// inout -> reference, interface blocks -> globals; no arithmetic is touched).  This file binds those globals to
// the caller's arrays the way the reference's descriptor sets do (restirPass.h:120-237, spatialReusePass.h:28-89,
// unbiasedReusePass.h:111-196, lightingPass.h:48-96) and dispatches main() once per pixel
// (restirPass.h:52-57, spatialReusePass.h:18-24, unbiasedReusePass.h:38-44; lighting is a full-screen triangle).
//
// Same C entry points as oracle/restir_oracle.cpp (oracle_* -> glslref_*), so tests/test_oracle_vs_glsl.py can
// run both on the same inputs and demand bit-identical reservoirs.

#include "glsl_shim.h"

#ifdef _OPENMP
#	include <omp.h>
#endif

#define main shader_main

namespace glsl {
namespace omni {
#include "glsl/restirOmniSoftware.comp"
}
namespace spatial {
#include "glsl/spatialReuse.comp"
}
namespace unbiased3 { // unbiasedReuse.glsl:47 as shipped: NUM_NEIGHBORS 3
#include "glsl/unbiasedReuseSoftware.comp"
}
#undef NUM_NEIGHBORS
namespace unbiased5 { // the same source with NUM_NEIGHBORS 5 (north-star configuration)
#include "glsl_n5/unbiasedReuseSoftware.comp"
}
namespace lighting {
#include "glsl/lighting.frag"
}
} // namespace glsl

#undef main

using namespace glsl;

// layouts the bindings below rely on
// (the variant builds — RESERVOIR_SIZE / UNBIASED_MIS set in the transliterated restirStructs.glsl — state their own sizes)
#ifndef GLSLREF_RESERVOIR_BYTES
#	define GLSLREF_RESERVOIR_BYTES 64
#	define GLSLREF_SAMPLE_BYTES 48
#endif
static_assert(sizeof(omni::Reservoir) == GLSLREF_RESERVOIR_BYTES && sizeof(omni::LightSample) == GLSLREF_SAMPLE_BYTES, "Reservoir layout");
static_assert(sizeof(spatial::Reservoir) == GLSLREF_RESERVOIR_BYTES && sizeof(unbiased3::Reservoir) == GLSLREF_RESERVOIR_BYTES &&
                  sizeof(unbiased5::Reservoir) == GLSLREF_RESERVOIR_BYTES && sizeof(lighting::Reservoir) == GLSLREF_RESERVOIR_BYTES,
              "Reservoir layout");
extern "C" int glslref_reservoir_bytes(void) { return GLSLREF_RESERVOIR_BYTES; }
static_assert(sizeof(omni::RestirUniforms) == 128, "RestirUniforms layout");
static_assert(sizeof(lighting::LightingPassUniforms) == 96, "LightingPassUniforms layout");
static_assert(sizeof(omni::AabbTreeNode) == 80 && sizeof(omni::Triangle) == 48, "AabbTree layout");
static_assert(sizeof(omni::pointLight) == 32 && sizeof(omni::triLight) == 80 && sizeof(omni::aliasTableColumn) == 16, "light layouts");

extern "C" {

struct glslref_gbuffer {
	const void *albedo, *normal, *material, *worldPos, *depth;
};
struct glslref_scene {
	const void *nodes, *tris, *pointBlob, *triBlob, *aliasBlob;
};

} // extern "C"

namespace {

const int kBlobHeader = 16; // sceneBuffers.h:100-124: {int32 count; pad to 16; T[count]}

int blobCount(const void *blob) {
	int n = 0;
	if (blob) {
		std::memcpy(&n, blob, 4);
	}
	return n;
}
template <class T> T *blobItems(const void *blob) { return blob ? (T *)((const char *)blob + kBlobHeader) : nullptr; }

sampler2D tex(TexelFormat f, const void *p, int w, int h) { return sampler2D{f, p, w, h}; }

template <class F> void forEachPixel(int w, int y0, int y1, F f) {
#pragma omp parallel for schedule(dynamic, 1)
	for (int y = y0; y < y1; ++y) {
		for (int x = 0; x < w; ++x) {
			gl_GlobalInvocationID.xy = uvec2((uint)x, (uint)y);
			f(x, y);
		}
	}
}

} // namespace

extern "C" {

// restirOmni.glsl: set0 {0 point, 1 tri, 2 alias, 3 uniforms}; set1 {0 worldPos, 1 albedo, 2 normal, 3 material,
// 4-7 previous frame, 8 reservoirs (out), 9 prevFrameReservoirs}; set2 {0 nodes, 1 triangles}
void glslref_restir_pass(const glslref_scene *sc, const void *uniforms, const glslref_gbuffer *cur, const glslref_gbuffer *prev,
                         const void *prevReservoirs, void *out, int y0, int y1) {
	using namespace omni;
	std::memcpy((void *)&omni::uniforms, uniforms, 128);
	int w = (int)omni::uniforms.screenSize.x, h = (int)omni::uniforms.screenSize.y;
	pointLights.count = blobCount(sc->pointBlob);
	pointLights.lights = blobItems<pointLight>(sc->pointBlob);
	triangleLights.count = blobCount(sc->triBlob);
	triangleLights.lights = blobItems<triLight>(sc->triBlob);
	aliasTable.count = blobCount(sc->aliasBlob);
	aliasTable.aliasCol = blobItems<aliasTableColumn>(sc->aliasBlob);
	uniWorldPosition = tex(kRGBA32F, cur->worldPos, w, h);
	uniAlbedo = tex(kRGBA8_SRGB, cur->albedo, w, h);
	uniNormal = tex(kRGBA16_SNORM, cur->normal, w, h);
	uniMaterialProperties = tex(kRG16_UNORM, cur->material, w, h);
	uniPrevFrameWorldPosition = tex(kRGBA32F, prev ? prev->worldPos : nullptr, w, h);
	uniPrevFrameAlbedo = tex(kRGBA8_SRGB, prev ? prev->albedo : nullptr, w, h);
	uniPrevFrameNormal = tex(kRGBA16_SNORM, prev ? prev->normal : nullptr, w, h);
	uniPrevDepth = tex(kD32F, prev ? prev->depth : nullptr, w, h);
	reservoirs = (Reservoir *)out;
	prevFrameReservoirs = (Reservoir *)prevReservoirs;
	aabbTree.nodes = (AabbTreeNode *)sc->nodes;
	triangles = (Triangle *)sc->tris;
	forEachPixel(w, y0, y1, [](int, int) { omni::shader_main(); });
}

// spatialReuse.comp: {0 uniforms, 1 worldPos, 2 albedo, 3 normal, 4 material, 5 depth, 6 reservoirs, 7 result} + push constant iter
void glslref_spatial_pass(const void *uniforms, const glslref_gbuffer *cur, const void *in, void *out, int iter, int y0, int y1) {
	using namespace spatial;
	std::memcpy((void *)&spatial::uniforms, uniforms, 128);
	int w = (int)spatial::uniforms.screenSize.x, h = (int)spatial::uniforms.screenSize.y;
	constant.iter = iter;
	uniWorldPosition = tex(kRGBA32F, cur->worldPos, w, h);
	uniAlbedo = tex(kRGBA8_SRGB, cur->albedo, w, h);
	uniNormal = tex(kRGBA16_SNORM, cur->normal, w, h);
	uniMaterialProperties = tex(kRG16_UNORM, cur->material, w, h);
	uniDepth = tex(kD32F, cur->depth, w, h);
	reservoirs = (Reservoir *)in;
	resultReservoirs = (Reservoir *)out;
	forEachPixel(w, y0, y1, [](int, int) { spatial::shader_main(); });
}

// unbiasedReuse.glsl: set0 {0 worldPos, 1 albedo, 2 normal, 3 material, 4 depth, 5 reservoirs, 6 result, 7 uniforms}; set1 {0 nodes, 1 triangles}
#define GLSLREF_UNBIASED(NS)                                                                                           \
	{                                                                                                                  \
		using namespace NS;                                                                                            \
		std::memcpy((void *)&NS::uniforms, uniforms, 128);                                                                    \
		int w = (int)NS::uniforms.screenSize.x, h = (int)NS::uniforms.screenSize.y;                                    \
		NS::uniWorldPosition = tex(kRGBA32F, cur->worldPos, w, h);                                                     \
		NS::uniAlbedo = tex(kRGBA8_SRGB, cur->albedo, w, h);                                                           \
		NS::uniNormal = tex(kRGBA16_SNORM, cur->normal, w, h);                                                         \
		NS::uniMaterialProperties = tex(kRG16_UNORM, cur->material, w, h);                                             \
		NS::uniDepth = tex(kD32F, cur->depth, w, h);                                                                   \
		NS::reservoirs = (NS::Reservoir *)in;                                                                          \
		NS::resultReservoirs = (NS::Reservoir *)out;                                                                   \
		NS::aabbTree.nodes = (NS::AabbTreeNode *)sc->nodes;                                                            \
		NS::triangles = (NS::Triangle *)sc->tris;                                                                      \
		forEachPixel(w, y0, y1, [](int, int) { NS::shader_main(); });                                                  \
	}

// numNeighbors: 3 = the source as shipped; 5 = the same source with `#define NUM_NEIGHBORS 5`; anything else fails
int glslref_unbiased_pass(const glslref_scene *sc, const void *uniforms, const glslref_gbuffer *cur, const void *in, void *out,
                          int numNeighbors, int y0, int y1) {
	if (numNeighbors == 3) {
		GLSLREF_UNBIASED(unbiased3)
	} else if (numNeighbors == 5) {
		GLSLREF_UNBIASED(unbiased5)
	} else {
		return -1;
	}
	return 0;
}

// lighting.frag: {0 albedo, 1 normal, 2 material, 3 worldPos, 4 uniforms, 5 reservoirs, 6 point, 7 tri}; writes linear RGBA32F
// (the reference's swapchain then quantises to BGRA8 sRGB)
void glslref_lighting_pass(const glslref_scene *sc, const void *lightingUniforms, const glslref_gbuffer *cur, const void *reservoirs,
                           float *outRGBA, int y0, int y1) {
	using namespace lighting;
	std::memcpy((void *)&lighting::uniforms, lightingUniforms, 96);
	int w = (int)lighting::uniforms.bufferSize.x, h = (int)lighting::uniforms.bufferSize.y;
	uniAlbedo = tex(kRGBA8_SRGB, cur->albedo, w, h);
	uniNormal = tex(kRGBA16_SNORM, cur->normal, w, h);
	uniMaterialProperties = tex(kRG16_UNORM, cur->material, w, h);
	uniWorldPosition = tex(kRGBA32F, cur->worldPos, w, h);
	lighting::reservoirs = (Reservoir *)reservoirs;
	pointLights.count = blobCount(sc->pointBlob);
	pointLights.lights = blobItems<pointLight>(sc->pointBlob);
	triangleLights.count = blobCount(sc->triBlob);
	triangleLights.lights = blobItems<triLight>(sc->triBlob);
	forEachPixel(w, y0, y1, [=](int x, int y) {
		gl_FragCoord.xy = vec2((float)x + 0.5f, (float)y + 0.5f);
		inUv = vec2(((float)x + 0.5f) / (float)w, ((float)y + 0.5f) / (float)h); // quad.vert: uv of the pixel centre
		lighting::shader_main();
		float *o = outRGBA + ((size_t)y * (size_t)w + (size_t)x) * 4;
		o[0] = outColor.x;
		o[1] = outColor.y;
		o[2] = outColor.z;
		o[3] = 1.0f;
	});
}

// visibilityTest.glsl:1-4,27-28 + softwareRaytracing.glsl:39-85 as compiled into restirOmniSoftware.comp
void glslref_trace_segments(const glslref_scene *sc, long long n, const float *p1, const float *p2, unsigned char *shadowed) {
	omni::aabbTree.nodes = (omni::AabbTreeNode *)sc->nodes;
	omni::triangles = (omni::Triangle *)sc->tris;
#pragma omp parallel for schedule(dynamic, 256)
	for (long long i = 0; i < n; ++i) {
		shadowed[i] = omni::testVisibility(vec3(p1[3 * i], p1[3 * i + 1], p1[3 * i + 2]), vec3(p2[3 * i], p2[3 * i + 1], p2[3 * i + 2])) ? 1 : 0;
	}
}

// rand.glsl as compiled: n draws of randUint / randFloat from seedRand(seed, seq)
void glslref_pcg32(unsigned long long seed, unsigned long long seq, int n, unsigned *outU, float *outF) {
	omni::Rand r = omni::seedRand(seed, seq);
	for (int i = 0; i < n; ++i) {
		if (outU) {
			omni::Rand c = r;
			outU[i] = omni::randUint(c);
		}
		float f = omni::randFloat(r);
		if (outF) {
			outF[i] = f;
		}
	}
}

// restirUtils.glsl:3-28 as compiled; args[i] = {worldPos, lightPos, camPos, normal, lightNormal, useLightNormal} as 16 floats
void glslref_evaluate_phat(const float *args, int n, float albedoLum, float emissionLum, float roughness, float metallic, float *out) {
	for (int i = 0; i < n; ++i) {
		const float *a = args + 16 * i;
		out[i] = omni::evaluatePHat(vec3(a[0], a[1], a[2]), vec3(a[3], a[4], a[5]), vec3(a[6], a[7], a[8]), vec3(a[9], a[10], a[11]),
		                            vec3(a[12], a[13], a[14]), a[15] > 0.5f, albedoLum, emissionLum, roughness, metallic);
	}
}

void glslref_set_num_threads(int n) {
#ifdef _OPENMP
	if (n > 0) omp_set_num_threads(n);
#endif
}

} // extern "C"
