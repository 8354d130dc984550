/* restir_layouts.h — byte-exact data layouts of the ReSTIR hot path's interface.
 *
 * Every struct here is the C spelling of a struct the reference shares between GLSL and C++ through
 * src/shaderIncludes.h (std430/std140-consistent; nvmath vec4 is 16-byte aligned,
 * thirdparty/nvmath/nvmath_glsltypes.h:36-52).  Sizes and offsets were measured by compiling the
 * reference header and are pinned below with static asserts.
 *
 * Plain C (C11) / C++ / CUDA.  No torch, no Vulkan.
 */
#ifndef RESTIR_LAYOUTS_H_
#define RESTIR_LAYOUTS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
#	define RESTIR_STATIC_ASSERT(c, m) static_assert(c, m)
#	define RESTIR_ALIGN16 alignas(16)
extern "C" {
#else
#	define RESTIR_STATIC_ASSERT(c, m) _Static_assert(c, m)
#	define RESTIR_ALIGN16 _Alignas(16)
#endif

/* reference: src/shaders/include/structs/restirStructs.glsl:19-29 (RESERVOIR_SIZE 1, UNBIASED_MIS off) */
typedef struct restir_light_sample {
	RESTIR_ALIGN16 float position_emissionLum[4]; /*  0: xyz = sample position, w = emission luminance */
	float normal[4];                              /* 16: xyz = light normal, w = 1 for triangle lights */
	int32_t lightIndex;                           /* 32: >=0 point light, -1-i triangle light i */
	float pHat;                                   /* 36 */
	float sumWeights;                             /* 40 */
	float w;                                      /* 44 */
} restir_light_sample;

/* reference: restirStructs.glsl:31-34; 64 bytes with tail padding */
typedef struct restir_reservoir {
	restir_light_sample sample;  /*  0 */
	uint32_t numStreamSamples;   /* 48: "M" */
	uint32_t _pad[3];            /* 52 */
} restir_reservoir;

#define RESTIR_VISIBILITY_REUSE_FLAG (1 << 0) /* restirStructs.glsl:37 */
#define RESTIR_TEMPORAL_REUSE_FLAG (1 << 1)   /* restirStructs.glsl:38 */

/* reference: restirStructs.glsl:40-56; the 128-byte uniform block, verbatim */
typedef struct restir_uniforms {
	RESTIR_ALIGN16 float prevFrameProjectionViewMatrix[16]; /*   0: column-major mat4 */
	float cameraPos[4];                                     /*  64 */
	uint32_t screenSize[2];                                 /*  80 */
	uint32_t frame;                                         /*  88 */
	uint32_t initialLightSampleCount;                       /*  92 */
	uint32_t temporalSampleCountMultiplier;                 /*  96 */
	float spatialPosThreshold;                              /* 100 */
	float spatialNormalThreshold;                           /* 104: degrees */
	uint32_t spatialNeighbors;                              /* 108 */
	float spatialRadius;                                    /* 112 */
	int32_t flags;                                          /* 116 */
	uint32_t _pad[2];                                       /* 120 */
} restir_uniforms;

/* reference: src/shaders/include/structs/lightingPassStructs.glsl:1-7 */
typedef struct restir_lighting_uniforms {
	RESTIR_ALIGN16 float prevFrameProjectionViewMatrix[16]; /*  0 */
	float cameraPos[4];                                     /* 64 */
	uint32_t bufferSize[2];                                 /* 80 */
	int32_t debugMode;                                      /* 88: only 0 (GBUFFER_DEBUG_NONE) is on the hot path */
	float gamma;                                            /* 92 */
} restir_lighting_uniforms;

/* reference: src/shaders/include/structs/aabbTree.glsl:1-8.  2-wide node: the boxes of BOTH children
 * live in the parent; child >= 0 is a node index, child < 0 is ~triangleIndex; node 0 is the root. */
typedef struct restir_aabb_node {
	RESTIR_ALIGN16 float leftAabbMin[4];
	float leftAabbMax[4];
	float rightAabbMin[4];
	float rightAabbMax[4];
	int32_t leftChild;  /* 64 */
	int32_t rightChild; /* 68 */
	int32_t _pad[2];    /* 72 */
} restir_aabb_node;

/* reference: aabbTree.glsl:9-13; w = 1 (from the matrix multiply in aabbTreeBuilder.cpp:66-68) */
typedef struct restir_triangle {
	RESTIR_ALIGN16 float p1[4];
	float p2[4];
	float p3[4];
} restir_triangle;

/* reference: src/shaders/include/structs/light.glsl:2-5 */
typedef struct restir_point_light {
	RESTIR_ALIGN16 float pos[4];
	float color_luminance[4]; /* w = luminance */
} restir_point_light;

/* reference: light.glsl:7-13 */
typedef struct restir_tri_light {
	RESTIR_ALIGN16 float p1[4];
	float p2[4];
	float p3[4];
	float emission_luminance[4]; /* w = luminance */
	float normalArea[4];         /* xyz = unit normal, w = area */
} restir_tri_light;

/* reference: light.glsl:15-20 */
typedef struct restir_alias_column {
	float prob;
	int32_t alias;
	float oriProb;
	float aliasOriProb;
} restir_alias_column;

/* The three light SSBOs are "blobs": int32 count at byte 0, the array at byte 16
 * (reference: src/sceneBuffers.h:100-124 sizes, :241-270 contents, src/misc.h:28-30). */
#define RESTIR_BLOB_HEADER_BYTES 16

/* G-buffer planes as the hot path samples them (texelFetch / nearest at pixel centres), in the
 * formats the reference's GBuffer::Formats picks on NVIDIA hardware
 * (src/passes/gBufferPass.cpp:75-108): 4 + 8 + 4 + 16 + 4 = 36 bytes per pixel.
 * Clears (gBufferPass.cpp:117-123): colour (0,0,0,1), depth 1.0 => background has normal 0, albedo.a 1. */
typedef enum restir_gbuffer_format {
	RESTIR_GBUFFER_NVIDIA_DEFAULT = 0
	/* albedo   R8G8B8A8_SRGB   (rgb sRGB-encoded, a linear = emissive flag)   4 B/px
	 * normal   R16G16B16A16_SNORM                                            8 B/px
	 * material R16G16_UNORM    (roughness, metallic)                         4 B/px
	 * worldPos R32G32B32A32_SFLOAT                                          16 B/px
	 * depth    D32_SFLOAT      (raw depth-buffer value)                      4 B/px */
} restir_gbuffer_format;

typedef struct restir_gbuffer_planes {
	const void *albedo;
	const void *normal;
	const void *material;
	const void *worldPos;
	const void *depth;
} restir_gbuffer_planes;

/* ---- what the G-buffer pass binds (src/passes/gBufferPass.cpp:116-157), for restir_pass_gbuffer --------------------
 * reference: src/vertex.h (vec4 position, normal, tangent, color; vec2 uv; 16-byte aligned => 80 bytes).  The vertex
 * shader reads position.xyz, normal.xyz, tangent.xyzw and uv (gBuffer.vert:12-16); color is carried and unused. */
typedef struct restir_vertex {
	float position[4];
	float normal[4];
	float tangent[4];
	float color[4];
	float uv[2];
	float pad_[2];
} restir_vertex;

/* One drawIndexed of GBufferPass::issueCommands (gBufferPass.cpp:135-152): the GltfPrimMesh fields it uses.  Draw d uses
 * matrices[d] (the dynamic offset is i * sizeof(ModelMatrices)). */
typedef struct restir_draw {
	uint32_t firstIndex;
	uint32_t indexCount;
	uint32_t vertexOffset;
	int32_t materialIndex;
} restir_draw;

/* reference: sceneStructs.glsl:1-4, filled at sceneBuffers.h:225-230 (column-major) */
typedef struct restir_model_matrices {
	float transform[16];
	float transformInverseTransposed[16];
} restir_model_matrices;

#define RESTIR_SHADING_MODEL_METALLIC_ROUGHNESS 0
#define RESTIR_SHADING_MODEL_SPECULAR_GLOSSINESS 1
#define RESTIR_ALPHA_MODE_OPAQUE 0
#define RESTIR_ALPHA_MODE_MASK 1
#define RESTIR_ALPHA_MODE_BLEND 2
/* reference: sceneStructs.glsl:15-34, filled at sceneBuffers.h:205-222 */
typedef struct restir_material_uniforms {
	float colorParam[4];    /* baseColorFactor | diffuseFactor */
	float materialParam[4]; /* (-, roughnessFactor, metallicFactor, -) | (specularFactor.rgb, glossinessFactor) */
	float emissiveFactor[4];
	int32_t shadingModel;
	int32_t alphaMode;
	float alphaCutoff;
	float normalTextureScale;
} restir_material_uniforms;

/* The four combined image samplers of a material's descriptor set (gBuffer.frag:11-14, written at
 * gBufferPass.cpp:200-247): indices into the texture array, -1 = the default texture of that binding
 * (white 255,255,255,255; for `normal` 127,127,255,255 — sceneBuffers.h:153-170). */
typedef struct restir_material_textures {
	int32_t albedo;   /* baseColor | diffuse */
	int32_t normal;
	int32_t material; /* metallicRoughness | specularGlossiness */
	int32_t emissive;
} restir_material_textures;

/* One texture image: R8G8B8A8_UNORM as SceneBuffers uploads it (sceneBuffers.h:126), level 0, tightly packed rows. */
typedef struct restir_texture {
	const void *rgba8;
	uint32_t width, height;
} restir_texture;

RESTIR_STATIC_ASSERT(sizeof(restir_vertex) == 80, "Vertex is 80 bytes");
RESTIR_STATIC_ASSERT(offsetof(restir_vertex, uv) == 64, "Vertex.uv");
RESTIR_STATIC_ASSERT(sizeof(restir_draw) == 16, "draw record is 16 bytes");
RESTIR_STATIC_ASSERT(sizeof(restir_model_matrices) == 128, "ModelMatrices is 128 bytes");
RESTIR_STATIC_ASSERT(sizeof(restir_material_uniforms) == 64, "MaterialUniforms is 64 bytes");
RESTIR_STATIC_ASSERT(offsetof(restir_material_uniforms, emissiveFactor) == 32, "MaterialUniforms.emissiveFactor");
RESTIR_STATIC_ASSERT(offsetof(restir_material_uniforms, shadingModel) == 48, "MaterialUniforms.shadingModel");
RESTIR_STATIC_ASSERT(offsetof(restir_material_uniforms, normalTextureScale) == 60, "MaterialUniforms.normalTextureScale");
RESTIR_STATIC_ASSERT(sizeof(restir_light_sample) == 48, "LightSample is 48 bytes");
RESTIR_STATIC_ASSERT(offsetof(restir_light_sample, normal) == 16, "LightSample.normal");
RESTIR_STATIC_ASSERT(offsetof(restir_light_sample, lightIndex) == 32, "LightSample.lightIndex");
RESTIR_STATIC_ASSERT(offsetof(restir_light_sample, pHat) == 36, "LightSample.pHat");
RESTIR_STATIC_ASSERT(offsetof(restir_light_sample, sumWeights) == 40, "LightSample.sumWeights");
RESTIR_STATIC_ASSERT(offsetof(restir_light_sample, w) == 44, "LightSample.w");
RESTIR_STATIC_ASSERT(sizeof(restir_reservoir) == 64, "Reservoir is 64 bytes");
RESTIR_STATIC_ASSERT(offsetof(restir_reservoir, numStreamSamples) == 48, "Reservoir.numStreamSamples");
RESTIR_STATIC_ASSERT(sizeof(restir_uniforms) == 128, "RestirUniforms is 128 bytes");
RESTIR_STATIC_ASSERT(offsetof(restir_uniforms, cameraPos) == 64, "RestirUniforms.cameraPos");
RESTIR_STATIC_ASSERT(offsetof(restir_uniforms, screenSize) == 80, "RestirUniforms.screenSize");
RESTIR_STATIC_ASSERT(offsetof(restir_uniforms, frame) == 88, "RestirUniforms.frame");
RESTIR_STATIC_ASSERT(offsetof(restir_uniforms, initialLightSampleCount) == 92, "RestirUniforms.initialLightSampleCount");
RESTIR_STATIC_ASSERT(offsetof(restir_uniforms, temporalSampleCountMultiplier) == 96, "RestirUniforms.temporalSampleCountMultiplier");
RESTIR_STATIC_ASSERT(offsetof(restir_uniforms, spatialPosThreshold) == 100, "RestirUniforms.spatialPosThreshold");
RESTIR_STATIC_ASSERT(offsetof(restir_uniforms, spatialNormalThreshold) == 104, "RestirUniforms.spatialNormalThreshold");
RESTIR_STATIC_ASSERT(offsetof(restir_uniforms, spatialNeighbors) == 108, "RestirUniforms.spatialNeighbors");
RESTIR_STATIC_ASSERT(offsetof(restir_uniforms, spatialRadius) == 112, "RestirUniforms.spatialRadius");
RESTIR_STATIC_ASSERT(offsetof(restir_uniforms, flags) == 116, "RestirUniforms.flags");
RESTIR_STATIC_ASSERT(sizeof(restir_lighting_uniforms) == 96, "LightingPassUniforms is 96 bytes");
RESTIR_STATIC_ASSERT(offsetof(restir_lighting_uniforms, bufferSize) == 80, "LightingPassUniforms.bufferSize");
RESTIR_STATIC_ASSERT(offsetof(restir_lighting_uniforms, debugMode) == 88, "LightingPassUniforms.debugMode");
RESTIR_STATIC_ASSERT(offsetof(restir_lighting_uniforms, gamma) == 92, "LightingPassUniforms.gamma");
RESTIR_STATIC_ASSERT(sizeof(restir_aabb_node) == 80, "AabbTreeNode is 80 bytes");
RESTIR_STATIC_ASSERT(offsetof(restir_aabb_node, leftChild) == 64, "AabbTreeNode.leftChild");
RESTIR_STATIC_ASSERT(offsetof(restir_aabb_node, rightChild) == 68, "AabbTreeNode.rightChild");
RESTIR_STATIC_ASSERT(sizeof(restir_triangle) == 48, "Triangle is 48 bytes");
RESTIR_STATIC_ASSERT(sizeof(restir_point_light) == 32, "pointLight is 32 bytes");
RESTIR_STATIC_ASSERT(sizeof(restir_tri_light) == 80, "triLight is 80 bytes");
RESTIR_STATIC_ASSERT(sizeof(restir_alias_column) == 16, "aliasTableColumn is 16 bytes");

#ifdef __cplusplus
} /* extern "C" */
#endif

#endif /* RESTIR_LAYOUTS_H_ */
