// restir_device.cuh — device-side views shared by the kernels and the C-ABI context.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/restir_layouts.h"
#include "restir_math.cuh"

namespace restir {

// Internal HBM layout of a reservoir: 32 bytes, lossless w.r.t. the reference's 64-byte struct
// (restirStructs.glsl:19-34) for every reservoir the path produces: `normal` and `emissionLum` are
// functions of lightIndex (restirOmni.glsl:117-133) and are re-read from the light tables when a
// reservoir is resampled; pHat == 0 marks "no sample ever selected" (all sample fields zero).
struct __align__(32) PackedReservoir {
	float px, py, pz;
	int lightIndex;
	float pHat, sumWeights, w;
	uint32_t M;
};
static_assert(sizeof(PackedReservoir) == 32, "packed reservoir is 32 bytes");

struct SceneView {
	const float4 *nodes;       // reference layout, 5 x float4 per node (80 B)
	const float4 *tris;        // reference layout, 3 x float4 per triangle (48 B)
	const float4 *image;       // derived at upload (traversal_image.h): the same nodes re-strided to 4 x float4 (64 B); null => literal 80-byte walk
	const float4 *triEdges;    // derived at upload: 64-byte (p1, e1, e2) records of `tris` for the walk over `image` (restir_trace.cuh)
	const restir_point_light *pointLights;
	const restir_tri_light *triLights;
	const restir_alias_column *alias;
	const float4 *pointPosLum; // derived at upload: (pos.xyz, luminance), 16 B per point light
	const float4 *triAux;      // derived at upload: (normal.xyz, luminance), 16 B per triangle light
	const float *srgbLut;      // P12: 256-entry sRGB8 -> linear table
	int pointCount, triCount, aliasCount;
	uint32_t nNodes, nTris;
};

struct GBufferView {
	const uchar4 *albedo;    // R8G8B8A8_SRGB
	const short4 *normal;    // R16G16B16A16_SNORM
	const ushort2 *material; // R16G16_UNORM (roughness, metallic)
	const float4 *worldPos;  // R32G32B32A32_SFLOAT
	const float *depth;      // D32_SFLOAT
};

// Screen and band geometry.  Per-pixel buffers cover rows [allocBegin, allocEnd); kernels shade rows
// [rowBegin, rowEnd).  Single GPU: rowBegin = allocBegin = 0, rowEnd = allocEnd = H.
struct Band {
	int W, H;
	int rowBegin, rowEnd;
	int allocBegin, allocEnd;
};

enum CounterSlot { kCounterRays = 0, kCounterOverflow = 1, kCounterHaloMiss = 2, kCounterWork = 3, kCounterTraced = 4, kCounterHaloTimeout = 5, kCounterCached = 6, kCounterCount = 7 };
// kCounterRays counts the reference's testVisibility calls that were answered, kCounterTraced the ones that needed a walk of the tree.
// kCounterCached (part of kCounterRays): answered by the occluder cache — one triangle test instead of a walk (restir_trace.cu).
// kCounterWork is the persistent trace kernel's work cursor (zeroed before every trace launch).

struct PassParams {
	SceneView scene;
	GBufferView cur, prev; // prev.* may be null (zero texels)
	Band band;
	restir_uniforms u;
	unsigned long long *counters;
};

} // namespace restir
