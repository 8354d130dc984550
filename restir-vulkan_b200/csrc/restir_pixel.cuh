// restir_pixel.cuh — what every per-pixel kernel shares: the thread -> pixel mapping, G-buffer texel decode, alias-table
// sampling, counters.  Included by restir_kernels.cu (the tuned path) and restir_generic.cu (the reference's compile-time variants).
#pragma once

#include "restir_device.cuh"

namespace restir {

// ------------------------------------------------------------------------------------------------
// thread -> pixel mapping
constexpr int kTileW = 32, kTileH = 8, kThreads = 256;

__device__ __forceinline__ bool pixel_of_thread(const Band &b, int &x, int &y) {
	int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	x = blockIdx.x * kTileW + (warp & 3) * 8 + (lane & 7);
	y = b.rowBegin + blockIdx.y * kTileH + (warp >> 2) * 4 + (lane >> 3);
	return x < b.W && y < b.rowEnd;
}
// Tile-ordered pixel id of this thread: (8x4 tile index, lane).  The trace kernel numbers its work items the
// same way (restir_trace.cu item_pixel).
__device__ __forceinline__ unsigned long long tile_pixel_id() {
	unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	unsigned tileX = blockIdx.x * 4u + (warp & 3u), tileY = blockIdx.y * 2u + (warp >> 2);
	return ((unsigned long long)tileY * (gridDim.x * 4u) + tileX) * 32ull + lane;
}
__device__ __forceinline__ size_t local_index(const Band &b, int x, int y) {
	return (size_t)(y - b.allocBegin) * (size_t)b.W + (size_t)x;
}

// ------------------------------------------------------------------------------------------------
// G-buffer texel decode (texelFetch of the NVIDIA-default formats)
__device__ __forceinline__ f3 fetch_albedo(const GBufferView &g, const float *lut, size_t i, float *alpha) {
	if (g.albedo == nullptr) {
		if (alpha) *alpha = 0.0f;
		return mk3(0.0f, 0.0f, 0.0f);
	}
	uchar4 c = __ldg(g.albedo + i);
	if (alpha) *alpha = div_unorm8((float)c.w);
	return mk3(__ldg(lut + c.x), __ldg(lut + c.y), __ldg(lut + c.z));
}
__device__ __forceinline__ f3 fetch_normal(const GBufferView &g, size_t i) {
	if (g.normal == nullptr) {
		return mk3(0.0f, 0.0f, 0.0f);
	}
	short4 n = __ldg(g.normal + i);
	return mk3(fmaxf(div_snorm16((float)n.x), -1.0f), fmaxf(div_snorm16((float)n.y), -1.0f), fmaxf(div_snorm16((float)n.z), -1.0f));
}
__device__ __forceinline__ void fetch_material(const GBufferView &g, size_t i, float &roughness, float &metallic) {
	ushort2 m = __ldg(g.material + i);
	roughness = div_unorm16((float)m.x);
	metallic = div_unorm16((float)m.y);
}
__device__ __forceinline__ f3 fetch_world_pos(const GBufferView &g, size_t i) {
	if (g.worldPos == nullptr) {
		return mk3(0.0f, 0.0f, 0.0f);
	}
	float4 p = __ldg(g.worldPos + i);
	return mk3(p.x, p.y, p.z);
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void add_counter(unsigned long long *counters, int slot, unsigned v) {
	// one atomic per warp
	unsigned total = __reduce_add_sync(0xffffffffu, v);
	if ((threadIdx.x & 31) == 0 && total != 0) {
		atomicAdd(counters + slot, (unsigned long long)total);
	}
}

// restirOmni.glsl:73-83
__device__ __forceinline__ void alias_sample(const SceneView &sc, float r1, float r2, int &index, float &prob) {
	int col = min((int)((float)sc.aliasCount * r1), sc.aliasCount - 1);
	float4 c = __ldg(reinterpret_cast<const float4 *>(sc.alias) + col);
	if (c.x > r2) {
		index = col;
		prob = c.z;
	} else {
		index = __float_as_int(c.y);
		prob = c.w;
	}
}

// lighting.frag's output through an 8-bit sRGB swapchain image
__device__ __forceinline__ float srgb_encode(float c) {
	c = fminf(fmaxf(c, 0.0f), 1.0f);
	return c <= 0.0031308f ? 12.92f * c : 1.055f * powf(c, 1.0f / 2.4f) - 0.055f;
}

} // namespace restir
