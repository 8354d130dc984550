// restir_trace.cu — the persistent shadow-ray kernel (sm_100a).
//
// One kernel traces every shadow ray of the path: the visibility-reuse ray of restirOmni.glsl:148-160, the
// neighbour and self rays of unbiasedReuse.glsl:126-166, and the stand-alone segments of
// restir_trace_segments.  The reference traces them inline, one thread per pixel, in whatever order the
// pixels come.  Here rays are work items of a persistent grid:
//
//   * every WARP pulls chunks of kChunk consecutive items (items are numbered by 8x4 screen tile, so a chunk
//     is a compact screen patch) from a global cursor;
//   * it sorts the chunk by the light the ray is aimed at (bitonic sort of 32-bit keys in shared memory), so
//     the 32 rays a warp then walks in lockstep start next to each other AND end at the same light: they
//     visit the same nodes at the same time (one L1 wavefront serves many lanes — the measured limiter of
//     this kernel is L1 wavefronts, not DRAM, profiles/) and finish at about the same time;
//   * each lane builds its segment from the G-buffer / reservoir (visibilityTest.glsl:1-4) and walks the tree
//     (restir_trace.cuh).
//
// Measured alternatives that lost (B200, Sponza 1080p, profiles/r1_b*, r1_d*): per-lane refill from a
// shared-memory ray queue (lanes decorrelate, every 16-byte node load becomes its own L1 wavefront: 1.8
// Grays/s), and a 4-wide re-layout of the tree in lockstep or with refill (same instruction count per ray as
// the 2-wide walk once the exact slab arithmetic is kept: 2.9-3.3 Grays/s) against 4.2 Grays/s for the plain
// 2-wide walk of this file's first version.
//
// Per item one byte is written: 1 = shadowed.

#include "restir_kernels.h"
#include "restir_trace.cuh"

namespace restir {

constexpr int kTraceThreads = 256;
constexpr int kTraceWarps = kTraceThreads / 32;
#ifndef RESTIR_TRACE_CHUNK
#define RESTIR_TRACE_CHUNK 128
#endif
constexpr int kChunk = RESTIR_TRACE_CHUNK; // items per warp fetch (power of two, <= 256: the local id is 8 bits of the sort key)
#ifndef RESTIR_TRACE_SORT
#define RESTIR_TRACE_SORT 1
#endif
constexpr unsigned kInvalidKey = 0xffffffffu;

// ---- item -> pixel, key, segment ---------------------------------------------------------------------------

__device__ __forceinline__ bool item_pixel(const TraceParams &tp, unsigned p, size_t &pix) {
	// p is a tile-ordered pixel id: (tile, lane) with 8x4 tiles, tilesX tiles per tile row (restir_kernels.cu
	// tile_pixel_id produces the same numbering)
	unsigned tile = p >> 5, l = p & 31u;
	unsigned ty = tile / tp.tilesX, tx = tile - ty * tp.tilesX;
	int x = (int)(tx * 8u + (l & 7u));
	int y = tp.band.rowBegin + (int)(ty * 4u + (l >> 3));
	if (x >= tp.band.W || y >= tp.band.rowEnd) {
		return false;
	}
	pix = (size_t)(y - tp.band.allocBegin) * (size_t)tp.band.W + (size_t)x;
	return true;
}

// Resolves an item to (pixel holding the sample, pixel the ray starts from); false = no ray for this item.
template <int MODE> __device__ __forceinline__ bool item_pixels(const TraceParams &tp, unsigned item, size_t &pix, size_t &opix) {
	unsigned p = item, slot = 0;
	if (MODE == kTraceUnbiased) {
		p = item / tp.slots;
		slot = item - p * tp.slots;
	}
	if (!item_pixel(tp, p, pix)) {
		return false;
	}
	opix = pix;
	if (MODE == kTraceUnbiased && slot + 1 < tp.slots) { // neighbour ray: starts at the neighbour's surface point
		int n = tp.neighborPix[(size_t)p * (tp.slots - 1) + slot];
		if (n < 0) {
			return false;
		}
		opix = (size_t)n;
	}
	return true;
}

// Sort key of an item: (light the ray is aimed at, position in the chunk); kInvalidKey = no ray.
template <int MODE> __device__ __forceinline__ unsigned item_key(const TraceParams &tp, unsigned item, unsigned local) {
	if (item >= tp.nItems) {
		return kInvalidKey;
	}
	if (MODE == kTraceSegments) {
		return local;
	}
	size_t pix, opix;
	if (!item_pixels<MODE>(tp, item, pix, opix)) {
		return kInvalidKey;
	}
	unsigned light = RESTIR_TRACE_SORT ? (unsigned)__ldg(reinterpret_cast<const int *>(tp.reservoirs + pix) + 3) : 0u; // PackedReservoir::lightIndex
	return ((light & 0x7fffffu) << 8) | local;
}

template <int MODE> __device__ __forceinline__ void item_segment(const TraceParams &tp, unsigned item, f3 &p1, f3 &p2) {
	if (MODE == kTraceSegments) {
		const float *a = tp.segP1 + (size_t)item * 3, *b = tp.segP2 + (size_t)item * 3;
		p1 = mk3(a[0], a[1], a[2]);
		p2 = mk3(b[0], b[1], b[2]);
		return;
	}
	size_t pix, opix;
	item_pixels<MODE>(tp, item, pix, opix);
	float4 w = __ldg(tp.worldPos + opix);
	float4 t = __ldg(reinterpret_cast<const float4 *>(tp.reservoirs + pix));
	p1 = mk3(w.x, w.y, w.z);
	p2 = mk3(t.x, t.y, t.z);
}

// bitonic sort of kChunk keys in shared memory by one warp
__device__ __forceinline__ void warp_sort(unsigned *keys, unsigned lane) {
#pragma unroll 1
	for (unsigned k = 2; k <= (unsigned)kChunk; k <<= 1) {
#pragma unroll 1
		for (unsigned j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
			for (unsigned t = lane; t < (unsigned)kChunk / 2; t += 32) {
				unsigned i = ((t & ~(j - 1u)) << 1) | (t & (j - 1u));
				unsigned a = keys[i], b = keys[i + j];
				bool ascending = (i & k) == 0u;
				if ((a > b) == ascending) {
					keys[i] = b;
					keys[i + j] = a;
				}
			}
			__syncwarp();
		}
	}
}

// ---- the kernel ------------------------------------------------------------------------------------------

template <int MODE, bool IMAGE> __global__ void __launch_bounds__(kTraceThreads) trace_kernel(const __grid_constant__ TraceParams tp) {
	__shared__ unsigned allKeys[kTraceWarps][kChunk];
	const unsigned lane = threadIdx.x & 31u;
	unsigned *keys = allKeys[threadIdx.x >> 5];
	const unsigned full = 0xffffffffu;
	unsigned rays = 0, overflow = 0;

	for (;;) {
		unsigned base = 0;
		if (lane == 0) {
			base = (unsigned)min(atomicAdd(tp.counters + kCounterWork, (unsigned long long)kChunk), 0xffffffffull);
		}
		base = __shfl_sync(full, base, 0);
		if (base >= tp.nItems) {
			break;
		}
#pragma unroll 1
		for (unsigned r = 0; r < (unsigned)kChunk / 32; ++r) {
			unsigned local = r * 32u + lane;
			keys[local] = item_key<MODE>(tp, base + local, local);
		}
		__syncwarp();
		if (MODE != kTraceSegments && RESTIR_TRACE_SORT) {
			warp_sort(keys, lane);
		}
#pragma unroll 1
		for (unsigned r = 0; r < (unsigned)kChunk / 32; ++r) {
			unsigned key = keys[r * 32u + lane];
			if (__ballot_sync(full, key != kInvalidKey) == 0u) {
				if (MODE != kTraceSegments && RESTIR_TRACE_SORT) break; // sorted: only holes follow
				continue;
			}
			if (key != kInvalidKey) {
				unsigned item = base + (key & 255u);
				f3 p1, p2, o, d;
				item_segment<MODE>(tp, item, p1, p2);
				segment_setup(p1, p2, o, d);
				bool clear = IMAGE ? trace_any_image(tp.image, tp.tris, o, d) : trace_any_reference(tp.nodes, tp.tris, o, d, overflow);
				tp.shadowed[item] = clear ? 0 : 1;
				rays++;
			}
		}
		__syncwarp();
	}
	// one atomic per warp
	rays = __reduce_add_sync(full, rays);
	overflow = __reduce_add_sync(full, overflow);
	if (lane == 0) {
		if (rays) atomicAdd(tp.counters + kCounterRays, (unsigned long long)rays);
		if (overflow) atomicAdd(tp.counters + kCounterOverflow, (unsigned long long)overflow);
	}
}

// ---- launcher ----------------------------------------------------------------------------------------------

template <int MODE, bool IMAGE> static cudaError_t launch_mode(const TraceParams &tp, int smCount, cudaStream_t s) {
	static int blocksPerSm = 0; // same for every device of one box
	if (blocksPerSm == 0) {
		cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSm, trace_kernel<MODE, IMAGE>, kTraceThreads, 0);
		if (e != cudaSuccess) {
			return e;
		}
		if (blocksPerSm < 1) blocksPerSm = 1;
	}
	// items are numbered with 32 bits inside the kernel; restir_capi.cu splits longer segment lists
	unsigned long long chunks = (tp.nItems + kChunk - 1) / kChunk;
	unsigned long long wanted = (chunks + kTraceWarps - 1) / kTraceWarps;
	unsigned grid = (unsigned)std::min<unsigned long long>((unsigned long long)smCount * blocksPerSm, std::max<unsigned long long>(wanted, 1));
	trace_kernel<MODE, IMAGE><<<grid, kTraceThreads, 0, s>>>(tp);
	return cudaGetLastError();
}

cudaError_t launch_trace(const TraceParams &tp, int mode, int smCount, cudaStream_t s) {
	if (tp.nItems == 0) {
		return cudaSuccess;
	}
	cudaError_t e = cudaMemsetAsync(tp.counters + kCounterWork, 0, sizeof(unsigned long long), s);
	if (e != cudaSuccess) {
		return e;
	}
	const bool image = tp.image != nullptr;
	switch (mode) {
	case kTracePixel: return image ? launch_mode<kTracePixel, true>(tp, smCount, s) : launch_mode<kTracePixel, false>(tp, smCount, s);
	case kTraceUnbiased: return image ? launch_mode<kTraceUnbiased, true>(tp, smCount, s) : launch_mode<kTraceUnbiased, false>(tp, smCount, s);
	default: return image ? launch_mode<kTraceSegments, true>(tp, smCount, s) : launch_mode<kTraceSegments, false>(tp, smCount, s);
	}
}

} // namespace restir
