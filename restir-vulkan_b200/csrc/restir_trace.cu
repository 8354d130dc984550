// restir_trace.cu — the persistent shadow-ray kernel (sm_100a).
//
// One kernel traces every shadow ray of the path: the visibility-reuse ray of restirOmni.glsl:148-160, the
// neighbour and self rays of unbiasedReuse.glsl:126-166, and the stand-alone segments of
// restir_trace_segments.  The reference traces them inline, one thread per pixel, so a warp runs as long as
// its slowest ray (measured on B200 for the first, fused version of this path: 10-18 of 32 lanes active in
// traversal, profiles/r1_a_summary.md).  Here rays are work items:
//
//   * the grid is persistent (CTAs resident on all SMs); every WARP pulls chunks of kChunk consecutive items
//     from a global cursor, builds their segments with all lanes active (G-buffer / reservoir gathers,
//     normalize, 1/dir) and parks them in its own shared-memory queue;
//   * each LANE then traces one ray at a time and, when its ray ends, takes the next one from the warp's
//     queue (ballot + prefix popcount, no atomics): lanes refill individually, so the warp stays full until
//     the global cursor runs dry;
//   * items are ordered by 8x4 screen tile, so a chunk is a compact screen patch aimed at a few lights.
//
// Per item one byte is written: 1 = shadowed.

#include "restir_kernels.h"
#include "restir_trace.cuh"

namespace restir {

constexpr int kTraceThreads = 256;
constexpr int kTraceWarps = kTraceThreads / 32;
constexpr int kChunk = 64; // items per warp fetch; kChunk / 32 items per lane per staging round
// lanes take new rays once at least this many of the warp's lanes are idle (1 = at once, 32 = lockstep)
#ifndef RESTIR_REFILL_THRESHOLD
#define RESTIR_REFILL_THRESHOLD 1
#endif

struct __align__(16) StagedRay {
	float4 a; // origin.xyz, dir.x
	float4 b; // dir.yz, inv.xy
	float4 c; // inv.z, item (low, high 32 bits as float bits), unused
};

// ---- item -> segment end points ------------------------------------------------------------------------

__device__ __forceinline__ bool item_pixel(const TraceParams &tp, unsigned long long p, size_t &pix) {
	// p is a tile-ordered pixel id: (tile, lane) with 8x4 tiles, tilesX tiles per tile row (restir_kernels.cu
	// pixel_of_thread produces the same numbering)
	unsigned tile = (unsigned)(p >> 5), l = (unsigned)p & 31u;
	unsigned ty = tile / tp.tilesX, tx = tile - ty * tp.tilesX;
	int x = (int)(tx * 8u + (l & 7u));
	int y = tp.band.rowBegin + (int)(ty * 4u + (l >> 3));
	if (x >= tp.band.W || y >= tp.band.rowEnd) {
		return false;
	}
	pix = (size_t)(y - tp.band.allocBegin) * (size_t)tp.band.W + (size_t)x;
	return true;
}

template <int MODE> __device__ __forceinline__ bool item_segment(const TraceParams &tp, unsigned long long item, f3 &p1, f3 &p2) {
	if (MODE == kTraceSegments) {
		const float *a = tp.segP1 + item * 3, *b = tp.segP2 + item * 3;
		p1 = mk3(a[0], a[1], a[2]);
		p2 = mk3(b[0], b[1], b[2]);
		return true;
	}
	unsigned long long p = item;
	unsigned slot = 0;
	if (MODE == kTraceUnbiased) {
		p = item / tp.slots;
		slot = (unsigned)(item - p * tp.slots);
	}
	size_t pix;
	if (!item_pixel(tp, p, pix)) {
		return false;
	}
	size_t opix = pix;
	if (MODE == kTraceUnbiased && slot + 1 < tp.slots) { // neighbour ray: starts at the neighbour's surface point
		int n = tp.neighborPix[p * (tp.slots - 1) + slot];
		if (n < 0) {
			return false;
		}
		opix = (size_t)n;
	}
	float4 w = __ldg(tp.worldPos + opix);
	float4 t = __ldg(reinterpret_cast<const float4 *>(tp.reservoirs + pix));
	p1 = mk3(w.x, w.y, w.z);
	p2 = mk3(t.x, t.y, t.z);
	return true;
}

// ---- the kernel ------------------------------------------------------------------------------------------

template <int MODE, bool WIDE> __global__ void __launch_bounds__(kTraceThreads, 3) trace_kernel(TraceParams tp) {
	__shared__ StagedRay queues[kTraceWarps][kChunk];
	const unsigned lane = threadIdx.x & 31u;
	StagedRay *q = queues[threadIdx.x >> 5];
	const unsigned full = 0xffffffffu;

	int qHead = 0, qCount = 0;
	bool exhausted = false;
	unsigned rays = 0, overflow = 0;

	bool alive = false;
	WideRay ray;
	unsigned long long item = 0;
	int cur = 0, top = 0;
	int stack[kWideStack];

	for (;;) {
		unsigned need = __ballot_sync(full, !alive);
		if (__popc(need) >= RESTIR_REFILL_THRESHOLD) {
			while (qHead == qCount && !exhausted) {
				unsigned long long base = 0;
				if (lane == 0) {
					base = atomicAdd(tp.counters + kCounterWork, (unsigned long long)kChunk);
				}
				base = __shfl_sync(full, base, 0);
				if (base >= tp.nItems) {
					exhausted = true;
					break;
				}
				// stage this chunk: every lane builds kChunk / 32 segments, valid ones are packed into the queue
				qHead = 0;
				qCount = 0;
				__syncwarp();
#pragma unroll
				for (int r = 0; r < kChunk / 32; ++r) {
					unsigned long long it = base + (unsigned)(r * 32) + lane;
					f3 p1, p2, o, d, inv;
					bool valid = it < tp.nItems && item_segment<MODE>(tp, it, p1, p2);
					bool queued = false;
					if (valid) {
						rays++;
						segment_setup(p1, p2, o, d);
						inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
						// 0 * inf = NaN breaks the slab test's monotonicity: such rays (axis-parallel, degenerate, NaN)
						// are traced right here in the reference's own order.  Rare.
						bool finite = fabsf(inv.x) < __int_as_float(0x7f800000) && fabsf(inv.y) < __int_as_float(0x7f800000) &&
						              fabsf(inv.z) < __int_as_float(0x7f800000);
						if (WIDE && finite) {
							queued = true;
						} else {
							tp.shadowed[it] = trace_any_reference(tp.nodes, tp.tris, o, d, overflow) ? 0 : 1;
						}
					}
					unsigned qm = __ballot_sync(full, queued);
					if (queued) {
						StagedRay s;
						s.a = make_float4(o.x, o.y, o.z, d.x);
						s.b = make_float4(d.y, d.z, inv.x, inv.y);
						s.c = make_float4(inv.z, __uint_as_float((unsigned)it), __uint_as_float((unsigned)(it >> 32)), 0.0f);
						q[qCount + __popc(qm & ((1u << lane) - 1u))] = s;
					}
					qCount += __popc(qm);
				}
				__syncwarp();
			}
			int avail = qCount - qHead;
			int rank = __popc(need & ((1u << lane) - 1u));
			if (!alive && rank < avail) {
				const StagedRay &s = q[qHead + rank];
				float4 a = s.a, b = s.b, c = s.c;
				wide_ray_init(ray, mk3(a.x, a.y, a.z), mk3(a.w, b.x, b.y), mk3(b.z, b.w, c.x));
				item = (unsigned long long)__float_as_uint(c.y) | ((unsigned long long)__float_as_uint(c.z) << 32);
				cur = 0;
				top = 0;
				alive = true;
			}
			qHead += min(__popc(need), avail);
			if (exhausted && __ballot_sync(full, alive) == 0u) {
				break;
			}
		}
		if (alive) {
			int r = wide_step(tp.wide, tp.tris, ray, cur, stack, top);
			if (r != kWideContinue) {
				bool shadowed = r == kWideHit;
				if (r == kWideStackFull) { // deeper than the lane's stack: redo this ray in reference order (does not happen on the shipped scenes)
					shadowed = !trace_any_reference(tp.nodes, tp.tris, ray.o, ray.d, overflow);
				}
				tp.shadowed[item] = shadowed ? 1 : 0;
				alive = false;
			}
		}
	}
	// one atomic per warp
	unsigned totalRays = __reduce_add_sync(full, rays), totalOverflow = __reduce_add_sync(full, overflow);
	if (lane == 0) {
		if (totalRays) atomicAdd(tp.counters + kCounterRays, (unsigned long long)totalRays);
		if (totalOverflow) atomicAdd(tp.counters + kCounterOverflow, (unsigned long long)totalOverflow);
	}
}

// ---- launcher ----------------------------------------------------------------------------------------------

template <int MODE, bool WIDE> static cudaError_t launch_mode(const TraceParams &tp, int smCount, cudaStream_t s) {
	static int blocksPerSm = 0; // same for every device of one box
	if (blocksPerSm == 0) {
		cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSm, trace_kernel<MODE, WIDE>, kTraceThreads, 0);
		if (e != cudaSuccess) {
			return e;
		}
		if (blocksPerSm < 1) blocksPerSm = 1;
	}
	unsigned long long chunks = (tp.nItems + kChunk - 1) / kChunk;
	unsigned long long wanted = (chunks + kTraceWarps - 1) / kTraceWarps;
	unsigned grid = (unsigned)std::min<unsigned long long>((unsigned long long)smCount * blocksPerSm, std::max<unsigned long long>(wanted, 1));
	trace_kernel<MODE, WIDE><<<grid, kTraceThreads, 0, s>>>(tp);
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
	const bool wide = tp.wide != nullptr;
	switch (mode) {
	case kTracePixel: return wide ? launch_mode<kTracePixel, true>(tp, smCount, s) : launch_mode<kTracePixel, false>(tp, smCount, s);
	case kTraceUnbiased: return wide ? launch_mode<kTraceUnbiased, true>(tp, smCount, s) : launch_mode<kTraceUnbiased, false>(tp, smCount, s);
	default: return wide ? launch_mode<kTraceSegments, true>(tp, smCount, s) : launch_mode<kTraceSegments, false>(tp, smCount, s);
	}
}

} // namespace restir
