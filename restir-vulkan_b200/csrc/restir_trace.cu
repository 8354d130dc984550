// restir_trace.cu — the persistent shadow-ray kernel (sm_100a).
//
// One kernel traces every shadow ray of the path: the visibility-reuse ray of restirOmni.glsl:148-160 and the pixels'
// own rays of unbiasedReuse.glsl:157-166 (kTracePixel), the neighbour rays of unbiasedReuse.glsl:139-156
// (kTraceUnbiased, launched after the own rays) and the stand-alone segments of restir_trace_segments.  The
// reference traces them inline, one thread per pixel, in whatever order the pixels come.  Here rays are work items
// of a persistent grid:
//
//   * every WARP pulls chunks of consecutive items (128, or 256 neighbour-ray slots; items are numbered by 8x4
//     screen tile, so a chunk is a compact screen patch) from a global cursor;
//   * items that need no walk are answered on the spot (item_resolve: neighbour rays of a pixel whose own ray is
//     shadowed, neighbour rays bit-identical to the neighbour's own ray — exact, see there);
//   * it sorts the rest by the light the ray is aimed at (bitonic sort of 32-bit keys in shared memory), so
//     the 32 rays a warp then walks in lockstep start next to each other AND end at the same light: they
//     visit the same nodes at the same time (one L1 tag lookup serves many lanes — the measured limiters of
//     this kernel are L1 lookups and instruction issue, not DRAM, profiles/) and finish at about the same time;
//   * each lane builds its segment from the G-buffer / reservoir (visibilityTest.glsl:1-4) and walks the tree
//     (restir_trace.cuh: two 32-byte loads per node, both boxes with packed FADD2 / FMUL2).
//
// Measured alternatives that lost (B200, Sponza 1080p, profiles/r1_f_summary.md, r1_m_summary.md): per-lane refill
// from a shared-memory ray queue (lanes decorrelate, every node load becomes its own L1 lookup: 1.8 Grays/s), a
// 4-wide re-layout of the tree (same instruction count per ray once the exact slab arithmetic is kept), a stepwise
// walk with warp votes, the stack in shared memory, a branch-free push/pop, a two-round walk with parked rays,
// 64- and 128-thread CTAs (flat), 40 registers for 6 CTAs per SM (flat; 36 and 32 lose).
//
// Per ray one byte is written: 1 = shadowed.

#include "restir_kernels.h"
#include "restir_trace.cuh"
#include "restir_wide.cuh"

namespace restir {

#ifndef RESTIR_TRACE_THREADS
#define RESTIR_TRACE_THREADS 256
#endif
constexpr int kTraceThreads = RESTIR_TRACE_THREADS;
constexpr int kTraceWarps = kTraceThreads / 32;
#ifndef RESTIR_TRACE_CHUNK
#define RESTIR_TRACE_CHUNK 128
#endif
// items per warp fetch (power of two, <= 1024: the local id is the low kLocalBits of the sort key).  Neighbour rays: about half of the
// items of a chunk are answered without a ray (item_resolve), so the chunk is twice as long to keep the batches full.
constexpr int kChunk = RESTIR_TRACE_CHUNK;
#ifndef RESTIR_TRACE_CHUNK_NEIGHBOURS
#define RESTIR_TRACE_CHUNK_NEIGHBOURS 256
#endif
template <int MODE> struct ChunkOf {
	static constexpr int value = MODE == kTraceUnbiased ? RESTIR_TRACE_CHUNK_NEIGHBOURS : kChunk;
};
#ifndef RESTIR_TRACE_SORT
#define RESTIR_TRACE_SORT 1
#endif
#ifndef RESTIR_TRACE_SHARE
#define RESTIR_TRACE_SHARE 1 // answer a neighbour ray from the neighbour's own ray when the two segments are identical
#endif
constexpr unsigned kInvalidKey = 0xffffffffu;
constexpr unsigned kLocalBits = 10, kLocalMask = (1u << kLocalBits) - 1u, kLightMask = (1u << (31 - kLocalBits)) - 1u; // sort key = light << kLocalBits | position in the chunk

// What a trace kernel walks: the uploaded 80-byte nodes in the reference's order, their 64-byte binary image, or the 4-wide
// quantised image (restir_wide.cuh) with the binary image for the rays outside its range.
enum TraceWalk { kWalkReference = 0, kWalkImage = 1, kWalkWide = 2 };

__device__ __forceinline__ unsigned fast_div(unsigned n, const FastDiv &f) { return f.M == 0ull ? n : (unsigned)__umul64hi((unsigned long long)n, f.M); }

// ---- item -> pixel, key, segment ---------------------------------------------------------------------------

__device__ __forceinline__ bool item_pixel(const TraceParams &tp, unsigned p, size_t &pix) {
	// p is a tile-ordered pixel id: (tile, lane) with 8x4 tiles, tilesX tiles per tile row (restir_kernels.cu
	// tile_pixel_id produces the same numbering)
	unsigned tile = p >> 5, l = p & 31u;
	unsigned ty = fast_div(tile, tp.divTilesX), tx = tile - ty * tp.tilesX;
	int x = (int)(tx * 8u + (l & 7u));
	int y = tp.band.rowBegin + (int)(ty * 4u + (l >> 3));
	if (x >= tp.band.W || y >= tp.band.rowEnd) {
		return false;
	}
	pix = (size_t)(y - tp.band.allocBegin) * (size_t)tp.band.W + (size_t)x;
	return true;
}

// Tile-ordered pixel id of local pixel index n (the inverse of item_pixel); false outside the band's own rows.
__device__ __forceinline__ bool pixel_id_of_local(const TraceParams &tp, unsigned n, unsigned &id) {
	unsigned W = (unsigned)tp.band.W;
	unsigned yl = fast_div(n, tp.divW), x = n - yl * W;
	int y = (int)yl + tp.band.allocBegin;
	if (y < tp.band.rowBegin || y >= tp.band.rowEnd) {
		return false;
	}
	unsigned ry = (unsigned)(y - tp.band.rowBegin);
	id = (((ry >> 2) * tp.tilesX + (x >> 3)) << 5) | ((ry & 3u) << 3) | (x & 7u);
	return true;
}

// What the reference does for an item, and what is left of it here.
enum ItemState {
	kItemNone = 0,     // the reference makes no testVisibility call for this item
	kItemAnswered = 1, // it does, and the answer is known without walking the tree (see item_resolve)
	kItemRay = 2,      // it does, and the segment has to be traced
};

// Resolves an item to its state, the pixel holding the sample (pix), the pixel the ray starts from (opix) and
// the index of its visibility byte (out).
//
// kTraceUnbiased items are the NEIGHBOUR rays of unbiasedReuse.glsl:139-156; the pixels' own rays (:157-166)
// were traced by the kTracePixel launch before this one.  Two exact shortcuts:
//   * the pixel's own ray is shadowed: the reference zeroes numSamples after the neighbour loop (:160-165), so
//     no neighbour ray of that pixel can change the result;
//   * the neighbour's own merged sample sits at the same position, bit for bit, as this pixel's (spatial reuse
//     makes neighbours share samples; with point lights every reservoir holding light k does): the segment
//     neighbour -> sample IS the neighbour's own ray (same p1, same p2 => same segment_setup, same walk), whose
//     answer the kTracePixel launch already wrote.
template <int MODE> __device__ __forceinline__ int item_resolve(const TraceParams &tp, unsigned item, size_t &pix, size_t &opix, size_t &out, bool answer) {
	if (MODE == kTracePixel) {
		if (!item_pixel(tp, item, pix)) {
			return kItemNone;
		}
		opix = pix;
		out = (size_t)item * tp.outStride + tp.outOffset;
		return kItemRay;
	}
	unsigned p = fast_div(item, tp.divSlots), slot = item - p * tp.slots;
	if (!item_pixel(tp, p, pix)) {
		return kItemNone;
	}
	int n = tp.neighborPix[(size_t)p * tp.slots + slot];
	if (n < 0) {
		return kItemNone;
	}
	opix = (size_t)n;
	const size_t row = (size_t)p * (tp.slots + 1);
	out = row + slot;
	if (!tp.elide) {
		return kItemRay;
	}
	if (tp.shadowed[row + tp.slots] != 0) {
		return kItemAnswered; // the byte is never read (unbiased_finalize_kernel drops the pixel)
	}
	unsigned nid;
	if (RESTIR_TRACE_SHARE && pixel_id_of_local(tp, (unsigned)n, nid)) {
		float4 a = __ldg(reinterpret_cast<const float4 *>(tp.reservoirs + pix));
		float4 b = __ldg(reinterpret_cast<const float4 *>(tp.reservoirs + opix));
		if (__float_as_uint(a.x) == __float_as_uint(b.x) && __float_as_uint(a.y) == __float_as_uint(b.y) && __float_as_uint(a.z) == __float_as_uint(b.z)) {
			if (answer) {
				tp.shadowed[out] = tp.shadowed[(size_t)nid * (tp.slots + 1) + tp.slots];
			}
			return kItemAnswered;
		}
	}
	return kItemRay;
}

// ---- occluder cache -------------------------------------------------------------------------------------------------------
// 86 % of restirOmni's selected candidates are shadowed (Sponza, 200 lights), and the triangle that shadows one pixel from a
// light shadows most of its neighbourhood from that light: measured on the CPU over three frames, the last occluder found for
// (64 x 32-pixel screen region, light) answers 3 of 4 shadowed rays of the next frame (profiles/r2_j_summary.md).  A shadowed
// answer needs ONE witness — (*) of wide_image.h: the triangle hit with the reference's arithmetic and its leaf box passed with
// the reference's arithmetic — so before a ray is queued for a walk the cached triangle of its (region, light) is tested,
// exactly; a hit answers the ray (and frees its lane: the sort moves answered items behind the rays), a miss costs one
// triangle test.  Walks that find an occluder record it.  The table only ever proposes witnesses, so its contents (racy
// plain stores, stale entries of earlier frames) cannot change a bit of the result.
// Entry = tag (bits 8..15 of the light index) << 24 | triangle record; 0xffffffff = empty (records stay below 2^24).
// Every (region, key) holds kOccluderWays = 2 entries, most recent first: where two occluders share a region's view of a light (a
// column in front of a wall) one entry made them evict each other.  A walk that finds an occluder other than the first way's moves
// that one to the second way; tp.occluderPretest says how many ways a launch tests (restirOmni's rays both: 0.260 -> 0.230 ms; the
// unbiased pass's mostly unshadowed rays the first only, profiles/r2_s_ways.log).
// With more lights than a region's entries can tell apart (tp.occluderByDirection: the 1 M-light frames, where a (region, light)
// pair practically never comes back) the entry is chosen by the DIRECTION of the segment instead — cube face and an 8 x 8 grid
// on it, 384 cells: rays of a region that leave in the same direction meet the same nearby geometry whatever light they aim at.
// CPU estimate (960 x 540, 1 M lights, 64 candidates): 36 % of the shadowed restirOmni rays answered, against 12 % by light index;
// with 200 lights it is the other way round (64 % against 76 %).
constexpr unsigned kOccluderSlots = 256;
__device__ __forceinline__ unsigned occluder_key(const TraceParams &tp, unsigned light, f3 p1, f3 p2) { // 16 bits: entry | tag << 8
	if (!tp.occluderByDirection) {
		return light & 0xffffu;
	}
	const f3 dir = p2 - p1;
	const float ax = fabsf(dir.x), ay = fabsf(dir.y), az = fabsf(dir.z);
	unsigned face;
	float u, v, m;
	if (ax >= ay && ax >= az) {
		face = dir.x > 0.0f ? 0u : 1u; m = ax; u = dir.y; v = dir.z;
	} else if (ay >= az) {
		face = dir.y > 0.0f ? 2u : 3u; m = ay; u = dir.x; v = dir.z;
	} else {
		face = dir.z > 0.0f ? 4u : 5u; m = az; u = dir.x; v = dir.y;
	}
	const float s = __fdividef(4.0f, m); // a heuristic key: any value is as correct as any other
	const unsigned iu = min((unsigned)max((int)(u * s + 4.0f), 0), 7u), iv = min((unsigned)max((int)(v * s + 4.0f), 0), 7u);
	return (face * 8u + iu) * 8u + iv;
}
__device__ __forceinline__ unsigned occluder_entry(const TraceParams &tp, size_t opix, unsigned key16) {
	const unsigned W = (unsigned)tp.band.W, n = (unsigned)opix;
	const unsigned yl = fast_div(n, tp.divW), x = n - yl * W;
	return (((yl >> 5) * tp.regionsX + (x >> 6)) * kOccluderSlots + (key16 & (kOccluderSlots - 1u))) * kOccluderWays; // the first way
}

// ---- one walk per segment ---------------------------------------------------------------------------------------------------
// Every pixel is the neighbour of about five others, and spatial reuse makes the pixels of a region hold the same few samples:
// a third of the neighbour rays that survive item_resolve ask for a segment (neighbour's position -> sample position) that
// another pixel's item asks for too (measured on the CPU: profiles/r2_j_summary.md).  The first item to enter a segment into
// an open-addressing table (atomicCAS on a 32-bit word: empty -> item number) walks it; a later item that finds an item with
// the same origin pixel and, bit for bit, the same sample position there is its ALIAS: it records (its visibility byte, the
// owner's) and alias_resolve_kernel copies the answer after the walks.  Same p1, same p2 => same segment_setup, same walk:
// exact.  A full table or alias list only means the ray is walked.
__device__ __forceinline__ unsigned segment_hash(unsigned opix, float4 t) {
	unsigned h = opix * 0x9E3779B1u ^ __float_as_uint(t.x) * 0x85EBCA6Bu ^ __float_as_uint(t.y) * 0xC2B2AE35u ^ __float_as_uint(t.z) * 0x27D4EB2Fu;
	h ^= h >> 15;
	h *= 0x2C1B3C6Du;
	return h ^ (h >> 13);
}
// true: `item` is an alias of `owner` (the item whose walk answers it).  Single exit, no return inside the loop: with early
// returns around the atomic the compiler wrapped the whole chunk body in one convergence barrier and the lockstep batches ran
// in pieces (trace_kernel<neighbours> 1.00 -> 2.38 ms with the table switched OFF; measured, profiles/r2_j_summary.md).
__device__ __forceinline__ bool segment_claim(const TraceParams &tp, unsigned item, size_t pix, size_t opix, unsigned &owner) {
	const float4 t = __ldg(reinterpret_cast<const float4 *>(tp.reservoirs + pix));
	unsigned h = segment_hash((unsigned)opix, t) & tp.dedupeMask;
	bool alias = false, settled = false;
#pragma unroll
	for (int probe = 0; probe < 4; ++probe) { // straight-line code: a LOOP around the atomic makes ptxas give up on reconvergence (see above)
		if (settled) {
			continue;
		}
		const unsigned e = atomicCAS(tp.dedupe + h, 0xffffffffu, item);
		if (e == 0xffffffffu) {
			settled = true; // first: this item walks the segment
		} else {
			const unsigned pe = fast_div(e, tp.divSlots);
			size_t pixe = 0;
			const bool inside = item_pixel(tp, pe, pixe);
			const float4 te = __ldg(reinterpret_cast<const float4 *>(tp.reservoirs + pixe));
			if (inside && (size_t)tp.neighborPix[e] == opix && __float_as_uint(te.x) == __float_as_uint(t.x) && __float_as_uint(te.y) == __float_as_uint(t.y) &&
			    __float_as_uint(te.z) == __float_as_uint(t.z)) {
				owner = e;
				alias = settled = true;
			}
			h = (h + 1u) & tp.dedupeMask;
		}
	}
	return alias; // not settled after four probes: a crowded neighbourhood, the ray is walked
}

// Sort key of an item: (light the ray is aimed at, position in the chunk); kInvalidKey = nothing to trace.  aliasOf: set to the
// owner item when this item turns out to be an alias (then the key is valid until the caller has found room in the alias list).
template <int MODE, int WALK> __device__ __forceinline__ unsigned item_key(const TraceParams &tp, unsigned item, unsigned local, unsigned &answered, unsigned &cached,
                                                                         unsigned &aliasOf) {
	unsigned key = kInvalidKey;
	if (MODE == kTraceSegments) {
		key = item < tp.nItems ? local : kInvalidKey;
	} else if (item < tp.nItems) {
		size_t pix, opix, out;
		const int state = item_resolve<MODE>(tp, item, pix, opix, out, true);
		answered += state == kItemAnswered ? 1u : 0u;
		if (state == kItemRay) {
			const unsigned light = (unsigned)__ldg(reinterpret_cast<const int *>(tp.reservoirs + pix) + 3); // PackedReservoir::lightIndex
			bool witnessed = false;
			bool alias = false;
			if (MODE == kTraceUnbiased && WALK == kWalkWide && tp.dedupe != nullptr) {
				alias = segment_claim(tp, item, pix, opix, aliasOf);
			}
			if (WALK == kWalkWide && tp.occluders != nullptr && tp.occluderPretest && !alias) {
				float4 w = make_float4(0.0f, 0.0f, 0.0f, 0.0f), t = w;
				if (tp.occluderByDirection) { // the key needs the segment
					w = __ldg(tp.worldPos + opix);
					t = __ldg(reinterpret_cast<const float4 *>(tp.reservoirs + pix));
				}
				const unsigned key16 = occluder_key(tp, light, mk3(w.x, w.y, w.z), mk3(t.x, t.y, t.z));
				const unsigned *entry = tp.occluders + occluder_entry(tp, opix, key16);
				unsigned e = __ldcg(entry), e2 = 0xffffffffu;
				if (kOccluderWays > 1 && tp.occluderPretest > 1) { // the witness before the last one: a second chance where two occluders share a region's view of a light
					e2 = __ldcg(entry + 1);
				}
				unsigned rec = e & 0xffffffu;
				if ((e >> 24) == (key16 >> 8) && rec < tp.nTris) {
					if (!tp.occluderByDirection) {
						w = __ldg(tp.worldPos + opix);
						t = __ldg(reinterpret_cast<const float4 *>(tp.reservoirs + pix));
					}
					f3 o, d;
					segment_setup(mk3(w.x, w.y, w.z), mk3(t.x, t.y, t.z), o, d);
					if (wide_ray_in_range(tp.grid, o, d)) {
						bool hit = wide_leaf_hit(tp.triEdges, rec, o, d);
						rec = e2 & 0xffffffu;
						if (!hit && kOccluderWays > 1 && (e2 >> 24) == (key16 >> 8) && rec < tp.nTris) {
							hit = wide_leaf_hit(tp.triEdges, rec, o, d);
						}
						if (hit) {
							tp.shadowed[out] = 1;
							cached++;
							witnessed = true;
						}
					}
				}
			}
			if (!witnessed) { // (an alias keeps a valid key until the caller has recorded it)
				key = (((RESTIR_TRACE_SORT ? light : 0u) & kLightMask) << kLocalBits) | local;
			}
		}
	}
	return key;
}

// Segment of an item item_key found to be a ray; returns the index of its visibility byte.
template <int MODE> __device__ __forceinline__ size_t item_segment(const TraceParams &tp, unsigned item, f3 &p1, f3 &p2, size_t &opix) {
	if (MODE == kTraceSegments) {
		const float *a = tp.segP1 + (size_t)item * 3, *b = tp.segP2 + (size_t)item * 3;
		p1 = mk3(a[0], a[1], a[2]);
		p2 = mk3(b[0], b[1], b[2]);
		opix = 0;
		return item;
	}
	size_t pix, out;
	if (MODE == kTracePixel) {
		item_resolve<MODE>(tp, item, pix, opix, out, false);
	} else {
		unsigned p = fast_div(item, tp.divSlots), slot = item - p * tp.slots;
		item_pixel(tp, p, pix);
		opix = (size_t)tp.neighborPix[(size_t)p * tp.slots + slot];
		out = (size_t)p * (tp.slots + 1) + slot;
	}
	float4 w = __ldg(tp.worldPos + opix);
	float4 t = __ldg(reinterpret_cast<const float4 *>(tp.reservoirs + pix));
	p1 = mk3(w.x, w.y, w.z);
	p2 = mk3(t.x, t.y, t.z);
	return out;
}

// bitonic sort of CHUNK keys in shared memory by one warp
template <int CHUNK> __device__ __forceinline__ void warp_sort(unsigned *keys, unsigned lane) {
#pragma unroll 1
	for (unsigned k = 2; k <= (unsigned)CHUNK; k <<= 1) {
#pragma unroll 1
		for (unsigned j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
			for (unsigned t = lane; t < (unsigned)CHUNK / 2; t += 32) {
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

// The same network over the first n = 32 R keys with the keys in REGISTERS (key i = r * 32 + lane lives in register r of its lane):
// a compare-exchange at distance j >= 32 pairs two registers of one lane, at distance j < 32 one shuffle fetches the partner's
// key and the lane keeps the smaller or the larger — no shared-memory round trip, no index arithmetic, no warp barrier per
// stage (n = 128: ~340 instructions against ~1 000 for the shared-memory loop; capture N: the sort was 8.5 % of the neighbour
// kernel).  Fully unrolled per R.
template <int R> __device__ __forceinline__ void warp_sort_regs(unsigned *keys, unsigned lane) {
	unsigned v[R];
#pragma unroll
	for (int r = 0; r < R; ++r) {
		v[r] = keys[r * 32 + lane];
	}
#pragma unroll
	for (int k = 2; k <= R * 32; k <<= 1) {
#pragma unroll
		for (int j = k >> 1; j > 0; j >>= 1) {
			if (j >= 32) {
#pragma unroll
				for (int r = 0; r < R; ++r) {
					const int pr = r ^ (j / 32);
					if (pr > r) {
						const bool ascending = ((r * 32) & k) == 0; // k >= 64 here: the bit is one of r's
						const unsigned lo = min(v[r], v[pr]), hi = max(v[r], v[pr]);
						v[r] = ascending ? lo : hi;
						v[pr] = ascending ? hi : lo;
					}
				}
			} else {
#pragma unroll
				for (int r = 0; r < R; ++r) {
					const unsigned other = __shfl_xor_sync(0xffffffffu, v[r], j);
					const bool ascending = (((unsigned)(r * 32) + lane) & (unsigned)k) == 0u;
					const bool lower = (lane & (unsigned)j) == 0u; // this lane holds the lower index of the pair
					v[r] = lower == ascending ? min(v[r], other) : max(v[r], other);
				}
			}
		}
	}
#pragma unroll
	for (int r = 0; r < R; ++r) {
		keys[r * 32 + lane] = v[r];
	}
}
// n in {32, 64, 128, 256, ...}: the register network up to 256 keys, the shared-memory loop beyond
__device__ __forceinline__ void warp_sort_n(unsigned *keys, unsigned n, unsigned lane) {
	switch (n) {
	case 32: warp_sort_regs<1>(keys, lane); break;
	case 64: warp_sort_regs<2>(keys, lane); break;
	case 128: warp_sort_regs<4>(keys, lane); break;
	case 256: warp_sort_regs<8>(keys, lane); break;
	default:
#pragma unroll 1
		for (unsigned k = 2; k <= n; k <<= 1) {
#pragma unroll 1
			for (unsigned j = k >> 1; j > 0; j >>= 1) {
#pragma unroll 1
				for (unsigned t = lane; t < n / 2; t += 32) {
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
	__syncwarp();
}

// ---- the kernel ------------------------------------------------------------------------------------------

#ifndef RESTIR_TRACE_MIN_BLOCKS
#define RESTIR_TRACE_MIN_BLOCKS 5
#endif
// RESTIR_TRACE_AFFINE 1 (experiment): chunks are not handed out from one global cursor but from one cursor per REGION of the work
// list — a contiguous run of chunks, i.e. a compact part of the screen — and the CTAs resident on one SM own neighbouring regions
// (region = %smid x CTAs per SM + arrival order on that SM), so the rays an SM walks one after the other cross the same part of
// the scene and find its deep nodes in that SM's L1 (70 % sector hits with the global cursor).  A CTA whose region is exhausted
// moves on to the following regions (work stealing), so every chunk is processed whatever the mapping.  0: one global cursor.
#ifndef RESTIR_TRACE_AFFINE
#define RESTIR_TRACE_AFFINE 0
#endif
// RESTIR_TRACE_REFILL = T > 0 (experiment, off): a warp does not wait for its last rays.  When T or fewer lanes are still
// walking, the idle lanes take the next rays of the sorted chunk (which start at the root together and are aimed at the
// same light) while the survivors keep their lanes and their stacks.  T = 0 (default): lockstep batches of 32.
// Measured (profiles/r2_a_trace_ab.md): T = 4 / 8 / 16 / 24 give 3.78 / 3.74 / 3.73 / 3.73 ms per frame against 3.46 in
// lockstep — restirOmni's rays (11 of 32 lanes in lockstep) gain 10 % between T = 4 and T = 24, but the neighbour rays
// lose as much (survivors deep in the tree next to fresh rays at the root: more distinct node loads per instruction) and
// the loop that can resume a walk costs 20 % over the one that cannot.
#ifndef RESTIR_TRACE_REFILL
#define RESTIR_TRACE_REFILL 0
#endif

// the binary walk for the rays the wide walk does not take (non-finite or out-of-range origin / direction: wide_image.h).
// Inlined: as a __noinline__ call it crashes ptxas 12.9 (segmentation fault at every -O level).
static __device__ __forceinline__ bool trace_any_image_call(const float4 *__restrict__ image, const float4 *__restrict__ tris, f3 o, f3 d) {
	return trace_any_image(image, tris, o, d, nullptr);
}

#ifndef RESTIR_TRACE_MIN_BLOCKS_WIDE
#define RESTIR_TRACE_MIN_BLOCKS_WIDE 4
#endif
// RESTIR_TRACE_GUIDED = g > 0: guided self-scheduling of the wide kernel's work list.  A warp takes g x (items left / warps of the
// grid) items, at most a chunk, at least RESTIR_TRACE_GUIDED_MIN: with 3-9 chunks per warp and ~0.1 ms per chunk the warps of the
// fixed-chunk kernel finish up to a chunk apart and the launch ends with its last warp.  0: every fetch is a full chunk.
#ifndef RESTIR_TRACE_GUIDED
#define RESTIR_TRACE_GUIDED 0
#endif
#ifndef RESTIR_TRACE_GUIDED_MIN
#define RESTIR_TRACE_GUIDED_MIN 32
#endif
template <int MODE, int WALK> __global__ void __launch_bounds__(kTraceThreads, RESTIR_TRACE_MIN_BLOCKS) trace_kernel(const __grid_constant__ TraceParams tp) {
	static_assert(WALK != kWalkWide, "the wide image is walked by trace_wide_kernel");
	constexpr bool IMAGE = WALK == kWalkImage;
	constexpr int CHUNK = ChunkOf<MODE>::value;
	__shared__ unsigned allKeys[kTraceWarps][CHUNK];
#if RESTIR_TRACE_TOP_SMEM > 0
	__shared__ float4 topNodes[RESTIR_TRACE_TOP_SMEM * 4];
	if (IMAGE) { // the first nodes of the image = the top of the tree (breadth-first numbering), once per CTA of the persistent grid
		const unsigned count = min((unsigned)RESTIR_TRACE_TOP_SMEM, tp.nNodes) * 4u;
		for (unsigned t = threadIdx.x; t < count; t += kTraceThreads) {
			topNodes[t] = __ldg(tp.image + t);
		}
		__syncthreads();
	}
	const float4 *const topOfTree = IMAGE ? topNodes : nullptr;
#else
	const float4 *const topOfTree = nullptr;
#endif
	const unsigned lane = threadIdx.x & 31u;
	unsigned *keys = allKeys[threadIdx.x >> 5];
	const unsigned full = 0xffffffffu;
	unsigned rays = 0, answered = 0, overflow = 0, cached = 0;
	const float4 *const walkTris = RESTIR_TRACE_TRI_EDGES ? tp.triEdges : tp.tris;

#if RESTIR_TRACE_AFFINE
	__shared__ unsigned homeRegion;
	const unsigned nRegions = gridDim.x;
	const unsigned nChunks = (unsigned)((tp.nItems + CHUNK - 1) / CHUNK);
	if (threadIdx.x == 0) {
		unsigned smid;
		asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
		unsigned slot = atomicAdd(tp.smSlots + smid, 1u);
		homeRegion = (smid * tp.blocksPerSm + slot) % nRegions;
	}
	__syncthreads();
	unsigned region = homeRegion, visited = 0;
#endif
	for (;;) {
		unsigned base = 0;
#if RESTIR_TRACE_AFFINE
		// chunks [regionBegin, regionEnd) of region r; its cursor counts the chunks handed out
		const unsigned regionBegin = (unsigned)((unsigned long long)nChunks * region / nRegions);
		const unsigned regionEnd = (unsigned)((unsigned long long)nChunks * (region + 1) / nRegions);
		unsigned taken = 0;
		if (lane == 0) {
			taken = atomicAdd(tp.regionCursors + region, 1u);
		}
		taken = __shfl_sync(full, taken, 0);
		if (taken >= regionEnd - regionBegin) { // exhausted: help the next region
			if (++visited >= nRegions) {
				break;
			}
			region = region + 1 == nRegions ? 0 : region + 1;
			continue;
		}
		base = (regionBegin + taken) * (unsigned)CHUNK;
#else
		if (lane == 0) {
			base = (unsigned)min(atomicAdd(tp.counters + kCounterWork, (unsigned long long)CHUNK), 0xffffffffull);
		}
		base = __shfl_sync(full, base, 0);
		if (base >= tp.nItems) {
			break;
		}
#endif
		unsigned valid = 0;
#pragma unroll 1
		for (unsigned r = 0; r < (unsigned)CHUNK / 32; ++r) {
			unsigned local = r * 32u + lane;
			unsigned aliasOf = 0xffffffffu; // (aliases exist in trace_wide_kernel only)
			unsigned key = item_key<MODE, WALK>(tp, base + local, local, answered, cached, aliasOf);
			keys[local] = key;
			valid += __popc(__ballot_sync(full, key != kInvalidKey));
		}
		__syncwarp();
		if (valid == 0) {
			continue;
		}
		constexpr bool kSorted = MODE != kTraceSegments && RESTIR_TRACE_SORT;
		if (kSorted) {
			warp_sort<CHUNK>(keys, lane); // holes (kInvalidKey) end up behind the `valid` rays
		}
		if (IMAGE && RESTIR_TRACE_REFILL > 0) {
			// rays of the chunk in key order; in an unsorted chunk holes are skipped as they come
			const unsigned count = kSorted ? valid : (unsigned)CHUNK;
			unsigned next = 0;
			bool walking = false;
			size_t out = 0;
			WalkRay ray;
			int stack[32];
			int top = 0, cur = 0;
			for (;;) {
				unsigned act = __ballot_sync(full, walking);
				if ((unsigned)__popc(act) <= (unsigned)RESTIR_TRACE_REFILL && next < count) {
					unsigned idle = ~act;
					unsigned mine = next + (unsigned)__popc(idle & ((1u << lane) - 1u));
					if (!walking && mine < count) {
						unsigned key = keys[mine];
						if (key != kInvalidKey) {
							f3 p1, p2, o, d;
							size_t opix;
							out = item_segment<MODE>(tp, base + (key & kLocalMask), p1, p2, opix);
							segment_setup(p1, p2, o, d);
							ray = walk_ray(o, d);
							top = 0;
							cur = 0;
							walking = true;
							rays++;
						}
					}
					next += (unsigned)__popc(idle);
					act = __ballot_sync(full, walking);
				}
				if (act == 0u) {
					if (next >= count) break;
					continue;
				}
				if (walking) {
					int st = walk_step(tp.image, walkTris, ray, cur, top, stack, topOfTree);
					if (st != kWalkOn) {
						tp.shadowed[out] = st == kWalkHit ? 1 : 0;
						walking = false;
					}
				}
			}
		} else {
#pragma unroll 1
			for (unsigned r = 0; r < (unsigned)CHUNK / 32; ++r) {
				unsigned key = keys[r * 32u + lane];
				if (__ballot_sync(full, key != kInvalidKey) == 0u) {
					if (kSorted) break; // sorted: only holes follow
					continue;
				}
				if (key != kInvalidKey) {
					unsigned item = base + (key & kLocalMask);
					f3 p1, p2, o, d;
					size_t opix;
					size_t out = item_segment<MODE>(tp, item, p1, p2, opix);
					segment_setup(p1, p2, o, d);
					bool clear = IMAGE ? trace_any_image(tp.image, walkTris, o, d, topOfTree) : trace_any_reference(tp.nodes, tp.tris, o, d, overflow);
					tp.shadowed[out] = clear ? 0 : 1;
					rays++;
				}
			}
		}
		__syncwarp();
	}
	// one atomic per warp
	rays = __reduce_add_sync(full, rays);
	answered = __reduce_add_sync(full, answered + cached);
	cached = __reduce_add_sync(full, cached);
	overflow = __reduce_add_sync(full, overflow);
	if (lane == 0) {
		if (rays + answered) atomicAdd(tp.counters + kCounterRays, (unsigned long long)(rays + answered));
		if (rays) atomicAdd(tp.counters + kCounterTraced, (unsigned long long)rays);
		if (overflow) atomicAdd(tp.counters + kCounterOverflow, (unsigned long long)overflow);
		if (cached) atomicAdd(tp.counters + kCounterCached, (unsigned long long)cached);
	}
}

// ---- the kernel that walks the wide image ---------------------------------------------------------------------------------------
// Same persistent structure (chunks from a global cursor, items resolved and keyed, sorted by light, lockstep batches), built
// around the walk's register budget: at 64 registers (4 CTAs per SM) everything that lives across a walk competes with the
// walk's own operands, so the segment's origin and direction wait in shared memory for the leaf tests, the counters are per-warp
// words in shared memory fed by votes, and what a finished ray needs (its visibility byte, its cache entry) is derived again
// from its key.  Rays outside the wide walk's range (wide_image.h) walk the binary image after the batch.
template <int MODE> __global__ void __launch_bounds__(kTraceThreads, RESTIR_TRACE_MIN_BLOCKS_WIDE) trace_wide_kernel(const __grid_constant__ TraceParams tp) {
	constexpr int CHUNK = ChunkOf<MODE>::value;
	constexpr int WALK = kWalkWide;
	__shared__ unsigned allKeys[kTraceWarps][CHUNK];
	__shared__ float rayOD[6][kTraceThreads];
	__shared__ unsigned stats[kTraceWarps][4]; // rays walked, answered without a walk (cached included), cached, unused
	const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	unsigned *keys = allKeys[warp];
	const unsigned full = 0xffffffffu;
	if (lane < 4) {
		stats[warp][lane] = 0;
	}
	__syncwarp();
	for (;;) {
		unsigned base = 0, size = (unsigned)CHUNK;
		if (lane == 0) {
#if RESTIR_TRACE_GUIDED > 0
			// the end of the work list is handed out in smaller pieces: what a warp takes is its share of what is left (times
			// RESTIR_TRACE_GUIDED), rounded down to a power of two, between 32 items and the full chunk
			unsigned long long seen;
			asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(tp.counters + kCounterWork));
			const float left = seen < tp.nItems ? (float)(tp.nItems - seen) : 0.0f;
			const unsigned share = (unsigned)min(left * tp.guidedShare, (float)CHUNK);
			if (share < (unsigned)CHUNK) {
				size = max(1u << (31 - __clz((int)max(share, 1u))), (unsigned)RESTIR_TRACE_GUIDED_MIN);
			}
#endif
			base = (unsigned)min(atomicAdd(tp.counters + kCounterWork, (unsigned long long)size), 0xffffffffull);
		}
		base = __shfl_sync(full, base, 0);
		if (base >= tp.nItems) {
			break;
		}
#if RESTIR_TRACE_GUIDED > 0
		size = __shfl_sync(full, size, 0);
#endif
		unsigned valid = 0;
#pragma unroll 1
		for (unsigned r = 0; r * 32u < size; ++r) {
			unsigned local = r * 32u + lane;
			unsigned answered = 0, cached = 0, aliasOf = 0xffffffffu;
			unsigned key = item_key<MODE, WALK>(tp, base + local, local, answered, cached, aliasOf);
			if (MODE == kTraceUnbiased) { // aliases: one slot of the list each, reserved with one atomic per warp
				const unsigned isAlias = __ballot_sync(full, aliasOf != 0xffffffffu);
				if (isAlias != 0u) {
					unsigned first = 0;
					if (lane == 0) {
						first = atomicAdd(tp.aliasCount, (unsigned)__popc(isAlias));
					}
					first = __shfl_sync(full, first, 0) + (unsigned)__popc(isAlias & ((1u << lane) - 1u));
					if (aliasOf != 0xffffffffu && first < tp.aliasCapacity) {
						const unsigned item = base + local, p = fast_div(item, tp.divSlots), pe = fast_div(aliasOf, tp.divSlots);
						tp.aliases[first] = make_uint2(p * (tp.slots + 1u) + (item - p * tp.slots), pe * (tp.slots + 1u) + (aliasOf - pe * tp.slots));
						key = kInvalidKey;
						answered = 1;
					}
				}
			}
			// the rays of the chunk are packed at the front of the key array: what is sorted is the rays, not the chunk (half of a
			// neighbour chunk and three quarters of a restirOmni chunk are answered without a walk)
			const unsigned isValid = __ballot_sync(full, key != kInvalidKey);
			if (key != kInvalidKey) {
				keys[valid + (unsigned)__popc(isValid & ((1u << lane) - 1u))] = key;
			}
			const unsigned nValid = __popc(isValid);
			const unsigned nAnswered = __popc(__ballot_sync(full, answered + cached != 0u)), nCached = __popc(__ballot_sync(full, cached != 0u));
			valid += nValid;
			if (lane == 0) {
				stats[warp][0] += nValid;
				stats[warp][1] += nAnswered;
				stats[warp][2] += nCached;
			}
		}
		__syncwarp();
		if (valid == 0) {
			continue;
		}
		constexpr bool kSorted = MODE != kTraceSegments && RESTIR_TRACE_SORT;
		if (kSorted && valid > 1) {
			unsigned n = 32;
			while (n < valid) {
				n <<= 1;
			}
			for (unsigned t = valid + lane; t < n; t += 32) {
				keys[t] = kInvalidKey; // padding sorts behind every ray
			}
			__syncwarp();
			warp_sort_n(keys, n, lane);
		}
#pragma unroll 1
		for (unsigned r = 0; r * 32u < valid; ++r) {
			const unsigned key = r * 32u + lane < valid ? keys[r * 32u + lane] : kInvalidKey;
			int rec = -1;       // >= 0: the record of the occluding triangle
			bool unsafe = false; // the wide walk does not take this ray
			if (key != kInvalidKey) {
				f3 p1, p2, o, d;
				size_t opix;
				item_segment<MODE>(tp, base + (key & kLocalMask), p1, p2, opix);
				segment_setup(p1, p2, o, d);
				float *od = &rayOD[0][threadIdx.x];
				od[0] = o.x; od[kTraceThreads] = o.y; od[2 * kTraceThreads] = o.z;
				od[3 * kTraceThreads] = d.x; od[4 * kTraceThreads] = d.y; od[5 * kTraceThreads] = d.z;
				WideLaneRay wr;
				if (wide_lane_setup(tp.grid, o, d, wr)) {
					rec = trace_any_wide(tp.wide, tp.triEdges, wr, od, kTraceThreads);
				} else {
					unsafe = true;
				}
			}
			if (__any_sync(full, unsafe)) { // rare: non-finite or out-of-range origin / direction
				if (unsafe) {
					const float *od = &rayOD[0][threadIdx.x];
					const f3 o = mk3(od[0], od[kTraceThreads], od[2 * kTraceThreads]), d = mk3(od[3 * kTraceThreads], od[4 * kTraceThreads], od[5 * kTraceThreads]);
					rec = trace_any_image_call(tp.image, tp.triEdges, o, d) ? -1 : -2;
				}
			}
			if (key != kInvalidKey) {
				f3 p1, p2;
				size_t opix;
				const size_t out = item_segment<MODE>(tp, base + (key & kLocalMask), p1, p2, opix);
				tp.shadowed[out] = rec != -1 ? 1 : 0;
				if (MODE != kTraceSegments && rec >= 0 && tp.occluders != nullptr) { // the witness for the next ray of this region at this light
					const unsigned key16 = occluder_key(tp, key >> kLocalBits, p1, p2);
					unsigned *entry = tp.occluders + occluder_entry(tp, opix, key16);
					const unsigned fresh = ((key16 >> 8) << 24) | (unsigned)rec;
					if (kOccluderWays > 1) { // most recent first; the one it displaces moves to the second way
						const unsigned last = __ldcg(entry);
						if (last != fresh) {
							entry[1] = last;
							entry[0] = fresh;
						}
					} else {
						entry[0] = fresh;
					}
				}
			}
		}
		__syncwarp();
	}
	__syncwarp();
	if (lane == 0) { // one atomic per warp and counter
		const unsigned rays = stats[warp][0], answered = stats[warp][1], cached = stats[warp][2];
		if (rays + answered) atomicAdd(tp.counters + kCounterRays, (unsigned long long)(rays + answered));
		if (rays) atomicAdd(tp.counters + kCounterTraced, (unsigned long long)rays);
		if (cached) atomicAdd(tp.counters + kCounterCached, (unsigned long long)cached);
	}
}

// the aliases take their owners' answers (after every walk of the launch before it)
__global__ void __launch_bounds__(256) alias_resolve_kernel(const uint2 *__restrict__ aliases, const unsigned *__restrict__ count, unsigned capacity,
                                                            unsigned char *__restrict__ shadowed) {
	const unsigned n = min(*count, capacity);
	for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const uint2 a = aliases[i];
		shadowed[a.x] = shadowed[a.y];
	}
}

// ---- launcher ----------------------------------------------------------------------------------------------

template <int MODE, int WALK> struct KernelOf {
	static constexpr auto value = trace_kernel<MODE, WALK>;
};
template <int MODE> struct KernelOf<MODE, kWalkWide> {
	static constexpr auto value = trace_wide_kernel<MODE>;
};

template <int MODE, int WALK> static cudaError_t launch_mode(const TraceParams &tp, int smCount, cudaStream_t s) {
	static int blocksPerSm = 0; // same for every device of one box
	if (blocksPerSm == 0) {
		cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSm, KernelOf<MODE, WALK>::value, kTraceThreads, 0);
		if (e != cudaSuccess) {
			return e;
		}
		if (blocksPerSm < 1) blocksPerSm = 1;
	}
	// items are numbered with 32 bits inside the kernel; restir_capi.cu splits longer segment lists
	constexpr int CHUNK = ChunkOf<MODE>::value;
	unsigned long long chunks = (tp.nItems + CHUNK - 1) / CHUNK;
	unsigned long long wanted = (chunks + kTraceWarps - 1) / kTraceWarps;
	unsigned grid = (unsigned)std::min<unsigned long long>((unsigned long long)smCount * blocksPerSm, std::max<unsigned long long>(wanted, 1));
	TraceParams launch = tp;
	launch.blocksPerSm = (unsigned)blocksPerSm;
	launch.guidedShare = (float)RESTIR_TRACE_GUIDED / (float)((unsigned long long)grid * kTraceWarps);
#if RESTIR_TRACE_AFFINE
	if (launch.regionCursors == nullptr || launch.smSlots == nullptr || chunks >= 0xffffffffull) {
		return cudaErrorInvalidValue;
	}
	cudaError_t ez = cudaMemsetAsync(launch.regionCursors, 0, sizeof(unsigned) * (kTraceMaxRegions + kTraceMaxSms), s); // cursors, then the per-SM arrival counters
	if (ez != cudaSuccess) {
		return ez;
	}
#endif
	KernelOf<MODE, WALK>::value<<<grid, kTraceThreads, 0, s>>>(launch);
	return cudaGetLastError();
}

static cudaError_t launch_walk(const TraceParams &tp, int mode, int smCount, cudaStream_t s);

cudaError_t launch_trace(const TraceParams &tp, int mode, int smCount, cudaStream_t s) {
	if (tp.nItems == 0) {
		return cudaSuccess;
	}
	cudaError_t e = cudaMemsetAsync(tp.counters + kCounterWork, 0, sizeof(unsigned long long), s);
	if (e != cudaSuccess) {
		return e;
	}
	const bool dedupe = mode == kTraceUnbiased && tp.dedupe != nullptr;
	if (dedupe) {
		e = cudaMemsetAsync(tp.dedupe, 0xff, ((size_t)tp.dedupeMask + 1) * sizeof(unsigned), s);
		if (e == cudaSuccess) e = cudaMemsetAsync(tp.aliasCount, 0, sizeof(unsigned), s);
		if (e != cudaSuccess) {
			return e;
		}
	}
	TraceParams launch = tp;
	launch.divW = fast_div_make((unsigned)std::max(tp.band.W, 1));
	launch.divTilesX = fast_div_make(std::max(tp.tilesX, 1u));
	launch.divSlots = fast_div_make(std::max(tp.slots, 1u));
	e = launch_walk(launch, mode, smCount, s);
	if (e == cudaSuccess && dedupe) {
		alias_resolve_kernel<<<smCount * 4, 256, 0, s>>>(tp.aliases, tp.aliasCount, tp.aliasCapacity, tp.shadowed);
		e = cudaGetLastError();
	}
	return e;
}

static cudaError_t launch_walk(const TraceParams &tp, int mode, int smCount, cudaStream_t s) {
	const int walk = tp.image == nullptr ? kWalkReference : (tp.wide != nullptr && RESTIR_TRACE_TRI_EDGES && !RESTIR_TRACE_REFILL) ? kWalkWide : kWalkImage;
	switch (mode * 3 + walk) {
	case kTracePixel * 3 + kWalkReference: return launch_mode<kTracePixel, kWalkReference>(tp, smCount, s);
	case kTracePixel * 3 + kWalkImage: return launch_mode<kTracePixel, kWalkImage>(tp, smCount, s);
	case kTracePixel * 3 + kWalkWide: return launch_mode<kTracePixel, kWalkWide>(tp, smCount, s);
	case kTraceUnbiased * 3 + kWalkReference: return launch_mode<kTraceUnbiased, kWalkReference>(tp, smCount, s);
	case kTraceUnbiased * 3 + kWalkImage: return launch_mode<kTraceUnbiased, kWalkImage>(tp, smCount, s);
	case kTraceUnbiased * 3 + kWalkWide: return launch_mode<kTraceUnbiased, kWalkWide>(tp, smCount, s);
	case kTraceSegments * 3 + kWalkReference: return launch_mode<kTraceSegments, kWalkReference>(tp, smCount, s);
	case kTraceSegments * 3 + kWalkImage: return launch_mode<kTraceSegments, kWalkImage>(tp, smCount, s);
	default: return launch_mode<kTraceSegments, kWalkWide>(tp, smCount, s);
	}
}

cudaError_t preload_trace_kernels() {
	cudaFuncAttributes a;
	cudaError_t e = cudaFuncGetAttributes(&a, trace_kernel<kTracePixel, kWalkImage>);
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, trace_kernel<kTracePixel, kWalkReference>);
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, trace_wide_kernel<kTracePixel>);
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, trace_kernel<kTraceUnbiased, kWalkImage>);
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, trace_kernel<kTraceUnbiased, kWalkReference>);
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, trace_wide_kernel<kTraceUnbiased>);
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, trace_kernel<kTraceSegments, kWalkImage>);
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, trace_kernel<kTraceSegments, kWalkReference>);
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, trace_wide_kernel<kTraceSegments>);
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, alias_resolve_kernel);
	return e;
}

} // namespace restir
