// restir_trace.cuh — shadow-ray device functions (sm_100a).
//
//   segment_setup / test_visibility   <- src/shaders/include/visibilityTest.glsl:1-4, 27-28 (software branch)
//   ray_box_reference, ray_triangle   <- src/shaders/include/softwareRaytracing.glsl:9-14, 15-37
//   trace_any_reference               <- softwareRaytracing.glsl:39-85, in the reference's node order
//   trace_any_image                   the same walk over the 64-byte image of the same tree (traversal_image.h)
#pragma once

#include "restir_device.cuh"

namespace restir {

// softwareRaytracing.glsl:9-14 with the division hoisted (P3): inv = 1/dir once per ray.
__device__ __forceinline__ bool ray_box_reference(f3 o, f3 inv, float4 bmin, float4 bmax) {
	float t1x = (bmin.x - o.x) * inv.x, t1y = (bmin.y - o.y) * inv.y, t1z = (bmin.z - o.z) * inv.z;
	float t2x = (bmax.x - o.x) * inv.x, t2y = (bmax.y - o.y) * inv.y, t2z = (bmax.z - o.z) * inv.z;
	float rmin = fmaxf(fminf(t1x, t2x), fmaxf(fminf(t1y, t2y), fminf(t1z, t2z)));
	float rmax = fminf(fmaxf(t1x, t2x), fminf(fmaxf(t1y, t2y), fmaxf(t1z, t2z)));
	return rmin < 1.0f && rmax >= rmin && rmax > 0.0f;
}

// One 32-byte read-only load (sm_100 LDG.E.256): a 64-byte node is two of them instead of four 16-byte loads —
// half the requests on the L1 tag stage this kernel is bound by.  p must be 32-byte aligned.
struct F8 {
	float v[8];
};
__device__ __forceinline__ F8 ldg256(const void *p) {
	F8 r;
	asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	    : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
	    : "l"(p));
	return r;
}

// softwareRaytracing.glsl:15-37
__device__ __forceinline__ bool ray_triangle(const float4 *__restrict__ tris, int id, f3 o, f3 d) {
	const float4 *t = tris + (size_t)id * 3;
	float4 a = __ldg(t), b = __ldg(t + 1), c = __ldg(t + 2);
	f3 p1 = mk3(a.x, a.y, a.z);
	f3 e1 = mk3(b.x, b.y, b.z) - p1;
	f3 e2 = mk3(c.x, c.y, c.z) - p1;
	f3 p = cross3(d, e2);
	float f = 1.0f / dot3(e1, p);
	f3 s = o - p1;
	float baryX = f * dot3(s, p);
	if (baryX < 0.0f || baryX > 1.0f) {
		return false;
	}
	f3 q = cross3(s, e1);
	float baryY = f * dot3(d, q);
	if (baryY < 0.0f || baryY + baryX > 1.0f) {
		return false;
	}
	f = f * dot3(e2, q);
	return f > 0.0f && f < 1.0f;
}

// softwareRaytracing.glsl:39-85 in the reference's own order on the reference's own nodes.  Any-hit: the
// answer does not depend on the order in which nodes and triangles are visited, only on which boxes /
// triangles the segment intersects, so triangles are tested as soon as their leaf box is hit instead of
// being deferred in batches of 8 node visits.  The stack is the reference's 32 entries with its push order
// (left, then right); a push onto a full stack is dropped and counted (UB in the reference).  Returns true
// when nothing is hit.  This literal path serves trees whose worst-case stack occupancy exceeds 32
// (traversal_image.h) and RESTIR_TRAVERSAL_REFERENCE_ORDER.
static __device__ __noinline__ bool trace_any_reference(const float4 *__restrict__ nodes, const float4 *__restrict__ tris, f3 o, f3 d, unsigned &overflow) {
	int stack[32];
	int top = 1;
	stack[0] = 0;
	f3 inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
	while (top > 0) {
		const float4 *n = nodes + (size_t)stack[--top] * 5;
		float4 lmin = __ldg(n), lmax = __ldg(n + 1), rmin = __ldg(n + 2), rmax = __ldg(n + 3);
		float4 ch = __ldg(n + 4);
		int left = __float_as_int(ch.x), right = __float_as_int(ch.y);
		if (ray_box_reference(o, inv, lmin, lmax)) {
			if (left < 0) {
				if (ray_triangle(tris, ~left, o, d)) {
					return false;
				}
			} else if (top < 32) {
				stack[top++] = left;
			} else {
				overflow++;
			}
		}
		if (ray_box_reference(o, inv, rmin, rmax)) {
			if (right < 0) {
				if (ray_triangle(tris, ~right, o, d)) {
					return false;
				}
			} else if (top < 32) {
				stack[top++] = right;
			} else {
				overflow++;
			}
		}
	}
	return true;
}

// visibilityTest.glsl:1-4, 27-28: the traced segment starts tMin = 0.001 world units after p1 and ends
// 0.001 before p2.
__device__ __forceinline__ void segment_setup(f3 p1, f3 p2, f3 &o, f3 &d) {
	f3 dir = p2 - p1;
	f3 offset = normalize3(dir) * 0.001f;
	o = p1 + offset;
	d = dir - offset * 2.0f;
}

// Returns SHADOWED (visibilityTest.glsl:27-28), reference-order traversal.
__device__ __forceinline__ bool test_visibility_reference(const SceneView &sc, f3 p1, f3 p2, unsigned &overflow) {
	f3 o, d;
	segment_setup(p1, p2, o, d);
	return !trace_any_reference(sc.nodes, sc.tris, o, d, overflow);
}

// ------------------------------------------------------------------------------------------------
// The walk over the 64-byte image of the same tree (traversal_image.h): two 32-byte loads per node from one
// 128-byte line.  Only used when the tree is no deeper than the 32-entry stack (checked at upload), so no push
// can be dropped and the stack needs no bound checks; the node about to be visited stays in a register instead
// of going through the stack.
//
// What is fixed by the reference (softwareRaytracing.glsl:39-85) is WHICH boxes and triangles a segment is tested
// against — a triangle is tested iff every box on the path from the root to its leaf is hit, with the reference's
// slab and Moller-Trumbore arithmetic — not the ORDER: the answer is any-hit.  Two exact choices, both measured
// (profiles/r2_a_trace_ab.md, B200, Sponza 1080p unbiased 5):
//   RESTIR_TRACE_NEAR_FIRST 0 (default): of two hit inner children the right one is visited first, the reference's
//     pop order.  1: the one the segment enters first (smaller slab entry parameter, already computed).  Near-first
//     LOSES here (3.43 -> 3.49 ms/frame): a batch of 32 rays ends with its slowest ray, and the unshadowed rays of a
//     batch visit the same nodes in any order, while lanes that pick their own order stop sharing node loads (the
//     kernel is bound by L1 tag lookups).
//   RESTIR_TRACE_TRI_EDGES 1 (default): triangles are read as (p1, e1 = p2 - p1, e2 = p3 - p1) from 64-byte records
//     derived at upload with the very subtractions softwareRaytracing.glsl:16-17 makes per test (same operands, same
//     IEEE operation => same bits): one 32-byte + one 4-byte load from one line and six subtractions less per test
//     (restirOmni's rays, where triangle tests are 28 % of the instructions: 0.630 -> 0.615 ms).
#ifndef RESTIR_TRACE_NEAR_FIRST
#define RESTIR_TRACE_NEAR_FIRST 0
#endif
#ifndef RESTIR_TRACE_TRI_EDGES
#define RESTIR_TRACE_TRI_EDGES 1
#endif
//   RESTIR_TRACE_TOP_SMEM 0 (default): every node is read from global memory (L1 / L2).  N > 0 (experiment): the trace kernel's
//     CTAs keep the first N nodes of the image — the top levels of the tree: AabbTree::build numbers nodes breadth-first — in
//     shared memory and visits to them read it (four 16-byte LDS) instead of two 32-byte LDG.  Measured: profiles/r2_i_summary.md.
#ifndef RESTIR_TRACE_TOP_SMEM
#define RESTIR_TRACE_TOP_SMEM 0
#endif

// softwareRaytracing.glsl:15-37 on a derived (p1, e1, e2) record
__device__ __forceinline__ bool ray_triangle_edges(const float4 *__restrict__ triEdges, int id, f3 o, f3 d) {
	const float4 *t = triEdges + (size_t)id * 4;
	F8 a = ldg256(t);
	float e2z = __ldg(reinterpret_cast<const float *>(t + 2));
	f3 p1 = mk3(a.v[0], a.v[1], a.v[2]);
	f3 e1 = mk3(a.v[3], a.v[4], a.v[5]);
	f3 e2 = mk3(a.v[6], a.v[7], e2z);
	f3 p = cross3(d, e2);
	float f = 1.0f / dot3(e1, p);
	f3 s = o - p1;
	float baryX = f * dot3(s, p);
	if (baryX < 0.0f || baryX > 1.0f) {
		return false;
	}
	f3 q = cross3(s, e1);
	float baryY = f * dot3(d, q);
	if (baryY < 0.0f || baryY + baryX > 1.0f) {
		return false;
	}
	f = f * dot3(e2, q);
	return f > 0.0f && f < 1.0f;
}

// Per-ray constants of the packed slab test: (b - o) * inv as FADD2 (b + (-o), the same IEEE operation) and FMUL2,
// the left box in the low half and the right box in the high half of every pair.
struct WalkRay {
	f3 o, d;
	float2 nox, noy, noz, ivx, ivy, ivz;
};
__device__ __forceinline__ WalkRay walk_ray(f3 o, f3 d) {
	WalkRay r;
	r.o = o;
	r.d = d;
	f3 inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
	r.nox = make_float2(-o.x, -o.x); r.noy = make_float2(-o.y, -o.y); r.noz = make_float2(-o.z, -o.z);
	r.ivx = make_float2(inv.x, inv.x); r.ivy = make_float2(inv.y, inv.y); r.ivz = make_float2(inv.z, inv.z);
	return r;
}

enum WalkStatus { kWalkOn = 0, kWalkClear = 1, kWalkHit = 2 };

// One node visit: both boxes, hit leaves, then the next node (or the end of the walk).
__device__ __forceinline__ int walk_step(const float4 *__restrict__ image, const float4 *__restrict__ tris, const WalkRay &r, int &cur, int &top, int *stack,
                                         const float4 *topNodes = nullptr) {
	const float4 *n = image + (unsigned)cur * 4u;
	F8 lo, hi;
	if (RESTIR_TRACE_TOP_SMEM > 0 && topNodes != nullptr && cur < RESTIR_TRACE_TOP_SMEM) {
		const float4 *s = topNodes + (unsigned)cur * 4u;
		float4 a = s[0], b = s[1], c = s[2], d = s[3];
		lo = F8{{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w}};
		hi = F8{{c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w}};
	} else {
		lo = ldg256(n);
		hi = ldg256(n + 2);
	}
	float4 qx = make_float4(lo.v[0], lo.v[1], lo.v[2], lo.v[3]), qy = make_float4(lo.v[4], lo.v[5], lo.v[6], lo.v[7]);
	float4 qz = make_float4(hi.v[0], hi.v[1], hi.v[2], hi.v[3]);
	int2 ch = make_int2(__float_as_int(hi.v[4]), __float_as_int(hi.v[5]));
	float2 t1x = __fmul2_rn(__fadd2_rn(make_float2(qx.x, qx.y), r.nox), r.ivx), t2x = __fmul2_rn(__fadd2_rn(make_float2(qx.z, qx.w), r.nox), r.ivx);
	float2 t1y = __fmul2_rn(__fadd2_rn(make_float2(qy.x, qy.y), r.noy), r.ivy), t2y = __fmul2_rn(__fadd2_rn(make_float2(qy.z, qy.w), r.noy), r.ivy);
	float2 t1z = __fmul2_rn(__fadd2_rn(make_float2(qz.x, qz.y), r.noz), r.ivz), t2z = __fmul2_rn(__fadd2_rn(make_float2(qz.z, qz.w), r.noz), r.ivz);
	// softwareRaytracing.glsl:11-13 per box
	float lmin = fmaxf(fminf(t1x.x, t2x.x), fmaxf(fminf(t1y.x, t2y.x), fminf(t1z.x, t2z.x)));
	float lmax = fminf(fmaxf(t1x.x, t2x.x), fminf(fmaxf(t1y.x, t2y.x), fmaxf(t1z.x, t2z.x)));
	float rmin = fmaxf(fminf(t1x.y, t2x.y), fmaxf(fminf(t1y.y, t2y.y), fminf(t1z.y, t2z.y)));
	float rmax = fminf(fmaxf(t1x.y, t2x.y), fminf(fmaxf(t1y.y, t2y.y), fmaxf(t1z.y, t2z.y)));
	bool hl = lmin < 1.0f && lmax >= lmin && lmax > 0.0f;
	bool hr = rmin < 1.0f && rmax >= rmin && rmax > 0.0f;
	// hit leaves: left first, then right (most visits have none: one branch skips the whole block)
	int t0 = (hl && ch.x < 0) ? ~ch.x : -1, t1 = (hr && ch.y < 0) ? ~ch.y : -1;
	if ((t0 & t1) >= 0) { // at least one of them is a triangle index
		if (t0 < 0) {
			t0 = t1;
			t1 = -1;
		}
#pragma unroll 1
		do {
			if (RESTIR_TRACE_TRI_EDGES ? ray_triangle_edges(tris, t0, r.o, r.d) : ray_triangle(tris, t0, r.o, r.d)) {
				return kWalkHit;
			}
			t0 = t1;
			t1 = -1;
		} while (t0 >= 0);
	}
	bool il = hl && ch.x >= 0, ir = hr && ch.y >= 0;
	if (il && ir) {
		// both inner children hit: one is visited now, the other pushed
		bool leftFirst = RESTIR_TRACE_NEAR_FIRST ? lmin < rmin : false;
		stack[top++] = leftFirst ? ch.y : ch.x;
		cur = leftFirst ? ch.x : ch.y;
	} else if (il || ir) {
		cur = il ? ch.x : ch.y;
	} else {
		if (top == 0) {
			return kWalkClear;
		}
		cur = stack[--top];
	}
	return kWalkOn;
}

// Returns true when nothing is hit.
__device__ __forceinline__ bool trace_any_image(const float4 *__restrict__ image, const float4 *__restrict__ tris, f3 o, f3 d,
                                                const float4 *topNodes = nullptr) {
	int stack[32];
	int top = 0, cur = 0;
	const WalkRay r = walk_ray(o, d);
	for (;;) {
		int st = walk_step(image, tris, r, cur, top, stack, topNodes);
		if (st != kWalkOn) {
			return st == kWalkClear;
		}
	}
}

} // namespace restir
