// restir_trace.cuh — shadow-ray device functions (sm_100a).
//
//   segment_setup / test_visibility   <- src/shaders/include/visibilityTest.glsl:1-4, 27-28 (software branch)
//   ray_box_reference, ray_triangle   <- src/shaders/include/softwareRaytracing.glsl:9-14, 15-37
//   trace_any_reference               <- softwareRaytracing.glsl:39-85, in the reference's node order
//   wide_step                         one visit of the 4-wide re-layout (wide_bvh.h) — same answer, see there
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
// when nothing is hit.  This is the path for rays whose 1/dir is not finite and for trees the 4-wide
// re-layout does not cover.
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
// 4-wide traversal state of one lane.
constexpr int kWideStack = 32;

struct WideRay {
	f3 o, d, inv;
	int nearX, nearY, nearZ; // float4 index of the near plane set per axis inside a WideNode (axis + 3 * (inv < 0))
};

__device__ __forceinline__ void wide_ray_init(WideRay &r, f3 o, f3 d, f3 inv) {
	r.o = o;
	r.d = d;
	r.inv = inv;
	r.nearX = inv.x < 0.0f ? 3 : 0;
	r.nearY = inv.y < 0.0f ? 4 : 1;
	r.nearZ = inv.z < 0.0f ? 5 : 2;
}

// With finite non-zero 1/dir and lo <= hi the reference's min(t1,t2) / max(t1,t2) are the near / far plane
// distances selected by the sign of 1/dir — bit for bit, since (b - o) * inv is monotone in b.
__device__ __forceinline__ bool wide_slab(const WideRay &r, float nx, float ny, float nz, float fx, float fy, float fz) {
	float tn = fmaxf((nx - r.o.x) * r.inv.x, fmaxf((ny - r.o.y) * r.inv.y, (nz - r.o.z) * r.inv.z));
	float tf = fminf((fx - r.o.x) * r.inv.x, fminf((fy - r.o.y) * r.inv.y, (fz - r.o.z) * r.inv.z));
	return tn < 1.0f && tf >= tn && tf > 0.0f;
}

enum WideStepResult { kWideContinue = 0, kWideMiss = 1, kWideHit = 2, kWideStackFull = 3 };

// Visits wide node `cur`: tests its four boxes, tests the triangles of hit leaf slots at once, makes the
// first hit inner slot the next node and pushes the others.
__device__ __forceinline__ int wide_step(const float4 *__restrict__ wide, const float4 *__restrict__ tris, const WideRay &r, int &cur, int *stack,
                                         int &top) {
	const float4 *n = wide + (size_t)cur * 8;
	float4 nx = __ldg(n + r.nearX), ny = __ldg(n + r.nearY), nz = __ldg(n + r.nearZ);
	float4 fx = __ldg(n + (3 - r.nearX)), fy = __ldg(n + (5 - r.nearY)), fz = __ldg(n + (7 - r.nearZ));
	int4 ch = __ldg(reinterpret_cast<const int4 *>(n + 6));
	unsigned hit = 0;
	hit |= wide_slab(r, nx.x, ny.x, nz.x, fx.x, fy.x, fz.x) ? 1u : 0u;
	hit |= wide_slab(r, nx.y, ny.y, nz.y, fx.y, fy.y, fz.y) ? 2u : 0u;
	hit |= wide_slab(r, nx.z, ny.z, nz.z, fx.z, fy.z, fz.z) ? 4u : 0u;
	hit |= wide_slab(r, nx.w, ny.w, nz.w, fx.w, fy.w, fz.w) ? 8u : 0u;
	unsigned leaf = (ch.x < 0 ? 1u : 0u) | (ch.y < 0 ? 2u : 0u) | (ch.z < 0 ? 4u : 0u) | (ch.w < 0 ? 8u : 0u);
	unsigned lh = hit & leaf;
	while (lh) {
		int c = __ffs(lh) - 1;
		lh &= lh - 1;
		int id = c == 0 ? ch.x : (c == 1 ? ch.y : (c == 2 ? ch.z : ch.w));
		if (ray_triangle(tris, ~id, r.o, r.d)) {
			return kWideHit;
		}
	}
	unsigned ih = hit & ~leaf;
	if (ih == 0) {
		if (top == 0) {
			return kWideMiss;
		}
		cur = stack[--top];
		return kWideContinue;
	}
	if (top + 3 > kWideStack) {
		return kWideStackFull;
	}
	bool first = true;
	if (ih & 1u) { cur = ch.x; first = false; }
	if (ih & 2u) { if (first) { cur = ch.y; first = false; } else { stack[top++] = ch.y; } }
	if (ih & 4u) { if (first) { cur = ch.z; first = false; } else { stack[top++] = ch.z; } }
	if (ih & 8u) { if (first) { cur = ch.w; } else { stack[top++] = ch.w; } }
	return kWideContinue;
}

} // namespace restir
