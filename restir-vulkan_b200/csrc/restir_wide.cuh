// restir_wide.cuh — the shadow-ray walk over the 4-wide quantised image of the uploaded tree (sm_100a).
//
// wide_image.h proves why the answer is the reference's: for rays in the finite range the reference tests a triangle iff its
// own leaf box passes rayAabIntersection, so     shadowed <=> exists t: leafBox(t) passes && triangle(t) hit     and the
// hierarchy above the leaves only has to be conservative.  Here the hierarchy is 4-wide with boxes on a 15-bit grid; the two
// exact factors are evaluated at the leaves with the reference's arithmetic (ray_triangle_edges, ray_box_reference).
//
// One node visit = two 32-byte loads (four boxes + four children; the binary image: the same 64 bytes for two boxes), per
// plane one byte permute (decode AND entry / exit selection) and half a packed fused multiply-add, per box two 3-input
// min / max and three compares.  Measured against the binary walk in profiles/r2_j_summary.md.
#pragma once

#include "restir_trace.cuh"
#include "wide_image.h"

#ifndef RESTIR_WIDE_PREFETCH
#define RESTIR_WIDE_PREFETCH 0 // experiment: 1 = prefetch the line of the first two children of every visited node into L1, 2 = both lines
#endif

// RESTIR_WIDE_SAT 1: the [0, 1] clamp of the box test is done by saturating the x axis' multiply-adds (FMA pipe) instead of two
// min / max per box (ALU pipe); 0: the clamps as separate operations.  Every box the reference's test passes is hit either way.
#ifndef RESTIR_WIDE_SAT
#define RESTIR_WIDE_SAT 1
#endif
#if RESTIR_WIDE_SAT
#define WIDE_CLAMP_LO(x) (x)
#define WIDE_CLAMP_HI(x) (x)
#else
#define WIDE_CLAMP_LO(x) fmaxf((x), 0.0f)
#define WIDE_CLAMP_HI(x) fminf((x), 1.0f)
#endif

// RESTIR_WIDE_FLAT_LOOP 1 (experiment): one loop branch per visit, the next group settled by predicated moves at the end of a
// visit.  Measured (profiles/r2_p_flat.log): neighbours 0.879 -> 0.887 ms — the predicates cost four more ALU-pipe operations than
// the convergence barriers and branches they replace, and the ALU pipe is what the walk is short of.
#ifndef RESTIR_WIDE_FLAT_LOOP
#define RESTIR_WIDE_FLAT_LOOP 0
#endif
// Also measured and removed (profiles/r2_p_imad.log): the four hit bits gathered on the FMA pipe (IMAD.HI by 2 = the sign bit, three
// IMADs to combine) instead of four funnel shifts on the ALU pipe: neighbours 0.878 -> 0.900 ms (IMAD.HI is not a one-pass operation).

namespace restir {

struct U8 {
	unsigned v[8];
};
__device__ __forceinline__ U8 ldg256u(const void *p) {
	U8 r;
#ifdef WIDE_NO_LDG256
	uint4 x = __ldg(reinterpret_cast<const uint4 *>(p)), y = __ldg(reinterpret_cast<const uint4 *>(p) + 1);
	r.v[0] = x.x; r.v[1] = x.y; r.v[2] = x.z; r.v[3] = x.w; r.v[4] = y.x; r.v[5] = y.y; r.v[6] = y.z; r.v[7] = y.w;
	return r;
#endif
	asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	    : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
	    : "l"(p));
	return r;
}

__device__ __forceinline__ float fmax3(float a, float b, float c) {
#ifdef WIDE_NO_MNMX3
	return fmaxf(a, fmaxf(b, c));
#else
	float r;
	asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); // FMNMX3
	return r;
#endif
}
__device__ __forceinline__ float fmin3(float a, float b, float c) {
#ifdef WIDE_NO_MNMX3
	return fminf(a, fminf(b, c));
#else
	float r;
	asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
	return r;
#endif
}

// Per-ray constants of the conservative box test (wide_image.h wide_ray_setup, same operations), duplicated into both halves
// of a register pair where a packed FFMA2 reads them.
struct WideLaneRay {
	float2 sx, sy, sz, cLoX, cLoY, cLoZ, cHiX, cHiY, cHiZ;
	unsigned selNx, selNy, selNz, selFx, selFy, selFz;
};

// The walk keeps its loop-invariant operands in registers.  Without this ptxas, held to 64 registers, prefers to RECOMPUTE them
// from the origin and the direction on every node visit (K * inv, (h - o) * inv, the margins, the selectors: twenty-odd
// instructions per visit, profiles/r2_j_summary.md); an empty asm makes the values opaque, so they are kept.
__device__ __forceinline__ void keep(float &x) { asm volatile("" : "+f"(x)); }
__device__ __forceinline__ void keep(unsigned &x) { asm volatile("" : "+r"(x)); }

__device__ __forceinline__ bool wide_lane_setup(const WideGrid &g, f3 o, f3 d, WideLaneRay &r) {
	const float of[3] = {o.x, o.y, o.z}, df[3] = {d.x, d.y, d.z};
	const float iv[3] = {1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
	WideRay w;
	bool ok = wide_ray_setup(g, of, df, iv, w);
	for (int a = 0; a < 3; ++a) {
		keep(w.s[a]);
		keep(w.cLo[a]);
		keep(w.cHi[a]);
		keep(w.selNear[a]);
		keep(w.selFar[a]);
	}
	r.sx = make_float2(w.s[0], w.s[0]); r.sy = make_float2(w.s[1], w.s[1]); r.sz = make_float2(w.s[2], w.s[2]);
	r.cLoX = make_float2(w.cLo[0], w.cLo[0]); r.cLoY = make_float2(w.cLo[1], w.cLo[1]); r.cLoZ = make_float2(w.cLo[2], w.cLo[2]);
	r.cHiX = make_float2(w.cHi[0], w.cHi[0]); r.cHiY = make_float2(w.cHi[1], w.cHi[1]); r.cHiZ = make_float2(w.cHi[2], w.cHi[2]);
	r.selNx = w.selNear[0]; r.selNy = w.selNear[1]; r.selNz = w.selNear[2];
	r.selFx = w.selFar[0]; r.selFy = w.selFar[1]; r.selFz = w.selFar[2];
	return ok;
}

// entry / exit parameters of two children at once on one axis
__device__ __forceinline__ void wide_axis_pair(unsigned w0, unsigned w1, unsigned selN, unsigned selF, float2 s, float2 cLo, float2 cHi, float2 &tn, float2 &tf) {
	float2 vn = make_float2(__uint_as_float(__byte_perm(w0, 0x3F000000u, selN)), __uint_as_float(__byte_perm(w1, 0x3F000000u, selN)));
	float2 vf = make_float2(__uint_as_float(__byte_perm(w0, 0x3F000000u, selF)), __uint_as_float(__byte_perm(w1, 0x3F000000u, selF)));
#ifdef WIDE_NO_FFMA2
	tn = make_float2(fmaf(vn.x, s.x, cLo.x), fmaf(vn.y, s.y, cLo.y));
	tf = make_float2(fmaf(vf.x, s.x, cHi.x), fmaf(vf.y, s.y, cHi.y));
#else
	tn = __ffma2_rn(vn, s, cLo);
	tf = __ffma2_rn(vf, s, cHi);
#endif
}

// The same on the axis whose parameters are SATURATED to [0, 1] (wide_image.h wide_box_hit: the clamp of the segment's range,
// max(entry, 0) <= min(exit, 1), rides on one axis' multiply-adds instead of two min / max per box on the ALU pipe, the busiest
// unit of this kernel at 61-66 %, ncu capture N; fma.sat has no packed form, so two scalar operations replace each FFMA2).
__device__ __forceinline__ float fma_sat(float a, float b, float c) {
	float r;
	asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
	return r;
}
__device__ __forceinline__ void wide_axis_pair_sat(unsigned w0, unsigned w1, unsigned selN, unsigned selF, float2 s, float2 cLo, float2 cHi, float2 &tn, float2 &tf) {
	tn = make_float2(fma_sat(__uint_as_float(__byte_perm(w0, 0x3F000000u, selN)), s.x, cLo.x), fma_sat(__uint_as_float(__byte_perm(w1, 0x3F000000u, selN)), s.y, cLo.y));
	tf = make_float2(fma_sat(__uint_as_float(__byte_perm(w0, 0x3F000000u, selF)), s.x, cHi.x), fma_sat(__uint_as_float(__byte_perm(w1, 0x3F000000u, selF)), s.y, cHi.y));
}

// (*) of wide_image.h at a leaf: the reference's triangle test, then the reference's slab test on the leaf's own fp32 box
// (kept in the spare floats of the 64-byte triangle record: e2.z, min.xyz | max.xyz, -)
__device__ __forceinline__ bool wide_leaf_hit(const float4 *__restrict__ triRec, unsigned id, f3 o, f3 d) {
	if (!ray_triangle_edges(triRec, (int)id, o, d)) {
		return false;
	}
	const float4 *t = triRec + (size_t)id * 4;
	float4 a = __ldg(t + 2), b = __ldg(t + 3);
	keep(d.x), keep(d.y), keep(d.z); // the three divisions belong to the (rare) hit, not in front of the loop over the leaves where the compiler hoists them
	f3 inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
	return ray_box_reference(o, inv, make_float4(a.y, a.z, a.w, 0.0f), make_float4(b.x, b.y, b.z, 0.0f));
}

// (*) of wide_image.h holds for this ray whatever finds the leaf: origin and direction finite, 1 / direction finite (|d| in
// [2^-39, 2^39] puts |1 / d| inside wide_ray_setup's [2^-40, 2^40] without a division), origin within the grid's range.
__device__ __forceinline__ bool wide_ray_in_range(const WideGrid &g, f3 o, f3 d) {
	const float lo = 1.8189894035458565e-12f, hi = 549755813888.0f;
	return fabsf(d.x) >= lo && fabsf(d.x) <= hi && fabsf(d.y) >= lo && fabsf(d.y) <= hi && fabsf(d.z) >= lo && fabsf(d.z) <= hi &&
	       fabsf(o.x) <= g.maxOrigin[0] && fabsf(o.y) <= g.maxOrigin[1] && fabsf(o.z) <= g.maxOrigin[2];
}

// Returns the record of the triangle that occludes the segment, -1 when nothing is hit.  The caller has checked wide_lane_setup.
//
// The walk's unit is a GROUP: (first child node << 4) | 4-bit mask of the children still to visit.  The inner children a visit
// finds hit become the current group in one step — no index word per child (wide_image.h: children of a node are numbered
// consecutively), no push per child — and the group left over from the level above goes on the stack as ONE entry, so the
// stack holds one entry per level of the wide tree (<= kWideStack, checked at upload).
// od: this thread's segment origin and direction in shared memory (od[c * stride], c = 0..5), read back only where a leaf is
// tested (1.7 times per ray): six registers less across the walk.
__device__ __forceinline__ int trace_any_wide(const uint4 *__restrict__ wide, const float4 *__restrict__ triRec, const WideLaneRay &r, const float *od,
                                              int stride) {
#if RESTIR_WIDE_FLAT_LOOP
	// One loop branch per visit: the group to visit next is settled at the END of a visit with predicated moves (push the siblings
	// left over, take the hit inner children, or pop), and an empty group at the bottom of the stack ends the walk — no nested
	// `if (empty) { if (top == 0) return; pop }` with its convergence barrier and two branches at the head of every visit.
	unsigned stack[kWideStack + 1];
	stack[0] = 0u; // the sentinel: popping it leaves an empty group
	int top = 1;
	int found = -1;
	unsigned group = 1u; // node 0, one child to visit: the root
	do {
		const unsigned slot = (unsigned)__ffs((int)(group & 15u)) - 1u;
#else
	unsigned stack[kWideStack];
	int top = 0;
	unsigned group = 1u; // node 0, one child to visit: the root
	for (;;) {
		if ((group & 15u) == 0u) {
			if (top == 0) {
				return -1;
			}
			group = stack[--top];
		}
		const unsigned slot = (unsigned)__ffs((int)(group & 15u)) - 1u;
#endif
		group &= group - 1u; // the lowest set bit lies in the mask
		const uint4 *n = wide + ((size_t)((group >> 4) + slot)) * 4u;
		const U8 a = ldg256u(n), b = ldg256u(n + 2); // a: x words of the four slots, y words; b: z words, (childGroup, recBase, innerMask, count)
#if RESTIR_WIDE_PREFETCH > 0
		// the children of a node are consecutive 64-byte nodes (two per 128-byte line): ask for them while the boxes are tested
		if (b.v[6] != 0u) {
			const uint4 *c = wide + (size_t)(b.v[4] >> 4) * 4u;
			asm volatile("prefetch.global.L1 [%0];" ::"l"(c));
#if RESTIR_WIDE_PREFETCH > 1
			if (b.v[6] > 3u) {
				asm volatile("prefetch.global.L1 [%0];" ::"l"(c + 8));
			}
#endif
		}
#endif
		float2 nx, fx, ny, fy, nz, fz;
		unsigned hits;
		// slots 0 and 1
#if RESTIR_WIDE_SAT
		wide_axis_pair_sat(a.v[0], a.v[1], r.selNx, r.selFx, r.sx, r.cLoX, r.cHiX, nx, fx);
#else
		wide_axis_pair(a.v[0], a.v[1], r.selNx, r.selFx, r.sx, r.cLoX, r.cHiX, nx, fx);
#endif
		wide_axis_pair(a.v[4], a.v[5], r.selNy, r.selFy, r.sy, r.cLoY, r.cHiY, ny, fy);
		wide_axis_pair(b.v[0], b.v[1], r.selNz, r.selFz, r.sz, r.cLoZ, r.cHiZ, nz, fz);
		// a slot is hit when max(entry, 0) < min(exit, 1) — strictly: the margins leave every box the reference passes a positive
		// gap (wide_image.h wide_box_hit) — so the sign bit of entry - exit says "hit".  With RESTIR_WIDE_SAT the x parameters arrive
		// saturated to [0, 1], which is that clamp; a box beyond either end of the segment on x alone collapses onto 1 - 1 or 0 - 0,
		// an exact +0: missed, as it must be.
		float2 gap01;
		{
			float n0 = WIDE_CLAMP_LO(fmax3(nx.x, ny.x, nz.x)), f0 = WIDE_CLAMP_HI(fmin3(fx.x, fy.x, fz.x));
			float n1 = WIDE_CLAMP_LO(fmax3(nx.y, ny.y, nz.y)), f1 = WIDE_CLAMP_HI(fmin3(fx.y, fy.y, fz.y));
			gap01 = __fadd2_rn(make_float2(n0, n1), make_float2(-f0, -f1));
		}
		// slots 2 and 3
#if RESTIR_WIDE_SAT
		wide_axis_pair_sat(a.v[2], a.v[3], r.selNx, r.selFx, r.sx, r.cLoX, r.cHiX, nx, fx);
#else
		wide_axis_pair(a.v[2], a.v[3], r.selNx, r.selFx, r.sx, r.cLoX, r.cHiX, nx, fx);
#endif
		wide_axis_pair(a.v[6], a.v[7], r.selNy, r.selFy, r.sy, r.cLoY, r.cHiY, ny, fy);
		wide_axis_pair(b.v[2], b.v[3], r.selNz, r.selFz, r.sz, r.cLoZ, r.cHiZ, nz, fz);
		{
			float n0 = WIDE_CLAMP_LO(fmax3(nx.x, ny.x, nz.x)), f0 = WIDE_CLAMP_HI(fmin3(fx.x, fy.x, fz.x));
			float n1 = WIDE_CLAMP_LO(fmax3(nx.y, ny.y, nz.y)), f1 = WIDE_CLAMP_HI(fmin3(fx.y, fy.y, fz.y));
			const float2 gap23 = __fadd2_rn(make_float2(n0, n1), make_float2(-f0, -f1));
			// the four sign bits, slot 0 in bit 0: each funnel shift appends one
			hits = __float_as_uint(gap23.y) >> 31;
			hits = __funnelshift_l(__float_as_uint(gap23.x), hits, 1);
			hits = __funnelshift_l(__float_as_uint(gap01.y), hits, 1);
			hits = __funnelshift_l(__float_as_uint(gap01.x), hits, 1);
		}
		// hit leaves: the slots outside innerMask, the record of slot j is recBase + j; empty slots are never hit (inverted boxes)
		unsigned leaf = hits & ~b.v[6];
		if (leaf != 0u) {
			const f3 o = mk3(od[0], od[stride], od[2 * stride]), d = mk3(od[3 * stride], od[4 * stride], od[5 * stride]);
			const unsigned triBase = b.v[5];
#pragma unroll 1
			do {
				const unsigned j = (unsigned)__ffs((int)leaf) - 1u;
				leaf &= leaf - 1u;
				if (wide_leaf_hit(triRec, triBase + j, o, d)) {
#if RESTIR_WIDE_FLAT_LOOP
					found = (int)(triBase + j);
					break;
#else
					return (int)(triBase + j);
#endif
				}
			} while (leaf != 0u);
#if RESTIR_WIDE_FLAT_LOOP
			if (found >= 0) {
				break;
			}
#endif
		}
		// hit inner children: the slots of innerMask, nodes firstChild + slot
		const unsigned inner = hits & b.v[6];
#if RESTIR_WIDE_FLAT_LOOP
		const bool has = inner != 0u, siblings = (group & 15u) != 0u;
		if (has && siblings) {
			stack[top] = group;
		}
		top += has && siblings ? 1 : 0;
		if (!has && !siblings) {
			group = stack[top - 1];
		}
		top -= !has && !siblings ? 1 : 0;
		group = has ? (b.v[4] | inner) : group;
	} while ((group & 15u) != 0u);
	return found;
#else
		if (inner != 0u) {
			if ((group & 15u) != 0u) {
				stack[top++] = group;
			}
			group = b.v[4] | inner;
		}
	}
#endif
}

} // namespace restir
