// wide_image.h — the 4-wide, 64-byte-per-node image of the uploaded AABB tree that the trace kernel walks (host, once per
// restir_upload_bvh), and the arithmetic of its box test (host + device: the CPU tests run the very same operations).
//
// WHY A DIFFERENT TREE MAY BE WALKED AT ALL.  softwareRaytracing.glsl:39-85 tests a triangle iff every box on the path from the
// root to its leaf passes rayAabIntersection (:9-14) and reports "some tested triangle is hit" (any-hit, order-free).  For a ray
// whose origin, direction and 1/direction are finite (no 0 * inf inside the slab test), the slab test is MONOTONE in the box:
// fl(fl(b - o) * inv) is non-decreasing in b for inv > 0 and non-increasing for inv < 0 (rounding is monotone), so for a box C
// inside a box P every per-axis entry parameter of C is >= P's and every exit parameter <= P's; hence rmin(C) >= rmin(P),
// rmax(C) <= rmax(P), and `rmin < 1 && rmax >= rmin && rmax > 0` for C implies it for P.  AabbTree::build stores, for every
// child, the exact union of its triangles' bounds (min / max are exact), so the boxes are nested — checked at upload, not
// assumed.  Therefore:  a triangle is tested by the reference  <=>  ITS OWN LEAF BOX passes.  And
//
//     shadowed  <=>  exists t:  rayAabIntersection(leaf box of t)  &&  rayTriangleIntersection(t)          (*)
//
// with the reference's exact arithmetic in both factors.  The inner boxes only steer the search; ANY conservative hierarchy
// over the leaves gives the same bits.  This file builds one that is cheap to walk:
//
//   * 4 children per node (the binary tree collapsed by surface area: half the levels, no test of the boxes in between);
//   * child boxes quantised OUTWARDS to a 15-bit grid over the scene, one 32-bit word per (axis, child) = lo | hi << 16:
//     64 bytes per node = two 32-byte loads for four boxes (the binary image: 64 bytes for two);
//   * the box test decodes a plane with ONE byte permute — 0x3F000000 | q << 8 is the float 0.5 + q / 65536 — whose selector
//     also picks lo or hi by the sign of the ray direction (no min / max per axis), and evaluates entry / exit parameters with
//     one fused multiply-add per plane: t = v * (K * inv) + (h - o) * inv, K = 65536 cells, h = grid origin - K / 2;
//   * that arithmetic is NOT the reference's, so it carries a margin: entry parameters are lowered and exit parameters raised
//     by m = 2^-20 * (|o| + H) * |inv| per axis, H = max |grid coordinate| (error budget: wide_ray_setup below), which makes
//     "reference leaf test passes => every wide box above the leaf passes" a theorem, not a hope;
//   * at a leaf the triangle is tested with the reference's arithmetic (restir_trace.cuh ray_triangle_edges) and, on a hit,
//     the leaf's ORIGINAL fp32 box with the reference's slab test — (*) literally.
//
// Rays outside the finite / sane range (wide_ray_setup returns false) and trees that are not nested, not finite, reference a
// triangle from two leaves or need more than kWideStack stack entries walk the binary image instead (restir_trace.cuh).
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/restir_layouts.h"

#if defined(__CUDACC__)
#define RESTIR_HD __host__ __device__ __forceinline__
#else
#define RESTIR_HD inline
#endif

namespace restir {

constexpr int kWideStack = 32; // entries of the walk's stack = levels of the wide tree (one entry per level: the pending siblings); deeper trees are not walked wide

// Slots [0, inner) are inner children: wide nodes firstChild + slot (the children of a node are numbered consecutively).  Slots
// [inner, count) are leaves: triangle RECORDS firstRecord + (slot - inner) — the 64-byte (p1, e1, e2 | leaf box) records are laid
// out in the order the wide leaves name them (triOrder), so neither kind of child needs an index word.  Slots >= count are
// empty: lo = 32767, hi = 0 on every axis, an inverted box that no ray hits.  The four words after the planes hold these
// numbers in the form the walk consumes them (one operation less each on the ALU pipe, the unit the walk is short of):
struct alignas(64) WideNode {
	uint32_t q[3][4];    // [axis][slot]: lo (15 bits) | hi (15 bits) << 16, grid cells
	uint32_t childGroup; // firstChild << 4: or-ed with the mask of the hit inner slots it IS the walk's next group
	uint32_t recBase;    // firstRecord - inner (mod 2^32): the record of leaf SLOT j is recBase + j
	uint32_t innerMask;  // (1 << inner) - 1: the inner slots
	uint32_t count;      // number of used slots, 2..4
};
RESTIR_HD uint32_t wide_inner_count(const WideNode &n) {
	return n.innerMask == 0u ? 0u : n.innerMask == 1u ? 1u : n.innerMask == 3u ? 2u : n.innerMask == 7u ? 3u : 4u;
}
static_assert(sizeof(WideNode) == 64, "wide node is 64 bytes");

// plane coordinate of grid value q on axis a:  h[a] + (0.5 + q / 65536) * K[a]   (K a power of two, h a float: all exact)
struct WideGrid {
	float h[3], K[3], H[3]; // H = max(|h|, |h + K|): bound of every grid coordinate
	float maxOrigin[3];     // rays starting farther than this from 0 on an axis walk the binary image (2^16 * K)
};

// what the builders quantise with: grid origin and cell size per axis, in double (cell a power of two: every product below is exact)
struct WideQuant {
	double gridMin[3], cell[3], invCell[3];
};
// The grid of a scene whose boxes lie in [lo, hi] (per axis K = 65536 cells, a power of two with 32767 cells >= the extent; h =
// origin - K / 2 a float).  false: coordinates out of the range the grid is defined for.
bool make_wide_grid(const double lo[3], const double hi[3], WideGrid &grid, WideQuant &quant);

// lo | hi << 16 of the fp32 interval [mn, mx] on one axis, rounded outwards; false when the decoded planes would not enclose it
// (host + device: the two builders must produce the same words)
RESTIR_HD bool wide_quantise(const WideQuant &g, int a, float mn, float mx, uint32_t &word) {
	double ql = floor(((double)mn - g.gridMin[a]) * g.invCell[a]), qh = ceil(((double)mx - g.gridMin[a]) * g.invCell[a]);
	ql = fmin(fmax(ql, 0.0), 32767.0);
	qh = fmin(fmax(qh, 0.0), 32767.0);
	if (!(g.gridMin[a] + ql * g.cell[a] <= (double)mn) || !(g.gridMin[a] + qh * g.cell[a] >= (double)mx)) {
		return false;
	}
	word = (uint32_t)ql | ((uint32_t)qh << 16);
	return true;
}

struct WideImageInfo {
	bool usable = false;
	std::string why;
	uint32_t nodes = 0;
	int depth = 0;
	int stackBound = 0;
};

// triOrder: record r holds triangle triOrder[r] (a permutation of the triangles that hang under a leaf; triangles no leaf
// names follow at the end).  leafBoxes: per RECORD the fp32 box its leaf carries in the uploaded tree (min.xyz, max.xyz).
bool build_wide_image(const restir_aabb_node *nodes, uint32_t nNodes, uint32_t nTris, std::vector<WideNode> &out, std::vector<uint32_t> &triOrder,
                      std::vector<float> &leafBoxes, WideGrid &grid, WideImageInfo &info);

// Host emulation of the device walk over the wide image of (nodes, tris) for n segments p1 -> p2 (3 floats each): shadowed[i]
// = 1 when hit; walkedWide[i] = 0 for the rays wide_ray_setup refuses (the device walks the binary image for those: shadowed[i]
// is then left 0 here); *visits = wide nodes visited in total.  false (with why) when the tree is not walked wide.  Tests only.
bool wide_walk_host(const restir_aabb_node *nodes, uint32_t nNodes, const restir_triangle *tris, uint32_t nTris, const float *p1, const float *p2,
                    uint64_t n, unsigned char *shadowed, unsigned char *walkedWide, uint64_t *visits, std::string &why);

// ---- the box test (host + device) -------------------------------------------------------------------------------------------

struct WideRay {
	float s[3], cLo[3], cHi[3]; // t_entry = v * s + cLo, t_exit = v * s + cHi
	uint32_t selNear[3], selFar[3]; // byte-permute selectors: which half of the (lo | hi << 16) word is the entry / exit plane
};

RESTIR_HD float wide_as_float(uint32_t u) {
	float f;
#if defined(__CUDA_ARCH__)
	f = __uint_as_float(u);
#else
	std::memcpy(&f, &u, 4);
#endif
	return f;
}

// Error budget (u = 2^-24, every operation round-to-nearest; T = (p - o) * inv in real arithmetic with the FLOAT inv, p a
// plane; Z = (|o| + H) * |inv| bounds |T|, |c| and every intermediate):
//   reference side  E = fl(fl(b - o) * inv):                     |E - T(b)|  <= 2.01 u Z
//   here            c = fl(fl(h - o) * inv):                     |c - (h - o) inv| <= 2.01 u Z
//                   cLo = fl(c - m), t = fma(v, s, cLo), s = K * inv exact (power of two), v exact:
//                   |t - (T(p') - m)| <= 2.01 u Z + u |c - m| + u |t| <= 4.1 u Z      (p' = h + v K, the quantised plane)
//   p' lies on the outer side of b (build_wide_image), so T(p') <= T(b) for an entry plane and >= for an exit plane.
//   => t_entry <= E_entry as soon as m >= 6.2 u Z; m = 16 u Z = 2^-20 Z leaves room for the roundings of m itself.
// Ranges (else false => the caller walks the binary image): o, d finite; 2^-40 <= |inv| <= 2^40; |o| <= 2^16 K.  With
// 2^-40 <= K <= 2^40 and H <= 2^10 K (build_wide_image) nothing overflows or becomes subnormal, and an empty slot (lo = 32767,
// hi = 0) has entry - exit >= 0.499 K |inv| - 2 m > 0: never hit.
RESTIR_HD bool wide_ray_setup(const WideGrid &g, const float o[3], const float d[3], const float inv[3], WideRay &r) {
	bool ok = true;
	for (int a = 0; a < 3; ++a) {
		float ai = fabsf(inv[a]), ao = fabsf(o[a]);
		ok = ok && ai >= 9.094947017729282e-13f && ai <= 1099511627776.0f && ao <= g.maxOrigin[a] && fabsf(d[a]) <= 3.0e38f;
		r.s[a] = g.K[a] * inv[a];
		float c = (g.h[a] - o[a]) * inv[a];
		float m = ((ao + g.H[a]) * ai) * 9.5367431640625e-07f;
		r.cLo[a] = c - m;
		r.cHi[a] = c + m;
		// result byte 0 <- 0x00 and byte 3 <- 0x3F of the constant operand (bytes 4..7), bytes 1..2 <- the chosen half
		bool pos = inv[a] > 0.0f;
		r.selNear[a] = pos ? 0x7104u : 0x7324u;
		r.selFar[a] = pos ? 0x7324u : 0x7104u;
	}
	return ok; // NaN anywhere compares false
}

RESTIR_HD float wide_plane_value(uint32_t word, uint32_t sel) {
#if defined(__CUDA_ARCH__)
	return __uint_as_float(__byte_perm(word, 0x3F000000u, sel));
#else
	uint32_t half = sel == 0x7104u ? (word & 0xffffu) : (word >> 16);
	return wide_as_float(0x3F000000u | (half << 8));
#endif
}

// fma.rn.sat.f32: the fused result clamped to [0, 1], NaN -> +0
RESTIR_HD float wide_fma_sat(float a, float b, float c) {
	float t = fmaf(a, b, c);
	return t > 0.0f ? (t < 1.0f ? t : 1.0f) : 0.0f;
}

// entry / exit parameters of slot c of node n (conservative: see above).  The segment is t in [0, 1]: with exact planes a slot
// would be hit when N = max(entry_x, entry_y, entry_z, 0) <= F = min(exit_x, exit_y, exit_z, 1).  Two refinements:
//   * STRICT: hit <=> N < F.  For a box the reference passes (rmin < 1, rmax >= rmin, rmax > 0) every entry parameter here lies
//     below the reference's and every exit parameter above it by more than 9 u Z > 0 (the margin m = 16 u Z against the 6.2 u Z the
//     roundings can use up), so exit - entry > 0, exit > 0 and entry < 1, strictly: N < F in all four clamp combinations, and the
//     difference of two different floats does not round to zero.  Nothing the reference passes sits on N == F.
//   * SATURATED x: the device kernel gets both clamps for free by saturating the x parameters (fma.rn.sat, restir_wide.cuh):
//     n' = max(sat(entry_x), entry_y, entry_z), f' = min(sat(exit_x), exit_y, exit_z).  If N < F then entry_x <= N < 1 and
//     exit_x >= F > 0, so sat(entry_x) = max(entry_x, 0), sat(exit_x) = min(exit_x, 1), n' = N, f' = F: hit.  A box that only the
//     clamp separates from the segment (beyond its end or behind its origin on x alone) collapses onto n' = f' = 1 or 0: the
//     strict test says missed, as it should (with <= every such box would be visited: measured, 2.7x the walk time).  In
//     general entry_x > 1 gives n' >= 1 >= f' and exit_x < 0 gives f' <= 0 <= n', misses like N >= F: n' < f' <=> N < F, the
//     saturated test visits exactly the boxes the clamped one does.
RESTIR_HD bool wide_box_hit(const WideNode &n, int c, const WideRay &r) {
	float tn[3], tf[3];
	for (int a = 0; a < 3; ++a) {
		const float vn = wide_plane_value(n.q[a][c], r.selNear[a]), vf = wide_plane_value(n.q[a][c], r.selFar[a]);
		tn[a] = a == 0 ? wide_fma_sat(vn, r.s[a], r.cLo[a]) : fmaf(vn, r.s[a], r.cLo[a]);
		tf[a] = a == 0 ? wide_fma_sat(vf, r.s[a], r.cHi[a]) : fmaf(vf, r.s[a], r.cHi[a]);
	}
	float nearT = fmaxf(tn[0], fmaxf(tn[1], tn[2])), farT = fminf(tf[0], fminf(tf[1], tf[2]));
	return nearT - farT < 0.0f; // the kernel reads the sign bit of this difference
}

} // namespace restir
