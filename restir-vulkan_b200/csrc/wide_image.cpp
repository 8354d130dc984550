// wide_image.cpp — see wide_image.h.  Host C++, runs once per restir_upload_bvh after build_traversal_image has accepted the
// upload as a tree (every child index in range, every node reached once).

#include "wide_image.h"

#include <algorithm>
#include <cstdio>

namespace restir {

namespace {

struct Slot {
	int32_t ref; // the binary tree's child word: >= 0 node, < 0 ~triangle
	float mn[3], mx[3];
};

inline double slot_area(const Slot &s) {
	double e[3] = {(double)s.mx[0] - s.mn[0], (double)s.mx[1] - s.mn[1], (double)s.mx[2] - s.mn[2]};
	return e[0] * e[1] + e[1] * e[2] + e[2] * e[0];
}

inline void node_slots(const restir_aabb_node &n, Slot &l, Slot &r) {
	l.ref = n.leftChild;
	r.ref = n.rightChild;
	for (int a = 0; a < 3; ++a) {
		l.mn[a] = n.leftAabbMin[a];
		l.mx[a] = n.leftAabbMax[a];
		r.mn[a] = n.rightAabbMin[a];
		r.mx[a] = n.rightAabbMax[a];
	}
}

inline bool finite_box(const Slot &s) {
	for (int a = 0; a < 3; ++a) {
		if (!std::isfinite(s.mn[a]) || !std::isfinite(s.mx[a]) || s.mn[a] > s.mx[a]) {
			return false;
		}
	}
	return true;
}

inline bool inside(const Slot &c, const Slot &p) {
	for (int a = 0; a < 3; ++a) {
		if (c.mn[a] < p.mn[a] || c.mx[a] > p.mx[a]) {
			return false;
		}
	}
	return true;
}

} // namespace

bool make_wide_grid(const double lo[3], const double hi[3], WideGrid &grid, WideQuant &quant) {
	for (int a = 0; a < 3; ++a) {
		if (!std::isfinite(lo[a]) || !std::isfinite(hi[a]) || lo[a] > hi[a]) {
			return false;
		}
		double extent = hi[a] - lo[a], H0 = std::max(std::fabs(lo[a]), std::fabs(hi[a]));
		double want = std::max({2.001 * extent, H0 / 512.0, std::ldexp(1.0, -40)});
		int e;
		std::frexp(want, &e); // want = f * 2^e, 0.5 <= f < 1
		for (;; ++e) {
			double K = std::ldexp(1.0, e);
			float h = (float)(lo[a] - 0.5 * K);
			if ((double)h + 0.5 * K > lo[a]) {
				h = std::nextafterf(h, -INFINITY);
			}
			double gmin = (double)h + 0.5 * K, c = K / 65536.0;
			if (gmin <= lo[a] && gmin + 32767.0 * c >= hi[a]) {
				grid.h[a] = h;
				grid.K[a] = (float)K;
				grid.H[a] = (float)std::max(std::fabs((double)h), std::fabs((double)h + K)) * 1.0000002f;
				grid.maxOrigin[a] = (float)(K * 65536.0);
				quant.gridMin[a] = gmin;
				quant.cell[a] = c;
				quant.invCell[a] = 65536.0 / K;
				break;
			}
		}
		if (!(grid.K[a] <= 1099511627776.0f) || !(grid.H[a] <= 1024.0f * grid.K[a])) {
			return false;
		}
	}
	return true;
}

bool build_wide_image(const restir_aabb_node *nodes, uint32_t nNodes, uint32_t nTris, std::vector<WideNode> &out, std::vector<uint32_t> &triOrder,
                      std::vector<float> &leafBoxes, WideGrid &grid, WideImageInfo &info) {
	out.clear();
	triOrder.clear();
	leafBoxes.clear();
	info = WideImageInfo{};
	char msg[256];
	auto refuse = [&](const char *why) {
		info.why = why;
		out.clear();
		triOrder.clear();
		leafBoxes.clear();
		return true; // not an error: the binary image is walked instead
	};

	// ---- what (*) of wide_image.h needs: finite boxes, nested boxes, every triangle in exactly one leaf ----------------------
	std::vector<uint8_t> triSeen(nTris, 0);
	double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
	{
		std::vector<std::pair<int32_t, Slot>> todo; // (node, the box its parent stores for it)
		Slot all;
		all.ref = 0;
		for (int a = 0; a < 3; ++a) {
			all.mn[a] = -INFINITY;
			all.mx[a] = INFINITY;
		}
		todo.push_back({0, all});
		while (!todo.empty()) {
			auto [n, parentBox] = todo.back();
			todo.pop_back();
			Slot s[2];
			node_slots(nodes[n], s[0], s[1]);
			for (int k = 0; k < 2; ++k) {
				if (!finite_box(s[k])) {
					std::snprintf(msg, sizeof(msg), "node %d stores a box that is not finite or has min > max", n);
					return refuse(msg);
				}
				if (!inside(s[k], parentBox)) {
					std::snprintf(msg, sizeof(msg), "the boxes of node %d are not inside the box its parent stores for it (not nested)", n);
					return refuse(msg);
				}
				for (int a = 0; a < 3; ++a) {
					lo[a] = std::min(lo[a], (double)s[k].mn[a]);
					hi[a] = std::max(hi[a], (double)s[k].mx[a]);
				}
				if (s[k].ref >= 0) {
					todo.push_back({s[k].ref, s[k]});
				} else {
					uint32_t t = (uint32_t)~s[k].ref;
					if (triSeen[t]) {
						std::snprintf(msg, sizeof(msg), "triangle %u is referenced by more than one leaf", t);
						return refuse(msg);
					}
					triSeen[t] = 1;
				}
			}
		}
	}

	WideQuant quant;
	if (!make_wide_grid(lo, hi, grid, quant)) {
		return refuse("scene coordinates out of the range the quantised grid is defined for");
	}
	auto quantise = [&](const Slot &s, int a, uint32_t &word) { return wide_quantise(quant, a, s.mn[a], s.mx[a], word); };

	// ---- collapse: a wide node takes the two children of a binary node and, while it has room, opens the inner child of
	// largest surface area into its own two children ------------------------------------------------------------------------------
	std::vector<int32_t> binaryOf{0}; // wide node -> the binary node it stands for (breadth-first numbering)
	std::vector<int32_t> depthOf{0};
	out.reserve(nNodes / 2 + 1);
	for (size_t w = 0; w < binaryOf.size(); ++w) {
		Slot s[4];
		int count = 2;
		node_slots(nodes[binaryOf[w]], s[0], s[1]);
		while (count < 4) {
			int best = -1;
			double bestArea = -1.0;
			for (int k = 0; k < count; ++k) {
				if (s[k].ref >= 0) {
					double area = slot_area(s[k]);
					if (area > bestArea) {
						bestArea = area;
						best = k;
					}
				}
			}
			if (best < 0) {
				break;
			}
			Slot l, r;
			node_slots(nodes[s[best].ref], l, r);
			for (int k = count; k > best + 1; --k) {
				s[k] = s[k - 1];
			}
			s[best] = l;
			s[best + 1] = r;
			++count;
		}
		// inner children first (the order of the slots is free: the answer is any-hit), each group in the tree's own order
		std::stable_partition(s, s + count, [](const Slot &x) { return x.ref >= 0; });
		WideNode n;
		const uint32_t firstChild = (uint32_t)binaryOf.size(), firstRecord = (uint32_t)triOrder.size();
		uint32_t inner = 0;
		if (binaryOf.size() >= (1u << 28)) {
			return refuse("more wide nodes than a group word can name");
		}
		n.count = (uint32_t)count;
		for (int c = 0; c < 4; ++c) {
			if (c >= count) {
				for (int a = 0; a < 3; ++a) {
					n.q[a][c] = 32767u; // lo = 32767, hi = 0: inverted
				}
				continue;
			}
			for (int a = 0; a < 3; ++a) {
				if (!quantise(s[c], a, n.q[a][c])) {
					return refuse("a box does not fit the quantised grid");
				}
			}
			if (s[c].ref >= 0) {
				++inner;
				binaryOf.push_back(s[c].ref);
				depthOf.push_back(depthOf[w] + 1);
			} else {
				triOrder.push_back((uint32_t)~s[c].ref);
				for (int a = 0; a < 3; ++a) {
					leafBoxes.push_back(s[c].mn[a]);
				}
				for (int a = 0; a < 3; ++a) {
					leafBoxes.push_back(s[c].mx[a]);
				}
			}
		}
		n.childGroup = firstChild << 4;
		n.recBase = firstRecord - inner;
		n.innerMask = (1u << inner) - 1u;
		info.depth = std::max(info.depth, depthOf[w] + 1);
		out.push_back(n);
	}
	// triangles no leaf names keep a record (the binary image never names them either)
	for (uint32_t t = 0; t < nTris; ++t) {
		if (!triSeen[t]) {
			triOrder.push_back(t);
			leafBoxes.insert(leafBoxes.end(), 6, 0.0f);
		}
	}

	// the walk (restir_wide.cuh trace_any_wide) keeps the pending siblings of one level in one stack entry
	info.stackBound = info.depth;
	info.nodes = (uint32_t)out.size();
	if (info.stackBound > kWideStack) {
		std::snprintf(msg, sizeof(msg), "the wide tree has %d levels (the walk's stack holds %d)", info.stackBound, kWideStack);
		return refuse(msg);
	}
	info.usable = true;
	return true;
}

} // namespace restir

// ---- host emulation of the device walk (restir_wide.cuh), for the CPU tests -------------------------------------------------------
// Same node format, same box arithmetic (wide_image.h, shared with the device), same leaf rule; the two exact factors are
// restated here for the host (compiled with -ffp-contract=off, like every host file of the library).  No pass calls this.
namespace restir {

namespace {

struct H3 {
	float x, y, z;
};
inline H3 hsub(H3 a, H3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float hdot(H3 a, H3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline H3 hcross(H3 a, H3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

// softwareRaytracing.glsl:9-14 (division hoisted: inv = 1 / dir)
inline bool host_box(H3 o, H3 inv, const float *mn, const float *mx) {
	float t1x = (mn[0] - o.x) * inv.x, t1y = (mn[1] - o.y) * inv.y, t1z = (mn[2] - o.z) * inv.z;
	float t2x = (mx[0] - o.x) * inv.x, t2y = (mx[1] - o.y) * inv.y, t2z = (mx[2] - o.z) * inv.z;
	float rmin = fmaxf(fminf(t1x, t2x), fmaxf(fminf(t1y, t2y), fminf(t1z, t2z)));
	float rmax = fminf(fmaxf(t1x, t2x), fminf(fmaxf(t1y, t2y), fmaxf(t1z, t2z)));
	return rmin < 1.0f && rmax >= rmin && rmax > 0.0f;
}

// softwareRaytracing.glsl:15-37
inline bool host_triangle(const restir_triangle &t, H3 o, H3 d) {
	H3 p1{t.p1[0], t.p1[1], t.p1[2]};
	H3 e1 = hsub(H3{t.p2[0], t.p2[1], t.p2[2]}, p1), e2 = hsub(H3{t.p3[0], t.p3[1], t.p3[2]}, p1);
	H3 p = hcross(d, e2);
	float f = 1.0f / hdot(e1, p);
	H3 s = hsub(o, p1);
	float bx = f * hdot(s, p);
	if (bx < 0.0f || bx > 1.0f) return false;
	H3 q = hcross(s, e1);
	float by = f * hdot(d, q);
	if (by < 0.0f || by + bx > 1.0f) return false;
	f = f * hdot(e2, q);
	return f > 0.0f && f < 1.0f;
}

} // namespace

bool wide_walk_host(const restir_aabb_node *nodes, uint32_t nNodes, const restir_triangle *tris, uint32_t nTris, const float *p1, const float *p2,
                    uint64_t n, unsigned char *shadowed, unsigned char *walkedWide, uint64_t *visits, std::string &why) {
	std::vector<WideNode> wide;
	std::vector<uint32_t> order;
	std::vector<float> leaf;
	WideGrid g;
	WideImageInfo info;
	build_wide_image(nodes, nNodes, nTris, wide, order, leaf, g, info);
	if (!info.usable) {
		why = info.why;
		return false;
	}
	uint64_t total = 0;
	for (uint64_t i = 0; i < n; ++i) {
		// visibilityTest.glsl:1-4, 27-28 (restir_trace.cuh segment_setup)
		H3 a{p1[i * 3], p1[i * 3 + 1], p1[i * 3 + 2]}, b{p2[i * 3], p2[i * 3 + 1], p2[i * 3 + 2]};
		H3 dir = hsub(b, a);
		float il = 1.0f / sqrtf(hdot(dir, dir));
		H3 nrm{dir.x * il, dir.y * il, dir.z * il};
		H3 off{nrm.x * 0.001f, nrm.y * 0.001f, nrm.z * 0.001f};
		H3 o{a.x + off.x, a.y + off.y, a.z + off.z}, d{dir.x - off.x * 2.0f, dir.y - off.y * 2.0f, dir.z - off.z * 2.0f};
		H3 inv{1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
		const float of[3] = {o.x, o.y, o.z}, df[3] = {d.x, d.y, d.z}, iv[3] = {inv.x, inv.y, inv.z};
		WideRay wr;
		if (!wide_ray_setup(g, of, df, iv, wr)) {
			walkedWide[i] = 0; // the device walks the binary image for this ray
			shadowed[i] = 0;
			continue;
		}
		walkedWide[i] = 1;
		bool hit = false;
		unsigned stack[kWideStack];
		int top = 0;
		unsigned group = 1u;
		for (;;) {
			if ((group & 15u) == 0u) {
				if (top == 0) break;
				group = stack[--top];
			}
			unsigned slot = (unsigned)__builtin_ctz(group & 15u);
			group &= group - 1u;
			const WideNode &nd = wide[(group >> 4) + slot];
			++total;
			unsigned hits = 0;
			for (int c = 0; c < 4; ++c) {
				if (wide_box_hit(nd, c, wr)) hits |= 1u << c;
			}
			for (unsigned lf = hits & ~nd.innerMask; lf != 0u && !hit; lf &= lf - 1u) {
				uint32_t rec = nd.recBase + (unsigned)__builtin_ctz(lf);
				hit = host_triangle(tris[order[rec]], o, d) && host_box(o, inv, &leaf[(size_t)rec * 6], &leaf[(size_t)rec * 6 + 3]);
			}
			if (hit) break;
			unsigned inner = hits & nd.innerMask;
			if (inner != 0u) {
				if ((group & 15u) != 0u) stack[top++] = group;
				group = nd.childGroup | inner;
			}
		}
		shadowed[i] = hit ? 1 : 0;
	}
	if (visits) *visits = total;
	return true;
}

} // namespace restir
