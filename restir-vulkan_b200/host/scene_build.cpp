// scene_build.cpp — host-side (once per scene) builders of the data the ReSTIR passes bind:
// the 2-wide AABB tree, the triangle-light list, the fallback random point lights and the Vose
// alias table.  They produce, byte for byte, what the reference's CPU code produces; tests compare
// against blobs dumped from the reference's own C++ (oracle/_ref/scene_baker).
//
//   restir_build_aabb_tree               <- src/aabbTreeBuilder.cpp:52-214 (AabbTree::build)
//   restir_collect_triangle_lights       <- src/misc.cpp:380-414
//   restir_generate_random_point_lights  <- src/misc.cpp:358-378
//   restir_create_alias_table            <- src/misc.cpp:418-497
//
// Host C++ like the reference's; compiled with -ffp-contract=off so that every float expression is
// evaluated as written (the reference is built for plain x86-64, where no FMA exists).

#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <random>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/restir_b200.h"

namespace {

struct Vec3 {
	float v[3];
	float &operator[](int i) { return v[i]; }
	float operator[](int i) const { return v[i]; }
};

// nvmath's nv_min / nv_max are `(a < b) ? a : b` / `(a > b) ? a : b` (thirdparty/nvmath/nvmath.inl:2415-2420)
inline float pickMin(float a, float b) { return (a < b) ? a : b; }
inline float pickMax(float a, float b) { return (a > b) ? a : b; }
inline void growMin(Vec3 &acc, const Vec3 &o) {
	for (int k = 0; k < 3; ++k) acc[k] = pickMin(acc[k], o[k]);
}
inline void growMax(Vec3 &acc, const Vec3 &o) {
	for (int k = 0; k < 3; ++k) acc[k] = pickMax(acc[k], o[k]);
}
// surfaceAreaHeuristic, aabbTreeBuilder.cpp:15-18
inline float halfArea(const Vec3 &lo, const Vec3 &hi) {
	float sx = hi[0] - lo[0], sy = hi[1] - lo[1], sz = hi[2] - lo[2];
	return sx * sy + sx * sz + sy * sz;
}

struct LeafRecord { // aabbTreeBuilder.cpp:28-32
	Vec3 centroid, lo, hi;
	int32_t geom;
	uint32_t bin;
};

struct Bin { // aabbTreeBuilder.cpp:33-50
	Vec3 lo{{FLT_MAX, FLT_MAX, FLT_MAX}};
	Vec3 hi{{-FLT_MAX, -FLT_MAX, -FLT_MAX}};
	size_t count = 0;
	float cost() const { return (float)count * halfArea(lo, hi); }
	void absorb(const Bin &o) {
		count += o.count;
		growMin(lo, o.lo);
		growMax(hi, o.hi);
	}
};

struct Job {
	int64_t slot; // -1: the root placeholder; otherwise node*2 + (0 left | 1 right) — where the child id is written
	size_t beg, end;
};

inline void setBox(float dst[4], const Vec3 &s) { // vec4(vec3) sets w = 1 (nvmath_types.h:392-397)
	dst[0] = s[0];
	dst[1] = s[1];
	dst[2] = s[2];
	dst[3] = 1.0f;
}

} // namespace

// Leaf records of the triangles, aabbForTriangle + centroid (aabbTreeBuilder.cpp:8-14, 70-75)
static void makeLeaves(const restir_triangle *tris, uint32_t n_triangles, std::vector<LeafRecord> &leaves) {
	leaves.resize(n_triangles);
	for (uint32_t i = 0; i < n_triangles; ++i) {
		LeafRecord &l = leaves[i];
		const float *a = tris[i].p1, *b = tris[i].p2, *c = tris[i].p3;
		for (int k = 0; k < 3; ++k) {
			float lo = a[k], hi = a[k];
			lo = pickMin(lo, b[k]);
			hi = pickMax(hi, b[k]);
			lo = pickMin(lo, c[k]);
			hi = pickMax(hi, c[k]);
			l.lo[k] = lo;
			l.hi[k] = hi;
			l.centroid[k] = 0.5f * (lo + hi);
		}
		l.geom = (int32_t)i;
		l.bin = 0;
	}
}

static inline void writeSlot(restir_aabb_node *nodes, int64_t slot, int32_t value) {
	if (slot < 0) {
		return; // dummyRoot (:81): always receives 0
	}
	restir_aabb_node &n = nodes[slot >> 1];
	if (slot & 1) {
		n.rightChild = value;
	} else {
		n.leftChild = value;
	}
}

// One BuildStep of the reference's loop (aabbTreeBuilder.cpp:86-209) on a range of more than two leaves: picks the
// split, partitions leaves[beg, end) in place with the reference's own swap sequence, fills node `id` and returns the
// pivot.  Touches nothing but its own range and its own node, so the steps of one breadth-first level are independent.
static size_t splitRange(std::vector<LeafRecord> &leaves, restir_aabb_node *nodes, int32_t id, size_t beg, size_t end) {
	constexpr size_t kBins = 12;
	// :107-122 bounds of centroids and of geometry
	Vec3 cLo = leaves[beg].centroid, cHi = cLo;
	Vec3 gLo = leaves[beg].lo, gHi = leaves[beg].hi;
	for (size_t i = beg + 1; i < end; ++i) {
		growMin(cLo, leaves[i].centroid);
		growMax(cHi, leaves[i].centroid);
		growMin(gLo, leaves[i].lo);
		growMax(gHi, leaves[i].hi);
	}
	const float outerArea = halfArea(gLo, gHi);
	// :124-129 split axis = widest centroid extent
	const float ext[3] = {cHi[0] - cLo[0], cHi[1] - cLo[1], cHi[2] - cLo[2]};
	int axis = ext[0] > ext[1] ? 0 : 1;
	if (ext[2] > ext[axis]) {
		axis = 2;
	}
	// :130-142 binning.  When every centroid coincides on the axis the quotient is 0/0 = NaN and the
	// reference's size_t cast is undefined; defined here as bin 0 (the median fallback then applies).
	Bin bins[kBins];
	const float binWidth = ext[axis] / (float)kBins;
	for (size_t i = beg; i < end; ++i) {
		LeafRecord &l = leaves[i];
		float q = (l.centroid[axis] - cLo[axis]) / binWidth;
		q = (q < 0.5f) ? 0.5f : q;
		q = (q > (float)kBins - 0.5f) ? (float)kBins - 0.5f : q;
		l.bin = (q == q) ? (uint32_t)q : 0u;
		Bin &b = bins[l.bin];
		growMin(b.lo, l.lo);
		growMax(b.hi, l.hi);
		++b.count;
	}
	// :143-151 suffix unions: rightOf[i] = bins[i+1..]
	Bin rightOf[kBins - 1];
	{
		Bin acc = bins[kBins - 1];
		for (size_t i = kBins - 1; i > 0;) {
			rightOf[--i] = acc;
			acc.absorb(bins[i]);
		}
	}
	// :152-171 cheapest of the 11 splits; an empty side costs 0 * inf = NaN and never wins
	size_t bestSplit = 0;
	Vec3 lLo{}, lHi{}, rLo{}, rHi{};
	{
		float bestCost = FLT_MAX;
		Bin leftAcc;
		for (size_t s = 0; s < kBins - 1; ++s) {
			leftAcc.absorb(bins[s]);
			const Bin &rightAcc = rightOf[s];
			float cost = 0.125f + (leftAcc.cost() + rightAcc.cost()) / outerArea;
			if (cost < bestCost) {
				bestCost = cost;
				bestSplit = s;
				lLo = leftAcc.lo;
				lHi = leftAcc.hi;
				rLo = rightAcc.lo;
				rHi = rightAcc.hi;
			}
		}
	}
	// :172-178 in-place partition, same swap sequence
	size_t pivot = beg;
	for (size_t i = beg; i < end; ++i) {
		if (leaves[i].bin <= bestSplit) {
			std::swap(leaves[i], leaves[pivot++]);
		}
	}
	// :179-196 everything on one side: split at the median, recompute both boxes
	if (pivot == beg || pivot == end) {
		pivot = (beg + end) / 2;
		lLo = leaves[beg].lo;
		lHi = leaves[beg].hi;
		for (size_t i = beg + 1; i < pivot; ++i) {
			growMin(lLo, leaves[i].lo);
			growMax(lHi, leaves[i].hi);
		}
		rLo = leaves[pivot].lo;
		rHi = leaves[pivot].hi;
		for (size_t i = pivot; i < end; ++i) {
			growMin(rLo, leaves[i].lo);
			growMax(rHi, leaves[i].hi);
		}
	}
	// :198-207
	restir_aabb_node &n = nodes[id];
	setBox(n.leftAabbMin, lLo);
	setBox(n.leftAabbMax, lHi);
	setBox(n.rightAabbMin, rLo);
	setBox(n.rightAabbMax, rHi);
	return pivot;
}

// :91-104
static void pairNode(const std::vector<LeafRecord> &leaves, restir_aabb_node *nodes, int32_t id, size_t beg) {
	const LeafRecord &l = leaves[beg], &r = leaves[beg + 1];
	restir_aabb_node &n = nodes[id];
	n.leftChild = ~l.geom;
	n.rightChild = ~r.geom;
	setBox(n.leftAabbMin, l.lo);
	setBox(n.leftAabbMax, l.hi);
	setBox(n.rightAabbMin, r.lo);
	setBox(n.rightAabbMax, r.hi);
}

extern "C" int restir_build_aabb_tree(const void *triangles, uint32_t n_triangles, void *nodes_out) {
	if (triangles == nullptr || nodes_out == nullptr || n_triangles < 2) {
		return RESTIR_E_INVALID; // the reference asserts on a 1-triangle scene (aabbTreeBuilder.cpp:212)
	}
	restir_aabb_node *nodes = static_cast<restir_aabb_node *>(nodes_out);
	std::memset(nodes, 0, sizeof(restir_aabb_node) * (size_t)(n_triangles - 1));
	std::vector<LeafRecord> leaves;
	makeLeaves(static_cast<const restir_triangle *>(triangles), n_triangles, leaves);

	int32_t nextNode = 0;
	std::vector<Job> fifo; // breadth-first, like the reference's std::deque (:80-85)
	fifo.reserve((size_t)n_triangles * 2);
	fifo.push_back(Job{-1, 0, n_triangles});
	for (size_t head = 0; head < fifo.size(); ++head) {
		const Job job = fifo[head];
		const size_t span = job.end - job.beg;
		if (span == 1) { // :88-90
			writeSlot(nodes, job.slot, ~leaves[job.beg].geom);
			continue;
		}
		int32_t id = nextNode++;
		writeSlot(nodes, job.slot, id);
		if (span == 2) {
			pairNode(leaves, nodes, id, job.beg);
			continue;
		}
		size_t pivot = splitRange(leaves, nodes, id, job.beg, job.end);
		fifo.push_back(Job{(int64_t)id * 2, job.beg, pivot});
		fifo.push_back(Job{(int64_t)id * 2 + 1, pivot, job.end});
	}
	return RESTIR_OK;
}

// The same tree, byte for byte, built level by level: the reference's queue is breadth-first, node ids are handed out in
// queue order and a step touches only its own leaf range, its own node and one child slot of its parent, so all steps of
// a level can run at once — on `n_threads` host threads here, and in the same formulation on a device (SURVEY.md §8f:
// rebuild for dynamic geometry).  Inside a step everything stays sequential and in the reference's order: its min/max
// (`a < b ? a : b`) is not commutative for signed zeros, and the partition is the reference's own swap sequence.
extern "C" int restir_build_aabb_tree_mt(const void *triangles, uint32_t n_triangles, void *nodes_out, uint32_t n_threads) {
	if (triangles == nullptr || nodes_out == nullptr || n_triangles < 2) {
		return RESTIR_E_INVALID;
	}
	if (n_threads == 0) {
		n_threads = std::max(1u, std::thread::hardware_concurrency());
	}
	restir_aabb_node *nodes = static_cast<restir_aabb_node *>(nodes_out);
	std::memset(nodes, 0, sizeof(restir_aabb_node) * (size_t)(n_triangles - 1));
	std::vector<LeafRecord> leaves;
	makeLeaves(static_cast<const restir_triangle *>(triangles), n_triangles, leaves);

	std::vector<Job> level{Job{-1, 0, n_triangles}}, next;
	std::vector<int32_t> ids;     // node id of every step of the level that makes a node
	std::vector<size_t> childAt;  // where a splitting step's two children go in the next level
	std::vector<size_t> pivots;
	int32_t nextNode = 0;
	while (!level.empty()) {
		// queue order fixes the node ids (alloc++ at :93 / :198) and the order of the next level (:208-209)
		ids.assign(level.size(), -1);
		childAt.assign(level.size(), 0);
		size_t children = 0;
		for (size_t j = 0; j < level.size(); ++j) {
			const size_t span = level[j].end - level[j].beg;
			if (span >= 2) {
				ids[j] = nextNode++;
			}
			if (span > 2) {
				childAt[j] = children;
				children += 2;
			}
		}
		next.assign(children, Job{0, 0, 0});
		std::atomic<size_t> cursor{0};
		// steps are handed out in queue order, a few at a time: one atomic per step would cost more than the small steps of
		// the deep levels themselves
		const size_t batch = std::max<size_t>(1, level.size() / ((size_t)n_threads * 8));
		auto work = [&]() {
			for (;;) {
				const size_t first = cursor.fetch_add(batch, std::memory_order_relaxed);
				if (first >= level.size()) {
					return;
				}
				const size_t last = std::min(level.size(), first + batch);
				for (size_t j = first; j < last; ++j) {
					const Job &job = level[j];
					const size_t span = job.end - job.beg;
					if (span == 1) {
						writeSlot(nodes, job.slot, ~leaves[job.beg].geom);
						continue;
					}
					writeSlot(nodes, job.slot, ids[j]);
					if (span == 2) {
						pairNode(leaves, nodes, ids[j], job.beg);
						continue;
					}
					size_t pivot = splitRange(leaves, nodes, ids[j], job.beg, job.end);
					next[childAt[j]] = Job{(int64_t)ids[j] * 2, job.beg, pivot};
					next[childAt[j] + 1] = Job{(int64_t)ids[j] * 2 + 1, pivot, job.end};
				}
			}
		};
		const size_t want = std::min<size_t>(n_threads, level.size());
		if (want <= 1) {
			work();
		} else {
			std::vector<std::thread> pool;
			pool.reserve(want - 1);
			for (size_t t = 1; t < want; ++t) {
				pool.emplace_back(work);
			}
			work();
			for (std::thread &t : pool) {
				t.join();
			}
		}
		level.swap(next);
	}
	return RESTIR_OK;
}

extern "C" int64_t restir_collect_triangle_lights(const void *triangles, const int32_t *tri_material, uint32_t n_triangles,
                                                   const float *material_emissive, uint32_t n_materials, restir_tri_light *out) {
	if (triangles == nullptr || tri_material == nullptr || material_emissive == nullptr || out == nullptr) {
		return RESTIR_E_INVALID;
	}
	const restir_triangle *tris = static_cast<const restir_triangle *>(triangles);
	int64_t count = 0;
	for (uint32_t i = 0; i < n_triangles; ++i) {
		int32_t m = tri_material[i];
		if (m < 0 || (uint32_t)m >= n_materials) {
			return RESTIR_E_INVALID;
		}
		const float *e = material_emissive + (size_t)m * 3;
		float sq = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
		if (!((double)sq > 1e-6)) { // misc.cpp:385
			continue;
		}
		const restir_triangle &t = tris[i];
		float ux = t.p2[0] - t.p1[0], uy = t.p2[1] - t.p1[1], uz = t.p2[2] - t.p1[2];
		float vx = t.p3[0] - t.p1[0], vy = t.p3[1] - t.p1[1], vz = t.p3[2] - t.p1[2];
		float nx = uy * vz - uz * vy, ny = uz * vx - ux * vz, nz = ux * vy - uy * vx; // nvmath cross, nvmath.inl:254-261
		float area = sqrtf(nx * nx + ny * ny + nz * nz);                              // :395
		nx /= area;                                                                   // :396
		ny /= area;
		nz /= area;
		area *= 0.5f;                                                                 // :397
		float lum = 0.2126f * e[0] + 0.7152f * e[1] + 0.0722f * e[2];               // common.glsl:7-9
		restir_tri_light &l = out[count++];
		std::memcpy(l.p1, t.p1, 16);
		std::memcpy(l.p2, t.p2, 16);
		std::memcpy(l.p3, t.p3, 16);
		l.emission_luminance[0] = e[0];
		l.emission_luminance[1] = e[1];
		l.emission_luminance[2] = e[2];
		l.emission_luminance[3] = lum;
		l.normalArea[0] = nx;
		l.normalArea[1] = ny;
		l.normalArea[2] = nz;
		l.normalArea[3] = area;
	}
	return count;
}

extern "C" int restir_generate_random_point_lights(uint64_t count, const float min_xyz[3], const float max_xyz[3], restir_point_light *out) {
	if (min_xyz == nullptr || max_xyz == nullptr || (out == nullptr && count != 0)) {
		return RESTIR_E_INVALID;
	}
	// The reference calls the standard library here (misc.cpp:364-367), so this does too: the sequence is
	// libstdc++'s std::default_random_engine (minstd_rand0), default-seeded — canonical for this repo.
	std::uniform_real_distribution<float> dx(min_xyz[0], max_xyz[0]), dy(min_xyz[1], max_xyz[1]), dz(min_xyz[2], max_xyz[2]);
	std::uniform_real_distribution<float> dr(0.0f, 1.0f), dg(0.0f, 1.0f), db(0.0f, 1.0f);
	std::default_random_engine engine;
	for (uint64_t i = 0; i < count; ++i) {
		restir_point_light &l = out[i];
		// misc.cpp:371-372 draws inside constructor argument lists, whose evaluation order C++ leaves
		// unspecified; g++ (and MSVC x64) evaluate them right to left, so the stream is z, y, x, then b, g, r.
		// Checked against the reference's own binary (tests/golden/random_lights_5000.npz).
		l.pos[2] = dz(engine);
		l.pos[1] = dy(engine);
		l.pos[0] = dx(engine);
		l.pos[3] = 1.0f;
		l.color_luminance[2] = db(engine);
		l.color_luminance[1] = dg(engine);
		l.color_luminance[0] = dr(engine);
		l.color_luminance[3] = 0.2126f * l.color_luminance[0] + 0.7152f * l.color_luminance[1] + 0.0722f * l.color_luminance[2];
	}
	return RESTIR_OK;
}

extern "C" int restir_create_alias_table(const restir_point_light *point, uint64_t n_point, const restir_tri_light *tri,
                                         uint64_t n_tri, restir_alias_column *out) {
	const uint64_t n = n_point != 0 ? n_point : n_tri;
	if (n == 0) {
		return RESTIR_OK;
	}
	if (out == nullptr || (n_point != 0 && point == nullptr) || (n_point == 0 && tri == nullptr)) {
		return RESTIR_E_INVALID;
	}
	// misc.cpp:426-444 powers and their float running sum, in light order
	std::vector<float> scaled(n);
	float total = 0.0f;
	for (uint64_t i = 0; i < n; ++i) {
		float power = n_point != 0 ? point[i].color_luminance[3] : tri[i].emission_luminance[3] * tri[i].normalArea[3];
		total += power;
		scaled[i] = power;
	}
	// :446-458
	std::vector<int> large, small; // FIFO queues with a read cursor (std::queue in the reference)
	size_t largeHead = 0, smallHead = 0;
	for (uint64_t i = 0; i < n; ++i) {
		out[i].prob = 0.0f;
		out[i].alias = -1;
		out[i].oriProb = scaled[i] / total;
		out[i].aliasOriProb = 0.0f;
		scaled[i] = (float)n * scaled[i] / total;
		(scaled[i] >= 1.0f ? large : small).push_back((int)i);
	}
	// :460-478
	while (largeHead < large.size() && smallHead < small.size()) {
		int g = large[largeHead++];
		int l = small[smallHead++];
		out[l].prob = scaled[l];
		out[l].alias = g;
		scaled[g] = (scaled[g] + scaled[l]) - 1.0f;
		(scaled[g] < 1.0f ? small : large).push_back(g);
	}
	// :480-492
	while (largeHead < large.size()) {
		int g = large[largeHead++];
		out[g].prob = 1.0f;
		out[g].alias = g;
	}
	while (smallHead < small.size()) {
		int l = small[smallHead++];
		out[l].prob = 1.0f;
		out[l].alias = l;
	}
	// :494-496
	for (uint64_t i = 0; i < n; ++i) {
		out[i].aliasOriProb = out[out[i].alias].oriProb;
	}
	return RESTIR_OK;
}

// Row-band boundaries of equal measured cost (no reference equivalent: the reference is single-GPU).  `seconds[r]` is
// what band r = [bounds_in[r], bounds_in[r+1]) cost; the cost is taken as uniform inside a measured band and the new
// boundaries cut its integral into n_bands equal parts, every band at least min_rows high (the halo must fit into a
// neighbour's band).  Same arithmetic as restir-vulkan_b200/bands.py::balanced_bounds (the tests hold the two together).
extern "C" int restir_band_balanced_bounds(uint32_t height, uint32_t n_bands, const uint32_t *bounds_in, const double *seconds,
                                            uint32_t min_rows, uint32_t *bounds_out) {
	if (n_bands == 0 || bounds_in == nullptr || seconds == nullptr || bounds_out == nullptr || bounds_in[0] != 0 || bounds_in[n_bands] != height ||
	    (uint64_t)min_rows * n_bands > height) {
		return RESTIR_E_INVALID;
	}
	std::vector<double> density(height);
	for (uint32_t r = 0; r < n_bands; ++r) {
		if (bounds_in[r + 1] <= bounds_in[r]) {
			return RESTIR_E_INVALID;
		}
		const uint32_t rows = bounds_in[r + 1] - bounds_in[r];
		const double d = std::max(seconds[r], 1e-9) / (double)rows;
		for (uint32_t y = bounds_in[r]; y < bounds_in[r + 1]; ++y) {
			density[y] = d;
		}
	}
	double total = 0.0;
	for (uint32_t y = 0; y < height; ++y) {
		total += density[y];
	}
	bounds_out[0] = 0;
	double acc = 0.0;
	uint32_t y = 0;
	for (uint32_t r = 1; r < n_bands; ++r) {
		const double target = total * (double)r / (double)n_bands;
		while (y < height && acc + density[y] <= target) {
			acc += density[y];
			++y;
		}
		const uint32_t lo = bounds_out[r - 1] + min_rows, hi = height - (n_bands - r) * min_rows;
		bounds_out[r] = std::min(std::max(y, lo), hi);
	}
	bounds_out[n_bands] = height;
	return RESTIR_OK;
}

