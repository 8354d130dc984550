// restir_bvh_build.cu — AabbTree::build (src/aabbTreeBuilder.cpp:52-214) on the device, byte for byte.
//
// The reference builds breadth-first from a queue: node ids are handed out in queue order and a build step touches only its
// own leaf range, its own node and one child slot of its parent.  So all steps of one queue generation ("level") are
// independent (host/scene_build.cpp restir_build_aabb_tree_mt runs them on host threads); here a level is a handful of flat
// kernels over all leaves and all jobs:
//
//   job_begin      node ids and child slots of the level, from two prefix sums over the jobs (queue order = id order);
//                  ranges of one and two leaves are finished on the spot (:88-104)
//   leaf_bounds    centroid and geometry bounds of every splitting range (:107-122)
//   job_axis       split axis and bin width (:124-131)
//   leaf_bin       bin of every leaf, bounds and counts of the 12 bins (:132-142)
//   job_split      the 11 candidate splits, the cheapest one, the node's boxes (:143-171, :198-207)
//   (prefix sum of the "goes left" flags)
//   job_children   pivot, median fallback with its own boxes (:179-196), the two child jobs (:208-209)
//   leaf_partition the in-place partition of :172-178 — the reference's swap sequence, reproduced as a permutation
//
// What makes the bytes come out the same:
//   * min / max are nvmath's `(a < b) ? a : b` / `(a > b) ? a : b` folded in leaf order: for values that compare equal (+0 and
//     -0) the LAST one wins.  Reductions here carry (value key, position) pairs ordered by key, then by later position — the
//     same fold, associatively — and the winning position's own bits are fetched;
//   * the partition `for i: if (left(i)) swap(a[i], a[pivot++])` keeps the left elements in order and rotates the right ones:
//     the element that ends at a position x of the right block is a[x] itself if it went right, else the element found by
//     following x -> beg + (number of left elements before x) until a right element is met (every swap moves the first
//     element of the right block to the scan position).  Pure function of the flags' prefix sums: no sequential pass;
//   * every float expression is the host builder's (-fmad=false), the 11 split costs are evaluated by one thread per job.
#include <cfloat>

#include "restir_kernels.h"

namespace restir {

namespace {

constexpr int kBins = 12;
constexpr unsigned long long kMinInit = ~0ull, kMaxInit = 0ull;

// a float as a key that orders like the float compares; +0 and -0 share a key (they compare equal)
__device__ __forceinline__ unsigned order_key(float f) {
	unsigned u = f == 0.0f ? 0u : __float_as_uint(f);
	return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ unsigned long long min_pack(float f, unsigned pos) { return ((unsigned long long)order_key(f) << 32) | (0xffffffffu - pos); }
__device__ __forceinline__ unsigned long long max_pack(float f, unsigned pos) { return ((unsigned long long)order_key(f) << 32) | pos; }
__device__ __forceinline__ unsigned min_pos(unsigned long long p) { return 0xffffffffu - (unsigned)(p & 0xffffffffull); }
__device__ __forceinline__ unsigned max_pos(unsigned long long p) { return (unsigned)(p & 0xffffffffull); }

__device__ __forceinline__ unsigned long long warp_min64(unsigned long long v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
		v = w < v ? w : v;
	}
	return v;
}
__device__ __forceinline__ unsigned long long warp_max64(unsigned long long v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
		v = w > v ? w : v;
	}
	return v;
}

__device__ __forceinline__ float pick_min(float a, float b) { return (a < b) ? a : b; }
__device__ __forceinline__ float pick_max(float a, float b) { return (a > b) ? a : b; }
__device__ __forceinline__ float half_area(const float lo[3], const float hi[3]) { // surfaceAreaHeuristic, aabbTreeBuilder.cpp:15-18
	float sx = hi[0] - lo[0], sy = hi[1] - lo[1], sz = hi[2] - lo[2];
	return sx * sy + sx * sz + sy * sz;
}
__device__ __forceinline__ void set_box(float *dst, const float s[3]) { // vec4(vec3): w = 1
	dst[0] = s[0];
	dst[1] = s[1];
	dst[2] = s[2];
	dst[3] = 1.0f;
}
__device__ __forceinline__ void write_slot(restir_aabb_node *nodes, long long slot, int value) {
	if (slot < 0) {
		return; // dummyRoot (:81)
	}
	restir_aabb_node &n = nodes[slot >> 1];
	if (slot & 1) {
		n.rightChild = value;
	} else {
		n.leftChild = value;
	}
}

} // namespace

// ---- leaves ------------------------------------------------------------------------------------------------------------
// aabbForTriangle + centroid (aabbTreeBuilder.cpp:8-14, 70-75)
__global__ void bvh_make_leaves_kernel(const float4 *__restrict__ tris, unsigned n, BvhLeaves out, int *__restrict__ jobOf) {
	unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) {
		return;
	}
	float4 a = tris[(size_t)i * 3], b = tris[(size_t)i * 3 + 1], c = tris[(size_t)i * 3 + 2];
	const float pa[3] = {a.x, a.y, a.z}, pb[3] = {b.x, b.y, b.z}, pc[3] = {c.x, c.y, c.z};
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		float lo = pa[k], hi = pa[k];
		lo = pick_min(lo, pb[k]);
		hi = pick_max(hi, pb[k]);
		lo = pick_min(lo, pc[k]);
		hi = pick_max(hi, pc[k]);
		out.lo[k][i] = lo;
		out.hi[k][i] = hi;
		out.cen[k][i] = 0.5f * (lo + hi);
	}
	out.geom[i] = (int)i;
	jobOf[i] = 0;
}

// ---- per level ---------------------------------------------------------------------------------------------------------
// The size of a level and the first node id it hands out live on the device: the host launches every level's kernels with grids
// sized by an upper bound (min(2^level, n) jobs) and looks at this record only every few levels, to see whether the queue is empty.
struct BvhLevel {
	unsigned nJobs;      // jobs of the level about to be processed
	int firstNode;       // id of the first node this level makes
	unsigned levels;     // levels processed so far that had jobs
	unsigned lastNodes;  // nodes made by the last such level (0: it only hung single leaves)
};

// flags for the two prefix sums over the jobs: makes a node (span >= 2), splits (span > 2); zero up to the launch bound
__global__ void bvh_job_flags_kernel(const BvhJob *__restrict__ jobs, const BvhLevel *__restrict__ lv, unsigned bound, unsigned *__restrict__ makesNode,
                                     unsigned *__restrict__ splits) {
	unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= bound) {
		return;
	}
	unsigned span = j < lv->nJobs ? (unsigned)(jobs[j].end - jobs[j].beg) : 0u;
	makesNode[j] = span >= 2u ? 1u : 0u;
	splits[j] = span > 2u ? 1u : 0u;
}

// after job_begin: the next level's size and first node id (queue order = id order)
__global__ void bvh_level_advance_kernel(BvhLevel *lv, const unsigned *__restrict__ nodeTotal, const unsigned *__restrict__ splitTotal) {
	if (lv->nJobs > 0u) {
		lv->levels += 1u;
		lv->lastNodes = *nodeTotal;
	}
	lv->firstNode += (int)*nodeTotal;
	lv->nJobs = 2u * *splitTotal;
}

__global__ void bvh_job_begin_kernel(BvhJob *__restrict__ jobs, const BvhLevel *__restrict__ lv, const unsigned *__restrict__ nodeScan,
                                     const unsigned *__restrict__ splitScan, BvhLeaves leaves, restir_aabb_node *__restrict__ nodes,
                                     BvhJobState *__restrict__ state) {
	unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= lv->nJobs) {
		return;
	}
	const int firstNode = lv->firstNode;
	BvhJob job = jobs[j];
	const unsigned span = (unsigned)(job.end - job.beg);
	if (span == 1u) { // :88-90
		write_slot(nodes, job.slot, ~leaves.geom[job.beg]);
		jobs[j].node = -1;
		jobs[j].split = -1;
		return;
	}
	const int id = firstNode + (int)nodeScan[j];
	write_slot(nodes, job.slot, id);
	jobs[j].node = id;
	if (span == 2u) { // :91-104
		restir_aabb_node &n = nodes[id];
		const int l = job.beg, r = job.beg + 1;
		n.leftChild = ~leaves.geom[l];
		n.rightChild = ~leaves.geom[r];
		const float llo[3] = {leaves.lo[0][l], leaves.lo[1][l], leaves.lo[2][l]}, lhi[3] = {leaves.hi[0][l], leaves.hi[1][l], leaves.hi[2][l]};
		const float rlo[3] = {leaves.lo[0][r], leaves.lo[1][r], leaves.lo[2][r]}, rhi[3] = {leaves.hi[0][r], leaves.hi[1][r], leaves.hi[2][r]};
		set_box(n.leftAabbMin, llo);
		set_box(n.leftAabbMax, lhi);
		set_box(n.rightAabbMin, rlo);
		set_box(n.rightAabbMax, rhi);
		jobs[j].split = -1;
		return;
	}
	const int s = (int)splitScan[j];
	jobs[j].split = s;
	BvhJobState &st = state[s];
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		st.cenMin[k] = kMinInit;
		st.geoMin[k] = kMinInit;
		st.cenMax[k] = kMaxInit;
		st.geoMax[k] = kMaxInit;
	}
	for (int b = 0; b < kBins; ++b) {
#pragma unroll
		for (int k = 0; k < 3; ++k) {
			st.binMin[b][k] = kMinInit;
			st.binMax[b][k] = kMaxInit;
		}
		st.binCount[b] = 0u;
	}
}

// :107-122: one (key, position) reduction per bound; whole warps inside one job reduce in registers first
__global__ void bvh_leaf_bounds_kernel(BvhLeaves leaves, const int *__restrict__ jobOf, const BvhJob *__restrict__ jobs, unsigned n,
                                       BvhJobState *__restrict__ state) {
	unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
	int j = i < n ? jobOf[i] : -1;
	int s = j >= 0 ? jobs[j].split : -1;
	const bool uniform = __all_sync(0xffffffffu, s == __shfl_sync(0xffffffffu, s, 0));
	if (uniform && s < 0) {
		return;
	}
	unsigned long long v[12];
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		if (s >= 0) {
			v[k] = min_pack(leaves.cen[k][i], i);
			v[3 + k] = max_pack(leaves.cen[k][i], i);
			v[6 + k] = min_pack(leaves.lo[k][i], i);
			v[9 + k] = max_pack(leaves.hi[k][i], i);
		}
	}
	if (uniform) {
#pragma unroll
		for (int k = 0; k < 3; ++k) {
			v[k] = warp_min64(v[k]);
			v[3 + k] = warp_max64(v[3 + k]);
			v[6 + k] = warp_min64(v[6 + k]);
			v[9 + k] = warp_max64(v[9 + k]);
		}
		if ((threadIdx.x & 31) != 0) {
			return;
		}
	} else if (s < 0) {
		return;
	}
	BvhJobState &st = state[s];
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		atomicMin(&st.cenMin[k], v[k]);
		atomicMax(&st.cenMax[k], v[3 + k]);
		atomicMin(&st.geoMin[k], v[6 + k]);
		atomicMax(&st.geoMax[k], v[9 + k]);
	}
}

// :124-131
__global__ void bvh_job_axis_kernel(const BvhJob *__restrict__ jobs, const BvhLevel *__restrict__ lv, BvhLeaves leaves, BvhJobState *__restrict__ state) {
	unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= lv->nJobs || jobs[j].split < 0) {
		return;
	}
	BvhJobState &st = state[jobs[j].split];
	float cLo[3], cHi[3], gLo[3], gHi[3];
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		cLo[k] = leaves.cen[k][min_pos(st.cenMin[k])];
		cHi[k] = leaves.cen[k][max_pos(st.cenMax[k])];
		gLo[k] = leaves.lo[k][min_pos(st.geoMin[k])];
		gHi[k] = leaves.hi[k][max_pos(st.geoMax[k])];
	}
	st.outerArea = half_area(gLo, gHi);
	const float ext[3] = {cHi[0] - cLo[0], cHi[1] - cLo[1], cHi[2] - cLo[2]};
	int axis = ext[0] > ext[1] ? 0 : 1;
	if (ext[2] > ext[axis]) {
		axis = 2;
	}
	st.axis = axis;
	st.axisLo = cLo[axis];
	st.binWidth = ext[axis] / (float)kBins;
}

// :132-142
__global__ void bvh_leaf_bin_kernel(BvhLeaves leaves, const int *__restrict__ jobOf, const BvhJob *__restrict__ jobs, unsigned n,
                                    BvhJobState *__restrict__ state) {
	unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) {
		return;
	}
	int j = jobOf[i];
	if (j < 0 || jobs[j].split < 0) {
		return;
	}
	BvhJobState &st = state[jobs[j].split];
	float q = (leaves.cen[st.axis][i] - st.axisLo) / st.binWidth;
	q = (q < 0.5f) ? 0.5f : q;
	q = (q > (float)kBins - 0.5f) ? (float)kBins - 0.5f : q;
	unsigned bin = (q == q) ? (unsigned)q : 0u; // NaN (every centroid on one plane: 0 / 0) is bin 0, as in the host builder
	leaves.bin[i] = bin;
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		atomicMin(&st.binMin[bin][k], min_pack(leaves.lo[k][i], i));
		atomicMax(&st.binMax[bin][k], max_pack(leaves.hi[k][i], i));
	}
	atomicAdd(&st.binCount[bin], 1u);
}

// :143-171 and :198-207.  One thread per job, the host builder's loop.
__global__ void bvh_job_split_kernel(const BvhJob *__restrict__ jobs, const BvhLevel *__restrict__ lv, BvhLeaves leaves, BvhJobState *__restrict__ state,
                                     restir_aabb_node *__restrict__ nodes) {
	unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= lv->nJobs || jobs[j].split < 0) {
		return;
	}
	BvhJobState &st = state[jobs[j].split];
	struct Bin {
		float lo[3], hi[3];
		unsigned count;
	};
	Bin bins[kBins];
	for (int b = 0; b < kBins; ++b) {
		bins[b].count = st.binCount[b];
#pragma unroll
		for (int k = 0; k < 3; ++k) {
			bins[b].lo[k] = bins[b].count ? leaves.lo[k][min_pos(st.binMin[b][k])] : FLT_MAX;
			bins[b].hi[k] = bins[b].count ? leaves.hi[k][max_pos(st.binMax[b][k])] : -FLT_MAX;
		}
	}
	// suffix unions: rightOf[i] = bins[i + 1 ..]
	Bin rightOf[kBins - 1];
	{
		Bin acc = bins[kBins - 1];
		for (int i = kBins - 1; i > 0;) {
			rightOf[--i] = acc;
			acc.count += bins[i].count;
#pragma unroll
			for (int k = 0; k < 3; ++k) {
				acc.lo[k] = pick_min(acc.lo[k], bins[i].lo[k]);
				acc.hi[k] = pick_max(acc.hi[k], bins[i].hi[k]);
			}
		}
	}
	int bestSplit = 0;
	float lLo[3] = {0.0f, 0.0f, 0.0f}, lHi[3] = {0.0f, 0.0f, 0.0f}, rLo[3] = {0.0f, 0.0f, 0.0f}, rHi[3] = {0.0f, 0.0f, 0.0f};
	{
		float bestCost = FLT_MAX;
		Bin left;
		left.count = 0;
#pragma unroll
		for (int k = 0; k < 3; ++k) {
			left.lo[k] = FLT_MAX;
			left.hi[k] = -FLT_MAX;
		}
		for (int s = 0; s < kBins - 1; ++s) {
			left.count += bins[s].count;
#pragma unroll
			for (int k = 0; k < 3; ++k) {
				left.lo[k] = pick_min(left.lo[k], bins[s].lo[k]);
				left.hi[k] = pick_max(left.hi[k], bins[s].hi[k]);
			}
			const Bin &right = rightOf[s];
			float costL = (float)left.count * half_area(left.lo, left.hi), costR = (float)right.count * half_area(right.lo, right.hi);
			float cost = 0.125f + (costL + costR) / st.outerArea;
			if (cost < bestCost) {
				bestCost = cost;
				bestSplit = s;
#pragma unroll
				for (int k = 0; k < 3; ++k) {
					lLo[k] = left.lo[k];
					lHi[k] = left.hi[k];
					rLo[k] = right.lo[k];
					rHi[k] = right.hi[k];
				}
			}
		}
	}
	st.bestSplit = bestSplit;
	restir_aabb_node &n = nodes[jobs[j].node];
	set_box(n.leftAabbMin, lLo);
	set_box(n.leftAabbMax, lHi);
	set_box(n.rightAabbMin, rLo);
	set_box(n.rightAabbMax, rHi);
}

// "goes left" flags of :173-178, for the prefix sum
__global__ void bvh_leaf_flags_kernel(BvhLeaves leaves, const int *__restrict__ jobOf, const BvhJob *__restrict__ jobs, unsigned n,
                                      const BvhJobState *__restrict__ state, unsigned *__restrict__ goesLeft) {
	unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) {
		return;
	}
	int j = jobOf[i];
	unsigned f = 0u;
	if (j >= 0 && jobs[j].split >= 0) {
		f = leaves.bin[i] <= (unsigned)state[jobs[j].split].bestSplit ? 1u : 0u;
	}
	goesLeft[i] = f;
}

// pivot, median fallback (:179-196), child jobs (:208-209)
__global__ void bvh_job_children_kernel(BvhJob *__restrict__ jobs, const BvhLevel *__restrict__ lv, const unsigned *__restrict__ leftScan, BvhLeaves leaves,
                                        restir_aabb_node *__restrict__ nodes, BvhJob *__restrict__ next) {
	unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= lv->nJobs || jobs[j].split < 0) {
		return;
	}
	BvhJob job = jobs[j];
	int pivot = job.beg + (int)(leftScan[job.end] - leftScan[job.beg]);
	int keep = 0;
	if (pivot == job.beg || pivot == job.end) {
		// nothing moved (every swap was a self-swap, or there was none): split the range as it stands at its median
		pivot = (int)(((long long)job.beg + job.end) / 2);
		keep = 1;
		float lLo[3], lHi[3], rLo[3], rHi[3];
#pragma unroll
		for (int k = 0; k < 3; ++k) {
			lLo[k] = leaves.lo[k][job.beg];
			lHi[k] = leaves.hi[k][job.beg];
			rLo[k] = leaves.lo[k][pivot];
			rHi[k] = leaves.hi[k][pivot];
		}
		for (int i = job.beg + 1; i < pivot; ++i) {
#pragma unroll
			for (int k = 0; k < 3; ++k) {
				lLo[k] = pick_min(lLo[k], leaves.lo[k][i]);
				lHi[k] = pick_max(lHi[k], leaves.hi[k][i]);
			}
		}
		for (int i = pivot; i < job.end; ++i) {
#pragma unroll
			for (int k = 0; k < 3; ++k) {
				rLo[k] = pick_min(rLo[k], leaves.lo[k][i]);
				rHi[k] = pick_max(rHi[k], leaves.hi[k][i]);
			}
		}
		restir_aabb_node &n = nodes[job.node];
		set_box(n.leftAabbMin, lLo);
		set_box(n.leftAabbMax, lHi);
		set_box(n.rightAabbMin, rLo);
		set_box(n.rightAabbMax, rHi);
	}
	jobs[j].pivot = pivot;
	jobs[j].keepOrder = keep;
	BvhJob l{}, r{};
	l.slot = (long long)job.node * 2;
	l.beg = job.beg;
	l.end = pivot;
	r.slot = (long long)job.node * 2 + 1;
	r.beg = pivot;
	r.end = job.end;
	next[2 * job.split] = l;
	next[2 * job.split + 1] = r;
}

// :172-178 as a permutation (see the header), into the other copy of the leaf arrays; and the job of every leaf in the next level
__global__ void bvh_leaf_partition_kernel(BvhLeaves in, BvhLeaves out, const int *__restrict__ jobOf, const BvhJob *__restrict__ jobs, unsigned n,
                                          const unsigned *__restrict__ leftScan, int *__restrict__ jobOfNext) {
	unsigned x = blockIdx.x * blockDim.x + threadIdx.x;
	if (x >= n) {
		return;
	}
	int j = jobOf[x];
	auto copy = [&](unsigned dst, unsigned src) {
#pragma unroll
		for (int k = 0; k < 3; ++k) {
			out.lo[k][dst] = in.lo[k][src];
			out.hi[k][dst] = in.hi[k][src];
			out.cen[k][dst] = in.cen[k][src];
		}
		out.geom[dst] = in.geom[src];
	};
	if (j < 0 || jobs[j].split < 0) { // finished ranges keep their place (nobody reads them again)
		copy(x, x);
		jobOfNext[x] = -1;
		return;
	}
	const BvhJob job = jobs[j];
	jobOfNext[x] = 2 * job.split + ((int)x >= job.pivot ? 1 : 0);
	if (job.keepOrder) {
		copy(x, x);
		return;
	}
	const unsigned base = leftScan[job.beg];
	const bool left = leftScan[x + 1] != leftScan[x];
	if (left) {
		copy((unsigned)job.beg + (leftScan[x] - base), x); // the left block keeps its order
	}
	if ((int)x >= job.pivot) {
		unsigned p = x;
		while (leftScan[p + 1] != leftScan[p]) { // a left element stood here: what the swaps put in its place came from further left
			p = (unsigned)job.beg + (leftScan[p] - base);
		}
		copy(x, p);
	}
}

// ---- exclusive prefix sum of 32-bit flags (n + 1 outputs: out[n] = total) ------------------------------------------------
constexpr int kScanBlock = 1024;

__global__ void __launch_bounds__(kScanBlock) scan_block_kernel(const unsigned *__restrict__ in, unsigned *__restrict__ out, unsigned n, unsigned *__restrict__ sums) {
	__shared__ unsigned warpSum[32];
	unsigned i = blockIdx.x * kScanBlock + threadIdx.x;
	unsigned v = i < n ? in[i] : 0u, incl = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
		if ((threadIdx.x & 31) >= o) incl += t;
	}
	if ((threadIdx.x & 31) == 31) warpSum[threadIdx.x >> 5] = incl;
	__syncthreads();
	if (threadIdx.x < 32) {
		unsigned w = warpSum[threadIdx.x], wi = w;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			unsigned t = __shfl_up_sync(0xffffffffu, wi, o);
			if (threadIdx.x >= o) wi += t;
		}
		warpSum[threadIdx.x] = wi - w;
		if (threadIdx.x == 31 && sums) sums[blockIdx.x] = wi;
	}
	__syncthreads();
	if (i < n) {
		out[i] = warpSum[threadIdx.x >> 5] + incl - v;
	}
}
__global__ void scan_add_kernel(unsigned *__restrict__ out, unsigned n, const unsigned *__restrict__ blockOffsets, const unsigned *__restrict__ in) {
	unsigned i = blockIdx.x * kScanBlock + threadIdx.x;
	if (i < n) {
		out[i] += blockOffsets[blockIdx.x];
	}
	if (i == n - 1) {
		out[n] = out[i] + in[i]; // the total
	}
}

// out must hold n + 1 values; tmp 2 * (n / 1024 + 2) + 2 values.  n <= 1024^3.
cudaError_t exclusive_scan_u32(const unsigned *in, unsigned *out, unsigned n, unsigned *tmp, cudaStream_t s) {
	if (n == 0) {
		return cudaMemsetAsync(out, 0, sizeof(unsigned), s);
	}
	const unsigned blocks = (n + kScanBlock - 1) / kScanBlock;
	unsigned *sums = tmp, *sumsScan = tmp + blocks + 1;
	scan_block_kernel<<<blocks, kScanBlock, 0, s>>>(in, out, n, sums);
	if (blocks > 1) {
		// offsets of the blocks: the same scan one level up (at most 1024^2 blocks)
		const unsigned blocks2 = (blocks + kScanBlock - 1) / kScanBlock;
		unsigned *sums2 = sumsScan + blocks + 1, *sums2Scan = sums2 + blocks2 + 1;
		scan_block_kernel<<<blocks2, kScanBlock, 0, s>>>(sums, sumsScan, blocks, sums2);
		if (blocks2 > 1) {
			scan_block_kernel<<<1, kScanBlock, 0, s>>>(sums2, sums2Scan, blocks2, nullptr);
			scan_add_kernel<<<blocks2, kScanBlock, 0, s>>>(sumsScan, blocks, sums2Scan, sums);
		}
	} else {
		cudaMemsetAsync(sumsScan, 0, sizeof(unsigned), s);
	}
	scan_add_kernel<<<blocks, kScanBlock, 0, s>>>(out, n, sumsScan, in);
	return cudaGetLastError();
}

// nodes (80 bytes, reference layout) -> the 64-byte traversal image (traversal_image.h), on the device
__global__ void bvh_image_kernel(const restir_aabb_node *__restrict__ nodes, unsigned n, float4 *__restrict__ image) {
	unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) {
		return;
	}
	const restir_aabb_node &nd = nodes[i];
	float4 *o = image + (size_t)i * 4;
	o[0] = make_float4(nd.leftAabbMin[0], nd.rightAabbMin[0], nd.leftAabbMax[0], nd.rightAabbMax[0]);
	o[1] = make_float4(nd.leftAabbMin[1], nd.rightAabbMin[1], nd.leftAabbMax[1], nd.rightAabbMax[1]);
	o[2] = make_float4(nd.leftAabbMin[2], nd.rightAabbMin[2], nd.leftAabbMax[2], nd.rightAabbMax[2]);
	o[3] = make_float4(__int_as_float(nd.leftChild), __int_as_float(nd.rightChild), 0.0f, 0.0f);
}

// any coordinate that is not finite: the min / max folds of the reference are order-dependent on NaN; such input goes to the host builder
__global__ void bvh_check_finite_kernel(const float4 *__restrict__ tris, unsigned n, unsigned *__restrict__ bad) {
	unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n * 3u) {
		return;
	}
	float4 p = tris[i];
	if (!(fabsf(p.x) <= FLT_MAX) || !(fabsf(p.y) <= FLT_MAX) || !(fabsf(p.z) <= FLT_MAX)) {
		atomicAdd(bad, 1u);
	}
}

// ---- the build ---------------------------------------------------------------------------------------------------------

size_t bvh_build_scratch_bytes(unsigned n) {
	const size_t leaves = 2 * (size_t)n * (9 * sizeof(float) + sizeof(int)) + (size_t)n * sizeof(unsigned);       // two copies + bins
	const size_t perLeaf = (size_t)n * (2 * sizeof(int) + sizeof(unsigned)) + ((size_t)n + 1) * sizeof(unsigned);   // jobOf x2, flags, scan
	const size_t jobs = 2 * ((size_t)n + 2) * sizeof(BvhJob) + 2 * ((size_t)n + 2) * sizeof(unsigned) * 2;         // two levels, flags + scans
	const size_t state = ((size_t)n / 3 + 2) * sizeof(BvhJobState);                                                // splitting ranges have >= 3 leaves
	const size_t tmp = (2 * ((size_t)n / kScanBlock + 4) + 2 * ((size_t)n / kScanBlock / kScanBlock + 4) + 8) * sizeof(unsigned);
	return leaves + perLeaf + jobs + state + tmp + 64 * 256; // every sub-array starts on a 256-byte boundary (about 40 of them)
}

// Builds the tree over `tris` (device, n x 3 float4) into `nodes` (device, (n - 1) x 80 bytes).  levelsOut = level of the deepest leaf (root node = level 0).
cudaError_t build_aabb_tree_device(const float4 *tris, unsigned n, restir_aabb_node *nodes, void *scratch, int *levelsOut, unsigned *nonFinite,
                                   cudaStream_t s) {
	unsigned char *p = static_cast<unsigned char *>(scratch);
	auto take = [&](size_t bytes) {
		void *r = p;
		p += (bytes + 255) & ~(size_t)255;
		return r;
	};
	BvhLeaves L[2];
	for (int c = 0; c < 2; ++c) {
		for (int k = 0; k < 3; ++k) {
			L[c].lo[k] = (float *)take((size_t)n * 4);
			L[c].hi[k] = (float *)take((size_t)n * 4);
			L[c].cen[k] = (float *)take((size_t)n * 4);
		}
		L[c].geom = (int *)take((size_t)n * 4);
	}
	unsigned *bins = (unsigned *)take((size_t)n * 4);
	L[0].bin = L[1].bin = bins;
	int *jobOf[2] = {(int *)take((size_t)n * 4), (int *)take((size_t)n * 4)};
	unsigned *goesLeft = (unsigned *)take((size_t)n * 4), *leftScan = (unsigned *)take(((size_t)n + 1) * 4);
	BvhJob *jobs[2] = {(BvhJob *)take(((size_t)n + 2) * sizeof(BvhJob)), (BvhJob *)take(((size_t)n + 2) * sizeof(BvhJob))};
	unsigned *makesNode = (unsigned *)take(((size_t)n + 2) * 4), *splits = (unsigned *)take(((size_t)n + 2) * 4);
	unsigned *nodeScan = (unsigned *)take(((size_t)n + 2) * 4), *splitScan = (unsigned *)take(((size_t)n + 2) * 4);
	BvhJobState *state = (BvhJobState *)take(((size_t)n / 3 + 2) * sizeof(BvhJobState));
	unsigned *tmp = (unsigned *)take((2 * ((size_t)n / kScanBlock + 4) + 2 * ((size_t)n / kScanBlock / kScanBlock + 4) + 8) * 4);
	unsigned *bad = (unsigned *)take(256);
	BvhLevel *lv = (BvhLevel *)take(256);

	const unsigned tb = 256;
	auto grid = [&](unsigned count) { return (count + tb - 1) / tb; };
	cudaError_t e;
	if ((e = cudaMemsetAsync(bad, 0, sizeof(unsigned), s)) != cudaSuccess) return e;
	bvh_check_finite_kernel<<<grid(n * 3), tb, 0, s>>>(tris, n, bad);
	if ((e = cudaMemsetAsync(nodes, 0, sizeof(restir_aabb_node) * (size_t)(n - 1), s)) != cudaSuccess) return e;
	bvh_make_leaves_kernel<<<grid(n), tb, 0, s>>>(tris, n, L[0], jobOf[0]);
	BvhJob root{};
	root.slot = -1;
	root.beg = 0;
	root.end = (int)n;
	if ((e = cudaMemcpyAsync(jobs[0], &root, sizeof(root), cudaMemcpyHostToDevice, s)) != cudaSuccess) return e;
	if ((e = cudaMemcpyAsync(nonFinite, bad, sizeof(unsigned), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;

	// the queue's first generation: the root range
	BvhLevel first{1u, 0, 0u, 0u};
	if ((e = cudaMemcpyAsync(lv, &first, sizeof(first), cudaMemcpyHostToDevice, s)) != cudaSuccess) return e;
	int cur = 0, levels = 0;
	for (unsigned level = 0;; ++level) {
		// a level has at most min(2^level, n) jobs: the grids and the scans are sized by that bound, the kernels read the level's
		// real size from `lv` — no host round trip per level (the build is launch-bound: 30 levels x ~18 launches for Sponza)
		const unsigned bound = level >= 31u ? n : std::min(n, 1u << level);
		BvhJob *J = jobs[cur & 1], *N = jobs[(cur & 1) ^ 1];
		BvhLeaves &in = L[cur & 1], &out = L[(cur & 1) ^ 1];
		bvh_job_flags_kernel<<<grid(bound), tb, 0, s>>>(J, lv, bound, makesNode, splits);
		if ((e = exclusive_scan_u32(makesNode, nodeScan, bound, tmp, s)) != cudaSuccess) return e;
		if ((e = exclusive_scan_u32(splits, splitScan, bound, tmp, s)) != cudaSuccess) return e;
		bvh_job_begin_kernel<<<grid(bound), tb, 0, s>>>(J, lv, nodeScan, splitScan, in, nodes, state);
		bvh_leaf_bounds_kernel<<<grid(n), tb, 0, s>>>(in, jobOf[cur & 1], J, n, state);
		bvh_job_axis_kernel<<<grid(bound), tb, 0, s>>>(J, lv, in, state);
		bvh_leaf_bin_kernel<<<grid(n), tb, 0, s>>>(in, jobOf[cur & 1], J, n, state);
		bvh_job_split_kernel<<<grid(bound), tb, 0, s>>>(J, lv, in, state, nodes);
		bvh_leaf_flags_kernel<<<grid(n), tb, 0, s>>>(in, jobOf[cur & 1], J, n, state, goesLeft);
		if ((e = exclusive_scan_u32(goesLeft, leftScan, n, tmp, s)) != cudaSuccess) return e;
		bvh_job_children_kernel<<<grid(bound), tb, 0, s>>>(J, lv, leftScan, in, nodes, N);
		bvh_leaf_partition_kernel<<<grid(n), tb, 0, s>>>(in, out, jobOf[cur & 1], J, n, leftScan, jobOf[(cur & 1) ^ 1]);
		bvh_level_advance_kernel<<<1, 1, 0, s>>>(lv, nodeScan + bound, splitScan + bound);
		if ((e = cudaGetLastError()) != cudaSuccess) return e;
		++cur;
		if ((level & 7u) == 7u || level >= 64u) { // is the queue empty?  (levels processed after it emptied are no-ops)
			BvhLevel now;
			if ((e = cudaMemcpyAsync(&now, lv, sizeof(now), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
			if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
			if (now.nJobs == 0u) {
				// a last generation of single leaves only hangs them one level higher than one that still made nodes
				levels = (int)now.levels - (now.lastNodes == 0u ? 1 : 0);
				break;
			}
		}
	}
	*levelsOut = levels;
	return cudaGetLastError();
}

void launch_bvh_image(const restir_aabb_node *nodes, unsigned n, float4 *image, cudaStream_t s) {
	if (n) bvh_image_kernel<<<(n + 255) / 256, 256, 0, s>>>(nodes, n, image);
}

cudaError_t preload_bvh_build_kernels() {
	cudaFuncAttributes a;
	cudaError_t e = cudaSuccess;
	const void *kernels[] = {(const void *)bvh_make_leaves_kernel,  (const void *)bvh_job_flags_kernel,    (const void *)bvh_job_begin_kernel,
	                         (const void *)bvh_leaf_bounds_kernel,  (const void *)bvh_job_axis_kernel,     (const void *)bvh_leaf_bin_kernel,
	                         (const void *)bvh_job_split_kernel,    (const void *)bvh_leaf_flags_kernel,   (const void *)bvh_job_children_kernel,
	                         (const void *)bvh_leaf_partition_kernel, (const void *)scan_block_kernel,      (const void *)scan_add_kernel,
	                         (const void *)bvh_image_kernel,        (const void *)bvh_check_finite_kernel,
	                         (const void *)bvh_level_advance_kernel};
	for (const void *k : kernels) {
		if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k);
	}
	return e;
}

} // namespace restir
