// restir_wide_build.cu — the 4-wide image (wide_image.h) of a tree that lives on the device: what restir_build_bvh_device
// installs, so that a tree rebuilt on the GPU is walked like an uploaded one.  Same algorithm as build_wide_image
// (wide_image.cpp), level-synchronous: the wide nodes of a level are expanded in parallel (one thread each: open the inner slot
// of largest surface area until four slots are filled, inner slots first, quantise), two prefix sums over the level give
// every node its children's indices (consecutive, level by level = the host's breadth-first numbering) and its triangle
// records (consecutive in node order = the host's record order), a second kernel links.  The result equals the host builder's
// byte for byte (tests/test_gpu_builders.py): same double-precision areas in the same order, same quantiser (wide_quantise).
#include "restir_kernels.h"
#include "wide_image.h"

#include <algorithm>

namespace restir {

cudaError_t exclusive_scan_u32(const unsigned *in, unsigned *out, unsigned n, unsigned *tmp, cudaStream_t s); // restir_bvh_build.cu

namespace {

struct WSlot {
	int ref;   // the binary tree's child word
	int src;   // binary node * 2 + side: where its fp32 box is stored
	float mn[3], mx[3];
};

__device__ __forceinline__ void load_slots(const restir_aabb_node *__restrict__ nodes, int n, WSlot &l, WSlot &r) {
	const float4 *p = reinterpret_cast<const float4 *>(nodes + n);
	const float4 a = p[0], b = p[1], c = p[2], d = p[3];
	const int2 ch = *reinterpret_cast<const int2 *>(p + 4);
	l.ref = ch.x; l.src = n * 2;
	l.mn[0] = a.x; l.mn[1] = a.y; l.mn[2] = a.z; l.mx[0] = b.x; l.mx[1] = b.y; l.mx[2] = b.z;
	r.ref = ch.y; r.src = n * 2 + 1;
	r.mn[0] = c.x; r.mn[1] = c.y; r.mn[2] = c.z; r.mx[0] = d.x; r.mx[1] = d.y; r.mx[2] = d.z;
}

__device__ __forceinline__ double slot_area(const WSlot &s) {
	const double ex = (double)s.mx[0] - s.mn[0], ey = (double)s.mx[1] - s.mn[1], ez = (double)s.mx[2] - s.mn[2];
	return ex * ey + ey * ez + ez * ex;
}

// One thread per wide node of the level [begin, begin + count): its slots, its quantised boxes, what the link step needs.
__global__ void wide_expand_kernel(const restir_aabb_node *__restrict__ nodes, const int *__restrict__ wideBinary, unsigned begin, unsigned count, WideQuant quant,
                                   WideNode *__restrict__ wide, int2 *__restrict__ slotRef, unsigned *__restrict__ innerCount, unsigned *__restrict__ leafCount,
                                   unsigned *__restrict__ bad) {
	const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count) {
		return;
	}
	const unsigned w = begin + i;
	WSlot s[4];
	int n = 2;
	load_slots(nodes, wideBinary[w], s[0], s[1]);
	while (n < 4) {
		int best = -1;
		double bestArea = -1.0;
		for (int k = 0; k < 4; ++k) {
			if (k < n && s[k].ref >= 0) {
				const double area = slot_area(s[k]);
				if (area > bestArea) {
					bestArea = area;
					best = k;
				}
			}
		}
		if (best < 0) {
			break;
		}
		WSlot l, r;
		load_slots(nodes, s[best].ref, l, r);
		// the opened slot is replaced by its two children, in place (static indexing: the slots live in registers)
#pragma unroll
		for (int k = 3; k >= 1; --k) {
			if (k > best + 1 && k <= n) {
				s[k] = s[k - 1];
			}
		}
#pragma unroll
		for (int k = 0; k < 4; ++k) {
			if (k == best) s[k] = l;
			if (k == best + 1) s[k] = r;
		}
		++n;
	}
	// inner slots first, each group in the tree's own order (std::stable_partition)
	WSlot t[4];
	int inner = 0, m = 0;
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		if (k < n && s[k].ref >= 0) {
#pragma unroll
			for (int q = 0; q < 4; ++q) {
				if (q == m) t[q] = s[k];
			}
			++m;
		}
	}
	inner = m;
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		if (k < n && s[k].ref < 0) {
#pragma unroll
			for (int q = 0; q < 4; ++q) {
				if (q == m) t[q] = s[k];
			}
			++m;
		}
	}
	WideNode out;
	out.childGroup = out.recBase = 0; // wide_link_kernel
	out.innerMask = (1u << inner) - 1u;
	out.count = (unsigned)n;
	bool ok = true;
#pragma unroll
	for (int c = 0; c < 4; ++c) {
		int2 ref = make_int2(0, -1);
		if (c < n) {
			for (int a = 0; a < 3; ++a) {
				ok = wide_quantise(quant, a, t[c].mn[a], t[c].mx[a], out.q[a][c]) && ok;
			}
			ref = make_int2(t[c].ref, t[c].src);
		} else {
			for (int a = 0; a < 3; ++a) {
				out.q[a][c] = 32767u; // lo = 32767, hi = 0: inverted
			}
		}
		slotRef[(size_t)w * 4 + c] = ref;
	}
	wide[w] = out;
	innerCount[i] = (unsigned)inner;
	leafCount[i] = (unsigned)(n - inner);
	if (!ok) {
		atomicAdd(bad, 1u);
	}
}

// Children and records of the level's nodes, from the prefix sums: node i's children are the wide nodes levelEnd + innerScan[i]
// ..., its records triTotal + leafScan[i] ...
__global__ void wide_link_kernel(const restir_aabb_node *__restrict__ nodes, unsigned begin, unsigned count, const unsigned *__restrict__ innerScan,
                                 const unsigned *__restrict__ leafScan, unsigned triTotal, WideNode *__restrict__ wide, const int2 *__restrict__ slotRef,
                                 int *__restrict__ wideBinary, unsigned *__restrict__ triOrder, float *__restrict__ leafBoxes, unsigned *__restrict__ recordOf) {
	const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count) {
		return;
	}
	const unsigned w = begin + i;
	const unsigned childBase = begin + count + innerScan[i], triBase = triTotal + leafScan[i];
	const unsigned inner = (unsigned)__popc(wide[w].innerMask), n = wide[w].count;
	wide[w].childGroup = childBase << 4;
	wide[w].recBase = triBase - inner;
	for (unsigned c = 0; c < n; ++c) {
		const int2 ref = slotRef[(size_t)w * 4 + c];
		if (c < inner) {
			wideBinary[childBase + c] = ref.x;
		} else {
			const unsigned rec = triBase + (c - inner);
			const unsigned tri = (unsigned)~ref.x;
			triOrder[rec] = tri;
			recordOf[tri] = rec;
			const float *box = reinterpret_cast<const float *>(nodes + (ref.y >> 1)) + (ref.y & 1) * 8; // (min.xyzw, max.xyzw) of that side
			float *o = leafBoxes + (size_t)rec * 6;
			o[0] = box[0]; o[1] = box[1]; o[2] = box[2];
			o[3] = box[4]; o[4] = box[5]; o[5] = box[6];
		}
	}
}

// the binary image names the same records (it is walked by the rays the wide walk does not take)
__global__ void wide_remap_image_kernel(float4 *__restrict__ image, unsigned nNodes, const unsigned *__restrict__ recordOf) {
	const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nNodes) {
		return;
	}
	int2 *ch = reinterpret_cast<int2 *>(image + (size_t)i * 4 + 3);
	int2 c = *ch;
	if (c.x < 0) c.x = ~(int)recordOf[(unsigned)~c.x];
	if (c.y < 0) c.y = ~(int)recordOf[(unsigned)~c.y];
	*ch = c;
}

} // namespace

size_t wide_build_scratch_bytes(unsigned nNodes, unsigned nTris) {
	const size_t n = (size_t)nNodes + 4;
	return n * sizeof(int) /* wideBinary */ + n * 4 * sizeof(int2) /* slotRef */ + 4 * (n + 2) * sizeof(unsigned) /* counts, scans */ +
	       (size_t)nTris * sizeof(unsigned) /* recordOf */ + (2 * (n / 1024 + 4) + 2 * (n / 1024 / 1024 + 4) + 8) * sizeof(unsigned) /* scan tmp */ + 1024;
}

// wide: room for nNodes nodes.  triOrder: nTris words, leafBoxes: 6 nTris floats.  image: the 64-byte binary image of the same
// tree (its leaf words are rewritten to name records).  Returns through nWide / depth; *usable = false when the tree is not
// walked wide (deeper than the walk's stack, a box off the grid, a triangle without a leaf).
cudaError_t build_wide_image_device(const restir_aabb_node *nodes, unsigned nNodes, unsigned nTris, const WideQuant &quant, WideNode *wide, unsigned *triOrder,
                                    float *leafBoxes, float4 *image, void *scratch, unsigned *nWide, int *depth, bool *usable, cudaStream_t s) {
	*usable = false;
	*nWide = 0;
	*depth = 0;
	unsigned char *p = static_cast<unsigned char *>(scratch);
	auto take = [&](size_t bytes) {
		void *r = p;
		p += (bytes + 255) & ~(size_t)255;
		return r;
	};
	const size_t n = (size_t)nNodes + 4;
	int *wideBinary = (int *)take(n * sizeof(int));
	int2 *slotRef = (int2 *)take(n * 4 * sizeof(int2));
	unsigned *innerCount = (unsigned *)take((n + 2) * 4), *leafCount = (unsigned *)take((n + 2) * 4);
	unsigned *innerScan = (unsigned *)take((n + 2) * 4), *leafScan = (unsigned *)take((n + 2) * 4);
	unsigned *recordOf = (unsigned *)take((size_t)nTris * 4);
	unsigned *tmp = (unsigned *)take((2 * (n / 1024 + 4) + 2 * (n / 1024 / 1024 + 4) + 8) * 4);
	unsigned *bad = (unsigned *)take(256);
	cudaError_t e;
	if ((e = cudaMemsetAsync(bad, 0, sizeof(unsigned), s)) != cudaSuccess) return e;
	if ((e = cudaMemsetAsync(wideBinary, 0, sizeof(int), s)) != cudaSuccess) return e; // wide node 0 stands for binary node 0
	unsigned begin = 0, count = 1, triTotal = 0;
	int levels = 0;
	while (count != 0) {
		if (begin + count > nNodes || ++levels > kWideStack) {
			return cudaSuccess; // not a tree of nNodes nodes / deeper than the walk's stack: the binary image is walked
		}
		const unsigned blocks = (count + 127) / 128;
		wide_expand_kernel<<<blocks, 128, 0, s>>>(nodes, wideBinary, begin, count, quant, wide, slotRef, innerCount, leafCount, bad);
		if ((e = exclusive_scan_u32(innerCount, innerScan, count, tmp, s)) != cudaSuccess) return e;
		if ((e = exclusive_scan_u32(leafCount, leafScan, count, tmp, s)) != cudaSuccess) return e;
		wide_link_kernel<<<blocks, 128, 0, s>>>(nodes, begin, count, innerScan, leafScan, triTotal, wide, slotRef, wideBinary, triOrder, leafBoxes, recordOf);
		unsigned totals[2];
		if ((e = cudaMemcpyAsync(&totals[0], innerScan + count, sizeof(unsigned), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
		if ((e = cudaMemcpyAsync(&totals[1], leafScan + count, sizeof(unsigned), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
		if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
		begin += count;
		count = totals[0];
		triTotal += totals[1];
		if (triTotal > nTris) {
			return cudaSuccess;
		}
	}
	unsigned nBad = 0;
	if ((e = cudaMemcpyAsync(&nBad, bad, sizeof(unsigned), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
	if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
	if (nBad != 0 || triTotal != nTris) {
		return cudaSuccess; // a box off the grid / a triangle no leaf names
	}
	wide_remap_image_kernel<<<(nNodes + 255) / 256, 256, 0, s>>>(image, nNodes, recordOf);
	if ((e = cudaGetLastError()) != cudaSuccess) return e;
	*nWide = begin;
	*depth = levels;
	*usable = true;
	return cudaSuccess;
}

cudaError_t preload_wide_build_kernels() {
	cudaFuncAttributes a;
	cudaError_t e = cudaFuncGetAttributes(&a, wide_expand_kernel);
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, wide_link_kernel);
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, wide_remap_image_kernel);
	return e;
}

} // namespace restir
