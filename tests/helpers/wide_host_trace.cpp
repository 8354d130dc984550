// TEST INFRASTRUCTURE (not product code): a scalar host restatement of the 4-wide traversal that
// restir-vulkan_b200/csrc/restir_trace.cuh runs on the GPU, over the wide nodes the product's own
// build_wide_bvh produces.  tests/test_wide_bvh.py compares it with the oracle's reference-order traversal
// on CPU, so the "folding a level does not change any visibility bit" argument of wide_bvh.h is checked
// without a GPU.  Built by the test with g++ -ffp-contract=off.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../restir-vulkan_b200/csrc/wide_bvh.h"

using restir::WideNode;

namespace {
struct V3 {
	float x, y, z;
};
inline V3 sub(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return V3{a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }

bool tri_hit(const float *t, V3 o, V3 d) {
	V3 p1{t[0], t[1], t[2]}, e1 = sub(V3{t[4], t[5], t[6]}, p1), e2 = sub(V3{t[8], t[9], t[10]}, p1);
	V3 p = cross(d, e2);
	float f = 1.0f / dot(e1, p);
	V3 s = sub(o, p1);
	float u = f * dot(s, p);
	if (u < 0.0f || u > 1.0f) return false;
	V3 q = cross(s, e1);
	float v = f * dot(d, q);
	if (v < 0.0f || v + u > 1.0f) return false;
	f = f * dot(e2, q);
	return f > 0.0f && f < 1.0f;
}
} // namespace

extern "C" {

// returns 0 ok, 1 tree rejected, 2 tree not usable for the wide traversal; info[0..5] = wide nodes, depth, folded,
// unfolded, reference stack bound, wide stack bound
int wide_host_build(const void *nodes, uint32_t nNodes, uint32_t nTris, void **handle, uint32_t *info, char *why, int whyLen) {
	auto *w = new std::vector<WideNode>();
	restir::WideBvhInfo wi;
	std::string err;
	if (!restir::build_wide_bvh(static_cast<const restir_aabb_node *>(nodes), nNodes, nTris, *w, wi, err)) {
		std::strncpy(why, err.c_str(), whyLen - 1);
		delete w;
		return 1;
	}
	info[0] = wi.wideNodes; info[1] = (uint32_t)wi.wideDepth; info[2] = wi.foldedNodes; info[3] = wi.keptUnfolded;
	info[4] = (uint32_t)wi.referenceStackBound; info[5] = (uint32_t)wi.wideStackBound;
	if (!wi.usable) {
		std::strncpy(why, wi.why.c_str(), whyLen - 1);
		delete w;
		return 2;
	}
	*handle = w;
	return 0;
}
void wide_host_free(void *handle) { delete static_cast<std::vector<WideNode> *>(handle); }

// shadowed[i] = 1 hit, 0 miss, 2 = "1/dir not finite: the product traces this ray in reference order"
void wide_host_trace(void *handle, const float *tris, int64_t n, const float *p1, const float *p2, uint8_t *shadowed, int32_t *maxStack) {
	const std::vector<WideNode> &wide = *static_cast<std::vector<WideNode> *>(handle);
	int deepest = 0;
#pragma omp parallel for schedule(dynamic, 1024) reduction(max : deepest)
	for (int64_t i = 0; i < n; ++i) {
		V3 a{p1[i * 3], p1[i * 3 + 1], p1[i * 3 + 2]}, b{p2[i * 3], p2[i * 3 + 1], p2[i * 3 + 2]};
		V3 dir = sub(b, a);
		float invLen = 1.0f / std::sqrt(dot(dir, dir));
		V3 off{(dir.x * invLen) * 0.001f, (dir.y * invLen) * 0.001f, (dir.z * invLen) * 0.001f};
		V3 o{a.x + off.x, a.y + off.y, a.z + off.z};
		V3 d{dir.x - off.x * 2.0f, dir.y - off.y * 2.0f, dir.z - off.z * 2.0f};
		float inv[3] = {1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
		if (!(std::fabs(inv[0]) < INFINITY && std::fabs(inv[1]) < INFINITY && std::fabs(inv[2]) < INFINITY)) {
			shadowed[i] = 2;
			continue;
		}
		const float oo[3] = {o.x, o.y, o.z};
		int nearIdx[3], farIdx[3];
		for (int k = 0; k < 3; ++k) {
			nearIdx[k] = inv[k] < 0.0f ? 3 + k : k;
			farIdx[k] = inv[k] < 0.0f ? k : 3 + k;
		}
		std::vector<int32_t> stack;
		int32_t cur = 0;
		uint8_t result = 0;
		for (;;) {
			const WideNode &w = wide[cur];
			int32_t inner[4];
			int m = 0;
			bool hitTri = false;
			for (int c = 0; c < 4 && !hitTri; ++c) {
				float tn = -INFINITY, tf = INFINITY;
				for (int k = 0; k < 3; ++k) {
					tn = std::fmax(tn, (w.planes[nearIdx[k]][c] - oo[k]) * inv[k]);
					tf = std::fmin(tf, (w.planes[farIdx[k]][c] - oo[k]) * inv[k]);
				}
				if (!(tn < 1.0f && tf >= tn && tf > 0.0f)) continue;
				if (w.child[c] < 0) {
					hitTri = tri_hit(tris + (size_t)(~w.child[c]) * 12, o, d);
				} else {
					inner[m++] = w.child[c];
				}
			}
			if (hitTri) {
				result = 1;
				break;
			}
			if (m == 0) {
				if (stack.empty()) break;
				cur = stack.back();
				stack.pop_back();
				continue;
			}
			cur = inner[0];
			for (int j = 1; j < m; ++j) stack.push_back(inner[j]);
			deepest = std::max(deepest, (int)stack.size());
		}
		shadowed[i] = result;
	}
	*maxStack = deepest;
}
}
