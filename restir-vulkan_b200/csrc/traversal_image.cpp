// traversal_image.cpp — see traversal_image.h.  Host C++, runs once per restir_upload_bvh.

#include "traversal_image.h"

#include <algorithm>
#include <cstdio>
#include <cstring>

namespace restir {

bool build_traversal_image(const restir_aabb_node *nodes, uint32_t nNodes, uint32_t nTris, std::vector<Node64> &out, TraversalImageInfo &info,
                           std::string &error) {
	out.clear();
	info = TraversalImageInfo{};
	char msg[256];

	// ---- structure: every child index in range, every node reached at most once ----------------------
	std::vector<uint8_t> seen(nNodes, 0);
	std::vector<int32_t> order; // pre-order of the reachable nodes
	std::vector<int32_t> depthOf(nNodes, 0);
	order.reserve(nNodes);
	{
		std::vector<int32_t> todo{0};
		seen[0] = 1;
		depthOf[0] = 0; // root = level 0, as SURVEY.md Appendix D counts it
		while (!todo.empty()) {
			int32_t n = todo.back();
			todo.pop_back();
			order.push_back(n);
			const int32_t ch[2] = {nodes[n].leftChild, nodes[n].rightChild};
			for (int s = 0; s < 2; ++s) {
				if (ch[s] >= 0) {
					if ((uint32_t)ch[s] >= nNodes) {
						std::snprintf(msg, sizeof(msg), "AABB tree node %d: child node index %d out of range [0,%u)", n, ch[s], nNodes);
						error = msg;
						return false;
					}
					if (seen[ch[s]]) {
						std::snprintf(msg, sizeof(msg), "AABB tree node %d is referenced more than once (not a tree)", ch[s]);
						error = msg;
						return false;
					}
					seen[ch[s]] = 1;
					depthOf[ch[s]] = depthOf[n] + 1;
					todo.push_back(ch[s]);
				} else if ((uint32_t)(~ch[s]) >= nTris) {
					std::snprintf(msg, sizeof(msg), "AABB tree node %d: triangle index %d out of range [0,%u)", n, ~ch[s], nTris);
					error = msg;
					return false;
				}
				info.depth = std::max(info.depth, depthOf[n] + 1);
			}
		}
	}
	info.reachableNodes = (uint32_t)order.size();

	// ---- worst-case occupancy of the reference's traversal stack (softwareRaytracing.glsl:44-67: pop, push
	// left then right, so right is popped first), children before parents --------------------------------------
	std::vector<int32_t> need(nNodes, 0);
	for (size_t k = order.size(); k-- > 0;) {
		const restir_aabb_node &n = nodes[order[k]];
		int li = n.leftChild >= 0, ri = n.rightChild >= 0;
		int v = li + ri;
		if (ri) v = std::max(v, li + need[n.rightChild]);
		if (li) v = std::max(v, need[n.leftChild]);
		need[order[k]] = v;
	}
	info.referenceStackBound = std::max(1, need[0]);
	// a walk that picks the order per visit (restir_trace.cuh, near child first) holds at most one pending sibling
	// per node on the path to the current one
	std::vector<int32_t> needAny(nNodes, 0);
	for (size_t k = order.size(); k-- > 0;) {
		const restir_aabb_node &n = nodes[order[k]];
		int li = n.leftChild >= 0, ri = n.rightChild >= 0;
		int v = 0;
		if (ri) v = std::max(v, li + needAny[n.rightChild]);
		if (li) v = std::max(v, ri + needAny[n.leftChild]);
		needAny[order[k]] = v;
	}
	info.anyOrderStackBound = std::max(1, needAny[0]);
	if (info.referenceStackBound > 32 || info.anyOrderStackBound > 32) {
		std::snprintf(msg, sizeof(msg), "the reference's 32-entry stack can overflow on this tree (worst case %d entries, %d in any visiting order)", info.referenceStackBound,
		              info.anyOrderStackBound);
		info.why = msg;
		return true; // usable stays false
	}

	// ---- the 64-byte image: same index, same boxes, same children, planes grouped per axis (traversal_image.h) -------------------------------------
	out.resize(nNodes);
	for (uint32_t i = 0; i < nNodes; ++i) {
		const restir_aabb_node &n = nodes[i];
		Node64 &o = out[i];
		for (int k = 0; k < 3; ++k) {
			o.box[4 * k + 0] = n.leftAabbMin[k];
			o.box[4 * k + 1] = n.rightAabbMin[k];
			o.box[4 * k + 2] = n.leftAabbMax[k];
			o.box[4 * k + 3] = n.rightAabbMax[k];
		}
		o.left = n.leftChild;
		o.right = n.rightChild;
		o.pad[0] = o.pad[1] = 0;
	}
	info.usable = true;
	return true;
}

} // namespace restir
