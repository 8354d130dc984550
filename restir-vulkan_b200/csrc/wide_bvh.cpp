// wide_bvh.cpp — see wide_bvh.h.  Host C++, runs once per restir_upload_bvh.

#include "wide_bvh.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

namespace restir {

namespace {

struct Box {
	float lo[3], hi[3];
};

struct Slot {
	Box box;
	int32_t child; // reference encoding: >= 0 binary node index, < 0 ~triangle
};

inline Box leftBox(const restir_aabb_node &n) {
	Box b;
	for (int k = 0; k < 3; ++k) {
		b.lo[k] = n.leftAabbMin[k];
		b.hi[k] = n.leftAabbMax[k];
	}
	return b;
}
inline Box rightBox(const restir_aabb_node &n) {
	Box b;
	for (int k = 0; k < 3; ++k) {
		b.lo[k] = n.rightAabbMin[k];
		b.hi[k] = n.rightAabbMax[k];
	}
	return b;
}
inline bool wellFormed(const Box &b) {
	for (int k = 0; k < 3; ++k) {
		if (!std::isfinite(b.lo[k]) || !std::isfinite(b.hi[k]) || !(b.lo[k] <= b.hi[k])) {
			return false;
		}
	}
	return true;
}
inline bool inside(const Box &inner, const Box &outer) {
	for (int k = 0; k < 3; ++k) {
		if (!(inner.lo[k] >= outer.lo[k]) || !(inner.hi[k] <= outer.hi[k])) {
			return false;
		}
	}
	return true;
}
inline double halfArea(const Box &b) {
	double x = (double)b.hi[0] - b.lo[0], y = (double)b.hi[1] - b.lo[1], z = (double)b.hi[2] - b.lo[2];
	return x * y + x * z + y * z;
}

} // namespace

bool build_wide_bvh(const restir_aabb_node *nodes, uint32_t nNodes, uint32_t nTris, std::vector<WideNode> &out, WideBvhInfo &info,
                    std::string &error) {
	out.clear();
	info = WideBvhInfo{};
	char msg[256];

	// ---- 1. structure: every child index in range, every node reached at most once -------------------
	std::vector<uint8_t> seen(nNodes, 0);
	std::vector<int32_t> order; // pre-order of reachable nodes
	order.reserve(nNodes);
	{
		std::vector<int32_t> todo{0};
		seen[0] = 1;
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
					todo.push_back(ch[s]);
				} else if ((uint32_t)(~ch[s]) >= nTris) {
					std::snprintf(msg, sizeof(msg), "AABB tree node %d: triangle index %d out of range [0,%u)", n, ~ch[s], nTris);
					error = msg;
					return false;
				}
			}
		}
	}

	// ---- 2. worst-case occupancy of the reference's traversal stack (softwareRaytracing.glsl:44-67:
	// pop, push left then right, right is popped first), children before parents --------------------------
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
	if (info.referenceStackBound > 32) {
		std::snprintf(msg, sizeof(msg), "the reference's 32-entry stack can overflow on this tree (worst case %d entries)", info.referenceStackBound);
		info.why = msg;
		return true; // usable stays false: reference-order traversal, dropped pushes counted
	}

	// ---- 3. boxes well-formed; per inner node: do its children's boxes nest inside its own box? ----------
	std::vector<uint8_t> nested(nNodes, 0);
	for (int32_t idx : order) {
		const restir_aabb_node &n = nodes[idx];
		Box bl = leftBox(n), br = rightBox(n);
		if (!wellFormed(bl) || !wellFormed(br)) {
			std::snprintf(msg, sizeof(msg), "node %d has a non-finite or inverted box", idx);
			info.why = msg;
			return true;
		}
		if (n.leftChild >= 0) {
			const restir_aabb_node &c = nodes[n.leftChild];
			nested[n.leftChild] = inside(leftBox(c), bl) && inside(rightBox(c), bl);
		}
		if (n.rightChild >= 0) {
			const restir_aabb_node &c = nodes[n.rightChild];
			nested[n.rightChild] = inside(leftBox(c), br) && inside(rightBox(c), br);
		}
	}

	// ---- 4. fold: breadth-first, each wide node starts as a binary node's two slots and greedily replaces
	// its largest foldable inner slot by that node's two slots until it has four -------------------------
	struct Pending {
		int32_t binary;
		int depth;
	};
	std::vector<Pending> queue{{0, 1}};
	out.reserve(nNodes / 2 + 1);
	for (size_t head = 0; head < queue.size(); ++head) {
		Pending cur = queue[head];
		const restir_aabb_node &root = nodes[cur.binary];
		Slot slots[4];
		int count = 2;
		slots[0] = Slot{leftBox(root), root.leftChild};
		slots[1] = Slot{rightBox(root), root.rightChild};
		while (count < 4) {
			int best = -1;
			double bestArea = -1.0;
			for (int s = 0; s < count; ++s) {
				if (slots[s].child >= 0 && nested[slots[s].child]) {
					double a = halfArea(slots[s].box);
					if (a > bestArea) {
						bestArea = a;
						best = s;
					}
				}
			}
			if (best < 0) {
				break;
			}
			const restir_aabb_node &c = nodes[slots[best].child];
			slots[best] = Slot{leftBox(c), c.leftChild};
			slots[count++] = Slot{rightBox(c), c.rightChild};
			info.foldedNodes++;
		}
		WideNode w;
		std::memset(&w, 0, sizeof(w));
		for (int s = 0; s < 4; ++s) {
			if (s < count) {
				for (int k = 0; k < 3; ++k) {
					w.planes[k][s] = slots[s].box.lo[k];
					w.planes[3 + k][s] = slots[s].box.hi[k];
				}
				if (slots[s].child >= 0) {
					w.child[s] = (int32_t)queue.size(); // wide index = position in the BFS queue
					queue.push_back({slots[s].child, cur.depth + 1});
					if (!nested[slots[s].child]) {
						info.keptUnfolded++;
					}
				} else {
					w.child[s] = slots[s].child;
				}
			} else {
				// a box no segment with finite 1/dir can hit: t = (3e38 - o) * inv is either >= 1 (inv > 0) or
				// negative (inv < 0) on every axis
				for (int k = 0; k < 6; ++k) {
					w.planes[k][s] = 3.0e38f;
				}
				w.child[s] = kWideEmptyChild;
			}
		}
		info.wideDepth = std::max(info.wideDepth, cur.depth);
		out.push_back(w);
	}
	info.wideNodes = (uint32_t)out.size();

	// ---- 5. worst-case occupancy of the wide traversal's stack (first hit inner child is visited next,
	// the others are pushed in slot order and popped last-pushed-first) ------------------------------------
	std::vector<int32_t> wneed(out.size(), 0);
	for (size_t k = out.size(); k-- > 0;) {
		int inner[4], m = 0;
		for (int s = 0; s < 4; ++s) {
			if (out[k].child[s] >= 0) inner[m++] = out[k].child[s];
		}
		int v = 0;
		if (m > 0) {
			v = (m - 1) + wneed[inner[0]];
			for (int j = m - 1, pending = m - 2; j >= 1; --j, --pending) {
				v = std::max(v, pending + wneed[inner[j]]);
			}
		}
		wneed[k] = v;
	}
	info.wideStackBound = wneed.empty() ? 0 : wneed[0];
	info.usable = true;
	return true;
}

} // namespace restir
