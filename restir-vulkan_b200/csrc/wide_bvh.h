// wide_bvh.h — device-side traversal image of the uploaded AABB tree (host builder, run once per upload).
//
// The interface layout stays the reference's (aabbTree.glsl:1-13: 80-byte 2-wide nodes, 48-byte
// triangles); restir_upload_bvh additionally derives a 4-wide, 128-byte-aligned re-layout of THE SAME
// tree (same boxes, same leaves, every second level folded into its parent) for the traversal kernel.
//
// Why this is exact, not approximate.  softwareRaytracing.glsl:39-85 is an any-hit search: it returns
// "occluded" iff some triangle T passes rayTriangleIntersection AND every box on the path root -> T passes
// rayAabIntersection.  A child box lies inside its parent's box (the builder takes unions of exact
// min/max, aabbTreeBuilder.cpp:112-119,180-196) and the slab test `(b - o) * inv` is monotone in b for
// finite inv, so "child box hit" implies "parent box hit": dropping the parent's test cannot change the
// answer.  build_wide_bvh CHECKS the nesting per node and only folds nodes where it holds, so the
// property is verified for the bytes actually uploaded, not assumed.  Rays whose 1/dir is not finite
// (0·inf = NaN breaks monotonicity) and trees the reference itself cannot traverse without overflowing
// its 32-entry stack take the reference-order traversal instead (restir_trace.cuh: trace_any_reference).
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/restir_layouts.h"

namespace restir {

// 8 x float4.  planes[axis + 3 * side][child]: side 0 = box min, side 1 = box max.  child[c] >= 0: wide
// node index; < 0: ~triangleIndex.  Empty slots hold a far-away degenerate box that no segment hits.
struct alignas(128) WideNode {
	float planes[6][4];
	int32_t child[4];
	int32_t pad[4];
};
static_assert(sizeof(WideNode) == 128, "wide node is one 128-byte line");

constexpr int32_t kWideEmptyChild = INT32_MIN;

struct WideBvhInfo {
	bool usable = false;      // false: keep the reference-order traversal (reason in `why`)
	std::string why;
	int referenceStackBound = 0; // worst-case occupancy of the reference's stack (all boxes hit)
	int wideStackBound = 0;      // same for the 4-wide traversal order
	uint32_t wideNodes = 0;
	uint32_t foldedNodes = 0;    // binary nodes whose own box test was folded away
	uint32_t keptUnfolded = 0;   // inner children left unfolded because nesting did not hold
	int wideDepth = 0;
};

// Returns false (with `error`) when the tree is not a tree over [0,nNodes) x [0,nTris) — such an upload
// is rejected: the reference would read out of bounds.  Otherwise fills `out`/`info`.
bool build_wide_bvh(const restir_aabb_node *nodes, uint32_t nNodes, uint32_t nTris, std::vector<WideNode> &out, WideBvhInfo &info,
                    std::string &error);

} // namespace restir
