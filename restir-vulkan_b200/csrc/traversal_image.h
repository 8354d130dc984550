// traversal_image.h — what restir_upload_bvh derives from the uploaded AABB tree (host, once per upload).
//
// The interface layout stays the reference's (aabbTree.glsl:1-13: 80-byte 2-wide nodes, 48-byte triangles).
// The trace kernel walks a device-side copy of THE SAME tree — same boxes, same children, same order — whose
// nodes are re-strided to 64 bytes: an 80-byte node is five 16-byte loads and, at an 80-byte stride, crosses
// a 128-byte line in 3 nodes out of 8; a 64-byte node is four loads from one line.  Nothing about the
// traversal changes, so no visibility bit can.
//
// The upload is also checked, because the kernels trust it: every child index must be in range and no node
// may be reached twice (the reference would read out of bounds / loop), and the worst-case occupancy of the
// reference's 32-entry stack (softwareRaytracing.glsl:44) is computed.  If that bound exceeds 32 the tree is
// traversed by the literal 80-byte path that drops and counts pushes on a full stack, like the reference's
// undefined behaviour would; otherwise no push can ever be dropped and the stack needs no bound checks.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/restir_layouts.h"

namespace restir {

// 4 x float4, one per axis then the children: (Lmin.a, Rmin.a, Lmax.a, Rmax.a) for a = x, y, z; (left, right, -, -).
// The left and the right box sit side by side so that one packed instruction (sm_100 FADD2 / FMUL2: two
// IEEE binary32 operations per issue slot, each rounded like the scalar one) serves both.
struct alignas(64) Node64 {
	float box[12];
	int32_t left, right; // reference encoding: >= 0 node index, < 0 ~triangleIndex
	int32_t pad[2];
};
static_assert(sizeof(Node64) == 64, "traversal node is 64 bytes");

struct TraversalImageInfo {
	bool usable = false;         // false: keep the literal reference traversal (reason in `why`)
	std::string why;
	int referenceStackBound = 0; // worst-case occupancy of the reference's stack (all boxes hit)
	int anyOrderStackBound = 0;  // the same for a walk that may visit the two children of a node in either order
	int depth = 0;               // level of the deepest leaf (root node = level 0)
	uint32_t reachableNodes = 0;
};

// Returns false (with `error`) when the upload is not a tree over [0,nNodes) x [0,nTris): it is rejected.
bool build_traversal_image(const restir_aabb_node *nodes, uint32_t nNodes, uint32_t nTris, std::vector<Node64> &out, TraversalImageInfo &info,
                           std::string &error);

} // namespace restir
