// Vulkan-free stand-in for the reference's src/aabbTreeBuilder.h, used ONLY by the recipe in
// oracle/ref_build/Makefile: the reference's own aabbTreeBuilder.cpp is copied next to this file
// into the git-ignored oracle/_ref/gen/ at build time, so that its `#include "aabbTreeBuilder.h"`
// resolves here.  It declares the same `AabbTree` (reference src/aabbTreeBuilder.h:11-17) and
// leaves out `AabbTreeBuffers` (the VMA upload, src/aabbTreeBuilder.h:19-52), which needs Vulkan.
#pragma once

#include <cassert>
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

#include <gltfscene.h>

#include "shaderIncludes.h"

struct AabbTree {
	std::vector<shader::AabbTreeNode> nodes;
	std::vector<shader::Triangle> triangles;
	int32_t root;

	[[nodiscard]] static AabbTree build(const nvh::GltfScene&);
};
