// capture_info — prints what a frame capture (include/restir_capture.h) holds: header fields and an FNV-1a checksum per
// section.  CPU-only (g++ -std=c++17 capture_info.cpp); the tests use it to hold the C++ reader to the Python writer.
#include <cstdio>
#include <iostream>

#include "capture.hpp"

static unsigned long long fnv1a(const void *data, size_t bytes, unsigned long long h = 1469598103934665603ull) {
	const unsigned char *p = static_cast<const unsigned char *>(data);
	for (size_t i = 0; i < bytes; ++i) {
		h = (h ^ p[i]) * 1099511628211ull;
	}
	return h;
}

int main(int argc, char **argv) {
	if (argc != 2) {
		std::cerr << "usage: capture_info <capture.rsc>\n";
		return 64;
	}
	try {
		restir::Capture c = restir::Capture::read(argv[1]);
		const restir_capture_header &h = c.header;
		unsigned long long scene = fnv1a(c.nodes.data(), c.nodes.size() * sizeof(restir_aabb_node));
		scene = fnv1a(c.triangles.data(), c.triangles.size() * sizeof(restir_triangle), scene);
		scene = fnv1a(c.pointBlob.data(), c.pointBlob.size(), scene);
		scene = fnv1a(c.triBlob.data(), c.triBlob.size(), scene);
		scene = fnv1a(c.aliasBlob.data(), c.aliasBlob.size(), scene);
		unsigned long long inputs = 1469598103934665603ull, expected = 1469598103934665603ull;
		for (const restir::CaptureFrame &f : c.frames) {
			inputs = fnv1a(&f.uniforms, sizeof(f.uniforms), inputs);
			inputs = fnv1a(&f.lightingUniforms, sizeof(f.lightingUniforms), inputs);
			for (int k = 0; k < 5; ++k) {
				inputs = fnv1a(f.plane[k].data(), f.plane[k].size(), inputs);
			}
			expected = fnv1a(f.initial.data(), f.initial.size() * sizeof(restir_reservoir), expected);
			expected = fnv1a(f.final.data(), f.final.size() * sizeof(restir_reservoir), expected);
			expected = fnv1a(f.rgba.data(), f.rgba.size() * sizeof(float), expected);
		}
		std::printf("{\"width\": %u, \"height\": %u, \"frames\": %u, \"unbiased\": %u, \"unbiased_neighbors\": %u, \"spatial_iterations\": %u, "
		            "\"expected\": %u, \"n_nodes\": %u, \"n_triangles\": %u, \"scene_fnv1a\": \"%016llx\", \"inputs_fnv1a\": \"%016llx\", "
		            "\"expected_fnv1a\": \"%016llx\"}\n",
		            h.width, h.height, h.frames, h.unbiased, h.unbiased_neighbors, h.spatial_iterations, h.expected, h.n_nodes, h.n_triangles, scene,
		            inputs, expected);
	} catch (const std::exception &e) {
		std::cerr << "capture_info: " << e.what() << "\n";
		return 1;
	}
	return 0;
}
