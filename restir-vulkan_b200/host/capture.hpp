// capture.hpp — reader of frame captures (include/restir_capture.h) for the C++ host side.  Header-only, no CUDA.
#pragma once

#include <cstdint>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/restir_b200.h"
#include "../../include/restir_capture.h"

namespace restir {

struct CaptureFrame {
	restir_uniforms uniforms;
	restir_lighting_uniforms lightingUniforms;
	std::vector<unsigned char> plane[5]; // albedo, normal, material, worldPos, depth
	std::vector<restir_reservoir> initial, final; // empty when the capture carries no expected output
	std::vector<float> rgba;
};

struct Capture {
	restir_capture_header header{};
	std::vector<restir_aabb_node> nodes;
	std::vector<restir_triangle> triangles;
	std::vector<unsigned char> pointBlob, triBlob, aliasBlob;
	std::vector<CaptureFrame> frames;

	static Capture read(const std::string &path) {
		std::ifstream f(path, std::ios::binary);
		if (!f) {
			throw std::runtime_error("cannot open capture " + path);
		}
		Capture c;
		get(f, &c.header, sizeof(c.header), path);
		if (std::memcmp(c.header.magic, RESTIR_CAPTURE_MAGIC, 8) != 0 || c.header.version != 1) {
			throw std::runtime_error(path + ": not a RSTRCAP1 version-1 capture");
		}
		const restir_capture_header &h = c.header;
		c.nodes.resize(h.n_nodes);
		c.triangles.resize(h.n_triangles);
		c.pointBlob.resize(h.point_blob_bytes);
		c.triBlob.resize(h.tri_blob_bytes);
		c.aliasBlob.resize(h.alias_blob_bytes);
		get(f, c.nodes.data(), c.nodes.size() * sizeof(restir_aabb_node), path);
		get(f, c.triangles.data(), c.triangles.size() * sizeof(restir_triangle), path);
		get(f, c.pointBlob.data(), c.pointBlob.size(), path);
		get(f, c.triBlob.data(), c.triBlob.size(), path);
		get(f, c.aliasBlob.data(), c.aliasBlob.size(), path);
		const size_t n = (size_t)h.width * h.height;
		const size_t bpp[5] = {4, 8, 4, 16, 4};
		c.frames.resize(h.frames);
		for (CaptureFrame &fr : c.frames) {
			get(f, &fr.uniforms, sizeof(fr.uniforms), path);
			get(f, &fr.lightingUniforms, sizeof(fr.lightingUniforms), path);
			for (int k = 0; k < 5; ++k) {
				fr.plane[k].resize(n * bpp[k]);
				get(f, fr.plane[k].data(), fr.plane[k].size(), path);
			}
			if (h.expected & RESTIR_CAPTURE_HAS_INITIAL) {
				fr.initial.resize(n);
				get(f, fr.initial.data(), n * sizeof(restir_reservoir), path);
			}
			if (h.expected & RESTIR_CAPTURE_HAS_FINAL) {
				fr.final.resize(n);
				get(f, fr.final.data(), n * sizeof(restir_reservoir), path);
			}
			if (h.expected & RESTIR_CAPTURE_HAS_RGBA) {
				fr.rgba.resize(n * 4);
				get(f, fr.rgba.data(), n * 16, path);
			}
		}
		if (f.peek() != std::ifstream::traits_type::eof()) {
			throw std::runtime_error(path + ": trailing bytes");
		}
		return c;
	}

private:
	static void get(std::ifstream &f, void *dst, size_t bytes, const std::string &path) {
		if (bytes == 0) {
			return;
		}
		f.read(static_cast<char *>(dst), (std::streamsize)bytes);
		if ((size_t)f.gcount() != bytes) {
			throw std::runtime_error(path + ": truncated capture");
		}
	}
};

// Reservoirs that differ between two buffers, comparing what the passes define: the 52 bytes of LightSample + M of a
// pixel (restirStructs.glsl:19-34), bit for bit, any NaN equal to any NaN.
inline size_t countReservoirMismatches(const std::vector<restir_reservoir> &a, const std::vector<restir_reservoir> &b) {
	size_t bad = 0;
	for (size_t i = 0; i < a.size() && i < b.size(); ++i) {
		uint32_t x[16], y[16];
		std::memcpy(x, &a[i], 64);
		std::memcpy(y, &b[i], 64);
		bool same = true;
		for (int k = 0; k < 13 && same; ++k) { // 12 words of the sample + numStreamSamples; the tail is padding
			if (x[k] == y[k]) {
				continue;
			}
			const bool isFloat = k != 8 && k != 12; // word 8 = lightIndex, word 12 = M
			const bool nanX = (x[k] & 0x7f800000u) == 0x7f800000u && (x[k] & 0x7fffffu), nanY = (y[k] & 0x7f800000u) == 0x7f800000u && (y[k] & 0x7fffffu);
			same = isFloat && nanX && nanY;
		}
		bad += same ? 0 : 1;
	}
	return bad + (a.size() > b.size() ? a.size() - b.size() : b.size() - a.size());
}

} // namespace restir
