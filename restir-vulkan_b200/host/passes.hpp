// passes.hpp — C++ host-side mirror of the reference's pass objects, over the C ABI.
//
// The reference drives the hot path through four pass classes with an `issueCommands` method and
// public fields that select the resources of a dispatch:
//     RestirPass         src/passes/restirPass.h:36-62      (fields :64-70)
//     SpatialReusePass   src/passes/spatialReusePass.h:11-26 (fields :91-93: descriptorSet, screenSize, iter)
//     UnbiasedReusePass  src/passes/unbiasedReusePass.h:23-49
//     LightingPass       src/passes/lightingPass.h:14-42
// and App wires them to its buffers (src/app.h:212-346).  The classes below keep those names, the
// `issueCommands` verb and the meaning of the fields; a Vulkan descriptor set becomes the plain ids
// of what it bound (G-buffer slot, reservoir buffer ids) and the command buffer becomes the context's
// CUDA stream.  Errors are thrown as restir::Error instead of aborting (src/misc.cpp:21-26).
//
// Header-only; link against librestir_b200.so.
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/restir_b200.h"

namespace restir {

struct Error : std::runtime_error {
	int code;
	Error(int c, const std::string &what) : std::runtime_error(what), code(c) {}
};

// Owns the restir_context: the device-side state App keeps (src/app.h:113-148).
class Device {
public:
	explicit Device(int cudaDevice = 0, void *stream = nullptr) {
		int rc = restir_create(&_ctx, cudaDevice, stream);
		if (rc != RESTIR_OK) {
			throw Error(rc, "restir_create failed: a CUDA device is required (no CPU fallback)");
		}
	}
	~Device() { restir_destroy(_ctx); }
	Device(const Device &) = delete;
	Device &operator=(const Device &) = delete;

	restir_context *get() const { return _ctx; }
	void check(int rc) const {
		if (rc != RESTIR_OK) {
			throw Error(rc, restir_last_error(_ctx));
		}
	}
	void waitIdle() const { check(restir_synchronize(_ctx)); } // vk::Device::waitIdle

private:
	restir_context *_ctx = nullptr;
};

// AabbTree (src/aabbTreeBuilder.h:11-17) + AabbTreeBuffers::create (:25-51)
struct AabbTree {
	std::vector<restir_aabb_node> nodes;
	std::vector<restir_triangle> triangles;

	// triangles: world space, in the reference's order (aabbTreeBuilder.cpp:58-76)
	static AabbTree build(std::vector<restir_triangle> worldTriangles) {
		AabbTree t;
		t.triangles = std::move(worldTriangles);
		if (t.triangles.size() < 2) {
			throw Error(RESTIR_E_INVALID, "AabbTree::build needs at least 2 triangles");
		}
		t.nodes.resize(t.triangles.size() - 1);
		int rc = restir_build_aabb_tree(t.triangles.data(), (uint32_t)t.triangles.size(), t.nodes.data());
		if (rc != RESTIR_OK) {
			throw Error(rc, "restir_build_aabb_tree failed");
		}
		return t;
	}
	void upload(const Device &dev) const {
		dev.check(restir_upload_bvh(dev.get(), nodes.data(), (uint32_t)nodes.size(), triangles.data(), (uint32_t)triangles.size()));
	}
};

// The light part of SceneBuffers (src/sceneBuffers.h:78-124, 241-270)
struct SceneLights {
	std::vector<restir_point_light> pointLights;
	std::vector<restir_tri_light> triangleLights;
	std::vector<restir_alias_column> aliasTable;

	// sceneBuffers.h:78-84: fall back to 200 random lights when the scene has none
	static SceneLights create(std::vector<restir_point_light> point, std::vector<restir_tri_light> tri, const float sceneMin[3],
	                          const float sceneMax[3]) {
		SceneLights s;
		s.pointLights = std::move(point);
		s.triangleLights = std::move(tri);
		if (s.pointLights.empty() && s.triangleLights.empty()) {
			s.pointLights.resize(200);
			restir_generate_random_point_lights(200, sceneMin, sceneMax, s.pointLights.data());
		}
		s.aliasTable.resize(s.pointLights.empty() ? s.triangleLights.size() : s.pointLights.size());
		int rc = restir_create_alias_table(s.pointLights.data(), s.pointLights.size(), s.triangleLights.data(), s.triangleLights.size(),
		                                   s.aliasTable.data());
		if (rc != RESTIR_OK) {
			throw Error(rc, "restir_create_alias_table failed");
		}
		return s;
	}
	template <typename T> static std::vector<unsigned char> blob(const std::vector<T> &items) {
		std::vector<unsigned char> b(RESTIR_BLOB_HEADER_BYTES + sizeof(T) * items.size(), 0);
		int32_t n = (int32_t)items.size();
		std::copy((unsigned char *)&n, (unsigned char *)&n + 4, b.begin());
		if (!items.empty()) {
			std::copy((const unsigned char *)items.data(), (const unsigned char *)items.data() + sizeof(T) * items.size(),
			          b.begin() + RESTIR_BLOB_HEADER_BYTES);
		}
		return b;
	}
	void upload(const Device &dev) const {
		auto p = blob(pointLights), t = blob(triangleLights), a = blob(aliasTable);
		dev.check(restir_upload_lights(dev.get(), p.data(), p.size(), t.data(), t.size(), a.data(), a.size()));
	}
};

// RestirPass (restirPass.h): frameDescriptorSet -> {gBuffer, outReservoirs, prevReservoirs}
class RestirPass {
public:
	int gBuffer = 0;                             // current G-buffer; the previous frame's is gBuffer ^ 1
	int reservoirBuffer = RESTIR_BUF_FRAME0;     // binding 8, set 1
	int prevFrameReservoirBuffer = RESTIR_BUF_FRAME1; // binding 9, set 1
	bool useSoftwareRayTracing = true;           // the hardware (VK_KHR_ray_tracing) path does not exist here

	void issueCommands(const Device &dev) const {
		if (!useSoftwareRayTracing) {
			throw Error(RESTIR_E_UNSUPPORTED, "hardware ray tracing is out of scope: B200 has no RT cores");
		}
		dev.check(restir_pass_restir(dev.get(), gBuffer, reservoirBuffer, prevFrameReservoirBuffer));
	}
};

// SpatialReusePass (spatialReusePass.h): descriptorSet -> {gBuffer, in, out}; push constant iter
class SpatialReusePass {
public:
	int gBuffer = 0;
	int reservoirBuffer = RESTIR_BUF_FRAME0;       // binding 6
	int resultReservoirBuffer = RESTIR_BUF_FRAME1; // binding 7
	int iter = 0;

	void issueCommands(const Device &dev) const {
		dev.check(restir_pass_spatial(dev.get(), gBuffer, reservoirBuffer, resultReservoirBuffer, iter));
	}
};

// UnbiasedReusePass (unbiasedReusePass.h)
class UnbiasedReusePass {
public:
	int gBuffer = 0;
	int reservoirBuffer = RESTIR_BUF_TEMP;         // binding 5
	int resultReservoirBuffer = RESTIR_BUF_FRAME0; // binding 6
	bool useSoftwareRayTracing = true;

	void issueCommands(const Device &dev) const {
		if (!useSoftwareRayTracing) {
			throw Error(RESTIR_E_UNSUPPORTED, "hardware ray tracing is out of scope: B200 has no RT cores");
		}
		dev.check(restir_pass_unbiased(dev.get(), gBuffer, reservoirBuffer, resultReservoirBuffer));
	}
};

// GBufferPass (gBufferPass.h / gBufferPass.cpp:116-157): scene / sceneBuffers / descriptorSets -> the geometry and materials
// uploaded once (restir_upload_geometry, restir_upload_materials); the uniform's projectionViewMatrix -> the camera it is made of
class GBufferPass {
public:
	int gBuffer = 0;
	restir_camera camera{};

	static void uploadScene(const Device &dev, const std::vector<restir_vertex> &vertices, const std::vector<uint32_t> &indices,
	                        const std::vector<restir_draw> &draws, const std::vector<restir_model_matrices> &matrices,
	                        const std::vector<restir_material_uniforms> &materials, const std::vector<restir_material_textures> &bindings,
	                        const std::vector<restir_texture> &textures) {
		dev.check(restir_upload_geometry(dev.get(), vertices.data(), vertices.size(), indices.data(), indices.size(), draws.data(), matrices.data(),
		                                 (uint32_t)draws.size()));
		dev.check(restir_upload_materials(dev.get(), materials.data(), bindings.data(), (uint32_t)materials.size(), textures.data(),
		                                  (uint32_t)textures.size()));
	}
	void issueCommands(const Device &dev) const { dev.check(restir_pass_gbuffer(dev.get(), gBuffer, &camera)); }
};

// LightingPass (lightingPass.h): descriptorSet -> {gBuffer, reservoirs}; the framebuffer becomes a device image
class LightingPass {
public:
	int gBuffer = 0;
	int reservoirBuffer = RESTIR_BUF_FRAME0;
	void *outImage = nullptr; // device memory
	int outFormat = RESTIR_OUT_RGBA8_SRGB;

	void issueCommands(const Device &dev) const {
		dev.check(restir_pass_lighting(dev.get(), gBuffer, reservoirBuffer, outImage, outFormat));
	}
};

// What App::_recordMainCommandBuffers + _updateRestirBuffers set up (src/app.h:212-336): the pass order
// and buffer roles for G-buffer index i.
class FrameRecorder {
public:
	bool unbiasedSpatialReuse = true; // src/app.h:174
	int spatialReuseIterations = 1;   // src/app.h:169

	void record(const Device &dev, int i) const {
		const int cur = i & 1, prev = cur ^ 1;
		RestirPass restir;
		restir.gBuffer = cur;
		restir.prevFrameReservoirBuffer = prev;
		if (unbiasedSpatialReuse) {
			restir.reservoirBuffer = RESTIR_BUF_TEMP;
			restir.issueCommands(dev);
			UnbiasedReusePass unbiased;
			unbiased.gBuffer = cur;
			unbiased.reservoirBuffer = RESTIR_BUF_TEMP;
			unbiased.resultReservoirBuffer = cur;
			unbiased.issueCommands(dev);
			return;
		}
		restir.reservoirBuffer = cur;
		restir.issueCommands(dev);
		for (int j = 0; j < spatialReuseIterations; ++j) {
			SpatialReusePass spatial;
			spatial.gBuffer = cur;
			spatial.reservoirBuffer = cur;
			spatial.resultReservoirBuffer = prev;
			spatial.iter = j * 2;
			spatial.issueCommands(dev);
			spatial.reservoirBuffer = prev;
			spatial.resultReservoirBuffer = cur;
			spatial.iter = j * 2 + 1;
			spatial.issueCommands(dev);
		}
	}
};

// Row bands in one process (no reference equivalent: the reference is single-GPU).  Each band is a Device of its own
// — on one GPU or on several with peer access enabled by the caller — allocated with restir_resize_band and wired to
// its neighbours with restir_band_connect, after which every band runs the same FrameRecorder and the library moves
// the halo rows itself (restir_b200.h, "row-band neighbours over peer memory").
class BandSet {
public:
	// bounds: bands + 1 ascending rows from 0 to height
	BandSet(const std::vector<Device *> &devices, uint32_t width, uint32_t height, const std::vector<uint32_t> &bounds, uint32_t halo)
	    : _devices(devices), _bounds(bounds) {
		if (devices.empty() || bounds.size() != devices.size() + 1 || bounds.front() != 0 || bounds.back() != height) {
			throw Error(RESTIR_E_INVALID, "BandSet: bounds must run from 0 to height, one band per device");
		}
		for (size_t r = 0; r < devices.size(); ++r) {
			// a neighbour pushes only the rows it shades itself: a halo taller than a band would leave the rows of the band
			// two hops away empty (restir_band_connect rejects it too)
			if (devices.size() > 1 && (bounds[r + 1] <= bounds[r] || bounds[r + 1] - bounds[r] < halo)) {
				throw Error(RESTIR_E_INVALID, "BandSet: band " + std::to_string(r) + " has " + std::to_string(bounds[r + 1] - bounds[r]) +
				                                  " rows, fewer than the " + std::to_string(halo) + "-row halo");
			}
		}
		for (size_t r = 0; r < devices.size(); ++r) {
			devices[r]->check(restir_resize_band(devices[r]->get(), width, height, bounds[r], bounds[r + 1], halo));
		}
		for (size_t r = 0; r < devices.size(); ++r) {
			restir_band_peer up{}, down{};
			if (r > 0) devices[r - 1]->check(restir_band_local_peer(devices[r - 1]->get(), &up));
			if (r + 1 < devices.size()) devices[r + 1]->check(restir_band_local_peer(devices[r + 1]->get(), &down));
			devices[r]->check(restir_band_connect(devices[r]->get(), 0, r > 0 ? &up : nullptr));
			devices[r]->check(restir_band_connect(devices[r]->get(), 1, r + 1 < devices.size() ? &down : nullptr));
		}
	}
	size_t size() const { return _devices.size(); }
	Device &device(size_t r) const { return *_devices[r]; }
	uint32_t rowBegin(size_t r) const { return _bounds[r]; }
	uint32_t rowEnd(size_t r) const { return _bounds[r + 1]; }
	// rows the band's per-pixel buffers cover (its G-buffer planes and images start at allocBegin)
	void allocRows(size_t r, uint32_t &begin, uint32_t &end) const {
		_devices[r]->check(restir_get_band(_devices[r]->get(), nullptr, nullptr, &begin, &end));
	}
	// Same pass sequence on every band, issued asynchronously: a band that waits for a neighbour's rows waits on the device.
	void record(const FrameRecorder &recorder, int i) const {
		for (Device *d : _devices) {
			recorder.record(*d, i);
		}
	}
	void waitIdle() const {
		for (Device *d : _devices) {
			d->waitIdle();
		}
	}

private:
	std::vector<Device *> _devices;
	std::vector<uint32_t> _bounds;
};

} // namespace restir
