// restir_driver — C++ host program that drives the sm_100a kernels through the C ABI the way the
// reference's App drives its passes: load scene data, build tree / lights / alias table, then per frame
// patch the uniforms (src/app.cpp:775-826), record the ReSTIR passes (src/app.h:212-262) and the
// lighting pass (src/app.cpp:861-862), flipping the G-buffer index (src/app.cpp:900).
//
//   restir_driver <scene_dir> <width> <height> <frames> <unbiased 0|1> [neighbors] [--bands N] [--halo ROWS]
//
// --bands N (no reference equivalent): the frame is split into N row bands, each a context of its own on GPU 0, wired
// to its neighbours with restir::BandSet (restir_band_connect): the library exchanges the halo rows itself and the
// checksums must equal the single-context run's.
//
// <scene_dir> holds triangles.bin (n x 48), tri_material.i32, materials.f32 (m x 16), dims.f32 and
// material_table.u32 (m x 4) — what scenes/_baked/<name>/ or restir-vulkan_b200/fixtures.py write.
//   restir_driver --replay <capture.rsc>
//
// --replay (no reference equivalent): runs a frame capture (include/restir_capture.h — scene buffers, uniforms and
// G-buffers of every frame, e.g. dumped from a run of the Vulkan application) through the passes and counts what
// differs from the outputs the capture carries.
//
// Prints one JSON line: FNV-1a checksums of the final reservoirs (64-byte layout) and of the RGBA8 image,
// ms per frame and ray count — tests compare the checksums with the same sequence driven from Python.

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "capture.hpp"
#include "passes.hpp"

namespace {

template <typename T> std::vector<T> readFile(const std::string &path, bool required = true) {
	std::ifstream f(path, std::ios::binary | std::ios::ate);
	if (!f) {
		if (required) {
			std::cerr << "restir_driver: cannot read " << path << "\n";
			std::exit(2);
		}
		return {};
	}
	std::streamsize bytes = f.tellg();
	f.seekg(0);
	std::vector<T> v((size_t)bytes / sizeof(T));
	f.read(reinterpret_cast<char *>(v.data()), (std::streamsize)(v.size() * sizeof(T)));
	return v;
}

uint64_t fnv1a(const void *data, size_t bytes, uint64_t h = 1469598103934665603ull) {
	const unsigned char *p = static_cast<const unsigned char *>(data);
	for (size_t i = 0; i < bytes; ++i) {
		h = (h ^ p[i]) * 1099511628211ull;
	}
	return h;
}

void cuda(cudaError_t e, const char *what) {
	if (e != cudaSuccess) {
		std::cerr << "restir_driver: " << what << ": " << cudaGetErrorString(e) << "\n";
		std::exit(3);
	}
}

struct DeviceGBuffer {
	void *plane[5] = {};
	void allocate(size_t pixels) {
		const size_t bpp[5] = {4, 8, 4, 16, 4};
		for (int k = 0; k < 5; ++k) {
			cuda(cudaMalloc(&plane[k], pixels * bpp[k]), "cudaMalloc G-buffer plane");
		}
	}
	restir_gbuffer_planes planes() const { return restir_gbuffer_planes{plane[0], plane[1], plane[2], plane[3], plane[4]}; }
};

// --replay: the capture's frames through the passes, compared with the outputs it carries.
int replay(const std::string &path) {
	restir::Capture cap = restir::Capture::read(path);
	const restir_capture_header &h = cap.header;
	const size_t pixels = (size_t)h.width * h.height;
	restir::Device device(0);
	device.check(restir_upload_bvh(device.get(), cap.nodes.data(), h.n_nodes, cap.triangles.data(), h.n_triangles));
	device.check(restir_upload_lights(device.get(), cap.pointBlob.data(), cap.pointBlob.size(), cap.triBlob.data(), cap.triBlob.size(),
	                                  cap.aliasBlob.data(), cap.aliasBlob.size()));
	device.check(restir_resize(device.get(), h.width, h.height));
	device.check(restir_set_unbiased_neighbors(device.get(), h.unbiased_neighbors ? h.unbiased_neighbors : 3));
	float *image = nullptr;
	cuda(cudaMalloc(&image, pixels * 16), "cudaMalloc image");
	std::vector<restir_reservoir> got(pixels);
	std::vector<float> rgba(pixels * 4);
	size_t badInitial = 0, badFinal = 0, badPixels = 0;
	double maxRel = 0.0;
	for (uint32_t f = 0; f < h.frames; ++f) {
		const restir::CaptureFrame &fr = cap.frames[f];
		const int i = (int)(f & 1u), p = i ^ 1;
		restir_gbuffer_planes planes{fr.plane[0].data(), fr.plane[1].data(), fr.plane[2].data(), fr.plane[3].data(), fr.plane[4].data()};
		device.check(restir_upload_gbuffer(device.get(), i, RESTIR_GBUFFER_NVIDIA_DEFAULT, &planes));
		device.check(restir_set_uniforms(device.get(), &fr.uniforms));
		device.check(restir_set_lighting_uniforms(device.get(), &fr.lightingUniforms));
		const int first = h.unbiased ? RESTIR_BUF_TEMP : i; // src/app.h:298-332
		device.check(restir_pass_restir(device.get(), i, first, p));
		if (!fr.initial.empty()) {
			device.check(restir_download_reservoirs(device.get(), first, got.data()));
			badInitial += restir::countReservoirMismatches(got, fr.initial);
		}
		if (h.unbiased) {
			device.check(restir_pass_unbiased(device.get(), i, RESTIR_BUF_TEMP, i));
		} else {
			for (uint32_t j = 0; j < h.spatial_iterations; ++j) {
				device.check(restir_pass_spatial(device.get(), i, i, p, (int)(2 * j)));
				device.check(restir_pass_spatial(device.get(), i, p, i, (int)(2 * j + 1)));
			}
		}
		if (!fr.final.empty()) {
			device.check(restir_download_reservoirs(device.get(), i, got.data()));
			badFinal += restir::countReservoirMismatches(got, fr.final);
		}
		device.check(restir_pass_lighting(device.get(), i, i, image, RESTIR_OUT_RGBA32F));
		device.waitIdle();
		if (!fr.rgba.empty()) {
			cuda(cudaMemcpy(rgba.data(), image, pixels * 16, cudaMemcpyDeviceToHost), "memcpy image");
			for (size_t k = 0; k < pixels; ++k) {
				bool bad = false;
				for (int c = 0; c < 3; ++c) { // tolerance of tests/parity_harness.py: rel 1e-3 with an absolute floor of 1e-6
					const double a = rgba[k * 4 + c], b = fr.rgba[k * 4 + c];
					if (a != a && b != b) continue;
					const double diff = a > b ? a - b : b - a, den = (b < 0 ? -b : b) > 1e-6 ? (b < 0 ? -b : b) : 1e-6;
					if (!(diff <= 1e-6)) {
						maxRel = diff / den > maxRel ? diff / den : maxRel;
						bad = bad || !(diff / den <= 1e-3);
					}
				}
				badPixels += bad ? 1 : 0;
			}
		}
	}
	restir_counters counters{};
	device.check(restir_get_counters(device.get(), &counters, 0));
	std::printf("{\"replay\": \"%s\", \"frames\": %u, \"width\": %u, \"height\": %u, \"unbiased\": %u, \"expected\": %u, "
	            "\"mismatching_initial_reservoirs\": %zu, \"mismatching_final_reservoirs\": %zu, \"mismatching_pixels\": %zu, "
	            "\"max_rel_rgb_error\": %.3g, \"shadow_rays\": %llu}\n",
	            path.c_str(), h.frames, h.width, h.height, h.unbiased, h.expected, badInitial, badFinal, badPixels, maxRel,
	            (unsigned long long)counters.shadow_rays);
	cudaFree(image);
	return (badInitial || badFinal || badPixels) ? 5 : 0;
}

} // namespace

int main(int argc, char **argv) {
	if (argc == 3 && std::string(argv[1]) == "--replay") {
		try {
			return replay(argv[2]);
		} catch (const restir::Error &e) {
			std::cerr << "restir_driver: error " << e.code << ": " << e.what() << "\n";
			return 1;
		} catch (const std::exception &e) {
			std::cerr << "restir_driver: " << e.what() << "\n";
			return 2;
		}
	}
	if (argc < 6) {
		std::cerr << "usage: restir_driver <scene_dir> <width> <height> <frames> <unbiased 0|1> [neighbors] [--bands N] [--halo ROWS]\n";
		return 64;
	}
	const std::string dir = argv[1];
	const uint32_t width = (uint32_t)std::atoi(argv[2]), height = (uint32_t)std::atoi(argv[3]);
	const int frames = std::atoi(argv[4]);
	const bool unbiased = std::atoi(argv[5]) != 0;
	uint32_t neighbors = unbiased ? 3u : 4u, nBands = 1, halo = 64;
	for (int a = 6; a < argc; ++a) {
		const std::string arg = argv[a];
		if (arg == "--bands" && a + 1 < argc) {
			nBands = (uint32_t)std::atoi(argv[++a]);
		} else if (arg == "--halo" && a + 1 < argc) {
			halo = (uint32_t)std::atoi(argv[++a]);
		} else {
			neighbors = (uint32_t)std::atoi(argv[a]);
		}
	}
	if (nBands < 1 || nBands > 64) {
		std::cerr << "restir_driver: --bands must be in 1..64\n";
		return 64;
	}

	try {
		// ---- scene (App::App: loadScene, SceneBuffers::create, AabbTree::build; src/app.cpp:285-360) ----
		auto triangles = readFile<restir_triangle>(dir + "/triangles.bin");
		auto triMaterial = readFile<int32_t>(dir + "/tri_material.i32");
		auto materials = readFile<float>(dir + "/materials.f32");
		auto dims = readFile<float>(dir + "/dims.f32");
		auto materialTable = readFile<uint32_t>(dir + "/material_table.u32");
		auto filePointLights = readFile<unsigned char>(dir + "/gltf_point_lights.bin", false);
		const uint32_t nMaterials = (uint32_t)(materials.size() / 16);
		std::vector<float> emissive(nMaterials * 3);
		for (uint32_t m = 0; m < nMaterials; ++m) {
			std::memcpy(&emissive[m * 3], &materials[m * 16 + 8], 12);
		}
		std::vector<restir_point_light> point;
		if (filePointLights.size() > RESTIR_BLOB_HEADER_BYTES) {
			point.resize((filePointLights.size() - RESTIR_BLOB_HEADER_BYTES) / sizeof(restir_point_light));
			std::memcpy(point.data(), filePointLights.data() + RESTIR_BLOB_HEADER_BYTES, point.size() * sizeof(restir_point_light));
		}
		std::vector<restir_tri_light> tri(triangles.size());
		int64_t nTri = restir_collect_triangle_lights(triangles.data(), triMaterial.data(), (uint32_t)triangles.size(), emissive.data(),
		                                              nMaterials, tri.data());
		if (nTri < 0) {
			throw restir::Error((int)nTri, "restir_collect_triangle_lights failed");
		}
		tri.resize((size_t)nTri);

		// one context per band (a single one covering the screen without --bands), the scene replicated into each
		restir::AabbTree tree = restir::AabbTree::build(triangles);
		restir::SceneLights lights = restir::SceneLights::create(point, tri, dims.data(), dims.data() + 3);
		std::vector<std::unique_ptr<restir::Device>> owned;
		std::vector<restir::Device *> devices;
		for (uint32_t r = 0; r < nBands; ++r) {
			owned.emplace_back(new restir::Device(0));
			devices.push_back(owned.back().get());
			tree.upload(*devices[r]);
			lights.upload(*devices[r]);
		}
		std::vector<uint32_t> bounds(nBands + 1);
		for (uint32_t r = 0; r <= nBands; ++r) {
			bounds[r] = (uint32_t)((uint64_t)height * r / nBands);
		}
		std::unique_ptr<restir::BandSet> bandSet;
		if (nBands == 1) {
			devices[0]->check(restir_resize(devices[0]->get(), width, height)); // App::_updateRestirBuffers
		} else {
			bandSet.reset(new restir::BandSet(devices, width, height, bounds, halo));
		}
		for (restir::Device *d : devices) {
			d->check(restir_set_unbiased_neighbors(d->get(), unbiased ? neighbors : 3));
		}

		// ---- inputs: two G-buffers from two camera positions (fixture tool), per band its allocated rows -------
		int32_t *dTriMaterial = nullptr;
		uint32_t *dMaterialTable = nullptr;
		cuda(cudaMalloc(&dTriMaterial, triMaterial.size() * 4), "cudaMalloc");
		cuda(cudaMalloc(&dMaterialTable, materialTable.size() * 4), "cudaMalloc");
		cuda(cudaMemcpy(dTriMaterial, triMaterial.data(), triMaterial.size() * 4, cudaMemcpyHostToDevice), "memcpy");
		cuda(cudaMemcpy(dMaterialTable, materialTable.data(), materialTable.size() * 4, cudaMemcpyHostToDevice), "memcpy");
		restir_camera cams[2];
		for (int k = 0; k < 2; ++k) { // src/camera.h:7-13 defaults, nudged along +x for the second view
			restir_camera c{};
			auto cam = readFile<float>(dir + "/camera.f32", false); // optional: position xyz, lookAt xyz
			float pos[3] = {3.0f, 4.0f, 5.0f}, look[3] = {0.0f, 0.0f, 0.0f};
			if (cam.size() >= 6) {
				std::memcpy(pos, cam.data(), 12);
				std::memcpy(look, cam.data() + 3, 12);
			}
			c.position[0] = pos[0] + 0.05f * (float)k;
			c.position[1] = pos[1];
			c.position[2] = pos[2];
			std::memcpy(c.lookAt, look, 12);
			c.worldUp[1] = 1.0f;
			c.zNear = 0.01f;
			c.zFar = 1000.0f;
			c.fovYRadians = 0.5f * 3.14159265358979f;
			c.aspectRatio = (float)width / (float)height;
			cams[k] = c;
		}
		std::vector<uint32_t> allocBegin(nBands), allocEnd(nBands);
		std::vector<void *> images(nBands, nullptr);
		std::vector<DeviceGBuffer> gbuf(2 * (size_t)nBands);
		for (uint32_t r = 0; r < nBands; ++r) {
			devices[r]->check(restir_get_band(devices[r]->get(), nullptr, nullptr, &allocBegin[r], &allocEnd[r]));
			const size_t pixels = (size_t)(allocEnd[r] - allocBegin[r]) * width;
			for (int k = 0; k < 2; ++k) {
				DeviceGBuffer &g = gbuf[2 * r + k];
				g.allocate(pixels);
				devices[r]->check(restir_tools_raycast_gbuffer(devices[r]->get(), &cams[k], dTriMaterial, dMaterialTable, g.plane[0], g.plane[1],
				                                               g.plane[2], g.plane[3], g.plane[4]));
				restir_gbuffer_planes p = g.planes();
				devices[r]->check(restir_bind_gbuffer(devices[r]->get(), k, RESTIR_GBUFFER_NVIDIA_DEFAULT, &p));
			}
			cuda(cudaMalloc(&images[r], pixels * 4), "cudaMalloc image");
		}

		// ---- main loop (App::mainLoop, src/app.cpp:690-902) ---------------------------------------------
		restir_uniforms u{};
		u.screenSize[0] = width;
		u.screenSize[1] = height;
		u.frame = 0;                          // app.cpp:421
		u.spatialPosThreshold = 0.1f;         // app.h:150
		u.spatialNormalThreshold = 25.0f;     // app.h:151
		u.flags = RESTIR_VISIBILITY_REUSE_FLAG | RESTIR_TEMPORAL_REUSE_FLAG;
		u.spatialNeighbors = unbiased ? 4u : neighbors; // app.cpp:430
		u.spatialRadius = 30.0f;              // app.cpp:431
		restir_lighting_uniforms lu{};
		lu.bufferSize[0] = width;
		lu.bufferSize[1] = height;
		lu.debugMode = 0;
		lu.gamma = 1.0f;

		restir::FrameRecorder recorder;
		recorder.unbiasedSpatialReuse = unbiased;
		int currentGBufferFrame = 0;
		restir_counters counters{};
		for (restir::Device *d : devices) {
			d->check(restir_get_counters(d->get(), &counters, 1));
			d->waitIdle();
		}
		auto t0 = std::chrono::steady_clock::now();
		for (int f = 0; f < frames; ++f) {
			const restir_camera &cam = cams[f & 1], &prevCam = cams[f > 0 ? ((f & 1) ^ 1) : 0];
			++u.frame;                                   // app.cpp:776
			u.initialLightSampleCount = 1u << 5;         // app.cpp:777, app.h:165
			restir_camera_matrix(&prevCam, u.prevFrameProjectionViewMatrix); // app.cpp:778
			u.temporalSampleCountMultiplier = 20;        // app.cpp:779, app.h:168
			std::memcpy(u.cameraPos, cam.position, 12);  // app.cpp:795
			u.cameraPos[3] = 1.0f;
			std::memcpy(lu.cameraPos, u.cameraPos, 16);
			std::memcpy(lu.prevFrameProjectionViewMatrix, u.prevFrameProjectionViewMatrix, 64);
			for (restir::Device *d : devices) {
				d->check(restir_set_uniforms(d->get(), &u));
				d->check(restir_set_lighting_uniforms(d->get(), &lu));
			}
			if (bandSet) {
				bandSet->record(recorder, currentGBufferFrame);
			} else {
				recorder.record(*devices[0], currentGBufferFrame); // app.cpp:828-832
			}
			for (uint32_t r = 0; r < nBands; ++r) {
				restir::LightingPass lighting;
				lighting.outImage = images[r];
				lighting.gBuffer = currentGBufferFrame;
				lighting.reservoirBuffer = currentGBufferFrame;
				lighting.issueCommands(*devices[r]);          // app.cpp:861-862
			}
			currentGBufferFrame ^= 1;                     // app.cpp:900
		}
		for (restir::Device *d : devices) {
			d->waitIdle();
		}
		double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();

		// the whole frame, assembled from the rows each band owns
		const size_t pixels = (size_t)width * height;
		std::vector<restir_reservoir> reservoirs(pixels);
		std::vector<unsigned char> rgba(pixels * 4);
		unsigned long long rays = 0, launches = 0, haloMisses = 0, haloTimeouts = 0;
		for (uint32_t r = 0; r < nBands; ++r) {
			devices[r]->check(restir_get_counters(devices[r]->get(), &counters, 0));
			rays += counters.shadow_rays;
			launches += counters.kernel_launches;
			haloMisses += counters.halo_misses;
			haloTimeouts += counters.halo_wait_timeouts;
			const size_t bandPixels = (size_t)(allocEnd[r] - allocBegin[r]) * width;
			std::vector<restir_reservoir> part(bandPixels);
			devices[r]->check(restir_download_reservoirs(devices[r]->get(), currentGBufferFrame ^ 1, part.data()));
			std::vector<unsigned char> img(bandPixels * 4);
			cuda(cudaMemcpy(img.data(), images[r], img.size(), cudaMemcpyDeviceToHost), "memcpy image");
			const size_t first = (size_t)(bounds[r] - allocBegin[r]) * width, count = (size_t)(bounds[r + 1] - bounds[r]) * width;
			std::copy(part.begin() + first, part.begin() + first + count, reservoirs.begin() + (size_t)bounds[r] * width);
			std::copy(img.begin() + first * 4, img.begin() + (first + count) * 4, rgba.begin() + (size_t)bounds[r] * width * 4);
		}
		std::printf("{\"frames\": %d, \"bands\": %u, \"ms_per_frame_wall\": %.4f, \"shadow_rays\": %llu, \"kernel_launches\": %llu, "
		            "\"halo_misses\": %llu, \"halo_wait_timeouts\": %llu, \"reservoir_fnv1a\": \"%016llx\", \"image_fnv1a\": \"%016llx\"}\n",
		            frames, nBands, ms / frames, rays, launches, haloMisses, haloTimeouts,
		            (unsigned long long)fnv1a(reservoirs.data(), reservoirs.size() * sizeof(restir_reservoir)),
		            (unsigned long long)fnv1a(rgba.data(), rgba.size()));
		if (haloMisses != 0 || haloTimeouts != 0) { // the frame is not the single-GPU frame: never a success
			std::cerr << "restir_driver: " << haloMisses << " halo misses, " << haloTimeouts << " halo wait timeouts\n";
			return 2;
		}
	} catch (const restir::Error &e) {
		std::cerr << "restir_driver: error " << e.code << ": " << e.what() << "\n";
		return 1;
	}
	return 0;
}
