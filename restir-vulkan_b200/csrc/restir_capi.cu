// restir_capi.cu — the context behind include/restir_b200.h: device memory, the reference's buffer
// roles, and one in-order CUDA stream.  No CPU fallback: every pass is a kernel launch or an error.

#include <cmath>
#include <cstdarg>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/restir_b200.h"
#include "restir_kernels.h"
#include "traversal_image.h"
#include "wide_image.h"

using namespace restir;

struct restir_context {
	int device = 0;
	cudaStream_t stream = nullptr;
	bool ownStream = false;
	std::string error;
	bool cudaFailed = false;

	// scene
	float4 *nodes = nullptr, *tris = nullptr;
	float4 *image = nullptr; // 64-byte image of `nodes` (traversal_image.h); null => literal 80-byte walk
	float4 *triEdges = nullptr; // 64-byte (p1, e1, e2) records of `tris` (restir_trace.cuh)
	const uint32_t *wideOrder = nullptr; // inside treeBlock: record r holds triangle wideOrder[r]
	uint4 *wide = nullptr;      // 4-wide quantised image of the same tree (wide_image.h); null => the binary image is walked
	WideGrid wideGrid{};
	WideImageInfo wideInfo;
	unsigned char *treeBlock = nullptr; // one allocation holding `image`, `triEdges`, then `wide`: what the trace kernel walks
	size_t treeBlockBytes = 0;
	TraversalImageInfo imageInfo;
	uint32_t nNodes = 0, nTris = 0;
	int smCount = 0;
	void *bvhScratch = nullptr; // restir_build_bvh_device: kept between builds (a rebuild per frame must not pay for cudaMalloc)
	size_t bvhScratchBytes = 0;
	unsigned char *pointBlob = nullptr, *triBlob = nullptr, *aliasBlob = nullptr;
	float4 *pointPosLum = nullptr, *triAux = nullptr;
	int pointCount = 0, triCount = 0, aliasCount = 0;
	float *srgbLut = nullptr;
	// the G-buffer pass's scene (restir_upload_geometry / restir_upload_materials)
	float4 *gbAttrs = nullptr;
	int *gbTriMaterial = nullptr;
	uint32_t gbTris = 0;
	restir_material_uniforms *gbUniforms = nullptr;
	restir_material_textures *gbBindings = nullptr;
	uchar4 *gbTexels = nullptr;
	uint4 *gbTextureTable = nullptr;
	float *gbSrgbThresholds = nullptr;
	int gbMaterials = 0, gbTextures = 0;

	// screen
	Band band{0, 0, 0, 0, 0, 0};
	PackedReservoir *reservoirs[3] = {nullptr, nullptr, nullptr};
	// restir_set_reservoir_variant: RESERVOIR_SIZE, UNBIASED_MIS, one kernel per shader.  `generic` = anything but the shipped
	// configuration on the tuned path; its three buffers hold the reference's std430 records instead of packed ones.
	int variantN = 1;
	bool variantMis = false, fusedPasses = false;
	unsigned char *genericReservoirs[3] = {nullptr, nullptr, nullptr};
	bool generic() const { return variantN != 1 || variantMis || fusedPasses; }
	size_t reservoirBytes() const { return generic() ? generic_reservoir_bytes(variantN, variantMis) : sizeof(restir_reservoir); }
	GBufferView gbuf[2] = {};
	void *ownedPlanes[2][5] = {};
	// restir_upload_gbuffer copies on its own stream, so the upload of frame f+1 runs under the reuse and
	// lighting passes of frame f: `uploaded[s]` orders the passes after the copy into slot s, `lastRead[s]`
	// orders the next copy into slot s after the last kernel that read it.
	cudaStream_t copyStream = nullptr;
	cudaEvent_t uploaded[2] = {nullptr, nullptr}, lastRead[2] = {nullptr, nullptr};
	bool uploadPending[2] = {false, false}, readRecorded[2] = {false, false};
	restir_reservoir *staging = nullptr; // device scratch for 64-byte <-> 32-byte conversion
	size_t stagingPixels = 0;
	// hand-over buffers between the halves of a cut pass and the trace kernel (restir_kernels.cu)
	unsigned char *shadowed = nullptr; // [tile-ordered pixel id][rays per pixel]
	size_t shadowedBytes = 0;
	int *neighborPix = nullptr;        // [tile-ordered pixel id][unbiased neighbours]
	size_t neighborPixCount = 0;
	uint32_t *neighborM = nullptr;     // [tile-ordered pixel id][unbiased neighbours + 1]: sample counts for the normalisation
	size_t neighborMCount = 0;

	// connected row-band neighbours (restir_band_connect): side 0 owns the rows above, side 1 the rows below
	struct PeerSide {
		bool connected = false;
		PackedReservoir *reservoirs[3] = {nullptr, nullptr, nullptr};
		unsigned long long *flags = nullptr;
		int allocBegin = 0, allocEnd = 0, rowBegin = 0, rowEnd = 0;
	} peers[2];
	unsigned long long *bandFlags = nullptr; // device, [side the push came from][buffer]: raised by the neighbours
	unsigned *haloTicket = nullptr;
	uint64_t produced[3] = {0, 0, 0};        // times each buffer has been produced since the neighbours were connected
	std::vector<void *> ipcOpened;
	std::vector<std::pair<void *, cudaExternalMemory_t>> imported; // restir_import_external_memory

	restir_uniforms uniforms{};
	bool haveUniforms = false;
	restir_lighting_uniforms lighting{};
	bool haveLighting = false;
	uint32_t unbiasedNeighbors = 3; // unbiasedReuse.glsl:48
	int traversal = RESTIR_TRAVERSAL_AUTO;
	int rayElision = 1; // restir_set_ray_elision
	bool spatialStaging = false; // restir_set_spatial_staging
	unsigned long long *counters = nullptr; // device, kCounterCount entries
	unsigned *dedupe = nullptr;             // device, segment table of the neighbour rays (restir_trace.cu segment_claim), dedupeEntries words
	size_t dedupeEntries = 0;
	uint2 *aliases = nullptr;               // device, aliasCapacity pairs
	size_t aliasCapacity = 0;
	unsigned *aliasCounter = nullptr;       // device counter
	unsigned *occluders = nullptr;          // device, occluder cache of the trace kernel (restir_trace.cu): [region][256] entries
	size_t occluderEntries = 0;
	unsigned regionsX = 0;
	int occluderCache = 1;
	unsigned *traceCursors = nullptr;       // device, kTraceMaxRegions + kTraceMaxSms words (restir_trace.cu, RESTIR_TRACE_AFFINE)
	uint64_t launches = 0;

	// optional per-kernel CUDA-event timing (restir_profile_begin / _end)
	struct ProfRecord {
		const char *name;
		cudaEvent_t begin, end;
	};
	bool profiling = false;
	std::vector<ProfRecord> prof;

	size_t allocPixels() const { return (size_t)(band.allocEnd - band.allocBegin) * (size_t)band.W; }
};

namespace {

int fail(restir_context *ctx, int code, const char *fmt, ...) {
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	if (ctx) {
		ctx->error = buf;
	}
	return code;
}

int cudaCheck(restir_context *ctx, cudaError_t e, const char *what) {
	if (e == cudaSuccess) {
		return RESTIR_OK;
	}
	ctx->cudaFailed = true;
	return fail(ctx, e == cudaErrorMemoryAllocation ? RESTIR_E_NOMEM : RESTIR_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

#define CU(ctx, call)                                              \
	do {                                                           \
		int rc_ = cudaCheck((ctx), (call), #call);                 \
		if (rc_ != RESTIR_OK) return rc_;                          \
	} while (0)

#define ENTER(ctx)                                                                            \
	do {                                                                                      \
		if ((ctx) == nullptr) return RESTIR_E_INVALID;                                        \
		if ((ctx)->cudaFailed) return RESTIR_E_CUDA; /* sticky */                             \
		cudaError_t e_ = cudaSetDevice((ctx)->device);                                        \
		if (e_ != cudaSuccess) return cudaCheck((ctx), e_, "cudaSetDevice");                  \
	} while (0)

template <typename T> void freeDev(T *&p) {
	if (p) {
		cudaFree(p);
		p = nullptr;
	}
}

void dropGBuffers(restir_context *ctx) {
	for (int s = 0; s < 2; ++s) {
		for (int k = 0; k < 5; ++k) {
			freeDev(ctx->ownedPlanes[s][k]);
		}
		ctx->gbuf[s] = GBufferView{};
	}
}

SceneView sceneView(const restir_context *ctx) {
	SceneView v{};
	v.nodes = ctx->nodes;
	v.tris = ctx->tris;
	v.image = ctx->image;
	v.triEdges = ctx->triEdges;
	v.pointLights = ctx->pointBlob ? reinterpret_cast<const restir_point_light *>(ctx->pointBlob + RESTIR_BLOB_HEADER_BYTES) : nullptr;
	v.triLights = ctx->triBlob ? reinterpret_cast<const restir_tri_light *>(ctx->triBlob + RESTIR_BLOB_HEADER_BYTES) : nullptr;
	v.alias = ctx->aliasBlob ? reinterpret_cast<const restir_alias_column *>(ctx->aliasBlob + RESTIR_BLOB_HEADER_BYTES) : nullptr;
	v.pointPosLum = ctx->pointPosLum;
	v.triAux = ctx->triAux;
	v.srgbLut = ctx->srgbLut;
	v.pointCount = ctx->pointCount;
	v.triCount = ctx->triCount;
	v.aliasCount = ctx->aliasCount;
	v.nNodes = ctx->nNodes;
	v.nTris = ctx->nTris;
	return v;
}

int checkBuffer(restir_context *ctx, int b) {
	if (b < 0 || b > 2 || (ctx->generic() ? ctx->genericReservoirs[b] == nullptr : ctx->reservoirs[b] == nullptr)) {
		return fail(ctx, RESTIR_E_INVALID, "reservoir buffer id %d invalid or restir_resize not called", b);
	}
	return RESTIR_OK;
}

int makeParams(restir_context *ctx, int gbuffer, bool needScene, bool needLights, PassParams &p) {
	if (ctx->band.W == 0) {
		return fail(ctx, RESTIR_E_INVALID, "restir_resize has not been called");
	}
	if (gbuffer < 0 || gbuffer > 1 || ctx->gbuf[gbuffer].worldPos == nullptr) {
		return fail(ctx, RESTIR_E_INVALID, "G-buffer slot %d is not bound", gbuffer);
	}
	if (needScene && ctx->nodes == nullptr) {
		return fail(ctx, RESTIR_E_INVALID, "restir_upload_bvh has not been called");
	}
	if (needLights && ctx->aliasBlob == nullptr) {
		return fail(ctx, RESTIR_E_INVALID, "restir_upload_lights has not been called");
	}
	if (!ctx->haveUniforms) {
		return fail(ctx, RESTIR_E_INVALID, "restir_set_uniforms has not been called");
	}
	if ((int)ctx->uniforms.screenSize[0] != ctx->band.W || (int)ctx->uniforms.screenSize[1] != ctx->band.H) {
		return fail(ctx, RESTIR_E_INVALID, "uniforms.screenSize %ux%u does not match restir_resize %dx%d", ctx->uniforms.screenSize[0],
		            ctx->uniforms.screenSize[1], ctx->band.W, ctx->band.H);
	}
	p.scene = sceneView(ctx);
	p.cur = ctx->gbuf[gbuffer];
	p.prev = ctx->gbuf[gbuffer ^ 1];
	p.band = ctx->band;
	p.u = ctx->uniforms;
	p.counters = ctx->counters;
	return RESTIR_OK;
}

// Every kernel launch of a pass goes between beforeLaunch and afterLaunch: error check, launch count and,
// when profiling is on, a pair of events on the launching stream.
void beforeLaunch(restir_context *ctx, const char *what) {
	if (ctx->profiling) {
		restir_context::ProfRecord r{what, nullptr, nullptr};
		if (cudaEventCreate(&r.begin) == cudaSuccess && cudaEventCreate(&r.end) == cudaSuccess) {
			cudaEventRecord(r.begin, ctx->stream);
			ctx->prof.push_back(r);
		}
	}
}
int afterLaunch(restir_context *ctx, const char *what) {
	ctx->launches++;
	if (ctx->profiling && !ctx->prof.empty() && ctx->prof.back().name == what) {
		cudaEventRecord(ctx->prof.back().end, ctx->stream);
	}
	return cudaCheck(ctx, cudaGetLastError(), what);
}
void dropProfile(restir_context *ctx) {
	for (auto &r : ctx->prof) {
		if (r.begin) cudaEventDestroy(r.begin);
		if (r.end) cudaEventDestroy(r.end);
	}
	ctx->prof.clear();
}

const size_t kPlaneBytes[5] = {4, 8, 4, 16, 4}; // albedo, normal, material, worldPos, depth

// The passes about to read G-buffer slot `slot` wait for an upload into it that is still in flight.
int waitUpload(restir_context *ctx, int slot) {
	if (ctx->uploadPending[slot]) {
		CU(ctx, cudaStreamWaitEvent(ctx->stream, ctx->uploaded[slot], 0));
		ctx->uploadPending[slot] = false;
	}
	return RESTIR_OK;
}
// Marks the point on the compute stream after which slot `slot` may be overwritten by the next upload
// (only slots whose planes restir_upload_gbuffer owns take part).
int markRead(restir_context *ctx, int slot);
int markReadIfOwned(restir_context *ctx, int slot) {
	if (ctx->ownedPlanes[slot][3] != nullptr && ctx->gbuf[slot].worldPos == ctx->ownedPlanes[slot][3]) {
		return markRead(ctx, slot);
	}
	return RESTIR_OK;
}
int markRead(restir_context *ctx, int slot) {
	if (ctx->lastRead[slot] == nullptr) {
		CU(ctx, cudaEventCreateWithFlags(&ctx->lastRead[slot], cudaEventDisableTiming));
	}
	CU(ctx, cudaEventRecord(ctx->lastRead[slot], ctx->stream));
	ctx->readRecorded[slot] = true;
	return RESTIR_OK;
}

// grow-only hand-over buffers: one visibility byte per ray, one neighbour index per unbiased neighbour slot
int ensureHandOver(restir_context *ctx, const PassGrid &g, unsigned raysPerPixel, unsigned neighbors) {
	size_t bytes = (size_t)g.pixelIds * raysPerPixel;
	if (ctx->shadowedBytes < bytes) {
		CU(ctx, cudaStreamSynchronize(ctx->stream));
		freeDev(ctx->shadowed);
		ctx->shadowedBytes = 0;
		CU(ctx, cudaMalloc(&ctx->shadowed, bytes));
		ctx->shadowedBytes = bytes;
	}
	size_t count = (size_t)g.pixelIds * neighbors;
	if (ctx->neighborPixCount < count) {
		CU(ctx, cudaStreamSynchronize(ctx->stream));
		freeDev(ctx->neighborPix);
		ctx->neighborPixCount = 0;
		CU(ctx, cudaMalloc(&ctx->neighborPix, count * sizeof(int)));
		ctx->neighborPixCount = count;
	}
	if (neighbors != 0 && ctx->rayElision == 2) { // experiment: one walk per distinct neighbour segment: table of 2^n >= items words, room for items / 2 aliases
		size_t entries = 1024;
		while (entries < count && entries < ((size_t)1 << 28)) {
			entries <<= 1;
		}
		if (ctx->dedupeEntries < entries || ctx->aliasCapacity < count / 2) {
			CU(ctx, cudaStreamSynchronize(ctx->stream));
			freeDev(ctx->dedupe);
			freeDev(ctx->aliases);
			ctx->dedupeEntries = ctx->aliasCapacity = 0;
			CU(ctx, cudaMalloc(&ctx->dedupe, entries * sizeof(unsigned)));
			CU(ctx, cudaMalloc(&ctx->aliases, (count / 2 + 1) * sizeof(uint2)));
			ctx->dedupeEntries = entries;
			ctx->aliasCapacity = count / 2;
		}
		if (ctx->aliasCounter == nullptr) {
			CU(ctx, cudaMalloc(&ctx->aliasCounter, sizeof(unsigned)));
		}
	}
	size_t countM = neighbors ? (size_t)g.pixelIds * (neighbors + 1) : 0;
	if (ctx->neighborMCount < countM) {
		CU(ctx, cudaStreamSynchronize(ctx->stream));
		freeDev(ctx->neighborM);
		ctx->neighborMCount = 0;
		CU(ctx, cudaMalloc(&ctx->neighborM, countM * sizeof(uint32_t)));
		ctx->neighborMCount = countM;
	}
	return RESTIR_OK;
}

bool bandConnected(const restir_context *ctx) { return ctx->peers[0].connected || ctx->peers[1].connected; }

void dropPeers(restir_context *ctx) {
	for (auto &p : ctx->peers) {
		p = restir_context::PeerSide{};
	}
	for (void *m : ctx->ipcOpened) {
		cudaIpcCloseMemHandle(m);
	}
	ctx->ipcOpened.clear();
	ctx->produced[0] = ctx->produced[1] = ctx->produced[2] = 0;
	if (ctx->bandFlags) {
		cudaMemsetAsync(ctx->bandFlags, 0, 6 * sizeof(unsigned long long), ctx->stream);
	}
	cudaGetLastError();
}

// The pass about to read the halo rows of `buffer` waits (on the device) until both neighbours have pushed their latest rows.
int haloWait(restir_context *ctx, int buffer) {
	if (!bandConnected(ctx)) {
		return RESTIR_OK;
	}
	const unsigned long long *a = ctx->peers[0].connected ? ctx->bandFlags + 0 * 3 + buffer : nullptr;
	const unsigned long long *b = ctx->peers[1].connected ? ctx->bandFlags + 1 * 3 + buffer : nullptr;
	beforeLaunch(ctx, "halo_wait_kernel");
	CU(ctx, launch_halo_wait(a, b, ctx->produced[buffer], ctx->counters, ctx->stream));
	return afterLaunch(ctx, "halo_wait_kernel");
}

// `buffer` has just been produced: store the rows each neighbour holds as halo into its copy and raise its counter.
int haloPush(restir_context *ctx, int buffer) {
	if (!bandConnected(ctx)) {
		return RESTIR_OK;
	}
	ctx->produced[buffer]++;
	HaloPush hp{};
	hp.local = ctx->reservoirs[buffer];
	hp.W = ctx->band.W;
	hp.localAllocBegin = ctx->band.allocBegin;
	hp.sequence = ctx->produced[buffer];
	hp.ticket = ctx->haloTicket;
	for (int side = 0; side < 2; ++side) {
		const auto &p = ctx->peers[side];
		if (!p.connected) {
			continue;
		}
		hp.peer[side] = p.reservoirs[buffer];
		hp.peerAllocBegin[side] = p.allocBegin;
		// this side's neighbour sees me on its other side
		hp.peerFlag[side] = p.flags + (size_t)(side ^ 1) * 3 + buffer;
		if (side == 0) { // rows above are the neighbour's: it holds my first rows up to its allocEnd
			hp.firstRow[side] = ctx->band.rowBegin;
			hp.rows[side] = std::max(0, std::min(p.allocEnd, ctx->band.rowEnd) - ctx->band.rowBegin);
		} else {         // it holds my last rows from its allocBegin on
			hp.firstRow[side] = std::max(p.allocBegin, ctx->band.rowBegin);
			hp.rows[side] = std::max(0, ctx->band.rowEnd - hp.firstRow[side]);
		}
	}
	beforeLaunch(ctx, "halo_push_kernel");
	CU(ctx, launch_halo_push(hp, ctx->smCount, ctx->stream));
	return afterLaunch(ctx, "halo_push_kernel");
}

int clearOccluders(restir_context *ctx) {
	if (ctx->occluders != nullptr) {
		CU(ctx, cudaMemsetAsync(ctx->occluders, 0xff, ctx->occluderEntries * sizeof(unsigned), ctx->stream));
	}
	return RESTIR_OK;
}

TraceParams traceParams(const restir_context *ctx) {
	TraceParams tp{};
	tp.nodes = ctx->nodes;
	tp.tris = ctx->tris;
	tp.image = ctx->image;
	tp.triEdges = ctx->triEdges;
	tp.wide = ctx->wide;
	tp.grid = ctx->wideGrid;
	tp.nNodes = ctx->nNodes;
	tp.nTris = ctx->nTris;
	tp.occluders = (ctx->occluderCache && ctx->wide != nullptr && ctx->nTris < (1u << 24)) ? ctx->occluders : nullptr;
	tp.regionsX = ctx->regionsX;
	tp.occluderPretest = kOccluderWays > 1 ? 2 : 1; // restirOmni's rays try both witnesses of an entry (86 % of them are shadowed); the unbiased pass's
	                                                 // mostly unshadowed rays only the most recent one (restir_pass_unbiased)
	// 1: by light index while a region's 256 entries (x 256 tags) can tell the lights apart, by direction beyond; 2 / 3 force one
	tp.occluderByDirection = ctx->occluderCache == 3 || (ctx->occluderCache == 1 && ctx->pointCount + ctx->triCount > 4096);
	tp.band = ctx->band;
	tp.shadowed = ctx->shadowed;
	tp.counters = ctx->counters;
	tp.regionCursors = ctx->traceCursors;
	tp.smSlots = ctx->traceCursors ? ctx->traceCursors + kTraceMaxRegions : nullptr;
	return tp;
}

} // namespace

// RESTIR_NEIGHBOUR_PRETEST 0 (experiment): the neighbour rays skip the occluder cache's pretest (81 % of them are unshadowed)
#ifndef RESTIR_NEIGHBOUR_PRETEST
#define RESTIR_NEIGHBOUR_PRETEST 1
#endif

namespace {
template <typename T> int uploadArray(restir_context *ctx, T *&dst, const T *src, size_t n, const char *what) {
	freeDev(dst);
	if (n == 0) {
		return RESTIR_OK;
	}
	int rc = cudaCheck(ctx, cudaMalloc(&dst, n * sizeof(T)), what);
	if (rc != RESTIR_OK) return rc;
	return cudaCheck(ctx, cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream), what);
}


struct H3 {
	float x, y, z;
};
inline H3 hsub(H3 a, H3 b) { return H3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float hdot(H3 a, H3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline H3 hcross(H3 a, H3 b) { return H3{a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline H3 hnorm(H3 v) {
	float inv = 1.0f / sqrtf(hdot(v, v));
	return H3{v.x * inv, v.y * inv, v.z * inv};
}
void cameraBasis(const restir_camera *c, H3 &pos, H3 &fwd, H3 &right, H3 &up) { // camera.h:25-28
	pos = H3{c->position[0], c->position[1], c->position[2]};
	fwd = hnorm(hsub(H3{c->lookAt[0], c->lookAt[1], c->lookAt[2]}, pos));
	right = hnorm(hcross(fwd, H3{c->worldUp[0], c->worldUp[1], c->worldUp[2]}));
	up = hcross(right, fwd);
}

void cameraBasisOf(const restir_camera *camera, RaycastCamera &rc) {
	H3 pos, fwd, right, up;
	cameraBasis(camera, pos, fwd, right, up);
	float f = 1.0f / tanf(0.5f * camera->fovYRadians);
	rc.pos[0] = pos.x; rc.pos[1] = pos.y; rc.pos[2] = pos.z;
	rc.fwd[0] = fwd.x; rc.fwd[1] = fwd.y; rc.fwd[2] = fwd.z;
	rc.right[0] = right.x; rc.right[1] = right.y; rc.right[2] = right.z;
	rc.up[0] = up.x; rc.up[1] = up.y; rc.up[2] = up.z;
	rc.sx = camera->aspectRatio / f;
	rc.sy = 1.0f / f;
	restir_camera_matrix(camera, rc.pv);
}
} // namespace

extern "C" {

int restir_create(restir_context **out, int device, void *stream) {
	if (out == nullptr) {
		return RESTIR_E_INVALID;
	}
	*out = nullptr;
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || device < 0 || device >= count) {
		return RESTIR_E_CUDA; // no context to carry a message: there is no CPU fallback
	}
	restir_context *ctx = new (std::nothrow) restir_context();
	if (ctx == nullptr) {
		return RESTIR_E_NOMEM;
	}
	ctx->device = device;
	int rc = RESTIR_OK;
	do {
		if ((rc = cudaCheck(ctx, cudaSetDevice(device), "cudaSetDevice")) != RESTIR_OK) break;
		if ((rc = cudaCheck(ctx, cudaDeviceGetAttribute(&ctx->smCount, cudaDevAttrMultiProcessorCount, device), "cudaDeviceGetAttribute")) != RESTIR_OK) break;
		if (stream != nullptr) {
			ctx->stream = static_cast<cudaStream_t>(stream);
		} else {
			if ((rc = cudaCheck(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking), "cudaStreamCreate")) != RESTIR_OK) break;
			ctx->ownStream = true;
		}
		if ((rc = cudaCheck(ctx, cudaMalloc(&ctx->counters, sizeof(unsigned long long) * kCounterCount), "cudaMalloc counters")) != RESTIR_OK) break;
		if ((rc = cudaCheck(ctx, cudaMemsetAsync(ctx->counters, 0, sizeof(unsigned long long) * kCounterCount, ctx->stream), "memset")) != RESTIR_OK) break;
		if ((rc = cudaCheck(ctx, cudaMalloc(&ctx->traceCursors, sizeof(unsigned) * (kTraceMaxRegions + kTraceMaxSms)), "cudaMalloc trace cursors")) != RESTIR_OK) break;
		if ((rc = cudaCheck(ctx, preload_pixel_kernels(), "loading kernels")) != RESTIR_OK) break;
		if ((rc = cudaCheck(ctx, preload_trace_kernels(), "loading kernels")) != RESTIR_OK) break;
		if ((rc = cudaCheck(ctx, preload_halo_kernels(), "loading kernels")) != RESTIR_OK) break;
		if ((rc = cudaCheck(ctx, preload_gbuffer_kernels(), "loading kernels")) != RESTIR_OK) break;
		if ((rc = cudaCheck(ctx, preload_bvh_build_kernels(), "loading kernels")) != RESTIR_OK) break;
		if ((rc = cudaCheck(ctx, preload_wide_build_kernels(), "loading kernels")) != RESTIR_OK) break;
		if ((rc = cudaCheck(ctx, preload_generic_kernels(), "loading kernels")) != RESTIR_OK) break;
		if ((rc = cudaCheck(ctx, cudaMalloc(&ctx->bandFlags, 6 * sizeof(unsigned long long)), "cudaMalloc band flags")) != RESTIR_OK) break;
		if ((rc = cudaCheck(ctx, cudaMemsetAsync(ctx->bandFlags, 0, 6 * sizeof(unsigned long long), ctx->stream), "memset")) != RESTIR_OK) break;
		if ((rc = cudaCheck(ctx, cudaMalloc(&ctx->haloTicket, sizeof(unsigned)), "cudaMalloc halo ticket")) != RESTIR_OK) break;
		if ((rc = cudaCheck(ctx, cudaMemsetAsync(ctx->haloTicket, 0, sizeof(unsigned), ctx->stream), "memset")) != RESTIR_OK) break;
		// P12: sRGB8 -> linear table, EOTF in double rounded to float
		float lut[256];
		for (int i = 0; i < 256; ++i) {
			double c = i / 255.0;
			lut[i] = (float)(c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4));
		}
		if ((rc = cudaCheck(ctx, cudaMalloc(&ctx->srgbLut, sizeof(lut)), "cudaMalloc lut")) != RESTIR_OK) break;
		if ((rc = cudaCheck(ctx, cudaMemcpy(ctx->srgbLut, lut, sizeof(lut), cudaMemcpyHostToDevice), "memcpy lut")) != RESTIR_OK) break;
	} while (false);
	if (rc != RESTIR_OK) {
		restir_destroy(ctx);
		return rc;
	}
	*out = ctx;
	return RESTIR_OK;
}

void restir_destroy(restir_context *ctx) {
	if (ctx == nullptr) {
		return;
	}
	cudaSetDevice(ctx->device);
	if (ctx->copyStream) {
		cudaStreamSynchronize(ctx->copyStream);
	}
	if (ctx->stream) {
		cudaStreamSynchronize(ctx->stream);
	}
	for (int s = 0; s < 2; ++s) {
		if (ctx->uploaded[s]) cudaEventDestroy(ctx->uploaded[s]);
		if (ctx->lastRead[s]) cudaEventDestroy(ctx->lastRead[s]);
	}
	if (ctx->copyStream) {
		cudaStreamDestroy(ctx->copyStream);
	}
	dropPeers(ctx);
	for (auto &im : ctx->imported) {
		cudaDestroyExternalMemory(im.second);
	}
	ctx->imported.clear();
	freeDev(ctx->bandFlags);
	freeDev(ctx->haloTicket);
	dropGBuffers(ctx);
	dropProfile(ctx);
	freeDev(ctx->nodes);
	freeDev(ctx->tris);
	freeDev(ctx->bvhScratch);
	freeDev(ctx->treeBlock); // image + triEdges
	freeDev(ctx->shadowed);
	freeDev(ctx->neighborPix);
	freeDev(ctx->neighborM);
	freeDev(ctx->pointBlob);
	freeDev(ctx->triBlob);
	freeDev(ctx->aliasBlob);
	freeDev(ctx->pointPosLum);
	freeDev(ctx->triAux);
	freeDev(ctx->srgbLut);
	freeDev(ctx->gbAttrs);
	freeDev(ctx->gbTriMaterial);
	freeDev(ctx->gbUniforms);
	freeDev(ctx->gbBindings);
	freeDev(ctx->gbTexels);
	freeDev(ctx->gbTextureTable);
	freeDev(ctx->gbSrgbThresholds);
	freeDev(ctx->staging);
	freeDev(ctx->counters);
	freeDev(ctx->traceCursors);
	freeDev(ctx->occluders);
	freeDev(ctx->dedupe);
	freeDev(ctx->aliases);
	freeDev(ctx->aliasCounter);
	for (auto &r : ctx->reservoirs) {
		freeDev(r);
	}
	for (auto &r : ctx->genericReservoirs) {
		freeDev(r);
	}
	if (ctx->ownStream && ctx->stream) {
		cudaStreamDestroy(ctx->stream);
	}
	delete ctx;
}

const char *restir_last_error(const restir_context *ctx) { return ctx ? ctx->error.c_str() : "null context"; }

int restir_synchronize(restir_context *ctx) {
	ENTER(ctx);
	if (ctx->copyStream) {
		CU(ctx, cudaStreamSynchronize(ctx->copyStream));
	}
	CU(ctx, cudaStreamSynchronize(ctx->stream));
	if (bandConnected(ctx)) {
		unsigned long long timeouts = 0;
		CU(ctx, cudaMemcpy(&timeouts, ctx->counters + kCounterHaloTimeout, sizeof(timeouts), cudaMemcpyDeviceToHost));
		if (timeouts != 0) {
			return fail(ctx, RESTIR_E_HALO, "%llu wait(s) for a neighbour's halo rows timed out: the passes after them read stale rows", timeouts);
		}
	}
	return RESTIR_OK;
}

int restir_upload_bvh(restir_context *ctx, const void *nodes, uint32_t n_nodes, const void *triangles, uint32_t n_triangles) {
	ENTER(ctx);
	if (nodes == nullptr || triangles == nullptr || n_nodes == 0 || n_triangles == 0) {
		return fail(ctx, RESTIR_E_INVALID, "restir_upload_bvh: empty tree");
	}
	// Child indices are checked here (the kernels trust them) and the traversal image is derived: the same
	// tree at a 64-byte stride (traversal_image.h).
	std::vector<Node64> image;
	TraversalImageInfo info;
	std::string why;
	if (!build_traversal_image(static_cast<const restir_aabb_node *>(nodes), n_nodes, n_triangles, image, info, why)) {
		return fail(ctx, RESTIR_E_INVALID, "restir_upload_bvh: %s", why.c_str());
	}
	const bool useImage = info.usable && ctx->traversal != RESTIR_TRAVERSAL_REFERENCE_ORDER;
	// the 4-wide quantised image (wide_image.h): walked instead of the binary one wherever its exactness argument applies
	std::vector<WideNode> wide;
	std::vector<uint32_t> triOrder;
	std::vector<float> leafBoxes;
	WideGrid wideGrid{};
	WideImageInfo wideInfo;
	if (useImage && ctx->traversal != RESTIR_TRAVERSAL_IMAGE) {
		build_wide_image(static_cast<const restir_aabb_node *>(nodes), n_nodes, n_triangles, wide, triOrder, leafBoxes, wideGrid, wideInfo);
	} else {
		wideInfo.why = "not requested";
	}
	if (wideInfo.usable) {
		// the triangle records are laid out in the order the wide leaves name them; the binary image (walked by the rays the
		// wide walk does not take, and by the single-kernel variants) names the same records
		std::vector<uint32_t> recordOf(n_triangles);
		for (uint32_t r = 0; r < n_triangles; ++r) {
			recordOf[triOrder[r]] = r;
		}
		for (Node64 &n : image) {
			if (n.left < 0) n.left = ~(int32_t)recordOf[(uint32_t)~n.left];
			if (n.right < 0) n.right = ~(int32_t)recordOf[(uint32_t)~n.right];
		}
	}
	CU(ctx, cudaStreamSynchronize(ctx->stream));
	freeDev(ctx->nodes);
	freeDev(ctx->tris);
	freeDev(ctx->treeBlock);
	ctx->image = ctx->triEdges = nullptr;
	ctx->wide = nullptr;
	ctx->wideOrder = nullptr;
	ctx->wideInfo = WideImageInfo{};
	ctx->nNodes = ctx->nTris = 0;
	freeDev(ctx->gbAttrs); // per-triangle attributes belong to the previous triangle list
	freeDev(ctx->gbTriMaterial);
	ctx->gbTris = 0;
	const size_t triBytes = (size_t)n_triangles * sizeof(restir_triangle);
	CU(ctx, cudaMalloc(&ctx->nodes, (size_t)n_nodes * sizeof(restir_aabb_node)));
	CU(ctx, cudaMalloc(&ctx->tris, triBytes));
	CU(ctx, cudaMemcpyAsync(ctx->nodes, nodes, (size_t)n_nodes * sizeof(restir_aabb_node), cudaMemcpyHostToDevice, ctx->stream));
	CU(ctx, cudaMemcpyAsync(ctx->tris, triangles, triBytes, cudaMemcpyHostToDevice, ctx->stream));
	if (useImage) {
		// what the trace kernel walks, in one allocation: the 64-byte binary nodes, the 64-byte (p1, e1, e2 | leaf box) triangle
		// records, the 64-byte 4-wide nodes (the leaf boxes pass through the tail of the block on their way into the records)
		const size_t imageBytes = (image.size() * sizeof(Node64) + 255) & ~(size_t)255;
		const size_t edgeBytes = (size_t)n_triangles * 64;
		const size_t wideBytes = (wide.size() * sizeof(WideNode) + 255) & ~(size_t)255;
		const size_t boxBytes = wideInfo.usable ? leafBoxes.size() * sizeof(float) : 0;
		const size_t orderBytes = wideInfo.usable ? triOrder.size() * sizeof(uint32_t) : 0;
		CU(ctx, cudaMalloc(&ctx->treeBlock, imageBytes + edgeBytes + wideBytes + boxBytes + orderBytes));
		ctx->treeBlockBytes = imageBytes + edgeBytes + wideBytes + boxBytes + orderBytes;
		ctx->image = reinterpret_cast<float4 *>(ctx->treeBlock);
		ctx->triEdges = reinterpret_cast<float4 *>(ctx->treeBlock + imageBytes);
		CU(ctx, cudaMemcpyAsync(ctx->image, image.data(), image.size() * sizeof(Node64), cudaMemcpyHostToDevice, ctx->stream));
		const float *boxes = nullptr;
		const uint32_t *order = nullptr;
		if (wideInfo.usable) {
			ctx->wide = reinterpret_cast<uint4 *>(ctx->treeBlock + imageBytes + edgeBytes);
			float *dstBoxes = reinterpret_cast<float *>(ctx->treeBlock + imageBytes + edgeBytes + wideBytes);
			uint32_t *dstOrder = reinterpret_cast<uint32_t *>(ctx->treeBlock + imageBytes + edgeBytes + wideBytes + boxBytes);
			CU(ctx, cudaMemcpyAsync(ctx->wide, wide.data(), wide.size() * sizeof(WideNode), cudaMemcpyHostToDevice, ctx->stream));
			CU(ctx, cudaMemcpyAsync(dstBoxes, leafBoxes.data(), boxBytes, cudaMemcpyHostToDevice, ctx->stream));
			CU(ctx, cudaMemcpyAsync(dstOrder, triOrder.data(), orderBytes, cudaMemcpyHostToDevice, ctx->stream));
			boxes = dstBoxes;
			order = dstOrder;
		}
		ctx->wideOrder = order;
		launch_derive_triangle_edges(ctx->tris, n_triangles, ctx->triEdges, order, boxes, ctx->stream);
		CU(ctx, cudaGetLastError());
	}
	CU(ctx, cudaStreamSynchronize(ctx->stream));
	ctx->nNodes = n_nodes;
	ctx->nTris = n_triangles;
	ctx->imageInfo = info;
	ctx->wideGrid = wideGrid;
	ctx->wideInfo = wideInfo;
	return clearOccluders(ctx); // its entries name triangle records of the previous tree
}

int restir_build_bvh_device(restir_context *ctx, const void *triangles, uint32_t n_triangles, void *nodes_out) {
	ENTER(ctx);
	if (triangles == nullptr || n_triangles < 2) {
		return fail(ctx, RESTIR_E_INVALID, "restir_build_bvh_device: at least two triangles (the reference asserts on one, aabbTreeBuilder.cpp:212)");
	}
	if (n_triangles > (1u << 30)) {
		return fail(ctx, RESTIR_E_UNSUPPORTED, "restir_build_bvh_device: more than 2^30 triangles");
	}
	CU(ctx, cudaStreamSynchronize(ctx->stream));
	freeDev(ctx->nodes);
	freeDev(ctx->tris);
	freeDev(ctx->treeBlock);
	ctx->image = ctx->triEdges = nullptr;
	ctx->wide = nullptr;
	ctx->wideOrder = nullptr;
	ctx->wideInfo = WideImageInfo{};
	ctx->nNodes = ctx->nTris = 0;
	freeDev(ctx->gbAttrs);
	freeDev(ctx->gbTriMaterial);
	ctx->gbTris = 0;
	const uint32_t nNodes = n_triangles - 1;
	const size_t triBytes = (size_t)n_triangles * sizeof(restir_triangle);
	CU(ctx, cudaMalloc(&ctx->nodes, (size_t)nNodes * sizeof(restir_aabb_node)));
	CU(ctx, cudaMalloc(&ctx->tris, triBytes));
	CU(ctx, cudaMemcpyAsync(ctx->tris, triangles, triBytes, cudaMemcpyHostToDevice, ctx->stream));
	if (ctx->bvhScratchBytes < bvh_build_scratch_bytes(n_triangles)) {
		freeDev(ctx->bvhScratch);
		ctx->bvhScratchBytes = 0;
		CU(ctx, cudaMalloc(&ctx->bvhScratch, bvh_build_scratch_bytes(n_triangles)));
		ctx->bvhScratchBytes = bvh_build_scratch_bytes(n_triangles);
	}
	void *scratch = ctx->bvhScratch;
	int levels = 0;
	unsigned nonFinite = 0;
	beforeLaunch(ctx, "bvh_build (all kernels)");
	cudaError_t e = build_aabb_tree_device(ctx->tris, n_triangles, reinterpret_cast<restir_aabb_node *>(ctx->nodes), scratch, &levels, &nonFinite,
	                                       ctx->stream);
	int rc = afterLaunch(ctx, "bvh_build (all kernels)");
	if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
	if (e != cudaSuccess) return cudaCheck(ctx, e, "restir_build_bvh_device");
	if (rc != RESTIR_OK) return rc;
	if (nonFinite != 0) {
		freeDev(ctx->nodes);
		freeDev(ctx->tris);
		return fail(ctx, RESTIR_E_UNSUPPORTED,
		            "restir_build_bvh_device: %u non-finite vertex coordinates (the reference's min / max folds depend on the order NaN is met in: "
		            "build such input with restir_build_aabb_tree on the host)", nonFinite);
	}
	// a tree built here is a tree over [0, n) by construction; a walk in any order holds at most one pending sibling per level
	TraversalImageInfo info;
	info.reachableNodes = nNodes;
	info.depth = levels;
	info.referenceStackBound = info.anyOrderStackBound = std::max(1, levels); // upper bounds (restir_check_aabb_tree computes the exact ones)
	info.usable = levels <= 32 && ctx->traversal != RESTIR_TRAVERSAL_REFERENCE_ORDER;
	WideGrid wideGrid{};
	WideImageInfo wideInfo;
	if (info.usable) {
		// one allocation: the 64-byte binary nodes, the 64-byte triangle records, then (worst case: as many as binary nodes) the
		// 64-byte 4-wide nodes, the leaf boxes and the record order on their way into the records
		const size_t imageBytes = ((size_t)nNodes * 64 + 255) & ~(size_t)255;
		const size_t edgeBytes = (size_t)n_triangles * 64;
		const bool wantWide = ctx->traversal != RESTIR_TRAVERSAL_IMAGE;
		const size_t wideBytes = wantWide ? ((size_t)nNodes * sizeof(WideNode) + 255) & ~(size_t)255 : 0;
		const size_t boxBytes = wantWide ? (((size_t)n_triangles * 6 * sizeof(float)) + 255) & ~(size_t)255 : 0;
		const size_t orderBytes = wantWide ? (size_t)n_triangles * sizeof(unsigned) : 0;
		CU(ctx, cudaMalloc(&ctx->treeBlock, imageBytes + edgeBytes + wideBytes + boxBytes + orderBytes));
		ctx->treeBlockBytes = imageBytes + edgeBytes + wideBytes + boxBytes + orderBytes;
		ctx->image = reinterpret_cast<float4 *>(ctx->treeBlock);
		ctx->triEdges = reinterpret_cast<float4 *>(ctx->treeBlock + imageBytes);
		launch_bvh_image(reinterpret_cast<const restir_aabb_node *>(ctx->nodes), nNodes, ctx->image, ctx->stream);
		const unsigned *order = nullptr;
		const float *boxes = nullptr;
		if (wantWide) {
			// the grid needs the scene's bounds: the two boxes of the root (the tree is nested by construction)
			restir_aabb_node root;
			CU(ctx, cudaMemcpyAsync(&root, ctx->nodes, sizeof(root), cudaMemcpyDeviceToHost, ctx->stream));
			CU(ctx, cudaStreamSynchronize(ctx->stream));
			double lo[3], hi[3];
			for (int a = 0; a < 3; ++a) {
				lo[a] = std::min((double)root.leftAabbMin[a], (double)root.rightAabbMin[a]);
				hi[a] = std::max((double)root.leftAabbMax[a], (double)root.rightAabbMax[a]);
			}
			WideQuant quant;
			if (!make_wide_grid(lo, hi, wideGrid, quant)) {
				wideInfo.why = "scene coordinates out of the range the quantised grid is defined for";
			} else {
				const size_t need = wide_build_scratch_bytes(nNodes, n_triangles);
				if (ctx->bvhScratchBytes < need) {
					freeDev(ctx->bvhScratch);
					ctx->bvhScratchBytes = 0;
					CU(ctx, cudaMalloc(&ctx->bvhScratch, need));
					ctx->bvhScratchBytes = need;
				}
				WideNode *wide = reinterpret_cast<WideNode *>(ctx->treeBlock + imageBytes + edgeBytes);
				float *dstBoxes = reinterpret_cast<float *>(ctx->treeBlock + imageBytes + edgeBytes + wideBytes);
				unsigned *dstOrder = reinterpret_cast<unsigned *>(ctx->treeBlock + imageBytes + edgeBytes + wideBytes + boxBytes);
				unsigned nWide = 0;
				int depth = 0;
				bool usable = false;
				beforeLaunch(ctx, "wide_build (all kernels)");
				cudaError_t we = build_wide_image_device(reinterpret_cast<const restir_aabb_node *>(ctx->nodes), nNodes, n_triangles, quant, wide, dstOrder, dstBoxes,
				                                         ctx->image, ctx->bvhScratch, &nWide, &depth, &usable, ctx->stream);
				int wrc = afterLaunch(ctx, "wide_build (all kernels)");
				if (we != cudaSuccess) return cudaCheck(ctx, we, "restir_build_bvh_device (wide image)");
				if (wrc != RESTIR_OK) return wrc;
				if (usable) {
					ctx->wide = reinterpret_cast<uint4 *>(wide);
					wideInfo.usable = true;
					wideInfo.nodes = nWide;
					wideInfo.depth = wideInfo.stackBound = depth;
					order = dstOrder;
					boxes = dstBoxes;
				} else {
					wideInfo.why = "the wide image could not be derived from this tree (deeper than the walk's stack, or a box off the grid)";
					launch_bvh_image(reinterpret_cast<const restir_aabb_node *>(ctx->nodes), nNodes, ctx->image, ctx->stream); // undo a partial remap
				}
			}
		} else {
			wideInfo.why = "not requested";
		}
		ctx->wideOrder = order;
		launch_derive_triangle_edges(ctx->tris, n_triangles, ctx->triEdges, order, boxes, ctx->stream);
		CU(ctx, cudaGetLastError());
	} else {
		info.why = "deeper than the 32-entry stack";
	}
	if (nodes_out != nullptr) {
		CU(ctx, cudaMemcpyAsync(nodes_out, ctx->nodes, (size_t)nNodes * sizeof(restir_aabb_node), cudaMemcpyDeviceToHost, ctx->stream));
	}
	CU(ctx, cudaStreamSynchronize(ctx->stream));
	ctx->nNodes = nNodes;
	ctx->nTris = n_triangles;
	ctx->imageInfo = info;
	ctx->wideGrid = wideGrid;
	ctx->wideInfo = wideInfo;
	return clearOccluders(ctx);
}

static int uploadBlob(restir_context *ctx, const void *blob, size_t bytes, size_t stride, unsigned char *&dst, int &count, const char *name) {
	freeDev(dst);
	count = 0;
	if (blob == nullptr || bytes < RESTIR_BLOB_HEADER_BYTES) {
		return fail(ctx, RESTIR_E_INVALID, "%s blob must be at least %d bytes", name, RESTIR_BLOB_HEADER_BYTES);
	}
	int32_t n;
	std::memcpy(&n, blob, 4);
	if (n < 0 || (size_t)n * stride + RESTIR_BLOB_HEADER_BYTES > bytes) {
		return fail(ctx, RESTIR_E_INVALID, "%s blob: count %d does not fit in %zu bytes", name, n, bytes);
	}
	CU(ctx, cudaMalloc(&dst, bytes));
	CU(ctx, cudaMemcpyAsync(dst, blob, bytes, cudaMemcpyHostToDevice, ctx->stream));
	count = n;
	return RESTIR_OK;
}

int restir_upload_lights(restir_context *ctx, const void *point_blob, size_t point_bytes, const void *tri_blob, size_t tri_bytes,
                         const void *alias_blob, size_t alias_bytes) {
	ENTER(ctx);
	CU(ctx, cudaStreamSynchronize(ctx->stream));
	int rc;
	if ((rc = uploadBlob(ctx, point_blob, point_bytes, sizeof(restir_point_light), ctx->pointBlob, ctx->pointCount, "point-light")) != RESTIR_OK) return rc;
	if ((rc = uploadBlob(ctx, tri_blob, tri_bytes, sizeof(restir_tri_light), ctx->triBlob, ctx->triCount, "triangle-light")) != RESTIR_OK) return rc;
	if ((rc = uploadBlob(ctx, alias_blob, alias_bytes, sizeof(restir_alias_column), ctx->aliasBlob, ctx->aliasCount, "alias-table")) != RESTIR_OK) return rc;
	int lights = ctx->pointCount != 0 ? ctx->pointCount : ctx->triCount; // restirOmni.glsl:116 picks the list the same way
	if (ctx->aliasCount == 0 || ctx->aliasCount != lights) {
		const int columns = ctx->aliasCount;
		freeDev(ctx->aliasBlob); // no pass may run on mismatched tables (makeParams checks the alias blob)
		ctx->aliasCount = 0;
		return fail(ctx, RESTIR_E_INVALID, "alias table has %d columns for %d lights", columns, lights);
	}
	// a stored reservoir names its light by index and its normal / emission are re-read from the tables (PackedReservoir):
	// history sampled from the previous tables means nothing under the new ones (and could index past them)
	for (auto &r : ctx->reservoirs) {
		if (r) CU(ctx, cudaMemsetAsync(r, 0, ctx->allocPixels() * sizeof(PackedReservoir), ctx->stream));
	}
	for (auto &r : ctx->genericReservoirs) {
		if (r) CU(ctx, cudaMemsetAsync(r, 0, ctx->allocPixels() * ctx->reservoirBytes(), ctx->stream));
	}
	{
		int rc = clearOccluders(ctx); // keyed by light index
		if (rc != RESTIR_OK) return rc;
	}
	freeDev(ctx->pointPosLum);
	freeDev(ctx->triAux);
	if (ctx->pointCount) CU(ctx, cudaMalloc(&ctx->pointPosLum, sizeof(float4) * (size_t)ctx->pointCount));
	if (ctx->triCount) CU(ctx, cudaMalloc(&ctx->triAux, sizeof(float4) * (size_t)ctx->triCount));
	SceneView v = sceneView(ctx);
	launch_derive_light_tables(v.pointLights, ctx->pointCount, ctx->pointPosLum, v.triLights, ctx->triCount, ctx->triAux, ctx->stream);
	CU(ctx, cudaGetLastError());
	CU(ctx, cudaStreamSynchronize(ctx->stream));
	return RESTIR_OK;
}

int restir_resize_band(restir_context *ctx, uint32_t width, uint32_t height, uint32_t row_begin, uint32_t row_end, uint32_t halo) {
	ENTER(ctx);
	if (width == 0 || height == 0 || row_begin >= row_end || row_end > height || width > (1u << 20) || height > (1u << 20)) {
		return fail(ctx, RESTIR_E_INVALID, "restir_resize: bad geometry %ux%u rows [%u,%u)", width, height, row_begin, row_end);
	}
	if (ctx->copyStream) {
		CU(ctx, cudaStreamSynchronize(ctx->copyStream));
	}
	CU(ctx, cudaStreamSynchronize(ctx->stream));
	dropPeers(ctx);
	dropGBuffers(ctx);
	for (int s = 0; s < 2; ++s) {
		ctx->uploadPending[s] = ctx->readRecorded[s] = false;
	}
	for (auto &r : ctx->reservoirs) {
		freeDev(r);
	}
	for (auto &r : ctx->genericReservoirs) {
		freeDev(r);
	}
	freeDev(ctx->staging);
	ctx->stagingPixels = 0;
	Band b;
	b.W = (int)width;
	b.H = (int)height;
	b.rowBegin = (int)row_begin;
	b.rowEnd = (int)row_end;
	b.allocBegin = (int)(row_begin > halo ? row_begin - halo : 0);
	b.allocEnd = (int)((uint64_t)row_end + halo < height ? row_end + halo : height);
	ctx->band = b;
	// occluder cache: 256 entries per 64 x 32-pixel region of the rows this context holds, all empty
	freeDev(ctx->occluders);
	ctx->regionsX = (width + 63) / 64;
	ctx->occluderEntries = (size_t)ctx->regionsX * (((size_t)(b.allocEnd - b.allocBegin) + 31) / 32) * 256 * kOccluderWays;
	CU(ctx, cudaMalloc(&ctx->occluders, ctx->occluderEntries * sizeof(unsigned)));
	CU(ctx, cudaMemsetAsync(ctx->occluders, 0xff, ctx->occluderEntries * sizeof(unsigned), ctx->stream));
	if (ctx->generic()) {
		size_t bytes = ctx->allocPixels() * ctx->reservoirBytes();
		for (auto &r : ctx->genericReservoirs) { // app.h:264-284: three buffers, zero-filled
			CU(ctx, cudaMalloc(&r, bytes));
			CU(ctx, cudaMemsetAsync(r, 0, bytes, ctx->stream));
		}
	} else {
		size_t bytes = ctx->allocPixels() * sizeof(PackedReservoir);
		for (auto &r : ctx->reservoirs) { // app.h:264-284: three buffers, zero-filled
			CU(ctx, cudaMalloc(&r, bytes));
			CU(ctx, cudaMemsetAsync(r, 0, bytes, ctx->stream));
		}
	}
	CU(ctx, cudaStreamSynchronize(ctx->stream));
	return RESTIR_OK;
}

int restir_resize(restir_context *ctx, uint32_t width, uint32_t height) { return restir_resize_band(ctx, width, height, 0, height, 0); }

int restir_get_band(const restir_context *ctx, uint32_t *row_begin, uint32_t *row_end, uint32_t *alloc_begin, uint32_t *alloc_end) {
	if (ctx == nullptr) {
		return RESTIR_E_INVALID;
	}
	if (row_begin) *row_begin = (uint32_t)ctx->band.rowBegin;
	if (row_end) *row_end = (uint32_t)ctx->band.rowEnd;
	if (alloc_begin) *alloc_begin = (uint32_t)ctx->band.allocBegin;
	if (alloc_end) *alloc_end = (uint32_t)ctx->band.allocEnd;
	return RESTIR_OK;
}

int restir_bind_gbuffer(restir_context *ctx, int slot, restir_gbuffer_format format, const restir_gbuffer_planes *pl) {
	ENTER(ctx);
	if (slot < 0 || slot > 1 || pl == nullptr) {
		return fail(ctx, RESTIR_E_INVALID, "restir_bind_gbuffer: bad slot");
	}
	if (format != RESTIR_GBUFFER_NVIDIA_DEFAULT) {
		return fail(ctx, RESTIR_E_UNSUPPORTED, "G-buffer format %d not supported", (int)format);
	}
	if (!pl->albedo || !pl->normal || !pl->material || !pl->worldPos || !pl->depth) {
		return fail(ctx, RESTIR_E_INVALID, "restir_bind_gbuffer: all five planes are required");
	}
	GBufferView v;
	v.albedo = static_cast<const uchar4 *>(pl->albedo);
	v.normal = static_cast<const short4 *>(pl->normal);
	v.material = static_cast<const ushort2 *>(pl->material);
	v.worldPos = static_cast<const float4 *>(pl->worldPos);
	v.depth = static_cast<const float *>(pl->depth);
	ctx->gbuf[slot] = v;
	ctx->uploadPending[slot] = false;
	return RESTIR_OK;
}

int restir_upload_gbuffer(restir_context *ctx, int slot, restir_gbuffer_format format, const restir_gbuffer_planes *pl) {
	ENTER(ctx);
	if (slot < 0 || slot > 1 || pl == nullptr || ctx->band.W == 0) {
		return fail(ctx, RESTIR_E_INVALID, "restir_upload_gbuffer: bad slot or restir_resize not called");
	}
	if (format != RESTIR_GBUFFER_NVIDIA_DEFAULT) {
		return fail(ctx, RESTIR_E_UNSUPPORTED, "G-buffer format %d not supported", (int)format);
	}
	const void *src[5] = {pl->albedo, pl->normal, pl->material, pl->worldPos, pl->depth};
	size_t px = ctx->allocPixels();
	for (int k = 0; k < 5; ++k) {
		if (src[k] == nullptr) {
			return fail(ctx, RESTIR_E_INVALID, "restir_upload_gbuffer: all five planes are required");
		}
	}
	if (ctx->copyStream == nullptr) {
		CU(ctx, cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking));
	}
	if (ctx->uploaded[slot] == nullptr) {
		CU(ctx, cudaEventCreateWithFlags(&ctx->uploaded[slot], cudaEventDisableTiming));
	}
	// the copy waits for the last kernel that read this slot (for frame f+1 that is frame f's temporal kernel,
	// restirOmni.glsl:163-209: the reuse and lighting passes of frame f only read the other slot), not for the
	// whole compute stream
	const bool ownedBefore = ctx->gbuf[slot].worldPos != nullptr && ctx->gbuf[slot].worldPos == ctx->ownedPlanes[slot][3];
	if (ownedBefore && ctx->readRecorded[slot]) {
		CU(ctx, cudaStreamWaitEvent(ctx->copyStream, ctx->lastRead[slot], 0));
	} else if (ctx->ownedPlanes[slot][0] != nullptr) {
		// planes exist but the slot was rebound in between: order after everything issued so far
		int rc = markRead(ctx, slot);
		if (rc != RESTIR_OK) return rc;
		CU(ctx, cudaStreamWaitEvent(ctx->copyStream, ctx->lastRead[slot], 0));
	}
	for (int k = 0; k < 5; ++k) {
		if (ctx->ownedPlanes[slot][k] == nullptr) {
			CU(ctx, cudaMalloc(&ctx->ownedPlanes[slot][k], px * kPlaneBytes[k]));
		}
		CU(ctx, cudaMemcpyAsync(ctx->ownedPlanes[slot][k], src[k], px * kPlaneBytes[k], cudaMemcpyHostToDevice, ctx->copyStream));
	}
	CU(ctx, cudaEventRecord(ctx->uploaded[slot], ctx->copyStream));
	restir_gbuffer_planes dev{ctx->ownedPlanes[slot][0], ctx->ownedPlanes[slot][1], ctx->ownedPlanes[slot][2], ctx->ownedPlanes[slot][3],
	                          ctx->ownedPlanes[slot][4]};
	int rc = restir_bind_gbuffer(ctx, slot, format, &dev);
	ctx->uploadPending[slot] = rc == RESTIR_OK;
	return rc;
}

// ---- the G-buffer pass -----------------------------------------------------------------------------------------------

int restir_upload_geometry(restir_context *ctx, const restir_vertex *vertices, uint64_t n_vertices, const uint32_t *indices, uint64_t n_indices,
                           const restir_draw *draws, const restir_model_matrices *matrices, uint32_t n_draws) {
	ENTER(ctx);
	if (ctx->nodes == nullptr) {
		return fail(ctx, RESTIR_E_INVALID, "restir_upload_geometry: call restir_upload_bvh first (the draws' triangles are the tree's triangles)");
	}
	if (!vertices || !indices || !draws || !matrices || n_vertices == 0 || n_indices == 0 || n_draws == 0) {
		return fail(ctx, RESTIR_E_INVALID, "restir_upload_geometry: empty geometry");
	}
	// the draws, validated (the kernels trust them), and which draw every triangle belongs to
	std::vector<uint32_t> triDraw, drawFirstTri(n_draws);
	std::vector<int> triMaterial;
	for (uint32_t d = 0; d < n_draws; ++d) {
		const restir_draw &dr = draws[d];
		if (dr.indexCount % 3 != 0 || (uint64_t)dr.firstIndex + dr.indexCount > n_indices) {
			return fail(ctx, RESTIR_E_INVALID, "restir_upload_geometry: draw %u reads indices [%u, %u + %u) of %llu", d, dr.firstIndex, dr.firstIndex,
			            dr.indexCount, (unsigned long long)n_indices);
		}
		for (uint32_t i = 0; i < dr.indexCount; ++i) {
			if ((uint64_t)dr.vertexOffset + indices[dr.firstIndex + i] >= n_vertices) {
				return fail(ctx, RESTIR_E_INVALID, "restir_upload_geometry: draw %u addresses vertex %llu of %llu", d,
				            (unsigned long long)dr.vertexOffset + indices[dr.firstIndex + i], (unsigned long long)n_vertices);
			}
		}
		drawFirstTri[d] = (uint32_t)triDraw.size();
		triDraw.insert(triDraw.end(), dr.indexCount / 3, d);
		triMaterial.insert(triMaterial.end(), dr.indexCount / 3, dr.materialIndex);
	}
	if (triDraw.size() != ctx->nTris) {
		return fail(ctx, RESTIR_E_INVALID, "restir_upload_geometry: the draws make %zu triangles, the uploaded tree has %u", triDraw.size(), ctx->nTris);
	}
	CU(ctx, cudaStreamSynchronize(ctx->stream));
	restir_vertex *dV = nullptr;
	uint32_t *dI = nullptr, *dTriDraw = nullptr, *dFirst = nullptr;
	restir_draw *dD = nullptr;
	restir_model_matrices *dM = nullptr;
	int rc = RESTIR_OK;
	do {
		if ((rc = uploadArray(ctx, dV, vertices, (size_t)n_vertices, "vertices")) != RESTIR_OK) break;
		if ((rc = uploadArray(ctx, dI, indices, (size_t)n_indices, "indices")) != RESTIR_OK) break;
		if ((rc = uploadArray(ctx, dD, draws, (size_t)n_draws, "draws")) != RESTIR_OK) break;
		if ((rc = uploadArray(ctx, dM, matrices, (size_t)n_draws, "matrices")) != RESTIR_OK) break;
		if ((rc = uploadArray(ctx, dTriDraw, triDraw.data(), triDraw.size(), "triangle -> draw")) != RESTIR_OK) break;
		if ((rc = uploadArray(ctx, dFirst, drawFirstTri.data(), drawFirstTri.size(), "draw -> first triangle")) != RESTIR_OK) break;
		if ((rc = uploadArray(ctx, ctx->gbTriMaterial, triMaterial.data(), triMaterial.size(), "triangle materials")) != RESTIR_OK) break;
		freeDev(ctx->gbAttrs);
		if ((rc = cudaCheck(ctx, cudaMalloc(&ctx->gbAttrs, (size_t)ctx->nTris * 8 * sizeof(float4)), "triangle attributes")) != RESTIR_OK) break;
		beforeLaunch(ctx, "vertex_stage_kernel"); // gBuffer.vert, once per upload
		launch_vertex_stage(dV, dI, dD, dM, dTriDraw, dFirst, ctx->nTris, ctx->gbAttrs, ctx->stream);
		if ((rc = afterLaunch(ctx, "vertex_stage_kernel")) != RESTIR_OK) break;
		rc = cudaCheck(ctx, cudaStreamSynchronize(ctx->stream), "vertex stage");
	} while (false);
	freeDev(dV);
	freeDev(dI);
	freeDev(dD);
	freeDev(dM);
	freeDev(dTriDraw);
	freeDev(dFirst);
	ctx->gbTris = rc == RESTIR_OK ? ctx->nTris : 0;
	return rc;
}

int restir_upload_materials(restir_context *ctx, const restir_material_uniforms *uniforms, const restir_material_textures *bindings,
                            uint32_t n_materials, const restir_texture *textures, uint32_t n_textures) {
	ENTER(ctx);
	if (!uniforms || !bindings || n_materials == 0 || (n_textures != 0 && textures == nullptr)) {
		return fail(ctx, RESTIR_E_INVALID, "restir_upload_materials: no materials");
	}
	std::vector<uint4> table(n_textures);
	size_t texels = 0;
	for (uint32_t t = 0; t < n_textures; ++t) {
		if (textures[t].rgba8 == nullptr || textures[t].width == 0 || textures[t].height == 0 || textures[t].width > 32768 || textures[t].height > 32768) {
			return fail(ctx, RESTIR_E_INVALID, "restir_upload_materials: texture %u is empty or larger than 32768 texels a side", t);
		}
		if (texels > 0xffffffffull) {
			return fail(ctx, RESTIR_E_UNSUPPORTED, "restir_upload_materials: more than 2^32 texels");
		}
		table[t] = make_uint4((unsigned)texels, textures[t].width, textures[t].height, 0u);
		texels += (size_t)textures[t].width * textures[t].height;
	}
	CU(ctx, cudaStreamSynchronize(ctx->stream));
	int rc;
	if ((rc = uploadArray(ctx, ctx->gbUniforms, uniforms, n_materials, "material uniforms")) != RESTIR_OK) return rc;
	if ((rc = uploadArray(ctx, ctx->gbBindings, bindings, n_materials, "material textures")) != RESTIR_OK) return rc;
	if ((rc = uploadArray(ctx, ctx->gbTextureTable, table.data(), table.size(), "texture table")) != RESTIR_OK) return rc;
	freeDev(ctx->gbTexels);
	if (texels) {
		CU(ctx, cudaMalloc(&ctx->gbTexels, texels * sizeof(uchar4)));
		for (uint32_t t = 0; t < n_textures; ++t) {
			CU(ctx, cudaMemcpyAsync(ctx->gbTexels + table[t].x, textures[t].rgba8, (size_t)table[t].y * table[t].z * 4, cudaMemcpyHostToDevice, ctx->stream));
		}
	}
	if (ctx->gbSrgbThresholds == nullptr) {
		// the smallest float that an R8G8B8A8_SRGB attachment stores as code c (round to nearest in the encoded domain): the
		// EOTF of the midpoint between codes c - 1 and c, in double, rounded up to float
		float thr[256];
		thr[0] = -INFINITY;
		for (int c = 1; c < 256; ++c) {
			double e = (c - 0.5) / 255.0;
			double lin = e <= 0.04045 ? e / 12.92 : std::pow((e + 0.055) / 1.055, 2.4);
			float f = (float)lin;
			if ((double)f < lin) f = std::nextafterf(f, INFINITY);
			thr[c] = f;
		}
		float *d = nullptr;
		if ((rc = uploadArray(ctx, d, thr, 256, "sRGB thresholds")) != RESTIR_OK) return rc;
		ctx->gbSrgbThresholds = d;
	}
	CU(ctx, cudaStreamSynchronize(ctx->stream));
	ctx->gbMaterials = (int)n_materials;
	ctx->gbTextures = (int)n_textures;
	return RESTIR_OK;
}

int restir_pass_gbuffer(restir_context *ctx, int slot, const restir_camera *camera) {
	ENTER(ctx);
	if (slot < 0 || slot > 1 || camera == nullptr || ctx->band.W == 0) {
		return fail(ctx, RESTIR_E_INVALID, "restir_pass_gbuffer: bad slot, null camera or restir_resize not called");
	}
	if (ctx->nodes == nullptr || ctx->gbAttrs == nullptr || ctx->gbTris != ctx->nTris || ctx->gbUniforms == nullptr) {
		return fail(ctx, RESTIR_E_INVALID, "restir_pass_gbuffer: upload the BVH, the geometry (restir_upload_geometry) and the materials (restir_upload_materials) first");
	}
	// the planes are the context's: allocated on first use, bound to the slot like restir_upload_gbuffer's.  An upload still in
	// flight into them (copy stream) is ordered before this pass.
	int rc;
	if ((rc = waitUpload(ctx, slot)) != RESTIR_OK) return rc;
	const size_t px = ctx->allocPixels();
	for (int k = 0; k < 5; ++k) {
		if (ctx->ownedPlanes[slot][k] == nullptr) {
			CU(ctx, cudaMalloc(&ctx->ownedPlanes[slot][k], px * kPlaneBytes[k]));
		}
	}
	RaycastCamera cam;
	cameraBasisOf(camera, cam);
	GBufferScene g{};
	g.attrs = ctx->gbAttrs;
	g.triMaterial = ctx->gbTriMaterial;
	g.uniforms = ctx->gbUniforms;
	g.bindings = ctx->gbBindings;
	g.texels = ctx->gbTexels;
	g.textureTable = ctx->gbTextureTable;
	g.srgbThresholds = ctx->gbSrgbThresholds;
	g.recordTri = ctx->wideOrder;
	g.nMaterials = ctx->gbMaterials;
	g.nTextures = ctx->gbTextures;
	Band full = ctx->band; // every row the context holds: the reuse passes gather from the halo rows
	full.rowBegin = full.allocBegin;
	full.rowEnd = full.allocEnd;
	SceneView sv = sceneView(ctx);
	if (ctx->traversal == RESTIR_TRAVERSAL_REFERENCE_ORDER) {
		sv.image = nullptr; // the literal walk of the uploaded 80-byte nodes (what trees without an image get): same planes, kept testable
	}
	beforeLaunch(ctx, "gbuffer_kernel");
	launch_gbuffer(sv, g, full, cam, camera->zNear, camera->zFar, ctx->ownedPlanes[slot][0], ctx->ownedPlanes[slot][1],
	               ctx->ownedPlanes[slot][2], ctx->ownedPlanes[slot][3], ctx->ownedPlanes[slot][4], ctx->stream);
	if ((rc = afterLaunch(ctx, "gbuffer_kernel")) != RESTIR_OK) return rc;
	restir_gbuffer_planes dev{ctx->ownedPlanes[slot][0], ctx->ownedPlanes[slot][1], ctx->ownedPlanes[slot][2], ctx->ownedPlanes[slot][3],
	                          ctx->ownedPlanes[slot][4]};
	return restir_bind_gbuffer(ctx, slot, RESTIR_GBUFFER_NVIDIA_DEFAULT, &dev);
}

int restir_gbuffer_device_planes(restir_context *ctx, int slot, restir_gbuffer_planes *out) {
	ENTER(ctx);
	if (slot < 0 || slot > 1 || out == nullptr || ctx->ownedPlanes[slot][0] == nullptr) {
		return fail(ctx, RESTIR_E_INVALID, "restir_gbuffer_device_planes: slot %d has no context-owned planes", slot);
	}
	*out = restir_gbuffer_planes{ctx->ownedPlanes[slot][0], ctx->ownedPlanes[slot][1], ctx->ownedPlanes[slot][2], ctx->ownedPlanes[slot][3],
	                             ctx->ownedPlanes[slot][4]};
	return RESTIR_OK;
}

int restir_import_external_memory(restir_context *ctx, int fd, uint64_t size, void **device_ptr) {
	ENTER(ctx);
	if (device_ptr == nullptr || fd < 0 || size == 0) {
		return fail(ctx, RESTIR_E_INVALID, "restir_import_external_memory: bad descriptor, size or result pointer");
	}
	*device_ptr = nullptr;
	cudaExternalMemoryHandleDesc hd{};
	hd.type = cudaExternalMemoryHandleTypeOpaqueFd;
	hd.handle.fd = fd;
	hd.size = size;
	cudaExternalMemory_t mem = nullptr;
	cudaError_t e = cudaImportExternalMemory(&mem, &hd);
	if (e != cudaSuccess) { // not sticky: a descriptor that is not an exported allocation says nothing about the device
		cudaGetLastError();
		return fail(ctx, RESTIR_E_INVALID, "cudaImportExternalMemory: %s", cudaGetErrorString(e));
	}
	cudaExternalMemoryBufferDesc bd{};
	bd.offset = 0;
	bd.size = size;
	void *ptr = nullptr;
	e = cudaExternalMemoryGetMappedBuffer(&ptr, mem, &bd);
	if (e != cudaSuccess) {
		cudaGetLastError();
		cudaDestroyExternalMemory(mem);
		return fail(ctx, RESTIR_E_INVALID, "cudaExternalMemoryGetMappedBuffer: %s", cudaGetErrorString(e));
	}
	ctx->imported.emplace_back(ptr, mem);
	*device_ptr = ptr;
	return RESTIR_OK;
}

int restir_release_external_memory(restir_context *ctx, void *device_ptr) {
	ENTER(ctx);
	for (size_t k = 0; k < ctx->imported.size(); ++k) {
		if (ctx->imported[k].first == device_ptr) {
			CU(ctx, cudaStreamSynchronize(ctx->stream));
			cudaFree(device_ptr); // the mapping; then the import itself
			cudaDestroyExternalMemory(ctx->imported[k].second);
			ctx->imported.erase(ctx->imported.begin() + (long)k);
			return RESTIR_OK;
		}
	}
	return fail(ctx, RESTIR_E_INVALID, "restir_release_external_memory: not a pointer restir_import_external_memory returned");
}

int restir_set_uniforms(restir_context *ctx, const restir_uniforms *u) {
	ENTER(ctx);
	if (u == nullptr) {
		return fail(ctx, RESTIR_E_INVALID, "null uniforms");
	}
	ctx->uniforms = *u;
	ctx->haveUniforms = true;
	return RESTIR_OK;
}

int restir_set_lighting_uniforms(restir_context *ctx, const restir_lighting_uniforms *u) {
	ENTER(ctx);
	if (u == nullptr) {
		return fail(ctx, RESTIR_E_INVALID, "null uniforms");
	}
	if (u->debugMode != 0) {
		return fail(ctx, RESTIR_E_UNSUPPORTED, "lighting debugMode %d: only GBUFFER_DEBUG_NONE (0) is on the hot path", u->debugMode);
	}
	ctx->lighting = *u;
	ctx->haveLighting = true;
	return RESTIR_OK;
}

int restir_set_reservoir_variant(restir_context *ctx, uint32_t reservoir_size, int unbiased_mis, int fused_passes) {
	ENTER(ctx);
	if (!generic_variant_supported((int)reservoir_size, unbiased_mis != 0)) {
		return fail(ctx, RESTIR_E_UNSUPPORTED, "RESERVOIR_SIZE %u is not built (1, 2 and 4 are)", reservoir_size);
	}
	if (bandConnected(ctx)) {
		return fail(ctx, RESTIR_E_UNSUPPORTED, "restir_set_reservoir_variant on connected bands: the halo kernels move packed reservoirs");
	}
	const bool was = ctx->generic();
	const size_t before = ctx->reservoirBytes();
	ctx->variantN = (int)reservoir_size;
	ctx->variantMis = unbiased_mis != 0;
	ctx->fusedPasses = fused_passes != 0;
	if (ctx->band.W != 0 && (was != ctx->generic() || before != ctx->reservoirBytes())) {
		// the three buffers change their record: reallocated and zero-filled like a resize (bound G-buffers are kept)
		CU(ctx, cudaStreamSynchronize(ctx->stream));
		for (auto &r : ctx->reservoirs) freeDev(r);
		for (auto &r : ctx->genericReservoirs) freeDev(r);
		freeDev(ctx->staging);
		ctx->stagingPixels = 0;
		const size_t bytes = ctx->allocPixels() * (ctx->generic() ? ctx->reservoirBytes() : sizeof(PackedReservoir));
		for (int b = 0; b < 3; ++b) {
			void *ptr = nullptr;
			CU(ctx, cudaMalloc(&ptr, bytes));
			CU(ctx, cudaMemsetAsync(ptr, 0, bytes, ctx->stream));
			if (ctx->generic()) ctx->genericReservoirs[b] = static_cast<unsigned char *>(ptr); else ctx->reservoirs[b] = static_cast<PackedReservoir *>(ptr);
		}
		CU(ctx, cudaStreamSynchronize(ctx->stream));
	}
	return RESTIR_OK;
}

int restir_get_reservoir_bytes(const restir_context *ctx, size_t *bytes) {
	if (ctx == nullptr || bytes == nullptr) {
		return RESTIR_E_INVALID;
	}
	*bytes = ctx->reservoirBytes();
	return RESTIR_OK;
}

int restir_set_unbiased_neighbors(restir_context *ctx, uint32_t count) {
	ENTER(ctx);
	if (count < 1 || count > 16) {
		return fail(ctx, RESTIR_E_INVALID, "unbiased neighbour count must be in 1..16");
	}
	ctx->unbiasedNeighbors = count;
	if (bandConnected(ctx)) { // see restir_band_connect
		return ensureHandOver(ctx, pass_grid(ctx->band), count + 1, count);
	}
	return RESTIR_OK;
}

int restir_set_traversal(restir_context *ctx, int mode) {
	ENTER(ctx);
	if (mode != RESTIR_TRAVERSAL_AUTO && mode != RESTIR_TRAVERSAL_REFERENCE_ORDER && mode != RESTIR_TRAVERSAL_IMAGE && mode != RESTIR_TRAVERSAL_WIDE) {
		return fail(ctx, RESTIR_E_INVALID, "unknown traversal mode %d", mode);
	}
	ctx->traversal = mode;
	return RESTIR_OK;
}

int restir_band_local_peer(restir_context *ctx, restir_band_peer *out) {
	ENTER(ctx);
	if (out == nullptr || ctx->band.W == 0) {
		return fail(ctx, RESTIR_E_INVALID, "restir_band_local_peer: null result or restir_resize_band not called");
	}
	if (ctx->generic()) {
		return fail(ctx, RESTIR_E_UNSUPPORTED, "restir_band_local_peer: the halo kernels move packed reservoirs, this context holds a reservoir variant");
	}
	for (int b = 0; b < 3; ++b) out->reservoirs[b] = ctx->reservoirs[b];
	out->flags = ctx->bandFlags;
	out->alloc_begin = (uint32_t)ctx->band.allocBegin;
	out->alloc_end = (uint32_t)ctx->band.allocEnd;
	out->row_begin = (uint32_t)ctx->band.rowBegin;
	out->row_end = (uint32_t)ctx->band.rowEnd;
	return RESTIR_OK;
}

int restir_band_export_ipc(restir_context *ctx, restir_band_ipc *out) {
	ENTER(ctx);
	if (out == nullptr || ctx->band.W == 0) {
		return fail(ctx, RESTIR_E_INVALID, "restir_band_export_ipc: null result or restir_resize_band not called");
	}
	if (ctx->generic()) {
		return fail(ctx, RESTIR_E_UNSUPPORTED, "restir_band_export_ipc: the halo kernels move packed reservoirs, this context holds a reservoir variant");
	}
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
	std::memset(out, 0, sizeof(*out));
	for (int b = 0; b < 3; ++b) {
		cudaIpcMemHandle_t h;
		CU(ctx, cudaIpcGetMemHandle(&h, ctx->reservoirs[b]));
		std::memcpy(out->reservoirs[b], &h, 64);
	}
	cudaIpcMemHandle_t h;
	CU(ctx, cudaIpcGetMemHandle(&h, ctx->bandFlags));
	std::memcpy(out->flags, &h, 64);
	out->alloc_begin = (uint32_t)ctx->band.allocBegin;
	out->alloc_end = (uint32_t)ctx->band.allocEnd;
	out->row_begin = (uint32_t)ctx->band.rowBegin;
	out->row_end = (uint32_t)ctx->band.rowEnd;
	return RESTIR_OK;
}

int restir_band_open_ipc(restir_context *ctx, const restir_band_ipc *in, restir_band_peer *out) {
	ENTER(ctx);
	if (in == nullptr || out == nullptr) {
		return fail(ctx, RESTIR_E_INVALID, "restir_band_open_ipc: null argument");
	}
	std::memset(out, 0, sizeof(*out));
	for (int b = 0; b < 4; ++b) {
		cudaIpcMemHandle_t h;
		std::memcpy(&h, b < 3 ? in->reservoirs[b] : in->flags, 64);
		void *p = nullptr;
		CU(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
		ctx->ipcOpened.push_back(p);
		if (b < 3) out->reservoirs[b] = p; else out->flags = p;
	}
	out->alloc_begin = in->alloc_begin;
	out->alloc_end = in->alloc_end;
	out->row_begin = in->row_begin;
	out->row_end = in->row_end;
	return RESTIR_OK;
}

int restir_band_connect(restir_context *ctx, int side, const restir_band_peer *peer) {
	ENTER(ctx);
	if (side < 0 || side > 1 || ctx->band.W == 0) {
		return fail(ctx, RESTIR_E_INVALID, "restir_band_connect: bad side or restir_resize_band not called");
	}
	auto &p = ctx->peers[side];
	p = restir_context::PeerSide{};
	if (peer == nullptr) {
		return RESTIR_OK;
	}
	if (ctx->generic()) {
		return fail(ctx, RESTIR_E_UNSUPPORTED, "restir_band_connect: the halo kernels move packed reservoirs; a context with a reservoir variant "
		                                       "(restir_set_reservoir_variant) exchanges its halo rows through restir_reservoir_device_ptr");
	}
	if (!peer->reservoirs[0] || !peer->reservoirs[1] || !peer->reservoirs[2] || !peer->flags) {
		return fail(ctx, RESTIR_E_INVALID, "restir_band_connect: incomplete peer");
	}
	// the neighbour must hold rows of this band as halo, on the right side
	const bool touches = side == 0 ? ((int)peer->alloc_end > ctx->band.rowBegin && (int)peer->alloc_begin < ctx->band.rowBegin)
	                               : ((int)peer->alloc_begin < ctx->band.rowEnd && (int)peer->alloc_end > ctx->band.rowEnd);
	if (!touches) {
		return fail(ctx, RESTIR_E_INVALID, "restir_band_connect: peer rows [%u,%u) do not overlap this band's edge on side %d", peer->alloc_begin,
		            peer->alloc_end, side);
	}
	// ... and own every halo row this band keeps on that side: a neighbour pushes only rows it shades itself, so rows of a
	// band two hops away would stay at their zero fill and be gathered as empty reservoirs without any halo_miss
	const bool covers = side == 0 ? ((int)peer->row_end == ctx->band.rowBegin && (int)peer->row_begin <= ctx->band.allocBegin)
	                              : ((int)peer->row_begin == ctx->band.rowEnd && (int)peer->row_end >= ctx->band.allocEnd);
	if (!covers) {
		return fail(ctx, RESTIR_E_INVALID,
		            "restir_band_connect: the neighbour on side %d owns rows [%u,%u), which do not cover this band's halo rows [%d,%d): the halo must "
		            "not be taller than the neighbouring band",
		            side, peer->row_begin, peer->row_end, side == 0 ? ctx->band.allocBegin : ctx->band.rowEnd,
		            side == 0 ? ctx->band.rowBegin : ctx->band.allocEnd);
	}
	for (int b = 0; b < 3; ++b) p.reservoirs[b] = static_cast<PackedReservoir *>(peer->reservoirs[b]);
	p.flags = static_cast<unsigned long long *>(peer->flags);
	p.allocBegin = (int)peer->alloc_begin;
	p.allocEnd = (int)peer->alloc_end;
	p.rowBegin = (int)peer->row_begin;
	p.rowEnd = (int)peer->row_end;
	p.connected = true;
	// the hand-over buffers of the cut passes are sized now: growing them later would synchronise this stream in the middle of
	// a frame, possibly behind a wait for a neighbour that the same host thread has not driven yet
	return ensureHandOver(ctx, pass_grid(ctx->band), ctx->unbiasedNeighbors + 1, ctx->unbiasedNeighbors);
}

int restir_set_spatial_staging(restir_context *ctx, int enable) {
	ENTER(ctx);
	ctx->spatialStaging = enable != 0;
	return RESTIR_OK;
}

int restir_set_ray_elision(restir_context *ctx, int enable) {
	ENTER(ctx);
	ctx->rayElision = enable == 2 ? 2 : enable ? 1 : 0;
	return RESTIR_OK;
}

int restir_set_occluder_cache(restir_context *ctx, int enable) {
	ENTER(ctx);
	ctx->occluderCache = enable >= 0 && enable <= 3 ? enable : 1;
	return clearOccluders(ctx);
}

int restir_get_bvh_info(const restir_context *ctx, restir_bvh_info *out) {
	if (ctx == nullptr || out == nullptr) {
		return RESTIR_E_INVALID;
	}
	std::memset(out, 0, sizeof(*out));
	out->nodes = ctx->nNodes;
	out->triangles = ctx->nTris;
	out->reachable_nodes = ctx->imageInfo.reachableNodes;
	out->depth = (uint32_t)ctx->imageInfo.depth;
	out->reference_stack_bound = (uint32_t)ctx->imageInfo.referenceStackBound;
	out->traversal = ctx->wide ? RESTIR_TRAVERSAL_WIDE : ctx->image ? RESTIR_TRAVERSAL_IMAGE : RESTIR_TRAVERSAL_REFERENCE_ORDER;
	out->wide_nodes = ctx->wideInfo.nodes;
	out->wide_depth = (uint32_t)ctx->wideInfo.depth;
	out->wide_stack_bound = (uint32_t)ctx->wideInfo.stackBound;
	return RESTIR_OK;
}

int restir_check_aabb_tree(const void *nodes, uint32_t n_nodes, uint32_t n_triangles, restir_bvh_info *out, char *message, size_t message_bytes) {
	if (message && message_bytes) message[0] = 0;
	if (nodes == nullptr || n_nodes == 0 || n_triangles == 0) {
		return RESTIR_E_INVALID;
	}
	std::vector<Node64> image;
	TraversalImageInfo info;
	std::string why;
	bool ok = build_traversal_image(static_cast<const restir_aabb_node *>(nodes), n_nodes, n_triangles, image, info, why);
	const std::string &text = ok ? info.why : why;
	if (message && message_bytes) {
		std::strncpy(message, text.c_str(), message_bytes - 1);
		message[message_bytes - 1] = 0;
	}
	if (out) {
		std::memset(out, 0, sizeof(*out));
		out->nodes = n_nodes;
		out->triangles = n_triangles;
		out->reachable_nodes = info.reachableNodes;
		out->depth = (uint32_t)info.depth;
		out->reference_stack_bound = (uint32_t)info.referenceStackBound;
		out->traversal = info.usable ? RESTIR_TRAVERSAL_IMAGE : RESTIR_TRAVERSAL_REFERENCE_ORDER;
		if (ok && info.usable) { // would restir_upload_bvh walk the 4-wide image?  (wide_image.h: nested finite boxes, one leaf per triangle)
			std::vector<WideNode> wide;
			std::vector<uint32_t> triOrder;
			std::vector<float> leafBoxes;
			WideGrid grid{};
			WideImageInfo wi;
			build_wide_image(static_cast<const restir_aabb_node *>(nodes), n_nodes, n_triangles, wide, triOrder, leafBoxes, grid, wi);
			if (wi.usable) {
				out->traversal = RESTIR_TRAVERSAL_WIDE;
				out->wide_nodes = wi.nodes;
				out->wide_depth = (uint32_t)wi.depth;
				out->wide_stack_bound = (uint32_t)wi.stackBound;
			} else if (message && message_bytes && text.empty()) {
				std::strncpy(message, wi.why.c_str(), message_bytes - 1);
				message[message_bytes - 1] = 0;
			}
		}
	}
	return ok ? RESTIR_OK : RESTIR_E_INVALID;
}

int restir_get_wide_image(restir_context *ctx, void *wide_nodes, uint32_t capacity, uint32_t *n_wide, uint32_t *tri_order) {
	ENTER(ctx);
	if (n_wide) *n_wide = ctx->wide ? ctx->wideInfo.nodes : 0;
	if (ctx->wide == nullptr) {
		return fail(ctx, RESTIR_E_UNSUPPORTED, "the tree is not walked wide: %s", ctx->wideInfo.why.c_str());
	}
	CU(ctx, cudaStreamSynchronize(ctx->stream));
	if (wide_nodes != nullptr) {
		if (capacity < ctx->wideInfo.nodes) {
			return fail(ctx, RESTIR_E_INVALID, "restir_get_wide_image: room for %u nodes, the image has %u", capacity, ctx->wideInfo.nodes);
		}
		CU(ctx, cudaMemcpy(wide_nodes, ctx->wide, (size_t)ctx->wideInfo.nodes * sizeof(WideNode), cudaMemcpyDeviceToHost));
	}
	if (tri_order != nullptr) { // recovered from the records: record r holds the triangle whose p1 ... simpler: kept next to the image
		if (ctx->wideOrder == nullptr) {
			return fail(ctx, RESTIR_E_UNSUPPORTED, "restir_get_wide_image: record order not kept");
		}
		CU(ctx, cudaMemcpy(tri_order, ctx->wideOrder, (size_t)ctx->nTris * sizeof(uint32_t), cudaMemcpyDeviceToHost));
	}
	return RESTIR_OK;
}

int restir_check_wide_walk(const void *nodes, uint32_t n_nodes, const void *triangles, uint32_t n_triangles, const float *p1, const float *p2,
                           uint64_t n, unsigned char *shadowed, unsigned char *walked_wide, uint64_t *visits, char *message, size_t message_bytes) {
	if (message && message_bytes) message[0] = 0;
	if (nodes == nullptr || triangles == nullptr || n_nodes == 0 || n_triangles == 0 || (n != 0 && (p1 == nullptr || p2 == nullptr || shadowed == nullptr || walked_wide == nullptr))) {
		return RESTIR_E_INVALID;
	}
	std::vector<Node64> image;
	TraversalImageInfo info;
	std::string why;
	if (!build_traversal_image(static_cast<const restir_aabb_node *>(nodes), n_nodes, n_triangles, image, info, why)) {
		if (message && message_bytes) {
			std::strncpy(message, why.c_str(), message_bytes - 1);
			message[message_bytes - 1] = 0;
		}
		return RESTIR_E_INVALID;
	}
	if (!wide_walk_host(static_cast<const restir_aabb_node *>(nodes), n_nodes, static_cast<const restir_triangle *>(triangles), n_triangles, p1, p2, n, shadowed,
	                    walked_wide, visits, why)) {
		if (message && message_bytes) {
			std::strncpy(message, why.c_str(), message_bytes - 1);
			message[message_bytes - 1] = 0;
		}
		return RESTIR_E_UNSUPPORTED;
	}
	return RESTIR_OK;
}

int restir_pass_restir(restir_context *ctx, int gbuffer, int out_buffer, int prev_buffer) {
	ENTER(ctx);
	PassParams p;
	int rc;
	if ((rc = makeParams(ctx, gbuffer, true, true, p)) != RESTIR_OK) return rc;
	if ((rc = checkBuffer(ctx, out_buffer)) != RESTIR_OK) return rc;
	if ((rc = checkBuffer(ctx, prev_buffer)) != RESTIR_OK) return rc;
	if (out_buffer == prev_buffer) {
		return fail(ctx, RESTIR_E_INVALID, "restir pass: out and prev buffers must differ");
	}
	if (ctx->generic()) { // one kernel, rays traced inline (restir_generic.cu)
		if ((rc = waitUpload(ctx, gbuffer)) != RESTIR_OK) return rc;
		if ((p.u.flags & RESTIR_TEMPORAL_REUSE_FLAG) != 0 && (rc = waitUpload(ctx, gbuffer ^ 1)) != RESTIR_OK) return rc;
		beforeLaunch(ctx, "generic_restir_kernel");
		launch_generic_restir(ctx->variantN, ctx->variantMis, p, ctx->genericReservoirs[out_buffer], ctx->genericReservoirs[prev_buffer], ctx->stream);
		if ((rc = afterLaunch(ctx, "generic_restir_kernel")) != RESTIR_OK) return rc;
		if ((rc = markReadIfOwned(ctx, gbuffer)) != RESTIR_OK) return rc;
		return markReadIfOwned(ctx, gbuffer ^ 1);
	}
	// restirOmni.glsl cut at its testVisibility call (:148-160): candidates | shadow rays | visibility + temporal
	const PassGrid g = pass_grid(ctx->band);
	const bool vis = (p.u.flags & RESTIR_VISIBILITY_REUSE_FLAG) != 0, temporal = (p.u.flags & RESTIR_TEMPORAL_REUSE_FLAG) != 0;
	if ((rc = ensureHandOver(ctx, g, 1, 0)) != RESTIR_OK) return rc;
	if ((rc = waitUpload(ctx, gbuffer)) != RESTIR_OK) return rc;
	if (temporal && (rc = waitUpload(ctx, gbuffer ^ 1)) != RESTIR_OK) return rc;
	// connected bands: the neighbours' rows of the previous frame's reservoirs (temporal reprojection reads them; waiting
	// even when temporal reuse is off keeps a fast rank from overwriting a halo its neighbour is still reading)
	if ((rc = haloWait(ctx, prev_buffer)) != RESTIR_OK) return rc;
	PackedReservoir *out = ctx->reservoirs[out_buffer];
	beforeLaunch(ctx, "omni_candidates_kernel");
	launch_omni_candidates(p, out, ctx->stream);
	if ((rc = afterLaunch(ctx, "omni_candidates_kernel")) != RESTIR_OK) return rc;
	if (vis) {
		TraceParams tp = traceParams(ctx);
		tp.tilesX = g.tilesX;
		tp.slots = 1;
		tp.outStride = 1;
		tp.outOffset = 0;
		tp.nItems = g.pixelIds;
		tp.worldPos = p.cur.worldPos;
		tp.reservoirs = out;
		beforeLaunch(ctx, "trace_kernel<pixel>");
		CU(ctx, launch_trace(tp, kTracePixel, ctx->smCount, ctx->stream));
		if ((rc = afterLaunch(ctx, "trace_kernel<pixel>")) != RESTIR_OK) return rc;
	}
	if (vis || temporal) {
		beforeLaunch(ctx, "omni_temporal_kernel");
		launch_omni_temporal(p, out, ctx->reservoirs[prev_buffer], ctx->shadowed, ctx->stream);
		if ((rc = afterLaunch(ctx, "omni_temporal_kernel")) != RESTIR_OK) return rc;
	}
	if ((rc = haloPush(ctx, out_buffer)) != RESTIR_OK) return rc;
	if ((rc = markReadIfOwned(ctx, gbuffer)) != RESTIR_OK) return rc;
	return markReadIfOwned(ctx, gbuffer ^ 1); // the previous frame's G-buffer is not read after this pass
}

namespace {
// lighting output of a pass: checked like restir_pass_lighting's
int checkLightingTarget(restir_context *ctx, const void *out_device, int out_format) {
	if (!ctx->haveLighting) {
		return fail(ctx, RESTIR_E_INVALID, "restir_set_lighting_uniforms has not been called");
	}
	if (out_device == nullptr || (out_format != RESTIR_OUT_RGBA32F && out_format != RESTIR_OUT_RGBA8_SRGB)) {
		return fail(ctx, RESTIR_E_INVALID, "lighting pass: bad output");
	}
	if ((int)ctx->lighting.bufferSize[0] != ctx->band.W || (int)ctx->lighting.bufferSize[1] != ctx->band.H) {
		return fail(ctx, RESTIR_E_INVALID, "lighting uniforms bufferSize does not match restir_resize");
	}
	return RESTIR_OK;
}
int passSpatial(restir_context *ctx, int gbuffer, int in_buffer, int out_buffer, int iter, void *lit_out, int lit_format);
int passUnbiased(restir_context *ctx, int gbuffer, int in_buffer, int out_buffer, void *lit_out, int lit_format);
} // namespace

int restir_pass_spatial(restir_context *ctx, int gbuffer, int in_buffer, int out_buffer, int iter) {
	ENTER(ctx);
	return passSpatial(ctx, gbuffer, in_buffer, out_buffer, iter, nullptr, 0);
}

int restir_pass_unbiased(restir_context *ctx, int gbuffer, int in_buffer, int out_buffer) {
	ENTER(ctx);
	return passUnbiased(ctx, gbuffer, in_buffer, out_buffer, nullptr, 0);
}

namespace {

// lit_out != null: the lighting pass of the same pixels runs inside the pass's last kernel (restir_frame_lit)
int passSpatial(restir_context *ctx, int gbuffer, int in_buffer, int out_buffer, int iter, void *lit_out, int lit_format) {
	PassParams p;
	int rc;
	if ((rc = makeParams(ctx, gbuffer, false, true, p)) != RESTIR_OK) return rc;
	if ((rc = checkBuffer(ctx, in_buffer)) != RESTIR_OK) return rc;
	if ((rc = checkBuffer(ctx, out_buffer)) != RESTIR_OK) return rc;
	if (in_buffer == out_buffer) {
		return fail(ctx, RESTIR_E_INVALID, "spatial pass: in and out buffers must differ");
	}
	if ((rc = waitUpload(ctx, gbuffer)) != RESTIR_OK) return rc;
	if (lit_out && (rc = checkLightingTarget(ctx, lit_out, lit_format)) != RESTIR_OK) return rc;
	if (ctx->generic()) {
		beforeLaunch(ctx, "generic_spatial_kernel");
		launch_generic_spatial(ctx->variantN, ctx->variantMis, p, ctx->genericReservoirs[in_buffer], ctx->genericReservoirs[out_buffer], iter, ctx->stream);
		if ((rc = afterLaunch(ctx, "generic_spatial_kernel")) != RESTIR_OK) return rc;
		if (lit_out && (rc = restir_pass_lighting(ctx, gbuffer, out_buffer, lit_out, lit_format)) != RESTIR_OK) return rc;
		return markReadIfOwned(ctx, gbuffer);
	}
	if ((rc = haloWait(ctx, in_buffer)) != RESTIR_OK) return rc;
	const char *name = lit_out ? "spatial_reuse_kernel+lighting" : "spatial_reuse_kernel";
	beforeLaunch(ctx, name);
	launch_spatial_reuse(p, ctx->reservoirs[in_buffer], ctx->reservoirs[out_buffer], iter, lit_out ? &ctx->lighting : nullptr, lit_out, lit_format,
	                     ctx->spatialStaging, ctx->stream);
	if ((rc = afterLaunch(ctx, name)) != RESTIR_OK) return rc;
	if ((rc = haloPush(ctx, out_buffer)) != RESTIR_OK) return rc;
	return markReadIfOwned(ctx, gbuffer);
}

int passUnbiased(restir_context *ctx, int gbuffer, int in_buffer, int out_buffer, void *lit_out, int lit_format) {
	PassParams p;
	int rc;
	if ((rc = makeParams(ctx, gbuffer, true, true, p)) != RESTIR_OK) return rc;
	if ((rc = checkBuffer(ctx, in_buffer)) != RESTIR_OK) return rc;
	if ((rc = checkBuffer(ctx, out_buffer)) != RESTIR_OK) return rc;
	if (in_buffer == out_buffer) {
		return fail(ctx, RESTIR_E_INVALID, "unbiased pass: in and out buffers must differ");
	}
	if (ctx->generic()) { // one kernel, rays traced inline (restir_generic.cu)
		if ((rc = waitUpload(ctx, gbuffer)) != RESTIR_OK) return rc;
		if (lit_out && (rc = checkLightingTarget(ctx, lit_out, lit_format)) != RESTIR_OK) return rc;
		beforeLaunch(ctx, "generic_unbiased_kernel");
		launch_generic_unbiased(ctx->variantN, ctx->variantMis, p, ctx->genericReservoirs[in_buffer], ctx->genericReservoirs[out_buffer],
		                        (int)ctx->unbiasedNeighbors, ctx->stream);
		if ((rc = afterLaunch(ctx, "generic_unbiased_kernel")) != RESTIR_OK) return rc;
		if (lit_out && (rc = restir_pass_lighting(ctx, gbuffer, out_buffer, lit_out, lit_format)) != RESTIR_OK) return rc;
		return markReadIfOwned(ctx, gbuffer);
	}
	// unbiasedReuse.glsl cut at its testVisibility calls (:139-166): merge | shadow rays | normalisation
	const PassGrid g = pass_grid(ctx->band);
	const unsigned k = ctx->unbiasedNeighbors;
	const bool vis = (p.u.flags & RESTIR_VISIBILITY_REUSE_FLAG) != 0;
	if (g.pixelIds * (k + 1) >= (1ull << 32)) {
		return fail(ctx, RESTIR_E_UNSUPPORTED, "unbiased pass: %llu rays exceed the trace kernel's 32-bit work list; use row bands", (unsigned long long)(g.pixelIds * (k + 1)));
	}
	if ((rc = ensureHandOver(ctx, g, k + 1, k)) != RESTIR_OK) return rc;
	if ((rc = waitUpload(ctx, gbuffer)) != RESTIR_OK) return rc;
	if ((rc = haloWait(ctx, in_buffer)) != RESTIR_OK) return rc;
	if (lit_out && (rc = checkLightingTarget(ctx, lit_out, lit_format)) != RESTIR_OK) return rc;
	const PackedReservoir *in = ctx->reservoirs[in_buffer];
	PackedReservoir *out = ctx->reservoirs[out_buffer];
	beforeLaunch(ctx, "unbiased_merge_kernel");
	launch_unbiased_merge(p, in, out, (int)k, ctx->neighborPix, ctx->neighborM, ctx->stream);
	if ((rc = afterLaunch(ctx, "unbiased_merge_kernel")) != RESTIR_OK) return rc;
	if (vis) {
		// the pixels' own rays first (:157-166): a shadowed pixel needs none of its neighbour rays, and a neighbour ray
		// whose segment is bit-identical to the neighbour's own ray is answered from it (restir_trace.cu item_resolve)
		TraceParams tp = traceParams(ctx);
		tp.tilesX = g.tilesX;
		tp.slots = 1;
		tp.outStride = k + 1;
		tp.outOffset = k;
		tp.nItems = g.pixelIds;
		tp.worldPos = p.cur.worldPos;
		tp.reservoirs = out;
		// the unbiased pass's rays are mostly unshadowed (89 % / 81 % on Sponza): with entries chosen by direction the key needs the
		// segment of every item, and the pretest costs those two kernels more than its few witnesses save (8K / 1 M lights: own
		// rays 5.58 -> 5.75 ms, neighbour rays 23.2 -> 24.0) — restirOmni's rays keep it, these walks still record what they find
		tp.occluderPretest = tp.occluderByDirection ? 0 : 1;
		beforeLaunch(ctx, "trace_kernel<own>");
		CU(ctx, launch_trace(tp, kTracePixel, ctx->smCount, ctx->stream));
		if ((rc = afterLaunch(ctx, "trace_kernel<own>")) != RESTIR_OK) return rc;
		tp.slots = k;
		tp.elide = ctx->rayElision != 0;
		if (ctx->rayElision == 2 && ctx->dedupe != nullptr) { // experiment: one walk per distinct segment (restir_trace.cu segment_claim)
			tp.dedupe = ctx->dedupe;
			tp.dedupeMask = (unsigned)(ctx->dedupeEntries - 1);
			tp.aliases = ctx->aliases;
			tp.aliasCapacity = (unsigned)std::min<size_t>(ctx->aliasCapacity, 0xffffffffu);
			tp.aliasCount = ctx->aliasCounter;
		}
		tp.nItems = g.pixelIds * k;
		tp.neighborPix = ctx->neighborPix;
#if !RESTIR_NEIGHBOUR_PRETEST
		tp.occluderPretest = 0;
#endif
		beforeLaunch(ctx, "trace_kernel<neighbours>");
		CU(ctx, launch_trace(tp, kTraceUnbiased, ctx->smCount, ctx->stream));
		if ((rc = afterLaunch(ctx, "trace_kernel<neighbours>")) != RESTIR_OK) return rc;
	}
	const char *name = lit_out ? "unbiased_finalize_kernel+lighting" : "unbiased_finalize_kernel";
	beforeLaunch(ctx, name);
	launch_unbiased_finalize(p, out, (int)k, ctx->neighborPix, ctx->neighborM, ctx->shadowed, lit_out ? &ctx->lighting : nullptr, lit_out, lit_format,
	                         ctx->stream);
	if ((rc = afterLaunch(ctx, name)) != RESTIR_OK) return rc;
	if ((rc = haloPush(ctx, out_buffer)) != RESTIR_OK) return rc;
	return markReadIfOwned(ctx, gbuffer);
}

} // namespace

int restir_pass_lighting(restir_context *ctx, int gbuffer, int buffer, void *out_device, int out_format) {
	ENTER(ctx);
	PassParams p;
	int rc;
	if ((rc = makeParams(ctx, gbuffer, false, true, p)) != RESTIR_OK) return rc;
	if ((rc = checkBuffer(ctx, buffer)) != RESTIR_OK) return rc;
	if ((rc = checkLightingTarget(ctx, out_device, out_format)) != RESTIR_OK) return rc;
	if ((rc = waitUpload(ctx, gbuffer)) != RESTIR_OK) return rc;
	if (ctx->generic()) {
		beforeLaunch(ctx, "generic_lighting_kernel");
		launch_generic_lighting(ctx->variantN, ctx->variantMis, p, ctx->lighting, ctx->genericReservoirs[buffer], out_device, out_format, ctx->stream);
		if ((rc = afterLaunch(ctx, "generic_lighting_kernel")) != RESTIR_OK) return rc;
		return markReadIfOwned(ctx, gbuffer);
	}
	beforeLaunch(ctx, "lighting_kernel");
	launch_lighting(p, ctx->lighting, ctx->reservoirs[buffer], out_device, out_format, ctx->stream);
	if ((rc = afterLaunch(ctx, "lighting_kernel")) != RESTIR_OK) return rc;
	return markReadIfOwned(ctx, gbuffer);
}

static int frameImpl(restir_context *ctx, int i, int unbiased, int spatial_iterations, void *lit_out, int lit_format) {
	if (i < 0 || i > 1 || spatial_iterations < 0) {
		return fail(ctx, RESTIR_E_INVALID, "restir_frame: bad frame index");
	}
	const bool edgeless = (ctx->band.rowBegin == 0 || ctx->peers[0].connected) && (ctx->band.rowEnd == ctx->band.H || ctx->peers[1].connected);
	if (!edgeless) {
		return fail(ctx, RESTIR_E_INVALID, "restir_frame on a band context needs its neighbours connected (restir_band_connect); otherwise interleave the halo copies yourself");
	}
	int rc;
	if (lit_out && (rc = checkLightingTarget(ctx, lit_out, lit_format)) != RESTIR_OK) return rc;
	const int cur = i, prev = i ^ 1; // app.h:298-332
	if (unbiased) {
		if ((rc = restir_pass_restir(ctx, i, RESTIR_BUF_TEMP, prev)) != RESTIR_OK) return rc;
		return passUnbiased(ctx, i, RESTIR_BUF_TEMP, cur, lit_out, lit_format);
	}
	if ((rc = restir_pass_restir(ctx, i, cur, prev)) != RESTIR_OK) return rc;
	for (int j = 0; j < spatial_iterations; ++j) {
		const bool last = j + 1 == spatial_iterations;
		if ((rc = passSpatial(ctx, i, cur, prev, j * 2, nullptr, 0)) != RESTIR_OK) return rc;
		if ((rc = passSpatial(ctx, i, prev, cur, j * 2 + 1, last ? lit_out : nullptr, lit_format)) != RESTIR_OK) return rc;
	}
	if (lit_out && spatial_iterations == 0) { // nothing to fuse the lighting into
		return restir_pass_lighting(ctx, i, cur, lit_out, lit_format);
	}
	return RESTIR_OK;
}

int restir_frame(restir_context *ctx, int i, int unbiased, int spatial_iterations) {
	ENTER(ctx);
	return frameImpl(ctx, i, unbiased, spatial_iterations, nullptr, 0);
}

int restir_frame_lit(restir_context *ctx, int i, int unbiased, int spatial_iterations, void *out_device, int out_format) {
	ENTER(ctx);
	if (out_device == nullptr) {
		return fail(ctx, RESTIR_E_INVALID, "restir_frame_lit: null output");
	}
	return frameImpl(ctx, i, unbiased, spatial_iterations, out_device, out_format);
}

static int ensureStaging(restir_context *ctx) {
	size_t px = ctx->allocPixels();
	if (ctx->stagingPixels < px) {
		freeDev(ctx->staging);
		ctx->stagingPixels = 0;
		CU(ctx, cudaMalloc(&ctx->staging, px * sizeof(restir_reservoir)));
		ctx->stagingPixels = px;
	}
	return RESTIR_OK;
}

int restir_download_reservoirs(restir_context *ctx, int buffer, restir_reservoir *dst_host) {
	ENTER(ctx);
	int rc;
	if ((rc = checkBuffer(ctx, buffer)) != RESTIR_OK) return rc;
	if (dst_host == nullptr) {
		return fail(ctx, RESTIR_E_INVALID, "null destination");
	}
	if (ctx->generic()) { // already the reference's records
		CU(ctx, cudaMemcpyAsync(dst_host, ctx->genericReservoirs[buffer], ctx->allocPixels() * ctx->reservoirBytes(), cudaMemcpyDeviceToHost, ctx->stream));
		CU(ctx, cudaStreamSynchronize(ctx->stream));
		return RESTIR_OK;
	}
	if ((rc = ensureStaging(ctx)) != RESTIR_OK) return rc;
	size_t px = ctx->allocPixels();
	launch_unpack_reservoirs(sceneView(ctx), ctx->reservoirs[buffer], ctx->staging, px, ctx->stream);
	if ((rc = afterLaunch(ctx, "unpack_reservoirs_kernel")) != RESTIR_OK) return rc;
	CU(ctx, cudaMemcpyAsync(dst_host, ctx->staging, px * sizeof(restir_reservoir), cudaMemcpyDeviceToHost, ctx->stream));
	CU(ctx, cudaStreamSynchronize(ctx->stream));
	return RESTIR_OK;
}

int restir_upload_reservoirs(restir_context *ctx, int buffer, const restir_reservoir *src_host) {
	ENTER(ctx);
	int rc;
	if ((rc = checkBuffer(ctx, buffer)) != RESTIR_OK) return rc;
	if (src_host == nullptr) {
		return fail(ctx, RESTIR_E_INVALID, "null source");
	}
	if (ctx->generic()) {
		CU(ctx, cudaMemcpyAsync(ctx->genericReservoirs[buffer], src_host, ctx->allocPixels() * ctx->reservoirBytes(), cudaMemcpyHostToDevice, ctx->stream));
		CU(ctx, cudaStreamSynchronize(ctx->stream));
		return RESTIR_OK;
	}
	if ((rc = ensureStaging(ctx)) != RESTIR_OK) return rc;
	size_t px = ctx->allocPixels();
	CU(ctx, cudaMemcpyAsync(ctx->staging, src_host, px * sizeof(restir_reservoir), cudaMemcpyHostToDevice, ctx->stream));
	launch_pack_reservoirs(ctx->staging, ctx->reservoirs[buffer], px, ctx->stream);
	if ((rc = afterLaunch(ctx, "pack_reservoirs_kernel")) != RESTIR_OK) return rc;
	CU(ctx, cudaStreamSynchronize(ctx->stream));
	return RESTIR_OK;
}

int restir_reservoir_device_ptr(restir_context *ctx, int buffer, void **ptr, size_t *row_pitch_bytes) {
	ENTER(ctx);
	int rc;
	if ((rc = checkBuffer(ctx, buffer)) != RESTIR_OK) return rc;
	if (ptr) *ptr = ctx->generic() ? static_cast<void *>(ctx->genericReservoirs[buffer]) : static_cast<void *>(ctx->reservoirs[buffer]);
	if (row_pitch_bytes) *row_pitch_bytes = (size_t)ctx->band.W * (ctx->generic() ? ctx->reservoirBytes() : sizeof(PackedReservoir));
	return RESTIR_OK;
}

int restir_trace_segments(restir_context *ctx, const float *p1, const float *p2, uint64_t n, uint8_t *shadowed) {
	ENTER(ctx);
	if (ctx->nodes == nullptr) {
		return fail(ctx, RESTIR_E_INVALID, "restir_upload_bvh has not been called");
	}
	if (n != 0 && (p1 == nullptr || p2 == nullptr || shadowed == nullptr)) {
		return fail(ctx, RESTIR_E_INVALID, "null segment arrays");
	}
	if (n == 0) {
		return RESTIR_OK;
	}
	// the kernel numbers its work items with 32 bits: longer lists go in slices
	const uint64_t slice = 1ull << 31;
	for (uint64_t first = 0; first < n; first += slice) {
		TraceParams tp = traceParams(ctx);
		tp.nItems = n - first < slice ? n - first : slice;
		tp.segP1 = p1 + first * 3;
		tp.segP2 = p2 + first * 3;
		tp.shadowed = shadowed + first;
		beforeLaunch(ctx, "trace_kernel<segments>");
		CU(ctx, launch_trace(tp, kTraceSegments, ctx->smCount, ctx->stream));
		int rc = afterLaunch(ctx, "trace_kernel<segments>");
		if (rc != RESTIR_OK) return rc;
	}
	return RESTIR_OK;
}

int restir_get_counters(restir_context *ctx, restir_counters *out, int reset) {
	ENTER(ctx);
	unsigned long long h[kCounterCount];
	CU(ctx, cudaMemcpyAsync(h, ctx->counters, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
	CU(ctx, cudaStreamSynchronize(ctx->stream));
	if (out) {
		out->shadow_rays = h[kCounterRays];
		out->shadow_rays_traced = h[kCounterTraced];
		out->shadow_rays_cached = h[kCounterCached];
		out->halo_wait_timeouts = h[kCounterHaloTimeout];
		out->stack_overflows = h[kCounterOverflow];
		out->halo_misses = h[kCounterHaloMiss];
		out->kernel_launches = ctx->launches;
	}
	if (reset) {
		CU(ctx, cudaMemsetAsync(ctx->counters, 0, sizeof(h), ctx->stream));
		ctx->launches = 0;
	}
	if (h[kCounterHaloTimeout] != 0) {
		return fail(ctx, RESTIR_E_HALO, "%llu wait(s) for a neighbour's halo rows timed out: the passes after them read stale rows", h[kCounterHaloTimeout]);
	}
	return RESTIR_OK;
}

int restir_profile_begin(restir_context *ctx) {
	ENTER(ctx);
	dropProfile(ctx);
	ctx->profiling = true;
	return RESTIR_OK;
}

int restir_profile_end(restir_context *ctx, restir_kernel_time *out, uint32_t capacity, uint32_t *count) {
	ENTER(ctx);
	ctx->profiling = false;
	CU(ctx, cudaStreamSynchronize(ctx->stream));
	uint32_t n = 0;
	for (auto &r : ctx->prof) {
		float ms = 0.0f;
		if (cudaEventElapsedTime(&ms, r.begin, r.end) != cudaSuccess) {
			continue;
		}
		uint32_t k = 0;
		while (k < n && std::strcmp(out[k].name, r.name) != 0) ++k;
		if (k == n) {
			if (n == capacity) continue;
			std::memset(&out[n], 0, sizeof(out[n]));
			std::strncpy(out[n].name, r.name, sizeof(out[n].name) - 1);
			++n;
		}
		out[k].launches++;
		out[k].total_ms += ms;
	}
	dropProfile(ctx);
	if (count) *count = n;
	return RESTIR_OK;
}

// ---- fixture tool ------------------------------------------------------------------------------

int restir_camera_matrix(const restir_camera *c, float out_pv[16]) {
	if (c == nullptr || out_pv == nullptr) {
		return RESTIR_E_INVALID;
	}
	H3 pos, fwd, right, up;
	cameraBasis(c, pos, fwd, right, up);
	H3 r0 = right, r1 = H3{-up.x, -up.y, -up.z}, r2 = fwd; // camera.h:30-33: row 1 carries the y flip
	float view[4][4] = {{r0.x, r0.y, r0.z, -hdot(r0, pos)}, {r1.x, r1.y, r1.z, -hdot(r1, pos)}, {r2.x, r2.y, r2.z, -hdot(r2, pos)}, {0, 0, 0, 1}};
	float f = 1.0f / tanf(0.5f * c->fovYRadians); // camera.h:41-47
	float proj[4][4] = {{f / c->aspectRatio, 0, 0, 0},
	                    {0, f, 0, 0},
	                    {0, 0, -c->zFar / (c->zNear - c->zFar), c->zNear * c->zFar / (c->zNear - c->zFar)},
	                    {0, 0, 1, 0}};
	for (int r = 0; r < 4; ++r) {
		for (int col = 0; col < 4; ++col) {
			float acc = 0.0f;
			for (int k = 0; k < 4; ++k) {
				acc = acc + proj[r][k] * view[k][col];
			}
			out_pv[col * 4 + r] = acc;
		}
	}
	return RESTIR_OK;
}

int restir_tools_raycast_gbuffer(restir_context *ctx, const restir_camera *camera, const int32_t *tri_material_device,
                                 const uint32_t *material_table_device, void *albedo, void *normal, void *material, void *worldPos,
                                 void *depth) {
	ENTER(ctx);
	if (ctx->nodes == nullptr || ctx->band.W == 0) {
		return fail(ctx, RESTIR_E_INVALID, "raycast: upload the BVH and call restir_resize first");
	}
	if (!camera || !tri_material_device || !material_table_device || !albedo || !normal || !material || !worldPos || !depth) {
		return fail(ctx, RESTIR_E_INVALID, "raycast: null argument");
	}
	RaycastCamera rc;
	cameraBasisOf(camera, rc);
	Band full = ctx->band; // the fixture covers every row the context holds, halo included
	full.rowBegin = full.allocBegin;
	full.rowEnd = full.allocEnd;
	launch_raycast_gbuffer(sceneView(ctx), full, rc, tri_material_device, reinterpret_cast<const uint4 *>(material_table_device), albedo, normal,
	                       material, worldPos, depth, ctx->stream);
	return afterLaunch(ctx, "raycast_gbuffer_kernel");
}

} // extern "C"
