// restir_kernels.h — launch wrappers implemented in restir_kernels.cu (internal to the library).
#pragma once

#include "restir_device.cuh"
#include "wide_image.h"

namespace restir {

// host-precomputed camera basis for the G-buffer fixture tool (twin of oracle_raycast_gbuffer's preamble)
struct RaycastCamera {
	float pos[3], fwd[3], right[3], up[3];
	float sx, sy;
	float pv[16];
};

// ---- the persistent shadow-ray kernel (restir_trace.cu) ---------------------------------------------------
enum TraceMode {
	kTracePixel = 0,    // one ray per pixel: G-buffer position -> reservoir sample (restirOmni.glsl:148-160, unbiasedReuse.glsl:157-166)
	kTraceUnbiased = 1, // `slots` neighbour rays per pixel: neighbour position -> the pixel's sample (unbiasedReuse.glsl:139-156)
	kTraceSegments = 2, // explicit segments (restir_trace_segments)
};
// n / d for any n < 2^32 as the high half of one 64-bit product: M = ceil(2^64 / d) (host, fast_div_make); d = 1 is M = 0.
// Exact: n M / 2^64 = n / d + n e / (d 2^64) with e = M d - 2^64 < d, and the excess stays below 2^-32 < 1 / d.
struct FastDiv {
	unsigned long long M;
	unsigned d;
};
inline FastDiv fast_div_make(unsigned d) {
	FastDiv f;
	f.d = d;
	f.M = d <= 1 ? 0ull : ~0ull / d + 1ull; // floor((2^64 - 1) / d) + 1 = ceil(2^64 / d) when d is not a power of two, and 2^64 / d when it is
	return f;
}

struct TraceParams {
	const float4 *nodes, *tris, *image; // image == null: literal walk of the 80-byte nodes (restir_trace.cuh)
	const float4 *triEdges;             // with image: 64-byte (p1, e1, e2) records derived from tris at upload (restir_trace.cuh)
	const uint4 *wide;                  // 4-wide quantised image of the same tree (wide_image.h, restir_wide.cuh); null: the binary image is walked
	WideGrid grid;
	unsigned *dedupe;                   // kTraceUnbiased: open-addressing table (segment -> first item that asked for it), restir_trace.cu; null: off
	unsigned dedupeMask;                // entries - 1 (a power of two)
	uint2 *aliases;                     // (visibility byte of a duplicate, visibility byte of the item that walks the segment)
	unsigned aliasCapacity;
	unsigned *aliasCount;               // device counter of `aliases`
	unsigned *occluders;                // occluder cache (restir_trace.cu): [screen region][256] -> tag << 24 | triangle record; null: off
	unsigned regionsX;                  // regions (64 x 32 pixels) per region row
	int occluderPretest;                // test the cached witness(es) before a ray is queued: 0 no, 1 the most recent one, 2 both ways (walks record witnesses either way)
	int occluderByDirection;            // entries chosen by the segment's direction instead of the light index (many lights)
	unsigned nTris;
	unsigned nNodes;
	FastDiv divW, divTilesX, divSlots;  // set by launch_trace
	Band band;
	unsigned tilesX;                   // 8x4 tiles per tile row of the pass grid (item numbering, see tile_pixel_id)
	unsigned slots;
	unsigned long long nItems;
	const float4 *worldPos;
	const PackedReservoir *reservoirs; // sample positions = ray targets
	const int *neighborPix;            // [pixel id][slots-1]: local pixel index of the neighbour, < 0 = no ray
	const float *segP1, *segP2;
	unsigned char *shadowed;           // 1 = shadowed.  kTracePixel: [pixel id * outStride + outOffset]; kTraceUnbiased:
	                                   // [pixel id * (slots + 1) + slot], the pixel's own ray (traced before, kTracePixel) at slot `slots`
	unsigned outStride, outOffset;
	int elide;                         // kTraceUnbiased: answer neighbour rays without a walk where that is exact (restir_trace.cu item_resolve)
	unsigned long long *counters;
	unsigned *regionCursors, *smSlots; // RESTIR_TRACE_AFFINE: one chunk cursor per region of the work list, one arrival counter per SM
	unsigned blocksPerSm;
	float guidedShare;                 // RESTIR_TRACE_GUIDED / warps of the grid (set by the launcher): the part of the remaining items one fetch takes
};
// witnesses kept per (screen region, key) of the occluder cache (restir_trace.cu), most recent first
#ifndef RESTIR_OCCLUDER_WAYS
#define RESTIR_OCCLUDER_WAYS 2
#endif
constexpr unsigned kOccluderWays = RESTIR_OCCLUDER_WAYS;
constexpr unsigned kTraceMaxRegions = 4096, kTraceMaxSms = 1024; // regionCursors holds kTraceMaxRegions + kTraceMaxSms words, smSlots = regionCursors + kTraceMaxRegions
cudaError_t launch_trace(const TraceParams &tp, int mode, int smCount, cudaStream_t s);

// ---- per-pixel kernels (restir_kernels.cu) ---------------------------------------------------------------
// Grid geometry shared by every per-pixel kernel and the trace kernel's item numbering.
struct PassGrid {
	unsigned gx, gy;     // CTAs (32x8 pixels each)
	unsigned tilesX;     // 8x4 tiles per tile row = 4 * gx
	unsigned long long pixelIds; // tile-ordered pixel ids = 256 * gx * gy (ids past the band edge are holes)
};
PassGrid pass_grid(const Band &b);

// restirOmni.glsl:108-142 (candidates) and :148-209 (apply visibility, temporal reuse); between the two the
// trace kernel runs in kTracePixel mode on `out`.
void launch_omni_candidates(const PassParams &p, PackedReservoir *out, cudaStream_t s);
void launch_omni_temporal(const PassParams &p, PackedReservoir *out, const PackedReservoir *prev, const unsigned char *shadowed, cudaStream_t s);
// lu != null: the lighting pass of the same pixels follows inside the kernel (restir_frame_lit), written to outPixels in format fmt
// staged: the gate data (depth, normal) of the tile and its apron go through shared memory first (spatial_reuse_staged_kernel)
void launch_spatial_reuse(const PassParams &p, const PackedReservoir *in, PackedReservoir *out, int iter, const restir_lighting_uniforms *lu,
                          void *outPixels, int fmt, bool staged, cudaStream_t s);
// unbiasedReuse.glsl:84-124 (merge) and :126-182 (normalisation from the visibility bits); the trace kernel
// runs in kTraceUnbiased mode between them.  neighborM: [pixel id][numNeighbors + 1] sample counts handed from the merge to the normalisation.
void launch_unbiased_merge(const PassParams &p, const PackedReservoir *in, PackedReservoir *out, int numNeighbors, int *neighborPix,
                           uint32_t *neighborM, cudaStream_t s);
void launch_unbiased_finalize(const PassParams &p, PackedReservoir *out, int numNeighbors, const int *neighborPix, const uint32_t *neighborM,
                              const unsigned char *shadowed, const restir_lighting_uniforms *lu, void *outPixels, int fmt, cudaStream_t s);
void launch_lighting(const PassParams &p, const restir_lighting_uniforms &lu, const PackedReservoir *res, void *out, int fmt, cudaStream_t s);
void launch_raycast_gbuffer(const SceneView &sc, const Band &band, const RaycastCamera &cam, const int *triMaterial, const uint4 *materialTable,
                            void *albedo, void *normal, void *material, void *worldPos, void *depth, cudaStream_t s);
void launch_unpack_reservoirs(const SceneView &sc, const PackedReservoir *in, restir_reservoir *out, size_t n, cudaStream_t s);
void launch_pack_reservoirs(const restir_reservoir *in, PackedReservoir *out, size_t n, cudaStream_t s);
void launch_derive_triangle_edges(const float4 *tris, uint32_t n, float4 *out, const uint32_t *order, const float *leafBoxes, cudaStream_t s);
void launch_derive_light_tables(const restir_point_light *pl, int np, float4 *pointOut, const restir_tri_light *tl, int nt, float4 *triOut,
                                cudaStream_t s);

// ---- the G-buffer pass (restir_gbuffer.cu) ------------------------------------------------------------------------
struct GBufferScene {
	const float4 *attrs;                       // per triangle, 8 x float4: what gBuffer.vert hands to the rasteriser (vertex_stage_kernel)
	const int *triMaterial;                    // per triangle: the material index of its draw
	const restir_material_uniforms *uniforms;  // per material
	const restir_material_textures *bindings;  // per material
	const uchar4 *texels;                      // every texture's level 0, back to back
	const uint4 *textureTable;                 // per texture: (first texel, width, height, -)
	const float *srgbThresholds;               // 256 floats: the smallest value that encodes to each 8-bit sRGB code
	const uint32_t *recordTri;                 // with the 4-wide image in use the binary image's leaves name triangle RECORDS: record -> triangle; null: identity
	int nMaterials, nTextures;
};
void launch_vertex_stage(const restir_vertex *vertices, const uint32_t *indices, const restir_draw *draws, const restir_model_matrices *matrices,
                         const uint32_t *triDraw, const uint32_t *drawFirstTri, uint32_t nTris, float4 *attrs, cudaStream_t s);
void launch_gbuffer(const SceneView &sc, const GBufferScene &g, const Band &band, const RaycastCamera &cam, float zNear, float zFar, void *albedo,
                    void *normal, void *material, void *worldPos, void *depth, cudaStream_t s);
cudaError_t preload_gbuffer_kernels();

// ---- the reference's compile-time variants: RESERVOIR_SIZE > 1, UNBIASED_MIS, one kernel per shader (restir_generic.cu) ----------
// Reservoir buffers of these kernels hold the reference's std430 records (generic_reservoir_bytes each).  false: (n, mis) is not instantiated.
bool generic_variant_supported(int n, bool mis);
size_t generic_reservoir_bytes(int n, bool mis);
bool launch_generic_restir(int n, bool mis, const PassParams &p, void *out, const void *prev, cudaStream_t s);
bool launch_generic_spatial(int n, bool mis, const PassParams &p, const void *in, void *out, int iter, cudaStream_t s);
bool launch_generic_unbiased(int n, bool mis, const PassParams &p, const void *in, void *out, int numNeighbors, cudaStream_t s);
bool launch_generic_lighting(int n, bool mis, const PassParams &p, const restir_lighting_uniforms &lu, const void *res, void *outPixels, int fmt,
                             cudaStream_t s);
cudaError_t preload_generic_kernels();

// ---- AabbTree::build on the device (restir_bvh_build.cu) ---------------------------------------------------------------
struct BvhLeaves { // aabbTreeBuilder.cpp:28-32 as arrays; `bin` is shared by both copies
	float *lo[3], *hi[3], *cen[3];
	int *geom;
	unsigned *bin;
};
struct BvhJob { // one BuildStep of the reference's queue (aabbTreeBuilder.cpp:19-27)
	long long slot; // where the id of the node (or leaf) of this range is written: node * 2 + (0 left | 1 right); -1: the root
	int beg, end;
	int node;       // id of the node this range makes; -1: a single leaf
	int split;      // ordinal among the ranges of this level that split (> 2 leaves); -1: finished
	int pivot, keepOrder;
};
struct BvhJobState { // reductions of one splitting range: (float key, leaf position) pairs, see restir_bvh_build.cu
	unsigned long long cenMin[3], cenMax[3], geoMin[3], geoMax[3];
	unsigned long long binMin[12][3], binMax[12][3];
	unsigned binCount[12];
	float outerArea, axisLo, binWidth;
	int axis, bestSplit;
};
size_t bvh_build_scratch_bytes(unsigned n);
cudaError_t build_aabb_tree_device(const float4 *tris, unsigned n, restir_aabb_node *nodes, void *scratch, int *levelsOut, unsigned *nonFinite,
                                   cudaStream_t s);
void launch_bvh_image(const restir_aabb_node *nodes, unsigned n, float4 *image, cudaStream_t s);
// the 4-wide image of a tree on the device (restir_wide_build.cu): byte for byte what build_wide_image makes on the host
size_t wide_build_scratch_bytes(unsigned nNodes, unsigned nTris);
cudaError_t build_wide_image_device(const restir_aabb_node *nodes, unsigned nNodes, unsigned nTris, const WideQuant &quant, WideNode *wide, unsigned *triOrder,
                                    float *leafBoxes, float4 *image, void *scratch, unsigned *nWide, int *depth, bool *usable, cudaStream_t s);
cudaError_t preload_wide_build_kernels();
cudaError_t preload_bvh_build_kernels();

// ---- halo exchange over peer memory (restir_halo.cu) ----------------------------------------------------------
// side 0 = the neighbour that owns the rows above this band, side 1 = the rows below.
struct HaloPush {
	const PackedReservoir *local;  // the buffer just produced (rows [localAllocBegin, ...))
	PackedReservoir *peer[2];      // the same buffer of the neighbour on each side, mapped into this process; null = no neighbour
	unsigned long long *peerFlag[2]; // the neighbour's counter for (this side as seen from there, this buffer)
	int W, localAllocBegin;
	int peerAllocBegin[2];
	int firstRow[2], rows[2];      // the rows of this band the neighbour holds as halo
	unsigned long long sequence;   // how many times this buffer has been produced, this launch included
	unsigned *ticket;              // zero-initialised scratch word (last-block detection)
};
cudaError_t launch_halo_push(const HaloPush &hp, int smCount, cudaStream_t s);
cudaError_t launch_halo_wait(const unsigned long long *flagA, const unsigned long long *flagB, unsigned long long sequence, unsigned long long *counters,
                             cudaStream_t s);

// every kernel of the library loaded up front (restir_create), see restir_kernels.cu preload_pixel_kernels
cudaError_t preload_pixel_kernels();
cudaError_t preload_trace_kernels();
cudaError_t preload_halo_kernels();

} // namespace restir
