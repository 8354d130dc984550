/* restir_b200.h — C ABI of the B200-native ReSTIR resampling passes.
 *
 * Drop-in boundary for the reference's four GPU passes and the resources they bind.  Each entry point
 * names the reference interface it replaces (paths relative to lukedan/ReSTIR-Vulkan).  Plain
 * pointers and sizes only; no Vulkan, no torch.  Every call returns 0 on success or a negative
 * RESTIR_E_* code (the reference's vkCheck aborts instead, src/misc.cpp:21-26); the message is
 * available from restir_last_error().  CUDA errors are sticky.
 *
 * Threading/ordering (reference: one queue, full barriers between passes, one frame in flight —
 * restirPass.h:37-41, app.cpp:771-773): calls on one context are issued in order on one CUDA stream
 * and are asynchronous with respect to the host unless stated; restir_synchronize() waits.  Uniform
 * blocks are snapshotted at call time.  A context is not thread-safe; contexts are independent.
 *
 * Ownership (reference: App owns every buffer, passes hold handles — src/app.h:113-148): the context
 * owns the device memory it allocates (reservoirs, scene copies, G-buffer copies made by
 * restir_upload_gbuffer); the caller owns every pointer it passes in.
 */
#ifndef RESTIR_B200_H_
#define RESTIR_B200_H_

#include <stddef.h>
#include <stdint.h>

#include "restir_layouts.h"

#ifdef __cplusplus
extern "C" {
#endif

#define RESTIR_OK 0
#define RESTIR_E_INVALID (-1)  /* bad argument / wrong call order */
#define RESTIR_E_CUDA (-2)     /* CUDA runtime error (sticky) */
#define RESTIR_E_NOMEM (-3)
#define RESTIR_E_UNSUPPORTED (-4)
#define RESTIR_E_HALO (-5)     /* connected bands: a wait for a neighbour's halo rows timed out — every frame since is suspect (sticky until
                                * the counters are reset) */

typedef struct restir_context restir_context;

/* Reservoir buffer ids — the three buffers App keeps (src/app.h:264-284): two per-frame buffers that
 * ping-pong with the G-buffers, and the temporary the unbiased path writes its initial samples to. */
#define RESTIR_BUF_FRAME0 0
#define RESTIR_BUF_FRAME1 1
#define RESTIR_BUF_TEMP 2

/* Output formats of the lighting pass (reference renders into the swapchain image, lightingPass.h). */
#define RESTIR_OUT_RGBA32F 0     /* linear, before any quantisation: the parity format */
#define RESTIR_OUT_RGBA8_SRGB 1  /* 8-bit sRGB-encoded, what a *_SRGB swapchain stores */

/* ---- lifetime -------------------------------------------------------------------------------- */

/* Replaces: pass construction in App::App (src/app.cpp:437-536, Pass::create<> src/passes/pass.h:112-119).
 * `device` is a CUDA device ordinal; `stream` is a cudaStream_t to issue on, or NULL for a stream the
 * context creates. */
int restir_create(restir_context **out, int device, void *stream);
void restir_destroy(restir_context *ctx);
const char *restir_last_error(const restir_context *ctx);
/* Replaces: waiting on _mainFence (src/app.cpp:771-773).  Returns RESTIR_E_HALO when a connected band gave up waiting for a
 * neighbour's halo rows since the counters were last reset (restir_get_counters): the passes after that wait read stale rows. */
int restir_synchronize(restir_context *ctx);

/* ---- scene resources (set 0 / set 2 descriptors) ------------------------------------------------ */

/* Replaces: AabbTreeBuffers::create (src/aabbTreeBuilder.h:25-51) + initializeSoftwareRayTracingDescriptorSet
 * (src/passes/restirPass.h:221-237).  Bytes exactly as AabbTree::build produces them: n_nodes x 80, n_tris x 48. */
int restir_upload_bvh(restir_context *ctx, const void *nodes, uint32_t n_nodes, const void *triangles, uint32_t n_triangles);

/* Replaces: AabbTree::build (src/aabbTreeBuilder.cpp:52-214) + AabbTreeBuffers::create in one call, ON THE DEVICE
 * (rebuild for dynamic geometry): uploads the n x 48-byte world-space triangles (HOST pointer, the
 * reference's order), builds the tree with the library's kernels — the reference's breadth-first build one queue generation
 * per round of launches, byte-identical to restir_build_aabb_tree and to the reference's own output — and installs it as the
 * context's tree like restir_upload_bvh.  nodes_out: NULL, or HOST memory for the (n - 1) x 80-byte nodes.
 * restir_get_bvh_info then reports depth and upper bounds of the stack occupancies.  RESTIR_E_UNSUPPORTED for non-finite
 * coordinates (build those on the host). */
int restir_build_bvh_device(restir_context *ctx, const void *triangles, uint32_t n_triangles, void *nodes_out);

/* Replaces: the three light SSBOs of SceneBuffers (src/sceneBuffers.h:100-124, 241-270) +
 * initializeStaticDescriptorSetFor (restirPass.h:120-150).  Blobs = {int32 count; pad to 16; array}. */
int restir_upload_lights(restir_context *ctx, const void *point_blob, size_t point_bytes, const void *tri_blob,
                         size_t tri_bytes, const void *alias_blob, size_t alias_bytes);

/* ---- per-resolution resources ------------------------------------------------------------------- */

/* Replaces: App::_updateRestirBuffers (src/app.h:264-284): allocates and ZERO-FILLS the three
 * reservoir buffers for a width x height screen and drops any bound G-buffers. */
int restir_resize(restir_context *ctx, uint32_t width, uint32_t height);

/* Row-band variant for multi-GPU (no reference equivalent — the reference is single-GPU): this
 * context shades rows [row_begin, row_end) of a width x height screen and keeps `halo` extra rows
 * on each side (clipped to the screen) for neighbour / reprojection reads.  All per-pixel pointers
 * passed to or returned by this context then cover rows [alloc_begin, alloc_end) =
 * [max(0,row_begin-halo), min(height,row_end+halo)), see restir_get_band(). */
int restir_resize_band(restir_context *ctx, uint32_t width, uint32_t height, uint32_t row_begin, uint32_t row_end,
                       uint32_t halo);
int restir_get_band(const restir_context *ctx, uint32_t *row_begin, uint32_t *row_end, uint32_t *alloc_begin,
                    uint32_t *alloc_end);

/* Replaces: the G-buffer image bindings of initializeFrameDescriptorSetFor (restirPass.h:152-219),
 * SpatialReusePass::initializeDescriptorSetFor (spatialReusePass.h:28-89), UnbiasedReusePass
 * (unbiasedReusePass.h:111-161) and LightingPass (lightingPass.h:48-96).  slot is the G-buffer index
 * (0/1, src/app.h numGBuffers).  Planes are DEVICE pointers that must stay valid while bound.
 * Formats and pitches: exactly one tuple is accepted, RESTIR_GBUFFER_NVIDIA_DEFAULT with tightly packed rows (pitch = width x texel
 * size) — what GBuffer::Formats::initialize picks on NVIDIA hardware.  The other candidates of gBufferPass.cpp:75-108 (RGB16_SNORM /
 * RGB16F / RGBA16F / RGB32F normals, D32S8 / D24S8 depth, RGB32F positions) and padded pitches are refused on purpose
 * (RESTIR_E_UNSUPPORTED): every kernel decodes the texels inline, a second format is a second set of kernels, and an importer that
 * holds such images converts or copies them once per frame at the interop boundary (linear, tightly packed images are what a
 * CUDA external-memory import of a Vulkan image needs anyway). */
int restir_bind_gbuffer(restir_context *ctx, int slot, restir_gbuffer_format format, const restir_gbuffer_planes *device_planes);
/* Same, from HOST memory (pinned or pageable): copies the planes into context-owned device memory and binds
 * them.  The copy runs on a copy stream the context owns (asynchronous when the host memory is pinned, which
 * must then stay untouched until restir_synchronize or the first pass that reads the slot has been waited
 * for): it starts as soon as the last pass that read this slot has finished — for frame f+1 that is frame
 * f's restir pass (restirOmni.glsl:163-209 is the only reader of the previous G-buffer), so the upload is
 * hidden under frame f's reuse and lighting passes — and every later pass that reads the slot waits for it.
 * The observable order is the in-order one. */
int restir_upload_gbuffer(restir_context *ctx, int slot, restir_gbuffer_format format, const restir_gbuffer_planes *host_planes);

/* ---- the G-buffer pass ------------------------------------------------------ */

typedef struct restir_camera { /* src/camera.h:7-13 */
	float position[3], lookAt[3], worldUp[3];
	float zNear, zFar, fovYRadians, aspectRatio;
} restir_camera;

/* Replaces: the vertex / index / matrix buffers of SceneBuffers (src/sceneBuffers.h:173-203, 225-230) and the draw
 * list of GBufferPass::issueCommands (src/passes/gBufferPass.cpp:135-152), plus the vertex stage gBuffer.vert:22-34,
 * which runs here once per upload: every triangle gets its three world-space normals (transformInverseTransposed,
 * normalised), tangents (transform, normalised, w kept) and texture coordinates.  The triangles of the draws, in draw
 * order, must be the triangles of restir_upload_bvh (AabbTree::build collects them in exactly this order,
 * src/aabbTreeBuilder.cpp:59-76): call restir_upload_bvh first.  Host pointers. */
int restir_upload_geometry(restir_context *ctx, const restir_vertex *vertices, uint64_t n_vertices, const uint32_t *indices, uint64_t n_indices,
                           const restir_draw *draws, const restir_model_matrices *matrices, uint32_t n_draws);
/* Replaces: the material uniform buffer (sceneBuffers.h:205-222), the texture images (sceneBuffers.h:126-171) and the
 * per-material descriptor sets (gBufferPass.cpp:193-247).  Textures are sampled like the reference's samplers as far as
 * this producer goes: repeat wrap, bilinear, level 0 of what is uploaded (no mip chain, no anisotropy).  Host pointers. */
int restir_upload_materials(restir_context *ctx, const restir_material_uniforms *uniforms, const restir_material_textures *bindings,
                            uint32_t n_materials, const restir_texture *textures, uint32_t n_textures);
/* Replaces: GBufferPass::issueCommands (gBufferPass.cpp:116-157) running gBuffer.vert / gBuffer.frag:27-80 into G-buffer
 * slot `slot`, for the camera whose projectionViewMatrix the reference writes into the pass's uniform buffer
 * (src/app.cpp:802-806, src/camera.h:25-50).  B200 has no rasteriser: primary visibility is a closest-hit ray cast
 * through each pixel centre against the uploaded tree (back faces culled, ALPHA_MODE_MASK fragments below the cutoff
 * discarded, near / far planes as clip planes, depth test LESS = nearest hit, ties to the earlier draw).  The planes are
 * context-owned, in the formats of RESTIR_GBUFFER_NVIDIA_DEFAULT with the pass's clears, and are bound to the slot
 * like restir_upload_gbuffer's — a full frame then needs no G-buffer traffic over PCIe at all. */
int restir_pass_gbuffer(restir_context *ctx, int slot, const restir_camera *camera);
/* The device planes of a slot the context owns (restir_pass_gbuffer / restir_upload_gbuffer), rows [alloc_begin, alloc_end). */
int restir_gbuffer_device_planes(restir_context *ctx, int slot, restir_gbuffer_planes *out);

/* ---- Vulkan interop -------------------------------------------------------- */

/* The CUDA half of sharing the reference's G-buffer images with this library instead of copying them: `fd` is the POSIX file
 * descriptor vkGetMemoryFdKHR returns for a VkDeviceMemory allocated with VkExportMemoryAllocateInfo{ handleTypes =
 * VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT } (src/vma.h allocations of gBufferPass.cpp:34-58 would need that pNext and LINEAR
 * tiling); `size` the allocation size.  Returns a device pointer to the whole allocation, valid until restir_release_external_memory
 * or restir_destroy; pass pointer + image offset to restir_bind_gbuffer.  CUDA takes ownership of the descriptor.  Ordering
 * between the Vulkan queue and this context's stream stays the caller's (vkQueueWaitIdle / restir_synchronize, or exported
 * semaphores).  Not exercised by the tests of this repository beyond its error path: the image has no Vulkan. */
int restir_import_external_memory(restir_context *ctx, int fd, uint64_t size, void **device_ptr);
int restir_release_external_memory(restir_context *ctx, void *device_ptr);

/* ---- uniforms ------------------------------------------------------------------------------------ */

/* Replaces: the mapped write of the RestirUniforms UBO each frame (src/app.cpp:775-826, initial values
 * :414-434).  The 128-byte block is taken verbatim.  In band mode screenSize must be the FULL screen. */
int restir_set_uniforms(restir_context *ctx, const restir_uniforms *uniforms);
/* Replaces: the LightingPassUniforms UBO write (src/app.cpp:793-800). */
int restir_set_lighting_uniforms(restir_context *ctx, const restir_lighting_uniforms *uniforms);
/* Replaces: the compile-time switches of include/structs/restirStructs.glsl:16-17 — `#define RESERVOIR_SIZE 1` and the
 * commented-out `#define UNBIASED_MIS` — and the choice between split and single-kernel passes the authors measured
 * (media/milestone3, slides 6-7).
 *   reservoir_size  RESERVOIR_SIZE: 1 (as shipped), 2 or 4 samples per reservoir, each streamed with its own random draw
 *                   (reservoir.glsl:28-42) and traced with its own shadow rays (restirOmni.glsl:149-160, unbiasedReuse.glsl:132-166);
 *   unbiased_mis    != 0: UNBIASED_MIS — LightSample carries sumPHat and the unbiased pass normalises by the sum of target pdfs
 *                   instead of sample counts (unbiasedReuse.glsl:74-82, 134-181), including the reference's own argument order at
 *                   reservoir.glsl:54-58;
 *   fused_passes    != 0: every pass is ONE kernel that traces its shadow rays inline, like the reference's shaders, also for the
 *                   shipped configuration (1, off) — by default that one runs on the tuned path, whose passes are cut at their rays.
 * Anything but (1, 0, 0) runs on the generic kernels (csrc/restir_generic.cu), whose three reservoir buffers hold the reference's
 * own std430 records: LightSample 48 bytes (64 with sumPHat), Reservoir = reservoir_size samples + numStreamSamples padded to 16,
 * i.e. reservoir_size * (48 | 64) + 16 bytes — what restir_download_reservoirs / restir_upload_reservoirs then move and
 * restir_get_reservoir_bytes reports.  Changing the record reallocates and zero-fills the buffers like restir_resize.  Not
 * available on connected bands (RESTIR_E_UNSUPPORTED).  Results: bit-identical to the reference's shader sources compiled with
 * the same defines (tests/test_variants.py). */
int restir_set_reservoir_variant(restir_context *ctx, uint32_t reservoir_size, int unbiased_mis, int fused_passes);
int restir_get_reservoir_bytes(const restir_context *ctx, size_t *bytes);
/* The unbiased pass's neighbour count is a compile-time 3 in the reference (unbiasedReuse.glsl:48) and
 * ignores uniforms.spatialNeighbors; this overrides it (1..16) for the north-star's 5-neighbour runs. */
int restir_set_unbiased_neighbors(restir_context *ctx, uint32_t count);

/* Shadow-ray traversal.  The uploaded tree is always the reference's (aabbTreeBuilder node / triangle layout) and the
 * visibility bits are always those of softwareRaytracing.glsl:39-85 on that tree; what differs is what the trace kernel reads.
 * AUTO (default): restir_upload_bvh derives, on the host,
 *   - the binary image: the same nodes at 64 bytes each (two 32-byte loads from one line instead of five 16-byte loads that
 *     straddle lines), walked without per-push bound checks once the reference's 32-entry stack is known not to overflow; and
 *   - the WIDE image (csrc/wide_image.h): the tree collapsed to 4 children per node with boxes quantised outwards to a 15-bit
 *     grid, 64 bytes per node.  For rays whose origin, direction and reciprocal direction are finite the reference's slab test
 *     is monotone in the box, its boxes are nested (checked), so the reference tests a triangle iff the triangle's own leaf box
 *     passes; the wide walk is a conservative search for those leaves and evaluates the leaf box and the triangle with the
 *     reference's arithmetic.  Used when the uploaded boxes are finite and nested and every triangle hangs under one leaf;
 *     rays outside the finite range walk the binary image.
 * IMAGE: the binary image only.  WIDE: same as AUTO.  REFERENCE_ORDER: walk the 80-byte nodes literally, dropped pushes
 * counted — also what AUTO falls back to when the stack could overflow.  restir_get_bvh_info tells which one is in use.
 * Takes effect at the next restir_upload_bvh (restir_pass_gbuffer, whose closest-hit walk reads the binary image and the same
 * triangle records, honours REFERENCE_ORDER at once: same planes either way).  Uploads whose child indices are out of range or that are not trees are rejected
 * with RESTIR_E_INVALID.  restir_build_bvh_device derives both images on the device (csrc/restir_wide_build.cu). */
#define RESTIR_TRAVERSAL_AUTO 0
#define RESTIR_TRAVERSAL_REFERENCE_ORDER 1
#define RESTIR_TRAVERSAL_IMAGE 2
#define RESTIR_TRAVERSAL_WIDE 3
int restir_set_traversal(restir_context *ctx, int mode);

/* No reference equivalent.  With the WIDE traversal the trace kernel keeps, per 64 x 32-pixel screen region and light, the last
 * triangle that occluded a ray, and tests it (the reference's triangle test and the reference's slab test on its leaf box)
 * before queueing a ray for a walk: "shadowed" needs one witness, and most shadowed rays of a region at a light share theirs.
 * The table only proposes witnesses — its contents cannot change a visibility bit — but shadow_rays_traced then depends on
 * what earlier frames left in it.  enable = 0 walks every ray that is not elided (A/B, deterministic counters); every call
 * clears the table.  Default: 1 = entries chosen by light index for up to 4 096 lights and by the segment's direction (cube face
 * + 8 x 8 grid) beyond, where a (region, light) pair practically never comes back; 2 / 3 force the first / the second (A/B). */
int restir_set_occluder_cache(restir_context *ctx, int enable);

typedef struct restir_bvh_info {
	uint32_t nodes, triangles;
	uint32_t reachable_nodes, depth;
	uint32_t reference_stack_bound; /* worst-case occupancy of the reference's 32-entry stack on this tree */
	int32_t traversal;              /* what the trace kernel walks: RESTIR_TRAVERSAL_WIDE, _IMAGE or _REFERENCE_ORDER */
	uint32_t wide_nodes, wide_depth, wide_stack_bound; /* the 4-wide image (0 when not in use) */
} restir_bvh_info;
int restir_get_bvh_info(const restir_context *ctx, restir_bvh_info *out);
/* Inspection (tests, tools): the 4-wide image the context walks — n_wide nodes of 64 bytes (csrc/wide_image.h WideNode: 12 words
 * of quantised planes, first child << 4, first triangle record - inner count, mask of the inner slots, slot count) and, per triangle record, the index of the
 * uploaded triangle it holds.  Either pointer may be NULL.  RESTIR_E_UNSUPPORTED when the tree is not walked wide.  The image of
 * a tree built by restir_build_bvh_device equals, byte for byte, the image restir_upload_bvh derives from the same nodes. */
int restir_get_wide_image(restir_context *ctx, void *wide_nodes, uint32_t capacity, uint32_t *n_wide, uint32_t *tri_order);
/* Host-only self-check of the WIDE traversal (no reference equivalent, no GPU, called by no pass): builds the wide image of
 * the tree exactly as restir_upload_bvh does and walks it on the CPU with the operations the kernel uses (the box arithmetic is
 * one header shared by host and device, csrc/wide_image.h) for n segments p1 -> p2 (3 floats each).  shadowed[i] = 1 when the
 * segment is occluded; walked_wide[i] = 0 marks the segments the wide walk hands to the binary image (non-finite or
 * out-of-range origin / direction: shadowed[i] is then 0); *visits (may be NULL) = wide nodes visited.  RESTIR_E_UNSUPPORTED
 * when the tree is not walked wide (message says why). */
int restir_check_wide_walk(const void *nodes, uint32_t n_nodes, const void *triangles, uint32_t n_triangles, const float *p1, const float *p2,
                           uint64_t n, unsigned char *shadowed, unsigned char *walked_wide, uint64_t *visits, char *message, size_t message_bytes);
/* No reference equivalent (the reference traces every ray it asks for).  The unbiased pass answers a neighbour ray
 * of unbiasedReuse.glsl:139-156 without walking the tree when the answer is already determined, exactly: the pixel's
 * own ray (:157-166) is shadowed, or the segment is bit-identical to the neighbour's own ray.  enable = 0 walks every ray
 * instead (A/B measurements and the tests that require both settings to give identical bits).  Default: 1.
 * enable = 2 (experiment, WIDE traversal only) adds one walk per DISTINCT segment: pixels of a region share samples and
 * neighbours, so many neighbour rays ask for the same (neighbour position, sample position) pair, bit for bit — the first item
 * to enter it into a hash table walks it, the others copy the answer afterwards.  Exact, and measured to lose: on Sponza 1080p
 * with 200 lights it saves 22 % of the neighbour walks and still costs 0.11 ms per frame more than it saves (one atomic into a
 * 64 MB table per ray); with 1 M lights 4 % of the rays are duplicates (profiles/r2_j_summary.md). */
int restir_set_ray_elision(restir_context *ctx, int enable);
/* No reference equivalent.  The biased spatial pass tests every neighbour's depth and normal before it merges the neighbour's
 * reservoir (spatialReuse.comp:62-70).  enable != 0: the CTA stages the depth and normal texels of its 32x8 tile and the
 * 31-pixel apron in shared memory first (79 KB) and the gates read those; 0 (default): every gate reads its two texels
 * through L1/L2 directly.  Same results bit for bit; which one is faster is a measurement (profiles/r2_e_summary.md). */
int restir_set_spatial_staging(restir_context *ctx, int enable);
/* The same checks restir_upload_bvh runs, on the host and without a context (no GPU needed): RESTIR_E_INVALID
 * and a message for an upload that would be rejected; otherwise RESTIR_OK, `out` filled (traversal says which
 * walk such an upload would get) and, for the reference-order fallback, the reason in `message`.  The
 * reference itself checks nothing (its only assert is dummyRoot == 0, src/aabbTreeBuilder.cpp:212). */
int restir_check_aabb_tree(const void *nodes, uint32_t n_nodes, uint32_t n_triangles, restir_bvh_info *out, char *message,
                           size_t message_bytes);

/* ---- the passes ---------------------------------------------------------------------------------- */

/* Replaces: RestirPass::issueCommands, software-ray-tracing pipeline (restirPass.h:36-62) running
 * restirOmniSoftware.comp -> restirOmni.glsl:86-212.  Reads G-buffer slots `gbuffer` (current) and
 * gbuffer^1 (previous; zero texels if unbound), reservoir buffer `prev_buffer`; writes `out_buffer`. */
int restir_pass_restir(restir_context *ctx, int gbuffer, int out_buffer, int prev_buffer);
/* Replaces: SpatialReusePass::issueCommands (spatialReusePass.h:11-26) running spatialReuse.comp:30-86
 * with push constant `iter`. */
int restir_pass_spatial(restir_context *ctx, int gbuffer, int in_buffer, int out_buffer, int iter);
/* Replaces: UnbiasedReusePass::issueCommands, software pipeline (unbiasedReusePass.h:23-49) running
 * unbiasedReuseSoftware.comp -> unbiasedReuse.glsl:50-185. */
int restir_pass_unbiased(restir_context *ctx, int gbuffer, int in_buffer, int out_buffer);
/* Replaces: LightingPass::issueCommands (lightingPass.h:14-42) running lighting.frag:43-71,103
 * (debugMode 0 only).  `out` is a DEVICE pointer to rows [alloc_begin, alloc_end) x width pixels of
 * `out_format`; only rows [row_begin,row_end) are written. */
int restir_pass_lighting(restir_context *ctx, int gbuffer, int buffer, void *out_device, int out_format);

/* Replaces: the pre-recorded command buffer of App::_recordMainCommandBuffers (src/app.h:212-262)
 * minus the G-buffer pass, with App's buffer roles (src/app.h:298-332): frame index i in {0,1};
 *   unbiased != 0: restir(out=TEMP, prev=FRAME[i^1]) ; unbiased(in=TEMP, out=FRAME[i])
 *   unbiased == 0: restir(out=FRAME[i], prev=FRAME[i^1]) ; for j < spatial_iterations:
 *                  spatial(in=FRAME[i], out=FRAME[i^1], iter=2j) ; spatial(in=FRAME[i^1], out=FRAME[i], iter=2j+1)
 * On a band context every band edge inside the screen must have its neighbour connected (restir_band_connect): the
 * context then exchanges the halo rows itself between the passes; otherwise RESTIR_E_INVALID (drive the passes one by
 * one and copy the halo rows through restir_reservoir_device_ptr instead). */
int restir_frame(restir_context *ctx, int i, int unbiased, int spatial_iterations);
/* Replaces: the same command buffer INCLUDING its lighting pass (src/app.h:212-262: the ReSTIR passes, then
 * LightingPass::issueCommands on the same G-buffer and FRAME[i]) = restir_frame followed by
 * restir_pass_lighting(ctx, i, FRAME[i], out_device, out_format), same results bit for bit.  The lighting of a pixel runs
 * at the end of the kernel that produces its final reservoir (the unbiased pass's normalisation, or the last spatial
 * pass), from registers: it does not read the reservoir and the G-buffer texels back.  The reference authors measured the
 * same fusion (media/milestone3 slides 6-7).  Lighting uniforms must be set. */
int restir_frame_lit(restir_context *ctx, int i, int unbiased, int spatial_iterations, void *out_device, int out_format);

/* ---- reservoir access (parity / replay / halo exchange) -------------------------------------- */

/* Copy reservoir buffer `buffer` to / from HOST memory in the reference's 64-byte layout
 * (restir_reservoir), rows [alloc_begin, alloc_end).  Synchronous.  No reference equivalent (the
 * reference cannot dump); SURVEY.md §5 "checkpoint / resume". */
int restir_download_reservoirs(restir_context *ctx, int buffer, restir_reservoir *dst_host);
int restir_upload_reservoirs(restir_context *ctx, int buffer, const restir_reservoir *src_host);
/* Device pointer and row pitch (bytes) of the context's internal packed reservoir buffer, for
 * GPU-to-GPU halo exchange (row r of the screen lives at ptr + (r - alloc_begin) * pitch). */
int restir_reservoir_device_ptr(restir_context *ctx, int buffer, void **ptr, size_t *row_pitch_bytes);

/* ---- stand-alone visibility (the minimum parity slice) ---------------------------------------- */

/* testVisibility(p1,p2) of visibilityTest.glsl:1-4,27-28 over n segments (raytrace of
 * softwareRaytracing.glsl:39-85).  p1/p2: DEVICE float[n][3]; shadowed: DEVICE uint8[n] (1 = shadowed). */
int restir_trace_segments(restir_context *ctx, const float *p1_device, const float *p2_device, uint64_t n,
                          uint8_t *shadowed_device);

/* ---- row-band neighbours over peer memory (no reference equivalent: the reference is single-GPU) ------------
 * Without these calls a band context leaves the halo rows of its reservoir buffers to the caller (copy them between
 * the passes through restir_reservoir_device_ptr).  With its neighbours connected, the context does it itself with
 * its own kernels over NVLink peer memory: when a pass has produced a buffer it stores the boundary rows straight into
 * each neighbour's copy of that buffer and raises a counter there; the pass that is about to read a halo first waits,
 * on the device, for both neighbours' counters (every rank must issue the same pass sequence).  No host
 * synchronisation and no collective on the data path.
 *   side 0 = the neighbour that owns the rows above this band, side 1 = the rows below.
 *   Same process (contexts on one or several devices with peer access): restir_band_local_peer of the neighbour,
 *   then restir_band_connect.  One process per GPU: restir_band_export_ipc, ship the struct to the neighbour (any
 *   transport), restir_band_open_ipc there, then restir_band_connect.  Connections end at restir_resize / restir_resize_band / restir_destroy;
 *   all ranks must (re)connect before the next pass and synchronise once after connecting. */
typedef struct restir_band_peer {
	void *reservoirs[3];            /* the neighbour's three packed reservoir buffers, addressable from this process */
	void *flags;                    /* the neighbour's counter block */
	uint32_t alloc_begin, alloc_end; /* rows its buffers cover */
	uint32_t row_begin, row_end;     /* rows it owns (shades and pushes): must cover this band's halo rows on that side */
} restir_band_peer;
typedef struct restir_band_ipc {
	unsigned char reservoirs[3][64]; /* cudaIpcMemHandle_t */
	unsigned char flags[64];
	uint32_t alloc_begin, alloc_end;
	uint32_t row_begin, row_end;
} restir_band_ipc;
int restir_band_local_peer(restir_context *ctx, restir_band_peer *out);
int restir_band_export_ipc(restir_context *ctx, restir_band_ipc *out);
int restir_band_open_ipc(restir_context *ctx, const restir_band_ipc *in, restir_band_peer *out);
/* peer == NULL: no neighbour on that side.  RESTIR_E_INVALID when the neighbour's own rows do not cover this band's halo
 * rows on that side (a halo taller than the neighbouring band would need a second hop, which is not made: those rows
 * would silently stay empty) or when it holds none of this band's rows. */
int restir_band_connect(restir_context *ctx, int side, const restir_band_peer *peer);
/* Host helper: band boundaries of equal measured COST instead of equal height.  bounds_in / bounds_out: n_bands + 1
 * ascending rows from 0 to height; seconds[r]: what band r took (its own kernels, restir_profile_end); every new band
 * is at least min_rows high (>= the halo).  The frame time of a band split is the slowest band's. */
int restir_band_balanced_bounds(uint32_t height, uint32_t n_bands, const uint32_t *bounds_in, const double *seconds, uint32_t min_rows,
                                uint32_t *bounds_out);

/* ---- counters --------------------------------------------------------------------------------- */

typedef struct restir_counters {
	uint64_t shadow_rays;       /* the reference's testVisibility calls answered since the last reset */
	uint64_t stack_overflows;   /* pushes dropped on a full 32-entry traversal stack (UB in the reference) */
	uint64_t halo_misses;       /* band mode: neighbour / reprojection reads outside [alloc_begin, alloc_end) */
	uint64_t kernel_launches;   /* kernels launched by this context since the last reset */
	uint64_t halo_wait_timeouts; /* connected bands: waits for a neighbour's rows that gave up after 5 s (a lost neighbour must not hang the GPU) */
	uint64_t shadow_rays_traced; /* of shadow_rays, the ones that needed a walk of the tree: the rest were answered exactly
	                              * without one (neighbour rays of a pixel whose own ray is shadowed, unbiasedReuse.glsl:157-166;
	                              * neighbour rays bit-identical to the neighbour's own ray) */
	uint64_t shadow_rays_cached; /* of shadow_rays, the ones answered "shadowed" by the occluder cache: one exact triangle + leaf-box test
	                              * of a triangle that occluded an earlier ray of the same screen region at the same light, no walk */
} restir_counters;
/* Synchronises the stream.  `out` is always filled; the return value is RESTIR_E_HALO when halo_wait_timeouts != 0. */
int restir_get_counters(restir_context *ctx, restir_counters *out, int reset);

/* Per-kernel device times (no reference equivalent: the reference has no GPU timestamps, only an FPS
 * counter, src/fpsCounter.h).  Between _begin and _end every kernel the context launches is bracketed by
 * CUDA events on the context's stream; _end synchronises and sums them by kernel name. */
typedef struct restir_kernel_time {
	char name[48];
	uint32_t launches;
	float total_ms;
} restir_kernel_time;
int restir_profile_begin(restir_context *ctx);
int restir_profile_end(restir_context *ctx, restir_kernel_time *out, uint32_t capacity, uint32_t *count);

/* ---- scene-side builders (host, once per scene) ------------------------------------------------ */

/* Replaces: AabbTree::build (src/aabbTreeBuilder.cpp:52-214) for world-space triangles already in the
 * reference's order.  triangles: n x 48 bytes (restir_triangle).  nodes_out: (n-1) x 80 bytes. n >= 2. */
int restir_build_aabb_tree(const void *triangles, uint32_t n_triangles, void *nodes_out);
/* The same tree, byte for byte, built one breadth-first level at a time on n_threads host threads (0 = all): the
 * reference's queue hands out node ids in breadth-first order and a build step touches only its own triangle range,
 * its own node and one child slot of its parent, so the steps of a level are independent (Sponza, 262 267 triangles:
 * about 150 ms -> 37 ms on 8 cores).  The formulation a device-side rebuild follows. */
int restir_build_aabb_tree_mt(const void *triangles, uint32_t n_triangles, void *nodes_out, uint32_t n_threads);
/* Replaces: collectTriangleLightsFromScene (src/misc.cpp:380-414).  tri_material[i] indexes
 * material_emissive (n_materials x float[3]); a triangle is a light if |emissive|^2 > 1e-6.
 * Returns the number of lights written (<= n_triangles) or a negative error. */
int64_t restir_collect_triangle_lights(const void *triangles, const int32_t *tri_material, uint32_t n_triangles,
                                        const float *material_emissive, uint32_t n_materials, restir_tri_light *out);
/* Replaces: generateRandomPointLights (src/misc.cpp:358-378) with the default colour ranges [0,1):
 * libstdc++'s std::default_random_engine, default-seeded; draw order z,y,x,b,g,r (g++ evaluates the
 * reference's constructor arguments right to left). */
int restir_generate_random_point_lights(uint64_t count, const float min_xyz[3], const float max_xyz[3], restir_point_light *out);
/* Replaces: createAliasTable (src/misc.cpp:418-497).  Exactly one of the two arrays is used: point lights
 * if n_point > 0, else triangle lights. */
int restir_create_alias_table(const restir_point_light *point, uint64_t n_point, const restir_tri_light *tri,
                              uint64_t n_tri, restir_alias_column *out);

/* ---- fixture tool (synthesises INPUTS; not part of the reference's hot path) ------------------ */

/* src/camera.h:25-50: column-major projectionViewMatrix. */
int restir_camera_matrix(const restir_camera *camera, float out_pv[16]);
/* Primary-visibility ray cast of the uploaded BVH into the five G-buffer planes (DEVICE pointers, rows
 * [alloc_begin, alloc_end) of the context's screen), semantics in SURVEY.md §8d / Appendix E.
 * tri_material: DEVICE int32[n_triangles]; material_table: DEVICE uint32[n_materials][4] =
 * {albedo RGBA8, material RG16, flags(bit0: discarded), 0}. */
int restir_tools_raycast_gbuffer(restir_context *ctx, const restir_camera *camera, const int32_t *tri_material_device,
                                 const uint32_t *material_table_device, void *albedo, void *normal, void *material,
                                 void *worldPos, void *depth);

#ifdef __cplusplus
} /* extern "C" */
#endif

#endif /* RESTIR_B200_H_ */
