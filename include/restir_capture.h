/* restir_capture.h — a frame capture of the hot path's inputs and outputs (fixture / wire format).
 *
 * One file holds everything the four passes read and, optionally, what the reference wrote, so that a run of the
 * real Vulkan application on another machine can be replayed through this library (or through the CPU oracle) and
 * compared: the scene buffers exactly as src/sceneBuffers.h / src/aabbTreeBuilder.h lay them out, and per frame the
 * two uniform blocks, the five G-buffer attachments in the NVIDIA-default formats (src/passes/gBufferPass.cpp:75-108)
 * and the reservoir SSBOs (restirStructs.glsl:19-34, 64 bytes per pixel).  On the reference side the dump is a
 * vkCmdCopyImageToBuffer of the G-buffer attachments and a vkCmdCopyBuffer of the reservoir buffers after
 * _recordMainCommandBuffers' passes (src/app.h:212-262), plus the host copies of the uniform structs.
 *
 * Little-endian, no padding between sections.  Pixel (x, y) of a plane is at index y * width + x (restirOmni.glsl:145).
 *
 *   restir_capture_header
 *   nodes      n_nodes     x 80 bytes   (restir_aabb_node)
 *   triangles  n_triangles x 48 bytes   (restir_triangle)
 *   point-light blob, triangle-light blob, alias-table blob   (point_blob_bytes, tri_blob_bytes, alias_blob_bytes)
 *   frames x {
 *       restir_uniforms (128 bytes), restir_lighting_uniforms (96 bytes)
 *       albedo RGBA8 (4 B/px), normal RGBA16_SNORM (8), material RG16_UNORM (4), worldPos RGBA32F (16), depth D32F (4)
 *       [RESTIR_CAPTURE_HAS_INITIAL]  reservoirs after the restir pass        (64 B/px)
 *       [RESTIR_CAPTURE_HAS_FINAL]    reservoirs after the reuse passes       (64 B/px)
 *       [RESTIR_CAPTURE_HAS_RGBA]     lighting output, linear RGBA32F         (16 B/px)
 *   }
 * Frame f uses G-buffer slot f & 1 and the buffer roles of src/app.h:298-332; the first frame's history is all zero.
 */
#ifndef RESTIR_CAPTURE_H
#define RESTIR_CAPTURE_H

#include <stdint.h>

#define RESTIR_CAPTURE_MAGIC "RSTRCAP1"
#define RESTIR_CAPTURE_HAS_INITIAL 1u
#define RESTIR_CAPTURE_HAS_FINAL 2u
#define RESTIR_CAPTURE_HAS_RGBA 4u

typedef struct restir_capture_header {
	char magic[8];               /* RESTIR_CAPTURE_MAGIC */
	uint32_t version;            /* 1 */
	uint32_t width, height, frames;
	uint32_t unbiased;           /* App::_unbiasedSpatialReuse (src/app.h:174) */
	uint32_t unbiased_neighbors; /* NUM_NEIGHBORS the unbiased shader was compiled with (unbiasedReuse.glsl:48) */
	uint32_t spatial_iterations; /* src/app.h:169 */
	uint32_t expected;           /* RESTIR_CAPTURE_HAS_* bits */
	uint32_t n_nodes, n_triangles;
	uint64_t point_blob_bytes, tri_blob_bytes, alias_blob_bytes;
} restir_capture_header;

#endif
