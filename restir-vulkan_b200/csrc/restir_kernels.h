// restir_kernels.h — launch wrappers implemented in restir_kernels.cu (internal to the library).
#pragma once

#include "restir_device.cuh"

namespace restir {

// host-precomputed camera basis for the G-buffer fixture tool (twin of oracle_raycast_gbuffer's preamble)
struct RaycastCamera {
	float pos[3], fwd[3], right[3], up[3];
	float sx, sy;
	float pv[16];
};

void launch_restir_omni(const PassParams &p, PackedReservoir *out, const PackedReservoir *prev, cudaStream_t s);
void launch_spatial_reuse(const PassParams &p, const PackedReservoir *in, PackedReservoir *out, int iter, cudaStream_t s);
void launch_unbiased_reuse(const PassParams &p, const PackedReservoir *in, PackedReservoir *out, int numNeighbors, cudaStream_t s);
void launch_lighting(const PassParams &p, const restir_lighting_uniforms &lu, const PackedReservoir *res, void *out, int fmt, cudaStream_t s);
void launch_trace_segments(const SceneView &sc, const float *p1, const float *p2, unsigned long long n, unsigned char *shadowed,
                           unsigned long long *counters, cudaStream_t s);
void launch_raycast_gbuffer(const SceneView &sc, const Band &band, const RaycastCamera &cam, const int *triMaterial, const uint4 *materialTable,
                            void *albedo, void *normal, void *material, void *worldPos, void *depth, cudaStream_t s);
void launch_unpack_reservoirs(const SceneView &sc, const PackedReservoir *in, restir_reservoir *out, size_t n, cudaStream_t s);
void launch_pack_reservoirs(const restir_reservoir *in, PackedReservoir *out, size_t n, cudaStream_t s);
void launch_derive_light_tables(const restir_point_light *pl, int np, float4 *pointOut, const restir_tri_light *tl, int nt, float4 *triOut,
                                cudaStream_t s);

} // namespace restir
