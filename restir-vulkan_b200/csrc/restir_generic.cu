// restir_generic.cu — the reference's compile-time variants, hand-written for sm_100a:
//
//   RESERVOIR_SIZE > 1 and UNBIASED_MIS   include/structs/restirStructs.glsl:16-17, 26; include/reservoir.glsl under both switches
//   single-kernel ("fused") passes        one kernel per reference shader, the shadow rays traced inline where the shader asks
//                                         for them — the formulation the authors measured against split passes (media/milestone3)
//
//   generic_restir_kernel    <- src/shaders/restirOmni.glsl:86-212   (candidates, visibility reuse, temporal reuse in one kernel)
//   generic_spatial_kernel   <- src/shaders/spatialReuse.comp:30-86
//   generic_unbiased_kernel  <- src/shaders/unbiasedReuse.glsl:50-185 (merge, neighbour rays, own ray, normalisation in one kernel)
//   generic_lighting_kernel  <- src/shaders/lighting.frag:43-71,103
//
// Templates on (RESERVOIR_SIZE, UNBIASED_MIS).  Reservoirs stay in the reference's own std430 layout in HBM here — LightSample 48
// bytes (64 with sumPHat), Reservoir = RESERVOIR_SIZE samples + numStreamSamples padded to 16 — so restir_download / _upload are
// plain copies.  The tuned path (restir_kernels.cu + restir_trace.cu: cut shaders, packed 32-byte reservoirs, sorted lockstep rays)
// serves the configuration the reference ships, (1, off); this file serves the others and, on request, (1, off) as well
// (restir_set_reservoir_variant(ctx, 1, 0, fused = 1)): the A/B of cutting the shaders at their rays (profiles/r2_g_summary.md).
// Same arithmetic policy (restir_math.cuh); the oracle twins are oracle_*_variant, themselves pinned against the reference's shader
// sources compiled with the same defines.
#include "restir_kernels.h"
#include "restir_pixel.cuh"
#include "restir_trace.cuh"

namespace restir {

namespace {

template <bool MIS> struct GSample {
	float px, py, pz, lum;
	float nx, ny, nz, nw;
	int lightIndex;
	float pHat, sumWeights, w;
};
template <> struct GSample<true> {
	float px, py, pz, lum;
	float nx, ny, nz, nw;
	int lightIndex;
	float pHat, sumWeights, w;
	float sumPHat;
	float pad_[3];
};
template <int N, bool MIS> struct alignas(16) GReservoir { // moved as float4: local copies must be 16-byte aligned too
	GSample<MIS> s[N];
	uint32_t M;
	uint32_t pad_[3];
};
static_assert(sizeof(GReservoir<1, false>) == 64 && sizeof(GReservoir<2, false>) == 112 && sizeof(GReservoir<1, true>) == 80 && sizeof(GReservoir<2, true>) == 144,
              "std430 sizes of Reservoir under RESERVOIR_SIZE / UNBIASED_MIS");

__device__ __forceinline__ float &sum_phat(GSample<true> &s) { return s.sumPHat; }
__device__ __forceinline__ float sum_phat(const GSample<true> &s) { return s.sumPHat; }
__device__ __forceinline__ float sum_phat(const GSample<false> &) { return 0.0f; }
__device__ __forceinline__ void set_sum_phat(GSample<true> &s, float v) { s.sumPHat = v; }
__device__ __forceinline__ void set_sum_phat(GSample<false> &, float) {}

template <int N, bool MIS> __device__ __forceinline__ GReservoir<N, MIS> load_g(const void *buf, size_t i) {
	GReservoir<N, MIS> r;
	const float4 *src = reinterpret_cast<const float4 *>(static_cast<const unsigned char *>(buf) + i * sizeof(r));
	float4 *dst = reinterpret_cast<float4 *>(&r);
#pragma unroll
	for (unsigned k = 0; k < sizeof(r) / 16; ++k) {
		dst[k] = src[k];
	}
	return r;
}
template <int N, bool MIS> __device__ __forceinline__ void store_g(void *buf, size_t i, const GReservoir<N, MIS> &r) {
	float4 *dst = reinterpret_cast<float4 *>(static_cast<unsigned char *>(buf) + i * sizeof(r));
	const float4 *src = reinterpret_cast<const float4 *>(&r);
#pragma unroll
	for (unsigned k = 0; k < sizeof(r) / 16; ++k) {
		dst[k] = src[k];
	}
}
template <int N, bool MIS> __device__ __forceinline__ GReservoir<N, MIS> new_g() { // newReservoir; what GLSL leaves unset is zero
	GReservoir<N, MIS> r;
	float4 *dst = reinterpret_cast<float4 *>(&r);
#pragma unroll
	for (unsigned k = 0; k < sizeof(r) / 16; ++k) {
		dst[k] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
	}
	return r;
}

// reservoir.glsl:6-26
template <int N, bool MIS>
__device__ __forceinline__ void g_update(GReservoir<N, MIS> &res, int i, float weight, f3 pos, float nx, float ny, float nz, float nw, float lum,
                                         int lightIdx, float pHat, float w, float sumPHat, Pcg32 &rng) {
	GSample<MIS> &s = res.s[i];
	s.sumWeights = s.sumWeights + weight;
	float replacePossibility = weight / s.sumWeights;
	if (pcg_float(rng) < replacePossibility) {
		s.px = pos.x; s.py = pos.y; s.pz = pos.z; s.lum = lum;
		s.nx = nx; s.ny = ny; s.nz = nz; s.nw = nw;
		s.lightIndex = lightIdx;
		s.pHat = pHat;
		s.w = w;
		if (MIS) {
			set_sum_phat(s, sum_phat(s) + sumPHat); // :22-24
		}
	}
}
// reservoir.glsl:44-64, with its call-site quirk under UNBIASED_MIS (:54-58 lists other.sumPHat before other.w, the signature
// has them the other way round): reproduced, not fixed
template <int N, bool MIS>
__device__ __forceinline__ void g_combine(GReservoir<N, MIS> &self, const GReservoir<N, MIS> &other, const float *pHat, Pcg32 &rng) {
	self.M += other.M;
#pragma unroll
	for (int i = 0; i < N; ++i) {
		const GSample<MIS> &o = other.s[i];
		float weight = (pHat[i] * o.w) * (float)other.M;
		if (weight > 0.0f) {
			if (MIS) {
				g_update(self, i, weight, mk3(o.px, o.py, o.pz), o.nx, o.ny, o.nz, o.nw, o.lum, o.lightIndex, pHat[i], sum_phat(o), o.w, rng);
			} else {
				g_update(self, i, weight, mk3(o.px, o.py, o.pz), o.nx, o.ny, o.nz, o.nw, o.lum, o.lightIndex, pHat[i], o.w, 0.0f, rng);
			}
		}
		if (self.s[i].w > 0.0f) {
			self.s[i].w = self.s[i].sumWeights / ((float)self.M * self.s[i].pHat);
		}
	}
}
template <bool MIS> __device__ __forceinline__ float g_phat(const Surface &sf, float albedoLum, const GSample<MIS> &s) {
	return evaluate_phat(sf, albedoLum, mk3(s.px, s.py, s.pz), mk3(s.nx, s.ny, s.nz), s.nw > 0.5f, s.lum);
}

// testVisibility (visibilityTest.glsl:1-4, 27-28), traced inline by the calling thread.  Returns SHADOWED.
__device__ __forceinline__ bool g_shadowed(const SceneView &sc, f3 p1, f3 p2, unsigned &overflow) {
	f3 o, d;
	segment_setup(p1, p2, o, d);
	if (sc.image != nullptr) {
		return !trace_any_image(sc.image, RESTIR_TRACE_TRI_EDGES ? sc.triEdges : sc.tris, o, d);
	}
	return !trace_any_reference(sc.nodes, sc.tris, o, d, overflow);
}

struct PixelInputs {
	f3 albedo, normal, worldPos;
	float roughness, metallic, albedoLum;
};
__device__ __forceinline__ PixelInputs fetch_pixel(const PassParams &p, size_t pix) {
	PixelInputs in;
	in.albedo = fetch_albedo(p.cur, p.scene.srgbLut, pix, nullptr);
	in.normal = fetch_normal(p.cur, pix);
	fetch_material(p.cur, pix, in.roughness, in.metallic);
	in.worldPos = fetch_world_pos(p.cur, pix);
	in.albedoLum = luminance3(in.albedo.x, in.albedo.y, in.albedo.z);
	return in;
}

} // namespace

// ---- restirOmni.glsl:86-212 ------------------------------------------------------------------------------------------
template <int N, bool MIS>
__global__ void __launch_bounds__(kThreads) generic_restir_kernel(PassParams p, void *__restrict__ out, const void *__restrict__ prev) {
	int x, y;
	const bool active = pixel_of_thread(p.band, x, y);
	unsigned rays = 0, overflow = 0, haloMiss = 0;
	if (active) {
		const SceneView &sc = p.scene;
		const size_t pix = local_index(p.band, x, y);
		const PixelInputs in = fetch_pixel(p, pix);
		const f3 cam = mk3(p.u.cameraPos[0], p.u.cameraPos[1], p.u.cameraPos[2]);
		const Surface sf = make_surface(in.worldPos, in.normal, cam, in.roughness, in.metallic);
		GReservoir<N, MIS> res = new_g<N, MIS>();                                       // :105
		Pcg32 rng = pcg_seed(p.u.frame, (uint32_t)y * 10007u + (uint32_t)x);            // :106
		if (dot3(in.normal, in.normal) != 0.0f) {                                       // :107
			const bool pointMode = sc.pointCount != 0;
			for (uint32_t c = 0; c < p.u.initialLightSampleCount; ++c) {                // :108-142
				float r1 = pcg_float(rng);
				float r2 = pcg_float(rng);
				int idx;
				float prob;
				alias_sample(sc, r1, r2, idx, prob);
				f3 lpos, ln;
				float lum, lnw;
				int lightIndex;
				if (pointMode) {
					float4 pl = __ldg(sc.pointPosLum + idx);
					lpos = mk3(pl.x, pl.y, pl.z);
					lum = pl.w;
					lightIndex = idx;
					ln = mk3(0.0f, 0.0f, 0.0f);
					lnw = 0.0f;
				} else {
					const float4 *tl = reinterpret_cast<const float4 *>(sc.triLights + idx);
					float4 a = __ldg(tl), b = __ldg(tl + 1), cc = __ldg(tl + 2), em = __ldg(tl + 3), na = __ldg(tl + 4);
					float r3 = pcg_float(rng);
					float r4 = pcg_float(rng);
					float sq = sqrtf(r3);
					lpos = (mk3(a.x, a.y, a.z) * (1.0f - sq) + mk3(b.x, b.y, b.z) * (sq * (1.0f - r4))) + mk3(cc.x, cc.y, cc.z) * (r4 * sq);
					lum = em.w;
					lightIndex = -1 - idx;
					f3 wi = normalize3(in.worldPos - lpos);
					ln = mk3(na.x, na.y, na.z);
					lnw = 1.0f;
					prob = prob / (fabsf(dot3(wi, ln)) * na.w);
				}
				float pHat = evaluate_phat(sf, in.albedoLum, lpos, ln, !pointMode, lum);
				// addSampleToReservoir, reservoir.glsl:28-42: every sample slot streams the candidate with its own draw
				float weight = pHat / prob;
				res.M += 1u;
#pragma unroll
				for (int i = 0; i < N; ++i) {
					float w = (res.s[i].sumWeights + weight) / ((float)res.M * pHat);
					g_update(res, i, weight, lpos, ln.x, ln.y, ln.z, lnw, lum, lightIndex, pHat, w, pHat, rng);
				}
			}
		}
		if ((p.u.flags & RESTIR_VISIBILITY_REUSE_FLAG) != 0) {                          // :148-160
#pragma unroll 1
			for (int i = 0; i < N; ++i) {
				bool shadowed = g_shadowed(sc, in.worldPos, mk3(res.s[i].px, res.s[i].py, res.s[i].pz), overflow);
				rays++;
				if (shadowed) {
					res.s[i].w = 0.0f;
					res.s[i].sumWeights = 0.0f;
					set_sum_phat(res.s[i], 0.0f);
				}
			}
		}
		if ((p.u.flags & RESTIR_TEMPORAL_REUSE_FLAG) != 0) {                            // :163-209
			const float *M = p.u.prevFrameProjectionViewMatrix;
			float px = ((M[0] * in.worldPos.x + M[4] * in.worldPos.y) + M[8] * in.worldPos.z) + M[12] * 1.0f;
			float py = ((M[1] * in.worldPos.x + M[5] * in.worldPos.y) + M[9] * in.worldPos.z) + M[13] * 1.0f;
			float pw = ((M[3] * in.worldPos.x + M[7] * in.worldPos.y) + M[11] * in.worldPos.z) + M[15] * 1.0f;
			float invW = 1.0f / pw;
			px = px * invW;
			py = py * invW;
			px = ((px + 1.0f) * 0.5f) * (float)p.band.W;
			py = ((py + 1.0f) * 0.5f) * (float)p.band.H;
			if (px > 0.0f && py > 0.0f && px < (float)p.band.W && py < (float)p.band.H) {
				int fx = (int)px, fy = (int)py;
				if (fy < p.band.allocBegin || fy >= p.band.allocEnd) {
					haloMiss = 1;
				} else {
					size_t ppix = local_index(p.band, fx, fy);
					f3 dp = in.worldPos - fetch_world_pos(p.prev, ppix);
					if (dot3(dp, dp) < 0.01f) {
						f3 da = in.albedo - fetch_albedo(p.prev, sc.srgbLut, ppix, nullptr);
						if (dot3(da, da) < 0.01f) {
							if (dot3(in.normal, fetch_normal(p.prev, ppix)) > 0.5f) {
								GReservoir<N, MIS> prevRes = load_g<N, MIS>(prev, ppix);
								prevRes.M = min(prevRes.M, p.u.temporalSampleCountMultiplier * res.M);
								float pHat[N];
#pragma unroll
								for (int i = 0; i < N; ++i) {
									pHat[i] = g_phat(sf, in.albedoLum, prevRes.s[i]);
								}
								g_combine(res, prevRes, pHat, rng);
							}
						}
					}
				}
			}
		}
		store_g(out, pix, res);                                                         // :211
	}
	add_counter(p.counters, kCounterRays, rays);
	add_counter(p.counters, kCounterTraced, rays);
	add_counter(p.counters, kCounterOverflow, overflow);
	add_counter(p.counters, kCounterHaloMiss, haloMiss);
}

// ---- spatialReuse.comp:30-86 -----------------------------------------------------------------------------------------
template <int N, bool MIS>
__global__ void __launch_bounds__(kThreads) generic_spatial_kernel(PassParams p, const void *__restrict__ in, void *__restrict__ out, int iter) {
	int x, y;
	const bool active = pixel_of_thread(p.band, x, y);
	unsigned haloMiss = 0;
	if (active) {
		const size_t pix = local_index(p.band, x, y);
		const PixelInputs px = fetch_pixel(p, pix);
		const float worldDepth = __ldg(p.cur.depth + pix);
		const f3 cam = mk3(p.u.cameraPos[0], p.u.cameraPos[1], p.u.cameraPos[2]);
		const Surface sf = make_surface(px.worldPos, px.normal, cam, px.roughness, px.metallic);
		float sinThr, cosThr;
		sincos_policy(p.u.spatialNormalThreshold * 0.017453292519943295f, sinThr, cosThr);
		GReservoir<N, MIS> res = load_g<N, MIS>(in, pix);
		Pcg32 rng = pcg_seed(p.u.frame * 31u + (uint32_t)iter, (uint32_t)y * 10007u + (uint32_t)x);
		for (uint32_t i = 0; i < p.u.spatialNeighbors; ++i) {
			float angle = (pcg_float(rng) * 2.0f) * RESTIR_PI_F;
			float radius = sqrtf(pcg_float(rng)) * p.u.spatialRadius;
			float sn, cs;
			sincos_policy(angle, sn, cs);
			int nx = x + (int)floorf(cs * radius), ny = y + (int)floorf(sn * radius);
			nx = max(0, min(nx, p.band.W - 1));
			ny = max(0, min(ny, p.band.H - 1));
			if (ny < p.band.allocBegin || ny >= p.band.allocEnd) {
				haloMiss = 1;
				continue;
			}
			size_t npix = local_index(p.band, nx, ny);
			float nDepth = __ldg(p.cur.depth + npix);
			f3 nNor = fetch_normal(p.cur, npix);
			if (fabsf(nDepth - worldDepth) > p.u.spatialPosThreshold * fabsf(worldDepth) || dot3(nNor, px.normal) < cosThr) {
				continue;
			}
			GReservoir<N, MIS> other = load_g<N, MIS>(in, npix);
			float pHat[N];
#pragma unroll
			for (int j = 0; j < N; ++j) {
				pHat[j] = g_phat(sf, px.albedoLum, other.s[j]);
			}
			g_combine(res, other, pHat, rng);
		}
		store_g(out, pix, res);
	}
	add_counter(p.counters, kCounterHaloMiss, haloMiss);
}

// ---- unbiasedReuse.glsl:50-185 ---------------------------------------------------------------------------------------
constexpr int kMaxNeighbors = 16;
template <int N, bool MIS>
__global__ void __launch_bounds__(kThreads) generic_unbiased_kernel(PassParams p, const void *__restrict__ in, void *__restrict__ out, int numNeighbors) {
	int x, y;
	const bool active = pixel_of_thread(p.band, x, y);
	unsigned rays = 0, overflow = 0, haloMiss = 0;
	if (active) {
		const SceneView &sc = p.scene;
		const size_t pix = local_index(p.band, x, y);
		const PixelInputs px = fetch_pixel(p, pix);
		const f3 cam = mk3(p.u.cameraPos[0], p.u.cameraPos[1], p.u.cameraPos[2]);
		const Surface sf = make_surface(px.worldPos, px.normal, cam, px.roughness, px.metallic);
		const bool vis = (p.u.flags & RESTIR_VISIBILITY_REUSE_FLAG) != 0;
		GReservoir<N, MIS> res = load_g<N, MIS>(in, pix);
		Pcg32 rng = pcg_seed(p.u.frame * 17u, (uint32_t)y * 10007u + (uint32_t)x);      // :72
		int npx[kMaxNeighbors];
		uint32_t neighborM[kMaxNeighbors];
		float neighborSumPHat[N][kMaxNeighbors];                                        // :74-82
		float originalSumPHat[N];
		const uint32_t originalM = res.M;
#pragma unroll
		for (int i = 0; i < N; ++i) {
			originalSumPHat[i] = sum_phat(res.s[i]);
		}
		for (int i = 0; i < numNeighbors; ++i) {                                        // :84-124
			float angle = (pcg_float(rng) * 2.0f) * RESTIR_PI_F;
			float radius = sqrtf(pcg_float(rng)) * p.u.spatialRadius;
			float sn, cs;
			sincos_policy(angle, sn, cs);
			int nx = x + (int)roundf(cs * radius), ny = y + (int)roundf(sn * radius);
			nx = max(0, min(nx, p.band.W - 1));
			ny = max(0, min(ny, p.band.H - 1));
			npx[i] = -1;
			neighborM[i] = 0u;
#pragma unroll
			for (int j = 0; j < N; ++j) {
				neighborSumPHat[j][i] = 0.0f;
			}
			if (ny < p.band.allocBegin || ny >= p.band.allocEnd) {
				haloMiss = 1;
				continue;
			}
			size_t npix = local_index(p.band, nx, ny);
			GReservoir<N, MIS> other = load_g<N, MIS>(in, npix);
			npx[i] = (int)npix;
			neighborM[i] = other.M;
			res.M += other.M;                                                           // :104
#pragma unroll
			for (int j = 0; j < N; ++j) {
				const GSample<MIS> &o = other.s[j];
				neighborSumPHat[j][i] = sum_phat(o);
				float newPHat = g_phat(sf, px.albedoLum, o);
				float weight = (newPHat * o.w) * (float)other.M;
				if (weight > 0.0f) {
					g_update(res, j, weight, mk3(o.px, o.py, o.pz), o.nx, o.ny, o.nz, o.nw, o.lum, o.lightIndex, newPHat, o.w, sum_phat(o), rng); // :108-121
				}
			}
		}
#pragma unroll 1
		for (int i = 0; i < N; ++i) {                                                   // :132-182
			GSample<MIS> &s = res.s[i];
			const f3 lightPos = mk3(s.px, s.py, s.pz);
			float sumPHat = originalSumPHat[i];
			uint32_t numSamples = originalM;
			for (int j = 0; j < numNeighbors; ++j) {
				if (npx[j] < 0) {
					continue;
				}
				f3 nPos = fetch_world_pos(p.cur, (size_t)npx[j]);
				f3 nNor = fetch_normal(p.cur, (size_t)npx[j]);
				if (dot3(lightPos - nPos, nNor) < 0.0f) {
					continue;
				}
				if (vis) {
					bool shadowed = g_shadowed(sc, nPos, lightPos, overflow);
					rays++;
					if (shadowed) {
						continue;
					}
				}
				sumPHat = sumPHat + neighborSumPHat[i][j];
				numSamples += neighborM[j];
			}
			if (vis) {
				bool shadowed = g_shadowed(sc, px.worldPos, lightPos, overflow);
				rays++;
				if (shadowed) {
					sumPHat = 0.0f;
					numSamples = 0u;
				}
			}
			if (MIS ? (sumPHat > 0.0f) : (numSamples > 0u)) {
				if (MIS) {
					s.w = (s.sumWeights * s.pHat) / (sumPHat * s.pHat);                 // :169
				} else {
					s.w = s.sumWeights / ((float)numSamples * s.pHat);
				}
			} else {
				s.w = 0.0f;
				s.sumWeights = 0.0f;
				set_sum_phat(s, 0.0f);
			}
		}
		store_g(out, pix, res);
	}
	add_counter(p.counters, kCounterRays, rays);
	add_counter(p.counters, kCounterTraced, rays);
	add_counter(p.counters, kCounterOverflow, overflow);
	add_counter(p.counters, kCounterHaloMiss, haloMiss);
}

// ---- lighting.frag:43-71,103 -----------------------------------------------------------------------------------------
template <int N, bool MIS>
__global__ void __launch_bounds__(kThreads) generic_lighting_kernel(PassParams p, restir_lighting_uniforms lu, const void *__restrict__ reservoirs,
                                                                   void *__restrict__ outPixels, int outFormat) {
	int x, y;
	if (!pixel_of_thread(p.band, x, y)) {
		return;
	}
	const SceneView &sc = p.scene;
	const size_t pix = local_index(p.band, x, y);
	float albedoA;
	f3 albedo = fetch_albedo(p.cur, sc.srgbLut, pix, &albedoA);
	f3 normal = fetch_normal(p.cur, pix);
	float roughness, metallic;
	fetch_material(p.cur, pix, roughness, metallic);
	f3 worldPos = fetch_world_pos(p.cur, pix);
	const Surface sf = make_surface(worldPos, normal, mk3(lu.cameraPos[0], lu.cameraPos[1], lu.cameraPos[2]), roughness, metallic);
	const GReservoir<N, MIS> r = load_g<N, MIS>(reservoirs, pix);
	f3 c = mk3(0.0f, 0.0f, 0.0f);
#pragma unroll
	for (int i = 0; i < N; ++i) {                                                       // :53-67
		const GSample<MIS> &s = r.s[i];
		f3 emission = mk3(0.0f, 0.0f, 0.0f);
		if (s.lightIndex < 0) {
			int ti = -1 - s.lightIndex;
			if (ti < sc.triCount) {
				float4 e = __ldg(reinterpret_cast<const float4 *>(sc.triLights + ti) + 3);
				emission = mk3(e.x, e.y, e.z);
			}
		} else if (s.lightIndex < sc.pointCount) {
			float4 e = __ldg(reinterpret_cast<const float4 *>(sc.pointLights + s.lightIndex) + 1);
			emission = mk3(e.x, e.y, e.z);
		}
		c = c + evaluate_phat_full(sf, albedo, mk3(s.px, s.py, s.pz), mk3(s.nx, s.ny, s.nz), s.nw > 0.5f, emission) * s.w;
	}
	c = c * (1.0f / (float)N);                                                          // :68, P3
	if (albedoA > 0.5f) {
		c = albedo;
	}
	if (lu.gamma != 1.0f) {
		float e = 1.0f / lu.gamma;
		c = mk3(powf(c.x, e), powf(c.y, e), powf(c.z, e));
	}
	if (outFormat == 0) {
		reinterpret_cast<float4 *>(outPixels)[pix] = make_float4(c.x, c.y, c.z, 1.0f);
	} else {
		uchar4 q;
		q.x = (unsigned char)rintf(srgb_encode(c.x) * 255.0f);
		q.y = (unsigned char)rintf(srgb_encode(c.y) * 255.0f);
		q.z = (unsigned char)rintf(srgb_encode(c.z) * 255.0f);
		q.w = 255;
		reinterpret_cast<uchar4 *>(outPixels)[pix] = q;
	}
}

// ---- launchers -------------------------------------------------------------------------------------------------------

namespace {
dim3 tiles(const Band &b) { return dim3((unsigned)((b.W + kTileW - 1) / kTileW), (unsigned)((b.rowEnd - b.rowBegin + kTileH - 1) / kTileH), 1); }
} // namespace

#define RESTIR_VARIANT_SWITCH(N_, MIS_, CALL)        \
	switch ((N_) * 2 + ((MIS_) ? 1 : 0)) {           \
	case 2: CALL(1, false); break;                   \
	case 3: CALL(1, true); break;                    \
	case 4: CALL(2, false); break;                   \
	case 5: CALL(2, true); break;                    \
	case 8: CALL(4, false); break;                   \
	case 9: CALL(4, true); break;                    \
	default: return false;                           \
	}

bool generic_variant_supported(int n, bool mis) { return n == 1 || n == 2 || n == 4; (void)mis; }
size_t generic_reservoir_bytes(int n, bool mis) { return (size_t)n * (mis ? 64 : 48) + 16; }

bool launch_generic_restir(int n, bool mis, const PassParams &p, void *out, const void *prev, cudaStream_t s) {
#define CALL(N, M) generic_restir_kernel<N, M><<<tiles(p.band), kThreads, 0, s>>>(p, out, prev)
	RESTIR_VARIANT_SWITCH(n, mis, CALL)
#undef CALL
	return true;
}
bool launch_generic_spatial(int n, bool mis, const PassParams &p, const void *in, void *out, int iter, cudaStream_t s) {
#define CALL(N, M) generic_spatial_kernel<N, M><<<tiles(p.band), kThreads, 0, s>>>(p, in, out, iter)
	RESTIR_VARIANT_SWITCH(n, mis, CALL)
#undef CALL
	return true;
}
bool launch_generic_unbiased(int n, bool mis, const PassParams &p, const void *in, void *out, int numNeighbors, cudaStream_t s) {
#define CALL(N, M) generic_unbiased_kernel<N, M><<<tiles(p.band), kThreads, 0, s>>>(p, in, out, numNeighbors)
	RESTIR_VARIANT_SWITCH(n, mis, CALL)
#undef CALL
	return true;
}
bool launch_generic_lighting(int n, bool mis, const PassParams &p, const restir_lighting_uniforms &lu, const void *res, void *outPixels, int fmt,
                             cudaStream_t s) {
#define CALL(N, M) generic_lighting_kernel<N, M><<<tiles(p.band), kThreads, 0, s>>>(p, lu, res, outPixels, fmt)
	RESTIR_VARIANT_SWITCH(n, mis, CALL)
#undef CALL
	return true;
}

cudaError_t preload_generic_kernels() {
	cudaFuncAttributes a;
	cudaError_t e = cudaSuccess;
#define LOAD(N, M)                                                                                   \
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, generic_restir_kernel<N, M>);                 \
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, generic_spatial_kernel<N, M>);                \
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, generic_unbiased_kernel<N, M>);               \
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, generic_lighting_kernel<N, M>)
	LOAD(1, false);
	LOAD(1, true);
	LOAD(2, false);
	LOAD(2, true);
	LOAD(4, false);
	LOAD(4, true);
#undef LOAD
	return e;
}

} // namespace restir
