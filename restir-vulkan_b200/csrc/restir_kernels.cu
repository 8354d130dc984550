// restir_kernels.cu — hand-written sm_100a per-pixel kernels of the ReSTIR resampling path.
//
//   omni_candidates_kernel    <- src/shaders/restirOmni.glsl:86-145   (RIS over the alias table)
//   omni_temporal_kernel      <- src/shaders/restirOmni.glsl:148-212  (apply the visibility bit, temporal reuse)
//   spatial_reuse_kernel      <- src/shaders/spatialReuse.comp:30-86
//   unbiased_merge_kernel     <- src/shaders/unbiasedReuse.glsl:50-124
//   unbiased_finalize_kernel  <- src/shaders/unbiasedReuse.glsl:126-185
//   lighting_kernel           <- src/shaders/lighting.frag:43-71,103  (debugMode 0)
//   raycast_gbuffer_kernel    fixture tool (primary visibility through the same tree)
//
// The reference traces its shadow rays inline in restirOmni / unbiasedReuse.  Here each of those shaders is
// cut at its testVisibility calls: the part before writes the (not yet visibility-tested) reservoir, the
// persistent trace kernel (restir_trace.cu) answers every ray of the pass with full warps, and the part
// after consumes one byte per ray.  The RNG stream of a pixel continues across the cut by an LCG jump.
//
// One thread shades one pixel, as in the reference, but a CTA covers a 32x8 screen tile made of
// eight 8x4 warp tiles (the reference uses 64x1 workgroups, restirStructs.glsl:10-14): the G-buffer /
// reservoir rows a warp touches are whole 32-byte sectors and the rays of a tile are neighbours in the trace
// kernel's work list.  Reservoirs live in HBM as 32-byte PackedReservoir records.

#include "restir_device.cuh"
#include "restir_kernels.h"
#include "restir_pixel.cuh"

namespace restir {

// ------------------------------------------------------------------------------------------------
// packed reservoir I/O: two 16-byte transactions per record
__device__ __forceinline__ PackedReservoir load_reservoir(const PackedReservoir *buf, size_t i) {
	const float4 *p = reinterpret_cast<const float4 *>(buf + i);
	float4 a = __ldg(p), b = __ldg(p + 1);
	PackedReservoir r;
	r.px = a.x; r.py = a.y; r.pz = a.z; r.lightIndex = __float_as_int(a.w);
	r.pHat = b.x; r.sumWeights = b.y; r.w = b.z; r.M = __float_as_uint(b.w);
	return r;
}
// same, through the coherent path: for a buffer this pass's earlier kernel wrote
__device__ __forceinline__ PackedReservoir load_reservoir_plain(const PackedReservoir *buf, size_t i) {
	const float4 *p = reinterpret_cast<const float4 *>(buf + i);
	float4 a = p[0], b = p[1];
	PackedReservoir r;
	r.px = a.x; r.py = a.y; r.pz = a.z; r.lightIndex = __float_as_int(a.w);
	r.pHat = b.x; r.sumWeights = b.y; r.w = b.z; r.M = __float_as_uint(b.w);
	return r;
}
__device__ __forceinline__ void store_reservoir(PackedReservoir *buf, size_t i, const PackedReservoir &r) {
	float4 *p = reinterpret_cast<float4 *>(buf + i);
	p[0] = make_float4(r.px, r.py, r.pz, __int_as_float(r.lightIndex));
	p[1] = make_float4(r.pHat, r.sumWeights, r.w, __uint_as_float(r.M));
}

// normal / useLightNormal / emissionLum of a stored sample, re-read from the light tables
// (restirOmni.glsl:117-133 wrote exactly these values into the reference's 64-byte reservoir).
__device__ __forceinline__ void sample_light_attrs(const SceneView &sc, const PackedReservoir &r, f3 &n, bool &useN, float &lum) {
	if (r.pHat == 0.0f) { // never selected: all-zero sample (oracle definition of reservoir.glsl:66-76)
		n = mk3(0.0f, 0.0f, 0.0f);
		useN = false;
		lum = 0.0f;
	} else if (r.lightIndex >= 0) {
		n = mk3(0.0f, 0.0f, 0.0f);
		useN = false;
		lum = r.lightIndex < sc.pointCount ? __ldg(sc.pointPosLum + r.lightIndex).w : 0.0f; // uploaded reservoirs may name any index
	} else if (-1 - r.lightIndex >= sc.triCount) {
		n = mk3(0.0f, 0.0f, 0.0f);
		useN = true;
		lum = 0.0f;
	} else {
		float4 a = __ldg(sc.triAux + (-1 - r.lightIndex));
		n = mk3(a.x, a.y, a.z);
		useN = true;
		lum = a.w;
	}
}

// reservoir.glsl:6-26 on the packed fields.  One RNG draw, always.
__device__ __forceinline__ void update_reservoir(PackedReservoir &res, float weight, f3 pos, int lightIdx, float pHat, float w, Pcg32 &rng) {
	res.sumWeights = res.sumWeights + weight;
	float replacePossibility = weight / res.sumWeights;
	if (pcg_float(rng) < replacePossibility) {
		res.px = pos.x; res.py = pos.y; res.pz = pos.z;
		res.lightIndex = lightIdx;
		res.pHat = pHat;
		res.w = w;
	}
}

// reservoir.glsl:44-64.  evaluatePHat of the other sample is only needed when the merge weight can be
// positive: weight = (pHat * other.w) * other.M is +-0 or NaN whenever other.w == 0 or other.M == 0, and
// then neither the update nor its RNG draw happens — an exact shortcut, not an approximation.
__device__ __forceinline__ void combine_reservoirs(PackedReservoir &self, const PackedReservoir &other, const SceneView &sc,
                                                   const Surface &sf, float albedoLum, Pcg32 &rng) {
	self.M += other.M;
	if (other.w != 0.0f && other.M != 0u) {
		f3 n; bool useN; float lum;
		sample_light_attrs(sc, other, n, useN, lum);
		float pHat = evaluate_phat(sf, albedoLum, mk3(other.px, other.py, other.pz), n, useN, lum);
		float weight = (pHat * other.w) * (float)other.M;
		if (weight > 0.0f) {
			update_reservoir(self, weight, mk3(other.px, other.py, other.pz), other.lightIndex, pHat, other.w, rng);
		}
	}
	if (self.w > 0.0f) {
		self.w = self.sumWeights / ((float)self.M * self.pHat);
	}
}

__device__ __forceinline__ void shade_pixel(const PassParams &p, const restir_lighting_uniforms &lu, size_t pix, const PackedReservoir &r,
                                            void *__restrict__ outPixels, int outFormat);

// ------------------------------------------------------------------------------------------------
// PCG32 stream position after n draws: state_n = A_n * state + G_n * inc (LCG jump-ahead), with A_n, G_n
// computed once per launch on the host (lcg_jump).  Lets the second half of a cut shader continue the
// pixel's RNG stream (rand.glsl:12-18) without storing it.
struct LcgJump {
	uint64_t A, G;
};
static LcgJump lcg_jump(uint64_t n) {
	uint64_t accA = 1, accG = 0, curA = 6364136223846793005ull, curG = 1;
	while (n) {
		if (n & 1) {
			accG = accG * curA + curG;
			accA = accA * curA;
		}
		curG = (curA + 1) * curG;
		curA = curA * curA;
		n >>= 1;
	}
	return LcgJump{accA, accG};
}

// draws consumed by restirOmni.glsl:108-142 per candidate: 2 (alias table) [+ 2 (point on triangle)] + 1 (reservoir update)
__host__ __device__ inline uint32_t draws_per_candidate(bool pointMode) { return pointMode ? 3u : 5u; }

// ------------------------------------------------------------------------------------------------
// restirOmni.glsl:86-145: candidate generation and streaming RIS.  Writes the reservoir BEFORE the
// visibility test (:148-160) and temporal reuse (:163-209), which omni_temporal_kernel applies.
//
// Exact shortcuts (same bits as the straightforward evaluation, tests/test_gpu_parity.py):
//   * `w` of addSampleToReservoir (reservoir.glsl:33) is only observable for the candidate that ends up
//     selected, so the division is done once after the loop from the (sumWeights, M) recorded at selection;
//   * a point light behind the surface has pHat = +0 (restirUtils.glsl:8-10) and, for prob > 0, weight +0:
//     sumWeights and the selection are unchanged and only M and the RNG advance (reservoir.glsl:6-26).
//
// RESTIR_CANDIDATES_SKIP_AHEAD 1 (experiment, off; point lights): such candidates cost ~25 instructions, the others ~290,
// and which is which differs from lane to lane: in the plain loop a lane whose light is behind its surface idles through
// the other lanes' evaluation (r1 capture M: 23.8 of 32 lanes).  With the switch on every lane first runs ahead, on its
// own, over the candidates that need no evaluation until it holds one that does, then the warp evaluates one candidate
// per lane together: the expensive part runs max-over-lanes(front-facing candidates) times instead of `count` times.
// Measured on B200 (profiles/r2_b_summary.md): 0.708 -> 0.950 ms on Sponza / 200 lights — with a quarter of the
// candidates back-facing, almost every round has some lane that runs ahead two or three times, so the ~35-instruction
// draw-and-fetch part is issued 2.5 times per round at a few lanes, which costs more than the 4 rounds it saves.
#ifndef RESTIR_TEMPORAL_MIN_BLOCKS
#define RESTIR_TEMPORAL_MIN_BLOCKS 5 // latency-bound gathers: 0.111 -> 0.099 ms; the candidate loop is issue-bound and loses with fewer registers (0.707 -> 0.744 at 5)
#endif
#ifndef RESTIR_CANDIDATES_SKIP_AHEAD
#define RESTIR_CANDIDATES_SKIP_AHEAD 0
#endif
// no minimum CTA count here: the loop is issue-bound, 64 registers (4 CTAs/SM) is what ptxas picks on its own and both 72 and 51 lose
#ifndef RESTIR_CANDIDATES_MIN_BLOCKS
#define RESTIR_CANDIDATES_MIN_BLOCKS 4 // 64 registers: 0.6085 -> 0.602 ms against the 70 ptxas takes when left alone
#endif
__global__ void __launch_bounds__(kThreads, RESTIR_CANDIDATES_MIN_BLOCKS) omni_candidates_kernel(PassParams p, PackedReservoir *__restrict__ out) {
	int x, y;
	const bool inside = pixel_of_thread(p.band, x, y); // no early return: the skip-ahead loop votes with the whole warp
	const SceneView &sc = p.scene;
	const size_t pix = inside ? local_index(p.band, x, y) : 0;
	f3 normal = mk3(0.0f, 0.0f, 0.0f), worldPos = normal, albedo = normal;
	float roughness = 0.0f, metallic = 0.0f;
	if (inside) {
		albedo = fetch_albedo(p.cur, sc.srgbLut, pix, nullptr);                 // :98-101
		normal = fetch_normal(p.cur, pix);
		fetch_material(p.cur, pix, roughness, metallic);
		worldPos = fetch_world_pos(p.cur, pix);
	}
	float albedoLum = luminance3(albedo.x, albedo.y, albedo.z);                 // :103
	f3 cam = mk3(p.u.cameraPos[0], p.u.cameraPos[1], p.u.cameraPos[2]);
	Surface sf = make_surface(worldPos, normal, cam, roughness, metallic);

	PackedReservoir res;                                                      // :105
	res.px = res.py = res.pz = 0.0f;
	res.lightIndex = 0;
	res.pHat = res.sumWeights = res.w = 0.0f;
	res.M = 0u;
	const bool lit = inside && dot3(normal, normal) != 0.0f;                  // :107
	Pcg32 rng = pcg_seed(p.u.frame, (uint32_t)y * 10007u + (uint32_t)x);      // :106
	const uint32_t count = lit ? p.u.initialLightSampleCount : 0u;
	const bool pointMode = sc.pointCount != 0;
	float selSum = 0.0f;
	uint32_t selM = 0u;
	if (pointMode && RESTIR_CANDIDATES_SKIP_AHEAD) {
		uint32_t i = 0;
		for (;;) {
			// run ahead over the candidates that only move M and the RNG, :116-122
			bool have = false;
			int idx = 0;
			float prob = 0.0f;
			float4 pl = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
			while (i < count) {
				float r1 = pcg_float(rng);
				float r2 = pcg_float(rng);
				alias_sample(sc, r1, r2, idx, prob);
				pl = __ldg(sc.pointPosLum + idx);
				++i;
				res.M += 1u;
				if (dot3(mk3(pl.x, pl.y, pl.z) - worldPos, normal) < 0.0f && prob > 0.0f) { // pHat = +0, weight = +0
					pcg_next(rng);
					continue;
				}
				have = true;
				break;
			}
			if (!__any_sync(0xffffffffu, have)) {
				break;
			}
			if (have) {
				f3 lpos = mk3(pl.x, pl.y, pl.z);
				float pHat = evaluate_phat(sf, albedoLum, lpos, mk3(0.0f, 0.0f, 0.0f), false, pl.w); // :135-139
				float weight = pHat / prob;                                       // reservoir.glsl:28-42, 6-26
				res.sumWeights = res.sumWeights + weight;
				float replacePossibility = weight / res.sumWeights;
				if (pcg_float(rng) < replacePossibility) {
					res.px = lpos.x; res.py = lpos.y; res.pz = lpos.z;
					res.lightIndex = idx;
					res.pHat = pHat;
					selSum = res.sumWeights;
					selM = res.M;
				}
			}
		}
	} else {
		for (uint32_t i = 0; i < count; ++i) {                                // :108-142
			float r1 = pcg_float(rng);
			float r2 = pcg_float(rng);
			int idx;
			float prob;
			alias_sample(sc, r1, r2, idx, prob);
			f3 lpos, ln;
			float lum;
			int lightIndex;
			if (pointMode) {                                                  // :116-122
				float4 pl = __ldg(sc.pointPosLum + idx);
				lpos = mk3(pl.x, pl.y, pl.z);
				lum = pl.w;
				lightIndex = idx;
				ln = mk3(0.0f, 0.0f, 0.0f);
				if (dot3(lpos - worldPos, normal) < 0.0f && prob > 0.0f) {    // pHat = +0, weight = +0: nothing but M and the RNG move
					res.M += 1u;
					pcg_next(rng);
					continue;
				}
			} else {                                                          // :123-133
				const float4 *tl = reinterpret_cast<const float4 *>(sc.triLights + idx);
				float4 a = __ldg(tl), b = __ldg(tl + 1), c = __ldg(tl + 2), em = __ldg(tl + 3), na = __ldg(tl + 4);
				float r3 = pcg_float(rng);
				float r4 = pcg_float(rng);
				float sq = sqrtf(r3);                                          // pickPointOnTriangle :68-71
				lpos = (mk3(a.x, a.y, a.z) * (1.0f - sq) + mk3(b.x, b.y, b.z) * (sq * (1.0f - r4))) + mk3(c.x, c.y, c.z) * (r4 * sq);
				lum = em.w;
				lightIndex = -1 - idx;
				f3 wi = normalize3(worldPos - lpos);
				ln = mk3(na.x, na.y, na.z);
				prob = prob / (fabsf(dot3(wi, ln)) * na.w);
			}
			// evaluatePHat (:135-139), then addSampleToReservoir + updateReservoirAt (reservoir.glsl:28-42, 6-26): eleven divisions,
			// reciprocals and square roots.  Point lights: the divider's fast sequences without their wrappers, validated once
			// (restir_math.cuh SpecOps), the ordinary operators when an operand was out of the range in which those sequences are the
			// divider's own.  Triangle lights keep the ordinary operators: measured, the speculative path loses there (office 2160p,
			// 54 198 lights: 4.67 -> 5.00 ms; half of the candidates lie behind the surface and leave the evaluation at its first
			// branch, and the lanes that stay pay the validation on top of the light-normal terms).
			float pHat, weight, sum, replacePossibility;
			if (pointMode) {
				OpGuard guard;
				pHat = evaluate_phat_t<SpecOps>(sf, albedoLum, lpos, ln, false, lum, guard);
				SpecOps::check(pHat, guard);
				SpecOps::check(prob, guard);
				weight = SpecOps::div(pHat, prob);
				sum = res.sumWeights + weight;
				SpecOps::check(weight, guard);
				SpecOps::check(sum, guard);
				replacePossibility = SpecOps::div(weight, sum);
				if (!guard.ok()) {
					pHat = evaluate_phat_call(sf, albedoLum, lpos, ln, false, lum);
					weight = pHat / prob;
					sum = res.sumWeights + weight;
					replacePossibility = weight / sum;
				}
			} else {
				pHat = evaluate_phat(sf, albedoLum, lpos, ln, true, lum);
				if (pHat == 0.0f && prob > 0.0f) {
					// Half of the triangle-light candidates lie behind the surface or face away (office: 54 198 emissive triangles all
					// around).  weight = +-0 / prob is +-0 for any prob > 0 (infinity included), sumWeights + +-0 is sumWeights, and
					// +-0 / sum is +-0 or NaN: never above a draw — only M and the RNG move (reservoir.glsl:6-42), as for a point light
					// behind the surface.  Not only the two divisions go: a zero numerator takes the divider's ~100-instruction slow
					// path (capture Q, office 2160p: 15 % of this kernel's instructions were those calls, at 10 of 32 lanes).
					res.M += 1u;
					pcg_next(rng);
					continue;
				}
				weight = pHat / prob;
				sum = res.sumWeights + weight;
				replacePossibility = weight / sum;
			}
			res.M += 1u;
			res.sumWeights = sum;
			if (pcg_float(rng) < replacePossibility) {
				res.px = lpos.x; res.py = lpos.y; res.pz = lpos.z;
				res.lightIndex = lightIndex;
				res.pHat = pHat;
				selSum = res.sumWeights;
				selM = res.M;
			}
		}
	}
	if (selM != 0u) { // w = (sumWeights + weight) / (M * pHat) as of the selection, reservoir.glsl:33
		res.w = selSum / ((float)selM * res.pHat);
	}
	if (inside) {
		store_reservoir(out, pix, res);                                       // handed to the trace kernel and omni_temporal_kernel
	}
}

// restirOmni.glsl:148-212 on the reservoir omni_candidates_kernel wrote.
__global__ void __launch_bounds__(kThreads, RESTIR_TEMPORAL_MIN_BLOCKS) omni_temporal_kernel(PassParams p, PackedReservoir *__restrict__ out,
                                                                const PackedReservoir *__restrict__ prevReservoirs,
                                                                const unsigned char *__restrict__ shadowed, LcgJump jump) {
	int x, y;
	bool active = pixel_of_thread(p.band, x, y);
	unsigned haloMiss = 0;
	if (active) {
		const SceneView &sc = p.scene;
		size_t pix = local_index(p.band, x, y);
		PackedReservoir res = load_reservoir_plain(out, pix);
		bool dirty = false;
		// visibility reuse, :148-160 (M is kept)
		if ((p.u.flags & RESTIR_VISIBILITY_REUSE_FLAG) != 0 && shadowed[tile_pixel_id()] != 0) {
			res.w = 0.0f;
			res.sumWeights = 0.0f;
			dirty = true;
		}
		// temporal reuse, :163-209.  A background pixel (normal == 0, gBufferPass.cpp:117-123) can never pass the
		// normal gate (:183, dot(0, n') = 0 > 0.5 is false), so its reprojection — which lands wherever the world
		// origin projects to, usually far outside a row band — is not even looked up.
		f3 normal = fetch_normal(p.cur, pix);
		if ((p.u.flags & RESTIR_TEMPORAL_REUSE_FLAG) != 0 && dot3(normal, normal) != 0.0f) {
			f3 worldPos = fetch_world_pos(p.cur, pix);
			const float *M = p.u.prevFrameProjectionViewMatrix;
			float px = ((M[0] * worldPos.x + M[4] * worldPos.y) + M[8] * worldPos.z) + M[12] * 1.0f;
			float py = ((M[1] * worldPos.x + M[5] * worldPos.y) + M[9] * worldPos.z) + M[13] * 1.0f;
			float pw = ((M[3] * worldPos.x + M[7] * worldPos.y) + M[11] * worldPos.z) + M[15] * 1.0f;
			float invW = 1.0f / pw;
			px = px * invW;
			py = py * invW;
			px = ((px + 1.0f) * 0.5f) * (float)p.band.W;
			py = ((py + 1.0f) * 0.5f) * (float)p.band.H;
			if (px > 0.0f && py > 0.0f && px < (float)p.band.W && py < (float)p.band.H) {
				int fx = (int)px, fy = (int)py;
				if (fy < p.band.allocBegin || fy >= p.band.allocEnd) {
					haloMiss = 1; // band too thin for this camera motion: reported, never silent
				} else {
					size_t ppix = local_index(p.band, fx, fy);
					// everything the gates and the merge read at the reprojected pixel is requested at once (one round trip
					// to L2/HBM instead of four dependent ones); the gates then run in the reference's order
					f3 prevPos = fetch_world_pos(p.prev, ppix);
					f3 prevAlbedo = fetch_albedo(p.prev, sc.srgbLut, ppix, nullptr);
					f3 prevNormal = fetch_normal(p.prev, ppix);
					PackedReservoir prevRes = load_reservoir(prevReservoirs, ppix);
					f3 albedo = fetch_albedo(p.cur, sc.srgbLut, pix, nullptr);
					float roughness, metallic;
					fetch_material(p.cur, pix, roughness, metallic);
					f3 dp = worldPos - prevPos;
					if (dot3(dp, dp) < 0.01f) {
						f3 da = albedo - prevAlbedo;
						if (dot3(da, da) < 0.01f) {
							float nd = dot3(normal, prevNormal);
							if (nd > 0.5f) {
								float albedoLum = luminance3(albedo.x, albedo.y, albedo.z);
								f3 cam = mk3(p.u.cameraPos[0], p.u.cameraPos[1], p.u.cameraPos[2]);
								Surface sf = make_surface(worldPos, normal, cam, roughness, metallic);
								// the pixel's RNG stream, advanced past the candidate loop's draws (:106-142)
								Pcg32 rng = pcg_seed(p.u.frame, (uint32_t)y * 10007u + (uint32_t)x);
								rng.state = jump.A * rng.state + jump.G * rng.inc;
								prevRes.M = min(prevRes.M, p.u.temporalSampleCountMultiplier * res.M); // :189-191
								combine_reservoirs(res, prevRes, sc, sf, albedoLum, rng);
								dirty = true;
							}
						}
					}
				}
			}
		}
		if (dirty) {
			store_reservoir(out, pix, res);                                       // :211
		}
	}
	add_counter(p.counters, kCounterHaloMiss, haloMiss);
}

// ------------------------------------------------------------------------------------------------
// spatialReuse.comp:30-86
#ifndef RESTIR_SPATIAL_MIN_BLOCKS
#define RESTIR_SPATIAL_MIN_BLOCKS 4 // same reasoning: 0.253 -> 0.222 ms per pass
#endif
// LIGHT: the lighting pass (lighting.frag, debugMode 0) of the same pixel follows from registers — the tail of
// restir_frame_lit: lighting_kernel would re-read the reservoir just written and the G-buffer texels just used.
template <bool LIGHT>
__global__ void __launch_bounds__(kThreads, RESTIR_SPATIAL_MIN_BLOCKS) spatial_reuse_kernel(PassParams p, const PackedReservoir *__restrict__ in,
                                                                PackedReservoir *__restrict__ out, int iter, restir_lighting_uniforms lu,
                                                                void *__restrict__ outPixels, int outFormat) {
	int x, y;
	bool active = pixel_of_thread(p.band, x, y);
	unsigned haloMiss = 0;
	if (active) {
		const SceneView &sc = p.scene;
		size_t pix = local_index(p.band, x, y);
		f3 albedo = fetch_albedo(p.cur, sc.srgbLut, pix, nullptr);
		f3 normal = fetch_normal(p.cur, pix);
		float roughness, metallic;
		fetch_material(p.cur, pix, roughness, metallic);
		f3 worldPos = fetch_world_pos(p.cur, pix);
		float worldDepth = __ldg(p.cur.depth + pix);
		float albedoLum = luminance3(albedo.x, albedo.y, albedo.z);
		f3 cam = mk3(p.u.cameraPos[0], p.u.cameraPos[1], p.u.cameraPos[2]);
		Surface sf = make_surface(worldPos, normal, cam, roughness, metallic);
		float sinThr, cosThr;
		sincos_policy(p.u.spatialNormalThreshold * 0.017453292519943295f, sinThr, cosThr); // :67

		PackedReservoir res = load_reservoir(in, pix);
		Pcg32 rng = pcg_seed(p.u.frame * 31u + (uint32_t)iter, (uint32_t)y * 10007u + (uint32_t)x); // :47
		const uint32_t k = p.u.spatialNeighbors;
		for (uint32_t i = 0; i < k; ++i) {
			float angle = (pcg_float(rng) * 2.0f) * RESTIR_PI_F;                  // :52
			float radius = sqrtf(pcg_float(rng)) * p.u.spatialRadius;             // :53
			float sn, cs;
			sincos_policy(angle, sn, cs);
			int nx = x + (int)floorf(cs * radius), ny = y + (int)floorf(sn * radius); // :55-57
			nx = max(0, min(nx, p.band.W - 1));
			ny = max(0, min(ny, p.band.H - 1));
			if (ny < p.band.allocBegin || ny >= p.band.allocEnd) {
				haloMiss = 1;
				continue;
			}
			size_t npix = local_index(p.band, nx, ny);
			float nDepth = __ldg(p.cur.depth + npix);
			f3 nNor = fetch_normal(p.cur, npix);
			if (fabsf(nDepth - worldDepth) > p.u.spatialPosThreshold * fabsf(worldDepth) || dot3(nNor, normal) < cosThr) { // :65-70
				continue;
			}
			PackedReservoir other = load_reservoir(in, npix);
			combine_reservoirs(res, other, sc, sf, albedoLum, rng);               // :72-83
		}
		store_reservoir(out, pix, res);
		if (LIGHT) {
			shade_pixel(p, lu, pix, res, outPixels, outFormat);
		}
	}
	add_counter(p.counters, kCounterHaloMiss, haloMiss);
}

// The same pass with the gate data staged in shared memory (north-star design): the CTA first copies the raw depth
// and the packed normal of its 32x8 tile and the +-31-pixel apron around it ((32 + 62) x (8 + 62) texels x 12 bytes = 79 KB),
// then every neighbour's gate (:62-70) reads shared memory; the reservoirs of accepted neighbours are still gathered from
// L2.  Same arithmetic, same bits.  A/B against the direct kernel: profiles/r2_e_summary.md; selected by
// restir_set_spatial_staging (default: whichever won).
constexpr int kApron = 31, kStageW = kTileW + 2 * kApron, kStageH = kTileH + 2 * kApron;
constexpr size_t kStageBytes = (size_t)kStageW * kStageH * (sizeof(float) + sizeof(short4));
template <bool LIGHT>
__global__ void __launch_bounds__(kThreads, 2) spatial_reuse_staged_kernel(PassParams p, const PackedReservoir *__restrict__ in,
                                                                         PackedReservoir *__restrict__ out, int iter, restir_lighting_uniforms lu,
                                                                         void *__restrict__ outPixels, int outFormat) {
	extern __shared__ __align__(16) unsigned char stage[];
	short4 *sNormal = reinterpret_cast<short4 *>(stage);
	float *sDepth = reinterpret_cast<float *>(stage + (size_t)kStageW * kStageH * sizeof(short4));
	const int x0 = (int)blockIdx.x * kTileW - kApron, y0 = p.band.rowBegin + (int)blockIdx.y * kTileH - kApron;
	for (int t = threadIdx.x; t < kStageW * kStageH; t += kThreads) {
		int sy = t / kStageW, sx = t - sy * kStageW;
		int gx = max(0, min(x0 + sx, p.band.W - 1)), gy = max(0, min(y0 + sy, p.band.H - 1));
		if (gy >= p.band.allocBegin && gy < p.band.allocEnd) { // rows outside the band's memory are never read back (halo miss, below)
			size_t g = local_index(p.band, gx, gy);
			sNormal[t] = __ldg(p.cur.normal + g);
			sDepth[t] = __ldg(p.cur.depth + g);
		}
	}
	__syncthreads();
	int x, y;
	bool active = pixel_of_thread(p.band, x, y);
	unsigned haloMiss = 0;
	if (active) {
		const SceneView &sc = p.scene;
		size_t pix = local_index(p.band, x, y);
		f3 albedo = fetch_albedo(p.cur, sc.srgbLut, pix, nullptr);
		f3 normal = fetch_normal(p.cur, pix);
		float roughness, metallic;
		fetch_material(p.cur, pix, roughness, metallic);
		f3 worldPos = fetch_world_pos(p.cur, pix);
		float worldDepth = __ldg(p.cur.depth + pix);
		float albedoLum = luminance3(albedo.x, albedo.y, albedo.z);
		f3 cam = mk3(p.u.cameraPos[0], p.u.cameraPos[1], p.u.cameraPos[2]);
		Surface sf = make_surface(worldPos, normal, cam, roughness, metallic);
		float sinThr, cosThr;
		sincos_policy(p.u.spatialNormalThreshold * 0.017453292519943295f, sinThr, cosThr); // :67

		PackedReservoir res = load_reservoir(in, pix);
		Pcg32 rng = pcg_seed(p.u.frame * 31u + (uint32_t)iter, (uint32_t)y * 10007u + (uint32_t)x); // :47
		const uint32_t k = p.u.spatialNeighbors;
		for (uint32_t i = 0; i < k; ++i) {
			float angle = (pcg_float(rng) * 2.0f) * RESTIR_PI_F;                  // :52
			float radius = sqrtf(pcg_float(rng)) * p.u.spatialRadius;             // :53
			float sn, cs;
			sincos_policy(angle, sn, cs);
			int nx = x + (int)floorf(cs * radius), ny = y + (int)floorf(sn * radius); // :55-57
			nx = max(0, min(nx, p.band.W - 1));
			ny = max(0, min(ny, p.band.H - 1));
			if (ny < p.band.allocBegin || ny >= p.band.allocEnd) {
				haloMiss = 1;
				continue;
			}
			// the clamped neighbour is inside the staged rectangle: |offset| <= 30 and clamping moves it towards the pixel
			const int st = (ny - y0) * kStageW + (nx - x0);
			float nDepth = sDepth[st];
			short4 nq = sNormal[st];
			f3 nNor = mk3(fmaxf(div_snorm16((float)nq.x), -1.0f), fmaxf(div_snorm16((float)nq.y), -1.0f), fmaxf(div_snorm16((float)nq.z), -1.0f));
			if (fabsf(nDepth - worldDepth) > p.u.spatialPosThreshold * fabsf(worldDepth) || dot3(nNor, normal) < cosThr) { // :65-70
				continue;
			}
			PackedReservoir other = load_reservoir(in, local_index(p.band, nx, ny));
			combine_reservoirs(res, other, sc, sf, albedoLum, rng);               // :72-83
		}
		store_reservoir(out, pix, res);
		if (LIGHT) {
			shade_pixel(p, lu, pix, res, outPixels, outFormat);
		}
	}
	add_counter(p.counters, kCounterHaloMiss, haloMiss);
}

// ------------------------------------------------------------------------------------------------
// unbiasedReuse.glsl:50-124: merge the neighbours' reservoirs (no rejection), then decide which neighbours
// take part in the normalisation (:135-138, sample in front of the neighbour's surface).  Writes the merged
// reservoir and, per neighbour slot, the neighbour's local pixel index (< 0: contributes nothing and needs no
// ray).  The rays themselves (:139-166) are the trace kernel's, the normalisation unbiased_finalize_kernel's.
constexpr int kMaxUnbiasedNeighbors = 16;

// K > 0: the neighbour count is the compile-time K (the reference's NUM_NEIGHBORS 3 and the north-star's 5): the
// neighbour loops unroll and the neighbour indices stay in registers; K == 0: any count up to kMaxUnbiasedNeighbors.
#ifndef RESTIR_REUSE_MIN_BLOCKS
#define RESTIR_REUSE_MIN_BLOCKS 4 // 64 registers instead of 78: the kernel waits on gathers (long scoreboard), a fourth CTA per SM hides more of them (0.304 -> 0.272 ms)
#endif
template <int K>
__global__ void __launch_bounds__(kThreads, RESTIR_REUSE_MIN_BLOCKS) unbiased_merge_kernel(PassParams p, const PackedReservoir *__restrict__ in,
                                                                 PackedReservoir *__restrict__ out, int numNeighborsArg,
                                                                 int *__restrict__ neighborPix, uint32_t *__restrict__ neighborM) {
	const int numNeighbors = K ? K : numNeighborsArg;
	int x, y;
	bool active = pixel_of_thread(p.band, x, y);
	unsigned haloMiss = 0;
	if (active) {
		const SceneView &sc = p.scene;
		size_t pix = local_index(p.band, x, y);
		f3 albedo = fetch_albedo(p.cur, sc.srgbLut, pix, nullptr);
		f3 normal = fetch_normal(p.cur, pix);
		float roughness, metallic;
		fetch_material(p.cur, pix, roughness, metallic);
		f3 worldPos = fetch_world_pos(p.cur, pix);
		float albedoLum = luminance3(albedo.x, albedo.y, albedo.z);
		f3 cam = mk3(p.u.cameraPos[0], p.u.cameraPos[1], p.u.cameraPos[2]);
		Surface sf = make_surface(worldPos, normal, cam, roughness, metallic);

		PackedReservoir res = load_reservoir(in, pix);
		Pcg32 rng = pcg_seed(p.u.frame * 17u, (uint32_t)y * 10007u + (uint32_t)x); // :72
		int npx[K ? K : kMaxUnbiasedNeighbors];
		// the sample counts the normalisation adds up (:126, :139-156), handed to unbiased_finalize_kernel: slot j = neighbour j,
		// slot numNeighbors = this pixel before the merge
		uint32_t *ms = neighborM + tile_pixel_id() * (unsigned long long)(numNeighbors + 1);
		ms[numNeighbors] = res.M;
#pragma unroll(K ? K : 1)
		for (int i = 0; i < numNeighbors; ++i) {                                  // :84-124
			float angle = (pcg_float(rng) * 2.0f) * RESTIR_PI_F;
			float radius = sqrtf(pcg_float(rng)) * p.u.spatialRadius;
			float sn, cs;
			sincos_policy(angle, sn, cs);
			int nx = x + (int)roundf(cs * radius), ny = y + (int)roundf(sn * radius); // :88-89
			nx = max(0, min(nx, p.band.W - 1));
			ny = max(0, min(ny, p.band.H - 1));
			if (ny < p.band.allocBegin || ny >= p.band.allocEnd) {
				haloMiss = 1;
				npx[i] = -1;
				continue;
			}
			size_t npix = local_index(p.band, nx, ny);
			PackedReservoir other = load_reservoir(in, npix);
			npx[i] = (int)npix;
			ms[i] = other.M;
			res.M += other.M;                                                     // :104
			if (other.w != 0.0f && other.M != 0u) {                               // see combine_reservoirs
				f3 n; bool useN; float lum;
				sample_light_attrs(sc, other, n, useN, lum);
				float pHat = evaluate_phat(sf, albedoLum, mk3(other.px, other.py, other.pz), n, useN, lum);
				float weight = (pHat * other.w) * (float)other.M;
				if (weight > 0.0f) {
					update_reservoir(res, weight, mk3(other.px, other.py, other.pz), other.lightIndex, pHat, other.w, rng);
				}
			}
		}
		store_reservoir(out, pix, res);
		// :135-138
		f3 lightPos = mk3(res.px, res.py, res.pz);
		int *slots = neighborPix + tile_pixel_id() * (unsigned long long)numNeighbors;
#pragma unroll(K ? K : 1)
		for (int j = 0; j < numNeighbors; ++j) {
			int n = npx[j];
			if (n >= 0) {
				f3 nPos = fetch_world_pos(p.cur, (size_t)n);
				f3 nNor = fetch_normal(p.cur, (size_t)n);
				if (dot3(lightPos - nPos, nNor) < 0.0f) {
					n = -1;
				}
			}
			slots[j] = n;
		}
	}
	add_counter(p.counters, kCounterHaloMiss, haloMiss);
}

// unbiasedReuse.glsl:126-182 given the visibility bytes: Z = own M + the M of every participating, unshadowed
// neighbour (the counts unbiased_merge_kernel handed over: no gather here); everything is dropped when the pixel
// itself is shadowed.  LIGHT: see spatial_reuse_kernel.
template <bool LIGHT>
__global__ void __launch_bounds__(kThreads) unbiased_finalize_kernel(PassParams p, PackedReservoir *__restrict__ out, int numNeighbors,
                                                                    const int *__restrict__ neighborPix, const uint32_t *__restrict__ neighborM,
                                                                    const unsigned char *__restrict__ shadowed, restir_lighting_uniforms lu,
                                                                    void *__restrict__ outPixels, int outFormat) {
	int x, y;
	if (!pixel_of_thread(p.band, x, y)) {
		return;
	}
	size_t pix = local_index(p.band, x, y);
	const bool vis = (p.u.flags & RESTIR_VISIBILITY_REUSE_FLAG) != 0;
	unsigned long long id = tile_pixel_id();
	const int *slots = neighborPix + id * (unsigned long long)numNeighbors;
	const uint32_t *ms = neighborM + id * (unsigned long long)(numNeighbors + 1);
	const unsigned char *bits = shadowed + id * (unsigned long long)(numNeighbors + 1);
	float4 *o = reinterpret_cast<float4 *>(out + pix);
	float4 b = o[1]; // pHat, sumWeights, w, M of the merged reservoir
	uint32_t numSamples = ms[numNeighbors]; // own M before the merge, :126
#pragma unroll 1
	for (int j = 0; j < numNeighbors; ++j) {
		if (slots[j] < 0 || (vis && bits[j] != 0)) {
			continue;
		}
		numSamples += ms[j];
	}
	if (vis && bits[numNeighbors] != 0) {                                         // :157-166
		numSamples = 0;
	}
	if (numSamples > 0) {                                                         // :171-181
		b.z = b.y / ((float)numSamples * b.x);
	} else {
		b.z = 0.0f;
		b.y = 0.0f;
	}
	o[1] = b;
	if (LIGHT) {
		float4 a = o[0];
		PackedReservoir r;
		r.px = a.x; r.py = a.y; r.pz = a.z; r.lightIndex = __float_as_int(a.w);
		r.pHat = b.x; r.sumWeights = b.y; r.w = b.z; r.M = __float_as_uint(b.w);
		shade_pixel(p, lu, pix, r, outPixels, outFormat);
	}
}

// ------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------
// lighting.frag:43-71,103 (debugMode 0)

// lighting.frag:43-71,103 for one pixel whose final reservoir is `r`
__device__ __forceinline__ void shade_pixel(const PassParams &p, const restir_lighting_uniforms &lu, size_t pix, const PackedReservoir &r,
                                            void *__restrict__ outPixels, int outFormat) {
	const SceneView &sc = p.scene;
	float albedoA;
	f3 albedo = fetch_albedo(p.cur, sc.srgbLut, pix, &albedoA);
	f3 normal = fetch_normal(p.cur, pix);
	float roughness, metallic;
	fetch_material(p.cur, pix, roughness, metallic);
	f3 worldPos = fetch_world_pos(p.cur, pix);
	f3 cam = mk3(lu.cameraPos[0], lu.cameraPos[1], lu.cameraPos[2]);
	Surface sf = make_surface(worldPos, normal, cam, roughness, metallic);

	f3 emission = mk3(0.0f, 0.0f, 0.0f);                                           // :55-60 (out-of-range reads give 0)
	if (r.lightIndex < 0) {
		int ti = -1 - r.lightIndex;
		if (ti < sc.triCount) {
			float4 e = __ldg(reinterpret_cast<const float4 *>(sc.triLights + ti) + 3);
			emission = mk3(e.x, e.y, e.z);
		}
	} else if (r.lightIndex < sc.pointCount) {
		float4 e = __ldg(reinterpret_cast<const float4 *>(sc.pointLights + r.lightIndex) + 1);
		emission = mk3(e.x, e.y, e.z);
	}
	f3 n; bool useN; float lum;
	sample_light_attrs(sc, r, n, useN, lum);
	f3 c = evaluate_phat_full(sf, albedo, mk3(r.px, r.py, r.pz), n, useN, emission) * r.w; // :61-66
	if (albedoA > 0.5f) {                                                          // :69-71
		c = albedo;
	}
	if (lu.gamma != 1.0f) {                                                        // :103, P11
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

__global__ void __launch_bounds__(kThreads) lighting_kernel(PassParams p, restir_lighting_uniforms lu,
                                                           const PackedReservoir *__restrict__ reservoirs, void *__restrict__ outPixels,
                                                           int outFormat) {
	int x, y;
	if (!pixel_of_thread(p.band, x, y)) {
		return;
	}
	size_t pix = local_index(p.band, x, y);
	shade_pixel(p, lu, pix, load_reservoir(reservoirs, pix), outPixels, outFormat);
}

// ------------------------------------------------------------------------------------------------
// fixture tool: primary-visibility G-buffer (closest front-facing hit), twin of oracle_raycast_gbuffer
__global__ void __launch_bounds__(kThreads) raycast_gbuffer_kernel(SceneView sc, Band band, RaycastCamera cam,
                                                                  const int *__restrict__ triMaterial, const uint4 *__restrict__ materialTable,
                                                                  uchar4 *albedo, short4 *normal, ushort2 *material, float4 *worldPos, float *depth) {
	int x, y;
	if (!pixel_of_thread(band, x, y)) {
		return;
	}
	size_t pix = local_index(band, x, y);
	float ndcx = (((float)x + 0.5f) / (float)band.W) * 2.0f - 1.0f;
	float ndcy = (((float)y + 0.5f) / (float)band.H) * 2.0f - 1.0f;
	f3 pos = mk3(cam.pos[0], cam.pos[1], cam.pos[2]);
	f3 fwd = mk3(cam.fwd[0], cam.fwd[1], cam.fwd[2]);
	f3 right = mk3(cam.right[0], cam.right[1], cam.right[2]);
	f3 up = mk3(cam.up[0], cam.up[1], cam.up[2]);
	f3 dir = (fwd + right * (ndcx * cam.sx)) - up * (ndcy * cam.sy);
	f3 inv = mk3(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z);
	float best = __int_as_float(0x7f800000), bu = 0.0f, bv = 0.0f;
	int bestTri = -1;
	int stack[64];
	int top = 1;
	stack[0] = 0;
	while (top > 0) {
		const float4 *n = sc.nodes + (size_t)stack[--top] * 5;
		float4 ch = __ldg(n + 4);
#pragma unroll
		for (int side = 0; side < 2; ++side) {
			float4 bmin = __ldg(n + side * 2), bmax = __ldg(n + side * 2 + 1);
			int child = __float_as_int(side ? ch.y : ch.x);
			float t1x = (bmin.x - pos.x) * inv.x, t1y = (bmin.y - pos.y) * inv.y, t1z = (bmin.z - pos.z) * inv.z;
			float t2x = (bmax.x - pos.x) * inv.x, t2y = (bmax.y - pos.y) * inv.y, t2z = (bmax.z - pos.z) * inv.z;
			float rmin = fmaxf(fminf(t1x, t2x), fmaxf(fminf(t1y, t2y), fminf(t1z, t2z)));
			float rmax = fminf(fmaxf(t1x, t2x), fminf(fmaxf(t1y, t2y), fmaxf(t1z, t2z)));
			if (!(rmin <= best && rmax >= rmin && rmax > 0.0f)) {
				continue;
			}
			if (child >= 0) {
				if (top < 64) {
					stack[top++] = child;
				}
				continue;
			}
			int ti = ~child;
			const float4 *t = sc.tris + (size_t)ti * 3;
			float4 a = __ldg(t), b = __ldg(t + 1), c = __ldg(t + 2);
			f3 p1 = mk3(a.x, a.y, a.z);
			f3 e1 = mk3(b.x, b.y, b.z) - p1;
			f3 e2 = mk3(c.x, c.y, c.z) - p1;
			f3 nn = cross3(e1, e2);
			if (!(dot3(nn, dir) < 0.0f)) {
				continue;
			}
			if (__ldg(materialTable + __ldg(triMaterial + ti)).z & 1u) {
				continue;
			}
			f3 pv = cross3(dir, e2);
			float fdet = 1.0f / dot3(e1, pv);
			f3 sv = pos - p1;
			float u_ = fdet * dot3(sv, pv);
			if (u_ < 0.0f || u_ > 1.0f) {
				continue;
			}
			f3 q = cross3(sv, e1);
			float v_ = fdet * dot3(dir, q);
			if (v_ < 0.0f || v_ + u_ > 1.0f) {
				continue;
			}
			float tt = fdet * dot3(e2, q);
			if (tt > 0.0f && (tt < best || (tt == best && ti < bestTri))) {
				best = tt;
				bestTri = ti;
				bu = u_;
				bv = v_;
			}
		}
	}
	if (bestTri < 0) {
		albedo[pix] = make_uchar4(0, 0, 0, 255);
		normal[pix] = make_short4(0, 0, 0, 32767);
		material[pix] = make_ushort2(0, 0);
		worldPos[pix] = make_float4(0.0f, 0.0f, 0.0f, 1.0f);
		depth[pix] = 1.0f;
		return;
	}
	const float4 *t = sc.tris + (size_t)bestTri * 3;
	float4 a = __ldg(t), b = __ldg(t + 1), c = __ldg(t + 2);
	f3 p1 = mk3(a.x, a.y, a.z), p2 = mk3(b.x, b.y, b.z), p3 = mk3(c.x, c.y, c.z);
	f3 nn = normalize3(cross3(p2 - p1, p3 - p1));
	f3 hit = (p1 * ((1.0f - bu) - bv) + p2 * bu) + p3 * bv;
	uint4 mt = __ldg(materialTable + __ldg(triMaterial + bestTri));
	albedo[pix] = make_uchar4(mt.x & 255u, (mt.x >> 8) & 255u, (mt.x >> 16) & 255u, mt.x >> 24);
	material[pix] = make_ushort2(mt.y & 65535u, mt.y >> 16);
	normal[pix] = make_short4((short)rintf(fminf(fmaxf(nn.x, -1.0f), 1.0f) * 32767.0f), (short)rintf(fminf(fmaxf(nn.y, -1.0f), 1.0f) * 32767.0f),
	                          (short)rintf(fminf(fmaxf(nn.z, -1.0f), 1.0f) * 32767.0f), 32767);
	worldPos[pix] = make_float4(hit.x, hit.y, hit.z, 1.0f);
	const float *PV = cam.pv;
	float cz = ((PV[2] * hit.x + PV[6] * hit.y) + PV[10] * hit.z) + PV[14];
	float cw = ((PV[3] * hit.x + PV[7] * hit.y) + PV[11] * hit.z) + PV[15];
	depth[pix] = cz / cw;
}

// ------------------------------------------------------------------------------------------------
// 64-byte reference layout <-> 32-byte packed layout (boundary conversions for download / upload)
__global__ void unpack_reservoirs_kernel(SceneView sc, const PackedReservoir *__restrict__ in, restir_reservoir *__restrict__ out, size_t n) {
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) {
		return;
	}
	PackedReservoir r = load_reservoir(in, i);
	f3 nrm; bool useN; float lum;
	sample_light_attrs(sc, r, nrm, useN, lum);
	float4 *o = reinterpret_cast<float4 *>(out + i);
	o[0] = make_float4(r.px, r.py, r.pz, lum);
	o[1] = make_float4(nrm.x, nrm.y, nrm.z, useN ? 1.0f : 0.0f);
	o[2] = make_float4(__int_as_float(r.lightIndex), r.pHat, r.sumWeights, r.w);
	o[3] = make_float4(__uint_as_float(r.M), 0.0f, 0.0f, 0.0f);
}
__global__ void pack_reservoirs_kernel(const restir_reservoir *__restrict__ in, PackedReservoir *__restrict__ out, size_t n) {
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) {
		return;
	}
	const float4 *s = reinterpret_cast<const float4 *>(in + i);
	float4 a = s[0], c = s[2], d = s[3];
	PackedReservoir r;
	r.px = a.x; r.py = a.y; r.pz = a.z;
	r.lightIndex = __float_as_int(c.x);
	r.pHat = c.y; r.sumWeights = c.z; r.w = c.w;
	r.M = __float_as_uint(d.x);
	store_reservoir(out, i, r);
}

// derived light tables (upload time)
__global__ void derive_point_table_kernel(const restir_point_light *__restrict__ lights, int n, float4 *__restrict__ out) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) {
		out[i] = make_float4(lights[i].pos[0], lights[i].pos[1], lights[i].pos[2], lights[i].color_luminance[3]);
	}
}
__global__ void derive_tri_table_kernel(const restir_tri_light *__restrict__ lights, int n, float4 *__restrict__ out) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) {
		out[i] = make_float4(lights[i].normalArea[0], lights[i].normalArea[1], lights[i].normalArea[2], lights[i].emission_luminance[3]);
	}
}

// (p1, e1 = p2 - p1, e2 = p3 - p1) records for the trace kernel (restir_trace.cuh ray_triangle_edges): the subtractions of
// softwareRaytracing.glsl:16-17, made once per upload instead of once per test
// order (may be null: identity): record i holds triangle order[i] — the wide image names its leaves' records consecutively
// (wide_image.h).  leafBoxes (may be null): per record the fp32 box of its leaf (min.xyz, max.xyz), kept in the spare floats
// of the record for the wide walk's exact leaf test (restir_wide.cuh wide_leaf_hit).
__global__ void derive_triangle_edges_kernel(const float4 *__restrict__ tris, uint32_t n, float4 *__restrict__ out, const uint32_t *__restrict__ order,
                                             const float *__restrict__ leafBoxes) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) {
		const size_t t = order ? order[i] : i;
		float4 a = tris[t * 3], b = tris[t * 3 + 1], c = tris[t * 3 + 2];
		f3 p1 = mk3(a.x, a.y, a.z);
		f3 e1 = mk3(b.x, b.y, b.z) - p1;
		f3 e2 = mk3(c.x, c.y, c.z) - p1;
		float4 *o = out + (size_t)i * 4;
		o[0] = make_float4(p1.x, p1.y, p1.z, e1.x);
		o[1] = make_float4(e1.y, e1.z, e2.x, e2.y);
		const float *lb = leafBoxes ? leafBoxes + (size_t)i * 6 : nullptr;
		o[2] = lb ? make_float4(e2.z, lb[0], lb[1], lb[2]) : make_float4(e2.z, 0.0f, 0.0f, 0.0f);
		o[3] = lb ? make_float4(lb[3], lb[4], lb[5], 0.0f) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
	}
}

// ------------------------------------------------------------------------------------------------
// launchers
static dim3 tile_grid(const Band &b) {
	return dim3((unsigned)((b.W + kTileW - 1) / kTileW), (unsigned)((b.rowEnd - b.rowBegin + kTileH - 1) / kTileH), 1);
}

PassGrid pass_grid(const Band &b) {
	dim3 g = tile_grid(b);
	return PassGrid{g.x, g.y, g.x * 4u, 256ull * g.x * g.y};
}

void launch_omni_candidates(const PassParams &p, PackedReservoir *out, cudaStream_t s) {
	omni_candidates_kernel<<<tile_grid(p.band), kThreads, 0, s>>>(p, out);
}
void launch_omni_temporal(const PassParams &p, PackedReservoir *out, const PackedReservoir *prev, const unsigned char *shadowed, cudaStream_t s) {
	LcgJump jump = lcg_jump((uint64_t)p.u.initialLightSampleCount * draws_per_candidate(p.scene.pointCount != 0));
	omni_temporal_kernel<<<tile_grid(p.band), kThreads, 0, s>>>(p, out, prev, shadowed, jump);
}
void launch_spatial_reuse(const PassParams &p, const PackedReservoir *in, PackedReservoir *out, int iter, const restir_lighting_uniforms *lu,
                          void *outPixels, int fmt, bool staged, cudaStream_t s) {
	// the staged apron is sized for the reference's radius (30) and G-buffer planes that exist
	if (staged && p.u.spatialRadius <= 30.0f && p.u.spatialRadius >= 0.0f && p.cur.normal != nullptr) {
		static bool configured = false;
		if (!configured) {
			cudaFuncSetAttribute(spatial_reuse_staged_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStageBytes);
			cudaFuncSetAttribute(spatial_reuse_staged_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStageBytes);
			configured = true;
		}
		if (lu) {
			spatial_reuse_staged_kernel<true><<<tile_grid(p.band), kThreads, kStageBytes, s>>>(p, in, out, iter, *lu, outPixels, fmt);
		} else {
			spatial_reuse_staged_kernel<false><<<tile_grid(p.band), kThreads, kStageBytes, s>>>(p, in, out, iter, restir_lighting_uniforms{}, nullptr, 0);
		}
		return;
	}
	if (lu) {
		spatial_reuse_kernel<true><<<tile_grid(p.band), kThreads, 0, s>>>(p, in, out, iter, *lu, outPixels, fmt);
	} else {
		spatial_reuse_kernel<false><<<tile_grid(p.band), kThreads, 0, s>>>(p, in, out, iter, restir_lighting_uniforms{}, nullptr, 0);
	}
}
void launch_unbiased_merge(const PassParams &p, const PackedReservoir *in, PackedReservoir *out, int numNeighbors, int *neighborPix,
                           uint32_t *neighborM, cudaStream_t s) {
	switch (numNeighbors) {
	case 3: unbiased_merge_kernel<3><<<tile_grid(p.band), kThreads, 0, s>>>(p, in, out, numNeighbors, neighborPix, neighborM); break;
	case 5: unbiased_merge_kernel<5><<<tile_grid(p.band), kThreads, 0, s>>>(p, in, out, numNeighbors, neighborPix, neighborM); break;
	default: unbiased_merge_kernel<0><<<tile_grid(p.band), kThreads, 0, s>>>(p, in, out, numNeighbors, neighborPix, neighborM); break;
	}
}
void launch_unbiased_finalize(const PassParams &p, PackedReservoir *out, int numNeighbors, const int *neighborPix, const uint32_t *neighborM,
                              const unsigned char *shadowed, const restir_lighting_uniforms *lu, void *outPixels, int fmt, cudaStream_t s) {
	if (lu) {
		unbiased_finalize_kernel<true><<<tile_grid(p.band), kThreads, 0, s>>>(p, out, numNeighbors, neighborPix, neighborM, shadowed, *lu, outPixels, fmt);
	} else {
		unbiased_finalize_kernel<false><<<tile_grid(p.band), kThreads, 0, s>>>(p, out, numNeighbors, neighborPix, neighborM, shadowed,
		                                                                      restir_lighting_uniforms{}, nullptr, 0);
	}
}
void launch_lighting(const PassParams &p, const restir_lighting_uniforms &lu, const PackedReservoir *res, void *out, int fmt, cudaStream_t s) {
	lighting_kernel<<<tile_grid(p.band), kThreads, 0, s>>>(p, lu, res, out, fmt);
}
void launch_raycast_gbuffer(const SceneView &sc, const Band &band, const RaycastCamera &cam, const int *triMaterial, const uint4 *materialTable,
                            void *albedo, void *normal, void *material, void *worldPos, void *depth, cudaStream_t s) {
	raycast_gbuffer_kernel<<<tile_grid(band), kThreads, 0, s>>>(sc, band, cam, triMaterial, materialTable, (uchar4 *)albedo, (short4 *)normal,
	                                                           (ushort2 *)material, (float4 *)worldPos, (float *)depth);
}
void launch_unpack_reservoirs(const SceneView &sc, const PackedReservoir *in, restir_reservoir *out, size_t n, cudaStream_t s) {
	if (n) unpack_reservoirs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(sc, in, out, n);
}
void launch_pack_reservoirs(const restir_reservoir *in, PackedReservoir *out, size_t n, cudaStream_t s) {
	if (n) pack_reservoirs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(in, out, n);
}
void launch_derive_triangle_edges(const float4 *tris, uint32_t n, float4 *out, const uint32_t *order, const float *leafBoxes, cudaStream_t s) {
	if (n) derive_triangle_edges_kernel<<<(n + 255) / 256, 256, 0, s>>>(tris, n, out, order, leafBoxes);
}
void launch_derive_light_tables(const restir_point_light *pl, int np, float4 *pointOut, const restir_tri_light *tl, int nt, float4 *triOut,
                                cudaStream_t s) {
	if (np) derive_point_table_kernel<<<(np + 255) / 256, 256, 0, s>>>(pl, np, pointOut);
	if (nt) derive_tri_table_kernel<<<(nt + 255) / 256, 256, 0, s>>>(tl, nt, triOut);
}

// With lazy module loading (the CUDA 12 default) the first launch of a kernel loads it, which can synchronise the device:
// fatal timing for a band whose stream is spinning in halo_wait_kernel for a neighbour driven by the same host thread.
// restir_create touches every kernel once instead.
cudaError_t preload_pixel_kernels() {
	cudaFuncAttributes a;
	cudaError_t e = cudaSuccess;
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, omni_candidates_kernel);
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, omni_temporal_kernel);
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, spatial_reuse_kernel<false>);
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, spatial_reuse_kernel<true>);
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, spatial_reuse_staged_kernel<false>);
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, spatial_reuse_staged_kernel<true>);
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, unbiased_merge_kernel<0>);
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, unbiased_merge_kernel<3>);
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, unbiased_merge_kernel<5>);
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, unbiased_finalize_kernel<false>);
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, unbiased_finalize_kernel<true>);
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, lighting_kernel);
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, raycast_gbuffer_kernel);
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, unpack_reservoirs_kernel);
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, pack_reservoirs_kernel);
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, derive_point_table_kernel);
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, derive_tri_table_kernel);
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, derive_triangle_edges_kernel);
	return e;
}

} // namespace restir
