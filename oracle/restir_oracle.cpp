// restir_oracle.cpp — CPU restatement of the reference's ReSTIR hot path.
//
// *** TEST INFRASTRUCTURE, NOT PRODUCT. ***  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library.  The product (restir-vulkan_b200/) never
// links, imports or calls it and has no CPU fallback.
//
// What it follows (all paths relative to /root/reference):
//   src/shaders/restirOmni.glsl, spatialReuse.comp, unbiasedReuse.glsl, lighting.frag (debugMode 0)
//   src/shaders/include/{rand,reservoir,restirUtils,disneyBRDF,common,softwareRaytracing,visibilityTest}.glsl
//   src/shaders/include/structs/*.glsl (layouts: include/restir_layouts.h)
// Each function cites the lines it restates.
//
// PARITY STATUS: PINNED BY THE REFERENCE'S OWN SOURCE.  The reference has no tests, golden vectors or dump
// facility for this path and no Vulkan/GLSL toolchain exists in this image, so the pin is
// built here: oracle/ref_build compiles the reference's shader sources themselves (restirOmni.glsl,
// spatialReuse.comp, unbiasedReuse.glsl, lighting.frag and their includes, from where they lie) as C++ into
// oracle/_ref/libglslref.so, and tests/test_oracle_vs_glsl.py demands that this file reproduces that build bit
// for bit — every reservoir field of every pixel over multi-frame sequences of cornellBox, Sponza, office and
// procedural scenes, every visibility bit, every output colour.  Frame sequences made by that build are committed
// as tests/golden/frames_*.npz (tests/make_golden_frames.py).  Also pinned: PCG32 (canonical pcg32 demo
// stream), struct layouts, and — via oracle/_ref/scene_baker, which runs the reference's real C++ — the BVH /
// light / alias-table inputs.  What stays a stated choice rather than a measured fact is the precision of the
// GLSL built-ins a driver is free to pick (below); the shader build and this file share that one reading.
//
// ARITHMETIC POLICY (DESIGN.md §Arithmetic policy).  GLSL leaves the precision of '/', sqrt, pow,
// sin, cos, normalize and FMA contraction to the driver.  This oracle fixes one IEEE-754 binary32
// reading of every such site, written out operation by operation, and the CUDA kernels implement
// the same reading independently; discrete decisions (reservoir replacement, neighbour choice, hit /
// miss) therefore agree bit for bit.  Build with -ffp-contract=off and no -ffast-math.
//   P1  + - * / sqrt are IEEE round-to-nearest-even, never contracted into FMA.
//   P2  normalize(v)    = v * (1.0f / sqrt(dot(v,v)))
//   P3  vec / scalar    = vec * (1.0f / scalar)          (wi /= sqrt(d); p.xyz /= p.w; (box - o) / dir)
//   P4  dot(a,b)        = (a.x*b.x + a.y*b.y) + a.z*b.z ; mat4*vec4 = ((c0*x + c1*y) + c2*z) + c3*w
//   P5  pow(x, 2.0)     = x * x
//   P6  mix(x,y,a)      = x*(1-a) + y*a ;  clamp = min(max(x,lo),hi) ; min/max = IEEE fminf/fmaxf
//   P7  sin/cos         = det_sincos below (Cody-Waite by pi/2 + Cephes single-precision polynomials)
//   P8  radians(d)      = d * 0.017453292519943295f ; M_PI = 3.14159274f
//   P9  floor/round/int = floorf / roundf (half away from zero) / truncation
//   P10 uint -> float   = round-to-nearest-even; randFloat = float(u) * 2^-32 (can be exactly 1.0)
//   P11 pow(c, 1/gamma) = identity when gamma == 1, else powf (lighting output only: toleranced)
//   P12 sRGB8 decode    = 256-entry table from the sRGB EOTF evaluated in double, rounded to float

#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#	include <omp.h>
#endif

#include "../include/restir_layouts.h"

namespace {

struct V3 {
	float x, y, z;
};
inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
inline V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; } // P4
inline V3 cross(V3 a, V3 b) {
	return V3{a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; // GLSL spec form
}
inline V3 normalize(V3 v) { // P2
	float inv = 1.0f / sqrtf(dot(v, v));
	return v * inv;
}
inline float mixf(float x, float y, float a) { return x * (1.0f - a) + y * a; } // P6
inline float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

const float kPi = 3.14159274f; // disneyBRDF.glsl:1 M_PI as binary32

// P7: deterministic sin/cos.  k = rint(a*2/pi); r = a - k*pi/2 in three steps; Cephes sinf/cosf kernels.
void det_sincos(float a, float *s, float *c) {
	float kf = rintf(a * 0.636619772f);
	int k = (int)kf;
	float r = a - kf * 1.5703125f;
	r = r - kf * 4.837512969970703125e-4f;
	r = r - kf * 7.54978995489188216e-8f;
	float z = r * r;
	float ps = -1.9515295891e-4f * z;
	ps = ps + 8.3321608736e-3f;
	ps = ps * z;
	ps = ps - 1.6666654611e-1f;
	ps = ps * z;
	ps = ps * r;
	float sr = ps + r;
	float pc = 2.443315711809948e-5f * z;
	pc = pc - 1.388731625493765e-3f;
	pc = pc * z;
	pc = pc + 4.166664568298827e-2f;
	pc = pc * z;
	pc = pc * z;
	float hz = 0.5f * z;
	pc = pc - hz;
	float cr = pc + 1.0f;
	switch (k & 3) {
	case 0: *s = sr; *c = cr; break;
	case 1: *s = cr; *c = -sr; break;
	case 2: *s = -sr; *c = -cr; break;
	default: *s = -cr; *c = sr; break;
	}
}

// ---------------------------------------------------------------------------------------------
// rand.glsl:7-32 — PCG32 (XSH-RR 64/32)
struct Rand {
	uint64_t state, inc;
};
uint32_t randUint(Rand &r) { // rand.glsl:12-18
	uint64_t oldState = r.state;
	r.state = oldState * 6364136223846793005ull + r.inc;
	uint32_t xorShifted = (uint32_t)(((oldState >> 18u) ^ oldState) >> 27u);
	uint32_t rot = (uint32_t)(oldState >> 59u);
	return (xorShifted >> rot) | (xorShifted << ((0u - rot) & 31u));
}
Rand seedRand(uint64_t seed, uint64_t seq) { // rand.glsl:20-28
	Rand result;
	result.state = 0;
	result.inc = (seq << 1u) | 1u;
	randUint(result);
	result.state += seed;
	randUint(result);
	return result;
}
float randFloat(Rand &r) { // rand.glsl:30-32, P10
	return (float)randUint(r) * 2.3283064365386963e-10f;
}

// ---------------------------------------------------------------------------------------------
// common.glsl:7-9
float luminance(float r, float g, float b) { return (0.2126f * r + 0.7152f * g) + 0.0722f * b; }

// disneyBRDF.glsl:5-10
float schlickFresnel(float c) {
	float m = clampf(1.0f - c, 0.0f, 1.0f);
	float sm = m * m;
	return (sm * sm) * m;
}
// disneyBRDF.glsl:13-18
float GTR2(float NdotH, float a) {
	float a2 = a * a;
	float t = 1.0f + ((a2 - 1.0f) * NdotH) * NdotH;
	return a2 / ((kPi * t) * t);
}
// disneyBRDF.glsl:20-25 (receives `a`, squares it again: SURVEY Appendix B.11)
float smithG_GGX(float NdotV, float alphaG) {
	float a = alphaG * alphaG;
	float b = NdotV * NdotV;
	return 1.0f / (fabsf(NdotV) + fmaxf(sqrtf((a + b) - a * b), 0.0001f));
}
// disneyBRDF.glsl:27-33
float disneyBrdfDiffuseFactor(float cosIn, float cosOut, float cosInHalf, float roughness, float metallic) {
	float fresnelIn = schlickFresnel(cosIn);
	float fresnelOut = schlickFresnel(cosOut);
	float fresnelDiffuse90 = 0.5f + ((2.0f * cosInHalf) * cosInHalf) * roughness;
	float fresnelDiffuse = mixf(1.0f, fresnelDiffuse90, fresnelIn) * mixf(1.0f, fresnelDiffuse90, fresnelOut);
	return (fresnelDiffuse * (1.0f - metallic)) / kPi;
}
// disneyBRDF.glsl:42-55; returns (fresnelInHalf, Gs*Ds)
void disneyBrdfSpecularFactors(float cosIn, float cosOut, float cosHalf, float cosInHalf, float roughness,
                               float *fresnelInHalf, float *gsds) {
	*fresnelInHalf = schlickFresnel(cosInHalf);
	float a = fmaxf(0.001f, roughness * roughness); // P5
	float Ds = GTR2(cosHalf, a);
	float Gs = smithG_GGX(cosIn, a);
	Gs = Gs * smithG_GGX(cosOut, a);
	*gsds = Gs * Ds;
}
// disneyBRDF.glsl:82-90 with :37-39 and :64-71
float disneyBrdfLuminance(float cosIn, float cosOut, float cosHalf, float cosInHalf, float albedoLum,
                          float roughness, float metallic) {
	if (cosIn < 0.0f) {
		return 0.0f;
	}
	float diffuse = albedoLum * disneyBrdfDiffuseFactor(cosIn, cosOut, cosInHalf, roughness, metallic);
	float f, gd;
	disneyBrdfSpecularFactors(cosIn, cosOut, cosHalf, cosInHalf, roughness, &f, &gd);
	float specularLuminance = mixf(0.04f, albedoLum, metallic);
	float Fs = mixf(specularLuminance, 1.0f, f);
	float specular = Fs * gd;
	return diffuse + specular;
}
// disneyBRDF.glsl:73-81 with :34-36 and :56-63
V3 disneyBrdfColor(float cosIn, float cosOut, float cosHalf, float cosInHalf, V3 albedo, float roughness,
                   float metallic) {
	if (cosIn < 0.0f) {
		return v3(0, 0, 0);
	}
	V3 diffuse = albedo * disneyBrdfDiffuseFactor(cosIn, cosOut, cosInHalf, roughness, metallic);
	float f, gd;
	disneyBrdfSpecularFactors(cosIn, cosOut, cosHalf, cosInHalf, roughness, &f, &gd);
	V3 specularColor = v3(mixf(0.04f, albedo.x, metallic), mixf(0.04f, albedo.y, metallic), mixf(0.04f, albedo.z, metallic));
	V3 Fs = v3(mixf(specularColor.x, 1.0f, f), mixf(specularColor.y, 1.0f, f), mixf(specularColor.z, 1.0f, f));
	return diffuse + Fs * gd;
}

// shared front half of restirUtils.glsl:3-28 and :30-55
struct PHatGeom {
	bool behind;
	float cosIn, cosOut, cosHalf, cosInHalf, geometry;
};
PHatGeom pHatGeometry(V3 worldPos, V3 lightPos, V3 camPos, V3 normal, V3 lightNormal, bool useLightNormal) {
	PHatGeom g{};
	V3 wi = lightPos - worldPos;
	if (dot(wi, normal) < 0.0f) {
		g.behind = true;
		return g;
	}
	float sqrDist = dot(wi, wi);
	wi = wi * (1.0f / sqrtf(sqrDist)); // P3
	V3 wo = normalize(camPos - worldPos);
	g.cosIn = dot(normal, wi);
	g.cosOut = dot(normal, wo);
	V3 halfVec = normalize(wi + wo);
	g.cosHalf = dot(normal, halfVec);
	g.cosInHalf = dot(wi, halfVec);
	g.geometry = g.cosIn / sqrDist;
	if (useLightNormal) {
		g.geometry = g.geometry * fabsf(dot(wi, lightNormal));
	}
	return g;
}
// restirUtils.glsl:3-28
float evaluatePHat(V3 worldPos, V3 lightPos, V3 camPos, V3 normal, V3 lightNormal, bool useLightNormal,
                   float albedoLum, float emissionLum, float roughness, float metallic) {
	PHatGeom g = pHatGeometry(worldPos, lightPos, camPos, normal, lightNormal, useLightNormal);
	if (g.behind) {
		return 0.0f;
	}
	return (emissionLum * disneyBrdfLuminance(g.cosIn, g.cosOut, g.cosHalf, g.cosInHalf, albedoLum, roughness, metallic)) * g.geometry;
}
// restirUtils.glsl:30-55
V3 evaluatePHatFull(V3 worldPos, V3 lightPos, V3 camPos, V3 normal, V3 lightNormal, bool useLightNormal,
                    V3 albedo, V3 emission, float roughness, float metallic) {
	PHatGeom g = pHatGeometry(worldPos, lightPos, camPos, normal, lightNormal, useLightNormal);
	if (g.behind) {
		return v3(0, 0, 0);
	}
	return (emission * disneyBrdfColor(g.cosIn, g.cosOut, g.cosHalf, g.cosInHalf, albedo, roughness, metallic)) * g.geometry;
}

// ---------------------------------------------------------------------------------------------
// reservoir.glsl (RESERVOIR_SIZE 1)
typedef restir_reservoir Reservoir;

inline V3 samplePos(const Reservoir &r) {
	return v3(r.sample.position_emissionLum[0], r.sample.position_emissionLum[1], r.sample.position_emissionLum[2]);
}
inline V3 sampleNormal(const Reservoir &r) { return v3(r.sample.normal[0], r.sample.normal[1], r.sample.normal[2]); }

// reservoir.glsl:6-26.  One RNG draw, always (even for weight 0: 0/0 = NaN, the compare is false).
void updateReservoirAt(Reservoir &res, float weight, V3 position, const float normal[4], float emissionLum,
                       int lightIdx, float pHat, float w, Rand &rand) {
	res.sample.sumWeights = res.sample.sumWeights + weight;
	float replacePossibility = weight / res.sample.sumWeights;
	if (randFloat(rand) < replacePossibility) {
		res.sample.position_emissionLum[0] = position.x;
		res.sample.position_emissionLum[1] = position.y;
		res.sample.position_emissionLum[2] = position.z;
		res.sample.position_emissionLum[3] = emissionLum;
		std::memcpy(res.sample.normal, normal, 16);
		res.sample.lightIndex = lightIdx;
		res.sample.pHat = pHat;
		res.sample.w = w;
	}
}
// reservoir.glsl:28-42.  `w` is computed at insertion time and goes stale.
void addSampleToReservoir(Reservoir &res, V3 position, const float normal[4], float emissionLum, int lightIdx,
                          float pHat, float sampleP, Rand &rand) {
	float weight = pHat / sampleP;
	res.numStreamSamples += 1;
	float w = (res.sample.sumWeights + weight) / ((float)res.numStreamSamples * pHat);
	updateReservoirAt(res, weight, position, normal, emissionLum, lightIdx, pHat, w, rand);
}
// reservoir.glsl:44-64
void combineReservoirs(Reservoir &self, const Reservoir &other, float pHat, Rand &rand) {
	self.numStreamSamples += other.numStreamSamples;
	float weight = (pHat * other.sample.w) * (float)other.numStreamSamples;
	if (weight > 0.0f) {
		updateReservoirAt(self, weight, samplePos(other), other.sample.normal, other.sample.position_emissionLum[3],
		                  other.sample.lightIndex, pHat, other.sample.w, rand);
	}
	if (self.sample.w > 0.0f) {
		self.sample.w = self.sample.sumWeights / ((float)self.numStreamSamples * self.sample.pHat);
	}
}
// reservoir.glsl:66-76.  Only sumWeights and M are set in GLSL; the oracle DEFINES the rest as zero
// (SURVEY Appendix B.2).
Reservoir newReservoir() {
	Reservoir r;
	std::memset(&r, 0, sizeof(r));
	return r;
}

// ---------------------------------------------------------------------------------------------
// softwareRaytracing.glsl + visibilityTest.glsl (software branch)
struct Scene {
	const restir_aabb_node *nodes;
	const restir_triangle *tris;
	const restir_point_light *pointLights;
	int pointCount;
	const restir_tri_light *triLights;
	int triCount;
	const restir_alias_column *alias;
	int aliasCount;
};

struct TraceStats {
	float margin;  // smallest distance of any decisive comparison from its threshold
	int overflow;  // pushes dropped because the 32-entry stack was full (UB in the reference)
};

inline void noteMargin(TraceStats *st, float a, float b) {
	if (st) {
		float m = fabsf(a - b);
		if (m < st->margin) {
			st->margin = m;
		}
	}
}

// softwareRaytracing.glsl:9-14 with P3: inv = 1/dir hoisted out of the loop by the caller.
bool rayAabIntersection(V3 origin, V3 inv, const float *bmin, const float *bmax, TraceStats *st) {
	V3 t1 = v3((bmin[0] - origin.x) * inv.x, (bmin[1] - origin.y) * inv.y, (bmin[2] - origin.z) * inv.z);
	V3 t2 = v3((bmax[0] - origin.x) * inv.x, (bmax[1] - origin.y) * inv.y, (bmax[2] - origin.z) * inv.z);
	float rmin = fmaxf(fminf(t1.x, t2.x), fmaxf(fminf(t1.y, t2.y), fminf(t1.z, t2.z)));
	float rmax = fminf(fmaxf(t1.x, t2.x), fminf(fmaxf(t1.y, t2.y), fmaxf(t1.z, t2.z)));
	noteMargin(st, rmin, 1.0f);
	noteMargin(st, rmax, rmin);
	noteMargin(st, rmax, 0.0f);
	return rmin < 1.0f && rmax >= rmin && rmax > 0.0f;
}
// softwareRaytracing.glsl:15-37
bool rayTriangleIntersection(const restir_triangle &tri, V3 origin, V3 dir, TraceStats *st) {
	V3 p1 = v3(tri.p1[0], tri.p1[1], tri.p1[2]);
	V3 e1 = v3(tri.p2[0], tri.p2[1], tri.p2[2]) - p1;
	V3 e2 = v3(tri.p3[0], tri.p3[1], tri.p3[2]) - p1;
	V3 p = cross(dir, e2);
	float f = 1.0f / dot(e1, p);
	V3 s = origin - p1;
	float baryX = f * dot(s, p);
	noteMargin(st, baryX, 0.0f);
	noteMargin(st, baryX, 1.0f);
	if (baryX < 0.0f || baryX > 1.0f) {
		return false;
	}
	V3 q = cross(s, e1);
	float baryY = f * dot(dir, q);
	noteMargin(st, baryY, 0.0f);
	noteMargin(st, baryY + baryX, 1.0f);
	if (baryY < 0.0f || baryY + baryX > 1.0f) {
		return false;
	}
	f = f * dot(e2, q);
	noteMargin(st, f, 0.0f);
	noteMargin(st, f, 1.0f);
	return f > 0.0f && f < 1.0f;
}
// softwareRaytracing.glsl:39-85: stack 32, 16 deferred candidates, flush every 8 node visits.
// Returns true when NOTHING is hit.
bool raytrace(const Scene &sc, V3 origin, V3 dir, TraceStats *st) {
	const int geomTestInterval = 8, aabbTreeStackSize = 32;
	int stack[aabbTreeStackSize], top = 1;
	stack[0] = 0;
	int candidates[geomTestInterval * 2], numCandidates = 0;
	int counter = 0;
	V3 inv = v3(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z);
	while (top > 0) {
		const restir_aabb_node &node = sc.nodes[stack[--top]];
		bool leftIsect = rayAabIntersection(origin, inv, node.leftAabbMin, node.leftAabbMax, st);
		bool rightIsect = rayAabIntersection(origin, inv, node.rightAabbMin, node.rightAabbMax, st);
		if (leftIsect) {
			if (node.leftChild < 0) {
				candidates[numCandidates++] = ~node.leftChild;
			} else if (top < aabbTreeStackSize) {
				stack[top++] = node.leftChild;
			} else if (st) {
				st->overflow++;
			}
		}
		if (rightIsect) {
			if (node.rightChild < 0) {
				candidates[numCandidates++] = ~node.rightChild;
			} else if (top < aabbTreeStackSize) {
				stack[top++] = node.rightChild;
			} else if (st) {
				st->overflow++;
			}
		}
		if (++counter == geomTestInterval) {
			for (int i = 0; i < numCandidates; ++i) {
				if (rayTriangleIntersection(sc.tris[candidates[i]], origin, dir, st)) {
					return false;
				}
			}
			numCandidates = 0;
			counter = 0;
		}
	}
	for (int i = 0; i < numCandidates; ++i) {
		if (rayTriangleIntersection(sc.tris[candidates[i]], origin, dir, st)) {
			return false;
		}
	}
	return true;
}
// visibilityTest.glsl:1-4, 27-28.  Returns SHADOWED.
bool testVisibility(const Scene &sc, V3 p1, V3 p2, TraceStats *st) {
	float tMin = 0.001f;
	V3 dir = p2 - p1;
	V3 offset = normalize(dir) * tMin;
	return !raytrace(sc, p1 + offset, dir - offset * 2.0f, st);
}

// ---------------------------------------------------------------------------------------------
// G-buffer access: texelFetch of the NVIDIA-default formats (include/restir_layouts.h)
struct SrgbTable {
	float v[256];
	SrgbTable() { // P12
		for (int i = 0; i < 256; ++i) {
			double c = i / 255.0;
			v[i] = (float)(c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4));
		}
	}
};
const SrgbTable kSrgb;

struct GBuffer {
	const uint8_t *albedo;   // RGBA8 (rgb sRGB)
	const int16_t *normal;   // RGBA16_SNORM
	const uint16_t *material; // RG16_UNORM
	const float *worldPos;   // RGBA32F
	const float *depth;      // D32F
};
struct Texel {
	V3 albedo;
	float albedoA;
	V3 normal;
	float roughness, metallic;
	V3 worldPos;
};
inline V3 fetchAlbedo(const GBuffer &g, size_t i, float *a) {
	if (!g.albedo) {
		if (a) *a = 0.0f;
		return v3(0, 0, 0);
	}
	const uint8_t *p = g.albedo + i * 4;
	if (a) *a = (float)p[3] / 255.0f;
	return v3(kSrgb.v[p[0]], kSrgb.v[p[1]], kSrgb.v[p[2]]);
}
inline V3 fetchNormal(const GBuffer &g, size_t i) {
	if (!g.normal) {
		return v3(0, 0, 0);
	}
	const int16_t *p = g.normal + i * 4;
	return v3(fmaxf((float)p[0] / 32767.0f, -1.0f), fmaxf((float)p[1] / 32767.0f, -1.0f), fmaxf((float)p[2] / 32767.0f, -1.0f));
}
inline void fetchMaterial(const GBuffer &g, size_t i, float *roughness, float *metallic) {
	const uint16_t *p = g.material + i * 2;
	*roughness = (float)p[0] / 65535.0f;
	*metallic = (float)p[1] / 65535.0f;
}
inline V3 fetchWorldPos(const GBuffer &g, size_t i) {
	if (!g.worldPos) {
		return v3(0, 0, 0);
	}
	const float *p = g.worldPos + i * 4;
	return v3(p[0], p[1], p[2]);
}

Scene makeScene(const void *nodes, const void *tris, const void *pointBlob, const void *triBlob, const void *aliasBlob) {
	Scene sc{};
	sc.nodes = (const restir_aabb_node *)nodes;
	sc.tris = (const restir_triangle *)tris;
	if (pointBlob) {
		std::memcpy(&sc.pointCount, pointBlob, 4);
		sc.pointLights = (const restir_point_light *)((const uint8_t *)pointBlob + RESTIR_BLOB_HEADER_BYTES);
	}
	if (triBlob) {
		std::memcpy(&sc.triCount, triBlob, 4);
		sc.triLights = (const restir_tri_light *)((const uint8_t *)triBlob + RESTIR_BLOB_HEADER_BYTES);
	}
	if (aliasBlob) {
		std::memcpy(&sc.aliasCount, aliasBlob, 4);
		sc.alias = (const restir_alias_column *)((const uint8_t *)aliasBlob + RESTIR_BLOB_HEADER_BYTES);
	}
	return sc;
}

// restirOmni.glsl:68-71
V3 pickPointOnTriangle(float r1, float r2, V3 p1, V3 p2, V3 p3) {
	float sqrt_r1 = sqrtf(r1);
	return (p1 * (1.0f - sqrt_r1) + p2 * (sqrt_r1 * (1.0f - r2))) + p3 * (r2 * sqrt_r1);
}
// restirOmni.glsl:73-83
void aliasTableSample(const Scene &sc, float r1, float r2, int *index, float *probability) {
	int selected = (int)((float)sc.aliasCount * r1);
	if (selected > sc.aliasCount - 1) {
		selected = sc.aliasCount - 1;
	}
	const restir_alias_column &col = sc.alias[selected];
	if (col.prob > r2) {
		*index = selected;
		*probability = col.oriProb;
	} else {
		*index = col.alias;
		*probability = col.aliasOriProb;
	}
}

} // namespace

extern "C" {

struct oracle_gbuffer {
	const void *albedo, *normal, *material, *worldPos, *depth;
};
struct oracle_scene {
	const void *nodes, *tris, *pointBlob, *triBlob, *aliasBlob;
};

static GBuffer toG(const oracle_gbuffer *g) {
	GBuffer r{};
	if (g) {
		r.albedo = (const uint8_t *)g->albedo;
		r.normal = (const int16_t *)g->normal;
		r.material = (const uint16_t *)g->material;
		r.worldPos = (const float *)g->worldPos;
		r.depth = (const float *)g->depth;
	}
	return r;
}

// KAT hooks -----------------------------------------------------------------------------------
void oracle_pcg32(uint64_t seed, uint64_t seq, int n, uint32_t *out) {
	Rand r = seedRand(seed, seq);
	for (int i = 0; i < n; ++i) {
		out[i] = randUint(r);
	}
}
void oracle_rand_floats(uint64_t seed, uint64_t seq, int n, float *out) {
	Rand r = seedRand(seed, seq);
	for (int i = 0; i < n; ++i) {
		out[i] = randFloat(r);
	}
}
void oracle_sincos(const float *a, int n, float *s, float *c) {
	for (int i = 0; i < n; ++i) {
		det_sincos(a[i], &s[i], &c[i]);
	}
}
// evaluatePHat on n independent argument tuples (SoA of 16 floats each:
// worldPos3 lightPos3 camPos3 normal3 lightNormal3 useLightNormal)
void oracle_evaluate_phat(const float *args, int n, float albedoLum, float emissionLum, float roughness,
                          float metallic, float *out) {
	for (int i = 0; i < n; ++i) {
		const float *a = args + (size_t)i * 16;
		out[i] = evaluatePHat(v3(a[0], a[1], a[2]), v3(a[3], a[4], a[5]), v3(a[6], a[7], a[8]), v3(a[9], a[10], a[11]),
		                      v3(a[12], a[13], a[14]), a[15] > 0.5f, albedoLum, emissionLum, roughness, metallic);
	}
}

// testVisibility(p1, p2) for n segments; shadowed[i] in {0,1}; margin/overflow optional.
void oracle_trace_segments(const oracle_scene *s, int64_t n, const float *p1, const float *p2, uint8_t *shadowed,
                           float *margin, int32_t *overflow) {
	Scene sc = makeScene(s->nodes, s->tris, nullptr, nullptr, nullptr);
#pragma omp parallel for schedule(dynamic, 256)
	for (int64_t i = 0; i < n; ++i) {
		TraceStats st{INFINITY, 0};
		bool sh = testVisibility(sc, v3(p1[i * 3], p1[i * 3 + 1], p1[i * 3 + 2]), v3(p2[i * 3], p2[i * 3 + 1], p2[i * 3 + 2]),
		                         (margin || overflow) ? &st : nullptr);
		shadowed[i] = sh ? 1 : 0;
		if (margin) margin[i] = st.margin;
		if (overflow) overflow[i] = st.overflow;
	}
}

// restirOmni.glsl:86-212 over rows [y0, y1).  `prev` may be NULL (first frame: zero G-buffer).
// reservoirs are indexed y*W + x over the whole screen.  rays += executed testVisibility calls.
void oracle_restir_pass(const oracle_scene *s, const restir_uniforms *u, const oracle_gbuffer *cur,
                        const oracle_gbuffer *prev, const restir_reservoir *prevFrameReservoirs,
                        restir_reservoir *reservoirs, int y0, int y1, uint64_t *rays) {
	Scene sc = makeScene(s->nodes, s->tris, s->pointBlob, s->triBlob, s->aliasBlob);
	GBuffer g = toG(cur), pg = toG(prev);
	const int W = (int)u->screenSize[0], H = (int)u->screenSize[1];
	const V3 camPos = v3(u->cameraPos[0], u->cameraPos[1], u->cameraPos[2]);
	const float *M = u->prevFrameProjectionViewMatrix;
	uint64_t rayCount = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : rayCount)
	for (int y = y0; y < y1; ++y) {
		for (int x = 0; x < W; ++x) {
			size_t pix = (size_t)y * W + x;
			// :98-103
			V3 albedo = fetchAlbedo(g, pix, nullptr);
			V3 normal = fetchNormal(g, pix);
			float roughness, metallic;
			fetchMaterial(g, pix, &roughness, &metallic);
			V3 worldPos = fetchWorldPos(g, pix);
			float albedoLum = luminance(albedo.x, albedo.y, albedo.z);

			Reservoir res = newReservoir();                                        // :105
			Rand rand = seedRand(u->frame, (uint32_t)((uint32_t)y * 10007u + (uint32_t)x)); // :106
			if (dot(normal, normal) != 0.0f) {                                     // :107
				for (uint32_t i = 0; i < u->initialLightSampleCount; ++i) {       // :108
					int selected_idx;
					float lightSampleProb;
					float r1 = randFloat(rand); // GLSL argument order: left to right (Appendix B.3)
					float r2 = randFloat(rand);
					aliasTableSample(sc, r1, r2, &selected_idx, &lightSampleProb);

					V3 lightSamplePos;
					float lightNormal[4];
					float lightSampleLum;
					int lightSampleIndex;
					if (sc.pointCount != 0) { // :116-122
						const restir_point_light &light = sc.pointLights[selected_idx];
						lightSamplePos = v3(light.pos[0], light.pos[1], light.pos[2]);
						lightSampleLum = light.color_luminance[3];
						lightSampleIndex = selected_idx;
						lightNormal[0] = lightNormal[1] = lightNormal[2] = lightNormal[3] = 0.0f;
					} else { // :123-133
						const restir_tri_light &light = sc.triLights[selected_idx];
						float r3 = randFloat(rand);
						float r4 = randFloat(rand);
						lightSamplePos = pickPointOnTriangle(r3, r4, v3(light.p1[0], light.p1[1], light.p1[2]),
						                                     v3(light.p2[0], light.p2[1], light.p2[2]),
						                                     v3(light.p3[0], light.p3[1], light.p3[2]));
						lightSampleLum = light.emission_luminance[3];
						lightSampleIndex = -1 - selected_idx;
						V3 wi = normalize(worldPos - lightSamplePos);
						V3 ln = v3(light.normalArea[0], light.normalArea[1], light.normalArea[2]);
						lightSampleProb = lightSampleProb / (fabsf(dot(wi, ln)) * light.normalArea[3]);
						lightNormal[0] = ln.x;
						lightNormal[1] = ln.y;
						lightNormal[2] = ln.z;
						lightNormal[3] = 1.0f;
					}
					float pHat = evaluatePHat(worldPos, lightSamplePos, camPos, normal,
					                          v3(lightNormal[0], lightNormal[1], lightNormal[2]), lightNormal[3] > 0.5f,
					                          albedoLum, lightSampleLum, roughness, metallic); // :135-139
					addSampleToReservoir(res, lightSamplePos, lightNormal, lightSampleLum, lightSampleIndex, pHat,
					                     lightSampleProb, rand); // :141
				}
			}

			// Visibility reuse :148-160 (M is kept: Appendix B.5)
			if ((u->flags & RESTIR_VISIBILITY_REUSE_FLAG) != 0) {
				bool shadowed = testVisibility(sc, worldPos, samplePos(res), nullptr);
				rayCount++;
				if (shadowed) {
					res.sample.w = 0.0f;
					res.sample.sumWeights = 0.0f;
				}
			}

			// Temporal reuse :163-209
			if ((u->flags & RESTIR_TEMPORAL_REUSE_FLAG) != 0) {
				float px = ((M[0] * worldPos.x + M[4] * worldPos.y) + M[8] * worldPos.z) + M[12] * 1.0f;
				float py = ((M[1] * worldPos.x + M[5] * worldPos.y) + M[9] * worldPos.z) + M[13] * 1.0f;
				float pw = ((M[3] * worldPos.x + M[7] * worldPos.y) + M[11] * worldPos.z) + M[15] * 1.0f;
				float invW = 1.0f / pw; // P3
				px = px * invW;
				py = py * invW;
				px = ((px + 1.0f) * 0.5f) * (float)W;
				py = ((py + 1.0f) * 0.5f) * (float)H;
				if (px > 0.0f && py > 0.0f && px < (float)W && py < (float)H) {
					int fx = (int)px, fy = (int)py;
					size_t ppix = (size_t)fy * W + fx;
					V3 positionDiff = worldPos - fetchWorldPos(pg, ppix);
					if (dot(positionDiff, positionDiff) < 0.01f) {
						V3 albedoDiff = albedo - fetchAlbedo(pg, ppix, nullptr);
						if (dot(albedoDiff, albedoDiff) < 0.01f) {
							float normalDot = dot(normal, fetchNormal(pg, ppix));
							if (normalDot > 0.5f) {
								Reservoir prevRes = prevFrameReservoirs[ppix];
								uint32_t cap = u->temporalSampleCountMultiplier * res.numStreamSamples;
								if (prevRes.numStreamSamples > cap) { // :189-191
									prevRes.numStreamSamples = cap;
								}
								float pHat = evaluatePHat(worldPos, samplePos(prevRes), camPos, normal, sampleNormal(prevRes),
								                          prevRes.sample.normal[3] > 0.5f, albedoLum,
								                          prevRes.sample.position_emissionLum[3], roughness, metallic);
								combineReservoirs(res, prevRes, pHat, rand);
							}
						}
					}
				}
			}
			reservoirs[pix] = res; // :211
		}
	}
	(void)H;
	if (rays) {
		*rays += rayCount;
	}
}

// spatialReuse.comp:30-86 over rows [y0, y1)
void oracle_spatial_pass(const restir_uniforms *u, const oracle_gbuffer *cur, const restir_reservoir *reservoirs,
                         restir_reservoir *resultReservoirs, int iter, int y0, int y1) {
	GBuffer g = toG(cur);
	const int W = (int)u->screenSize[0], H = (int)u->screenSize[1];
	const V3 camPos = v3(u->cameraPos[0], u->cameraPos[1], u->cameraPos[2]);
	float sinThr, cosThr;
	det_sincos(u->spatialNormalThreshold * 0.017453292519943295f, &sinThr, &cosThr); // :67, P7/P8
#pragma omp parallel for schedule(dynamic, 1)
	for (int y = y0; y < y1; ++y) {
		for (int x = 0; x < W; ++x) {
			size_t pix = (size_t)y * W + x;
			V3 albedo = fetchAlbedo(g, pix, nullptr);
			V3 normal = fetchNormal(g, pix);
			float roughness, metallic;
			fetchMaterial(g, pix, &roughness, &metallic);
			V3 worldPos = fetchWorldPos(g, pix);
			float worldDepth = g.depth[pix];
			float albedoLum = luminance(albedo.x, albedo.y, albedo.z);

			Reservoir res = reservoirs[pix];
			Rand rand = seedRand((uint32_t)(u->frame * 31u + (uint32_t)iter), (uint32_t)((uint32_t)y * 10007u + (uint32_t)x)); // :47
			for (uint32_t i = 0; i < u->spatialNeighbors; ++i) {
				float angle = (randFloat(rand) * 2.0f) * kPi;                   // :52
				float radius = sqrtf(randFloat(rand)) * u->spatialRadius;       // :53
				float sn, cs;
				det_sincos(angle, &sn, &cs);
				int ox = (int)floorf(cs * radius), oy = (int)floorf(sn * radius); // :55
				int nx = x + ox, ny = y + oy;                                    // :56-57
				nx = nx < 0 ? 0 : (nx > W - 1 ? W - 1 : nx);
				ny = ny < 0 ? 0 : (ny > H - 1 ? H - 1 : ny);
				size_t npix = (size_t)ny * W + nx;

				float neighborDepth = g.depth[npix];
				V3 neighborNor = fetchNormal(g, npix);
				if (fabsf(neighborDepth - worldDepth) > u->spatialPosThreshold * fabsf(worldDepth) ||
				    dot(neighborNor, normal) < cosThr) { // :65-70
					continue;
				}
				Reservoir randRes = reservoirs[npix];
				float newPHat = evaluatePHat(worldPos, samplePos(randRes), camPos, normal, sampleNormal(randRes),
				                             randRes.sample.normal[3] > 0.5f, albedoLum,
				                             randRes.sample.position_emissionLum[3], roughness, metallic);
				combineReservoirs(res, randRes, newPHat, rand);
			}
			resultReservoirs[pix] = res;
		}
	}
}

// unbiasedReuse.glsl:50-185 over rows [y0, y1); NUM_NEIGHBORS is the reference's #define 3 unless
// numNeighbors overrides it (north_star asks for 5 as well).
void oracle_unbiased_pass(const oracle_scene *s, const restir_uniforms *u, const oracle_gbuffer *cur,
                          const restir_reservoir *reservoirs, restir_reservoir *resultReservoirs, int numNeighbors,
                          int y0, int y1, uint64_t *rays) {
	Scene sc = makeScene(s->nodes, s->tris, nullptr, nullptr, nullptr);
	GBuffer g = toG(cur);
	const int W = (int)u->screenSize[0], H = (int)u->screenSize[1];
	const V3 camPos = v3(u->cameraPos[0], u->cameraPos[1], u->cameraPos[2]);
	const int NN = numNeighbors > 0 ? (numNeighbors > 16 ? 16 : numNeighbors) : 3;
	uint64_t rayCount = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : rayCount)
	for (int y = y0; y < y1; ++y) {
		for (int x = 0; x < W; ++x) {
			size_t pix = (size_t)y * W + x;
			V3 albedo = fetchAlbedo(g, pix, nullptr);
			V3 normal = fetchNormal(g, pix);
			float roughness, metallic;
			fetchMaterial(g, pix, &roughness, &metallic);
			V3 worldPos = fetchWorldPos(g, pix);
			float albedoLum = luminance(albedo.x, albedo.y, albedo.z);

			Reservoir res = reservoirs[pix];
			Rand rand = seedRand((uint32_t)(u->frame * 17u), (uint32_t)((uint32_t)y * 10007u + (uint32_t)x)); // :72
			size_t neighborPix[16];
			uint32_t neighborNumSamples[16];
			uint32_t originalNumSamples = res.numStreamSamples;
			for (int i = 0; i < NN; ++i) { // :84-124
				float angle = (randFloat(rand) * 2.0f) * kPi;
				float radius = sqrtf(randFloat(rand)) * u->spatialRadius;
				float sn, cs;
				det_sincos(angle, &sn, &cs);
				int nx = x + (int)roundf(cs * radius), ny = y + (int)roundf(sn * radius); // :88-89, P9
				nx = nx < 0 ? 0 : (nx > W - 1 ? W - 1 : nx);
				ny = ny < 0 ? 0 : (ny > H - 1 ? H - 1 : ny);
				size_t npix = (size_t)ny * W + nx;
				Reservoir randRes = reservoirs[npix];
				neighborPix[i] = npix;
				neighborNumSamples[i] = randRes.numStreamSamples;

				res.numStreamSamples += randRes.numStreamSamples; // :104 (regardless of visibility, B.10)
				float newPHat = evaluatePHat(worldPos, samplePos(randRes), camPos, normal, sampleNormal(randRes),
				                             randRes.sample.normal[3] > 0.5f, albedoLum,
				                             randRes.sample.position_emissionLum[3], roughness, metallic);
				float weight = (newPHat * randRes.sample.w) * (float)randRes.numStreamSamples;
				if (weight > 0.0f) {
					updateReservoirAt(res, weight, samplePos(randRes), randRes.sample.normal,
					                  randRes.sample.position_emissionLum[3], randRes.sample.lightIndex, newPHat,
					                  randRes.sample.w, rand);
				}
			}
			// :126-182
			V3 lightPos = samplePos(res);
			uint32_t numSamples = originalNumSamples;
			for (int j = 0; j < NN; ++j) {
				V3 nPos = fetchWorldPos(g, neighborPix[j]);
				V3 nNor = fetchNormal(g, neighborPix[j]);
				if (dot(lightPos - nPos, nNor) < 0.0f) {
					continue;
				}
				if ((u->flags & RESTIR_VISIBILITY_REUSE_FLAG) != 0) {
					bool shadowed = testVisibility(sc, nPos, lightPos, nullptr);
					rayCount++;
					if (shadowed) {
						continue;
					}
				}
				numSamples += neighborNumSamples[j];
			}
			if ((u->flags & RESTIR_VISIBILITY_REUSE_FLAG) != 0) {
				bool shadowed = testVisibility(sc, worldPos, lightPos, nullptr);
				rayCount++;
				if (shadowed) {
					numSamples = 0;
				}
			}
			if (numSamples > 0) {
				res.sample.w = res.sample.sumWeights / ((float)numSamples * res.sample.pHat);
			} else {
				res.sample.w = 0.0f;
				res.sample.sumWeights = 0.0f;
			}
			resultReservoirs[pix] = res;
		}
	}
	if (rays) {
		*rays += rayCount;
	}
}

// lighting.frag:43-71,103 (debugMode 0) over rows [y0, y1); writes linear RGBA32F (a = 1), i.e. the
// value BEFORE the swapchain's 8-bit sRGB quantisation.
void oracle_lighting_pass(const oracle_scene *s, const restir_lighting_uniforms *u, const oracle_gbuffer *cur,
                          const restir_reservoir *reservoirs, float *outRgba, int y0, int y1) {
	Scene sc = makeScene(nullptr, nullptr, s->pointBlob, s->triBlob, nullptr);
	GBuffer g = toG(cur);
	const int W = (int)u->bufferSize[0];
	const V3 camPos = v3(u->cameraPos[0], u->cameraPos[1], u->cameraPos[2]);
#pragma omp parallel for schedule(dynamic, 4)
	for (int y = y0; y < y1; ++y) {
		for (int x = 0; x < W; ++x) {
			size_t pix = (size_t)y * W + x;
			float albedoA;
			V3 albedo = fetchAlbedo(g, pix, &albedoA);
			V3 normal = fetchNormal(g, pix);
			float roughness, metallic;
			fetchMaterial(g, pix, &roughness, &metallic);
			V3 worldPos = fetchWorldPos(g, pix);
			const Reservoir &r = reservoirs[pix];
			V3 emission;
			int lightIndex = r.sample.lightIndex;
			emission = v3(0, 0, 0); // :55-60; an index outside the bound SSBO reads 0 (robust buffer access)
			if (lightIndex < 0) {
				if (-1 - lightIndex < sc.triCount) {
					const restir_tri_light &l = sc.triLights[-1 - lightIndex];
					emission = v3(l.emission_luminance[0], l.emission_luminance[1], l.emission_luminance[2]);
				}
			} else if (lightIndex < sc.pointCount) {
				const restir_point_light &l = sc.pointLights[lightIndex];
				emission = v3(l.color_luminance[0], l.color_luminance[1], l.color_luminance[2]);
			}
			V3 pHat = evaluatePHatFull(worldPos, samplePos(r), camPos, normal, sampleNormal(r), r.sample.normal[3] > 0.5f,
			                           albedo, emission, roughness, metallic);
			V3 c = pHat * r.sample.w; // :66; "/ RESERVOIR_SIZE" is / 1
			if (albedoA > 0.5f) {     // :69-71 (emissive AND background, Appendix B.12)
				c = albedo;
			}
			if (u->gamma != 1.0f) { // :103, P11
				float e = 1.0f / u->gamma;
				c = v3(powf(c.x, e), powf(c.y, e), powf(c.z, e));
			}
			float *o = outRgba + pix * 4;
			o[0] = c.x;
			o[1] = c.y;
			o[2] = c.z;
			o[3] = 1.0f;
		}
	}
}

// ---------------------------------------------------------------------------------------------
// Fixture generator (SURVEY §8d "Concrete synthetic inputs", Appendix E): primary-visibility ray
// cast of the same triangle list through the same AABB tree, one sample at the pixel centre,
// factor-only materials, quantised to the NVIDIA-default formats.  Camera = src/camera.h:25-50.
// Not part of the reference's hot path (its G-buffer comes from a rasteriser): this only
// synthesises INPUTS.  The CUDA twin is restir_tools_raycast_gbuffer.
struct oracle_camera {
	float position[3], lookAt[3], worldUp[3];
	float zNear, zFar, fovYRadians, aspectRatio;
};

// camera.h:25-50: projectionViewMatrix (column-major out[16])
void oracle_camera_matrix(const oracle_camera *c, float *outPV) {
	V3 pos = v3(c->position[0], c->position[1], c->position[2]);
	V3 fwd = normalize(v3(c->lookAt[0], c->lookAt[1], c->lookAt[2]) - pos);
	V3 right = normalize(cross(fwd, v3(c->worldUp[0], c->worldUp[1], c->worldUp[2])));
	V3 up = cross(right, fwd);
	V3 r0 = right, r1 = v3(-up.x, -up.y, -up.z), r2 = fwd;
	float off[3] = {-dot(r0, pos), -dot(r1, pos), -dot(r2, pos)};
	float view[4][4] = {{r0.x, r0.y, r0.z, off[0]}, {r1.x, r1.y, r1.z, off[1]}, {r2.x, r2.y, r2.z, off[2]}, {0, 0, 0, 1}};
	float f = 1.0f / tanf(0.5f * c->fovYRadians);
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
			outPV[col * 4 + r] = acc;
		}
	}
}

// materialTable: per material {u32 albedoRGBA8, u32 materialRG16, u32 flags(bit0 = discard), u32 pad}
void oracle_raycast_gbuffer(const oracle_scene *s, const int32_t *triMaterial, const uint32_t *materialTable,
                            const oracle_camera *c, int W, int H, int y0, int y1, uint8_t *albedo, int16_t *normal,
                            uint16_t *material, float *worldPos, float *depth) {
	const restir_aabb_node *nodes = (const restir_aabb_node *)s->nodes;
	const restir_triangle *tris = (const restir_triangle *)s->tris;
	V3 pos = v3(c->position[0], c->position[1], c->position[2]);
	V3 fwd = normalize(v3(c->lookAt[0], c->lookAt[1], c->lookAt[2]) - pos);
	V3 right = normalize(cross(fwd, v3(c->worldUp[0], c->worldUp[1], c->worldUp[2])));
	V3 up = cross(right, fwd);
	float f = 1.0f / tanf(0.5f * c->fovYRadians);
	float sx = c->aspectRatio / f, sy = 1.0f / f;
	float PV[16];
	oracle_camera_matrix(c, PV);
#pragma omp parallel for schedule(dynamic, 1)
	for (int y = y0; y < y1; ++y) {
		for (int x = 0; x < W; ++x) {
			size_t pix = (size_t)y * W + x;
			float ndcx = (((float)x + 0.5f) / (float)W) * 2.0f - 1.0f;
			float ndcy = (((float)y + 0.5f) / (float)H) * 2.0f - 1.0f;
			V3 dir = (fwd + right * (ndcx * sx)) - up * (ndcy * sy);
			V3 inv = v3(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z);
			float best = INFINITY, bu = 0, bv = 0;
			int bestTri = -1;
			int stack[64], top = 1;
			stack[0] = 0;
			while (top > 0) {
				const restir_aabb_node &node = nodes[stack[--top]];
				for (int side = 0; side < 2; ++side) {
					const float *bmin = side ? node.rightAabbMin : node.leftAabbMin;
					const float *bmax = side ? node.rightAabbMax : node.leftAabbMax;
					int child = side ? node.rightChild : node.leftChild;
					V3 t1 = v3((bmin[0] - pos.x) * inv.x, (bmin[1] - pos.y) * inv.y, (bmin[2] - pos.z) * inv.z);
					V3 t2 = v3((bmax[0] - pos.x) * inv.x, (bmax[1] - pos.y) * inv.y, (bmax[2] - pos.z) * inv.z);
					float rmin = fmaxf(fminf(t1.x, t2.x), fmaxf(fminf(t1.y, t2.y), fminf(t1.z, t2.z)));
					float rmax = fminf(fmaxf(t1.x, t2.x), fminf(fmaxf(t1.y, t2.y), fmaxf(t1.z, t2.z)));
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
					const restir_triangle &tri = tris[ti];
					V3 p1 = v3(tri.p1[0], tri.p1[1], tri.p1[2]);
					V3 e1 = v3(tri.p2[0], tri.p2[1], tri.p2[2]) - p1;
					V3 e2 = v3(tri.p3[0], tri.p3[1], tri.p3[2]) - p1;
					V3 n = cross(e1, e2);
					if (!(dot(n, dir) < 0.0f)) { // back-face culling, CCW front (pass.h:24-32)
						continue;
					}
					if (materialTable[(size_t)triMaterial[ti] * 4 + 2] & 1u) { // ALPHA_MODE_MASK below cutoff
						continue;
					}
					V3 p = cross(dir, e2);
					float fdet = 1.0f / dot(e1, p);
					V3 sv = pos - p1;
					float u_ = fdet * dot(sv, p);
					if (u_ < 0.0f || u_ > 1.0f) {
						continue;
					}
					V3 q = cross(sv, e1);
					float v_ = fdet * dot(dir, q);
					if (v_ < 0.0f || v_ + u_ > 1.0f) {
						continue;
					}
					float t = fdet * dot(e2, q);
					if (t > 0.0f && (t < best || (t == best && ti < bestTri))) {
						best = t;
						bestTri = ti;
						bu = u_;
						bv = v_;
					}
				}
			}
			uint8_t *a = albedo + pix * 4;
			int16_t *nq = normal + pix * 4;
			uint16_t *m = material + pix * 2;
			float *wp = worldPos + pix * 4;
			if (bestTri < 0) { // clears: gBufferPass.cpp:117-123
				a[0] = a[1] = a[2] = 0;
				a[3] = 255;
				nq[0] = nq[1] = nq[2] = 0;
				nq[3] = 32767;
				m[0] = m[1] = 0;
				wp[0] = wp[1] = wp[2] = 0.0f;
				wp[3] = 1.0f;
				depth[pix] = 1.0f;
				continue;
			}
			const restir_triangle &tri = tris[bestTri];
			V3 p1 = v3(tri.p1[0], tri.p1[1], tri.p1[2]);
			V3 p2 = v3(tri.p2[0], tri.p2[1], tri.p2[2]);
			V3 p3 = v3(tri.p3[0], tri.p3[1], tri.p3[2]);
			V3 n = normalize(cross(p2 - p1, p3 - p1));
			V3 hit = (p1 * ((1.0f - bu) - bv) + p2 * bu) + p3 * bv;
			uint32_t mat = materialTable[(size_t)triMaterial[bestTri] * 4];
			uint32_t mp = materialTable[(size_t)triMaterial[bestTri] * 4 + 1];
			std::memcpy(a, &mat, 4);
			std::memcpy(m, &mp, 4);
			nq[0] = (int16_t)rintf(clampf(n.x, -1.0f, 1.0f) * 32767.0f);
			nq[1] = (int16_t)rintf(clampf(n.y, -1.0f, 1.0f) * 32767.0f);
			nq[2] = (int16_t)rintf(clampf(n.z, -1.0f, 1.0f) * 32767.0f);
			nq[3] = 32767;
			wp[0] = hit.x;
			wp[1] = hit.y;
			wp[2] = hit.z;
			wp[3] = 1.0f;
			float cz = ((PV[2] * hit.x + PV[6] * hit.y) + PV[10] * hit.z) + PV[14];
			float cw = ((PV[3] * hit.x + PV[7] * hit.y) + PV[11] * hit.z) + PV[15];
			depth[pix] = cz / cw;
		}
	}
}

// torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm of bench.py asks for every host core explicitly
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
	if (n > 0) omp_set_num_threads(n);
#endif
}

int oracle_num_threads(void) {
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

} // extern "C"
// ---------------------------------------------------------------------------------------------
// The G-buffer pass: CPU twin of restir_pass_gbuffer.  Follows src/shaders/gBuffer.vert:22-34,
// src/shaders/gBuffer.frag:27-80, src/passes/gBufferPass.cpp:116-157 (clears, draw order, depth test LESS) and
// src/passes/pass.h:24-40 (back-face culling).  Primary visibility by ray casting, like oracle_raycast_gbuffer.
// Texture policy: R8G8B8A8_UNORM (sceneBuffers.h:126), repeat wrap, bilinear at level 0, texel = code / 255, the two
// lerps written out; sRGB attachment encode = largest code whose lower threshold (EOTF of the code midpoint, in double) the
// value reaches; SNORM16 / UNORM16 = rintf of the clamped value.

namespace {
struct F4o {
	float x, y, z, w;
};
struct GbTextures {
	const uint8_t *texels;   // every texture's level 0, RGBA8, back to back
	const uint32_t *table;   // per texture: first texel, width, height, -
	int count;
};
V3 matMul(const float *m, V3 v, float w) { // column-major mat4 * (v, w), P4
	return v3(((m[0] * v.x + m[4] * v.y) + m[8] * v.z) + m[12] * w, ((m[1] * v.x + m[5] * v.y) + m[9] * v.z) + m[13] * w,
	          ((m[2] * v.x + m[6] * v.y) + m[10] * v.z) + m[14] * w);
}
F4o sampleTexture(const GbTextures &tx, int index, bool normalBinding, float u, float v) {
	if (index < 0 || index >= tx.count) {
		if (normalBinding) {
			return F4o{127.0f / 255.0f, 127.0f / 255.0f, 1.0f, 1.0f};
		}
		return F4o{1.0f, 1.0f, 1.0f, 1.0f};
	}
	const uint32_t *d = tx.table + (size_t)index * 4;
	const int W = (int)d[1], H = (int)d[2];
	float x = u * (float)W - 0.5f, y = v * (float)H - 0.5f;
	if (!(fabsf(x) < 1.0e9f) || !(fabsf(y) < 1.0e9f)) {
		x = 0.0f;
		y = 0.0f;
	}
	float fx = floorf(x), fy = floorf(y);
	float ax = x - fx, ay = y - fy;
	int x0 = (int)fx % W, y0 = (int)fy % H;
	x0 = x0 < 0 ? x0 + W : x0;
	y0 = y0 < 0 ? y0 + H : y0;
	int x1 = x0 + 1 == W ? 0 : x0 + 1, y1 = y0 + 1 == H ? 0 : y0 + 1;
	const uint8_t *t = tx.texels + (size_t)d[0] * 4;
	const uint8_t *c00 = t + ((size_t)y0 * W + x0) * 4, *c10 = t + ((size_t)y0 * W + x1) * 4;
	const uint8_t *c01 = t + ((size_t)y1 * W + x0) * 4, *c11 = t + ((size_t)y1 * W + x1) * 4;
	float bx = 1.0f - ax, by = 1.0f - ay;
	float r[4];
	for (int k = 0; k < 4; ++k) {
		float top = (float)c00[k] / 255.0f * bx + (float)c10[k] / 255.0f * ax;
		float bot = (float)c01[k] / 255.0f * bx + (float)c11[k] / 255.0f * ax;
		r[k] = top * by + bot * ay;
	}
	return F4o{r[0], r[1], r[2], r[3]};
}
struct SrgbThresholds {
	float thr[256];
	SrgbThresholds() {
		thr[0] = -INFINITY;
		for (int c = 1; c < 256; ++c) {
			double e = (c - 0.5) / 255.0;
			double lin = e <= 0.04045 ? e / 12.92 : pow((e + 0.055) / 1.055, 2.4);
			float f = (float)lin;
			if ((double)f < lin) f = nextafterf(f, INFINITY);
			thr[c] = f;
		}
	}
	uint8_t code(float c) const {
		int best = 0;
		for (int k = 255; k >= 1; --k) {
			if (c >= thr[k]) {
				best = k;
				break;
			}
		}
		return (uint8_t)best;
	}
};
} // namespace

extern "C" {

// gBuffer.vert:22-34 per triangle corner, in draw order; out: nTris x 32 floats (N0 u0 | T0 | N1 u1 | T1 | N2 u2 | T2 | v0 v1 v2 - | -)
void oracle_vertex_stage(const restir_vertex *vertices, const uint32_t *indices, const restir_draw *draws, const restir_model_matrices *matrices,
                         uint32_t nDraws, float *attrs, int32_t *triMaterial) {
	size_t t = 0;
	for (uint32_t d = 0; d < nDraws; ++d) {
		const restir_draw &dr = draws[d];
		for (uint32_t i = 0; i < dr.indexCount; i += 3, ++t) {
			float *o = attrs + t * 32;
			std::memset(o, 0, 32 * sizeof(float));
			for (int k = 0; k < 3; ++k) {
				const restir_vertex &vx = vertices[(size_t)dr.vertexOffset + indices[dr.firstIndex + i + k]];
				V3 n = normalize(matMul(matrices[d].transformInverseTransposed, v3(vx.normal[0], vx.normal[1], vx.normal[2]), 0.0f));
				V3 tg = normalize(matMul(matrices[d].transform, v3(vx.tangent[0], vx.tangent[1], vx.tangent[2]), 0.0f));
				o[8 * k + 0] = n.x; o[8 * k + 1] = n.y; o[8 * k + 2] = n.z; o[8 * k + 3] = vx.uv[0];
				o[8 * k + 4] = tg.x; o[8 * k + 5] = tg.y; o[8 * k + 6] = tg.z; o[8 * k + 7] = vx.tangent[3];
				o[24 + k] = vx.uv[1];
			}
			triMaterial[t] = dr.materialIndex;
		}
	}
}

void oracle_gbuffer_pass(const oracle_scene *s, const float *attrs, const int32_t *triMaterial, const restir_material_uniforms *uniforms,
                         const restir_material_textures *bindings, int nMaterials, const uint8_t *texels, const uint32_t *textureTable,
                         int nTextures, const oracle_camera *c, int W, int H, int y0, int y1, uint8_t *albedoOut, int16_t *normalOut,
                         uint16_t *materialOut, float *worldPosOut, float *depthOut) {
	static const SrgbThresholds srgb;
	const restir_aabb_node *nodes = (const restir_aabb_node *)s->nodes;
	const restir_triangle *tris = (const restir_triangle *)s->tris;
	const GbTextures tx{texels, textureTable, nTextures};
	V3 pos = v3(c->position[0], c->position[1], c->position[2]);
	V3 fwd = normalize(v3(c->lookAt[0], c->lookAt[1], c->lookAt[2]) - pos);
	V3 right = normalize(cross(fwd, v3(c->worldUp[0], c->worldUp[1], c->worldUp[2])));
	V3 up = cross(right, fwd);
	float f = 1.0f / tanf(0.5f * c->fovYRadians);
	float sx = c->aspectRatio / f, sy = 1.0f / f;
	float PV[16];
	oracle_camera_matrix(c, PV);
#pragma omp parallel for schedule(dynamic, 1)
	for (int y = y0; y < y1; ++y) {
		for (int x = 0; x < W; ++x) {
			size_t pix = (size_t)y * W + x;
			float ndcx = (((float)x + 0.5f) / (float)W) * 2.0f - 1.0f;
			float ndcy = (((float)y + 0.5f) / (float)H) * 2.0f - 1.0f;
			V3 dir = (fwd + right * (ndcx * sx)) - up * (ndcy * sy);
			V3 inv = v3(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z);
			float best = INFINITY, bu = 0, bv = 0;
			int bestTri = -1;
			// nearer child first (slab entry parameter; ties: left first), like the kernel: the pruning by `best` makes the
			// result depend on the order in the last bit of near-ties, so both sides walk the same way
			int stack[64], top = 1;
			stack[0] = 0;
			while (top > 0) {
				const restir_aabb_node &node = nodes[stack[--top]];
				const int child[2] = {node.leftChild, node.rightChild};
				float rminOf[2];
				bool hitBox[2];
				for (int side = 0; side < 2; ++side) {
					const float *bmin = side ? node.rightAabbMin : node.leftAabbMin;
					const float *bmax = side ? node.rightAabbMax : node.leftAabbMax;
					V3 t1 = v3((bmin[0] - pos.x) * inv.x, (bmin[1] - pos.y) * inv.y, (bmin[2] - pos.z) * inv.z);
					V3 t2 = v3((bmax[0] - pos.x) * inv.x, (bmax[1] - pos.y) * inv.y, (bmax[2] - pos.z) * inv.z);
					float rmin = fmaxf(fminf(t1.x, t2.x), fmaxf(fminf(t1.y, t2.y), fminf(t1.z, t2.z)));
					float rmax = fminf(fmaxf(t1.x, t2.x), fminf(fmaxf(t1.y, t2.y), fmaxf(t1.z, t2.z)));
					rminOf[side] = rmin;
					hitBox[side] = rmin <= best && rmax >= rmin && rmax > 0.0f;
				}
				for (int side = 0; side < 2; ++side) {
					if (!hitBox[side] || child[side] >= 0) {
						continue;
					}
					int ti = ~child[side];
					const restir_triangle &tri = tris[ti];
					V3 p1 = v3(tri.p1[0], tri.p1[1], tri.p1[2]);
					V3 e1 = v3(tri.p2[0], tri.p2[1], tri.p2[2]) - p1;
					V3 e2 = v3(tri.p3[0], tri.p3[1], tri.p3[2]) - p1;
					if (!(dot(cross(e1, e2), dir) < 0.0f)) { // back-face culling, CCW front (pass.h:24-32)
						continue;
					}
					V3 p = cross(dir, e2);
					float fdet = 1.0f / dot(e1, p);
					V3 sv = pos - p1;
					float u_ = fdet * dot(sv, p);
					if (u_ < 0.0f || u_ > 1.0f) {
						continue;
					}
					V3 q = cross(sv, e1);
					float v_ = fdet * dot(dir, q);
					if (v_ < 0.0f || v_ + u_ > 1.0f) {
						continue;
					}
					float t = fdet * dot(e2, q);
					if (!(t >= c->zNear && t <= c->zFar) || !(t < best || (t == best && ti < bestTri))) { // clip planes; depth test LESS, earlier draw wins ties
						continue;
					}
					int mi = triMaterial[ti];
					if (mi >= 0 && mi < nMaterials && uniforms[mi].alphaMode == RESTIR_ALPHA_MODE_MASK) { // gBuffer.frag:29-34
						const float *at = attrs + (size_t)ti * 32;
						float w0 = (1.0f - u_) - v_;
						float uu = (at[3] * w0 + at[11] * u_) + at[19] * v_, vw = (at[24] * w0 + at[25] * u_) + at[26] * v_;
						float alpha = sampleTexture(tx, bindings[mi].albedo, false, uu, vw).w * uniforms[mi].colorParam[3];
						if (alpha < uniforms[mi].alphaCutoff) {
							continue;
						}
					}
					best = t;
					bestTri = ti;
					bu = u_;
					bv = v_;
				}
				const bool il = hitBox[0] && child[0] >= 0, ir = hitBox[1] && child[1] >= 0;
				if (il && ir) {
					const bool leftNear = rminOf[0] <= rminOf[1];
					if (top < 63) {
						stack[top++] = leftNear ? child[1] : child[0];
						stack[top++] = leftNear ? child[0] : child[1];
					}
				} else if (il || ir) {
					if (top < 64) {
						stack[top++] = il ? child[0] : child[1];
					}
				}
			}
			uint8_t *a = albedoOut + pix * 4;
			int16_t *nq = normalOut + pix * 4;
			uint16_t *m = materialOut + pix * 2;
			float *wp = worldPosOut + pix * 4;
			if (bestTri < 0) { // clears: gBufferPass.cpp:117-123
				a[0] = a[1] = a[2] = 0;
				a[3] = 255;
				nq[0] = nq[1] = nq[2] = 0;
				nq[3] = 32767;
				m[0] = m[1] = 0;
				wp[0] = wp[1] = wp[2] = 0.0f;
				wp[3] = 1.0f;
				depthOut[pix] = 1.0f;
				continue;
			}
			const restir_triangle &tri = tris[bestTri];
			const float w0 = (1.0f - bu) - bv, w1 = bu, w2 = bv;
			V3 hit = (v3(tri.p1[0], tri.p1[1], tri.p1[2]) * w0 + v3(tri.p2[0], tri.p2[1], tri.p2[2]) * w1) + v3(tri.p3[0], tri.p3[1], tri.p3[2]) * w2;
			const float *at = attrs + (size_t)bestTri * 32;
			V3 N = (v3(at[0], at[1], at[2]) * w0 + v3(at[8], at[9], at[10]) * w1) + v3(at[16], at[17], at[18]) * w2;
			V3 T = (v3(at[4], at[5], at[6]) * w0 + v3(at[12], at[13], at[14]) * w1) + v3(at[20], at[21], at[22]) * w2;
			float Tw = (at[7] * w0 + at[15] * w1) + at[23] * w2;
			float u = (at[3] * w0 + at[11] * w1) + at[19] * w2, v = (at[24] * w0 + at[25] * w1) + at[26] * w2;
			int mi = triMaterial[bestTri];
			restir_material_uniforms mu;
			std::memset(&mu, 0, sizeof(mu));
			restir_material_textures mt{-1, -1, -1, -1};
			if (mi >= 0 && mi < nMaterials) {
				mu = uniforms[mi];
				mt = bindings[mi];
			}
			F4o tex = sampleTexture(tx, mt.albedo, false, u, v);
			V3 albedo = v3(tex.x * mu.colorParam[0], tex.y * mu.colorParam[1], tex.z * mu.colorParam[2]);              // :29
			V3 bitangent = cross(N, T) * Tw;                                                                           // :40
			F4o nt4 = sampleTexture(tx, mt.normal, true, u * mu.normalTextureScale, v * mu.normalTextureScale);
			V3 nt = v3(nt4.x * 2.0f - 1.0f, nt4.y * 2.0f - 1.0f, nt4.z * 2.0f - 1.0f);                                 // :41
			V3 outN = normalize((T * nt.x + bitangent * nt.y) + N * nt.z);                                             // :42
			F4o mp4 = sampleTexture(tx, mt.material, false, u, v);
			F4o mp{mp4.x * mu.materialParam[0], mp4.y * mu.materialParam[1], mp4.z * mu.materialParam[2], mp4.w * mu.materialParam[3]}; // :45
			float roughness = 0.0f, metallic = 0.0f;
			V3 outAlbedo = albedo;
			if (mu.shadingModel == RESTIR_SHADING_MODEL_METALLIC_ROUGHNESS) {                                          // :48-50
				roughness = mp.y;
				metallic = mp.z;
			} else if (mu.shadingModel == RESTIR_SHADING_MODEL_SPECULAR_GLOSSINESS) {                                  // :51-67
				roughness = 1.0f - mp.w;
				V3 average = (albedo + v3(mp.x, mp.y, mp.z)) * 0.5f;
				V3 under = average * average - albedo * 0.04f;
				V3 sqrtTerm = v3(sqrtf(under.x), sqrtf(under.y), sqrtf(under.z));
				V3 metallicRgb = average * 25.0f - sqrtTerm;
				metallic = ((metallicRgb.x + metallicRgb.y) + metallicRgb.z) / 3.0f;
				outAlbedo = average + sqrtTerm;
			}
			uint8_t alphaCode = 0;
			V3 em = v3(mu.emissiveFactor[0], mu.emissiveFactor[1], mu.emissiveFactor[2]);
			if (sqrtf(dot(em, em)) > 0.0f) {                                                                           // :73-79
				F4o et = sampleTexture(tx, mt.emissive, false, u, v);
				outAlbedo = (v3(mu.colorParam[0], mu.colorParam[1], mu.colorParam[2]) * em) * v3(et.x, et.y, et.z);
				alphaCode = 255;
			}
			a[0] = srgb.code(outAlbedo.x);
			a[1] = srgb.code(outAlbedo.y);
			a[2] = srgb.code(outAlbedo.z);
			a[3] = alphaCode;
			// NaN -> 0 in the fixed-point conversions (NVIDIA / D3D behaviour; Vulkan leaves it open): a zero tangent (MikkTSpace on
			// degenerate texture coordinates) normalised in gBuffer.vert:30 poisons the normal, which is then stored as (0, 0, 0)
			nq[0] = outN.x != outN.x ? 0 : (int16_t)rintf(clampf(outN.x, -1.0f, 1.0f) * 32767.0f);
			nq[1] = outN.y != outN.y ? 0 : (int16_t)rintf(clampf(outN.y, -1.0f, 1.0f) * 32767.0f);
			nq[2] = outN.z != outN.z ? 0 : (int16_t)rintf(clampf(outN.z, -1.0f, 1.0f) * 32767.0f);
			nq[3] = 32767;
			m[0] = (uint16_t)rintf(clampf(roughness, 0.0f, 1.0f) * 65535.0f);
			m[1] = (uint16_t)rintf(clampf(metallic, 0.0f, 1.0f) * 65535.0f);
			wp[0] = hit.x;
			wp[1] = hit.y;
			wp[2] = hit.z;
			wp[3] = 1.0f;
			float cz = ((PV[2] * hit.x + PV[6] * hit.y) + PV[10] * hit.z) + PV[14];
			float cw = ((PV[3] * hit.x + PV[7] * hit.y) + PV[11] * hit.z) + PV[15];
			depthOut[pix] = cz / cw;
		}
	}
}

} // extern "C"


// ---------------------------------------------------------------------------------------------
// The reference's compile-time variants: RESERVOIR_SIZE > 1 and UNBIASED_MIS
// (include/structs/restirStructs.glsl:16-17, 26; include/reservoir.glsl under both switches;
// restirOmni.glsl:149-160, 195-205; spatialReuse.comp:73-83; unbiasedReuse.glsl:74-82, 96-121, 132-181;
// lighting.frag:53-68).  Restated for every (RESERVOIR_SIZE, UNBIASED_MIS) as templates; (1, off) is the
// configuration the rest of this file restates and must agree with it bit for bit (tests/test_variants.py);
// the other instantiations are pinned against the reference's shader sources compiled with the same defines
// (oracle/ref_build, libglslref_rs<N>[_mis].so).  Records are the reference's std430 structs:
// LightSample 48 bytes (64 with sumPHat), Reservoir = RESERVOIR_SIZE samples + numStreamSamples padded to 16.

namespace {

template <bool MIS> struct SampleV {
	float position_emissionLum[4];
	float normal[4];
	int32_t lightIndex;
	float pHat, sumWeights, w;
};
template <> struct SampleV<true> {
	float position_emissionLum[4];
	float normal[4];
	int32_t lightIndex;
	float pHat, sumWeights, w;
	float sumPHat;
	float pad_[3];
};
template <int N, bool MIS> struct ReservoirV {
	SampleV<MIS> samples[N];
	uint32_t numStreamSamples;
	uint32_t pad_[3];
};
static_assert(sizeof(ReservoirV<1, false>) == 64 && sizeof(ReservoirV<2, false>) == 112 && sizeof(ReservoirV<1, true>) == 80 &&
                  sizeof(ReservoirV<2, true>) == 144,
              "std430 sizes of Reservoir under RESERVOIR_SIZE / UNBIASED_MIS");

inline float &sumPHatOf(SampleV<true> &s) { return s.sumPHat; }
inline float sumPHatOf(const SampleV<true> &s) { return s.sumPHat; }
inline float &sumPHatOf(SampleV<false> &) {
	static thread_local float unused;
	return unused;
}
inline float sumPHatOf(const SampleV<false> &) { return 0.0f; }
template <bool MIS> inline V3 posOf(const SampleV<MIS> &s) { return v3(s.position_emissionLum[0], s.position_emissionLum[1], s.position_emissionLum[2]); }
template <bool MIS> inline V3 normalOf(const SampleV<MIS> &s) { return v3(s.normal[0], s.normal[1], s.normal[2]); }

// reservoir.glsl:6-26
template <int N, bool MIS>
void updateReservoirAtV(ReservoirV<N, MIS> &res, int i, float weight, V3 position, const float normal[4], float emissionLum, int lightIdx,
                        float pHat, float w, float sumPHat, Rand &rand) {
	SampleV<MIS> &s = res.samples[i];
	s.sumWeights = s.sumWeights + weight;
	float replacePossibility = weight / s.sumWeights;
	if (randFloat(rand) < replacePossibility) {
		s.position_emissionLum[0] = position.x;
		s.position_emissionLum[1] = position.y;
		s.position_emissionLum[2] = position.z;
		s.position_emissionLum[3] = emissionLum;
		std::memcpy(s.normal, normal, 16);
		s.lightIndex = lightIdx;
		s.pHat = pHat;
		s.w = w;
		if (MIS) {
			sumPHatOf(s) = sumPHatOf(s) + sumPHat; // :22-24
		}
	}
}
// reservoir.glsl:28-42
template <int N, bool MIS>
void addSampleToReservoirV(ReservoirV<N, MIS> &res, V3 position, const float normal[4], float emissionLum, int lightIdx, float pHat,
                           float sampleP, Rand &rand) {
	float weight = pHat / sampleP;
	res.numStreamSamples += 1;
	for (int i = 0; i < N; ++i) {
		float w = (res.samples[i].sumWeights + weight) / ((float)res.numStreamSamples * pHat);
		updateReservoirAtV(res, i, weight, position, normal, emissionLum, lightIdx, pHat, w, pHat, rand);
	}
}
// reservoir.glsl:44-64
template <int N, bool MIS> void combineReservoirsV(ReservoirV<N, MIS> &self, const ReservoirV<N, MIS> &other, const float pHat[N], Rand &rand) {
	self.numStreamSamples += other.numStreamSamples;
	for (int i = 0; i < N; ++i) {
		const SampleV<MIS> &o = other.samples[i];
		float weight = (pHat[i] * o.w) * (float)other.numStreamSamples;
		if (weight > 0.0f) {
			// reservoir.glsl:54-58: under UNBIASED_MIS the call site lists `other.sumPHat` BEFORE `other.w`, the signature (:6-11) has them
			// the other way round — the selected sample's w becomes the neighbour's sumPHat and its sumPHat grows by the neighbour's w.
			// Reproduced, not fixed (the unbiased pass, unbiasedReuse.glsl:108-121, passes them in signature order).
			if (MIS) {
				updateReservoirAtV(self, i, weight, posOf(o), o.normal, o.position_emissionLum[3], o.lightIndex, pHat[i], sumPHatOf(o), o.w, rand);
			} else {
				updateReservoirAtV(self, i, weight, posOf(o), o.normal, o.position_emissionLum[3], o.lightIndex, pHat[i], o.w, 0.0f, rand);
			}
		}
		if (self.samples[i].w > 0.0f) {
			self.samples[i].w = self.samples[i].sumWeights / ((float)self.numStreamSamples * self.samples[i].pHat);
		}
	}
}

// restirOmni.glsl:86-212
template <int N, bool MIS>
void restirPassV(const oracle_scene *s, const restir_uniforms *u, const oracle_gbuffer *cur, const oracle_gbuffer *prev, const void *prevFrameReservoirs,
                 void *reservoirsOut, int y0, int y1, uint64_t *rays) {
	typedef ReservoirV<N, MIS> R;
	const R *prevRes = (const R *)prevFrameReservoirs;
	R *reservoirs = (R *)reservoirsOut;
	Scene sc = makeScene(s->nodes, s->tris, s->pointBlob, s->triBlob, s->aliasBlob);
	GBuffer g = toG(cur), pg = toG(prev);
	const int W = (int)u->screenSize[0], H = (int)u->screenSize[1];
	const V3 camPos = v3(u->cameraPos[0], u->cameraPos[1], u->cameraPos[2]);
	const float *M = u->prevFrameProjectionViewMatrix;
	uint64_t rayCount = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : rayCount)
	for (int y = y0; y < y1; ++y) {
		for (int x = 0; x < W; ++x) {
			size_t pix = (size_t)y * W + x;
			V3 albedo = fetchAlbedo(g, pix, nullptr);
			V3 normal = fetchNormal(g, pix);
			float roughness, metallic;
			fetchMaterial(g, pix, &roughness, &metallic);
			V3 worldPos = fetchWorldPos(g, pix);
			float albedoLum = luminance(albedo.x, albedo.y, albedo.z);
			R res;
			std::memset(&res, 0, sizeof(res)); // newReservoir; fields GLSL leaves unset are zero (Appendix B.2)
			Rand rand = seedRand(u->frame, (uint32_t)((uint32_t)y * 10007u + (uint32_t)x));
			if (dot(normal, normal) != 0.0f) {
				for (uint32_t i = 0; i < u->initialLightSampleCount; ++i) {
					int selected_idx;
					float lightSampleProb;
					float r1 = randFloat(rand);
					float r2 = randFloat(rand);
					aliasTableSample(sc, r1, r2, &selected_idx, &lightSampleProb);
					V3 lightSamplePos;
					float lightNormal[4];
					float lightSampleLum;
					int lightSampleIndex;
					if (sc.pointCount != 0) {
						const restir_point_light &light = sc.pointLights[selected_idx];
						lightSamplePos = v3(light.pos[0], light.pos[1], light.pos[2]);
						lightSampleLum = light.color_luminance[3];
						lightSampleIndex = selected_idx;
						lightNormal[0] = lightNormal[1] = lightNormal[2] = lightNormal[3] = 0.0f;
					} else {
						const restir_tri_light &light = sc.triLights[selected_idx];
						float r3 = randFloat(rand);
						float r4 = randFloat(rand);
						lightSamplePos = pickPointOnTriangle(r3, r4, v3(light.p1[0], light.p1[1], light.p1[2]), v3(light.p2[0], light.p2[1], light.p2[2]),
						                                     v3(light.p3[0], light.p3[1], light.p3[2]));
						lightSampleLum = light.emission_luminance[3];
						lightSampleIndex = -1 - selected_idx;
						V3 wi = normalize(worldPos - lightSamplePos);
						V3 ln = v3(light.normalArea[0], light.normalArea[1], light.normalArea[2]);
						lightSampleProb = lightSampleProb / (fabsf(dot(wi, ln)) * light.normalArea[3]);
						lightNormal[0] = ln.x;
						lightNormal[1] = ln.y;
						lightNormal[2] = ln.z;
						lightNormal[3] = 1.0f;
					}
					float pHat = evaluatePHat(worldPos, lightSamplePos, camPos, normal, v3(lightNormal[0], lightNormal[1], lightNormal[2]),
					                          lightNormal[3] > 0.5f, albedoLum, lightSampleLum, roughness, metallic);
					addSampleToReservoirV(res, lightSamplePos, lightNormal, lightSampleLum, lightSampleIndex, pHat, lightSampleProb, rand);
				}
			}
			if ((u->flags & RESTIR_VISIBILITY_REUSE_FLAG) != 0) { // :148-160
				for (int i = 0; i < N; ++i) {
					bool shadowed = testVisibility(sc, worldPos, posOf(res.samples[i]), nullptr);
					rayCount++;
					if (shadowed) {
						res.samples[i].w = 0.0f;
						res.samples[i].sumWeights = 0.0f;
						if (MIS) {
							sumPHatOf(res.samples[i]) = 0.0f;
						}
					}
				}
			}
			if ((u->flags & RESTIR_TEMPORAL_REUSE_FLAG) != 0) { // :163-209
				float px = ((M[0] * worldPos.x + M[4] * worldPos.y) + M[8] * worldPos.z) + M[12] * 1.0f;
				float py = ((M[1] * worldPos.x + M[5] * worldPos.y) + M[9] * worldPos.z) + M[13] * 1.0f;
				float pw = ((M[3] * worldPos.x + M[7] * worldPos.y) + M[11] * worldPos.z) + M[15] * 1.0f;
				float invW = 1.0f / pw;
				px = px * invW;
				py = py * invW;
				px = ((px + 1.0f) * 0.5f) * (float)W;
				py = ((py + 1.0f) * 0.5f) * (float)H;
				if (px > 0.0f && py > 0.0f && px < (float)W && py < (float)H) {
					int fx = (int)px, fy = (int)py;
					size_t ppix = (size_t)fy * W + fx;
					V3 positionDiff = worldPos - fetchWorldPos(pg, ppix);
					if (dot(positionDiff, positionDiff) < 0.01f) {
						V3 albedoDiff = albedo - fetchAlbedo(pg, ppix, nullptr);
						if (dot(albedoDiff, albedoDiff) < 0.01f) {
							float normalDot = dot(normal, fetchNormal(pg, ppix));
							if (normalDot > 0.5f) {
								R p = prevRes[ppix];
								uint32_t cap = u->temporalSampleCountMultiplier * res.numStreamSamples;
								if (p.numStreamSamples > cap) {
									p.numStreamSamples = cap;
								}
								float pHat[N];
								for (int i = 0; i < N; ++i) {
									pHat[i] = evaluatePHat(worldPos, posOf(p.samples[i]), camPos, normal, normalOf(p.samples[i]), p.samples[i].normal[3] > 0.5f,
									                       albedoLum, p.samples[i].position_emissionLum[3], roughness, metallic);
								}
								combineReservoirsV(res, p, pHat, rand);
							}
						}
					}
				}
			}
			reservoirs[pix] = res;
		}
	}
	if (rays) {
		*rays += rayCount;
	}
}

// spatialReuse.comp:30-86
template <int N, bool MIS>
void spatialPassV(const restir_uniforms *u, const oracle_gbuffer *cur, const void *in, void *out, int iter, int y0, int y1) {
	typedef ReservoirV<N, MIS> R;
	const R *reservoirs = (const R *)in;
	R *result = (R *)out;
	GBuffer g = toG(cur);
	const int W = (int)u->screenSize[0], H = (int)u->screenSize[1];
	const V3 camPos = v3(u->cameraPos[0], u->cameraPos[1], u->cameraPos[2]);
	float sinThr, cosThr;
	det_sincos(u->spatialNormalThreshold * 0.017453292519943295f, &sinThr, &cosThr);
#pragma omp parallel for schedule(dynamic, 1)
	for (int y = y0; y < y1; ++y) {
		for (int x = 0; x < W; ++x) {
			size_t pix = (size_t)y * W + x;
			V3 albedo = fetchAlbedo(g, pix, nullptr);
			V3 normal = fetchNormal(g, pix);
			float roughness, metallic;
			fetchMaterial(g, pix, &roughness, &metallic);
			V3 worldPos = fetchWorldPos(g, pix);
			float worldDepth = g.depth[pix];
			float albedoLum = luminance(albedo.x, albedo.y, albedo.z);
			R res = reservoirs[pix];
			Rand rand = seedRand((uint32_t)(u->frame * 31u + (uint32_t)iter), (uint32_t)((uint32_t)y * 10007u + (uint32_t)x));
			for (uint32_t i = 0; i < u->spatialNeighbors; ++i) {
				float angle = (randFloat(rand) * 2.0f) * kPi;
				float radius = sqrtf(randFloat(rand)) * u->spatialRadius;
				float sn, cs;
				det_sincos(angle, &sn, &cs);
				int nx = x + (int)floorf(cs * radius), ny = y + (int)floorf(sn * radius);
				nx = nx < 0 ? 0 : (nx > W - 1 ? W - 1 : nx);
				ny = ny < 0 ? 0 : (ny > H - 1 ? H - 1 : ny);
				size_t npix = (size_t)ny * W + nx;
				float neighborDepth = g.depth[npix];
				V3 neighborNor = fetchNormal(g, npix);
				if (fabsf(neighborDepth - worldDepth) > u->spatialPosThreshold * fabsf(worldDepth) || dot(neighborNor, normal) < cosThr) {
					continue;
				}
				const R &randRes = reservoirs[npix];
				float newPHats[N]; // :73-80
				for (int j = 0; j < N; ++j) {
					newPHats[j] = evaluatePHat(worldPos, posOf(randRes.samples[j]), camPos, normal, normalOf(randRes.samples[j]),
					                           randRes.samples[j].normal[3] > 0.5f, albedoLum, randRes.samples[j].position_emissionLum[3], roughness, metallic);
				}
				combineReservoirsV(res, randRes, newPHats, rand);
			}
			result[pix] = res;
		}
	}
}

// unbiasedReuse.glsl:50-185
template <int N, bool MIS>
void unbiasedPassV(const oracle_scene *s, const restir_uniforms *u, const oracle_gbuffer *cur, const void *in, void *out, int numNeighbors, int y0,
                   int y1, uint64_t *rays) {
	typedef ReservoirV<N, MIS> R;
	const R *reservoirs = (const R *)in;
	R *result = (R *)out;
	Scene sc = makeScene(s->nodes, s->tris, nullptr, nullptr, nullptr);
	GBuffer g = toG(cur);
	const int W = (int)u->screenSize[0], H = (int)u->screenSize[1];
	const V3 camPos = v3(u->cameraPos[0], u->cameraPos[1], u->cameraPos[2]);
	const int NN = numNeighbors > 0 ? (numNeighbors > 16 ? 16 : numNeighbors) : 3;
	uint64_t rayCount = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : rayCount)
	for (int y = y0; y < y1; ++y) {
		for (int x = 0; x < W; ++x) {
			size_t pix = (size_t)y * W + x;
			V3 albedo = fetchAlbedo(g, pix, nullptr);
			V3 normal = fetchNormal(g, pix);
			float roughness, metallic;
			fetchMaterial(g, pix, &roughness, &metallic);
			V3 worldPos = fetchWorldPos(g, pix);
			float albedoLum = luminance(albedo.x, albedo.y, albedo.z);
			R res = reservoirs[pix];
			Rand rand = seedRand((uint32_t)(u->frame * 17u), (uint32_t)((uint32_t)y * 10007u + (uint32_t)x));
			size_t neighborPix[16];
			float neighborSumPHat[N][16]; // :74-82
			float originalSumPHat[N];
			uint32_t neighborNumSamples[16];
			uint32_t originalNumSamples = res.numStreamSamples;
			for (int i = 0; i < N; ++i) {
				originalSumPHat[i] = sumPHatOf(res.samples[i]);
			}
			for (int i = 0; i < NN; ++i) { // :84-124
				float angle = (randFloat(rand) * 2.0f) * kPi;
				float radius = sqrtf(randFloat(rand)) * u->spatialRadius;
				float sn, cs;
				det_sincos(angle, &sn, &cs);
				int nx = x + (int)roundf(cs * radius), ny = y + (int)roundf(sn * radius);
				nx = nx < 0 ? 0 : (nx > W - 1 ? W - 1 : nx);
				ny = ny < 0 ? 0 : (ny > H - 1 ? H - 1 : ny);
				size_t npix = (size_t)ny * W + nx;
				const R &randRes = reservoirs[npix];
				neighborPix[i] = npix;
				for (int j = 0; j < N; ++j) {
					neighborSumPHat[j][i] = sumPHatOf(randRes.samples[j]);
				}
				neighborNumSamples[i] = randRes.numStreamSamples;
				res.numStreamSamples += randRes.numStreamSamples;
				for (int j = 0; j < N; ++j) {
					const SampleV<MIS> &o = randRes.samples[j];
					float newPHat = evaluatePHat(worldPos, posOf(o), camPos, normal, normalOf(o), o.normal[3] > 0.5f, albedoLum, o.position_emissionLum[3],
					                             roughness, metallic);
					float weight = (newPHat * o.w) * (float)randRes.numStreamSamples;
					if (weight > 0.0f) {
						updateReservoirAtV(res, j, weight, posOf(o), o.normal, o.position_emissionLum[3], o.lightIndex, newPHat, o.w, sumPHatOf(o), rand);
					}
				}
			}
			V3 neighborWorldPos[16], neighborNormal[16]; // :126-130
			for (int i = 0; i < NN; ++i) {
				neighborWorldPos[i] = fetchWorldPos(g, neighborPix[i]);
				neighborNormal[i] = fetchNormal(g, neighborPix[i]);
			}
			for (int i = 0; i < N; ++i) { // :132-182
				SampleV<MIS> &smp = res.samples[i];
				V3 lightPos = posOf(smp);
				float sumPHat = originalSumPHat[i];
				uint32_t numSamples = originalNumSamples;
				for (int j = 0; j < NN; ++j) {
					if (dot(lightPos - neighborWorldPos[j], neighborNormal[j]) < 0.0f) {
						continue;
					}
					if ((u->flags & RESTIR_VISIBILITY_REUSE_FLAG) != 0) {
						bool shadowed = testVisibility(sc, neighborWorldPos[j], lightPos, nullptr);
						rayCount++;
						if (shadowed) {
							continue;
						}
					}
					sumPHat = sumPHat + neighborSumPHat[i][j];
					numSamples += neighborNumSamples[j];
				}
				if ((u->flags & RESTIR_VISIBILITY_REUSE_FLAG) != 0) {
					bool shadowed = testVisibility(sc, worldPos, lightPos, nullptr);
					rayCount++;
					if (shadowed) {
						sumPHat = 0.0f;
						numSamples = 0;
					}
				}
				if (MIS ? (sumPHat > 0.0f) : (numSamples > 0)) {
					if (MIS) {
						smp.w = (smp.sumWeights * smp.pHat) / (sumPHat * smp.pHat); // :169
					} else {
						smp.w = smp.sumWeights / ((float)numSamples * smp.pHat);
					}
				} else {
					smp.w = 0.0f;
					smp.sumWeights = 0.0f;
					if (MIS) {
						sumPHatOf(smp) = 0.0f;
					}
				}
			}
			result[pix] = res;
		}
	}
	if (rays) {
		*rays += rayCount;
	}
}

// lighting.frag:43-71,103
template <int N, bool MIS>
void lightingPassV(const oracle_scene *s, const restir_lighting_uniforms *u, const oracle_gbuffer *cur, const void *in, float *outRgba, int y0, int y1) {
	typedef ReservoirV<N, MIS> R;
	const R *reservoirs = (const R *)in;
	Scene sc = makeScene(nullptr, nullptr, s->pointBlob, s->triBlob, nullptr);
	GBuffer g = toG(cur);
	const int W = (int)u->bufferSize[0];
	const V3 camPos = v3(u->cameraPos[0], u->cameraPos[1], u->cameraPos[2]);
#pragma omp parallel for schedule(dynamic, 4)
	for (int y = y0; y < y1; ++y) {
		for (int x = 0; x < W; ++x) {
			size_t pix = (size_t)y * W + x;
			float albedoA;
			V3 albedo = fetchAlbedo(g, pix, &albedoA);
			V3 normal = fetchNormal(g, pix);
			float roughness, metallic;
			fetchMaterial(g, pix, &roughness, &metallic);
			V3 worldPos = fetchWorldPos(g, pix);
			const R &r = reservoirs[pix];
			V3 c = v3(0, 0, 0);
			for (int i = 0; i < N; ++i) { // :53-67
				const SampleV<MIS> &smp = r.samples[i];
				V3 emission = v3(0, 0, 0);
				if (smp.lightIndex < 0) {
					if (-1 - smp.lightIndex < sc.triCount) {
						const restir_tri_light &l = sc.triLights[-1 - smp.lightIndex];
						emission = v3(l.emission_luminance[0], l.emission_luminance[1], l.emission_luminance[2]);
					}
				} else if (smp.lightIndex < sc.pointCount) {
					const restir_point_light &l = sc.pointLights[smp.lightIndex];
					emission = v3(l.color_luminance[0], l.color_luminance[1], l.color_luminance[2]);
				}
				V3 pHat = evaluatePHatFull(worldPos, posOf(smp), camPos, normal, normalOf(smp), smp.normal[3] > 0.5f, albedo, emission, roughness, metallic);
				c = c + pHat * smp.w;
			}
			c = c * (1.0f / (float)N); // :68, P3
			if (albedoA > 0.5f) {
				c = albedo;
			}
			if (u->gamma != 1.0f) {
				float e = 1.0f / u->gamma;
				c = v3(powf(c.x, e), powf(c.y, e), powf(c.z, e));
			}
			float *o = outRgba + pix * 4;
			o[0] = c.x;
			o[1] = c.y;
			o[2] = c.z;
			o[3] = 1.0f;
		}
	}
}

} // namespace

#define ORACLE_VARIANT_DISPATCH(CALL)                         \
	switch (reservoirSize * 2 + (unbiasedMis ? 1 : 0)) {     \
	case 2: CALL(1, false); return 0;                        \
	case 3: CALL(1, true); return 0;                         \
	case 4: CALL(2, false); return 0;                        \
	case 5: CALL(2, true); return 0;                         \
	case 8: CALL(4, false); return 0;                        \
	case 9: CALL(4, true); return 0;                         \
	default: return -1;                                      \
	}

extern "C" {

int oracle_variant_reservoir_bytes(int reservoirSize, int unbiasedMis) { return reservoirSize * (unbiasedMis ? 64 : 48) + 16; }

int oracle_restir_pass_variant(int reservoirSize, int unbiasedMis, const oracle_scene *s, const restir_uniforms *u, const oracle_gbuffer *cur,
                               const oracle_gbuffer *prev, const void *prevFrameReservoirs, void *reservoirs, int y0, int y1, uint64_t *rays) {
#define CALL(N, M) restirPassV<N, M>(s, u, cur, prev, prevFrameReservoirs, reservoirs, y0, y1, rays)
	ORACLE_VARIANT_DISPATCH(CALL)
#undef CALL
}
int oracle_spatial_pass_variant(int reservoirSize, int unbiasedMis, const restir_uniforms *u, const oracle_gbuffer *cur, const void *in, void *out,
                                int iter, int y0, int y1) {
#define CALL(N, M) spatialPassV<N, M>(u, cur, in, out, iter, y0, y1)
	ORACLE_VARIANT_DISPATCH(CALL)
#undef CALL
}
int oracle_unbiased_pass_variant(int reservoirSize, int unbiasedMis, const oracle_scene *s, const restir_uniforms *u, const oracle_gbuffer *cur,
                                 const void *in, void *out, int numNeighbors, int y0, int y1, uint64_t *rays) {
#define CALL(N, M) unbiasedPassV<N, M>(s, u, cur, in, out, numNeighbors, y0, y1, rays)
	ORACLE_VARIANT_DISPATCH(CALL)
#undef CALL
}
int oracle_lighting_pass_variant(int reservoirSize, int unbiasedMis, const oracle_scene *s, const restir_lighting_uniforms *u, const oracle_gbuffer *cur,
                                 const void *in, float *outRgba, int y0, int y1) {
#define CALL(N, M) lightingPassV<N, M>(s, u, cur, in, outRgba, y0, y1)
	ORACLE_VARIANT_DISPATCH(CALL)
#undef CALL
}

} // extern "C"
