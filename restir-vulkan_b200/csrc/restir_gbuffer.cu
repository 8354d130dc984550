// restir_gbuffer.cu — the G-buffer pass on the GPU, hand-written for sm_100a.
//
//   vertex_stage_kernel   <- src/shaders/gBuffer.vert:22-34 (once per upload: per-triangle world-space normals, tangents, uvs)
//   gbuffer_kernel        <- src/passes/gBufferPass.cpp:116-157 (clears, draw order, depth test LESS, back-face culling
//                            src/passes/pass.h:24-40) + src/shaders/gBuffer.frag:27-80 (alpha-mask discard, TBN normal
//                            mapping, metallic-roughness / specular-glossiness conversion, emissive flag)
//
// B200 has no rasteriser: primary visibility is a closest-hit ray cast through every pixel centre against the tree the
// shadow rays use.  What the rasteriser would interpolate perspective-correctly (gBuffer.vert's outputs) is interpolated
// with the hit's barycentric coordinates, which is the same thing.  Textures are R8G8B8A8_UNORM (sceneBuffers.h:126),
// repeat wrap, bilinear at level 0 (the reference's samplers add a mip chain and 16x anisotropy, which a driver defines).
// Arithmetic policy as everywhere (restir_math.cuh, -fmad=false); the oracle twin is oracle_gbuffer_pass.
#include "restir_kernels.h"
#include "restir_trace.cuh" // ldg256: one 32-byte load

namespace restir {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ bool pixel_of_thread(const Band &b, int &x, int &y) { // same 32x8 tile of 8x4 warp tiles as restir_kernels.cu
	int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	x = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);
	y = b.rowBegin + blockIdx.y * 8 + (warp >> 2) * 4 + (lane >> 3);
	return x < b.W && y < b.rowEnd;
}

// column-major mat4 * (v, w), P4
__device__ __forceinline__ f3 mat_mul(const float *m, f3 v, float w) {
	return mk3(((m[0] * v.x + m[4] * v.y) + m[8] * v.z) + m[12] * w, ((m[1] * v.x + m[5] * v.y) + m[9] * v.z) + m[13] * w,
	           ((m[2] * v.x + m[6] * v.y) + m[10] * v.z) + m[14] * w);
}

struct F4 {
	float x, y, z, w;
};

// texture(sampler, uv): repeat, bilinear, level 0.  index < 0: the binding's default texture (a constant).
__device__ __forceinline__ F4 sample_texture(const GBufferScene &g, int index, bool normalBinding, float u, float v) {
	if (index < 0 || index >= g.nTextures) {
		if (normalBinding) {
			return F4{127.0f / 255.0f, 127.0f / 255.0f, 1.0f, 1.0f}; // sceneBuffers.h:155
		}
		return F4{1.0f, 1.0f, 1.0f, 1.0f};                             // sceneBuffers.h:164
	}
	uint4 d = __ldg(g.textureTable + index); // x = first texel, y = width, z = height
	const int W = (int)d.y, H = (int)d.z;
	float x = u * (float)W - 0.5f, y = v * (float)H - 0.5f;
	if (!(fabsf(x) < 1.0e9f) || !(fabsf(y) < 1.0e9f)) { // NaN or absurd coordinates: texel (0, 0)
		x = 0.0f;
		y = 0.0f;
	}
	float fx = floorf(x), fy = floorf(y);
	float ax = x - fx, ay = y - fy;
	int x0 = (int)fx % W, y0 = (int)fy % H;
	x0 = x0 < 0 ? x0 + W : x0;
	y0 = y0 < 0 ? y0 + H : y0;
	int x1 = x0 + 1 == W ? 0 : x0 + 1, y1 = y0 + 1 == H ? 0 : y0 + 1;
	const uchar4 *t = g.texels + d.x;
	uchar4 c00 = __ldg(t + (size_t)y0 * W + x0), c10 = __ldg(t + (size_t)y0 * W + x1);
	uchar4 c01 = __ldg(t + (size_t)y1 * W + x0), c11 = __ldg(t + (size_t)y1 * W + x1);
	float bx = 1.0f - ax, by = 1.0f - ay;
#define RESTIR_BILERP(ch) \
	((div_unorm8((float)c00.ch) * bx + div_unorm8((float)c10.ch) * ax) * by + (div_unorm8((float)c01.ch) * bx + div_unorm8((float)c11.ch) * ax) * ay)
	F4 r{RESTIR_BILERP(x), RESTIR_BILERP(y), RESTIR_BILERP(z), RESTIR_BILERP(w)};
#undef RESTIR_BILERP
	return r;
}

// float -> the 8-bit code an R8G8B8A8_SRGB attachment stores: the largest code whose lower threshold the value reaches
// (thresholds = EOTF of the code midpoints, computed in double on the host; NaN -> 0)
__device__ __forceinline__ unsigned srgb8_code(const float *__restrict__ thr, float c) {
	unsigned lo = 0, hi = 255;
#pragma unroll
	for (int i = 0; i < 8; ++i) {
		unsigned mid = (lo + hi + 1) >> 1;
		if (c >= __ldg(thr + mid)) {
			lo = mid;
		} else {
			hi = mid - 1;
		}
	}
	return lo;
}
// float -> SNORM16 / UNORM16 as the attachment stores it: clamped, round to nearest even, NaN -> 0 (what NVIDIA hardware and
// D3D define; Vulkan leaves it open).  A NaN normal — a zero tangent from MikkTSpace on degenerate texture coordinates,
// normalised in gBuffer.vert:30 — is therefore stored as (0, 0, 0): that pixel is background to the ReSTIR passes.
__device__ __forceinline__ short snorm16(float v) { return v != v ? (short)0 : (short)rintf(fminf(fmaxf(v, -1.0f), 1.0f) * 32767.0f); }
__device__ __forceinline__ unsigned short unorm16(float v) { return (unsigned short)rintf(fminf(fmaxf(v, 0.0f), 1.0f) * 65535.0f); }

struct TriAttr {
	f3 n[3];
	F4 t[3];
	float u[3], v[3];
};
__device__ __forceinline__ TriAttr load_attr(const float4 *__restrict__ attrs, int tri) {
	const float4 *a = attrs + (size_t)tri * 8;
	TriAttr r;
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		float4 n = __ldg(a + 2 * k), t = __ldg(a + 2 * k + 1);
		r.n[k] = mk3(n.x, n.y, n.z);
		r.u[k] = n.w;
		r.t[k] = F4{t.x, t.y, t.z, t.w};
	}
	float4 vv = __ldg(a + 6);
	r.v[0] = vv.x;
	r.v[1] = vv.y;
	r.v[2] = vv.z;
	return r;
}

} // namespace

// gBuffer.vert:22-34 per triangle corner, in draw order (= the order AabbTree::build collects triangles in).
// Record (8 x float4): (N0, u0) (T0) (N1, u1) (T1) (N2, u2) (T2) (v0, v1, v2, -) (-).
__global__ void vertex_stage_kernel(const restir_vertex *__restrict__ vertices, const uint32_t *__restrict__ indices,
                                    const restir_draw *__restrict__ draws, const restir_model_matrices *__restrict__ matrices,
                                    const uint32_t *__restrict__ triDraw, const uint32_t *__restrict__ drawFirstTri, uint32_t nTris,
                                    float4 *__restrict__ attrs) {
	uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= nTris) {
		return;
	}
	const uint32_t d = triDraw[t];
	const restir_draw draw = draws[d];
	const float *M = matrices[d].transform, *MIT = matrices[d].transformInverseTransposed;
	const uint32_t *idx = indices + draw.firstIndex + (size_t)(t - drawFirstTri[d]) * 3;
	float4 *o = attrs + (size_t)t * 8;
	float vs[3];
	for (int k = 0; k < 3; ++k) {
		const restir_vertex &vx = vertices[(size_t)draw.vertexOffset + idx[k]];
		f3 n = normalize3(mat_mul(MIT, mk3(vx.normal[0], vx.normal[1], vx.normal[2]), 0.0f));     // :29
		f3 tg = normalize3(mat_mul(M, mk3(vx.tangent[0], vx.tangent[1], vx.tangent[2]), 0.0f));    // :30
		o[2 * k] = make_float4(n.x, n.y, n.z, vx.uv[0]);
		o[2 * k + 1] = make_float4(tg.x, tg.y, tg.z, vx.tangent[3]);                              // :31
		vs[k] = vx.uv[1];
	}
	o[6] = make_float4(vs[0], vs[1], vs[2], 0.0f);
	o[7] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
}

template <bool IMAGE>
__global__ void __launch_bounds__(kThreads) gbuffer_kernel(SceneView sc, GBufferScene g, Band band, RaycastCamera cam, float zNear, float zFar,
                                                          uchar4 *albedoOut, short4 *normalOut, ushort2 *materialOut, float4 *worldPosOut,
                                                          float *depthOut) {
	int x, y;
	if (!pixel_of_thread(band, x, y)) {
		return;
	}
	size_t pix = (size_t)(y - band.allocBegin) * (size_t)band.W + (size_t)x;
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
	// Closest hit = the depth test LESS over all fragments of the pixel.  The walk visits the nearer child first (slab entry
	// parameter; ties: left first) so that `best` shrinks early and the farther subtrees are pruned by `rmin <= best`; the oracle
	// twin walks in the same order (the pruning makes the result order-dependent in the last bit of near-ties).
	// One candidate triangle: back-face culling, Moller-Trumbore, the near / far planes, the depth test with its tie rule, the mask.
	auto consider = [&](int ti, f3 p1, f3 e1, f3 e2) {
		if (!(dot3(cross3(e1, e2), dir) < 0.0f)) { // back-face culling, CCW front (pass.h:24-32)
			return;
		}
		f3 pv = cross3(dir, e2);
		float fdet = 1.0f / dot3(e1, pv);
		f3 sv = pos - p1;
		float u_ = fdet * dot3(sv, pv);
		if (u_ < 0.0f || u_ > 1.0f) {
			return;
		}
		f3 q = cross3(sv, e1);
		float v_ = fdet * dot3(dir, q);
		if (v_ < 0.0f || v_ + u_ > 1.0f) {
			return;
		}
		float tt = fdet * dot3(e2, q);
		// the view-space depth of the hit is tt (dir . fwd = 1): near and far planes clip like the rasteriser's
		if (!(tt >= zNear && tt <= zFar) || !(tt < best || (tt == best && ti < bestTri))) {
			return;
		}
		int mi = __ldg(g.triMaterial + ti);
		if ((unsigned)mi < (unsigned)g.nMaterials && __ldg(&g.uniforms[mi].alphaMode) == RESTIR_ALPHA_MODE_MASK) { // gBuffer.frag:29-34
			const float4 *at = g.attrs + (size_t)ti * 8;
			float4 n0 = __ldg(at), n1 = __ldg(at + 2), n2 = __ldg(at + 4), vv = __ldg(at + 6);
			float w0 = (1.0f - u_) - v_;
			float uu = (n0.w * w0 + n1.w * u_) + n2.w * v_, vw = (vv.x * w0 + vv.y * u_) + vv.z * v_;
			float alpha = sample_texture(g, __ldg(&g.bindings[mi].albedo), false, uu, vw).w * __ldg(&g.uniforms[mi].colorParam[3]);
			if (alpha < __ldg(&g.uniforms[mi].alphaCutoff)) {
				return;
			}
		}
		best = tt;
		bestTri = ti;
		bu = u_;
		bv = v_;
	};
	int stack[64];
	int top = 1;
	stack[0] = 0;
	if (IMAGE) {
		// The same walk — same boxes, same children, same order, same operations — over the 64-byte image of the tree and the 64-byte
		// (p1, e1, e2) triangle records the shadow rays use (traversal_image.h, restir_trace.cuh): two 32-byte loads per node instead of
		// five 16-byte ones from an 80-byte stride, both boxes per packed FADD2 / FMUL2 (each half an IEEE operation: b + (-o) IS
		// b - o), one 32-byte + one 4-byte load per triangle with the edges already subtracted.  With the 4-wide image in use the
		// leaves of the binary image name triangle RECORDS; recordTri gives the triangle back (attributes, material, tie rule).
		const float2 nox = make_float2(-pos.x, -pos.x), noy = make_float2(-pos.y, -pos.y), noz = make_float2(-pos.z, -pos.z);
		const float2 ivx = make_float2(inv.x, inv.x), ivy = make_float2(inv.y, inv.y), ivz = make_float2(inv.z, inv.z);
		while (top > 0) {
			const float4 *n = sc.image + (size_t)(unsigned)stack[--top] * 4u;
			const F8 lo = ldg256(n), hi = ldg256(n + 2); // (Lmin, Rmin, Lmax, Rmax) for x, for y | for z, (left, right, -, -)
			const int child[2] = {__float_as_int(hi.v[4]), __float_as_int(hi.v[5])};
			const float2 t1x = __fmul2_rn(__fadd2_rn(make_float2(lo.v[0], lo.v[1]), nox), ivx), t2x = __fmul2_rn(__fadd2_rn(make_float2(lo.v[2], lo.v[3]), nox), ivx);
			const float2 t1y = __fmul2_rn(__fadd2_rn(make_float2(lo.v[4], lo.v[5]), noy), ivy), t2y = __fmul2_rn(__fadd2_rn(make_float2(lo.v[6], lo.v[7]), noy), ivy);
			const float2 t1z = __fmul2_rn(__fadd2_rn(make_float2(hi.v[0], hi.v[1]), noz), ivz), t2z = __fmul2_rn(__fadd2_rn(make_float2(hi.v[2], hi.v[3]), noz), ivz);
			float rminOf[2];
			bool hitBox[2];
			{
				const float rmin = fmaxf(fminf(t1x.x, t2x.x), fmaxf(fminf(t1y.x, t2y.x), fminf(t1z.x, t2z.x)));
				const float rmax = fminf(fmaxf(t1x.x, t2x.x), fminf(fmaxf(t1y.x, t2y.x), fmaxf(t1z.x, t2z.x)));
				rminOf[0] = rmin;
				hitBox[0] = rmin <= best && rmax >= rmin && rmax > 0.0f;
			}
			{
				const float rmin = fmaxf(fminf(t1x.y, t2x.y), fmaxf(fminf(t1y.y, t2y.y), fminf(t1z.y, t2z.y)));
				const float rmax = fminf(fmaxf(t1x.y, t2x.y), fminf(fmaxf(t1y.y, t2y.y), fmaxf(t1z.y, t2z.y)));
				rminOf[1] = rmin;
				hitBox[1] = rmin <= best && rmax >= rmin && rmax > 0.0f;
			}
#pragma unroll
			for (int side = 0; side < 2; ++side) {
				if (!hitBox[side] || child[side] >= 0) {
					continue;
				}
				const int rec = ~child[side];
				const float4 *t = sc.triEdges + (size_t)rec * 4;
				const F8 a = ldg256(t);
				const float e2z = __ldg(reinterpret_cast<const float *>(t + 2));
				const int ti = g.recordTri != nullptr ? (int)__ldg(g.recordTri + rec) : rec;
				consider(ti, mk3(a.v[0], a.v[1], a.v[2]), mk3(a.v[3], a.v[4], a.v[5]), mk3(a.v[6], a.v[7], e2z));
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
	} else {
		while (top > 0) {
			const float4 *n = sc.nodes + (size_t)stack[--top] * 5;
			float4 ch = __ldg(n + 4);
			int child[2] = {__float_as_int(ch.x), __float_as_int(ch.y)};
			float rminOf[2];
			bool hitBox[2];
#pragma unroll
			for (int side = 0; side < 2; ++side) {
				float4 bmin = __ldg(n + side * 2), bmax = __ldg(n + side * 2 + 1);
				float t1x = (bmin.x - pos.x) * inv.x, t1y = (bmin.y - pos.y) * inv.y, t1z = (bmin.z - pos.z) * inv.z;
				float t2x = (bmax.x - pos.x) * inv.x, t2y = (bmax.y - pos.y) * inv.y, t2z = (bmax.z - pos.z) * inv.z;
				float rmin = fmaxf(fminf(t1x, t2x), fmaxf(fminf(t1y, t2y), fminf(t1z, t2z)));
				float rmax = fminf(fmaxf(t1x, t2x), fminf(fmaxf(t1y, t2y), fmaxf(t1z, t2z)));
				rminOf[side] = rmin;
				hitBox[side] = rmin <= best && rmax >= rmin && rmax > 0.0f;
			}
			// leaves first (left, then right), then the inner children: the farther one is pushed first, so the nearer one is popped first
#pragma unroll
			for (int side = 0; side < 2; ++side) {
				if (!hitBox[side] || child[side] >= 0) {
					continue;
				}
				int ti = ~child[side];
				const float4 *t = sc.tris + (size_t)ti * 3;
				float4 a = __ldg(t), b = __ldg(t + 1), c = __ldg(t + 2);
				f3 p1 = mk3(a.x, a.y, a.z);
				consider(ti, p1, mk3(b.x, b.y, b.z) - p1, mk3(c.x, c.y, c.z) - p1);
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
	}
	if (bestTri < 0) { // clears, gBufferPass.cpp:117-123
		albedoOut[pix] = make_uchar4(0, 0, 0, 255);
		normalOut[pix] = make_short4(0, 0, 0, 32767);
		materialOut[pix] = make_ushort2(0, 0);
		worldPosOut[pix] = make_float4(0.0f, 0.0f, 0.0f, 1.0f);
		depthOut[pix] = 1.0f;
		return;
	}
	// ---- what the rasteriser interpolates (gBuffer.vert's outputs), at the hit ----
	const float4 *t = sc.tris + (size_t)bestTri * 3;
	float4 a = __ldg(t), b = __ldg(t + 1), c = __ldg(t + 2);
	const float w0 = (1.0f - bu) - bv, w1 = bu, w2 = bv;
	f3 hit = (mk3(a.x, a.y, a.z) * w0 + mk3(b.x, b.y, b.z) * w1) + mk3(c.x, c.y, c.z) * w2;
	TriAttr at = load_attr(g.attrs, bestTri);
	f3 N = (at.n[0] * w0 + at.n[1] * w1) + at.n[2] * w2;
	f3 T = (mk3(at.t[0].x, at.t[0].y, at.t[0].z) * w0 + mk3(at.t[1].x, at.t[1].y, at.t[1].z) * w1) + mk3(at.t[2].x, at.t[2].y, at.t[2].z) * w2;
	float Tw = (at.t[0].w * w0 + at.t[1].w * w1) + at.t[2].w * w2;
	float u = (at.u[0] * w0 + at.u[1] * w1) + at.u[2] * w2, v = (at.v[0] * w0 + at.v[1] * w1) + at.v[2] * w2;
	// ---- gBuffer.frag:27-80 ----
	int mi = __ldg(g.triMaterial + bestTri);
	restir_material_uniforms mu{};
	restir_material_textures mt{-1, -1, -1, -1};
	if ((unsigned)mi < (unsigned)g.nMaterials) {
		mu = g.uniforms[mi];
		mt = g.bindings[mi];
	}
	F4 tex = sample_texture(g, mt.albedo, false, u, v);
	f3 albedo = mk3(tex.x * mu.colorParam[0], tex.y * mu.colorParam[1], tex.z * mu.colorParam[2]);                      // :29
	f3 bitangent = cross3(N, T) * Tw;                                                                                   // :40
	F4 nt4 = sample_texture(g, mt.normal, true, u * mu.normalTextureScale, v * mu.normalTextureScale);
	f3 nt = mk3(nt4.x * 2.0f - 1.0f, nt4.y * 2.0f - 1.0f, nt4.z * 2.0f - 1.0f);                                         // :41
	f3 outN = normalize3((T * nt.x + bitangent * nt.y) + N * nt.z);                                                     // :42
	F4 mp4 = sample_texture(g, mt.material, false, u, v);
	F4 mp{mp4.x * mu.materialParam[0], mp4.y * mu.materialParam[1], mp4.z * mu.materialParam[2], mp4.w * mu.materialParam[3]}; // :45
	float roughness = 0.0f, metallic = 0.0f;
	f3 outAlbedo = albedo;
	if (mu.shadingModel == RESTIR_SHADING_MODEL_METALLIC_ROUGHNESS) {                                                   // :48-50
		roughness = mp.y;
		metallic = mp.z;
	} else if (mu.shadingModel == RESTIR_SHADING_MODEL_SPECULAR_GLOSSINESS) {                                           // :51-67
		roughness = 1.0f - mp.w;
		f3 average = (albedo + mk3(mp.x, mp.y, mp.z)) * 0.5f;
		f3 under = average * average - albedo * 0.04f;
		f3 sqrtTerm = mk3(sqrtf(under.x), sqrtf(under.y), sqrtf(under.z));
		f3 metallicRgb = average * 25.0f - sqrtTerm;
		metallic = ((metallicRgb.x + metallicRgb.y) + metallicRgb.z) / 3.0f;
		outAlbedo = average + sqrtTerm;
	}
	unsigned alphaCode = 0;
	f3 em = mk3(mu.emissiveFactor[0], mu.emissiveFactor[1], mu.emissiveFactor[2]);
	if (sqrtf(dot3(em, em)) > 0.0f) {                                                                                   // :73-79
		F4 et = sample_texture(g, mt.emissive, false, u, v);
		outAlbedo = (mk3(mu.colorParam[0], mu.colorParam[1], mu.colorParam[2]) * em) * mk3(et.x, et.y, et.z);
		alphaCode = 255;
	}
	albedoOut[pix] = make_uchar4(srgb8_code(g.srgbThresholds, outAlbedo.x), srgb8_code(g.srgbThresholds, outAlbedo.y),
	                             srgb8_code(g.srgbThresholds, outAlbedo.z), alphaCode);
	normalOut[pix] = make_short4(snorm16(outN.x), snorm16(outN.y), snorm16(outN.z), 32767);
	materialOut[pix] = make_ushort2(unorm16(roughness), unorm16(metallic));
	worldPosOut[pix] = make_float4(hit.x, hit.y, hit.z, 1.0f);
	const float *PV = cam.pv;
	float cz = ((PV[2] * hit.x + PV[6] * hit.y) + PV[10] * hit.z) + PV[14];
	float cw = ((PV[3] * hit.x + PV[7] * hit.y) + PV[11] * hit.z) + PV[15];
	depthOut[pix] = cz / cw;
}

void launch_vertex_stage(const restir_vertex *vertices, const uint32_t *indices, const restir_draw *draws, const restir_model_matrices *matrices,
                         const uint32_t *triDraw, const uint32_t *drawFirstTri, uint32_t nTris, float4 *attrs, cudaStream_t s) {
	if (nTris) {
		vertex_stage_kernel<<<(nTris + 255) / 256, 256, 0, s>>>(vertices, indices, draws, matrices, triDraw, drawFirstTri, nTris, attrs);
	}
}

void launch_gbuffer(const SceneView &sc, const GBufferScene &g, const Band &band, const RaycastCamera &cam, float zNear, float zFar, void *albedo,
                    void *normal, void *material, void *worldPos, void *depth, cudaStream_t s) {
	dim3 grid((unsigned)((band.W + 31) / 32), (unsigned)((band.rowEnd - band.rowBegin + 7) / 8), 1);
	if (sc.image != nullptr && sc.triEdges != nullptr) {
		gbuffer_kernel<true><<<grid, kThreads, 0, s>>>(sc, g, band, cam, zNear, zFar, (uchar4 *)albedo, (short4 *)normal, (ushort2 *)material,
		                                               (float4 *)worldPos, (float *)depth);
	} else {
		gbuffer_kernel<false><<<grid, kThreads, 0, s>>>(sc, g, band, cam, zNear, zFar, (uchar4 *)albedo, (short4 *)normal, (ushort2 *)material,
		                                                (float4 *)worldPos, (float *)depth);
	}
}

cudaError_t preload_gbuffer_kernels() {
	cudaFuncAttributes a;
	cudaError_t e = cudaFuncGetAttributes(&a, vertex_stage_kernel);
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, gbuffer_kernel<true>);
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, gbuffer_kernel<false>);
	return e;
}

} // namespace restir
