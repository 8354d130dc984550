// glsl_shim.h — the GLSL types and built-in functions the reference's shaders use, as C++ (TEST INFRASTRUCTURE).
//
// Lets g++ compile the reference's own shader sources (transliterated by glsl2cpp.py, arithmetic untouched)
// into oracle/_ref/libglslref.so, which pins oracle/restir_oracle.cpp: the restatement must reproduce, bit for
// bit, what the authors' source text computes.
//
// GLSL leaves the precision of '/', sqrt, pow, sin, cos, normalize and FMA contraction to the driver; the
// built-ins below implement the one IEEE-754 binary32 reading DESIGN.md §Arithmetic policy fixes (P1-P12) — the
// same reading the oracle and the CUDA kernels implement, written a third time here, independently, so that the
// comparison checks the *shader logic* (which operations, on which operands, in which order, under which
// conditions) against the reference's text.  Build with -ffp-contract=off, no -ffast-math, and
// -ftrivial-auto-var-init=zero (P-definition: fields GLSL leaves uninitialised are zero).
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#undef M_PI

namespace glsl {

typedef unsigned int uint;

struct vec2;
struct vec3;
struct vec4;

// ---- swizzles: views of N floats that read/write a subset as a vec2 / vec3 ------------------------------------
template <int N, int A, int B> struct Swz2 {
	float v[N];
	operator vec2() const;
	Swz2 &operator=(const vec2 &o);
	Swz2 &operator=(const Swz2 &o);
	Swz2 &operator+=(const vec2 &o);
	Swz2 &operator-=(const vec2 &o);
	Swz2 &operator*=(float s);
	Swz2 &operator/=(float s);
};
template <int N, int A, int B, int C> struct Swz3 {
	float v[N];
	operator vec3() const;
	Swz3 &operator=(const vec3 &o);
	Swz3 &operator=(const Swz3 &o);
	Swz3 &operator+=(const vec3 &o);
	Swz3 &operator-=(const vec3 &o);
	Swz3 &operator*=(float s);
	Swz3 &operator/=(float s);
};

struct vec2 {
	union {
		struct { float x, y; };
		struct { float r, g; };
		float v[2];
	};
	vec2() = default;
	explicit vec2(float s) : x(s), y(s) {}
	vec2(float a, float b) : x(a), y(b) {}
	explicit vec2(const struct uvec2 &u);
	explicit vec2(const struct ivec2 &u);
};

struct vec3 {
	union {
		struct { float x, y, z; };
		struct { float r, g, b; };
		float v[3];
		Swz2<3, 0, 1> xy;
	};
	vec3() = default;
	vec3(const vec3 &o) : x(o.x), y(o.y), z(o.z) {}
	vec3 &operator=(const vec3 &o) { x = o.x; y = o.y; z = o.z; return *this; }
	explicit vec3(float s) : x(s), y(s), z(s) {}
	vec3(float a, float b_, float c) : x(a), y(b_), z(c) {}
	vec3(const vec2 &a, float c) : x(a.x), y(a.y), z(c) {}
};

struct alignas(16) vec4 {
	union {
		struct { float x, y, z, w; };
		struct { float r, g, b, a; };
		float v[4];
		Swz2<4, 0, 1> xy;
		Swz3<4, 0, 1, 2> xyz;
		Swz3<4, 0, 1, 2> rgb;
	};
	vec4() = default;
	vec4(const vec4 &o) : x(o.x), y(o.y), z(o.z), w(o.w) {}
	vec4 &operator=(const vec4 &o) { x = o.x; y = o.y; z = o.z; w = o.w; return *this; }
	explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
	vec4(float a_, float b_, float c, float d) : x(a_), y(b_), z(c), w(d) {}
	vec4(const vec3 &a_, float d) : x(a_.x), y(a_.y), z(a_.z), w(d) {}
};

struct ivec2 {
	int x, y;
	ivec2() = default;
	explicit ivec2(int s) : x(s), y(s) {}
	ivec2(int a, int b) : x(a), y(b) {}
	explicit ivec2(const vec2 &f) : x((int)f.x), y((int)f.y) {} // P9: float -> int truncates
	explicit ivec2(const struct uvec2 &u);
};
struct uvec2 {
	uint x, y;
	uvec2() = default;
	explicit uvec2(uint s) : x(s), y(s) {}
	uvec2(uint a, uint b) : x(a), y(b) {}
	explicit uvec2(const vec2 &f) : x((uint)f.x), y((uint)f.y) {}
};
struct uvec3 { // gl_GlobalInvocationID: only .xy is used
	uvec2 xy;
	uint z;
};
struct fragcoord { // gl_FragCoord: only .xy is used
	vec2 xy;
};
struct bvec2 {
	bool x, y;
};
inline vec2::vec2(const uvec2 &u) : x((float)u.x), y((float)u.y) {} // P10: uint -> float rounds to nearest even
inline vec2::vec2(const ivec2 &u) : x((float)u.x), y((float)u.y) {}
inline ivec2::ivec2(const uvec2 &u) : x((int)u.x), y((int)u.y) {}

struct mat4 { // column-major, like the uniform block
	float m[16];
};

// ---- swizzle members ------------------------------------------------------------------------------------------
template <int N, int A, int B> Swz2<N, A, B>::operator vec2() const { return vec2(v[A], v[B]); }
template <int N, int A, int B> Swz2<N, A, B> &Swz2<N, A, B>::operator=(const vec2 &o) { v[A] = o.x; v[B] = o.y; return *this; }
template <int N, int A, int B> Swz2<N, A, B> &Swz2<N, A, B>::operator=(const Swz2 &o) { return *this = vec2(o); }
template <int N, int A, int B, int C> Swz3<N, A, B, C>::operator vec3() const { return vec3(v[A], v[B], v[C]); }
template <int N, int A, int B, int C> Swz3<N, A, B, C> &Swz3<N, A, B, C>::operator=(const vec3 &o) { v[A] = o.x; v[B] = o.y; v[C] = o.z; return *this; }
template <int N, int A, int B, int C> Swz3<N, A, B, C> &Swz3<N, A, B, C>::operator=(const Swz3 &o) { return *this = vec3(o); }

// ---- arithmetic: component-wise, every operation an IEEE binary32 operation (P1) -------------------------------
// P3: vector / scalar and vector / vector multiply by the reciprocal(s); scalar / scalar is a true division.
#define GLSL_VEC_OPS(V, EXPR2, EXPRS, EXPRSL)                                                                          \
	inline V operator+(const V &a, const V &b) { return EXPR2(+); }                                                    \
	inline V operator-(const V &a, const V &b) { return EXPR2(-); }                                                    \
	inline V operator*(const V &a, const V &b) { return EXPR2(*); }                                                    \
	inline V operator+(const V &a, float s) { return EXPRS(+); }                                                       \
	inline V operator-(const V &a, float s) { return EXPRS(-); }                                                       \
	inline V operator*(const V &a, float s) { return EXPRS(*); }                                                       \
	inline V operator+(float s, const V &a) { return EXPRSL(+); }                                                      \
	inline V operator-(float s, const V &a) { return EXPRSL(-); }                                                      \
	inline V operator*(float s, const V &a) { return EXPRSL(*); }                                                      \
	inline V operator/(const V &a, float d) { float s = 1.0f / d; return EXPRS(*); }                                   \
	inline V &operator+=(V &a, const V &b) { a = a + b; return a; }                                                    \
	inline V &operator-=(V &a, const V &b) { a = a - b; return a; }                                                    \
	inline V &operator*=(V &a, const V &b) { a = a * b; return a; }                                                    \
	inline V &operator*=(V &a, float s) { a = a * s; return a; }                                                       \
	inline V &operator/=(V &a, float s) { a = a / s; return a; }

#define E2_2(op) vec2(a.x op b.x, a.y op b.y)
#define ES_2(op) vec2(a.x op s, a.y op s)
#define ESL_2(op) vec2(s op a.x, s op a.y)
#define E2_3(op) vec3(a.x op b.x, a.y op b.y, a.z op b.z)
#define ES_3(op) vec3(a.x op s, a.y op s, a.z op s)
#define ESL_3(op) vec3(s op a.x, s op a.y, s op a.z)
#define E2_4(op) vec4(a.x op b.x, a.y op b.y, a.z op b.z, a.w op b.w)
#define ES_4(op) vec4(a.x op s, a.y op s, a.z op s, a.w op s)
#define ESL_4(op) vec4(s op a.x, s op a.y, s op a.z, s op a.w)
GLSL_VEC_OPS(vec2, E2_2, ES_2, ESL_2)
GLSL_VEC_OPS(vec3, E2_3, ES_3, ESL_3)
GLSL_VEC_OPS(vec4, E2_4, ES_4, ESL_4)
inline vec3 operator/(const vec3 &a, const vec3 &b) { return vec3(a.x * (1.0f / b.x), a.y * (1.0f / b.y), a.z * (1.0f / b.z)); } // P3
inline vec2 operator/(const vec2 &a, const vec2 &b) { return vec2(a.x * (1.0f / b.x), a.y * (1.0f / b.y)); }
inline vec3 operator-(const vec3 &a) { return vec3(-a.x, -a.y, -a.z); }

template <int N, int A, int B> Swz2<N, A, B> &Swz2<N, A, B>::operator+=(const vec2 &o) { return *this = vec2(*this) + o; }
template <int N, int A, int B> Swz2<N, A, B> &Swz2<N, A, B>::operator-=(const vec2 &o) { return *this = vec2(*this) - o; }
template <int N, int A, int B> Swz2<N, A, B> &Swz2<N, A, B>::operator*=(float s) { return *this = vec2(*this) * s; }
template <int N, int A, int B> Swz2<N, A, B> &Swz2<N, A, B>::operator/=(float s) { return *this = vec2(*this) / s; }
template <int N, int A, int B, int C> Swz3<N, A, B, C> &Swz3<N, A, B, C>::operator+=(const vec3 &o) { return *this = vec3(*this) + o; }
template <int N, int A, int B, int C> Swz3<N, A, B, C> &Swz3<N, A, B, C>::operator-=(const vec3 &o) { return *this = vec3(*this) - o; }
template <int N, int A, int B, int C> Swz3<N, A, B, C> &Swz3<N, A, B, C>::operator*=(float s) { return *this = vec3(*this) * s; }
template <int N, int A, int B, int C> Swz3<N, A, B, C> &Swz3<N, A, B, C>::operator/=(float s) { return *this = vec3(*this) / s; }

inline ivec2 operator+(const ivec2 &a, const ivec2 &b) { return ivec2(a.x + b.x, a.y + b.y); }
inline ivec2 operator-(const ivec2 &a, const ivec2 &b) { return ivec2(a.x - b.x, a.y - b.y); }
inline uvec2 operator-(const uvec2 &a, uint s) { return uvec2(a.x - s, a.y - s); }
inline uvec2 operator+(const uvec2 &a, uint s) { return uvec2(a.x + s, a.y + s); }

// mat4 * vec4, P4: ((c0*x + c1*y) + c2*z) + c3*w
inline vec4 operator*(const mat4 &M, const vec4 &p) {
	vec4 r;
	for (int i = 0; i < 4; ++i) {
		r.v[i] = ((M.m[i] * p.x + M.m[4 + i] * p.y) + M.m[8 + i] * p.z) + M.m[12 + i] * p.w;
	}
	return r;
}

// ---- built-ins ------------------------------------------------------------------------------------------------
inline float sqrt(float x) { return ::sqrtf(x); }                     // P1
inline float abs(float x) { return ::fabsf(x); }
inline float floor(float x) { return ::floorf(x); }                   // P9
inline float round(float x) { return ::roundf(x); }                   // P9: half away from zero
inline vec2 round(const vec2 &a) { return vec2(::roundf(a.x), ::roundf(a.y)); }
inline float min(float a, float b) { return ::fminf(a, b); }          // P6
inline float max(float a, float b) { return ::fmaxf(a, b); }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline uint min(uint a, uint b) { return a < b ? a : b; }
inline uint max(uint a, uint b) { return a > b ? a : b; }
inline vec3 min(const vec3 &a, const vec3 &b) { return vec3(::fminf(a.x, b.x), ::fminf(a.y, b.y), ::fminf(a.z, b.z)); }
inline vec3 max(const vec3 &a, const vec3 &b) { return vec3(::fmaxf(a.x, b.x), ::fmaxf(a.y, b.y), ::fmaxf(a.z, b.z)); }
inline float clamp(float x, float lo, float hi) { return ::fminf(::fmaxf(x, lo), hi); } // P6
inline int clamp(int x, int lo, int hi) { return min(max(x, lo), hi); }
inline ivec2 clamp(const ivec2 &a, const ivec2 &lo, const ivec2 &hi) { return ivec2(clamp(a.x, lo.x, hi.x), clamp(a.y, lo.y, hi.y)); }
inline float mix(float x, float y, float a) { return x * (1.0f - a) + y * a; } // P6
inline vec3 mix(const vec3 &x, const vec3 &y, float a) { return vec3(mix(x.x, y.x, a), mix(x.y, y.y, a), mix(x.z, y.z, a)); }
inline float dot(const vec3 &a, const vec3 &b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; } // P4
inline float dot(const vec2 &a, const vec2 &b) { return a.x * b.x + a.y * b.y; }
inline vec3 cross(const vec3 &a, const vec3 &b) { // the GLSL specification's expression
	return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
inline vec3 normalize(const vec3 &a) { return a * (1.0f / ::sqrtf(dot(a, a))); } // P2
inline float length(const vec3 &a) { return ::sqrtf(dot(a, a)); }
inline float radians(float d) { return d * 0.017453292519943295f; }   // P8
// P5 / P11: pow(x, 2) = x * x, pow(x, 1) = x; any other exponent is the C library's powf (lighting output only)
inline float pow(float x, float y) { return y == 2.0f ? x * x : (y == 1.0f ? x : ::powf(x, y)); }
inline vec3 pow(const vec3 &a, const vec3 &e) { return vec3(pow(a.x, e.x), pow(a.y, e.y), pow(a.z, e.z)); }

// P7: k = rint(a * 2/pi); r = a - k * pi/2 in three Cody-Waite steps; Cephes single-precision sine and cosine
// polynomials on r, every step rounded; quadrant fix-up.
inline void sincos_p7(float a, float &s, float &c) {
	float kf = ::rintf(a * 0.636619772f);
	int k = (int)kf;
	float r = a - kf * 1.5703125f;
	r = r - kf * 4.837512969970703125e-4f;
	r = r - kf * 7.54978995489188216e-8f;
	float z = r * r;
	float sr = ((((-1.9515295891e-4f * z + 8.3321608736e-3f) * z) - 1.6666654611e-1f) * z) * r + r;
	float cr = ((((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z) * z - 0.5f * z) + 1.0f;
	switch (k & 3) {
	case 0: s = sr; c = cr; break;
	case 1: s = cr; c = -sr; break;
	case 2: s = -sr; c = -cr; break;
	default: s = -cr; c = sr; break;
	}
}
inline float sin(float a) { float s, c; sincos_p7(a, s, c); return s; }
inline float cos(float a) { float s, c; sincos_p7(a, s, c); return c; }

inline bvec2 greaterThanEqual(const uvec2 &a, const uvec2 &b) { return bvec2{a.x >= b.x, a.y >= b.y}; }
inline bvec2 greaterThan(const vec2 &a, const vec2 &b) { return bvec2{a.x > b.x, a.y > b.y}; }
inline bvec2 lessThan(const vec2 &a, const vec2 &b) { return bvec2{a.x < b.x, a.y < b.y}; }
inline bool any(const bvec2 &b) { return b.x || b.y; }
inline bool all(const bvec2 &b) { return b.x && b.y; }

// ---- textures: nearest fetch of the reference's NVIDIA-default G-buffer formats (gBufferPass.cpp:75-108) -------
// A null plane reads as zero (the first frame's "previous" G-buffer, which the reference leaves undefined).
enum TexelFormat { kRGBA8_SRGB, kRGBA16_SNORM, kRG16_UNORM, kRGBA32F, kD32F };
struct sampler2D {
	TexelFormat format;
	const void *data;
	int width, height;
};
inline float srgb8_to_linear(unsigned c) { // P12: the sRGB EOTF in double, rounded once
	static float table[256];
	static bool ready = [] {
		for (int i = 0; i < 256; ++i) {
			double v = i / 255.0;
			table[i] = (float)(v <= 0.04045 ? v / 12.92 : std::pow((v + 0.055) / 1.055, 2.4));
		}
		return true;
	}();
	(void)ready;
	return table[c];
}
inline vec4 texelFetch(const sampler2D &s, const ivec2 &p, int) {
	if (!s.data) {
		return vec4(0.0f);
	}
	size_t i = (size_t)p.y * (size_t)s.width + (size_t)p.x;
	switch (s.format) {
	case kRGBA8_SRGB: {
		const uint8_t *t = (const uint8_t *)s.data + i * 4;
		return vec4(srgb8_to_linear(t[0]), srgb8_to_linear(t[1]), srgb8_to_linear(t[2]), (float)t[3] / 255.0f);
	}
	case kRGBA16_SNORM: {
		const int16_t *t = (const int16_t *)s.data + i * 4;
		return vec4(::fmaxf((float)t[0] / 32767.0f, -1.0f), ::fmaxf((float)t[1] / 32767.0f, -1.0f), ::fmaxf((float)t[2] / 32767.0f, -1.0f),
		            ::fmaxf((float)t[3] / 32767.0f, -1.0f));
	}
	case kRG16_UNORM: {
		const uint16_t *t = (const uint16_t *)s.data + i * 2;
		return vec4((float)t[0] / 65535.0f, (float)t[1] / 65535.0f, 0.0f, 1.0f);
	}
	case kRGBA32F: {
		const float *t = (const float *)s.data + i * 4;
		return vec4(t[0], t[1], t[2], t[3]);
	}
	default: {
		const float *t = (const float *)s.data + i;
		return vec4(t[0], 0.0f, 0.0f, 1.0f);
	}
	}
}
// texture() with the nearest sampler of restirPass.h:362 at a pixel-centre uv
inline vec4 texture(const sampler2D &s, const vec2 &uv) {
	return texelFetch(s, ivec2((int)::floorf(uv.x * (float)s.width), (int)::floorf(uv.y * (float)s.height)), 0);
}

// ---- per-invocation built-in variables --------------------------------------------------------------------------
static thread_local uvec3 gl_GlobalInvocationID;
static thread_local fragcoord gl_FragCoord;

} // namespace glsl
