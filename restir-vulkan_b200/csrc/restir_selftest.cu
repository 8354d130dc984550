// restir_selftest.cu — device self-test of the packed arithmetic (restir_math2.cuh) against the scalar policy
// (restir_math.cuh): the same inputs through both, bitwise comparison, mismatch counts per operation.
// Exposed as restir_tools_selftest_packed_math for tests/test_gpu_parity.py.

#include "restir_kernels.h"
#include "restir_math2.cuh"

namespace restir {

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
	x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
	return x;
}
// a float whose exponent is spread over `spread` binades around 1, random sign when `signedValue`
__device__ __forceinline__ float random_float(uint32_t h, int spread, bool signedValue) {
	uint32_t mant = h & 0x7fffffu;
	int e = 127 + (int)((h >> 23) % (uint32_t)(2 * spread + 1)) - spread;
	uint32_t bits = ((uint32_t)e << 23) | mant | ((signedValue && (h >> 31)) ? 0x80000000u : 0u);
	return __uint_as_float(bits);
}
__device__ __forceinline__ bool same_bits(float a, float b) { return __float_as_uint(a) == __float_as_uint(b) || (a != a && b != b); }

__global__ void selftest_packed_kernel(uint64_t n, uint32_t seed, unsigned long long *mismatch) {
	unsigned bad[5] = {0, 0, 0, 0, 0};
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		uint32_t h0 = mix32((uint32_t)i * 2654435761u + seed), h1 = mix32(h0 + 0x9e3779b9u), h2 = mix32(h1 + 0x9e3779b9u), h3 = mix32(h2 + 0x9e3779b9u);
		// wide exponent spread every 8th sample (exercises the out-of-range fallbacks), a few binades otherwise
		int spread = (i & 7u) == 0 ? 126 : 12;
		float a0 = random_float(h0, spread, true), a1 = random_float(h1, spread, true);
		float b0 = random_float(h2, spread, true), b1 = random_float(h3, spread, true);
		if ((i & 1023u) == 1) a0 = 0.0f;
		if ((i & 1023u) == 2) a1 = -0.0f;
		if ((i & 4095u) == 3) b0 = 0.0f;
		f2 q = div2(mk2(a0, a1), mk2(b0, b1));
		bad[0] += !same_bits(q.x, a0 / b0) + !same_bits(q.y, a1 / b1);
		f2 r = rcp2(mk2(b0, b1));
		bad[1] += !same_bits(r.x, 1.0f / b0) + !same_bits(r.y, 1.0f / b1);
		f2 s = sqrt2(mk2(fabsf(a0), a1));
		bad[2] += !same_bits(s.x, sqrtf(fabsf(a0))) + !same_bits(s.y, sqrtf(a1));
		// evaluatePHat on a random surface and two random lights (values of scene scale)
		f3 pos = mk3(random_float(h0, 3, true), random_float(h1, 3, true), random_float(h2, 3, true));
		f3 nrm = normalize3(mk3(random_float(h3, 2, true), random_float(mix32(h3), 2, true), random_float(mix32(h3 + 1), 2, true)));
		f3 cam = mk3(3.0f, 4.0f, 5.0f);
		float rough = (float)(h1 & 0xffffu) / 65535.0f, metal = (i & 3u) == 0 ? 1.0f : (float)(h2 & 0xffffu) / 65535.0f;
		Surface sf = make_surface(pos, nrm, cam, rough, metal);
		f3 l0 = mk3(random_float(mix32(h0 + 7), 4, true), random_float(mix32(h1 + 7), 4, true), random_float(mix32(h2 + 7), 4, true));
		f3 l1 = mk3(random_float(mix32(h0 + 9), 4, true), random_float(mix32(h1 + 9), 4, true), random_float(mix32(h2 + 9), 4, true));
		f3 n0 = normalize3(mk3(random_float(mix32(h0 + 11), 2, true), random_float(mix32(h1 + 11), 2, true), random_float(mix32(h2 + 11), 2, true)));
		f3 n1 = normalize3(mk3(random_float(mix32(h0 + 13), 2, true), random_float(mix32(h1 + 13), 2, true), random_float(mix32(h2 + 13), 2, true)));
		float albedoLum = (float)(h3 & 0xffu) / 255.0f, lum0 = random_float(mix32(h3 + 5), 3, false), lum1 = random_float(mix32(h3 + 6), 3, false);
		bool useN = (i & 2u) != 0;
		f2 ph = evaluate_phat2(sf, albedoLum, mk32(l0, l1), mk32(n0, n1), useN, mk2(lum0, lum1));
		float p0 = evaluate_phat(sf, albedoLum, l0, n0, useN, lum0), p1 = evaluate_phat(sf, albedoLum, l1, n1, useN, lum1);
		bad[3] += !same_bits(ph.x, p0) + !same_bits(ph.y, p1);
		bad[4] += 2;
	}
	for (int k = 0; k < 5; ++k) {
		unsigned total = __reduce_add_sync(0xffffffffu, bad[k]);
		if ((threadIdx.x & 31) == 0 && total) {
			atomicAdd(mismatch + k, (unsigned long long)total);
		}
	}
}

void launch_selftest_packed(uint64_t n, uint32_t seed, unsigned long long *mismatch, cudaStream_t s) {
	selftest_packed_kernel<<<148 * 8, 256, 0, s>>>(n, seed, mismatch);
}

} // namespace restir
