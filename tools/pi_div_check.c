#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
int main() {
	const float pi = 3.14159274f;
	const float c = 1.0f / pi; // RN(1/pi_f)
	printf("c = %.9g (0x%08x)\n", c, *(uint32_t *)&c);
	long bad = 0, n = 0;
	for (int e = 1; e <= 253; e += 1) { // every normal binade whose quotient stays normal
		for (uint32_t m = 0; m < (1u << 23); m += (e >= 100 && e <= 140) ? 1 : 64) { // exhaustive around the values that occur, sampled elsewhere
			uint32_t bits = ((uint32_t)e << 23) | m;
			float x; memcpy(&x, &bits, 4);
			float want = x / pi;
			float q = x * c;
			float r = fmaf(-pi, q, x);
			float got = fmaf(r, c, q);
			++n;
			if (got != want) { if (bad < 10) printf("x=%.9g want %.9g got %.9g\n", x, want, got); ++bad; }
		}
	}
	printf("%ld values, %ld mismatches\n", n, bad);
	return 0;
}
