// const_div_check.c — x / D as q = x c, q + (x - D q) c with c = RN(1 / D) (csrc/restir_pixel.cuh div_snorm16 / div_unorm16 / div_unorm8):
// the sequence is compared with the division for EVERY integer the format holds.  gcc -O2 -ffp-contract=off const_div_check.c -lm
// Prints the mismatch count per divisor (0 0 0) and exits 1 on any.
#include <math.h>
#include <stdio.h>
#include <string.h>
static unsigned bits(float f) {
	unsigned u;
	memcpy(&u, &f, 4);
	return u;
}
static float seq(float x, float D, float c) {
	float q = x * c;
	return fmaf(fmaf(-D, q, x), c, q);
}
int main(void) {
	const struct { float D, c; int lo, hi; } k[3] = {
		{32767.0f, 3.0518509447574615e-05f, -32768, 32767}, {65535.0f, 1.5259021893143654e-05f, 0, 65535}, {255.0f, 0.0039215688593685627f, 0, 255}};
	int total = 0;
	for (int i = 0; i < 3; ++i) {
		volatile float one = 1.0f, D = k[i].D;
		int bad = bits(one / D) != bits(k[i].c); // the literal IS RN(1 / D)
		for (int x = k[i].lo; x <= k[i].hi; ++x) {
			volatile float xf = (float)x;
			bad += bits(xf / D) != bits(seq(xf, k[i].D, k[i].c));
		}
		printf("%d%c", bad, i == 2 ? '\n' : ' ');
		total += bad;
	}
	return total != 0;
}
