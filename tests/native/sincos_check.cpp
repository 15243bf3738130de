// Unit check of rs_sinf / rs_cosf (rsb_math.h) against the HOST libm -- the library the reference's sinf / cosf calls resolve to.
// usage: sincos_check [stride]   (stride 1 = all 2.2e9 floats |x| < 120, ~25 s)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <cmath>
#define RSB_HOSTCHECK
#include "sim/rsb_math.h"
int main(int argc, char** argv) {
	const uint64_t stride = argc > 1 ? strtoull(argv[1], nullptr, 10) : 61;
	unsigned long long n = 0, bs = 0, bc = 0;
	for (uint64_t b = 0; b < (1ull << 32); b += stride) {
		float f; uint32_t u = (uint32_t)b; memcpy(&f, &u, 4);
		if (!(fabsf(f) < 120.f)) continue;
		n++;
		float a = sinf(f), c = rsb::rs_sinf(f);
		if (memcmp(&a, &c, 4)) bs++;
		a = cosf(f); c = rsb::rs_cosf(f);
		if (memcmp(&a, &c, 4)) bc++;
	}
	printf("checked %llu sin_mismatch %llu cos_mismatch %llu\n", n, bs, bc);
	return 0;
}
