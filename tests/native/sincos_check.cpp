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
	// atanf over every stride-th float, atan2f over pseudo-random pairs (raw bit patterns and comparable magnitudes)
	unsigned long long na = 0, ba = 0, n2 = 0, b2 = 0;
	for (uint64_t b = 0; b < (1ull << 32); b += stride) {
		float f; uint32_t u = (uint32_t)b; memcpy(&f, &u, 4);
		if (f != f) continue;
		na++;
		float a = atanf(f), c = rsb::rs_atanf(f);
		if (memcmp(&a, &c, 4)) ba++;
	}
	uint64_t s = 88172645463325252ull;
	for (uint64_t i = 0; i < (1ull << 32) / stride; i++) {
		s ^= s << 13; s ^= s >> 7; s ^= s << 17;
		float y, x; uint32_t lo = (uint32_t)s, hi = (uint32_t)(s >> 32);
		memcpy(&y, &lo, 4); memcpy(&x, &hi, 4);
		if (i & 1) { x = (float)((int32_t)hi) * (1.0f / 1073741824.0f); y = (float)((int32_t)lo) * (1.0f / 1073741824.0f); if ((i & 6) == 2) y *= 1e-3f; if ((i & 6) == 4) x *= 1e-4f; }
		if (x != x || y != y) continue;
		n2++;
		float a = atan2f(y, x), c = rsb::rs_atan2f(y, x);
		if (memcmp(&a, &c, 4)) b2++;
	}
	printf("atanf checked %llu mismatch %llu atan2f checked %llu mismatch %llu\n", na, ba, n2, b2);
	return 0;
}
