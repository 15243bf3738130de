/* rsb_bench_inputs.h -- the pinned "README bench" control schedule.
 *
 * The reference README (README.md:49-53) benchmarks Soccar 2v2 with "inputs randomly
 * pre-generated, changed every 2-60 ticks per car".  It does not publish the
 * distribution; this header is our pinned restatement so that the reference CPU arm,
 * the CUDA arm and the parity tests all consume the SAME control stream:
 *
 *   at tick t >= next:  throttle, steer, pitch, yaw, roll ~ U(-1,1)
 *                       jump ~ B(0.25), boost ~ B(0.5), handbrake ~ B(0.2)
 *                       next = t + 2 + (r % 59)            (hold 2..60 ticks)
 *
 * RNG: counter-based (splitmix64 finaliser) keyed by (seed, arena, car, draw#), so
 * any (arena, car) stream can be generated independently on host or device.
 *
 * Plain C, usable from C, C++, CUDA (__host__ __device__) and via ctypes wrappers.
 */
#ifndef RSB_BENCH_INPUTS_H
#define RSB_BENCH_INPUTS_H

#include <stdint.h>

#if defined(__CUDACC__)
#define RSB_INP_HD __host__ __device__ __forceinline__
#else
#define RSB_INP_HD static inline
#endif

typedef struct RsbInputSched {
	uint32_t next_tick; /* first tick at which a new draw happens */
	uint32_t draws;     /* number of draws made so far */
	float throttle, steer, pitch, yaw, roll;
	uint8_t jump, boost, handbrake, _pad;
} RsbInputSched;

RSB_INP_HD uint64_t rsb_inp_mix(uint64_t z) {
	z += 0x9E3779B97F4A7C15ull;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

RSB_INP_HD uint64_t rsb_inp_rand(uint64_t seed, uint32_t arena, uint32_t car, uint32_t draw, uint32_t lane) {
	uint64_t k = rsb_inp_mix(seed ^ rsb_inp_mix(((uint64_t)arena << 32) | car));
	return rsb_inp_mix(k ^ (((uint64_t)draw << 8) | lane));
}

/* uniform in [-1, 1): 24 random bits -> exact float */
RSB_INP_HD float rsb_inp_unit(uint64_t r) {
	return (float)(int32_t)((r >> 40) & 0xFFFFFF) * (1.0f / 8388608.0f) - 1.0f;
}

RSB_INP_HD void rsb_inp_init(RsbInputSched* s) {
	s->next_tick = 0; s->draws = 0;
	s->throttle = s->steer = s->pitch = s->yaw = s->roll = 0.f;
	s->jump = s->boost = s->handbrake = s->_pad = 0;
}

/* Advance the schedule to `tick` (call once per tick, ticks ascending). */
RSB_INP_HD void rsb_inp_update(RsbInputSched* s, uint64_t seed, uint32_t arena, uint32_t car, uint32_t tick) {
	if (tick < s->next_tick) return;
	uint32_t d = s->draws++;
	s->throttle = rsb_inp_unit(rsb_inp_rand(seed, arena, car, d, 0));
	s->steer    = rsb_inp_unit(rsb_inp_rand(seed, arena, car, d, 1));
	s->pitch    = rsb_inp_unit(rsb_inp_rand(seed, arena, car, d, 2));
	s->yaw      = rsb_inp_unit(rsb_inp_rand(seed, arena, car, d, 3));
	s->roll     = rsb_inp_unit(rsb_inp_rand(seed, arena, car, d, 4));
	uint64_t r = rsb_inp_rand(seed, arena, car, d, 5);
	s->jump      = ((r & 0xFF) < 64) ? 1 : 0;          /* 0.25 */
	s->boost     = (((r >> 8) & 0xFF) < 128) ? 1 : 0;  /* 0.5  */
	s->handbrake = (((r >> 16) & 0xFF) < 51) ? 1 : 0;  /* ~0.2 */
	s->next_tick = tick + 2 + (uint32_t)((r >> 24) % 59);
}

#endif
