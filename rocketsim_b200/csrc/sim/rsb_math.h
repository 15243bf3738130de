// rsb_math.h -- fp32 vector / matrix / quaternion primitives for the batched arena tick.
//
// Semantics follow the math the reference path runs on (vendored Bullet LinearMath, SSE
// build: /root/reference/libsrc/bullet3-3.24/LinearMath/btVector3.h:229-249 dot =
// (x*x'+y*y')+z*z'; btMatrix3x3.h:216-288 setRotation, :421-499 getRotation (SIMD
// variant), :1093-1104 inverse; btQuaternion.h:108-115; btTransformUtil.h:32-87), but the
// types are our own plain structs usable from CUDA device code and from a host compile
// (tests only).  Compile with -fmad=false / -ffp-contract=off: the reference build has no
// FMA contraction.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define RSB_HD __host__ __device__ __forceinline__
#define RSB_HD_NOINLINE __host__ __device__ __noinline__
#else
#define RSB_HD inline
#define RSB_HD_NOINLINE
#endif
// GJK pair detector inlined at its call sites (measured faster than the out-of-line call with by-value aggregates)
#define RSB_CC_ATTR RSB_HD
// A/B switch (developer): -DRSB_ACP_NOINLINE makes the contact-add hook one out-of-line function (miscompiles with nvcc 12.9: car-car contacts vanish)
#if !defined(RSB_ACP_NOINLINE)
#define RSB_ACP_ATTR RSB_HD
#else
#define RSB_ACP_ATTR RSB_HD_NOINLINE
#endif

namespace rsb {

constexpr float kEps = 1.1920928955078125e-7f;  // FLT_EPSILON (SIMD_EPSILON)
constexpr float kPi = 3.1415926535897932384626433832795029f;
constexpr float k2Pi = 2.0f * kPi;
constexpr float kHalfPi = kPi * 0.5f;
constexpr float kSqrt12 = 0.7071067811865475244008443621048490f;
constexpr float UU_TO_BT = 1.f / 50.f;
constexpr float BT_TO_UU = 50.f;

RSB_HD float fminf_(float a, float b) { return a < b ? a : b; }
RSB_HD float fmaxf_(float a, float b) { return a > b ? a : b; }
// RS_CLAMP(val,min,max) = RS_MIN(RS_MAX(val,min),max) with (a>b)?a:b macros (Framework.h:47-50)
RSB_HD float rs_clamp(float v, float lo, float hi) {
	float t = (v > lo) ? v : lo;
	return (t < hi) ? t : hi;
}
RSB_HD int rs_sgn(float v) { return (v > 0) - (v < 0); }

// sinf / cosf exactly as the reference's libm computes them.  glibc >= 2.28 (sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c,
// sincosf.h -- the ARM optimized-routines algorithm): argument reduced with one multiply-subtract in DOUBLE precision, a
// double-precision minimax polynomial per quadrant, one final rounding to float.  x86-64 hosts with FMA (every current server
// CPU) run the -mfma ifunc variant, in which every a * b + c below is a fused operation; this restatement was compared with
// the host's sinf / cosf on ALL 2,246,049,792 floats |x| < 120: 0 mismatches (without FMA: 12 / 22 one-ulp differences).
// The device libm's sinf / cosf are only 1-2 ulp accurate; with these, rotations integrate bit-identically to the reference.
struct SinCosTab { double c0, c1, c2, c3, c4, s1, s2, s3; };
RSB_HD double rs_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
	return __fma_rn(a, b, c);
#else
	return __builtin_fma(a, b, c);
#endif
}
RSB_HD float sincosf_poly(double x, double x2, bool neg, int n) {
	// __sincosf_table[0] / [1]: the second entry negates the cosine coefficients (quadrants 2, 3)
	const double sg = neg ? -1.0 : 1.0;
	if ((n & 1) == 0) {
		const double x3 = x * x2;
		const double s1 = rs_fma(x2, -0x1.994eb3774cf24p-13, 0x1.1107605230bc4p-7);
		const double x7 = x3 * x2;
		const double s = rs_fma(x3, -0x1.555545995a603p-3, x);
		return (float)rs_fma(x7, s1, s);
	}
	const double x4 = x2 * x2;
	const double c2 = rs_fma(x2, sg * 0x1.99343027bf8c3p-16, sg * -0x1.6c087e89a359dp-10);
	const double c1 = rs_fma(x2, sg * -0x1.ffffffd0c621cp-2, sg * 0x1p0);
	const double x6 = x4 * x2;
	const double c = rs_fma(x4, sg * 0x1.55553e1068f19p-5, c1);
	return (float)rs_fma(x6, c2, c);
}
RSB_HD uint32_t rs_abstop12(float x) {
#if defined(__CUDA_ARCH__)
	return (__float_as_uint(x) >> 20) & 0x7ffu;
#else
	uint32_t u; memcpy(&u, &x, 4); return (u >> 20) & 0x7ffu;
#endif
}
// which = 0: sinf, 1: cosf
RSB_HD float rs_sincosf(float y, int which) {
	double x = y;
	const uint32_t top = rs_abstop12(y);
	if (top < 0x3f4u) {                                   // |y| < pi/4 (abstop12(0x1.921FB6p-1f))
		if (top < 0x398u) return which ? 1.0f : y;        // |y| < 2^-12
		return sincosf_poly(x, x * x, false, which);
	}
	if (top < 0x42fu) {                                   // |y| < 120
		const double r = x * 0x1.45F306DC9C883p+23;       // 2/pi * 2^24: the quadrant ends up in bits 24..31
		const int n = ((int32_t)r + 0x800000) >> 24;
		x = rs_fma(-(double)n, 0x1.921FB54442D18p0, x);
		const double sgn = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
		return sincosf_poly(x * sgn, x * x, (n & 2) != 0, n ^ which);
	}
	return which ? cosf(y) : sinf(y);                     // huge arguments do not occur on this path (large-argument reduction not restated)
}
RSB_HD float rs_sinf(float y) { return rs_sincosf(y, 0); }
RSB_HD float rs_cosf(float y) { return rs_sincosf(y, 1); }

// atanf / atan2f as the reference's libm computes them: glibc 2.39 still carries the fdlibm single-precision algorithm
// (sysdeps/ieee754/flt-32/s_atanf.c, e_atan2f.c), plain fp32 without contraction.  Verified against the host's atanf on all
// 4,278,190,082 non-NaN floats and against atan2f on 1.49e9 argument pairs: 0 mismatches (tests/native/sincos_check.cpp).
RSB_HD uint32_t rs_fbits(float f) {
#if defined(__CUDA_ARCH__)
	return __float_as_uint(f);
#else
	uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
RSB_HD float rs_bitsf(uint32_t u) {
#if defined(__CUDA_ARCH__)
	return __uint_as_float(u);
#else
	float f; memcpy(&f, &u, 4); return f;
#endif
}
RSB_HD float rs_atanf(float x) {
	const float hi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
	const float lo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
	const int32_t hx = (int32_t)rs_fbits(x), ix = hx & 0x7fffffff;
	int id;
	if (ix >= 0x4c000000) {  // |x| >= 2^25
		if (ix > 0x7f800000) return x + x;
		return hx > 0 ? hi[3] + lo[3] : -hi[3] - lo[3];
	}
	if (ix < 0x3ee00000) {   // |x| < 0.4375
		if (ix < 0x31000000) return x;
		id = -1;
	} else {
		x = fabsf(x);
		if (ix < 0x3f980000) {
			if (ix < 0x3f300000) { id = 0; x = (2.0f * x - 1.0f) / (2.0f + x); }
			else { id = 1; x = (x - 1.0f) / (x + 1.0f); }
		} else {
			if (ix < 0x401c0000) { id = 2; x = (x - 1.5f) / (1.0f + 1.5f * x); }
			else { id = 3; x = -1.0f / x; }
		}
	}
	const float z = x * x, w = z * z;
	const float s1 = z * (3.3333334327e-01f + w * (1.4285714924e-01f + w * (9.0908870101e-02f + w * (6.6610731184e-02f + w * (4.9768779427e-02f + w * 1.6285819933e-02f)))));
	const float s2 = w * (-2.0000000298e-01f + w * (-1.1111110449e-01f + w * (-7.6918758452e-02f + w * (-5.8335702866e-02f + w * -3.6531571299e-02f))));
	if (id < 0) return x - x * (s1 + s2);
	const float r = hi[id] - ((x * (s1 + s2) - lo[id]) - x);
	return hx < 0 ? -r : r;
}
RSB_HD float rs_atan2f(float y, float x) {
	const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
	const int32_t hx = (int32_t)rs_fbits(x), ix = hx & 0x7fffffff, hy = (int32_t)rs_fbits(y), iy = hy & 0x7fffffff;
	if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
	if (hx == 0x3f800000) return rs_atanf(y);
	const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
	if (iy == 0) return m < 2 ? y : (m == 2 ? pi + tiny : -pi - tiny);
	if (ix == 0) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
	if (ix == 0x7f800000) {
		if (iy == 0x7f800000) return m == 0 ? pi_o_4 + tiny : (m == 1 ? -pi_o_4 - tiny : (m == 2 ? 3.0f * pi_o_4 + tiny : -3.0f * pi_o_4 - tiny));
		return m == 0 ? 0.0f : (m == 1 ? -0.0f : (m == 2 ? pi + tiny : -pi - tiny));
	}
	if (iy == 0x7f800000) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
	const int k = (iy - ix) >> 23;
	float z;
	if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
	else if (hx < 0 && k < -60) z = 0.0f;
	else z = rs_atanf(fabsf(y / x));
	switch (m) {
	case 0: return z;
	case 1: return rs_bitsf(rs_fbits(z) ^ 0x80000000u);
	case 2: return pi - (z - pi_lo);
	default: return (z - pi_lo) - pi;
	}
}

struct V3 {
	float x, y, z;
	RSB_HD V3() {}
	RSB_HD V3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
	RSB_HD float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
	RSB_HD void set(int i, float v) { if (i == 0) x = v; else if (i == 1) y = v; else z = v; }
};

RSB_HD V3 v3(float x, float y, float z) { return V3(x, y, z); }
RSB_HD V3 v3zero() { return V3(0.f, 0.f, 0.f); }
RSB_HD V3 operator+(const V3& a, const V3& b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
RSB_HD V3 operator-(const V3& a, const V3& b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
RSB_HD V3 operator-(const V3& a) { return V3(-a.x, -a.y, -a.z); }
RSB_HD V3 operator*(const V3& a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
RSB_HD V3 operator*(float s, const V3& a) { return V3(a.x * s, a.y * s, a.z * s); }
RSB_HD V3 operator*(const V3& a, const V3& b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
RSB_HD V3 operator/(const V3& a, float s) { float r = 1.0f / s; return V3(a.x * r, a.y * r, a.z * r); }  // btVector3 /= : multiply by reciprocal
RSB_HD V3 vdiv(const V3& a, float s) { return V3(a.x / s, a.y / s, a.z / s); }                             // RocketSim::Vec / float : true division
RSB_HD V3& operator+=(V3& a, const V3& b) { a = a + b; return a; }
RSB_HD V3& operator-=(V3& a, const V3& b) { a = a - b; return a; }
RSB_HD V3& operator*=(V3& a, float s) { a = a * s; return a; }
RSB_HD float dot(const V3& a, const V3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
RSB_HD V3 cross(const V3& a, const V3& b) {
	return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
RSB_HD float len2(const V3& a) { return dot(a, a); }
RSB_HD float len(const V3& a) { return sqrtf(dot(a, a)); }
RSB_HD bool is_zero(const V3& a) { return a.x == 0.f && a.y == 0.f && a.z == 0.f; }
RSB_HD bool fuzzy_zero(const V3& a) { return len2(a) < kEps * kEps; }
// btVector3::normalize() (bt/LinearMath/btVector3.h:304-340).  The reference's x86-64 build takes the SSE branch: the 12-bit
// rsqrtss ESTIMATE of x.x followed by one Newton-Raphson step, y = y0 * (1.5 - ((x * 0.5) * y0) * y0), and the vector times y.
// The estimate is a pure function of (exponent parity, top 10 mantissa bits) on Intel CPUs; kRsqrtTab holds it (generated and
// verified against the instruction by scripts/gen_rsqrt_table.c), so the normalised vectors match the reference bit for bit.
#if defined(__CUDACC__)
#ifndef RSB_RSQRT_TAB_SPACE
#define RSB_RSQRT_TAB_SPACE __device__  // developer A/B: __constant__
#endif
static RSB_RSQRT_TAB_SPACE const uint16_t kRsqrtTabDev[2048] = {
#include "rsb_rsqrt_tab.inc"
};
#endif
static const uint16_t kRsqrtTabHost[2048] = {
#include "rsb_rsqrt_tab.inc"
};
RSB_HD float rsqrtss_estimate(float x) {
#if defined(__CUDA_ARCH__)
	const uint32_t b = __float_as_uint(x);
#else
	uint32_t b; memcpy(&b, &x, 4);
#endif
	const uint32_t e = (b >> 23) & 0xffu, m = b & 0x7fffffu;
	uint32_t r;
	if (b >> 31) r = (e == 0) ? 0xff800000u : 0xffc00000u;  // -0 / negative denormal -> -inf, negative -> NaN
	else if (e == 0) r = 0x7f800000u;                       // +0 and denormals -> +inf
	else if (e == 255) r = m ? b : 0u;                      // NaN -> NaN, +inf -> 0
	else {
		const uint32_t idx = ((e & 1u) << 10) | (m >> 13);
		const int k = ((int)e - ((e & 1u) ? 127 : 128)) / 2;
#if defined(__CUDA_ARCH__)
		const uint32_t t = kRsqrtTabDev[idx];
#else
		const uint32_t t = kRsqrtTabHost[idx];
#endif
		r = (t << 11) + (126u << 23) - ((uint32_t)k << 23);    // entry 0x1000 carries into the exponent: exactly 1.0
	}
#if defined(__CUDA_ARCH__)
	return __uint_as_float(r);
#else
	float f; memcpy(&f, &r, 4); return f;
#endif
}
RSB_HD V3 normalized(const V3& a) {
	float vd = len2(a);
#if defined(RSB_EXP_PLAIN_NORMALIZE)  // developer A/B only (breaks bit-exactness): cost of the table lookup
	{ const float r = 1.0f / sqrtf(vd); return V3(a.x * r, a.y * r, a.z * r); }
#endif
	const float y0 = rsqrtss_estimate(vd);
	vd = vd * 0.5f;
	vd = vd * y0;
	vd = vd * y0;
	const float y = y0 * (1.5f - vd);
	return V3(a.x * y, a.y * y, a.z * y);
}
// btVector3::safeNormalize (btVector3.h:292-305)
RSB_HD V3 safe_normalized(const V3& a) {
	float l2 = len2(a);
	if (l2 >= kEps * kEps) return a * (1.0f / sqrtf(l2));
	return V3(1.f, 0.f, 0.f);
}
// RocketSim::Vec::Normalized (MathTypes.h:96-103): length via sqrt, true division, zero if tiny
RSB_HD V3 rs_normalized(const V3& a) {
	float l2 = len2(a);
	float l = (l2 > 0.f) ? sqrtf(l2) : 0.f;
	if (l > kEps * kEps) return vdiv(a, l);
	return v3zero();
}
RSB_HD V3 vmin(const V3& a, const V3& b) { return V3(fminf_(a.x, b.x), fminf_(a.y, b.y), fminf_(a.z, b.z)); }
RSB_HD V3 vmax(const V3& a, const V3& b) { return V3(fmaxf_(a.x, b.x), fmaxf_(a.y, b.y), fmaxf_(a.z, b.z)); }
RSB_HD V3 vabs(const V3& a) { return V3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
RSB_HD V3 lerp(const V3& a, const V3& b, float t) {  // btVector3::setInterpolate3: s*v0 + rt*v1, s = 1-rt
	float s = 1.0f - t;
	return V3(s * a.x + t * b.x, s * a.y + t * b.y, s * a.z + t * b.z);
}

// btPlaneSpace1 (btVector3.h:1266-1294)
RSB_HD void plane_space1(const V3& n, V3& p, V3& q) {
	if (fabsf(n.z) > kSqrt12) {
		float a = n.y * n.y + n.z * n.z;
		float k = 1.0f / sqrtf(a);
		p = V3(0.f, -n.z * k, n.y * k);
		q = V3(a * k, -n.x * p.z, n.x * p.y);
	} else {
		float a = n.x * n.x + n.y * n.y;
		float k = 1.0f / sqrtf(a);
		p = V3(-n.y * k, n.x * k, 0.f);
		q = V3(-n.z * p.y, n.z * p.x, a * k);
	}
}

// Row-major 3x3 (btMatrix3x3 layout: r[i] = row i).  RocketSim's RotMat columns
// forward/right/up are the COLUMNS of this matrix (MathTypes.h:178-186).
struct M3 {
	V3 r[3];
	RSB_HD M3() {}
	RSB_HD M3(const V3& a, const V3& b, const V3& c) { r[0] = a; r[1] = b; r[2] = c; }
	RSB_HD V3 col(int j) const { return V3(r[0][j], r[1][j], r[2][j]); }
};

RSB_HD M3 m3identity() { return M3(V3(1, 0, 0), V3(0, 1, 0), V3(0, 0, 1)); }
RSB_HD V3 operator*(const M3& m, const V3& v) { return V3(dot(m.r[0], v), dot(m.r[1], v), dot(m.r[2], v)); }
// row-vector * matrix (btMatrix3x3.h:1225-1262)
RSB_HD V3 vmul(const V3& v, const M3& m) {
	return V3((v.x * m.r[0].x + v.y * m.r[1].x) + v.z * m.r[2].x,
	          (v.x * m.r[0].y + v.y * m.r[1].y) + v.z * m.r[2].y,
	          (v.x * m.r[0].z + v.y * m.r[1].z) + v.z * m.r[2].z);
}
RSB_HD M3 operator*(const M3& a, const M3& b) {
	M3 o;
	for (int i = 0; i < 3; i++) {
		const V3 ai = a.r[i];
		o.r[i] = V3((ai.x * b.r[0].x + ai.y * b.r[1].x) + ai.z * b.r[2].x,
		            (ai.x * b.r[0].y + ai.y * b.r[1].y) + ai.z * b.r[2].y,
		            (ai.x * b.r[0].z + ai.y * b.r[1].z) + ai.z * b.r[2].z);
	}
	return o;
}
RSB_HD M3 transpose(const M3& m) { return M3(m.col(0), m.col(1), m.col(2)); }
RSB_HD M3 scaled(const M3& m, const V3& s) { return M3(m.r[0] * s, m.r[1] * s, m.r[2] * s); }
RSB_HD float cofac(const M3& m, int r1, int c1, int r2, int c2) { return m.r[r1][c1] * m.r[r2][c2] - m.r[r1][c2] * m.r[r2][c1]; }
RSB_HD M3 inverse(const M3& m) {
	V3 co(cofac(m, 1, 1, 2, 2), cofac(m, 1, 2, 2, 0), cofac(m, 1, 0, 2, 1));
	float det = dot(m.r[0], co);
	float s = 1.0f / det;
	return M3(V3(co.x * s, cofac(m, 0, 2, 2, 1) * s, cofac(m, 0, 1, 1, 2) * s),
	          V3(co.y * s, cofac(m, 0, 0, 2, 2) * s, cofac(m, 0, 2, 1, 0) * s),
	          V3(co.z * s, cofac(m, 0, 1, 2, 0) * s, cofac(m, 0, 0, 1, 1) * s));
}

struct Q4 { float x, y, z, w; };
RSB_HD Q4 q4(float x, float y, float z, float w) { Q4 q; q.x = x; q.y = y; q.z = z; q.w = w; return q; }
// operator*(btQuaternion, btQuaternion), SSE variant (btQuaternion.h:619-650): (A0 - A3) + (A1 + A2) per lane
RSB_HD Q4 qmul(const Q4& a, const Q4& b) {
	return q4((a.w * b.x - a.z * b.y) + (a.x * b.w + a.y * b.z),
	          (a.w * b.y - a.x * b.z) + (a.y * b.w + a.z * b.x),
	          (a.w * b.z - a.y * b.x) + (a.z * b.w + a.x * b.y),
	          (a.w * b.w - a.z * b.z) + (-(a.x * b.x + a.y * b.y)));
}
RSB_HD float qlen2(const Q4& q) { return (q.x * q.x + q.z * q.z) + (q.y * q.y + q.w * q.w); }
// btQuaternion(axis, angle) (btQuaternion.h:108-115)
RSB_HD Q4 q_axis_angle(const V3& axis, float angle) {
	float d = len(axis);
	float s = rs_sinf(angle * 0.5f) / d;
	return q4(axis.x * s, axis.y * s, axis.z * s, rs_cosf(angle * 0.5f));
}
// btMatrix3x3::setRotation, SSE variant (btMatrix3x3.h:222-274): products of the raw components, pairwise sums, THEN the
// scale s = 2 / |q|^2, then + identity -- not the scalar branch's x * (y * s) form, which rounds differently
RSB_HD M3 m3_from_quat(const Q4& q) {
	const float d = qlen2(q);
	const float s = 2.0f / d;
	const float xx = q.x * q.x, yy = q.y * q.y, zz = q.z * q.z;
	const float xy = q.x * q.y, xz = q.x * q.z, yz = q.y * q.z;
	const float wx = q.w * q.x, wy = q.w * q.y, wz = q.w * q.z;
	return M3(V3(((-yy) + (-zz)) * s + 1.0f, (xy + (-wz)) * s + 0.0f, (xz + wy) * s + 0.0f),
	          V3((xy + wz) * s + 0.0f, ((-xx) + (-zz)) * s + 1.0f, (yz + (-wx)) * s + 0.0f),
	          V3((xz + (-wy)) * s + 0.0f, (yz + wx) * s + 0.0f, ((-xx) + (-yy)) * s + 1.0f));
}
// btMatrix3x3::getRotation, SIMD variant (btMatrix3x3.h:421-499)
RSB_HD Q4 quat_from_m3(const M3& m) {
	float trace = m.r[0].x + m.r[1].y + m.r[2].z;
	float t[4];
	float x;
	if (trace > 0.f) {
		x = trace + 1.0f;
		t[0] = m.r[2].y - m.r[1].z;
		t[1] = m.r[0].z - m.r[2].x;
		t[2] = m.r[1].x - m.r[0].y;
		t[3] = x;
	} else {
		int i, j, k;
		if (m.r[0].x < m.r[1].y) {
			if (m.r[1].y < m.r[2].z) { i = 2; j = 0; k = 1; } else { i = 1; j = 2; k = 0; }
		} else {
			if (m.r[0].x < m.r[2].z) { i = 2; j = 0; k = 1; } else { i = 0; j = 1; k = 2; }
		}
		x = m.r[i][i] - m.r[j][j] - m.r[k][k] + 1.0f;
		t[3] = m.r[k][j] - m.r[j][k];
		t[j] = m.r[j][i] + m.r[i][j];
		t[k] = m.r[k][i] + m.r[i][k];
		t[i] = x;
	}
	float s = 0.5f / sqrtf(x);
	return q4(t[0] * s, t[1] * s, t[2] * s, t[3] * s);
}

struct Xform { M3 basis; V3 origin; };
RSB_HD V3 xf_apply(const Xform& t, const V3& v) { return t.basis * v + t.origin; }  // btTransform::operator()
// btTransform::invXform: (v - origin) * basis  (row-vector product = basis^T * v)
RSB_HD V3 xf_inv_apply(const Xform& t, const V3& v) { return vmul(v - t.origin, t.basis); }

// btTransformUtil::integrateTransform (btTransformUtil.h:37-87)
RSB_HD void integrate_transform(const Xform& cur, const V3& linvel, const V3& angvel, float dt, bool noRot, Xform& out) {
	out.origin = cur.origin + linvel * dt;
	if (noRot) { out.basis = cur.basis; return; }
	float fAngle2 = len2(angvel);
	float fAngle = 0.f;
	if (fAngle2 > kEps) fAngle = sqrtf(fAngle2);
	const float kAngMotionThresh = 0.5f * kHalfPi;
	if (fAngle * dt > kAngMotionThresh) fAngle = kAngMotionThresh / dt;
	V3 axis;
	if (fAngle < 0.001f) {
		axis = angvel * (0.5f * dt - (dt * dt * dt) * 0.020833333333f * fAngle * fAngle);
	} else {
		axis = angvel * (rs_sinf(0.5f * fAngle * dt) / fAngle);
	}
	Q4 dorn = q4(axis.x, axis.y, axis.z, rs_cosf(fAngle * dt * 0.5f));
	Q4 orn0 = quat_from_m3(cur.basis);
	Q4 pred = qmul(dorn, orn0);
	float l2 = qlen2(pred);
	if (l2 > kEps) {  // safeNormalize -> normalize: q * (1/sqrt(l2))
		float inv = 1.0f / sqrtf(l2);
		pred = q4(pred.x * inv, pred.y * inv, pred.z * inv, pred.w * inv);
	}
	if (qlen2(pred) > kEps) out.basis = m3_from_quat(pred);
	else out.basis = cur.basis;
}

// btRigidBody::updateInertiaTensor (btRigidBody.cpp:253-256)
RSB_HD M3 world_inv_inertia(const M3& basis, const V3& invInertiaLocal) {
	return scaled(basis, invInertiaLocal) * transpose(basis);
}

// LinearPieceCurve::GetOutput (src/Math/Math.cpp:5-35) over a small sorted knot table.
template <int N>
struct Curve { float x[N > 0 ? N : 1]; float y[N > 0 ? N : 1]; };
template <int N>
RSB_HD float curve_eval(const Curve<N>& c, float input, float defaultOutput = 1.f) {
	if (N == 0) return defaultOutput;
	if (input <= c.x[0]) return c.y[0];
	for (int i = 1; i < N; i++) {
		if (c.x[i] > input) {
			float rangeBetween = c.x[i] - c.x[i - 1];
			float valDiffBetween = c.y[i] - c.y[i - 1];
			float f = (input - c.x[i - 1]) / rangeBetween;
			return c.y[i - 1] + valDiffBetween * f;
		}
	}
	return c.y[N - 1];
}

}  // namespace rsb
