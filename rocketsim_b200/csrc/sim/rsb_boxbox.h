// rsb_boxbox.h -- car-car contact generation: 15-axis SAT between two oriented boxes, then
// either the closest points of the two touching edges or the incident face clipped against the
// reference face, culled to 4 points.  Restates the ODE-derived detector the reference uses,
//   btBoxBoxDetector::getClosestPoints / dBoxBox2 / intersectRectQuad2 / cullPoints2
//   bt/BulletCollision/CollisionDispatch/btBoxBoxDetector.cpp:60-767
// with the matrices held as three row vectors instead of 4-float padded rows.
#pragma once
#include <stdio.h>
#include "rsb_work.h"
#include "rsb_simplex.h"

namespace rsb {

RSB_ACP_ATTR void add_contact_point(const Ctx& cx, int mi, const V3& normalOnB, const V3& pointInWorld, float depth, int triMesh, int triIndex, bool dynPair);

// R[i*4+j] of the ODE code is element (row i, col j); a "column j" walk is R + j with stride 4.
struct Mat12 { float m[12]; };
RSB_HD Mat12 mat12_from(const M3& b) {
	Mat12 r;
	for (int j = 0; j < 3; j++) { r.m[0 + 4 * j] = b.r[j].x; r.m[1 + 4 * j] = b.r[j].y; r.m[2 + 4 * j] = b.r[j].z; r.m[3 + 4 * j] = 0.f; }
	return r;
}
RSB_HD float ddot_pq(const float* a, const float* b, int p, int q) { return (a[0] * b[0] + a[p] * b[q]) + a[2 * p] * b[2 * q]; }

RSB_HD void line_closest_approach(const float* pa, const float* ua, const float* pb, const float* ub, float* alpha, float* beta) {
	float p[3] = {pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]};
	float uaub = ddot_pq(ua, ub, 1, 1);
	float q1 = ddot_pq(ua, p, 1, 1);
	float q2 = -ddot_pq(ub, p, 1, 1);
	float d = 1 - uaub * uaub;
	if (d <= 0.0001f) { *alpha = 0; *beta = 0; }
	else { d = 1.f / d; *alpha = (q1 + uaub * q2) * d; *beta = (uaub * q1 + q2) * d; }
}

// intersectRectQuad2 (btBoxBoxDetector.cpp:160-215): clip quad p against the rectangle +-h, one half-plane at a time.
// Written with value copies between the clip stages (the reference ping-pongs two pointers between `ret` and a scratch
// buffer; data-dependent pointers into local arrays are avoided in device code).
RSB_HD int intersect_rect_quad2(const float h[2], const float p[8], float ret[16]) {
	float q[16];
	int nq = 4, nr = 0;
	for (int i = 0; i < 8; i++) q[i] = p[i];
	bool done = false;
	for (int dir = 0; dir <= 1 && !done; dir++) {
		for (int sign = -1; sign <= 1 && !done; sign += 2) {
			float r[16];
			nr = 0;
			for (int i = 0; i < nq; i++) {
				const float pq0 = q[2 * i], pq1 = q[2 * i + 1];
				const float pqd = dir ? pq1 : pq0, pqo = dir ? pq0 : pq1;
				if (sign * pqd < h[dir]) {
					r[2 * nr] = pq0; r[2 * nr + 1] = pq1; nr++;
					if (nr & 8) { done = true; break; }
				}
				const int ni = (i < nq - 1) ? i + 1 : 0;
				const float nq0 = q[2 * ni], nq1 = q[2 * ni + 1];
				const float nqd = dir ? nq1 : nq0, nqo = dir ? nq0 : nq1;
				if ((sign * pqd < h[dir]) ^ (sign * nqd < h[dir])) {
					const float vo = pqo + (nqo - pqo) / (nqd - pqd) * (sign * h[dir] - pqd);
					const float vd = sign * h[dir];
					r[2 * nr + (1 - dir)] = vo; r[2 * nr + dir] = vd; nr++;
					if (nr & 8) { done = true; break; }
				}
			}
			for (int i = 0; i < nr * 2; i++) q[i] = r[i];
			nq = nr;
		}
	}
	for (int i = 0; i < nr * 2; i++) ret[i] = q[i];
	return nr;
}

RSB_HD void cull_points2(int n, float p[], int m, int i0, int iret[]) {
	const float M__PI = 3.14159265f;
	int i, j;
	float a, cx, cy, q;
	if (n == 1) { cx = p[0]; cy = p[1]; }
	else if (n == 2) { cx = 0.5f * (p[0] + p[2]); cy = 0.5f * (p[1] + p[3]); }
	else {
		a = 0; cx = 0; cy = 0;
		for (i = 0; i < (n - 1); i++) {
			q = p[i * 2] * p[i * 2 + 3] - p[i * 2 + 2] * p[i * 2 + 1];
			a += q;
			cx += q * (p[i * 2] + p[i * 2 + 2]);
			cy += q * (p[i * 2 + 1] + p[i * 2 + 3]);
		}
		q = p[n * 2 - 2] * p[1] - p[0] * p[n * 2 - 1];
		if (fabsf(a + q) > kEps) a = 1.f / (3.0f * (a + q));
		else a = kLargeFloat;
		cx = a * (cx + q * (p[n * 2 - 2] + p[0]));
		cy = a * (cy + q * (p[n * 2 - 1] + p[1]));
	}
	float A[8];
	for (i = 0; i < n; i++) A[i] = rs_atan2f(p[i * 2 + 1] - cy, p[i * 2] - cx);
	int avail[8];
	for (i = 0; i < n; i++) avail[i] = 1;
	avail[i0] = 0;
	iret[0] = i0;
	iret++;
	for (j = 1; j < m; j++) {
		a = (float)j * (2 * M__PI / m) + A[i0];
		if (a > M__PI) a -= 2 * M__PI;
		float maxdiff = 1e9f, diff;
		*iret = i0;
		for (i = 0; i < n; i++) {
			if (avail[i]) {
				diff = fabsf(A[i] - a);
				if (diff > M__PI) diff = 2 * M__PI - diff;
				if (diff < maxdiff) { maxdiff = diff; *iret = i; }
			}
		}
		avail[*iret] = 0;
		iret++;
	}
}

// boxes 1/2 = (shape, world transform); the manifold's body0 = owner of box 1.  Pure detector: returns the number of contacts
// (<= 4) and writes the common normal on B, the points and the signed depths; the commit phase adds them in order.
RSB_HD int box_box_detect(const Ctx& cx, const Convex& box1, const Xform& xf1, const Convex& box2, const Xform& xf2, V3& outNormalOnB, V3 outPts[4], float outDep[4]) {
	const float fudge_factor = 1.05f;
#if defined(RSB_PROFILE)
	if (cx.epa) cx.epa->profCalls += 1;
#endif
	Mat12 R1m = mat12_from(xf1.basis), R2m = mat12_from(xf2.basis);
	const float* R1 = R1m.m; const float* R2 = R2m.m;
	const float p1[3] = {xf1.origin.x, xf1.origin.y, xf1.origin.z};
	const float p2[3] = {xf2.origin.x, xf2.origin.y, xf2.origin.z};
	V3 he1 = box1.he + V3(box1.margin, box1.margin, box1.margin), he2 = box2.he + V3(box2.margin, box2.margin, box2.margin);
	const float side1[3] = {2.f * he1.x, 2.f * he1.y, 2.f * he1.z};
	const float side2[3] = {2.f * he2.x, 2.f * he2.y, 2.f * he2.z};
	float p[3], pp[3], normalC[3] = {0.f, 0.f, 0.f};
	float nsel0 = 0.f, nsel1 = 0.f, nsel2 = 0.f; int haveR = 0;  // the selected face normal = column `norm` of R1/R2 (copied: no pointers into local arrays)
	float A[3], B[3], R11, R12, R13, R21, R22, R23, R31, R32, R33, Q11, Q12, Q13, Q21, Q22, Q23, Q31, Q32, Q33, s, s2, l;
	int i, j, invert_normal, code;
	float normal[3], depth;

	p[0] = p2[0] - p1[0]; p[1] = p2[1] - p1[1]; p[2] = p2[2] - p1[2];
	pp[0] = ddot_pq(R1 + 0, p, 4, 1); pp[1] = ddot_pq(R1 + 1, p, 4, 1); pp[2] = ddot_pq(R1 + 2, p, 4, 1);
	A[0] = side1[0] * 0.5f; A[1] = side1[1] * 0.5f; A[2] = side1[2] * 0.5f;
	B[0] = side2[0] * 0.5f; B[1] = side2[1] * 0.5f; B[2] = side2[2] * 0.5f;
	R11 = ddot_pq(R1 + 0, R2 + 0, 4, 4); R12 = ddot_pq(R1 + 0, R2 + 1, 4, 4); R13 = ddot_pq(R1 + 0, R2 + 2, 4, 4);
	R21 = ddot_pq(R1 + 1, R2 + 0, 4, 4); R22 = ddot_pq(R1 + 1, R2 + 1, 4, 4); R23 = ddot_pq(R1 + 1, R2 + 2, 4, 4);
	R31 = ddot_pq(R1 + 2, R2 + 0, 4, 4); R32 = ddot_pq(R1 + 2, R2 + 1, 4, 4); R33 = ddot_pq(R1 + 2, R2 + 2, 4, 4);
	Q11 = fabsf(R11); Q12 = fabsf(R12); Q13 = fabsf(R13);
	Q21 = fabsf(R21); Q22 = fabsf(R22); Q23 = fabsf(R23);
	Q31 = fabsf(R31); Q32 = fabsf(R32); Q33 = fabsf(R33);

#define RSB_TST(expr1, expr2, norm, cc) \
	s2 = fabsf(expr1) - (expr2); \
	if (s2 > 0) return 0; \
	if (s2 > s) { s = s2; nsel0 = (norm)[0]; nsel1 = (norm)[4]; nsel2 = (norm)[8]; haveR = 1; invert_normal = ((expr1) < 0); code = (cc); }
	s = -3.402823466e+38f;
	invert_normal = 0;
	code = 0;
	RSB_TST(pp[0], (A[0] + B[0] * Q11 + B[1] * Q12 + B[2] * Q13), R1 + 0, 1);
	RSB_TST(pp[1], (A[1] + B[0] * Q21 + B[1] * Q22 + B[2] * Q23), R1 + 1, 2);
	RSB_TST(pp[2], (A[2] + B[0] * Q31 + B[1] * Q32 + B[2] * Q33), R1 + 2, 3);
	RSB_TST(ddot_pq(R2 + 0, p, 4, 1), (A[0] * Q11 + A[1] * Q21 + A[2] * Q31 + B[0]), R2 + 0, 4);
	RSB_TST(ddot_pq(R2 + 1, p, 4, 1), (A[0] * Q12 + A[1] * Q22 + A[2] * Q32 + B[1]), R2 + 1, 5);
	RSB_TST(ddot_pq(R2 + 2, p, 4, 1), (A[0] * Q13 + A[1] * Q23 + A[2] * Q33 + B[2]), R2 + 2, 6);
#undef RSB_TST
#define RSB_TST(expr1, expr2, n1, n2, n3, cc) \
	s2 = fabsf(expr1) - (expr2); \
	if (s2 > kEps) return 0; \
	l = sqrtf((n1) * (n1) + (n2) * (n2) + (n3) * (n3)); \
	if (l > kEps) { \
		s2 /= l; \
		if (s2 * fudge_factor > s) { \
			s = s2; haveR = 0; \
			normalC[0] = (n1) / l; normalC[1] = (n2) / l; normalC[2] = (n3) / l; \
			invert_normal = ((expr1) < 0); code = (cc); \
		} \
	}
	const float fudge2 = 1.0e-5f;
	Q11 += fudge2; Q12 += fudge2; Q13 += fudge2;
	Q21 += fudge2; Q22 += fudge2; Q23 += fudge2;
	Q31 += fudge2; Q32 += fudge2; Q33 += fudge2;
	RSB_TST(pp[2] * R21 - pp[1] * R31, (A[1] * Q31 + A[2] * Q21 + B[1] * Q13 + B[2] * Q12), 0, -R31, R21, 7);
	RSB_TST(pp[2] * R22 - pp[1] * R32, (A[1] * Q32 + A[2] * Q22 + B[0] * Q13 + B[2] * Q11), 0, -R32, R22, 8);
	RSB_TST(pp[2] * R23 - pp[1] * R33, (A[1] * Q33 + A[2] * Q23 + B[0] * Q12 + B[1] * Q11), 0, -R33, R23, 9);
	RSB_TST(pp[0] * R31 - pp[2] * R11, (A[0] * Q31 + A[2] * Q11 + B[1] * Q23 + B[2] * Q22), R31, 0, -R11, 10);
	RSB_TST(pp[0] * R32 - pp[2] * R12, (A[0] * Q32 + A[2] * Q12 + B[0] * Q23 + B[2] * Q21), R32, 0, -R12, 11);
	RSB_TST(pp[0] * R33 - pp[2] * R13, (A[0] * Q33 + A[2] * Q13 + B[0] * Q22 + B[1] * Q21), R33, 0, -R13, 12);
	RSB_TST(pp[1] * R11 - pp[0] * R21, (A[0] * Q21 + A[1] * Q11 + B[1] * Q33 + B[2] * Q32), -R21, R11, 0, 13);
	RSB_TST(pp[1] * R12 - pp[0] * R22, (A[0] * Q22 + A[1] * Q12 + B[0] * Q33 + B[2] * Q31), -R22, R12, 0, 14);
	RSB_TST(pp[1] * R13 - pp[0] * R23, (A[0] * Q23 + A[1] * Q13 + B[0] * Q32 + B[1] * Q31), -R23, R13, 0, 15);
#undef RSB_TST
	if (!code) return 0;

	if (haveR) { normal[0] = nsel0; normal[1] = nsel1; normal[2] = nsel2; }
	else { normal[0] = ddot_pq(R1, normalC, 1, 1); normal[1] = ddot_pq(R1 + 4, normalC, 1, 1); normal[2] = ddot_pq(R1 + 8, normalC, 1, 1); }
	if (invert_normal) { normal[0] = -normal[0]; normal[1] = -normal[1]; normal[2] = -normal[2]; }
	depth = -s;
	outNormalOnB = V3(-normal[0], -normal[1], -normal[2]);
	int nOut = 0;  // <= maxc = 4 contacts

	if (code > 6) {
		float pa[3], pb[3], sign;
		for (i = 0; i < 3; i++) pa[i] = p1[i];
		for (j = 0; j < 3; j++) {
			sign = (ddot_pq(normal, R1 + j, 1, 4) > 0) ? 1.0f : -1.0f;
			for (i = 0; i < 3; i++) pa[i] += sign * A[j] * R1[i * 4 + j];
		}
		for (i = 0; i < 3; i++) pb[i] = p2[i];
		for (j = 0; j < 3; j++) {
			sign = (ddot_pq(normal, R2 + j, 1, 4) > 0) ? -1.0f : 1.0f;
			for (i = 0; i < 3; i++) pb[i] += sign * B[j] * R2[i * 4 + j];
		}
		float alpha, beta, ua[3], ub[3];
		for (i = 0; i < 3; i++) ua[i] = R1[((code)-7) / 3 + i * 4];
		for (i = 0; i < 3; i++) ub[i] = R2[((code)-7) % 3 + i * 4];
		line_closest_approach(pa, ua, pb, ub, &alpha, &beta);
		for (i = 0; i < 3; i++) pa[i] += ua[i] * alpha;
		for (i = 0; i < 3; i++) pb[i] += ub[i] * beta;
		outPts[0] = V3(pb[0], pb[1], pb[2]); outDep[0] = -depth; nOut = 1;
		goto emit;
	}
	{

	// reference/incident box selected by VALUE (element-wise copies; no data-dependent pointers into local arrays)
	float Ra[12], Rb[12], pa[3], pb[3], Sa[3], Sb[3];
	{
		const bool first = code <= 3;
		for (i = 0; i < 12; i++) { Ra[i] = first ? R1[i] : R2[i]; Rb[i] = first ? R2[i] : R1[i]; }
		for (i = 0; i < 3; i++) { pa[i] = first ? p1[i] : p2[i]; pb[i] = first ? p2[i] : p1[i]; Sa[i] = first ? A[i] : B[i]; Sb[i] = first ? B[i] : A[i]; }
	}
	float normal2[3], nr[3], anr[3];
	if (code <= 3) { normal2[0] = normal[0]; normal2[1] = normal[1]; normal2[2] = normal[2]; }
	else { normal2[0] = -normal[0]; normal2[1] = -normal[1]; normal2[2] = -normal[2]; }
	nr[0] = ddot_pq(Rb + 0, normal2, 4, 1); nr[1] = ddot_pq(Rb + 1, normal2, 4, 1); nr[2] = ddot_pq(Rb + 2, normal2, 4, 1);
	anr[0] = fabsf(nr[0]); anr[1] = fabsf(nr[1]); anr[2] = fabsf(nr[2]);
	int lanr, a1, a2;
	if (anr[1] > anr[0]) {
		if (anr[1] > anr[2]) { a1 = 0; lanr = 1; a2 = 2; }
		else { a1 = 0; a2 = 1; lanr = 2; }
	} else {
		if (anr[0] > anr[2]) { lanr = 0; a1 = 1; a2 = 2; }
		else { a1 = 0; a2 = 1; lanr = 2; }
	}
	float center[3];
	if (nr[lanr] < 0) { for (i = 0; i < 3; i++) center[i] = pb[i] - pa[i] + Sb[lanr] * Rb[i * 4 + lanr]; }
	else { for (i = 0; i < 3; i++) center[i] = pb[i] - pa[i] - Sb[lanr] * Rb[i * 4 + lanr]; }
	int codeN, code1, code2;
	if (code <= 3) codeN = code - 1; else codeN = code - 4;
	if (codeN == 0) { code1 = 1; code2 = 2; }
	else if (codeN == 1) { code1 = 0; code2 = 2; }
	else { code1 = 0; code2 = 1; }
	float quad[8];
	float c1, c2, m11, m12, m21, m22;
	c1 = ddot_pq(center, Ra + code1, 1, 4);
	c2 = ddot_pq(center, Ra + code2, 1, 4);
	m11 = ddot_pq(Ra + code1, Rb + a1, 4, 4);
	m12 = ddot_pq(Ra + code1, Rb + a2, 4, 4);
	m21 = ddot_pq(Ra + code2, Rb + a1, 4, 4);
	m22 = ddot_pq(Ra + code2, Rb + a2, 4, 4);
	{
		float k1 = m11 * Sb[a1], k2 = m21 * Sb[a1], k3 = m12 * Sb[a2], k4 = m22 * Sb[a2];
		quad[0] = c1 - k1 - k3; quad[1] = c2 - k2 - k4;
		quad[2] = c1 - k1 + k3; quad[3] = c2 - k2 + k4;
		quad[4] = c1 + k1 + k3; quad[5] = c2 + k2 + k4;
		quad[6] = c1 + k1 - k3; quad[7] = c2 + k2 - k4;
	}
	float rect[2] = {Sa[code1], Sa[code2]};
	float ret[16];
	int n = intersect_rect_quad2(rect, quad, ret);
	if (n < 1) return 0;
	float point[3 * 8], dep[8];
	float det1 = 1.f / (m11 * m22 - m12 * m21);
	m11 *= det1; m12 *= det1; m21 *= det1; m22 *= det1;
	int cnum = 0;
	for (j = 0; j < n; j++) {
		float k1 = m22 * (ret[j * 2] - c1) - m12 * (ret[j * 2 + 1] - c2);
		float k2 = -m21 * (ret[j * 2] - c1) + m11 * (ret[j * 2 + 1] - c2);
		for (i = 0; i < 3; i++) point[cnum * 3 + i] = center[i] + k1 * Rb[i * 4 + a1] + k2 * Rb[i * 4 + a2];
		dep[cnum] = Sa[codeN] - ddot_pq(normal2, point + cnum * 3, 1, 1);
		if (dep[cnum] >= 0) {
			ret[cnum * 2] = ret[j * 2];
			ret[cnum * 2 + 1] = ret[j * 2 + 1];
			cnum++;
		}
	}
	if (cnum < 1) return 0;
	int maxc = 4;
	if (maxc > cnum) maxc = cnum;
	if (maxc < 1) maxc = 1;
	const V3 nrm(normal[0], normal[1], normal[2]);
	if (cnum <= maxc) {
		for (j = 0; j < cnum; j++) {
			V3 pw(point[j * 3 + 0] + pa[0], point[j * 3 + 1] + pa[1], point[j * 3 + 2] + pa[2]);
			if (code < 4) outPts[nOut] = pw;
			else outPts[nOut] = V3(point[j * 3 + 0] + pa[0] - normal[0] * dep[j], point[j * 3 + 1] + pa[1] - normal[1] * dep[j], point[j * 3 + 2] + pa[2] - normal[2] * dep[j]);
			outDep[nOut++] = -dep[j];
		}
	} else {
		int i1 = 0;
		float maxdepth = dep[0];
		for (i = 1; i < cnum; i++) if (dep[i] > maxdepth) { maxdepth = dep[i]; i1 = i; }
		int iret[8];
		cull_points2(cnum, ret, maxc, i1, iret);
		for (j = 0; j < maxc; j++) {
			V3 posInWorld(point[iret[j] * 3 + 0] + pa[0], point[iret[j] * 3 + 1] + pa[1], point[iret[j] * 3 + 2] + pa[2]);
			if (code < 4) outPts[nOut] = posInWorld;
			else outPts[nOut] = posInWorld - nrm * dep[iret[j]];
			outDep[nOut++] = -dep[iret[j]];
		}
	}
	}
emit:
	return nOut;
}

}  // namespace rsb
