// rsb_boxbox.h -- car-car contact generation between two oriented boxes (hitbox + margin).
//
// What has to come out is fixed by the reference (btBoxBoxDetector::getClosestPoints, the ODE box-box test,
// bt/BulletCollision/CollisionDispatch/btBoxBoxDetector.cpp:60-767): the same separating-axis choice incl. its tie-breaks and
// the 1.05 edge-axis handicap, the same contact points and depths bit for bit -- so every floating-point expression below
// keeps the reference's operand order.  How it is organised is ours:
//   * a box is an Obb record (centre, three world axes, half sizes); the 15 candidate axes are ENUMERATED (face k of A, face k
//     of B, edge pair (i, j)) and evaluated by one routine each from index arithmetic, instead of 15 expanded macro bodies --
//     a third of the instructions, and the axis evaluations are independent of each other (the selection pass that follows is
//     the only sequential part);
//   * the face case works on value copies of the reference / incident box, a fixed-size 2-D polygon clipper and a fixed-size
//     point reducer -- no pointers into thread-local arrays, which would force them into local memory on the GPU.
#pragma once
#include "rsb_work.h"
#include "rsb_simplex.h"

namespace rsb {

RSB_ACP_ATTR void add_contact_point(const Ctx& cx, int mi, const V3& normalOnB, const V3& pointInWorld, float depth, int triMesh, int triIndex, bool dynPair);

struct Obb {
	V3 c;       // centre (world)
	V3 ax[3];   // world directions of the local x / y / z axes (columns of the basis)
	float h[3]; // half sizes, margin included
};
RSB_HD Obb obb_of(const Convex& box, const Xform& xf) {
	Obb o;
	o.c = xf.origin;
	for (int k = 0; k < 3; k++) o.ax[k] = xf.basis.col(k);
	// the reference doubles the half extents into side lengths and halves them again: exact in fp32
	o.h[0] = box.he.x + box.margin; o.h[1] = box.he.y + box.margin; o.h[2] = box.he.z + box.margin;
	return o;
}
struct V2 { float x, y; };
RSB_HD float v2get(const V2& p, int k) { return k ? p.y : p.x; }

// Result of the axis search: which of the 15 axes separates least (code 1-3 face of A, 4-6 face of B, 7-15 edge pair), the
// world normal pointing from A to B, and the (negative) separation along it.
struct SatHit { int code; V3 normal; float sep; };

RSB_HD bool obb_sat(const Obb& A, const Obb& B, SatHit& out) {
	const V3 d = B.c - A.c;
	float dA[3], R[3][3], Q[3][3];
	for (int i = 0; i < 3; i++) {
		dA[i] = dot(A.ax[i], d);
		for (int j = 0; j < 3; j++) { R[i][j] = dot(A.ax[i], B.ax[j]); Q[i][j] = fabsf(R[i][j]); }
	}
	float best = -3.402823466e+38f;
	int code = 0;
	bool flip = false, edgeAxis = false;
	V3 nFace = v3zero(), nEdgeLocal = v3zero();
	// face normals of A, then of B: an axis separates when the centre distance exceeds the summed projected half sizes
	for (int k = 0; k < 6; k++) {
		const int i = k < 3 ? k : k - 3;
		const float proj = k < 3 ? dA[i] : dot(B.ax[i], d);
		const float reach = k < 3 ? (((A.h[i] + B.h[0] * Q[i][0]) + B.h[1] * Q[i][1]) + B.h[2] * Q[i][2])
		                          : (((A.h[0] * Q[0][i] + A.h[1] * Q[1][i]) + A.h[2] * Q[2][i]) + B.h[i]);
		const float s = fabsf(proj) - reach;
		if (s > 0.f) return false;
		if (s > best) { best = s; nFace = k < 3 ? A.ax[i] : B.ax[i]; edgeAxis = false; flip = proj < 0.f; code = k + 1; }
	}
	// edge x edge axes a_i x b_j, expressed in A's frame; near-parallel edges are kept from winning by a small bias on |R|
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++) Q[i][j] += 1.0e-5f;
	for (int i = 0; i < 3; i++) {
		const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
		const int a0 = i == 0 ? 1 : 0, a1 = i == 2 ? 1 : 2;      // the two axes of A other than i, ascending
		for (int j = 0; j < 3; j++) {
			const int b0 = j == 0 ? 1 : 0, b1 = j == 2 ? 1 : 2;  // the two axes of B other than j, ascending
			const float proj = dA[i2] * R[i1][j] - dA[i1] * R[i2][j];
			const float reach = ((A.h[a0] * Q[a1][j] + A.h[a1] * Q[a0][j]) + B.h[b0] * Q[i][b1]) + B.h[b1] * Q[i][b0];
			float s = fabsf(proj) - reach;
			if (s > kEps) return false;
			float n[3];
			n[i] = 0.f; n[i1] = -R[i2][j]; n[i2] = R[i1][j];
			const float l = sqrtf((n[0] * n[0] + n[1] * n[1]) + n[2] * n[2]);
			if (l > kEps) {
				s /= l;
				if (s * 1.05f > best) {  // an edge axis must beat the best face axis by 5 %
					best = s; edgeAxis = true; flip = proj < 0.f; code = 7 + 3 * i + j;
					nEdgeLocal = V3(n[0] / l, n[1] / l, n[2] / l);
				}
			}
		}
	}
	if (!code) return false;
	V3 nrm = nFace;
	if (edgeAxis) nrm = V3((A.ax[0].x * nEdgeLocal.x + A.ax[1].x * nEdgeLocal.y) + A.ax[2].x * nEdgeLocal.z,
	                       (A.ax[0].y * nEdgeLocal.x + A.ax[1].y * nEdgeLocal.y) + A.ax[2].y * nEdgeLocal.z,
	                       (A.ax[0].z * nEdgeLocal.x + A.ax[1].z * nEdgeLocal.y) + A.ax[2].z * nEdgeLocal.z);
	if (flip) nrm = -nrm;
	out.code = code; out.normal = nrm; out.sep = best;
	return true;
}

// Closest points of the two touching edges (one contact, on B's edge).
RSB_HD V3 obb_edge_contact(const Obb& A, const Obb& B, const SatHit& h) {
	V3 pa = A.c, pb = B.c;
	for (int j = 0; j < 3; j++) {  // walk to the edge of A that faces B ...
		const float sg = dot(h.normal, A.ax[j]) > 0.f ? 1.0f : -1.0f;
		pa = pa + A.ax[j] * (sg * A.h[j]);
	}
	for (int j = 0; j < 3; j++) {  // ... and to the edge of B that faces A
		const float sg = dot(h.normal, B.ax[j]) > 0.f ? -1.0f : 1.0f;
		pb = pb + B.ax[j] * (sg * B.h[j]);
	}
	const V3 ua = A.ax[(h.code - 7) / 3], ub = B.ax[(h.code - 7) % 3];
	// closest approach of the lines pa + alpha ua, pb + beta ub
	const V3 p = pb - pa;
	const float uaub = dot(ua, ub), q1 = dot(ua, p), q2 = -dot(ub, p);
	float det = 1.f - uaub * uaub, beta = 0.f;
	if (!(det <= 0.0001f)) { det = 1.f / det; beta = (uaub * q1 + q2) * det; }
	return pb + ub * beta;
}

// Clip a convex quad against the rectangle |x| <= hx, |y| <= hy: four half-plane passes, at most 8 vertices.
RSB_HD int clip_quad_to_rect(const float half[2], const V2 quad[4], V2 out[8]) {
	V2 cur[8];
	int n = 4, m = 0;
	for (int i = 0; i < 4; i++) cur[i] = quad[i];
	bool full = false;
	for (int pass = 0; pass < 4 && !full; pass++) {
		const int axis = pass >> 1;
		const float sg = (pass & 1) ? 1.f : -1.f;
		V2 nxt[8];
		m = 0;
		for (int i = 0; i < n && !full; i++) {
			const V2 a = cur[i], b = cur[i + 1 < n ? i + 1 : 0];
			const float ad = v2get(a, axis), bd = v2get(b, axis);
			const bool ain = sg * ad < half[axis], bin = sg * bd < half[axis];
			if (ain) { nxt[m++] = a; if (m & 8) { full = true; break; } }
			if (ain != bin) {  // the edge crosses the clip line: emit the crossing
				const float ao = v2get(a, 1 - axis), bo = v2get(b, 1 - axis);
				const float cross_o = ao + (bo - ao) / (bd - ad) * (sg * half[axis] - ad);
				V2 c;
				if (axis) { c.x = cross_o; c.y = sg * half[axis]; } else { c.x = sg * half[axis]; c.y = cross_o; }
				nxt[m++] = c;
				if (m & 8) full = true;
			}
		}
		for (int i = 0; i < m; i++) cur[i] = nxt[i];
		n = m;
	}
	for (int i = 0; i < m; i++) out[i] = cur[i];
	return m;
}

// Choose `want` of the n polygon vertices, spread evenly in angle around the centroid and always containing vertex `first`.
RSB_HD void spread_points(int n, const V2 p[8], int want, int first, int pick[4]) {
	const float M__PI = 3.14159265f;
	float cx, cy;
	if (n == 1) { cx = p[0].x; cy = p[0].y; }
	else if (n == 2) { cx = 0.5f * (p[0].x + p[1].x); cy = 0.5f * (p[0].y + p[1].y); }
	else {
		float a = 0.f, q;
		cx = 0.f; cy = 0.f;
		for (int i = 0; i < n - 1; i++) {
			q = p[i].x * p[i + 1].y - p[i + 1].x * p[i].y;
			a += q;
			cx += q * (p[i].x + p[i + 1].x);
			cy += q * (p[i].y + p[i + 1].y);
		}
		q = p[n - 1].x * p[0].y - p[0].x * p[n - 1].y;
		if (fabsf(a + q) > kEps) a = 1.f / (3.0f * (a + q));
		else a = kLargeFloat;
		cx = a * (cx + q * (p[n - 1].x + p[0].x));
		cy = a * (cy + q * (p[n - 1].y + p[0].y));
	}
	float ang[8];
	unsigned freeMask = 0;
	for (int i = 0; i < n; i++) { ang[i] = rs_atan2f(p[i].y - cy, p[i].x - cx); freeMask |= 1u << i; }
	freeMask &= ~(1u << first);
	pick[0] = first;
	for (int j = 1; j < want; j++) {
		float target = (float)j * (2 * M__PI / want) + ang[first];
		if (target > M__PI) target -= 2 * M__PI;
		float bestDiff = 1e9f;
		int sel = first;
		for (int i = 0; i < n; i++) {
			if (!((freeMask >> i) & 1u)) continue;
			float diff = fabsf(ang[i] - target);
			if (diff > M__PI) diff = 2 * M__PI - diff;
			if (diff < bestDiff) { bestDiff = diff; sel = i; }
		}
		freeMask &= ~(1u << sel);
		pick[j] = sel;
	}
}

// boxes 1/2 = (shape, world transform); the manifold's body0 = owner of box 1.  Pure detector: returns the number of contacts
// (<= 4) and writes the common normal on B, the points and the signed depths; the commit phase adds them in order.
RSB_HD int box_box_detect(const Ctx& cx, const Convex& box1, const Xform& xf1, const Convex& box2, const Xform& xf2, V3& outNormalOnB, V3 outPts[4], float outDep[4]) {
#if defined(RSB_PROFILE)
	if (cx.epa) cx.epa->profCalls += 1;
#endif
	const Obb A = obb_of(box1, xf1), B = obb_of(box2, xf2);
	SatHit h;
	if (!obb_sat(A, B, h)) return 0;
	const float depth = -h.sep;
	outNormalOnB = -h.normal;
	if (h.code > 6) {
		outPts[0] = obb_edge_contact(A, B, h);
		outDep[0] = -depth;
		return 1;
	}
	// face contact: the box owning the separating face is the reference, the other one the incident box
	const bool aIsRef = h.code <= 3;
	const Obb ref = aIsRef ? A : B, inc = aIsRef ? B : A;
	const V3 n = aIsRef ? h.normal : -h.normal;  // out of the reference face
	// incident face = the face of `inc` most anti-parallel to n
	float nr[3], anr[3];
	for (int k = 0; k < 3; k++) { nr[k] = dot(inc.ax[k], n); anr[k] = fabsf(nr[k]); }
	int face, u, v;
	if (anr[1] > anr[0]) { if (anr[1] > anr[2]) { face = 1; u = 0; v = 2; } else { face = 2; u = 0; v = 1; } }
	else                 { if (anr[0] > anr[2]) { face = 0; u = 1; v = 2; } else { face = 2; u = 0; v = 1; } }
	const V3 off = inc.ax[face] * inc.h[face];
	const V3 centre = nr[face] < 0.f ? (inc.c - ref.c) + off : (inc.c - ref.c) - off;  // incident face centre, relative to ref.c
	const int fN = (h.code - 1) % 3;                 // reference face axis; the other two span the face
	const int f1 = fN == 0 ? 1 : 0, f2 = fN == 2 ? 1 : 2;
	const float c1 = dot(centre, ref.ax[f1]), c2 = dot(centre, ref.ax[f2]);
	float m11 = dot(ref.ax[f1], inc.ax[u]), m12 = dot(ref.ax[f1], inc.ax[v]);
	float m21 = dot(ref.ax[f2], inc.ax[u]), m22 = dot(ref.ax[f2], inc.ax[v]);
	V2 quad[4];
	{
		const float k1 = m11 * inc.h[u], k2 = m21 * inc.h[u], k3 = m12 * inc.h[v], k4 = m22 * inc.h[v];
		quad[0].x = c1 - k1 - k3; quad[0].y = c2 - k2 - k4;
		quad[1].x = c1 - k1 + k3; quad[1].y = c2 - k2 + k4;
		quad[2].x = c1 + k1 + k3; quad[2].y = c2 + k2 + k4;
		quad[3].x = c1 + k1 - k3; quad[3].y = c2 + k2 - k4;
	}
	const float half[2] = {ref.h[f1], ref.h[f2]};
	V2 poly[8];
	const int np = clip_quad_to_rect(half, quad, poly);
	if (np < 1) return 0;
	// back to 3-D on the incident face; keep the points that lie below the reference face
	const float det = 1.f / (m11 * m22 - m12 * m21);
	m11 *= det; m12 *= det; m21 *= det; m22 *= det;
	V3 pt[8];
	float dep[8];
	int cnt = 0;
	for (int j = 0; j < np; j++) {
		const float k1 = m22 * (poly[j].x - c1) - m12 * (poly[j].y - c2);
		const float k2 = -m21 * (poly[j].x - c1) + m11 * (poly[j].y - c2);
		pt[cnt] = (centre + inc.ax[u] * k1) + inc.ax[v] * k2;
		dep[cnt] = ref.h[fN] - dot(n, pt[cnt]);
		if (dep[cnt] >= 0.f) { poly[cnt] = poly[j]; cnt++; }
	}
	if (cnt < 1) return 0;
	int pick[4] = {0, 1, 2, 3};
	int nOut = cnt;
	if (cnt > 4) {
		int deepest = 0;
		for (int i = 1; i < cnt; i++) if (dep[i] > dep[deepest]) deepest = i;
		spread_points(cnt, poly, 4, deepest, pick);
		nOut = 4;
	}
	for (int j = 0; j < nOut; j++) {
		const int k = pick[j];
		const V3 pw = pt[k] + ref.c;
		outPts[j] = aIsRef ? pw : pw - h.normal * dep[k];
		outDep[j] = -dep[k];
	}
	return nOut;
}

}  // namespace rsb
