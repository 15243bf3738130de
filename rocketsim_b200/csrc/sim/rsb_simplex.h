// rsb_simplex.h -- closest-point-to-origin on a 1..4 vertex simplex (Voronoi regions) and the
// support mappings of the convex shapes in an arena.  Restates
//   bt/BulletCollision/NarrowPhaseCollision/btVoronoiSimplexSolver.cpp:33-580
//   support mappings: btBoxShape.h (localGetSupportingVertex[WithoutMargin]),
//   btSphereShape.cpp:29-49, btConvexHullShape.cpp (max-dot scan), btTriangleShape.h
// in a compact value-type form (no virtual interfaces), because GJK, the sub-simplex ray
// cast and EPA all build on it and results are iteration-path dependent.
#pragma once
#include "rsb_math.h"

namespace rsb {

constexpr float kLargeFloat = 1e18f;  // BT_LARGE_FLOAT

struct Simplex {
	int n;
	V3 W[5], P[5], Q[5];
	V3 cachedP1, cachedP2, cachedV, lastW;
	float bc[4];
	unsigned used;  // bit0..3 = usedVertexA..D
	bool degenerate, cachedValid, needsUpdate;

	RSB_HD void reset() {
		cachedValid = false; n = 0; needsUpdate = true;
		lastW = V3(kLargeFloat, kLargeFloat, kLargeFloat);
		bc_reset();
	}
	RSB_HD void bc_reset() { degenerate = false; bc[0] = bc[1] = bc[2] = bc[3] = 0.f; used = 0; }
	RSB_HD void set_bc(float a, float b = 0.f, float c = 0.f, float d = 0.f) { bc[0] = a; bc[1] = b; bc[2] = c; bc[3] = d; }
	RSB_HD bool bc_valid() const { return bc[0] >= 0.f && bc[1] >= 0.f && bc[2] >= 0.f && bc[3] >= 0.f; }
	RSB_HD void add_vertex(const V3& w, const V3& p, const V3& q) {
		lastW = w; needsUpdate = true;
		W[n] = w; P[n] = p; Q[n] = q; n++;
	}
	RSB_HD void remove_vertex(int i) { n--; W[i] = W[n]; P[i] = P[n]; Q[i] = Q[n]; }
	RSB_HD void reduce(unsigned u) {
		if (n >= 4 && !(u & 8)) remove_vertex(3);
		if (n >= 3 && !(u & 4)) remove_vertex(2);
		if (n >= 2 && !(u & 2)) remove_vertex(1);
		if (n >= 1 && !(u & 1)) remove_vertex(0);
	}
	RSB_HD bool in_simplex(const V3& w) const {
		bool found = false;
		for (int i = 0; i < n; i++) {
			if (len2(W[i] - w) <= 0.0001f) { found = true; break; }  // BT_USE_EQUAL_VERTEX_THRESHOLD
		}
		if (w.x == lastW.x && w.y == lastW.y && w.z == lastW.z) return true;
		return found;
	}
	RSB_HD float max_vertex() const {
		float m = 0.f;
		for (int i = 0; i < n; i++) { float l = len2(W[i]); if (m < l) m = l; }
		return m;
	}
};

struct SubResult { V3 closest; unsigned used; float bc[4]; bool degenerate; };

// closestPtPointTriangle (btVoronoiSimplexSolver.cpp:314-404)
RSB_HD void closest_pt_point_triangle(const V3& p, const V3& a, const V3& b, const V3& c, SubResult& r) {
	r.used = 0;
	V3 ab = b - a, ac = c - a, ap = p - a;
	float d1 = dot(ab, ap), d2 = dot(ac, ap);
	if (d1 <= 0.f && d2 <= 0.f) { r.closest = a; r.used = 1; r.bc[0] = 1; r.bc[1] = 0; r.bc[2] = 0; r.bc[3] = 0; return; }
	V3 bp = p - b;
	float d3 = dot(ab, bp), d4 = dot(ac, bp);
	if (d3 >= 0.f && d4 <= d3) { r.closest = b; r.used = 2; r.bc[0] = 0; r.bc[1] = 1; r.bc[2] = 0; r.bc[3] = 0; return; }
	float vc = d1 * d4 - d3 * d2;
	if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) {
		float v = d1 / (d1 - d3);
		r.closest = a + v * ab; r.used = 3; r.bc[0] = 1 - v; r.bc[1] = v; r.bc[2] = 0; r.bc[3] = 0; return;
	}
	V3 cp = p - c;
	float d5 = dot(ab, cp), d6 = dot(ac, cp);
	if (d6 >= 0.f && d5 <= d6) { r.closest = c; r.used = 4; r.bc[0] = 0; r.bc[1] = 0; r.bc[2] = 1; r.bc[3] = 0; return; }
	float vb = d5 * d2 - d1 * d6;
	if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) {
		float w = d2 / (d2 - d6);
		r.closest = a + w * ac; r.used = 5; r.bc[0] = 1 - w; r.bc[1] = 0; r.bc[2] = w; r.bc[3] = 0; return;
	}
	float va = d3 * d6 - d5 * d4;
	if (va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f) {
		float w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
		r.closest = b + w * (c - b); r.used = 6; r.bc[0] = 0; r.bc[1] = 1 - w; r.bc[2] = w; r.bc[3] = 0; return;
	}
	float denom = 1.0f / (va + vb + vc);
	float v = vb * denom, w = vc * denom;
	r.closest = a + ab * v + ac * w; r.used = 7;
	r.bc[0] = 1 - v - w; r.bc[1] = v; r.bc[2] = w; r.bc[3] = 0;
}

RSB_HD int point_outside_of_plane(const V3& p, const V3& a, const V3& b, const V3& c, const V3& d) {
	V3 normal = cross(b - a, c - a);
	float signp = dot(p - a, normal);
	float signd = dot(d - a, normal);
	if (signd * signd < (1e-4f * 1e-4f)) return -1;
	return (signp * signd < 0.f) ? 1 : 0;
}

// closestPtPointTetrahedron (btVoronoiSimplexSolver.cpp:430-580)
RSB_HD bool closest_pt_point_tetrahedron(const V3& p, const V3& a, const V3& b, const V3& c, const V3& d, SubResult& fin) {
	SubResult t;
	fin.closest = p; fin.used = 15;
	int oABC = point_outside_of_plane(p, a, b, c, d);
	int oACD = point_outside_of_plane(p, a, c, d, b);
	int oADB = point_outside_of_plane(p, a, d, b, c);
	int oBDC = point_outside_of_plane(p, b, d, c, a);
	if (oABC < 0 || oACD < 0 || oADB < 0 || oBDC < 0) { fin.degenerate = true; return false; }
	if (!oABC && !oACD && !oADB && !oBDC) return false;
	float best = 3.402823466e+38f;
	if (oABC) {
		closest_pt_point_triangle(p, a, b, c, t);
		V3 q = t.closest; float sq = dot(q - p, q - p);
		if (sq < best) {
			best = sq; fin.closest = q;
			fin.used = ((t.used & 1) ? 1 : 0) | ((t.used & 2) ? 2 : 0) | ((t.used & 4) ? 4 : 0);
			fin.bc[0] = t.bc[0]; fin.bc[1] = t.bc[1]; fin.bc[2] = t.bc[2]; fin.bc[3] = 0;
		}
	}
	if (oACD) {
		closest_pt_point_triangle(p, a, c, d, t);
		V3 q = t.closest; float sq = dot(q - p, q - p);
		if (sq < best) {
			best = sq; fin.closest = q;
			fin.used = ((t.used & 1) ? 1 : 0) | ((t.used & 2) ? 4 : 0) | ((t.used & 4) ? 8 : 0);
			fin.bc[0] = t.bc[0]; fin.bc[1] = 0; fin.bc[2] = t.bc[1]; fin.bc[3] = t.bc[2];
		}
	}
	if (oADB) {
		closest_pt_point_triangle(p, a, d, b, t);
		V3 q = t.closest; float sq = dot(q - p, q - p);
		if (sq < best) {
			best = sq; fin.closest = q;
			fin.used = ((t.used & 1) ? 1 : 0) | ((t.used & 4) ? 2 : 0) | ((t.used & 2) ? 8 : 0);
			fin.bc[0] = t.bc[0]; fin.bc[1] = t.bc[2]; fin.bc[2] = 0; fin.bc[3] = t.bc[1];
		}
	}
	if (oBDC) {
		closest_pt_point_triangle(p, b, d, c, t);
		V3 q = t.closest; float sq = dot(q - p, q - p);
		if (sq < best) {
			best = sq; fin.closest = q;
			fin.used = ((t.used & 1) ? 2 : 0) | ((t.used & 4) ? 4 : 0) | ((t.used & 2) ? 8 : 0);
			fin.bc[0] = 0; fin.bc[1] = t.bc[0]; fin.bc[2] = t.bc[2]; fin.bc[3] = t.bc[1];
		}
	}
	return true;
}

// updateClosestVectorAndPoints (btVoronoiSimplexSolver.cpp:83-232)
RSB_HD bool simplex_update(Simplex& s) {
	if (!s.needsUpdate) return s.cachedValid;
	s.bc_reset();
	s.needsUpdate = false;
	switch (s.n) {
	case 0: s.cachedValid = false; break;
	case 1:
		s.cachedP1 = s.P[0]; s.cachedP2 = s.Q[0];
		s.cachedV = s.cachedP1 - s.cachedP2;
		s.bc_reset(); s.set_bc(1.f, 0.f, 0.f, 0.f);
		s.cachedValid = s.bc_valid();
		break;
	case 2: {
		const V3 from = s.W[0], to = s.W[1];
		V3 diff = v3zero() - from;
		V3 v = to - from;
		float t = dot(v, diff);
		if (t > 0.f) {
			float dotVV = dot(v, v);
			if (t < dotVV) { t /= dotVV; diff -= t * v; s.used |= 3; }
			else { t = 1.f; diff -= v; s.used |= 2; }
		} else { t = 0.f; s.used |= 1; }
		s.set_bc(1 - t, t);
		s.cachedP1 = s.P[0] + t * (s.P[1] - s.P[0]);
		s.cachedP2 = s.Q[0] + t * (s.Q[1] - s.Q[0]);
		s.cachedV = s.cachedP1 - s.cachedP2;
		s.reduce(s.used);
		s.cachedValid = s.bc_valid();
		break;
	}
	case 3: {
		SubResult r; r.degenerate = false;
		closest_pt_point_triangle(v3zero(), s.W[0], s.W[1], s.W[2], r);
		s.used = r.used; s.set_bc(r.bc[0], r.bc[1], r.bc[2], r.bc[3]);
		s.cachedP1 = s.P[0] * s.bc[0] + s.P[1] * s.bc[1] + s.P[2] * s.bc[2];
		s.cachedP2 = s.Q[0] * s.bc[0] + s.Q[1] * s.bc[1] + s.Q[2] * s.bc[2];
		s.cachedV = s.cachedP1 - s.cachedP2;
		s.reduce(s.used);
		s.cachedValid = s.bc_valid();
		break;
	}
	case 4: {
		SubResult r; r.degenerate = false; r.bc[0] = r.bc[1] = r.bc[2] = r.bc[3] = 0.f; r.used = 0;
		bool hasSeparation = closest_pt_point_tetrahedron(v3zero(), s.W[0], s.W[1], s.W[2], s.W[3], r);
		s.used = r.used; s.set_bc(r.bc[0], r.bc[1], r.bc[2], r.bc[3]); s.degenerate = r.degenerate;
		if (hasSeparation) {
			s.cachedP1 = s.P[0] * s.bc[0] + s.P[1] * s.bc[1] + s.P[2] * s.bc[2] + s.P[3] * s.bc[3];
			s.cachedP2 = s.Q[0] * s.bc[0] + s.Q[1] * s.bc[1] + s.Q[2] * s.bc[2] + s.Q[3] * s.bc[3];
			s.cachedV = s.cachedP1 - s.cachedP2;
			s.reduce(s.used);
		} else {
			if (s.degenerate) s.cachedValid = false;
			else { s.cachedValid = true; s.cachedV = v3zero(); }
			break;
		}
		s.cachedValid = s.bc_valid();
		break;
	}
	default: s.cachedValid = false;
	}
	return s.cachedValid;
}
RSB_HD bool simplex_closest(Simplex& s, V3& v) { bool ok = simplex_update(s); v = s.cachedV; return ok; }
RSB_HD void simplex_points(Simplex& s, V3& p1, V3& p2) { simplex_update(s); p1 = s.cachedP1; p2 = s.cachedP2; }

// ----------------------------------------------------------------------------------------
// Convex shapes (value types).  kind: 0 point (radius-0 sphere), 1 sphere, 2 box, 3 triangle, 4 hull
enum ShapeKind : int { SH_POINT = 0, SH_SPHERE = 1, SH_BOX = 2, SH_TRI = 3, SH_HULL = 4 };

struct Convex {
	int kind;
	float margin;       // collision margin (sphere: radius)
	V3 he;              // box: implicit half extents (WITHOUT margin)
	V3 tri[3];          // triangle vertices (local)
	const V3* pts;      // hull points
	int npts;
};

// localGetSupportVertexWithoutMarginNonVirtual (btConvexShape.cpp:117-243)
RSB_HD V3 support_no_margin(const Convex& c, const V3& v) {
	switch (c.kind) {
	case SH_POINT:
	case SH_SPHERE: return v3zero();
	case SH_BOX: return V3(v.x >= 0.f ? c.he.x : -c.he.x, v.y >= 0.f ? c.he.y : -c.he.y, v.z >= 0.f ? c.he.z : -c.he.z);
	case SH_TRI: {
		V3 dots(dot(v, c.tri[0]), dot(v, c.tri[1]), dot(v, c.tri[2]));
		// btVector3::maxAxis: x<y ? (y<z?2:1) : (x<z?2:0)
		int i = dots.x < dots.y ? (dots.y < dots.z ? 2 : 1) : (dots.x < dots.z ? 2 : 0);
		return c.tri[i];
	}
	default: {
		// btVector3::maxDot: first index of the maximum dot product
		float best = -3.402823466e+38f; int bi = -1;
		for (int i = 0; i < c.npts; i++) { float d = dot(c.pts[i], v); if (d > best) { best = d; bi = i; } }
		return bi >= 0 ? c.pts[bi] : v3zero();
	}
	}
}

// localGetSupportingVertex (WITH margin) as used by the convex cast and EPA's support
RSB_HD V3 support_with_margin(const Convex& c, const V3& v) {
	switch (c.kind) {
	case SH_POINT: return v3zero();
	case SH_SPHERE: {
		V3 n = v;
		if (len2(n) < kEps * kEps) n = normalized(V3(-1.f, -1.f, -1.f));
		else n = normalized(n);
		return c.margin * n;
	}
	case SH_BOX: {
		V3 h(c.he.x + c.margin, c.he.y + c.margin, c.he.z + c.margin);
		return V3(v.x >= 0.f ? h.x : -h.x, v.y >= 0.f ? h.y : -h.y, v.z >= 0.f ? h.z : -h.z);
	}
	default: {
		V3 s = support_no_margin(c, v);
		if (c.margin != 0.f) {
			V3 n = v;
			if (len2(n) < kEps * kEps) n = V3(-1.f, -1.f, -1.f);
			n = normalized(n);
			s += c.margin * n;
		}
		return s;
	}
	}
}

}  // namespace rsb
