// rsb_gjk.h -- convex-convex narrowphase: GJK closest points with margin handling, the
// EPA penetration fallback, and the box-box (SAT + clipping) detector.  Restates
//   btConvexConvexAlgorithm::processCollision    bt/BulletCollision/CollisionDispatch/btConvexConvexAlgorithm.cpp:268-590
//   btGjkPairDetector::getClosestPointsNonVirtual bt/BulletCollision/NarrowPhaseCollision/btGjkPairDetector.cpp:686-959
//   btGjkEpaSolver2::Penetration (GJK+EPA)        bt/BulletCollision/NarrowPhaseCollision/btGjkEpa2.cpp
//   btGjkEpaPenetrationDepthSolver::calcPenDepth  bt/BulletCollision/NarrowPhaseCollision/btGjkEpaPenetrationDepthSolver.cpp:23-79
//   btBoxBoxDetector (ODE dBoxBox2)               bt/BulletCollision/CollisionDispatch/btBoxBoxDetector.cpp:272-767
#pragma once
#include "rsb_work.h"
#include "rsb_simplex.h"

namespace rsb {

RSB_ACP_ATTR void add_contact_point(const Ctx& cx, int mi, const V3& normalOnB, const V3& pointInWorld, float depth, int triMesh, int triIndex, bool dynPair);
RSB_HD void transform_aabb_he(const V3& he, float margin, const M3& basis, const V3& origin, V3& mn, V3& mx);
RSB_HD void transform_aabb_mm(const V3& lmin, const V3& lmax, float margin, const M3& basis, const V3& origin, V3& mn, V3& mx);

// shape->getAabb(t) as used by the GJK normal-direction fix (only the centre is consumed)
RSB_HD V3 convex_aabb_center(const Ctx& cx, const Convex& s, const Xform& t) {
	V3 mn, mx;
	switch (s.kind) {
	case SH_SPHERE: { float mg = s.margin + 0.08f; V3 e(mg, mg, mg); mn = t.origin - e; mx = t.origin + e; break; }
	case SH_BOX: transform_aabb_he(s.he, s.margin, t.basis, t.origin, mn, mx); break;
	case SH_TRI: {
		// btConvexInternalShape::getAabbSlow (btConvexInternalShape.cpp:51-73)
		float margin = s.margin;
		for (int i = 0; i < 3; i++) {
			V3 vec(0.f, 0.f, 0.f); vec.set(i, 1.f);
			V3 sv = support_with_margin(s, vmul(vec, t.basis));
			V3 tmp = xf_apply(t, sv);
			mx.set(i, tmp[i] + margin);
			vec.set(i, -1.f);
			tmp = xf_apply(t, support_with_margin(s, vmul(vec, t.basis)));
			mn.set(i, tmp[i] - margin);
		}
		break;
	}
	default: transform_aabb_mm(cx.p->puckLocalAabbMin, cx.p->puckLocalAabbMax, s.margin, t.basis, t.origin, mn, mx); break;
	}
	return (mx + mn) * 0.5f;
}

// (out-of-line functions take their read-only aggregates BY VALUE: with nvcc 12.9 a caller-local object whose last use
// is being passed by reference to a __noinline__ callee can share its stack slot with the callee's own locals)
RSB_HD_NOINLINE bool epa_penetration(const Ctx& cx, Convex A, Xform tA, Convex B, Xform tB, V3& v, V3& wa, V3& wb);

// ---- distance query between two convex cores ----------------------------------------------------------------------------
// The reference's pair detector (btGjkPairDetector::getClosestPointsNonVirtual, btGjkPairDetector.cpp:686-959) is one 270-line
// loop-and-flags routine.  Here the same decisions are three stages with explicit results, so that each stage can be read --
// and later scheduled -- on its own:
//   gjk_descend   the simplex iteration; reports WHY it stopped (GjkStop) and the last separating axis;
//   gjk_surface   shrinks the core-to-core segment by the two margins -> signed surface distance;
//   the caller    decides from GjkStop + distance whether the penetration solver (EPA) has to have a say.
// The stop reasons keep the reference's meaning (its numeric "degenerate simplex" codes are named instead of numbered); the
// arithmetic and the order of the exit tests are the reference's, so the contacts come out bit for bit.
enum GjkStop : int {
	GJK_RUNNING = 0,
	GJK_SEPARATED,       // the support plane already proves a gap wider than the contact range
	GJK_REPEAT_VERTEX,   // the new support point is already in the simplex
	GJK_NO_PROGRESS,     // the support point does not get closer than the current distance (relative 1e-6) ...
	GJK_TOUCHING,        // ... and not at all: the cores touch or overlap
	GJK_BAD_SIMPLEX,     // the Voronoi solver rejected the simplex
	GJK_TINY_AXIS,       // the closest vector collapsed (|v|^2 < 1e-6): origin reached
	GJK_STALLED,         // the squared distance stopped decreasing (relative FLT_EPSILON)
	GJK_FULL_SIMPLEX,    // four vertices: the origin is enclosed
	GJK_ITER_LIMIT
};

struct GjkState {
	Simplex sx;
	V3 axis;        // last accepted closest vector (from B towards A)
	float dist2;    // its squared length at acceptance time
	GjkStop stop;
	bool usable;    // the simplex may be read for closest points (every stop except the enclosed-origin / iteration-limit ones)
};

RSB_HD void gjk_descend(const Convex& A, const Xform& rotA, const Xform& locA, const Convex& B, const Xform& rotB, const Xform& locB,
                        float reach2, GjkState& g) {
	const float kRel2 = 1.0e-6f;
	g.sx.reset();
	g.axis = V3(0.f, 1.f, 0.f);
	g.dist2 = kLargeFloat;
	g.stop = GJK_RUNNING;
	g.usable = true;
	for (int iter = 0;;) {
		// support point of the Minkowski difference A - B against the current axis
		const V3 pw = xf_apply(locA, support_no_margin(A, vmul(-g.axis, rotA.basis)));
		const V3 qw = xf_apply(locB, support_no_margin(B, vmul(g.axis, rotB.basis)));
		const V3 w = pw - qw;
		const float along = dot(g.axis, w);
		if (along > 0.f && along * along > g.dist2 * reach2) { g.stop = GJK_SEPARATED; return; }
		if (g.sx.in_simplex(w)) { g.stop = GJK_REPEAT_VERTEX; return; }
		const float gain = g.dist2 - along;
		if (gain <= g.dist2 * kRel2) { g.stop = gain <= 0.f ? GJK_TOUCHING : GJK_NO_PROGRESS; return; }
		g.sx.add_vertex(w, pw, qw);
		V3 v;
		if (!simplex_closest(g.sx, v)) { g.stop = GJK_BAD_SIMPLEX; return; }
		if (len2(v) < kRel2) { g.axis = v; g.stop = GJK_TINY_AXIS; return; }
		const float before = g.dist2;
		g.dist2 = len2(v);
		if (before - g.dist2 <= kEps * before) { g.stop = GJK_STALLED; return; }
		g.axis = v;
		if (iter++ > 1000) { g.stop = GJK_ITER_LIMIT; g.usable = false; return; }
		if (g.sx.n == 4) { g.stop = GJK_FULL_SIMPLEX; g.usable = false; return; }
	}
}

struct SurfaceHit { V3 onA, onB, normal; float dist; bool valid; bool shallowSuspect; };

// closest points of the cores -> closest points of the surfaces (cores grown by their margins)
RSB_HD SurfaceHit gjk_surface(GjkState& g, float marginA, float marginB) {
	SurfaceHit h;
	h.onA = v3zero(); h.onB = v3zero(); h.normal = v3zero(); h.dist = 0.f; h.valid = false;
	bool degenerate = g.stop != GJK_RUNNING;  // every stop the loop reports counts as a degenerate simplex for the fallback rule
	if (g.usable) {
		simplex_points(g.sx, h.onA, h.onB);
		h.normal = g.axis;
		const float l2 = len2(g.axis);
		if (l2 < 0.0001f) degenerate = true;
		if (l2 > kEps * kEps) {
			const float rl = 1.f / sqrtf(l2);
			h.normal *= rl;
			const float s = sqrtf(g.dist2);
			h.onA -= g.axis * (marginA / s);
			h.onB += g.axis * (marginB / s);
			h.dist = (1.f / rl) - (marginA + marginB);
			h.valid = true;
		}
	}
	// a degenerate stop with the surfaces within a centimetre of touching is re-examined by the penetration solver
	h.shallowSuspect = degenerate && ((h.dist + (marginA + marginB)) < 0.01f);
	return h;
}

// GJK closest points (+ EPA when the cores overlap) of A/B at their world transforms.  Pure detector: reports the contact
// (normal on B, point on B in world space, signed distance) and leaves the manifold alone -- PH_NARROW_DETECT runs it on one
// thread per candidate, the commit phases add the results in order.
RSB_CC_ATTR bool convex_convex_detect(const Ctx& cx, float breaking, const Convex& A, const Xform& xfA, const Convex& B, const Xform& xfB,
                                      V3& outNormal, V3& outPoint, float& outDist) {
#if defined(RSB_PROFILE)
	if (cx.epa) cx.epa->profCalls += 1;
#endif
	float reach2 = A.margin + B.margin + breaking;
	reach2 *= reach2;
	// work around the midpoint of the two origins (keeps the simplex arithmetic small-valued, as the reference does)
	const V3 mid = (xfA.origin + xfB.origin) * 0.5f;
	Xform locA = xfA, locB = xfB;
	locA.origin -= mid;
	locB.origin -= mid;

	GjkState g;
	gjk_descend(A, xfA, locA, B, xfB, locB, reach2, g);
	SurfaceHit h = gjk_surface(g, A.margin, B.margin);

	if (!h.valid || h.shallowSuspect) {
		V3 sep = v3zero(), pa, pb;
		const float margin = A.margin + B.margin;
		if (epa_penetration(cx, A, locA, B, locB, sep, pa, pb)) {
			// penetration found: direction from the two witness points (or the solver's axis when they coincide)
			V3 n = pb - pa;
			float l2 = len2(n);
			if (l2 <= kEps * kEps) { n = sep; l2 = len2(sep); }
			if (l2 > kEps * kEps) {
				n = n / sqrtf(l2);
				const float d = -len(pa - pb);
				if (!h.valid || d < h.dist) { h.dist = d; h.onA = pa; h.onB = pb; h.normal = n; h.valid = true; }
			}
		} else if (len2(sep) > 0.f) {
			// no penetration, but the solver's GJK pass produced a separating axis: distance between its witness points
			const float d = len(pa - pb) - margin;
			if (!h.valid || d < h.dist) {
				h.dist = d;
				h.onA = pa - sep * A.margin;
				h.onB = pb + sep * B.margin;
				h.normal = normalized(sep);
				h.valid = true;
			}
		}
	}
	if (!(h.valid && (h.dist < 0.f || h.dist * h.dist < reach2))) return false;
	// the normal must point from B to A: compare with the offset of the two bounding-box centres
	const V3 between = convex_aabb_center(cx, A, locA) - convex_aabb_center(cx, B, locB);
	if (dot(between, h.normal) < 0.f) h.normal *= -1.f;
	outNormal = h.normal; outPoint = h.onB + mid; outDist = h.dist;
	return true;
}

}  // namespace rsb

#include "rsb_epa.h"
#include "rsb_boxbox.h"
