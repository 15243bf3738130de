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

// GJK closest points (+ EPA when the cores overlap) of A/B at their world transforms.  Pure detector: reports the contact
// (normal on B, point on B in world space, signed distance) and leaves the manifold alone -- PH_NARROW_DETECT runs it on one
// thread per candidate, the commit phases add the results in order.
RSB_CC_ATTR bool convex_convex_detect(const Ctx& cx, float breaking, const Convex& A, const Xform& xfA, const Convex& B, const Xform& xfB,
                                      V3& outNormal, V3& outPoint, float& outDist) {
#if defined(RSB_PROFILE)
	if (cx.epa) cx.epa->profCalls += 1;
#endif
	float maxDist2 = A.margin + B.margin + breaking;
	maxDist2 *= maxDist2;

	float distance = 0.f;
	V3 normalInB = v3zero();
	V3 pointOnA = v3zero(), pointOnB = v3zero();
	Xform localA = xfA, localB = xfB;
	V3 positionOffset = (localA.origin + localB.origin) * 0.5f;
	localA.origin -= positionOffset;
	localB.origin -= positionOffset;
	const float marginA = A.margin, marginB = B.margin;
	int curIter = 0;
	const int gGjkMaxIter = 1000;
	V3 cachedSeparatingAxis(0.f, 1.f, 0.f);
	bool isValid = false, checkSimplex = false;
	int degenerateSimplex = 0;
	const float REL_ERROR2 = 1.0e-6f;
	float squaredDistance = kLargeFloat;
	float delta = 0.f;
	const float margin = marginA + marginB;
	Simplex sx;
	sx.reset();
	while (true) {
		V3 sepAxisInA = vmul(-cachedSeparatingAxis, xfA.basis);
		V3 sepAxisInB = vmul(cachedSeparatingAxis, xfB.basis);
		V3 pInA = support_no_margin(A, sepAxisInA);
		V3 qInB = support_no_margin(B, sepAxisInB);
		V3 pWorld = xf_apply(localA, pInA);
		V3 qWorld = xf_apply(localB, qInB);
		V3 w = pWorld - qWorld;
		delta = dot(cachedSeparatingAxis, w);
		if ((delta > 0.f) && (delta * delta > squaredDistance * maxDist2)) { degenerateSimplex = 10; checkSimplex = true; break; }
		if (sx.in_simplex(w)) { degenerateSimplex = 1; checkSimplex = true; break; }
		float f0 = squaredDistance - delta;
		float f1 = squaredDistance * REL_ERROR2;
		if (f0 <= f1) { degenerateSimplex = (f0 <= 0.f) ? 2 : 11; checkSimplex = true; break; }
		sx.add_vertex(w, pWorld, qWorld);
		V3 newAxis;
		if (!simplex_closest(sx, newAxis)) { degenerateSimplex = 3; checkSimplex = true; break; }
		if (len2(newAxis) < REL_ERROR2) { cachedSeparatingAxis = newAxis; degenerateSimplex = 6; checkSimplex = true; break; }
		float previousSquaredDistance = squaredDistance;
		squaredDistance = len2(newAxis);
		if (previousSquaredDistance - squaredDistance <= kEps * previousSquaredDistance) { checkSimplex = true; degenerateSimplex = 12; break; }
		cachedSeparatingAxis = newAxis;
		if (curIter++ > gGjkMaxIter) break;
		if (sx.n == 4) { degenerateSimplex = 13; break; }
	}
	if (checkSimplex) {
		simplex_points(sx, pointOnA, pointOnB);
		normalInB = cachedSeparatingAxis;
		float lenSqr = len2(cachedSeparatingAxis);
		if (lenSqr < 0.0001f) degenerateSimplex = 5;
		if (lenSqr > kEps * kEps) {
			float rlen = 1.f / sqrtf(lenSqr);
			normalInB *= rlen;
			float s = sqrtf(squaredDistance);
			pointOnA -= cachedSeparatingAxis * (marginA / s);
			pointOnB += cachedSeparatingAxis * (marginB / s);
			distance = ((1.f / rlen) - margin);
			isValid = true;
		}
	}
	bool catchDegeneratePenetrationCase = (degenerateSimplex != 0) && ((distance + margin) < 0.01f);
	if (!isValid || catchDegeneratePenetrationCase) {
		V3 tmpPointOnA, tmpPointOnB;
		cachedSeparatingAxis = v3zero();
		bool isValid2 = epa_penetration(cx, A, localA, B, localB, cachedSeparatingAxis, tmpPointOnA, tmpPointOnB);
		if (isValid2) {
			V3 tmpNormalInB = tmpPointOnB - tmpPointOnA;
			float lenSqr = len2(tmpNormalInB);
			if (lenSqr <= (kEps * kEps)) { tmpNormalInB = cachedSeparatingAxis; lenSqr = len2(cachedSeparatingAxis); }
			if (lenSqr > (kEps * kEps)) {
				tmpNormalInB = tmpNormalInB / sqrtf(lenSqr);
				float distance2 = -len(tmpPointOnA - tmpPointOnB);
				if (!isValid || (distance2 < distance)) {
					distance = distance2; pointOnA = tmpPointOnA; pointOnB = tmpPointOnB; normalInB = tmpNormalInB; isValid = true;
				}
			}
		} else {
			if (len2(cachedSeparatingAxis) > 0.f) {
				float distance2 = len(tmpPointOnA - tmpPointOnB) - margin;
				if (!isValid || (distance2 < distance)) {
					distance = distance2; pointOnA = tmpPointOnA; pointOnB = tmpPointOnB;
					pointOnA -= cachedSeparatingAxis * marginA;
					pointOnB += cachedSeparatingAxis * marginB;
					normalInB = normalized(cachedSeparatingAxis);
					isValid = true;
				}
			}
		}
	}
	if (isValid && ((distance < 0.f) || (distance * distance < maxDist2))) {
		V3 posA = convex_aabb_center(cx, A, localA);
		V3 posB = convex_aabb_center(cx, B, localB);
		V3 diff = posA - posB;
		if (dot(diff, normalInB) < 0.f) normalInB *= -1.f;
		outNormal = normalInB; outPoint = pointOnB + positionOffset; outDist = distance;
		return true;
	}
	return false;
}

}  // namespace rsb

#include "rsb_epa.h"
#include "rsb_boxbox.h"
