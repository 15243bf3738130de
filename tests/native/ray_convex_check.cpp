// Unit check of the device ray-vs-convex cast (rsb_ray.h ray_convex) against the reference's
// btSubsimplexConvexCast driven exactly like btCollisionWorld::rayTestSingleInternal does
// (bt/BulletCollision/CollisionDispatch/btCollisionWorld.cpp:277-311).  TEST ONLY: links oracle/_ref.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <random>
#include "BulletCollision/CollisionShapes/btBoxShape.h"
#include "BulletCollision/CollisionShapes/btSphereShape.h"
#include "BulletCollision/NarrowPhaseCollision/btSubSimplexConvexCast.h"
#include "BulletCollision/NarrowPhaseCollision/btVoronoiSimplexSolver.h"
#define RSB_HOSTCHECK
#include "sim/rsb_tick.h"

int main(int argc, char** argv) {
	int N = argc > 1 ? atoi(argv[1]) : 20000;
	int verbose = argc > 2 ? atoi(argv[2]) : 0;
	std::mt19937 rng(12345);
	std::uniform_real_distribution<float> U(-1.f, 1.f);
	int hitsRef = 0, hitsMine = 0, mismatchHit = 0, loose = 0; double worstFrac = 0, worstN = 0;
	for (int it = 0; it < N; it++) {
		bool sphere = (it % 4) == 3;
		btVector3 he(1.2f, 0.87f, 0.39f);
		btBoxShape box(he);
		btSphereShape ball(1.825f);
		btConvexShape* shp = sphere ? (btConvexShape*)&ball : (btConvexShape*)&box;
		btQuaternion q(U(rng), U(rng), U(rng), U(rng)); q.normalize();
		btTransform xf(q, btVector3(U(rng) * 3, U(rng) * 3, U(rng) * 3));
		// ray: start 0.5..2.5 from the centre, pointing roughly at it, length 1.3
		btVector3 dir(U(rng), U(rng), U(rng)); dir.normalize();
		float d0 = 1.0f + 0.9f * (U(rng) + 1.f);
		btVector3 from = xf.getOrigin() + dir * d0;
		btVector3 aim = xf.getOrigin() + btVector3(U(rng), U(rng), U(rng)) * 0.8f;
		btVector3 to = from + (aim - from).normalized() * 1.3f;
		// reference
		btSphereShape pointShape(btScalar(0.0)); pointShape.setMargin(0.f);
		btTransform rayFrom; rayFrom.setIdentity(); rayFrom.setOrigin(from);
		btTransform rayTo; rayTo.setIdentity(); rayTo.setOrigin(to);
		btConvexCast::CastResult cr; cr.m_fraction = 1.f;
		btVoronoiSimplexSolver ss;
		btSubsimplexConvexCast caster(&pointShape, shp, &ss);
		bool hr = caster.calcTimeOfImpact(rayFrom, rayTo, xf, xf, cr) && cr.m_normal.length2() > 0.0001f && cr.m_fraction < 1.f;
		btVector3 nr = cr.m_normal; if (hr) nr.normalize();
		// mine
		rsb::Convex c;
		if (sphere) { c.kind = rsb::SH_SPHERE; c.margin = 1.825f; c.he = rsb::v3zero(); }
		else { c.kind = rsb::SH_BOX; c.margin = box.getMargin(); c.he = rsb::V3(he.x() - 0.04f, he.y() - 0.04f, he.z() - 0.04f); }
		c.pts = nullptr; c.npts = 0;
		rsb::Xform mx;
		const btMatrix3x3& B = xf.getBasis();
		mx.basis = rsb::M3(rsb::V3(B[0][0], B[0][1], B[0][2]), rsb::V3(B[1][0], B[1][1], B[1][2]), rsb::V3(B[2][0], B[2][1], B[2][2]));
		mx.origin = rsb::V3(xf.getOrigin().x(), xf.getOrigin().y(), xf.getOrigin().z());
		float frac = 1.f; rsb::V3 n = rsb::v3zero();
		bool hm = rsb::ray_convex(c, mx, rsb::V3(from.x(), from.y(), from.z()), rsb::V3(to.x(), to.y(), to.z()), frac, n);
		hitsRef += hr; hitsMine += hm;
		if (hr != hm) { mismatchHit++; if (verbose) printf("case %d: hit mismatch ref %d mine %d\n", it, hr, hm); continue; }
		if (hr) {
			double df = fabs(frac - cr.m_fraction);
			double dn = fmax(fmax(fabs(n.x - nr.x()), fabs(n.y - nr.y())), fabs(n.z - nr.z()));
			if (verbose && (df > 1e-6 || dn > 1e-5)) printf("case %d (%s): frac %.8f vs %.8f  n (%.6f %.6f %.6f) vs (%.6f %.6f %.6f)\n", it, sphere ? "sphere" : "box", frac, cr.m_fraction, n.x, n.y, n.z, nr.x(), nr.y(), nr.z());
			if (df > 1e-3) loose++;
			if (df > worstFrac) worstFrac = df;
			if (dn > worstN) worstN = dn;
		}
	}
	printf("cases %d hits_ref %d hits_mine %d hit_mismatch %d fraction_diff_gt_1e-3 %d worst_fraction_diff %.3g worst_normal_diff %.3g\n", N, hitsRef, hitsMine, mismatchHit, loose, worstFrac, worstN);
	return 0;
}
