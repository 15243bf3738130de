// Unit check of the device box-box detector (rsb_boxbox.h box_box_detect) against the reference's btBoxBoxDetector on random
// near-contact box pairs: same number of contacts, bit-identical normal, points and depths, in the same order.
// TEST ONLY: includes the Bullet headers from /root/reference and links oracle/_ref.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <random>
#include <vector>
#include "BulletCollision/CollisionShapes/btBoxShape.h"
#include "BulletCollision/CollisionDispatch/btBoxBoxDetector.h"
#define RSB_HOSTCHECK
#include "sim/rsb_tick.h"

struct Collect : btDiscreteCollisionDetectorInterface::Result {
	std::vector<btVector3> n, p; std::vector<float> d;
	void setShapeIdentifiersA(int, int) override {}
	void setShapeIdentifiersB(int, int) override {}
	void addContactPoint(const btVector3& normalOnBInWorld, const btVector3& pointInWorld, btScalar depth) override { n.push_back(normalOnBInWorld); p.push_back(pointInWorld); d.push_back(depth); }
};
static bool same(float a, float b) { return memcmp(&a, &b, 4) == 0; }

int main(int argc, char** argv) {
	const int N = argc > 1 ? atoi(argv[1]) : 200000;
	std::mt19937 rng(777);
	std::uniform_real_distribution<float> U(-1.f, 1.f);
	long hits = 0, bad = 0, edge = 0, multi = 0, culled = 0;
	for (int it = 0; it < N; it++) {
		btVector3 he1(1.2f + 0.3f * U(rng), 0.85f + 0.2f * U(rng), 0.4f + 0.1f * U(rng)), he2(1.2f + 0.3f * U(rng), 0.85f + 0.2f * U(rng), 0.4f + 0.1f * U(rng));
		btBoxShape b1(he1), b2(he2);
		btQuaternion q1(U(rng), U(rng), U(rng), U(rng)), q2(U(rng), U(rng), U(rng), U(rng));
		if ((it % 3) == 0) { q1 = btQuaternion(btVector3(0, 0, 1), 3.14159f * U(rng)); q2 = btQuaternion(btVector3(0, 0, 1), 3.14159f * U(rng)); }   // flat on the ground: face contacts
		if ((it % 3) == 1) { q2 = q1 * btQuaternion(btVector3(U(rng), U(rng), U(rng)).normalized(), 0.05f * U(rng)); }                               // nearly aligned: >4 clip points
		q1.normalize(); q2.normalize();
		btTransform x1(q1, btVector3(U(rng) * 2, U(rng) * 2, U(rng) * 2));
		btVector3 dir(U(rng), U(rng), (it % 3) == 0 ? 0.3f * U(rng) : U(rng)); dir.normalize();
		btTransform x2(q2, x1.getOrigin() + dir * (0.8f + 1.7f * (U(rng) + 1.f) * 0.5f));
		Collect res;
		btBoxBoxDetector det(&b1, &b2);
		btDiscreteCollisionDetectorInterface::ClosestPointInput in;
		in.m_maximumDistanceSquared = BT_LARGE_FLOAT; in.m_transformA = x1; in.m_transformB = x2;
		det.getClosestPoints(in, res, false);
		auto conv = [](const btBoxShape& b) { rsb::Convex c; btVector3 h = b.getHalfExtentsWithoutMargin(); c.kind = rsb::SH_BOX; c.margin = b.getMargin(); c.he = rsb::V3(h.x(), h.y(), h.z()); c.pts = nullptr; c.npts = 0; return c; };
		auto xf = [](const btTransform& t) { rsb::Xform x; const btMatrix3x3& B = t.getBasis(); x.basis = rsb::M3(rsb::V3(B[0][0], B[0][1], B[0][2]), rsb::V3(B[1][0], B[1][1], B[1][2]), rsb::V3(B[2][0], B[2][1], B[2][2])); x.origin = rsb::V3(t.getOrigin().x(), t.getOrigin().y(), t.getOrigin().z()); return x; };
		rsb::Ctx cx; memset(&cx, 0, sizeof(cx));
		rsb::V3 nrm, pts[4]; float dep[4];
		int n = rsb::box_box_detect(cx, conv(b1), xf(x1), conv(b2), xf(x2), nrm, pts, dep);
		bool ok = n == (int)res.d.size();
		for (int k = 0; ok && k < n; k++)
			ok = same(nrm.x, res.n[k].x()) && same(nrm.y, res.n[k].y()) && same(nrm.z, res.n[k].z()) && same(pts[k].x, res.p[k].x()) && same(pts[k].y, res.p[k].y()) && same(pts[k].z, res.p[k].z()) && same(dep[k], res.d[k]);
		if (!ok) { if (bad < 5) printf("case %d: mine %d contacts, reference %zu\n", it, n, res.d.size()); bad++; }
		if (n) hits++;
		if (n == 1) edge++;
		if (n > 1) multi++;
		if (n == 4) culled++;
	}
	printf("cases %d with_contact %ld single %ld multi %ld four %ld mismatches %ld\n", N, hits, edge, multi, culled, bad);
	return 0;
}
