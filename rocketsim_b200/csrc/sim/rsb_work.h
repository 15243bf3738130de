// rsb_work.h -- per-tick scratch (ArenaW) and the execution context of one arena.
#pragma once
#include "rsb_state.h"
#include "rsb_consts.h"

namespace rsb {

// Object ids used inside a tick: dynamic bodies 0..nBodies-1 (0 = ball, 1+i = car i);
// static objects are STATIC_BASE + index (meshes first, then planes) -- the reference's
// world object order (Arena.cpp:480-496, 1004-1068).
constexpr int OBJ_NONE = -1;
constexpr int STATIC_BASE = 64;

enum Activation : int { ACT_ACTIVE = 1, ACT_SLEEPING = 2, ACT_WANTS_DEACT = 3, ACT_DISABLE_DEACT = 4, ACT_DISABLE_SIM = 5 };

struct WheelW {  // btWheelInfoRL working values (btVehicleRL.h:9-27, btWheelInfo.h)
	V3 hardPointWS, wheelDirWS, wheelAxleWS;
	V3 contactPointWS, contactNormalWS;
	V3 impulse;
	M3 wtBasis;  // wheel world transform basis
	float suspLen, suspRelVel, clippedInv, suspForce;
	int isInContact, isInContactWithWorld, groundObj;
};

constexpr int MAX_CAND = 14;
struct BodyW {  // btRigidBody working values of one dynamic body
	M3 invInertiaW;
	V3 totalForce, totalTorque;
	V3 invInertiaLocal;
	float invMass;
	Xform interp;       // predicted transform (m_interpolationWorldTransform)
	V3 aabbMin, aabbMax;
	int newCell, newI, newJ, newK;  // broadphase cell of this tick's AABB min corner (applied in PH_PAIRS)
	int activation;     // Activation
	int noResponse;     // CF_NO_CONTACT_RESPONSE
	int islandTag;
	float friction, restitution;
	// cars: world transform of the hitbox (compound child) and its AABB at the CURRENT pose, shared by every narrowphase pair of
	// the body this tick (computed once in PH_PREDICT; the pose does not change before the solver)
	Xform childXf;
	V3 childMin, childMax;
	// broadphase: static proxies of the body's (new) cell whose AABB overlaps the body's and that pass the collision filter,
	// as positions in Scene::cellStatics, found by the body's own thread (PH_PREDICT); nCand < 0 = list overflow, PH_PAIRS rescans
	int nCand;
	uint16_t cand[MAX_CAND];
	uint8_t candObj[MAX_CAND + 2];  // static object index of each candidate
};

struct ContactPt {  // btManifoldPoint subset
	V3 localA, localB;
	V3 posA, posB;      // world
	V3 normalB;         // m_normalWorldOnB
	float dist;
	float friction, restitution;
	int isSpecial;
	int partId0, index0, partId1, index1;
};

struct Manifold {  // btPersistentManifold subset; slot k belongs to broadphase pair k, body0 < 0 = not created
	int body0, body1;   // object ids (dynamic index or STATIC_BASE+k)
	int nPts;
	float breakingThresh, processingThresh;
	int _pad;
	ContactPt* pts;     // 4 points of this slot in the cold scratch (ArenaW::points)
};

// broadphase pair list in creation order (btHashedOverlappingPairCache array) + solver manifold order
struct PairList {
	int nPairs, nManifolds;
	uint32_t mfMask;     // bit k: pair k has a manifold (new_manifold was called for it this tick)
	uint32_t deferMask;  // bit k: the contacts of pair k go through candidate records (committed in PH_NARROW_COMMIT / PH_DYN_SOLVE)
	uint32_t dynMask;    // bit k: pair k joins two dynamic bodies (the only pairs that merge islands)
	uint8_t a[MAX_MANIFOLDS], b[MAX_MANIFOLDS];
	uint8_t manifoldOrder[MAX_MANIFOLDS];
	uint16_t entry[MAX_MANIFOLDS];  // static pairs: position of the static object in the cell list (Scene::cellStatics)
};

// Contact candidate of the staged narrowphase (rsb_collide.h).  Bodies whose shape needs an iterative detector (car box, puck
// hull) do not run it inside the per-body pair walk: the walk only RECORDS what has to be tested -- one record per
// (pair, triangle), per plane contact, per dynamic pair -- in pair order / BVH leaf order; PH_NARROW_DETECT then runs one
// detector per pending record on its own thread (so the triangles of one car-mesh pair, and GJK / EPA calls of different
// arenas, proceed in parallel instead of one after the other on the body's thread), and the records are committed to the
// manifolds in index order, which is the order the reference adds them in.
constexpr int MAX_CANDS = 120;
enum CandState : int { CAND_PENDING = 0, CAND_CONTACT = 1, CAND_NONE = 2 };
struct CandRec {
	int16_t pair;   // manifold slot = broadphase pair index
	int16_t mesh;   // arena mesh of the triangle, -1 none
	int32_t tri;    // triangle index, -1 none
	int32_t state;  // CandState
	float depth;
	V3 normal, point;  // normalOnB / pointInWorld as the detector reports them
};

struct SolverBody {
	V3 dLin, dAng, push, turn;
	V3 lin, ang, extF, extT;
	V3 invMassV;
	Xform xf;
	int orig;  // dynamic body index or -1 (fixed)
};

struct SolverRow {
	V3 n1, n2, rp1xn, rp2xn, angA, angB;
	float jacDiagInv, rhs, rhsPen, cfm, lower, upper, friction, applied, appliedPush;
	int bodyA, bodyB, frictionIndex, isSpecial;
};

// ---- EPA scratch (see rsb_epa.h) ----
constexpr int EPA_MAX_VERTICES = 128, EPA_MAX_FACES = 256, EPA_MAX_ITERATIONS = 255;
constexpr float GJK2_ACCURACY = 0.0001f, GJK2_MIN_DISTANCE = 0.0001f, GJK2_DUPLICATED_EPS = 0.0001f;
constexpr float EPA_ACCURACY = 0.0001f, EPA_PLANE_EPS = 0.00001f;

struct SV { V3 d, w; };
struct Simplex2 { SV c[4]; float p[4]; int rank; };

struct EpaFace {
	V3 n; float d;
	int16_t c[3];      // vertex indices into EpaScratch::sv
	int16_t f[3];      // adjacent faces
	int16_t l[2];      // list links (prev, next), -1 = none
	uint8_t e[3];
	uint8_t pass;
};
struct EpaScratch {
	SV sv[EPA_MAX_VERTICES + 4];
	EpaFace fc[EPA_MAX_FACES];
	int16_t stackF[EPA_MAX_FACES + 8]; uint8_t stackE[EPA_MAX_FACES + 8]; uint8_t stackS[EPA_MAX_FACES + 8];
#if defined(RSB_PROFILE)
	int profCalls;  // developer builds: narrowphase detector calls of the executing thread in the current phase (+1 GJK/box-box, +256 EPA)
#endif
};

// Solver scratch of the common case (few rows): counters, solver bodies and the first SOLVER_ROWS_FAST rows of
// each pool.  In the kernel it ALIASES the wheel scratch, which is dead once the car pre-tick is over.
constexpr int SOLVER_ROWS_FAST = 6;
constexpr int SOLV_BYTES_MAX = 16 + (MAX_BODIES + 1) * (int)sizeof(SolverBody) + 2 * SOLVER_ROWS_FAST * (int)sizeof(SolverRow);
RSB_HD constexpr int solv_bytes(int nb) { return 16 + (nb + 1) * (int)sizeof(SolverBody) + 2 * SOLVER_ROWS_FAST * (int)sizeof(SolverRow); }

// Hot per-tick scratch: lives in shared memory next to the arena record while a launch runs (the kernel
// lays out only the used prefix of each array); ArenaH is its full-size global mirror (debug probes, host check).
struct ArenaH {
	BodyW body[MAX_BODIES];
	WheelW wheels[MAX_WHEELS];
	int padLocked[MAX_PADS];  // BoostPadState::curLockedCar as car index, -1 none
	int misc[4];              // [0] = this tick needs the general solver, [1] = #manifolds handed to it, [2] = bodies written back by the solver, [3] = #contact candidates
	PairList pairs;
	Manifold mf[MAX_MANIFOLDS];
	alignas(8) unsigned char solv[SOLV_BYTES_MAX];
};

// Cold per-tick scratch (global memory): only touched when there are contacts.
struct ArenaW {
	CandRec cand[MAX_CANDS];
	ContactPt points[MAX_POINTS];
	SolverRow rows[MAX_ROWS];    // overflow beyond SOLVER_ROWS_FAST (indexed like the full pool)
	SolverRow frows[MAX_ROWS];
};

// developer builds (-DRSB_PROFILE): cycle split inside PH_DYN_SOLVE, summed over arenas and ticks
#if defined(RSB_PROFILE) && defined(__CUDACC__)
static __device__ unsigned long long g_subprof[64];
#endif
#if defined(RSB_PROFILE) && defined(__CUDA_ARCH__)
#define RSB_SUBPROF_BEGIN() long long _sp0 = clock64()
#define RSB_SUBPROF_BEGIN2() _sp0 = clock64()
#define RSB_SUBPROF_MARK(k) do { long long _n = clock64(); atomicAdd(&g_subprof[k], (unsigned long long)(_n - _sp0)); atomicAdd(&g_subprof[32 + (k)], 1ull); _sp0 = _n; } while (0)
#else
#define RSB_SUBPROF_BEGIN() ((void)0)
#define RSB_SUBPROF_BEGIN2() ((void)0)
#define RSB_SUBPROF_MARK(k) ((void)0)
#endif

struct EventSink {
	struct Ev { uint32_t arena, type, a, b; uint64_t tick; };
	Ev* buf;
	uint32_t* count;
	uint32_t cap;
};

struct Ctx {
	ArenaS* s;         // the arena record the tick works on (shared memory inside the step kernel)
	ArenaS* g;         // the same arena's record in HBM (cold tail only; == s in the host check build)
	BodyW* body;       // hot scratch (see ArenaH)
	WheelW* wheels;
	int* padLocked;
	int* misc;
	PairList* pl;
	Manifold* mf;
	unsigned char* solv;  // solver scratch (see SolverView)
	ArenaW* w;
	EpaScratch* epa;   // EPA working set of the EXECUTING THREAD (several bodies of one arena may run EPA concurrently)
	const BatchParams* p;
	const Scene* sc;
	EventSink ev;
	int arena;  // global arena index inside the batch
	// Batch constants that select code paths (game mode, ball shape, Dropshot tiles).  The step kernel is instantiated once per
	// game mode and writes compile-time constants here (make_ctx): after inlining, every `cx.mode == ...` / `cx.ballIsSphere`
	// test folds away and the other modes' code (Heatseeker steering, puck hull paths, tile pairs) leaves the instantiation.
	// Everywhere else (host check build, helper kernels) they are copied from BatchParams / Scene (ctx_set_mode).
	int mode, ballIsSphere, nTiles;
	// detector work queue of the executing block (PH_NARROW_DETECT): entries = arLocal * 128 + candidate index
	uint16_t* q; unsigned* nq; int qcap; int arLocal;
};

RSB_HD void ctx_set_mode(Ctx& cx, const BatchParams& P, const Scene* sc) { cx.mode = P.mode; cx.ballIsSphere = P.ballIsSphere; cx.nTiles = sc ? sc->nTiles : 0; }

struct SolverView {
	int* cnt;  // [0] solver bodies, [1] contact rows, [2] friction rows
	SolverBody* sbody;
	SolverRow *rows, *frows, *orows, *ofrows;
	RSB_HD SolverRow& row(int j) const { return j < SOLVER_ROWS_FAST ? rows[j] : orows[j]; }
	RSB_HD SolverRow& frow(int j) const { return j < SOLVER_ROWS_FAST ? frows[j] : ofrows[j]; }
};
RSB_HD SolverView solver_view(const Ctx& cx) {
	SolverView v;
	v.cnt = (int*)cx.solv;
	v.sbody = (SolverBody*)(cx.solv + 16);
	v.rows = (SolverRow*)(v.sbody + (cx.p->nCars + 2));
	v.frows = v.rows + SOLVER_ROWS_FAST;
	v.orows = cx.w->rows; v.ofrows = cx.w->frows;
	return v;
}

RSB_HD void set_error(const Ctx& cx, uint32_t flag) {
#if defined(__CUDA_ARCH__)
	atomicOr(&cx.s->errorFlags, flag);
#else
	cx.s->errorFlags |= flag;
#endif
}

RSB_HD RigidS& rigid_of(const Ctx& cx, int b) { return b == 0 ? cx.s->ball.rb : cx.s->cars[b - 1].rb; }
RSB_HD int n_bodies(const Ctx& cx) { return cx.p->nCars + 1; }

RSB_HD void push_event(const Ctx& cx, uint32_t type, uint32_t a, uint32_t b) {
	if (!cx.ev.buf) return;
#if defined(__CUDA_ARCH__)
	uint32_t i = atomicAdd(cx.ev.count, 1u);
#else
	uint32_t i = (*cx.ev.count)++;
#endif
	if (i < cx.ev.cap) {
		EventSink::Ev e; e.arena = (uint32_t)(cx.p->arenaOffset + cx.arena); e.type = type; e.a = a; e.b = b; e.tick = cx.s->tickCount;
		cx.ev.buf[i] = e;
	} else {
		set_error(cx, ERR_EVENT_OVERFLOW);
	}
}

// same with an explicit tick stamp (the tracker reports Arena::tickCount as its Update() sees it, i.e. after the step)
RSB_HD void push_event_at(const Ctx& cx, uint32_t type, uint32_t a, uint32_t b, uint64_t tick) {
	if (!cx.ev.buf) return;
#if defined(__CUDA_ARCH__)
	uint32_t i = atomicAdd(cx.ev.count, 1u);
#else
	uint32_t i = (*cx.ev.count)++;
#endif
	if (i < cx.ev.cap) { EventSink::Ev e; e.arena = (uint32_t)(cx.p->arenaOffset + cx.arena); e.type = type; e.a = a; e.b = b; e.tick = tick; cx.ev.buf[i] = e; }
	else set_error(cx, ERR_EVENT_OVERFLOW);
}

RSB_HD bool body_active(const BodyW& b) { return b.activation != ACT_SLEEPING && b.activation != ACT_DISABLE_SIM; }

// ---- btRigidBody helpers on (RigidS, BodyW) pairs ----
RSB_HD V3 vel_at(const RigidS& rb, const V3& rel) { return rb.vel + cross(rb.angvel, rel); }  // getVelocityInLocalPoint
RSB_HD void apply_central_impulse(RigidS& rb, const BodyW& bw, const V3& imp) { rb.vel += imp * bw.invMass; }
RSB_HD void apply_torque_impulse(RigidS& rb, const BodyW& bw, const V3& t) { rb.angvel += bw.invInertiaW * t; }
RSB_HD void apply_impulse(RigidS& rb, const BodyW& bw, const V3& imp, const V3& rel) {
	if (bw.invMass != 0.f) {
		apply_central_impulse(rb, bw, imp);
		apply_torque_impulse(rb, bw, cross(rel, imp));
	}
}
RSB_HD void apply_central_force(BodyW& bw, const V3& f) { bw.totalForce += f; }
RSB_HD void apply_torque(BodyW& bw, const V3& t) { bw.totalTorque += t; }
// btRigidBody::computeImpulseDenominator (btRigidBody.h:495-504)
RSB_HD float impulse_denominator(const RigidS& rb, const BodyW& bw, const V3& pos, const V3& normal) {
	V3 r0 = pos - rb.pos;
	V3 c0 = cross(r0, normal);
	V3 vec = cross(vmul(c0, bw.invInertiaW), r0);
	return bw.invMass + dot(normal, vec);
}

}  // namespace rsb
