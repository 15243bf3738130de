// oracle/ref_harness.cpp -- ORACLE shim (test infrastructure only; see ref_harness.h).
//
// Our own code.  Wraps the public C++ surface of the unmodified reference
// (/root/reference/src/RocketSim.h) in a C ABI so that Python tests and bench.py's
// CPU-baseline leg can drive the real reference implementation of Arena::Step.
#include "ref_harness.h"
#include "../include/rsb_bench_inputs.h"

#include "RocketSim.h"
#include "Sim/GameEventTracker/GameEventTracker.h"
#include "Sim/BallPredTracker/BallPredTracker.h"
#include "../libsrc/bullet3-3.24/BulletCollision/CollisionShapes/btBvhTriangleMeshShape.h"
#include "../libsrc/bullet3-3.24/BulletCollision/CollisionShapes/btTriangleInfoMap.h"
#include "../libsrc/bullet3-3.24/BulletCollision/CollisionShapes/btOptimizedBvh.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

using namespace RocketSim;

static_assert(sizeof(RsbVec) == sizeof(Vec), "Vec");
static_assert(sizeof(RsbRotMat) == sizeof(RotMat), "RotMat");
static_assert(sizeof(RsbCarControls) == sizeof(CarControls), "CarControls");
static_assert(sizeof(RsbBallHitInfo) == sizeof(BallHitInfo), "BallHitInfo");
static_assert(sizeof(RsbCarState) == sizeof(CarState), "CarState");
static_assert(sizeof(RsbBallState) == sizeof(BallState), "BallState");
static_assert(sizeof(RsbMutatorConfig) == sizeof(MutatorConfig), "MutatorConfig");
static_assert(sizeof(RsbCarConfig) == sizeof(CarConfig), "CarConfig");
static_assert(offsetof(RsbCarState, boost) == offsetof(CarState, boost), "boost");
static_assert(offsetof(RsbCarState, worldContact) == offsetof(CarState, worldContact), "worldContact");
static_assert(offsetof(RsbCarState, carContact) == offsetof(CarState, carContact), "carContact");
static_assert(offsetof(RsbCarState, isDemoed) == offsetof(CarState, isDemoed), "isDemoed");
static_assert(offsetof(RsbCarState, ballHitInfo) == offsetof(CarState, ballHitInfo), "ballHitInfo");
static_assert(offsetof(RsbCarState, lastControls) == offsetof(CarState, lastControls), "lastControls");
static_assert(offsetof(RsbBallState, hsInfo) == offsetof(BallState, hsInfo), "hsInfo");
static_assert(offsetof(RsbBallState, dsInfo) == offsetof(BallState, dsInfo), "dsInfo");
static_assert(offsetof(RsbMutatorConfig, ballRadius) == offsetof(MutatorConfig, ballRadius), "ballRadius");
static_assert(offsetof(RsbMutatorConfig, goalBaseThresholdY) == offsetof(MutatorConfig, goalBaseThresholdY), "goalY");
static_assert(offsetof(RsbCarConfig, dodgeDeadzone) == offsetof(CarConfig, dodgeDeadzone), "dodgeDeadzone");

struct RefArena {
	Arena* arena = nullptr;
	std::vector<Car*> cars; // in AddCar order (index = id-1)
	std::vector<RsbEvent> events;
	GameEventTracker* tracker = nullptr;  // optional: the reference's own shot / goal / save detection, updated after every tick
};

static thread_local std::string g_err;

static const CarConfig* preset(int i) {
	static const CarConfig* P[7] = { &CAR_CONFIG_OCTANE, &CAR_CONFIG_DOMINUS, &CAR_CONFIG_PLANK, &CAR_CONFIG_BREAKOUT,
	                                 &CAR_CONFIG_HYBRID, &CAR_CONFIG_MERC, &CAR_CONFIG_PSYCLOPS };
	return P[(i < 0 || i > 6) ? 0 : i];
}

extern "C" {

const char* ref_last_error(void) { return g_err.c_str(); }

int ref_init_from_mem(int n_meshes, const int* mesh_modes, const uint8_t* const* mesh_ptrs, const size_t* mesh_sizes) {
	try {
		if (GetStage() == RocketSimStage::INITIALIZED) return 0;
		std::map<GameMode, std::vector<FileData>> m;
		for (int i = 0; i < n_meshes; i++)
			m[(GameMode)mesh_modes[i]].push_back(FileData(mesh_ptrs[i], mesh_ptrs[i] + mesh_sizes[i]));
		InitFromMem(m, true);
		return 0;
	} catch (std::exception& e) { g_err = e.what(); return -1; }
}

RefArena* ref_arena_create(int game_mode, float tick_rate) { return ref_arena_create_cfg(game_mode, tick_rate, 1, -1, nullptr); }

// ArenaConfig with noBallRot and (n_pads >= 0) custom boost pads: pads = n_pads x (x, y, z, isBig)
RefArena* ref_arena_create_cfg(int game_mode, float tick_rate, int no_ball_rot, int n_pads, const float* pads) {
	try {
		RefArena* r = new RefArena();
		ArenaConfig cfg;
		cfg.noBallRot = no_ball_rot != 0;
		if (n_pads >= 0) {
			cfg.useCustomBoostPads = true;
			for (int i = 0; i < n_pads; i++) { BoostPadConfig pc; pc.pos = Vec(pads[4 * i], pads[4 * i + 1], pads[4 * i + 2]); pc.isBig = pads[4 * i + 3] != 0.f; cfg.customBoostPads.push_back(pc); }
		}
		r->arena = Arena::Create((GameMode)game_mode, cfg, tick_rate);
		if ((GameMode)game_mode != GameMode::THE_VOID) {
			r->arena->SetGoalScoreCallback([](Arena* a, Team t, void* u) {
				RefArena* ra = (RefArena*)u;
				ra->events.push_back(RsbEvent{ 0, RSB_EVENT_GOAL, (uint32_t)t, 0, a->tickCount });
			}, r);
		}
		r->arena->SetCarBumpCallback([](Arena* a, Car* bumper, Car* victim, bool isDemo, void* u) {
			RefArena* ra = (RefArena*)u;
			ra->events.push_back(RsbEvent{ 0, (uint32_t)(isDemo ? RSB_EVENT_DEMO : RSB_EVENT_BUMP), bumper->id, victim->id, a->tickCount });
		}, r);
		return r;
	} catch (std::exception& e) { g_err = e.what(); return nullptr; }
}

void ref_arena_destroy(RefArena* r) { if (r) { delete r->tracker; delete r->arena; delete r; } }

uint32_t ref_add_car(RefArena* r, int team, int car_preset) {
	Car* c = r->arena->AddCar((Team)team, *preset(car_preset));
	r->cars.push_back(c);
	return c->id;
}
int ref_num_cars(RefArena* r) { return (int)r->cars.size(); }
int ref_num_pads(RefArena* r) { return (int)r->arena->_boostPads.size(); }
void ref_seed_thread_rng(uint32_t seed) { Math::GetRandEngine().seed(seed); }
void ref_reset_kickoff(RefArena* r, int seed) { r->arena->ResetToRandomKickoff(seed); }
void ref_set_controls(RefArena* r, uint32_t id, const RsbCarControls* c) { memcpy(&r->cars[id - 1]->controls, c, sizeof(CarControls)); }
void ref_step(RefArena* r, int n) {
	if (!r->tracker) { r->arena->Step(n); return; }
	for (int i = 0; i < n; i++) { r->arena->Step(1); r->tracker->Update(r->arena); }
}

/* GameEventTracker (src/Sim/GameEventTracker/GameEventTracker.h) with the given config (7 floats, field order of
 * GameEventTrackerConfig); its callbacks land in the event queue as RSB_EVENT_SHOT / _GOAL_CREDIT / _SAVE. */
void ref_enable_event_tracker(RefArena* r, const float* cfg7) {
	delete r->tracker;
	r->tracker = new GameEventTracker();
	if (cfg7) {
		GameEventTrackerConfig& c = r->tracker->config;
		c.shotMinSpeed = cfg7[0]; c.shotTouchMinDelay = cfg7[1]; c.predScoreExtraMargin = cfg7[2]; c.shotEventCooldown = cfg7[3];
		c.shotMinScoreTime = cfg7[4]; c.goalMaxTouchTime = cfg7[5]; c.passMaxTouchTime = cfg7[6];
	}
	r->tracker->SetShotCallback([](Arena* a, Car* shooter, Car* passer, void* u) {
		((RefArena*)u)->events.push_back(RsbEvent{ 0, RSB_EVENT_SHOT, shooter->id, passer ? passer->id : 0u, a->tickCount });
	}, r);
	r->tracker->SetGoalCallback([](Arena* a, Car* scorer, Car* passer, void* u) {
		((RefArena*)u)->events.push_back(RsbEvent{ 0, RSB_EVENT_GOAL_CREDIT, scorer->id, passer ? passer->id : 0u, a->tickCount });
	}, r);
	r->tracker->SetSaveCallback([](Arena* a, Car* saver, void* u) {
		((RefArena*)u)->events.push_back(RsbEvent{ 0, RSB_EVENT_SAVE, saver->id, 0u, a->tickCount });
	}, r);
}
/* Arena::Serialize / Arena::DeserializeNew (Arena.cpp:546-644) through in-memory streams. */
size_t ref_serialize(RefArena* r, uint8_t* out, size_t cap) {
	DataStreamOut s; r->arena->Serialize(s);
	if (out && cap >= s.data.size()) memcpy(out, s.data.data(), s.data.size());
	return s.data.size();
}
RefArena* ref_deserialize_new(const uint8_t* data, size_t n) {
	try {
		DataStreamIn in; in.data.assign(data, data + n);
		RefArena* r = new RefArena();
		r->arena = Arena::DeserializeNew(in);
		std::vector<Car*> cars(r->arena->_cars.begin(), r->arena->_cars.end());
		std::sort(cars.begin(), cars.end(), [](Car* a, Car* b) { return a->id < b->id; });
		r->cars = cars;
		return r;
	} catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
/* Arena::Clone (Arena.cpp:646-674). */
RefArena* ref_clone(RefArena* r) {
	RefArena* c = new RefArena();
	c->arena = r->arena->Clone(false);
	std::vector<Car*> cars(c->arena->_cars.begin(), c->arena->_cars.end());
	std::sort(cars.begin(), cars.end(), [](Car* a, Car* b) { return a->id < b->id; });
	c->cars = cars;
	return c;
}
/* BallPredTracker (src/Sim/BallPredTracker/BallPredTracker.cpp): predData of a fresh tracker built on the arena. */
int ref_ball_pred(RefArena* r, int n_ticks, RsbBallState* out) {
	BallPredTracker t(r->arena, (size_t)n_ticks);
	for (size_t i = 0; i < t.predData.size(); i++) memcpy(&out[i], &t.predData[i], sizeof(RsbBallState));
	return (int)t.predData.size();
}
int ref_is_ball_probably_going_in(RefArena* r, float maxTime, float extraMargin, int* teamOut) {
	Team t = Team::BLUE;
	bool in = r->arena->IsBallProbablyGoingIn(maxTime, extraMargin, &t);
	if (teamOut) *teamOut = (int)t;
	return in ? 1 : 0;
}
uint64_t ref_tick_count(RefArena* r) { return r->arena->tickCount; }
int ref_is_ball_scored(RefArena* r) { return r->arena->IsBallScored() ? 1 : 0; }

void ref_get_car_state(RefArena* r, uint32_t id, RsbCarState* out) {
	CarState s = r->cars[id - 1]->GetState();
	memset(out, 0, sizeof(*out));
	memcpy(out, &s, sizeof(s));
}
void ref_set_car_state(RefArena* r, uint32_t id, const RsbCarState* in) {
	CarState s; memcpy(&s, in, sizeof(s));
	r->cars[id - 1]->SetState(s);
}
void ref_get_car_internal(RefArena* r, uint32_t id, RsbCarInternal* o) {
	Car* c = r->cars[id - 1];
	for (int i = 0; i < 4; i++) {
		auto& w = c->_bulletVehicle.m_wheelInfo[i];
		o->steerAngle[i] = w.m_steerAngle; o->engineForce[i] = w.m_engineForce; o->brake[i] = w.m_brake;
		o->latFriction[i] = w.m_latFriction; o->longFriction[i] = w.m_longFriction; o->extraPushback[i] = w.m_extraPushback;
	}
	memcpy(o->velocityImpulseCache, &c->_velocityImpulseCache, 16);
	const btVector3& org = c->_rigidBody.getWorldTransform().getOrigin();
	for (int k = 0; k < 3; k++) { o->rbPos[k] = org[k]; o->rbVel[k] = c->_rigidBody.m_linearVelocity[k]; }
	o->rbPos[3] = o->rbVel[3] = 0.f;
	o->tickCountSinceUpdate = c->_internalState.tickCountSinceUpdate;
	o->hasRigidBody = 3; o->_pad = 0;
	const btMatrix3x3& bs = c->_rigidBody.getWorldTransform().getBasis();
	for (int r3 = 0; r3 < 3; r3++) { for (int k = 0; k < 3; k++) o->rbBasis[4 * r3 + k] = bs[r3][k]; o->rbBasis[4 * r3 + 3] = 0.f; }
}
void ref_set_car_internal(RefArena* r, uint32_t id, const RsbCarInternal* o) {
	Car* c = r->cars[id - 1];
	for (int i = 0; i < 4; i++) {
		auto& w = c->_bulletVehicle.m_wheelInfo[i];
		w.m_steerAngle = o->steerAngle[i]; w.m_engineForce = o->engineForce[i]; w.m_brake = o->brake[i];
		w.m_latFriction = o->latFriction[i]; w.m_longFriction = o->longFriction[i]; w.m_extraPushback = o->extraPushback[i];
	}
	memcpy(&c->_velocityImpulseCache, o->velocityImpulseCache, 16);
	if (o->hasRigidBody & 1) {
		c->_rigidBody.getWorldTransform().setOrigin(btVector3(o->rbPos[0], o->rbPos[1], o->rbPos[2]));
		c->_rigidBody.m_linearVelocity = btVector3(o->rbVel[0], o->rbVel[1], o->rbVel[2]);
		c->_internalState.tickCountSinceUpdate = o->tickCountSinceUpdate;
	}
	if (o->hasRigidBody & 2) {
		btMatrix3x3 bs(o->rbBasis[0], o->rbBasis[1], o->rbBasis[2], o->rbBasis[4], o->rbBasis[5], o->rbBasis[6], o->rbBasis[8], o->rbBasis[9], o->rbBasis[10]);
		c->_rigidBody.getWorldTransform().setBasis(bs);
		c->_rigidBody.updateInertiaTensor();
	}
}
void ref_get_ball_state(RefArena* r, RsbBallState* out) {
	BallState s = r->arena->ball->GetState();
	memset(out, 0, sizeof(*out));
	memcpy(out, &s, sizeof(s));
	/* BallState::DropshotInfo::lastDamageTick has no initialiser and is "only valid if hasDamaged" (Ball.h:44-45): report 0
	 * instead of the indeterminate value until the first damage */
	if (!out->dsInfo.hasDamaged) out->dsInfo.lastDamageTick = 0;
}
void ref_set_ball_state(RefArena* r, const RsbBallState* in) {
	BallState s; memcpy(&s, in, sizeof(s));
	r->arena->ball->SetState(s);
}
void ref_get_ball_internal(RefArena* r, RsbBallInternal* o) {
	memset(o, 0, sizeof(*o));
	memcpy(o->velocityImpulseCache, &r->arena->ball->_velocityImpulseCache, 16);
	o->groundStickApplied = r->arena->ball->_groundStickApplied;
	const btVector3& org = r->arena->ball->_rigidBody.getWorldTransform().getOrigin();
	for (int k = 0; k < 3; k++) { o->rbPos[k] = org[k]; o->rbVel[k] = r->arena->ball->_rigidBody.m_linearVelocity[k]; }
	o->tickCountSinceUpdate = r->arena->ball->_internalState.tickCountSinceUpdate;
	o->hasRigidBody = 1;
}
void ref_set_ball_internal(RefArena* r, const RsbBallInternal* o) {
	memcpy(&r->arena->ball->_velocityImpulseCache, o->velocityImpulseCache, 16);
	r->arena->ball->_groundStickApplied = o->groundStickApplied;
	if (o->hasRigidBody) {
		r->arena->ball->_rigidBody.getWorldTransform().setOrigin(btVector3(o->rbPos[0], o->rbPos[1], o->rbPos[2]));
		r->arena->ball->_rigidBody.m_linearVelocity = btVector3(o->rbVel[0], o->rbVel[1], o->rbVel[2]);
		r->arena->ball->_internalState.tickCountSinceUpdate = o->tickCountSinceUpdate;
	}
}
void ref_get_pad_states(RefArena* r, RsbBoostPadState* out) {
	auto& pads = r->arena->_boostPads;
	for (size_t i = 0; i < pads.size(); i++) {
		BoostPadState s = pads[i]->GetState();
		memset(&out[i], 0, sizeof(out[i]));
		out[i].isActive = s.isActive; out[i].cooldown = s.cooldown; out[i].prevLockedCarID = s.prevLockedCarID;
	}
}
void ref_set_pad_states(RefArena* r, const RsbBoostPadState* in) {
	auto& pads = r->arena->_boostPads;
	for (size_t i = 0; i < pads.size(); i++) {
		BoostPadState s;
		s.isActive = in[i].isActive; s.cooldown = in[i].cooldown; s.prevLockedCarID = in[i].prevLockedCarID; s.curLockedCar = NULL;
		pads[i]->SetState(s);
	}
}
/* Arena::GetDropshotTilesState / SetDropshotTilesState (Arena.h:178-179): 140 damage-state bytes, [team][70] */
void ref_get_dropshot_tiles(RefArena* r, uint8_t* out) {
	DropshotTilesState s = r->arena->GetDropshotTilesState();
	for (int t = 0; t < 2; t++) for (int i = 0; i < 70; i++) out[t * 70 + i] = s.states[t][i].damageState;
}
void ref_set_dropshot_tiles(RefArena* r, const uint8_t* in) {
	DropshotTilesState s;
	for (int t = 0; t < 2; t++) for (int i = 0; i < 70; i++) s.states[t][i].damageState = in[t * 70 + i];
	r->arena->SetDropshotTilesState(s);
}
/* DropshotTiles::GetTilePos (DropshotTiles.cpp:13-20), uu */
void ref_get_dropshot_tile_pos(int team, int idx, float* xyz) {
	Vec p = DropshotTiles::GetTilePos(team, idx);
	xyz[0] = p.x; xyz[1] = p.y; xyz[2] = p.z;
}
void ref_get_pad_positions(RefArena* r, float* o) {
	auto& pads = r->arena->_boostPads;
	for (size_t i = 0; i < pads.size(); i++) {
		o[i * 4 + 0] = pads[i]->config.pos.x; o[i * 4 + 1] = pads[i]->config.pos.y; o[i * 4 + 2] = pads[i]->config.pos.z;
		o[i * 4 + 3] = pads[i]->config.isBig ? 1.f : 0.f;
	}
}
void ref_get_mutator_config(RefArena* r, RsbMutatorConfig* out) { memset(out, 0, sizeof(*out)); MutatorConfig m = r->arena->GetMutatorConfig(); memcpy(out, &m, sizeof(m)); }
void ref_set_mutator_config(RefArena* r, const RsbMutatorConfig* in) { MutatorConfig m(r->arena->gameMode); memcpy(&m, in, sizeof(m)); r->arena->SetMutatorConfig(m); }
void ref_default_mutator_config(int mode, RsbMutatorConfig* out) { memset(out, 0, sizeof(*out)); MutatorConfig m((GameMode)mode); memcpy(out, &m, sizeof(m)); }
void ref_get_car_config(int p, RsbCarConfig* out) { memset(out, 0, sizeof(*out)); memcpy(out, preset(p), sizeof(CarConfig)); }

int ref_drain_events(RefArena* r, RsbEvent* out, int cap) {
	int n = (int)r->events.size(); if (n > cap) n = cap;
	for (int i = 0; i < n; i++) out[i] = r->events[i];
	r->events.clear();
	return n;
}

int ref_num_meshes(int mode) { return (int)GetArenaCollisionShapes((GameMode)mode).size(); }

int ref_bvh_counts(int mode, int mesh, int* c) {
	auto& shapes = GetArenaCollisionShapes((GameMode)mode);
	if (mesh < 0 || mesh >= (int)shapes.size()) return -1;
	btBvhTriangleMeshShape* s = shapes[mesh];
	btOptimizedBvh* bvh = s->getOptimizedBvh();
	c[0] = bvh->getQuantizedNodeArray().size();
	c[1] = bvh->getSubtreeInfoArray().size();
	const unsigned char *vb, *ib; int nv, vs, is, nf;
	s->getMeshInterface()->getLockedReadOnlyVertexIndexBase(&vb, nv, vs, &ib, is, nf);
	c[2] = nf; c[3] = nv;
	c[4] = s->getTriangleInfoMap() ? s->getTriangleInfoMap()->size() : 0;
	return 0;
}

int ref_bvh_dump(int mode, int mesh, void* nodes, void* headers, float* quant9, void* triinfo) {
	auto& shapes = GetArenaCollisionShapes((GameMode)mode);
	if (mesh < 0 || mesh >= (int)shapes.size()) return -1;
	btBvhTriangleMeshShape* s = shapes[mesh];
	btOptimizedBvh* bvh = s->getOptimizedBvh();
	auto& na = bvh->getQuantizedNodeArray();
	for (int i = 0; i < na.size(); i++) memcpy((char*)nodes + 16 * i, &na[i], 16);
	auto& ha = bvh->getSubtreeInfoArray();
	for (int i = 0; i < ha.size(); i++) {
		char* p = (char*)headers + 20 * i;
		memcpy(p, ha[i].m_quantizedAabbMin, 6); memcpy(p + 6, ha[i].m_quantizedAabbMax, 6);
		memcpy(p + 12, &ha[i].m_rootNodeIndex, 4); memcpy(p + 16, &ha[i].m_subtreeSize, 4);
	}
	// m_bvhAabbMin/Max/Quantization are protected: recover them through the public (un)quantize API.
	unsigned short q0[3] = { 0, 0, 0 }, q1[3] = { 1, 1, 1 };
	btVector3 mn = bvh->unQuantize(q0), one = bvh->unQuantize(q1);
	for (int k = 0; k < 3; k++) { quant9[k] = mn[k]; quant9[3 + k] = 0.f; quant9[6 + k] = one[k] - mn[k]; }
	const unsigned char *vb, *ib; int nv, vs, is, nf;
	s->getMeshInterface()->getLockedReadOnlyVertexIndexBase(&vb, nv, vs, &ib, is, nf);
	btTriangleInfoMap* tim = s->getTriangleInfoMap();
	for (int t = 0; t < nf; t++) {
		int32_t flags = -1; float a[3] = { SIMD_2_PI, SIMD_2_PI, SIMD_2_PI };
		if (tim) {
			btTriangleInfo* info = tim->find(btHashInt(t));
			if (info) { flags = info->m_flags; a[0] = info->m_edgeV0V1Angle; a[1] = info->m_edgeV1V2Angle; a[2] = info->m_edgeV2V0Angle; }
		}
		char* p = (char*)triinfo + 16 * t;
		memcpy(p, &flags, 4); memcpy(p + 4, a, 12);
	}
	return 0;
}

void ref_get_wheel_debug(RefArena* r, uint32_t id, float* o) {
	Car* c = r->cars[id - 1];
	for (int i = 0; i < 4; i++) {
		auto& w = c->_bulletVehicle.m_wheelInfo[i];
		float* p = o + 14 * i;
		p[0] = w.m_raycastInfo.m_isInContact; p[1] = w.m_isInContactWithWorld; p[2] = w.m_raycastInfo.m_suspensionLength;
		for (int k = 0; k < 3; k++) { p[3 + k] = w.m_raycastInfo.m_contactPointWS[k]; p[6 + k] = w.m_raycastInfo.m_contactNormalWS[k]; p[11 + k] = w.m_impulse[k]; }
		p[9] = w.m_suspensionRelativeVelocity; p[10] = w.m_clippedInvContactDotSuspension;
	}
}
void ref_get_car_body(RefArena* r, uint32_t id, float* o) {
	Car* c = r->cars[id - 1];
	const btTransform& t = c->_rigidBody.getWorldTransform();
	for (int k = 0; k < 3; k++) { o[k] = t.getOrigin()[k]; o[12 + k] = c->_rigidBody.m_linearVelocity[k]; o[15 + k] = c->_rigidBody.m_angularVelocity[k]; }
	for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) o[3 + i * 3 + j] = t.getBasis()[i][j];
}

// Contact manifolds alive after the last step (they persist until the next tick's pair wipe):
// 16 floats per point {obj0, obj1, point#, dist, normalOnB xyz, posA xyz, posB xyz, friction, restitution, isSpecial};
// obj = index in the world's collision object array (meshes, planes, ball, cars).
int ref_get_manifolds(RefArena* r, float* out, int cap_points) {
	btCollisionDispatcher* d = r->arena->_bulletWorld.getDispatcher();
	int n = 0;
	for (int m = 0; m < d->getNumManifolds(); m++) {
		btPersistentManifold* pm = d->getManifoldByIndexInternal(m);
		for (int k = 0; k < pm->getNumContacts() && n < cap_points; k++, n++) {
			const btManifoldPoint& p = pm->getContactPoint(k);
			float* o = out + 16 * n;
			o[0] = (float)pm->getBody0()->getWorldArrayIndex(); o[1] = (float)pm->getBody1()->getWorldArrayIndex(); o[2] = (float)k;
			o[3] = p.m_distance1;
			for (int c = 0; c < 3; c++) { o[4 + c] = p.m_normalWorldOnB[c]; o[7 + c] = p.m_positionWorldOnA[c]; o[10 + c] = p.m_positionWorldOnB[c]; }
			o[13] = p.m_combinedFriction; o[14] = p.m_combinedRestitution; o[15] = (float)p.m_isSpecial;
		}
	}
	return n;
}

// Solver rows of the last step (the pools keep their memory after solveGroupCacheFriendlyFinish resets the sizes):
// which = 0 contact rows, 1 friction rows; 24 floats per row {n1, rp1xn, n2, rp2xn, angA, angB (3 each), jacDiagInv, rhs,
// rhsPenetration, appliedImpulse, appliedPushImpulse, friction}.
void ref_get_solver_rows(RefArena* r, int which, int n, float* out) {
	btSequentialImpulseConstraintSolver* s = r->arena->_bulletWorld.getConstraintSolver();
	btConstraintArray& pool = which ? s->m_tmpSolverContactFrictionConstraintPool : s->m_tmpSolverContactConstraintPool;
	if (pool.capacity() < n) n = pool.capacity();
	for (int i = 0; i < n; i++) {
		const btSolverConstraint& c = pool[i]; float* o = out + 24 * i;
		for (int k = 0; k < 3; k++) {
			o[k] = c.m_contactNormal1[k]; o[3 + k] = c.m_relpos1CrossNormal[k]; o[6 + k] = c.m_contactNormal2[k];
			o[9 + k] = c.m_relpos2CrossNormal[k]; o[12 + k] = c.m_angularComponentA[k]; o[15 + k] = c.m_angularComponentB[k];
		}
		o[18] = c.m_jacDiagABInv; o[19] = c.m_rhs; o[20] = c.m_rhsPenetration; o[21] = c.m_appliedImpulse; o[22] = c.m_appliedPushImpulse; o[23] = c.m_friction;
	}
}

// Solver bodies of the last step: 12 floats each {linVel, angVel, externalForceImpulse, externalTorqueImpulse} (as at setup
// for lin/ang: m_linearVelocity/m_angularVelocity of btSolverBody are not modified by the iterations).
void ref_get_solver_bodies(RefArena* r, int n, float* out) {
	btSequentialImpulseConstraintSolver* s = r->arena->_bulletWorld.getConstraintSolver();
	if (s->m_tmpSolverBodyPool.capacity() < n) n = s->m_tmpSolverBodyPool.capacity();
	for (int i = 0; i < n; i++) {
		const btSolverBody& b = s->m_tmpSolverBodyPool[i]; float* o = out + 12 * i;
		for (int k = 0; k < 3; k++) { o[k] = b.m_linearVelocity[k]; o[3 + k] = b.m_angularVelocity[k]; o[6 + k] = b.m_externalForceImpulse[k]; o[9 + k] = b.m_externalTorqueImpulse[k]; }
	}
}

/* Steps an EXISTING arena `ticks` ticks with the pinned README-bench input stream of global arena id `gid` (include/rsb_bench_inputs.h),
 * starting at stream tick `tick0`, entirely in C (no per-tick FFI); no kickoff reset.  Returns the stepping wall time in seconds.
 * Used for BASELINE config 1 (THE_VOID 1v1, 1 M ticks): the CPU number and the state the GPU rollout is compared with. */
double ref_run_bench_inputs(RefArena* r, int ticks, uint64_t seed, int gid, int tick0) {
	const int ncars = (int)r->cars.size();
	std::vector<RsbInputSched> ss(ncars);
	for (auto& s : ss) rsb_inp_init(&s);
	// replay the schedule up to tick0 so that a rollout can be continued in several calls
	for (int k = 0; k < tick0; k++) for (int i = 0; i < ncars; i++) rsb_inp_update(&ss[i], seed, gid, i, k);
	auto t0 = std::chrono::steady_clock::now();
	for (int k = tick0; k < tick0 + ticks; k++) {
		for (int i = 0; i < ncars; i++) {
			RsbInputSched& s = ss[i];
			rsb_inp_update(&s, seed, gid, i, k);
			CarControls& c = r->cars[i]->controls;
			c.throttle = s.throttle; c.steer = s.steer; c.pitch = s.pitch; c.yaw = s.yaw; c.roll = s.roll;
			c.jump = s.jump; c.boost = s.boost; c.handbrake = s.handbrake;
		}
		r->arena->Step(1);
	}
	return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

double ref_bench(int mode, int cars_per_team, int n_threads, int arenas_per_thread, int ticks, uint64_t seed, double* wall_s) {
	std::atomic<int> ready(0);
	std::atomic<bool> go(false);
	std::vector<std::thread> th;
	std::vector<double> secs(n_threads, 0.0);
	for (int t = 0; t < n_threads; t++) {
		th.emplace_back([&, t]() {
			const int ncars = cars_per_team * 2;
			std::vector<Arena*> arenas; std::vector<std::vector<Car*>> cars; std::vector<std::vector<RsbInputSched>> sched;
			Math::GetRandEngine().seed((uint32_t)(seed + 7919u * t));
			for (int a = 0; a < arenas_per_thread; a++) {
				Arena* ar = Arena::Create((GameMode)mode);
				std::vector<Car*> cs;
				for (int i = 0; i < ncars; i++) cs.push_back(ar->AddCar(i < cars_per_team ? Team::BLUE : Team::ORANGE));
				int gid = t * arenas_per_thread + a;
				ar->ResetToRandomKickoff((int)(seed + gid));
				arenas.push_back(ar); cars.push_back(cs);
				std::vector<RsbInputSched> ss(ncars); for (auto& s : ss) rsb_inp_init(&s);
				sched.push_back(ss);
			}
			ready++;
			while (!go.load()) std::this_thread::yield();
			auto t0 = std::chrono::steady_clock::now();
			for (int k = 0; k < ticks; k++) {
				for (int a = 0; a < arenas_per_thread; a++) {
					int gid = t * arenas_per_thread + a;
					for (int i = 0; i < ncars; i++) {
						RsbInputSched& s = sched[a][i];
						rsb_inp_update(&s, seed, gid, i, k);
						CarControls& c = cars[a][i]->controls;
						c.throttle = s.throttle; c.steer = s.steer; c.pitch = s.pitch; c.yaw = s.yaw; c.roll = s.roll;
						c.jump = s.jump; c.boost = s.boost; c.handbrake = s.handbrake;
					}
					arenas[a]->Step(1);
					if (arenas[a]->IsBallScored()) arenas[a]->ResetToRandomKickoff((int)(seed + gid + k));
				}
			}
			secs[t] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
			for (Arena* ar : arenas) delete ar;
		});
	}
	while (ready.load() < n_threads) std::this_thread::yield();
	auto t0 = std::chrono::steady_clock::now();
	go.store(true);
	for (auto& x : th) x.join();
	(void)t0;
	// stepping time only: the slowest thread's loop (arena construction/destruction is untimed)
	double wall = 0.0;
	for (double s : secs) wall = s > wall ? s : wall;
	if (wall_s) *wall_s = wall;
	return (double)n_threads * arenas_per_thread * ticks / wall;
}

} // extern "C"
