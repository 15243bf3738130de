// rsb_state.h -- device-resident data model of the batched arena tick.
//
//   ArenaS  : everything the reference keeps alive BETWEEN ticks for one arena: the rigid
//             body pose/velocity of ball and cars (Bullet units), CarState/BallState game
//             fields (reference src/Sim/Car/Car.h:17-120, src/Sim/Ball/Ball.h:18-57), the
//             wheel values that the reference consumes one tick late, pad
//             state, the respawn RNG and the broadphase cell history stamps.
//             One contiguous record per arena: a warp owns an arena and streams its record
//             with coalesced 16-byte accesses.
//   ArenaW  : per-tick scratch of one arena (wheel ray results, swept AABBs, contact
//             manifolds, solver rows).  Never read across ticks.
//   BatchParams / Scene : read-only constants shared by all arenas of a batch / game mode.
#pragma once

#include "rsb_math.h"

namespace rsb {

constexpr int MAX_CARS = 6;          // 3v3
constexpr int MAX_BODIES = MAX_CARS + 1;  // dynamic bodies: ball (index 0) + cars (1..n)
constexpr int MAX_PADS = 40;         // Soccar 34, Hoops 20
constexpr int MAX_WHEELS = MAX_CARS * 4;
constexpr int MAX_MANIFOLDS = 32;    // broadphase pairs per arena and tick (= manifold slots)
constexpr int MAX_POINTS = MAX_MANIFOLDS * 4;
constexpr int MAX_ROWS = MAX_POINTS + MAX_BODIES;
constexpr int MAX_EVENTS_PER_ARENA = 8;
constexpr int TILES_PER_TEAM = 70, MAX_TILES = 2 * TILES_PER_TEAM;  // RLConst::Dropshot (RLConst.h:223-225)

enum GameMode : int { SOCCAR = 0, HOOPS = 1, HEATSEEKER = 2, SNOWDAY = 3, DROPSHOT = 4, THE_VOID = 5 };

// ---------------------------------------------------------------- persistent state
struct RigidS {
	V3 pos;      // world transform origin (Bullet units)
	M3 basis;    // world transform basis (row-major)
	V3 vel;      // linear velocity (bt/s)
	V3 angvel;   // rad/s
};

struct CtrlS {
	float throttle, steer, pitch, yaw, roll;
	uint8_t jump, boost, handbrake, _pad;
};

struct CarS {
	RigidS rb;
	// --- CarState game fields (Car.h:17-120) ---
	V3 flipRelTorque;
	V3 worldContactNormal;
	float jumpTime, flipTime, airTime, airTimeSinceJump;
	float boost, timeSinceBoosted, boostingTime, supersonicTime;
	float handbrakeVal, autoFlipTimer, autoFlipTorqueScale;
	float carContactCooldown, demoRespawnTimer;
	uint32_t carContactOtherID;
	uint8_t isOnGround, hasJumped, hasDoubleJumped, hasFlipped;
	uint8_t isFlipping, isJumping, isBoosting, isSupersonic;
	uint8_t isAutoFlipping, worldContactHas, isDemoed, bhValid;
	uint8_t wheelsWithContact[4];
	// ballHitInfo (BallHitInfo.h:9-28), vectors in UU
	V3 bhRelPos, bhBallPos, bhExtraVel;
	uint64_t bhTickHit, bhTickExtra;
	CtrlS lastControls;
	CtrlS controls;  // Car::controls (Car.h:149)
	uint64_t tickCountSinceUpdate;
	// --- state the reference keeps outside CarState ---
	float wSteer[4], wEngine[4], wBrake[4], wLatFriction[4], wLongFriction[4], wExtraPushback[4];
	V3 velImpulseCache;  // Car::_velocityImpulseCache, Bullet units
};

struct BallS {
	RigidS rb;
	float hsYTargetDir, hsCurTargetSpeed, hsTimeSinceHit;  // BallState::HeatseekerInfo
	V3 velImpulseCache;  // Ball::_velocityImpulseCache, Bullet units
	uint64_t tickCountSinceUpdate;
	uint8_t groundStickApplied;  // Ball::_groundStickApplied (Ball.h:82)
	uint8_t _pad[7];
};

struct PadS {
	float cooldown;
	uint32_t prevLockedCarID;
	uint8_t isActive;
	uint8_t _pad[3];
};

struct ArenaS {
	uint64_t tickCount;
	uint32_t rng;          // minstd_rand0 state of the respawn RNG (Math.cpp:44-64)
	uint32_t stampCounter; // broadphase cell-history counter
	// btRSBroadphase dynamic proxies: cell of the AABB min corner and the order in which
	// bodies last (re-)entered their 27-cell neighbourhood (btRSBroadphase.cpp:184-202,286-323).
	// dynOrder lists the bodies by ascending stamp (= the order of every cell's dynHandles vector),
	// cellI/J/K are the grid coordinates of cellIdx.
	int32_t cellIdx[MAX_BODIES];
	uint32_t cellStamp[MAX_BODIES];
	uint8_t dynOrder[MAX_BODIES + 1];
	uint8_t cellI[MAX_BODIES + 1], cellJ[MAX_BODIES + 1], cellK[MAX_BODIES + 1];
	uint32_t errorFlags;
	uint32_t goalScoredTick;  // debug
	// GameEventTracker persistent info (GameEventTracker.h:72-87); only advanced when the batch enabled the tracker
	uint64_t trkLastBallUpdate;
	float trkShotCooldown;
	uint8_t trkBallShot, trkShotGoalTeam, trkScoredLast;
	uint8_t worldFresh;    // 1 until the arena's physics world has stepped once (a new or freshly deserialised arena)
	BallS ball;
	PadS pads[MAX_PADS];
	CarS cars[MAX_CARS];   // last of the staged part: a batch with fewer cars stages only the used prefix of the record
	// ---- cold tail: never staged into shared memory, touched through Ctx::g (the record in HBM) ----
	// CarState::rotMat as Car::GetState reports it while a car is demoed: the reference refreshes _internalState.rotMat only at the
	// top of Car::_PostTickUpdate, which returns early for a demoed car (Car.cpp:133-138), while the victim's rigid body still
	// goes through the solve + integration of the tick it was demoed in.  Captured at the moment of the demo.
	M3 demoRotMat[MAX_CARS];
	// Dropshot (mode DROPSHOT only): BallState::dsInfo (Ball.h:34-47) and the tile field.  tiles[team * 70 + i]: bits 0-1 =
	// DropshotTileState::damageState (Arena::_dropshotTilesState), bit 2 = the tile's rigid body currently has
	// CF_NO_CONTACT_RESPONSE.  The two differ between the tick a tile breaks and the end of that tick (Arena.cpp:741-746).
	int32_t dsChargeLevel;
	float dsAccumulatedHitForce, dsYTargetDir;
	uint32_t dsHasDamaged;
	uint64_t dsLastDamageTick;
	uint8_t tiles[MAX_TILES];
	uint8_t _padTiles[4];
};

enum ErrorFlags : uint32_t {
	ERR_MANIFOLD_OVERFLOW = 1u, ERR_NAN = 2u, ERR_EVENT_OVERFLOW = 4u, ERR_BVH_STACK = 8u, ERR_EPA_FAIL = 16u,
	ERR_TRI_OVERFLOW = 32u,
};

// ---------------------------------------------------------------- batch constants
struct CarCfgD {            // derived from CarConfig (CarConfig.h:21-38) in Car::_BulletSetup (Car.cpp:195-292)
	V3 hitboxHalf;          // (hitboxSize * UU_TO_BT) / 2 as passed to btBoxShape
	V3 hitboxImplicit;      // hitboxHalf - 0.04 (btBoxShape.cpp:17-26: shrunk by the DEFAULT margin)
	float hitboxMargin;     // min(0.04, 0.1 * min half extent) (setSafeMargin, btConvexInternalShape.h:63-78)
	V3 hitboxOffset;        // compound child origin
	V3 invInertiaLocal;
	V3 wheelConn[4];        // chassisConnectionPointCS
	float wheelRadius[4];
	float suspRest[4];      // suspensionRestLength1 (config rest - travel), Bullet units
	float forceScale[4];
	float dodgeDeadzone;
	int threeWheels;
};

struct PadCfgD {
	V3 posBT, boxMin, boxMax;
	float cylRadBT;
	int isBig;
};

struct MutD {  // MutatorConfig (MutatorConfig.h:16-79), raw values in reference units
	V3 gravity;
	float carMass, carWorldFriction, carWorldRestitution;
	float ballMass, ballMaxSpeed, ballDrag, ballWorldFriction, ballWorldRestitution;
	float jumpAccel, jumpImmediateForce, boostAccelGround, boostAccelAir, boostUsedPerSecond;
	float respawnDelay, bumpCooldownTime, boostPadCooldown_Big, boostPadCooldown_Small, carSpawnBoostAmount;
	float ballHitExtraForceScale, bumpForceScale, ballRadius;
	float rechargeBoostPerSecond, rechargeBoostDelay, goalBaseThresholdY;
	int unlimitedFlips, unlimitedDoubleJumps, rechargeBoostEnabled, demoMode, enableTeamDemos;
};

constexpr int PUCK_POINTS = 40;

struct TrackerCfgD { float shotMinSpeed, shotTouchMinDelay, predScoreExtraMargin, shotEventCooldown, shotMinScoreTime, goalMaxTouchTime, passMaxTouchTime; };

struct BatchParams {
	int mode, nCars, nPads, nArenas;
	float dt;
	MutD mut;
	V3 gravityBT;
	// ball rigid body / shape (Ball.cpp:53-115)
	int ballIsSphere, ballNoRot;
	float ballRadiusBT, ballInvMass, ballMass;
	V3 ballInvInertiaLocal;
	V3 puckPts[PUCK_POINTS];
	V3 puckLocalAabbMin, puckLocalAabbMax;  // recalcLocalAabb result
	CarCfgD carCfg[MAX_CARS];
	int team[MAX_CARS];
	PadCfgD pads[MAX_PADS];
	int padGrid[8 * 10];  // BoostPadGrid (BoostPadGrid.h:10-29): pad index or -1
	// powf results whose arguments are batch constants, computed on the HOST (glibc, like the reference): the ball's linear damping
	// factor powf(1 - ballDrag, dt) (btRigidBody::applyDamping) and the flip z-damp factor powf(1 - FLIP_Z_DAMP_120, dt * 120)
	float ballLinDampFactor, flipZDampFactor;
	int customPads;       // ArenaConfig::useCustomBoostPads: no grid, every pad is checked against every car (Arena.cpp:724-730)
	// README bench loop on device: when set, a goal is followed by ResetToRandomKickoff(autoResetSeed + arena + tick)
	int autoReset; uint32_t autoResetSeed;
	// spawn orientations [team][slot], computed on the HOST at batch creation: the reference builds them with glibc's sinf / cosf
	// (Angle::ToRotMat), which the device libm only matches to 1 ulp -- a kickoff or respawn must start bit-identical
	M3 kickBasis[2][5], respBasis[2][4];
	int arenaOffset;  // global index of the batch's arena 0 in a sharded job (auto-reset seeds and event arena ids are global)
	int trackerEnabled; TrackerCfgD tracker;  // device-side GameEventTracker (rsb_set_event_tracker)
	float boundRadius[MAX_BODIES]; // radius of a sphere around the shape centre that contains the shape incl. margins
	float motionDisc[MAX_BODIES];  // btCollisionShape::getAngularMotionDisc of each dynamic body's shape (constant per batch)
};

// ---------------------------------------------------------------- static scene (per game mode)
struct BvhNode { uint16_t qmin[3]; uint16_t qmax[3]; int32_t idx; };  // btQuantizedBvhNode (16 B)
struct SubtreeHdr { uint16_t qmin[3]; uint16_t qmax[3]; int32_t root; int32_t size; };
struct TriInfo { int32_t flags; float a01, a12, a20; };  // btTriangleInfo, flags = -1 when absent

struct MeshD {
	const BvhNode* nodes; int nNodes;
	const float* nodeBounds;  // per node: unquantized box, min xyz_ max xyz_ (8 floats), = btQuantizedBvh::unQuantize of the node's corners, precomputed at scene build
	const SubtreeHdr* hdrs; int nHdrs;
	const int32_t* tris;   // 3 indices per triangle
	const float* verts;    // 4 floats per vertex (16-byte stride like btVector3)
	const TriInfo* info;
	int nTris, nVerts;
	V3 bvhMin, bvhMax, bvhQuant;  // quantization frame
	V3 aabbMin, aabbMax;          // shape local AABB (= world: meshes sit at the origin)
	int group, mask;              // collision filter (HOOPS_NET for the 505-vertex mesh)
};

struct PlaneD { V3 n; float d; V3 pos; V3 aabbMin, aabbMax; int group, mask; };  // world plane n.x = d ; body origin pos ; collision filter
// Dropshot floor tile (DropshotTiles.cpp:22-62, Arena.cpp:1070-1085): a btConvexHullShape of 6 coplanar points given in WORLD
// coordinates (Bullet units), on a static body at (0, 0, -margin); nb1 / nb2 = the tiles damaged by a charge-level 2 / 3 hit
// (indices within the tile's team, ascending, incl. the tile itself; DropshotTiles.cpp:92-112)
struct TileD {
	V3 pts[6];
	V3 aabbMin, aabbMax;  // world AABB of the shape (getAabb of the body's transform)
	float posY;           // GetTilePos(team, idx).y (uu)
	uint8_t n1, n2, nb1[7], nb2[20];
	uint8_t _pad[3];
};

constexpr int MAX_MESHES = 20;
constexpr int MAX_PLANES = 6;
constexpr int MAX_STATICS = MAX_MESHES + MAX_PLANES + MAX_TILES;  // static object index: meshes, then planes, then tiles (world order)

struct Scene {
	int nMeshes, nPlanes;
	int nTiles; const TileD* tiles;  // Dropshot only
	MeshD meshes[MAX_MESHES];
	PlaneD planes[MAX_PLANES];
	// btRSBroadphase grid (btRSBroadphase.cpp:43-84): static object lists per cell, CSR
	V3 gridMin, gridMax;
	float cellSize;
	int cellsX, cellsY, cellsZ;
	const int32_t* cellStart;   // [totalCells+1]
	const uint8_t* cellStatics; // static object indices (0..nMeshes-1 meshes, then planes)
	// per cellStatics entry of a mesh: candidate BVH subtree headers for queries issued from that cell
	// (creation order for AABB queries, root order for ray queries); see rsb_scene.h
	const int32_t* cellHdrStart;
	const uint16_t* cellHdrs;
	const uint16_t* cellHdrsRay;
	// per cellStatics entry of a mesh: the LEAF nodes (ascending = BVH leaf order) that a ray of length <= rayLeafReach starting
	// inside the cell can reach (node box vs the cell box grown by the reach).  A leaf is reported by the reference's walk
	// exactly when the leaf's own quantised-box and ray-box tests pass (both tests are monotone under box containment, and
	// every ancestor box contains the leaf box), so testing the listed leaves in order reproduces the walk (rsb_ray.h).
	const int32_t* cellRayLeafStart;
	const int32_t* cellRayLeaves;
	float rayLeafReach;
};

}  // namespace rsb
