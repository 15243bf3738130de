/* rsb_types.h -- plain-C POD types of the drop-in boundary.
 *
 * Each struct is field-for-field and byte-for-byte layout compatible with the reference
 * type it replaces, so a RocketSim binding can pass its own objects straight through
 * (layouts verified against the compiled reference with offsetof probes; the sizes are
 * re-checked by static_asserts in oracle/ref_harness.cpp and tests/test_abi.py):
 *
 *   RsbVec          <- RocketSim::Vec          src/Math/MathTypes/MathTypes.h:7-21   (16 B, x y z _w)
 *   RsbRotMat       <- RocketSim::RotMat       src/Math/MathTypes/MathTypes.h:169-176 (forward,right,up columns)
 *   RsbCarControls  <- RocketSim::CarControls  src/Sim/CarControls.h:7-33            (24 B)
 *   RsbBallHitInfo  <- RocketSim::BallHitInfo  src/Sim/BallHitInfo/BallHitInfo.h:9-28 (80 B)
 *   RsbCarState     <- RocketSim::CarState     src/Sim/Car/Car.h:17-120              (352 B)
 *   RsbBallState    <- RocketSim::BallState    src/Sim/Ball/Ball.h:18-57             (144 B)
 *   RsbMutatorConfig<- RocketSim::MutatorConfig src/Sim/MutatorConfig/MutatorConfig.h:16-79 (128 B)
 *   RsbCarConfig    <- RocketSim::CarConfig    src/Sim/Car/CarConfig/CarConfig.h:21-38 (112 B)
 *
 * Units at this boundary are the reference's: positions/velocities in UU (1 uu = 0.02
 * Bullet units, src/BulletLink.h:12-15), angular velocity in rad/s.
 */
#ifndef RSB_TYPES_H
#define RSB_TYPES_H

#include <stdint.h>

#ifdef __cplusplus
#define RSB_ALIGN16 alignas(16)
#else
#define RSB_ALIGN16 _Alignas(16)
#endif

typedef struct RSB_ALIGN16 RsbVec { float x, y, z, _w; } RsbVec;
typedef struct RSB_ALIGN16 RsbRotMat { RsbVec forward, right, up; } RsbRotMat;

typedef struct RsbCarControls {
	float throttle, steer;
	float pitch, yaw, roll;
	uint8_t jump, boost, handbrake;
	uint8_t _pad;
} RsbCarControls;

typedef struct RsbBallHitInfo {
	uint8_t isValid;
	uint8_t _pad0[15];
	RsbVec relativePosOnBall;
	RsbVec ballPos;
	RsbVec extraHitVel;
	uint64_t tickCountWhenHit;
	uint64_t tickCountWhenExtraImpulseApplied;
} RsbBallHitInfo;

typedef struct RsbCarState {
	/* PhysState (src/Sim/PhysState/PhysState.h:6-21) */
	RsbVec pos;
	RsbRotMat rotMat;
	RsbVec vel;
	RsbVec angVel;
	/* CarState */
	uint64_t tickCountSinceUpdate;   /*  96 */
	uint8_t isOnGround;              /* 104 */
	uint8_t wheelsWithContact[4];    /* 105 */
	uint8_t hasJumped;               /* 109 */
	uint8_t hasDoubleJumped;         /* 110 */
	uint8_t hasFlipped;              /* 111 */
	RsbVec flipRelTorque;            /* 112 */
	float jumpTime;                  /* 128 */
	float flipTime;                  /* 132 */
	uint8_t isFlipping;              /* 136 */
	uint8_t isJumping;               /* 137 */
	uint8_t _pad0[2];
	float airTime;                   /* 140 */
	float airTimeSinceJump;          /* 144 */
	float boost;                     /* 148 */
	float timeSinceBoosted;          /* 152 */
	uint8_t isBoosting;              /* 156 */
	uint8_t _pad1[3];
	float boostingTime;              /* 160 */
	uint8_t isSupersonic;            /* 164 */
	uint8_t _pad2[3];
	float supersonicTime;            /* 168 */
	float handbrakeVal;              /* 172 */
	uint8_t isAutoFlipping;          /* 176 */
	uint8_t _pad3[3];
	float autoFlipTimer;             /* 180 */
	float autoFlipTorqueScale;       /* 184 */
	uint8_t _pad4[4];
	struct { uint8_t hasContact; uint8_t _pad[15]; RsbVec contactNormal; } worldContact; /* 192 */
	struct { uint32_t otherCarID; float cooldownTimer; } carContact;                     /* 224 */
	uint8_t isDemoed;                /* 232 */
	uint8_t _pad5[3];
	float demoRespawnTimer;          /* 236 */
	RsbBallHitInfo ballHitInfo;      /* 240 */
	RsbCarControls lastControls;     /* 320 */
	uint8_t _pad6[8];
} RsbCarState;

typedef struct RsbBallState {
	RsbVec pos;
	RsbRotMat rotMat;
	RsbVec vel;
	RsbVec angVel;
	uint64_t tickCountSinceUpdate;                                        /*  96 */
	struct { float yTargetDir, curTargetSpeed, timeSinceHit; } hsInfo;    /* 104 */
	uint8_t _pad0[4];
	struct { int32_t chargeLevel; float accumulatedHitForce; float yTargetDir;
	         uint8_t hasDamaged; uint8_t _pad[3]; uint64_t lastDamageTick; } dsInfo; /* 120 */
} RsbBallState;

enum { RSB_DEMO_NORMAL = 0, RSB_DEMO_ON_CONTACT = 1, RSB_DEMO_DISABLED = 2 };

typedef struct RsbMutatorConfig {
	RsbVec gravity;
	float carMass, carWorldFriction, carWorldRestitution;
	float ballMass, ballMaxSpeed, ballDrag, ballWorldFriction, ballWorldRestitution;
	float jumpAccel, jumpImmediateForce;
	float boostAccelGround, boostAccelAir, boostUsedPerSecond;
	float respawnDelay, bumpCooldownTime;
	float boostPadCooldown_Big, boostPadCooldown_Small;
	float carSpawnBoostAmount;                 /*  84 */
	float ballHitExtraForceScale, bumpForceScale; /* 88, 92 */
	float ballRadius;                          /*  96 */
	uint8_t unlimitedFlips, unlimitedDoubleJumps, rechargeBoostEnabled, _pad0; /* 100..103 */
	float rechargeBoostPerSecond, rechargeBoostDelay; /* 104, 108 */
	uint8_t demoMode, enableTeamDemos, _pad1[2];      /* 112 */
	float goalBaseThresholdY;                         /* 116 */
	uint8_t _pad2[8];
} RsbMutatorConfig;

typedef struct RsbWheelPairConfig {
	float wheelRadius, suspensionRestLength;
	uint8_t _pad[8];
	RsbVec connectionPointOffset;
} RsbWheelPairConfig;

typedef struct RsbCarConfig {
	RsbVec hitboxSize;
	RsbVec hitboxPosOffset;
	RsbWheelPairConfig frontWheels, backWheels;
	uint8_t threeWheels;
	uint8_t _pad0[3];
	float dodgeDeadzone;
	uint8_t _pad1[8];
} RsbCarConfig;

/* Boost pad state.  The reference struct (src/Sim/BoostPad/BoostPad.h:23-32) holds a raw
 * Car* (curLockedCar, intra-tick only); across the boundary we carry the three
 * serialised fields (BOOSTPAD_SERIALIZATION_FIELDS, BoostPad.h:33-34). */
typedef struct RsbBoostPadState {
	uint8_t isActive;
	uint8_t _pad[3];
	float cooldown;
	uint32_t prevLockedCarID;
} RsbBoostPadState;

/* Game modes: values of RocketSim::GameMode (src/Sim/GameMode.h:6-17). */
enum { RSB_MODE_SOCCAR = 0, RSB_MODE_HOOPS = 1, RSB_MODE_HEATSEEKER = 2, RSB_MODE_SNOWDAY = 3,
       RSB_MODE_DROPSHOT = 4, RSB_MODE_THE_VOID = 5 };
enum { RSB_TEAM_BLUE = 0, RSB_TEAM_ORANGE = 1 };
/* Car presets, order of src/Sim/Car/CarConfig/CarConfig.cpp:20-100 */
enum { RSB_CAR_OCTANE = 0, RSB_CAR_DOMINUS, RSB_CAR_PLANK, RSB_CAR_BREAKOUT, RSB_CAR_HYBRID, RSB_CAR_MERC, RSB_CAR_PSYCLOPS };

/* State that the reference carries between ticks OUTSIDE CarState:
 * btWheelInfoRL fields consumed one tick late (src/Sim/btVehicleRL/btVehicleRL.cpp:76,
 * 340-374; written at src/Sim/Car/Car.cpp:396-416,466-467), m_extraPushback
 * (btVehicleRL.cpp:181-198,205) and Car::_velocityImpulseCache (Car.h:188).
 * Exposed for exact mid-rollout state transfer in parity tests. */
typedef struct RsbCarInternal {
	float steerAngle[4];
	float engineForce[4];
	float brake[4];
	float latFriction[4];
	float longFriction[4];
	float extraPushback[4];
	float velocityImpulseCache[4]; /* Bullet units */
	/* exact rigid-body position / linear velocity in Bullet units (CarState carries them x50,
	 * which does not round-trip bit-exactly); applied by the setter when hasRigidBody != 0 */
	float rbPos[4];
	float rbVel[4];
	uint64_t tickCountSinceUpdate;
	uint32_t hasRigidBody;  /* bit 0: rbPos / rbVel / tickCountSinceUpdate are valid; bit 1: rbBasis is valid */
	uint32_t _pad;
	/* exact rigid-body basis, 3 rows x 4 floats.  Equal to CarState::rotMat except while a car is demoed: then the state keeps
	 * the rotMat of the last post-tick update (Car.cpp:133-138) while the body keeps the pose it was integrated to. */
	float rbBasis[12];
} RsbCarInternal;

typedef struct RsbBallInternal {
	float velocityImpulseCache[4]; /* Bullet units */
	uint8_t groundStickApplied;
	uint8_t _pad[3];
	uint32_t hasRigidBody;
	uint64_t tickCountSinceUpdate;
	float rbPos[4];
	float rbVel[4];
} RsbBallInternal;

/* BoostPadConfig (src/Sim/BoostPad/BoostPad.h:12-18): same layout (Vec pos; bool isBig; padded to 32 bytes). */
typedef struct RsbBoostPadConfig {
	RsbVec pos;
	uint8_t isBig;
	uint8_t _pad[15];
} RsbBoostPadConfig;

/* ArenaConfig (src/Sim/Arena/ArenaConfig/ArenaConfig.h:18-50) as a POD: same fields, the std::vector of custom pads as
 * pointer + count.  rsb_default_arena_config() fills in the reference's defaults.  What the batched back end supports:
 *   noBallRot            both values (ignored in Snowday, as in the reference)
 *   useCustomBoostPads   up to 40 pads; like the reference this replaces the pad grid by an all-pads check per car
 *   maxObjects           accepted, unused (capacity of the reference's proxy array)
 *   memWeightMode, minPos, maxPos, maxAABBLen, useCustomBroadphase   must keep their default values: the broadphase grid is
 *                        part of the static scene built once per game mode (rsb_init_from_mem); other values are rejected
 *                        with an error that names the field. */
enum { RSB_ARENA_MEM_HEAVY = 0, RSB_ARENA_MEM_LIGHT = 1 };
typedef struct RsbArenaConfig {
	uint8_t memWeightMode;
	uint8_t noBallRot;
	uint8_t useCustomBroadphase;
	uint8_t useCustomBoostPads;
	float maxAABBLen;
	int32_t maxObjects;
	uint32_t numCustomBoostPads;
	RsbVec minPos, maxPos;
	const RsbBoostPadConfig* customBoostPads;
	uint64_t _reserved;
} RsbArenaConfig;

/* Events produced by a step (goal callback Arena.cpp:748-752, bump callback :399-400). */
enum { RSB_EVENT_GOAL = 1, RSB_EVENT_BUMP = 2, RSB_EVENT_DEMO = 3,
       /* GameEventTracker events (src/Sim/GameEventTracker/GameEventTracker.h:6-8), produced when the tracker is enabled:
        * a = shooter / scorer / saver car id, b = passer car id or 0 */
       RSB_EVENT_SHOT = 4, RSB_EVENT_GOAL_CREDIT = 5, RSB_EVENT_SAVE = 6 };

/* GameEventTrackerConfig (src/Sim/GameEventTracker/GameEventTracker.h:10-38), same field order and defaults. */
typedef struct RsbEventTrackerConfig {
	float shotMinSpeed, shotTouchMinDelay, predScoreExtraMargin, shotEventCooldown, shotMinScoreTime, goalMaxTouchTime, passMaxTouchTime;
} RsbEventTrackerConfig;
typedef struct RsbEvent {
	uint32_t arena;
	uint32_t type;     /* RSB_EVENT_* */
	uint32_t a;        /* goal: scoring team; bump/demo: bumper car id */
	uint32_t b;        /* bump/demo: victim car id */
	uint64_t tick;     /* arena tickCount at which it happened */
} RsbEvent;

#endif
