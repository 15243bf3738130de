"""numpy structured dtypes mirroring ``include/rsb_types.h`` byte for byte.

These are the Python view of the drop-in boundary PODs (which are themselves layout
compatible with the reference structs, e.g. ``CarState`` at reference
``src/Sim/Car/Car.h:17-120``).  ``tests/test_abi.py`` checks every itemsize/offset against
the C header through the loaded shared library.
"""
import numpy as np

Vec = np.dtype({"names": ["x", "y", "z", "_w"], "formats": ["<f4"] * 4, "itemsize": 16})
RotMat = np.dtype({"names": ["forward", "right", "up"], "formats": [Vec] * 3, "itemsize": 48})

CarControls = np.dtype({
    "names": ["throttle", "steer", "pitch", "yaw", "roll", "jump", "boost", "handbrake"],
    "formats": ["<f4"] * 5 + ["u1"] * 3,
    "offsets": [0, 4, 8, 12, 16, 20, 21, 22],
    "itemsize": 24,
})

BallHitInfo = np.dtype({
    "names": ["isValid", "relativePosOnBall", "ballPos", "extraHitVel", "tickCountWhenHit",
              "tickCountWhenExtraImpulseApplied"],
    "formats": ["u1", Vec, Vec, Vec, "<u8", "<u8"],
    "offsets": [0, 16, 32, 48, 64, 72],
    "itemsize": 80,
})

WorldContact = np.dtype({"names": ["hasContact", "contactNormal"], "formats": ["u1", Vec],
                         "offsets": [0, 16], "itemsize": 32})
CarContact = np.dtype({"names": ["otherCarID", "cooldownTimer"], "formats": ["<u4", "<f4"],
                       "offsets": [0, 4], "itemsize": 8})

CarState = np.dtype({
    "names": ["pos", "rotMat", "vel", "angVel", "tickCountSinceUpdate", "isOnGround", "wheelsWithContact",
              "hasJumped", "hasDoubleJumped", "hasFlipped", "flipRelTorque", "jumpTime", "flipTime",
              "isFlipping", "isJumping", "airTime", "airTimeSinceJump", "boost", "timeSinceBoosted",
              "isBoosting", "boostingTime", "isSupersonic", "supersonicTime", "handbrakeVal",
              "isAutoFlipping", "autoFlipTimer", "autoFlipTorqueScale", "worldContact", "carContact",
              "isDemoed", "demoRespawnTimer", "ballHitInfo", "lastControls"],
    "formats": [Vec, RotMat, Vec, Vec, "<u8", "u1", ("u1", 4),
                "u1", "u1", "u1", Vec, "<f4", "<f4",
                "u1", "u1", "<f4", "<f4", "<f4", "<f4",
                "u1", "<f4", "u1", "<f4", "<f4",
                "u1", "<f4", "<f4", WorldContact, CarContact,
                "u1", "<f4", BallHitInfo, CarControls],
    "offsets": [0, 16, 64, 80, 96, 104, 105,
                109, 110, 111, 112, 128, 132,
                136, 137, 140, 144, 148, 152,
                156, 160, 164, 168, 172,
                176, 180, 184, 192, 224,
                232, 236, 240, 320],
    "itemsize": 352,
})

HsInfo = np.dtype({"names": ["yTargetDir", "curTargetSpeed", "timeSinceHit"], "formats": ["<f4"] * 3, "itemsize": 12})
DsInfo = np.dtype({"names": ["chargeLevel", "accumulatedHitForce", "yTargetDir", "hasDamaged", "lastDamageTick"],
                   "formats": ["<i4", "<f4", "<f4", "u1", "<u8"], "offsets": [0, 4, 8, 12, 16], "itemsize": 24})
BallState = np.dtype({
    "names": ["pos", "rotMat", "vel", "angVel", "tickCountSinceUpdate", "hsInfo", "dsInfo"],
    "formats": [Vec, RotMat, Vec, Vec, "<u8", HsInfo, DsInfo],
    "offsets": [0, 16, 64, 80, 96, 104, 120],
    "itemsize": 144,
})

MutatorConfig = np.dtype({
    "names": ["gravity", "carMass", "carWorldFriction", "carWorldRestitution", "ballMass", "ballMaxSpeed",
              "ballDrag", "ballWorldFriction", "ballWorldRestitution", "jumpAccel", "jumpImmediateForce",
              "boostAccelGround", "boostAccelAir", "boostUsedPerSecond", "respawnDelay", "bumpCooldownTime",
              "boostPadCooldown_Big", "boostPadCooldown_Small", "carSpawnBoostAmount", "ballHitExtraForceScale",
              "bumpForceScale", "ballRadius", "unlimitedFlips", "unlimitedDoubleJumps", "rechargeBoostEnabled",
              "rechargeBoostPerSecond", "rechargeBoostDelay", "demoMode", "enableTeamDemos", "goalBaseThresholdY"],
    "formats": [Vec] + ["<f4"] * 21 + ["u1"] * 3 + ["<f4", "<f4", "u1", "u1", "<f4"],
    "offsets": [0] + [16 + 4 * i for i in range(21)] + [100, 101, 102, 104, 108, 112, 113, 116],
    "itemsize": 128,
})

WheelPairConfig = np.dtype({"names": ["wheelRadius", "suspensionRestLength", "connectionPointOffset"],
                            "formats": ["<f4", "<f4", Vec], "offsets": [0, 4, 16], "itemsize": 32})
CarConfig = np.dtype({
    "names": ["hitboxSize", "hitboxPosOffset", "frontWheels", "backWheels", "threeWheels", "dodgeDeadzone"],
    "formats": [Vec, Vec, WheelPairConfig, WheelPairConfig, "u1", "<f4"],
    "offsets": [0, 16, 32, 64, 96, 100],
    "itemsize": 112,
})

BoostPadState = np.dtype({"names": ["isActive", "cooldown", "prevLockedCarID"], "formats": ["u1", "<f4", "<u4"],
                          "offsets": [0, 4, 8], "itemsize": 12})

BoostPadConfig = np.dtype({"names": ["pos", "isBig"], "formats": [Vec, "u1"], "offsets": [0, 16], "itemsize": 32})
# RsbArenaConfig (include/rsb_types.h); customBoostPads is a raw pointer filled in by ArenaBatch
ArenaConfig = np.dtype({"names": ["memWeightMode", "noBallRot", "useCustomBroadphase", "useCustomBoostPads", "maxAABBLen", "maxObjects",
                                  "numCustomBoostPads", "minPos", "maxPos", "customBoostPads", "_reserved"],
                        "formats": ["u1", "u1", "u1", "u1", "<f4", "<i4", "<u4", Vec, Vec, "<u8", "<u8"],
                        "offsets": [0, 1, 2, 3, 4, 8, 12, 16, 32, 48, 56], "itemsize": 64})

CarInternal = np.dtype({
    "names": ["steerAngle", "engineForce", "brake", "latFriction", "longFriction", "extraPushback",
              "velocityImpulseCache", "rbPos", "rbVel", "tickCountSinceUpdate", "hasRigidBody", "rbBasis"],
    "formats": [("<f4", 4)] * 9 + ["<u8", "<u4", ("<f4", 12)],
    "offsets": [16 * i for i in range(9)] + [144, 152, 160],
    "itemsize": 208,
})
BallInternal = np.dtype({"names": ["velocityImpulseCache", "groundStickApplied", "hasRigidBody", "tickCountSinceUpdate",
                                   "rbPos", "rbVel"],
                         "formats": [("<f4", 4), "u1", "<u4", "<u8", ("<f4", 4), ("<f4", 4)],
                         "offsets": [0, 16, 20, 24, 32, 48], "itemsize": 64})

Event = np.dtype({"names": ["arena", "type", "a", "b", "tick"], "formats": ["<u4", "<u4", "<u4", "<u4", "<u8"],
                  "itemsize": 24})

EVENT_GOAL, EVENT_BUMP, EVENT_DEMO = 1, 2, 3
EVENT_SHOT, EVENT_GOAL_CREDIT, EVENT_SAVE = 4, 5, 6
EventTrackerConfig = np.dtype([(n, "<f4") for n in ("shotMinSpeed", "shotTouchMinDelay", "predScoreExtraMargin", "shotEventCooldown",
                                                     "shotMinScoreTime", "goalMaxTouchTime", "passMaxTouchTime")])
MODE_SOCCAR, MODE_HOOPS, MODE_HEATSEEKER, MODE_SNOWDAY, MODE_DROPSHOT, MODE_THE_VOID = range(6)
TEAM_BLUE, TEAM_ORANGE = 0, 1
CAR_OCTANE, CAR_DOMINUS, CAR_PLANK, CAR_BREAKOUT, CAR_HYBRID, CAR_MERC, CAR_PSYCLOPS = range(7)

SIZES = {
    "RsbVec": 16, "RsbRotMat": 48, "RsbCarControls": 24, "RsbBallHitInfo": 80, "RsbCarState": 352,
    "RsbBallState": 144, "RsbMutatorConfig": 128, "RsbCarConfig": 112, "RsbBoostPadState": 12,
    "RsbCarInternal": 208, "RsbBoostPadConfig": 32, "RsbArenaConfig": 64, "RsbBallInternal": 64, "RsbEvent": 24,
}
