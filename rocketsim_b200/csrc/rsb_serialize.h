// rsb_serialize.h -- RocketSim's arena wire format (host side), so that a single arena of a batch can be handed to / taken
// from the CPU library.  Restates the layout written by
//   Arena::Serialize / DeserializeNew          src/Sim/Arena/Arena.cpp:546-644
//   ArenaConfig::Serialize                      src/Sim/Arena/ArenaConfig/ArenaConfig.cpp:5-14, fields ArenaConfig.h:52-53
//   Car::Serialize / CarState::Serialize        src/Sim/Car/Car.cpp:308-337, fields Car.h:122-128, CarControls.h:35-38, CarConfig.h:18-43
//   BallHitInfo::Serialize                      src/Sim/BallHitInfo/BallHitInfo.cpp:5-10, fields BallHitInfo.h:29-30
//   BoostPadState / BallState / MutatorConfig   BoostPad.h:33-34, Ball.h:59-62, MutatorConfig.cpp:42-55 + MutatorConfig.h:81-88
//   DataStreamOut::WriteMultiple                src/DataStream/DataStreamOut.h:35-44: u32 field count, then each field's raw
//                                               little-endian bytes (Vec = 16 B incl. the 4th lane, RotMat = 48 B, bool = 1 B)
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/rsb_types.h"

namespace rsbser {

struct Writer {
	std::vector<uint8_t> d;
	void raw(const void* p, size_t n) { const uint8_t* b = (const uint8_t*)p; d.insert(d.end(), b, b + n); }
	template <class T> void put(const T& v) { raw(&v, sizeof(T)); }
	void count(uint32_t n) { put<uint32_t>(n); }
	void b(bool v) { uint8_t x = v ? 1 : 0; raw(&x, 1); }
};
struct Reader {
	const uint8_t* p; size_t n, pos = 0;
	Reader(const uint8_t* p, size_t n) : p(p), n(n) {}
	void raw(void* out, size_t k) { if (pos + k > n) throw std::runtime_error("Arena::Deserialize(): input data overflown"); memcpy(out, p + pos, k); pos += k; }
	template <class T> T get() { T v; raw(&v, sizeof(T)); return v; }
	bool b() { return get<uint8_t>() != 0; }
	void count(uint32_t expect) {
		uint32_t k = get<uint32_t>();
		if (k != expect) throw std::runtime_error("DataStreamIn::ReadMultipleFromList(): Prop count mismatch, expected " + std::to_string(expect) + " but have " + std::to_string(k) + ".");
	}
};

struct ArenaImage {  // everything the format carries for one arena
	uint8_t gameMode = 0; float tickTime = 0; uint64_t tickCount = 0; uint32_t lastCarID = 0;
	RsbVec minPos{-5600, -6000, 0, 0}, maxPos{5600, 6000, 2200, 0}; float maxAABBLen = 370; bool noBallRot = true, useCustomBroadphase = true;
	struct CarImg { uint8_t team; uint32_t id; RsbCarControls controls; RsbCarConfig config; RsbCarState state; };
	std::vector<CarImg> cars;
	std::vector<RsbBoostPadState> pads;
	RsbBallState ball;
	RsbMutatorConfig mut;
};

inline void write_controls(Writer& w, const RsbCarControls& c) {
	w.count(8); w.put(c.throttle); w.put(c.steer); w.put(c.pitch); w.put(c.yaw); w.put(c.roll); w.b(c.boost); w.b(c.jump); w.b(c.handbrake);
}
inline void read_controls(Reader& r, RsbCarControls& c) {
	r.count(8); c.throttle = r.get<float>(); c.steer = r.get<float>(); c.pitch = r.get<float>(); c.yaw = r.get<float>(); c.roll = r.get<float>();
	c.boost = r.b(); c.jump = r.b(); c.handbrake = r.b(); c._pad = 0;
}
inline void write_car_config(Writer& w, const RsbCarConfig& c) {
	w.count(9); w.put(c.dodgeDeadzone); w.put(c.hitboxPosOffset); w.put(c.hitboxSize);
	w.put(c.frontWheels.connectionPointOffset); w.put(c.frontWheels.suspensionRestLength); w.put(c.frontWheels.wheelRadius);
	w.put(c.backWheels.connectionPointOffset); w.put(c.backWheels.suspensionRestLength); w.put(c.backWheels.wheelRadius);
}
inline void read_car_config(Reader& r, RsbCarConfig& c) {
	memset(&c, 0, sizeof(c));
	r.count(9); c.dodgeDeadzone = r.get<float>(); c.hitboxPosOffset = r.get<RsbVec>(); c.hitboxSize = r.get<RsbVec>();
	c.frontWheels.connectionPointOffset = r.get<RsbVec>(); c.frontWheels.suspensionRestLength = r.get<float>(); c.frontWheels.wheelRadius = r.get<float>();
	c.backWheels.connectionPointOffset = r.get<RsbVec>(); c.backWheels.suspensionRestLength = r.get<float>(); c.backWheels.wheelRadius = r.get<float>();
}
inline void write_car_state(Writer& w, const RsbCarState& s) {
	w.b(s.ballHitInfo.isValid);
	if (s.ballHitInfo.isValid) { w.count(5); w.put(s.ballHitInfo.relativePosOnBall); w.put(s.ballHitInfo.ballPos); w.put(s.ballHitInfo.extraHitVel); w.put(s.ballHitInfo.tickCountWhenHit); w.put(s.ballHitInfo.tickCountWhenExtraImpulseApplied); }
	w.count(30);
	w.put(s.pos); w.put(s.rotMat); w.put(s.vel); w.put(s.angVel); w.b(s.isOnGround); w.b(s.hasJumped); w.b(s.hasDoubleJumped); w.b(s.hasFlipped);
	w.put(s.flipRelTorque); w.put(s.jumpTime); w.b(s.isFlipping); w.put(s.flipTime); w.b(s.isJumping); w.put(s.airTimeSinceJump);
	w.put(s.boost); w.b(s.isBoosting); w.put(s.boostingTime); w.put(s.timeSinceBoosted); w.put(s.supersonicTime); w.put(s.handbrakeVal); w.b(s.isAutoFlipping);
	w.put(s.autoFlipTimer); w.put(s.autoFlipTorqueScale); w.b(s.isDemoed); w.put(s.demoRespawnTimer); w.put(s.lastControls);
	w.b(s.worldContact.hasContact); w.put(s.worldContact.contactNormal); w.put(s.carContact.otherCarID); w.put(s.carContact.cooldownTimer);
}
inline void read_car_state(Reader& r, RsbCarState& s) {
	s.ballHitInfo.isValid = r.b();
	if (s.ballHitInfo.isValid) { r.count(5); s.ballHitInfo.relativePosOnBall = r.get<RsbVec>(); s.ballHitInfo.ballPos = r.get<RsbVec>(); s.ballHitInfo.extraHitVel = r.get<RsbVec>(); s.ballHitInfo.tickCountWhenHit = r.get<uint64_t>(); s.ballHitInfo.tickCountWhenExtraImpulseApplied = r.get<uint64_t>(); }
	r.count(30);
	s.pos = r.get<RsbVec>(); s.rotMat = r.get<RsbRotMat>(); s.vel = r.get<RsbVec>(); s.angVel = r.get<RsbVec>(); s.isOnGround = r.b(); s.hasJumped = r.b(); s.hasDoubleJumped = r.b(); s.hasFlipped = r.b();
	s.flipRelTorque = r.get<RsbVec>(); s.jumpTime = r.get<float>(); s.isFlipping = r.b(); s.flipTime = r.get<float>(); s.isJumping = r.b(); s.airTimeSinceJump = r.get<float>();
	s.boost = r.get<float>(); s.isBoosting = r.b(); s.boostingTime = r.get<float>(); s.timeSinceBoosted = r.get<float>(); s.supersonicTime = r.get<float>(); s.handbrakeVal = r.get<float>(); s.isAutoFlipping = r.b();
	s.autoFlipTimer = r.get<float>(); s.autoFlipTorqueScale = r.get<float>(); s.isDemoed = r.b(); s.demoRespawnTimer = r.get<float>(); s.lastControls = r.get<RsbCarControls>();
	s.worldContact.hasContact = r.b(); s.worldContact.contactNormal = r.get<RsbVec>(); s.carContact.otherCarID = r.get<uint32_t>(); s.carContact.cooldownTimer = r.get<float>();
}
inline void write_ball(Writer& w, const RsbBallState& s) {
	w.count(12); w.put(s.pos); w.put(s.rotMat); w.put(s.vel); w.put(s.angVel); w.put(s.hsInfo.yTargetDir); w.put(s.hsInfo.curTargetSpeed); w.put(s.hsInfo.timeSinceHit);
	w.put(s.dsInfo.chargeLevel); w.put(s.dsInfo.accumulatedHitForce); w.put(s.dsInfo.yTargetDir); w.b(s.dsInfo.hasDamaged); w.put(s.dsInfo.lastDamageTick);
}
inline void read_ball(Reader& r, RsbBallState& s) {
	r.count(12); s.pos = r.get<RsbVec>(); s.rotMat = r.get<RsbRotMat>(); s.vel = r.get<RsbVec>(); s.angVel = r.get<RsbVec>(); s.hsInfo.yTargetDir = r.get<float>(); s.hsInfo.curTargetSpeed = r.get<float>(); s.hsInfo.timeSinceHit = r.get<float>();
	s.dsInfo.chargeLevel = r.get<int32_t>(); s.dsInfo.accumulatedHitForce = r.get<float>(); s.dsInfo.yTargetDir = r.get<float>(); s.dsInfo.hasDamaged = r.b(); s.dsInfo.lastDamageTick = r.get<uint64_t>();
}
inline void write_mutators(Writer& w, const RsbMutatorConfig& m) {
	w.put<uint16_t>(30); w.count(30);
	w.put(m.gravity); w.put(m.carMass); w.put(m.carWorldFriction); w.put(m.carWorldRestitution); w.put(m.ballMass);
	w.put(m.ballMaxSpeed); w.put(m.ballDrag); w.put(m.ballWorldFriction); w.put(m.ballWorldRestitution); w.put(m.jumpAccel);
	w.put(m.jumpImmediateForce); w.put(m.boostAccelGround); w.put(m.boostAccelAir); w.put(m.boostUsedPerSecond); w.put(m.respawnDelay);
	w.put(m.carSpawnBoostAmount); w.put(m.bumpCooldownTime); w.put(m.boostPadCooldown_Big); w.put(m.boostPadCooldown_Small);
	w.put(m.ballHitExtraForceScale); w.put(m.bumpForceScale); w.put(m.ballRadius); w.b(m.unlimitedFlips); w.b(m.unlimitedDoubleJumps);
	w.b(m.rechargeBoostEnabled); w.put(m.rechargeBoostPerSecond); w.put(m.rechargeBoostDelay); w.put<uint8_t>(m.demoMode); w.b(m.enableTeamDemos); w.put(m.goalBaseThresholdY);
}
inline void read_mutators(Reader& r, RsbMutatorConfig& m) {
	memset(&m, 0, sizeof(m));
	if (r.get<uint16_t>() != 30) throw std::runtime_error(" MutatorConfig::Deserialize(): Mutator config is from a different version of RocketSim, fields don't match");
	r.count(30);
	m.gravity = r.get<RsbVec>(); m.carMass = r.get<float>(); m.carWorldFriction = r.get<float>(); m.carWorldRestitution = r.get<float>(); m.ballMass = r.get<float>();
	m.ballMaxSpeed = r.get<float>(); m.ballDrag = r.get<float>(); m.ballWorldFriction = r.get<float>(); m.ballWorldRestitution = r.get<float>(); m.jumpAccel = r.get<float>();
	m.jumpImmediateForce = r.get<float>(); m.boostAccelGround = r.get<float>(); m.boostAccelAir = r.get<float>(); m.boostUsedPerSecond = r.get<float>(); m.respawnDelay = r.get<float>();
	m.carSpawnBoostAmount = r.get<float>(); m.bumpCooldownTime = r.get<float>(); m.boostPadCooldown_Big = r.get<float>(); m.boostPadCooldown_Small = r.get<float>();
	m.ballHitExtraForceScale = r.get<float>(); m.bumpForceScale = r.get<float>(); m.ballRadius = r.get<float>(); m.unlimitedFlips = r.b(); m.unlimitedDoubleJumps = r.b();
	m.rechargeBoostEnabled = r.b(); m.rechargeBoostPerSecond = r.get<float>(); m.rechargeBoostDelay = r.get<float>(); m.demoMode = r.get<uint8_t>(); m.enableTeamDemos = r.b(); m.goalBaseThresholdY = r.get<float>();
}

inline std::vector<uint8_t> write_arena(const ArenaImage& a) {
	Writer w;
	w.count(4); w.put(a.gameMode); w.put(a.tickTime); w.put(a.tickCount); w.put(a.lastCarID);
	w.count(5); w.put(a.minPos); w.put(a.maxPos); w.put(a.maxAABBLen); w.b(a.noBallRot); w.b(a.useCustomBroadphase);
	w.b(false);  // useCustomBoostPads
	w.put<uint32_t>((uint32_t)a.cars.size());
	for (const auto& c : a.cars) { w.put(c.team); w.put(c.id); write_controls(w, c.controls); write_car_config(w, c.config); write_car_state(w, c.state); }
	if (!a.pads.empty()) {
		w.put<uint32_t>((uint32_t)a.pads.size());
		for (const auto& p : a.pads) { w.count(3); w.b(p.isActive); w.put(p.cooldown); w.put(p.prevLockedCarID); }
	}
	write_ball(w, a.ball);
	write_mutators(w, a.mut);
	return w.d;
}

// nPadsOfMode: the pad count of an arena of the stream's game mode (the stream only carries a pad block when it is > 0)
template <class PadCountFn>
inline ArenaImage read_arena(const uint8_t* p, size_t n, PadCountFn nPadsOfMode) {
	Reader r(p, n);
	ArenaImage a;
	r.count(4); a.gameMode = r.get<uint8_t>(); a.tickTime = r.get<float>(); a.tickCount = r.get<uint64_t>(); a.lastCarID = r.get<uint32_t>();
	r.count(5); a.minPos = r.get<RsbVec>(); a.maxPos = r.get<RsbVec>(); a.maxAABBLen = r.get<float>(); a.noBallRot = r.b(); a.useCustomBroadphase = r.b();
	if (r.b()) throw std::runtime_error("Arena::Deserialize(): custom boost pads are not supported by the batched back end");
	const uint32_t nCars = r.get<uint32_t>();
	if (nCars > 64) throw std::runtime_error("Arena::Deserialize(): implausible car count");
	a.cars.resize(nCars);
	for (auto& c : a.cars) {
		c.team = r.get<uint8_t>(); c.id = r.get<uint32_t>();
		for (const auto& o : a.cars) if (&o != &c && &o < &c && o.id == c.id) throw std::runtime_error("Arena::Deserialize(): Failed to load, got repeated car ID of " + std::to_string(c.id));
		read_controls(r, c.controls); read_car_config(r, c.config); memset(&c.state, 0, sizeof(c.state)); read_car_state(r, c.state);
	}
	const int nPads = nPadsOfMode(a.gameMode);
	if (nPads > 0) {
		const uint32_t k = r.get<uint32_t>();
		if ((int)k != nPads) throw std::runtime_error("Arena::Deserialize(): Failed to load, different boost pad amount written in file (" + std::to_string(k) + "/" + std::to_string(nPads) + ")");
		a.pads.resize(k);
		for (auto& pd : a.pads) { r.count(3); memset(&pd, 0, sizeof(pd)); pd.isActive = r.b(); pd.cooldown = r.get<float>(); pd.prevLockedCarID = r.get<uint32_t>(); }
	}
	memset(&a.ball, 0, sizeof(a.ball)); read_ball(r, a.ball);
	read_mutators(r, a.mut);
	return a;
}

}  // namespace rsbser
