// rocketsim_b200.hpp -- header-only C++ front end over the C ABI (rsb_c_api.h), in the reference's vocabulary.
//
// Mirrors the part of RocketSim's public C++ surface that sits on the Arena::Step path:
//
//   RocketSim::Init / InitFromMem        src/RocketSim.h:27,32            -> RocketSimB200::Init / InitFromMem
//   Arena::Create / AddCar / GetCar      src/Sim/Arena/Arena.h:111,71,82  -> Arena::Create / AddCar / GetCar
//   Arena::Step / IsBallScored           Arena.h:132,144                  -> Arena::Step / IsBallScored
//   Arena::ResetToRandomKickoff          Arena.h (Arena.cpp:113-223)      -> Arena::ResetToRandomKickoff
//   Arena::Set/GetMutatorConfig          Arena.h:52-53                    -> same names
//   SetGoalScoreCallback / SetCarBumpCallback  Arena.h:24-25,102,108      -> same names, dispatched after Step() from the event buffer
//   Car::controls / GetState / SetState  src/Sim/Car/Car.h:149-152        -> same names
//   Ball::GetState / SetState            src/Sim/Ball/Ball.h:68-69        -> same names
//   tickCount, tickTime, GetTickRate()   Arena.h:56-64                    -> same names
//
// plus ArenaBatch, the batched front end (N identical arenas advanced in lockstep by one CUDA launch).  A facade
// `Arena` is an ArenaBatch of one arena.  All compute happens in librsb.so on the GPU; there is no CPU path, and
// errors surface as std::runtime_error like the reference's RS_ERR_CLOSE (src/Framework.h:67-72).
//
// Differences from the reference, by construction of the batched design:
//   * cars are fixed at the first Step/GetState/SetState of an arena (the device record is sized then): AddCar / RemoveCar
//     after that throw;
//   * callbacks run on the calling thread right after Step() returns the events of the stepped ticks (the reference
//     calls them from inside the tick); with Step(n > 1) they arrive in tick order after the n ticks;
//   * ArenaConfig: noBallRot and custom boost pads are honoured; fields that would change the broadphase grid
//     (memWeightMode, minPos, maxPos, maxAABBLen, useCustomBroadphase) must keep their defaults, other values throw.
#pragma once

#include <cstdint>
#include <cstring>
#include <algorithm>
#include <filesystem>
#include <fstream>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "rsb_c_api.h"

namespace RocketSimB200 {

using Vec = RsbVec;
using RotMat = RsbRotMat;
using CarControls = RsbCarControls;
using CarState = RsbCarState;
using BallState = RsbBallState;
using BallHitInfo = RsbBallHitInfo;
using MutatorConfig = RsbMutatorConfig;
using CarConfig = RsbCarConfig;
using BoostPadState = RsbBoostPadState;

enum class GameMode : uint8_t { SOCCAR = 0, HOOPS = 1, HEATSEEKER = 2, SNOWDAY = 3, DROPSHOT = 4, THE_VOID = 5 };
constexpr const char* GAMEMODE_STRS[] = {"soccar", "hoops", "heatseeker", "snowday", "dropshot", "void"};  // src/Sim/GameMode.h:19-26

// ArenaConfig (src/Sim/Arena/ArenaConfig/ArenaConfig.h:11-50), same fields and defaults
enum class ArenaMemWeightMode : uint8_t { HEAVY, LIGHT };
struct BoostPadConfig { Vec pos = {}; bool isBig = false; };
struct ArenaConfig {
	ArenaMemWeightMode memWeightMode = ArenaMemWeightMode::HEAVY;
	Vec minPos = Vec{-5600, -6000, 0, 0}, maxPos = Vec{5600, 6000, 2200, 0};
	float maxAABBLen = 370;
	bool noBallRot = true;
	bool useCustomBroadphase = true;
	int maxObjects = 512;
	bool useCustomBoostPads = false;
	std::vector<BoostPadConfig> customBoostPads = {};
};
enum class Team : uint8_t { BLUE = 0, ORANGE = 1 };

// DropshotTileState / DropshotTilesState (src/Sim/Arena/DropshotTiles/DropshotTiles.h:7-33): same layout, one byte per tile
struct DropshotTileState {
	enum { STATE_FULL = 0, STATE_DAMAGED, STATE_BROKEN };
	uint8_t damageState = STATE_FULL;
};
struct DropshotTilesState {
	enum { STATE_FULL = 0, STATE_DAMAGED, STATE_BROKEN };
	DropshotTileState states[2][70];  // [RLConst::Dropshot::TEAM_AMOUNT][NUM_TILES_PER_TEAM]
};
static_assert(sizeof(DropshotTilesState) == 140, "DropshotTilesState is 140 damage-state bytes");

inline void Check(int rc) { if (rc != 0) throw std::runtime_error(std::string("RocketSimB200: ") + rsb_last_error()); }

inline CarConfig CarConfigPreset(int preset) { CarConfig c; Check(rsb_car_config_preset(preset, &c)); return c; }
inline const CarConfig& CAR_CONFIG_OCTANE_() { static const CarConfig c = CarConfigPreset(RSB_CAR_OCTANE); return c; }
#define RSB200_CAR_CONFIG(name) inline const CarConfig& CAR_CONFIG_##name() { static const CarConfig c = CarConfigPreset(RSB_CAR_##name); return c; }
RSB200_CAR_CONFIG(OCTANE) RSB200_CAR_CONFIG(DOMINUS) RSB200_CAR_CONFIG(PLANK) RSB200_CAR_CONFIG(BREAKOUT)
RSB200_CAR_CONFIG(HYBRID) RSB200_CAR_CONFIG(MERC) RSB200_CAR_CONFIG(PSYCLOPS)
#undef RSB200_CAR_CONFIG

inline MutatorConfig DefaultMutatorConfig(GameMode mode) { MutatorConfig m; Check(rsb_default_mutator_config((int)mode, &m)); return m; }

// RocketSim::InitFromMem (src/RocketSim.cpp:92-185): .cmf bytes per game mode.
inline void InitFromMem(const std::map<GameMode, std::vector<std::vector<uint8_t>>>& meshFilesMap, bool /*silent*/ = false) {
	std::vector<int> modes; std::vector<const uint8_t*> ptrs; std::vector<size_t> sizes;
	for (auto& kv : meshFilesMap)
		for (auto& f : kv.second) { modes.push_back((int)kv.first); ptrs.push_back(f.data()); sizes.push_back(f.size()); }
	Check(rsb_init_from_mem((int)modes.size(), modes.data(), ptrs.data(), sizes.data()));
}

// ---------------------------------------------------------------------------------------------------------------
// ArenaBatch: N identical arenas (same mode, same car slots) on one GPU.
class ArenaBatch {
public:
	ArenaBatch(GameMode mode, int numArenas, const std::vector<Team>& teams, const std::vector<CarConfig>& configs = {},
	           const MutatorConfig* mutator = nullptr, float tickRate = 120.f, uint32_t seed = 0, int device = 0, const ArenaConfig& arenaConfig = {})
		: gameMode(mode), numArenas(numArenas), numCars((int)teams.size()) {
		std::vector<int> t; for (Team x : teams) t.push_back((int)x);
		if (!configs.empty() && configs.size() != teams.size()) throw std::runtime_error("RocketSimB200: one CarConfig per car slot expected");
		RsbArenaConfig ac; rsb_default_arena_config(&ac);
		ac.memWeightMode = (uint8_t)arenaConfig.memWeightMode; ac.minPos = arenaConfig.minPos; ac.maxPos = arenaConfig.maxPos; ac.maxAABBLen = arenaConfig.maxAABBLen;
		ac.noBallRot = arenaConfig.noBallRot; ac.useCustomBroadphase = arenaConfig.useCustomBroadphase; ac.maxObjects = arenaConfig.maxObjects;
		ac.useCustomBoostPads = arenaConfig.useCustomBoostPads;
		std::vector<RsbBoostPadConfig> pads;
		for (const BoostPadConfig& pc : arenaConfig.customBoostPads) { RsbBoostPadConfig r; std::memset(&r, 0, sizeof(r)); r.pos = pc.pos; r.isBig = pc.isBig; pads.push_back(r); }
		ac.customBoostPads = pads.empty() ? nullptr : pads.data(); ac.numCustomBoostPads = (uint32_t)pads.size();
		Check(rsb_batch_create_ex(device, (int)mode, &ac, numArenas, numCars, t.data(), configs.empty() ? nullptr : configs.data(), mutator, tickRate, seed, &h_));
		Check(rsb_batch_info(h_, nullptr, nullptr, &numPads, nullptr));
	}
	~ArenaBatch() { if (h_) rsb_batch_destroy(h_); }
	ArenaBatch(const ArenaBatch&) = delete;
	ArenaBatch& operator=(const ArenaBatch&) = delete;

	// `car->controls = ...` for every car: controls is [numArenas][numCars]
	void SetControls(const CarControls* controls) { Check(rsb_set_controls(h_, controls)); }
	void Step(int ticksToSimulate = 1) { Check(rsb_step(h_, ticksToSimulate)); }
	// for t: controls = seq[t]; Step(1)  -- one launch; seq is [nTicks][numArenas][numCars]
	void StepWithControls(const CarControls* seq, int nTicks) { Check(rsb_step_with_controls(h_, seq, nTicks)); }
	void Sync() { Check(rsb_sync(h_)); }
	void SetStream(void* cudaStream) { Check(rsb_set_stream(h_, cudaStream)); }

	std::vector<CarState> GetCarStates() { std::vector<CarState> v((size_t)numArenas * numCars); if (!v.empty()) Check(rsb_get_car_states(h_, v.data())); return v; }
	std::vector<BallState> GetBallStates() { std::vector<BallState> v(numArenas); Check(rsb_get_ball_states(h_, v.data())); return v; }
	std::vector<BoostPadState> GetPadStates() { std::vector<BoostPadState> v((size_t)numArenas * numPads); if (!v.empty()) Check(rsb_get_pad_states(h_, v.data())); return v; }
	void SetCarState(uint32_t arena, uint32_t carSlot, const CarState& s) { Check(rsb_set_car_states(h_, &arena, &carSlot, 1, &s)); }
	void SetBallState(uint32_t arena, const BallState& s) { Check(rsb_set_ball_states(h_, &arena, 1, &s)); }
	// Arena::GetDropshotTilesState / SetDropshotTilesState (Arena.h:178-179) of a Dropshot batch
	std::vector<DropshotTilesState> GetDropshotTilesStates() { std::vector<DropshotTilesState> v(numArenas); Check(rsb_get_dropshot_tiles(h_, (uint8_t*)v.data())); return v; }
	void SetDropshotTilesState(uint32_t arena, const DropshotTilesState& s) { Check(rsb_set_dropshot_tiles(h_, &arena, 1, (const uint8_t*)&s)); }
	std::vector<uint64_t> GetTickCounts() { std::vector<uint64_t> v(numArenas); Check(rsb_get_tick_counts(h_, v.data())); return v; }
	std::vector<uint8_t> IsBallScored() { std::vector<uint8_t> v(numArenas); Check(rsb_is_ball_scored(h_, v.data())); return v; }
	std::vector<RsbEvent> GetEvents(int cap = 65536) {
		std::vector<RsbEvent> all;
		for (;;) {  // rc 1 = more than `cap` events pending: nothing is dropped, fetch the next slice
			std::vector<RsbEvent> v(cap); int n = 0;
			const int rc = rsb_get_events(h_, v.data(), cap, &n);
			if (rc != 0 && rc != 1) Check(rc);
			all.insert(all.end(), v.begin(), v.begin() + n);
			if (rc == 0) break;
		}
		return all;
	}
	std::vector<uint32_t> GetErrorFlags() { std::vector<uint32_t> v(numArenas); Check(rsb_get_error_flags(h_, v.data())); return v; }
	// Arena::ResetToRandomKickoff(seed) for every arena (seeds[a]) or for the listed ones
	void ResetToRandomKickoff(const std::vector<int32_t>& seeds, const std::vector<uint32_t>* arenas = nullptr) {
		Check(rsb_reset_to_random_kickoff(h_, arenas ? arenas->data() : nullptr, arenas ? (int)arenas->size() : numArenas, seeds.data()));
	}
	void CloneArena(uint32_t src, uint32_t dst) { Check(rsb_clone_arena(h_, src, dst)); }  // Arena::Clone
	// Arena::Serialize / Arena::DeserializeNew of one arena in RocketSim's own byte format (interchange with the CPU library)
	std::vector<uint8_t> SerializeArena(uint32_t arena) {
		size_t n = 0; Check(rsb_serialize_arena(h_, arena, nullptr, 0, &n));
		std::vector<uint8_t> v(n); Check(rsb_serialize_arena(h_, arena, v.data(), v.size(), &n)); return v;
	}
	void DeserializeArena(uint32_t arena, const std::vector<uint8_t>& data) { Check(rsb_deserialize_arena(h_, arena, data.data(), data.size())); }
	// BallPredTracker::predData for every arena: [nTicks][numArenas], row 0 = now
	std::vector<BallState> PredictBall(int nTicks) { std::vector<BallState> v((size_t)nTicks * numArenas); Check(rsb_predict_ball(h_, nTicks, v.data())); return v; }
	void SetAutoReset(bool enable, uint32_t seed) { Check(rsb_set_auto_reset(h_, enable ? 1 : 0, seed)); }
	// GameEventTracker on the device (shots, goal credits, saves in GetEvents) and Arena::IsBallProbablyGoingIn
	void SetEventTracker(bool enable, const RsbEventTrackerConfig* cfg = nullptr) { Check(rsb_set_event_tracker(h_, enable ? 1 : 0, cfg)); }
	std::vector<uint8_t> IsBallProbablyGoingIn(float maxTime = 2.f, float extraMargin = 0, std::vector<uint8_t>* goalTeams = nullptr) {
		std::vector<uint8_t> v(numArenas), t(numArenas); Check(rsb_is_ball_probably_going_in(h_, maxTime, extraMargin, v.data(), t.data()));
		if (goalTeams) *goalTeams = t;
		return v;
	}
	RsbBatch* Handle() { return h_; }

	const GameMode gameMode;
	const int numArenas, numCars;
	int numPads = 0;

private:
	RsbBatch* h_ = nullptr;
};

// ---------------------------------------------------------------------------------------------------------------
// Single-arena facade with the reference's object model.
class Arena;

class Car {
public:
	CarControls controls = {};          // Car::controls (Car.h:149)
	uint32_t id = 0;                    // 1-based, AddCar order (Arena.cpp:64-78)
	Team team = Team::BLUE;
	CarConfig config = {};
	CarState GetState();                // Car::GetState (Car.cpp:10-20)
	void SetState(const CarState& s);   // Car::SetState (Car.cpp:23-36)
private:
	friend class Arena;
	Arena* arena_ = nullptr;
	uint32_t slot_ = 0;                 // car slot in the device record (differs from id - 1 after a RemoveCar)
};

class Ball {
public:
	BallState GetState();               // Ball::GetState (Ball.cpp:27-37)
	void SetState(const BallState& s);  // Ball::SetState (Ball.cpp:39-51)
private:
	friend class Arena;
	Arena* arena_ = nullptr;
};

typedef std::function<void(Arena* arena, Team scoringTeam, void* userInfo)> GoalScoreEventFn;
typedef std::function<void(Arena* arena, Car* bumper, Car* victim, bool isDemo, void* userInfo)> CarBumpEventFn;
// GameEventTracker callbacks (src/Sim/GameEventTracker/GameEventTracker.h:6-8)
typedef std::function<void(Arena* arena, Car* shooter, Car* passer, void* userInfo)> ShotEventFn;
typedef std::function<void(Arena* arena, Car* scorer, Car* passer, void* userInfo)> GoalEventFn;
typedef std::function<void(Arena* arena, Car* saver, void* userInfo)> SaveEventFn;

class Arena {
public:
	GameMode gameMode;
	Ball* ball;
	float tickTime;
	uint64_t tickCount = 0;
	float GetTickRate() const { return 1 / tickTime; }

	// Arena::Create(gameMode, arenaConfig = {}, tickRate = 120) (Arena.h:111); `device` / `rngSeed` are ours (which GPU, seed of
	// the respawn RNG that the reference takes from a thread-local engine).  The caller deletes the arena, as in the reference.
	static Arena* Create(GameMode gameMode, const ArenaConfig& arenaConfig = {}, float tickRate = 120, int device = 0, uint32_t rngSeed = 0) {
		if (!(tickRate >= 15 && tickRate <= 120)) throw std::runtime_error("RocketSimB200: Tickrate must be from 15 to 120tps");
		Arena* a = new Arena(gameMode, tickRate, device, rngSeed);
		a->config_ = arenaConfig;
		return a;
	}
	const ArenaConfig& GetArenaConfig() const { return config_; }
	~Arena() { delete ball; for (Car* c : cars_) delete c; }
	Arena(const Arena&) = delete;
	Arena& operator=(const Arena&) = delete;

	Car* AddCar(Team team, const CarConfig& config = CAR_CONFIG_OCTANE()) {
		if (batch_) throw std::runtime_error("RocketSimB200: AddCar after the arena's first Step/GetState/SetState is not supported by the batched back end");
		Car* c = new Car(); c->id = ++lastCarID_; c->team = team; c->config = config; c->arena_ = this;  // ids keep counting up (Arena.cpp:64-78)
		cars_.push_back(c);
		return c;
	}
	// Arena::RemoveCar (Arena.h:74, Arena.cpp:80-94): false if the id is unknown.  Only before the arena's first Step.
	bool RemoveCar(uint32_t id) {
		auto it = std::find_if(cars_.begin(), cars_.end(), [&](Car* c) { return c->id == id; });
		if (it == cars_.end()) return false;
		if (batch_) throw std::runtime_error("RocketSimB200: RemoveCar after the arena's first Step/GetState/SetState is not supported by the batched back end");
		delete *it; cars_.erase(it);
		return true;
	}
	bool RemoveCar(Car* car) { return RemoveCar(car->id); }
	Car* GetCar(uint32_t id) { for (Car* c : cars_) if (c->id == id) return c; return nullptr; }
	const std::vector<Car*>& GetCars() const { return cars_; }

	const MutatorConfig& GetMutatorConfig() { return mutator_; }
	void SetMutatorConfig(const MutatorConfig& m) {
		if (batch_) throw std::runtime_error("RocketSimB200: SetMutatorConfig after the arena's first Step is not supported by the batched back end");
		mutator_ = m;
	}
	void SetGoalScoreCallback(GoalScoreEventFn fn, void* userInfo = nullptr) { goalFn_ = fn; goalInfo_ = userInfo; }
	void SetCarBumpCallback(CarBumpEventFn fn, void* userInfo = nullptr) { bumpFn_ = fn; bumpInfo_ = userInfo; }
	// GameEventTracker folded into the arena: enable once, then the callbacks fire from Step() like the arena's own
	void EnableEventTracker(const RsbEventTrackerConfig* cfg = nullptr) { Realize(); batch_->SetEventTracker(true, cfg); }
	void SetShotCallback(ShotEventFn fn, void* userInfo = nullptr) { shotFn_ = fn; shotInfo_ = userInfo; }
	void SetGoalCallback(GoalEventFn fn, void* userInfo = nullptr) { goalCreditFn_ = fn; goalCreditInfo_ = userInfo; }
	void SetSaveCallback(SaveEventFn fn, void* userInfo = nullptr) { saveFn_ = fn; saveInfo_ = userInfo; }
	bool IsBallProbablyGoingIn(float maxTime = 2.f, float extraMargin = 0, Team* goalTeamOut = nullptr) {
		Realize(); std::vector<uint8_t> t; bool r = batch_->IsBallProbablyGoingIn(maxTime, extraMargin, &t)[0] != 0;
		if (r && goalTeamOut) *goalTeamOut = (Team)t[0];
		return r;
	}

	// Arena::Step (Arena.cpp:689-756): uploads every car's `controls`, advances n ticks in one launch, dispatches callbacks.
	void Step(int ticksToSimulate = 1) {
		Realize();
		std::vector<CarControls> c; for (Car* car : cars_) c.push_back(car->controls);
		if (!c.empty()) batch_->SetControls(c.data());
		batch_->Step(ticksToSimulate);
		tickCount += (uint64_t)ticksToSimulate;
		for (const RsbEvent& e : batch_->GetEvents(4096)) {
			if (e.type == RSB_EVENT_GOAL) { if (goalFn_) goalFn_(this, (Team)e.a, goalInfo_); }
			else if (e.type == RSB_EVENT_SHOT) { if (shotFn_) shotFn_(this, CarOfSlotId(e.a), CarOfSlotId(e.b), shotInfo_); }
			else if (e.type == RSB_EVENT_GOAL_CREDIT) { if (goalCreditFn_) goalCreditFn_(this, CarOfSlotId(e.a), CarOfSlotId(e.b), goalCreditInfo_); }
			else if (e.type == RSB_EVENT_SAVE) { if (saveFn_) saveFn_(this, CarOfSlotId(e.a), saveInfo_); }
			else if (bumpFn_) bumpFn_(this, CarOfSlotId(e.a), CarOfSlotId(e.b), e.type == RSB_EVENT_DEMO, bumpInfo_);
		}
	}
	void ResetToRandomKickoff(int seed = -1) { Realize(); batch_->ResetToRandomKickoff({(int32_t)seed}); }
	bool IsBallScored() { Realize(); return batch_->IsBallScored()[0] != 0; }
	// Arena::GetDropshotTilesState / SetDropshotTilesState (Arena.h:178-179, Arena.cpp:225-242); Dropshot arenas only
	DropshotTilesState GetDropshotTilesState() { Realize(); return batch_->GetDropshotTilesStates()[0]; }
	void SetDropshotTilesState(const DropshotTilesState& tilesState) { Realize(); batch_->SetDropshotTilesState(0, tilesState); }
	ArenaBatch& Batch() { Realize(); return *batch_; }

private:
	friend class Car;
	friend class Ball;
	Arena(GameMode mode, float tickRate, int device, uint32_t seed)
		: gameMode(mode), ball(new Ball()), tickTime(1 / tickRate), mutator_(DefaultMutatorConfig(mode)), device_(device), seed_(seed) { ball->arena_ = this; }
	void Realize() {
		if (batch_) return;
		std::vector<Team> teams; std::vector<CarConfig> cfgs;
		for (Car* c : cars_) { teams.push_back(c->team); cfgs.push_back(c->config); }
		batch_.reset(new ArenaBatch(gameMode, 1, teams, cfgs, &mutator_, 1 / tickTime, seed_, device_, config_));
		for (size_t i = 0; i < cars_.size(); i++) cars_[i]->slot_ = (uint32_t)i;  // device car slot = position in AddCar order
	}
	Car* CarOfSlotId(uint32_t slotId) { return (slotId >= 1 && slotId <= cars_.size()) ? cars_[slotId - 1] : nullptr; }  // event ids are slot + 1
	ArenaConfig config_;
	uint32_t lastCarID_ = 0;
	std::vector<Car*> cars_;
	std::unique_ptr<ArenaBatch> batch_;
	MutatorConfig mutator_;
	GoalScoreEventFn goalFn_; void* goalInfo_ = nullptr;
	CarBumpEventFn bumpFn_; void* bumpInfo_ = nullptr;
	ShotEventFn shotFn_; void* shotInfo_ = nullptr;
	GoalEventFn goalCreditFn_; void* goalCreditInfo_ = nullptr;
	SaveEventFn saveFn_; void* saveInfo_ = nullptr;
	int device_; uint32_t seed_;
};

inline CarState Car::GetState() { ArenaBatch& b = arena_->Batch(); return b.GetCarStates()[slot_]; }
inline void Car::SetState(const CarState& s) { ArenaBatch& b = arena_->Batch(); b.SetCarState(0, slot_, s); }
inline BallState Ball::GetState() { return arena_->Batch().GetBallStates()[0]; }
inline void Ball::SetState(const BallState& s) { arena_->Batch().SetBallState(0, s); }

// RocketSim::Init (src/RocketSim.cpp:55-90): reads every *.cmf listed in <folder>/<mode>/index.txt-less directory scan is not
// portable in a header; callers pass the file lists explicitly.
inline std::vector<uint8_t> ReadFile(const std::string& path) {
	std::ifstream f(path, std::ios::binary);
	if (!f) throw std::runtime_error("RocketSimB200: cannot read " + path);
	return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
inline void Init(const std::map<GameMode, std::vector<std::string>>& meshPaths, bool silent = false) {
	std::map<GameMode, std::vector<std::vector<uint8_t>>> files;
	for (auto& kv : meshPaths) for (auto& p : kv.second) files[kv.first].push_back(ReadFile(p));
	InitFromMem(files, silent);
}
// RocketSim::Init(collisionMeshesFolder) (src/RocketSim.h:27, RocketSim.cpp:55-90): every *.cmf under <folder>/soccar, /hoops,
// /dropshot (the game modes with meshes of their own); missing sub-folders are skipped, as in the reference.
inline void Init(const std::filesystem::path& collisionMeshesFolder, bool silent = false) {
	std::map<GameMode, std::vector<std::vector<uint8_t>>> files;
	for (GameMode gm : {GameMode::SOCCAR, GameMode::HOOPS, GameMode::DROPSHOT}) {
		const std::filesystem::path dir = collisionMeshesFolder / GAMEMODE_STRS[(int)gm];
		if (!std::filesystem::exists(dir)) continue;
		std::vector<std::filesystem::path> paths;
		for (auto& entry : std::filesystem::directory_iterator(dir))
			if (entry.path().has_extension() && entry.path().extension() == ".cmf") paths.push_back(entry.path());
		std::sort(paths.begin(), paths.end());  // directory order is unspecified; mesh order fixes the static-object order
		for (auto& p : paths) files[gm].push_back(ReadFile(p.string()));
	}
	InitFromMem(files, silent);
}

}  // namespace RocketSimB200
