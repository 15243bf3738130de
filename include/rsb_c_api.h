/* rsb_c_api.h -- C ABI of the B200 batched Arena::Step.
 *
 * The reference has no FFI layer: its boundary is the public C++ surface of
 * RocketSim::Arena / Car / Ball (reference src/Sim/Arena/Arena.h, src/Sim/Car/Car.h,
 * src/Sim/Ball/Ball.h).  These entry points are what a binding of that
 * surface calls instead of the CPU implementation; every function cites the reference
 * interface it replaces.  Plain pointers and sizes only; all POD layouts are those of
 * include/rsb_types.h (byte-compatible with the reference structs).
 *
 * All functions return 0 on success, a negative value on failure (rsb_last_error() then
 * describes it).  Nothing throws across this boundary.  A batch is not re-entrant: one host
 * thread drives one batch (the reference's "one thread per Arena" rule, README.md:11).
 * There is NO CPU fallback: without a CUDA device every batch call fails.
 */
#ifndef RSB_C_API_H
#define RSB_C_API_H

#include <stddef.h>
#include <stdint.h>
#include "rsb_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct RsbBatch RsbBatch;

const char* rsb_last_error(void);
/* ABI self-description (used by tests/test_abi.py): sizeof of the named boundary struct, or -1 */
int rsb_sizeof(const char* type_name);

/* Replaces RocketSim::InitFromMem (src/RocketSim.h:32, src/RocketSim.cpp:92-185): registers the
 * .cmf collision meshes of each game mode (mesh i belongs to mesh_modes[i]) and builds the
 * static scene (quantised BVH, internal-edge info, broadphase cell lists) once per mode.
 * n_meshes == 0 is valid (THE_VOID only).  Idempotent like the reference. */
int rsb_init_from_mem(int n_meshes, const int* mesh_modes, const uint8_t* const* mesh_ptrs, const size_t* mesh_sizes);

/* The reference's constant tables: CAR_CONFIG_OCTANE ... CAR_CONFIG_PSYCLOPS (src/Sim/Car/CarConfig/CarConfig.cpp:20-100,
 * preset = RSB_CAR_*) and MutatorConfig(GameMode) (src/Sim/MutatorConfig/MutatorConfig.cpp:5-40). Host only, no device needed. */
int rsb_car_config_preset(int preset, RsbCarConfig* out);
int rsb_default_mutator_config(int game_mode, RsbMutatorConfig* out);

/* Replaces Arena::Create + n_cars x Arena::AddCar (src/Sim/Arena/Arena.h:111,71;
 * Arena.cpp:416-544, 50-62) for n_arenas identical arenas on CUDA device `device`.
 * teams[i] / car_configs[i] describe car slot i (car id = i + 1, the reference's ++_lastCarID).
 * mutator == NULL selects MutatorConfig(game_mode).  Cars start at the reference's
 * CarState() defaults, the ball at BallState(); `seed` seeds the per-arena respawn RNG
 * (arena a uses seed + a), the reference's thread-local engine (src/Math/Math.cpp:59-64). */
int rsb_batch_create(int device, int game_mode, int n_arenas, int n_cars, const int* teams,
                     const RsbCarConfig* car_configs, const RsbMutatorConfig* mutator, float tick_rate,
                     uint32_t seed, RsbBatch** out);
/* Arena::Create(gameMode, arenaConfig, tickRate) (src/Sim/Arena/Arena.h:111, Arena.cpp:416-545) for every arena of the batch:
 * rsb_batch_create with an explicit ArenaConfig (NULL = the default one).  Unsupported values fail with an error naming the
 * field (see RsbArenaConfig). */
void rsb_default_arena_config(RsbArenaConfig* out);
int rsb_batch_create_ex(int device, int game_mode, const RsbArenaConfig* arena_config, int n_arenas, int n_cars, const int* teams,
                        const RsbCarConfig* car_configs, const RsbMutatorConfig* mutator, float tick_rate, uint32_t seed, RsbBatch** out);
int rsb_batch_destroy(RsbBatch* b);
int rsb_batch_info(RsbBatch* b, int* n_arenas, int* n_cars, int* n_pads, int* game_mode);
/* CUDA stream (cudaStream_t) on which all batch work is enqueued; 0 = legacy default stream */
int rsb_set_stream(RsbBatch* b, void* cuda_stream);

/* Replaces `car->controls = ...` (src/Sim/Car/Car.h:149) for every car: controls is
 * [n_arenas][n_cars] in HOST memory. */
int rsb_set_controls(RsbBatch* b, const RsbCarControls* controls);
/* Same, from DEVICE memory (no host round trip). */
int rsb_set_controls_device(RsbBatch* b, const void* d_controls);

/* Replaces Arena::Step(ticksToSimulate) (src/Sim/Arena/Arena.h:132, Arena.cpp:689-756) for all
 * arenas of the batch: one fused kernel launch advancing n_ticks ticks. Asynchronous. */
int rsb_step(RsbBatch* b, int n_ticks);
/* Like rsb_step(1) repeated n_ticks times with a different control set per tick:
 * d_controls_seq is DEVICE memory [n_ticks][n_arenas][n_cars]. One launch. */
int rsb_step_with_controls_device(RsbBatch* b, const void* d_controls_seq, int n_ticks);
/* Same with the control sequence [n_ticks][n_arenas][n_cars] in HOST memory (one H2D copy + one launch):
 * the reference loop `for t: car->controls = seq[t]; arena->Step(1)` (Car.h:149, Arena.h:132). */
int rsb_step_with_controls(RsbBatch* b, const RsbCarControls* controls_seq, int n_ticks);
/* README bench loop on device (README.md:49-53; oracle ref_bench): when enabled, a tick that ends with
 * IsBallScored() is followed by ResetToRandomKickoff(seed + arena + tick) inside the same launch
 * (Arena.cpp:113-223, 744-752). Off by default, as in the reference. */
int rsb_set_auto_reset(RsbBatch* b, int enable, uint32_t seed);
/* Sharded jobs (one batch per GPU): global index of this batch's arena 0.  Auto-reset seeds and the arena ids of events are
 * formed with the GLOBAL arena index, so per-arena results do not depend on how a job is split over GPUs.  Default 0. */
int rsb_set_arena_offset(RsbBatch* b, uint32_t global_index_of_arena0);
int rsb_sync(RsbBatch* b);

/* Replaces GameEventTracker (src/Sim/GameEventTracker/GameEventTracker.h:45-100, .cpp:5-157) used as
 * `arena->Step(1); tracker.Update(arena);`: when enabled, every tick of every arena ends with the tracker's Update and its
 * shot / goal / save callbacks arrive through rsb_get_events as RSB_EVENT_SHOT / _GOAL_CREDIT / _SAVE (a = shooter, scorer or
 * saver car id, b = passer id or 0).  cfg == NULL selects GameEventTrackerConfig's defaults.  Soccar, Snowday and Hoops only
 * (Arena::IsBallProbablyGoingIn raises for the other modes, Arena.cpp:887). */
int rsb_set_event_tracker(RsbBatch* b, int enable, const RsbEventTrackerConfig* cfg);
/* Wire compatibility with the CPU library.  rsb_serialize_arena writes arena `arena_idx` in the byte format of
 * Arena::Serialize (src/Sim/Arena/Arena.cpp:546-574; no version header, like the in-memory DataStreamOut) -- a stream that
 * RocketSim's Arena::DeserializeNew loads; out == NULL only returns the size in *n_out.  rsb_deserialize_arena is
 * Arena::DeserializeNew (Arena.cpp:576-644) into an existing arena of the batch: the stream's game mode, tick rate, cars
 * (ids 1..n_cars, teams, configs) and mutators must equal what the batch was created with. */
int rsb_serialize_arena(RsbBatch* b, uint32_t arena_idx, uint8_t* out, size_t cap, size_t* n_out);
/* Replaces Arena::Clone (Arena.cpp:646-674): arena `dst` becomes a clone of arena `src` of the same batch. */
int rsb_clone_arena(RsbBatch* b, uint32_t src, uint32_t dst);
int rsb_deserialize_arena(RsbBatch* b, uint32_t arena_idx, const uint8_t* data, size_t n);

/* Replaces BallPredTracker::ForceUpdateAllPred / predData (src/Sim/BallPredTracker/BallPredTracker.cpp:58-66) for every arena
 * of the batch at once: out is [n_ticks][n_arenas]; row 0 = the arena's current Ball::GetState(), row i = the ball state after i
 * ticks of a ball-only arena of the same game mode, mutators and tick rate started from it (cars are not simulated). */
int rsb_predict_ball(RsbBatch* b, int n_ticks, RsbBallState* out);

/* Replaces Arena::IsBallProbablyGoingIn(maxTime, extraMargin, &goalTeam) (Arena.h, Arena.cpp:771-891); out / goal_team_out are
 * [n_arenas] (goal_team_out may be NULL; it is only meaningful where out is 1). */
int rsb_is_ball_probably_going_in(RsbBatch* b, float max_time, float extra_margin, uint8_t* out, uint8_t* goal_team_out);

/* Replace Car::GetState / Car::SetState (src/Sim/Car/Car.h:151-152, Car.cpp:10-36).
 * get: out is [n_arenas][n_cars] host memory.  set: n entries addressed by (arena_idx[i], car_idx[i]),
 * car_idx is the 0-based slot (id - 1). */
int rsb_get_car_states(RsbBatch* b, RsbCarState* out);
int rsb_set_car_states(RsbBatch* b, const uint32_t* arena_idx, const uint32_t* car_idx, int n, const RsbCarState* states);
/* Replace Ball::GetState / SetState (src/Sim/Ball/Ball.h:68-69, Ball.cpp:27-51); out is [n_arenas]. */
int rsb_get_ball_states(RsbBatch* b, RsbBallState* out);
int rsb_set_ball_states(RsbBatch* b, const uint32_t* arena_idx, int n, const RsbBallState* states);
/* Replace BoostPad::GetState / SetState (src/Sim/BoostPad/BoostPad.h:44-45); [n_arenas][n_pads]. */
int rsb_get_pad_states(RsbBatch* b, RsbBoostPadState* out);
int rsb_set_pad_states(RsbBatch* b, const uint32_t* arena_idx, int n, const RsbBoostPadState* states /* [n][n_pads] */);
/* Replace Arena::GetDropshotTilesState / SetDropshotTilesState (src/Sim/Arena/Arena.h:178-179, Arena.cpp:225-242) of a Dropshot
 * batch: one DropshotTileState::damageState byte (0 full, 1 damaged, 2 broken) per tile, [n_arenas][2 teams][70 tiles]
 * (src/Sim/Arena/DropshotTiles/DropshotTiles.h:7-33).  Setting also updates the tiles' contact response, as the reference does. */
int rsb_get_dropshot_tiles(RsbBatch* b, uint8_t* out /* [n_arenas][140] */);
int rsb_set_dropshot_tiles(RsbBatch* b, const uint32_t* arena_idx, int n, const uint8_t* states /* [n][140] */);
/* Arena::tickCount (Arena.h:60) and Arena::IsBallScored (Arena.h:144, Arena.cpp:893-927); [n_arenas]. */
int rsb_get_tick_counts(RsbBatch* b, uint64_t* out);
int rsb_is_ball_scored(RsbBatch* b, uint8_t* out);
/* Goal / bump / demo events since the last call (the reference delivers them through
 * SetGoalScoreCallback / SetCarBumpCallback, Arena.h:102,108).  Sorted by (arena, tick).  Nothing is dropped when more than `cap`
 * events are pending: the call hands out the first `cap` and returns 1 (rsb_last_error says how many remain); call again. */
int rsb_get_events(RsbBatch* b, RsbEvent* out, int cap, int* n_out);
/* Device-side error flags per arena (manifold overflow etc.), 0 = clean; [n_arenas]. */
int rsb_get_error_flags(RsbBatch* b, uint32_t* out);
int rsb_seed_rng(RsbBatch* b, uint32_t arena_idx, uint32_t seed);

/* Compact per-tick observation download for training loops: for each arena
 * ball pos(3) vel(3) angvel(3), then per car pos(3) rotMat fwd(3) up(3) vel(3) angvel(3) boost(1)
 * flags(1: bit0 onGround, bit1 demoed, bit2 supersonic, bit3 hasFlipOrJump) as float32
 * -> 9 + 17 * n_cars floats per arena, units as in CarState/BallState. */
int rsb_get_observations(RsbBatch* b, float* out);
int rsb_observation_floats(RsbBatch* b);

/* Exact transfer of the state the reference keeps outside CarState/BallState;
 * used by the parity tests to resynchronise mid-rollout. */
int rsb_get_car_internals(RsbBatch* b, RsbCarInternal* out);
int rsb_set_car_internals(RsbBatch* b, const uint32_t* arena_idx, const uint32_t* car_idx, int n, const RsbCarInternal* in);
int rsb_get_ball_internals(RsbBatch* b, RsbBallInternal* out);
int rsb_set_ball_internals(RsbBatch* b, const uint32_t* arena_idx, int n, const RsbBallInternal* in);

/* Replaces Arena::ResetToRandomKickoff(seed) (Arena.h, Arena.cpp:113-223) for the listed arenas
 * (all arenas when arena_idx == NULL; then seeds[a] applies to arena a). */
int rsb_reset_to_random_kickoff(RsbBatch* b, const uint32_t* arena_idx, int n, const int32_t* seeds);

/* Debug probe: per wheel {isInContact, isInContactWithWorld, suspLen, contactPoint xyz, contactNormal xyz,
 * suspRelVel, clippedInv, impulse xyz} of the last tick for arena `arena`; out is [n_cars*4][14]. */
int rsb_debug_get_wheels(RsbBatch* b, int arena, float* out);

/* Debug probe: contact points of the last tick of arena `arena`, 16 floats each {obj0, obj1, point#, dist, normalOnB xyz,
 * posA xyz, posB xyz, friction, restitution, isSpecial}; obj = world object index (meshes, planes, ball, cars). */
int rsb_debug_get_manifolds(RsbBatch* b, int arena, float* out, int cap_points, int* n_out);

/* Debug probe: solver rows of the last tick (which: 0 contact, 1 friction), 24 floats each {n1, rp1xn, n2, rp2xn, angA, angB,
 * jacDiagInv, rhs, rhsPenetration, appliedImpulse, appliedPushImpulse, friction}. */
int rsb_debug_get_solver_rows(RsbBatch* b, int arena, int which, float* out, int cap, int* n_out);

/* Debug probe: solver bodies of the last tick, 12 floats each {linVel, angVel, externalForceImpulse, externalTorqueImpulse}. */
int rsb_debug_get_solver_bodies(RsbBatch* b, int arena, float* out, int cap, int* n_out);

/* Debug probe: the four per-arena scratch ints of the last tick ([0] solver ran, [1] #manifolds solved, [2..3] developer use). */
int rsb_debug_get_misc(RsbBatch* b, int arena, int* out4);
/* Test hook: usable capacity of the per-block detector queue (1..512).  Results never depend on it (records that do not fit
 * are swept up after the queued ones); tests shrink it to exercise that sweep. */
int rsb_debug_set_detect_queue_cap(RsbBatch* b, int cap);

/* Developer builds (-DRSB_PROFILE) only: per tick phase p: out[p] block wall cycles, out[16+p] busy-thread cycles,
 * out[32+p] busy threads, out[48+p] slowest-thread cycles (each summed over blocks and ticks); zeros otherwise. */
int rsb_debug_profile(uint64_t* out64, int reset);

/* Static-scene introspection for the scene-builder unit tests. */
int rsb_scene_counts(int game_mode, int mesh, int* counts5);
int rsb_scene_dump(int game_mode, int mesh, void* nodes, void* headers, float* quant9, void* triinfo);
int rsb_scene_num_meshes(int game_mode);
/* CSR of the broadphase cell lists: n_cells+1 offsets and the static indices */
int rsb_scene_cells(int game_mode, int32_t* cell_start, int cap_start, uint8_t* statics, int cap_statics, int* n_cells, int* n_entries);

/* Bench helper: fills d_controls_seq [n_ticks][n_arenas][n_cars] (DEVICE memory) with the pinned
 * "README bench" schedule of include/rsb_bench_inputs.h starting at tick `tick0`. */
int rsb_bench_fill_controls_device(RsbBatch* b, void* d_controls_seq, int n_ticks, uint32_t tick0, uint64_t seed);

#ifdef __cplusplus
}
#endif
#endif
