/* oracle/ref_harness.h -- C interface of the ORACLE (test infrastructure only).
 *
 * A thin extern "C" shim over the UNMODIFIED reference library (RocketSim 2.2 compiled
 * from /root/reference by oracle/Makefile into oracle/_ref/librocketsim_ref.so).  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product (rocketsim_b200/csrc) never links or calls anything in here.
 *
 * All state crosses this interface in the layouts of include/rsb_types.h, which are
 * byte-compatible with the reference structs (static_asserts in ref_harness.cpp).
 */
#ifndef REF_HARNESS_H
#define REF_HARNESS_H

#include <stddef.h>
#include <stdint.h>
#include "../include/rsb_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct RefArena RefArena;

/* RocketSim::InitFromMem (src/RocketSim.cpp:92-185).  mesh i belongs to game mode
 * mesh_modes[i]; bytes are .cmf files (src/CollisionMeshFile/CollisionMeshFile.cpp:11-61).
 * n_meshes == 0 initialises with no meshes (THE_VOID only).  Idempotent. Returns 0 on success. */
int ref_init_from_mem(int n_meshes, const int* mesh_modes, const uint8_t* const* mesh_ptrs, const size_t* mesh_sizes);
const char* ref_last_error(void);

RefArena* ref_arena_create(int game_mode, float tick_rate);
RefArena* ref_arena_create_cfg(int game_mode, float tick_rate, int no_ball_rot, int n_pads, const float* pads /* n_pads x (x, y, z, isBig) */);          /* Arena::Create, Arena.cpp:542 */
void ref_arena_destroy(RefArena*);
uint32_t ref_add_car(RefArena*, int team, int car_preset);           /* Arena::AddCar, Arena.cpp:50 */
int ref_num_cars(RefArena*);
int ref_num_pads(RefArena*);
void ref_seed_thread_rng(uint32_t seed);                             /* Math::GetRandEngine().seed, Math.cpp:59-64 */
void ref_reset_kickoff(RefArena*, int seed);                         /* Arena::ResetToRandomKickoff, Arena.cpp:113 */
void ref_set_controls(RefArena*, uint32_t car_id, const RsbCarControls*);
void ref_step(RefArena*, int n_ticks);                               /* Arena::Step, Arena.cpp:689 */
uint64_t ref_tick_count(RefArena*);
int ref_is_ball_scored(RefArena*);
void ref_enable_event_tracker(RefArena*, const float* cfg7);              /* GameEventTracker, GameEventTracker.h:45-100 */
size_t ref_serialize(RefArena*, uint8_t* out, size_t cap);               /* Arena::Serialize, Arena.cpp:546 */
RefArena* ref_deserialize_new(const uint8_t* data, size_t n);           /* Arena::DeserializeNew, Arena.cpp:576 */
RefArena* ref_clone(RefArena*);                                         /* Arena::Clone, Arena.cpp:646 */
int ref_ball_pred(RefArena*, int n_ticks, RsbBallState* out);            /* BallPredTracker, BallPredTracker.cpp:5-66 */
int ref_is_ball_probably_going_in(RefArena*, float maxTime, float extraMargin, int* teamOut); /* Arena.cpp:771-891 */                                   /* Arena::IsBallScored, Arena.cpp:893 */

void ref_get_car_state(RefArena*, uint32_t car_id, RsbCarState* out);  /* Car::GetState, Car.cpp:10 */
void ref_set_car_state(RefArena*, uint32_t car_id, const RsbCarState* in);
void ref_get_car_internal(RefArena*, uint32_t car_id, RsbCarInternal* out);
void ref_set_car_internal(RefArena*, uint32_t car_id, const RsbCarInternal* in);
void ref_get_ball_state(RefArena*, RsbBallState* out);                 /* Ball::GetState, Ball.cpp:27 */
void ref_set_ball_state(RefArena*, const RsbBallState* in);
void ref_get_ball_internal(RefArena*, RsbBallInternal* out);
void ref_set_ball_internal(RefArena*, const RsbBallInternal* in);
void ref_get_pad_states(RefArena*, RsbBoostPadState* out);             /* n = ref_num_pads */
void ref_set_pad_states(RefArena*, const RsbBoostPadState* in);
void ref_get_pad_positions(RefArena*, float* out_xyz_isbig);           /* 4 floats per pad */
void ref_get_dropshot_tiles(RefArena*, uint8_t* out140);               /* Arena::GetDropshotTilesState, Arena.h:178 */
void ref_set_dropshot_tiles(RefArena*, const uint8_t* in140);          /* Arena::SetDropshotTilesState, Arena.cpp:225 */
void ref_get_dropshot_tile_pos(int team, int idx, float* xyz);         /* DropshotTiles::GetTilePos, DropshotTiles.cpp:13 */
void ref_get_mutator_config(RefArena*, RsbMutatorConfig* out);
void ref_set_mutator_config(RefArena*, const RsbMutatorConfig* in);    /* Arena::SetMutatorConfig, Arena.cpp:17 */
void ref_default_mutator_config(int game_mode, RsbMutatorConfig* out); /* MutatorConfig(GameMode) */
void ref_get_car_config(int car_preset, RsbCarConfig* out);

/* Events recorded by the goal/bump callbacks since the last call; returns count written. */
int ref_drain_events(RefArena*, RsbEvent* out, int cap);

/* ---- static scene dumps (for the BVH builder / scene unit tests) ---- */
int ref_num_meshes(int game_mode);
/* counts: [0]=numNodes [1]=numSubtreeHeaders [2]=numTris [3]=numVerts [4]=numTriInfo */
int ref_bvh_counts(int game_mode, int mesh, int* counts5);
/* nodes: numNodes x {u16 min[3], u16 max[3], i32 escapeOrTri} (16 B, btQuantizedBvhNode);
 * headers: numHeaders x {u16 min[3], u16 max[3], i32 root, i32 size} (20 B packed);
 * quant: bvhAabbMin[3], bvhAabbMax[3], bvhQuantization[3] (9 floats);
 * triinfo: numTris x {i32 flags, f32 edgeV0V1, edgeV1V2, edgeV2V0} (flags=-1 when absent) */
int ref_bvh_dump(int game_mode, int mesh, void* nodes, void* headers, float* quant9, void* triinfo);

/* ---- debug probes used to localise divergence ---- */
/* wheel ray result of car after last step: per wheel {isInContact, isInContactWithWorld, suspLen,
 * contactPoint xyz, contactNormal xyz, suspRelVel, clippedInv, impulse xyz} = 14 floats */
void ref_get_wheel_debug(RefArena*, uint32_t car_id, float* out4x14);
/* rigid-body transform of car: origin(3) + basis row-major(9) + linvel(3) + angvel(3), Bullet units */
void ref_get_car_body(RefArena*, uint32_t car_id, float* out18);

/* contact points alive after the last step, 16 floats each (see ref_harness.cpp); returns the count */
int ref_get_manifolds(RefArena*, float* out, int cap_points);

/* solver rows of the last step (which: 0 contact, 1 friction), 24 floats each (see ref_harness.cpp) */
void ref_get_solver_rows(RefArena*, int which, int n, float* out);

void ref_get_solver_bodies(RefArena*, int n, float* out); /* 12 floats each: linVel, angVel, extForceImpulse, extTorqueImpulse */

/* ---- CPU throughput harness: T threads x A arenas, README input schedule ---- */
/* returns arena-ticks/s; writes wall seconds to *wall_s.  mesh data must be initialised first. */
double ref_run_bench_inputs(RefArena*, int ticks, uint64_t seed, int gid, int tick0);  /* Arena::Step x ticks with the pinned bench input stream, in C */
double ref_bench(int game_mode, int cars_per_team, int n_threads, int arenas_per_thread,
                 int ticks, uint64_t seed, double* wall_s);

#ifdef __cplusplus
}
#endif
#endif
