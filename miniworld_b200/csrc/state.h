// state.h -- the flat structure-of-arrays holding N independent environments in HBM.
//
// Layout rule: every per-env scalar is an array [N] (env index fastest) so that the
// one-thread-per-env physics kernel reads and writes fully coalesced; per-entity fields are
// [max_ents][N].  Static room geometry is either one shared template (all envs of a level
// without per-episode topology) or [N][capacity] blocks (Maze).  Nothing here is ever
// re-laid-out between kernels: the physics kernel and the rasteriser read the same arrays.
#pragma once
#include "../../include/mwb.h"
#include "np_rng.cuh"

#define MWB_MAX_BINS 640         // half-tiles of one entity's screen box that can be binned (160x120 frame: 600)
#define MWB_BIN_REFS 6           // bin references per listed triangle the index buffer has room for

struct TriRec;
struct MeshSegInfo;
struct MazeDev;

struct DevState {
  int32_t N, E, R, Q, S;        // envs, entity slots, room / quad / segment capacity
  int32_t shared_geom;
  int32_t obs_w, obs_h, msaa;

  // ---- dynamic per-env state ----
  int32_t* ent_proto;           // [E][N]  -1 = empty / removed
  double* ent_px;               // [E][N]
  double* ent_py;
  double* ent_pz;
  double* ent_dir;
  double* ent_col;              // [E][3][N]  Box colour after randomize
  double* ent_size;             // [E][N]  per-episode Box edge length drawn by the level (PutNext); 0 = the prototype's
  int32_t* num_slots;           // [N] entity-list length
  int32_t* agent_slot;          // [N]
  int32_t* carrying;            // [N] slot or -1
  int32_t* step_count;          // [N]
  int32_t* num_picked;          // [N]
  int32_t* needs_reset;         // [N] set by a terminated|truncated step when autoreset
  unsigned long long* episodes_done;   // [1] device counter of episode-ending steps
  int32_t* fault;               // [1] capacity faults: K2 triangle lists that overflowed, device world generation that ran
                                //     out of room / quad / segment capacity (mwb_overflow_count; must stay 0)
  double* cam;                  // [4][N]  cam_height, cam_fwd_disp, cam_pitch, cam_fov_y
  double* envp;                 // [12][N] sky_color, light_pos, light_color, light_ambient
  // object removed by the level rule AFTER this step's observation (pickupobjects.py:86-90)
  int32_t* ghost_slot;          // [N] -1 = none
  int32_t* ghost_proto;         // [N]
  double* ghost_pose;           // [4][N] x, y, z, dir
  double* ghost_col;            // [3][N]
  // numpy PCG64 stream
  uint64_t* rng_s_hi;
  uint64_t* rng_s_lo;
  uint64_t* rng_inc_hi;
  uint64_t* rng_inc_lo;
  int32_t* rng_has32;
  uint32_t* rng_cache;

  // ---- geometry: [1 or N][capacity] ----
  int32_t* num_rooms;           // [1 or N]
  int32_t* num_quads;
  int32_t* num_segs;
  mwb_room* rooms;
  mwb_quad* quads;
  mwb_seg* segs;
  int32_t* room_tex;            // [N][R][3] texture id in use (domain-rand variants)

  // ---- per-frame mesh triangle lists written by mesh_setup_kernel: [N][E][mesh_cap] ----
  TriRec* mesh_tris;
  MeshSegInfo* mesh_seg;        // [N][E]
  uint2* mesh_bbox;             // [N][E][mesh_cap] packed (bx, by) of each listed triangle, coalesced for the tile scan
  // the same lists binned by half-tile of the entity's screen box (mesh_setup_kernel): bin b holds
  // mesh_bin_idx[mesh_bin_off[b] .. mesh_bin_off[b + 1]) = triangles that can touch that half-tile
  uint16_t* mesh_bin_idx;       // [N][E][MWB_BIN_REFS * mesh_cap]
  int32_t* mesh_bin_off;        // [N][E][MWB_MAX_BINS + 1]
  int32_t mesh_cap;             // 0 = the level has no mesh entities
  // per-frame trigonometry, written by frame_trig_kernel right before every render launch: the glibc-exact
  // cos / sin of the camera's three angles and of every entity slot's model rotation (one thread each), so that
  // neither K2's nor mesh_setup_kernel's blocks wait for a thread that evaluates them
  double* cam_trig;             // [6][N]  cos, sin of heading, pitch, half field of view
  float* ent_cs;                // [E][2][N]  cos, sin of the slot's glRotatef angle (the form its prototype's render() uses)
  const float* depth_lut;       // [65536] depth16 code -> metres (depth_code_to_metres of every code), or null
  TriRec* room_tris;            // [N][tri_cap] room + box triangle lists in HBM for levels whose lists do
                                //   not fit shared memory (Maze); null = lists live in shared memory

  // ---- level definition (shared) ----
  const MazeDev* maze;          // Maze templates (mwb_set_maze) or null
  const double* maze_cdf;
  const mwb_proto* protos;
  int32_t num_protos;
  const mwb_op* ops;
  int32_t num_ops;
  mwb_params params;
  int32_t rule_kind, rule_arg;
  int32_t domain_rand;
  int32_t max_episode_steps;
  int32_t autoreset;
  double near_extra;            // 1.1 * max_forward_step
  // StochasticActionWrapper on the device (reference wrappers.py:49-71): per step one uniform() draw from
  // the env's own stream; below act_prob the chosen action stands, else act_random (< 0: integers(0, 6))
  int32_t act_noise, act_random;
  double act_prob;
};

MWB_DEV int geom_index(const DevState& S, int i) { return S.shared_geom ? 0 : i; }

MWB_DEV NpRng load_rng(const DevState& S, int i) {
  NpRng r;
  r.s_hi = S.rng_s_hi[i];
  r.s_lo = S.rng_s_lo[i];
  r.inc_hi = S.rng_inc_hi[i];
  r.inc_lo = S.rng_inc_lo[i];
  r.has32 = S.rng_has32[i];
  r.cache = S.rng_cache[i];
  return r;
}

MWB_DEV void store_rng(const DevState& S, int i, const NpRng& r) {
  S.rng_s_hi[i] = r.s_hi;
  S.rng_s_lo[i] = r.s_lo;
  S.rng_has32[i] = r.has32;
  S.rng_cache[i] = r.cache;
}
