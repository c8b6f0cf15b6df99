// pb_state.cuh -- structure-of-arrays layout of the per-env state in HBM, and its register form.
//
// One env = 4 x 16 B packed words + 28 x 4 B pellet words = 176 B, stored field-major so that a
// warp stepping 32 consecutive envs issues fully coalesced 128-bit loads/stores:
//
//   core0[i] (uint4)  x: curr_ticks
//                     y: update_period | mode << 8 | mode_steps << 16 | flags << 24
//                        (flags bit0 paused, bit1 pause_on_update)
//                     z: level_steps | curr_score << 16
//                     w: curr_level | curr_lives << 8 | ghost_combo << 16 | fruit_steps << 24
//   core1[i] (uint4)  x: pac_row | pac_col << 8 | pac_dir << 16 | last_action << 24
//                     y: num_pellets | last_score << 16
//                     z: ticks_per_step | cfg_flags << 8 | ghost_flags << 16
//                        (ghost_flags bit g spawning, bit 4+g eaten)
//                     w: rng_ctr
//   ghosts[p * npad + i] (uint4), p = 0 (red, pink), 1 (cyan, orange); per ghost 8 B:
//                     lo: row | col << 8 | dir << 16 | fright_steps << 24        (loc)
//                     hi: row | col << 8 | dir << 16 | trapped_steps << 24       (next_loc)
//   pel[w * npad + i] (uint32), w < 28: pellet bits in observation order (pb_tables.cuh fbit)
//
// Reference fields: PacmanGym (env.rs:96-110) + the GameState fields of SURVEY.md Appendix A.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "pb_tables.cuh"

namespace pb {

constexpr int STATE_BYTES_PER_ENV = 4 * 16 + PEL_WORDS * 4;  // 176

struct StateView {
    uint4 *core0;
    uint4 *core1;
    uint4 *ghosts;   // [2][npad]
    uint32_t *pel;   // [28][npad]
    long long npad;  // env count rounded up to 32
    long long n;
};

struct GhostR {
    int row, col, dir, fright;
    int nrow, ncol, ndir, trapped;
};

// flags bits
constexpr int F_PAUSED = 1, F_PAUSE_ON_UPDATE = 2;

struct EnvR {
    uint32_t ticks;
    int period, mode, mode_steps, flags;
    int level_steps, score;
    int level, lives, combo, fruit_steps;
    int prow, pcol, pdir, last_action;
    int num_pellets, last_score;
    int tps, cfg, gflags;
    uint32_t rng_ctr;
    GhostR g[4];
};

__host__ __device__ inline int sx8(uint32_t v) { return (int)(int8_t)(v & 0xffu); }

__host__ __device__ inline void unpack_ghost(uint32_t lo, uint32_t hi, GhostR &g) {
    g.row = sx8(lo);
    g.col = sx8(lo >> 8);
    g.dir = (lo >> 16) & 0xff;
    g.fright = (lo >> 24) & 0xff;
    g.nrow = sx8(hi);
    g.ncol = sx8(hi >> 8);
    g.ndir = (hi >> 16) & 0xff;
    g.trapped = (hi >> 24) & 0xff;
}
__host__ __device__ inline void pack_ghost(const GhostR &g, uint32_t &lo, uint32_t &hi) {
    lo = ((uint32_t)g.row & 0xffu) | (((uint32_t)g.col & 0xffu) << 8) |
         (((uint32_t)g.dir & 0xffu) << 16) | (((uint32_t)g.fright & 0xffu) << 24);
    hi = ((uint32_t)g.nrow & 0xffu) | (((uint32_t)g.ncol & 0xffu) << 8) |
         (((uint32_t)g.ndir & 0xffu) << 16) | (((uint32_t)g.trapped & 0xffu) << 24);
}

__host__ __device__ inline void unpack_env(const uint4 &c0, const uint4 &c1, const uint4 &ga,
                                           const uint4 &gb, EnvR &s) {
    s.ticks = c0.x;
    s.period = c0.y & 0xff;
    s.mode = (c0.y >> 8) & 0xff;
    s.mode_steps = (c0.y >> 16) & 0xff;
    s.flags = (c0.y >> 24) & 0xff;
    s.level_steps = c0.z & 0xffff;
    s.score = (c0.z >> 16) & 0xffff;
    s.level = c0.w & 0xff;
    s.lives = (c0.w >> 8) & 0xff;
    s.combo = (c0.w >> 16) & 0xff;
    s.fruit_steps = (c0.w >> 24) & 0xff;
    s.prow = sx8(c1.x);
    s.pcol = sx8(c1.x >> 8);
    s.pdir = (c1.x >> 16) & 0xff;
    s.last_action = (c1.x >> 24) & 0xff;
    s.num_pellets = c1.y & 0xffff;
    s.last_score = (c1.y >> 16) & 0xffff;
    s.tps = c1.z & 0xff;
    s.cfg = (c1.z >> 8) & 0xff;
    s.gflags = (c1.z >> 16) & 0xff;
    s.rng_ctr = c1.w;
    unpack_ghost(ga.x, ga.y, s.g[0]);
    unpack_ghost(ga.z, ga.w, s.g[1]);
    unpack_ghost(gb.x, gb.y, s.g[2]);
    unpack_ghost(gb.z, gb.w, s.g[3]);
}

__host__ __device__ inline void pack_env(const EnvR &s, uint4 &c0, uint4 &c1, uint4 &ga,
                                         uint4 &gb) {
    c0.x = s.ticks;
    c0.y = ((uint32_t)s.period & 0xffu) | (((uint32_t)s.mode & 0xffu) << 8) |
           (((uint32_t)s.mode_steps & 0xffu) << 16) | (((uint32_t)s.flags & 0xffu) << 24);
    c0.z = ((uint32_t)s.level_steps & 0xffffu) | (((uint32_t)s.score & 0xffffu) << 16);
    c0.w = ((uint32_t)s.level & 0xffu) | (((uint32_t)s.lives & 0xffu) << 8) |
           (((uint32_t)s.combo & 0xffu) << 16) | (((uint32_t)s.fruit_steps & 0xffu) << 24);
    c1.x = ((uint32_t)s.prow & 0xffu) | (((uint32_t)s.pcol & 0xffu) << 8) |
           (((uint32_t)s.pdir & 0xffu) << 16) | (((uint32_t)s.last_action & 0xffu) << 24);
    c1.y = ((uint32_t)s.num_pellets & 0xffffu) | (((uint32_t)s.last_score & 0xffffu) << 16);
    c1.z = ((uint32_t)s.tps & 0xffu) | (((uint32_t)s.cfg & 0xffu) << 8) |
           (((uint32_t)s.gflags & 0xffu) << 16);
    c1.w = s.rng_ctr;
    pack_ghost(s.g[0], ga.x, ga.y);
    pack_ghost(s.g[1], ga.z, ga.w);
    pack_ghost(s.g[2], gb.x, gb.y);
    pack_ghost(s.g[3], gb.z, gb.w);
}

// ---- counter-based RNG (documented in include/pacbot_b200.h) ---------------------------------
__host__ __device__ inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
__host__ __device__ inline uint64_t env_key(uint64_t seed, uint64_t gid) {
    return mix64(seed + 0x9e3779b97f4a7c15ull * (gid + 1ull));
}
// game-draw stream: frightened-ghost turns and reset randomisation
__host__ __device__ inline uint32_t draw_raw(uint64_t seed, uint64_t gid, uint32_t ctr) {
    return (uint32_t)(mix64(env_key(seed, gid) + (uint64_t)ctr * 0x9e3779b97f4a7c15ull +
                            0x632be59bd9b4e019ull) >> 32);
}
// policy stream: random actions / epsilon tests, keyed by a caller-supplied call index
__host__ __device__ inline uint32_t action_raw(uint64_t seed, uint64_t gid, uint64_t call) {
    return (uint32_t)(mix64(env_key(seed, gid) + call * 0xd1342543de82ef95ull +
                            0xa0761d6478bd642full) >> 32);
}
// replay-sampling stream (pb_replay_sample): draw j of sampling call `call`
__host__ __device__ inline uint32_t replay_raw(uint64_t seed, uint64_t call, uint32_t j) {
    return (uint32_t)(mix64(mix64(seed + 0x9e3779b97f4a7c15ull * (call + 1ull)) +
                            (uint64_t)j * 0xd1342543de82ef95ull + 0x8ebc6af09c88c6e3ull) >> 32);
}
// seed of the game-draw stream used by MCTS SIMULATIONS (pb_mcts.cuh): the cloned env of search
// iteration `epoch` keeps its gid and rng_ctr but draws from a different seed, so a simulation
// never sees the draws the real env is going to get
__host__ __device__ inline uint64_t sim_seed(uint64_t seed, uint64_t epoch) {
    return mix64(seed ^ ((epoch + 1ull) * 0xd6e8feb86659fd93ull));
}
__host__ __device__ inline uint32_t below(uint32_t raw, uint32_t n) {
    return (uint32_t)(((uint64_t)raw * (uint64_t)n) >> 32);
}

}  // namespace pb
