// host_engine.cpp -- TEST INFRASTRUCTURE: the device game logic of the product
// (pacbot-rl_b200/csrc/pb_engine.cuh: gym_step / gym_reset / action_mask_bits, the functions
// k_step and k_reset call per thread) compiled for the HOST with g++ through the shim
// cuda_runtime.h next to this file, so that tests/test_host_engine.py can run long differential
// campaigns of the *same source the GPU executes* against the scalar C oracle in a container
// without a GPU.  It mirrors what k_step / k_reset do around those calls
// (pacbot-rl_b200/csrc/pb_kernels.cu) and converts to/from pb_env_state exactly like
// pb_get_state / pb_set_state.  It is NOT a CPU path of the product: nothing under
// pacbot-rl_b200/ builds, loads or links it.
#include <cstring>

#include "../../include/pacbot_b200.h"
#include "pb_engine.cuh"

using namespace pb;

namespace {

void rows_to_flat(const uint32_t *rows, uint32_t *f) {
    for (int w = 0; w < PEL_WORDS; w++) f[w] = 0;
    for (int r = 0; r < ROWS; r++)
        for (int c = 0; c < COLS; c++)
            if ((rows[r] >> c) & 1u) {
                const int i = fbit(r, c);
                f[i >> 5] |= 1u << (i & 31);
            }
}
void flat_to_rows(const uint32_t *f, uint32_t *rows) {
    for (int r = 0; r < PB_ROWS; r++) rows[r] = 0;
    for (int r = 0; r < ROWS; r++)
        for (int c = 0; c < COLS; c++) {
            const int i = fbit(r, c);
            if ((f[i >> 5] >> (i & 31)) & 1u) rows[r] |= 1u << c;
        }
}

void from_host(const pb_env_state &o, EnvR &s, uint32_t *f) {
    s.ticks = o.curr_ticks;
    s.rng_ctr = o.rng_ctr;
    s.level_steps = o.level_steps;
    s.score = o.curr_score;
    s.num_pellets = o.num_pellets;
    s.last_score = o.last_score;
    s.period = o.update_period;
    s.mode = o.mode;
    s.flags = (o.paused ? F_PAUSED : 0) | (o.pause_on_update ? F_PAUSE_ON_UPDATE : 0);
    s.mode_steps = o.mode_steps;
    s.level = o.curr_level;
    s.lives = o.curr_lives;
    s.combo = o.ghost_combo;
    s.fruit_steps = o.fruit_steps;
    s.last_action = o.last_action;
    s.tps = o.ticks_per_step;
    s.cfg = o.cfg_flags & 0x0f;
    s.prow = o.pac_row;
    s.pcol = o.pac_col;
    s.pdir = o.pac_dir;
    s.gflags = 0;
    for (int k = 0; k < 4; k++) {
        s.g[k].row = o.g_row[k];
        s.g[k].col = o.g_col[k];
        s.g[k].dir = o.g_dir[k];
        s.g[k].nrow = o.g_next_row[k];
        s.g[k].ncol = o.g_next_col[k];
        s.g[k].ndir = o.g_next_dir[k];
        s.g[k].fright = o.g_fright[k];
        s.g[k].trapped = o.g_trapped[k];
        s.gflags |= (o.g_spawning[k] ? 1 : 0) << k;
        s.gflags |= (o.g_eaten[k] ? 1 : 0) << (4 + k);
    }
    rows_to_flat(o.pellets, f);
    // the device keeps the state PACKED between launches: go through pack/unpack so that field
    // widths (u8 / u16 truncation, sign extension of coordinates) are exercised here too
    uint4 c0, c1, ga, gb;
    pack_env(s, c0, c1, ga, gb);
    unpack_env(c0, c1, ga, gb, s);
}

void to_host(const EnvR &s_in, const uint32_t *f, pb_env_state &o) {
    uint4 c0, c1, ga, gb;
    pack_env(s_in, c0, c1, ga, gb);
    EnvR s;
    unpack_env(c0, c1, ga, gb, s);
    memset(&o, 0, sizeof(o));
    o.curr_ticks = s.ticks;
    o.rng_ctr = s.rng_ctr;
    flat_to_rows(f, o.pellets);
    o.level_steps = (uint16_t)s.level_steps;
    o.curr_score = (uint16_t)s.score;
    o.num_pellets = (uint16_t)s.num_pellets;
    o.last_score = (uint16_t)s.last_score;
    o.update_period = (uint8_t)s.period;
    o.mode = (uint8_t)s.mode;
    o.paused = (s.flags & F_PAUSED) ? 1 : 0;
    o.pause_on_update = (s.flags & F_PAUSE_ON_UPDATE) ? 1 : 0;
    o.mode_steps = (uint8_t)s.mode_steps;
    o.curr_level = (uint8_t)s.level;
    o.curr_lives = (uint8_t)s.lives;
    o.ghost_combo = (uint8_t)s.combo;
    o.fruit_steps = (uint8_t)s.fruit_steps;
    o.last_action = (uint8_t)s.last_action;
    o.ticks_per_step = (uint8_t)s.tps;
    o.cfg_flags = (uint8_t)s.cfg;
    o.pac_row = (int8_t)s.prow;
    o.pac_col = (int8_t)s.pcol;
    o.pac_dir = (uint8_t)s.pdir;
    for (int k = 0; k < 4; k++) {
        o.g_row[k] = (int8_t)s.g[k].row;
        o.g_col[k] = (int8_t)s.g[k].col;
        o.g_dir[k] = (uint8_t)s.g[k].dir;
        o.g_next_row[k] = (int8_t)s.g[k].nrow;
        o.g_next_col[k] = (int8_t)s.g[k].ncol;
        o.g_next_dir[k] = (uint8_t)s.g[k].ndir;
        o.g_fright[k] = (uint8_t)s.g[k].fright;
        o.g_trapped[k] = (uint8_t)s.g[k].trapped;
        o.g_spawning[k] = (s.gflags >> k) & 1;
        o.g_eaten[k] = (s.gflags >> (4 + k)) & 1;
    }
}

RngCtx make_rng(uint64_t seed, int64_t gid_base, const uint32_t *tape, int64_t tape_len,
                int64_t i) {
    RngCtx r;
    r.seed = seed;
    r.gid = (uint64_t)(gid_base + i);
    r.tape = tape ? tape + i * tape_len : nullptr;
    r.tape_len = (uint32_t)tape_len;
    return r;
}

void write_mask(uint8_t *out, uint32_t m) {
    for (int k = 0; k < 5; k++) out[k] = (m >> k) & 1u;
}

}  // namespace

extern "C" {

size_t he_sizeof_env_state(void) { return sizeof(pb_env_state); }

// what one k_step thread does (pb_kernels.cu k_step, actions given by the caller)
int64_t he_step(pb_env_state *st, int64_t n, const uint8_t *actions, int32_t *reward,
                uint8_t *done_out, uint8_t *mask_out, int auto_reset, uint64_t seed,
                int64_t gid_base, const uint32_t *tape, int64_t tape_len) {
    int64_t bad = -1;
    for (int64_t i = 0; i < n; i++) {
        EnvR s;
        uint32_t f[PEL_WORDS];
        from_host(st[i], s, f);
        const int action = actions[i];
        if (action > 4) {
            if (bad < 0) bad = i;
            reward[i] = 0;
            done_out[i] = 0;
            if (mask_out) write_mask(mask_out + i * 5, action_mask_bits(s, H_TABLES.walls));
            continue;
        }
        const PelRef pel{f, 1};
        const RngCtx rng = make_rng(seed, gid_base, tape, tape_len, i);
        int rew;
        bool done;
        gym_step(s, pel, H_TABLES, H_TABLES.walls, rng, action, rew, done);
        if (done && auto_reset) gym_reset(s, pel, H_TABLES, H_TABLES.walls, rng, true);
        to_host(s, f, st[i]);
        reward[i] = rew;
        done_out[i] = done ? 1 : 0;
        if (mask_out) write_mask(mask_out + i * 5, action_mask_bits(s, H_TABLES.walls));
    }
    return bad;
}

// what one k_reset thread does; randomise == 0 is the constructor state (pb_create)
void he_reset(pb_env_state *st, int64_t n, const uint8_t *mask, int randomise, uint64_t seed,
              int64_t gid_base, const uint32_t *tape, int64_t tape_len) {
    for (int64_t i = 0; i < n; i++) {
        if (mask && !mask[i]) continue;
        EnvR s;
        uint32_t f[PEL_WORDS];
        from_host(st[i], s, f);
        const PelRef pel{f, 1};
        gym_reset(s, pel, H_TABLES, H_TABLES.walls,
                  make_rng(seed, gid_base, tape, tape_len, i), randomise != 0);
        to_host(s, f, st[i]);
    }
}

void he_action_mask(const pb_env_state *st, int64_t n, uint8_t *out) {
    for (int64_t i = 0; i < n; i++) {
        EnvR s;
        uint32_t f[PEL_WORDS];
        from_host(st[i], s, f);
        write_mask(out + i * 5, action_mask_bits(s, H_TABLES.walls));
    }
}

}  // extern "C"
