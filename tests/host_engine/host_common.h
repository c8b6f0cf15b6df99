// host_common.h -- TEST INFRASTRUCTURE shared by host_engine.cpp and host_mcts.cpp: conversion
// between the public host view of one env (pb_env_state, include/pacbot_b200.h) and the register
// form the device code works on (EnvR + pellet words in observation order), written the way
// pb_get_state / pb_set_state do it (pacbot-rl_b200/csrc/pb_kernels.cu).
#pragma once
#include <cstring>

#include "../../include/pacbot_b200.h"
#include "pb_engine.cuh"

namespace host_common {
using namespace pb;

inline void rows_to_flat(const uint32_t *rows, uint32_t *f) {
    for (int w = 0; w < PEL_WORDS; w++) f[w] = 0;
    for (int r = 0; r < ROWS; r++)
        for (int c = 0; c < COLS; c++)
            if ((rows[r] >> c) & 1u) {
                const int i = fbit(r, c);
                f[i >> 5] |= 1u << (i & 31);
            }
}
inline void flat_to_rows(const uint32_t *f, uint32_t *rows) {
    for (int r = 0; r < PB_ROWS; r++) rows[r] = 0;
    for (int r = 0; r < ROWS; r++)
        for (int c = 0; c < COLS; c++) {
            const int i = fbit(r, c);
            if ((f[i >> 5] >> (i & 31)) & 1u) rows[r] |= 1u << c;
        }
}

inline void from_host(const pb_env_state &o, EnvR &s, uint32_t *f) {
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

inline void to_host(const EnvR &s_in, const uint32_t *f, pb_env_state &o) {
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

}  // namespace host_common
