// pb_tables.cuh -- maze constants and derived lookup tables (host constexpr + device __constant__).
//
// Source of truth for the geometry:
//   walls   = complement of /root/reference/computed_data/node_coords.json in the 31x28 grid
//             (forced by /root/reference/pacbot_rs/build.rs:78-108), SURVEY.md Appendix D.1
//   pellets = SURVEY.md Appendix D.2 (244 pellets, super pellets at the four cells named in
//             /root/reference/pacbot_rs/src/env.rs:283)
// Everything else in this file is DERIVED from those two bitboards at compile time.
#pragma once
#include <cstdint>

namespace pb {

constexpr int ROWS = 31;
constexpr int COLS = 28;
constexpr int NUM_NODES = 288;
constexpr int CH_FLOATS = COLS * ROWS;         // 868 floats per observation channel
constexpr int OBS_CH = 17;
constexpr int OBS_FLOATS = OBS_CH * CH_FLOATS; // 14 756
constexpr int OBS_BYTES = OBS_FLOATS * 4;      // 59 024 = 3 689 * 16
constexpr int PEL_WORDS = 28;                  // ceil(868 / 32): pellets in observation order

// ---- engine constants (pacbot_rs_2::variables; DESIGN.md "Engine assumption ledger") ---------
constexpr int INIT_UPDATE_PERIOD = 12;
constexpr int INIT_LEVEL = 1;
constexpr int INIT_LIVES = 3;
constexpr int SCATTER_DURATION = 60;
constexpr int CHASE_DURATION = 180;
constexpr int LEVEL_DURATION = 960;
constexpr int LEVEL_PENALTY_DURATION = 240;
constexpr int FRUIT_THRESHOLD_1 = 174;
constexpr int FRUIT_THRESHOLD_2 = 74;
constexpr int FRUIT_DURATION = 30;
constexpr int ANGER_THRESHOLD_1 = 20;
constexpr int ANGER_THRESHOLD_2 = 10;
constexpr int PELLET_POINTS = 10;
constexpr int SUPER_PELLET_POINTS = 50;
constexpr int FRUIT_POINTS = 100;
constexpr int COMBO_MULTIPLIER = 200;
constexpr int GHOST_FRIGHT_STEPS = 40;
constexpr int EMPTY = 32;
constexpr int FRUIT_ROW = 17, FRUIT_COL = 13;
constexpr int PAC_SPAWN_ROW = 23, PAC_SPAWN_COL = 13;
constexpr int EXIT_ROW = 12, EXIT_COL = 13;          // ghost-house exit
constexpr int RED_SPAWN_ROW = 11, RED_SPAWN_COL = 13; // red's spawn cell, just outside the house
constexpr int INIT_PELLET_COUNT = 244;

// ---- wrapper constants (/root/reference/pacbot_rs/src/env.rs:23-38) ---------------------------
constexpr int NORMAL_TICKS_PER_STEP = 8;
constexpr int MIN_TICKS_PER_STEP = 4;
constexpr int MAX_TICKS_PER_STEP = 14; // exclusive
constexpr int TURN_PENALTY = -2;
constexpr int LAST_REWARD = 3000;

enum Dir : int { UP = 0, LEFT = 1, DOWN = 2, RIGHT = 3, NONE = 4 };
enum Act : int { A_STAY = 0, A_DOWN = 1, A_UP = 2, A_LEFT = 3, A_RIGHT = 4 };
enum Mode : int { SCATTER = 1, CHASE = 2 };

constexpr uint32_t WALL_ROWS[ROWS] = {
    0xfffffff, 0x8006001, 0xbdf6fbd, 0xbdf6fbd, 0xbdf6fbd, 0x8000001, 0xbdbfdbd, 0xbdbfdbd,
    0x8186181, 0xfdf6fbf, 0xfdf6fbf, 0xfd801bf, 0xfdbfdbf, 0xfdbfdbf, 0xfc3fc3f, 0xfdbfdbf,
    0xfdbfdbf, 0xfd801bf, 0xfdbfdbf, 0xfdbfdbf, 0x8006001, 0xbdf6fbd, 0xbdf6fbd, 0x8c00031,
    0xedbfdb7, 0xedbfdb7, 0x8186181, 0xbff6ffd, 0xbff6ffd, 0x8000001, 0xfffffff};

constexpr uint32_t PELLET_ROWS[ROWS] = {
    0x0000000, 0x7ff9ffe, 0x4209042, 0x4209042, 0x4209042, 0x7fffffe, 0x4240242, 0x4240242,
    0x7e79e7e, 0x0200040, 0x0200040, 0x0200040, 0x0200040, 0x0200040, 0x0200040, 0x0200040,
    0x0200040, 0x0200040, 0x0200040, 0x0200040, 0x7ff9ffe, 0x4209042, 0x4209042, 0x73f9fce,
    0x1240248, 0x1240248, 0x7e79e7e, 0x4009002, 0x4009002, 0x7fffffe, 0x0000000};

// Pellets live in HBM in OBSERVATION ORDER: bit index = col * 31 + (30 - row), i.e. exactly the
// flat position of the cell inside one observation channel (env.rs:124,425: [col][30-row]).
// The encoder then turns 4 consecutive bits into one 128-bit store with no transposition.
__host__ __device__ constexpr int fbit(int row, int col) { return col * ROWS + (ROWS - 1 - row); }

constexpr int SUPER_ROW[4] = {3, 3, 23, 23};
constexpr int SUPER_COL[4] = {1, 26, 1, 26};

struct MazeTables {
    uint32_t walls[32];                   // row bitboards (bit c of row r <=> wall)
    uint32_t init_pel[PEL_WORDS];         // initial pellets, observation order
    uint32_t super_mask[PEL_WORDS];       // the 4 super-pellet bits
    uint32_t wipe_keep[5][PEL_WORDS];     // env.rs:185-207: bits that SURVIVE wipe type k
    uint16_t node_rc[NUM_NODES];          // NODE_COORDS: row | col << 8 (row-major walkable cells)
    // IEEE-rounded quotients of the observation's timer channels (env.rs:463,468): u8 / 40.0 and
    // u8 / 180.0 for every possible u8 numerator.  Table lookups keep the divergent single-cell
    // code of the encoder free of division slow paths (out-of-line calls).
    float fright_q[256];
    float mode_q[256];
};

constexpr MazeTables make_tables() {
    MazeTables t{};
    for (int r = 0; r < ROWS; r++) t.walls[r] = WALL_ROWS[r];
    t.walls[31] = 0xfffffffu;
    int node = 0;
    for (int r = 0; r < ROWS; r++)
        for (int c = 0; c < COLS; c++) {
            const int i = fbit(r, c);
            if ((PELLET_ROWS[r] >> c) & 1u) t.init_pel[i >> 5] |= 1u << (i & 31);
            if (!((WALL_ROWS[r] >> c) & 1u)) {
                if (node < NUM_NODES) t.node_rc[node] = (uint16_t)(r | (c << 8));
                node++;
            }
            // wipe type 1: col < 14, 2: col >= 14, 3: row < 15, 4: row >= 15 are CLEARED
            const bool keep[5] = {true, !(c < COLS / 2), !(c >= COLS / 2), !(r < ROWS / 2),
                                  !(r >= ROWS / 2)};
            for (int k = 0; k < 5; k++)
                if (keep[k]) t.wipe_keep[k][i >> 5] |= 1u << (i & 31);
        }
    for (int k = 0; k < 4; k++) {
        const int i = fbit(SUPER_ROW[k], SUPER_COL[k]);
        t.super_mask[i >> 5] |= 1u << (i & 31);
    }
    for (int v = 0; v < 256; v++) {
        t.fright_q[v] = (float)v / (float)GHOST_FRIGHT_STEPS;
        t.mode_q[v] = (float)v / (float)CHASE_DURATION;
    }
    return t;
}

constexpr MazeTables H_TABLES = make_tables();

constexpr int count_nodes() {
    int n = 0;
    for (int r = 0; r < ROWS; r++)
        for (int c = 0; c < COLS; c++) n += !((WALL_ROWS[r] >> c) & 1u);
    return n;
}
constexpr int count_pellets() {
    int n = 0;
    for (int r = 0; r < ROWS; r++)
        for (int c = 0; c < COLS; c++) n += (PELLET_ROWS[r] >> c) & 1u;
    return n;
}
static_assert(count_nodes() == NUM_NODES, "walls must leave exactly 288 walkable nodes");
static_assert(count_pellets() == INIT_PELLET_COUNT, "pellet layout must hold 244 pellets");
static_assert(OBS_BYTES % 16 == 0 && CH_FLOATS % 4 == 0, "128-bit stores need 16 B granules");

}  // namespace pb
