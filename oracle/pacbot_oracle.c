/*
 * pacbot_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).  See pacbot_oracle.h.
 *
 * PARITY UNPINNED at the engine boundary (crate pacbot-rs @423622af is not in /root/reference and
 * cannot be fetched/built; the reference has no tests).  The wrapper half below follows
 * /root/reference/pacbot_rs/src/env.rs line by line (every function cites its lines); the engine
 * half restates the published Pacbot-2 engine algorithm that the crate ports, anchored on the
 * reference's call sites -- every engine assumption is tagged [E<n>] and
 * listed in DESIGN.md "Engine assumption ledger".
 */
#include "pacbot_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

/* ======================================================================================== */
/* Maze tables                                                                              */
/* ======================================================================================== */

/* Walls: bit c of row r set <=> wall.  Complement of computed_data/node_coords.json in the 31x28
 * grid -- forced by build.rs:78-108 (BFS over !wall_at indexes a map of node coords). */
const uint32_t PO_WALLS[PO_ROWS] = {
    0xfffffff, 0x8006001, 0xbdf6fbd, 0xbdf6fbd, 0xbdf6fbd, 0x8000001, 0xbdbfdbd, 0xbdbfdbd,
    0x8186181, 0xfdf6fbf, 0xfdf6fbf, 0xfd801bf, 0xfdbfdbf, 0xfdbfdbf, 0xfc3fc3f, 0xfdbfdbf,
    0xfdbfdbf, 0xfd801bf, 0xfdbfdbf, 0xfdbfdbf, 0x8006001, 0xbdf6fbd, 0xbdf6fbd, 0x8c00031,
    0xedbfdb7, 0xedbfdb7, 0x8186181, 0xbff6ffd, 0xbff6ffd, 0x8000001, 0xfffffff};

/* [E1] Initial pellets: every walkable cell except rows 9..19 outside cols {6,21} and the two
 * cells (23,13),(23,14); 244 pellets incl. the 4 super pellets at (3,1),(3,26),(23,1),(23,26)
 * (super-pellet cells pinned by env.rs:283,345,379,428,491). */
const uint32_t PO_INIT_PELLETS[PO_ROWS] = {
    0x0000000, 0x7ff9ffe, 0x4209042, 0x4209042, 0x4209042, 0x7fffffe, 0x4240242, 0x4240242,
    0x7e79e7e, 0x0200040, 0x0200040, 0x0200040, 0x0200040, 0x0200040, 0x0200040, 0x0200040,
    0x0200040, 0x0200040, 0x0200040, 0x0200040, 0x7ff9ffe, 0x4209042, 0x4209042, 0x73f9fce,
    0x1240248, 0x1240248, 0x7e79e7e, 0x4009002, 0x4009002, 0x7fffffe, 0x0000000};
#define PO_INIT_PELLET_COUNT 244

/* grid.rs:13 NODE_COORDS = computed_data/node_coords.json (row, col), row-major walkable cells */
const int8_t PO_NODE_COORDS[PO_NUM_NODES][2] = {
    {1,1}, {1,2}, {1,3}, {1,4}, {1,5}, {1,6}, {1,7}, {1,8}, {1,9}, {1,10}, {1,11}, {1,12},
    {1,15}, {1,16}, {1,17}, {1,18}, {1,19}, {1,20}, {1,21}, {1,22}, {1,23}, {1,24}, {1,25}, {1,26},
    {2,1}, {2,6}, {2,12}, {2,15}, {2,21}, {2,26}, {3,1}, {3,6}, {3,12}, {3,15}, {3,21}, {3,26},
    {4,1}, {4,6}, {4,12}, {4,15}, {4,21}, {4,26}, {5,1}, {5,2}, {5,3}, {5,4}, {5,5}, {5,6},
    {5,7}, {5,8}, {5,9}, {5,10}, {5,11}, {5,12}, {5,13}, {5,14}, {5,15}, {5,16}, {5,17}, {5,18},
    {5,19}, {5,20}, {5,21}, {5,22}, {5,23}, {5,24}, {5,25}, {5,26}, {6,1}, {6,6}, {6,9}, {6,18},
    {6,21}, {6,26}, {7,1}, {7,6}, {7,9}, {7,18}, {7,21}, {7,26}, {8,1}, {8,2}, {8,3}, {8,4},
    {8,5}, {8,6}, {8,9}, {8,10}, {8,11}, {8,12}, {8,15}, {8,16}, {8,17}, {8,18}, {8,21}, {8,22},
    {8,23}, {8,24}, {8,25}, {8,26}, {9,6}, {9,12}, {9,15}, {9,21}, {10,6}, {10,12}, {10,15}, {10,21},
    {11,6}, {11,9}, {11,10}, {11,11}, {11,12}, {11,13}, {11,14}, {11,15}, {11,16}, {11,17}, {11,18}, {11,21},
    {12,6}, {12,9}, {12,18}, {12,21}, {13,6}, {13,9}, {13,18}, {13,21}, {14,6}, {14,7}, {14,8}, {14,9},
    {14,18}, {14,19}, {14,20}, {14,21}, {15,6}, {15,9}, {15,18}, {15,21}, {16,6}, {16,9}, {16,18}, {16,21},
    {17,6}, {17,9}, {17,10}, {17,11}, {17,12}, {17,13}, {17,14}, {17,15}, {17,16}, {17,17}, {17,18}, {17,21},
    {18,6}, {18,9}, {18,18}, {18,21}, {19,6}, {19,9}, {19,18}, {19,21}, {20,1}, {20,2}, {20,3}, {20,4},
    {20,5}, {20,6}, {20,7}, {20,8}, {20,9}, {20,10}, {20,11}, {20,12}, {20,15}, {20,16}, {20,17}, {20,18},
    {20,19}, {20,20}, {20,21}, {20,22}, {20,23}, {20,24}, {20,25}, {20,26}, {21,1}, {21,6}, {21,12}, {21,15},
    {21,21}, {21,26}, {22,1}, {22,6}, {22,12}, {22,15}, {22,21}, {22,26}, {23,1}, {23,2}, {23,3}, {23,6},
    {23,7}, {23,8}, {23,9}, {23,10}, {23,11}, {23,12}, {23,13}, {23,14}, {23,15}, {23,16}, {23,17}, {23,18},
    {23,19}, {23,20}, {23,21}, {23,24}, {23,25}, {23,26}, {24,3}, {24,6}, {24,9}, {24,18}, {24,21}, {24,24},
    {25,3}, {25,6}, {25,9}, {25,18}, {25,21}, {25,24}, {26,1}, {26,2}, {26,3}, {26,4}, {26,5}, {26,6},
    {26,9}, {26,10}, {26,11}, {26,12}, {26,15}, {26,16}, {26,17}, {26,18}, {26,21}, {26,22}, {26,23}, {26,24},
    {26,25}, {26,26}, {27,1}, {27,12}, {27,15}, {27,26}, {28,1}, {28,12}, {28,15}, {28,26}, {29,1}, {29,2},
    {29,3}, {29,4}, {29,5}, {29,6}, {29,7}, {29,8}, {29,9}, {29,10}, {29,11}, {29,12}, {29,13}, {29,14},
    {29,15}, {29,16}, {29,17}, {29,18}, {29,19}, {29,20}, {29,21}, {29,22}, {29,23}, {29,24}, {29,25}, {29,26},
};

/* ======================================================================================== */
/* Engine constants (pacbot_rs_2::variables) -- [E2]                                         */
/* ======================================================================================== */
#define INIT_UPDATE_PERIOD 12 /* env.rs:25-26 "Ghosts move every 12 ticks" */
#define INIT_LEVEL 1          /* env.rs:278 (fresh game is not done)      */
#define INIT_LIVES 3          /* env.rs:242,278                            */
#define SCATTER_DURATION 60
#define CHASE_DURATION 180    /* GameMode::CHASE.duration(), env.rs:468   */
#define LEVEL_DURATION 960
#define LEVEL_PENALTY_DURATION 240
#define FRUIT_THRESHOLD_1 174
#define FRUIT_THRESHOLD_2 74
#define FRUIT_DURATION 30
#define ANGER_THRESHOLD_1 20
#define ANGER_THRESHOLD_2 10
#define PELLET_POINTS 10       /* variables::PELLET_POINTS, env.rs:431       */
#define SUPER_PELLET_POINTS 50 /* variables::SUPER_PELLET_POINTS, env.rs:429 */
#define FRUIT_POINTS 100       /* variables::FRUIT_POINTS, env.rs:437        */
#define COMBO_MULTIPLIER 200   /* variables::COMBO_MULTIPLIER, env.rs:441    */
#define GHOST_FRIGHT_STEPS 40  /* variables::GHOST_FRIGHT_STEPS, env.rs:463  */
#define EMPTY_COORD 32         /* env.rs:123 sentinel                        */

static const int8_t D_ROW[5] = {-1, 0, 1, 0, 0};
static const int8_t D_COL[5] = {0, -1, 0, 1, 0};
static const po_loc PACMAN_SPAWN = {23, 13, PO_RIGHT};
static const po_loc FRUIT_SPAWN = {17, 13, PO_NONE};
static const po_loc EMPTY_LOC = {EMPTY_COORD, EMPTY_COORD, PO_NONE};
static const po_loc GHOST_SPAWN[4] = {
    {11, 13, PO_LEFT}, {13, 13, PO_DOWN}, {14, 11, PO_UP}, {14, 15, PO_UP}};
static const int8_t GHOST_SCATTER[4][2] = {{-3, 25}, {-3, 2}, {31, 27}, {31, 0}};
static const uint8_t GHOST_TRAPPED_STEPS[4] = {0, 5, 16, 32};
#define HOUSE_EXIT_ROW 12
#define HOUSE_EXIT_COL 13

/* ---- env.rs:8 LAST_REWARD etc. (wrapper constants, authoritative) ------------------------- */
#define NORMAL_TICKS_PER_STEP 8 /* env.rs:27 */
#define MIN_TICKS_PER_STEP 4    /* env.rs:30 */
#define MAX_TICKS_PER_STEP 14   /* env.rs:33 (exclusive bound of gen_range, env.rs:148) */
#define TURN_PENALTY (-2)       /* env.rs:36 */
#define LAST_REWARD 3000        /* env.rs:38 */

/* ======================================================================================== */
/* Injected draws                                                                           */
/* ======================================================================================== */
static uint32_t draw_below(po_gym *e, uint32_t n) {
    uint32_t raw = e->draw ? e->draw(e->draw_ctx, e->rng_ctr) : 0u;
    e->rng_ctr += 1u;
    return (uint32_t)(((uint64_t)raw * (uint64_t)n) >> 32);
}

/* ======================================================================================== */
/* Engine                                                                                   */
/* ======================================================================================== */

static int in_bounds(int row, int col) {
    return row >= 0 && row < PO_ROWS && col >= 0 && col < PO_COLS;
}

/* GameState::wall_at (call sites env.rs:297-300,404,426): out of range is a wall [E3] */
int po_wall_at(int row, int col) {
    if (!in_bounds(row, col)) return 1;
    return (int)((PO_WALLS[row] >> col) & 1u);
}

/* GameState::pellet_at (call sites env.rs:189,283,427,490): out of range is no pellet [E3] */
int po_pellet_at(const po_game_state *g, int row, int col) {
    if (!in_bounds(row, col)) return 0;
    return (int)((g->pellets[row] >> col) & 1u);
}

static int loc_is_empty(const po_loc *l) { return l->row == EMPTY_COORD && l->col == EMPTY_COORD; }
static int loc_collides(const po_loc *a, const po_loc *b) {
    return a->row == b->row && a->col == b->col;
}
static uint8_t reversed_dir(uint8_t d) {
    switch (d) {
    case PO_UP: return PO_DOWN;
    case PO_LEFT: return PO_RIGHT;
    case PO_DOWN: return PO_UP;
    case PO_RIGHT: return PO_LEFT;
    default: return PO_NONE;
    }
}
static int ghost_is_frightened(const po_ghost *g) { return g->fright_steps > 0; } /* env.rs:253 */
static int ghost_is_trapped(const po_ghost *g) { return g->trapped_steps > 0; }
static int fruit_exists(const po_game_state *g) { return g->fruit_steps > 0; } /* env.rs:341 */

/* [E4] ghost house interior: rows 13..14, cols 11..15 */
static int ghost_spawn_at(int row, int col) {
    if (!in_bounds(row, col)) return 0;
    return row >= 13 && row <= 14 && col >= 11 && col <= 15;
}

static int dist_sq(int r1, int c1, int r2, int c2) {
    int dr = r1 - r2, dc = c1 - c2;
    return dr * dr + dc * dc;
}

/* [E5] ghost initial / reset state: off the board (loc = EMPTY), next_loc = spawn, spawning */
static void ghost_reset(po_ghost *gh) {
    gh->spawning = 1;
    gh->trapped_steps = GHOST_TRAPPED_STEPS[gh->color];
    gh->fright_steps = 0;
    gh->loc = EMPTY_LOC;
    gh->next_loc = GHOST_SPAWN[gh->color];
}

/* GameState::new() -- SURVEY B.1 [E6] */
void po_game_new(po_game_state *g) {
    memset(g, 0, sizeof(*g));
    g->curr_ticks = 0;
    g->update_period = INIT_UPDATE_PERIOD;
    g->mode = PO_MODE_SCATTER;
    g->paused = 1; /* env.rs:134,211 force it to false */
    g->pause_on_update = 0;
    g->mode_steps = SCATTER_DURATION;
    g->level_steps = LEVEL_DURATION;
    g->curr_score = 0;
    g->curr_level = INIT_LEVEL;
    g->curr_lives = INIT_LIVES;
    g->ghost_combo = 0;
    g->pacman_loc = PACMAN_SPAWN;
    g->fruit_loc = FRUIT_SPAWN;
    g->fruit_steps = 0;
    for (int i = 0; i < 4; i++) {
        g->ghosts[i].color = (uint8_t)i;
        g->ghosts[i].eaten = 0;
        ghost_reset(&g->ghosts[i]);
    }
    memcpy(g->pellets, PO_INIT_PELLETS, sizeof(g->pellets));
    g->num_pellets = PO_INIT_PELLET_COUNT;
    /* (the engine collects the pellet under the spawn cell here; (23,13) holds none [E1]) */
}

/* GameState::decrement_num_pellets (call sites env.rs:203,381): pure counter, no side effects */
void po_game_decrement_num_pellets(po_game_state *g) {
    if (g->num_pellets > 0) g->num_pellets--;
}

static void reset_all_ghosts(po_game_state *g) {
    for (int i = 0; i < 4; i++) ghost_reset(&g->ghosts[i]);
}

/* [E7] reversal = "trap for one step": plan() of a trapped ghost keeps moving one cell and flips
 * its direction.  Used by frighten and by mode switches. */
static void reverse_all_ghosts(po_game_state *g) {
    for (int i = 0; i < 4; i++)
        if (!ghost_is_trapped(&g->ghosts[i])) g->ghosts[i].trapped_steps = 1;
}
static void frighten_all_ghosts(po_game_state *g) {
    g->ghost_combo = 0;
    for (int i = 0; i < 4; i++) {
        g->ghosts[i].fright_steps = GHOST_FRIGHT_STEPS;
        if (!ghost_is_trapped(&g->ghosts[i])) g->ghosts[i].trapped_steps = 1;
    }
}

static void lower_update_period(po_game_state *g) {
    int p = (int)g->update_period - 2;
    g->update_period = (uint8_t)(p < 1 ? 1 : p);
}

/* SURVEY B.7 [E8] */
static void death_reset(po_game_state *g) {
    g->pause_on_update = 1;
    g->pacman_loc = EMPTY_LOC;
    if (g->curr_lives > 0) g->curr_lives--;
    if (g->curr_lives > 0) {
        g->mode = PO_MODE_SCATTER;
        g->mode_steps = SCATTER_DURATION;
    }
    g->fruit_steps = 0;
    reset_all_ghosts(g);
}

/* [E9] board cleared: fresh board, next level (the wrapper reports done, env.rs:278) */
static void level_reset(po_game_state *g) {
    g->pause_on_update = 1;
    g->pacman_loc = EMPTY_LOC;
    g->level_steps = LEVEL_DURATION;
    g->mode = PO_MODE_SCATTER;
    g->mode_steps = SCATTER_DURATION;
    g->fruit_steps = 0;
    reset_all_ghosts(g);
    memcpy(g->pellets, PO_INIT_PELLETS, sizeof(g->pellets));
    g->num_pellets = PO_INIT_PELLET_COUNT;
}
static void increment_level(po_game_state *g) {
    if (g->curr_level == 255) return;
    g->curr_level++;
    int p = INIT_UPDATE_PERIOD - 2 * ((int)g->curr_level - 1);
    g->update_period = (uint8_t)(p < 1 ? 1 : p);
}

/* SURVEY B.5 [E10].  A collision needs Pac-Man on the board: right after a board clear or a death
 * everything sits on the (32,32) sentinel and must not "collide" (otherwise clearing the board
 * through set_pacman_location would immediately cost a life and env.rs:245-247 could never pay
 * LAST_REWARD). */
static void check_collisions(po_game_state *g) {
    if (loc_is_empty(&g->pacman_loc)) return;
    uint8_t respawn_flag = 0;
    int n_respawn = 0;
    for (int i = 0; i < 4; i++) {
        po_ghost *gh = &g->ghosts[i];
        if (loc_collides(&g->pacman_loc, &gh->loc)) {
            if (gh->eaten) continue;
            if (ghost_is_frightened(gh)) {
                respawn_flag |= (uint8_t)(1u << i);
                n_respawn++;
            } else {
                death_reset(g);
                return;
            }
        }
    }
    if (n_respawn == 0) return;
    for (int i = 0; i < 4; i++) {
        if (!(respawn_flag & (1u << i))) continue;
        po_ghost *gh = &g->ghosts[i];
        /* GhostState::respawn */
        gh->spawning = 1;
        gh->eaten = 1;
        gh->loc = EMPTY_LOC;
        gh->next_loc = GHOST_SPAWN[i == PO_RED ? PO_PINK : i];
        gh->next_loc.dir = PO_UP;
        g->curr_score = (uint16_t)(g->curr_score + (uint16_t)(COMBO_MULTIPLIER << g->ghost_combo));
        g->ghost_combo++;
    }
}

/* SURVEY B.3 collect_pellet [E11] */
static void collect_pellet(po_game_state *g, int row, int col) {
    if (fruit_exists(g) && loc_collides(&g->pacman_loc, &g->fruit_loc)) {
        g->fruit_steps = 0;
        g->curr_score = (uint16_t)(g->curr_score + FRUIT_POINTS);
    }
    if (!po_pellet_at(g, row, col)) return;
    g->pellets[row] &= ~(1u << col);
    po_game_decrement_num_pellets(g);
    int super_pellet = ((row == 3) || (row == 23)) && ((col == 1) || (col == 26));
    if (super_pellet) {
        frighten_all_ghosts(g);
        g->curr_score = (uint16_t)(g->curr_score + SUPER_PELLET_POINTS);
    } else {
        g->curr_score = (uint16_t)(g->curr_score + PELLET_POINTS);
    }
    uint16_t n = g->num_pellets;
    if ((n == FRUIT_THRESHOLD_1 || n == FRUIT_THRESHOLD_2) && !fruit_exists(g))
        g->fruit_steps = FRUIT_DURATION;
    if (n == ANGER_THRESHOLD_1 || n == ANGER_THRESHOLD_2) {
        lower_update_period(g);
        g->mode = PO_MODE_CHASE;
        g->mode_steps = CHASE_DURATION;
    } else if (n == 0) {
        level_reset(g);
        increment_level(g);
    }
}

/* GameState::set_pacman_location (call site env.rs:405) -- SURVEY B.3 [E12] */
void po_game_set_pacman_location(po_game_state *g, int8_t row, int8_t col) {
    if (g->paused) return;
    int dr = row - g->pacman_loc.row, dc = col - g->pacman_loc.col;
    if (dr == -1 && dc == 0) g->pacman_loc.dir = PO_UP;
    else if (dr == 0 && dc == -1) g->pacman_loc.dir = PO_LEFT;
    else if (dr == 1 && dc == 0) g->pacman_loc.dir = PO_DOWN;
    else if (dr == 0 && dc == 1) g->pacman_loc.dir = PO_RIGHT;
    /* any other delta (incl. staying put): facing direction unchanged */
    g->pacman_loc.row = row;
    g->pacman_loc.col = col;
    collect_pellet(g, row, col);
    check_collisions(g);
}

/* GhostState::update -- SURVEY B.2 [E13] */
static void ghost_update(po_ghost *gh) {
    if (loc_collides(&gh->loc, &GHOST_SPAWN[PO_RED])) gh->spawning = 0;
    if (gh->eaten) {
        gh->eaten = 0;
        gh->fright_steps = 0;
    }
    if (ghost_is_frightened(gh)) gh->fright_steps--;
    gh->loc = gh->next_loc;
}

/* chase targets -- SURVEY B.4 [E14] */
static void ahead_coords(const po_loc *l, int spaces, int *row, int *col) {
    int ro = D_ROW[l->dir] * spaces, co = D_COL[l->dir] * spaces;
    if (l->dir == PO_UP) co -= spaces; /* arcade overflow quirk */
    *row = l->row + ro;
    *col = l->col + co;
}
static void chase_target(const po_game_state *g, int color, int *tr, int *tc) {
    const po_loc *p = &g->pacman_loc;
    switch (color) {
    case PO_RED:
        *tr = p->row;
        *tc = p->col;
        break;
    case PO_PINK:
        ahead_coords(p, 4, tr, tc);
        break;
    case PO_CYAN: {
        int pr, pc;
        ahead_coords(p, 2, &pr, &pc);
        *tr = 2 * pr - g->ghosts[PO_RED].loc.row;
        *tc = 2 * pc - g->ghosts[PO_RED].loc.col;
        break;
    }
    default: { /* orange */
        const po_loc *o = &g->ghosts[PO_ORANGE].loc;
        if (dist_sq(o->row, o->col, p->row, p->col) >= 64) {
            *tr = p->row;
            *tc = p->col;
        } else {
            *tr = GHOST_SCATTER[PO_ORANGE][0];
            *tc = GHOST_SCATTER[PO_ORANGE][1];
        }
    }
    }
}

/* GhostState::plan -- SURVEY B.4 [E15]; the only RNG use inside step() */
static void ghost_plan(po_game_state *g, int i, po_gym *rng_owner) {
    po_ghost *gh = &g->ghosts[i];
    if (loc_is_empty(&gh->loc)) return;
    /* next_loc = loc advanced one cell along loc.dir */
    gh->next_loc.row = (int8_t)(gh->loc.row + D_ROW[gh->loc.dir]);
    gh->next_loc.col = (int8_t)(gh->loc.col + D_COL[gh->loc.dir]);
    gh->next_loc.dir = gh->loc.dir;
    if (ghost_is_trapped(gh)) {
        gh->next_loc.dir = reversed_dir(gh->next_loc.dir);
        gh->trapped_steps--;
        return;
    }
    int tr = 0, tc = 0;
    if (g->mode == PO_MODE_CHASE) {
        chase_target(g, i, &tr, &tc);
    } else if (g->mode == PO_MODE_SCATTER) {
        tr = GHOST_SCATTER[i][0];
        tc = GHOST_SCATTER[i][1];
    }
    /* [E15] a spawning ghost is lured out of the house: until it stands on red's spawn cell
     * (11,13, the cell above the house exit) or is about to step onto it, that cell is its
     * target; from there on it targets by mode even though `spawning` is only cleared by the
     * next update() */
    if (gh->spawning && !loc_collides(&gh->loc, &GHOST_SPAWN[PO_RED]) &&
        !loc_collides(&gh->next_loc, &GHOST_SPAWN[PO_RED])) {
        tr = GHOST_SPAWN[PO_RED].row;
        tc = GHOST_SPAWN[PO_RED].col;
    }
    int n_valid = 0, valid[4], d2[4];
    uint8_t rev = reversed_dir(gh->next_loc.dir);
    for (int d = 0; d < 4; d++) {
        int r = gh->next_loc.row + D_ROW[d], c = gh->next_loc.col + D_COL[d];
        d2[d] = dist_sq(r, c, tr, tc);
        valid[d] = !po_wall_at(r, c);
        if (gh->spawning) {
            if (ghost_spawn_at(r, c)) valid[d] = 1;
            if (r == HOUSE_EXIT_ROW && c == HOUSE_EXIT_COL) valid[d] = 1;
        }
        if (d == rev) valid[d] = 0;
        if (valid[d]) n_valid++;
    }
    if (n_valid == 0) return; /* boxed in: keep heading */
    if (gh->fright_steps > 1) {
        int k = (int)draw_below(rng_owner, (uint32_t)n_valid);
        for (int d = 0, cnt = 0; d < 4; d++) {
            if (!valid[d]) continue;
            if (cnt == k) {
                gh->next_loc.dir = (uint8_t)d;
                return;
            }
            cnt++;
        }
    }
    int best = PO_UP, best_d = 0x7fffffff;
    for (int d = 0; d < 4; d++) {
        if (!valid[d]) continue;
        if (d2[d] < best_d) {
            best = d;
            best_d = d2[d];
        }
    }
    gh->next_loc.dir = (uint8_t)best;
}

/* SURVEY B.6 [E16] */
static void handle_step_events(po_game_state *g) {
    if (g->mode_steps == 0) {
        if (g->mode == PO_MODE_CHASE) {
            g->mode = PO_MODE_SCATTER;
            g->mode_steps = SCATTER_DURATION;
        } else {
            g->mode = PO_MODE_CHASE;
            g->mode_steps = CHASE_DURATION;
        }
        reverse_all_ghosts(g);
    } else {
        g->mode_steps--;
    }
    if (g->level_steps == 0) {
        lower_update_period(g);
        g->level_steps = LEVEL_PENALTY_DURATION;
    } else {
        g->level_steps--;
    }
    if (g->fruit_steps > 0) g->fruit_steps--;
}

/* GameState::step (call site env.rs:227) -- SURVEY B.2 [E17] */
void po_game_step(po_game_state *g, po_gym *rng_owner) {
    if (g->paused) return;
    if (g->curr_ticks % (uint32_t)g->update_period == 0) {
        for (int i = 0; i < 4; i++) ghost_update(&g->ghosts[i]);
        if (loc_is_empty(&g->pacman_loc)) g->pacman_loc = PACMAN_SPAWN; /* try_respawn_pacman */
        if (g->pause_on_update) {
            g->paused = 1;
            g->pause_on_update = 0;
        }
        check_collisions(g);
        handle_step_events(g);
        for (int i = 0; i < 4; i++) ghost_plan(g, i, rng_owner);
    }
    g->curr_ticks++;
}

/* ======================================================================================== */
/* Wrapper: env.rs                                                                          */
/* ======================================================================================== */

/* env.rs:112-119 */
static void modify_bit_u32(uint32_t *num, int bit_idx, int bit_val) {
    if (bit_val) *num |= (1u << bit_idx);
    else *num &= ~(1u << bit_idx);
}

/* env.rs:377-384 */
static void apply_configuration(po_gym *e) {
    if (e->remove_super_pellets) {
        static const int8_t sp[4][2] = {{3, 1}, {23, 1}, {3, 26}, {23, 26}};
        for (int i = 0; i < 4; i++) {
            modify_bit_u32(&e->gs.pellets[sp[i][0]], sp[i][1], 0);
            po_game_decrement_num_pellets(&e->gs);
        }
    }
}

/* env.rs:362-375 new_with_state */
void po_gym_new_with_config(po_gym *e, int random_start, int random_ticks, int randomize_ghosts,
                            int remove_super_pellets) {
    memset(e, 0, sizeof(*e));
    po_game_new(&e->gs);
    e->gs.paused = 0; /* env.rs:134 */
    e->random_start = (uint8_t)(random_start != 0);
    e->last_score = 0;
    e->last_action = PO_ACT_STAY;
    e->ticks_per_step = NORMAL_TICKS_PER_STEP;
    e->random_ticks = (uint8_t)(random_ticks != 0);
    e->randomize_ghosts = (uint8_t)(randomize_ghosts != 0);
    e->remove_super_pellets = (uint8_t)(remove_super_pellets != 0);
    apply_configuration(e);
}

/* env.rs:133-138 */
void po_gym_new(po_gym *e, int random_start, int random_ticks) {
    po_gym_new_with_config(e, random_start, random_ticks, 0, 0);
}

void po_gym_set_draws(po_gym *e, po_draw_fn fn, void *ctx) {
    e->draw = fn;
    e->draw_ctx = ctx;
}

/* build.rs:141-149 VALID_ACTIONS[node] = [true, (r+1,c), (r-1,c), (r,c-1), (r,c+1) in nodes] */
static void valid_actions_at(int row, int col, uint8_t out[5]) {
    out[0] = 1;
    out[1] = (uint8_t)!po_wall_at(row + 1, col);
    out[2] = (uint8_t)!po_wall_at(row - 1, col);
    out[3] = (uint8_t)!po_wall_at(row, col - 1);
    out[4] = (uint8_t)!po_wall_at(row, col + 1);
}

/* env.rs:140-212 */
void po_gym_reset(po_gym *e) {
    e->last_score = 0;                      /* :141 */
    po_game_new(&e->gs);                    /* :142 */
    apply_configuration(e);                 /* :144 */
    if (e->random_ticks)                    /* :147-149 gen_range(4..14) */
        e->ticks_per_step =
            MIN_TICKS_PER_STEP + draw_below(e, MAX_TICKS_PER_STEP - MIN_TICKS_PER_STEP);
    if (e->random_start) {                  /* :151 */
        uint32_t node = draw_below(e, PO_NUM_NODES); /* :152-154 NODE_COORDS.choose */
        e->gs.pacman_loc.row = PO_NODE_COORDS[node][0];
        e->gs.pacman_loc.col = PO_NODE_COORDS[node][1];
        e->gs.pacman_loc.dir = PO_UP;       /* :155 (direct assignment: no pellet eaten) */
        if (e->randomize_ghosts) {          /* :157-182 */
            for (int i = 0; i < 4; i++) {
                po_ghost *gh = &e->gs.ghosts[i];
                gh->trapped_steps = 0;      /* :159 */
                uint32_t gn = draw_below(e, PO_NUM_NODES);
                int r = PO_NODE_COORDS[gn][0], c = PO_NODE_COORDS[gn][1];
                uint8_t va[5];
                valid_actions_at(r, c, va);
                int seen = 0, idx = -1;     /* :163-168 second `true` entry (.nth(1)) */
                for (int k = 0; k < 5; k++)
                    if (va[k] && seen++ == 1) {
                        idx = k;
                        break;
                    }
                uint8_t d = PO_NONE;
                switch (idx) {              /* :172-177 (3 => Right, 4 => Left: sic) */
                case 1: d = PO_DOWN; break;
                case 2: d = PO_UP; break;
                case 3: d = PO_RIGHT; break;
                case 4: d = PO_LEFT; break;
                default: break;             /* unreachable: every node has degree >= 2 */
                }
                gh->loc.row = (int8_t)r;
                gh->loc.col = (int8_t)c;
                gh->loc.dir = d;
                gh->next_loc = gh->loc;     /* :180 */
            }
        }
        uint32_t wipe_type = draw_below(e, 5); /* :185 gen_range(0..=4) */
        if (wipe_type != 0) {
            for (int row = 0; row < 31; row++) {
                for (int col = 0; col < 28; col++) {
                    int in_half = 0;
                    switch (wipe_type) {    /* :190-196 */
                    case 1: in_half = col < 28 / 2; break;
                    case 2: in_half = col >= 28 / 2; break;
                    case 3: in_half = row < 31 / 2; break;
                    case 4: in_half = row >= 31 / 2; break;
                    }
                    if (po_pellet_at(&e->gs, row, col) && in_half) {
                        modify_bit_u32(&e->gs.pellets[row], col, 0);
                        po_game_decrement_num_pellets(&e->gs);
                    }
                }
            }
        }
    }
    e->last_action = PO_ACT_STAY;           /* :210 */
    e->gs.paused = 0;                       /* :211 */
}

/* i8 helpers mirroring Rust's saturating_sub on i8 (env.rs:400-401) */
static int8_t i8_saturating_sub1(int8_t v) { return v == INT8_MIN ? v : (int8_t)(v - 1); }

/* env.rs:395-407 */
static void move_one_cell(po_gym *e, uint8_t action) {
    po_loc old = e->gs.pacman_loc;
    int8_t nr = old.row, nc = old.col;
    switch (action) {
    case PO_ACT_STAY: break;
    case PO_ACT_RIGHT: nc = (int8_t)(old.col + 1); break;
    case PO_ACT_LEFT: nc = i8_saturating_sub1(old.col); break;
    case PO_ACT_UP: nr = i8_saturating_sub1(old.row); break;
    case PO_ACT_DOWN: nr = (int8_t)(old.row + 1); break;
    }
    if (!po_wall_at(nr, nc)) po_game_set_pacman_location(&e->gs, nr, nc);
}

/* env.rs:277-279 */
int po_gym_is_done(const po_gym *e) {
    return e->gs.curr_lives < 3 || e->gs.curr_level != INIT_LEVEL;
}

/* env.rs:216-267 */
int po_gym_step(po_gym *e, uint8_t action, int32_t *reward_out, uint8_t *done_out) {
    if (action > 4) return -1;              /* :83-88 */
    move_one_cell(e, action);               /* :218 */
    int turn_penalty =                      /* :221-225 */
        (e->last_action == action || e->last_action == PO_ACT_STAY) ? 0 : TURN_PENALTY;
    for (uint32_t t = 0; t < e->ticks_per_step; t++) { /* :226-231 */
        po_game_step(&e->gs, e);
        if (po_gym_is_done(e)) break;
    }
    e->last_action = action;                /* :232 */
    const po_game_state *g = &e->gs;
    int done = po_gym_is_done(e);           /* :236 */
    int32_t reward = ((int32_t)g->curr_score - (int32_t)e->last_score) + turn_penalty; /* :240 */
    if (done) {                             /* :241-249 */
        if (g->curr_lives < 3) reward += -200;
        else reward += LAST_REWARD;
    }
    for (int i = 0; i < 4; i++) {           /* :251-263 */
        const po_ghost *gh = &g->ghosts[i];
        if (!ghost_is_frightened(gh)) {
            int dr = (int8_t)(g->pacman_loc.row - gh->loc.row);
            int dc = (int8_t)(g->pacman_loc.col - gh->loc.col);
            int d = abs(dr) + abs(dc);
            reward += (d == 1) ? -50 : (d == 2) ? -20 : 0;
        }
    }
    e->last_score = g->curr_score;          /* :264 */
    *reward_out = reward;
    *done_out = (uint8_t)done;
    return 0;
}

uint32_t po_gym_score(const po_gym *e) { return e->gs.curr_score; }             /* :269-271 */
uint8_t po_gym_lives(const po_gym *e) { return e->gs.curr_lives; }              /* :273-275 */
uint16_t po_gym_remaining_pellets(const po_gym *e) { return e->gs.num_pellets; } /* :287-289 */

/* env.rs:281-285 */
int po_gym_first_ai_done(const po_gym *e) {
    static const int8_t sp[4][2] = {{3, 1}, {3, 26}, {23, 1}, {23, 26}};
    for (int i = 0; i < 4; i++)
        if (po_pellet_at(&e->gs, sp[i][0], sp[i][1])) return 0;
    for (int i = 0; i < 4; i++)
        if (ghost_is_frightened(&e->gs.ghosts[i])) return 0;
    return 1;
}

/* env.rs:293-302 */
void po_gym_action_mask(const po_gym *e, uint8_t out[5]) {
    int r = e->gs.pacman_loc.row, c = e->gs.pacman_loc.col;
    out[0] = 1;
    out[1] = (uint8_t)!po_wall_at(r + 1, c);
    out[2] = (uint8_t)!po_wall_at(r - 1, c);
    out[3] = (uint8_t)!po_wall_at(r, c - 1);
    out[4] = (uint8_t)!po_wall_at(r, c + 1);
}

/* env.rs:122-128 loc_to_pos */
static int loc_to_pos(const po_loc *l, int *x, int *y) {
    if (l->row != 32 && l->col != 32) {
        *x = l->col;
        *y = 31 - l->row - 1;
        return 1;
    }
    return 0;
}

#define OBS_AT(ch, x, y) out[((ch) * 28 + (x)) * 31 + (y)]

/* env.rs:410-500 */
void po_gym_obs(const po_gym *e, float *out) {
    const po_game_state *g = &e->gs;
    memset(out, 0, sizeof(float) * PO_OBS_FLOATS); /* :412 */
    for (int row = 0; row < 31; row++) {           /* :423-443 */
        for (int col = 0; col < 28; col++) {
            int obs_row = 31 - row - 1;
            OBS_AT(0, col, obs_row) = (float)(uint8_t)po_wall_at(row, col);
            uint16_t v;
            if (po_pellet_at(g, row, col)) {
                v = (uint16_t)((((row == 3) || (row == 23)) && ((col == 1) || (col == 26))
                                    ? SUPER_PELLET_POINTS
                                    : PELLET_POINTS) +
                               (g->num_pellets == 1 ? LAST_REWARD : 0));
            } else if (fruit_exists(g) && col == g->fruit_loc.col && row == g->fruit_loc.row) {
                v = FRUIT_POINTS;
            } else {
                v = 0;
            }
            OBS_AT(1, col, obs_row) = (float)v / (float)COMBO_MULTIPLIER;
        }
    }
    int px, py;                                    /* :446-457 */
    int has_p = loc_to_pos(&g->pacman_loc, &px, &py);
    if (has_p) {
        OBS_AT(2, px, py) = 1.0f;
        OBS_AT(3, px, py) = 1.0f;
    }
    for (int i = 0; i < 4; i++) {                  /* :459-471 */
        const po_ghost *gh = &g->ghosts[i];
        int gx, gy;
        if (!loc_to_pos(&gh->loc, &gx, &gy)) continue;
        OBS_AT(4 + i, gx, gy) = 1.0f;
        if (ghost_is_frightened(gh)) {
            OBS_AT(14, gx, gy) = (float)gh->fright_steps / (float)GHOST_FRIGHT_STEPS;
            OBS_AT(1, gx, gy) += (float)(int32_t)(1 << g->ghost_combo); /* 2_i32.pow(combo) */
        } else {
            int state_index = (g->mode == PO_MODE_CHASE) ? 1 : 0;
            OBS_AT(12 + state_index, gx, gy) = (float)g->mode_steps / (float)CHASE_DURATION;
        }
        OBS_AT(8 + i, gx, gy) = 1.0f;              /* :476-480 */
    }
    float speed = (float)e->ticks_per_step / (float)g->update_period; /* :482-484 */
    for (int x = 0; x < 28; x++)
        for (int y = 0; y < 31; y++) OBS_AT(15, x, y) = speed;
    for (int row = 0; row < 31; row++)             /* :487-497 */
        for (int col = 0; col < 28; col++)
            if (po_pellet_at(g, row, col) && ((row == 3) || (row == 23)) &&
                ((col == 1) || (col == 26)))
                OBS_AT(16, col, 31 - row - 1) = 1.0f;
}

/* ======================================================================================== */
/* Flat state export / import                                                               */
/* ======================================================================================== */
void po_gym_export(const po_gym *e, po_flat_state *o) {
    const po_game_state *g = &e->gs;
    memset(o, 0, sizeof(*o));
    o->curr_ticks = g->curr_ticks;
    o->rng_ctr = e->rng_ctr;
    memcpy(o->pellets, g->pellets, sizeof(o->pellets));
    o->level_steps = g->level_steps;
    o->curr_score = g->curr_score;
    o->num_pellets = g->num_pellets;
    o->last_score = e->last_score;
    o->update_period = g->update_period;
    o->mode = g->mode;
    o->paused = g->paused;
    o->pause_on_update = g->pause_on_update;
    o->mode_steps = g->mode_steps;
    o->curr_level = g->curr_level;
    o->curr_lives = g->curr_lives;
    o->ghost_combo = g->ghost_combo;
    o->fruit_steps = g->fruit_steps;
    o->last_action = e->last_action;
    o->ticks_per_step = (uint8_t)e->ticks_per_step;
    o->random_start = e->random_start;
    o->random_ticks = e->random_ticks;
    o->randomize_ghosts = e->randomize_ghosts;
    o->remove_super_pellets = e->remove_super_pellets;
    o->pac_row = g->pacman_loc.row;
    o->pac_col = g->pacman_loc.col;
    o->pac_dir = g->pacman_loc.dir;
    for (int i = 0; i < 4; i++) {
        const po_ghost *gh = &g->ghosts[i];
        o->g_row[i] = gh->loc.row;
        o->g_col[i] = gh->loc.col;
        o->g_dir[i] = gh->loc.dir;
        o->g_next_row[i] = gh->next_loc.row;
        o->g_next_col[i] = gh->next_loc.col;
        o->g_next_dir[i] = gh->next_loc.dir;
        o->g_fright[i] = gh->fright_steps;
        o->g_trapped[i] = gh->trapped_steps;
        o->g_spawning[i] = gh->spawning;
        o->g_eaten[i] = gh->eaten;
    }
}

void po_gym_import(po_gym *e, const po_flat_state *o) {
    po_game_state *g = &e->gs;
    g->curr_ticks = o->curr_ticks;
    e->rng_ctr = o->rng_ctr;
    memcpy(g->pellets, o->pellets, sizeof(g->pellets));
    g->level_steps = o->level_steps;
    g->curr_score = o->curr_score;
    g->num_pellets = o->num_pellets;
    e->last_score = o->last_score;
    g->update_period = o->update_period;
    g->mode = o->mode;
    g->paused = o->paused;
    g->pause_on_update = o->pause_on_update;
    g->mode_steps = o->mode_steps;
    g->curr_level = o->curr_level;
    g->curr_lives = o->curr_lives;
    g->ghost_combo = o->ghost_combo;
    g->fruit_steps = o->fruit_steps;
    g->fruit_loc = FRUIT_SPAWN;
    e->last_action = o->last_action;
    e->ticks_per_step = o->ticks_per_step;
    e->random_start = o->random_start;
    e->random_ticks = o->random_ticks;
    e->randomize_ghosts = o->randomize_ghosts;
    e->remove_super_pellets = o->remove_super_pellets;
    g->pacman_loc.row = o->pac_row;
    g->pacman_loc.col = o->pac_col;
    g->pacman_loc.dir = o->pac_dir;
    for (int i = 0; i < 4; i++) {
        po_ghost *gh = &g->ghosts[i];
        gh->color = (uint8_t)i;
        gh->loc.row = o->g_row[i];
        gh->loc.col = o->g_col[i];
        gh->loc.dir = o->g_dir[i];
        gh->next_loc.row = o->g_next_row[i];
        gh->next_loc.col = o->g_next_col[i];
        gh->next_loc.dir = o->g_next_dir[i];
        gh->fright_steps = o->g_fright[i];
        gh->trapped_steps = o->g_trapped[i];
        gh->spawning = o->g_spawning[i];
        gh->eaten = o->g_eaten[i];
    }
}

size_t po_sizeof_gym(void) { return sizeof(po_gym); }
size_t po_sizeof_flat(void) { return sizeof(po_flat_state); }

/* ======================================================================================== */
/* Batch helpers (tests)                                                                    */
/* ======================================================================================== */
struct po_batch {
    int64_t n;
    po_gym *envs;
    const uint32_t *tape;
    int64_t tape_len;
    struct tape_ctx {
        const uint32_t *base;
        int64_t len;
        uint64_t key; /* counter mode: env_key(seed, gid) */
    } *ctx;
};

static uint32_t tape_draw(void *vctx, uint32_t ctr) {
    struct tape_ctx *c = (struct tape_ctx *)vctx;
    if (!c->base || c->len <= 0) return 0u;
    return c->base[(int64_t)(ctr % (uint32_t)c->len)];
}

/* ---- the product's counter RNG (include/pacbot_b200.h "Counter RNG"), restated so that a run
 * WITHOUT a draw tape (what bench.py times) can be replayed on the oracle --------------------- */
static uint64_t po_mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
static uint64_t po_env_key(uint64_t seed, uint64_t gid) {
    return po_mix64(seed + 0x9e3779b97f4a7c15ull * (gid + 1ull));
}
uint32_t po_counter_draw_raw(uint64_t seed, uint64_t gid, uint32_t ctr) {
    return (uint32_t)(po_mix64(po_env_key(seed, gid) + (uint64_t)ctr * 0x9e3779b97f4a7c15ull +
                               0x632be59bd9b4e019ull) >> 32);
}
uint32_t po_counter_action_raw(uint64_t seed, uint64_t gid, uint64_t call) {
    return (uint32_t)(po_mix64(po_env_key(seed, gid) + call * 0xd1342543de82ef95ull +
                               0xa0761d6478bd642full) >> 32);
}
static uint32_t counter_key_draw(void *vctx, uint32_t ctr) {
    struct tape_ctx *c = (struct tape_ctx *)vctx;
    return (uint32_t)(po_mix64(c->key + (uint64_t)ctr * 0x9e3779b97f4a7c15ull +
                               0x632be59bd9b4e019ull) >> 32);
}

/* ---- the batch loops run on all host threads once the batch is large enough to pay: a plain
 * pthread fork/join per call (no OpenMP runtime in this image's gcc) -------------------------- */
#define PO_PAR_MIN 2048
typedef void (*po_range_fn)(void *arg, int64_t lo, int64_t hi);
typedef struct {
    po_range_fn fn;
    void *arg;
    int64_t lo, hi;
} po_par_task;
static void *po_par_thread(void *v) {
    po_par_task *t = (po_par_task *)v;
    t->fn(t->arg, t->lo, t->hi);
    return NULL;
}
static int po_host_threads(void) {
    static int n = 0;
    if (n == 0) {
        const char *e = getenv("PO_THREADS");
        long v = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
        n = v < 1 ? 1 : (v > 256 ? 256 : (int)v);
    }
    return n;
}
static void po_parallel_for(int64_t n, int64_t min_par, po_range_fn fn, void *arg) {
    int nt = po_host_threads();
    if (n < min_par || nt <= 1) {
        fn(arg, 0, n);
        return;
    }
    if ((int64_t)nt > n) nt = (int)n;
    pthread_t th[256];
    po_par_task task[256];
    int started = 0;
    for (int t = 0; t < nt; t++) {
        task[t].fn = fn;
        task[t].arg = arg;
        task[t].lo = n * t / nt;
        task[t].hi = n * (t + 1) / nt;
        if (t == nt - 1 || pthread_create(&th[t], NULL, po_par_thread, &task[t]) != 0) {
            /* last share (or a thread that could not be created) runs on the caller */
            fn(arg, task[t].lo, t == nt - 1 ? task[t].hi : n);
            break;
        }
        started++;
    }
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
}

po_batch *po_batch_create(int64_t n, const uint8_t *cfg, const uint32_t *tape, int64_t tape_len) {
    po_batch *b = (po_batch *)calloc(1, sizeof(po_batch));
    b->n = n;
    b->envs = (po_gym *)calloc((size_t)n, sizeof(po_gym));
    b->ctx = (struct tape_ctx *)calloc((size_t)n, sizeof(struct tape_ctx));
    b->tape = tape;
    b->tape_len = tape_len;
    for (int64_t i = 0; i < n; i++) {
        uint8_t f = cfg ? cfg[i] : 0;
        po_gym_new_with_config(&b->envs[i], f & 1, (f >> 1) & 1, (f >> 2) & 1, (f >> 3) & 1);
        b->ctx[i].base = tape ? tape + i * tape_len : NULL;
        b->ctx[i].len = tape_len;
        po_gym_set_draws(&b->envs[i], tape_draw, &b->ctx[i]);
    }
    return b;
}
/* switch every env to counter draws keyed (seed, gid_base + i): draw number c of env i becomes
 * pb_draw_raw(seed, gid_base + i, c), exactly what the CUDA engine uses without a tape */
void po_batch_set_counter_rng(po_batch *b, uint64_t seed, int64_t gid_base) {
    for (int64_t i = 0; i < b->n; i++) {
        b->ctx[i].key = po_env_key(seed, (uint64_t)(gid_base + i));
        po_gym_set_draws(&b->envs[i], counter_key_draw, &b->ctx[i]);
    }
}
void po_batch_destroy(po_batch *b) {
    if (!b) return;
    free(b->envs);
    free(b->ctx);
    free(b);
}
po_gym *po_batch_env(po_batch *b, int64_t i) { return &b->envs[i]; }

typedef struct {
    po_batch *b;
    const uint8_t *in8;   /* reset mask / actions */
    uint8_t *out8;        /* done / mask / sampled actions */
    int32_t *reward;
    float *obs;
    po_flat_state *flat_out;
    const po_flat_state *flat_in;
    int64_t first;        /* obs_range */
    int auto_reset;
    uint64_t seed, call_idx;
    int64_t gid_base;
    int64_t first_bad;    /* per call: min over shares, merged under `lock` */
    pthread_mutex_t lock;
} po_call;

static void r_reset(void *v, int64_t lo, int64_t hi) {
    po_call *c = (po_call *)v;
    for (int64_t i = lo; i < hi; i++)
        if (!c->in8 || c->in8[i]) po_gym_reset(&c->b->envs[i]);
}
void po_batch_reset(po_batch *b, const uint8_t *mask) {
    po_call c = {.b = b, .in8 = mask};
    po_parallel_for(b->n, PO_PAR_MIN, r_reset, &c);
}
static void r_step(void *v, int64_t lo, int64_t hi) {
    po_call *c = (po_call *)v;
    int64_t bad = -1;
    for (int64_t i = lo; i < hi; i++) {
        int32_t r = 0;
        uint8_t d = 0;
        if (po_gym_step(&c->b->envs[i], c->in8[i], &r, &d) != 0) {
            if (bad < 0) bad = i;
            c->reward[i] = 0;
            c->out8[i] = 0;
            continue;
        }
        c->reward[i] = r;
        c->out8[i] = d;
        if (d && c->auto_reset) po_gym_reset(&c->b->envs[i]);
    }
    if (bad >= 0) {
        pthread_mutex_lock(&c->lock);
        if (c->first_bad < 0 || bad < c->first_bad) c->first_bad = bad;
        pthread_mutex_unlock(&c->lock);
    }
}
int64_t po_batch_step(po_batch *b, const uint8_t *actions, int32_t *reward, uint8_t *done,
                      int auto_reset) {
    po_call c = {.b = b, .in8 = actions, .out8 = done, .reward = reward, .auto_reset = auto_reset,
                 .first_bad = -1};
    pthread_mutex_init(&c.lock, NULL);
    po_parallel_for(b->n, PO_PAR_MIN, r_step, &c);
    pthread_mutex_destroy(&c.lock);
    return c.first_bad;
}
static void r_mask(void *v, int64_t lo, int64_t hi) {
    po_call *c = (po_call *)v;
    for (int64_t i = lo; i < hi; i++) po_gym_action_mask(&c->b->envs[i], c->out8 + 5 * i);
}
void po_batch_action_mask(po_batch *b, uint8_t *out) {
    po_call c = {.b = b, .out8 = out};
    po_parallel_for(b->n, PO_PAR_MIN, r_mask, &c);
}
static void r_obs(void *v, int64_t lo, int64_t hi) {
    po_call *c = (po_call *)v;
    for (int64_t i = lo; i < hi; i++)
        po_gym_obs(&c->b->envs[c->first + i], c->obs + (size_t)PO_OBS_FLOATS * (size_t)i);
}
void po_batch_obs(po_batch *b, float *out) {
    po_call c = {.b = b, .obs = out, .first = 0};
    po_parallel_for(b->n, 256, r_obs, &c);
}
/* observations of envs [first, first + count) only (large batches are compared chunk by chunk) */
void po_batch_obs_range(po_batch *b, int64_t first, int64_t count, float *out) {
    po_call c = {.b = b, .obs = out, .first = first};
    po_parallel_for(count, 256, r_obs, &c);
}
static void r_export(void *v, int64_t lo, int64_t hi) {
    po_call *c = (po_call *)v;
    for (int64_t i = lo; i < hi; i++) po_gym_export(&c->b->envs[i], &c->flat_out[i]);
}
void po_batch_export(po_batch *b, po_flat_state *out) {
    po_call c = {.b = b, .flat_out = out};
    po_parallel_for(b->n, PO_PAR_MIN, r_export, &c);
}
static void r_import(void *v, int64_t lo, int64_t hi) {
    po_call *c = (po_call *)v;
    for (int64_t i = lo; i < hi; i++) po_gym_import(&c->b->envs[i], &c->flat_in[i]);
}
void po_batch_import(po_batch *b, const po_flat_state *in) {
    po_call c = {.b = b, .flat_in = in};
    po_parallel_for(b->n, PO_PAR_MIN, r_import, &c);
}
/* uniformly random VALID action per env from the policy stream of the counter RNG (the product's
 * pb_sample_actions / pb_step_random): k-th set bit of the mask, k = (raw * popcount) >> 32 */
static void r_sample(void *v, int64_t lo, int64_t hi) {
    po_call *c = (po_call *)v;
    for (int64_t i = lo; i < hi; i++) {
        uint8_t m[5], valid[5];
        int nv = 0;
        po_gym_action_mask(&c->b->envs[i], m);
        for (int k = 0; k < 5; k++)
            if (m[k]) valid[nv++] = (uint8_t)k;
        const uint32_t raw = po_counter_action_raw(c->seed, (uint64_t)(c->gid_base + i), c->call_idx);
        c->out8[i] = valid[((uint64_t)raw * (uint64_t)nv) >> 32];
    }
}
void po_batch_sample_actions(po_batch *b, uint64_t seed, int64_t gid_base, uint64_t call_idx,
                             uint8_t *out) {
    po_call c = {.b = b, .out8 = out, .seed = seed, .gid_base = gid_base, .call_idx = call_idx};
    po_parallel_for(b->n, PO_PAR_MIN, r_sample, &c);
}

/* ======================================================================================== */
/* CPU baseline rollout                                                                     */
/* ======================================================================================== */
typedef struct {
    uint64_t s;
} splitmix;
static uint32_t sm_next(splitmix *r) {
    uint64_t z = (r->s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return (uint32_t)((z ^ (z >> 31)) >> 32);
}
static uint32_t sm_draw(void *ctx, uint32_t ctr) {
    (void)ctr;
    return sm_next((splitmix *)ctx);
}

typedef struct {
    int64_t env_lo, env_hi, n_envs_total, n_steps, warmup;
    uint64_t seed;
    uint64_t checksum;
    int64_t episodes;
    pthread_barrier_t *bar;
    double t0, t1;
} rollout_arg;

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void *rollout_thread(void *va) {
    rollout_arg *a = (rollout_arg *)va;
    int64_t n = a->env_hi - a->env_lo;
    po_gym *envs = (po_gym *)calloc((size_t)(n > 0 ? n : 1), sizeof(po_gym));
    splitmix *rngs = (splitmix *)calloc((size_t)(n > 0 ? n : 1), sizeof(splitmix));
    /* one obs slot per env of the shard: every env-step writes a full 59 024 B observation, as
     * obs_numpy() does in the reference (fresh zeroed array per call, env.rs:412) */
    float *obs = (float *)malloc(sizeof(float) * PO_OBS_FLOATS * (size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < n; i++) {
        int64_t gid = a->env_lo + i;
        rngs[i].s = a->seed * 0x2545f4914f6cdd1dull + (uint64_t)gid * 0x9e3779b97f4a7c15ull;
        po_gym_new(&envs[i], gid < a->n_envs_total / 2, 1); /* replay_buffer.py:75-80 */
        po_gym_set_draws(&envs[i], sm_draw, &rngs[i]);
        po_gym_reset(&envs[i]);
    }
    uint64_t cs = 0;
    int64_t episodes = 0;
    for (int64_t s = -a->warmup; s < a->n_steps; s++) {
        if (s == 0) {
            pthread_barrier_wait(a->bar);
            a->t0 = now_s();
        }
        for (int64_t i = 0; i < n; i++) {
            uint8_t m[5], valid[5];
            int nv = 0;
            po_gym_action_mask(&envs[i], m);
            for (int k = 0; k < 5; k++)
                if (m[k]) valid[nv++] = (uint8_t)k;
            uint8_t act = valid[((uint64_t)sm_next(&rngs[i]) * (uint64_t)nv) >> 32];
            int32_t r;
            uint8_t d;
            po_gym_step(&envs[i], act, &r, &d);
            if (d) {
                po_gym_reset(&envs[i]);
                episodes++;
            }
            float *o = obs + (size_t)PO_OBS_FLOATS * (size_t)i;
            po_gym_obs(&envs[i], o);
            cs += (uint64_t)(uint32_t)r + (uint64_t)d;
            uint32_t bits;
            memcpy(&bits, &o[1 * 868 + envs[i].gs.pacman_loc.col % 28 * 31], 4);
            cs = cs * 31 + bits;
        }
    }
    a->t1 = now_s();
    pthread_barrier_wait(a->bar);
    a->checksum = cs;
    a->episodes = episodes;
    free(envs);
    free(rngs);
    free(obs);
    return NULL;
}

double po_rollout(int64_t n_envs, int64_t n_steps, int64_t warmup_steps, uint64_t seed,
                  int n_threads, uint64_t *checksum, int64_t *episodes) {
    if (n_threads < 1) n_threads = 1;
    if ((int64_t)n_threads > n_envs) n_threads = (int)n_envs;
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    rollout_arg *args = (rollout_arg *)calloc((size_t)n_threads, sizeof(rollout_arg));
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, NULL, (unsigned)n_threads);
    for (int t = 0; t < n_threads; t++) {
        args[t].env_lo = n_envs * t / n_threads;
        args[t].env_hi = n_envs * (t + 1) / n_threads;
        args[t].n_envs_total = n_envs;
        args[t].n_steps = n_steps;
        args[t].warmup = warmup_steps;
        args[t].seed = seed;
        args[t].bar = &bar;
        pthread_create(&th[t], NULL, rollout_thread, &args[t]);
    }
    double t0 = 1e300, t1 = 0;
    uint64_t cs = 0;
    int64_t ep = 0;
    for (int t = 0; t < n_threads; t++) {
        pthread_join(th[t], NULL);
        if (args[t].t0 < t0) t0 = args[t].t0;
        if (args[t].t1 > t1) t1 = args[t].t1;
        cs ^= args[t].checksum + (uint64_t)t;
        ep += args[t].episodes;
    }
    pthread_barrier_destroy(&bar);
    if (checksum) *checksum = cs;
    if (episodes) *episodes = ep;
    free(th);
    free(args);
    return t1 - t0;
}
