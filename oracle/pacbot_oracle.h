/*
 * pacbot_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Scalar, one-env-per-object C restatement of the reference hot path
 *   - wrapper  : /root/reference/pacbot_rs/src/env.rs (whole file; authoritative, in tree)
 *   - engine   : crate `pacbot-rs` 0.1.0 (git RIT-MDRC/pacbot-rs, rev
 *                423622afedc0b0c33c42497cb4e1f2b73641e747; pinned at
 *                /root/reference/pacbot_rs/Cargo.toml:33, Cargo.lock:636-645), imported by the
 *                reference as `pacbot_rs_2`.  NOT vendored under /root/reference, no network, no Rust
 *                toolchain => restated here from the published algorithm (Harvard Pacbot-2 game
 *                engine rules) and anchored on the reference's call sites.
 *
 * PARITY UNPINNED at the engine boundary: the reference ships no tests / golden vectors and the
 * crate cannot be built here.  What IS pinned by in-tree data (walls, action mask, obs layout,
 * reward algebra, reset algebra, constants) is checked in tests/test_oracle_kat.py against
 * tests/golden/ (generated from /root/reference/computed_data/node_coords.json).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (pacbot-rl_b200/) never links, imports or calls it.
 *
 * Randomness: every random draw is INJECTED.  A draw is a raw u32 obtained from
 * `po_draw_fn(ctx, ctr)` (ctr = per-env draw counter, incremented per draw) and mapped to
 * k in [0, n) by k = (raw * n) >> 32.  The reference uses an unseedable thread_rng
 * (env.rs:146) so bitwise equality with a live pacbot_rs run is only possible on RNG-free
 * prefixes anyway.
 */
#ifndef PACBOT_ORACLE_H
#define PACBOT_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PO_ROWS 31
#define PO_COLS 28
#define PO_NUM_NODES 288
#define PO_OBS_CH 17
#define PO_OBS_FLOATS (17 * 28 * 31) /* env.rs:23 OBS_SHAPE */

/* pacbot_rs_2::location::Direction numeric order (candle_inference.rs:78) */
enum { PO_UP = 0, PO_LEFT = 1, PO_DOWN = 2, PO_RIGHT = 3, PO_NONE = 4 };
/* env.rs:40-48 Action */
enum { PO_ACT_STAY = 0, PO_ACT_DOWN = 1, PO_ACT_UP = 2, PO_ACT_LEFT = 3, PO_ACT_RIGHT = 4 };
/* game modes (numbering of the Go engine; internal only) */
enum { PO_MODE_SCATTER = 1, PO_MODE_CHASE = 2 };
/* ghost order env.rs:322 */
enum { PO_RED = 0, PO_PINK = 1, PO_CYAN = 2, PO_ORANGE = 3 };

typedef struct {
    int8_t row, col;
    uint8_t dir;
} po_loc;

typedef struct {
    po_loc loc, next_loc;
    uint8_t color;
    uint8_t fright_steps;
    uint8_t trapped_steps;
    uint8_t spawning;
    uint8_t eaten;
} po_ghost;

typedef struct {
    uint32_t curr_ticks;
    uint8_t update_period;
    uint8_t mode;
    uint8_t paused;
    uint8_t pause_on_update;
    uint8_t mode_steps;
    uint16_t level_steps;
    uint16_t curr_score;
    uint8_t curr_level;
    uint8_t curr_lives;
    uint8_t ghost_combo;
    po_loc pacman_loc;
    po_loc fruit_loc;
    uint8_t fruit_steps;
    po_ghost ghosts[4];
    uint32_t pellets[PO_ROWS];
    uint16_t num_pellets;
} po_game_state;

typedef uint32_t (*po_draw_fn)(void *ctx, uint32_t ctr);

typedef struct {
    po_game_state gs;     /* env.rs:99  */
    uint8_t random_start; /* env.rs:101 */
    uint16_t last_score;  /* env.rs:102 */
    uint8_t last_action;  /* env.rs:103 */
    uint32_t ticks_per_step;      /* env.rs:105 */
    uint8_t random_ticks;         /* env.rs:107 */
    uint8_t randomize_ghosts;     /* env.rs:108 */
    uint8_t remove_super_pellets; /* env.rs:109 */
    /* injected randomness */
    po_draw_fn draw;
    void *draw_ctx;
    uint32_t rng_ctr;
} po_gym;

/* ---- maze tables ------------------------------------------------------------------------ */
extern const uint32_t PO_WALLS[PO_ROWS];        /* complement of node_coords.json */
extern const uint32_t PO_INIT_PELLETS[PO_ROWS]; /* SURVEY D.2 (hypothesis K9, 244 pellets)    */
extern const int8_t PO_NODE_COORDS[PO_NUM_NODES][2]; /* computed_data/node_coords.json        */

/* ---- engine (pacbot_rs_2 restatement) ----------------------------------------------------- */
void po_game_new(po_game_state *g);
int po_wall_at(int row, int col);
int po_pellet_at(const po_game_state *g, int row, int col);
void po_game_step(po_game_state *g, po_gym *rng_owner);
void po_game_set_pacman_location(po_game_state *g, int8_t row, int8_t col);
void po_game_decrement_num_pellets(po_game_state *g);

/* ---- wrapper (env.rs restatement) ---------------------------------------------------------- */
void po_gym_new(po_gym *e, int random_start, int random_ticks);
void po_gym_new_with_config(po_gym *e, int random_start, int random_ticks, int randomize_ghosts,
                            int remove_super_pellets);
void po_gym_set_draws(po_gym *e, po_draw_fn fn, void *ctx);
void po_gym_reset(po_gym *e);
/* returns 0, or -1 for an invalid action (env.rs:83-88 "Invalid action"; state untouched) */
int po_gym_step(po_gym *e, uint8_t action, int32_t *reward, uint8_t *done);
uint32_t po_gym_score(const po_gym *e);
uint8_t po_gym_lives(const po_gym *e);
int po_gym_is_done(const po_gym *e);
int po_gym_first_ai_done(const po_gym *e);
uint16_t po_gym_remaining_pellets(const po_gym *e);
void po_gym_action_mask(const po_gym *e, uint8_t out[5]);
void po_gym_obs(const po_gym *e, float *out /* PO_OBS_FLOATS */);

/* ---- flat state export (for comparing with the product's pb_env_state in tests) ----------- */
typedef struct {
    uint32_t curr_ticks;
    uint32_t rng_ctr;
    uint32_t pellets[PO_ROWS];
    uint16_t level_steps;
    uint16_t curr_score;
    uint16_t num_pellets;
    uint16_t last_score;
    uint8_t update_period;
    uint8_t mode;
    uint8_t paused;
    uint8_t pause_on_update;
    uint8_t mode_steps;
    uint8_t curr_level;
    uint8_t curr_lives;
    uint8_t ghost_combo;
    uint8_t fruit_steps;
    uint8_t last_action;
    uint8_t ticks_per_step;
    uint8_t random_start;
    uint8_t random_ticks;
    uint8_t randomize_ghosts;
    uint8_t remove_super_pellets;
    int8_t pac_row, pac_col;
    uint8_t pac_dir;
    int8_t g_row[4], g_col[4];
    uint8_t g_dir[4];
    int8_t g_next_row[4], g_next_col[4];
    uint8_t g_next_dir[4];
    uint8_t g_fright[4], g_trapped[4], g_spawning[4], g_eaten[4];
} po_flat_state;

void po_gym_export(const po_gym *e, po_flat_state *out);
void po_gym_import(po_gym *e, const po_flat_state *in);
size_t po_sizeof_gym(void);
size_t po_sizeof_flat(void);

/* ---- batch helpers with a draw tape (tests) ------------------------------------------------ */
/* A tape is uint32 raw draws, env-major: tape[env * tape_len + (ctr % tape_len)]. */
typedef struct po_batch po_batch;
po_batch *po_batch_create(int64_t n, const uint8_t *cfg_flags /* per env, bit0 random_start,
                          bit1 random_ticks, bit2 randomize_ghosts, bit3 remove_super_pellets */,
                          const uint32_t *tape, int64_t tape_len);
void po_batch_destroy(po_batch *b);
po_gym *po_batch_env(po_batch *b, int64_t i);
void po_batch_reset(po_batch *b, const uint8_t *mask /* nullable = all */);
/* returns index of first invalid action or -1; auto_reset: reset envs that finished */
int64_t po_batch_step(po_batch *b, const uint8_t *actions, int32_t *reward, uint8_t *done,
                      int auto_reset);
void po_batch_action_mask(po_batch *b, uint8_t *out /* n*5 */);
void po_batch_obs(po_batch *b, float *out /* n*PO_OBS_FLOATS */);
void po_batch_export(po_batch *b, po_flat_state *out /* n */);
void po_batch_import(po_batch *b, const po_flat_state *in /* n */);
void po_batch_obs_range(po_batch *b, int64_t first, int64_t count, float *out);
/* The product's counter RNG (include/pacbot_b200.h), restated: game-draw stream and policy stream.
 * po_batch_set_counter_rng replaces the tape by draws keyed (seed, gid_base + i);
 * po_batch_sample_actions is pb_sample_actions / pb_step_random's action choice. */
uint32_t po_counter_draw_raw(uint64_t seed, uint64_t gid, uint32_t ctr);
uint32_t po_counter_action_raw(uint64_t seed, uint64_t gid, uint64_t call);
void po_batch_set_counter_rng(po_batch *b, uint64_t seed, int64_t gid_base);
void po_batch_sample_actions(po_batch *b, uint64_t seed, int64_t gid_base, uint64_t call_idx,
                             uint8_t *out /* n */);

/* ---- CPU baseline rollout (bench.py cpu_baseline / --impl reference) ----------------------- */
/* Random-valid-action rollout of n_envs for n_steps env-steps each, step + full obs encode per
 * env-step, auto reset on done, random_start for the first n_envs/2 envs, random_ticks for all
 * (mirrors /root/reference/src/replay_buffer.py:75-80).  Sharded over n_threads pthreads.
 * Returns elapsed seconds of the timed region; *checksum gets a digest so the work cannot be
 * optimised away; *episodes the number of finished episodes. */
double po_rollout(int64_t n_envs, int64_t n_steps, int64_t warmup_steps, uint64_t seed,
                  int n_threads, uint64_t *checksum, int64_t *episodes);

#ifdef __cplusplus
}
#endif
#endif
