/*
 * pacbot_b200.h -- C ABI of the B200-native batched PacBot environment engine.
 *
 * Drop-in boundary for the one hot path of qxzcode/pacbot-rl: the pyo3 class `PacmanGym`
 * (reference: pacbot_rs/src/lib.rs:17-25, pacbot_rs/src/env.rs:130-359, stub
 * pacbot_rs/pacbot_rs.pyi:6-19).  Every entry point below names the reference method it
 * replaces.  Plain pointers and sizes only; no torch / C++ types cross this boundary.
 *
 * Conventions
 *   - `*_dev` pointers are device pointers on the engine's device, `*_host` are host pointers.
 *   - `stream` is a `cudaStream_t` passed as `void*` (NULL = legacy default stream).  Calls that
 *     take a stream are asynchronous on it unless stated otherwise.
 *   - All functions return PB_OK (0) or a negative PB_ERR_* code and never throw.
 *     `pb_last_error()` returns a thread-local human-readable message for the last failure.
 *   - A handle is not thread-safe (the reference holds the GIL / a PyCell borrow for every call).
 *   - Per-env config flags: bit0 random_start, bit1 random_ticks (the two ctor args,
 *     env.rs:133), bit2 randomize_ghosts, bit3 remove_super_pellets (Rust-only config,
 *     env.rs:14-21).
 *   - Randomness is counter based: draw number `ctr` of env `g` (global env id = env_id_base + i)
 *     is raw = pb_draw_raw(seed, g, ctr) and is mapped to [0, n) by (raw * n) >> 32.  The
 *     per-env counter lives in the state (`rng_ctr`).  With a draw tape installed
 *     (pb_set_draw_tape) raw = tape[i * tape_len + ctr % tape_len] instead: this is how parity
 *     tests replay identical injected draws on the oracle and on the GPU.
 */
#ifndef PACBOT_B200_H
#define PACBOT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB_VERSION 100 /* 0.1.0 */

#define PB_OK 0
#define PB_ERR_INVALID_ARG (-1)
#define PB_ERR_CUDA (-2)
#define PB_ERR_INVALID_ACTION (-3) /* reference: PyValueError("Invalid action"), env.rs:83-88 */
#define PB_ERR_NO_DEVICE (-4)

#define PB_ROWS 31
#define PB_COLS 28
#define PB_NUM_ACTIONS 5
#define PB_OBS_CHANNELS 17
#define PB_OBS_FLOATS (17 * 28 * 31) /* env.rs:23 OBS_SHAPE = (17, 28, 31); 59 024 bytes */

#define PB_CFG_RANDOM_START 1u
#define PB_CFG_RANDOM_TICKS 2u
#define PB_CFG_RANDOMIZE_GHOSTS 4u
#define PB_CFG_REMOVE_SUPER_PELLETS 8u

/* pb_query selectors -- the scalar getters of PacmanGym */
#define PB_Q_SCORE 0             /* score()             env.rs:269-271 */
#define PB_Q_LIVES 1             /* lives()             env.rs:273-275 */
#define PB_Q_IS_DONE 2           /* is_done()           env.rs:277-279 */
#define PB_Q_FIRST_AI_DONE 3     /* first_ai_done()     env.rs:281-285 */
#define PB_Q_REMAINING_PELLETS 4 /* remaining_pellets() env.rs:287-289 */
#define PB_Q_TICKS_PER_STEP 5
#define PB_Q_CURR_TICKS 6
#define PB_Q_LEVEL 7
#define PB_Q_UPDATE_PERIOD 8
#define PB_Q_COUNT 9

typedef struct pb_engine pb_engine; /* opaque: owns the device-resident SoA state of n envs */

/* Host-side view of ONE env: the reference's `PacmanGym` fields (env.rs:96-110) plus the engine
 * `GameState` fields its wrapper touches.  Directions: 0 up, 1 left,
 * 2 down, 3 right, 4 none.  Modes: 1 scatter, 2 chase.  (32, 32) is the empty location. */
typedef struct pb_env_state {
    uint32_t curr_ticks;
    uint32_t rng_ctr;
    uint32_t pellets[PB_ROWS]; /* bit `col` of pellets[row], as GameState.pellets (env.rs:199) */
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
    uint8_t cfg_flags;
    int8_t pac_row, pac_col;
    uint8_t pac_dir;
    uint8_t reserved0;
    int8_t g_row[4], g_col[4];
    uint8_t g_dir[4];
    int8_t g_next_row[4], g_next_col[4];
    uint8_t g_next_dir[4];
    uint8_t g_fright[4], g_trapped[4], g_spawning[4], g_eaten[4];
} pb_env_state;

/* ---- library -------------------------------------------------------------------------------- */
int pb_version(void);
const char *pb_last_error(void);
size_t pb_sizeof_env_state(void);
/* packed device bytes per env (the S of SURVEY.md 8(d)): what one step reads and writes */
int pb_state_bytes_per_env(void);
/* raw draw of the counter RNG (host mirror of the device function; for tests and docs) */
uint32_t pb_draw_raw(uint64_t seed, uint64_t env_gid, uint32_t ctr);
/* raw draw used by pb_sample_actions */
uint32_t pb_action_raw(uint64_t seed, uint64_t env_gid, uint64_t call_idx);

/* maze tables the kernels use (host copies; derived from computed_data/node_coords.json):
 * wall / initial-pellet row bitboards (bit `col` of row `row`), and NODE_COORDS[k]
 * (pacbot_rs/src/grid.rs:13).  Return 0 / PB_ERR_INVALID_ARG. */
uint32_t pb_maze_wall_row(int row);
uint32_t pb_maze_pellet_row(int row);
int pb_maze_node(int k, int *row, int *col);

/* ---- lifetime --------------------------------------------------------------------------------
 * pb_create replaces PacmanGym::new(random_start, random_ticks) (env.rs:133-138) for n envs:
 * fresh unpaused game, ticks_per_step = 8, NOT randomised (call pb_reset for that).
 * cfg_flags_host: n bytes or NULL (all zero). */
int pb_create(pb_engine **out, int64_t n_envs, int device, uint64_t seed, int64_t env_id_base,
              const uint8_t *cfg_flags_host);
int pb_destroy(pb_engine *e);
int64_t pb_num_envs(const pb_engine *e);
int pb_device(const pb_engine *e);
/* `random_start` property setter (env.rs:100-101) and friends; synchronous */
int pb_set_cfg_flags(pb_engine *e, int64_t first, int64_t count, const uint8_t *cfg_flags_host);
/* tape_dev: uint32 [n_envs][tape_len] on the device, borrowed; NULL restores the counter RNG */
int pb_set_draw_tape(pb_engine *e, const uint32_t *tape_dev, int64_t tape_len);

/* ---- the hot path ---------------------------------------------------------------------------- */
/* PacmanGym::reset (env.rs:140-212) for every env whose mask byte is non-zero (NULL = all). */
int pb_reset(pb_engine *e, const uint8_t *mask_dev, void *stream);

/* PacmanGym::step (env.rs:216-267) for all envs.
 *   actions_dev  uint8 [n]  Action encoding env.rs:40-48 (0 stay, 1 down, 2 up, 3 left, 4 right)
 *   reward_dev   int32 [n]
 *   done_dev     uint8 [n]
 *   mask_out_dev uint8 [n][5] or NULL: action_mask() (env.rs:293-302) of the state AFTER the step
 *                (after the auto reset, if one happened)
 *   auto_reset   non-zero: an env that finishes is reset in the same launch (reward/done still
 *                describe the finished transition), so rollouts never leave the device.
 * An action > 4 leaves that env untouched (reward 0, done 0) and records the lowest offending
 * env index; see pb_check_actions.  Stepping an env that is already done is legal and does what
 * the reference does (the engine keeps going). */
int pb_step(pb_engine *e, const uint8_t *actions_dev, int32_t *reward_dev, uint8_t *done_dev,
            uint8_t *mask_out_dev, int auto_reset, void *stream);

/* PacmanGym::obs / obs_numpy (env.rs:305-307, 410-500) for all envs: float32 [17][28][31] per env
 * written at out_dev + i * env_stride_floats (stride >= PB_OBS_FLOATS and a multiple of 4 floats;
 * 0 means dense).  out_dev must be 16-byte aligned. */
int pb_encode_obs(pb_engine *e, float *out_dev, int64_t env_stride_floats, void *stream);
/* obs() of envs [first, first+count) in the layout the networks' first convolution takes:
 * channels-last, zero-padded to 32 channels -- float32 [count][28][31][32] (a torch tensor of
 * shape [count, 32, 28, 31] in channels_last memory format), channels 17..31 zero.  Same values
 * as pb_encode_obs followed by a layout change (src/models.py:70-77 consumes it), one launch, one
 * CTA per env.  out_dev must be 16-byte aligned. */
int pb_encode_obs_nhwc32(pb_engine *e, int64_t first, int64_t count, float *out_dev, void *stream);

/* Same for envs [first, first + count) only; env first + k is written at out_dev + k * stride.
 * Lets a caller stream a large env population through a smaller observation ring. */
int pb_encode_obs_range(pb_engine *e, int64_t first, int64_t count, float *out_dev,
                        int64_t env_stride_floats, void *stream);

/* Observation of all envs into slot *slot_index_dev of a ring: env i is written at
 * ring_dev + (*slot_index_dev) * slot_stride_floats + i * PB_OBS_FLOATS.  The slot number is
 * read on the DEVICE at kernel time, so a captured CUDA graph keeps working while the replay
 * ring advances (src/replay_buffer.py's deque append, done in place). */
int pb_encode_obs_ring(pb_engine *e, float *ring_dev, int64_t slot_stride_floats,
                       const int64_t *slot_index_dev, void *stream);
/* pb_encode_obs_ring that ALSO hands the same observations out in the networks' input layout
 * (channels-last, zero-padded to 32 channels: nhwc_out_dev float32 [n][28][31][32], see
 * pb_encode_obs_nhwc32), from the same launch (one CTA per env): the replay ring of
 * src/replay_buffer.py keeps the channel-major observation, the policy's next forward
 * (src/replay_buffer.py:107-112) reads its input without a layout conversion in between. */
int pb_encode_obs_ring_nhwc32(pb_engine *e, float *ring_dev, int64_t slot_stride_floats,
                              const int64_t *slot_index_dev, float *nhwc_out_dev, void *stream);

/* pb_step followed by pb_encode_obs of the resulting states, back to back on `stream`. */
int pb_step_encode(pb_engine *e, const uint8_t *actions_dev, int32_t *reward_dev,
                   uint8_t *done_dev, uint8_t *mask_out_dev, int auto_reset, float *obs_out_dev,
                   int64_t env_stride_floats, void *stream);
/* pb_step + the observation of the resulting states for a SMALL engine (the DQN loop's 32 envs, an
 * evaluation env), one launch, one CTA per env: in pb_step the envs of a small batch share one
 * warp and their divergent game logic is serialised; here every env steps on a warp of its own
 * and its CTA then encodes the observation -- channel-major at obs_out_dev + i * env_stride_floats
 * (+ *slot_index_dev * slot_stride_floats when slot_index_dev is given: the replay ring's slot,
 * read at kernel time) and / or channels-last padded to 32 channels at nhwc_out_dev (see
 * pb_encode_obs_nhwc32); either output may be NULL.  Same results as pb_step + pb_encode_obs*. */
int pb_step_encode_small(pb_engine *e, const uint8_t *actions_dev, int32_t *reward_dev, uint8_t *done_dev,
                         uint8_t *mask_out_dev, int auto_reset, float *obs_out_dev, int64_t env_stride_floats,
                         const int64_t *slot_index_dev, int64_t slot_stride_floats, float *nhwc_out_dev,
                         void *stream);

/* PacmanGym::action_mask (env.rs:293-302): uint8 [n][5]. */
int pb_action_mask(pb_engine *e, uint8_t *mask_out_dev, void *stream);

/* Scalar getters (PB_Q_*): int32 [n]. */
int pb_query(pb_engine *e, int what, int32_t *out_dev, void *stream);

/* Uniformly random VALID action per env from its current action mask (what EpsilonGreedy does
 * at epsilon = 1, /root/reference/src/policies.py:29-32), drawn from the counter RNG's action
 * stream keyed by (seed, env id, call_idx). */
int pb_sample_actions(pb_engine *e, uint8_t *actions_dev, uint64_t call_idx, void *stream);

/* pb_sample_actions(call_idx) + pb_step in ONE launch (random-policy rollouts: the policy is
 * EpsilonGreedy at epsilon = 1).  The action each env took is written to actions_out_dev
 * (uint8 [n], may be NULL).  Bit-identical to calling the two entry points in sequence. */
int pb_step_random(pb_engine *e, uint64_t call_idx, uint8_t *actions_out_dev, int32_t *reward_dev,
                   uint8_t *done_dev, uint8_t *mask_out_dev, int auto_reset, void *stream);
/* Same for envs [first, first + count) only; every array is still indexed by the absolute env
 * index (pass the same full-size arrays).  Together with pb_encode_obs_range this lets a caller
 * software-pipeline a rollout over two streams: step chunk B of step t+1 while chunk A of step t
 * is still being encoded (disjoint envs, so no hazard) -- see rollout.py. */
int pb_step_random_range(pb_engine *e, int64_t first, int64_t count, uint64_t call_idx,
                         uint8_t *actions_out_dev, int32_t *reward_dev, uint8_t *done_dev,
                         uint8_t *mask_out_dev, int auto_reset, void *stream);
/* pb_step_random with DOUBLE-BUFFERED state: reads the current state buffer, writes the stepped
 * state into the engine's second buffer (allocated on first use, +176 B per env) and makes that
 * the current one.  The buffer just read stays intact, so an observation encode of the previous
 * state that is still running on ANOTHER stream is not disturbed: step t+1 can overlap encode t
 * (rollout.py).  The caller orders the streams: this launch must come after the encode of state
 * t-1 has finished (it overwrites that buffer). */
int pb_step_random_flip(pb_engine *e, uint64_t call_idx, uint8_t *actions_out_dev,
                        int32_t *reward_dev, uint8_t *done_dev, uint8_t *mask_out_dev,
                        int auto_reset, void *stream);
/* The same double-buffered step with caller-supplied actions (pb_step semantics; actions_dev ==
 * NULL draws random valid actions like pb_step_random, then call_idx / actions_out_dev apply). */
int pb_step_flip(pb_engine *e, const uint8_t *actions_dev, uint64_t call_idx,
                 uint8_t *actions_out_dev, int32_t *reward_dev, uint8_t *done_dev,
                 uint8_t *mask_out_dev, int auto_reset, void *stream);

/* ---- pipelined rollout steps (one C call per step) ------------------------------------------- */
/* A rollout step = pb_step_flip + pb_encode_obs on ENGINE-OWNED streams, so that the step
 * kernel of step t+1 runs while the encoder is still streaming step t to HBM (the step only
 * waits for the encode of step t-1, whose state buffer it overwrites).  `stream` is consulted at
 * the first step after pb_create / pb_rollout_join only: the engine's streams start after the
 * work already queued on it.  The observations and the device-side outputs are complete for a
 * consumer on `stream` after pb_rollout_join(e, stream); until then do not touch the engine
 * through other entry points.  Results are bit-identical to pb_step(_random) + pb_encode_obs. */
/* random valid actions, everything stays on the device (pb_step_random semantics) */
int pb_rollout_random_step(pb_engine *e, uint64_t call_idx, uint8_t *actions_out_dev,
                           int32_t *reward_dev, uint8_t *done_dev, uint8_t *mask_out_dev,
                           int auto_reset, float *obs_out_dev, int64_t env_stride_floats,
                           void *stream);
/* host actions in, reward/done/mask back on the host when the call returns (pb_step_host
 * semantics; the observation is complete after pb_rollout_join or after a later step returned) */
int pb_rollout_host_step(pb_engine *e, const uint8_t *actions_host, int32_t *reward_host,
                         uint8_t *done_host, uint8_t *mask_host, int auto_reset, float *obs_out_dev,
                         int64_t env_stride_floats, void *stream);
/* The same overlap for env populations walked chunk by chunk (the 2 M-env shard of a GPU: one
 * encoder launch per ring slot), IN PLACE -- no second state buffer: k_step of envs
 * [first, first + count) runs on the engine's step stream as soon as the encoder that read these
 * envs in the previous sweep has finished, i.e. while the encoders of the other chunks stream to
 * HBM; the chunk's encoder follows on one of two encode streams (alternating).  All arrays are
 * indexed by the ABSOLUTE env index (like pb_step_random_range); obs_out_dev points at the
 * observation of env `first`.  actions_dev == NULL draws random valid actions (pb_step_random
 * semantics, call_idx / actions_out_dev apply), otherwise pb_step semantics with actions produced
 * on `stream` (the step waits for the work queued there so far).  Per-env results are
 * bit-identical to pb_step(_random) + pb_encode_obs_range; complete after pb_rollout_join.
 * Replaces the per-env loop of src/replay_buffer.py:107-137 / the call pair env.rs:216-267 +
 * env.rs:305-307 at shard scale. */
int pb_rollout_chunk_step(pb_engine *e, int64_t first, int64_t count, const uint8_t *actions_dev,
                          uint64_t call_idx, uint8_t *actions_out_dev, int32_t *reward_dev,
                          uint8_t *done_dev, uint8_t *mask_out_dev, int auto_reset,
                          float *obs_out_dev, int64_t env_stride_floats, void *stream);
/* Host buffers (full-size arrays, absolute env index): the chunk's actions go host -> device on
 * the step stream, its reward / done / mask come back on the copy stream while the encoders run.
 * Nothing blocks: pb_rollout_host_sync returns when the results of every chunk issued so far are
 * in host memory (and reports an invalid action, PB_ERR_INVALID_ACTION, like pb_step_host). */
int pb_rollout_chunk_host_step(pb_engine *e, int64_t first, int64_t count,
                               const uint8_t *actions_host, int32_t *reward_host,
                               uint8_t *done_host, uint8_t *mask_host, int auto_reset,
                               float *obs_out_dev, int64_t env_stride_floats, void *stream);
int pb_rollout_host_sync(pb_engine *e);
int pb_rollout_join(pb_engine *e, void *stream);
/* the engine-owned streams (cudaStream_t), e.g. to record timing events on them: the step stream
 * and the encode stream of the MOST RECENT step (encoders alternate between two streams) */
int pb_rollout_streams(pb_engine *e, void **step_stream, void **encode_stream);

/* ---- the per-env drop-in's round trip (engines of up to PB_SNAPSHOT_MAX_ENVS envs) ------------ */
/* What the reference's loop reads per env and step (src/replay_buffer.py:107-137,
 * src/algorithms/train_dqn.py:95-113) in ONE kernel launch and no copy operation (the kernel
 * reads the actions from, and writes every answer to, pinned host memory; the call returns when
 * its last CTA has published the results, with `stream` drained up to it): optionally step every
 * env with `actions_host` (uint8 [n]; NULL = no step; pb_step semantics, reward int32 [n] / done
 * uint8 [n] returned), then info_host int32 [n][8] = score, lives, is_done, first_ai_done,
 * remaining_pellets, ticks_per_step, level, curr_ticks (env.rs:269-289), mask_host uint8 [n][5]
 * (env.rs:293-302) and obs_host float [n][17][28][31] (env.rs:305-307); any output may be NULL.
 * Returns PB_ERR_INVALID_ACTION if an action was > 4 (that env untouched, env.rs:83-88). */
#define PB_SNAPSHOT_MAX_ENVS 1024
int pb_step_snapshot_host(pb_engine *e, const uint8_t *actions_host, int auto_reset,
                          int32_t *reward_host, uint8_t *done_host, int32_t *info_host,
                          uint8_t *mask_host, float *obs_host, void *stream);
/* Zero-copy form: the engine's own pinned buffers.  *block_host receives the block the kernel
 * fills and offsets[0..5] the byte offsets of info, reward, done, mask, obs inside it and its
 * total size; *actions_host the pinned uint8 [n] the kernel reads.  Passing exactly these
 * addresses to pb_step_snapshot_host means "in place": nothing is copied on the host, and the
 * contents are valid until the next pb_step_snapshot_host call of this engine. */
int pb_snapshot_block(pb_engine *e, void **block_host, uint8_t **actions_host, size_t *offsets);

/* Synchronises `stream`, then returns PB_ERR_INVALID_ACTION (and the lowest offending env index
 * in *first_bad_env) if any pb_step since the last check saw an action > 4; clears the record. */
int pb_check_actions(pb_engine *e, int64_t *first_bad_env, void *stream);

/* ---- state access (parity tests, MCTS-style cloning, checkpointing); synchronous ------------- */
int pb_get_state(pb_engine *e, int64_t first, int64_t count, pb_env_state *out_host);
int pb_set_state(pb_engine *e, int64_t first, int64_t count, const pb_env_state *in_host);
/* The same, ordered on `stream`: only that stream and the engine's own streams are synchronised
 * (pb_get_state / pb_set_state wait for the whole device).  This is what the per-env drop-in
 * getters (score(), lives(), ... env.rs:269-289) go through: a training graph replaying on
 * another stream is not stalled by them. */
int pb_get_state_on(pb_engine *e, int64_t first, int64_t count, pb_env_state *out_host, void *stream);
int pb_set_state_on(pb_engine *e, int64_t first, int64_t count, const pb_env_state *in_host,
                    void *stream);
/* #[derive(Clone)] on PacmanGym (env.rs:96): copy env src_i of `src` over env dst_i of `dst` */
int pb_clone_env(pb_engine *dst, int64_t dst_i, const pb_engine *src, int64_t src_i, void *stream);

/* ---- host-buffer entry point (what a per-call FFI binding would use) ------------------------- */
/* Copies actions host->device, runs pb_step (+ pb_encode_obs when obs_out_dev != NULL) on
 * `stream` and copies reward/done/mask device->host.  Returns as soon as those RESULTS are on the
 * host: they only depend on the step kernel and travel on an internal side stream while the
 * observation encode is still running; the observation itself is ordered on `stream` like any
 * other asynchronous call (work queued on `stream` afterwards sees it; a host reader must
 * synchronise `stream`).  Host pointers may be pageable or pinned; mask_host may be NULL.
 * Returns PB_ERR_INVALID_ACTION like pb_check_actions. */
int pb_step_host(pb_engine *e, const uint8_t *actions_host, int32_t *reward_host,
                 uint8_t *done_host, uint8_t *mask_host, int auto_reset, float *obs_out_dev,
                 int64_t env_stride_floats, void *stream);
/* pb_step_host without the observation, split in two for callers that enqueue their own work
 * between a step and the moment they need its results (e.g. the chunked observation encodes of a
 * population larger than its ring, pb_encode_obs_range): _begin enqueues the H2D copy of the
 * actions, the step on `stream` and the D2H copies of reward / done / mask on a side stream, and
 * returns; _end returns when the results are in the host buffers (PB_ERR_INVALID_ACTION as for
 * pb_step_host).  One step in flight per engine; the host buffers must stay valid until _end. */
int pb_step_host_begin(pb_engine *e, const uint8_t *actions_host, int32_t *reward_host,
                       uint8_t *done_host, uint8_t *mask_host, int auto_reset, void *stream);
int pb_step_host_end(pb_engine *e, void *stream);
/* obs_numpy() (env.rs:305-307) of envs [first, first+count) straight to a host buffer.  The
 * range is encoded in chunks of 1 024 envs into two device staging buffers; each chunk crosses
 * PCIe on an internal stream while the next one is being encoded.  A pinned / registered
 * destination receives the copies directly (PCIe-bound); a pageable one goes through two pinned
 * bounce buffers with the host memcpy overlapped.  Returns when obs_host is complete. */
int pb_obs_host(pb_engine *e, int64_t first, int64_t count, float *obs_host, void *stream);

/* ---- on-device DQN helpers (reference: src/policies.py, src/replay_buffer.py) ---------------- */
/* EpsilonGreedy(MaxQPolicy) (policies.py:16-59): with probability eps a uniformly random valid
 * action, else argmax over valid actions of q (ties -> lowest index, like torch.argmax).
 * q_dev float32 [n][5]; mask_dev uint8 [n][5] or NULL (= current action_mask()). */
int pb_select_actions(pb_engine *e, const float *q_dev, const uint8_t *mask_dev, float epsilon,
                      uint8_t *actions_dev, uint64_t call_idx, void *stream);
/* pb_select_actions with epsilon (float32) and the call index (int64) read from DEVICE memory at
 * kernel time: a captured CUDA graph follows the epsilon schedule and the step counter while it
 * is replayed. */
int pb_select_actions_dev(pb_engine *e, const float *q_dev, const uint8_t *mask_dev,
                          const float *epsilon_dev, const int64_t *call_idx_dev,
                          uint8_t *actions_dev, void *stream);
/* Gather `count` observations (float32 [PB_OBS_FLOATS] each) from ring slots idx_dev[k]
 * (int64 slot numbers) of ring_dev (slot stride in floats) into the dense out_dev, writing zeros
 * where zero_mask_dev[k] != 0 (the `next_obs is None` case, train_dqn.py:154-159). */
int pb_gather_obs(const float *ring_dev, int64_t slot_stride_floats, const int64_t *idx_dev,
                  const uint8_t *zero_mask_dev, int64_t count, float *out_dev, int device,
                  void *stream);

/* ---- replay ring (reference: src/replay_buffer.py:57-143, collate of train_dqn.py:152-165) ---
 * A TIME ring of R slots x n envs.  Slot s = t % R holds the observation before action t (written
 * by pb_encode_obs_ring) and, per env, action uint8 / reward int32 / done uint8 / next action mask
 * uint8[5] / valid uint8 of transition t, each a dense [R][n] array owned by the caller; flat index
 * = s * n + env; next_obs of a transition is slot (s + 1) % R.  The step counter t is an int64 in
 * DEVICE memory read at kernel time (CUDA-graph replay). */
/* deque.append for one step of all envs (replay_buffer.py:107-137): slot t % R receives the
 * transition (next mask = mask_dev & ~done; valid = 1, or !in_phase1_dev[i] for the steps of
 * reset_env's old policy, replay_buffer.py:44-54, which are not recorded); slot (t+1) % R becomes
 * "current observation only" (valid = 0).  mask_dev is the mask AFTER the step (pb_step's). */
int pb_replay_record(int64_t n_envs, int64_t ring_slots, const int64_t *t_dev,
                     const uint8_t *actions_dev, const int32_t *reward_dev, const uint8_t *done_dev,
                     const uint8_t *mask_dev, const uint8_t *in_phase1_dev, uint8_t *ring_actions,
                     int32_t *ring_rewards, uint8_t *ring_done, uint8_t *ring_next_mask,
                     uint8_t *ring_valid, int device, void *stream);
/* random.sample(buffer, batch) (replay_buffer.py:141-143): `batch` DISTINCT flat indices of valid
 * transitions, uniform over all such subsets and orders: the valid flags are compacted in
 * ascending order into a list of m entries, then position j = 0..batch-1 swaps with position
 * j + ((raw_j * (m - j)) >> 32), raw_j = pb_replay_raw(seed, *call_idx_dev, j) (partial
 * Fisher-Yates).  *num_valid_out_dev (may be NULL) receives m; if m < batch the draws are WITH
 * replacement over the m entries (index 0 if m == 0) -- check m before training.
 * Limits: batch <= 4 096, total = R * n <= pb_replay_sample_capacity() (the list lives in
 * shared memory). */
/* The ring advances by one step: *t_dev += 1, *next_slot_dev = (*t_dev + 1) % ring_slots (the slot
 * pb_encode_obs_ring writes the next observation to); device-side so that a captured graph keeps
 * working (the deque append of replay_buffer.py:130 moving on). */
int pb_replay_advance(int64_t *t_dev, int64_t *next_slot_dev, int64_t ring_slots, int device, void *stream);
uint32_t pb_replay_raw(uint64_t seed, uint64_t call_idx, uint32_t j);
int64_t pb_replay_sample_capacity(void);
int pb_replay_sample(const uint8_t *ring_valid_dev, int64_t total, int64_t batch, uint64_t seed,
                     const int64_t *call_idx_dev, int64_t *idx_out_dev, int32_t *num_valid_out_dev,
                     int device, void *stream);
/* The collate of train_dqn.py:153-165 for the transitions idx_dev[0..batch): obs and next_obs
 * (zeros where done: `next_obs is None`) as float32 [batch][17][28][31], dense NCHW
 * (channels_last == 0) or in channels-last memory order [batch][28][31][C] (what the NHWC
 * convolutions of the Q-network read without a layout conversion): channels_last == 1 or 17
 * gives C = 17; a multiple of 4 in 20..64 gives that many channels, those >= 17 all zero
 * (cuDNN's tensor-core kernels for the first convolution need C % 8 == 0; the network pads its
 * first weight with zero input channels, which leaves the result unchanged).  actions int64,
 * rewards float32 (raw, unscaled), done uint8, next masks uint8 [batch][5]. */
int pb_replay_collate(const float *ring_obs_dev, int64_t n_envs, int64_t ring_slots,
                      const int64_t *idx_dev, int64_t batch, const uint8_t *ring_actions,
                      const int32_t *ring_rewards, const uint8_t *ring_done,
                      const uint8_t *ring_next_mask, float *obs_out_dev, float *next_obs_out_dev,
                      int channels_last, int64_t *actions_out_dev, float *rewards_out_dev,
                      uint8_t *done_out_dev, uint8_t *next_mask_out_dev, int device, void *stream);
/* last_obs of the policy forward (replay_buffer.py:109): the n observations of slot t % R, dense
 * NCHW or channels-last (channels_last as for pb_replay_collate). */
int pb_replay_slot_obs(const float *ring_obs_dev, int64_t n_envs, int64_t ring_slots,
                       const int64_t *t_dev, float *out_dev, int channels_last, int device,
                       void *stream);

/* evaluate_episode's bookkeeping (src/algorithms/train_dqn.py:99-111) after one step of all envs,
 * one launch: stats_dev is int32 [7][n] with rows 0 score, 1 steps, 2 lives, 3 remaining pellets,
 * 4 finished (done within max_steps), 5 alive, 6 steps left.  An env that is alive with steps left
 * counts the step and re-reads score / lives / pellets from its state AFTER the step; done_dev[i]
 * (pb_step's output for this step) then ends its episode and freezes the row values.  The caller
 * initialises rows 5 (1) and 6 (max_steps) and zeroes the rest. */
int pb_eval_track(pb_engine *e, const uint8_t *done_dev, int32_t *stats_dev, void *stream);

/* ---- Q-network convolution-block tail (reference: src/models.py:70-90) -------------------------
 * `Conv2d -> MaxPool2d(2, ceil_mode=True) -> SiLU` (pool != 0) and `Conv2d -> SiLU` (pool == 0) of
 * QNetV2 minus the convolution (which stays cuDNN, run WITHOUT its bias): bias add, 2x2 max pool
 * and SiLU in one kernel, and their backward in one kernel.  Channels-last float32: x is
 * [n][h][w][c]; z / y are [n][ho][wo][c] with ho = ceil(h / 2), wo = ceil(w / 2) (h, w without
 * pool); c a multiple of 4 with (c / 4) dividing 256.
 * fwd: z = max over the window of (x + bias) (first maximum in row-major window order, NaN wins:
 * torch's rule), y = z / (1 + exp(-z)); idx = arg-max 0..3 per element (pool only).  z_out_dev
 * may be NULL when no backward will follow; without pool it may alias x_dev.
 * bwd: dx = dy * silu'(z) routed to the arg-max cell of each window (zeros elsewhere; every cell
 * of dx is written), dbias[c] = sum of dy * silu'(z) (zeroed inside the call). */
int pb_bias_pool_silu_fwd(const float *x_dev, const float *bias_dev, int64_t n, int64_t h, int64_t w,
                          int64_t c, int pool, float *z_out_dev, float *y_out_dev, uint8_t *idx_out_dev,
                          int device, void *stream);
int pb_bias_pool_silu_bwd(const float *dy_dev, const float *z_dev, const uint8_t *idx_dev, int64_t n,
                          int64_t h, int64_t w, int64_t c, int pool, float *dx_out_dev,
                          float *dbias_out_dev, int device, void *stream);

/* The head of QNetV2 / NetV2 (src/models.py:92-110, :139-141) for inference, one launch: global
 * max over the map of the last convolution's output x (float32 [n][hw][128], channels-last,
 * WITHOUT its bias) + conv_bias, SiLU, Linear 128 -> 256 (w1 [256][128], b1), SiLU, then
 * dueling != 0: Q = V - mean(A) + A with wv [1][256], bv [1], wa [n_out][256], ba [n_out];
 * dueling == 0: one plain linear layer wa / ba (wv / bv unused).  out float32 [n][n_out],
 * n_out <= 7.  float32 arithmetic, summation order differs from cuBLAS (rtol 1e-5). */
int pb_qnet_head_fwd(const float *x_dev, int64_t n, int64_t hw, const float *conv_bias_dev,
                     const float *w1_dev, const float *b1_dev, const float *wv_dev, const float *bv_dev,
                     const float *wa_dev, const float *ba_dev, int64_t n_out, int dueling,
                     float *out_dev, int device, void *stream);

/* pb_qnet_head_fwd + MaxQPolicy (src/policies.py:44-59) in the same launch: actions_out uint8 [n] =
 * argmax over the actions with mask_dev uint8 [n][n_out] != 0 (first maximum; 0 if none), q_out
 * float32 [n][n_out] as above.  One launch per greedy evaluation step instead of four more. */
int pb_qnet_head_greedy(const float *x_dev, int64_t n, int64_t hw, const float *conv_bias_dev,
                        const float *w1_dev, const float *b1_dev, const float *wv_dev, const float *bv_dev,
                        const float *wa_dev, const float *ba_dev, int64_t n_out, int dueling,
                        float *q_out_dev, const uint8_t *mask_dev, uint8_t *actions_out_dev, int device,
                        void *stream);
/* pb_qnet_head_fwd (plain head, n_out = actions + 1 rows) + the AlphaZero evaluator's
 * post-processing (src/algorithms/train_alphazero.py:83-95) in the same launch: value_out float32
 * [n] = last row * value_scale, policy_out float32 [n][n_out - 1] = softmax of the other rows with
 * the actions whose mask_dev uint8 [n][n_out - 1] is 0 at -inf; out float32 [n][n_out] as above. */
int pb_qnet_head_policy_value(const float *x_dev, int64_t n, int64_t hw, const float *conv_bias_dev,
                              const float *w1_dev, const float *b1_dev, const float *wo_dev,
                              const float *bo_dev, int64_t n_out, const uint8_t *mask_dev, float value_scale,
                              float *out_dev, float *value_out_dev, float *policy_out_dev, int device,
                              void *stream);
/* Training form of the head: the same forward that also keeps p_out float32 [n][128] (pooled
 * maximum + bias), idx_out uint8 [n][128] (map position of the maximum, first one in row-major
 * order as torch chooses; hw <= 255) and u_out float32 [n][256] (input of the second SiLU) ... */
int pb_qnet_head_fwd_train(const float *x_dev, int64_t n, int64_t hw, const float *conv_bias_dev,
                           const float *w1_dev, const float *b1_dev, const float *wv_dev,
                           const float *bv_dev, const float *wa_dev, const float *ba_dev, int64_t n_out,
                           int dueling, float *out_dev, float *p_out_dev, uint8_t *idx_out_dev,
                           float *u_out_dev, int device, void *stream);
/* ... and its backward, two launches: dq float32 [n][n_out] in; dx float32 [n][hw][128] (every
 * cell written: the gradient at the arg-max position, 0 elsewhere), the convolution bias gradient
 * [128], dW1 [256][128], db1 [256], dWv [256] / dbv [1] (dueling only), dWa [n_out][256], dba
 * [n_out] out.  scratch_dev: pb_qnet_head_bwd_scratch_floats(n) floats of device memory.  The
 * arithmetic is the chain rule of src/models.py:92-110 in float32; batch sums in sample order. */
int pb_qnet_head_bwd(const float *dq_dev, const float *p_dev, const uint8_t *idx_dev, const float *u_dev,
                     const float *w1_dev, const float *wv_dev, const float *wa_dev, int64_t n, int64_t hw,
                     int64_t n_out, int dueling, float *dx_dev, float *dconv_bias_dev, float *dw1_dev,
                     float *db1_dev, float *dwv_dev, float *dbv_dev, float *dwa_dev, float *dba_dev,
                     float *scratch_dev, int device, void *stream);
int64_t pb_qnet_head_bwd_scratch_floats(int64_t n);

/* ---- the arithmetic of a DQN iteration around the network (src/algorithms/train_dqn.py:167-206) */
/* Double-DQN target (:167-183): target[b] = reward[b] * reward_scale + (done[b] ? 0 : gamma *
 * Q_target(s')[b][argmax_a masked Q_online(s')[b][a]]); masked-out actions count as -inf, ties go
 * to the first maximum (torch.argmax).  q_*_dev float32 [n][num_actions], mask_dev / done_dev
 * uint8 (bool), reward_dev float32 [n]; same float32 operations as the reference expressions. */
int pb_dqn_target(const float *q_target_dev, const float *q_online_dev, const uint8_t *mask_dev,
                  const float *reward_dev, const uint8_t *done_dev, double gamma, double reward_scale,
                  float *target_out_dev, int64_t n, int64_t num_actions, int device, void *stream);
/* Loss (:193-197): out4[0] = mean_b (Q[b][a_b] - target[b])^2; dq_out [n][num_actions] = its
 * gradient with respect to the Q table (2 / n * (Q[b][a_b] - target[b]) at a_b, 0 elsewhere);
 * the logged statistics out4[1] = stat_scale * mean_b max_a Q[b][a], out4[2] = stat_scale *
 * mean(target), out4[3] = 1.0 if the loss is finite (the reference's error_if_nonfinite tripwire,
 * :200-204).  One CTA, sums in a fixed order. */
int pb_dqn_loss(const float *all_q_dev, const int64_t *actions_dev, const float *target_dev, int64_t n,
                int64_t num_actions, double stat_scale, float *out4_dev, float *dq_out_dev, int device,
                void *stream);
/* ---- end of a DQN iteration (src/algorithms/train_dqn.py:200-206) -----------------------------
 * torch.nn.utils.clip_grad_norm_(parameters, max_norm) followed by torch.optim.Adam.step()
 * (default Adam: no weight decay, no amsgrad) over `count` <= 32 float32 tensors, two launches.
 * params_dev / grads_dev: HOST arrays of `count` device pointers; sizes[t] = elements of tensor t;
 * a parameter and its gradient share one memory layout (storage order is what is paired).
 * exp_avg_dev / exp_avg_sq_dev: float32 [sum sizes], the moments in tensor order (zero at start);
 * step_dev: int64 [1] device step counter, advanced by the call (zero at start);
 * partials_dev: float32 [count * 8 + 2] scratch; grad_norm_out_dev: float32 [1], the total norm BEFORE
 * clipping, what clip_grad_norm_ returns (may be NULL).  The gradients are read, not rewritten.
 * The hyper-parameters are doubles, as torch takes them (1 - beta is rounded to float32 once). */
int pb_clip_adam(int count, float *const *params_dev, const float *const *grads_dev,
                 const int64_t *sizes, float *exp_avg_dev, float *exp_avg_sq_dev, int64_t *step_dev,
                 float *partials_dev, double max_norm, double lr, double beta1, double beta2, double eps,
                 float *grad_norm_out_dev, int device, void *stream);



/* ---- device timeline (diagnostics; no counterpart in the reference) --------------------------
 * Between pb_trace_begin and pb_trace_end every k_step / k_encode_obs launch of the engine
 * records, in nanoseconds of the GPU's %globaltimer: begin = first CTA started, ready = first CTA
 * past its wait for the preceding grid (programmatic dependent launch; == begin for k_step),
 * end = last warp finished.  pb_trace_mark puts a one-thread kernel on `stream` that records the
 * time it ran (the boundaries of a timed region in the same clock).  Records come back in launch
 * order; launches beyond `capacity` are not recorded.  pb_trace_begin / pb_trace_end synchronise
 * the device.  bench.py does not trace its timed loops; tools/timeline.py does. */
#define PB_TRACE_MARK 0u
#define PB_TRACE_STEP 1u
#define PB_TRACE_ENCODE 2u
typedef struct pb_trace_record {
    uint64_t begin_ns, ready_ns, end_ns;
    uint32_t kind, reserved;
} pb_trace_record;
int pb_trace_begin(pb_engine *e, int64_t capacity);
int pb_trace_mark(pb_engine *e, void *stream);
int pb_trace_end(pb_engine *e, pb_trace_record *out_host, int64_t max_records, int64_t *count);

/* Batched #[derive(Clone)] (env.rs:96) between engines on one device: for k in [0, count) copy
 * env src_idx_dev[k] of `src` over env dst_idx_dev[k] of `dst` (int64 device arrays; NULL =
 * identity; negative or out-of-range entries are skipped).  Used for MCTS clones and for state
 * snapshots (an experience buffer of 176-byte states instead of 59 KB observations). */
int pb_copy_envs(pb_engine *dst, const int64_t *dst_idx_dev, const pb_engine *src,
                 const int64_t *src_idx_dev, int64_t count, void *stream);

/* ---- batched MCTS (reference: pacbot_rs/src/mcts.rs) --------------------- */
/* One search tree per env of `root`, all trees advanced in lock step.  `sim` is a second engine
 * of the same size on the same device: slot i receives the clone of env i that a simulation
 * steps (select_trajectory's env.clone(), mcts.rs:293), so the leaf observation of tree i is
 * pb_encode_obs(sim) row i.  Both engines are borrowed and must outlive the pb_mcts.
 * max_nodes = LIVE node capacity per tree (2..32767); a search iteration adds at most one node.
 * (Physically each tree has two pools of 8 x max_nodes 64-byte nodes: take_action moves the root
 * index and compacts the kept subtree into the other pool only when the free tail runs short;
 * env PB_MCTS_POOL_FACTOR overrides the 8.)
 * Simulated env steps draw from the counter RNG with seed sim_seed(seed, epoch) =
 * mix64(seed ^ ((epoch + 1) * 0xd6e8feb86659fd93)), same gid and draw counter as the real env;
 * epoch = number of pb_mcts_backprop calls so far.  All float arithmetic is single IEEE f32
 * operations in the reference's order. */
typedef struct pb_mcts pb_mcts;
/* the seed simulations of search iteration `epoch` draw from (host helper, like pb_draw_raw) */
uint64_t pb_sim_seed(uint64_t seed, uint64_t epoch);
int pb_mcts_create(pb_mcts **out, pb_engine *root, pb_engine *sim, int max_nodes);
int pb_mcts_destroy(pb_mcts *m);
int pb_mcts_max_nodes(const pb_mcts *m);
/* reset_root without the noise (mcts.rs:252-255): root := new_visited(value, policy) for every
 * env with mask_dev[i] != 0 (NULL = all); value_dev float32 [n], policy_dev float32 [n][5]. */
int pb_mcts_reset_root(pb_mcts *m, const uint8_t *mask_dev, const float *value_dev,
                       const float *policy_dev, void *stream);
/* add_root_policy_noise (mcts.rs:259-281): prior = (1 - mix) * prior + mix * noise over the
 * valid actions of the real env; the k-th valid action takes noise_dev[i][k] (float32 [n][5]). */
int pb_mcts_root_noise(pb_mcts *m, const uint8_t *mask_dev, const float *noise_dev, float mix,
                       void *stream);
/* The same with the noise drawn on the device: Dirichlet(alpha / num_valid) over the valid actions
 * (mcts.rs:263-271; the reference draws from thread_rng, so only the distribution is defined) from
 * the engine's counter RNG, keyed by (seed, tree id, number of draws that tree has taken). */
int pb_mcts_root_noise_dirichlet(pb_mcts *m, const uint8_t *mask_dev, float alpha, float mix, void *stream);
/* select_trajectory (mcts.rs:290-321) for every env with active_dev[i] != 0 (NULL = all).
 * leaf_kind_dev uint8 [n] (may be NULL): 0 inactive, 1 trajectory ended in a terminal state,
 * 2 unexpanded leaf (its state is slot i of `sim`: encode + evaluate it);
 * leaf_mask_dev uint8 [n][5] (may be NULL): action mask of that state. */
int pb_mcts_select(pb_mcts *m, const uint8_t *active_dev, uint8_t *leaf_kind_dev,
                   uint8_t *leaf_mask_dev, void *stream);
/* backprop_trajectory (mcts.rs:325-381) of the trajectories of the last pb_mcts_select;
 * value_dev / policy_dev are read only for leaf_kind 2.  Increments the epoch. */
int pb_mcts_backprop(pb_mcts *m, const float *value_dev, const float *policy_dev, void *stream);
/* take_action (mcts.rs:153-170) for every env with mask_dev[i] != 0 (NULL = all): steps the real
 * env (no auto reset), writes reward/done, and status_dev uint8 [n]: 0 untouched, 1 episode done,
 * 2 re-rooted to the existing child (caller adds root noise), 3 child absent (caller evaluates
 * the new root: pb_mcts_reset_root + noise), 4 env was already done (untouched; the reference
 * panics), 5 action > 4 (untouched). */
int pb_mcts_take_action(pb_mcts *m, const uint8_t *mask_dev, const uint8_t *actions_dev,
                        int32_t *reward_dev, uint8_t *done_dev, uint8_t *status_dev, void *stream);
/* The env-step half of ExperienceCollector::generate_experience_step (alphazero.rs:141-195) for
 * all trees in one launch.  remaining_dev int32 [n] (search iterations left before the tree
 * acts) is decremented; a tree that reaches 0 stores its ExperienceItem at slot
 * i * max_episode_length + ep_len_dev[i] -- the env STATE into engine `snap` (which needs
 * n * max_episode_length envs; encode it later instead of keeping a 59 KB observation),
 * ep_mask_dev uint8 [n][L][5], ep_dist_dev float32 [n][L][5] (visit distribution), ep_value_dev
 * float32 [n][L] (the reward) -- then plays its most visited action like pb_mcts_take_action,
 * increments ep_len_dev[i] and, if it was re-rooted, sets remaining_dev[i] = tree_size - nodes.
 * status_dev uint8 [n]: the pb_mcts_take_action status, plus 8 if the episode hit
 * max_episode_length without ending.  *flags_dev (uint32, or-ed): bit 0 an episode ended (done or
 * truncated: the caller finalises it and resets the env), bit 1 a tree needs a fresh root. */
int pb_mcts_collect_act(pb_mcts *m, pb_engine *snap, int max_episode_length, int tree_size,
                        int32_t *remaining_dev, int32_t *ep_len_dev, uint8_t *ep_mask_dev,
                        float *ep_dist_dev, float *ep_value_dev, uint8_t *status_dev,
                        uint32_t *flags_dev, void *stream);
/* pb_mcts_backprop + pb_mcts_collect_act + pb_mcts_root_noise_dirichlet (for the trees that
 * collect_act re-rooted, mcts.rs:160-163) as ONE launch: the tail of a collector step
 * (alphazero.rs:141-195) after the evaluator.  A tree's three stages only touch that tree's own
 * nodes, env and counters, so the results are those of the three separate calls. */
int pb_mcts_backprop_collect(pb_mcts *m, const float *value_dev, const float *policy_dev, pb_engine *snap,
                             int max_episode_length, int tree_size, int32_t *remaining_dev,
                             int32_t *ep_len_dev, uint8_t *ep_mask_dev, float *ep_dist_dev,
                             float *ep_value_dev, uint8_t *status_dev, uint32_t *flags_dev, float alpha,
                             float mix, int publish, void *stream);
/* publish != 0: the launch also reports to the host without a copy or a stream synchronisation.
 * Its last CTA takes the accumulated flag word (and clears *flags_dev), and writes
 * (number of publishing launches so far << 32) | flags to the pinned word handed out by
 * pb_mcts_step_word; the host polls that word for the count it expects. */
int pb_mcts_step_word(pb_mcts *m, uint64_t **host_word_out);
/* Per-tree getters (mcts.rs:173-243) */
#define PB_MCTS_Q_BEST_ACTION 0         /* int32 [n]: most visited valid child, ties -> last */
#define PB_MCTS_Q_ACTION_DISTRIBUTION 1 /* float32 [n][5]: child visits / (root visits - 1) */
#define PB_MCTS_Q_POLICY_PRIOR 2        /* float32 [n][5] */
#define PB_MCTS_Q_ACTION_VALUES 3       /* float32 [n][5]: child value, root value where absent */
#define PB_MCTS_Q_VALUE 4               /* float32 [n]: total_return / visit_count of the root */
#define PB_MCTS_Q_MAX_DEPTH 5           /* int32 [n] */
#define PB_MCTS_Q_NODE_COUNT 6          /* int32 [n] */
#define PB_MCTS_Q_ROOT_VISITS 7         /* int32 [n] */
int pb_mcts_query(pb_mcts *m, int what, void *out_dev, void *stream);
/* Synchronises `stream`; returns PB_ERR_INVALID_ARG (text in pb_last_error) if any kernel since
 * the last check hit a full node pool, a double expansion or a search from a finished env;
 * *epoch_out (may be NULL) receives the epoch counter. */
int pb_mcts_check(pb_mcts *m, uint64_t *epoch_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PACBOT_B200_H */
