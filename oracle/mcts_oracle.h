/*
 * mcts_oracle.h -- CPU ORACLE for the MCTS / AlphaZero experience path (test infrastructure,
 * NOT product code; same rules as pacbot_oracle.h: only tests/, smoke() and bench.py's CPU
 * baseline legs may load it).
 *
 * Scalar C restatement of
 *   /root/reference/pacbot_rs/src/mcts.rs       (SearchTreeNode, MCTSContext)
 *   /root/reference/pacbot_rs/src/alphazero.rs  (AlphaZeroConfig, ExperienceCollector)
 * on top of the env oracle (pacbot_oracle.h).  Each function cites the lines it follows.
 *
 * What the reference leaves to chance is INJECTED here so that the CUDA path can be compared
 * bit for bit:
 *   - leaf / root evaluations come from a callback (the reference calls a Python evaluator,
 *     mcts.rs:384-416);
 *   - the Dirichlet root noise (mcts.rs:259-281, rand::thread_rng) comes from a callback that
 *     returns `num_valid` weights;
 *   - random draws of the SIMULATED env steps (the reference's cloned env shares the thread RNG,
 *     mcts.rs:293-299) come from the counter RNG documented in include/pacbot_b200.h, keyed by
 *     (sim_seed(seed, epoch), gid, rng_ctr) where `epoch` counts batched search iterations;
 *     the real env uses (seed, gid, rng_ctr).
 * Floating point: built with -ffp-contract=off; every operation is a single IEEE f32 op in the
 * reference's order (Rust never fuses).
 */
#ifndef MCTS_ORACLE_H
#define MCTS_ORACLE_H

#include "pacbot_oracle.h"

#ifdef __cplusplus
extern "C" {
#endif

#define PM_DISCOUNT 0.99f          /* mcts.rs:15 DISCOUNT_FACTOR */
#define PM_EXPLORATION_RATE 100.0f /* mcts.rs:59 */
#define PM_NOISE_MIX 0.25f         /* mcts.rs:262 */
#define PM_DIRICHLET_ALPHA 11.0f   /* mcts.rs:263 (alpha / num_valid per component) */

/* evaluator: obs (17*28*31 f32) + mask (5) of tree `tree` -> value, policy[5] (mcts.rs:384-398) */
typedef void (*pm_eval_fn)(void *ctx, int64_t tree, const float *obs, const uint8_t *mask,
                           float *value, float *policy);
/* Dirichlet(alpha/num_valid) sample of size num_valid for tree `tree` (mcts.rs:268-271) */
typedef void (*pm_noise_fn)(void *ctx, int64_t tree, int num_valid, float *noise);

typedef struct pm_batch pm_batch;

/* counter RNG of the product (include/pacbot_b200.h), restated */
uint32_t pm_draw_raw(uint64_t seed, uint64_t gid, uint32_t ctr);
uint64_t pm_sim_seed(uint64_t seed, uint64_t epoch);

/* n trees; env i = PacmanGym::new(random_start[i], random_ticks[i]) with counter draws keyed
 * (seed, gid_base + i); NOT reset and NO root yet (call pm_batch_reset). */
pm_batch *pm_batch_create(int64_t n, const uint8_t *random_start, const uint8_t *random_ticks,
                          uint64_t seed, uint64_t gid_base, pm_eval_fn eval, void *eval_ctx,
                          pm_noise_fn noise, void *noise_ctx);
void pm_batch_destroy(pm_batch *b);
po_gym *pm_batch_env(pm_batch *b, int64_t i);
uint64_t pm_batch_epoch(const pm_batch *b);

/* MCTSContext::reset (mcts.rs:138-145) for masked trees (mask NULL = all): env.reset + reset_root.
 * reset_env == 0: only reset_root (MCTSContext::new on an existing env, mcts.rs:131-135). */
void pm_batch_reset(pm_batch *b, const uint8_t *mask, int reset_env);
/* one search iteration (body of ponder_and_choose, mcts.rs:213-221 == one
 * generate_experience_step search, alphazero.rs:106-145) for every active tree; epoch += 1 */
void pm_batch_search_iteration(pm_batch *b, const uint8_t *active);
/* MCTSContext::take_action (mcts.rs:153-170) for masked trees */
void pm_batch_take_action(pm_batch *b, const uint8_t *mask, const uint8_t *actions, int32_t *reward,
                          uint8_t *done);

/* per-tree queries (mcts.rs:173-243) */
int pm_best_action(pm_batch *b, int64_t i);
void pm_action_distribution(pm_batch *b, int64_t i, float out[5]);
void pm_policy_prior(pm_batch *b, int64_t i, float out[5]);
void pm_action_values(pm_batch *b, int64_t i, float out[5]);
float pm_value(pm_batch *b, int64_t i);
int64_t pm_max_depth(pm_batch *b, int64_t i);
int64_t pm_node_count(pm_batch *b, int64_t i);
uint32_t pm_root_visits(pm_batch *b, int64_t i);

/* ---- ExperienceCollector (alphazero.rs:58-197) --------------------------------------------- */
typedef struct {
    int64_t tree_size, max_episode_length;
    float discount_factor;
    int64_t num_parallel_envs;
} pm_config;
typedef struct pm_collector pm_collector;
typedef struct {
    float obs[PO_OBS_FLOATS];
    uint8_t action_mask[5];
    float value;
    float action_distribution[5];
} pm_item;

/* envs are PacmanGym::new(true, false), reset, roots evaluated (alphazero.rs:69-83) */
pm_collector *pm_collector_create(pm_config cfg, uint64_t seed, uint64_t gid_base, pm_eval_fn eval,
                                  void *eval_ctx, pm_noise_fn noise, void *noise_ctx);
void pm_collector_destroy(pm_collector *c);
/* generate_experience (alphazero.rs:86-92): returns the number of items, *items points at an
 * internal array valid until the next call */
int64_t pm_collector_generate(pm_collector *c, const pm_item **items);
pm_batch *pm_collector_batch(pm_collector *c);

/* CPU baseline of bench.py's "mcts" leg: the collector with the untrained-model evaluator
 * (train_alphazero.py:97-101) and uniform root noise, single thread */
void pm_eval_uniform(void *ctx, int64_t tree, const float *obs, const uint8_t *mask, float *value,
                     float *policy);
void pm_noise_uniform(void *ctx, int64_t tree, int num_valid, float *noise);
double pm_collector_bench(pm_config cfg, uint64_t seed, int64_t steps, int64_t *env_steps);

#ifdef __cplusplus
}
#endif
#endif
