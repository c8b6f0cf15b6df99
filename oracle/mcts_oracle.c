/*
 * mcts_oracle.c -- CPU ORACLE (test infrastructure): scalar restatement of the reference's MCTS
 * (/root/reference/pacbot_rs/src/mcts.rs) and AlphaZero experience collector
 * (/root/reference/pacbot_rs/src/alphazero.rs).  See mcts_oracle.h for what is injected.
 * Pointer-linked tree exactly like the reference's Box<SearchTreeNode>; the product uses index
 * pools on the GPU and shares no code with this file.
 */
#include "mcts_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ---- counter RNG (include/pacbot_b200.h "Counter RNG") ------------------------------------ */
static uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
uint32_t pm_draw_raw(uint64_t seed, uint64_t gid, uint32_t ctr) {
    uint64_t key = mix64(seed + 0x9e3779b97f4a7c15ull * (gid + 1ull));
    return (uint32_t)(mix64(key + (uint64_t)ctr * 0x9e3779b97f4a7c15ull + 0x632be59bd9b4e019ull) >> 32);
}
uint64_t pm_sim_seed(uint64_t seed, uint64_t epoch) {
    return mix64(seed ^ ((epoch + 1ull) * 0xd6e8feb86659fd93ull));
}
typedef struct {
    uint64_t seed, gid;
} rng_key;
static uint32_t counter_draw(void *ctx, uint32_t ctr) {
    const rng_key *k = (const rng_key *)ctx;
    return pm_draw_raw(k->seed, k->gid, ctr);
}

/* ---- SearchTreeNode (mcts.rs:17-111) ------------------------------------------------------- */
typedef struct pm_node {
    uint32_t visit_count;
    float total_return;
    float policy_priors[5];
    struct pm_node *children[5];
} pm_node;

static pm_node *node_new(const float policy[5]) { /* mcts.rs:37-40 */
    pm_node *n = (pm_node *)calloc(1, sizeof(pm_node));
    if (policy) memcpy(n->policy_priors, policy, sizeof(float) * 5);
    return n;
}
static void node_free(pm_node *n) {
    if (!n) return;
    for (int a = 0; a < 5; a++) node_free(n->children[a]);
    free(n);
}
static float node_value(const pm_node *n) { /* mcts.rs:82-86 */
    return n->total_return / (float)n->visit_count;
}
static float child_value(const pm_node *n, const pm_node *child) { /* mcts.rs:89-92 */
    return child ? node_value(child) : node_value(n);
}
/* mcts.rs:49-68 PUCT; Iterator::max_by_key returns the LAST maximal element */
static int pick_action(const pm_node *n, const po_gym *env) {
    uint8_t mask[5];
    po_gym_action_mask(env, mask);
    int best = -1;
    float best_key = 0.0f;
    for (int a = 0; a < 5; a++) {
        if (!mask[a]) continue;
        const pm_node *child = n->children[a];
        uint32_t child_visits = child ? child->visit_count : 0u;
        float cv = child_value(n, child);
        float exploration = n->policy_priors[a] * sqrtf((float)n->visit_count) / (float)(1u + child_visits);
        float key = cv + PM_EXPLORATION_RATE * exploration;
        if (best < 0 || key >= best_key) {
            best = a;
            best_key = key;
        }
    }
    return best;
}
static int best_action(const pm_node *n, const po_gym *env) { /* mcts.rs:71-79 */
    uint8_t mask[5];
    po_gym_action_mask(env, mask);
    int best = -1;
    uint32_t best_key = 0;
    for (int a = 0; a < 5; a++) {
        if (!mask[a]) continue;
        uint32_t v = n->children[a] ? n->children[a]->visit_count : 0u;
        if (best < 0 || v >= best_key) {
            best = a;
            best_key = v;
        }
    }
    return best;
}
static int64_t node_max_depth(const pm_node *n) { /* mcts.rs:95-101 */
    int64_t d = 0;
    for (int a = 0; a < 5; a++)
        if (n->children[a]) {
            int64_t c = 1 + node_max_depth(n->children[a]);
            if (c > d) d = c;
        }
    return d;
}
static int64_t node_count(const pm_node *n) { /* mcts.rs:104-110 */
    int64_t c = 1;
    for (int a = 0; a < 5; a++)
        if (n->children[a]) c += node_count(n->children[a]);
    return c;
}

/* ---- MCTSContext (mcts.rs:113-382) ----------------------------------------------------------- */
typedef struct {
    po_gym env;
    pm_node *root;
    rng_key root_key, sim_key;
} pm_tree;

struct pm_batch {
    int64_t n;
    pm_tree *t;
    uint64_t seed, epoch;
    pm_eval_fn eval;
    void *eval_ctx;
    pm_noise_fn noise;
    void *noise_ctx;
    float *obs_scratch;
};

typedef struct {
    int action_index;
    int32_t reward;
} transition;

static void add_root_policy_noise(pm_batch *b, int64_t i) { /* mcts.rs:259-281 */
    pm_tree *t = &b->t[i];
    uint8_t mask[5];
    po_gym_action_mask(&t->env, mask);
    int num_valid = 0;
    for (int a = 0; a < 5; a++) num_valid += mask[a] ? 1 : 0;
    float noise[5] = {0, 0, 0, 0, 0};
    b->noise(b->noise_ctx, i, num_valid, noise);
    int k = 0;
    for (int a = 0; a < 5; a++) {
        if (!mask[a]) continue;
        float *prior = &t->root->policy_priors[a];
        *prior = (1.0f - PM_NOISE_MIX) * (*prior) + PM_NOISE_MIX * noise[k++];
    }
}
static void reset_root(pm_batch *b, int64_t i) { /* mcts.rs:252-257 */
    pm_tree *t = &b->t[i];
    uint8_t mask[5];
    float value, policy[5];
    po_gym_obs(&t->env, b->obs_scratch);
    po_gym_action_mask(&t->env, mask);
    b->eval(b->eval_ctx, i, b->obs_scratch, mask, &value, policy);
    node_free(t->root);
    t->root = node_new(policy); /* new_visited, mcts.rs:27-34 */
    t->root->visit_count = 1;
    t->root->total_return = value;
    add_root_policy_noise(b, i);
}

/* select_trajectory (mcts.rs:290-321); returns the trajectory length, *leaf = 1 if it ended at an
 * unexpanded node (then *leaf_env holds the state to evaluate) */
static int64_t select_trajectory(pm_batch *b, int64_t i, transition *traj, int64_t cap, int *leaf,
                                 po_gym *leaf_env) {
    pm_tree *t = &b->t[i];
    po_gym env = t->env; /* env.clone() */
    t->sim_key.seed = pm_sim_seed(b->seed, b->epoch);
    t->sim_key.gid = t->root_key.gid;
    po_gym_set_draws(&env, counter_draw, &t->sim_key);
    const pm_node *node = t->root;
    int64_t len = 0;
    for (;;) {
        int a = pick_action(node, &env);
        int32_t reward;
        uint8_t done;
        po_gym_step(&env, (uint8_t)a, &reward, &done);
        if (len >= cap) abort();
        traj[len].action_index = a;
        traj[len].reward = reward;
        len++;
        if (done) {
            *leaf = 0;
            return len;
        } else if (node->children[a]) {
            node = node->children[a];
        } else {
            *leaf = 1;
            *leaf_env = env;
            return len;
        }
    }
}

/* backprop_helper (mcts.rs:331-372) */
static float backprop_helper(pm_node *node, const transition *traj, int64_t len, int has_eval,
                             float value, const float *policy) {
    pm_node **child = &node->children[traj[0].action_index];
    float subsequent_return;
    if (len == 1) {
        if (has_eval) {
            if (*child) abort(); /* "tried to expand the same child twice" */
            *child = node_new(policy);
            subsequent_return = value;
        } else {
            if (!*child) *child = node_new(NULL); /* new_terminal */
            subsequent_return = 0.0f;
        }
    } else {
        if (!*child) abort(); /* "backprop hit an unexpanded child" */
        subsequent_return = backprop_helper(*child, traj + 1, len - 1, has_eval, value, policy);
    }
    float this_return = (float)traj[0].reward + PM_DISCOUNT * subsequent_return;
    (*child)->visit_count += 1;
    (*child)->total_return += this_return;
    return this_return;
}
static void backprop_trajectory(pm_tree *t, const transition *traj, int64_t len, int has_eval,
                                float value, const float *policy) { /* mcts.rs:325-381 */
    float subsequent_return = backprop_helper(t->root, traj, len, has_eval, value, policy);
    t->root->visit_count += 1;
    t->root->total_return += PM_DISCOUNT * subsequent_return;
}

pm_batch *pm_batch_create(int64_t n, const uint8_t *random_start, const uint8_t *random_ticks,
                          uint64_t seed, uint64_t gid_base, pm_eval_fn eval, void *eval_ctx,
                          pm_noise_fn noise, void *noise_ctx) {
    pm_batch *b = (pm_batch *)calloc(1, sizeof(pm_batch));
    b->n = n;
    b->t = (pm_tree *)calloc((size_t)n, sizeof(pm_tree));
    b->seed = seed;
    b->eval = eval;
    b->eval_ctx = eval_ctx;
    b->noise = noise;
    b->noise_ctx = noise_ctx;
    b->obs_scratch = (float *)malloc(sizeof(float) * PO_OBS_FLOATS);
    for (int64_t i = 0; i < n; i++) {
        pm_tree *t = &b->t[i];
        po_gym_new(&t->env, random_start ? random_start[i] : 0, random_ticks ? random_ticks[i] : 0);
        t->root_key.seed = seed;
        t->root_key.gid = gid_base + (uint64_t)i;
        po_gym_set_draws(&t->env, counter_draw, &t->root_key);
    }
    return b;
}
void pm_batch_destroy(pm_batch *b) {
    if (!b) return;
    for (int64_t i = 0; i < b->n; i++) node_free(b->t[i].root);
    free(b->t);
    free(b->obs_scratch);
    free(b);
}
po_gym *pm_batch_env(pm_batch *b, int64_t i) { return &b->t[i].env; }
uint64_t pm_batch_epoch(const pm_batch *b) { return b->epoch; }

void pm_batch_reset(pm_batch *b, const uint8_t *mask, int reset_env) {
    for (int64_t i = 0; i < b->n; i++) {
        if (mask && !mask[i]) continue;
        if (reset_env) po_gym_reset(&b->t[i].env);
        reset_root(b, i);
    }
}

void pm_batch_search_iteration(pm_batch *b, const uint8_t *active) {
    transition *traj = (transition *)malloc(sizeof(transition) * 65536);
    for (int64_t i = 0; i < b->n; i++) {
        if (active && !active[i]) continue;
        pm_tree *t = &b->t[i];
        int leaf = 0;
        po_gym leaf_env;
        int64_t len = select_trajectory(b, i, traj, 65536, &leaf, &leaf_env);
        float value = 0.0f, policy[5] = {0, 0, 0, 0, 0};
        if (leaf) {
            uint8_t mask[5];
            po_gym_obs(&leaf_env, b->obs_scratch);
            po_gym_action_mask(&leaf_env, mask);
            b->eval(b->eval_ctx, i, b->obs_scratch, mask, &value, policy);
        }
        backprop_trajectory(t, traj, len, leaf, value, policy);
    }
    free(traj);
    b->epoch += 1;
}

void pm_batch_take_action(pm_batch *b, const uint8_t *mask, const uint8_t *actions, int32_t *reward,
                          uint8_t *done) { /* mcts.rs:153-170 */
    for (int64_t i = 0; i < b->n; i++) {
        if (mask && !mask[i]) continue;
        pm_tree *t = &b->t[i];
        po_gym_step(&t->env, actions[i], &reward[i], &done[i]);
        if (!done[i]) {
            pm_node *child = t->root->children[actions[i]];
            if (child) {
                t->root->children[actions[i]] = NULL;
                node_free(t->root);
                t->root = child;
                add_root_policy_noise(b, i);
            } else {
                reset_root(b, i);
            }
        }
    }
}

int pm_best_action(pm_batch *b, int64_t i) { return best_action(b->t[i].root, &b->t[i].env); }
void pm_action_distribution(pm_batch *b, int64_t i, float out[5]) { /* mcts.rs:180-186 */
    const pm_node *r = b->t[i].root;
    float total_child_visits = (float)(r->visit_count - 1u);
    for (int a = 0; a < 5; a++)
        out[a] = r->children[a] ? (float)r->children[a]->visit_count / total_child_visits : 0.0f;
}
void pm_policy_prior(pm_batch *b, int64_t i, float out[5]) {
    memcpy(out, b->t[i].root->policy_priors, sizeof(float) * 5);
}
void pm_action_values(pm_batch *b, int64_t i, float out[5]) { /* mcts.rs:196-198 */
    const pm_node *r = b->t[i].root;
    for (int a = 0; a < 5; a++) out[a] = child_value(r, r->children[a]);
}
float pm_value(pm_batch *b, int64_t i) { return node_value(b->t[i].root); }
int64_t pm_max_depth(pm_batch *b, int64_t i) { return node_max_depth(b->t[i].root); }
int64_t pm_node_count(pm_batch *b, int64_t i) { return node_count(b->t[i].root); }
uint32_t pm_root_visits(pm_batch *b, int64_t i) { return b->t[i].root->visit_count; }

/* ---- ExperienceCollector (alphazero.rs:58-197) ----------------------------------------------- */
typedef struct {
    int64_t num_search_iters_remaining;
    pm_item *experience;
    int64_t len, cap;
} parallel_env;

struct pm_collector {
    pm_config cfg;
    pm_batch *b;
    parallel_env *envs;
    pm_item *final_items;
    int64_t final_len, final_cap;
};

pm_collector *pm_collector_create(pm_config cfg, uint64_t seed, uint64_t gid_base, pm_eval_fn eval,
                                  void *eval_ctx, pm_noise_fn noise, void *noise_ctx) {
    pm_collector *c = (pm_collector *)calloc(1, sizeof(pm_collector));
    c->cfg = cfg;
    int64_t n = cfg.num_parallel_envs;
    uint8_t *rs = (uint8_t *)malloc((size_t)n), *rt = (uint8_t *)malloc((size_t)n);
    memset(rs, 1, (size_t)n); /* PacmanGym::new(true, false), alphazero.rs:72 */
    memset(rt, 0, (size_t)n);
    c->b = pm_batch_create(n, rs, rt, seed, gid_base, eval, eval_ctx, noise, noise_ctx);
    free(rs);
    free(rt);
    pm_batch_reset(c->b, NULL, 1); /* env.reset() + MCTSContext::new -> reset_root */
    c->envs = (parallel_env *)calloc((size_t)n, sizeof(parallel_env));
    for (int64_t i = 0; i < n; i++) c->envs[i].num_search_iters_remaining = cfg.tree_size - 1;
    return c;
}
void pm_collector_destroy(pm_collector *c) {
    if (!c) return;
    for (int64_t i = 0; i < c->cfg.num_parallel_envs; i++) free(c->envs[i].experience);
    free(c->envs);
    free(c->final_items);
    pm_batch_destroy(c->b);
    free(c);
}
pm_batch *pm_collector_batch(pm_collector *c) { return c->b; }

static void end_episode(pm_collector *c, int64_t i) { /* alphazero.rs:163-176 */
    parallel_env *e = &c->envs[i];
    float next_state_value = 0.0f;
    for (int64_t k = e->len - 1; k >= 0; k--) {
        e->experience[k].value += c->cfg.discount_factor * next_state_value;
        next_state_value = e->experience[k].value;
    }
    if (c->final_len + e->len > c->final_cap) {
        c->final_cap = (c->final_len + e->len) * 2;
        c->final_items = (pm_item *)realloc(c->final_items, sizeof(pm_item) * (size_t)c->final_cap);
    }
    memcpy(c->final_items + c->final_len, e->experience, sizeof(pm_item) * (size_t)e->len);
    c->final_len += e->len;
    e->len = 0;
    uint8_t *mask = (uint8_t *)calloc((size_t)c->b->n, 1);
    mask[i] = 1;
    pm_batch_reset(c->b, mask, 1); /* mcts_context.reset() */
    free(mask);
}

static void generate_experience_step(pm_collector *c) { /* alphazero.rs:97-196 */
    pm_batch *b = c->b;
    pm_batch_search_iteration(b, NULL); /* select for all, batched eval, backprop for all */
    for (int64_t i = 0; i < b->n; i++) {
        parallel_env *e = &c->envs[i];
        e->num_search_iters_remaining -= 1;
        if (e->num_search_iters_remaining != 0) continue;
        if (e->len == e->cap) {
            e->cap = e->cap ? e->cap * 2 : 64;
            e->experience = (pm_item *)realloc(e->experience, sizeof(pm_item) * (size_t)e->cap);
        }
        pm_item *it = &e->experience[e->len];
        po_gym_obs(&b->t[i].env, it->obs);
        po_gym_action_mask(&b->t[i].env, it->action_mask);
        pm_action_distribution(b, i, it->action_distribution);
        int action = pm_best_action(b, i);
        int32_t reward;
        uint8_t done;
        /* take_action for this one tree */
        uint8_t *mask = (uint8_t *)calloc((size_t)b->n, 1);
        uint8_t *acts = (uint8_t *)calloc((size_t)b->n, 1);
        int32_t *rew = (int32_t *)calloc((size_t)b->n, sizeof(int32_t));
        uint8_t *dn = (uint8_t *)calloc((size_t)b->n, 1);
        mask[i] = 1;
        acts[i] = (uint8_t)action;
        pm_batch_take_action(b, mask, acts, rew, dn);
        reward = rew[i];
        done = dn[i];
        free(mask);
        free(acts);
        free(rew);
        free(dn);
        it->value = (float)reward;
        e->len += 1;
        if (done) {
            end_episode(c, i);
        } else if (e->len >= c->cfg.max_episode_length) {
            e->experience[e->len - 1].value += c->cfg.discount_factor * pm_value(b, i);
            end_episode(c, i);
        }
        int64_t nodes = pm_node_count(b, i);
        e->num_search_iters_remaining = c->cfg.tree_size > nodes ? c->cfg.tree_size - nodes : 0;
    }
}

/* ---- CPU baseline (bench.py "mcts" leg): collector with the untrained-model evaluator --------- */
/* train_alphazero.py:97-101: value 0, policy = mask / sum(mask) */
void pm_eval_uniform(void *ctx, int64_t tree, const float *obs, const uint8_t *mask, float *value,
                     float *policy) {
    (void)ctx; (void)tree; (void)obs;
    int k = 0;
    for (int a = 0; a < 5; a++) k += mask[a] ? 1 : 0;
    *value = 0.0f;
    for (int a = 0; a < 5; a++) policy[a] = mask[a] ? 1.0f / (float)k : 0.0f;
}
void pm_noise_uniform(void *ctx, int64_t tree, int num_valid, float *noise) {
    (void)ctx; (void)tree;
    for (int k = 0; k < num_valid; k++) noise[k] = 1.0f / (float)num_valid;
}
/* runs `steps` generate_experience_steps (each = one search iteration of every tree, leaf
 * observation encode included, plus the env steps that become due); returns seconds */
double pm_collector_bench(pm_config cfg, uint64_t seed, int64_t steps, int64_t *env_steps) {
    pm_collector *c = pm_collector_create(cfg, seed, 0, pm_eval_uniform, NULL, pm_noise_uniform, NULL);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    int64_t items = 0;
    for (int64_t s = 0; s < steps; s++) {
        generate_experience_step(c);
        items += c->final_len;
        c->final_len = 0;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    for (int64_t i = 0; i < cfg.num_parallel_envs; i++) items += c->envs[i].len;
    if (env_steps) *env_steps = items;
    pm_collector_destroy(c);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

int64_t pm_collector_generate(pm_collector *c, const pm_item **items) { /* alphazero.rs:86-92 */
    c->final_len = 0;
    while (c->final_len == 0) generate_experience_step(c);
    *items = c->final_items;
    return c->final_len;
}
