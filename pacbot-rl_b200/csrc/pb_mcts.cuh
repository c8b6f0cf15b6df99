// pb_mcts.cuh -- batched Monte-Carlo tree search on the device: one search tree per env, all trees
// advanced in lock step by one kernel launch per phase.
//
// Reference semantics: /root/reference/pacbot_rs/src/mcts.rs
//   SearchTreeNode {visit_count, total_return, policy_priors[5], children[5]}      :17-23
//   pick_action (PUCT, exploration rate 100, ties -> LAST maximum)                 :49-68
//   best_action / value / child_value / max_depth / node_count                     :71-110
//   MCTSContext::take_action (step, re-root to the child or fresh root)            :153-170
//   select_trajectory (clone env, walk until done or unexpanded)                   :290-321
//   backprop_trajectory (expand leaf / terminal child, discounted returns, 0.99)   :325-381
//   add_root_policy_noise (0.75 * prior + 0.25 * noise over the valid actions)     :259-281
//
// B200 shape: the reference keeps one Box-linked tree per MCTSContext on the host and calls a
// Python evaluator per leaf (or per batch of leaves, alphazero.rs:106-133), copying every leaf
// observation host->device.  Here the trees live in HBM as index pools (64-byte nodes; re-rooting
// moves the root index and only occasionally compacts the tree into a second pool, see
// MctsView), the env clone a simulation steps is a slot
// of a second engine (`sim`), the leaf observation is written by k_encode_obs straight from that
// slot, and the evaluator consumes it on the device.  One thread walks one tree: the walk is a
// dependent chain of node loads and env steps, so trees -- not nodes -- are the parallel axis.
//
// Floating point: every operation is a single IEEE f32 op in the reference's order
// (__fdiv_rn / __fmul_rn / __fadd_rn / __fsqrt_rn: no FMA contraction), so tree statistics are
// bit-identical to the scalar oracle (oracle/mcts_oracle.c) given the same evaluations.
#pragma once
#include "pb_engine.cuh"

namespace pb {

struct __align__(16) MctsNode {
    uint32_t visit;
    float total;
    float prior[5];
    int32_t child[5];  // index into the same pool, -1 = None
    uint32_t size;     // nodes in this subtree, itself included (node_count(), mcts.rs:104-110)
    uint32_t pad[3];
};
static_assert(sizeof(MctsNode) == 64, "MctsNode is 64 bytes");

constexpr float MCTS_DISCOUNT = 0.99f;          // mcts.rs:15
constexpr float MCTS_EXPLORATION_RATE = 100.0f; // mcts.rs:59
constexpr int MCTS_THREADS = 32;

// leaf_kind values
enum { MCTS_INACTIVE = 0, MCTS_TERMINAL = 1, MCTS_NEEDS_EVAL = 2 };
// take_action status values
enum {
    MCTS_ST_UNTOUCHED = 0,
    MCTS_ST_DONE = 1,         // episode ended, tree left as is
    MCTS_ST_REROOTED = 2,     // root := existing child (caller adds root noise)
    MCTS_ST_NEEDS_ROOT = 3,   // child did not exist: caller evaluates the new root state
    MCTS_ST_WAS_DONE = 4,     // env was already done (reference panics): untouched
    MCTS_ST_BAD_ACTION = 5,   // action > 4: untouched
};
// error bits (sticky, read by pb_mcts_check)
enum { MCTS_ERR_POOL_FULL = 1u, MCTS_ERR_EXPAND_TWICE = 2u, MCTS_ERR_ROOT_DONE = 4u };

// Node storage: M physical slots per tree and pool, of which at most L (= max_nodes) are LIVE
// (reachable from the root).  take_action normally just moves the root index to the chosen child
// and leaves the rest of the old tree behind as garbage; only when the free tail of the pool
// could not hold the live-node budget any more is the kept subtree copied, compacted, into the
// other pool (M = 8 L by default: about one compaction per ten moves instead of one per move).
struct MctsView {
    MctsNode *pool;            // [2][n][M]
    int32_t *cur;              // [n] pool in use
    int32_t *root;             // [n] index of the root node in that pool
    int32_t *n_alloc;          // [n] slots handed out in that pool (live + garbage)
    int32_t *traj_len;         // [n]
    int32_t *path;             // [n][L] node at depth k (parent of transition k)
    uint8_t *traj_act;         // [n][L]
    int32_t *traj_rew;         // [n][L]
    uint8_t *leaf_kind;        // [n]
    int32_t *stack;            // [n][L] scratch (compaction, max_depth)
    unsigned long long *epoch; // number of completed search iterations
    int32_t *noise_calls;      // [n] built-in root-noise draws taken so far (RNG counter)
    unsigned int *error;       // sticky error bits
    long long n;
    int M;                     // physical slots per tree and pool
    int L;                     // live-node capacity (max_nodes)
    int tpw;                   // trees per warp (1..32), see mcts_tree_of_thread
    __device__ __forceinline__ MctsNode *nodes(long long i, int which) const {
        return pool + ((long long)which * n + i) * M;
    }
};

// Thread -> tree map of every MCTS kernel (blocks of one warp).  A tree walk is branchy (PUCT,
// the env step, variable depth), so the 32 lanes of a warp that each walk their own tree execute
// the UNION of 32 instruction streams one after the other: ncu showed ~9 700 warp instructions for
// ~1 500 per tree.  While there are fewer trees than the GPU has warp slots, only the first `tpw`
// lanes of a warp own a tree (tpw = 1 for the reference's 128 trees: one tree per warp, 128 SMs
// busy instead of 4); idle lanes get index n and leave through the kernels' bounds check.
__device__ __forceinline__ long long mcts_tree_of_thread(const MctsView &v) {
    return (int)threadIdx.x < v.tpw ? (long long)blockIdx.x * v.tpw + threadIdx.x : v.n;
}

__device__ __forceinline__ void node_init(MctsNode &nd, uint32_t visit, float total, const float *prior) {
    nd.visit = visit;
    nd.total = total;
#pragma unroll
    for (int a = 0; a < 5; a++) {
        nd.prior[a] = prior ? prior[a] : 0.0f;
        nd.child[a] = -1;
    }
    nd.size = 1u;
    nd.pad[0] = nd.pad[1] = nd.pad[2] = 0u;
}

// ---------------------------------------------------------------------------------------------
// reset_root (mcts.rs:252-257, without the noise): root := new_visited(evaluation)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MCTS_THREADS)
k_mcts_reset_root(MctsView v, const uint8_t *__restrict__ mask, const float *__restrict__ value,
                  const float *__restrict__ policy) {
    const long long i = mcts_tree_of_thread(v);
    if (i >= v.n || (mask && !mask[i])) return;
    MctsNode nd;
    node_init(nd, 1u, value[i], policy + i * 5);
    v.nodes(i, v.cur[i])[0] = nd;  // the old tree, wherever it was, is dropped
    v.root[i] = 0;
    v.n_alloc[i] = 1;
}

// add_root_policy_noise (mcts.rs:259-281): the k-th VALID action gets noise[i][k]
__global__ void __launch_bounds__(MCTS_THREADS)
k_mcts_root_noise(MctsView v, StateView root, const uint8_t *__restrict__ mask,
                  const float *__restrict__ noise, float keep, float mix) {
    __shared__ uint32_t s_walls[32];
    load_walls(s_walls);
    const long long i = mcts_tree_of_thread(v);
    if (i >= v.n || (mask && !mask[i])) return;
    EnvR s;
    load_env(root, i, s);
    const uint32_t m = action_mask_bits(s, s_walls);
    MctsNode *nd = v.nodes(i, v.cur[i]) + v.root[i];
    int k = 0;
    for (int a = 0; a < 5; a++) {
        if (!((m >> a) & 1u)) continue;
        nd->prior[a] = __fadd_rn(__fmul_rn(keep, nd->prior[a]), __fmul_rn(mix, noise[i * 5 + k]));
        k++;
    }
}

// add_root_policy_noise with the noise drawn on the device (mcts.rs:259-281: Dirichlet with
// concentration alpha / num_valid over the valid actions, mixed into the root prior).  The
// reference draws from rand_distr::Dirichlet on thread_rng, so there is no stream to reproduce:
// this is the same distribution from the engine's counter RNG, keyed (seed, tree id, call number
// of that tree), one launch instead of the ~14 of the torch formulation (mask, sum, arange,
// _standard_gamma, normalise ...).  gamma(a), a >= 1 (alpha = 11, at most 5 valid actions):
// Marsaglia-Tsang, normals by Box-Muller.
__device__ __forceinline__ float noise_uniform(uint64_t key, uint32_t &ctr) {
    const uint32_t raw = (uint32_t)(mix64(key + (uint64_t)(ctr++) * 0x9e3779b97f4a7c15ull) >> 32);
    return ((float)raw + 0.5f) * (1.0f / 4294967296.0f);   // (0, 1)
}
__device__ __forceinline__ float noise_gamma(float a, uint64_t key, uint32_t &ctr) {
    const float boost = a < 1.0f ? powf(noise_uniform(key, ctr), 1.0f / a) : 1.0f;   // gamma(a) = gamma(a+1) U^(1/a)
    if (a < 1.0f) a += 1.0f;
    const float d = a - 1.0f / 3.0f, c = 1.0f / sqrtf(9.0f * d);
    for (int tries = 0; tries < 64; tries++) {
        const float u1 = noise_uniform(key, ctr), u2 = noise_uniform(key, ctr);
        const float x = sqrtf(-2.0f * logf(u1)) * cosf(6.283185307179586f * u2);
        float v = 1.0f + c * x;
        if (v <= 0.0f) continue;
        v = v * v * v;
        const float u = noise_uniform(key, ctr);
        if (logf(u) < 0.5f * x * x + d - d * v + d * logf(v)) return d * v * boost;
    }
    return d * boost;   // (the mean; 64 rejections in a row do not happen)
}

// add_root_policy_noise with the built-in Dirichlet draw, for tree i (one thread)
__device__ __forceinline__ void mcts_dirichlet_noise_tree(const MctsView &v, const StateView &root,
                                                          const uint32_t *s_walls, long long i, float alpha,
                                                          float keep, float mix, unsigned long long seed,
                                                          long long gid_base) {
    EnvR s;
    load_env(root, i, s);
    const uint32_t m = action_mask_bits(s, s_walls);
    const int nv = __popc(m);
    const uint32_t call = (uint32_t)v.noise_calls[i];
    v.noise_calls[i] = (int32_t)(call + 1u);
    const uint64_t key = mix64(env_key(seed ^ 0xd1b54a32d192ed03ull, (uint64_t)(gid_base + i)) +
                               (uint64_t)call * 0xd6e8feb86659fd93ull);
    uint32_t ctr = 0;
    float g[5], sum = 0.0f;
    for (int k = 0; k < nv; k++) {
        g[k] = noise_gamma(alpha / (float)nv, key, ctr);
        sum += g[k];
    }
    MctsNode *nd = v.nodes(i, v.cur[i]) + v.root[i];
    int k = 0;
    for (int a = 0; a < 5; a++) {
        if (!((m >> a) & 1u)) continue;
        const float noise = sum > 0.0f ? g[k] / sum : 1.0f / (float)nv;
        nd->prior[a] = __fadd_rn(__fmul_rn(keep, nd->prior[a]), __fmul_rn(mix, noise));
        k++;
    }
}

__global__ void __launch_bounds__(MCTS_THREADS)
k_mcts_root_noise_dirichlet(MctsView v, StateView root, const uint8_t *__restrict__ mask, float alpha,
                            float keep, float mix, unsigned long long seed, long long gid_base) {
    __shared__ uint32_t s_walls[32];
    load_walls(s_walls);
    const long long i = mcts_tree_of_thread(v);
    if (i >= v.n || (mask && !mask[i])) return;
    mcts_dirichlet_noise_tree(v, root, s_walls, i, alpha, keep, mix, seed, gid_base);
}

// ---------------------------------------------------------------------------------------------
// select_trajectory (mcts.rs:290-321)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int mcts_pick_action(const MctsNode &nd, const MctsNode *nodes, uint32_t m) {
    const float parent_value = __fdiv_rn(nd.total, (float)nd.visit);   // mcts.rs:82-86
    const float sqrt_visits = __fsqrt_rn((float)nd.visit);
    // children's statistics: independent loads, issued together
    uint32_t cvis[5];
    float ctot[5];
#pragma unroll
    for (int a = 0; a < 5; a++) {
        const int c = nd.child[a];
        cvis[a] = c >= 0 ? nodes[c].visit : 0u;
        ctot[a] = c >= 0 ? nodes[c].total : 0.0f;
    }
    int best = -1;
    float best_key = 0.0f;
#pragma unroll
    for (int a = 0; a < 5; a++) {
        if (!((m >> a) & 1u)) continue;
        const float cv = nd.child[a] >= 0 ? __fdiv_rn(ctot[a], (float)cvis[a]) : parent_value;
        const float exploration = __fdiv_rn(__fmul_rn(nd.prior[a], sqrt_visits), (float)(1u + cvis[a]));
        const float key = __fadd_rn(cv, __fmul_rn(MCTS_EXPLORATION_RATE, exploration));
        if (best < 0 || key >= best_key) {  // max_by_key: the LAST maximum wins
            best = a;
            best_key = key;
        }
    }
    return best;
}

__global__ void __launch_bounds__(MCTS_THREADS)
k_mcts_select(MctsView v, StateView root, StateView sim, uint64_t seed, long long gid_base,
              const uint8_t *__restrict__ active, uint8_t *__restrict__ leaf_kind_out,
              uint8_t *__restrict__ leaf_mask_out) {
    __shared__ uint32_t s_walls[32];
    // the simulation's pellet words live in shared memory while the tree is walked: every tick
    // of every simulated env step reads (and eating writes) them, and with one tree per warp
    // nothing hides a global-memory round trip per access.  Word w of the tree owned by lane l
    // sits at s_pel[w * 32 + l] (a lane only ever touches its own column: no synchronisation);
    // the 28 loads that bring them in are requested together.
    __shared__ uint32_t s_pel[PEL_WORDS * MCTS_THREADS];
    load_walls(s_walls);
    const int lane = threadIdx.x;
    const long long i = mcts_tree_of_thread(v);
    const bool owner = i < v.n;
    if (!owner) return;
    EnvR s;
    load_env(root, i, s);
    int kind = MCTS_INACTIVE;
    if (!active || active[i]) {
        if (is_done(s)) {  // reference: assert!(!self.env.is_done())
            atomicOr(v.error, MCTS_ERR_ROOT_DONE);
        } else {
            // env.clone(): the simulation steps slot i of the `sim` engine
            uint32_t words[PEL_WORDS];
#pragma unroll
            for (int w = 0; w < PEL_WORDS; w++) words[w] = root.pel[(long long)w * root.npad + i];
#pragma unroll
            for (int w = 0; w < PEL_WORDS; w++) s_pel[w * MCTS_THREADS + lane] = words[w];
            const PelRef pel{s_pel + lane, MCTS_THREADS};
            RngCtx rng;
            rng.seed = sim_seed(seed, *v.epoch);
            rng.gid = (uint64_t)(gid_base + i);
            rng.tape = nullptr;
            rng.tape_len = 0;
            const MctsNode *nodes = v.nodes(i, v.cur[i]);
            int32_t *path = v.path + i * v.L;
            uint8_t *act = v.traj_act + i * v.L;
            int32_t *rew = v.traj_rew + i * v.L;
            int node = v.root[i], len = 0;
            for (;;) {
                const MctsNode nd = nodes[node];
                const int a = mcts_pick_action(nd, nodes, action_mask_bits(s, s_walls));
                int reward;
                bool done;
                gym_step(s, pel, c_tab, s_walls, rng, a, reward, done);
                path[len] = node;
                act[len] = (uint8_t)a;
                rew[len] = reward;
                len++;
                if (done) {
                    kind = MCTS_TERMINAL;
                    break;
                }
                if (nd.child[a] < 0) {
                    kind = MCTS_NEEDS_EVAL;
                    break;
                }
                if (len >= v.L) {  // cannot happen with <= L live nodes (depth < live nodes)
                    atomicOr(v.error, MCTS_ERR_POOL_FULL);
                    kind = MCTS_INACTIVE;
                    break;
                }
                node = nd.child[a];
            }
            v.traj_len[i] = len;
            store_env(sim, i, s);
#pragma unroll
            for (int w = 0; w < PEL_WORDS; w++)
                sim.pel[(long long)w * sim.npad + i] = s_pel[w * MCTS_THREADS + lane];
        }
    }
    v.leaf_kind[i] = (uint8_t)kind;
    if (leaf_kind_out) leaf_kind_out[i] = (uint8_t)kind;
    if (leaf_mask_out) write_mask(leaf_mask_out, i, action_mask_bits(s, s_walls));
}

// ---------------------------------------------------------------------------------------------
// backprop_trajectory (mcts.rs:325-381), iterative: expand the leaf, then walk the path upwards
// ---------------------------------------------------------------------------------------------
// backprop_trajectory for tree i (one thread): expand the leaf, then walk the path upwards
__device__ __forceinline__ void mcts_backprop_tree(const MctsView &v, long long i,
                                                   const float *__restrict__ value,
                                                   const float *__restrict__ policy) {
    const int kind = v.leaf_kind[i];
    if (kind == MCTS_INACTIVE) return;
    v.leaf_kind[i] = MCTS_INACTIVE;  // a trajectory is consumed exactly once
    MctsNode *nodes = v.nodes(i, v.cur[i]);
    const int32_t *path = v.path + i * v.L;
    const uint8_t *act = v.traj_act + i * v.L;
    const int32_t *rew = v.traj_rew + i * v.L;
    const int len = v.traj_len[i];
    const int parent = path[len - 1], a_last = act[len - 1];
    int c = nodes[parent].child[a_last];
    float subsequent;
    if (kind == MCTS_NEEDS_EVAL) {
        if (c >= 0) {  // "tried to expand the same child twice"
            atomicOr(v.error, MCTS_ERR_EXPAND_TWICE);
            return;
        }
        subsequent = value[i];
    } else {
        subsequent = 0.0f;
    }
    uint32_t grown = 0u;
    if (c < 0) {
        c = v.n_alloc[i];
        if (c >= v.M || nodes[path[0]].size >= (uint32_t)v.L) {  // physical / live capacity
            atomicOr(v.error, MCTS_ERR_POOL_FULL);
            return;
        }
        v.n_alloc[i] = c + 1;
        MctsNode nd;  // SearchTreeNode::new(policy) / new_terminal()
        node_init(nd, 0u, 0.0f, kind == MCTS_NEEDS_EVAL ? policy + i * 5 : nullptr);
        nodes[c] = nd;
        nodes[parent].child[a_last] = c;
        grown = 1u;  // every node on the path has one more descendant
    }
    int child = c;
    for (int k = len - 1; k >= 0; k--) {
        const float this_return = __fadd_rn((float)rew[k], __fmul_rn(MCTS_DISCOUNT, subsequent));
        nodes[child].visit += 1u;
        nodes[child].total = __fadd_rn(nodes[child].total, this_return);
        subsequent = this_return;
        child = path[k];
        nodes[child].size += grown;
    }
    // child == path[0] == root
    nodes[child].visit += 1u;
    nodes[child].total = __fadd_rn(nodes[child].total, __fmul_rn(MCTS_DISCOUNT, subsequent));
}

__global__ void __launch_bounds__(MCTS_THREADS)
k_mcts_backprop(MctsView v, const float *__restrict__ value, const float *__restrict__ policy) {
    const long long i = mcts_tree_of_thread(v);
    if (i == 0) *v.epoch += 1ull;  // one search iteration done (select of the next one sees it)
    if (i >= v.n) return;
    mcts_backprop_tree(v, i, value, policy);
}

// ---------------------------------------------------------------------------------------------
// take_action (mcts.rs:153-170): step the real env; keep the chosen child's subtree as the new
// tree (compacting copy into the other pool) or ask for a fresh root
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int mcts_take_action(const MctsView &v, const StateView &root,
                                                const RngParams &rp, const uint32_t *s_walls,
                                                long long i, int a, int &rew, bool &done) {
    rew = 0;
    done = false;
    EnvR s;
    load_env(root, i, s);
    if (is_done(s)) return MCTS_ST_WAS_DONE;
    if (a > 4) return MCTS_ST_BAD_ACTION;
    const PelRef pel{root.pel + i, root.npad};
    const RngCtx rng = make_rng(rp, i);
    gym_step(s, pel, c_tab, s_walls, rng, a, rew, done);
    store_env(root, i, s);
    if (done) return MCTS_ST_DONE;
    const int which = v.cur[i];
    const MctsNode *old_nodes = v.nodes(i, which);
    const int c = old_nodes[v.root[i]].child[a];
    if (c < 0) return MCTS_ST_NEEDS_ROOT;
    // self.root = *child.  Normally that is just a new root index (the old root and its other
    // subtrees stay behind as garbage); the kept subtree can still grow to L live nodes, so when
    // the free tail of the pool is shorter than that growth it is compacted into the other pool.
    const int live = (int)old_nodes[c].size;
    if (v.M - v.n_alloc[i] >= v.L - live) {
        v.root[i] = c;
        return MCTS_ST_REROOTED;
    }
    // compacting copy, depth first, new root at index 0
    MctsNode *new_nodes = v.nodes(i, which ^ 1);
    int32_t *stack = v.stack + i * v.L;
    new_nodes[0] = old_nodes[c];
    int count = 1, sp = 0;
    stack[sp++] = 0;
    while (sp > 0) {
        const int ni = stack[--sp];
#pragma unroll
        for (int b = 0; b < 5; b++) {
            const int oc = new_nodes[ni].child[b];  // still an index into the OLD pool
            if (oc >= 0) {
                const int nn = count++;
                new_nodes[nn] = old_nodes[oc];
                new_nodes[ni].child[b] = nn;
                stack[sp++] = nn;
            }
        }
    }
    v.cur[i] = which ^ 1;
    v.root[i] = 0;
    v.n_alloc[i] = count;
    return MCTS_ST_REROOTED;
}

__global__ void __launch_bounds__(MCTS_THREADS)
k_mcts_take_action(MctsView v, StateView root, RngParams rp, const uint8_t *__restrict__ mask,
                   const uint8_t *__restrict__ actions, int32_t *__restrict__ reward,
                   uint8_t *__restrict__ done_out, uint8_t *__restrict__ status) {
    __shared__ uint32_t s_walls[32];
    load_walls(s_walls);
    const long long i = mcts_tree_of_thread(v);
    if (i >= v.n) return;
    if (mask && !mask[i]) {
        status[i] = MCTS_ST_UNTOUCHED;
        return;
    }
    int rew;
    bool done;
    status[i] = (uint8_t)mcts_take_action(v, root, rp, s_walls, i, actions[i], rew, done);
    reward[i] = rew;
    done_out[i] = done ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// The env-step half of ExperienceCollector::generate_experience_step (alphazero.rs:141-195) for
// all trees in one launch: count the finished search iteration; a tree whose budget is used up
// records its experience (state snapshot instead of the 59 KB observation, action mask, visit
// distribution), plays its most visited action (take_action) and gets its next search budget.
// Episodes that END here (done, or max_episode_length reached) are flagged for the host, which
// finalises them (value recursion, hand-out, env reset + fresh root); so are trees that need a
// fresh root.  Experience slot of (tree i, step t): i * max_len + t.
// ---------------------------------------------------------------------------------------------
enum { MCTS_ST_TRUNCATED_BIT = 8 };  // or-ed into the take_action status when the episode is cut

// the env-step half of generate_experience_step for tree i (one thread); returns the status byte
__device__ __forceinline__ int mcts_collect_act_tree(
    const MctsView &v, const StateView &root, const RngParams &rp, const StateView &snap,
    const uint32_t *s_walls, long long i, int max_len, int tree_size, int32_t *__restrict__ remaining,
    int32_t *__restrict__ ep_len, uint8_t *__restrict__ ep_mask, float *__restrict__ ep_dist,
    float *__restrict__ ep_value, uint8_t *__restrict__ status, unsigned int *__restrict__ host_flags) {
    const int left = remaining[i] - 1;  // alphazero.rs:142
    remaining[i] = left;
    if (left != 0) {
        status[i] = MCTS_ST_UNTOUCHED;
        return MCTS_ST_UNTOUCHED;
    }
    const int t = ep_len[i];
    const long long slot = i * max_len + t;
    // ExperienceItem (alphazero.rs:145-161): the state stands in for root_obs_numpy()
    snap.core0[slot] = root.core0[i];
    snap.core1[slot] = root.core1[i];
    snap.ghosts[slot] = root.ghosts[i];
    snap.ghosts[snap.npad + slot] = root.ghosts[root.npad + i];
    for (int w = 0; w < PEL_WORDS; w++)
        snap.pel[(long long)w * snap.npad + slot] = root.pel[(long long)w * root.npad + i];
    EnvR s;
    load_env(root, i, s);
    const uint32_t m = action_mask_bits(s, s_walls);
    const MctsNode *nodes = v.nodes(i, v.cur[i]);
    const MctsNode r = nodes[v.root[i]];
    const float total_child_visits = (float)(r.visit - 1u);  // mcts.rs:180-186
    int best = -1;
    uint32_t best_key = 0;
    for (int a = 0; a < 5; a++) {
        const uint32_t vis = r.child[a] >= 0 ? nodes[r.child[a]].visit : 0u;
        ep_mask[slot * 5 + a] = (m >> a) & 1u;
        ep_dist[slot * 5 + a] = r.child[a] >= 0 ? __fdiv_rn((float)vis, total_child_visits) : 0.0f;
        if (((m >> a) & 1u) && (best < 0 || vis >= best_key)) {  // best_action, mcts.rs:71-79
            best = a;
            best_key = vis;
        }
    }
    int rew;
    bool done;
    int st = mcts_take_action(v, root, rp, s_walls, i, best, rew, done);
    ep_value[slot] = (float)rew;  // alphazero.rs:157
    ep_len[i] = t + 1;
    unsigned int flags = 0;
    if (st == MCTS_ST_DONE) flags |= 1u;
    if (st != MCTS_ST_DONE && t + 1 >= max_len) {  // alphazero.rs:177-184
        st |= MCTS_ST_TRUNCATED_BIT;
        flags |= 1u;
    }
    if ((st & 7) == MCTS_ST_NEEDS_ROOT || (st & 7) == MCTS_ST_WAS_DONE) flags |= 2u;
    if (st == MCTS_ST_REROOTED) {  // alphazero.rs:192-193
        const int nodes_now = (int)v.nodes(i, v.cur[i])[v.root[i]].size;
        remaining[i] = tree_size > nodes_now ? tree_size - nodes_now : 0;
    }
    status[i] = (uint8_t)st;
    if (flags) atomicOr(host_flags, flags);
    return st;
}

__global__ void __launch_bounds__(MCTS_THREADS)
k_mcts_collect_act(MctsView v, StateView root, RngParams rp, StateView snap, int max_len,
                   int tree_size, int32_t *__restrict__ remaining, int32_t *__restrict__ ep_len,
                   uint8_t *__restrict__ ep_mask, float *__restrict__ ep_dist,
                   float *__restrict__ ep_value, uint8_t *__restrict__ status,
                   unsigned int *__restrict__ host_flags) {
    __shared__ uint32_t s_walls[32];
    load_walls(s_walls);
    const long long i = mcts_tree_of_thread(v);
    if (i >= v.n) return;
    mcts_collect_act_tree(v, root, rp, snap, s_walls, i, max_len, tree_size, remaining, ep_len, ep_mask,
                          ep_dist, ep_value, status, host_flags);
}

// One search iteration's tail for the collector, ONE launch instead of three (plus the two torch
// kernels that derived the noise mask from the status bytes): backprop of the evaluated leaf, the
// env step that became due (experience record, take_action, next budget) and, for a tree that
// take_action re-rooted, the Dirichlet noise on its new root (mcts.rs:160-163).  A tree's three
// stages only touch that tree's own nodes, env and counters, so running them back to back in the
// tree's thread is the order the separate launches give.
// `host_word` (optional, pinned host memory): the step's report to the host without a copy or a
// stream synchronisation.  The last CTA to finish takes the flag word accumulated by all trees
// (and clears it), counts the launch in `launches`, and publishes (launch count << 32) | flags after
// a system fence; the host polls that word (alphazero.py).  `arrive` is zero between launches.
__global__ void __launch_bounds__(MCTS_THREADS)
k_mcts_backprop_collect(MctsView v, const float *__restrict__ value, const float *__restrict__ policy,
                        StateView root, RngParams rp, StateView snap, int max_len, int tree_size,
                        int32_t *__restrict__ remaining, int32_t *__restrict__ ep_len,
                        uint8_t *__restrict__ ep_mask, float *__restrict__ ep_dist,
                        float *__restrict__ ep_value, uint8_t *__restrict__ status,
                        unsigned int *__restrict__ host_flags, float alpha, float keep, float mix,
                        unsigned long long noise_seed, long long gid_base, unsigned int *arrive,
                        unsigned long long *launches, unsigned long long *host_word) {
    __shared__ uint32_t s_walls[32];
    load_walls(s_walls);
    const long long i = mcts_tree_of_thread(v);
    if (i == 0) *v.epoch += 1ull;
    if (i < v.n) {
        mcts_backprop_tree(v, i, value, policy);
        const int st = mcts_collect_act_tree(v, root, rp, snap, s_walls, i, max_len, tree_size, remaining,
                                             ep_len, ep_mask, ep_dist, ep_value, status, host_flags);
        if ((st & 7) == MCTS_ST_REROOTED)
            mcts_dirichlet_noise_tree(v, root, s_walls, i, alpha, keep, mix, noise_seed, gid_base);
    }
    if (!host_word) return;
    __syncthreads();   // every tree of this CTA has raised its flags
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int before = atomicAdd(arrive, 1u);
        if (before == gridDim.x - 1) {
            __threadfence();
            const unsigned long long flags = atomicExch(host_flags, 0u);
            const unsigned long long seq = *launches + 1ull;
            *launches = seq;
            atomicExch(arrive, 0u);
            __threadfence_system();
            *reinterpret_cast<volatile unsigned long long *>(host_word) = (seq << 32) | flags;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// queries (mcts.rs:173-243)
// ---------------------------------------------------------------------------------------------
enum {
    MCTS_Q_BEST_ACTION = 0,          // int32 [n]
    MCTS_Q_ACTION_DISTRIBUTION = 1,  // float32 [n][5]
    MCTS_Q_POLICY_PRIOR = 2,         // float32 [n][5]
    MCTS_Q_ACTION_VALUES = 3,        // float32 [n][5]
    MCTS_Q_VALUE = 4,                // float32 [n]
    MCTS_Q_MAX_DEPTH = 5,            // int32 [n]
    MCTS_Q_NODE_COUNT = 6,           // int32 [n]
    MCTS_Q_ROOT_VISITS = 7,          // int32 [n]
};

__global__ void __launch_bounds__(MCTS_THREADS)
k_mcts_query(MctsView v, StateView root, int what, void *__restrict__ out) {
    __shared__ uint32_t s_walls[32];
    load_walls(s_walls);
    const long long i = mcts_tree_of_thread(v);
    if (i >= v.n) return;
    const MctsNode *nodes = v.nodes(i, v.cur[i]);
    const int root_idx = v.root[i];
    const MctsNode r = nodes[root_idx];
    float *fo = reinterpret_cast<float *>(out);
    int32_t *io = reinterpret_cast<int32_t *>(out);
    switch (what) {
    case MCTS_Q_BEST_ACTION: {  // mcts.rs:71-79 (ties -> last)
        EnvR s;
        load_env(root, i, s);
        const uint32_t m = action_mask_bits(s, s_walls);
        int best = -1;
        uint32_t best_key = 0;
        for (int a = 0; a < 5; a++) {
            if (!((m >> a) & 1u)) continue;
            const uint32_t vis = r.child[a] >= 0 ? nodes[r.child[a]].visit : 0u;
            if (best < 0 || vis >= best_key) {
                best = a;
                best_key = vis;
            }
        }
        io[i] = best;
        break;
    }
    case MCTS_Q_ACTION_DISTRIBUTION: {  // mcts.rs:180-186
        const float total_child_visits = (float)(r.visit - 1u);
        for (int a = 0; a < 5; a++)
            fo[i * 5 + a] = r.child[a] >= 0
                                ? __fdiv_rn((float)nodes[r.child[a]].visit, total_child_visits)
                                : 0.0f;
        break;
    }
    case MCTS_Q_POLICY_PRIOR:
        for (int a = 0; a < 5; a++) fo[i * 5 + a] = r.prior[a];
        break;
    case MCTS_Q_ACTION_VALUES: {  // mcts.rs:196-198
        const float pv = __fdiv_rn(r.total, (float)r.visit);
        for (int a = 0; a < 5; a++) {
            const int c = r.child[a];
            fo[i * 5 + a] = c >= 0 ? __fdiv_rn(nodes[c].total, (float)nodes[c].visit) : pv;
        }
        break;
    }
    case MCTS_Q_VALUE:
        fo[i] = __fdiv_rn(r.total, (float)r.visit);
        break;
    case MCTS_Q_MAX_DEPTH: {  // mcts.rs:95-101; stack entries = node | depth << 16
        int32_t *stack = v.stack + i * v.L;
        int sp = 0, best = 0;
        stack[sp++] = root_idx;
        while (sp > 0) {
            const int e = stack[--sp];
            const int ni = e & 0xffff, d = e >> 16;
            best = d > best ? d : best;
            for (int a = 0; a < 5; a++) {
                const int c = nodes[ni].child[a];
                if (c >= 0) stack[sp++] = c | ((d + 1) << 16);
            }
        }
        io[i] = best;
        break;
    }
    case MCTS_Q_NODE_COUNT:
        io[i] = (int32_t)r.size;
        break;
    case MCTS_Q_ROOT_VISITS:
        io[i] = (int32_t)r.visit;
        break;
    default:
        break;
    }
}

// ---------------------------------------------------------------------------------------------
// pb_copy_envs: batched #[derive(Clone)] between engines (one warp per copy)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_copy_envs(StateView dst, const long long *__restrict__ dst_idx, StateView src,
            const long long *__restrict__ src_idx, long long count) {
    const long long k = ((long long)blockIdx.x * 128 + threadIdx.x) >> 5;
    const int t = threadIdx.x & 31;
    if (k >= count) return;
    const long long di = dst_idx ? dst_idx[k] : k, si = src_idx ? src_idx[k] : k;
    if (di < 0 || si < 0 || di >= dst.n || si >= src.n) return;
    if (t == 0) dst.core0[di] = src.core0[si];
    if (t == 1) dst.core1[di] = src.core1[si];
    if (t == 2) dst.ghosts[di] = src.ghosts[si];
    if (t == 3) dst.ghosts[dst.npad + di] = src.ghosts[src.npad + si];
    if (t >= 4 && t < 4 + PEL_WORDS)
        dst.pel[(long long)(t - 4) * dst.npad + di] = src.pel[(long long)(t - 4) * src.npad + si];
}

}  // namespace pb
