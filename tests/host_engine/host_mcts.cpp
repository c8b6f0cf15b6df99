// host_mcts.cpp -- TEST INFRASTRUCTURE: the device MCTS kernels of the product
// (pacbot-rl_b200/csrc/pb_mcts.cuh: k_mcts_reset_root / root_noise / select / backprop /
// take_action / collect_act / query, k_copy_envs) and a host engine object around the device game
// logic (create / reset / step / state access), compiled for the HOST with g++ and run thread by
// thread, so that tests/test_host_mcts.py and tests/test_host_replay.py can run the product's
// Python host code over the *same source the GPU executes* and diff it against the scalar
// oracles (oracle/) in a container without a GPU.  A "launch" here is a loop over
// (block, thread) with threadIdx / blockIdx set before every call; the kernels have no
// intra-launch communication (one thread owns one tree), so the order does not matter.
// The glue around the kernels mirrors pacbot-rl_b200/csrc/pb_kernels.cu (load_env, store_env,
// RngParams, write_mask) and pb_mcts_abi.inl (allocation sizes, grid size, pool factor).
// NOT a CPU path of the product: nothing under pacbot-rl_b200/ builds, loads or links it.
#include <cmath>
#include <cstdlib>
#include <vector>

#include "host_common.h"

// ---- what pb_mcts.cuh needs beyond the cuda_runtime.h shim ----------------------------------
#define __global__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) alignas(n)
struct HostDim3 {
    unsigned x, y, z;
};
static HostDim3 threadIdx, blockIdx, gridDim;
// the publishing tail of k_mcts_backprop_collect (last-CTA hand-over to a pinned host word) only
// has to compile here: the host build never passes a host word, the tail returns before it
static inline void __syncthreads() {}
static inline void __threadfence() {}
static inline void __threadfence_system() {}
template <class T>
static inline T atomicAdd(T *p, T v) {
    const T old = *p;
    *p += v;
    return old;
}
template <class T>
static inline T atomicExch(T *p, T v) {
    const T old = *p;
    *p = v;
    return old;
}
static inline unsigned atomicOr(unsigned *p, unsigned v) {
    const unsigned old = *p;
    *p |= v;
    return old;
}
// single IEEE f32 operations (this file is built with -ffp-contract=off, no fast-math)
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }

namespace pb {

static const MazeTables c_tab = H_TABLES;

inline void load_env(const StateView &st, long long i, EnvR &s) {
    unpack_env(st.core0[i], st.core1[i], st.ghosts[i], st.ghosts[st.npad + i], s);
}
inline void store_env(const StateView &st, long long i, const EnvR &s) {
    uint4 c0, c1, ga, gb;
    pack_env(s, c0, c1, ga, gb);
    st.core0[i] = c0;
    st.core1[i] = c1;
    st.ghosts[i] = ga;
    st.ghosts[st.npad + i] = gb;
}
struct RngParams {
    uint64_t seed;
    long long gid_base;
    const uint32_t *tape;
    long long tape_len;
};
inline RngCtx make_rng(const RngParams &rp, long long i) {
    RngCtx r;
    r.seed = rp.seed;
    r.gid = (uint64_t)(rp.gid_base + i);
    r.tape = rp.tape ? rp.tape + i * rp.tape_len : nullptr;
    r.tape_len = (uint32_t)rp.tape_len;
    return r;
}
inline void load_walls(uint32_t *s_walls) {
    for (int k = 0; k < 32; k++) s_walls[k] = c_tab.walls[k];
}
inline void write_mask(uint8_t *mask_out, long long i, uint32_t m) {
    for (int k = 0; k < 5; k++) mask_out[i * 5 + k] = (m >> k) & 1u;
}

}  // namespace pb

#include "pb_mcts.cuh"

using namespace pb;
using namespace host_common;

namespace {

struct HEngine {
    StateView st{};
    uint64_t seed = 0;
    long long gid_base = 0;
    const uint32_t *tape = nullptr;   // pb_set_draw_tape: borrowed, [n][tape_len]
    long long tape_len = 0;
    std::vector<uint4> core0, core1, ghosts;
    std::vector<uint32_t> pel;
};

struct HMcts {
    HEngine *root = nullptr, *sim = nullptr;
    MctsView v{};
    std::vector<MctsNode> pool;
    std::vector<int32_t> cur, rootv, n_alloc, traj_len, path, traj_rew, stack, noise_calls;
    std::vector<uint8_t> traj_act, leaf_kind;
    unsigned long long epoch = 0;
    unsigned int error = 0;
};

RngParams rng_params(const HEngine *e) { return RngParams{e->seed, e->gid_base, e->tape, e->tape_len}; }

template <class F>
void launch(const MctsView &v, F kernel) {
    const unsigned grid = (unsigned)((v.n + v.tpw - 1) / v.tpw);
    for (unsigned b = 0; b < grid; b++)
        for (unsigned t = 0; t < (unsigned)MCTS_THREADS; t++) {
            blockIdx.x = b;
            threadIdx.x = t;
            kernel();
        }
}

}  // namespace

extern "C" {

// ---- engines (pb_create / pb_reset / pb_get_state / pb_set_state / pb_action_mask) -------------
void *hm_engine_create(int64_t n, const uint8_t *cfg_flags, uint64_t seed, int64_t gid_base) {
    HEngine *e = new HEngine();
    const long long npad = (n + 31) / 32 * 32;
    e->core0.assign(npad, uint4{0, 0, 0, 0});
    e->core1.assign(npad, uint4{0, 0, 0, 0});
    e->ghosts.assign(2 * npad, uint4{0, 0, 0, 0});
    e->pel.assign((size_t)PEL_WORDS * npad, 0u);
    e->st = StateView{e->core0.data(), e->core1.data(), e->ghosts.data(), e->pel.data(), npad, n};
    e->seed = seed;
    e->gid_base = gid_base;
    uint32_t walls[32];
    load_walls(walls);
    for (int64_t i = 0; i < n; i++) {  // pb_create: zeroed state + cfg flags, k_reset(randomise = 0)
        e->core1[i].z = (uint32_t)(cfg_flags ? (cfg_flags[i] & 0x0f) : 0) << 8;
        EnvR s;
        load_env(e->st, i, s);
        const PelRef pel{e->st.pel + i, e->st.npad};
        gym_reset(s, pel, c_tab, walls, make_rng(rng_params(e), i), false);
        store_env(e->st, i, s);
    }
    return e;
}
void hm_engine_destroy(void *h) { delete static_cast<HEngine *>(h); }

void hm_engine_reset(void *h, const uint8_t *mask) {
    HEngine *e = static_cast<HEngine *>(h);
    uint32_t walls[32];
    load_walls(walls);
    for (long long i = 0; i < e->st.n; i++) {
        if (mask && !mask[i]) continue;
        EnvR s;
        load_env(e->st, i, s);
        const PelRef pel{e->st.pel + i, e->st.npad};
        gym_reset(s, pel, c_tab, walls, make_rng(rng_params(e), i), true);
        store_env(e->st, i, s);
    }
}

void hm_engine_set_tape(void *h, const uint32_t *tape, int64_t tape_len) {
    HEngine *e = static_cast<HEngine *>(h);
    e->tape = tape;
    e->tape_len = tape ? tape_len : 0;
}

// pb_step: what one k_step thread does (pb_kernels.cu), actions given by the caller; returns the
// index of the first invalid action or -1
int64_t hm_engine_step(void *h, const uint8_t *actions, int32_t *reward, uint8_t *done_out,
                       uint8_t *mask_out, int auto_reset) {
    HEngine *e = static_cast<HEngine *>(h);
    uint32_t walls[32];
    load_walls(walls);
    int64_t bad = -1;
    for (long long i = 0; i < e->st.n; i++) {
        EnvR s;
        load_env(e->st, i, s);
        const int action = actions[i];
        if (action > 4) {   // env.rs:83-88: "Invalid action" -- env untouched
            if (bad < 0) bad = i;
            reward[i] = 0;
            done_out[i] = 0;
            if (mask_out) write_mask(mask_out, i, action_mask_bits(s, walls));
            continue;
        }
        const PelRef pel{e->st.pel + i, e->st.npad};
        const RngCtx rng = make_rng(rng_params(e), i);
        int rew;
        bool done;
        gym_step(s, pel, c_tab, walls, rng, action, rew, done);
        if (done && auto_reset) gym_reset(s, pel, c_tab, walls, rng, true);
        store_env(e->st, i, s);
        reward[i] = rew;
        done_out[i] = done ? 1 : 0;
        if (mask_out) write_mask(mask_out, i, action_mask_bits(s, walls));
    }
    return bad;
}

void hm_engine_get_state(void *h, pb_env_state *out) {
    HEngine *e = static_cast<HEngine *>(h);
    for (long long i = 0; i < e->st.n; i++) {
        EnvR s;
        load_env(e->st, i, s);
        uint32_t f[PEL_WORDS];
        for (int w = 0; w < PEL_WORDS; w++) f[w] = e->st.pel[(long long)w * e->st.npad + i];
        to_host(s, f, out[i]);
    }
}
void hm_engine_set_state(void *h, const pb_env_state *in) {
    HEngine *e = static_cast<HEngine *>(h);
    for (long long i = 0; i < e->st.n; i++) {
        EnvR s;
        uint32_t f[PEL_WORDS];
        from_host(in[i], s, f);
        store_env(e->st, i, s);
        for (int w = 0; w < PEL_WORDS; w++) e->st.pel[(long long)w * e->st.npad + i] = f[w];
    }
}
void hm_engine_action_mask(void *h, uint8_t *out) {
    HEngine *e = static_cast<HEngine *>(h);
    uint32_t walls[32];
    load_walls(walls);
    for (long long i = 0; i < e->st.n; i++) {
        EnvR s;
        load_env(e->st, i, s);
        write_mask(out, i, action_mask_bits(s, walls));
    }
}

// ---- trees (pb_mcts_*) ---------------------------------------------------------------------------
void *hm_mcts_create(void *root_h, void *sim_h, int max_nodes, int pool_factor, int tpw) {
    HMcts *m = new HMcts();
    m->root = static_cast<HEngine *>(root_h);
    m->sim = static_cast<HEngine *>(sim_h);
    MctsView &v = m->v;
    v.n = m->root->st.n;
    v.L = max_nodes;
    const long long phys = (long long)max_nodes * (pool_factor > 0 ? pool_factor : 8);
    v.M = (int)(phys > 65535 ? 65535 : phys);
    v.tpw = tpw >= 1 && tpw <= 32 ? tpw : 1;
    const size_t n = (size_t)v.n, L = (size_t)max_nodes, M = (size_t)v.M;
    m->pool.assign(2 * n * M, MctsNode{});
    m->cur.assign(n, 0);
    m->rootv.assign(n, 0);
    m->n_alloc.assign(n, 0);
    m->traj_len.assign(n, 0);
    m->path.assign(n * L, 0);
    m->traj_act.assign(n * L, 0);
    m->traj_rew.assign(n * L, 0);
    m->leaf_kind.assign(n, 0);
    m->stack.assign(n * L, 0);
    m->noise_calls.assign(n, 0);
    v.pool = m->pool.data();
    v.cur = m->cur.data();
    v.root = m->rootv.data();
    v.n_alloc = m->n_alloc.data();
    v.traj_len = m->traj_len.data();
    v.path = m->path.data();
    v.traj_act = m->traj_act.data();
    v.traj_rew = m->traj_rew.data();
    v.leaf_kind = m->leaf_kind.data();
    v.stack = m->stack.data();
    v.epoch = &m->epoch;
    v.noise_calls = m->noise_calls.data();
    v.error = &m->error;
    return m;
}
void hm_mcts_destroy(void *h) { delete static_cast<HMcts *>(h); }

void hm_mcts_reset_root(void *h, const uint8_t *mask, const float *value, const float *policy) {
    HMcts *m = static_cast<HMcts *>(h);
    launch(m->v, [&] { k_mcts_reset_root(m->v, mask, value, policy); });
}
void hm_mcts_root_noise(void *h, const uint8_t *mask, const float *noise, float mix) {
    HMcts *m = static_cast<HMcts *>(h);
    launch(m->v, [&] { k_mcts_root_noise(m->v, m->root->st, mask, noise, 1.0f - mix, mix); });
}
void hm_mcts_root_noise_dirichlet(void *h, const uint8_t *mask, float alpha, float mix) {
    HMcts *m = static_cast<HMcts *>(h);
    launch(m->v, [&] {
        k_mcts_root_noise_dirichlet(m->v, m->root->st, mask, alpha, 1.0f - mix, mix, m->root->seed,
                                    m->root->gid_base);
    });
}
void hm_mcts_select(void *h, const uint8_t *active, uint8_t *leaf_kind_out, uint8_t *leaf_mask_out) {
    HMcts *m = static_cast<HMcts *>(h);
    launch(m->v, [&] {
        k_mcts_select(m->v, m->root->st, m->sim->st, m->root->seed, m->root->gid_base, active,
                      leaf_kind_out, leaf_mask_out);
    });
}
void hm_mcts_backprop(void *h, const float *value, const float *policy) {
    HMcts *m = static_cast<HMcts *>(h);
    launch(m->v, [&] { k_mcts_backprop(m->v, value, policy); });
}
void hm_mcts_take_action(void *h, const uint8_t *mask, const uint8_t *actions, int32_t *reward,
                         uint8_t *done, uint8_t *status) {
    HMcts *m = static_cast<HMcts *>(h);
    launch(m->v, [&] {
        k_mcts_take_action(m->v, m->root->st, rng_params(m->root), mask, actions, reward, done, status);
    });
}
void hm_mcts_query(void *h, int what, void *out) {
    HMcts *m = static_cast<HMcts *>(h);
    launch(m->v, [&] { k_mcts_query(m->v, m->root->st, what, out); });
}
// pb_mcts_collect_act: the env-step half of generate_experience_step for all trees
int hm_mcts_collect_act(void *h, void *snap_h, int max_len, int tree_size, int32_t *remaining,
                        int32_t *ep_len, uint8_t *ep_mask, float *ep_dist, float *ep_value,
                        uint8_t *status, uint32_t *flags) {
    HMcts *m = static_cast<HMcts *>(h);
    HEngine *snap = static_cast<HEngine *>(snap_h);
    if (snap->st.n < m->v.n * (long long)max_len) return PB_ERR_INVALID_ARG;
    launch(m->v, [&] {
        k_mcts_collect_act(m->v, m->root->st, rng_params(m->root), snap->st, max_len, tree_size,
                           remaining, ep_len, ep_mask, ep_dist, ep_value, status, flags);
    });
    return PB_OK;
}

// pb_mcts_backprop_collect: backprop + env steps + built-in root noise, one "launch" (no host word:
// the publishing tail is not run here)
int hm_mcts_backprop_collect(void *h, const float *value, const float *policy, void *snap_h, int max_len,
                             int tree_size, int32_t *remaining, int32_t *ep_len, uint8_t *ep_mask,
                             float *ep_dist, float *ep_value, uint8_t *status, uint32_t *flags, float alpha,
                             float mix) {
    HMcts *m = static_cast<HMcts *>(h);
    HEngine *snap = static_cast<HEngine *>(snap_h);
    if (snap->st.n < m->v.n * (long long)max_len) return PB_ERR_INVALID_ARG;
    launch(m->v, [&] {
        k_mcts_backprop_collect(m->v, value, policy, m->root->st, rng_params(m->root), snap->st, max_len,
                                tree_size, remaining, ep_len, ep_mask, ep_dist, ep_value, status, flags, alpha,
                                1.0f - mix, mix, m->root->seed, m->root->gid_base, nullptr, nullptr, nullptr);
    });
    return PB_OK;
}

// pb_copy_envs (k_copy_envs: one warp per copy, 128-thread blocks)
int hm_engine_copy_envs(void *dst_h, const int64_t *dst_idx, void *src_h, const int64_t *src_idx,
                        int64_t count) {
    HEngine *dst = static_cast<HEngine *>(dst_h), *src = static_cast<HEngine *>(src_h);
    if ((!dst_idx && count > dst->st.n) || (!src_idx && count > src->st.n)) return PB_ERR_INVALID_ARG;
    const unsigned blocks = (unsigned)((count * 32 + 127) / 128);
    for (unsigned b = 0; b < blocks; b++)
        for (unsigned t = 0; t < 128; t++) {
            blockIdx.x = b;
            threadIdx.x = t;
            k_copy_envs(dst->st, reinterpret_cast<const long long *>(dst_idx), src->st,
                        reinterpret_cast<const long long *>(src_idx), count);
        }
    return PB_OK;
}

// returns the sticky error bits (and clears them), *epoch_out = completed search iterations
unsigned hm_mcts_check(void *h, uint64_t *epoch_out) {
    HMcts *m = static_cast<HMcts *>(h);
    if (epoch_out) *epoch_out = m->epoch;
    const unsigned err = m->error;
    m->error = 0;
    return err;
}

}  // extern "C"
