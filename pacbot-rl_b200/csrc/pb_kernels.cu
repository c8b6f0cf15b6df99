// pb_kernels.cu -- sm_100a kernels and the C ABI (include/pacbot_b200.h) of the batched PacBot
// environment engine.  Build: nvcc -gencode arch=compute_100a,code=sm_100a (see build.py).
//
// Kernels
//   k_step       one env per thread; packed state in, reward/done/mask out, optional auto reset
//   k_reset      masked in-place PacmanGym::reset / constructor
//   k_encode_obs (pb_encode.cuh) one env per warp; the observation image is assembled in shared
//                memory and shipped to HBM by the TMA engine with cp.async.bulk (SASS UBLKCP)
//   k_mask / k_query / k_sample_actions / k_select_actions / k_gather_obs / k_clone
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/pacbot_b200.h"
#include "pb_engine.cuh"

namespace pb {

__constant__ MazeTables c_tab;

}  // namespace pb

#include "pb_encode.cuh"

namespace pb {

// =============================================================================================
// state load / store for the one-env-per-thread kernels
// =============================================================================================
__device__ __forceinline__ void load_env(const StateView &st, long long i, EnvR &s) {
    const uint4 c0 = st.core0[i], c1 = st.core1[i];
    const uint4 ga = st.ghosts[i], gb = st.ghosts[st.npad + i];
    unpack_env(c0, c1, ga, gb, s);
}
__device__ __forceinline__ void store_env(const StateView &st, long long i, const EnvR &s) {
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
__device__ __forceinline__ RngCtx make_rng(const RngParams &rp, long long i) {
    RngCtx r;
    r.seed = rp.seed;
    r.gid = (uint64_t)(rp.gid_base + i);
    r.tape = rp.tape ? rp.tape + i * rp.tape_len : nullptr;
    r.tape_len = (uint32_t)rp.tape_len;
    return r;
}

__device__ __forceinline__ void load_walls(uint32_t *s_walls) {
    if (threadIdx.x < 32) s_walls[threadIdx.x] = c_tab.walls[threadIdx.x];
    __syncthreads();
}

__device__ __forceinline__ void write_mask(uint8_t *mask_out, long long i, uint32_t m) {
    uint8_t *o = mask_out + i * 5;
#pragma unroll
    for (int k = 0; k < 5; k++) o[k] = (m >> k) & 1u;
}

// =============================================================================================
// k_step
// =============================================================================================
constexpr int STEP_THREADS = 128;

__device__ __forceinline__ int kth_set_bit(uint32_t m, int k) {
    while (k-- > 0) m &= m - 1;
    return __ffs(m) - 1;
}

// actions == nullptr: every env draws a uniformly random VALID action from its own mask (the
// action stream of the counter RNG, keyed by call_idx), i.e. k_sample_actions fused in; the
// chosen action is reported through actions_out (may be nullptr).
// Launched with STEP_THREADS threads per CTA, or with STEP_THREADS_WIDE when it has to share the
// SMs with the observation encoder (pb_step_flip): the encoder's CTA leaves 6 000 B of an SM's
// shared memory, every resident CTA of this kernel takes 1 152 B of it (128 B of walls + the 1 KB
// system reservation), so six 128-thread CTAs on one SM keep that SM's encoder CTA from
// launching.  With 512-thread CTAs at most four fit an SM (2 048 threads) and the encoder always
// finds room, whichever of the two kernels the hardware happens to start first.
constexpr int STEP_THREADS_WIDE = 512;
__global__ void __launch_bounds__(STEP_THREADS_WIDE)
k_step(StateView st, StateView dst, RngParams rp, const uint8_t *__restrict__ actions,
       int32_t *__restrict__ reward, uint8_t *__restrict__ done_out,
       uint8_t *__restrict__ mask_out, int auto_reset, unsigned long long *bad_action,
       uint8_t *__restrict__ actions_out, unsigned long long call_idx, long long env_first,
       long long env_count, TraceRec *__restrict__ trace) {
    if (trace && threadIdx.x == 0) {
        const unsigned long long t = global_timer_ns();
        atomicMin(&trace->begin, t);
        atomicMin(&trace->ready, t);
    }
    // let a programmatically-dependent kernel (the observation encoder) start its prologue now;
    // it still waits for this whole grid (griddepcontrol.wait) before reading the state
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    __shared__ uint32_t s_walls[32];
    load_walls(s_walls);
    // envs [env_first, env_first + env_count); every array is indexed by the absolute env index
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long i = env_first + k;
    if (k >= env_count) return;
    EnvR s;
    load_env(st, i, s);
    // double-buffered stepping (pb_step_random_flip): the new state goes to `dst`, `st` stays
    // intact for an encoder launch that is still reading it; in place when dst == st
    if (dst.pel != st.pel) {
#pragma unroll 4
        for (int w = 0; w < PEL_WORDS; w++)
            dst.pel[(long long)w * dst.npad + i] = st.pel[(long long)w * st.npad + i];
    }
    int action;
    if (actions) {
        action = actions[i];
    } else {
        const uint32_t m = action_mask_bits(s, s_walls);
        const uint32_t raw = action_raw(rp.seed, (uint64_t)(rp.gid_base + i), call_idx);
        action = kth_set_bit(m, (int)below(raw, (uint32_t)__popc(m)));
    }
    if (actions_out) actions_out[i] = (uint8_t)action;
    if (action > 4) {  // env.rs:83-88: "Invalid action" -- env untouched
        atomicMin(bad_action, (unsigned long long)i);
        reward[i] = 0;
        done_out[i] = 0;
        if (mask_out) write_mask(mask_out, i, action_mask_bits(s, s_walls));
        if (dst.pel != st.pel) store_env(dst, i, s);
        return;
    }
    const PelRef pel{dst.pel + i, dst.npad};
    const RngCtx rng = make_rng(rp, i);
    int rew;
    bool done;
    gym_step(s, pel, c_tab, s_walls, rng, action, rew, done);
    if (done && auto_reset) gym_reset(s, pel, c_tab, s_walls, rng, true);
    store_env(dst, i, s);
    reward[i] = rew;
    done_out[i] = done ? 1 : 0;
    if (mask_out) write_mask(mask_out, i, action_mask_bits(s, s_walls));
    if (trace && (threadIdx.x & 31) == 0) atomicMax(&trace->end, global_timer_ns());
}

// A mark on a stream's timeline (pb_trace_mark): one thread, one timestamp.
__global__ void k_trace_mark(TraceRec *trace) {
    const unsigned long long t = global_timer_ns();
    trace->begin = t;
    trace->ready = t;
    trace->end = t;
}

// =============================================================================================
// k_reset: mask == nullptr -> all envs; randomise == 0 -> constructor state (env.rs:133-138)
// =============================================================================================
__global__ void __launch_bounds__(STEP_THREADS)
k_reset(StateView st, RngParams rp, const uint8_t *__restrict__ mask, int randomise) {
    __shared__ uint32_t s_walls[32];
    load_walls(s_walls);
    const long long i = (long long)blockIdx.x * STEP_THREADS + threadIdx.x;
    if (i >= st.n) return;
    if (mask && !mask[i]) return;
    EnvR s;
    load_env(st, i, s);
    const PelRef pel{st.pel + i, st.npad};
    gym_reset(s, pel, c_tab, s_walls, make_rng(rp, i), randomise != 0);
    store_env(st, i, s);
}

// =============================================================================================
// k_mask / k_query
// =============================================================================================
__global__ void __launch_bounds__(STEP_THREADS) k_mask(StateView st, uint8_t *__restrict__ out) {
    __shared__ uint32_t s_walls[32];
    load_walls(s_walls);
    const long long i = (long long)blockIdx.x * STEP_THREADS + threadIdx.x;
    if (i >= st.n) return;
    EnvR s;
    load_env(st, i, s);
    write_mask(out, i, action_mask_bits(s, s_walls));
}

__device__ __forceinline__ bool first_ai_done(const PelRef &pel, const EnvR &s) {
    // env.rs:281-285: all four super pellets gone and no ghost frightened
    bool any = false;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int b = fbit(k < 2 ? 3 : 23, (k & 1) ? 26 : 1);
        any |= (pel.get(b >> 5) >> (b & 31)) & 1u;
    }
    return !any && s.g[0].fright == 0 && s.g[1].fright == 0 && s.g[2].fright == 0 &&
           s.g[3].fright == 0;
}
__device__ __forceinline__ bool first_ai_done(const StateView &st, long long i, const EnvR &s) {
    return first_ai_done(PelRef{st.pel + i, st.npad}, s);
}

__global__ void __launch_bounds__(STEP_THREADS)
k_query(StateView st, int what, int32_t *__restrict__ out) {
    const long long i = (long long)blockIdx.x * STEP_THREADS + threadIdx.x;
    if (i >= st.n) return;
    EnvR s;
    load_env(st, i, s);
    int v = 0;
    switch (what) {
    case PB_Q_SCORE: v = s.score; break;
    case PB_Q_LIVES: v = s.lives; break;
    case PB_Q_IS_DONE: v = is_done(s); break;
    case PB_Q_FIRST_AI_DONE: v = first_ai_done(st, i, s); break;
    case PB_Q_REMAINING_PELLETS: v = s.num_pellets; break;
    case PB_Q_TICKS_PER_STEP: v = s.tps; break;
    case PB_Q_CURR_TICKS: v = (int)s.ticks; break;
    case PB_Q_LEVEL: v = s.level; break;
    case PB_Q_UPDATE_PERIOD: v = s.period; break;
    }
    out[i] = v;
}

}  // namespace pb

// the small-batch forms: k_obs_small (one CTA per env) and k_gym_call (the per-env drop-in's whole
// round trip in one launch); they use load_env / store_env / first_ai_done / write_mask above
#include "pb_encode_small.cuh"

namespace pb {

// =============================================================================================
// k_sample_actions / k_select_actions (policies.py:16-59)
// =============================================================================================
// evaluate_episode's per-step bookkeeping (train_dqn.py:99-111) for all envs in one launch.  An env
// is "live" while its episode has not ended and it has steps of max_steps left; a live env counts
// the step and re-reads score / lives / remaining pellets AFTER it; the step that ends the episode
// (done) freezes them.  stats rows: 0 score, 1 steps, 2 lives, 3 pellets_end, 4 finished
// (done within max_steps), 5 alive, 6 steps left -- int32 [7][n].
__global__ void __launch_bounds__(STEP_THREADS)
k_eval_track(StateView st, const uint8_t *__restrict__ done, int32_t *__restrict__ stats) {
    const long long i = (long long)blockIdx.x * STEP_THREADS + threadIdx.x;
    if (i >= st.n) return;
    const long long n = st.n;
    const int left = stats[6 * n + i];
    const bool live = stats[5 * n + i] != 0 && left > 0;
    stats[6 * n + i] = left - 1;
    if (!live) return;
    EnvR s;
    load_env(st, i, s);
    stats[0 * n + i] = s.score;
    stats[1 * n + i] += 1;
    stats[2 * n + i] = s.lives;
    stats[3 * n + i] = s.num_pellets;
    if (done[i]) {
        stats[4 * n + i] = 1;
        stats[5 * n + i] = 0;
    }
}

__global__ void __launch_bounds__(STEP_THREADS)
k_sample_actions(StateView st, RngParams rp, uint8_t *__restrict__ actions,
                 unsigned long long call_idx) {
    __shared__ uint32_t s_walls[32];
    load_walls(s_walls);
    const long long i = (long long)blockIdx.x * STEP_THREADS + threadIdx.x;
    if (i >= st.n) return;
    const uint4 c1 = st.core1[i];
    EnvR s;
    s.prow = sx8(c1.x);
    s.pcol = sx8(c1.x >> 8);
    const uint32_t m = action_mask_bits(s, s_walls);
    const uint32_t raw = action_raw(rp.seed, (uint64_t)(rp.gid_base + i), call_idx);
    actions[i] = (uint8_t)kth_set_bit(m, (int)below(raw, (uint32_t)__popc(m)));
}

__global__ void __launch_bounds__(STEP_THREADS)
k_select_actions(StateView st, RngParams rp, const float *__restrict__ q,
                 const uint8_t *__restrict__ mask_in, float epsilon,
                 uint8_t *__restrict__ actions, unsigned long long call_idx,
                 const float *__restrict__ epsilon_ptr, const long long *__restrict__ call_ptr) {
    __shared__ uint32_t s_walls[32];
    load_walls(s_walls);
    const long long i = (long long)blockIdx.x * STEP_THREADS + threadIdx.x;
    if (i >= st.n) return;
    // pb_select_actions_dev: both live in device memory and are read at kernel time, so that a
    // captured CUDA graph follows the epsilon schedule and the step counter while it is replayed
    if (epsilon_ptr) epsilon = *epsilon_ptr;
    if (call_ptr) call_idx = (unsigned long long)*call_ptr;
    uint32_t m;
    if (mask_in) {
        m = 0;
#pragma unroll
        for (int k = 0; k < 5; k++) m |= (mask_in[i * 5 + k] ? 1u : 0u) << k;
    } else {
        const uint4 c1 = st.core1[i];
        EnvR s;
        s.prow = sx8(c1.x);
        s.pcol = sx8(c1.x >> 8);
        m = action_mask_bits(s, s_walls);
    }
    const uint64_t gid = (uint64_t)(rp.gid_base + i);
    // greedy iff rand > epsilon (policies.py:36), rand uniform in [0, 1) with 24 bits
    const float u = (float)(action_raw(rp.seed, gid, 2ull * call_idx) >> 8) * (1.0f / 16777216.0f);
    int a;
    if (u > epsilon || m == 0u) {
        float best = -INFINITY;
        a = 0;
        bool first = true;
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const float v = ((m >> k) & 1u) ? q[i * 5 + k] : -INFINITY;  // policies.py:57-58
            if (first || v > best) {
                best = v;
                a = k;
                first = false;
            }
        }
    } else {
        const uint32_t raw = action_raw(rp.seed, gid, 2ull * call_idx + 1ull);
        a = kth_set_bit(m, (int)below(raw, (uint32_t)__popc(m)));
    }
    actions[i] = (uint8_t)a;
}

// =============================================================================================
// k_gather_obs (train_dqn.py:153-159 collate) and k_clone
// =============================================================================================
__global__ void __launch_bounds__(256)
k_gather_obs(const float4 *__restrict__ ring, long long slot_stride4,
             const long long *__restrict__ idx, const uint8_t *__restrict__ zero_mask,
             float4 *__restrict__ out) {
    const long long k = blockIdx.x;
    const bool zero = zero_mask && zero_mask[k];
    const float4 *src = ring + idx[k] * slot_stride4;
    float4 *dst = out + k * (OBS_FLOATS / 4);
    for (int q = threadIdx.x; q < OBS_FLOATS / 4; q += 256)
        dst[q] = zero ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldg(src + q);
}

__global__ void k_clone(StateView dst, long long di, StateView src, long long si) {
    const int t = threadIdx.x;
    if (t == 0) dst.core0[di] = src.core0[si];
    if (t == 1) dst.core1[di] = src.core1[si];
    if (t == 2) dst.ghosts[di] = src.ghosts[si];
    if (t == 3) dst.ghosts[dst.npad + di] = src.ghosts[src.npad + si];
    if (t >= 4 && t < 4 + PEL_WORDS)
        dst.pel[(long long)(t - 4) * dst.npad + di] = src.pel[(long long)(t - 4) * src.npad + si];
}

}  // namespace pb

#include "pb_mcts.cuh"

// =============================================================================================
// Host side: the C ABI
// =============================================================================================
using namespace pb;

static thread_local std::string g_err;
static int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
#define PB_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (call);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return fail(PB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e));      \
    } while (0)

struct pb_engine {
    int device = 0;
    long long n = 0, npad = 0;
    uint64_t seed = 0;
    long long gid_base = 0;
    StateView st{};
    StateView st_alt{};  // second state buffer, allocated by the first pb_step_random_flip
    const uint32_t *tape = nullptr;
    long long tape_len = 0;
    unsigned long long *bad_action = nullptr;  // device
    // staging for the host-buffer entry points
    uint8_t *stg_actions = nullptr;
    int32_t *stg_reward = nullptr;
    uint8_t *stg_done = nullptr;
    uint8_t *stg_mask = nullptr;
    // pb_obs_host: two device staging buffers (encode chunk k+1 while chunk k crosses PCIe) and,
    // for pageable destinations, two pinned bounce buffers
    float *stg_obs[2] = {nullptr, nullptr};
    float *pin_obs[2] = {nullptr, nullptr};
    long long stg_obs_envs = 0, pin_obs_envs = 0;
    cudaEvent_t ev_obs_enc[2] = {nullptr, nullptr}, ev_obs_copy[2] = {nullptr, nullptr};

    int num_sms = 148;
    unsigned int *tickets = nullptr;  // ring of work counters for k_encode_obs launches
    unsigned int ticket_next = 0;
    // pb_step_host: results travel device->host on a side stream while the encoder runs
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_step = nullptr, ev_host_copied = nullptr;
    unsigned long long *host_bad = nullptr;  // pinned: invalid-action record of the step in flight
    bool host_step_pending = false;
    // pb_rollout_*: two internal streams, double-buffered state (created on first use)
    // two encode streams used alternately: the first CTAs of encoder t+1 start on the SMs that
    // encoder t has already left (its tail and the next launch's prologue overlap)
    cudaStream_t roll_step = nullptr, roll_enc[2] = {nullptr, nullptr};
    int roll_enc_streams = 2;
    const float *roll_last_obs = nullptr;   // output range of the previous rollout step's encoder
    size_t roll_last_obs_bytes = 0;
    cudaEvent_t ev_roll_step = nullptr, ev_roll_fork = nullptr, ev_roll_join = nullptr;
    cudaEvent_t ev_enc_done[2] = {nullptr, nullptr};
    unsigned long long roll_t = 0;
    bool roll_started = false;
    // pb_rollout_chunk_*: in-place chunked rollout; one record per encoder launch whose env range
    // a later in-place step of the same envs has to wait for
    struct RollChunk {
        long long first, count;     // env range the encoder reads
        const char *out_lo;         // output range it writes
        size_t out_bytes;
        int stream;                 // index into roll_enc
        cudaEvent_t enc_done;
    };
    std::vector<RollChunk> roll_chunks;
    unsigned long long roll_chunk_ctr = 0;
    int roll_last_enc = 0;                   // roll_enc index of the most recent encoder launch
    cudaEvent_t ev_roll_copied = nullptr;    // last result copy of a host chunk
    bool roll_host_pending = false;
    // pb_step_snapshot_host: the pinned block k_gym_call fills in place, pinned actions it reads
    // in place (zero copy both ways), the device counter its CTAs arrive at, the call number
    unsigned char *snap_pin = nullptr;
    uint8_t *snap_actions_pin = nullptr;
    unsigned int *snap_arrive = nullptr;
    unsigned long long snap_seq = 0;
    // pb_trace_*: device timeline of this engine's k_step / k_encode_obs launches
    TraceRec *trace = nullptr;
    long long trace_cap = 0, trace_n = 0;
    std::vector<uint32_t> trace_kinds;
};
constexpr unsigned int TICKET_RING = 64;

namespace {
inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#elif defined(__aarch64__)
    asm volatile("yield" ::: "memory");
#endif
}
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};

RngParams rng_params(const pb_engine *e) {
    return RngParams{e->seed, e->gid_base, e->tape, e->tape_len};
}
unsigned grid_for(long long n) { return (unsigned)((n + STEP_THREADS - 1) / STEP_THREADS); }

// the next record of the engine's device timeline, or nullptr when tracing is off / full
TraceRec *trace_slot(pb_engine *e, uint32_t kind) {
    if (!e->trace || e->trace_n >= e->trace_cap) return nullptr;
    e->trace_kinds[(size_t)e->trace_n] = kind;
    return e->trace + e->trace_n++;
}

void flat_to_rows(const uint32_t *f, uint32_t rows[PB_ROWS]) {
    for (int r = 0; r < PB_ROWS; r++) rows[r] = 0;
    for (int r = 0; r < ROWS; r++)
        for (int c = 0; c < COLS; c++) {
            const int i = fbit(r, c);
            if ((f[i >> 5] >> (i & 31)) & 1u) rows[r] |= 1u << c;
        }
}
void rows_to_flat(const uint32_t rows[PB_ROWS], uint32_t *f) {
    for (int w = 0; w < PEL_WORDS; w++) f[w] = 0;
    for (int r = 0; r < ROWS; r++)
        for (int c = 0; c < COLS; c++)
            if ((rows[r] >> c) & 1u) {
                const int i = fbit(r, c);
                f[i >> 5] |= 1u << (i & 31);
            }
}

int encode_attrs(pb_engine *e) {
    static thread_local int attr_set_for = -1;
    if (attr_set_for != e->device) {
        PB_CUDA(cudaFuncSetAttribute(k_encode_obs, cudaFuncAttributeMaxDynamicSharedMemorySize, ENC_SMEM));
        PB_CUDA(cudaFuncSetAttribute(k_obs_small, cudaFuncAttributeMaxDynamicSharedMemorySize, SMALL_SMEM));
        PB_CUDA(cudaFuncSetAttribute(k_gym_call, cudaFuncAttributeMaxDynamicSharedMemorySize, SMALL_SMEM));
        PB_CUDA(cudaFuncSetAttribute(k_step_obs_small, cudaFuncAttributeMaxDynamicSharedMemorySize, SMALL_SMEM));
        // the kernel in front of a small encode is k_step, which runs under the largest carve-out
        // (see pb_create): ask for the same one, so that the SMs are not reconfigured in between.
        // PB_SMALL_CARVEOUT=0 leaves the driver's choice (A/B runs).
        const char *cv = getenv("PB_SMALL_CARVEOUT");
        if (!(cv && cv[0] == '0'))
            PB_CUDA(cudaFuncSetAttribute(k_obs_small, cudaFuncAttributePreferredSharedMemoryCarveout,
                                         cudaSharedmemCarveoutMaxShared));
        attr_set_for = e->device;
    }
    return PB_OK;
}

// `nhwc_out` (small launches only, see pb_encode_obs_ring_nhwc32): the same observations also in
// the padded channels-last layout, from the same launch
int launch_encode(pb_engine *e, float *out_dev, long long stride, long long first,
                  long long count, cudaStream_t s, const long long *slot_index = nullptr,
                  long long slot_stride = 0, float *nhwc_out = nullptr) {
    if (count <= 0) return PB_OK;
    {
        const int rc = encode_attrs(e);
        if (rc != PB_OK) return rc;
    }
    // Small batches (the DQN loop's 32 envs, an evaluation env, the leaves of a few hundred MCTS
    // trees): one CTA per env instead of the persistent kernel, whose prologue alone costs more
    // than these launches (pb_encode_small.cuh).  PB_SMALL_ENC_MAX overrides the switch-over
    // count for A/B runs (0: always the persistent kernel).
    static const long long small_max = [] {
        const char *v = getenv("PB_SMALL_ENC_MAX");
        return v ? atoll(v) : (long long)SMALL_ENC_MAX;
    }();
    if (count <= small_max || nhwc_out) {
        TraceRec *trace = trace_slot(e, PB_TRACE_ENCODE);
        k_obs_small<<<(unsigned)count, OBS_SMALL_THREADS, SMALL_SMEM, s>>>(
            e->st, out_dev, stride, first, slot_index, slot_stride, nhwc_out, trace);
        PB_CUDA(cudaGetLastError());
        return PB_OK;
    }
    long long blocks = (count + ENC_WARPS - 1) / ENC_WARPS;
    if (blocks > e->num_sms) blocks = e->num_sms;
    // {ticket, finished-warp count} pair; the last warp of a launch zeroes both again, so no
    // memset node is needed per launch (the ring only matters for launches on different streams)
    unsigned int *ticket = e->tickets + 2 * (e->ticket_next++ % TICKET_RING);
    // Programmatic dependent launch: the encoder's prologue (zeroing 222 KB of shared memory,
    // wall template) overlaps the tail of the preceding kernel in the stream (k_step); the
    // kernel executes griddepcontrol.wait before it touches the env state or the ticket.
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)blocks);
    cfg.blockDim = dim3(ENC_THREADS);
    cfg.dynamicSmemBytes = ENC_SMEM;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    // PB_NO_PDL=1 launches the encoder with plain stream ordering (A/B runs)
    static const bool no_pdl = [] {
        const char *v = getenv("PB_NO_PDL");
        return v && v[0] == '1';
    }();
    cfg.numAttrs = no_pdl ? 0 : 1;
    TraceRec *trace = trace_slot(e, PB_TRACE_ENCODE);
    PB_CUDA(cudaLaunchKernelEx(&cfg, k_encode_obs, e->st, out_dev, stride, first, count, ticket,
                               slot_index, slot_stride, trace));
    return PB_OK;
}
}  // namespace

extern "C" {

int pb_version(void) { return PB_VERSION; }
const char *pb_last_error(void) { return g_err.c_str(); }
size_t pb_sizeof_env_state(void) { return sizeof(pb_env_state); }
int pb_state_bytes_per_env(void) { return STATE_BYTES_PER_ENV; }
uint32_t pb_draw_raw(uint64_t seed, uint64_t env_gid, uint32_t ctr) {
    return draw_raw(seed, env_gid, ctr);
}
uint32_t pb_action_raw(uint64_t seed, uint64_t env_gid, uint64_t call_idx) {
    return action_raw(seed, env_gid, call_idx);
}

uint32_t pb_maze_wall_row(int row) { return (row >= 0 && row < ROWS) ? H_TABLES.walls[row] : 0u; }
uint32_t pb_maze_pellet_row(int row) {
    if (row < 0 || row >= ROWS) return 0u;
    uint32_t rows[PB_ROWS];
    flat_to_rows(H_TABLES.init_pel, rows);
    return rows[row];
}
int pb_maze_node(int k, int *row, int *col) {
    if (k < 0 || k >= NUM_NODES || !row || !col)
        return fail(PB_ERR_INVALID_ARG, "pb_maze_node: node index out of range or null output");
    *row = H_TABLES.node_rc[k] & 0xff;
    *col = H_TABLES.node_rc[k] >> 8;
    return PB_OK;
}

int pb_destroy(pb_engine *e) {
    if (!e) return PB_OK;
    DeviceGuard g(e->device);
    cudaFree(e->st.core0);
    cudaFree(e->st.core1);
    cudaFree(e->st.ghosts);
    cudaFree(e->st.pel);
    cudaFree(e->st_alt.core0);
    cudaFree(e->st_alt.core1);
    cudaFree(e->st_alt.ghosts);
    cudaFree(e->st_alt.pel);
    cudaFree(e->bad_action);
    cudaFree(e->stg_actions);
    cudaFree(e->stg_reward);
    cudaFree(e->stg_done);
    cudaFree(e->stg_mask);
    for (int b = 0; b < 2; b++) {
        cudaFree(e->stg_obs[b]);
        if (e->pin_obs[b]) cudaFreeHost(e->pin_obs[b]);
        if (e->ev_obs_enc[b]) cudaEventDestroy(e->ev_obs_enc[b]);
        if (e->ev_obs_copy[b]) cudaEventDestroy(e->ev_obs_copy[b]);
    }
    cudaFree(e->tickets);
    cudaFree(e->trace);
    cudaFree(e->snap_arrive);
    if (e->snap_pin) cudaFreeHost(e->snap_pin);
    if (e->snap_actions_pin) cudaFreeHost(e->snap_actions_pin);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    if (e->ev_step) cudaEventDestroy(e->ev_step);
    if (e->ev_host_copied) cudaEventDestroy(e->ev_host_copied);
    if (e->host_bad) cudaFreeHost(e->host_bad);
    if (e->roll_step) cudaStreamDestroy(e->roll_step);
    for (cudaStream_t s : e->roll_enc)
        if (s) cudaStreamDestroy(s);
    for (auto &r : e->roll_chunks) cudaEventDestroy(r.enc_done);
    for (cudaEvent_t ev : {e->ev_roll_step, e->ev_roll_fork, e->ev_roll_join, e->ev_enc_done[0],
                           e->ev_enc_done[1], e->ev_roll_copied})
        if (ev) cudaEventDestroy(ev);
    delete e;
    return PB_OK;
}

int pb_create(pb_engine **out, int64_t n_envs, int device, uint64_t seed, int64_t env_id_base,
              const uint8_t *cfg_flags_host) {
    if (!out || n_envs <= 0) return fail(PB_ERR_INVALID_ARG, "pb_create: bad arguments");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(PB_ERR_NO_DEVICE, "pb_create: no CUDA device (this engine has no CPU path)");
    }
    if (device < 0 || device >= ndev) return fail(PB_ERR_INVALID_ARG, "pb_create: bad device");
    DeviceGuard g(device);
    if (!g.ok) return fail(PB_ERR_CUDA, "pb_create: cudaSetDevice failed");
    pb_engine *e = new pb_engine();
    e->device = device;
    e->n = n_envs;
    e->npad = (n_envs + 31) / 32 * 32;
    e->seed = seed;
    e->gid_base = env_id_base;
    e->st.n = e->n;
    e->st.npad = e->npad;
#define PB_CUDA_OR_FREE(call)                                                                  \
    do {                                                                                       \
        cudaError_t _e = (call);                                                               \
        if (_e != cudaSuccess) {                                                               \
            pb_destroy(e);                                                                     \
            return fail(PB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e));      \
        }                                                                                      \
    } while (0)
    PB_CUDA_OR_FREE(cudaDeviceGetAttribute(&e->num_sms, cudaDevAttrMultiProcessorCount, device));
    PB_CUDA_OR_FREE(cudaMemcpyToSymbol(c_tab, &H_TABLES, sizeof(MazeTables)));
    // k_step shares the SMs with the observation encoder (pb_rollout_*), whose CTA needs the
    // LARGEST shared-memory carve-out (225 KB).  An SM running k_step CTAs under a small carve-out
    // cannot take an encoder CTA until they have drained and the SM is reconfigured: whenever
    // k_step reached the SMs first, the encoder started 15 us late on 128 of 148 SMs (device
    // timeline, profiles/r02_timeline_*: encoder 524 us instead of 514 us).  With the same
    // carve-out preference on both kernels the launch order no longer matters.
    // PB_STEP_CARVEOUT=0 restores the default (A/B runs).
    {
        const char *v = getenv("PB_STEP_CARVEOUT");
        if (!(v && v[0] == '0'))
            PB_CUDA_OR_FREE(cudaFuncSetAttribute(k_step, cudaFuncAttributePreferredSharedMemoryCarveout,
                                                 cudaSharedmemCarveoutMaxShared));
    }
    PB_CUDA_OR_FREE(cudaMalloc(&e->st.core0, sizeof(uint4) * e->npad));
    PB_CUDA_OR_FREE(cudaMalloc(&e->st.core1, sizeof(uint4) * e->npad));
    PB_CUDA_OR_FREE(cudaMalloc(&e->st.ghosts, sizeof(uint4) * 2 * e->npad));
    PB_CUDA_OR_FREE(cudaMalloc(&e->st.pel, sizeof(uint32_t) * PEL_WORDS * e->npad));
    PB_CUDA_OR_FREE(cudaMalloc(&e->bad_action, sizeof(unsigned long long)));
    PB_CUDA_OR_FREE(cudaMalloc(&e->stg_actions, e->npad));
    PB_CUDA_OR_FREE(cudaMalloc(&e->stg_reward, sizeof(int32_t) * e->npad));
    PB_CUDA_OR_FREE(cudaMalloc(&e->stg_done, e->npad));
    PB_CUDA_OR_FREE(cudaMalloc(&e->stg_mask, 5 * e->npad));
    PB_CUDA_OR_FREE(cudaMalloc(&e->tickets, sizeof(unsigned int) * 2 * TICKET_RING));
    PB_CUDA_OR_FREE(cudaMemset(e->tickets, 0, sizeof(unsigned int) * 2 * TICKET_RING));
    PB_CUDA_OR_FREE(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
    PB_CUDA_OR_FREE(cudaEventCreateWithFlags(&e->ev_step, cudaEventDisableTiming));
    PB_CUDA_OR_FREE(cudaEventCreateWithFlags(&e->ev_host_copied, cudaEventDisableTiming));
    PB_CUDA_OR_FREE(cudaMemset(e->st.core0, 0, sizeof(uint4) * e->npad));
    PB_CUDA_OR_FREE(cudaMemset(e->st.core1, 0, sizeof(uint4) * e->npad));
    PB_CUDA_OR_FREE(cudaMemset(e->st.ghosts, 0, sizeof(uint4) * 2 * e->npad));
    PB_CUDA_OR_FREE(cudaMemset(e->st.pel, 0, sizeof(uint32_t) * PEL_WORDS * e->npad));
    PB_CUDA_OR_FREE(cudaMemset(e->bad_action, 0xff, sizeof(unsigned long long)));
    // config flags live in core1.z bits 8..15; write them, then run the constructor path
    {
        std::vector<uint4> c1((size_t)e->npad, make_uint4(0, 0, 0, 0));
        for (long long i = 0; i < e->n; i++)
            c1[(size_t)i].z = (uint32_t)(cfg_flags_host ? (cfg_flags_host[i] & 0x0f) : 0) << 8;
        PB_CUDA_OR_FREE(cudaMemcpy(e->st.core1, c1.data(), sizeof(uint4) * e->npad,
                                   cudaMemcpyHostToDevice));
    }
    k_reset<<<grid_for(e->n), STEP_THREADS>>>(e->st, rng_params(e), nullptr, 0);
    PB_CUDA_OR_FREE(cudaGetLastError());
    PB_CUDA_OR_FREE(cudaDeviceSynchronize());
#undef PB_CUDA_OR_FREE
    *out = e;
    return PB_OK;
}

int64_t pb_num_envs(const pb_engine *e) { return e ? e->n : 0; }
int pb_device(const pb_engine *e) { return e ? e->device : -1; }

int pb_set_cfg_flags(pb_engine *e, int64_t first, int64_t count, const uint8_t *cfg) {
    if (!e || !cfg || first < 0 || count < 0 || first + count > e->n)
        return fail(PB_ERR_INVALID_ARG, "pb_set_cfg_flags: bad arguments");
    if (count == 0) return PB_OK;
    DeviceGuard g(e->device);
    std::vector<uint4> c1((size_t)count);
    PB_CUDA(cudaMemcpy(c1.data(), e->st.core1 + first, sizeof(uint4) * count,
                       cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < count; i++)
        c1[(size_t)i].z = (c1[(size_t)i].z & ~0xff00u) | ((uint32_t)(cfg[i] & 0x0f) << 8);
    PB_CUDA(cudaMemcpy(e->st.core1 + first, c1.data(), sizeof(uint4) * count,
                       cudaMemcpyHostToDevice));
    return PB_OK;
}

int pb_set_draw_tape(pb_engine *e, const uint32_t *tape_dev, int64_t tape_len) {
    if (!e || (tape_dev && tape_len <= 0) || tape_len > 0x7fffffffLL)
        return fail(PB_ERR_INVALID_ARG, "pb_set_draw_tape: bad arguments");
    e->tape = tape_dev;
    e->tape_len = tape_dev ? tape_len : 0;
    return PB_OK;
}

int pb_reset(pb_engine *e, const uint8_t *mask_dev, void *stream) {
    if (!e) return fail(PB_ERR_INVALID_ARG, "pb_reset: null engine");
    DeviceGuard g(e->device);
    k_reset<<<grid_for(e->n), STEP_THREADS, 0, (cudaStream_t)stream>>>(e->st, rng_params(e),
                                                                       mask_dev, 1);
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int pb_step(pb_engine *e, const uint8_t *actions_dev, int32_t *reward_dev, uint8_t *done_dev,
            uint8_t *mask_out_dev, int auto_reset, void *stream) {
    if (!e || !actions_dev || !reward_dev || !done_dev)
        return fail(PB_ERR_INVALID_ARG, "pb_step: null argument");
    DeviceGuard g(e->device);
    k_step<<<grid_for(e->n), STEP_THREADS, 0, (cudaStream_t)stream>>>(
        e->st, e->st, rng_params(e), actions_dev, reward_dev, done_dev, mask_out_dev, auto_reset,
        e->bad_action, nullptr, 0ull, 0ll, e->n, trace_slot(e, PB_TRACE_STEP));
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int pb_step_random(pb_engine *e, uint64_t call_idx, uint8_t *actions_out_dev, int32_t *reward_dev,
                   uint8_t *done_dev, uint8_t *mask_out_dev, int auto_reset, void *stream) {
    if (!e || !reward_dev || !done_dev)
        return fail(PB_ERR_INVALID_ARG, "pb_step_random: null argument");
    return pb_step_random_range(e, 0, e->n, call_idx, actions_out_dev, reward_dev, done_dev,
                                mask_out_dev, auto_reset, stream);
}

int pb_step_random_range(pb_engine *e, int64_t first, int64_t count, uint64_t call_idx,
                         uint8_t *actions_out_dev, int32_t *reward_dev, uint8_t *done_dev,
                         uint8_t *mask_out_dev, int auto_reset, void *stream) {
    if (!e || !reward_dev || !done_dev || first < 0 || count < 0 || first + count > e->n)
        return fail(PB_ERR_INVALID_ARG, "pb_step_random_range: bad arguments");
    if (count == 0) return PB_OK;
    DeviceGuard g(e->device);
    k_step<<<grid_for(count), STEP_THREADS, 0, (cudaStream_t)stream>>>(
        e->st, e->st, rng_params(e), nullptr, reward_dev, done_dev, mask_out_dev, auto_reset,
        e->bad_action, actions_out_dev, call_idx, first, count, trace_slot(e, PB_TRACE_STEP));
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int pb_step_random_flip(pb_engine *e, uint64_t call_idx, uint8_t *actions_out_dev,
                        int32_t *reward_dev, uint8_t *done_dev, uint8_t *mask_out_dev,
                        int auto_reset, void *stream) {
    return pb_step_flip(e, nullptr, call_idx, actions_out_dev, reward_dev, done_dev, mask_out_dev,
                        auto_reset, stream);
}

int pb_step_flip(pb_engine *e, const uint8_t *actions_dev, uint64_t call_idx,
                 uint8_t *actions_out_dev, int32_t *reward_dev, uint8_t *done_dev,
                 uint8_t *mask_out_dev, int auto_reset, void *stream) {
    if (!e || !reward_dev || !done_dev)
        return fail(PB_ERR_INVALID_ARG, "pb_step_flip: null argument");
    DeviceGuard g(e->device);
    if (!e->st_alt.pel) {  // first use: the second state buffer (same layout)
        StateView a{};
        a.npad = e->st.npad;
        a.n = e->st.n;
        cudaError_t err = cudaMalloc(&a.core0, sizeof(uint4) * e->npad);
        if (err == cudaSuccess) err = cudaMalloc(&a.core1, sizeof(uint4) * e->npad);
        if (err == cudaSuccess) err = cudaMalloc(&a.ghosts, sizeof(uint4) * 2 * e->npad);
        if (err == cudaSuccess) err = cudaMalloc(&a.pel, sizeof(uint32_t) * PEL_WORDS * e->npad);
        if (err != cudaSuccess) {
            cudaFree(a.core0);
            cudaFree(a.core1);
            cudaFree(a.ghosts);
            cudaFree(a.pel);
            return fail(PB_ERR_CUDA, std::string("pb_step_random_flip: ") + cudaGetErrorString(err));
        }
        e->st_alt = a;
    }
    const unsigned wide_grid = (unsigned)((e->n + STEP_THREADS_WIDE - 1) / STEP_THREADS_WIDE);
    k_step<<<wide_grid, STEP_THREADS_WIDE, 0, (cudaStream_t)stream>>>(
        e->st, e->st_alt, rng_params(e), actions_dev, reward_dev, done_dev, mask_out_dev, auto_reset,
        e->bad_action, actions_out_dev, call_idx, 0ll, e->n, trace_slot(e, PB_TRACE_STEP));
    PB_CUDA(cudaGetLastError());
    // from here on "the state" is the buffer just written; launches already enqueued keep the
    // StateView they captured
    std::swap(e->st, e->st_alt);
    return PB_OK;
}

int pb_encode_obs_range(pb_engine *e, int64_t first, int64_t count, float *out_dev,
                        int64_t env_stride_floats, void *stream) {
    if (!e || !out_dev || first < 0 || count < 0 || first + count > e->n)
        return fail(PB_ERR_INVALID_ARG, "pb_encode_obs: bad arguments");
    if (env_stride_floats == 0) env_stride_floats = OBS_FLOATS;
    if (env_stride_floats < OBS_FLOATS || (env_stride_floats & 3) ||
        (reinterpret_cast<uintptr_t>(out_dev) & 15))
        return fail(PB_ERR_INVALID_ARG,
                    "pb_encode_obs: stride must be >= 14756 and a multiple of 4 floats, "
                    "out must be 16-byte aligned");
    DeviceGuard g(e->device);
    return launch_encode(e, out_dev, env_stride_floats, first, count, (cudaStream_t)stream);
}

int pb_encode_obs(pb_engine *e, float *out_dev, int64_t env_stride_floats, void *stream) {
    if (!e) return fail(PB_ERR_INVALID_ARG, "pb_encode_obs: null engine");
    return pb_encode_obs_range(e, 0, e->n, out_dev, env_stride_floats, stream);
}

int pb_encode_obs_nhwc32(pb_engine *e, int64_t first, int64_t count, float *out_dev, void *stream) {
    if (!e || !out_dev || first < 0 || count < 0 || first + count > e->n ||
        (reinterpret_cast<uintptr_t>(out_dev) & 15))
        return fail(PB_ERR_INVALID_ARG, "pb_encode_obs_nhwc32: bad arguments (out must be 16-byte aligned)");
    if (count == 0) return PB_OK;
    DeviceGuard g(e->device);
    const int rc = encode_attrs(e);
    if (rc != PB_OK) return rc;
    k_obs_small<<<(unsigned)count, OBS_SMALL_THREADS, SMALL_SMEM, (cudaStream_t)stream>>>(
        e->st, nullptr, 0ll, first, nullptr, 0ll, out_dev, trace_slot(e, PB_TRACE_ENCODE));
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int pb_encode_obs_ring(pb_engine *e, float *ring_dev, int64_t slot_stride_floats,
                       const int64_t *slot_index_dev, void *stream) {
    if (!e || !ring_dev || !slot_index_dev || (slot_stride_floats & 3) ||
        slot_stride_floats < (int64_t)OBS_FLOATS * e->n || (reinterpret_cast<uintptr_t>(ring_dev) & 15))
        return fail(PB_ERR_INVALID_ARG, "pb_encode_obs_ring: bad arguments");
    DeviceGuard g(e->device);
    return launch_encode(e, ring_dev, OBS_FLOATS, 0, e->n, (cudaStream_t)stream,
                         reinterpret_cast<const long long *>(slot_index_dev), slot_stride_floats);
}

int pb_encode_obs_ring_nhwc32(pb_engine *e, float *ring_dev, int64_t slot_stride_floats,
                              const int64_t *slot_index_dev, float *nhwc_out_dev, void *stream) {
    if (!e || !ring_dev || !slot_index_dev || !nhwc_out_dev || (slot_stride_floats & 3) ||
        slot_stride_floats < (int64_t)OBS_FLOATS * e->n || (reinterpret_cast<uintptr_t>(ring_dev) & 15) ||
        (reinterpret_cast<uintptr_t>(nhwc_out_dev) & 15))
        return fail(PB_ERR_INVALID_ARG, "pb_encode_obs_ring_nhwc32: bad arguments");
    DeviceGuard g(e->device);
    return launch_encode(e, ring_dev, OBS_FLOATS, 0, e->n, (cudaStream_t)stream,
                         reinterpret_cast<const long long *>(slot_index_dev), slot_stride_floats,
                         nhwc_out_dev);
}

int pb_step_encode(pb_engine *e, const uint8_t *actions_dev, int32_t *reward_dev,
                   uint8_t *done_dev, uint8_t *mask_out_dev, int auto_reset, float *obs_out_dev,
                   int64_t env_stride_floats, void *stream) {
    // a small engine: one launch, every env on a warp of its own (pb_step_encode_small)
    if (e && actions_dev && reward_dev && done_dev && obs_out_dev && e->n <= SMALL_ENC_MAX)
        return pb_step_encode_small(e, actions_dev, reward_dev, done_dev, mask_out_dev, auto_reset, obs_out_dev,
                                    env_stride_floats, nullptr, 0, nullptr, stream);
    int rc = pb_step(e, actions_dev, reward_dev, done_dev, mask_out_dev, auto_reset, stream);
    if (rc != PB_OK) return rc;
    return pb_encode_obs(e, obs_out_dev, env_stride_floats, stream);
}

int pb_step_encode_small(pb_engine *e, const uint8_t *actions_dev, int32_t *reward_dev, uint8_t *done_dev,
                         uint8_t *mask_out_dev, int auto_reset, float *obs_out_dev, int64_t env_stride_floats,
                         const int64_t *slot_index_dev, int64_t slot_stride_floats, float *nhwc_out_dev,
                         void *stream) {
    if (!e || !actions_dev || !reward_dev || !done_dev)
        return fail(PB_ERR_INVALID_ARG, "pb_step_encode_small: null argument");
    if (obs_out_dev && ((env_stride_floats & 3) || env_stride_floats < OBS_FLOATS ||
                        (reinterpret_cast<uintptr_t>(obs_out_dev) & 15) || (slot_stride_floats & 3)))
        return fail(PB_ERR_INVALID_ARG, "pb_step_encode_small: strides must be multiples of 4 floats (>= 14756), "
                                        "outputs 16-byte aligned");
    if (nhwc_out_dev && (reinterpret_cast<uintptr_t>(nhwc_out_dev) & 15))
        return fail(PB_ERR_INVALID_ARG, "pb_step_encode_small: nhwc_out must be 16-byte aligned");
    DeviceGuard g(e->device);
    const int rc = encode_attrs(e);
    if (rc != PB_OK) return rc;
    k_step_obs_small<<<(unsigned)e->n, GYM_CALL_THREADS, SMALL_SMEM, (cudaStream_t)stream>>>(
        e->st, rng_params(e), actions_dev, reward_dev, done_dev, mask_out_dev, auto_reset, e->bad_action,
        obs_out_dev, env_stride_floats, reinterpret_cast<const long long *>(slot_index_dev),
        slot_stride_floats, nhwc_out_dev, trace_slot(e, PB_TRACE_STEP));
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int pb_action_mask(pb_engine *e, uint8_t *mask_out_dev, void *stream) {
    if (!e || !mask_out_dev) return fail(PB_ERR_INVALID_ARG, "pb_action_mask: null argument");
    DeviceGuard g(e->device);
    k_mask<<<grid_for(e->n), STEP_THREADS, 0, (cudaStream_t)stream>>>(e->st, mask_out_dev);
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int pb_query(pb_engine *e, int what, int32_t *out_dev, void *stream) {
    if (!e || !out_dev || what < 0 || what >= PB_Q_COUNT)
        return fail(PB_ERR_INVALID_ARG, "pb_query: bad arguments");
    DeviceGuard g(e->device);
    k_query<<<grid_for(e->n), STEP_THREADS, 0, (cudaStream_t)stream>>>(e->st, what, out_dev);
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int pb_eval_track(pb_engine *e, const uint8_t *done_dev, int32_t *stats_dev, void *stream) {
    if (!e || !done_dev || !stats_dev) return fail(PB_ERR_INVALID_ARG, "pb_eval_track: null argument");
    DeviceGuard g(e->device);
    k_eval_track<<<grid_for(e->n), STEP_THREADS, 0, (cudaStream_t)stream>>>(e->st, done_dev, stats_dev);
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int pb_sample_actions(pb_engine *e, uint8_t *actions_dev, uint64_t call_idx, void *stream) {
    if (!e || !actions_dev) return fail(PB_ERR_INVALID_ARG, "pb_sample_actions: null argument");
    DeviceGuard g(e->device);
    k_sample_actions<<<grid_for(e->n), STEP_THREADS, 0, (cudaStream_t)stream>>>(
        e->st, rng_params(e), actions_dev, call_idx);
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int pb_select_actions(pb_engine *e, const float *q_dev, const uint8_t *mask_dev, float epsilon,
                      uint8_t *actions_dev, uint64_t call_idx, void *stream) {
    if (!e || !q_dev || !actions_dev)
        return fail(PB_ERR_INVALID_ARG, "pb_select_actions: null argument");
    DeviceGuard g(e->device);
    k_select_actions<<<grid_for(e->n), STEP_THREADS, 0, (cudaStream_t)stream>>>(
        e->st, rng_params(e), q_dev, mask_dev, epsilon, actions_dev, call_idx, nullptr, nullptr);
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int pb_select_actions_dev(pb_engine *e, const float *q_dev, const uint8_t *mask_dev,
                          const float *epsilon_dev, const int64_t *call_idx_dev,
                          uint8_t *actions_dev, void *stream) {
    if (!e || !q_dev || !actions_dev || !epsilon_dev || !call_idx_dev)
        return fail(PB_ERR_INVALID_ARG, "pb_select_actions_dev: null argument");
    DeviceGuard g(e->device);
    k_select_actions<<<grid_for(e->n), STEP_THREADS, 0, (cudaStream_t)stream>>>(
        e->st, rng_params(e), q_dev, mask_dev, 0.0f, actions_dev, 0ull, epsilon_dev,
        reinterpret_cast<const long long *>(call_idx_dev));
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int pb_check_actions(pb_engine *e, int64_t *first_bad_env, void *stream) {
    if (!e) return fail(PB_ERR_INVALID_ARG, "pb_check_actions: null engine");
    DeviceGuard g(e->device);
    unsigned long long bad = ~0ull;
    cudaStream_t s = (cudaStream_t)stream;
    PB_CUDA(cudaMemcpyAsync(&bad, e->bad_action, sizeof(bad), cudaMemcpyDeviceToHost, s));
    PB_CUDA(cudaStreamSynchronize(s));
    if (bad == ~0ull) return PB_OK;
    PB_CUDA(cudaMemsetAsync(e->bad_action, 0xff, sizeof(bad), s));
    if (first_bad_env) *first_bad_env = (int64_t)bad;
    return fail(PB_ERR_INVALID_ACTION, "Invalid action");
}

// Work of the engine's own streams (pipelined rollout steps, result copies) that a state access
// ordered on the caller's stream has to wait for.
static int sync_engine_streams(pb_engine *e) {
    for (cudaStream_t s : {e->roll_step, e->roll_enc[0], e->roll_enc[1], e->copy_stream})
        if (s) PB_CUDA(cudaStreamSynchronize(s));
    return PB_OK;
}

// whole_device: pb_get_state (no stream argument: waits for the whole device).  Otherwise only the
// caller's stream and the engine's own streams are synchronised, and the copies are ordered on the
// caller's stream: other streams of the process (a training graph, another engine) keep running.
static int get_state_impl(pb_engine *e, int64_t first, int64_t count, pb_env_state *out,
                          bool whole_device, cudaStream_t s) {
    if (!e || !out || first < 0 || count < 0 || first + count > e->n)
        return fail(PB_ERR_INVALID_ARG, "pb_get_state: bad arguments");
    if (count == 0) return PB_OK;
    DeviceGuard g(e->device);
    if (whole_device) {
        PB_CUDA(cudaDeviceSynchronize());
        s = nullptr;
    } else {
        int rc = sync_engine_streams(e);
        if (rc != PB_OK) return rc;
    }
    const size_t c = (size_t)count;
    std::vector<uint4> c0(c), c1(c), ga(c), gb(c);
    std::vector<uint32_t> pel(c * PEL_WORDS);
    PB_CUDA(cudaMemcpyAsync(c0.data(), e->st.core0 + first, sizeof(uint4) * c, cudaMemcpyDeviceToHost, s));
    PB_CUDA(cudaMemcpyAsync(c1.data(), e->st.core1 + first, sizeof(uint4) * c, cudaMemcpyDeviceToHost, s));
    PB_CUDA(cudaMemcpyAsync(ga.data(), e->st.ghosts + first, sizeof(uint4) * c, cudaMemcpyDeviceToHost, s));
    PB_CUDA(cudaMemcpyAsync(gb.data(), e->st.ghosts + e->npad + first, sizeof(uint4) * c,
                            cudaMemcpyDeviceToHost, s));
    PB_CUDA(cudaMemcpy2DAsync(pel.data(), sizeof(uint32_t) * c, e->st.pel + first,
                              sizeof(uint32_t) * e->npad, sizeof(uint32_t) * c, PEL_WORDS,
                              cudaMemcpyDeviceToHost, s));
    PB_CUDA(cudaStreamSynchronize(s));
    for (size_t i = 0; i < c; i++) {
        EnvR s;
        unpack_env(c0[i], c1[i], ga[i], gb[i], s);
        pb_env_state &o = out[i];
        memset(&o, 0, sizeof(o));
        o.curr_ticks = s.ticks;
        o.rng_ctr = s.rng_ctr;
        uint32_t f[PEL_WORDS];
        for (int w = 0; w < PEL_WORDS; w++) f[w] = pel[(size_t)w * c + i];
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
    return PB_OK;
}

int pb_get_state(pb_engine *e, int64_t first, int64_t count, pb_env_state *out) {
    return get_state_impl(e, first, count, out, true, nullptr);
}
int pb_get_state_on(pb_engine *e, int64_t first, int64_t count, pb_env_state *out, void *stream) {
    return get_state_impl(e, first, count, out, false, (cudaStream_t)stream);
}

static int set_state_impl(pb_engine *e, int64_t first, int64_t count, const pb_env_state *in,
                          bool whole_device, cudaStream_t stream) {
    if (!e || !in || first < 0 || count < 0 || first + count > e->n)
        return fail(PB_ERR_INVALID_ARG, "pb_set_state: bad arguments");
    if (count == 0) return PB_OK;
    const size_t c = (size_t)count;
    std::vector<uint4> c0(c), c1(c), ga(c), gb(c);
    std::vector<uint32_t> pel(c * PEL_WORDS);
    for (size_t i = 0; i < c; i++) {
        const pb_env_state &o = in[i];
        if (o.update_period == 0)
            return fail(PB_ERR_INVALID_ARG, "pb_set_state: update_period must be >= 1");
        EnvR s;
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
        pack_env(s, c0[i], c1[i], ga[i], gb[i]);
        uint32_t f[PEL_WORDS];
        rows_to_flat(o.pellets, f);
        for (int w = 0; w < PEL_WORDS; w++) pel[(size_t)w * c + i] = f[w];
    }
    DeviceGuard g(e->device);
    if (whole_device) {
        PB_CUDA(cudaDeviceSynchronize());
        stream = nullptr;
    } else {
        int rc = sync_engine_streams(e);
        if (rc != PB_OK) return rc;
    }
    PB_CUDA(cudaMemcpyAsync(e->st.core0 + first, c0.data(), sizeof(uint4) * c, cudaMemcpyHostToDevice, stream));
    PB_CUDA(cudaMemcpyAsync(e->st.core1 + first, c1.data(), sizeof(uint4) * c, cudaMemcpyHostToDevice, stream));
    PB_CUDA(cudaMemcpyAsync(e->st.ghosts + first, ga.data(), sizeof(uint4) * c, cudaMemcpyHostToDevice, stream));
    PB_CUDA(cudaMemcpyAsync(e->st.ghosts + e->npad + first, gb.data(), sizeof(uint4) * c,
                            cudaMemcpyHostToDevice, stream));
    PB_CUDA(cudaMemcpy2DAsync(e->st.pel + first, sizeof(uint32_t) * e->npad, pel.data(),
                              sizeof(uint32_t) * c, sizeof(uint32_t) * c, PEL_WORDS,
                              cudaMemcpyHostToDevice, stream));
    // the host vectors above die with this call: the copies must have left them
    PB_CUDA(cudaStreamSynchronize(stream));
    return PB_OK;
}

int pb_set_state(pb_engine *e, int64_t first, int64_t count, const pb_env_state *in) {
    return set_state_impl(e, first, count, in, true, nullptr);
}
int pb_set_state_on(pb_engine *e, int64_t first, int64_t count, const pb_env_state *in, void *stream) {
    return set_state_impl(e, first, count, in, false, (cudaStream_t)stream);
}

int pb_clone_env(pb_engine *dst, int64_t dst_i, const pb_engine *src, int64_t src_i,
                 void *stream) {
    if (!dst || !src || dst_i < 0 || dst_i >= dst->n || src_i < 0 || src_i >= src->n ||
        dst->device != src->device)
        return fail(PB_ERR_INVALID_ARG, "pb_clone_env: bad arguments");
    DeviceGuard g(dst->device);
    k_clone<<<1, 32, 0, (cudaStream_t)stream>>>(dst->st, dst_i, src->st, src_i);
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int pb_step_host_begin(pb_engine *e, const uint8_t *actions_host, int32_t *reward_host,
                       uint8_t *done_host, uint8_t *mask_host, int auto_reset, void *stream) {
    if (!e || !actions_host || !reward_host || !done_host)
        return fail(PB_ERR_INVALID_ARG, "pb_step_host: null argument");
    if (e->host_step_pending)
        return fail(PB_ERR_INVALID_ARG, "pb_step_host_begin: the previous step has not been ended");
    DeviceGuard g(e->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (!e->host_bad) PB_CUDA(cudaMallocHost(&e->host_bad, sizeof(unsigned long long)));
    PB_CUDA(cudaMemcpyAsync(e->stg_actions, actions_host, (size_t)e->n, cudaMemcpyHostToDevice, s));
    int rc = pb_step(e, e->stg_actions, e->stg_reward, e->stg_done, mask_host ? e->stg_mask : nullptr,
                     auto_reset, stream);
    if (rc != PB_OK) return rc;
    // reward / done / mask only depend on k_step: ship them to the host on the side stream while
    // whatever the caller enqueues next on its stream (the much longer observation encode) runs
    PB_CUDA(cudaEventRecord(e->ev_step, s));
    cudaStream_t c = e->copy_stream;
    PB_CUDA(cudaStreamWaitEvent(c, e->ev_step, 0));
    *e->host_bad = ~0ull;
    PB_CUDA(cudaMemcpyAsync(reward_host, e->stg_reward, sizeof(int32_t) * (size_t)e->n,
                            cudaMemcpyDeviceToHost, c));
    PB_CUDA(cudaMemcpyAsync(done_host, e->stg_done, (size_t)e->n, cudaMemcpyDeviceToHost, c));
    if (mask_host)
        PB_CUDA(cudaMemcpyAsync(mask_host, e->stg_mask, 5 * (size_t)e->n, cudaMemcpyDeviceToHost, c));
    PB_CUDA(cudaMemcpyAsync(e->host_bad, e->bad_action, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c));
    // the next step's kernel overwrites the staging buffers: it has to wait for these copies
    PB_CUDA(cudaEventRecord(e->ev_host_copied, c));
    e->host_step_pending = true;
    return PB_OK;
}

int pb_step_host_end(pb_engine *e, void *stream) {
    if (!e) return fail(PB_ERR_INVALID_ARG, "pb_step_host_end: null engine");
    if (!e->host_step_pending) return PB_OK;
    DeviceGuard g(e->device);
    e->host_step_pending = false;
    PB_CUDA(cudaStreamSynchronize(e->copy_stream));
    PB_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, e->ev_host_copied, 0));
    if (*e->host_bad != ~0ull) {
        PB_CUDA(cudaMemsetAsync(e->bad_action, 0xff, sizeof(unsigned long long), (cudaStream_t)stream));
        return fail(PB_ERR_INVALID_ACTION, "Invalid action");
    }
    return PB_OK;
}

int pb_step_host(pb_engine *e, const uint8_t *actions_host, int32_t *reward_host,
                 uint8_t *done_host, uint8_t *mask_host, int auto_reset, float *obs_out_dev,
                 int64_t env_stride_floats, void *stream) {
    int rc = pb_step_host_begin(e, actions_host, reward_host, done_host, mask_host, auto_reset, stream);
    if (rc != PB_OK) return rc;
    if (obs_out_dev) {
        rc = pb_encode_obs(e, obs_out_dev, env_stride_floats, stream);
        if (rc != PB_OK) {
            pb_step_host_end(e, stream);
            return rc;
        }
    }
    return pb_step_host_end(e, stream);
}

// The small-batch host round trip of the per-env drop-in (gym.py): what the reference's loop does
// per env and step -- step(), is_done(), obs_numpy(), action_mask(), score() ...
// (src/replay_buffer.py:107-137, src/algorithms/train_dqn.py:95-113) -- as ONE kernel launch and
// no copy operation: k_gym_call (pb_encode_small.cuh) reads the actions in place from pinned host
// memory, steps, answers the getters, encodes the observation and writes all of it straight into
// a pinned block; its last CTA publishes the call's sequence number there and the host waits for
// that word (a stream query every few thousand polls catches a failed launch).
namespace {
// the pinned block, the pinned actions and the arrival counter of pb_step_snapshot_host
int snap_alloc(pb_engine *e) {
    const size_t n = (size_t)e->n;
    const GymCallLayout l = gym_call_layout(n);
    if (!e->snap_pin) {
        PB_CUDA(cudaMallocHost(&e->snap_pin, l.total));
        memset(e->snap_pin, 0, l.off_obs);
    }
    if (!e->snap_actions_pin) PB_CUDA(cudaMallocHost(&e->snap_actions_pin, n));
    if (!e->snap_arrive) {
        PB_CUDA(cudaMalloc(&e->snap_arrive, sizeof(unsigned int)));
        PB_CUDA(cudaMemset(e->snap_arrive, 0, sizeof(unsigned int)));
    }
    static thread_local int attr_set_for = -1;
    if (attr_set_for != e->device) {
        PB_CUDA(cudaFuncSetAttribute(k_gym_call, cudaFuncAttributeMaxDynamicSharedMemorySize, SMALL_SMEM));
        attr_set_for = e->device;
    }
    return PB_OK;
}
}  // namespace

int pb_snapshot_block(pb_engine *e, void **block_host, uint8_t **actions_host, size_t *offsets) {
    if (!e || !block_host || !actions_host || !offsets)
        return fail(PB_ERR_INVALID_ARG, "pb_snapshot_block: null argument");
    if (e->n > PB_SNAPSHOT_MAX_ENVS)
        return fail(PB_ERR_INVALID_ARG, "pb_snapshot_block: engines of up to 1024 envs");
    DeviceGuard g(e->device);
    const int rc = snap_alloc(e);
    if (rc != PB_OK) return rc;
    const GymCallLayout l = gym_call_layout((size_t)e->n);
    *block_host = e->snap_pin;
    *actions_host = e->snap_actions_pin;
    offsets[0] = l.off_info;
    offsets[1] = l.off_reward;
    offsets[2] = l.off_done;
    offsets[3] = l.off_mask;
    offsets[4] = l.off_obs;
    offsets[5] = l.total;
    return PB_OK;
}

int pb_step_snapshot_host(pb_engine *e, const uint8_t *actions_host, int auto_reset,
                          int32_t *reward_host, uint8_t *done_host, int32_t *info_host,
                          uint8_t *mask_host, float *obs_host, void *stream) {
    if (!e) return fail(PB_ERR_INVALID_ARG, "pb_step_snapshot_host: null engine");
    if (e->n > PB_SNAPSHOT_MAX_ENVS)
        return fail(PB_ERR_INVALID_ARG, "pb_step_snapshot_host: engines of up to 1024 envs");
    if (actions_host && (!reward_host || !done_host))
        return fail(PB_ERR_INVALID_ARG, "pb_step_snapshot_host: a step needs reward / done outputs");
    DeviceGuard g(e->device);
    cudaStream_t s = (cudaStream_t)stream;
    // the call waits for its kernel on the host: inside a stream capture the kernel would only be
    // recorded and the wait below would never end
    cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &capturing) != cudaSuccess || capturing != cudaStreamCaptureStatusNone) {
        cudaGetLastError();
        return fail(PB_ERR_INVALID_ARG, "pb_step_snapshot_host: the stream is being captured (a host round trip cannot be part of a CUDA graph)");
    }
    const size_t n = (size_t)e->n;
    const GymCallLayout l = gym_call_layout(n);
    const int rc = snap_alloc(e);
    if (rc != PB_OK) return rc;
    unsigned char *h = e->snap_pin;
    // the previous call waited for its kernel: the pinned actions and the block are free.
    // Pointers INTO the engine's own pinned buffers (pb_snapshot_block) mean "in place": no copy.
    if (actions_host && actions_host != e->snap_actions_pin) memcpy(e->snap_actions_pin, actions_host, n);
    const unsigned long long seq = ++e->snap_seq;
    k_gym_call<<<(unsigned)n, GYM_CALL_THREADS, SMALL_SMEM, s>>>(
        e->st, rng_params(e), actions_host ? e->snap_actions_pin : nullptr, auto_reset,
        e->bad_action, h, obs_host ? 1 : 0, e->snap_arrive, seq,
        trace_slot(e, actions_host ? PB_TRACE_STEP : PB_TRACE_ENCODE));
    PB_CUDA(cudaGetLastError());
    volatile unsigned long long *seq_word = reinterpret_cast<volatile unsigned long long *>(h + l.off_seq);
    for (unsigned int polls = 1; *seq_word != seq; polls++) {
        // a failed launch never publishes: ask the stream now and then
        if ((polls & 0xfff) == 0) {
            const cudaError_t q = cudaStreamQuery(s);
            if (q != cudaSuccess && q != cudaErrorNotReady) PB_CUDA(q);
        }
        cpu_relax();
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    auto out = [&](void *dst, size_t off, size_t bytes) {
        if (dst && dst != h + off) memcpy(dst, h + off, bytes);
    };
    if (actions_host) {
        out(reward_host, l.off_reward, 4 * n);
        out(done_host, l.off_done, n);
    }
    out(info_host, l.off_info, 32 * n);
    out(mask_host, l.off_mask, 5 * n);
    out(obs_host, l.off_obs, sizeof(float) * OBS_FLOATS * n);
    unsigned long long bad;
    memcpy(&bad, h + l.off_bad, sizeof(bad));
    if (bad != ~0ull) return fail(PB_ERR_INVALID_ACTION, "Invalid action");
    return PB_OK;
}

// Observations straight to host memory, PCIe-bound instead of pageable-bound: the range is cut
// into chunks of OBS_HOST_CHUNK envs; chunk k is encoded into device staging buffer k & 1 on the
// caller's stream and crosses PCIe on the engine's copy stream while chunk k + 1 is being
// encoded.  A pinned (or registered) destination receives the copies directly; a pageable one
// goes through two pinned bounce buffers, the host memcpy of chunk k overlapping the transfer of
// chunk k + 1 (a cudaMemcpy into pageable memory is staged by the driver, synchronously, at a
// fraction of the link rate).
constexpr long long OBS_HOST_CHUNK = 1024;  // 60 MB per staging buffer

int pb_obs_host(pb_engine *e, int64_t first, int64_t count, float *obs_host, void *stream) {
    if (!e || !obs_host || first < 0 || count < 0 || first + count > e->n)
        return fail(PB_ERR_INVALID_ARG, "pb_obs_host: bad arguments");
    if (count == 0) return PB_OK;
    DeviceGuard g(e->device);
    cudaStream_t s = (cudaStream_t)stream;
    cudaStream_t c = e->copy_stream;
    const long long chunk = count < OBS_HOST_CHUNK ? count : OBS_HOST_CHUNK;
    const size_t env_bytes = sizeof(float) * OBS_FLOATS;
    if (!e->ev_obs_enc[0])
        for (int b = 0; b < 2; b++) {
            PB_CUDA(cudaEventCreateWithFlags(&e->ev_obs_enc[b], cudaEventDisableTiming));
            PB_CUDA(cudaEventCreateWithFlags(&e->ev_obs_copy[b], cudaEventDisableTiming));
        }
    const long long nchunks = (count + chunk - 1) / chunk;
    if (e->stg_obs_envs < chunk) {  // grow-only; the second buffer only when there are >= 2 chunks
        PB_CUDA(cudaStreamSynchronize(s));
        PB_CUDA(cudaStreamSynchronize(c));
        for (int b = 0; b < 2; b++) {
            cudaFree(e->stg_obs[b]);
            e->stg_obs[b] = nullptr;
        }
        e->stg_obs_envs = 0;
        for (int b = 0; b < (chunk == OBS_HOST_CHUNK ? 2 : 1); b++)
            PB_CUDA(cudaMalloc(&e->stg_obs[b], env_bytes * (size_t)chunk));
        e->stg_obs_envs = chunk;
    }
    if (nchunks > 1 && !e->stg_obs[1]) PB_CUDA(cudaMalloc(&e->stg_obs[1], env_bytes * (size_t)e->stg_obs_envs));
    cudaPointerAttributes pa{};
    const bool pinned_dst = cudaPointerGetAttributes(&pa, obs_host) == cudaSuccess &&
                            (pa.type == cudaMemoryTypeHost || pa.type == cudaMemoryTypeManaged);
    cudaGetLastError();
    if (!pinned_dst && e->pin_obs_envs < chunk) {
        for (int b = 0; b < 2; b++) {
            if (e->pin_obs[b]) cudaFreeHost(e->pin_obs[b]);
            e->pin_obs[b] = nullptr;
        }
        e->pin_obs_envs = 0;
        for (int b = 0; b < 2; b++) PB_CUDA(cudaMallocHost(&e->pin_obs[b], env_bytes * (size_t)chunk));
        e->pin_obs_envs = chunk;
    }
    // the copy stream starts after whatever it was doing for an earlier call
    for (long long k = 0; k < nchunks + 2; k++) {
        const int b = (int)(k & 1);
        if (k >= 2) {
            const long long kk = k - 2;  // chunk whose staging buffer is reused / drained now
            const long long cnt = (kk + 1) * chunk <= count ? chunk : count - kk * chunk;
            if (pinned_dst) {
                if (k < nchunks) PB_CUDA(cudaStreamWaitEvent(s, e->ev_obs_copy[b], 0));
            } else {
                PB_CUDA(cudaEventSynchronize(e->ev_obs_copy[b]));
                memcpy(obs_host + (size_t)kk * chunk * OBS_FLOATS, e->pin_obs[b], env_bytes * (size_t)cnt);
            }
        }
        if (k >= nchunks) continue;
        const long long cnt = (k + 1) * chunk <= count ? chunk : count - k * chunk;
        int rc = launch_encode(e, e->stg_obs[b], OBS_FLOATS, first + k * chunk, cnt, s);
        if (rc != PB_OK) return rc;
        PB_CUDA(cudaEventRecord(e->ev_obs_enc[b], s));
        PB_CUDA(cudaStreamWaitEvent(c, e->ev_obs_enc[b], 0));
        float *dst = pinned_dst ? obs_host + (size_t)k * chunk * OBS_FLOATS : e->pin_obs[b];
        PB_CUDA(cudaMemcpyAsync(dst, e->stg_obs[b], env_bytes * (size_t)cnt, cudaMemcpyDeviceToHost, c));
        PB_CUDA(cudaEventRecord(e->ev_obs_copy[b], c));
    }
    if (pinned_dst) PB_CUDA(cudaStreamSynchronize(c));
    // a later call on `s` may reuse the staging buffers: order it after the last copies
    PB_CUDA(cudaStreamWaitEvent(s, e->ev_obs_copy[0], 0));
    if (nchunks > 1) PB_CUDA(cudaStreamWaitEvent(s, e->ev_obs_copy[1], 0));
    return PB_OK;
}

int pb_gather_obs(const float *ring_dev, int64_t slot_stride_floats, const int64_t *idx_dev,
                  const uint8_t *zero_mask_dev, int64_t count, float *out_dev, int device,
                  void *stream) {
    if (!ring_dev || !idx_dev || !out_dev || count < 0 || (slot_stride_floats & 3) ||
        slot_stride_floats < OBS_FLOATS || (reinterpret_cast<uintptr_t>(ring_dev) & 15) ||
        (reinterpret_cast<uintptr_t>(out_dev) & 15))
        return fail(PB_ERR_INVALID_ARG, "pb_gather_obs: bad arguments");
    if (count == 0) return PB_OK;
    DeviceGuard g(device);
    k_gather_obs<<<(unsigned)count, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4 *>(ring_dev), slot_stride_floats / 4,
        reinterpret_cast<const long long *>(idx_dev), zero_mask_dev,
        reinterpret_cast<float4 *>(out_dev));
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

}  // extern "C"

// ---- device timeline -------------------------------------------------------------------------
extern "C" {

int pb_trace_begin(pb_engine *e, int64_t capacity) {
    if (!e || capacity <= 0 || capacity > (1 << 22))
        return fail(PB_ERR_INVALID_ARG, "pb_trace_begin: bad arguments");
    DeviceGuard g(e->device);
    PB_CUDA(cudaDeviceSynchronize());
    cudaFree(e->trace);
    e->trace = nullptr;
    e->trace_cap = e->trace_n = 0;
    PB_CUDA(cudaMalloc(&e->trace, sizeof(TraceRec) * (size_t)capacity));
    std::vector<TraceRec> init((size_t)capacity, TraceRec{~0ull, ~0ull, 0ull, 0ull});
    PB_CUDA(cudaMemcpy(e->trace, init.data(), sizeof(TraceRec) * (size_t)capacity, cudaMemcpyHostToDevice));
    e->trace_kinds.assign((size_t)capacity, 0u);
    e->trace_cap = capacity;
    return PB_OK;
}

int pb_trace_mark(pb_engine *e, void *stream) {
    if (!e) return fail(PB_ERR_INVALID_ARG, "pb_trace_mark: null engine");
    DeviceGuard g(e->device);
    TraceRec *t = trace_slot(e, PB_TRACE_MARK);
    if (!t) return PB_OK;
    k_trace_mark<<<1, 1, 0, (cudaStream_t)stream>>>(t);
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int pb_trace_end(pb_engine *e, pb_trace_record *out_host, int64_t max_records, int64_t *count) {
    if (!e || !count || (max_records > 0 && !out_host))
        return fail(PB_ERR_INVALID_ARG, "pb_trace_end: bad arguments");
    DeviceGuard g(e->device);
    PB_CUDA(cudaDeviceSynchronize());
    const long long n = e->trace_n < max_records ? e->trace_n : max_records;
    std::vector<TraceRec> recs((size_t)(n > 0 ? n : 0));
    if (n > 0)
        PB_CUDA(cudaMemcpy(recs.data(), e->trace, sizeof(TraceRec) * (size_t)n, cudaMemcpyDeviceToHost));
    for (long long i = 0; i < n; i++) {
        out_host[i].begin_ns = recs[(size_t)i].begin;
        out_host[i].ready_ns = recs[(size_t)i].ready;
        out_host[i].end_ns = recs[(size_t)i].end;
        out_host[i].kind = e->trace_kinds[(size_t)i];
        out_host[i].reserved = 0;
    }
    *count = e->trace_n;
    cudaFree(e->trace);
    e->trace = nullptr;
    e->trace_cap = e->trace_n = 0;
    return PB_OK;
}

}  // extern "C"

#include "pb_mcts_abi.inl"
#include "pb_rollout_abi.inl"
#include "pb_replay_abi.inl"
#include "pb_qnet_abi.inl"
#include "pb_dqn_abi.inl"
