// pb_encode_small.cuh -- the small-batch forms of the hot path: one CTA per env.
//
// k_encode_obs (pb_encode.cuh) is a persistent kernel sized for the 65 536-env shard: every one of
// its 148 CTAs zeroes 222 KB of shared memory and builds the wall template before the first ticket
// is taken, ~10 us that a 3.9 GB write stream hides and a 32-env launch does not.  The DQN loop
// (32 envs per experience step, 1 env per evaluation step), the MCTS collector (one leaf per tree)
// and the per-env drop-in (gym.py, ONE env) only ever launch small batches, so they get:
//
//   k_obs_small   PacmanGym::obs (env.rs:410-500) for <= SMALL_ENC_MAX envs: one 512-thread CTA
//                 per env assembles the whole 59 024 B observation in shared memory (base values
//                 with 128-bit stores, then the <= 29 single cells of cell_role(), the SAME
//                 per-lane roles as the persistent kernel) and copies it out with coalesced
//                 128-bit stores;
//   k_gym_call    the per-env drop-in's whole round trip in ONE launch and NO copy operation:
//                 thread 0 of the env's CTA steps it (env.rs:216-267, auto reset as k_step) and
//                 answers every getter of env.rs:269-302, the CTA encodes the observation, and
//                 all of it is written straight into the caller's block of pinned host memory;
//                 the last CTA to finish publishes a sequence number there, which is what the
//                 host waits for (pb_step_snapshot_host).
//
// Bit-exactness: the values are cell_role()'s and the fills of k_encode_obs (pv nibbles for
// channel 1, tps / period for channel 15, the wall bits for channel 0); tests diff both kernels
// against the oracle and against each other.
#pragma once

namespace pb {

// threads per CTA (= per env); -DPB_OBS_SMALL_THREADS / -DPB_GYM_CALL_THREADS for A/B builds
#ifndef PB_OBS_SMALL_THREADS
#define PB_OBS_SMALL_THREADS 512
#endif
#ifndef PB_GYM_CALL_THREADS
#define PB_GYM_CALL_THREADS 512
#endif
constexpr int OBS_SMALL_THREADS = PB_OBS_SMALL_THREADS;
constexpr int GYM_CALL_THREADS = PB_GYM_CALL_THREADS;
constexpr int SMALL_IMG_BYTES = OBS_FLOATS * 4;                      // 59 024: all 17 channels
constexpr int SMALL_SMEM = SMALL_IMG_BYTES + 64 * 4;                 // + pellet words, wall rows
static_assert(SMALL_IMG_BYTES % 16 == 0, "the image is copied out as float4");
// launches of up to this many envs go to k_obs_small, larger ones to the persistent k_encode_obs
// (tools/small_encode_latency.py: 128 envs 8.2 vs 10.3 us per launch, 512 envs 14.4 vs 10.7 us)
constexpr long long SMALL_ENC_MAX = 256;

// One env's observation into `img` (shared memory, OBS_FLOATS floats), by all THREADS threads of
// the CTA; `words` (shared, 64 words) receives the pellet bitboard (0..27) and the wall rows
// (32..62: a per-lane row index into __constant__ memory would be serialised address by address,
// ncu: short-scoreboard stalls; shared memory serves 32 different rows at once).  Ends with a
// __syncthreads(): the image is complete and visible to the whole CTA on return.
template <int THREADS>
__device__ __forceinline__ void build_obs_image(const StateView &st, long long env, float *img,
                                                uint32_t *words, bool have_pellet_words = false) {
    const int t = threadIdx.x, lane = t & 31;
    if (t < 32) {
        if (!have_pellet_words) words[t] = (t < PEL_WORDS) ? st.pel[(long long)t * st.npad + env] : 0u;
    } else if (t < 64) {
        words[t] = c_tab.walls[t - 32];
    }
    EnvR s;
    load_env(st, env, s);
    __syncthreads();
    // warp 0 evaluates the single-cell roles (a few hundred instructions of register work) and
    // stores them after the fill; every thread needs the two fill values only
    CellRole role{};
    if (t < 32) role = cell_role(s, words[lane], lane);
    const float pv = obs_pellet_value(s), sp = obs_speed_value(s);
    const uint32_t *walls = words + 32;
    float4 *img4 = reinterpret_cast<float4 *>(img);
    for (int j = t; j < OBS_CH * CH_VEC; j += THREADS) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < CH_VEC) {                   // ch 0: walls (env.rs:426), cell c = col * ROWS + (30 - row)
            float w[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int c = 4 * j + u;
                const int col = c / ROWS, row = ROWS - 1 - (c % ROWS);
                w[u] = ((walls[row] >> col) & 1u) ? 1.0f : 0.0f;
            }
            v = make_float4(w[0], w[1], w[2], w[3]);
        } else if (j < 2 * CH_VEC) {        // ch 1: pellet values, one nibble of the bitboard
            const int q = j - CH_VEC;
            const uint32_t nib = words[q >> 3] >> (4 * (q & 7));
            v.x = (nib & 1u) ? pv : 0.0f;
            v.y = (nib & 2u) ? pv : 0.0f;
            v.z = (nib & 4u) ? pv : 0.0f;
            v.w = (nib & 8u) ? pv : 0.0f;
        } else if (j >= 15 * CH_VEC && j < 16 * CH_VEC) {   // ch 15: ticks per step / update period (env.rs:484)
            v = make_float4(sp, sp, sp, sp);
        }
        img4[j] = v;
    }
    __syncthreads();
    if (t < 32 && role.do_store) img[CH_FLOATS + role.off] = role.val;   // disjoint cells
    __syncthreads();
}

template <int THREADS>
__device__ __forceinline__ void copy_obs_image(const float *img, float *dst) {
    const float4 *src4 = reinterpret_cast<const float4 *>(img);
    float4 *dst4 = reinterpret_cast<float4 *>(dst);
#pragma unroll 4
    for (int j = threadIdx.x; j < OBS_CH * CH_VEC; j += THREADS) dst4[j] = src4[j];
}

// The image as a channels-last observation zero-padded to 32 channels (float32 [28][31][32] per
// env = [cell][channel], what the networks' first convolution takes: models.obs_to_padded_nhwc):
// one float4 of four consecutive channels per thread and step, coalesced 128-bit stores.
constexpr int NHWC_CH = 32;
constexpr int NHWC_FLOATS = CH_FLOATS * NHWC_CH;                      // 27 776 per env
template <int THREADS>
__device__ __forceinline__ void copy_obs_image_nhwc32(const float *img, float *dst) {
    float4 *dst4 = reinterpret_cast<float4 *>(dst);
    for (int j = threadIdx.x; j < CH_FLOATS * (NHWC_CH / 4); j += THREADS) {
        const int cell = j >> 3, c = (j & 7) * 4;
        float4 v;
        v.x = c + 0 < OBS_CH ? img[(c + 0) * CH_FLOATS + cell] : 0.0f;
        v.y = c + 1 < OBS_CH ? img[(c + 1) * CH_FLOATS + cell] : 0.0f;
        v.z = c + 2 < OBS_CH ? img[(c + 2) * CH_FLOATS + cell] : 0.0f;
        v.w = c + 3 < OBS_CH ? img[(c + 3) * CH_FLOATS + cell] : 0.0f;
        dst4[j] = v;
    }
}

// Same contract as k_encode_obs for envs [env_first, env_first + gridDim.x): out + k * env_stride,
// optional ring-slot indirection read from device memory (CUDA-graph replays while the ring
// advances), optional device-timeline record.  `out_nhwc` (may be given instead of, or together
// with, `out`): the padded channels-last form of the same observations, env k at
// out_nhwc + k * NHWC_FLOATS -- the replay ring keeps the channel-major observation while the
// policy's next forward gets its input layout from the same launch.
__global__ void __launch_bounds__(OBS_SMALL_THREADS)
k_obs_small(StateView st, float *__restrict__ out, long long env_stride, long long env_first,
            const long long *__restrict__ slot_index, long long slot_stride,
            float *__restrict__ out_nhwc, TraceRec *__restrict__ trace) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *img = reinterpret_cast<float *>(smem_raw);
    uint32_t *words = reinterpret_cast<uint32_t *>(smem_raw + SMALL_IMG_BYTES);
    if (trace && threadIdx.x == 0) {
        const unsigned long long t = global_timer_ns();
        stamp_min(&trace->begin, t);
        stamp_min(&trace->ready, t);
    }
    const long long k = blockIdx.x;
    build_obs_image<OBS_SMALL_THREADS>(st, env_first + k, img, words);
    if (out) {
        if (slot_index) out += *slot_index * slot_stride;
        copy_obs_image<OBS_SMALL_THREADS>(img, out + k * env_stride);
    }
    if (out_nhwc) copy_obs_image_nhwc32<OBS_SMALL_THREADS>(img, out_nhwc + k * (long long)NHWC_FLOATS);
    if (trace && threadIdx.x == 0) stamp_max(&trace->end, global_timer_ns());
}

// Layout of the block k_gym_call fills (pb_step_snapshot_host computes the same offsets):
//   seq u64 | bad u64 | info int32[n][8] | reward int32[n] | done u8[n] | mask u8[n][5] | pad to 16
//   | obs float32[n][17][28][31]
struct GymCallLayout {
    size_t off_seq, off_bad, off_info, off_reward, off_done, off_mask, off_obs, total;
};
__host__ __device__ inline GymCallLayout gym_call_layout(size_t n) {
    GymCallLayout l;
    l.off_seq = 0;
    l.off_bad = 8;
    l.off_info = 16;
    l.off_reward = l.off_info + 32 * n;
    l.off_done = l.off_reward + 4 * n;
    l.off_mask = l.off_done + n;
    l.off_obs = (l.off_mask + 5 * n + 15) & ~(size_t)15;
    l.total = l.off_obs + sizeof(float) * (size_t)OBS_FLOATS * n;
    return l;
}

// One CTA per env of a small engine.  `actions` (pinned host memory, read in place) may be null:
// no step, only the answers.  `blk` is the caller's pinned block (GymCallLayout); `arrive` a
// device counter that is zero between launches; `seq` the number to publish when the whole block
// is in host memory.
__global__ void __launch_bounds__(GYM_CALL_THREADS, 1)
k_gym_call(StateView st, RngParams rp, const uint8_t *__restrict__ actions, int auto_reset,
           unsigned long long *bad_action, unsigned char *__restrict__ blk, int want_obs,
           unsigned int *arrive, unsigned long long seq, TraceRec *__restrict__ trace) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *img = reinterpret_cast<float *>(smem_raw);
    uint32_t *words = reinterpret_cast<uint32_t *>(smem_raw + SMALL_IMG_BYTES);
    __shared__ uint32_t s_walls[32];
    if (trace && threadIdx.x == 0) {
        const unsigned long long t = global_timer_ns();
        stamp_min(&trace->begin, t);
        stamp_min(&trace->ready, t);
    }
    // the env's pellet words go to shared memory first: the step reads them at every tick (and
    // eating writes them), and a single stepping thread hides no global-memory round trip
    const long long i = blockIdx.x;
    if (threadIdx.x < 32)
        words[threadIdx.x] = threadIdx.x < PEL_WORDS ? st.pel[(long long)threadIdx.x * st.npad + i] : 0u;
    load_walls(s_walls);   // ends with a __syncthreads()
    const GymCallLayout l = gym_call_layout((size_t)st.n);
    if (threadIdx.x == 0) {
        EnvR s;
        load_env(st, i, s);
        const PelRef pel{words, 1};
        int32_t *reward = reinterpret_cast<int32_t *>(blk + l.off_reward);
        uint8_t *done_out = blk + l.off_done;
        if (actions) {
            const int action = actions[i];
            if (action > 4) {  // env.rs:83-88: "Invalid action" -- env untouched
                atomicMin(bad_action, (unsigned long long)i);
                reward[i] = 0;
                done_out[i] = 0;
            } else {
                const RngCtx rng = make_rng(rp, i);
                int rew;
                bool done;
                gym_step(s, pel, c_tab, s_walls, rng, action, rew, done);
                if (done && auto_reset) gym_reset(s, pel, c_tab, s_walls, rng, true);
                store_env(st, i, s);
                reward[i] = rew;
                done_out[i] = done ? 1 : 0;
            }
        }
        int32_t *o = reinterpret_cast<int32_t *>(blk + l.off_info) + i * 8;   // env.rs:269-289
        o[0] = s.score;
        o[1] = s.lives;
        o[2] = is_done(s);
        o[3] = first_ai_done(pel, s);
        o[4] = s.num_pellets;
        o[5] = s.tps;
        o[6] = s.level;
        o[7] = (int)s.ticks;
        write_mask(blk + l.off_mask, i, action_mask_bits(s, s_walls));          // env.rs:293-302
    }
    __syncthreads();   // thread 0's state (global memory) and pellet words (shared) are what the CTA encodes
    if (actions && threadIdx.x < PEL_WORDS) st.pel[(long long)threadIdx.x * st.npad + i] = words[threadIdx.x];
    if (want_obs) {
        build_obs_image<GYM_CALL_THREADS>(st, i, img, words, true);
        copy_obs_image<GYM_CALL_THREADS>(img, reinterpret_cast<float *>(blk + l.off_obs) + i * (long long)OBS_FLOATS);
    }
    // The CTA's stores to host memory (all threads', made visible to thread 0 by the barrier; the
    // fence is cumulative) are performed before thread 0 counts the CTA in; the last CTA hands the
    // invalid-action record over and publishes the sequence number.  A one-env engine -- the
    // drop-in's case -- is its own last CTA and needs a single system fence.
    __syncthreads();
    if (threadIdx.x == 0) {
        if (trace) stamp_max(&trace->end, global_timer_ns());
        volatile unsigned long long *bad_word = reinterpret_cast<volatile unsigned long long *>(blk + l.off_bad);
        volatile unsigned long long *seq_word = reinterpret_cast<volatile unsigned long long *>(blk + l.off_seq);
        if (gridDim.x == 1) {
            *bad_word = atomicExch(bad_action, ~0ull);
            __threadfence_system();
            *seq_word = seq;
        } else {
            __threadfence_system();
            const unsigned int before = atomicAdd(arrive, 1u);
            if (before == gridDim.x - 1) {
                __threadfence();
                *bad_word = atomicExch(bad_action, ~0ull);
                atomicExch(arrive, 0u);
                __threadfence_system();
                *seq_word = seq;
            }
        }
    }
}

// pb_step + the observation of the resulting state for a small batch, ONE launch, device buffers:
// what the DQN loop's experience step (32 envs) and an evaluation step (1 env) need.  One CTA per
// env: thread 0 steps the env on pellet words held in shared memory (in k_step the 32 envs of a
// small batch share ONE warp and their divergent game logic is executed lane group by lane group:
// 8 us per 32-env launch; here every env has a warp of its own on its own SM), then the CTA
// encodes the observation -- channel-major at out + k * env_stride (optionally through the
// ring-slot indirection) and / or channels-last padded to 32 channels at out_nhwc + k * NHWC_FLOATS.
// Same results as k_step followed by k_obs_small (tests diff them and the oracle).
__global__ void __launch_bounds__(GYM_CALL_THREADS, 1)
k_step_obs_small(StateView st, RngParams rp, const uint8_t *__restrict__ actions,
                 int32_t *__restrict__ reward, uint8_t *__restrict__ done_out,
                 uint8_t *__restrict__ mask_out, int auto_reset, unsigned long long *bad_action,
                 float *__restrict__ out, long long env_stride, const long long *__restrict__ slot_index,
                 long long slot_stride, float *__restrict__ out_nhwc, TraceRec *__restrict__ trace) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *img = reinterpret_cast<float *>(smem_raw);
    uint32_t *words = reinterpret_cast<uint32_t *>(smem_raw + SMALL_IMG_BYTES);
    __shared__ uint32_t s_walls[32];
    if (trace && threadIdx.x == 0) {
        const unsigned long long t = global_timer_ns();
        stamp_min(&trace->begin, t);
        stamp_min(&trace->ready, t);
    }
    const long long i = blockIdx.x;
    if (threadIdx.x < 32)
        words[threadIdx.x] = threadIdx.x < PEL_WORDS ? st.pel[(long long)threadIdx.x * st.npad + i] : 0u;
    load_walls(s_walls);   // ends with a __syncthreads()
    if (threadIdx.x == 0) {
        EnvR s;
        load_env(st, i, s);
        const PelRef pel{words, 1};
        const int action = actions[i];
        if (action > 4) {  // env.rs:83-88: "Invalid action" -- env untouched
            atomicMin(bad_action, (unsigned long long)i);
            reward[i] = 0;
            done_out[i] = 0;
        } else {
            const RngCtx rng = make_rng(rp, i);
            int rew;
            bool done;
            gym_step(s, pel, c_tab, s_walls, rng, action, rew, done);
            if (done && auto_reset) gym_reset(s, pel, c_tab, s_walls, rng, true);
            store_env(st, i, s);
            reward[i] = rew;
            done_out[i] = done ? 1 : 0;
        }
        if (mask_out) write_mask(mask_out, i, action_mask_bits(s, s_walls));
    }
    __syncthreads();   // thread 0's state (global memory) and pellet words (shared) are what the CTA encodes
    if (threadIdx.x < PEL_WORDS) st.pel[(long long)threadIdx.x * st.npad + i] = words[threadIdx.x];
    if (out || out_nhwc) {
        build_obs_image<GYM_CALL_THREADS>(st, i, img, words, true);
        if (out) {
            if (slot_index) out += *slot_index * slot_stride;
            copy_obs_image<GYM_CALL_THREADS>(img, out + i * env_stride);
        }
        if (out_nhwc) copy_obs_image_nhwc32<GYM_CALL_THREADS>(img, out_nhwc + i * (long long)NHWC_FLOATS);
    }
    if (trace && threadIdx.x == 0) stamp_max(&trace->end, global_timer_ns());
}

}  // namespace pb
