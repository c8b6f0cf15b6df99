// pb_replay_abi.inl -- pb_replay_*: the replay ring of the DQN loop on the device
// (reference: /root/reference/src/replay_buffer.py:57-143 and the collate of
// src/algorithms/train_dqn.py:152-165); included at the end of pb_kernels.cu.
//
// The ring is a TIME ring: slot s = t % R holds, for all n envs, the observation before action t
// (written by k_encode_obs straight into the ring) and (action, reward, done, next mask, valid)
// of transition t; next_obs of a transition is slot (s + 1) % R.  Every kernel below reads the
// step counter t from DEVICE memory at kernel time, so one captured CUDA graph keeps working
// while the ring advances.
//
//   k_replay_record   deque.append of one step for all envs (replay_buffer.py:130)
//   k_replay_sample   random.sample(buffer, B): B distinct valid transitions, uniform
//                     (replay_buffer.py:141-143): ordered compaction of the valid flags + a
//                     partial Fisher-Yates shuffle in shared memory, counter RNG
//   k_replay_collate  torch.stack of obs / next_obs (zeros where done) and the scalar columns
//                     (train_dqn.py:153-165), obs optionally emitted CHANNELS-LAST so that the
//                     Q-network's NHWC convolutions read it without a layout-conversion kernel
//   k_replay_slot     last_obs of the policy forward: ring slot t % R, NCHW or NHWC

namespace pb {

__global__ void __launch_bounds__(128)
k_replay_record(long long n, long long ring_slots, const long long *__restrict__ t_dev,
                const uint8_t *__restrict__ actions, const int32_t *__restrict__ reward,
                const uint8_t *__restrict__ done, const uint8_t *__restrict__ mask,
                const uint8_t *__restrict__ in_phase1, uint8_t *__restrict__ ring_actions,
                int32_t *__restrict__ ring_rewards, uint8_t *__restrict__ ring_done,
                uint8_t *__restrict__ ring_next_mask, uint8_t *__restrict__ ring_valid) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long t = *t_dev;
    const long long s = (t % ring_slots) * n + i, nx = ((t + 1) % ring_slots) * n + i;
    const uint8_t d = done[i];
    ring_actions[s] = actions[i];
    ring_rewards[s] = reward[i];
    ring_done[s] = d;
    // next_action_mask is [False] * 5 when the episode ended (replay_buffer.py:115-117)
#pragma unroll
    for (int k = 0; k < 5; k++) ring_next_mask[s * 5 + k] = d ? 0 : mask[i * 5 + k];
    // steps played by the old policy of reset_env (replay_buffer.py:44-54) are not recorded
    ring_valid[s] = in_phase1 ? (in_phase1[i] ? 0 : 1) : 1;
    ring_valid[nx] = 0;  // slot t + 1 holds the current observation only
}

// the ring moves on by one step: t += 1 and the slot the next observation goes to, (t + 1) % R
// (one launch instead of three tiny torch kernels per experience step)
__global__ void k_replay_advance(long long *__restrict__ t_dev, long long *__restrict__ next_slot_dev,
                                 long long ring_slots) {
    const long long t = *t_dev + 1;
    *t_dev = t;
    *next_slot_dev = (t + 1) % ring_slots;
}

constexpr int SAMPLE_THREADS = 1024;
constexpr int SAMPLE_SMEM_MAX = 227 * 1024 - SAMPLE_THREADS * 4;  // dynamic part
// dynamic shared memory: int32 list[total] + uint32 raw[batch]
__global__ void __launch_bounds__(SAMPLE_THREADS)
k_replay_sample(const uint8_t *__restrict__ ring_valid, long long total, int batch, uint64_t seed,
                const long long *__restrict__ call_dev, long long *__restrict__ idx_out,
                int *__restrict__ num_valid_out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    int *list = reinterpret_cast<int *>(smem_raw);
    uint32_t *raw = reinterpret_cast<uint32_t *>(list + total);
    __shared__ int s_scan[SAMPLE_THREADS];
    const int tid = threadIdx.x;
    const long long per = (total + SAMPLE_THREADS - 1) / SAMPLE_THREADS;
    const long long lo = per * tid, hi = (lo + per < total) ? lo + per : total;
    int cnt = 0;
    for (long long j = lo; j < hi; j++) cnt += ring_valid[j] ? 1 : 0;
    // block-wide inclusive scan of the per-thread counts (ordered compaction: the list is
    // ascending in the flat ring index, whatever the thread schedule)
    s_scan[tid] = cnt;
    __syncthreads();
    for (int off = 1; off < SAMPLE_THREADS; off <<= 1) {
        const int v = (tid >= off) ? s_scan[tid - off] : 0;
        __syncthreads();
        s_scan[tid] += v;
        __syncthreads();
    }
    int w = s_scan[tid] - cnt;
    for (long long j = lo; j < hi; j++)
        if (ring_valid[j]) list[w++] = (int)j;
    const int m = s_scan[SAMPLE_THREADS - 1];
    const unsigned long long call = (unsigned long long)*call_dev;
    // the draws, and (when there are enough candidates) the partner index of every position, are
    // computed by all threads; only the dependent chain of swaps is left to one thread
    for (int j = tid; j < batch; j += SAMPLE_THREADS) {
        const uint32_t u = replay_raw(seed, call, (uint32_t)j);
        raw[j] = (m >= batch) ? (uint32_t)j + below(u, (uint32_t)(m - j)) : u;
    }
    __syncthreads();
    if (tid == 0) {
        if (num_valid_out) *num_valid_out = m;
        if (m >= batch) {
            // partial Fisher-Yates: position j takes a uniformly chosen element of list[j..m)
            for (int j = 0; j < batch; j++) {
                const int r = (int)raw[j];
                const int a = list[j], b = list[r];
                list[r] = a;
                list[j] = b;
            }
        }
    }
    __syncthreads();
    // fewer valid transitions than the batch (the caller checks num_valid before it trains):
    // sample with replacement over what there is instead of reading past the list
    for (int j = tid; j < batch; j += SAMPLE_THREADS)
        idx_out[j] = (m >= batch) ? list[j] : (m > 0 ? list[below(raw[j], (uint32_t)m)] : 0);
}

// One observation (17 x 868 floats, channel-major in the ring) to `dst`: a straight float4 copy,
// or transposed to channels-last ([cell][cpad]) through a padded shared-memory tile.  An
// observation is cut into OBS_SPLIT parts of 124 cells, one CTA each: a batch of 32 policy
// observations is 224 CTAs instead of 32 (11 -> ~4 us per k_replay_slot call), a 512-sample
// collate 7 168 small CTAs instead of 1 024 big ones that ran three to an SM in 2.3 waves.
constexpr int OBS_SPLIT = 7;                          // 217 float4 per channel = 7 x 31
constexpr int PART_VEC = CH_VEC / OBS_SPLIT;          // 31 float4 per channel and part
constexpr int PART_CELLS = PART_VEC * 4;              // 124 cells
constexpr int PART_STRIDE = PART_CELLS + 1;           // 125: odd stride, conflict-free transposed reads
constexpr int COLLATE_THREADS = 128;
static_assert(PART_VEC * OBS_SPLIT == CH_VEC, "the parts tile a channel");
// `cpad`: 0 = channel-major copy (NCHW); otherwise the channel count of the channels-last output,
// 17 (dense) or a multiple of 4 up to 64 whose channels >= 17 are written as zeros: cuDNN only has
// tensor-core kernels for the first convolution when its input channels are a multiple of 8
// (17 -> 32: 230 -> 140 us forward at batch 512, tools/experiments/pad_channels_probe.py).
__device__ __forceinline__ void emit_obs_part(const float4 *__restrict__ src, float4 *__restrict__ dst,
                                              bool zero, int cpad, int part, float *tile) {
    // the part's 17 x 31 float4 are requested up front (5 independent loads per thread), then used
    constexpr int IN_VEC = OBS_CH * PART_VEC, IN_ITERS = (IN_VEC + COLLATE_THREADS - 1) / COLLATE_THREADS;
    const int out_vec = cpad ? PART_CELLS * cpad / 4 : 0;
    float4 *o4 = dst + (long long)part * out_vec;
    if (zero) {
        if (!cpad) {
            for (int q = threadIdx.x; q < IN_VEC; q += COLLATE_THREADS) {
                const int ch = q / PART_VEC, v = q - ch * PART_VEC;
                dst[ch * CH_VEC + part * PART_VEC + v] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else {
            for (int q = threadIdx.x; q < out_vec; q += COLLATE_THREADS) o4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        return;
    }
    float4 in[IN_ITERS];
#pragma unroll
    for (int u = 0; u < IN_ITERS; u++) {
        const int q = threadIdx.x + u * COLLATE_THREADS;
        if (q < IN_VEC) {
            const int ch = q / PART_VEC, v = q - ch * PART_VEC;
            in[u] = __ldg(src + ch * CH_VEC + part * PART_VEC + v);
        }
    }
#pragma unroll
    for (int u = 0; u < IN_ITERS; u++) {
        const int q = threadIdx.x + u * COLLATE_THREADS;
        if (q < IN_VEC) {
            const int ch = q / PART_VEC, v = q - ch * PART_VEC;
            if (!cpad) {   // channel-major copy
                dst[ch * CH_VEC + part * PART_VEC + v] = in[u];
            } else {       // into the transposing tile
                float *t = tile + ch * PART_STRIDE + 4 * v;
                t[0] = in[u].x;
                t[1] = in[u].y;
                t[2] = in[u].z;
                t[3] = in[u].w;
            }
        }
    }
    if (!cpad) return;
    __syncthreads();
    for (int q = threadIdx.x; q < out_vec; q += COLLATE_THREADS) {
        float o[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int j = q * 4 + u, cell = j / cpad, ch = j - cell * cpad;
            o[u] = ch < OBS_CH ? tile[ch * PART_STRIDE + cell] : 0.0f;
        }
        o4[q] = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// grid: batch x {obs, next_obs} x OBS_SPLIT
__global__ void __launch_bounds__(COLLATE_THREADS)
k_replay_collate(const float4 *__restrict__ ring_obs, long long n, long long ring_slots,
                 const long long *__restrict__ idx, const uint8_t *__restrict__ ring_actions,
                 const int32_t *__restrict__ ring_rewards, const uint8_t *__restrict__ ring_done,
                 const uint8_t *__restrict__ ring_next_mask, float4 *__restrict__ obs_out,
                 float4 *__restrict__ next_obs_out, int nhwc, long long *__restrict__ actions_out,
                 float *__restrict__ rewards_out, uint8_t *__restrict__ done_out,
                 uint8_t *__restrict__ next_mask_out) {
    __shared__ float tile[OBS_CH * PART_STRIDE];
    const int part = blockIdx.x % OBS_SPLIT;
    const long long item = blockIdx.x / OBS_SPLIT;
    const long long k = item >> 1;
    const bool is_next = item & 1;
    const long long flat = idx[k];
    const long long slot = flat / n, env = flat - slot * n;
    const bool done = ring_done[flat] != 0;
    const long long out_vec = (nhwc ? CH_FLOATS * nhwc : OBS_FLOATS) / 4;   // float4 per output obs
    if (!is_next) {
        emit_obs_part(ring_obs + flat * (OBS_FLOATS / 4), obs_out + k * out_vec, false, nhwc, part, tile);
        if (part == 0) {
            if (threadIdx.x == 0) {
                actions_out[k] = ring_actions[flat];
                rewards_out[k] = (float)ring_rewards[flat];
                done_out[k] = done ? 1 : 0;
            }
            if (threadIdx.x < 5) next_mask_out[k * 5 + threadIdx.x] = ring_next_mask[flat * 5 + threadIdx.x];
        }
    } else {
        // `next_obs is None` -> zeros (train_dqn.py:154-159); otherwise the next time slot
        const long long nflat = ((slot + 1) % ring_slots) * n + env;
        emit_obs_part(ring_obs + nflat * (OBS_FLOATS / 4), next_obs_out + k * out_vec, done, nhwc, part, tile);
    }
}

// grid: n envs x OBS_SPLIT
__global__ void __launch_bounds__(COLLATE_THREADS)
k_replay_slot(const float4 *__restrict__ ring_obs, long long n, long long ring_slots,
              const long long *__restrict__ t_dev, float4 *__restrict__ out, int nhwc) {
    __shared__ float tile[OBS_CH * PART_STRIDE];
    const int part = blockIdx.x % OBS_SPLIT;
    const long long env = blockIdx.x / OBS_SPLIT;
    const long long flat = (*t_dev % ring_slots) * n + env;
    const long long out_vec = (nhwc ? CH_FLOATS * nhwc : OBS_FLOATS) / 4;
    emit_obs_part(ring_obs + flat * (OBS_FLOATS / 4), out + env * out_vec, false, nhwc, part, tile);
}

}  // namespace pb

namespace {
bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
// channels_last argument of the ABI -> channel count of the channels-last output (0 = NCHW),
// or -1 if it is not 0, 1 (= 17), 17, or a multiple of 4 in 20..64
int channels_arg(int channels_last) {
    if (channels_last == 0) return 0;
    if (channels_last == 1 || channels_last == OBS_CH) return OBS_CH;
    if (channels_last > OBS_CH && channels_last <= 64 && (channels_last & 3) == 0) return channels_last;
    return -1;
}
}  // namespace

extern "C" {

uint32_t pb_replay_raw(uint64_t seed, uint64_t call_idx, uint32_t j) { return replay_raw(seed, call_idx, j); }

int64_t pb_replay_sample_capacity(void) {
    // list (4 B per ring entry) + 4 096 raw draws + the 4 KB scan array must fit 227 KB
    return (int64_t)((SAMPLE_SMEM_MAX - 4 * 4096) / 4);
}

int pb_replay_record(int64_t n_envs, int64_t ring_slots, const int64_t *t_dev,
                     const uint8_t *actions_dev, const int32_t *reward_dev, const uint8_t *done_dev,
                     const uint8_t *mask_dev, const uint8_t *in_phase1_dev, uint8_t *ring_actions,
                     int32_t *ring_rewards, uint8_t *ring_done, uint8_t *ring_next_mask,
                     uint8_t *ring_valid, int device, void *stream) {
    if (n_envs <= 0 || ring_slots < 2 || !t_dev || !actions_dev || !reward_dev || !done_dev ||
        !mask_dev || !ring_actions || !ring_rewards || !ring_done || !ring_next_mask || !ring_valid)
        return fail(PB_ERR_INVALID_ARG, "pb_replay_record: bad arguments");
    DeviceGuard g(device);
    k_replay_record<<<(unsigned)((n_envs + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        n_envs, ring_slots, reinterpret_cast<const long long *>(t_dev), actions_dev, reward_dev,
        done_dev, mask_dev, in_phase1_dev, ring_actions, ring_rewards, ring_done, ring_next_mask,
        ring_valid);
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int pb_replay_advance(int64_t *t_dev, int64_t *next_slot_dev, int64_t ring_slots, int device, void *stream) {
    if (!t_dev || !next_slot_dev || ring_slots < 1)
        return fail(PB_ERR_INVALID_ARG, "pb_replay_advance: bad arguments");
    DeviceGuard g(device);
    k_replay_advance<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<long long *>(t_dev),
                                                       reinterpret_cast<long long *>(next_slot_dev), ring_slots);
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int pb_replay_sample(const uint8_t *ring_valid_dev, int64_t total, int64_t batch, uint64_t seed,
                     const int64_t *call_idx_dev, int64_t *idx_out_dev, int32_t *num_valid_out_dev,
                     int device, void *stream) {
    if (!ring_valid_dev || !call_idx_dev || !idx_out_dev || total <= 0 || batch <= 0 ||
        batch > 4096 || total > pb_replay_sample_capacity())
        return fail(PB_ERR_INVALID_ARG,
                    "pb_replay_sample: bad arguments (batch <= 4096, ring entries <= "
                    "pb_replay_sample_capacity())");
    DeviceGuard g(device);
    const size_t smem = sizeof(int) * (size_t)total + sizeof(uint32_t) * (size_t)batch;
    static thread_local int set_for = -1;
    if (set_for != device) {
        PB_CUDA(cudaFuncSetAttribute(k_replay_sample, cudaFuncAttributeMaxDynamicSharedMemorySize, SAMPLE_SMEM_MAX));
        set_for = device;
    }
    k_replay_sample<<<1, SAMPLE_THREADS, smem, (cudaStream_t)stream>>>(
        ring_valid_dev, total, (int)batch, seed, reinterpret_cast<const long long *>(call_idx_dev),
        reinterpret_cast<long long *>(idx_out_dev), num_valid_out_dev);
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int pb_replay_collate(const float *ring_obs_dev, int64_t n_envs, int64_t ring_slots,
                      const int64_t *idx_dev, int64_t batch, const uint8_t *ring_actions,
                      const int32_t *ring_rewards, const uint8_t *ring_done,
                      const uint8_t *ring_next_mask, float *obs_out_dev, float *next_obs_out_dev,
                      int channels_last, int64_t *actions_out_dev, float *rewards_out_dev,
                      uint8_t *done_out_dev, uint8_t *next_mask_out_dev, int device, void *stream) {
    if (!ring_obs_dev || !idx_dev || !ring_actions || !ring_rewards || !ring_done || !ring_next_mask ||
        !obs_out_dev || !next_obs_out_dev || !actions_out_dev || !rewards_out_dev || !done_out_dev ||
        !next_mask_out_dev || n_envs <= 0 || ring_slots < 2 || batch < 0 || !aligned16(ring_obs_dev) ||
        !aligned16(obs_out_dev) || !aligned16(next_obs_out_dev) || channels_arg(channels_last) < 0)
        return fail(PB_ERR_INVALID_ARG, "pb_replay_collate: bad arguments");
    channels_last = channels_arg(channels_last);
    if (batch == 0) return PB_OK;
    DeviceGuard g(device);
    k_replay_collate<<<(unsigned)(2 * batch * OBS_SPLIT), COLLATE_THREADS, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4 *>(ring_obs_dev), n_envs, ring_slots,
        reinterpret_cast<const long long *>(idx_dev), ring_actions, ring_rewards, ring_done,
        ring_next_mask, reinterpret_cast<float4 *>(obs_out_dev),
        reinterpret_cast<float4 *>(next_obs_out_dev), channels_last,
        reinterpret_cast<long long *>(actions_out_dev), rewards_out_dev, done_out_dev,
        next_mask_out_dev);
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int pb_replay_slot_obs(const float *ring_obs_dev, int64_t n_envs, int64_t ring_slots,
                       const int64_t *t_dev, float *out_dev, int channels_last, int device,
                       void *stream) {
    if (!ring_obs_dev || !t_dev || !out_dev || n_envs <= 0 || ring_slots < 1 ||
        !aligned16(ring_obs_dev) || !aligned16(out_dev) || channels_arg(channels_last) < 0)
        return fail(PB_ERR_INVALID_ARG, "pb_replay_slot_obs: bad arguments");
    channels_last = channels_arg(channels_last);
    DeviceGuard g(device);
    k_replay_slot<<<(unsigned)(n_envs * OBS_SPLIT), COLLATE_THREADS, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4 *>(ring_obs_dev), n_envs, ring_slots,
        reinterpret_cast<const long long *>(t_dev), reinterpret_cast<float4 *>(out_dev), channels_last);
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

}  // extern "C"
