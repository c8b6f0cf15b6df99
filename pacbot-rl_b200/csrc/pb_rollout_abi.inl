// pb_rollout_abi.inl -- pb_rollout_*: pipelined rollout steps with double-buffered env state on two
// engine-owned streams (include/pacbot_b200.h); included at the end of pb_kernels.cu.
//
//     encode stream 0 : enc(t)                                enc(t+2)         ...
//     encode stream 1 :                     enc(t+1)                           ...
//     step stream     :     step(t+1)            step(t+2)         ...
//
// step(t+1) reads state buffer t and writes buffer t+1 (pb_step_flip) while enc(t) is still
// streaming buffer t to HBM; it waits only for enc(t-1), whose buffer it overwrites.  The whole
// orchestration of a step is one C call (a handful of CUDA API calls), so a busy host does not
// put bubbles into the stream of 0.52 ms encoder launches.
//
// Consecutive encoders go to two streams alternately and depend only on their own step kernel:
// the encoder is one persistent CTA per SM (all of an SM's shared memory), so the CTAs of
// enc(t+1) start SM by SM as the CTAs of enc(t) run out of tickets and exit.  On one stream the
// device timeline (tools/timeline.py) showed 6 us between the last warp of one encoder and the
// first working CTA of the next (launch + prologue: 222 KB of shared memory to zero, the wall
// template, the L2 pinning of the state), i.e. 1.1 % of every step.  Encoders that write
// overlapping output ranges (the same obs buffer twice in a row) stay serialised.
//
// pb_rollout_chunk_*: the same overlap for an env population larger than one encoder launch
// should cover (the 2 M-env shard of BASELINE configs[2]: 352 MB of state, a second buffer would
// only add HBM traffic).  The caller walks the envs chunk by chunk; a chunk's k_step runs IN PLACE
// on the step stream while the encoders of other chunks stream to HBM, and waits only for the
// encoder that read the same envs one sweep earlier:
//
//     encode stream 0 : enc(c0)            enc(c2)              ...
//     encode stream 1 :           enc(c1)            enc(c3)    ...
//     step stream     : step(c0) step(c1) step(c2) ...  (each as soon as its envs' last encoder is done)

namespace {

int rollout_ensure(pb_engine *e) {
    if (e->roll_step) return PB_OK;
    PB_CUDA(cudaStreamCreateWithFlags(&e->roll_step, cudaStreamNonBlocking));
    for (cudaStream_t &s : e->roll_enc) PB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    // PB_ROLL_ENC_STREAMS=1: every encoder on one stream (A/B runs)
    const char *v = getenv("PB_ROLL_ENC_STREAMS");
    e->roll_enc_streams = (v && v[0] == '1') ? 1 : 2;
    for (cudaEvent_t *ev : {&e->ev_roll_step, &e->ev_roll_fork, &e->ev_roll_join, &e->ev_enc_done[0],
                            &e->ev_enc_done[1], &e->ev_roll_copied})
        PB_CUDA(cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
    return PB_OK;
}

// the engine's streams start after everything already queued on the caller's stream
int rollout_fork(pb_engine *e, cudaStream_t user) {
    if (e->roll_started) return PB_OK;
    PB_CUDA(cudaEventRecord(e->ev_roll_fork, user));
    PB_CUDA(cudaStreamWaitEvent(e->roll_step, e->ev_roll_fork, 0));
    for (cudaStream_t s : e->roll_enc) PB_CUDA(cudaStreamWaitEvent(s, e->ev_roll_fork, 0));
    e->roll_last_obs = nullptr;
    PB_CUDA(cudaStreamWaitEvent(e->copy_stream, e->ev_roll_fork, 0));
    e->roll_started = true;
    return PB_OK;
}

// step (flip) on the step stream, encode on the encode stream; reward/done/mask/actions device ptrs
int rollout_step_device(pb_engine *e, const uint8_t *actions_dev, uint64_t call_idx,
                        uint8_t *actions_out_dev, int32_t *reward_dev, uint8_t *done_dev,
                        uint8_t *mask_dev, int auto_reset, float *obs_out_dev, long long stride) {
    const int slot = (int)(e->roll_t & 1ull);
    // this step overwrites the buffer that the encode of two steps ago read
    PB_CUDA(cudaStreamWaitEvent(e->roll_step, e->ev_enc_done[slot], 0));
    // chunk encoders of pb_rollout_chunk_* read whichever buffer was current when they were
    // launched: a caller that mixes the two forms gets them serialised here
    for (auto &r : e->roll_chunks) {
        PB_CUDA(cudaStreamWaitEvent(e->roll_step, r.enc_done, 0));
        PB_CUDA(cudaEventDestroy(r.enc_done));
    }
    e->roll_chunks.clear();
    int rc = pb_step_flip(e, actions_dev, call_idx, actions_out_dev, reward_dev, done_dev, mask_dev,
                          auto_reset, e->roll_step);
    if (rc != PB_OK) return rc;
    PB_CUDA(cudaEventRecord(e->ev_roll_step, e->roll_step));
    cudaStream_t enc = e->roll_enc[e->roll_enc_streams == 2 ? slot : 0];
    PB_CUDA(cudaStreamWaitEvent(enc, e->ev_roll_step, 0));
    if (obs_out_dev) {
        // an encoder that writes where the previous one wrote must not overtake it
        const size_t bytes = sizeof(float) * (size_t)stride * (size_t)e->n;
        const char *lo = reinterpret_cast<const char *>(obs_out_dev);
        const char *plo = reinterpret_cast<const char *>(e->roll_last_obs);
        if (e->roll_last_obs && lo < plo + e->roll_last_obs_bytes && plo < lo + bytes)
            PB_CUDA(cudaStreamWaitEvent(enc, e->ev_enc_done[slot ^ 1], 0));
        rc = launch_encode(e, obs_out_dev, stride, 0, e->n, enc);
        if (rc != PB_OK) return rc;
        e->roll_last_obs = obs_out_dev;
        e->roll_last_obs_bytes = bytes;
        e->roll_last_enc = e->roll_enc_streams == 2 ? slot : 0;
    }
    PB_CUDA(cudaEventRecord(e->ev_enc_done[slot], enc));
    e->roll_t += 1;
    return PB_OK;
}

// One chunk of the in-place chunked rollout: k_step of envs [first, first + count) on the step
// stream, their observations on one of the two encode streams (alternating launch by launch).
// Every encoder launch leaves a record {env range, output range, stream, event}: a later in-place
// step of overlapping envs waits for it (and retires it), a later encoder on the OTHER stream
// whose output overlaps waits for it (ring slots are reused within a sweep).
int rollout_chunk_device(pb_engine *e, long long first, long long count, const uint8_t *actions_dev,
                         uint64_t call_idx, uint8_t *actions_out_dev, int32_t *reward_dev,
                         uint8_t *done_dev, uint8_t *mask_dev, int auto_reset, float *obs_out_dev,
                         long long stride) {
    // the in-place step overwrites state that earlier encoders of these envs may still be reading
    for (cudaEvent_t ev : e->ev_enc_done) PB_CUDA(cudaStreamWaitEvent(e->roll_step, ev, 0));
    cudaEvent_t mine = nullptr;
    auto retire = [&](size_t k) -> int {
        PB_CUDA(cudaStreamWaitEvent(e->roll_step, e->roll_chunks[k].enc_done, 0));
        if (!mine)
            mine = e->roll_chunks[k].enc_done;
        else
            PB_CUDA(cudaEventDestroy(e->roll_chunks[k].enc_done));
        e->roll_chunks.erase(e->roll_chunks.begin() + (long)k);
        return PB_OK;
    };
    for (size_t k = 0; k < e->roll_chunks.size();) {
        const pb_engine::RollChunk &r = e->roll_chunks[k];
        if (r.first < first + count && first < r.first + r.count) {
            int rc = retire(k);
            if (rc != PB_OK) return rc;
        } else {
            k++;
        }
    }
    if (e->roll_chunks.size() >= 1024) {  // a caller sliding its ranges: fold the oldest record
        int rc = retire(0);
        if (rc != PB_OK) return rc;
    }
    // 512-thread CTAs: the kernel shares the SMs with running encoders (see k_step)
    const unsigned wide_grid = (unsigned)((count + STEP_THREADS_WIDE - 1) / STEP_THREADS_WIDE);
    k_step<<<wide_grid, STEP_THREADS_WIDE, 0, e->roll_step>>>(
        e->st, e->st, rng_params(e), actions_dev, reward_dev, done_dev, mask_dev, auto_reset,
        e->bad_action, actions_out_dev, call_idx, first, count, trace_slot(e, PB_TRACE_STEP));
    PB_CUDA(cudaGetLastError());
    PB_CUDA(cudaEventRecord(e->ev_roll_step, e->roll_step));
    if (!obs_out_dev) {
        if (mine) PB_CUDA(cudaEventDestroy(mine));
        return PB_OK;
    }
    const int which = e->roll_enc_streams == 2 ? (int)(e->roll_chunk_ctr & 1ull) : 0;
    e->roll_chunk_ctr += 1;
    cudaStream_t enc = e->roll_enc[which];
    PB_CUDA(cudaStreamWaitEvent(enc, e->ev_roll_step, 0));
    // an encoder that writes where an earlier one wrote must not overtake it
    const size_t bytes = sizeof(float) * (size_t)stride * (size_t)count;
    const char *lo = reinterpret_cast<const char *>(obs_out_dev);
    for (const pb_engine::RollChunk &r : e->roll_chunks)
        if (r.stream != which && lo < r.out_lo + r.out_bytes && r.out_lo < lo + bytes)
            PB_CUDA(cudaStreamWaitEvent(enc, r.enc_done, 0));
    const char *plo = reinterpret_cast<const char *>(e->roll_last_obs);   // of a whole-batch encoder
    if (e->roll_last_obs && lo < plo + e->roll_last_obs_bytes && plo < lo + bytes)
        for (cudaEvent_t ev : e->ev_enc_done) PB_CUDA(cudaStreamWaitEvent(enc, ev, 0));
    int rc = launch_encode(e, obs_out_dev, stride, first, count, enc);
    if (rc != PB_OK) {
        if (mine) cudaEventDestroy(mine);
        return rc;
    }
    e->roll_last_enc = which;
    if (!mine) PB_CUDA(cudaEventCreateWithFlags(&mine, cudaEventDisableTiming));
    PB_CUDA(cudaEventRecord(mine, enc));
    e->roll_chunks.push_back(pb_engine::RollChunk{first, count, lo, bytes, which, mine});
    return PB_OK;
}

int rollout_check_obs(const float *obs_out_dev, int64_t &stride) {
    if (stride == 0) stride = OBS_FLOATS;
    if (obs_out_dev && (stride < OBS_FLOATS || (stride & 3) ||
                        (reinterpret_cast<uintptr_t>(obs_out_dev) & 15)))
        return fail(PB_ERR_INVALID_ARG,
                    "pb_rollout: stride must be >= 14756 and a multiple of 4 floats, obs_out must be "
                    "16-byte aligned");
    return PB_OK;
}

}  // namespace

extern "C" {

int pb_rollout_random_step(pb_engine *e, uint64_t call_idx, uint8_t *actions_out_dev,
                           int32_t *reward_dev, uint8_t *done_dev, uint8_t *mask_out_dev,
                           int auto_reset, float *obs_out_dev, int64_t env_stride_floats,
                           void *stream) {
    if (!e || !reward_dev || !done_dev)
        return fail(PB_ERR_INVALID_ARG, "pb_rollout_random_step: null argument");
    int rc = rollout_check_obs(obs_out_dev, env_stride_floats);
    if (rc != PB_OK) return rc;
    DeviceGuard g(e->device);
    if ((rc = rollout_ensure(e)) != PB_OK) return rc;
    if ((rc = rollout_fork(e, (cudaStream_t)stream)) != PB_OK) return rc;
    return rollout_step_device(e, nullptr, call_idx, actions_out_dev, reward_dev, done_dev,
                               mask_out_dev, auto_reset, obs_out_dev, env_stride_floats);
}

int pb_rollout_host_step(pb_engine *e, const uint8_t *actions_host, int32_t *reward_host,
                         uint8_t *done_host, uint8_t *mask_host, int auto_reset, float *obs_out_dev,
                         int64_t env_stride_floats, void *stream) {
    if (!e || !actions_host || !reward_host || !done_host)
        return fail(PB_ERR_INVALID_ARG, "pb_rollout_host_step: null argument");
    int rc = rollout_check_obs(obs_out_dev, env_stride_floats);
    if (rc != PB_OK) return rc;
    DeviceGuard g(e->device);
    if ((rc = rollout_ensure(e)) != PB_OK) return rc;
    if ((rc = rollout_fork(e, (cudaStream_t)stream)) != PB_OK) return rc;
    // actions host -> device on the step stream (the previous step kernel has consumed the
    // staging buffer: same stream)
    PB_CUDA(cudaMemcpyAsync(e->stg_actions, actions_host, (size_t)e->n, cudaMemcpyHostToDevice,
                            e->roll_step));
    rc = rollout_step_device(e, e->stg_actions, 0, nullptr, e->stg_reward, e->stg_done,
                             mask_host ? e->stg_mask : nullptr, auto_reset, obs_out_dev,
                             env_stride_floats);
    if (rc != PB_OK) return rc;
    // results device -> host on the copy stream as soon as the step kernel is done; the encoder
    // keeps running
    cudaStream_t c = e->copy_stream;
    PB_CUDA(cudaStreamWaitEvent(c, e->ev_roll_step, 0));
    unsigned long long bad = ~0ull;
    PB_CUDA(cudaMemcpyAsync(reward_host, e->stg_reward, sizeof(int32_t) * (size_t)e->n,
                            cudaMemcpyDeviceToHost, c));
    PB_CUDA(cudaMemcpyAsync(done_host, e->stg_done, (size_t)e->n, cudaMemcpyDeviceToHost, c));
    if (mask_host)
        PB_CUDA(cudaMemcpyAsync(mask_host, e->stg_mask, 5 * (size_t)e->n, cudaMemcpyDeviceToHost, c));
    PB_CUDA(cudaMemcpyAsync(&bad, e->bad_action, sizeof(bad), cudaMemcpyDeviceToHost, c));
    PB_CUDA(cudaStreamSynchronize(c));
    if (bad != ~0ull) {
        PB_CUDA(cudaMemsetAsync(e->bad_action, 0xff, sizeof(bad), e->roll_step));
        return fail(PB_ERR_INVALID_ACTION, "Invalid action");
    }
    return PB_OK;
}

int pb_rollout_chunk_step(pb_engine *e, int64_t first, int64_t count, const uint8_t *actions_dev,
                          uint64_t call_idx, uint8_t *actions_out_dev, int32_t *reward_dev,
                          uint8_t *done_dev, uint8_t *mask_out_dev, int auto_reset,
                          float *obs_out_dev, int64_t env_stride_floats, void *stream) {
    if (!e || !reward_dev || !done_dev || first < 0 || count < 0 || first + count > e->n)
        return fail(PB_ERR_INVALID_ARG, "pb_rollout_chunk_step: bad arguments");
    int rc = rollout_check_obs(obs_out_dev, env_stride_floats);
    if (rc != PB_OK) return rc;
    if (count == 0) return PB_OK;
    DeviceGuard g(e->device);
    if ((rc = rollout_ensure(e)) != PB_OK) return rc;
    if ((rc = rollout_fork(e, (cudaStream_t)stream)) != PB_OK) return rc;
    if (actions_dev) {  // the caller's stream produced the actions
        PB_CUDA(cudaEventRecord(e->ev_roll_fork, (cudaStream_t)stream));
        PB_CUDA(cudaStreamWaitEvent(e->roll_step, e->ev_roll_fork, 0));
    }
    return rollout_chunk_device(e, first, count, actions_dev, call_idx, actions_out_dev, reward_dev,
                                done_dev, mask_out_dev, auto_reset, obs_out_dev, env_stride_floats);
}

int pb_rollout_chunk_host_step(pb_engine *e, int64_t first, int64_t count,
                               const uint8_t *actions_host, int32_t *reward_host,
                               uint8_t *done_host, uint8_t *mask_host, int auto_reset,
                               float *obs_out_dev, int64_t env_stride_floats, void *stream) {
    if (!e || !actions_host || !reward_host || !done_host || first < 0 || count < 0 ||
        first + count > e->n)
        return fail(PB_ERR_INVALID_ARG, "pb_rollout_chunk_host_step: bad arguments");
    int rc = rollout_check_obs(obs_out_dev, env_stride_floats);
    if (rc != PB_OK) return rc;
    if (count == 0) return PB_OK;
    DeviceGuard g(e->device);
    if ((rc = rollout_ensure(e)) != PB_OK) return rc;
    if ((rc = rollout_fork(e, (cudaStream_t)stream)) != PB_OK) return rc;
    if (!e->host_bad) PB_CUDA(cudaMallocHost(&e->host_bad, sizeof(unsigned long long)));
    if (!e->roll_host_pending) *e->host_bad = ~0ull;
    // the staging buffers of these envs are free once their previous result copies have left
    PB_CUDA(cudaStreamWaitEvent(e->roll_step, e->ev_roll_copied, 0));
    PB_CUDA(cudaMemcpyAsync(e->stg_actions + first, actions_host + first, (size_t)count,
                            cudaMemcpyHostToDevice, e->roll_step));
    rc = rollout_chunk_device(e, first, count, e->stg_actions, 0, nullptr, e->stg_reward, e->stg_done,
                              mask_host ? e->stg_mask : nullptr, auto_reset, obs_out_dev,
                              env_stride_floats);
    if (rc != PB_OK) return rc;
    // this chunk's results leave on the copy stream while the encoders keep running
    cudaStream_t c = e->copy_stream;
    PB_CUDA(cudaStreamWaitEvent(c, e->ev_roll_step, 0));
    PB_CUDA(cudaMemcpyAsync(reward_host + first, e->stg_reward + first, sizeof(int32_t) * (size_t)count,
                            cudaMemcpyDeviceToHost, c));
    PB_CUDA(cudaMemcpyAsync(done_host + first, e->stg_done + first, (size_t)count,
                            cudaMemcpyDeviceToHost, c));
    if (mask_host)
        PB_CUDA(cudaMemcpyAsync(mask_host + 5 * first, e->stg_mask + 5 * first, 5 * (size_t)count,
                                cudaMemcpyDeviceToHost, c));
    PB_CUDA(cudaMemcpyAsync(e->host_bad, e->bad_action, sizeof(unsigned long long),
                            cudaMemcpyDeviceToHost, c));
    PB_CUDA(cudaEventRecord(e->ev_roll_copied, c));
    e->roll_host_pending = true;
    return PB_OK;
}

int pb_rollout_host_sync(pb_engine *e) {
    if (!e) return fail(PB_ERR_INVALID_ARG, "pb_rollout_host_sync: null engine");
    if (!e->roll_host_pending) return PB_OK;
    DeviceGuard g(e->device);
    e->roll_host_pending = false;
    PB_CUDA(cudaStreamSynchronize(e->copy_stream));
    if (*e->host_bad != ~0ull) {
        PB_CUDA(cudaMemsetAsync(e->bad_action, 0xff, sizeof(unsigned long long), e->roll_step));
        return fail(PB_ERR_INVALID_ACTION, "Invalid action");
    }
    return PB_OK;
}

int pb_rollout_join(pb_engine *e, void *stream) {
    if (!e) return fail(PB_ERR_INVALID_ARG, "pb_rollout_join: null engine");
    if (!e->roll_step) return PB_OK;
    DeviceGuard g(e->device);
    cudaStream_t user = (cudaStream_t)stream;
    for (cudaStream_t s : {e->roll_step, e->roll_enc[0], e->roll_enc[1], e->copy_stream}) {
        PB_CUDA(cudaEventRecord(e->ev_roll_join, s));
        PB_CUDA(cudaStreamWaitEvent(user, e->ev_roll_join, 0));
    }
    e->roll_started = false;
    return PB_OK;
}

int pb_rollout_streams(pb_engine *e, void **step_stream, void **encode_stream) {
    if (!e) return fail(PB_ERR_INVALID_ARG, "pb_rollout_streams: null engine");
    DeviceGuard g(e->device);
    int rc = rollout_ensure(e);
    if (rc != PB_OK) return rc;
    if (step_stream) *step_stream = e->roll_step;
    // the stream the most recent encoder launch went to
    if (encode_stream) *encode_stream = e->roll_enc[e->roll_last_enc];
    return PB_OK;
}

}  // extern "C"
