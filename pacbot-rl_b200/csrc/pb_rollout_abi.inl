// pb_rollout_abi.inl -- pb_rollout_*: pipelined rollout steps with double-buffered env state on two
// engine-owned streams (include/pacbot_b200.h); included at the end of pb_kernels.cu.
//
//     encode stream : enc(t)                enc(t+1)               ...
//     step stream   :     step(t+1)              step(t+2)         ...
//
// step(t+1) reads state buffer t and writes buffer t+1 (pb_step_flip) while enc(t) is still
// streaming buffer t to HBM; it waits only for enc(t-1), whose buffer it overwrites.  The whole
// orchestration of a step is one C call (a handful of CUDA API calls), so a busy host does not
// put bubbles into the stream of 0.52 ms encoder launches.

namespace {

int rollout_ensure(pb_engine *e) {
    if (e->roll_step) return PB_OK;
    PB_CUDA(cudaStreamCreateWithFlags(&e->roll_step, cudaStreamNonBlocking));
    PB_CUDA(cudaStreamCreateWithFlags(&e->roll_enc, cudaStreamNonBlocking));
    for (cudaEvent_t *ev : {&e->ev_roll_step, &e->ev_roll_fork, &e->ev_roll_join, &e->ev_enc_done[0],
                            &e->ev_enc_done[1]})
        PB_CUDA(cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
    return PB_OK;
}

// the engine's streams start after everything already queued on the caller's stream
int rollout_fork(pb_engine *e, cudaStream_t user) {
    if (e->roll_started) return PB_OK;
    PB_CUDA(cudaEventRecord(e->ev_roll_fork, user));
    PB_CUDA(cudaStreamWaitEvent(e->roll_step, e->ev_roll_fork, 0));
    PB_CUDA(cudaStreamWaitEvent(e->roll_enc, e->ev_roll_fork, 0));
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
    int rc = pb_step_flip(e, actions_dev, call_idx, actions_out_dev, reward_dev, done_dev, mask_dev,
                          auto_reset, e->roll_step);
    if (rc != PB_OK) return rc;
    PB_CUDA(cudaEventRecord(e->ev_roll_step, e->roll_step));
    PB_CUDA(cudaStreamWaitEvent(e->roll_enc, e->ev_roll_step, 0));
    if (obs_out_dev) {
        rc = launch_encode(e, obs_out_dev, stride, 0, e->n, e->roll_enc);
        if (rc != PB_OK) return rc;
    }
    PB_CUDA(cudaEventRecord(e->ev_enc_done[slot], e->roll_enc));
    e->roll_t += 1;
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

int pb_rollout_join(pb_engine *e, void *stream) {
    if (!e) return fail(PB_ERR_INVALID_ARG, "pb_rollout_join: null engine");
    if (!e->roll_step) return PB_OK;
    DeviceGuard g(e->device);
    cudaStream_t user = (cudaStream_t)stream;
    for (cudaStream_t s : {e->roll_step, e->roll_enc, e->copy_stream}) {
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
    if (encode_stream) *encode_stream = e->roll_enc;
    return PB_OK;
}

}  // extern "C"
