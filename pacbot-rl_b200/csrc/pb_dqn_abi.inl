// pb_dqn_abi.inl -- the arithmetic of a DQN iteration around the network (reference:
// /root/reference/src/algorithms/train_dqn.py:167-206); included at the end of pb_kernels.cu.
//
//   pb_dqn_target   the double-DQN target of :167-183 from the two next-state Q tables: ONE launch
//                   (torch: bitwise_not, 2 x masked_fill, argmax, gather, zeros_like, 2 x mul,
//                   where, add).  Same float32 operations in the same order: bit-identical.
//   pb_dqn_loss     Q(s)[a], the MSE loss of :193-197, its gradient with respect to the Q table
//                   (2 / B * (Q(s)[a] - target) at the taken action, 0 elsewhere) and the
//                   iteration's logged statistics (mean max Q, mean target, finite-loss flag) in
//                   ONE single-CTA launch with fixed-order sums (torch: gather, mse_loss, amax,
//                   2 x mean, isfinite, and backward mse_loss_backward, zeros, scatter).
//   pb_clip_adam    `clip_grad_norm_(parameters, max_norm)` then `Adam.step()` (:200-206) as TWO
//                   launches over the network's parameter list:
//
// torch runs this tail as ~8 launches (foreach norm, stack, norm, add, reciprocal, clamp, foreach
// mul, the fused multi-tensor Adam and its step counters): ~0.1 ms of a 2.1 ms iteration for
// 156 934 parameters, all of it launch and dependency latency.  Here:
//   k_grad_sqnorm   one CTA row per tensor: partial sums of g^2 (fixed order: deterministic);
//                   CTA (0, 0) also advances the device-resident step counter and derives the
//                   step's bias corrections (double arithmetic, once)
//   k_clip_adam     every CTA adds the partials up in the same order, derives the clip
//                   coefficient min(1, max_norm / (norm + 1e-6)) (torch/nn/utils/clip_grad.py) and
//                   applies Adam to its slice with the clipped gradient g * coef:
//                     m = lerp(m, g, 1 - beta1);  v = beta2 * v + (1 - beta2) * g * g
//                     p -= (lr / (1 - beta1^t)) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps)
//                   (torch.optim.Adam defaults: no weight decay, no amsgrad).
// The moments live in two flat float32 buffers indexed by (tensor offset + storage index), so a
// parameter and its gradient must share one memory layout (autograd's gradient layout contract;
// the Python wrapper checks the strides).  The gradients themselves are NOT rewritten with their
// clipped values (nothing reads them after the step).
//
// Numerics vs torch: the norm is one sum of squares instead of a norm of per-tensor norms, the
// bias corrections are computed in double like torch's fused kernel; tests hold the result to
// torch.optim.Adam + clip_grad_norm_ within rtol 1e-5 over many steps.

namespace pb {

constexpr int OPT_MAX_TENSORS = 32;
constexpr int OPT_ROW = 8;          // CTAs per tensor
constexpr int OPT_THREADS = 256;

struct OptTensors {
    float *p[OPT_MAX_TENSORS];
    const float *g[OPT_MAX_TENSORS];
    long long start[OPT_MAX_TENSORS + 1];   // offsets into the flat moment buffers
    int count;
};

__device__ __forceinline__ float block_sum(float v, float *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < OPT_THREADS / 32; w++) t += red[w];
    return t;
}

// partials: float32 [count * OPT_ROW] sums of squares, followed by 2 floats written by CTA (0, 0):
// the step size lr / (1 - beta1^t) and sqrt(1 - beta2^t) of THIS step (double arithmetic, once)
__global__ void __launch_bounds__(OPT_THREADS)
k_grad_sqnorm(OptTensors tl, float *__restrict__ partials, long long *step, double lr, double beta1,
              double beta2) {
    __shared__ float red[OPT_THREADS / 32];
    const int t = blockIdx.y;
    const float *g = tl.g[t];
    const long long size = tl.start[t + 1] - tl.start[t];
    float acc = 0.0f;
    for (long long i = (long long)blockIdx.x * OPT_THREADS + threadIdx.x; i < size;
         i += (long long)OPT_ROW * OPT_THREADS) {
        const float v = g[i];
        acc = fmaf(v, v, acc);
    }
    const float s = block_sum(acc, red);
    if (threadIdx.x == 0) {
        partials[t * OPT_ROW + blockIdx.x] = s;
        if (blockIdx.x == 0 && blockIdx.y == 0) {
            const long long now = *step + 1;
            *step = now;
            const double bc1 = 1.0 - pow(beta1, (double)now), bc2 = 1.0 - pow(beta2, (double)now);
            partials[tl.count * OPT_ROW] = (float)(lr / bc1);
            partials[tl.count * OPT_ROW + 1] = (float)sqrt(bc2);
        }
    }
}

// grid (columns, count): CTA (x, t) owns elements x * 1024 .. of tensor t in steps of
// gridDim.x * 1024, four independent elements per thread and step (the loads of a step are all
// requested before the first use: the kernel is latency bound, 157 K parameters in all)
constexpr int ADAM_PER_THREAD = 4;
__global__ void __launch_bounds__(OPT_THREADS)
k_clip_adam(OptTensors tl, float *__restrict__ exp_avg, float *__restrict__ exp_avg_sq,
            const float *__restrict__ partials, float max_norm, double beta1_d, double beta2_d,
            float eps, float *__restrict__ grad_norm_out) {
    const int t = blockIdx.y;
    const long long size = tl.start[t + 1] - tl.start[t];
    constexpr int CHUNK = OPT_THREADS * ADAM_PER_THREAD;
    if ((long long)blockIdx.x * CHUNK >= size) return;     // whole CTA: no barrier is skipped
    __shared__ float red[OPT_THREADS / 32];
    __shared__ float s_coef;
    const int nparts = tl.count * OPT_ROW;
    const float mine = threadIdx.x < nparts ? partials[threadIdx.x] : 0.0f;   // nparts <= 256
    const float total = block_sum(mine, red);
    if (threadIdx.x == 0) {
        const float norm = sqrtf(total);
        const float coef = max_norm / (norm + 1e-6f);
        s_coef = coef < 1.0f ? coef : 1.0f;
        if (blockIdx.x == 0 && blockIdx.y == 0 && grad_norm_out) *grad_norm_out = norm;
    }
    __syncthreads();
    const float coef = s_coef, step_size = partials[nparts], bc2_sqrt = partials[nparts + 1];
    float *p = tl.p[t];
    const float *g = tl.g[t];
    float *m = exp_avg + tl.start[t], *v = exp_avg_sq + tl.start[t];
    // the hyper-parameters are doubles like torch's; 1 - beta is rounded to float32 once, from the
    // double difference (1.0f - 0.999f is 1.3e-5 away from float(1 - 0.999))
    const float w = (float)(1.0 - beta1_d), beta2 = (float)beta2_d, one_minus_beta2 = (float)(1.0 - beta2_d);
    for (long long base = (long long)blockIdx.x * CHUNK; base < size; base += (long long)gridDim.x * CHUNK) {
        float gi[ADAM_PER_THREAD], m0[ADAM_PER_THREAD], v0[ADAM_PER_THREAD], p0[ADAM_PER_THREAD];
#pragma unroll
        for (int k = 0; k < ADAM_PER_THREAD; k++) {
            const long long i = base + k * OPT_THREADS + threadIdx.x;
            const bool in = i < size;
            gi[k] = in ? g[i] : 0.0f;
            m0[k] = in ? m[i] : 0.0f;
            v0[k] = in ? v[i] : 0.0f;
            p0[k] = in ? p[i] : 0.0f;
        }
#pragma unroll
        for (int k = 0; k < ADAM_PER_THREAD; k++) {
            const long long i = base + k * OPT_THREADS + threadIdx.x;
            if (i >= size) continue;
            const float gc = gi[k] * coef;
            // ATen lerp: start + weight * (end - start) for weight < 0.5
            const float m1 = w < 0.5f ? m0[k] + w * (gc - m0[k]) : gc - (gc - m0[k]) * (1.0f - w);
            const float v1 = beta2 * v0[k] + one_minus_beta2 * gc * gc;
            m[i] = m1;
            v[i] = v1;
            const float denom = sqrtf(v1) / bc2_sqrt + eps;
            p[i] = p0[k] - step_size * m1 / denom;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// double-DQN target (train_dqn.py:167-183) and loss (:193-197)
// ---------------------------------------------------------------------------------------------
constexpr int DQN_MAX_ACTIONS = 8;

__global__ void __launch_bounds__(256)
k_dqn_target(const float *__restrict__ q_target, const float *__restrict__ q_online,
             const uint8_t *__restrict__ mask, const float *__restrict__ reward,
             const uint8_t *__restrict__ done, float gamma, float reward_scale,
             float *__restrict__ target, long long n, int na) {
    const long long b = (long long)blockIdx.x * 256 + threadIdx.x;
    if (b >= n) return;
    const float ninf = -INFINITY;
    // next_actions = online_next_q.masked_fill(~mask, -inf).argmax(1): first maximum, NaN wins
    int best = 0;
    float best_v = 0.0f;
    for (int a = 0; a < na; a++) {
        const float v = mask[b * na + a] ? q_online[b * na + a] : ninf;
        if (a == 0 || v > best_v || (v != v && best_v == best_v)) {
            best = a;
            best_v = v;
        }
    }
    const float ret = mask[b * na + best] ? q_target[b * na + best] : ninf;
    const float discounted = done[b] ? 0.0f : __fmul_rn(gamma, ret);
    target[b] = __fadd_rn(__fmul_rn(reward[b], reward_scale), discounted);
}

// out[0] = mean((Q[b][a_b] - target[b])^2), out[1] = stat_scale * mean_b max_a Q[b][a],
// out[2] = stat_scale * mean(target), out[3] = 1 if the loss is finite else 0;
// dq[b][a] = 2 / n * (Q[b][a_b] - target[b]) for a == a_b, else 0
__global__ void __launch_bounds__(OPT_THREADS)
k_dqn_loss(const float *__restrict__ all_q, const long long *__restrict__ actions,
           const float *__restrict__ target, long long n, int na, float stat_scale,
           float *__restrict__ out, float *__restrict__ dq) {
    __shared__ float red[OPT_THREADS / 32];
    float sq = 0.0f, mx = 0.0f, tg = 0.0f;
    const float norm = 2.0f / (float)n;
    for (long long b = threadIdx.x; b < n; b += OPT_THREADS) {
        const int act = (int)actions[b];
        const float t = target[b];
        float best = all_q[b * na], pred = 0.0f;
        for (int a = 0; a < na; a++) {
            const float q = all_q[b * na + a];
            best = (q > best || q != q) ? q : best;
            pred = a == act ? q : pred;
        }
        const float d = pred - t;
        for (int a = 0; a < na; a++) dq[b * na + a] = a == act ? norm * d : 0.0f;
        sq = fmaf(d, d, sq);
        mx += best;
        tg += t;
    }
    const float s_sq = block_sum(sq, red), s_mx = block_sum(mx, red), s_tg = block_sum(tg, red);
    if (threadIdx.x == 0) {
        const float loss = s_sq / (float)n;
        out[0] = loss;
        out[1] = stat_scale * (s_mx / (float)n);
        out[2] = stat_scale * (s_tg / (float)n);
        out[3] = (loss == loss && fabsf(loss) != INFINITY) ? 1.0f : 0.0f;
    }
}

}  // namespace pb

extern "C" {

int pb_clip_adam(int count, float *const *params_dev, const float *const *grads_dev,
                 const int64_t *sizes, float *exp_avg_dev, float *exp_avg_sq_dev, int64_t *step_dev,
                 float *partials_dev, double max_norm, double lr, double beta1, double beta2, double eps,
                 float *grad_norm_out_dev, int device, void *stream) {
    using namespace pb;
    if (count <= 0 || count > OPT_MAX_TENSORS || !params_dev || !grads_dev || !sizes || !exp_avg_dev ||
        !exp_avg_sq_dev || !step_dev || !partials_dev)
        return fail(PB_ERR_INVALID_ARG, "pb_clip_adam: bad arguments (1..32 tensors per call)");
    OptTensors tl{};
    tl.count = count;
    long long off = 0;
    for (int t = 0; t < count; t++) {
        if (!params_dev[t] || !grads_dev[t] || sizes[t] <= 0)
            return fail(PB_ERR_INVALID_ARG, "pb_clip_adam: null tensor or empty tensor");
        tl.p[t] = params_dev[t];
        tl.g[t] = grads_dev[t];
        tl.start[t] = off;
        off += sizes[t];
    }
    tl.start[count] = off;
    DeviceGuard g(device);
    cudaStream_t s = (cudaStream_t)stream;
    const dim3 grid(OPT_ROW, (unsigned)count);
    k_grad_sqnorm<<<grid, OPT_THREADS, 0, s>>>(tl, partials_dev, reinterpret_cast<long long *>(step_dev), lr,
                                              beta1, beta2);
    PB_CUDA(cudaGetLastError());
    long long largest = 0;
    for (int t = 0; t < count; t++) largest = sizes[t] > largest ? sizes[t] : largest;
    constexpr long long chunk = OPT_THREADS * ADAM_PER_THREAD;
    long long cols = (largest + chunk - 1) / chunk;
    if (cols > 256) cols = 256;
    k_clip_adam<<<dim3((unsigned)cols, (unsigned)count), OPT_THREADS, 0, s>>>(
        tl, exp_avg_dev, exp_avg_sq_dev, partials_dev, (float)max_norm, beta1, beta2, (float)eps,
        grad_norm_out_dev);
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int pb_dqn_target(const float *q_target_dev, const float *q_online_dev, const uint8_t *mask_dev,
                  const float *reward_dev, const uint8_t *done_dev, double gamma, double reward_scale,
                  float *target_out_dev, int64_t n, int64_t num_actions, int device, void *stream) {
    using namespace pb;
    if (!q_target_dev || !q_online_dev || !mask_dev || !reward_dev || !done_dev || !target_out_dev || n <= 0 ||
        num_actions <= 0 || num_actions > DQN_MAX_ACTIONS)
        return fail(PB_ERR_INVALID_ARG, "pb_dqn_target: bad arguments (1..8 actions)");
    DeviceGuard g(device);
    k_dqn_target<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        q_target_dev, q_online_dev, mask_dev, reward_dev, done_dev, (float)gamma, (float)reward_scale,
        target_out_dev, n, (int)num_actions);
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int pb_dqn_loss(const float *all_q_dev, const int64_t *actions_dev, const float *target_dev, int64_t n,
                int64_t num_actions, double stat_scale, float *out4_dev, float *dq_out_dev, int device,
                void *stream) {
    using namespace pb;
    if (!all_q_dev || !actions_dev || !target_dev || !out4_dev || !dq_out_dev || n <= 0 || num_actions <= 0 ||
        num_actions > DQN_MAX_ACTIONS)
        return fail(PB_ERR_INVALID_ARG, "pb_dqn_loss: bad arguments (1..8 actions)");
    DeviceGuard g(device);
    k_dqn_loss<<<1, OPT_THREADS, 0, (cudaStream_t)stream>>>(
        all_q_dev, reinterpret_cast<const long long *>(actions_dev), target_dev, n, (int)num_actions,
        (float)stat_scale, out4_dev, dq_out_dev);
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

}  // extern "C"
