// pb_qnet_abi.inl -- pb_bias_pool_silu_*: the elementwise tail of a QNetV2 convolution block
// (reference: /root/reference/src/models.py:70-90, `Conv2d -> MaxPool2d(2, ceil_mode=True) -> SiLU`
// and the first block `Conv2d -> SiLU`) as ONE kernel forward and ONE backward, channels-last
// float32; included at the end of pb_kernels.cu.  The convolution itself stays cuDNN.
//
// torch runs the tail as separate kernels: add_ (the bias; cuDNN's convolution does not fuse it),
// max_pool2d_with_indices, silu -- and backward silu_backward, max_pool2d_with_indices_backward,
// a sum for the bias gradient.  Each of them streams the activation through HBM again; at batch
// 512 that was 0.9 ms of a 3.3 ms DQN iteration (profiles/r02_dqn_torch_profiler.txt).  Fused:
//   forward   z = max over the 2x2 window (partial at the border: ceil_mode) of (x + bias),
//             y = z / (1 + exp(-z)); writes z (the SiLU input, needed by backward), y and the
//             arg-max (0..3) -- reads x once, writes a quarter of it
//   backward  dz = dy * s * (1 + z * (1 - s)), s = 1 / (1 + exp(-z)); dx = dz at the window's
//             arg-max and 0 at its other cells (2x2 windows with stride 2 tile the map: every
//             cell of dx is written exactly once, no zero-fill, no atomics); dbias = per-channel
//             sum of dz (per-thread partial sums, one shared-memory reduction and one atomicAdd
//             per channel per CTA)
// Same arithmetic as the torch composition: fl(a + b) is monotone in a, so pooling x + bias
// equals pooling x and adding the bias; the arg-max is taken over the SUMS with torch's rule
// (first maximum in row-major window order, NaN wins); SiLU and its derivative are written as in
// ATen (activation kernels: x / (1 + exp(-x)); dy * s * (1 + x * (1 - s))).

namespace pb {

constexpr int QN_THREADS = 256;

__device__ __forceinline__ float silu_f(float z) { return z / (1.0f + expf(-z)); }
__device__ __forceinline__ float silu_grad_f(float dy, float z) {
    const float s = 1.0f / (1.0f + expf(-z));
    return dy * s * (1.0f + z * (1.0f - s));
}
__device__ __forceinline__ void take_max(float v, int k, float &best, int &arg) {
    if (v > best || v != v) {  // ATen max_pool2d: (val > maxval) || isnan(val)
        best = v;
        arg = k;
    }
}

// one thread per (n, ho, wo, 4 channels); x / z / y are [n][h][w][c] resp. [n][ho][wo][c]
template <bool POOL>
__global__ void __launch_bounds__(QN_THREADS)
k_bias_pool_silu_fwd(const float4 *__restrict__ x, const float4 *__restrict__ bias, long long n, int h,
                     int w, int c4, float4 *__restrict__ z_out, float4 *__restrict__ y_out,
                     uchar4 *__restrict__ idx_out) {
    const int ho = POOL ? (h + 1) / 2 : h, wo = POOL ? (w + 1) / 2 : w;
    const long long total = n * ho * wo * c4;
    for (long long t = (long long)blockIdx.x * QN_THREADS + threadIdx.x; t < total;
         t += (long long)gridDim.x * QN_THREADS) {
        const int cc = (int)(t % c4);
        long long r = t / c4;
        const int ox = (int)(r % wo);
        r /= wo;
        const int oy = (int)(r % ho);
        const long long img = r / ho;
        const float4 b = __ldg(bias + cc);
        float4 z;
        if (POOL) {
            const float ninf = -INFINITY;
            float bx = ninf, by = ninf, bz = ninf, bw = ninf;
            int ax = 0, ay = 0, az = 0, aw = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int iy = 2 * oy + (k >> 1), ix = 2 * ox + (k & 1);
                if (iy < h && ix < w) {
                    const float4 v = __ldg(x + ((img * h + iy) * w + ix) * c4 + cc);
                    take_max(v.x + b.x, k, bx, ax);
                    take_max(v.y + b.y, k, by, ay);
                    take_max(v.z + b.z, k, bz, az);
                    take_max(v.w + b.w, k, bw, aw);
                }
            }
            z = make_float4(bx, by, bz, bw);
            if (idx_out) idx_out[t] = make_uchar4((unsigned char)ax, (unsigned char)ay, (unsigned char)az, (unsigned char)aw);
        } else {
            const float4 v = __ldg(x + t);
            z = make_float4(v.x + b.x, v.y + b.y, v.z + b.z, v.w + b.w);
        }
        if (z_out) z_out[t] = z;
        y_out[t] = make_float4(silu_f(z.x), silu_f(z.y), silu_f(z.z), silu_f(z.w));
    }
}

// grid-stride with gridDim.x * QN_THREADS a multiple of c4: a thread always works on the same four
// channels, so the bias gradient accumulates in registers
template <bool POOL>
__global__ void __launch_bounds__(QN_THREADS)
k_bias_pool_silu_bwd(const float4 *__restrict__ dy, const float4 *__restrict__ z,
                     const uchar4 *__restrict__ idx, long long n, int h, int w, int c4,
                     float4 *__restrict__ dx, float *__restrict__ dbias) {
    __shared__ float4 s_part[QN_THREADS];
    const int ho = POOL ? (h + 1) / 2 : h, wo = POOL ? (w + 1) / 2 : w;
    const long long total = n * ho * wo * c4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long t = (long long)blockIdx.x * QN_THREADS + threadIdx.x; t < total;
         t += (long long)gridDim.x * QN_THREADS) {
        const float4 g = __ldg(dy + t), zz = __ldg(z + t);
        const float4 dz = make_float4(silu_grad_f(g.x, zz.x), silu_grad_f(g.y, zz.y),
                                      silu_grad_f(g.z, zz.z), silu_grad_f(g.w, zz.w));
        acc.x += dz.x;
        acc.y += dz.y;
        acc.z += dz.z;
        acc.w += dz.w;
        if (POOL) {
            const int cc = (int)(t % c4);
            long long r = t / c4;
            const int ox = (int)(r % wo);
            r /= wo;
            const int oy = (int)(r % ho);
            const long long img = r / ho;
            const uchar4 a = idx[t];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int iy = 2 * oy + (k >> 1), ix = 2 * ox + (k & 1);
                if (iy < h && ix < w)
                    dx[((img * h + iy) * w + ix) * c4 + cc] =
                        make_float4(a.x == k ? dz.x : 0.f, a.y == k ? dz.y : 0.f, a.z == k ? dz.z : 0.f,
                                    a.w == k ? dz.w : 0.f);
            }
        } else {
            dx[t] = dz;
        }
    }
    // threads tid, tid + c4, tid + 2 c4, ... of a CTA hold the same channels (QN_THREADS % c4 == 0)
    s_part[threadIdx.x] = acc;
    __syncthreads();
    if ((int)threadIdx.x < c4) {
        float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j = threadIdx.x; j < QN_THREADS; j += c4) {
            const float4 v = s_part[j];
            sum.x += v.x;
            sum.y += v.y;
            sum.z += v.z;
            sum.w += v.w;
        }
        // the first channel group of this CTA: (blockIdx.x * QN_THREADS) % c4 == 0, so thread tid
        // works on channel group tid
        float *o = dbias + 4 * threadIdx.x;
        atomicAdd(o + 0, sum.x);
        atomicAdd(o + 1, sum.y);
        atomicAdd(o + 2, sum.z);
        atomicAdd(o + 3, sum.w);
    }
}


// ---- the head of QNetV2 / NetV2 as one kernel (forward, inference) -----------------------------
// reference models.py:92-110: `AdaptiveMaxPool2d((1, 1)) -> Flatten -> SiLU -> Linear(128, 256) ->
// SiLU`, then the dueling combine `V - mean(A) + A` of two linear heads (QNetV2) or one plain
// linear layer (NetV2, models.py:139-141).  torch runs this as ~11 launches per forward (max-pool,
// silu, addmm + epilogue, silu, two addmm, mean, sub, add, the grouped convolution's bias add) on
// [N, 128] / [N, 256] tensors: at batch 32 (the four policy forwards of an iteration, every
// evaluation step) and 512 (the two target forwards) that is launch latency, not work.  Here:
//   p[c] = max over the map of x[n, :, c] + conv_bias[c]   (fl(a + b) is monotone in a: equals
//          the max of the biased map, with torch's NaN rule)
//   a = SiLU(p); u = W1 a + b1; f = SiLU(u); head rows = Wh f + bh; Q = V - mean(A) + A
// One CTA handles S samples: W1 (128 KB, L2 resident) is read once per CTA, row by row, a warp
// per row with the 128 inputs split over the lanes (coalesced 512 B loads) and a shuffle tree.
// float32 FMA throughout; the sums are ordered differently from cuBLAS (tests: rtol 1e-5).
constexpr int HEAD_C = 128, HEAD_H = 256, HEAD_THREADS = 512, HEAD_MAX_OUT = 8;
constexpr int HEAD_WARPS = HEAD_THREADS / 32, HEAD_RB = 8;      // rows of W1 in flight per warp
constexpr int HEAD_PGROUPS = HEAD_THREADS / (HEAD_C / 4);   // 16 map positions in flight per sample

// TRAIN: additionally keeps what the backward kernels need -- p_out [n][128] (the pooled maximum +
// bias, i.e. the first SiLU's input), idx_out [n][128] (map position of the maximum: the first
// one in row-major order, torch's choice), u_out [n][256] (the second SiLU's input).
template <int S, bool TRAIN>
__global__ void __launch_bounds__(HEAD_THREADS)
k_qnet_head_fwd(const float *__restrict__ x, int hw, const float *__restrict__ conv_bias,
                const float *__restrict__ w1, const float *__restrict__ b1,
                const float *__restrict__ wv, const float *__restrict__ bv,
                const float *__restrict__ wa, const float *__restrict__ ba, int n_out, int dueling,
                long long n, float *__restrict__ out, float *__restrict__ p_out,
                unsigned char *__restrict__ idx_out, float *__restrict__ u_out,
                const unsigned char *__restrict__ act_mask, unsigned char *__restrict__ act_out,
                float *__restrict__ pv_value, float *__restrict__ pv_policy, float value_scale) {
    __shared__ __align__(16) float s_a[S][HEAD_C];
    __shared__ __align__(16) float s_m[S][HEAD_PGROUPS][HEAD_C];
    __shared__ unsigned char s_pos[TRAIN ? S : 1][HEAD_PGROUPS][HEAD_C];
    __shared__ float s_f[S][HEAD_H];
    __shared__ float s_h[S][HEAD_MAX_OUT + 1];
    __shared__ float s_b1[HEAD_H];
    const long long n0 = (long long)blockIdx.x * S;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float4 wrow[HEAD_RB];
#pragma unroll
    for (int r = 0; r < HEAD_RB; r++)
        wrow[r] = __ldg(reinterpret_cast<const float4 *>(w1 + (long long)(r * HEAD_WARPS + warp) * HEAD_C) + lane);
    if (threadIdx.x < HEAD_H) s_b1[threadIdx.x] = __ldg(b1 + threadIdx.x);
    // global max over the map: thread (g, c4) takes positions g, g + 16, ... of four channels
    // (independent 16-byte loads, 512 B per position across a warp), then the 16 partial maxima
    // of a channel meet in shared memory.  max is exact, so the grouping does not matter; a NaN
    // anywhere in the map wins as in ATen's max_pool2d.
    {
        const int c4 = threadIdx.x % (HEAD_C / 4), g = threadIdx.x / (HEAD_C / 4);
#pragma unroll
        for (int smp = 0; smp < S; smp++) {
            const long long i = n0 + smp;
            float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
            int ax = g, ay = g, az = g, aw = g;   // TRAIN: position of each channel's maximum so far
            if (i < n) {
                const float4 *src = reinterpret_cast<const float4 *>(x + i * hw * HEAD_C) + c4;
#pragma unroll 4
                for (int p = g; p < hw; p += HEAD_PGROUPS) {
                    const float4 v = __ldg(src + (long long)p * (HEAD_C / 4));
                    if (v.x > m.x || v.x != v.x) { m.x = v.x; ax = p; }
                    if (v.y > m.y || v.y != v.y) { m.y = v.y; ay = p; }
                    if (v.z > m.z || v.z != v.z) { m.z = v.z; az = p; }
                    if (v.w > m.w || v.w != v.w) { m.w = v.w; aw = p; }
                }
            }
            *reinterpret_cast<float4 *>(&s_m[smp][g][4 * c4]) = m;
            if (TRAIN)
                *reinterpret_cast<uchar4 *>(&s_pos[smp][g][4 * c4]) =
                    make_uchar4((unsigned char)ax, (unsigned char)ay, (unsigned char)az, (unsigned char)aw);
        }
        __syncthreads();
        for (int t = threadIdx.x; t < S * HEAD_C; t += HEAD_THREADS) {
            const int smp = t / HEAD_C, c = t % HEAD_C;
            float best = s_m[smp][0][c];
            int bpos = TRAIN ? s_pos[smp][0][c] : 0;
#pragma unroll
            for (int k = 1; k < HEAD_PGROUPS; k++) {
                const float v = s_m[smp][k][c];
                if (TRAIN) {   // equal maxima: the smaller map position (first in row-major order)
                    const int vp = s_pos[smp][k][c];
                    if (best == best && (v > best || v != v || (v == best && vp < bpos))) {
                        best = v;
                        bpos = vp;
                    }
                } else if (best == best && (v > best || v != v)) {
                    best = v;
                }
            }
            const bool live = n0 + smp < n;
            const float pre = best + __ldg(conv_bias + c);
            s_a[smp][c] = live ? silu_f(pre) : 0.f;
            if (TRAIN && live) {
                p_out[(n0 + smp) * HEAD_C + c] = pre;
                idx_out[(n0 + smp) * HEAD_C + c] = (unsigned char)bpos;
            }
        }
        __syncthreads();
    }
    // Linear 128 -> 256: warp w owns rows w, w + 16, ...; the rows of a batch of HEAD_RB are all
    // requested before any of them is used (their first batch was requested at the top of the
    // kernel, under the max pooling), lanes split the 128 inputs, a butterfly sums them, and
    // lane r finishes row r of the batch (bias, SiLU) so that the tail is not serialised either
    float4 av[S];
#pragma unroll
    for (int smp = 0; smp < S; smp++) av[smp] = *reinterpret_cast<const float4 *>(&s_a[smp][4 * lane]);
#pragma unroll
    for (int jb = 0; jb < HEAD_H / (HEAD_WARPS * HEAD_RB); jb++) {
        float acc[HEAD_RB][S];
#pragma unroll
        for (int r = 0; r < HEAD_RB; r++)
#pragma unroll
            for (int smp = 0; smp < S; smp++)
                acc[r][smp] = fmaf(wrow[r].w, av[smp].w,
                                   fmaf(wrow[r].z, av[smp].z, fmaf(wrow[r].y, av[smp].y, wrow[r].x * av[smp].x)));
        if (jb + 1 < HEAD_H / (HEAD_WARPS * HEAD_RB)) {
#pragma unroll
            for (int r = 0; r < HEAD_RB; r++)
                wrow[r] = __ldg(reinterpret_cast<const float4 *>(
                                    w1 + (long long)(((jb + 1) * HEAD_RB + r) * HEAD_WARPS + warp) * HEAD_C) + lane);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
#pragma unroll
            for (int r = 0; r < HEAD_RB; r++)
#pragma unroll
                for (int smp = 0; smp < S; smp++)
                    acc[r][smp] += __shfl_xor_sync(0xffffffffu, acc[r][smp], off);
        const int row = (jb * HEAD_RB + (lane & (HEAD_RB - 1))) * HEAD_WARPS + warp;
#pragma unroll
        for (int smp = 0; smp < S; smp++) {
            float v = acc[0][smp];
#pragma unroll
            for (int r = 1; r < HEAD_RB; r++)
                if ((lane & (HEAD_RB - 1)) == r) v = acc[r][smp];
            if (lane < HEAD_RB) {
                const float pre = v + s_b1[row];
                s_f[smp][row] = silu_f(pre);
                if (TRAIN && n0 + smp < n) u_out[(n0 + smp) * HEAD_H + row] = pre;
            }
        }
    }
    __syncthreads();
    // head rows: row 0 = the value head (dueling), rows 1.. = the advantage / output rows
    const int rows = n_out + (dueling ? 1 : 0);
    for (int r = warp; r < rows; r += HEAD_THREADS / 32) {
        const bool is_v = dueling && r == 0;
        const float *wr = is_v ? wv : wa + (long long)(r - (dueling ? 1 : 0)) * HEAD_H;
        const float bias = is_v ? __ldg(bv) : __ldg(ba + r - (dueling ? 1 : 0));
#pragma unroll
        for (int smp = 0; smp < S; smp++) {
            float acc = 0.f;
#pragma unroll
            for (int t = 0; t < HEAD_H / 32; t++) acc = fmaf(__ldg(wr + lane + 32 * t), s_f[smp][lane + 32 * t], acc);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
            if (lane == 0) s_h[smp][r] = acc + bias;
        }
    }
    __syncthreads();
    float q = 0.f;
    const bool q_thread = (int)threadIdx.x < S * n_out && n0 + threadIdx.x / n_out < n;
    if (q_thread) {
        const int smp = threadIdx.x / n_out, k = threadIdx.x % n_out;
        if (dueling) {   // value - adv.mean(dim=1, keepdim=True) + adv  (models.py:108-110)
            float sum = 0.f;
            for (int t = 0; t < n_out; t++) sum += s_h[smp][1 + t];
            q = (s_h[smp][0] - sum / (float)n_out) + s_h[smp][1 + k];
        } else {
            q = s_h[smp][k];
        }
        out[(n0 + smp) * n_out + k] = q;
    }
    if (act_out) {   // the Q-values take the place of the advantage rows for the arg-max below
        __syncthreads();
        if (q_thread) s_h[threadIdx.x / n_out][(dueling ? 1 : 0) + threadIdx.x % n_out] = q;
    }
    // The AlphaZero evaluator's post-processing (src/algorithms/train_alphazero.py:83-95) in the same
    // launch, plain head only: value = last row * value_scale, policy = softmax over the other
    // rows with invalid actions at -inf (exp(x - max) / sum, as torch.softmax).
    if (pv_policy) {
        if ((int)threadIdx.x < S && n0 + threadIdx.x < n) {
            const int smp = threadIdx.x, na = n_out - 1;
            const long long i = n0 + smp;
            float mx = -INFINITY;
            for (int k = 0; k < na; k++)
                if (act_mask[i * na + k]) mx = fmaxf(mx, s_h[smp][k]);
            float e[HEAD_MAX_OUT], sum = 0.f;
            for (int k = 0; k < na; k++) {
                e[k] = act_mask[i * na + k] ? expf(s_h[smp][k] - mx) : 0.f;
                sum += e[k];
            }
            for (int k = 0; k < na; k++) pv_policy[i * na + k] = e[k] / sum;
            pv_value[i] = s_h[smp][na] * value_scale;
        }
        return;
    }
    // MaxQPolicy (src/policies.py:44-59) in the same launch: q.masked_fill(~mask, -inf).argmax(1),
    // first maximum, a NaN counts as the maximum, 0 when every action is masked
    if (act_out) {
        __syncthreads();
        if ((int)threadIdx.x < S && n0 + threadIdx.x < n) {
            const int smp = threadIdx.x;
            const long long i = n0 + smp;
            float best = -INFINITY;
            int arg = 0;
            for (int k = 0; k < n_out; k++) {
                const float v = act_mask[i * n_out + k] ? s_h[smp][(dueling ? 1 : 0) + k] : -INFINITY;
                if (best == best && (v > best || v != v)) {
                    best = v;
                    arg = k;
                }
            }
            act_out[i] = (unsigned char)arg;
        }
    }
}

// ---- backward of the head (training forward = k_qnet_head_fwd<S, true>) ------------------------
// k_qnet_head_bwd: the per-sample chain, one CTA per S samples like the forward:
//   dh   = d(head rows): dueling  dV = sum_k dq_k, dA_k = dq_k - (sum_j dq_j) / A; plain  dq
//   df_j = sum_r Wh[r][j] dh_r;  du_j = df_j SiLU'(u_j)
//   da_c = sum_j W1[j][c] du_j   (column walk of W1: row j is one coalesced 512 B load, four
//          groups of 64 rows in flight over the CTA, partial sums meet in shared memory)
//   dp_c = da_c SiLU'(p_c);  dx[pos][c] = dp_c at the arg-max position, 0 elsewhere -- every
//          cell of dx is written exactly once, no zero fill
// and it leaves du, dh, dp plus the recomputed activations a = SiLU(p), f = SiLU(u) for
// k_qnet_head_wgrad, which reduces over the batch: dW1 = du^T a (32 x 32 tiles, shared-memory
// staged), db1, dWh = dh^T f, dbh, and the convolution bias gradient sum_n dp.
constexpr int HEAD_R = HEAD_MAX_OUT;   // head rows incl. the value row (padded to 8)

template <int S>
__global__ void __launch_bounds__(HEAD_THREADS)
k_qnet_head_bwd(const float *__restrict__ dq, const float *__restrict__ p, const unsigned char *__restrict__ idx,
                const float *__restrict__ u, const float *__restrict__ w1, const float *__restrict__ wv,
                const float *__restrict__ wa, int hw, int n_out, int dueling, long long n,
                float *__restrict__ dx, float *__restrict__ dp_out, float *__restrict__ du_out,
                float *__restrict__ a_out, float *__restrict__ f_out, float *__restrict__ dh_out) {
    __shared__ float s_dh[S][HEAD_R];
    __shared__ float s_du[S][HEAD_H];
    __shared__ float s_part[HEAD_THREADS / HEAD_C][S][HEAD_C];
    __shared__ float s_dp[S][HEAD_C];
    __shared__ unsigned char s_idx[S][HEAD_C];
    const long long n0 = (long long)blockIdx.x * S;
    const int rows = n_out + (dueling ? 1 : 0);
    if ((int)threadIdx.x < S * HEAD_R) {
        const int smp = threadIdx.x / HEAD_R, r = threadIdx.x % HEAD_R;
        const long long i = n0 + smp;
        float v = 0.f;
        if (i < n && r < rows) {
            const float *g = dq + i * n_out;
            if (dueling) {
                float sum = 0.f;
                for (int k = 0; k < n_out; k++) sum += g[k];
                v = r == 0 ? sum : g[r - 1] - sum / (float)n_out;
            } else {
                v = g[r];
            }
        }
        s_dh[smp][r] = v;
        if (i < n) dh_out[i * HEAD_R + r] = v;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < S * HEAD_H; t += HEAD_THREADS) {
        const int smp = t / HEAD_H, j = t % HEAD_H;
        const long long i = n0 + smp;
        float du = 0.f;
        if (i < n) {
            const float uu = __ldg(u + i * HEAD_H + j);
            float w[HEAD_R];
#pragma unroll
            for (int r = 0; r < HEAD_R; r++)   // independent loads, requested together
                w[r] = r >= rows ? 0.f
                       : (dueling && r == 0) ? __ldg(wv + j)
                                             : __ldg(wa + (long long)(r - (dueling ? 1 : 0)) * HEAD_H + j);
            float df = 0.f;
#pragma unroll
            for (int r = 0; r < HEAD_R; r++) df = fmaf(w[r], s_dh[smp][r], df);
            du = silu_grad_f(df, uu);
            du_out[i * HEAD_H + j] = du;
            f_out[i * HEAD_H + j] = silu_f(uu);
        }
        s_du[smp][j] = du;
    }
    __syncthreads();
    {
        const int c = threadIdx.x % HEAD_C, jg = threadIdx.x / HEAD_C;   // 4 groups of 64 rows
        constexpr int ROWS_PER = HEAD_H / (HEAD_THREADS / HEAD_C);
        float acc[S];
#pragma unroll
        for (int smp = 0; smp < S; smp++) acc[smp] = 0.f;
#pragma unroll 8
        for (int jj = 0; jj < ROWS_PER; jj++) {
            const int j = jg * ROWS_PER + jj;
            const float w = __ldg(w1 + (long long)j * HEAD_C + c);
#pragma unroll
            for (int smp = 0; smp < S; smp++) acc[smp] = fmaf(w, s_du[smp][j], acc[smp]);
        }
#pragma unroll
        for (int smp = 0; smp < S; smp++) s_part[jg][smp][c] = acc[smp];
    }
    __syncthreads();
    for (int t = threadIdx.x; t < S * HEAD_C; t += HEAD_THREADS) {
        const int smp = t / HEAD_C, c = t % HEAD_C;
        const long long i = n0 + smp;
        float dp = 0.f;
        unsigned char pos = 0;
        if (i < n) {
            float da = 0.f;
#pragma unroll
            for (int k = 0; k < HEAD_THREADS / HEAD_C; k++) da += s_part[k][smp][c];
            const float pp = __ldg(p + i * HEAD_C + c);
            dp = silu_grad_f(da, pp);
            dp_out[i * HEAD_C + c] = dp;
            a_out[i * HEAD_C + c] = silu_f(pp);
            pos = idx[i * HEAD_C + c];
        }
        s_dp[smp][c] = dp;
        s_idx[smp][c] = pos;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < S * hw * (HEAD_C / 4); t += HEAD_THREADS) {
        const int c4 = t % (HEAD_C / 4);
        const int r = t / (HEAD_C / 4);
        const int pos = r % hw, smp = r / hw;
        const long long i = n0 + smp;
        if (i >= n) continue;
        const int c = 4 * c4;
        const float4 v = make_float4(s_idx[smp][c] == pos ? s_dp[smp][c] : 0.f,
                                     s_idx[smp][c + 1] == pos ? s_dp[smp][c + 1] : 0.f,
                                     s_idx[smp][c + 2] == pos ? s_dp[smp][c + 2] : 0.f,
                                     s_idx[smp][c + 3] == pos ? s_dp[smp][c + 3] : 0.f);
        reinterpret_cast<float4 *>(dx + (i * hw + pos) * HEAD_C)[c4] = v;
    }
}

constexpr int WG_THREADS = 256, WG_T = 32, WG_NS = 64;   // 32 x 32 output tile, 64 samples per stage
constexpr int WG_TILES = (HEAD_H / WG_T) * (HEAD_C / WG_T);   // 32 tiles of dW1
// blocks [0, WG_TILES): dW1 tiles (+ db1 in the tiles of the first column block);
// block WG_TILES: dWh / dbh; block WG_TILES + 1: the convolution bias gradient.
// Every block walks the batch in stages whose global loads are all requested before the first one
// is used (16 per thread in the tile blocks, 32 in the head-row block, 16 in the bias block): the
// kernel is a chain of L2 round trips, so the number of stages is what it costs.
__global__ void __launch_bounds__(WG_THREADS)
k_qnet_head_wgrad(const float *__restrict__ du, const float *__restrict__ a, const float *__restrict__ f,
                  const float *__restrict__ dh, const float *__restrict__ dp, long long n, int n_out,
                  int dueling, float *__restrict__ dw1, float *__restrict__ db1, float *__restrict__ dwv,
                  float *__restrict__ dbv, float *__restrict__ dwa, float *__restrict__ dba,
                  float *__restrict__ dconv_bias) {
    __shared__ float s_x[WG_NS][WG_T + 1];
    __shared__ float s_y[WG_NS][WG_T + 1];
    const int tid = threadIdx.x;
    if ((int)blockIdx.x < WG_TILES) {
        const int jt = blockIdx.x / (HEAD_C / WG_T), ct = blockIdx.x % (HEAD_C / WG_T);
        const int c = tid % WG_T, tj = tid / WG_T;   // rows tj, tj + 8, tj + 16, tj + 24 of the tile
        float acc[4] = {0.f, 0.f, 0.f, 0.f}, accb[4] = {0.f, 0.f, 0.f, 0.f};
        for (long long nb = 0; nb < n; nb += WG_NS) {
            float vx[WG_NS / 8], vy[WG_NS / 8];
#pragma unroll
            for (int k = 0; k < WG_NS / 8; k++) {   // 64 samples x 32 columns of du and of a
                const long long i = nb + tj + 8 * k;
                vx[k] = i < n ? __ldg(du + i * HEAD_H + jt * WG_T + c) : 0.f;
                vy[k] = i < n ? __ldg(a + i * HEAD_C + ct * WG_T + c) : 0.f;
            }
#pragma unroll
            for (int k = 0; k < WG_NS / 8; k++) {
                s_x[tj + 8 * k][c] = vx[k];
                s_y[tj + 8 * k][c] = vy[k];
            }
            __syncthreads();
#pragma unroll 8
            for (int nn = 0; nn < WG_NS; nn++) {
                const float y = s_y[nn][c];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float x = s_x[nn][tj + 8 * k];
                    acc[k] = fmaf(x, y, acc[k]);
                    accb[k] += x;
                }
            }
            __syncthreads();
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int j = jt * WG_T + tj + 8 * k;
            dw1[(long long)j * HEAD_C + ct * WG_T + c] = acc[k];
            if (ct == 0 && c == 0) db1[j] = accb[k];
        }
        return;
    }
    if ((int)blockIdx.x == WG_TILES) {   // head rows: thread j owns column j of every row
        const int rows = n_out + (dueling ? 1 : 0);
        float acc[HEAD_R], accb = 0.f;
#pragma unroll
        for (int r = 0; r < HEAD_R; r++) acc[r] = 0.f;
        float *s_dh = &s_x[0][0];   // 64 samples x 8 rows per stage
        for (long long nb = 0; nb < n; nb += WG_NS) {
            float fv[WG_NS];
#pragma unroll
            for (int nn = 0; nn < WG_NS; nn++) fv[nn] = (nb + nn < n) ? __ldg(f + (nb + nn) * HEAD_H + tid) : 0.f;
#pragma unroll
            for (int k = 0; k < WG_NS * HEAD_R / WG_THREADS; k++) {
                const int t = tid + k * WG_THREADS;
                const long long i = nb + t / HEAD_R;
                s_dh[t] = i < n ? __ldg(dh + i * HEAD_R + t % HEAD_R) : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int nn = 0; nn < WG_NS; nn++) {
#pragma unroll
                for (int r = 0; r < HEAD_R; r++) acc[r] = fmaf(s_dh[nn * HEAD_R + r], fv[nn], acc[r]);
                if (tid < HEAD_R) accb += s_dh[nn * HEAD_R + tid];
            }
            __syncthreads();
        }
#pragma unroll
        for (int r = 0; r < HEAD_R; r++) {
            if (r >= rows) break;
            if (dueling && r == 0) dwv[tid] = acc[r];
            else dwa[(long long)(r - (dueling ? 1 : 0)) * HEAD_H + tid] = acc[r];
        }
        if (tid < rows) {
            if (dueling && tid == 0) dbv[0] = accb;
            else dba[tid - (dueling ? 1 : 0)] = accb;
        }
        return;
    }
    {   // convolution bias gradient: two halves of the batch per channel, 16 loads in flight each
        const int c = tid % HEAD_C, half = tid / HEAD_C;
        float acc = 0.f;
        for (long long i0 = half; i0 < n; i0 += 32) {
            float v[16];
#pragma unroll
            for (int k = 0; k < 16; k++) v[k] = (i0 + 2 * k < n) ? __ldg(dp + (i0 + 2 * k) * HEAD_C + c) : 0.f;
#pragma unroll
            for (int k = 0; k < 16; k++) acc += v[k];
        }
        float *s_half = &s_x[0][0];
        if (half == 1) s_half[c] = acc;
        __syncthreads();
        if (half == 0) dconv_bias[c] = acc + s_half[c];
    }
}

}  // namespace pb

namespace {
int qnet_check(const void *a, const void *b, int64_t n, int64_t h, int64_t w, int64_t c, const char *what) {
    if (!a || !b || n <= 0 || h <= 0 || w <= 0 || c <= 0 || (c & 3) || QN_THREADS % (c / 4) != 0 ||
        h > 32767 || w > 32767 || !aligned16(a) || !aligned16(b))
        return fail(PB_ERR_INVALID_ARG, std::string(what) + ": bad arguments (channels-last float32, channels "
                                                           "a multiple of 4 that divides 1024, 16-byte aligned)");
    return PB_OK;
}
unsigned qnet_grid(long long total, int num_sms) {
    const long long need = (total + QN_THREADS - 1) / QN_THREADS;
    const long long cap = (long long)num_sms * 8;
    return (unsigned)(need < cap ? need : cap);
}
int qnet_head_launch(const float *x_dev, int64_t n, int64_t hw, const float *conv_bias_dev,
                     const float *w1_dev, const float *b1_dev, const float *wv_dev, const float *bv_dev,
                     const float *wa_dev, const float *ba_dev, int64_t n_out, int dueling, float *out_dev,
                     float *p_out_dev, uint8_t *idx_out_dev, float *u_out_dev, const uint8_t *mask_dev,
                     uint8_t *actions_out_dev, int device, void *stream, float *pv_value_dev = nullptr,
                     float *pv_policy_dev = nullptr, float value_scale = 1.0f);
int qnet_sms(int device) {
    static thread_local int cached_dev = -1, sms = 148;
    if (cached_dev != device) {
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) sms = 148;
        cached_dev = device;
    }
    return sms;
}
}  // namespace

extern "C" {

int pb_bias_pool_silu_fwd(const float *x_dev, const float *bias_dev, int64_t n, int64_t h, int64_t w,
                          int64_t c, int pool, float *z_out_dev, float *y_out_dev, uint8_t *idx_out_dev,
                          int device, void *stream) {
    int rc = qnet_check(x_dev, y_out_dev, n, h, w, c, "pb_bias_pool_silu_fwd");
    if (rc != PB_OK) return rc;
    if (!bias_dev || !aligned16(bias_dev) || (z_out_dev && !aligned16(z_out_dev)) ||
        (idx_out_dev && (reinterpret_cast<uintptr_t>(idx_out_dev) & 3)))
        return fail(PB_ERR_INVALID_ARG, "pb_bias_pool_silu_fwd: bad arguments");
    DeviceGuard g(device);
    const int c4 = (int)(c / 4);
    const long long ho = pool ? (h + 1) / 2 : h, wo = pool ? (w + 1) / 2 : w;
    const unsigned grid = qnet_grid(n * ho * wo * c4, qnet_sms(device));
    auto xs = reinterpret_cast<const float4 *>(x_dev);
    auto bs = reinterpret_cast<const float4 *>(bias_dev);
    auto zs = reinterpret_cast<float4 *>(z_out_dev);
    auto ys = reinterpret_cast<float4 *>(y_out_dev);
    if (pool)
        k_bias_pool_silu_fwd<true><<<grid, QN_THREADS, 0, (cudaStream_t)stream>>>(
            xs, bs, n, (int)h, (int)w, c4, zs, ys, reinterpret_cast<uchar4 *>(idx_out_dev));
    else
        k_bias_pool_silu_fwd<false><<<grid, QN_THREADS, 0, (cudaStream_t)stream>>>(
            xs, bs, n, (int)h, (int)w, c4, zs, ys, nullptr);
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int pb_bias_pool_silu_bwd(const float *dy_dev, const float *z_dev, const uint8_t *idx_dev, int64_t n,
                          int64_t h, int64_t w, int64_t c, int pool, float *dx_out_dev,
                          float *dbias_out_dev, int device, void *stream) {
    int rc = qnet_check(dy_dev, z_dev, n, h, w, c, "pb_bias_pool_silu_bwd");
    if (rc != PB_OK) return rc;
    if (!dx_out_dev || !dbias_out_dev || !aligned16(dx_out_dev) || (pool && !idx_dev))
        return fail(PB_ERR_INVALID_ARG, "pb_bias_pool_silu_bwd: bad arguments");
    DeviceGuard g(device);
    const int c4 = (int)(c / 4);
    const long long ho = pool ? (h + 1) / 2 : h, wo = pool ? (w + 1) / 2 : w;
    const unsigned grid = qnet_grid(n * ho * wo * c4, qnet_sms(device));
    PB_CUDA(cudaMemsetAsync(dbias_out_dev, 0, sizeof(float) * (size_t)c, (cudaStream_t)stream));
    auto gs = reinterpret_cast<const float4 *>(dy_dev);
    auto zs = reinterpret_cast<const float4 *>(z_dev);
    auto ds = reinterpret_cast<float4 *>(dx_out_dev);
    if (pool)
        k_bias_pool_silu_bwd<true><<<grid, QN_THREADS, 0, (cudaStream_t)stream>>>(
            gs, zs, reinterpret_cast<const uchar4 *>(idx_dev), n, (int)h, (int)w, c4, ds, dbias_out_dev);
    else
        k_bias_pool_silu_bwd<false><<<grid, QN_THREADS, 0, (cudaStream_t)stream>>>(
            gs, zs, nullptr, n, (int)h, (int)w, c4, ds, dbias_out_dev);
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int pb_qnet_head_fwd(const float *x_dev, int64_t n, int64_t hw, const float *conv_bias_dev,
                     const float *w1_dev, const float *b1_dev, const float *wv_dev, const float *bv_dev,
                     const float *wa_dev, const float *ba_dev, int64_t n_out, int dueling,
                     float *out_dev, int device, void *stream) {
    return pb_qnet_head_fwd_train(x_dev, n, hw, conv_bias_dev, w1_dev, b1_dev, wv_dev, bv_dev, wa_dev,
                                  ba_dev, n_out, dueling, out_dev, nullptr, nullptr, nullptr, device, stream);
}

int pb_qnet_head_fwd_train(const float *x_dev, int64_t n, int64_t hw, const float *conv_bias_dev,
                           const float *w1_dev, const float *b1_dev, const float *wv_dev,
                           const float *bv_dev, const float *wa_dev, const float *ba_dev, int64_t n_out,
                           int dueling, float *out_dev, float *p_out_dev, uint8_t *idx_out_dev,
                           float *u_out_dev, int device, void *stream) {
    return qnet_head_launch(x_dev, n, hw, conv_bias_dev, w1_dev, b1_dev, wv_dev, bv_dev, wa_dev, ba_dev, n_out,
                            dueling, out_dev, p_out_dev, idx_out_dev, u_out_dev, nullptr, nullptr, device, stream);
}

int pb_qnet_head_greedy(const float *x_dev, int64_t n, int64_t hw, const float *conv_bias_dev,
                        const float *w1_dev, const float *b1_dev, const float *wv_dev, const float *bv_dev,
                        const float *wa_dev, const float *ba_dev, int64_t n_out, int dueling,
                        float *q_out_dev, const uint8_t *mask_dev, uint8_t *actions_out_dev, int device,
                        void *stream) {
    if (!mask_dev || !actions_out_dev) return fail(PB_ERR_INVALID_ARG, "pb_qnet_head_greedy: null argument");
    return qnet_head_launch(x_dev, n, hw, conv_bias_dev, w1_dev, b1_dev, wv_dev, bv_dev, wa_dev, ba_dev, n_out,
                            dueling, q_out_dev, nullptr, nullptr, nullptr, mask_dev, actions_out_dev, device,
                            stream);
}

int pb_qnet_head_policy_value(const float *x_dev, int64_t n, int64_t hw, const float *conv_bias_dev,
                              const float *w1_dev, const float *b1_dev, const float *wo_dev,
                              const float *bo_dev, int64_t n_out, const uint8_t *mask_dev, float value_scale,
                              float *out_dev, float *value_out_dev, float *policy_out_dev, int device,
                              void *stream) {
    if (!mask_dev || !value_out_dev || !policy_out_dev || n_out < 2)
        return fail(PB_ERR_INVALID_ARG, "pb_qnet_head_policy_value: bad arguments");
    return qnet_head_launch(x_dev, n, hw, conv_bias_dev, w1_dev, b1_dev, nullptr, nullptr, wo_dev, bo_dev, n_out, 0,
                            out_dev, nullptr, nullptr, nullptr, mask_dev, nullptr, device, stream,
                            value_out_dev, policy_out_dev, value_scale);
}

}  // extern "C"

namespace {
int qnet_head_launch(const float *x_dev, int64_t n, int64_t hw, const float *conv_bias_dev,
                     const float *w1_dev, const float *b1_dev, const float *wv_dev, const float *bv_dev,
                     const float *wa_dev, const float *ba_dev, int64_t n_out, int dueling, float *out_dev,
                     float *p_out_dev, uint8_t *idx_out_dev, float *u_out_dev, const uint8_t *mask_dev,
                     uint8_t *actions_out_dev, int device, void *stream, float *pv_value_dev,
                     float *pv_policy_dev, float value_scale) {
    const bool train = p_out_dev || idx_out_dev || u_out_dev;
    if (train && (!p_out_dev || !idx_out_dev || !u_out_dev || hw > 255))
        return fail(PB_ERR_INVALID_ARG, "pb_qnet_head_fwd_train: p / idx / u outputs come together, map of <= 255 cells");
    if (!x_dev || !conv_bias_dev || !w1_dev || !b1_dev || !wa_dev || !ba_dev || !out_dev || n <= 0 ||
        hw <= 0 || hw > 4096 || n_out <= 0 || n_out > HEAD_MAX_OUT - (dueling ? 1 : 0) ||
        (dueling && (!wv_dev || !bv_dev)) || !aligned16(w1_dev) || !aligned16(x_dev))
        return fail(PB_ERR_INVALID_ARG, "pb_qnet_head_fwd: bad arguments (x float32 [n][hw][128], "
                                        "Linear 128 -> 256, up to 7 output rows)");
    DeviceGuard g(device);
    cudaStream_t s = (cudaStream_t)stream;
    // samples per CTA: enough CTAs to cover the SMs at small batches, W1 amortised at large ones
    const int sms = qnet_sms(device);
#define PB_HEAD_LAUNCH(S, T)                                                                       \
    k_qnet_head_fwd<S, T><<<(unsigned)((n + (S) - 1) / (S)), HEAD_THREADS, 0, s>>>(                \
        x_dev, (int)hw, conv_bias_dev, w1_dev, b1_dev, wv_dev, bv_dev, wa_dev, ba_dev, (int)n_out,  \
        dueling, n, out_dev, p_out_dev, idx_out_dev, u_out_dev, mask_dev, actions_out_dev, pv_value_dev,     \
        pv_policy_dev, value_scale)
    if (train) {
        if (n <= sms)
            PB_HEAD_LAUNCH(1, true);
        else if (n <= 2 * sms)
            PB_HEAD_LAUNCH(2, true);
        else
            PB_HEAD_LAUNCH(4, true);
    } else if (n <= sms) {
        PB_HEAD_LAUNCH(1, false);
    } else if (n <= 2 * sms) {
        PB_HEAD_LAUNCH(2, false);
    } else {
        PB_HEAD_LAUNCH(4, false);
    }
#undef PB_HEAD_LAUNCH
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}
}  // namespace

extern "C" {

int pb_qnet_head_bwd(const float *dq_dev, const float *p_dev, const uint8_t *idx_dev, const float *u_dev,
                     const float *w1_dev, const float *wv_dev, const float *wa_dev, int64_t n, int64_t hw,
                     int64_t n_out, int dueling, float *dx_dev, float *dconv_bias_dev, float *dw1_dev,
                     float *db1_dev, float *dwv_dev, float *dbv_dev, float *dwa_dev, float *dba_dev,
                     float *scratch_dev, int device, void *stream) {
    if (!dq_dev || !p_dev || !idx_dev || !u_dev || !w1_dev || !wa_dev || !dx_dev || !dconv_bias_dev ||
        !dw1_dev || !db1_dev || !dwa_dev || !dba_dev || !scratch_dev || n <= 0 || hw <= 0 || hw > 255 ||
        n_out <= 0 || n_out > HEAD_MAX_OUT - (dueling ? 1 : 0) || (dueling && (!wv_dev || !dwv_dev || !dbv_dev)) ||
        !aligned16(dx_dev))
        return fail(PB_ERR_INVALID_ARG, "pb_qnet_head_bwd: bad arguments");
    DeviceGuard g(device);
    cudaStream_t s = (cudaStream_t)stream;
    // scratch: dp [n][128] | du [n][256] | a [n][128] | f [n][256] | dh [n][8]
    float *dp = scratch_dev, *du = dp + n * HEAD_C, *a = du + n * HEAD_H, *f = a + n * HEAD_C,
          *dh = f + n * HEAD_H;
    const int sms = qnet_sms(device);
#define PB_HEAD_BWD(S)                                                                             \
    k_qnet_head_bwd<S><<<(unsigned)((n + (S) - 1) / (S)), HEAD_THREADS, 0, s>>>(                   \
        dq_dev, p_dev, idx_dev, u_dev, w1_dev, wv_dev, wa_dev, (int)hw, (int)n_out, dueling, n,    \
        dx_dev, dp, du, a, f, dh)
    if (n <= sms)
        PB_HEAD_BWD(1);
    else if (n <= 2 * sms)
        PB_HEAD_BWD(2);
    else
        PB_HEAD_BWD(4);
#undef PB_HEAD_BWD
    PB_CUDA(cudaGetLastError());
    k_qnet_head_wgrad<<<WG_TILES + 2, WG_THREADS, 0, s>>>(du, a, f, dh, dp, n, (int)n_out, dueling, dw1_dev,
                                                         db1_dev, dwv_dev, dbv_dev, dwa_dev, dba_dev,
                                                         dconv_bias_dev);
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int64_t pb_qnet_head_bwd_scratch_floats(int64_t n) { return n * (2 * HEAD_C + 2 * HEAD_H + HEAD_R); }

}  // extern "C"
