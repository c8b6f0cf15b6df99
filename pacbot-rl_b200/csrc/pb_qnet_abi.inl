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

template <int S>
__global__ void __launch_bounds__(HEAD_THREADS)
k_qnet_head_fwd(const float *__restrict__ x, int hw, const float *__restrict__ conv_bias,
                const float *__restrict__ w1, const float *__restrict__ b1,
                const float *__restrict__ wv, const float *__restrict__ bv,
                const float *__restrict__ wa, const float *__restrict__ ba, int n_out, int dueling,
                long long n, float *__restrict__ out) {
    __shared__ __align__(16) float s_a[S][HEAD_C];
    __shared__ __align__(16) float s_m[S][HEAD_PGROUPS][HEAD_C];
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
            if (i < n) {
                const float4 *src = reinterpret_cast<const float4 *>(x + i * hw * HEAD_C) + c4;
#pragma unroll 4
                for (int p = g; p < hw; p += HEAD_PGROUPS) {
                    const float4 v = __ldg(src + (long long)p * (HEAD_C / 4));
                    if (v.x > m.x || v.x != v.x) m.x = v.x;
                    if (v.y > m.y || v.y != v.y) m.y = v.y;
                    if (v.z > m.z || v.z != v.z) m.z = v.z;
                    if (v.w > m.w || v.w != v.w) m.w = v.w;
                }
            }
            *reinterpret_cast<float4 *>(&s_m[smp][g][4 * c4]) = m;
        }
        __syncthreads();
        for (int t = threadIdx.x; t < S * HEAD_C; t += HEAD_THREADS) {
            const int smp = t / HEAD_C, c = t % HEAD_C;
            float best = s_m[smp][0][c];
#pragma unroll
            for (int k = 1; k < HEAD_PGROUPS; k++) {
                const float v = s_m[smp][k][c];
                if (best == best && (v > best || v != v)) best = v;
            }
            s_a[smp][c] = (n0 + smp < n) ? silu_f(best + __ldg(conv_bias + c)) : 0.f;
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
            if (lane < HEAD_RB) s_f[smp][row] = silu_f(v + s_b1[row]);
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
    if ((int)threadIdx.x < S * n_out) {
        const int smp = threadIdx.x / n_out, k = threadIdx.x % n_out;
        const long long i = n0 + smp;
        if (i < n) {
            float q;
            if (dueling) {   // value - adv.mean(dim=1, keepdim=True) + adv  (models.py:108-110)
                float sum = 0.f;
                for (int t = 0; t < n_out; t++) sum += s_h[smp][1 + t];
                q = (s_h[smp][0] - sum / (float)n_out) + s_h[smp][1 + k];
            } else {
                q = s_h[smp][k];
            }
            out[i * n_out + k] = q;
        }
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
    if (!x_dev || !conv_bias_dev || !w1_dev || !b1_dev || !wa_dev || !ba_dev || !out_dev || n <= 0 ||
        hw <= 0 || hw > 4096 || n_out <= 0 || n_out > HEAD_MAX_OUT - (dueling ? 1 : 0) ||
        (dueling && (!wv_dev || !bv_dev)) || !aligned16(w1_dev) || !aligned16(x_dev))
        return fail(PB_ERR_INVALID_ARG, "pb_qnet_head_fwd: bad arguments (x float32 [n][hw][128], "
                                        "Linear 128 -> 256, up to 7 output rows)");
    DeviceGuard g(device);
    cudaStream_t s = (cudaStream_t)stream;
    // samples per CTA: enough CTAs to cover the SMs at small batches, W1 amortised at large ones
    const int sms = qnet_sms(device);
#define PB_HEAD_LAUNCH(S)                                                                          \
    k_qnet_head_fwd<S><<<(unsigned)((n + (S) - 1) / (S)), HEAD_THREADS, 0, s>>>(                   \
        x_dev, (int)hw, conv_bias_dev, w1_dev, b1_dev, wv_dev, bv_dev, wa_dev, ba_dev, (int)n_out,  \
        dueling, n, out_dev)
    if (n <= sms)
        PB_HEAD_LAUNCH(1);
    else if (n <= 2 * sms)
        PB_HEAD_LAUNCH(2);
    else
        PB_HEAD_LAUNCH(4);
#undef PB_HEAD_LAUNCH
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

}  // extern "C"
