// Microbenchmark (not product code): how fast can 65 536 x 59 024 B blocks be streamed to HBM
// from one B200 by different store paths?  Build: nvcc -gencode arch=compute_100a,code=sm_100a
//   A  TMA bulk store of a whole 59 024 B image per warp, W warps/CTA (W = 1..3), ticket
//   B  TMA bulk store split in 17 channel copies
//   C  direct st.global.v4 from registers, env per warp, many warps
//   D  direct st.global.v4, grid-stride over the flat buffer (what torch.fill_ does)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int OBS_BYTES = 59024;
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int CHUNKS>
__global__ void k_bulk(float *out, long long n, unsigned *ticket, int warps) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float4 *img4 = reinterpret_cast<float4 *>(smem + (size_t)warp * OBS_BYTES);
    for (int q = lane; q < OBS_BYTES / 16; q += 32) img4[q] = make_float4(1.f, 2.f, 3.f, 4.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    unsigned next = 0;
    if (lane == 0) next = atomicAdd(ticket, 1u);
    next = __shfl_sync(~0u, next, 0);
    while ((long long)next < n) {
        long long k = next;
        if (lane == 0) {
            next = atomicAdd(ticket, 1u);
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            char *dst = reinterpret_cast<char *>(out) + k * OBS_BYTES;
            constexpr int CB = OBS_BYTES / CHUNKS;
#pragma unroll
            for (int c = 0; c < CHUNKS; c++)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + c * CB),
                             "r"(smem_u32(reinterpret_cast<char *>(img4) + c * CB)), "r"(CB) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        next = __shfl_sync(~0u, next, 0);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
// A + the ingredients of the real encoder, one flag each:
//   1: fence.proxy.async + __syncwarp before every issue   2: rewrite 2 channels (434 STS.128)
//   4: 176 B of global loads per env (the packed state)     8: 4 warps x 55 552 B + shared 3 472 B
//  16: L2 evict_first cache hint on the bulk stores         32: late ticket (taken after wait.read)
//  64: L2 evict_last hint instead (negative control)
template <int FLAGS>
__global__ void k_bulk_x(float *out, long long n, unsigned *ticket, const uint4 *state) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int IMG = (FLAGS & 8) ? 55552 : OBS_BYTES;
    float4 *img4 = reinterpret_cast<float4 *>(smem + (size_t)warp * IMG);
    unsigned char *shared_ch = smem + (size_t)(blockDim.x >> 5) * IMG;
    for (int q = lane; q < IMG / 16; q += 32) img4[q] = make_float4(1.f, 2.f, 3.f, 4.f);
    if (FLAGS & 8) for (int q = threadIdx.x; q < 3472 / 16; q += blockDim.x) reinterpret_cast<float4 *>(shared_ch)[q] = make_float4(1.f, 1.f, 0.f, 0.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    unsigned next = 0;
    if (lane == 0) next = atomicAdd(ticket, 1u);
    next = __shfl_sync(~0u, next, 0);
    float acc = 0.f;
    uint64_t pol = 0;
    if (FLAGS & 16) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    if (FLAGS & 64) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    while ((long long)next < n) {
        long long k = next;
        if (!(FLAGS & 32) && lane == 0) next = atomicAdd(ticket, 1u);
        uint4 s0 = make_uint4(0, 0, 0, 0);
        if (FLAGS & 4) { s0 = state[k * 4 + (lane & 3)]; uint4 s1 = state[(n + k) * 4 + (lane & 3)]; s0.x ^= s1.y; }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        if (FLAGS & 2) {
            const float v = (float)(s0.x & 1) + acc;
#pragma unroll
            for (int it = 0; it < 7; it++) {
                const int q = lane + 32 * it;
                if (q < 217) { img4[q] = make_float4(v, v, v, v); img4[217 * 14 + q] = make_float4(v, 1.f, v, 2.f); }
            }
        }
        if (FLAGS & 1) { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); __syncwarp(); }
        if (lane == 0) {
            char *dst = reinterpret_cast<char *>(out) + k * OBS_BYTES;
            if (FLAGS & (16 | 64)) {
                if (FLAGS & 8) {
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(smem_u32(shared_ch)), "r"(3472), "l"(pol) : "memory");
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst + 3472), "r"(smem_u32(img4)), "r"(55552), "l"(pol) : "memory");
                } else {
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(smem_u32(img4)), "r"(OBS_BYTES), "l"(pol) : "memory");
                }
            } else if (FLAGS & 8) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(shared_ch)), "r"(3472) : "memory");
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + 3472), "r"(smem_u32(img4)), "r"(55552) : "memory");
            } else {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(img4)), "r"(OBS_BYTES) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        acc += (float)(s0.y & 1);
        if ((FLAGS & 32) && lane == 0) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            next = atomicAdd(ticket, 1u);
        }
        next = __shfl_sync(~0u, next, 0);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
// ticket granularity / copy size sensitivity of the bare TMA store stream:
// BATCH envs per ticket, each env written as PIECES equal bulk copies (one ticket per piece
// when PIECES > 1 and BATCH == 1)
template <int BATCH, int PIECES>
__global__ void k_bulk_gran(float *out, long long n, unsigned *ticket) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int PB = OBS_BYTES / PIECES / 16 * 16;   // piece bytes (16 B granules)
    float4 *img4 = reinterpret_cast<float4 *>(smem + (size_t)warp * OBS_BYTES);
    for (int q = lane; q < OBS_BYTES / 16; q += 32) img4[q] = make_float4(1.f, 2.f, 3.f, 4.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    const long long units = n * PIECES;
    unsigned next = 0;
    if (lane == 0) next = atomicAdd(ticket, (unsigned)BATCH);
    next = __shfl_sync(~0u, next, 0);
    while ((long long)next < units) {
        const long long u0 = next;
        if (lane == 0) {
            next = atomicAdd(ticket, (unsigned)BATCH);
            for (int b = 0; b < BATCH && u0 + b < units; b++) {
                const long long u = u0 + b, k = u / PIECES; const int piece = (int)(u % PIECES);
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                char *dst = reinterpret_cast<char *>(out) + k * OBS_BYTES + (long long)piece * PB;
                const int bytes = (piece == PIECES - 1) ? OBS_BYTES - piece * PB : PB;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                             "r"(smem_u32(reinterpret_cast<char *>(img4) + piece * PB)), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        next = __shfl_sync(~0u, next, 0);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__global__ void k_direct_env(float4 *out, long long n, unsigned *ticket) {
    const int lane = threadIdx.x & 31;
    unsigned next = 0;
    if (lane == 0) next = atomicAdd(ticket, 1u);
    next = __shfl_sync(~0u, next, 0);
    const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    while ((long long)next < n) {
        float4 *dst = out + (long long)next * (OBS_BYTES / 16);
        if (lane == 0) next = atomicAdd(ticket, 1u);
#pragma unroll 8
        for (int q = lane; q < OBS_BYTES / 16; q += 32) dst[q] = v;
        next = __shfl_sync(~0u, next, 0);
    }
}
__global__ void k_direct_flat(float4 *out, long long total4) {
    const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x)
        out[i] = v;
}
template <typename F> float timeit(F f, int reps = 8) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9f;
    for (int r = 0; r < reps + 2; r++) {
        cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (r >= 2 && ms < best) best = ms;
    }
    return best;
}
int main() {
    const long long n = 65536; const size_t bytes = (size_t)n * OBS_BYTES;
    float *buf[3]; for (auto &b : buf) cudaMalloc(&b, bytes);
    unsigned *ticket; cudaMalloc(&ticket, 4);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaFuncSetAttribute(k_bulk<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * OBS_BYTES);
    cudaFuncSetAttribute(k_bulk<17>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * OBS_BYTES);
    int it = 0;
    for (int w = 1; w <= 3; w++) {
        float ms = timeit([&] { cudaMemsetAsync(ticket, 0, 4); k_bulk<1><<<sms, 32 * w, w * OBS_BYTES>>>(buf[it++ % 3], n, ticket, w); });
        printf("A bulk 59024B x1   warps/SM=%d : %.3f ms  %.0f GB/s\n", w, ms, bytes / ms / 1e6);
    }
    { float ms = timeit([&] { cudaMemsetAsync(ticket, 0, 4); k_bulk<17><<<sms, 96, 3 * OBS_BYTES>>>(buf[it++ % 3], n, ticket, 3); });
      printf("B bulk 3472B x17   warps/SM=3 : %.3f ms  %.0f GB/s\n", ms, bytes / ms / 1e6); }
    uint4 *state; cudaMalloc(&state, (size_t)n * 8 * 16); cudaMemset(state, 1, (size_t)n * 8 * 16);
#define RUNX(F, W, label) { cudaFuncSetAttribute(k_bulk_x<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024); \
        int smem = ((F) & 8) ? (W) * 55552 + 3472 : (W) * OBS_BYTES; \
        float ms = timeit([&] { cudaMemsetAsync(ticket, 0, 4); k_bulk_x<F><<<sms, 32 * (W), smem>>>(buf[it++ % 3], n, ticket, state); }); \
        printf("X %-44s warps/SM=%d : %.3f ms  %.0f GB/s\n", label, W, ms, bytes / ms / 1e6); }
    RUNX(0, 3, "baseline (same as A)");
    RUNX(1, 3, "+fence.proxy.async per env");
    RUNX(2, 3, "+rewrite 2 channels");
    RUNX(3, 3, "+fence +rewrite");
    RUNX(4, 3, "+state loads");
    RUNX(7, 3, "+fence +rewrite +loads");
    RUNX(8, 4, "split 3472+55552, 4 warps");
    RUNX(15, 4, "split, 4 warps, +fence +rewrite +loads");
    RUNX(15, 3, "split, 3 warps, +fence +rewrite +loads");
    RUNX(15, 2, "split, 2 warps, +fence +rewrite +loads");
    RUNX(16, 3, "baseline + L2 evict_first");
    RUNX(64, 3, "baseline + L2 evict_last");
    RUNX(8 + 16, 4, "split, 4 warps + evict_first");
    RUNX(15 + 32, 4, "split, 4 warps, full chain, late ticket");
    RUNX(15 + 16, 4, "split, 4 warps, full chain, evict_first");
    RUNX(15 + 16 + 32, 4, "split, 4w, full chain, evict_first, late");
    RUNX(15 + 16 + 32, 3, "split, 3w, full chain, evict_first, late");
#define RUNG(B, P, label) { cudaFuncSetAttribute(k_bulk_gran<B, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * OBS_BYTES); \
        float ms = timeit([&] { cudaMemsetAsync(ticket, 0, 4); k_bulk_gran<B, P><<<sms, 96, 3 * OBS_BYTES>>>(buf[it++ % 3], n, ticket); }); \
        printf("G %-44s : %.3f ms  %.0f GB/s\n", label, ms, bytes / ms / 1e6); }
    RUNG(1, 1, "ticket = 1 env (same as A)");
    RUNG(2, 1, "ticket = 2 envs");
    RUNG(4, 1, "ticket = 4 envs");
    RUNG(8, 1, "ticket = 8 envs");
    RUNG(1, 2, "ticket = half env (2 x 29.5 KB)");
    RUNG(1, 4, "ticket = quarter env (4 x 14.7 KB)");
    for (int wps : {8, 16, 32, 64}) {
        float ms = timeit([&] { cudaMemsetAsync(ticket, 0, 4); k_direct_env<<<sms * wps / 8, 256>>>((float4 *)buf[it++ % 3], n, ticket); });
        printf("C direct v4 env/warp warps/SM=%d : %.3f ms  %.0f GB/s\n", wps, ms, bytes / ms / 1e6);
    }
    for (int bps : {4, 8, 16}) {
        float ms = timeit([&] { k_direct_flat<<<sms * bps, 256>>>((float4 *)buf[it++ % 3], (long long)(bytes / 16)); });
        printf("D direct v4 flat  blocks/SM=%d : %.3f ms  %.0f GB/s\n", bps, ms, bytes / ms / 1e6);
    }
    { float ms = timeit([&] { cudaMemsetAsync(buf[it++ % 3], 0, bytes); });
      printf("E cudaMemsetAsync              : %.3f ms  %.0f GB/s\n", ms, bytes / ms / 1e6); }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
