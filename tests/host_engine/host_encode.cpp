// host_encode.cpp -- TEST INFRASTRUCTURE: the observation encoder kernel of the product
// (pacbot-rl_b200/csrc/pb_encode.cuh: k_encode_obs, PacmanGym::obs, env.rs:410-500) compiled for
// the HOST and executed by ONE OS THREAD PER LANE (4 warps x 32 lanes = one CTA), so that
// tests/test_host_encode.py can diff the *same source the GPU executes* -- lane roles, shuffles,
// votes, the work ticket, the un-patching of the previous env's cells -- against the oracle's
// observation without a GPU.  Warp shuffles / votes are a barrier-protected exchange between the
// 32 threads of a warp, __syncthreads a 128-thread barrier, atomics are real atomics; the copies
// that cp.async and the TMA engine do on the GPU are plain memcpy (PB_HOST_EMULATION branch at
// the top of pb_encode.cuh; the GPU build's token stream is unaffected by it).
// NOT a CPU path of the product: nothing under pacbot-rl_b200/ builds, loads or links it.
#include <barrier>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#define PB_HOST_EMULATION 1
#include "host_common.h"

#define __global__
#define __launch_bounds__(...)
#define __shared__
#define __align__(n)

struct float4 {
    float x, y, z, w;
};
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

struct HostDim3 {
    unsigned x, y, z;
};
static thread_local HostDim3 threadIdx, blockIdx;
static HostDim3 blockDim, gridDim;

struct WarpExchange {
    std::barrier<> bar{32};
    uint32_t slot[32];
};
static WarpExchange *g_warp;       // one per warp of the CTA
static std::barrier<> *g_cta_bar;  // all threads of the CTA

static inline uint32_t __shfl_sync(unsigned, uint32_t v, int src_lane) {
    WarpExchange &w = g_warp[threadIdx.x >> 5];
    w.slot[threadIdx.x & 31] = v;
    w.bar.arrive_and_wait();
    const uint32_t r = w.slot[src_lane & 31];
    w.bar.arrive_and_wait();
    return r;
}
static inline bool __all_sync(unsigned, bool pred) {
    WarpExchange &w = g_warp[threadIdx.x >> 5];
    w.slot[threadIdx.x & 31] = pred ? 1u : 0u;
    w.bar.arrive_and_wait();
    bool all = true;
    for (int k = 0; k < 32; k++) all = all && w.slot[k];
    w.bar.arrive_and_wait();
    return all;
}
static inline void __syncwarp() { g_warp[threadIdx.x >> 5].bar.arrive_and_wait(); }
static inline void __syncthreads() { g_cta_bar->arrive_and_wait(); }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicExch(unsigned *p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
static inline void __trap() { abort(); }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fadd_rn(float a, float b) { return a + b; }

namespace pb {
static const MazeTables c_tab = H_TABLES;
}

#include "pb_encode.cuh"

namespace pb {
// the CTA's dynamic shared memory (`extern __shared__ unsigned char smem_raw[]` in the kernel)
alignas(128) unsigned char smem_raw[ENC_SMEM];
}  // namespace pb

using namespace pb;
using namespace host_common;

extern "C" {

// k_encode_obs over `n` envs given as pb_env_state records; out = float32 [n][17][28][31].
// `first` / `count` select a sub-range like pb_encode_obs_range.
int he_encode_obs(const pb_env_state *states, int64_t n, int64_t first, int64_t count, float *out) {
    const long long npad = (n + 31) / 32 * 32;
    std::vector<uint4> core0(npad), core1(npad), ghosts(2 * npad);
    std::vector<uint32_t> pel((size_t)PEL_WORDS * npad, 0u);
    for (int64_t i = 0; i < n; i++) {
        EnvR s;
        uint32_t f[PEL_WORDS];
        from_host(states[i], s, f);
        pack_env(s, core0[i], core1[i], ghosts[i], ghosts[npad + i]);
        for (int w = 0; w < PEL_WORDS; w++) pel[(size_t)w * npad + i] = f[w];
    }
    StateView st{core0.data(), core1.data(), ghosts.data(), pel.data(), npad, n};
    unsigned ticket[2] = {0u, 0u};
    std::vector<WarpExchange> warps(ENC_WARPS);
    std::barrier<> cta_bar(ENC_THREADS);
    g_warp = warps.data();
    g_cta_bar = &cta_bar;
    blockDim = HostDim3{(unsigned)ENC_THREADS, 1, 1};
    gridDim = HostDim3{1, 1, 1};   // one persistent CTA takes every ticket
    std::vector<std::thread> lanes;
    for (unsigned t = 0; t < (unsigned)ENC_THREADS; t++)
        lanes.emplace_back([&, t] {
            threadIdx = HostDim3{t, 0, 0};
            blockIdx = HostDim3{0, 0, 0};
            k_encode_obs(st, out, (long long)OBS_FLOATS, first, count, ticket, nullptr, 0, nullptr);
        });
    for (auto &th : lanes) th.join();
    return (ticket[0] == 0u && ticket[1] == 0u) ? 0 : -1;   // the last warp re-arms the ticket pair
}

}  // extern "C"
