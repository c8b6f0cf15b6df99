// pb_encode.cuh -- k_encode_obs: PacmanGym::obs (/root/reference/pacbot_rs/src/env.rs:410-500)
// for a batch of envs, written to HBM by the TMA engine.
//
// Shape of the kernel (one CTA per SM, persistent, ENC_WARPS = 4 warps, one env and one image per
// warp at a time):
//   * shared memory holds ONE wall-channel template (ch 0, constant for every env, 3 472 B) and
//     FOUR private images of channels 1..16 (55 552 B each): 225 680 B of the 227 KB;
//   * per env a warp rewrites only what changes: channel 1 (pellet values: a nibble of the
//     observation-ordered pellet bitboard becomes one 128-bit shared store), channel 15 (constant
//     fill) and <= 19 single cells (Pac-Man x2, ghosts x8, mode/fright timers, super pellets,
//     fruit, frightened-ghost bonus); the zeros of the sparse channels persist across envs and
//     only the patched cells are cleared again;
//   * ONE lane then issues two cp.async.bulk shared->global copies (SASS UBLKCP): wall template
//     -> channel 0, private image -> channels 1..16.  No register staging, full-line HBM writes;
//   * before an image is rebuilt the warp waits for the bulk group's shared-memory READ only
//     (cp.async.bulk.wait_group.read), so global-write latency is never on the critical path;
//   * envs are handed out by a global atomic ticket (SMs drain to HBM at different rates; a
//     static map left fast SMs idle: ncu sm__cycles_active min/max 0.65, +22 % time), and a warp
//     takes its ticket AS LATE AS POSSIBLE: only once its image is free again (see below).
//
// Why TMA: tools/microbench/store_paths.cu measures, for this exact access pattern (65 536 x
// 59 024 B), bare TMA bulk stores at 7.48 TB/s (0.517 ms) vs 7.15 TB/s for the best
// st.global.v4 variant and 7.34 TB/s for cudaMemset, provided >= 2 copies per SM are in flight
// (1 in flight: 6.8 TB/s).
//
// Why late tickets: HBM write efficiency depends on WHICH addresses are being written at the same
// time.  The 592 warps of the grid each write one 59 KB block; the tighter the window of blocks in
// flight, the better the DRAM pages/channels are used.  The same microbenchmark (variant G) with
// tickets of 1 / 2 / 4 / 8 consecutive envs per warp takes 0.520 / 0.527 / 0.534 / 0.540 ms, and
// in this kernel every scheme that holds a ticket ahead of its write widens the window:
//     ticket + state prefetched two envs ahead (software pipeline)      0.547-0.551 ms
//     tickets in batches of 4, state staged through cp.async            0.569 ms
//     ticket taken when the previous image is handed to the TMA engine  0.542 ms
//     ticket taken when the previous copy has finished reading          0.539 ms  (7.20 TB/s)
//     ... plus the L2 evict_first hint on the bulk stores (this kernel) 0.532 ms  (7.29 TB/s)
// although the late ticket puts the atomic's and the state fetch's latency back on the warp's
// critical path: with 4 warps per SM the TMA engine still has ~2 copies queued on average.
// The evict_first cache hint (createpolicy + cp.async.bulk ... .L2::cache_hint) tells L2 that the
// 3.9 GB write stream has no reuse, so it drains to DRAM in arrival order instead of displacing
// other lines: +1.4 % here (A/B on one box: 0.5319 vs 0.5391 ms); with the hint an earlier ticket
// is again slower (0.5351 ms).  Also measured, equal within noise or slower: 3 x 59 024 B images
// with channel 0 inside, 2 warps x 2 double-buffered images, channel-1 values prepared in
// registers before the wait or not, prefetch.global.L2 of the state of the env 256 / 1 024
// tickets ahead (0.5413 / 0.5427 vs 0.5413 ms: the DRAM reads are only issued earlier, they stay
// scattered through the write stream; see "the last 2 %" below),
// and the shared-template encoder with 16 envs in flight per SM (pb_encode2.cuh, 0.602 ms: five
// copies and <= 18 scattered 4-byte stores per env disturb the write stream more than they
// help).
//
// The last 2 %: state reads.  Two diagnostic builds (profiles/r01_ab_encoder.txt) separated the
// suspects: without the per-env state fetch (every warp re-encodes one fixed state) the kernel
// runs at 0.5211 ms = the bare-TMA rate, with the ticket atomic still on the critical path; with a
// static env->warp map and no atomics it is far slower (0.62-0.64 ms).  So the residual was the
// 176 B of state per env, and six images per SM instead of four (duplicate channels copied twice
// by the TMA engine, 34 KB images) did NOT help (0.5233 vs 0.5205 ms): it is not latency hidden by
// more copies in flight, it is the 11.5 MB of state arriving as ~360 k DRAM READ sectors scattered
// through a 3.9 GB WRITE stream (read/write turnarounds on every HBM channel).  The state was
// written by the step kernel just before and is still in L2 when this kernel starts, so the
// prologue marks those lines evict_last (prefetch.global.L2::evict_last, 5 instructions per thread)
// and the per-env fetches then hit L2: 0.5320 -> 0.5204 ms (7.46 TB/s, 99.5 % of the bare-TMA
// ceiling of 0.517-0.518 ms).
//
// Per-lane roles are written BRANCH FREE (selects + one predicated store per lane).  A first
// version used an if/else chain per lane group with nested conditionals; for particular states
// (tests/test_parity_gpu.py fuzz seed 2) lanes at the end of the chain did not get their store in
// before lane 0 issued the copy, ptxas having dropped every __syncwarp() it considered redundant;
// the mismatch vanished with printf in the kernel.  Hence also warp_arrived(): a vote whose result
// is consumed, which cannot be dropped.
//
// Bit-exactness: every float is either an exactly representable constant computed with IEEE
// division at compile time (10/200, 50/200, ...), or an IEEE division / addition done with
// __fdiv_rn / __fadd_rn in the reference's order (env.rs:441,463,464,468,484).
#pragma once
#include "pb_engine.cuh"

namespace pb {

// defined in pb_kernels.cu before this header is included
// (same translation unit): __constant__ MazeTables c_tab;

constexpr int ENC_WARPS = 4;   // warps per CTA
constexpr int ENC_THREADS = ENC_WARPS * 32;
constexpr int CH_BYTES = CH_FLOATS * 4;                 // 3 472
constexpr int IMG_CH = OBS_CH - 1;                      // channels 1..16 live in the private image
constexpr int IMG_FLOATS = IMG_CH * CH_FLOATS;          // 13 888
constexpr int IMG_BYTES = IMG_FLOATS * 4;               // 55 552
// per-warp staging area for one env's packed state: core0, core1, ghosts x2 (4 x 16 B) followed
// by the 28 pellet words
constexpr int STAGE_WORDS = 48;
constexpr int ENC_SMEM = ENC_WARPS * IMG_BYTES + CH_BYTES + ENC_WARPS * STAGE_WORDS * 4;
constexpr int CH_VEC = CH_FLOATS / 4;                   // 217 float4 per channel
static_assert(ENC_SMEM <= 227 * 1024, "encoder images must fit the 227 KB shared memory");
static_assert(IMG_BYTES % 16 == 0 && CH_BYTES % 16 == 0, "bulk copies move 16 B granules");

#ifdef PB_HOST_EMULATION
// The test suite compiles this kernel for the HOST (one OS thread per lane, tests/README.md) to
// diff it against the oracle without a GPU.  Only the data movement primitives differ there:
// plain copies instead of cp.async / TMA bulk copies, nothing to fence or wait for.  The GPU
// build never defines PB_HOST_EMULATION; its token stream is the #else branch, unchanged.
#ifndef PB_ENC_STORE_HINT
#define PB_ENC_STORE_HINT 1
#endif
#ifndef PB_ENC_PIN_STATE
#define PB_ENC_PIN_STATE 0
#endif
inline void bulk_store_s2g(void *gdst, const void *ssrc, uint32_t bytes) { memcpy(gdst, ssrc, bytes); }
inline uint64_t l2_policy_evict_first() { return 0; }
inline void bulk_store_s2g_hint(void *gdst, const void *ssrc, uint32_t bytes, uint64_t) {
    memcpy(gdst, ssrc, bytes);
}
inline void bulk_commit() {}
template <int N>
inline void bulk_wait_read() {}
inline void bulk_wait_all() {}
inline void fence_proxy_async_smem() {}
inline void cp_async_4(void *sdst, const void *gsrc) { memcpy(sdst, gsrc, 4); }
inline void cp_async_16(void *sdst, const void *gsrc) { memcpy(sdst, gsrc, 16); }
inline void cp_async_commit() {}
inline void cp_async_wait_all() {}
inline void grid_dependency_wait() {}
inline unsigned long long global_timer_ns() { return 0; }
inline void stamp_min(unsigned long long *, unsigned long long) {}   // pb_trace_*: GPU only
inline void stamp_max(unsigned long long *, unsigned long long) {}
#else
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void bulk_store_s2g(void *gdst, const void *ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"(smem_u32(ssrc)), "r"(bytes)
                 : "memory");
}
// PB_ENC_STORE_HINT (default 1): the observation stream is written once and not read by this
// kernel, so its lines are marked evict_first in L2; 0 = plain stores, for A/B runs
// (tools/ab_encoder.sh)
#ifndef PB_ENC_STORE_HINT
#define PB_ENC_STORE_HINT 1
#endif
// PB_ENC_PIN_STATE (default 1): prefetch.global.L2::evict_last over the launch's packed state
// before the first copy is issued; 0 for A/B runs
#ifndef PB_ENC_PIN_STATE
#define PB_ENC_PIN_STATE 1
#endif
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_store_s2g_hint(void *gdst, const void *ssrc, uint32_t bytes, uint64_t pol) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst),
                 "r"(smem_u32(ssrc)), "r"(bytes), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// wait until at most N of this thread's bulk groups are still reading shared memory
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() {
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void cp_async_4(void *sdst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_16(void *sdst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// wait for the preceding grid of the stream (programmatic dependent launch)
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// nanoseconds of the GPU's global timer (pb_trace_*: device timeline of the launches)
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void stamp_min(unsigned long long *p, unsigned long long t) { atomicMin(p, t); }
__device__ __forceinline__ void stamp_max(unsigned long long *p, unsigned long long t) { atomicMax(p, t); }
#endif  // PB_HOST_EMULATION

// One record of the device timeline (pb_trace_*): written by the launches of an engine while
// tracing is on.  begin = earliest CTA start, ready = earliest moment a CTA had its dependencies
// (the preceding grid) resolved, end = latest warp exit; all in %globaltimer nanoseconds.
struct TraceRec {
    unsigned long long begin, ready, end, kind;
};

__device__ __forceinline__ bool on_board(int row, int col) {
    return (unsigned)row < (unsigned)ROWS && (unsigned)col < (unsigned)COLS;
}
__device__ __forceinline__ int obs_cell(int row, int col) {  // env.rs:124: [col][30 - row]
    return col * ROWS + (ROWS - 1 - row);
}
// Warp-wide rendezvous that ptxas cannot drop.  `__syncwarp()` (bar.warp.sync) is removed by
// ptxas wherever its convergence analysis says the warp is already converged (the SASS of this
// kernel had a single WARPSYNC left out of five), and with the divergent per-lane roles below that
// was observed to let lane 0 hand an image to the TMA engine before slower lanes had written
// their cells (state-dependent, deterministic mismatches in tests/test_parity_gpu.py fuzz seed 2;
// gone with printf in the kernel).  A vote.sync whose result is consumed has to wait for all 32
// lanes.  `v` must be a value the compiler cannot fold (the prefetched ticket).
// The vote is a rendezvous, not a memory fence: CUDA only documents __syncwarp() (bar.warp.sync)
// as ordering memory accesses among the participating lanes, so the barrier comes first (it is
// the memory ordering between one lane's shared-memory stores / cp.async data and another lane's
// reads or lane 0's bulk copy) and the consumed vote keeps ptxas from dropping the rendezvous.
__device__ __forceinline__ bool warp_arrived(unsigned int v) {
    __syncwarp();
    return __all_sync(0xffffffffu, v != 0xfffffffeu);
}
// offset of (channel, cell) inside the private image (channel 0 is not part of it)
__device__ __forceinline__ int img_off(int ch, int cell) { return (ch - 1) * CH_FLOATS + cell; }

// The two fill values of an env's observation: a plain pellet's channel-1 value (env.rs:427-441:
// (points + LAST_REWARD if this is the last pellet) as f32 / 200.0, IEEE-rounded compile-time
// constants) and channel 15, ticks per step / update period (env.rs:484, one IEEE division).
__device__ __forceinline__ float obs_pellet_value(const EnvR &s) {
    return s.num_pellets == 1 ? (float)(PELLET_POINTS + LAST_REWARD) / 200.0f : (float)PELLET_POINTS / 200.0f;
}
__device__ __forceinline__ float obs_speed_value(const EnvR &s) {
    return __fdiv_rn((float)s.tps, (float)s.period);
}

// The single cells of one env's observation as per-lane roles: pure register work, branch free.
// Called by all 32 lanes of a warp; lane l < PEL_WORDS holds word l of the observation-ordered
// pellet bitboard in F.  Every lane ends up with at most ONE store into the image of channels
// 1..16 (`off` floats from the start of channel 1); `unpatch`: the cell belongs to a sparse
// channel and has to be cleared again before the image is reused for another env.  Also hands
// back the two fill values of the env: pv (a plain pellet's channel-1 value) and sp (channel 15).
struct CellRole {
    bool do_store, unpatch;
    int off;
    float val, pv, sp;
};
__device__ __forceinline__ CellRole cell_role(const EnvR &s, const uint32_t F, const int lane) {
    // Every lane ends up with at most ONE store (offset, value, needs-un-patch):
    //   lane 0,1    Pac-Man, ch 2 / 3                       (env.rs:452-457)
    //   lane 4..7   ghost g = lane & 3, ch 4+g              (env.rs:461)
    //   lane 8..11  ghost g, ch 8+g                         (env.rs:476-480)
    //   lane 12..15 ghost g timer, ch 12 / 13 / 14          (env.rs:462-469)
    //   lane 16..19 ghost g frightened bonus, ch 1          (env.rs:464)
    //   lane 20..23 super pellet k = lane & 3, ch 16        (env.rs:487-497)
    //   lane 24..27 super pellet k value, ch 1              (env.rs:428-432)
    //   lane 28     fruit value, ch 1                       (env.rs:433-437)
    const int grp = lane >> 2, gi = lane & 3;
    int grow = s.g[0].row, gcol = s.g[0].col, gfr = s.g[0].fright;
#pragma unroll
    for (int g = 1; g < 4; g++) {
        grow = (g == gi) ? s.g[g].row : grow;
        gcol = (g == gi) ? s.g[g].col : gcol;
        gfr = (g == gi) ? s.g[g].fright : gfr;
    }
    const bool g_on = on_board(grow, gcol);
    const int g_cell = g_on ? obs_cell(grow, gcol) : 0;
    const int sup_cell = obs_cell(gi < 2 ? 3 : 23, (gi & 1) ? 26 : 1);
    const bool pac_on = on_board(s.prow, s.pcol);
    const int pac_cell = pac_on ? obs_cell(s.prow, s.pcol) : 0;
    // the cell whose pellet bit / frightened occupants this lane needs (ch-1 finalisers)
    const int cell = (grp == 4) ? g_cell : ((grp == 5 || grp == 6) ? sup_cell
                                           : obs_cell(FRUIT_ROW, FRUIT_COL));
    const uint32_t cell_word = __shfl_sync(0xffffffffu, F, cell >> 5);
    const bool cell_pellet = (cell_word >> (cell & 31)) & 1u;
    bool fright_here = false, earlier_fright_here = false, later_same_timer = false;
    const int my_timer_ch = gfr > 0 ? 14 : (s.mode == CHASE ? 13 : 12);
    float bonus_sum_v = 0.0f;  // filled below for the bonus lanes
    {
        // base value under a frightened ghost (env.rs:427-441), then one IEEE add per
        // frightened ghost on that cell, in ghost order, starting with this lane's ghost
        const bool last_pellet0 = s.num_pellets == 1;
        const float pv0 = last_pellet0 ? (float)(PELLET_POINTS + LAST_REWARD) / 200.0f
                                       : (float)PELLET_POINTS / 200.0f;
        const float sv0 = last_pellet0 ? (float)(SUPER_PELLET_POINTS + LAST_REWARD) / 200.0f
                                       : (float)SUPER_PELLET_POINTS / 200.0f;
        const bool is_sup = (grow == 3 || grow == 23) && (gcol == 1 || gcol == 26);
        const bool is_fruit = s.fruit_steps > 0 && grow == FRUIT_ROW && gcol == FRUIT_COL;
        float v = cell_pellet ? (is_sup ? sv0 : pv0)
                              : (is_fruit ? (float)FRUIT_POINTS / 200.0f : 0.0f);
        const float bonus = (float)(1 << (s.combo & 31));
#pragma unroll
        for (int g = 0; g < 4; g++) {
            const bool on = on_board(s.g[g].row, s.g[g].col);
            const bool f_here = s.g[g].fright > 0 && on && obs_cell(s.g[g].row, s.g[g].col) == cell;
            fright_here |= f_here;
            earlier_fright_here |= f_here && g < gi;
            const int gch = s.g[g].fright > 0 ? 14 : (s.mode == CHASE ? 13 : 12);
            later_same_timer |= g > gi && on && s.g[g].row == grow && s.g[g].col == gcol &&
                                gch == my_timer_ch;
            const bool add = g >= gi && s.g[g].fright > 0 && s.g[g].row == grow && s.g[g].col == gcol;
            v = add ? __fadd_rn(v, bonus) : v;
        }
        bonus_sum_v = v;
    }
    // env.rs:427-441: (points + LAST_REWARD if this is the last pellet) as f32 / 200.0; the
    // four possible quotients are IEEE-rounded compile-time constants
    const bool last_pellet = s.num_pellets == 1;
    const float pv = obs_pellet_value(s);
    const float sv = last_pellet ? (float)(SUPER_PELLET_POINTS + LAST_REWARD) / 200.0f
                                 : (float)SUPER_PELLET_POINTS / 200.0f;
    const float sp = obs_speed_value(s);
    const float timer_v = gfr > 0 ? c_tab.fright_q[gfr & 0xff] : c_tab.mode_q[s.mode_steps & 0xff];
    // one store per lane
    bool do_store, unpatch;
    int off;
    float val;
    if (grp == 0) {             // Pac-Man (lanes 0, 1; lanes 2, 3 idle)
        do_store = pac_on && lane < 2; unpatch = true; off = img_off(2 + lane, pac_cell); val = 1.0f;
    } else if (grp <= 2) {      // ghost position, ch 4+g (lanes 4..7) and 8+g (lanes 8..11)
        do_store = g_on; unpatch = true; off = img_off(lane, g_cell); val = 1.0f;
    } else if (grp == 3) {      // timers; "last ghost wins" on a shared cell of one channel
        do_store = g_on && !later_same_timer; unpatch = true; off = img_off(my_timer_ch, g_cell); val = timer_v;
    } else if (grp == 4) {      // bonus: the FIRST frightened ghost on a cell finalises ch 1 there
        do_store = g_on && gfr > 0 && !earlier_fright_here; unpatch = false; off = g_cell; val = bonus_sum_v;
    } else if (grp == 5) {      // super pellet still there: ch 16
        do_store = cell_pellet; unpatch = true; off = img_off(16, sup_cell); val = 1.0f;
    } else if (grp == 6) {      // ... and its ch-1 value unless a frightened ghost owns the cell
        do_store = cell_pellet && !fright_here; unpatch = false; off = sup_cell; val = sv;
    } else {                    // fruit (lane 28): only where there is no pellet
        do_store = lane == 28 && s.fruit_steps > 0 && !cell_pellet && !fright_here;
        unpatch = false; off = cell; val = (float)FRUIT_POINTS / 200.0f;
    }
    return CellRole{do_store, unpatch, off, val, pv, sp};
}

__global__ void __launch_bounds__(ENC_THREADS, 1)
k_encode_obs(StateView st, float *__restrict__ out,
             long long env_stride, long long env_first, long long env_count,
             unsigned int *__restrict__ ticket, const long long *__restrict__ slot_index,
             long long slot_stride, TraceRec *__restrict__ trace) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (trace && threadIdx.x == 0) stamp_min(&trace->begin, global_timer_ns());
    float *img = reinterpret_cast<float *>(smem_raw + (size_t)warp * IMG_BYTES);
    float4 *img4 = reinterpret_cast<float4 *>(img);
    float *wall_ch = reinterpret_cast<float *>(smem_raw + (size_t)ENC_WARPS * IMG_BYTES);
    uint32_t *stage = reinterpret_cast<uint32_t *>(smem_raw + (size_t)ENC_WARPS * IMG_BYTES + CH_BYTES) +
                      warp * STAGE_WORDS;

    // ---- one-time init: wall template (env.rs:426), zeroed private image ----
    for (int q = lane; q < IMG_FLOATS / 4; q += 32) img4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = threadIdx.x; c < CH_FLOATS; c += blockDim.x) {
        const int col = c / ROWS, row = ROWS - 1 - (c % ROWS);
        wall_ch[c] = ((c_tab.walls[row] >> col) & 1u) ? 1.0f : 0.0f;
    }
    fence_proxy_async_smem();
    __syncthreads();  // the template is read by every warp's bulk copies
    // everything above overlapped the previous kernel (programmatic dependent launch); from here
    // on the env state and the ticket are read, so wait for the preceding grid to finish
    grid_dependency_wait();
    if (trace && threadIdx.x == 0) stamp_min(&trace->ready, global_timer_ns());
    // ring-slot indirection: the slot number lives in device memory so that a CUDA graph can be
    // replayed while the replay ring advances.  Read AFTER the wait: under programmatic dependent
    // launch nothing written by the preceding grid is visible before it.
    if (slot_index) out += *reinterpret_cast<const volatile long long *>(slot_index) * slot_stride;

#if PB_ENC_PIN_STATE
    // The packed state of this launch (176 B per env, 11.5 MB for 65 536 envs) was written by the
    // step kernel a moment ago and is (mostly) still in L2.  Mark its lines evict_last before the
    // 3.9 GB write stream starts, so that the per-env state fetches below hit L2 instead of
    // becoming DRAM reads scattered through the write stream (see the header).
    {
        const long long tid = (long long)blockIdx.x * ENC_THREADS + threadIdx.x;
        const long long nthr = (long long)gridDim.x * ENC_THREADS;
        auto pin = [](const void *p) { asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(p)); };
        for (long long e = tid * 8; e < env_count; e += nthr * 8) {  // 16 B records: 8 per 128 B line
            pin(st.core0 + env_first + e);
            pin(st.core1 + env_first + e);
            pin(st.ghosts + env_first + e);
            pin(st.ghosts + st.npad + env_first + e);
        }
        for (long long e = tid * 32; e < env_count; e += nthr * 32)  // 4 B words: 32 per line
            for (int w = 0; w < PEL_WORDS; w++) pin(st.pel + (long long)w * st.npad + env_first + e);
    }
#endif

    // image offset this lane patched for the previous env (sparse channels only)
    int patched = -1;

    auto next_ticket = [&]() -> unsigned int {
        unsigned int t = 0;
        if (lane == 0) t = atomicAdd(ticket, 1u);
        return __shfl_sync(0xffffffffu, t, 0);
    };
    unsigned int cur = next_ticket();
    while ((long long)cur < env_count) {
        const long long k = cur;
        // ---- this env's packed state (176 B): every lane fetches one piece straight into the
        // warp's staging area (LDGSTS: 28 pellet words + 4 x 16 B), then all lanes read it back;
        // one memory round trip, no shuffles to spread the 16-byte records ----
        {
            const long long env = env_first + k;
            const int w = lane - PEL_WORDS;
            const void *src = (lane < PEL_WORDS) ? (const void *)(st.pel + (long long)lane * st.npad + env)
                              : (w == 0) ? (const void *)(st.core0 + env)
                              : (w == 1) ? (const void *)(st.core1 + env)
                              : (w == 2) ? (const void *)(st.ghosts + env)
                                         : (const void *)(st.ghosts + st.npad + env);
            if (lane < PEL_WORDS) cp_async_4(stage + 16 + lane, src);
            else cp_async_16(stage + 4 * w, src);
            cp_async_commit();
            cp_async_wait_all();
        }
        if (!warp_arrived((unsigned int)k)) __trap();  // every lane's piece has landed
        const uint32_t F = (lane < PEL_WORDS) ? stage[16 + lane] : 0u;
        const uint4 c0 = *reinterpret_cast<const uint4 *>(stage);
        const uint4 c1 = *reinterpret_cast<const uint4 *>(stage + 4);
        const uint4 ga = *reinterpret_cast<const uint4 *>(stage + 8);
        const uint4 gb = *reinterpret_cast<const uint4 *>(stage + 12);
        EnvR s;
        unpack_env(c0, c1, ga, gb, s);

        // ---- per-lane role for the single cells (cell_role above): at most one store per lane
        const CellRole role = cell_role(s, F, lane);
        const bool do_store = role.do_store, unpatch = role.unpatch;
        const int off = role.off;
        const float val = role.val, pv = role.pv, sp = role.sp;

        // ---- rebuild the image (it is free: the ticket was only taken after the previous copy
        // had finished reading it).  Clear last env's single cell, then channel 1 base values (a
        // nibble of the pellet bitboard -> one 128-bit store) and the channel 15 fill ----
        if (patched >= 0) img[patched] = 0.0f;
#pragma unroll
        for (int it = 0; it < 7; it++) {
            const int q = lane + 32 * it;  // float4 index inside a channel
            const uint32_t w = __shfl_sync(0xffffffffu, F, (lane >> 3) + 4 * it);
            if (q < CH_VEC) {
                const uint32_t nib = w >> (4 * (lane & 7));
                float4 v;
                v.x = (nib & 1u) ? pv : 0.0f;
                v.y = (nib & 2u) ? pv : 0.0f;
                v.z = (nib & 4u) ? pv : 0.0f;
                v.w = (nib & 8u) ? pv : 0.0f;
                img4[q] = v;                                             // channel 1
                img4[CH_VEC * 14 + q] = make_float4(sp, sp, sp, sp);     // channel 15
            }
        }
        // base values and un-patching land before the single-cell writers
        if (!warp_arrived((unsigned int)k)) __trap();

        // ---- single cells: one predicated store per lane (disjoint cells by construction) ----
        if (do_store) img[off] = val;
        patched = (do_store && unpatch) ? off : -1;

        // ---- hand channel 0 (template) + channels 1..16 (image) to the TMA engine ----
        fence_proxy_async_smem();
        const bool all_written = warp_arrived((unsigned int)k);  // every lane's cells are in the image
        if (lane == 0 && all_written) {
            char *dst = reinterpret_cast<char *>(out + k * env_stride);
#if PB_ENC_STORE_HINT
            const uint64_t pol = l2_policy_evict_first();
            bulk_store_s2g_hint(dst, wall_ch, CH_BYTES, pol);
            bulk_store_s2g_hint(dst + CH_BYTES, img, IMG_BYTES, pol);
#else
            bulk_store_s2g(dst, wall_ch, CH_BYTES);
            bulk_store_s2g(dst + CH_BYTES, img, IMG_BYTES);
#endif
            bulk_commit();
        }
        // ---- the image is free once the bulk copies have finished READING it; only then take
        // the next ticket, so that the blocks being written at any one time stay close together
        // in memory (see the header: late tickets are worth 2 % of HBM write bandwidth) ----
        if (lane == 0) bulk_wait_read<0>();
        cur = next_ticket();
    }
    if (lane == 0) {
        bulk_wait_all();
        if (trace) stamp_max(&trace->end, global_timer_ns());
        // ticket[1] counts finished warps; the last one re-arms {ticket, count} for the next launch
        const unsigned int finished = atomicAdd(ticket + 1, 1u);
        if (finished == gridDim.x * ENC_WARPS - 1) {
            atomicExch(ticket, 0u);
            atomicExch(ticket + 1, 0u);
        }
    }
}

}  // namespace pb
