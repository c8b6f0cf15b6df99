// pb_encode.cuh -- k_encode_obs: PacmanGym::obs (/root/reference/pacbot_rs/src/env.rs:410-500)
// for a batch of envs, written to HBM by the TMA engine.
//
// Shape of the kernel (one CTA per SM, persistent, ENC_WARPS = 4 warps, one env per warp at a
// time, ENC_BUFS = 1 image per warp):
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
//     static map left fast SMs idle: ncu sm__cycles_active min/max 0.65, +22 % time); ticket and
//     packed state are fetched one to two envs ahead of their use.
//
// Why TMA: tools/microbench/store_paths.cu measures, for this exact access pattern (65 536 x
// 59 024 B), bare TMA bulk stores at 7.48 TB/s (0.517 ms) vs 7.15 TB/s for the best
// st.global.v4 variant and 7.34 TB/s for cudaMemset, provided >= 2 copies per SM are in flight
// (1 in flight: 6.8 TB/s).  What the real encoder adds on top is the dependent chain
// "image free -> state -> rebuild -> fence.proxy.async -> issue": the same microbenchmark with that
// chain (loads + 434 dependent STS.128 + fence) drops to 7.2-7.25 TB/s, and this kernel reaches
// 7.05-7.1 TB/s (0.547-0.551 ms).  Tried and measured equal within noise (0.547-0.561 ms):
// 3 x 59 024 B images with channel 0 inside, 4 x 1 images, 2 warps x 2 double-buffered images,
// with and without the two-deep software pipeline.
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
constexpr int ENC_BUFS = 1;    // images per warp (2 = double buffering; measured no faster)
constexpr int ENC_THREADS = ENC_WARPS * 32;
constexpr int CH_BYTES = CH_FLOATS * 4;                 // 3 472
constexpr int IMG_CH = OBS_CH - 1;                      // channels 1..16 live in the private image
constexpr int IMG_FLOATS = IMG_CH * CH_FLOATS;          // 13 888
constexpr int IMG_BYTES = IMG_FLOATS * 4;               // 55 552
constexpr int ENC_SMEM = ENC_WARPS * ENC_BUFS * IMG_BYTES + CH_BYTES;  // 225 680
constexpr int CH_VEC = CH_FLOATS / 4;                   // 217 float4 per channel
static_assert(ENC_SMEM <= 227 * 1024, "encoder images must fit the 227 KB shared memory");
static_assert(IMG_BYTES % 16 == 0 && CH_BYTES % 16 == 0, "bulk copies move 16 B granules");

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void bulk_store_s2g(void *gdst, const void *ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"(smem_u32(ssrc)), "r"(bytes)
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
__device__ __forceinline__ bool warp_arrived(unsigned int v) {
    return __all_sync(0xffffffffu, v != 0xfffffffeu);
}
// offset of (channel, cell) inside the private image (channel 0 is not part of it)
__device__ __forceinline__ int img_off(int ch, int cell) { return (ch - 1) * CH_FLOATS + cell; }

__global__ void __launch_bounds__(ENC_THREADS, 1)
k_encode_obs(StateView st, float *__restrict__ out,
             long long env_stride, long long env_first, long long env_count,
             unsigned int *__restrict__ ticket, const long long *__restrict__ slot_index,
             long long slot_stride) {
    // ring-slot indirection: the slot number lives in device memory so that a CUDA graph can be
    // replayed while the replay ring advances
    if (slot_index) out += *slot_index * slot_stride;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *wall_ch = reinterpret_cast<float *>(smem_raw + (size_t)ENC_WARPS * ENC_BUFS * IMG_BYTES);
    float *img_base = reinterpret_cast<float *>(smem_raw + (size_t)warp * ENC_BUFS * IMG_BYTES);

    // ---- one-time init: wall template (env.rs:426), zeroed private images ----
    for (int q = lane; q < ENC_BUFS * IMG_FLOATS / 4; q += 32)
        reinterpret_cast<float4 *>(img_base)[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = threadIdx.x; c < CH_FLOATS; c += blockDim.x) {
        const int col = c / ROWS, row = ROWS - 1 - (c % ROWS);
        wall_ch[c] = ((c_tab.walls[row] >> col) & 1u) ? 1.0f : 0.0f;
    }
    fence_proxy_async_smem();
    __syncthreads();  // the template is read by every warp's bulk copies
    // everything above overlapped the previous kernel (programmatic dependent launch); from here
    // on the env state and the ticket are read, so wait for the preceding grid to finish
    asm volatile("griddepcontrol.wait;" ::: "memory");

    // image offset this lane patched in buffer 0 / 1 for an earlier env (sparse channels only)
    int patched0 = -1, patched1 = -1;
    int buf = 0;

    // Software pipeline, two envs deep: while env `cur` is encoded, the packed state of env `nxt`
    // is already being loaded and the ticket after that is being fetched, so neither the atomic's
    // nor the loads' latency sits on the per-env critical path of the warp.
    // `pend` (lane 0 only) is a ticket whose atomic is still in flight: it is requested right
    // AFTER a bulk copy is issued and first read one whole env later, because fence.proxy.async
    // (MEMBAR.ALL.CTA in SASS) makes a lane wait for all of its outstanding memory operations --
    // an atomic requested before the fence would put its L2 round trip on the critical path.
    unsigned int cur = 0, nxt = 0, pend = 0;
    if (lane == 0) {
        cur = atomicAdd(ticket, 1u);
        nxt = atomicAdd(ticket, 1u);
        pend = atomicAdd(ticket, 1u);
    }
    cur = __shfl_sync(0xffffffffu, cur, 0);
    nxt = __shfl_sync(0xffffffffu, nxt, 0);
    uint32_t F = 0u;
    uint4 c0 = make_uint4(0, 0, 0, 0), c1 = c0, ga = c0, gb = c0;
    if ((long long)cur < env_count) {
        const long long i = env_first + cur;
        F = (lane < PEL_WORDS) ? st.pel[(long long)lane * st.npad + i] : 0u;
        c0 = st.core0[i];
        c1 = st.core1[i];
        ga = st.ghosts[i];
        gb = st.ghosts[st.npad + i];
    }
    while ((long long)cur < env_count) {
        const long long k = cur;
        float *img = img_base + buf * IMG_FLOATS;
        float4 *img4 = reinterpret_cast<float4 *>(img);
        // ---- prefetch the state of the next env ----
        uint32_t F_n = 0u;
        uint4 c0_n = make_uint4(0, 0, 0, 0), c1_n = c0_n, ga_n = c0_n, gb_n = c0_n;
        if ((long long)nxt < env_count) {
            const long long i = env_first + nxt;
            F_n = (lane < PEL_WORDS) ? st.pel[(long long)lane * st.npad + i] : 0u;
            c0_n = st.core0[i];
            c1_n = st.core1[i];
            ga_n = st.ghosts[i];
            gb_n = st.ghosts[st.npad + i];
        }
        EnvR s;
        unpack_env(c0, c1, ga, gb, s);

        // ---- per-lane role for the single cells: pure register work, branch free, done before
        // the wait.  Every lane ends up with at most ONE store (offset, value, needs-un-patch):
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
        const float pv = last_pellet ? (float)(PELLET_POINTS + LAST_REWARD) / 200.0f
                                     : (float)PELLET_POINTS / 200.0f;
        const float sv = last_pellet ? (float)(SUPER_PELLET_POINTS + LAST_REWARD) / 200.0f
                                     : (float)SUPER_PELLET_POINTS / 200.0f;
        const float sp = __fdiv_rn((float)s.tps, (float)s.period);      // env.rs:484
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

        // ---- channel-1 vectors in registers: everything that can be computed without touching
        // the image happens BEFORE the wait, so that the chain "image free -> stores -> fence ->
        // issue" is as short as possible (that chain is what keeps the encoder below the bare
        // TMA store rate, see the header) ----
        float4 ch1_v[7];
#pragma unroll
        for (int it = 0; it < 7; it++) {
            const uint32_t w = __shfl_sync(0xffffffffu, F, (lane >> 3) + 4 * it);
            const uint32_t nib = w >> (4 * (lane & 7));
            ch1_v[it].x = (nib & 1u) ? pv : 0.0f;
            ch1_v[it].y = (nib & 2u) ? pv : 0.0f;
            ch1_v[it].z = (nib & 4u) ? pv : 0.0f;
            ch1_v[it].w = (nib & 8u) ? pv : 0.0f;
        }

        // ---- the image is free once the previous bulk copies have finished READING it ----
        // (the copy of the OTHER buffer, issued for the previous env, stays in flight: this is what
        // keeps the TMA engine busy while this warp rebuilds an image)
        if (lane == 0) bulk_wait_read<ENC_BUFS - 1>();
        if (!warp_arrived((unsigned int)k)) __trap();  // nobody touches the image before lane 0's wait is over
        {
            const int patched = buf ? patched1 : patched0;
            if (patched >= 0) img[patched] = 0.0f;
        }

        // ---- channel 1 base values and channel 15 fill: 2 x 7 128-bit stores per lane (the
        // values were prepared in registers before the wait) ----
#pragma unroll
        for (int it = 0; it < 7; it++) {
            const int q = lane + 32 * it;  // float4 index inside a channel
            if (q < CH_VEC) {
                img4[q] = ch1_v[it];                                     // channel 1
                img4[CH_VEC * 14 + q] = make_float4(sp, sp, sp, sp);     // channel 15
            }
        }
        // base values and un-patching land before the single-cell writers
        if (!warp_arrived((unsigned int)k)) __trap();

        // ---- single cells: one predicated store per lane (disjoint cells by construction) ----
        if (do_store) img[off] = val;
        {
            const int patched = (do_store && unpatch) ? off : -1;
            patched0 = buf ? patched0 : patched;
            patched1 = buf ? patched : patched1;
        }

        // ---- hand channel 0 (template) + channels 1..16 (image) to the TMA engine ----
        fence_proxy_async_smem();
        const bool all_written = warp_arrived((unsigned int)k);  // every lane's cells are in the image
        if (lane == 0 && all_written) {
            char *dst = reinterpret_cast<char *>(out + k * env_stride);
            bulk_store_s2g(dst, wall_ch, CH_BYTES);
            bulk_store_s2g(dst + CH_BYTES, img, IMG_BYTES);
            bulk_commit();
        }
        buf = (buf + 1) % ENC_BUFS;
        cur = nxt;
        nxt = __shfl_sync(0xffffffffu, pend, 0);          // requested one env ago: long since back
        if (lane == 0) pend = atomicAdd(ticket, 1u);     // in flight during the next env
        F = F_n;
        c0 = c0_n;
        c1 = c1_n;
        ga = ga_n;
        gb = gb_n;
    }
    if (lane == 0) {
        bulk_wait_all();
        // ticket[1] counts finished warps; the last one re-arms {ticket, count} for the next launch
        const unsigned int finished = atomicAdd(ticket + 1, 1u);
        if (finished == gridDim.x * ENC_WARPS - 1) {
            atomicExch(ticket, 0u);
            atomicExch(ticket + 1, 0u);
        }
    }
}

}  // namespace pb
