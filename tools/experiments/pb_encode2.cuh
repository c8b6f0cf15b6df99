// pb_encode2.cuh -- k_encode_obs_tpl: the observation encoder built on SHARED TEMPLATES.
//
// Same output as k_encode_obs (pb_encode.cuh), different shape.  Of the 17 channels of an
// observation only two are dense and env specific (ch 1 pellet values, ch 15 speed); channel 0 is
// the same for every env and channels 2..14 and 16 are zero except for <= 18 single cells.  So:
//   * one CTA-wide zero region (13 channels, 45 136 B) and one wall template (3 472 B) live in
//     shared memory ONCE; a warp's private footprint is just ch 1 + ch 15 (6 944 B);
//   * per env the warp issues five TMA bulk stores (wall, ch 1, zeros -> ch 2..14, ch 15,
//     zeros -> ch 16: together the full 59 024 B), waits for their completion and then drops the
//     <= 18 single cells into the freshly written zeros with plain 4-byte global stores (the
//     lines are still in L2: no extra DRAM traffic);
//   * with 6.9 KB per warp, E2_WARPS = 16 warps share an SM, so the dependent chain
//     "state -> rebuild -> fence -> issue" of any one env is hidden behind 15 others and the TMA
//     engine never runs dry (the image-per-warp encoder tops out at 4 images per SM).
#pragma once
#include "pb_encode.cuh"

namespace pb {

constexpr int E2_WARPS = 16;
constexpr int E2_THREADS = E2_WARPS * 32;
constexpr int E2_ZERO_CH = 13;                                  // channels 2..14
constexpr int E2_ZERO_BYTES = E2_ZERO_CH * CH_BYTES;            // 45 136
constexpr int E2_PRIV_BYTES = 2 * CH_BYTES;                     // ch 1 + ch 15
constexpr int E2_SMEM = E2_ZERO_BYTES + CH_BYTES + E2_WARPS * E2_PRIV_BYTES;
static_assert(E2_SMEM <= 227 * 1024, "template encoder must fit the 227 KB shared memory");

__global__ void __launch_bounds__(E2_THREADS, 1)
k_encode_obs_tpl(StateView st, float *__restrict__ out, long long env_stride, long long env_first,
                 long long env_count, unsigned int *__restrict__ ticket,
                 const long long *__restrict__ slot_index, long long slot_stride) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *zeros = reinterpret_cast<float *>(smem_raw);
    float *wall_ch = reinterpret_cast<float *>(smem_raw + E2_ZERO_BYTES);
    float *ch1 = reinterpret_cast<float *>(smem_raw + E2_ZERO_BYTES + CH_BYTES +
                                           (size_t)warp * E2_PRIV_BYTES);
    float *ch15 = ch1 + CH_FLOATS;
    float4 *ch1_4 = reinterpret_cast<float4 *>(ch1), *ch15_4 = reinterpret_cast<float4 *>(ch15);

    for (int q = threadIdx.x; q < E2_ZERO_BYTES / 16; q += E2_THREADS)
        reinterpret_cast<float4 *>(zeros)[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = threadIdx.x; c < CH_FLOATS; c += E2_THREADS) {
        const int col = c / ROWS, row = ROWS - 1 - (c % ROWS);
        wall_ch[c] = ((c_tab.walls[row] >> col) & 1u) ? 1.0f : 0.0f;
    }
    fence_proxy_async_smem();
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // ring-slot indirection: the slot number lives in device memory so that a CUDA graph can be
    // replayed while the replay ring advances.  Read AFTER the wait: under programmatic dependent
    // launch nothing written by the preceding grid is visible before it.
    if (slot_index) out += *reinterpret_cast<const volatile long long *>(slot_index) * slot_stride;

    // tickets: `cur` in hand, `pend` (lane 0) requested one env ago
    unsigned int cur = 0, pend = 0;
    if (lane == 0) {
        cur = atomicAdd(ticket, 1u);
        pend = atomicAdd(ticket, 1u);
    }
    cur = __shfl_sync(0xffffffffu, cur, 0);
    while ((long long)cur < env_count) {
        const long long k = cur;
        const long long i = env_first + k;
        const uint32_t F = (lane < PEL_WORDS) ? st.pel[(long long)lane * st.npad + i] : 0u;
        const uint4 c0 = st.core0[i], c1 = st.core1[i];
        const uint4 ga = st.ghosts[i], gb = st.ghosts[st.npad + i];
        EnvR s;
        unpack_env(c0, c1, ga, gb, s);

        // ---- per-lane role (branch free; same lane map as k_encode_obs) ----
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
        const int cell = (grp == 4) ? g_cell : ((grp == 5 || grp == 6) ? sup_cell
                                               : obs_cell(FRUIT_ROW, FRUIT_COL));
        const uint32_t cell_word = __shfl_sync(0xffffffffu, F, cell >> 5);
        const bool cell_pellet = (cell_word >> (cell & 31)) & 1u;
        const bool last_pellet = s.num_pellets == 1;
        const float pv = last_pellet ? (float)(PELLET_POINTS + LAST_REWARD) / 200.0f
                                     : (float)PELLET_POINTS / 200.0f;
        const float sv = last_pellet ? (float)(SUPER_PELLET_POINTS + LAST_REWARD) / 200.0f
                                     : (float)SUPER_PELLET_POINTS / 200.0f;
        const float sp = __fdiv_rn((float)s.tps, (float)s.period);
        const int my_timer_ch = gfr > 0 ? 14 : (s.mode == CHASE ? 13 : 12);
        bool fright_here = false, earlier_fright_here = false, later_same_timer = false;
        float bonus_v;
        {
            const bool is_sup = (grow == 3 || grow == 23) && (gcol == 1 || gcol == 26);
            const bool is_fruit = s.fruit_steps > 0 && grow == FRUIT_ROW && gcol == FRUIT_COL;
            float v = cell_pellet ? (is_sup ? sv : pv)
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
            bonus_v = v;
        }
        const float timer_v = gfr > 0 ? c_tab.fright_q[gfr & 0xff] : c_tab.mode_q[s.mode_steps & 0xff];
        // where each lane's single cell goes: `in_ch1` cells are written into the private ch-1
        // buffer before the copy, the others into global memory after it (offsets in floats
        // from the start of this env's observation)
        bool do_store, in_ch1;
        int off;
        float val;
        if (grp == 0) {
            do_store = pac_on && lane < 2; in_ch1 = false; off = (2 + lane) * CH_FLOATS + pac_cell; val = 1.0f;
        } else if (grp <= 2) {
            do_store = g_on; in_ch1 = false; off = lane * CH_FLOATS + g_cell; val = 1.0f;
        } else if (grp == 3) {
            do_store = g_on && !later_same_timer; in_ch1 = false; off = my_timer_ch * CH_FLOATS + g_cell; val = timer_v;
        } else if (grp == 4) {
            do_store = g_on && gfr > 0 && !earlier_fright_here; in_ch1 = true; off = g_cell; val = bonus_v;
        } else if (grp == 5) {
            do_store = cell_pellet; in_ch1 = false; off = 16 * CH_FLOATS + sup_cell; val = 1.0f;
        } else if (grp == 6) {
            do_store = cell_pellet && !fright_here; in_ch1 = true; off = sup_cell; val = sv;
        } else {
            do_store = lane == 28 && s.fruit_steps > 0 && !cell_pellet && !fright_here;
            in_ch1 = true; off = cell; val = (float)FRUIT_POINTS / 200.0f;
        }

        // ---- private channels (free: the previous env's copies were waited for completely) ----
#pragma unroll
        for (int it = 0; it < 7; it++) {
            const int q = lane + 32 * it;
            const uint32_t w = __shfl_sync(0xffffffffu, F, (lane >> 3) + 4 * it);
            if (q < CH_VEC) {
                const uint32_t nib = w >> (4 * (lane & 7));
                float4 v;
                v.x = (nib & 1u) ? pv : 0.0f;
                v.y = (nib & 2u) ? pv : 0.0f;
                v.z = (nib & 4u) ? pv : 0.0f;
                v.w = (nib & 8u) ? pv : 0.0f;
                ch1_4[q] = v;
                ch15_4[q] = make_float4(sp, sp, sp, sp);
            }
        }
        if (!warp_arrived((unsigned int)k)) __trap();
        if (do_store && in_ch1) ch1[off] = val;
        fence_proxy_async_smem();
        const bool ready = warp_arrived((unsigned int)k);
        float *dst = out + k * env_stride;
        if (lane == 0 && ready) {
            char *d = reinterpret_cast<char *>(dst);
            bulk_store_s2g(d, wall_ch, CH_BYTES);                               // ch 0
            bulk_store_s2g(d + CH_BYTES, ch1, CH_BYTES);                        // ch 1
            bulk_store_s2g(d + 2 * CH_BYTES, zeros, E2_ZERO_BYTES);             // ch 2..14
            bulk_store_s2g(d + 15 * CH_BYTES, ch15, CH_BYTES);                  // ch 15
            bulk_store_s2g(d + 16 * CH_BYTES, zeros, CH_BYTES);                 // ch 16
            bulk_commit();
            bulk_wait_all();  // complete (written), not just read: the cells below land on top
        }
        // fetch the next ticket while lane 0 waits
        const unsigned int nxt = __shfl_sync(0xffffffffu, pend, 0);
        if (lane == 0) pend = atomicAdd(ticket, 1u);
        if (!warp_arrived((unsigned int)k)) __trap();
        if (do_store && !in_ch1) dst[off] = val;
        cur = nxt;
    }
    if (lane == 0) {
        const unsigned int finished = atomicAdd(ticket + 1, 1u);
        if (finished == gridDim.x * E2_WARPS - 1) {
            atomicExch(ticket, 0u);
            atomicExch(ticket + 1, 0u);
        }
    }
}

}  // namespace pb
