// pb_engine.cuh -- device game logic, one env per thread, whole env state in registers.
//
// Reference semantics:
//   wrapper  /root/reference/pacbot_rs/src/env.rs (step :216-267, move_one_cell :395-407,
//            reset :140-212, is_done :277-279, action_mask :293-302, apply_configuration :377-384)
//   engine   crate pacbot-rs @423622af (`pacbot_rs_2`, not in tree): GameState::step,
//            set_pacman_location, ghost update/plan, collisions, step events -- SURVEY.md
//            Appendix A/B; assumptions listed in DESIGN.md "Engine assumption ledger".
//
// This is an independent implementation (registers + bit tricks + tick skipping); it is checked
// bit-for-bit against the scalar C oracle in tests/ but shares no code with it.
#pragma once
#include "pb_state.cuh"

namespace pb {

struct RngCtx {
    uint64_t seed;
    uint64_t gid;            // global env id
    const uint32_t *tape;    // this env's tape row (or nullptr)
    uint32_t tape_len;
};

__device__ __forceinline__ uint32_t draw_below(EnvR &s, const RngCtx &r, uint32_t n) {
    const uint32_t raw =
        r.tape ? r.tape[s.rng_ctr % r.tape_len] : draw_raw(r.seed, r.gid, s.rng_ctr);
    s.rng_ctr += 1u;
    return below(raw, n);
}

// pellet words of one env, strided by npad in HBM
struct PelRef {
    uint32_t *p;
    long long stride;
    __device__ __forceinline__ uint32_t get(int w) const { return p[(long long)w * stride]; }
    __device__ __forceinline__ void set(int w, uint32_t v) const { p[(long long)w * stride] = v; }
};

__device__ __forceinline__ int d_row(int d) { return d == UP ? -1 : (d == DOWN ? 1 : 0); }
__device__ __forceinline__ int d_col(int d) { return d == LEFT ? -1 : (d == RIGHT ? 1 : 0); }
__device__ __forceinline__ int reversed(int d) { return d < 4 ? (d ^ 2) : NONE; }

__device__ __forceinline__ int spawn_row(int i) { return i == 0 ? 11 : (i == 1 ? 13 : 14); }
__device__ __forceinline__ int spawn_col(int i) { return i == 2 ? 11 : (i == 3 ? 15 : 13); }
__device__ __forceinline__ int spawn_dir(int i) { return i == 0 ? LEFT : (i == 1 ? DOWN : UP); }
__device__ __forceinline__ int trapped_init(int i) {
    return i == 0 ? 0 : (i == 1 ? 5 : (i == 2 ? 16 : 32));
}
__device__ __forceinline__ int scatter_row(int i) { return i < 2 ? -3 : 31; }
__device__ __forceinline__ int scatter_col(int i) {
    return i == 0 ? 25 : (i == 1 ? 2 : (i == 2 ? 27 : 0));
}

// walls: 32 row bitboards in shared memory (row index differs per lane: 31 banks, no conflicts)
__device__ __forceinline__ bool wall_at(const uint32_t *walls, int r, int c) {
    if ((unsigned)r >= (unsigned)ROWS || (unsigned)c >= (unsigned)COLS) return true;
    return (walls[r] >> c) & 1u;
}

__device__ __forceinline__ bool is_done(const EnvR &s) {  // env.rs:277-279
    return s.lives < INIT_LIVES || s.level != INIT_LEVEL;
}

__device__ __forceinline__ void add_score(EnvR &s, int pts) { s.score = (s.score + pts) & 0xffff; }

__device__ __forceinline__ void reset_all_ghosts(EnvR &s) {
    s.gflags |= 0x0f;  // spawning; `eaten` is left alone
#pragma unroll
    for (int i = 0; i < 4; i++) {
        GhostR &g = s.g[i];
        g.trapped = trapped_init(i);
        g.fright = 0;
        g.row = EMPTY;
        g.col = EMPTY;
        g.dir = NONE;
        g.nrow = spawn_row(i);
        g.ncol = spawn_col(i);
        g.ndir = spawn_dir(i);
    }
}

// GameState::new() with the wrapper's paused=false (env.rs:134,211); pellets handled by caller
__device__ __forceinline__ void fresh_game(EnvR &s) {
    s.ticks = 0;
    s.period = INIT_UPDATE_PERIOD;
    s.mode = SCATTER;
    s.mode_steps = SCATTER_DURATION;
    s.flags = 0;
    s.level_steps = LEVEL_DURATION;
    s.score = 0;
    s.level = INIT_LEVEL;
    s.lives = INIT_LIVES;
    s.combo = 0;
    s.fruit_steps = 0;
    s.prow = PAC_SPAWN_ROW;
    s.pcol = PAC_SPAWN_COL;
    s.pdir = RIGHT;
    s.num_pellets = INIT_PELLET_COUNT;
    s.gflags = 0;
    reset_all_ghosts(s);
}

__device__ __forceinline__ void lower_update_period(EnvR &s) { s.period = max(1, s.period - 2); }

__device__ __forceinline__ void death_reset(EnvR &s) {
    s.flags |= F_PAUSE_ON_UPDATE;
    s.prow = EMPTY;
    s.pcol = EMPTY;
    s.pdir = NONE;
    if (s.lives > 0) s.lives--;
    if (s.lives > 0) {
        s.mode = SCATTER;
        s.mode_steps = SCATTER_DURATION;
    }
    s.fruit_steps = 0;
    reset_all_ghosts(s);
}

__device__ __forceinline__ void check_collisions(EnvR &s) {
    if (s.prow == EMPTY && s.pcol == EMPTY) return;  // Pac-Man is off the board: nothing collides
    int eat = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const GhostR &g = s.g[i];
        if (g.row == s.prow && g.col == s.pcol) {
            if ((s.gflags >> (4 + i)) & 1) continue;  // already eaten
            if (g.fright > 0) {
                eat |= 1 << i;
            } else {
                death_reset(s);
                return;
            }
        }
    }
    if (eat == 0) return;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (!((eat >> i) & 1)) continue;
        GhostR &g = s.g[i];
        s.gflags |= (1 << i) | (1 << (4 + i));  // spawning + eaten
        g.row = EMPTY;
        g.col = EMPTY;
        g.dir = NONE;
        const int sp = (i == 0) ? 1 : i;  // red respawns inside the house, on pink's spot
        g.nrow = spawn_row(sp);
        g.ncol = spawn_col(sp);
        g.ndir = UP;
        add_score(s, COMBO_MULTIPLIER << s.combo);
        s.combo++;
    }
}

__device__ __forceinline__ void level_reset(EnvR &s, const PelRef &pel, const MazeTables &tab) {
    s.flags |= F_PAUSE_ON_UPDATE;
    s.prow = EMPTY;
    s.pcol = EMPTY;
    s.pdir = NONE;
    s.level_steps = LEVEL_DURATION;
    s.mode = SCATTER;
    s.mode_steps = SCATTER_DURATION;
    s.fruit_steps = 0;
    reset_all_ghosts(s);
    for (int w = 0; w < PEL_WORDS; w++) pel.set(w, tab.init_pel[w]);
    s.num_pellets = INIT_PELLET_COUNT;
    if (s.level != 255) {
        s.level++;
        s.period = max(1, INIT_UPDATE_PERIOD - 2 * (s.level - 1));
    }
}

__device__ __forceinline__ void collect_pellet(EnvR &s, const PelRef &pel, const MazeTables &tab,
                                               int row, int col) {
    if (s.fruit_steps > 0 && s.prow == FRUIT_ROW && s.pcol == FRUIT_COL) {
        s.fruit_steps = 0;
        add_score(s, FRUIT_POINTS);
    }
    const int i = fbit(row, col);
    const int w = i >> 5;
    const uint32_t m = 1u << (i & 31);
    const uint32_t word = pel.get(w);
    if (!(word & m)) return;
    pel.set(w, word & ~m);
    if (s.num_pellets > 0) s.num_pellets--;
    const bool super_pellet = (row == 3 || row == 23) && (col == 1 || col == 26);
    if (super_pellet) {
        s.combo = 0;
#pragma unroll
        for (int g = 0; g < 4; g++) {
            s.g[g].fright = GHOST_FRIGHT_STEPS;
            if (s.g[g].trapped == 0) s.g[g].trapped = 1;  // forces a reversal
        }
        add_score(s, SUPER_PELLET_POINTS);
    } else {
        add_score(s, PELLET_POINTS);
    }
    const int n = s.num_pellets;
    if ((n == FRUIT_THRESHOLD_1 || n == FRUIT_THRESHOLD_2) && s.fruit_steps == 0)
        s.fruit_steps = FRUIT_DURATION;
    if (n == ANGER_THRESHOLD_1 || n == ANGER_THRESHOLD_2) {
        lower_update_period(s);
        s.mode = CHASE;
        s.mode_steps = CHASE_DURATION;
    } else if (n == 0) {
        level_reset(s, pel, tab);
    }
}

// GameState::set_pacman_location (env.rs:405)
__device__ __forceinline__ void set_pacman_location(EnvR &s, const PelRef &pel,
                                                    const MazeTables &tab, int row, int col) {
    if (s.flags & F_PAUSED) return;
    const int dr = row - s.prow, dc = col - s.pcol;
    if (dr == -1 && dc == 0) s.pdir = UP;
    else if (dr == 0 && dc == -1) s.pdir = LEFT;
    else if (dr == 1 && dc == 0) s.pdir = DOWN;
    else if (dr == 0 && dc == 1) s.pdir = RIGHT;
    s.prow = row;
    s.pcol = col;
    collect_pellet(s, pel, tab, row, col);
    check_collisions(s);
}

__device__ __forceinline__ void plan_ghost(EnvR &s, const int i, const uint32_t *walls,
                                           const RngCtx &rng) {
    GhostR &g = s.g[i];
    if (g.row == EMPTY && g.col == EMPTY) return;
    g.nrow = g.row + d_row(g.dir);
    g.ncol = g.col + d_col(g.dir);
    g.ndir = g.dir;
    if (g.trapped > 0) {
        g.ndir = reversed(g.ndir);
        g.trapped--;
        return;
    }
    int tr = 0, tc = 0;
    if (s.mode == CHASE) {
        if (i == 0) {
            tr = s.prow;
            tc = s.pcol;
        } else if (i == 1 || i == 2) {
            const int k = (i == 1) ? 4 : 2;
            int ar = s.prow + d_row(s.pdir) * k;
            int ac = s.pcol + d_col(s.pdir) * k - (s.pdir == UP ? k : 0);
            if (i == 1) {
                tr = ar;
                tc = ac;
            } else {
                tr = 2 * ar - s.g[0].row;
                tc = 2 * ac - s.g[0].col;
            }
        } else {
            const int dr = g.row - s.prow, dc = g.col - s.pcol;
            if (dr * dr + dc * dc >= 64) {
                tr = s.prow;
                tc = s.pcol;
            } else {
                tr = scatter_row(3);
                tc = scatter_col(3);
            }
        }
    } else if (s.mode == SCATTER) {
        tr = scatter_row(i);
        tc = scatter_col(i);
    }
    // a spawning ghost heads for red's spawn cell (11,13), just outside the ghost house, until it
    // stands on it or is about to (ledger item E15)
    const bool spawning = (s.gflags >> i) & 1;
    const bool out_here = g.row == RED_SPAWN_ROW && g.col == RED_SPAWN_COL;
    const bool out_next = g.nrow == RED_SPAWN_ROW && g.ncol == RED_SPAWN_COL;
    if (spawning && !out_here && !out_next) {
        tr = RED_SPAWN_ROW;
        tc = RED_SPAWN_COL;
    }
    const int rev = reversed(g.ndir);
    int valid = 0, best = UP, best_d = 0x7fffffff;
#pragma unroll
    for (int d = 0; d < 4; d++) {
        const int r = g.nrow + d_row(d), c = g.ncol + d_col(d);
        bool ok = !wall_at(walls, r, c);
        if (spawning) {
            if (r >= 13 && r <= 14 && c >= 11 && c <= 15) ok = true;  // inside the ghost house
            if (r == EXIT_ROW && c == EXIT_COL) ok = true;
        }
        if (d == rev) ok = false;
        if (ok) {
            valid |= 1 << d;
            const int dr = r - tr, dc = c - tc;
            const int dd = dr * dr + dc * dc;
            if (dd < best_d) {
                best_d = dd;
                best = d;
            }
        }
    }
    if (valid == 0) return;
    if (g.fright > 1) {
        // k-th set bit of `valid`, k uniform in [0, popcount)
        int k = (int)draw_below(s, rng, (uint32_t)__popc(valid));
        int v = valid;
        while (k-- > 0) v &= v - 1;
        g.ndir = __ffs(v) - 1;
        return;
    }
    g.ndir = best;
}

// one ghost-update tick of GameState::step (everything except the tick counter)
__device__ __forceinline__ void update_tick(EnvR &s, const uint32_t *walls, const RngCtx &rng) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
        GhostR &g = s.g[i];
        if (g.row == RED_SPAWN_ROW && g.col == RED_SPAWN_COL) s.gflags &= ~(1 << i);  // spawned
        if ((s.gflags >> (4 + i)) & 1) {
            s.gflags &= ~(1 << (4 + i));
            g.fright = 0;
        }
        if (g.fright > 0) g.fright--;
        g.row = g.nrow;
        g.col = g.ncol;
        g.dir = g.ndir;
    }
    if (s.prow == EMPTY && s.pcol == EMPTY) {  // try_respawn_pacman
        s.prow = PAC_SPAWN_ROW;
        s.pcol = PAC_SPAWN_COL;
        s.pdir = RIGHT;
    }
    if (s.flags & F_PAUSE_ON_UPDATE) s.flags = (s.flags | F_PAUSED) & ~F_PAUSE_ON_UPDATE;
    check_collisions(s);
    // handle_step_events
    if (s.mode_steps == 0) {
        if (s.mode == CHASE) {
            s.mode = SCATTER;
            s.mode_steps = SCATTER_DURATION;
        } else {
            s.mode = CHASE;
            s.mode_steps = CHASE_DURATION;
        }
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (s.g[i].trapped == 0) s.g[i].trapped = 1;
    } else {
        s.mode_steps--;
    }
    if (s.level_steps == 0) {
        lower_update_period(s);
        s.level_steps = LEVEL_PENALTY_DURATION;
    } else {
        s.level_steps--;
    }
    if (s.fruit_steps > 0) s.fruit_steps--;
#pragma unroll
    for (int i = 0; i < 4; i++) plan_ghost(s, i, walls, rng);
}

// PacmanGym::step (env.rs:216-267).  `action` must be in 0..4.
__device__ __forceinline__ void gym_step(EnvR &s, const PelRef &pel, const MazeTables &tab,
                                         const uint32_t *walls, const RngCtx &rng, int action,
                                         int &reward, bool &done) {
    // move_one_cell (env.rs:395-407)
    {
        int nr = s.prow, nc = s.pcol;
        if (action == A_RIGHT) nc += 1;
        else if (action == A_LEFT) nc -= 1;
        else if (action == A_UP) nr -= 1;
        else if (action == A_DOWN) nr += 1;
        if (!wall_at(walls, nr, nc)) set_pacman_location(s, pel, tab, nr, nc);
    }
    const int turn_penalty =
        (s.last_action == action || s.last_action == A_STAY) ? 0 : TURN_PENALTY;  // :221-225
    // tick loop (env.rs:226-231): ghosts only act when curr_ticks % update_period == 0, every
    // other tick is a counter increment -> jump from update tick to update tick.
    {
        int remaining = is_done(s) ? 1 : s.tps;  // at least one tick always runs
        while (remaining > 0 && !(s.flags & F_PAUSED)) {
            const uint32_t r = s.ticks % (uint32_t)s.period;
            if (r != 0) {
                const int skip = min((int)((uint32_t)s.period - r), remaining);
                s.ticks += (uint32_t)skip;
                remaining -= skip;
                continue;
            }
            update_tick(s, walls, rng);
            s.ticks += 1u;
            remaining -= 1;
            if (is_done(s)) break;
        }
    }
    s.last_action = action;
    done = is_done(s);
    int rew = (s.score - s.last_score) + turn_penalty;  // :240
    if (done) rew += (s.lives < 3) ? -200 : LAST_REWARD;  // :241-249
#pragma unroll
    for (int i = 0; i < 4; i++) {  // :251-263
        if (s.g[i].fright == 0) {
            const int d = abs(s.prow - s.g[i].row) + abs(s.pcol - s.g[i].col);
            rew += (d == 1) ? -50 : ((d == 2) ? -20 : 0);
        }
    }
    s.last_score = s.score;  // :264
    reward = rew;
}

// PacmanGym::new / reset (env.rs:133-138, 140-212).  randomise=false is the constructor path.
__device__ __forceinline__ void gym_reset(EnvR &s, const PelRef &pel, const MazeTables &tab,
                                          const uint32_t *walls, const RngCtx &rng,
                                          bool randomise) {
    s.last_score = 0;
    fresh_game(s);
    int wipe = 0;
    if (!randomise) {
        s.tps = NORMAL_TICKS_PER_STEP;  // env.rs:368
    } else {
        if (s.cfg & 2)  // random_ticks, :147-149
            s.tps = MIN_TICKS_PER_STEP +
                    (int)draw_below(s, rng, MAX_TICKS_PER_STEP - MIN_TICKS_PER_STEP);
        if (s.cfg & 1) {  // random_start, :151-208
            const uint32_t node = draw_below(s, rng, NUM_NODES);
            const int rc = tab.node_rc[node];
            s.prow = rc & 0xff;
            s.pcol = rc >> 8;
            s.pdir = UP;
            if (s.cfg & 4) {  // randomize_ghosts, :157-182 (Rust-only config)
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    GhostR &g = s.g[i];
                    g.trapped = 0;
                    const int grc = tab.node_rc[draw_below(s, rng, NUM_NODES)];
                    const int r = grc & 0xff, c = grc >> 8;
                    // first valid entry after Stay in [down, up, left, right]; the reference maps
                    // slot 3 -> Right and slot 4 -> Left (sic, env.rs:172-177)
                    int d = NONE;
                    if (!wall_at(walls, r + 1, c)) d = DOWN;
                    else if (!wall_at(walls, r - 1, c)) d = UP;
                    else if (!wall_at(walls, r, c - 1)) d = RIGHT;
                    else if (!wall_at(walls, r, c + 1)) d = LEFT;
                    g.row = g.nrow = r;
                    g.col = g.ncol = c;
                    g.dir = g.ndir = d;
                }
            }
            wipe = (int)draw_below(s, rng, 5);  // :185
        }
    }
    int count = 0;
    const bool no_super = s.cfg & 8;  // apply_configuration, :377-384
    for (int w = 0; w < PEL_WORDS; w++) {
        uint32_t v = tab.init_pel[w] & tab.wipe_keep[wipe][w];
        if (no_super) v &= ~tab.super_mask[w];
        pel.set(w, v);
        count += __popc(v);
    }
    s.num_pellets = count;
    s.last_action = A_STAY;  // :210
}

__device__ __forceinline__ uint32_t action_mask_bits(const EnvR &s, const uint32_t *walls) {
    // env.rs:293-302: [true, !wall(r+1,c), !wall(r-1,c), !wall(r,c-1), !wall(r,c+1)]
    uint32_t m = 1u;
    m |= (wall_at(walls, s.prow + 1, s.pcol) ? 0u : 1u) << 1;
    m |= (wall_at(walls, s.prow - 1, s.pcol) ? 0u : 1u) << 2;
    m |= (wall_at(walls, s.prow, s.pcol - 1) ? 0u : 1u) << 3;
    m |= (wall_at(walls, s.prow, s.pcol + 1) ? 0u : 1u) << 4;
    return m;
}

}  // namespace pb
