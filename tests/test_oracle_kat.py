"""CPU tests: the oracle against the in-tree facts that pin behaviour (SURVEY.md 8c K1-K7,
Appendix G) and against tests/golden/maze.json (generated from the reference's own data by
tools/make_golden.py).  The reference has NO golden vectors of its own; the engine half of the
oracle is therefore 'parity unpinned' -- tests that depend on engine assumptions live in
test_oracle_engine_assumptions below and are named as such."""
import numpy as np
import pytest

STAY, DOWN, UP, LEFT, RIGHT = 0, 1, 2, 3, 4


def fresh(oracle, n=1, cfg=0, tape=None):
    b = oracle.OracleBatch(n, np.full(n, cfg, np.uint8), tape)
    return b


def place(b, row, col, **kw):
    st = b.export_state()
    st["pac_row"], st["pac_col"] = row, col
    for k, v in kw.items():
        st[k] = v
    b.import_state(st)
    return st


# ---------------------------------------------------------------- K1 / K4: walls and wall channel
def test_walls_match_node_coords(oracle, golden):
    L = oracle.lib()
    nodes = {tuple(n) for n in golden["node_coords"]}
    for r in range(-2, 34):
        for c in range(-2, 34):
            exp = 1 if not (0 <= r < 31 and 0 <= c < 28) else int((r, c) not in nodes)
            assert L.po_wall_at(r, c) == exp, (r, c)


def test_obs_wall_channel_and_shape(oracle, golden):
    b = fresh(oracle, 2, cfg=3, tape=np.arange(2 * 8, dtype=np.uint32).reshape(2, 8) * 0x9E3779B1)
    b.reset()
    o = b.obs()
    assert o.shape == (2, 17, 28, 31) and o.dtype == np.float32
    for i in range(2):
        assert np.array_equal(o[i, 0].ravel(), np.array(golden["wall_channel_flat"], np.float32))
        assert o[i, 0].sum() == 580.0


# ---------------------------------------------------------------- K2: action mask
def test_action_mask_every_node(oracle, golden):
    nodes = golden["node_coords"]
    b = fresh(oracle, len(nodes))
    st = b.export_state()
    st["pac_row"] = [n[0] for n in nodes]
    st["pac_col"] = [n[1] for n in nodes]
    b.import_state(st)
    assert b.action_mask().tolist() == golden["valid_actions"]
    for key, exp in golden["mask_kat"].items():
        r, c = map(int, key.split(","))
        assert b.action_mask()[nodes.index([r, c])].tolist() == exp


# ---------------------------------------------------------------- K3 / K4: obs index map
def test_obs_index_map_and_identities(oracle):
    b = fresh(oracle, 1)
    b.step(np.array([STAY], np.uint8))  # first update tick puts the ghosts on the board
    st = b.export_state()[0]
    o = b.obs()[0]
    # super pellets appear in ch16 at (x, y) = (col, 30 - row)
    assert sorted(map(tuple, np.argwhere(o[16] == 1.0).tolist())) == sorted(
        [(1, 27), (26, 27), (1, 7), (26, 7)])
    assert np.array_equal(o[2], o[3])
    assert np.array_equal(o[8:12], o[4:8])
    assert np.all(o[15] == np.float32(st["ticks_per_step"]) / np.float32(st["update_period"]))
    assert o[3, st["pac_col"], 30 - st["pac_row"]] == 1.0 and o[3].sum() == 1.0
    for g in range(4):
        assert o[4 + g].sum() == 1.0
        assert o[4 + g, st["g_col"][g], 30 - st["g_row"][g]] == 1.0
    ghost_cells = {(int(st["g_col"][g]), 30 - int(st["g_row"][g])) for g in range(4)}
    for ch in (12, 13, 14):
        assert {tuple(x) for x in np.argwhere(o[ch] != 0).tolist()} <= ghost_cells
    assert not (o[12].any() and o[13].any())
    # channel 1: pellets are worth 10/200, super pellets 50/200 (env.rs:427-441)
    assert o[1, 1, 27] == np.float32(50) / np.float32(200)
    assert o[1, 2, 29] == np.float32(10) / np.float32(200)


def test_obs_last_pellet_and_fruit_and_fright_bonus(oracle):
    b = fresh(oracle, 1)
    st = b.export_state()
    st["pellets"][:] = 0
    st["pellets"][0, 5] = 1 << 6          # one ordinary pellet left at (5, 6)
    st["num_pellets"] = 1
    st["fruit_steps"] = 7
    st["ghost_combo"] = 2
    st["g_row"][0, 1], st["g_col"][0, 1], st["g_fright"][0, 1] = 5, 6, 33   # frightened, on pellet
    st["g_row"][0, 2], st["g_col"][0, 2], st["g_fright"][0, 2] = 5, 6, 33   # second one, same cell
    st["g_row"][0, 0], st["g_col"][0, 0] = 8, 1
    st["mode"], st["mode_steps"] = 2, 77
    b.import_state(st)
    o = b.obs()[0]
    f = np.float32
    assert o[1, 6, 25] == f(f(3010) / f(200)) + f(4) + f(4)   # env.rs:432,464 (+= twice)
    assert o[1, 13, 13] == f(100) / f(200)                    # fruit cell (17, 13)
    assert o[14, 6, 25] == f(33) / f(40)
    assert o[13, 1, 22] == f(77) / f(180) and not o[12].any()


# ---------------------------------------------------------------- K5: reward algebra
@pytest.mark.parametrize("last,act,exp_pen", [
    (STAY, LEFT, 0), (LEFT, LEFT, 0), (LEFT, STAY, -2), (LEFT, RIGHT, -2), (UP, DOWN, -2),
    (STAY, STAY, 0), (RIGHT, RIGHT, 0), (DOWN, LEFT, -2),
])
def test_turn_penalty_table(oracle, last, act, exp_pen):
    b = fresh(oracle, 1)
    # park Pac-Man in a pellet-free corridor cell far from everything: (14, 7)
    place(b, 14, 7, last_action=last, pellets=0, num_pellets=100)
    r, d = b.step(np.array([act], np.uint8))
    assert (int(r[0]), bool(d[0])) == (exp_pen, False)


def test_proximity_penalty_and_score_delta(oracle):
    # Appendix G: delta=10, last=Left, act=Stay, one non-frightened ghost at Manhattan 2 -> -12
    b = fresh(oracle, 1)
    st = b.export_state()
    st["pac_row"], st["pac_col"] = 5, 6
    st["last_action"] = LEFT
    st["curr_ticks"] = 1                       # no ghost update during the 8 ticks (1..8)
    st["g_row"][0, 0], st["g_col"][0, 0] = 5, 8   # distance 2
    st["g_next_row"][0, 0], st["g_next_col"][0, 0] = 5, 8
    b.import_state(st)
    r, d = b.step(np.array([STAY], np.uint8))     # Stay re-enters the cell: eats its pellet (+10)
    assert (int(r[0]), bool(d[0])) == (10 - 2 - 20, False)
    # distance 1 -> -50 ; frightened ghosts are ignored ; distance 0 matches neither arm (Q7)
    for dist_cells, fright, exp in [((5, 7), 0, -50), ((5, 7), 5, 0), ((5, 9), 0, 0)]:
        b = fresh(oracle, 1)
        st = b.export_state()
        st["pac_row"], st["pac_col"], st["curr_ticks"] = 5, 6, 1
        st["pellets"][:] = 0
        st["g_row"][0, 3], st["g_col"][0, 3] = dist_cells
        st["g_fright"][0, 3] = fright
        b.import_state(st)
        r, d = b.step(np.array([STAY], np.uint8))
        assert int(r[0]) == exp, (dist_cells, fright)


def test_terminal_rewards(oracle):
    # death: -200 ; ghost walks... simplest: non-frightened ghost on Pac-Man's target cell
    b = fresh(oracle, 1)
    st = b.export_state()
    st["pac_row"], st["pac_col"], st["curr_ticks"] = 5, 6, 1
    st["pellets"][:] = 0
    st["g_row"][0, 0], st["g_col"][0, 0] = 5, 7
    b.import_state(st)
    r, d = b.step(np.array([RIGHT], np.uint8))
    assert bool(d[0]) and int(r[0]) == -200        # all locs are (32,32) afterwards: no proximity term
    assert b.export_state()["curr_lives"][0] == 2
    # board cleared: +3000 on top of the last pellet's 10 points
    b = fresh(oracle, 1)
    st = b.export_state()
    st["pac_row"], st["pac_col"], st["curr_ticks"] = 5, 6, 1
    st["pellets"][:] = 0
    st["pellets"][0, 5] = 1 << 7
    st["num_pellets"] = 1
    b.import_state(st)
    r, d = b.step(np.array([RIGHT], np.uint8))
    assert bool(d[0]) and int(r[0]) == 10 + 3000
    assert b.export_state()["curr_level"][0] == 2 and b.export_state()["curr_lives"][0] == 3


# ---------------------------------------------------------------- move rules
def test_blocked_move_and_stay(oracle):
    b = fresh(oracle, 1)
    place(b, 1, 1, curr_ticks=1)
    before = b.export_state()[0].copy()
    r, d = b.step(np.array([UP], np.uint8))         # wall above (1,1): no move, no pellet eaten
    after = b.export_state()[0]
    assert (after["pac_row"], after["pac_col"]) == (1, 1)
    assert after["num_pellets"] == before["num_pellets"] and int(r[0]) == 0
    r, d = b.step(np.array([STAY], np.uint8))       # Stay calls set_pacman_location(same cell)
    after2 = b.export_state()[0]
    assert after2["num_pellets"] == before["num_pellets"] - 1
    assert int(r[0]) == 10 - 2                       # and pays the turn penalty (Q6)


def test_invalid_action_is_value_error(oracle):
    b = fresh(oracle, 2)
    before = b.export_state().copy()
    with pytest.raises(ValueError, match="Invalid action"):
        b.step(np.array([0, 5], np.uint8))
    assert np.array_equal(b.export_state()[1], before[1])


# ---------------------------------------------------------------- K6 / K7: reset algebra
def test_reset_wipe_halves_and_tick_bounds(oracle, pb_maze=None):
    from importlib import import_module

    maze = import_module("pacbot_rl_b200").maze
    n = 5 * 10 * 4
    tape = np.zeros((n, 3), np.uint32)
    k = 0
    combos = []
    for wipe in range(5):
        for tick in range(10):
            for node in (0, 100, 200, 287):
                # draw order: ticks (env.rs:148), node (:152), wipe type (:185); k = (raw*n)>>32
                tape[k] = [((tick << 32) // 10) + 1, ((node << 32) // 288) + 1, ((wipe << 32) // 5) + 1]
                combos.append((wipe, tick, node))
                k += 1
    b = fresh(oracle, n, cfg=3, tape=tape)
    b.reset()
    st = b.export_state()
    for i, (wipe, tick, node) in enumerate(combos):
        assert st["ticks_per_step"][i] == 4 + tick
        assert (st["pac_row"][i], st["pac_col"][i]) == maze.NODE_COORDS[node]
        assert st["pac_dir"][i] == 0 and st["last_action"][i] == 0 and st["last_score"][i] == 0
        exp = []
        for r in range(31):
            keep = 0
            for c in range(28):
                cleared = [False, c < 14, c >= 14, r < 15, r >= 15][wipe]
                if not cleared:
                    keep |= 1 << c
            exp.append(maze.PELLET_ROWS[r] & keep)
        assert st["pellets"][i].tolist() == exp
        assert st["num_pellets"][i] == sum(bin(w).count("1") for w in exp)
        assert st["rng_ctr"][i] == 3 and st["paused"][i] == 0
    # not random: ticks stay 8, spawn cell, full board, no draws consumed
    b = fresh(oracle, 1, cfg=0)
    b.reset()
    s = b.export_state()[0]
    assert (s["ticks_per_step"], s["pac_row"], s["pac_col"], s["num_pellets"], s["rng_ctr"]) == (8, 23, 13, 244, 0)


def test_first_ai_done(oracle):
    b = fresh(oracle, 1)
    assert not b.first_ai_done()[0]
    st = b.export_state()
    for r, c in ((3, 1), (3, 26), (23, 1), (23, 26)):
        st["pellets"][0, r] &= ~np.uint32(1 << c)
    b.import_state(st)
    assert b.first_ai_done()[0]
    st["g_fright"][0, 2] = 3
    b.import_state(st)
    assert not b.first_ai_done()[0]


# ---------------------------------------------------------------- engine-assumption dependent
def test_oracle_engine_assumptions_fresh_game(oracle):
    """[UPSTREAM-RECALL] dependent: documents what the oracle assumes about a fresh game."""
    b = fresh(oracle, 1)
    s = b.export_state()[0]
    assert (s["curr_lives"], s["curr_level"], s["curr_score"], s["update_period"]) == (3, 1, 0, 12)
    assert (s["mode"], s["mode_steps"], s["level_steps"]) == (1, 60, 960)
    assert (s["pac_row"], s["pac_col"], s["pac_dir"]) == (23, 13, 3)
    assert s["g_trapped"].tolist() == [0, 5, 16, 32]
    r, d = b.step(np.array([LEFT], np.uint8))      # first step Left from spawn: +10, not done
    assert (int(r[0]), bool(d[0])) == (10, False)
    s = b.export_state()[0]
    assert list(zip(s["g_row"].tolist(), s["g_col"].tolist())) == [(11, 13), (13, 13), (14, 11), (14, 15)]
    o = b.obs()[0]
    assert o[12, 13, 19] == np.float32(59) / np.float32(180)
    assert o[15, 0, 0] == np.float32(8) / np.float32(12)


def test_oracle_long_random_rollout_invariants(oracle):
    """Size-independent invariants over a long random rollout with deaths and resets."""
    n, steps = 64, 600
    rng = np.random.default_rng(5)
    tape = rng.integers(0, 2**32, size=(n, 251), dtype=np.uint32)
    b = fresh(oracle, n, cfg=3, tape=tape)
    b.reset()
    episodes = 0
    for t in range(steps):
        m = b.action_mask()
        a = np.array([rng.choice(np.flatnonzero(x)) for x in m], np.uint8)
        r, d = b.step(a, auto_reset=True)
        episodes += int(d.sum())
        st = b.export_state()
        assert np.array_equal(st["num_pellets"], [sum(bin(int(w)).count("1") for w in p) for p in st["pellets"]])
        assert (st["curr_lives"] == 3).all() and (st["curr_level"] == 1).all()  # auto reset
        assert ((st["ticks_per_step"] >= 4) & (st["ticks_per_step"] <= 13)).all()
        if t % 100 == 0:
            o = b.obs()
            assert (o[:, 0].reshape(n, -1).sum(1) == 580).all()
            assert np.array_equal(o[:, 2], o[:, 3]) and np.array_equal(o[:, 4:8], o[:, 8:12])
    assert episodes > 20


def test_oracle_engine_assumption_house_exit_paths(oracle):
    """[UPSTREAM-RECALL] dependent (ledger E5/E15): on a fresh game with Pac-Man parked on his
    spawn cell (12 ticks per step = one ghost update per step) the ghosts leave the house after
    0 / 5 / 16 / 32 trapped bounces, are lured to red's spawn cell (11,13) while `spawning`, and
    from that cell on head for their scatter corners: orange turns LEFT along row 11 (a target of
    the house exit (12,13) would send it up at (11,12))."""
    b = fresh(oracle, 1)
    place(b, 23, 13, ticks_per_step=12)
    cells = [[] for _ in range(4)]
    for t in range(44):
        r, d = b.step(np.array([STAY], np.uint8))
        assert not d[0]
        s = b.export_state()[0]
        for i in range(4):
            cells[i].append((int(s["g_row"][i]), int(s["g_col"][i]), int(s["g_spawning"][i])))
    assert cells[0][:4] == [(11, 13, 1), (11, 12, 0), (10, 12, 0), (9, 12, 0)]
    assert cells[1][5:11] == [(14, 13, 1), (13, 13, 1), (12, 13, 1), (11, 13, 1), (11, 12, 0), (10, 12, 0)]
    assert cells[2][17:24] == [(13, 11, 1), (13, 12, 1), (13, 13, 1), (12, 13, 1), (11, 13, 1),
                               (11, 14, 0), (11, 15, 0)]
    assert cells[3][33:43] == [(13, 15, 1), (13, 14, 1), (13, 13, 1), (12, 13, 1), (11, 13, 1),
                               (11, 12, 0), (11, 11, 0), (11, 10, 0), (11, 9, 0), (12, 9, 0)]


def test_action_mask_literal_examples(oracle, golden):
    """SURVEY.md Appendix G, spelled out: masks are [Stay, Down, Up, Left, Right] (env.rs:293-302);
    the only four-way junctions are (5,6), (5,21), (20,6), (20,21); degree histogram 254/30/4."""
    T, F = True, False
    want = {(23, 13): [T, F, F, T, T], (1, 1): [T, T, F, F, T], (5, 6): [T, T, T, T, T],
            (29, 26): [T, F, T, T, F], (14, 9): [T, T, T, T, F], (11, 13): [T, F, F, T, T]}
    b = fresh(oracle, len(want))
    st = b.export_state()
    for i, (r, c) in enumerate(want):
        st["pac_row"][i], st["pac_col"][i] = r, c
    b.import_state(st)
    assert b.action_mask().tolist() == list(want.values())
    nodes = [tuple(n) for n in golden["node_coords"]]
    four = [n for n, va in zip(nodes, golden["valid_actions"]) if sum(va[1:]) == 4]
    assert four == [(5, 6), (5, 21), (20, 6), (20, 21)]
    assert golden["degree_histogram"] == {"2": 254, "3": 30, "4": 4}
