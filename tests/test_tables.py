"""CPU tests: maze tables, C-ABI surface, host logic (no GPU compute)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_maze_py_matches_golden(golden):
    import pacbot_rl_b200 as pb

    assert list(pb.maze.WALL_ROWS) == golden["wall_rows"]
    assert [list(n) for n in pb.maze.NODE_COORDS] == golden["node_coords"]
    for (r, c), va in zip(pb.maze.NODE_COORDS, golden["valid_actions"]):
        assert pb.maze.action_mask_at(r, c) == va
    assert sum(bin(w).count("1") for w in pb.maze.PELLET_ROWS) == 244
    for r, c in golden["super_pellets"]:
        assert (pb.maze.PELLET_ROWS[r] >> c) & 1
    # pellets only on walkable cells
    for r in range(31):
        assert pb.maze.PELLET_ROWS[r] & pb.maze.WALL_ROWS[r] == 0


def test_library_tables_match_golden(golden, pb):
    L = pb._lib.load()
    assert [L.pb_maze_wall_row(r) for r in range(31)] == golden["wall_rows"]
    assert [L.pb_maze_pellet_row(r) for r in range(31)] == list(pb.maze.PELLET_ROWS)
    for k, (r, c) in enumerate(golden["node_coords"]):
        rr, cc = C.c_int(), C.c_int()
        assert L.pb_maze_node(k, C.byref(rr), C.byref(cc)) == 0
        assert (rr.value, cc.value) == (r, c)
    assert L.pb_maze_node(288, C.byref(rr), C.byref(cc)) != 0


def test_header_symbols_all_exported(pb):
    """Every function include/pacbot_b200.h declares is exported by the .so and bound."""
    hdr = open(os.path.join(ROOT, "include", "pacbot_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(pb_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    L = pb._lib.load()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert declared == set(pb._lib.SIGNATURES), declared ^ set(pb._lib.SIGNATURES)


def test_library_constants(pb):
    L = pb._lib.load()
    assert L.pb_version() == 100
    assert L.pb_sizeof_env_state() == pb.ENV_STATE_DTYPE.itemsize
    assert L.pb_state_bytes_per_env() == 176
    assert pb.OBS_SHAPE == (17, 28, 31) and pb.OBS_FLOATS == 14756


def test_counter_rng_host_mirror(pb):
    """pb_draw_raw is a pure function of (seed, env, ctr); numpy restatement of the same hash."""
    L = pb._lib.load()
    M = (1 << 64) - 1

    def mix(z):
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        return z ^ (z >> 31)

    def draw(seed, gid, ctr):
        key = mix((seed + 0x9E3779B97F4A7C15 * (gid + 1)) & M)
        return mix((key + ctr * 0x9E3779B97F4A7C15 + 0x632BE59BD9B4E019) & M) >> 32

    for seed, gid, ctr in [(0, 0, 0), (1, 2, 3), (12345, 65535, 99), (2**63, 2**20, 2**31)]:
        assert L.pb_draw_raw(seed, gid, ctr) == draw(seed, gid, ctr)
    vals = {L.pb_draw_raw(7, g, c) for g in range(64) for c in range(64)}
    assert len(vals) > 4000  # no obvious collisions / stuck bits


def test_no_cpu_fallback(pb):
    """Without a CUDA device engine creation must fail loudly (no silent CPU path)."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pb.PacbotB200Error):
        pb.BatchedPacmanGym(4)
    with pytest.raises(pb.PacbotB200Error):
        pb.PacmanGym(True, True)
    # and the raw C ABI reports PB_ERR_NO_DEVICE instead of crashing
    L = pb._lib.load()
    h = C.c_void_p()
    assert L.pb_create(C.byref(h), 4, 0, 0, 0, None) == -4
    assert b"no CPU path" in L.pb_last_error()


def test_product_does_not_import_oracle():
    """Neither the oracle nor any other test infrastructure (the MCTS oracle, the host build of
    the device engine under tests/host_engine/) is reachable from the product package, the C-ABI
    header or the root alias module."""
    import re

    # imports, includes and library loads -- a comment naming an oracle file is not a dependency
    banned = re.compile(r"pyoracle|pymcts|libpacbot_oracle|host_engine|from oracle|import oracle|"
                        r"#\s*include[^\n]*oracle")
    files = [os.path.join(ROOT, "pacbot_rl_b200.py"), os.path.join(ROOT, "include", "pacbot_b200.h")]
    for dirpath, _, names in os.walk(os.path.join(ROOT, "pacbot-rl_b200")):
        files += [os.path.join(dirpath, fn) for fn in names
                  if fn.endswith((".py", ".cu", ".cuh", ".h", ".inl"))]
    assert len(files) > 20
    for path in files:
        src = open(path).read()
        hit = banned.search(src)
        assert hit is None, (path, hit.group(0))


def test_create_rejects_empty_batch(pb):
    """Edge case: zero / negative env counts are argument errors, not crashes."""
    L = pb._lib.load()
    h = C.c_void_p()
    assert L.pb_create(C.byref(h), 0, 0, 0, 0, None) == -1
    assert L.pb_create(C.byref(h), -5, 0, 0, 0, None) == -1
    assert L.pb_create(None, 4, 0, 0, 0, None) == -1
    with pytest.raises(ValueError):
        pb._lib.check(-1)
