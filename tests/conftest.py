import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "maze.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure); built on demand with gcc."""
    from oracle import pyoracle

    pyoracle.build()
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def pb():
    """The product package; builds the CUDA extension in-tree if it is stale."""
    import __graft_entry__ as ge

    ge.build_extension()
    import pacbot_rl_b200

    return pacbot_rl_b200


# ---- state comparison helper (oracle po_flat_state  vs  product pb_env_state) ----------------
COMMON_FIELDS = [
    "curr_ticks", "rng_ctr", "pellets", "level_steps", "curr_score", "num_pellets", "last_score",
    "update_period", "mode", "paused", "pause_on_update", "mode_steps", "curr_level",
    "curr_lives", "ghost_combo", "fruit_steps", "last_action", "ticks_per_step",
    "pac_row", "pac_col", "pac_dir", "g_row", "g_col", "g_dir", "g_next_row", "g_next_col",
    "g_next_dir", "g_fright", "g_trapped", "g_spawning", "g_eaten",
]


def oracle_cfg_flags(st):
    return (st["random_start"].astype(np.uint8) | (st["random_ticks"].astype(np.uint8) << 1)
            | (st["randomize_ghosts"].astype(np.uint8) << 2)
            | (st["remove_super_pellets"].astype(np.uint8) << 3))


def assert_states_equal(oracle_state, gpu_state, context=""):
    for f in COMMON_FIELDS:
        a, b = oracle_state[f], gpu_state[f]
        if not np.array_equal(a, b):
            bad = np.argwhere(np.asarray(a != b).reshape(len(a), -1).any(axis=1)).ravel()
            i = int(bad[0])
            raise AssertionError(
                f"{context}: field {f!r} differs for {len(bad)} envs; first env {i}: "
                f"oracle={a[i]!r} gpu={b[i]!r}"
            )
    assert np.array_equal(oracle_cfg_flags(oracle_state), gpu_state["cfg_flags"]), context


def oracle_to_gpu_state(pbmod, st):
    """Convert oracle flat states to the product's pb_env_state records."""
    out = np.zeros(len(st), pbmod.ENV_STATE_DTYPE)
    for f in COMMON_FIELDS:
        out[f] = st[f]
    out["cfg_flags"] = oracle_cfg_flags(st)
    return out


def assert_obs_equal(oracle_obs, gpu_obs, context="", state=None):
    """Bit-exact float32 comparison with a compact report (a bare `assert np.array_equal(...)`
    on multi-hundred-MB arrays makes pytest spend minutes building the failure repr)."""
    a = np.ascontiguousarray(oracle_obs).view(np.uint32)
    b = np.ascontiguousarray(gpu_obs).view(np.uint32)
    if a.shape == b.shape and np.array_equal(a, b):
        return
    assert a.shape == b.shape, f"{context}: obs shape {a.shape} vs {b.shape}"
    bad = np.argwhere(a != b)
    lines = [f"{context}: observation differs in {len(bad)} cells of {len(np.unique(bad[:, 0]))} envs"]
    for e, ch, x, y in bad[:12]:
        lines.append(f"  env {e} ch {ch} cell (row {30 - y}, col {x}): oracle "
                     f"{oracle_obs[e, ch, x, y]!r} gpu {gpu_obs[e, ch, x, y]!r}")
    if state is not None:
        e = int(bad[0][0])
        s = state[e]
        lines.append("  state of first env: " + ", ".join(
            f"{k}={s[k].tolist() if hasattr(s[k], 'tolist') else s[k]}" for k in
            ("pac_row", "pac_col", "g_row", "g_col", "g_fright", "ghost_combo", "fruit_steps",
             "num_pellets", "mode", "mode_steps")))
    raise AssertionError("\n".join(lines))


@pytest.fixture(autouse=True)
def _watchdog():
    """PB_TEST_WATCHDOG=<seconds>: dump the Python stack and exit if a single test exceeds it (a
    hung CUDA call cannot be interrupted by pytest-timeout's signal method)."""
    secs = int(os.environ.get("PB_TEST_WATCHDOG", "0"))
    if secs > 0:
        import faulthandler

        faulthandler.dump_traceback_later(secs, exit=True)
        yield
        faulthandler.cancel_dump_traceback_later()
    else:
        yield
