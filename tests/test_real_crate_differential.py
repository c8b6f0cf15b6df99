"""Differential test against the REAL `pacbot_rs` extension (the Rust crate built with maturin), for
whoever has it: it cannot be built in this container (no cargo, engine crate not vendored), so
here the test is skipped and only its harness is exercised (against an oracle-backed stand-in).

What it does when `import pacbot_rs` yields the real extension: plays the same action tapes on
`pacbot_rs.PacmanGym(False, False)` and on the C oracle and compares reward, done, score, lives,
remaining pellets, action mask and the full float32 observation after every step.  The real
crate draws from an unseedable `thread_rng` (env.rs:146), so only RNG-free prefixes can be
compared: fixed start, fixed ticks, and each tape stops being compared at the first step after
which a ghost is frightened (frightened ghosts are the only consumers of randomness inside
`step`).  Any mismatch is an entry of the engine assumption ledger proven
wrong -- the message says which step and which quantity.

    PACBOT_RS_DIFF_TAPES=200 python -m pytest tests/test_real_crate_differential.py -q"""
import os

import numpy as np
import pytest

from replay_reference import MAIN_POLICY_CHANNELS, bfs_action


def real_pacbot_rs():
    try:
        import pacbot_rs
    except Exception:
        return None
    f = getattr(pacbot_rs, "__file__", "") or ""
    return pacbot_rs if f.endswith((".so", ".pyd")) or hasattr(pacbot_rs, "pacbot_rs") else None


def tapes(n_tapes, length, seed=0):
    """Action scripts: seeded mixtures of 'walk towards the nearest pellet' and random moves (the
    choice is made from the ORACLE's observation; the same action goes to both sides)."""
    rng = np.random.default_rng(seed)
    for k in range(n_tapes):
        yield k, rng.random(length), rng.integers(0, 5, length), float(rng.choice([0.0, 0.1, 0.3, 0.6]))


def run_differential(make_gym, oracle, n_tapes=24, length=400):
    """`make_gym()` -> an object with the pyo3 surface (pacbot_rs.pyi:6-19).  Returns the number of
    steps compared; raises AssertionError on the first difference."""
    compared = 0
    for k, coin, wild, p_wild in tapes(n_tapes, length):
        gym = make_gym()
        ob = oracle.OracleBatch(1, np.zeros(1, np.uint8), None)
        gym.reset()
        ob.reset()
        for t in range(length):
            ctx = f"tape {k} step {t}"
            obs, mask = ob.obs()[0], ob.action_mask()[0]
            assert gym.action_mask() == mask.tolist(), f"{ctx}: action mask"
            got = np.asarray(gym.obs_numpy())
            assert got.shape == (17, 28, 31) and got.dtype == np.float32, f"{ctx}: obs shape/dtype"
            if not np.array_equal(got.view(np.uint32), obs.view(np.uint32)):
                ch, x, y = np.argwhere(got != obs)[0]
                raise AssertionError(f"{ctx}: obs[{ch}, col {x}, row {30 - y}] crate {got[ch, x, y]!r} "
                                     f"oracle {obs[ch, x, y]!r}")
            a = int(wild[t]) if coin[t] < p_wild else bfs_action(obs, mask, MAIN_POLICY_CHANNELS)
            r_o, d_o = ob.step(np.array([a], np.uint8))
            r_c, d_c = gym.step(a)
            st = ob.export_state()[0]
            assert (int(r_c), bool(d_c)) == (int(r_o[0]), bool(d_o[0])), f"{ctx} action {a}: reward/done"
            assert int(gym.score()) == int(st["curr_score"]), f"{ctx}: score"
            assert int(gym.lives()) == int(st["curr_lives"]), f"{ctx}: lives"
            assert int(gym.remaining_pellets()) == int(st["num_pellets"]), f"{ctx}: remaining pellets"
            assert bool(gym.is_done()) == bool(d_o[0]), f"{ctx}: is_done"
            compared += 1
            if d_o[0] or (st["g_fright"] > 0).any():
                break   # from here on the crate's thread_rng decides
    return compared


@pytest.mark.skipif(real_pacbot_rs() is None, reason="the real pacbot_rs extension is not installed")
def test_oracle_against_the_real_crate(oracle):
    mod = real_pacbot_rs()
    n = int(os.environ.get("PACBOT_RS_DIFF_TAPES", "24"))
    compared = run_differential(lambda: mod.PacmanGym(False, False), oracle, n_tapes=n)
    assert compared > 10 * n


def test_differential_harness_on_an_oracle_backed_gym(oracle):
    """The harness itself, run against a stand-in with the pyo3 surface built on the oracle: it
    must walk long RNG-free prefixes, and it must catch a planted difference."""

    class Gym:
        def __init__(self, lie_about_score_at=None):
            self._b = oracle.OracleBatch(1, np.zeros(1, np.uint8), None)
            self._steps, self._lie = 0, lie_about_score_at

        def reset(self):
            self._b.reset()

        def step(self, action):
            r, d = self._b.step(np.array([action], np.uint8))
            self._steps += 1
            return int(r[0]), bool(d[0])

        def score(self):
            s = int(self._b.export_state()["curr_score"][0])
            return s + 10 if self._lie is not None and self._steps >= self._lie else s

        def lives(self):
            return int(self._b.export_state()["curr_lives"][0])

        def remaining_pellets(self):
            return int(self._b.export_state()["num_pellets"][0])

        def is_done(self):
            return bool(self._b.is_done()[0])

        def action_mask(self):
            return self._b.action_mask()[0].tolist()

        def obs_numpy(self):
            return self._b.obs()[0].copy()

    compared = run_differential(Gym, oracle, n_tapes=6, length=300)
    assert compared > 150, compared      # the prefixes are long enough to mean something
    with pytest.raises(AssertionError, match="step 7: score"):
        run_differential(lambda: Gym(lie_about_score_at=8), oracle, n_tapes=1, length=50)
