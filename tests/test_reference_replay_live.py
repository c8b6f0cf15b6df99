"""LIVE differential test against the reference's own experience-generation code.

Where /root/reference exists (the build container; skipped elsewhere), the UNMODIFIED
src/replay_buffer.py (`ReplayBuffer`, `reset_env`) is run in a subprocess over a `pacbot_rs` shim
backed by the scalar oracle with an injected draw tape (tests/ref_replay_driver.py).  What lands
in the reference's replay buffer -- per env: observation, action, reward, done, next action mask
of every recorded transition, incl. the unrecorded old-policy phase after each reset -- must equal
`replay_reference.restated_env_loop`, the per-env restatement that the GPU test of the batched
ReplayBuffer uses as its expectation.  SURVEY.md 8(a) rows a13 / a14; integer + exact-float
comparison, tolerance none."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from replay_reference import (MAIN_POLICY_CHANNELS, OLD_POLICY_CHANNELS, bfs_policy,
                              restated_env_loop)

N_ENVS, STEPS, TAPE_LEN = 6, 160, 4093


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree not present")
def test_reference_replay_buffer_live_equals_restated_loop(oracle, tmp_path):
    rng = np.random.default_rng(20261020)
    tape = rng.integers(0, 2**32, size=(N_ENVS, TAPE_LEN), dtype=np.uint32)
    tape_path, out_path = str(tmp_path / "tape.npy"), str(tmp_path / "ref_out.npz")
    np.save(tape_path, tape)
    res = subprocess.run(
        [sys.executable, os.path.join(ROOT, "tests", "ref_replay_driver.py"), tape_path, out_path,
         str(N_ENVS), str(STEPS)], capture_output=True, text=True, timeout=900,
        env={**os.environ, "OMP_NUM_THREADS": "1"})  # tiny tensors: threads only add overhead
    assert res.returncode == 0, res.stderr[-3000:]
    with np.load(out_path) as z:
        ref = {k: z[k] for k in z.files}  # decompress once, not per access
    assert ref["draws_used"].max() < TAPE_LEN  # no tape row wrapped around

    main, old = bfs_policy(MAIN_POLICY_CHANNELS), bfs_policy(OLD_POLICY_CHANNELS)
    episodes = phase1_envs = 0
    for i in range(N_ENVS):
        # replay_buffer.py:75-80: random_start for the first N * proportion envs, random_ticks all
        cfg = np.array([(1 if i < N_ENVS * 0.5 else 0) | 2], np.uint8)
        ob = oracle.OracleBatch(1, cfg, tape[i:i + 1])
        ob.reset()   # reset_env starts with env.reset() (replay_buffer.py:45)
        rec = restated_env_loop(ob, main, old, max_env_steps=100 * STEPS, max_recorded=STEPS)
        assert len(rec) == STEPS, (i, len(rec))
        for t, (obs_bits, a, r, d, nm) in enumerate(rec):
            ctx = f"env {i} transition {t}"
            assert a == ref["action"][t, i] and r == ref["reward"][t, i], ctx
            assert d == ref["done"][t, i], ctx
            assert np.array_equal(nm, ref["next_mask"][t, i]), ctx
            assert np.array_equal(obs_bits[0], ref["obs"][t, i]), ctx
        episodes += int(ref["done"][:, i].sum())
        phase1_envs += int(ob.export_state()["rng_ctr"][0] > 3)
    # the scenario must actually exercise the terminal / re-reset / old-policy paths
    assert episodes >= 2, episodes
    assert int(ref["reward"].max()) >= 10 and int(ref["reward"].min()) <= -200
