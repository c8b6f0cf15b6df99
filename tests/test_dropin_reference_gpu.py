"""The drop-in claim exercised on the GPU: the reference's UNMODIFIED src/replay_buffer.py
(`ReplayBuffer`, `reset_env`, /root/reference/src/replay_buffer.py:44-143) runs over this repo's
`pacbot_rl_b200.PacmanGym` (one 1-env CUDA engine per object) as its `pacbot_rs.PacmanGym`, and
over the oracle-backed shim, on ONE draw tape; everything that lands in the reference's replay
buffer -- observation bits, action, reward, done, next action mask of every transition, and the
number of random draws each env consumed -- must be identical.

The reference's file is imported from /root/reference/src where that exists and otherwise from the
staged, git-ignored copy baseline/_ref/src (tools/stage_reference.py, run by build()): the GPU box
has no /root/reference.  Both runs are subprocesses of tests/ref_replay_driver.py."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

N_ENVS, STEPS, TAPE_LEN = 6, 120, 4093
HAVE_REF = any(os.path.isdir(d) for d in ("/root/reference/src", os.path.join(ROOT, "baseline", "_ref", "src")))


def _run(backend, tape_path, out_path):
    res = subprocess.run(
        [sys.executable, os.path.join(ROOT, "tests", "ref_replay_driver.py"), tape_path, out_path,
         str(N_ENVS), str(STEPS), backend], capture_output=True, text=True, timeout=900,
        env={**os.environ, "OMP_NUM_THREADS": "1"})
    assert res.returncode == 0, f"{backend}: " + res.stderr[-3000:]
    with np.load(out_path) as z:
        return {k: z[k] for k in z.files}


@pytest.mark.skipif(not HAVE_REF, reason="reference sources neither at /root/reference nor staged in baseline/_ref")
def test_reference_replay_buffer_runs_unmodified_on_the_b200_engine(pb, oracle, tmp_path):
    rng = np.random.default_rng(20261021)
    tape = rng.integers(0, 2**32, size=(N_ENVS, TAPE_LEN), dtype=np.uint32)
    tape_path = str(tmp_path / "tape.npy")
    np.save(tape_path, tape)
    (tmp_path / "o").mkdir()
    (tmp_path / "p").mkdir()
    ref = _run("oracle", tape_path, str(tmp_path / "o" / "out.npz"))
    got = _run("product", tape_path, str(tmp_path / "p" / "out.npz"))
    assert ref["draws_used"].max() < TAPE_LEN
    for key in ("action", "reward", "done", "next_mask", "draws_used"):
        assert np.array_equal(ref[key], got[key]), key
    bad = np.argwhere((ref["obs"] != got["obs"]).reshape(STEPS, N_ENVS, -1).any(-1))
    assert len(bad) == 0, f"observation bits differ for (step, env) {bad[:8].tolist()}"
    # the scenario exercises terminal transitions, the old-policy phase and big rewards
    assert int(ref["done"].sum()) >= 2
    assert int(ref["reward"].max()) >= 10 and int(ref["reward"].min()) <= -200
    assert got["final_score"].shape == (N_ENVS,)
