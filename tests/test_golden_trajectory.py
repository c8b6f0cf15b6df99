"""Committed golden trajectory (tests/golden/trajectory_v1.npz, made by
tools/make_golden_trajectory.py from the CPU oracle): the oracle must keep reproducing it (CPU),
and the CUDA path must reproduce it bit for bit through the C ABI (GPU)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.fixture(scope="module")
def golden_traj():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "trajectory_v1.npz")))


def _replay(batch, g, state_of, obs_of, mask_of, step):
    from make_golden_trajectory import obs_digest, state_digest

    batch.reset()
    for t in range(g["actions"].shape[0]):
        assert np.array_equal(mask_of(), g["masks"][t]), f"mask differs at step {t}"
        r, d = step(g["actions"][t])
        assert np.array_equal(r, g["reward"][t]), f"reward differs at step {t}"
        assert np.array_equal(d, g["done"][t]), f"done differs at step {t}"
        assert state_digest(state_of()) == g["state_digest"][t], f"state differs at step {t}"
        assert obs_digest(obs_of()) == g["obs_digest"][t], f"observation differs at step {t}"


def test_oracle_reproduces_golden_trajectory(oracle, golden_traj):
    g = golden_traj
    b = oracle.OracleBatch(len(g["cfg"]), g["cfg"], g["tape"])
    _replay(b, g, b.export_state, b.obs, b.action_mask, lambda a: b.step(a, auto_reset=True))
    assert g["done"].sum() > 50          # the fixture covers many deaths / resets


@pytest.mark.gpu
def test_gpu_reproduces_golden_trajectory(pb, golden_traj):
    import torch

    g = golden_traj
    cfg = g["cfg"]
    env = pb.BatchedPacmanGym(len(cfg), (cfg & 1).astype(bool), (cfg & 2).astype(bool),
                              device="cuda:0", auto_reset=True)
    env.set_draw_tape(torch.from_numpy(g["tape"].view(np.int32)).cuda())

    def step(a):
        r, d = env.step(torch.from_numpy(a).cuda())
        return r.cpu().numpy(), d.cpu().numpy()

    _replay(env, g, env.get_state, lambda: env.obs().cpu().numpy(),
            lambda: env.action_mask().cpu().numpy(), step)
