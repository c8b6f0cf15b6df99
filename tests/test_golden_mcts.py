"""tests/golden/mcts_v1.npz (tools/make_golden_mcts.py): the MCTS oracle must keep reproducing
its committed fixture (CPU), and the CUDA path must reproduce it bit for bit WITHOUT the oracle
(GPU).  Evaluator / noise: tests/mcts_common.py."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_golden_mcts as G  # noqa: E402
from mcts_common import NoiseSource, np_evaluator  # noqa: E402


@pytest.fixture(scope="module")
def golden_mcts():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "mcts_v1.npz")))


def assert_same(got, want, label):
    assert set(got) == set(want)
    for k in sorted(want):
        assert np.array_equal(np.asarray(got[k]), want[k]), f"{label}: field {k} differs"


def test_oracle_reproduces_golden_mcts(oracle, golden_mcts):
    assert_same(G.run_oracle(), golden_mcts, "oracle vs fixture")


class _GpuTreeGetters:
    """Adapts BatchedMCTS (batched getters) to the per-index getters snapshot() expects."""

    def __init__(self, m):
        self.c = {k: getattr(m, k)().cpu().numpy() for k in (
            "root_visits", "node_count", "max_depth", "best_action", "value", "action_distribution",
            "policy_prior", "action_values")}

    def __getattr__(self, name):
        return lambda i: self.c[name][i]


@pytest.mark.gpu
def test_gpu_reproduces_golden_mcts(pb, golden_mcts):
    import torch

    noise = NoiseSource(G.N)

    def evaluator(obs, mask):
        v, p = np_evaluator(obs.cpu().numpy(), mask.cpu().numpy())
        return torch.from_numpy(v).to(obs.device), torch.from_numpy(p).to(obs.device)

    env = pb.BatchedPacmanGym(G.N, G.RANDOM_START.astype(bool), False, device="cuda:0", seed=G.SEED,
                              env_id_base=G.BASE)
    m = pb.mcts.BatchedMCTS(env, evaluator, G.PHASE1 + G.PHASE2 + 2, noise_fn=noise.torch_fn)
    m.reset()
    got = {}
    for _ in range(G.PHASE1):
        m.search_iteration()
    got.update({f"p1_{k}": v for k, v in G.snapshot(_GpuTreeGetters(m), G.N).items()})
    reward, done, _ = m.take_action(torch.from_numpy(got["p1_best_action"].astype(np.uint8)))
    got["move_reward"], got["move_done"] = reward.cpu().numpy(), done.cpu().numpy()
    for _ in range(G.PHASE2):
        m.search_iteration()
    got.update({f"p2_{k}": v for k, v in G.snapshot(_GpuTreeGetters(m), G.N).items()})
    m.check()
    assert_same(got, golden_mcts, "CUDA vs fixture")
