"""Generate tests/golden/mcts_v1.npz: the scalar MCTS oracle (oracle/mcts_oracle.c) run on 12 trees
with the deterministic evaluator / root noise of tests/mcts_common.py: 25 search iterations, every
tree plays its most visited action, 15 more iterations.  Recorded: root statistics after each
phase (float fields as raw bits), rewards and done flags of the move.  Pins the ORACLE over time
(CPU suite) and gives the GPU suite a committed fixture to reproduce without the oracle.  It
does NOT pin the oracle to the reference (nothing can: the reference ships no MCTS vectors).
Usage: python tools/make_golden_mcts.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from mcts_common import NoiseSource, np_evaluator  # noqa: E402

N, SEED, BASE, PHASE1, PHASE2 = 12, 31, 4000, 25, 15
RANDOM_START = (np.arange(N) % 2).astype(np.uint8)


def snapshot(tree, n):
    """tree: object with per-tree getters taking an index (OracleMCTS) -> dict of arrays."""
    f = lambda x: np.ascontiguousarray(np.asarray(x, np.float32)).view(np.uint32)  # noqa: E731
    return {
        "root_visits": np.array([tree.root_visits(i) for i in range(n)], np.int64),
        "node_count": np.array([tree.node_count(i) for i in range(n)], np.int64),
        "max_depth": np.array([tree.max_depth(i) for i in range(n)], np.int64),
        "best_action": np.array([tree.best_action(i) for i in range(n)], np.int64),
        "value_bits": f([tree.value(i) for i in range(n)]),
        "dist_bits": f([tree.action_distribution(i) for i in range(n)]),
        "prior_bits": f([tree.policy_prior(i) for i in range(n)]),
        "q_bits": f([tree.action_values(i) for i in range(n)]),
    }


def run_oracle():
    from oracle import pymcts, pyoracle

    pyoracle.build()
    noise = NoiseSource(N)
    om = pymcts.OracleMCTS(N, RANDOM_START, 0, SEED, BASE, np_evaluator, noise.oracle)
    om.reset()
    out = {}
    for _ in range(PHASE1):
        om.search_iteration()
    out.update({f"p1_{k}": v for k, v in snapshot(om, N).items()})
    actions = out["p1_best_action"].astype(np.uint8)
    reward, done = om.take_action(actions)
    out["move_reward"], out["move_done"] = reward, done
    assert not done.any(), "fixture assumes nobody dies on the first move"
    for _ in range(PHASE2):
        om.search_iteration()
    out.update({f"p2_{k}": v for k, v in snapshot(om, N).items()})
    return out


if __name__ == "__main__":
    data = run_oracle()
    path = os.path.join(ROOT, "tests", "golden", "mcts_v1.npz")
    np.savez_compressed(path, **data)
    print(path, {k: v.shape for k, v in data.items()})
