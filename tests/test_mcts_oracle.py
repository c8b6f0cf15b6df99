"""Known-answer tests that pin the MCTS oracle (oracle/mcts_oracle.c) to what the reference's
source fixes (pacbot_rs/src/mcts.rs, alphazero.rs): tie-breaking of max_by_key, the return
algebra of backprop_trajectory, visit bookkeeping, re-rooting, the collector's value recursion.
The reference ships no MCTS tests or vectors, so these are derived from its code, not recorded
from it (parity unpinned at the engine boundary)."""
import numpy as np
import pytest

from mcts_common import NoiseSource, np_evaluator, uniform_evaluator


@pytest.fixture(scope="module")
def pymcts(oracle):
    from oracle import pymcts

    pymcts.lib()
    return pymcts


def uniform_noise(tree, k):
    return np.full(k, 1.0 / k, np.float32)


def test_counter_rng_restatement_matches_the_product(pymcts, pb):
    L = pb._lib.load()
    for seed, gid, ctr in [(0, 0, 0), (1, 2, 3), (2**63 + 5, 10**9, 2**32 - 1)]:
        assert pymcts.lib().pm_draw_raw(seed, gid, ctr) == L.pb_draw_raw(seed, gid, ctr)
    for seed, epoch in [(0, 0), (7, 1), (2**64 - 1, 123456789)]:
        assert pymcts.lib().pm_sim_seed(seed, epoch) == L.pb_sim_seed(seed, epoch)
    assert pymcts.lib().pm_sim_seed(7, 0) != pymcts.lib().pm_sim_seed(7, 1)


def test_first_iteration_takes_the_last_maximal_action(pymcts):
    # uniform priors, no children: every PUCT key is equal, Iterator::max_by_key keeps the LAST
    # maximum (mcts.rs:55-66), i.e. the highest-numbered valid action
    m = pymcts.OracleMCTS(4, [1, 1, 0, 0], 0, 3, 500, uniform_evaluator, uniform_noise)
    m.reset()
    for i in range(4):
        assert m.root_visits(i) == 1 and m.node_count(i) == 1 and m.max_depth(i) == 0
        assert m.value(i) == 0.0
    masks = [m.root_mask(i) for i in range(4)]
    m.search_iteration()
    for i in range(4):
        last_valid = int(np.nonzero(masks[i])[0][-1])
        dist = m.action_distribution(i)
        assert m.root_visits(i) == 2 and m.node_count(i) == 2 and m.max_depth(i) == 1
        assert dist[last_valid] == 1.0 and dist.sum() == 1.0
        assert m.best_action(i) == last_valid
    assert m.epoch == 1


def test_backprop_return_algebra(pymcts):
    # value 0 leaf: child.total = reward, root.total = 0 + 0.99 * reward (mcts.rs:365-380)
    a = pymcts.OracleMCTS(1, 0, 0, 9, 40, uniform_evaluator, uniform_noise)
    b = pymcts.OracleMCTS(1, 0, 0, 9, 40, uniform_evaluator, uniform_noise)
    a.reset()
    b.reset()
    a.search_iteration()
    action = a.best_action(0)
    reward, done = b.take_action([action])   # fresh game: no frightened ghosts, no draws
    assert not done[0]
    r = np.float32(reward[0])
    assert a.action_values(0)[action] == r
    assert a.value(0) == (np.float32(0.0) + np.float32(0.99) * r) / np.float32(2.0)
    # absent children report the root's own value (mcts.rs:89-92)
    other = [x for x in range(5) if x != action][0]
    assert a.action_values(0)[other] == a.value(0)


def test_visit_bookkeeping_and_reroot(pymcts):
    n = 6
    noise = NoiseSource(n)
    m = pymcts.OracleMCTS(n, [1, 0] * 3, 0, 21, 0, np_evaluator, noise.oracle)
    m.reset()
    iters = 40
    for _ in range(iters):
        m.search_iteration()
    for i in range(n):
        assert m.root_visits(i) == 1 + iters
        assert 2 <= m.node_count(i) <= 1 + iters
        assert abs(float(m.action_distribution(i).sum()) - 1.0) < 1e-6
        assert (m.action_distribution(i)[~m.root_mask(i)] == 0).all()
    best = [m.best_action(i) for i in range(n)]
    visits_before = [m.action_distribution(i)[best[i]] * iters for i in range(n)]
    prior_before = [m.policy_prior(i).copy() for i in range(n)]
    reward, done = m.take_action(best)
    for i in range(n):
        if done[i]:
            continue
        # the most visited child became the root: its visit count carries over (mcts.rs:160-163)
        assert m.root_visits(i) == round(float(visits_before[i]))
        assert m.node_count(i) <= iters
        # and root noise was mixed into ITS priors, not the old root's
        assert not np.array_equal(m.policy_prior(i), prior_before[i])
    # one mask-selected tree acting on a never-visited action gets a fresh root (mcts.rs:164-166)
    i = int(np.nonzero(~done)[0][0])
    unvisited = [a for a in range(5) if m.root_mask(i)[a] and m.action_distribution(i)[a] == 0]
    if unvisited:
        mask = np.zeros(n, np.uint8)
        mask[i] = 1
        acts = np.zeros(n, np.uint8)
        acts[i] = unvisited[0]
        _, d = m.take_action(acts, mask)
        if not d[i]:
            assert m.root_visits(i) == 1 and m.node_count(i) == 1


def test_sim_draws_differ_from_real_draws(pymcts):
    # frightened ghosts move at random: a simulated step must not reuse the real env's draws
    assert pymcts.lib().pm_draw_raw(pymcts.lib().pm_sim_seed(5, 0), 1, 0) != pymcts.lib().pm_draw_raw(5, 1, 0)


def test_collector_value_recursion_and_shapes(pymcts):
    noise = NoiseSource(4)
    c = pymcts.OracleCollector(12, 25, 0.99, 4, 77, 1000, np_evaluator, noise.oracle)
    total = 0
    for _ in range(3):
        items = c.generate_experience()
        assert len(items) >= 1
        total += len(items)
        assert items["obs"].shape[1:] == (17, 28, 31)
        # every recorded state is a live game position with at least 2 legal actions
        assert (items["action_mask"].sum(1) >= 2).all() and (items["action_mask"][:, 0] == 1).all()
        # visit distributions live on the legal actions and sum to one
        assert (items["action_distribution"][items["action_mask"] == 0] == 0).all()
        assert np.allclose(items["action_distribution"].sum(1), 1.0, atol=1e-6)
        assert np.isfinite(items["value"]).all()
    assert total <= 3 * 4 * 25


def test_collector_truncation_bootstraps_from_the_root_value(pymcts):
    # max_episode_length 1: every action ends an episode; its value is reward + 0.99 * value of
    # the NEW root (alphazero.rs:180-184) unless the game ended.  Rebuild the expected items from
    # the MCTS primitives (same seed, ids, evaluator and noise => same trees).
    tree_size, P, seed, base = 8, 3, 5, 40
    noise_c, noise_m = NoiseSource(P), NoiseSource(P)
    c = pymcts.OracleCollector(tree_size, 1, 0.99, P, seed, base, np_evaluator, noise_c.oracle)
    m = pymcts.OracleMCTS(P, 1, 0, seed, base, np_evaluator, noise_m.oracle)   # new(true, false)
    m.reset()
    for _ in range(tree_size - 1):          # num_search_iters_remaining = tree_size - 1
        m.search_iteration()
    obs = [m.root_obs(i) for i in range(P)]
    masks = [m.root_mask(i) for i in range(P)]
    dists = [m.action_distribution(i) for i in range(P)]
    best = [m.best_action(i) for i in range(P)]
    reward, done = m.take_action(best)
    expected = [np.float32(reward[i]) if done[i]
                else np.float32(reward[i]) + np.float32(0.99) * m.value(i) for i in range(P)]
    items = c.generate_experience()
    assert len(items) == P                  # all trees act in the same step, env order
    for i in range(P):
        assert np.array_equal(items["obs"][i].view(np.uint32), obs[i].view(np.uint32))
        assert np.array_equal(items["action_mask"][i].astype(bool), masks[i])
        assert np.array_equal(items["action_distribution"][i].view(np.uint32), dists[i].view(np.uint32))
        assert items["value"][i].view(np.uint32) == np.float32(expected[i]).view(np.uint32)


def test_collector_discounted_returns_within_an_episode(pymcts):
    # end_episode (alphazero.rs:163-170): value[k] = reward[k] + 0.99 * value[k+1], last one
    # bootstrapped if truncated.  With max_episode_length 3 rebuild one episode by hand.
    tree_size, L, seed, base = 6, 3, 9, 7
    noise_c, noise_m = NoiseSource(1), NoiseSource(1)
    c = pymcts.OracleCollector(tree_size, L, 0.99, 1, seed, base, uniform_evaluator, noise_c.oracle)
    m = pymcts.OracleMCTS(1, 1, 0, seed, base, uniform_evaluator, noise_m.oracle)
    m.reset()
    rewards, ended = [], False
    remaining = tree_size - 1
    while len(rewards) < L and not ended:
        for _ in range(remaining):
            m.search_iteration()
        r, d = m.take_action([m.best_action(0)])
        rewards.append(np.float32(r[0]))
        ended = bool(d[0])
        remaining = max(tree_size - m.node_count(0), 0)
    values = list(rewards)
    if not ended:
        values[-1] = values[-1] + np.float32(0.99) * m.value(0)
    nxt = np.float32(0.0)
    for k in range(len(values) - 1, -1, -1):
        values[k] = np.float32(values[k] + np.float32(0.99) * nxt)
        nxt = values[k]
    items = c.generate_experience()
    assert len(items) == len(values)
    assert np.array_equal(items["value"].view(np.uint32), np.array(values, np.float32).view(np.uint32))
