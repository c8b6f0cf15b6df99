"""The two-stream pipelined rollout (pacbot-rl_b200/rollout.py: k_step of one chunk overlaps the
observation encode of another) must be bit-identical to the sequential order, and the ranged step
entry point (pb_step_random_range) must touch only its envs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def make(pb, n, seed=13):
    env = pb.BatchedPacmanGym(n, np.arange(n) < n // 2, True, device="cuda:0", seed=seed, auto_reset=True)
    env.reset()
    return env


@pytest.mark.parametrize("n,num_chunks", [(4096, 2), (5000, 3), (777, 5)])
def test_pipelined_rollout_matches_sequential(pb, n, num_chunks):
    from pacbot_rl_b200.rollout import PipelinedRandomRollout

    steps = 40
    a, b = make(pb, n), make(pb, n)
    obs_a = torch.empty((n,) + pb.OBS_SHAPE, device="cuda:0")
    obs_b = torch.empty((n,) + pb.OBS_SHAPE, device="cuda:0")
    pipe = PipelinedRandomRollout(b, num_chunks=num_chunks)
    assert sum(c for _, c in pipe.chunks) == n
    for t in range(steps):
        r_a, d_a = a.step_random(call_idx=t)
        a.obs(obs_a)
        pipe.step(t, lambda c, first, cnt: obs_b[first:first + cnt])
        if t % 13 == 12 or t == steps - 1:
            pipe.join()
            torch.cuda.synchronize()
            assert torch.equal(r_a, pipe.reward) and torch.equal(d_a, pipe.done), f"step {t}"
            assert a.get_state().tobytes() == b.get_state().tobytes(), f"state differs at step {t}"
            assert torch.equal(obs_a, obs_b), f"observations differ at step {t}"
            assert torch.equal(a.action_mask(), pipe.mask)


def test_ranged_step_touches_only_its_envs(pb):
    n = 1000
    a, b = make(pb, n), make(pb, n)
    before = b.get_state()
    r_a, d_a = a.step_random(call_idx=5)
    rew = torch.full((n,), -7, dtype=torch.int32, device="cuda:0")
    done = torch.zeros(n, dtype=torch.bool, device="cuda:0")
    b.step_random(call_idx=5, reward_out=rew, done_out=done, first=123, count=456)
    torch.cuda.synchronize()
    after, full = b.get_state(), a.get_state()
    assert after[123:579].tobytes() == full[123:579].tobytes()
    assert after[:123].tobytes() == before[:123].tobytes() and after[579:].tobytes() == before[579:].tobytes()
    assert torch.equal(rew[123:579], r_a[123:579]) and (rew[:123] == -7).all() and (rew[579:] == -7).all()
    with pytest.raises(ValueError):
        b.step_random(call_idx=6, first=900, count=200)


@pytest.mark.parametrize("n,chunks", [(4096, None), (3001, [(0, 1500), (1500, 1501)])])
def test_double_buffered_rollout_matches_sequential(pb, n, chunks):
    """pb_step_random_flip: step t+1 writes the engine's second state buffer while the encode of
    step t still reads the first; everything else (state getters, masks, a plain step afterwards)
    keeps seeing "the" state."""
    from pacbot_rl_b200.rollout import DoubleBufferedRandomRollout

    steps = 41   # odd: the rollout ends on the second buffer
    a, b = make(pb, n, seed=29), make(pb, n, seed=29)
    obs_a = torch.empty((n,) + pb.OBS_SHAPE, device="cuda:0")
    obs_b = torch.empty((n,) + pb.OBS_SHAPE, device="cuda:0")
    roll = DoubleBufferedRandomRollout(b, chunks=chunks)
    for t in range(steps):
        r_a, d_a = a.step_random(call_idx=t)
        a.obs(obs_a)
        roll.step(t, lambda c, first, cnt: obs_b[first:first + cnt])
        if t % 10 == 9 or t == steps - 1:
            roll.join()
            torch.cuda.synchronize()
            assert torch.equal(r_a, roll.reward) and torch.equal(d_a, roll.done), f"step {t}"
            assert a.get_state().tobytes() == b.get_state().tobytes(), f"state differs at step {t}"
            assert torch.equal(obs_a, obs_b), f"observations differ at step {t}"
            assert torch.equal(a.action_mask(), roll.mask)
    # the engine is an ordinary engine afterwards: in-place steps, resets, scores agree
    acts = a.sample_actions(call_idx=99)
    assert torch.equal(acts, b.sample_actions(call_idx=99))
    ra, da = a.step(acts)
    rb, db = b.step(acts)
    assert torch.equal(ra, rb) and torch.equal(da, db) and torch.equal(a.score(), b.score())
    m = torch.arange(n, device="cuda:0") % 3 == 0
    a.reset(m)
    b.reset(m)
    assert a.get_state().tobytes() == b.get_state().tobytes()


def test_host_action_rollout_matches_step_host(pb):
    """HostActionRollout (pinned actions in, pinned results out, double-buffered state) == the
    one-call path BatchedPacmanGym.step_host, step for step."""
    from pacbot_rl_b200.rollout import HostActionRollout

    n, steps = 3000, 25
    a, b = make(pb, n, seed=41), make(pb, n, seed=41)
    rng = np.random.default_rng(1)
    obs_a = torch.empty((n,) + pb.OBS_SHAPE, device="cuda:0")
    obs_b = [torch.empty((n,) + pb.OBS_SHAPE, device="cuda:0") for _ in range(2)]
    roll = HostActionRollout(b)
    pinned = torch.empty(n, dtype=torch.uint8).pin_memory()
    for t in range(steps):
        acts = rng.integers(0, 5, n, dtype=np.uint8)
        r_a, d_a, m_a = a.step_host(acts, obs_out=obs_a)
        pinned.numpy()[:] = acts
        r_b, d_b, m_b = roll.step(pinned, obs_b[t & 1])
        assert np.array_equal(r_a, r_b.numpy()) and np.array_equal(d_a, d_b.numpy()), f"step {t}"
        assert np.array_equal(m_a, m_b.numpy()), f"mask at step {t}"
        if t % 6 == 5 or t == steps - 1:
            roll.join()
            torch.cuda.synchronize()
            assert torch.equal(obs_a, obs_b[t & 1]), f"observations differ at step {t}"
            assert a.get_state().tobytes() == b.get_state().tobytes()
