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


def test_c_rollout_random_step_matches_sequential(pb):
    """pb_rollout_random_step: the same pipeline as DoubleBufferedRandomRollout, one C call per
    step on engine-owned streams."""
    n, steps = 5000, 33
    a, b = make(pb, n, seed=77), make(pb, n, seed=77)
    obs_a = torch.empty((n,) + pb.OBS_SHAPE, device="cuda:0")
    obs_b = [torch.empty((n,) + pb.OBS_SHAPE, device="cuda:0") for _ in range(2)]
    rew = torch.empty(n, dtype=torch.int32, device="cuda:0")
    done = torch.empty(n, dtype=torch.bool, device="cuda:0")
    mask = torch.empty((n, 5), dtype=torch.bool, device="cuda:0")
    acts = torch.empty(n, dtype=torch.uint8, device="cuda:0")
    for t in range(steps):
        exp_actions = a.sample_actions(call_idx=t)
        r_a, d_a = a.step(exp_actions)
        a.obs(obs_a)
        b.rollout_random_step(obs_b[t & 1], call_idx=t, actions_out=acts, mask_out=mask, reward_out=rew,
                              done_out=done)
        if t % 8 == 7 or t == steps - 1:
            b.rollout_join()
            torch.cuda.synchronize()
            assert torch.equal(exp_actions, acts) and torch.equal(r_a, rew) and torch.equal(d_a, done), f"step {t}"
            assert torch.equal(a.action_mask(), mask)
            assert a.get_state().tobytes() == b.get_state().tobytes(), f"state differs at step {t}"
            assert torch.equal(obs_a, obs_b[t & 1]), f"observations differ at step {t}"
    s, e = b.rollout_streams()
    assert s.cuda_stream != e.cuda_stream


@pytest.mark.parametrize("pinned", [False, True], ids=["pageable-staged", "pinned-zero-copy"])
def test_c_rollout_host_step_matches_step_host(pb, pinned):
    """pb_rollout_host_step with pageable and with pinned host buffers (both go through the engine's
    device staging buffers: actions H2D on the step stream, reward / done / mask D2H on the copy
    stream while the encoder runs, one copy-stream synchronisation per step) against pb_step_host."""
    n, steps = 3000, 21
    a, b = make(pb, n, seed=5), make(pb, n, seed=5)
    rng = np.random.default_rng(2)
    obs_a = torch.empty((n,) + pb.OBS_SHAPE, device="cuda:0")
    obs_b = [torch.empty((n,) + pb.OBS_SHAPE, device="cuda:0") for _ in range(2)]
    if pinned:
        keep = [torch.empty(n, dtype=torch.int32).pin_memory(), torch.empty(n, dtype=torch.bool).pin_memory(),
                torch.empty((n, 5), dtype=torch.bool).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()]
        outs = tuple(t.numpy() for t in keep[:3])
        act_buf = keep[3].numpy()
    else:
        outs = (np.empty(n, np.int32), np.empty(n, np.bool_), np.empty((n, 5), np.bool_))
        act_buf = np.empty(n, np.uint8)
    for t in range(steps):
        act_buf[:] = rng.integers(0, 5, n, dtype=np.uint8)
        r_a, d_a, m_a = a.step_host(act_buf.copy(), obs_out=obs_a)
        b.rollout_host_step(act_buf, obs_b[t & 1], outs)
        assert np.array_equal(r_a, outs[0]) and np.array_equal(d_a, outs[1]) and np.array_equal(m_a, outs[2]), f"step {t}"
        if t % 5 == 4 or t == steps - 1:
            b.rollout_join()
            torch.cuda.synchronize()
            assert torch.equal(obs_a, obs_b[t & 1]), f"observations differ at step {t}"
            assert a.get_state().tobytes() == b.get_state().tobytes()
    before = b.get_state()
    act_buf[:] = rng.integers(0, 5, n, dtype=np.uint8)
    act_buf[17] = 9
    with pytest.raises(ValueError, match="Invalid action"):
        b.rollout_host_step(act_buf, obs_b[0], outs)
    b.rollout_join()
    torch.cuda.synchronize()
    assert b.get_state()[17:18].tobytes() == before[17:18].tobytes()   # env.rs:83-88: env untouched
    act_buf[17] = 0
    b.rollout_host_step(act_buf, obs_b[0], outs)                        # and the engine keeps working
    b.rollout_join()


def test_device_timeline_records_every_launch(pb):
    """pb_trace_*: k_step / k_encode_obs launches record %globaltimer begin <= ready <= end, marks
    bracket them on the caller's stream, records come back in launch order, tracing is off again
    afterwards (the timed loops of bench.py run untraced)."""
    n = 4096
    env = make(pb, n, seed=3)
    obs = torch.empty((2, n) + pb.OBS_SHAPE, device="cuda:0")
    env.rollout_random_step(obs[0], call_idx=0)
    env.rollout_join()
    torch.cuda.synchronize()
    env.trace_begin(64)
    env.trace_mark()
    for t in range(5):
        env.rollout_random_step(obs[t & 1], call_idx=1 + t)
    env.rollout_join()
    env.trace_mark()
    rec = env.trace_end()
    assert rec["kind"].tolist() == [0] + [1, 2] * 5 + [0]
    assert (rec["begin_ns"] <= rec["ready_ns"]).all() and (rec["ready_ns"] <= rec["end_ns"]).all()
    steps, encs, marks = rec[rec["kind"] == 1], rec[rec["kind"] == 2], rec[rec["kind"] == 0]
    assert marks[0]["end_ns"] <= steps[0]["begin_ns"] and encs[-1]["end_ns"] <= marks[1]["begin_ns"]
    # an encoder reads the state its step wrote; the next step overlaps it
    assert (steps["end_ns"] <= encs["ready_ns"]).all()
    # consecutive encoders run on two streams: the next one starts on the SMs the previous one has
    # left, so they may overlap by a tail, never by more (one CTA per SM holds all its shared memory)
    assert (encs["end_ns"][:-1] <= encs["end_ns"][1:]).all()
    assert (encs["ready_ns"][1:] + 50_000 >= encs["end_ns"][:-1]).all()
    assert (encs["end_ns"] - encs["ready_ns"]).min() > 10_000      # 4 096 obs = 242 MB: > 10 us
    env.rollout_random_step(obs[0], call_idx=9)                    # untraced again
    env.rollout_join()
    torch.cuda.synchronize()
    env.trace_begin(4)
    assert len(env.trace_end()) == 0


@pytest.mark.parametrize("n", [96, 700, 5000])
def test_consecutive_encoders_into_one_buffer_stay_ordered(pb, n):
    """Consecutive rollout encoders run on two streams and may overlap -- unless they write
    overlapping output ranges: the SAME observation buffer handed to every step must end up
    holding the last step's observations (small n: the two encoders would otherwise run fully
    concurrently on disjoint SMs), also when the second range only partly overlaps the first."""
    a, b = make(pb, n, seed=21), make(pb, n, seed=21)
    obs_a = torch.empty((n,) + pb.OBS_SHAPE, device="cuda:0")
    big = torch.zeros((n + 8,) + pb.OBS_SHAPE, device="cuda:0")
    for t in range(12):
        a.step_random(call_idx=t)
        a.obs(obs_a)
        # every step writes the same rows, except the last one, shifted by 8 envs (partial overlap)
        dst = big[8:8 + n] if t == 11 else big[:n]
        b.rollout_random_step(dst, call_idx=t)
        if t in (5, 10):
            b.rollout_join()
            torch.cuda.synchronize()
            assert torch.equal(big[:n], obs_a), f"step {t}"
    b.rollout_join()
    torch.cuda.synchronize()
    assert torch.equal(big[8:8 + n], obs_a)
    assert a.get_state().tobytes() == b.get_state().tobytes()


def test_split_host_step_matches_step_host(pb):
    """pb_step_host_begin / _end with the caller's own chunked encodes in between == pb_step_host;
    an invalid action is reported by _end and leaves that env untouched; a second _begin before
    _end is refused."""
    n, chunk = 5000, 2048
    a, b = make(pb, n, seed=8), make(pb, n, seed=8)
    rng = np.random.default_rng(4)
    obs_a = torch.empty((n,) + pb.OBS_SHAPE, device="cuda:0")
    obs_b = torch.empty((n,) + pb.OBS_SHAPE, device="cuda:0")
    outs = tuple(t.pin_memory().numpy() for t in (torch.empty(n, dtype=torch.int32), torch.empty(n, dtype=torch.bool),
                                                  torch.empty((n, 5), dtype=torch.bool)))
    for t in range(15):
        acts = rng.integers(0, 5, n, dtype=np.uint8)
        r_a, d_a, m_a = a.step_host(acts, obs_out=obs_a)
        b.step_host(acts, out=outs, begin_only=True)
        for first in range(0, n, chunk):
            cnt = min(chunk, n - first)
            b.obs_range(first, cnt, obs_b[first:first + cnt])
        r_b, d_b, m_b = b.step_host_end()
        assert r_b is outs[0]
        assert np.array_equal(r_a, r_b) and np.array_equal(d_a, d_b) and np.array_equal(m_a, m_b), t
    torch.cuda.synchronize()
    assert torch.equal(obs_a, obs_b) and a.get_state().tobytes() == b.get_state().tobytes()
    acts = rng.integers(0, 5, n, dtype=np.uint8)
    b.step_host(acts, out=outs, begin_only=True)
    with pytest.raises(ValueError, match="has not been ended"):
        b.step_host(acts, out=outs, begin_only=True)
    b.step_host_end()
    before = b.get_state()
    acts[77] = 9
    b.step_host(acts, out=outs, begin_only=True)
    with pytest.raises(ValueError, match="Invalid action"):
        b.step_host_end()
    after = b.get_state()
    assert after[77].tobytes() == before[77].tobytes()       # env.rs:83-88: env untouched


@pytest.mark.parametrize("n,chunk", [(5000, 2048), (4096, 1024), (777, 100)])
def test_chunked_rollout_matches_sequential(pb, n, chunk):
    """pb_rollout_chunk_step: the shard is walked chunk by chunk, every chunk's k_step in place on
    the step stream beside the encoders of the other chunks (two encode streams).  Per env the
    sequence of launches is unchanged, so state / reward / done / mask / actions / observations are
    those of step_random + obs, sweep after sweep -- also when consecutive encoders share a buffer."""
    sweeps = 24
    a, b = make(pb, n, seed=61), make(pb, n, seed=61)
    obs_a = torch.empty((n,) + pb.OBS_SHAPE, device="cuda:0")
    obs_b = [torch.empty((n,) + pb.OBS_SHAPE, device="cuda:0") for _ in range(2)]
    rew = torch.empty(n, dtype=torch.int32, device="cuda:0")
    done = torch.empty(n, dtype=torch.bool, device="cuda:0")
    mask = torch.empty((n, 5), dtype=torch.bool, device="cuda:0")
    acts = torch.empty(n, dtype=torch.uint8, device="cuda:0")
    chunks = [(f, min(chunk, n - f)) for f in range(0, n, chunk)]
    for t in range(sweeps):
        exp_actions = a.sample_actions(call_idx=t)
        r_a, d_a = a.step(exp_actions)
        a.obs(obs_a)
        for first, cnt in chunks:
            b.rollout_chunk_step(first, cnt, obs_b[t & 1][first:first + cnt], call_idx=t, actions_out=acts,
                                 mask_out=mask, reward_out=rew, done_out=done)
        if t % 7 == 6 or t == sweeps - 1:
            b.rollout_join()
            torch.cuda.synchronize()
            assert torch.equal(exp_actions, acts) and torch.equal(r_a, rew) and torch.equal(d_a, done), f"sweep {t}"
            assert torch.equal(a.action_mask(), mask)
            assert a.get_state().tobytes() == b.get_state().tobytes(), f"state differs at sweep {t}"
            assert torch.equal(obs_a, obs_b[t & 1]), f"observations differ at sweep {t}"
    # every chunk into the SAME small buffer: the encoders must stay ordered, the last one wins
    small = torch.empty((chunk,) + pb.OBS_SHAPE, device="cuda:0")
    a.step_random(call_idx=sweeps)
    a.obs(obs_a)
    for first, cnt in chunks:
        b.rollout_chunk_step(first, cnt, small[:cnt], call_idx=sweeps)
    b.rollout_join()
    torch.cuda.synchronize()
    first, cnt = chunks[-1]
    assert torch.equal(small[:cnt], obs_a[first:first + cnt])
    if len(chunks) > 1 and cnt < chunk:
        f2 = chunks[-2][0]
        assert torch.equal(small[cnt:], obs_a[f2 + cnt:f2 + chunk])
    # with explicit device actions, and mixed with the double-buffered form on the same engine
    for t in range(sweeps + 1, sweeps + 7):
        exp_actions = a.sample_actions(call_idx=t)
        r_a, d_a = a.step(exp_actions)
        a.obs(obs_a)
        if t % 3 == 0:
            b.rollout_random_step(obs_b[t & 1], call_idx=t, reward_out=rew, done_out=done, mask_out=mask)
        else:
            for first, cnt in chunks:
                b.rollout_chunk_step(first, cnt, obs_b[t & 1][first:first + cnt], actions=exp_actions,
                                     mask_out=mask, reward_out=rew, done_out=done)
    b.rollout_join()
    torch.cuda.synchronize()
    assert torch.equal(r_a, rew) and torch.equal(d_a, done) and torch.equal(a.action_mask(), mask)
    assert a.get_state().tobytes() == b.get_state().tobytes()
    assert torch.equal(obs_a, obs_b[t & 1])


def test_chunked_host_rollout_matches_step_host(pb):
    """pb_rollout_chunk_host_step / pb_rollout_host_sync: pinned host actions in, reward / done /
    mask back per chunk on the copy stream, one host synchronisation per sweep == pb_step_host;
    an invalid action is reported by the sync and leaves that env untouched."""
    n, chunk, sweeps = 5000, 1536, 18
    a, b = make(pb, n, seed=19), make(pb, n, seed=19)
    rng = np.random.default_rng(6)
    obs_a = torch.empty((n,) + pb.OBS_SHAPE, device="cuda:0")
    obs_b = torch.empty((n,) + pb.OBS_SHAPE, device="cuda:0")
    keep = [torch.empty(n, dtype=torch.int32).pin_memory(), torch.empty(n, dtype=torch.bool).pin_memory(),
            torch.empty((n, 5), dtype=torch.bool).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()]
    outs = tuple(t.numpy() for t in keep[:3])
    act_buf = keep[3].numpy()
    chunks = [(f, min(chunk, n - f)) for f in range(0, n, chunk)]
    for t in range(sweeps):
        act_buf[:] = rng.integers(0, 5, n, dtype=np.uint8)
        r_a, d_a, m_a = a.step_host(act_buf.copy(), obs_out=obs_a)
        for first, cnt in chunks:
            b.rollout_chunk_host_step(first, cnt, act_buf, obs_b[first:first + cnt], outs)
        b.rollout_host_sync()
        assert np.array_equal(r_a, outs[0]) and np.array_equal(d_a, outs[1]) and np.array_equal(m_a, outs[2]), f"sweep {t}"
        if t % 5 == 4 or t == sweeps - 1:
            b.rollout_join()
            torch.cuda.synchronize()
            assert torch.equal(obs_a, obs_b), f"observations differ at sweep {t}"
            assert a.get_state().tobytes() == b.get_state().tobytes()
    before = b.get_state()
    act_buf[:] = rng.integers(0, 5, n, dtype=np.uint8)
    act_buf[1700] = 7
    for first, cnt in chunks:
        b.rollout_chunk_host_step(first, cnt, act_buf, obs_b[first:first + cnt], outs)
    with pytest.raises(ValueError, match="Invalid action"):
        b.rollout_host_sync()
    b.rollout_join()
    torch.cuda.synchronize()
    assert b.get_state()[1700:1701].tobytes() == before[1700:1701].tobytes()   # env.rs:83-88
    act_buf[1700] = 0
    for first, cnt in chunks:
        b.rollout_chunk_host_step(first, cnt, act_buf, obs_b[first:first + cnt], outs)
    b.rollout_host_sync()                                                      # and it keeps working
    b.rollout_join()
