"""GPU tests of the device-resident DQN loop (replay ring, policies, trainer): mirrors of
/root/reference/src/replay_buffer.py, src/policies.py and src/algorithms/train_dqn.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


class FirstValidPolicy:
    """Deterministic policy: lowest-index valid action other than Stay when there is one."""

    def __call__(self, obs, masks):
        m = masks.clone()
        m[:, 0] = False
        has = m.any(1)
        a = m.float().argmax(1)
        return a * has


def test_replay_ring_matches_oracle_rollout(pb, oracle):
    """The ring stores exactly the transitions an env-by-env reference loop would produce
    (replay_buffer.py:102-139): obs before the action, action, reward, next_obs = obs after the
    step (None/zeros when done), next_action_mask ([False]*5 when done), and the observation that
    starts the next episode comes from the reset env."""
    import torch
    from pacbot_rl_b200.replay_buffer import ReplayBuffer

    n, maxlen, steps = 32, 320, 25
    rng = np.random.default_rng(0)
    tape = rng.integers(0, 2**32, size=(n, 257), dtype=np.uint32)
    buf = ReplayBuffer(maxlen, FirstValidPolicy(), n, 0.5, torch.device("cuda:0"), seed=0)
    cfg = (np.arange(n) < n * 0.5).astype(np.uint8) | 2
    # rebuild the same envs under a draw tape so the oracle can follow
    buf.envs.set_draw_tape(torch.from_numpy(tape.view(np.int32)).cuda())
    ob = oracle.OracleBatch(n, cfg, tape)
    buf.envs.set_state(_to_gpu_state(pb, ob.export_state()))
    ob.reset()
    buf.envs.reset()
    buf.envs.obs(out=buf._obs[0])
    buf.envs.action_mask(out=buf._mask)
    policy = FirstValidPolicy()
    for t in range(steps):
        o_obs = ob.obs()
        o_mask = ob.action_mask()
        a = policy(torch.from_numpy(o_obs), torch.from_numpy(o_mask)).numpy().astype(np.uint8)
        buf.generate_experience_step()
        r, d = ob.step(a, auto_reset=True)
        slot, nxt = t % buf._ring, (t + 1) % buf._ring
        assert np.array_equal(buf._obs[slot].cpu().numpy().view(np.uint32), o_obs.view(np.uint32)), t
        assert np.array_equal(buf._actions[slot].cpu().numpy(), a), t
        assert np.array_equal(buf._rewards[slot].cpu().numpy(), r), t
        assert np.array_equal(buf._done[slot].cpu().numpy(), d), t
        exp_mask = ob.action_mask() & ~d[:, None]
        assert np.array_equal(buf._next_mask[slot].cpu().numpy(), exp_mask), t
        assert np.array_equal(buf._obs[nxt].cpu().numpy().view(np.uint32), ob.obs().view(np.uint32)), t
    assert len(buf) == min(steps, buf._slots) * n == int(buf._valid.sum().item())


def _to_gpu_state(pb, st):
    from conftest import oracle_to_gpu_state

    return oracle_to_gpu_state(pb, st)


def test_sample_batch_collate_semantics(pb):
    import torch
    from pacbot_rl_b200.policies import EpsilonGreedy, MaxQPolicy
    from pacbot_rl_b200.models import QNetV2
    from pacbot_rl_b200.replay_buffer import ReplayBuffer

    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    q = QNetV2(pb.OBS_SHAPE, 5).to(dev)
    buf = ReplayBuffer(10_000, EpsilonGreedy(MaxQPolicy(q), 5, 1.0), 32, 0.5, dev, seed=1)
    buf.fill()
    assert buf.capacity == 313 * 32 and len(buf) == buf.capacity
    for _ in range(40):                      # wrap the ring
        buf.generate_experience_step()
    assert len(buf) == buf.capacity and int(buf._valid.sum().item()) == buf.capacity
    batch = buf.sample_batch(512)
    assert batch.obs.shape == (512, 17, 28, 31) and batch.next_obs.shape == batch.obs.shape
    assert batch.actions.dtype == torch.int64 and batch.rewards.dtype == torch.float32
    # every sampled obs is a real observation: wall channel sums to 580; done rows are zeros
    assert torch.equal(batch.obs[:, 0].sum((1, 2)), torch.full((512,), 580.0, device=dev))
    nz = batch.next_obs.reshape(512, -1).abs().sum(1)
    assert bool(((nz == 0) == batch.done).all())
    assert not bool(batch.next_action_masks[batch.done].any())
    assert bool(batch.next_action_masks[~batch.done][:, 0].all())   # Stay is always valid
    # without replacement, and only recorded transitions
    flat = buf.sample_indices(512).clone()
    assert flat.unique().numel() == 512 and buf.num_valid == buf.capacity
    assert bool(buf._valid.view(-1)[flat].all())
    # the slot holding only the current observation is never sampled
    cur = buf._t % buf._ring
    assert int(buf._valid[cur].sum()) == 0
    assert not bool((flat // buf.num_envs == cur).any())
    # reward sanity: sampled rewards are the stored ones
    assert bool((batch.rewards.abs() <= 3000 + 3000).all())


def test_epsilon_greedy_policy_semantics(pb):
    import torch
    from pacbot_rl_b200.policies import EpsilonGreedy, MaxQPolicy

    dev = torch.device("cuda:0")

    class FakeQ(torch.nn.Module):
        def forward(self, obs):
            return obs[:, :5, 0, 0]

    b = 20000
    g = torch.Generator(device=dev).manual_seed(1)
    obs = torch.randn((b, 17, 28, 31), device=dev, generator=g)
    masks = torch.rand((b, 5), device=dev, generator=g) > 0.4
    masks[:, 0] = True
    greedy_ref = obs[:, :5, 0, 0].masked_fill(~masks, -torch.inf).argmax(1)
    assert torch.equal(MaxQPolicy(FakeQ())(obs, masks), greedy_ref)
    pol = EpsilonGreedy(MaxQPolicy(FakeQ()), 5, 0.0)
    assert torch.equal(pol(obs, masks), greedy_ref)
    pol.epsilon = 1.0
    a = pol(obs, masks)
    assert bool(masks[torch.arange(b, device=dev), a].all())
    pol.epsilon = 0.25
    a = pol(obs, masks)
    assert bool(masks[torch.arange(b, device=dev), a].all())
    nv = masks.sum(1).float()
    expect = (0.75 + 0.25 / nv).mean().item()
    assert abs((a == greedy_ref).float().mean().item() - expect) < 0.015


def test_short_training_run(pb, tmp_path):
    import torch
    from pacbot_rl_b200 import train_dqn

    cfg = train_dqn.build_parser().parse_args([
        "--num_iters", "40", "--replay_buffer_size", "2048", "--batch_size", "256",
        "--log_every", "10", "--evaluate_steps", "2", "--device", "cuda:0",
        "--checkpoint-dir", str(tmp_path / "ckpt"), "--metrics-file", str(tmp_path / "m.jsonl")])
    trainer = train_dqn.train(cfg)
    assert trainer.iter_num == 40 and trainer._graph is not None    # captured CUDA graph path
    # host mirror of the ring position == the device-side counter the kernels used
    assert int(trainer.replay._t_dev.item()) == trainer.replay._t
    assert len(trainer.replay) == trainer.replay.capacity
    # the eager path runs the same body
    cfg2 = train_dqn.build_parser().parse_args([
        "--num_iters", "6", "--replay_buffer_size", "1024", "--batch_size", "128", "--log_every", "3",
        "--device", "cuda:0", "--no-checkpoints", "--no-cuda-graph"])
    eager = train_dqn.train(cfg2)
    assert eager._graph is None and eager.iter_num == 6
    assert int(eager.replay._t_dev.item()) == eager.replay._t
    import json

    rows = [json.loads(l) for l in open(tmp_path / "m.jsonl")]
    assert len(rows) == 4 and all(np.isfinite(r["loss"]) for r in rows)
    assert "eval_episode_score" in rows[1] and rows[1]["eval_pellets_start"] == 244
    assert rows[-1]["exploration_epsilon"] < 0.1
    # heads start at zero; after training they moved
    assert float(trainer.q_net.advantage_head[0].weight.abs().sum()) > 0
    # reference-format checkpoints (train_dqn.py:239-253)
    assert (tmp_path / "ckpt" / "q_net-latest.pt").exists()
    assert (tmp_path / "ckpt" / "q_net-iter0000000.pt").exists()
    assert (tmp_path / "ckpt" / "q_net-latest.safetensors").exists()
    loaded = torch.load(tmp_path / "ckpt" / "q_net-latest.pt", map_location="cuda:0", weights_only=False)
    assert sum(p.numel() for p in loaded.parameters()) == 156_934


class LastValidPolicy:
    """Deterministic 'old' policy for the two-stage reset test: highest-index valid action."""

    def __call__(self, obs, masks):
        idx = __import__("torch").arange(5, device=masks.device)
        return (masks.long() * idx).argmax(1)


def test_two_stage_reset_matches_reference_loop(pb, oracle):
    """reset_env (replay_buffer.py:44-54): after every reset the OLD policy plays, unrecorded,
    until first_ai_done(); only then are transitions recorded.  The batched buffer interleaves
    envs, but each env's own trajectory must equal the reference's sequential per-env loop."""
    import torch
    from conftest import oracle_to_gpu_state
    from pacbot_rl_b200.replay_buffer import ReplayBuffer

    n, total_steps = 16, 150
    rng = np.random.default_rng(4)
    tape = rng.integers(0, 2**32, size=(n, 263), dtype=np.uint32)
    main, old = FirstValidPolicy(), LastValidPolicy()
    buf = ReplayBuffer(n * (total_steps + 2), main, n, 0.5, torch.device("cuda:0"), seed=0, phase1_policy=old)
    buf.envs.set_draw_tape(torch.from_numpy(tape.view(np.int32)).cuda())
    # start state: one super pellet left, directly above Pac-Man (the old policy walks into it),
    # a few ordinary pellets, ghosts far away; half the envs have NO super pellet at all
    proto = oracle.OracleBatch(n, (np.arange(n) < n * 0.5).astype(np.uint8) | 2, tape)
    st = proto.export_state()
    st["paused"] = 0
    st["pellets"][:] = 0
    st["pellets"][:, 5] = 0x7FFFFFE
    st["num_pellets"] = 26
    with_super = np.arange(n) % 2 == 0
    st["pellets"][with_super, 3] = 1 << 1
    st["num_pellets"][with_super] = 27
    st["pac_row"], st["pac_col"], st["pac_dir"] = 4, 1, 0
    st["curr_ticks"] = 1
    buf.envs.set_state(oracle_to_gpu_state(pb, st))
    buf.envs.obs(out=buf._obs[0])
    buf.envs.action_mask(out=buf._mask)
    buf._in_phase1 = ~buf.envs.first_ai_done()
    assert buf._in_phase1.cpu().numpy().tolist() == with_super.tolist()

    # ---- reference semantics, env by env, on the oracle: replay_reference.restated_env_loop, which
    # tests/test_reference_replay_live.py checks against the reference's own replay_buffer.py ----
    from replay_reference import restated_env_loop

    def as_numpy(policy):
        return lambda obs, mask: policy(torch.from_numpy(obs), torch.from_numpy(mask)).numpy()

    expected = []   # per env: list of (obs_bits, action, reward, done, next_mask)
    for i in range(n):
        ob = oracle.OracleBatch(1, ((np.arange(n) < n * 0.5).astype(np.uint8) | 2)[i:i + 1], tape[i:i + 1])
        ob.import_state(st[i:i + 1])
        expected.append(restated_env_loop(ob, as_numpy(main), as_numpy(old), max_env_steps=total_steps))

    for _ in range(total_steps):
        buf.generate_experience_step()
    w = buf._valid.cpu().numpy()
    obs_ring = buf._obs.cpu().numpy().view(np.uint32)
    acts, rews, dones = buf._actions.cpu().numpy(), buf._rewards.cpu().numpy(), buf._done.cpu().numpy()
    next_masks = buf._next_mask.cpu().numpy()
    passed_through_phase1 = 0
    for i in range(n):
        got = [(obs_ring[s, i], int(acts[s, i]), int(rews[s, i]), bool(dones[s, i]), next_masks[s, i])
               for s in range(total_steps) if w[s, i] > 0]
        assert len(got) == len(expected[i]), (i, len(got), len(expected[i]))
        for k, (g, e) in enumerate(zip(got, expected[i])):
            assert g[1:4] == e[1:4], (i, k, g[1:4], e[1:4])
            assert np.array_equal(g[0], e[0][0]), (i, k)
            assert np.array_equal(g[4], e[4]), (i, k, "next_action_mask")
        if with_super[i] and len(got) > 0:
            passed_through_phase1 += 1
    assert passed_through_phase1 >= 1


def test_synchronous_reset_env_matches_reference_loop(pb, oracle):
    """`replay_buffer.reset_env` (the form evaluate_episode uses, train_dqn.py:95-96): every env is
    reset and played by the old policy until first_ai_done(); envs that are through (here: the ones
    built without super pellets) are parked and must come out untouched.  Expectation: the
    reference's loop (replay_buffer.py:44-54) env by env on the oracle, same draw tape."""
    import torch
    from conftest import assert_obs_equal, assert_states_equal
    from pacbot_rl_b200.replay_buffer import reset_env
    from replay_reference import OLD_POLICY_CHANNELS, bfs_policy

    n = 8
    rng = np.random.default_rng(9)
    tape = rng.integers(0, 2**32, size=(n, 1021), dtype=np.uint32)
    no_super = np.arange(n) % 4 == 3
    env = pb.BatchedPacmanGym(n, False, False, device="cuda:0", seed=0, remove_super_pellets=no_super)
    env.set_draw_tape(torch.from_numpy(tape.view(np.int32)).cuda())
    old_np = bfs_policy(OLD_POLICY_CHANNELS)
    calls = []

    def old(obs, masks):   # exact-cell BFS on the host: the same function on both sides
        calls.append(1)
        return torch.from_numpy(old_np(obs.cpu().numpy(), masks.cpu().numpy())).to(obs.device)

    reset_env(env, old)
    assert not env.auto_reset and bool(env.first_ai_done().all())
    ob = oracle.OracleBatch(n, no_super.astype(np.uint8) << 3, tape)
    ob.reset()
    longest = 0
    for i in range(n):
        one = oracle.OracleBatch(1, (no_super.astype(np.uint8) << 3)[i:i + 1], tape[i:i + 1])
        one.reset()
        k = 0
        while not one.first_ai_done()[0]:
            one.step(old_np(one.obs(), one.action_mask()), auto_reset=True)
            k += 1
        longest = max(longest, k)
        st = ob.export_state()
        st[i] = one.export_state()[0]
        ob.import_state(st)
        assert (k == 0) == bool(no_super[i])
    assert len(calls) == longest > 100          # stops as soon as the last env is through
    assert_states_equal(ob.export_state(), env.get_state(), "after reset_env")
    assert_obs_equal(ob.obs(), env.obs().cpu().numpy(), "after reset_env")
    # no old policy: a plain reset
    reset_env(env, None)
    ob.reset()
    assert_states_equal(ob.export_state(), env.get_state(), "plain reset")


# ------------------------------------------------------------------------------------------------
# the replay kernels behind the C ABI (csrc/pb_replay_abi.inl)
# ------------------------------------------------------------------------------------------------
def _sample(pb, valid, batch, seed, call):
    """pb_replay_sample through the C ABI -> (idx int64 [batch] host, num_valid)."""
    import ctypes as C

    import torch

    L = pb._lib.load()
    dev = valid.device
    call_dev = torch.tensor([call], dtype=torch.int64, device=dev)
    idx = torch.full((batch,), -1, dtype=torch.int64, device=dev)
    m = torch.zeros(1, dtype=torch.int32, device=dev)
    pb._lib.check(L.pb_replay_sample(
        C.c_void_p(valid.data_ptr()), valid.numel(), batch, seed, C.c_void_p(call_dev.data_ptr()),
        C.c_void_p(idx.data_ptr()), C.c_void_p(m.data_ptr()), dev.index,
        C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return idx.cpu().numpy(), int(m.item())


def test_replay_sample_is_a_partial_fisher_yates_over_the_valid_list(pb):
    """random.sample(buffer, k) (replay_buffer.py:141-143): k DISTINCT recorded transitions.  The
    kernel is specified exactly (include/pacbot_b200.h): ascending list of the valid flat indices,
    then position j swaps with j + ((raw_j * (m - j)) >> 32), raw_j = pb_replay_raw(seed, call, j).
    Replayed here on the host draw by draw; plus the m < k and m == 0 edge cases, the size limits,
    and uniformity over many calls."""
    import torch

    L = pb._lib.load()
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(5)
    total, batch, seed = 314 * 32, 512, 0xC0FFEE
    valid = (torch.rand(total, generator=g) > 0.35)
    valid_dev = valid.to(dev)
    for call in (0, 1, 977):
        idx, m = _sample(pb, valid_dev, batch, seed, call)
        lst = np.flatnonzero(valid.numpy()).tolist()
        assert m == len(lst)
        for j in range(batch):
            r = j + ((L.pb_replay_raw(seed, call, j) * (m - j)) >> 32)
            lst[j], lst[r] = lst[r], lst[j]
        assert idx.tolist() == lst[:batch], call
        assert len(set(idx.tolist())) == batch and valid.numpy()[idx].all()
    # fewer valid transitions than the batch: with replacement over what there is; none: index 0
    few = torch.zeros(total, dtype=torch.bool)
    few[[7, 4000, 9999]] = True
    idx, m = _sample(pb, few.to(dev), 64, seed, 3)
    assert m == 3 and set(idx.tolist()) <= {7, 4000, 9999}
    idx, m = _sample(pb, torch.zeros(total, dtype=torch.bool, device=dev), 64, seed, 3)
    assert m == 0 and not idx.any()
    # exactly as many as the batch: a permutation of all of them
    exact = torch.zeros(total, dtype=torch.bool)
    exact[torch.randperm(total, generator=g)[:64]] = True
    idx, m = _sample(pb, exact.to(dev), 64, seed, 4)
    assert m == 64 and sorted(idx.tolist()) == np.flatnonzero(exact.numpy()).tolist()
    # limits are refused, not silently wrong
    cap = L.pb_replay_sample_capacity()
    assert cap >= total
    big = torch.ones(cap + 1, dtype=torch.bool, device=dev)
    with pytest.raises(Exception):
        _sample(pb, big, 64, seed, 0)
    with pytest.raises(Exception):
        _sample(pb, valid_dev, 4097, seed, 0)
    # uniformity: every valid transition is drawn equally often (4 sigma per cell, chi-square overall)
    small = torch.zeros(2048, dtype=torch.bool)
    small[::2] = True
    small_dev, calls, k = small.to(dev), 1500, 128
    hits = np.zeros(2048, np.int64)
    first_pos = np.zeros(2048, np.int64)
    for call in range(calls):
        idx, m = _sample(pb, small_dev, k, seed + 1, call)
        np.add.at(hits, idx, 1)
        first_pos[idx[0]] += 1
    assert hits[1::2].sum() == 0 and m == 1024
    expect = calls * k / 1024
    chi2 = ((hits[::2] - expect) ** 2 / expect).sum()
    # without replacement the per-cell variance is expect * (1 - k/m): chi2 ~ (1 - k/m) * 1023
    assert 0.875 * 1023 - 5 * np.sqrt(2 * 1023) < chi2 < 0.875 * 1023 + 5 * np.sqrt(2 * 1023), chi2
    assert first_pos.max() < 12      # the first element is uniform too (mean 1.46 per cell)


@pytest.mark.parametrize("channels_last,pad", [(False, False), (True, False), (True, True)],
                         ids=["nchw", "nhwc", "nhwc-padded32"])
def test_replay_record_collate_and_slot_obs_match_plain_indexing(pb, channels_last, pad):
    """k_replay_record / k_replay_collate / k_replay_slot against the plain torch expressions they
    replace (the collate of train_dqn.py:153-165), NCHW and channels-last output, values compared
    bit for bit; a channels-last batch must have channels-last strides (that is what removes the
    layout conversion in front of the NHWC convolutions)."""
    import torch
    from pacbot_rl_b200.models import QNetV2
    from pacbot_rl_b200.policies import EpsilonGreedy, MaxQPolicy
    from pacbot_rl_b200.replay_buffer import ReplayBuffer

    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    q = QNetV2(pb.OBS_SHAPE, 5).to(dev)
    if channels_last:
        q.use_channels_last()
    n, maxlen = 32, 2048
    buf = ReplayBuffer(maxlen, EpsilonGreedy(MaxQPolicy(q), 5, 1.0), n, 0.5, dev, seed=9,
                       channels_last=channels_last, pad_channels=pad)
    r = buf._ring
    assert buf.obs_channels == (32 if pad else 17)

    def real(x):   # the 17 observation channels; padded ones must be zero
        assert x.shape[1] == buf.obs_channels and not bool(x[:, 17:].any())
        return x[:, :17]

    dual = channels_last and pad                 # the encoder writes the policy's next input itself
    for t in range(r + 30):                      # wrap the ring
        before = buf._obs[t % r].clone()
        if dual and t > 0:                       # what the policy is about to see == ring slot t
            assert torch.equal(real(buf._last_obs), before), t
        buf.generate_experience_step()
        if dual:     # pb_encode_obs_ring_nhwc32: slot t + 1 and its channels-last twin, one launch
            assert torch.equal(real(buf._last_obs), buf._obs[(t + 1) % r]), t
        else:        # last_obs handed to the policy == the ring slot of step t
            assert torch.equal(real(buf._last_obs), before), t
        assert buf._last_obs.is_contiguous(
            memory_format=torch.channels_last if channels_last else torch.contiguous_format)
        # what k_replay_record appended == the step's own outputs
        s = t % r
        assert torch.equal(buf._actions[s], buf._act) and torch.equal(buf._rewards[s], buf._reward)
        assert torch.equal(buf._done[s], buf._done_step)
        assert torch.equal(buf._next_mask[s], buf._mask & ~buf._done_step[:, None])
        assert bool(buf._valid[s].all()) and not bool(buf._valid[(t + 1) % r].any())
    assert int(buf._done.sum()) > 0              # finished episodes are in the ring
    for b in (1, 7, 512):
        batch = buf.sample_batch(b)
        flat = buf._batch_buffers(b)["idx"]
        slot, env = flat // n, flat % n
        done = buf._done.view(-1)[flat]
        assert torch.equal(real(batch.obs), buf._obs.view(-1, *pb.OBS_SHAPE)[flat])
        nxt = buf._obs[(slot + 1) % r, env]
        nxt[done] = 0.0                          # `next_obs is None` -> zeros (train_dqn.py:154-159)
        assert torch.equal(real(batch.next_obs), nxt)
        assert batch.actions.dtype == torch.int64 and torch.equal(batch.actions, buf._actions.view(-1)[flat].long())
        assert batch.rewards.dtype == torch.float32 and torch.equal(batch.rewards, buf._rewards.view(-1)[flat].float())
        assert torch.equal(batch.done, done)
        assert torch.equal(batch.next_action_masks, buf._next_mask.view(-1, 5)[flat])
        fmt = torch.channels_last if channels_last else torch.contiguous_format
        assert batch.obs.is_contiguous(memory_format=fmt) and batch.next_obs.is_contiguous(memory_format=fmt)
    assert bool(batch.done.any()) and not bool(batch.done.all())
    # the network gives the same Q-values for the padded batch as for the 17 channels alone
    old_tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            torch.testing.assert_close(q(batch.obs), q(real(batch.obs).contiguous()), rtol=1e-5, atol=1e-6)
    finally:
        torch.backends.cudnn.allow_tf32 = old_tf32


def test_fill_with_old_policy_waits_for_recorded_transitions(pb):
    """With the two-stage reset every env starts in the old policy's phase and nothing is recorded
    until it has eaten the super pellets: a ring that has merely wrapped can hold fewer recorded
    transitions than a batch.  `fill(min_valid)` goes on until there are enough, `len()` reports
    recorded transitions (len(deque) in the reference), and sampling only returns those."""
    import torch
    from pacbot_rl_b200.replay_buffer import ReplayBuffer
    from replay_reference import OLD_POLICY_CHANNELS, bfs_policy

    dev = torch.device("cuda:0")
    old_np = bfs_policy(OLD_POLICY_CHANNELS)

    def old(obs, masks):
        return torch.from_numpy(old_np(obs.cpu().numpy(), masks.cpu().numpy())).to(obs.device)

    n = 4
    buf = ReplayBuffer(64, FirstValidPolicy(), n, 0.5, dev, seed=2, phase1_policy=old)
    assert len(buf) == 0 and bool(buf._in_phase1.all())
    for _ in range(buf._slots):
        buf.generate_experience_step()
    assert len(buf) == 0                         # the ring wrapped, nothing recorded yet
    buf.fill(min_valid=24)
    assert 24 <= len(buf) <= buf.capacity
    idx = buf.sample_indices(16)
    assert bool(buf._valid.view(-1)[idx].all()) and idx.unique().numel() == 16
    assert buf.num_valid == len(buf)

    class Never:   # an old policy that never gets through phase 1: fill() must say so, not spin
        def __call__(self, obs, masks):
            return torch.zeros(obs.shape[0], dtype=torch.int64, device=obs.device)

    stuck = ReplayBuffer(64, FirstValidPolicy(), n, 0.0, dev, seed=2, phase1_policy=Never())
    with pytest.raises(RuntimeError, match="recorded transitions"):
        stuck.fill(min_valid=8)


# ------------------------------------------------------------------------------------------------
# a17: the double-DQN target / loss arithmetic against the reference's own expressions
# ------------------------------------------------------------------------------------------------
def _reference_target_and_loss(q_net, target_q_net, obs_batch, next_obs_batch, done_mask,
                               next_action_masks, action_batch, raw_rewards, reward_scale, discount_factor):
    """train_dqn.py:153-197 with the same statements (index assignment, fancy indexing)."""
    import torch
    import torch.nn.functional as F

    n = obs_batch.shape[0]
    reward_batch = torch.tensor([r * reward_scale for r in raw_rewards.tolist()], device=obs_batch.device)
    with torch.no_grad():
        next_q_values = target_q_net(next_obs_batch)
        next_q_values[~next_action_masks] = -torch.inf
        online_next_q_values = q_net(next_obs_batch)
        online_next_q_values[~next_action_masks] = -torch.inf
        next_actions = online_next_q_values.argmax(dim=1)
        returns = next_q_values[range(n), next_actions]
        discounted_returns = discount_factor * returns
        discounted_returns[done_mask] = 0.0
        target_q_values = reward_batch + discounted_returns
    all_predicted_q_values = q_net(obs_batch)
    predicted_q_values = all_predicted_q_values[range(n), action_batch]
    loss = F.mse_loss(predicted_q_values, target_q_values)
    return target_q_values, loss, all_predicted_q_values


def test_double_dqn_target_and_loss_match_reference_expressions(pb):
    """One fixed ReplayBatch through `dqn_targets` / `dqn_loss` (what `_iteration_body` runs) vs
    the reference's statements (train_dqn.py:167-197): on the same device with the same networks
    (tolerance 1e-6: same cuDNN kernels underneath, only the target/loss arithmetic differs), on
    the CPU in float32 with exactly-representable stand-in networks (rtol 1e-6 target, 1e-5 loss)
    and on the CPU in float64 with the real networks (5e-4: cuDNN fp32 convolutions, TF32 off).  The batch has finished
    transitions (done, all-False next mask, zero next_obs), partial next masks, negative and large
    rewards; reward_scale and gamma are not the defaults."""
    import copy

    import torch
    from pacbot_rl_b200 import train_dqn
    from pacbot_rl_b200.models import QNetV2
    from pacbot_rl_b200.replay_buffer import ReplayBatch

    dev = torch.device("cuda:0")
    old_tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        torch.manual_seed(11)
        q_net = QNetV2(pb.OBS_SHAPE, 5).to(dev)
        with torch.no_grad():                    # the heads start at zero: make Q depend on the input
            for head in (q_net.value_head, q_net.advantage_head):
                for p in head.parameters():
                    p.normal_(0.0, 0.5)
        target_q_net = copy.deepcopy(q_net)
        with torch.no_grad():
            for p in target_q_net.parameters():
                p.add_(0.05 * torch.randn_like(p))
        b = 96
        g = torch.Generator(device="cpu").manual_seed(12)
        obs = (torch.rand((b,) + pb.OBS_SHAPE, generator=g) < 0.1).float() * torch.rand((b,) + pb.OBS_SHAPE, generator=g)
        next_obs = (torch.rand((b,) + pb.OBS_SHAPE, generator=g) < 0.1).float() * torch.rand((b,) + pb.OBS_SHAPE, generator=g)
        done = torch.rand(b, generator=g) < 0.3
        masks = torch.rand((b, 5), generator=g) < 0.6
        masks[:, 0] = True
        masks[done] = False                      # [False] * 5 when the episode ended
        next_obs[done] = 0.0
        actions = torch.randint(0, 5, (b,), generator=g)
        rewards = torch.randint(-250, 3200, (b,), generator=g).float()
        batch = ReplayBatch(obs.to(dev), actions.to(dev), rewards.to(dev), next_obs.to(dev), done.to(dev),
                            masks.to(dev))
        gamma, scale = 0.97, 1 / 40
        target = train_dqn.dqn_targets(batch, q_net, target_q_net, gamma, scale)
        loss, all_q = train_dqn.dqn_loss(batch, q_net, target)
        q_net.zero_grad()
        loss.backward()
        grads = [p.grad.clone() for p in q_net.parameters()]
        assert bool(torch.isfinite(target).all()) and bool((target[done.to(dev)] == (rewards.to(dev) * scale)[done.to(dev)]).all())
        # (1) the reference statements on the same device
        q_net.zero_grad()
        t_ref, l_ref, q_ref = _reference_target_and_loss(
            q_net, target_q_net, batch.obs, batch.next_obs, batch.done, batch.next_action_masks,
            batch.actions, rewards, scale, gamma)
        l_ref.backward()
        torch.testing.assert_close(target, t_ref, rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(loss, l_ref, rtol=1e-6, atol=1e-7)
        torch.testing.assert_close(all_q, q_ref, rtol=1e-6, atol=1e-6)
        for got, p in zip(grads, q_net.parameters()):   # cuDNN's backward accumulates with atomics
            torch.testing.assert_close(got, p.grad, rtol=1e-4, atol=1e-5)
        # (2) the reference statements on the CPU, float32, with networks whose arithmetic is exact
        # in any evaluation order (two cells, a power-of-two scale, one add): what is compared is
        # the target / loss arithmetic alone, GPU product path vs CPU reference statements
        class CellNet(torch.nn.Module):
            def __init__(self, shift):
                super().__init__()
                self.w = torch.nn.Parameter(torch.tensor(2.0))
                self.shift = shift

            def forward(self, x):
                return x[:, :5, 3, 4] * self.w + x[:, 5:10, 7 + self.shift, 9] - 0.5

        q_cell, t_cell = CellNet(0), CellNet(1)
        dense = ReplayBatch(torch.rand((b,) + pb.OBS_SHAPE, generator=g).to(dev), batch.actions, batch.rewards,
                            (torch.rand((b,) + pb.OBS_SHAPE, generator=g) * (~done)[:, None, None, None]).to(dev),
                            batch.done, batch.next_action_masks)
        tgt = train_dqn.dqn_targets(dense, q_cell.to(dev), t_cell.to(dev), gamma, scale)
        lss, _ = train_dqn.dqn_loss(dense, q_cell, tgt)
        t_cpu, l_cpu, _ = _reference_target_and_loss(
            copy.deepcopy(q_cell).cpu(), copy.deepcopy(t_cell).cpu(), dense.obs.cpu(), dense.next_obs.cpu(),
            done, masks, actions, rewards, scale, gamma)
        torch.testing.assert_close(tgt.cpu(), t_cpu, rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(lss.cpu(), l_cpu, rtol=1e-5, atol=0)
        # (3) the real networks on the CPU in float64: cuDNN's fp32 convolutions against a float64
        # evaluation bound the agreement here, not the target arithmetic
        q64, t64 = copy.deepcopy(q_net).cpu().double(), copy.deepcopy(target_q_net).cpu().double()
        t_cpu, l_cpu, _ = _reference_target_and_loss(
            q64, t64, obs.double(), next_obs.double(), done, masks, actions, rewards.double(), scale, gamma)
        torch.testing.assert_close(target.cpu().double(), t_cpu.double(), rtol=5e-4, atol=5e-4)
        torch.testing.assert_close(loss.cpu().double(), l_cpu.double(), rtol=5e-4, atol=1e-5)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old_tf32


# ------------------------------------------------------------------------------------------------
# f2: batched evaluate_episode against the reference's per-env loop on the oracle
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("max_steps,with_old_policy", [(400, True), (75, False)])
def test_evaluate_episode_matches_reference_loop(pb, oracle, max_steps, with_old_policy):
    """`run_eval_episodes(num_envs=K)` vs `evaluate_episode` (train_dqn.py:88-113) run env by env on
    the oracle with the same draws: score, steps, cleared, pellets_start and pellets_end per env.
    The policies are host functions of exactly representable observation cells, so both sides
    compute the same actions; 75 steps is shorter than any episode (the truncated case)."""
    import torch
    from pacbot_rl_b200 import train_dqn
    from replay_reference import MAIN_POLICY_CHANNELS, OLD_POLICY_CHANNELS, bfs_policy

    dev = torch.device("cuda:0")
    k, seed, base = 6, 23, 10_000_000
    main_np, old_np = bfs_policy(MAIN_POLICY_CHANNELS), bfs_policy(OLD_POLICY_CHANNELS)

    def on_device(f):
        return lambda obs, masks: torch.from_numpy(f(obs.cpu().numpy(), masks.cpu().numpy())).to(obs.device)

    per = train_dqn.run_eval_episodes(on_device(main_np), dev, k, max_steps, seed=seed, env_id_base=base,
                                      phase1_policy=on_device(old_np) if with_old_policy else None)
    finished = 0
    for i in range(k):
        gym = oracle.OracleBatch(1, np.zeros(1, np.uint8), None)   # random_start=False, random_ticks=False
        gym.set_counter_rng(seed, base + i)

        def st():
            return gym.export_state()[0]

        # reset_env(gym) (replay_buffer.py:44-54)
        gym.reset()
        if with_old_policy:
            while not gym.first_ai_done()[0]:
                _, d = gym.step(old_np(gym.obs(), gym.action_mask()))
                if d[0]:
                    gym.reset()
        pellets_start = int(st()["num_pellets"])
        done = False
        for step_num in range(1, max_steps + 1):
            _, d = gym.step(main_np(gym.obs(), gym.action_mask()))
            done = bool(d[0])
            if done:
                break
        cleared = done and int(st()["curr_lives"]) == 3
        pellets_end = 0 if cleared else int(st()["num_pellets"])
        expect = (int(st()["curr_score"]), step_num, int(cleared), pellets_start, pellets_end, int(done))
        got = tuple(int(per[f][i]) for f in train_dqn.EVAL_FIELDS)
        assert got == expect, (i, got, expect)
        finished += int(done)
    if max_steps == 75:
        assert finished == 0
    else:
        assert finished >= 1


def test_global_max_pool_matches_adaptive_max_pool_on_gpu(pb):
    """models.GlobalMaxPool2d vs nn.AdaptiveMaxPool2d((1, 1)) on the device kernels the trainer
    runs (channels-last, batch 512): same values, same gradient routing, ties included."""
    import torch
    from pacbot_rl_b200.models import GlobalMaxPool2d

    torch.manual_seed(5)
    x = torch.randn(512, 128, 4, 4, device="cuda:0")
    x[:, ::7] = x[:, ::7].round()                 # plenty of tied maxima
    for fmt in (torch.contiguous_format, torch.channels_last):
        a = x.clone().contiguous(memory_format=fmt).requires_grad_(True)
        b = x.clone().contiguous(memory_format=fmt).requires_grad_(True)
        ya, yb = torch.nn.AdaptiveMaxPool2d((1, 1))(a), GlobalMaxPool2d()(b)
        assert torch.equal(ya, yb)
        w = torch.randn_like(ya)
        (ya * w).sum().backward()
        (yb * w).sum().backward()
        assert torch.equal(a.grad, b.grad)


# ------------------------------------------------------------------------------------------------
# the fused tail of QNetV2's convolution blocks (csrc/pb_qnet_abi.inl)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape,pool", [((512, 16, 28, 31), False), ((64, 32, 28, 31), True),
                                        ((9, 64, 14, 16), True), ((3, 128, 7, 8), True), ((1, 32, 1, 1), True)])
def test_bias_pool_silu_matches_torch_composition(pb, shape, pool):
    """pb_bias_pool_silu_fwd / _bwd against what torch runs for `Conv2d -> [MaxPool2d(2,
    ceil_mode=True)] -> SiLU` after the convolution (reference models.py:70-90): bias add, max
    pool, SiLU and their backward.  Odd sizes (partial border windows), tied maxima (arg-max must
    be torch's: first in row-major order) and NaN.  Forward bit for bit; backward to float32
    rounding (torch sums the bias gradient in another order; FMA contraction may differ by an ulp)."""
    import torch
    import torch.nn.functional as F
    from pacbot_rl_b200.models import bias_pool_silu

    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(hash(shape) % 1000)
    x = torch.randn(shape, generator=g)
    x[:, ::3] = (x[:, ::3] * 2).round() / 2            # many ties inside windows
    bias = torch.randn(shape[1], generator=g)
    x = x.to(dev).contiguous(memory_format=torch.channels_last)
    bias = bias.to(dev)
    xa, ba = x.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    xb, bb = x.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    za = xa + ba.view(1, -1, 1, 1)
    if pool:
        za = F.max_pool2d(za, 2, ceil_mode=True)
    ya = F.silu(za)
    yb = bias_pool_silu(xb, bb, pool)
    assert yb.shape == ya.shape and yb.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(ya, yb)
    w = torch.randn(ya.shape, generator=g).to(dev)
    (ya * w).sum().backward()
    (yb * w).sum().backward()
    torch.testing.assert_close(xb.grad, xa.grad, rtol=1e-5, atol=1e-6)
    assert torch.equal(xb.grad == 0, xa.grad == 0)     # the same cells carry the gradient
    torch.testing.assert_close(bb.grad, ba.grad, rtol=1e-4, atol=1e-4)
    # NaN propagates like torch's pooling does; no backward buffers under no_grad
    xn = x.clone()
    xn[0, 0, 0, 0] = float("nan")
    with torch.no_grad():
        zn = xn + bias.view(1, -1, 1, 1)
        ref = F.silu(F.max_pool2d(zn, 2, ceil_mode=True) if pool else zn)
        got = bias_pool_silu(xn, bias, pool)
    assert torch.equal(torch.isnan(ref), torch.isnan(got)) and bool(torch.isnan(got[0, 0, 0, 0]))
    assert torch.equal(torch.nan_to_num(ref), torch.nan_to_num(got))


@pytest.mark.parametrize("n", [1, 32, 300, 512, 1100])
def test_qnet_head_matches_torch_composition(pb, n):
    """pb_qnet_head_fwd (global max + bias, SiLU, Linear, SiLU, dueling / plain heads in one
    launch) against the torch modules it replaces (reference models.py:92-110, :139-141):
    rtol 1e-5 / atol 1e-6 in float32 (cuBLAS orders the sums differently); the pooled maximum
    itself follows torch's NaN rule."""
    import torch
    from torch import nn
    from pacbot_rl_b200.models import GlobalMaxPool2d, qnet_head

    dev = torch.device("cuda:0")
    torch.manual_seed(n)
    x = torch.randn((n, 128, 4, 4), device=dev).contiguous(memory_format=torch.channels_last)
    conv_bias = torch.randn(128, device=dev) * 0.5
    lin, adv, val, plain = nn.Linear(128, 256).to(dev), nn.Linear(256, 5).to(dev), nn.Linear(256, 1).to(dev), \
        nn.Linear(256, 6).to(dev)
    with torch.no_grad():
        feat = nn.SiLU()(lin(nn.SiLU()(nn.Flatten()(GlobalMaxPool2d()(x + conv_bias.view(1, -1, 1, 1))))))
        a = adv(feat)
        ref_q = val(feat) - a.mean(dim=1, keepdim=True) + a
        torch.testing.assert_close(qnet_head(x, conv_bias, lin, adv, val), ref_q, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(qnet_head(x, conv_bias, lin, plain), plain(feat), rtol=1e-5, atol=1e-6)
        # a 7 x 8 map (another hw), and a NaN in one map poisons exactly that sample
        x2 = torch.randn((min(n, 40), 128, 7, 8), device=dev).contiguous(memory_format=torch.channels_last)
        x2[0, 5, 3, 2] = float("nan")
        feat2 = nn.SiLU()(lin(nn.SiLU()(nn.Flatten()(GlobalMaxPool2d()(x2 + conv_bias.view(1, -1, 1, 1))))))
        got = qnet_head(x2, conv_bias, lin, plain)
        assert bool(torch.isnan(got[0]).all()) and not bool(torch.isnan(got[1:]).any())
        torch.testing.assert_close(got[1:], plain(feat2)[1:], rtol=1e-5, atol=1e-6)
    with pytest.raises(ValueError):
        qnet_head(torch.randn((2, 64, 4, 4), device=dev), conv_bias, lin, adv, val)


@pytest.mark.parametrize("n,dueling", [(32, True), (300, True), (512, True), (1100, True), (40, False)])
def test_qnet_head_train_matches_torch_autograd(pb, n, dueling):
    """`qnet_head_train` (pb_qnet_head_fwd_train + pb_qnet_head_bwd): Q-values and EVERY gradient
    (the convolution output incl. the routing to the arg-max cells, the convolution bias, both
    linear layers, the dueling heads) against autograd through the torch modules it replaces
    (reference models.py:92-110 / :139-141).  float32 both ways; the batch reductions run in
    another order than cuBLAS: rtol 1e-4, atol 1e-5 of each tensor's largest entry."""
    import torch
    from torch import nn
    from pacbot_rl_b200.models import GlobalMaxPool2d, qnet_head_train

    dev = torch.device("cuda:0")
    torch.manual_seed(100 + n)
    x0 = torch.randn((n, 128, 4, 4), device=dev)
    x0[:, 7] = x0[:, 7].round()            # ties inside a map: the first maximum must get the gradient
    lin, out = nn.Linear(128, 256).to(dev), nn.Linear(256, 5 if dueling else 6).to(dev)
    val = nn.Linear(256, 1).to(dev) if dueling else None
    conv_bias = (torch.randn(128, device=dev) * 0.5).requires_grad_()
    weight = torch.randn((n, out.out_features), device=dev)      # d(loss)/dq

    def run(fused: bool):
        for m in (lin, out, val):
            if m is not None:
                m.zero_grad()
        conv_bias.grad = None
        x = x0.clone().contiguous(memory_format=torch.channels_last).requires_grad_()
        if fused:
            q = qnet_head_train(x, conv_bias, lin, out, val)
        else:
            feat = nn.SiLU()(lin(nn.SiLU()(nn.Flatten()(GlobalMaxPool2d()(x + conv_bias.view(1, -1, 1, 1))))))
            a = out(feat)
            q = a if val is None else val(feat) - a.mean(dim=1, keepdim=True) + a
        (q * weight).sum().backward()
        grads = {"x": x.grad, "conv_bias": conv_bias.grad, "w1": lin.weight.grad, "b1": lin.bias.grad,
                 "wa": out.weight.grad, "ba": out.bias.grad}
        if val is not None:
            grads.update(wv=val.weight.grad, bv=val.bias.grad)
        return q.detach().clone(), {k: v.detach().clone() for k, v in grads.items()}

    q_ref, g_ref = run(False)
    q_got, g_got = run(True)
    torch.testing.assert_close(q_got, q_ref, rtol=1e-5, atol=1e-6)
    for name, ref in g_ref.items():
        got = g_got[name]
        assert got.shape == ref.shape, name
        scale = float(ref.abs().max())
        torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-5 * scale, msg=lambda m, name=name: f"{name}: {m}")
    # the gradient of the map is non-zero in exactly one cell per (sample, channel): the arg-max
    nz = (g_got["x"] != 0).sum(dim=(2, 3))
    assert int(nz.max()) <= 1 and torch.equal(g_got["x"] != 0, g_ref["x"] != 0)


def test_greedy_head_matches_max_q_policy(pb):
    """`QNetV2.greedy_actions` / `MaxQPolicy.actions_u8` (the masked arg-max taken inside the head
    launch, pb_qnet_head_greedy) == `q.masked_fill(~mask, -inf).argmax(1)` on the SAME Q-values
    (reference policies.py:44-59), incl. rows with a single valid action and rows with none."""
    import copy

    import torch
    from pacbot_rl_b200.models import QNetV2
    from pacbot_rl_b200.policies import MaxQPolicy

    dev = torch.device("cuda:0")
    torch.manual_seed(9)
    net = QNetV2(pb.OBS_SHAPE, 5).to(dev)
    with torch.no_grad():
        for head in (net.value_head, net.advantage_head):
            head[0].weight.normal_(0, 0.3)
            head[0].bias.normal_(0, 0.3)
    plain = copy.deepcopy(net).use_channels_last().eval()
    net.use_fused_blocks().eval()
    for n in (1, 32, 200, 700):
        obs = (torch.rand((n,) + pb.OBS_SHAPE, device=dev) < 0.15).float()
        mask = torch.rand((n, 5), device=dev) < 0.6
        mask[:, 0] = True
        if n > 2:
            mask[1] = False              # no valid action: argmax of all -inf is 0
            mask[2] = False
            mask[2, 3] = True            # exactly one
        with torch.no_grad():
            q = net(obs)                                                      # the fused head's own Q-values
            want = q.masked_fill(~mask, -torch.inf).argmax(dim=1)
            got = net.greedy_actions(obs, mask)
            assert got.dtype == torch.uint8 and torch.equal(got.long(), want)
            assert torch.equal(MaxQPolicy(net).actions_u8(obs, mask), got)
            # and the plain network's policy agrees wherever its two best valid Q-values are not within rounding
            qp = plain(obs).masked_fill(~mask, -torch.inf)
            top2 = qp.topk(2, dim=1).values
            clear = (top2[:, 0] - top2[:, 1]) > 1e-4
            assert torch.equal(MaxQPolicy(plain)(obs, mask)[clear], want[clear])
            assert torch.equal(MaxQPolicy(plain).actions_u8(obs, mask).long(), MaxQPolicy(plain)(obs, mask))


def test_netv2_policy_value_matches_the_evaluator_expressions(pb):
    """`NetV2.policy_value` with the fused blocks on (value scaling + masked softmax inside the head
    launch, pb_qnet_head_policy_value) against the evaluator of train_alphazero.py:83-95 written
    with torch ops on the plain network: rtol 1e-5 / atol 1e-6; invalid actions get exactly 0."""
    import copy

    import torch
    from pacbot_rl_b200.models import NetV2

    dev = torch.device("cuda:0")
    torch.manual_seed(12)
    plain = NetV2(pb.OBS_SHAPE, 6).to(dev).use_channels_last().eval()
    with torch.no_grad():
        plain.network[-1].weight.normal_(0, 0.5)
        plain.network[-1].bias.normal_(0, 0.5)
    fused = copy.deepcopy(plain).use_fused_blocks().eval()
    old_tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        for n in (1, 128, 500):
            obs = (torch.rand((n,) + pb.OBS_SHAPE, device=dev) < 0.15).float()
            mask = torch.rand((n, 5), device=dev) < 0.6
            mask[:, 0] = True
            with torch.no_grad():
                pred = plain(obs)
                want_v = pred[:, -1] * 100.0
                want_p = torch.softmax(pred[:, :-1].masked_fill(~mask, -torch.inf), dim=-1)
                got_v, got_p = fused.policy_value(obs, mask, 100.0)
                ref_v, ref_p = plain.policy_value(obs, mask, 100.0)       # the unfused route of the same method
            torch.testing.assert_close(got_v, want_v, rtol=1e-5, atol=1e-4)
            torch.testing.assert_close(got_p, want_p, rtol=1e-5, atol=1e-6)
            assert torch.equal(ref_v, want_v) and torch.equal(ref_p, want_p)
            assert not bool(got_p[~mask].any()) and torch.allclose(got_p.sum(1), torch.ones(n, device=dev), atol=1e-6)
    finally:
        torch.backends.cudnn.allow_tf32 = old_tf32


def test_obs_to_padded_nhwc_is_a_pure_layout_change(pb):
    import torch
    from pacbot_rl_b200.models import obs_to_padded_nhwc

    for n in (1, 33, 128):
        x = torch.rand((n,) + pb.OBS_SHAPE, device="cuda:0")
        y = obs_to_padded_nhwc(x)
        assert y.shape == (n, 32, 28, 31) and y.is_contiguous(memory_format=torch.channels_last)
        assert torch.equal(y[:, :17], x) and not bool(y[:, 17:].any())


def test_netv2_fused_blocks_match_the_plain_network(pb):
    """NetV2.use_fused_blocks() (the MCTS evaluator's network): inference through the fused block
    tails + `qnet_head` and training through the fused tails agree with the plain module."""
    import copy

    import torch
    from pacbot_rl_b200.models import NetV2

    dev = torch.device("cuda:0")
    old_tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        torch.manual_seed(5)
        plain = NetV2(pb.OBS_SHAPE, 6).to(dev).use_channels_last()
        with torch.no_grad():
            for p in plain.parameters():
                if p.dim() == 1:
                    p.normal_(0, 0.3)
            plain.network[-1].weight.normal_(0, 0.3)
        fused = copy.deepcopy(plain).use_fused_blocks()
        x = (torch.rand((40,) + pb.OBS_SHAPE, device=dev) < 0.15).float() * torch.rand((40,) + pb.OBS_SHAPE, device=dev)
        with torch.no_grad():
            torch.testing.assert_close(fused(x), plain(x), rtol=1e-5, atol=1e-6)
        qa, qb = plain(x), fused(x)                # raw layout: fused pads to 32 channels itself
        torch.testing.assert_close(qb, qa, rtol=1e-5, atol=1e-6)
        xp = torch.nn.functional.pad(x, (0, 0, 0, 0, 0, 15)).contiguous(memory_format=torch.channels_last)
        torch.testing.assert_close(fused(xp), qb, rtol=0, atol=0)   # == the padded NHWC batch given directly
        (qa ** 2).sum().backward()
        (qb ** 2).sum().backward()
        for (name, pa), pf in zip(plain.named_parameters(), fused.parameters()):
            scale = float(pa.grad.abs().max())
            assert float((pf.grad - pa.grad).abs().max()) <= 5e-3 * scale, name
    finally:
        torch.backends.cudnn.allow_tf32 = old_tf32


def test_qnetv2_fused_blocks_match_the_plain_network(pb):
    """QNetV2.use_fused_blocks(): same Q-values (TF32 off: the same cuDNN convolutions, the block
    tails bit-identical, the head to float32 rounding) and the same gradients as the plain torch
    module, for the 17-channel and the zero-padded 32-channel input, train and no_grad."""
    import copy

    import torch
    from pacbot_rl_b200.models import QNetV2

    dev = torch.device("cuda:0")
    old_tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        torch.manual_seed(2)
        plain = QNetV2(pb.OBS_SHAPE, 5).to(dev).use_channels_last()
        with torch.no_grad():
            for p in plain.parameters():
                if p.dim() == 1:
                    p.normal_(0, 0.3)            # non-zero biases
            for head in (plain.value_head, plain.advantage_head):
                head[0].weight.normal_(0, 0.3)
        fused = copy.deepcopy(plain).use_fused_blocks()
        twin = copy.deepcopy(plain)              # the run-to-run noise of the plain module itself
        g = torch.Generator(device="cpu").manual_seed(3)
        obs = ((torch.rand((48,) + pb.OBS_SHAPE, generator=g) < 0.15).float()
               * torch.rand((48,) + pb.OBS_SHAPE, generator=g)).to(dev)
        for x in (obs, torch.nn.functional.pad(obs, (0, 0, 0, 0, 0, 15))):
            with torch.no_grad():
                # inference runs the head as one kernel (pb_qnet_head_fwd): float32 sums in another
                # order than cuBLAS, everything before it is bit-identical
                torch.testing.assert_close(fused(x), plain(x), rtol=1e-5, atol=1e-6)
            for net in (plain, fused, twin):
                net.zero_grad()
            qa, qb = plain(x), fused(x)
            # the head is one kernel with gradients too (float32 sums in another order than cuBLAS);
            # for the raw encoder layout the fused network also pads the batch to 32 channels itself
            torch.testing.assert_close(qb, qa, rtol=1e-5, atol=1e-6)
            (qa ** 2).sum().backward()
            (qb ** 2).sum().backward()
            # the reference for "noise": the PLAIN module once more -- on the 32-channel padded batch
            # when the fused network pads the raw layout itself (another cuDNN weight-gradient
            # kernel for the first convolution than for the 17-channel input)
            (twin(torch.nn.functional.pad(obs, (0, 0, 0, 0, 0, 15)) if x is obs else x) ** 2).sum().backward()
            # gradients, relative to each tensor's largest entry: 5e-6 (bias and head gradients are
            # summed in another order), or the difference between two runs of the PLAIN module
            # where that is larger (cuDNN's fp32 weight-gradient kernel for the first convolution
            # is not run-to-run deterministic: 1.5e-3 between identical plain networks)
            for (name, pa), pf, pt in zip(plain.named_parameters(), fused.parameters(), twin.parameters()):
                scale = float(pa.grad.abs().max())
                err = float((pf.grad - pa.grad).abs().max()) / scale
                noise = float((pt.grad - pa.grad).abs().max()) / scale
                assert err <= max(5e-6, 3 * noise), (name, err, noise)
        # checkpoints are written from a plain copy
        from pacbot_rl_b200.models import portable_copy

        assert not getattr(portable_copy(fused), "_fused_blocks", False)
    finally:
        torch.backends.cudnn.allow_tf32 = old_tf32


@pytest.mark.parametrize("max_norm", [10.0, 0.05])
def test_clip_adam_matches_clip_grad_norm_and_torch_adam(pb, max_norm):
    """optim.ClipAdam (pb_clip_adam: two launches) against the reference's pair
    `clip_grad_norm_(params, max_norm)` + `torch.optim.Adam.step()` (train_dqn.py:133,200-206) on
    QNetV2's own parameter list in channels-last layout, 25 steps of random gradients of very
    different scales (clipping active and inactive): returned norm, weights and both moments
    within float32 rounding (rtol 1e-5); also inside a replayed CUDA graph."""
    import torch

    from pacbot_rl_b200 import models
    from pacbot_rl_b200.optim import ClipAdam

    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net_a = models.QNetV2(pb.OBS_SHAPE, 5).to(dev).use_channels_last()
    net_b = models.QNetV2(pb.OBS_SHAPE, 5).to(dev).use_channels_last()
    net_b.load_state_dict(net_a.state_dict())
    pa, pbs = list(net_a.parameters()), list(net_b.parameters())
    ours = ClipAdam(pa, lr=3e-4)
    ref = torch.optim.Adam(pbs, lr=3e-4)
    gen = torch.Generator(device=dev).manual_seed(1)
    for step in range(25):
        scale = 10.0 ** ((step % 5) - 3)
        for x, y in zip(pa, pbs):
            g = torch.randn(x.shape, device=dev, generator=gen) * scale
            x.grad = torch.empty_like(x).copy_(g)          # the parameter's own layout
            y.grad = torch.empty_like(y).copy_(g)
        n_ref = torch.nn.utils.clip_grad_norm_(pbs, max_norm=max_norm)
        ref.step()
        n_ours = ours.clip_and_step(max_norm)
        torch.testing.assert_close(n_ours, n_ref, rtol=1e-5, atol=0)
        for x, y in zip(pa, pbs):
            torch.testing.assert_close(x, y, rtol=1e-5, atol=1e-7)
    assert int(ours.step_count) == 25
    off = 0
    for x, y in zip(pa, pbs):
        st = ref.state[y]
        k = x.numel()
        m = torch.as_strided(ours.exp_avg[off:off + k], x.shape, x.stride())
        v = torch.as_strided(ours.exp_avg_sq[off:off + k], x.shape, x.stride())
        torch.testing.assert_close(m, st["exp_avg"], rtol=1e-5, atol=1e-6)      # |g| up to ~40: sums cancel
        torch.testing.assert_close(v, st["exp_avg_sq"], rtol=1e-5, atol=1e-7)
        off += k
    # graph replay: fixed gradient addresses, device-resident step counter
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            ours.clip_and_step(max_norm)
    torch.cuda.current_stream(dev).wait_stream(side)
    w_before = pa[0].detach().clone()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    assert int(ours.step_count) == 28 and not torch.equal(pa[0], w_before)


def test_evaluator_channels_last_observations_give_the_same_episodes(pb):
    """EpisodeEvaluator hands a fused QNetV2 its observations channels-last and padded to 32
    channels straight from the encoder (pb_encode_obs_nhwc32).  The same network behind a policy
    object that hides what it is gets the channel-major observation and converts it itself: the
    bits that reach the first convolution are the same, so every episode must be identical."""
    import torch

    from pacbot_rl_b200 import models, train_dqn
    from pacbot_rl_b200.policies import MaxQPolicy

    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    net = models.QNetV2(pb.OBS_SHAPE, 5).to(dev).use_fused_blocks().eval()
    with torch.no_grad():
        for head in (net.value_head, net.advantage_head):
            for p in head.parameters():
                p.normal_(0.0, 0.5)

    class Opaque:                        # no `q_net` attribute: the evaluator keeps channel-major buffers
        def actions_u8(self, obs, mask):
            assert tuple(obs.shape[1:]) == (17, 28, 31)
            return net.greedy_actions(obs, mask)

    direct = train_dqn.run_eval_episodes(MaxQPolicy(net), dev, 24, 150, seed=5, env_id_base=40)
    via = train_dqn.run_eval_episodes(Opaque(), dev, 24, 150, seed=5, env_id_base=40)
    assert set(direct) == set(via)
    for k in direct:
        assert torch.equal(direct[k], via[k]), k
    assert int(direct["steps"].max()) > 20
