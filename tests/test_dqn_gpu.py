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
    assert len(buf) == min(steps, buf._slots) * n == int(buf._weight.sum().item())


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
    assert len(buf) == buf.capacity and int(buf._weight.sum().item()) == buf.capacity
    batch = buf.sample_batch(512)
    assert batch.obs.shape == (512, 17, 28, 31) and batch.next_obs.shape == batch.obs.shape
    assert batch.actions.dtype == torch.int64 and batch.rewards.dtype == torch.float32
    # every sampled obs is a real observation: wall channel sums to 580; done rows are zeros
    assert torch.equal(batch.obs[:, 0].sum((1, 2)), torch.full((512,), 580.0, device=dev))
    nz = batch.next_obs.reshape(512, -1).abs().sum(1)
    assert bool(((nz == 0) == batch.done).all())
    assert not bool(batch.next_action_masks[batch.done].any())
    assert bool(batch.next_action_masks[~batch.done][:, 0].all())   # Stay is always valid
    # without replacement
    flat = torch.multinomial(buf._weight.view(-1), 512, replacement=False)
    assert flat.unique().numel() == 512
    # the slot holding only the current observation is never sampled
    cur = buf._t % buf._ring
    assert float(buf._weight[cur].sum()) == 0.0
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
    w = buf._weight.cpu().numpy()
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
