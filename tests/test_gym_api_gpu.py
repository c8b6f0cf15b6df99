"""GPU tests of the reference-facing API: the per-env `PacmanGym` drop-in (pacbot_rs.pyi:6-19),
the host-buffer C-ABI entry, and the on-device policy / collate helpers."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_pacman_gym_drop_in_surface(pb, oracle):
    """Same call sequence as src/replay_buffer.py / train_dqn.py make on pacbot_rs.PacmanGym."""
    gym = pb.PacmanGym(random_start=True, random_ticks=True)
    gym.reset()
    obs = gym.obs_numpy()
    assert isinstance(obs, np.ndarray) and obs.shape == (17, 28, 31) and obs.dtype == np.float32
    assert obs is not gym.obs_numpy()                      # fresh array per call (env.rs:305-307)
    mask = gym.action_mask()
    assert isinstance(mask, list) and len(mask) == 5 and mask[0] is True
    out = gym.step(0)
    assert isinstance(out, tuple) and isinstance(out[0], int) and isinstance(out[1], bool)
    assert isinstance(gym.score(), int) and isinstance(gym.lives(), int)
    assert isinstance(gym.is_done(), bool) and isinstance(gym.first_ai_done(), bool)
    assert isinstance(gym.remaining_pellets(), int)
    assert gym.random_start is True
    gym.random_start = False
    assert gym.random_start is False
    gym.reset()
    s = gym._batch.get_state()[0]
    assert (s["pac_row"], s["pac_col"]) == (23, 13)        # no random start any more
    with pytest.raises(ValueError, match="Invalid action"):
        gym.step(5)
    with pytest.raises(OverflowError):
        gym.step(256)
    with pytest.raises(OverflowError):
        gym.step(-1)
    with pytest.raises(TypeError):
        gym.step(1.5)
    gym.print_game_state()


def test_pacman_gym_episode_matches_oracle(pb, oracle):
    """One full greedy-random episode through the per-env API vs the oracle (RNG-free: fixed
    start, fixed ticks; the draw tape covers the frightened phases)."""
    import torch

    rng = np.random.default_rng(2)
    tape = rng.integers(0, 2**32, size=(1, 64), dtype=np.uint32)
    ob = oracle.OracleBatch(1, np.zeros(1, np.uint8), tape)
    gym = pb.PacmanGym(False, False)
    gym._batch.set_draw_tape(torch.from_numpy(tape.view(np.int32)).cuda())
    ob.reset()
    gym.reset()
    for t in range(400):
        m = ob.action_mask()[0]
        assert gym.action_mask() == m.tolist()
        a = int(rng.choice(np.flatnonzero(m)))
        r_o, d_o = ob.step(np.array([a], np.uint8))
        r_g, d_g = gym.step(a)
        assert (int(r_o[0]), bool(d_o[0])) == (r_g, d_g), t
        if t % 25 == 0 or d_g:
            assert np.array_equal(ob.obs()[0].view(np.uint32), gym.obs_numpy().view(np.uint32))
            assert gym.score() == int(ob.export_state()["curr_score"][0])
            assert gym.first_ai_done() == bool(ob.first_ai_done()[0])
        if d_g:
            assert gym.is_done()
            break
    else:
        pytest.fail("episode did not finish in 400 steps of random play")


def test_clone_and_view(pb):
    import torch

    gym = pb.PacmanGym(True, True, seed=5)
    gym.reset()
    for a in (3, 3, 1):
        gym.step(a)
    twin = gym.clone()
    assert np.array_equal(gym._batch.get_state(), twin._batch.get_state())
    gym.step(0)
    assert not np.array_equal(gym._batch.get_state(), twin._batch.get_state())
    batch = pb.BatchedPacmanGym(4, True, True, seed=1)
    batch.reset()
    v = batch.view(2)
    assert v.action_mask() == [bool(x) for x in batch.action_mask()[2].tolist()]
    assert np.array_equal(v.obs_numpy(), batch.obs()[2].cpu().numpy())
    assert v.score() == int(batch.score()[2]) and v.lives() == 3
    assert v.remaining_pellets() == int(batch.remaining_pellets()[2])
    with pytest.raises(IndexError):
        batch.view(4)
    del torch


def test_step_host_matches_device_path(pb):
    import torch

    n = 1000
    a_env = pb.BatchedPacmanGym(n, True, True, seed=9, auto_reset=True)
    b_env = pb.BatchedPacmanGym(n, True, True, seed=9, auto_reset=True)
    a_env.reset()
    b_env.reset()
    obs_a = torch.empty((n,) + pb.OBS_SHAPE, dtype=torch.float32, device="cuda")
    mask_b = torch.empty((n, 5), dtype=torch.bool, device="cuda")
    for t in range(30):
        acts = a_env.sample_actions(call_idx=t)
        r_h, d_h, m_h = a_env.step_host(acts.cpu().numpy(), obs_out=obs_a)
        r_d, d_d = b_env.step(acts, mask_out=mask_b)
        assert np.array_equal(r_h, r_d.cpu().numpy()) and np.array_equal(d_h, d_d.cpu().numpy())
        assert np.array_equal(m_h, mask_b.cpu().numpy())
        assert torch.equal(obs_a, b_env.obs())
    with pytest.raises(ValueError, match="Invalid action"):
        a_env.step_host(np.full(n, 6, np.uint8))


def test_sample_actions_are_valid_and_uniform(pb):
    import torch

    n = 20000
    env = pb.BatchedPacmanGym(n, True, True, seed=4)
    env.reset()
    mask = env.action_mask()
    counts = torch.zeros((n, 5), device="cuda")
    K = 40
    for k in range(K):
        a = env.sample_actions().long()
        assert bool(mask[torch.arange(n, device="cuda"), a].all())
        counts[torch.arange(n, device="cuda"), a] += 1
    # pooled over envs with the same number of valid actions the split must be even
    nv = mask.sum(1)
    for v in (2, 3):
        sel = nv == v
        if sel.sum() < 100:
            continue
        frac = (counts[sel] * mask[sel]).sum(0) / (sel.sum() * K)
        share = (mask[sel].float().mean(0) / v)
        assert torch.allclose(frac, share, atol=0.01), (v, frac, share)
    # host mirror of the action stream
    lib = pb._lib.load()
    a0 = env.sample_actions(call_idx=123).cpu().numpy()
    m0 = mask.cpu().numpy()
    for i in (0, 1, 77, n - 1):
        valid = np.flatnonzero(m0[i])
        k = (lib.pb_action_raw(4, i, 123) * len(valid)) >> 32
        assert a0[i] == valid[k]


def test_select_actions_matches_policies_py(pb):
    """EpsilonGreedy(MaxQPolicy) semantics (src/policies.py:24-59) on device."""
    import torch

    n = 50000
    env = pb.BatchedPacmanGym(n, True, True, seed=8)
    env.reset()
    mask = env.action_mask()
    g = torch.Generator(device="cuda").manual_seed(0)
    q = torch.randn((n, 5), device="cuda", generator=g)
    # epsilon = 0: pure masked argmax (MaxQPolicy: q[~mask] = -inf; argmax(dim=1))
    ref = q.masked_fill(~mask, -torch.inf).argmax(1)
    assert torch.equal(env.select_actions(q, 0.0).long(), ref)
    # explicit mask argument is honoured
    m2 = mask.clone()
    m2[:, 0] = False
    has = m2.any(1)
    ref2 = q.masked_fill(~m2, -torch.inf).argmax(1)
    got2 = env.select_actions(q, 0.0, mask=m2).long()
    assert torch.equal(got2[has], ref2[has])
    # ties -> lowest index (torch.argmax)
    qz = torch.zeros((n, 5), device="cuda")
    assert torch.equal(env.select_actions(qz, 0.0).long(), torch.zeros(n, dtype=torch.long, device="cuda"))
    # epsilon = 1: always a valid random action; 0 < eps < 1: greedy fraction ~ 1 - eps
    a1 = env.select_actions(q, 1.0).long()
    assert bool(mask[torch.arange(n, device="cuda"), a1].all())
    eps = 0.3
    a = env.select_actions(q, eps).long()
    assert bool(mask[torch.arange(n, device="cuda"), a].all())
    nv = mask.sum(1).float()
    expect_greedy = ((1 - eps) + eps / nv).mean().item()   # random pick may coincide with greedy
    assert abs((a == ref).float().mean().item() - expect_greedy) < 0.01


def test_gather_obs_matches_torch_indexing(pb):
    import ctypes as C

    import torch

    slots, count = 300, 512
    ring = torch.randn((slots, pb.OBS_FLOATS), device="cuda")
    idx = torch.randint(0, slots, (count,), device="cuda")
    zero = torch.rand(count, device="cuda") < 0.25
    out = torch.empty((count, pb.OBS_FLOATS), device="cuda")
    L = pb._lib.load()
    pb._lib.check(L.pb_gather_obs(C.c_void_p(ring.data_ptr()), pb.OBS_FLOATS, C.c_void_p(idx.data_ptr()),
                                  C.c_void_p(zero.view(torch.uint8).data_ptr()), count,
                                  C.c_void_p(out.data_ptr()), 0, None))
    torch.cuda.synchronize()
    ref = ring[idx]
    ref[zero] = 0
    assert torch.equal(out, ref)


def test_step_random_equals_sample_then_step(pb):
    """pb_step_random == pb_sample_actions(call_idx) followed by pb_step, bit for bit."""
    import torch

    n = 5000
    a_env = pb.BatchedPacmanGym(n, True, True, seed=21, auto_reset=True)
    b_env = pb.BatchedPacmanGym(n, True, True, seed=21, auto_reset=True)
    a_env.reset()
    b_env.reset()
    acts = torch.empty(n, dtype=torch.uint8, device="cuda")
    for t in range(60):
        r_a, d_a = a_env.step_random(call_idx=t, actions_out=acts)
        sampled = b_env.sample_actions(call_idx=t)
        r_b, d_b = b_env.step(sampled)
        assert torch.equal(acts, sampled) and torch.equal(r_a, r_b) and torch.equal(d_a, d_b)
    assert np.array_equal(a_env.get_state(), b_env.get_state())


def test_step_snapshot_matches_separate_calls(pb):
    """pb_step_snapshot_host (the per-env drop-in's one round trip per step): reward / done / info /
    mask / observation equal what the separate entry points return for the same envs and actions;
    an invalid action raises like pb_step_host and leaves that env untouched; a snapshot without
    a step changes nothing."""
    import torch

    n = 37
    mk = lambda: pb.BatchedPacmanGym(n, np.arange(n) < n // 2, True, device="cuda:0", seed=3)  # noqa: E731
    a, b = mk(), mk()
    a.reset()
    b.reset()
    rng = np.random.default_rng(0)
    for t in range(60):
        acts = rng.integers(0, 5, n, dtype=np.uint8)
        r_a, d_a, m_a = a.step_host(acts)
        snap = b.step_snapshot(acts)
        assert np.array_equal(snap["reward"], r_a) and np.array_equal(snap["done"], d_a), t
        assert np.array_equal(snap["mask"], m_a)
        info = snap["info"]
        assert np.array_equal(info[:, 0], a.score().cpu().numpy())
        assert np.array_equal(info[:, 1], a.lives().cpu().numpy())
        assert np.array_equal(info[:, 2].astype(bool), a.is_done().cpu().numpy())
        assert np.array_equal(info[:, 3].astype(bool), a.first_ai_done().cpu().numpy())
        assert np.array_equal(info[:, 4], a.remaining_pellets().cpu().numpy())
        assert np.array_equal(info[:, 5], a.ticks_per_step().cpu().numpy())
        assert np.array_equal(snap["obs"].view(np.uint32), a.obs().cpu().numpy().view(np.uint32))
        if t % 9 == 8:
            done = torch.from_numpy(d_a.copy()).to("cuda:0")
            a.reset(done)
            b.reset(done)
    before = b.get_state()
    again = b.step_snapshot(None, want_obs=False)          # no step: nothing moves
    assert b.get_state().tobytes() == before.tobytes() == a.get_state().tobytes()
    assert np.array_equal(again["info"][:, 0], a.score().cpu().numpy())
    acts = rng.integers(0, 5, n, dtype=np.uint8)
    acts[11] = 6
    with pytest.raises(ValueError, match="Invalid action"):
        b.step_snapshot(acts)
    assert b.get_state()[11].tobytes() == before[11].tobytes()      # env.rs:83-88: env untouched
    b.step_snapshot(np.zeros(n, np.uint8))                 # the record was cleared: works again
    with pytest.raises(ValueError):
        pb.BatchedPacmanGym(2000, False, False, device="cuda:0").step_snapshot(None)


def test_pacman_gym_snapshot_follows_every_mutation(pb):
    """The drop-in answers its getters from the snapshot of its last mutation: step / reset / the
    random_start setter / clone / MCTSContext.take_action must all refresh it."""
    gym = pb.PacmanGym(False, False, seed=5, env_id=0)
    gym.reset()
    p0, o0 = gym.remaining_pellets(), gym.obs_numpy()
    for a in (3, 3, 3):
        gym.step(a)
    st = gym._batch.get_state()[0]
    assert gym.score() == int(st["curr_score"]) and gym.remaining_pellets() == int(st["num_pellets"]) < p0
    assert not np.array_equal(gym.obs_numpy(), o0)
    twin = gym.clone()
    assert twin.score() == gym.score() and np.array_equal(twin.obs_numpy(), gym.obs_numpy())
    twin.step(3)
    assert gym.remaining_pellets() == int(st["num_pellets"])        # the original did not move
    gym.reset()
    assert gym.remaining_pellets() == p0 and gym.score() == 0 and np.array_equal(gym.obs_numpy(), o0)

    def evaluator(obs, mask):
        return np.zeros(len(obs), np.float32), np.full((len(obs), 5), 0.2, np.float32)

    ctx = pb.mcts.MCTSContext(gym, evaluator)
    root0 = ctx.root_obs_numpy()
    assert np.array_equal(root0, o0)
    ctx.take_action(3)
    assert not np.array_equal(ctx.root_obs_numpy(), root0)
    assert np.array_equal(ctx.root_obs_numpy(), ctx.env.obs_numpy())
    ctx.reset()
    assert np.array_equal(ctx.root_obs_numpy(), ctx.env.obs_numpy())


def test_step_snapshot_refuses_a_capturing_stream(pb):
    """pb_step_snapshot_host waits on the host for a word its kernel publishes: on a stream that
    is being captured the kernel would only be recorded, so the call must refuse instead of
    waiting forever."""
    import torch

    b = pb.BatchedPacmanGym(4, False, False, device="cuda:0", seed=2)
    b.reset()
    before = b.step_snapshot(np.zeros(4, np.uint8))["info"].copy()     # buffers exist, path works
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        with pytest.raises(ValueError, match="captured"):
            b.step_snapshot(None, want_obs=False)
    torch.cuda.synchronize()
    again = b.step_snapshot(None, want_obs=False)["info"]              # and still works afterwards
    assert np.array_equal(again, before)
