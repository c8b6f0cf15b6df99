"""GPU parity of the batched MCTS / AlphaZero collector (pb_mcts_* behind BatchedMCTS) against
the scalar oracle (oracle/mcts_oracle.c, a restatement of pacbot_rs/src/mcts.rs + alphazero.rs).
Both sides get the same evaluator outputs, the same root noise and the same counter-RNG draws;
tree statistics, rewards and experience items must agree BIT FOR BIT."""
import numpy as np
import pytest
import torch

from conftest import assert_obs_equal, assert_states_equal
from mcts_common import NoiseSource, np_evaluator, uniform_evaluator

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pymcts(oracle):
    from oracle import pymcts

    pymcts.lib()
    return pymcts


def torch_eval(fn):
    def run(obs, mask):
        v, p = fn(obs.cpu().numpy(), mask.cpu().numpy())
        return torch.from_numpy(v).to(obs.device), torch.from_numpy(p).to(obs.device)

    return run


def bits(x):
    return np.ascontiguousarray(np.asarray(x, np.float32)).view(np.uint32)


def compare_trees(om, gm, n, ctx):
    gv = {
        "root_visits": gm.root_visits().cpu().numpy(),
        "node_count": gm.node_count().cpu().numpy(),
        "max_depth": gm.max_depth().cpu().numpy(),
        "best_action": gm.best_action().cpu().numpy(),
        "value": gm.value().cpu().numpy(),
        "action_distribution": gm.action_distribution().cpu().numpy(),
        "policy_prior": gm.policy_prior().cpu().numpy(),
        "action_values": gm.action_values().cpu().numpy(),
    }
    for i in range(n):
        assert om.root_visits(i) == gv["root_visits"][i], f"{ctx} tree {i} root visits"
        assert om.node_count(i) == gv["node_count"][i], f"{ctx} tree {i} node count"
        assert om.max_depth(i) == gv["max_depth"][i], f"{ctx} tree {i} max depth"
        assert np.array_equal(bits(om.policy_prior(i)), bits(gv["policy_prior"][i])), f"{ctx} tree {i} priors"
        assert bits(om.value(i)) == bits(gv["value"][i]), f"{ctx} tree {i} value"
        assert np.array_equal(bits(om.action_values(i)), bits(gv["action_values"][i])), f"{ctx} tree {i} Q"
        if om.root_visits(i) > 1:
            assert np.array_equal(bits(om.action_distribution(i)), bits(gv["action_distribution"][i])), \
                f"{ctx} tree {i} visit distribution"
        assert om.best_action(i) == gv["best_action"][i], f"{ctx} tree {i} best action"


@pytest.mark.parametrize("pool_factor", ["8", "1"], ids=["pool8x", "pool1x"])
@pytest.mark.parametrize("evaluator", [np_evaluator, uniform_evaluator], ids=["hashed", "uniform"])
def test_search_take_action_and_reset_match_oracle(pb, pymcts, evaluator, pool_factor, monkeypatch):
    # pool1x: every re-root compacts the kept subtree into the other pool; pool8x: re-rooting only
    # moves the root index and compaction is rare; the tree statistics must not care
    monkeypatch.setenv("PB_MCTS_POOL_FACTOR", pool_factor)
    n, seed, base, iters_per_move, moves = 24, 11, 3000, 30, 14
    rs = (np.arange(n) % 3 != 0)
    noise = NoiseSource(n)
    om = pymcts.OracleMCTS(n, rs.astype(np.uint8), 0, seed, base, evaluator, noise.oracle)
    env = pb.BatchedPacmanGym(n, rs, False, device="cuda:0", seed=seed, env_id_base=base)
    gm = pb.mcts.BatchedMCTS(env, torch_eval(evaluator), moves * iters_per_move + 2, noise_fn=noise.torch_fn)
    om.reset()
    gm.reset()
    assert_states_equal(om.export_state(), env.get_state(), "after reset")
    compare_trees(om, gm, n, "fresh roots")
    finished = 0
    for move in range(moves):
        for it in range(iters_per_move):
            om.search_iteration()
            gm.search_iteration()
            if it % 10 == 3:
                compare_trees(om, gm, n, f"move {move} iteration {it}")
        compare_trees(om, gm, n, f"move {move} before acting")
        actions = gm.best_action().to(torch.uint8)
        # every third tree plays its LEAST likely legal action instead: exercises the
        # "child absent -> fresh root" branch (mcts.rs:164-166)
        mask = env.action_mask().cpu().numpy()
        a = actions.cpu().numpy().copy()
        for i in range(0, n, 3):
            dist = gm.action_distribution()[i].cpu().numpy()
            legal = np.nonzero(mask[i])[0]
            a[i] = legal[np.argmin(dist[legal])]
        o_rew, o_done = om.take_action(a)
        g_rew, g_done, status = gm.take_action(torch.from_numpy(a))
        assert np.array_equal(o_rew, g_rew.cpu().numpy()), f"move {move} rewards"
        assert np.array_equal(o_done, g_done.cpu().numpy()), f"move {move} done"
        assert_states_equal(om.export_state(), env.get_state(), f"move {move} after take_action")
        if o_done.any():  # MCTSContext::reset for the finished games
            finished += int(o_done.sum())
            om.reset(o_done.astype(np.uint8))
            gm.reset(torch.from_numpy(o_done).cuda())
        compare_trees(om, gm, n, f"move {move} after acting")
    assert gm.check() == om.epoch == moves * iters_per_move
    assert np.array_equal(noise.calls_oracle, noise.calls_gpu)


def test_ponder_grows_each_tree_to_the_requested_size(pb, pymcts):
    n, seed, base, size = 9, 4, 10, 37
    noise = NoiseSource(n)
    om = pymcts.OracleMCTS(n, 1, 0, seed, base, np_evaluator, noise.oracle)
    env = pb.BatchedPacmanGym(n, True, False, device="cuda:0", seed=seed, env_id_base=base)
    gm = pb.mcts.BatchedMCTS(env, torch_eval(np_evaluator), size, noise_fn=noise.torch_fn)
    om.reset()
    gm.reset()
    for _ in range(size - 1):   # ponder_and_choose: max_tree_size - node_count iterations
        om.search_iteration()
    best = gm.ponder(size)
    compare_trees(om, gm, n, "after ponder")
    assert (gm.node_count() <= size).all() and int(gm.root_visits().min()) == size
    assert np.array_equal(best.cpu().numpy(), [om.best_action(i) for i in range(n)])
    # a second ponder to the same size is a no-op (saturating_sub, mcts.rs:209)
    gm.ponder(size)
    compare_trees(om, gm, n, "after a no-op ponder")
    with pytest.raises(ValueError):
        gm.ponder(size + 1)


def test_builtin_root_noise_is_dirichlet(pb, n=6000):
    """pb_mcts_root_noise_dirichlet (the built-in root noise, drawn on the device): what reaches the
    root prior is `0.75 prior + 0.25 noise` with noise ~ Dirichlet(11 / num_valid) over the valid
    actions (mcts.rs:259-281): on the simplex, zero on invalid actions, component mean 1 / k and
    variance (1/k)(1 - 1/k) / (alpha + 1); different per tree and per call."""
    env = pb.BatchedPacmanGym(n, True, False, device="cuda:0", seed=21, env_id_base=5)
    gm = pb.mcts.BatchedMCTS(env, torch_eval(uniform_evaluator), 8)        # noise_fn=None: built in
    gm.reset()                                                             # reset_root -> prior + noise
    mask = env.action_mask().cpu().numpy().astype(bool)
    nv = mask.sum(1)
    uniform = mask / nv[:, None]
    prior1 = gm.policy_prior().cpu().numpy().astype(np.float64)
    noise1 = (prior1 - 0.75 * uniform) / 0.25
    gm.add_root_noise()                                                    # a second draw on top
    prior2 = gm.policy_prior().cpu().numpy().astype(np.float64)
    noise2 = (prior2 - 0.75 * prior1) / 0.25
    for noise in (noise1, noise2):
        assert np.abs(noise[~mask]).max() < 1e-6 and noise.min() > -1e-6
        assert np.abs(noise.sum(1) - 1.0).max() < 1e-5
    assert np.abs(noise1 - noise2)[mask & (nv[:, None] > 1)].mean() > 0.05   # per call
    alpha = 11.0
    checked = 0
    for k in range(3, 6):                 # Stay is always valid: 3 (corridor) .. 5 (crossing)
        rows = np.nonzero(nv == k)[0]
        if len(rows) < 300:
            continue
        comp = np.concatenate([noise1[rows][mask[rows]].reshape(len(rows), k),
                               noise2[rows][mask[rows]].reshape(len(rows), k)])
        var = (1 / k) * (1 - 1 / k) / (alpha + 1)
        sem = (var / len(comp)) ** 0.5
        assert np.abs(comp.mean(0) - 1 / k).max() < 5 * sem, (k, comp.mean(0))
        assert np.abs(comp.var(0) / var - 1).max() < 0.2, (k, comp.var(0), var)
        assert len(np.unique(np.round(comp[:, 0], 6))) > 0.95 * len(comp)     # per tree
        checked += 1
    assert checked >= 1 and (nv == 3).sum() >= 300      # corridors dominate the maze


def test_search_from_a_finished_env_is_reported(pb):
    env = pb.BatchedPacmanGym(2, False, False, device="cuda:0")
    gm = pb.mcts.BatchedMCTS(env, torch_eval(uniform_evaluator), 8)
    gm.reset()
    st = env.get_state()
    st["curr_lives"][1] = 2          # is_done (env.rs:277-279)
    env.set_state(st)
    gm.search_iteration()
    with pytest.raises(ValueError, match="finished env"):
        gm.check()
    with pytest.raises(RuntimeError):
        gm.take_action(torch.zeros(2, dtype=torch.uint8))


def test_copy_envs_batched_clone(pb):
    src = pb.BatchedPacmanGym(64, True, True, device="cuda:0", seed=8)
    src.reset()
    for k in range(5):
        src.step_random(call_idx=k)
    dst = pb.BatchedPacmanGym(200, False, False, device="cuda:0", seed=1)
    before = dst.get_state()
    si = torch.tensor([3, 7, 63, 0, 5], device="cuda:0")
    di = torch.tensor([199, 0, 17, -1, 100], device="cuda:0")   # -1: skipped
    dst.copy_envs_from(src, src_index=si, dst_index=di)
    after, s = dst.get_state(), src.get_state()
    for a, b in [(199, 3), (0, 7), (17, 63), (100, 5)]:
        assert after[a:a + 1].tobytes() == s[b:b + 1].tobytes()
    untouched = np.setdiff1d(np.arange(200), [199, 0, 17, 100])
    assert np.array_equal(after[untouched], before[untouched])
    out = torch.empty((3, 17, 28, 31), device="cuda:0")
    dst.obs_range(17, 1, out[:1])
    assert torch.equal(out[0], src.obs()[63])


def test_collector_matches_oracle(pb, pymcts):
    P, tree, L, seed, base = 8, 16, 12, 23, 700
    noise = NoiseSource(P)
    oc = pymcts.OracleCollector(tree, L, 0.99, P, seed, base, np_evaluator, noise.oracle)
    cfg = pb.alphazero.AlphaZeroConfig(tree, L, 0.99, P)
    gc = pb.alphazero.BatchedExperienceCollector(torch_eval(np_evaluator), cfg, device="cuda:0", seed=seed,
                                                 env_id_base=base, noise_fn=noise.torch_fn)
    total = 0
    for call in range(6):
        items = oc.generate_experience()
        batch = gc.generate_experience()
        assert len(batch) == len(items), f"call {call}: number of items"
        total += len(items)
        assert np.array_equal(items["action_mask"].astype(bool), batch.action_mask.cpu().numpy()), f"call {call} masks"
        assert np.array_equal(bits(items["value"]), bits(batch.value.cpu().numpy())), f"call {call} values"
        assert np.array_equal(bits(items["action_distribution"]),
                              bits(batch.action_distribution.cpu().numpy())), f"call {call} distributions"
        assert_obs_equal(items["obs"], batch.obs.cpu().numpy(), f"call {call} observations")
        # the compact replay form: the snapshot states re-encode to the same observations
        out = torch.empty_like(batch.obs)
        stage = pb.BatchedPacmanGym(len(batch), False, False, device="cuda:0")
        stage.copy_envs_from(batch.state_engine, src_index=batch.state_slots)
        assert torch.equal(stage.obs(out), batch.obs)
    assert total >= 6
    assert_states_equal(oc.export_state(), gc.env.get_state(), "collector envs at the end")
    assert np.array_equal(noise.calls_oracle, noise.calls_gpu)


def test_mcts_context_drop_in_surface(pb):
    """MCTSContext with the reference's NumPy evaluator (pacbot_rs.pyi:26-41), as
    train_alphazero.py:119-131 drives it."""
    calls = []

    def evaluator(obs, mask):
        assert isinstance(obs, np.ndarray) and obs.dtype == np.float32 and obs.shape[1:] == (17, 28, 31)
        assert isinstance(mask, np.ndarray) and mask.dtype == np.bool_ and mask.shape[1:] == (5,)
        calls.append(obs.shape[0])
        return uniform_evaluator(obs, mask)

    gym = pb.PacmanGym(random_start=True, random_ticks=False, device="cuda:0", seed=2)
    gym.reset()
    mc = pb.mcts.MCTSContext(gym, evaluator)
    assert mc.node_count() == 1 and mc.max_depth() == 0 and calls
    with pytest.raises(RuntimeError):
        mc.action_distribution()        # assert!(self.root.visit_count > 1)
    mc.reset()
    steps = 0
    while steps < 5:
        action = mc.ponder_and_choose(24)
        assert mc.node_count() == 24 and isinstance(action, int) and mc.env.action_mask()[action]
        assert abs(sum(mc.action_distribution()) - 1.0) < 1e-6
        assert len(mc.policy_prior()) == 5 and len(mc.action_values()) == 5
        assert isinstance(mc.value(), float) and mc.max_depth() >= 1
        assert mc.root_obs_numpy().shape == (17, 28, 31)
        score_before = mc.env.score()
        reward, done = mc.take_action(action)
        assert isinstance(reward, int) and isinstance(done, bool)
        assert mc.env.score() >= score_before
        steps += 1
        if done:
            break
    # the caller's env object was copied, not captured (pyo3 extracts PacmanGym by value)
    assert gym.score() == 0
    with pytest.raises(ValueError):
        mc.take_action(7)
    mc.evaluator = uniform_evaluator    # #[pyo3(get, set)] evaluator
    mc.ponder_and_choose(30)


def test_experience_collector_drop_in_surface(pb):
    cfg = pb.alphazero.AlphaZeroConfig(tree_size=10, max_episode_length=6, discount_factor=0.99,
                                       num_parallel_envs=4)
    assert "tree_size: 10" in repr(cfg)
    col = pb.alphazero.ExperienceCollector(uniform_evaluator, cfg, device="cuda:0", seed=1)
    items = col.generate_experience()
    assert col.config is cfg and len(items) >= 1
    it = items[0]
    assert it.obs.shape == (17, 28, 31) and it.obs.dtype == np.float32
    assert len(it.action_mask) == 5 and all(isinstance(b, bool) for b in it.action_mask)
    assert isinstance(it.value, float) and len(it.action_distribution) == 5


def test_collector_cuda_graph_matches_eager(pb):
    """The captured collector step (search iteration + env steps + root noise in one CUDA graph)
    must leave envs, trees and episode buffers exactly where eager execution leaves them."""
    # the built-in root noise is drawn on the device from the counter RNG, keyed by (seed, tree,
    # draws taken by that tree): both collectors see the same Dirichlet samples
    def evaluator(obs, mask):
        value = (obs[:, 1] != 0).sum(dim=(1, 2)).to(torch.float32) * 0.25 - obs[:, 2].flatten(1).argmax(1) % 7
        m = mask.to(torch.float32) * (1.0 + obs[:, 2].flatten(1).argmax(1)[:, None] % 3)
        return value, m / m.sum(dim=1, keepdim=True)

    cfg = pb.alphazero.AlphaZeroConfig(20, 15, 0.99, 32)
    cols = [pb.alphazero.BatchedExperienceCollector(evaluator, cfg, device="cuda:0", seed=3, env_id_base=50)
            for _ in range(2)]
    cols[1].enable_cuda_graph()           # = 3 ordinary steps, then capture
    fin = [[], cols[1]._pending]
    for _ in range(3):
        cols[0].generate_experience_step(fin[0])
    for step in range(330):
        for c, f in zip(cols, fin):
            c.generate_experience_step(f)
        if step % 30 == 29:
            a, b = cols
            assert_states_equal_gpu = a.env.get_state().tobytes() == b.env.get_state().tobytes()
            assert assert_states_equal_gpu, f"env states differ at step {step}"
            assert torch.equal(a.remaining, b.remaining) and torch.equal(a.ep_len, b.ep_len)
            assert torch.equal(a.mcts.root_visits(), b.mcts.root_visits())
            assert torch.equal(a.mcts.node_count(), b.mcts.node_count())
            assert torch.equal(a.mcts.value(), b.mcts.value())
            assert torch.equal(a.mcts.policy_prior(), b.mcts.policy_prior())
            assert torch.equal(a.ep_value, b.ep_value) and torch.equal(a.ep_dist, b.ep_dist)
    assert len(fin[0]) == len(fin[1]) and len(fin[0]) > 0
    for x, y in zip(fin[0], fin[1]):
        assert all(torch.equal(p, q) for p, q in zip(x, y))
    assert cols[0].mcts.check() == cols[1].mcts.check()


def test_alphazero_trainer_short_run(pb, tmp_path):
    """train_alphazero mirror end to end: collect with MCTS, state-replay, encode-on-sample,
    two optimiser iterations, greedy evaluation, checkpoint in the reference's format."""
    from pacbot_rl_b200 import train_alphazero as ta

    args = ta.build_parser().parse_args([
        "--device", "cuda:0", "--num_iters", "2", "--batch_size", "32", "--train_instances_per_iter", "64",
        "--num_parallel_envs", "8", "--experience_steps", "16", "--experience_buffer_size", "64",
        "--tree_size", "8", "--max_episode_length", "10", "--eval_envs", "4",
        "--checkpoint-dir", str(tmp_path / "ckpt"), "--metrics-file", str(tmp_path / "m.jsonl")])
    tr = ta.AlphaZeroTrainer(args)
    assert sum(p.numel() for p in tr.model.parameters()) == 156_934 - 256 - 1 - 5 * 256 - 5 + 6 * 256 + 6
    w0 = tr.model.network[0].weight.detach().clone()
    tr.train()
    assert tr.has_model_updated and len(tr.buffer) == 64
    assert not torch.equal(w0, tr.model.network[0].weight)
    rows = [__import__("json").loads(l) for l in open(tmp_path / "m.jsonl")]
    assert len(rows) == 2 and all(np.isfinite(r["avg_loss"]) for r in rows)
    assert rows[0]["eval_episode_steps"] >= 1
    model = torch.load(tmp_path / "ckpt" / "model-latest.pt", weights_only=False)
    assert sorted(k for k in model.state_dict() if k.endswith("weight")) == [
        f"network.{i}.weight" for i in (0, 11, 15, 17, 2, 5, 8)]
    # sampled observations are exactly the encoder's output for the stored states
    obs, mask, value, dist = tr.buffer.sample(32)
    assert obs.shape == (32, 17, 28, 31) and torch.isfinite(obs).all()
    assert (dist[~mask] == 0).all() and torch.allclose(dist.sum(1), torch.ones(32, device="cuda:0"), atol=1e-5)


def test_nhwc32_leaf_layout_gives_the_same_search(pb):
    """BatchedMCTS(leaf_layout="nhwc32"): the encoder writes leaf observations channels-last and
    zero-padded to 32 channels (pb_encode_obs_nhwc32) and NetV2.policy_value takes them as they
    are.  (1) That tensor is bit for bit what the layout-conversion launch makes of obs();
    (2) two collectors that differ only in the leaf layout build identical trees and experience."""
    from pacbot_rl_b200 import alphazero, models

    dev = torch.device("cuda:0")
    env = pb.BatchedPacmanGym(300, True, True, device=dev, seed=4)
    env.reset()
    for k in range(7):
        env.step_random(call_idx=k)
    direct = env.obs_nhwc32()
    via = models.obs_to_padded_nhwc(env.obs())
    assert direct.shape == via.shape == (300, 32, 28, 31) and direct.stride() == via.stride()
    assert torch.equal(direct.view(torch.int32), via.view(torch.int32))
    part = torch.full((40, 32, 28, 31), -1.0, device=dev).contiguous(memory_format=torch.channels_last)
    env.obs_nhwc32(out=part, first=211, count=40)
    assert torch.equal(part, via[211:251])
    with pytest.raises(ValueError):
        env.obs_nhwc32(out=torch.empty((300, 32, 28, 31), device=dev))     # not channels_last

    torch.manual_seed(0)
    model = models.NetV2((17, 28, 31), 6).to(dev).use_fused_blocks().eval()
    with torch.no_grad():                    # the final layer starts at zero: make the outputs vary
        for p in model.network[17].parameters():
            p.normal_(0.0, 0.3)

    @torch.no_grad()
    def evaluator(obs, mask):
        return model.policy_value(obs, mask, 100.0)

    cfg = alphazero.AlphaZeroConfig(8, 6, 0.99, 48)     # episodes are cut after 6 env steps: hand-outs within 100 iterations
    cols = [alphazero.BatchedExperienceCollector(evaluator, cfg, device=dev, seed=9, leaf_layout=layout)
            for layout in ("nchw", "nhwc32")]
    batches = []
    for col in cols:
        finished = []
        for _ in range(100):
            col.generate_experience_step(finished)
        col.mcts.check()
        assert finished
        batches.append((col.mcts.node_count().cpu(), col.mcts.root_visits().cpu(), col.mcts.value().cpu(),
                        torch.cat([f[1] for f in finished]).cpu(), torch.cat([f[3] for f in finished]).cpu()))
    assert len(batches[0][3]) > 0
    for a, b in zip(*batches):
        assert torch.equal(a, b)


def test_fused_collector_tail_equals_the_three_launches(pb):
    """pb_mcts_backprop_collect (backprop + env steps + built-in Dirichlet root noise in one
    launch) against the three separate launches it replaces: two collectors that differ only in
    `fused_tail`, same seed, 200 steps through several env steps, re-roots, episode ends and
    resets: tree statistics, priors (noise), budgets, episode records and hand-outs identical."""
    from pacbot_rl_b200 import alphazero

    dev = torch.device("cuda:0")

    def evaluator(obs, mask):          # deterministic, input dependent, no network
        flat = obs.reshape(obs.shape[0], -1)
        value = flat[:, 1000:1400].sum(dim=1) * 0.01 - 0.3
        m = mask.to(torch.float32)
        pol = m * (1.0 + flat[:, 2000:2005].abs())
        return value, pol / pol.sum(dim=1, keepdim=True)

    cfg = alphazero.AlphaZeroConfig(10, 8, 0.99, 96)
    out = []
    for fused in (True, False):
        col = alphazero.BatchedExperienceCollector(evaluator, cfg, device=dev, seed=31, env_id_base=5)
        col.fused_tail = fused
        finished = []
        for _ in range(200):
            col.generate_experience_step(finished)
        col.mcts.check()
        assert finished
        out.append(dict(
            nodes=col.mcts.node_count().cpu(), visits=col.mcts.root_visits().cpu(), value=col.mcts.value().cpu(),
            prior=col.mcts.policy_prior().cpu(), remaining=col.remaining.cpu(), ep_len=col.ep_len.cpu(),
            ep_value=col.ep_value.cpu(), ep_dist=col.ep_dist.cpu(), state=col.env.get_state().tobytes(),
            handed=torch.cat([f[1] for f in finished]).cpu(), dist=torch.cat([f[3] for f in finished]).cpu()))
    a, b = out
    for k in a:
        same = a[k] == b[k] if isinstance(a[k], bytes) else torch.equal(a[k], b[k])
        assert same, k
