"""CPU tests against tests/golden/models_v1.json -- OUTPUTS OF THE REFERENCE ITSELF: the fixture
was produced by importing /root/reference/src/{models,policies,utils}.py in the build container
and parsing the hyper-parameter tables of its training scripts (tools/make_golden_models.py).
Pins SURVEY.md 8(a) rows a15 (MaxQPolicy), a16 (QNetV2, and NetV2 of row f4) and the
hyper-parameter defaults of a17 / f4.  Floating point: same torch build on CPU, still compared with
an explicit tolerance of 1e-5 relative (written here) rather than bitwise."""
import json
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
from make_golden_models import OBS_SHAPE, det_masks, det_obs, det_params  # noqa: E402  (no reference import at module level)

RTOL = 1e-5


@pytest.fixture(scope="module")
def fx():
    with open(os.path.join(ROOT, "tests", "golden", "models_v1.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def models(pb):
    from pacbot_rl_b200 import models as m

    return m


@pytest.mark.parametrize("cls,out_dim", [("QNetV2", 5), ("NetV2", 6), ("QNet", 5), ("DebugMLPQNet", 5)])
def test_network_matches_reference_module(models, fx, cls, out_dim):
    g = fx[cls]
    torch.set_num_threads(1)
    torch.manual_seed(g["init_seed"])
    net = getattr(models, cls)(OBS_SHAPE, out_dim)
    # same parameter names, order and shapes => state dicts and checkpoints interchange
    assert [[k, list(v.shape)] for k, v in net.state_dict().items()] == g["keys"]
    assert sum(p.numel() for p in net.parameters()) == g["num_params"]
    if cls in ("QNetV2", "NetV2"):
        assert g["num_params"] == 156934
    # same initialisation under the same seed (orthogonal, heads / last layer zero; the debugging
    # MLP keeps torch's default init, which consumes the generator in the same order)
    sums = [float(p.detach().double().sum()) for p in net.parameters()]
    abs_sums = [float(p.detach().double().abs().sum()) for p in net.parameters()]
    np.testing.assert_allclose(sums, g["init_sums"], rtol=RTOL, atol=1e-6)
    np.testing.assert_allclose(abs_sums, g["init_abs_sums"], rtol=RTOL, atol=1e-6)
    # same function: deterministic weights and inputs, outputs from the reference's module
    det_params(net, seed=7)
    net.eval()
    with torch.no_grad():
        out = net(det_obs(6, seed=11))
    np.testing.assert_allclose(out.double().numpy(), np.array(g["det_forward"]), rtol=RTOL, atol=1e-6)
    # the NHWC mode the trainers use computes the same function
    if hasattr(net, "use_channels_last"):
        with torch.no_grad():
            out_cl = net.use_channels_last()(det_obs(6, seed=11))
        np.testing.assert_allclose(out_cl.double().numpy(), np.array(g["det_forward"]), rtol=1e-4, atol=1e-5)
    if cls == "QNetV2":
        # ... and so does the zero-padded 32-channel input of ReplayBuffer(pad_channels=True)
        x = torch.nn.functional.pad(det_obs(6, seed=11), (0, 0, 0, 0, 0, 15))
        with torch.no_grad():
            out_pad = net(x)
        np.testing.assert_allclose(out_pad.double().numpy(), np.array(g["det_forward"]), rtol=1e-4, atol=1e-5)


def test_maxq_policy_matches_reference(models, fx, pb):
    from pacbot_rl_b200.policies import EpsilonGreedy, MaxQPolicy

    net = models.QNetV2(OBS_SHAPE, 5)
    det_params(net, seed=7)
    net.eval()
    obs, masks = det_obs(24, seed=17), det_masks(24, seed=13)
    pol = MaxQPolicy(net)
    assert pol(obs, masks).tolist() == fx["QNetV2"]["maxq_actions"]
    assert len(set(fx["QNetV2"]["maxq_actions"])) >= 4   # the fixture exercises the masking
    # EpsilonGreedy (policies.py:16-41): greedy iff rand > epsilon.  epsilon = 0 is always greedy;
    # epsilon = 1 never is, and then only valid actions are drawn
    assert EpsilonGreedy(pol, 5, 0.0)(obs, masks).tolist() == fx["QNetV2"]["maxq_actions"]
    torch.manual_seed(0)
    eg = EpsilonGreedy(pol, 5, 1.0)
    seen = set()
    for _ in range(50):
        a = eg(obs, masks)
        assert bool(masks[torch.arange(24), a].all())
        seen.update(a.tolist())
    assert seen == {0, 1, 2, 3, 4}


def test_lerp_and_hyperparameter_defaults(fx, pb):
    from pacbot_rl_b200 import train_alphazero, train_dqn

    for a, b, t, want in fx["lerp"]:
        assert train_dqn.lerp(a, b, t) == want
    for mod, key in ((train_dqn, "hyperparams_dqn"), (train_alphazero, "hyperparams_alphazero")):
        ours = [[k, v] for k, v in mod.hyperparam_defaults.items()]
        assert ours == fx[key], key
        # and they are what the command line parser hands out
        ns = mod.build_parser().parse_args([])
        for k, v in fx[key]:
            assert getattr(ns, k) == v and type(getattr(ns, k)) is type(v), k


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree not present")
def test_live_reference_checkpoint_round_trip(models, tmp_path):
    """Where the reference tree exists (the build container): a checkpoint written the way
    train_dqn.save_checkpoint writes `q_net-latest.safetensors` loads STRICTLY into the reference's
    own `models.QNetV2` and computes the same Q-values, and the other way round (row f3)."""
    import importlib.util

    import safetensors.torch

    spec = importlib.util.spec_from_file_location("_ref_models", "/root/reference/src/models.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    ours = models.QNetV2(OBS_SHAPE, 5).use_channels_last()
    det_params(ours, seed=3)
    assert not ours.backbone[0].weight.is_contiguous()          # NHWC weights, as the trainers keep them
    # what save_checkpoint pickles: a row-major copy.  The reference's export script
    # (to_safetensors.py: torch.load + save_file(q_net.state_dict())) refuses non-contiguous tensors
    pt = str(tmp_path / "q_net-latest.pt")
    torch.save(models.portable_copy(ours), pt)
    loaded = torch.load(pt, map_location="cpu", weights_only=False)
    assert all(v.is_contiguous() for v in loaded.state_dict().values())
    path = str(tmp_path / "q_net-latest.safetensors")
    safetensors.torch.save_file(loaded.state_dict(), path)      # the line of to_safetensors.py
    theirs = ref.QNetV2(OBS_SHAPE, 5)
    theirs.load_state_dict(safetensors.torch.load_file(path), strict=True)
    obs = det_obs(5, seed=23)
    with torch.no_grad():
        np.testing.assert_allclose(ours.eval()(obs).numpy(), theirs.eval()(obs).numpy(), rtol=1e-4, atol=1e-5)
    back = models.QNetV2(OBS_SHAPE, 5)
    back.load_state_dict(theirs.state_dict(), strict=True)
    with torch.no_grad():
        np.testing.assert_allclose(back.eval()(obs).numpy(), theirs(obs).numpy(), rtol=RTOL, atol=1e-6)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree not present")
def test_live_reference_epsilon_greedy_distribution(pb):
    """EpsilonGreedy is random on both sides (the reference draws with Python's `random` and
    `torch.rand`, policies.py:29-39), so the comparison is statistical: per mask pattern the
    empirical action distribution of the reference's class, run live, and of ours agree within
    5 sigma of the binomial noise, and both match the closed form
    P(a) = (1 - eps) * [a == greedy] + eps * valid(a) / n_valid."""
    import importlib.util
    import random

    from pacbot_rl_b200.policies import EpsilonGreedy

    sys.path.insert(0, "/root/reference/src")   # policies.py imports `models` by bare name
    try:
        spec = importlib.util.spec_from_file_location("_ref_policies", "/root/reference/src/policies.py")
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
    finally:
        sys.path.remove("/root/reference/src")
    masks = torch.tensor([[1, 1, 0, 0, 1], [1, 0, 0, 0, 0], [1, 1, 1, 1, 1], [1, 0, 1, 1, 0]], dtype=torch.bool)
    greedy = torch.tensor([4, 0, 2, 3])
    eps, reps = 0.3, 4000
    big_masks, big_greedy = masks.repeat(reps, 1), greedy.repeat(reps)

    class Fixed:   # stands in for MaxQPolicy: the greedy action travels in the "observation", so it
        def __call__(self, o, m):   # works for the greedy subset (reference) and the full batch (ours)
            return o[:, 0].long()

    obs = big_greedy.float().unsqueeze(1)
    random.seed(1)
    torch.manual_seed(1)
    a_ref = ref.EpsilonGreedy(Fixed(), 5, eps)(obs, big_masks)
    a_our = EpsilonGreedy(Fixed(), 5, eps)(obs, big_masks)
    for k in range(len(masks)):
        valid = masks[k].float()
        want = eps * valid / valid.sum()
        want[greedy[k]] += 1 - eps
        for a in (a_ref, a_our):
            freq = torch.bincount(a[k::len(masks)], minlength=5).float() / reps
            sigma = torch.sqrt(want * (1 - want) / reps) + 1e-9
            assert bool(((freq - want).abs() <= 5 * sigma + 1e-12).all()), (k, freq, want)


def test_global_max_pool_is_adaptive_max_pool(models):
    """GlobalMaxPool2d (one max-pool window over the map) vs nn.AdaptiveMaxPool2d((1, 1)), the
    module the reference uses (models.py:96): same values and the same gradient routing, including
    maps with tied maxima (first maximum in row-major order) -- NCHW and channels-last."""
    torch.manual_seed(3)
    x = torch.randn(5, 7, 4, 4)
    x[0, 0] = 1.5                      # every cell tied
    x[1, 2, 1, 2] = x[1, 2, 3, 0] = 9.0   # two tied maxima
    for fmt in (torch.contiguous_format, torch.channels_last):
        a = x.clone().contiguous(memory_format=fmt).requires_grad_(True)
        b = x.clone().contiguous(memory_format=fmt).requires_grad_(True)
        ya = torch.nn.AdaptiveMaxPool2d((1, 1))(a)
        yb = models.GlobalMaxPool2d()(b)
        assert torch.equal(ya, yb) and yb.shape == (5, 7, 1, 1)
        w = torch.randn_like(ya)
        (ya * w).sum().backward()
        (yb * w).sum().backward()
        assert torch.equal(a.grad, b.grad)
