"""CPU differential tests of the DEVICE MCTS kernels (pacbot-rl_b200/csrc/pb_mcts.cuh) compiled for
the host (tests/host_engine/host_mcts.cpp) against the scalar MCTS oracle (oracle/mcts_oracle.c, a
restatement of pacbot_rs/src/mcts.rs): the GPU suite's own test bodies
(test_mcts_gpu.test_search_take_action_and_reset_match_oracle, ..ponder..) are re-run with a host
stand-in for the engine (`HostEnv`), so tree statistics, rewards and env states are
compared bit for bit over search / re-root / fresh-root / compaction / reset cycles without a GPU.

The product's host code -- pacbot-rl_b200/mcts.py (`BatchedMCTS`) and alphazero.py
(`BatchedExperienceCollector`) -- runs UNMODIFIED in these tests: only the engine object is the
host stand-in `HostEnv` (the product's `BatchedPacmanGym` refuses to exist without a CUDA device,
by design) and the C calls go through `HostABI` into the host build.  Leaf observations come from
the env oracle (the encoder is GPU-only).  Test infrastructure only, like tests/test_host_engine.py;
the GPU tests remain the parity tests proper."""
import pytest

from host_stand_ins import hm, host_pb  # noqa: F401  (pytest fixtures)


@pytest.fixture(scope="module")
def pymcts(oracle):
    from oracle import pymcts

    pymcts.lib()
    return pymcts


@pytest.mark.parametrize("pool_factor", ["8", "1"], ids=["pool8x", "pool1x"])
@pytest.mark.parametrize("evaluator", ["hashed", "uniform"])
def test_device_mcts_source_matches_oracle_on_cpu(host_pb, pymcts, evaluator, pool_factor, monkeypatch):
    import test_mcts_gpu as tg
    from mcts_common import np_evaluator, uniform_evaluator

    tg.test_search_take_action_and_reset_match_oracle(
        host_pb, pymcts, np_evaluator if evaluator == "hashed" else uniform_evaluator, pool_factor, monkeypatch)


def test_device_mcts_ponder_on_cpu(host_pb, pymcts):
    import test_mcts_gpu as tg

    tg.test_ponder_grows_each_tree_to_the_requested_size(host_pb, pymcts)


def test_builtin_root_noise_is_dirichlet_on_cpu(host_pb):
    """k_mcts_root_noise_dirichlet (counter-RNG gamma sampling on the device) executed on the host."""
    import test_mcts_gpu as tg

    tg.test_builtin_root_noise_is_dirichlet(host_pb, n=3000)


def test_device_mcts_reports_search_from_finished_env(host_pb):
    import test_mcts_gpu as tg

    tg.test_search_from_a_finished_env_is_reported(host_pb)


@pytest.mark.parametrize("tpw", ["4", "32"])
def test_trees_per_warp_mapping_does_not_change_results(host_pb, pymcts, tpw, monkeypatch):
    """mcts_tree_of_thread: the lane -> tree map used when trees outnumber warps."""
    import test_mcts_gpu as tg

    monkeypatch.setenv("PB_MCTS_TREES_PER_WARP", tpw)
    tg.test_ponder_grows_each_tree_to_the_requested_size(host_pb, pymcts)


def test_product_collector_runs_over_the_host_kernels_and_matches_oracle(host_pb, pymcts):
    """test_mcts_gpu.test_collector_matches_oracle: the product's BatchedExperienceCollector
    (pacbot-rl_b200/alphazero.py, unmodified -- search budget, episode records, truncation
    bootstrap, discounted returns, hand-out order) over the host build of k_mcts_collect_act and
    friends, against the oracle's ExperienceCollector: items bit for bit."""
    import test_mcts_gpu as tg

    tg.test_collector_matches_oracle(host_pb, pymcts)


def test_mcts_context_drop_in_surface_on_cpu(host_pb):
    """pacbot_rs.MCTSContext drop-in (mcts.py over gym.py, both unmodified) with the reference's
    NumPy evaluator, as train_alphazero.py:119-131 drives it."""
    import test_mcts_gpu as tg

    tg.test_mcts_context_drop_in_surface(host_pb)


def test_experience_collector_drop_in_surface_on_cpu(host_pb):
    """pacbot_rs.ExperienceCollector drop-in: NumPy evaluator in, list of ExperienceItem out."""
    import test_mcts_gpu as tg

    tg.test_experience_collector_drop_in_surface(host_pb)
