"""CPU runs of the product's experience-generation host code (pacbot-rl_b200/replay_buffer.py:
`ReplayBuffer.generate_experience_step` with the interleaved old-policy phase, and the synchronous
`reset_env`) UNMODIFIED over the host build of the device engine: the GPU suite's own test bodies
(tests/test_dqn_gpu.py) re-run with `HostEnv` in place of `BatchedPacmanGym` (host_stand_ins.py).

Together with tests/test_reference_replay_live.py this closes the chain on the CPU:
reference replay_buffer.py (live) == restated per-env loop == product ReplayBuffer, and the GPU
suite repeats the last link on the compiled sm_100a code.  Not covered here: `sample_batch`
(pb_gather_obs is a CUDA kernel) and the observation encoder (observations come from the oracle)."""
import pytest

from host_stand_ins import hm, host_pb  # noqa: F401  (pytest fixtures)


def test_product_replay_buffer_two_stage_reset_on_cpu(host_pb, oracle):
    import test_dqn_gpu as tg

    tg.test_two_stage_reset_matches_reference_loop(host_pb, oracle)


def test_product_replay_ring_matches_oracle_rollout_on_cpu(host_pb, oracle):
    import test_dqn_gpu as tg

    tg.test_replay_ring_matches_oracle_rollout(host_pb, oracle)


def test_product_synchronous_reset_env_on_cpu(host_pb, oracle):
    import test_dqn_gpu as tg

    tg.test_synchronous_reset_env_matches_reference_loop(host_pb, oracle)
