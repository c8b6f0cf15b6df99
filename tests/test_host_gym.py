"""CPU runs of the per-env drop-in (pacbot-rl_b200/gym.py: `PacmanGym`, `PacmanGymView` -- the
pyo3 class surface of pacbot_rs.pyi:6-19 with its types, error behaviour and clone semantics)
UNMODIFIED over the host build of the device engine: the GPU suite's own test bodies
(tests/test_gym_api_gpu.py) re-run with `HostEnv` in place of `BatchedPacmanGym`."""
from host_stand_ins import hm, host_pb  # noqa: F401  (pytest fixtures)


def test_pacman_gym_drop_in_surface_on_cpu(host_pb, oracle):
    import test_gym_api_gpu as tg

    tg.test_pacman_gym_drop_in_surface(host_pb, oracle)


def test_pacman_gym_episode_matches_oracle_on_cpu(host_pb, oracle):
    import test_gym_api_gpu as tg

    tg.test_pacman_gym_episode_matches_oracle(host_pb, oracle)


def test_clone_and_view_on_cpu(host_pb):
    import test_gym_api_gpu as tg

    tg.test_clone_and_view(host_pb)
