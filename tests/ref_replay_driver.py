"""Subprocess driver for tests/test_reference_replay_live.py and
tests/test_dropin_reference_gpu.py: runs the reference's OWN src/replay_buffer.py (unmodified,
imported from /root/reference/src, or from the staged copy baseline/_ref/src on the GPU box) over a
`pacbot_rs` shim and dumps what ended up in the reference's replay buffer.  The shim's `PacmanGym`
is either the scalar C oracle (`oracle`) or this repo's drop-in `pacbot_rl_b200.PacmanGym` on the
GPU (`product`: one 1-env CUDA engine per object), env i drawing from row i of the same tape.

    python tests/ref_replay_driver.py <tape.npy> <out.npz> <n_envs> <steps> [oracle|product]

What has to be faked for the import to succeed, and nothing else:
  * `pacbot_rs`            the Rust extension (not buildable here) -> oracle-backed PacmanGym
  * `gymnasium`            imported (never used) by debug_probe_envs.py; not installed
  * checkpoints/q_net-old.safetensors  loaded at import time (replay_buffer.py:38-39);
    written to a scratch cwd with random weights, then `policy_old` is replaced by the test's
    deterministic old policy."""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF_SRC = next((d for d in ("/root/reference/src", os.path.join(ROOT, "baseline", "_ref", "src"))
                if os.path.isdir(d)), None)


def main():
    tape_path, out_path, n_envs, steps = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    backend = sys.argv[5] if len(sys.argv) > 5 else "oracle"
    if REF_SRC is None:
        raise SystemExit("reference sources not found (neither /root/reference/src nor baseline/_ref/src)")
    tape = np.load(tape_path)
    from oracle import pyoracle
    from replay_reference import MAIN_POLICY_CHANNELS, OLD_POLICY_CHANNELS, bfs_policy

    pyoracle.build()
    created = []

    class OraclePacmanGym:  # pacbot_rs.pyi:6-19 on top of the oracle
        def __init__(self, random_start: bool, random_ticks: bool):
            i = len(created)
            cfg = np.array([int(bool(random_start)) | (int(bool(random_ticks)) << 1)], np.uint8)
            row = tape[i % len(tape)][None, :]
            self._b = pyoracle.OracleBatch(1, cfg, row)
            created.append(self)

        def reset(self):
            self._b.reset()

        def step(self, action):
            r, d = self._b.step(np.array([action], np.uint8))
            return int(r[0]), bool(d[0])

        def first_ai_done(self):
            return bool(self._b.first_ai_done()[0])

        def is_done(self):
            return bool(self._b.is_done()[0])

        def action_mask(self):
            return self._b.action_mask()[0].tolist()

        def obs_numpy(self):
            return self._b.obs()[0].copy()

    if backend == "product":
        import pacbot_rl_b200 as pb

        class PacmanGym(pb.PacmanGym):   # the drop-in itself; only the draw tape is injected
            def __init__(self, random_start: bool, random_ticks: bool):
                i = len(created)
                super().__init__(random_start, random_ticks, device="cuda:0")
                row = np.ascontiguousarray(tape[i % len(tape)][None, :])
                self._batch.set_draw_tape(torch.from_numpy(row.view(np.int32)).to("cuda:0"))
                created.append(self)

        def draws_used(e):
            return int(e._batch.get_state()["rng_ctr"][0])
    else:
        PacmanGym = OraclePacmanGym

        def draws_used(e):
            return int(e._b.export_state()["rng_ctr"][0])

    shim = types.ModuleType("pacbot_rs")
    shim.PacmanGym = PacmanGym
    sys.modules["pacbot_rs"] = shim
    sys.modules["gymnasium"] = types.ModuleType("gymnasium")

    scratch = os.path.dirname(os.path.abspath(out_path))
    os.makedirs(os.path.join(scratch, "checkpoints"), exist_ok=True)
    os.chdir(scratch)
    sys.path.insert(0, REF_SRC)
    import models as ref_models
    import safetensors.torch

    safetensors.torch.save_file(ref_models.QNetV2((17, 28, 31), 5).state_dict(),
                                os.path.join("checkpoints", "q_net-old.safetensors"))
    import replay_buffer as ref_rb  # the reference module, unmodified

    created.clear()  # the import built one throw-away env for the obs shape (replay_buffer.py:34)

    def as_torch(policy):
        def f(obs, masks):
            return torch.from_numpy(policy(obs.numpy(), masks.numpy()).astype(np.int64))
        return f

    ref_rb.policy_old = as_torch(bfs_policy(OLD_POLICY_CHANNELS))
    rb = ref_rb.ReplayBuffer(
        maxlen=n_envs * steps, policy=as_torch(bfs_policy(MAIN_POLICY_CHANNELS)),
        num_parallel_envs=n_envs, random_start_proportion=0.5, device=torch.device("cpu"))
    assert len(created) == n_envs
    for _ in range(steps):
        rb.generate_experience_step()
    items = list(rb._buffer)
    assert len(items) == n_envs * steps
    # items were appended env by env within a step (replay_buffer.py:107-129)
    obs = np.stack([it.obs.numpy() for it in items]).view(np.uint32).reshape(steps, n_envs, 17, 28, 31)
    np.savez_compressed(
        out_path, obs=obs,
        action=np.array([it.action for it in items], np.int64).reshape(steps, n_envs),
        reward=np.array([it.reward for it in items], np.int64).reshape(steps, n_envs),
        done=np.array([it.next_obs is None for it in items]).reshape(steps, n_envs),
        next_mask=np.array([it.next_action_mask for it in items]).reshape(steps, n_envs, 5),
        draws_used=np.array([draws_used(e) for e in created]),
        final_score=np.array([e.score() for e in created]) if backend == "product" else np.zeros(0))


if __name__ == "__main__":
    main()
