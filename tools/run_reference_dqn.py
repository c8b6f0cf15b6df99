"""Run the reference's UNMODIFIED src/algorithms/train_dqn.py for a few iterations and report its
iterations/s: the reference-side number for BASELINE.json configs[3] ("train_dqn.py default
hyperparams") that the B200 trainer's it/s is read against.  Used by `bench.py --impl reference`
(on the GPU box, `--device cuda`) and by hand.

    python tools/run_reference_dqn.py [--iters 30] [--device cpu|cuda] [--env oracle|product]
                                      [--ref-src DIR] [--out FILE]

The script is executed as it is (runpy, `--no-wandb --device D --num_iters N`), from
/root/reference/src in the build container or from the staged copy baseline/_ref/src elsewhere
(tools/stage_reference.py).  What has to be supplied around it, because the Rust extension cannot
be built here (no cargo, engine crate not vendored) -- nothing else is touched:
  * `pacbot_rs.PacmanGym`  -> `--env oracle`: the scalar C oracle (oracle/), one env per object,
                              per-env random draw tape;  `--env product`: this repo's drop-in
                              `pacbot_rl_b200.PacmanGym` (one 1-env CUDA engine per object): the
                              reference's loop running on the B200 engine, unchanged
  * `gymnasium`            -> empty stub (imported, never used, by debug_probe_envs.py)
  * `checkpoints/q_net-old.safetensors` (loaded at import, replay_buffer.py:38-39) -> random
    weights of the reference's own QNetV2 in a scratch directory; a random old policy never
    finishes `reset_env`'s phase 1, so `policy_old` plays on a board whose super pellets are
    removed (PacmanGym built with remove_super_pellets): phase 1 is empty, as in
    `python -m pacbot_rl_b200.train_dqn` without --old-policy
  * `tqdm.tqdm`            -> a pass-through that timestamps every iteration
  * `wandb`                -> a stub when the package is missing; `wandb.finish` exits the process
                              (the script would go on to an interactive viewer)
The number is a REPORTED baseline with a C restatement of the env underneath (`--env oracle`), not
the Rust crate."""
import argparse
import json
import os
import runpy
import sys
import tempfile
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def default_ref_src() -> str:
    for cand in ("/root/reference/src", os.path.join(ROOT, "baseline", "_ref", "src")):
        if os.path.isdir(cand):
            return cand
    raise SystemExit("reference sources not found (neither /root/reference/src nor baseline/_ref/src)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--env", default="oracle", choices=["oracle", "product"])
    ap.add_argument("--ref-src", default=None)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    REF_SRC = args.ref_src or default_ref_src()
    sys.path.insert(0, ROOT)
    from oracle import pyoracle

    pyoracle.build()
    L = pyoracle.lib()
    seed_rng = np.random.default_rng(0)
    counter = [0]

    class OraclePacmanGym:  # pacbot_rs.pyi:6-19 on the oracle; draws from a per-env random tape
        def __init__(self, random_start: bool, random_ticks: bool):
            cfg = np.array([int(bool(random_start)) | (int(bool(random_ticks)) << 1) | 8], np.uint8)
            tape = seed_rng.integers(0, 2**32, size=(1, 65521), dtype=np.uint32)
            self._b = pyoracle.OracleBatch(1, cfg, tape)
            self._e = L.po_batch_env(self._b._h, 0)
            counter[0] += 1

        def reset(self):
            self._b.reset()

        def step(self, action):
            r, d = self._b.step(np.array([action], np.uint8))
            return int(r[0]), bool(d[0])

        def first_ai_done(self):
            return bool(L.po_gym_first_ai_done(self._e))

        def is_done(self):
            return bool(L.po_gym_is_done(self._e))

        def score(self):
            return int(L.po_gym_score(self._e))

        def lives(self):
            return int(L.po_gym_lives(self._e))

        def remaining_pellets(self):
            return int(L.po_gym_remaining_pellets(self._e))

        def action_mask(self):
            return self._b.action_mask()[0].tolist()

        def obs_numpy(self):
            return self._b.obs()[0]

    if args.env == "product":
        import pacbot_rl_b200 as pb

        class PacmanGym(pb.PacmanGym):   # the drop-in itself; only the super-pellet config is set
            def __init__(self, random_start: bool, random_ticks: bool):
                super().__init__(random_start, random_ticks)
                st = self._batch.get_state()
                st["cfg_flags"] |= 8     # remove_super_pellets (see the module docstring)
                self._batch.set_state(st)
                counter[0] += 1
    else:
        PacmanGym = OraclePacmanGym

    from importlib.machinery import ModuleSpec

    shim = types.ModuleType("pacbot_rs")
    shim.PacmanGym = PacmanGym
    shim.__spec__ = ModuleSpec("pacbot_rs", None)
    sys.modules["pacbot_rs"] = shim
    gym_stub = types.ModuleType("gymnasium")
    gym_stub.__spec__ = ModuleSpec("gymnasium", None)
    sys.modules["gymnasium"] = gym_stub

    stamps = []

    def tqdm(iterable, **_kw):
        for x in iterable:
            stamps.append(time.perf_counter())
            yield x
        stamps.append(time.perf_counter())

    import tqdm as real_tqdm   # the real package stays importable (torch looks its spec up)

    real_tqdm.tqdm = tqdm       # `from tqdm import tqdm` in the script picks this up

    scratch = tempfile.mkdtemp(prefix="ref_dqn_")
    os.makedirs(os.path.join(scratch, "checkpoints"))
    os.chdir(scratch)
    sys.path.insert(0, REF_SRC)
    import models as ref_models
    import safetensors.torch
    import torch

    try:
        import wandb
    except ImportError:   # the script only needs init / config / log / finish / Artifact
        wandb = types.ModuleType("wandb")
        wandb.__spec__ = ModuleSpec("wandb", None)
        wandb.config = types.SimpleNamespace()
        wandb.run = None

        def _init(config=None, **_kw):
            for k, v in (config or {}).items():
                setattr(wandb.config, k, v)

        wandb.init = _init
        wandb.log = lambda *_a, **_k: None
        sys.modules["wandb"] = wandb

    safetensors.torch.save_file(ref_models.QNetV2((17, 28, 31), 5).state_dict(),
                                os.path.join("checkpoints", "q_net-old.safetensors"))

    def finish(*_a, **_k):
        raise SystemExit(0)

    wandb.finish = finish
    sys.argv = ["train_dqn.py", "--no-wandb", "--device", args.device, "--num_iters", str(args.iters),
                "--checkpoint-dir", os.path.join(scratch, "checkpoints")]
    real_stdout = sys.stdout
    sys.stdout = open(os.devnull, "w")   # the script prints a timing line per block per iteration
    t0 = time.perf_counter()
    try:
        runpy.run_path(os.path.join(REF_SRC, "algorithms", "train_dqn.py"), run_name="__main__")
    except SystemExit:
        pass
    finally:
        sys.stdout = real_stdout
    total = time.perf_counter() - t0
    d = np.diff(np.array(stamps))
    # iteration 0 builds the target network and runs an evaluation episode like every 10th one;
    # report the plain mean over all timed iterations (what the reference's tqdm bar would show)
    out = {
        "what": f"reference src/algorithms/train_dqn.py, unmodified, --device {args.device}, default "
                "hyper-parameters (batch 512, 32 envs, 4 experience steps/iter, eval episode every 10 "
                "iters), env = " + ("C restatement of pacbot_rs (oracle/)" if args.env == "oracle" else
                                    "pacbot_rl_b200.PacmanGym (the B200 engine, one env per object)")
                + " behind a pacbot_rs.PacmanGym shim",
        "kind": "port" if args.env == "oracle" else "reference loop on the product engine",
        "device": args.device,
        "env": args.env,
        "iters": int(len(d)),
        "iters_per_sec": float(len(d) / d.sum()),
        "median_iter_s": float(np.median(d)),
        "fill_and_setup_s": float(stamps[0] - t0),
        "total_s": float(total),
        "torch_threads": torch.get_num_threads(),
        "host_cpus": os.cpu_count(),
        "envs_created": counter[0],
    }
    line = json.dumps(out)
    print(line)
    if args.out:
        with open(os.path.join(ROOT, args.out) if not os.path.isabs(args.out) else args.out, "w") as f:
            f.write(line + "\n")


if __name__ == "__main__":
    main()
