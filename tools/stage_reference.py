"""Stage the reference's Python DQN path, UNMODIFIED, into the git-ignored `baseline/_ref/src/`.

    python tools/stage_reference.py            (also called by __graft_entry__.build())

The GPU box has no /root/reference, but two things must run the reference's own files there:
`bench.py --impl reference` (its DQN leg times src/algorithms/train_dqn.py at `--device cuda`)
and tests/test_dropin_reference_gpu.py (the reference's replay_buffer.py over the product's
`PacmanGym`).  `baseline/_ref/` is in .gitignore (the reference's sources never enter this repo's
history) and not in .gpurunignore, so the copy travels with the snapshot like the built .so
files.  Nothing under pacbot-rl_b200/ reads it.  A no-op where /root/reference does not exist."""
import filecmp
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src"
DST = os.path.join(ROOT, "baseline", "_ref", "src")
FILES = ["models.py", "policies.py", "replay_buffer.py", "utils.py", "timing.py",
         "debug_probe_envs.py", os.path.join("algorithms", "train_dqn.py")]


def stage() -> bool:
    if not os.path.isdir(REF):
        return os.path.isdir(DST)
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
    return True


if __name__ == "__main__":
    ok = stage()
    print(f"{'staged' if ok else 'nothing to stage'}: {DST}")
    sys.exit(0)
