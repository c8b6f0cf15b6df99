"""A/B of capturing the DQN iteration on a high-priority stream (`torch.cuda.Stream(priority=-1)`,
`torch.cuda.graph(..., stream=that)`) so that its kernels win against the evaluation episode's
batch-1 kernels.  Needs the two lines described in tools/experiments/README.md patched into
train_dqn.py (`--no-train-priority` switched it off); kept as the record of the measurement."""
import sys, time, json
sys.path.insert(0, "/root/repo")
import torch
from pacbot_rl_b200 import train_dqn
dev = torch.device("cuda:0")
def run(extra):
    cfg = train_dqn.build_parser().parse_args(["--device", "cuda:0", "--no-checkpoints", "--log_every", "1000000"])
    tr = train_dqn.DQNTrainer(cfg, dev)
    tr.replay.fill(cfg.batch_size); tr.epsilon = 0.1
    tr.capture_graph()
    for _ in range(10): tr.train_iteration()
    tr.maybe_evaluate(force=True); tr.finish_evaluation()
    torch.cuda.synchronize()
    out = {}
    for with_eval in (False, True, False, True):
        t0 = time.perf_counter()
        for _ in range(300):
            tr.train_iteration()
            if with_eval: tr.maybe_evaluate()
        if with_eval: tr.finish_evaluation()
        torch.cuda.synchronize()
        out.setdefault("eval" if with_eval else "noeval", []).append(round(300 / (time.perf_counter() - t0), 1))
    out["eval_steps"] = [e["eval_episode_steps"] for e in tr.eval_log[-5:]]
    return out
print("priority", run([]))
print("no priority", run(["--no-train-priority"]))
