"""Device timeline of bench.py's timed loop: where every microsecond of the N-step protocol goes.

    python tools/timeline.py [--steps 20] [--warmup 5] [--envs 65536] [--repeat 3] [--out FILE]

Runs exactly the calls of bench.py's `value` loop (pb_rollout_random_step per step: k_step on the
step stream, k_encode_obs on the encode stream, double-buffered state; the same gate, the same
events) with the engine's device timeline on (pb_trace_*: every launch records its first-CTA
begin, dependency-resolved and last-warp end time in %globaltimer ns; one-thread mark kernels
bracket the timed region in the same clock), and prints the account:

    region   = mark1 - mark0                         (what the CUDA events of bench.py time)
    encoders = sum of (end - ready) of k_encode_obs  (the HBM-bound work)
    gaps     = encoder-to-encoder idle time on the encode stream (launch + prologue)
    head     = first encoder's ready - mark0         (un-hidden first k_step + launch latency)
    tail     = mark1 - last encoder's end            (join)
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--repeat", type=int, default=3)
    ap.add_argument("--gate-cycles", type=int, default=2_000_000)
    ap.add_argument("--out", default=None)
    ap.add_argument("--summary-only", action="store_true")
    args = ap.parse_args()
    import numpy as np
    import torch

    import pacbot_rl_b200 as pb

    dev = torch.device("cuda:0")
    n = args.envs
    env = pb.BatchedPacmanGym(n, np.arange(n) < n // 2, True, device=dev, seed=0, auto_reset=True)
    env.reset()
    ring = torch.empty((3, n) + pb.OBS_SHAPE, dtype=torch.float32, device=dev)
    actions = torch.empty(n, dtype=torch.uint8, device=dev)
    reward = torch.empty(n, dtype=torch.int32, device=dev)
    done = torch.empty(n, dtype=torch.bool, device=dev)
    mask = torch.empty((n, 5), dtype=torch.bool, device=dev)
    call = [0]

    def hot_step():
        env.rollout_random_step(ring[call[0] % 3], call_idx=call[0], actions_out=actions, mask_out=mask,
                                reward_out=reward, done_out=done)
        call[0] += 1

    runs = []
    for rep in range(args.repeat):
        for _ in range(max(3, args.warmup)):
            hot_step()
        env.rollout_join()
        torch.cuda.synchronize(dev)
        env.trace_begin(4 * args.steps + 16)
        if args.gate_cycles:
            torch.cuda._sleep(args.gate_cycles)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        env.trace_mark()
        for _ in range(args.steps):
            hot_step()
        env.rollout_join()
        env.trace_mark()
        ev1.record()
        torch.cuda.synchronize(dev)
        rec = env.trace_end()
        marks = rec[rec["kind"] == 0]
        steps = rec[rec["kind"] == 1]
        encs = rec[rec["kind"] == 2]
        t0 = int(marks[0]["end_ns"])
        us = lambda t: (int(t) - t0) / 1e3   # noqa: E731
        region = us(marks[1]["begin_ns"])
        enc_busy = sum(us(e["end_ns"]) - us(e["ready_ns"]) for e in encs)
        gaps = [us(encs[i + 1]["ready_ns"]) - us(encs[i]["end_ns"]) for i in range(len(encs) - 1)]
        rows = []
        for i, (s, e) in enumerate(zip(steps, encs)):
            rows.append({"step": i, "k_step_begin": us(s["begin_ns"]), "k_step_end": us(s["end_ns"]),
                         "enc_begin": us(e["begin_ns"]), "enc_ready": us(e["ready_ns"]),
                         "enc_end": us(e["end_ns"]), "enc_us": us(e["end_ns"]) - us(e["ready_ns"])})
        runs.append({
            "event_ms_per_step": ev0.elapsed_time(ev1) / args.steps,
            "region_us": region, "region_us_per_step": region / args.steps,
            "encoder_busy_us": enc_busy, "encoder_us_mean": enc_busy / len(encs),
            "encoder_us_first3": [r["enc_us"] for r in rows[:3]],
            "encoder_us_min_max": [min(r["enc_us"] for r in rows), max(r["enc_us"] for r in rows)],
            "gap_us_sum": sum(gaps), "gap_us_mean": sum(gaps) / max(1, len(gaps)),
            "gap_us_min_max": [min(gaps), max(gaps)] if gaps else None,
            "head_us": rows[0]["enc_ready"], "first_k_step_us": rows[0]["k_step_end"] - rows[0]["k_step_begin"],
            "k_step_us_mean": sum(r["k_step_end"] - r["k_step_begin"] for r in rows) / len(rows),
            "tail_us": region - rows[-1]["enc_end"],
            "account_us": {"head": rows[0]["enc_ready"], "encoders": enc_busy, "gaps": sum(gaps),
                           "tail": region - rows[-1]["enc_end"]},
            "rows": rows,
        })
    out = {"steps": args.steps, "warmup": args.warmup, "envs": n, "gate_cycles": args.gate_cycles, "runs": runs}
    text = json.dumps(out, indent=1)
    if args.out:
        with open(args.out, "w") as f:
            f.write(text)
    for r in runs:
        print(json.dumps({k: v for k, v in r.items() if k != "rows"}))
    if args.summary_only:
        return
    print("first run, per step:")
    for r in runs[0]["rows"]:
        print("  step {step:3d}  k_step {k_step_begin:9.1f}..{k_step_end:9.1f}  enc begin {enc_begin:9.1f} "
              "ready {enc_ready:9.1f} end {enc_end:9.1f}  ({enc_us:6.1f} us)".format(**r))


if __name__ == "__main__":
    main()
