#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched PacBot hot path (step + full f32 observation encode).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one batch: draw a random valid action per env
(device RNG), PacmanGym.step for every env (auto reset on done), and the 17x28x31 float32
observation of every env written into a GPU-resident ring.  Workload at N=1 is BASELINE.json
configs[1]: 65 536 envs on one B200; with N GPUs every rank runs the same shard size on its own
env-id range (weak scaling, no data-path collective).  Prints ONE JSON line (rank 0).

`--impl reference` times the CPU implementation of the same path (the C restatement under
oracle/, since the Rust crate cannot be built here) on all host threads, same workload shape.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

OBS_BYTES = 17 * 28 * 31 * 4
STATE_BYTES = 176
# algorithmic bytes per env-step (DESIGN.md "Roofline"): state read + state write + action in
# + reward/done/mask out + observation write
STEP_IO_BYTES = 2 * STATE_BYTES + 1 + 4 + 1 + 5
BYTES_PER_ENV_STEP = STEP_IO_BYTES + OBS_BYTES
# the dominant kernel (k_encode_obs) alone: reads the packed state, writes the observation
ENCODE_BYTES_PER_ENV = STATE_BYTES + OBS_BYTES


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=65536)
    ap.add_argument("--ring-gb", type=float, default=12.0, help="observation ring size per GPU")
    ap.add_argument("--chunk-envs", type=int, default=65536, help="envs per encode launch")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-sample-envs", type=int, default=8192)
    ap.add_argument("--cpu-sample-steps", type=int, default=48)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-dqn", action="store_true")
    ap.add_argument("--dqn-iters", type=int, default=200)
    ap.add_argument("--no-mcts", action="store_true")
    ap.add_argument("--pipeline", default="auto", choices=["auto", "flip", "chunks", "none"],
                    help="auto: flip while both state buffers fit L2 comfortably (<= 64 MB), else none; "
                         "flip: double-buffered state, k_step of step t+1 overlaps k_encode_obs of step t "
                         "(two streams); chunks: in place, chunk-skewed two-stream pipeline; none: k_step "
                         "then k_encode_obs on one stream")
    ap.add_argument("--mcts-steps", type=int, default=400, help="search iterations timed in the mcts leg")
    return ap.parse_args()


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def encode_traffic(envs_per_launch: int):
    """DRAM bytes per k_encode_obs launch from the committed ncu capture (scaled per env)."""
    try:
        with open(os.path.join(ROOT, "profiles", "encode_traffic.json")) as f:
            d = json.load(f)
        per_env = (d["dram_bytes_read"] + d["dram_bytes_write"]) / d["envs_per_launch"]
        return per_env * envs_per_launch
    except Exception:
        return None


def dqn_throughput(dev, seed: int, iters: int, graph: bool = True):
    """BASELINE.json configs[3]: train_dqn defaults with GPU-resident envs, iterations/s."""
    import torch

    from pacbot_rl_b200 import train_dqn

    cfg = train_dqn.build_parser().parse_args(
        ["--device", str(dev), "--seed", str(seed), "--no-checkpoints", "--log_every", "1000000"])
    trainer = train_dqn.DQNTrainer(cfg, dev)
    trainer.replay.fill()
    trainer.epsilon = cfg.initial_epsilon
    if graph:
        trainer.capture_graph()
    for _ in range(10):
        trainer.train_iteration()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(iters):
        trainer.train_iteration()
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    m = trainer.pop_metrics()
    return {"iters_per_sec": iters / dt, "iters": iters, "cuda_graph": graph,
            "env_steps_per_sec": iters * cfg.experience_steps * cfg.num_parallel_envs / dt,
            "config": "train_dqn defaults: batch 512, 32 envs, 4 experience steps/iter, buffer "
                      "10k, QNetV2 fp32 (cuDNN, NHWC weights), double DQN, Adam",
            "loss": m["loss"]}


def mcts_throughput(dev, seed: int, num_envs: int, tree_size: int, steps: int, graph: bool,
                    cpu_steps: int = 0):
    """the AlphaZero experience collector (alphazero.rs generate_experience_step:
    one MCTS search iteration per tree -- select, leaf observation encode, NetV2 evaluation,
    backprop -- plus the env steps / re-roots / resets that become due), simulations per second."""
    import torch

    from pacbot_rl_b200 import alphazero, models

    model = models.NetV2((17, 28, 31), 6).to(dev).use_channels_last().eval()

    @torch.no_grad()
    def evaluator(obs, mask):  # train_alphazero.py:83-95 (reward_scale 1/100)
        pred = model(obs)
        return pred[:, -1] * 100.0, torch.softmax(pred[:, :-1].masked_fill(~mask, -torch.inf), dim=-1)

    cfg = alphazero.AlphaZeroConfig(tree_size, 1000, 0.99, num_envs)
    col = alphazero.BatchedExperienceCollector(evaluator, cfg, device=dev, seed=seed)
    if graph:
        col.enable_cuda_graph()
    finished: list = []
    for _ in range(max(10, tree_size)):  # warm-up incl. the first round of env steps
        col.generate_experience_step(finished)
    torch.cuda.synchronize(dev)
    items0 = int(col.ep_len.sum()) + sum(int(f[0].numel()) for f in finished)
    t0 = time.perf_counter()
    for _ in range(steps):
        col.generate_experience_step(finished)
    col.mcts.check()
    dt = time.perf_counter() - t0
    items = int(col.ep_len.sum()) + sum(int(f[0].numel()) for f in finished) - items0
    out = {"sims_per_sec": steps * num_envs / dt, "env_steps_per_sec": items / dt,
           "search_iterations": steps, "ms_per_iteration": 1e3 * dt / steps, "cuda_graph": graph,
           "config": f"{num_envs} trees x tree_size {tree_size}, NetV2 fp32 evaluator (cuDNN, NHWC) on "
                     "device-resident leaf observations, Dirichlet root noise, max_episode_length 1000"}
    if cpu_steps:
        import ctypes as C

        from oracle import pymcts

        L = pymcts.lib()
        L.pm_collector_bench.restype = C.c_double
        L.pm_collector_bench.argtypes = [pymcts._Config, C.c_uint64, C.c_int64, C.POINTER(C.c_int64)]
        n_items = C.c_int64(0)
        secs = L.pm_collector_bench(pymcts._Config(tree_size, 1000, 0.99, num_envs), seed, cpu_steps,
                                    C.byref(n_items))
        out["cpu_port_sims_per_sec"] = cpu_steps * num_envs / secs
        out["cpu_port_note"] = ("C restatement of mcts.rs/alphazero.rs (oracle/), 1 thread, tree + env "
                                "+ leaf observation encode only: its evaluator is the constant "
                                "untrained-model branch, no network")
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.lines = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [ln for (ts, ln) in self.lines if t0 - 0.05 <= ts <= t1 + 0.15] or [ln for _, ln in self.lines]
        for ln in rows:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": statistics.median(sm) if sm else None,
            "sm_max_mhz": max(smax) if smax else None,
            "power_w_max": max(power) if power else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


def cpu_baseline(n_envs: int, n_steps: int, threads: int, seed: int):
    from oracle import pyoracle

    pyoracle.build()
    # calibrate, then size the sample for roughly 15 s of wall time on this host
    c_secs, c_steps, _, _ = pyoracle.rollout(n_envs, 4, 1, seed, threads)
    rate = c_steps / max(c_secs, 1e-6)
    n_steps = int(min(max(n_steps, 15.0 * rate / n_envs), 4000))
    warm = max(2, n_steps // 8)
    secs, steps, _cs, episodes = pyoracle.rollout(n_envs, n_steps, warm, seed, threads)
    return {
        "value": steps / secs,
        "unit": "env-steps/s",
        "cores": threads,
        "kind": "port",
        "sample": (f"C restatement of pacbot_rs (oracle/, NOT the Rust crate): {n_envs} envs x "
                   f"{n_steps} steps (+{warm} warm-up), step + full f32 obs encode per env-step, "
                   f"random valid actions, auto reset, {threads} pthreads, {secs:.2f} s, "
                   f"{episodes} episodes; host has {os.cpu_count()} logical CPUs"),
    }


def workload_name(envs_per_gpu: int) -> str:
    """config.workload, identical in both arms (the reference arm times a bounded sample of it)"""
    return (f"{envs_per_gpu} envs/GPU batched step+observation encode, random valid actions, "
            "auto reset, random_start for half the envs, random_ticks (BASELINE.json configs[1])")


def run_reference(args, rank: int, world: int):
    """Reference arm: the CPU implementation of the path on all host threads (rank 0 only)."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_envs = args.cpu_sample_envs
    # bounded sample of the 65 536-env workload: every "step" steps n_envs envs once
    per_step = []
    from oracle import pyoracle

    pyoracle.build()
    pyoracle.rollout(n_envs, max(1, args.warmup), 0, args.seed, threads)  # warm-up (untimed)
    t_total, steps_total = 0.0, 0
    secs, steps, _cs, _ep = pyoracle.rollout(n_envs, args.steps, 2, args.seed + 1, threads)
    t_total += secs
    steps_total += steps
    value = steps_total / t_total
    line = {
        "impl": "reference",
        "metric": "env_steps_per_sec_incl_obs",
        "value": value,
        "unit": "env-steps/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "int32+f32",
        "data": "synthetic",
        "config": {"workload": workload_name(args.envs_per_gpu),
                   "envs_per_gpu": args.envs_per_gpu,
                   "sample_envs_per_step": n_envs, "host_threads": threads},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": threads, "kind": "port",
                         "sample": f"{n_envs} of {args.envs_per_gpu} envs per step x {args.steps} steps, C "
                                   f"restatement of pacbot_rs under oracle/ (Rust crate not buildable "
                                   f"here), {threads} pthreads"},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    del per_step
    emit(line)


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line goes to the real stdout; everything else any library prints (NCCL's
    version banner, torchrun notices) was redirected to stderr in main()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    args = parse_args()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import pacbot_rl_b200 as pb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG")  # keep stdout to the single JSON line
        dist.init_process_group("nccl", device_id=dev)

    n = args.envs_per_gpu
    chunk = min(args.chunk_envs, n)
    slot_bytes = chunk * OBS_BYTES
    n_slots = max(2, int(args.ring_gb * 1e9 // slot_bytes))
    ring = torch.empty((n_slots, chunk) + pb.OBS_SHAPE, dtype=torch.float32, device=dev)
    env = pb.BatchedPacmanGym(n, np.arange(n) < n // 2, True, device=dev, seed=args.seed,
                              env_id_base=rank * n, auto_reset=True)
    env.reset()
    actions = torch.empty(n, dtype=torch.uint8, device=dev)
    reward = torch.empty(n, dtype=torch.int32, device=dev)
    done = torch.empty(n, dtype=torch.bool, device=dev)
    mask = torch.empty((n, 5), dtype=torch.bool, device=dev)
    chunks = [(f, min(chunk, n - f)) for f in range(0, n, chunk)]
    launches_per_step = 1 + len(chunks)
    slot_ctr = [0]
    enc_events = []

    # Default (--pipeline flip): two streams and double-buffered env state (pacbot-rl_b200/rollout.py):
    # the step kernel of step t+1 runs while the encoder is still streaming step t, so the 10 us of
    # k_step are no longer exposed every step.  --pipeline chunks: the in-place variant (a single
    # encode chunk is split in two halves that share a ring slot).  --pipeline none: k_step then
    # k_encode_obs back to back on one stream.
    pipe = None
    mode = args.pipeline
    if mode == "auto":  # double buffering pays while both state buffers stay L2 resident
        # (measured at 2 097 152 envs per GPU, 352 MB of state: flip 123.5, chunks 124.0-124.4,
        # none 124.1-124.2 M env-steps/s -- nothing to win once the state streams from HBM)
        mode = "flip" if n * 176 * 2 <= 64 * 2**20 else "none"
    # flip with a single encode launch per step: ONE C call per step (pb_rollout_random_step does
    # the stream / event choreography inside the library; a busy host cannot put bubbles between
    # the 0.52 ms encoder launches)
    flipc = mode == "flip" and len(chunks) == 1 and n >= 2
    flip_window = [None, 0, 0]  # [open event, envs, launches] of the current timing window
    if flipc:
        flip_enc_stream = env.rollout_streams()[1]
        launches_per_step = 2
    if mode != "none" and n >= 2 and not flipc:
        from pacbot_rl_b200.rollout import DoubleBufferedRandomRollout, PipelinedRandomRollout

        if mode == "flip":   # double-buffered state: step t+1 overlaps encode t
            pchunks = chunks
            pipe = DoubleBufferedRandomRollout(env, chunks=pchunks)
            launches_per_step = 1 + len(pchunks)
        else:                         # "chunks": in place, step of one chunk overlaps encode of another
            half = (n + 1) // 2
            pchunks = chunks if len(chunks) >= 2 else [(0, half), (half, n - half)]
            pipe = PipelinedRandomRollout(env, chunks=pchunks)
            launches_per_step = 2 * len(pchunks)
        done = pipe.done
        timing = [None]
        window = [False]

        def obs_out(c: int, first: int, cnt: int):
            if len(chunks) >= 2:
                slot = ring[slot_ctr[0] % n_slots]
                slot_ctr[0] += 1
                return slot[:cnt]
            if c == 0:
                slot_ctr[0] += 1
            return ring[(slot_ctr[0] - 1) % n_slots][first:first + cnt]

        def on_encode(c: int, stream, before: bool):
            # encoder launches run back to back on the encode stream (each one's prologue overlaps
            # the previous one's tail); timing events go around WINDOWS of 8 consecutive steps so
            # that no event sits between two launches inside a window
            if timing[0] in ("open", "open+close") and before and c == 0:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(stream)
                enc_events.append([ev, None, 0, 0])
                window[0] = True
            if not before and window[0]:
                enc_events[-1][2] += pchunks[c][1]
                enc_events[-1][3] += 1
                if timing[0] in ("close", "open+close") and c == len(pchunks) - 1:
                    ev = torch.cuda.Event(enable_timing=True)
                    ev.record(stream)
                    enc_events[-1][1] = ev
                    window[0] = False

    def hot_step(call_idx: int, time_encode: bool, it: int = -1):
        if flipc:
            # timing events go around windows of 8 consecutive steps on the engine's encode stream
            # (no event between two encoder launches inside a window), one window every 64 steps:
            # a timing event is a host-visible completion and costs the running encoder ~14 us
            # (DESIGN.md section 9), so they are kept rare
            if it >= 0 and it % 64 == 0:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(flip_enc_stream)
                flip_window[:] = [ev, 0, 0]
            slot = ring[slot_ctr[0] % n_slots]
            slot_ctr[0] += 1
            env.rollout_random_step(slot[:n], call_idx=call_idx, actions_out=actions, mask_out=mask,
                                    reward_out=reward, done_out=done)
            if flip_window[0] is not None:
                flip_window[1] += n
                flip_window[2] += 1
                if it % 64 == 7 or it == args.steps - 1:
                    ev = torch.cuda.Event(enable_timing=True)
                    ev.record(flip_enc_stream)
                    enc_events.append([flip_window[0], ev, flip_window[1], flip_window[2]])
                    flip_window[0] = None
            return
        if pipe is not None:
            last = it % 64 == 7 or it == args.steps - 1
            timing[0] = "open" if it % 64 == 0 else ("close" if last else None)
            if it % 64 == 0 and last:  # a one-step window (tiny --steps)
                timing[0] = "open+close"
            pipe.step(call_idx, obs_out, on_encode)
            return
        # random valid action per env + PacmanGym.step, one launch (pb_step_random)
        env.step_random(call_idx=call_idx, actions_out=actions, mask_out=mask, reward_out=reward,
                        done_out=done)
        if time_encode:
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
        for first, cnt in chunks:
            slot = ring[slot_ctr[0] % n_slots]
            slot_ctr[0] += 1
            env.obs_range(first, cnt, slot[:cnt])
        if time_encode:
            e1.record()
            enc_events.append([e0, e1, n, len(chunks)])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident throughput (`value`) ----
    call = 0
    for _ in range(max(3, args.warmup)):
        hot_step(call, False)
        call += 1
    if pipe is not None:
        pipe.join()
    if flipc:
        env.rollout_join()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    t_wall0 = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for it in range(args.steps):
        # the encoder launches are bracketed by CUDA events on every 8th step only: an event
        # record between k_step and k_encode_obs breaks their programmatic-dependent-launch
        # adjacency, which the other 7 steps keep (as real callers of step_obs do)
        hot_step(call, it % 8 == 0, it)
        call += 1
    if pipe is not None:
        pipe.join()  # the timing stream waits for both pipeline streams
    if flipc:
        env.rollout_join()
    ev1.record()
    barrier()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    ms_total = ev0.elapsed_time(ev1)
    enc_events = [e for e in enc_events if e[1] is not None]  # drop a window left open at the end
    enc_total_ms = sum(e[0].elapsed_time(e[1]) for e in enc_events)
    enc_total_envs = sum(e[2] for e in enc_events)
    enc_launches = sum(e[3] for e in enc_events)
    if enc_launches == 0:  # cannot happen with the windows above; keep the line printable anyway
        enc_total_ms, enc_total_envs, enc_launches = ms_total, n * args.steps, len(chunks) * args.steps
    enc_ms = enc_total_ms / enc_launches
    enc_envs_per_launch = enc_total_envs / enc_launches
    episodes = int(done.sum().item())
    ms_by_rank = [ms_total / args.steps]
    if world > 1:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        every = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(every, t)
        ms_by_rank = [float(x.item()) / args.steps for x in every]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    value = n * world * args.steps / (ms_total * 1e-3)

    # ---- end to end through the public API with HOST buffers (`e2e`) ----
    e2e = None
    e2e_full = None
    if not args.no_e2e:
        k_e2e = max(3, min(args.steps, 500))
        # The host side owns the actions here (as a host-side policy would): a pool of pinned
        # action buffers, uniform over 0..4 (a blocked move is a legal no-op, env.rs:404), and
        # pinned result buffers; every step copies its actions H2D and its results D2H.
        rng = np.random.default_rng(args.seed + rank)
        pool = 8
        host_actions = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(pool)]
        for buf in host_actions:
            buf.numpy()[:] = rng.integers(0, 5, n, dtype=np.uint8)
        h_reward = torch.empty(n, dtype=torch.int32).pin_memory().numpy()
        h_done = torch.empty(n, dtype=torch.bool).pin_memory().numpy()
        h_mask = torch.empty((n, 5), dtype=torch.bool).pin_memory().numpy()
        outs = (h_reward, h_done, h_mask)
        checksum = 0

        # flip mode: the same host-in / host-out step as ONE pipelined C call
        # (pb_rollout_host_step): the step kernel of call t+1 does not queue behind the encode of
        # call t
        host_roll = flipc

        def e2e_step(i: int):
            if host_roll:
                slot = ring[slot_ctr[0] % n_slots]
                slot_ctr[0] += 1
                env.rollout_host_step(host_actions[i % pool].numpy(), slot[:n], outs)
                return
            a = host_actions[i % pool].numpy()
            if n <= chunk:  # one C-ABI call: H2D actions, step, encode, D2H reward/done/mask
                slot = ring[slot_ctr[0] % n_slots]
                slot_ctr[0] += 1
                env.step_host(a, obs_out=slot[:n], out=outs)
            else:  # env population larger than a ring slot: encode chunk by chunk
                env.step_host(a, obs_out=None, out=outs)
                for first, cnt in chunks:
                    slot = ring[slot_ctr[0] % n_slots]
                    slot_ctr[0] += 1
                    env.obs_range(first, cnt, slot[:cnt])

        for i in range(3):
            e2e_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(k_e2e):
            e2e_step(i)
            checksum += int(h_reward[0]) + int(h_done[-1])   # the host reads the step's result
        if host_roll:
            env.rollout_join()
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": n * world * k_e2e / dt, "unit": "env-steps/s",
               "h2d_bytes_per_step": n, "d2h_bytes_per_step": n * (4 + 1 + 5), "steps": k_e2e,
               "what": ("BatchedPacmanGym.rollout_host_step -> pb_rollout_host_step (pb_step_flip + "
                        "pb_encode_obs on two engine-owned streams)"
                        if host_roll else "BatchedPacmanGym.step_host -> pb_step_host")
                       + ": pinned host actions in, reward/done/mask back to pinned host memory and "
                         "read by the host every step (the D2H copies ride a side stream while the "
                         "encoder runs), observation encoded into the device-resident ring (it is "
                         "consumed on the GPU)"}
        # variant that also drags the full observation back to the host (PCIe-bound by design)
        if rank == 0 and world == 1:
            n_small = min(n, 4096)
            t0 = time.perf_counter()
            reps = 3
            for _ in range(reps):
                env.obs_numpy(0, n_small)
            dt2 = time.perf_counter() - t0
            e2e_full = {"value": n_small * reps / dt2, "unit": "obs/s",
                        "d2h_bytes_per_obs": OBS_BYTES,
                        "what": f"obs_numpy of {n_small} envs to pageable host memory"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    enc_gbs = enc_envs_per_launch * ENCODE_BYTES_PER_ENV / (enc_ms * 1e-3) / 1e9
    line = {
        "metric": "env_steps_per_sec_incl_obs",
        "value": value,
        "unit": "env-steps/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": max(3, args.warmup),
        "ms_per_step": ms_total / args.steps,
        "ms_per_step_by_rank": [round(x, 5) for x in ms_by_rank],
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "int32+f32",
        "data": "synthetic",
        "config": {
            "workload": workload_name(n),
            "envs_per_gpu": n,
            "obs_ring": f"{n_slots} slots x {slot_bytes / 1e9:.2f} GB per GPU (outputs >> 126 MB L2; "
                        "the 11.5 MB packed state is re-read every step by construction)",
            "parallelism": f"env-shard x{world}, no collective",
            "bytes_per_env_step": BYTES_PER_ENV_STEP,
        },
        "hbm_roofline_frac_whole_step": (value / world) * BYTES_PER_ENV_STEP / 1e9 / peak,
        "roofline": {
            "kernel": "k_encode_obs",
            "bound": "hbm",
            "achieved": enc_gbs,
            "peak": peak,
            "unit": "GB/s",
            "frac": enc_gbs / peak,
            "traffic": encode_traffic(int(enc_envs_per_launch)),
            "traffic_source": "profiles/encode_traffic.json (dram__bytes_read.sum + "
                              "dram__bytes_write.sum of one ncu --set full capture, per launch)",
            "peak_source": peak_src,
            "bytes_per_launch": int(enc_envs_per_launch * ENCODE_BYTES_PER_ENV),
            "avg_launch_ms": enc_ms,
            "envs_per_launch": int(enc_envs_per_launch),
        },
        "pipeline": {
            "flip": "two streams, double-buffered state: k_step of step t+1 overlaps k_encode_obs of "
                    "step t (one pb_rollout_random_step call per step)",
            "chunks": "two streams, in place: k_step of one chunk overlaps k_encode_obs of another",
            "none": "none: k_step then k_encode_obs on one stream",
        }[mode if (pipe is not None or flipc) else "none"],
        "gpu_launches": launches_per_step * args.steps * world,
        "clocks": clocks,
        "episodes_finished_last_step": episodes,
    }
    if e2e is not None:
        line["e2e"] = e2e
    if e2e_full is not None:
        line["e2e_obs_to_host"] = e2e_full
    if world == 1 and not args.no_dqn:
        line["dqn"] = dqn_throughput(dev, args.seed, args.dqn_iters, graph=True)
        line["dqn_eager"] = dqn_throughput(dev, args.seed, args.dqn_iters, graph=False)
    if world == 1 and not args.no_mcts:
        cpu_steps = 0 if args.no_cpu_baseline else 400
        line["mcts"] = mcts_throughput(dev, args.seed, 128, 100, args.mcts_steps, True, cpu_steps)
        line["mcts_eager"] = mcts_throughput(dev, args.seed, 128, 100, max(50, args.mcts_steps // 4), False)
        line["mcts_4096"] = mcts_throughput(dev, args.seed, 4096, 100, max(50, args.mcts_steps // 4), True)
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args.cpu_sample_envs, args.cpu_sample_steps,
                                            os.cpu_count() or 1, args.seed)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
