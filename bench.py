#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched PacBot hot path (step + full f32 observation encode).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one batch: draw a random valid action per env
(device RNG), PacmanGym.step for every env (auto reset on done), and the 17x28x31 float32
observation of every env written into a GPU-resident ring.  Workload at N=1 is BASELINE.json
configs[1]: 65 536 envs on one B200; with N GPUs every rank runs the same shard size on its own
env-id range (weak scaling, no data-path collective).  Prints ONE JSON line (rank 0).

Legs of one run (all in the one JSON line):
  value / roofline / e2e / verified   configs[1], 65 536 envs per GPU (the headline)
  config2                             configs[2], 2 097 152 envs per GPU, chunked encode
  dqn / dqn_dp                        configs[3] (1 GPU) / configs[4] (N GPUs, NCCL all-reduce)
  cpu_baseline                        the C restatement under oracle/ on the host cores (N=1)

`verified`: after the timed loop an UNTIMED pass takes the states of a 4 096-env sample, runs
one more hot step through the same calls, and has the CPU oracle replay it from those states
(same counter-RNG draws): actions, reward, done, mask, full state and the observation bits in
the ring slot the step wrote must all be identical.

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
CONFIG2_ENVS = 2 * 1024 * 1024


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
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--no-config2", action="store_true")
    ap.add_argument("--config2-steps", type=int, default=12)
    ap.add_argument("--no-dqn", action="store_true")
    ap.add_argument("--dqn-iters", type=int, default=200)
    ap.add_argument("--no-mcts", action="store_true")
    ap.add_argument("--no-affinity", action="store_true",
                    help="do not pin this rank's host threads to its own slice of the GPU-local CPUs")
    ap.add_argument("--pipeline", default="auto", choices=["auto", "flip", "chunks", "none"],
                    help="auto: flip while both state buffers fit L2 comfortably (<= 64 MB), else chunks; "
                         "chunks: in place, k_step of one 65 536-env chunk beside the encoders of the "
                         "other chunks (engine-owned streams, one C call per chunk); "
                         "flip: double-buffered state, k_step of step t+1 overlaps k_encode_obs of step t "
                         "(two engine-owned streams, one C call per step); none: k_step then "
                         "k_encode_obs on one stream")
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


# =================================================================================================
# host plumbing: CPU affinity, clock sampling
# =================================================================================================
def bind_rank_to_cpus(local_rank: int, local_world: int):
    """Pin this process to its own slice of the CPUs NVML reports as local to its GPU.  With one
    process per GPU every rank spins on a stream synchronisation once per e2e step; unpinned,
    eight such processes (plus their NCCL / CUDA helper threads) migrate over the same cores."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
    except Exception:
        cpus = []
    allowed = sorted(os.sched_getaffinity(0))
    cpus = [c for c in cpus if c in allowed] or allowed
    if local_world > 1:
        # ranks whose GPUs share a CPU set split it evenly (all 8 GPUs of this pool's boxes report
        # the same NUMA node)
        per = max(1, len(cpus) // local_world)
        mine = cpus[(local_rank * per) % len(cpus):][:per] or cpus
    else:
        mine = cpus
    try:
        os.sched_setaffinity(0, mine)
    except OSError:
        return None
    return mine


class ClockSampler:
    """SM clock / throttle reasons / power sampled DURING the timed region: an NVML polling thread
    (2 ms period; the timed region of a 20-step run is 10 ms, too short for `nvidia-smi -lms`),
    `nvidia-smi` as the fallback.  Started before the warm-up so nothing idles between warm-up and
    the first timed step."""

    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
               "sw_power_cap": 0x4}

    def __init__(self, index: int, period_s: float = 0.002):
        self.samples = []   # (wall time, sm MHz, sm max MHz, power W, reasons bitmask)
        self._stop = threading.Event()
        self.kind = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self._max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self.kind = "nvml"
            self._period = period_s
            self._thread = threading.Thread(target=self._poll, daemon=True)
            self._thread.start()
        except Exception:
            self._start_smi(index)

    def _poll(self):
        nv, h = self._nv, self._h
        while not self._stop.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    reasons = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    reasons = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                power = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                self.samples.append((time.time(), sm, self._max, power, reasons))
            except Exception:
                pass
            time.sleep(self._period)

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def _start_smi(self, index: int):
        try:
            self._proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "20", "-i", str(index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.kind = "nvidia-smi"
            self._thread = threading.Thread(target=self._pump, daemon=True)
            self._thread.start()
        except Exception:
            self._proc = None

    def _pump(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self._proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                bits = sum(self.REASONS[n] for n, v in zip(names, parts[3:7]) if v.lower().startswith("active"))
                self.samples.append((time.time(), float(parts[0]), float(parts[1]), float(parts[2]), bits))
            except ValueError:
                continue

    def stop(self):
        self._stop.set()
        if self.kind == "nvidia-smi" and getattr(self, "_proc", None) is not None:
            time.sleep(0.05)
            self._proc.terminate()

    def summary(self, t0: float, t1: float):
        """Clocks over the wall-clock window [t0, t1] (the timed region)."""
        if self.kind is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi / NVML unavailable"]}
        rows = [s for s in self.samples if t0 <= s[0] <= t1]
        window = "timed region"
        if len(rows) < 2:   # a very short region: widen to the samples right around it
            rows = [s for s in self.samples if t0 - 0.05 <= s[0] <= t1 + 0.05] or list(self.samples)
            window = "timed region +-50 ms"
        bits = 0
        for s in rows:
            bits |= s[4]
        return {
            "sm_mhz": statistics.median(s[1] for s in rows) if rows else None,
            "sm_max_mhz": max(s[2] for s in rows) if rows else None,
            "power_w_max": max(s[3] for s in rows) if rows else None,
            "samples": len(rows),
            "window": window,
            "source": self.kind,
            "reasons": sorted(n for n, b in self.REASONS.items() if bits & b),
        }


# =================================================================================================
# CPU arms (the oracle under oracle/ is the checker / the reported baseline, never the product)
# =================================================================================================
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


def workload_config(args, world: int) -> dict:
    """`config` of the JSON line: identical in both arms (the reference arm runs the same
    workload on the host cores)."""
    n = args.envs_per_gpu
    chunk = min(args.chunk_envs, n)
    slot_bytes = chunk * OBS_BYTES
    n_slots = max(2, int(args.ring_gb * 1e9 // slot_bytes))
    return {
        "workload": (f"{n} envs/GPU batched step+observation encode, random valid actions, "
                     "auto reset, random_start for half the envs, random_ticks "
                     "(BASELINE.json configs[1])"),
        "envs_per_gpu": n,
        "obs_ring": f"{n_slots} slots x {slot_bytes / 1e9:.2f} GB per GPU (outputs >> 126 MB L2; no "
                    "flush needed: every step writes 3.87 GB of new lines)",
        "parallelism": f"env-shard x{world}, no collective",
        "bytes_per_env_step": BYTES_PER_ENV_STEP,
    }


def run_reference(args, rank: int, world: int):
    """Reference arm: the CPU implementation of the path on all host threads (rank 0 only), on
    the product arm's config: every step steps + encodes all `envs_per_gpu` envs once."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_envs = args.envs_per_gpu
    from oracle import pyoracle

    pyoracle.build()
    warm = max(1, args.warmup)
    secs, steps_total, _cs, _ep = pyoracle.rollout(n_envs, args.steps, warm, args.seed + 1, threads)
    value = steps_total / secs
    line = {
        "impl": "reference",
        "metric": "env_steps_per_sec_incl_obs",
        "value": value,
        "unit": "env-steps/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": warm,
        "ms_per_step": 1e3 * secs / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "int32+f32",
        "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": threads, "kind": "port",
                         "sample": f"all {n_envs} envs of one GPU's shard per step x {args.steps} steps "
                                   f"(+{warm} warm-up), C restatement of pacbot_rs under oracle/ (the "
                                   f"Rust crate is not buildable here), {threads} pthreads, step + "
                                   "full f32 obs encode per env-step, random valid actions, auto reset"},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    dqn_ref = reference_dqn(args)
    if dqn_ref is not None:
        line["dqn"] = dqn_ref
    emit(line)


def reference_dqn(args, env: str = "oracle"):
    """configs[3] on the reference side: its UNMODIFIED src/algorithms/train_dqn.py (staged by
    build() into the git-ignored baseline/_ref/, never committed) over a `pacbot_rs` shim,
    `--device cuda` when there is a GPU.  env "oracle": the C restatement under the shim (the
    reference arm); env "product": this repo's drop-in `PacmanGym` under the shim, i.e. the
    reference's own loop on the B200 engine (reported in the product arm).  None when disabled."""
    if args.no_dqn:
        return None
    tool = os.path.join(ROOT, "tools", "run_reference_dqn.py")
    ref_src = os.path.join(ROOT, "baseline", "_ref", "src")
    if not (os.path.exists(tool) and os.path.isdir(ref_src)):
        return {"unavailable": "baseline/_ref/src not staged (build() stages it where /root/reference exists)"}
    try:
        import torch

        device = "cuda" if torch.cuda.is_available() else "cpu"
    except Exception:
        device = "cpu"
    iters = 30 if device == "cuda" else 12
    try:
        res = subprocess.run([sys.executable, tool, "--iters", str(iters), "--device", device,
                              "--env", env, "--ref-src", ref_src],
                             capture_output=True, text=True, timeout=600)
        lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
        return json.loads(lines[-1]) if lines else {"unavailable": (res.stderr or "no output")[-300:]}
    except Exception as exc:  # the env-step line above must still be printed
        return {"unavailable": repr(exc)[:300]}


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


# =================================================================================================
# the rollout leg: value, roofline, e2e, verified for one env population
# =================================================================================================
class Ctx:
    pass


def rollout_leg(cx, n: int, steps: int, warmup: int, ring_gb: float, chunk_envs: int, pipeline: str,
                seed: int, want_e2e: bool, e2e_steps: int, verify: bool, sampler=None):
    torch, np, pb, dist = cx.torch, cx.np, cx.pb, cx.dist
    dev, rank, world = cx.dev, cx.rank, cx.world
    chunk = min(chunk_envs, n)
    slot_bytes = chunk * OBS_BYTES
    n_slots = max(2, int(ring_gb * 1e9 // slot_bytes))
    ring = torch.empty((n_slots, chunk) + pb.OBS_SHAPE, dtype=torch.float32, device=dev)
    gid_base = rank * n
    env = pb.BatchedPacmanGym(n, np.arange(n) < n // 2, True, device=dev, seed=seed,
                              env_id_base=gid_base, auto_reset=True)
    env.reset()
    actions = torch.empty(n, dtype=torch.uint8, device=dev)
    reward = torch.empty(n, dtype=torch.int32, device=dev)
    done = torch.empty(n, dtype=torch.bool, device=dev)
    mask = torch.empty((n, 5), dtype=torch.bool, device=dev)
    chunks = [(f, min(chunk, n - f)) for f in range(0, n, chunk)]
    mode = pipeline
    if mode == "auto":  # double buffering pays while both state buffers stay L2 resident
        # (measured at 2 097 152 envs per GPU, 352 MB of state: flip 123.5, none 124.1-124.2 M
        # env-steps/s -- nothing to win once the state streams from HBM)
        mode = "flip" if n * STATE_BYTES * 2 <= 64 * 2**20 else "chunks"
    if mode == "flip" and (len(chunks) != 1 or n < 2):
        mode = "chunks"
    if mode == "chunks" and len(chunks) < 2:
        mode = "none"
    flip = mode == "flip"
    chunked = mode == "chunks"
    launches_per_step = 2 * len(chunks) if chunked else 1 + len(chunks)
    slot_ctr = [0]
    last_slots = []     # (ring slot index, first env, count) of the most recent step

    def hot_step(call_idx: int):
        last_slots.clear()
        if flip:
            # ONE C call per step (pb_rollout_random_step does the stream / event choreography
            # inside the library): k_step of this call overlaps k_encode_obs of the previous one
            s = slot_ctr[0] % n_slots
            slot_ctr[0] += 1
            last_slots.append((s, 0, n))
            env.rollout_random_step(ring[s][:n], call_idx=call_idx, actions_out=actions, mask_out=mask,
                                    reward_out=reward, done_out=done)
            return
        if chunked:
            # one C call per chunk (pb_rollout_chunk_step): the chunk's k_step runs in place on the
            # engine's step stream beside the encoders of the other chunks
            for first, cnt in chunks:
                s = slot_ctr[0] % n_slots
                slot_ctr[0] += 1
                last_slots.append((s, first, cnt))
                env.rollout_chunk_step(first, cnt, ring[s][:cnt], call_idx=call_idx, actions_out=actions,
                                       mask_out=mask, reward_out=reward, done_out=done)
            return
        # random valid action per env + PacmanGym.step, one launch (pb_step_random) ...
        env.step_random(call_idx=call_idx, actions_out=actions, mask_out=mask, reward_out=reward,
                        done_out=done)
        # ... then the observation of every env, one encoder launch per ring slot
        for first, cnt in chunks:
            s = slot_ctr[0] % n_slots
            slot_ctr[0] += 1
            last_slots.append((s, first, cnt))
            env.obs_range(first, cnt, ring[s][:cnt])

    def join():
        if flip or chunked:
            env.rollout_join()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident throughput (`value`) ----
    call = 0
    for _ in range(max(3, warmup)):
        hot_step(call)
        call += 1
    join()
    barrier()
    # Gate: ~1 ms of device-side sleep in front of the timed region, so that the host has the
    # first steps enqueued before the GPU starts on them (the timed region then measures the GPU,
    # not the launch latency of its first kernels).  No event is recorded inside the timed loop.
    torch.cuda._sleep(2_000_000)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    ev0.record()
    for _ in range(steps):
        hot_step(call)
        call += 1
    join()   # the current stream waits for the engine's streams
    ev1.record()
    torch.cuda.synchronize(dev)
    t_wall1 = time.time()
    ms_total = ev0.elapsed_time(ev1)
    barrier()
    clocks = sampler.summary(t_wall0, t_wall1) if sampler is not None else None
    episodes = int(done.sum().item())
    ms_by_rank = [ms_total / steps]
    if world > 1:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        every = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(every, t)
        ms_by_rank = [float(x.item()) / steps for x in every]
        ms_total = max(ms_by_rank) * steps
    value = n * world * steps / (ms_total * 1e-3)

    # ---- roofline of the dominant kernel: a SEPARATE short pass over the same calls, so that no
    # timing event sits inside the `value` loop.  flip: the window opens on the encode stream
    # AFTER the window's first encoder launch is enqueued and closes after 16 more, i.e. it spans
    # 16 back-to-back launches and nothing else (no host launch latency, no pipeline fill). ----
    enc_windows = []   # (ms, launches, envs)
    if flip:
        per_window = 16
        for _ in range(3):
            hot_step(call)
            call += 1
            # consecutive encoders alternate between two engine-owned streams:
            # rollout_streams()[1] is the stream of the encoder launch just enqueued
            e_open = torch.cuda.Event(enable_timing=True)
            e_open.record(env.rollout_streams()[1])
            for _ in range(per_window):
                hot_step(call)
                call += 1
            e_close = torch.cuda.Event(enable_timing=True)
            e_close.record(env.rollout_streams()[1])
            enc_windows.append((e_open, e_close, per_window, per_window * n))
        join()
    elif chunked:
        # one window per sweep: from the end of the sweep's first encoder launch to the end of its
        # last one (the encoders alternate between two streams; rollout_streams()[1] is the stream
        # of the launch just enqueued)
        for _ in range(3):
            opened = None
            for first, cnt in chunks:
                s = slot_ctr[0] % n_slots
                slot_ctr[0] += 1
                env.rollout_chunk_step(first, cnt, ring[s][:cnt], call_idx=call, actions_out=actions,
                                       mask_out=mask, reward_out=reward, done_out=done)
                if opened is None:
                    opened = torch.cuda.Event(enable_timing=True)
                    opened.record(env.rollout_streams()[1])
            e_close = torch.cuda.Event(enable_timing=True)
            e_close.record(env.rollout_streams()[1])
            call += 1
            enc_windows.append((opened, e_close, len(chunks) - 1, n - chunks[0][1]))
        join()
    else:
        for _ in range(max(2, min(8, 64 // len(chunks)))):
            env.step_random(call_idx=call, actions_out=actions, mask_out=mask, reward_out=reward,
                            done_out=done)
            call += 1
            e_open = torch.cuda.Event(enable_timing=True)
            e_open.record()
            for first, cnt in chunks:
                s = slot_ctr[0] % n_slots
                slot_ctr[0] += 1
                env.obs_range(first, cnt, ring[s][:cnt])
            e_close = torch.cuda.Event(enable_timing=True)
            e_close.record()
            enc_windows.append((e_open, e_close, len(chunks), n))
    torch.cuda.synchronize(dev)
    enc_ms = sum(a.elapsed_time(b) for a, b, _, _ in enc_windows)
    enc_launches = sum(w[2] for w in enc_windows)
    enc_envs = sum(w[3] for w in enc_windows)
    peak, peak_src = measured_peak()
    avg_launch_ms = enc_ms / enc_launches
    envs_per_launch = enc_envs / enc_launches
    enc_gbs = envs_per_launch * ENCODE_BYTES_PER_ENV / (avg_launch_ms * 1e-3) / 1e9
    roofline = {
        "kernel": "k_encode_obs",
        "bound": "hbm",
        "achieved": enc_gbs,
        "peak": peak,
        "unit": "GB/s",
        "frac": enc_gbs / peak,
        "traffic": encode_traffic(int(envs_per_launch)),
        "traffic_source": "profiles/encode_traffic.json (dram__bytes_read.sum + "
                          "dram__bytes_write.sum of one ncu --set full capture, per launch)",
        "peak_source": peak_src,
        "bytes_per_launch": int(envs_per_launch * ENCODE_BYTES_PER_ENV),
        "avg_launch_ms": avg_launch_ms,
        "envs_per_launch": int(envs_per_launch),
        "launches_timed": enc_launches,
        "how": ("CUDA events on the encode streams (end of launch i to end of launch i + 16) around "
                "windows of 16 back-to-back launches, in a separate pass right after the timed loop "
                "(same calls, k_step of the next step running beside the encoder as in the timed "
                "loop): the steady-state period of the launch" if flip else
                "CUDA events on the encode streams from the end of a sweep's first encoder launch to "
                "the end of its last one (31 back-to-back launches on two alternating streams, the "
                "k_step launches of the other chunks running beside them), in a separate pass right "
                "after the timed loop" if chunked else
                "CUDA events around the encoder launches of a step (one stream), in a separate "
                "pass right after the timed loop"),
    }

    # ---- verification (untimed): one more hot step replayed by the CPU oracle ----
    verified = None
    if verify:
        verified = verify_hot_step(cx, env, ring, hot_step, join, last_slots, call, n, seed, gid_base,
                                   actions, reward, done, mask)
        call += 1

    # ---- end to end through the public API with HOST buffers (`e2e`) ----
    e2e = None
    if want_e2e:
        # The host side owns the actions here (as a host-side policy would): a pool of pinned
        # action buffers, uniform over 0..4 (a blocked move is a legal no-op, env.rs:404), and
        # pinned result buffers; every step copies its actions H2D and its results D2H.
        rng = np.random.default_rng(seed + rank)
        pool = 8
        host_actions = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(pool)]
        for buf in host_actions:
            buf.numpy()[:] = rng.integers(0, 5, n, dtype=np.uint8)
        h_reward = torch.empty(n, dtype=torch.int32).pin_memory().numpy()
        h_done = torch.empty(n, dtype=torch.bool).pin_memory().numpy()
        h_mask = torch.empty((n, 5), dtype=torch.bool).pin_memory().numpy()
        outs = (h_reward, h_done, h_mask)
        checksum = 0

        def e2e_step(i: int):
            a = host_actions[i % pool].numpy()
            if flip:   # ONE pipelined C call: H2D, k_step, k_encode_obs, D2H (pb_rollout_host_step)
                s = slot_ctr[0] % n_slots
                slot_ctr[0] += 1
                env.rollout_host_step(a, ring[s][:n], outs)
            elif len(chunks) == 1:  # one C-ABI call: H2D actions, step, encode, D2H reward/done/mask
                s = slot_ctr[0] % n_slots
                slot_ctr[0] += 1
                env.step_host(a, obs_out=ring[s][:n], out=outs)
            elif chunked:  # one C call per chunk: actions H2D, k_step in place, encode, results D2H,
                # nothing blocks until the sweep's results are awaited
                for first, cnt in chunks:
                    s = slot_ctr[0] % n_slots
                    slot_ctr[0] += 1
                    env.rollout_chunk_host_step(first, cnt, a, ring[s][:cnt], outs)
                env.rollout_host_sync()
            else:  # env population larger than a ring slot: encode chunk by chunk; the result
                # copies of the step travel while the chunks are being encoded
                env.step_host(a, out=outs, begin_only=True)
                for first, cnt in chunks:
                    s = slot_ctr[0] % n_slots
                    slot_ctr[0] += 1
                    env.obs_range(first, cnt, ring[s][:cnt])
                env.step_host_end()

        for i in range(3):
            e2e_step(i)
        join()
        barrier()
        # every rank times ITS OWN steps (wall clock from its first call to the moment its last
        # observation is in HBM); the slowest rank defines the step time.  No collective sits
        # inside a rank's timed region.
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            e2e_step(i)
            checksum += int(h_reward[0]) + int(h_done[-1])   # the host reads the step's result
        join()
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        dt_by_rank = [dt]
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            every = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(every, t)
            dt_by_rank = [float(x.item()) for x in every]
            dt = max(dt_by_rank)
        e2e = {"value": n * world * e2e_steps / dt, "unit": "env-steps/s",
               "h2d_bytes_per_step": n, "d2h_bytes_per_step": n * (4 + 1 + 5), "steps": e2e_steps,
               "ms_per_step": 1e3 * dt / e2e_steps,
               "ms_per_step_by_rank": [round(1e3 * x / e2e_steps, 5) for x in dt_by_rank],
               "what": ("BatchedPacmanGym.rollout_host_step -> pb_rollout_host_step (pb_step_flip + "
                        "pb_encode_obs on two engine-owned streams)"
                        if flip else "BatchedPacmanGym.rollout_chunk_host_step -> "
                        "pb_rollout_chunk_host_step per 65 536-env chunk (actions H2D, k_step in place "
                        "on the step stream, pb_encode_obs_range on two alternating encode streams, "
                        "results D2H), pb_rollout_host_sync once per step"
                        if chunked else "BatchedPacmanGym.step_host -> pb_step_host"
                        if len(chunks) == 1 else "BatchedPacmanGym.step_host(begin_only) -> "
                        "pb_step_host_begin, pb_encode_obs_range per ring slot, pb_step_host_end")
                       + ": pinned host actions in, reward/done/mask back to pinned host memory and "
                         "read by the host every step (the D2H copies ride a side stream while the "
                         "encoder runs), observation encoded into the device-resident ring (it is "
                         "consumed on the GPU); wall clock per rank, max over ranks"}

    out = {
        "value": value, "ms_per_step": ms_total / steps, "ms_by_rank": ms_by_rank, "roofline": roofline,
        "e2e": e2e, "verified": verified, "clocks": clocks, "episodes": episodes, "mode": mode,
        "launches_per_step": launches_per_step, "n_slots": n_slots, "slot_bytes": slot_bytes,
        "peak": peak,
    }
    cx.last_env = env   # the e2e_obs_to_host leg reuses the main leg's engine
    return out


def verify_hot_step(cx, env, ring, hot_step, join, last_slots, call, n, seed, gid_base,
                    actions, reward, done, mask):
    """Untimed: take the states of a 4 096-env sample, run ONE more hot step through the same
    calls as the timed loop, and have the CPU oracle replay that step from those states with the
    same counter-RNG draws.  Everything the step produced for the sample must be identical:
    actions, reward, done, mask, full state, and the observation BITS in the ring."""
    torch, np, dist = cx.torch, cx.np, cx.dist
    from oracle import pyoracle

    pyoracle.build()
    sample_n = min(n, 4096)
    pieces = 4 if n >= 4 * 1024 else 1
    per = sample_n // pieces
    firsts = [(n // pieces) * p for p in range(pieces)]   # spread over both random_start halves
    # an env population larger than the ring (configs[2]: 32 encoder launches per step through a
    # 3-slot ring) leaves only the step's last launches resident: the observation bits are checked
    # on pieces inside those, everything else on the spread pieces as well
    obs_firsts = set(firsts)
    launches = list(last_slots)        # (slot, first env, count) of the previous step's launches
    if len(launches) > ring.shape[0]:
        extra = [first + (cnt - per) // 2 for _s, first, cnt in launches[-ring.shape[0]:]]
        obs_firsts = set(extra)
        firsts = firsts + extra
    join()
    torch.cuda.synchronize(cx.dev)
    before = [env.get_state(f, per) for f in firsts]
    hot_step(call)
    join()
    torch.cuda.synchronize(cx.dev)
    after = [env.get_state(f, per) for f in firsts]
    a_h, r_h = actions.cpu().numpy(), reward.cpu().numpy()
    d_h, m_h = done.cpu().numpy(), mask.cpu().numpy()
    detail, ok = [], True
    for f, st0, st1 in zip(firsts, before, after):
        ob = pyoracle.OracleBatch(per, st0["cfg_flags"].copy(), None)
        ob.import_state(pyoracle.flat_from_product(st0))
        ob.set_counter_rng(seed, gid_base + f)
        a_o = ob.sample_actions(seed, gid_base + f, call)
        r_o, d_o = ob.step(a_o, auto_reset=True)
        sl = slice(f, f + per)
        checks = {
            "actions": np.array_equal(a_o, a_h[sl]),
            "reward": np.array_equal(r_o, r_h[sl]),
            "done": np.array_equal(d_o, d_h[sl]),
            "mask": np.array_equal(ob.action_mask(), m_h[sl]),
            "state": pyoracle.states_differ(ob.export_state(), st1) is None,
        }
        # the observation the step wrote into the ring, bit for bit
        if f in obs_firsts:
            obs_ok = False
            for s, first, cnt in last_slots:
                if first <= f and f + per <= first + cnt:
                    got = ring[s][f - first:f - first + per].cpu().numpy().view(np.uint32)
                    obs_ok = np.array_equal(ob.obs().view(np.uint32), got)
            checks["obs_bits"] = obs_ok
        if not all(checks.values()):
            ok = False
            detail.append({"first_env": f, **{k: bool(v) for k, v in checks.items()}})
    if cx.world > 1:
        t = torch.tensor([1 if ok else 0], device=cx.dev, dtype=torch.int32)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = bool(t.item())
    cx.verify_detail = {
        "sample": f"{len(firsts)} x {per} envs per rank, one extra untimed hot step after the timed loop, "
                  "replayed by the CPU oracle (oracle/) from the sampled states with the same "
                  "counter-RNG draws: actions, reward, done, mask and full state compared for every "
                  f"piece, the observation bits of the ring slot for {len(obs_firsts)} of them"
                  + ("" if len(obs_firsts) == len(firsts) else
                     " (the pieces inside the step's last encoder launches: earlier launches' ring "
                     "slots were reused within the step)"),
        "failures": detail,
    }
    return ok


# =================================================================================================
# DQN legs (BASELINE.json configs[3] and configs[4])
# =================================================================================================
def dqn_trainer(dev, seed: int, graph: bool, rank: int = 0, world: int = 1, extra=()):
    from pacbot_rl_b200 import train_dqn

    cfg = train_dqn.build_parser().parse_args(
        ["--device", str(dev), "--seed", str(seed), "--no-checkpoints", "--log_every", "1000000",
         *extra])
    trainer = train_dqn.DQNTrainer(cfg, dev, rank, world)
    trainer.replay.fill()
    trainer.epsilon = cfg.initial_epsilon
    if graph:
        trainer.capture_graph()
    return trainer, cfg


def dqn_time(cx, trainer, iters: int, with_eval: bool):
    """iterations/s of `train_iteration` (+ the reference's evaluation cadence: an evaluation
    episode every `evaluate_steps` = 10 iterations, train_dqn.py:220-229)."""
    torch = cx.torch
    for _ in range(10):
        trainer.train_iteration()
    if with_eval:
        trainer.maybe_evaluate(force=True)
        trainer.finish_evaluation()
    torch.cuda.synchronize(cx.dev)
    if cx.world > 1:
        cx.dist.barrier()
    t0 = time.perf_counter()
    for _ in range(iters):
        trainer.train_iteration()
        if with_eval:
            trainer.maybe_evaluate()
    if with_eval:
        trainer.finish_evaluation()
    torch.cuda.synchronize(cx.dev)
    dt = time.perf_counter() - t0
    if cx.world > 1:
        t = torch.tensor([dt], device=cx.dev, dtype=torch.float64)
        cx.dist.all_reduce(t, op=cx.dist.ReduceOp.MAX)
        dt = float(t.item())
    return iters / dt


def dqn_leg(cx, seed: int, iters: int):
    """configs[3]: train_dqn defaults with GPU-resident envs on one B200."""
    trainer, cfg = dqn_trainer(cx.dev, seed, graph=True)
    with_eval = dqn_time(cx, trainer, iters, True)
    no_eval = dqn_time(cx, trainer, iters, False)
    m = trainer.pop_metrics()
    evals = trainer.eval_log[-3:]
    del trainer
    eager, _ = dqn_trainer(cx.dev, seed, graph=False)
    eager_rate = dqn_time(cx, eager, max(20, iters // 4), False)
    return {
        "iters_per_sec": with_eval, "iters": iters, "cuda_graph": True,
        "eval_cadence": f"one greedy evaluation episode every evaluate_steps = {cfg.evaluate_steps} "
                        "iterations as in train_dqn.py:220-229, run on a side stream with a weight "
                        "snapshot while training goes on",
        "iters_per_sec_no_eval": no_eval,
        "iters_per_sec_eager_no_eval": eager_rate,
        "env_steps_per_sec": with_eval * cfg.experience_steps * cfg.num_parallel_envs,
        "config": "train_dqn defaults: batch 512, 32 envs, 4 experience steps/iter, buffer 10k, "
                  "QNetV2 fp32 (cuDNN, NHWC), double DQN, Adam",
        "loss": m["loss"], "last_evals": evals,
    }


def dqn_dp_leg(cx, seed: int, iters: int):
    """configs[4]: data-parallel DQN, envs sharded per GPU, ONE NCCL all-reduce of the flat
    156 934-float gradient per iteration (inside the captured graph)."""
    torch, dist = cx.torch, cx.dist
    trainer, cfg = dqn_trainer(cx.dev, seed, graph=True, rank=cx.rank, world=cx.world)
    dp = dqn_time(cx, trainer, iters, False)
    dp_eval = dqn_time(cx, trainer, iters, True)
    del trainer
    # the same iteration without the all-reduce (every rank trains alone): what the collective
    # and the rank skew it exposes cost together
    local, _ = dqn_trainer(cx.dev, seed, graph=True, rank=cx.rank, world=1)
    alone = dqn_time(cx, local, iters, False)
    del local
    # the collective by itself: 628 KB fp32 all-reduce, back to back
    flat = torch.zeros(156934, device=cx.dev)
    for _ in range(20):
        dist.all_reduce(flat)
    torch.cuda.synchronize(cx.dev)
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 200
    for _ in range(reps):
        dist.all_reduce(flat)
    e1.record()
    torch.cuda.synchronize(cx.dev)
    ar_us = 1e3 * e0.elapsed_time(e1) / reps
    t = torch.tensor([ar_us], device=cx.dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ar_us = float(t.item())
    ms_dp, ms_alone = 1e3 / dp, 1e3 / alone
    return {
        "iters_per_sec": dp, "iters_per_sec_with_eval": dp_eval, "iters": iters,
        "global_batch": cfg.batch_size * cx.world, "envs_total": cfg.num_parallel_envs * cx.world,
        "env_steps_per_sec": dp * cfg.experience_steps * cfg.num_parallel_envs * cx.world,
        "samples_per_sec": dp * cfg.batch_size * cx.world,
        "iters_per_sec_without_allreduce": alone,
        "ms_per_iter": ms_dp, "ms_per_iter_without_allreduce": ms_alone,
        "allreduce_us_isolated": ar_us,
        "allreduce_share_of_iteration": max(0.0, (ms_dp - ms_alone) / ms_dp),
        "rank_skew_us_per_iter": max(0.0, 1e3 * (ms_dp - ms_alone) - ar_us),
        "what": "train_dqn defaults per rank (batch 512, 32 envs), one CUDA graph per iteration with "
                "the NCCL all-reduce of the 628 KB flat gradient inside; slowest rank defines the rate",
    }


def mcts_throughput(dev, seed: int, num_envs: int, tree_size: int, steps: int, graph: bool,
                    cpu_steps: int = 0):
    """the AlphaZero experience collector (alphazero.rs generate_experience_step:
    one MCTS search iteration per tree -- select, leaf observation encode, NetV2 evaluation,
    backprop -- plus the env steps / re-roots / resets that become due), simulations per second."""
    import torch

    from pacbot_rl_b200 import alphazero, models

    model = models.NetV2((17, 28, 31), 6).to(dev).use_fused_blocks().eval()

    @torch.no_grad()
    def evaluator(obs, mask):  # train_alphazero.py:83-95 (reward_scale 1/100)
        return model.policy_value(obs, mask, 100.0)

    cfg = alphazero.AlphaZeroConfig(tree_size, 1000, 0.99, num_envs)
    col = alphazero.BatchedExperienceCollector(evaluator, cfg, device=dev, seed=seed, leaf_layout="nhwc32")
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
           "config": f"{num_envs} trees x tree_size {tree_size}, NetV2 fp32 evaluator (cuDNN convolutions, NHWC, fused block tails and head) on "
                     "device-resident leaf observations written channels-last by the encoder, Dirichlet root noise, max_episode_length 1000"}
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


def dropin_leg(cx, seed: int, steps: int = 3000):
    """configs[0]'s call pattern on the B200 engine: ONE env object driven the reference's way
    (src/replay_buffer.py:107-137 -- action_mask(), step(), is_done(), obs_numpy(), action_mask()
    per env-step, reset() on done), through the per-env drop-in `PacmanGym`.  Wall clock per
    env-step; the old path (one device round trip per question) is timed beside it."""
    np, pb = cx.np, cx.pb
    rng = np.random.default_rng(seed)

    def drive(gym, k, ask):
        t0 = time.perf_counter()
        for _ in range(k):
            mask = ask["mask"](gym)
            valid = [a for a in range(5) if mask[a]]
            _reward, done = gym.step(valid[int(rng.integers(len(valid)))])
            if ask["done"](gym) or done:
                gym.reset()
            ask["obs"](gym)
            ask["mask"](gym)
        return (time.perf_counter() - t0) / k

    gym = pb.PacmanGym(True, True, device=cx.dev, seed=seed, env_id=777)
    gym.reset()
    new_way = {"mask": lambda g: g.action_mask(), "done": lambda g: g.is_done(), "obs": lambda g: g.obs_numpy()}
    surface = pb.gym._SingleEnvSurface     # the general per-question entry points (what views use)
    old_way = {"mask": lambda g: surface.action_mask(g), "done": lambda g: surface.is_done(g),
               "obs": lambda g: surface.obs_numpy(g)}
    drive(gym, 200, new_way)
    us_new = 1e6 * drive(gym, steps, new_way)
    drive(gym, 50, old_way)
    us_old = 1e6 * drive(gym, max(200, steps // 10), old_way)
    return {"us_per_env_step": us_new, "env_steps_per_sec": 1e6 / us_new,
            "us_per_env_step_one_round_trip_per_question": us_old,
            "what": "pacbot_rl_b200.PacmanGym, one env per object: action_mask + step + is_done + "
                    "obs_numpy + action_mask per env-step as replay_buffer.py:107-137 calls them; step() "
                    "is ONE device round trip (pb_step_snapshot_host) and the getters answer from its "
                    "snapshot"}


def obs_to_host_leg(cx, env, n: int):
    """The reference loop's own per-env call (`obs_numpy`, env.rs:305-307) at batch scale: encode +
    D2H of full observations into PINNED host memory, chunked and double buffered
    (pb_obs_host); GB/s against a plain pinned cudaMemcpy of the same size on this box."""
    torch, np = cx.torch, cx.np
    n_small = min(n, 16384)
    host = torch.empty((n_small,) + cx.pb.OBS_SHAPE, dtype=torch.float32).pin_memory()
    env.obs_numpy(0, n_small, out=host.numpy())
    reps = 3
    t0 = time.perf_counter()
    for _ in range(reps):
        env.obs_numpy(0, n_small, out=host.numpy())
    dt = (time.perf_counter() - t0) / reps
    # PCIe reference: one plain D2H copy of the same bytes from a device tensor
    devbuf = torch.empty((n_small,) + cx.pb.OBS_SHAPE, dtype=torch.float32, device=cx.dev)
    host.copy_(devbuf)
    torch.cuda.synchronize(cx.dev)
    t0 = time.perf_counter()
    for _ in range(reps):
        host.copy_(devbuf, non_blocking=True)
        torch.cuda.synchronize(cx.dev)
    dt_copy = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    pageable = env.obs_numpy(0, min(n_small, 4096))
    dt_page = time.perf_counter() - t0
    gb = n_small * OBS_BYTES / 1e9
    return {"value": n_small / dt, "unit": "obs/s", "d2h_bytes_per_obs": OBS_BYTES,
            "gb_per_s": gb / dt, "pcie_d2h_gb_per_s_plain_copy": gb / dt_copy,
            "frac_of_plain_copy": dt_copy / dt,
            "pageable_obs_per_s": len(pageable) / dt_page,
            "what": f"obs_numpy(out=pinned) of {n_small} envs: k_encode_obs in chunks into two device "
                    "staging buffers, each chunk's D2H copy overlapping the next chunk's encode"}


def main():
    global _REAL_STDOUT
    args = parse_args()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    cpus = None if args.no_affinity else bind_rank_to_cpus(local_rank, local_world)

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
    sampler = ClockSampler(local_rank) if rank == 0 else None

    cx = Ctx()
    cx.torch, cx.np, cx.pb, cx.dist = torch, np, pb, dist
    cx.dev, cx.rank, cx.world = dev, rank, world
    cx.verify_detail = None
    cx.last_env = None

    n = args.envs_per_gpu
    warmup = max(3, args.warmup)
    e2e_steps = max(3, min(max(args.steps, 200), 500))
    main_leg = rollout_leg(cx, n, args.steps, warmup, args.ring_gb, args.chunk_envs, args.pipeline,
                           args.seed, want_e2e=not args.no_e2e, e2e_steps=e2e_steps,
                           verify=not args.no_verify, sampler=sampler)
    main_detail = cx.verify_detail
    e2e_full = None
    if rank == 0 and world == 1 and not args.no_e2e:
        e2e_full = obs_to_host_leg(cx, cx.last_env, n)
    cx.last_env = None
    torch.cuda.empty_cache()

    # ---- configs[2]: 2 097 152 envs per GPU, the observation streamed through the ring in
    # 65 536-env chunks (the whole population's obs would be 118 GB) ----
    config2 = None
    if not args.no_config2 and n != CONFIG2_ENVS:
        k2 = max(8, args.config2_steps)
        leg = rollout_leg(cx, CONFIG2_ENVS, k2, 3, args.ring_gb, 65536,
                          "none" if args.pipeline == "none" else "chunks", args.seed + 1,
                          want_e2e=not args.no_e2e, e2e_steps=k2, verify=not args.no_verify)
        config2 = {
            "workload": f"{CONFIG2_ENVS} envs/GPU x {world} GPU(s) = {CONFIG2_ENVS * world} envs "
                        "(BASELINE.json configs[2]), same step + full f32 observation encode, 32 encoder "
                        "launches of 65 536 envs per step into the ring"
                        + (", each chunk's k_step in place beside the other chunks' encoders"
                           if leg["mode"] == "chunks" else ""),
            "value": leg["value"], "unit": "env-steps/s", "steps": k2, "warmup": 3,
            "ms_per_step": leg["ms_per_step"],
            "ms_per_step_by_rank": [round(x, 4) for x in leg["ms_by_rank"]],
            "hbm_roofline_frac_whole_step": (leg["value"] / world) * BYTES_PER_ENV_STEP / 1e9 / leg["peak"],
            "roofline": leg["roofline"], "pipeline": leg["mode"],
            "gpu_launches": leg["launches_per_step"] * k2 * world,
            "verified": leg["verified"], "verify_failures": (cx.verify_detail or {}).get("failures"),
        }
        if leg["e2e"] is not None:
            config2["e2e"] = leg["e2e"]
        torch.cuda.empty_cache()

    dqn = dqn_dp = None
    if not args.no_dqn:
        if world == 1:
            dqn = dqn_leg(cx, args.seed, args.dqn_iters)
        else:
            dqn_dp = dqn_dp_leg(cx, args.seed, args.dqn_iters)

    if sampler is not None:
        sampler.stop()
    if rank != 0:
        shutdown(dist, world)
        return

    cfg = workload_config(args, world)
    line = {
        "metric": "env_steps_per_sec_incl_obs",
        "value": main_leg["value"],
        "unit": "env-steps/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": warmup,
        "ms_per_step": main_leg["ms_per_step"],
        "ms_per_step_by_rank": [round(x, 5) for x in main_leg["ms_by_rank"]],
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "int32+f32",
        "data": "synthetic",
        "config": cfg,
        "hbm_roofline_frac_whole_step": (main_leg["value"] / world) * BYTES_PER_ENV_STEP / 1e9 / main_leg["peak"],
        "roofline": main_leg["roofline"],
        "pipeline": {
            "flip": "engine-owned streams, double-buffered state: k_step of step t+1 overlaps "
                    "k_encode_obs of step t, consecutive encoders alternate between two streams (one "
                    "pb_rollout_random_step call per step)",
            "chunks": "engine-owned streams, in place: k_step of a chunk runs beside the encoders of "
                      "the other chunks, consecutive encoders alternate between two streams (one "
                      "pb_rollout_chunk_step call per chunk)",
            "none": "none: k_step then k_encode_obs on one stream",
        }[main_leg["mode"]],
        "gpu_launches": main_leg["launches_per_step"] * args.steps * world,
        "clocks": main_leg["clocks"],
        "episodes_finished_last_step": main_leg["episodes"],
        "host_cpus_bound": cpus,
    }
    if main_leg["verified"] is not None:
        line["verified"] = main_leg["verified"]
        line["verify"] = main_detail
    if main_leg["e2e"] is not None:
        line["e2e"] = main_leg["e2e"]
    if e2e_full is not None:
        line["e2e_obs_to_host"] = e2e_full
    if config2 is not None:
        line["config2"] = config2
    if dqn is not None:
        # the drop-in claim, timed: the reference's unmodified train_dqn.py + replay_buffer.py with
        # this engine's per-env `PacmanGym` as its `pacbot_rs.PacmanGym` (32 one-env engines, the
        # reference's own host loop: obs_numpy / action_mask / step per env per step)
        dqn["reference_loop_on_b200_engine"] = reference_dqn(args, env="product")
        dqn["dropin_per_env"] = dropin_leg(cx, args.seed)
        line["dqn"] = dqn
    if dqn_dp is not None:
        line["dqn_dp"] = dqn_dp
    if world == 1 and not args.no_mcts:
        cpu_steps = 0 if args.no_cpu_baseline else 400
        line["mcts"] = mcts_throughput(dev, args.seed, 128, 100, args.mcts_steps, True, cpu_steps)
        line["mcts_4096"] = mcts_throughput(dev, args.seed, 4096, 100, max(50, args.mcts_steps // 4), True)
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args.cpu_sample_envs, args.cpu_sample_steps,
                                            os.cpu_count() or 1, args.seed)
    emit(line)
    shutdown(dist, world)


def shutdown(dist, world: int) -> None:
    """Tear the process group down, but never sit in it: destroy_process_group() has been seen to
    wait forever while a CUDA graph with captured NCCL kernels is still alive.  The JSON line is
    out by now; a rank that is still here after 20 s leaves with exit code 0."""
    if world <= 1:
        return
    sys.stdout.flush()
    sys.stderr.flush()
    timer = threading.Timer(20.0, lambda: os._exit(0))
    timer.daemon = True
    timer.start()
    dist.destroy_process_group()
    timer.cancel()


if __name__ == "__main__":
    main()
