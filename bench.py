#!/usr/bin/env python
"""bench.py — headline benchmark of the batched offline bounce path.

Workload (BASELINE.json configs[1], "C2"): a 4096-patch drum sweep (1024 each of kick / snare / hi-hat / tom,
every FFI-reachable parameter drawn U[0,1), one trigger at frame 0), 2 s each at 44.1 kHz = 88 200 samples per
voice, per-voice envelopes + filters, per-voice f32 output kept.  One "step" = one full render of the sweep.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

* ours: `value` = voice-samples/s with patches and output resident in HBM (device time, CUDA events on the
  library's launching streams, max over ranks); `e2e` = the same through the C ABI with HOST buffers
  (event tables H2D + kernels + D2H of every voice's audio into pinned host memory inside the timed region).
* reference: the reference's CPU implementation of the same path.  The Rust reference cannot be built here
  (no toolchain in the image), so this arm times the C++ restatement in oracle/ ("port") on all host cores,
  on a bounded sample of the same patches.
Multi-GPU (torchrun, one rank per GPU): voices shard with no collective; weak scaling — every rank renders its
own 4096-patch sweep (rank-dependent seed); value = all ranks' voice-samples / max-over-ranks time.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SR = 44100.0
N_PATCHES = 4096
FRAMES = 88200
SEED = 0x600E7
# Algorithmic bytes per voice-sample for this workload: one f32 stored, nothing loaded.
BYTES_PER_VOICE_SAMPLE = 4


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_port_throughput(n_sample, threads):
    """C++ restatement of the reference render (oracle/) on `threads` host threads over the first
    n_sample patches of the workload; returns voice-samples/s."""
    import oracle_lib as O
    from workloads import drum_sweep_patches
    patches, vel, _ = drum_sweep_patches(n_sample, seed=SEED)
    trig = [(i, 0, float(vel[i])) for i in range(n_sample)]
    O.render_voices(patches[:4], 2048, triggers=trig[:4], threads=1)  # warm the library
    t0 = time.perf_counter()
    O.render_voices(patches, FRAMES, triggers=trig, threads=threads)
    dt = time.perf_counter() - t0
    return n_sample * FRAMES / dt, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_sample = 32 * cores
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt = cpu_port_throughput(n_sample, cores)
        if i >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals])) * 1e3
    line = {
        "impl": "reference", "metric": "voice-samples/sec", "value": value, "unit": "voice-samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2: 4096-patch drum sweep (kick/snare/hihat/tom), 88200 samples each @44.1kHz", "seed": SEED},
        "cpu_baseline": {"value": value, "unit": "voice-samples/s", "cores": cores, "kind": "port",
                         "sample": f"first {n_sample} of the 4096 patches x {FRAMES} frames per step, one voice per worker thread"},
        "e2e": {"value": value, "unit": "voice-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "Rust reference not buildable in this image (no cargo/rustc); C++ restatement of the reference render (oracle/), all host cores",
    }
    print(json.dumps(line), flush=True)


def run_ours(args, rank, world, local_rank):
    import torch
    from libgooey_b200 import lib, voices as V
    from workloads import drum_sweep_patches
    L = lib()
    if L.gooey_b200_device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device — libgooey_b200 has no CPU fallback")
    dev = local_rank
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    patches, vel, kinds = drum_sweep_patches(N_PATCHES, seed=SEED + rank)
    batch = V.VoiceBatch(patches, SR, device=dev)
    stride = FRAMES  # multiple of 4
    out_dev = torch.empty((N_PATCHES, stride), dtype=torch.float32, device=f"cuda:{dev}")
    out_host = torch.empty((N_PATCHES, FRAMES), dtype=torch.float32).pin_memory()
    out_np = out_host.numpy()

    launches0 = L.gooey_b200_launch_count()

    def step_device():
        batch.trigger_all(0, vel)
        batch.render_device(FRAMES, out_dev.data_ptr(), stride)
        return L.gooey_b200_last_kernel_ms()

    def step_e2e():
        batch.trigger_all(0, vel)
        batch.render(FRAMES, out_np)

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(dev)
    barrier()
    sampler.start()
    launches_before = L.gooey_b200_launch_count()
    t0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        dev_ms += step_device()
    barrier()
    wall_dev = time.perf_counter() - t0
    launches = L.gooey_b200_launch_count() - launches_before
    clocks = sampler.stop()

    # end-to-end through the C ABI with host buffers
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    wall_e2e = time.perf_counter() - t0
    checksum = float(np.abs(out_np[:, ::97]).sum())

    t = torch.tensor([dev_ms, wall_dev, wall_e2e], dtype=torch.float64, device=f"cuda:{dev}")
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall_dev, wall_e2e = (float(x) for x in t.tolist())

    if rank == 0:
        units = world * N_PATCHES * FRAMES * args.steps
        value = units / (dev_ms * 1e-3)
        e2e = units / wall_e2e
        peak, peak_kind = measured_peaks()
        # dominant kernel = the type bucket that sets the critical path; the four buckets run concurrently, so
        # the step's device time is the dominant bucket's launch duration.
        achieved = N_PATCHES * FRAMES * BYTES_PER_VOICE_SAMPLE / (dev_ms / args.steps * 1e-3) / 1e9
        cores = os.cpu_count() or 1
        n_sample = 16 * cores
        cpu_v, cpu_dt = cpu_port_throughput(n_sample, cores)
        line = {
            "metric": "voice-samples/sec", "value": value, "unit": "voice-samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C2: 4096-patch drum sweep (kick/snare/hihat/tom), 88200 samples each @44.1kHz", "seed": SEED,
                       "voices_per_gpu": N_PATCHES, "frames": FRAMES, "sample_rate": SR,
                       "l2": "output 1.44 GB per step >> 126 MB L2; nothing is re-read between steps",
                       "parallelism": f"independent voice shards x{world}, no collective"},
            "e2e": {"value": e2e, "unit": "voice-samples/s", "h2d_bytes_per_step": int(N_PATCHES * (12 + 8)),
                    "d2h_bytes_per_step": int(N_PATCHES * FRAMES * 4), "ms_per_step": wall_e2e / args.steps * 1e3},
            "gpu_launches": int(launches),
            "wall_ms_per_step_device_resident": wall_dev / args.steps * 1e3,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "peak_kind": peak_kind,
                         "note": "whole-step store rate; the dominant buckets (kick/snare additive oscillators) are FP32/FP64-pipe bound, see DESIGN.md"},
            "cpu_baseline": {"value": cpu_v, "unit": "voice-samples/s", "cores": cores, "kind": "port",
                             "sample": f"first {n_sample} of the 4096 patches x {FRAMES} frames, {cores} threads, {cpu_dt:.1f} s"},
            "checksum": checksum,
        }
        print(json.dumps(line), flush=True)
    batch.close()
    if dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import __graft_entry__ as g
    if rank == 0 or not os.path.exists(os.path.join(ROOT, "libgooey_b200", "lib", "libgooey_b200.so")):
        g.build()
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
