#!/usr/bin/env python
"""bench.py — headline benchmark of the batched offline bounce path.

Workload (BASELINE.json configs[1], "C2"): a 4096-patch drum sweep (1024 each of kick / snare / hi-hat / tom,
every FFI-reachable parameter drawn U[0,1), one trigger at frame 0), 2 s each at 44.1 kHz = 88 200 samples per
voice, per-voice envelopes + filters, per-voice f32 output kept.  One "step" = one full render of the sweep.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

* ours: `value` = voice-samples/s with patches and output resident in HBM (device time, CUDA events on the
  library's launching streams, max over ranks); `e2e` = the same through the C ABI with HOST buffers
  (event tables H2D + kernels + D2H of every voice's audio into pinned host memory inside the timed region; the
  library drains finished 8192-frame chunks on a copy stream while later chunks render).
  `roofline` is the contract's store-bandwidth figure for the dominant back-end kernel (algorithmic bytes per launch /
  its mean launch duration, CUDA events inside the library; `traffic` from the committed ncu capture); `kernels` lists
  all four back ends.  The kernels are latency-bound, not byte-bound: DESIGN.md section 4 and profiles/README.md.
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

# The library drives ~14 streams (4 type buckets x {front, back, general} + copy); the default 8 hardware queues alias
# them and serialise independent kernels.  Must be set before the CUDA context exists (libgooey_b200 sets it too).
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

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

    def mark(self):
        self.first = max(len(self.rows) - 1, 0)   # keep the row in flight when the timed region starts

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        rows = self.rows[getattr(self, "first", 0):] or self.rows[-1:]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


KERNELS = ["wave_kernel<KickW>", "wave_kernel<SnareW>", "wave_kernel<HatW>", "wave_kernel<TomW>"]


def kernel_stats(L):
    """Per back-end kernel: launches, mean launch duration (CUDA events on the launching stream, measured inside the
    library over the timed region) and voice-frames per launch."""
    L.gooey_b200_kernel_stat.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    out = {}
    for k in KERNELS:
        n, ms, vf = ctypes.c_uint64(0), ctypes.c_double(0.0), ctypes.c_double(0.0)
        L.gooey_b200_kernel_stat(k.encode(), ctypes.byref(n), ctypes.byref(ms), ctypes.byref(vf))
        if n.value:
            out[k] = {"launches": int(n.value), "avg_ms": ms.value / n.value, "total_ms": ms.value, "voice_frames_per_launch": vf.value / n.value}
    return out


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full` capture."""
    p = os.path.join(ROOT, "profiles", "ncu_dram_traffic.json")
    if not os.path.exists(p):
        return None, None
    with open(p) as f:
        d = json.load(f)
    e = d.get(kernel)
    return (e["dram_bytes_per_launch"], e.get("source")) if e else (None, None)


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
        # NCCL prints its version banner (and any NCCL_DEBUG output) to stdout; stdout carries exactly one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
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

    sampler = ClockSampler(dev)
    sampler.start()                      # nvidia-smi needs ~0.3 s to produce its first row: start it under the warm-up load
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler.mark()                       # only rows sampled from here on (the timed region) are reported
    L.gooey_b200_kernel_stats_reset()
    launches_before = L.gooey_b200_launch_count()
    t0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        dev_ms += step_device()
    barrier()
    wall_dev = time.perf_counter() - t0
    launches = L.gooey_b200_launch_count() - launches_before
    clocks = sampler.stop()
    kstats = kernel_stats(L)

    # end-to-end through the C ABI with host buffers
    for _ in range(max(args.warmup, 3 if world == 1 else 6)):   # the host-buffer path has its own first-call costs (staging allocations, page touch; with 8 ranks draining at once the first steps were 160 / 140 / 100 ms)
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_steps, e2e_kernel_ms = [], []
    for _ in range(args.steps):
        t1 = time.perf_counter()
        step_e2e()
        e2e_steps.append((time.perf_counter() - t1) * 1e3)
        e2e_kernel_ms.append(L.gooey_b200_last_kernel_ms())
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
        # dominant kernel = the back-end launch with the largest share of device time.  achieved = algorithmic bytes per
        # launch (4 B stored per voice-frame) / its mean launch duration, both measured over the timed region.
        # (The four buckets overlap on the device, so live durations include contention; the serialised ncu launch list
        # in profiles/ ranks wave_kernel<TomW> first, and it is kept as the dominant kernel while it is within 10 % of the
        # live maximum so that both views name the same kernel.)
        dom = max(kstats, key=lambda k: kstats[k]["total_ms"]) if kstats else None
        if dom and "wave_kernel<TomW>" in kstats and kstats["wave_kernel<TomW>"]["total_ms"] >= 0.9 * kstats[dom]["total_ms"]:
            dom = "wave_kernel<TomW>"
        if dom:
            bytes_per_launch = kstats[dom]["voice_frames_per_launch"] * BYTES_PER_VOICE_SAMPLE
            achieved = bytes_per_launch / (kstats[dom]["avg_ms"] * 1e-3) / 1e9
        else:
            bytes_per_launch, achieved = None, N_PATCHES * FRAMES * BYTES_PER_VOICE_SAMPLE / (dev_ms / args.steps * 1e-3) / 1e9
        traffic, traffic_src = ncu_traffic(dom) if dom else (None, None)
        cores = os.cpu_count() or 1
        n_sample = min(N_PATCHES, 256 * cores)          # 16 cores: the whole sweep, ~7 s of CPU work
        cpu_v, cpu_dt = cpu_port_throughput(n_sample, cores)
        line = {
            "metric": "voice-samples/sec", "value": value, "unit": "voice-samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C2: 4096-patch drum sweep (kick/snare/hihat/tom), 88200 samples each @44.1kHz", "seed": SEED,
                       "voices_per_gpu": N_PATCHES, "frames": FRAMES, "sample_rate": SR,
                       "l2": "output 1.44 GB per step >> 126 MB L2; nothing is re-read between steps",
                       "parallelism": f"independent voice shards x{world}, no collective"},
            "e2e": {"value": e2e, "unit": "voice-samples/s", "h2d_bytes_per_step": int(world * N_PATCHES * (16 + 12)),
                    "d2h_bytes_per_step": int(world * N_PATCHES * FRAMES * 4), "ms_per_step": wall_e2e / args.steps * 1e3,
                    "ms_each_step_rank0": [round(x, 2) for x in e2e_steps],
                    "kernel_ms_each_step_rank0": [round(x, 2) for x in e2e_kernel_ms]},
            "gpu_launches": int(launches),
            "wall_ms_per_step_device_resident": wall_dev / args.steps * 1e3,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_kind": peak_kind, "kernel": dom, "algorithmic_bytes_per_launch": bytes_per_launch,
                         "avg_launch_ms": kstats[dom]["avg_ms"] if dom else None, "traffic_source": traffic_src,
                         "whole_step_store_gbs": N_PATCHES * FRAMES * BYTES_PER_VOICE_SAMPLE / (dev_ms / args.steps * 1e-3) / 1e9,
                         "note": "HBM is the contract's roofline for this store-only path, but the voice kernels are bound by dependent-issue latency (replayed recurrences + shuffle scans), not by bytes: see DESIGN.md section 4"},
            "kernels": kstats,
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
