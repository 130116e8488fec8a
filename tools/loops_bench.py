"""Sample-playback sources at batch size: N FFI engines, each with two loop channels over ONE shared 10 s stereo loop
(varispeed + sample-rate conversion on one, a Resample tempo warp on a wrap-around window on the other) and a sampler rack with four
pads fired at the start, bounced `bars` bars in one device pass.  Prints one JSON object: device / end-to-end times, engine-samples/s,
the bytes the path moves per engine-sample (8 B stored per source row pair + the gathers), a spot-check against the oracle and the
oracle's own time for one engine on one host core.  Run by bench.py in a subprocess (rank 0, N = 1) so that nothing here can disturb
the headline line; also usable on its own:  python tools/loops_bench.py [engines] [bars]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
SR = 44100.0


def source(seed=0x600E7):
    rng = np.random.default_rng(seed)
    n = int(10 * 48000)
    t = np.arange(n) / 48000.0
    left = 0.4 * np.sin(2 * np.pi * 110.0 * t) * (0.6 + 0.4 * np.sin(2 * np.pi * 0.5 * t)) + 0.1 * rng.uniform(-1, 1, n)
    right = 0.4 * np.sin(2 * np.pi * 165.0 * t + 0.3) + 0.1 * rng.uniform(-1, 1, n)
    pads = [(0.5 * np.sin(2 * np.pi * f * np.arange(m) / SR) * np.exp(-np.arange(m) / (0.2 * m))).astype(np.float32)
            for f, m in ((60.0, 22050), (180.0, 11025), (400.0, 6000), (3000.0, 3000))]
    return np.stack([left, right], 1).astype(np.float32), pads


def script(e, i, src, pads, first=None):
    """The same calls on the product and on the oracle (the product shares engine 0's buffer on the device)."""
    rng = np.random.default_rng(1000 + i)
    e.set_bpm(120.0)
    if first is None or not e.loop_share_buffer(0, first, 0):
        assert e.loop_load(0, src, 48000.0)
    if first is None or not e.loop_share_buffer(1, first, 0):
        assert e.loop_load(1, src, 48000.0)
    e.loop_set_speed(0, float(rng.uniform(0.5, 1.5))); e.loop_set_gain(0, float(rng.uniform(0.3, 0.9)))
    e.loop_set_start(0, float(rng.uniform(0.0, 0.4))); e.loop_set_end(0, float(rng.uniform(0.6, 1.0)))
    e.loop_restart(0); e.loop_set_playing(0, True)
    e.loop_set_start(1, 0.8); e.loop_set_end(1, 0.2); e.loop_restart(1)
    e.loop_set_source_bpm(1, float(rng.uniform(90.0, 150.0))); e.loop_set_pitch_mode(1, 1); e.loop_set_gain(1, 0.5)
    e.loop_set_playing(1, True)
    r = e.sampler_register()
    e.mixer_route_source(5 + r, 0)
    for slot, p in enumerate(pads):
        e.sampler_set_slot_buffer(r, slot, p, SR)
        e.sampler_trigger(r, slot, float(rng.uniform(0.4, 1.0)))
    e.mixer_set_track_pan(3, float(rng.uniform(0.3, 0.7)))


def main():
    n_eng = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    bars = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    from libgooey_b200 import engine as G
    from libgooey_b200._lib import lib
    import oracle_lib as O
    L = lib()
    L.gooey_b200_last_kernel_ms.restype = __import__("ctypes").c_float
    src, pads = source()
    frames = int(round(bars * 4.0 * 0.5 * SR))

    def build():
        engines = [G.Engine() for _ in range(n_eng)]
        for i, e in enumerate(engines):
            script(e, i, src, pads, None if i == 0 else engines[0])
        return engines
    warm = build()                                   # one-time allocations and uploads on a throw-away batch
    G.batch_bounce_host(warm, bars)
    for e in warm:
        e.close()
    engines = build()
    t0 = time.perf_counter()
    out = G.batch_bounce_host(engines, bars)
    wall = time.perf_counter() - t0
    dev_ms = float(L.gooey_b200_last_kernel_ms())
    errs = [bool(e.has_error()) for e in engines]
    for e in engines:
        e.close()
    check = [0, n_eng // 2, n_eng - 1]
    err, cpu_s = 0.0, 0.0
    for i in check:
        o = O.oracle_engine()
        script(o, i, src, pads)
        t1 = time.perf_counter()
        want = o.bounce_to_buffer(bars)
        cpu_s += time.perf_counter() - t1
        o.close()
        err = max(err, float(np.abs(out[i][:len(want)] - want).max()))
    cpu_per_engine = cpu_s / len(check)
    res = {
        "workload": f"sample playback: {n_eng} FFI engines x {bars} bars, 2 loop channels each over one shared 10 s 48 kHz stereo loop (varispeed + rate conversion; "
                    f"Resample warp on a wrap-around window) + a 4-pad sampler rack, mono bounce; no WSOLA channel (serial per channel, see DESIGN.md 9.4)",
        "engines": n_eng, "frames": frames, "device_ms": dev_ms, "e2e_ms": wall * 1e3,
        "engine_samples_per_s": n_eng * frames / (dev_ms * 1e-3) if dev_ms > 0 else None,
        "e2e_engine_samples_per_s": n_eng * frames / wall,
        "note_device_ms": "whole bounce (idle drum voices, the two ext_source kernels, the mixers), CUDA events inside the library",
        "algorithmic_bytes_per_engine_sample": {"stored": 4 + 2 * 8, "gathered_from_l2": 2 * 32 + 4 * 8},
        "cpu_port_one_engine_s": cpu_per_engine, "cpu_port_engine_samples_per_s_per_core": frames / cpu_per_engine,
        "parity_max_err_vs_oracle": err, "parity_engines_checked": check, "engine_errors": int(sum(errs)),
    }
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
