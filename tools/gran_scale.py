#!/usr/bin/env python
"""BASELINE.json config C4 at scale: N granulators (64 + 16 grain slots each, density 80 grains/s, ~0.9 s grains -> the
pool saturates) over ONE shared synthetic 60 s source buffer, rendered `seconds` s in one batch.

    python tools/gran_scale.py --engines 1600 --seconds 10 [--check 2]
"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--engines", type=int, default=1600)
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--check", type=int, default=2)
    args = ap.parse_args()
    import torch
    import __graft_entry__ as g
    g.build()
    from libgooey_b200 import engine as G, lib
    import oracle_lib as O
    n = np.arange(2646000, dtype=np.float64)
    rng = np.random.default_rng(0x600E7)
    src = (0.5 * np.sin(2 * np.pi * 220.0 * n / 44100.0) * (0.5 + 0.5 * np.sin(2 * np.pi * 0.1 * n / 44100.0)) + 0.1 * rng.uniform(-1, 1, len(n))).astype(np.float32)
    pitch = rng.uniform(0.3, 0.7, args.engines); tex = rng.random(args.engines)

    def script(e, i, first=None):
        assert (e.granulator_set_buffer(src, 44100.0) if first is None else e.granulator_share_buffer(first))
        for p, v in [(4, 1.0), (1, 0.55), (2, 0.5), (3, float(pitch[i])), (6, 0.3), (5, float(tex[i])), (9, 0.3), (10, 0.3), (7, 1.0)]:
            e.granulator_set_param(p, v)
        e.granulator_set_seed(i + 1)
        e.granulator_snap_params()
        e.granulator_trigger(1.0)
    engines = [G.Engine() for _ in range(args.engines)]
    for i, e in enumerate(engines):
        script(e, i, None if i == 0 else engines[0])
    frames = int(args.seconds * 44100)
    L = lib()
    import ctypes
    L.gooey_batch_render.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p]
    hs = (ctypes.c_void_p * args.engines)(*[e._h for e in engines])
    out = np.zeros((args.engines, frames, 2), np.float32)
    t0 = time.perf_counter()
    rc = L.gooey_batch_render(hs, args.engines, frames, out.ctypes.data)
    dt = time.perf_counter() - t0
    assert rc == 0, L.gooey_b200_last_error()
    res = {"engines": args.engines, "frames": frames, "wall_s": round(dt, 3), "kernel_ms": L.gooey_b200_last_kernel_ms(),
           "engine_samples_per_s": args.engines * frames / (L.gooey_b200_last_kernel_ms() * 1e-3),
           "grain_slot_samples_per_s": 80 * args.engines * frames / (L.gooey_b200_last_kernel_ms() * 1e-3), "peak": float(np.abs(out).max())}
    errs = []
    for i in np.linspace(0, args.engines - 1, args.check).astype(int):
        o = O.oracle_engine(); script(o, int(i)); want = o.render(frames); o.close()
        errs.append(float(np.abs(out[i] - want).max()))
    res["max_abs_err_vs_oracle"] = errs
    print(json.dumps(res))


if __name__ == "__main__":
    main()
