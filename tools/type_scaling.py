#!/usr/bin/env python
"""How one voice type's render time scales with the number of voices (is the device saturated by 1024 warps?)."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch
from libgooey_b200 import voices as V, lib
from workloads import drum_sweep_patches

FRAMES = 88200
EXACT = os.environ.get("EXACT_TIER") == "1"
COUNTS = [int(x) for x in os.environ.get("COUNTS", "").split(",") if x]
patches, vel, kinds = drum_sweep_patches(max([16384] + [4 * c for c in COUNTS]), seed=0x600E7, exact_tier=EXACT)
L = lib()
only = sys.argv[1].split(",") if len(sys.argv) > 1 else ["tom", "hat", "snare", "kick"]
for kind, name in [(3, "tom"), (2, "hat"), (1, "snare"), (0, "kick")]:
    if name not in only:
        continue
    idx = [i for i, k in enumerate(kinds) if k == kind]
    for n in (COUNTS or ([1024, 2048, 4096] if len(sys.argv) > 1 else [256, 512, 1024, 2048, 4096])):
        sel = idx[:n]
        b = V.VoiceBatch([patches[i] for i in sel], 44100.0)
        v = np.ascontiguousarray(vel[sel])
        out = torch.empty((n, FRAMES), dtype=torch.float32, device="cuda:0")
        ms = []
        for r in range(3):
            b.trigger_all(0, v)
            b.render_device(FRAMES, out.data_ptr(), FRAMES)
            ms.append(L.gooey_b200_last_kernel_ms())
        b.close()
        print(name, n, "voices:", round(min(ms), 2), "ms  ->", round(n * FRAMES / min(ms) / 1e6, 1), "M voice-samples/ms... per-voice us:", round(min(ms) * 1e3 / n, 2), flush=True)
