import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from libgooey_b200 import lib
L = lib()
for rep in range(2):
    r = bench.config_c4(L)
    print("C4 alone:", r["device_ms"], r["e2e_ms"], flush=True)
import torch
import engine_scripts as S
r = bench.engine_config(L, torch, 0, list(range(256)), bench._c3_script(S), 8, {0}, 6542.1, None, "c3-small")
print("C3 small:", r["device_ms"], r["device_ms_settled"], r["e2e_ms"], flush=True)
r = bench.config_c4(L)
print("C4 after C3:", r["device_ms"], r["e2e_ms"], flush=True)
