#!/usr/bin/env python3
"""Small, short-running drivers for ncu captures of single kernels (used under gpurun).
  python tools/prof_kernels.py merkle [log_h] [width]   # linearhash leaves + merkle levels on random columns
  python tools/prof_kernels.py lde [log_n] [width]      # coset LDE (blowup 2) on random columns
  python tools/prof_kernels.py ntt [log_n] [width]
  python tools/prof_kernels.py msm [log_n]              # BN254 G1 MSM on deterministic random points
"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from eigen_zkvm_b200 import _lib
L = _lib.lib()
what = sys.argv[1]
a1 = int(sys.argv[2]) if len(sys.argv) > 2 else 22
w = int(sys.argv[3]) if len(sys.argv) > 3 else 2
P = 0xFFFFFFFF00000001
def rnd(n):
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    x = torch.randint(0, 2**62, (n,), dtype=torch.int64, device="cuda", generator=g)
    return x  # < 2^62 < p: canonical
reps = int(os.environ.get("REPS", "3"))
L.b200_timing_enable(1)
if what == "merkle":
    h = 1 << a1
    leaves = rnd(h * w)
    nodes = torch.empty(L.b200_gl_merkle_n_nodes(h) * 4, dtype=torch.int64, device="cuda")
    for _ in range(reps):
        _lib.check(L.b200_gl_merkelize_dev(ctypes.c_void_p(leaves.data_ptr()), w, h, ctypes.c_void_p(nodes.data_ptr())))
elif what == "lde":
    n = 1 << a1
    src = rnd(n * w); dst = torch.empty(2 * n * w, dtype=torch.int64, device="cuda")
    for _ in range(reps):
        _lib.check(L.b200_gl_lde_dev(ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(dst.data_ptr()), w, a1, a1 + 1))
elif what == "ntt":
    n = 1 << a1
    src = rnd(n * w); dst = torch.empty(n * w, dtype=torch.int64, device="cuda")
    for _ in range(reps):
        _lib.check(L.b200_gl_ntt_dev(ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(dst.data_ptr()), w, a1, 0))
elif what == "msm":
    import numpy as np
    from eigen_zkvm_b200 import groth16 as g16
    n = 1 << a1
    d_b = torch.empty(n * 8, dtype=torch.int64, device="cuda")
    g16.random_points_dev(d_b.data_ptr(), n, 0xB254)
    d_s = rnd(n * 4)
    tab = g16.MsmTable(device_ptr=d_b.data_ptr(), n=n)          # the per-circuit table mode (bench.py's headline MSM)
    print("table: window %d bits, %d windows" % (tab.window_bits, tab.windows))
    for _ in range(reps):
        tab.run_dev(d_s.data_ptr())
torch.cuda.synchronize()
from eigen_zkvm_b200 import starky
for r in starky.timing_report():
    per = r["ms"] / r["launches"]
    print("%-20s launches %4d  total %9.3f ms  per-launch %8.4f ms  %8.1f GB/s" % (r["name"], r["launches"], r["ms"], per, r["bytes"] / r["launches"] / per / 1e6))
print("done", what)
