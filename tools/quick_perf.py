#!/usr/bin/env python3
"""Quick throughput probe of the resident plan (not the bench): synthetic log-normal-ish genes."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time

import numpy as np
import torch

from scdd_b200 import hotpath as hp
from scdd_b200 import _lib

ngenes = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ncells = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
nperm = int(sys.argv[3]) if len(sys.argv) > 3 else 16
rng = np.random.default_rng(0)
X = np.exp(rng.normal(3.0, 1.0, size=(ngenes, ncells)))
X[rng.random(X.shape) < 0.12] = 0.0
cond = np.repeat([1, 2], ncells // 2)
logy, c01, off = hp.gene_csr(X, cond)
seeds = hp.gene_seeds(1, ngenes)
import ctypes
clk = ctypes.c_double()
print("fp64 peak TFLOP/s:", _lib.lib().scdd_measure_fp64_peak(0, ctypes.byref(clk)), "clock MHz", clk.value)
plan = hp.Plan(0, logy, c01, off, nperm, seeds, cfg=hp.make_cfg(alpha=0.01))
s = torch.cuda.current_stream().cuda_stream
times = []
for it in range(int(os.environ.get('QP_CHUNKS', '3'))):
    torch.cuda.synchronize()
    t0 = time.time()
    plan.run_chunk(s, nperm)
    torch.cuda.synchronize()
    dt = time.time() - t0
    times.append(dt)
    print(f"chunk {it}: {dt*1e3:.1f} ms, {ngenes*nperm/dt:.0f} gene-perms/s")
st = plan.stats()
print(st)
T, P = st["triples"], st["point_iters"]
print("alg GFLOP/s (24T+4P):", (24 * T + 4 * P) / (st["fit_kernel_ms"] * 1e-3) / 1e9, "triples/s:", T / (st["fit_kernel_ms"] * 1e-3))

print(f"BEST chunk {min(times)*1e3:.2f} ms  median {sorted(times)[len(times)//2]*1e3:.2f} ms")
