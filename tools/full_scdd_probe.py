#!/usr/bin/env python3
"""Wall time of the whole scDD() call (host mirror) on the config-3 workload with a reduced number of permutations."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scdd_b200 import api
from scdd_b200.simulate import simulate_set_style
ng = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
X, cond = simulate_set_style(ng, 1000, seed=284)
sce = api.SingleCellExperiment({"normcounts": X}, {"condition": cond})
pri = dict(alpha=0.01, mu0=0, s0=0.01, a0=0.01, b0=0.01)
api.scDD(api.SingleCellExperiment({"normcounts": X[:64]}, {"condition": cond}), prior_param=pri, permutations=2, testZeroes=False, verbose=False)
t0 = time.time()
out = api.scDD(sce, prior_param=pri, permutations=B, testZeroes=False, param={"RNGseed": 1}, verbose=False)
dt = time.time() - t0
st = out.metadata["_scdd_b200"]["stats"]
print(f"scDD({ng} genes x 2000 cells, permutations={B}): {dt:.2f} s wall; observed fit kernel {st['observed']['fit_kernel_ms']:.0f} ms; "
      f"permutation fit kernel {st['permutations']['fit_kernel_ms']:.0f} ms, call total {st['permutations']['total_ms']:.0f} ms")
for k in ("observed", "permutations"):
    print(f"  {k}: {st[k]['fit_sets']:.0f} fit-sets, {st[k]['triples']:.3e} triples, "
          f"{st[k]['triples'] / (st[k]['fit_kernel_ms'] * 1e-3):.3e} triples/s, em_iters {st[k]['em_iters']:.0f}")
print("host-side share:", 1 - (st['observed']['fit_kernel_ms'] + st['permutations']['fit_kernel_ms']) / (dt * 1e3))
fd = st.get("feDP", {})
if fd:
    cats = out.metadata["Genes"]["DDcategory"]
    import collections
    print(f"feDP on the device: {fd['total_ms']:.1f} ms (H2D + statistics + {fd['fit_sets']:.0f} Fisher jobs + D2H), "
          f"tables too large: {fd['failed_sets']:.0f}; categories: {dict(collections.Counter(str(c) for c in cats))}")
