#!/usr/bin/env python3
"""Where does the observed-fit call (3 fit-sets per gene, one launch) spend its time?  Kernel time against the number
of genes: a flat curve means the launch is bound by its slowest fit-sets, not by throughput."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scdd_b200 import hotpath as hp
from scdd_b200.simulate import simulate_set_style
X, cond = simulate_set_style(20000, 1000, seed=284)
cfg = hp.make_cfg(alpha=0.01)
hp.fit_observed(*hp.gene_csr(X[:64], cond), cfg=cfg)
for ng in (500, 2000, 5000, 20000):
    logy, c01, off = hp.gene_csr(X[:ng], cond)
    fits, _, _, _, _, st = hp.fit_observed(logy, c01, off, cfg=cfg)
    it = np.array([[f["n"] for f in row] for row in fits])
    print(f"{ng:6d} genes: fit kernel {st['fit_kernel_ms']:7.1f} ms, {st['triples']:.3e} triples, {st['triples'] / st['fit_kernel_ms'] / 1e-3:.3e} triples/s, "
          f"em_iters {st['em_iters']:.0f}, max n {it.max()}")
