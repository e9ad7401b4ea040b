#!/usr/bin/env python3
"""Scaled-down probe of BASELINE config 5: genes x 50,000 cells, adjust.perms=TRUE (CDR covariate, 3 fit-sets per
permutation, CTA-per-fit-set team with an L2-resident slab).  usage: cfg5_probe.py [genes] [cells] [perms]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scdd_b200 import hotpath as hp
from scdd_b200.simulate import simulate_set_style
ng = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cells = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
B = int(sys.argv[3]) if len(sys.argv) > 3 else 4
X, cond = simulate_set_style(ng, cells // 2, seed=284)
logy, c01, off = hp.gene_csr(X, cond)
C = (X > 0).sum(axis=0) / float(ng)
cdr = np.broadcast_to(C, X.shape)[X > 0]
seeds = hp.gene_seeds(1, ng)
cfg = hp.make_cfg(alpha=0.01)
for adj in (False, True):
    hp.perm_bf(logy, c01, off, 1, seed6=seeds, cdr=cdr if adj else None, cfg=cfg)       # warm-up
    t0 = time.time()
    out, st = hp.perm_bf(logy, c01, off, B, seed6=seeds, cdr=cdr if adj else None, cfg=cfg)
    dt = time.time() - t0
    print(f"adjust.perms={adj}: {ng} genes x {cells} cells x {B} perms: {dt:.2f} s, {ng*B/dt:.1f} gene-perms/s, "
          f"fit kernel {st['fit_kernel_ms']:.0f} ms, rng {st['rng_kernel_ms']:.0f} ms, triples/s {st['triples']/(st['fit_kernel_ms']*1e-3):.3e}, "
          f"fit_sets {st['fit_sets']:.0f}, failed {st['failed_sets']:.0f}")
