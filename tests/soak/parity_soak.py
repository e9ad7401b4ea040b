#!/usr/bin/env python3
"""Parity soak: GPU path vs CPU oracle on a large random sample of the benchmark workload.
Reports how many fit-sets / Bayes factors differ beyond the tolerances of BASELINE.json and the near-tie counters.
usage: parity_soak.py [genes] [perms]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle import pyoracle as po
from scdd_b200 import hotpath as hp
from scdd_b200.simulate import simulate_set_style

ng = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
PRI = dict(alpha=0.01, mu0=0.0, s0=0.01, a0=0.01, b0=0.01)
X, cond, cats = simulate_set_style(20000, 1000, seed=284, return_categories=True)
rng = np.random.default_rng(99)
rows = np.sort(rng.choice(20000, size=ng, replace=False))
logy, c01, off = hp.gene_csr(X[rows], cond)
print(f"sample: {ng} genes of the 20000 x 2000 workload ({dict(zip(*np.unique(cats[rows], return_counts=True)))}), {B} permutations")

t0 = time.time()
fits, cls_oa, cls_c, bf, den, st = hp.fit_observed(logy, c01, off, cfg=hp.make_cfg(**PRI))
t1 = time.time()
ofits, ocls_oa, ocls_c, obf, oden = po.genefit_batch(logy, c01, off, cfg=po.make_cfg(**PRI))
t2 = time.time()
gm = sum(int(fits[g, w]["G"] != ofits[g][w]["G"]) for g in range(ng) for w in range(3))
bic_rel = max(abs(fits[g, w]["bic"] - ofits[g][w]["bic"]) / abs(ofits[g][w]["bic"]) for g in range(ng) for w in range(3))
print(f"observed fits: {3*ng} fit-sets, GPU {t1-t0:.2f} s, oracle {t2-t1:.1f} s")
print(f"  selected-G mismatches: {gm};  label mismatches (cells): oa {int((cls_oa != ocls_oa).sum())}, c1/c2 {int((cls_c != ocls_c).sum())} of {cls_oa.size}")
print(f"  max relative difference: BIC {bic_rel:.2e}, bf {np.max(np.abs(bf-obf)/np.abs(obf)):.2e}, den {np.max(np.abs(den-oden)/np.abs(oden)):.2e}")
print(f"  kernel near-tie counters: tol_near {st['tol_near']:.0f}, bic_near {st['bic_near']:.0f}, iter_cap_hits {st['iter_cap_hits']:.0f}, failed {st['failed_sets']:.0f}")

seeds = hp.gene_seeds(7, ng)
t0 = time.time()
out, st = hp.perm_bf(logy, c01, off, B, seed6=seeds, cfg=hp.make_cfg(**PRI))
t1 = time.time()
ref = po.perm_batch(logy, c01, off, B, seed6=seeds, cfg=po.make_cfg(**PRI))
t2 = time.time()
rel = np.abs(out - ref) / np.abs(ref)
print(f"permutations: {ng*B} gene-permutations ({2*ng*B} fit-sets), GPU {t1-t0:.2f} s, oracle {t2-t1:.1f} s")
print(f"  |rel diff| > 1e-8: {int((rel > 1e-8).sum())};  max {rel.max():.2e};  median {np.median(rel):.2e}")
print(f"  p-value counts identical: {bool(np.array_equal((out > bf[:, None]).sum(1), (ref > obf[:, None]).sum(1)))}")
print(f"  kernel near-tie counters: tol_near {st['tol_near']:.0f}, bic_near {st['bic_near']:.0f}, iter_cap_hits {st['iter_cap_hits']:.0f}, failed {st['failed_sets']:.0f}")
