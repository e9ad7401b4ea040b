"""Spread of the device-vs-oracle differences on config 2 (scDatEx, 500 genes x 100 permutations): quantiles of the
relative error of the jointPosterior terms and the list of discrete flips.  usage: python tools/parity_spread.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle as po            # noqa: E402
from scdd_b200 import hotpath as hp          # noqa: E402

d = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scDatEx.npz"), allow_pickle=True)
X, cond = d["normcounts"], d["condition"]
B = int(os.environ.get("PS_B", "100"))
pri = dict(alpha=0.01, mu0=0.0, s0=0.01, a0=0.01, b0=0.01)
logy, c01, off = hp.gene_csr(X, cond)
seeds = hp.gene_seeds(816, X.shape[0])
out, st, terms = hp.perm_bf(logy, c01, off, B, seed6=seeds, cfg=hp.make_cfg(**pri), return_terms=True)
ref = po.perm_batch_ex(logy, c01, off, B, seed6=seeds, cfg=po.make_cfg(**pri))
rel = np.abs(terms - ref["jp_terms"]) / np.abs(ref["jp_terms"])
print("lib", os.environ.get("SCDD_LIB_OVERRIDE", "default"))
print("quantiles 50/90/99/99.9/max:", [float(np.quantile(rel, q)) for q in (0.5, 0.9, 0.99, 0.999, 1.0)])
bad = np.argwhere(rel > 1e-8)
for g, b, w in bad[:20]:
    print("flip gene", g, "perm", b, "set", w, "n", int(off[g + 1] - off[g]), terms[g, b, w], ref["jp_terms"][g, b, w])
print("em_iters", st["em_iters"], "triples", st["triples"], "tol_near", st["tol_near"], "bic_near", st["bic_near"])
