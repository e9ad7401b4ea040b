"""debug helper: the permuted fit-set (gene 174, permutation 86 of config 2) through scdd_fit_observed, next to the oracle"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle as po
from scdd_b200 import hotpath as hp
d = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scDatEx.npz"), allow_pickle=True)
X, cond = d["normcounts"], d["condition"]
logy, c01, off = hp.gene_csr(X, cond)
seeds = hp.gene_seeds(816, X.shape[0])
g, b = int(os.environ.get("DBG_G", 174)), int(os.environ.get("DBG_B", 86))
n = int(off[g + 1] - off[g]); y = logy[off[g]:off[g + 1]]; c = c01[off[g]:off[g + 1]]
p = hp.sample_perms(seeds[g], n, b + 1)[0][b] - 1
yp = np.ascontiguousarray(y[p])
pri = dict(alpha=0.01, mu0=0.0, s0=0.01, a0=0.01, b0=0.01)
fits, cls_oa, cls_c, bf, den, st = hp.fit_observed(yp, c, np.array([0, n], dtype=np.int64), cfg=hp.make_cfg(**pri), allow_failed=True)
print("lib", os.environ.get("SCDD_LIB_OVERRIDE", "default"), "em_iters", st["em_iters"], "bf", bf, "den", den)
for w, name in enumerate(["oa", "c1", "c2"]):
    f = fits[0, w]
    print(name, "G", f["G"], "bic", f["bic"], "loglik", f["loglik"], "mean", f["mean"][:f["G"]], "var", f["var"][:f["G"]])
ocfg = po.make_cfg(**pri)
for name, v in (("c1", yp[c == 0]), ("c2", yp[c != 0])):
    rows = po.fit_table(v, ocfg)
    print("oracle", name, [(r["G"], r["valid"], r.get("iters"), round(r.get("bic", 0), 6)) for r in rows])
    f = po.mclust_restricted(v, cfg=ocfg)
    print("oracle", name, "bic", f["bic"], "mean", f["mean"], "var", f["var"])
