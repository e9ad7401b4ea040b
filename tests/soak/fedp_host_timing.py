#!/usr/bin/env python3
"""Times the feDP restatement (oracle/fedp_oracle.py: numpy/scipy, one gene at a time like the R loop) on a sample of
the config-3 workload, next to the device time tools/full_scdd_probe.py prints.  Test infrastructure (uses oracle/)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle import fedp_oracle as fo
from scdd_b200 import hotpath as hp
from scdd_b200.simulate import simulate_set_style
ng = int(sys.argv[1]) if len(sys.argv) > 1 else 200
X, cond = simulate_set_style(ng, 1000, seed=284)
logy, c01, off = hp.gene_csr(X, cond)
fits, cls_oa, cls_c, _, _, _ = hp.fit_observed(logy, c01, off, cfg=hp.make_cfg(alpha=0.01))
C = (X > 0).sum(axis=0) / float(ng)
t0 = time.time()
for j in range(ng):
    s = slice(off[j], off[j + 1])
    fo.fedp_gene(X[j], cond, C, cls_oa[s], cls_c[s][c01[s] == 0], cls_c[s][c01[s] == 1])
dt = time.time() - t0
hp.fedp(logy, c01, off, cls_oa, cls_c)
t0 = time.time()
st = hp.fedp(logy, c01, off, cls_oa, cls_c)[3]
print(f"feDP restatement on the host: {dt / ng * 1e3:.2f} ms/gene on {ng} genes (x 20000 = {dt / ng * 20000:.1f} s); "
      f"scdd_fedp on the same {ng} genes: {(time.time() - t0) * 1e3:.1f} ms wall, {st['total_ms']:.1f} ms device")
