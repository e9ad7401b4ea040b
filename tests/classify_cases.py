"""Clusterings whose classifyDD pattern is unambiguous (shared by the oracle's CPU test and the device test)"""
import numpy as np


def clear_cases(seed=3):
    rng = np.random.default_rng(seed)

    def gene(c1_spec, c2_spec, oa_k):
        """spec: list of (mean, count); returns y, cond01, class vectors with labels ordered by mean"""
        ys, cs, l1, l2 = [], [], [], []
        for k, (m, cnt) in enumerate(c1_spec):
            ys.append(rng.normal(m, 0.1, cnt)); cs.append(np.zeros(cnt, dtype=np.int32)); l1.append(np.full(cnt, k + 1))
        for k, (m, cnt) in enumerate(c2_spec):
            ys.append(rng.normal(m, 0.1, cnt)); cs.append(np.ones(cnt, dtype=np.int32)); l2.append(np.full(cnt, k + 1))
        y, c = np.concatenate(ys), np.concatenate(cs)
        cls_c = np.concatenate(l1 + l2).astype(np.int32)
        # overall clustering: oa_k clusters by thresholding between the distinct means
        means = sorted(set(m for m, _ in c1_spec + c2_spec))
        cuts = [(means[i] + means[i + 1]) / 2 for i in range(len(means) - 1)][:oa_k - 1]
        cls_oa = (1 + sum((y > t).astype(np.int32) for t in cuts)).astype(np.int32) if cuts else np.ones(y.size, dtype=np.int32)
        return y, c, cls_oa, cls_c

    cases = [("DE", gene([(1.0, 50)], [(3.0, 50)], 2)),
             ("NC", gene([(2.0, 50)], [(2.0, 50)], 1)),
             ("DB", gene([(1.0, 30), (5.0, 30)], [(3.0, 60)], 3)),
             ("DM", gene([(1.0, 30), (5.0, 30)], [(1.0, 60)], 2)),
             ("DP", gene([(1.0, 18), (5.0, 42)], [(1.0, 42), (5.0, 18)], 2)),
             ("DE", gene([(1.0, 30), (3.0, 30)], [(5.0, 30), (7.0, 30)], 4))]       # multimodal DE: 2 + 2 clusters, 4 overall
    ys, cs, oas, ccs = zip(*[g for _, g in cases])
    off = np.r_[0, np.cumsum([y.size for y in ys])].astype(np.int64)
    return ([w for w, _ in cases], np.concatenate(ys), np.concatenate(cs), off, np.concatenate(oas).astype(np.int32),
            np.concatenate(ccs).astype(np.int32))
