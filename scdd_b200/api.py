"""Host-side mirror of the reference's R interface for the accelerated path.

R is not available in the build image, so the reference-facing layer is mirrored in Python with the SAME
names, argument meaning, error strings and result columns as the R functions it stands in for:

    scDD()      R/scDD.R:217-602     (dispatch only: filters, observed fits, permutations / KS, p-values,
                                      cluster counts, Zhat matrices, Genes table)
    testKS()    R/Test_KS.R:55-83
    results()   R/helper.R:62-66

All numerical work of the hot path happens in libscdd_b200.so (CUDA); this module contains no fitting code.
    feDP()      R/classify.dd.R:314-383  (the follow-up test of the non-significant genes)

    classifyDD() R/classify.dd.R:86-238  (the DE / DP / DM / DB / NC pattern of the significant genes;
                                      its Monte-Carlo test uses numpy's Student-t draws, not R's stream -- same
                                      distribution, borderline genes may differ as between two R seeds)

testZeroes (R/classify.dd.R:262-286, arm::bayesglm) runs on the device as well (csrc/zeroes.cuh): scDD(testZeroes=True)
adds the DZ category, the zero.pvalue columns and Fisher's combined p-value exactly as R/scDD.R:517-526, :568-588.
"""
import numpy as np

from . import hotpath as hp


class SingleCellExperiment:
    """Minimal stand-in for Bioconductor's SingleCellExperiment: assays, colData, metadata, dimnames."""

    def __init__(self, assays, colData, rownames=None, colnames=None):
        self.assays = {k: np.asarray(v, dtype=np.float64) for k, v in assays.items()}
        self.colData = {k: np.asarray(v) for k, v in colData.items()}
        first = next(iter(self.assays.values())) if self.assays else np.zeros((0, 0))
        self.rownames = list(rownames) if rownames is not None else [f"gene{i + 1}" for i in range(first.shape[0])]
        self.colnames = list(colnames) if colnames is not None else [f"Sample{i + 1}" for i in range(first.shape[1])]
        self.metadata = {}

    def assayNames(self):
        return list(self.assays.keys())


def normcounts(sce):
    return sce.assays["normcounts"]


def p_adjust_BH(p):
    """stats::p.adjust(p, method = "BH") with NA handling (NaN kept, n = number of non-NA)."""
    p = np.asarray(p, dtype=np.float64)
    out = np.full(p.shape, np.nan)
    ok = ~np.isnan(p)
    pv = p[ok]
    n = pv.size
    if n == 0:
        return out
    o = np.argsort(-pv, kind="stable")           # order(p, decreasing = TRUE)
    ro = np.argsort(o, kind="stable")
    adj = np.minimum(1.0, np.minimum.accumulate(n / np.arange(n, 0, -1) * pv[o]))
    out[ok] = adj[ro]
    return out


def _unique_in_order(x):
    seen, out = set(), []
    for v in x.tolist():
        if v not in seen:
            seen.add(v)
            out.append(v)
    return out


def luOutlier(x, min_size=3):
    """R/Cluster_actions.R:25  sum(table(x) >= min.size)"""
    _, cnt = np.unique(np.asarray(x), return_counts=True)
    return int((cnt >= min_size).sum())


def testKS(dat, condition, inclZero=True, numDE=None, DEIndex=None, rownames=None):
    """R/Test_KS.R:55-83.  dat: genes x cells matrix.  Returns dict(genes, p, p.unadj[, power, fdr])."""
    dat = np.asarray(dat, dtype=np.float64)
    condition = np.asarray(condition)
    lev = _unique_in_order(condition)
    keep = (condition == lev[0]) | (condition == lev[1])   # x1 / x2 use the first two levels only
    vals, c01, off = hp.gene_csr(dat[:, keep], condition[keep], ref=lev[0], log=False, incl_zero=bool(inclZero))
    _, p_unadj = hp.ks(vals, c01, off)
    p_adj = p_adjust_BH(p_unadj)
    sig = np.nonzero(p_adj < 0.05)[0]                      # hard-coded 0.05, R/Test_KS.R:71
    names = rownames if rownames is not None else [str(i + 1) for i in range(dat.shape[0])]
    out = {"genes": {names[i]: int(i + 1) for i in sig}, "p": p_adj, "p.unadj": p_unadj}
    if numDE is not None:
        de = set(int(i) for i in np.asarray(DEIndex).tolist())
        sig1 = [int(i + 1) for i in sig]
        out["power"] = sum(1 for i in sig1 if i in de) / numDE
        out["fdr"] = (sum(1 for i in sig1 if i not in de) / len(sig1)) if sig1 else float("nan")
    return out


def testZeroes(dat, cond, these=None):
    """R/classify.dd.R:262-286: p-value of a difference in the proportion of zeroes between conditions that the
    detection rate does not explain (Bayesian logistic regression, one device kernel for all genes).  dat: genes x
    cells on the original scale; these: 0-based rows (R's `these` is 1-based).  NaN = NA (gene without zeros)."""
    return hp.test_zeroes(dat, cond, these=these)[0]


testZeroes.__test__ = False


def fishersCombinedPval(x):
    """R/scDD.R:570-578, one row of cbind(pvals.all, pvals.z).  pchisq(-2 sum(log x), df = 4, lower = FALSE) has the
    closed form P (1 - log P), P the product of the two p-values."""
    x = np.asarray(x, dtype=np.float64)
    nas = int(np.isnan(x).sum())
    if nas == 0:
        with np.errstate(divide="ignore"):
            h = -float(np.sum(np.log(x)))              # chi-square statistic / 2
        if np.isinf(h):
            return 0.0
        k = len(x)                                     # upper tail of chi-square with 2k df: e^-h sum_{i<k} h^i / i!
        term, acc = 1.0, 0.0
        for i in range(k):
            acc += term
            term *= h / (i + 1)
        return float(np.exp(-h) * acc)
    if nas == 1:
        return float(x[~np.isnan(x)][0])
    return float("nan")


def feDP(logy, cond01, gene_off, sig_genes, cls_oa, cls_c, testZeroes=False, cdr=None, min_size=3, stats_out=None,
         precomputed=None):
    """R/classify.dd.R:314-383 over the CSR gene block (what `pe_mat[tofit, ]`, `condition` and the oa / c1 / c2
    class vectors are to the R function).  sig_genes: 0-based indices of the significant genes; cdr given means
    adjust.perms = TRUE.  Returns (ns_genes, pval_ns, cat): the BH-adjusted mean-shift p-values of the
    non-significant genes and their names "NS" / "DP" / "NC" -- R's `names(pval.ns) <- cat`.
    R quirk kept: with no significant gene `(1:nrow(pe_mat))[-sig_genes]` is EMPTY, so nothing is tested."""
    ng = len(gene_off) - 1
    thresh = 0.025 if testZeroes else 0.05
    sig = np.zeros(ng, dtype=bool)
    sig[np.asarray(sig_genes, dtype=np.int64)] = True
    ns = np.nonzero(~sig)[0] if sig.any() else np.zeros(0, dtype=np.int64)
    if ns.size == 0:
        return ns, np.zeros(0), np.zeros(0, dtype=object)
    p_shift, p_fisher, cnt, st = precomputed if precomputed is not None else \
        hp.fedp(logy, cond01, gene_off, cls_oa, cls_c, cdr=cdr, min_size=min_size)
    if stats_out is not None:
        stats_out.update(st)
    pv = p_adjust_BH(p_shift[ns])
    with np.errstate(invalid="ignore"):
        cat = np.where(p_fisher[ns] < 0.05, "DP", "NS").astype(object)       # :366-373 (NaN: test not run -> "NS")
        low = pv < thresh
    c = cnt[ns]
    cat[low & (cat != "DP")] = "NC"                                          # :378
    cat[low & (c[:, 1] == c[:, 2]) & (c[:, 1] == c[:, 0]) & (c[:, 1] > 1)] = "DP"   # :379
    return ns, pv, cat


def classifyDD(logy, cond01, gene_off, sig_genes, cls_oa, cls_c, prior_param, adjust_perms=False, cdr=None, min_size=3,
               seed=1, p_shift=None):
    """classifyDD, R/classify.dd.R:86-238, over the CSR gene block: the DE / DP / DM / DB / NC pattern of each gene in
    sig_genes (0-based) from the three clusterings -- one library call (scdd_classify): label bookkeeping and posterior
    parameters on host threads, the Monte-Carlo comparison of cluster means on the device.

    The reference decides whether two cluster means differ by 10,000 Student-t draws per cluster from R's own RNG
    stream (`rt`, :170-187), "different" iff all sampled differences share a sign.  The SAME test is made on the device
    from Philox streams keyed by (seed, comparison): the reference's distribution, not R's stream, so a gene whose
    clusters sit right at the threshold can be labelled differently -- exactly as between two R runs with different
    seeds.  p_shift: the unadjusted mean-shift p-values of scdd_fedp (Welch t / lm coefficient), used at :206-214;
    computed here when not given."""
    sig_genes = np.asarray(sig_genes, dtype=np.int64)
    if sig_genes.size == 0:
        return np.zeros(0, dtype=object)
    if p_shift is None:
        p_shift = hp.fedp(logy, cond01, gene_off, cls_oa, cls_c, cdr=cdr if adjust_perms else None, min_size=min_size)[0]
    cfg = hp.make_cfg(alpha=prior_param.get("alpha", 0.1), mu0=prior_param["mu0"], s0=prior_param["s0"],
                      a0=prior_param["a0"], b0=prior_param["b0"], min_size=min_size)
    return hp.classify(logy, cond01, gene_off, sig_genes, cls_oa, cls_c, p_shift=p_shift, cfg=cfg, seed=seed)[0]


def scDD(SCdat, prior_param=None, permutations=0, testZeroes=True, adjust_perms=False, param=None,
         parallelBy=("Genes", "Permutations"), condition="condition", min_size=3, min_nonzero=None, level=0.05,
         categorize=True, perm_source="device", verbose=True, classify=True):
    """R/scDD.R:217-602 with the per-gene work done on the GPU.

    param:       stands in for the BiocParallel param: dict(RNGseed=int[, devices=[...]]) -- RNGseed seeds the
                 per-gene L'Ecuyer-CMRG streams exactly as bplapply would.
    perm_source: "device" draws the permutations on the GPU (bit-exact R sampler); "host" draws them with the
                 library's host implementation of the same sampler and ships the indices (configs 1-2).
    classify:    True runs classifyDD (DE / DP / DM / DB / NC labels; Monte-Carlo on the device -- see classifyDD);
                 False leaves None for the genes the reference hands to classifyDD.
    """
    msg = (lambda s: print(s)) if verbose else (lambda s: None)
    import time as _time
    phases = {}
    _t = [_time.perf_counter()]

    def lap(name):                     # host wall time of the phase that just ended (seconds)
        now = _time.perf_counter()
        phases[name] = phases.get(name, 0.0) + now - _t[0]
        _t[0] = now
    if prior_param is None:
        prior_param = dict(alpha=0.10, mu0=0, s0=0.01, a0=0.01, b0=0.01)
    if not isinstance(SCdat, SingleCellExperiment):
        raise ValueError("Please provide a valid 'SingleCellExperiment' object.")
    if not SCdat.assayNames() or "normcounts" not in SCdat.assayNames():
        raise ValueError("Please make sure the 'SingleCellExperiment' object includes an assays slot named "
                         "'normcounts' that contains normalized counts on the original scale")
    if (not categorize) and permutations != 0:
        raise ValueError("Categorization will be carried out if permuations are run.")
    if isinstance(parallelBy, (tuple, list)):
        parallelBy = parallelBy[0]
    if parallelBy not in ("Genes", "Permutations"):
        raise ValueError("'arg' should be one of 'Genes', 'Permutations'")
    X = normcounts(SCdat)
    cond = SCdat.colData.get(condition)
    if cond is None or len(_unique_in_order(cond)) != 2 or len(cond) != X.shape[1]:
        raise ValueError("Error: Please specify valid condition labels.")
    ref = _unique_in_order(cond)[0]
    # one threaded pass over the matrix: nonzero cells per condition, constant-value flags, negatives
    nnz0, nnz1, const0, const1, n_negative = hp.dense_scan(X, cond, ref=ref)
    if n_negative > 0:
        raise ValueError("Error: Negative values for Normalized Expression counts detected. "
                         "Please ensure all counts are non-negative")
    if min_nonzero is None:
        min_nonzero = min_size
    isref = cond == ref
    need = max(min_size, 2, min_nonzero)
    tofit = np.nonzero((nnz0 >= need) & (nnz1 >= need))[0]
    ngenes = X.shape[0]
    todo = "Only testing for DZ for these genes." if testZeroes else "Skipping these genes."
    cut = level / 2 if testZeroes else level            # R/scDD.R:362-366, :475-479, :503-507
    if tofit.size < ngenes:
        msg(f"Notice: {ngenes - tofit.size} genes have less than {min_nonzero} nonzero cells per condition.  {todo}")
    # genes whose nonzero values are all identical within a condition (R/scDD.R:292-310)
    skip = const0[tofit] | const1[tofit]
    if skip.any():
        msg(f"Notice: {int(skip.sum())} Genes have constant nonzero values.  {todo}")
        tofit = tofit[~skip]
    param = dict(param or {})
    if param.get("devices") is not None:
        hp.set_devices(param["devices"])
    msg(f"Setting up parallel back-end using {hp.device_count()} GPUs")
    cfg = hp.make_cfg(alpha=prior_param["alpha"], mu0=prior_param["mu0"], s0=prior_param["s0"], a0=prior_param["a0"],
                      b0=prior_param["b0"], min_size=min_size)
    nf = tofit.size
    sel = None if nf == ngenes else tofit
    lap("filters")
    logy, c01, off = hp.gene_csr(X, cond, ref=ref, genes=sel, nnz=(nnz0 + nnz1)[tofit])
    lap("gene_csr")
    stats = {}
    if categorize:
        msg("Clustering observed expression data for each gene")
    fits = cls_oa = cls_c = None
    bf = den = None
    comps_all = comps_c1 = comps_c2 = np.full(nf, np.nan)
    if categorize:
        fits, cls_oa, cls_c, bf, den, st = hp.fit_observed(logy, c01, off, cfg=cfg, with_bf=permutations != 0)
        stats["observed"] = st
        # luOutlier (R/Cluster_actions.R:25) for every gene at once
        comps_all = hp.cluster_counts(cls_oa, c01, off, min_size=min_size).astype(np.float64)
        comps_c1 = hp.cluster_counts(cls_c, c01, off, min_size=min_size, which=0).astype(np.float64)
        comps_c2 = hp.cluster_counts(cls_c, c01, off, min_size=min_size, which=1).astype(np.float64)
    cdr = None
    if adjust_perms:
        # per-cell detection rate over the fitted genes (R/scDD.R:457, R/classify.dd.R:331), spread over the nonzero cells
        C = ((X if sel is None else X[tofit]) > 0).sum(axis=0) / float(nf)
        cdr = hp.cell_covariate(X, C, off, genes=sel)
    lap("observed_fits")
    if permutations == 0:
        msg("Notice: Number of permutations is set to zero; using Kolmogorov-Smirnov to test for differences "
            "in distributions instead of the Bayes Factor permutation test")
        res_ks = testKS(X if sel is None else X[tofit], cond, inclZero=False)
        sig = np.nonzero(res_ks["p"] < cut)[0]
        pvals = res_ks["p.unadj"]
    else:
        msg("Performing permutations to evaluate independence of clustering and condition for each gene")
        msg(f"Parallelizing by {parallelBy}")
        seed = int(param.get("RNGseed", 1))
        # RNG-stream shape of the two dispatch modes: "Genes" = one L'Ecuyer-CMRG stream per gene
        # (bplapply over genes, R/scDD.R:458); "Permutations" = permutation b of EVERY gene from stream b, because
        # scDD() loops over the genes and each permMclust / permMclustCov call runs bplapply(1:nperms) under the same
        # RNGseed (R/scDD.R:414-455, R/permutations.R:177, :99)
        by_perm = parallelBy == "Permutations"
        seeds = hp.gene_seeds(seed, permutations if by_perm else nf)
        if perm_source == "host":
            def draw(g):
                n = int(off[g + 1] - off[g])
                if by_perm:
                    return np.stack([hp.sample_perms(seeds[b], n, 1)[0][0] for b in range(permutations)]).ravel()
                return hp.sample_perms(seeds[g], n, permutations)[0].ravel()
            perm_idx = np.concatenate([draw(g) for g in range(nf)]) if nf else np.zeros(0, dtype=np.int32)
            bf_perm, st = hp.perm_bf(logy, c01, off, permutations, perm_idx=perm_idx, cdr=cdr, cfg=cfg, allow_failed=True)
        else:
            bf_perm, st = hp.perm_bf(logy, c01, off, permutations, seed6=seeds, cdr=cdr, cfg=cfg, allow_failed=True,
                                     seed_mode=int(by_perm), interrupt=param.get("interrupt"))
        if st["const_sets"] > 0:
            msg(f"Notice: {int(st['const_sets'])} permuted subsets had identical values and were jittered "
                f"(runif(-0.1, 0.1), as mclustRestricted does)")
        if st["failed_sets"] > 0:
            msg(f"Warning: {int(st['failed_sets'])} permuted fit-sets had no valid model; the p-values of their genes are NA")
        stats["permutations"] = st
        obs = (bf - den) if adjust_perms else bf                      # R/scDD.R:467-473
        pvals = (bf_perm > obs[:, None]).sum(axis=1) / float(permutations)
        pvals = np.where(np.isnan(bf_perm).any(axis=1) | np.isnan(obs), np.nan, pvals)   # R: sum() over an NA is NA
        with np.errstate(invalid="ignore"):
            sig = np.nonzero(p_adjust_BH(pvals) < cut)[0]
    lap("test")
    pvals_all = np.full(ngenes, np.nan)
    pvals_all[tofit] = pvals
    cats_all = [None] * ngenes
    if categorize:
        msg("Classifying significant genes into patterns")
        # R/scDD.R:485-512: significant genes go to classifyDD (R; None here), the others start as "NS" and feDP
        # promotes some to "DP" or to classifyDD candidates ("NC" with adjusted p > level -> None here)
        cats = np.full(nf, "NS", dtype=object)
        # the unadjusted mean-shift p-values of every gene (one kernel): feDP needs them for the non-significant genes,
        # classifyDD for its multimodal-DE rule
        fe_all = hp.fedp(logy, c01, off, cls_oa, cls_c, cdr=cdr if adjust_perms else None, min_size=min_size) \
            if nf else (np.zeros(0), np.zeros(0), np.zeros((0, 3), dtype=np.int32), {})
        p_shift_all = fe_all[0]
        lap("feDP_kernels")
        cseed = int(param.get("classify_seed", param.get("RNGseed", 1)))
        if classify:
            cats[sig] = classifyDD(logy, c01, off, sig, cls_oa, cls_c, prior_param, adjust_perms=adjust_perms,
                                   cdr=cdr if adjust_perms else None, min_size=min_size, seed=cseed, p_shift=p_shift_all)
        else:
            cats[sig] = None
        ns, _, cat_ns = feDP(logy, c01, off, sig, cls_oa, cls_c, testZeroes=testZeroes, cdr=cdr if adjust_perms else None,
                             min_size=min_size, stats_out=stats.setdefault("feDP", {}), precomputed=fe_all)
        cats[ns] = cat_ns
        with np.errstate(invalid="ignore"):
            to_classify = np.nonzero((p_adjust_BH(pvals) > cut) & (cats == "NC"))[0]
        if classify:
            cats[to_classify] = classifyDD(logy, c01, off, to_classify, cls_oa, cls_c, prior_param,
                                           adjust_perms=adjust_perms, cdr=cdr if adjust_perms else None,
                                           min_size=min_size, seed=cseed + 1, p_shift=p_shift_all)
        else:
            cats[to_classify] = None
        for j, g in enumerate(tofit):
            cats_all[g] = cats[j]
        lap("classify")
    # zero test (R/scDD.R:517-526): DZ for the genes the nonzero test did not categorise, NS for the rest
    pvals_z = np.full(ngenes, np.nan)
    if testZeroes:
        pvals_z, stz = hp.test_zeroes(X, cond)
        stats["testZeroes"] = stz
        pz = p_adjust_BH(pvals_z)
        for g in range(ngenes):
            if cats_all[g] in ("DE", "DP", "DM", "DB") or np.isnan(pz[g]):
                continue
            if pz[g] < level / 2:
                cats_all[g] = "DZ"
            elif categorize:
                cats_all[g] = "NS"
    lap("testZeroes")
    # Zhat matrices (R/scDD.R:528-552): 0 = zero count, 1 = default for untested nonzero, else cluster label
    if categorize:
        MAP = hp.labels_to_dense(X, cond, cls_oa, off, genes=tofit, ref=ref)
        MAP1 = hp.labels_to_dense(X, cond, cls_c, off, genes=tofit, which=0, ref=ref)
        MAP2 = hp.labels_to_dense(X, cond, cls_c, off, genes=tofit, which=1, ref=ref)
    else:
        MAP1 = MAP2 = MAP = "Categorize set to FALSE; no clustering provided."

    lap("Zhat")

    def _spread(v):
        out = np.full(ngenes, np.nan)
        out[tofit] = v
        return out
    Genes = {"gene": list(SCdat.rownames), "DDcategory": cats_all,
             "Clusters.combined": _spread(comps_all), "Clusters.c1": _spread(comps_c1),
             "Clusters.c2": _spread(comps_c2), "nonzero.pvalue": pvals_all,
             "nonzero.pvalue.adj": p_adjust_BH(pvals_all)}
    if testZeroes:                                      # R/scDD.R:568-588
        comb = np.array([fishersCombinedPval((a, b)) for a, b in zip(pvals_all, pvals_z)])
        Genes.update({"zero.pvalue": pvals_z, "zero.pvalue.adj": p_adjust_BH(pvals_z),
                      "combined.pvalue": comb, "combined.pvalue.adj": p_adjust_BH(comb)})
    SCdat.metadata["Genes"] = Genes
    SCdat.metadata["Zhat.combined"] = MAP
    SCdat.metadata["Zhat.c1"] = MAP1
    SCdat.metadata["Zhat.c2"] = MAP2
    lap("assemble")
    stats["host_phases_s"] = phases
    SCdat.metadata["_scdd_b200"] = {"stats": stats, "tofit": tofit, "fits": fits, "bf": bf, "den": den}
    return SCdat


def results(SCdat, type="Genes"):
    """R/helper.R:62-66"""
    if type not in ("Genes", "Zhat.c1", "Zhat.c2", "Zhat.combined"):
        raise ValueError("'arg' should be one of 'Genes', 'Zhat.c1', 'Zhat.c2', 'Zhat.combined'")
    return SCdat.metadata[type]
