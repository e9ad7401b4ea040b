# scDD_gpu.R -- R-side helpers a maintainer adds to the reference package: each replaces one bplapply fan-out (or one
# per-gene R loop) of the hot path with ONE .Call into libscdd_b200.so.  r/R/scDD_b200.R is the dispatcher built on
# them: scDD() with the reference's exact signature.  NAMESPACE gains  useDynLib(scDD, .registration = TRUE).
#
# The glue (r/src/scdd_glue.c) is type-strict on purpose -- INTEGER() on a double vector is undefined behaviour in R's
# C API -- so every helper coerces with as.integer() / as.double() before the call.

.cond01 <- function(condition, ref) as.integer(condition != ref)

.gene_csr <- function(dat, condition, ref, log = TRUE, inclZero = FALSE) {
  # dat: genes x cells; returns the CSR-over-genes layout of include/scdd_b200.h (vals, cond01, off) -- the
  # library walks R's column-major matrix directly, one threaded pass to count and one to gather
  storage.mode(dat) <- "double"
  .Call(C_scdd_gene_csr, dat, nrow(dat), ncol(dat), .cond01(condition, ref), log, inclZero)
}

# the gene filters of R/scDD.R:262-298 in one pass: nonzero cells per condition, constant-value flags, negatives
dense_scan_gpu <- function(dat, condition, ref) {
  storage.mode(dat) <- "double"
  .Call(C_scdd_dense_scan, dat, nrow(dat), ncol(dat), .cond01(condition, ref))
}

# a per-cell covariate (the detection rate C, R/scDD.R:457) on the nonzero cells of every gene, in CSR order
cell_covariate_gpu <- function(dat, values, off) {
  storage.mode(dat) <- "double"
  .Call(C_scdd_cell_covariate, dat, nrow(dat), ncol(dat), as.double(values), as.double(off))
}

# replaces the MAP/MAP1/MAP2 loop of R/scDD.R:528-552; which = -1 (Zhat.combined), 0 (Zhat.c1), 1 (Zhat.c2);
# labels = the flat oa class vector (combined) or the c1/c2 classes in cell order ("class.c")
zhat_gpu <- function(dat, condition, ref, tofit, off, labels, which = -1L) {
  storage.mode(dat) <- "double"
  .Call(C_scdd_zhat, dat, nrow(dat), ncol(dat), .cond01(condition, ref), as.integer(which),
        as.integer(tofit), as.double(off), as.integer(labels))
}

# replaces  comps.all / comps.c1 / comps.c2 <- luOutlier(...)  of R/scDD.R:335-337 for all genes at once
luOutlier_gpu <- function(labels, cond01, off, min.size, which = -1L)
  .Call(C_scdd_cluster_counts, as.integer(labels), as.integer(cond01), as.double(off), as.integer(min.size),
        as.integer(which))

# replaces:  out <- bplapply(1:nrow(...), function(x) genefit(normcounts(SCdat)[tofit[x],]))   R/scDD.R:343, :393
# Returns the oa / c1 / c2 lists exactly as mclustRestricted returns them (class, mean, var, df, n, bic, loglik, model;
# R/mclust.restricted.R:249-252) -- classifyDD() consumes them unchanged -- plus bf, den and the flat vectors.
genefit_gpu <- function(dat, condition, ref, prior, min.size, with.bf = TRUE) {
  g <- .gene_csr(dat, condition, ref)
  r <- .Call(C_scdd_fit_observed, g$vals, g$off, g$cond01, as.double(prior), as.integer(min.size), with.bf)
  ng <- nrow(dat)
  # the per-gene lists are built with split() / Map(): 3 x ngenes small lists without an R-level loop over cells
  gene <- rep.int(seq_len(ng), diff(g$off))
  is1 <- g$cond01 == 0L
  cls.oa <- split(r$class.oa, factor(gene, levels = seq_len(ng)))
  cls.c1 <- split(r$class.c[is1], factor(gene[is1], levels = seq_len(ng)))
  cls.c2 <- split(r$class.c[!is1], factor(gene[!is1], levels = seq_len(ng)))
  mk <- function(i, cls) {
    G <- r$G[i]
    k <- seq_len(G)
    list(class = cls, mean = setNames(r$mean[k, i], k), var = r$var[k, i],
         df = if (G > 1) 3 * G - 1 else 2, n = r$n[i], bic = r$bic[i], loglik = r$loglik[i],
         model = if (G > 1) "V" else "X")
  }
  j <- seq_len(ng)
  list(oa = unname(Map(mk, 3 * j - 2, cls.oa)), c1 = unname(Map(mk, 3 * j - 1, cls.c1)),
       c2 = unname(Map(mk, 3 * j, cls.c2)), bf = r$bf, den = r$den, raw = r, csr = g)
}

# The L'Ecuyer-CMRG state BiocParallel hands to element i of bplapply(seq_len(n), ..., BPPARAM = param) for
# RNGseed = seed: set.seed(seed, "L'Ecuyer-CMRG") advanced by i - 1 substreams.  6 x n integer matrix.
# parallelBy = "Genes": n = number of genes (R/scDD.R:458); "Permutations": n = permutations (R/permutations.R:99,177).
rng_streams <- function(RNGseed, n) {
  old <- RNGkind("L'Ecuyer-CMRG"); on.exit(do.call(RNGkind, as.list(old)))
  set.seed(RNGseed)
  s <- .Random.seed
  out <- matrix(0L, 6, n)
  for (i in seq_len(n)) { out[, i] <- s[2:7]; s <- parallel::nextRNGSubStream(s) }
  out
}

# replaces:  bf.perm <- bplapply(..., permMclustGene(...))   R/scDD.R:458-463  and the per-gene loops over
# permMclust / permMclustCov (:414-455).  The device draws the reference's own permutations: sample(1:n) from the
# streams above with a bit-exact sampler, including the runif jitter mclustRestricted draws mid-stream when a permuted
# subset is constant.  One call for all genes, sharded over the devices of scdd_set_devices(); Ctrl-C is honoured
# between chunks of 16 permutations.
permMclust_gpu <- function(dat, condition, ref, permutations, prior, min.size, RNGseed,
                           parallelBy = c("Genes", "Permutations"), C = NULL, perm_idx = NULL) {
  parallelBy <- match.arg(parallelBy)
  by.perm <- parallelBy == "Permutations"
  g <- .gene_csr(dat, condition, ref)
  cdr <- if (is.null(C)) NULL else cell_covariate_gpu(dat, C, g$off)      # C[y>0] of every gene, R/permutations.R:216
  seed6 <- if (is.null(perm_idx)) rng_streams(RNGseed, if (by.perm) permutations else nrow(dat)) else NULL
  r <- .Call(C_scdd_perm_bf, g$vals, g$off, g$cond01, as.integer(permutations),
             if (is.null(perm_idx)) NULL else as.integer(perm_idx), seed6, by.perm, cdr,
             as.double(prior), as.integer(min.size))
  if (r$const.sets > 0)
    message("Notice: ", r$const.sets, " permuted subsets had identical values and were jittered as mclustRestricted does")
  if (r$failed.sets > 0)
    warning(r$failed.sets, " permuted fit-sets had no valid model; the p-values of their genes are NA")
  m <- r$bf.perm                                   # permutations x genes
  lapply(seq_len(ncol(m)), function(j) m[, j])     # bf.perm as the list scDD() expects
}

# replaces the bplapply inside testKS, R/Test_KS.R:67
testKS_gpu <- function(dat, condition, inclZero = TRUE) {
  lev <- unique(condition)
  keep <- condition %in% lev[1:2]
  g <- .gene_csr(dat[, keep, drop = FALSE], condition[keep], lev[1], log = FALSE, inclZero = inclZero)
  .Call(C_scdd_ks, g$vals, g$off, g$cond01)$p
}

# the per-gene pieces of feDP for every gene of the block in one call: p.shift (Welch t / lm coefficient), p.fisher,
# counts (luOutlier of oa / c1 / c2).  classifyDD_gpu() needs p.shift too, so scDD() computes this once.
fedp_stats_gpu <- function(g, fit, cdr = NULL, min.size = 3)
  .Call(C_scdd_fedp, g$vals, g$off, g$cond01, as.integer(fit$class.oa), as.integer(fit$class.c),
        if (is.null(cdr)) NULL else as.double(cdr), as.integer(min.size))

# replaces feDP(), R/classify.dd.R:314-383: the per-gene loop runs on the device, the labelling (:377-382) stays R.
# g: the .gene_csr() list of pe_mat; fit: the list C_scdd_fit_observed returned (class.oa, class.c); cdr: the cellular
# detection rate of every nonzero cell in CSR order when adjust.perms (else NULL); stats: fedp_stats_gpu() if at hand.
feDP_gpu <- function(g, fit, sig_genes, nrow_pe, cdr = NULL, testZeroes = FALSE, min.size = 3, stats = NULL) {
  pval.thresh <- if (testZeroes) 0.025 else 0.05
  r <- if (is.null(stats)) fedp_stats_gpu(g, fit, cdr, min.size) else stats
  ns_genes <- (1:nrow_pe)[-sig_genes]              # R quirk kept: empty when there is no significant gene
  cnt <- r$counts[, ns_genes, drop = FALSE]
  cat <- ifelse(!is.na(r$p.fisher[ns_genes]) & r$p.fisher[ns_genes] < 0.05, "DP", "NS")
  pval.ns <- p.adjust(r$p.shift[ns_genes], method = "BH")
  cat[pval.ns < pval.thresh & cat != "DP"] <- "NC"
  cat[pval.ns < pval.thresh & cnt[2, ] == cnt[3, ] & cnt[2, ] == cnt[1, ] & cnt[2, ] > 1] <- "DP"
  names(pval.ns) <- cat
  pval.ns
}

# replaces classifyDD(), R/classify.dd.R:86-238, for the genes `these` (1-based) of the fitted block: the label
# bookkeeping and the posterior parameters run on host threads of the library, the 10,000-draw Monte-Carlo comparison
# of every pair of cluster means (:165-187) on the device.  The draws come from Philox streams keyed by `seed`, not from
# R's RNG: same distribution, another stream (the reference's own classifyDD consumes .Random.seed); set
# options(scdd_b200.classify = "R") to keep the reference's function instead.
classifyDD_gpu <- function(g, fit, these, prior, min.size, p.shift, seed = 1L) {
  if (length(these) == 0) return(character(0))
  code <- .Call(C_scdd_classify, g$vals, g$off, g$cond01, as.integer(fit$class.oa), as.integer(fit$class.c),
                as.integer(these), as.double(p.shift), as.double(prior), as.integer(min.size), as.integer(seed))
  c("NC", "DE", "DP", "DM", "DB")[code + 1L]
}

# replaces the bplapply of arm::bayesglm fits inside testZeroes, R/classify.dd.R:262-286.  factor(cond) sorts its
# levels, so the baseline of the condition dummy is the smallest label -- not scDD()'s `ref` (first seen).
testZeroes_gpu <- function(dat, cond, these = seq_len(nrow(dat))) {
  storage.mode(dat) <- "double"
  .Call(C_scdd_test_zeroes, dat, nrow(dat), ncol(dat), as.integer(factor(cond)) - 1L, as.integer(these))
}

# BiocParallel::register(param) of R/scDD.R:314 becomes a device list
scdd_set_devices <- function(devices = NULL) invisible(.Call(C_scdd_set_devices, if (is.null(devices)) NULL else as.integer(devices)))
scdd_device_count <- function() .Call(C_scdd_device_count)
scdd_release_cache <- function() invisible(.Call(C_scdd_release_cache))
