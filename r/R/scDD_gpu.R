# scDD_gpu.R -- the R-side shim a maintainer adds to the reference package: three helpers that replace the
# three bplapply fan-outs of the hot path (R/scDD.R:343/393, R/scDD.R:420-465, R/Test_KS.R:67) with one
# .Call each.  Everything else of scDD() (argument checks, tofit/skipConstant, classifyDD, feDP, testZeroes,
# Genes table) stays the reference's own R code.  NAMESPACE gains  useDynLib(scDD, .registration = TRUE).

.gene_csr <- function(dat, condition, ref, log = TRUE, inclZero = FALSE) {
  # dat: genes x cells; returns the CSR-over-genes layout of include/scdd_b200.h (vals, cond01, off) -- the
  # library walks R's column-major matrix directly, one threaded pass to count and one to gather
  storage.mode(dat) <- "double"
  .Call(C_scdd_gene_csr, dat, nrow(dat), ncol(dat), as.integer(condition != ref), log, inclZero)
}

# replaces the MAP/MAP1/MAP2 loop of R/scDD.R:528-552; which = -1 (Zhat.combined), 0 (Zhat.c1), 1 (Zhat.c2);
# labels = unlist of oa$class (combined) or of the c1/c2 classes scattered back to cell order ("class.c")
zhat_gpu <- function(dat, condition, ref, tofit, off, labels, which = -1L) {
  storage.mode(dat) <- "double"
  .Call(C_scdd_zhat, dat, nrow(dat), ncol(dat), as.integer(condition != ref), as.integer(which),
        as.integer(tofit), off, as.integer(labels))
}

# replaces  comps.all / comps.c1 / comps.c2 <- luOutlier(...)  of R/scDD.R:335-337 for all genes at once
luOutlier_gpu <- function(labels, cond01, off, min.size, which = -1L)
  .Call(C_scdd_cluster_counts, as.integer(labels), as.integer(cond01), off, as.integer(min.size), as.integer(which))

# replaces:  out <- bplapply(1:nrow(...), function(x) genefit(normcounts(SCdat)[tofit[x],]))   R/scDD.R:393
genefit_gpu <- function(dat, condition, ref, prior, min.size, with.bf = TRUE) {
  g <- .gene_csr(dat, condition, ref)
  r <- .Call(C_scdd_fit_observed, g$vals, g$off, g$cond01, as.double(prior), as.integer(min.size), with.bf)
  ng <- nrow(dat)
  mk <- function(i, cls) {                        # the list mclustRestricted returns, R/mclust.restricted.R:249-252
    G <- r$G[i]
    list(class = cls, mean = setNames(r$mean[seq_len(G), i], seq_len(G)), var = r$var[seq_len(G), i],
         df = if (G > 1) 3 * G - 1 else 2, n = r$n[i], bic = r$bic[i], loglik = r$loglik[i],
         model = if (G > 1) "V" else "X")
  }
  oa <- c1 <- c2 <- vector("list", ng)
  for (j in seq_len(ng)) {
    s <- (g$off[j] + 1):g$off[j + 1]
    c0 <- g$cond01[s] == 0L
    oa[[j]] <- mk(3 * j - 2, r$class.oa[s])
    c1[[j]] <- mk(3 * j - 1, r$class.c[s][c0])
    c2[[j]] <- mk(3 * j, r$class.c[s][!c0])
  }
  # raw / csr: the flat class vectors and the CSR block, for feDP_gpu(), zhat_gpu() and luOutlier_gpu()
  list(oa = oa, c1 = c1, c2 = c2, bf = r$bf, den = r$den, raw = r, csr = g)
}

# replaces:  bf.perm <- bplapply(..., permMclustGene(...))   R/scDD.R:458-463  (and :420-455)
# Permutations identical to the reference's: either draw them here with R's own sample() under the same
# bplapply(RNGseed) streams and pass perm_idx, or pass the per-gene L'Ecuyer states (.Random.seed[2:7]) and let
# the device draw them with the bit-exact sampler.
permMclust_gpu <- function(dat, condition, ref, permutations, prior, min.size, seed6 = NULL, perm_idx = NULL,
                           C = NULL) {
  g <- .gene_csr(dat, condition, ref)
  cdr <- if (is.null(C)) NULL else as.double(rep(C, nrow(dat))[as.vector(t(dat > 0))])
  m <- .Call(C_scdd_perm_bf, g$vals, g$off, g$cond01, as.integer(permutations), perm_idx, seed6, cdr,
             as.double(prior), as.integer(min.size))
  lapply(seq_len(ncol(m)), function(j) m[, j])   # bf.perm as the list scDD() expects
}

gene_streams <- function(RNGseed, ngenes) {
  # the stream BiocParallel hands to element i of bplapply(seq_len(ngenes), ...)
  old <- RNGkind("L'Ecuyer-CMRG"); on.exit(do.call(RNGkind, as.list(old)))
  set.seed(RNGseed)
  s <- .Random.seed
  out <- matrix(0L, 6, ngenes)
  for (i in seq_len(ngenes)) { out[, i] <- s[2:7]; s <- parallel::nextRNGSubStream(s) }
  out
}

# replaces the bplapply inside testKS, R/Test_KS.R:67
testKS_gpu <- function(dat, condition, inclZero = TRUE) {
  lev <- unique(condition)
  keep <- condition %in% lev[1:2]
  g <- .gene_csr(dat[, keep, drop = FALSE], condition[keep], lev[1], log = FALSE, inclZero = inclZero)
  .Call(C_scdd_ks, g$vals, g$off, g$cond01)$p
}

# replaces feDP(), R/classify.dd.R:314-383: the per-gene loop runs on the device, the labelling (:377-382) stays R.
# g: the .gene_csr() list of pe_mat; fit: the list C_scdd_fit_observed returned (class.oa, class.c); C: the cellular
# detection rate per cell when adjust.perms (else NULL).  NaN p.shift marks genes where the reference would jitter
# with runif(): those few are recomputed by the reference's own loop body.
feDP_gpu <- function(g, fit, sig_genes, nrow_pe, C = NULL, testZeroes = FALSE, min.size = 3, keep = NULL) {
  pval.thresh <- if (testZeroes) 0.025 else 0.05
  cdr <- if (is.null(C)) NULL else as.double(rep(C, nrow_pe)[keep])   # keep = as.vector(t(pe_mat > 0))
  r <- .Call(C_scdd_fedp, g$vals, g$off, g$cond01, fit$class.oa, fit$class.c, cdr, as.integer(min.size))
  ns_genes <- (1:nrow_pe)[-sig_genes]
  cnt <- r$counts[, ns_genes, drop = FALSE]
  cat <- ifelse(!is.na(r$p.fisher[ns_genes]) & r$p.fisher[ns_genes] < 0.05, "DP", "NS")
  pval.ns <- p.adjust(r$p.shift[ns_genes], method = "BH")
  cat[pval.ns < pval.thresh & cat != "DP"] <- "NC"
  cat[pval.ns < pval.thresh & cnt[2, ] == cnt[3, ] & cnt[2, ] == cnt[1, ] & cnt[2, ] > 1] <- "DP"
  names(pval.ns) <- cat
  pval.ns
}
