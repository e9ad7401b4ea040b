# scDD_b200.R -- drop-in dispatcher: scDD() with the reference's exact signature, argument checks, messages and result
# objects (R/scDD.R:217-602), with the per-gene work of the hot path done by libscdd_b200.so through the helpers of
# scDD_gpu.R.  A maintainer replaces the body of R/scDD.R by this file (or sources it after the package) -- nothing
# else of the package changes: results() is the reference's own, classifyDD() too when options(scdd_b200.classify = "R")
# (default: the library's, see classifyDD_gpu); testKS() and testZeroes() are redefined at the end of this file with the
# reference's signatures.
#
# Written from the reference's behaviour, not from its text: the three bplapply fan-outs and the per-gene R loops are
# gone, so the function is organised around flat vectors (one CSR block for all fitted genes) instead of lists.
# UNTESTED IN R -- neither the build image nor the GPU box has R; the same dispatch runs in Python as scdd_b200/api.py
# (tests/test_gpu_parity.py::test_scdd_api_end_to_end) and every .Call routine is executed through a mock of R's C API
# (tests/test_r_glue.py).
scDD <- function(SCdat,
                 prior_param = list(alpha = 0.10, mu0 = 0, s0 = 0.01, a0 = 0.01, b0 = 0.01),
                 permutations = 0,
                 testZeroes = TRUE, adjust.perms = FALSE,
                 param = bpparam(),
                 parallelBy = c("Genes", "Permutations"),
                 condition = "condition", min.size = 3,
                 min.nonzero = NULL,
                 level = 0.05, categorize = TRUE) {
  ## ---- argument checks (same conditions, same messages) ----
  if (!("SingleCellExperiment" %in% class(SCdat)))
    stop("Please provide a valid 'SingleCellExperiment' object.")
  if (is.null(assayNames(SCdat)) || !("normcounts" %in% assayNames(SCdat)))
    stop("Please make sure the 'SingleCellExperiment' object includes ",
         "an assays slot named 'normcounts' that contains normalized ",
         "counts on the original scale")
  if (!categorize && permutations != 0)
    stop("Categorization will be carried out if permuations are run.")
  parallelBy <- match.arg(parallelBy)
  prior <- c(prior_param$alpha, prior_param$mu0, prior_param$s0, prior_param$a0, prior_param$b0)
  X <- as.matrix(normcounts(SCdat))
  storage.mode(X) <- "double"
  cond <- colData(SCdat)[[condition]]
  if (length(unique(cond)) != 2 | length(cond) != ncol(X))
    stop("Error: Please specify valid condition labels.")
  ref <- unique(cond)[1]
  isref <- cond == ref

  ## ---- gene filters: one threaded pass instead of rowSums() + two apply(unique) ----
  sc <- dense_scan_gpu(X, cond, ref)
  if (sc$n.negative > 0)
    stop(paste0("Error: Negative values for Normalized Expression counts ",
                "detected. Please ensure all counts are non-negative"))
  if (is.null(min.nonzero)) min.nonzero <- min.size
  need <- max(min.size, 2, min.nonzero)
  tofit <- which(sc$nnz0 >= need & sc$nnz1 >= need)
  todo <- if (testZeroes) " Only testing for DZ for these genes." else " Skipping these genes."
  if (length(tofit) < nrow(X))
    message(paste0("Notice: ", nrow(X) - length(tofit), " genes have less than ", min.nonzero,
                   " nonzero cells per condition. ", todo))
  skipConstant <- which(sc$const0[tofit] | sc$const1[tofit])
  if (length(skipConstant) > 0) {
    message(paste0("Notice: ", length(skipConstant), " Genes have constant nonzero values. ", todo))
    tofit <- tofit[-skipConstant]
  }

  ## ---- "parallel back-end": the devices of the box; the BiocParallel param only carries the RNG seed ----
  ndev <- scdd_device_count()
  if (ndev < 1) stop("scdd_b200: no CUDA device available (the library has no CPU fallback)")
  message(paste0("Setting up parallel back-end using ", ndev, " GPUs"))
  scdd_set_devices(seq_len(ndev) - 1L)
  RNGseed <- tryCatch(BiocParallel::bpRNGseed(param), error = function(e) NULL)
  if (is.null(RNGseed) || is.na(RNGseed)) RNGseed <- sample.int(.Machine$integer.max, 1)

  ng <- length(tofit)
  Xf <- X[tofit, , drop = FALSE]
  oa <- c1 <- c2 <- vector("list", ng)
  bf <- den <- comps.all <- comps.c1 <- comps.c2 <- rep(NA, ng)
  fit <- NULL

  ## ---- observed fits: genefit for every gene in one call (R/scDD.R:327-353, :373-404) ----
  if (categorize) {
    message("Clustering observed expression data for each gene")
    fit <- genefit_gpu(Xf, cond, ref, prior, min.size, with.bf = permutations != 0)
    oa <- fit$oa; c1 <- fit$c1; c2 <- fit$c2
    comps.all <- luOutlier_gpu(fit$raw$class.oa, fit$csr$cond01, fit$csr$off, min.size, -1L)
    comps.c1  <- luOutlier_gpu(fit$raw$class.c,  fit$csr$cond01, fit$csr$off, min.size, 0L)
    comps.c2  <- luOutlier_gpu(fit$raw$class.c,  fit$csr$cond01, fit$csr$off, min.size, 1L)
  }
  C <- if (adjust.perms) colSums(Xf > 0) / nrow(Xf) else NULL        # R/scDD.R:416-417, :457

  if (permutations == 0) {
    message("Notice: Number of permutations is set to zero; using
            Kolmogorov-Smirnov to test for differences in distributions
            instead of the Bayes Factor permutation test")
    p.unadj <- testKS_gpu(Xf, cond, inclZero = FALSE)                # R/Test_KS.R:55-83 through the device
    p.adj <- p.adjust(p.unadj, method = "BH")
    sig <- which(p.adj < if (testZeroes) level / 2 else level)
    pvals <- p.unadj
  } else {
    bf <- fit$bf; den <- fit$den
    message("Performing permutations to evaluate independence of clustering and condition for each gene")
    message(paste0("Parallelizing by ", parallelBy))
    bf.perm <- permMclust_gpu(Xf, cond, ref, permutations, prior, min.size, RNGseed, parallelBy = parallelBy, C = C)
    obs <- if (adjust.perms) bf - den else bf                        # R/scDD.R:467-473: strict >, divisor = permutations
    pvals <- vapply(seq_len(ng), function(x) sum(bf.perm[[x]] > obs[x]), numeric(1)) / permutations
    sig <- which(p.adjust(pvals, method = "BH") < if (testZeroes) level / 2 else level)
  }

  cats.all <- pvals.all <- rep(NA, nrow(X))
  pvals.all[tofit] <- pvals

  ## ---- categories: feDP's per-gene loop and classifyDD's Monte-Carlo comparisons are kernels ----
  if (categorize) {
    message("Classifying significant genes into patterns")
    cdr <- if (adjust.perms) cell_covariate_gpu(Xf, C, fit$csr$off) else NULL
    fe <- fedp_stats_gpu(fit$csr, fit$raw, cdr, min.size)            # p.shift / p.fisher / counts of every gene, once
    on.device <- getOption("scdd_b200.classify", "gpu") == "gpu"
    cseed <- if (is.null(RNGseed)) 1L else as.integer(RNGseed)
    classify <- function(these, k) {
      if (on.device) {
        classifyDD_gpu(fit$csr, fit$raw, these, prior, min.size, fe$p.shift, seed = cseed + k)
      } else {
        classifyDD(Xf, cond, these, oa, c1, c2, alpha = prior[1], m0 = prior[2], s0 = prior[3], a0 = prior[4],
                   b0 = prior[5], log.nonzero = TRUE, ref = ref, min.size = min.size)
      }
    }
    cats <- rep("NS", ng)
    cats[sig] <- classify(sig, 0L)
    extraDP <- feDP_gpu(fit$csr, fit$raw, sig, ng, cdr = cdr, testZeroes = testZeroes, min.size = min.size, stats = fe)
    cats[-sig] <- names(extraDP)
    NCs <- which(p.adjust(pvals, method = "BH") > (if (testZeroes) level / 2 else level) & cats == "NC")
    cats[NCs] <- classify(NCs, 1L)
    cats.all[tofit] <- cats
  }

  ## ---- zero test: the Bayesian logistic regressions of arm::bayesglm as one kernel (testZeroes below) ----
  pvals.z <- rep(NA, nrow(X))
  if (testZeroes) {
    pvals.z <- testZeroes(X, cond)
    pz <- p.adjust(pvals.z, method = "BH")
    nodd <- !(cats.all %in% c("DE", "DP", "DM", "DB"))
    cats.all[pz < level / 2 & nodd] <- "DZ"
    if (categorize) cats.all[pz >= level / 2 & nodd] <- "NS"
  }

  ## ---- Zhat matrices: three threaded passes instead of a per-gene S4 subsetting loop (R/scDD.R:528-552) ----
  if (categorize) {
    MAP  <- zhat_gpu(X, cond, ref, tofit, fit$csr$off, fit$raw$class.oa, -1L)
    MAP1 <- zhat_gpu(X, cond, ref, tofit, fit$csr$off, fit$raw$class.c, 0L)
    MAP2 <- zhat_gpu(X, cond, ref, tofit, fit$csr$off, fit$raw$class.c, 1L)
    rownames(MAP1) <- rownames(MAP2) <- rownames(MAP) <- rownames(SCdat)
    colnames(MAP1) <- colnames(SCdat)[isref]
    colnames(MAP2) <- colnames(SCdat)[!isref]
    colnames(MAP) <- colnames(SCdat)
  } else {
    MAP1 <- MAP2 <- MAP <- "Categorize set to FALSE; no clustering provided."
  }

  spread <- function(v) { out <- rep(NA, nrow(X)); out[tofit] <- v; out }
  Genes <- data.frame(gene = rownames(SCdat),
                      DDcategory = cats.all,
                      Clusters.combined = spread(comps.all),
                      Clusters.c1 = spread(comps.c1),
                      Clusters.c2 = spread(comps.c2),
                      nonzero.pvalue = pvals.all,
                      nonzero.pvalue.adj = p.adjust(pvals.all, method = "BH"))
  if (testZeroes) {
    fisherCombined <- function(x) {                                   # Fisher's method over the available p-values
      ok <- !is.na(x)
      if (all(ok)) pchisq(-2 * sum(log(x)), df = 2 * length(x), lower.tail = FALSE)
      else if (sum(!ok) == 1) x[ok] else NA
    }
    pvals.comb <- apply(cbind(pvals.all, pvals.z), 1, fisherCombined)
    Genes <- cbind(Genes,
                   zero.pvalue = pvals.z,
                   zero.pvalue.adj = p.adjust(pvals.z, method = "BH"),
                   combined.pvalue = pvals.comb,
                   combined.pvalue.adj = p.adjust(pvals.comb, method = "BH"))
  }
  rownames(Genes) <- rownames(SCdat)

  metadata(SCdat)[["Genes"]] <- Genes
  metadata(SCdat)[["Zhat.combined"]] <- MAP
  metadata(SCdat)[["Zhat.c1"]] <- MAP1
  metadata(SCdat)[["Zhat.c2"]] <- MAP2
  return(SCdat)
}

# testKS() with the reference's signature and return value (R/Test_KS.R:55-83); the per-gene ks.test calls are one kernel
testKS <- function(dat, condition, inclZero = TRUE, numDE = NULL, DEIndex) {
  ks.pval.unadj <- testKS_gpu(dat, condition, inclZero)
  ks.pval <- p.adjust(ks.pval.unadj, method = "BH")
  sig_genes_ks <- which(ks.pval < 0.05)
  names(sig_genes_ks) <- rownames(dat)[sig_genes_ks]
  names(ks.pval) <- rownames(dat)
  if (!is.null(numDE)) {
    power <- sum(sig_genes_ks %in% DEIndex) / numDE
    fdr <- sum(!(sig_genes_ks %in% DEIndex)) / length(sig_genes_ks)
    return(list(genes = sig_genes_ks, p = ks.pval, p.unadj = ks.pval.unadj, power = power, fdr = fdr))
  }
  list(genes = sig_genes_ks, p = ks.pval, p.unadj = ks.pval.unadj)
}

# testZeroes() with the reference's signature and return value (R/classify.dd.R:262-286): the unadjusted p-value of the
# condition coefficient per gene of `these`, NA for a gene without zeros
testZeroes <- function(dat, cond, these = 1:nrow(dat)) testZeroes_gpu(as.matrix(dat), cond, these)
