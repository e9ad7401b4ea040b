#!/usr/bin/env Rscript
# dump_golden.R -- pins the CPU oracle (oracle/scdd_oracle.c) to the REAL reference.
#
# The build image has no R, so every mclust / base-R / BiocParallel fact the oracle restates is from memory
# (SURVEY.md App. A-C, "[MEM]").  One run of this script on ANY machine with R >= 4.0, scDD, mclust and BiocParallel
# writes the reference's own numbers for the two bundled data sets into tests/golden/r_golden_<name>.json;
# tests/test_r_golden.py then checks the oracle against them (it is skipped while the files are absent).
#
#   Rscript tools/dump_golden.R [outdir = tests/golden] [ngenes = all] [B = 20]
#
# What is dumped, per data set (scDatExSim 30 genes, scDatEx 500 genes), per gene, per fit-set (oa, c1, c2):
#   * mclust: the G = 1..5 table of Mclust(y, modelNames = "V", G = g): bic, loglik, mean, sigmasq, pro, classification
#     (A.2-A.5), plus the qclass() start of every G (A.4)
#   * scDD:::mclustRestricted(y, restrict = TRUE, min.size = 3): class, mean, var, bic, loglik       (R/mclust.restricted.R)
#   * scDD:::jointPosterior of the three fits, bf and den                                             (R/scDD.R:381-383)
#   * permutation numerators of scDD:::permMclustGene under bplapply(SerialParam(RNGseed = 816)), B permutations,
#     and the permutations themselves (sample(1:n) per stream)                                        (R/permutations.R:211-286)
#   * ks.test(x1, x2, exact = FALSE): statistic and p.value on the nonzero raw values                 (R/Test_KS.R:56-65)
#   * scDD:::testZeroes over all rows, and per gene the arm::bayesglm fit behind it: coefficients, standard errors,
#     iterations, deviance, prior.scale, prior.sd                                                      (R/classify.dd.R:262-286)
# and once: versions, RNG known answers (Mersenne-Twister and L'Ecuyer-CMRG sample()/runif, nextRNGStream,
# nextRNGSubStream, the .Random.seed BiocParallel hands to bplapply elements 1..4 for RNGseed = 7).
suppressPackageStartupMessages({
  library(scDD); library(mclust); library(BiocParallel); library(SingleCellExperiment)
})
args <- commandArgs(trailingOnly = TRUE)
outdir <- if (length(args) >= 1) args[1] else "tests/golden"
ngmax  <- if (length(args) >= 2) as.integer(args[2]) else NA
B      <- if (length(args) >= 3) as.integer(args[3]) else 20L
prior  <- list(alpha = 0.01, mu0 = 0, s0 = 0.01, a0 = 0.01, b0 = 0.01)     # R/scDD.R:203

num <- function(x) {                                   # JSON number array with 17 significant digits, NA -> null
  if (length(x) == 0) return("[]")
  s <- sprintf("%.17g", as.numeric(x)); s[is.na(x)] <- "null"; s[is.infinite(x)] <- "null"
  paste0("[", paste(s, collapse = ","), "]")
}
int <- function(x) { if (length(x) == 0) return("[]"); s <- as.character(as.integer(x)); s[is.na(x)] <- "null"; paste0("[", paste(s, collapse = ","), "]") }
obj <- function(...) { a <- list(...); paste0("{", paste(sprintf("\"%s\":%s", names(a), unlist(a)), collapse = ","), "}") }
arr <- function(v) paste0("[", paste(v, collapse = ","), "]")
str <- function(s) paste0("\"", gsub("\"", "'", s), "\"")

qclass <- get("qclass", envir = asNamespace("mclust"))

fit_table <- function(y) {
  rows <- lapply(1:5, function(g) {
    m <- if (g <= length(y)) suppressWarnings(tryCatch(Mclust(y, warn = FALSE, modelNames = "V", G = g, verbose = FALSE),
                                                       error = function(e) NULL)) else NULL
    q <- if (g >= 2 && g <= length(y)) tryCatch(qclass(y, g), error = function(e) rep(NA, length(y))) else rep(1, length(y))
    if (is.null(m)) return(obj(G = g, valid = "false", qclass = int(q)))
    obj(G = g, valid = "true", bic = num(m$bic), loglik = num(m$loglik), mean = num(m$parameters$mean),
        sigmasq = num(m$parameters$variance$sigmasq), pro = num(m$parameters$pro),
        classification = int(m$classification), qclass = int(q), modelName = str(m$modelName), df = num(m$df))
  })
  arr(rows)
}
restricted <- function(y) {
  r <- suppressWarnings(scDD:::mclustRestricted(y, restrict = TRUE, min.size = 3))
  jp <- scDD:::jointPosterior(y, r, prior$alpha, prior$mu0, prior$s0, prior$a0, prior$b0)
  obj(class = int(r$class), mean = num(r$mean), var = num(r$var), df = num(r$df), n = num(r$n), bic = num(r$bic),
      loglik = num(r$loglik), model = str(r$model), jointPosterior = num(jp))
}

dump_one <- function(name, sce) {
  X <- normcounts(sce); cond <- colData(sce)$condition; ref <- unique(cond)[1]
  ng <- if (is.na(ngmax)) nrow(X) else min(ngmax, nrow(X))
  genes <- lapply(seq_len(ng), function(g) {
    y0 <- X[g, ]; c0 <- cond[y0 > 0]; y <- log(y0[y0 > 0])
    ks <- suppressWarnings(ks.test(y0[cond == ref & y0 > 0], y0[cond != ref & y0 > 0], exact = FALSE))
    obj(gene = str(rownames(X)[g]), n = length(y), n1 = sum(c0 == ref),
        oa_table = fit_table(y), c1_table = fit_table(y[c0 == ref]), c2_table = fit_table(y[c0 != ref]),
        oa = restricted(y), c1 = restricted(y[c0 == ref]), c2 = restricted(y[c0 != ref]),
        ks_D = num(ks$statistic), ks_p = num(ks$p.value))
  })
  # permutations exactly as scDD(parallelBy = "Genes") runs them: one bplapply element per gene (R/scDD.R:458)
  C <- apply(X[seq_len(ng), , drop = FALSE], 2, function(x) sum(x > 0) / length(x))
  perm <- bplapply(seq_len(ng), function(g) {
    seed0 <- .Random.seed
    n <- sum(X[g, ] > 0)
    bf <- scDD:::permMclustGene(X[g, ], FALSE, B, cond, remove.zeroes = TRUE, log.transf = TRUE, restrict = TRUE,
                                min.size = 3, prior$alpha, prior$mu0, prior$s0, prior$a0, prior$b0, C, ref)
    assign(".Random.seed", seed0, envir = globalenv())
    first <- sample(1:n, replace = FALSE)                     # the first permutation of this element's stream
    list(seed = seed0, bf = bf, first = first)
  }, BPPARAM = SerialParam(RNGseed = 816))
  perms <- lapply(perm, function(p) obj(random_seed = int(p$seed), bf_perm = num(p$bf), first_perm = int(p$first)))
  # testZeroes: the p-values the package returns, and the bayesglm objects behind them (all rows count in `detection`)
  pz <- scDD:::testZeroes(X, cond)
  detection <- colSums(X > 0) / nrow(X)
  zfits <- lapply(seq_len(ng), function(g) {
    y <- X[g, ]
    if (sum(y == 0) == 0) return(obj(na = "true"))
    M1 <- suppressWarnings(arm::bayesglm(y > 0 ~ detection + factor(cond), family = binomial(link = "logit"), Warning = FALSE))
    cf <- summary(M1)$coefficients
    obj(na = "false", coef = num(cf[, 1]), se = num(cf[, 2]), p = num(cf[, 4]), iter = num(M1$iter), deviance = num(M1$deviance),
        prior_scale = num(M1$prior.scale), prior_sd = num(M1$prior.sd), converged = str(as.character(M1$converged)))
  })
  json <- obj(name = str(name), B = B, RNGseed = 816, genes = arr(genes), perms = arr(perms),
              zero_p = num(pz), zero_fits = arr(zfits), arm = str(as.character(packageVersion("arm"))))
  writeLines(json, file.path(outdir, paste0("r_golden_", name, ".json")))
  message("wrote ", name, ": ", ng, " genes")
}

# ---- RNG known answers + versions
RNGkind("Mersenne-Twister", sample.kind = "Rejection"); set.seed(1); mt1 <- sample(1:10); mtu <- runif(3)
RNGkind("L'Ecuyer-CMRG"); set.seed(1); le1 <- sample(1:10); leu <- runif(3)
set.seed(123); s123 <- .Random.seed
ns <- parallel::nextRNGStream(s123); nss <- parallel::nextRNGSubStream(s123)
bp <- bplapply(1:4, function(i) .Random.seed, BPPARAM = SerialParam(RNGseed = 7))
set.seed(5); big <- sample(1:40000)[1:8]                      # > 32768: two 16-bit chunks per index
RNGkind("Mersenne-Twister")
meta <- obj(R = str(R.version.string), mclust = str(as.character(packageVersion("mclust"))),
            scDD = str(as.character(packageVersion("scDD"))), BiocParallel = str(as.character(packageVersion("BiocParallel"))),
            mt_sample_1_10_seed1 = int(mt1), mt_runif3_after = num(mtu),
            lecuyer_sample_1_10_seed1 = int(le1), lecuyer_runif3_after = num(leu),
            lecuyer_seed123 = int(s123), nextRNGStream = int(ns), nextRNGSubStream = int(nss),
            bplapply_seeds_RNGseed7 = arr(lapply(bp, int)), lecuyer_sample_40000_seed5_first8 = int(big))
dir.create(outdir, showWarnings = FALSE, recursive = TRUE)
writeLines(meta, file.path(outdir, "r_golden_meta.json"))
data(scDatExSim); dump_one("scDatExSim", scDatExSim)
data(scDatEx);    dump_one("scDatEx", scDatEx)
