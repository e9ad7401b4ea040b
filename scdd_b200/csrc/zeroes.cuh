// zeroes.cuh -- device side of testZeroes (R/classify.dd.R:262-286): per gene, the Bayesian logistic regression
//     arm::bayesglm(y > 0 ~ detection + factor(cond), family = binomial(link = "logit"))
// and the Wald p-value of the condition coefficient (summary(M1)$coefficients[3, 4]).
//
// The response of a gene is one BIT per cell (nonzero or not) and the two predictors are shared by all genes, so the
// whole problem of a 20,000 x 2,000 matrix is 5 MB of bit rows plus two vectors.  One warp per gene runs bayesglm.fit's
// loop (Gelman et al. 2008, section 3; restated with its defaults in oracle/zeroes_oracle.py): every pass over the
// cells evaluates the deviance at the current coefficients AND accumulates the weighted normal equations of the next
// IRLS step (6 + 3 sums), the 3 x 3 system augmented with the prior precisions 1 / prior.sd^2 is solved by Cholesky in
// registers, and the EM step re-estimates prior.sd from the coefficients and diag(V).  fp64 throughout; the cost is
// exp + log + three divisions per cell and pass (~1e9 cell-passes for config 3: milliseconds), three orders of
// magnitude below fit_kernel, so the kernel is written for clarity, not tuned.
#pragma once
#include <cstdint>

namespace scdd {

struct ZeroGlmArgs {
    const uint32_t* mask;     // [ngenes * W] bit c % 32 of word c / 32 = (y[c] > 0)
    const uint32_t* lvl;      // [W] bit = cell is in the 2nd level of factor(cond)
    const double* det;        // [ncells] detection rate, R/classify.dd.R:263
    const int32_t* nzero;     // [ngenes] sum(y == 0); 0 -> NA (:267)
    int64_t W;
    int ncells, ngenes;
    double scale[3];          // prior.scale after bayesglm's scaled = TRUE rule: intercept, detection, condition
    double df;                // prior.df (1 = Cauchy)
    double epsilon;           // glm.control: 1e-8
    int maxit;                // 100
    double* pvalue;           // [ngenes], NaN = NA
    double* detail;           // nullable [ngenes * 8]: coef[3], se[3], iterations, deviance
};

constexpr double GLM_EPS = 2.220446049250313e-16;   // DBL_EPSILON
constexpr double GLM_THRESH = 30.0;                 // family.c THRESH / MTHRESH

__device__ __forceinline__ double zsum(double v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}

__global__ void zero_glm_kernel(ZeroGlmArgs a) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < a.ngenes; g += warps) {
        if (a.nzero[g] <= 0) {                          // no zero to explain: NA
            if (lane == 0) {
                a.pvalue[g] = __longlong_as_double(0x7ff8000000000000LL);
                if (a.detail) for (int k = 0; k < 8; k++) a.detail[(size_t)g * 8 + k] = __longlong_as_double(0x7ff8000000000000LL);
            }
            continue;
        }
        const uint32_t* mrow = a.mask + (size_t)g * a.W;
        double b0 = 0, b1 = 0, b2 = 0;
        double sd0 = a.scale[0], sd1 = a.scale[1], sd2 = a.scale[2];   // prior.sd starts at prior.scale
        double v00 = 0, v11 = 0, v22 = 0;
        double dev = 0, devold = 0;
        int iter = 0;
        // S = X' W X (s00 s01 s02 s11 s12 s22; s22 = s02 because the dummy is 0/1), t = X' W z
        double S[5], T[3];
        // one pass over the cells: deviance at the current linear predictor, normal equations of the next step
        auto pass = [&](bool first) {
            double s00 = 0, s01 = 0, s02 = 0, s11 = 0, s12 = 0, t0 = 0, t1 = 0, t2 = 0, dv = 0;
            for (int64_t k = 0; k < a.W; k++) {
                const int c = (int)(k * 32) + lane;
                if (c >= a.ncells) continue;
                const bool y = (mrow[k] >> lane) & 1u;
                const bool d = (a.lvl[k] >> lane) & 1u;
                const double x = a.det[c];
                double eta;
                if (first) {
                    const double mu0 = y ? 0.75 : 0.25;                 // mustart = (y + 0.5) / 2, binomial()$initialize
                    eta = log(mu0 / (1.0 - mu0));
                } else {
                    eta = b0 + b1 * x + (d ? b2 : 0.0);
                }
                // logit_linkinv / logit_mu_eta of R's family.c, with their clamps
                const double e = exp(eta);
                const double tmp = (eta < -GLM_THRESH) ? GLM_EPS : ((eta > GLM_THRESH) ? 1.0 / GLM_EPS : e);
                const double mu = tmp / (1.0 + tmp);
                const double opexp = 1.0 + e;
                const double me = (eta > GLM_THRESH || eta < -GLM_THRESH) ? GLM_EPS : e / (opexp * opexp);
                dv += 2.0 * (y ? log(1.0 / mu) : log(1.0 / (1.0 - mu)));   // binomial_dev_resids, y in {0, 1}
                const double z = eta + ((y ? 1.0 : 0.0) - mu) / me;
                const double w2 = me * me / (mu * (1.0 - mu));
                s00 += w2; s01 += w2 * x; s11 += w2 * x * x;
                t0 += w2 * z; t1 += w2 * z * x;
                if (d) { s02 += w2; s12 += w2 * x; t2 += w2 * z; }
            }
            S[0] = zsum(s00); S[1] = zsum(s01); S[2] = zsum(s02); S[3] = zsum(s11); S[4] = zsum(s12);
            T[0] = zsum(t0); T[1] = zsum(t1); T[2] = zsum(t2);
            return zsum(dv);
        };
        devold = pass(true);
        dev = devold;
        for (iter = 1; iter <= a.maxit; iter++) {
            // lm.fit on [X; I] with weights [w; 1 / prior.sd] (prior.mean = 0): (S + diag(1 / sd^2)) b = t
            const double a00 = S[0] + 1.0 / (sd0 * sd0), a10 = S[1], a20 = S[2];
            const double a11 = S[3] + 1.0 / (sd1 * sd1), a21 = S[4], a22 = S[2] + 1.0 / (sd2 * sd2);
            const double l00 = sqrt(a00), l10 = a10 / l00, l20 = a20 / l00;
            const double l11 = sqrt(a11 - l10 * l10), l21 = (a21 - l20 * l10) / l11;
            const double l22 = sqrt(a22 - l20 * l20 - l21 * l21);
            const double y0 = T[0] / l00, y1 = (T[1] - l10 * y0) / l11, y2 = (T[2] - l20 * y0 - l21 * y1) / l22;
            b2 = y2 / l22;
            b1 = (y1 - l21 * b2) / l11;
            b0 = (y0 - l10 * b1 - l20 * b2) / l00;
            // V = chol2inv(R) = inverse of the augmented cross-product
            const double i00 = 1.0 / l00, i11 = 1.0 / l11, i22 = 1.0 / l22;
            const double i10 = -l10 * i00 * i11, i21 = -l21 * i11 * i22, i20 = -(l20 * i00 + l21 * i10) * i22;
            v00 = i00 * i00 + i10 * i10 + i20 * i20;
            v11 = i11 * i11 + i21 * i21;
            v22 = i22 * i22;
            dev = pass(false);
            // EM step of the t prior: prior.sd^2 = ((coef - mean)^2 + V_jj dispersion + df scale^2) / (1 + df)
            sd0 = sqrt((b0 * b0 + v00 + a.df * a.scale[0] * a.scale[0]) / (1.0 + a.df));
            sd1 = sqrt((b1 * b1 + v11 + a.df * a.scale[1] * a.scale[1]) / (1.0 + a.df));
            sd2 = sqrt((b2 * b2 + v22 + a.df * a.scale[2] * a.scale[2]) / (1.0 + a.df));
            if (iter > 1 && fabs(dev - devold) / (0.1 + fabs(dev)) < a.epsilon) break;
            devold = dev;
        }
        if (iter > a.maxit) iter = a.maxit;
        if (lane == 0) {
            // summary.glm: z = coef / sqrt(diag(chol2inv(qr))) of the last fit, p = 2 pnorm(-|z|) = erfc(|z| / sqrt 2)
            const double zv = b2 / sqrt(v22);
            a.pvalue[g] = erfc(fabs(zv) * 0.70710678118654752440);
            if (a.detail) {
                double* o = a.detail + (size_t)g * 8;
                o[0] = b0; o[1] = b1; o[2] = b2;
                o[3] = sqrt(v00); o[4] = sqrt(v11); o[5] = sqrt(v22);
                o[6] = (double)iter; o[7] = dev;
            }
        }
    }
}

}  // namespace scdd
