// dense_host.h -- host-side, multi-threaded conversion between the caller's dense genes x cells matrix
// (an R matrix is column-major, numpy's is row-major) and the CSR-over-genes layout of the hot path.
// This is marshalling on either side of the GPU path, not a compute fallback: the per-gene filters of
// R/scDD.R:270-310, `y <- log(y[y>0])` (R/scDD.R:374-375), luOutlier (R/Cluster_actions.R:25) and the Zhat
// matrices (R/scDD.R:528-552).  The reference does these with apply()/for loops over genes.
#pragma once
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <mutex>
#include <thread>
#include <vector>

namespace dense {

// runs fn(lo, hi) over [0, n) split into contiguous ranges, one per host thread
template <class F>
inline void parallel_ranges(int64_t n, int64_t grain, F fn) {
    unsigned hw = std::thread::hardware_concurrency();
    int64_t T = std::max<int64_t>(1, std::min<int64_t>({(int64_t)(hw ? hw : 1), (int64_t)32, (n + grain - 1) / grain}));
    if (T == 1) { fn((int64_t)0, n); return; }
    std::vector<std::thread> th;
    int64_t per = (n + T - 1) / T;
    for (int64_t t = 0; t < T; t++) {
        int64_t lo = t * per, hi = std::min(n, lo + per);
        if (lo >= hi) break;
        th.emplace_back([=] { fn(lo, hi); });
    }
    for (auto& x : th) x.join();
}

struct View {
    const double* X;
    int64_t ngenes, ncells;
    bool col_major;
    inline double at(int64_t g, int64_t c) const { return col_major ? X[g + c * ngenes] : X[g * ncells + c]; }
};

constexpr int64_t BLK = 128;   // genes walked together in column-major order (one cache line run per cell)

// R/scDD.R:270-274 (nonzero cells per condition), :292-298 (all nonzero values identical), :262 (negatives)
inline void scan(const View& v, const int32_t* cc, int32_t* nnz0, int32_t* nnz1, uint8_t* const0, uint8_t* const1,
                 int64_t* n_negative) {
    std::atomic<int64_t> neg{0};
    parallel_ranges(v.ngenes, BLK, [&](int64_t lo, int64_t hi) {
        int64_t myneg = 0;
        int32_t n[2][BLK];
        double mn[2][BLK], mx[2][BLK];
        for (int64_t b = lo; b < hi; b += BLK) {
            int64_t w = std::min(BLK, hi - b);
            for (int k = 0; k < 2; k++)
                for (int64_t j = 0; j < w; j++) { n[k][j] = 0; mn[k][j] = INFINITY; mx[k][j] = -INFINITY; }
            auto visit = [&](int64_t j, int k, double x) {
                if (x > 0) {
                    n[k][j]++;
                    mn[k][j] = std::min(mn[k][j], x);
                    mx[k][j] = std::max(mx[k][j], x);
                } else if (x < 0) myneg++;
            };
            if (v.col_major) {
                for (int64_t c = 0; c < v.ncells; c++) {
                    const double* col = v.X + c * v.ngenes + b;
                    int k = cc[c] != 0;
                    for (int64_t j = 0; j < w; j++) visit(j, k, col[j]);
                }
            } else {
                for (int64_t j = 0; j < w; j++) {
                    const double* row = v.X + (b + j) * v.ncells;
                    for (int64_t c = 0; c < v.ncells; c++) visit(j, cc[c] != 0, row[c]);
                }
            }
            for (int64_t j = 0; j < w; j++) {
                if (nnz0) nnz0[b + j] = n[0][j];
                if (nnz1) nnz1[b + j] = n[1][j];
                if (const0) const0[b + j] = n[0][j] > 0 && mn[0][j] == mx[0][j];
                if (const1) const1[b + j] = n[1][j] > 0 && mn[1][j] == mx[1][j];
            }
        }
        neg += myneg;
    });
    if (n_negative) *n_negative = neg.load();
}

// walks the selected genes in blocks; for every kept cell of gene slot j (in cell order) calls
// emit(j_global, cell, x, position-in-CSR).  Returns false if gene_off does not match the matrix.
template <class Keep, class Emit>
inline bool walk_selected(const View& v, const int32_t* genes, int64_t nsel, const int64_t* gene_off, Keep keep, Emit emit) {
    std::atomic<bool> ok{true};
    parallel_ranges(nsel, BLK, [&](int64_t lo, int64_t hi) {
        int64_t cur[BLK], row[BLK];
        bool good = true;
        for (int64_t b = lo; b < hi; b += BLK) {
            int64_t w = std::min(BLK, hi - b);
            for (int64_t j = 0; j < w; j++) {
                cur[j] = gene_off[b + j];
                row[j] = genes ? genes[b + j] : b + j;
                if (row[j] < 0 || row[j] >= v.ngenes) { good = false; row[j] = 0; cur[j] = gene_off[b + j + 1]; }
            }
            if (v.col_major) {
                for (int64_t c = 0; c < v.ncells; c++) {
                    const double* col = v.X + c * v.ngenes;
                    for (int64_t j = 0; j < w; j++) {
                        double x = col[row[j]];
                        if (!keep(x)) continue;
                        if (cur[j] < gene_off[b + j + 1]) emit(b + j, c, x, cur[j]);
                        cur[j]++;
                    }
                }
            } else {
                for (int64_t j = 0; j < w; j++) {
                    const double* r = v.X + row[j] * v.ncells;
                    const int64_t end = gene_off[b + j + 1];
                    int64_t p = cur[j];
                    for (int64_t c = 0; c < v.ncells; c++) {
                        double x = r[c];
                        if (!keep(x)) continue;
                        if (p < end) emit(b + j, c, x, p);
                        p++;
                    }
                    cur[j] = p;
                }
            }
            for (int64_t j = 0; j < w; j++) good = good && cur[j] == gene_off[b + j + 1];
        }
        if (!good) ok = false;
    });
    return ok.load();
}

// R/scDD.R:374-375  y <- log(y[y>0]); cond0 <- condition[y>0]   (incl_zero: R/Test_KS.R:57-64 keeps zeros, no log)
inline bool to_csr(const View& v, const int32_t* cc, const int32_t* genes, int64_t nsel, bool take_log, bool incl_zero,
                   const int64_t* gene_off, double* vals, int32_t* cond01) {
    auto emit = [&](int64_t, int64_t c, double x, int64_t p) {
        vals[p] = take_log ? std::log(x) : x;
        cond01[p] = cc[c] != 0;
    };
    if (incl_zero) return walk_selected(v, genes, nsel, gene_off, [](double) { return true; }, emit);
    return walk_selected(v, genes, nsel, gene_off, [](double x) { return x > 0; }, emit);
}

// a per-cell covariate (the cellular detection rate C of R/scDD.R:457) spread over the nonzero cells of the selected
// genes in CSR order: what `C[y>0]` is for every gene (R/permutations.R:216, R/classify.dd.R:341)
inline bool cell_values_to_csr(const View& v, const double* cell_values, const int32_t* genes, int64_t nsel,
                               const int64_t* gene_off, double* out) {
    return walk_selected(v, genes, nsel, gene_off, [](double x) { return x > 0; },
                         [&](int64_t, int64_t c, double, int64_t p) { out[p] = cell_values[c]; });
}

// R/scDD.R:528-552: Z = 1, Z[X == 0] = 0, then the class labels of the fitted genes on their nonzero cells.
// which = -1: all cells (Zhat.combined); 0 / 1: only the cells of that condition (Zhat.c1 / Zhat.c2), Z then has
// that many columns.  labels are in CSR order over ALL nonzero cells of the selected genes.
inline bool labels_to_dense(const View& v, const int32_t* cc, int which, const int32_t* genes, int64_t nsel,
                            const int64_t* gene_off, const int32_t* labels, double* Z) {
    std::vector<int64_t> colmap(v.ncells);
    int64_t ncols = 0;
    for (int64_t c = 0; c < v.ncells; c++) colmap[c] = (which < 0 || (cc[c] != 0) == (which != 0)) ? ncols++ : -1;
    if (v.col_major) {
        parallel_ranges(v.ncells, 8, [&](int64_t lo, int64_t hi) {
            for (int64_t c = lo; c < hi; c++) {
                if (colmap[c] < 0) continue;
                const double* src = v.X + c * v.ngenes;
                double* dst = Z + colmap[c] * v.ngenes;
                for (int64_t g = 0; g < v.ngenes; g++) dst[g] = src[g] == 0 ? 0.0 : 1.0;
            }
        });
    } else {
        parallel_ranges(v.ngenes, 64, [&](int64_t lo, int64_t hi) {
            for (int64_t g = lo; g < hi; g++) {
                const double* src = v.X + g * v.ncells;
                double* dst = Z + g * ncols;
                for (int64_t c = 0; c < v.ncells; c++)
                    if (colmap[c] >= 0) dst[colmap[c]] = src[c] == 0 ? 0.0 : 1.0;
            }
        });
    }
    if (!labels || nsel == 0) return true;
    const int64_t ng = v.ngenes;
    const bool cm = v.col_major;
    return walk_selected(v, genes, nsel, gene_off, [](double x) { return x > 0; },
                         [&](int64_t j, int64_t c, double, int64_t p) {
                             if (colmap[c] < 0) return;
                             int64_t g = genes ? genes[j] : j;
                             Z[cm ? g + colmap[c] * ng : g * ncols + colmap[c]] = (double)labels[p];
                         });
}

// testZeroes marshalling (R/classify.dd.R:263-267): detection[c] = (number of rows with X > 0 in cell c) / nrow over
// ALL rows; for the selected rows one bit per cell (y > 0; bit c % 32 of word c / 32, W words per row, zeroed here)
// and the count of exact zeros (the `sum(y == 0) > 0` rule).
inline void zero_mask(const View& v, const int32_t* genes, int64_t nsel, int64_t W, uint32_t* mask, int32_t* nzero,
                      double* detection) {
    std::vector<int64_t> cnt(v.ncells, 0);
    if (v.col_major) {
        parallel_ranges(v.ncells, 16, [&](int64_t lo, int64_t hi) {
            for (int64_t c = lo; c < hi; c++) {
                const double* col = v.X + c * v.ngenes;
                int64_t k = 0;
                for (int64_t g = 0; g < v.ngenes; g++) k += col[g] > 0;
                cnt[c] = k;
            }
        });
    } else {
        std::mutex mu;
        parallel_ranges(v.ngenes, 64, [&](int64_t lo, int64_t hi) {
            std::vector<int64_t> mine(v.ncells, 0);
            for (int64_t g = lo; g < hi; g++) {
                const double* row = v.X + g * v.ncells;
                for (int64_t c = 0; c < v.ncells; c++) mine[c] += row[c] > 0;
            }
            std::lock_guard<std::mutex> lk(mu);
            for (int64_t c = 0; c < v.ncells; c++) cnt[c] += mine[c];
        });
    }
    for (int64_t c = 0; c < v.ncells; c++) detection[c] = (double)cnt[c] / (double)v.ngenes;
    parallel_ranges(nsel, BLK, [&](int64_t lo, int64_t hi) {
        for (int64_t b = lo; b < hi; b += BLK) {
            const int64_t w = std::min(BLK, hi - b);
            int64_t row[BLK];
            for (int64_t j = 0; j < w; j++) {
                row[j] = genes ? genes[b + j] : b + j;
                nzero[b + j] = 0;
                std::fill(mask + (b + j) * W, mask + (b + j + 1) * W, 0u);
            }
            if (v.col_major) {
                for (int64_t c = 0; c < v.ncells; c++) {
                    const double* col = v.X + c * v.ngenes;
                    const uint32_t bit = 1u << (c & 31);
                    for (int64_t j = 0; j < w; j++) {
                        const double x = col[row[j]];
                        if (x > 0) mask[(b + j) * W + (c >> 5)] |= bit;
                        else if (x == 0) nzero[b + j]++;
                    }
                }
            } else {
                for (int64_t j = 0; j < w; j++) {
                    const double* r = v.X + row[j] * v.ncells;
                    uint32_t* m = mask + (b + j) * W;
                    int32_t nz = 0;
                    for (int64_t c = 0; c < v.ncells; c++) {
                        if (r[c] > 0) m[c >> 5] |= 1u << (c & 31);
                        else if (r[c] == 0) nz++;
                    }
                    nzero[b + j] = nz;
                }
            }
        }
    });
}

// R/Cluster_actions.R:25  luOutlier: sum(table(class) >= min.size), per gene, optionally restricted to one condition
inline void cluster_counts(const int32_t* labels, const int32_t* cond01, const int64_t* gene_off, int64_t ngenes,
                           int min_size, int which, int maxlabel, int32_t* counts) {
    parallel_ranges(ngenes, 256, [&](int64_t lo, int64_t hi) {
        std::vector<int32_t> cnt(maxlabel + 1);
        for (int64_t g = lo; g < hi; g++) {
            std::fill(cnt.begin(), cnt.end(), 0);
            for (int64_t p = gene_off[g]; p < gene_off[g + 1]; p++) {
                if (which >= 0 && (cond01[p] != 0) != (which != 0)) continue;
                int l = labels[p];
                if (l >= 1 && l <= maxlabel) cnt[l]++;
            }
            int k = 0;
            for (int l = 1; l <= maxlabel; l++) k += cnt[l] >= min_size;
            counts[g] = k;
        }
    });
}

}  // namespace dense
