// scdd_lib.cu -- host side of libscdd_b200.so: the C ABI of include/scdd_b200.h on top of the kernels.
//
// One Pipeline object per device holds a gene block resident in HBM and drives
//   permgen (R sampler) -> fit (mixture fits + restricted selection + PPM score) -> combine
// per chunk of permutations.  Multi-GPU calls shard genes over devices (one host thread per device,
// no collective: every gene is independent -- SURVEY 8e).
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "kernels.cuh"
#include "fedp.cuh"
#include "zeroes.cuh"
#include "classify.cuh"
#include "dense_host.h"

using namespace scdd;

// ------------------------------------------------------------------------------------------------
// errors / device list
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static std::vector<int> g_devices;   // empty = current device
static std::mutex g_devices_mu;

struct ScddError {
    int code;
    std::string msg;
};

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            throw ScddError{SCDD_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)};    \
    } while (0)

static int fail(int code, const std::string& m) {
    g_err = m;
    return code;
}

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { if (p) cudaFree(p); }
    size_t cap = 0;
    // grows only: pipelines are cached per device, so repeated calls reuse their allocations
    void alloc(size_t count) {
        if (count <= cap) { n = count; return; }
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        n = count;
        if (count) { CK(cudaMalloc((void**)&p, count * sizeof(T))); cap = count; }
    }
};

static int device_count_quiet() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// ------------------------------------------------------------------------------------------------
// R's RNG on the host (set.seed for L'Ecuyer-CMRG and Mersenne-Twister, stream jumps)
// ------------------------------------------------------------------------------------------------
namespace {

uint32_t lcg50(uint32_t seed) {   // RNG_Init: initial scrambling
    for (int j = 0; j < 50; j++) seed = 69069u * seed + 1u;
    return seed;
}

void set_seed_lecuyer(uint32_t seed, uint32_t* s6) {
    seed = lcg50(seed);
    for (int j = 0; j < 6; j++) { seed = 69069u * seed + 1u; s6[j] = seed; }
    // FixupSeeds: components must be below the moduli (R re-randomises otherwise; practically unreachable)
    for (int j = 0; j < 3; j++) if (s6[j] >= (uint32_t)MRG_M1) s6[j] = 0;
    for (int j = 3; j < 6; j++) if (s6[j] >= (uint32_t)MRG_M2) s6[j] = 0;
}

typedef unsigned __int128 u128;
struct Mat3 { uint64_t a[3][3]; };

Mat3 matmul(const Mat3& x, const Mat3& y, uint64_t m) {
    Mat3 r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            u128 s = 0;
            for (int k = 0; k < 3; k++) s += (u128)x.a[i][k] * y.a[k][j];
            r.a[i][j] = (uint64_t)(s % m);
        }
    return r;
}

// transition matrices of the two MRG components raised to 2^e (parallel::nextRNGStream: e = 127,
// nextRNGSubStream: e = 76)
void jump_matrices(int e, Mat3& A1, Mat3& A2) {
    A1 = Mat3{{{0, 1, 0}, {0, 0, 1}, {(uint64_t)(MRG_M1 - 810728), 1403580, 0}}};
    A2 = Mat3{{{0, 1, 0}, {0, 0, 1}, {(uint64_t)(MRG_M2 - 1370589), 0, 527612}}};
    for (int i = 0; i < e; i++) { A1 = matmul(A1, A1, MRG_M1); A2 = matmul(A2, A2, MRG_M2); }
}

void apply_jump(const Mat3& A1, const Mat3& A2, const uint32_t* in, uint32_t* out) {
    uint32_t t[6];
    for (int i = 0; i < 3; i++) {
        u128 s1 = 0, s2 = 0;
        for (int k = 0; k < 3; k++) { s1 += (u128)A1.a[i][k] * in[k]; s2 += (u128)A2.a[i][k] * in[3 + k]; }
        t[i] = (uint32_t)(s1 % (uint64_t)MRG_M1);
        t[3 + i] = (uint32_t)(s2 % (uint64_t)MRG_M2);
    }
    std::memcpy(out, t, sizeof(t));
}

// coefficients {b0, b1, b2} of z[j+3] = b2 z[j+2] + b1 z[j+1] + b0 z[j] (mod m) obeyed by every 2^e-decimated
// subsequence: the characteristic polynomial of M = A^(2^e), lambda^3 = tr(M) lambda^2 - minors(M) lambda + det(M)
void charpoly3(const Mat3& M, uint64_t m, uint32_t* b) {
    auto mulm = [&](uint64_t x, uint64_t y) { return (uint64_t)((u128)x * y % m); };
    auto subm = [&](uint64_t x, uint64_t y) { return (x + m - y % m) % m; };
    const auto& a = M.a;
    uint64_t tr = (a[0][0] + a[1][1] + a[2][2]) % m;
    uint64_t minors = (subm(mulm(a[0][0], a[1][1]), mulm(a[0][1], a[1][0])) +
                       subm(mulm(a[0][0], a[2][2]), mulm(a[0][2], a[2][0])) +
                       subm(mulm(a[1][1], a[2][2]), mulm(a[1][2], a[2][1]))) % m;
    uint64_t det = 0;
    det = (det + mulm(a[0][0], subm(mulm(a[1][1], a[2][2]), mulm(a[1][2], a[2][1])))) % m;
    det = subm(det, mulm(a[0][1], subm(mulm(a[1][0], a[2][2]), mulm(a[1][2], a[2][0]))));
    det = (det + mulm(a[0][2], subm(mulm(a[1][0], a[2][1]), mulm(a[1][1], a[2][0])))) % m;
    b[0] = (uint32_t)det;
    b[1] = (uint32_t)((m - minors) % m);
    b[2] = (uint32_t)tr;
}

MrgDecim decimation_by_pow2(int log2) {
    Mat3 A1, A2;
    jump_matrices(log2, A1, A2);   // A^(2^log2)
    MrgDecim d;
    charpoly3(A1, (uint64_t)MRG_M1, d.b1);
    charpoly3(A2, (uint64_t)MRG_M2, d.b2);
    return d;
}

struct MersenneTwister {   // R's default kind
    uint32_t mt[624];
    int mti;
    void seed(uint32_t s) {
        s = lcg50(s);
        s = 69069u * s + 1u;   // i_seed[0] (the mti slot), overwritten by FixupSeeds
        for (int j = 0; j < 624; j++) { s = 69069u * s + 1u; mt[j] = s; }
        mti = 624;
    }
    SCDD_RNG_HD double unif() {
        uint32_t y;
        if (mti >= 624) {
            int kk;
            for (kk = 0; kk < 624 - 397; kk++) {
                y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
                mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            for (; kk < 623; kk++) {
                y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
                mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
            mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            mti = 0;
        }
        y = mt[mti++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return r_fixup((double)y * 2.3283064365386963e-10);
    }
};

}  // namespace

// ------------------------------------------------------------------------------------------------
// Pipeline: one gene block resident on one device
// ------------------------------------------------------------------------------------------------
struct Pipeline {
    int device = 0;
    int ngenes = 0;
    size_t nnz = 0;
    int max_n = 0, max_half = 0;
    int n_sm = 0;
    DevCfg cfg;
    DevBuf<double> d_logy, d_cdr, d_fitted, d_resid, d_jp, d_bf, d_stats, d_srt_vals;
    DevBuf<uint32_t> d_srt_idx;
    bool have_presort = false;
    DevBuf<int64_t> d_off;
    DevBuf<int32_t> d_cond, d_n1;
    DevBuf<uint32_t> d_part, d_seeds;
    DevBuf<unsigned long long> d_counter;
    DevBuf<unsigned char> d_perm, d_scratch;
    DevBuf<uint32_t> d_jit;          // stream states of the constant-subset jitter, per fit-set of a chunk
    int seed_mode = 0;               // 0: one stream per gene, 1: one stream per permutation (parallelBy = "Permutations")
    int perms_done = 0;              // permutations already processed (indexes the per-permutation seeds)
    int n_seed_rows = 0;
    DevBuf<double> d_gene_cost;      // EM work per gene measured on the first chunk ...
    DevBuf<int32_t> d_gene_order;    // ... and the resulting longest-first gene order used by later chunks
    bool have_order = false;
    uint64_t layout_hash = 0;
    std::vector<int64_t> h_off;
    bool adjusted = false;
    bool wide_idx = false;   // permutation indices need 32 bits
    int chunk_perms = 0;
    // per-launch CUDA-event pairs, recorded on the launching stream and read back in read_stats()
    struct Timed { cudaEvent_t a, b; int kind; };
    std::vector<Timed> ev_free, ev_pending;
    double fit_ms = 0, rng_ms = 0, fit_launches = 0;
    bool time_kernels = true;

    ~Pipeline() {
        for (auto& t : ev_free) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
        for (auto& t : ev_pending) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
    }

    Timed tick_begin(int kind, cudaStream_t s) {
        Timed t;
        if (!ev_free.empty()) { t = ev_free.back(); ev_free.pop_back(); }
        else { CK(cudaEventCreate(&t.a)); CK(cudaEventCreate(&t.b)); }
        t.kind = kind;
        CK(cudaEventRecord(t.a, s));
        return t;
    }
    void tick_end(Timed t, cudaStream_t s) {
        CK(cudaEventRecord(t.b, s));
        ev_pending.push_back(t);
    }

    static DevCfg to_dev(const scdd_cfg* c) {
        DevCfg d;
        d.alpha = c->alpha; d.m0 = c->m0; d.s0 = c->s0; d.a0 = c->a0; d.b0 = c->b0;
        d.min_size = c->min_size; d.restrict_ = c->restrict_; d.extra_mstep = c->extra_mstep; d.pad = 0;
        return d;
    }

    void upload(int dev, const double* logy, const int64_t* gene_off, const int32_t* cond01, int ng,
                const scdd_cfg* c, const double* cdr, cudaStream_t s) {
        device = dev;
        CK(cudaSetDevice(dev));
        {
            static std::mutex mu;
            static std::map<int, bool> done;
            std::lock_guard<std::mutex> lk(mu);
            if (!done[dev]) { CK(init_exp_table()); done[dev] = true; }
        }
        if (n_sm == 0) CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        ngenes = ng;
        cfg = to_dev(c);
        h_off.assign(gene_off, gene_off + ng + 1);
        const int64_t base = h_off[0];
        for (auto& v : h_off) v -= base;
        nnz = (size_t)h_off[ng];
        {   // the cost order survives a re-upload of the same gene layout (repeated calls on one data set)
            uint64_t h = 1469598103934665603ULL;
            for (auto v : h_off) { h ^= (uint64_t)v; h *= 1099511628211ULL; }
            if (h != layout_hash) have_order = false;
            layout_hash = h;
        }
        max_n = 0;
        for (int g = 0; g < ng; g++) max_n = std::max<int>(max_n, (int)(h_off[g + 1] - h_off[g]));
        d_logy.alloc(nnz); d_off.alloc(ng + 1); d_cond.alloc(nnz); d_part.alloc(nnz); d_n1.alloc(ng);
        d_stats.alloc(16); d_counter.alloc(1);
        CK(cudaMemcpyAsync(d_logy.p, logy + base, nnz * sizeof(double), cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(d_off.p, h_off.data(), (ng + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(d_cond.p, cond01 + base, nnz * sizeof(int32_t), cudaMemcpyHostToDevice, s));
        CK(cudaMemsetAsync(d_stats.p, 0, 16 * sizeof(double), s));
        const int wpb = 8;
        const int blocks = (ng + wpb - 1) / wpb;
        if (ng > 0) prep_partition_kernel<<<blocks, wpb * 32, 0, s>>>(d_cond.p, d_off.p, ng, d_part.p, d_n1.p);
        CK(cudaGetLastError());
        adjusted = cdr != nullptr;
        // sorted copy of every gene for the warp teams' presorted subsets (unadjusted path; env switch for A/B)
        have_presort = false;
        if (ng > 0 && !adjusted && max_n <= PRESORT_MAX_CELLS && !(std::getenv("SCDD_NO_PRESORT") && std::atoi(std::getenv("SCDD_NO_PRESORT")))) {
            d_srt_vals.alloc(nnz); d_srt_idx.alloc(nnz);
            int np = 2;
            while (np < max_n) np <<= 1;
            const size_t sm = (size_t)np * 12;
            CK(cudaFuncSetAttribute(prep_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            prep_sort_kernel<<<ng, 256, sm, s>>>(d_logy.p, d_off.p, ng, d_srt_vals.p, d_srt_idx.p);
            CK(cudaGetLastError());
            have_presort = true;
        }
        if (adjusted) {
            d_cdr.alloc(nnz); d_fitted.alloc(nnz); d_resid.alloc(nnz);
            CK(cudaMemcpyAsync(d_cdr.p, cdr + base, nnz * sizeof(double), cudaMemcpyHostToDevice, s));
            if (ng > 0) prep_lm_kernel<<<blocks, wpb * 32, 0, s>>>(d_logy.p, d_cdr.p, d_off.p, ng, d_fitted.p, d_resid.p);
            CK(cudaGetLastError());
        }
        std::vector<int32_t> h_n1(ng);
        CK(cudaMemcpyAsync(h_n1.data(), d_n1.p, ng * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        max_half = 0;
        for (int g = 0; g < ng; g++) {
            int n = (int)(h_off[g + 1] - h_off[g]);
            max_half = std::max(max_half, std::max(h_n1[g], n - h_n1[g]));
        }
        wide_idx = max_n > 65535;
    }

    size_t idx_bytes() const { return wide_idx ? 4 : 2; }

    // launches the persistent fit kernel over ngenes x nb x sets fit-sets
    DevBuf<double> d_xs_scratch;
    void launch_fit(cudaStream_t s, FitArgs a, int nmax_points, bool force_cta = false) {
        int np = 2;
        while (np < nmax_points) np <<= 1;
        const long long total = (long long)a.ngenes * a.nb * a.sets;
        if (total == 0) return;
        const size_t smem_cap = 227 * 1024 - FIT_TAB_BYTES;
        a.np_max = np;
        a.work_counter = d_counter.p;
        a.stats = d_stats.p;
        a.cfg = cfg;
        a.xs_global = nullptr;
        a.sort_tile = 0;
        int grid, threads, tw = 1;
        size_t smem;
        const bool warp_team = nmax_points <= FIT_WARP_TEAM_MAX_POINTS && !force_cta;
        if (warp_team) {
            const size_t per_warp = fit_smem_per_warp(np);
            int warps = (int)std::min<size_t>(FIT_MAX_WARPS, smem_cap / per_warp);
            smem = per_warp * warps;             // dynamic part; the exp table is static shared memory
            threads = warps * 32;
            grid = (int)std::min<long long>(n_sm, (total + warps - 1) / warps);
            if (a.iter_budget) CK(cudaFuncSetAttribute(fit_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            else CK(cudaFuncSetAttribute(fit_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        } else {
            // several warps per fit-set.  Default: the whole CTA (16 warps) per fit-set -- the slabs of 148 concurrent
            // fit-sets of config-5 size (352 KB each) stay L2-resident, and the two-pass observed fits are bound by
            // their longest single fit-set.  4-warp teams, four to a CTA, only for bulk work on mid-size vectors whose
            // slabs all fit shared memory (measured in round 2, profiles/sub_cta_teams_r02.txt: with 4-warp teams
            // config 5 drops from 0.377 to 0.250 of peak -- 592 slabs no longer fit L2 -- and the observed-fit tail grows)
            const bool bulk_mid = np <= 4096 && total > 8LL * n_sm && !a.only_deferred &&
                                  fit_team_region(FIT_SUB_WARPS, np, true) * (FIT_MAX_WARPS / FIT_SUB_WARPS) <= smem_cap;
            tw = bulk_mid ? FIT_SUB_WARPS : FIT_CTA_WARPS;
            const int teams = FIT_MAX_WARPS / tw;
            threads = FIT_MAX_WARPS * 32;
            grid = (int)std::min<long long>(n_sm, (total + teams - 1) / teams);
            if (fit_team_region(tw, np, true) * teams <= smem_cap) {
                smem = fit_team_region(tw, np, true) * teams;
            } else {
                // slab per team in global memory (L2); the sort goes through the largest power-of-two tile that fits
                // the team's share of shared memory
                const size_t head = fit_team_region(tw, np, false);
                int tile = 1;
                while ((size_t)tile * 2 * sizeof(double) + head <= smem_cap / teams) tile <<= 1;
                a.sort_tile = (tile >= 1024 && tile < np) ? tile : 0;
                // the E-step stages such a slab through the same tile (2 buffers x >= 1 point per thread)
                if (a.sort_tile < 2 * tw * 32) throw ScddError{SCDD_ERR_TOO_LARGE, "no room for the staging tile of a global-memory slab"};
                smem = (head + (size_t)a.sort_tile * sizeof(double)) * teams;
                const size_t need = (size_t)grid * teams * np;
                if (d_xs_scratch.n < need) d_xs_scratch.alloc(need);
                a.xs_global = d_xs_scratch.p;
            }
            if (tw == FIT_SUB_WARPS) CK(cudaFuncSetAttribute(fit_kernel<FIT_SUB_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            else CK(cudaFuncSetAttribute(fit_kernel<FIT_CTA_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
        CK(cudaMemsetAsync(d_counter.p, 0, sizeof(unsigned long long), s));
        Timed t;
        if (time_kernels) t = tick_begin(0, s);
        if (warp_team && a.iter_budget) fit_kernel<1, true><<<grid, threads, smem, s>>>(a);
        else if (warp_team) fit_kernel<1><<<grid, threads, smem, s>>>(a);
        else if (tw == FIT_SUB_WARPS) fit_kernel<FIT_SUB_WARPS><<<grid, threads, smem, s>>>(a);
        else fit_kernel<FIT_CTA_WARPS><<<grid, threads, smem, s>>>(a);
        CK(cudaGetLastError());
        if (time_kernels) tick_end(t, s);
        fit_launches += 1;
    }

    void harvest_times() {   // call after the stream is synchronised
        for (auto& t : ev_pending) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess) (t.kind == 0 ? fit_ms : rng_ms) += ms;
            else cudaGetLastError();
            ev_free.push_back(t);
        }
        ev_pending.clear();
    }

    FitArgs base_args() const {
        FitArgs a;
        std::memset(&a, 0, sizeof(a));
        a.vals = adjusted ? d_resid.p : d_logy.p;
        a.fitted = adjusted ? d_fitted.p : nullptr;
        a.gene_off = d_off.p;
        a.part = d_part.p;
        a.n1 = d_n1.p;
        a.ngenes = ngenes;
        a.gene_order = have_order ? d_gene_order.p : nullptr;
        a.srt_vals = have_presort ? d_srt_vals.p : nullptr;
        a.srt_idx = have_presort ? d_srt_idx.p : nullptr;
        a.gene_cost = nullptr;
        a.iter_budget = 0; a.only_deferred = 0; a.deferred = nullptr;
        return a;
    }

    // Longest-processing-time-first scheduling of the persistent fit kernel.  The EM work of a gene's permuted
    // fit-sets is a property of the gene's pooled distribution, so the work measured on one chunk predicts the
    // next: genes are then started in decreasing cost order and the kernel's tail consists of cheap fit-sets.
    void begin_cost_measurement(cudaStream_t s) {
        d_gene_cost.alloc(ngenes);
        CK(cudaMemsetAsync(d_gene_cost.p, 0, (size_t)ngenes * sizeof(double), s));
    }
    void finish_cost_measurement(cudaStream_t s) {
        std::vector<double> cost(ngenes);
        CK(cudaMemcpyAsync(cost.data(), d_gene_cost.p, (size_t)ngenes * sizeof(double), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        std::vector<int32_t> order(ngenes);
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return cost[x] > cost[y]; });
        d_gene_order.alloc(ngenes);
        CK(cudaMemcpyAsync(d_gene_order.p, order.data(), (size_t)ngenes * sizeof(int32_t), cudaMemcpyHostToDevice, s));
        CK(cudaStreamSynchronize(s));
        have_order = true;
    }

    // device buffers for `bc` permutations per gene drawn on the device
    // seed_rows: ngenes (mode 0) or the total number of permutations B (mode 1)
    void alloc_perm_chunk(int bc, const uint32_t* seed6, cudaStream_t s, int mode = 0, int seed_rows = -1) {
        chunk_perms = bc;
        seed_mode = mode;
        perms_done = 0;
        n_seed_rows = seed_rows < 0 ? ngenes : seed_rows;
        d_perm.alloc(nnz * (size_t)bc * idx_bytes());
        d_scratch.alloc(nnz * idx_bytes());
        d_seeds.alloc((size_t)n_seed_rows * 6);
        CK(cudaMemcpyAsync(d_seeds.p, seed6, (size_t)n_seed_rows * 6 * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
        const int sets = adjusted ? 3 : 2;
        d_jp.alloc((size_t)ngenes * bc * sets);
        d_jit.alloc((size_t)ngenes * bc * sets * 6);
    }

    // GS lanes per gene for permgen_group_kernel: the smallest group that still leaves several warps per SM
    // dynamic shared memory a permutation-kernel CTA may ask for (B200: 227 KB per CTA).  224 KB lets TWO 50,000-cell
    // genes (100.6 KB each with their candidate blocks) share an SM: the kernel is latency-bound, so config 5's
    // permutation time halves against the 200 KB used before
    static constexpr size_t PERMGEN_SMEM_CAP = 224 * 1024;
    int pick_group_size() const {
        const size_t cap = PERMGEN_SMEM_CAP;
        const int ib = (int)idx_bytes();
        auto fits = [&](int gs, size_t min_warps) {
            return cap / (permgen_smem_per_group(max_n, ib, gs) * (32 / gs)) >= min_warps;
        };
        if (const char* e = std::getenv("SCDD_PERMGEN_GS")) {   // A/B switch; ignored when the genes do not fit
            const int v = std::atoi(e);
            if ((v == 4 || v == 8 || v == 32) && fits(v, 1)) return v;
        }
        // measured on B200 (profiles/permgen_groups_r02.txt): 8 lanes per gene is the fastest of the three where it fits
        // -- unless the shard is so small that 8-lane groups leave the SMs with ~4 warps each (a 2,500-gene shard of
        // an 8-GPU job: 32 lanes per gene is 17 % faster there, profiles/permgen_fewgenes_r02.txt)
        if (ngenes <= 24 * n_sm && fits(32, 4)) return 32;
        for (int gs : {8})
            if (fits(gs, 4)) return gs;
        return 32;
    }

    template <class IdxT, int GS>
    void launch_permgen_gs(cudaStream_t s, PermgenArgs a) {
        static const MrgDecim dc = decimation_by_pow2(GS == 32 ? 5 : (GS == 8 ? 3 : 2));
        a.dc = dc;
        const size_t cap = PERMGEN_SMEM_CAP;
        const size_t per_warp = permgen_smem_per_group(max_n, (int)sizeof(IdxT), GS) * (32 / GS);
        const int gpw = 32 / GS;
        size_t wpb_max = 4;
        if (const char* e = std::getenv("SCDD_PERMGEN_WPB")) wpb_max = (size_t)std::max(1, std::atoi(e));   // A/B switch
        int wpb = (int)std::max<size_t>(1, std::min<size_t>(wpb_max, cap / per_warp));
        wpb = std::max(1, std::min(wpb, (ngenes + gpw - 1) / gpw));
        const size_t smem = per_warp * wpb;
        const int blocks_needed = (ngenes + gpw * wpb - 1) / (gpw * wpb);
        const int grid = std::min(blocks_needed, n_sm * std::max(1, std::min(16, (int)(cap / smem))));
        // whole-group consumption of the stream where long genes leave few groups per SM; the leader's serial loop
        // elsewhere (kernels.cuh, "BATCH": slower whenever enough genes share an SM to hide the leader's chain)
        bool batch = GS == 32 && PERMGEN_SMEM_CAP / per_warp < 8;   // fewer than 8 genes per SM (> ~14,000 cells)
        if (const char* e = std::getenv("SCDD_PERMGEN_BATCH")) batch = std::atoi(e) != 0;   // A/B switch
        if (batch) {
            CK(cudaFuncSetAttribute(permgen_group_kernel<IdxT, GS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            permgen_group_kernel<IdxT, GS, true><<<grid, wpb * 32, smem, s>>>(a);
        } else {
            CK(cudaFuncSetAttribute(permgen_group_kernel<IdxT, GS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            permgen_group_kernel<IdxT, GS, false><<<grid, wpb * 32, smem, s>>>(a);
        }
    }
    template <class IdxT>
    void launch_permgen(cudaStream_t s, const PermgenArgs& a, int gs) {
        if (gs == 4) launch_permgen_gs<IdxT, 4>(s, a);
        else if (gs == 8) launch_permgen_gs<IdxT, 8>(s, a);
        else launch_permgen_gs<IdxT, 32>(s, a);
    }

    // permgen + fit + combine for nb permutations; bf_out[g * ld + col0 + b]; jp_out (nullable, device):
    // the chunk's jointPosterior terms go to jp_out[(g * ld + col0 + b) * sets + w]
    void run_chunk_device_perms(cudaStream_t s, int nb, double* bf_out, long long ld, int col0, double* jp_out = nullptr) {
        if (ngenes == 0 || nb == 0) return;
        const int sets = adjusted ? 3 : 2;
        Timed tr;
        if (time_kernels) tr = tick_begin(1, s);
        bool with_jit = false;
        {
            const int ib = (int)idx_bytes();
            const size_t cap = PERMGEN_SMEM_CAP;
            if (permgen_smem_per_group(max_n, ib, 32) <= cap) {
                with_jit = true;
                CK(cudaMemsetAsync(d_jit.p, 0xff, (size_t)ngenes * nb * sets * 6 * sizeof(uint32_t), s));
                PermgenArgs pa;
                std::memset(&pa, 0, sizeof(pa));
                pa.off = d_off.p; pa.ngenes = ngenes; pa.seeds = d_seeds.p; pa.seed_mode = seed_mode; pa.perm_b0 = perms_done;
                pa.perm_out = d_perm.p; pa.nb = nb; pa.nmax = max_n;
                pa.vals = adjusted ? d_resid.p : d_logy.p; pa.fitted = adjusted ? d_fitted.p : nullptr;
                pa.part = d_part.p; pa.n1 = d_n1.p; pa.sets = sets; pa.jit_state = d_jit.p;
                const int gs = pick_group_size();
                if (wide_idx) launch_permgen<uint32_t>(s, pa, gs);
                else launch_permgen<uint16_t>(s, pa, gs);
            } else {
                // genes too long for the shared-memory Fisher-Yates array: thread-per-gene kernel, global scratch
                // (per-gene streams only; constant subsets are not jittered on this path and fail as "constant")
                if (seed_mode != 0) throw ScddError{SCDD_ERR_TOO_LARGE, "per-permutation streams need genes that fit shared memory"};
                const int tpb = 64;
                if (wide_idx)
                    permgen_kernel<uint32_t><<<(ngenes + tpb - 1) / tpb, tpb, 0, s>>>(d_off.p, ngenes, d_seeds.p,
                        (uint32_t*)d_scratch.p, (uint32_t*)d_perm.p, nb);
                else
                    permgen_kernel<uint16_t><<<(ngenes + tpb - 1) / tpb, tpb, 0, s>>>(d_off.p, ngenes, d_seeds.p,
                        (uint16_t*)d_scratch.p, (uint16_t*)d_perm.p, nb);
            }
        }
        CK(cudaGetLastError());
        if (time_kernels) tick_end(tr, s);
        FitArgs a = base_args();
        a.perm = d_perm.p;
        a.perm_bstride = nb;
        a.perm_kind = wide_idx ? PERM_U32 : PERM_U16;
        a.perm_b0 = 0;
        a.nb = nb;
        a.sets = sets;
        a.jp = d_jp.p;
        a.jit_state = with_jit ? d_jit.p : nullptr;
        const bool measure = !have_order && ngenes >= 2 * n_sm;   // worth ordering only when genes outnumber the SMs
        if (measure) { begin_cost_measurement(s); a.gene_cost = d_gene_cost.p; }
        launch_fit(s, a, adjusted ? max_n : max_half);
        if (measure) finish_cost_measurement(s);
        const long long units = (long long)ngenes * nb;
        combine_kernel<<<(int)((units + 255) / 256), 256, 0, s>>>(d_jp.p, ngenes, nb, a.sets, bf_out, ld, col0, jp_out);
        CK(cudaGetLastError());
        perms_done += nb;
    }

    void read_stats(scdd_stats* st, cudaStream_t s) {
        double h[16];
        CK(cudaMemcpyAsync(h, d_stats.p, sizeof(h), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        harvest_times();
        st->triples += h[0]; st->point_iters += h[1]; st->em_iters += h[2]; st->iter_cap_hits += h[3];
        st->tol_near += h[4]; st->bic_near += h[5]; st->const_sets += h[6]; st->failed_sets += h[7];
        st->fit_sets += h[8];
        st->debug_violations += h[9];
        st->fit_kernel_ms += fit_ms; st->rng_kernel_ms += rng_ms; st->fit_kernel_launches += fit_launches;
    }
};

// ------------------------------------------------------------------------------------------------
// gene sharding over devices (LPT greedy on the number of nonzero cells; contiguous ranges are kept
// simple here: genes are dealt to the least-loaded device in decreasing size, then each device's
// genes are processed as gathered sub-blocks)
// ------------------------------------------------------------------------------------------------
struct Shard {
    int device;
    std::vector<int> genes;           // original gene indices
    std::vector<int64_t> off;         // local CSR
    std::vector<double> logy, cdr;
    std::vector<int32_t> cond;
};

static std::vector<int> active_devices() {
    {
        std::lock_guard<std::mutex> lk(g_devices_mu);
        if (!g_devices.empty()) return g_devices;
    }
    int cur = 0;
    if (cudaGetDevice(&cur) != cudaSuccess) { cudaGetLastError(); cur = 0; }
    return std::vector<int>{cur};
}

static int check_common(const void* a, const void* b, const void* c, int ngenes, const scdd_cfg* cfg) {
    if (ngenes < 0 || (ngenes > 0 && (!a || !b || !c)) || !cfg) return fail(SCDD_ERR_BAD_ARG, "null pointer or negative gene count");
    if (device_count_quiet() < 1) return fail(SCDD_ERR_NO_DEVICE, "no CUDA device available: libscdd_b200 has no CPU fallback");
    return SCDD_OK;
}

template <class F>
static int guarded(F&& f) {
    try {
        return f();
    } catch (const ScddError& e) {
        cudaGetLastError();
        return fail(e.code, e.msg);
    } catch (const std::exception& e) {
        return fail(SCDD_ERR_CUDA, e.what());
    }
}

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int32_t scdd_abi_version(void) { return SCDD_ABI_VERSION; }
const char* scdd_last_error(void) { return g_err.c_str(); }
int32_t scdd_device_count(void) { return device_count_quiet(); }

int32_t scdd_set_devices(const int32_t* devices, int32_t n) {
    if (n < 0 || (n > 0 && !devices)) return fail(SCDD_ERR_BAD_ARG, "bad device list");
    int cnt = device_count_quiet();
    std::vector<int> v;
    for (int i = 0; i < n; i++) {
        if (devices[i] < 0 || devices[i] >= cnt) return fail(SCDD_ERR_NO_DEVICE, "device index out of range");
        v.push_back(devices[i]);
    }
    std::lock_guard<std::mutex> lk(g_devices_mu);
    g_devices = v;
    return SCDD_OK;
}

void scdd_default_cfg(scdd_cfg* c) {
    c->alpha = 0.10; c->m0 = 0.0; c->s0 = 0.01; c->a0 = 0.01; c->b0 = 0.01;   // R/scDD.R:218
    c->min_size = 3; c->restrict_ = 1; c->extra_mstep = 1; c->ks_old_pvalue = 0;
}

static void merge_stats(scdd_stats& dst, const scdd_stats& src);

// ---------------------------------------------------------------- observed fits
// EM iterations one warp spends on a single mixture before the fit-set is handed to a CTA (observed fits)
static const int OBSERVED_ITER_BUDGET = 500;

static void fit_observed_one(int dev, const double* logy, const int64_t* off, const int32_t* cond, int ng,
                             const scdd_cfg* cfg, scdd_fit* fits, int32_t* cls_oa, int32_t* cls_c,
                             double* bf, double* den, scdd_stats* st) {
    if (ng == 0) return;
    Pipeline P;
    cudaStream_t s;
    CK(cudaSetDevice(dev));
    CK(cudaStreamCreate(&s));
    try {
        P.upload(dev, logy, off, cond, ng, cfg, nullptr, s);
        DevBuf<scdd_fit> d_fits;
        DevBuf<int32_t> d_oa, d_c;
        d_fits.alloc(3 * (size_t)ng); d_oa.alloc(P.nnz); d_c.alloc(P.nnz);
        P.d_jp.alloc(3 * (size_t)ng);
        // no measured costs exist for a first look at the data: start the genes with the most cells first, so the
        // persistent kernel's tail is made of small fit-sets
        if (ng >= 2 * P.n_sm) {
            std::vector<int32_t> order(ng);
            std::iota(order.begin(), order.end(), 0);
            std::stable_sort(order.begin(), order.end(),
                             [&](int32_t x, int32_t y) { return P.h_off[x + 1] - P.h_off[x] > P.h_off[y + 1] - P.h_off[y]; });
            P.d_gene_order.alloc(ng);
            CK(cudaMemcpyAsync(P.d_gene_order.p, order.data(), (size_t)ng * sizeof(int32_t), cudaMemcpyHostToDevice, s));
            CK(cudaStreamSynchronize(s));
            P.have_order = true;
        }
        FitArgs a = P.base_args();
        a.nb = 1; a.sets = 3; a.jp = P.d_jp.p;
        a.fits = d_fits.p; a.cls_oa = d_oa.p; a.cls_c = d_c.p;
        if (P.max_n <= FIT_WARP_TEAM_MAX_POINTS) {
            // pass 1: one warp per fit-set with an EM iteration budget; pass 2: the ~1 % of fit-sets that exceeded it,
            // one CTA each (a launch would otherwise last as long as its slowest fit-set on a single warp)
            DevBuf<unsigned char> d_deferred;
            d_deferred.alloc(3 * (size_t)ng);
            CK(cudaMemsetAsync(d_deferred.p, 0, 3 * (size_t)ng, s));
            a.deferred = d_deferred.p;
            a.iter_budget = OBSERVED_ITER_BUDGET;
            P.launch_fit(s, a, P.max_n);
            a.iter_budget = 0; a.only_deferred = 1;
            P.launch_fit(s, a, P.max_n, /*force_cta=*/true);
            CK(cudaStreamSynchronize(s));   // d_deferred goes out of scope
        } else {
            P.launch_fit(s, a, P.max_n);
        }
        std::vector<double> jp(3 * (size_t)ng);
        const int64_t base = off[0];
        CK(cudaMemcpyAsync(fits, d_fits.p, 3 * (size_t)ng * sizeof(scdd_fit), cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(cls_oa + base, d_oa.p, P.nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(cls_c + base, d_c.p, P.nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(jp.data(), P.d_jp.p, jp.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
        P.read_stats(st, s);
        for (int g = 0; g < ng; g++) {
            if (bf) bf[g] = jp[3 * (size_t)g] + jp[3 * (size_t)g + 1];   // R/scDD.R:381-382
            if (den) den[g] = jp[3 * (size_t)g + 2];                     // :383
        }
        st->h2d_bytes += (double)P.nnz * 12 + (ng + 1) * 8.0;
        st->d2h_bytes += 3.0 * ng * sizeof(scdd_fit) + P.nnz * 8.0 + jp.size() * 8.0;
    } catch (...) {
        cudaStreamDestroy(s);
        throw;
    }
    CK(cudaStreamDestroy(s));
}

int32_t scdd_fit_observed(const double* logy, const int64_t* gene_off, const int32_t* cond01, int32_t ngenes,
                          const scdd_cfg* cfg, scdd_fit* fits, int32_t* cls_oa, int32_t* cls_c,
                          double* bf, double* den, scdd_stats* stats) {
    int rc = check_common(logy, gene_off, cond01, ngenes, cfg);
    if (rc) return rc;
    if (ngenes > 0 && (!fits || !cls_oa || !cls_c)) return fail(SCDD_ERR_BAD_ARG, "null output");
    scdd_stats st;
    std::memset(&st, 0, sizeof(st));
    return guarded([&]() -> int {
        // 3 fit-sets per gene.  With several devices the genes are cut into contiguous ranges of equal cell counts (no
        // gather: every range is a sub-block of the caller's arrays) and fitted concurrently, one host thread per device
        std::vector<int> devs = active_devices();
        if (devs.size() == 1 || ngenes < 64 * (int)devs.size()) {
            fit_observed_one(devs[0], logy, gene_off, cond01, ngenes, cfg, fits, cls_oa, cls_c, bf, den, &st);
        } else {
            const size_t nd = devs.size();
            std::vector<int> cut(nd + 1, ngenes);
            cut[0] = 0;
            const int64_t base = gene_off[0], total = gene_off[ngenes] - base;
            for (size_t d = 1; d < nd; d++) {
                const int64_t want = base + total * (int64_t)d / (int64_t)nd;
                cut[d] = (int)(std::lower_bound(gene_off, gene_off + ngenes + 1, want) - gene_off);
                cut[d] = std::max(cut[d], cut[d - 1]);
            }
            std::vector<scdd_stats> sts(nd);
            std::vector<ScddError> errs(nd, ScddError{0, ""});
            std::vector<std::thread> th;
            for (size_t d = 0; d < nd; d++) {
                th.emplace_back([&, d]() {
                    std::memset(&sts[d], 0, sizeof(scdd_stats));
                    const int g0 = cut[d], n = cut[d + 1] - cut[d];
                    try {
                        fit_observed_one(devs[d], logy, gene_off + g0, cond01, n, cfg, fits + 3 * (size_t)g0, cls_oa, cls_c,
                                         bf ? bf + g0 : nullptr, den ? den + g0 : nullptr, &sts[d]);
                    } catch (const ScddError& e) {
                        errs[d] = e;
                    } catch (const std::exception& e) {
                        errs[d] = ScddError{SCDD_ERR_CUDA, e.what()};
                    }
                });
            }
            for (auto& t : th) t.join();
            for (auto& e : errs) if (e.code) throw e;
            double ms = 0.0;
            for (auto& s1 : sts) { merge_stats(st, s1); ms = std::max(ms, s1.fit_kernel_ms); }
            st.fit_kernel_ms = ms;      // concurrent devices: the slowest one
        }
        if (stats) *stats = st;
        if (st.failed_sets > 0) return fail(SCDD_ERR_FIT_FAILED, "some fit-sets had no valid model (constant values?)");
        return SCDD_OK;
    });
}

// ---------------------------------------------------------------- permutations
// One cached pipeline (device buffers + stream + pinned staging) per device: a call reuses the previous call's
// allocations, so the end-to-end cost of scdd_perm_bf is copies + kernels, not cudaMalloc/cudaFree of gigabytes.
// scdd_release_cache() frees them.
struct PinnedBuf {
    void* p = nullptr;
    size_t cap = 0;
    void* get(size_t bytes) {
        if (bytes <= cap) return p;
        if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
        CK(cudaHostAlloc(&p, bytes, cudaHostAllocDefault));
        cap = bytes;
        return p;
    }
    void release() { if (p) { cudaFreeHost(p); p = nullptr; cap = 0; } }
};
struct CachedPipeline {
    std::mutex mu;
    std::unique_ptr<Pipeline> P;
    cudaStream_t s = nullptr;
};
// pinned staging of one gene shard of a multi-device call (logy, cond, cdr, seeds, results); kept per shard SLOT, not
// per device, so that a device may be listed twice (two shards on one GPU: how the sharded path is tested on one GPU)
struct ShardStage { PinnedBuf stage[5]; };
static std::mutex g_stage_mu;
static std::vector<std::unique_ptr<ShardStage>> g_stage;
static ShardStage& shard_stage(size_t slot) {
    std::lock_guard<std::mutex> lk(g_stage_mu);
    while (g_stage.size() <= slot) g_stage.emplace_back(new ShardStage());
    return *g_stage[slot];
}
static std::mutex g_cache_mu;
static std::map<int, std::unique_ptr<CachedPipeline>> g_cache;
static CachedPipeline& cached_pipeline(int dev) {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    auto& e = g_cache[dev];
    if (!e) e.reset(new CachedPipeline());
    return *e;
}

typedef int32_t (*scdd_interrupt_fn)(void*);
struct PermCall {
    int B = 0;              // permutations of the whole call = row stride of bf_perm / jp_terms
    int b0 = 0, nbp = 0;    // this invocation computes permutations [b0, b0 + nbp)
    const int32_t* perm_idx = nullptr;
    const uint32_t* seed6 = nullptr;
    int seed_mode = 0;
    const double* cdr = nullptr;
    const scdd_cfg* cfg = nullptr;
    double* bf_perm = nullptr;      // [ng * B]
    double* jp_terms = nullptr;     // nullable [ng * B * sets]
    uint32_t* seed6_out = nullptr;  // nullable [ng * 6] (seed_mode 0)
    double* gene_cost = nullptr;    // nullable [ng]: EM work per gene measured on the first chunk
};

// Interruption: opts->interrupt is polled between permutation chunks, always on the CALLING thread (R's API is
// main-thread only: the glue's callback wraps R_CheckUserInterrupt in R_ToplevelExec).  Worker threads of a
// multi-device call only read the shared `cancel` flag.
static void perm_bf_one(int dev, const double* logy, const int64_t* off, const int32_t* cond, int ng,
                        const PermCall& pc, scdd_stats* st, scdd_interrupt_fn intr, void* intr_arg,
                        std::atomic<int>* cancel = nullptr) {
    const int B = pc.B;
    auto cancelled = [&]() {
        if (cancel && cancel->load()) return true;
        if (intr && intr(intr_arg)) { if (cancel) cancel->store(1); return true; }
        return false;
    };
    const bool polls = intr != nullptr || cancel != nullptr;
    if (ng == 0 || pc.nbp == 0) return;
    CachedPipeline& C = cached_pipeline(dev);
    std::lock_guard<std::mutex> lk(C.mu);
    CK(cudaSetDevice(dev));
    if (!C.P) { C.P.reset(new Pipeline()); CK(cudaStreamCreate(&C.s)); }
    Pipeline& P = *C.P;
    P.fit_ms = P.rng_ms = P.fit_launches = 0;
    cudaStream_t s = C.s;
    try {
        cudaEvent_t t0, t1;
        CK(cudaEventCreate(&t0)); CK(cudaEventCreate(&t1));
        CK(cudaEventRecord(t0, s));
        P.upload(dev, logy, off, cond, ng, pc.cfg, pc.cdr, s);
        const int sets = P.adjusted ? 3 : 2;
        const int nbp = pc.nbp;
        P.d_bf.alloc((size_t)ng * nbp);
        DevBuf<double> d_terms;
        if (pc.jp_terms) d_terms.alloc((size_t)ng * nbp * sets);
        const int64_t base = off[0];
        double h2d = (double)P.nnz * 12 + (ng + 1) * 8.0 + (pc.cdr ? P.nnz * 8.0 : 0.0);
        bool interrupted = false;
        if (pc.perm_idx) {
            // host-drawn permutations (the R shim's own sample() calls): whole table resident
            const size_t cnt = P.nnz * (size_t)B;
            DevBuf<int32_t> d_pi;
            d_pi.alloc(cnt);
            CK(cudaMemcpyAsync(d_pi.p, pc.perm_idx + (size_t)base * B, cnt * sizeof(int32_t), cudaMemcpyHostToDevice, s));
            h2d += cnt * 4.0;
            // jp chunked over permutations to bound memory
            int bc = (int)std::max<long long>(1, std::min<long long>(nbp, (1LL << 27) / std::max(1, ng * sets)));
            P.d_jp.alloc((size_t)ng * bc * sets);
            for (int c0 = 0; c0 < nbp && !interrupted; c0 += bc) {
                int nb = std::min(bc, nbp - c0);
                FitArgs a = P.base_args();
                a.perm = d_pi.p; a.perm_bstride = B; a.perm_kind = PERM_I32_1BASED; a.perm_b0 = pc.b0 + c0;
                a.nb = nb; a.sets = sets; a.jp = P.d_jp.p;
                P.launch_fit(s, a, P.adjusted ? P.max_n : P.max_half);
                const long long units = (long long)ng * nb;
                combine_kernel<<<(int)((units + 255) / 256), 256, 0, s>>>(P.d_jp.p, ng, nb, sets, P.d_bf.p, nbp, c0, d_terms.p);
                CK(cudaGetLastError());
                if (polls) { CK(cudaStreamSynchronize(s)); if (cancelled()) interrupted = true; }
            }
            CK(cudaStreamSynchronize(s));
        } else {
            // device-drawn permutations, chunked so the index buffer stays below ~2 GiB
            long long per_perm = (long long)P.nnz * (long long)P.idx_bytes();
            int bc = (int)std::max<long long>(1, std::min<long long>(nbp, (2LL << 30) / std::max<long long>(1, per_perm)));
            bc = (int)std::min<long long>(bc, std::max<long long>(1, (1LL << 27) / std::max(1, ng * sets)));
            if (polls || pc.gene_cost) bc = std::min(bc, 16);   // interruptible calls / cost probes: chunks of <= 16 permutations
            P.alloc_perm_chunk(bc, pc.seed6, s, pc.seed_mode, pc.seed_mode ? B : ng);
            P.perms_done = pc.seed_mode ? pc.b0 : 0;
            h2d += (pc.seed_mode ? B : ng) * 24.0;
            for (int c0 = 0; c0 < nbp && !interrupted; c0 += bc) {
                int nb = std::min(bc, nbp - c0);
                P.run_chunk_device_perms(s, nb, P.d_bf.p, nbp, c0, d_terms.p);
                if (polls) { CK(cudaStreamSynchronize(s)); if (cancelled()) interrupted = true; }
            }
            if (pc.seed6_out && pc.seed_mode == 0)
                CK(cudaMemcpyAsync(pc.seed6_out, P.d_seeds.p, (size_t)ng * 6 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
            if (pc.gene_cost) {
                if (P.have_order && P.d_gene_cost.n >= (size_t)ng)
                    CK(cudaMemcpyAsync(pc.gene_cost, P.d_gene_cost.p, (size_t)ng * sizeof(double), cudaMemcpyDeviceToHost, s));
                else
                    for (int g = 0; g < ng; g++) pc.gene_cost[g] = (double)(off[g + 1] - off[g]);
            }
        }
        // results: rows of nbp values into the caller's rows of B values
        CK(cudaMemcpy2DAsync(pc.bf_perm + pc.b0, (size_t)B * sizeof(double), P.d_bf.p, (size_t)nbp * sizeof(double),
                             (size_t)nbp * sizeof(double), (size_t)ng, cudaMemcpyDeviceToHost, s));
        if (pc.jp_terms)
            CK(cudaMemcpy2DAsync(pc.jp_terms + (size_t)pc.b0 * sets, (size_t)B * sets * sizeof(double), d_terms.p,
                                 (size_t)nbp * sets * sizeof(double), (size_t)nbp * sets * sizeof(double), (size_t)ng,
                                 cudaMemcpyDeviceToHost, s));
        CK(cudaEventRecord(t1, s));
        P.read_stats(st, s);
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, t0, t1));
        st->total_ms += ms;
        st->h2d_bytes += h2d;
        st->d2h_bytes += (double)ng * nbp * 8.0 * (pc.jp_terms ? 1 + sets : 1);
        cudaEventDestroy(t0); cudaEventDestroy(t1);
        if (interrupted) throw ScddError{SCDD_ERR_INTERRUPTED, "interrupted by the caller between permutation chunks"};
    } catch (...) {
        cudaStreamSynchronize(s);
        cudaGetLastError();
        throw;
    }
}

// adds the counters of one shard / phase; device times of concurrent shards overlap, so total_ms is handled by the caller
static void merge_stats(scdd_stats& dst, const scdd_stats& src) {
    const double* a = reinterpret_cast<const double*>(&src);
    double* d = reinterpret_cast<double*>(&dst);
    const double keep = dst.total_ms;
    for (size_t k = 0; k < sizeof(scdd_stats) / sizeof(double); k++) d[k] += a[k];
    dst.total_ms = keep;
}

// Multi-device call: each device's genes are gathered into PINNED staging (one pass over the caller's pageable arrays,
// then full-speed asynchronous H2D), the shards run concurrently on one host thread per device, results are scattered
// back.  Two phases for device-drawn permutations: a short first phase on shards balanced by the number of nonzero
// cells measures the EM work of every gene (it is a property of the gene, not of the permutation); the remaining
// permutations run on shards re-dealt longest-first by that MEASURED cost, every gene's stream continuing from the
// state the first phase left (balance by cell counts alone leaves the slowest of eight devices ~6 % behind).
static void perm_bf_sharded(const std::vector<int>& devs, const double* logy, const int64_t* gene_off,
                            const int32_t* cond01, int ngenes, const PermCall& call, scdd_stats& st,
                            scdd_interrupt_fn intr, void* intr_arg) {
    const int B = call.B;
    const int sets = call.cdr ? 3 : 2;
    const bool device_perms = call.perm_idx == nullptr;
    std::vector<uint32_t> seeds_cur;          // per-gene stream states carried between the phases (seed_mode 0)
    if (device_perms && call.seed_mode == 0) seeds_cur.assign(call.seed6, call.seed6 + (size_t)ngenes * 6);
    std::vector<double> cost(ngenes);
    for (int g = 0; g < ngenes; g++) cost[g] = (double)(gene_off[g + 1] - gene_off[g]);
    const int probe = (device_perms && B > 32 && ngenes >= 4 * (int)devs.size()) ? 8 : 0;   // phase-1 permutations
    struct Phase { int b0, nbp; bool measure; };
    std::vector<Phase> phases;
    if (probe) { phases.push_back({0, probe, true}); phases.push_back({probe, B - probe, false}); }
    else phases.push_back({0, B, false});
    double total_ms = 0.0;
    std::atomic<int> cancel(0);
    for (const Phase& ph : phases) {
        // longest-processing-time-first deal of the genes by the current cost estimate
        std::vector<std::vector<int>> assign(devs.size());
        {
            std::vector<int> order(ngenes);
            std::iota(order.begin(), order.end(), 0);
            std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return cost[x] > cost[y]; });
            std::vector<double> load(devs.size(), 0.0);
            for (int g : order) {
                size_t best = 0;
                for (size_t d = 1; d < devs.size(); d++) if (load[d] < load[best]) best = d;
                assign[best].push_back(g);
                load[best] += cost[g];
            }
            for (auto& v : assign) std::sort(v.begin(), v.end());
        }
        std::vector<scdd_stats> sts(devs.size());
        std::vector<ScddError> errs(devs.size(), ScddError{0, ""});
        std::vector<std::thread> th;
        std::atomic<int> running((int)devs.size());
        for (size_t d = 0; d < devs.size(); d++) {
            th.emplace_back([&, d]() {
                std::memset(&sts[d], 0, sizeof(scdd_stats));
                try {
                    const std::vector<int>& gs = assign[d];
                    const int ngl = (int)gs.size();
                    if (ngl == 0) { running.fetch_sub(1); return; }
                    CK(cudaSetDevice(devs[d]));
                    ShardStage& C = shard_stage(d);
                    std::vector<int64_t> off(ngl + 1, 0);
                    for (int i = 0; i < ngl; i++) off[i + 1] = off[i] + (gene_off[gs[i] + 1] - gene_off[gs[i]]);
                    const size_t nz = (size_t)off[ngl];
                    double* ly = (double*)C.stage[0].get(std::max<size_t>(8, nz * 8));
                    int32_t* cc = (int32_t*)C.stage[1].get(std::max<size_t>(8, nz * 4));
                    double* cd = call.cdr ? (double*)C.stage[2].get(std::max<size_t>(8, nz * 8)) : nullptr;
                    uint32_t* sd = (uint32_t*)C.stage[3].get((size_t)std::max(ngl, B) * 6 * 4 * 2);
                    uint32_t* sdo = sd + (size_t)std::max(ngl, B) * 6;
                    const size_t out_cnt = (size_t)ngl * ph.nbp;
                    double* out = (double*)C.stage[4].get((out_cnt * (call.jp_terms ? 1 + sets : 1) + ngl) * 8);
                    double* terms = call.jp_terms ? out + out_cnt : nullptr;
                    double* gcost = out + out_cnt * (call.jp_terms ? 1 + sets : 1);
                    std::vector<int32_t> pi;
                    if (call.perm_idx) pi.resize(nz * (size_t)B);
                    for (int i = 0; i < ngl; i++) {
                        const int g = gs[i];
                        const size_t o = (size_t)gene_off[g], n = (size_t)(gene_off[g + 1] - gene_off[g]);
                        std::memcpy(ly + off[i], logy + o, n * 8);
                        std::memcpy(cc + off[i], cond01 + o, n * 4);
                        if (cd) std::memcpy(cd + off[i], call.cdr + o, n * 8);
                        if (call.perm_idx) std::memcpy(&pi[(size_t)off[i] * B], call.perm_idx + o * (size_t)B, n * (size_t)B * 4);
                        if (device_perms && call.seed_mode == 0) std::memcpy(sd + (size_t)i * 6, &seeds_cur[(size_t)g * 6], 24);
                    }
                    if (device_perms && call.seed_mode == 1) std::memcpy(sd, call.seed6, (size_t)B * 24);
                    PermCall pc = call;
                    pc.B = ph.nbp; pc.b0 = 0; pc.nbp = ph.nbp;       // the shard's own result block has ph.nbp columns
                    pc.perm_idx = call.perm_idx ? pi.data() : nullptr;
                    pc.seed6 = device_perms ? sd : nullptr;
                    pc.cdr = cd;
                    pc.bf_perm = out; pc.jp_terms = terms;
                    pc.seed6_out = (device_perms && call.seed_mode == 0) ? sdo : nullptr;
                    pc.gene_cost = ph.measure ? gcost : nullptr;
                    if (device_perms && call.seed_mode == 1) pc.seed6 = sd + (size_t)ph.b0 * 6;   // this phase's streams
                    perm_bf_one(devs[d], ly, off.data(), cc, ngl, pc, &sts[d], nullptr, nullptr, intr ? &cancel : nullptr);
                    for (int i = 0; i < ngl; i++) {
                        const int g = gs[i];
                        std::memcpy(call.bf_perm + (size_t)g * B + ph.b0, out + (size_t)i * ph.nbp, (size_t)ph.nbp * 8);
                        if (terms)
                            std::memcpy(call.jp_terms + ((size_t)g * B + ph.b0) * sets, terms + (size_t)i * ph.nbp * sets,
                                        (size_t)ph.nbp * sets * 8);
                        if (pc.seed6_out) std::memcpy(&seeds_cur[(size_t)g * 6], sdo + (size_t)i * 6, 24);
                        if (ph.measure) cost[g] = gcost[i];
                    }
                } catch (const ScddError& e) {
                    errs[d] = e;
                } catch (const std::exception& e) {
                    errs[d] = ScddError{SCDD_ERR_CUDA, e.what()};
                }
                running.fetch_sub(1);
            });
        }
        // the calling thread polls the caller's interrupt hook while the shards run
        while (intr && running.load() > 0) {
            std::this_thread::sleep_for(std::chrono::milliseconds(20));
            if (!cancel.load() && intr(intr_arg)) cancel.store(1);
        }
        for (auto& t : th) t.join();
        for (auto& e : errs) if (e.code) throw e;
        double phase_ms = 0.0;
        for (auto& s1 : sts) { merge_stats(st, s1); phase_ms = std::max(phase_ms, s1.total_ms); }
        total_ms += phase_ms;     // shards run concurrently: the phase lasts as long as its slowest device
    }
    st.total_ms = total_ms;
    if (call.seed6_out && device_perms && call.seed_mode == 0)
        std::memcpy(call.seed6_out, seeds_cur.data(), (size_t)ngenes * 24);
}

static int32_t perm_bf_entry(const double* logy, const int64_t* gene_off, const int32_t* cond01, int32_t ngenes,
                             int32_t B, const int32_t* perm_idx, const uint32_t* seed6, const double* cdr,
                             const scdd_cfg* cfg, const scdd_perm_opts* opts, double* bf_perm, uint32_t* seed6_out,
                             scdd_stats* stats) {
    int rc = check_common(logy, gene_off, cond01, ngenes, cfg);
    if (rc) return rc;
    if (B < 0 || (ngenes > 0 && B > 0 && !bf_perm)) return fail(SCDD_ERR_BAD_ARG, "bad B / null output");
    if ((perm_idx != nullptr) == (seed6 != nullptr) && ngenes > 0 && B > 0)
        return fail(SCDD_ERR_BAD_ARG, "give exactly one of perm_idx (host-drawn) or seed6 (device-drawn)");
    const int seed_mode = opts ? opts->seed_mode : 0;
    if (seed_mode != 0 && seed_mode != 1) return fail(SCDD_ERR_BAD_ARG, "seed_mode must be 0 (per gene) or 1 (per permutation)");
    if (seed_mode == 1 && !seed6 && ngenes > 0 && B > 0) return fail(SCDD_ERR_BAD_ARG, "per-permutation streams need seed6");
    scdd_stats st;
    std::memset(&st, 0, sizeof(st));
    return guarded([&]() -> int {
        std::vector<int> devs = active_devices();
        PermCall call;
        call.B = B; call.b0 = 0; call.nbp = B;
        call.perm_idx = perm_idx; call.seed6 = seed6; call.seed_mode = seed_mode; call.cdr = cdr; call.cfg = cfg;
        call.bf_perm = bf_perm; call.jp_terms = opts ? opts->jp_terms : nullptr; call.seed6_out = seed6_out;
        scdd_interrupt_fn intr = opts ? opts->interrupt : nullptr;
        void* intr_arg = opts ? opts->interrupt_arg : nullptr;
        if (ngenes > 0 && B > 0) {
            if (devs.size() == 1) perm_bf_one(devs[0], logy, gene_off, cond01, ngenes, call, &st, intr, intr_arg);
            else perm_bf_sharded(devs, logy, gene_off, cond01, ngenes, call, st, intr, intr_arg);
        }
        if (stats) *stats = st;
        if (st.failed_sets > 0) return fail(SCDD_ERR_FIT_FAILED, "some permuted fit-sets had no valid model; their bf_perm is NaN");
        return SCDD_OK;
    });
}

int32_t scdd_perm_bf(const double* logy, const int64_t* gene_off, const int32_t* cond01, int32_t ngenes,
                     int32_t B, const int32_t* perm_idx, const uint32_t* seed6, const double* cdr,
                     const scdd_cfg* cfg, double* bf_perm, uint32_t* seed6_out, scdd_stats* stats) {
    return perm_bf_entry(logy, gene_off, cond01, ngenes, B, perm_idx, seed6, cdr, cfg, nullptr, bf_perm, seed6_out, stats);
}

int32_t scdd_perm_bf_ex(const double* logy, const int64_t* gene_off, const int32_t* cond01, int32_t ngenes,
                        int32_t B, const int32_t* perm_idx, const uint32_t* seed6, const double* cdr,
                        const scdd_cfg* cfg, const scdd_perm_opts* opts, double* bf_perm, uint32_t* seed6_out,
                        scdd_stats* stats) {
    return perm_bf_entry(logy, gene_off, cond01, ngenes, B, perm_idx, seed6, cdr, cfg, opts, bf_perm, seed6_out, stats);
}

void scdd_release_cache(void) {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    for (auto& kv : g_cache) {
        CachedPipeline& C = *kv.second;
        std::lock_guard<std::mutex> l2(C.mu);
        if (cudaSetDevice(kv.first) != cudaSuccess) { cudaGetLastError(); continue; }
        if (C.s) cudaStreamSynchronize(C.s);
        C.P.reset();
        if (C.s) { cudaStreamDestroy(C.s); C.s = nullptr; }
    }
    std::lock_guard<std::mutex> l3(g_stage_mu);
    for (auto& sp : g_stage) for (auto& b : sp->stage) b.release();
}

// ---------------------------------------------------------------- KS
// asymptotic two-sided p-value of ks.test (R >= 4.2: psmirnov/pkolmogorov_two_limit upper tail;
// older R: 1 - pKS2) -- evaluated on the host, one scalar series per gene
static double ks_pvalue(double stat, int old_form) {
    const double tol = 1e-6;
    const int k_max = (int)std::sqrt(2.0 - std::log(tol));
    const double PI = 3.14159265358979323846;
    const double inv_sqrt_2pi = 0.398942280401432677939946059934;
    double p;
    if (stat <= 0) p = old_form ? 1.0 - 0.0 : 1.0;
    else if (stat < 1.0) {
        double z = -(PI / 2.0 * PI / 4.0) / (stat * stat), w = std::log(stat), s = 0.0;
        for (int k = 1; k < k_max; k += 2) s += std::exp(k * k * z - w);
        p = 1.0 - s / inv_sqrt_2pi;
    } else if (!old_form) {
        double z = -2.0 * stat * stat, sg = -1.0, old = 0.0, nw = 2.0 * std::exp(z);
        int k = 2;
        while (std::fabs(old - nw) > tol) { old = nw; nw += 2.0 * sg * std::exp(z * k * k); sg *= -1.0; k++; }
        p = nw;
    } else {
        double z = -2.0 * stat * stat, sg = -1.0, old = 0.0, nw = 1.0;
        int k = 1;
        while (std::fabs(old - nw) > tol) { old = nw; nw += 2.0 * sg * std::exp(z * k * k); sg *= -1.0; k++; }
        p = 1.0 - nw;
    }
    return std::min(1.0, std::max(0.0, p));
}

int32_t scdd_ks(const double* vals, const int64_t* gene_off, const int32_t* cond01, int32_t ngenes,
                int32_t old_pvalue, double* D, double* p, scdd_stats* stats) {
    scdd_cfg cfg;
    scdd_default_cfg(&cfg);
    int rc = check_common(vals, gene_off, cond01, ngenes, &cfg);
    if (rc) return rc;
    if (ngenes > 0 && (!D || !p)) return fail(SCDD_ERR_BAD_ARG, "null output");
    if (ngenes == 0) return SCDD_OK;
    return guarded([&]() -> int {
        std::vector<int> devs = active_devices();
        Pipeline P;
        cudaStream_t s;
        CK(cudaSetDevice(devs[0]));
        CK(cudaStreamCreate(&s));
        std::vector<int32_t> h_n1(ngenes);
        try {
            P.upload(devs[0], vals, gene_off, cond01, ngenes, &cfg, nullptr, s);
            int np = 2;
            while (np < P.max_n) np <<= 1;
            size_t smem = (size_t)np * 12;
            DevBuf<double> d_D;
            DevBuf<uint64_t> d_key;
            DevBuf<uint32_t> d_idx;
            d_D.alloc(ngenes);
            int grid;
            if (smem <= 200 * 1024) {
                CK(cudaFuncSetAttribute(ks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (227 * 1024) / std::max<size_t>(smem, 1)));
                grid = std::min(ngenes, P.n_sm * per_sm);
            } else {
                grid = std::min(ngenes, P.n_sm * 2);
                d_key.alloc((size_t)grid * np);
                d_idx.alloc((size_t)grid * np);
                smem = 0;
            }
            cudaEvent_t t0, t1;
            CK(cudaEventCreate(&t0)); CK(cudaEventCreate(&t1));
            CK(cudaEventRecord(t0, s));
            ks_kernel<<<grid, smem ? 128 : 512, smem, s>>>(P.d_logy.p, P.d_off.p, P.d_part.p, P.d_n1.p, ngenes, np, d_D.p, d_key.p, d_idx.p);
            CK(cudaGetLastError());
            CK(cudaEventRecord(t1, s));
            CK(cudaMemcpyAsync(D, d_D.p, ngenes * sizeof(double), cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(h_n1.data(), P.d_n1.p, ngenes * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            if (stats) {
                std::memset(stats, 0, sizeof(*stats));
                float ms = 0;
                CK(cudaEventElapsedTime(&ms, t0, t1));
                stats->total_ms = ms;
                stats->h2d_bytes = (double)P.nnz * 12 + (ngenes + 1) * 8.0;
                stats->d2h_bytes = ngenes * 12.0;
            }
            cudaEventDestroy(t0); cudaEventDestroy(t1);
        } catch (...) {
            cudaStreamDestroy(s);
            throw;
        }
        CK(cudaStreamDestroy(s));
        for (int g = 0; g < ngenes; g++) {
            double nx = h_n1[g], ny = (double)(gene_off[g + 1] - gene_off[g]) - nx;
            if (nx < 1 || ny < 1) { p[g] = NAN; continue; }
            double en = nx * ny / (nx + ny);
            p[g] = ks_pvalue(std::sqrt(en) * D[g], old_pvalue);
        }
        return SCDD_OK;
    });
}

// ---------------------------------------------------------------- feDP (R/classify.dd.R:314-383)
// two-sided Student p-value 2 * pt(-|t|, df) = I_{df/(df+t^2)}(df/2, 1/2): continued fraction of the regularised
// incomplete beta function (modified Lentz), with R's normal approximation for df > 4e5 (nmath/pt.c)
static double beta_cf(double a, double b, double x) {
    const double tiny = 1e-300, eps = 1e-16;
    double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0, d = 1.0 - qab * x / qap;
    if (std::fabs(d) < tiny) d = tiny;
    d = 1.0 / d;
    double h = d;
    for (int m = 1; m <= 200000; m++) {
        double m2 = 2.0 * m;
        double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d; if (std::fabs(d) < tiny) d = tiny;
        c = 1.0 + aa / c; if (std::fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        h *= d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d; if (std::fabs(d) < tiny) d = tiny;
        c = 1.0 + aa / c; if (std::fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        double del = d * c;
        h *= del;
        if (std::fabs(del - 1.0) < eps) break;
    }
    return h;
}

static double student_two_sided(double t, double df) {
    if (std::isnan(t) || std::isnan(df) || !(df > 0)) return NAN;
    const double t2 = t * t;
    if (t2 == 0) return 1.0;
    if (std::isinf(t2)) return 0.0;
    if (df > 4e5) {
        double val = 1.0 / (4.0 * df);
        return std::erfc(std::fabs(t) * (1.0 - val) / std::sqrt(1.0 + t2 * 2.0 * val) / std::sqrt(2.0));
    }
    const double a = 0.5 * df, b = 0.5;
    const double lx = -std::log1p(t2 / df), l1x = std::log(t2 / (df + t2));   // log x, log(1 - x), x = df/(df+t^2)
    const double x = std::exp(lx), x1 = t2 / (df + t2);
    // log Beta(a, 1/2) = lgamma(1/2) - [lgamma(a + 1/2) - lgamma(a)]; the bracket by its asymptotic series for large a
    // (the difference of two lgammas of magnitude a log a loses ~ log10(a log a) digits)
    double dl;
    if (a >= 20.0) {
        const double i1 = 1.0 / a, i2 = i1 * i1;
        dl = 0.5 * std::log(a) + i1 * (-1.0 / 8 + i2 * (1.0 / 192 + i2 * (-1.0 / 640 + i2 * (17.0 / 14336 - i2 * (31.0 / 18432)))));
    } else {
        dl = std::lgamma(a + b) - std::lgamma(a);
    }
    const double lbeta = 0.57236494292470008707 - dl;   // lgamma(1/2) = log(sqrt(pi))
    const double front = std::exp(a * lx + b * l1x - lbeta);
    if (x < (a + 1.0) / (a + b + 2.0)) return std::min(1.0, front * beta_cf(a, b, x) / a);
    return std::min(1.0, std::max(0.0, 1.0 - front * beta_cf(b, a, x1) / b));
}

int32_t scdd_fedp(const double* logy, const int64_t* gene_off, const int32_t* cond01, int32_t ngenes,
                  const int32_t* cls_oa, const int32_t* cls_c, const double* cdr, int32_t min_size,
                  double* p_shift, double* p_fisher, int32_t* counts3, scdd_stats* stats) {
    scdd_cfg cfg;
    scdd_default_cfg(&cfg);
    int rc = check_common(logy, gene_off, cond01, ngenes, &cfg);
    if (rc) return rc;
    if (ngenes > 0 && (!cls_oa || !cls_c || !p_shift || !p_fisher)) return fail(SCDD_ERR_BAD_ARG, "null pointer");
    if (ngenes == 0) return SCDD_OK;
    return guarded([&]() -> int {
        std::vector<int> devs = active_devices();
        CK(cudaSetDevice(devs[0]));
        cudaStream_t s;
        CK(cudaStreamCreate(&s));
        struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } sg{s};
        const int64_t base = gene_off[0];
        const size_t nnz = (size_t)(gene_off[ngenes] - base);
        std::vector<int64_t> h_off(gene_off, gene_off + ngenes + 1);
        int max_n = 0;
        for (int g = 0; g <= ngenes; g++) h_off[g] -= base;
        for (int g = 0; g < ngenes; g++) max_n = std::max<int>(max_n, (int)(h_off[g + 1] - h_off[g]));
        DevBuf<double> d_y, d_cdr, d_lf, d_part;
        DevBuf<int64_t> d_off;
        DevBuf<int32_t> d_cond, d_oa, d_c;
        DevBuf<FedpGene> d_out;
        DevBuf<FisherJob> d_jobs;
        d_y.alloc(nnz); d_off.alloc(ngenes + 1); d_cond.alloc(nnz); d_oa.alloc(nnz); d_c.alloc(nnz); d_out.alloc(ngenes);
        struct Events {
            cudaEvent_t a = nullptr, b = nullptr;
            ~Events() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
        } ev;
        CK(cudaEventCreate(&ev.a)); CK(cudaEventCreate(&ev.b));
        cudaEvent_t t0 = ev.a, t1 = ev.b;
        CK(cudaEventRecord(t0, s));
        CK(cudaMemcpyAsync(d_y.p, logy + base, nnz * sizeof(double), cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(d_off.p, h_off.data(), (ngenes + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(d_cond.p, cond01 + base, nnz * sizeof(int32_t), cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(d_oa.p, cls_oa + base, nnz * sizeof(int32_t), cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(d_c.p, cls_c + base, nnz * sizeof(int32_t), cudaMemcpyHostToDevice, s));
        if (cdr) {
            d_cdr.alloc(nnz);
            CK(cudaMemcpyAsync(d_cdr.p, cdr + base, nnz * sizeof(double), cudaMemcpyHostToDevice, s));
        }
        int n_sm = 0;
        CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, devs[0]));
        const int wpb = 8;
        const int blocks = std::min((ngenes + wpb - 1) / wpb, n_sm * 8);
        fedp_stats_kernel<<<blocks, wpb * 32, 0, s>>>(d_y.p, d_off.p, d_cond.p, d_oa.p, d_c.p, cdr ? d_cdr.p : nullptr,
                                                      ngenes, min_size, d_out.p);
        CK(cudaGetLastError());
        std::vector<FedpGene> h(ngenes);
        CK(cudaMemcpyAsync(h.data(), d_out.p, ngenes * sizeof(FedpGene), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        // ---- Fisher jobs: rows sorted ascending so the two largest are the inner / dependent dimensions;
        // tables whose enumeration is long are split over several CTAs (partials summed here in job order)
        std::vector<FisherJob> jobs;
        std::vector<int> job_gene;
        const long long CHUNK = 1LL << 22, MAX_JOBS = 1LL << 20;
        double skipped = 0;
        for (int g = 0; g < ngenes; g++) {
            p_fisher[g] = NAN;
            const FedpGene& f = h[g];
            if (f.r < 2) continue;
            int ord[SCDD_GMAX];
            for (int i = 0; i < f.r; i++) ord[i] = i;
            std::stable_sort(ord, ord + f.r, [&](int x, int y) { return f.row[x] < f.row[y]; });
            FisherJob jb{};
            jb.r = f.r; jb.c0 = f.c0;
            for (int i = 0; i < f.r; i++) { jb.row[i] = f.row[ord[i]]; jb.a[i] = f.a[ord[i]]; }
            jb.tol = f.r == 2 ? std::log1p(1e-7) : 3.45254e-7;
            long double items = (jb.row[f.r - 2] + FISHER_SEG) / FISHER_SEG;
            for (int i = 0; i < f.r - 2; i++) items *= (long double)(jb.row[i] + 1);
            if (items / CHUNK + jobs.size() >= (long double)MAX_JOBS) { skipped++; continue; }
            const long long it = (long long)items;
            for (long long b0 = 0; b0 < it; b0 += CHUNK) {
                jb.slot = (int32_t)jobs.size();
                jb.item_begin = b0; jb.item_end = std::min(it, b0 + CHUNK);
                jobs.push_back(jb);
                job_gene.push_back(g);
            }
        }
        if (!jobs.empty()) {
            std::vector<double> lf(max_n + 1);
            for (int i = 0; i <= max_n; i++) lf[i] = std::lgamma((double)i + 1.0);
            d_lf.alloc(lf.size()); d_jobs.alloc(jobs.size()); d_part.alloc(jobs.size());
            CK(cudaMemcpyAsync(d_lf.p, lf.data(), lf.size() * sizeof(double), cudaMemcpyHostToDevice, s));
            CK(cudaMemcpyAsync(d_jobs.p, jobs.data(), jobs.size() * sizeof(FisherJob), cudaMemcpyHostToDevice, s));
            const int grid = (int)std::min<size_t>(jobs.size(), (size_t)n_sm * 8);
            fisher_kernel<<<grid, FISHER_THREADS, 0, s>>>(d_jobs.p, (int)jobs.size(), d_lf.p, d_part.p);
            CK(cudaGetLastError());
            std::vector<double> part(jobs.size());
            CK(cudaMemcpyAsync(part.data(), d_part.p, jobs.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
            CK(cudaEventRecord(t1, s));
            CK(cudaStreamSynchronize(s));
            for (size_t j = 0; j < jobs.size(); j++) {
                double& p = p_fisher[job_gene[j]];
                p = std::isnan(p) ? part[j] : p + part[j];
            }
        } else {
            CK(cudaEventRecord(t1, s));
            CK(cudaStreamSynchronize(s));
        }
        double undefined = 0;
        for (int g = 0; g < ngenes; g++) {
            p_shift[g] = student_two_sided(h[g].tstat, h[g].df);
            if (h[g].flag & 1) { p_shift[g] = NAN; undefined++; }   // the reference would jitter with runif here
            if (counts3) { counts3[3 * g] = h[g].c_oa; counts3[3 * g + 1] = h[g].c_c1; counts3[3 * g + 2] = h[g].c_c2; }
        }
        if (stats) {
            std::memset(stats, 0, sizeof(*stats));
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, t0, t1));
            stats->total_ms = ms;
            stats->fit_sets = (double)jobs.size();
            stats->const_sets = undefined;
            stats->failed_sets = skipped;
            stats->h2d_bytes = (double)nnz * (20 + (cdr ? 8 : 0)) + (ngenes + 1) * 8.0;
            stats->d2h_bytes = (double)ngenes * sizeof(FedpGene);
        }
        return SCDD_OK;
    });
}

int32_t scdd_selftest_student_p(const double* t, const double* df, int32_t n, double* out) {
    if (n < 0 || (n > 0 && (!t || !df || !out))) return fail(SCDD_ERR_BAD_ARG, "bad arguments");
    for (int i = 0; i < n; i++) out[i] = student_two_sided(t[i], df[i]);
    return SCDD_OK;
}

// ---------------------------------------------------------------- host RNG helpers
int32_t scdd_rng_gene_seeds(uint32_t seed, int32_t ngenes, uint32_t* out6) {
    if (ngenes < 0 || (ngenes > 0 && !out6)) return fail(SCDD_ERR_BAD_ARG, "bad arguments");
    uint32_t s[6];
    set_seed_lecuyer(seed, s);
    Mat3 A1, A2;
    jump_matrices(76, A1, A2);   // nextRNGSubStream
    for (int g = 0; g < ngenes; g++) {
        std::memcpy(out6 + (size_t)g * 6, s, sizeof(s));
        apply_jump(A1, A2, s, s);
    }
    return SCDD_OK;
}

int32_t scdd_rng_sample_perms(uint32_t* state6, int32_t n, int32_t B, int32_t* out) {
    if (!state6 || n < 0 || B < 0 || (!out && (size_t)n * B > 0)) return fail(SCDD_ERR_BAD_ARG, "bad arguments");
    LecuyerGen gen;
    std::memcpy(gen.st.s, state6, 24);
    std::vector<int32_t> scratch(std::max(1, n));
    for (int b = 0; b < B; b++) {
        int32_t* o = out + (size_t)b * n;
        r_sample_perm(gen, (uint32_t)n, scratch.data(), o);
        for (int i = 0; i < n; i++) o[i] += 1;
    }
    std::memcpy(state6, gen.st.s, 24);
    return SCDD_OK;
}

int32_t scdd_rng_mt_sample_perms(uint32_t seed, int32_t n, int32_t count, int32_t* out) {
    if (n < 0 || count < 0 || (!out && (size_t)n * count > 0)) return fail(SCDD_ERR_BAD_ARG, "bad arguments");
    MersenneTwister mt;
    mt.seed(seed);
    std::vector<int32_t> scratch(std::max(1, n));
    for (int b = 0; b < count; b++) {
        int32_t* o = out + (size_t)b * n;
        r_sample_perm(mt, (uint32_t)n, scratch.data(), o);
        for (int i = 0; i < n; i++) o[i] += 1;
    }
    return SCDD_OK;
}

// ---------------------------------------------------------------- dense matrix <-> CSR (host marshalling)
static bool dense_view(const double* X, int64_t ngenes, int64_t ncells, int32_t col_major, const int32_t* cc, dense::View& v) {
    if (ngenes < 0 || ncells < 0 || ((!X || !cc) && ngenes * ncells > 0)) return false;
    v = dense::View{X, ngenes, ncells, col_major != 0};
    return true;
}

int32_t scdd_dense_scan(const double* X, int64_t ngenes, int64_t ncells, int32_t col_major, const int32_t* cell_cond01,
                        int32_t* nnz0, int32_t* nnz1, uint8_t* const0, uint8_t* const1, int64_t* n_negative) {
    dense::View v;
    if (!dense_view(X, ngenes, ncells, col_major, cell_cond01, v)) return fail(SCDD_ERR_BAD_ARG, "bad arguments");
    dense::scan(v, cell_cond01, nnz0, nnz1, const0, const1, n_negative);
    return SCDD_OK;
}

int32_t scdd_dense_to_csr(const double* X, int64_t ngenes, int64_t ncells, int32_t col_major, const int32_t* cell_cond01,
                          const int32_t* genes, int32_t nsel, int32_t take_log, int32_t incl_zero,
                          const int64_t* gene_off, double* vals, int32_t* cond01) {
    dense::View v;
    if (!dense_view(X, ngenes, ncells, col_major, cell_cond01, v) || nsel < 0 || !gene_off ||
        (gene_off[nsel] > 0 && (!vals || !cond01)))
        return fail(SCDD_ERR_BAD_ARG, "bad arguments");
    if (!dense::to_csr(v, cell_cond01, genes, nsel, take_log != 0, incl_zero != 0, gene_off, vals, cond01))
        return fail(SCDD_ERR_BAD_ARG, "gene_off does not match the number of kept cells of the selected genes");
    return SCDD_OK;
}

int32_t scdd_dense_cell_covariate(const double* X, int64_t ngenes, int64_t ncells, int32_t col_major,
                                  const double* cell_values, const int32_t* genes, int32_t nsel, const int64_t* gene_off,
                                  double* out) {
    if (ngenes < 0 || ncells < 0 || nsel < 0 || !gene_off || ((!X || !cell_values) && ngenes * ncells > 0) ||
        (gene_off[nsel] > 0 && !out))
        return fail(SCDD_ERR_BAD_ARG, "bad arguments");
    dense::View v{X, ngenes, ncells, col_major != 0};
    if (!dense::cell_values_to_csr(v, cell_values, genes, nsel, gene_off, out))
        return fail(SCDD_ERR_BAD_ARG, "gene_off does not match the nonzero cells of the selected genes");
    return SCDD_OK;
}

// ---- classifyDD, R/classify.dd.R:86-238 ----
namespace {
struct GeneClusters {           // what classifyDD derives from one gene's three clusterings
    int n_oa, n_c1, n_c2;       // luOutlier of oa / c1 / c2
    int first, count;           // its comparisons in the flat list
};

// R/classify.dd.R:124-158: labels above the FIRST missing one move down by one (the reference handles one gap)
inline void compress_labels(std::vector<int>& cls) {
    int mx = 0, present = 0;
    bool has[SCDD_GMAX + 2] = {false};
    for (int l : cls) if (l >= 1 && l <= SCDD_GMAX) { has[l] = true; mx = std::max(mx, l); }
    for (int l = 1; l <= mx; l++) present += has[l];
    if (present == mx) return;
    int missing = 1;
    while (has[missing]) missing++;
    for (int& l : cls) if (l > missing) l--;
}

struct Posterior { double nk[SCDD_GMAX], mk[SCDD_GMAX], sk[SCDD_GMAX], ak[SCDD_GMAX], bk[SCDD_GMAX]; };

// getPosteriorParams, R/classify.dd.R:19-37: normal-gamma posterior of every cluster label 1..GMAX
inline void posterior_params(const std::vector<double>& y, const std::vector<int>& cls, const scdd_cfg& c, Posterior& P) {
    double sy[SCDD_GMAX] = {0}, syy[SCDD_GMAX] = {0};
    for (int k = 0; k < SCDD_GMAX; k++) P.nk[k] = 0;
    for (size_t i = 0; i < y.size(); i++) {
        const int l = cls[i];
        if (l < 1 || l > SCDD_GMAX) continue;
        P.nk[l - 1] += 1.0; sy[l - 1] += y[i]; syy[l - 1] += y[i] * y[i];
    }
    for (int k = 0; k < SCDD_GMAX; k++) {
        P.sk[k] = c.s0 + P.nk[k];
        P.mk[k] = (c.s0 * c.m0 + sy[k]) / P.sk[k];
        P.ak[k] = c.a0 + P.nk[k];
        P.bk[k] = c.b0 + syy[k] + c.s0 * c.m0 * c.m0 - P.sk[k] * P.mk[k] * P.mk[k];
    }
}
}  // namespace

void scdd_default_classify_opts(scdd_classify_opts* o) {
    if (!o) return;
    std::memset(o, 0, sizeof(*o));
    o->ndraws = 10000;
    o->seed = 1;
}

int32_t scdd_classify(const double* logy, const int64_t* gene_off, const int32_t* cond01, int32_t ngenes,
                      const int32_t* cls_oa, const int32_t* cls_c, const int32_t* genes, int32_t nsel,
                      const double* p_shift, const scdd_cfg* cfg, const scdd_classify_opts* opts,
                      int32_t* category, scdd_stats* stats) {
    scdd_cfg dcfg;
    scdd_default_cfg(&dcfg);
    if (!cfg) cfg = &dcfg;
    int rc = check_common(logy, gene_off, cond01, ngenes, cfg);
    if (rc) return rc;
    if (nsel < 0 || (nsel > 0 && (!genes || !category || !cls_oa || !cls_c))) return fail(SCDD_ERR_BAD_ARG, "null pointer");
    scdd_classify_opts o;
    scdd_default_classify_opts(&o);
    if (opts) o = *opts;
    if (o.ndraws < 1) return fail(SCDD_ERR_BAD_ARG, "ndraws must be positive");
    for (int j = 0; j < nsel; j++)
        if (genes[j] < 0 || genes[j] >= ngenes) return fail(SCDD_ERR_BAD_ARG, "gene index out of range");
    if (stats) std::memset(stats, 0, sizeof(*stats));
    if (nsel == 0) return SCDD_OK;
    return guarded([&]() -> int {
        // ---- host, threaded over the selected genes: label bookkeeping, posteriors, the list of comparisons ----
        std::vector<GeneClusters> gc((size_t)nsel);
        std::vector<std::vector<McPair>> per_gene((size_t)nsel);
        const int ms = cfg->min_size;
        dense::parallel_ranges(nsel, 64, [&](int64_t lo, int64_t hi) {
            std::vector<int> oa, c1, c2;
            std::vector<double> y1, y2;
            for (int64_t j = lo; j < hi; j++) {
                const int g = genes[j];
                const int64_t b = gene_off[g], e = gene_off[g + 1];
                oa.clear(); c1.clear(); c2.clear(); y1.clear(); y2.clear();
                for (int64_t i = b; i < e; i++) {
                    oa.push_back(cls_oa[i]);
                    if (cond01[i] == 0) { c1.push_back(cls_c[i]); y1.push_back(logy[i]); }
                    else { c2.push_back(cls_c[i]); y2.push_back(logy[i]); }
                }
                compress_labels(oa); compress_labels(c1); compress_labels(c2);
                auto big = [&](const std::vector<int>& cls, int* labs) {    // findOutliers: labels with >= min.size members
                    int cnt[SCDD_GMAX + 1] = {0}, n = 0;
                    for (int l : cls) if (l >= 1 && l <= SCDD_GMAX) cnt[l]++;
                    for (int l = 1; l <= SCDD_GMAX; l++) if (cnt[l] >= ms && cnt[l] > 0) labs[n++] = l;
                    return n;
                };
                int l_oa[SCDD_GMAX], l1[SCDD_GMAX], l2[SCDD_GMAX];
                GeneClusters& G = gc[(size_t)j];
                G.n_oa = big(oa, l_oa); G.n_c1 = big(c1, l1); G.n_c2 = big(c2, l2);
                Posterior P1, P2;
                posterior_params(y1, c1, *cfg, P1);
                posterior_params(y2, c2, *cfg, P2);
                for (int a = 0; a < G.n_c1; a++)
                    for (int q = 0; q < G.n_c2; q++) {
                        const int ka = l1[a] - 1, kb = l2[q] - 1;
                        McPair m;
                        m.df1 = std::max(1.0, std::nearbyint(P1.ak[ka]));              // rt(df = round(ak)), :173
                        m.scale1 = std::sqrt(P1.bk[ka] / (P1.ak[ka] * P1.sk[ka]));
                        m.m1 = P1.mk[ka];
                        m.df2 = std::max(1.0, std::nearbyint(P2.ak[kb]));
                        m.scale2 = std::sqrt(P2.bk[kb] / (P2.ak[kb] * P2.sk[kb]));
                        m.m2 = P2.mk[kb];
                        per_gene[(size_t)j].push_back(m);
                    }
            }
        });
        std::vector<McPair> pairs;
        for (int j = 0; j < nsel; j++) {
            gc[(size_t)j].first = (int)pairs.size();
            gc[(size_t)j].count = (int)per_gene[(size_t)j].size();
            pairs.insert(pairs.end(), per_gene[(size_t)j].begin(), per_gene[(size_t)j].end());
        }
        const int ncomp = (int)pairs.size();
        std::vector<int32_t> flags((size_t)ncomp, 0);
        std::vector<double> lohi(o.comp_detail ? (size_t)ncomp * 2 : 0);
        // ---- device: the Monte-Carlo draws, one warp per comparison ----
        if (ncomp > 0) {
            std::vector<int> devs = active_devices();
            CK(cudaSetDevice(devs[0]));
            cudaStream_t s;
            CK(cudaStreamCreate(&s));
            struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } sg{s};
            struct Events {
                cudaEvent_t a = nullptr, b = nullptr;
                ~Events() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
            } ev;
            CK(cudaEventCreate(&ev.a)); CK(cudaEventCreate(&ev.b));
            DevBuf<McPair> d_pairs;
            DevBuf<int32_t> d_flags;
            DevBuf<double> d_lohi;
            d_pairs.alloc((size_t)ncomp); d_flags.alloc((size_t)ncomp);
            if (o.comp_detail) d_lohi.alloc((size_t)ncomp * 2);
            CK(cudaMemcpyAsync(d_pairs.p, pairs.data(), (size_t)ncomp * sizeof(McPair), cudaMemcpyHostToDevice, s));
            int n_sm = 0;
            CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, devs[0]));
            const int wpb = 4;
            const int blocks = std::max(1, std::min((ncomp + wpb - 1) / wpb, n_sm * 16));
            CK(cudaEventRecord(ev.a, s));
            classify_mc_kernel<<<blocks, wpb * 32, 0, s>>>(d_pairs.p, ncomp, o.ndraws, (unsigned long long)o.seed, d_flags.p,
                                                           o.comp_detail ? d_lohi.p : nullptr);
            CK(cudaGetLastError());
            CK(cudaEventRecord(ev.b, s));
            CK(cudaMemcpyAsync(flags.data(), d_flags.p, (size_t)ncomp * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
            if (o.comp_detail) CK(cudaMemcpyAsync(lohi.data(), d_lohi.p, (size_t)ncomp * 16, cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            if (stats) {
                float ms_k = 0;
                CK(cudaEventElapsedTime(&ms_k, ev.a, ev.b));
                stats->total_ms = ms_k;
                stats->h2d_bytes = (double)ncomp * sizeof(McPair);
                stats->d2h_bytes = (double)ncomp * (o.comp_detail ? 20 : 4);
            }
        }
        // ---- the decision tree, :189-234 ----
        for (int j = 0; j < nsel; j++) {
            const GeneClusters& G = gc[(size_t)j];
            int tot = 0;
            for (int t = 0; t < G.count; t++) tot += flags[(size_t)(G.first + t)];
            int cat = SCDD_CAT_NC;
            if (G.n_c1 == G.n_c2) {
                if (G.n_c1 == 1) cat = tot > 0 ? SCDD_CAT_DE : SCDD_CAT_NC;
                else if (G.n_oa == G.n_c1) cat = tot <= G.n_c1 ? SCDD_CAT_DP : SCDD_CAT_NC;
                else if (G.n_oa > G.n_c1) {
                    // multimodal DE: every cross pair but the matching ones differs AND the overall mean shifts (:206-214)
                    if (tot > G.n_c1 * (G.n_c1 - 1)) cat = (p_shift && p_shift[genes[j]] < 0.01) ? SCDD_CAT_DE : SCDD_CAT_NC;
                }
            } else {
                cat = tot == G.count ? SCDD_CAT_DB : SCDD_CAT_DM;
            }
            category[j] = cat;
        }
        if (o.comp_off) {
            for (int j = 0; j < nsel; j++) o.comp_off[j] = gc[(size_t)j].first;
            o.comp_off[nsel] = ncomp;
        }
        if (o.comp_detail) {
            const int64_t m = std::min<int64_t>(ncomp, o.comp_cap);
            for (int64_t t = 0; t < m; t++) {
                double* d = o.comp_detail + t * 9;
                const McPair& q = pairs[(size_t)t];
                d[0] = q.df1; d[1] = q.scale1; d[2] = q.m1; d[3] = q.df2; d[4] = q.scale2; d[5] = q.m2;
                d[6] = (double)flags[(size_t)t]; d[7] = lohi[(size_t)t * 2]; d[8] = lohi[(size_t)t * 2 + 1];
            }
        }
        return SCDD_OK;
    });
}

int32_t scdd_selftest_student_t(double df, int32_t n, uint64_t seed, double* out) {
    if (n < 0 || (n > 0 && !out) || !(df >= 1.0)) return fail(SCDD_ERR_BAD_ARG, "bad arguments");
    if (device_count_quiet() < 1) return fail(SCDD_ERR_NO_DEVICE, "no CUDA device available: libscdd_b200 has no CPU fallback");
    if (n == 0) return SCDD_OK;
    return guarded([&]() -> int {
        std::vector<int> devs = active_devices();
        CK(cudaSetDevice(devs[0]));
        DevBuf<double> d;
        d.alloc((size_t)n);
        student_t_kernel<<<(n + 127) / 128, 128>>>(df, n, (unsigned long long)seed, d.p);
        CK(cudaGetLastError());
        CK(cudaMemcpy(out, d.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
        return SCDD_OK;
    });
}

void scdd_default_glm_opts(scdd_glm_opts* o) {
    if (!o) return;
    o->prior_scale = 2.5;
    o->prior_scale_for_intercept = 10.0;
    o->prior_df = 1.0;
    o->epsilon = 1e-8;
    o->maxit = 100;
    o->reserved = 0;
}

// bayesglm's scaled = TRUE rule for one predictor column: 1 for a constant, max - min for a two-valued column,
// 2 sd(x) (n - 1 denominator) otherwise
static double predictor_scale(std::vector<double> x) {
    if (x.empty()) return 1.0;
    const size_t n = x.size();
    double mean = 0;
    for (double v : x) mean += v;
    mean /= (double)n;
    double ss = 0;
    for (double v : x) ss += (v - mean) * (v - mean);
    std::sort(x.begin(), x.end());
    const size_t ncat = (size_t)(std::unique(x.begin(), x.end()) - x.begin());
    if (ncat == 2) return x[1] - x[0];
    if (ncat > 2) return 2.0 * std::sqrt(ss / (double)(n - 1));
    return 1.0;
}

int32_t scdd_test_zeroes(const double* X, int64_t ngenes, int64_t ncells, int32_t col_major, const int32_t* cell_level2,
                         const int32_t* these, int64_t nthese, const scdd_glm_opts* opts, double* pvalue,
                         double* detail, scdd_stats* stats) {
    if (ngenes < 0 || ncells < 0 || nthese < 0 || nthese > INT32_MAX || ncells > INT32_MAX - 64 ||
        (ngenes * ncells > 0 && (!X || !cell_level2)) || (nthese > 0 && !pvalue))
        return fail(SCDD_ERR_BAD_ARG, "bad arguments");
    if (!these && nthese > ngenes) return fail(SCDD_ERR_BAD_ARG, "nthese exceeds the number of rows");
    if (these)
        for (int64_t j = 0; j < nthese; j++)
            if (these[j] < 0 || these[j] >= ngenes) return fail(SCDD_ERR_BAD_ARG, "row index out of range");
    if (stats) std::memset(stats, 0, sizeof(*stats));
    if (nthese == 0) return SCDD_OK;
    scdd_glm_opts o;
    scdd_default_glm_opts(&o);
    if (opts) o = *opts;
    if (!(o.prior_scale > 0) || !(o.prior_scale_for_intercept > 0) || !(o.prior_df > 0) || !(o.epsilon > 0) || o.maxit < 1)
        return fail(SCDD_ERR_BAD_ARG, "bad glm options");
    int64_t n2 = 0;
    for (int64_t c = 0; c < ncells; c++) n2 += cell_level2[c] != 0;
    if (n2 == 0 || n2 == ncells) return fail(SCDD_ERR_BAD_ARG, "factor(cond) needs two levels");
    if (device_count_quiet() < 1) return fail(SCDD_ERR_NO_DEVICE, "no CUDA device available: libscdd_b200 has no CPU fallback");
    return guarded([&]() -> int {
        const int64_t W = (ncells + 31) / 32;
        dense::View v{X, ngenes, ncells, col_major != 0};
        std::vector<uint32_t> mask((size_t)nthese * W), lvl((size_t)W, 0u);
        std::vector<int32_t> nzero((size_t)nthese);
        std::vector<double> det((size_t)ncells);
        dense::zero_mask(v, these, nthese, W, mask.data(), nzero.data(), det.data());
        for (int64_t c = 0; c < ncells; c++)
            if (cell_level2[c] != 0) lvl[c >> 5] |= 1u << (c & 31);
        ZeroGlmArgs a{};
        constexpr double MIN_PRIOR_SCALE = 1e-12;   // bayesglm's min.prior.scale
        a.scale[0] = o.prior_scale_for_intercept;
        a.scale[1] = std::max(o.prior_scale / predictor_scale(det), MIN_PRIOR_SCALE);
        a.scale[2] = o.prior_scale;                 // the 0/1 dummy: max - min = 1
        a.df = o.prior_df;
        a.epsilon = o.epsilon;
        a.maxit = o.maxit;
        a.W = W;
        a.ncells = (int)ncells;
        a.ngenes = (int)nthese;
        std::vector<int> devs = active_devices();
        CK(cudaSetDevice(devs[0]));
        cudaStream_t s;
        CK(cudaStreamCreate(&s));
        try {
            DevBuf<uint32_t> d_mask, d_lvl;
            DevBuf<int32_t> d_nz;
            DevBuf<double> d_det, d_p, d_detail;
            d_mask.alloc(mask.size()); d_lvl.alloc(lvl.size()); d_nz.alloc(nzero.size());
            d_det.alloc(det.size()); d_p.alloc((size_t)nthese);
            if (detail) d_detail.alloc((size_t)nthese * 8);
            cudaEvent_t t0, t1;
            CK(cudaEventCreate(&t0)); CK(cudaEventCreate(&t1));
            CK(cudaMemcpyAsync(d_mask.p, mask.data(), mask.size() * 4, cudaMemcpyHostToDevice, s));
            CK(cudaMemcpyAsync(d_lvl.p, lvl.data(), lvl.size() * 4, cudaMemcpyHostToDevice, s));
            CK(cudaMemcpyAsync(d_nz.p, nzero.data(), nzero.size() * 4, cudaMemcpyHostToDevice, s));
            CK(cudaMemcpyAsync(d_det.p, det.data(), det.size() * 8, cudaMemcpyHostToDevice, s));
            a.mask = d_mask.p; a.lvl = d_lvl.p; a.det = d_det.p; a.nzero = d_nz.p;
            a.pvalue = d_p.p; a.detail = detail ? d_detail.p : nullptr;
            cudaDeviceProp prop;
            CK(cudaGetDeviceProperties(&prop, devs[0]));
            const int threads = 128;   // 4 genes per CTA
            const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((nthese + 3) / 4, (int64_t)prop.multiProcessorCount * 16));
            CK(cudaEventRecord(t0, s));
            zero_glm_kernel<<<grid, threads, 0, s>>>(a);
            CK(cudaGetLastError());
            CK(cudaEventRecord(t1, s));
            CK(cudaMemcpyAsync(pvalue, d_p.p, (size_t)nthese * 8, cudaMemcpyDeviceToHost, s));
            if (detail) CK(cudaMemcpyAsync(detail, d_detail.p, (size_t)nthese * 64, cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            if (stats) {
                float ms = 0;
                CK(cudaEventElapsedTime(&ms, t0, t1));
                stats->total_ms = ms;
                stats->h2d_bytes = (double)(mask.size() + lvl.size() + nzero.size()) * 4 + (double)det.size() * 8;
                stats->d2h_bytes = (double)nthese * (detail ? 72 : 8);
            }
            cudaEventDestroy(t0); cudaEventDestroy(t1);
        } catch (...) {
            cudaStreamDestroy(s);
            throw;
        }
        CK(cudaStreamDestroy(s));
        return SCDD_OK;
    });
}

int32_t scdd_labels_to_dense(const double* X, int64_t ngenes, int64_t ncells, int32_t col_major, const int32_t* cell_cond01,
                             int32_t which, const int32_t* genes, int32_t nsel, const int64_t* gene_off,
                             const int32_t* labels, double* Z) {
    dense::View v;
    if (!dense_view(X, ngenes, ncells, col_major, cell_cond01, v) || nsel < 0 || (ngenes * ncells > 0 && !Z) ||
        (nsel > 0 && labels && !gene_off))
        return fail(SCDD_ERR_BAD_ARG, "bad arguments");
    if (!dense::labels_to_dense(v, cell_cond01, which, genes, nsel, gene_off, labels, Z))
        return fail(SCDD_ERR_BAD_ARG, "gene_off does not match the nonzero cells of the selected genes");
    return SCDD_OK;
}

int32_t scdd_cluster_counts(const int32_t* labels, const int32_t* cond01, const int64_t* gene_off, int32_t ngenes,
                            int32_t min_size, int32_t which, int32_t* counts) {
    if (ngenes < 0 || (ngenes > 0 && (!gene_off || !counts)) || (ngenes > 0 && gene_off[ngenes] > 0 && !labels) ||
        (which >= 0 && ngenes > 0 && gene_off[ngenes] > 0 && !cond01))
        return fail(SCDD_ERR_BAD_ARG, "bad arguments");
    dense::cluster_counts(labels, cond01, gene_off, ngenes, min_size, which, SCDD_GMAX, counts);
    return SCDD_OK;
}

// ---------------------------------------------------------------- resident plan
struct scdd_plan {
    Pipeline P;
    cudaStream_t own = nullptr;
};

scdd_plan* scdd_plan_create(int32_t device, const double* logy, const int64_t* gene_off, const int32_t* cond01,
                            int32_t ngenes, const scdd_cfg* cfg, int32_t chunk_perms, const uint32_t* seed6,
                            const double* cdr) {
    if (check_common(logy, gene_off, cond01, ngenes, cfg)) return nullptr;
    if (!seed6 || chunk_perms < 1 || ngenes < 1) { fail(SCDD_ERR_BAD_ARG, "plan needs seed6, chunk_perms >= 1, ngenes >= 1"); return nullptr; }
    scdd_plan* pl = nullptr;
    int rc = guarded([&]() -> int {
        pl = new scdd_plan();
        CK(cudaSetDevice(device));
        CK(cudaStreamCreate(&pl->own));
        pl->P.upload(device, logy, gene_off, cond01, ngenes, cfg, cdr, pl->own);
        pl->P.alloc_perm_chunk(chunk_perms, seed6, pl->own);
        pl->P.d_bf.alloc((size_t)ngenes * chunk_perms);
        CK(cudaStreamSynchronize(pl->own));
        return SCDD_OK;
    });
    if (rc) { delete pl; return nullptr; }
    return pl;
}

int32_t scdd_plan_run_chunk(scdd_plan* plan, void* cuda_stream, int32_t nperms) {
    if (!plan || nperms < 1 || nperms > plan->P.chunk_perms) return fail(SCDD_ERR_BAD_ARG, "bad plan / nperms");
    return guarded([&]() -> int {
        CK(cudaSetDevice(plan->P.device));
        plan->P.run_chunk_device_perms((cudaStream_t)cuda_stream, nperms, plan->P.d_bf.p, plan->P.chunk_perms, 0);
        return SCDD_OK;
    });
}

void* scdd_plan_bf_device_ptr(scdd_plan* plan) { return plan ? plan->P.d_bf.p : nullptr; }

int32_t scdd_plan_stats(scdd_plan* plan, scdd_stats* stats) {
    if (!plan || !stats) return fail(SCDD_ERR_BAD_ARG, "null");
    return guarded([&]() -> int {
        CK(cudaSetDevice(plan->P.device));
        CK(cudaDeviceSynchronize());
        std::memset(stats, 0, sizeof(*stats));
        plan->P.read_stats(stats, plan->own);
        return SCDD_OK;
    });
}

int32_t scdd_plan_gene_costs(scdd_plan* plan, double* cost) {
    if (!plan || !cost) return fail(SCDD_ERR_BAD_ARG, "null");
    if (!plan->P.have_order || plan->P.d_gene_cost.n < (size_t)plan->P.ngenes)
        return fail(SCDD_ERR_BAD_ARG, "no chunk has been measured yet (run one chunk first; blocks with fewer genes than 2 x SMs are not measured)");
    return guarded([&]() -> int {
        CK(cudaSetDevice(plan->P.device));
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(cost, plan->P.d_gene_cost.p, (size_t)plan->P.ngenes * sizeof(double), cudaMemcpyDeviceToHost));
        return SCDD_OK;
    });
}

void scdd_plan_destroy(scdd_plan* plan) {
    if (!plan) return;
    cudaSetDevice(plan->P.device);
    cudaDeviceSynchronize();
    if (plan->own) cudaStreamDestroy(plan->own);
    delete plan;
}

// ---------------------------------------------------------------- self tests / peak
int32_t scdd_selftest_x87_cumsum(const double* steps, int32_t n, double* out) {
    Ext80 run = ext_zero();
    for (int i = 0; i < n; i++) {
        run = ext_add(run, ext_from_double(steps[i]));
        out[i] = ext_to_double(run);
    }
    return SCDD_OK;
}

int32_t scdd_selftest_exp(const double* x, int32_t n, double* out) {
    // the kernels' softmax-term arithmetic with shift 0: out[i] = exp(x[i]).  (Inside the kernels the shift is
    // the per-point maximum, which cancels the range-reduction constant's error; with shift 0 that error,
    // |N| * 1.5e-19, shows up and is bounded by the test.)
    const double* tab = exp_table_host();
    for (int i = 0; i < n; i++) {
        const double t = term_t(x[i]);
        out[i] = term_e(term_s(term_r(x[i], t)), term_n(x[i], t), 0, tab);
    }
    return SCDD_OK;
}

// softmax of G log-terms exactly as the E-step computes it (integer shift, table exp, Newton reciprocal; a point whose
// every term is saturated is redone on a_k - max a_k, as the kernel's saturating pass does)
int32_t scdd_selftest_softmax(const double* a_in, int32_t G, double* z, double* logsumexp) {
    if (G < 1 || G > SCDD_GMAX) return fail(SCDD_ERR_BAD_ARG, "G out of range");
    const double* tab = exp_table_host();
    double a[SCDD_GMAX], t[SCDD_GMAX], s[SCDD_GMAX], e[SCDD_GMAX];
    int nn[SCDD_GMAX], M = 0;
    double far = 0.0;
    for (int k = 0; k < G; k++) a[k] = a_in[k];
    for (int pass = 0; pass < 2; pass++) {
        for (int k = 0; k < G; k++) {
            t[k] = term_t(a[k]); nn[k] = term_n(a[k], t[k]);
            M = (k == 0 || term_h(nn[k]) > M) ? term_h(nn[k]) : M;
        }
        if (M != expc::HSAT || pass == 1) break;
        far = a[0];
        for (int k = 1; k < G; k++) far = std::fmax(far, a[k]);
        for (int k = 0; k < G; k++) a[k] -= far;
    }
    double S = 0.0;
    for (int k = 0; k < G; k++) { s[k] = term_s(term_r(a[k], t[k])); e[k] = term_e(s[k], nn[k], M, tab); S += e[k]; }
    for (int k = 0; k < G; k++) z[k] = e[k] / S;
    if (logsumexp) *logsumexp = fast_log_tab(S, log_table_host()) + (double)M * expc::LN2_STEP + far;
    return SCDD_OK;
}

// the kernels' table-driven logarithm (host build of the same source): out[i] = log(x[i]) for positive normal x
int32_t scdd_selftest_log(const double* x, int32_t n, double* out) {
    if (n < 0 || (n > 0 && (!x || !out))) return fail(SCDD_ERR_BAD_ARG, "bad arguments");
    const LogTab* tab = log_table_host();
    for (int i = 0; i < n; i++) out[i] = fast_log_tab(x[i], tab);
    return SCDD_OK;
}

double scdd_measure_fp64_peak(int32_t device, double* sm_clock_mhz_out) {
    if (device_count_quiet() < 1) { fail(SCDD_ERR_NO_DEVICE, "no CUDA device"); return -1.0; }
    double result = -1.0;
    guarded([&]() -> int {
        CK(cudaSetDevice(device));
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, device));
        const int blocks = prop.multiProcessorCount * 4, threads = 512, iters = 1 << 14;
        DevBuf<double> out;
        out.alloc((size_t)blocks * threads);
        cudaEvent_t a, b;
        CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
        double best = 0.0;
        for (int rep = 0; rep < 6; rep++) {
            CK(cudaEventRecord(a));
            fp64_peak_kernel<<<blocks, threads>>>(out.p, iters, 1.0 + rep);
            CK(cudaEventRecord(b));
            CK(cudaEventSynchronize(b));
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, a, b));
            double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
            if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
        }
        cudaEventDestroy(a); cudaEventDestroy(b);
        if (sm_clock_mhz_out) {
            int khz = 0;
            cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
            *sm_clock_mhz_out = khz / 1000.0;
        }
        result = best;
        return SCDD_OK;
    });
    return result;
}

}  // extern "C"
