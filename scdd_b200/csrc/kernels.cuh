// kernels.cuh -- the __global__ kernels of libscdd_b200 (sm_100a).
//
//   prep_partition_kernel   per gene: positions of the reference-condition cells first (stable), n1
//   prep_lm_kernel          per gene: lm(y ~ C) fitted values and residuals (adjust.perms path)
//   permgen_kernel          per gene: B x sample(1:n) with R's L'Ecuyer-CMRG + rejection sampler
//   fit_kernel              persistent warps, one fit-set per warp at a time (fitset.cuh)
//   combine_kernel          bf.perm = JP(c1) + JP(c2) [- JP(oa)]
//   ks_kernel               per gene: stable sort of c(x, y), x87-exact cumsum, D
//   fp64_peak_kernel        DFMA throughput microbenchmark (roofline denominator)
#pragma once
#include <cstdint>

#include "../../include/scdd_b200.h"
#include "fitset.cuh"
#include "rrng.h"

namespace scdd {

// ------------------------------------------------------------------------------------------------
// gene preparation
// ------------------------------------------------------------------------------------------------
// part[off_g + j], j < n1_g: positions (within the gene's nonzero cells) of the cells with cond == 0,
// in cell order; then the cond != 0 cells.  new.y[cond == ref] of the reference is a gather through it.
__global__ void prep_partition_kernel(const int32_t* __restrict__ cond, const int64_t* __restrict__ off,
                                      int ngenes, uint32_t* __restrict__ part, int32_t* __restrict__ n1) {
    int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (g >= ngenes) return;
    int64_t o = off[g];
    int n = (int)(off[g + 1] - o);
    int c0 = 0;
    for (int i = lane; i < n; i += 32) c0 += (cond[o + i] == 0) ? 1 : 0;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) c0 += __shfl_xor_sync(FULLMASK, c0, s);
    if (lane == 0) n1[g] = c0;
    int r0 = 0, r1 = c0;
    for (int base = 0; base < n; base += 32) {
        int i = base + lane;
        bool v = i < n;
        int c = v ? cond[o + i] : 0;
        unsigned m0 = __ballot_sync(FULLMASK, v && c == 0);
        unsigned m1 = __ballot_sync(FULLMASK, v && c != 0);
        unsigned lt = (1u << lane) - 1u;
        if (v) {
            if (c == 0) part[o + r0 + __popc(m0 & lt)] = (uint32_t)i;
            else part[o + r1 + __popc(m1 & lt)] = (uint32_t)i;
        }
        r0 += __popc(m0);
        r1 += __popc(m1);
    }
}

// lm(y ~ C) per gene (R/permutations.R:232-234): ordinary least squares with intercept.
__global__ void prep_lm_kernel(const double* __restrict__ y, const double* __restrict__ C,
                               const int64_t* __restrict__ off, int ngenes,
                               double* __restrict__ fitted, double* __restrict__ resid) {
    int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (g >= ngenes) return;
    int64_t o = off[g];
    int n = (int)(off[g + 1] - o);
    double sx = 0.0, sy = 0.0;
    for (int i = lane; i < n; i += 32) { sx += C[o + i]; sy += y[o + i]; }
    sx = warp_sum(sx); sy = warp_sum(sy);
    double mx = sx / n, my = sy / n;
    double sxx = 0.0, sxy = 0.0;
    for (int i = lane; i < n; i += 32) {
        double dx = C[o + i] - mx;
        sxx = fma(dx, dx, sxx);
        sxy = fma(dx, y[o + i] - my, sxy);
    }
    sxx = warp_sum(sxx); sxy = warp_sum(sxy);
    double b1 = sxy / sxx, b0 = my - b1 * mx;
    for (int i = lane; i < n; i += 32) {
        double f = __dadd_rn(b0, __dmul_rn(b1, C[o + i]));
        fitted[o + i] = f;
        resid[o + i] = y[o + i] - f;
    }
}

// ------------------------------------------------------------------------------------------------
// permutations: thread per gene, nb consecutive sample(1:n) draws from the gene's own stream
// (BiocParallel gives every bplapply element its own L'Ecuyer-CMRG stream, R/scDD.R:458).
// perm_out layout: gene g, permutation b (local), element i -> perm_out[off_g * nb + b * n_g + i], 0-based.
// ------------------------------------------------------------------------------------------------
template <class IdxT>
__global__ void permgen_kernel(const int64_t* __restrict__ off, int ngenes, uint32_t* __restrict__ seeds,
                               IdxT* __restrict__ scratch, IdxT* __restrict__ perm_out, int nb) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ngenes) return;
    int64_t o = off[g];
    uint32_t n = (uint32_t)(off[g + 1] - o);
    LecuyerGen gen;
#pragma unroll
    for (int j = 0; j < 6; j++) gen.st.s[j] = seeds[(size_t)g * 6 + j];
    IdxT* sc = scratch + o;
    IdxT* out = perm_out + (size_t)o * nb;
    for (int b = 0; b < nb; b++) {
        r_sample_perm(gen, n, sc, out);
        out += n;
    }
#pragma unroll
    for (int j = 0; j < 6; j++) seeds[(size_t)g * 6 + j] = gen.st.s[j];
}

// ------------------------------------------------------------------------------------------------
// permutations, a group of GS lanes per gene (GS = 32, 8 or 4; 32 / GS genes per warp).
//
// R's stream is serial per gene, but only the Fisher-Yates consumption has to be: the MRG32k3a outputs themselves
// are produced GS at a time.  Decimating a linear recurrence of order 3 by GS gives another order-3 recurrence,
// z[j+3] = b2 z[j+2] + b1 z[j+1] + b0 z[j] (mod m), whose coefficients are those of the characteristic polynomial of
// A^GS (computed on the host); lane l of the group produces outputs l, l+GS, l+2GS, ... of the stream.  The group's
// leader lane then consumes the 16-bit chunks exactly as R_unif_index / do_sample do, with the Fisher-Yates array in
// shared memory, and the finished permutation is written out by the whole group.  With GS < 32 the leaders of the
// 32 / GS groups of a warp run the serial loop in lock-step, each on its own gene: the issue slots the loop costs
// (it is issue-bound: ~25 instructions per candidate) are shared by 4 or 8 genes instead of one.
// Bit-identical to permgen_kernel and to R.
//
// seed_mode 0: seeds = [ngenes*6], gene g continues its own stream over the permutations of all chunks (the advanced
//              state is written back) -- BiocParallel's stream per bplapply element, parallelBy = "Genes".
// seed_mode 1: seeds = [B*6], permutation b of every gene starts stream b afresh -- bplapply(1:nperms) inside
//              permMclust / permMclustCov, parallelBy = "Permutations" (R/permutations.R:177, :99).
//
// Constant subsets (R/mclust.restricted.R:48-50): when all values of a permuted subset are identical the reference
// adds runif(n, -0.1, 0.1), drawn from the SAME stream between this permutation's sample() and the next one's.  The
// leader therefore tests each subset of the permutation it just drew (first two values, then the rest: one compare
// per permutation on untied data), and for a constant subset records the stream state at that point in
// jit_state[item] -- fit_kernel regenerates the n uniforms from it -- and skips n outputs of its own stream.
// ------------------------------------------------------------------------------------------------
struct MrgDecim {
    uint32_t b1[3], b2[3];   // {b0, b1, b2} for the two components (moduli MRG_M1, MRG_M2)
};

constexpr int PG_BLOCK = 256;   // uniforms generated per refill and gene
constexpr uint32_t JIT_NONE = 0xffffffffu;   // jit_state[6 * item]: no jitter recorded (not a valid MRG component)

// x mod (2^32 - c) for the two MRG moduli (c = 209, 22853): 2^32 = c (mod m), so hi * 2^32 + lo = hi * c + lo
__device__ __forceinline__ uint64_t mrg_fold(uint64_t x, uint32_t c) { return (x >> 32) * c + (x & 0xffffffffULL); }
__device__ __forceinline__ uint32_t mrg_rec3(const uint32_t* co, uint32_t z0, uint32_t z1, uint32_t z2, uint32_t c) {
    // each folded product < 2^32 (c + 1); their sum < 2^49; second fold < 2^33; third < 2^32 + c
    uint64_t s = mrg_fold((uint64_t)co[0] * z0, c) + mrg_fold((uint64_t)co[1] * z1, c) + mrg_fold((uint64_t)co[2] * z2, c);
    s = mrg_fold(mrg_fold(s, c), c);
    const uint64_t m = 4294967296ULL - c;
    if (s >= m) s -= m;
    return (uint32_t)s;
}
constexpr uint32_t MRG_C1 = (uint32_t)(4294967296LL - MRG_M1), MRG_C2 = (uint32_t)(4294967296LL - MRG_M2);

__host__ __device__ constexpr size_t permgen_smem_per_group(int nmax, int idx_bytes, int gs) {
    return ((size_t)nmax * idx_bytes + 15) / 16 * 16 + PG_BLOCK * sizeof(uint16_t) + (size_t)6 * gs * sizeof(uint32_t);
}

struct PermgenArgs {
    const int64_t* off;
    int ngenes;
    uint32_t* seeds;
    int seed_mode;            // 0: per gene (in/out), 1: per permutation (read-only, indexed perm_b0 + b)
    int perm_b0;
    void* perm_out;           // gene g, local permutation b, element i -> [off_g * nb + b * n_g + i], 0-based
    int nb;
    MrgDecim dc;              // decimation by the group size
    int nmax;
    // constant-subset jitter; jit_state == nullptr switches the test off
    const double* vals;
    const double* fitted;     // adjusted path: the permuted quantity is log(exp(fitted[i] + vals[perm[i]]))
    const uint32_t* part;
    const int32_t* n1;
    int sets;
    uint32_t* jit_state;      // [ngenes * nb * sets * 6], pre-set to JIT_NONE
};

// the value the fit kernel will see at position j (partition order) of a subset -- same expression as fit_kernel's gather
__device__ __forceinline__ double permuted_value(const double* vals, const double* fitted, int64_t o, uint32_t i, uint32_t src) {
    double v = vals[o + src];
    if (fitted) v = log(exp(fitted[o + i] + v));   // R/permutations.R:244,251-252
    return v;
}

// BATCH: the GS lanes of a group consume GS candidates of the stream per round (see the comment in the loop) instead of
// the leader alone one by one.  Measured (profiles/permgen_batched_r02.txt): slower when many groups share an SM
// (config 3: 8-lane groups, ~40 genes per SM hide the leader's dependent chain), so it is used for the 32-lane groups
// that long genes force (config 5: a 50,000-cell gene fills 100 KB, two groups per SM).
template <class IdxT, int GS, bool BATCH>
__global__ void permgen_group_kernel(PermgenArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int GPW = 32 / GS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int sub = lane / GS, sl = lane % GS, lead = lane - sl;
    const bool leader = sl == 0;
    unsigned char* base = smem_raw + (size_t)(warp * GPW + sub) * permgen_smem_per_group(a.nmax, (int)sizeof(IdxT), GS);
    IdxT* arr = reinterpret_cast<IdxT*>(base);
    uint16_t* cand = reinterpret_cast<uint16_t*>(base + ((size_t)a.nmax * sizeof(IdxT) + 15) / 16 * 16);
    uint32_t* ini = reinterpret_cast<uint32_t*>(cand + PG_BLOCK);   // [2][3 * GS] raw outputs of the first 3 GS steps
    const int groups_total = gridDim.x * wpb * GPW;

    for (int gb = (blockIdx.x * wpb + warp) * GPW; gb < a.ngenes; gb += groups_total) {   // warp-uniform
        const int g = gb + sub;
        const bool act = g < a.ngenes;
        int64_t o = 0;
        uint32_t n = 0;
        if (act) { o = a.off[g]; n = (uint32_t)(a.off[g + 1] - o); }
        int nh0 = 0;
        if (act && a.jit_state) nh0 = a.n1[g];
        uint32_t x0 = 0, x1 = 0, x2 = 0, y0 = 0, y1 = 0, y2 = 0;   // this lane's window of the decimated sequences
        uint32_t xl = 0, yl = 0;                                  // raw values of the lane's last generated output
        uint32_t bs[6] = {0, 0, 0, 0, 0, 0};    // stream state at the start of the block held in cand (leader)
        uint32_t nbs[6] = {0, 0, 0, 0, 0, 0};   // ... of the block after it
        int used = PG_BLOCK;
        bool have_block = false;

        // stream start: the leader steps 3 GS outputs serially, the lanes pick up their windows
        auto init_stream = [&](const uint32_t* seed) {
            __syncwarp();
            if (leader && act) {
                Lecuyer st;
#pragma unroll
                for (int j = 0; j < 6; j++) { st.s[j] = seed[j]; bs[j] = seed[j]; }
                for (int t = 0; t < 3 * GS; t++) {
                    lecuyer_unif(st);
                    ini[t] = st.s[2];
                    ini[3 * GS + t] = st.s[5];
                }
            }
            __syncwarp();
            if (act) {
                x0 = ini[sl]; x1 = ini[GS + sl]; x2 = ini[2 * GS + sl];
                y0 = ini[3 * GS + sl]; y1 = ini[4 * GS + sl]; y2 = ini[5 * GS + sl];
            }
            used = PG_BLOCK;
            have_block = false;
            __syncwarp();
        };
        // next PG_BLOCK outputs of the groups that ask for them (want is group-uniform; called warp-uniformly)
        auto refill = [&](bool want) {
            if (!__any_sync(FULLMASK, want)) return;
            if (want) {
#pragma unroll 1
                for (int j = 0; j < PG_BLOCK / GS; j++) {
                    const int t = GS * j + sl;
                    const int64_t p1 = (int64_t)x0, p2 = (int64_t)y0;
                    const int64_t d = (p1 > p2) ? (p1 - p2) : (p1 - p2 + MRG_M1);
                    const double u = r_fixup((double)d * 2.328306549295727688e-10);
                    cand[t] = (uint16_t)(int)floor(u * 65536.0);
                    xl = x0; yl = y0;
                    const uint32_t xn = mrg_rec3(a.dc.b1, x0, x1, x2, MRG_C1);
                    const uint32_t yn = mrg_rec3(a.dc.b2, y0, y1, y2, MRG_C2);
                    x0 = x1; x1 = x2; x2 = xn;
                    y0 = y1; y1 = y2; y2 = yn;
                }
            }
            __syncwarp();
            // the state after this block = its last three raw outputs (lanes GS-3 .. GS-1 of the last round)
            uint32_t t6[6];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                t6[k] = __shfl_sync(FULLMASK, xl, lead + GS - 3 + k);
                t6[3 + k] = __shfl_sync(FULLMASK, yl, lead + GS - 3 + k);
            }
            if (want) {
#pragma unroll
                for (int k = 0; k < 6; k++) { if (have_block) bs[k] = nbs[k]; nbs[k] = t6[k]; }
                have_block = true;
                used = 0;
            }
        };
        // state after `used` outputs of the current block (leader; at most PG_BLOCK serial steps, rarely needed)
        auto state_now = [&](uint32_t* out6) {
            Lecuyer st;
#pragma unroll
            for (int j = 0; j < 6; j++) st.s[j] = (used == PG_BLOCK && have_block) ? nbs[j] : bs[j];
            if (!(used == PG_BLOCK && have_block))
                for (int t = 0; t < used && have_block; t++) lecuyer_unif(st);
#pragma unroll
            for (int j = 0; j < 6; j++) out6[j] = st.s[j];
        };

        if (a.seed_mode == 0) init_stream(a.seeds + (size_t)(act ? g : 0) * 6);
        IdxT* gout = reinterpret_cast<IdxT*>(a.perm_out) + (size_t)o * a.nb;
        for (int b = 0; b < a.nb; b++) {
            if (a.seed_mode != 0) init_stream(a.seeds + (size_t)(a.perm_b0 + b) * 6);
            for (uint32_t i = sl; i < n; i += GS) arr[i] = (IdxT)i;
            __syncwarp();
            uint32_t i = 0, m = n;
            int bits = index_bits(m);
            int need = bits / 16 + 1;
            uint64_t vacc = 0;
            if constexpr (BATCH) {
                const unsigned grp = (GS == 32) ? FULLMASK : (((1u << (GS & 31)) - 1u) << lead);   // this group's lanes
                while (__any_sync(FULLMASK, act && i < n)) {
                    refill(act && i < n && used == PG_BLOCK);
                    // ---- the candidates of the current block ----
                    while (__any_sync(FULLMASK, act && i < n && used < PG_BLOCK)) {
                        const bool live = act && i < n && used < PG_BLOCK;
                        // (1) A batch of GS candidates at once, one per lane of the group, whenever every decision of the
                        // batch is certain without knowing the others: R accepts candidate v iff v < m, and m drops by one
                        // per accepted pick, so lane l's candidate is surely accepted if v + l < m and surely rejected if
                        // v >= m.  The batch must not reach a power of two (the bit count of R_unif_index changes there),
                        // must not finish the permutation (candidates after the last pick are not drawn), and its swaps
                        // must not touch a location twice -- then "all loads, then all stores" equals the serial order.
                        const int cpc = bits / 16 + 1;                                  // 16-bit chunks per candidate
                        const int navail = (PG_BLOCK - used) / cpc;
                        const uint32_t P = m > 1 ? (1u << (31 - __clz((int)(m - 1)))) : 0u;   // largest power of two below m
                        const bool can = live && need == cpc && bits <= 31 && navail >= 1 && m > 2u * GS && m - (uint32_t)GS > P;
                        const int nbat = min(GS, navail);
                        const bool have = can && sl < nbat;
                        uint32_t v = 0xffffffffu;
                        if (have) {
                            const uint16_t* cp = cand + used + sl * cpc;
                            v = cp[0];
                            if (cpc == 2) v = (v << 16) | cp[1];
                            v &= (bits >= 32) ? 0xffffffffu : ((1u << bits) - 1u);
                        }
                        const bool acc = have && v < m;
                        const bool unsure = acc && !(v + (uint32_t)sl < m);
                        const unsigned b_acc = __ballot_sync(FULLMASK, acc) & grp;
                        const int na = __popc(b_acc);
                        const uint32_t my_m = m - (uint32_t)__popc(b_acc & ((1u << lane) - 1u));   // m when this lane's pick is made
                        const unsigned same = __match_any_sync(FULLMASK, acc ? v : (0x80000000u | (unsigned)lane));
                        const bool hazard = unsure || (acc && (__popc(same & grp) > 1 || (v + (uint32_t)na >= m && v + 2u <= my_m)));
                        const unsigned b_haz = __ballot_sync(FULLMASK, hazard);   // outside the &&: every lane must vote
                        const bool go = can && (b_haz & grp) == 0;
                        IdxT pick = 0, last = 0;
                        if (go && acc) { pick = arr[v]; last = arr[my_m - 1]; }
                        __syncwarp();
                        if (go && acc) { arr[v] = last; arr[my_m - 1] = pick; }   // y[i] = x[j]; x[j] = x[--m]: the array ends reversed
                        __syncwarp();
                        if (go) { i += (uint32_t)na; m -= (uint32_t)na; used += nbat * cpc; }
                        // (2) otherwise the leader takes up to 2 GS chunks one by one, exactly as R_unif_index / do_sample do
                        if (leader && live && !go) {
                            int budget = 2 * GS;
                            if (bits <= 15) {
                                // ceil(log2 m) <= 15: one 16-bit chunk per attempt, everything fits 32-bit arithmetic
                                uint32_t mask = (1u << bits) - 1u;
                                while (i < n && used < PG_BLOCK && budget-- > 0) {
                                    const uint32_t c = (uint32_t)cand[used++] & mask;
                                    if (c < m) {
                                        const IdxT pk = arr[c];
                                        --m;
                                        arr[c] = arr[m];
                                        arr[m] = pk;
                                        i++;
                                        if ((m & (m - 1)) == 0) {   // ceil(log2 m) changes only at powers of two
                                            bits = index_bits(m);
                                            mask = (1u << bits) - 1u;
                                        }
                                    }
                                }
                                need = bits / 16 + 1;
                            } else {
                                const uint64_t one = 1;
                                uint64_t mask = (bits >= 64) ? ~0ULL : ((one << bits) - one);
                                while (i < n && used < PG_BLOCK && bits > 15 && budget-- > 0) {
                                    vacc = 65536ULL * vacc + (uint64_t)cand[used++];
                                    if (--need == 0) {
                                        const uint64_t c = vacc & mask;
                                        if (c < (uint64_t)m) {
                                            const IdxT pk = arr[c];
                                            --m;
                                            arr[c] = arr[m];
                                            arr[m] = pk;
                                            i++;
                                            if ((m & (m - 1)) == 0) {
                                                bits = index_bits(m);
                                                mask = (bits >= 64) ? ~0ULL : ((one << bits) - one);
                                            }
                                        }
                                        vacc = 0;
                                        need = bits / 16 + 1;
                                    }
                                }
                            }
                        }
                        __syncwarp();
                        i = __shfl_sync(FULLMASK, i, lead);
                        m = __shfl_sync(FULLMASK, m, lead);
                        used = __shfl_sync(FULLMASK, used, lead);
                        bits = __shfl_sync(FULLMASK, bits, lead);
                        need = __shfl_sync(FULLMASK, need, lead);
                    }
                }
            } else {
                while (__any_sync(FULLMASK, act && i < n)) {
                    refill(act && i < n && used == PG_BLOCK);
                    if (leader && act && i < n) {
                        // ---- R_unif_index rejection sampling + the swap of do_sample, serial ----
                        if (bits <= 15) {
                            // ceil(log2 m) <= 15 (m <= 32768; bits only shrinks while a permutation is drawn, so a gene of up
                            // to 65535 cells spends three quarters of its picks here): one 16-bit chunk per attempt,
                            // everything fits 32-bit arithmetic
                            uint32_t mask = (1u << bits) - 1u;
                            // y[i] = x[j]; x[j] = x[--m]  (do_sample).  The slot freed at the end of the active range
                            // receives y[i]: the array ends up holding the permutation reversed.
                            auto take = [&](uint32_t c16) {
                                const uint32_t v = c16 & mask;
                                if (v < m) {
                                    const IdxT pick = arr[v];
                                    --m;
                                    arr[v] = arr[m];
                                    arr[m] = pick;
                                    i++;
                                    if ((m & (m - 1)) == 0) {   // ceil(log2 m) changes only at powers of two
                                        bits = index_bits(m);
                                        mask = (1u << bits) - 1u;
                                    }
                                }
                            };
                            // four candidates per shared-memory load while a whole aligned quad is available: the loop is
                            // latency-bound (few genes per SM fit shared memory), and the candidate load is half of the
                            // dependent chain of a step
                            while (i + 4 <= n && (used & 3) == 0 && used + 4 <= PG_BLOCK) {
                                const uint2 q = *reinterpret_cast<const uint2*>(cand + used);
                                used += 4;
                                take(q.x & 0xffffu); take(q.x >> 16); take(q.y & 0xffffu); take(q.y >> 16);
                            }
                            while (i < n && used < PG_BLOCK) take((uint32_t)cand[used++]);
                        } else {
                            const uint64_t one = 1;
                            uint64_t mask = (bits >= 64) ? ~0ULL : ((one << bits) - one);
                            while (i < n && used < PG_BLOCK && bits > 15) {
                                vacc = 65536ULL * vacc + (uint64_t)cand[used++];
                                if (--need == 0) {
                                    uint64_t v = vacc & mask;
                                    if (v < (uint64_t)m) {
                                        const IdxT pick = arr[v];
                                        --m;
                                        arr[v] = arr[m];
                                        arr[m] = pick;
                                        i++;
                                        if ((m & (m - 1)) == 0) {
                                            bits = index_bits(m);
                                            mask = (bits >= 64) ? ~0ULL : ((one << bits) - one);
                                        }
                                    }
                                    vacc = 0;
                                    need = bits / 16 + 1;
                                }
                            }
                        }
                    }
                    i = __shfl_sync(FULLMASK, i, lead);
                    used = __shfl_sync(FULLMASK, used, lead);
                    __syncwarp();
                }
            }
            for (uint32_t k = sl; k < n; k += GS) gout[(size_t)b * n + k] = arr[n - 1 - k];
            __syncwarp();
            if (a.jit_state) {
                // subsets in the order the reference fits them: (oa,) c1, c2 -- R/permutations.R:261-263, :278-279
                for (int q = 0; q < a.sets; q++) {
                    const int w = (a.sets == 3) ? (q == 0 ? 2 : q - 1) : q;   // fit_kernel's set index: 0 = c1, 1 = c2, 2 = oa
                    uint32_t skip = 0;
                    if (leader && act) {
                        const int nh = (w == 0) ? nh0 : (w == 1 ? (int)n - nh0 : (int)n);
                        const int pofs = (w == 1) ? nh0 : 0;
                        bool constant = nh >= 1;
                        double v0 = 0.0;
                        for (int j = 0; j < nh && constant; j++) {
                            const uint32_t ci = a.part[o + pofs + j];
                            const double v = permuted_value(a.vals, a.fitted, o, ci, (uint32_t)arr[n - 1 - ci]);
                            if (j == 0) v0 = v; else if (v != v0) constant = false;
                        }
                        if (constant) {
                            state_now(a.jit_state + (((size_t)g * a.nb + b) * a.sets + w) * 6);
                            skip = (uint32_t)nh;
                        }
                    }
                    skip = __shfl_sync(FULLMASK, skip, lead);
                    while (__any_sync(FULLMASK, skip > 0)) {
                        refill(skip > 0 && used == PG_BLOCK);
                        if (skip > 0) {
                            const uint32_t take = min(skip, (uint32_t)(PG_BLOCK - used));
                            used += (int)take;
                            skip -= take;
                        }
                    }
                }
            }
        }
        // ---- state after exactly the consumed outputs ----
        if (a.seed_mode == 0 && leader && act) state_now(a.seeds + (size_t)g * 6);
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// Per-gene sorted copy of the values with the original position of every element (one CTA per gene, bitonic sort of
// (value, index) pairs in shared memory).  A permuted subset of a gene is a sub-multiset of the SAME values, so its
// sorted order can be read off this copy with a stable partition instead of sorting the subset again for every
// permutation (fit_kernel, "presorted").  Genes beyond PRESORT_MAX_CELLS are skipped (they run on CTA teams).
// ------------------------------------------------------------------------------------------------
constexpr int PRESORT_MAX_CELLS = 4096;   // a warp-team launch has subsets <= 2048 points, hence genes <= 4096 cells
__global__ void prep_sort_kernel(const double* __restrict__ vals, const int64_t* __restrict__ off, int ngenes,
                                 double* __restrict__ srt_vals, uint32_t* __restrict__ srt_idx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int g = blockIdx.x;
    if (g >= ngenes) return;
    const int64_t o = off[g];
    const int n = (int)(off[g + 1] - o);
    if (n > PRESORT_MAX_CELLS || n < 1) return;
    int np = 2;
    while (np < n) np <<= 1;
    double* key = reinterpret_cast<double*>(smem_raw);
    uint32_t* idx = reinterpret_cast<uint32_t*>(key + np);
    for (int j = threadIdx.x; j < np; j += blockDim.x) { key[j] = j < n ? vals[o + j] : INFINITY; idx[j] = (uint32_t)j; }
    __syncthreads();
    for (int k = 2; k <= np; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < np; i += blockDim.x) {
                const int l = i ^ j;
                if (l > i) {
                    const double a = key[i], b = key[l];
                    const bool up = (i & k) == 0;
                    // ties broken by index: a total order, so the result does not depend on the network's schedule
                    const bool gt = a > b || (a == b && idx[i] > idx[l]);
                    if (gt == up) {
                        key[i] = b; key[l] = a;
                        const uint32_t t = idx[i]; idx[i] = idx[l]; idx[l] = t;
                    }
                }
            }
            __syncthreads();
        }
    for (int j = threadIdx.x; j < n; j += blockDim.x) { srt_vals[o + j] = key[j]; srt_idx[o + j] = idx[j]; }
}

// ------------------------------------------------------------------------------------------------
// mixture fits
// ------------------------------------------------------------------------------------------------
enum PermKind { PERM_NONE = 0, PERM_I32_1BASED = 1, PERM_U16 = 2, PERM_U32 = 3 };

// compute-sanitizer is closed on the build pool, so index checks are compiled in on demand
// (-DSCDD_DEBUG_BOUNDS): a failed check bumps stats[9] and the tests assert it stays 0.
#ifdef SCDD_DEBUG_BOUNDS
#define SCDD_CHECK(stats, cond) do { if (!(cond)) atomicAdd(&(stats)[9], 1.0); } while (0)
#else
#define SCDD_CHECK(stats, cond) do { } while (0)
#endif

struct FitArgs {
    const double* vals;        // logy, or lm residuals on the adjusted path
    const double* fitted;      // nullptr, or lm fitted values (adjusted path): x = log(exp(fitted + resid[perm]))
    const int64_t* gene_off;
    const uint32_t* part;
    const int32_t* n1;
    const void* perm;          // nullptr = identity (observed data)
    long long perm_bstride;    // permutations per gene stored in `perm`
    int perm_kind;
    int perm_b0;               // first permutation of this launch inside `perm`
    int ngenes, nb, sets;      // sets per (gene, permutation): 2 = (c1, c2), 3 = (c1, c2, oa)
    int np_max;                // slab capacity (power of two)
    double* xs_global;         // CTA teams only: per-team slabs in global memory when np_max exceeds shared memory
    int sort_tile;             // ... then the sort runs through a shared-memory tile of this many doubles per team (0: none)
    const int32_t* gene_order; // nullable: genes in decreasing expected cost (longest-processing-time-first)
    double* gene_cost;         // nullable: [ngenes] accumulated EM work (triples) per gene, feeds gene_order
    double* jp;                // [ngenes * nb * sets] log joint posterior of each fit-set
    scdd_fit* fits;            // observed-data detail outputs (nullable): [3 * ngenes] as (oa, c1, c2)
    int32_t* cls_oa;
    int32_t* cls_c;
    DevCfg cfg;
    unsigned long long* work_counter;
    double* stats;             // [9] global accumulators (WarpStats order)
    // Two-pass launches (observed fits): pass 1 runs warp teams with an EM iteration budget and marks the fit-sets that
    // exceed it; pass 2 (only_deferred) re-runs just those with one CTA per fit-set.  A launch is otherwise as long as
    // its slowest fit-set on ONE warp (thousands of EM iterations for ~1 % of the observed fit-sets).
    int iter_budget;           // 0 = none
    int only_deferred;
    unsigned char* deferred;   // [ngenes * nb * sets], nullable
    // Constant subsets (R/mclust.restricted.R:48-50): stream state at which the reference draws runif(n, -0.1, 0.1),
    // recorded by permgen_group_kernel per fit-set (JIT_NONE in word 0 = none); nullable (observed data, host-drawn
    // permutations: a constant vector then has no valid model)
    const uint32_t* jit_state;
    // Per-gene sorted values and the original position of each (prep_sort_kernel); nullable.  Warp teams on the
    // unadjusted path build the sorted subset from them by a stable partition instead of a sort ("presorted").
    const double* srt_vals;
    const uint32_t* srt_idx;
};

__device__ __forceinline__ uint32_t perm_lookup(const FitArgs& a, int64_t o, int n, int b, uint32_t i) {
    if (a.perm_kind == PERM_NONE) return i;
    size_t at = (size_t)o * (size_t)a.perm_bstride + (size_t)(a.perm_b0 + b) * (size_t)n + i;
    if (a.perm_kind == PERM_U16) return (uint32_t)((const uint16_t*)a.perm)[at];
    if (a.perm_kind == PERM_U32) return ((const uint32_t*)a.perm)[at];
    return (uint32_t)(((const int32_t*)a.perm)[at] - 1);
}

template <int G>
__device__ __noinline__ int label_of(const Row& r, double x) {
    ParamsReg<G, true> p;
#pragma unroll
    for (int k = 0; k < G; k++) { p.mu[k] = r.mu_em[k]; p.nh[k] = r.nh_em[k]; p.lc[k] = r.lc_em[k]; }
    return map_label<G>(x, p);
}

#ifndef SCDD_FIT_WARPS
#define SCDD_FIT_WARPS 16
#endif
constexpr int FIT_MAX_WARPS = SCDD_FIT_WARPS;

// uploads the 2^(j/2^TB) table to the current device (once per device per process)
inline cudaError_t init_exp_table() {
    cudaError_t e = cudaMemcpyToSymbol(EXP_TABLE_GLOBAL, exp_table_host(), SCDD_EXP_TABLE_SIZE * sizeof(double));
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(LOG_TABLE_GLOBAL, log_table_host(), SCDD_LOG_TABLE_SIZE * sizeof(LogTab));
}
constexpr int FIT_CTA_WARPS = SCDD_FIT_WARPS;     // team size of the CTA-per-fit-set variant (few, very large fit-sets)
constexpr int FIT_SUB_WARPS = 4;                   // team size of the default multi-warp variant: 4 teams per CTA
constexpr int FIT_WARP_TEAM_MAX_POINTS = 2048;
constexpr int FIT_TAB_BYTES = SCDD_EXP_TABLE_SIZE * 8 + SCDD_LOG_TABLE_SIZE * 16;   // the CTA's 2^(j/2048) table (16 KB of static shared memory)    // larger fit-sets are run by a whole CTA

// per-team shared-memory region: {mu, nh, lc, pad} x GMAX | EmCtx | Row[GMAX] | WarpStats | cross-warp reduction
// scratch (CTA teams) | xs[np] (absent when the vector lives in global memory)
constexpr int PRESORT_FLAG_BYTES = PRESORT_MAX_CELLS / 8;   // one membership bit per cell of the gene (warp teams)
__host__ __device__ constexpr size_t fit_region_head(int tw) {
    return 4 * GMAX * sizeof(double) + sizeof(EmCtx) + GMAX * sizeof(Row) + sizeof(WarpStats) +
           (tw > 1 ? (size_t)tw * 16 * sizeof(double) : (size_t)PRESORT_FLAG_BYTES);
}
__host__ __device__ constexpr size_t fit_smem_per_warp(int np) {
    return fit_region_head(1) + (size_t)np * sizeof(double);
}
static_assert(fit_region_head(1) % 16 == 0 && fit_region_head(FIT_CTA_WARPS) % 16 == 0 &&
              fit_region_head(FIT_SUB_WARPS) % 16 == 0, "xs must stay 16-byte aligned");
// bytes of one team's region when the CTA holds 16 / tw teams: head, plus the slab when it is kept in shared memory
__host__ __device__ constexpr size_t fit_team_region(int tw, int np, bool slab_in_smem) {
    return fit_region_head(tw) + (slab_in_smem ? (size_t)np * sizeof(double) : 0);
}

// BUDGET: pass 1 of a two-pass launch (see FitArgs::iter_budget).  A separate instantiation, so that the kernel of the
// permutation chunks is compiled exactly as if the budget logic did not exist.
template <int TW, bool BUDGET = false>
__global__ void __launch_bounds__(FIT_MAX_WARPS * 32, 1) fit_kernel(FitArgs a) {
    typedef Team<TW> T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tid = T::tid(lane);
    for (int i = threadIdx.x; i < SCDD_EXP_TABLE_SIZE; i += blockDim.x) EXP_TABLE_SH[i] = EXP_TABLE_GLOBAL[i];   // 2^(j/2048)
    for (int i = threadIdx.x; i < SCDD_LOG_TABLE_SIZE; i += blockDim.x) LOG_TABLE_SH[i] = LOG_TABLE_GLOBAL[i];
    __syncthreads();
    const int team = T::index(), teams = FIT_MAX_WARPS / TW;
    const size_t stride = (TW == 1 || a.xs_global == nullptr) ? fit_team_region(TW, a.np_max, true)
                                                              : fit_region_head(TW) + (size_t)a.sort_tile * sizeof(double);
    unsigned char* region = smem_raw + (size_t)team * stride;
    double* pm = reinterpret_cast<double*>(region);               // {mu, nh, lc, pad} x GMAX, 16-byte aligned
    Row* rows = reinterpret_cast<Row*>(reinterpret_cast<unsigned char*>(pm + 4 * GMAX) + sizeof(EmCtx));   // EmCtx sits between them
    WarpStats* st = reinterpret_cast<WarpStats*>(rows + GMAX);
    double* red = reinterpret_cast<double*>(st + 1);
    double* xs = (TW > 1 && a.xs_global) ? a.xs_global + ((size_t)blockIdx.x * teams + team) * a.np_max
                                         : reinterpret_cast<double*>(region + fit_region_head(TW));
    const bool lead0 = T::lead() && lane == 0;
    if (lead0) {
        *st = WarpStats{0, 0, 0, 0, 0, 0, 0, 0, 0};
        // slab in global memory: the E-step stages it through the sort's tile, 2 buffers x stage_k points per thread
        EmCtx* cx = reinterpret_cast<EmCtx*>(pm + 4 * GMAX);
        const int k = (TW > 1 && a.xs_global) ? a.sort_tile / (2 * T::NT) : 0;
        cx->stage = k > 0 ? reinterpret_cast<double*>(region + fit_region_head(TW)) : nullptr;
        cx->stage_k = k > STAGE_K ? STAGE_K : k;
        cx->pad = 0;
    }
    T::sync();
    const long long total = (long long)a.ngenes * a.nb * a.sets;
    const double NaN = nan("");

    for (;;) {
        long long item = 0;
        if (TW == 1) {
            if (lane == 0) item = (long long)atomicAdd(a.work_counter, 1ULL);
            item = __shfl_sync(FULLMASK, item, 0);
        } else {
            T::sync();
            if (tid == 0) *reinterpret_cast<long long*>(red) = (long long)atomicAdd(a.work_counter, 1ULL);
            T::sync();
            item = *reinterpret_cast<long long*>(red);
            T::sync();
        }
        if (item >= total) break;
        if (a.only_deferred && !a.deferred[item]) continue;
        const int w = (int)(item % a.sets);
        const long long u = item / a.sets;
        const int b = (int)(u % a.nb);
        const int g = a.gene_order ? a.gene_order[(int)(u / a.nb)] : (int)(u / a.nb);
        const int64_t o = a.gene_off[g];
        const double work0 = st->triples;
        const int ng = (int)(a.gene_off[g + 1] - o);
        const int n1 = a.n1[g];
        int n, pofs;
        if (w == 0) { n = n1; pofs = 0; }
        else if (w == 1) { n = ng - n1; pofs = n1; }
        else { n = ng; pofs = 0; }                       // oa on c(y1, y2)

        // ---- the (permuted) subset, sorted, in the slab (padded with +Inf to a power of two) ----
        int np = 2;
        while (np < n) np <<= 1;
        if (TW == 1 && a.srt_vals != nullptr && a.fitted == nullptr && ng <= PRESORT_MAX_CELLS) {
            // presorted: the subset is a sub-multiset of the gene's values, whose sorted order is known.  Mark the
            // ORIGINAL positions that this permutation sends into the subset (one bit per cell), then walk the gene's
            // sorted copy and keep the marked elements in order: O(n) coalesced work instead of a bitonic sort per
            // fit-set, and bit-identical to it (equal values are interchangeable).
            if (w == 2) {
                for (int j = lane; j < np; j += 32) xs[j] = j < n ? a.srt_vals[o + j] : INFINITY;
            } else {
                uint32_t* flag = reinterpret_cast<uint32_t*>(red);
                for (int k = lane; k < (ng + 31) / 32; k += 32) flag[k] = 0u;
                __syncwarp();
                // four independent load chains per lane and step (the loops are bound by L2 latency, not by work)
                constexpr int UF = 4;
                for (int j0 = 0; j0 < n; j0 += 32 * UF) {
                    uint32_t pi[UF], src[UF];
#pragma unroll
                    for (int u = 0; u < UF; u++) {
                        const int j = j0 + 32 * u + lane;
                        pi[u] = j < n ? a.part[o + pofs + j] : 0u;
                    }
#pragma unroll
                    for (int u = 0; u < UF; u++) src[u] = perm_lookup(a, o, ng, b, pi[u]);
#pragma unroll
                    for (int u = 0; u < UF; u++) {
                        const int j = j0 + 32 * u + lane;
                        SCDD_CHECK(a.stats, j >= n || (pi[u] < (uint32_t)ng && src[u] < (uint32_t)ng && np <= a.np_max));
                        if (j < n) atomicOr(&flag[src[u] >> 5], 1u << (src[u] & 31));
                    }
                }
                __syncwarp();
                int cnt = 0;
                for (int k0 = 0; k0 < ng; k0 += 32 * UF) {
                    uint32_t src[UF];
                    double v[UF];
#pragma unroll
                    for (int u = 0; u < UF; u++) {
                        const int k = k0 + 32 * u + lane;
                        src[u] = k < ng ? a.srt_idx[o + k] : 0xffffffffu;
                        v[u] = k < ng ? a.srt_vals[o + k] : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < UF; u++) {
                        const bool f = src[u] != 0xffffffffu && ((flag[src[u] >> 5] >> (src[u] & 31)) & 1u);
                        const unsigned m = __ballot_sync(FULLMASK, f);
                        if (f) xs[cnt + __popc(m & ((1u << lane) - 1u))] = v[u];
                        cnt += __popc(m);
                    }
                }
                SCDD_CHECK(a.stats, cnt == n);
                for (int j = n + lane; j < np; j += 32) xs[j] = INFINITY;
            }
            __syncwarp();
        } else {
            for (int j = tid; j < np; j += T::NT) {
                double v = INFINITY;
                if (j < n) {
                    uint32_t i = a.part[o + pofs + j];
                    uint32_t src = perm_lookup(a, o, ng, b, i);
                    SCDD_CHECK(a.stats, i < (uint32_t)ng && src < (uint32_t)ng && np <= a.np_max);
                    v = permuted_value(a.vals, a.fitted, o, i, src);
                }
                xs[j] = v;
            }
            T::sync();
            if (TW > 1 && a.xs_global && a.sort_tile > 0 && np > a.sort_tile)
                team_sort_tiled<TW>(xs, np, lane, reinterpret_cast<double*>(region + fit_region_head(TW)), a.sort_tile);
            else
                team_sort<TW>(xs, np, lane);
        }

        double jp = NaN;
        int comps = 0;
        // all values identical: the reference jitters with runif(n, -0.1, 0.1) drawn mid-stream
        // (R/mclust.restricted.R:48-50).  With device-drawn permutations the stream state at that point was
        // recorded by permgen_group_kernel and the n uniforms are regenerated here, in element order; the jittered
        // copy is local to the fit (jointPosterior gets the identical values, R/permutations.R:278-282).
        bool jittered = false;
        double vconst = 0.0;
        if (n >= 1 && xs[0] == xs[n - 1] && a.jit_state) {
            const uint32_t* js = a.jit_state + (size_t)(((long long)g * a.nb + b) * a.sets + w) * 6;
            if (js[0] != JIT_NONE) {       // team-uniform (global read of the same word)
                jittered = true;
                vconst = xs[0];
                T::sync();
                if (lead0) {
                    Lecuyer rs;
#pragma unroll
                    for (int j = 0; j < 6; j++) rs.s[j] = js[j];
                    for (int j = 0; j < n; j++) {
                        const double u = lecuyer_unif(rs);
                        xs[j] = __dadd_rn(vconst, __dadd_rn(-0.1, __dmul_rn(0.2, u)));   // y + (a + (b - a) * unif_rand())
                    }
                }
                T::sync();
                team_sort<TW>(xs, np, lane);
            }
        }
        if (n < 1 || xs[0] == xs[n - 1]) {
            // no stream to draw the jitter from (observed data, host-drawn permutations): reported, never silently fitted
            if (lead0) { st->const_sets += 1.0; st->failed_sets += 1.0; st->fit_sets += 1.0; }
        } else {
            if (jittered && lead0) st->const_sets += 1.0;
            if (BUDGET) {
                WarpStats snap;
                if (lead0) snap = *st;
                comps = fit_set<TW, true>(xs, n, lane, rows, a.cfg, st, pm, red, a.iter_budget);
                if (comps == FIT_DEFERRED) {   // team-uniform
                    if (lead0) { *st = snap; a.deferred[item] = 1; }
                    T::sync();
                    continue;
                }
            } else {
                comps = fit_set<TW>(xs, n, lane, rows, a.cfg, st, pm, red);
            }
            if (comps > 0) jp = joint_posterior(rows[comps - 1], comps, a.cfg, jittered, vconst);
        }
        if (lead0) {
            a.jp[((long long)g * a.nb + b) * a.sets + w] = jp;
            if (a.gene_cost) atomicAdd(&a.gene_cost[g], st->triples - work0);
        }

        // ---- observed-data detail: the list mclustRestricted returns ----
        if (a.fits) {
            scdd_fit* f = &a.fits[3 * (size_t)g + (w == 2 ? 0 : w + 1)];
            int32_t* cls = (w == 2) ? a.cls_oa : a.cls_c;
            int ord[GMAX] = {0, 1, 2, 3, 4};
            if (comps > 0) {
                const Row& mc = rows[comps - 1];
                if (comps > 1) mean_order(mc, comps, ord);
                if (lead0) {
                    f->G = comps; f->model_v = comps > 1; f->n = n; f->df = comps > 1 ? 3 * comps - 1 : 2;
                    f->bic = mc.bic; f->loglik = mc.loglik;
                    for (int k = 0; k < GMAX; k++) {
                        f->mean[k] = k < comps ? mc.mean[ord[k]] : NaN;
                        f->var[k] = k < comps ? mc.sigsq[ord[k]] : NaN;
                    }
                }
                for (int j = tid; j < n; j += T::NT) {
                    uint32_t i = a.part[o + pofs + j];
                    double x = a.vals[o + perm_lookup(a, o, ng, b, i)];
                    int l = 0;
                    switch (comps) {
                        case 2: l = label_of<2>(mc, x); break;
                        case 3: l = label_of<3>(mc, x); break;
                        case 4: l = label_of<4>(mc, x); break;
                        case 5: l = label_of<5>(mc, x); break;
                        default: l = 0;
                    }
                    cls[o + i] = (comps > 1 ? ord[l] : l) + 1;   // R/mclust.restricted.R:243-247
                }
            } else {
                if (lead0) {
                    f->G = 0; f->model_v = 0; f->n = n; f->df = 0; f->bic = NaN; f->loglik = NaN;
                    for (int k = 0; k < GMAX; k++) { f->mean[k] = NaN; f->var[k] = NaN; }
                }
                for (int j = tid; j < n; j += T::NT) cls[o + a.part[o + pofs + j]] = 0;
            }
        }
        T::sync();
    }

    // ---- flush the work counters ----
    T::sync();
    if (lead0) {
        double v[9] = {st->triples, st->point_iters, st->em_iters, st->cap_hits, st->tol_near,
                       st->bic_near, st->const_sets, st->failed_sets, st->fit_sets};
#pragma unroll
        for (int k = 0; k < 9; k++) if (v[k] != 0.0) atomicAdd(&a.stats[k], v[k]);
    }
}

// bf.perm[g, b] = JP(c1) + JP(c2) [- JP(oa)]   (R/permutations.R:265-267, :281-282); jp_out (nullable) keeps the terms
__global__ void combine_kernel(const double* __restrict__ jp, int ngenes, int nb, int sets,
                               double* __restrict__ bf, long long ld, int col0, double* __restrict__ jp_out) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)ngenes * nb) return;
    int b = (int)(t % nb);
    long long g = t / nb;
    const double* p = jp + t * sets;
    double v = __dadd_rn(p[0], p[1]);
    if (sets == 3) v = __dsub_rn(v, p[2]);
    bf[g * ld + col0 + b] = v;
    if (jp_out) {
        double* q = jp_out + (g * ld + col0 + b) * sets;
        for (int w = 0; w < sets; w++) q[w] = p[w];
    }
}

// ------------------------------------------------------------------------------------------------
// KS: D of stats::ks.test(x, y, exact = FALSE) -- z <- cumsum(ifelse(order(c(x, y)) <= n.x, 1/n.x, -1/n.y)),
// ties keep the last element of each run, D = max |z|.  order() is stable, cumsum() accumulates in
// x87 extended precision (LDOUBLE) and stores doubles; both are reproduced exactly.
// One CTA per gene: bitonic sort of (value, position-in-c(x,y)) pairs in shared memory, then a serial
// extended-precision scan by thread 0 (the rounding sequence is inherently serial).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t orderable(double v) {
    if (v == 0.0) v = 0.0;   // -0 -> +0
    uint64_t b = (uint64_t)__double_as_longlong(v);
    return (b & 0x8000000000000000ULL) ? ~b : (b | 0x8000000000000000ULL);
}

__global__ void ks_kernel(const double* __restrict__ vals, const int64_t* __restrict__ off,
                          const uint32_t* __restrict__ part, const int32_t* __restrict__ n1, int ngenes,
                          int np_max, double* __restrict__ D, uint64_t* key_global, uint32_t* idx_global) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // genes too long for shared memory are sorted in a per-CTA global (L2-resident) slab
    uint64_t* key = key_global ? key_global + (size_t)blockIdx.x * np_max : reinterpret_cast<uint64_t*>(smem_raw);
    uint32_t* idx = idx_global ? idx_global + (size_t)blockIdx.x * np_max : reinterpret_cast<uint32_t*>(key + np_max);
    for (int g = blockIdx.x; g < ngenes; g += gridDim.x) {
        int64_t o = off[g];
        int n = (int)(off[g + 1] - o);
        int nx = n1[g], ny = n - nx;
        if (nx < 1 || ny < 1) {   // ks.test: "not enough 'x' data" / "not enough 'y' data"
            if (threadIdx.x == 0) D[g] = nan("");
            continue;
        }
        int np = 2;
        while (np < n) np <<= 1;
        for (int j = threadIdx.x; j < np; j += blockDim.x) {
            key[j] = (j < n) ? orderable(vals[o + part[o + j]]) : ~0ULL;
            idx[j] = (uint32_t)j;
        }
        __syncthreads();
        for (int k = 2; k <= np; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = threadIdx.x; t < (np >> 1); t += blockDim.x) {
                    int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    int l = i | j;
                    bool up = ((i & k) == 0);
                    uint64_t ka = key[i], kb = key[l];
                    uint32_t ia = idx[i], ib = idx[l];
                    bool gt = (ka > kb) || (ka == kb && ia > ib);
                    if (gt == up) { key[i] = kb; key[l] = ka; idx[i] = ib; idx[l] = ia; }
                }
                __syncthreads();
            }
        }
        if (threadIdx.x == 0) {
            const double cx = 1.0 / (double)nx, cy = -1.0 / (double)ny;
            const Ext80 ex = ext_from_double(cx), ey = ext_from_double(cy);
            Ext80 run = ext_zero();
            double dmax = 0.0;
            for (int i = 0; i < n; i++) {
                run = ext_add(run, idx[i] < (uint32_t)nx ? ex : ey);
                if (i == n - 1 || key[i + 1] != key[i]) {
                    double z = fabs(ext_to_double(run));
                    if (z > dmax) dmax = z;
                }
            }
            D[g] = dmax;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// DFMA throughput: 8 independent chains per thread
// ------------------------------------------------------------------------------------------------
__global__ void fp64_peak_kernel(double* out, int iters, double seed) {
    double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
    const double m = 0.999999, c = 1e-9 * (threadIdx.x + 1);
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace scdd
