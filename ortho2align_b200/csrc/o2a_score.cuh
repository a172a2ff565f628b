// o2a_score.cuh — K3+K4+K5: the streaming kernel of the path (the only stage that touches every HSP).
//
// One CTA per alignment (= lncRNA x syntenic region, the scope of the reference's multiple-testing
// step, pipeline.py:763-776):
//   filter_by_score(isf(alpha))  strict '>'                       alignment.py:618-646, pipeline.py:765
//   pvals = background.sf(scores of survivors)                     pipeline.py:766
//   fdr:  q = (m * p) / rank(p),  keep q <= alpha  (no cummin)      fitting.py:359-364, pipeline.py:769-776
// Scores are read once with coalesced 128-bit loads; survivors are compacted IN ORDER (ballot +
// CTA scan) into the alignment's slot [hsp_off[a], hsp_off[a]+n_surv) of the survivor arrays, so no
// global scan is needed and only survivors cost write traffic. Ranks use the build's tie policy:
// stable (ties keep HSP order) — SURVEY.md §8a row 7.
#pragma once
#include "o2a_device.cuh"

namespace o2a {

constexpr int SCORE_THREADS = 256;
constexpr int SCORE_ROWS = 4;                        // double2 per thread per row
constexpr int SCORE_TILE = SCORE_THREADS * 2 * SCORE_ROWS;
constexpr int RANK_SMALL_MAX = 1024;                 // survivors ranked inside the streaming kernel
constexpr int TBL_SMEM_MAX = 256;

struct ScoreArgs {
    long long n_aln;
    const long long* hsp_off;
    const int* aln_lnc;
    const double* score;
    const double* thr;
    const int* lnc_status;
    // hist
    const LncFit* fits;
    const int* tbl_k;
    const double* tbl_cum;
    const long long* bg_off;
    // kde / empirical (compressed background: distinct values + counts, see o2a_kde.cuh)
    const int* uv_val;
    const int* uv_cnt;          // kde: multiplicities; empirical: #{bg > value} suffix counts
    const int* uv_n;
    const double* kde_factor;
    int fitter;
    int fdr;
    double alpha;
    // outputs (slot layout)
    int* n_surv;
    int* n_keep;
    int* surv_idx;
    double* surv_p;
    double* surv_pq;
    unsigned char* surv_keep;
    int* big_list;              // alignments whose survivors exceed RANK_SMALL_MAX (fdr only)
    int* big_count;
};

__device__ __forceinline__ double2 ld_stream_f64x2(const double* p)
{
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}

// ndtr(z) = 0.5*erfc(-z/sqrt2): scipy.special.ndtr is cephes-derived and not bit-reproducible with
// CUDA's erfc (<= 2.3e-16 absolute apart) — KDE parity is to tolerance.
__device__ __forceinline__ double ndtr_dev(double z) { return 0.5 * erfc(-z * 0.70710678118654752440); }

// KernelFitter.sf (fitting.py:282-310) on the value-compressed background:
// cdf = sum_v cnt_v * ndtr((x - v)/factor) / n
__device__ double kde_sf_compressed(double x, const int* __restrict__ val, const int* __restrict__ cnt, int nv,
                                    double factor, double n)
{
    double acc = 0.0;
    for (int i = 0; i < nv; ++i) acc += (double)cnt[i] * ndtr_dev((x - (double)val[i]) / factor);
    return 1.0 - acc / n;
}

// empirical sf (extension): #{bg > x}/n from the sorted distinct values + suffix counts
__device__ double emp_sf_compressed(double x, const int* __restrict__ val, const int* __restrict__ gt, int nv, double n)
{
    int lo = 0, hi = nv;                              // first distinct value > x
    while (lo < hi) { int mid = (lo + hi) >> 1; if ((double)val[mid] > x) hi = mid; else lo = mid + 1; }
    double c = lo < nv ? (double)(gt[lo]) : 0.0;      // gt[i] = #{bg >= val[i]}
    return c / n;
}

__global__ void __launch_bounds__(SCORE_THREADS)
k_score(ScoreArgs A)
{
    __shared__ double s_p[RANK_SMALL_MAX];
    __shared__ int s_tk[TBL_SMEM_MAX];
    __shared__ double s_tc[TBL_SMEM_MAX];
    __shared__ int s_cnt[SCORE_ROWS][SCORE_THREADS / 32];
    __shared__ int s_base;
    __shared__ int s_keepcnt;
    const int tid = threadIdx.x, lane = lane_id(), w = warp_id();
    constexpr int NW = SCORE_THREADS / 32;

    for (long long a = blockIdx.x; a < A.n_aln; a += gridDim.x) {
        const int l = A.aln_lnc[a];
        const long long h0 = A.hsp_off[a];
        const int m = (int)(A.hsp_off[a + 1] - h0);
        if (A.lnc_status[l] != 0) {                   // fit failed: the reference raised before the loop
            if (tid == 0) { A.n_surv[a] = 0; A.n_keep[a] = 0; }
            continue;
        }
        const double thr = A.thr[l];
        const double* sc = A.score + h0;
        // fitter state
        LncFit f; const int* tk = nullptr; const double* tc = nullptr;
        const int* uval = nullptr; const int* ucnt = nullptr; int unv = 0; double kfac = 1.0, nbg = 1.0;
        if (A.fitter == 0) {
            f = A.fits[l];
            const long long tb = A.bg_off[l] >> 1;
            tk = A.tbl_k + tb; tc = A.tbl_cum + tb;
            if (f.tbl_n <= TBL_SMEM_MAX) {
                for (int t = tid; t < f.tbl_n; t += SCORE_THREADS) { s_tk[t] = tk[t]; s_tc[t] = tc[t]; }
                tk = s_tk; tc = s_tc;
            }
        } else {
            const long long b0 = A.bg_off[l];
            uval = A.uv_val + b0; ucnt = A.uv_cnt + b0; unv = A.uv_n[l];
            nbg = (double)(A.bg_off[l + 1] - b0);
            if (A.fitter == 1) kfac = A.kde_factor[l];
        }
        if (tid == 0) { s_base = 0; s_keepcnt = 0; }
        __syncthreads();
        const bool aligned16 = ((h0 & 1) == 0);       // score + h0 is 16-byte aligned
        for (int t0 = 0; t0 < m; t0 += SCORE_TILE) {
            double2 v[SCORE_ROWS];
            unsigned flag = 0;                        // 2 bits per row
#pragma unroll
            for (int r = 0; r < SCORE_ROWS; ++r) {
                int e = t0 + (r * SCORE_THREADS + tid) * 2;
                if (aligned16 && e + 1 < m) v[r] = ld_stream_f64x2(sc + e);
                else {
                    v[r].x = e < m ? sc[e] : 0.0;
                    v[r].y = e + 1 < m ? sc[e + 1] : 0.0;
                }
                bool fx = (e < m) && (v[r].x > thr);
                bool fy = (e + 1 < m) && (v[r].y > thr);
                unsigned bx = __ballot_sync(FULL_MASK, fx), by = __ballot_sync(FULL_MASK, fy);
                if (lane == 0) s_cnt[r][w] = __popc(bx) + __popc(by);
                // survivors of this warp-row before me, in element order (x0 y0 x1 y1 ...)
                int before = __popc(bx & lanemask_lt()) + __popc(by & lanemask_lt());
                flag |= (fx ? 1u : 0u) << (2 * r);
                flag |= (fy ? 1u : 0u) << (2 * r + 1);
                flag |= (unsigned)before << (8 + 6 * r);   // before <= 62 fits 6 bits
            }
            __syncthreads();
            const int base = s_base;
            int tile_total = 0;
#pragma unroll
            for (int r = 0; r < SCORE_ROWS; ++r) {
                int row_before = 0, row_total = 0;
#pragma unroll
                for (int ww = 0; ww < NW; ++ww) { int c = s_cnt[r][ww]; if (ww < w) row_before += c; row_total += c; }
                int pos = base + tile_total + row_before + (int)((flag >> (8 + 6 * r)) & 63u);
                int e = t0 + (r * SCORE_THREADS + tid) * 2;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    if ((flag >> (2 * r + c)) & 1u) {
                        double x = c ? v[r].y : v[r].x;
                        double p;
                        if (A.fitter == 0) p = hist_sf(x, f, tk, tc);
                        else if (A.fitter == 1) p = kde_sf_compressed(x, uval, ucnt, unv, kfac, nbg);
                        else p = emp_sf_compressed(x, uval, ucnt, unv, nbg);
                        A.surv_idx[h0 + pos] = e + c;
                        A.surv_p[h0 + pos] = p;
                        if (pos < RANK_SMALL_MAX) s_p[pos] = p;
                        ++pos;
                    }
                }
                tile_total += row_total;
            }
            __syncthreads();
            if (tid == 0) s_base = base + tile_total;
            __syncthreads();
        }
        const int ns = s_base;
        if (!A.fdr) {                                 // pipeline.py:778-780: hsp.pval = p, nothing dropped
            for (int k = tid; k < ns; k += SCORE_THREADS) {
                A.surv_pq[h0 + k] = (k < RANK_SMALL_MAX) ? s_p[k] : A.surv_p[h0 + k];
                A.surv_keep[h0 + k] = 1;
            }
            if (tid == 0) { A.n_surv[a] = ns; A.n_keep[a] = ns; }
        } else if (ns <= RANK_SMALL_MAX) {
            const double md = (double)m;              // total_hsps = len(alignment._all_HSPs)
            int kept = 0;
            for (int k = tid; k < ns; k += SCORE_THREADS) {
                const double p = s_p[k];
                int rank = 1;
                for (int j = 0; j < ns; ++j) { double pj = s_p[j]; rank += (pj < p) || (pj == p && j < k); }
                double q = __ddiv_rn(__dmul_rn(md, p), (double)rank);
                bool keep = q <= A.alpha;
                A.surv_pq[h0 + k] = q;
                A.surv_keep[h0 + k] = keep ? 1 : 0;
                kept += keep;
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) kept += __shfl_xor_sync(FULL_MASK, kept, d);
            if (lane == 0 && kept) atomicAdd(&s_keepcnt, kept);
            __syncthreads();
            if (tid == 0) { A.n_surv[a] = ns; A.n_keep[a] = s_keepcnt; }
        } else {
            if (tid == 0) {
                A.n_surv[a] = ns; A.n_keep[a] = 0;
                int slot = atomicAdd(A.big_count, 1);
                A.big_list[slot] = (int)a;
            }
        }
        __syncthreads();
    }
}

// Survivor sets larger than RANK_SMALL_MAX: stable CTA radix sort of the p-values in global scratch.
constexpr int RANKBIG_THREADS = 512;

__global__ void __launch_bounds__(RANKBIG_THREADS)
k_rank_big(const int* __restrict__ big_list, const int* __restrict__ big_count,
           const long long* __restrict__ hsp_off, double alpha,
           const int* __restrict__ n_surv, int* __restrict__ n_keep,
           const double* __restrict__ surv_p, double* __restrict__ surv_pq, unsigned char* __restrict__ surv_keep,
           unsigned long long* __restrict__ g_k0, unsigned long long* __restrict__ g_k1,
           unsigned int* __restrict__ g_v0, unsigned int* __restrict__ g_v1, long long cap)
{
    __shared__ int s_sort[32 + (RANKBIG_THREADS / 32) * 16 + 1];
    __shared__ int s_keepcnt;
    const int tid = threadIdx.x;
    unsigned long long* k0 = g_k0 + (long long)blockIdx.x * cap;
    unsigned long long* k1 = g_k1 + (long long)blockIdx.x * cap;
    unsigned int* v0 = g_v0 + (long long)blockIdx.x * cap;
    unsigned int* v1 = g_v1 + (long long)blockIdx.x * cap;
    const int nbig = *big_count;
    for (int b = blockIdx.x; b < nbig; b += gridDim.x) {
        const int a = big_list[b];
        const long long h0 = hsp_off[a];
        const int m = (int)(hsp_off[a + 1] - h0);
        const int ns = n_surv[a];
        for (int k = tid; k < ns; k += RANKBIG_THREADS) { k0[k] = sortable_f64(surv_p[h0 + k]); v0[k] = (unsigned)k; }
        if (tid == 0) s_keepcnt = 0;
        __syncthreads();
        int where = cta_radix_sort(k0, v0, k1, v1, ns, s_sort);
        __syncthreads();
        const unsigned int* vs = where ? v1 : v0;
        const double md = (double)m;
        int kept = 0;
        for (int r = tid; r < ns; r += RANKBIG_THREADS) {
            int k = (int)vs[r];
            double p = surv_p[h0 + k];
            double q = __ddiv_rn(__dmul_rn(md, p), (double)(r + 1));
            bool keep = q <= alpha;
            surv_pq[h0 + k] = q;
            surv_keep[h0 + k] = keep ? 1 : 0;
            kept += keep;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) kept += __shfl_xor_sync(FULL_MASK, kept, d);
        if (lane_id() == 0 && kept) atomicAdd(&s_keepcnt, kept);
        __syncthreads();
        if (tid == 0) n_keep[a] = s_keepcnt;
        __syncthreads();
    }
}

}  // namespace o2a
