// o2a_fit.cuh — K1/K2: per-lncRNA histogram background fit + isf(alpha).
//
// Replaces HistogramFitter.__init__ / .isf (fitting.py:142-155, 206-215), i.e. numpy 2.3.5
// np.histogram(bins=len//2, density=True) + scipy 1.18.1 rv_histogram.__init__ + rv_continuous.isf,
// with their exact operation order:
//   edges   = linspace(min, max, B+1)            (k*step + first, last edge = max)
//   index   = int(((x-first)/denom)*B) with numpy's three edge fix-ups
//   dens    = n/diff(edges)/n.sum();  hp0 = dens/diff(edges)   (rv_histogram's second division)
//   hpdf    = hp0/np.sum(hp0*w)  — np.sum is numpy's PAIRWISE sum (128-blocks, 8 accumulators)
//   hcdf    = np.cumsum(hpdf*w)  — sequential, left to right
//   isf(a)  = np.interp(1-a, hcdf, hbins)
// One CTA per lncRNA. The only per-lncRNA output besides thr is a COMPACT table of the non-empty
// bins (bin index, cumulative value): empty bins add exactly 0.0 everywhere, so hcdf is a step
// function of the non-empty ones — the 2x(B+1) doubles of hbins/hcdf never touch HBM.
#pragma once
#include "o2a_device.cuh"

namespace o2a {

constexpr int FIT_THREADS = 256;
constexpr int FIT_SMEM_BINS = 8192;     // bin counters kept in shared memory up to this B
constexpr int FIT_SMEM_LEAVES = 512;    // pairwise-sum leaves kept in shared memory

struct PwNode { long long off; long long n; };

// t_k = hp0_k * w_k of bin k (the summand of rv_histogram's normaliser)
__device__ __forceinline__ double fit_term(int cnt, long long k, const LncFit& f, double total)
{
    double w = __dsub_rn(hist_edge(k + 1, f), hist_edge(k, f));
    double dens = __ddiv_rn(__ddiv_rn((double)cnt, w), total);
    double hp0 = __ddiv_rn(dens, w);
    return __dmul_rn(hp0, w);
}

// numpy DOUBLE_pairwise_sum leaf (n <= 128) over bins [off, off+n), zero bins skipped (adding +0.0
// to a non-negative accumulator is exact)
__device__ double fit_leaf_sum(const int* cnt, long long off, long long n, const LncFit& f, double total)
{
    if (n < 8) {
        double res = 0.;
        for (long long i = 0; i < n; ++i) { int c = cnt[off + i]; if (c) res = __dadd_rn(res, fit_term(c, off + i, f, total)); }
        return res;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { int c = cnt[off + j]; r[j] = c ? fit_term(c, off + j, f, total) : 0.0; }
    long long i;
    long long lim = n - (n % 8);
    for (i = 8; i < lim; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { int c = cnt[off + i + j]; if (c) r[j] = __dadd_rn(r[j], fit_term(c, off + i + j, f, total)); }
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) { int c = cnt[off + i]; if (c) res = __dadd_rn(res, fit_term(c, off + i, f, total)); }
    return res;
}

// Kernel. Global scratch (per CTA): g_cnt [gridDim.x * max_B] ints (only used when B > FIT_SMEM_BINS),
// g_leaf_off/g_leaf_n/g_leaf_sum [gridDim.x * max_leaves].
__global__ void __launch_bounds__(FIT_THREADS)
k_fit_hist(long long n_lnc, const long long* __restrict__ bg_off, const int* __restrict__ bg, double alpha,
           double* __restrict__ thr, LncFit* __restrict__ fits, int* __restrict__ tbl_k,
           double* __restrict__ tbl_cum, int* __restrict__ status,
           int* __restrict__ g_cnt, long long max_B,
           PwNode* __restrict__ g_leaf, double* __restrict__ g_leaf_sum, long long max_leaves)
{
    __shared__ int s_cnt[FIT_SMEM_BINS];
    __shared__ PwNode s_leaf[FIT_SMEM_LEAVES];
    __shared__ double s_leaf_sum[FIT_SMEM_LEAVES];
    __shared__ int s_red[2 * (FIT_THREADS / 32)];
    __shared__ int s_scan[FIT_THREADS / 32 + 2];
    __shared__ LncFit s_fit;
    __shared__ long long s_nleaf;
    __shared__ double s_S;
    __shared__ int s_bad;
    const int tid = threadIdx.x;

    for (long long l = blockIdx.x; l < n_lnc; l += gridDim.x) {
        const long long b0 = bg_off[l];
        const long long n = bg_off[l + 1] - b0;
        const int* x = bg + b0;
        const long long B = n / 2;
        const long long tbase = b0 >> 1;           // table slot of this lncRNA (capacity >= B)
        // ---- min / max -------------------------------------------------------------------
        int mn = INT32_MAX, mx = INT32_MIN;
        for (long long i = tid; i < n; i += FIT_THREADS) { int v = x[i]; mn = min(mn, v); mx = max(mx, v); }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            mn = min(mn, __shfl_xor_sync(FULL_MASK, mn, d));
            mx = max(mx, __shfl_xor_sync(FULL_MASK, mx, d));
        }
        if (lane_id() == 0) { s_red[warp_id()] = mn; s_red[FIT_THREADS / 32 + warp_id()] = mx; }
        if (tid == 0) s_bad = 0;
        __syncthreads();
        if (tid == 0) {
            for (int w = 0; w < FIT_THREADS / 32; ++w) { mn = min(mn, s_red[w]); mx = max(mx, s_red[FIT_THREADS / 32 + w]); }
            LncFit f;
            f.first = (double)mn; f.last = (double)mx;
            if (mn == mx) { f.first = __dsub_rn(f.first, 0.5); f.last = __dadd_rn(f.last, 0.5); }
            f.denom = __dsub_rn(f.last, f.first);
            f.B = B;
            f.step = B > 0 ? __ddiv_rn(f.denom, (double)B) : 0.0;
            f.tbl_n = 0; f.pad = 0;
            s_fit = f;
        }
        __syncthreads();
        const LncFit f = s_fit;
        if (B < 1) {                                  // cannot happen after the reference's patches
            if (tid == 0) { thr[l] = __longlong_as_double(0x7ff8000000000000ll); fits[l] = f; status[l] = 3; }
            __syncthreads();
            continue;
        }
        // np.histogram raises "Too many bins for data range" unless edges strictly increase
        for (long long k = tid; k < B; k += FIT_THREADS)
            if (!(hist_edge(k, f) < hist_edge(k + 1, f))) s_bad = 1;
        // ---- bin counts ------------------------------------------------------------------
        int* cnt = (B <= FIT_SMEM_BINS) ? s_cnt : (g_cnt + (long long)blockIdx.x * max_B);
        for (long long k = tid; k < B; k += FIT_THREADS) cnt[k] = 0;
        __syncthreads();
        if (s_bad) {
            if (tid == 0) { thr[l] = __longlong_as_double(0x7ff8000000000000ll); fits[l] = f; status[l] = 3; }
            __syncthreads();
            continue;
        }
        for (long long i = tid; i < n; i += FIT_THREADS) {
            double v = (double)x[i];
            double g = __dmul_rn(__ddiv_rn(__dsub_rn(v, f.first), f.denom), (double)B);
            long long idx = (long long)g;
            if (idx == B) idx -= 1;
            if (v < hist_edge(idx, f)) idx -= 1;
            if (v >= hist_edge(idx + 1, f) && idx != B - 1) idx += 1;
            atomicAdd(&cnt[idx], 1);
        }
        __syncthreads();
        // ---- normaliser: numpy pairwise sum over the B terms ---------------------------------
        const double total = (double)n;
        PwNode* leaf = (B <= (long long)FIT_SMEM_LEAVES * 64) ? s_leaf : (g_leaf + (long long)blockIdx.x * max_leaves);
        double* leaf_sum = (B <= (long long)FIT_SMEM_LEAVES * 64) ? s_leaf_sum : (g_leaf_sum + (long long)blockIdx.x * max_leaves);
        if (tid == 0) {
            // enumerate the leaves of the recursion (n <= 128) left to right
            PwNode stack[64];
            int sp = 0; long long nl = 0;
            stack[sp++] = PwNode{0, B};
            while (sp > 0) {
                PwNode nd = stack[--sp];
                if (nd.n <= 128) { leaf[nl++] = nd; }
                else {
                    long long n2 = nd.n / 2; n2 -= n2 % 8;
                    stack[sp++] = PwNode{nd.off + n2, nd.n - n2};   // right first: left pops first
                    stack[sp++] = PwNode{nd.off, n2};
                }
            }
            s_nleaf = nl;
        }
        __syncthreads();
        const long long nleaf = s_nleaf;
        for (long long q = tid; q < nleaf; q += FIT_THREADS)
            leaf_sum[q] = fit_leaf_sum(cnt, leaf[q].off, leaf[q].n, f, total);
        __syncthreads();
        if (tid == 0) {
            // combine the leaf sums in the recursion's order: post-order walk with a value stack
            struct Fr { long long off, n; int state; };
            Fr st[64]; double vs[64];
            int sp = 0, vp = 0; long long next_leaf = 0;
            st[sp++] = Fr{0, B, 0};
            while (sp > 0) {
                Fr& t = st[sp - 1];
                if (t.n <= 128) { vs[vp++] = leaf_sum[next_leaf++]; --sp; continue; }
                long long n2 = t.n / 2; n2 -= n2 % 8;
                if (t.state == 0) { t.state = 1; st[sp++] = Fr{t.off, n2, 0}; }
                else if (t.state == 1) { t.state = 2; st[sp++] = Fr{t.off + n2, t.n - n2, 0}; }
                else { double b = vs[--vp]; double a = vs[--vp]; vs[vp++] = __dadd_rn(a, b); --sp; }
            }
            s_S = vs[0];
        }
        __syncthreads();
        const double S = s_S;
        // ---- compact the non-empty bins, in bin order ------------------------------------------
        int nnz = 0;
        for (long long k0 = 0; k0 < B; k0 += FIT_THREADS) {
            long long k = k0 + tid;
            int c = (k < B) ? cnt[k] : 0;
            int tot;
            int pos = cta_exclusive_scan(c != 0, s_scan, &tot);
            if (c != 0) {
                // c_k = hpdf_k * w_k with hpdf_k = hp0_k / S
                double w = __dsub_rn(hist_edge(k + 1, f), hist_edge(k, f));
                double dens = __ddiv_rn(__ddiv_rn((double)c, w), total);
                double hp0 = __ddiv_rn(dens, w);
                double hpdf = __ddiv_rn(hp0, S);
                tbl_k[tbase + nnz + pos] = (int)k;
                tbl_cum[tbase + nnz + pos] = __dmul_rn(hpdf, w);
            }
            nnz += tot;
            __syncthreads();
        }
        // ---- np.cumsum (sequential) + isf ---------------------------------------------------------
        if (tid == 0) {
            double acc = 0.0;
            for (int t = 0; t < nnz; ++t) { acc = __dadd_rn(acc, tbl_cum[tbase + t]); tbl_cum[tbase + t] = acc; }
            double r;
            if (alpha == 1.0) r = f.first;
            else if (alpha == 0.0) r = f.last;
            else if (!(alpha > 0.0 && alpha < 1.0)) r = __longlong_as_double(0x7ff8000000000000ll);
            else {
                double y = __dsub_rn(1.0, alpha);
                // first table entry with cum > y
                int lo = 0, hi = nnz;
                while (lo < hi) { int mid = (lo + hi) >> 1; if (tbl_cum[tbase + mid] > y) hi = mid; else lo = mid + 1; }
                if (lo == nnz) r = f.last;             // y >= hcdf[B]
                else {
                    long long j = tbl_k[tbase + lo];
                    double c0 = lo > 0 ? tbl_cum[tbase + lo - 1] : 0.0, c1 = tbl_cum[tbase + lo];
                    double xj = hist_edge(j, f);
                    if (c0 == y) r = xj;
                    else {
                        double slope = __ddiv_rn(__dsub_rn(hist_edge(j + 1, f), xj), __dsub_rn(c1, c0));
                        r = __dadd_rn(__dmul_rn(slope, __dsub_rn(y, c0)), xj);
                    }
                }
            }
            thr[l] = r;
            LncFit fo = f; fo.tbl_n = nnz;
            fits[l] = fo;
            status[l] = 0;
        }
        __syncthreads();
    }
}

}  // namespace o2a
