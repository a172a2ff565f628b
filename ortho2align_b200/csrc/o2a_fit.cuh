// o2a_fit.cuh — K1/K2: per-lncRNA background fit + isf(alpha) for all three fitters.
//
// hist  — replaces HistogramFitter.__init__ / .isf (fitting.py:142-155, 206-215), i.e. numpy 2.3.5
//         np.histogram(bins=len//2, density=True) + scipy 1.18.1 rv_histogram.__init__ +
//         rv_continuous.isf, with their exact operation order:
//           edges   = linspace(min, max, B+1)            (k*step + first, last edge = max)
//           index   = int(((x-first)/denom)*B) with numpy's three edge fix-ups
//           dens    = n/diff(edges)/n.sum();  hp0 = dens/diff(edges)  (rv_histogram's 2nd division)
//           hpdf    = hp0/np.sum(hp0*w)  — np.sum is numpy's PAIRWISE sum (128-blocks, 8 accumulators)
//           hcdf    = np.cumsum(hpdf*w)  — sequential, left to right
//           isf(a)  = np.interp(1-a, hcdf, hbins)
// kde   — KernelFitter (fitting.py:245-356) on the value-compressed background (see below)
// emp   — extension
//
// One CTA per lncRNA. The background is read ONCE (the HBM-bound part) and reduced to its distinct
// values with multiplicities (integer BLAST scores: ~50 distinct values among 10^4..10^6 scores).
// Everything after that is exact arithmetic on the few non-empty bins: empty bins contribute
// exactly +0.0 to numpy's pairwise sum and cumsum, so the B-long hbins/hcdf arrays never exist —
// the result per lncRNA is a COMPACT table (bin index, cumulative value) of the non-empty bins.
#pragma once
#include "o2a_device.cuh"

namespace o2a {

constexpr int FIT_THREADS = 64;
constexpr int FIT_SMEM_RANGE = 1024;    // distinct-value counters kept in shared memory up to this range
constexpr int FIT_SMEM_NNZ = 512;       // distinct values / non-empty bins staged in shared memory

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

// empirical sf (extension): #{bg > x}/n from the sorted distinct values + suffix counts #{bg >= v}
__device__ double emp_sf_compressed(double x, const int* __restrict__ val, const int* __restrict__ ge, int nv, double n)
{
    int lo = 0, hi = nv;                              // first distinct value > x
    while (lo < hi) { int mid = (lo + hi) >> 1; if ((double)val[mid] > x) hi = mid; else lo = mid + 1; }
    double c = lo < nv ? (double)(ge[lo]) : 0.0;
    return c / n;
}

// CTA-wide deterministic sum: per-thread partials, then a fixed shuffle/shared tree.
__device__ double cta_sum(double v, double* s_red)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL_MASK, v, d);
    __syncthreads();
    if (lane_id() == 0) s_red[warp_id()] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_red[w];
    return t;
}

__device__ double kde_sf_cta(double x, const int* val, const int* cnt, int nv, double factor, double n, double* s_red)
{
    double acc = 0.0;
    for (int i = threadIdx.x; i < nv; i += blockDim.x) acc += (double)cnt[i] * ndtr_dev((x - (double)val[i]) / factor);
    return 1.0 - cta_sum(acc, s_red) / n;
}

// scipy/optimize/Zeros/brentq.c with scipy.optimize.brentq's defaults (xtol=2e-12, rtol=4*eps,
// maxiter=100); every thread of the CTA executes the same control flow. err: 0 ok, -1 sign, -2 conv.
__device__ double brentq_cta(double xa, double xb, double q, const int* val, const int* cnt, int nv,
                             double factor, double n, double* s_red, int* err)
{
    const double xtol = 2e-12, rtol = 8.881784197001252e-16;
    double xpre = xa, xcur = xb, xblk = 0., fpre, fcur, fblk = 0., spre = 0., scur = 0., sbis, delta, stry, dpre, dblk;
    *err = 0;
    fpre = kde_sf_cta(xpre, val, cnt, nv, factor, n, s_red) - q;
    fcur = kde_sf_cta(xcur, val, cnt, nv, factor, n, s_red) - q;
    if (fpre == 0) return xpre;
    if (fcur == 0) return xcur;
    if (signbit(fpre) == signbit(fcur)) { *err = -1; return 0.; }
    for (int i = 0; i < 100; ++i) {
        if (fpre != 0 && fcur != 0 && (signbit(fpre) != signbit(fcur))) { xblk = xpre; fblk = fpre; spre = scur = xcur - xpre; }
        if (fabs(fblk) < fabs(fcur)) { xpre = xcur; xcur = xblk; xblk = xpre; fpre = fcur; fcur = fblk; fblk = fpre; }
        delta = (xtol + rtol * fabs(xcur)) / 2;
        sbis = (xblk - xcur) / 2;
        if (fcur == 0 || fabs(sbis) < delta) return xcur;
        if (fabs(spre) > delta && fabs(fcur) < fabs(fpre)) {
            if (xpre == xblk) stry = -fcur * (xcur - xpre) / (fcur - fpre);
            else {
                dpre = (fpre - fcur) / (xpre - xcur);
                dblk = (fblk - fcur) / (xblk - xcur);
                stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre));
            }
            double m1 = fabs(spre), m2 = 3 * fabs(sbis) - delta;
            if (2 * fabs(stry) < (m1 < m2 ? m1 : m2)) { spre = scur; scur = stry; }
            else { spre = sbis; scur = sbis; }
        } else { spre = sbis; scur = sbis; }
        xpre = xcur; fpre = fcur;
        if (fabs(scur) > delta) xcur += scur; else xcur += (sbis > 0 ? delta : -delta);
        fcur = kde_sf_cta(xcur, val, cnt, nv, factor, n, s_red) - q;
    }
    *err = -2;
    return xcur;
}

// t_k = hp0_k * w_k of bin k (summand of rv_histogram's normaliser); *hp0w = (hp0, w) for reuse
__device__ __forceinline__ double fit_term(int cnt, long long k, const LncFit& f, double total, double* hp0_out, double* w_out)
{
    double w = __dsub_rn(hist_edge(k + 1, f), hist_edge(k, f));
    double dens = __ddiv_rn(__ddiv_rn((double)cnt, w), total);
    double hp0 = __ddiv_rn(dens, w);
    *hp0_out = hp0; *w_out = w;
    return __dmul_rn(hp0, w);
}

// numpy DOUBLE_pairwise_sum over a B-long array that is zero except at a few sorted positions.
// The recursion (n <= 128: 8-accumulator leaf; else split at n/2 rounded down to a multiple of 8)
// only matters through the non-empty entries: an all-zero subtree returns exactly 0.0 and
// x + 0.0 == x. So (1) every entry finds its leaf by descending the recursion, (2) each non-empty
// leaf is summed with numpy's accumulator order, (3) the leaves are combined in recursion order,
// which for leaves in position order is an operator-precedence evaluation where the operator
// between two adjacent leaves binds tighter the deeper their lowest common ancestor is.
struct PwLeaf { long long off; int n; int depth; unsigned long long path; };

__device__ __forceinline__ PwLeaf pw_find_leaf(long long pos, long long B)
{
    PwLeaf L; L.off = 0; L.depth = 0; L.path = 0ull;
    long long n = B;
    while (n > 128) {
        long long n2 = n / 2; n2 -= n2 % 8;
        if (pos - L.off < n2) { n = n2; L.path <<= 1; }
        else { L.off += n2; n -= n2; L.path = (L.path << 1) | 1ull; }
        ++L.depth;
    }
    L.n = (int)n;
    return L;
}

// depth of the lowest common ancestor of two different leaves
__device__ __forceinline__ int pw_lca_depth(const PwLeaf& a, const PwLeaf& b)
{
    int m = a.depth < b.depth ? a.depth : b.depth;
    unsigned long long x = (a.path >> (a.depth - m)) ^ (b.path >> (b.depth - m));
    return m - (64 - __clzll((long long)x));
}

// numpy leaf (n <= 128) over the entries e in [lo, hi) (positions pos[e] inside the leaf, values t[e])
__device__ double pw_leaf_sum(const int* pos, const double* t, int lo, int hi, long long off, int n)
{
    double res;
    if (n < 8) {
        res = 0.;
        for (int e = lo; e < hi; ++e) res = __dadd_rn(res, t[e]);
        return res;
    }
    double r0 = 0., r1 = 0., r2 = 0., r3 = 0., r4 = 0., r5 = 0., r6 = 0., r7 = 0.;
    const int lim = n - (n % 8);
    int e = lo;
    for (; e < hi && (int)((long long)pos[e] - off) < lim; ++e) {
        const int j = (int)((long long)pos[e] - off) & 7;
        const double v = t[e];
        switch (j) {
            case 0: r0 = __dadd_rn(r0, v); break; case 1: r1 = __dadd_rn(r1, v); break;
            case 2: r2 = __dadd_rn(r2, v); break; case 3: r3 = __dadd_rn(r3, v); break;
            case 4: r4 = __dadd_rn(r4, v); break; case 5: r5 = __dadd_rn(r5, v); break;
            case 6: r6 = __dadd_rn(r6, v); break; default: r7 = __dadd_rn(r7, v); break;
        }
    }
    res = __dadd_rn(__dadd_rn(__dadd_rn(r0, r1), __dadd_rn(r2, r3)), __dadd_rn(__dadd_rn(r4, r5), __dadd_rn(r6, r7)));
    for (; e < hi; ++e) res = __dadd_rn(res, t[e]);
    return res;
}

struct FitArgs {
    long long n_lnc;
    const long long* bg_off;
    const void* bg;                      // int32 (BG = int) or int16 (BG = short)
    const double* kde_factor;
    int fitter;
    double alpha;
    double* thr;
    LncFit* fits;
    int* tbl_k; double* tbl_cum;        // hist: compact table, slot bg_off[l] >> 1
    int* uv_val; int* uv_cnt; int* uv_n; // kde / empirical: distinct values, slot bg_off[l]
    int* status;
    int* g_cnt; long long max_range;     // per-CTA global counters for value ranges > FIT_SMEM_RANGE
    // Backgrounds beyond the counting envelope (value range > max_range, or more distinct values than the shared
    // staging holds) need per-CTA global scratch: a launch without it (scratch == nullptr) lists them in list2 and a
    // second, small-grid launch with scratch takes that list. np.histogram / gaussian_kde have no such limits.
    unsigned char* scratch; long long scratch_stride; long long scratch_cap;   // bytes per CTA, entries per CTA
    int* list2; int* list2_n;
    // hist fast path (k_fit_count + k_fit_hist_tail): fixed-capacity (value, count) slots per lncRNA;
    // lncRNAs it cannot take (wide value range, > FITW_CAP distinct values) are listed for k_fit
    int* fc_val; int* fc_cnt; int* fc_n;
    int* list; int* list_n;              // when list != nullptr, k_fit processes list[0..*list_n) only
};

// per-CTA global scratch of the general fit kernel (entries = scratch_cap >= the largest background)
struct FitScratch {
    unsigned long long* k0; unsigned long long* k1; unsigned int* v0; unsigned int* v1;   // radix sort
    int* K; int* C; double* T; double* LS; short* LD;                                     // value/bin, count, term, leaf sums
};
__device__ __forceinline__ FitScratch fit_scratch(const FitArgs& A)
{
    FitScratch S{};
    if (!A.scratch) return S;
    unsigned char* p = A.scratch + (long long)blockIdx.x * A.scratch_stride;
    const long long n = A.scratch_cap;
    S.k0 = (unsigned long long*)p; p += n * 8; S.k1 = (unsigned long long*)p; p += n * 8;
    S.T = (double*)p; p += n * 8; S.LS = (double*)p; p += n * 8;
    S.v0 = (unsigned int*)p; p += n * 4; S.v1 = (unsigned int*)p; p += n * 4;
    S.K = (int*)p; p += n * 4; S.C = (int*)p; p += n * 4;
    S.LD = (short*)p;
    return S;
}
constexpr int FIT_SCRATCH_BYTES_PER_ENTRY = 8 * 4 + 4 * 4 + 2;

// every background value packed in one 32-bit word of a 16-byte load (BG = int, short or signed char)
#define O2A_BG_WORD(OP, w) do {                                                                                   \
        if (PER == 4) { OP(w); }                                                                                  \
        else if (PER == 8) { OP((int)(short)((w) & 0xffff)); OP((w) >> 16); }                                     \
        else { OP((int)(signed char)((w) & 0xff)); OP((int)(signed char)(((w) >> 8) & 0xff));                     \
               OP((int)(signed char)(((w) >> 16) & 0xff)); OP((w) >> 24); } } while (0)
#define O2A_BG_VEC(OP, v) do { O2A_BG_WORD(OP, (v).x); O2A_BG_WORD(OP, (v).y); O2A_BG_WORD(OP, (v).z); O2A_BG_WORD(OP, (v).w); } while (0)

// General fit kernel: one CTA per lncRNA, any integer background (any value range, any number of distinct values).
//   counting   values in [0, FIT_SMEM_RANGE): shared counters during the min/max pass; range <= FIT_SMEM_RANGE:
//              shared counters relative to min; range <= max_range: per-CTA global counters; beyond: the values are
//              radix-sorted in per-CTA global scratch and run-length encoded
//   tables     up to FIT_SMEM_NNZ distinct values / non-empty bins in shared memory, more in the global scratch
// A launch without scratch lists what needs it (list2) for a second, small-grid launch that has it.
template <typename BG>
__global__ void __launch_bounds__(FIT_THREADS)
k_fit(FitArgs A)
{
    constexpr int PER = 16 / (int)sizeof(BG);           // background scores per 16-byte load
    __shared__ int s_cnt[FIT_SMEM_RANGE];
    __shared__ int s_k[FIT_SMEM_NNZ];        // distinct value (then bin index)
    __shared__ int s_c[FIT_SMEM_NNZ];        // multiplicity (then bin count)
    __shared__ double s_t[FIT_SMEM_NNZ];     // term / cumulative
    __shared__ int s_red_i[2 * (FIT_THREADS / 32)];
    __shared__ double s_red[FIT_THREADS / 32];
    __shared__ int s_scan[FIT_THREADS / 32 + 2];
    __shared__ int s_mn, s_mx, s_bad, s_oor;
    __shared__ double s_S;
    __shared__ double s_lsum[FIT_SMEM_NNZ];            // leaf sums (at the leaf's first entry)
    __shared__ short s_ldep[FIT_SMEM_NNZ];             // LCA depth between this leaf and the previous one
    __shared__ double s_vstk[72];
    __shared__ short s_ostk[72];
    const int tid = threadIdx.x;
    const double NaN = __longlong_as_double(0x7ff8000000000000ll);
    const FitScratch G = fit_scratch(A);

    const long long n_items = A.list ? (long long)*A.list_n : A.n_lnc;
    for (long long it = blockIdx.x; it < n_items; it += gridDim.x) {
        const long long l = A.list ? (long long)A.list[it] : it;
        const long long b0 = A.bg_off[l];
        const long long n = A.bg_off[l + 1] - b0;
        const BG* x = reinterpret_cast<const BG*>(A.bg) + b0;
        if (A.scratch && n > A.scratch_cap) continue;    // understated max_lnc_bg: the call fails on the host side
        // ---- pass over the background: min / max, and (optimistically) the multiplicities ----------
        // BLAST scores are small non-negative integers: values in [0, FIT_SMEM_RANGE) are counted
        // directly in shared memory during the same pass; anything else falls back to the general
        // two-pass path (min first, then counts relative to it).
        __syncthreads();                                 // the previous lncRNA is done with the shared tables
        {
            int4* z = reinterpret_cast<int4*>(s_cnt);
            for (int k = tid; k < FIT_SMEM_RANGE / 4; k += FIT_THREADS) z[k] = make_int4(0, 0, 0, 0);
            if (tid == 0) { s_bad = 0; s_oor = 0; }
        }
        __syncthreads();
        int mn = INT32_MAX, mx = INT32_MIN;
        {
            bool oor = false;
#define O2A_FIT_ADD(v) do { int v__ = (v); mn = min(mn, v__); mx = max(mx, v__);                              \
                if ((unsigned)v__ < (unsigned)FIT_SMEM_RANGE) atomicAdd(&s_cnt[v__], 1); else oor = true; } while (0)
            // 16-byte vector loads over the aligned middle part
            long long head = ((PER - (b0 & (PER - 1))) & (PER - 1)); if (head > n) head = n;
            for (long long i = tid; i < head; i += FIT_THREADS) O2A_FIT_ADD(x[i]);
            const int4* x4 = reinterpret_cast<const int4*>(x + head);
            const int n4 = (int)((n - head) / PER);
            for (int i = tid; i < n4; i += FIT_THREADS) {
                const int4 v = x4[i];
                O2A_BG_VEC(O2A_FIT_ADD, v);
            }
            for (long long i = head + (long long)n4 * PER + tid; i < n; i += FIT_THREADS) O2A_FIT_ADD(x[i]);
#undef O2A_FIT_ADD
            if (oor) s_oor = 1;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { mn = min(mn, __shfl_xor_sync(FULL_MASK, mn, d)); mx = max(mx, __shfl_xor_sync(FULL_MASK, mx, d)); }
        if (lane_id() == 0) { s_red_i[warp_id()] = mn; s_red_i[FIT_THREADS / 32 + warp_id()] = mx; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 0; w < FIT_THREADS / 32; ++w) { mn = min(mn, s_red_i[w]); mx = max(mx, s_red_i[FIT_THREADS / 32 + w]); }
            s_mn = mn; s_mx = mx;
        }
        __syncthreads();
        mn = s_mn; mx = s_mx;
        const bool fast = (s_oor == 0);
        const long long R = (long long)mx - (long long)mn + 1;
        const long long B = n / 2;
        if (A.fitter == 0 && B < 1) {                   // np.histogram(bins=0): ValueError
            if (tid == 0) { A.thr[l] = NaN; A.status[l] = 3; LncFit f{}; A.fits[l] = f; }
            continue;
        }
        const bool by_sort = !fast && R > A.max_range && R > FIT_SMEM_RANGE;
        if (by_sort && !A.scratch) {                    // needs the sort scratch: second launch
            if (tid == 0) { const int slot = atomicAdd(A.list2_n, 1); A.list2[slot] = (int)l; }
            continue;
        }
        // ---- distinct values with multiplicities, ascending: (K, Cn), in shared memory when they fit -----------
        int* uval = A.fitter == 0 ? nullptr : A.uv_val + b0;
        int* ucnt = A.fitter == 0 ? nullptr : A.uv_cnt + b0;
        int nv = 0;
        if (by_sort) {
            for (long long i = tid; i < n; i += FIT_THREADS) { G.k0[i] = (unsigned long long)((unsigned)(int)x[i] ^ 0x80000000u); G.v0[i] = 0u; }
            __syncthreads();
            const int where = cta_radix_sort(G.k0, G.v0, G.k1, G.v1, (int)n, s_cnt);     // s_cnt: free in this path
            __syncthreads();
            const unsigned long long* ks = where ? G.k1 : G.k0;
            // run heads -> (value, first index); counts from the distance to the next head
            for (long long i0 = 0; i0 < n; i0 += FIT_THREADS) {
                const long long i = i0 + tid;
                const bool headf = i < n && (i == 0 || ks[i] != ks[i - 1]);
                int tot;
                const int pos = cta_exclusive_scan(headf ? 1 : 0, s_scan, &tot);
                if (headf) { G.K[nv + pos] = (int)((unsigned)ks[i] ^ 0x80000000u); G.C[nv + pos] = (int)i; }
                nv += tot;
                __syncthreads();
            }
            for (int t = tid; t < nv; t += FIT_THREADS) {
                const int nxt = t + 1 < nv ? G.C[t + 1] : (int)n;
                G.T[t] = (double)(nxt - G.C[t]);          // parked: C[t + 1] is still needed by the neighbour
            }
            __syncthreads();
            for (int t = tid; t < nv; t += FIT_THREADS) {
                G.C[t] = (int)G.T[t];
                if (t < FIT_SMEM_NNZ) { s_k[t] = G.K[t]; s_c[t] = G.C[t]; }
                if (uval) { uval[t] = G.K[t]; ucnt[t] = G.C[t]; }
            }
            __syncthreads();
        } else {
            // cnt[k] = multiplicity of value mn + k
            int* cnt = s_cnt + mn;
            if (!fast) {
                cnt = (R <= FIT_SMEM_RANGE) ? s_cnt : (A.g_cnt + (long long)blockIdx.x * A.max_range);
                for (long long k = tid; k < R; k += FIT_THREADS) cnt[k] = 0;
                __syncthreads();
                for (long long i0 = 0; i0 < n; i0 += FIT_THREADS) {
                    long long i = i0 + tid;
                    bool valid = i < n;
                    unsigned m = __ballot_sync(FULL_MASK, valid);
                    if (valid) {
                        int c = (int)x[i] - mn;
                        unsigned peers = __match_any_sync(m, c);
                        if ((peers & lanemask_lt()) == 0) atomicAdd(&cnt[c], __popc(peers));
                    }
                }
                __syncthreads();
            }
            for (long long k0 = 0; k0 < R; k0 += FIT_THREADS) {
                long long k = k0 + tid;
                int c = k < R ? cnt[k] : 0;
                int tot;
                int pos = cta_exclusive_scan(c != 0, s_scan, &tot);
                if (c != 0) {
                    int p = nv + pos;
                    if (p < FIT_SMEM_NNZ) { s_k[p] = (int)(mn + k); s_c[p] = c; }
                    if (G.K && p < A.scratch_cap) { G.K[p] = (int)(mn + k); G.C[p] = c; }
                    if (uval) { uval[p] = (int)(mn + k); ucnt[p] = c; }
                }
                nv += tot;
                __syncthreads();
            }
        }
        const bool overflow = nv > FIT_SMEM_NNZ;
        if (A.fitter == 0 && overflow && !A.scratch) {  // more distinct values than the shared staging: second launch
            if (tid == 0) { const int slot = atomicAdd(A.list2_n, 1); A.list2[slot] = (int)l; }
            continue;
        }
        if (A.fitter == 1) {
            int err = 0;
            const double fac = A.kde_factor[l];
            const int* v = overflow ? uval : s_k; const int* c = overflow ? ucnt : s_c;
            double r = brentq_cta((double)mn, (double)mx, A.alpha, v, c, nv, fac, (double)n, s_red, &err);
            int st = 0;
            if (err == -1) {        // ValueError branch of inverse_approximate_collection (fitting.py:240-241)
                double fb = fabs(kde_sf_cta((double)mx, v, c, nv, fac, (double)n, s_red) - A.alpha);
                double fa = fabs(kde_sf_cta((double)mn, v, c, nv, fac, (double)n, s_red) - A.alpha);
                r = fb < fa ? (double)mx : (double)mn;
            } else if (err == -2) st = 2;
            if (tid == 0) { A.thr[l] = r; A.uv_n[l] = nv; A.status[l] = st; }
            continue;
        }
        if (A.fitter == 2) {
            if (tid == 0) {
                // isf = sorted[ceil((1-alpha)*n) - 1]; suffix counts #{bg >= val[i]} replace the counts
                long long k = (long long)ceil((1.0 - A.alpha) * (double)n) - 1;
                if (k < 0) k = 0;
                if (k > n - 1) k = n - 1;
                long long acc = 0; int t = 0;
                for (int i = 0; i < nv; ++i) { t = i; if (acc + ucnt[i] > k) break; acc += ucnt[i]; }
                A.thr[l] = (double)uval[t];
                long long suf = 0;
                for (int i = nv - 1; i >= 0; --i) { suf += ucnt[i]; ucnt[i] = (int)suf; }
                A.uv_n[l] = nv; A.status[l] = 0;
            }
            continue;
        }
        // ---- histogram fitter ------------------------------------------------------------------------
        int* K = overflow ? G.K : s_k; int* Cn = overflow ? G.C : s_c;
        double* T = overflow ? G.T : s_t; double* LS = overflow ? G.LS : s_lsum; short* LD = overflow ? G.LD : s_ldep;
        LncFit f;
        f.first = (double)mn; f.last = (double)mx;
        if (mn == mx) { f.first = __dsub_rn(f.first, 0.5); f.last = __dadd_rn(f.last, 0.5); }
        f.denom = __dsub_rn(f.last, f.first);
        f.B = B;
        f.step = __ddiv_rn(f.denom, (double)B);
        f.tbl_n = 0; f.pad = 0;
        // np.histogram raises "Too many bins for data range" unless the edges strictly increase.
        // step > 8 ulp(max |edge|) guarantees it (each edge carries <= 1.5 ulp of rounding); only
        // otherwise is the O(B) check needed.
        {
            double amax = fmax(fabs(f.first), fabs(f.last));
            if (!(f.step > amax * 1.7763568394002505e-15)) {
                for (long long k = tid; k < B; k += FIT_THREADS)
                    if (!(hist_edge(k, f) < hist_edge(k + 1, f))) s_bad = 1;
            }
        }
        // bin index of every distinct value (np.histogram uniform-bin path incl. its fix-ups)
        __syncthreads();
        if (s_bad) {
            if (tid == 0) { A.thr[l] = NaN; A.fits[l] = f; A.status[l] = 3; }
            continue;
        }
        for (int i = tid; i < nv; i += FIT_THREADS) {
            double v = (double)K[i];
            double g = __dmul_rn(__ddiv_rn(__dsub_rn(v, f.first), f.denom), (double)B);
            long long idx = (long long)g;
            if (idx == B) idx -= 1;
            if (v < hist_edge(idx, f)) idx -= 1;
            if (v >= hist_edge(idx + 1, f) && idx != B - 1) idx += 1;
            K[i] = (int)idx;              // value -> bin index (non-decreasing in i)
        }
        __syncthreads();
        // merge distinct values that share a bin: head flags -> bin slots
        int nnz = 0;
        for (int i0 = 0; i0 < nv; i0 += FIT_THREADS) {
            int i = i0 + tid;
            bool head = i < nv && (i == 0 || K[i] != K[i - 1]);
            int tot;
            int pos = cta_exclusive_scan(head ? 1 : 0, s_scan, &tot);
            // slot of element i = (number of heads up to and including i) - 1
            int slot = nnz + pos + (head ? 1 : 0) - 1;
            int kk = i < nv ? K[i] : 0, cc = i < nv ? Cn[i] : 0;
            __syncthreads();                       // everyone has read K/Cn of this chunk
            if (i < nv) {
                if (head) { K[slot] = kk; Cn[slot] = 0; }
            }
            __syncthreads();
            if (i < nv) atomicAdd(&Cn[slot], cc);
            nnz += tot;
            __syncthreads();
        }
        // terms of the normaliser, then S = np.sum(hp0*w) (pairwise), then c_k = (hp0/S)*w
        const double total = (double)n;
        for (int t = tid; t < nnz; t += FIT_THREADS) {
            double hp0, w;
            T[t] = fit_term(Cn[t], K[t], f, total, &hp0, &w);
        }
        __syncthreads();
        // (1)+(2): leaf of every entry; the first entry of each leaf sums its leaf
        for (int t = tid; t < nnz; t += FIT_THREADS) {
            const PwLeaf L = pw_find_leaf(K[t], B);
            bool head = true; int dep = -1;
            if (t > 0) {
                const PwLeaf Pl = pw_find_leaf(K[t - 1], B);
                head = (Pl.off != L.off);
                if (head) dep = pw_lca_depth(Pl, L);
            }
            LD[t] = (short)(head ? dep : -2);          // -2: not a leaf head
            if (head) {
                int hi = t + 1;
                while (hi < nnz && (long long)K[hi] < L.off + L.n) ++hi;
                LS[t] = pw_leaf_sum(K, T, t, hi, L.off, L.n);
            }
        }
        __syncthreads();
        // (3): precedence evaluation over the leaf heads, left to right
        if (tid == 0) {
            int vp = 0, op = 0;
            for (int t = 0; t < nnz; ++t) {
                const int d = LD[t];
                if (d == -2) continue;
                if (t > 0) {                                // operator between the previous leaf and this one
                    while (op > 0 && s_ostk[op - 1] > d) {
                        double b = s_vstk[--vp], a = s_vstk[--vp];
                        s_vstk[vp++] = __dadd_rn(a, b); --op;
                    }
                    s_ostk[op++] = (short)d;
                }
                s_vstk[vp++] = LS[t];
            }
            while (op > 0) { double b = s_vstk[--vp], a = s_vstk[--vp]; s_vstk[vp++] = __dadd_rn(a, b); --op; }
            s_S = nnz > 0 ? s_vstk[0] : 0.0;
        }
        __syncthreads();
        const double S = s_S;
        for (int t = tid; t < nnz; t += FIT_THREADS) {
            double hp0, w;
            fit_term(Cn[t], K[t], f, total, &hp0, &w);
            T[t] = __dmul_rn(__ddiv_rn(hp0, S), w);
        }
        __syncthreads();
        if (tid == 0) {
            // np.cumsum: sequential left to right
            double acc = 0.0;
            for (int t = 0; t < nnz; ++t) { acc = __dadd_rn(acc, T[t]); T[t] = acc; }
            // isf(alpha) = np.interp(1 - alpha, hcdf, hbins)
            double r;
            const double alpha = A.alpha;
            if (alpha == 1.0) r = f.first;
            else if (alpha == 0.0) r = f.last;
            else if (!(alpha > 0.0 && alpha < 1.0)) r = NaN;
            else {
                double y = __dsub_rn(1.0, alpha);
                int lo = 0, hi = nnz;                       // first entry with cum > y
                while (lo < hi) { int mid = (lo + hi) >> 1; if (T[mid] > y) hi = mid; else lo = mid + 1; }
                if (lo == nnz) r = f.last;                 // y >= hcdf[B]
                else {
                    long long j = K[lo];
                    double c0 = lo > 0 ? T[lo - 1] : 0.0, c1 = T[lo];
                    double xj = hist_edge(j, f);
                    if (c0 == y) r = xj;
                    else {
                        double slope = __ddiv_rn(__dsub_rn(hist_edge(j + 1, f), xj), __dsub_rn(c1, c0));
                        r = __dadd_rn(__dmul_rn(slope, __dsub_rn(y, c0)), xj);
                    }
                }
            }
            A.thr[l] = r;
            LncFit fo = f; fo.tbl_n = nnz;
            A.fits[l] = fo;
            A.status[l] = 0;
        }
        __syncthreads();
        const long long tbase = b0 >> 1;                   // table slot (capacity >= B >= nnz)
        for (int t = tid; t < nnz; t += FIT_THREADS) { A.tbl_k[tbase + t] = K[t]; A.tbl_cum[tbase + t] = T[t]; }
    }
}

// ---------------------------------------------------------------------------------------------
// Histogram fast path, stage 1: one CTA per lncRNA streams the background once and counts the
// distinct values (values in [0, FIT_SMEM_RANGE), the BLAST case) -> up to FITW_CAP (value, count)
// pairs per lncRNA. Anything else is flagged (fc_n = -1) for the general kernel k_fit.
// ---------------------------------------------------------------------------------------------
constexpr int FITC_THREADS = 128;
constexpr int FITW_CAP = 128;            // distinct background values handled by the warp-level tail

template <typename BG>
__global__ void __launch_bounds__(FITC_THREADS)
k_fit_count(FitArgs A)
{
    constexpr int PER = 16 / (int)sizeof(BG);
    __shared__ int s_cnt[FIT_SMEM_RANGE];
    __shared__ int s_scan[FITC_THREADS / 32 + 2];
    __shared__ int s_oor, s_lo, s_hi;
    const int tid = threadIdx.x;
    {   // the counters are zeroed once; every lncRNA re-zeroes only the range it touched
        int4* z = reinterpret_cast<int4*>(s_cnt);
        for (int k = tid; k < FIT_SMEM_RANGE / 4; k += FITC_THREADS) z[k] = make_int4(0, 0, 0, 0);
    }
    for (long long l = blockIdx.x; l < A.n_lnc; l += gridDim.x) {
        const long long b0 = A.bg_off[l];
        const long long n = A.bg_off[l + 1] - b0;
        const BG* x = reinterpret_cast<const BG*>(A.bg) + b0;
        if (tid == 0) { s_oor = 0; s_lo = FIT_SMEM_RANGE; s_hi = -1; }
        __syncthreads();
        {
            bool oor = false;
            int lo = FIT_SMEM_RANGE, hi = -1;           // range of the values this thread counted
#define O2A_CNT(v) do { int v__ = (v); if ((unsigned)v__ < (unsigned)FIT_SMEM_RANGE) { atomicAdd(&s_cnt[v__], 1); lo = min(lo, v__); hi = max(hi, v__); } else oor = true; } while (0)
            long long head = ((PER - (b0 & (PER - 1))) & (PER - 1)); if (head > n) head = n;
            for (long long i = tid; i < head; i += FITC_THREADS) O2A_CNT(x[i]);
            const int4* x4 = reinterpret_cast<const int4*>(x + head);
            const int n4 = (int)((n - head) / PER);
#define O2A_CNT16(v) O2A_BG_VEC(O2A_CNT, v)
            int i = tid;
            // eight 16-byte loads in flight per thread before the first shared-memory atomic
            for (; i + 7 * FITC_THREADS < n4; i += 8 * FITC_THREADS) {
                int4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = x4[i + u * FITC_THREADS];
#pragma unroll
                for (int u = 0; u < 8; ++u) O2A_CNT16(v[u]);
            }
            for (; i < n4; i += FITC_THREADS) { const int4 v = x4[i]; O2A_CNT16(v); }
#undef O2A_CNT16
            for (long long i = head + (long long)n4 * PER + tid; i < n; i += FITC_THREADS) O2A_CNT(x[i]);
#undef O2A_CNT
            if (oor) s_oor = 1;
            lo = __reduce_min_sync(FULL_MASK, lo); hi = __reduce_max_sync(FULL_MASK, hi);
            if (lane_id() == 0 && hi >= 0) { atomicMin(&s_lo, lo); atomicMax(&s_hi, hi); }
        }
        __syncthreads();
        int nv = 0;
        const int k_lo = s_hi >= 0 ? (s_lo / FITC_THREADS) * FITC_THREADS : 0, k_hi = s_hi;   // touched counters only
        const bool in_range = s_oor == 0;
        int* ov = A.fc_val + l * FITW_CAP; int* oc = A.fc_cnt + l * FITW_CAP;
        for (int k0 = k_lo; k0 <= k_hi; k0 += FITC_THREADS) {
            const int c = s_cnt[k0 + tid];
            s_cnt[k0 + tid] = 0;                                  // ready for the next lncRNA
            int tot;
            const int pos = cta_exclusive_scan(c != 0, s_scan, &tot);
            if (in_range && c != 0 && nv + pos < FITW_CAP) { ov[nv + pos] = k0 + tid; oc[nv + pos] = c; }
            nv += tot;
            __syncthreads();
        }
        if (tid == 0) A.fc_n[l] = (in_range && nv <= FITW_CAP) ? nv : -1;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Histogram fast path, stage 2: one WARP per lncRNA does the exact numpy/scipy arithmetic on the
// distinct values (same steps as k_fit's histogram branch, warp-synchronous): many lncRNAs in
// flight per SM hide the serial parts (pairwise-sum combine, cumsum, isf).
// ---------------------------------------------------------------------------------------------
constexpr int FITW_WARPS = 4;

__global__ void __launch_bounds__(FITW_WARPS * 32)
k_fit_hist_tail(FitArgs A)
{
    __shared__ int s_k[FITW_WARPS][FITW_CAP];
    __shared__ int s_c[FITW_WARPS][FITW_CAP];
    __shared__ double s_t[FITW_WARPS][FITW_CAP];
    __shared__ double s_lsum[FITW_WARPS][FITW_CAP];
    __shared__ short s_ldep[FITW_WARPS][FITW_CAP];
    __shared__ short s_rep[FITW_WARPS][FITW_CAP];
    __shared__ double s_hp0[FITW_WARPS][FITW_CAP];       // hp0 and bin width of every non-empty bin: three divisions and two
    __shared__ double s_w[FITW_WARPS][FITW_CAP];         // edges per bin that the normalised pass would otherwise repeat
    const int lane = lane_id(), w = warp_id();
    int* K = s_k[w]; int* Cn = s_c[w]; double* T = s_t[w]; double* LS = s_lsum[w]; short* LD = s_ldep[w];
    short* RP = s_rep[w]; double* HP = s_hp0[w]; double* WW = s_w[w];
    const double NaN = __longlong_as_double(0x7ff8000000000000ll);
    const long long nwarps = (long long)gridDim.x * FITW_WARPS;
    for (long long l = (long long)blockIdx.x * FITW_WARPS + w; l < A.n_lnc; l += nwarps) {
        const int nv = A.fc_n[l];
        if (nv < 0) {                                    // general kernel takes it
            if (lane == 0) { const int slot = atomicAdd(A.list_n, 1); A.list[slot] = (int)l; }
            continue;
        }
        const long long b0 = A.bg_off[l];
        const long long n = A.bg_off[l + 1] - b0;
        const long long B = n / 2;
        __syncwarp();
        for (int i = lane; i < nv; i += 32) { K[i] = A.fc_val[l * FITW_CAP + i]; Cn[i] = A.fc_cnt[l * FITW_CAP + i]; }
        __syncwarp();
        int st = 0; double thr_v = NaN;
        LncFit f{};
        if (B < 1 || nv < 1) st = 3;
        else {
            const int mn = K[0], mx = K[nv - 1];
            f.first = (double)mn; f.last = (double)mx;
            if (mn == mx) { f.first = __dsub_rn(f.first, 0.5); f.last = __dadd_rn(f.last, 0.5); }
            f.denom = __dsub_rn(f.last, f.first);
            f.B = B;
            f.step = __ddiv_rn(f.denom, (double)B);
            // "Too many bins for data range" unless the edges strictly increase (see k_fit)
            const double amax = fmax(fabs(f.first), fabs(f.last));
            if (!(f.step > amax * 1.7763568394002505e-15)) {
                bool bad = false;
                for (long long k = lane; k < B; k += 32) if (!(hist_edge(k, f) < hist_edge(k + 1, f))) bad = true;
                if (__any_sync(FULL_MASK, bad)) st = 3;
            }
        }
        int nnz = 0;
        if (st == 0) {
            // bin index of every distinct value (np.histogram uniform-bin path incl. its fix-ups)
            for (int i = lane; i < nv; i += 32) {
                const double v = (double)K[i];
                const double g = __dmul_rn(__ddiv_rn(__dsub_rn(v, f.first), f.denom), (double)B);
                long long idx = (long long)g;
                if (idx == B) idx -= 1;
                if (v < hist_edge(idx, f)) idx -= 1;
                if (v >= hist_edge(idx + 1, f) && idx != B - 1) idx += 1;
                K[i] = (int)idx;
            }
            __syncwarp();
            // merge distinct values that share a bin (bins are non-decreasing): head flags -> slots
            for (int i0 = 0; i0 < nv; i0 += 32) {
                const int i = i0 + lane;
                const bool head = i < nv && (i == 0 || K[i] != K[i - 1]);
                const unsigned hb = __ballot_sync(FULL_MASK, head);
                const int slot = nnz + __popc(hb & (lanemask_lt() | (1u << lane))) - 1;
                const int kk = i < nv ? K[i] : 0, cc = i < nv ? Cn[i] : 0;
                __syncwarp();
                if (head) { K[slot] = kk; Cn[slot] = 0; }
                __syncwarp();
                if (i < nv) atomicAdd(&Cn[slot], cc);
                nnz += __popc(hb);
                __syncwarp();
            }
            const double total = (double)n;
            for (int t = lane; t < nnz; t += 32) { double hp0, ww; T[t] = fit_term(Cn[t], K[t], f, total, &hp0, &ww); HP[t] = hp0; WW[t] = ww; }
            __syncwarp();
            // numpy pairwise sum: leaves (heads sum their leaf), then precedence evaluation
            for (int t0 = 0; t0 < nnz; t0 += 32) {
                const int t = t0 + lane;
                // every lane takes part in the shuffles; lanes past the end compute a leaf nobody uses
                const PwLeaf L = pw_find_leaf(K[t < nnz ? t : nnz - 1], B);
                // the leaf of entry t - 1: the neighbouring lane has just found it (lane 0: an entry of the previous round)
                PwLeaf P;
                P.off = __shfl_up_sync(FULL_MASK, L.off, 1); P.n = __shfl_up_sync(FULL_MASK, L.n, 1);
                P.depth = __shfl_up_sync(FULL_MASK, L.depth, 1); P.path = __shfl_up_sync(FULL_MASK, L.path, 1);
                if (lane == 0 && t > 0) P = pw_find_leaf(K[t - 1], B);
                if (t >= nnz) continue;
                bool head = true; int dep = -1;
                if (t > 0) {
                    head = (P.off != L.off);
                    if (head) dep = pw_lca_depth(P, L);
                }
                LD[t] = (short)(head ? dep : -2);
                if (head) {
                    int hi = t + 1;
                    while (hi < nnz && (long long)K[hi] < L.off + L.n) ++hi;
                    LS[t] = pw_leaf_sum(K, T, t, hi, L.off, L.n);
                }
            }
            __syncwarp();
            // Combine the leaf sums exactly as numpy's recursion does: the "operator" between two consecutive non-empty
            // leaves is the tree node that is their lowest common ancestor, and a deeper node is summed first. Every
            // node joins the finished group on its left (which starts after the nearest shallower operator) with the
            // finished group on its right (which starts at its own leaf); nodes of equal depth belong to disjoint
            // subtrees, so one pass per depth, deepest first, evaluates the whole tree with all lanes busy.
            int dmax = -1;
            for (int t = lane; t < nnz; t += 32) {
                const int d = LD[t];
                int rep = 0;
                if (d >= 0) {
                    int u = t - 1;
                    for (; u > 0; --u) { const int du = LD[u]; if (du != -2 && du < d) break; }
                    rep = u;
                    dmax = max(dmax, d);
                }
                RP[t] = (short)rep;
            }
            dmax = __reduce_max_sync(FULL_MASK, dmax);
            __syncwarp();
            for (int D = dmax; D >= 0; --D) {
                for (int t = lane; t < nnz; t += 32)
                    if (LD[t] == D) { const int rp = RP[t]; LS[rp] = __dadd_rn(LS[rp], LS[t]); }
                __syncwarp();
            }
            double S = nnz > 0 ? LS[0] : 0.0;
            for (int t = lane; t < nnz; t += 32) T[t] = __dmul_rn(__ddiv_rn(HP[t], S), WW[t]);
            __syncwarp();
            if (lane == 0) {
                double acc = 0.0;                          // np.cumsum: sequential left to right
                for (int t = 0; t < nnz; ++t) { acc = __dadd_rn(acc, T[t]); T[t] = acc; }
                const double alpha = A.alpha;
                double r;
                if (alpha == 1.0) r = f.first;
                else if (alpha == 0.0) r = f.last;
                else if (!(alpha > 0.0 && alpha < 1.0)) r = NaN;
                else {
                    const double y = __dsub_rn(1.0, alpha);
                    int lo = 0, hi = nnz;                   // first entry with cum > y
                    while (lo < hi) { int mid = (lo + hi) >> 1; if (T[mid] > y) hi = mid; else lo = mid + 1; }
                    if (lo == nnz) r = f.last;
                    else {
                        const long long j = K[lo];
                        const double c0 = lo > 0 ? T[lo - 1] : 0.0, c1 = T[lo];
                        const double xj = hist_edge(j, f);
                        if (c0 == y) r = xj;
                        else {
                            const double slope = __ddiv_rn(__dsub_rn(hist_edge(j + 1, f), xj), __dsub_rn(c1, c0));
                            r = __dadd_rn(__dmul_rn(slope, __dsub_rn(y, c0)), xj);
                        }
                    }
                }
                thr_v = r;
            }
            thr_v = __shfl_sync(FULL_MASK, thr_v, 0);
            __syncwarp();
            const long long tbase = b0 >> 1;
            for (int t = lane; t < nnz; t += 32) { A.tbl_k[tbase + t] = K[t]; A.tbl_cum[tbase + t] = T[t]; }
        }
        if (lane == 0) {
            LncFit fo = f; fo.tbl_n = st == 0 ? nnz : 0;
            A.thr[l] = st == 0 ? thr_v : NaN; A.fits[l] = fo; A.status[l] = st;
        }
        __syncwarp();
    }
}

}  // namespace o2a
