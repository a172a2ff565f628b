// o2a_kde.cuh — background "fit" for the KDE fitter and the empirical extension.
//
// KernelFitter (fitting.py:245-356): cdf(x) = mean(ndtr((x - data)/factor)) (:295), sf = 1 - cdf
// (:310), isf = brentq(sf(x) - q, min(data), max(data)) with end-point fallback (:218-242, :350).
// Backgrounds are integer BLAST scores with few distinct values, so the background is compressed
// to (distinct value, multiplicity) pairs once per lncRNA: cdf(x) = sum_v cnt_v*ndtr((x-v)/factor)/n.
// That turns an FP64-pipe-bound O(survivors x n_bg) loop into O(survivors x distinct) and changes
// only the summation order; scipy's ndtr is not bit-reproducible on the GPU anyway (SURVEY.md
// §8a row 14), so KDE parity is to |dp| <= max(1e-12*p, 16*2^-53) by construction.
//
// Empirical fitter (extension): sf(x) = #{bg > x}/n,
// isf(q) = sorted_bg[ceil((1-q)*n) - 1]; stored as distinct values + suffix counts #{bg >= v}.
#pragma once
#include "o2a_device.cuh"
#include "o2a_score.cuh"

namespace o2a {

constexpr int UV_THREADS = 256;
constexpr int UV_SMEM_RANGE = 8192;

// CTA-wide deterministic sum: per-thread strided partials, then a fixed shuffle/shared tree.
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

// One CTA per lncRNA: distinct values + counts (dense counting over [min, max]), then isf(alpha).
// status 4 = value range larger than the counting scratch (max_range) — extension limit.
__global__ void __launch_bounds__(UV_THREADS)
k_fit_uv(long long n_lnc, const long long* __restrict__ bg_off, const int* __restrict__ bg,
         const double* __restrict__ kde_factor, int fitter, double alpha,
         double* __restrict__ thr, int* __restrict__ uv_val, int* __restrict__ uv_cnt, int* __restrict__ uv_n,
         int* __restrict__ status, int* __restrict__ g_cnt, long long max_range)
{
    __shared__ int s_cnt[UV_SMEM_RANGE];
    __shared__ int s_red_i[2 * (UV_THREADS / 32)];
    __shared__ double s_red[UV_THREADS / 32];
    __shared__ int s_scan[UV_THREADS / 32 + 2];
    __shared__ int s_mn, s_mx;
    const int tid = threadIdx.x;
    for (long long l = blockIdx.x; l < n_lnc; l += gridDim.x) {
        const long long b0 = bg_off[l];
        const long long n = bg_off[l + 1] - b0;
        const int* x = bg + b0;
        int mn = INT32_MAX, mx = INT32_MIN;
        for (long long i = tid; i < n; i += UV_THREADS) { int v = x[i]; mn = min(mn, v); mx = max(mx, v); }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { mn = min(mn, __shfl_xor_sync(FULL_MASK, mn, d)); mx = max(mx, __shfl_xor_sync(FULL_MASK, mx, d)); }
        if (lane_id() == 0) { s_red_i[warp_id()] = mn; s_red_i[UV_THREADS / 32 + warp_id()] = mx; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 0; w < UV_THREADS / 32; ++w) { mn = min(mn, s_red_i[w]); mx = max(mx, s_red_i[UV_THREADS / 32 + w]); }
            s_mn = mn; s_mx = mx;
        }
        __syncthreads();
        mn = s_mn; mx = s_mx;
        const long long R = (long long)mx - (long long)mn + 1;
        if (R > max_range && R > UV_SMEM_RANGE) {
            if (tid == 0) { thr[l] = __longlong_as_double(0x7ff8000000000000ll); uv_n[l] = 0; status[l] = 4; }
            __syncthreads();
            continue;
        }
        int* cnt = (R <= UV_SMEM_RANGE) ? s_cnt : (g_cnt + (long long)blockIdx.x * max_range);
        for (long long k = tid; k < R; k += UV_THREADS) cnt[k] = 0;
        __syncthreads();
        for (long long i = tid; i < n; i += UV_THREADS) atomicAdd(&cnt[x[i] - mn], 1);
        __syncthreads();
        int nv = 0;
        for (long long k0 = 0; k0 < R; k0 += UV_THREADS) {
            long long k = k0 + tid;
            int c = k < R ? cnt[k] : 0;
            int tot;
            int pos = cta_exclusive_scan(c != 0, s_scan, &tot);
            if (c != 0) { uv_val[b0 + nv + pos] = (int)(mn + k); uv_cnt[b0 + nv + pos] = c; }
            nv += tot;
            __syncthreads();
        }
        int* val = uv_val + b0; int* ucnt = uv_cnt + b0;
        if (fitter == 1) {
            int err = 0;
            double fac = kde_factor[l];
            double r = brentq_cta((double)mn, (double)mx, alpha, val, ucnt, nv, fac, (double)n, s_red, &err);
            int st = 0;
            if (err == -1) {        // ValueError branch of inverse_approximate_collection (fitting.py:240-241)
                double fb = fabs(kde_sf_cta((double)mx, val, ucnt, nv, fac, (double)n, s_red) - alpha);
                double fa = fabs(kde_sf_cta((double)mn, val, ucnt, nv, fac, (double)n, s_red) - alpha);
                r = fb < fa ? (double)mx : (double)mn;
            } else if (err == -2) st = 2;
            if (tid == 0) { thr[l] = r; uv_n[l] = nv; status[l] = st; }
        } else {
            // suffix counts #{bg >= val[i]} and isf = sorted[ceil((1-alpha)*n) - 1]
            if (tid == 0) {
                long long k = (long long)ceil((1.0 - alpha) * (double)n) - 1;
                if (k < 0) k = 0;
                if (k > n - 1) k = n - 1;
                long long acc = 0; int t = 0;
                for (int i = 0; i < nv; ++i) { if (acc + ucnt[i] > k) { t = i; break; } acc += ucnt[i]; t = i; }
                thr[l] = (double)val[t];
                long long suf = 0;
                for (int i = nv - 1; i >= 0; --i) { suf += ucnt[i]; ucnt[i] = (int)suf; }
                uv_n[l] = nv; status[l] = 0;
            }
        }
        __syncthreads();
    }
}

}  // namespace o2a
