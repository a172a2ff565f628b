// o2a_device.cuh — device-side building blocks shared by the kernels of libo2a_cuda.so.
//
// Everything whose bits are compared with the reference is fp64 with explicit round-to-nearest
// intrinsics (__dadd_rn/__dmul_rn/__ddiv_rn never contract into FMA), and the translation unit is
// compiled with -fmad=false on top of that.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define O2A_WARP 32
#define FULL_MASK 0xffffffffu

namespace o2a {

// ---------------------------------------------------------------------------------------------
// per-lncRNA result of the background fit, consumed by the scoring kernel
// ---------------------------------------------------------------------------------------------
struct LncFit {
    double first, last;   // outer histogram edges (numpy _get_outer_edges)
    double step;          // np.linspace step = (last-first)/B
    double denom;         // norm_denom of np.histogram's uniform-bin path (last-first)
    long long B;          // number of bins = len(bg)//2 (fitting.py:152-154)
    int tbl_n;            // number of non-empty bins (compact hcdf table length)
    int pad;
};

__device__ __forceinline__ double hist_edge(long long k, const LncFit& f)
{   // np.linspace: y = arange(num)*step + start; y[-1] = stop
    return k >= f.B ? f.last : __dadd_rn(__dmul_rn((double)k, f.step), f.first);
}

// hcdf[e] (e = edge index 0..B) from the compact table of non-empty bins:
// tbl_k ascending bin indices, tbl_cum[t] = np.cumsum value at the right edge of bin tbl_k[t].
// Returns the number of table entries with bin index < e (so hcdf[e] = cum[cnt-1] or 0).
__device__ __forceinline__ int tbl_count_lt(const int* __restrict__ tbl_k, int n, long long e)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if ((long long)tbl_k[mid] < e) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// rv_histogram.sf(x) (scipy rv_continuous.sf + _cdf = np.interp(x, hbins, hcdf)); fitting.py:184-193
__device__ __forceinline__ double hist_sf(double x, const LncFit& f, const int* __restrict__ tbl_k,
                                          const double* __restrict__ tbl_cum)
{
    if (x != x) return x;
    if (x <= f.first) return 1.0;
    if (!(x < f.last)) return 0.0;
    // locate j with hbins[j] <= x < hbins[j+1] (what np.interp's binary search returns); the
    // uniform-bin guess is corrected against the actual linspace edges
    double g = __dmul_rn(__ddiv_rn(__dsub_rn(x, f.first), f.denom), (double)f.B);
    long long j = (long long)g;
    if (j < 0) j = 0;
    if (j > f.B - 1) j = f.B - 1;
    while (j > 0 && x < hist_edge(j, f)) --j;
    while (j < f.B - 1 && x >= hist_edge(j + 1, f)) ++j;
    int c = tbl_count_lt(tbl_k, f.tbl_n, j);
    double c0 = c > 0 ? tbl_cum[c - 1] : 0.0;
    double c1 = (c < f.tbl_n && (long long)tbl_k[c] == j) ? tbl_cum[c] : c0;
    double xj = hist_edge(j, f);
    double r;
    if (xj == x) r = c0;
    else {
        double slope = __ddiv_rn(__dsub_rn(c1, c0), __dsub_rn(hist_edge(j + 1, f), xj));
        r = __dadd_rn(__dmul_rn(slope, __dsub_rn(x, xj)), c0);
    }
    return __dsub_rn(1.0, r);
}

// order-preserving map double -> uint64 (handles negatives; -0.0 < +0.0 is harmless here: p is
// produced by 1.0 - r and never -0.0)
__device__ __forceinline__ unsigned long long sortable_f64(double x)
{
    unsigned long long u = (unsigned long long)__double_as_longlong(x);
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}

// ---------------------------------------------------------------------------------------------
// CTA-wide helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }
__device__ __forceinline__ unsigned lanemask_lt() { return (1u << lane_id()) - 1u; }

// exclusive prefix of `v` over the CTA in thread order; *total gets the CTA sum.
// `s_warp` must hold blockDim.x/32 + 1 ints. Contains two __syncthreads().
__device__ __forceinline__ int cta_exclusive_scan(int v, int* s_warp, int* total)
{
    int lane = lane_id(), w = warp_id(), nw = blockDim.x >> 5;
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int y = __shfl_up_sync(FULL_MASK, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) s_warp[w] = x;
    __syncthreads();
    if (w == 0) {
        int t = lane < nw ? s_warp[lane] : 0;
        int s = t;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int y = __shfl_up_sync(FULL_MASK, s, d);
            if (lane >= d) s += y;
        }
        if (lane < nw) s_warp[lane] = s - t;
        if (lane == 31) s_warp[nw] = s;
    }
    __syncthreads();
    int base = s_warp[w];
    *total = s_warp[nw];
    return base + x - v;
}

// ---------------------------------------------------------------------------------------------
// Stable CTA-level LSD radix sort of (64-bit key, 32-bit payload), 4-bit digits, n elements in
// global (or shared) memory, ping-pong with (k1, v1). Passes in which every key has the same
// digit are skipped. Returns 0 if the result is in (k0, v0), 1 if it is in (k1, v1).
// Needs shared scratch: int s[16 + 16 + (blockDim/32)*16 + 1].
// ---------------------------------------------------------------------------------------------
__device__ int cta_radix_sort(unsigned long long* k0, unsigned int* v0, unsigned long long* k1,
                              unsigned int* v1, int n, int* s)
{
    int* hist = s;                 // 16
    int* base = s + 16;            // 16
    int* wcnt = s + 32;            // nw*16
    const int nw = blockDim.x >> 5;
    const int lane = lane_id(), w = warp_id();
    int flip = 0;
    for (int shift = 0; shift < 64; shift += 4) {
        unsigned long long* ki = flip ? k1 : k0;
        unsigned int* vi = flip ? v1 : v0;
        unsigned long long* ko = flip ? k0 : k1;
        unsigned int* vo = flip ? v0 : v1;
        if (threadIdx.x < 16) hist[threadIdx.x] = 0;
        __syncthreads();
        for (int i0 = 0; i0 < n; i0 += blockDim.x) {
            int i = i0 + threadIdx.x;
            bool valid = i < n;
            unsigned m = __ballot_sync(FULL_MASK, valid);
            if (valid) {
                int d = (int)((ki[i] >> shift) & 15ull);
                unsigned peers = __match_any_sync(m, d);
                if ((peers & lanemask_lt()) == 0) atomicAdd(&hist[d], __popc(peers));
            }
        }
        __syncthreads();
        bool skip = false;
        for (int d = 0; d < 16; ++d) if (hist[d] == n) skip = true;
        if (skip) { __syncthreads(); continue; }
        if (threadIdx.x == 0) {
            int acc = 0;
            for (int d = 0; d < 16; ++d) { base[d] = acc; acc += hist[d]; }
        }
        __syncthreads();
        for (int i0 = 0; i0 < n; i0 += blockDim.x) {
            for (int t = threadIdx.x; t < nw * 16; t += blockDim.x) wcnt[t] = 0;
            __syncthreads();
            int i = i0 + threadIdx.x;
            bool valid = i < n;
            unsigned m = __ballot_sync(FULL_MASK, valid);
            unsigned long long key = 0; unsigned int val = 0; int d = 0, r = 0;
            if (valid) {
                key = ki[i]; val = vi[i];
                d = (int)((key >> shift) & 15ull);
                unsigned peers = __match_any_sync(m, d);
                r = __popc(peers & lanemask_lt());
                if (r == 0) wcnt[w * 16 + d] = __popc(peers);
            }
            __syncthreads();
            if (valid) {
                int off = base[d] + r;
                for (int ww = 0; ww < w; ++ww) off += wcnt[ww * 16 + d];
                ko[off] = key; vo[off] = val;
            }
            __syncthreads();
            if (threadIdx.x < 16) {
                int add = 0;
                for (int ww = 0; ww < nw; ++ww) add += wcnt[ww * 16 + threadIdx.x];
                base[threadIdx.x] += add;
            }
            __syncthreads();
        }
        flip ^= 1;
    }
    return flip;
}

}  // namespace o2a
