// o2a_score.cuh — K3 (streaming score filter) and K4+K5 (p-values, rank, q, keep).
//
// k_filter — the only stage that touches every HSP; HBM-bound. One CTA per alignment
//   (= lncRNA x syntenic region): filter_by_score(isf(alpha)), strict '>'
//   (alignment.py:618-646, pipeline.py:765). Scores are read once with coalesced 128-bit loads and
//   the survivors (position + score) are compacted IN ORDER (ballot + CTA scan) into the
//   alignment's slot [hsp_off[a], hsp_off[a]+n_surv) of the survivor arrays: no global scan, and
//   only survivors cost write traffic.
// k_pq — per alignment, on the ~7 % that survived: pvals = background.sf(scores) (pipeline.py:766);
//   fdr: q = (m*p)/rank(p), keep q <= alpha, no cummin (fitting.py:359-364, pipeline.py:769-776).
//   Ranks use the build's tie policy: stable (ties keep HSP order) — SURVEY.md §8a row 7.
#pragma once
#include "o2a_device.cuh"
#include "o2a_fit.cuh"

namespace o2a {

constexpr int FILTER_THREADS = 256;
constexpr int FILTER_ROWS = 4;                       // double2 per thread per row
constexpr int FILTER_TILE = FILTER_THREADS * 2 * FILTER_ROWS;
constexpr int PQ_THREADS = 128;
constexpr int RANK_SMALL_MAX = 1024;                 // survivors ranked in shared memory
constexpr int TBL_SMEM_MAX = 256;

__device__ __forceinline__ double2 ld_stream_f64x2(const double* p)
{
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}

// Layout inside a tile of FILTER_TILE elements: warp w owns the contiguous chunk
// [w*FILTER_WCHUNK, (w+1)*FILTER_WCHUNK), as FILTER_ROWS rows of 64 elements (32 lanes x double2,
// 512 contiguous bytes per warp load). Order-preserving positions then need only: the lane prefix
// (ballots), the running row prefix (registers) and ONE cross-warp prefix per tile (8 counters).
__global__ void __launch_bounds__(FILTER_THREADS, 5)
k_filter(long long n_aln, const long long* __restrict__ hsp_off, const int* __restrict__ aln_lnc,
         const double* __restrict__ score, const double* __restrict__ thr, const int* __restrict__ lnc_status,
         int* __restrict__ n_surv, int* __restrict__ surv_idx, double* __restrict__ surv_x)
{
    constexpr int NW = FILTER_THREADS / 32;
    constexpr int WCHUNK = FILTER_ROWS * 64;
    __shared__ int s_wtot[2][NW];
    const int lane = lane_id(), w = warp_id();
    const unsigned lt = lanemask_lt();
    int buf = 0;

    for (long long a = blockIdx.x; a < n_aln; a += gridDim.x) {
        const int l = aln_lnc[a];
        const long long h0 = hsp_off[a];
        const int m = (int)(hsp_off[a + 1] - h0);
        if (lnc_status[l] != 0) {                     // fit failed: the reference raised before the loop
            if (threadIdx.x == 0) n_surv[a] = 0;
            continue;
        }
        const double t = thr[l];
        // the stream is read from the 16-byte aligned address at or just below score + h0: stream
        // element i is HSP i - sh of the alignment (sh = 0 or 1); element -1 belongs to the neighbour
        const int sh = (int)(h0 & 1);
        const double* sc = score + h0 - sh;
        const int ms = m + sh;                        // stream length
        int* out_idx = surv_idx + h0;
        double* out_x = surv_x + h0;
        const double NEG_INF = __longlong_as_double(0xfff0000000000000ll);   // never passes the strict '>'
        int base = 0;
        for (int t0 = 0; t0 < ms; t0 += FILTER_TILE) {
            const int e0 = t0 + w * WCHUNK + lane * 2;            // my first stream element of row 0
            double2 v[FILTER_ROWS];
#pragma unroll
            for (int r = 0; r < FILTER_ROWS; ++r) {
                const int rs = t0 + w * WCHUNK + r * 64;          // row start (warp-uniform)
                if (rs + 64 <= ms) v[r] = ld_stream_f64x2(sc + e0 + r * 64);
                else if (rs >= ms) { v[r].x = NEG_INF; v[r].y = NEG_INF; }
                else {
                    const int e = e0 + r * 64;
                    v[r].x = e < ms ? sc[e] : NEG_INF;
                    v[r].y = e + 1 < ms ? sc[e + 1] : NEG_INF;
                }
            }
            if (sh && t0 == 0 && w == 0 && lane == 0) v[0].x = NEG_INF;   // the neighbour's last HSP
            unsigned bx[FILTER_ROWS], by[FILTER_ROWS];
            int wtot = 0;
#pragma unroll
            for (int r = 0; r < FILTER_ROWS; ++r) {
                bx[r] = __ballot_sync(FULL_MASK, v[r].x > t);
                by[r] = __ballot_sync(FULL_MASK, v[r].y > t);
                wtot += __popc(bx[r]) + __popc(by[r]);
            }
            if (lane == 0) s_wtot[buf][w] = wtot;
            __syncthreads();
            int wbase = 0, tile_total = 0;
#pragma unroll
            for (int ww = 0; ww < NW; ++ww) { const int c = s_wtot[buf][ww]; wbase += (ww < w) ? c : 0; tile_total += c; }
            buf ^= 1;                                             // double-buffered: one barrier per tile
            if (wtot) {
                int pos = base + wbase;
#pragma unroll
                for (int r = 0; r < FILTER_ROWS; ++r) {
                    const unsigned mx = bx[r], my = by[r];
                    if (mx | my) {
                        int p = pos + __popc(mx & lt) + __popc(my & lt);
                        const int e = e0 + r * 64 - sh;          // HSP position inside the alignment
                        if ((mx >> lane) & 1u) { out_idx[p] = e; out_x[p] = v[r].x; ++p; }
                        if ((my >> lane) & 1u) { out_idx[p] = e + 1; out_x[p] = v[r].y; }
                        pos += __popc(mx) + __popc(my);
                    }
                }
            }
            base += tile_total;
        }
        if (threadIdx.x == 0) n_surv[a] = base;
    }
}

struct PqArgs {
    long long n_aln;
    const long long* hsp_off;
    const int* aln_lnc;
    const int* lnc_status;
    const int* n_surv;
    // hist
    const LncFit* fits; const int* tbl_k; const double* tbl_cum; const long long* bg_off;
    // kde / empirical
    const int* uv_val; const int* uv_cnt; const int* uv_n; const double* kde_factor;
    int fitter; int fdr; double alpha;
    double* surv_p;            // in: survivor scores (written by k_filter); out: p-values
    double* surv_pq; unsigned char* surv_keep; int* n_keep;
    int* big_list; int* big_count;
};

__global__ void __launch_bounds__(PQ_THREADS)
k_pq(PqArgs A)
{
    __shared__ double s_p[RANK_SMALL_MAX];
    __shared__ int s_tk[TBL_SMEM_MAX];
    __shared__ double s_tc[TBL_SMEM_MAX];
    __shared__ int s_keepcnt;
    const int tid = threadIdx.x, lane = lane_id();

    for (long long a = blockIdx.x; a < A.n_aln; a += gridDim.x) {
        const int ns = A.n_surv[a];
        if (ns == 0) { if (tid == 0) A.n_keep[a] = 0; continue; }
        const int l = A.aln_lnc[a];
        const long long h0 = A.hsp_off[a];
        const int m = (int)(A.hsp_off[a + 1] - h0);
        LncFit f; const int* tk = nullptr; const double* tc = nullptr;
        const int* uval = nullptr; const int* ucnt = nullptr; int unv = 0; double kfac = 1.0, nbg = 1.0;
        if (A.fitter == 0) {
            f = A.fits[l];
            const long long tb = A.bg_off[l] >> 1;
            tk = A.tbl_k + tb; tc = A.tbl_cum + tb;
            if (f.tbl_n <= TBL_SMEM_MAX) {
                for (int t = tid; t < f.tbl_n; t += PQ_THREADS) { s_tk[t] = tk[t]; s_tc[t] = tc[t]; }
                tk = s_tk; tc = s_tc;
            }
        } else {
            const long long b0 = A.bg_off[l];
            uval = A.uv_val + b0; ucnt = A.uv_cnt + b0; unv = A.uv_n[l];
            nbg = (double)(A.bg_off[l + 1] - b0);
            if (A.fitter == 1) kfac = A.kde_factor[l];
        }
        if (tid == 0) s_keepcnt = 0;
        __syncthreads();
        for (int k = tid; k < ns; k += PQ_THREADS) {
            const double x = A.surv_p[h0 + k];
            double p;
            if (A.fitter == 0) p = hist_sf(x, f, tk, tc);
            else if (A.fitter == 1) p = kde_sf_compressed(x, uval, ucnt, unv, kfac, nbg);
            else p = emp_sf_compressed(x, uval, ucnt, unv, nbg);
            A.surv_p[h0 + k] = p;
            if (k < RANK_SMALL_MAX) s_p[k] = p;
            if (!A.fdr) { A.surv_pq[h0 + k] = p; A.surv_keep[h0 + k] = 1; }   // pipeline.py:778-780
        }
        if (!A.fdr) {
            if (tid == 0) A.n_keep[a] = ns;
            __syncthreads();
            continue;
        }
        __syncthreads();
        if (ns <= RANK_SMALL_MAX) {
            const double md = (double)m;              // total_hsps = len(alignment._all_HSPs)
            int kept = 0;
            for (int k = tid; k < ns; k += PQ_THREADS) {
                const double p = s_p[k];
                // stable rank: 1 + #{j < k : p_j <= p} + #{j > k : p_j < p}
                int rank = 1;
#pragma unroll 4
                for (int j = 0; j < k; ++j) rank += (s_p[j] <= p);
#pragma unroll 4
                for (int j = k + 1; j < ns; ++j) rank += (s_p[j] < p);
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
            if (tid == 0) A.n_keep[a] = s_keepcnt;
        } else if (tid == 0) {
            A.n_keep[a] = 0;
            int slot = atomicAdd(A.big_count, 1);
            A.big_list[slot] = (int)a;
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
