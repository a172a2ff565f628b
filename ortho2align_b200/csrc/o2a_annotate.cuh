// o2a_annotate.cuh — interval join of annotate_orthologs:
//   GenomicRangesList.get_neighbours -> BaseGenomicRange.find_neighbours   genomicranges.py:2040-2067, 635-685
//   BaseGenomicRange.distance / intersect / calc_JI / calc_OC              genomicranges.py:439-458, 460-495, 497-547
// for every ortholog ("self") against the annotation, which is a list sorted by (chrom, start, end, name)
// (_genomic_key_func, genomicranges.py:1566-1572).
//
// find_neighbours, restated: idx = bisect_left(annotation, self). The LEFT loop walks idx-1, idx-2, ... down to
// the first range on another chromosome (InconsistentChromosomesError ends it) and keeps every range with
// |distance| <= d; the RIGHT loop walks idx, idx+1, ... and stops at the first range that is not a neighbour.
// The reference's left loop is O(chromosome); here the candidates are bounded with a running maximum of `end`
// over the sorted list: a left range can only be a neighbour when end + d >= self.start (for inverted ranges,
// end < start, which the reference's loops accept: when max(start, end) >= min(self.start - d, self.end + 1)),
// and the running maximum is monotone, so the first candidate is a binary search. The chromosome rank is packed above `end`,
// which restarts the maximum at every chromosome without a segmented scan.
// Count pass -> exclusive scan -> write pass (order = the sorted list's) -> JI / OC per pair.
#pragma once
#include "o2a_device.cuh"

namespace o2a {

struct AnnArgs {
    long long n_q, n_a;
    const int* q_chrom; const int* q_start; const int* q_end; const int* q_name; const signed char* q_strand;
    const int* a_chrom; const int* a_start; const int* a_end; const int* a_name; const signed char* a_strand;
    long long distance;
    int strandness;
    unsigned long long* prefmax;   // [n_a] running max of (chrom, max(start, end))
    unsigned long long* chunk_max; // [chunks + 1]
    int* blkmax;                   // [ceil(n_a / ANN_BLK)] largest max(start, end) of each block of ANN_BLK ranges
    int* cnt;                      // [n_q] neighbours per ortholog
    int4* span;                    // [n_q] (lo, idx, -, -) found by the count pass: first left candidate, bisect_left position
    const long long* off;          // [n_q + 1]
    int* nb_idx; double* ji; double* oc;
    int* pair_q;                   // [pairs] owner of each pair, for k_ann_ji_oc
    long long n_pairs;
};

constexpr int ANN_THREADS = 256;
constexpr int ANN_CHUNK = 1024;
constexpr int ANN_BLK = 32;          // annotation ranges per skip block (one warp of k_ann_prefmax)

__device__ __forceinline__ unsigned long long ann_pack(int chrom, int end)
{
    return ((unsigned long long)(unsigned)chrom << 32) | (unsigned long long)((unsigned)end ^ 0x80000000u);
}

// ---- running maximum of (chrom, end) over the sorted annotation: chunk maxima, scan of the maxima, per-chunk rescan ----
__global__ void __launch_bounds__(256) k_ann_chunk_max(AnnArgs A)
{
    __shared__ unsigned long long s_w[8];
    for (long long c = blockIdx.x; c * ANN_CHUNK < A.n_a; c += gridDim.x) {
        const long long i0 = c * ANN_CHUNK, i1 = min(A.n_a, i0 + ANN_CHUNK);
        unsigned long long m = 0;
        for (long long i = i0 + threadIdx.x; i < i1; i += 256) m = max(m, ann_pack(A.a_chrom[i], max(A.a_start[i], A.a_end[i])));
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_xor_sync(FULL_MASK, m, d));
        if (lane_id() == 0) s_w[warp_id()] = m;
        __syncthreads();
        if (threadIdx.x == 0) { unsigned long long t = 0; for (int w = 0; w < 8; ++w) t = max(t, s_w[w]); A.chunk_max[c] = t; }
        __syncthreads();
    }
}

// single CTA: exclusive running maximum of the chunk maxima, in place
__global__ void __launch_bounds__(1024) k_ann_scan_max(AnnArgs A)
{
    __shared__ unsigned long long s_w[32];
    __shared__ unsigned long long s_base;
    const long long n_chunks = (A.n_a + ANN_CHUNK - 1) / ANN_CHUNK;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (long long i0 = 0; i0 < n_chunks; i0 += 1024) {
        const long long i = i0 + threadIdx.x;
        const unsigned long long v = i < n_chunks ? A.chunk_max[i] : 0;
        unsigned long long x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { unsigned long long y = __shfl_up_sync(FULL_MASK, x, d); if (lane_id() >= d) x = max(x, y); }
        if (lane_id() == 31) s_w[warp_id()] = x;
        __syncthreads();
        unsigned long long before = s_base;                    // maxima of everything in earlier warps / rounds
        for (int w = 0; w < warp_id(); ++w) before = max(before, s_w[w]);
        unsigned long long ex = __shfl_up_sync(FULL_MASK, x, 1);
        ex = lane_id() == 0 ? before : max(before, ex);
        __syncthreads();
        if (i < n_chunks) A.chunk_max[i] = ex;
        if (threadIdx.x == 1023) s_base = max(before, x);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) k_ann_prefmax(AnnArgs A)
{
    __shared__ unsigned long long s_w[8];
    const long long n_chunks = (A.n_a + ANN_CHUNK - 1) / ANN_CHUNK;
    for (long long c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const long long i0 = c * ANN_CHUNK, i1 = min(A.n_a, i0 + ANN_CHUNK);
        unsigned long long base = A.chunk_max[c];
        for (long long j0 = i0; j0 < i1; j0 += 256) {
            const long long i = j0 + threadIdx.x;
            const int raw = i < i1 ? max(A.a_start[i], A.a_end[i]) : INT32_MIN;
            unsigned long long x = i < i1 ? ann_pack(A.a_chrom[i], raw) : 0;
            int bm = raw;                                               // a warp covers exactly one skip block
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) bm = max(bm, __shfl_xor_sync(FULL_MASK, bm, d));
            if (lane_id() == 0 && i < i1) A.blkmax[i / ANN_BLK] = bm;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { unsigned long long y = __shfl_up_sync(FULL_MASK, x, d); if (lane_id() >= d) x = max(x, y); }
            if (lane_id() == 31) s_w[warp_id()] = x;
            __syncthreads();
            unsigned long long before = base;
            for (int w = 0; w < warp_id(); ++w) before = max(before, s_w[w]);
            x = max(x, before);
            if (i < i1) A.prefmax[i] = x;
            unsigned long long tot = base;
            for (int w = 0; w < 8; ++w) tot = max(tot, s_w[w]);
            base = tot;
            __syncthreads();
        }
    }
}

// |BaseGenomicRange.distance| (genomicranges.py:439-458) on one chromosome; the caller has checked the chromosomes.
// Either difference is positive and below 2^32, so it is exact in unsigned 32-bit arithmetic; the caller compares
// with min(distance, 2^32 - 1). Only "is it zero" matters beyond that (intersect() :479).
__device__ __forceinline__ unsigned ann_distance(int qs, int qe, int as, int ae)
{
    if (qe < as) return (unsigned)as - (unsigned)qe;
    if (ae < qs) return (unsigned)qs - (unsigned)ae;
    return 0u;
}

__device__ __forceinline__ bool ann_strand_ok(int strandness, int qstrand, int astrand)
{   // find_neighbours :670-677; strands are +1 / -1 / 0 ('.')
    return !strandness || qstrand == 0 || astrand == qstrand || astrand == 0;
}

// Jaccard index and overlap coefficient (calc_JI :497-521, calc_OC :523-547) with the reference's intersect()
// (:460-495): the other range's start/end are taken only when STRICTLY inside self, so a range that merely touches
// self "intersects" it in all of self. Python int / int is the correctly rounded quotient; below 2^53 that is
// __ddiv_rn of the two exact doubles. ZeroDivisionError -> 0.
__device__ __forceinline__ void ann_ji_oc(int qs, int qe, int as, int ae, unsigned dist, double* ji, double* oc)
{
    if (dist != 0) { *ji = 0.0; *oc = 0.0; return; }
    const long long is = (qs < as && as < qe) ? as : qs;
    const long long ie = (qs < ae && ae < qe) ? ae : qe;
    const long long li = ie - is, l0 = (long long)qe - qs, l1 = (long long)ae - as;
    const long long uni = l0 + l1 - li, mn = l0 < l1 ? l0 : l1;
    *ji = uni == 0 ? 0.0 : __ddiv_rn((double)li, (double)uni);
    *oc = mn == 0 ? 0.0 : __ddiv_rn((double)li, (double)mn);
}

// One LANE per ortholog for the searches (32 binary searches per warp run in lockstep), the whole WARP for every
// block of ranges that has to be looked at: the lane owning the ortholog broadcasts it, 32 lanes test 32 ranges, one
// ballot orders the survivors. (First version: 8 lanes per ortholog for everything — the four groups of a warp
// diverged in every loop and the kernel was bound by instruction issue, 2.0 ms per pass for 4M orthologs.)
template <bool WRITE>
__global__ void __launch_bounds__(ANN_THREADS) k_annotate(AnnArgs A)
{
    const int lane = lane_id();
    const unsigned lt_mask = (1u << lane) - 1u;
    const unsigned dmax = A.distance > 0xffffffffll ? 0xffffffffu : (unsigned)A.distance;
    const long long warps = (long long)gridDim.x * (ANN_THREADS / 32);
    for (long long q0 = ((long long)blockIdx.x * (ANN_THREADS / 32) + warp_id()) * 32; q0 < A.n_q; q0 += warps * 32) {
        const long long q = q0 + lane;
        const bool valid = q < A.n_q;
        int qc = 0, qs = 0, qe = 0, qstr = 0, lo = 0, idx = 0;
        if (valid) { qc = A.q_chrom[q]; qs = A.q_start[q]; qe = A.q_end[q]; qstr = A.q_strand[q]; }
        if (!WRITE) {
            if (valid) {
                const int qn = A.q_name[q];
                // idx = bisect_left(annotation, self) on (chrom, start, end, name)
                int l = 0, h = (int)A.n_a;
                while (l < h) {
                    const int m = (l + h) >> 1;
                    const int c = A.a_chrom[m];
                    bool less;
                    if (c != qc) less = c < qc;
                    else {
                        const int s = A.a_start[m];
                        if (s != qs) less = s < qs;
                        else {
                            const int e = A.a_end[m];
                            less = e != qe ? e < qe : A.a_name[m] < qn;
                        }
                    }
                    if (less) l = m + 1; else h = m;
                }
                idx = l;
                // first left candidate: running max of (chrom, max(start, end)) >= (self.chrom, t). A left range is a
                // neighbour through distance()'s second branch only if end >= self.start - d, and through its first
                // branch (self.end < start, possible on the left only when self is inverted) only if start > self.end.
                const long long t64 = min((long long)qs - A.distance, (long long)qe + 1);
                if (t64 > INT32_MAX) lo = idx;
                else {
                    const unsigned long long key = ann_pack(qc, t64 < INT32_MIN ? INT32_MIN : (int)t64);
                    l = 0; h = idx;
                    while (l < h) { const int m = (l + h) >> 1; if (A.prefmax[m] < key) l = m + 1; else h = m; }
                    lo = l;
                }
            }
        } else if (valid) {
            const int4 sp = A.span[q];
            lo = sp.x; idx = sp.y;
        }
        long long w = (WRITE && valid) ? A.off[q] : 0;
        int n = 0;
        // ---- left side: every candidate in [lo, idx) is on self's chromosome. The range is walked in blocks of ANN_BLK
        // ranges; a block whose largest max(start, end) is below the threshold cannot hold a neighbour and is skipped (a
        // few long genes keep `lo` hundreds of ranges behind idx, but almost every block in between ends before self
        // starts). Each lane finds the next block worth reading for its ortholog, then the warp reads those blocks.
        const long long t64 = min((long long)qs - A.distance, (long long)qe + 1);
        const int t = t64 < INT32_MIN ? INT32_MIN : (int)min(t64, (long long)INT32_MAX);
        int b = lo / ANN_BLK;
        const int b_last = lo < idx ? (idx - 1) / ANN_BLK : -1;
        for (;;) {
            bool have = false;
            while (b <= b_last) { if (A.blkmax[b] >= t) { have = true; break; } ++b; }
            unsigned m = __ballot_sync(FULL_MASK, have);
            if (!m) break;
            while (m) {
                const int L = __ffs(m) - 1;
                m &= m - 1;
                const int qs_L = __shfl_sync(FULL_MASK, qs, L), qe_L = __shfl_sync(FULL_MASK, qe, L);
                const int qstr_L = __shfl_sync(FULL_MASK, qstr, L);
                const int lo_L = __shfl_sync(FULL_MASK, lo, L), idx_L = __shfl_sync(FULL_MASK, idx, L);
                const int k = __shfl_sync(FULL_MASK, b, L) * ANN_BLK + lane;
                bool keep = false;
                int as = 0, ae = 0;
                if (k >= lo_L && k < idx_L) {
                    as = A.a_start[k]; ae = A.a_end[k];
                    keep = ann_distance(qs_L, qe_L, as, ae) <= dmax && ann_strand_ok(A.strandness, qstr_L, A.a_strand[k]);
                }
                const unsigned bal = __ballot_sync(FULL_MASK, keep);
                if (WRITE) {
                    const long long w_L = __shfl_sync(FULL_MASK, w, L);
                    if (keep) {
                        const long long o = w_L + __popc(bal & lt_mask);
                        A.nb_idx[o] = k; A.pair_q[o] = (int)(q0 + L);
                    }
                }
                if (lane == L) { const int c = __popc(bal); n += c; w += c; }
            }
            if (have) ++b;
        }
        // ---- right side: the run of neighbours that starts at idx, 32 ranges per step
        int rk = idx;
        bool open = valid;
        for (;;) {
            unsigned m = __ballot_sync(FULL_MASK, open);
            if (!m) break;
            while (m) {
                const int L = __ffs(m) - 1;
                m &= m - 1;
                const int qc_L = __shfl_sync(FULL_MASK, qc, L);
                const int qs_L = __shfl_sync(FULL_MASK, qs, L), qe_L = __shfl_sync(FULL_MASK, qe, L);
                const int qstr_L = __shfl_sync(FULL_MASK, qstr, L);
                const long long k = (long long)__shfl_sync(FULL_MASK, rk, L) + lane;
                bool nb = false, ok = false;
                int as = 0, ae = 0;
                if (k < A.n_a && A.a_chrom[k] == qc_L) {
                    as = A.a_start[k]; ae = A.a_end[k];
                    nb = ann_distance(qs_L, qe_L, as, ae) <= dmax;
                    ok = nb && ann_strand_ok(A.strandness, qstr_L, A.a_strand[k]);
                }
                const unsigned bn = __ballot_sync(FULL_MASK, nb);
                const int run = bn == FULL_MASK ? 32 : __ffs(~bn) - 1;      // neighbours before the first miss
                const bool keep = ok && lane < run;
                const unsigned bal = __ballot_sync(FULL_MASK, keep);
                if (WRITE) {
                    const long long w_L = __shfl_sync(FULL_MASK, w, L);
                    if (keep) {
                        const long long o = w_L + __popc(bal & lt_mask);
                        A.nb_idx[o] = (int)k; A.pair_q[o] = (int)(q0 + L);
                    }
                }
                if (lane == L) {
                    const int c = __popc(bal);
                    n += c; w += c;
                    if (run < 32) open = false; else rk += 32;
                }
            }
        }
        if (!WRITE && valid) { A.cnt[q] = n; A.span[q] = make_int4(lo, idx, 0, 0); }
    }
}

// JI / OC of every (ortholog, neighbour) pair, one thread per pair. Kept out of k_annotate<true>: there the two fp64
// divisions ran under divergence, once per block that held a neighbour, and cost more than the scan itself.
__global__ void __launch_bounds__(256) k_ann_ji_oc(AnnArgs A)
{
    for (long long p = (long long)blockIdx.x * 256 + threadIdx.x; p < A.n_pairs; p += (long long)gridDim.x * 256) {
        const int q = A.pair_q[p], k = A.nb_idx[p];
        const int qs = A.q_start[q], qe = A.q_end[q], as = A.a_start[k], ae = A.a_end[k];
        double ji, oc;
        ann_ji_oc(qs, qe, as, ae, ann_distance(qs, qe, as, ae), &ji, &oc);
        A.ji[p] = ji; A.oc[p] = oc;
    }
}

}  // namespace o2a
