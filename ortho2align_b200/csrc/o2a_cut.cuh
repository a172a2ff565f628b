// o2a_cut.cuh — producer side of the alignment JSON on the GPU.
//
// What `_align_syntenies` (pipeline.py:525-532) does to the HSPs of every BLAST table before the
// build step ever sees them, for a whole batch of alignments in one pass over HBM:
//
//   alignment.to_genomic()                          genomicranges.py:1108-1129
//   alignment.cut_coordinates(qleft, qright, ...)   alignment.py:734-803 (+ _process_boundary :267-333)
//   Alignment.__init__ qlen / slen                  alignment.py:477-487
//
// The output is the columnar layout `o2a_build_batch` consumes (hsp_off, qstart..send, score, qlen,
// slen), so a caller can chain cut -> build on the device without the JSON round trip.
//
// Work decomposition: an alignment is split into tiles of CUT_TILE consecutive HSPs, one warp per tile.
// Kept HSPs keep their input order, so the output position of an HSP is the number of kept HSPs before it
// in the batch: pass 1 counts per tile, a scan turns the counts into tile bases, pass 2 recomputes the
// (cheap) transform and writes with a ballot/popc rank inside the warp. Both passes stream the inputs with
// fully coalesced 128 B warp loads; nothing but the per-alignment ranges is re-read from cache.
#pragma once
#include "o2a_device.cuh"

namespace o2a {

constexpr int CUT_TILE = 256;            // HSPs per warp-tile
constexpr int CUT_WARPS = 8;             // warps per CTA
constexpr long long CUT_NO_BOUND = INT64_MIN;
constexpr int CUT_ST_RANGE = 1;          // a kept coordinate does not fit int32
constexpr int CUT_ST_INEXACT = 2;        // an intermediate beyond 2^53: Python's exact integers would differ

struct CutArgs {
    long long n_aln;
    const long long* hsp_off;
    const int *qs, *qe, *ss, *se;
    const double* score;
    const unsigned char* is_int;                   // null: every input score is an int
    const long long *q_start, *q_end, *s_start, *s_end;
    const signed char *q_minus, *s_minus;
    const unsigned char* relative;                 // null: every alignment is relative
    const long long *qleft, *qright, *sleft, *sright;   // null: None for every alignment
    // tiles
    const long long* tile_off;   // [n_aln + 1] first tile of each alignment
    const int* tile_aln;         // [n_tiles]
    int* tile_cnt;               // [n_tiles] kept HSPs per tile
    const long long* tile_out;   // [n_tiles + 1] exclusive scan of tile_cnt
    // outputs
    int *oqs, *oqe, *oss, *ose;
    double* oscore;
    unsigned char* ocut;
    int *max_q, *max_s;          // per alignment, atomicMax, start at INT32_MIN
    int* status;                 // per alignment, atomicOr
    long long *out_off, *qlen, *slen;
};

enum { CUT_OR_NONE = 0, CUT_OR_DIRECT = 1, CUT_OR_REVERSE = 2 };

struct CutBounds { long long ql, qr, sl, sr; };

struct CutRange {
    long long q_start, q_end, s_start, s_end;
    bool q_minus, s_minus, relative;
};

struct CutHsp {
    long long qs, qe, ss, se;
    double score;
    bool cut;        // score multiplied by a fraction at least once: a Python float from here on
    bool inexact;
};

// round((a * b) / c + d): Python evaluates a*b exactly, divides two ints with a correctly rounded quotient,
// adds an int converted to float, and rounds half to even (alignment.py:289-292, 297-300, 303-306, 317-320).
// With |a*b|, |c|, |d| < 2^53 the double operations below are the same operations.
__device__ __forceinline__ long long cut_interpolate(long long a, long long b, long long c, long long d, bool& inexact)
{
    const double limit = 9007199254740992.0;
    double pa = fabs(__dmul_rn((double)a, (double)b));
    if (pa >= limit || fabs((double)c) >= limit || fabs((double)d) >= limit) inexact = true;
    double v = __dadd_rn(__ddiv_rn((double)(a * b), (double)c), (double)d);
    return (long long)rint(v);
}

__device__ __forceinline__ void cut_scale(CutHsp& h, long long num, long long den)
{
    h.score = __dmul_rn(h.score, __ddiv_rn((double)num, (double)den));      // hsp.score *= fraction, alignment.py:330
    h.cut = true;
}

// to_genomic (genomicranges.py:1112-1128) then _cut_hsps (alignment.py:757-803). Returns false if dropped.
__device__ __forceinline__ bool cut_transform(CutHsp& h, const CutRange& r, const CutBounds& b)
{
    if (r.relative) {
        if (!r.q_minus) { h.qs += r.q_start; h.qe += r.q_start; } else { h.qs = r.q_end - h.qs; h.qe = r.q_end - h.qe; }
        if (!r.s_minus) { h.ss += r.s_start; h.se += r.s_start; } else { h.ss = r.s_end - h.ss; h.se = r.s_end - h.se; }
    }
    // orientation as hsp.copy() computes it at the top of _cut_hsps (alignment.py:107-110, 757); never refreshed
    const int orient = (h.qe - h.qs > 0) ? ((h.se - h.ss > 0) ? CUT_OR_DIRECT : CUT_OR_REVERSE) : CUT_OR_NONE;
    if (b.ql != CUT_NO_BOUND) {                                            // :758-766
        if (h.qs < b.ql && b.ql <= h.qe) {                                 // 'qleft' :287-293
            const long long span = h.qe - h.qs;
            const long long num = h.qe - b.ql;
            h.ss = cut_interpolate(b.ql - h.qs, h.se - h.ss, span, h.ss, h.inexact);
            h.qs = b.ql;
            cut_scale(h, num, span);
        } else if (b.ql > h.qs) return false;
    }
    if (b.qr != CUT_NO_BOUND) {                                            // :767-775
        if (h.qs <= b.qr && b.qr < h.qe) {                                 // 'qright' :294-301
            const long long span = h.qe - h.qs;
            const long long num = b.qr - h.qs;
            h.se = cut_interpolate(num, h.se - h.ss, span, h.ss, h.inexact);
            h.qe = b.qr;
            cut_scale(h, num, span);
        } else if (h.qe > b.qr) return false;
    }
    if (b.sl != CUT_NO_BOUND) {                                            // :776-789, an if/elif chain: order matters
        bool straddle;
        if (h.ss < b.sl && b.sl <= h.se) straddle = true;
        else if (b.sl <= h.ss && orient == CUT_OR_DIRECT) straddle = false;
        else if (h.se < b.sl && b.sl <= h.ss) straddle = true;
        else if (b.sl <= h.se && orient == CUT_OR_REVERSE) straddle = false;
        else return false;
        if (straddle) {                                                    // 'sleft' :302-315
            const long long qpoint = cut_interpolate(b.sl - h.ss, h.qe - h.qs, h.se - h.ss, h.qs, h.inexact);
            if (orient == CUT_OR_DIRECT) { cut_scale(h, h.se - b.sl, h.se - h.ss); h.qs = qpoint; h.ss = b.sl; }
            else { cut_scale(h, h.ss - b.sl, h.ss - h.se); h.qe = qpoint; h.se = b.sl; }
        }
    }
    if (b.sr != CUT_NO_BOUND) {
        // :790-802 — the reference collects the survivors of this boundary in new_hsps but returns the previous
        // list (:803), so nothing is dropped here; HSPs straddling the boundary are still cut (in place).
        const bool straddle = (h.ss <= b.sr && b.sr < h.se)
                              || (!(h.se <= b.sr && orient == CUT_OR_DIRECT) && h.se <= b.sr && b.sr < h.ss);
        if (straddle) {                                                    // 'sright' :316-329
            const long long qpoint = cut_interpolate(b.sr - h.ss, h.qe - h.qs, h.se - h.ss, h.qs, h.inexact);
            if (orient == CUT_OR_DIRECT) { cut_scale(h, b.sr - h.ss, h.se - h.ss); h.qe = qpoint; h.se = b.sr; }
            else { cut_scale(h, b.sr - h.se, h.ss - h.se); h.qs = qpoint; h.ss = b.sr; }
        }
    }
    return true;
}

__device__ __forceinline__ void cut_load_alignment(const CutArgs& A, long long a, CutRange& r, CutBounds& b)
{
    r.relative = A.relative ? A.relative[a] != 0 : true;
    r.q_start = A.q_start[a]; r.q_end = A.q_end[a]; r.s_start = A.s_start[a]; r.s_end = A.s_end[a];
    r.q_minus = A.q_minus[a] != 0; r.s_minus = A.s_minus[a] != 0;
    b.ql = A.qleft ? A.qleft[a] : CUT_NO_BOUND; b.qr = A.qright ? A.qright[a] : CUT_NO_BOUND;
    b.sl = A.sleft ? A.sleft[a] : CUT_NO_BOUND; b.sr = A.sright ? A.sright[a] : CUT_NO_BOUND;
}

// tiles per alignment (int, scanned into tile_off by the host-side launcher)
__global__ void k_cut_tiles_per_aln(long long n_aln, const long long* __restrict__ hsp_off, int* __restrict__ n_tiles)
{
    for (long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x; a < n_aln; a += (long long)gridDim.x * blockDim.x)
        n_tiles[a] = (int)((hsp_off[a + 1] - hsp_off[a] + CUT_TILE - 1) / CUT_TILE);
}

__global__ void k_cut_fill_tile_aln(long long n_aln, const long long* __restrict__ tile_off, int* __restrict__ tile_aln,
                                    int* __restrict__ max_q, int* __restrict__ max_s, int* __restrict__ status)
{
    for (long long a = blockIdx.x; a < n_aln; a += gridDim.x) {
        for (long long t = tile_off[a] + threadIdx.x; t < tile_off[a + 1]; t += blockDim.x) tile_aln[t] = (int)a;
        if (threadIdx.x == 0) { max_q[a] = INT32_MIN; max_s[a] = INT32_MIN; status[a] = 0; }
    }
}

// Pass 1 (WRITE = false): kept HSPs per tile. Pass 2 (WRITE = true): the same transform, written in order.
template <bool WRITE>
__global__ void __launch_bounds__(CUT_WARPS * 32) k_cut(CutArgs A)
{
    const int lane = lane_id();
    const long long n_tiles = A.tile_off[A.n_aln];
    const long long warps = (long long)gridDim.x * CUT_WARPS;
    for (long long t = (long long)blockIdx.x * CUT_WARPS + warp_id(); t < n_tiles; t += warps) {
        const long long a = A.tile_aln[t];
        const long long h0 = A.hsp_off[a] + (t - A.tile_off[a]) * CUT_TILE;
        const long long h1 = min(A.hsp_off[a + 1], h0 + CUT_TILE);
        CutRange r; CutBounds b;
        cut_load_alignment(A, a, r, b);
        long long base = WRITE ? A.tile_out[t] : 0;
        int cnt = 0, mq = INT32_MIN, ms = INT32_MIN, st = 0;
        for (long long i0 = h0; i0 < h1; i0 += 32) {
            const long long i = i0 + lane;
            bool keep = false;
            CutHsp h;
            if (i < h1) {
                h.qs = A.qs[i]; h.qe = A.qe[i]; h.ss = A.ss[i]; h.se = A.se[i];
                h.score = WRITE ? A.score[i] : 0.0;
                h.cut = false; h.inexact = false;
                keep = cut_transform(h, r, b);
            }
            const unsigned m = __ballot_sync(FULL_MASK, keep);
            if (WRITE) {
                if (keep) {
                    const long long o = base + __popc(m & lanemask_lt());
                    const bool fits = h.qs == (int)h.qs && h.qe == (int)h.qe && h.ss == (int)h.ss && h.se == (int)h.se;
                    if (!fits) st |= CUT_ST_RANGE;
                    if (h.inexact) st |= CUT_ST_INEXACT;
                    A.oqs[o] = (int)h.qs; A.oqe[o] = (int)h.qe; A.oss[o] = (int)h.ss; A.ose[o] = (int)h.se;
                    A.oscore[o] = h.score;
                    A.ocut[o] = (unsigned char)((h.cut ? 1 : 0) | ((A.is_int && !A.is_int[i]) ? 2 : 0));
                    mq = max(mq, (int)h.qe);
                    ms = max(ms, (int)max(h.ss, h.se));
                }
                base += __popc(m);
            } else cnt += __popc(m);
        }
        if (WRITE) {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                mq = max(mq, __shfl_xor_sync(FULL_MASK, mq, d)); ms = max(ms, __shfl_xor_sync(FULL_MASK, ms, d));
                st |= __shfl_xor_sync(FULL_MASK, st, d);
            }
            if (lane == 0) {
                if (mq != INT32_MIN) { atomicMax(&A.max_q[a], mq); atomicMax(&A.max_s[a], ms); }
                if (st) atomicOr(&A.status[a], st);
            }
        } else if (lane == 0) A.tile_cnt[t] = cnt;
    }
}

// out_off, qlen = max(qend) + 1, slen = max(max sstart, max send) + 1; an alignment left without HSPs gets the
// reference's default vertex HSPVertex(0,0,0,0,0): qlen = slen = 1 (alignment.py:477-487)
__global__ void k_cut_finish(CutArgs A)
{
    for (long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x; a <= A.n_aln; a += (long long)gridDim.x * blockDim.x) {
        const long long t0 = A.tile_off[a < A.n_aln ? a : A.n_aln];
        A.out_off[a] = A.tile_out[t0];
        if (a < A.n_aln) {
            const int mq = A.max_q[a], ms = A.max_s[a];
            A.qlen[a] = mq == INT32_MIN ? 1 : (long long)mq + 1;
            A.slen[a] = ms == INT32_MIN ? 1 : (long long)ms + 1;
        }
    }
}

// ---- exclusive scan int -> int64 over up to ~10^7 items: per-chunk sums, scan of the sums, per-chunk rescan ----
constexpr int SCAN_CHUNK = 8192;

__global__ void __launch_bounds__(256) k_scan_chunk_sums(const long long* __restrict__ n_ptr, const int* __restrict__ v,
                                                         long long* __restrict__ sums)
{
    __shared__ long long s_w[8];
    const long long n = *n_ptr;
    for (long long c = blockIdx.x; c * SCAN_CHUNK < n; c += gridDim.x) {
        const long long i0 = c * SCAN_CHUNK, i1 = min(n, i0 + SCAN_CHUNK);
        long long s = 0;
        for (long long i = i0 + threadIdx.x; i < i1; i += 256) s += v[i];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(FULL_MASK, s, d);
        if (lane_id() == 0) s_w[warp_id()] = s;
        __syncthreads();
        if (threadIdx.x == 0) { long long tot = 0; for (int w = 0; w < 8; ++w) tot += s_w[w]; sums[c] = tot; }
        __syncthreads();
    }
}

// single CTA: exclusive scan of the chunk sums in place; sums[n_chunks] = grand total
__global__ void __launch_bounds__(1024) k_scan_sums(const long long* __restrict__ n_ptr, long long* __restrict__ sums)
{
    __shared__ long long s_w[33];
    __shared__ long long s_base;
    const long long n_chunks = (*n_ptr + SCAN_CHUNK - 1) / SCAN_CHUNK;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (long long i0 = 0; i0 < n_chunks; i0 += 1024) {
        const long long i = i0 + threadIdx.x;
        const long long v = i < n_chunks ? sums[i] : 0;
        long long x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { long long y = __shfl_up_sync(FULL_MASK, x, d); if (lane_id() >= d) x += y; }
        if (lane_id() == 31) s_w[warp_id()] = x;
        __syncthreads();
        if (warp_id() == 0) {
            long long tw = s_w[lane_id()], s = tw;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { long long y = __shfl_up_sync(FULL_MASK, s, d); if (lane_id() >= d) s += y; }
            s_w[lane_id()] = s - tw;
            if (lane_id() == 31) s_w[32] = s;
        }
        __syncthreads();
        const long long base = s_base;
        if (i < n_chunks) sums[i] = base + s_w[warp_id()] + x - v;
        __syncthreads();
        if (threadIdx.x == 0) s_base = base + s_w[32];
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[n_chunks] = s_base;
}

__global__ void __launch_bounds__(256) k_scan_chunks(const long long* __restrict__ n_ptr, const int* __restrict__ v,
                                                     const long long* __restrict__ sums, long long* __restrict__ out)
{
    __shared__ int s_scan[256 / 32 + 2];
    const long long n = *n_ptr;
    const long long n_chunks = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
    for (long long c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const long long i0 = c * SCAN_CHUNK, i1 = min(n, i0 + SCAN_CHUNK);
        long long base = sums[c];
        for (long long j0 = i0; j0 < i1; j0 += 256) {
            const long long i = j0 + threadIdx.x;
            const int x = i < i1 ? v[i] : 0;
            int tot;
            const int ex = cta_exclusive_scan(x, s_scan, &tot);
            if (i < i1) out[i] = base + ex;
            base += tot;
            __syncthreads();
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = sums[n_chunks];
}

}  // namespace o2a
