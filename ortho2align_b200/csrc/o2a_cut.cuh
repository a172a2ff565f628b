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
// (cheap) transform and writes with a ballot/popc rank inside the warp. A warp walks its tile in chunks of
// CUT_CHUNK HSPs held in registers (all loads of a chunk are issued before the first use; coalesced 128 B rows).
//
// Alignments come in two classes, each with its own pair of kernels:
//   fast    — only query boundaries are set, qleft <= qright (what `_align_syntenies` asks for,
//             pipeline.py:530-531) and all ranges fit 30 bits: every HSP is classified keep / drop / straddle
//             with a handful of 32-bit operations; the count pass reads the query coordinates only. An HSP that
//             straddles a boundary (< 1 %) is always kept, so the streaming passes never evaluate the cut: the
//             write pass emits it untransformed with a marker and `k_cut_fixup` evaluates the reference's
//             64-bit / float64 arithmetic for the marked HSPs afterwards, densely packed (one lane each)
//             instead of one active lane per warp.
//   general — everything else (subject boundaries, inverted or huge boundaries): the full restatement inline.
//
// A single-pass variant (chained scan with decoupled look-back over CTA groups, tile held in registers) was
// built and measured: 3.1 ms against 2.2 ms for the two passes on 10^8 HSPs — the look-back latency with only
// 16 resident warps per SM cost more than the 16 B/HSP it saved — and removed.
#pragma once
#include "o2a_device.cuh"

namespace o2a {

constexpr int CUT_TILE = 512;            // HSPs per warp-tile (granularity of the scan)
constexpr int CUT_CHUNK = 128;           // HSPs a warp holds in registers at a time
constexpr int CUT_PER_LANE = CUT_CHUNK / 32;
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
    long long* gen_tiles;        // tiles of the general-class alignments (any order)
    unsigned long long* gen_n;   // their number
    unsigned char* aln_fast;     // [n_aln] class of each alignment
    int* span_aln;               // [n_out / FIXUP_SPAN + 1] alignment holding the first output of each fix-up span
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
// General path in 64-bit integers: every boundary, every corner.
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

// Warp-uniform view of one alignment for the streaming kernels. `fast`: only query boundaries are set (what
// `_align_syntenies` asks for, pipeline.py:530-531) and every range / boundary is small enough for the
// classification keep-untouched / drop / straddle to be done in 32-bit integers.
struct CutAln {
    bool fast, qm, sm;     // qm / sm: relative and on the minus strand
    int q_off, s_off;      // genomic = minus ? off - rel : off + rel   (0 / plus when the alignment is not relative)
    int ql, qr;            // INT32_MIN / INT32_MAX stand for None
};

__device__ __forceinline__ void cut_load_aln(const CutArgs& A, long long a, CutAln& al)
{
    const bool rel = A.relative ? A.relative[a] != 0 : true;
    const bool qminus = A.q_minus[a] != 0, sminus = A.s_minus[a] != 0;
    al.qm = rel && qminus; al.sm = rel && sminus;
    const long long qo = rel ? (qminus ? A.q_end[a] : A.q_start[a]) : 0;
    const long long so = rel ? (sminus ? A.s_end[a] : A.s_start[a]) : 0;
    const long long ql = A.qleft ? A.qleft[a] : CUT_NO_BOUND, qr = A.qright ? A.qright[a] : CUT_NO_BOUND;
    const bool no_s = (!A.sleft || A.sleft[a] == CUT_NO_BOUND) && (!A.sright || A.sright[a] == CUT_NO_BOUND);
    const long long lim = 1ll << 30;
    const bool ql_ok = ql == CUT_NO_BOUND || (ql > INT32_MIN && ql < INT32_MAX);
    const bool qr_ok = qr == CUT_NO_BOUND || (qr > INT32_MIN && qr < INT32_MAX);
    const bool ordered = ql == CUT_NO_BOUND || qr == CUT_NO_BOUND || ql <= qr;     // else a straddler may still be dropped
    al.fast = no_s && ordered && qo >= 0 && qo <= lim && so >= 0 && so <= lim && ql_ok && qr_ok;
    al.q_off = (int)qo; al.s_off = (int)so;
    al.ql = ql == CUT_NO_BOUND ? INT32_MIN : (int)ql;
    al.qr = qr == CUT_NO_BOUND ? INT32_MAX : (int)qr;
}

// tiles per alignment (int, scanned into tile_off by the host-side launcher)
__global__ void k_cut_tiles_per_aln(long long n_aln, const long long* __restrict__ hsp_off, int* __restrict__ n_tiles,
                                    long long* __restrict__ n_aln_dev)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) *n_aln_dev = n_aln;          // length of the scan that follows
    for (long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x; a < n_aln; a += (long long)gridDim.x * blockDim.x)
        n_tiles[a] = (int)((hsp_off[a + 1] - hsp_off[a] + CUT_TILE - 1) / CUT_TILE);
}

// tile -> alignment map, per-alignment accumulators, alignment class; the (few) tiles of general-class alignments
// are listed so that k_cut_general touches nothing else
__global__ void k_cut_fill_tile_aln(CutArgs A, int* __restrict__ tile_aln)
{
    for (long long a = blockIdx.x; a < A.n_aln; a += gridDim.x) {
        CutAln al;
        cut_load_aln(A, a, al);
        const long long t0 = A.tile_off[a], t1 = A.tile_off[a + 1];
        for (long long t = t0 + threadIdx.x; t < t1; t += blockDim.x) tile_aln[t] = (int)a;
        if (!al.fast)
            for (long long t = t0 + threadIdx.x; t < t1; t += blockDim.x) A.gen_tiles[atomicAdd(A.gen_n, 1ull)] = t;
        if (threadIdx.x == 0) { A.max_q[a] = INT32_MIN; A.max_s[a] = INT32_MIN; A.status[a] = 0; A.aln_fast[a] = al.fast ? 1 : 0; }
    }
}

constexpr unsigned char CUT_MARK = 0x80;     // ocut bit: straddling HSP emitted untransformed, k_cut_fixup pending

// keep / drop / straddle of one HSP of a fast alignment (alignment.py:758-775 with qleft <= qright):
// 0 = dropped, 1 = kept untouched, 2 = straddles qleft and / or qright (kept, to be cut).
// T = int when the relative coordinates fit 30 bits (always, in practice), long long otherwise.
template <typename T>
__device__ __forceinline__ int cut_classify(const CutAln& al, T qs, T qe, T& gqs, T& gqe)
{
    gqs = al.qm ? (T)al.q_off - qs : (T)al.q_off + qs;
    gqe = al.qm ? (T)al.q_off - qe : (T)al.q_off + qe;
    const bool left_of_ql = gqs < (T)al.ql, right_of_qr = gqe > (T)al.qr;
    const bool drop_l = left_of_ql && gqe < (T)al.ql, cut_l = left_of_ql && !drop_l;
    const bool drop_r = right_of_qr && gqs > (T)al.qr, cut_r = right_of_qr && !drop_r;
    if (drop_l || (!cut_l && drop_r)) return 0;
    return (cut_l || cut_r) ? 2 : 1;
}

// Fast class, pass 1 (WRITE = false): kept HSPs per tile from qstart / qend alone.
// Pass 2 (WRITE = true): genomic coordinates of the kept HSPs in order; straddlers raw + CUT_MARK.
template <bool WRITE>
__global__ void __launch_bounds__(CUT_WARPS * 32, 4) k_cut_fast(CutArgs A)
{
    const int lane = lane_id();
    const long long n_tiles = A.tile_off[A.n_aln];
    const long long warps = (long long)gridDim.x * CUT_WARPS;
    for (long long t = (long long)blockIdx.x * CUT_WARPS + warp_id(); t < n_tiles; t += warps) {
        const long long a = A.tile_aln[t];
        if (!A.aln_fast[a]) continue;                               // k_cut_general's tile
        CutAln al;
        cut_load_aln(A, a, al);
        const long long h0 = A.hsp_off[a] + (t - A.tile_off[a]) * CUT_TILE;
        const int n_tile = (int)min((long long)CUT_TILE, A.hsp_off[a + 1] - h0);
        // tile-relative addressing: one 64-bit base per array, 32-bit offsets from there on
        const int* pqs = A.qs + h0 + lane; const int* pqe = A.qe + h0 + lane;
        const int* pss = A.ss + h0 + lane; const int* pse = A.se + h0 + lane;
        const double* psc = A.score + h0 + lane;
        const unsigned char* pin = A.is_int ? A.is_int + h0 + lane : nullptr;
        const long long o0 = WRITE ? A.tile_out[t] : 0;
        int* oqs = A.oqs + o0; int* oqe = A.oqe + o0; int* oss = A.oss + o0; int* ose = A.ose + o0;
        double* osc = A.oscore + o0; unsigned char* ocut = A.ocut + o0;
        int base = 0, mq = INT32_MIN, ms = INT32_MIN;              // base: kept so far in this tile
        for (int c0 = 0; c0 < n_tile; c0 += CUT_CHUNK) {
            const int n = n_tile - c0;                              // rows k with k*32 + lane < n are inside the tile
            int rqs[CUT_PER_LANE], rqe[CUT_PER_LANE], rss[CUT_PER_LANE], rse[CUT_PER_LANE];
            double rsc[CUT_PER_LANE];
            unsigned char rfl[CUT_PER_LANE];
#pragma unroll
            for (int k = 0; k < CUT_PER_LANE; ++k) {
                const bool in = k * 32 + lane < n;
                rqs[k] = in ? pqs[c0 + k * 32] : 0; rqe[k] = in ? pqe[c0 + k * 32] : 0;
                if (WRITE) {
                    rss[k] = in ? pss[c0 + k * 32] : 0; rse[k] = in ? pse[c0 + k * 32] : 0;
                    rsc[k] = in ? psc[c0 + k * 32] : 0.0;
                    rfl[k] = (in && pin && !pin[c0 + k * 32]) ? 2 : 0;
                }
            }
#pragma unroll
            for (int k = 0; k < CUT_PER_LANE; ++k) {
                int gqs = 0, gqe = 0, cls = 0;
                if (k * 32 + lane < n) {
                    if ((unsigned)(rqs[k] | rqe[k]) < (1u << 30)) cls = cut_classify<int>(al, rqs[k], rqe[k], gqs, gqe);
                    else {      // beyond the 30-bit guarantee of the 32-bit arithmetic: decide in 64 bits, let the fix-up write
                        long long g0, g1;
                        cls = cut_classify<long long>(al, rqs[k], rqe[k], g0, g1) ? 2 : 0;
                    }
                    if (WRITE && cls == 1 && (unsigned)(rss[k] | rse[k]) >= (1u << 30)) cls = 2;
                }
                const unsigned m = __ballot_sync(FULL_MASK, cls != 0);
                if (WRITE && cls != 0) {
                    const int o = base + __popc(m & lanemask_lt());
                    if (cls == 1) {
                        const int gss = al.sm ? al.s_off - rss[k] : al.s_off + rss[k];
                        const int gse = al.sm ? al.s_off - rse[k] : al.s_off + rse[k];
                        oqs[o] = gqs; oqe[o] = gqe; oss[o] = gss; ose[o] = gse; ocut[o] = rfl[k];
                        mq = max(mq, gqe);
                        ms = max(ms, max(gss, gse));
                    } else {                                        // straddler (or oversized): raw values, evaluated by k_cut_fixup
                        oqs[o] = rqs[k]; oqe[o] = rqe[k]; oss[o] = rss[k]; ose[o] = rse[k]; ocut[o] = rfl[k] | CUT_MARK;
                    }
                    osc[o] = rsc[k];
                }
                base += __popc(m);
            }
        }
        if (WRITE) {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                mq = max(mq, __shfl_xor_sync(FULL_MASK, mq, d)); ms = max(ms, __shfl_xor_sync(FULL_MASK, ms, d));
            }
            if (lane == 0) {
                if (mq != INT32_MIN) { atomicMax(&A.max_q[a], mq); atomicMax(&A.max_s[a], ms); }
            }
        } else if (lane == 0) A.tile_cnt[t] = base;
    }
}

// General class: the full restatement for every HSP of the alignments k_cut_fast skips.
template <bool WRITE>
__global__ void __launch_bounds__(CUT_WARPS * 32) k_cut_general(CutArgs A)
{
    const int lane = lane_id();
    const long long n_gen = (long long)*A.gen_n;
    const long long warps = (long long)gridDim.x * CUT_WARPS;
    for (long long g = (long long)blockIdx.x * CUT_WARPS + warp_id(); g < n_gen; g += warps) {
        const long long t = A.gen_tiles[g];
        const long long a = A.tile_aln[t];
        CutRange r; CutBounds b;
        cut_load_alignment(A, a, r, b);
        const long long h0 = A.hsp_off[a] + (t - A.tile_off[a]) * CUT_TILE;
        const long long h1 = min(A.hsp_off[a + 1], h0 + CUT_TILE);
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

constexpr int FIXUP_THREADS = 256;
constexpr int FIXUP_PER_THREAD = 64;                // flag bytes per thread per iteration (4 x 16 B loads)
constexpr int FIXUP_SPAN = FIXUP_THREADS * FIXUP_PER_THREAD;      // outputs scanned per CTA iteration (< 65536)
constexpr int FIXUP_STAGE = 512;

// out_off[a] = base of the alignment's first tile; span_aln[s] = the alignment that holds output s * FIXUP_SPAN
// (k_cut_fixup brackets its alignment search with it)
__global__ void k_cut_offsets(CutArgs A)
{
    for (long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x; a <= A.n_aln; a += (long long)gridDim.x * blockDim.x) {
        const long long lo = A.tile_out[A.tile_off[a]];
        A.out_off[a] = lo;
        if (a < A.n_aln) {
            const long long hi = A.tile_out[A.tile_off[a + 1]];
            for (long long sp = (lo + FIXUP_SPAN - 1) / FIXUP_SPAN; sp * FIXUP_SPAN < hi; ++sp) A.span_aln[sp] = (int)a;
        }
    }
}

// Straddling HSPs of the fast class: find the marked outputs, pack them, one lane each through the reference's
// arithmetic (to_genomic + both query cuts in 64-bit integers / float64), results written in place.

__global__ void __launch_bounds__(FIXUP_THREADS) k_cut_fixup(CutArgs A)
{
    __shared__ unsigned short s_list[FIXUP_SPAN];
    __shared__ int s_n;
    __shared__ long long s_alo, s_ahi;
    __shared__ long long s_off[FIXUP_STAGE];            // out_off of the span's alignments, when they are few enough
    const long long n_out = A.out_off[A.n_aln];
    // last alignment a with out_off[a] <= o, searched in [lo, hi]
    auto find_aln = [&](long long o, long long lo, long long hi) {
        while (lo < hi) {
            const long long mid = (lo + hi + 1) >> 1;
            if (A.out_off[mid] <= o) lo = mid; else hi = mid - 1;
        }
        return lo;
    };
    for (long long c0 = (long long)blockIdx.x * FIXUP_SPAN; c0 < n_out; c0 += (long long)gridDim.x * FIXUP_SPAN) {
        if (threadIdx.x == 0) {
            // the outputs of this span belong to a short run of alignments, bracketed by the span table
            s_n = 0;
            s_alo = A.span_aln[c0 / FIXUP_SPAN];
            s_ahi = c0 + FIXUP_SPAN < n_out ? A.span_aln[c0 / FIXUP_SPAN + 1] : max(A.n_aln - 1, 0ll);
        }
        __syncthreads();
        // flag bytes, 16 per load (c0 is a multiple of the span, the buffer is 256 B aligned); thread-interleaved
        // so that a warp reads 512 contiguous bytes per load
#pragma unroll
        for (int v4 = 0; v4 < FIXUP_PER_THREAD / 16; ++v4) {
            const int off = (v4 * FIXUP_THREADS + threadIdx.x) * 16;
            const long long o = c0 + off;
            if (o + 16 <= n_out) {
                const uint4 v = *reinterpret_cast<const uint4*>(A.ocut + o);
                const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    unsigned marks = w[j] & 0x80808080u;
                    while (marks) {
                        const int bit = __ffs(marks) - 1;
                        marks &= marks - 1;
                        s_list[atomicAdd(&s_n, 1)] = (unsigned short)(off + j * 4 + (bit >> 3));
                    }
                }
            } else {
                for (int j = 0; j < 16 && o + j < n_out; ++j)
                    if (A.ocut[o + j] & CUT_MARK) s_list[atomicAdd(&s_n, 1)] = (unsigned short)(off + j);
            }
        }
        __syncthreads();
        const int n = s_n;
        const long long alo = s_alo, ahi = s_ahi;
        const bool staged = ahi - alo < FIXUP_STAGE;
        if (staged && n > 0)
            for (long long k = threadIdx.x; k <= ahi - alo; k += FIXUP_THREADS) s_off[k] = A.out_off[alo + k];
        __syncthreads();
        for (int e = threadIdx.x; e < n; e += FIXUP_THREADS) {
            const long long o = c0 + s_list[e];
            long long a;
            if (staged) {
                int lo = 0, hi = (int)(ahi - alo);
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (s_off[mid] <= o) lo = mid; else hi = mid - 1;
                }
                a = alo + lo;
            } else a = find_aln(o, alo, ahi);
            CutRange r; CutBounds b;
            cut_load_alignment(A, a, r, b);
            CutHsp h;
            h.qs = A.oqs[o]; h.qe = A.oqe[o]; h.ss = A.oss[o]; h.se = A.ose[o]; h.score = A.oscore[o];
            h.cut = false; h.inexact = false;
            cut_transform(h, r, b);                                 // kept by construction (classified in k_cut_fast)
            const bool fits = h.qs == (int)h.qs && h.qe == (int)h.qe && h.ss == (int)h.ss && h.se == (int)h.se;
            const int st = (fits ? 0 : CUT_ST_RANGE) | (h.inexact ? CUT_ST_INEXACT : 0);
            A.oqs[o] = (int)h.qs; A.oqe[o] = (int)h.qe; A.oss[o] = (int)h.ss; A.ose[o] = (int)h.se;
            A.oscore[o] = h.score;
            A.ocut[o] = (unsigned char)((A.ocut[o] & 2) | (h.cut ? 1 : 0));
            atomicMax(&A.max_q[a], (int)h.qe);
            atomicMax(&A.max_s[a], (int)max(h.ss, h.se));
            if (st) atomicOr(&A.status[a], st);
        }
        __syncthreads();
    }
}

// qlen = max(qend) + 1, slen = max(max sstart, max send) + 1; an alignment left without HSPs gets the
// reference's default vertex HSPVertex(0,0,0,0,0): qlen = slen = 1 (alignment.py:477-487)
__global__ void k_cut_finish(CutArgs A)
{
    for (long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x; a < A.n_aln; a += (long long)gridDim.x * blockDim.x) {
        const int mq = A.max_q[a], ms = A.max_s[a];
        A.qlen[a] = mq == INT32_MIN ? 1 : (long long)mq + 1;
        A.slen[a] = ms == INT32_MIN ? 1 : (long long)ms + 1;
    }
}

// ---- exclusive scan int -> int64 over up to ~10^7 items: per-chunk sums, scan of the sums, per-chunk rescan ----
constexpr int SCAN_CHUNK = 1024;

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
