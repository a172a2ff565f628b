// o2a_front.cuh — K3 + K4 + K5 fused: score filter, p-values, ranks, q-values and keep flags of ONE
// alignment by ONE warp, survivors never leaving the SM between the stages.
//
//   alignment.filter_by_score(isf(alpha))     strict '>'            alignment.py:618-646, pipeline.py:765
//   pvals = background.sf(scores)                                   fitting.py:184-193 / 282-310, pipeline.py:766
//   q = (m * p) / ranking(p); keep q <= alpha  (no cummin)          fitting.py:359-364, pipeline.py:769-776
//
// Why fused: the streaming filter is HBM-bound with idle issue slots, the p/q stage is issue-bound with idle
// HBM (round 1: 0.334 ms + 0.338 ms as two kernels on BASELINE config 2). With one warp per alignment doing
// both, 32 independent warps per SM are always in different phases, so the score stream of some hides under the
// rank arithmetic of others, one launch and one global round trip of the survivor scores disappear, and no
// barrier is left (everything is warp-synchronous).
//
// Work distribution: groups of consecutive alignments (about one lncRNA's worth, FrontArgs::chunk_aln) are handed to
// the warps through a ticket counter. Consecutive alignments belong to the same lncRNA (one background fit, one
// threshold), so what depends on the lncRNA only is kept in shared memory across them:
//   ptab[s]  p = sf(floor(thr) + s) of the integer score classes ("direct slots"), computed lazily up to the
//            highest slot any alignment of the lncRNA has needed so far (hist: one table search per slot and
//            lncRNA instead of one per class and alignment; KDE: one value-compressed sum per slot and lncRNA)
//   grp[s]   first slot of the run of equal p that s belongs to (sf is non-increasing in the score; the table is
//            CHECKED to be non-increasing — anything else sends the lncRNA's alignments to k_pq)
//
// The filter streams the alignment's scores (float64: 32-byte loads covering 1 KB of consecutive memory per warp request,
// a 4 KB tile in flight per warp) and writes the survivors IN ORDER (popc + one packed warp scan per tile) into the alignment's slot [hsp_off[a], hsp_off[a] + n_surv) of the
// survivor arrays.
//
// Ranks (np.argsort(kind='stable') of the p-values, rank_k = 1 + #{p_j < p_k} + #{j < k : p_j == p_k}) come from
// KEYS that order the survivors by p without comparing p-values pairwise:
//   direct survivor in slot s                          key 2 * grp[s]        (all members share one p)
//   other score x (cut HSP, huge score) with p equal to a neighbouring slot's   that slot's key
//   other score strictly between slots s and s + 1     key 2 * s + 1         (p[s + 1] < p < p[s] is CHECKED)
// A larger key means a strictly smaller p, so #{p_j < p_k} is a suffix sum of the key histogram, and the ties of
// an even key are counted in list order while the histogram is built (match.any, like k_pq's ordered walk). The
// few members of one odd key are ordered among themselves by (p, list position).
// Alignments with more than FRONT_SCAP survivors or more than FRONT_OVER non-class scores, and everything that
// fails a monotonicity check, are only filtered here and queued for the CTA-wide k_pq (o2a_score.cuh), which
// makes no such assumption.
#pragma once
#include "o2a_chain.cuh"
#include "o2a_device.cuh"
#include "o2a_fit.cuh"
#include "o2a_score.cuh"

namespace o2a {

constexpr int FRONT_WARPS = 4;            // warps per CTA; every warp owns its alignments
#ifndef O2A_FRONT_CTAS
#define O2A_FRONT_CTAS 8
#endif
constexpr int FRONT_CTAS_PER_SM = O2A_FRONT_CTAS;      // 32 warps per SM
constexpr int FRONT_SCAP = 192;           // survivors per alignment kept on chip
constexpr int FRONT_DIRECT = 128;         // integer score classes floor(thr) + [0, 128)
constexpr int FRONT_OVER = 32;            // other scores (cut HSPs, huge values): one lane each
constexpr int FRONT_KEYS = 2 * FRONT_DIRECT;
constexpr int FRONT_PASS = 32;            // passing exceptions of a narrow-score alignment kept on chip
constexpr int FRONT_XR = FRONT_SCAP / 32; // survivor scores per lane

struct FrontArgs {
    PqArgs P;                             // geometry, fit tables, p / q / keep outputs (n_surv, surv_x are outputs here)
    const double* score;                  // float64 layout (score_bits == 0)
    const void* score_q;                  // narrow layout: int8 / int16 / int32 + exception lists
    const long long* exc_off; const int* exc_pos; const double* exc_val;
    long long n_hsp_total;
    int* n_surv; int* surv_idx; double* surv_x;
    int* pq_list; int* pq_count;          // alignments left to k_pq
    // the retained HSPs of the alignments finished here, for the chaining (slot layout, see o2a_chain.cuh):
    int* ret_slot; double* ret_score;     // position in the survivor list, score
    int4* ret_coord;                      // coordinates — fetched here only when fetch_coords != 0 (device memory)
    int fetch_coords;
    const int4* coords4; const int* qstart; const int* qend; const int* sstart; const int* send;
    const long long* qlen; const long long* slen;
    int* mid_list; int* mid_count;        // alignments the warp path of the chaining cannot take (-> k_chain_small)
    int chunk_aln; int* ticket;           // groups of chunk_aln (>= 1) consecutive alignments are handed out through *ticket
};

struct __align__(16) FrontWarpSmem {
    double ptab[FRONT_DIRECT];            // per lncRNA: p of the direct slots [0, tab_n)
    double ox[FRONT_OVER];                // scores of the non-class survivors
    double op[FRONT_OVER];                // their p (narrow filter phase: values of the passing exceptions)
    double x[FRONT_SCAP];                 // survivor scores, list order
    int sidx[FRONT_SCAP];                 // HSP position of every survivor
    unsigned short cnt[FRONT_KEYS];       // members per key, then: survivors with a larger key (narrow filter phase: positions of the passing exceptions)
    unsigned char grp[FRONT_DIRECT];      // per lncRNA: head slot of the run of equal p
    unsigned char slot[FRONT_SCAP];       // direct slot, or FRONT_DIRECT + o for non-class survivor o
    unsigned char key[FRONT_SCAP];
    unsigned char aux[FRONT_SCAP];        // even key: earlier members of the key; odd key: o
    unsigned char kept[FRONT_SCAP];       // survivor-list positions of the retained HSPs
    unsigned char okey[FRONT_OVER];
    unsigned char oin[FRONT_OVER];        // members of the same odd key that rank before o
};

static_assert(FRONT_KEYS * sizeof(unsigned short) >= FRONT_PASS * sizeof(int) && FRONT_PASS <= FRONT_OVER, "aliases of the filter phase");
static_assert(FRONT_SCAP <= 256 && FRONT_KEYS <= 256 && FRONT_DIRECT + FRONT_OVER <= 256, "byte-sized indices");
static_assert(FRONT_KEYS == 8 * 32, "the suffix sums take eight keys per lane");

template <typename T> struct FrontTraits;
template <> struct FrontTraits<double> { static constexpr int PER = 2; };
template <> struct FrontTraits<int> { static constexpr int PER = 4; };
template <> struct FrontTraits<short> { static constexpr int PER = 8; };
template <> struct FrontTraits<signed char> { static constexpr int PER = 16; };

__device__ __forceinline__ int4 ld_stream_16(const void* p)
{
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// ---- filter phase, float64 scores: returns n_surv; survivor positions to global slots and (first FRONT_SCAP) S.sidx --
// Tile = 512 scores = 4 KB per warp in flight (see front_filter_f64 for the layout).
__device__ __forceinline__ void ld_stream_f64x4(const double* p, double& a, double& b, double& c, double& d)
{
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}

// 16 scores of one lane -> bit c set when score c passes the strict '>' (alignment.py:32-33)
__device__ __forceinline__ unsigned front_mask16(const double (&v)[16], double t)
{
    unsigned mask = 0;
#pragma unroll
    for (int c = 0; c < 16; ++c) mask |= (v[c] > t ? 1u : 0u) << c;
    return mask;
}

#ifndef O2A_FRONT_CPX
#define O2A_FRONT_CPX 1
#endif
// survivor score -> S.x[p] by an asynchronous 8-byte global->shared copy (LDGSTS: no register, no scoreboard wait at the
// point of issue); the line has just been streamed, so the copy is served by L2
__device__ __forceinline__ void front_copy_score_async(double* dst_smem, const double* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void front_copy_wait()
{
    asm volatile("cp.async.wait_all;" ::: "memory");
}

// Row layout: the 512-score tile is four rows of 128 scores and lane i owns scores [4 i, 4 i + 4) of EVERY row, so each
// 32-byte load of the warp (LDG.E.256) covers 1 KB of consecutive memory — eight full lines per request.
// v[4 q + c] / mask bit 4 q + c = score 128 q + 4 i + c of the tile. Order-preserving positions need the lanes' counts
// per row: the four counts (<= 4 each, row totals <= 128) are scanned together as the bytes of one word, so the scan
// costs what a single count would. The loads of tile t + 1 are issued right after tile t's compares, so their DRAM
// latency hides under the scan and the placement of tile t.
// (Until round 2's last change lane i owned 16 CONSECUTIVE scores: one scan of plain counts, but every request touched
// one 32-byte sector of 32 different lines, four requests per line: 0.539 ms per launch on BASELINE config 2 against
// 0.484 ms with full-line requests. Scanning four tiles' counts together as packed shuffle chains saves shuffles but
// leaves nothing to overlap the loads with: 0.67 against 0.65 ms.)
__device__ __forceinline__ void front_load_rows(const double* __restrict__ sc, long long g0, long long n_total, int t0, int lane,
                                                double t, double (&v)[16], bool inside)
{
    const double* p0 = sc + t0 + 4 * lane;
    if (inside) {
#pragma unroll
        for (int q = 0; q < 4; ++q) ld_stream_f64x4(p0 + 128 * q, v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        return;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const long long left = n_total - (g0 + t0 + 128 * q + 4 * lane);
        if (left >= 4) ld_stream_f64x4(p0 + 128 * q, v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        else {
#pragma unroll
            for (int c = 0; c < 4; ++c) v[4 * q + c] = c < left ? p0[128 * q + c] : t;
        }
    }
}

__device__ __forceinline__ int front_filter_f64(const double* __restrict__ score, long long n_total, long long h0, int m,
                                                     double t, int* __restrict__ out_idx, FrontWarpSmem& S)
{
    const int lane = lane_id();
    const int sh = (int)(h0 & 3);
    const double* sc = score + h0 - sh;
    const long long g0 = h0 - sh;
    const int ms = m + sh;
    int base = 0;
    double v[16];
    front_load_rows(sc, g0, n_total, 0, lane, t, v, g0 + 512 <= n_total);
#pragma unroll 1
    for (int t0 = 0; t0 < ms; t0 += 512) {
        unsigned mask = front_mask16(v, t);
        if (t0 + 512 < ms) front_load_rows(sc, g0, n_total, t0 + 512, lane, t, v, g0 + t0 + 1024 <= n_total);
        const int e0 = t0 + 4 * lane;                  // my first score of row 0
        if (t0 + 512 > ms) {                           // the alignment ends inside this tile (warp-uniform)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int left = ms - (e0 + 128 * q);
                const unsigned keep = left >= 4 ? 0xfu : (left <= 0 ? 0u : ((1u << left) - 1u));
                mask &= ~(0xfu << (4 * q)) | (keep << (4 * q));
            }
        }
        if (t0 == 0 && lane == 0) mask &= ~((1u << sh) - 1u);
        const unsigned pk = (unsigned)__popc(mask & 0xfu) | ((unsigned)__popc(mask & 0xf0u) << 8) |
                            ((unsigned)__popc(mask & 0xf00u) << 16) | ((unsigned)__popc(mask & 0xf000u) << 24);
        unsigned incl = pk;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned y = __shfl_up_sync(FULL_MASK, incl, d); if (lane >= d) incl += y; }
        const unsigned tot = __shfl_sync(FULL_MASK, incl, 31);
        const unsigned excl = incl - pk;
        // first position of my survivors of row q, relative to `base`: the totals of the rows before + the lanes before me
        // (row totals can reach 128: the running sums over the rows need more than a byte)
        const int T0 = tot & 0xff, T1 = (tot >> 8) & 0xff, T2 = (tot >> 16) & 0xff, T3 = tot >> 24;
        const int s0 = (int)(excl & 0xff), s1 = T0 + (int)((excl >> 8) & 0xff), s2 = T0 + T1 + (int)((excl >> 16) & 0xff),
                  s3 = T0 + T1 + T2 + (int)(excl >> 24);
        const unsigned long long W = (unsigned long long)s0 | ((unsigned long long)s1 << 16) | ((unsigned long long)s2 << 32) |
                                     ((unsigned long long)s3 << 48);
        int curq = -1, p = 0;
#pragma unroll 1
        while (mask) {
            const int c = __ffs(mask) - 1;
            mask &= mask - 1;
            const int q = c >> 2;
            if (q != curq) { curq = q; p = base + (int)((W >> (16 * q)) & 0xffffull); }
            const int e = e0 + (q << 7) + (c & 3);
            if (p < FRONT_SCAP) {
                S.sidx[p] = e - sh;
#if O2A_FRONT_CPX
                front_copy_score_async(&S.x[p], sc + e);
#endif
            } else out_idx[p] = e - sh;
            ++p;
        }
        base += T0 + T1 + T2 + T3;
    }
    // the first FRONT_SCAP positions go to HBM as a few full-width stores instead of one scattered 4-byte store per
    // survivor (133 store requests per alignment on BASELINE config 2 become 5)
    __syncwarp();
    for (int k = lane; k < min(base, FRONT_SCAP); k += 32) out_idx[k] = S.sidx[k];
#if O2A_FRONT_CPX
    front_copy_wait();
#endif
    return base;
}

// One 16-byte load of narrow scores -> bit c of `msk`: score c > tq, bit c of `sent`: score c is the exception sentinel.
// int8: the 16 scores are compared inside their 32-bit words (__vcmpgts4 / __vcmpeq4: about 3.5 instructions per score
// instead of extract + two compares + two inserts each) and the per-byte flags are packed by one multiplication.
// tq is clamped to the type's maximum first: no score can exceed it, none may pass.
template <typename T>
__device__ __forceinline__ void front_cmp_row(const int4& raw, int tq, unsigned& msk, unsigned& sent)
{
    if constexpr (sizeof(T) == 1) {
        const unsigned t4 = (unsigned)(tq > 127 ? 127 : tq) & 0xffu, tb = t4 * 0x01010101u;
        const unsigned w[4] = {(unsigned)raw.x, (unsigned)raw.y, (unsigned)raw.z, (unsigned)raw.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned gt = __vcmpgts4(w[j], tb) & 0x01010101u, eq = __vcmpeq4(w[j], 0x80808080u) & 0x01010101u;
            msk |= ((gt * 0x01020408u) >> 24) << (4 * j);          // flag of byte b -> bit b
            sent |= ((eq * 0x01020408u) >> 24) << (4 * j);
        }
    } else {                                             // int16 measured no faster with __vcmpgts2 (0.886 against 0.862 ms)
        constexpr int SENT = QTraits<T>::SENT;
#pragma unroll
        for (int c = 0; c < QTraits<T>::PER; ++c) {
            const int q = q_at<T>(raw, c);
            msk |= (q > tq ? 1u : 0u) << c;
            sent |= (q == SENT ? 1u : 0u) << c;
        }
    }
}

// ---- filter phase, narrow scores (int8 / int16 / int32 + exception lists; see QTraits' contract) -----------------
// Survivor positions and scores: the first FRONT_SCAP to S.sidx / S.x (positions then to out_idx as full-width stores),
// the rest straight to out_idx / out_x.
template <typename T>
__device__ __forceinline__ int front_filter_q(const T* __restrict__ score_q, long long n_hsp_total, long long h0, int m,
                                              double t, const int* __restrict__ exc_pos, const double* __restrict__ exc_val,
                                              long long x0, int ne, int* __restrict__ out_idx, double* __restrict__ out_x,
                                              FrontWarpSmem& S)
{
    constexpr int PER = FrontTraits<T>::PER;
    constexpr int SENT = QTraits<T>::SENT;
    constexpr int ROWS = 2;                              // 16-byte loads in flight per lane
    constexpr int ROWLEN = 32 * PER;
    const int lane = lane_id();
    const unsigned lt = lanemask_lt();
    // q > thr  <=>  q > floor(thr) for integers; clamped so that the sentinel can never pass
    const double tf = floor(t);
    const int tq = tf >= 2147483647.0 ? 2147483647 : (tf <= (double)SENT ? SENT : (int)tf);
    // exceptions that pass (value > thr): positions in e_pos, values in e_val (both free during the filter phase)
    int* e_pos = reinterpret_cast<int*>(S.cnt);
    double* e_val = S.op;
    int np = 0;
    for (int i0 = 0; i0 < ne; i0 += 32) {
        const int i = i0 + lane;
        const double v = i < ne ? exc_val[x0 + i] : 0.0;
        const bool pass = i < ne && v > t;
        const unsigned b = __ballot_sync(FULL_MASK, pass);
        const int p = np + __popc(b & lt);
        if (pass && p < FRONT_PASS) { e_pos[p] = exc_pos[x0 + i]; e_val[p] = v; }
        np += __popc(b);
    }
    __syncwarp();
    const int sh = (int)(h0 & (PER - 1));
    const T* sc = score_q + h0 - sh;
    const int ms = m + sh;
    const long long g0 = h0 - sh;                        // global index of stream element 0
    int base = 0;
    // (Loading the rows of tile t + 1 before tile t is compared and placed: int8 0.632 -> 0.690 ms per C2 launch, int16
    // 0.633 -> 0.617 — not kept.)
    for (int t0 = 0; t0 < ms; t0 += ROWS * ROWLEN) {
        int4 raw[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const int e = t0 + r * ROWLEN + lane * PER;
            if (e >= ms) raw[r] = make_int4(0, 0, 0, 0);                        // masked below
            else if (g0 + e + PER <= n_hsp_total) raw[r] = ld_stream_16(sc + e);
            else {                                                              // tail of the whole score array
                T q[PER];
#pragma unroll
                for (int c = 0; c < PER; ++c) q[c] = (g0 + e + c < n_hsp_total) ? sc[e + c] : (T)SENT;
                memcpy(&raw[r], q, 16);
            }
        }
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const int e = t0 + r * ROWLEN + lane * PER;  // stream index of my first score of this row
            unsigned msk = 0, sent = 0;
            front_cmp_row<T>(raw[r], tq, msk, sent);
            if (e < sh || e + PER > ms) {                // lanes straddling the alignment borders
                unsigned inside = 0;
#pragma unroll
                for (int c = 0; c < PER; ++c) inside |= ((e + c >= sh && e + c < ms) ? 1u : 0u) << c;
                msk &= inside; sent &= inside;
            }
            unsigned sent_pass = 0;                      // exceptions of mine that pass
            while (sent) {
                const int c = __ffs(sent) - 1; sent &= sent - 1;
                const int pos = e + c - sh;
                bool ok = false;
                if (np <= FRONT_PASS) { for (int i = 0; i < np; ++i) if (e_pos[i] == pos) { ok = true; break; } }
                else {
                    int lo = 0, hi = ne;
                    while (lo < hi) { const int mid = (lo + hi) >> 1; if (exc_pos[x0 + mid] < pos) lo = mid + 1; else hi = mid; }
                    ok = lo < ne && exc_pos[x0 + lo] == pos && exc_val[x0 + lo] > t;
                }
                if (ok) sent_pass |= 1u << c;
            }
            msk |= sent_pass;
            const int cnt = __popc(msk);
            int incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(FULL_MASK, incl, d); if (lane >= d) incl += y; }
            const int rowtot = __shfl_sync(FULL_MASK, incl, 31);
            {
                // one iteration per survivor of the lane (a few per row), not one per score (a fully predicated walk over
                // the row's 16 / 8 scores cost more than the compares: 1.06 ms per C2 launch with int8 scores)
                int p = base + incl - cnt;
                const int ep = e - sh;                   // HSP position of my first score of this row
                unsigned mm = msk;
#pragma unroll 1
                while (mm) {
                    const int c = __ffs(mm) - 1; mm &= mm - 1;
                    double x = (double)q_at<T>(raw[r], c);
                    if ((sent_pass >> c) & 1u) {         // a passing exception: its exact float64 value
                        if (np <= FRONT_PASS) { for (int i = 0; i < np; ++i) if (e_pos[i] == ep + c) { x = e_val[i]; break; } }
                        else {
                            int lo = 0, hi = ne;
                            while (lo < hi) { const int mid = (lo + hi) >> 1; if (exc_pos[x0 + mid] < ep + c) lo = mid + 1; else hi = mid; }
                            x = exc_val[x0 + lo];
                        }
                    }
                    if (p < FRONT_SCAP) { S.sidx[p] = ep + c; S.x[p] = x; }
                    else { out_idx[p] = ep + c; out_x[p] = x; }
                    ++p;
                }
            }
            base += rowtot;
        }
    }
    __syncwarp();
    for (int k = lane; k < min(base, FRONT_SCAP); k += 32) out_idx[k] = S.sidx[k];       // full-width stores
    return base;
}

template <int FIT>
__device__ __forceinline__ double front_sf(double x, const PqFit& F)
{
    if (FIT == 0) return hist_sf(x, F.f, F.tk, F.tc);
    if (F.fitter == 1) return kde_sf_compressed(x, F.uval, F.ucnt, F.unv, F.kfac, F.nbg);
    return emp_sf_compressed(x, F.uval, F.ucnt, F.unv, F.nbg);
}

template <int FIT>
__device__ __forceinline__ PqFit front_fit_of(const PqArgs& P, int l)
{
    PqFit F; F.fitter = FIT == 0 ? 0 : P.fitter; F.tk = nullptr; F.tc = nullptr; F.uval = nullptr; F.ucnt = nullptr;
    F.unv = 0; F.kfac = 1.0; F.nbg = 1.0;
    if (FIT == 0) {
        F.f = P.fits[l];
        const long long tb = P.bg_off[l] >> 1;
        F.tk = P.tbl_k + tb; F.tc = P.tbl_cum + tb;
    } else {
        const long long b0 = P.bg_off[l];
        F.uval = P.uv_val + b0; F.ucnt = P.uv_cnt + b0; F.unv = P.uv_n[l];
        F.nbg = (double)(P.bg_off[l + 1] - b0);
        if (P.fitter == 1) F.kfac = P.kde_factor[l];
    }
    return F;
}

// ---- the lncRNA's class table: slots [tab_n, need) ----------------------------------------------------------------
// Returns false when p rises somewhere (or is NaN): the run structure the ranks rely on does not hold.
template <int FIT>
__device__ __forceinline__ bool front_extend_table(const PqArgs& P, int l, long long xlo, int tab_n, int need, FrontWarpSmem& S)
{
    const int lane = lane_id();
    const unsigned le = lanemask_lt() | (1u << lane);
    const PqFit F = front_fit_of<FIT>(P, l);
    bool ok = true;
#pragma unroll 1
    for (int s0 = tab_n; s0 < need; s0 += 32) {
        const int s = s0 + lane;
        const bool valid = s < need;
        double p = 0.0;
        if (valid) { p = front_sf<FIT>((double)(xlo + s), F); S.ptab[s] = p; }
        double pprev = __shfl_up_sync(FULL_MASK, p, 1);
        if (lane == 0 && s > 0) pprev = S.ptab[s - 1];              // written by an earlier round / an earlier call
        const bool head = valid && (s == 0 || p != pprev);
        if (valid && s > 0 && !(p <= pprev)) ok = false;
        const unsigned H = __ballot_sync(FULL_MASK, head);
        if (valid) S.grp[s] = (unsigned char)((H & le) ? s0 + 31 - __clz(H & le) : S.grp[s0 - 1]);
        __syncwarp();
    }
    return __all_sync(FULL_MASK, ok);
}

// ---- p / q phase on the survivors (positions in S.sidx, scores in S.x) ------------------------------------------------
// Returns false when the alignment has to go to k_pq. On success S.kept[0..*kept_out) holds the survivor-list
// positions of the retained HSPs. `tab_n` (warp-uniform, per lncRNA) is the number of valid ptab / grp slots, -1 once
// the lncRNA's table has failed its check.
template <int FIT>
__device__ __forceinline__ bool front_pq(const PqArgs& A, double thr, int ns, int m, long long h0,
                                         long long a, int l, int& tab_n, FrontWarpSmem& S, int* kept_out)
{
    const int lane = lane_id();
    const unsigned lt = lanemask_lt();
    if (tab_n < 0) return false;
    // ---- slots ----------------------------------------------------------------------------------------------
    const long long xlo = (long long)floor(thr);
    int nover = 0, maxslot = -1;
#pragma unroll 1
    for (int k0 = 0; k0 < ns; k0 += 32) {
        const int k = k0 + lane;
        const bool valid = k < ns;
        const double x = valid ? S.x[k] : 0.0;
        const long long xi = (long long)x;
        const long long off = xi - xlo;
        const bool direct = valid && (double)xi == x && off >= 0 && off < FRONT_DIRECT;
        const bool over = valid && !direct;
        const unsigned ob = __ballot_sync(FULL_MASK, over);
        int slot = 0;
        if (direct) { slot = (int)off; maxslot = max(maxslot, slot); }
        else if (over) {
            const int o = nover + __popc(ob & lt);
            slot = FRONT_DIRECT + (o < FRONT_OVER ? o : 0);
            if (o < FRONT_OVER) S.ox[o] = x;
        }
        nover += __popc(ob);
        if (valid) S.slot[k] = (unsigned char)slot;
    }
    if (nover > FRONT_OVER) return false;
    // zero the key histogram while the slot stores settle
    reinterpret_cast<uint4*>(S.cnt)[lane] = make_uint4(0u, 0u, 0u, 0u);
    // ---- non-class scores: the slots around them ---------------------------------------------------------------
    int o_lo = 0; bool o_beyond = false, bad = false;
    __syncwarp();
    const double ox = lane < nover ? S.ox[lane] : 0.0;
    if (lane < nover) {
        const double d = floor(ox) - (double)xlo;      // slots below x (exact: both are integers of moderate size, or d is huge)
        if (!(d >= 0.0)) bad = true;                   // x below floor(thr) (or NaN): cannot have passed a sane threshold
        else if (d >= (double)(FRONT_DIRECT - 1)) { o_beyond = true; maxslot = max(maxslot, FRONT_DIRECT - 1); }
        else { o_lo = (int)d; maxslot = max(maxslot, o_lo + 1); }
    }
    if (__any_sync(FULL_MASK, bad)) return false;
    const int need = __reduce_max_sync(FULL_MASK, maxslot) + 1;
    if (need > tab_n) {
        const bool ok = front_extend_table<FIT>(A, l, xlo, tab_n, need, S);
        tab_n = ok ? need : -1;
        if (!ok) return false;
    }
    // ---- non-class scores: p, key, order inside an odd key -------------------------------------------------------
    if (nover > 0) {
        double p = 0.0; int key = 0;
        const PqFit F = front_fit_of<FIT>(A, l);
        bool have_p = false;
        if (FIT == 1 && F.fitter == 1 && nover * ((F.unv + 31) >> 5) < F.unv) {
            // KDE: a few scores against many background values — the warp shares each sum (lane-strided partial sums,
            // fixed xor tree: the same order for every score, so the results stay monotone in the score among
            // themselves; against the table's sequential sums the order can differ in the last bit, which the checks
            // below catch) instead of one lane walking all values per score
#pragma unroll 1
            for (int o = 0; o < nover; ++o) {
                const double xo = __shfl_sync(FULL_MASK, ox, o);
                double acc = 0.0;
                for (int i = lane; i < F.unv; i += 32) acc += (double)F.ucnt[i] * ndtr_dev((xo - (double)F.uval[i]) / F.kfac);
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(FULL_MASK, acc, d);
                if (lane == o) p = 1.0 - acc / F.nbg;
            }
            have_p = true;
        }
        if (lane < nover) {
            if (!have_p) p = front_sf<FIT>(ox, F);
            if (o_beyond) {
                const double pl = S.ptab[FRONT_DIRECT - 1];
                if (!(p <= pl)) bad = true;
                key = p == pl ? 2 * S.grp[FRONT_DIRECT - 1] : FRONT_KEYS - 1;
            } else {
                const double pl = S.ptab[o_lo], ph = S.ptab[o_lo + 1];
                if (!(ph <= p && p <= pl)) bad = true;
                key = p == pl ? 2 * S.grp[o_lo] : (p == ph ? 2 * S.grp[o_lo + 1] : 2 * o_lo + 1);
            }
            S.okey[lane] = (unsigned char)key; S.op[lane] = p;
        }
        if (__any_sync(FULL_MASK, bad)) return false;
        // several survivors in one odd key: earlier in (p, list order) first; o is list order
        const unsigned om = __ballot_sync(FULL_MASK, lane < nover);
        int inner = 0;
        if (lane < nover) {
            const unsigned peers = __match_any_sync(om, (key & 1) ? key : FRONT_KEYS + lane);
            bad = __popc(peers) > 1;
        }
        if (__any_sync(FULL_MASK, bad)) {
#pragma unroll 1
            for (int o2 = 0; o2 < nover; ++o2) {
                const int k2 = __shfl_sync(FULL_MASK, key, o2);
                const double p2 = __shfl_sync(FULL_MASK, p, o2);
                if ((key & 1) && k2 == key && o2 != lane && (p2 < p || (p2 == p && o2 < lane))) ++inner;
            }
        }
        if (lane < nover) S.oin[lane] = (unsigned char)inner;
    }
    __syncwarp();
    double* xs = A.surv_p + h0;
    if (!A.fdr) {                                      // pipeline.py:778-780: hsp.pval = p, nothing dropped
        for (int k = lane; k < ns; k += 32) {
            const int v = S.slot[k];
            const double p = v < FRONT_DIRECT ? S.ptab[v] : S.op[v - FRONT_DIRECT];
            xs[k] = p; A.surv_pq[h0 + k] = p; A.surv_keep[h0 + k] = 1;
            S.kept[k] = (unsigned char)k;              // everything is retained
        }
        if (lane == 0) A.n_keep[a] = ns;
        *kept_out = ns;
        return true;
    }
    // ---- key histogram; ties of an even key in list order ---------------------------------------------------------
#pragma unroll 1
    for (int k0 = 0; k0 < ns; k0 += 32) {
        const int k = k0 + lane;
        const bool valid = k < ns;
        const unsigned vm = __ballot_sync(FULL_MASK, valid);
        int key = 0, seen = 0, before = 0, v = 0; unsigned peers = 0;
        if (valid) {
            v = S.slot[k];
            key = v < FRONT_DIRECT ? 2 * S.grp[v] : S.okey[v - FRONT_DIRECT];
            peers = __match_any_sync(vm, key);
            before = __popc(peers & lt);
            seen = S.cnt[key];
        }
        __syncwarp();
        if (valid && before == 0) S.cnt[key] = (unsigned short)(seen + __popc(peers));
        if (valid) {
            S.key[k] = (unsigned char)key;
            S.aux[k] = (unsigned char)((key & 1) ? v - FRONT_DIRECT : seen + before);
        }
        __syncwarp();
    }
    // ---- survivors with a larger key (= strictly smaller p): suffix sums, eight keys per lane --------------------
    {
        const uint4 c = reinterpret_cast<const uint4*>(S.cnt)[lane];
        const int c0 = c.x & 0xffff, c1 = c.x >> 16, c2 = c.y & 0xffff, c3 = c.y >> 16;
        const int c4 = c.z & 0xffff, c5 = c.z >> 16, c6 = c.w & 0xffff, c7 = c.w >> 16;
        const int tot = c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7;
        int incl = tot;                                // inclusive suffix over the lanes
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_down_sync(FULL_MASK, incl, d); if (lane + d < 32) incl += y; }
        const int s7 = incl - tot, s6 = s7 + c7, s5 = s6 + c6, s4 = s5 + c5, s3 = s4 + c4, s2 = s3 + c3, s1 = s2 + c2, s0 = s1 + c1;
        reinterpret_cast<uint4*>(S.cnt)[lane] = make_uint4((unsigned)s0 | ((unsigned)s1 << 16), (unsigned)s2 | ((unsigned)s3 << 16),
                                                           (unsigned)s4 | ((unsigned)s5 << 16), (unsigned)s6 | ((unsigned)s7 << 16));
    }
    __syncwarp();
    // ---- rank, q, keep ------------------------------------------------------------------------------------------------
    const double md = (double)m;                       // total_hsps = len(alignment._all_HSPs), pipeline.py:770
    int kept = 0;
#pragma unroll 1
    for (int k0 = 0; k0 < ns; k0 += 32) {
        const int k = k0 + lane;
        bool is_kept = false;
        if (k < ns) {
            const int key = S.key[k], aux = S.aux[k];
            const double p = (key & 1) ? S.op[aux] : S.ptab[key >> 1];
            const int rank = 1 + S.cnt[key] + ((key & 1) ? (int)S.oin[aux] : aux);
            const double q = __ddiv_rn(__dmul_rn(md, p), (double)rank);     // multiply first, like the reference
            is_kept = q <= A.alpha;
            xs[k] = p; A.surv_pq[h0 + k] = q; A.surv_keep[h0 + k] = is_kept ? 1 : 0;
        }
        const unsigned kb = __ballot_sync(FULL_MASK, is_kept);
        if (is_kept) S.kept[kept + __popc(kb & lt)] = (unsigned char)k;
        kept += __popc(kb);
    }
    if (lane == 0) A.n_keep[a] = kept;
    *kept_out = kept;
    return true;
}

// FIT = 0: histogram fitter only (its sf is the only one compiled in: the kernel's hot code has to stay small — the
// instruction caches hold 32 KB, and 32 warps per SM in different phases of a large kernel miss them constantly: a variant
// of this kernel that also gathered coordinates and chained ran at 44 % issue utilisation with 6.9 stall cycles per
// instruction on instruction fetch); FIT = 1: KDE / empirical.
// Registers: 64 per thread (8 CTAs per SM) with next to nothing spilled (16 bytes, touched once per group) — what is carried from one alignment to the next is
// kept to a handful of 32-bit values on purpose (a build with 264 bytes of spills ran 8 % slower, one at 48 registers
// and 10 CTAs per SM 60 % slower: the L1 left beside 170 KB of shared memory does not hold the local memory of 1024 threads).
// Also measured and not kept (BASELINE config 2, ms per launch): the next alignment's header prefetched through lanes
// or through cp.async into shared memory (0.563 -> 0.581), L2::128B / evict_last hints on the stream loads (0.557),
// scans of four tiles as packed shuffle chains after the round instead of one scan per tile (0.674 against 0.647).
template <typename T, int FIT>
__global__ void __launch_bounds__(FRONT_WARPS * 32, FRONT_CTAS_PER_SM)
k_front(const __grid_constant__ FrontArgs A)
{
    __shared__ FrontWarpSmem s_all[FRONT_WARPS];
    FrontWarpSmem& S = s_all[warp_id()];
    const int lane = lane_id();
    const PqArgs& P = A.P;
    // groups of chunk_aln consecutive alignments (about one lncRNA) are handed out through a ticket counter; the next
    // ticket is drawn while the current group is processed
    int ticket = 0;
    if (lane == 0) ticket = atomicAdd(A.ticket, 1);
    ticket = __shfl_sync(FULL_MASK, ticket, 0);
    int cur_l = -1, tab_n = 0;
    double thr = 0.0;
#pragma unroll 1
    while ((long long)ticket * A.chunk_aln < P.n_aln) {
    const int a_lo = ticket * A.chunk_aln;
    if (lane == 0) ticket = atomicAdd(A.ticket, 1);           // read after the group
#pragma unroll 1
    for (int a = a_lo; a < min(a_lo + A.chunk_aln, (int)P.n_aln); ++a) {
        const int l = P.aln_lnc[a];
        const long long hbeg = P.hsp_off[a];
        const int m = (int)(P.hsp_off[a + 1] - hbeg);
        if (l != cur_l) {
            cur_l = l; tab_n = 0;
            // fit failed: the reference raised before its alignment loop — nothing passes an infinite threshold
            thr = P.lnc_status[l] != 0 ? __longlong_as_double(0x7ff0000000000000ll) : P.thr[l];
        }
        int ns;
        __syncwarp();                                  // the previous alignment is done with S
        if constexpr (sizeof(T) == 8) ns = front_filter_f64(A.score, A.n_hsp_total, hbeg, m, thr, A.surv_idx + hbeg, S);
        else {
            const long long x0 = A.exc_off[a];
            ns = front_filter_q<T>(reinterpret_cast<const T*>(A.score_q), A.n_hsp_total, hbeg, m, thr, A.exc_pos, A.exc_val,
                                   x0, (int)(A.exc_off[a + 1] - x0), A.surv_idx + hbeg, A.surv_x + hbeg, S);
        }
        if (lane == 0) A.n_surv[a] = ns;
        if (ns == 0) { if (lane == 0) P.n_keep[a] = 0; continue; }
        __syncwarp();
        // the survivors' scores are in S.x. float64: copied asynchronously while the filter placed them (or, without
        // O2A_FRONT_CPX, independent loads of the L2-resident lines all in flight together); narrow: stored by the filter.
        // surv_x in HBM is read by k_pq and k_retained_slots only, i.e. for the alignments queued below.
        double* out_x = A.surv_x + hbeg;
        {
            if constexpr (sizeof(T) == 8) {
#if !O2A_FRONT_CPX
                double xr[FRONT_XR];
                const double* al = A.score + hbeg;
#pragma unroll
                for (int j = 0; j < FRONT_XR; ++j) { const int k = 32 * j + lane; xr[j] = k < ns ? al[S.sidx[k]] : 0.0; }
#pragma unroll
                for (int j = 0; j < FRONT_XR; ++j) { const int k = 32 * j + lane; if (k < ns) S.x[k] = xr[j]; }
#endif
            }
        }
        __syncwarp();
        bool done = false;
        int kept = 0;
        if (ns <= FRONT_SCAP) done = front_pq<FIT>(P, thr, ns, m, hbeg, a, l, tab_n, S, &kept);
        if (!done) {                                   // k_pq takes it (many survivors / many non-class scores / a failed check)
            if constexpr (sizeof(T) == 8) {            // ... and reads the survivors' scores from HBM
                const double* al = A.score + hbeg;
#pragma unroll 1
                for (int k = lane; k < ns; k += 32) out_x[k] = k < FRONT_SCAP ? S.x[k] : al[A.surv_idx[hbeg + k]];
            } else {                                   // narrow: the filter wrote the scores beyond FRONT_SCAP itself
#pragma unroll 1
                for (int k = lane; k < min(ns, FRONT_SCAP); k += 32) out_x[k] = S.x[k];
            }
            if (lane == 0) {
                P.n_keep[a] = 0;
                const int slot = atomicAdd(A.pq_count, 1);
                A.pq_list[slot] = a;
            }
            continue;
        }
        // ---- the retained HSPs for the chaining: survivor slot, score and (device-resident records) coordinates ------
        if (kept == 0) continue;
        if (lane == 0 && chain_needs_cta(kept, A.qlen[a], A.slen[a])) { const int slot = atomicAdd(A.mid_count, 1); A.mid_list[slot] = a; }
        __syncwarp();
        const long long xlo = (long long)floor(thr);
#pragma unroll 1
        for (int p = lane; p < kept; p += 32) {
            const int k = S.kept[p];
            const int v = S.slot[k];
            A.ret_slot[hbeg + p] = k;
            A.ret_score[hbeg + p] = v < FRONT_DIRECT ? (double)(xlo + v) : S.ox[v - FRONT_DIRECT];
            if (A.fetch_coords) {
                const long long h = hbeg + S.sidx[k];
                int4 c;
                if (A.coords4) c = A.coords4[h];
                else { c.x = A.qstart[h]; c.y = A.qend[h]; c.z = A.sstart[h]; c.w = A.send[h]; }
                A.ret_coord[hbeg + p] = c;
            }
        }
    }
    ticket = __shfl_sync(FULL_MASK, ticket, 0);
    }
}

}  // namespace o2a
