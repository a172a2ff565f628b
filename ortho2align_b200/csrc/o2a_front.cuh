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
// The warp streams the alignment's scores (float64: 32-byte loads, a 4 KB tile in flight per warp) and compacts the
// survivors IN ORDER (popc + warp scan) into the alignment's slot [hsp_off[a], hsp_off[a] + n_surv) of the
// survivor arrays and into shared memory. The p/q stage works on SCORE CLASSES like k_pq (o2a_score.cuh): sf once
// per distinct score, L = #{survivors with smaller p} per class, then one ordered walk that hands out the stable
// ranks 1 + L + #{earlier survivors with the same p} = np.argsort(kind='stable').
// Alignments with more than FRONT_SCAP survivors or more than FRONT_OVER non-class scores are only filtered here
// and queued for the CTA-wide k_pq.
#pragma once
#include "o2a_chain.cuh"
#include "o2a_device.cuh"
#include "o2a_fit.cuh"
#include "o2a_score.cuh"

namespace o2a {

constexpr int FRONT_WARPS = 4;            // warps per CTA; every warp owns its alignments
constexpr int FRONT_CTAS_PER_SM = 8;      // 32 warps per SM
constexpr int FRONT_SCAP = 192;           // survivors per alignment kept on chip
constexpr int FRONT_DIRECT = 128;         // integer score classes floor(thr) + [0, 128)
constexpr int FRONT_OVER = 32;            // other scores (cut HSPs, huge values): one class each
constexpr int FRONT_SLOTS = FRONT_DIRECT + FRONT_OVER;
constexpr int FRONT_PASS = 32;            // passing exceptions of a narrow-score alignment kept on chip

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
};

struct FrontWarpSmem {
    double x[FRONT_SCAP];                 // survivor scores, list order
    double pe[FRONT_SLOTS];               // p of the class
    double ox[FRONT_OVER];                // scores of the overflow classes (narrow filter phase: values of the passing exceptions)
    int cnt[FRONT_SLOTS];                 // class multiplicity, later: survivors seen per tie group (narrow filter phase: positions of the passing exceptions)
    int L[FRONT_SLOTS];                   // survivors with strictly smaller p    } float64 filter phase: L and g together hold
    short g[FRONT_SLOTS];                 // tie group = first class with the same p } the survivors' positions (FRONT_SCAP ints)
    short dense[FRONT_SLOTS];             // non-empty classes in slot order
    short slot[FRONT_SCAP];               // class of every survivor
};

static_assert(FRONT_SLOTS * (sizeof(int) + sizeof(short)) >= FRONT_SCAP * sizeof(int) && FRONT_PASS <= FRONT_OVER, "aliases of the filter phase");

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

// ---- filter phase, float64 scores: returns n_surv; survivors to global slots and (first FRONT_SCAP) to S.x ------
// Tile = 512 scores; lane i owns the 16 CONSECUTIVE scores [16 i, 16 i + 16) of the tile (four 32-byte loads, one
// full DRAM sector each: LDG.E.256 on sm_100), so the order-preserving position of a survivor is
//   (survivors of earlier tiles) + (exclusive warp scan of the per-lane counts) + (rank inside the lane's 16 bits):
// one popc + one 5-step shuffle scan per 512 scores instead of two ballots per 64. The few survivors (~1 per lane and
// tile) are then written in a loop over the set bits; their value is re-read (an L2 hit) rather than selected out of
// 16 registers with a run-time index.
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

__device__ __forceinline__ void front_load16(const double* __restrict__ sc, long long g0, long long n_total, int e0, double t,
                                             double (&v)[16], bool inside)
{
    // vector loads may run past the alignment's end (the next alignment's scores, masked by the caller) but not past
    // the end of the score array; `inside` (warp-uniform): the whole tile is inside the array
    if (inside) {
#pragma unroll
        for (int q = 0; q < 4; ++q) ld_stream_f64x4(sc + e0 + 4 * q, v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        return;
    }
#pragma unroll
    for (int c = 0; c < 16; ++c) v[c] = (g0 + e0 + c < n_total) ? sc[e0 + c] : t;      // the last tile of the whole array
}

__device__ __forceinline__ int front_filter_f64(const double* __restrict__ score, long long n_total, long long h0, int m,
                                                double t, int* __restrict__ out_idx, double* __restrict__ out_x,
                                                FrontWarpSmem& S)
{
    const int lane = lane_id();
    const int sh = (int)(h0 & 3);                       // stream starts at the 32-byte aligned address at / below h0
    const double* sc = score + h0 - sh;
    const long long g0 = h0 - sh;                       // global index of stream element 0
    const int ms = m + sh;
    int* s_idx = S.L;                                   // survivor positions (S.L and S.g are free until the p/q phase)
    int base = 0;
    double v[16];
    front_load16(sc, g0, n_total, lane * 16, t, v, g0 + 512 <= n_total);
#pragma unroll 1
    for (int t0 = 0; t0 < ms; t0 += 512) {
        const int e0 = t0 + lane * 16;                  // my first stream element
        unsigned mask = front_mask16(v, t);
        // the next tile's loads are in flight while this tile's survivors are placed
        if (t0 + 512 < ms) front_load16(sc, g0, n_total, e0 + 512, t, v, g0 + t0 + 1024 <= n_total);
        const int left = ms - e0;                       // my scores inside the alignment
        if (left < 16) mask &= left <= 0 ? 0u : ((1u << left) - 1u);
        if (t0 == 0 && lane == 0) mask &= ~((1u << sh) - 1u);          // the neighbour's last HSPs
        const int cnt = __popc(mask);
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(FULL_MASK, incl, d); if (lane >= d) incl += y; }
        const int total = __shfl_sync(FULL_MASK, incl, 31);
        if (total) {
            int p = base + incl - cnt;
#pragma unroll 1
            while (mask) {
                const int c = __ffs(mask) - 1;
                mask &= mask - 1;
                const int pos = e0 + c - sh;            // HSP position inside the alignment
                out_idx[p] = pos;
                if (p < FRONT_SCAP) s_idx[p] = pos;
                ++p;
            }
        }
        base += total;
    }
    // the survivors' scores: one independent (L2-resident) load each, all in flight together
    __syncwarp();
    const double* al = score + h0;
#pragma unroll 1
    for (int k = lane; k < base; k += 32) {
        const double x = al[k < FRONT_SCAP ? s_idx[k] : out_idx[k]];
        out_x[k] = x;
        if (k < FRONT_SCAP) S.x[k] = x;
    }
    return base;
}

// ---- filter phase, narrow scores (int8 / int16 / int32 + exception lists; see k_filter_q's contract) --------------
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
    // exceptions that pass (value > thr): positions in S.cnt, values in S.ox (both free during the filter phase)
    int np = 0;
    for (int i0 = 0; i0 < ne; i0 += 32) {
        const int i = i0 + lane;
        const double v = i < ne ? exc_val[x0 + i] : 0.0;
        const bool pass = i < ne && v > t;
        const unsigned b = __ballot_sync(FULL_MASK, pass);
        const int p = np + __popc(b & lt);
        if (pass && p < FRONT_PASS) { S.cnt[p] = exc_pos[x0 + i]; S.ox[p] = v; }
        np += __popc(b);
    }
    __syncwarp();
    const int sh = (int)(h0 & (PER - 1));
    const T* sc = score_q + h0 - sh;
    const int ms = m + sh;
    const long long g0 = h0 - sh;                        // global index of stream element 0
    int base = 0;
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
#pragma unroll
            for (int c = 0; c < PER; ++c) {
                const int q = q_at<T>(raw[r], c);
                msk |= (q > tq ? 1u : 0u) << c;
                sent |= (q == SENT ? 1u : 0u) << c;
            }
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
                if (np <= FRONT_PASS) { for (int i = 0; i < np; ++i) if (S.cnt[i] == pos) { ok = true; break; } }
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
            if (msk) {
                int p = base + incl - cnt;
                const int ep = e - sh;                   // HSP position of my first score of this row
#pragma unroll
                for (int c = 0; c < PER; ++c) {
                    if ((msk >> c) & 1u) {
                        double x = (double)q_at<T>(raw[r], c);
                        if ((sent_pass >> c) & 1u) {     // a passing exception: its exact float64 value
                            if (np <= FRONT_PASS) { for (int i = 0; i < np; ++i) if (S.cnt[i] == ep + c) { x = S.ox[i]; break; } }
                            else {
                                int lo = 0, hi = ne;
                                while (lo < hi) { const int mid = (lo + hi) >> 1; if (exc_pos[x0 + mid] < ep + c) lo = mid + 1; else hi = mid; }
                                x = exc_val[x0 + lo];
                            }
                        }
                        out_idx[p] = ep + c; out_x[p] = x;
                        if (p < FRONT_SCAP) S.x[p] = x;
                        ++p;
                    }
                }
            }
            base += rowtot;
        }
    }
    return base;
}

template <int FIT>
__device__ __forceinline__ double front_sf(double x, const PqFit& F)
{
    if (FIT == 0) return hist_sf(x, F.f, F.tk, F.tc);
    if (F.fitter == 1) return kde_sf_compressed(x, F.uval, F.ucnt, F.unv, F.kfac, F.nbg);
    return emp_sf_compressed(x, F.uval, F.ucnt, F.unv, F.nbg);
}

// ---- p / q phase on the survivors in S.x[0..ns) --------------------------------------------------------------------
// Returns false when the alignment has more than FRONT_OVER non-class scores (left to k_pq).
// On success S.slot[0..kept) holds the survivor-list positions of the retained HSPs and *kept_out their number.
template <int FIT>
__device__ __forceinline__ bool front_pq(const PqArgs& A, const PqFit& F, double thr, int ns, int m, long long h0,
                                         long long a, FrontWarpSmem& S, int* kept_out)
{
    const int lane = lane_id();
    const unsigned lt = lanemask_lt();
    for (int t = lane; t < FRONT_SLOTS; t += 32) S.cnt[t] = 0;
    __syncwarp();
    // ---- classes ---------------------------------------------------------------------------------
    const long long xlo = (long long)floor(thr);
    int nover = 0, maxslot = 0;
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
        if (direct) { slot = (int)off; atomicAdd(&S.cnt[slot], 1); maxslot = max(maxslot, slot); }
        else if (over) {
            const int o = nover + __popc(ob & lt);
            slot = FRONT_DIRECT + (o < FRONT_OVER ? o : 0);
            if (o < FRONT_OVER) { S.ox[o] = x; S.cnt[slot] = 1; }
        }
        nover += __popc(ob);
        if (valid) S.slot[k] = (short)slot;
    }
    if (nover > FRONT_OVER) return false;
    maxslot = __reduce_max_sync(FULL_MASK, maxslot);
    __syncwarp();
    // ---- sf once per class; dense list of the non-empty classes (slot order) ---------------------------
    const int ndir = maxslot + 1;
    const int ncls = ndir + nover;
    int nd = 0;
#pragma unroll 1
    for (int e0 = 0; e0 < ncls; e0 += 32) {
        const int e = e0 + lane;
        const int slot = e < ndir ? e : FRONT_DIRECT + (e - ndir);
        const bool used = e < ncls && S.cnt[slot] > 0;
        if (used) S.pe[slot] = front_sf<FIT>(e < ndir ? (double)(xlo + slot) : S.ox[e - ndir], F);
        const unsigned ub = __ballot_sync(FULL_MASK, used);
        if (used) S.dense[nd + __popc(ub & lt)] = (short)slot;
        nd += __popc(ub);
    }
    __syncwarp();
    double* xs = A.surv_p + h0;
    if (!A.fdr) {                                      // pipeline.py:778-780: hsp.pval = p, nothing dropped
        for (int k = lane; k < ns; k += 32) {
            const double p = S.pe[S.slot[k]];
            xs[k] = p; A.surv_pq[h0 + k] = p; A.surv_keep[h0 + k] = 1;
        }
        __syncwarp();
        for (int k = lane; k < ns; k += 32) S.slot[k] = (short)k;       // everything is retained
        if (lane == 0) A.n_keep[a] = ns;
        *kept_out = ns;
        return true;
    }
    // ---- per class: survivors with strictly smaller p (L), and the tie group (g) ------------------------------------
    // sf is non-increasing in the score, so along the dense list (ascending integer scores) p normally never rises:
    // tie groups are runs, L = survivors after the run. That takes two ballots and two scans for up to 64 classes.
    // Anything else (overflow classes, more than 64 classes, a p that rises by an ulp) takes the all-pairs form.
    bool fast = nover == 0 && nd <= 64;
    int slot_u[2], cn_u[2]; double p_u[2]; bool head_u[2];
    if (fast) {
        bool mono = true;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int d = lane + 32 * u;
            const bool valid = d < nd;
            slot_u[u] = valid ? S.dense[d] : 0;
            cn_u[u] = valid ? S.cnt[slot_u[u]] : 0;
            p_u[u] = valid ? S.pe[slot_u[u]] : 0.0;
            const double pprev = (valid && d > 0) ? S.pe[S.dense[d - 1]] : 0.0;
            head_u[u] = valid && (d == 0 || p_u[u] != pprev);
            if (valid && d > 0 && p_u[u] > pprev) mono = false;
        }
        fast = __all_sync(FULL_MASK, mono);
    }
    if (fast) {
        const unsigned H0 = __ballot_sync(FULL_MASK, head_u[0]), H1 = __ballot_sync(FULL_MASK, head_u[1]);
        int incl0 = cn_u[0], incl1 = cn_u[1];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y0 = __shfl_up_sync(FULL_MASK, incl0, d), y1 = __shfl_up_sync(FULL_MASK, incl1, d);
            if (lane >= d) { incl0 += y0; incl1 += y1; }
        }
        incl1 += __shfl_sync(FULL_MASK, incl0, 31);
        const int total = __shfl_sync(FULL_MASK, incl1, 31);
        const int cprev0 = incl0 - cn_u[0], cprev1 = incl1 - cn_u[1];      // survivors before this class
        const unsigned le = lt | (1u << lane), gt = ~le;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const unsigned Hu = u ? H1 : H0;
            // head of my run (at or below me) and the next head (above me), as dense indices
            int hidx, nidx;
            if (Hu & le) hidx = 32 * u + 31 - __clz(Hu & le); else hidx = 31 - __clz(H0);
            if (Hu & gt) nidx = 32 * u + __ffs(Hu & gt) - 1;
            else if (u == 0 && H1) nidx = 32 + __ffs(H1) - 1;
            else nidx = -1;
            const int c0 = __shfl_sync(FULL_MASK, cprev0, nidx & 31), c1 = __shfl_sync(FULL_MASK, cprev1, nidx & 31);
            const int before_next = nidx < 0 ? total : (nidx < 32 ? c0 : c1);
            if (lane + 32 * u < nd) { S.L[slot_u[u]] = total - before_next; S.g[slot_u[u]] = S.dense[hidx]; }
        }
    } else {
#pragma unroll 1
        for (int d = lane; d < nd; d += 32) {
            const int slot = S.dense[d];
            const double p = S.pe[slot];
            int L = 0, g = slot;
#pragma unroll 1
            for (int d2 = 0; d2 < nd; ++d2) {
                const int s2 = S.dense[d2];
                const double p2 = S.pe[s2];
                L += (p2 < p) ? S.cnt[s2] : 0;
                if (p2 == p && s2 < g) g = s2;
            }
            S.L[slot] = L; S.g[slot] = (short)g;
        }
    }
    __syncwarp();
    for (int d = lane; d < nd; d += 32) S.cnt[S.dense[d]] = 0;         // now: survivors seen per tie group
    __syncwarp();
    // ---- ordered walk: rank = 1 + L + #{earlier survivors of the tie group} ------------------------------------
    const double md = (double)m;                       // total_hsps = len(alignment._all_HSPs), pipeline.py:770
    int kept = 0;
#pragma unroll 1
    for (int k0 = 0; k0 < ns; k0 += 32) {
        const int k = k0 + lane;
        const bool valid = k < ns;
        const unsigned vm = __ballot_sync(FULL_MASK, valid);
        int slot = 0, g = 0, seen = 0, before = 0; unsigned peers = 0;
        bool is_kept = false;
        if (valid) {
            slot = S.slot[k]; g = S.g[slot];
            peers = __match_any_sync(vm, g);
            before = __popc(peers & lt);
            seen = S.cnt[g];
        }
        __syncwarp();
        if (valid && before == 0) S.cnt[g] = seen + __popc(peers);
        __syncwarp();
        if (valid) {
            const double p = S.pe[slot];
            const int rank = 1 + S.L[slot] + seen + before;
            const double q = __ddiv_rn(__dmul_rn(md, p), (double)rank);     // multiply first, like the reference
            const bool keep = q <= A.alpha;
            xs[k] = p; A.surv_pq[h0 + k] = q; A.surv_keep[h0 + k] = keep ? 1 : 0;
            is_kept = keep;
        }
        // the retained HSPs in list order, compacted in place: the slot entries at or below k have been read
        const unsigned kb = __ballot_sync(FULL_MASK, is_kept);
        if (is_kept) S.slot[kept + __popc(kb & lt)] = (short)k;
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
template <typename T, int FIT>
__global__ void __launch_bounds__(FRONT_WARPS * 32, FRONT_CTAS_PER_SM)
k_front(FrontArgs A)
{
    __shared__ FrontWarpSmem s_all[FRONT_WARPS];
    FrontWarpSmem& S = s_all[warp_id()];
    const int lane = lane_id();
    const PqArgs& P = A.P;
    const long long nwarps = (long long)gridDim.x * FRONT_WARPS;
    for (long long a = (long long)blockIdx.x * FRONT_WARPS + warp_id(); a < P.n_aln; a += nwarps) {
        const int l = P.aln_lnc[a];
        const long long h0 = P.hsp_off[a];
        const int m = (int)(P.hsp_off[a + 1] - h0);
        if (P.lnc_status[l] != 0) {                    // fit failed: the reference raised before its alignment loop
            if (lane == 0) { A.n_surv[a] = 0; P.n_keep[a] = 0; }
            continue;
        }
        const double thr = P.thr[l];
        int ns;
        __syncwarp();                                  // the previous alignment is done with S
        if constexpr (sizeof(T) == 8) ns = front_filter_f64(A.score, A.n_hsp_total, h0, m, thr, A.surv_idx + h0, A.surv_x + h0, S);
        else {
            const long long x0 = A.exc_off[a];
            ns = front_filter_q<T>(reinterpret_cast<const T*>(A.score_q), A.n_hsp_total, h0, m, thr, A.exc_pos, A.exc_val,
                                   x0, (int)(A.exc_off[a + 1] - x0), A.surv_idx + h0, A.surv_x + h0, S);
        }
        if (lane == 0) A.n_surv[a] = ns;
        if (ns == 0) { if (lane == 0) P.n_keep[a] = 0; continue; }
        bool done = false;
        int kept = 0;
        if (ns <= FRONT_SCAP) {
            __syncwarp();
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
            done = front_pq<FIT>(P, F, thr, ns, m, h0, a, S, &kept);
        }
        if (!done) {                                   // k_pq takes it (many survivors / many non-class scores)
            if (lane == 0) {
                P.n_keep[a] = 0;
                const int slot = atomicAdd(A.pq_count, 1);
                A.pq_list[slot] = (int)a;
            }
            continue;
        }
        // ---- the retained HSPs for the chaining: survivor slot, score and (device-resident records) coordinates ------
        if (kept == 0) continue;
        if (lane == 0 && chain_needs_cta(kept, A.qlen[a], A.slen[a])) { const int slot = atomicAdd(A.mid_count, 1); A.mid_list[slot] = (int)a; }
        __syncwarp();
#pragma unroll 1
        for (int p = lane; p < kept; p += 32) {
            const int k = S.slot[p];
            A.ret_slot[h0 + p] = k; A.ret_score[h0 + p] = S.x[k];
            if (A.fetch_coords) {
                const long long h = h0 + A.surv_idx[h0 + k];
                int4 c;
                if (A.coords4) c = A.coords4[h];
                else { c.x = A.qstart[h]; c.y = A.qend[h]; c.z = A.sstart[h]; c.w = A.send[h]; }
                A.ret_coord[h0 + p] = c;
            }
        }
    }
}

}  // namespace o2a
