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
// The warp streams the alignment's scores with 16-byte loads, eight in flight per lane, and compacts the
// survivors IN ORDER (ballot / popc prefixes) into the alignment's slot [hsp_off[a], hsp_off[a] + n_surv) of the
// survivor arrays and into shared memory. The p/q stage works on SCORE CLASSES like k_pq (o2a_score.cuh): sf once
// per distinct score, L = #{survivors with smaller p} per class, then one ordered walk that hands out the stable
// ranks 1 + L + #{earlier survivors with the same p} = np.argsort(kind='stable').
// Alignments with more than FRONT_SCAP survivors or more than FRONT_OVER non-class scores are only filtered here
// and queued for the CTA-wide k_pq.
#pragma once
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
constexpr int FRONT_ROWS = 8;             // 16-byte loads in flight per lane (f64: 8 x 512 B per warp)
constexpr int FRONT_PASS = 32;            // passing exceptions of a narrow-score alignment kept on chip

struct FrontArgs {
    PqArgs P;                             // geometry, fit tables, p / q / keep outputs (n_surv, surv_x are outputs here)
    const double* score;                  // float64 layout (score_bits == 0)
    const void* score_q;                  // narrow layout: int8 / int16 / int32 + exception lists
    const long long* exc_off; const int* exc_pos; const double* exc_val;
    long long n_hsp_total;
    int* n_surv; int* surv_idx; double* surv_x;
    int* pq_list; int* pq_count;          // alignments left to k_pq
};

struct FrontWarpSmem {
    double x[FRONT_SCAP];                 // survivor scores, list order
    double pe[FRONT_SLOTS];               // p of the class (narrow filter phase: values of the passing exceptions)
    double ox[FRONT_OVER];                // scores of the overflow classes
    int cnt[FRONT_SLOTS];                 // class multiplicity, later: survivors seen per tie group
    int L[FRONT_SLOTS];                   // survivors with strictly smaller p (narrow filter phase: positions of passing exceptions)
    short g[FRONT_SLOTS];                 // tie group = first class (dense order) with the same p
    short dense[FRONT_SLOTS];             // non-empty classes in slot order
    short slot[FRONT_SCAP];               // class of every survivor
};

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
__device__ __forceinline__ int front_filter_f64(const double* __restrict__ score, long long h0, int m, double t,
                                                int* __restrict__ out_idx, double* __restrict__ out_x, FrontWarpSmem& S)
{
    const int lane = lane_id();
    const unsigned lt = lanemask_lt();
    const int sh = (int)(h0 & 1);                       // stream starts at the 16-byte aligned address at / below h0
    const double* sc = score + h0 - sh;
    const int ms = m + sh;
    const double NEG_INF = __longlong_as_double(0xfff0000000000000ll);
    int base = 0;
    for (int t0 = 0; t0 < ms; t0 += FRONT_ROWS * 64) {
        double2 v[FRONT_ROWS];
#pragma unroll
        for (int r = 0; r < FRONT_ROWS; ++r) {
            const int rs = t0 + r * 64;                 // row start (warp-uniform)
            const int e = rs + lane * 2;
            if (rs + 64 <= ms) v[r] = ld_stream_f64x2(sc + e);
            else if (rs >= ms) { v[r].x = NEG_INF; v[r].y = NEG_INF; }
            else { v[r].x = e < ms ? sc[e] : NEG_INF; v[r].y = e + 1 < ms ? sc[e + 1] : NEG_INF; }
        }
        if (sh && t0 == 0 && lane == 0) v[0].x = NEG_INF;    // the neighbour's last HSP
#pragma unroll
        for (int r = 0; r < FRONT_ROWS; ++r) {
            const unsigned mx = __ballot_sync(FULL_MASK, v[r].x > t), my = __ballot_sync(FULL_MASK, v[r].y > t);
            if (mx | my) {
                int p = base + __popc(mx & lt) + __popc(my & lt);
                const int e = t0 + r * 64 + lane * 2 - sh;   // HSP position inside the alignment
                if ((mx >> lane) & 1u) { out_idx[p] = e; out_x[p] = v[r].x; if (p < FRONT_SCAP) S.x[p] = v[r].x; ++p; }
                if ((my >> lane) & 1u) { out_idx[p] = e + 1; out_x[p] = v[r].y; if (p < FRONT_SCAP) S.x[p] = v[r].y; }
                base += __popc(mx) + __popc(my);
            }
        }
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
    // exceptions that pass (value > thr): positions in S.L, values in S.pe (both free during the filter phase)
    int np = 0;
    for (int i0 = 0; i0 < ne; i0 += 32) {
        const int i = i0 + lane;
        const double v = i < ne ? exc_val[x0 + i] : 0.0;
        const bool pass = i < ne && v > t;
        const unsigned b = __ballot_sync(FULL_MASK, pass);
        const int p = np + __popc(b & lt);
        if (pass && p < FRONT_PASS) { S.L[p] = exc_pos[x0 + i]; S.pe[p] = v; }
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
                if (np <= FRONT_PASS) { for (int i = 0; i < np; ++i) if (S.L[i] == pos) { ok = true; break; } }
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
                            if (np <= FRONT_PASS) { for (int i = 0; i < np; ++i) if (S.L[i] == ep + c) { x = S.pe[i]; break; } }
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

// ---- p / q phase on the survivors in S.x[0..ns) --------------------------------------------------------------------
// Returns false when the alignment has more than FRONT_OVER non-class scores (left to k_pq).
__device__ __forceinline__ bool front_pq(const PqArgs& A, const PqFit& F, double thr, int ns, int m, long long h0,
                                         long long a, FrontWarpSmem& S)
{
    const int lane = lane_id();
    const unsigned lt = lanemask_lt();
    for (int t = lane; t < FRONT_SLOTS; t += 32) S.cnt[t] = 0;
    __syncwarp();
    // ---- classes ---------------------------------------------------------------------------------
    const long long xlo = (long long)floor(thr);
    int nover = 0, maxslot = 0;
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
    for (int e0 = 0; e0 < ncls; e0 += 32) {
        const int e = e0 + lane;
        const int slot = e < ndir ? e : FRONT_DIRECT + (e - ndir);
        const bool used = e < ncls && S.cnt[slot] > 0;
        if (used) S.pe[slot] = pq_sf(e < ndir ? (double)(xlo + slot) : S.ox[e - ndir], F);
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
        if (lane == 0) A.n_keep[a] = ns;
        return true;
    }
    // ---- per class: survivors with strictly smaller p, and the tie group ------------------------------------
    for (int d = lane; d < nd; d += 32) {
        const int slot = S.dense[d];
        const double p = S.pe[slot];
        int L = 0, g = slot;
        for (int d2 = 0; d2 < nd; ++d2) {
            const int s2 = S.dense[d2];
            const double p2 = S.pe[s2];
            L += (p2 < p) ? S.cnt[s2] : 0;
            if (p2 == p && s2 < g) g = s2;
        }
        S.L[slot] = L; S.g[slot] = (short)g;
    }
    __syncwarp();
    for (int d = lane; d < nd; d += 32) S.cnt[S.dense[d]] = 0;         // now: survivors seen per tie group
    __syncwarp();
    // ---- ordered walk: rank = 1 + L + #{earlier survivors of the tie group} ------------------------------------
    const double md = (double)m;                       // total_hsps = len(alignment._all_HSPs), pipeline.py:770
    int kept = 0;
    for (int k0 = 0; k0 < ns; k0 += 32) {
        const int k = k0 + lane;
        const bool valid = k < ns;
        const unsigned vm = __ballot_sync(FULL_MASK, valid);
        int slot = 0, g = 0, seen = 0, before = 0; unsigned peers = 0;
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
            kept += keep;
        }
    }
    kept = __reduce_add_sync(FULL_MASK, kept);
    if (lane == 0) A.n_keep[a] = kept;
    return true;
}

template <typename T>
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
        if constexpr (sizeof(T) == 8) ns = front_filter_f64(A.score, h0, m, thr, A.surv_idx + h0, A.surv_x + h0, S);
        else {
            const long long x0 = A.exc_off[a];
            ns = front_filter_q<T>(reinterpret_cast<const T*>(A.score_q), A.n_hsp_total, h0, m, thr, A.exc_pos, A.exc_val,
                                   x0, (int)(A.exc_off[a + 1] - x0), A.surv_idx + h0, A.surv_x + h0, S);
        }
        if (lane == 0) A.n_surv[a] = ns;
        if (ns == 0) { if (lane == 0) P.n_keep[a] = 0; continue; }
        bool done = false;
        if (ns <= FRONT_SCAP) {
            __syncwarp();
            PqFit F; F.fitter = P.fitter; F.tk = nullptr; F.tc = nullptr; F.uval = nullptr; F.ucnt = nullptr;
            F.unv = 0; F.kfac = 1.0; F.nbg = 1.0;
            if (P.fitter == 0) {
                F.f = P.fits[l];
                const long long tb = P.bg_off[l] >> 1;
                F.tk = P.tbl_k + tb; F.tc = P.tbl_cum + tb;
            } else {
                const long long b0 = P.bg_off[l];
                F.uval = P.uv_val + b0; F.ucnt = P.uv_cnt + b0; F.unv = P.uv_n[l];
                F.nbg = (double)(P.bg_off[l + 1] - b0);
                if (P.fitter == 1) F.kfac = P.kde_factor[l];
            }
            done = front_pq(P, F, thr, ns, m, h0, a, S);
        }
        if (!done && lane == 0) {                      // k_pq takes it (many survivors / many non-class scores)
            P.n_keep[a] = 0;
            const int slot = atomicAdd(A.pq_count, 1);
            A.pq_list[slot] = (int)a;
        }
    }
}

}  // namespace o2a
