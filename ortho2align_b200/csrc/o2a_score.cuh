// o2a_score.cuh — K4+K5 (p-values, rank, q, keep) for the alignments the fused front kernel
// (o2a_front.cuh: filter + p/q by one warp per alignment) leaves behind, plus helpers it shares.
//
// k_pq — CTA per queued alignment (more than FRONT_SCAP survivors, or many non-integral scores):
//   pvals = background.sf(scores) (pipeline.py:766); fdr: q = (m*p)/rank(p), keep q <= alpha, no cummin
//   (fitting.py:359-364, pipeline.py:769-776). Ranks use the build's tie policy: stable (ties keep HSP
//   order) — SURVEY.md §8a row 7.
// k_rank_big — survivor sets beyond shared memory: stable radix sort of the p-values in global scratch.
#pragma once
#include "o2a_device.cuh"
#include "o2a_fit.cuh"

namespace o2a {

constexpr int RANK_SMALL_MAX = 1024;                 // survivors ranked in shared memory
constexpr int TBL_SMEM_MAX = 256;

__device__ __forceinline__ double2 ld_stream_f64x2(const double* p)
{
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}

// Narrow scores: integral scores as int8 / int16 / int32 (BLAST raw scores are integers; only HSPs
// cut at a gene boundary carry fractional scores, alignment.py:330). The sentinel (most negative
// value of T) marks an HSP whose score is not representable: its exact float64 value sits in the
// alignment's exception list (exc_pos ascending HSP positions, exc_val). One 16-byte load carries 16 (int8),
// 8 (int16) or 4 (int32) scores per lane, `score > thr` is the integer test `q > floor(thr)` (the sentinel
// never passes it), and the few exceptions that DO pass (value > thr) are resolved once per alignment.
template <typename T> struct QTraits;
template <> struct QTraits<short> { static constexpr int PER = 8; static constexpr int SENT = -32768; };
template <> struct QTraits<int> { static constexpr int PER = 4; static constexpr int SENT = INT32_MIN; };
template <> struct QTraits<signed char> { static constexpr int PER = 16; static constexpr int SENT = -128; };

template <typename T>
__device__ __forceinline__ int q_at(const int4& raw, int c)
{
    if (sizeof(T) == 4) return c == 0 ? raw.x : c == 1 ? raw.y : c == 2 ? raw.z : raw.w;
    if (sizeof(T) == 1) {
        const int ws = c >> 2;
        const int word = ws == 0 ? raw.x : ws == 1 ? raw.y : ws == 2 ? raw.z : raw.w;
        return (int)(signed char)((word >> ((c & 3) * 8)) & 0xff);
    }
    const int wsel = c >> 1;
    const int word = wsel == 0 ? raw.x : wsel == 1 ? raw.y : wsel == 2 ? raw.z : raw.w;
    return (c & 1) ? (word >> 16) : (int)(short)(word & 0xffff);
}

struct PqArgs {
    long long n_aln;
    const long long* hsp_off;
    const int* aln_lnc;
    const int* lnc_status;
    const int* n_surv;
    const double* thr;
    // hist
    const LncFit* fits; const int* tbl_k; const double* tbl_cum; const long long* bg_off;
    // kde / empirical
    const int* uv_val; const int* uv_cnt; const int* uv_n; const double* kde_factor;
    int fitter; int fdr; double alpha;
    const double* surv_x;      // survivor scores (written by the filter phase of k_front)
    double* surv_p;            // out: p-values
    double* surv_pq; unsigned char* surv_keep; int* n_keep;
    const int* pq_list; const int* pq_count;   // alignments k_front queued for k_pq
    int* big_list; int* big_count;             // alignments k_pq queues for k_rank_big
};

// Survivor scores of one alignment are mostly small integers (BLAST raw scores) with few distinct
// values, and p = sf(score) depends on the score only. k_pq therefore works on SCORE CLASSES:
//   direct classes   : integer-valued scores in [floor(thr), floor(thr) + PQ_DIRECT)  (slot = offset)
//   overflow classes : anything else (cut HSPs with fractional scores, huge scores), one per HSP
// sf is evaluated once per class; for every class L = #{survivors with strictly smaller p} and the
// tie group g = first class with the same p; one warp then walks the survivors in list order and
// counts, per tie group, the survivors seen so far (match.any) -> the stable rank
//   rank_k = 1 + L[class_k] + #{j < k : p_j == p_k}
// exactly as np.argsort(kind='stable') would assign it. More than PQ_OVER overflow classes fall back
// to the plain all-pairs rank. (The common case — at most FRONT_SCAP survivors — never gets here: the
// same class algorithm runs warp-synchronously inside k_front.)

struct PqFit {
    LncFit f; const int* tk; const double* tc;
    const int* uval; const int* ucnt; int unv; double kfac, nbg; int fitter;
};

__device__ __forceinline__ double pq_sf(double x, const PqFit& F)
{
    if (F.fitter == 0) return hist_sf(x, F.f, F.tk, F.tc);
    if (F.fitter == 1) return kde_sf_compressed(x, F.uval, F.ucnt, F.unv, F.kfac, F.nbg);
    return emp_sf_compressed(x, F.uval, F.ucnt, F.unv, F.nbg);
}

template <int PQ_THREADS, int MIN_BLOCKS, int SMAX, int PQ_DIRECT, int PQ_OVER, int TBL_MAX>
__global__ void __launch_bounds__(PQ_THREADS, MIN_BLOCKS)
k_pq(PqArgs A)
{
    constexpr int PQ_SLOTS = PQ_DIRECT + PQ_OVER;
    // class tables; the all-pairs fallback (s_p) aliases them — the two are never live together
    __shared__ __align__(16) unsigned char s_raw[PQ_SLOTS * (8 + 4 + 4 + 2 + 2) + PQ_OVER * 8];
    static_assert(sizeof(s_raw) >= SMAX * sizeof(double), "fallback buffer must fit");
    double* s_p = reinterpret_cast<double*>(s_raw);                        // fallback path only
    double* s_pe = reinterpret_cast<double*>(s_raw);                       // p of the class
    double* s_ox = s_pe + PQ_SLOTS;                                        // scores of the overflow classes
    int* s_cnt = reinterpret_cast<int*>(s_ox + PQ_OVER);                   // class multiplicity, later tie counters
    int* s_L = s_cnt + PQ_SLOTS;                                           // survivors with strictly smaller p
    short* s_g = reinterpret_cast<short*>(s_L + PQ_SLOTS);                 // tie group = first slot with the same p
    short* s_dense = s_g + PQ_SLOTS;                                       // non-empty classes, slot order
    __shared__ int s_tk[TBL_MAX];
    __shared__ double s_tc[TBL_MAX];
    __shared__ short s_slot[SMAX];        // class of every survivor
    __shared__ int s_scan[PQ_THREADS / 32 + 2];
    __shared__ int s_keepcnt, s_nover, s_maxslot;
    const int tid = threadIdx.x, lane = lane_id();

    const int n_items = *A.pq_count;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const long long a = A.pq_list[it];
        const int ns = A.n_surv[a];
        if (ns == 0) { if (tid == 0) A.n_keep[a] = 0; continue; }
        const int l = A.aln_lnc[a];
        const long long h0 = A.hsp_off[a];
        const int m = (int)(A.hsp_off[a + 1] - h0);
        PqFit F; F.fitter = A.fitter; F.tk = nullptr; F.tc = nullptr; F.uval = nullptr; F.ucnt = nullptr;
        F.unv = 0; F.kfac = 1.0; F.nbg = 1.0;
        __syncthreads();                              // previous alignment is done with the shared tables
        if (A.fitter == 0) {
            F.f = A.fits[l];
            const long long tb = A.bg_off[l] >> 1;
            F.tk = A.tbl_k + tb; F.tc = A.tbl_cum + tb;
            if (F.f.tbl_n <= TBL_MAX) {
                for (int t = tid; t < F.f.tbl_n; t += PQ_THREADS) { s_tk[t] = F.tk[t]; s_tc[t] = F.tc[t]; }
                F.tk = s_tk; F.tc = s_tc;
            }
        } else {
            const long long b0 = A.bg_off[l];
            F.uval = A.uv_val + b0; F.ucnt = A.uv_cnt + b0; F.unv = A.uv_n[l];
            F.nbg = (double)(A.bg_off[l + 1] - b0);
            if (A.fitter == 1) F.kfac = A.kde_factor[l];
        }
        const double* xin = A.surv_x + h0;            // survivor scores
        double* xs = A.surv_p + h0;                   // out: p-values
        if (ns > SMAX) {
            // big survivor set: p per survivor here, ranks by radix sort in k_rank_big
            __syncthreads();
            for (int k = tid; k < ns; k += PQ_THREADS) {
                const double p = pq_sf(xin[k], F);
                xs[k] = p;
                if (!A.fdr) { A.surv_pq[h0 + k] = p; A.surv_keep[h0 + k] = 1; }
            }
            if (tid == 0) {
                if (A.fdr) { A.n_keep[a] = 0; int slot = atomicAdd(A.big_count, 1); A.big_list[slot] = (int)a; }
                else A.n_keep[a] = ns;
            }
            continue;
        }
        // ---- classes ---------------------------------------------------------------------------------
        for (int t = tid; t < PQ_SLOTS; t += PQ_THREADS) s_cnt[t] = 0;
        if (tid == 0) { s_keepcnt = 0; s_nover = 0; s_maxslot = 0; }
        __syncthreads();
        const double thr = A.thr[l];
        const long long xlo = (long long)floor(thr);
        for (int k = tid; k < ns; k += PQ_THREADS) {
            const double x = xin[k];
            const long long xi = (long long)x;
            const long long off = xi - xlo;
            int slot;
            if ((double)xi == x && off >= 0 && off < PQ_DIRECT) {
                slot = (int)off;
                atomicAdd(&s_cnt[slot], 1);
                atomicMax(&s_maxslot, slot);
            } else {
                const int o = atomicAdd(&s_nover, 1);
                slot = PQ_DIRECT + (o < PQ_OVER ? o : 0);
                if (o < PQ_OVER) { s_ox[o] = x; s_cnt[slot] = 1; }
            }
            s_slot[k] = (short)slot;
        }
        __syncthreads();
        const int nover = s_nover;
        if (nover > PQ_OVER) {
            // ---- fallback: all-pairs stable rank ---------------------------------------------------------
            for (int k = tid; k < ns; k += PQ_THREADS) {
                const double p = pq_sf(xin[k], F);
                s_p[k] = p; xs[k] = p;
                if (!A.fdr) { A.surv_pq[h0 + k] = p; A.surv_keep[h0 + k] = 1; }
            }
            __syncthreads();
            if (!A.fdr) { if (tid == 0) A.n_keep[a] = ns; continue; }
            const double md = (double)m;
            int kept = 0;
            for (int k = tid; k < ns; k += PQ_THREADS) {
                const double p = s_p[k];
                int rank = 1;
                for (int j = 0; j < k; ++j) rank += (s_p[j] <= p);
                for (int j = k + 1; j < ns; ++j) rank += (s_p[j] < p);
                const double q = __ddiv_rn(__dmul_rn(md, p), (double)rank);
                const bool keep = q <= A.alpha;
                A.surv_pq[h0 + k] = q; A.surv_keep[h0 + k] = keep ? 1 : 0;
                kept += keep;
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) kept += __shfl_xor_sync(FULL_MASK, kept, d);
            if (lane == 0 && kept) atomicAdd(&s_keepcnt, kept);
            __syncthreads();
            if (tid == 0) A.n_keep[a] = s_keepcnt;
            continue;
        }
        // ---- sf once per class -----------------------------------------------------------------------------
        const int ndir = s_maxslot + 1;
        const int ncls = ndir + nover;                 // class index e -> slot
        for (int e = tid; e < ncls; e += PQ_THREADS) {
            const int slot = e < ndir ? e : PQ_DIRECT + (e - ndir);
            if (s_cnt[slot] > 0) s_pe[slot] = pq_sf(e < ndir ? (double)(xlo + slot) : s_ox[e - ndir], F);
        }
        __syncthreads();
        if (!A.fdr) {                                  // pipeline.py:778-780: hsp.pval = p, nothing dropped
            for (int k = tid; k < ns; k += PQ_THREADS) {
                const double p = s_pe[s_slot[k]];
                xs[k] = p; A.surv_pq[h0 + k] = p; A.surv_keep[h0 + k] = 1;
            }
            if (tid == 0) A.n_keep[a] = ns;
            continue;
        }
        // ---- per class: survivors with strictly smaller p, and the tie group -----------------------------------
        // dense list of the non-empty classes (slot order) so the all-pairs loop runs on ~35 entries
        int nd = 0;
        for (int e0 = 0; e0 < ncls; e0 += PQ_THREADS) {
            const int e = e0 + tid;
            const int slot = e < ndir ? e : PQ_DIRECT + (e - ndir);
            const bool used = e < ncls && s_cnt[slot] > 0;
            int tot;
            const int pos = cta_exclusive_scan(used ? 1 : 0, s_scan, &tot);
            if (used) s_dense[nd + pos] = (short)slot;
            nd += tot;
            __syncthreads();
        }
        for (int d = tid; d < nd; d += PQ_THREADS) {
            const int slot = s_dense[d];
            const double p = s_pe[slot];
            int L = 0, g = slot;
            for (int d2 = 0; d2 < nd; ++d2) {
                const int s2 = s_dense[d2];
                const double p2 = s_pe[s2];
                L += (p2 < p) ? s_cnt[s2] : 0;
                if (p2 == p && s2 < g) g = s2;
            }
            s_L[slot] = L; s_g[slot] = (short)g;
        }
        __syncthreads();
        for (int t = tid; t < PQ_SLOTS; t += PQ_THREADS) s_cnt[t] = 0;     // now: survivors seen per tie group
        __syncthreads();
        // ---- one warp walks the survivors in list order ----------------------------------------------------------
        if (tid < 32) {
            const double md = (double)m;              // total_hsps = len(alignment._all_HSPs)
            int kept = 0;
            for (int k0 = 0; k0 < ns; k0 += 32) {
                const int k = k0 + lane;
                const bool valid = k < ns;
                const unsigned vm = __ballot_sync(FULL_MASK, valid);
                int slot = 0, g = 0, base = 0, before = 0; unsigned peers = 0;
                if (valid) {
                    slot = s_slot[k]; g = s_g[slot];
                    peers = __match_any_sync(vm, g);
                    before = __popc(peers & lanemask_lt());
                    base = s_cnt[g];
                }
                __syncwarp();
                if (valid && before == 0) s_cnt[g] = base + __popc(peers);
                __syncwarp();
                if (valid) {
                    const double p = s_pe[slot];
                    const int rank = 1 + s_L[slot] + base + before;
                    const double q = __ddiv_rn(__dmul_rn(md, p), (double)rank);
                    const bool keep = q <= A.alpha;
                    xs[k] = p; A.surv_pq[h0 + k] = q; A.surv_keep[h0 + k] = keep ? 1 : 0;
                    kept += keep;
                }
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) kept += __shfl_xor_sync(FULL_MASK, kept, d);
            if (lane == 0) A.n_keep[a] = kept;
        }
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
        if ((long long)ns > cap) continue;                  // understated max_aln_hsps: the call fails on the host side
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
