// o2a_chain.cuh — K6: dynamic-programming chaining of the retained HSPs of one alignment.
//
// Replaces Alignment.best_chain and everything below it (alignment.py:863-927, 982-1046):
//   _split_orientations   stable sort by (orientation, qstart, sstart)              :863-873
//   _build_hsp_graph      begin/end sentinels, O(n^2) precede/distance, prune loop  :875-927
//   _find_best_chain      relax in list order, strict '>', backtrack from end       :982-1013
//   best_chain            direct wins only if strictly greater                      :1032-1046
// The O(n^2) graph build is replaced by the equivalent windowed rule (chaining
// spec): with f(i) = first j > i that i precedes, vertex i keeps exactly the edges i -> j with
// precede(i, j) and f(i) <= j < f(f(i)). Vertices are pushed in ascending order; all threads of
// the CTA relax the window of vertex i in parallel (distinct targets, no conflicts), which keeps
// the reference's "first strict improvement wins" tie-break. Path costs mix genomic coordinates
// (begin vertex at 0) and possibly fractional scores -> fp64 with the reference's operation order
// (total + ((G + E*(qd+sd)) - score)), no FMA.
#pragma once
#include "o2a_device.cuh"

namespace o2a {

constexpr int CHAIN_THREADS = 128;
constexpr int CHAIN_WARP_MAX = 32;        // retained HSPs per alignment handled by ONE warp
constexpr int CHAIN_SMALL_MAX = 192;      // retained HSPs per alignment handled by a CTA in shared memory
constexpr int CHAINBIG_THREADS = 512;

// per-orientation vertex storage: n HSPs, vertex v in 1..n is HSP slot v-1; v=0 begin, v=n+1 end
struct ChainMem {
    int* qs; int* qe; int* ss; int* se;   // [n]
    double* sc;                           // [n]
    int* id;                              // [n] survivor-slot position of the HSP (chain output)
    double* total;                        // [n+2]
    int* prev;                            // [n+2]
    int* f;                               // [n+2]
};

struct ChainGeom {
    int n; bool direct; long long qlen, slen; double G, E;
};

__device__ __forceinline__ long long v_qs(const ChainMem& M, const ChainGeom& g, int v)
{ return v == 0 ? 0ll : (v == g.n + 1 ? g.qlen : (long long)M.qs[v - 1]); }
__device__ __forceinline__ long long v_qe(const ChainMem& M, const ChainGeom& g, int v)
{ return v == 0 ? 0ll : (v == g.n + 1 ? g.qlen : (long long)M.qe[v - 1]); }
__device__ __forceinline__ long long v_ss(const ChainMem& M, const ChainGeom& g, int v)
{   // begin: direct (0,0,0,0) / reverse (0,0,slen,slen); end: direct (..,slen,slen) / reverse (..,0,0)
    if (v == 0) return g.direct ? 0ll : g.slen;
    if (v == g.n + 1) return g.direct ? g.slen : 0ll;
    return (long long)M.ss[v - 1];
}
__device__ __forceinline__ long long v_se(const ChainMem& M, const ChainGeom& g, int v)
{
    if (v == 0) return g.direct ? 0ll : g.slen;
    if (v == g.n + 1) return g.direct ? g.slen : 0ll;
    return (long long)M.se[v - 1];
}

// HSP.precede (alignment.py:137-168). Quirk kept: both sentinels have orientation None, so a
// begin->end test uses the `else` (reverse) rule even in the direct graph.
__device__ __forceinline__ bool v_prec(const ChainMem& M, const ChainGeom& g, int i, int j)
{
    bool dir_rule = g.direct && !(i == 0 && j == g.n + 1);
    bool qp = v_qe(M, g, i) < v_qs(M, g, j);
    bool sp = dir_rule ? (v_se(M, g, i) < v_ss(M, g, j)) : (v_se(M, g, i) > v_ss(M, g, j));
    return qp && sp;
}

// HSP.distance for i preceding j (alignment.py:170-217): gapopen + gapextend*(qdist+sdist) - score_j
__device__ __forceinline__ double v_weight(const ChainMem& M, const ChainGeom& g, int i, int j)
{
    long long qd = v_qs(M, g, j) - v_qe(M, g, i);
    long long sd = g.direct ? (v_ss(M, g, j) - v_se(M, g, i)) : (v_se(M, g, i) - v_ss(M, g, j));
    double pen = __dadd_rn(g.G, __dmul_rn(g.E, (double)(qd + sd)));
    double sj = (j == 0 || j == g.n + 1) ? 0.0 : M.sc[j - 1];
    return __dsub_rn(pen, sj);
}

// Best chain of one orientation (HSP slots already sorted by (qstart, sstart, list order)).
// All threads of the CTA call this. Returns score = -total[end]; chain (survivor-slot positions, in
// chain order) is written to chain_out by thread 0, its length to *s_len (shared).
__device__ double chain_orientation(const ChainMem& M, const ChainGeom& g, int* chain_out, int* s_len)
{
    const int n = g.n, N = n + 2;
    const int lane = lane_id(), w = warp_id(), nw = blockDim.x >> 5;
    // ---- f(i): first successor -------------------------------------------------------------
    for (int i = w; i < N; i += nw) {            // one warp per vertex, lanes scan 32 candidates
        int fi = N;
        if (i < N - 1) {
            long long qe_i = v_qe(M, g, i);
            // first HSP vertex j > i with qs_j > qe_i (HSP vertices are sorted by qs)
            int lo = i + 1, hi = n + 1;         // search in [i+1, n+1)
            while (lo < hi) { int mid = (lo + hi) >> 1; if ((long long)M.qs[mid - 1] > qe_i) hi = mid; else lo = mid + 1; }
            int found = -1;
            for (int j0 = lo; j0 <= n && found < 0; j0 += 32) {
                int j = j0 + lane;
                bool ok = (j <= n) && v_prec(M, g, i, j);
                unsigned b = __ballot_sync(FULL_MASK, ok);
                if (b) found = j0 + __ffs(b) - 1;
            }
            if (found >= 0) fi = found;
            else if (v_prec(M, g, i, N - 1)) fi = N - 1;
        }
        if (lane == 0) M.f[i] = fi;
    }
    for (int v = threadIdx.x; v < N; v += blockDim.x) {
        M.total[v] = v == 0 ? 0.0 : __longlong_as_double(0x7ff0000000000000ll);
        M.prev[v] = -1;
    }
    __syncthreads();
    // ---- push DP in list order ----------------------------------------------------------------
    for (int i = 0; i < N - 1; ++i) {
        const double ti = M.total[i];
        const int fi = M.f[i];
        if (fi < N && ti < __longlong_as_double(0x7ff0000000000000ll)) {   // uniform across the CTA
            const int gi = M.f[fi] < N ? M.f[fi] : N;      // f(f(i)); f(N-1) = N
            for (int j = fi + threadIdx.x; j < gi; j += blockDim.x) {
                if (v_prec(M, g, i, j)) {
                    double ns = __dadd_rn(ti, v_weight(M, g, i, j));       // HSPVertex._relax
                    if (M.total[j] > ns) { M.total[j] = ns; M.prev[j] = i; }
                }
            }
        }
        __syncthreads();
    }
    const double score = -M.total[N - 1];
    // ---- backtrack (alignment.py:1007-1013) -----------------------------------------------------
    if (threadIdx.x == 0) {
        int len = 0, cur = N - 1;
        while (M.prev[cur] != -1) { cur = M.prev[cur]; if (cur != 0) ++len; }
        // second walk fills chain_out back to front
        int pos = len; cur = N - 1;
        while (M.prev[cur] != -1) { cur = M.prev[cur]; if (cur != 0) chain_out[--pos] = M.id[cur - 1]; }
        *s_len = len;
    }
    __syncthreads();
    return score;
}

struct ChainArgs {
    long long n_aln;
    const long long* hsp_off;
    const int* aln_lnc;
    const long long* qlen; const long long* slen;
    const int* qstart; const int* qend; const int* sstart; const int* send; const double* score;
    const int4* coords4;       // optional AoS (qstart,qend,sstart,send) per HSP; replaces the four SoA arrays
    const int* n_surv; const int* n_keep;
    const int* surv_idx; const unsigned char* surv_keep;
    const double* surv_x;      // survivor scores (slot layout)
    // retained HSPs of alignment a, list order, at [hsp_off[a], hsp_off[a] + n_keep[a]) (k_gather_retained)
    int4* ret_coord; double* ret_score; int* ret_slot;
    int gapopen, gapextend;
    int best_value;
    // outputs
    int* chain_len; unsigned char* chain_orient; double* chain_score; double* orient_scores;
    int* chain_slot;           // slot layout: positions in the retained list of the chain HSPs, chain order
    double* best_key;          // per alignment: get_best_orthologs key of the chosen chain
    int* lnc_status;
    int* mid_list; int* mid_count;     // alignments with > CHAIN_WARP_MAX retained HSPs (-> k_chain_small)
    int* big_list; int* big_count;     // alignments with > CHAIN_SMALL_MAX retained HSPs (-> k_chain_big)
};

// key of get_best_orthologs on the QUERY bed12 line (pipeline.py:981-987) for the chosen chain
__device__ double best_key_of_chain(const ChainArgs& A, long long h0, const int* slots, int len, int which)
{
    long long blen = 0; int lo = INT32_MAX, hi = INT32_MIN; double wsum = 0.0;
    for (int c = 0; c < len; ++c) {
        const int4 q = A.ret_coord[h0 + slots[c]];
        int qs = q.x, qe = q.y;
        int d = qe - qs; blen += d < 0 ? -d : d;
        lo = min(lo, min(qs, qe)); hi = max(hi, max(qs, qe));
        wsum = __dadd_rn(wsum, A.ret_score[h0 + slots[c]]);   // Python sum(): 0 + s0 + s1 + ...
    }
    if (which == 0) return (double)blen;
    if (which == 1) return (double)((long long)hi - (long long)lo);
    if (which == 2) return (double)len;
    return wsum;
}

// Gather stage: one warp per alignment compacts the retained HSPs (keep flags of k_pq) in list order
// and copies their coordinates + score next to each other. This is the only place that touches the
// coordinate arrays — ~1 % of the HSPs — and, for host batches in pinned memory, the only traffic
// they ever cause over PCIe (zero-copy reads straight from the caller's arrays).
__global__ void __launch_bounds__(256)
k_gather_retained(ChainArgs A)
{
    const int lane = lane_id();
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long a = warp; a < A.n_aln; a += nwarps) {
        const int r = A.n_keep[a];
        if (r == 0) continue;
        // alignments too large for the warp path are queued here, so that k_chain_small can run beside k_chain_warp
        if (r > CHAIN_WARP_MAX && lane == 0) { int slot = atomicAdd(A.mid_count, 1); A.mid_list[slot] = (int)a; }
        const long long h0 = A.hsp_off[a];
        const int ns = A.n_surv[a];
        int got = 0;
        // up to 128 survivors per round; each dependent level (keep flags -> survivor index -> coordinates / score)
        // is issued for all four rows before the next level uses it
        for (int k0 = 0; k0 < ns && got < r; k0 += 128) {
            bool keep[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int k = k0 + u * 32 + lane;
                keep[u] = (k < ns) && A.surv_keep[h0 + k];
            }
            long long pos[4], h[4];
            double x[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int k = k0 + u * 32 + lane;
                const unsigned b = __ballot_sync(FULL_MASK, keep[u]);
                pos[u] = h0 + got + __popc(b & lanemask_lt());
                got += __popc(b);
                h[u] = keep[u] ? h0 + A.surv_idx[h0 + k] : 0;
                x[u] = keep[u] ? A.surv_x[h0 + k] : 0.0;
            }
            int4 c[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (keep[u]) {
                    if (A.coords4) c[u] = A.coords4[h[u]];
                    else { c[u].x = A.qstart[h[u]]; c[u].y = A.qend[h[u]]; c[u].z = A.sstart[h[u]]; c[u].w = A.send[h[u]]; }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (keep[u]) { A.ret_coord[pos[u]] = c[u]; A.ret_score[pos[u]] = x[u]; A.ret_slot[pos[u]] = k0 + u * 32 + lane; }
            }
        }
    }
}

// Host-gather mode (host batches whose coordinate records stay in HOST memory): instead of reading the ~1 % retained
// records in place over PCIe (32-byte reads that compete with the upload DMA for the link's read-request slots), the
// GPU lists the global HSP indices of the retained HSPs straight into mapped pinned host memory (posted writes), the
// host gathers the 16-byte records with its own threads and sends them back as one small DMA, and k_scatter_coords
// puts them where the chaining kernels expect them. Everything else k_gather_retained does (score, survivor slot,
// queueing the alignments beyond the warp path) happens here.
__global__ void __launch_bounds__(256)
k_list_retained(ChainArgs A, long long* __restrict__ ret_base, long long* __restrict__ cursor,
                long long* __restrict__ host_list, long long cap)
{
    const int lane = lane_id();
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long a = warp; a < A.n_aln; a += nwarps) {
        const int r = A.n_keep[a];
        if (r == 0) continue;
        if (r > CHAIN_WARP_MAX && lane == 0) { int slot = atomicAdd(A.mid_count, 1); A.mid_list[slot] = (int)a; }
        long long base = 0;
        if (lane == 0) { base = (long long)atomicAdd((unsigned long long*)cursor, (unsigned long long)r); ret_base[a] = base; }
        base = __shfl_sync(FULL_MASK, base, 0);
        const long long h0 = A.hsp_off[a];
        const int ns = A.n_surv[a];
        int got = 0;
        for (int k0 = 0; k0 < ns && got < r; k0 += 32) {
            const int k = k0 + lane;
            const bool keep = (k < ns) && A.surv_keep[h0 + k];
            const unsigned b = __ballot_sync(FULL_MASK, keep);
            if (keep) {
                const int pos = got + __popc(b & lanemask_lt());
                A.ret_score[h0 + pos] = A.surv_x[h0 + k]; A.ret_slot[h0 + pos] = k;
                if (base + pos < cap) host_list[base + pos] = h0 + A.surv_idx[h0 + k];
            }
            got += __popc(b);
        }
    }
}

__global__ void __launch_bounds__(256)
k_scatter_coords(ChainArgs A, const long long* __restrict__ ret_base, const int4* __restrict__ stage)
{
    const int lane = lane_id();
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long a = warp; a < A.n_aln; a += nwarps) {
        const int r = A.n_keep[a];
        if (r == 0) continue;
        const long long h0 = A.hsp_off[a], base = ret_base[a];
        for (int i = lane; i < r; i += 32) A.ret_coord[h0 + i] = stage[base + i];
    }
}

// ret_coord of the retained HSPs from device-resident AoS coordinate records (ret_slot already filled by
// k_list_retained): the dense-chunk fallback of the host-gather mode.
__global__ void __launch_bounds__(256)
k_fill_coords(ChainArgs A)
{
    const int lane = lane_id();
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long a = warp; a < A.n_aln; a += nwarps) {
        const int r = A.n_keep[a];
        const long long h0 = A.hsp_off[a];
        for (int i = lane; i < r; i += 32) A.ret_coord[h0 + i] = A.coords4[h0 + A.surv_idx[h0 + A.ret_slot[h0 + i]]];
    }
}

// Warp path: one warp per alignment with at most CHAIN_WARP_MAX retained HSPs (the common case:
// ~10-20 retained of ~130 survivors). Everything lives in the warp's slice of shared memory; the
// begin/end sentinels are stored as ordinary vertices (64-bit coordinates). The two orientations are
// independent graphs (alignment.py:1015-1030) and are chained AT THE SAME TIME: their vertices sit in
// consecutive index ranges [0, N0) and [N0, N0 + N1), lanes 0-15 push the window of the direct
// graph's vertex `step` while lanes 16-31 push the reverse graph's, so the serial part of the DP is
// max(N0, N1) steps instead of N0 + N1.
constexpr int CHAINW_WARPS = 4;

__global__ void __launch_bounds__(CHAINW_WARPS * 32)
k_chain_warp(ChainArgs A)
{
    constexpr int C = CHAIN_WARP_MAX, V = CHAIN_WARP_MAX + 4;
    __shared__ long long s_qs[CHAINW_WARPS][V], s_qe[CHAINW_WARPS][V], s_ss[CHAINW_WARPS][V], s_se[CHAINW_WARPS][V];
    __shared__ double s_sc[CHAINW_WARPS][V], s_total[CHAINW_WARPS][V];
    __shared__ int s_id[CHAINW_WARPS][V], s_prev[CHAINW_WARPS][V], s_f[CHAINW_WARPS][V];
    __shared__ int s_chain[CHAINW_WARPS][2][C];
    const int lane = lane_id(), w = warp_id();
    const int half = lane >> 4, hl = lane & 15;          // half 0: direct graph, half 1: reverse graph
    const double NEG_INF = __longlong_as_double(0xfff0000000000000ll);
    const double POS_INF = __longlong_as_double(0x7ff0000000000000ll);
    const long long nwarps_total = (long long)gridDim.x * CHAINW_WARPS;
    long long* Qs = s_qs[w]; long long* Qe = s_qe[w]; long long* Ss = s_ss[w]; long long* Se = s_se[w];
    double* Sc = s_sc[w]; double* Tot = s_total[w]; int* Id = s_id[w]; int* Prev = s_prev[w]; int* F = s_f[w];

    for (long long a = (long long)blockIdx.x * CHAINW_WARPS + w; a < A.n_aln; a += nwarps_total) {
        const int r = A.n_keep[a];
        const int l = A.aln_lnc[a];
        if (r > C) continue;                              // queued for k_chain_small by k_gather_retained
        if (r == 0 || A.lnc_status[l] >= 2) {
            if (lane == 0) {
                A.chain_len[a] = 0; A.chain_orient[a] = 1; A.chain_score[a] = NEG_INF;
                A.orient_scores[2 * a] = A.orient_scores[2 * a + 1] = NEG_INF; A.best_key[a] = 0.0;
            }
            continue;
        }
        const long long h0 = A.hsp_off[a];
        // lane p owns the p-th retained HSP (list order)
        int qs = 0, qe = 0, ss = 0, se = 0; double sc = 0.0; bool dir = false; bool bad = false;
        if (lane < r) {
            const int4 c = A.ret_coord[h0 + lane];
            qs = c.x; qe = c.y; ss = c.z; se = c.w; sc = A.ret_score[h0 + lane];
            dir = (se - ss > 0);                         // alignment.py:107-110
            bad = !(qe - qs > 0);                        // orientation None
        }
        if (__any_sync(FULL_MASK, bad)) {                // reference raises in _split_orientations
            if (lane == 0) {
                atomicMax(&A.lnc_status[l], 1);
                A.chain_len[a] = 0; A.chain_orient[a] = 1; A.chain_score[a] = NEG_INF;
                A.orient_scores[2 * a] = A.orient_scores[2 * a + 1] = NEG_INF; A.best_key[a] = 0.0;
            }
            continue;
        }
        const long long qlen = A.qlen[a], slen = A.slen[a];
        const double G = (double)A.gapopen, E = (double)A.gapextend;
        const unsigned mem_d = __ballot_sync(FULL_MASK, lane < r && dir);
        const unsigned mem_r = __ballot_sync(FULL_MASK, lane < r && !dir);
        const int n0 = __popc(mem_d), n1 = __popc(mem_r);
        const int N0 = n0 ? n0 + 2 : 0, N1 = n1 ? n1 + 2 : 0;     // vertices per graph incl. sentinels
        // stable rank by (qstart, sstart, list order) among the HSPs of the own orientation
        int rank = 0;
        for (int q = 0; q < r; ++q) {
            const int oqs = __shfl_sync(FULL_MASK, qs, q), oss = __shfl_sync(FULL_MASK, ss, q);
            const bool same = (((mem_d >> q) & 1u) != 0) == dir;
            rank += same && ((oqs < qs) || (oqs == qs && (oss < ss || (oss == ss && q < lane))));
        }
        __syncwarp();
        if (lane < r) {
            const int v = (dir ? 0 : N0) + rank + 1;
            Qs[v] = qs; Qe[v] = qe; Ss[v] = ss; Se[v] = se; Sc[v] = sc; Id[v] = lane;
        }
        if (hl == 0) {           // sentinels (alignment.py:893-902): lane 0 for the direct graph, lane 16 for the reverse one
            const int Nn = half ? N1 : N0, base = half ? N0 : 0;
            if (Nn) {
                const bool d = half == 0;
                Qs[base] = 0; Qe[base] = 0; Ss[base] = d ? 0 : slen; Se[base] = Ss[base]; Sc[base] = 0.0;
                const int e = base + Nn - 1;
                Qs[e] = qlen; Qe[e] = qlen; Ss[e] = d ? slen : 0; Se[e] = Ss[e]; Sc[e] = 0.0;
            }
        }
        __syncwarp();
        // HSP.precede (alignment.py:137-168) inside graph [base, end); the begin-end pair uses the reverse rule (both None)
#define O2A_PREC(isdir, base, end, i, j) ((Qe[i] < Qs[j]) && (((isdir) && !((i) == (base) && (j) == (end) - 1)) ? (Se[i] < Ss[j]) : (Se[i] > Ss[j])))
        // f(i): every vertex scans forward inside its own graph
        for (int v = lane; v < N0 + N1; v += 32) {
            const bool isd = v < N0;
            const int base = isd ? 0 : N0, end = isd ? N0 : N0 + N1;
            int fi = end;
            for (int j = v + 1; j < end; ++j) if (O2A_PREC(isd, base, end, v, j)) { fi = j; break; }
            F[v] = fi;
            Tot[v] = v == base ? 0.0 : POS_INF;
            Prev[v] = -1;
        }
        __syncwarp();
        {
            const bool isd = half == 0;
            const int Nn = isd ? N0 : N1, base = isd ? 0 : N0, end = base + Nn;
            const int steps = max(N0, N1) - 1;
            for (int step = 0; step < steps; ++step) {
                if (step < Nn - 1) {
                    const int i = base + step;
                    const double ti = Tot[i];
                    const int fi = F[i];
                    if (fi < end && ti < POS_INF) {
                        const int gi = F[fi];                    // f(f(i)), `end` if none
                        for (int j = fi + hl; j < gi; j += 16) {
                            if (O2A_PREC(isd, base, end, i, j)) {
                                // HSP.distance (alignment.py:170-217) + HSPVertex._relax (:388-411)
                                const long long qd = Qs[j] - Qe[i];
                                const long long sd = isd ? (Ss[j] - Se[i]) : (Se[i] - Ss[j]);
                                const double wgt = __dsub_rn(__dadd_rn(G, __dmul_rn(E, (double)(qd + sd))), Sc[j]);
                                const double nsc = __dadd_rn(ti, wgt);
                                if (Tot[j] > nsc) { Tot[j] = nsc; Prev[j] = i; }
                            }
                        }
                    }
                }
                __syncwarp();
            }
        }
#undef O2A_PREC
        // per graph (half-warp): score, backtrack, key of get_best_orthologs
        const bool isd = half == 0;
        const int Nn = isd ? N0 : N1, base = isd ? 0 : N0, end = base + Nn;
        const double sco_h = Nn ? -Tot[end - 1] : NEG_INF;
        int ln = 0;
        if (hl == 0 && Nn) {     // backtrack (alignment.py:1007-1013): reversed chain into F[base..]
            int cur = end - 1;
            while (Prev[cur] != -1) { cur = Prev[cur]; if (cur != base) F[base + ln++] = cur; }
        }
        ln = __shfl_sync(FULL_MASK, ln, half << 4);
        __syncwarp();
        double key = 0.0;
        {
            long long blen = 0, lo = INT64_MAX, hi = INT64_MIN;
            for (int c = hl; c < ln; c += 16) {
                const int v = F[base + ln - 1 - c];
                s_chain[w][half][c] = Id[v];
                const long long d = Qe[v] - Qs[v];
                blen += d < 0 ? -d : d;
                lo = min(lo, min(Qs[v], Qe[v])); hi = max(hi, max(Qs[v], Qe[v]));
            }
#pragma unroll
            for (int d = 8; d > 0; d >>= 1) {        // stays inside the half-warp
                blen += __shfl_xor_sync(FULL_MASK, blen, d);
                const long long olo = __shfl_xor_sync(FULL_MASK, lo, d), ohi = __shfl_xor_sync(FULL_MASK, hi, d);
                lo = olo < lo ? olo : lo; hi = ohi > hi ? ohi : hi;
            }
            if (A.best_value == 0) key = (double)blen;
            else if (A.best_value == 1) key = ln ? (double)(hi - lo) : 0.0;
            else if (A.best_value == 2) key = (double)ln;
            else {              // weight: Python sum() left to right (genomicranges.py:1225)
                double wsum = 0.0;
                for (int c = ln - 1; c >= 0; --c) wsum = __dadd_rn(wsum, Sc[F[base + c]]);
                key = wsum;
            }
        }
        const double sco0 = __shfl_sync(FULL_MASK, sco_h, 0), sco1 = __shfl_sync(FULL_MASK, sco_h, 16);
        const double key0 = __shfl_sync(FULL_MASK, key, 0), key1 = __shfl_sync(FULL_MASK, key, 16);
        const int len0 = __shfl_sync(FULL_MASK, ln, 0), len1 = __shfl_sync(FULL_MASK, ln, 16);
        __syncwarp();
        const int pick = (sco0 > sco1) ? 0 : 1;          // alignment.py:1043-1046
        const int lnp = pick ? len1 : len0;
        for (int c = lane; c < lnp; c += 32) A.chain_slot[h0 + c] = s_chain[w][pick][c];
        if (lane == 0) {
            A.chain_len[a] = lnp; A.chain_orient[a] = (unsigned char)pick; A.chain_score[a] = pick ? sco1 : sco0;
            A.orient_scores[2 * a] = sco0; A.orient_scores[2 * a + 1] = sco1;
            A.best_key[a] = lnp ? (pick ? key1 : key0) : 0.0;
        }
        __syncwarp();
    }
}

// Shared-memory path: alignments with at most CHAIN_SMALL_MAX retained HSPs (the common case:
// ~10-20). Larger ones are queued for k_chain_big.
__global__ void __launch_bounds__(CHAIN_THREADS)
k_chain_small(ChainArgs A)
{
    constexpr int C = CHAIN_SMALL_MAX;
    __shared__ int s_qs[C], s_qe[C], s_ss[C], s_se[C], s_id[C];
    __shared__ double s_sc[C];
    __shared__ int t_qs[C], t_qe[C], t_ss[C], t_se[C], t_id[C];     // unsorted gather
    __shared__ double t_sc[C];
    __shared__ unsigned char t_dir[C];
    __shared__ double s_total[C + 2];
    __shared__ int s_prev[C + 2], s_f[C + 2];
    __shared__ int s_chain[2][C];
    __shared__ int s_len[2];
    __shared__ int s_bad;
    const int tid = threadIdx.x;

    const int nmid = *A.mid_count;
    for (int bi = blockIdx.x; bi < nmid; bi += gridDim.x) {
        const long long a = A.mid_list[bi];
        const long long h0 = A.hsp_off[a];
        const int r = A.n_keep[a];
        const int l = A.aln_lnc[a];
        if (r > C) {
            if (tid == 0) { int slot = atomicAdd(A.big_count, 1); A.big_list[slot] = (int)a; }
            continue;
        }
        if (tid == 0) s_bad = 0;
        // ---- the retained HSPs, list order (k_gather_retained) --------------------------------------
        __syncthreads();
        for (int p = tid; p < r; p += CHAIN_THREADS) {
            const int4 c = A.ret_coord[h0 + p];
            t_qs[p] = c.x; t_qe[p] = c.y; t_ss[p] = c.z; t_se[p] = c.w; t_sc[p] = A.ret_score[h0 + p]; t_id[p] = p;
            t_dir[p] = (c.w - c.z > 0) ? 1 : 0;                    // alignment.py:107-110
            if (!(c.y - c.x > 0)) s_bad = 1;                        // orientation None
        }
        __syncthreads();
        if (s_bad) {                                   // reference raises in _split_orientations
            if (tid == 0) {
                atomicMax(&A.lnc_status[l], 1);
                A.chain_len[a] = 0; A.chain_orient[a] = 1;
                A.chain_score[a] = __longlong_as_double(0xfff0000000000000ll);
                A.orient_scores[2 * a] = A.orient_scores[2 * a + 1] = __longlong_as_double(0xfff0000000000000ll);
                A.best_key[a] = 0.0;
            }
            __syncthreads();
            continue;
        }
        double sco[2];
        for (int o = 0; o < 2; ++o) {                  // o = 0 direct, 1 reverse
            // stable rank of each HSP of this orientation by (qstart, sstart, list order)
            int n_o = 0;
            for (int p = 0; p < r; ++p) n_o += (t_dir[p] == (o == 0 ? 1 : 0));
            for (int p = tid; p < r; p += CHAIN_THREADS) {
                if (t_dir[p] != (o == 0 ? 1 : 0)) continue;
                int rank = 0;
                for (int q = 0; q < r; ++q) {
                    if (t_dir[q] != t_dir[p]) continue;
                    bool less = (t_qs[q] < t_qs[p]) || (t_qs[q] == t_qs[p] && (t_ss[q] < t_ss[p] || (t_ss[q] == t_ss[p] && q < p)));
                    rank += less;
                }
                s_qs[rank] = t_qs[p]; s_qe[rank] = t_qe[p]; s_ss[rank] = t_ss[p]; s_se[rank] = t_se[p];
                s_sc[rank] = t_sc[p]; s_id[rank] = t_id[p];
            }
            __syncthreads();
            if (n_o == 0) {
                sco[o] = __longlong_as_double(0xfff0000000000000ll);
                if (tid == 0) s_len[o] = 0;
                __syncthreads();
                continue;
            }
            ChainMem M{s_qs, s_qe, s_ss, s_se, s_sc, s_id, s_total, s_prev, s_f};
            ChainGeom g{n_o, o == 0, A.qlen[a], A.slen[a], (double)A.gapopen, (double)A.gapextend};
            sco[o] = chain_orientation(M, g, s_chain[o], &s_len[o]);
            __syncthreads();
        }
        const int pick = (sco[0] > sco[1]) ? 0 : 1;    // alignment.py:1043-1046
        const int len = s_len[pick];
        for (int c = tid; c < len; c += CHAIN_THREADS) A.chain_slot[h0 + c] = s_chain[pick][c];
        __syncthreads();
        if (tid == 0) {
            A.chain_len[a] = len; A.chain_orient[a] = (unsigned char)pick; A.chain_score[a] = sco[pick];
            A.orient_scores[2 * a] = sco[0]; A.orient_scores[2 * a + 1] = sco[1];
            A.best_key[a] = len ? best_key_of_chain(A, h0, s_chain[pick], len, A.best_value) : 0.0;
        }
        __syncthreads();
    }
}

// Global-scratch path for dense alignments (BASELINE config 4: ~25k retained HSPs per orientation).
// Per-CTA scratch of capacity `cap` HSPs (cap = max alignment size).
struct ChainBigScratch {
    unsigned long long* k0; unsigned long long* k1; unsigned int* v0; unsigned int* v1;   // sort
    int* qs; int* qe; int* ss; int* se; double* sc; int* id;                              // sorted HSPs
    double* total; int* prev; int* f;                                                     // cap + 2
    int* chain;                                                                           // 2*cap
    long long cap;
};

__global__ void __launch_bounds__(CHAINBIG_THREADS)
k_chain_big(ChainArgs A, ChainBigScratch S)
{
    __shared__ int s_sort[32 + (CHAINBIG_THREADS / 32) * 16 + 1];
    __shared__ int s_scan[CHAINBIG_THREADS / 32 + 2];
    __shared__ int s_len[2];
    __shared__ int s_bad;
    const int tid = threadIdx.x;
    const long long so = (long long)blockIdx.x * S.cap, so2 = (long long)blockIdx.x * (S.cap + 2);
    unsigned long long *k0 = S.k0 + so, *k1 = S.k1 + so;
    unsigned int *v0 = S.v0 + so, *v1 = S.v1 + so;
    ChainMem M{S.qs + so, S.qe + so, S.ss + so, S.se + so, S.sc + so, S.id + so, S.total + so2, S.prev + so2, S.f + so2};
    int* chain_buf = S.chain + 2 * so;
    const int nbig = *A.big_count;
    for (int b = blockIdx.x; b < nbig; b += gridDim.x) {
        const long long a = A.big_list[b];
        const long long h0 = A.hsp_off[a];
        const int l = A.aln_lnc[a];
        if ((long long)A.n_keep[a] > S.cap) continue;      // understated max_aln_hsps: the call fails on the host side
        if (tid == 0) s_bad = 0;
        __syncthreads();
        double sco[2];
        for (int o = 0; o < 2; ++o) {
            // retained HSPs of this orientation, in list order, as (key, retained-list position) pairs
            int n_o = 0;
            const int r = A.n_keep[a];
            for (int k0i = 0; k0i < r; k0i += CHAINBIG_THREADS) {
                int k = k0i + tid;
                bool take = false; unsigned long long key = 0;
                if (k < r) {
                    const int4 c = A.ret_coord[h0 + k];
                    if (!(c.y - c.x > 0)) s_bad = 1;
                    bool dir = (c.w - c.z > 0);
                    take = (dir == (o == 0));
                    key = ((unsigned long long)((unsigned)c.x ^ 0x80000000u) << 32) | (unsigned long long)((unsigned)c.z ^ 0x80000000u);
                }
                int tot;
                int pos = cta_exclusive_scan(take ? 1 : 0, s_scan, &tot);
                if (take) { k0[n_o + pos] = key; v0[n_o + pos] = (unsigned)k; }
                n_o += tot;
                __syncthreads();
            }
            if (s_bad) break;
            if (n_o == 0) {
                sco[o] = __longlong_as_double(0xfff0000000000000ll);
                if (tid == 0) s_len[o] = 0;
                __syncthreads();
                continue;
            }
            int where = cta_radix_sort(k0, v0, k1, v1, n_o, s_sort);
            __syncthreads();
            const unsigned int* vs = where ? v1 : v0;
            for (int p = tid; p < n_o; p += CHAINBIG_THREADS) {
                int k = (int)vs[p];
                const int4 c = A.ret_coord[h0 + k];
                M.qs[p] = c.x; M.qe[p] = c.y; M.ss[p] = c.z; M.se[p] = c.w;
                M.sc[p] = A.ret_score[h0 + k]; M.id[p] = k;
            }
            __syncthreads();
            ChainGeom g{n_o, o == 0, A.qlen[a], A.slen[a], (double)A.gapopen, (double)A.gapextend};
            sco[o] = chain_orientation(M, g, chain_buf + (long long)o * S.cap, &s_len[o]);
            __syncthreads();
        }
        if (s_bad) {
            if (tid == 0) {
                atomicMax(&A.lnc_status[l], 1);
                A.chain_len[a] = 0; A.chain_orient[a] = 1;
                A.chain_score[a] = __longlong_as_double(0xfff0000000000000ll);
                A.orient_scores[2 * a] = A.orient_scores[2 * a + 1] = __longlong_as_double(0xfff0000000000000ll);
                A.best_key[a] = 0.0;
            }
            __syncthreads();
            continue;
        }
        const int pick = (sco[0] > sco[1]) ? 0 : 1;
        const int len = s_len[pick];
        const int* ch = chain_buf + (long long)pick * S.cap;
        for (int c = tid; c < len; c += CHAINBIG_THREADS) A.chain_slot[h0 + c] = ch[c];
        __syncthreads();
        if (tid == 0) {
            A.chain_len[a] = len; A.chain_orient[a] = (unsigned char)pick; A.chain_score[a] = sco[pick];
            A.orient_scores[2 * a] = sco[0]; A.orient_scores[2 * a + 1] = sco[1];
            A.best_key[a] = len ? best_key_of_chain(A, h0, ch, len, A.best_value) : 0.0;
        }
        __syncthreads();
    }
}

// K7: get_best_orthologs selection per lncRNA (pipeline.py:1004-1016): first extremal key wins.
// Also clears the chains of lncRNAs the reference would have aborted (status != 0).
__global__ void k_best(long long n_lnc, const long long* __restrict__ aln_off, const int* __restrict__ lnc_status,
                       int* __restrict__ n_surv, int* __restrict__ n_keep,
                       int* __restrict__ chain_len, double* __restrict__ chain_score, double* __restrict__ orient_scores,
                       const double* __restrict__ best_key, int best_fn, long long* __restrict__ best_aln)
{
    long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_lnc) return;
    long long best = -1; double bk = 0.0;
    const bool failed = lnc_status[l] != 0;
    for (long long a = aln_off[l]; a < aln_off[l + 1]; ++a) {
        if (failed) {   // the reference aborted this lncRNA: it contributes nothing
            n_surv[a] = 0; n_keep[a] = 0; chain_len[a] = 0; chain_score[a] = __longlong_as_double(0xfff0000000000000ll);
            orient_scores[2 * a] = orient_scores[2 * a + 1] = __longlong_as_double(0xfff0000000000000ll);
            continue;
        }
        if (chain_len[a] > 0) {
            double k = best_key[a];
            if (best < 0 || (best_fn == 0 ? k > bk : k < bk)) { best = a; bk = k; }
        }
    }
    best_aln[l] = best;
}

}  // namespace o2a
