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

constexpr int CHAIN_THREADS = 32;         // k_chain_small: one warp per queued alignment (31..192 retained HSPs)
constexpr int CHAIN_WARP_MAX = 30;        // retained HSPs per alignment handled by ONE warp (one graph + 2 sentinels <= 32 lanes)
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
    // k_retained_slots / k_fetch_coords: the alignments of work_list[0..*work_count) only (nullptr: all)
    const int* work_list; const int* work_count;
    int* mid_list; int* mid_count;     // alignments with > CHAIN_WARP_MAX retained HSPs (-> k_chain_small)
    int* big_list; int* big_count;     // alignments with > CHAIN_SMALL_MAX retained HSPs (-> k_chain_big)
};

// Neumaier-compensated running sum, the float branch of CPython >= 3.12's sum() (Python/bltinmodule.c)
__device__ __forceinline__ void neumaier_add(double& f_result, double& c, double x)
{
    const double t = __dadd_rn(f_result, x);
    if (fabs(f_result) >= fabs(x)) c = __dadd_rn(c, __dadd_rn(__dsub_rn(f_result, t), x));
    else c = __dadd_rn(c, __dadd_rn(__dsub_rn(x, t), f_result));
    f_result = t;
}

// sum(scores) as the pinned interpreter computes it for the BED score column (genomicranges.py:1225): ints exactly while
// every item is an int, Neumaier-compensated floats from the first float on. The columnar batch does not carry the JSON
// number type: an integral score counts as an int (the two sums are equal below 2^53).
struct PySum {
    long long isum = 0; bool in_int = true; double fr = 0.0, cc = 0.0;
    __device__ __forceinline__ void add(double x)
    {
        if (in_int && x == (double)(long long)x && fabs(x) < 9007199254740992.0) { isum += (long long)x; return; }
        if (in_int) { in_int = false; fr = (double)isum; cc = 0.0; }
        neumaier_add(fr, cc, x);
    }
    __device__ __forceinline__ double value() const
    {
        if (in_int) return (double)isum;
        return (cc != 0.0 && cc == cc && fabs(cc) <= 1.79769313486231570815e308) ? __dadd_rn(fr, cc) : fr;
    }
};

// key of get_best_orthologs on the QUERY bed12 line (pipeline.py:981-987) for the chosen chain
__device__ double best_key_of_chain(const ChainArgs& A, long long h0, const int* slots, int len, int which)
{
    long long blen = 0; int lo = INT32_MAX, hi = INT32_MIN; PySum w;
    for (int c = 0; c < len; ++c) {
        const int4 q = A.ret_coord[h0 + slots[c]];
        int qs = q.x, qe = q.y;
        int d = qe - qs; blen += d < 0 ? -d : d;
        lo = min(lo, min(qs, qe)); hi = max(hi, max(qs, qe));
        if (which == 3) w.add(A.ret_score[h0 + slots[c]]);
    }
    if (which == 0) return (double)blen;
    if (which == 1) return (double)((long long)hi - (long long)lo);
    if (which == 2) return (double)len;
    return w.value();                                  // weight = float(str(sum(scores))), pipeline.py:986
}

// Alignments the warp path of the chaining cannot take: too many retained HSPs, or sentinels beyond 32 bits
__device__ __forceinline__ bool chain_needs_cta(int r, long long qlen, long long slen)
{
    return r > CHAIN_WARP_MAX || qlen > 0x7fffffffll || slen > 0x7fffffffll || qlen < 0 || slen < 0;
}

// Gather stage. The retained HSPs of alignment a (list order) live in the slot [hsp_off[a], hsp_off[a] + n_keep[a]) of
//   ret_slot (position in the survivor list), ret_score, ret_coord (qstart, qend, sstart, send).
// k_front fills ret_slot / ret_score (and ret_coord when the coordinate records are in device memory) for the alignments
// it finished itself; what is left:
//   k_retained_slots  ret_slot / ret_score + queueing for k_chain_small, from the keep flags, for the alignments of
//                     work_list (those that went through k_pq) or for all (work_list == nullptr)
//   k_fetch_coords    ret_coord from device-readable coordinate records — device memory, or pinned host memory read in
//                     place over PCIe (zero-copy: the only traffic the coordinate arrays ever cause, ~1 % of the HSPs)
//   k_list_retained / k_scatter_coords   the host-gather mode (see below)
__global__ void __launch_bounds__(256)
k_retained_slots(ChainArgs A)
{
    const int lane = lane_id();
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long n_items = A.work_list ? (long long)*A.work_count : A.n_aln;
    for (long long it = warp; it < n_items; it += nwarps) {
        const long long a = A.work_list ? (long long)A.work_list[it] : it;
        const int r = A.n_keep[a];
        if (r == 0) continue;
        if (lane == 0 && chain_needs_cta(r, A.qlen[a], A.slen[a])) { int slot = atomicAdd(A.mid_count, 1); A.mid_list[slot] = (int)a; }
        const long long h0 = A.hsp_off[a];
        const int ns = A.n_surv[a];
        int got = 0;
        for (int k0 = 0; k0 < ns && got < r; k0 += 32) {
            const int k = k0 + lane;
            const bool keep = (k < ns) && A.surv_keep[h0 + k];
            const unsigned b = __ballot_sync(FULL_MASK, keep);
            if (keep) {
                const long long pos = h0 + got + __popc(b & lanemask_lt());
                A.ret_slot[pos] = k; A.ret_score[pos] = A.surv_x[h0 + k];
            }
            got += __popc(b);
        }
    }
}

__global__ void __launch_bounds__(256)
k_fetch_coords(ChainArgs A)
{
    const int lane = lane_id();
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long n_items = A.work_list ? (long long)*A.work_count : A.n_aln;
    for (long long it = warp; it < n_items; it += nwarps) {
        const long long a = A.work_list ? (long long)A.work_list[it] : it;
        const int r = A.n_keep[a];
        const long long h0 = A.hsp_off[a];
        for (int i = lane; i < r; i += 32) {
            const long long h = h0 + A.surv_idx[h0 + A.ret_slot[h0 + i]];
            int4 c;
            if (A.coords4) c = A.coords4[h];
            else { c.x = A.qstart[h]; c.y = A.qend[h]; c.z = A.sstart[h]; c.w = A.send[h]; }
            A.ret_coord[h0 + i] = c;
        }
    }
}

// Host-gather mode (host batches whose coordinate records stay in HOST memory): instead of reading the ~1 % retained
// records in place over PCIe (32-byte reads that compete with the upload DMA for the link's read-request slots), the
// GPU lists the global HSP indices of the retained HSPs straight into mapped pinned host memory (posted writes), the
// host gathers the 16-byte records with its own threads and sends them back as one small DMA, and k_scatter_coords
// puts them where the chaining kernels expect them.
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
        long long base = 0;
        if (lane == 0) { base = (long long)atomicAdd((unsigned long long*)cursor, (unsigned long long)r); ret_base[a] = base; }
        base = __shfl_sync(FULL_MASK, base, 0);
        const long long h0 = A.hsp_off[a];
        for (int i = lane; i < r; i += 32)
            if (base + i < cap) host_list[base + i] = h0 + A.surv_idx[h0 + A.ret_slot[h0 + i]];
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

// Warp path: one warp per alignment with at most CHAIN_WARP_MAX retained HSPs (the common case: ~10-20 retained of
// ~130 survivors), entirely in REGISTERS: every vertex of an orientation graph — begin sentinel, HSPs in
// (qstart, sstart, list order), end sentinel — is one lane; a vertex's data is broadcast with shuffles when it is its
// turn. The two graphs are independent (alignment.py:1015-1030): when both fit the warp together (n_direct + n_reverse +
// 4 sentinels <= 32) the direct graph takes lanes [0, N0), the reverse graph [N0, N0 + N1) and they advance in the
// same instructions (at step s every lane looks at vertex s of ITS graph); otherwise they are chained one after the other.
//   pass 1  successor masks M_i = {j > i : precede(i, j)} by ballot; f(i) = lowest bit; edges of i = M_i below f(f(i))
//   pass 2  push DP in vertex order: the lanes in the edge mask relax themselves (distinct targets, strict '>')
//   result  chain = vertices on the prev-path from the end sentinel; ascending vertex index IS chain order
// Coordinates and sentinels are 32-bit here: alignments whose qlen / slen do not fit (or with negative coordinates) are
// queued for k_chain_small like the ones with too many retained HSPs (chain_needs_cta).
// (Round 1 kept the graphs in shared memory as 64-bit arrays: 2.6 k warp instructions per alignment.)
constexpr int CHAINW_WARPS = 4;
constexpr int CHAINW_SMEM = 128;          // bytes of shared memory per warp: inverse permutation, two staged chains, flags

struct WarpChainResult { int pick; int len; double sco0, sco1; double key; bool bad; };

// All 32 lanes call this. Lane p < r holds the p-th retained HSP (list order); r <= CHAIN_WARP_MAX and qlen, slen fit
// int32. `sm`: CHAINW_SMEM bytes of shared memory owned by the warp. The chosen chain (retained-list positions, chain
// order) goes to chain_out[0..len).
//
// Inside, the subject coordinates of the REVERSE graph are stored bit-complemented (~x = -x - 1 reverses the order and
// cannot overflow): `a.send > b.sstart` becomes `~a.send < ~b.sstart` and `a.send - b.sstart` becomes
// `~b.sstart - ~a.send`, so both graphs run the direct graph's formulas
//     precede(i, j) = qend_i < qstart_j and send_i < sstart_j            (alignment.py:137-168)
//     distance      = (qstart_j - qend_i) + (sstart_j - send_i)          (alignment.py:170-217)
// and a step of either pass is the same dozen instructions for every lane. The integer distance is evaluated as
// (qstart_j + sstart_j) - (qend_i + send_i) — exact in 64 bits, hence the same double as the reference's.
__device__ __forceinline__ WarpChainResult warp_chain(int r, int qs, int qe, int ss, int se, double sc, int qlen, int slen,
                                                      double G, double E, int best_value, unsigned char* sm,
                                                      int* __restrict__ chain_out)
{
    const int lane = lane_id();
    const unsigned lt = lanemask_lt();
    const double NEG_INF = __longlong_as_double(0xfff0000000000000ll);
    const double POS_INF = __longlong_as_double(0x7ff0000000000000ll);
    unsigned char* inv = sm;                                      // vertex lane -> retained-list position
    WarpChainResult R; R.pick = 1; R.len = 0; R.sco0 = R.sco1 = NEG_INF; R.key = 0.0; R.bad = false;
    const bool mine = lane < r;
    const bool dir = mine && (se - ss > 0);                       // alignment.py:107-110
    // orientation None: the reference raises in _split_orientations
    if (__any_sync(FULL_MASK, mine && !(qe - qs > 0))) { R.bad = true; return R; }
    const unsigned mem_d = __ballot_sync(FULL_MASK, dir), mem_r = __ballot_sync(FULL_MASK, mine && !dir);
    const int n0 = __popc(mem_d), n1 = __popc(mem_r);
    // stable rank by (qstart, sstart, list order) among the HSPs of the own orientation:
    //   #{peers with a smaller key} + #{earlier peers with the same key}
    const unsigned peers = dir ? mem_d : mem_r;
    const unsigned long long key = ((unsigned long long)((unsigned)qs ^ 0x80000000u) << 32) | (unsigned long long)((unsigned)ss ^ 0x80000000u);
    int rank = __popc(__match_any_sync(FULL_MASK, key) & peers & lt);
#pragma unroll 1
    for (int q = 0; q < r; ++q) {
        const unsigned long long ok = __shfl_sync(FULL_MASK, key, q);
        if (((peers >> q) & 1u) && ok < key) ++rank;
    }
    const bool together = (n0 ? n0 + 2 : 0) + (n1 ? n1 + 2 : 0) <= 32;
    int len_d = 0, len_r = 0;
    sm[96 + lane] = 0;                                            // flag[p]: retained HSP p is on its graph's best chain
#pragma unroll 1
    for (int round = 0; round < (together ? 1 : 2); ++round) {
        const bool use_d = together || round == 0, use_r = together || round == 1;
        const int N0 = (use_d && n0) ? n0 + 2 : 0, N1 = (use_r && n1) ? n1 + 2 : 0;     // vertices incl. sentinels
        __syncwarp();
        if (mine && (dir ? use_d : use_r)) inv[(dir ? 0 : N0) + 1 + rank] = (unsigned char)lane;
        __syncwarp();
        // ---- my vertex: lane v of graph D = [0, N0) or R = [N0, N0 + N1) ---------------------------------------------
        const int v = lane;
        const bool active = v < N0 + N1;
        const bool gD = v < N0;
        const int base = gD ? 0 : N0, Nn = gD ? N0 : N1;
        const int me = active ? v - base : -1;                    // my index inside my graph (-1: no vertex)
        const bool is_begin = me == 0, is_end = active && me == Nn - 1;
        const bool is_hsp = active && !is_begin && !is_end;
        const int src = is_hsp ? inv[v] : 0;                      // retained-list position of my HSP
        int Qs = __shfl_sync(FULL_MASK, qs, src), Qe = __shfl_sync(FULL_MASK, qe, src);
        int Ss = __shfl_sync(FULL_MASK, ss, src), Se = __shfl_sync(FULL_MASK, se, src);
        double Sc = __shfl_sync(FULL_MASK, sc, src);
        if (is_begin) { Qs = 0; Qe = 0; Ss = gD ? 0 : slen; Se = Ss; Sc = 0.0; }        // alignment.py:893-902
        if (is_end) { Qs = qlen; Qe = qlen; Ss = gD ? slen : 0; Se = Ss; Sc = 0.0; }
        if (!gD) { Ss = ~Ss; Se = ~Se; }                          // reverse graph: complemented subject axis
        // what an edge INTO me adds to the distance and what an edge OUT of me subtracts, as doubles: both sums are
        // integers below 2^33, so is their difference, hence din - dout_i is exact and equals (double)(in_sum - out_sum_i)
        // — one DADD per step instead of a 64-bit subtraction and an I2F.F64.S64
        const double din = (double)((long long)Qs + (long long)Ss);
        const double dout = (double)((long long)Qe + (long long)Se);
        const unsigned below_end = (N0 + N1) >= 32 ? 0xffffffffu : ((1u << (N0 + N1)) - 1u);
        const unsigned gmask = !active ? 0u : (gD ? ((N0 >= 32) ? 0xffffffffu : ((1u << N0) - 1u)) : (below_end & ~((1u << N0) - 1u)));
        const int steps = max(N0, N1);
        // ---- pass 1: successor masks M_i = {j > i : precede(i, j)} ---------------------------------------------------
        unsigned M = 0;
#pragma unroll 1
        for (int s = 0; s < steps; ++s) {
            const int i = (base + s) & 31;                        // the vertex of my graph whose turn it is
            const int qe_i = __shfl_sync(FULL_MASK, Qe, i), se_i = __shfl_sync(FULL_MASK, Se, i);
            const unsigned bm = __ballot_sync(FULL_MASK, me > s && qe_i < Qs && se_i < Ss);
            if (me == s) M = bm & gmask;
        }
        // the begin-end pair is tested with the REVERSE rule in both graphs (both sentinels have orientation None,
        // alignment.py:162-165): in the direct graph that is 0 > slen, never true
        if (is_begin && gD) M &= ~(1u << (N0 - 1));
        const int f = M ? __ffs(M) - 1 : 32;                      // first successor (32: none)
        const int ff = __shfl_sync(FULL_MASK, f, f & 31);
        const int g = f >= 32 ? 32 : ff;                          // f(f(i))
        const unsigned Ed = g >= 32 ? M : (M & ((1u << g) - 1u)); // kept edges: precede(i, j) and j < f(f(i))
        // ---- pass 2: push DP in vertex order (HSPVertex._relax, alignment.py:388-411): strict '>', first wins ----------
        double tot = is_begin ? 0.0 : POS_INF;
        int prev = -1;
#pragma unroll 1
        for (int s = 0; s + 1 < steps; ++s) {
            const int i = (base + s) & 31;
            const unsigned e_i = __shfl_sync(FULL_MASK, Ed, i);
            const double t_i = __shfl_sync(FULL_MASK, tot, i);
            const double o_i = __shfl_sync(FULL_MASK, dout, i);
            // an unreached source carries +inf and cannot improve anything; a lane outside the edge mask does not relax
            const double nsc = __dadd_rn(t_i, __dsub_rn(__dadd_rn(G, __dmul_rn(E, __dsub_rn(din, o_i))), Sc));
            if (((e_i >> v) & 1u) && me > s && tot > nsc) { tot = nsc; prev = i; }
        }
        // ---- scores and backtrack (alignment.py:1007-1013) --------------------------------------------------------
        const double t0 = __shfl_sync(FULL_MASK, tot, (N0 - 1) & 31), t1 = __shfl_sync(FULL_MASK, tot, (N0 + N1 - 1) & 31);
        if (N0) R.sco0 = -t0;
        if (N1) R.sco1 = -t1;
        unsigned cm = 0;                                          // vertices on the path of MY graph (begin included)
        {
            int cur = active ? base + Nn - 1 : 0;
#pragma unroll 1
            for (int s = 0; s + 1 < steps; ++s) {
                const int pv = __shfl_sync(FULL_MASK, prev, cur & 31);
                if (active && pv >= 0) { cm |= 1u << pv; cur = pv; }
            }
        }
        const bool on_chain = is_hsp && ((cm >> v) & 1u);
        const unsigned cb_d = __ballot_sync(FULL_MASK, on_chain && gD), cb_r = __ballot_sync(FULL_MASK, on_chain && !gD);
        // stage both chains (retained-list positions in chain order = ascending vertex index)
        if (on_chain) { sm[32 + (gD ? 0 : 32) + __popc((gD ? cb_d : cb_r) & lt)] = (unsigned char)src; sm[96 + src] = 1; }
        if (N0) len_d = __popc(cb_d);
        if (N1) len_r = __popc(cb_r);
    }
    __syncwarp();
    R.pick = (R.sco0 > R.sco1) ? 0 : 1;                           // alignment.py:1043-1046
    R.len = R.pick ? len_r : len_d;
    const unsigned char* staged = sm + 32 + (R.pick ? 32 : 0);
    if (lane < R.len) chain_out[lane] = staged[lane];
    // ---- key of get_best_orthologs on the query BED12 line of the chosen chain (pipeline.py:981-987) -----------------
    if (R.len) {
        const bool on_chain = mine && sm[96 + lane] != 0 && (dir ? R.pick == 0 : R.pick == 1);   // lane p holds retained HSP p
        if (best_value == 0) {                                    // block_length: sum of abs(qend - qstart)
            long long d = on_chain ? (long long)(qe - qs < 0 ? qs - qe : qe - qs) : 0;
#pragma unroll
            for (int k = 16; k > 0; k >>= 1) d += __shfl_xor_sync(FULL_MASK, d, k);
            R.key = (double)d;
        } else if (best_value == 1) {                             // total_length: chromEnd - chromStart
            int lo = on_chain ? min(qs, qe) : INT32_MAX, hi = on_chain ? max(qs, qe) : INT32_MIN;
            lo = __reduce_min_sync(FULL_MASK, lo); hi = __reduce_max_sync(FULL_MASK, hi);
            R.key = (double)((long long)hi - (long long)lo);
        } else if (best_value == 2) R.key = (double)R.len;        // block_count
        else {
            // weight = float(str(sum(scores))) of the BED score column (genomicranges.py:1225): CPython >= 3.12's sum()
            // adds ints exactly while every item is an int, then switches to Neumaier-compensated float addition
            PySum w;
#pragma unroll 1
            for (int c = 0; c < R.len; ++c) w.add(__shfl_sync(FULL_MASK, sc, staged[c]));
            R.key = w.value();
        }
    }
    return R;
}

// 72 registers: 7 CTAs per SM. Measured: capping at 64 registers (8 CTAs, 20 bytes spilled) 0.2625 -> 0.271 ms per C2 launch;
// a minimum-blocks hint of 1 makes ptxas take 75 registers (6 CTAs): 0.287 ms.
__global__ void __launch_bounds__(CHAINW_WARPS * 32)
k_chain_warp(ChainArgs A)
{
    __shared__ unsigned char s_mem[CHAINW_WARPS][CHAINW_SMEM];
    const int lane = lane_id(), w = warp_id();
    const double NEG_INF = __longlong_as_double(0xfff0000000000000ll);
    const long long nwarps_total = (long long)gridDim.x * CHAINW_WARPS;
    for (long long a = (long long)blockIdx.x * CHAINW_WARPS + w; a < A.n_aln; a += nwarps_total) {
        const int r = A.n_keep[a];
        const int l = A.aln_lnc[a];
        const long long qlen = A.qlen[a], slen = A.slen[a];
        if (r > 0 && chain_needs_cta(r, qlen, slen)) continue;    // queued for k_chain_small by the gather
        WarpChainResult R; R.pick = 1; R.len = 0; R.sco0 = R.sco1 = NEG_INF; R.key = 0.0; R.bad = false;
        const long long h0 = A.hsp_off[a];
        if (r > 0 && A.lnc_status[l] < 2) {
            int qs = 0, qe = 0, ss = 0, se = 0; double sc = 0.0;
            if (lane < r) {
                const int4 c = A.ret_coord[h0 + lane];
                qs = c.x; qe = c.y; ss = c.z; se = c.w; sc = A.ret_score[h0 + lane];
            }
            R = warp_chain(r, qs, qe, ss, se, sc, (int)qlen, (int)slen, (double)A.gapopen, (double)A.gapextend,
                           A.best_value, s_mem[w], A.chain_slot + h0);
        }
        if (lane == 0) {
            if (R.bad) atomicMax(&A.lnc_status[l], 1);   // the reference raises in _split_orientations
            A.chain_len[a] = R.len; A.chain_orient[a] = (unsigned char)R.pick;
            A.chain_score[a] = R.pick ? R.sco1 : R.sco0;
            A.orient_scores[2 * a] = R.sco0; A.orient_scores[2 * a + 1] = R.sco1;
            A.best_key[a] = R.len ? R.key : 0.0;
        }
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
