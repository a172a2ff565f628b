// o2a_api.cu — C ABI (include/o2a.h) of libo2a_cuda.so: context, device workspace, stage launches.
// sm_100a only; there is no CPU fallback anywhere in this library.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <map>
#include <vector>

#include "../../include/o2a.h"
#include "o2a_annotate.cuh"
#include "o2a_chain.cuh"
#include "o2a_cut.cuh"
#include "o2a_device.cuh"
#include "o2a_fit.cuh"
#include "o2a_front.cuh"
#include "o2a_probe.cuh"
#include "o2a_score.cuh"

using namespace o2a;

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

constexpr int N_EVENTS = O2A_STAGE_COUNT + 1;

// A few persistent host threads for the host-gather mode (collecting the retained HSPs' coordinate records from the
// caller's host arrays): started on first use, parked on a condition variable between jobs.
class HostPool {
public:
    explicit HostPool(int n) : n_(n < 1 ? 1 : n) {
        for (int i = 0; i < n_; ++i) threads_.emplace_back([this, i]() { worker(i); });
    }
    ~HostPool() {
        { std::lock_guard<std::mutex> g(m_); stop_ = true; ++gen_; }
        cv_.notify_all();
        for (auto& t : threads_) t.join();
    }
    int size() const { return n_; }
    // fn(k) for k in [0, n): contiguous blocks, one per worker; returns when all are done
    void run(long long n, const std::function<void(long long)>& fn) {
        if (n <= 0) return;
        if (n < 8192) { for (long long k = 0; k < n; ++k) fn(k); return; }
        {
            std::lock_guard<std::mutex> g(m_);
            fn_ = &fn; total_ = n; pending_ = n_; ++gen_;
        }
        cv_.notify_all();
        std::unique_lock<std::mutex> g(m_);
        done_.wait(g, [this]() { return pending_ == 0; });
        fn_ = nullptr;
    }
private:
    void worker(int id) {
        unsigned long long seen = 0;
        for (;;) {
            const std::function<void(long long)>* fn; long long n;
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&]() { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                fn = fn_; n = total_;
            }
            const long long lo = n * id / n_, hi = n * (id + 1) / n_;
            for (long long k = lo; k < hi; ++k) (*fn)(k);
            {
                std::lock_guard<std::mutex> g(m_);
                if (--pending_ == 0) done_.notify_one();
            }
        }
    }
    int n_;
    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(long long)>* fn_ = nullptr;
    long long total_ = 0;
    int pending_ = 0;
    unsigned long long gen_ = 0;
    bool stop_ = false;
};
constexpr int SM_COUNT_FALLBACK = 148;
constexpr long long UV_MAX_RANGE = 1ll << 18;   // 1 MiB of counters per resident CTA

}  // namespace

struct o2a_ctx {
    int device = 0;
    int sm_count = SM_COUNT_FALLBACK;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    std::string err;
    long long launches = 0;
    bool timing = false;
    cudaEvent_t ev[N_EVENTS] = {};
    bool ev_valid = false;
    // batch geometry of the last build
    bool have_result = false;
    o2a_batch b{};          // device pointers of the inputs actually used
    o2a_params prm{};
    o2a_counts counts{};
    // owned copies of host inputs
    DevBuf in_bg_off, in_bg, in_kde, in_aln_off, in_hsp_off, in_qlen, in_slen, in_qs, in_qe, in_ss, in_se, in_score;
    // workspace
    DevBuf aln_lnc, thr, fits, tbl_k, tbl_cum, status, uv_val, uv_cnt, uv_n;
    DevBuf n_surv, n_keep, surv_idx, surv_x, surv_p, surv_pq, surv_keep;
    DevBuf in_score_q, in_exc_off, in_exc_pos, in_exc_val;
    DevBuf chain_len, chain_orient, chain_score, orient_scores, chain_slot, best_key, best_aln;
    DevBuf ret_coord, ret_score, ret_slot, in_coords;
    const int4* coords4 = nullptr;    // AoS coordinates actually used (device or mapped host), or null
    DevBuf big_list, mid_list, big_count, counters;
    // per-chunk work lists / counters: stages of different chunks overlap on different streams, so nothing is shared.
    // stage_cnt: 4 ints per chunk = (alignments queued for k_pq, for k_rank_big, for k_chain_small, for k_chain_big)
    DevBuf pq_list, cbig_list, stage_cnt;
    DevBuf csort_k0, csort_k1, csort_v0, csort_v1;   // k_chain_big's sort scratch (k_rank_big has sort_*)
    DevBuf maxchk;                                   // device-side check of the caller's max_aln_hsps / max_lnc_bg
    // coordinates of a host batch: how the retained HSPs' records reach the chaining kernels
    int coords_mode_req = O2A_COORDS_AUTO, coords_mode_used = O2A_COORDS_AUTO;
    bool coords_on_device = false;                   // the coordinate records are in device memory: k_front reads them itself
    cudaStream_t tail_stream = nullptr;              // host-gather mode: scatter + chaining of chunk c beside the front of chunk c+1
    cudaEvent_t front_ev[16] = {};
    cudaEvent_t tail_join = nullptr;
    DevBuf ret_base, list_cursor, d_stage;
    long long* h_list = nullptr; size_t h_list_cap = 0;     // mapped pinned: global HSP indices of the retained HSPs
    int4* h_stage = nullptr; size_t h_stage_cap = 0;        // pinned: their gathered coordinate records
    long long* h_cursor = nullptr;                          // pinned [16]: retained HSPs listed per chunk
    int gather_threads = 0;
    HostPool* pool = nullptr;
    DevBuf fit_cnt, fc_val, fc_cnt, fc_n, fit_list, fit_list_n, fit_list2, fit_list2_n, fit_scratch;
    bool coords_zero_copy = false;   // qstart..send read in place from pinned, device-mapped host memory
    // chunk view [vl0, vl1) lncRNAs / [va0, va1) alignments the stage launchers work on (whole batch by default)
    long long vl0 = 0, vl1 = 0, va0 = 0, va1 = 0;
    cudaStream_t aux_stream = nullptr;       // kernels over disjoint alignment classes run beside the main stream
    cudaEvent_t aux_fork = nullptr, aux_join = nullptr;
    cudaStream_t copy_stream = nullptr;      // host path: H2D of chunk c+1 overlaps the kernels of chunk c
    cudaEvent_t copy_ev[16] = {};
    int pipelined_chunks = 0;
    // CUDA graph of the launch sequence of a device-resident batch (same pointers, sizes and parameters as the call
    // before): captured on the second identical call, replayed from the third; dropped when a buffer is reallocated
    unsigned long long alloc_epoch = 0;
    std::vector<long long> graph_key, seen_key;
    cudaGraphExec_t graph_exec = nullptr;
    long long graph_launches = 0;
    bool graph_disabled = false;
    int graph_state = 0;                     // last call: 0 eager, 1 captured + launched, 2 replayed
    DevBuf sort_k0, sort_k1, sort_v0, sort_v1;
    DevBuf cb_qs, cb_qe, cb_ss, cb_se, cb_sc, cb_id, cb_total, cb_prev, cb_f, cb_chain;
    DevBuf off_a, off_b, out_i32, out_f64a, out_f64b, out_u8, maxred;
    // producer side (o2a_cut_batch)
    DevBuf ci_isint, ci_qstart, ci_qend, ci_sstart, ci_send, ci_qminus, ci_sminus, ci_rel, ci_ql, ci_qr, ci_sl, ci_sr;
    DevBuf ct_n, ct_off, ct_aln, ct_cnt, ct_out, ct_sums, ct_maxq, ct_maxs, ct_gen, ct_gen_n, ct_fast, ct_nal, ct_span;
    DevBuf co_off, co_qs, co_qe, co_ss, co_se, co_score, co_cut, co_qlen, co_slen, co_status;
    bool have_cut = false;
    long long cut_n_aln = 0, cut_n_kept = 0;
    cudaEvent_t cut_ev[5] = {};
    bool cut_ev_valid = false;
    // annotate_orthologs' interval join (o2a_annotate_batch)
    DevBuf an_qc, an_qs, an_qe, an_qn, an_qstr, an_ac, an_as, an_ae, an_an, an_astr;
    DevBuf an_prefmax, an_chunkmax, an_blkmax, an_cnt, an_span, an_off, an_sums, an_n, an_idx, an_ji, an_oc, an_pq;
    bool have_ann = false;
    long long ann_n_q = 0, ann_n_pairs = 0;
    cudaEvent_t ann_ev[5] = {};
    bool ann_ev_valid = false;
};

#define CU(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);                       \
            return O2A_ECUDA;                                                                     \
        }                                                                                         \
    } while (0)

#define LAUNCH_CHECK()                                                                            \
    do {                                                                                          \
        ctx->launches++;                                                                          \
        cudaError_t e__ = cudaGetLastError();                                                     \
        if (e__ != cudaSuccess) {                                                                 \
            ctx->err = std::string("kernel launch: ") + cudaGetErrorString(e__);                  \
            return O2A_ECUDA;                                                                     \
        }                                                                                         \
    } while (0)

static int ensure(o2a_ctx* ctx, DevBuf& b, size_t bytes)
{
    if (bytes == 0) bytes = 16;
    if (b.cap >= bytes) return 0;
    if (b.p) { cudaFree(b.p); b.p = nullptr; b.cap = 0; }
    ctx->alloc_epoch++;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        ctx->err = std::string("cudaMalloc(") + std::to_string(want) + "): " + cudaGetErrorString(e);
        cudaGetLastError();
        return O2A_ENOMEM;
    }
    b.cap = want;
    return 0;
}
#define ENS(buf, bytes) do { int r__ = ensure(ctx, ctx->buf, (size_t)(bytes)); if (r__) return r__; } while (0)

template <typename T> static T* P(const DevBuf& b) { return reinterpret_cast<T*>(b.p); }

// --------------------------------------------------------------------------------------------
// small utility kernels
// --------------------------------------------------------------------------------------------
__global__ void k_fill_aln_lnc(long long n_lnc, const long long* __restrict__ aln_off, int* __restrict__ aln_lnc)
{
    for (long long l = blockIdx.x; l < n_lnc; l += gridDim.x)
        for (long long a = aln_off[l] + threadIdx.x; a < aln_off[l + 1]; a += blockDim.x) aln_lnc[a] = (int)l;
}

__global__ void k_max_diff(long long n, const long long* __restrict__ off, unsigned long long* __restrict__ out)
{
    long long m = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        m = max(m, off[i + 1] - off[i]);
    for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_xor_sync(FULL_MASK, m, d));
    if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)m);
}

__global__ void k_counts(long long n_aln, long long n_lnc, const int* __restrict__ n_surv, const int* __restrict__ n_keep,
                         const int* __restrict__ chain_len, const int* __restrict__ status,
                         unsigned long long* __restrict__ out /*5*/)
{
    unsigned long long c[5] = {0, 0, 0, 0, 0};
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x; a < n_aln; a += stride) {
        c[0] += n_surv[a]; c[1] += n_keep[a]; c[2] += chain_len[a]; c[3] += chain_len[a] > 0;
    }
    for (long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x; l < n_lnc; l += stride) c[4] += status[l] != 0;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        for (int d = 16; d > 0; d >>= 1) c[k] += __shfl_xor_sync(FULL_MASK, c[k], d);
        if ((threadIdx.x & 31) == 0 && c[k]) atomicAdd(&out[k], c[k]);
    }
}

// exclusive scan of int counts -> int64 offsets [n+1]; single CTA, sequential over chunks
__global__ void k_scan_offsets(long long n, const int* __restrict__ cnt, long long* __restrict__ off)
{
    __shared__ int s_scan[1024 / 32 + 2];
    __shared__ long long s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (long long i0 = 0; i0 < n; i0 += blockDim.x) {
        long long i = i0 + threadIdx.x;
        int v = i < n ? cnt[i] : 0;
        int tot;
        int ex = cta_exclusive_scan(v, s_scan, &tot);
        long long base = s_base;
        if (i < n) off[i] = base + ex;
        __syncthreads();
        if (threadIdx.x == 0) s_base = base + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) off[n] = s_base;
}

__global__ void k_gather_surv(long long n_aln, const long long* __restrict__ hsp_off, const long long* __restrict__ off,
                              const int* __restrict__ surv_idx, const double* __restrict__ surv_p,
                              const double* __restrict__ surv_pq, const unsigned char* __restrict__ surv_keep,
                              int* __restrict__ o_idx, double* __restrict__ o_p, double* __restrict__ o_pq,
                              unsigned char* __restrict__ o_keep)
{
    for (long long a = blockIdx.x; a < n_aln; a += gridDim.x) {
        const long long h0 = hsp_off[a], o0 = off[a];
        const int n = (int)(off[a + 1] - o0);
        for (int k = threadIdx.x; k < n; k += blockDim.x) {
            o_idx[o0 + k] = surv_idx[h0 + k]; o_p[o0 + k] = surv_p[h0 + k];
            o_pq[o0 + k] = surv_pq[h0 + k]; o_keep[o0 + k] = surv_keep[h0 + k];
        }
    }
}

__global__ void k_gather_chain(long long n_aln, const long long* __restrict__ hsp_off, const long long* __restrict__ off,
                               const int* __restrict__ chain_slot, const int* __restrict__ ret_slot,
                               const int* __restrict__ surv_idx,
                               const double* __restrict__ surv_pq, int* __restrict__ o_idx, double* __restrict__ o_pq)
{
    for (long long a = blockIdx.x; a < n_aln; a += gridDim.x) {
        const long long h0 = hsp_off[a], o0 = off[a];
        const int n = (int)(off[a + 1] - o0);
        for (int c = threadIdx.x; c < n; c += blockDim.x) {
            int s = ret_slot[h0 + chain_slot[h0 + c]];     // retained-list position -> survivor slot
            o_idx[o0 + c] = surv_idx[h0 + s]; o_pq[o0 + c] = surv_pq[h0 + s];
        }
    }
}

// sf at arbitrary points for the fitter seam (o2a_fitter_eval)
__global__ void k_sf_eval(int fitter, const double* __restrict__ x, long long n_x, const LncFit* __restrict__ fits,
                          const int* __restrict__ tbl_k, const double* __restrict__ tbl_cum,
                          const int* __restrict__ uv_val, const int* __restrict__ uv_cnt, const int* __restrict__ uv_n,
                          double kde_factor, double n_bg, double* __restrict__ out)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_x) return;
    double v = x[i], p;
    if (fitter == 0) { LncFit f = fits[0]; p = hist_sf(v, f, tbl_k, tbl_cum); }
    else if (fitter == 1) p = kde_sf_compressed(v, uv_val, uv_cnt, uv_n[0], kde_factor, n_bg);
    else p = emp_sf_compressed(v, uv_val, uv_cnt, uv_n[0], n_bg);
    out[i] = p;
}

__global__ void k_iota_keep(int n, const double* __restrict__ score, int* __restrict__ idx, unsigned char* __restrict__ keep,
                            double* __restrict__ surv_x)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { idx[i] = i; keep[i] = 1; surv_x[i] = score[i]; }
}

// --------------------------------------------------------------------------------------------
// context
// --------------------------------------------------------------------------------------------
extern "C" int o2a_abi_version(void) { return O2A_ABI_VERSION; }

extern "C" int o2a_create(o2a_ctx** out, int device)
{
    if (!out) return O2A_EINVAL;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
        cudaGetLastError();
        return O2A_ECUDA;   // no CUDA device: there is no CPU fallback
    }
    o2a_ctx* ctx = new o2a_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return O2A_ECUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return O2A_ECUDA; }
    for (int i = 0; i < N_EVENTS; ++i) cudaEventCreate(&ctx->ev[i]);
    cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    {   // highest priority: its few CTAs are placed before the main stream's machine-filling grids
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        cudaStreamCreateWithPriority(&ctx->aux_stream, cudaStreamNonBlocking, hi);
    }
    cudaEventCreateWithFlags(&ctx->aux_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->aux_join, cudaEventDisableTiming);
    for (int i = 0; i < 16; ++i) cudaEventCreateWithFlags(&ctx->copy_ev[i], cudaEventDisableTiming);
    for (int i = 0; i < 16; ++i) cudaEventCreateWithFlags(&ctx->front_ev[i], cudaEventDisableTiming);
    cudaStreamCreateWithFlags(&ctx->tail_stream, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&ctx->tail_join, cudaEventDisableTiming);
    cudaHostAlloc((void**)&ctx->h_cursor, 16 * sizeof(long long), cudaHostAllocPortable);
    {
        // host-gather threads: the box's cores shared among the ranks of this node (torchrun exports LOCAL_WORLD_SIZE)
        unsigned hc = std::thread::hardware_concurrency();
        int share = 1;
        if (const char* e = getenv("LOCAL_WORLD_SIZE")) { if (atoi(e) > 1) share = atoi(e); }
        int nt = hc ? (int)hc / share : 4;
        ctx->gather_threads = nt < 2 ? 2 : (nt > 16 ? 16 : nt);
        if (const char* e = getenv("O2A_GATHER_THREADS")) { if (atoi(e) > 0) ctx->gather_threads = atoi(e); }
    }
    for (int i = 0; i < 5; ++i) cudaEventCreate(&ctx->cut_ev[i]);
    for (int i = 0; i < 5; ++i) cudaEventCreate(&ctx->ann_ev[i]);
    *out = ctx;
    return 0;
}

static void free_all(o2a_ctx* ctx)
{
    DevBuf* bufs[] = {&ctx->in_bg_off, &ctx->in_bg, &ctx->in_kde, &ctx->in_aln_off, &ctx->in_hsp_off, &ctx->in_qlen,
                      &ctx->in_slen, &ctx->in_qs, &ctx->in_qe, &ctx->in_ss, &ctx->in_se, &ctx->in_score,
                      &ctx->aln_lnc, &ctx->thr, &ctx->fits, &ctx->tbl_k, &ctx->tbl_cum, &ctx->status, &ctx->uv_val,
                      &ctx->uv_cnt, &ctx->uv_n, &ctx->n_surv, &ctx->n_keep, &ctx->surv_idx, &ctx->surv_x, &ctx->in_score_q,
                      &ctx->in_exc_off, &ctx->in_exc_pos, &ctx->in_exc_val, &ctx->surv_p,
                      &ctx->surv_pq, &ctx->surv_keep, &ctx->chain_len, &ctx->chain_orient, &ctx->chain_score,
                      &ctx->orient_scores, &ctx->chain_slot, &ctx->best_key, &ctx->best_aln, &ctx->ret_coord,
                      &ctx->ret_score, &ctx->ret_slot, &ctx->in_coords, &ctx->big_list,
                      &ctx->mid_list, &ctx->big_count, &ctx->counters, &ctx->pq_list, &ctx->cbig_list, &ctx->stage_cnt,
                      &ctx->csort_k0, &ctx->csort_k1, &ctx->csort_v0, &ctx->csort_v1, &ctx->maxchk, &ctx->ret_base,
                      &ctx->list_cursor, &ctx->d_stage, &ctx->fit_list2, &ctx->fit_list2_n, &ctx->fit_scratch, &ctx->fit_cnt, &ctx->fc_val, &ctx->fc_cnt, &ctx->fc_n, &ctx->fit_list,
                      &ctx->fit_list_n, &ctx->sort_k0, &ctx->sort_k1, &ctx->sort_v0, &ctx->sort_v1, &ctx->cb_qs,
                      &ctx->cb_qe, &ctx->cb_ss, &ctx->cb_se, &ctx->cb_sc, &ctx->cb_id, &ctx->cb_total, &ctx->cb_prev,
                      &ctx->cb_f, &ctx->cb_chain, &ctx->off_a, &ctx->off_b, &ctx->out_i32, &ctx->out_f64a,
                      &ctx->out_f64b, &ctx->out_u8, &ctx->maxred,
                      &ctx->ci_isint, &ctx->ci_qstart, &ctx->ci_qend, &ctx->ci_sstart, &ctx->ci_send, &ctx->ci_qminus, &ctx->ci_sminus,
                      &ctx->ci_rel, &ctx->ci_ql, &ctx->ci_qr, &ctx->ci_sl, &ctx->ci_sr, &ctx->ct_n, &ctx->ct_off,
                      &ctx->ct_aln, &ctx->ct_cnt, &ctx->ct_out, &ctx->ct_sums, &ctx->ct_maxq, &ctx->ct_maxs, &ctx->ct_gen, &ctx->ct_gen_n, &ctx->ct_fast, &ctx->ct_nal, &ctx->ct_span,
                      &ctx->co_off, &ctx->co_qs, &ctx->co_qe, &ctx->co_ss, &ctx->co_se, &ctx->co_score, &ctx->co_cut,
                      &ctx->co_qlen, &ctx->co_slen, &ctx->co_status,
                      &ctx->an_qc, &ctx->an_qs, &ctx->an_qe, &ctx->an_qn, &ctx->an_qstr, &ctx->an_ac, &ctx->an_as, &ctx->an_ae, &ctx->an_an, &ctx->an_astr,
                      &ctx->an_prefmax, &ctx->an_chunkmax, &ctx->an_blkmax, &ctx->an_cnt, &ctx->an_span, &ctx->an_off, &ctx->an_sums, &ctx->an_n, &ctx->an_idx, &ctx->an_ji, &ctx->an_oc, &ctx->an_pq};
    for (DevBuf* b : bufs) if (b->p) { cudaFree(b->p); b->p = nullptr; b->cap = 0; }
}

extern "C" void o2a_destroy(o2a_ctx* ctx)
{
    if (ctx && ctx->graph_exec) { cudaSetDevice(ctx->device); cudaGraphExecDestroy(ctx->graph_exec); ctx->graph_exec = nullptr; }
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    free_all(ctx);
    for (int i = 0; i < N_EVENTS; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    for (int i = 0; i < 16; ++i) if (ctx->copy_ev[i]) cudaEventDestroy(ctx->copy_ev[i]);
    for (int i = 0; i < 16; ++i) if (ctx->front_ev[i]) cudaEventDestroy(ctx->front_ev[i]);
    if (ctx->tail_join) cudaEventDestroy(ctx->tail_join);
    if (ctx->tail_stream) cudaStreamDestroy(ctx->tail_stream);
    if (ctx->h_list) cudaFreeHost(ctx->h_list);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->h_cursor) cudaFreeHost(ctx->h_cursor);
    delete ctx->pool;
    for (int i = 0; i < 5; ++i) if (ctx->cut_ev[i]) cudaEventDestroy(ctx->cut_ev[i]);
    for (int i = 0; i < 5; ++i) if (ctx->ann_ev[i]) cudaEventDestroy(ctx->ann_ev[i]);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->aux_fork) cudaEventDestroy(ctx->aux_fork);
    if (ctx->aux_join) cudaEventDestroy(ctx->aux_join);
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char* o2a_last_error(const o2a_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }

extern "C" int o2a_set_stream(o2a_ctx* ctx, void* cuda_stream)
{
    if (!ctx) return O2A_EINVAL;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (cuda_stream) { ctx->stream = (cudaStream_t)cuda_stream; ctx->own_stream = false; }
    else {
        ctx->own_stream = true;
        CU(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    }
    return 0;
}

extern "C" void* o2a_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 16, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void o2a_host_free(void* p) { if (p) cudaFreeHost(p); }

extern "C" int64_t o2a_launch_count(const o2a_ctx* ctx) { return ctx ? (int64_t)ctx->launches : 0; }
extern "C" int o2a_graph_state(const o2a_ctx* ctx) { return ctx ? ctx->graph_state : 0; }
extern "C" int o2a_coords_zero_copy(const o2a_ctx* ctx) { return ctx && ctx->coords_zero_copy ? 1 : 0; }
extern "C" int o2a_pipeline_chunks(const o2a_ctx* ctx) { return ctx ? ctx->pipelined_chunks : 0; }
extern "C" int o2a_set_coords_mode(o2a_ctx* ctx, int mode)
{
    if (!ctx || mode < O2A_COORDS_AUTO || mode > O2A_COORDS_UPLOAD) return O2A_EINVAL;
    ctx->coords_mode_req = mode;
    return 0;
}
extern "C" int o2a_coords_mode(const o2a_ctx* ctx) { return ctx ? ctx->coords_mode_used : 0; }
extern "C" int o2a_set_timing(o2a_ctx* ctx, int enabled) { if (!ctx) return O2A_EINVAL; ctx->timing = enabled != 0; return 0; }
extern "C" int o2a_get_stage_ms(o2a_ctx* ctx, float* ms)
{
    if (!ctx || !ms) return O2A_EINVAL;
    if (!ctx->ev_valid) { ctx->err = "no timed build_batch yet (o2a_set_timing)"; return O2A_ESTATE; }
    for (int i = 0; i < O2A_STAGE_COUNT; ++i) CU(cudaEventElapsedTime(&ms[i], ctx->ev[i], ctx->ev[i + 1]));
    return 0;
}

// --------------------------------------------------------------------------------------------
// stages
// --------------------------------------------------------------------------------------------
static int stage_event(o2a_ctx* ctx, int i)
{
    if (ctx->timing) CU(cudaEventRecord(ctx->ev[i], ctx->stream));
    return 0;
}

template <typename T>
static int upload(o2a_ctx* ctx, DevBuf& dst, const T* src, size_t n, const T** out)
{
    int r = ensure(ctx, dst, n * sizeof(T));
    if (r) return r;
    if (n) CU(cudaMemcpyAsync(dst.p, src, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    *out = reinterpret_cast<const T*>(dst.p);
    return 0;
}

// Bulk per-HSP / per-background / per-exception arrays of a host batch: when the call is pipelined the
// copy is deferred and issued slice by slice on the copy stream (o2a_build_batch), else it is made here.
enum { SLICE_BG = 0, SLICE_HSP = 1, SLICE_EXC = 2 };
struct DeferredCopy { unsigned char* dst; const unsigned char* src; size_t esz; int kind; };

static int upload_bulk(o2a_ctx* ctx, DevBuf& dst, const void* src, size_t n, size_t esz, int kind,
                       std::vector<DeferredCopy>* deferred, const void** out)
{
    int r = ensure(ctx, dst, n * esz);
    if (r) return r;
    if (n) {
        if (deferred) deferred->push_back({(unsigned char*)dst.p, (const unsigned char*)src, esz, kind});
        else CU(cudaMemcpyAsync(dst.p, src, n * esz, cudaMemcpyHostToDevice, ctx->stream));
    }
    *out = dst.p;
    return 0;
}

static int grid_for(o2a_ctx* ctx, long long items, int per_sm)
{
    long long g = (long long)ctx->sm_count * per_sm;
    if (items < g) g = items;
    if (g < 1) g = 1;
    return (int)g;
}

// Grid of a persistent (grid-stride) kernel with items of EQUAL cost: exactly the CTAs the device holds at once, so no
// partly filled last wave is left. The occupancy query is cached per kernel.
template <typename K>
static int grid_resident(o2a_ctx* ctx, K kernel, int block, long long items)
{
    static std::map<const void*, int> cache;               // per process: the device kind is the same for every ctx
    static std::mutex mu;
    int per_sm = 0;
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = cache.find((const void*)kernel);
        if (it != cache.end()) per_sm = it->second;
        else {
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, 0) != cudaSuccess || per_sm < 1) { cudaGetLastError(); per_sm = 4; }
            cache[(const void*)kernel] = per_sm;
        }
    }
    return grid_for(ctx, items, per_sm);
}

static int run_fit(o2a_ctx* ctx)
{
    const o2a_batch& b = ctx->b;
    const o2a_params& p = ctx->prm;
    ENS(thr, b.n_lnc * sizeof(double));
    ENS(status, b.n_lnc * sizeof(int));
    const long long l0 = ctx->vl0, nl = ctx->vl1 - ctx->vl0;
    if (nl <= 0) return 0;
    FitArgs A{};
    A.n_lnc = nl; A.bg_off = (const long long*)b.bg_off + l0; A.bg = b.bg;
    A.kde_factor = b.kde_factor ? b.kde_factor + l0 : nullptr;
    A.fitter = p.fitter; A.alpha = p.alpha;
    if (p.fitter == O2A_FIT_HIST) {
        ENS(fits, b.n_lnc * sizeof(LncFit));
        ENS(tbl_k, (b.n_bg / 2 + 2) * sizeof(int));
        ENS(tbl_cum, (b.n_bg / 2 + 2) * sizeof(double));
    } else {
        if (p.fitter == O2A_FIT_KDE && !b.kde_factor) { ctx->err = "kde_factor is required for the KDE fitter"; return O2A_EINVAL; }
        ENS(uv_val, (b.n_bg + 1) * sizeof(int));
        ENS(uv_cnt, (b.n_bg + 1) * sizeof(int));
        ENS(uv_n, b.n_lnc * sizeof(int));
    }
    int grid = grid_for(ctx, nl, 12);
    // per-CTA global counters for backgrounds whose value range exceeds the shared-memory budget
    long long max_range = UV_MAX_RANGE;
    ENS(fit_cnt, (size_t)grid_for(ctx, b.n_lnc, 12) * max_range * sizeof(int));
    // per-lncRNA arrays are indexed relative to the view; tables / uv slots by absolute bg offsets
    A.thr = P<double>(ctx->thr) + l0; A.fits = P<LncFit>(ctx->fits) + l0; A.tbl_k = P<int>(ctx->tbl_k);
    A.tbl_cum = P<double>(ctx->tbl_cum); A.uv_val = P<int>(ctx->uv_val); A.uv_cnt = P<int>(ctx->uv_cnt);
    A.uv_n = P<int>(ctx->uv_n) + l0; A.status = P<int>(ctx->status) + l0; A.g_cnt = P<int>(ctx->fit_cnt);
    A.max_range = max_range;
    if (p.fitter == O2A_FIT_HIST) {
        // fast path: count distinct values (CTA per lncRNA), exact numpy/scipy arithmetic (warp per lncRNA);
        // what it cannot take is listed and goes through the general kernel
        ENS(fc_val, (size_t)b.n_lnc * FITW_CAP * sizeof(int)); ENS(fc_cnt, (size_t)b.n_lnc * FITW_CAP * sizeof(int));
        ENS(fc_n, b.n_lnc * sizeof(int)); ENS(fit_list, (b.n_lnc + 1) * sizeof(int)); ENS(fit_list_n, sizeof(int));
        CU(cudaMemsetAsync(ctx->fit_list_n.p, 0, sizeof(int), ctx->stream));
        A.fc_val = P<int>(ctx->fc_val) + l0 * FITW_CAP; A.fc_cnt = P<int>(ctx->fc_cnt) + l0 * FITW_CAP;
        A.fc_n = P<int>(ctx->fc_n) + l0;
        A.list = P<int>(ctx->fit_list); A.list_n = P<int>(ctx->fit_list_n);
        int fit_per_sm = 10;
        if (const char* e = getenv("O2A_FITC_PER_SM")) { if (atoi(e) > 0) fit_per_sm = atoi(e); }
        const int gc = grid_for(ctx, nl, fit_per_sm);
        if (b.bg_bits == 8) k_fit_count<signed char><<<gc, FITC_THREADS, 0, ctx->stream>>>(A);
        else if (b.bg_bits == 16) k_fit_count<short><<<gc, FITC_THREADS, 0, ctx->stream>>>(A);
        else k_fit_count<int><<<gc, FITC_THREADS, 0, ctx->stream>>>(A);
        LAUNCH_CHECK();
        k_fit_hist_tail<<<grid_resident(ctx, k_fit_hist_tail, FITW_WARPS * 32, (nl + FITW_WARPS - 1) / FITW_WARPS), FITW_WARPS * 32, 0, ctx->stream>>>(A);
        LAUNCH_CHECK();
    }
    // what this launch cannot take without per-CTA scratch (value range beyond the counters, more distinct values
    // than the shared staging) is listed for a second launch on a few CTAs that own sort / table scratch in HBM
    ENS(fit_list2, (b.n_lnc + 1) * sizeof(int)); ENS(fit_list2_n, sizeof(int));
    CU(cudaMemsetAsync(ctx->fit_list2_n.p, 0, sizeof(int), ctx->stream));
    A.list2 = P<int>(ctx->fit_list2); A.list2_n = P<int>(ctx->fit_list2_n);
    A.scratch = nullptr; A.scratch_stride = 0; A.scratch_cap = 0;
    if (b.bg_bits == 8) k_fit<signed char><<<grid, FIT_THREADS, 0, ctx->stream>>>(A);
    else if (b.bg_bits == 16) k_fit<short><<<grid, FIT_THREADS, 0, ctx->stream>>>(A);
    else k_fit<int><<<grid, FIT_THREADS, 0, ctx->stream>>>(A);
    LAUNCH_CHECK();
    {
        const int g2 = grid_for(ctx, nl, 1) < 32 ? grid_for(ctx, nl, 1) : 32;
        const long long cap = b.max_lnc_bg > 0 ? b.max_lnc_bg : 1;
        const long long stride = ((cap * FIT_SCRATCH_BYTES_PER_ENTRY + 255) / 256) * 256;
        ENS(fit_scratch, (size_t)g2 * stride);
        A.list = A.list2; A.list_n = A.list2_n; A.list2 = nullptr; A.list2_n = nullptr;
        A.scratch = P<unsigned char>(ctx->fit_scratch); A.scratch_stride = stride; A.scratch_cap = cap;
        if (b.bg_bits == 8) k_fit<signed char><<<g2, FIT_THREADS, 0, ctx->stream>>>(A);
        else if (b.bg_bits == 16) k_fit<short><<<g2, FIT_THREADS, 0, ctx->stream>>>(A);
        else k_fit<int><<<g2, FIT_THREADS, 0, ctx->stream>>>(A);
        LAUNCH_CHECK();
    }
    return 0;
}

static int ensure_sort_scratch(o2a_ctx* ctx, int grid, long long cap)
{
    ENS(sort_k0, (size_t)grid * cap * 8); ENS(sort_k1, (size_t)grid * cap * 8);
    ENS(sort_v0, (size_t)grid * cap * 4); ENS(sort_v1, (size_t)grid * cap * 4);
    return 0;
}

// per-chunk counters inside stage_cnt (see o2a_ctx)
enum { CNT_PQ = 0, CNT_RANKBIG = 1, CNT_MID = 2, CNT_CBIG = 3, CNT_FRONT = 4, CNT_PER_CHUNK = 5 };
static int* chunk_cnt(o2a_ctx* ctx, int chunk, int which) { return P<int>(ctx->stage_cnt) + chunk * CNT_PER_CHUNK + which; }

static void fill_chain_args(o2a_ctx* ctx, int chunk, ChainArgs& A);

// Front of the path for the view's alignments, on stream `st`: score filter + p / q / keep (k_front: one warp per
// alignment), then the alignments it queued (k_pq, k_rank_big).
static int run_score(o2a_ctx* ctx, int chunk, cudaStream_t st)
{
    const o2a_batch& b = ctx->b;
    const o2a_params& p = ctx->prm;
    // per-alignment arrays are indexed relative to the view; everything per HSP / per lncRNA absolutely
    const long long a0 = ctx->va0, na = ctx->va1 - ctx->va0;
    if (na <= 0) return stage_event(ctx, O2A_STAGE_PQ);
    const long long* hsp_off_v = (const long long*)b.hsp_off + a0;
    const int* aln_lnc_v = P<int>(ctx->aln_lnc) + a0;
    FrontArgs F{};
    PqArgs& A = F.P;
    A.n_aln = na; A.hsp_off = hsp_off_v; A.aln_lnc = aln_lnc_v;
    A.lnc_status = P<int>(ctx->status); A.n_surv = P<int>(ctx->n_surv) + a0; A.thr = P<double>(ctx->thr);
    A.fits = P<LncFit>(ctx->fits); A.tbl_k = P<int>(ctx->tbl_k); A.tbl_cum = P<double>(ctx->tbl_cum);
    A.bg_off = (const long long*)b.bg_off;
    A.uv_val = P<int>(ctx->uv_val); A.uv_cnt = P<int>(ctx->uv_cnt); A.uv_n = P<int>(ctx->uv_n);
    A.kde_factor = b.kde_factor; A.fitter = p.fitter; A.fdr = p.fdr; A.alpha = p.alpha;
    A.surv_x = P<double>(ctx->surv_x); A.surv_p = P<double>(ctx->surv_p); A.surv_pq = P<double>(ctx->surv_pq);
    A.surv_keep = P<unsigned char>(ctx->surv_keep); A.n_keep = P<int>(ctx->n_keep) + a0;
    A.pq_list = P<int>(ctx->pq_list) + a0; A.pq_count = chunk_cnt(ctx, chunk, CNT_PQ);
    A.big_list = P<int>(ctx->big_list) + a0; A.big_count = chunk_cnt(ctx, chunk, CNT_RANKBIG);
    F.score = b.score; F.score_q = b.score_q; F.exc_off = b.exc_off ? (const long long*)b.exc_off + a0 : nullptr;
    F.exc_pos = b.exc_pos; F.exc_val = b.exc_val; F.n_hsp_total = b.n_hsp;
    F.n_surv = P<int>(ctx->n_surv) + a0; F.surv_idx = P<int>(ctx->surv_idx); F.surv_x = P<double>(ctx->surv_x);
    F.pq_list = P<int>(ctx->pq_list) + a0; F.pq_count = chunk_cnt(ctx, chunk, CNT_PQ);
    {
        ChainArgs C{};
        fill_chain_args(ctx, chunk, C);
        F.ret_slot = C.ret_slot; F.ret_score = C.ret_score; F.ret_coord = C.ret_coord;
        F.fetch_coords = ctx->coords_on_device ? 1 : 0;
        F.coords4 = C.coords4; F.qstart = C.qstart; F.qend = C.qend; F.sstart = C.sstart; F.send = C.send;
        F.qlen = C.qlen; F.slen = C.slen; F.mid_list = C.mid_list; F.mid_count = C.mid_count;
    }
    const int grid = grid_for(ctx, (na + FRONT_WARPS - 1) / FRONT_WARPS, FRONT_CTAS_PER_SM);
    // one ticket = about one lncRNA's alignments (they share the class table k_front keeps on chip), but never so many
    // that the warps get fewer than ~4 tickets each
    F.ticket = chunk_cnt(ctx, chunk, CNT_FRONT);
    {
        const long long nl_view = ctx->vl1 - ctx->vl0 > 0 ? ctx->vl1 - ctx->vl0 : 1;
        long long per_lnc = (na + nl_view / 2) / nl_view;
        const long long balance = na / (4ll * grid * FRONT_WARPS);
        if (per_lnc > balance) per_lnc = balance;
        if (per_lnc > 64) per_lnc = 64;
        if (per_lnc < 1) per_lnc = 1;
        F.chunk_aln = (int)per_lnc;
        if (const char* e = getenv("O2A_FRONT_CHUNK")) { if (atoi(e) > 0) F.chunk_aln = atoi(e); }
    }
#define O2A_FRONT(T) do { if (p.fitter == O2A_FIT_HIST) k_front<T, 0><<<grid, FRONT_WARPS * 32, 0, st>>>(F); \
                         else k_front<T, 1><<<grid, FRONT_WARPS * 32, 0, st>>>(F); } while (0)
    if (b.score_bits == 8) O2A_FRONT(signed char);
    else if (b.score_bits == 16) O2A_FRONT(short);
    else if (b.score_bits == 32) O2A_FRONT(int);
    else O2A_FRONT(double);
#undef O2A_FRONT
    LAUNCH_CHECK();
    if (ctx->timing) CU(cudaEventRecord(ctx->ev[O2A_STAGE_PQ], st));
    // what k_front queued: more than FRONT_SCAP survivors, or many non-integral scores (a small grid: the list is
    // empty for BLAST-like data, and a CTA of an empty list exits at once)
    k_pq<64, 14, RANK_SMALL_MAX, 256, 128, TBL_SMEM_MAX><<<grid_for(ctx, na, b.max_aln_hsps > 4 * FRONT_SCAP ? 14 : 2), 64, 0, st>>>(A);
    LAUNCH_CHECK();
    if (p.fdr && b.max_aln_hsps > RANK_SMALL_MAX) {
        int g2 = grid_for(ctx, b.n_aln, 2);
        int r = ensure_sort_scratch(ctx, g2, b.max_aln_hsps);
        if (r) return r;
        k_rank_big<<<g2, RANKBIG_THREADS, 0, st>>>(
            A.big_list, A.big_count, hsp_off_v, p.alpha, P<int>(ctx->n_surv) + a0,
            P<int>(ctx->n_keep) + a0, P<double>(ctx->surv_p), P<double>(ctx->surv_pq), P<unsigned char>(ctx->surv_keep),
            P<unsigned long long>(ctx->sort_k0), P<unsigned long long>(ctx->sort_k1), P<unsigned int>(ctx->sort_v0),
            P<unsigned int>(ctx->sort_v1), b.max_aln_hsps);
        LAUNCH_CHECK();
    }
    return 0;
}

static void fill_chain_args(o2a_ctx* ctx, int chunk, ChainArgs& A)
{
    const o2a_batch& b = ctx->b;
    const o2a_params& p = ctx->prm;
    const long long a0 = ctx->va0, na = ctx->va1 - ctx->va0;
    A.n_aln = na; A.hsp_off = (const long long*)b.hsp_off + a0; A.aln_lnc = P<int>(ctx->aln_lnc) + a0;
    A.qlen = (const long long*)b.qlen + a0; A.slen = (const long long*)b.slen + a0;
    A.qstart = b.qstart; A.qend = b.qend; A.sstart = b.sstart; A.send = b.send; A.score = b.score;
    A.coords4 = ctx->coords4;
    A.ret_coord = P<int4>(ctx->ret_coord); A.ret_score = P<double>(ctx->ret_score); A.ret_slot = P<int>(ctx->ret_slot);
    A.n_surv = P<int>(ctx->n_surv) + a0; A.n_keep = P<int>(ctx->n_keep) + a0; A.surv_idx = P<int>(ctx->surv_idx);
    A.surv_keep = P<unsigned char>(ctx->surv_keep); A.surv_x = P<double>(ctx->surv_x);
    A.gapopen = p.gapopen; A.gapextend = p.gapextend;
    A.best_value = p.best_value;
    A.chain_len = P<int>(ctx->chain_len) + a0; A.chain_orient = P<unsigned char>(ctx->chain_orient) + a0;
    A.chain_score = P<double>(ctx->chain_score) + a0; A.orient_scores = P<double>(ctx->orient_scores) + 2 * a0;
    A.chain_slot = P<int>(ctx->chain_slot); A.best_key = P<double>(ctx->best_key) + a0;
    A.lnc_status = P<int>(ctx->status);
    A.mid_list = P<int>(ctx->mid_list) + a0; A.mid_count = chunk_cnt(ctx, chunk, CNT_MID);
    A.big_list = P<int>(ctx->cbig_list) + a0; A.big_count = chunk_cnt(ctx, chunk, CNT_CBIG);
    // what k_front left to k_pq still needs its retained list (k_retained_slots) and, on the device, its coordinates
    A.work_list = P<int>(ctx->pq_list) + a0; A.work_count = chunk_cnt(ctx, chunk, CNT_PQ);
}

static int ensure_chain_buffers(o2a_ctx* ctx)
{
    const o2a_batch& b = ctx->b;
    ENS(chain_len, b.n_aln * sizeof(int)); ENS(chain_orient, b.n_aln); ENS(chain_score, b.n_aln * sizeof(double));
    ENS(orient_scores, 2 * b.n_aln * sizeof(double)); ENS(chain_slot, b.n_hsp * sizeof(int));
    ENS(best_key, b.n_aln * sizeof(double)); ENS(best_aln, b.n_lnc * sizeof(long long));
    ENS(ret_coord, b.n_hsp * sizeof(int4)); ENS(ret_score, b.n_hsp * sizeof(double)); ENS(ret_slot, b.n_hsp * sizeof(int));
    return 0;
}

// What k_front left open for the chaining: the retained lists of the alignments that went through k_pq, and the
// coordinates of the retained HSPs wherever k_front could not read them itself (`all` = every alignment: the operator
// seam o2a_best_chain, which does not run k_front).
static int run_gather(o2a_ctx* ctx, int chunk, cudaStream_t st, bool all = false)
{
    const long long na = ctx->va1 - ctx->va0;
    if (na <= 0) return 0;
    ChainArgs A{};
    fill_chain_args(ctx, chunk, A);
    if (all) { A.work_list = nullptr; A.work_count = nullptr; }
    const int small_grid = grid_for(ctx, (na + 7) / 8, 2);
    k_retained_slots<<<all ? grid_for(ctx, (na + 7) / 8, 8) : small_grid, 256, 0, st>>>(A);
    LAUNCH_CHECK();
    if (ctx->coords_mode_used == O2A_COORDS_HOST_GATHER && !all) return 0;     // the host supplies the coordinates
    if (ctx->coords_on_device && !all) {
        k_fetch_coords<<<small_grid, 256, 0, st>>>(A);                          // only what went through k_pq
        LAUNCH_CHECK();
        return 0;
    }
    // Coordinates read in place over PCIe: the SMs' 32-byte reads and the copy engine's reads of the next chunk
    // compete for the link's read-request slots, and a machine-filling gather starves the DMA (8 CTAs per SM:
    // 13.0 ms per C2 batch, 1 per SM: 12.6 ms). A small grid keeps enough requests in flight for the gather
    // itself (it is request-latency bound either way) and leaves the rest to the uploads. Measured per C2 batch, one
    // context / two contexts alternating: 1184 CTAs 13.0 / 12.07 ms, 148: 12.57 / 11.53, 74: 12.04 / 11.14, 32: 12.51 / 10.92.
    int gather_grid = grid_for(ctx, (na + 7) / 8, 8);
    if (ctx->coords_zero_copy && ctx->pipelined_chunks > 1) {      // only then do uploads run beside the gather
        int want = ctx->sm_count / 2;
        if (const char* e = getenv("O2A_GATHER_CTAS")) { if (atoi(e) > 0) want = atoi(e); }
        if (want < gather_grid) gather_grid = want;
    }
    A.work_list = nullptr; A.work_count = nullptr;
    k_fetch_coords<<<gather_grid, 256, 0, st>>>(A);
    LAUNCH_CHECK();
    return 0;
}

// Chaining of the view's alignments + best ortholog of its lncRNAs, on stream `st` (ret_* filled by the gather)
static int run_chain(o2a_ctx* ctx, int chunk, cudaStream_t st)
{
    const o2a_batch& b = ctx->b;
    const o2a_params& p = ctx->prm;
    const long long l0 = ctx->vl0, nl = ctx->vl1 - ctx->vl0, na = ctx->va1 - ctx->va0;
    if (ctx->timing) CU(cudaEventRecord(ctx->ev[O2A_STAGE_CHAIN], st));
    if (na > 0) {
        ChainArgs A{};
        fill_chain_args(ctx, chunk, A);
        // The few alignments beyond the warp path (queued by the gather) are chained beside it: k_chain_small goes
        // first on `st` (a grid of mostly idle CTAs around ~10^2 long serial ones) and the machine-filling
        // k_chain_warp follows on the second stream as soon as the gather is done, not after k_chain_small.
        const bool fork = b.max_aln_hsps > CHAIN_WARP_MAX;
        if (fork) {
            CU(cudaEventRecord(ctx->aux_fork, st));
            CU(cudaStreamWaitEvent(ctx->aux_stream, ctx->aux_fork, 0));
            k_chain_small<<<grid_for(ctx, na, 12), CHAIN_THREADS, 0, st>>>(A);
            LAUNCH_CHECK();
        }
        // more CTAs than the device holds at once (7 per SM at 72 registers), on purpose: the alignments' costs differ,
        // and the block scheduler hands the queued CTAs to whichever SM frees a slot first (exactly one resident wave:
        // 0.283 -> 0.323 ms on BASELINE config 2; 16 per SM 0.283, 32: 0.273, 64: 0.266)
        int chain_per_sm = 64;
        if (const char* e = getenv("O2A_CHAINW_PER_SM")) { if (atoi(e) > 0) chain_per_sm = atoi(e); }
        k_chain_warp<<<grid_for(ctx, (na + CHAINW_WARPS - 1) / CHAINW_WARPS, chain_per_sm), CHAINW_WARPS * 32, 0,
                       fork ? ctx->aux_stream : st>>>(A);
        LAUNCH_CHECK();
        if (fork) {
            CU(cudaEventRecord(ctx->aux_join, ctx->aux_stream));
            CU(cudaStreamWaitEvent(st, ctx->aux_join, 0));
        }
        if (b.max_aln_hsps > CHAIN_SMALL_MAX) {
            int g2 = grid_for(ctx, b.n_aln, 2);
            const long long cap = b.max_aln_hsps;
            ENS(csort_k0, (size_t)g2 * cap * 8); ENS(csort_k1, (size_t)g2 * cap * 8);
            ENS(csort_v0, (size_t)g2 * cap * 4); ENS(csort_v1, (size_t)g2 * cap * 4);
            ENS(cb_qs, (size_t)g2 * cap * 4); ENS(cb_qe, (size_t)g2 * cap * 4); ENS(cb_ss, (size_t)g2 * cap * 4);
            ENS(cb_se, (size_t)g2 * cap * 4); ENS(cb_sc, (size_t)g2 * cap * 8); ENS(cb_id, (size_t)g2 * cap * 4);
            ENS(cb_total, (size_t)g2 * (cap + 2) * 8); ENS(cb_prev, (size_t)g2 * (cap + 2) * 4);
            ENS(cb_f, (size_t)g2 * (cap + 2) * 4); ENS(cb_chain, (size_t)g2 * cap * 2 * 4);
            ChainBigScratch S{P<unsigned long long>(ctx->csort_k0), P<unsigned long long>(ctx->csort_k1),
                              P<unsigned int>(ctx->csort_v0), P<unsigned int>(ctx->csort_v1),
                              P<int>(ctx->cb_qs), P<int>(ctx->cb_qe), P<int>(ctx->cb_ss), P<int>(ctx->cb_se),
                              P<double>(ctx->cb_sc), P<int>(ctx->cb_id), P<double>(ctx->cb_total), P<int>(ctx->cb_prev),
                              P<int>(ctx->cb_f), P<int>(ctx->cb_chain), cap};
            k_chain_big<<<g2, CHAINBIG_THREADS, 0, st>>>(A, S);
            LAUNCH_CHECK();
        }
    }
    if (ctx->timing) CU(cudaEventRecord(ctx->ev[O2A_STAGE_BEST], st));
    if (nl > 0) {
        // aln_off values are absolute alignment indices: per-alignment arrays are passed unshifted here
        k_best<<<(unsigned)((nl + 127) / 128), 128, 0, st>>>(
            nl, (const long long*)b.aln_off + l0, P<int>(ctx->status) + l0, P<int>(ctx->n_surv), P<int>(ctx->n_keep),
            P<int>(ctx->chain_len), P<double>(ctx->chain_score), P<double>(ctx->orient_scores),
            P<double>(ctx->best_key), p.best_function, P<long long>(ctx->best_aln) + l0);
        LAUNCH_CHECK();
    }
    return 0;
}

static int ensure_pinned(o2a_ctx* ctx, void** p, size_t* cap, size_t bytes)
{
    if (*cap >= bytes && *p) return 0;
    if (*p) { cudaFreeHost(*p); *p = nullptr; *cap = 0; }
    const size_t want = bytes + bytes / 4 + 4096;
    if (cudaHostAlloc(p, want, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
        cudaGetLastError();
        ctx->err = "cudaHostAlloc(" + std::to_string(want) + ") failed";
        return O2A_ENOMEM;
    }
    *cap = want;
    return 0;
}

static int find_max_diff(o2a_ctx* ctx, long long n, const long long* off_dev, long long* out)
{
    ENS(maxred, 8);
    CU(cudaMemsetAsync(ctx->maxred.p, 0, 8, ctx->stream));
    if (n > 0) {
        k_max_diff<<<grid_for(ctx, (n + 255) / 256, 4), 256, 0, ctx->stream>>>(n, off_dev, P<unsigned long long>(ctx->maxred));
        LAUNCH_CHECK();
    }
    unsigned long long v = 0;
    CU(cudaMemcpyAsync(&v, ctx->maxred.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    *out = (long long)v;
    return 0;
}

extern "C" int o2a_build_batch(o2a_ctx* ctx, const o2a_batch* batch, const o2a_params* params, o2a_counts* counts)
{
    if (!ctx) return O2A_EINVAL;
    if (!batch || !params) { ctx->err = "null batch/params"; return O2A_EINVAL; }
    if (batch->n_lnc < 0 || batch->n_aln < 0 || batch->n_hsp < 0 || batch->n_bg < 0) { ctx->err = "negative size"; return O2A_EINVAL; }
    if (params->fitter < 0 || params->fitter > 2) { ctx->err = "unknown fitter"; return O2A_EINVAL; }
    if (batch->n_hsp > 0x7fffffffll * 64) { ctx->err = "batch too large"; return O2A_EINVAL; }
    CU(cudaSetDevice(ctx->device));
    ctx->have_result = false;
    ctx->ev_valid = false;
    ctx->prm = *params;
    o2a_batch& b = ctx->b;
    b = *batch;
    int r;
    if ((r = stage_event(ctx, O2A_STAGE_H2D))) return r;
    // Host batches large enough to amortise the extra launches are processed as a few lncRNA chunks:
    // the bulk arrays of chunk c+1 cross PCIe on the copy stream while the kernels of chunk c run.
    std::vector<DeferredCopy> deferred_copies;
    int n_chunks = 1;
    long long pipeline_min_hsps = 4ll << 20;
    if (const char* e = getenv("O2A_PIPELINE_MIN_HSPS")) pipeline_min_hsps = atoll(e);
    if (batch->mem == O2A_MEM_HOST && batch->n_lnc >= 8 && batch->n_hsp >= pipeline_min_hsps && !getenv("O2A_NO_PIPELINE")) {
        n_chunks = 8;
        // every chunk must still fill the machine: kernels that take one CTA per (large) alignment would otherwise
        // run chunk after chunk at a fraction of the SMs (600 x 50 k-HSP alignments: 8 chunks tripled the step)
        const long long by_alignments = batch->n_aln / 2048;
        if (by_alignments < n_chunks) n_chunks = (int)(by_alignments < 1 ? 1 : by_alignments);
        if (const char* e = getenv("O2A_CHUNKS")) n_chunks = atoi(e);
        if (n_chunks < 1) n_chunks = 1;
        if (n_chunks > 16) n_chunks = 16;
    }
    std::vector<DeferredCopy>* deferred = n_chunks > 1 ? &deferred_copies : nullptr;
    ctx->pipelined_chunks = n_chunks;
    int coords_mode = O2A_COORDS_AUTO;           // how the retained HSPs' coordinate records reach the device
    if (batch->mem == O2A_MEM_HOST) {
        // the reference pickles file names to workers which json.load them (pipeline.py:743-754);
        // here the columnar arrays cross PCIe once
        if ((r = upload(ctx, ctx->in_bg_off, batch->bg_off, (size_t)batch->n_lnc + 1, &b.bg_off))) return r;
        if (batch->bg_bits != 0 && batch->bg_bits != 8 && batch->bg_bits != 16 && batch->bg_bits != 32) { ctx->err = "bg_bits must be 0, 8, 16 or 32"; return O2A_EINVAL; }
        {
            const void* d8 = nullptr;
            const size_t esz = batch->bg_bits == 8 ? 1 : batch->bg_bits == 16 ? 2 : 4;
            if ((r = upload_bulk(ctx, ctx->in_bg, batch->bg, (size_t)batch->n_bg, esz, SLICE_BG, deferred, &d8))) return r;
            b.bg = (const int32_t*)d8;
        }
        if (batch->kde_factor && params->fitter == O2A_FIT_KDE) {
            if ((r = upload(ctx, ctx->in_kde, batch->kde_factor, (size_t)batch->n_lnc, &b.kde_factor))) return r;
        } else b.kde_factor = nullptr;
        if ((r = upload(ctx, ctx->in_aln_off, batch->aln_off, (size_t)batch->n_lnc + 1, &b.aln_off))) return r;
        if ((r = upload(ctx, ctx->in_hsp_off, batch->hsp_off, (size_t)batch->n_aln + 1, &b.hsp_off))) return r;
        if ((r = upload(ctx, ctx->in_qlen, batch->qlen, (size_t)batch->n_aln, &b.qlen))) return r;
        if ((r = upload(ctx, ctx->in_slen, batch->slen, (size_t)batch->n_aln, &b.slen))) return r;
        if (batch->score_bits == 8 || batch->score_bits == 16 || batch->score_bits == 32) {
            if (!batch->score_q || !batch->exc_off) { ctx->err = "score_bits set but score_q/exc_off missing"; return O2A_EINVAL; }
            const size_t esz = batch->score_bits / 8;
            const void* d8 = nullptr;
            if ((r = upload_bulk(ctx, ctx->in_score_q, batch->score_q, (size_t)batch->n_hsp, esz, SLICE_HSP, deferred, &d8))) return r;
            b.score_q = d8;
            if ((r = upload(ctx, ctx->in_exc_off, batch->exc_off, (size_t)batch->n_aln + 1, &b.exc_off))) return r;
            const size_t n_exc = batch->n_aln > 0 ? (size_t)batch->exc_off[batch->n_aln] : 0;
            if ((r = upload_bulk(ctx, ctx->in_exc_pos, batch->exc_pos, n_exc, sizeof(int32_t), SLICE_EXC, deferred, &d8))) return r;
            b.exc_pos = (const int32_t*)d8;
            if ((r = upload_bulk(ctx, ctx->in_exc_val, batch->exc_val, n_exc, sizeof(double), SLICE_EXC, deferred, &d8))) return r;
            b.exc_val = (const double*)d8;
            b.score = nullptr;
        } else if (batch->score_bits != 0) { ctx->err = "score_bits must be 0, 8, 16 or 32"; return O2A_EINVAL; }
        else {
            const void* d8 = nullptr;
            if ((r = upload_bulk(ctx, ctx->in_score, batch->score, (size_t)batch->n_hsp, sizeof(double), SLICE_HSP, deferred, &d8))) return r;
            b.score = (const double*)d8;
        }
        // Coordinates are only ever read for the HSPs that reach the chaining (~1 %). Three ways to get those records
        // to the device (o2a_set_coords_mode / env O2A_COORDS=zerocopy|gather|upload; AUTO = host gather):
        //   zero-copy   the gather kernel reads them in place over PCIe (32-byte reads)
        //   host gather the GPU lists the retained HSP indices into pinned host memory, host threads collect the
        //               16-byte records, one small DMA brings them back (works with pageable caller memory)
        //   upload      all 16 B per HSP cross PCIe (what a dense batch, where everything is retained, wants)
        ctx->coords_zero_copy = false;
        ctx->coords4 = nullptr;
        auto mapped = [](const void* hp) -> const void* {
            cudaPointerAttributes at;
            if (cudaPointerGetAttributes(&at, hp) != cudaSuccess || at.type != cudaMemoryTypeHost || !at.devicePointer) {
                cudaGetLastError();
                return nullptr;
            }
            return at.devicePointer;
        };
        int want = ctx->coords_mode_req;
        if (const char* e = getenv("O2A_COORDS")) {
            if (!strcmp(e, "zerocopy")) want = O2A_COORDS_ZERO_COPY;
            else if (!strcmp(e, "gather")) want = O2A_COORDS_HOST_GATHER;
            else if (!strcmp(e, "upload")) want = O2A_COORDS_UPLOAD;
        }
        if (getenv("O2A_NO_ZERO_COPY") && want == O2A_COORDS_ZERO_COPY) want = O2A_COORDS_HOST_GATHER;
        const void* dp[4] = {nullptr, nullptr, nullptr, nullptr};
        bool can_map = false;
        if (batch->n_hsp > 0 && want == O2A_COORDS_ZERO_COPY) {
            if (batch->coords) { dp[0] = mapped(batch->coords); can_map = dp[0] != nullptr; }
            else {
                dp[0] = mapped(batch->qstart); dp[1] = mapped(batch->qend); dp[2] = mapped(batch->sstart); dp[3] = mapped(batch->send);
                can_map = dp[0] && dp[1] && dp[2] && dp[3];
            }
        }
        // AUTO = host gather: measured on a B200 box it is as fast or faster than zero-copy reads (C2 batch, one context:
        // 11.85 vs 12.29 ms; C5: 4.82 vs 5.61 ms), and it does not need the caller's coordinate arrays to be pinned
        if (want == O2A_COORDS_UPLOAD || batch->n_hsp == 0) coords_mode = O2A_COORDS_UPLOAD;
        else if (want == O2A_COORDS_ZERO_COPY && can_map) coords_mode = O2A_COORDS_ZERO_COPY;
        else coords_mode = O2A_COORDS_HOST_GATHER;
        if (coords_mode == O2A_COORDS_ZERO_COPY) {
            ctx->coords_zero_copy = true;
            if (batch->coords) { ctx->coords4 = (const int4*)dp[0]; b.qstart = b.qend = b.sstart = b.send = nullptr; }
            else { b.qstart = (const int*)dp[0]; b.qend = (const int*)dp[1]; b.sstart = (const int*)dp[2]; b.send = (const int*)dp[3]; }
        } else if (coords_mode == O2A_COORDS_UPLOAD) {
            const void* d = nullptr;
            if (batch->coords) {
                if ((r = upload_bulk(ctx, ctx->in_coords, batch->coords, (size_t)batch->n_hsp, 16, SLICE_HSP, deferred, &d))) return r;
                ctx->coords4 = (const int4*)d;
                b.qstart = b.qend = b.sstart = b.send = nullptr;
            } else {
                if ((r = upload_bulk(ctx, ctx->in_qs, batch->qstart, (size_t)batch->n_hsp, 4, SLICE_HSP, deferred, &d))) return r;
                b.qstart = (const int32_t*)d;
                if ((r = upload_bulk(ctx, ctx->in_qe, batch->qend, (size_t)batch->n_hsp, 4, SLICE_HSP, deferred, &d))) return r;
                b.qend = (const int32_t*)d;
                if ((r = upload_bulk(ctx, ctx->in_ss, batch->sstart, (size_t)batch->n_hsp, 4, SLICE_HSP, deferred, &d))) return r;
                b.sstart = (const int32_t*)d;
                if ((r = upload_bulk(ctx, ctx->in_se, batch->send, (size_t)batch->n_hsp, 4, SLICE_HSP, deferred, &d))) return r;
                b.send = (const int32_t*)d;
            }
        } else {
            b.qstart = b.qend = b.sstart = b.send = nullptr;       // never dereferenced on the device in this mode
        }
        {   // the caller's maxima are hints: a host batch is checked here (two passes over the offset arrays)
            long long m = 0;
            for (long long a = 0; a < batch->n_aln; ++a) { long long d = batch->hsp_off[a + 1] - batch->hsp_off[a]; if (d > m) m = d; }
            if (b.max_aln_hsps > 0 && b.max_aln_hsps < m) { ctx->err = "max_aln_hsps is smaller than the largest alignment"; return O2A_EINVAL; }
            b.max_aln_hsps = m;
            m = 0;
            for (long long l = 0; l < batch->n_lnc; ++l) { long long d = batch->bg_off[l + 1] - batch->bg_off[l]; if (d > m) m = d; }
            if (b.max_lnc_bg > 0 && b.max_lnc_bg < m) { ctx->err = "max_lnc_bg is smaller than the largest background"; return O2A_EINVAL; }
            b.max_lnc_bg = m;
        }
    } else {
        ctx->coords4 = (const int4*)batch->coords;
        ctx->coords_zero_copy = false;
        if (b.max_aln_hsps <= 0 && (r = find_max_diff(ctx, b.n_aln, (const long long*)b.hsp_off, (long long*)&b.max_aln_hsps))) return r;
        if (b.max_lnc_bg <= 0 && (r = find_max_diff(ctx, b.n_lnc, (const long long*)b.bg_off, (long long*)&b.max_lnc_bg))) return r;
    }
    ctx->coords_mode_used = coords_mode;
    ctx->coords_on_device = batch->mem != O2A_MEM_HOST || coords_mode == O2A_COORDS_UPLOAD;

    // ---- everything below up to the result counters is stream work without a host decision in between. For a
    // device-resident batch it is captured into a CUDA graph on the second call with identical arguments and replayed
    // from then on (17 launches of BASELINE config 2 -> one graph launch; what it saves is the launch gaps, which is
    // most of the step for small shards: config 1, config 5 sharded over 8 GPUs).
    std::vector<long long> key;
    bool replay = false, capture = false;
    ctx->graph_state = 0;
    if (batch->mem != O2A_MEM_HOST && !ctx->timing && !ctx->graph_disabled && !getenv("O2A_NO_GRAPH") &&
        batch->max_aln_hsps > 0 && batch->max_lnc_bg > 0) {
        const long long fields[] = {
            batch->n_lnc, batch->n_aln, batch->n_hsp, batch->n_bg, (long long)batch->bg_off, (long long)batch->bg,
            (long long)batch->kde_factor, (long long)batch->aln_off, (long long)batch->hsp_off, (long long)batch->qlen,
            (long long)batch->slen, (long long)batch->qstart, (long long)batch->qend, (long long)batch->sstart,
            (long long)batch->send, (long long)batch->score, batch->mem, batch->max_aln_hsps, batch->max_lnc_bg,
            (long long)batch->coords, batch->score_bits, batch->bg_bits, (long long)batch->score_q, (long long)batch->exc_off,
            (long long)batch->exc_pos, (long long)batch->exc_val, params->fitter, params->fdr,
            __builtin_bit_cast(long long, params->alpha), params->gapopen, params->gapextend, params->best_value,
            params->best_function, (long long)ctx->stream, (long long)ctx->alloc_epoch};
        key.assign(fields, fields + sizeof(fields) / sizeof(fields[0]));
        if (ctx->graph_exec && key == ctx->graph_key) replay = true;
        else if (key == ctx->seen_key) capture = true;
    }
    if (!replay && ctx->graph_exec) { cudaGraphExecDestroy(ctx->graph_exec); ctx->graph_exec = nullptr; ctx->graph_key.clear(); }

    auto stream_work = [&]() -> int {
    int r;
    if (batch->mem != O2A_MEM_HOST) {
        // caller-supplied maxima of a device batch are verified on the device, without an extra synchronisation:
        // the true maxima travel back with the result counters and an understated hint fails the call
        ENS(maxchk, 16);
        CU(cudaMemsetAsync(ctx->maxchk.p, 0, 16, ctx->stream));
        if (b.n_aln > 0) { k_max_diff<<<grid_for(ctx, (b.n_aln + 255) / 256, 4), 256, 0, ctx->stream>>>(b.n_aln, (const long long*)b.hsp_off, P<unsigned long long>(ctx->maxchk)); LAUNCH_CHECK(); }
        if (b.n_lnc > 0) { k_max_diff<<<grid_for(ctx, (b.n_lnc + 255) / 256, 4), 256, 0, ctx->stream>>>(b.n_lnc, (const long long*)b.bg_off, P<unsigned long long>(ctx->maxchk) + 1); LAUNCH_CHECK(); }
    }
    ENS(aln_lnc, b.n_aln * sizeof(int));
    if (b.n_lnc > 0) {
        k_fill_aln_lnc<<<grid_for(ctx, b.n_lnc, 16), 64, 0, ctx->stream>>>(b.n_lnc, (const long long*)b.aln_off, P<int>(ctx->aln_lnc));
        LAUNCH_CHECK();
    }
    // workspace of every stage, sized for the whole batch (slot layout: see DESIGN.md)
    ENS(n_surv, b.n_aln * sizeof(int)); ENS(n_keep, b.n_aln * sizeof(int));
    ENS(surv_idx, b.n_hsp * sizeof(int)); ENS(surv_p, b.n_hsp * sizeof(double)); ENS(surv_x, b.n_hsp * sizeof(double));
    ENS(surv_pq, b.n_hsp * sizeof(double)); ENS(surv_keep, b.n_hsp);
    ENS(big_list, (b.n_aln + 1) * sizeof(int)); ENS(mid_list, (b.n_aln + 1) * sizeof(int));
    ENS(pq_list, (b.n_aln + 1) * sizeof(int)); ENS(cbig_list, (b.n_aln + 1) * sizeof(int));
    ENS(stage_cnt, 16 * CNT_PER_CHUNK * sizeof(int));
    CU(cudaMemsetAsync(ctx->stage_cnt.p, 0, 16 * CNT_PER_CHUNK * sizeof(int), ctx->stream));
    if ((r = ensure_chain_buffers(ctx))) return r;

    // lncRNA boundaries that split the HSPs evenly (lncRNAs are independent units of the path)
    long long lb[17];
    lb[0] = 0; lb[n_chunks] = b.n_lnc;
    for (int c = 1; c < n_chunks; ++c) {
        const long long target = batch->n_hsp / n_chunks * c;
        long long lo = lb[c - 1], hi = batch->n_lnc;
        while (lo < hi) {
            long long mid = (lo + hi) >> 1;
            if (batch->hsp_off[batch->aln_off[mid]] < target) lo = mid + 1; else hi = mid;
        }
        lb[c] = lo;
    }
    auto set_view = [&](int c) {
        ctx->vl0 = lb[c]; ctx->vl1 = lb[c + 1];
        if (n_chunks > 1) { ctx->va0 = batch->aln_off[lb[c]]; ctx->va1 = batch->aln_off[lb[c + 1]]; }
        else { ctx->va0 = 0; ctx->va1 = b.n_aln; }
    };
    if (n_chunks > 1) {
        // the copy stream starts after everything already queued on the compute stream
        CU(cudaEventRecord(ctx->copy_ev[0], ctx->stream));
        CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_ev[0], 0));
        for (int c = 0; c < n_chunks; ++c) {
            const long long l0 = lb[c], l1 = lb[c + 1];
            const long long a0 = batch->aln_off[l0], a1 = batch->aln_off[l1];
            const long long lo_of[3] = {batch->bg_off[l0], batch->hsp_off[a0], batch->exc_off ? batch->exc_off[a0] : 0};
            const long long hi_of[3] = {batch->bg_off[l1], batch->hsp_off[a1], batch->exc_off ? batch->exc_off[a1] : 0};
            for (const DeferredCopy& d : deferred_copies) {
                const long long lo = lo_of[d.kind], hi = hi_of[d.kind];
                if (hi > lo)
                    CU(cudaMemcpyAsync(d.dst + lo * d.esz, d.src + lo * d.esz, (size_t)(hi - lo) * d.esz,
                                       cudaMemcpyHostToDevice, ctx->copy_stream));
            }
            CU(cudaEventRecord(ctx->copy_ev[c], ctx->copy_stream));
        }
    }
    if (coords_mode != O2A_COORDS_HOST_GATHER) {
        for (int c = 0; c < n_chunks; ++c) {
            if (n_chunks > 1) CU(cudaStreamWaitEvent(ctx->stream, ctx->copy_ev[c], 0));
            set_view(c);
            // stage events keep the times of the last chunk; "h2d" then spans the earlier chunks too
            if ((r = stage_event(ctx, O2A_STAGE_FIT))) return r;
            if ((r = run_fit(ctx))) return r;
            if ((r = stage_event(ctx, O2A_STAGE_FILTER))) return r;
            if ((r = run_score(ctx, c, ctx->stream))) return r;       // records O2A_STAGE_PQ itself
            if ((r = stage_event(ctx, O2A_STAGE_GATHER))) return r;
            if ((r = run_gather(ctx, c, ctx->stream))) return r;
            if ((r = run_chain(ctx, c, ctx->stream))) return r;       // records O2A_STAGE_CHAIN and O2A_STAGE_BEST itself
        }
    } else {
        // ---- host gather: front of every chunk on the compute stream; the host collects the retained records of chunk
        // c while the front of chunk c+1 runs; scatter + chaining of chunk c on the tail stream ------------------------
        long long cap_of[16], base_of[17];
        base_of[0] = 0;
        for (int c = 0; c < n_chunks; ++c) {
            const long long h_lo = batch->hsp_off[n_chunks > 1 ? batch->aln_off[lb[c]] : 0];
            const long long h_hi = batch->hsp_off[n_chunks > 1 ? batch->aln_off[lb[c + 1]] : batch->n_aln];
            long long cap = (h_hi - h_lo) / 8;                     // ~1 % is retained on BLAST-like data; beyond 12.5 %
            if (cap < 65536) cap = 65536;                          // the chunk's records are uploaded instead
            if (cap > h_hi - h_lo) cap = h_hi - h_lo;
            cap_of[c] = cap; base_of[c + 1] = base_of[c] + cap;
        }
        const size_t tot = (size_t)base_of[n_chunks];
        if ((r = ensure_pinned(ctx, (void**)&ctx->h_list, &ctx->h_list_cap, tot * 8 + 16))) return r;
        if ((r = ensure_pinned(ctx, (void**)&ctx->h_stage, &ctx->h_stage_cap, tot * 16 + 16))) return r;
        if (!ctx->h_cursor) { ctx->err = "pinned cursor missing"; return O2A_ENOMEM; }
        ENS(d_stage, tot * 16 + 16); ENS(ret_base, (b.n_aln + 1) * 8); ENS(list_cursor, 16 * 8);
        CU(cudaMemsetAsync(ctx->list_cursor.p, 0, 16 * 8, ctx->stream));
        long long* d_list = nullptr;
        CU(cudaHostGetDevicePointer((void**)&d_list, ctx->h_list, 0));
        for (int c = 0; c < n_chunks; ++c) {
            if (n_chunks > 1) CU(cudaStreamWaitEvent(ctx->stream, ctx->copy_ev[c], 0));
            set_view(c);
            if ((r = stage_event(ctx, O2A_STAGE_FIT))) return r;
            if ((r = run_fit(ctx))) return r;
            if ((r = stage_event(ctx, O2A_STAGE_FILTER))) return r;
            if ((r = run_score(ctx, c, ctx->stream))) return r;
            if ((r = stage_event(ctx, O2A_STAGE_GATHER))) return r;
            const long long na = ctx->va1 - ctx->va0;
            if ((r = run_gather(ctx, c, ctx->stream))) return r;      // retained lists of what went through k_pq
            if (na > 0) {
                ChainArgs A{};
                fill_chain_args(ctx, c, A);
                k_list_retained<<<grid_for(ctx, (na + 7) / 8, 8), 256, 0, ctx->stream>>>(
                    A, P<long long>(ctx->ret_base) + ctx->va0, P<long long>(ctx->list_cursor) + c, d_list + base_of[c], cap_of[c]);
                LAUNCH_CHECK();
            }
            CU(cudaMemcpyAsync(ctx->h_cursor + c, P<long long>(ctx->list_cursor) + c, 8, cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaEventRecord(ctx->front_ev[c], ctx->stream));
        }
        const int4* h_coords = (const int4*)batch->coords;
        if (!ctx->pool) ctx->pool = new HostPool(ctx->gather_threads);
        HostPool& pool = *ctx->pool;
        for (int c = 0; c < n_chunks; ++c) {
            CU(cudaEventSynchronize(ctx->front_ev[c]));
            set_view(c);
            const long long R = ctx->h_cursor[c];
            const long long na = ctx->va1 - ctx->va0;
            if (R > 0 && na > 0) {
                ChainArgs A{};
                fill_chain_args(ctx, c, A);
                if (R <= cap_of[c]) {
                    const long long* list = ctx->h_list + base_of[c];
                    int4* stage = ctx->h_stage + base_of[c];
                    if (h_coords) pool.run(R, [=](long long k) {
                        if (k + 8 < R) __builtin_prefetch(h_coords + list[k + 8]);     // random 16-byte records of a multi-GB array
                        stage[k] = h_coords[list[k]];
                    });
                    else {
                        const int32_t *qs = batch->qstart, *qe = batch->qend, *ss = batch->sstart, *se = batch->send;
                        pool.run(R, [=](long long k) {
                            const long long h = list[k];
                            stage[k] = make_int4(qs[h], qe[h], ss[h], se[h]);
                        });
                    }
                    int4* d_st = P<int4>(ctx->d_stage) + base_of[c];
                    CU(cudaMemcpyAsync(d_st, stage, (size_t)R * 16, cudaMemcpyHostToDevice, ctx->tail_stream));
                    k_scatter_coords<<<grid_for(ctx, (na + 7) / 8, 8), 256, 0, ctx->tail_stream>>>(A, P<long long>(ctx->ret_base) + ctx->va0, d_st);
                    LAUNCH_CHECK();
                } else {
                    // a dense chunk: its coordinate records are uploaded whole and read on the device
                    const long long h_lo = batch->hsp_off[ctx->va0], h_hi = batch->hsp_off[ctx->va1];
                    ENS(in_coords, (size_t)b.n_hsp * 16);
                    int4* d_c = P<int4>(ctx->in_coords);
                    if (h_coords) CU(cudaMemcpyAsync(d_c + h_lo, h_coords + h_lo, (size_t)(h_hi - h_lo) * 16, cudaMemcpyHostToDevice, ctx->tail_stream));
                    else {
                        // SoA caller arrays: interleave on the host into the staging records of this chunk, piecewise
                        const int32_t *qs = batch->qstart, *qe = batch->qend, *ss = batch->sstart, *se = batch->send;
                        const long long piece = (long long)(ctx->h_stage_cap / 16);
                        for (long long h = h_lo; h < h_hi; h += piece) {
                            const long long n = h_hi - h < piece ? h_hi - h : piece;
                            int4* stage = ctx->h_stage;
                            CU(cudaStreamSynchronize(ctx->tail_stream));            // the staging buffer is reused
                            pool.run(n, [=](long long k) { stage[k] = make_int4(qs[h + k], qe[h + k], ss[h + k], se[h + k]); });
                            CU(cudaMemcpyAsync(d_c + h, stage, (size_t)n * 16, cudaMemcpyHostToDevice, ctx->tail_stream));
                        }
                    }
                    A.coords4 = d_c; A.work_list = nullptr; A.work_count = nullptr;
                    k_fetch_coords<<<grid_for(ctx, (na + 7) / 8, 8), 256, 0, ctx->tail_stream>>>(A);
                    LAUNCH_CHECK();
                }
            }
            if ((r = run_chain(ctx, c, ctx->tail_stream))) return r;
        }
        CU(cudaEventRecord(ctx->tail_join, ctx->tail_stream));
        CU(cudaStreamWaitEvent(ctx->stream, ctx->tail_join, 0));
    }
    ctx->vl0 = 0; ctx->vl1 = b.n_lnc; ctx->va0 = 0; ctx->va1 = b.n_aln;
    ENS(counters, 5 * 8);
    CU(cudaMemsetAsync(ctx->counters.p, 0, 5 * 8, ctx->stream));
    {
        long long items = b.n_aln > b.n_lnc ? b.n_aln : b.n_lnc;
        k_counts<<<grid_for(ctx, (items + 255) / 256, 4), 256, 0, ctx->stream>>>(
            b.n_aln, b.n_lnc, P<int>(ctx->n_surv), P<int>(ctx->n_keep), P<int>(ctx->chain_len), P<int>(ctx->status),
            P<unsigned long long>(ctx->counters));
        LAUNCH_CHECK();
    }
    if ((r = stage_event(ctx, O2A_STAGE_COUNT))) return r;
    return 0;
    };   // stream_work

    if (replay) {
        CU(cudaGraphLaunch(ctx->graph_exec, ctx->stream));
        ctx->launches += ctx->graph_launches;
        ctx->vl0 = 0; ctx->vl1 = b.n_lnc; ctx->va0 = 0; ctx->va1 = b.n_aln;
        ctx->graph_state = 2;
    } else {
        bool captured = false;
        if (capture) {
            const long long launches0 = ctx->launches;
            const unsigned long long epoch0 = ctx->alloc_epoch;
            cudaGraph_t g = nullptr;
            if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                const int rc = stream_work();
                const cudaError_t ee = cudaStreamEndCapture(ctx->stream, &g);
                if (rc == 0 && ee == cudaSuccess && g && ctx->alloc_epoch == epoch0 &&
                    cudaGraphInstantiate(&ctx->graph_exec, g, 0) == cudaSuccess &&
                    cudaGraphLaunch(ctx->graph_exec, ctx->stream) == cudaSuccess) {
                    ctx->graph_key = key;
                    ctx->graph_launches = ctx->launches - launches0;
                    ctx->graph_state = 1;
                    captured = true;
                } else {
                    // anything the capture could not take: this context goes back to plain launches for good
                    if (ctx->graph_exec) { cudaGraphExecDestroy(ctx->graph_exec); ctx->graph_exec = nullptr; }
                    ctx->graph_disabled = true;
                    ctx->launches = launches0;
                }
                if (g) cudaGraphDestroy(g);
            } else ctx->graph_disabled = true;
            cudaGetLastError();
        }
        if (!captured) {
            if ((r = stream_work())) return r;
        }
        ctx->seen_key = key;
        if (!key.empty()) ctx->seen_key.back() = (long long)ctx->alloc_epoch;   // the epoch after this call's allocations
    }
    unsigned long long c[5], mx[2] = {0, 0};
    CU(cudaMemcpyAsync(c, ctx->counters.p, sizeof(c), cudaMemcpyDeviceToHost, ctx->stream));
    if (batch->mem != O2A_MEM_HOST) CU(cudaMemcpyAsync(mx, ctx->maxchk.p, sizeof(mx), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (batch->mem != O2A_MEM_HOST && ((long long)mx[0] > b.max_aln_hsps || (long long)mx[1] > b.max_lnc_bg)) {
        ctx->err = "max_aln_hsps / max_lnc_bg smaller than the batch's real maxima (" + std::to_string(mx[0]) + ", " +
                   std::to_string(mx[1]) + "): the result is not valid";
        return O2A_EINVAL;
    }
    ctx->counts.n_survivors = (long long)c[0]; ctx->counts.n_retained = (long long)c[1];
    ctx->counts.n_chain_hsps = (long long)c[2]; ctx->counts.n_chains = (long long)c[3];
    ctx->counts.n_failed_lnc = (long long)c[4];
    if (counts) *counts = ctx->counts;
    ctx->have_result = true;
    ctx->ev_valid = ctx->timing;
    return 0;
}

#define NEED_RESULT() do { if (!ctx) return O2A_EINVAL; if (!ctx->have_result) { ctx->err = "no result: call o2a_build_batch first"; return O2A_ESTATE; } CU(cudaSetDevice(ctx->device)); } while (0)

template <typename T>
static int download(o2a_ctx* ctx, T* dst, const void* src, size_t n)
{
    if (dst && n) CU(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    return 0;
}

extern "C" int o2a_fetch_lnc(o2a_ctx* ctx, double* thr, int32_t* status, int64_t* best_aln)
{
    NEED_RESULT();
    int r;
    size_t n = (size_t)ctx->b.n_lnc;
    if ((r = download(ctx, thr, ctx->thr.p, n))) return r;
    if ((r = download(ctx, status, ctx->status.p, n))) return r;
    if ((r = download(ctx, best_aln, ctx->best_aln.p, n))) return r;
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int o2a_fetch_aln(o2a_ctx* ctx, int32_t* n_surv, int32_t* n_keep, int32_t* chain_len, uint8_t* chain_orient,
                             double* chain_score, double* orient_scores)
{
    NEED_RESULT();
    int r;
    size_t n = (size_t)ctx->b.n_aln;
    if ((r = download(ctx, n_surv, ctx->n_surv.p, n))) return r;
    if ((r = download(ctx, n_keep, ctx->n_keep.p, n))) return r;
    if ((r = download(ctx, chain_len, ctx->chain_len.p, n))) return r;
    if ((r = download(ctx, chain_orient, ctx->chain_orient.p, n))) return r;
    if ((r = download(ctx, chain_score, ctx->chain_score.p, n))) return r;
    if ((r = download(ctx, orient_scores, ctx->orient_scores.p, 2 * n))) return r;
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int o2a_fetch_survivors(o2a_ctx* ctx, int64_t* surv_off, int32_t* surv_idx, double* surv_p, double* surv_pq,
                                   uint8_t* surv_keep)
{
    NEED_RESULT();
    const o2a_batch& b = ctx->b;
    const size_t tot = (size_t)ctx->counts.n_survivors;
    ENS(off_a, (b.n_aln + 1) * 8);
    ENS(out_i32, tot * 4); ENS(out_f64a, tot * 8); ENS(out_f64b, tot * 8); ENS(out_u8, tot);
    k_scan_offsets<<<1, 1024, 0, ctx->stream>>>(b.n_aln, P<int>(ctx->n_surv), P<long long>(ctx->off_a));
    LAUNCH_CHECK();
    if (b.n_aln > 0 && tot > 0) {
        k_gather_surv<<<grid_for(ctx, b.n_aln, 16), 128, 0, ctx->stream>>>(
            b.n_aln, (const long long*)b.hsp_off, P<long long>(ctx->off_a), P<int>(ctx->surv_idx), P<double>(ctx->surv_p),
            P<double>(ctx->surv_pq), P<unsigned char>(ctx->surv_keep), P<int>(ctx->out_i32), P<double>(ctx->out_f64a),
            P<double>(ctx->out_f64b), P<unsigned char>(ctx->out_u8));
        LAUNCH_CHECK();
    }
    int r;
    if ((r = download(ctx, surv_off, ctx->off_a.p, (size_t)b.n_aln + 1))) return r;
    if ((r = download(ctx, surv_idx, ctx->out_i32.p, tot))) return r;
    if ((r = download(ctx, surv_p, ctx->out_f64a.p, tot))) return r;
    if ((r = download(ctx, surv_pq, ctx->out_f64b.p, tot))) return r;
    if ((r = download(ctx, surv_keep, ctx->out_u8.p, tot))) return r;
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int o2a_fetch_chains(o2a_ctx* ctx, int64_t* chain_off, int32_t* chain_idx, double* chain_pq)
{
    NEED_RESULT();
    const o2a_batch& b = ctx->b;
    const size_t tot = (size_t)ctx->counts.n_chain_hsps;
    ENS(off_b, (b.n_aln + 1) * 8);
    ENS(out_i32, tot * 4); ENS(out_f64a, tot * 8);
    k_scan_offsets<<<1, 1024, 0, ctx->stream>>>(b.n_aln, P<int>(ctx->chain_len), P<long long>(ctx->off_b));
    LAUNCH_CHECK();
    if (b.n_aln > 0 && tot > 0) {
        k_gather_chain<<<grid_for(ctx, b.n_aln, 16), 64, 0, ctx->stream>>>(
            b.n_aln, (const long long*)b.hsp_off, P<long long>(ctx->off_b), P<int>(ctx->chain_slot), P<int>(ctx->ret_slot),
            P<int>(ctx->surv_idx),
            P<double>(ctx->surv_pq), P<int>(ctx->out_i32), P<double>(ctx->out_f64a));
        LAUNCH_CHECK();
    }
    int r;
    if ((r = download(ctx, chain_off, ctx->off_b.p, (size_t)b.n_aln + 1))) return r;
    if ((r = download(ctx, chain_idx, ctx->out_i32.p, tot))) return r;
    if ((r = download(ctx, chain_pq, ctx->out_f64a.p, tot))) return r;
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// --------------------------------------------------------------------------------------------
// instrumentation: FP64 pipe probes (bench.py reports KDE evaluations against them)
// --------------------------------------------------------------------------------------------
extern "C" int o2a_fp64_probe(o2a_ctx* ctx, int kind, int64_t iters, double* per_second)
{
    if (!ctx || !per_second || kind < 0 || kind > 1 || iters < 1) return O2A_EINVAL;
    CU(cudaSetDevice(ctx->device));
    ENS(maxred, 8);
    const int grid = ctx->sm_count * 8, block = 256;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    k_fp64_probe<<<grid, block, 0, ctx->stream>>>(kind, iters / 8 + 1, 0.5, P<double>(ctx->maxred));     // warm-up
    LAUNCH_CHECK();
    CU(cudaEventRecord(e0, ctx->stream));
    k_fp64_probe<<<grid, block, 0, ctx->stream>>>(kind, iters, 0.5, P<double>(ctx->maxred));
    LAUNCH_CHECK();
    CU(cudaEventRecord(e1, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    const double ops = (double)grid * block * (double)iters * (kind == 0 ? 8.0 : 4.0);
    *per_second = ops / (ms * 1e-3);
    return 0;
}

// --------------------------------------------------------------------------------------------
// operator-level seams
// --------------------------------------------------------------------------------------------
extern "C" int o2a_fitter_eval(o2a_ctx* ctx, int32_t fitter, const int32_t* bg, int64_t n_bg, double kde_factor,
                               const double* x, int64_t n_x, double* sf_out, const double* q, int64_t n_q,
                               double* isf_out, int32_t* status)
{
    if (!ctx) return O2A_EINVAL;
    if (!bg || n_bg < 2 || fitter < 0 || fitter > 2) { ctx->err = "bad fitter arguments"; return O2A_EINVAL; }
    CU(cudaSetDevice(ctx->device));
    ctx->have_result = false;
    long long off[2] = {0, n_bg};
    long long zero2[2] = {0, 0};
    o2a_batch& b = ctx->b;
    b = o2a_batch{};
    b.n_lnc = 1; b.n_bg = n_bg; b.max_lnc_bg = n_bg; b.mem = O2A_MEM_DEVICE;
    int r;
    if ((r = upload(ctx, ctx->in_bg_off, (const int64_t*)off, 2, &b.bg_off))) return r;
    if ((r = upload(ctx, ctx->in_aln_off, (const int64_t*)zero2, 2, &b.aln_off))) return r;
    if ((r = upload(ctx, ctx->in_bg, bg, (size_t)n_bg, &b.bg))) return r;
    b.bg_bits = 32;
    if ((r = upload(ctx, ctx->in_kde, &kde_factor, 1, &b.kde_factor))) return r;
    ctx->prm = o2a_params{};
    ctx->prm.fitter = fitter;
    ctx->vl0 = 0; ctx->vl1 = 1; ctx->va0 = 0; ctx->va1 = 0;
    int st = 0;
    const int nq_eff = n_q > 0 ? (int)n_q : 1;
    for (int k = 0; k < nq_eff; ++k) {
        ctx->prm.alpha = n_q > 0 ? q[k] : 0.5;
        if ((r = run_fit(ctx))) return r;
        double t; int s;
        CU(cudaMemcpyAsync(&t, ctx->thr.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(&s, ctx->status.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (n_q > 0 && isf_out) isf_out[k] = t;
        if (s) st = s;
    }
    if (status) *status = st;
    if (st) { ctx->err = "background fit failed (status " + std::to_string(st) + ")"; return O2A_EINVAL; }
    if (n_x > 0 && sf_out) {
        ENS(out_f64a, (size_t)n_x * 8); ENS(out_f64b, (size_t)n_x * 8);
        CU(cudaMemcpyAsync(ctx->out_f64a.p, x, (size_t)n_x * 8, cudaMemcpyHostToDevice, ctx->stream));
        k_sf_eval<<<(unsigned)((n_x + 127) / 128), 128, 0, ctx->stream>>>(
            fitter, P<double>(ctx->out_f64a), n_x, P<LncFit>(ctx->fits), P<int>(ctx->tbl_k), P<double>(ctx->tbl_cum),
            P<int>(ctx->uv_val), P<int>(ctx->uv_cnt), P<int>(ctx->uv_n), kde_factor, (double)n_bg, P<double>(ctx->out_f64b));
        LAUNCH_CHECK();
        CU(cudaMemcpyAsync(sf_out, ctx->out_f64b.p, (size_t)n_x * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return 0;
}

extern "C" int o2a_best_chain(o2a_ctx* ctx, int64_t n, const int32_t* qstart, const int32_t* qend, const int32_t* sstart,
                              const int32_t* send, const double* score, int64_t qlen, int64_t slen, int32_t gapopen,
                              int32_t gapextend, int32_t* chain_idx, int64_t* chain_len, int32_t* orient,
                              double* chain_score, double* orient_scores, int32_t* status)
{
    if (!ctx) return O2A_EINVAL;
    if (n < 0 || n > 0x7fffffff) { ctx->err = "bad n"; return O2A_EINVAL; }
    CU(cudaSetDevice(ctx->device));
    ctx->have_result = false;
    long long loff[2] = {0, 1}, hoff[2] = {0, n};
    o2a_batch& b = ctx->b;
    b = o2a_batch{};
    b.n_lnc = 1; b.n_aln = 1; b.n_hsp = n; b.max_aln_hsps = n; b.mem = O2A_MEM_DEVICE;
    int r;
    if ((r = upload(ctx, ctx->in_aln_off, (const int64_t*)loff, 2, &b.aln_off))) return r;
    if ((r = upload(ctx, ctx->in_hsp_off, (const int64_t*)hoff, 2, &b.hsp_off))) return r;
    if ((r = upload(ctx, ctx->in_qlen, &qlen, 1, &b.qlen))) return r;
    if ((r = upload(ctx, ctx->in_slen, &slen, 1, &b.slen))) return r;
    if ((r = upload(ctx, ctx->in_qs, qstart, (size_t)n, &b.qstart))) return r;
    if ((r = upload(ctx, ctx->in_qe, qend, (size_t)n, &b.qend))) return r;
    if ((r = upload(ctx, ctx->in_ss, sstart, (size_t)n, &b.sstart))) return r;
    if ((r = upload(ctx, ctx->in_se, send, (size_t)n, &b.send))) return r;
    if ((r = upload(ctx, ctx->in_score, score, (size_t)n, &b.score))) return r;
    ctx->prm = o2a_params{};
    ctx->coords4 = nullptr;
    ctx->prm.gapopen = gapopen; ctx->prm.gapextend = gapextend;
    ENS(aln_lnc, 4); ENS(status, 4); ENS(n_surv, 4); ENS(n_keep, 4);
    ENS(surv_idx, (size_t)n * 4); ENS(surv_keep, (size_t)n); ENS(surv_pq, (size_t)n * 8); ENS(surv_x, (size_t)n * 8);
    ENS(big_list, 2 * sizeof(int)); ENS(mid_list, 2 * sizeof(int)); ENS(pq_list, 2 * sizeof(int)); ENS(cbig_list, 2 * sizeof(int));
    ENS(stage_cnt, 16 * CNT_PER_CHUNK * sizeof(int));
    if ((r = ensure_chain_buffers(ctx))) return r;
    int ni = (int)n;
    CU(cudaMemsetAsync(ctx->aln_lnc.p, 0, 4, ctx->stream));
    CU(cudaMemsetAsync(ctx->status.p, 0, 4, ctx->stream));
    CU(cudaMemsetAsync(ctx->stage_cnt.p, 0, 16 * CNT_PER_CHUNK * sizeof(int), ctx->stream));
    CU(cudaMemcpyAsync(ctx->n_surv.p, &ni, 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->n_keep.p, &ni, 4, cudaMemcpyHostToDevice, ctx->stream));
    if (n > 0) {
        k_iota_keep<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ni, b.score, P<int>(ctx->surv_idx), P<unsigned char>(ctx->surv_keep), P<double>(ctx->surv_x));
        LAUNCH_CHECK();
    }
    bool timing = ctx->timing;
    ctx->timing = false;
    ctx->vl0 = 0; ctx->vl1 = 1; ctx->va0 = 0; ctx->va1 = 1;
    ctx->coords_zero_copy = false; ctx->pipelined_chunks = 1; ctx->coords_on_device = true; ctx->coords_mode_used = O2A_COORDS_AUTO;
    r = run_gather(ctx, 0, ctx->stream, true);
    if (!r) r = run_chain(ctx, 0, ctx->stream);
    ctx->timing = timing;
    if (r) return r;
    int len = 0, st = 0; unsigned char orr = 1; double cs = 0, os2[2] = {0, 0};
    CU(cudaMemcpyAsync(&len, ctx->chain_len.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(&st, ctx->status.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(&orr, ctx->chain_orient.p, 1, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(&cs, ctx->chain_score.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(os2, ctx->orient_scores.p, 16, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (len > 0 && chain_idx) {
        // chain_slot holds retained-list positions == HSP positions here (everything is retained)
        CU(cudaMemcpyAsync(chain_idx, ctx->chain_slot.p, (size_t)len * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    if (chain_len) *chain_len = len;
    if (orient) *orient = orr;
    if (chain_score) *chain_score = cs;
    if (orient_scores) { orient_scores[0] = os2[0]; orient_scores[1] = os2[1]; }
    if (status) *status = st;
    return 0;
}


// --------------------------------------------------------------------------------------------
// producer side: to_genomic + cut_coordinates over a batch of alignments (o2a_cut.cuh)
// --------------------------------------------------------------------------------------------
extern "C" int o2a_cut_batch(o2a_ctx* ctx, const o2a_cut_input* in, int64_t* n_kept)
{
    if (!ctx) return O2A_EINVAL;
    if (!in) { ctx->err = "null input"; return O2A_EINVAL; }
    if (in->n_aln < 0 || in->n_hsp < 0) { ctx->err = "negative size"; return O2A_EINVAL; }
    if (in->n_aln > 0 && (!in->hsp_off || !in->q_start || !in->q_end || !in->q_minus || !in->s_start || !in->s_end || !in->s_minus)) {
        ctx->err = "hsp_off and the six per-alignment range arrays are required"; return O2A_EINVAL;
    }
    if (in->n_hsp > 0 && (!in->qstart || !in->qend || !in->sstart || !in->send || !in->score)) {
        ctx->err = "null HSP array"; return O2A_EINVAL;
    }
    CU(cudaSetDevice(ctx->device));
    ctx->have_cut = false;
    ctx->cut_ev_valid = false;
    const long long n_aln = in->n_aln, n_hsp = in->n_hsp;
    int r;
    if (ctx->timing) CU(cudaEventRecord(ctx->cut_ev[0], ctx->stream));
    CutArgs A{};
    A.n_aln = n_aln;
    if (in->mem == O2A_MEM_HOST) {
        const int64_t* p64 = nullptr; const int32_t* p32 = nullptr; const double* pd = nullptr;
        const int8_t* p8 = nullptr; const uint8_t* pu8 = nullptr;
        if ((r = upload(ctx, ctx->in_hsp_off, in->hsp_off, (size_t)n_aln + 1, &p64))) return r;
        A.hsp_off = (const long long*)p64;
        if ((r = upload(ctx, ctx->in_qs, in->qstart, (size_t)n_hsp, &p32))) return r; A.qs = p32;
        if ((r = upload(ctx, ctx->in_qe, in->qend, (size_t)n_hsp, &p32))) return r; A.qe = p32;
        if ((r = upload(ctx, ctx->in_ss, in->sstart, (size_t)n_hsp, &p32))) return r; A.ss = p32;
        if ((r = upload(ctx, ctx->in_se, in->send, (size_t)n_hsp, &p32))) return r; A.se = p32;
        if ((r = upload(ctx, ctx->in_score, in->score, (size_t)n_hsp, &pd))) return r; A.score = pd;
        if (in->score_is_int) { if ((r = upload(ctx, ctx->ci_isint, in->score_is_int, (size_t)n_hsp, &pu8))) return r; A.is_int = pu8; }
        if ((r = upload(ctx, ctx->ci_qstart, in->q_start, (size_t)n_aln, &p64))) return r; A.q_start = (const long long*)p64;
        if ((r = upload(ctx, ctx->ci_qend, in->q_end, (size_t)n_aln, &p64))) return r; A.q_end = (const long long*)p64;
        if ((r = upload(ctx, ctx->ci_sstart, in->s_start, (size_t)n_aln, &p64))) return r; A.s_start = (const long long*)p64;
        if ((r = upload(ctx, ctx->ci_send, in->s_end, (size_t)n_aln, &p64))) return r; A.s_end = (const long long*)p64;
        if ((r = upload(ctx, ctx->ci_qminus, in->q_minus, (size_t)n_aln, &p8))) return r; A.q_minus = (const signed char*)p8;
        if ((r = upload(ctx, ctx->ci_sminus, in->s_minus, (size_t)n_aln, &p8))) return r; A.s_minus = (const signed char*)p8;
        if (in->relative) { if ((r = upload(ctx, ctx->ci_rel, in->relative, (size_t)n_aln, &pu8))) return r; A.relative = pu8; }
        if (in->qleft) { if ((r = upload(ctx, ctx->ci_ql, in->qleft, (size_t)n_aln, &p64))) return r; A.qleft = (const long long*)p64; }
        if (in->qright) { if ((r = upload(ctx, ctx->ci_qr, in->qright, (size_t)n_aln, &p64))) return r; A.qright = (const long long*)p64; }
        if (in->sleft) { if ((r = upload(ctx, ctx->ci_sl, in->sleft, (size_t)n_aln, &p64))) return r; A.sleft = (const long long*)p64; }
        if (in->sright) { if ((r = upload(ctx, ctx->ci_sr, in->sright, (size_t)n_aln, &p64))) return r; A.sright = (const long long*)p64; }
    } else {
        A.hsp_off = (const long long*)in->hsp_off;
        A.qs = in->qstart; A.qe = in->qend; A.ss = in->sstart; A.se = in->send; A.score = in->score;
        A.is_int = in->score_is_int;
        A.q_start = (const long long*)in->q_start; A.q_end = (const long long*)in->q_end;
        A.s_start = (const long long*)in->s_start; A.s_end = (const long long*)in->s_end;
        A.q_minus = (const signed char*)in->q_minus; A.s_minus = (const signed char*)in->s_minus;
        A.relative = in->relative;
        A.qleft = (const long long*)in->qleft; A.qright = (const long long*)in->qright;
        A.sleft = (const long long*)in->sleft; A.sright = (const long long*)in->sright;
    }
    if (ctx->timing) CU(cudaEventRecord(ctx->cut_ev[1], ctx->stream));
    // tile table: ceil(n/CUT_TILE) tiles per alignment, at most n_hsp/CUT_TILE + n_aln in total
    const long long max_tiles = n_hsp / CUT_TILE + n_aln;
    const long long max_chunks = (max_tiles + SCAN_CHUNK - 1) / SCAN_CHUNK + 1;
    ENS(ct_n, (n_aln + 1) * sizeof(int)); ENS(ct_off, (n_aln + 1) * 8); ENS(ct_aln, (max_tiles + 1) * sizeof(int));
    ENS(ct_cnt, (max_tiles + 1) * sizeof(int)); ENS(ct_out, (max_tiles + 2) * 8); ENS(ct_sums, (max_chunks + 1) * 8);
    ENS(ct_maxq, (n_aln + 1) * sizeof(int)); ENS(ct_maxs, (n_aln + 1) * sizeof(int));
    ENS(co_off, (n_aln + 1) * 8); ENS(co_qlen, (n_aln + 1) * 8); ENS(co_slen, (n_aln + 1) * 8); ENS(co_status, (n_aln + 1) * sizeof(int));
    ENS(co_qs, n_hsp * 4); ENS(co_qe, n_hsp * 4); ENS(co_ss, n_hsp * 4); ENS(co_se, n_hsp * 4);
    ENS(co_score, n_hsp * 8); ENS(co_cut, n_hsp);
    ENS(ct_gen, (max_tiles + 1) * 8); ENS(ct_gen_n, 16); ENS(ct_fast, n_aln + 1);
    ENS(ct_span, (n_hsp / FIXUP_SPAN + 2) * sizeof(int));
    A.span_aln = P<int>(ctx->ct_span);
    CU(cudaMemsetAsync(ctx->ct_gen_n.p, 0, 16, ctx->stream));
    A.gen_tiles = P<long long>(ctx->ct_gen); A.gen_n = P<unsigned long long>(ctx->ct_gen_n); A.aln_fast = P<unsigned char>(ctx->ct_fast);
    A.tile_off = P<long long>(ctx->ct_off); A.tile_aln = P<int>(ctx->ct_aln); A.tile_cnt = P<int>(ctx->ct_cnt);
    A.tile_out = P<long long>(ctx->ct_out);
    A.oqs = P<int>(ctx->co_qs); A.oqe = P<int>(ctx->co_qe); A.oss = P<int>(ctx->co_ss); A.ose = P<int>(ctx->co_se);
    A.oscore = P<double>(ctx->co_score); A.ocut = P<unsigned char>(ctx->co_cut);
    A.max_q = P<int>(ctx->ct_maxq); A.max_s = P<int>(ctx->ct_maxs); A.status = P<int>(ctx->co_status);
    A.out_off = P<long long>(ctx->co_off); A.qlen = P<long long>(ctx->co_qlen); A.slen = P<long long>(ctx->co_slen);
    // tiles per alignment -> tile_off (exclusive scan over the alignments)
    ENS(ct_nal, 16);
    k_cut_tiles_per_aln<<<grid_for(ctx, (n_aln + 255) / 256, 4), 256, 0, ctx->stream>>>(n_aln, A.hsp_off, P<int>(ctx->ct_n), P<long long>(ctx->ct_nal));
    LAUNCH_CHECK();
    {
        const long long aln_chunks = (n_aln + SCAN_CHUNK - 1) / SCAN_CHUNK + 1;
        ENS(ct_sums, ((max_chunks > aln_chunks ? max_chunks : aln_chunks) + 1) * 8);
        const int g = grid_for(ctx, aln_chunks, 8);
        k_scan_chunk_sums<<<g, 256, 0, ctx->stream>>>(P<long long>(ctx->ct_nal), P<int>(ctx->ct_n), P<long long>(ctx->ct_sums));
        LAUNCH_CHECK();
        k_scan_sums<<<1, 1024, 0, ctx->stream>>>(P<long long>(ctx->ct_nal), P<long long>(ctx->ct_sums));
        LAUNCH_CHECK();
        k_scan_chunks<<<g, 256, 0, ctx->stream>>>(P<long long>(ctx->ct_nal), P<int>(ctx->ct_n), P<long long>(ctx->ct_sums), P<long long>(ctx->ct_off));
        LAUNCH_CHECK();
    }
    if (n_aln > 0) {
        k_cut_fill_tile_aln<<<grid_for(ctx, n_aln, 32), 32, 0, ctx->stream>>>(A, P<int>(ctx->ct_aln));
        LAUNCH_CHECK();
    }
    const int tile_grid = grid_for(ctx, (max_tiles + CUT_WARPS - 1) / CUT_WARPS, 8);
    // count pass (fast class + general class), scan of the tile counts, write pass, straddler fix-up
    k_cut_fast<false><<<tile_grid, CUT_WARPS * 32, 0, ctx->stream>>>(A);
    LAUNCH_CHECK();
    k_cut_general<false><<<tile_grid, CUT_WARPS * 32, 0, ctx->stream>>>(A);
    LAUNCH_CHECK();
    if (ctx->timing) CU(cudaEventRecord(ctx->cut_ev[2], ctx->stream));
    const long long* n_tiles_ptr = A.tile_off + n_aln;
    const int scan_grid = grid_for(ctx, max_chunks, 8);
    k_scan_chunk_sums<<<scan_grid, 256, 0, ctx->stream>>>(n_tiles_ptr, A.tile_cnt, P<long long>(ctx->ct_sums));
    LAUNCH_CHECK();
    k_scan_sums<<<1, 1024, 0, ctx->stream>>>(n_tiles_ptr, P<long long>(ctx->ct_sums));
    LAUNCH_CHECK();
    k_scan_chunks<<<scan_grid, 256, 0, ctx->stream>>>(n_tiles_ptr, A.tile_cnt, P<long long>(ctx->ct_sums), P<long long>(ctx->ct_out));
    LAUNCH_CHECK();
    k_cut_offsets<<<grid_for(ctx, (n_aln + 256) / 256, 4), 256, 0, ctx->stream>>>(A);
    LAUNCH_CHECK();
    if (ctx->timing) CU(cudaEventRecord(ctx->cut_ev[3], ctx->stream));
    k_cut_fast<true><<<tile_grid, CUT_WARPS * 32, 0, ctx->stream>>>(A);
    LAUNCH_CHECK();
    k_cut_general<true><<<tile_grid, CUT_WARPS * 32, 0, ctx->stream>>>(A);
    LAUNCH_CHECK();
    k_cut_fixup<<<grid_for(ctx, n_hsp / FIXUP_SPAN + 1, 4), FIXUP_THREADS, 0, ctx->stream>>>(A);
    LAUNCH_CHECK();
    if (n_aln > 0) {
        k_cut_finish<<<grid_for(ctx, (n_aln + 255) / 256, 4), 256, 0, ctx->stream>>>(A);
        LAUNCH_CHECK();
    }
    if (ctx->timing) CU(cudaEventRecord(ctx->cut_ev[4], ctx->stream));
    long long kept = 0;
    CU(cudaMemcpyAsync(&kept, P<long long>(ctx->co_off) + n_aln, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->cut_n_aln = n_aln; ctx->cut_n_kept = kept;
    ctx->have_cut = true;
    ctx->cut_ev_valid = ctx->timing;
    if (n_kept) *n_kept = kept;
    return 0;
}

#define NEED_CUT() do { if (!ctx) return O2A_EINVAL; if (!ctx->have_cut) { ctx->err = "no result: call o2a_cut_batch first"; return O2A_ESTATE; } CU(cudaSetDevice(ctx->device)); } while (0)

extern "C" int o2a_fetch_cut(o2a_ctx* ctx, int64_t* hsp_off, int32_t* qstart, int32_t* qend, int32_t* sstart, int32_t* send,
                             double* score, uint8_t* cut, int64_t* qlen, int64_t* slen, int32_t* status)
{
    NEED_CUT();
    int r;
    const size_t na = (size_t)ctx->cut_n_aln, nk = (size_t)ctx->cut_n_kept;
    if ((r = download(ctx, hsp_off, ctx->co_off.p, na + 1))) return r;
    if ((r = download(ctx, qstart, ctx->co_qs.p, nk))) return r;
    if ((r = download(ctx, qend, ctx->co_qe.p, nk))) return r;
    if ((r = download(ctx, sstart, ctx->co_ss.p, nk))) return r;
    if ((r = download(ctx, send, ctx->co_se.p, nk))) return r;
    if ((r = download(ctx, score, ctx->co_score.p, nk))) return r;
    if ((r = download(ctx, cut, ctx->co_cut.p, nk))) return r;
    if ((r = download(ctx, qlen, ctx->co_qlen.p, na))) return r;
    if ((r = download(ctx, slen, ctx->co_slen.p, na))) return r;
    if ((r = download(ctx, status, ctx->co_status.p, na))) return r;
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int o2a_cut_view(o2a_ctx* ctx, o2a_batch* out)
{
    NEED_CUT();
    if (!out) { ctx->err = "null out"; return O2A_EINVAL; }
    memset(out, 0, sizeof(*out));
    out->n_aln = ctx->cut_n_aln; out->n_hsp = ctx->cut_n_kept;
    out->hsp_off = P<int64_t>(ctx->co_off); out->qlen = P<int64_t>(ctx->co_qlen); out->slen = P<int64_t>(ctx->co_slen);
    out->qstart = P<int32_t>(ctx->co_qs); out->qend = P<int32_t>(ctx->co_qe);
    out->sstart = P<int32_t>(ctx->co_ss); out->send = P<int32_t>(ctx->co_se);
    out->score = P<double>(ctx->co_score);
    out->mem = O2A_MEM_DEVICE;
    return 0;
}

extern "C" int o2a_get_cut_ms(o2a_ctx* ctx, float* ms)
{
    if (!ctx || !ms) return O2A_EINVAL;
    if (!ctx->cut_ev_valid) { ctx->err = "no timed o2a_cut_batch yet (o2a_set_timing)"; return O2A_ESTATE; }
    for (int i = 0; i < 4; ++i) CU(cudaEventElapsedTime(&ms[i], ctx->cut_ev[i], ctx->cut_ev[i + 1]));
    return 0;
}


// --------------------------------------------------------------------------------------------
// annotate_orthologs: neighbours of every ortholog in the sorted annotation + JI / OC (o2a_annotate.cuh)
// --------------------------------------------------------------------------------------------
extern "C" int o2a_annotate_batch(o2a_ctx* ctx, const o2a_ann_input* in, int64_t* n_pairs)
{
    if (!ctx) return O2A_EINVAL;
    if (!in) { ctx->err = "null input"; return O2A_EINVAL; }
    if (in->n_q < 0 || in->n_q > INT32_MAX - 64 || in->n_a < 0 || in->n_a > INT32_MAX - ANN_CHUNK) { ctx->err = "bad size"; return O2A_EINVAL; }
    if (in->distance < 0) { ctx->err = "negative distance"; return O2A_EINVAL; }
    if (in->n_q > 0 && (!in->q_chrom || !in->q_start || !in->q_end || !in->q_name || !in->q_strand)) {
        ctx->err = "null ortholog array"; return O2A_EINVAL;
    }
    if (in->n_a > 0 && (!in->a_chrom || !in->a_start || !in->a_end || !in->a_name || !in->a_strand)) {
        ctx->err = "null annotation array"; return O2A_EINVAL;
    }
    CU(cudaSetDevice(ctx->device));
    ctx->have_ann = false;
    ctx->ann_ev_valid = false;
    const long long n_q = in->n_q, n_a = in->n_a;
    int r;
    if (ctx->timing) CU(cudaEventRecord(ctx->ann_ev[0], ctx->stream));
    AnnArgs A{};
    A.n_q = n_q; A.n_a = n_a; A.distance = in->distance; A.strandness = in->strandness;
    if (in->mem == O2A_MEM_HOST) {
        const int32_t* p32 = nullptr; const int8_t* p8 = nullptr;
        if ((r = upload(ctx, ctx->an_qc, in->q_chrom, (size_t)n_q, &p32))) return r; A.q_chrom = p32;
        if ((r = upload(ctx, ctx->an_qs, in->q_start, (size_t)n_q, &p32))) return r; A.q_start = p32;
        if ((r = upload(ctx, ctx->an_qe, in->q_end, (size_t)n_q, &p32))) return r; A.q_end = p32;
        if ((r = upload(ctx, ctx->an_qn, in->q_name, (size_t)n_q, &p32))) return r; A.q_name = p32;
        if ((r = upload(ctx, ctx->an_qstr, in->q_strand, (size_t)n_q, &p8))) return r; A.q_strand = (const signed char*)p8;
        if ((r = upload(ctx, ctx->an_ac, in->a_chrom, (size_t)n_a, &p32))) return r; A.a_chrom = p32;
        if ((r = upload(ctx, ctx->an_as, in->a_start, (size_t)n_a, &p32))) return r; A.a_start = p32;
        if ((r = upload(ctx, ctx->an_ae, in->a_end, (size_t)n_a, &p32))) return r; A.a_end = p32;
        if ((r = upload(ctx, ctx->an_an, in->a_name, (size_t)n_a, &p32))) return r; A.a_name = p32;
        if ((r = upload(ctx, ctx->an_astr, in->a_strand, (size_t)n_a, &p8))) return r; A.a_strand = (const signed char*)p8;
    } else {
        A.q_chrom = in->q_chrom; A.q_start = in->q_start; A.q_end = in->q_end; A.q_name = in->q_name;
        A.q_strand = (const signed char*)in->q_strand;
        A.a_chrom = in->a_chrom; A.a_start = in->a_start; A.a_end = in->a_end; A.a_name = in->a_name;
        A.a_strand = (const signed char*)in->a_strand;
    }
    if (ctx->timing) CU(cudaEventRecord(ctx->ann_ev[1], ctx->stream));
    const long long a_chunks = (n_a + ANN_CHUNK - 1) / ANN_CHUNK, q_chunks = (n_q + SCAN_CHUNK - 1) / SCAN_CHUNK + 1;
    ENS(an_prefmax, (n_a + 1) * 8); ENS(an_chunkmax, (a_chunks + 1) * 8); ENS(an_blkmax, (n_a / ANN_BLK + 2) * sizeof(int));
    ENS(an_cnt, (n_q + 1) * sizeof(int)); ENS(an_span, (n_q + 1) * sizeof(int4)); ENS(an_off, (n_q + 1) * 8);
    ENS(an_sums, (q_chunks + 1) * 8); ENS(an_n, 16);
    A.prefmax = P<unsigned long long>(ctx->an_prefmax); A.chunk_max = P<unsigned long long>(ctx->an_chunkmax); A.blkmax = P<int>(ctx->an_blkmax);
    A.cnt = P<int>(ctx->an_cnt); A.span = P<int4>(ctx->an_span); A.off = P<long long>(ctx->an_off);
    if (n_a > 0) {
        const int g = grid_for(ctx, a_chunks, 8);
        k_ann_chunk_max<<<g, 256, 0, ctx->stream>>>(A);
        LAUNCH_CHECK();
        k_ann_scan_max<<<1, 1024, 0, ctx->stream>>>(A);
        LAUNCH_CHECK();
        k_ann_prefmax<<<g, 256, 0, ctx->stream>>>(A);
        LAUNCH_CHECK();
    }
    if (ctx->timing) CU(cudaEventRecord(ctx->ann_ev[2], ctx->stream));
    const int per_cta = ANN_THREADS;                 // one lane per ortholog
    const int qgrid = grid_for(ctx, (n_q + per_cta - 1) / per_cta, 8);
    if (n_q > 0) {
        k_annotate<false><<<qgrid, ANN_THREADS, 0, ctx->stream>>>(A);
        LAUNCH_CHECK();
    }
    CU(cudaMemcpyAsync(ctx->an_n.p, &n_q, 8, cudaMemcpyHostToDevice, ctx->stream));
    {
        const int g = grid_for(ctx, q_chunks, 8);
        k_scan_chunk_sums<<<g, 256, 0, ctx->stream>>>(P<long long>(ctx->an_n), A.cnt, P<long long>(ctx->an_sums));
        LAUNCH_CHECK();
        k_scan_sums<<<1, 1024, 0, ctx->stream>>>(P<long long>(ctx->an_n), P<long long>(ctx->an_sums));
        LAUNCH_CHECK();
        k_scan_chunks<<<g, 256, 0, ctx->stream>>>(P<long long>(ctx->an_n), A.cnt, P<long long>(ctx->an_sums), P<long long>(ctx->an_off));
        LAUNCH_CHECK();
    }
    long long pairs = 0;
    CU(cudaMemcpyAsync(&pairs, P<long long>(ctx->an_off) + n_q, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));      // the pair count sizes the outputs
    if (ctx->timing) CU(cudaEventRecord(ctx->ann_ev[3], ctx->stream));
    ENS(an_idx, (pairs + 1) * sizeof(int)); ENS(an_ji, (pairs + 1) * 8); ENS(an_oc, (pairs + 1) * 8); ENS(an_pq, (pairs + 1) * sizeof(int));
    A.nb_idx = P<int>(ctx->an_idx); A.ji = P<double>(ctx->an_ji); A.oc = P<double>(ctx->an_oc);
    A.pair_q = P<int>(ctx->an_pq); A.n_pairs = pairs;
    if (n_q > 0 && pairs > 0) {
        k_annotate<true><<<qgrid, ANN_THREADS, 0, ctx->stream>>>(A);
        LAUNCH_CHECK();
        k_ann_ji_oc<<<grid_for(ctx, (pairs + 255) / 256, 8), 256, 0, ctx->stream>>>(A);
        LAUNCH_CHECK();
    }
    if (ctx->timing) CU(cudaEventRecord(ctx->ann_ev[4], ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->ann_n_q = n_q; ctx->ann_n_pairs = pairs;
    ctx->have_ann = true;
    ctx->ann_ev_valid = ctx->timing;
    if (n_pairs) *n_pairs = pairs;
    return 0;
}

extern "C" int o2a_fetch_annotation(o2a_ctx* ctx, int64_t* nb_off, int32_t* nb_idx, double* ji, double* oc)
{
    if (!ctx) return O2A_EINVAL;
    if (!ctx->have_ann) { ctx->err = "no result: call o2a_annotate_batch first"; return O2A_ESTATE; }
    CU(cudaSetDevice(ctx->device));
    int r;
    const size_t nq = (size_t)ctx->ann_n_q, np = (size_t)ctx->ann_n_pairs;
    if ((r = download(ctx, nb_off, ctx->an_off.p, nq + 1))) return r;
    if ((r = download(ctx, nb_idx, ctx->an_idx.p, np))) return r;
    if ((r = download(ctx, ji, ctx->an_ji.p, np))) return r;
    if ((r = download(ctx, oc, ctx->an_oc.p, np))) return r;
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int o2a_get_annotate_ms(o2a_ctx* ctx, float* ms)
{
    if (!ctx || !ms) return O2A_EINVAL;
    if (!ctx->ann_ev_valid) { ctx->err = "no timed o2a_annotate_batch yet (o2a_set_timing)"; return O2A_ESTATE; }
    for (int i = 0; i < 4; ++i) CU(cudaEventElapsedTime(&ms[i], ctx->ann_ev[i], ctx->ann_ev[i + 1]));
    return 0;
}
