/*
 * o2a.h — C ABI of libo2a_cuda.so: ortho2align's build_orthologs / get_best_orthologs hot path
 * on one B200 (sm_100a). Plain pointers and sizes only; no torch / C++ types cross this boundary.
 *
 * The reference (dmitrymyl/ortho2align, pure Python) has no FFI; its seams for this path are
 * Python callables. Each entry point below names the reference interface it replaces
 * (file:line under /root/reference/ortho2align) — INTEGRATION.md shows the ctypes stub a
 * maintainer would add on the reference side.
 *
 * Conventions
 *   - every function returning int returns 0 on success, a negative O2A_E* code otherwise;
 *     o2a_last_error(ctx) gives the message. No exceptions cross the boundary.
 *   - one ctx per GPU; a ctx is single-threaded (caller serialises); different ctxs may be used
 *     from different threads / processes (one process per GPU under torchrun).
 *   - the library never frees caller memory. Input pointers are host pointers (mem = O2A_MEM_HOST;
 *     the library copies them host->device on its stream, pinned memory from o2a_host_alloc makes
 *     that copy asynchronous) or device pointers on ctx's device (mem = O2A_MEM_DEVICE; zero copy).
 *   - besides the caller's stream (o2a_set_stream) a ctx owns two internal streams: one for the chunked
 *     host->device copies of large host batches, one for kernels that run beside the main ones. Both are
 *     forked from and joined back into the caller's stream with events inside every call; when a call
 *     returns, all of its work has completed.
 *   - there is NO CPU fallback: without a usable CUDA device o2a_create fails.
 *
 * Data layout (three CSR levels; ortho2align_b200/columnar.py):
 *   lncRNA l    : bg[bg_off[l] .. bg_off[l+1])  background scores AFTER the reference's patches
 *                 (pipeline.py:746-754); alignments aln_off[l] .. aln_off[l+1]
 *   alignment a : HSPs hsp_off[a] .. hsp_off[a+1]; qlen[a], slen[a] as stored in the JSON
 *                 (alignment.py:477-487) — inputs, never recomputed
 *   HSP h       : qstart,qend,sstart,send (int32 genomic coordinates), score (float64)
 */
#ifndef O2A_H
#define O2A_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define O2A_ABI_VERSION 1

/* fitter: pipeline.py:839-842 maps 'hist' -> HistogramFitter, 'kde' -> KernelFitter.
 * O2A_FIT_EMPIRICAL is an extension with no reference counterpart. */
#define O2A_FIT_HIST 0
#define O2A_FIT_KDE 1
#define O2A_FIT_EMPIRICAL 2

/* get_best_orthologs -value / -function (pipeline.py:981-989, cli_scripts.py:339-350) */
#define O2A_BEST_BLOCK_LENGTH 0
#define O2A_BEST_TOTAL_LENGTH 1
#define O2A_BEST_BLOCK_COUNT 2
#define O2A_BEST_WEIGHT 3
#define O2A_BEST_MAX 0
#define O2A_BEST_MIN 1

#define O2A_MEM_HOST 0
#define O2A_MEM_DEVICE 1

/* per-lncRNA status: 0, or what the reference would have raised for this lncRNA (the lncRNA then
 * belongs in query_exceptions.bed, pipeline.py:797-798, 862-868) */
#define O2A_STATUS_OK 0
#define O2A_STATUS_ORIENTATION_NONE 1  /* alignment.py:865-872 KeyError/TypeError */
#define O2A_STATUS_BRENTQ 2            /* fitting.py:239 RuntimeError (no convergence) */
#define O2A_STATUS_TOO_MANY_BINS 3     /* fitting.py:152 np.histogram ValueError */
#define O2A_STATUS_UNSUPPORTED 4       /* round 1: library limit on the background's value range / distinct values.
                                          No longer produced — any integer background is fitted (the general path
                                          sorts it in HBM scratch); the value stays reserved */

#define O2A_EINVAL (-1)
#define O2A_ECUDA (-2)
#define O2A_ENOMEM (-3)
#define O2A_ESTATE (-4)

typedef struct o2a_ctx o2a_ctx;

typedef struct o2a_batch {
    int64_t n_lnc, n_aln, n_hsp, n_bg;
    const int64_t *bg_off;     /* [n_lnc+1] */
    const int32_t *bg;         /* [n_bg] (int16 data when bg_bits == 16) */
    const double  *kde_factor; /* [n_lnc] gaussian_kde.factor per lncRNA (host-side scalar of
                                  fitting.py:264,295); may be NULL unless fitter == O2A_FIT_KDE */
    const int64_t *aln_off;    /* [n_lnc+1] */
    const int64_t *hsp_off;    /* [n_aln+1] */
    const int64_t *qlen;       /* [n_aln] */
    const int64_t *slen;       /* [n_aln] */
    const int32_t *qstart, *qend, *sstart, *send; /* [n_hsp] SoA coordinates; ignored when `coords` is set */
    const double  *score;      /* [n_hsp] */
    int32_t mem;               /* O2A_MEM_HOST | O2A_MEM_DEVICE (applies to every pointer above) */
    int32_t reserved;
    int64_t max_aln_hsps;      /* max over a of hsp_off[a+1]-hsp_off[a]; 0 = let the library find out. A hint: it is */
    int64_t max_lnc_bg;        /* max over l of bg_off[l+1]-bg_off[l];   verified, and an understated value fails the call (O2A_EINVAL) */
    const int32_t *coords;     /* optional [4*n_hsp] AoS (qstart,qend,sstart,send) per HSP, 16-byte aligned:
                                  one 16 B read per retained HSP instead of four (matters for zero-copy
                                  reads from pinned host memory); NULL = use the four SoA arrays */
    /* Optional narrow scores (BLAST raw scores are integers; only HSPs cut at a gene boundary are
     * fractional, alignment.py:330). score_bits = 8, 16 or 32: `score_q` holds the scores as int8 /
     * int16 / int32; the most negative value marks an HSP whose exact float64 score is listed in the
     * alignment's exception list exc_pos[exc_off[a]..exc_off[a+1]) (ascending HSP positions inside
     * the alignment) / exc_val. score_bits = 0: `score` (float64) is used. Lossless either way. */
    int32_t score_bits;
    int32_t bg_bits;           /* 0 or 32: `bg` is int32 [n_bg]; 16 / 8: `bg` points to int16 / int8 [n_bg]
                                  (the caller picks a width every background score fits) */
    const void    *score_q;    /* [n_hsp] int8, int16 or int32 */
    const int64_t *exc_off;    /* [n_aln+1] */
    const int32_t *exc_pos;    /* [n_exc] */
    const double  *exc_val;    /* [n_exc] */
} o2a_batch;

/* build_orthologs(threshold, gapopen, gapextend, fdr, fitting) + get_best_orthologs(value,
 * function): pipeline.py:801-812, 959-968 */
typedef struct o2a_params {
    int32_t fitter;
    int32_t fdr;
    double  alpha;
    int32_t gapopen, gapextend;
    int32_t best_value, best_function;
} o2a_params;

typedef struct o2a_counts {
    int64_t n_survivors;   /* HSPs with score > isf(alpha)               (pipeline.py:765) */
    int64_t n_retained;    /* HSPs entering the chaining                 (pipeline.py:774-776) */
    int64_t n_chain_hsps;  /* HSPs in the chosen chains                  (alignment.py:1032) */
    int64_t n_chains;      /* alignments with a non-empty chain = ortholog records */
    int64_t n_failed_lnc;  /* lncRNAs with status != 0 */
} o2a_counts;

int  o2a_abi_version(void);
int  o2a_create(o2a_ctx **out, int device);
void o2a_destroy(o2a_ctx *ctx);
const char *o2a_last_error(const o2a_ctx *ctx);
/* run on an existing cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream); NULL = own stream */
int  o2a_set_stream(o2a_ctx *ctx, void *cuda_stream);

/* pinned (page-locked, device-mapped) host memory for batch inputs / fetched outputs */
void *o2a_host_alloc(size_t bytes);
void  o2a_host_free(void *p);

/*
 * Replaces the pool fan-out of pipeline.py:846-860: `_build` (pipeline.py:722-798) for EVERY
 * lncRNA of the batch in one call, plus the selection of get_best_orthologs (pipeline.py:959-1024):
 *   background fit + isf(alpha)      fitting.py:142-155, 206-215 (hist) / 254-264, 335-356 (kde)
 *   score filter, strict '>'         alignment.py:618-646, pipeline.py:765
 *   p-values sf(score)               fitting.py:184-193 (hist) / 282-310 (kde), pipeline.py:766
 *   q = (m*p)/rank, keep q <= alpha  fitting.py:359-364, pipeline.py:769-776 (ties: stable order)
 *   best chain per alignment         alignment.py:863-927, 982-1046
 *   best ortholog per lncRNA         pipeline.py:981-1016
 * Results stay on the device inside ctx until fetched. Blocks until the counts are known.
 */
int o2a_build_batch(o2a_ctx *ctx, const o2a_batch *batch, const o2a_params *params, o2a_counts *counts);

/* Results of the last o2a_build_batch. Output pointers are HOST pointers sized as noted; any may
 * be NULL to skip that array. */
int o2a_fetch_lnc(o2a_ctx *ctx, double *thr /*n_lnc*/, int32_t *status /*n_lnc*/, int64_t *best_aln /*n_lnc*/);
int o2a_fetch_aln(o2a_ctx *ctx, int32_t *n_surv, int32_t *n_keep, int32_t *chain_len /*n_aln each*/,
                  uint8_t *chain_orient /*n_aln: 0 direct, 1 reverse*/, double *chain_score /*n_aln*/,
                  double *orient_scores /*2*n_aln: direct, reverse*/);
int o2a_fetch_survivors(o2a_ctx *ctx, int64_t *surv_off /*n_aln+1*/, int32_t *surv_idx, double *surv_p,
                        double *surv_pq, uint8_t *surv_keep /*n_survivors each*/);
int o2a_fetch_chains(o2a_ctx *ctx, int64_t *chain_off /*n_aln+1*/, int32_t *chain_idx, double *chain_pq /*n_chain_hsps*/);

/* Operator-level seams (used by the Python mirrors of the reference classes):
 *  - fitter plugin `F(data).sf(x)`, `.isf(q)`           fitting.py:8-85 (AbstractFitter seam)
 *  - `Alignment.best_chain(gapopen, gapextend)`          alignment.py:1032-1046
 * Host pointers. `bg` must already be patched. Returns 0, or O2A_EINVAL with status in *status. */
int o2a_fitter_eval(o2a_ctx *ctx, int32_t fitter, const int32_t *bg, int64_t n_bg, double kde_factor,
                    const double *x, int64_t n_x, double *sf_out,
                    const double *q, int64_t n_q, double *isf_out, int32_t *status);
int o2a_best_chain(o2a_ctx *ctx, int64_t n, const int32_t *qstart, const int32_t *qend,
                   const int32_t *sstart, const int32_t *send, const double *score,
                   int64_t qlen, int64_t slen, int32_t gapopen, int32_t gapextend,
                   int32_t *chain_idx /*n*/, int64_t *chain_len, int32_t *orient, double *chain_score,
                   double *orient_scores /*2*/, int32_t *status);

/* Instrumentation for bench.py: kernels launched by this ctx since creation, and CUDA-event
 * durations (ms, on the launching stream) of the stages of the last o2a_build_batch when timing
 * is enabled. Stage ids: */
#define O2A_STAGE_H2D 0
#define O2A_STAGE_FIT 1
#define O2A_STAGE_FILTER 2    /* score filter + ordered compaction (the HBM-bound streaming kernel) */
#define O2A_STAGE_PQ 3        /* p-values, ranks, q, keep (survivors only) */
#define O2A_STAGE_GATHER 4    /* coordinates + score of the retained HSPs (zero-copy over PCIe for pinned host input) */
#define O2A_STAGE_CHAIN 5
#define O2A_STAGE_BEST 6
#define O2A_STAGE_COUNT 7
int64_t o2a_launch_count(const o2a_ctx *ctx);
/* How the last o2a_build_batch issued its kernels: 0 = plain launches; 1 = captured into a CUDA graph and launched;
 * 2 = replay of that graph (one launch). A DEVICE batch whose pointers, sizes, maxima hints and parameters equal the
 * previous call's, with stage timing off, is captured on the second such call and replayed from the third (the
 * reference has no counterpart: its unit of dispatch is one process-pool task per lncRNA, parallel.py:296-488).
 * Replays count their kernels in o2a_launch_count. Environment O2A_NO_GRAPH=1 disables it. */
int  o2a_graph_state(const o2a_ctx *ctx);
/* 1 if the last host-memory o2a_build_batch read the HSP coordinates in place (pinned, mapped) */
int  o2a_coords_zero_copy(const o2a_ctx *ctx);
/* number of lncRNA chunks the last o2a_build_batch pipelined its host uploads in (1 = single pass).
 * Host batches of >= 4 Mi HSPs are cut into chunks (default 8; env O2A_CHUNKS=1..16, O2A_NO_PIPELINE=1,
 * O2A_PIPELINE_MIN_HSPS) so that the H2D copy of chunk c+1 overlaps the kernels of chunk c; with more than
 * one chunk the stage times of o2a_get_stage_ms are those of the LAST chunk and "h2d" spans the others. */
int  o2a_pipeline_chunks(const o2a_ctx *ctx);
/* How the coordinate records of the retained HSPs (~1 % of a BLAST-like batch) of a HOST batch reach the device:
 *   ZERO_COPY    read in place over PCIe by the gather kernel (needs pinned, device-mapped caller memory)
 *   HOST_GATHER  the GPU lists the retained HSP indices into pinned host memory, host threads (O2A_GATHER_THREADS,
 *                default min(16, cores)) collect the 16-byte records, one small DMA per chunk brings them back;
 *                works with pageable caller memory; a chunk that retains more than 1/8 of its HSPs is uploaded whole
 *   UPLOAD       all 16 B per HSP cross PCIe
 *   AUTO         HOST_GATHER (as fast or faster than ZERO_COPY on a B200 box, and indifferent to how the caller
 *                allocated its coordinate arrays)
 * Environment override: O2A_COORDS=zerocopy|gather|upload. o2a_coords_mode: what the last host call used. */
#define O2A_COORDS_AUTO 0
#define O2A_COORDS_ZERO_COPY 1
#define O2A_COORDS_HOST_GATHER 2
#define O2A_COORDS_UPLOAD 3
int  o2a_set_coords_mode(o2a_ctx *ctx, int mode);
int  o2a_coords_mode(const o2a_ctx *ctx);
/* FP64 pipe probes of the whole device (what the KDE p-value evaluations are bound by; MEASURED_PEAKS.json has no
 * FP64 figure): kind 0 = dependent-chain DFMA rate (fma/s; x2 = FLOP/s), kind 1 = ndtr (0.5 erfc) evaluations/s with the
 * library's own device function. iters = chain length per thread. */
int  o2a_fp64_probe(o2a_ctx *ctx, int kind, int64_t iters, double *per_second);
int  o2a_set_timing(o2a_ctx *ctx, int enabled);
int  o2a_get_stage_ms(o2a_ctx *ctx, float *ms /*O2A_STAGE_COUNT*/);

/* ---------------------------------------------------------------------------------------------
 * Producer side of the alignment JSON: what `_align_syntenies`
 * (pipeline.py:525-532) applies to the HSPs of every BLAST table between `from_file_blast` and
 * `to_dict`, for a whole batch of alignments in one call:
 *     alignment.to_genomic()                                    genomicranges.py:1108-1129
 *     alignment.cut_coordinates(qleft, qright, sleft, sright)   alignment.py:734-803, 267-333;
 *                                                               genomicranges.py:1153-1158
 *     qlen / slen of the new alignment                          alignment.py:477-487
 * Input coordinates are what `from_file_blast` leaves (0-based, half-open, sequence-relative;
 * alignment.py:574-581). The result stays on the device in the layout o2a_build_batch takes
 * (o2a_cut_view) and can be fetched (o2a_fetch_cut). Kept HSPs keep their order.
 * ------------------------------------------------------------------------------------------- */
#define O2A_NO_BOUND INT64_MIN             /* a boundary that is None */
#define O2A_CUT_STATUS_RANGE 1             /* a kept coordinate does not fit int32: the alignment's coordinates,
                                              qlen and slen are not meaningful */
#define O2A_CUT_STATUS_INEXACT 2           /* an intermediate product beyond 2^53 (Python's exact ints would differ) */

typedef struct o2a_cut_input {
    int64_t n_aln, n_hsp;
    const int64_t *hsp_off;                       /* [n_aln+1] */
    const int32_t *qstart, *qend, *sstart, *send; /* [n_hsp] */
    const double  *score;                         /* [n_hsp] */
    const uint8_t *score_is_int;                  /* [n_hsp] optional: 0 = the table held a non-integral score (a Python
                                                     float before any cut); NULL = all ints. Only carried to `cut` */
    /* per alignment [n_aln]: query range (the flanked neighbourhood) and subject range (the syntenic
     * window): start, end, strand == '-' (genomicranges.py:1112, 1121) */
    const int64_t *q_start, *q_end;
    const int8_t  *q_minus;
    const int64_t *s_start, *s_end;
    const int8_t  *s_minus;
    const uint8_t *relative;                      /* [n_aln] 0 = to_genomic is a no-op (:1109-1110); NULL = all 1 */
    const int64_t *qleft, *qright, *sleft, *sright; /* [n_aln] each; NULL array or O2A_NO_BOUND entry = None */
    int32_t mem;                                  /* O2A_MEM_HOST | O2A_MEM_DEVICE, applies to every pointer */
    int32_t reserved;
} o2a_cut_input;

int o2a_cut_batch(o2a_ctx *ctx, const o2a_cut_input *in, int64_t *n_kept);
/* any pointer may be NULL. hsp_off[n_aln+1]; per kept HSP: coordinates, score, cut (bit 0 = the score was
 * scaled by a cut, bit 1 = it was a float already: the reference's JSON number is an int iff cut == 0);
 * per alignment: qlen, slen, status */
int o2a_fetch_cut(o2a_ctx *ctx, int64_t *hsp_off, int32_t *qstart, int32_t *qend, int32_t *sstart, int32_t *send,
                  double *score, uint8_t *cut, int64_t *qlen, int64_t *slen, int32_t *status);
/* device pointers of the last cut in `out` (n_aln, n_hsp, hsp_off, qlen, slen, qstart..send, score,
 * mem = O2A_MEM_DEVICE; everything else zero): add aln_off / bg (device) and hand it to o2a_build_batch.
 * Valid until the next o2a_cut_batch on this ctx. */
int o2a_cut_view(o2a_ctx *ctx, o2a_batch *out);
/* CUDA-event durations (ms) of the last o2a_cut_batch when timing is enabled: h2d, count pass, scan, write pass */
int o2a_get_cut_ms(o2a_ctx *ctx, float *ms /*4*/);

/* ---------------------------------------------------------------------------------------------
 * annotate_orthologs' interval join (pipeline.py:1027-1094):
 *     subject_orthologs.get_neighbours(annotation, relation='ortholog', strandness=True)
 *                                       genomicranges.py:2040-2067 -> find_neighbours :635-685
 *     ortholog.calc_JI(range), ortholog.calc_OC(range)    genomicranges.py:497-547 (intersect :460-495)
 * for a whole list of orthologs in one call. Both lists are columnar; chromosomes and names are the
 * caller's ranks in ONE dictionary shared by both lists (rank order = Python's str order, None first),
 * strands are +1 '+', -1 '-', 0 '.'. The annotation must be sorted by (chrom, start, end, name) like the
 * reference's GenomicRangesList (_genomic_key_func :1566-1572). Per ortholog the neighbours come out in the
 * order of that sorted list, like `relations['ortholog']`.
 * ------------------------------------------------------------------------------------------- */
typedef struct o2a_ann_input {
    int64_t n_q, n_a;                                  /* orthologs, annotation ranges (both < 2^31) */
    const int32_t *q_chrom, *q_start, *q_end, *q_name; /* [n_q] */
    const int8_t  *q_strand;                           /* [n_q] */
    const int32_t *a_chrom, *a_start, *a_end, *a_name; /* [n_a], sorted */
    const int8_t  *a_strand;                           /* [n_a] */
    int64_t distance;                                  /* find_neighbours' distance (annotate_orthologs: 0) */
    int32_t strandness;                                /* find_neighbours' strandness (annotate_orthologs: 1) */
    int32_t mem;                                       /* O2A_MEM_HOST | O2A_MEM_DEVICE, applies to every pointer */
} o2a_ann_input;

int o2a_annotate_batch(o2a_ctx *ctx, const o2a_ann_input *in, int64_t *n_pairs);
/* any pointer may be NULL. nb_off[n_q+1]; per (ortholog, neighbour) pair: index into the sorted annotation,
 * Jaccard index, overlap coefficient (float64, the reference's values bit for bit) */
int o2a_fetch_annotation(o2a_ctx *ctx, int64_t *nb_off, int32_t *nb_idx, double *ji, double *oc);
/* CUDA-event durations (ms) of the last o2a_annotate_batch when timing is enabled: h2d, running max, count + scan, write */
int o2a_get_annotate_ms(o2a_ctx *ctx, float *ms /*4*/);

#ifdef __cplusplus
}
#endif
#endif /* O2A_H */
