/*
 * o2a_ingest.h — C ABI of libo2a_ingest.so: native ingest of ortho2align's per-lncRNA JSON files
 * (align_files/<name>.json + bg_files/<name>.json) into the columnar batch of include/o2a.h.
 *
 * Replaces, for the build_orthologs path, the ingest half of `_build`:
 *   json.load + GenomicRangesAlignment.from_dict   pipeline.py:743-745, genomicranges.py:1084-1106
 *   HSPVertex.from_dict / HSP.__init__             alignment.py:89-110, 233-249
 *   background load + degenerate-set patches       pipeline.py:746-754
 * (94 % of the reference's build_orthologs time.)
 *
 * Host-only C++ (no CUDA). Files are parsed in parallel by `n_threads` std::threads.
 */
#ifndef O2A_INGEST_H
#define O2A_INGEST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct o2a_ingest o2a_ingest;

o2a_ingest *o2a_ingest_create(int n_threads /* <= 0: hardware concurrency */);
void o2a_ingest_destroy(o2a_ingest *g);

/* Parse n_files (alignment file, background file) pairs. A missing background file is the
 * reference's `[0, 1]` case; a missing/malformed alignment file gets file status 1; values outside
 * the int32 columnar types get status 2. Always returns 0 unless the arguments are invalid. */
/* Size and modification time (ns) of n files, stat'ed on n_threads threads (<= 0: hardware concurrency); size -1 =
 * missing. Guards the columnar sidecar (ortho2align_b200/ingest.py::_manifest); no reference counterpart. */
int o2a_stat_files(const char *const *paths, int64_t n, int n_threads, int64_t *size, int64_t *mtime_ns);

int o2a_ingest_load(o2a_ingest *g, const char *const *align_paths, const char *const *bg_paths, int64_t n_files);

/* Totals over the files with status 0 (in input order; each such file is one lncRNA). */
int o2a_ingest_sizes(const o2a_ingest *g, int64_t *n_lnc, int64_t *n_aln, int64_t *n_hsp, int64_t *n_bg);

/* Copy out. Any pointer may be NULL. Sizes: file_status[n_files]; lnc_file[n_lnc] (file index of
 * each lncRNA); bg_off/aln_off[n_lnc+1]; bg[n_bg]; hsp_off[n_aln+1]; qlen/slen[n_aln];
 * coords[4*n_hsp] (AoS qstart,qend,sstart,send) and/or the four SoA arrays [n_hsp]; score[n_hsp];
 * score_is_int[n_hsp] (1 if the JSON number was an integer literal — the reference's record text
 * depends on it); spans[4*n_aln] = byte spans (qrange begin,end, srange begin,end) for o2a_ingest_span;
 * relative[n_aln] = 1 where the alignment's 'relative' flag is truthy: GenomicRangesAlignmentChain.to_record
 * raises ValueError for those (genomicranges.py:1259), so build_orthologs reports them as dropped. */
int o2a_ingest_export(const o2a_ingest *g, int32_t *file_status, int64_t *lnc_file,
                      int64_t *bg_off, int32_t *bg, int64_t *aln_off, int64_t *hsp_off,
                      int64_t *qlen, int64_t *slen, int32_t *coords,
                      int32_t *qstart, int32_t *qend, int32_t *sstart, int32_t *send,
                      double *score, uint8_t *score_is_int, int64_t *spans, uint8_t *relative);

/*
 * The same batch in the PCIe-optimal LOSSLESS host layout o2a_build_batch accepts (include/o2a.h: `score_q` +
 * `exc_*`, narrow `bg`, AoS `coords` = the `coords` array of o2a_ingest_export), written by the parser's thread
 * pool straight into caller-provided buffers — o2a_host_alloc'ed (pinned) ones in the product path, so that
 * what the ingest returns is what crosses PCIe.
 *   o2a_ingest_narrow_plan   picks the widths: score_bits = the narrowest of 8 / 16 / 32 that sends at most 10 %
 *                            of the HSPs to the exception lists (fractional scores of cut HSPs, alignment.py:330,
 *                            and values outside the type), bg_bits = the narrowest that holds every background
 *                            score; *n_exc = exception entries at that score width.
 *   o2a_ingest_export_narrow score_q[n_hsp] (int8 / int16 / int32; the most negative value marks an exception),
 *                            exc_off[n_aln+1], exc_pos / exc_val[n_exc] (position inside the alignment, exact
 *                            float64), bg_q[n_bg] (int8 / int16 / int32).
 */
int o2a_ingest_narrow_plan(const o2a_ingest *g, int32_t *score_bits, int32_t *bg_bits, int64_t *n_exc);
int o2a_ingest_export_narrow(const o2a_ingest *g, int32_t score_bits, int32_t bg_bits, void *score_q,
                             int64_t *exc_off, int32_t *exc_pos, double *exc_val, void *bg_q);

/* The same encoding for callers that already hold columnar arrays (lncRNA shards cut for several GPUs, the
 * producer side, bench.py's synthetic batches). o2a_ingest_pack_plan: widths as above, n_exc[3] = exception
 * entries at 8 / 16 / 32 bits. o2a_ingest_pack_narrow: any of the three groups (score_q + exc_*, bg_q, coords)
 * may be skipped by passing NULL for its output. n_threads <= 0: hardware concurrency. */
int o2a_ingest_pack_plan(const double *score, int64_t n_hsp, const int32_t *bg, int64_t n_bg, int n_threads,
                         int32_t *score_bits, int32_t *bg_bits, int64_t *n_exc);
int o2a_ingest_pack_narrow(int64_t n_aln, const int64_t *hsp_off, const double *score, int32_t score_bits,
                           void *score_q, int64_t *exc_off, int32_t *exc_pos, double *exc_val,
                           const int32_t *bg, int64_t n_bg, int32_t bg_bits, void *bg_q,
                           const int32_t *qstart, const int32_t *qend, const int32_t *sstart, const int32_t *send,
                           int32_t *coords, int n_threads);

/* Every qrange / srange text of the batch in one call (alignment order q0, s0, q1, s1, ...): range_off[2*n_aln+1]
 * and, when blob != NULL and cap is large enough, the concatenated bytes. Returns the total size. */
int64_t o2a_ingest_range_texts(const o2a_ingest *g, int64_t *range_off, char *blob, int64_t cap);

/* What the record formatter (o2a_records.h) needs from the range objects, decoded natively: per alignment the three
 * strings str(qrange.chrom), str(qrange.name) (with GenomicRange.name's fall-back to "chrom:start-endstrand",
 * genomicranges.py:427-433) and str(srange.chrom) as name_off[3*n_aln+1] offsets into blob (UTF-8), q_plus[n_aln] =
 * qrange.strand in ('+', '.') (nxor_strands' test, alignment.py:48-64), simple[n_aln] = 1 when every value was a JSON
 * string / null / int as json.dump writes them (0: decode the object with a full JSON reader). Returns the blob size. */
int64_t o2a_ingest_range_fields(const o2a_ingest *g, int64_t *name_off, char *blob, int64_t cap, uint8_t *q_plus,
                                uint8_t *simple);

/* The JSON text of a qrange / srange object of `file` (span from o2a_ingest_export): returns its
 * length; copies it when out != NULL and cap is large enough. */
int64_t o2a_ingest_span(const o2a_ingest *g, int64_t file, int64_t begin, int64_t end, char *out, int64_t cap);

const char *o2a_ingest_file_error(const o2a_ingest *g, int64_t file);

#ifdef __cplusplus
}
#endif
#endif /* O2A_INGEST_H */
