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
int o2a_ingest_load(o2a_ingest *g, const char *const *align_paths, const char *const *bg_paths, int64_t n_files);

/* Totals over the files with status 0 (in input order; each such file is one lncRNA). */
int o2a_ingest_sizes(const o2a_ingest *g, int64_t *n_lnc, int64_t *n_aln, int64_t *n_hsp, int64_t *n_bg);

/* Copy out. Any pointer may be NULL. Sizes: file_status[n_files]; lnc_file[n_lnc] (file index of
 * each lncRNA); bg_off/aln_off[n_lnc+1]; bg[n_bg]; hsp_off[n_aln+1]; qlen/slen[n_aln];
 * coords[4*n_hsp] (AoS qstart,qend,sstart,send) and/or the four SoA arrays [n_hsp]; score[n_hsp];
 * score_is_int[n_hsp] (1 if the JSON number was an integer literal — the reference's record text
 * depends on it); spans[4*n_aln] = byte spans (qrange begin,end, srange begin,end) for o2a_ingest_span. */
int o2a_ingest_export(const o2a_ingest *g, int32_t *file_status, int64_t *lnc_file,
                      int64_t *bg_off, int32_t *bg, int64_t *aln_off, int64_t *hsp_off,
                      int64_t *qlen, int64_t *slen, int32_t *coords,
                      int32_t *qstart, int32_t *qend, int32_t *sstart, int32_t *send,
                      double *score, uint8_t *score_is_int, int64_t *spans);

/* The JSON text of a qrange / srange object of `file` (span from o2a_ingest_export): returns its
 * length; copies it when out != NULL and cap is large enough. */
int64_t o2a_ingest_span(const o2a_ingest *g, int64_t file, int64_t begin, int64_t end, char *out, int64_t cap);

const char *o2a_ingest_file_error(const o2a_ingest *g, int64_t file);

#ifdef __cplusplus
}
#endif
#endif /* O2A_INGEST_H */
