/*
 * o2a_blast.h — C ABI (in libo2a_ingest.so) of the native reader of BLAST tabular output,
 * `blastn -outfmt "7 std score"`, the stream ortho2align parses with
 *   Alignment.from_file_blast            alignment.py:515-586   (numberize: utils.py:6-41)
 * for every (query neighbourhood, syntenic region) pair in `_align_syntenies` (pipeline.py:525-528,
 * genomicranges.py:920-934). Output: the HSP arrays `o2a_cut_batch` (include/o2a.h) takes — 0-based,
 * half-open, sequence-relative coordinates — so the producer side never builds Python objects or JSON
 * (SURVEY.md 8f rank 2). Host-only C++; tables are parsed in parallel by `n_threads` std::threads.
 */
#ifndef O2A_BLAST_H
#define O2A_BLAST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct o2a_blast o2a_blast;

#define O2A_BLAST_OK 0
#define O2A_BLAST_ERROR 1        /* the reference raises (ValueError / UnboundLocalError / TypeError / IndexError):
                                    `_align_syntenies` turns it into an ExceptionLogger, pipeline.py:545-546 */
#define O2A_BLAST_UNSUPPORTED 2  /* the reference goes on but the columnar types cannot hold the value
                                    (non-integral or > int32 coordinate, non-numeric score) */

o2a_blast *o2a_blast_create(int n_threads /* <= 0: hardware concurrency */);
void o2a_blast_destroy(o2a_blast *b);

/* Parse n_tables text buffers (texts[i], lens[i] bytes; not NUL-terminated).
 * start_type: 1 = 1-based coordinates in the table (alignment.py:574-578); end_inclusive: 1 = ends are
 * inclusive (:579-581) — the reference's defaults, which `align_blast` uses. */
int o2a_blast_parse(o2a_blast *b, const char *const *texts, const int64_t *lens, int64_t n_tables,
                    int start_type, int end_inclusive);

int o2a_blast_sizes(const o2a_blast *b, int64_t *n_tables, int64_t *n_hsp);

/* Copy out; any pointer may be NULL. status[n_tables]; hsp_off[n_tables+1] (tables with status != 0 are
 * empty); per HSP: coordinates, score, score_is_int (numberize made an int of it); per table qlen / slen as
 * Alignment.__init__ sets them (alignment.py:477-487; from 'query length' / 'subject length' columns of the
 * LAST hit when the table has them, :583-584). */
int o2a_blast_export(const o2a_blast *b, int32_t *status, int64_t *hsp_off,
                     int32_t *qstart, int32_t *qend, int32_t *sstart, int32_t *send,
                     double *score, uint8_t *score_is_int, int64_t *qlen, int64_t *slen);

/* name of the exception the reference raises for table i ("" when status is 0), e.g. "ValueError" */
const char *o2a_blast_error(const o2a_blast *b, int64_t table);

/* ---------------------------------------------------------------------------------------------
 * Background collection: `_estimate_bg_for_single_query_blast/_seqfile`
 * (pipeline.py:223-233, 259-269) keep at most `observations` of the BLAST scores found for a query:
 *     if score_size > observations: scores = random.Random(); .seed(seed); .sample(scores, observations)
 * The algorithm lives in CPython, not in the reference tree: Lib/random.py (Random.sample,
 * _randbelow_with_getrandbits; unchanged 3.11 .. 3.13) over Modules/_randommodule.c (MT19937, seeding by
 * init_by_array over the 32-bit words of abs(seed), getrandbits(k) = genrand_uint32() >> (32 - k)).
 * o2a_bg_sample restates it: out receives min(n, observations) scores — all of them in order when
 * n <= observations, else the sample in the order `sample` returns it. Returns the count written, -1 on
 * invalid arguments.
 * ------------------------------------------------------------------------------------------- */
int64_t o2a_bg_sample(const int32_t *scores, int64_t n, int64_t observations, uint64_t seed, int32_t *out);

#ifdef __cplusplus
}
#endif
#endif /* O2A_BLAST_H */
