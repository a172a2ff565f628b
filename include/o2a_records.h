/*
 * o2a_records.h — C ABI (libo2a_ingest.so, host-only C++) of the record formatter: the text of the four
 * `significant.*_orthologs.{bed,tsv}` files for a batch of chains, byte for byte what the reference writes.
 *
 * Replaces, per non-empty chain, `chain.to_record(mode='str')` / `chain.to_bed12(mode='str')`:
 *   GenomicRangesAlignmentChain._prepare_side / to_record / to_bed12   genomicranges.py:1192-1301
 *   nxor_strands                                                        alignment.py:48-64
 *   the file loops of build_orthologs                                   pipeline.py:923-932
 * Numbers are rendered like Python's `str()`: ints in decimal; floats (a score that was a JSON float or became one in
 * the left-to-right `sum` of genomicranges.py:1225, and every p/q-value) as `repr(float)` — shortest round-trip
 * digits, exponent form when the decimal point falls at or before 10^-4 or beyond 10^16, '.0' appended to integers.
 */
#ifndef O2A_RECORDS_H
#define O2A_RECORDS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct o2a_records_input {
    int64_t n_chain;
    const int64_t *hsp_off;                        /* [n_chain+1] into the per-HSP arrays below, HSPs in chain order */
    const int32_t *qstart, *qend, *sstart, *send;  /* genomic coordinates */
    const double  *score;
    const uint8_t *score_is_int;                   /* 1 = the JSON number was an integer literal (a Python int) */
    const double  *pval;                           /* hsp.pval as assigned by _build (q or p), pipeline.py:772, 779 */
    const uint8_t *qstrand_plus;                   /* [n_chain] 1 if qrange.strand in ('+', '.') (nxor_strands' test) */
    const int64_t *str_off;                        /* [3*n_chain+1] offsets into str_blob: str(qrange.chrom), */
    const char    *str_blob;                       /*   str(name), str(srange.chrom) of every chain, UTF-8 */
} o2a_records_input;

/* out[4] / out_len[4]: query TSV, query BED12, subject TSV, subject BED12 — one '\n'-terminated line per chain, in
 * input order. The buffers are malloc'ed by the library: release each with o2a_records_free. Returns 0, or -1 for
 * invalid arguments / out of memory. n_threads <= 0: hardware concurrency. */
int o2a_records_format(const o2a_records_input *in, int n_threads, char **out, int64_t *out_len);
void o2a_records_free(char *p);
/* repr(float) of one value into buf (at least 32 bytes); returns its length. Exposed for the tests. */
int o2a_records_repr_double(double x, char *buf);

/*
 * get_best_orthologs on files (pipeline.py:959-1024): the four significant.* files are read in lock-step, consecutive
 * records with the same query name (column 4 of the query BED12 line) form a group, and the FIRST record with the
 * maximal / minimal key(query BED12 line) of each group is written to the four output files.
 *   value    0 block_length  = sum(int(i) for i in col11.split(','))     1 total_length = int(col3) - int(col2)
 *            2 block_count   = int(col10)                                3 weight       = float(col5)
 *   function 0 max, 1 min
 * Returns 0; -1 on an I/O error; 1 when the input holds anything Python would treat differently from this reader
 * (carriage returns, short records, numbers that are not plain decimal literals, NaN): nothing has been written then and
 * the caller runs the text-level restatement, which raises or decides exactly like the reference.
 */
int o2a_best_orthologs_files(const char *q_bed, const char *q_tsv, const char *s_bed, const char *s_tsv,
                             int value, int function, const char *out_q_bed, const char *out_q_tsv,
                             const char *out_s_bed, const char *out_s_tsv, int64_t *n_records, int64_t *n_groups);

#ifdef __cplusplus
}
#endif
#endif /* O2A_RECORDS_H */
