/*
 * o2a_oracle_producer.c — CPU restatement of the producer side of the alignment JSON
 * (SURVEY.md §8f rank 2): what `_align_syntenies` (pipeline.py:525-532) does to the HSPs of one
 * BLAST table between `from_file_blast` and `to_dict`.
 *
 * TEST INFRASTRUCTURE ONLY — same rules as o2a_oracle.c: only tests/, smoke() and the cpu_baseline
 * legs of the bench tools may load it; the product path never does.
 *
 * Parity pinning: the reference has no tests for this path; the restatement is pinned against the
 * reference itself run in the build container (tests/test_oracle_vs_reference.py, where /root/reference
 * exists) and against tests/golden/producer.json recorded from it (oracle/make_golden_producer.py).
 *
 * Restated (paths relative to /root/reference/ortho2align):
 *   GenomicRangesAlignment.to_genomic        genomicranges.py:1108-1129
 *   HSP.__init__ strands / orientation       alignment.py:84-85, 107-110 (evaluated by hsp.copy(), :258-265)
 *   Alignment._cut_hsps                      alignment.py:734-803
 *   HSPVertex._process_boundary              alignment.py:267-333  (Python round() = round-half-even;
 *                                            int/int true division = correctly rounded quotient)
 *   Alignment.__init__ qlen / slen defaults  alignment.py:477-487
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define O2A_NO_BOUND INT64_MIN
#define TWO53 9007199254740992.0

enum { OR_NONE = 0, OR_DIRECT = 1, OR_REVERSE = 2 };

typedef struct {
    int64_t qs, qe, ss, se;
    double score;
    int cut;       /* score was multiplied by a fraction at least once (-> a Python float) */
    int orient;    /* as computed when _cut_hsps copies the HSP, never refreshed afterwards */
    int inexact;   /* an intermediate left the range where double arithmetic equals Python's int arithmetic */
} hsp_t;

/* (a * b) / c + d, then round(): alignment.py:289-292, 297-300, 303-306, 317-320 */
static int64_t interpolate(int64_t a, int64_t b, int64_t c, int64_t d, int *inexact)
{
    double prod = (double)a * (double)b;
    if (fabs(prod) >= TWO53 || llabs(c) >= (1ll << 53) || llabs(d) >= (1ll << 53)) *inexact = 1;
    double v = (double)(a * b) / (double)c + (double)d;
    return (int64_t)rint(v);     /* default rounding mode: half to even, like Python's round(float) */
}

static double ratio(int64_t num, int64_t den) { return (double)num / (double)den; }

/* returns 1 if the HSP is kept */
static int cut_one(hsp_t *h, int64_t qleft, int64_t qright, int64_t sleft, int64_t sright)
{
    if (qleft != O2A_NO_BOUND) {                                     /* alignment.py:758-766 */
        if (h->qs < qleft && qleft <= h->qe) {
            double f = ratio(h->qe - qleft, h->qe - h->qs);          /* :288-293 */
            h->ss = interpolate(qleft - h->qs, h->se - h->ss, h->qe - h->qs, h->ss, &h->inexact);
            h->qs = qleft;
            h->score *= f; h->cut = 1;
        } else if (!(qleft <= h->qs)) return 0;
    }
    if (qright != O2A_NO_BOUND) {                                    /* :767-775 */
        if (h->qs <= qright && qright < h->qe) {
            double f = ratio(qright - h->qs, h->qe - h->qs);         /* :295-301 */
            h->se = interpolate(qright - h->qs, h->se - h->ss, h->qe - h->qs, h->ss, &h->inexact);
            h->qe = qright;
            h->score *= f; h->cut = 1;
        } else if (!(h->qe <= qright)) return 0;
    }
    if (sleft != O2A_NO_BOUND) {                                     /* :776-789 */
        int cut = 0;
        if (h->ss < sleft && sleft <= h->se) cut = 1;
        else if (sleft <= h->ss && h->orient == OR_DIRECT) cut = 0;
        else if (h->se < sleft && sleft <= h->ss) cut = 1;
        else if (sleft <= h->se && h->orient == OR_REVERSE) cut = 0;
        else return 0;
        if (cut) {                                                   /* :302-315 */
            int64_t qpoint = interpolate(sleft - h->ss, h->qe - h->qs, h->se - h->ss, h->qs, &h->inexact);
            double f;
            if (h->orient == OR_DIRECT) { f = ratio(h->se - sleft, h->se - h->ss); h->qs = qpoint; h->ss = sleft; }
            else { f = ratio(h->ss - sleft, h->ss - h->se); h->qe = qpoint; h->se = sleft; }
            h->score *= f; h->cut = 1;
        }
    }
    if (sright != O2A_NO_BOUND) {
        /* :790-802. QUIRK: this block builds new_hsps but never assigns it back (`return hsps`, :803), so the
         * right subject boundary drops nothing; HSPs that straddle it are still cut, in place. */
        int cut = 0;
        if (h->ss <= sright && sright < h->se) cut = 1;
        else if (h->se <= sright && h->orient == OR_DIRECT) cut = 0;
        else if (h->se <= sright && sright < h->ss) cut = 1;
        if (cut) {                                                   /* :316-329 */
            int64_t qpoint = interpolate(sright - h->ss, h->qe - h->qs, h->se - h->ss, h->qs, &h->inexact);
            double f;
            if (h->orient == OR_DIRECT) { f = ratio(sright - h->ss, h->se - h->ss); h->qe = qpoint; h->se = sright; }
            else { f = ratio(sright - h->se, h->ss - h->se); h->qs = qpoint; h->ss = sright; }
            h->score *= f; h->cut = 1;
        }
    }
    return 1;
}

static int fits32(int64_t v) { return v >= INT32_MIN && v <= INT32_MAX; }

/*
 * One call = every alignment of a batch: to_genomic (when relative[a], NULL = all), then
 * cut_coordinates(qleft[a], qright[a], sleft[a], sright[a]) (NULL array or O2A_NO_BOUND entry = None),
 * order-preserving. ocut[k]: bit 0 = the score was scaled by a cut, bit 1 = it was a float already.
 * Outputs are compacted: out_off[n_aln+1]; status[a] = 0 ok, 1 = a kept coordinate
 * does not fit int32, 2 = an intermediate beyond 2^53 (Python's exact integers would differ).
 */
int o2a_oracle_cut_batch(
    int64_t n_aln, const int64_t *hsp_off,
    const int32_t *qs, const int32_t *qe, const int32_t *ss, const int32_t *se, const double *score,
    const uint8_t *score_is_int /* NULL = every input score is a Python int */,
    const int64_t *q_start, const int64_t *q_end, const int8_t *q_minus,
    const int64_t *s_start, const int64_t *s_end, const int8_t *s_minus, const uint8_t *relative,
    const int64_t *qleft, const int64_t *qright, const int64_t *sleft, const int64_t *sright, int n_threads,
    int64_t *out_off, int32_t *oqs, int32_t *oqe, int32_t *oss, int32_t *ose, double *oscore, uint8_t *ocut,
    int64_t *qlen, int64_t *slen, int32_t *status)
{
    int64_t *cnt = (int64_t *)calloc((size_t)n_aln + 1, sizeof(int64_t));
    if (!cnt) return -1;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    /* pass 1: every alignment writes its kept HSPs at the start of its own input slot */
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t a = 0; a < n_aln; ++a) {
        const int64_t h0 = hsp_off[a], h1 = hsp_off[a + 1];
        const int rel = relative ? relative[a] : 1;
        const int64_t ql = qleft ? qleft[a] : O2A_NO_BOUND, qr = qright ? qright[a] : O2A_NO_BOUND;
        const int64_t sl = sleft ? sleft[a] : O2A_NO_BOUND, sr = sright ? sright[a] : O2A_NO_BOUND;
        int64_t k = h0, mq = 0, ms = 0;
        int st = 0, any = 0;
        for (int64_t i = h0; i < h1; ++i) {
            hsp_t h = {qs[i], qe[i], ss[i], se[i], score[i], 0, OR_NONE, 0};
            if (rel) {                                               /* genomicranges.py:1112-1128 */
                if (!q_minus[a]) { h.qs += q_start[a]; h.qe += q_start[a]; }
                else { h.qs = q_end[a] - h.qs; h.qe = q_end[a] - h.qe; }
                if (!s_minus[a]) { h.ss += s_start[a]; h.se += s_start[a]; }
                else { h.ss = s_end[a] - h.ss; h.se = s_end[a] - h.se; }
            }
            const int qstrand = h.qe - h.qs > 0, sstrand = h.se - h.ss > 0;     /* alignment.py:107-110 */
            h.orient = qstrand ? (sstrand ? OR_DIRECT : OR_REVERSE) : OR_NONE;
            if (!cut_one(&h, ql, qr, sl, sr)) continue;
            if (h.inexact) st = 2;
            if (!(fits32(h.qs) && fits32(h.qe) && fits32(h.ss) && fits32(h.se))) { if (!st) st = 1; }
            oqs[k] = (int32_t)h.qs; oqe[k] = (int32_t)h.qe; oss[k] = (int32_t)h.ss; ose[k] = (int32_t)h.se;
            /* bit 0: scaled by a cut; bit 1: already a float in the table -> the JSON number is an int iff 0 */
            oscore[k] = h.score; ocut[k] = (uint8_t)(h.cut | ((score_is_int && !score_is_int[i]) ? 2 : 0));
            if (!any || h.qe > mq) mq = h.qe;
            int64_t m = h.ss > h.se ? h.ss : h.se;
            if (!any || m > ms) ms = m;
            any = 1;
            ++k;
        }
        cnt[a] = k - h0;
        qlen[a] = any ? mq + 1 : 1;                                  /* alignment.py:477-487 (default HSPVertex(0,0,0,0,0)) */
        slen[a] = any ? ms + 1 : 1;
        status[a] = st;
    }
    /* pass 2: close the gaps, in order */
    int64_t o = 0;
    for (int64_t a = 0; a < n_aln; ++a) {
        const int64_t h0 = hsp_off[a], c = cnt[a];
        out_off[a] = o;
        if (o != h0 && c > 0) {
            memmove(oqs + o, oqs + h0, (size_t)c * 4); memmove(oqe + o, oqe + h0, (size_t)c * 4);
            memmove(oss + o, oss + h0, (size_t)c * 4); memmove(ose + o, ose + h0, (size_t)c * 4);
            memmove(oscore + o, oscore + h0, (size_t)c * 8); memmove(ocut + o, ocut + h0, (size_t)c);
        }
        o += c;
    }
    out_off[n_aln] = o;
    free(cnt);
    return 0;
}
