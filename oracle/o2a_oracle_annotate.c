/* o2a_oracle_annotate.c — CPU restatement of annotate_orthologs' interval join. TEST INFRASTRUCTURE ONLY:
 * imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg, never by the product.
 *
 * Follows the reference literally, loop for loop (files under /root/reference/ortho2align/):
 *   GenomicRangesList.get_neighbours      genomicranges.py:2040-2067  (one find_neighbours per ortholog)
 *   BaseGenomicRange.find_neighbours      genomicranges.py:635-685    (bisect_left, left loop to the chromosome
 *                                                                      boundary, right loop to the first miss,
 *                                                                      strand filter, re-sorted list)
 *   BaseGenomicRange.distance             genomicranges.py:439-458
 *   BaseGenomicRange.intersect            genomicranges.py:460-495
 *   BaseGenomicRange.calc_JI / calc_OC    genomicranges.py:497-521 / 523-547
 *   sort key (chrom, start, end, name)    genomicranges.py:1566-1572
 * The left loop is kept O(chromosome) on purpose: this file is the checker (and the CPU baseline), not a design.
 *
 * Pinned against the reference itself executed in the build container: tests/test_oracle_vs_reference.py
 * (random lists through GenomicRangesList.get_neighbours / calc_JI / calc_OC) and the committed
 * tests/golden/annotate/ files recorded by oracle/make_golden_annotate.py. The reference has no tests or golden
 * vectors of its own for this path.
 *
 * One observable-free difference: the reference rebuilds the neighbour list as a SortedKeyList, so left-side
 * neighbours with IDENTICAL keys (chrom, start, end, name) come out in reversed list order. Name, length, JI and
 * OC — everything annotate_orthologs prints — are functions of the key, so this restatement keeps list order.
 */
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    const int32_t *chrom, *start, *end, *name;
    const int8_t *strand;
} ranges_t;

/* BaseGenomicRange.distance (:451-458); returns 1 in *err for InconsistentChromosomesError */
static int64_t gr_distance(int32_t c0, int32_t s0, int32_t e0, int32_t c1, int32_t s1, int32_t e1, int *err)
{
    *err = 0;
    if (c0 != c1) { *err = 1; return 0; }
    if (e0 < s1) return (int64_t)s1 - e0;
    if (e1 < s0) return (int64_t)e1 - s0;
    return 0;
}

/* calc_JI (:511-521) and calc_OC (:537-547) through intersect (:479-495) */
static void gr_ji_oc(int32_t s0, int32_t e0, int32_t s1, int32_t e1, int64_t dist, double *ji, double *oc)
{
    if (dist != 0) { *ji = 0.0; *oc = 0.0; return; }                 /* intersect() returned None -> 0 */
    int64_t is = (s0 < s1 && s1 < e0) ? s1 : s0;                     /* :480-483 */
    int64_t ie = (s0 < e1 && e1 < e0) ? e1 : e0;                     /* :484-487 */
    int64_t li = ie - is, l0 = (int64_t)e0 - s0, l1 = (int64_t)e1 - s1;
    int64_t uni = l0 + l1 - li, mn = l0 < l1 ? l0 : l1;
    *ji = uni == 0 ? 0.0 : (double)li / (double)uni;                 /* ZeroDivisionError -> 0 */
    *oc = mn == 0 ? 0.0 : (double)li / (double)mn;
}

static int key_less(const ranges_t *a, int64_t k, int32_t c, int32_t s, int32_t e, int32_t n)
{   /* tuple comparison (chrom, start, end, name) < (c, s, e, n) */
    if (a->chrom[k] != c) return a->chrom[k] < c;
    if (a->start[k] != s) return a->start[k] < s;
    if (a->end[k] != e) return a->end[k] < e;
    return a->name[k] < n;
}

static int strand_ok(int strandness, int qs, int as)
{   /* :670-677 */
    return !strandness || qs == 0 || as == qs || as == 0;
}

/* One ortholog. If out_idx is NULL only counts. Returns the number of neighbours. */
static int64_t one_query(const ranges_t *q, int64_t i, const ranges_t *a, int64_t n_a, int64_t distance, int strandness,
                         int32_t *out_idx, double *out_ji, double *out_oc)
{
    const int32_t c = q->chrom[i], s = q->start[i], e = q->end[i], nm = q->name[i];
    const int qs = q->strand[i];
    int64_t lo = 0, hi = n_a;                                        /* bisect_left (:657) */
    while (lo < hi) {
        int64_t mid = (lo + hi) / 2;
        if (key_less(a, mid, c, s, e, nm)) lo = mid + 1; else hi = mid;
    }
    const int64_t index = lo;
    int64_t first_left = index;                                      /* left loop (:660-668): down to the chromosome boundary */
    for (int64_t k = index - 1; k >= 0; --k) {
        int err;
        gr_distance(c, s, e, a->chrom[k], a->start[k], a->end[k], &err);
        if (err) break;
        first_left = k;
    }
    int64_t n = 0;
    for (int64_t k = first_left; k < index; ++k) {                   /* ascending = the order of the re-sorted list */
        int err;
        int64_t d = gr_distance(c, s, e, a->chrom[k], a->start[k], a->end[k], &err);
        if ((d < 0 ? -d : d) <= distance && strand_ok(strandness, qs, a->strand[k])) {
            if (out_idx) { out_idx[n] = (int32_t)k; gr_ji_oc(s, e, a->start[k], a->end[k], d, &out_ji[n], &out_oc[n]); }
            ++n;
        }
    }
    for (int64_t k = index; k < n_a; ++k) {                          /* right loop (:669-676): stops at the first miss */
        int err;
        int64_t d = gr_distance(c, s, e, a->chrom[k], a->start[k], a->end[k], &err);
        if (err || (d < 0 ? -d : d) > distance) break;
        if (strand_ok(strandness, qs, a->strand[k])) {
            if (out_idx) { out_idx[n] = (int32_t)k; gr_ji_oc(s, e, a->start[k], a->end[k], d, &out_ji[n], &out_oc[n]); }
            ++n;
        }
    }
    return n;
}

/* Pass 1 (nb_idx == NULL): nb_off[0..n_q] = exclusive scan of the neighbour counts.
 * Pass 2 (nb_idx != NULL): nb_off as produced by pass 1; fills nb_idx / ji / oc. */
int32_t o2a_oracle_annotate(int64_t n_q, const int32_t *q_chrom, const int32_t *q_start, const int32_t *q_end,
                            const int32_t *q_name, const int8_t *q_strand,
                            int64_t n_a, const int32_t *a_chrom, const int32_t *a_start, const int32_t *a_end,
                            const int32_t *a_name, const int8_t *a_strand,
                            int64_t distance, int32_t strandness, int32_t n_threads,
                            int64_t *nb_off, int32_t *nb_idx, double *ji, double *oc)
{
    ranges_t q = {q_chrom, q_start, q_end, q_name, q_strand}, a = {a_chrom, a_start, a_end, a_name, a_strand};
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#else
    n_threads = 1;
#endif
    if (!nb_idx) {
        int64_t *cnt = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_q + 1));
        if (!cnt) return -1;
#pragma omp parallel for schedule(dynamic, 64) num_threads(n_threads)
        for (int64_t i = 0; i < n_q; ++i) cnt[i] = one_query(&q, i, &a, n_a, distance, strandness, 0, 0, 0);
        int64_t acc = 0;
        for (int64_t i = 0; i < n_q; ++i) { nb_off[i] = acc; acc += cnt[i]; }
        nb_off[n_q] = acc;
        free(cnt);
        return 0;
    }
#pragma omp parallel for schedule(dynamic, 64) num_threads(n_threads)
    for (int64_t i = 0; i < n_q; ++i)
        one_query(&q, i, &a, n_a, distance, strandness, nb_idx + nb_off[i], ji + nb_off[i], oc + nb_off[i]);
    return 0;
}
