/*
 * o2a_oracle.c — CPU restatement of ortho2align's build_orthologs / get_best_orthologs hot path.
 *
 * TEST INFRASTRUCTURE ONLY. This file is the parity oracle for the CUDA library
 * (ortho2align_b200/csrc). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load it; the product path never does.
 *
 * Parity pinning: the reference ships no tests or golden vectors, so this
 * restatement is pinned against the reference ITSELF executed in the build container
 * (oracle/ref_harness.py, numpy 2.3.5 / scipy 1.18.1) — tests/test_oracle_vs_reference.py — and
 * against the golden vectors recorded from it under tests/golden/ (oracle/make_golden.py).
 * KDE: scipy.special.ndtr (xsf/cephes) is not bit-reproducible with libm erfc, so the KDE leg is
 * pinned to |dp| <= max(1e-12*p, 16*2^-53) only; everything else is bit-exact.
 *
 * Each function cites the reference file:line (paths relative to /root/reference/ortho2align)
 * or the third-party routine (numpy 2.3.5 / scipy 1.18.1) whose arithmetic it restates.
 *
 * Build: gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off -fno-fast-math (oracle/build.py)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define O2A_FIT_HIST 0
#define O2A_FIT_KDE 1
#define O2A_FIT_EMPIRICAL 2

/* ------------------------------------------------------------------------------------------
 * numpy pairwise summation (numpy/_core/src/umath/loops_utils.h.src, DOUBLE_pairwise_sum),
 * used by np.sum (normaliser, scipy rv_histogram.__init__) and ndarray.mean (fitting.py:295).
 * ---------------------------------------------------------------------------------------- */
static double pairwise_sum(const double *a, int64_t n)
{
    if (n < 8) {
        double res = 0.;
        for (int64_t i = 0; i < n; i++) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        int64_t i;
        for (int j = 0; j < 8; j++) r[j] = a[j];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return pairwise_sum(a, n2) + pairwise_sum(a + n2, n - n2);
    }
}

double o2a_oracle_pairwise_sum(const double *a, int64_t n) { return pairwise_sum(a, n); }

/* np.interp for one point (numpy/_core/src/multiarray/compiled_base.c: arr_interp +
 * binary_search_with_guess): j = the unique index with xp[j] <= x < xp[j+1]. */
static double np_interp(double x, const double *xp, const double *fp, int64_t n)
{
    if (isnan(x)) return x;
    if (x > xp[n - 1]) return fp[n - 1];
    if (x < xp[0]) return fp[0];
    int64_t lo = 0, hi = n; /* first index with xp[idx] > x */
    while (lo < hi) {
        int64_t mid = lo + ((hi - lo) >> 1);
        if (x >= xp[mid]) lo = mid + 1; else hi = mid;
    }
    int64_t j = lo - 1;
    if (j == n - 1) return fp[j];
    if (xp[j] == x) return fp[j];
    double slope = (fp[j + 1] - fp[j]) / (xp[j + 1] - xp[j]);
    double r = slope * (x - xp[j]) + fp[j];
    if (isnan(r)) {
        r = slope * (x - xp[j + 1]) + fp[j + 1];
        if (isnan(r) && fp[j] == fp[j + 1]) r = fp[j];
    }
    return r;
}

double o2a_oracle_interp(double x, const double *xp, const double *fp, int64_t n)
{
    return np_interp(x, xp, fp, n);
}

/* ------------------------------------------------------------------------------------------
 * HistogramFitter.__init__ (fitting.py:142-155):
 *   np.histogram(data, bins=len(data)//2, density=True)  -> numpy/lib/_histograms_impl.py
 *   scipy.stats.rv_histogram(hist)                        -> scipy/stats/_continuous_distns.py
 * hbins, hcdf: B+1 doubles each (hcdf already padded with the leading 0). Returns B, or <0.
 * ---------------------------------------------------------------------------------------- */
int64_t o2a_oracle_hist_fit(const int32_t *bg, int64_t n, double *hbins, double *hcdf)
{
    int64_t B = n / 2;
    if (B < 1) return -1;
    int32_t mn = bg[0], mx = bg[0];
    for (int64_t i = 1; i < n; i++) { if (bg[i] < mn) mn = bg[i]; if (bg[i] > mx) mx = bg[i]; }
    double first = (double)mn, last = (double)mx;
    if (mn == mx) { first -= 0.5; last += 0.5; }             /* _get_outer_edges */
    /* np.linspace(first, last, B+1): y = arange(B+1)*step + start; y[-1] = stop */
    double delta = last - first;
    double step = delta / (double)B;
    for (int64_t k = 0; k <= B; k++) {
        double y = (double)k * step;
        hbins[k] = y + first;
    }
    hbins[B] = last;
    for (int64_t k = 0; k < B; k++) if (!(hbins[k] < hbins[k + 1])) return -2; /* "Too many bins" */
    /* uniform-bin fast path of np.histogram */
    int64_t *cnt = (int64_t *)calloc((size_t)B, sizeof(int64_t));
    double *tmp = (double *)malloc((size_t)B * sizeof(double));
    if (!cnt || !tmp) { free(cnt); free(tmp); return -3; }
    double norm_denom = (mn == mx) ? 1.0 : (double)((uint64_t)((int64_t)mx - (int64_t)mn));
    for (int64_t i = 0; i < n; i++) {
        double x = (double)bg[i];
        double f = ((x - first) / norm_denom) * (double)B;
        int64_t idx = (int64_t)f;
        if (idx == B) idx -= 1;
        if (x < hbins[idx]) idx -= 1;
        if (x >= hbins[idx + 1] && idx != B - 1) idx += 1;
        cnt[idx] += 1;
    }
    /* density=True: n / db / n.sum() */
    double total = (double)n;
    /* rv_histogram: density=None is falsy and widths are constant -> hpdf = hpdf / widths
     * (second division); hpdf /= float(np.sum(hpdf*widths)); hcdf = cumsum(hpdf*widths) */
    for (int64_t k = 0; k < B; k++) {
        double w = hbins[k + 1] - hbins[k];
        double dens = ((double)cnt[k] / w) / total;
        double hp0 = dens / w;
        tmp[k] = hp0 * w;
    }
    double S = pairwise_sum(tmp, B);
    hcdf[0] = 0.0;
    double acc = 0.0;
    for (int64_t k = 0; k < B; k++) {
        double w = hbins[k + 1] - hbins[k];
        double dens = ((double)cnt[k] / w) / total;
        double hp0 = dens / w;
        double hpdf = hp0 / S;
        double c = hpdf * w;
        acc = acc + c;                   /* np.cumsum: sequential left-to-right */
        hcdf[k + 1] = acc;
    }
    free(cnt); free(tmp);
    return B;
}

/* rv_continuous.sf for rv_histogram (scipy/stats/_distn_infrastructure.py:sf;
 * _sf = 1.0 - _cdf; _cdf = np.interp(x, hbins, hcdf)) — fitting.py:184-193 */
double o2a_oracle_hist_sf(double x, const double *hbins, const double *hcdf, int64_t B)
{
    if (isnan(x)) return NAN;
    if (x <= hbins[0]) return 1.0;
    if (!(x < hbins[B])) return 0.0;
    return 1.0 - np_interp(x, hbins, hcdf, B + 1);
}

/* rv_continuous.isf -> _ppf(1.0 - q) = np.interp(1-q, hcdf, hbins) — fitting.py:206-215 */
double o2a_oracle_hist_isf(double q, const double *hbins, const double *hcdf, int64_t B)
{
    if (q == 1.0) return hbins[0];
    if (q == 0.0) return hbins[B];
    if (!(q > 0.0 && q < 1.0)) return NAN;
    return np_interp(1.0 - q, hcdf, hbins, B + 1);
}

/* ------------------------------------------------------------------------------------------
 * KernelFitter (fitting.py:245-356). cdf(x) = ndtr((x - data)/factor).mean()  (:295),
 * sf = 1 - cdf (:310), isf = brentq(sf(x) - q, min, max) with end-point fallback (:218-242,350).
 * ndtr here is 0.5*erfc(-z/sqrt2) with libm erfc (scipy's xsf ndtr is not bit-reproducible).
 * ---------------------------------------------------------------------------------------- */
static double ndtr_libm(double z) { return 0.5 * erfc(-z * 0.70710678118654752440); }

double o2a_oracle_kde_cdf(double x, const int32_t *bg, int64_t n, double factor, double *scratch)
{
    for (int64_t i = 0; i < n; i++) scratch[i] = ndtr_libm((x - (double)bg[i]) / factor);
    return pairwise_sum(scratch, n) / (double)n;
}

typedef struct { const int32_t *bg; int64_t n; double factor; double q; double *scratch; } kde_ctx;

static double kde_f(double x, kde_ctx *c)
{
    return (1.0 - o2a_oracle_kde_cdf(x, c->bg, c->n, c->factor, c->scratch)) - c->q;
}

/* scipy/optimize/Zeros/brentq.c, defaults of scipy.optimize.brentq: xtol=2e-12,
 * rtol=8.881784197001252e-16 (4*eps), maxiter=100. err: 0 ok, -1 sign error, -2 no convergence. */
static double brentq_kde(kde_ctx *c, double xa, double xb, int *err)
{
    const double xtol = 2e-12, rtol = 8.881784197001252e-16;
    const int iter = 100;
    double xpre = xa, xcur = xb, xblk = 0., fpre, fcur, fblk = 0., spre = 0., scur = 0., sbis;
    double delta, stry, dpre, dblk;
    *err = 0;
    fpre = kde_f(xpre, c);
    fcur = kde_f(xcur, c);
    if (fpre == 0) return xpre;
    if (fcur == 0) return xcur;
    if (signbit(fpre) == signbit(fcur)) { *err = -1; return 0.; }
    for (int i = 0; i < iter; i++) {
        if (fpre != 0 && fcur != 0 && (signbit(fpre) != signbit(fcur))) {
            xblk = xpre; fblk = fpre; spre = scur = xcur - xpre;
        }
        if (fabs(fblk) < fabs(fcur)) {
            xpre = xcur; xcur = xblk; xblk = xpre;
            fpre = fcur; fcur = fblk; fblk = fpre;
        }
        delta = (xtol + rtol * fabs(xcur)) / 2;
        sbis = (xblk - xcur) / 2;
        if (fcur == 0 || fabs(sbis) < delta) return xcur;
        if (fabs(spre) > delta && fabs(fcur) < fabs(fpre)) {
            if (xpre == xblk) {
                stry = -fcur * (xcur - xpre) / (fcur - fpre);
            } else {
                dpre = (fpre - fcur) / (xpre - xcur);
                dblk = (fblk - fcur) / (xblk - xcur);
                stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre));
            }
            double m1 = fabs(spre), m2 = 3 * fabs(sbis) - delta;
            if (2 * fabs(stry) < (m1 < m2 ? m1 : m2)) { spre = scur; scur = stry; }
            else { spre = sbis; scur = sbis; }
        } else { spre = sbis; scur = sbis; }
        xpre = xcur; fpre = fcur;
        if (fabs(scur) > delta) xcur += scur;
        else xcur += (sbis > 0 ? delta : -delta);
        fcur = kde_f(xcur, c);
    }
    *err = -2;
    return xcur;
}

double o2a_oracle_kde_isf(double q, const int32_t *bg, int64_t n, double factor, double *scratch, int *err)
{
    int32_t mn = bg[0], mx = bg[0];
    for (int64_t i = 1; i < n; i++) { if (bg[i] < mn) mn = bg[i]; if (bg[i] > mx) mx = bg[i]; }
    kde_ctx c = { bg, n, factor, q, scratch };
    double r = brentq_kde(&c, (double)mn, (double)mx, err);
    if (*err == -1) { /* ValueError branch of inverse_approximate_collection (fitting.py:240-241) */
        double fb = fabs(kde_f((double)mx, &c)), fa = fabs(kde_f((double)mn, &c));
        *err = 0;
        return fb < fa ? (double)mx : (double)mn;
    }
    return r;
}

/* ------------------------------------------------------------------------------------------
 * "empirical" fitter — EXTENSION, not in the reference.
 * sf(x) = #{bg > x}/n on the sorted background; isf(q) = sorted[ceil((1-q)*n) - 1].
 * Oracle for it is scipy.stats.ecdf (tests); parity unpinned by the reference.
 * ---------------------------------------------------------------------------------------- */
static int cmp_i32(const void *a, const void *b)
{
    int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return (x > y) - (x < y);
}

double o2a_oracle_emp_sf(double x, const int32_t *sorted, int64_t n)
{
    int64_t lo = 0, hi = n; /* first index with sorted[idx] > x */
    while (lo < hi) { int64_t mid = lo + ((hi - lo) >> 1); if ((double)sorted[mid] > x) hi = mid; else lo = mid + 1; }
    return (double)(n - lo) / (double)n;
}

double o2a_oracle_emp_isf(double q, const int32_t *sorted, int64_t n)
{
    int64_t k = (int64_t)ceil((1.0 - q) * (double)n) - 1;
    if (k < 0) k = 0;
    if (k > n - 1) k = n - 1;
    return (double)sorted[k];
}

/* ------------------------------------------------------------------------------------------
 * ranking (fitting.py:359-364) with the build's tie policy: stable (ties keep HSP order).
 * rank[i] = 1 + #{j: p[j] < p[i]} + #{j < i: p[j] == p[i]}
 * ---------------------------------------------------------------------------------------- */
typedef struct { double p; int64_t i; } pi_t;
static int cmp_pi(const void *a, const void *b)
{
    const pi_t *x = (const pi_t *)a, *y = (const pi_t *)b;
    if (x->p < y->p) return -1;
    if (x->p > y->p) return 1;
    return (x->i > y->i) - (x->i < y->i);
}

void o2a_oracle_ranking(const double *p, int64_t n, double *rank)
{
    pi_t *t = (pi_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(pi_t));
    for (int64_t i = 0; i < n; i++) { t[i].p = p[i]; t[i].i = i; }
    qsort(t, (size_t)n, sizeof(pi_t), cmp_pi);
    for (int64_t r = 0; r < n; r++) rank[t[r].i] = (double)(r + 1);
    free(t);
}

/* ------------------------------------------------------------------------------------------
 * Chaining — Alignment.best_chain (alignment.py:1032-1046) and everything below it, restated
 * the way the reference does it: O(n^2) graph build with the prune loop (alignment.py:875-927),
 * relax in list order with strict '>' (:388-411, :1004-1006), backtrack (:1007-1013).
 * Vertex arrays include begin (index 0) and end (index N-1).
 * ---------------------------------------------------------------------------------------- */
typedef struct { double qs, qe, ss, se, score; int64_t id; } vert_t;

static int precede(const vert_t *a, const vert_t *b, int direct)
{   /* HSP.precede, alignment.py:137-168 */
    /* quirk: begin/end sentinels have orientation None, so a sentinel-sentinel test falls to the
     * `else` (reverse) rule even inside the direct graph (alignment.py:162-165) */
    int dir_rule = direct && !(a->id < 0 && b->id < 0);
    int qp = a->qe < b->qs;
    int sp = dir_rule ? (a->se < b->ss) : (a->se > b->ss);
    return qp && sp;
}

static double edge_weight(const vert_t *a, const vert_t *b, int direct, double G, double E)
{   /* HSP.distance, alignment.py:170-217 (a precedes b) */
    double qd = b->qs - a->qe;
    double sd = direct ? (b->ss - a->se) : (a->se - b->ss);
    return (G + E * (qd + sd)) - b->score;
}

/* verts: N vertices sorted (begin, HSPs..., end). Returns chain length; chain = vertex indices
 * (1..N-2) in chain order; *score = -total[end]. */
static int64_t find_best_chain(const vert_t *v, int64_t N, int direct, double G, double E,
                               int64_t *chain, double *score)
{
    double *total = (double *)malloc((size_t)N * sizeof(double));
    int64_t *prev = (int64_t *)malloc((size_t)N * sizeof(int64_t));
    int64_t *succ = (int64_t *)malloc((size_t)N * sizeof(int64_t));
    for (int64_t i = 0; i < N; i++) { total[i] = (i == 0) ? 0.0 : INFINITY; prev[i] = -1; }
    /* graph build + prune + relax fused per vertex i in ascending order: identical to building
     * the whole graph first (edges of i depend on coordinates only) and relaxing in list order */
    for (int64_t i = 0; i < N; i++) {
        int64_t ns = 0;
        for (int64_t j = i + 1; j < N; j++) if (precede(&v[i], &v[j], direct)) succ[ns++] = j;
        /* prune loop (alignment.py:919-926): i stays 0, cut at first k with next[0] < next[k] */
        int64_t keep = ns;
        for (int64_t k = 1; k < ns; k++) if (precede(&v[succ[0]], &v[succ[k]], direct)) { keep = k; break; }
        for (int64_t e = 0; e < keep; e++) {
            int64_t j = succ[e];
            double w = edge_weight(&v[i], &v[j], direct, G, E);
            double nscore = total[i] + w;                    /* HSPVertex._relax */
            if (total[j] > nscore) { total[j] = nscore; prev[j] = i; }
        }
    }
    *score = -total[N - 1];
    /* backtrack (alignment.py:1007-1013): collect while best_prev is not None, drop end */
    int64_t len = 0, cur = N - 1;
    while (prev[cur] != -1) { chain[len++] = cur; cur = prev[cur]; }
    /* chain[0] is the end vertex: best_hsps[1:][::-1] */
    int64_t m = len > 0 ? len - 1 : 0;
    for (int64_t a = 0; a < m; a++) chain[a] = chain[a + 1];
    for (int64_t a = 0; a < m / 2; a++) { int64_t t = chain[a]; chain[a] = chain[m - 1 - a]; chain[m - 1 - a] = t; }
    free(total); free(prev); free(succ);
    return m;
}

static int cmp_vert(const void *a, const void *b)
{   /* sorted(key=(orientation, qstart, sstart)) is stable: ties keep list order (id) */
    const vert_t *x = (const vert_t *)a, *y = (const vert_t *)b;
    if (x->qs != y->qs) return x->qs < y->qs ? -1 : 1;
    if (x->ss != y->ss) return x->ss < y->ss ? -1 : 1;
    return (x->id > y->id) - (x->id < y->id);
}

/* best_chain for one alignment. idx: positions (into the alignment's HSP arrays) of the retained
 * HSPs in list order. Outputs chain positions (subset of idx, chain order).
 * Returns chain length, or -1 if an HSP has orientation None (reference raises, alignment.py:865-872).
 * scores2[0] = direct score, scores2[1] = reverse score (-inf when that orientation is empty). */
int64_t o2a_oracle_best_chain(int64_t n, const int64_t *idx,
                              const int32_t *qstart, const int32_t *qend,
                              const int32_t *sstart, const int32_t *send, const double *score,
                              int64_t qlen, int64_t slen, double G, double E,
                              int64_t *chain_out, int *orient_out, double *score_out, double *scores2)
{
    vert_t *vd = (vert_t *)malloc((size_t)(n + 2) * sizeof(vert_t));
    vert_t *vr = (vert_t *)malloc((size_t)(n + 2) * sizeof(vert_t));
    int64_t *cd = (int64_t *)malloc((size_t)(n + 2) * sizeof(int64_t));
    int64_t *cr = (int64_t *)malloc((size_t)(n + 2) * sizeof(int64_t));
    int64_t nd = 0, nr = 0, ret = 0;
    for (int64_t k = 0; k < n; k++) {
        int64_t h = idx ? idx[k] : k;
        int qstrand = (qend[h] - qstart[h]) > 0;            /* alignment.py:107-110 */
        int sstrand = (send[h] - sstart[h]) > 0;
        if (!qstrand) { ret = -1; goto done; }
        vert_t t = { (double)qstart[h], (double)qend[h], (double)sstart[h], (double)send[h], score[h], h };
        if (sstrand) vd[1 + nd++] = t; else vr[1 + nr++] = t;
    }
    {
        double sd = -INFINITY, sr = -INFINITY;
        int64_t ld = 0, lr = 0;
        if (nd > 0) {
            qsort(vd + 1, (size_t)nd, sizeof(vert_t), cmp_vert);
            vert_t b = { 0, 0, 0, 0, 0, -1 }, e = { (double)qlen, (double)qlen, (double)slen, (double)slen, 0, -1 };
            vd[0] = b; vd[nd + 1] = e;
            ld = find_best_chain(vd, nd + 2, 1, G, E, cd, &sd);
        }
        if (nr > 0) {
            qsort(vr + 1, (size_t)nr, sizeof(vert_t), cmp_vert);
            vert_t b = { 0, 0, (double)slen, (double)slen, 0, -1 }, e = { (double)qlen, (double)qlen, 0, 0, 0, -1 };
            vr[0] = b; vr[nr + 1] = e;
            lr = find_best_chain(vr, nr + 2, 0, G, E, cr, &sr);
        }
        scores2[0] = sd; scores2[1] = sr;
        if (sd > sr) {                                       /* alignment.py:1043-1046 */
            *orient_out = 0; *score_out = sd; ret = ld;
            for (int64_t a = 0; a < ld; a++) chain_out[a] = vd[cd[a]].id;
        } else {
            *orient_out = 1; *score_out = sr; ret = lr;
            for (int64_t a = 0; a < lr; a++) chain_out[a] = vr[cr[a]].id;
        }
    }
done:
    free(vd); free(vr); free(cd); free(cr);
    return ret;
}

/* ------------------------------------------------------------------------------------------
 * Whole batch: _build for every lncRNA (pipeline.py:755-796) + get_best_orthologs selection
 * (pipeline.py:959-1024) on the columnar layout of include/o2a.h. Output arrays use the same
 * "slot" layout as the CUDA library's device workspace: survivors / chain of alignment a live at
 * [hsp_off[a], hsp_off[a] + n_surv[a]) and [hsp_off[a], hsp_off[a] + chain_len[a]).
 * ---------------------------------------------------------------------------------------- */
int o2a_oracle_build_batch(
    int64_t n_lnc, const int64_t *bg_off, const int32_t *bg, const double *kde_factor,
    const int64_t *aln_off, const int64_t *hsp_off, const int64_t *qlen, const int64_t *slen,
    const int32_t *qstart, const int32_t *qend, const int32_t *sstart, const int32_t *send,
    const double *score,
    int fitter, int fdr, double alpha, int gapopen, int gapextend, int best_value, int best_fn,
    int n_threads,
    double *thr, int32_t *status, int64_t *best_aln,
    int32_t *n_surv, int32_t *n_keep, int32_t *chain_len, uint8_t *chain_orient, double *chain_score,
    double *orient_scores /* 2*n_aln */,
    int32_t *surv_idx, double *surv_p, double *surv_pq, uint8_t *surv_keep, int32_t *chain_idx)
{
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t l = 0; l < n_lnc; l++) {
        const int32_t *b = bg + bg_off[l];
        int64_t nb = bg_off[l + 1] - bg_off[l];
        int64_t B = nb / 2;
        double *hbins = NULL, *hcdf = NULL, *scratch = NULL;
        int32_t *sorted = NULL;
        int st = 0;
        double t = NAN;
        if (fitter == O2A_FIT_HIST) {
            hbins = (double *)malloc((size_t)(B + 2) * sizeof(double));
            hcdf = (double *)malloc((size_t)(B + 2) * sizeof(double));
            int64_t r = o2a_oracle_hist_fit(b, nb, hbins, hcdf);
            if (r < 0) st = 3; else t = o2a_oracle_hist_isf(alpha, hbins, hcdf, B);
        } else if (fitter == O2A_FIT_KDE) {
            scratch = (double *)malloc((size_t)nb * sizeof(double));
            int err = 0;
            t = o2a_oracle_kde_isf(alpha, b, nb, kde_factor[l], scratch, &err);
            if (err == -2) st = 2;
        } else {
            sorted = (int32_t *)malloc((size_t)nb * sizeof(int32_t));
            memcpy(sorted, b, (size_t)nb * sizeof(int32_t));
            qsort(sorted, (size_t)nb, sizeof(int32_t), cmp_i32);
            t = o2a_oracle_emp_isf(alpha, sorted, nb);
        }
        thr[l] = t;
        double best_key = 0; int64_t best = -1;
        for (int64_t a = aln_off[l]; a < aln_off[l + 1]; a++) {
            int64_t h0 = hsp_off[a], m = hsp_off[a + 1] - hsp_off[a];
            n_surv[a] = 0; n_keep[a] = 0; chain_len[a] = 0; chain_orient[a] = 1;
            chain_score[a] = -INFINITY; orient_scores[2 * a] = orient_scores[2 * a + 1] = -INFINITY;
            if (st) continue;
            /* filter_by_score: strict '>' (alignment.py:618-646, pipeline.py:765) */
            int64_t ns = 0;
            for (int64_t k = 0; k < m; k++) if (score[h0 + k] > t) surv_idx[h0 + ns++] = (int32_t)k;
            n_surv[a] = (int32_t)ns;
            for (int64_t k = 0; k < ns; k++) {               /* pipeline.py:766-767 */
                double x = score[h0 + surv_idx[h0 + k]];
                double p;
                if (fitter == O2A_FIT_HIST) p = o2a_oracle_hist_sf(x, hbins, hcdf, B);
                else if (fitter == O2A_FIT_KDE) p = 1.0 - o2a_oracle_kde_cdf(x, b, nb, kde_factor[l], scratch);
                else p = o2a_oracle_emp_sf(x, sorted, nb);
                surv_p[h0 + k] = p;
            }
            int64_t *ret = (int64_t *)malloc((size_t)(ns + 1) * sizeof(int64_t));
            int64_t nr = 0;
            if (fdr) {                                        /* pipeline.py:769-776 */
                double *rank = (double *)malloc((size_t)(ns + 1) * sizeof(double));
                o2a_oracle_ranking(surv_p + h0, ns, rank);
                for (int64_t k = 0; k < ns; k++) {
                    double q = ((double)m * surv_p[h0 + k]) / rank[k];
                    surv_pq[h0 + k] = q;
                    surv_keep[h0 + k] = (uint8_t)(q <= alpha);
                    if (q <= alpha) ret[nr++] = surv_idx[h0 + k];
                }
                free(rank);
            } else {                                          /* pipeline.py:778-780 */
                for (int64_t k = 0; k < ns; k++) { surv_pq[h0 + k] = surv_p[h0 + k]; surv_keep[h0 + k] = 1; ret[nr++] = surv_idx[h0 + k]; }
            }
            n_keep[a] = (int32_t)nr;
            int64_t *ch = (int64_t *)malloc((size_t)(nr + 2) * sizeof(int64_t));
            int orient = 1; double cs = -INFINITY;
            int64_t cl = o2a_oracle_best_chain(nr, ret, qstart + h0, qend + h0, sstart + h0, send + h0,
                                               score + h0, qlen[a], slen[a], (double)gapopen, (double)gapextend,
                                               ch, &orient, &cs, orient_scores + 2 * a);
            if (cl < 0) { st = 1; free(ret); free(ch); continue; }
            chain_len[a] = (int32_t)cl; chain_orient[a] = (uint8_t)orient; chain_score[a] = cs;
            for (int64_t k = 0; k < cl; k++) chain_idx[h0 + k] = (int32_t)ch[k];
            if (cl > 0) {
                /* get_best_orthologs key on the QUERY bed12 line (pipeline.py:981-987) */
                double key;
                int64_t blen = 0; int32_t lo = INT32_MAX, hi = INT32_MIN;
                /* weight = float(str(sum(scores))) (pipeline.py:986, genomicranges.py:1225) under the pinned interpreter
                 * (CPython 3.12 bltinmodule.c, builtin_sum): ints are added exactly while every item is an int; from the
                 * first float on, Neumaier-compensated addition. A score is taken for an int when it is integral (the
                 * columnar batch does not carry the JSON number type; the sums are equal either way below 2^53). */
                int64_t isum = 0; int in_int = 1; double fr = 0.0, cc = 0.0;
                for (int64_t k = 0; k < cl; k++) {
                    int64_t h = h0 + ch[k];
                    int32_t d = qend[h] - qstart[h]; blen += d < 0 ? -d : d;
                    if (qstart[h] < lo) lo = qstart[h];
                    if (qend[h] < lo) lo = qend[h];
                    if (qstart[h] > hi) hi = qstart[h];
                    if (qend[h] > hi) hi = qend[h];
                    {
                        double x = score[h];
                        if (in_int && x == (double)(int64_t)x && fabs(x) < 9007199254740992.0) isum += (int64_t)x;
                        else {
                            if (in_int) { in_int = 0; fr = (double)isum; cc = 0.0; }
                            double t = fr + x;
                            if (fabs(fr) >= fabs(x)) cc += (fr - t) + x; else cc += (x - t) + fr;
                            fr = t;
                        }
                    }
                }
                double w = in_int ? (double)isum : ((cc != 0.0 && isfinite(cc)) ? fr + cc : fr);
                if (best_value == 0) key = (double)blen;
                else if (best_value == 1) key = (double)((int64_t)hi - (int64_t)lo);
                else if (best_value == 2) key = (double)cl;
                else key = w;
                /* max()/min() keep the first extremal element (pipeline.py:1009-1014) */
                if (best < 0 || (best_fn == 0 ? key > best_key : key < best_key)) { best = a; best_key = key; }
            }
            free(ret); free(ch);
        }
        status[l] = st;
        if (st) {
            best = -1;
            for (int64_t a = aln_off[l]; a < aln_off[l + 1]; a++) {   /* the reference aborted this lncRNA */
                n_surv[a] = 0; n_keep[a] = 0; chain_len[a] = 0; chain_score[a] = -INFINITY;
                orient_scores[2 * a] = orient_scores[2 * a + 1] = -INFINITY;
            }
        }
        best_aln[l] = best;
        free(hbins); free(hcdf); free(scratch); free(sorted);
    }
    return 0;
}

int o2a_oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
