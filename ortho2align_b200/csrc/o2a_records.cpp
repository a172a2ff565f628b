// o2a_records.cpp — include/o2a_records.h: BED12 / TSV text of the ortholog chains, byte-identical to the reference's
// str()-based formatting (genomicranges.py:1192-1301). Host-only C++17, compiled into libo2a_ingest.so.
#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/o2a_records.h"

namespace {

// repr(float) (CPython float_repr_style 'short': PyOS_double_to_string(x, 'r', 0, Py_DTSF_ADD_DOT_0)):
// shortest digits that round-trip; exponent form iff decpt <= -4 or decpt > 16, where x = 0.DIGITS * 10^decpt.
int repr_double(double x, char* out)
{
    if (std::isnan(x)) { memcpy(out, "nan", 3); return 3; }
    if (std::isinf(x)) { if (x < 0) { memcpy(out, "-inf", 4); return 4; } memcpy(out, "inf", 3); return 3; }
    char* p = out;
    if (std::signbit(x)) { *p++ = '-'; x = -x; }
    if (x == 0.0) { memcpy(p, "0.0", 3); return (int)(p - out) + 3; }
    char tmp[40];
    auto r = std::to_chars(tmp, tmp + sizeof(tmp), x, std::chars_format::scientific);   // d[.ddd]e[+-]XX, shortest
    char digits[24];
    int nd = 0, i = 0;
    const int len = (int)(r.ptr - tmp);
    for (; i < len && tmp[i] != 'e'; ++i)
        if (tmp[i] != '.') digits[nd++] = tmp[i];
    int e10 = 0;
    {
        int j = i + 1;
        bool neg = false;
        if (tmp[j] == '+') ++j; else if (tmp[j] == '-') { neg = true; ++j; }
        for (; j < len; ++j) e10 = e10 * 10 + (tmp[j] - '0');
        if (neg) e10 = -e10;
    }
    while (nd > 1 && digits[nd - 1] == '0') --nd;          // to_chars never pads, but be safe
    const int decpt = e10 + 1;
    if (decpt <= -4 || decpt > 16) {
        *p++ = digits[0];
        if (nd > 1) { *p++ = '.'; memcpy(p, digits + 1, nd - 1); p += nd - 1; }
        *p++ = 'e';
        int e = decpt - 1;
        if (e < 0) { *p++ = '-'; e = -e; } else *p++ = '+';
        char eb[8]; int ne = 0;
        do { eb[ne++] = (char)('0' + e % 10); e /= 10; } while (e);
        if (ne < 2) eb[ne++] = '0';
        while (ne) *p++ = eb[--ne];
    } else if (decpt <= 0) {
        *p++ = '0'; *p++ = '.';
        for (int k = 0; k < -decpt; ++k) *p++ = '0';
        memcpy(p, digits, nd); p += nd;
    } else if (decpt >= nd) {
        memcpy(p, digits, nd); p += nd;
        for (int k = 0; k < decpt - nd; ++k) *p++ = '0';
        *p++ = '.'; *p++ = '0';
    } else {
        memcpy(p, digits, decpt); p += decpt;
        *p++ = '.';
        memcpy(p, digits + decpt, nd - decpt); p += nd - decpt;
    }
    return (int)(p - out);
}

inline void put_int(std::string& s, long long v)
{
    char b[24];
    auto r = std::to_chars(b, b + sizeof(b), v);
    s.append(b, r.ptr);
}

inline void put_double(std::string& s, double v)
{
    char b[40];
    s.append(b, (size_t)repr_double(v, b));
}

struct Out { std::string f[4]; };

// One chain -> four lines (chain_records in ortho2align_b200/records.py is the Python statement of the same thing).
void format_chain(const o2a_records_input& in, int64_t c, Out& o, std::vector<long long>& sizes, std::vector<long long>& starts)
{
    const int64_t h0 = in.hsp_off[c], h1 = in.hsp_off[c + 1], n = h1 - h0;
    const char* qchrom = in.str_blob + in.str_off[3 * c];     const size_t qchrom_n = (size_t)(in.str_off[3 * c + 1] - in.str_off[3 * c]);
    const char* name = in.str_blob + in.str_off[3 * c + 1];   const size_t name_n = (size_t)(in.str_off[3 * c + 2] - in.str_off[3 * c + 1]);
    const char* schrom = in.str_blob + in.str_off[3 * c + 2]; const size_t schrom_n = (size_t)(in.str_off[3 * c + 3] - in.str_off[3 * c + 2]);
    // score = sum([hsp.score for hsp in HSPs]) (genomicranges.py:1225) as CPython >= 3.12 evaluates it (Python/bltinmodule.c,
    // builtin_sum_impl): ints add exactly; the first float makes the accumulator int + float; from then on floats are added
    // with Neumaier's compensated summation, ints plainly, and the compensation is folded in at the end.
    long long isum = 0; double fsum = 0.0, comp = 0.0; bool is_float = false;
    for (int64_t h = h0; h < h1; ++h) {
        const double x = in.score[h];
        if (!is_float) {
            if (in.score_is_int[h]) isum += (long long)x;
            else { is_float = true; fsum = (double)isum + x; }
        } else if (in.score_is_int[h]) fsum += x;
        else {
            const double t = fsum + x;
            if (std::fabs(fsum) >= std::fabs(x)) comp += (fsum - t) + x; else comp += (x - t) + fsum;
            fsum = t;
        }
    }
    if (is_float && comp != 0.0 && std::isfinite(comp)) fsum += comp;
    const bool qflag = in.qend[h0] - (long long)in.qstart[h0] > 0, sflag = in.send[h0] - (long long)in.sstart[h0] > 0;
    const bool qplus = in.qstrand_plus[c] != 0;
    for (int side = 0; side < 2; ++side) {
        const int32_t* st = side == 0 ? in.qstart : in.sstart;
        const int32_t* en = side == 0 ? in.qend : in.send;
        long long c0 = st[h0], c1 = st[h0];
        for (int64_t h = h0; h < h1; ++h) {
            c0 = std::min(c0, (long long)std::min(st[h], en[h]));
            c1 = std::max(c1, (long long)std::max(st[h], en[h]));
        }
        const char strand = ((side == 0 ? qflag : sflag) == qplus) ? '+' : '-';       // nxor_strands
        sizes.clear(); starts.clear();
        const bool rev = side == 1 && strand == '-';             // subject side on '-': send-based starts, lists reversed
        for (int64_t h = h0; h < h1; ++h) {
            sizes.push_back(std::llabs((long long)en[h] - st[h]));
            starts.push_back((rev ? (long long)en[h] : (long long)st[h]) - c0);
        }
        if (rev) { std::reverse(sizes.begin(), sizes.end()); std::reverse(starts.begin(), starts.end()); }
        std::string& bed = o.f[2 * side + 1];
        const size_t b0 = bed.size();
        if (side == 0) bed.append(qchrom, qchrom_n); else bed.append(schrom, schrom_n);
        bed += '\t'; put_int(bed, c0); bed += '\t'; put_int(bed, c1); bed += '\t'; bed.append(name, name_n); bed += '\t';
        if (is_float) put_double(bed, fsum); else put_int(bed, isum);
        bed += '\t'; bed += strand; bed += '\t'; put_int(bed, c0); bed += '\t'; put_int(bed, c1); bed += "\t0\t"; put_int(bed, n); bed += '\t';
        for (int64_t k = 0; k < n; ++k) { if (k) bed += ','; put_int(bed, sizes[k]); }
        bed += '\t';
        for (int64_t k = 0; k < n; ++k) { if (k) bed += ','; put_int(bed, starts[k]); }
        std::string& tsv = o.f[2 * side];
        tsv.append(bed, b0, std::string::npos);
        tsv += '\t';
        for (int64_t k = 0; k < n; ++k) { if (k) tsv += ','; put_double(tsv, in.pval[rev ? h1 - 1 - k : h0 + k]); }
        tsv += '\n';
        bed += '\n';
    }
}

}  // namespace

extern "C" int o2a_records_repr_double(double x, char* buf) { return repr_double(x, buf); }

extern "C" void o2a_records_free(char* p) { free(p); }

extern "C" int o2a_records_format(const o2a_records_input* in, int n_threads, char** out, int64_t* out_len)
{
    if (!in || !out || !out_len || in->n_chain < 0) return -1;
    const int64_t n = in->n_chain;
    if (n > 0 && (!in->hsp_off || !in->qstart || !in->qend || !in->sstart || !in->send || !in->score || !in->score_is_int ||
                  !in->pval || !in->qstrand_plus || !in->str_off || !in->str_blob)) return -1;
    for (int64_t c = 0; c < n; ++c) if (in->hsp_off[c + 1] <= in->hsp_off[c]) return -1;    // empty chains have no record
    int T = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (T < 1) T = 1;
    if ((int64_t)T > n / 256 + 1) T = (int)(n / 256 + 1);
    std::vector<Out> parts((size_t)T);
    auto work = [&](int t) {
        const int64_t c0 = n * t / T, c1 = n * (t + 1) / T;
        std::vector<long long> sizes, starts;
        for (int64_t c = c0; c < c1; ++c) format_chain(*in, c, parts[(size_t)t], sizes, starts);
    };
    if (T == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    for (int f = 0; f < 4; ++f) {
        size_t tot = 0;
        for (auto& p : parts) tot += p.f[f].size();
        char* buf = (char*)malloc(tot ? tot : 1);
        if (!buf) { for (int g = 0; g < f; ++g) { free(out[g]); out[g] = nullptr; } return -1; }
        size_t at = 0;
        for (auto& p : parts) { memcpy(buf + at, p.f[f].data(), p.f[f].size()); at += p.f[f].size(); }
        out[f] = buf; out_len[f] = (int64_t)tot;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// get_best_orthologs on files (pipeline.py:959-1024): the four significant.* files are read in lock-step, CONSECUTIVE
// records with the same query name (column 4 of the query BED12 line) form a group, and max() / min() over
// key(query BED12 line) keeps the FIRST extremal record of each group:
//   total_length = int(col3) - int(col2)   block_count = int(col10)   block_length = sum(int(i) for i in col11.split(','))
//   weight = float(col5)
// Anything Python would treat differently from this reader — '\r', a record with fewer columns, a number that is not a
// plain decimal literal, NaN keys, integers beyond int64 — returns 1: the caller then runs the text-level Python
// restatement, which raises or decides exactly like the reference.
// ---------------------------------------------------------------------------------------------------------------------
namespace {

bool slurp(const char* path, std::string& out)
{
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize((size_t)(n > 0 ? n : 0));
    size_t got = n > 0 ? fread(&out[0], 1, (size_t)n, f) : 0;
    fclose(f);
    return got == out.size();
}

struct LineIter {
    const char* p; const char* end;
    bool next(const char*& b, const char*& e) {              // [b, e) includes the '\n' when there is one
        if (p >= end) return false;
        b = p;
        const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
        e = nl ? nl + 1 : end;
        p = e;
        return true;
    }
};

inline bool is_py_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v'; }

// field `idx` (0-based) of line.strip().split('\t'); false if the record has fewer fields
bool field(const char* b, const char* e, int idx, const char*& fb, const char*& fe)
{
    while (b < e && is_py_space(*b)) ++b;
    while (e > b && is_py_space(e[-1])) --e;
    const char* s = b;
    for (int k = 0;; ++k) {
        const char* t = (const char*)memchr(s, '\t', (size_t)(e - s));
        const char* stop = t ? t : e;
        if (k == idx) { fb = s; fe = stop; return true; }
        if (!t) return false;
        s = t + 1;
    }
}

bool parse_i64(const char* b, const char* e, long long& v)   // int(): plain optional sign + digits only (else: fallback)
{
    if (b >= e) return false;
    const char* s = b;
    if (*s == '+' ) ++s;
    auto r = std::from_chars(s, e, v);
    return s < e && r.ec == std::errc() && r.ptr == e && (*s == '-' || (*s >= '0' && *s <= '9'));
}

}  // namespace

extern "C" int o2a_best_orthologs_files(const char* q_bed, const char* q_tsv, const char* s_bed, const char* s_tsv,
                                        int value, int function, const char* out_q_bed, const char* out_q_tsv,
                                        const char* out_s_bed, const char* out_s_tsv, int64_t* n_records, int64_t* n_groups)
{
    if (value < 0 || value > 3 || function < 0 || function > 1) return -1;
    std::string in[4];
    const char* paths[4] = {q_bed, q_tsv, s_bed, s_tsv};
    for (int k = 0; k < 4; ++k) {
        if (!paths[k] || !slurp(paths[k], in[k])) return -1;
        if (memchr(in[k].data(), '\r', in[k].size())) return 1;          // universal newlines would rewrite the records
    }
    LineIter it[4];
    for (int k = 0; k < 4; ++k) it[k] = {in[k].data(), in[k].data() + in[k].size()};
    std::string out[4];
    for (int k = 0; k < 4; ++k) out[k].reserve(in[k].size() / 2 + 64);
    const char *lb[4], *le[4];               // current record
    const char *bb[4] = {nullptr, nullptr, nullptr, nullptr}, *be[4] = {nullptr, nullptr, nullptr, nullptr};   // best of the group
    const char *gname_b = nullptr, *gname_e = nullptr;
    bool have = false;
    long long best_i = 0; double best_d = 0.0;
    int64_t nrec = 0, ngrp = 0;
    auto flush = [&]() { for (int k = 0; k < 4; ++k) out[k].append(bb[k], (size_t)(be[k] - bb[k])); ++ngrp; };
    for (;;) {
        bool all = true;
        for (int k = 0; k < 4; ++k) all = it[k].next(lb[k], le[k]) && all;   // zip(): stops at the shortest file
        if (!all) break;
        ++nrec;
        const char *nb, *ne;
        if (!field(lb[0], le[0], 3, nb, ne)) return 1;
        // key of this record
        long long ki = 0; double kd = 0.0;
        const char *fb, *fe;
        if (value == 1) {                    // total_length
            long long a, b2;
            if (!field(lb[0], le[0], 2, fb, fe) || !parse_i64(fb, fe, a)) return 1;
            if (!field(lb[0], le[0], 1, fb, fe) || !parse_i64(fb, fe, b2)) return 1;
            if (__builtin_sub_overflow(a, b2, &ki)) return 1;
        } else if (value == 2) {             // block_count
            if (!field(lb[0], le[0], 9, fb, fe) || !parse_i64(fb, fe, ki)) return 1;
        } else if (value == 0) {             // block_length
            if (!field(lb[0], le[0], 10, fb, fe)) return 1;
            const char* s = fb;
            for (;;) {
                const char* c = (const char*)memchr(s, ',', (size_t)(fe - s));
                long long v;
                if (!parse_i64(s, c ? c : fe, v) || __builtin_add_overflow(ki, v, &ki)) return 1;
                if (!c) break;
                s = c + 1;
            }
        } else {                             // weight: float(col5)
            if (!field(lb[0], le[0], 4, fb, fe) || fb >= fe) return 1;
            for (const char* c = fb; c < fe; ++c)
                if (!((*c >= '0' && *c <= '9') || *c == '.' || *c == 'e' || *c == 'E' || *c == '-' || *c == '+')) return 1;   // inf, nan, '_' ...
            const char* s = *fb == '+' ? fb + 1 : fb;
            auto r = std::from_chars(s, fe, kd);
            if (r.ec != std::errc() || r.ptr != fe || kd != kd) return 1;
        }
        const bool same = have && (size_t)(ne - nb) == (size_t)(gname_e - gname_b) && !memcmp(nb, gname_b, (size_t)(ne - nb));
        if (have && !same) { flush(); have = false; }
        bool take;
        if (!have) take = true;
        else if (value == 3) take = function == 0 ? kd > best_d : kd < best_d;       // strict: the first extremal record wins
        else take = function == 0 ? ki > best_i : ki < best_i;
        if (take) { for (int k = 0; k < 4; ++k) { bb[k] = lb[k]; be[k] = le[k]; } best_i = ki; best_d = kd; }
        // the reference compares the name with the LAST record of the group (purge[-1]): all of them carry the same name
        gname_b = nb; gname_e = ne;
        have = true;
    }
    if (have) flush();
    const char* outs[4] = {out_q_bed, out_q_tsv, out_s_bed, out_s_tsv};
    for (int k = 0; k < 4; ++k) {
        FILE* f = fopen(outs[k], "wb");
        if (!f) return -1;
        const size_t w = out[k].empty() ? 0 : fwrite(out[k].data(), 1, out[k].size(), f);
        fclose(f);
        if (w != out[k].size()) return -1;
    }
    if (n_records) *n_records = nrec;
    if (n_groups) *n_groups = ngrp;
    return 0;
}
