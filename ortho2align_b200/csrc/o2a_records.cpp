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
