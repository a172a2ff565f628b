// o2a_blast.cpp — native reader of `blastn -outfmt "7 std score"` tables into the columnar HSP arrays
// (include/o2a_blast.h). Follows Alignment.from_file_blast (alignment.py:515-586) decision by decision,
// including the ones that look accidental: the header search gives up after 100 lines and returns an empty
// alignment (:554-560); `lstrip("# Fields: ")` strips a SET of characters (:562); the hit count is the second
// space-separated token of the next line (:568); a Fields line followed by zero hits leaves `hsp` unbound
// (:583, UnboundLocalError); qlen / slen come from the LAST hit's extra columns (:583-584).
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <string_view>
#include <thread>
#include <vector>

#include "../../include/o2a_blast.h"

namespace {

struct Table {
    int status = 0;
    const char* error = "";
    std::vector<int32_t> qs, qe, ss, se;
    std::vector<double> score;
    std::vector<uint8_t> score_is_int;
    int64_t qlen = 1, slen = 1;
};

enum Kind { K_STR, K_INT, K_FLOAT };
struct Value { Kind kind = K_STR; double f = 0.0; };

bool py_space(char c) { return c == ' ' || (c >= '\t' && c <= '\r') || c == '\x1c' || c == '\x1d' || c == '\x1e' || c == '\x1f'; }

std::string_view strip(std::string_view s)
{
    while (!s.empty() && py_space(s.front())) s.remove_prefix(1);
    while (!s.empty() && py_space(s.back())) s.remove_suffix(1);
    return s;
}

// utils.numberize (utils.py:21-41): float(string) if it parses, int if that float is integral
Value numberize(std::string_view tok)
{
    Value v;
    std::string_view t = strip(tok);
    if (t.empty() || t.size() > 400) return v;
    // Python's float() takes decimal literals, inf/infinity/nan; strtod would also take hex floats
    for (size_t i = 0; i < t.size(); ++i) if (t[i] == 'x' || t[i] == 'X' || t[i] == '_') return v;
    char buf[404];
    memcpy(buf, t.data(), t.size());
    buf[t.size()] = 0;
    char* end = nullptr;
    double f = strtod(buf, &end);
    if (end != buf + t.size()) return v;
    v.f = f;
    v.kind = (std::isfinite(f) && f == std::floor(f)) ? K_INT : K_FLOAT;
    return v;
}

struct Lines {
    std::string_view text;
    size_t pos = 0;
    // file_object.readline(): up to and including '\n'; "" at EOF
    std::string_view readline()
    {
        if (pos >= text.size()) return std::string_view();
        size_t nl = text.find('\n', pos);
        size_t end = nl == std::string_view::npos ? text.size() : nl + 1;
        std::string_view ln = text.substr(pos, end - pos);
        pos = end;
        return ln;
    }
};

bool starts_with(std::string_view s, const char* p) { size_t n = strlen(p); return s.size() >= n && memcmp(s.data(), p, n) == 0; }

void fail(Table& T, int status, const char* what)
{
    T.status = status; T.error = what;
    T.qs.clear(); T.qe.clear(); T.ss.clear(); T.se.clear(); T.score.clear(); T.score_is_int.clear();
}

enum Col { C_QSTART, C_QEND, C_SSTART, C_SEND, C_SCORE, C_QLEN, C_SLEN, C_OTHER };

void parse_table(std::string_view text, int start_type, int end_inclusive, Table& T)
{
    Lines in{text};
    std::string_view line = in.readline();
    if (!starts_with(line, "#")) return fail(T, O2A_BLAST_ERROR, "ValueError");            // :550-552
    bool found = false;
    for (int k = 0; k < 100; ++k) {                                                          // :554-560
        line = in.readline();
        if (starts_with(line, "# Fields:")) { found = true; break; }
    }
    if (!found) return;                                  // cls([], **kwargs): empty alignment, qlen = slen = 1
    // :562 line.lstrip("# Fields: ").rstrip("\n").split(", ")
    {
        const char* set = "# Fields:";
        while (!line.empty() && strchr(set, line.front())) line.remove_prefix(1);
        while (!line.empty() && line.back() == '\n') line.remove_suffix(1);
    }
    std::vector<Col> cols;
    bool has_score = false;
    for (size_t p = 0;;) {
        size_t q = line.find(", ", p);
        std::string_view f = line.substr(p, q == std::string_view::npos ? std::string_view::npos : q - p);
        Col c = C_OTHER;
        if (f == "q. start" || f == "qstart") c = C_QSTART;                                  // replace_dict :440-445
        else if (f == "q. end" || f == "qend") c = C_QEND;
        else if (f == "s. start" || f == "sstart") c = C_SSTART;
        else if (f == "s. end" || f == "send") c = C_SEND;
        else if (f == "score") { c = C_SCORE; has_score = true; }
        else if (f == "query length" || f == "qlen") c = C_QLEN;
        else if (f == "subject length" || f == "slen") c = C_SLEN;
        cols.push_back(c);
        if (q == std::string_view::npos) break;
        p = q + 2;
    }
    if (!has_score) return fail(T, O2A_BLAST_ERROR, "ValueError");                           // :563-565
    // :568 int(file_object.readline().split(" ")[1])
    int64_t hit_number = 0;
    {
        line = in.readline();
        size_t a = line.find(' ');
        if (a == std::string_view::npos) return fail(T, O2A_BLAST_ERROR, "IndexError");
        size_t b = line.find(' ', a + 1);
        std::string_view tok = strip(line.substr(a + 1, b == std::string_view::npos ? std::string_view::npos : b - a - 1));
        bool neg = false;
        if (!tok.empty() && (tok.front() == '+' || tok.front() == '-')) { neg = tok.front() == '-'; tok.remove_prefix(1); }
        if (tok.empty() || tok.size() > 18) return fail(T, O2A_BLAST_ERROR, "ValueError");
        for (char c : tok) { if (c < '0' || c > '9') return fail(T, O2A_BLAST_ERROR, "ValueError"); hit_number = hit_number * 10 + (c - '0'); }
        if (neg) hit_number = -hit_number;
    }
    if (hit_number <= 0) return fail(T, O2A_BLAST_ERROR, "UnboundLocalError");               // :583 `hsp` never bound
    T.qs.reserve((size_t)hit_number); T.qe.reserve((size_t)hit_number); T.ss.reserve((size_t)hit_number);
    T.se.reserve((size_t)hit_number); T.score.reserve((size_t)hit_number); T.score_is_int.reserve((size_t)hit_number);
    Value last_qlen, last_slen;
    bool have_qlen = false, have_slen = false;
    for (int64_t h = 0; h < hit_number; ++h) {
        std::string_view row = strip(in.readline());                                         // :571
        Value v[5];
        bool have[5] = {false, false, false, false, false};
        have_qlen = have_slen = false;
        size_t p = 0;
        for (size_t c = 0; c < cols.size(); ++c) {               // dict(zip(fields, values)): a later duplicate wins
            size_t q = row.find('\t', p);
            std::string_view tok = row.substr(p, q == std::string_view::npos ? std::string_view::npos : q - p);
            Value val = numberize(tok);
            if (cols[c] <= C_SCORE) { v[cols[c]] = val; have[cols[c]] = true; }
            else if (cols[c] == C_QLEN) { last_qlen = val; have_qlen = true; }
            else if (cols[c] == C_SLEN) { last_slen = val; have_slen = true; }
            if (q == std::string_view::npos) break;
            p = q + 1;
        }
        for (int k = 0; k < 5; ++k) if (!have[k]) return fail(T, O2A_BLAST_ERROR, "TypeError");     // HSPVertex(**dict) :573
        double c4[4];
        for (int k = 0; k < 4; ++k) {
            if (v[k].kind == K_STR) return fail(T, O2A_BLAST_ERROR, "TypeError");            // `hsp.qstart -= 1` on a str :575
            if (v[k].kind == K_FLOAT) return fail(T, O2A_BLAST_UNSUPPORTED, "non-integral coordinate");
            c4[k] = v[k].f;
        }
        if (start_type == 1) for (int k = 0; k < 4; ++k) c4[k] -= 1;                          // :574-578
        if (end_inclusive) { c4[1] += 1; c4[3] += 1; }                                       // :579-581
        for (int k = 0; k < 4; ++k)
            if (!(c4[k] >= -2147483648.0 && c4[k] <= 2147483647.0)) return fail(T, O2A_BLAST_UNSUPPORTED, "coordinate outside int32");
        if (v[C_SCORE].kind == K_STR) return fail(T, O2A_BLAST_UNSUPPORTED, "non-numeric score");
        T.qs.push_back((int32_t)c4[0]); T.qe.push_back((int32_t)c4[1]); T.ss.push_back((int32_t)c4[2]); T.se.push_back((int32_t)c4[3]);
        T.score.push_back(v[C_SCORE].f);
        T.score_is_int.push_back(v[C_SCORE].kind == K_INT ? 1 : 0);
    }
    // Alignment.__init__ :477-487 with qlen / slen from the last hit's kwargs (:583-584), else max + 1
    int32_t mq = T.qe[0], ms = T.ss[0] > T.se[0] ? T.ss[0] : T.se[0];
    for (size_t i = 1; i < T.qe.size(); ++i) {
        if (T.qe[i] > mq) mq = T.qe[i];
        if (T.ss[i] > ms) ms = T.ss[i];
        if (T.se[i] > ms) ms = T.se[i];
    }
    T.qlen = (int64_t)mq + 1; T.slen = (int64_t)ms + 1;
    if (have_qlen) {
        if (last_qlen.kind != K_INT || std::fabs(last_qlen.f) > 9e15) return fail(T, O2A_BLAST_UNSUPPORTED, "non-integral query length");
        T.qlen = (int64_t)last_qlen.f;
    }
    if (have_slen) {
        if (last_slen.kind != K_INT || std::fabs(last_slen.f) > 9e15) return fail(T, O2A_BLAST_UNSUPPORTED, "non-integral subject length");
        T.slen = (int64_t)last_slen.f;
    }
}

}  // namespace

struct o2a_blast {
    int n_threads = 1;
    std::vector<Table> tables;
};

extern "C" o2a_blast* o2a_blast_create(int n_threads)
{
    o2a_blast* b = new o2a_blast();
    b->n_threads = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (b->n_threads < 1) b->n_threads = 1;
    return b;
}

extern "C" void o2a_blast_destroy(o2a_blast* b) { delete b; }

extern "C" int o2a_blast_parse(o2a_blast* b, const char* const* texts, const int64_t* lens, int64_t n_tables,
                               int start_type, int end_inclusive)
{
    if (!b || n_tables < 0 || (n_tables > 0 && (!texts || !lens))) return -1;
    b->tables.assign((size_t)n_tables, Table());
    std::atomic<int64_t> next(0);
    auto work = [&]() {
        for (;;) {
            int64_t i = next.fetch_add(1);
            if (i >= n_tables) break;
            parse_table(std::string_view(texts[i], (size_t)(lens[i] > 0 ? lens[i] : 0)), start_type, end_inclusive, b->tables[(size_t)i]);
        }
    };
    int nt = b->n_threads;
    if (nt > n_tables) nt = (int)(n_tables > 0 ? n_tables : 1);
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    return 0;
}

extern "C" int o2a_blast_sizes(const o2a_blast* b, int64_t* n_tables, int64_t* n_hsp)
{
    if (!b) return -1;
    int64_t h = 0;
    for (const Table& T : b->tables) h += (int64_t)T.score.size();
    if (n_tables) *n_tables = (int64_t)b->tables.size();
    if (n_hsp) *n_hsp = h;
    return 0;
}

extern "C" int o2a_blast_export(const o2a_blast* b, int32_t* status, int64_t* hsp_off,
                                int32_t* qstart, int32_t* qend, int32_t* sstart, int32_t* send,
                                double* score, uint8_t* score_is_int, int64_t* qlen, int64_t* slen)
{
    if (!b) return -1;
    int64_t o = 0;
    for (size_t i = 0; i < b->tables.size(); ++i) {
        const Table& T = b->tables[i];
        const size_t n = T.score.size();
        if (status) status[i] = T.status;
        if (hsp_off) hsp_off[i] = o;
        if (qlen) qlen[i] = T.qlen;
        if (slen) slen[i] = T.slen;
        if (n) {
            if (qstart) memcpy(qstart + o, T.qs.data(), n * 4);
            if (qend) memcpy(qend + o, T.qe.data(), n * 4);
            if (sstart) memcpy(sstart + o, T.ss.data(), n * 4);
            if (send) memcpy(send + o, T.se.data(), n * 4);
            if (score) memcpy(score + o, T.score.data(), n * 8);
            if (score_is_int) memcpy(score_is_int + o, T.score_is_int.data(), n);
        }
        o += (int64_t)n;
    }
    if (hsp_off) hsp_off[b->tables.size()] = o;
    return 0;
}

extern "C" const char* o2a_blast_error(const o2a_blast* b, int64_t i)
{
    if (!b || i < 0 || i >= (int64_t)b->tables.size()) return "";
    return b->tables[(size_t)i].error;
}


// --------------------------------------------------------------------------------------------------
// random.Random(seed).sample(population, k) of CPython (see include/o2a_blast.h for the sources restated)
// --------------------------------------------------------------------------------------------------
namespace {

struct MT19937 {
    static constexpr int N = 624, M = 397;
    uint32_t mt[N];
    int idx = N;

    void init_genrand(uint32_t s)
    {
        mt[0] = s;
        for (int i = 1; i < N; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        idx = N;
    }
    void init_by_array(const uint32_t* key, int len)
    {
        init_genrand(19650218u);
        int i = 1, j = 0;
        for (int k = N > len ? N : len; k; --k) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
            ++i; ++j;
            if (i >= N) { mt[0] = mt[N - 1]; i = 1; }
            if (j >= len) j = 0;
        }
        for (int k = N - 1; k; --k) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
            ++i;
            if (i >= N) { mt[0] = mt[N - 1]; i = 1; }
        }
        mt[0] = 0x80000000u;
    }
    uint32_t next()
    {
        if (idx >= N) {
            static const uint32_t mag01[2] = {0u, 0x9908b0dfu};
            int kk = 0;
            for (; kk < N - M; ++kk) {
                uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
                mt[kk] = mt[kk + M] ^ (y >> 1) ^ mag01[y & 1u];
            }
            for (; kk < N - 1; ++kk) {
                uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
                mt[kk] = mt[kk + (M - N)] ^ (y >> 1) ^ mag01[y & 1u];
            }
            uint32_t y = (mt[N - 1] & 0x80000000u) | (mt[0] & 0x7fffffffu);
            mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ mag01[y & 1u];
            idx = 0;
        }
        uint32_t y = mt[idx++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
    // Random._randbelow_with_getrandbits for 0 < n < 2^32
    uint32_t randbelow(uint32_t n)
    {
        int k = 32 - __builtin_clz(n);              // n.bit_length()
        uint32_t r = next() >> (32 - k);
        while (r >= n) r = next() >> (32 - k);
        return r;
    }
};

}  // namespace

extern "C" int64_t o2a_bg_sample(const int32_t* scores, int64_t n, int64_t observations, uint64_t seed, int32_t* out)
{
    if (n < 0 || observations < 0 || n >= (1ll << 32) || (n > 0 && (!scores || !out))) return -1;
    if (n <= observations) {                         // pipeline.py:227, 263: sampled only when there are more
        if (n) memcpy(out, scores, (size_t)n * sizeof(int32_t));
        return n;
    }
    const int64_t k = observations;
    MT19937 g;
    uint32_t key[2] = {(uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32)};
    g.init_by_array(key, key[1] ? 2 : 1);            // random_seed: 32-bit words of abs(seed), at least one
    // setsize = 21 (+ 4 ** ceil(log(3k, 4)) for k > 5): 3k is never a power of 4, so the ceiling is the smallest e with 4^e >= 3k
    int64_t setsize = 21;
    if (k > 5) { int64_t p = 1; while (p < 3 * k) p *= 4; setsize += p; }
    if (n <= setsize) {
        std::vector<int32_t> pool(scores, scores + n);
        for (int64_t i = 0; i < k; ++i) {
            uint32_t j = g.randbelow((uint32_t)(n - i));
            out[i] = pool[j];
            pool[j] = pool[(size_t)(n - i - 1)];     // move non-selected item into vacancy
        }
    } else {
        std::vector<uint8_t> selected((size_t)n, 0); // `j in selected` — membership only, so a bitmap does
        for (int64_t i = 0; i < k; ++i) {
            uint32_t j = g.randbelow((uint32_t)n);
            while (selected[j]) j = g.randbelow((uint32_t)n);
            selected[j] = 1;
            out[i] = scores[j];
        }
    }
    return k;
}
