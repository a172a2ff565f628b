// o2a_ingest.cpp — native ingest of the reference's per-lncRNA JSON files into the columnar batch
// of include/o2a.h (the reference spends 94 % of build_orthologs in
// json.load + HSP/Alignment object construction, pipeline.py:743-754, genomicranges.py:1084-1106,
// alignment.py:89-110, 233-249).
//
// align_files/<name>.json = JSON list of alignment dicts (to_dict: genomicranges.py:1077-1082,
// alignment.py:588-601); bg_files/<name>.json = JSON list of ints (pipeline.py:266-273).
// What is extracted per alignment: the HSPs of key "HSPs" ONLY (the reference rebuilds the working
// list from 'HSPs' and ignores 'filtered_HSPs', genomicranges.py:1094-1100), qlen/slen (null ->
// max+1 like alignment.py:477-487), the truthiness of 'relative' (to_record raises for relative
// alignments, genomicranges.py:1259), and the BYTE SPANS of the "qrange"/"srange" objects so the
// host decodes only those few hundred bytes when it formats a record. Scores keep their JSON type
// (int vs float) in `score_is_int` because the reference's record text depends on it.
// Background patches of pipeline.py:746-754 are applied here, in the reference's order.
//
// Two output forms (include/o2a_ingest.h): the plain columnar arrays (o2a_ingest_export), and the
// PCIe-optimal lossless layout o2a_build_batch takes from host memory (o2a_ingest_export_narrow:
// int8/int16/int32 scores + exception lists, narrow background; coordinates AoS) written straight
// into caller-provided — normally pinned — buffers by the same thread pool that parsed the files.
//
// Plain C ABI; files are parsed by a pool of std::threads.
#include <atomic>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../../include/o2a_ingest.h"
#include "o2a_pack.h"

namespace {

struct FileResult {
    int status = 0;                 // 0 ok, 1 parse error, 2 range error (coordinate / background not int32)
    std::string error;
    std::vector<int64_t> hsp_count, qlen, slen;
    std::vector<int64_t> qr_begin, qr_end, sr_begin, sr_end;   // byte spans inside `meta`
    std::string meta;               // the qrange / srange JSON texts of every alignment, back to back
    std::vector<int32_t> coords;    // AoS qstart,qend,sstart,send
    std::vector<double> score;
    std::vector<uint8_t> score_is_int;
    std::vector<int32_t> bg;
    std::vector<uint8_t> relative;  // per alignment: the 'relative' flag is truthy (to_record raises, genomicranges.py:1259)
    // per alignment, for the record formatter: str(qrange.chrom), str(qrange.name), str(srange.chrom) back to back in
    // `names` (offsets 3 per alignment + 1), qrange.strand in ('+', '.'), and whether these were plain JSON strings /
    // nulls / ints (else the host decodes the range object itself)
    std::string names;
    std::vector<int64_t> name_off;
    std::vector<uint8_t> q_plus, simple;
    // HSPs of the file whose score does not fit the lossless int8 / int16 / int32 encodings of include/o2a.h
    int64_t misfit[3] = {0, 0, 0};
    int32_t bg_min = 0, bg_max = 0;
    std::string text;               // file contents when they were read (not mapped); released after parsing
    const char* view = nullptr;     // file contents: mapped (MappedFile) or `text`
    size_t view_len = 0;
};

// read-only mapping of a whole file; falls back to read() into `fallback` when mmap is not possible
struct MappedFile {
    void* map = MAP_FAILED; size_t len = 0;
    bool open_file(const char* path, std::string& fallback, const char*& data, size_t& n, bool& missing);
    ~MappedFile() { if (map != MAP_FAILED) munmap(map, len); }
};

struct Parser {
    const char* base; const char* p; const char* end; std::string err;
    bool fail(const char* m) { if (err.empty()) { err = m; err += " at byte "; err += std::to_string(p - base); } return false; }
    void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p; }
    bool lit(char c) { ws(); if (p < end && *p == c) { ++p; return true; } return false; }
    bool expect(char c) { if (lit(c)) return true; return fail("unexpected character"); }
    // string: returns [b, e) of the raw contents (escapes left as they are)
    bool str(const char*& b, const char*& e) {
        ws();
        if (p >= end || *p != '"') return fail("expected string");
        b = ++p;
        while (p < end && *p != '"') { if (*p == '\\') ++p; ++p; }
        if (p >= end) return fail("unterminated string");
        e = p++;
        return true;
    }
    bool skip() {                   // any JSON value
        ws();
        if (p >= end) return fail("unexpected end");
        char c = *p;
        if (c == '"') { const char *b, *e; return str(b, e); }
        if (c == '{' || c == '[') {
            // bracket matching that respects strings: no recursion, any depth
            int depth = 0;
            while (p < end) {
                char d = *p;
                if (d == '"') { const char *b, *e; if (!str(b, e)) return false; continue; }
                if (d == '{' || d == '[') ++depth;
                else if (d == '}' || d == ']') { --depth; if (depth == 0) { ++p; return true; } }
                ++p;
            }
            return fail("unterminated container");
        }
        while (p < end && *p != ',' && *p != '}' && *p != ']' && *p != ' ' && *p != '\n' && *p != '\t' && *p != '\r') ++p;
        return true;
    }
    // number or null. kind: 0 null, 1 int, 2 float
    bool num(int& kind, int64_t& iv, double& dv) {
        ws();
        if (p + 4 <= end && !memcmp(p, "null", 4)) { p += 4; kind = 0; return true; }
        const char* s = p;
        bool is_float = false;
        if (p < end && (*p == '-' || *p == '+')) ++p;
        while (p < end && ((*p >= '0' && *p <= '9') || *p == '.' || *p == 'e' || *p == 'E' || *p == '-' || *p == '+')) {
            if (*p == '.' || *p == 'e' || *p == 'E') is_float = true;
            ++p;
        }
        if (p == s) return fail("expected number");
        if (!is_float) {
            auto r = std::from_chars(s, p, iv);
            if (r.ec == std::errc() && r.ptr == p) { kind = 1; dv = (double)iv; return true; }
            is_float = true;            // out of int64: keep the value as a double
        }
        auto r = std::from_chars(s, p, dv);
        if (r.ec != std::errc() || r.ptr != p) return fail("bad number");
        kind = 2; iv = 0;
        return true;
    }
    // Python truthiness of any JSON value (for 'relative'): false / null / 0 / "" / [] / {} are falsy
    bool truthy(bool& out) {
        ws();
        if (p >= end) return fail("unexpected end");
        const char* s = p;
        if (!skip()) return false;
        const size_t n = (size_t)(p - s);
        if ((n == 5 && !memcmp(s, "false", 5)) || (n == 4 && !memcmp(s, "null", 4))) { out = false; return true; }
        if (n == 4 && !memcmp(s, "true", 4)) { out = true; return true; }
        if (*s == '"') { out = n > 2; return true; }
        if (*s == '[' || *s == '{') {
            const char* q = s + 1;
            while (q < p - 1 && (*q == ' ' || *q == '\n' || *q == '\t' || *q == '\r')) ++q;
            out = q < p - 1;
            return true;
        }
        double dv = 0;
        auto r = std::from_chars(s, p, dv);
        out = !(r.ec == std::errc() && dv == 0.0);
        return true;
    }
};

inline bool key_is(const char* b, const char* e, const char* name) {
    size_t n = strlen(name);
    return (size_t)(e - b) == n && !memcmp(b, name, n);
}

bool read_file(const char* path, std::string& out, bool& missing) {
    missing = false;
    int fd = open(path, O_RDONLY);
    if (fd < 0) { missing = true; return false; }
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); return false; }
    const size_t n = (size_t)(st.st_size > 0 ? st.st_size : 0);
    out.resize(n);
    size_t got = 0;
    while (got < n) {
        ssize_t r = read(fd, &out[got], n - got);
        if (r <= 0) break;
        got += (size_t)r;
    }
    close(fd);
    return got == n;
}

bool MappedFile::open_file(const char* path, std::string& fallback, const char*& data, size_t& n, bool& missing)
{
    missing = false;
    int fd = open(path, O_RDONLY);
    if (fd < 0) { missing = true; return false; }
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); return false; }
    n = (size_t)(st.st_size > 0 ? st.st_size : 0);
    if (n >= (1u << 22)) {                       // below 4 MB read() is cheaper than a mapping: 16 threads mapping and unmapping 10^5 small files serialise on the address-space lock
        map = mmap(nullptr, n, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
        if (map != MAP_FAILED) { len = n; close(fd); data = (const char*)map; return true; }
    }
    fallback.resize(n);
    size_t got = 0;
    while (got < n) {
        ssize_t r = read(fd, &fallback[got], n - got);
        if (r <= 0) break;
        got += (size_t)r;
    }
    close(fd);
    data = fallback.data();
    return got == n;
}

// which narrow encodings cannot hold score `sc` (bit 0: int8, bit 1: int16, bit 2: int32): o2a_pack.h
using o2a_pack::narrow_misfit;


// JSON string contents [b, e) -> UTF-8 (json.dump writes non-ASCII as \uXXXX). Returns false on a malformed escape.
bool json_unescape(const char* b, const char* e, std::string& out)
{
    auto hex4 = [](const char* p, unsigned& v) {
        v = 0;
        for (int k = 0; k < 4; ++k) {
            const char c = p[k]; unsigned d;
            if (c >= '0' && c <= '9') d = (unsigned)(c - '0'); else if (c >= 'a' && c <= 'f') d = (unsigned)(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F') d = (unsigned)(c - 'A' + 10); else return false;
            v = v * 16 + d;
        }
        return true;
    };
    auto utf8 = [&](unsigned cp) {
        if (cp < 0x80) out.push_back((char)cp);
        else if (cp < 0x800) { out.push_back((char)(0xC0 | (cp >> 6))); out.push_back((char)(0x80 | (cp & 0x3F))); }
        else if (cp < 0x10000) { out.push_back((char)(0xE0 | (cp >> 12))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F))); }
        else { out.push_back((char)(0xF0 | (cp >> 18))); out.push_back((char)(0x80 | ((cp >> 12) & 0x3F))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F))); }
    };
    for (const char* p = b; p < e; ++p) {
        if (*p != '\\') { out.push_back(*p); continue; }
        if (++p >= e) return false;
        switch (*p) {
            case '"': out.push_back('"'); break; case '\\': out.push_back('\\'); break; case '/': out.push_back('/'); break;
            case 'b': out.push_back('\b'); break; case 'f': out.push_back('\f'); break; case 'n': out.push_back('\n'); break;
            case 'r': out.push_back('\r'); break; case 't': out.push_back('\t'); break;
            case 'u': {
                unsigned cp;
                if (e - p < 5 || !hex4(p + 1, cp)) return false;
                p += 4;
                if (cp >= 0xD800 && cp < 0xDC00) {                       // surrogate pair
                    unsigned lo;
                    if (e - p < 7 || p[1] != '\\' || p[2] != 'u' || !hex4(p + 3, lo) || lo < 0xDC00 || lo > 0xDFFF) return false;
                    cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                    p += 6;
                } else if (cp >= 0xDC00 && cp <= 0xDFFF) return false;   // lone surrogate: Python keeps it, UTF-8 cannot
                utf8(cp);
                break;
            }
            default: return false;
        }
    }
    return true;
}

// What the record formatter needs from a qrange / srange object (GenomicRange.from_dict, genomicranges.py:816-855):
// str(chrom), the strand, and str(name) with the reference's fall-back to "chrom:start-endstrand" for a null name
// (genomicranges.py:427-433). ok = false: a value has a JSON type this reader does not render (numbers as chrom / name,
// nested values, lone surrogates): the host then decodes the object with Python's json.
struct RangeMini { std::string chrom, name, strand; bool ok = true; };

void extract_range(const char* b, const char* e, RangeMini& out)
{
    Parser P{b, b, e, {}};
    bool have_chrom = false, have_name = false, name_null = false, have_strand = false, have_s = false, have_e = false;
    int64_t start = 0, end = 0;
    auto text_value = [&](std::string& dst, bool* is_null) -> bool {       // string or null
        P.ws();
        if (P.p < P.end && *P.p == '"') { const char *sb, *se; return P.str(sb, se) && json_unescape(sb, se, dst); }
        if (P.p + 4 <= P.end && !memcmp(P.p, "null", 4)) { P.p += 4; if (is_null) *is_null = true; else dst = "None"; return true; }
        return false;
    };
    if (!P.expect('{')) { out.ok = false; return; }
    if (!P.lit('}')) {
        do {
            const char *kb, *ke;
            if (!P.str(kb, ke) || !P.expect(':')) { out.ok = false; return; }
            if (key_is(kb, ke, "chrom")) { if (!text_value(out.chrom, nullptr)) { out.ok = false; return; } have_chrom = true; }
            else if (key_is(kb, ke, "name")) { if (!text_value(out.name, &name_null)) { out.ok = false; return; } have_name = true; }
            else if (key_is(kb, ke, "strand")) { if (!text_value(out.strand, nullptr)) { out.ok = false; return; } have_strand = true; }
            else if (key_is(kb, ke, "start") || key_is(kb, ke, "end")) {
                int kind; int64_t iv; double dv;
                if (!P.num(kind, iv, dv) || kind != 1) { out.ok = false; return; }
                if (kb[0] == 's') { start = iv; have_s = true; } else { end = iv; have_e = true; }
            } else if (!P.skip()) { out.ok = false; return; }
        } while (P.lit(','));
    }
    if (!(have_chrom && have_name && have_strand && have_s && have_e)) { out.ok = false; return; }
    if (name_null) out.name = out.chrom + ":" + std::to_string(start) + "-" + std::to_string(end) + out.strand;
}

// Fast path for the text json.dump writes for one HSP (alignment.py:233-249 -> to_dict):
//   {"qstart": I, "qend": I, "sstart": I, "send": I, "score": N, "pval": null, "kwargs": {}}
// Matches the five leading keys with their default separators and plain integer literals; anything
// else returns false with `p` untouched and the general key loop takes the object.
inline bool fast_int(const char*& q, const char* end, int64_t& v)
{
    const char* s = q;
    bool neg = false;
    if (s < end && *s == '-') { neg = true; ++s; }
    const char* d0 = s;
    uint64_t acc = 0;
    while (s < end && *s >= '0' && *s <= '9') { acc = acc * 10 + (uint64_t)(*s - '0'); ++s; }
    const int nd = (int)(s - d0);
    if (nd == 0 || nd > 18) return false;
    if (s < end && (*s == '.' || *s == 'e' || *s == 'E')) return false;
    v = neg ? -(int64_t)acc : (int64_t)acc;
    q = s;
    return true;
}

inline bool fast_hsp(Parser& P, int64_t c[4], double& sc, uint8_t& isint, bool& closed)
{
    static const char K0[] = "\"qstart\": ", K1[] = ", \"qend\": ", K2[] = ", \"sstart\": ", K3[] = ", \"send\": ",
                      K4[] = ", \"score\": ", TAIL[] = ", \"pval\": null, \"kwargs\": {}}";
    const char* q = P.p;
    const char* end = P.end;
#define O2A_KEY(K) if ((size_t)(end - q) < sizeof(K) || memcmp(q, K, sizeof(K) - 1)) return false; q += sizeof(K) - 1
    O2A_KEY(K0); if (!fast_int(q, end, c[0])) return false;
    O2A_KEY(K1); if (!fast_int(q, end, c[1])) return false;
    O2A_KEY(K2); if (!fast_int(q, end, c[2])) return false;
    O2A_KEY(K3); if (!fast_int(q, end, c[3])) return false;
    O2A_KEY(K4);
#undef O2A_KEY
    int64_t iv;
    const char* s = q;
    if (fast_int(q, end, iv)) { sc = (double)iv; isint = 1; }
    else {
        while (q < end && ((*q >= '0' && *q <= '9') || *q == '.' || *q == 'e' || *q == 'E' || *q == '-' || *q == '+')) ++q;
        if (q == s) return false;
        bool is_float = false;
        for (const char* t = s; t < q; ++t) if (*t == '.' || *t == 'e' || *t == 'E') is_float = true;
        if (!is_float) return false;                                    // > 18 digits: the general path decides
        auto r = std::from_chars(s, q, sc);
        if (r.ec != std::errc() || r.ptr != q) return false;
        isint = 0;
    }
    closed = false;
    if ((size_t)(end - q) >= sizeof(TAIL) - 1 && !memcmp(q, TAIL, sizeof(TAIL) - 1)) { q += sizeof(TAIL) - 1; closed = true; }
    P.p = q;
    return true;
}

bool parse_hsps(Parser& P, FileResult& R, int64_t& count, int64_t& max_qend, int64_t& max_s, bool& range_err)
{
    count = 0; max_qend = 0; max_s = 0;
    if (!P.expect('[')) return false;
    if (P.lit(']')) return true;
    do {
        if (!P.expect('{')) return false;
        int64_t c[4] = {0, 0, 0, 0}; bool have[5] = {false, false, false, false, false};
        double sc = 0; uint8_t isint = 1;
        bool closed = false;
        bool more = false;              // the general key loop continues after the leading keys
        if (fast_hsp(P, c, sc, isint, closed)) {
            have[0] = have[1] = have[2] = have[3] = have[4] = true;
            if (!closed) { if (P.lit(',')) more = true; else if (!P.expect('}')) return false; }
        } else if (!P.lit('}')) more = true;
        if (more) {
            do {
                const char *kb, *ke;
                if (!P.str(kb, ke) || !P.expect(':')) return false;
                int which = key_is(kb, ke, "qstart") ? 0 : key_is(kb, ke, "qend") ? 1 : key_is(kb, ke, "sstart") ? 2
                          : key_is(kb, ke, "send") ? 3 : key_is(kb, ke, "score") ? 4 : -1;
                if (which < 0) { if (!P.skip()) return false; continue; }
                int kind; int64_t iv; double dv;
                if (!P.num(kind, iv, dv)) return false;
                if (kind == 0) return P.fail("null HSP field");
                if (which < 4) { if (kind != 1) return P.fail("non-integer coordinate"); c[which] = iv; }
                else { sc = dv; isint = kind == 1; }
                have[which] = true;
            } while (P.lit(','));
            if (!P.expect('}')) return false;
        }
        if (!(have[0] && have[1] && have[2] && have[3] && have[4])) return P.fail("HSP without qstart/qend/sstart/send/score");
        for (int k = 0; k < 4; ++k) {
            if (c[k] > INT32_MAX || c[k] < -(int64_t)INT32_MAX) range_err = true;
            R.coords.push_back((int32_t)c[k]);
        }
        R.score.push_back(sc); R.score_is_int.push_back(isint);
        const unsigned mf = narrow_misfit(sc);
        R.misfit[0] += mf & 1u; R.misfit[1] += (mf >> 1) & 1u; R.misfit[2] += (mf >> 2) & 1u;
        if (count == 0 || c[1] > max_qend) max_qend = c[1];
        int64_t ms = c[2] > c[3] ? c[2] : c[3];
        if (count == 0 || ms > max_s) max_s = ms;
        ++count;
    } while (P.lit(','));
    return P.expect(']');
}

void parse_alignment_file(FileResult& R) {
    Parser P{R.view, R.view, R.view + R.view_len, {}};
    bool range_err = false;
    {   // ~232 bytes of JSON per HSP (each HSP is written twice, alignment.py:595-601): one allocation per array
        const size_t guess = R.view_len / 200 + 16;
        R.coords.reserve(4 * guess); R.score.reserve(guess); R.score_is_int.reserve(guess);
    }
    if (!P.expect('[')) { R.status = 1; R.error = P.err; return; }
    if (!P.lit(']')) {
        do {
            if (!P.expect('{')) break;
            int64_t cnt = 0, mq = 0, msub = 0, ql = 0, sl = 0; bool hq = false, hs = false, hh = false;
            int64_t qb = -1, qe = -1, sb = -1, se = -1;
            bool rel = false;
            const char *hsps_b = nullptr, *hsps_e = nullptr;        // text of the "HSPs" value
            if (!P.lit('}')) {
                do {
                    const char *kb, *ke;
                    if (!P.str(kb, ke) || !P.expect(':')) break;
                    if (key_is(kb, ke, "HSPs")) {
                        P.ws();
                        hsps_b = P.p;
                        if (!parse_hsps(P, R, cnt, mq, msub, range_err)) break;
                        hsps_e = P.p;
                        hh = true;
                    }
                    else if (key_is(kb, ke, "qlen") || key_is(kb, ke, "slen")) {
                        int kind; int64_t iv; double dv;
                        if (!P.num(kind, iv, dv)) break;
                        if (kind == 2) { P.fail("non-integer qlen/slen"); break; }
                        if (kb[0] == 'q') { hq = kind == 1; ql = iv; } else { hs = kind == 1; sl = iv; }
                    } else if (key_is(kb, ke, "qrange") || key_is(kb, ke, "srange")) {
                        P.ws();
                        const char* vb = P.p;
                        if (!P.skip()) break;
                        int64_t b = (int64_t)R.meta.size();
                        R.meta.append(vb, (size_t)(P.p - vb));
                        int64_t e = (int64_t)R.meta.size();
                        if (kb[0] == 'q') { qb = b; qe = e; } else { sb = b; se = e; }
                    } else if (key_is(kb, ke, "relative")) {
                        if (!P.truthy(rel)) break;
                    } else if (key_is(kb, ke, "filtered_HSPs") && hsps_b) {
                        // to_dict writes the working list next to the full one; before any filtering the two texts are
                        // byte-identical, and the reference ignores this key's content anyway (genomicranges.py:1099-1100)
                        P.ws();
                        const size_t n = (size_t)(hsps_e - hsps_b);
                        if ((size_t)(P.end - P.p) > n && !memcmp(P.p, hsps_b, n)) P.p += n;
                        else if (!P.skip()) break;
                    } else if (!P.skip()) break;
                } while (P.lit(','));
                if (!P.err.empty() || !P.expect('}')) break;
            }
            if (!hh) { P.fail("alignment without HSPs"); break; }        // i.get('HSPs') is None -> TypeError
            if (qb < 0 || sb < 0) { P.fail("alignment without qrange/srange"); break; }
            R.hsp_count.push_back(cnt);
            R.qlen.push_back(hq ? ql : (cnt ? mq : 0) + 1);              // alignment.py:477-480
            R.slen.push_back(hs ? sl : (cnt ? msub : 0) + 1);            // alignment.py:481-487
            R.qr_begin.push_back(qb); R.qr_end.push_back(qe); R.sr_begin.push_back(sb); R.sr_end.push_back(se);
            R.relative.push_back(rel ? 1 : 0);
            {
                RangeMini q, sr;
                // consecutive alignments of a lncRNA repeat the same qrange text: decode it once
                const size_t na = R.qr_begin.size();
                const bool same_q = na >= 2 && (qe - qb) == (R.qr_end[na - 2] - R.qr_begin[na - 2]) &&
                                    !memcmp(R.meta.data() + qb, R.meta.data() + R.qr_begin[na - 2], (size_t)(qe - qb));
                if (R.name_off.empty()) R.name_off.push_back(0);
                if (same_q) {
                    const size_t o = R.name_off.size() - 4;       // previous alignment's three strings
                    const std::string qc = R.names.substr((size_t)R.name_off[o], (size_t)(R.name_off[o + 1] - R.name_off[o]));
                    const std::string qn = R.names.substr((size_t)R.name_off[o + 1], (size_t)(R.name_off[o + 2] - R.name_off[o + 1]));
                    q.chrom = qc; q.name = qn; q.ok = R.simple.back() & 1;
                    R.q_plus.push_back(R.q_plus.back());
                } else {
                    extract_range(R.meta.data() + qb, R.meta.data() + qe, q);
                    R.q_plus.push_back((q.strand == "+" || q.strand == ".") ? 1 : 0);
                }
                extract_range(R.meta.data() + sb, R.meta.data() + se, sr);
                R.names += q.chrom; R.name_off.push_back((int64_t)R.names.size());
                R.names += q.name; R.name_off.push_back((int64_t)R.names.size());
                R.names += sr.chrom; R.name_off.push_back((int64_t)R.names.size());
                R.simple.push_back((q.ok ? 1 : 0) | (sr.ok ? 2 : 0));
            }
        } while (P.lit(','));
        if (P.err.empty()) P.expect(']');
    }
    if (!P.err.empty()) { R.status = 1; R.error = P.err; return; }
    if (range_err) { R.status = 2; R.error = "HSP coordinates exceed int32"; }
}

void parse_background(const std::string& text, bool missing, FileResult& R) {
    std::vector<int64_t> v;
    if (missing) { v = {0, 1}; }                                         // pipeline.py:749-750
    else {
        Parser P{text.data(), text.data(), text.data() + text.size(), {}};
        v.reserve(text.size() / 3 + 4);
        if (!P.expect('[')) { R.status = 1; R.error = "background: " + P.err; return; }
        if (!P.lit(']')) {
            do {
                P.ws();
                int64_t iv;
                const char* q = P.p;
                if (fast_int(q, P.end, iv)) { P.p = q; v.push_back(iv); continue; }
                int kind; double dv;
                if (!P.num(kind, iv, dv)) break;
                if (kind != 1) { R.status = 2; R.error = "background scores must be integers within int32"; return; }
                v.push_back(iv);
            } while (P.lit(','));
            if (P.err.empty()) P.expect(']');
        }
        if (!P.err.empty()) { R.status = 1; R.error = "background: " + P.err; return; }
    }
    bool distinct2 = false;                                              // len(set(bg)) <= 1 -> += [1]
    for (size_t i = 1; i < v.size(); ++i) if (v[i] != v[0]) { distinct2 = true; break; }
    if (!distinct2) v.push_back(1);
    if (v[0] == 1) v.push_back(0);                                       // bg[0] == 1 -> append(0)
    R.bg.reserve(v.size());
    int64_t mn = v[0], mx = v[0];
    for (int64_t x : v) {
        if (x > INT32_MAX || x < -(int64_t)INT32_MAX) { R.status = 2; R.error = "background scores must be integers within int32"; return; }
        R.bg.push_back((int32_t)x);
        if (x < mn) mn = x;
        if (x > mx) mx = x;
    }
    R.bg_min = (int32_t)mn; R.bg_max = (int32_t)mx;
}

}  // namespace

struct o2a_ingest {
    int n_threads = 1;
    std::vector<FileResult> files;
    std::string err;
};

extern "C" o2a_ingest* o2a_ingest_create(int n_threads)
{
    o2a_ingest* g = new o2a_ingest();
    g->n_threads = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (g->n_threads < 1) g->n_threads = 1;
    return g;
}

extern "C" void o2a_ingest_destroy(o2a_ingest* g)
{
    if (!g) return;
    // 10^5 files hold ~10^6 small vectors and strings: released on the pool, not one by one on the caller's thread
    // (0.35 s of a genome-scale run went into this destructor)
    if (g->files.size() >= 1024 && g->n_threads > 1)
        o2a_pack::parallel_for(g->n_threads, (int64_t)g->files.size(), [&](int64_t i) { FileResult empty; std::swap(empty, g->files[(size_t)i]); });
    delete g;
}

extern "C" const char* o2a_ingest_file_error(const o2a_ingest* g, int64_t i)
{
    if (!g || i < 0 || i >= (int64_t)g->files.size()) return "";
    return g->files[(size_t)i].error.c_str();
}

extern "C" int o2a_ingest_load(o2a_ingest* g, const char* const* align_paths, const char* const* bg_paths, int64_t n_files)
{
    if (!g || n_files < 0) return -1;
    g->files.assign((size_t)n_files, FileResult());
    o2a_pack::parallel_for(g->n_threads, n_files, [&](int64_t i) {
        FileResult& R = g->files[(size_t)i];
        bool missing = false;
        {
            MappedFile mf;
            if (!mf.open_file(align_paths[i], R.text, R.view, R.view_len, missing)) { R.status = 1; R.error = "cannot read alignment file"; return; }
            parse_alignment_file(R);
            R.view = nullptr; R.view_len = 0;
            std::string().swap(R.text);
        }
        if (R.status) return;
        std::string bgtext;
        bool bg_missing = false;
        bool ok = read_file(bg_paths[i], bgtext, bg_missing);
        if (!ok && !bg_missing) { R.status = 1; R.error = "cannot read background file"; return; }
        parse_background(bgtext, bg_missing, R);
    });
    return 0;
}

// size and mtime (ns) of n files on a pool of threads; size -1 = the file cannot be stat'ed (sidecar manifest check:
// 2 x 10^5 os.stat calls from Python cost more than memory-mapping the sidecar they guard)
extern "C" int o2a_stat_files(const char* const* paths, int64_t n, int n_threads, int64_t* size, int64_t* mtime_ns)
{
    if (!paths || !size || !mtime_ns || n < 0) return 1;
    const int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    o2a_pack::parallel_for(nt, n, [&](int64_t i) {
        struct stat st;
        if (paths[i] && stat(paths[i], &st) == 0) {
            size[i] = (int64_t)st.st_size;
            mtime_ns[i] = (int64_t)st.st_mtim.tv_sec * 1000000000ll + (int64_t)st.st_mtim.tv_nsec;
        } else { size[i] = -1; mtime_ns[i] = 0; }
    });
    return 0;
}

extern "C" int o2a_ingest_sizes(const o2a_ingest* g, int64_t* n_lnc, int64_t* n_aln, int64_t* n_hsp, int64_t* n_bg)
{
    if (!g) return -1;
    int64_t l = 0, a = 0, h = 0, b = 0;
    for (const FileResult& R : g->files) {
        if (R.status) continue;
        ++l; a += (int64_t)R.hsp_count.size(); h += (int64_t)R.score.size(); b += (int64_t)R.bg.size();
    }
    if (n_lnc) *n_lnc = l; if (n_aln) *n_aln = a; if (n_hsp) *n_hsp = h; if (n_bg) *n_bg = b;
    return 0;
}

namespace {
// first lncRNA / alignment / HSP / background index of every file (files with status != 0 own nothing)
struct Bases { std::vector<int64_t> l, a, h, b; };
Bases file_bases(const o2a_ingest* g)
{
    Bases B;
    const size_t n = g->files.size();
    B.l.resize(n + 1); B.a.resize(n + 1); B.h.resize(n + 1); B.b.resize(n + 1);
    int64_t l = 0, a = 0, h = 0, b = 0;
    for (size_t i = 0; i < n; ++i) {
        B.l[i] = l; B.a[i] = a; B.h[i] = h; B.b[i] = b;
        const FileResult& R = g->files[i];
        if (R.status) continue;
        ++l; a += (int64_t)R.hsp_count.size(); h += (int64_t)R.score.size(); b += (int64_t)R.bg.size();
    }
    B.l[n] = l; B.a[n] = a; B.h[n] = h; B.b[n] = b;
    return B;
}
}  // namespace

extern "C" int o2a_ingest_export(const o2a_ingest* g, int32_t* file_status, int64_t* lnc_file,
                                 int64_t* bg_off, int32_t* bg, int64_t* aln_off, int64_t* hsp_off,
                                 int64_t* qlen, int64_t* slen, int32_t* coords,
                                 int32_t* qstart, int32_t* qend, int32_t* sstart, int32_t* send,
                                 double* score, uint8_t* score_is_int, int64_t* spans, uint8_t* relative)
{
    if (!g) return -1;
    const Bases B = file_bases(g);
    if (bg_off) bg_off[0] = 0; if (aln_off) aln_off[0] = 0; if (hsp_off) hsp_off[0] = 0;
    o2a_pack::parallel_for(g->n_threads, (int64_t)g->files.size(), [&](int64_t fi) {
        const size_t i = (size_t)fi;
        const FileResult& R = g->files[i];
        if (file_status) file_status[i] = R.status;
        if (R.status) return;
        const int64_t l = B.l[i], a = B.a[i], h = B.h[i], b = B.b[i];
        if (lnc_file) lnc_file[l] = (int64_t)i;
        const int64_t na = (int64_t)R.hsp_count.size(), nh = (int64_t)R.score.size(), nb = (int64_t)R.bg.size();
        if (bg) memcpy(bg + b, R.bg.data(), (size_t)nb * 4);
        int64_t hcur = h;
        for (int64_t k = 0; k < na; ++k) {
            hcur += R.hsp_count[(size_t)k];
            if (hsp_off) hsp_off[a + k + 1] = hcur;
            if (qlen) qlen[a + k] = R.qlen[(size_t)k];
            if (slen) slen[a + k] = R.slen[(size_t)k];
            if (relative) relative[a + k] = R.relative[(size_t)k];
            if (spans) {
                int64_t* s = spans + 4 * (a + k);
                s[0] = R.qr_begin[(size_t)k]; s[1] = R.qr_end[(size_t)k]; s[2] = R.sr_begin[(size_t)k]; s[3] = R.sr_end[(size_t)k];
            }
        }
        if (coords) memcpy(coords + 4 * h, R.coords.data(), (size_t)nh * 16);
        if (qstart && qend && sstart && send) {
            for (int64_t k = 0; k < nh; ++k) {
                qstart[h + k] = R.coords[(size_t)(4 * k)]; qend[h + k] = R.coords[(size_t)(4 * k + 1)];
                sstart[h + k] = R.coords[(size_t)(4 * k + 2)]; send[h + k] = R.coords[(size_t)(4 * k + 3)];
            }
        }
        if (score) memcpy(score + h, R.score.data(), (size_t)nh * 8);
        if (score_is_int) memcpy(score_is_int + h, R.score_is_int.data(), (size_t)nh);
        if (bg_off) bg_off[l + 1] = b + nb;
        if (aln_off) aln_off[l + 1] = a + na;
    });
    return 0;
}

extern "C" int o2a_ingest_narrow_plan(const o2a_ingest* g, int32_t* score_bits, int32_t* bg_bits, int64_t* n_exc)
{
    if (!g) return -1;
    int64_t mis[3] = {0, 0, 0}, n_hsp = 0;
    int32_t mn = 0, mx = 0;
    for (const FileResult& R : g->files) {
        if (R.status) continue;
        for (int k = 0; k < 3; ++k) mis[k] += R.misfit[k];
        n_hsp += (int64_t)R.score.size();
        if (R.bg_min < mn) mn = R.bg_min;
        if (R.bg_max > mx) mx = R.bg_max;
    }
    const int w = o2a_pack::choose_score_bits(mis, n_hsp);
    if (score_bits) *score_bits = w;
    if (bg_bits) *bg_bits = o2a_pack::choose_bg_bits(mn, mx);
    if (n_exc) *n_exc = mis[w == 8 ? 0 : w == 16 ? 1 : 2];
    return 0;
}

extern "C" int o2a_ingest_export_narrow(const o2a_ingest* g, int32_t score_bits, int32_t bg_bits, void* score_q,
                                        int64_t* exc_off, int32_t* exc_pos, double* exc_val, void* bg_q)
{
    if (!g || (score_bits != 8 && score_bits != 16 && score_bits != 32) || (bg_bits != 8 && bg_bits != 16 && bg_bits != 32)) return -1;
    if (!score_q || !exc_off || !bg_q) return -1;
    const Bases B = file_bases(g);
    const int wi = score_bits == 8 ? 0 : score_bits == 16 ? 1 : 2;
    // exception base of every file, then every file fills its own alignments
    std::vector<int64_t> xbase(g->files.size() + 1, 0);
    for (size_t i = 0; i < g->files.size(); ++i) xbase[i + 1] = xbase[i] + (g->files[i].status ? 0 : g->files[i].misfit[wi]);
    if (xbase.back() > 0 && (!exc_pos || !exc_val)) return -1;
    exc_off[0] = 0;
    o2a_pack::parallel_for(g->n_threads, (int64_t)g->files.size(), [&](int64_t fi) {
        const size_t i = (size_t)fi;
        const FileResult& R = g->files[i];
        if (R.status) return;
        const int64_t a0 = B.a[i], h0 = B.h[i];
        int64_t x = xbase[i], hk = 0;
        for (size_t k = 0; k < R.hsp_count.size(); ++k) {
            const int64_t n = R.hsp_count[k];
            x += o2a_pack::quantize_segment(R.score.data() + hk, n, score_bits, score_q, h0 + hk,
                                            exc_pos ? exc_pos + x : nullptr, exc_val ? exc_val + x : nullptr);
            exc_off[a0 + (int64_t)k + 1] = x;
            hk += n;
        }
        o2a_pack::narrow_bg(R.bg.data(), (int64_t)R.bg.size(), bg_bits, bg_q, B.b[i]);
    });
    return 0;
}

extern "C" int64_t o2a_ingest_span(const o2a_ingest* g, int64_t file, int64_t begin, int64_t end, char* out, int64_t cap)
{
    if (!g || file < 0 || file >= (int64_t)g->files.size()) return -1;
    const std::string& t = g->files[(size_t)file].meta;
    if (begin < 0 || end < begin || end > (int64_t)t.size()) return -1;
    int64_t n = end - begin;
    if (out && cap >= n) memcpy(out, t.data() + begin, (size_t)n);
    return n;
}

// str(qrange.chrom), str(qrange.name), str(srange.chrom) of every alignment (3 strings each, UTF-8) as offsets
// name_off[3*n_aln+1] into `blob`; q_plus[n_aln] = qrange.strand in ('+', '.'); simple[n_aln] = 1 when all three could be
// rendered natively (else the host decodes the range objects with Python's json). Returns the blob size.
extern "C" int64_t o2a_ingest_range_fields(const o2a_ingest* g, int64_t* name_off, char* blob, int64_t cap, uint8_t* q_plus,
                                           uint8_t* simple)
{
    if (!g) return -1;
    int64_t pos = 0, a = 0;
    if (name_off) name_off[0] = 0;
    for (const FileResult& R : g->files) {
        if (R.status) continue;
        const size_t na = R.hsp_count.size();
        if (blob && pos + (int64_t)R.names.size() <= cap) memcpy(blob + pos, R.names.data(), R.names.size());
        for (size_t k = 0; k < na; ++k) {
            if (name_off) for (int t = 1; t <= 3; ++t) name_off[3 * (a + (int64_t)k) + t] = pos + R.name_off[3 * k + (size_t)t];
            if (q_plus) q_plus[a + (int64_t)k] = R.q_plus[k];
            if (simple) simple[a + (int64_t)k] = R.simple[k] == 3 ? 1 : 0;
        }
        pos += (int64_t)R.names.size();
        a += (int64_t)na;
    }
    return pos;
}

// All qrange / srange texts of the batch at once (alignment order: q0, s0, q1, s1, ...): range_off[2*n_aln+1] and,
// when `blob` is given, the concatenated bytes. Returns the blob size.
extern "C" int64_t o2a_ingest_range_texts(const o2a_ingest* g, int64_t* range_off, char* blob, int64_t cap)
{
    if (!g) return -1;
    int64_t pos = 0, k = 0;
    if (range_off) range_off[0] = 0;
    for (const FileResult& R : g->files) {
        if (R.status) continue;
        for (size_t a = 0; a < R.hsp_count.size(); ++a) {
            const int64_t span[4] = {R.qr_begin[a], R.qr_end[a], R.sr_begin[a], R.sr_end[a]};
            for (int w = 0; w < 2; ++w) {
                const int64_t n = span[2 * w + 1] - span[2 * w];
                if (blob && pos + n <= cap) memcpy(blob + pos, R.meta.data() + span[2 * w], (size_t)n);
                pos += n;
                ++k;
                if (range_off) range_off[k] = pos;
            }
        }
    }
    return pos;
}
