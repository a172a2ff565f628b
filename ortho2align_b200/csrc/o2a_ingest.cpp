// o2a_ingest.cpp — native ingest of the reference's per-lncRNA JSON files into the columnar batch
// of include/o2a.h (the reference spends 94 % of build_orthologs in
// json.load + HSP/Alignment object construction, pipeline.py:743-754, genomicranges.py:1084-1106,
// alignment.py:89-110, 233-249).
//
// align_files/<name>.json = JSON list of alignment dicts (to_dict: genomicranges.py:1077-1082,
// alignment.py:588-601); bg_files/<name>.json = JSON list of ints (pipeline.py:266-273).
// What is extracted per alignment: the HSPs of key "HSPs" ONLY (the reference rebuilds the working
// list from 'HSPs' and ignores 'filtered_HSPs', genomicranges.py:1094-1100), qlen/slen (null ->
// max+1 like alignment.py:477-487), and the BYTE SPANS of the "qrange"/"srange" objects so the
// host decodes only those few hundred bytes when it formats a record. Scores keep their JSON type
// (int vs float) in `score_is_int` because the reference's record text depends on it.
// Background patches of pipeline.py:746-754 are applied here, in the reference's order.
//
// Plain C ABI (declared in include/o2a_ingest.h); files are parsed by a pool of std::threads.
#include <atomic>
#include <charconv>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/o2a_ingest.h"

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
    std::string text;               // file contents; released after parsing
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
};

inline bool key_is(const char* b, const char* e, const char* name) {
    size_t n = strlen(name);
    return (size_t)(e - b) == n && !memcmp(b, name, n);
}

bool read_file(const char* path, std::string& out, bool& missing) {
    missing = false;
    FILE* f = fopen(path, "rb");
    if (!f) { missing = true; return false; }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize((size_t)(n > 0 ? n : 0));
    size_t got = n > 0 ? fread(&out[0], 1, (size_t)n, f) : 0;
    fclose(f);
    return got == (size_t)(n > 0 ? n : 0);
}

bool parse_hsps(Parser& P, FileResult& R, int64_t& count, int64_t& max_qend, int64_t& max_s, bool& range_err) {
    count = 0; max_qend = 0; max_s = 0;
    if (!P.expect('[')) return false;
    if (P.lit(']')) return true;
    do {
        if (!P.expect('{')) return false;
        int64_t c[4] = {0, 0, 0, 0}; bool have[5] = {false, false, false, false, false};
        double sc = 0; uint8_t isint = 1;
        if (!P.lit('}')) {
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
        if (count == 0 || c[1] > max_qend) max_qend = c[1];
        int64_t ms = c[2] > c[3] ? c[2] : c[3];
        if (count == 0 || ms > max_s) max_s = ms;
        ++count;
    } while (P.lit(','));
    return P.expect(']');
}

void parse_alignment_file(FileResult& R) {
    Parser P{R.text.data(), R.text.data(), R.text.data() + R.text.size(), {}};
    bool range_err = false;
    if (!P.expect('[')) { R.status = 1; R.error = P.err; return; }
    if (!P.lit(']')) {
        do {
            if (!P.expect('{')) break;
            int64_t cnt = 0, mq = 0, msub = 0, ql = 0, sl = 0; bool hq = false, hs = false, hh = false;
            int64_t qb = -1, qe = -1, sb = -1, se = -1;
            if (!P.lit('}')) {
                do {
                    const char *kb, *ke;
                    if (!P.str(kb, ke) || !P.expect(':')) break;
                    if (key_is(kb, ke, "HSPs")) { if (!parse_hsps(P, R, cnt, mq, msub, range_err)) break; hh = true; }
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
        if (!P.expect('[')) { R.status = 1; R.error = "background: " + P.err; return; }
        if (!P.lit(']')) {
            do {
                int kind; int64_t iv; double dv;
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
    for (int64_t x : v) {
        if (x > INT32_MAX || x < -(int64_t)INT32_MAX) { R.status = 2; R.error = "background scores must be integers within int32"; return; }
        R.bg.push_back((int32_t)x);
    }
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

extern "C" void o2a_ingest_destroy(o2a_ingest* g) { delete g; }

extern "C" const char* o2a_ingest_file_error(const o2a_ingest* g, int64_t i)
{
    if (!g || i < 0 || i >= (int64_t)g->files.size()) return "";
    return g->files[(size_t)i].error.c_str();
}

extern "C" int o2a_ingest_load(o2a_ingest* g, const char* const* align_paths, const char* const* bg_paths, int64_t n_files)
{
    if (!g || n_files < 0) return -1;
    g->files.assign((size_t)n_files, FileResult());
    std::atomic<int64_t> next(0);
    auto work = [&]() {
        for (;;) {
            int64_t i = next.fetch_add(1);
            if (i >= n_files) break;
            FileResult& R = g->files[(size_t)i];
            bool missing = false;
            if (!read_file(align_paths[i], R.text, missing)) { R.status = 1; R.error = "cannot read alignment file"; continue; }
            parse_alignment_file(R);
            std::string().swap(R.text);
            if (R.status) continue;
            std::string bgtext;
            bool bg_missing = false;
            bool ok = read_file(bg_paths[i], bgtext, bg_missing);
            if (!ok && !bg_missing) { R.status = 1; R.error = "cannot read background file"; continue; }
            parse_background(bgtext, bg_missing, R);
        }
    };
    int nt = g->n_threads;
    if (nt > n_files) nt = (int)(n_files > 0 ? n_files : 1);
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
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

extern "C" int o2a_ingest_export(const o2a_ingest* g, int32_t* file_status, int64_t* lnc_file,
                                 int64_t* bg_off, int32_t* bg, int64_t* aln_off, int64_t* hsp_off,
                                 int64_t* qlen, int64_t* slen, int32_t* coords,
                                 int32_t* qstart, int32_t* qend, int32_t* sstart, int32_t* send,
                                 double* score, uint8_t* score_is_int, int64_t* spans)
{
    if (!g) return -1;
    int64_t l = 0, a = 0, h = 0, b = 0;
    if (bg_off) bg_off[0] = 0; if (aln_off) aln_off[0] = 0; if (hsp_off) hsp_off[0] = 0;
    for (size_t i = 0; i < g->files.size(); ++i) {
        const FileResult& R = g->files[i];
        if (file_status) file_status[i] = R.status;
        if (R.status) continue;
        if (lnc_file) lnc_file[l] = (int64_t)i;
        const int64_t na = (int64_t)R.hsp_count.size(), nh = (int64_t)R.score.size(), nb = (int64_t)R.bg.size();
        if (bg) memcpy(bg + b, R.bg.data(), (size_t)nb * 4);
        int64_t hcur = h;
        for (int64_t k = 0; k < na; ++k) {
            hcur += R.hsp_count[(size_t)k];
            if (hsp_off) hsp_off[a + k + 1] = hcur;
            if (qlen) qlen[a + k] = R.qlen[(size_t)k];
            if (slen) slen[a + k] = R.slen[(size_t)k];
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
        ++l; a += na; h += nh; b += nb;
        if (bg_off) bg_off[l] = b;
        if (aln_off) aln_off[l] = a;
    }
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
