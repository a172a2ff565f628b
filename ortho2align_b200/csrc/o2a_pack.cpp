// o2a_pack.cpp — columnar batch (float64 scores, SoA coordinates, int32 background) -> the PCIe-optimal
// lossless host layout of include/o2a.h (narrow `score_q` + per-alignment exception lists, narrow `bg`,
// AoS `coords`), multi-threaded, written into caller-provided (normally pinned) buffers.
// The JSON ingest produces the same layout directly (o2a_ingest_export_narrow); this entry point serves
// callers that already hold columnar arrays: lncRNA shards cut for several GPUs, the producer side, bench.py.
// Host-only C++ (part of libo2a_ingest.so), declared in include/o2a_ingest.h.
#include <cstring>
#include <vector>

#include "../../include/o2a_ingest.h"
#include "o2a_pack.h"

namespace {
// [lo, hi) item ranges of roughly equal size
std::vector<int64_t> blocks_of(int64_t n, int64_t want)
{
    if (want < 1) want = 1;
    if (want > n) want = n > 0 ? n : 1;
    std::vector<int64_t> b((size_t)want + 1);
    for (int64_t k = 0; k <= want; ++k) b[(size_t)k] = n * k / want;
    return b;
}
}  // namespace

extern "C" int o2a_ingest_pack_plan(const double* score, int64_t n_hsp, const int32_t* bg, int64_t n_bg, int n_threads,
                                    int32_t* score_bits, int32_t* bg_bits, int64_t* n_exc)
{
    if ((n_hsp > 0 && !score) || (n_bg > 0 && !bg) || n_hsp < 0 || n_bg < 0) return -1;
    const int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    const std::vector<int64_t> hb = blocks_of(n_hsp, (int64_t)nt * 4), bb = blocks_of(n_bg, (int64_t)nt * 4);
    const int64_t nhb = (int64_t)hb.size() - 1, nbb = (int64_t)bb.size() - 1;
    std::vector<int64_t> mis((size_t)nhb * 3, 0), mn((size_t)nbb, 0), mx((size_t)nbb, 0);
    o2a_pack::parallel_for(nt, nhb + nbb, [&](int64_t t) {
        if (t < nhb) {
            int64_t m0 = 0, m1 = 0, m2 = 0;
            for (int64_t k = hb[(size_t)t]; k < hb[(size_t)t + 1]; ++k) {
                const unsigned mf = o2a_pack::narrow_misfit(score[k]);
                m0 += mf & 1u; m1 += (mf >> 1) & 1u; m2 += (mf >> 2) & 1u;
            }
            mis[(size_t)t * 3] = m0; mis[(size_t)t * 3 + 1] = m1; mis[(size_t)t * 3 + 2] = m2;
        } else {
            const int64_t u = t - nhb;
            int32_t a = 0, b = 0;
            for (int64_t k = bb[(size_t)u]; k < bb[(size_t)u + 1]; ++k) { if (bg[k] < a) a = bg[k]; if (bg[k] > b) b = bg[k]; }
            mn[(size_t)u] = a; mx[(size_t)u] = b;
        }
    });
    int64_t tot[3] = {0, 0, 0}, lo = 0, hi = 0;
    for (int64_t t = 0; t < nhb; ++t) for (int k = 0; k < 3; ++k) tot[k] += mis[(size_t)t * 3 + k];
    for (int64_t u = 0; u < nbb; ++u) { if (mn[(size_t)u] < lo) lo = mn[(size_t)u]; if (mx[(size_t)u] > hi) hi = mx[(size_t)u]; }
    if (score_bits) *score_bits = o2a_pack::choose_score_bits(tot, n_hsp);
    if (bg_bits) *bg_bits = o2a_pack::choose_bg_bits(lo, hi);
    if (n_exc) { n_exc[0] = tot[0]; n_exc[1] = tot[1]; n_exc[2] = tot[2]; }
    return 0;
}

extern "C" int o2a_ingest_pack_narrow(int64_t n_aln, const int64_t* hsp_off, const double* score, int32_t score_bits,
                                      void* score_q, int64_t* exc_off, int32_t* exc_pos, double* exc_val,
                                      const int32_t* bg, int64_t n_bg, int32_t bg_bits, void* bg_q,
                                      const int32_t* qstart, const int32_t* qend, const int32_t* sstart,
                                      const int32_t* send, int32_t* coords, int n_threads)
{
    if (n_aln < 0 || !hsp_off) return -1;
    const int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    const int64_t n_hsp = hsp_off[n_aln];
    if (score_q) {
        if (!score || !exc_off || (score_bits != 8 && score_bits != 16 && score_bits != 32)) return -1;
        const unsigned bit = score_bits == 8 ? 1u : score_bits == 16 ? 2u : 4u;
        const std::vector<int64_t> ab = blocks_of(n_aln, (int64_t)nt * 8);
        const int64_t nab = (int64_t)ab.size() - 1;
        // pass 1: exceptions per alignment; sequential prefix; pass 2: fill
        exc_off[0] = 0;
        o2a_pack::parallel_for(nt, nab, [&](int64_t t) {
            for (int64_t a = ab[(size_t)t]; a < ab[(size_t)t + 1]; ++a) {
                int64_t nx = 0;
                for (int64_t k = hsp_off[a]; k < hsp_off[a + 1]; ++k) nx += (o2a_pack::narrow_misfit(score[k]) & bit) != 0;
                exc_off[a + 1] = nx;
            }
        });
        for (int64_t a = 0; a < n_aln; ++a) exc_off[a + 1] += exc_off[a];
        if (exc_off[n_aln] > 0 && (!exc_pos || !exc_val)) return -1;
        o2a_pack::parallel_for(nt, nab, [&](int64_t t) {
            for (int64_t a = ab[(size_t)t]; a < ab[(size_t)t + 1]; ++a) {
                const int64_t x = exc_off[a];
                o2a_pack::quantize_segment(score + hsp_off[a], hsp_off[a + 1] - hsp_off[a], score_bits, score_q, hsp_off[a],
                                           exc_pos ? exc_pos + x : nullptr, exc_val ? exc_val + x : nullptr);
            }
        });
    }
    if (bg_q) {
        if ((n_bg > 0 && !bg) || (bg_bits != 8 && bg_bits != 16 && bg_bits != 32)) return -1;
        const std::vector<int64_t> bb = blocks_of(n_bg, (int64_t)nt * 4);
        o2a_pack::parallel_for(nt, (int64_t)bb.size() - 1, [&](int64_t u) {
            o2a_pack::narrow_bg(bg + bb[(size_t)u], bb[(size_t)u + 1] - bb[(size_t)u], bg_bits, bg_q, bb[(size_t)u]);
        });
    }
    if (coords) {
        if (n_hsp > 0 && (!qstart || !qend || !sstart || !send)) return -1;
        const std::vector<int64_t> hb = blocks_of(n_hsp, (int64_t)nt * 4);
        o2a_pack::parallel_for(nt, (int64_t)hb.size() - 1, [&](int64_t t) {
            for (int64_t k = hb[(size_t)t]; k < hb[(size_t)t + 1]; ++k) {
                int32_t* c = coords + 4 * k;
                c[0] = qstart[k]; c[1] = qend[k]; c[2] = sstart[k]; c[3] = send[k];
            }
        });
    }
    return 0;
}
