// o2a_pack.h — host-side helpers shared by the JSON ingest (o2a_ingest.cpp) and the columnar packer
// (o2a_pack.cpp): the lossless narrow encodings of include/o2a.h (`score_q` + exception lists,
// narrow `bg`) and a small thread-pool loop. Host-only C++, no CUDA.
#pragma once
#include <atomic>
#include <cmath>
#include <cstdint>
#include <thread>
#include <vector>

namespace o2a_pack {

// run fn(i) for i in [0, n) on up to n_threads std::threads (dynamic distribution; the caller is one of them)
template <typename F>
inline void parallel_for(int n_threads, int64_t n, F fn)
{
    std::atomic<int64_t> next(0);
    auto work = [&]() { for (;;) { int64_t i = next.fetch_add(1); if (i >= n) break; fn(i); } };
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (nt > n) nt = (int)(n > 0 ? n : 1);
    if (nt < 1) nt = 1;
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
}

// Which narrow encodings cannot hold score `sc` (bit 0: int8, bit 1: int16, bit 2: int32). A score is
// representable when it is integral and inside (min, max] of the type — the most negative value is the
// sentinel that sends the HSP to its alignment's exception list (include/o2a.h, `score_bits`).
inline unsigned narrow_misfit(double sc)
{
    if (sc != sc || std::fabs(sc) >= 4294967296.0) return 7u;
    if (sc != (double)(int64_t)sc) return 7u;                      // fractional: a cut HSP (alignment.py:330)
    if (sc == 0.0 && std::signbit(sc)) return 7u;                  // -0.0 keeps its sign
    unsigned m = 0;
    if (!(sc > -128.0 && sc <= 127.0)) m |= 1u;
    if (!(sc > -32768.0 && sc <= 32767.0)) m |= 2u;
    if (!(sc > -2147483648.0 && sc <= 2147483647.0)) m |= 4u;
    return m;
}

// smallest width that sends at most 10 % of the HSPs to the exception lists (12 bytes each)
inline int choose_score_bits(const int64_t misfit[3], int64_t n_hsp)
{
    if (misfit[0] * 10 <= n_hsp) return 8;
    if (misfit[1] * 10 <= n_hsp) return 16;
    return 32;
}

inline int choose_bg_bits(int64_t mn, int64_t mx)
{
    if (mn >= -128 && mx <= 127) return 8;
    if (mn >= -32768 && mx <= 32767) return 16;
    return 32;
}

// scores[0..n) of ONE alignment -> score_q[base..base+n) of the chosen width; exceptions (position inside the
// alignment, exact float64) appended to exc_pos / exc_val. Returns the number of exceptions.
inline int64_t quantize_segment(const double* sc, int64_t n, int bits, void* score_q, int64_t base,
                                int32_t* exc_pos, double* exc_val)
{
    int64_t nx = 0;
    const unsigned bit = bits == 8 ? 1u : bits == 16 ? 2u : 4u;
    for (int64_t k = 0; k < n; ++k) {
        const double v = sc[k];
        const bool bad = (narrow_misfit(v) & bit) != 0;
        if (bad) { if (exc_pos) { exc_pos[nx] = (int32_t)k; exc_val[nx] = v; } ++nx; }
        if (bits == 8) ((int8_t*)score_q)[base + k] = bad ? (int8_t)INT8_MIN : (int8_t)v;
        else if (bits == 16) ((int16_t*)score_q)[base + k] = bad ? (int16_t)INT16_MIN : (int16_t)v;
        else ((int32_t*)score_q)[base + k] = bad ? INT32_MIN : (int32_t)v;
    }
    return nx;
}

inline void narrow_bg(const int32_t* bg, int64_t n, int bits, void* out, int64_t base)
{
    if (bits == 8) { int8_t* o = (int8_t*)out + base; for (int64_t k = 0; k < n; ++k) o[k] = (int8_t)bg[k]; }
    else if (bits == 16) { int16_t* o = (int16_t*)out + base; for (int64_t k = 0; k < n; ++k) o[k] = (int16_t)bg[k]; }
    else { int32_t* o = (int32_t*)out + base; for (int64_t k = 0; k < n; ++k) o[k] = bg[k]; }
}

}  // namespace o2a_pack
