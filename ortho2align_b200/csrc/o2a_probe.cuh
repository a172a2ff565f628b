// o2a_probe.cuh — FP64 pipe probes for bench.py (the KDE p-value kernel is bound by the FP64 pipe, and
// MEASURED_PEAKS.json carries no FP64 figure): sustained DFMA rate and sustained ndtr (= 0.5 erfc) evaluation rate of
// the whole device, measured with the same erfc the KDE fitter uses (ndtr_dev, o2a_fit.cuh).
#pragma once
#include "o2a_fit.cuh"

namespace o2a {

// kind 0: 8 independent DFMA chains per thread; kind 1: 4 independent ndtr chains per thread
__global__ void __launch_bounds__(256)
k_fp64_probe(int kind, long long iters, double seed, double* __restrict__ sink)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    double acc = 0.0;
    if (kind == 0) {
        double a[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = seed + 1e-3 * (t + k);
        const double m = 1.0000001, c = 1e-9;
        for (long long i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = fma(a[k], m, c);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += a[k];
    } else {
        double z[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) z[k] = seed + 1e-3 * ((t & 1023) + k) - 2.0;
        for (long long i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 4; ++k) z[k] = ndtr_dev(z[k] * 3.0 - 1.5);      // stays inside (-1.5, 1.5): the erfc main branch
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) acc += z[k];
    }
    if (acc == 123.456) sink[0] = acc;               // never true: keeps the chains alive
}

}  // namespace o2a
