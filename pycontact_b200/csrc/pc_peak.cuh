// pc_peak.cuh -- FP32 issue-rate microbenchmark (bench.py's second roof): every thread runs
// eight independent FFMA chains, so the FP32 pipe of each SM sub-partition always has an instruction to issue.
// The measured rate (FFMA instructions / s over the whole GPU) divided by the instructions of one pair
// evaluation is the pair-evaluation roof the search kernels are reported against.
#pragma once
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) pc_ffma_chain_kernel(float *out, const int iters, const float a, const float b) {
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f,
          x7 = x0 + 7.f;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    const float s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456f) out[0] = s;          // keeps the chains alive without a store on the common path
}
