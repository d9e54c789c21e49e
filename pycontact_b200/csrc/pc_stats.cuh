// pc_stats.cuh -- per-key statistics over the (keys x frames) score matrix, one CTA per key.
//
// Replaces, for all AccumulatedContacts at once, the per-object Python loops of
//   mean_score (core/Biochemistry.py:179-186), median_score (:188-192), hbond_percentage (:150-158),
//   total_time (:160-167), life_time / mean_life_time / median_life_time (:169-213),
// which every filter, sort and export of the reference walks again (core/ContactFilters.py:185-287).
//
// life_time's list, restated: one entry per maximal run of frames with score > threshold, in units of
// ns_per_frame, plus an entry 0 when the last frame is not above the threshold (the reference appends the
// running contactTime at the last frame whatever it is, :203-205).  sum(entries) == frames above.
//
// A key's row of scores lives in shared memory as float64 (frames <= PC_STATS_MAX_FRAMES); medians come
// from a bitonic sort of that row / of the run list.
#pragma once
#include "pc_common.cuh"

#define PC_STATS_MAX_FRAMES 16384
#define PC_STATS_NT 512
#define PC_STATS_COLS 8
// columns of the output table
enum { PC_STAT_MEAN = 0, PC_STAT_MEDIAN = 1, PC_STAT_HBOND_PCT = 2, PC_STAT_TOTAL_TIME = 3, PC_STAT_MEAN_LIFE = 4,
       PC_STAT_MEDIAN_LIFE = 5, PC_STAT_LIFE_ENTRIES = 6, PC_STAT_FRAMES_ABOVE = 7 };

template <typename T>
__device__ __forceinline__ void pc_bitonic_sort_smem(T *v, int npow2) {
    for (int k = 2; k <= npow2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
                const int p = i ^ j;
                if (p > i) {
                    const T a = v[i], b = v[p];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { v[i] = b; v[p] = a; }
                }
            }
            __syncthreads();
        }
}

__device__ __forceinline__ double pc_block_sum(double x, double *scratch) {
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(PC_FULL_MASK, x, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = x;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += scratch[w];
    return t;
}

__global__ void __launch_bounds__(PC_STATS_NT) pc_key_stats_kernel(const double *scores, const uint32_t *hbonds, long long n_keys,
                                                                   int n_frames, int fpad, int rpad, double ns_per_frame,
                                                                   double threshold, double *out) {
    extern __shared__ __align__(16) unsigned char smem[];
    double *sc = reinterpret_cast<double *>(smem);                                 // fpad scores
    unsigned int *runs = reinterpret_cast<unsigned int *>(smem + sizeof(double) * fpad);   // rpad life-time entries
    __shared__ double scratch[PC_STATS_NT / 32];
    __shared__ unsigned int s_nruns;
    const int tid = threadIdx.x, F = n_frames;
    for (long long key = blockIdx.x; key < n_keys; key += gridDim.x) {
        const double *row = scores + key * (long long)F;
        const uint32_t *hrow = hbonds ? hbonds + key * (long long)F : nullptr;
        double sum = 0.0, above = 0.0, hb = 0.0;
        for (int i = tid; i < fpad; i += PC_STATS_NT) {
            double v = INFINITY;
            if (i < F) {
                v = row[i];
                sum += v;
                above += v > threshold ? 1.0 : 0.0;
                if (hrow) hb += hrow[i] > 0u ? 1.0 : 0.0;
            }
            sc[i] = v;
        }
        for (int i = tid; i < rpad; i += PC_STATS_NT) runs[i] = 0xffffffffu;
        if (tid == 0) s_nruns = 0;
        sum = pc_block_sum(sum, scratch);
        above = pc_block_sum(above, scratch);
        hb = pc_block_sum(hb, scratch);
        // ---- life-time entries: a thread that sits on the last frame of a run walks back to its start
        for (int i = tid; i < F; i += PC_STATS_NT) {
            if (!(sc[i] > threshold)) {
                if (i == F - 1) runs[atomicAdd(&s_nruns, 1u)] = 0u;          // the trailing entry of an inactive contact
                continue;
            }
            if (i + 1 < F && sc[i + 1] > threshold) continue;                // not the end of its run
            int s = i;
            while (s > 0 && sc[s - 1] > threshold) s--;
            runs[atomicAdd(&s_nruns, 1u)] = (unsigned)(i - s + 1);
        }
        __syncthreads();
        const unsigned L = s_nruns;
        pc_bitonic_sort_smem(runs, rpad);                                    // unused slots (0xffffffff) sort to the end
        double med_life = 0.0;
        if (L > 0) med_life = (L & 1u) ? (double)runs[L / 2] : 0.5 * ((double)runs[L / 2 - 1] + (double)runs[L / 2]);
        __syncthreads();
        pc_bitonic_sort_smem(sc, fpad);                                      // the +inf padding sorts to the end
        if (tid == 0) {
            double *o = out + key * PC_STATS_COLS;
            const double nan = __longlong_as_double(0x7ff8000000000000ll);
            o[PC_STAT_MEAN] = F > 0 ? sum / (double)F : nan;
            o[PC_STAT_MEDIAN] = F > 0 ? ((F & 1) ? sc[F / 2] : 0.5 * (sc[F / 2 - 1] + sc[F / 2])) : nan;
            o[PC_STAT_HBOND_PCT] = F > 0 ? hb / (double)F * 100.0 : nan;
            o[PC_STAT_TOTAL_TIME] = above * ns_per_frame;
            o[PC_STAT_MEAN_LIFE] = L > 0 ? above * ns_per_frame / (double)L : nan;
            o[PC_STAT_MEDIAN_LIFE] = L > 0 ? med_life * ns_per_frame : nan;
            o[PC_STAT_LIFE_ENTRIES] = (double)L;
            o[PC_STAT_FRAMES_ABOVE] = above;
        }
        __syncthreads();
    }
}
