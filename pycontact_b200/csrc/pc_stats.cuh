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
// A key's row of scores lives in shared memory as float64 (frames <= PC_STATS_MAX_FRAMES).  It comes either from
// the dense row-major matrix (scoreArray of each key) or -- ROWS mode -- straight from the key's (frame, score,
// H-bonds) cells as the accumulation emits them (CSR over keys): the matrix is assembled in shared memory and never
// exists in HBM.  Medians are order statistics, found by an MSB-first radix SELECT over the row (8-bit digits, a
// 256-bin histogram per pass, warp-aggregated so that the many equal scores -- zeros -- cost one atomic per warp):
// at most 8 passes over the row instead of the 105 compare-exchange stages of a bitonic sort of 16 384 slots.
// Run lengths come from a chunked scan (every thread owns a contiguous piece of the row and receives the start of
// the run that enters its piece), not from walking back along a run.
#pragma once
#include "pc_common.cuh"

#define PC_STATS_MAX_FRAMES 16384
#define PC_STATS_NT 512
#define PC_STATS_COLS 8
// columns of the output table
enum { PC_STAT_MEAN = 0, PC_STAT_MEDIAN = 1, PC_STAT_HBOND_PCT = 2, PC_STAT_TOTAL_TIME = 3, PC_STAT_MEAN_LIFE = 4,
       PC_STAT_MEDIAN_LIFE = 5, PC_STAT_LIFE_ENTRIES = 6, PC_STAT_FRAMES_ABOVE = 7 };

// order-preserving map float64 -> uint64 (and back)
__device__ __forceinline__ unsigned long long pc_f64_key(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double pc_key_f64(unsigned long long u) {
    const unsigned long long b = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u;
    return __longlong_as_double((long long)b);
}

struct PcStatsShared {
    double scratch[PC_STATS_NT / 32];
    unsigned long long kmin[PC_STATS_NT / 32];
    unsigned int hist[256];
    unsigned int sel[2];
    unsigned int cnt[PC_STATS_NT / 32];
    int carry[PC_STATS_NT / 32];
    unsigned int nruns;
};

__device__ __forceinline__ double pc_block_sum(double x, double *scratch) {
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(PC_FULL_MASK, x, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = x;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += scratch[w];
    return t;
}

// The key of rank `rank` (0-based) among key(0..n): MSB-first radix select, digits of 8 bits from bit top_shift down.
// All threads of the CTA call it with the same arguments and get the same result.
template <typename KeyFn>
__device__ unsigned long long pc_block_select(KeyFn key, const int n, unsigned rank, const int top_shift, PcStatsShared *S) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long prefix = 0ull, mask = 0ull;
    for (int shift = top_shift; shift >= 0; shift -= 8) {
        for (int b = threadIdx.x; b < 256; b += blockDim.x) S->hist[b] = 0u;
        __syncthreads();
        for (int i0 = warp * 32; i0 < n; i0 += blockDim.x) {
            const int i = i0 + lane;
            unsigned d = 0xffffffffu;
            if (i < n) {
                const unsigned long long u = key(i);
                if ((u & mask) == prefix) d = (unsigned)(u >> shift) & 255u;
            }
            const unsigned grp = __match_any_sync(PC_FULL_MASK, d);
            if (d != 0xffffffffu && lane == __ffs(grp) - 1) atomicAdd(&S->hist[d], (unsigned)__popc(grp));
        }
        __syncthreads();
        if (warp == 0) {
            unsigned c[8], s = 0;
#pragma unroll
            for (int q = 0; q < 8; q++) { c[q] = S->hist[lane * 8 + q]; s += c[q]; }
            const unsigned incl = warp_incl_scan(s, lane), excl = incl - s;
            if (rank >= excl && rank < incl) {                    // exactly one lane
                unsigned r = rank - excl;
                int dig = -1;
#pragma unroll
                for (int q = 0; q < 8; q++)
                    if (dig < 0) { if (r < c[q]) dig = q; else r -= c[q]; }
                S->sel[0] = (unsigned)(lane * 8 + dig);
                S->sel[1] = r;
            }
        }
        __syncthreads();
        prefix |= (unsigned long long)S->sel[0] << shift;
        mask |= 0xffull << shift;
        rank = S->sel[1];
        __syncthreads();                                          // hist and sel are rewritten by the next pass
    }
    return prefix;
}

// The two middle keys of key(0..n), n > 0 (equal when n is odd): one select, then one pass that counts the keys
// <= the lower middle and finds the smallest key above it.
template <typename KeyFn>
__device__ void pc_block_middle(KeyFn key, const int n, const int top_shift, PcStatsShared *S, unsigned long long &lo,
                                unsigned long long &hi) {
    lo = pc_block_select(key, n, (unsigned)((n - 1) / 2), top_shift, S);
    hi = lo;
    if (n & 1) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned le = 0;
    unsigned long long mn = ~0ull;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const unsigned long long u = key(i);
        if (u <= lo) le++;
        else if (u < mn) mn = u;
    }
    for (int o = 16; o > 0; o >>= 1) {
        le += __shfl_xor_sync(PC_FULL_MASK, le, o);
        const unsigned long long t = __shfl_xor_sync(PC_FULL_MASK, mn, o);
        mn = t < mn ? t : mn;
    }
    if (lane == 0) { S->cnt[warp] = le; S->kmin[warp] = mn; }
    __syncthreads();
    unsigned tle = 0;
    unsigned long long tmn = ~0ull;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) { tle += S->cnt[w]; tmn = S->kmin[w] < tmn ? S->kmin[w] : tmn; }
    __syncthreads();
    if (tle <= (unsigned)(n / 2)) hi = tmn;                      // rank n/2 lies above every key <= lo
}

// ROWS = false: scores[n_keys][n_frames] float64, hbonds[n_keys][n_frames] (may be null).
// ROWS = true:  key_ptr[n_keys + 1] into row_frame / scores / hbonds (one entry per (key, frame) cell, any order inside a key).
template <bool ROWS>
__global__ void __launch_bounds__(PC_STATS_NT) pc_key_stats_kernel(const double *scores, const uint32_t *hbonds, const long long *key_ptr,
                                                                   const int *row_frame, long long n_keys, int n_frames, int fpad,
                                                                   double ns_per_frame, double threshold, double *out) {
    extern __shared__ __align__(16) unsigned char smem[];
    double *sc = reinterpret_cast<double *>(smem);                                         // fpad scores
    unsigned int *runs = reinterpret_cast<unsigned int *>(smem + sizeof(double) * fpad);   // <= F / 2 + 2 life-time entries
    __shared__ PcStatsShared S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, F = n_frames;
    const int ch = (F + PC_STATS_NT - 1) / PC_STATS_NT;          // frames per thread in the run scan
    for (long long key = blockIdx.x; key < n_keys; key += gridDim.x) {
        double sum = 0.0, above = 0.0, hb = 0.0;
        if (ROWS) {
            for (int i = tid; i < F; i += PC_STATS_NT) sc[i] = 0.0;
            __syncthreads();
            const long long r0 = key_ptr[key], r1 = key_ptr[key + 1];
            for (long long r = r0 + tid; r < r1; r += PC_STATS_NT) {
                const int f = row_frame[r];
                const double v = scores[r];
                if (f >= 0 && f < F) { sc[f] = v; sum += v; if (hbonds && hbonds[r] > 0u) hb += 1.0; }
            }
            __syncthreads();
            for (int i = tid; i < F; i += PC_STATS_NT) above += sc[i] > threshold ? 1.0 : 0.0;
        } else {
            const double *row = scores + key * (long long)F;
            const uint32_t *hrow = hbonds ? hbonds + key * (long long)F : nullptr;
            for (int i = tid; i < F; i += PC_STATS_NT) {
                const double v = row[i];
                sum += v;
                above += v > threshold ? 1.0 : 0.0;
                if (hrow) hb += hrow[i] > 0u ? 1.0 : 0.0;
                sc[i] = v;
            }
        }
        if (tid == 0) S.nruns = 0;
        sum = pc_block_sum(sum, S.scratch);
        above = pc_block_sum(above, S.scratch);
        hb = pc_block_sum(hb, S.scratch);                        // (its barriers also publish sc[] and nruns)
        // ---- life-time entries.  A run starts at a frame above the threshold whose predecessor is not; thread t owns
        //      frames [t * ch, (t + 1) * ch) and needs the last start before its piece: exclusive prefix maximum
        const int b = min(tid * ch, F), e = min(b + ch, F);
        int ls = -1;
        for (int i = b; i < e; i++)
            if (sc[i] > threshold && !(i > 0 && sc[i - 1] > threshold)) ls = i;
        int incl = ls;
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(PC_FULL_MASK, incl, o);
            if (lane >= o) incl = max(incl, t);
        }
        if (lane == 31) S.carry[warp] = incl;
        __syncthreads();
        int cur = __shfl_up_sync(PC_FULL_MASK, incl, 1);
        if (lane == 0) cur = -1;
        for (int w = 0; w < warp; w++) cur = max(cur, S.carry[w]);
        for (int i = b; i < e; i++) {
            const bool up = sc[i] > threshold;
            if (up && !(i > 0 && sc[i - 1] > threshold)) cur = i;
            if (up && (i == F - 1 || !(sc[i + 1] > threshold))) runs[atomicAdd(&S.nruns, 1u)] = (unsigned)(i - cur + 1);
            if (!up && i == F - 1) runs[atomicAdd(&S.nruns, 1u)] = 0u;      // the trailing entry of an inactive contact
        }
        __syncthreads();
        const int L = (int)S.nruns;
        double med_life = 0.0, med = 0.0;
        if (L > 0) {
            unsigned long long lo, hi;
            pc_block_middle([&](int i) { return (unsigned long long)runs[i]; }, L, 8, &S, lo, hi);      // run lengths < 2^16
            med_life = 0.5 * ((double)lo + (double)hi);          // exact: integers below 2^15
        }
        if (F > 0) {
            unsigned long long lo, hi;
            pc_block_middle([&](int i) { return pc_f64_key(sc[i]); }, F, 56, &S, lo, hi);
            med = lo == hi ? pc_key_f64(lo) : 0.5 * (pc_key_f64(lo) + pc_key_f64(hi));
        }
        if (tid == 0) {
            double *o = out + key * PC_STATS_COLS;
            const double nan = __longlong_as_double(0x7ff8000000000000ll);
            o[PC_STAT_MEAN] = F > 0 ? sum / (double)F : nan;
            o[PC_STAT_MEDIAN] = F > 0 ? med : nan;
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
