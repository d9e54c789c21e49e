// pc_accumulate.cuh -- per-frame accumulation of contacts into (frame, key) rows.
//
// Replaces the first loop of analyze_contactResultsWithMaps (core/ContactAnalyzer.py:527-559):
// for every frame, group its contacts by key = (properties of atom1 under map1, properties of
// atom2 under map2) and sum fscore / bb1score / bb2score, count contributing contacts and their
// hydrogen bonds (TempContactAccumulate, core/Biochemistry.py:283-296).  Keys are integers
// gid1*G2 + gid2 where the group ids come from the static topology (pc_set_maps); the string keys
// of the reference ("r.14rn.VAL-r.22rn.ILE") are rebuilt by the Python layer from the group ids.
//
// One CTA per frame, open-addressing hash table in shared memory, float64 sums like the reference.
// Frames whose keys do not fit the table set PC_ST_TABLE_OVERFLOW (the host retries with a bigger
// table).
#pragma once
#include "pc_common.cuh"

struct PcAccArgs {
    const pc_contact *contacts;
    const pc_hbond *hbonds;
    const unsigned long long *order_keys;   // may be NULL
    const uint32_t *frame_cstart, *frame_ccount, *frame_hstart, *frame_hcount;
    const int *lgid1, *lgid2;               // per local index
    const uint8_t *lflag1, *lflag2;
    unsigned long long ngroups2;
    int nframes, cap;                       // cap = table slots, power of two
    pc_row *rows;
    long long cap_rows;
    unsigned long long *counters;
    uint32_t *frame_rstart, *frame_rcount;
};

#define PC_EMPTY_KEY 0xffffffffffffffffull

__device__ __forceinline__ unsigned pc_hash64(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (unsigned)k;
}

template <int NT>
__global__ void __launch_bounds__(NT) pc_accumulate_kernel(const PcAccArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int cap = a.cap;
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem);
    double *score = reinterpret_cast<double *>(keys + cap);
    double *bb1 = score + cap;
    double *bb2 = bb1 + cap;
    unsigned long long *first = reinterpret_cast<unsigned long long *>(bb2 + cap);
    unsigned int *ncont = reinterpret_cast<unsigned int *>(first + cap);
    unsigned int *nhb = ncont + cap;
    __shared__ uint32_t wsum[32];
    __shared__ long long s_base;
    __shared__ unsigned int s_status;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned mask = (unsigned)cap - 1u;

    for (int f = blockIdx.x; f < a.nframes; f += gridDim.x) {
        for (int s = tid; s < cap; s += NT) {
            keys[s] = PC_EMPTY_KEY; score[s] = 0.0; bb1[s] = 0.0; bb2[s] = 0.0;
            first[s] = PC_EMPTY_KEY; ncont[s] = 0; nhb[s] = 0;
        }
        if (tid == 0) s_status = 0;
        __syncthreads();
        const uint32_t c0 = a.frame_cstart[f], cn = a.frame_ccount[f];
        for (uint32_t c = tid; c < cn; c += NT) {
            const pc_contact rec = a.contacts[c0 + c];
            const unsigned long long key =
                (unsigned long long)__ldg(a.lgid1 + rec.i) * a.ngroups2 + (unsigned long long)__ldg(a.lgid2 + rec.j);
            unsigned s = pc_hash64(key) & mask;
            int probes = 0;
            for (;;) {
                const unsigned long long prev = atomicCAS(&keys[s], PC_EMPTY_KEY, key);
                if (prev == PC_EMPTY_KEY || prev == key) break;
                s = (s + 1) & mask;
                if (++probes >= cap) { s = 0xffffffffu; break; }
            }
            if (s == 0xffffffffu) { atomicOr(&s_status, (unsigned)PC_ST_TABLE_OVERFLOW); continue; }
            const double w = (double)rec.weight;
            atomicAdd(&score[s], w);
            if (__ldg(a.lflag1 + rec.i) & PC_ATOM_BACKBONE) atomicAdd(&bb1[s], w);
            if (__ldg(a.lflag2 + rec.j) & PC_ATOM_BACKBONE) atomicAdd(&bb2[s], w);
            atomicAdd(&ncont[s], 1u);
            const unsigned long long ok = a.order_keys ? a.order_keys[c0 + c]
                                                       : (((unsigned long long)(unsigned)rec.i << 32) | (unsigned)rec.j);
            atomicMin(&first[s], ok);
        }
        __syncthreads();
        const uint32_t h0 = a.frame_hstart[f], hn = a.frame_hcount[f];
        for (uint32_t h = tid; h < hn; h += NT) {
            const pc_hbond rec = a.hbonds[h0 + h];
            const unsigned long long key =
                (unsigned long long)__ldg(a.lgid1 + rec.i) * a.ngroups2 + (unsigned long long)__ldg(a.lgid2 + rec.j);
            unsigned s = pc_hash64(key) & mask;
            for (int probes = 0; probes < cap; probes++) {
                const unsigned long long k = keys[s];
                if (k == key) { atomicAdd(&nhb[s], 1u); break; }
                if (k == PC_EMPTY_KEY) break;          // only after a table overflow
                s = (s + 1) & mask;
            }
        }
        __syncthreads();
        // compact the occupied slots into the batch's row buffer
        const int ch = (cap + NT - 1) / NT;
        const int b = min(tid * ch, cap), e = min(b + ch, cap);
        uint32_t cnt = 0;
        for (int s = b; s < e; s++) cnt += keys[s] != PC_EMPTY_KEY;
        const uint32_t incl = warp_incl_scan(cnt, lane);
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const uint32_t v = lane < NT / 32 ? wsum[lane] : 0;
            const uint32_t vi = warp_incl_scan(v, lane);
            wsum[lane] = vi - v;
            const uint32_t total = __shfl_sync(PC_FULL_MASK, vi, 31);
            if (lane == 0) {
                long long base = (long long)atomicAdd(&a.counters[PC_CNT_ROWS], (unsigned long long)total);
                if (base + total > a.cap_rows) { atomicOr(&s_status, (unsigned)PC_ST_OUT_OVERFLOW); base = -1; }
                s_base = base;
                a.frame_rstart[f] = (uint32_t)(base < 0 ? 0 : base);
                a.frame_rcount[f] = base < 0 ? 0u : total;
            }
        }
        __syncthreads();
        if (s_base >= 0) {
            long long o = s_base + wsum[warp] + incl - cnt;
            for (int s = b; s < e; s++) {
                if (keys[s] == PC_EMPTY_KEY) continue;
                pc_row r;
                r.key = keys[s]; r.score = score[s]; r.bb1 = bb1[s]; r.bb2 = bb2[s];
                r.ncontacts = ncont[s]; r.nhbonds = nhb[s]; r.first = first[s];
                a.rows[o++] = r;
            }
        }
        if (tid == 0 && s_status) atomicOr(&a.counters[PC_CNT_STATUS], (unsigned long long)s_status);
        __syncthreads();
    }
}

// Time-aggregated maps: fold the (frame, key) rows of a batch into a dense per-key table
// [n_keys][4] = (sum of scores, sum of bb1, sum of bb2, number of frames with an H-bond) --
// the sums behind AccumulatedContact.mean_score / bb1 / bb2 / hbond_percentage
// (core/Biochemistry.py:150-186, core/ContactAnalyzer.py:588-591).  This is the buffer the ranks
// all-reduce at the end of a frame-sharded run.
__global__ void __launch_bounds__(256) pc_fold_rows_kernel(const pc_row *rows, const unsigned long long *counters,
                                                           long long cap_rows, double *totals,
                                                           unsigned long long n_keys) {
    const long long n = min((long long)counters[PC_CNT_ROWS], cap_rows);
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
        const pc_row row = rows[r];
        if (row.key >= n_keys) continue;
        double *t = totals + 4ull * row.key;
        atomicAdd(t + 0, row.score);
        if (row.bb1 != 0.0) atomicAdd(t + 1, row.bb1);
        if (row.bb2 != 0.0) atomicAdd(t + 2, row.bb2);
        if (row.nhbonds) atomicAdd(t + 3, 1.0);
    }
}
