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
#include "pc_scan_small.cuh"   // PC_EMPTY_KEY, pc_hash64

struct PcAccArgs {
    const pc_contact *contacts;
    const pc_hbond *hbonds;
    const unsigned long long *order_keys;   // may be NULL
    const uint32_t *frame_cstart, *frame_ccount, *frame_hstart, *frame_hcount;
    const int *lgid1, *lgid2;               // per local index
    const uint8_t *lflag1, *lflag2;
    const uint32_t *lgb1, *lgb2;            // group id | backbone << 31 per local index (one load instead of two)
    unsigned long long ngroups2;
    int nframes, cap;                       // cap = table slots, power of two
    pc_row *rows;
    long long cap_rows;
    unsigned long long *counters;
    uint32_t *frame_rstart, *frame_rcount;
};

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
            // a full table makes every further key walk 256 slots, and the batch is redone on the global table
            // anyway (PC_ST_TABLE_OVERFLOW): stop at the first overflow instead of finishing the frame
            if (*(volatile unsigned int *)&s_status & (unsigned)PC_ST_TABLE_OVERFLOW) break;
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
    // a batch that overflowed a buffer is re-run by the caller: do not fold its partial rows
    if (counters[PC_CNT_STATUS] & PC_ST_RERUN_MASK) return;
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

// ---------------------------------------------------------------------------------------------------
// Mid-size frames (thousands of keys per frame, e.g. protein vs. solvent at 1 M atoms): still one CTA
// per frame and a shared-memory table, but with 28-byte slots -- key, three float32 sums, two 32-bit
// counts -- so that 8192 slots fit one SM.  float32 sums stay within ~1e-6 relative of the float64 sum
// while a (frame, key) cell holds at most PC_FLOAT_SUM_MAX contacts; a cell that grows beyond it (coarse
// maps on large frames) raises PC_ST_PRECISION and the caller redoes the batch on the global-table path,
// whose sums are exact.  No first-appearance key: only used when no order keys were requested.
// pc_compact_frame is the body for one frame; the large-frame tail kernel (pc_scan_tail.cuh) runs it on
// the frame it has just searched.
// ---------------------------------------------------------------------------------------------------
#define PC_COMPACT_SLOTS 8192
#define PC_COMPACT_SLOT_BYTES 28
#define PC_COMPACT_SMEM (PC_COMPACT_SLOTS * PC_COMPACT_SLOT_BYTES)

template <int NT>
__device__ __forceinline__ void pc_compact_frame(const PcAccArgs &a, const int f, unsigned char *smem, uint32_t *wsum,
                                                 long long *s_base, unsigned int *s_status) {
    constexpr int cap = PC_COMPACT_SLOTS;
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem);
    float *score = reinterpret_cast<float *>(keys + cap);
    float *bb1 = score + cap;
    float *bb2 = bb1 + cap;
    unsigned int *ncont = reinterpret_cast<unsigned int *>(bb2 + cap);
    unsigned int *nhb = ncont + cap;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr unsigned mask = (unsigned)cap - 1u;
    for (int s = tid; s < cap; s += NT) { keys[s] = PC_EMPTY_KEY; score[s] = 0.f; bb1[s] = 0.f; bb2[s] = 0.f; ncont[s] = 0u; nhb[s] = 0u; }
    if (tid == 0) *s_status = 0;
    __syncthreads();
    const uint32_t c0 = a.frame_cstart[f], cn = a.frame_ccount[f];
    // four contacts per thread and round: their records, then their eight group words, are loaded back to back (a
    // thread's contacts used to wait for each other's L2 round trips: two dependent loads per contact)
    for (uint32_t cb = tid; cb < cn; cb += 4 * NT) {
        // a full table makes every further key walk 256 slots, and the batch is redone on the global table
        // anyway (PC_ST_TABLE_OVERFLOW): stop at the first overflow instead of finishing the frame
        if (*(volatile unsigned int *)s_status & (unsigned)PC_ST_TABLE_OVERFLOW) break;
        pc_contact rec[4];
        uint32_t g1[4], g2[4];
#pragma unroll
        for (int u = 0; u < 4; u++) rec[u] = a.contacts[c0 + min(cb + (uint32_t)u * NT, cn - 1u)];
#pragma unroll
        for (int u = 0; u < 4; u++) { g1[u] = __ldg(a.lgb1 + rec[u].i); g2[u] = __ldg(a.lgb2 + rec[u].j); }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (cb + (uint32_t)u * NT >= cn) break;
            // group id and backbone flag of an atom in one word: two random loads per contact, not four
            const unsigned long long key = (unsigned long long)(g1[u] & 0x7fffffffu) * a.ngroups2 + (unsigned long long)(g2[u] & 0x7fffffffu);
            unsigned s = pc_hash64(key) & mask;
            int probes = 0;
            for (;;) {
                const unsigned long long prev = atomicCAS(&keys[s], PC_EMPTY_KEY, key);
                if (prev == PC_EMPTY_KEY || prev == key) break;
                s = (s + 1) & mask;
                if (++probes >= 256) { s = 0xffffffffu; break; }
            }
            if (s == 0xffffffffu) { atomicOr(s_status, (unsigned)PC_ST_TABLE_OVERFLOW); continue; }
            atomicAdd(&score[s], rec[u].weight);
            if (g1[u] >> 31) atomicAdd(&bb1[s], rec[u].weight);
            if (g2[u] >> 31) atomicAdd(&bb2[s], rec[u].weight);
            if (atomicAdd(&ncont[s], 1u) == PC_FLOAT_SUM_MAX) atomicOr(s_status, (unsigned)PC_ST_PRECISION);
        }
    }
    __syncthreads();
    const uint32_t h0 = a.frame_hstart[f], hn = a.frame_hcount[f];
    for (uint32_t h = tid; h < hn; h += NT) {
        const pc_hbond rec = a.hbonds[h0 + h];
        const unsigned long long key =
            (unsigned long long)__ldg(a.lgid1 + rec.i) * a.ngroups2 + (unsigned long long)__ldg(a.lgid2 + rec.j);
        unsigned s = pc_hash64(key) & mask;
        for (int probes = 0; probes < 256; probes++) {
            const unsigned long long k = keys[s];
            if (k == key) { atomicAdd(&nhb[s], 1u); break; }
            if (k == PC_EMPTY_KEY) break;
            s = (s + 1) & mask;
        }
    }
    __syncthreads();
    // Rows out.  A warp owns cap / (NT / 32) consecutive slots and walks them 32 at a time: the lanes read consecutive
    // slots (no bank conflicts) and the occupied ones write one contiguous run of rows, as three 16-byte stores each.
    // (A thread per run of eight consecutive slots read shared memory 16-way bank-conflicted and wrote eight scattered
    // 48-byte rows with 8-byte stores: 12 % of the per-frame kernel's stall samples sat on those stores.)
    static_assert(sizeof(pc_row) == 48, "rows are written as three 16-byte pieces");
    constexpr int per_warp = cap / (NT / 32);
    const unsigned ltmask = (1u << lane) - 1u;
    uint32_t n = 0;                                               // (warp-uniform) occupied slots of this warp
    for (int s0 = warp * per_warp; s0 < (warp + 1) * per_warp; s0 += 32)
        n += (uint32_t)__popc(__ballot_sync(PC_FULL_MASK, keys[s0 + lane] != PC_EMPTY_KEY));
    if (lane == 0) wsum[warp] = n;
    __syncthreads();
    if (warp == 0) {
        const uint32_t v = lane < NT / 32 ? wsum[lane] : 0;
        const uint32_t vi = warp_incl_scan(v, lane);
        wsum[lane] = vi - v;
        const uint32_t total = __shfl_sync(PC_FULL_MASK, vi, 31);
        if (lane == 0) {
            long long base = (long long)atomicAdd(&a.counters[PC_CNT_ROWS], (unsigned long long)total);
            if (base + total > a.cap_rows) { atomicOr(s_status, (unsigned)PC_ST_OUT_OVERFLOW); base = -1; }
            *s_base = base;
            a.frame_rstart[f] = (uint32_t)(base < 0 ? 0 : base);
            a.frame_rcount[f] = base < 0 ? 0u : total;
        }
    }
    __syncthreads();
    if (*s_base >= 0) {
        long long o = *s_base + wsum[warp];
        for (int s0 = warp * per_warp; s0 < (warp + 1) * per_warp; s0 += 32) {
            const int s = s0 + lane;
            const unsigned long long k = keys[s];
            const unsigned m = __ballot_sync(PC_FULL_MASK, k != PC_EMPTY_KEY);
            if (k != PC_EMPTY_KEY) {
                uint4 *dst = reinterpret_cast<uint4 *>(a.rows + o + __popc(m & ltmask));
                const unsigned long long sc = (unsigned long long)__double_as_longlong((double)score[s]);
                const unsigned long long b1 = (unsigned long long)__double_as_longlong((double)bb1[s]);
                const unsigned long long b2 = (unsigned long long)__double_as_longlong((double)bb2[s]);
                // pc_row: key, score | bb1, bb2 | ncontacts, nhbonds, first (= key: no first-appearance order here)
                dst[0] = make_uint4((unsigned)k, (unsigned)(k >> 32), (unsigned)sc, (unsigned)(sc >> 32));
                dst[1] = make_uint4((unsigned)b1, (unsigned)(b1 >> 32), (unsigned)b2, (unsigned)(b2 >> 32));
                dst[2] = make_uint4(ncont[s], nhb[s], (unsigned)k, (unsigned)(k >> 32));
            }
            o += __popc(m);
        }
    }
    if (tid == 0 && *s_status) atomicOr(&a.counters[PC_CNT_STATUS], (unsigned long long)*s_status);
    __syncthreads();
}

template <int NT>
__global__ void __launch_bounds__(NT, 1) pc_accumulate_compact_kernel(const PcAccArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint32_t wsum[32];
    __shared__ long long s_base;
    __shared__ unsigned int s_status;
    for (int f = blockIdx.x; f < a.nframes; f += gridDim.x) pc_compact_frame<NT>(a, f, smem, wsum, &s_base, &s_status);
}

// ---------------------------------------------------------------------------------------------------
// Frames with more keys than a shared-memory table holds (1 M-atom systems, residue-pair maps of a
// capsid): one open-addressing table in global memory for the whole batch, keyed by
// frame * n_keys + key.  Inserts are 64-bit CAS + float64 atomics in L2; a key's first insertion
// bumps its frame's row count, so rows can afterwards be emitted grouped by frame.
// ---------------------------------------------------------------------------------------------------
struct __align__(64) PcGlobalEntry {         // 64 bytes = two 32-byte sectors per probe + update
    unsigned long long key, first;
    double score, bb1, bb2;
    unsigned int ncont, nhb;
    unsigned int rowidx;                     // position of this key among its frame's rows (insertion order)
    unsigned int pad[3];
};

struct PcGlobalTable {
    PcGlobalEntry *e;                        // cap entries; key / first initialised to PC_EMPTY_KEY, sums to 0
    unsigned long long cap;                  // power of two
    unsigned long long nkeys;                // G1 * G2
    unsigned int *fresh;                     // slots in order of first insertion (cap_rows entries)
    unsigned long long *nfresh;              // their number
};

__global__ void __launch_bounds__(256) pc_gacc_init_kernel(PcGlobalTable t) {
    for (unsigned long long s = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; s < t.cap;
         s += (unsigned long long)gridDim.x * blockDim.x) {
        PcGlobalEntry v;
        v.key = PC_EMPTY_KEY; v.first = PC_EMPTY_KEY; v.score = 0.0; v.bb1 = 0.0; v.bb2 = 0.0; v.ncont = 0; v.nhb = 0;
        v.rowidx = 0; v.pad[0] = v.pad[1] = v.pad[2] = 0;
        t.e[s] = v;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *t.nfresh = 0ull;
}

__device__ __forceinline__ long long pc_gacc_find(const PcGlobalTable &t, unsigned long long ck, bool insert, bool &fresh) {
    const unsigned long long mask = t.cap - 1ull;
    unsigned long long h = ck;
    h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
    unsigned long long s = h & mask;
    fresh = false;
    // a table sized from a hint can turn out too small: give up after 512 probes (-> PC_ST_TABLE_OVERFLOW,
    // the caller retries with the real count) instead of walking a nearly full table
    for (unsigned long long probes = 0; probes < 512 && probes < t.cap; probes++) {
        if (insert) {
            const unsigned long long prev = atomicCAS(&t.e[s].key, PC_EMPTY_KEY, ck);
            if (prev == PC_EMPTY_KEY) { fresh = true; return (long long)s; }
            if (prev == ck) return (long long)s;
        } else {
            const unsigned long long k = t.e[s].key;
            if (k == ck) return (long long)s;
            if (k == PC_EMPTY_KEY) return -1;
        }
        s = (s + 1) & mask;
    }
    return -1;
}

// group id and backbone flag of every selection atom in one word (pc_set_maps)
__global__ void pc_pack_gid_flag_kernel(const int *gid, const uint8_t *flag, int n, uint32_t *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t)gid[i] | ((flag[i] & PC_ATOM_BACKBONE) ? 0x80000000u : 0u);
}

// one (frame, key) update of the global table: `cnt` contacts with weight sum w (backbone parts b1, b2) and
// smallest order key `first`
__device__ __forceinline__ unsigned long long pc_gacc_home(const PcGlobalTable &t, unsigned long long ck) {
    unsigned long long h = ck;
    h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
    return h & (t.cap - 1ull);
}

// the update behind a probe: s = slot of the key (< 0: table overflow), fresh = this call inserted it
__device__ __forceinline__ void pc_gacc_update(const PcAccArgs &a, const PcGlobalTable &t, int f, long long s, bool fresh,
                                               double w, double b1, double b2, unsigned int cnt, unsigned long long first,
                                               unsigned int &status) {
    if (s < 0) { status |= (unsigned)PC_ST_TABLE_OVERFLOW; return; }
    if (fresh) {
        t.e[s].rowidx = atomicAdd(&a.frame_rcount[f], 1u);
        const unsigned long long q = atomicAdd(t.nfresh, 1ull);
        if ((long long)q < a.cap_rows) t.fresh[q] = (unsigned int)s;
    }
    PcGlobalEntry *e = t.e + s;
    atomicAdd(&e->score, w);
    if (b1 != 0.0) atomicAdd(&e->bb1, b1);
    if (b2 != 0.0) atomicAdd(&e->bb2, b2);
    atomicAdd(&e->ncont, cnt);
    atomicMin(&e->first, first);
}

__device__ __forceinline__ void pc_gacc_apply(const PcAccArgs &a, const PcGlobalTable &t, int f, unsigned long long key,
                                              double w, double b1, double b2, unsigned int cnt, unsigned long long first,
                                              unsigned int &status) {
    bool fresh;
    const long long s = pc_gacc_find(t, (unsigned long long)f * t.nkeys + key, true, fresh);
    pc_gacc_update(a, t, f, s, fresh, w, b1, b2, cnt, first, status);
}

// Contacts reach the table in the order the pair search found them: neighbouring records share their A atoms'
// residues and the B residues next to them, so a tile of PC_GACC_TILE consecutive contacts holds only a few
// hundred distinct keys.  Each CTA therefore folds a tile into a small shared-memory table first and sends ONE
// update per distinct key to the global table.  The tile sums are 2^-40 fixed point, exact and order-independent,
// so score and its backbone parts stay consistent (sc = score - bb is formed by the caller and must not see two
// differently rounded float sums); each is kept as two 32-bit halves (bits 20.. and 0..19 of the 41-bit weight,
// neither half overflows over 4096 weights <= 1) because 32-bit shared-memory atomics are native while 64-bit
// ones are CAS loops -- the L2 atomics, which bound this kernel, drop by the tile's contacts-per-key ratio.  A
// contact that finds no slot within 16 probes goes to the global table directly.
#ifndef PC_GACC_TILE
#define PC_GACC_TILE 2048
#endif
#ifndef PC_GACC_SLOTS
#define PC_GACC_SLOTS 1024
#endif
#define PC_GACC_FIX 1099511627776.0     // 2^40
// dynamic shared memory: key, first (u64 each), 4 classes x (hi, lo), count
#define PC_GACC_SMEM (PC_GACC_SLOTS * (8 + 8 + 4 * 2 * 4 + 4))
// Every contact belongs to exactly ONE of four classes (atom 1 backbone?, atom 2 backbone?), so it costs three
// shared-memory atomics (its class's two halves + the count) instead of up to seven; score = sum of the classes,
// bb1 = classes 2 + 3, bb2 = classes 1 + 3, in exact integer arithmetic at flush time.
__global__ void __launch_bounds__(256) pc_gacc_insert_kernel(const PcAccArgs a, PcGlobalTable t) {
    extern __shared__ __align__(16) unsigned char gacc_smem[];
    unsigned long long *s_key = reinterpret_cast<unsigned long long *>(gacc_smem);
    unsigned long long *s_first = s_key + PC_GACC_SLOTS;
    unsigned int *s_hi = reinterpret_cast<unsigned int *>(s_first + PC_GACC_SLOTS);      // [4][SLOTS]
    unsigned int *s_lo = s_hi + 4 * PC_GACC_SLOTS;                                       // [4][SLOTS]
    unsigned int *s_cnt = s_lo + 4 * PC_GACC_SLOTS;
    const int f = blockIdx.y, tid = threadIdx.x;
    const uint32_t c0 = a.frame_cstart[f], cn = a.frame_ccount[f];
    unsigned int status = 0;
    for (uint32_t base = blockIdx.x * (uint32_t)PC_GACC_TILE; base < cn; base += gridDim.x * (uint32_t)PC_GACC_TILE) {
        for (int s = tid; s < PC_GACC_SLOTS; s += 256) { s_key[s] = PC_EMPTY_KEY; s_first[s] = PC_EMPTY_KEY; s_cnt[s] = 0u; }
        for (int s = tid; s < 8 * PC_GACC_SLOTS; s += 256) s_hi[s] = 0u;                 // s_hi and s_lo are adjacent
        __syncthreads();
        const uint32_t end = min(cn, base + (uint32_t)PC_GACC_TILE);
        for (uint32_t c = base + tid; c < end; c += 256) {
            const pc_contact rec = a.contacts[c0 + c];
            const uint32_t g1 = __ldg(a.lgb1 + rec.i), g2 = __ldg(a.lgb2 + rec.j);
            const unsigned long long key = (unsigned long long)(g1 & 0x7fffffffu) * a.ngroups2 + (unsigned long long)(g2 & 0x7fffffffu);
            const bool bk1 = g1 >> 31, bk2 = g2 >> 31;
            const unsigned long long ok = a.order_keys ? a.order_keys[c0 + c]
                                                       : (((unsigned long long)(unsigned)rec.i << 32) | (unsigned)rec.j);
            unsigned int h = (unsigned int)(pc_hash64(key) & (PC_GACC_SLOTS - 1));
            int slot = -1;
#pragma unroll 1
            for (int p = 0; p < 16; p++) {
                const unsigned long long prev = atomicCAS(&s_key[h], PC_EMPTY_KEY, key);
                if (prev == PC_EMPTY_KEY || prev == key) { slot = (int)h; break; }
                h = (h + 1) & (PC_GACC_SLOTS - 1);
            }
            if (slot >= 0) {
                const unsigned long long wq = __double2ull_rn((double)fminf(rec.weight, 1.0f) * PC_GACC_FIX);
                const int cls = (bk1 ? 2 : 0) + (bk2 ? 1 : 0);
                atomicAdd(&s_hi[cls * PC_GACC_SLOTS + slot], (unsigned int)(wq >> 20));
                atomicAdd(&s_lo[cls * PC_GACC_SLOTS + slot], (unsigned int)(wq & 0xfffffu));
                atomicAdd(&s_cnt[slot], 1u);
                if (ok < s_first[slot]) atomicMin(&s_first[slot], ok);     // a CAS loop: only when it can lower the minimum
            } else {
                pc_gacc_apply(a, t, f, key, (double)rec.weight, bk1 ? (double)rec.weight : 0.0,
                              bk2 ? (double)rec.weight : 0.0, 1u, ok, status);
            }
        }
        __syncthreads();
        // flush: the global probe is a CAS that returns (an L2 round trip), and a thread owns SLOTS / 256 slots --
        // the first probes of all of them are issued before any is consumed (this loop was 48 % of the kernel's
        // stall samples when each slot waited for the previous one)
        {
            constexpr int PER = PC_GACC_SLOTS / 256;
            unsigned long long ck[PER], home[PER], prev[PER];
#pragma unroll
            for (int u = 0; u < PER; u++) {
                const int s = tid + u * 256;
                ck[u] = s_key[s];
                if (ck[u] != PC_EMPTY_KEY) {
                    ck[u] += (unsigned long long)f * t.nkeys;
                    home[u] = pc_gacc_home(t, ck[u]);
                    prev[u] = atomicCAS(&t.e[home[u]].key, PC_EMPTY_KEY, ck[u]);
                } else {
                    ck[u] = PC_EMPTY_KEY; home[u] = 0; prev[u] = 0;
                }
            }
#pragma unroll
            for (int u = 0; u < PER; u++) {
                const int s = tid + u * 256;
                if (ck[u] == PC_EMPTY_KEY) continue;
                long long gs = (long long)home[u];
                bool fresh = prev[u] == PC_EMPTY_KEY;
                if (!fresh && prev[u] != ck[u]) {                    // collision at home: the ordinary probe sequence from the next slot
                    gs = -1;
                    unsigned long long p = (home[u] + 1) & (t.cap - 1ull);
                    for (unsigned long long probes = 1; probes < 512 && probes < t.cap; probes++) {
                        const unsigned long long pv = atomicCAS(&t.e[p].key, PC_EMPTY_KEY, ck[u]);
                        if (pv == PC_EMPTY_KEY) { fresh = true; gs = (long long)p; break; }
                        if (pv == ck[u]) { gs = (long long)p; break; }
                        p = (p + 1) & (t.cap - 1ull);
                    }
                }
                unsigned long long v[4];
#pragma unroll
                for (int k = 0; k < 4; k++)
                    v[k] = ((unsigned long long)s_hi[k * PC_GACC_SLOTS + s] << 20) + (unsigned long long)s_lo[k * PC_GACC_SLOTS + s];
                pc_gacc_update(a, t, f, gs, fresh, (double)(v[0] + v[1] + v[2] + v[3]) * (1.0 / PC_GACC_FIX),
                               (double)(v[2] + v[3]) * (1.0 / PC_GACC_FIX), (double)(v[1] + v[3]) * (1.0 / PC_GACC_FIX), s_cnt[s],
                               s_first[s], status);
            }
        }
        __syncthreads();
    }
    status = __reduce_or_sync(PC_FULL_MASK, status);
    if ((threadIdx.x & 31) == 0 && status) atomicOr(&a.counters[PC_CNT_STATUS], (unsigned long long)status);
}

__global__ void __launch_bounds__(256) pc_gacc_hbond_kernel(const PcAccArgs a, PcGlobalTable t) {
    const int f = blockIdx.y;
    const uint32_t h0 = a.frame_hstart[f], hn = a.frame_hcount[f];
    for (uint32_t h = blockIdx.x * blockDim.x + threadIdx.x; h < hn; h += gridDim.x * blockDim.x) {
        const pc_hbond rec = a.hbonds[h0 + h];
        const unsigned long long key =
            (unsigned long long)__ldg(a.lgid1 + rec.i) * a.ngroups2 + (unsigned long long)__ldg(a.lgid2 + rec.j);
        bool fresh;
        const long long s = pc_gacc_find(t, (unsigned long long)f * t.nkeys + key, false, fresh);
        if (s >= 0) atomicAdd(&t.e[s].nhb, 1u);
    }
}

// one CTA: frame_rstart = exclusive scan of frame_rcount, total -> counters, cursors zeroed
__global__ void __launch_bounds__(1024) pc_gacc_offsets_kernel(const PcAccArgs a, uint32_t *cursor) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < a.nframes; base += 1024) {
        const int f = base + tid;
        const uint32_t v = f < a.nframes ? a.frame_rcount[f] : 0;
        const uint32_t incl = warp_incl_scan(v, lane);
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) { const uint32_t w = wsum[lane]; const uint32_t wi = warp_incl_scan(w, lane); wsum[lane] = wi - w; }
        __syncthreads();
        if (f < a.nframes) { a.frame_rstart[f] = carry + wsum[warp] + incl - v; cursor[f] = 0; }
        __syncthreads();
        if (tid == 1023) carry += wsum[warp] + incl;
        __syncthreads();
    }
    if (tid == 0) {
        a.counters[PC_CNT_ROWS] = carry;
        if ((long long)carry > a.cap_rows) atomicOr(&a.counters[PC_CNT_STATUS], PC_ST_OUT_OVERFLOW);
    }
}

__global__ void __launch_bounds__(256) pc_gacc_emit_kernel(const PcAccArgs a, PcGlobalTable t, uint32_t *cursor) {
    const unsigned long long n = min(*t.nfresh, (unsigned long long)a.cap_rows);
    for (unsigned long long q = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; q < n;
         q += (unsigned long long)gridDim.x * blockDim.x) {
        const PcGlobalEntry e = t.e[t.fresh[q]];
        const unsigned long long ck = e.key;
        const unsigned long long f = ck / t.nkeys;
        const long long pos = (long long)a.frame_rstart[f] + e.rowidx;
        if (pos >= a.cap_rows) continue;
        pc_row r;
        r.key = ck - f * t.nkeys; r.score = e.score; r.bb1 = e.bb1; r.bb2 = e.bb2;
        r.ncontacts = e.ncont; r.nhbonds = e.nhb; r.first = e.first;
        a.rows[pos] = r;
    }
}

// pc_row -> pc_row_packed for the D2H transfer of throughput runs
__global__ void __launch_bounds__(256) pc_pack_rows_kernel(const pc_row *rows, const unsigned long long *counters,
                                                           long long cap_rows, pc_row_packed *out) {
    const long long n = min((long long)counters[PC_CNT_ROWS], cap_rows);
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
        const pc_row v = rows[r];
        pc_row_packed p;
        p.key = v.key; p.score = (float)v.score; p.bb1 = (float)v.bb1; p.bb2 = (float)v.bb2;
        p.ncontacts = (uint16_t)min(v.ncontacts, 65535u); p.nhbonds = (uint16_t)min(v.nhbonds, 65535u);
        out[r] = p;
    }
}

__global__ void __launch_bounds__(256) pc_pack_rows_slim_kernel(const pc_row *rows, const unsigned long long *counters,
                                                                long long cap_rows, pc_row_slim *out) {
    const long long n = min((long long)counters[PC_CNT_ROWS], cap_rows);
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
        const pc_row v = rows[r];
        pc_row_slim p;
        p.key = (uint32_t)v.key; p.score = (float)v.score;
        p.ncontacts = (uint16_t)min(v.ncontacts, 65535u); p.nhbonds = (uint16_t)min(v.nhbonds, 65535u);
        out[r] = p;
    }
}

// ---- sparse time-aggregated maps (key spaces too large for pc_fold_rows' dense table) ----------------------
__global__ void __launch_bounds__(256) pc_fold_sparse_clear_kernel(pc_fold_entry *tab, unsigned long long cap, unsigned long long *meta) {
    for (unsigned long long s = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; s < cap;
         s += (unsigned long long)gridDim.x * blockDim.x) {
        pc_fold_entry e;
        e.key = PC_EMPTY_KEY; e.score = 0.0; e.bb1 = 0.0; e.bb2 = 0.0; e.hbframes = 0.0;
        tab[s] = e;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { meta[0] = 0; meta[1] = 0; }      // [0] overflow flag, [1] compaction cursor
}

__global__ void __launch_bounds__(256) pc_fold_sparse_kernel(const pc_row *rows, const unsigned long long *counters,
                                                             long long cap_rows, pc_fold_entry *tab, unsigned long long cap,
                                                             unsigned long long *meta) {
    if (counters[PC_CNT_STATUS] & PC_ST_RERUN_MASK) return;
    const long long n = min((long long)counters[PC_CNT_ROWS], cap_rows);
    const unsigned long long mask = cap - 1ull;
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
        const pc_row row = rows[r];
        unsigned long long h = row.key;
        h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
        unsigned long long s = h & mask;
        bool found = false;
        for (int probes = 0; probes < 512; probes++) {
            const unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long *>(&tab[s].key), PC_EMPTY_KEY,
                                                       (unsigned long long)row.key);
            if (prev == PC_EMPTY_KEY || prev == row.key) { found = true; break; }
            s = (s + 1) & mask;
        }
        if (!found) { meta[0] = 1ull; continue; }
        atomicAdd(&tab[s].score, row.score);
        if (row.bb1 != 0.0) atomicAdd(&tab[s].bb1, row.bb1);
        if (row.bb2 != 0.0) atomicAdd(&tab[s].bb2, row.bb2);
        if (row.nhbonds) atomicAdd(&tab[s].hbframes, 1.0);
    }
}

__global__ void __launch_bounds__(256) pc_fold_sparse_compact_kernel(const pc_fold_entry *tab, unsigned long long cap,
                                                                     pc_fold_entry *out, unsigned long long max_out,
                                                                     unsigned long long *meta) {
    for (unsigned long long s = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; s < cap;
         s += (unsigned long long)gridDim.x * blockDim.x) {
        const pc_fold_entry e = tab[s];
        if (e.key == PC_EMPTY_KEY) continue;
        const unsigned long long p = atomicAdd(&meta[1], 1ull);
        if (p < max_out) out[p] = e;
    }
}

// ---- exchange of sparse aggregated maps between ranks (frame-sharded runs) ---------------------
// export: the table's occupied entries, compacted into a FIXED-capacity buffer (rest padded with empty keys) so that
// every rank contributes the same number of bytes to one all-gather; import: another rank's buffer folded into
// this table.  meta[0] = overflow flag, meta[1] = compaction cursor (cleared by the caller).
__global__ void __launch_bounds__(256) pc_fold_sparse_export_kernel(const pc_fold_entry *tab, unsigned long long cap,
                                                                    pc_fold_entry *out, unsigned long long max_out,
                                                                    unsigned long long *meta) {
    for (unsigned long long s = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; s < cap;
         s += (unsigned long long)gridDim.x * blockDim.x) {
        const pc_fold_entry e = tab[s];
        if (e.key == PC_EMPTY_KEY) continue;
        const unsigned long long p = atomicAdd(&meta[1], 1ull);
        if (p < max_out) out[p] = e;
        else meta[0] = 1ull;
    }
}

__global__ void __launch_bounds__(256) pc_fold_sparse_pad_kernel(pc_fold_entry *out, unsigned long long max_out,
                                                                 const unsigned long long *meta) {
    pc_fold_entry e;
    e.key = PC_EMPTY_KEY; e.score = 0.0; e.bb1 = 0.0; e.bb2 = 0.0; e.hbframes = 0.0;
    for (unsigned long long s = min(meta[1], max_out) + blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; s < max_out;
         s += (unsigned long long)gridDim.x * blockDim.x)
        out[s] = e;
}

__global__ void __launch_bounds__(256) pc_fold_sparse_import_kernel(const pc_fold_entry *in, unsigned long long n,
                                                                    pc_fold_entry *tab, unsigned long long cap,
                                                                    unsigned long long *meta) {
    const unsigned long long mask = cap - 1ull;
    for (unsigned long long r = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; r < n;
         r += (unsigned long long)gridDim.x * blockDim.x) {
        const pc_fold_entry row = in[r];
        if (row.key == PC_EMPTY_KEY) continue;
        unsigned long long h = row.key;
        h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
        unsigned long long s = h & mask;
        bool found = false;
        for (int probes = 0; probes < 512; probes++) {
            const unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long *>(&tab[s].key), PC_EMPTY_KEY,
                                                       (unsigned long long)row.key);
            if (prev == PC_EMPTY_KEY || prev == row.key) { found = true; break; }
            s = (s + 1) & mask;
        }
        if (!found) { meta[0] = 1ull; continue; }
        atomicAdd(&tab[s].score, row.score);
        if (row.bb1 != 0.0) atomicAdd(&tab[s].bb1, row.bb1);
        if (row.bb2 != 0.0) atomicAdd(&tab[s].bb2, row.bb2);
        if (row.hbframes != 0.0) atomicAdd(&tab[s].hbframes, row.hbframes);
    }
}
