// pc_scan_small.cuh -- "small frame" path: one CTA analyses one whole frame out of shared memory.
//
// Replaces, for one frame, the body of loop_trajectory_grid (core/multi_trajectory.py:65-219):
//   find_contacts / vmd_gridsearch3 (cy_modules/src/gridsearch.C:408-427, 902-1162)
//   + heavy-atom filter (:89) + self-interaction filter (:93-95) + distance/weight (:98-107)
//   + H-bond test (:112-209), and optionally the per-frame accumulation of
//   analyze_contactResultsWithMaps (core/ContactAnalyzer.py:527-559).
//
// Phases (persistent CTAs, frames strided over the grid; CTA barriers between phases):
//   1  first pass over the frame: bounding boxes of the heavy atoms of both selections
//   2  grid over (bboxA (+) c) n (bboxB (+) c): only atoms inside it can have a partner
//   3  second pass over the frame (an L2 hit): atoms inside the grid -> candidate pool in shared memory
//      (x, y, z, packed atom word) + cell id, atoms per cell (shared atomics)
//   4  exclusive scans
//   5  scatter the candidates into cell order
//   6  pair search: one thread per (A atom, z-layer of its stencil); the three B x-runs of the layer
//      are contiguous ranges of the sorted B array; fused pre-test + exact float32 test
//      (pc_common.cuh: pair_within); hits -> shared queue
//   7  drain: one thread per hit -> float64 distance, weight, H-bond test on N/O/S pairs, contact
//      record; with fuse_acc the hit is also added to the frame's (key -> score) table in shared memory
//   8  H-bond records and accumulated rows -> global
// Only in-grid atoms are ever staged, so a frame needs little shared memory: the "fast" plan runs two
// 512-thread CTAs per SM (one frame's barriers and loads hide behind the other's work); frames whose
// in-grid atoms or hits do not fit it are queued on the device and redone by the "full" plan (one
// 1024-thread CTA per SM with the largest staging) in a second launch.
//
// Hydrogens never enter the search (the reference discards every pair with a hydrogen at :89); they
// are read from global memory only inside the H-bond test.
#pragma once
#include "pc_common.cuh"

// One pass over the frame instead of two when a CTA's previous frame predicts this one: the first pass (bounding
// boxes) also keeps every heavy atom inside the previous frame's grid box grown by PC_SMALL_PRED_MARGIN on each side.
// When this frame's exact box turns out to lie inside that box, the in-grid atoms are among the kept ones and are
// filtered out of shared memory (phase 3'); otherwise the second pass over the frame runs as before.  The pair set does
// not depend on which of the two happened: atoms outside the exact box cannot have a partner.
#ifndef PC_SMALL_PRED
#define PC_SMALL_PRED 1
#endif
#ifndef PC_SMALL_PRED_MARGIN
#define PC_SMALL_PRED_MARGIN 1.0f
#endif

struct PcSmallSmem {
    // offsets (bytes) into dynamic shared memory, computed on the host
    int off_cand, off_ccell, off_hits, off_hb, off_tab, off_sort, off_cellA, off_cellB, total;
    int capCand, capTab;
};

// Shared-memory plan.  One region is time-shared inside a frame: the candidate pool (atoms inside the
// grid, unsorted, + their cell ids) lives in phases 3-5; hit queue, H-bond staging and the
// accumulation table live in phases 6-8 and are laid over it.  Both pools hold selection A from the
// bottom up and selection B from the top down, so only the SUM of in-grid atoms is bounded.
__host__ inline PcSmallSmem pc_small_layout(int self_mode, int capCells, int capHits, int capHb, int capCand, int capTab) {
    PcSmallSmem L;
    int o = 0;
    auto take = [&](int bytes) { int r = o; o += (bytes + 15) & ~15; return r; };
    L.off_cand = take(16 * capCand);
    L.off_ccell = take(2 * capCand);
    const int end_early = o;
    o = 0;
    L.off_hits = take(4 * capHits);
    L.off_hb = take((int)sizeof(pc_hbond) * capHb);
    L.off_tab = take(40 * capTab);
    o = o > end_early ? o : end_early;
    L.off_sort = take(16 * capCand);
    L.off_cellA = take(4 * (capCells + 1));
    L.off_cellB = self_mode ? L.off_cellA : take(4 * (capCells + 1));
    L.capCand = capCand; L.capTab = capTab;
    L.total = o;
    return L;
}

// Cells are numbered with the axis that has the fewest cells running fastest ("f"), the one with the most
// slowest ("s"): contact interfaces are slabs, and when the thin axis has <= 3 cells an atom's three-cell
// run is the whole row, so the three rows of a stencil layer are one contiguous range of the sorted array.
struct PcGridSm {
    float lo[3];
    float inv;
    int nx, ny, nz, ncell;
    int sx, sy, sz;       // cell id = cx * sx + cy * sy + cz * sz
    int pf, pm, ps;       // axis (0 = x, 1 = y, 2 = z) that runs fastest / middle / slowest
    int nf, nm, ns;       // cells along those axes
};

#define PC_EMPTY_KEY 0xffffffffffffffffull

__device__ __forceinline__ unsigned pc_hash64(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (unsigned)k;
}

// Slot of a key in the per-frame shared-memory table: multiplicative hash, top bits (the table is private to
// one frame of one CTA, so this hash need not match any other table's).
__device__ __forceinline__ unsigned pc_hash_tab(unsigned long long k, int shift) {
    const unsigned h = (unsigned)k * 0x9E3779B1u + (unsigned)(k >> 32) * 0x85EBCA6Bu;
    return (h * 0x2545F491u) >> shift;
}

template <int NT>
__device__ __forceinline__ void block_excl_scan_inplace(uint32_t *arr, int n, uint32_t *wsum) {
    // arr[0..n) -> exclusive prefix sums, arr[n] = total.  Each thread owns a contiguous chunk.
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ch = (n + NT - 1) / NT;
    const int b = min(tid * ch, n), e = min(b + ch, n);
    uint32_t s = 0;
    for (int i = b; i < e; i++) s += arr[i];
    const uint32_t incl = warp_incl_scan(s, lane);
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const uint32_t v = lane < NT / 32 ? wsum[lane] : 0;
        const uint32_t vi = warp_incl_scan(v, lane);
        wsum[lane] = vi - v;
    }
    __syncthreads();
    uint32_t off = wsum[warp] + incl - s;
    for (int i = b; i < e; i++) { const uint32_t t = arr[i]; arr[i] = off; off += t; }
    if (tid == NT - 1) arr[n] = off;
    __syncthreads();
}

// One launch covers either all frames of the batch (frame_list == nullptr) or the frames another
// launch could not stage (frame_list / frame_count on the device).  With retry_list set, a frame whose
// in-grid atoms or hits overflow this launch's staging is appended there instead of being reported.
struct PcSmallFrames {
    const int *frame_list;
    const unsigned int *frame_count;
    int *retry_list;
    unsigned int *retry_count;
};

template <int NT>
__global__ void __launch_bounds__(NT, NT >= 1024 ? 1 : NT >= 512 ? 2 : NT >= 384 ? 3 : 4) pc_scan_small_kernel(const PcScanArgs a, const PcSmallSmem L,
                                                                               const PcSmallFrames fl) {
    extern __shared__ __align__(16) unsigned char smem[];
    float4 *cand = reinterpret_cast<float4 *>(smem + L.off_cand);
    uint16_t *ccell = reinterpret_cast<uint16_t *>(smem + L.off_ccell);
    float4 *sortp = reinterpret_cast<float4 *>(smem + L.off_sort);
    uint32_t *cellA = reinterpret_cast<uint32_t *>(smem + L.off_cellA);
    uint32_t *cellB = reinterpret_cast<uint32_t *>(smem + L.off_cellB);
    uint32_t *hits = reinterpret_cast<uint32_t *>(smem + L.off_hits);
    pc_hbond *hbq = reinterpret_cast<pc_hbond *>(smem + L.off_hb);
    // accumulation table (fuse_acc): structure of arrays over capTab slots
    const int capTab = L.capTab, capCand = L.capCand;
    // sums are float32: shared-memory float atomics are native (float64 ones are CAS loops), and a sum of a
    // few dozen float32 weights stays within ~1e-7 relative of the float64 sum (the parity bar is 1e-5)
    unsigned long long *tkey = reinterpret_cast<unsigned long long *>(smem + L.off_tab);
    unsigned long long *tfirst = tkey + capTab;
    // every contact belongs to ONE class (atom 1 backbone ? 2 : 0) + (atom 2 backbone ? 1 : 0) and adds its weight to
    // that class only: one float atomic per hit instead of one to three (score, bb1, bb2); the row's score is the
    // sum of the four classes, bb1 = classes 2 + 3, bb2 = classes 1 + 3
    float *tcls = reinterpret_cast<float *>(tfirst + capTab);            // [4][capTab]
    unsigned int *tncont = reinterpret_cast<unsigned int *>(tcls + 4 * capTab);
    unsigned int *tnhb = tncont + capTab;

    __shared__ float red[NT / 32][12];
    __shared__ uint32_t wsum[32];
    __shared__ PcGridSm g;
    __shared__ unsigned int s_nhits, s_nhb, s_nA, s_nB, s_ngate, s_next;
    __shared__ long long s_cbase, s_hbase, s_rbase;
    __shared__ unsigned int s_status;
    __shared__ float s_pbox[6];           // predicted box (lo xyz, hi xyz) from this CTA's previous frame
    __shared__ int s_pred, s_usepred;     // s_pred: s_pbox is valid; s_usepred: this frame's exact box lies inside it
    if (threadIdx.x == 0) { s_pred = 0; s_usepred = 0; s_nA = 0; s_nB = 0; }
    __syncthreads();

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool self = a.self_mode != 0;
    const bool fuse = a.fuse_acc != 0;
    const int n1h = a.t.n1h, n2h = self ? a.t.n1h : a.t.n2h;
    const float sqdist = a.sqdist;
    unsigned long long evals = 0;

    const int nfr = fl.frame_list ? (int)*fl.frame_count : a.nframes;
    for (int fi = blockIdx.x; fi < nfr; fi += gridDim.x) {
        const int f = fl.frame_list ? fl.frame_list[fi] : fi;
        const float *fr1 = a.xyz1 + (long long)f * a.pitch1;
        const float *fr2 = self ? fr1 : a.xyz2 + (long long)f * a.pitch2;
        PC_PHASE_BEGIN();

        // ---- phase 1: bounding boxes of the heavy atoms (first pass over the frame) -----------------
        float mn[6], mx[6];
#pragma unroll
        for (int k = 0; k < 6; k++) { mn[k] = INFINITY; mx[k] = -INFINITY; }
        // four atoms per iteration: the index loads, then the coordinate loads, are issued back to back
        const bool pred = PC_SMALL_PRED && s_pred != 0;
        auto bbox_pass = [&](const float *fr, const uint32_t *hinfo, int nh, float *lo3, float *hi3, unsigned int *counter,
                             const bool top) {
            for (int k0 = tid; k0 < nh; k0 += 4 * NT) {
                uint32_t info[4];
                float x[4], y[4], z[4];
#pragma unroll
                for (int u = 0; u < 4; u++) info[u] = __ldg(hinfo + min(k0 + u * NT, nh - 1));
#pragma unroll
                for (int u = 0; u < 4; u++) load_xyz(fr, PC_W_INDEX(info[u]), a.stride, x[u], y[u], z[u]);
#pragma unroll
                for (int u = 0; u < 4; u++) {       // a clamped duplicate of the last atom leaves min/max unchanged
                    lo3[0] = fminf(lo3[0], x[u]); lo3[1] = fminf(lo3[1], y[u]); lo3[2] = fminf(lo3[2], z[u]);
                    hi3[0] = fmaxf(hi3[0], x[u]); hi3[1] = fmaxf(hi3[1], y[u]); hi3[2] = fmaxf(hi3[2], z[u]);
                    if (pred && k0 + u * NT < nh && x[u] >= s_pbox[0] && x[u] <= s_pbox[3] && y[u] >= s_pbox[1] &&
                        y[u] <= s_pbox[4] && z[u] >= s_pbox[2] && z[u] <= s_pbox[5]) {
                        const unsigned p = atomicAdd(counter, 1u);
                        if (p < (unsigned)capCand)
                            cand[top ? (unsigned)capCand - 1u - p : p] = make_float4(x[u], y[u], z[u], __uint_as_float(info[u]));
                    }
                }
            }
        };
        bbox_pass(fr1, a.t.hinfo1, n1h, mn, mx, &s_nA, false);
        if (!self) bbox_pass(fr2, a.t.hinfo2, n2h, mn + 3, mx + 3, &s_nB, true);
#pragma unroll
        for (int k = 0; k < 6; k++) { mn[k] = warp_min(mn[k]); mx[k] = warp_max(mx[k]); }
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 6; k++) { red[warp][k] = mn[k]; red[warp][6 + k] = mx[k]; }
        }
        if (tid == 0) { s_nhits = 0; s_nhb = 0; s_status = 0; s_ngate = 0; s_next = 0; }     // (s_nA, s_nB: reset at the frame's end)
        __syncthreads();
        PC_PHASE_MARK(0);

        // ---- phase 2: grid geometry (warp 0 reduces the per-warp boxes) ---------------------------
        if (warp == 0) {
            float v = (lane < 12) ? red[0][lane] : 0.f;
            if (lane < 12)
                for (int w = 1; w < NT / 32; w++) v = lane < 6 ? fminf(v, red[w][lane]) : fmaxf(v, red[w][lane]);
            // lanes 0..5: min of A xyz, B xyz; lanes 6..11: max of A xyz, B xyz
            const float minA = __shfl_sync(PC_FULL_MASK, v, lane % 3), minB = __shfl_sync(PC_FULL_MASK, v, 3 + lane % 3);
            const float maxA = __shfl_sync(PC_FULL_MASK, v, 6 + lane % 3), maxB = __shfl_sync(PC_FULL_MASK, v, 9 + lane % 3);
            // lanes 0..2 hold axis k = lane
            float l, h;
            if (self) { l = minA - a.edge; h = maxA + a.edge; }
            else { l = fmaxf(minA, minB) - a.edge; h = fminf(maxA, maxB) + a.edge; }
            bool empty_k = !(l <= h), bad_k = !isfinite(l) || !isfinite(h);
            const unsigned em = __ballot_sync(PC_FULL_MASK, lane < 3 && empty_k);
            const unsigned bm = __ballot_sync(PC_FULL_MASK, lane < 3 && bad_k);
            const float ext = h - l;
            const float ex = __shfl_sync(PC_FULL_MASK, ext, 0), ey = __shfl_sync(PC_FULL_MASK, ext, 1),
                        ez = __shfl_sync(PC_FULL_MASK, ext, 2);
            const float lx = __shfl_sync(PC_FULL_MASK, l, 0), ly = __shfl_sync(PC_FULL_MASK, l, 1),
                        lz = __shfl_sync(PC_FULL_MASK, l, 2);
            {
                // does the exact box lie inside the box the first pass kept atoms for?  Then: the next frame's prediction
                const bool in_k = l >= s_pbox[lane % 3] && h <= s_pbox[3 + lane % 3];
                const unsigned outm = __ballot_sync(PC_FULL_MASK, lane < 3 && !in_k);
                const bool okbox = em == 0 && bm == 0 && n1h != 0 && n2h != 0;
                if (lane < 3) { s_pbox[lane] = l - PC_SMALL_PRED_MARGIN; s_pbox[3 + lane] = h + PC_SMALL_PRED_MARGIN; }
                if (lane == 0) { s_usepred = (pred && okbox && outm == 0) ? 1 : 0; s_pred = okbox ? 1 : 0; }
            }
            if (lane == 0) {
                bool empty = em != 0 || n1h == 0 || n2h == 0;
                if (bm && !empty) { atomicOr(&s_status, (unsigned)PC_ST_BAD_COORDS); empty = true; }
                int nx = 0, ny = 0, nz = 0;
                float inv = 0.f;
                if (!empty) pc_grid_dims(ex, ey, ez, a.edge, (long long)a.capCells, nx, ny, nz, inv);
                g.lo[0] = lx; g.lo[1] = ly; g.lo[2] = lz;
                g.inv = inv; g.nx = nx; g.ny = ny; g.nz = nz; g.ncell = nx * ny * nz;
                // axes by cell count, ascending (ties keep x < y < z)
                int pf = 0, pm = 1, ps = 2;
                int nn[3] = {nx, ny, nz};
                if (nn[pm] < nn[pf]) { const int t = pf; pf = pm; pm = t; }
                if (nn[ps] < nn[pm]) { const int t = pm; pm = ps; ps = t; }
                if (nn[pm] < nn[pf]) { const int t = pf; pf = pm; pm = t; }
                g.pf = pf; g.pm = pm; g.ps = ps;
                g.nf = nn[pf]; g.nm = nn[pm]; g.ns = nn[ps];
                int st[3];
                st[pf] = 1; st[pm] = nn[pf]; st[ps] = nn[pf] * nn[pm];
                g.sx = st[0]; g.sy = st[1]; g.sz = st[2];
            }
        }
        __syncthreads();
        PC_PHASE_MARK(1);
        const int ncell = g.ncell;
        int nInA = 0, nInB = 0;
        bool staged = true;               // false: the frame does not fit this launch's staging
        float4 *sortA = sortp, *sortB = sortp;
        if (ncell > 0) {
            const int nx = g.nx, ny = g.ny, nz = g.nz;
            const float lox = g.lo[0], loy = g.lo[1], loz = g.lo[2], inv = g.inv;
            const int sx = g.sx, sy = g.sy, sz = g.sz;

            // ---- phase 3: second pass over the frame (it is in L2 now): atoms inside the grid go to
            //      the candidate pool with their cell id; atoms per cell (shared atomics) -----------------
            // (with the prediction's atoms in the pool: phase 3' below instead)
            const int stA = (int)s_nA, stB = self ? 0 : (int)s_nB;            // kept by the first pass
            const bool filt = s_usepred != 0 && stA + stB <= capCand;
            __syncthreads();                                                  // everybody has read the counters
            for (int c = tid; c <= ncell; c += NT) { cellA[c] = 0; if (!self) cellB[c] = 0; }
            if (!filt && tid == 0) { s_nA = 0; s_nB = 0; }
            __syncthreads();
            auto cell_of = [&](float x, float y, float z) -> int {
                // atoms outside the grid box are farther than one cell edge (> cutoff) from the other
                // selection's bounding box
                const float ux = (x - lox) * inv, uy = (y - loy) * inv, uz = (z - loz) * inv;
                if (!(ux >= 0.f && uy >= 0.f && uz >= 0.f)) return -1;
                const int cx = (int)ux, cy = (int)uy, cz = (int)uz;
                if (cx >= nx || cy >= ny || cz >= nz) return -1;
                return cx * sx + cy * sy + cz * sz;
            };
            // selection A fills the pool from the bottom, B from the top; four atoms per iteration so that
            // the loads of an iteration are in flight together
            auto stage_pass = [&](const float *fr, const uint32_t *hinfo, int nh, unsigned int *counter, uint32_t *cells,
                                  const bool top) {
                for (int k0 = tid; k0 < nh; k0 += 4 * NT) {
                    uint32_t info[4];
                    float x[4], y[4], z[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) info[u] = __ldg(hinfo + min(k0 + u * NT, nh - 1));
#pragma unroll
                    for (int u = 0; u < 4; u++) load_xyz(fr, PC_W_INDEX(info[u]), a.stride, x[u], y[u], z[u]);
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int c = k0 + u * NT < nh ? cell_of(x[u], y[u], z[u]) : -1;
                        if (c >= 0) {
                            // (the compiler already merges the same-address atomics of a warp into one)
                            const unsigned p = atomicAdd(counter, 1u);
                            if (p < (unsigned)capCand) {
                                const unsigned q = top ? (unsigned)capCand - 1u - p : p;
                                cand[q] = make_float4(x[u], y[u], z[u], __uint_as_float(info[u]));
                                ccell[q] = (uint16_t)c;
                            }
                            atomicAdd(&cells[c], 1u);
                        }
                    }
                }
            };
            if (filt) {
                // ---- phase 3': the kept atoms' exact cells; atoms outside the exact grid are marked and left behind
                for (int p = tid; p < stA; p += NT) {
                    const float4 v = cand[p];
                    const int c = cell_of(v.x, v.y, v.z);
                    ccell[p] = (uint16_t)(c < 0 ? 0xffff : c);
                    if (c >= 0) atomicAdd(&cellA[c], 1u);
                }
                for (int p = tid; p < stB; p += NT) {
                    const float4 v = cand[capCand - 1 - p];
                    const int c = cell_of(v.x, v.y, v.z);
                    ccell[capCand - 1 - p] = (uint16_t)(c < 0 ? 0xffff : c);
                    if (c >= 0) atomicAdd(&cellB[c], 1u);
                }
            } else {
                stage_pass(fr1, a.t.hinfo1, n1h, &s_nA, cellA, false);
                if (!self) stage_pass(fr2, a.t.hinfo2, n2h, &s_nB, cellB, true);
            }
            __syncthreads();
            PC_PHASE_MARK(2);
            nInA = (int)s_nA; nInB = self ? nInA : (int)s_nB;      // atoms in the pool
#ifdef PC_PHASE_CLOCKS
            if (tid == 0) {
                atomicAdd(&a.counters[PC_CNT_PHASE + 16], (unsigned long long)(nInA + (self ? 0 : nInB)));
                atomicMax(&a.counters[PC_CNT_PHASE + 17], (unsigned long long)(nInA + (self ? 0 : nInB)));
                atomicAdd(&a.counters[PC_CNT_PHASE + 18], 1ull);
            }
#endif
            if (nInA + (self ? 0 : nInB) > capCand) {
                staged = false;
            } else {
                // ---- phase 4: scans ---------------------------------------------------------------------
                block_excl_scan_inplace<NT>(cellA, ncell, wsum);
                if (!self) block_excl_scan_inplace<NT>(cellB, ncell, wsum);
                PC_PHASE_MARK(3);
                // ---- phase 5: scatter into cell order; afterwards cell[c] == END of cell c -----------
                const int poolA = nInA, poolB = nInB;
                nInA = (int)cellA[ncell]; nInB = self ? nInA : (int)cellB[ncell];     // atoms inside the grid (all of the pool
                sortA = sortp;                                                        // unless phase 3' left some behind)
                sortB = self ? sortp : sortp + (capCand - nInB);
                for (int p = tid; p < poolA; p += NT) {
                    const unsigned c = ccell[p];
                    if (c != 0xffffu) sortA[atomicAdd(&cellA[c], 1u)] = cand[p];
                }
                if (!self)
                    for (int p = tid; p < poolB; p += NT) {
                        const unsigned c = ccell[capCand - 1 - p];
                        if (c != 0xffffu) sortB[atomicAdd(&cellB[c], 1u)] = cand[capCand - 1 - p];
                    }
                __syncthreads();
                PC_PHASE_MARK(4);

                // ---- phase 6: pair search, one thread per (A atom, stencil layer) --------------------
                // items are layer-major: the lanes of a warp hold neighbouring atoms (same or adjacent
                // cells, A is cell-sorted) on the same layer, so their run lengths are alike
                const int pf = g.pf, pm = g.pm, ps = g.ps, nf = g.nf, nm = g.nm, ns = g.ns;
                // batches of 32 items are handed out dynamically: the warps finish the phase together
                const int nitems = 3 * nInA;
                for (;;) {
                    int w = 0;
                    if (lane == 0) w = (int)atomicAdd(&s_next, 32u);
                    w = __shfl_sync(PC_FULL_MASK, w, 0);
                    if (w >= nitems) break;
                    w += lane;
                    if (w >= nitems) continue;
                    const int layer = (w >= nInA) + (w >= 2 * nInA), ia = w - layer * nInA;
                    const float4 pa = sortA[ia];
                    const int cx = (int)((pa.x - lox) * inv), cy = (int)((pa.y - loy) * inv),
                              cz = (int)((pa.z - loz) * inv);
                    const int cf = pf == 0 ? cx : pf == 1 ? cy : cz, cm = pm == 0 ? cx : pm == 1 ? cy : cz,
                              cs = ps == 0 ? cx : ps == 1 ? cy : cz;
                    const int sl = cs + layer - 1;
                    if (sl < 0 || sl >= ns) continue;
                    const int flo = max(cf - 1, 0), fhi = min(cf + 1, nf - 1);
                    const float2 nxy = make_float2(-pa.x, -pa.y);
                    const float naz = -pa.z;
                    int ares = 0, aseg = 0;
                    if (self) {
                        const int li = PC_W_INDEX(__float_as_uint(pa.w));
                        ares = __ldg(a.t.lres1 + li); aseg = __ldg(a.t.lseg1 + li);
                    }
                    // The (up to) three rows of the layer -> up to three runs of the sorted B array; rows that
                    // follow each other in the array are merged (always, when the fast axis has <= 3 cells).
                    // Built before the candidate loop so that lanes with one merged run stay converged.
                    int ab = 0, ae = 0, bb = 0, be = 0, cb = 0, ce = 0;
                    if (flo == 0 && fhi == nf - 1) {
                        // whole rows: the cells of rows cm-1 .. cm+1 are one contiguous range (two lookups)
                        const int first = (sl * nm + max(cm - 1, 0)) * nf, last = (sl * nm + min(cm + 1, nm - 1)) * nf + nf - 1;
                        ab = first ? (int)cellB[first - 1] : 0; ae = (int)cellB[last];
                    } else {
                        const int row1 = (sl * nm + cm) * nf;
                        if (cm > 0) {
                            const int row0 = row1 - nf;
                            ab = (row0 + flo) ? (int)cellB[row0 + flo - 1] : 0; ae = (int)cellB[row0 + fhi];
                        }
                        const int b0 = (row1 + flo) ? (int)cellB[row1 + flo - 1] : 0, b1 = (int)cellB[row1 + fhi];
                        if (ab == ae) { ab = b0; ae = b1; }
                        else if (b0 == ae) ae = b1;
                        else if (b0 != b1) { bb = b0; be = b1; }
                        if (cm + 1 < nm) {
                            const int row2 = row1 + nf;
                            const int c0 = (int)cellB[row2 + flo - 1], c1 = (int)cellB[row2 + fhi];
                            if (c0 != c1) {
                                if (bb != be) { if (c0 == be) be = c1; else { cb = c0; ce = c1; } }
                                else if (ab == ae) { ab = c0; ae = c1; }
                                else if (c0 == ae) ae = c1;
                                else { bb = c0; be = c1; }
                            }
                        }
                    }
                    evals += (unsigned)((ae - ab) + (be - bb) + (ce - cb));
                    for (int ri = 0; ri < 3; ri++) {
                        const int rb0 = ri == 0 ? ab : ri == 1 ? bb : cb, rb1 = ri == 0 ? ae : ri == 1 ? be : ce;
                        if (rb0 == rb1) break;
                        // 32 candidates at a time into a hit mask; the queue is touched once per chunk
                        for (int jb = rb0; jb < rb1; jb += 32) {
                            const int cnt = min(32, rb1 - jb);
                            unsigned m = 0;
                            int k0 = 0;
                            // no further unrolling: a 16-wide and a 4-wide body would be two paths that the
                            // lanes of a warp (different run lengths) execute one after the other
#pragma unroll 1
                            for (; k0 + 4 <= cnt; k0 += 4) {
                                unsigned m4 = 0;
#pragma unroll
                                for (int u = 0; u < 4; u++)
                                    if (pair_within_packed(nxy, naz, sortB[jb + k0 + u], sqdist)) m4 |= 1u << u;
                                m |= m4 << k0;
                            }
#pragma unroll 1
                            for (; k0 < cnt; k0++)
                                if (pair_within_packed(nxy, naz, sortB[jb + k0], sqdist)) m |= 1u << k0;
                            if (self && m)
                                m = pc_self_filter(m, ares, aseg, a.t.lres1, a.t.lseg1,
                                                   [&](int k) { return __float_as_uint(sortB[jb + k].w); });
                            if (!m) continue;
                            unsigned p = atomicAdd(&s_nhits, (unsigned)__popc(m));
                            while (m) {
                                const int j = jb + __ffs(m) - 1;
                                m &= m - 1;
                                if (p < (unsigned)a.capHits) hits[p] = ((unsigned)ia << 16) | (unsigned)j;
                                p++;
                            }
                        }
                    }
                }
                PC_PHASE_MARK(5);       // thread 0's own share of the search
            }
        }
        // The accumulation table is laid over the candidate pool, which nobody reads any more: every
        // warp that gets here is past the barrier that ended phase 5 (or skipped the grid).
        if (fuse)
            for (int s = tid; s < capTab; s += NT) {
                tkey[s] = PC_EMPTY_KEY; tcls[s] = 0.f; tcls[capTab + s] = 0.f; tcls[2 * capTab + s] = 0.f; tcls[3 * capTab + s] = 0.f;
                tfirst[s] = PC_EMPTY_KEY; tncont[s] = 0; tnhb[s] = 0;
            }
        __syncthreads();
        PC_PHASE_MARK(6);

        // A frame that does not fit this launch's staging (in-grid atoms, hits) is handed to the launch
        // with the larger plan when there is one; nothing of it has been written to global memory yet.
        unsigned nh = s_nhits;
        if (!staged || nh > (unsigned)a.capHits) {
            if (fl.retry_list != nullptr) {
#ifdef PC_PHASE_CLOCKS
                if (tid == 0) atomicAdd(&a.counters[PC_CNT_PHASE + (staged ? 13 : 12)], 1ull);   // why frames are handed on
#endif
                if (tid == 0) fl.retry_list[atomicAdd(fl.retry_count, 1u)] = f;
                __syncthreads();
                continue;
            }
            if (tid == 0) atomicOr(&s_status, (unsigned)(staged ? PC_ST_HITS_OVERFLOW : PC_ST_STAGE_OVERFLOW));
            nh = staged ? (unsigned)a.capHits : 0u;
        }

        // ---- phase 7: drain -----------------------------------------------------------------------
        if (tid == 0) {
            long long base = (long long)atomicAdd(&a.counters[PC_CNT_CONTACTS], (unsigned long long)nh);
            if (base + nh > a.cap_contacts) { atomicOr(&s_status, (unsigned)PC_ST_OUT_OVERFLOW); base = -1; }
            s_cbase = base;
        }
        __syncthreads();
        PC_PHASE_MARK(7);
        const long long cbase = s_cbase;
        const unsigned tmask = (unsigned)capTab - 1u;
        const int tshift = 33 - __ffs(capTab);        // 32 - log2(capTab); capTab is a power of two
        // One donor-H...acceptor side of a gated pair (core/multi_trajectory.py:112-209).
        // side 0: donor = atom1, acceptor = atom2; side 1: donor = atom2, acceptor = atom1
        auto hbond_side = [&](const unsigned u, const int side) {
            const float4 pa = sortA[u >> 16], pb = sortB[u & 0xffffu];
            const int li = PC_W_INDEX(__float_as_uint(pa.w)), lj = PC_W_INDEX(__float_as_uint(pb.w));
            const int *hptr = side == 0 ? a.t.lhptr1 : a.t.lhptr2;
            const int *hidx = side == 0 ? a.t.lhidx1 : a.t.lhidx2;
            const int heavy = side == 0 ? li : lj;
            const int p0 = __ldg(hptr + heavy), p1 = __ldg(hptr + heavy + 1);
            if (p0 == p1) return;
            const double ax = pa.x, ay = pa.y, az = pa.z, bx = pb.x, by = pb.y, bz = pb.z;
            for (int p = p0; p < p1; p++) {
                const int hl = __ldg(hidx + p);
                if (hl < 0) { atomicOr(&s_status, (unsigned)PC_ST_MISSING_H); continue; }
                float hxf, hyf, hzf;
                load_xyz(side == 0 ? fr1 : fr2, hl, a.stride, hxf, hyf, hzf);
                double d = 0, ang = 0;
                const bool ok = side == 0
                    ? hbond_test(hxf, hyf, hzf, ax, ay, az, bx, by, bz, a.hb_cutoff, a.hb_angle, d, ang)
                    : hbond_test(hxf, hyf, hzf, bx, by, bz, ax, ay, az, a.hb_cutoff, a.hb_angle, d, ang);
                if (!ok) continue;
                const unsigned q = atomicAdd(&s_nhb, 1u);
                if (q < (unsigned)a.capHb) {
                    pc_hbond r;
                    r.i = li; r.j = lj; r.hyd = (uint32_t)hl | (side ? 0x80000000u : 0u);
                    r.distance = (float)d; r.angle = (float)ang;
                    hbq[q] = r;
                }
                if (fuse) {
                    // the pair's key was inserted by 7a (unless the table overflowed, which 7a reports)
                    const unsigned long long key =
                        (unsigned long long)__ldg(a.t.lgid1 + li) * a.t.ngroups2 + (unsigned long long)__ldg(a.t.lgid2 + lj);
                    unsigned s = pc_hash_tab(key, tshift);
                    for (int probes = 0; probes < capTab; probes++) {
                        const unsigned long long k = *(volatile unsigned long long *)&tkey[s];
                        if (k == key) { atomicAdd(&tnhb[s], 1u); break; }
                        if (k == PC_EMPTY_KEY) break;
                        s = (s + 1) & tmask;
                    }
                }
            }
        };

        // 7a: one thread per hit -> contact record and accumulation.  Neighbouring hits come from the same
        // A atom and B run, i.e. mostly from the same (group, group) key: the lanes of a warp take hits
        // NT/32 apart so that their shared-memory atomics land on different table slots.  Pairs that pass
        // the H-bond gate (both names start with O/N/S, core/multi_trajectory.py:112) are queued for 7b in
        // the memory of the cell arrays, which the search no longer needs.
        uint32_t *gateq = cellA;
        const unsigned capGate = (unsigned)((self ? 1 : 2) * (a.capCells + 1));
        constexpr int NW = NT / 32;
        for (unsigned base = 0; base < nh; base += NT) {
            const unsigned h = base + (unsigned)(lane * NW + ((warp + lane) % NW));
            if (h >= nh) continue;
            const unsigned u = hits[h];
            const float4 pa = sortA[u >> 16], pb = sortB[u & 0xffffu];
            const unsigned wa = __float_as_uint(pa.w), wb = __float_as_uint(pb.w);
            const int li = PC_W_INDEX(wa), lj = PC_W_INDEX(wb);
            const double dx = (double)pa.x - (double)pb.x, dy = (double)pa.y - (double)pb.y,
                         dz = (double)pa.z - (double)pb.z;
            const double dist = sqrt(dx * dx + dy * dy + dz * dz);
            const float wgt = contact_weight_f32(dist);
            if (cbase >= 0) {
                pc_contact rec;
                rec.i = li; rec.j = lj; rec.distance = (float)dist; rec.weight = wgt;
                // the queue has no order of its own: in a full round the records are written in thread order
                // (one 512-byte run per warp) instead of the scattered order in which the hits are taken
                a.contacts[cbase + (base + NT <= nh ? base + (unsigned)tid : h)] = rec;
            }
            if (fuse) {
                const unsigned long long key =
                    (unsigned long long)__ldg(a.t.lgid1 + li) * a.t.ngroups2 + (unsigned long long)__ldg(a.t.lgid2 + lj);
                unsigned slot = 0xffffffffu, s = pc_hash_tab(key, tshift);
                for (int probes = 0; probes < capTab; probes++) {
                    const unsigned long long prev = atomicCAS(&tkey[s], PC_EMPTY_KEY, key);
                    if (prev == PC_EMPTY_KEY || prev == key) { slot = s; break; }
                    s = (s + 1) & tmask;
                }
                if (slot == 0xffffffffu) atomicOr(&s_status, (unsigned)PC_ST_TABLE_OVERFLOW);
                else {
                    const int cls = ((wa & PC_W_BACKBONE) ? 2 : 0) + ((wb & PC_W_BACKBONE) ? 1 : 0);
                    atomicAdd(&tcls[cls * capTab + slot], wgt);
                    // float32 class sums: a cell with more contacts than PC_FLOAT_SUM_MAX is redone exactly (global table)
                    if (atomicAdd(&tncont[slot], 1u) == PC_FLOAT_SUM_MAX) atomicOr(&s_status, (unsigned)PC_ST_PRECISION);
                    // the minimum only ever decreases: a (possibly stale) read that is already smaller settles it
                    const unsigned long long mine = ((unsigned long long)(unsigned)li << 32) | (unsigned)lj;
                    if (*(volatile unsigned long long *)&tfirst[slot] > mine) atomicMin(&tfirst[slot], mine);
                }
            }
            if (wa & wb & PC_W_HBCLASS) {
                const unsigned q = atomicAdd(&s_ngate, 1u);
                if (q < capGate) gateq[q] = u;
                else { hbond_side(u, 0); hbond_side(u, 1); }      // queue full: tested in place
            }
        }
        PC_PHASE_MARK(8);           // thread 0's own share of the drain
        __syncthreads();
        PC_PHASE_MARK(9);

        // 7b: one thread per (gated pair, side)
        const unsigned ngate = min(s_ngate, capGate);
        for (unsigned it = tid; it < 2u * ngate; it += NT) hbond_side(gateq[it >> 1], (int)(it & 1u));
        __syncthreads();
        PC_PHASE_MARK(11);

        // ---- phase 8: H-bond records and accumulated rows -> global -----------------------------
        unsigned nb = s_nhb;
        if (nb > (unsigned)a.capHb) { nb = a.capHb; if (tid == 0) atomicOr(&s_status, (unsigned)PC_ST_HB_OVERFLOW); }
        uint32_t cnt = 0, incl = 0;
        int tb = 0, te = 0;
        if (fuse) {
            const int ch = (capTab + NT - 1) / NT;
            tb = min(tid * ch, capTab); te = min(tb + ch, capTab);
            for (int s = tb; s < te; s++) cnt += tkey[s] != PC_EMPTY_KEY;
            incl = warp_incl_scan(cnt, lane);
            if (lane == 31) wsum[warp] = incl;
        }
        __syncthreads();
        if (warp == 0) {
            uint32_t total = 0;
            if (fuse) {
                const uint32_t v = lane < NT / 32 ? wsum[lane] : 0;
                const uint32_t vi = warp_incl_scan(v, lane);
                wsum[lane] = vi - v;
                total = __shfl_sync(PC_FULL_MASK, vi, 31);
            }
            if (lane == 0) {
                // both claims are issued before either result is looked at (one round trip to L2, not two)
                long long base = (long long)atomicAdd(&a.counters[PC_CNT_HBONDS], (unsigned long long)nb);
                long long rb = fuse ? (long long)atomicAdd(&a.counters[PC_CNT_ROWS], (unsigned long long)total) : 0;
                if (base + nb > a.cap_hbonds) { atomicOr(&s_status, (unsigned)PC_ST_OUT_OVERFLOW); base = -1; }
                s_hbase = base;
                a.frame_cstart[f] = (uint32_t)(cbase < 0 ? 0 : cbase);
                a.frame_ccount[f] = cbase < 0 ? 0u : nh;
                a.frame_hstart[f] = (uint32_t)(base < 0 ? 0 : base);
                a.frame_hcount[f] = base < 0 ? 0u : nb;
                if (fuse) {
                    if (rb + total > a.cap_rows) { atomicOr(&s_status, (unsigned)PC_ST_OUT_OVERFLOW); rb = -1; }
                    s_rbase = rb;
                    a.frame_rstart[f] = (uint32_t)(rb < 0 ? 0 : rb);
                    a.frame_rcount[f] = rb < 0 ? 0u : total;
                }
            }
        }
        __syncthreads();
        if (s_hbase >= 0)
            for (unsigned q = tid; q < nb; q += NT) a.hbonds[s_hbase + q] = hbq[q];
        if (fuse && s_rbase >= 0) {
            long long o = s_rbase + wsum[warp] + incl - cnt;
            for (int s = tb; s < te; s++) {
                if (tkey[s] == PC_EMPTY_KEY) continue;
                pc_row r;
                const double c0 = tcls[s], c1 = tcls[capTab + s], c2 = tcls[2 * capTab + s], c3 = tcls[3 * capTab + s];
                r.key = tkey[s]; r.score = (c0 + c1) + (c2 + c3); r.bb1 = c2 + c3; r.bb2 = c1 + c3;
                r.ncontacts = tncont[s]; r.nhbonds = tnhb[s]; r.first = tfirst[s];
                a.rows[o++] = r;
            }
        }
        if (tid == 0 && s_status) atomicOr(&a.counters[PC_CNT_STATUS], (unsigned long long)s_status);
        if (tid == 0) { s_nA = 0; s_nB = 0; }          // the next frame's first pass appends kept atoms right away
        __syncthreads();
        PC_PHASE_MARK(10);
    }
    // evaluation counter -> global (one atomic per warp)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) evals += __shfl_xor_sync(PC_FULL_MASK, evals, o);
    if (lane == 0 && evals) atomicAdd(&a.counters[PC_CNT_EVALS], evals);
}
