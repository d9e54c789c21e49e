// pc_scan_group.cuh -- large-frame path, dense frames: pair search + records + H-bonds + per-frame accumulation
// in ONE kernel whose CTAs own whole accumulation groups of selection A.
//
// A 3 M-atom capsid against itself has 1.1e7 contacts and 1.1e6 (frame, key) cells per
// frame.  The multi-kernel sequence (pc_scan_large.cuh) writes the hits, re-reads them for the records, re-reads
// the records for the accumulation and pushes every one of them through a multi-GB hash table in global memory
// (pc_gacc_insert_kernel: random DRAM probes, ~29x the row output in traffic).  The reference's accumulation
// (core/ContactAnalyzer.py:527-559) groups a frame's contacts by key = (group of atom 1, group of atom 2); the
// groups of the usual maps (residue, residue name, segment) are runs of consecutive atoms of a selection.  So
// a CTA that searches a run of consecutive A atoms made of WHOLE groups sees every contact of its keys, and a
// small table in shared memory is all the accumulation needs:
//   G1  a thread per heavy A atom of the unit (read straight from the frame through the heavy-atom list, like the
//       tail kernel's AFRAME mode); the unit's atoms are rank-sorted by cell inside the CTA (the table does not care
//       in which order they are searched) and stay in shared memory with their packed group words
//   G2  pair search of gridsearch.C:1111-1142, a warp per occupied cell of the unit: the B atoms of the cell's 27-cell
//       neighbourhood are nine contiguous runs of the cell-sorted array; the warp walks them as ONE list, 32 at a time
//       (one coalesced load each), against the cell's A atoms (broadcast from shared memory) with the reference's exact
//       float32 expression (pair_within_packed) and the self-interaction filter (the B atoms' residue / segment come
//       cell-sorted from the scatter kernel); ballot -> per-warp queues in shared memory.  Out of line, no calls inside
//   G3  whenever a queue might not take the next round, its warp drains it: float64 distance, weight, one coalesced run
//       of 16-byte contact records into the frame's slice, the donor-H...acceptor tests of N/O/S pairs
//       (core/multi_trajectory.py:112-209), and the hit's weight into the unit's (key -> class sums) table
//   G4  the table's rows go to the frame's slice of a row buffer; pc_group_offsets_kernel / pc_group_gather_kernel
//       then lay the frames' rows out back to back like every other accumulation path does
// Units are cut on the host (pc_set_maps): consecutive groups, at most PC_GROUP_NT heavy atoms.  Maps whose groups
// are not runs of the selection's order (e.g. atom name only), groups larger than a unit, order keys, or a table /
// float32-sum overflow (PC_ST_TABLE_OVERFLOW / PC_ST_PRECISION -> table_slots 16384) use the global-table path.
#pragma once
#include "pc_scan_large.cuh"
#include "pc_accumulate.cuh"

#define PC_GROUP_NT 256
#define PC_GROUP_TAB 1024                 // table slots per unit (power of two)
#define PC_GROUP_WQ 384                   // hits per warp queue (>= 160; its memory first stages the unit's atoms: 32 B each)
#ifndef PC_GROUP_MINB
#define PC_GROUP_MINB 4                   // CTAs per SM the kernel is compiled for
#endif
#define PC_GROUP_SMEM (PC_GROUP_NT * 36 + 16 + PC_GROUP_TAB * 32 + (PC_GROUP_NT / 32) * PC_GROUP_WQ * 4)   // == PC_GOFF_WQ + queues

// Dynamic shared memory of the group kernel (byte offsets; the out-of-line device functions address it directly, so
// that their loads and atomics are shared-memory instructions rather than generic ones)
#define PC_GOFF_SA     0                                               // float4 [NT]  -x, -y, -z, atom word (cell order)
#define PC_GOFF_TKEY   (PC_GOFF_SA + PC_GROUP_NT * 16)                 // u64 [TAB]
#define PC_GOFF_TCLS   (PC_GOFF_TKEY + PC_GROUP_TAB * 8)               // float [4][TAB] class sums (pc_scan_small.cuh)
#define PC_GOFF_TNCONT (PC_GOFF_TCLS + PC_GROUP_TAB * 16)              // u32 [TAB]
#define PC_GOFF_TNHB   (PC_GOFF_TNCONT + PC_GROUP_TAB * 4)             // u32 [TAB]
#define PC_GOFF_SRS    (PC_GOFF_TNHB + PC_GROUP_TAB * 4)               // int2 [NT] residue, segment (self filter)
#define PC_GOFF_SG     (PC_GOFF_SRS + PC_GROUP_NT * 8)                 // u32 [NT] group id | backbone << 31
#define PC_GOFF_SKEY   (PC_GOFF_SG + PC_GROUP_NT * 4)                  // u32 [NT] cell << 8 | thread, sorted
#define PC_GOFF_CLIST  (PC_GOFF_SKEY + PC_GROUP_NT * 4)                // u32 [NT + 4] first atom of each occupied cell
#define PC_GOFF_WQ     (PC_GOFF_CLIST + (PC_GROUP_NT + 4) * 4)         // u32 [NW][WQ] hit queues; before the search: staging

struct PcGroupArgs {
    const uint32_t *unit_start;           // nunits + 1 heavy ordinals of selection A
    int nunits;
    pc_row *slices;                       // nframes x cap_r_frame
    long long cap_r_frame;
};

struct PcGroupCtx {                       // per-CTA state of the out-of-line record / H-bond code, in shared memory
    const float *fr1, *fr2;
    const int *hptr1, *hidx1, *hptr2, *hidx2;
    double hb_cutoff, hb_angle;
    pc_hbond *hslice;
    long long cap_h;
    unsigned int *hcursor;
    unsigned int status;
    // records + accumulation
    const float4 *sB;                     // the frame's cell-sorted B atoms (global)
    const uint32_t *lgb2;                 // group id | backbone << 31 per local index of selection 2
    pc_contact *slice;
    long long cap_c;
    unsigned long long ngroups2;
    unsigned int *ccursor;
    int self;
};

// One donor-H...acceptor side of an N/O/S pair; returns the number of hydrogen bonds found.
__device__ __noinline__ unsigned pc_group_hbond(PcGroupCtx *c, const float4 pa, const float4 pb, const int side) {
    const int li = PC_W_INDEX(__float_as_uint(pa.w)), lj = PC_W_INDEX(__float_as_uint(pb.w));
    const int *hptr = side == 0 ? c->hptr1 : c->hptr2;
    const int *hidx = side == 0 ? c->hidx1 : c->hidx2;
    const float *frd = side == 0 ? c->fr1 : c->fr2;
    const int heavy = side == 0 ? li : lj;
    const int p0 = __ldg(hptr + heavy), p1 = __ldg(hptr + heavy + 1);
    unsigned found = 0;
    const double ax = pa.x, ay = pa.y, az = pa.z, bx = pb.x, by = pb.y, bz = pb.z;
    for (int p = p0; p < p1; p++) {
        const int hl = __ldg(hidx + p);
        if (hl < 0) { atomicOr(&c->status, (unsigned)PC_ST_MISSING_H); continue; }
        float hx, hy, hz;
        load_xyz(frd, hl, 3, hx, hy, hz);
        double d = 0, ang = 0;
        const bool ok = side == 0 ? hbond_test(hx, hy, hz, ax, ay, az, bx, by, bz, c->hb_cutoff, c->hb_angle, d, ang)
                                  : hbond_test(hx, hy, hz, bx, by, bz, ax, ay, az, c->hb_cutoff, c->hb_angle, d, ang);
        if (!ok) continue;
        found++;
        const unsigned hq = atomicAdd(c->hcursor, 1u);
        if ((long long)hq < c->cap_h) {
            pc_hbond r;
            r.i = li; r.j = lj; r.hyd = (uint32_t)hl | (side ? 0x80000000u : 0u);
            r.distance = (float)d; r.angle = (float)ang;
            c->hslice[hq] = r;
        } else {
            atomicOr(&c->status, (unsigned)PC_ST_OUT_OVERFLOW);
        }
    }
    return found;
}

// One hit (A atom of the unit << 24 | B position) -> contact record `slot` of the frame's slice, the unit's table, and the
// H-bond tests of an N/O/S pair.  Out of line: the search loop keeps the kernel's registers to itself.
__device__ __noinline__ void pc_group_emit(PcGroupCtx *c, const unsigned u, const long long slot) {
    constexpr int TAB = PC_GROUP_TAB;
    extern __shared__ __align__(16) unsigned char pc_gsm[];
    unsigned long long *tkey = reinterpret_cast<unsigned long long *>(pc_gsm + PC_GOFF_TKEY);
    float *tcls = reinterpret_cast<float *>(pc_gsm + PC_GOFF_TCLS);
    unsigned int *tncont = reinterpret_cast<unsigned int *>(pc_gsm + PC_GOFF_TNCONT);
    unsigned int *tnhb = reinterpret_cast<unsigned int *>(pc_gsm + PC_GOFF_TNHB);
    float4 qa = reinterpret_cast<const float4 *>(pc_gsm + PC_GOFF_SA)[u >> 24];
    qa.x = -qa.x; qa.y = -qa.y; qa.z = -qa.z;                          // (the search keeps the unit's atoms negated)
    const float4 qb = __ldg(c->sB + (u & 0xffffffu));
    const unsigned wa = __float_as_uint(qa.w), wb = __float_as_uint(qb.w);
    const int lj = PC_W_INDEX(wb);
    const uint32_t g1 = reinterpret_cast<const uint32_t *>(pc_gsm + PC_GOFF_SG)[u >> 24], g2 = __ldg(c->lgb2 + lj);
    const double dx = (double)qa.x - (double)qb.x, dy = (double)qa.y - (double)qb.y, dz = (double)qa.z - (double)qb.z;
    const double dist = sqrt(dx * dx + dy * dy + dz * dz);
    const float wgt = contact_weight_f32(dist);
    if (slot < c->cap_c) {
        pc_contact rec;
        rec.i = PC_W_INDEX(wa); rec.j = lj; rec.distance = (float)dist; rec.weight = wgt;
        c->slice[slot] = rec;
    } else {
        atomicOr(&c->status, (unsigned)PC_ST_OUT_OVERFLOW);
    }
    const unsigned long long key = (unsigned long long)(g1 & 0x7fffffffu) * c->ngroups2 + (unsigned long long)(g2 & 0x7fffffffu);
    unsigned s = pc_hash_tab(key, 33 - __ffs(TAB));
    int probes = 0;
    for (;;) {
        const unsigned long long prev = atomicCAS(&tkey[s], PC_EMPTY_KEY, key);
        if (prev == PC_EMPTY_KEY || prev == key) break;
        s = (s + 1) & (unsigned)(TAB - 1);
        if (++probes >= TAB) { s = 0xffffffffu; break; }
    }
    if (s == 0xffffffffu) {
        atomicOr(&c->status, (unsigned)PC_ST_TABLE_OVERFLOW);
    } else {
        atomicAdd(&tcls[((g1 >> 31) * 2 + (g2 >> 31)) * TAB + s], wgt);
        if (atomicAdd(&tncont[s], 1u) == PC_FLOAT_SUM_MAX) atomicOr(&c->status, (unsigned)PC_ST_PRECISION);
    }
    if (wa & wb & PC_W_HBCLASS) {
        const unsigned nh = pc_group_hbond(c, qa, qb, 0) + pc_group_hbond(c, qa, qb, 1);
        if (nh && s != 0xffffffffu) atomicAdd(&tnhb[s], nh);
    }
}

// A warp's queued hits -> records + table; called by whole warps: lanes over hits, ONE claim of the frame's cursor per
// call.  Out of line: its state does not compete with the search loop's for registers.
__device__ __noinline__ void pc_group_drain(PcGroupCtx *c, const unsigned n) {
    extern __shared__ __align__(16) unsigned char pc_gsm[];
    const int lane = threadIdx.x & 31;
    const uint32_t *mywq = reinterpret_cast<const uint32_t *>(pc_gsm + PC_GOFF_WQ) + (threadIdx.x >> 5) * PC_GROUP_WQ;
    __syncwarp();
    if (n) {
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(c->ccursor, n);
        base = __shfl_sync(PC_FULL_MASK, base, 0);
        for (unsigned k = lane; k < n; k += 32) pc_group_emit(c, mywq[k], (long long)base + k);
    }
    __syncwarp();
}

// Per-CTA constants and per-warp resume state of the search (shared memory).
struct PcGroupSearch {
    const float4 *sB;                     // the frame's cell-sorted B atoms
    const int2 *sBrs;                     // their (residue, segment), self contacts only
    const uint32_t *cB;                   // cB[c] == END of cell c in sB
    int nx, ny, nz, ncl;
    float inv_nx, inv_ny, sqdist;
    int next;                             // next cell of the unit's cell list nobody has taken yet
    int ci[PC_GROUP_NT / 32], t0[PC_GROUP_NT / 32], k0[PC_GROUP_NT / 32];     // where each warp resumes (ci < 0: take a cell)
    unsigned long long evals[PC_GROUP_NT / 32];
};

// n / d and n % d for 0 <= n < 2^22 through a float32 reciprocal (exact after one correction step)
__device__ __forceinline__ void pc_divmod_small(const int n, const int d, const float inv_d, int &q, int &r) {
    q = (int)((float)n * inv_d);
    r = n - q * d;
    if (r < 0) { q--; r += d; } else if (r >= d) { q++; r -= d; }
}

// G2: a warp per occupied cell of the unit (cells are handed out one at a time: the warps of a CTA finish together).  The B
// atoms of the cell's 27-cell neighbourhood are nine contiguous runs of the cell-sorted array (one per stencil row); the
// warp walks them as ONE list, 32 atoms at a time (lane -> list position -> run by a four-step search over the runs' ends,
// held in lanes 0..8), one coalesced load each, and tests them against the cell's A atoms (broadcast from shared memory,
// at most eight per round) with the reference's exact float32 expression (gridsearch.C:48-55, :1126); ballot -> the warp's queue.
// Returns the queue's fill when the queue might not take the next round (the caller drains it and calls again: the
// position is kept in S) or when no cell is left (S->ci[warp] >= S->ncl).  Out of line and free of calls: the loop
// nest has the register file to itself.
template <bool SELF>
__device__ __noinline__ unsigned pc_group_search(PcGroupSearch *S) {
    constexpr int WQ = PC_GROUP_WQ;
    extern __shared__ __align__(16) unsigned char pc_gsm[];
    const float4 *sA = reinterpret_cast<const float4 *>(pc_gsm + PC_GOFF_SA);
    const uint32_t *skey = reinterpret_cast<const uint32_t *>(pc_gsm + PC_GOFF_SKEY);
    const uint32_t *clist = reinterpret_cast<const uint32_t *>(pc_gsm + PC_GOFF_CLIST);
    const int2 *sRS = reinterpret_cast<const int2 *>(pc_gsm + PC_GOFF_SRS);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *mywq = reinterpret_cast<uint32_t *>(pc_gsm + PC_GOFF_WQ) + warp * WQ;
    const unsigned ltmask = (1u << lane) - 1u;
    const float4 *sB = S->sB;
    const uint32_t *cB = S->cB;
    const int nx = S->nx, ny = S->ny, nz = S->nz, ncl = S->ncl;
    const float sqdist = S->sqdist;
    int ci = S->ci[warp], t0 = S->t0[warp], k0 = S->k0[warp];
    unsigned long long evals = 0;
    unsigned wn = 0;
    auto take = [&]() {
        int v = 0;
        if (lane == 0) v = atomicAdd(&S->next, 1);
        return __shfl_sync(PC_FULL_MASK, v, 0);
    };
    if (ci < 0) ci = take();
    while (ci < ncl) {
        const int a0 = (int)clist[ci], a1 = (int)clist[ci + 1];
        int cx, cy, cz, rowzy;
        pc_divmod_small((int)(skey[a0] >> 8), nx, S->inv_nx, rowzy, cx);
        pc_divmod_small(rowzy, ny, S->inv_ny, cz, cy);
        int b0 = 0, len = 0;
        if (lane < 9) {
            const int z = cz + lane / 3 - 1, y = cy + lane % 3 - 1;
            if (z >= 0 && z < nz && y >= 0 && y < ny) {
                const int row = (z * ny + y) * nx, xlo = max(cx - 1, 0), xhi = min(cx + 1, nx - 1);
                b0 = (row + xlo) ? (int)__ldg(cB + row + xlo - 1) : 0;
                len = (int)__ldg(cB + row + xhi) - b0;
            }
        }
        const int end = (int)warp_incl_scan((uint32_t)len, lane);      // lanes 0..8: list position behind run `lane`; 9..31: nb
        const int nb = __shfl_sync(PC_FULL_MASK, end, 31);
        const int delta = b0 - (end - len);                            // list position t of run `lane` -> B position t + delta
        for (; t0 < nb; t0 += 32, k0 = 0) {
            const int t = t0 + lane;
            // first run whose end lies behind t (ends are non-decreasing; lanes >= 8 hold nb > t for every valid t)
            int r = 0;
#pragma unroll
            for (int step = 8; step > 0; step >>= 1) {
                const int e = __shfl_sync(PC_FULL_MASK, end, r + step - 1);
                if (e <= t) r += step;
            }
            const int j = t + __shfl_sync(PC_FULL_MASK, delta, min(r, 8));
            // lanes past the end test a point at infinity: never within the cutoff
            const float4 pb = t < nb ? __ldg(sB + j) : make_float4(INFINITY, 0.f, 0.f, 0.f);
            int2 brs = make_int2(0, 0);
            if (SELF && t < nb) brs = __ldg(S->sBrs + j);
            // the cell's A atoms, at most eight per round: a round's pushes (<= 32 per atom) must fit the queue
            for (int ks = a0 + k0; ks < a1; ks += 8) {
                const int ke = min(ks + 8, a1);
                if (wn + 32u * (unsigned)(ke - ks) > (unsigned)WQ) {   // (warp-uniform) hand back: the caller drains and calls again
                    if (lane == 0) { S->ci[warp] = ci; S->t0[warp] = t0; S->k0[warp] = ks - a0; S->evals[warp] += evals; }
                    return wn;
                }
#pragma unroll 1
                for (int k = ks; k < ke; k++) {
                    const float4 na = sA[k];
                    bool hit = pair_within_packed(make_float2(na.x, na.y), na.z, pb, sqdist);
                    if (SELF) {
                        // self-interaction filter (core/multi_trajectory.py:93-95): dropped when (resid1 - resid2) < 5,
                        // signed, in the same segment
                        const int2 ars = sRS[k];
                        hit = hit && !((ars.x - brs.x) < 5 && ars.y == brs.y);
                    }
                    const unsigned m = __ballot_sync(PC_FULL_MASK, hit);
                    if (hit) mywq[wn + __popc(m & ltmask)] = ((unsigned)k << 24) | (unsigned)j;
                    wn += (unsigned)__popc(m);
                }
            }
        }
        if (lane == 0) evals += (unsigned long long)(a1 - a0) * (unsigned)nb;      // (counted once, when the cell is done)
        ci = take(); t0 = 0; k0 = 0;
    }
    if (lane == 0) { S->ci[warp] = ci; S->evals[warp] += evals; }
    return wn;
}

__global__ void __launch_bounds__(PC_GROUP_NT, PC_GROUP_MINB) pc_large_group_kernel(const PcScanArgs a, PcLargeFrame *fr, const int f0,
                                                                                    const uint32_t *cellsB, const int capCells,
                                                                                    const float4 *sortB, const int2 *sortRS,
                                                                                    const int nBh_alloc, const PcLargeOut o,
                                                                                    const PcAccArgs acc,
                                                                                    const PcGroupArgs g) {
    constexpr int NT = PC_GROUP_NT, NW = NT / 32, TAB = PC_GROUP_TAB, WQ = PC_GROUP_WQ;
    extern __shared__ __align__(16) unsigned char smem[];
    float4 *sA = reinterpret_cast<float4 *>(smem + PC_GOFF_SA);
    unsigned long long *tkey = reinterpret_cast<unsigned long long *>(smem + PC_GOFF_TKEY);
    float *tcls = reinterpret_cast<float *>(smem + PC_GOFF_TCLS);
    unsigned int *tncont = reinterpret_cast<unsigned int *>(smem + PC_GOFF_TNCONT);
    unsigned int *tnhb = reinterpret_cast<unsigned int *>(smem + PC_GOFF_TNHB);
    int2 *sRS = reinterpret_cast<int2 *>(smem + PC_GOFF_SRS);
    uint32_t *sG = reinterpret_cast<uint32_t *>(smem + PC_GOFF_SG);
    uint32_t *skey = reinterpret_cast<uint32_t *>(smem + PC_GOFF_SKEY);
    uint32_t *clist = reinterpret_cast<uint32_t *>(smem + PC_GOFF_CLIST);
    uint32_t *wq = reinterpret_cast<uint32_t *>(smem + PC_GOFF_WQ);
    __shared__ PcLargeFrame F;
    __shared__ PcGroupCtx C;
    __shared__ PcGroupSearch S;
    __shared__ uint32_t wsum[32];
    __shared__ long long s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fb = blockIdx.y, f = f0 + fb;
    const bool self = a.self_mode != 0;
    if (tid == 0) {
        F = fr[fb];
        C.fr1 = a.xyz1 + (long long)f * a.pitch1;
        C.fr2 = self ? C.fr1 : a.xyz2 + (long long)f * a.pitch2;
        C.hptr1 = a.t.lhptr1; C.hidx1 = a.t.lhidx1; C.hptr2 = a.t.lhptr2; C.hidx2 = a.t.lhidx2;
        C.hb_cutoff = a.hb_cutoff; C.hb_angle = a.hb_angle;
        C.hslice = a.hbonds + (long long)f * o.cap_h_frame; C.cap_h = o.cap_h_frame;
        C.hcursor = &fr[fb].hcursor; C.status = 0;
        C.sB = sortB + (size_t)fb * nBh_alloc; C.lgb2 = acc.lgb2;
        C.slice = a.contacts + (long long)f * o.cap_c_frame; C.cap_c = o.cap_c_frame; C.ngroups2 = acc.ngroups2;
        C.ccursor = &fr[fb].ccursor; C.self = a.self_mode;
    }
    for (int s = tid; s < TAB; s += NT) {
        tkey[s] = PC_EMPTY_KEY; tcls[s] = 0.f; tcls[TAB + s] = 0.f; tcls[2 * TAB + s] = 0.f; tcls[3 * TAB + s] = 0.f;
        tncont[s] = 0u; tnhb[s] = 0u;
    }
    __syncthreads();
    if (F.ncell == 0) return;                                   // no grid: no contacts, no rows (rcount stays 0)
    const int nx = F.nx, ny = F.ny, nz = F.nz;
    const uint32_t h0 = __ldg(g.unit_start + blockIdx.x), h1 = __ldg(g.unit_start + blockIdx.x + 1);
    const float *fr1 = a.xyz1 + (long long)f * a.pitch1;
    const uint32_t *cB = cellsB + (size_t)fb * (capCells + 1);  // after the scatter: cB[c] == END of cell c
    const float4 *sB = sortB + (size_t)fb * nBh_alloc;
    const unsigned ltmask = (1u << lane) - 1u;
    unsigned long long evals = 0;

    // ---- G1: the unit's heavy A atoms, read from the frame, sorted by cell inside the CTA -------------------------
    // (cell order: the lanes of a warp then work on ONE cell's neighbourhood, whose B atoms they load coalesced; the
    // unit's table does not care in which order its atoms are searched)
    {
        float4 *stA = reinterpret_cast<float4 *>(wq);           // staging in the (not yet used) queue memory, thread order
        int2 *stRS = reinterpret_cast<int2 *>(stA + NT);
        uint32_t *stG = reinterpret_cast<uint32_t *>(stRS + NT);
        uint32_t key = 0xffffff00u | (uint32_t)tid;
        if (h0 + (uint32_t)tid < h1) {
            const uint32_t info = __ldg(a.t.hinfo1 + h0 + tid);
            const int li = PC_W_INDEX(info);
            const float *p = fr1 + 3ll * li;
            const float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
            stA[tid] = make_float4(-x, -y, -z, __uint_as_float(info));
            stG[tid] = __ldg(acc.lgb1 + li);
            stRS[tid] = self ? make_int2(__ldg(a.t.lres1 + li), __ldg(a.t.lseg1 + li)) : make_int2(0, 0);
            // same expressions as the bin pass (pc_large_cell_of): an atom outside the grid has no partner
            const float ux = (x - F.lo[0]) * F.inv, uy = (y - F.lo[1]) * F.inv, uz = (z - F.lo[2]) * F.inv;
            if (ux >= 0.f && uy >= 0.f && uz >= 0.f) {
                const int ix = (int)ux, iy = (int)uy, iz = (int)uz;
                if (ix < nx && iy < ny && iz < nz) key = ((uint32_t)((iz * ny + iy) * nx + ix) << 8) | (uint32_t)tid;
            }
        }
        // rank sort (keys are unique: cell << 8 | thread; atoms outside the grid sort behind every cell): two barriers
        // instead of a sorting network's 36
        uint32_t *ukey = reinterpret_cast<uint32_t *>(stG + NT);
        ukey[tid] = key;
        __syncthreads();
        {
            int rank = 0;
            const uint4 *uk4 = reinterpret_cast<const uint4 *>(ukey);
#pragma unroll 8
            for (int q = 0; q < NT / 4; q++) {
                const uint4 v = uk4[q];
                rank += (v.x < key) + (v.y < key) + (v.z < key) + (v.w < key);
            }
            skey[rank] = key;
        }
        __syncthreads();
        const uint32_t mine = skey[tid];
        const bool valid = mine < 0xffffff00u;
        if (valid) { const int src = (int)(mine & 255u); sA[tid] = stA[src]; sG[tid] = stG[src]; sRS[tid] = stRS[src]; }
        const bool head = valid && (tid == 0 || (skey[tid - 1] >> 8) != (mine >> 8));
        const unsigned bal = __ballot_sync(PC_FULL_MASK, head), vbal = __ballot_sync(PC_FULL_MASK, valid);
        if (lane == 0) wsum[warp] = (uint32_t)__popc(bal) | ((uint32_t)__popc(vbal) << 16);
        __syncthreads();
        uint32_t before = 0, total = 0;
        for (int w = 0; w < NW; w++) { const uint32_t v = wsum[w]; if (w < warp) before += v & 0xffffu; total += v; }
        if (head) clist[before + __popc(bal & ltmask)] = (uint32_t)tid;
        if (tid == 0) { clist[total & 0xffffu] = total >> 16; clist[NT + 3] = total & 0xffffu; }
        __syncthreads();                                        // (the staging memory becomes the hit queues)
    }
    const int ncl = (int)clist[NT + 3];

    // ---- G2 / G3: search (pc_group_search) and drain (pc_group_drain) alternate until the warp's cells are done -----------
    if (tid == 0) {
        S.sB = sB; S.sBrs = sortRS ? sortRS + (size_t)fb * nBh_alloc : nullptr; S.cB = cB; S.nx = nx; S.ny = ny; S.nz = nz; S.ncl = ncl; S.sqdist = a.sqdist;
        S.inv_nx = 1.0f / (float)nx; S.inv_ny = 1.0f / (float)ny; S.next = 0;
    }
    if (tid < NW) { S.ci[tid] = -1; S.t0[tid] = 0; S.k0[tid] = 0; S.evals[tid] = 0ull; }
    __syncthreads();
    for (;;) {
        const unsigned wn = self ? pc_group_search<true>(&S) : pc_group_search<false>(&S);
        __syncwarp();
        const bool done = *(volatile int *)&S.ci[warp] >= ncl;
        pc_group_drain(&C, wn);
        if (done) break;
    }
    if (lane == 0) evals = S.evals[warp];
    __syncthreads();

    // ---- G4: the unit's rows -> the frame's row slice --------------------------------------------------------------
    constexpr int ch = TAB / NT;
    const int sb = tid * ch;
    uint32_t n = 0;
#pragma unroll
    for (int s = 0; s < ch; s++) n += tkey[sb + s] != PC_EMPTY_KEY;
    const uint32_t incl = warp_incl_scan(n, lane);
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const uint32_t v = lane < NW ? wsum[lane] : 0;
        const uint32_t vi = warp_incl_scan(v, lane);
        wsum[lane] = vi - v;
        const uint32_t total = __shfl_sync(PC_FULL_MASK, vi, 31);
        if (lane == 0) {
            long long base = -1;
            if (total) {
                // the count keeps growing past the slice's end: pc_group_offsets_kernel reports what would be needed
                base = (long long)atomicAdd(&acc.frame_rcount[f], total);
                if (base + total > g.cap_r_frame) base = -1;
            }
            s_base = base;
        }
    }
    __syncthreads();
    if (s_base >= 0) {
        pc_row *out = g.slices + (long long)f * g.cap_r_frame + s_base + wsum[warp] + incl - n;
#pragma unroll
        for (int s = 0; s < ch; s++) {
            const int q = sb + s;
            if (tkey[q] == PC_EMPTY_KEY) continue;
            pc_row r;
            const double c0 = tcls[q], c1 = tcls[TAB + q], c2 = tcls[2 * TAB + q], c3 = tcls[3 * TAB + q];
            r.key = tkey[q]; r.score = (c0 + c1) + (c2 + c3); r.bb1 = c2 + c3; r.bb2 = c1 + c3;
            r.ncontacts = tncont[q]; r.nhbonds = tnhb[q]; r.first = tkey[q];
            *out++ = r;
        }
    }
    if (tid == 0 && C.status) atomicOr(&a.counters[PC_CNT_STATUS], (unsigned long long)C.status);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) evals += __shfl_xor_sync(PC_FULL_MASK, evals, s);
    if (lane == 0 && evals) atomicAdd(&a.counters[PC_CNT_EVALS], evals);
}

// One CTA: frame_rstart = exclusive scan of the frames' row counts, total -> counters.  A frame whose rows did not fit
// its slice raises PC_ST_OUT_OVERFLOW and reports nframes x the largest frame as the need (slices are equal).
__global__ void __launch_bounds__(1024) pc_group_offsets_kernel(const PcAccArgs a, const long long cap_r_frame) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry, s_max;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { carry = 0; s_max = 0; }
    __syncthreads();
    for (int base = 0; base < a.nframes; base += 1024) {
        const int f = base + tid;
        const uint32_t raw = f < a.nframes ? a.frame_rcount[f] : 0;
        if (raw) atomicMax(&s_max, raw);
        const uint32_t v = (uint32_t)min((long long)raw, cap_r_frame);
        if (f < a.nframes && v != raw) a.frame_rcount[f] = v;
        const uint32_t incl = warp_incl_scan(v, lane);
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) { const uint32_t w = wsum[lane]; const uint32_t wi = warp_incl_scan(w, lane); wsum[lane] = wi - w; }
        __syncthreads();
        if (f < a.nframes) a.frame_rstart[f] = carry + wsum[warp] + incl - v;
        __syncthreads();
        if (tid == 1023) carry += wsum[warp] + incl;
        __syncthreads();
    }
    if (tid == 0) {
        if ((long long)s_max > cap_r_frame) {
            a.counters[PC_CNT_ROWS] = (unsigned long long)s_max * (unsigned long long)a.nframes;
            atomicOr(&a.counters[PC_CNT_STATUS], PC_ST_OUT_OVERFLOW);
        } else {
            a.counters[PC_CNT_ROWS] = carry;
            if ((long long)carry > a.cap_rows) atomicOr(&a.counters[PC_CNT_STATUS], PC_ST_OUT_OVERFLOW);
        }
    }
}

// the frames' row slices -> back to back in a.rows (16-byte words: a pc_row is three of them)
__global__ void __launch_bounds__(256) pc_group_gather_kernel(const PcAccArgs a, const pc_row *slices, const long long cap_r_frame) {
    const int f = blockIdx.y;
    const long long n = 3ll * a.frame_rcount[f], dst0 = 3ll * a.frame_rstart[f];
    if (dst0 + n > 3ll * a.cap_rows) return;
    const uint4 *src = reinterpret_cast<const uint4 *>(slices + (long long)f * cap_r_frame);
    uint4 *dst = reinterpret_cast<uint4 *>(a.rows) + dst0;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) dst[k] = src[k];
}
