// pc_scan_tail.cuh -- large-frame path, everything behind the streaming kernels in ONE kernel per batch.
//
// A 1 M-atom frame (protein vs. lipid + water) is 12 MB of coordinates of which a few
// thousand atoms per selection can have a partner: the streaming kernels of pc_scan_large.cuh (the HBM-bound
// part) compact exactly those -- B atoms inside the grid around selection A, A atoms whose 27-cell
// neighbourhood holds a B atom -- into small float4 arrays that stay in L2.  Everything after that
// (cell sort, pair search of gridsearch.C:1111-1142, filters, distance + weight of
// core/multi_trajectory.py:98-107, the H-bond test of :112-209 and the per-frame accumulation of
// core/ContactAnalyzer.py:527-559) used to be eight more launches over all frames, each latency-bound on a few
// thousand atoms per frame.  Here ONE CTA owns one frame and does all of it out of shared memory:
//   T1  the frame's B cell counts (left in global memory by the bin kernel) -> shared memory, exclusive scan
//   T2  compacted B atoms -> cell order in shared memory (float4: x, y, z, packed atom word)
//   T3  pair search + drain, warp-autonomous: a warp takes 32 (A atom, stencil z-layer) items at a time (A atoms
//       come straight from the compacted array in L2, already in residue order), tests the layer's three B runs
//       out of shared memory with the reference's exact float32 expression (pair_within_packed), queues its
//       hits in its own shared-memory queue and, whenever >= 96 are queued, turns them into contact records
//       (float64 distance, weight; one coalesced run of 16-byte records per round).  No CTA barrier inside.
//   T3b N/O/S pairs queued by the drain: one thread per (pair, side) runs the donor-H...acceptor test
//   T4  (maps set, no order keys) the frame's records are grouped by key in the compact shared-memory table of
//       pc_accumulate.cuh (pc_compact_frame), laid over the memory T1-T3 used
// Frames whose in-grid B atoms + cell offsets do not fit the pool raise PC_ST_TAIL_OVERFLOW; the caller then
// re-runs the batch on the multi-kernel sequence (pc_set_tail(ctx, 0)), which has no such limit.
#pragma once
#include "pc_scan_large.cuh"
#include "pc_accumulate.cuh"

#define PC_TAIL_NT 1024
#ifndef PC_TAIL_UNROLL
#define PC_TAIL_UNROLL 4
#endif
constexpr int pc_tail_unroll = PC_TAIL_UNROLL;
#define PC_TAIL_WQ 256                   // hits per warp queue
#ifndef PC_TAIL_DRAIN_AT
#define PC_TAIL_DRAIN_AT 96              // a warp drains its queue once this many hits wait (>= 3 full rounds of lanes)
#endif
#define PC_TAIL_GATE 4096                // N/O/S pairs per frame queued for the H-bond pass (more are tested in place)
#define PC_TAIL_SMEM PC_COMPACT_SMEM     // dynamic shared memory: the accumulation table of T4 is the larger tenant
#define PC_TAIL_POOL (PC_TAIL_SMEM - (PC_TAIL_NT / 32) * PC_TAIL_WQ * 4 - PC_TAIL_GATE * 4)   // cell offsets + sorted B

// Every phase is a lambda over the kernel's locals.  (Measured against it and dropped: the rarely executed phases -- records,
// H-bond tests, accumulation -- compiled out of line with their state in shared memory, 225-290 us per 148 frames of c3
// against 209 us; and a "walk" kernel that balances the candidates of 32 A atoms over the lanes of a warp, 2.3x slower:
// spills under the 64-register cap.  Both are in the repository's history, commits ac37fd8 and 8a0aded.)
template <bool AFRAME>
__global__ void __launch_bounds__(PC_TAIL_NT, 1) pc_large_tail_inl_kernel(const PcScanArgs a, PcLargeFrame *fr, const int f0,
                                                                          const int nfb, const uint32_t *cellsA, const uint32_t *cellsB,
                                                                          const int capCells, const float4 *candA, const uint32_t *ccellA,
                                                                          float4 *sortA_g, const float4 *candB, const uint32_t *ccellB,
                                                                          const int n1h_alloc, const int n2h_alloc, const PcLargeOut o,
                                                                          const PcAccArgs acc, const int fuse_acc) {
    constexpr int NT = PC_TAIL_NT, NW = NT / 32;
    extern __shared__ __align__(16) unsigned char smem[];
    uint32_t *wq = reinterpret_cast<uint32_t *>(smem);                 // [NW][PC_TAIL_WQ]
    uint32_t *gateq = wq + NW * PC_TAIL_WQ;                            // [PC_TAIL_GATE]
    unsigned char *pool = reinterpret_cast<unsigned char *>(gateq + PC_TAIL_GATE);
    __shared__ PcLargeFrame F;
    __shared__ uint32_t wsum[32];
    __shared__ unsigned int wqn[NW];
    __shared__ unsigned int s_next, s_ccur, s_hcur, s_ngate, s_status;
    __shared__ unsigned int s_keep[2];
    __shared__ long long s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float sqdist = a.sqdist;
    unsigned long long evals = 0;
    uint32_t *mywq = wq + warp * PC_TAIL_WQ;
    volatile unsigned int *vqn = wqn;

    for (int fb = blockIdx.x; fb < nfb; fb += gridDim.x) {
        const int f = f0 + fb;
        if (tid == 0) { F = fr[fb]; s_next = 0; s_ccur = 0; s_hcur = 0; s_ngate = 0; s_status = 0; }
        if (tid < NW) wqn[tid] = 0;
        __syncthreads();
        const int ncell = F.ncell, nB = (int)F.nB_in;
        int nA = AFRAME ? a.t.n1h : (int)F.nA_in;                 // AFRAME: the in-grid count comes out of T0
        const size_t cell_bytes = ((size_t)(ncell + 1) * 4 + 15) & ~(size_t)15;
        // AFRAME: A's cell offsets + the occupancy bitmap live in the (not yet used) hit-queue memory during T0
        const bool fits = nA <= 65535 && nB <= 65535 && cell_bytes + 16ull * (size_t)nB <= (size_t)PC_TAIL_POOL &&
                          (!AFRAME || (size_t)(ncell + 1 + (ncell + 31) / 32) * 4 <= (size_t)(NW * PC_TAIL_WQ + PC_TAIL_GATE) * 4);
        uint32_t *cellB_s = reinterpret_cast<uint32_t *>(pool);
        float4 *sortB = reinterpret_cast<float4 *>(pool + cell_bytes);
        float4 *sA = sortA_g + (size_t)fb * n1h_alloc;           // this frame's A atoms in cell order (written in T0, plain loads)
        pc_contact *slice = a.contacts + (long long)f * o.cap_c_frame;
        pc_hbond *hslice = a.hbonds + (long long)f * o.cap_h_frame;
        const float *fr1 = a.xyz1 + (long long)f * a.pitch1;
        const float *fr2 = a.xyz2 + (long long)f * a.pitch2;

        // One donor-H...acceptor side of a gated pair (core/multi_trajectory.py:112-209).
        // side 0: donor = atom1, acceptor = atom2; side 1: donor = atom2, acceptor = atom1
        auto hbond_side = [&](const unsigned u, const int side) {
            const float4 pa = sA[u >> 16], pb = sortB[u & 0xffffu];
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
                const unsigned hq = atomicAdd(&s_hcur, 1u);
                if ((long long)hq < o.cap_h_frame) {
                    pc_hbond r;
                    r.i = li; r.j = lj; r.hyd = (uint32_t)hl | (side ? 0x80000000u : 0u);
                    r.distance = (float)d; r.angle = (float)ang;
                    hslice[hq] = r;
                } else {
                    atomicOr(&s_status, (unsigned)PC_ST_OUT_OVERFLOW);
                }
            }
        };
        // one hit -> contact record `slot` of the frame's slice; N/O/S pairs are queued for T3b
        auto emit = [&](const unsigned u, const long long slot) {
            const float4 pa = sA[u >> 16], pb = sortB[u & 0xffffu];
            const unsigned wa = __float_as_uint(pa.w), wb = __float_as_uint(pb.w);
            const double dx = (double)pa.x - (double)pb.x, dy = (double)pa.y - (double)pb.y, dz = (double)pa.z - (double)pb.z;
            const double dist = sqrt(dx * dx + dy * dy + dz * dz);
            if (slot < o.cap_c_frame) {
                pc_contact rec;
                rec.i = PC_W_INDEX(wa); rec.j = PC_W_INDEX(wb); rec.distance = (float)dist; rec.weight = contact_weight_f32(dist);
                slice[slot] = rec;
            } else {
                atomicOr(&s_status, (unsigned)PC_ST_OUT_OVERFLOW);
            }
            if (wa & wb & PC_W_HBCLASS) {
                const unsigned q = atomicAdd(&s_ngate, 1u);
                if (q < (unsigned)PC_TAIL_GATE) gateq[q] = u;
                else { hbond_side(u, 0); hbond_side(u, 1); }         // queue full: tested in place
            }
        };
        // the warp's queued hits -> records (lanes over hits, one claim of the frame's cursor per call)
        auto drain_queue = [&]() {
            __syncwarp();
            const unsigned n = min(vqn[warp], (unsigned)PC_TAIL_WQ);
            if (n) {
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(&s_ccur, n);
                base = __shfl_sync(PC_FULL_MASK, base, 0);
                for (unsigned k = lane; k < n; k += 32) emit(mywq[k], (long long)base + k);
            }
            __syncwarp();
            if (lane == 0) wqn[warp] = 0;
            __syncwarp();
        };

        if (!fits) {
            if (tid == 0) atomicOr(&s_status, (unsigned)PC_ST_TAIL_OVERFLOW);
        } else if (ncell > 0 && nA > 0 && nB > 0) {                // (AFRAME: nA is still the number of heavy A atoms here)
            const int nx = F.nx, ny = F.ny, nz = F.nz;
            const float lox = F.lo[0], loy = F.lo[1], loz = F.lo[2], inv = F.inv;
            // ---- T0: selection A's compacted atoms -> cell order, in a global scratch array (L2): the lanes of a warp
            //      then hold atoms of the same few cells, walk the same B runs and read the same shared-memory words
            //      (broadcasts instead of the bank conflicts of selection-ordered atoms).  The cell offsets live where
            //      B's will be in a moment.
            if (!AFRAME) {
                const uint32_t *gA = cellsA + (size_t)fb * (capCells + 1);
                for (int c = tid; c < ncell; c += NT) cellB_s[c] = __ldg(gA + c);
                __syncthreads();
                block_excl_scan_inplace<NT>(cellB_s, ncell, wsum);
                const float4 *cA = candA + (size_t)fb * n1h_alloc;
                const uint32_t *ccA = ccellA + (size_t)fb * n1h_alloc;
                for (int p = tid; p < nA; p += NT) sA[atomicAdd(&cellB_s[__ldg(ccA + p)], 1u)] = __ldg(cA + p);
                __syncthreads();
            }
            // ---- T1: B cell counts -> shared memory, exclusive scan -------------------------------------------
            const uint32_t *gB = cellsB + (size_t)fb * (capCells + 1);
            for (int c = tid; c < ncell; c += NT) cellB_s[c] = __ldg(gB + c);
            __syncthreads();
            block_excl_scan_inplace<NT>(cellB_s, ncell, wsum);
            // ---- T2: compacted B atoms -> cell order; afterwards cellB_s[c] == END of cell c -----------------
            const float4 *cB = candB + (size_t)fb * n2h_alloc;
            const uint32_t *ccB = ccellB + (size_t)fb * n2h_alloc;
            for (int p = tid; p < nB; p += NT) sortB[atomicAdd(&cellB_s[__ldg(ccB + p)], 1u)] = __ldg(cB + p);
            __syncthreads();
            // x-triples, in place: cell c's END becomes start | length << 16 of the B atoms in cells cx-1 .. cx+1 of c's
            // row (ONE lookup per stencil row in the search).  NT cells at a time, lowest first; the two ends just below a
            // chunk are kept aside, because the chunk before has already overwritten them (a triple reads the ends of
            // cells c-2 .. c+1).
            uint32_t *tri = cellB_s;
            if (tid < 2) s_keep[tid] = 0u;
            __syncthreads();
            for (int lo = 0; lo < ncell; lo += NT) {
                const int c = lo + tid;
                unsigned j0 = 0, j1 = 0, mine = 0;
                if (c < ncell) {
                    const int cx = c % nx, row = c - cx;
                    const int i0 = row + max(cx - 1, 0) - 1;
                    j0 = i0 < 0 ? 0u : (i0 >= lo ? cellB_s[i0] : s_keep[i0 - (lo - 2)]);
                    j1 = cellB_s[row + min(cx + 1, nx - 1)];
                    mine = cellB_s[c];
                }
                __syncthreads();
                if (c < ncell) {
                    cellB_s[c] = j0 | ((j1 - j0) << 16);          // nB <= 65535: both halves fit
                    if (tid >= NT - 2) s_keep[tid - (NT - 2)] = mine;
                }
                __syncthreads();
            }
            if (AFRAME) {
                // ---- T0 (A from the frame): no streaming bin and no occupancy pass for selection A.  Its heavy atoms are
                //      read through the heavy-atom list straight from the frame (0.6 MB per 1 M-atom frame), kept when their
                //      cell's 27-neighbourhood holds a B atom (bitmap from the x-triples), and counting-sorted by cell into
                //      the global scratch array: count pass, scan, scatter pass (the second pass hits L1 / L2).
                uint32_t *cellA_s = wq;
                uint32_t *occ = cellA_s + (ncell + 1);
                for (int c0 = warp * 32; c0 < ncell; c0 += NT) {
                    const int c = c0 + lane;
                    bool any = false;
                    if (c < ncell) {
                        const int cx = c % nx, cy = (c / nx) % ny, cz = c / (nx * ny);
                        for (int z = max(cz - 1, 0); z <= min(cz + 1, nz - 1); z++)
                            for (int y = max(cy - 1, 0); y <= min(cy + 1, ny - 1); y++)
                                any |= (tri[(z * ny + y) * nx + cx] >> 16) != 0u;
                    }
                    const unsigned bal = __ballot_sync(PC_FULL_MASK, any);
                    if (lane == 0) occ[c0 >> 5] = bal;
                }
                for (int c = tid; c <= ncell; c += NT) cellA_s[c] = 0u;
                __syncthreads();
                const uint32_t *hinfo1 = a.t.hinfo1;
                const int n1h = a.t.n1h;
                // cell of an A atom when it can have a partner, else -1 (same expressions as the search's)
                auto cell_a = [&](const float x, const float y, const float z) -> int {
                    const float ux = (x - lox) * inv, uy = (y - loy) * inv, uz = (z - loz) * inv;
                    if (!(ux >= 0.f && uy >= 0.f && uz >= 0.f)) return -1;
                    const int cx = (int)ux, cy = (int)uy, cz = (int)uz;
                    if (cx >= nx || cy >= ny || cz >= nz) return -1;
                    const int c = (cz * ny + cy) * nx + cx;
                    return ((occ[c >> 5] >> (c & 31)) & 1u) ? c : -1;
                };
                // pass 1 over the frame: kept atoms are compacted (arrival order, CTA-local cursor) and counted per cell;
                // eight atoms per thread and round: the index loads, then the coordinate loads, are issued back to back (the
                // pass is a chain of two dependent loads per atom, DRAM-cold: latency, not bandwidth)
                float4 *cAw = const_cast<float4 *>(candA) + (size_t)fb * n1h_alloc;
                uint32_t *ccAw = const_cast<uint32_t *>(ccellA) + (size_t)fb * n1h_alloc;
                for (int k0 = tid; k0 < n1h; k0 += 8 * NT) {
                    uint32_t info[8];
                    float x[8], y[8], z[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) info[u] = __ldg(hinfo1 + min(k0 + u * NT, n1h - 1));
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const float *p = fr1 + 3ll * PC_W_INDEX(info[u]);
                        x[u] = __ldg(p); y[u] = __ldg(p + 1); z[u] = __ldg(p + 2);
                    }
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const int c = k0 + u * NT < n1h ? cell_a(x[u], y[u], z[u]) : -1;
                        if (c < 0) continue;
                        const unsigned pos = atomicAdd(&s_next, 1u);          // (the search's item counter, not in use yet)
                        cAw[pos] = make_float4(x[u], y[u], z[u], __uint_as_float(info[u]));
                        ccAw[pos] = (uint32_t)c;
                        atomicAdd(&cellA_s[c], 1u);
                    }
                }
                __syncthreads();
                block_excl_scan_inplace<NT>(cellA_s, ncell, wsum);
                // pass 2 over the compacted atoms (this CTA wrote them: plain loads): scatter into cell order
                const int kept = (int)s_next;
                for (int p = tid; p < kept; p += NT) sA[atomicAdd(&cellA_s[ccAw[p]], 1u)] = cAw[p];
                __syncthreads();
                if (tid == 0) s_next = 0;
                nA = (int)cellA_s[ncell];
                __syncthreads();                                  // the hit queues take their memory back
            }
            // ---- T3: pair search + drain, warp-autonomous -----------------------------------------------------
            const int nitems = 3 * nA;
            for (;;) {
                int w = 0;
                if (lane == 0) w = (int)atomicAdd(&s_next, 32u);
                w = __shfl_sync(PC_FULL_MASK, w, 0);
                if (w >= nitems) break;
                w += lane;
                if (w < nitems) {
                    // layer-major items over the cell-sorted A atoms: the lanes of a warp hold the atoms of a few
                    // neighbouring cells on the same stencil layer
                    const int layer = (w >= nA) + (w >= 2 * nA), ia = w - layer * nA;
                    const float4 pa = sA[ia];
                    const int cx = (int)((pa.x - lox) * inv), cy = (int)((pa.y - loy) * inv), cz = (int)((pa.z - loz) * inv);
                    const int z = cz + layer - 1;
                    if (z >= 0 && z < nz) {
                        // the layer's three rows as ONE list of candidates: run r covers list positions [e_{r-1}, e_r)
                        const int row1 = (z * ny + cy) * nx + cx;
                        const unsigned t0 = cy > 0 ? tri[row1 - nx] : 0u, t1 = tri[row1], t2 = cy + 1 < ny ? tri[row1 + nx] : 0u;
                        const int e0 = (int)(t0 >> 16), e1 = e0 + (int)(t1 >> 16), e2 = e1 + (int)(t2 >> 16);
                        // list position k -> B position: k + (start of its run - list position of that start)
                        const int d0 = (int)(t0 & 0xffffu), d1 = (int)(t1 & 0xffffu) - e0, d2 = (int)(t2 & 0xffffu) - e1;
                        const float2 nxy = make_float2(-pa.x, -pa.y);
                        const float naz = -pa.z;
                        evals += (unsigned)e2;
                        for (int kb = 0; kb < e2; kb += 32) {
                            const int cnt = min(32, e2 - kb);
                            unsigned m = 0;
#pragma unroll pc_tail_unroll
                            for (int k = 0; k < cnt; k++) {
                                const int kk = kb + k;
                                const int j = kk + (kk < e0 ? d0 : (kk < e1 ? d1 : d2));
                                if (pair_within_packed(nxy, naz, sortB[j], sqdist)) m |= 1u << k;
                            }
                            if (!m) continue;
                            unsigned p = atomicAdd(&wqn[warp], (unsigned)__popc(m));
                            while (m) {
                                const int kk = kb + __ffs(m) - 1;
                                m &= m - 1;
                                const unsigned u = ((unsigned)ia << 16) | (unsigned)(kk + (kk < e0 ? d0 : (kk < e1 ? d1 : d2)));
                                if (p < (unsigned)PC_TAIL_WQ) mywq[p] = u;
                                else emit(u, (long long)atomicAdd(&s_ccur, 1u));     // queue full: this lane drains its own hit
                                p++;
                            }
                        }
                    }
                }
                __syncwarp();
                if (vqn[warp] >= (unsigned)PC_TAIL_DRAIN_AT) drain_queue();
            }
            drain_queue();
            __syncthreads();
            // ---- T3b: donor-H...acceptor tests, one thread per (gated pair, side) ------------------------------
            const unsigned ngate = min(s_ngate, (unsigned)PC_TAIL_GATE);
            for (unsigned it = tid; it < 2u * ngate; it += NT) hbond_side(gateq[it >> 1], (int)(it & 1u));
        }
        __syncthreads();
        // ---- per-frame counts and totals (what pc_large_finish_kernel does on the multi-kernel sequence) -----------
        if (tid == 0) {
            const unsigned cc = s_ccur, hc = s_hcur;
            a.frame_cstart[f] = (uint32_t)((long long)f * o.cap_c_frame);
            a.frame_ccount[f] = (uint32_t)min((long long)cc, o.cap_c_frame);
            a.frame_hstart[f] = (uint32_t)((long long)f * o.cap_h_frame);
            a.frame_hcount[f] = (uint32_t)min((long long)hc, o.cap_h_frame);
            // the totals report what WOULD be needed so the host can re-reserve after an overflow
            if (cc) atomicAdd(&a.counters[PC_CNT_CONTACTS], (unsigned long long)cc);
            if (hc) atomicAdd(&a.counters[PC_CNT_HBONDS], (unsigned long long)hc);
            atomicMax(&a.counters[6], (unsigned long long)cc);
            atomicMax(&a.counters[7], (unsigned long long)hc);
            if (s_status) atomicOr(&a.counters[PC_CNT_STATUS], (unsigned long long)s_status);
            if (fuse_acc && cc == 0) { acc.frame_rstart[f] = 0; acc.frame_rcount[f] = 0; }
            // the frame's work slot is left ready for the next batch (no reset launch in front of it)
            PcLargeFrame &G = fr[fb];
            for (int k = 0; k < 3; k++) { G.bbox[k] = 0xffffffffu; G.bbox[3 + k] = 0u; G.bbox[6 + k] = 0xffffffffu; G.bbox[9 + k] = 0u; }
            G.nA_in = G.nB_in = 0; G.ccursor = G.hcursor = 0; G.ncell = 0; G.pad[0] = G.pad[1] = 0;
        }
        __syncthreads();
        // ---- T4: per-frame accumulation of the records just written (they are in L2) -----------------------------
        const bool have = fuse_acc && s_ccur != 0;
        const bool do_acc = have && !(s_status & (unsigned)(PC_ST_OUT_OVERFLOW | PC_ST_TAIL_OVERFLOW));
        __syncthreads();                   // pc_compact_frame reuses s_status
        if (do_acc) pc_compact_frame<NT>(acc, f, smem, wsum, &s_base, &s_status);
        else if (have && tid == 0) { acc.frame_rstart[f] = 0; acc.frame_rcount[f] = 0; }
        __syncthreads();
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) evals += __shfl_xor_sync(PC_FULL_MASK, evals, s);
    if (lane == 0 && evals) atomicAdd(&a.counters[PC_CNT_EVALS], evals);
}

// ---- host side of the large-frame path --------------------------------------------------------------------------
static inline int pc_large_fail(std::string &err, const char *what, cudaError_t e) {
    err = std::string(what) + ": " + cudaGetErrorString(e);
    return PC_ERR_CUDA;
}
#define PCL_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return pc_large_fail(err, #expr, _e); } while (0)

// CTAs per frame of a streaming kernel over n atoms: `waves` waves of resident CTAs (2 per SM) over the nfb frames of the
// launch, but every CTA walks at least four chunks; atom ranges start on multiples of 32.  Every CTA pays a prologue
// (frame state, mbarriers, chunk offsets, occupancy bitmap) and a pipeline fill of ~5 us, so few, long-running CTAs
// beat many short ones as long as the last wave is full: 148 frames x 2 CTAs = one wave exactly.
static inline void pc_stream_geom(int n, int nfb, int num_sms, int waves, int &ctas, int &per) {
    const int atoms_chunk = PC_STREAM_CHUNK / 12;
    const int want = std::max(1, (waves * 2 * num_sms + nfb / 2) / nfb);
    const int most = std::max(1, n / (4 * atoms_chunk));
    ctas = std::min(want, most);
    per = ((n + ctas - 1) / ctas + 31) & ~31;
    ctas = (n + per - 1) / per;
}

// tail: 0 = multi-kernel sequence, 1 = fused tail kernel on the compacted A and B atoms, 3 = fused tail kernel that reads
// selection A from the frame (default); two selections only.  acc != nullptr with a tail kernel: rows are produced too.
static inline int pc_large_scan(PcLargeWork &W, const PcScanArgs &a, int num_sms, cudaStream_t s, std::string &err,
                                unsigned long long *launches, cudaEvent_t probe0, cudaEvent_t probe1, int tail,
                                const PcAccArgs *acc, const PcGroupArgs *group = nullptr) {
    const int n1 = a.t.n1, n2 = a.self_mode ? a.t.n1 : a.t.n2;
    const int n1h = a.t.n1h, n2h = a.self_mode ? a.t.n1h : a.t.n2h;
    const bool self = a.self_mode != 0;
    if (a.stride != 3) { err = "the large-frame path reads packed xyz (PC_LAYOUT_XYZ) only"; return PC_ERR_INVALID; }
    if ((((uintptr_t)a.xyz1 | (uintptr_t)a.xyz2) & 3) != 0) { err = "coordinates must be 4-byte aligned"; return PC_ERR_INVALID; }
    if (self) tail = 0;
    // cells per frame: a few per atom of the smaller side, bounded like the reference's MAXBOXES
    long long cc = 8ll * std::min(n1h, n2h);
    cc = std::max(4096ll, std::min(cc, 4000000ll));
    const int capCells = (int)cc;
    const int a1 = std::max(n1h, 1), a2 = std::max(n2h, 1);
    // frames in flight: bound the work memory to ~6 GB
    const size_t per_frame = (size_t)(capCells + 1) * 4 * (self ? 1 : 2) + (size_t)a1 * (self ? 44 : 36) + (self ? 0 : (size_t)a2 * 36);
    int FB = (int)std::max<size_t>(1, std::min<size_t>(296, (6ull << 30) / std::max<size_t>(per_frame, 1)));
    FB = std::min(FB, a.nframes);
    if (W.FB < FB || W.capCells != capCells || W.n1h != n1h || W.n2h != n2h || W.self_mode != a.self_mode) {
        W.release();
        PCL_TRY(cudaMalloc(&W.frames, sizeof(PcLargeFrame) * FB));
        PCL_TRY(cudaMalloc(&W.cellA, (size_t)FB * (capCells + 1) * 4));
        PCL_TRY(cudaMalloc(&W.candA, (size_t)FB * a1 * sizeof(float4)));
        PCL_TRY(cudaMalloc(&W.ccellA, (size_t)FB * a1 * 4));
        PCL_TRY(cudaMalloc(&W.sortA, (size_t)FB * a1 * sizeof(float4)));
        if (self) PCL_TRY(cudaMalloc(&W.sortRS, (size_t)FB * a1 * sizeof(int2)));
        if (!self) {
            PCL_TRY(cudaMalloc(&W.cellB, (size_t)FB * (capCells + 1) * 4));
            PCL_TRY(cudaMalloc(&W.candB, (size_t)FB * a2 * sizeof(float4)));
            PCL_TRY(cudaMalloc(&W.ccellB, (size_t)FB * a2 * 4));
            PCL_TRY(cudaMalloc(&W.sortB, (size_t)FB * a2 * sizeof(float4)));
            PCL_TRY(cudaMalloc(&W.occ, (size_t)FB * capCells));
            PCL_TRY(cudaMalloc(&W.occ2, (size_t)FB * capCells));
        }
        W.FB = FB; W.capCells = capCells; W.n1h = n1h; W.n2h = n2h; W.self_mode = a.self_mode;
        PCL_TRY(cudaFuncSetAttribute(pc_large_stream_kernel<PC_MODE_BBOX>, cudaFuncAttributeMaxDynamicSharedMemorySize, PC_STREAM_SMEM));
        PCL_TRY(cudaFuncSetAttribute(pc_large_stream_kernel<PC_MODE_BIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, PC_STREAM_SMEM));
        PCL_TRY(cudaFuncSetAttribute(pc_large_tail_inl_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PC_TAIL_SMEM));
        PCL_TRY(cudaFuncSetAttribute(pc_large_tail_inl_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PC_TAIL_SMEM));
        PCL_TRY(cudaFuncSetAttribute(pc_large_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PC_GROUP_SMEM));
        W.clean = false;
    }
    FB = W.FB;
    PcLargeOut o;
    o.cap_c_frame = a.cap_contacts / a.nframes;
    o.cap_h_frame = a.cap_hbonds / a.nframes;
    if (o.cap_c_frame < 1 || o.cap_h_frame < 1) { err = "pc_reserve: capacities smaller than the number of frames"; return PC_ERR_INVALID; }
    const int use_bboxB = (!self && (long long)n2h <= 4ll * n1h) ? 1 : 0;
    for (int f0 = 0; f0 < a.nframes; f0 += FB) {
        const int nfb = std::min(FB, a.nframes - f0);
        int ctasA, perA, ctasB = 0, perB = 0;
        static const int wavesA = getenv("PC_WAVES_A") ? atoi(getenv("PC_WAVES_A")) : 1;       // experiments
        static const int wavesB = getenv("PC_WAVES_B") ? atoi(getenv("PC_WAVES_B")) : 1;
        pc_stream_geom(n1, nfb, num_sms, self ? wavesB : wavesA, ctasA, perA);
        if (!self) pc_stream_geom(n2, nfb, num_sms, wavesB, ctasB, perB);
        // fused tail kernels leave every frame slot reset; the LAST CTA of a frame computes the grid behind the bbox pass
        // and the dilated occupancy behind the bin pass of B (one launch less each)
        const bool fused = tail != 0;
        if (!(fused && W.clean)) {
            pc_large_reset_kernel<<<(W.FB + 63) / 64, 64, 0, s>>>(W.frames, W.FB);
            *launches += 1;
        }
        W.clean = fused;
        PcStreamEpi none_epi;
        memset(&none_epi, 0, sizeof none_epi);
        PcStreamEpi grid_epi = none_epi, occ_epi = none_epi;
        grid_epi.what = PC_EPI_GRID; grid_epi.capCells = capCells; grid_epi.use_bboxB = 0; grid_epi.cellsA = W.cellA; grid_epi.cellsB = W.cellB;
        occ_epi.what = PC_EPI_OCC; occ_epi.capCells = capCells; occ_epi.cellsB = W.cellB; occ_epi.occ = W.occ; occ_epi.occ2 = W.occ2;
        // L2 hints (tail mode 3 reads selection A again in the per-frame kernel): keep A's frames, do not keep B's stream
        static const int l2hint = getenv("PC_L2HINT") ? atoi(getenv("PC_L2HINT")) : 1;           // experiments: 0 = off
        PcStreamEpi b_epi = tail == 1 ? occ_epi : none_epi;
        if (l2hint && tail == 3 && !self) { grid_epi.l2 = 2; none_epi.l2 = 0; b_epi.l2 = 1; }
        const bool epi = fused && !use_bboxB;
        pc_large_stream_kernel<PC_MODE_BBOX><<<dim3(ctasA, nfb), PC_STREAM_NT, PC_STREAM_SMEM, s>>>(
            a, W.frames, f0, 0, 0, perA, nullptr, 0, nullptr, nullptr, 0, nullptr, epi ? grid_epi : none_epi);
        if (use_bboxB)
            pc_large_stream_kernel<PC_MODE_BBOX><<<dim3(ctasB, nfb), PC_STREAM_NT, PC_STREAM_SMEM, s>>>(
                a, W.frames, f0, 1, 6, perB, nullptr, 0, nullptr, nullptr, 0, nullptr, none_epi);
        if (!epi) {
            pc_large_grid_kernel<<<nfb, 256, 0, s>>>(a, W.frames, capCells, use_bboxB, W.cellA, W.cellB);
            *launches += 1;
        }
        if (!self) {
            // B first: its cell counts tell which A atoms can have a partner at all
            if (probe0 && f0 == 0) cudaEventRecord(probe0, s);
            // (measured and dropped: a bin kernel that box-tests whole coordinates before it looks an atom up in the heavy-atom
            // list -- 5.3 TB/s like the list walk)
            pc_large_stream_kernel<PC_MODE_BIN><<<dim3(ctasB, nfb), PC_STREAM_NT, PC_STREAM_SMEM, s>>>(
                a, W.frames, f0, 1, 0, perB, W.cellB, capCells, W.candB, W.ccellB, a2, nullptr, b_epi);
            if (probe1 && f0 == 0) cudaEventRecord(probe1, s);
            if (tail == 3) {
                PcAccArgs none;
                memset(&none, 0, sizeof none);
                pc_large_tail_inl_kernel<true><<<std::min(nfb, num_sms), PC_TAIL_NT, PC_TAIL_SMEM, s>>>(
                    a, W.frames, f0, nfb, W.cellA, W.cellB, capCells, W.candA, W.ccellA, W.sortA, W.candB, W.ccellB, a1, a2, o,
                    acc ? *acc : none, acc ? 1 : 0);
                PCL_TRY(cudaGetLastError());
                *launches += 3 + (use_bboxB ? 1 : 0);
                continue;
            }
            if (tail != 1 && !group) {
                pc_large_occ_kernel<<<dim3(std::max(1, std::min((capCells + 255) / 256, 4 * num_sms / nfb + 1)), nfb), 256, 0, s>>>(
                    W.frames, W.cellB, capCells, W.occ);
                *launches += 1;
            }
        }
        if (self && probe0 && f0 == 0) cudaEventRecord(probe0, s);
        if (self || !group)         // (the group kernel reads selection A from the frame: no bin pass for it)
            pc_large_stream_kernel<PC_MODE_BIN><<<dim3(ctasA, nfb), PC_STREAM_NT, PC_STREAM_SMEM, s>>>(
                a, W.frames, f0, 0, 0, perA, W.cellA, capCells, W.candA, W.ccellA, a1, self ? nullptr : W.occ, none_epi);
        if (self && probe1 && f0 == 0) cudaEventRecord(probe1, s);
        if (tail) {
            PcAccArgs none;
            memset(&none, 0, sizeof none);
            pc_large_tail_inl_kernel<false><<<std::min(nfb, num_sms), PC_TAIL_NT, PC_TAIL_SMEM, s>>>(
                a, W.frames, f0, nfb, W.cellA, W.cellB, capCells, W.candA, W.ccellA, W.sortA, W.candB, W.ccellB, a1, a2, o,
                acc ? *acc : none, acc ? 1 : 0);
            PCL_TRY(cudaGetLastError());
            *launches += 4 + (use_bboxB ? 1 : 0);
            continue;
        }
        pc_large_scan_kernel<<<dim3(nfb, self ? 1 : 2), 1024, 0, s>>>(W.frames, W.cellA, W.cellB, capCells);
        const int per_frame_blocks = std::max(1, 4 * num_sms / nfb);
        const int blocksA = std::max(1, std::min((n1h + 255) / 256, per_frame_blocks));
        const int blocksB = std::max(1, std::min((n2h + 255) / 256, per_frame_blocks));
        pc_large_scatter_kernel<<<dim3(blocksA, nfb), 256, 0, s>>>(W.frames, 0, W.cellA, capCells, W.candA, W.ccellA, W.sortA, a1,
                                                                   self && group ? W.sortRS : nullptr, a.t.lres1, a.t.lseg1);
        if (!self)
            pc_large_scatter_kernel<<<dim3(blocksB, nfb), 256, 0, s>>>(W.frames, 1, W.cellB, capCells, W.candB, W.ccellB, W.sortB, a2);
        if (group) {
            // search + records + H-bonds + rows of the units' keys in one kernel (pc_scan_group.cuh)
            pc_large_group_kernel<<<dim3(group->nunits, nfb), PC_GROUP_NT, PC_GROUP_SMEM, s>>>(
                a, W.frames, f0, self ? W.cellA : W.cellB, capCells, self ? W.sortA : W.sortB, self ? W.sortRS : nullptr,
                self ? a1 : a2, o, *acc, *group);
            pc_large_finish_kernel<<<(nfb + 63) / 64, 64, 0, s>>>(a, W.frames, f0, nfb, o);
            PCL_TRY(cudaGetLastError());
            *launches += 5 + (use_bboxB ? 1 : 0) + (self ? 0 : 2);
            continue;
        }
        // the number of in-grid A atoms is only known on the device: launch for the worst case, surplus CTAs exit
        const int pblocks = (int)std::max<long long>(1, std::min<long long>((3ll * n1h + 255) / 256, std::max(64, 2 * per_frame_blocks)));
        pc_large_pair_kernel<<<dim3(pblocks, nfb), 256, 0, s>>>(a, W.frames, f0, W.cellA, W.cellB, capCells, W.sortA, W.sortB,
                                                               a1, a2, o);
        // every frame is searched by exactly one of the two kernels (pc_large_is_dense, decided on the device)
        const int dper = 8 * 4;
        const int dblocks_pair = (int)std::max<long long>(1, std::min<long long>((3ll * capCells + dper - 1) / dper, std::max(64, 2 * per_frame_blocks)));
        pc_large_pair_dense_kernel<<<dim3(dblocks_pair, nfb), 256, 0, s>>>(a, W.frames, f0, W.cellA, W.cellB, capCells, W.sortA,
                                                                          W.sortB, a1, a2, o);
        const int dblocks = (int)std::max<long long>(1, std::min<long long>((o.cap_c_frame + 255) / 256, per_frame_blocks));
        pc_large_drain_kernel<<<dim3(dblocks, nfb), 256, 0, s>>>(a, W.frames, f0, W.sortA, W.sortB, a1, a2, o);
        pc_large_finish_kernel<<<(nfb + 63) / 64, 64, 0, s>>>(a, W.frames, f0, nfb, o);
        PCL_TRY(cudaGetLastError());
        *launches += 7 + (use_bboxB ? 1 : 0) + (self ? 0 : 2);
    }
    return PC_OK;
}
