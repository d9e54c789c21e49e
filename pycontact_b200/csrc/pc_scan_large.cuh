// pc_scan_large.cuh -- "large frame" path: frames whose heavy atoms do not fit one CTA's shared
// memory (1 M-atom membrane system, 3 M-atom capsid).
//
// Same semantics as pc_scan_small.cuh (loop_trajectory_grid, core/multi_trajectory.py:65-219) with
// the cell list in global memory and many CTAs per frame.  Per sub-batch of FB frames, every kernel
// covers all FB frames (grid.y = frame):
//   L1 bbox     stream A (and B when it is not much larger than A): bounding boxes of the heavy atoms
//   L2 grid     grid over (bboxA (+) c) n (bboxB (+) c), cell edge grown until it fits capCells; zero cells
//   L3 bin      stream A and B ONCE: heavy atoms inside the grid are compacted to (x, y, z, atom word)
//               + cell id, cells counted with global atomics            <- the HBM-bound kernel
//   L4 scan     exclusive scan of the cell counts (one CTA per frame and side)
//   L5 scatter  compacted atoms -> cell order (float4)
//   L6 pairs    one thread per (A atom, stencil z-layer): fused pre-test + exact float32 test; hits are
//               appended to the frame's slice of the contact buffer as (sorted A position, sorted B position)
//   L7 drain    one thread per hit: float64 distance, weight, contact record (in place), H-bond test
//   L8 finish   per-frame counts, totals, overflow status
// L1 and L3 read the packed coordinates through the bulk-copy ring of pc_stream.cuh.
// Each frame owns a fixed slice of the batch's contact / H-bond buffers (capacity / nframes).
#pragma once
#include <string>
#include "pc_common.cuh"
#include "pc_stream.cuh"

struct PcLargeFrame {
    unsigned int bbox[12];            // order-preserving encoded floats: minA, maxA, minB, maxB
    float lo[3];
    float inv;
    int nx, ny, nz, ncell;
    unsigned int nA_in, nB_in;        // atoms inside the grid
    unsigned int ccursor, hcursor;    // records written into this frame's slices
    unsigned int pad[2];              // CTAs of the frame that have finished the bbox pass / the bin pass of selection B
};

struct PcLargeWork {
    int FB = 0, capCells = 0, n1h = 0, n2h = 0, self_mode = 0;
    bool clean = false;                               // every frame slot is in its reset state (the walk kernel leaves it so)
    PcLargeFrame *frames = nullptr;
    uint32_t *cellA = nullptr, *cellB = nullptr;      // FB x (capCells + 1)
    float4 *candA = nullptr, *candB = nullptr;        // FB x n_h   (compacted, unsorted)
    uint32_t *ccellA = nullptr, *ccellB = nullptr;    // FB x n_h   cell id of each compacted atom
    float4 *sortA = nullptr, *sortB = nullptr;        // FB x n_h   (cell order)
    uint8_t *occ = nullptr;                           // FB x capCells: cell has a B atom in its 27-cell neighbourhood
    uint8_t *occ2 = nullptr;                          // FB x capCells: scratch of the separable dilation
    int2 *sortRS = nullptr;                           // FB x n_h  (self contacts: residue, segment in cell order)
    void release() {
        void *ps[] = {frames, cellA, cellB, candA, candB, ccellA, ccellB, sortA, sortB, occ, occ2, sortRS};
        for (void *p : ps) if (p) cudaFree(p);
        frames = nullptr; cellA = cellB = nullptr; candA = candB = sortA = sortB = nullptr;
        ccellA = ccellB = nullptr; occ = occ2 = nullptr; sortRS = nullptr; FB = 0;
    }
    ~PcLargeWork() { release(); }
};

__device__ __forceinline__ unsigned int pc_enc_float(float f) {
    const unsigned int b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float pc_dec_float(unsigned int e) {
    return __uint_as_float((e & 0x80000000u) ? (e & 0x7fffffffu) : ~e);
}

// ---- L0: reset per-frame state ---------------------------------------------------------------
__global__ void pc_large_reset_kernel(PcLargeFrame *fr, int nfb) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nfb) return;
    for (int k = 0; k < 3; k++) {
        fr[f].bbox[k] = 0xffffffffu; fr[f].bbox[3 + k] = 0u;
        fr[f].bbox[6 + k] = 0xffffffffu; fr[f].bbox[9 + k] = 0u;
    }
    fr[f].nA_in = fr[f].nB_in = 0; fr[f].ccursor = fr[f].hcursor = 0; fr[f].ncell = 0;
    fr[f].pad[0] = fr[f].pad[1] = 0;          // arrival counters of the streaming kernels' epilogues
}

__device__ __forceinline__ int pc_large_cell_of(const PcLargeFrame &F, float x, float y, float z) {
    const float ux = (x - F.lo[0]) * F.inv, uy = (y - F.lo[1]) * F.inv, uz = (z - F.lo[2]) * F.inv;
    if (!(ux >= 0.f && uy >= 0.f && uz >= 0.f)) return -1;
    const int cx = (int)ux, cy = (int)uy, cz = (int)uz;
    if (cx >= F.nx || cy >= F.ny || cz >= F.nz) return -1;
    return (cz * F.ny + cy) * F.nx + cx;
}

// ---- L2: grid geometry + zeroed cell counters (one CTA per frame) -----------------------------------
// All threads of a CTA call this for frame fb; thread 0 reads the frame's bounding boxes (through L2: other CTAs
// wrote them with atomics) and fixes the geometry.  Returns the number of cells.
__device__ __forceinline__ int pc_large_grid_frame(const PcScanArgs &a, PcLargeFrame *fr, const int fb, const int capCells,
                                                   const int use_bboxB, uint32_t *cellsA, uint32_t *cellsB, int *s_ncell) {
    if (threadIdx.x == 0) {
        PcLargeFrame &F = fr[fb];
        bool empty = false, bad = false;
        float lo[3], ext[3];
        for (int k = 0; k < 3; k++) {
            const float minA = pc_dec_float(__ldcg(&F.bbox[k])), maxA = pc_dec_float(__ldcg(&F.bbox[3 + k]));
            float l = minA - a.edge, h = maxA + a.edge;
            if (!a.self_mode && use_bboxB) {
                const float minB = pc_dec_float(__ldcg(&F.bbox[6 + k])), maxB = pc_dec_float(__ldcg(&F.bbox[9 + k]));
                l = fmaxf(minA, minB) - a.edge; h = fminf(maxA, maxB) + a.edge;
            }
            if (!(l <= h)) empty = true;
            else if (!isfinite(l) || !isfinite(h)) bad = true;
            lo[k] = l; ext[k] = h - l;
        }
        if (bad && !empty) { atomicOr(&a.counters[PC_CNT_STATUS], PC_ST_BAD_COORDS); empty = true; }
        int nx = 0, ny = 0, nz = 0;
        float inv = 0.f;
        if (!empty) pc_grid_dims(ext[0], ext[1], ext[2], a.edge, (long long)capCells, nx, ny, nz, inv);
        F.lo[0] = lo[0]; F.lo[1] = lo[1]; F.lo[2] = lo[2]; F.inv = inv;
        F.nx = nx; F.ny = ny; F.nz = nz; F.ncell = nx * ny * nz;
        *s_ncell = F.ncell;
    }
    __syncthreads();
    const int ncell = *s_ncell;
    uint32_t *cA = cellsA + (size_t)fb * (capCells + 1);
    uint32_t *cB = a.self_mode ? nullptr : cellsB + (size_t)fb * (capCells + 1);
    for (int c = threadIdx.x; c <= ncell; c += blockDim.x) { cA[c] = 0; if (cB) cB[c] = 0; }
    return ncell;
}

// Dilated B occupancy of one frame by ONE CTA (the streaming bin's last CTA of the frame): separable -- x, then y, then z,
// three loads per cell and pass instead of 27 -- through L2 (the counts were written by other CTAs' atomics, the
// intermediate bytes by this CTA's other threads).
__device__ __forceinline__ void pc_large_occ_frame(const int nx, const int ny, const int nz, const uint32_t *cB, uint8_t *o, uint8_t *t) {
    const int ncell = nx * ny * nz, nxy = nx * ny;
    for (int c = threadIdx.x; c < ncell; c += blockDim.x) {
        const int cx = c % nx;
        unsigned v = __ldcg(cB + c);
        if (cx > 0) v |= __ldcg(cB + c - 1);
        if (cx + 1 < nx) v |= __ldcg(cB + c + 1);
        o[c] = v ? 1 : 0;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < ncell; c += blockDim.x) {
        const int cy = (c / nx) % ny;
        unsigned v = __ldcg(o + c);
        if (cy > 0) v |= __ldcg(o + c - nx);
        if (cy + 1 < ny) v |= __ldcg(o + c + nx);
        t[c] = (uint8_t)v;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < ncell; c += blockDim.x) {
        const int cz = c / nxy;
        unsigned v = __ldcg(t + c);
        if (cz > 0) v |= __ldcg(t + c - nxy);
        if (cz + 1 < nz) v |= __ldcg(t + c + nxy);
        o[c] = (uint8_t)v;
    }
}

// ---- L1 / L3: streaming kernels -----------------------------------------------------------------
#define PC_STREAM_NT 512
#define PC_STREAM_CHUNK 24576            // 2048 atoms per chunk
#ifndef PC_STREAM_STAGES
#define PC_STREAM_STAGES 3
#endif
#define PC_STREAM_OCCW 1024                                  // words of the occupancy bitmap in shared memory (32768 cells)
#define PC_STREAM_MAXCH 512                                  // chunk boundaries whose heavy ordinals are looked up ahead of the loop
#define PC_STREAM_HINFO (PC_STREAM_CHUNK / 12 * 4 + 32)      // a chunk's slice of the heavy-atom list
#define PC_STREAM_SLOT (PC_STREAM_CHUNK + 16 + PC_STREAM_HINFO)
#define PC_STREAM_SMEM (PC_STREAM_STAGES * PC_STREAM_SLOT + PC_STREAM_STAGES * 16 + 16)      // ring + full / empty barriers

struct PcStreamDesc { int cnt, hoff, off0, abase; };          // per ring slot, written by the issuing thread

enum { PC_MODE_BBOX = 0, PC_MODE_BIN = 1 };

// What the LAST CTA of a frame does when it leaves a streaming kernel (one launch less per batch each):
//   PC_EPI_GRID  after the bbox pass of selection A: the frame's grid geometry + zeroed cell counters
//   PC_EPI_OCC   after the bin pass of selection B: the dilated B occupancy that the bin pass of A filters with
enum { PC_EPI_NONE = 0, PC_EPI_GRID = 1, PC_EPI_OCC = 2 };
struct PcStreamEpi {
    int what, capCells, use_bboxB;
    int l2;                           // L2 hint of the coordinate copies: 0 = none, 1 = evict_first, 2 = evict_last
    uint32_t *cellsA, *cellsB;
    uint8_t *occ, *occ2;
};

template <int MODE>
__global__ void __launch_bounds__(PC_STREAM_NT) pc_large_stream_kernel(const PcScanArgs a, PcLargeFrame *fr, int f0, int side,
                                                                       int slot0, int atoms_per_cta, uint32_t *cells,
                                                                       int capCells, float4 *cand, uint32_t *ccell,
                                                                       int nh_alloc, const uint8_t *occ, const PcStreamEpi epi) {
    constexpr int NT = PC_STREAM_NT, CHUNK = PC_STREAM_CHUNK, STAGES = PC_STREAM_STAGES, SLOT = PC_STREAM_SLOT;
    extern __shared__ __align__(16) unsigned char smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + STAGES * SLOT);
    __shared__ PcLargeFrame F;
    __shared__ PcStreamDesc desc[STAGES];
    __shared__ float red[NT / 32][6];
    __shared__ unsigned int s_bad;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fb = blockIdx.y, f = f0 + fb;
    const int n = side == 0 ? a.t.n1 : a.t.n2;
    const int r0 = blockIdx.x * atoms_per_cta;
    if (r0 >= n) return;
    if (tid == 0) {
        F = fr[fb];
        s_bad = 0;
        // full[s]: the producer's expect_tx + the copies' bytes; empty[s]: one arrival per consumer warp
        for (int s = 0; s < STAGES; s++) { pc_mbar_init(bars + s, 1); pc_mbar_init(bars + STAGES + s, NT / 32 - 1); }
        pc_mbar_fence_init();
    }
    __syncthreads();
    if (MODE == PC_MODE_BIN && F.ncell == 0) return;
    const float *frame = side == 0 ? a.xyz1 + (long long)f * a.pitch1 : a.xyz2 + (long long)f * a.pitch2;
    const uint32_t *hmask = side == 0 ? a.t.hmask1 : a.t.hmask2;
    const uint32_t *hpre = side == 0 ? a.t.hpre1 : a.t.hpre2;
    const uint32_t *hinfo = side == 0 ? a.t.hinfo1 : a.t.hinfo2;
    PcSegment seg;
    seg.sb = reinterpret_cast<const unsigned char *>(frame) + 12ll * r0;
    seg.natoms = min(atoms_per_cta, n - r0);
    seg.atom0 = r0;
    const int nch = pc_seg_chunks<CHUNK>(seg);

    // Producer side (thread 0): chunk q = CHUNK bytes of coordinates + the slice of the heavy-atom list
    // (hinfo: local index + packed atom word per heavy atom, in index order) that falls into it, both
    // by bulk copy onto the slot's mbarrier.  Consumers never touch global memory for their input.
    auto heavy_before = [&](int i) -> int {       // heavy ordinal of local atom i
        return (int)(__ldg(hpre + (i >> 5)) + __popc(__ldg(hmask + (i >> 5)) & ((1u << (i & 31)) - 1u)));
    };
    // The heavy ordinals of the chunk boundaries are looked up by many threads at once BEFORE the loop: done by the
    // producer as it went, their two dependent global loads sat in front of every bulk copy it issued, and the whole
    // CTA waited for its warp at each chunk's barrier (1.3 us per 24 KB chunk and CTA).
    __shared__ int s_hb[PC_STREAM_MAXCH + 1];
    for (int q = tid; q <= min(nch, PC_STREAM_MAXCH); q += NT) {
        const int k = q < nch ? pc_chunk_of<CHUNK>(seg, q).k0 : seg.natoms;
        s_hb[q] = heavy_before(r0 + k);
    }
    __syncthreads();
    auto issue = [&](int q) {
        const PcChunk c = pc_chunk_of<CHUNK>(seg, q);
        const int h0 = q <= PC_STREAM_MAXCH ? s_hb[q] : heavy_before(r0 + c.k0);
        const int h1 = q + 1 <= PC_STREAM_MAXCH ? s_hb[q + 1] : heavy_before(r0 + c.k1);
        const uintptr_t hs = (uintptr_t)(hinfo + h0), hs_al = hs & ~(uintptr_t)15;
        uintptr_t he = ((uintptr_t)(hinfo + h1) + 15) & ~(uintptr_t)15;
        if (he <= hs_al) he = hs_al + 16;
        const uint32_t hbytes = (uint32_t)(he - hs_al);
        const int st = q % STAGES;
        PcStreamDesc d;
        d.cnt = h1 - h0; d.hoff = (int)(hs - hs_al); d.off0 = c.off0; d.abase = r0 + c.k0;
        desc[st] = d;
        unsigned char *slot = smem + (size_t)st * SLOT;
        pc_mbar_expect_tx(bars + st, c.bytes + hbytes);          // release: desc is visible to whoever sees the phase flip
        if (epi.l2) pc_bulk_g2s_hint(slot, c.g0, c.bytes, bars + st, pc_l2_policy(epi.l2));
        else pc_bulk_g2s(slot, c.g0, c.bytes, bars + st);
        pc_bulk_g2s(slot + CHUNK + 16, reinterpret_cast<const void *>(hs_al), hbytes, bars + st);
    };

    uint32_t *mycells = nullptr;
    float4 *mycand = nullptr;
    uint32_t *myccell = nullptr;
    unsigned int *cursor = nullptr;
    if (MODE == PC_MODE_BIN) {
        mycells = cells + (size_t)fb * (capCells + 1);
        mycand = cand + (size_t)fb * nh_alloc;
        myccell = ccell + (size_t)fb * nh_alloc;
        cursor = side == 0 ? &fr[fb].nA_in : &fr[fb].nB_in;
    }
    // the frame's dilated B occupancy as a bitmap in shared memory (bin pass of A): the filter's lookup was a global byte
    // load per atom in the consumers' loop
    __shared__ uint32_t s_occ[PC_STREAM_OCCW];
    const bool occ_smem = MODE == PC_MODE_BIN && occ != nullptr && F.ncell <= 32 * PC_STREAM_OCCW;
    if (occ_smem) {
        const uint8_t *ob = occ + (size_t)fb * capCells;
        for (int c0 = warp * 32; c0 < F.ncell; c0 += NT) {
            const int c = c0 + lane;
            const unsigned bal = __ballot_sync(PC_FULL_MASK, c < F.ncell && ob[c] != 0);
            if (lane == 0) s_occ[c0 >> 5] = bal;
        }
        __syncthreads();
    }
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    // conservative box of the grid (BIN mode): cell coordinate u = (x - lo) * inv; x < lo - 0.01 gives u < 0, and
    // x > lo + n * e * (1 + 1e-5) + 0.01 with e = 1 / inv (e * inv = 1 +- 2e-7) gives u > n
    float plo[3] = {0.f, 0.f, 0.f}, phi[3] = {0.f, 0.f, 0.f};
    if (MODE == PC_MODE_BIN) {
        const float e = 1.0f / F.inv;
        const int nn[3] = {F.nx, F.ny, F.nz};
#pragma unroll
        for (int k = 0; k < 3; k++) { plo[k] = F.lo[k] - 0.01f; phi[k] = F.lo[k] + (float)nn[k] * e * 1.00001f + 0.01f; }
    }

    // The last warp only issues copies (lane 0), as far ahead as the ring allows: it refills a slot as soon as every
    // consumer warp has arrived on the slot's empty barrier.  The other warps consume: no CTA barrier per chunk, and the
    // issue of a chunk (64-bit divisions of pc_chunk_of, two bulk copies) is off their critical path -- done by thread 0
    // between two barriers it bounded a CTA at 22 GB/s.
    constexpr int CNT = NT - 32;                      // consumer threads
    if (warp == NT / 32 - 1) {
        if (lane == 0)
            for (int q = 0; q < nch; q++) {
                if (q >= STAGES && !pc_mbar_wait(bars + STAGES + q % STAGES, (unsigned)(q / STAGES - 1) & 1u)) { s_bad = 1; break; }
                issue(q);
            }
    } else
    for (int q = 0; q < nch; q++) {
        const int st = q % STAGES;
        if (!pc_mbar_wait(bars + st, (unsigned)(q / STAGES) & 1u)) { s_bad = 1; break; }      // (a bug, not a data condition)
        const PcStreamDesc d = desc[st];
        const unsigned char *slot = smem + (size_t)st * SLOT;
        const unsigned char *xyz = slot + d.off0;
        const uint32_t *hin = reinterpret_cast<const uint32_t *>(slot + CHUNK + 16 + d.hoff);
        const int cnt = d.cnt, abase = d.abase;
        for (int kk = tid; kk < ((cnt + 31) & ~31); kk += CNT) {
            const bool heavy = kk < cnt;
            uint32_t info = 0;
            float x = 0.f, y = 0.f, z = 0.f;
            if (heavy) {
                info = hin[kk];
                const float *p = reinterpret_cast<const float *>(xyz + 12 * (PC_W_INDEX(info) - abase));
                x = p[0]; y = p[1]; z = p[2];
            }
            if (MODE == PC_MODE_BBOX) {
                if (heavy) {
                    mn[0] = fminf(mn[0], x); mn[1] = fminf(mn[1], y); mn[2] = fminf(mn[2], z);
                    mx[0] = fmaxf(mx[0], x); mx[1] = fmaxf(mx[1], y); mx[2] = fmaxf(mx[2], z);
                }
            } else {
                // most atoms of a large selection are far outside the grid: a box test with a safety margin first (an
                // atom it rejects has a cell coordinate < 0 or >= n for sure), the exact cell only for the few that pass
                const bool maybe = heavy && x >= plo[0] && x <= phi[0] && y >= plo[1] && y <= phi[1] && z >= plo[2] && z <= phi[2];
                if (!__any_sync(PC_FULL_MASK, maybe)) continue;
                int cell = maybe ? pc_large_cell_of(F, x, y, z) : -1;
                // an A atom whose 27-cell neighbourhood holds no B atom cannot have a partner
                if (occ != nullptr && cell >= 0 &&
                    !(occ_smem ? (s_occ[cell >> 5] >> (cell & 31)) & 1u : (unsigned)occ[(size_t)fb * capCells + cell])) cell = -1;
                const unsigned m = __ballot_sync(PC_FULL_MASK, cell >= 0);
                if (m) {
                    unsigned base = 0;
                    const int leader = __ffs(m) - 1;
                    if (lane == leader) base = atomicAdd(cursor, (unsigned)__popc(m));
                    base = __shfl_sync(PC_FULL_MASK, base, leader);
                    if (cell >= 0) {
                        const unsigned p = base + __popc(m & ((1u << lane) - 1u));
                        mycand[p] = make_float4(x, y, z, __uint_as_float(info));
                        myccell[p] = (uint32_t)cell;
                        atomicAdd(&mycells[cell], 1u);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) pc_mbar_arrive(bars + STAGES + st);        // the slot (and its descriptor) may be refilled once every warp is here
    }
    if (MODE == PC_MODE_BBOX) {
#pragma unroll
        for (int k = 0; k < 3; k++) { mn[k] = warp_min(mn[k]); mx[k] = warp_max(mx[k]); }
        if (lane == 0) for (int k = 0; k < 3; k++) { red[warp][k] = mn[k]; red[warp][3 + k] = mx[k]; }
        __syncthreads();
        if (tid < 6) {
            float v = red[0][tid];
            for (int w = 1; w < NT / 32; w++) v = tid < 3 ? fminf(v, red[w][tid]) : fmaxf(v, red[w][tid]);
            // a selection without heavy atoms leaves +-inf here; the grid kernel treats that as empty
            if (tid < 3) atomicMin(&fr[fb].bbox[slot0 + tid], pc_enc_float(v));
            else atomicMax(&fr[fb].bbox[slot0 + tid], pc_enc_float(v));
        }
    }
    if (tid == 0 && s_bad) atomicOr(&a.counters[PC_CNT_STATUS], (unsigned long long)PC_ST_BAD_COORDS);
    if (epi.what != PC_EPI_NONE) {
        // every CTA of the frame arrives here (pc_stream_geom launches no CTA without atoms; frames without a grid left
        // the bin pass above and need no occupancy); the last one does the epilogue
        __shared__ int s_last, s_ncell;
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = atomicAdd(&fr[fb].pad[epi.what - 1], 1u) == gridDim.x - 1u;
        __syncthreads();
        if (s_last) {
            __threadfence();
            if (epi.what == PC_EPI_GRID) {
                pc_large_grid_frame(a, fr, fb, epi.capCells, epi.use_bboxB, epi.cellsA, epi.cellsB, &s_ncell);
            } else {
                pc_large_occ_frame(F.nx, F.ny, F.nz, epi.cellsB + (size_t)fb * (epi.capCells + 1),
                                   epi.occ + (size_t)fb * epi.capCells, epi.occ2 + (size_t)fb * epi.capCells);
            }
        }
    }
}


// ---- L2: grid geometry + zeroed cell counters (one CTA per frame; pc_large_grid_frame above) ---------------
__global__ void __launch_bounds__(256) pc_large_grid_kernel(const PcScanArgs a, PcLargeFrame *fr, int capCells, int use_bboxB,
                                                            uint32_t *cellsA, uint32_t *cellsB) {
    __shared__ int s_ncell;
    pc_large_grid_frame(a, fr, blockIdx.x, capCells, use_bboxB, cellsA, cellsB, &s_ncell);
}

// ---- L3b: dilated B occupancy (after B has been binned, before A is) -------------------------------------
__global__ void __launch_bounds__(256) pc_large_occ_kernel(const PcLargeFrame *fr, const uint32_t *cellsB, int capCells,
                                                           uint8_t *occ) {
    const int fb = blockIdx.y;
    const int nx = fr[fb].nx, ny = fr[fb].ny, nz = fr[fb].nz, ncell = fr[fb].ncell;
    const uint32_t *cB = cellsB + (size_t)fb * (capCells + 1);
    uint8_t *o = occ + (size_t)fb * capCells;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ncell; c += gridDim.x * blockDim.x) {
        const int cx = c % nx, cy = (c / nx) % ny, cz = c / (nx * ny);
        bool any = false;
        for (int z = max(cz - 1, 0); z <= min(cz + 1, nz - 1) && !any; z++)
            for (int y = max(cy - 1, 0); y <= min(cy + 1, ny - 1) && !any; y++)
                for (int x = max(cx - 1, 0); x <= min(cx + 1, nx - 1); x++)
                    if (cB[(z * ny + y) * nx + x]) { any = true; break; }
        o[c] = any ? 1 : 0;
    }
}

// ---- L4: exclusive scan of cell counts (one CTA per frame and side) ---------------------------
__global__ void __launch_bounds__(1024) pc_large_scan_kernel(PcLargeFrame *fr, uint32_t *cellsA, uint32_t *cellsB,
                                                             int capCells) {
    const int fb = blockIdx.x;
    uint32_t *arr = (blockIdx.y == 0 ? cellsA : cellsB) + (size_t)fb * (capCells + 1);
    const int n = fr[fb].ncell;
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    constexpr int PER = 4;
    for (int base = 0; base < n; base += 1024 * PER) {
        uint32_t v[PER];
        uint32_t s = 0;
#pragma unroll
        for (int q = 0; q < PER; q++) {
            const int i = base + tid * PER + q;
            v[q] = i < n ? arr[i] : 0;
            s += v[q];
        }
        const uint32_t incl = warp_incl_scan(s, lane);
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const uint32_t w = wsum[lane];
            const uint32_t wi = warp_incl_scan(w, lane);
            wsum[lane] = wi - w;
        }
        __syncthreads();
        uint32_t off = carry + wsum[warp] + incl - s;
#pragma unroll
        for (int q = 0; q < PER; q++) {
            const int i = base + tid * PER + q;
            if (i < n) arr[i] = off;
            off += v[q];
        }
        __syncthreads();
        if (tid == 1023) carry = off;
        __syncthreads();
    }
    if (tid == 0) arr[n] = carry;
}

// ---- L5: scatter compacted atoms into cell order; afterwards cells[c] == END of cell c ---------
__global__ void __launch_bounds__(256) pc_large_scatter_kernel(PcLargeFrame *fr, int side, uint32_t *cells, int capCells,
                                                               const float4 *cand, const uint32_t *ccell, float4 *sorted,
                                                               int nh_alloc, int2 *sorted_rs = nullptr, const int *lres = nullptr,
                                                               const int *lseg = nullptr) {
    const int fb = blockIdx.y;
    const unsigned n = side == 0 ? fr[fb].nA_in : fr[fb].nB_in;
    uint32_t *mycells = cells + (size_t)fb * (capCells + 1);
    const float4 *mycand = cand + (size_t)fb * nh_alloc;
    const uint32_t *myccell = ccell + (size_t)fb * nh_alloc;
    float4 *mysorted = sorted + (size_t)fb * nh_alloc;
    for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const float4 v = mycand[p];
        const unsigned dst = atomicAdd(&mycells[myccell[p]], 1u);
        mysorted[dst] = v;
        // (residue, segment) of the atoms in the same order: the group kernel's self-interaction filter reads them
        // coalesced, next to the coordinates
        if (sorted_rs) {
            const int l = PC_W_INDEX(__float_as_uint(v.w));
            sorted_rs[(size_t)fb * nh_alloc + dst] = make_int2(__ldg(lres + l), __ldg(lseg + l));
        }
    }
}

// ---- L6: pair search ---------------------------------------------------------------------------------
struct PcLargeOut {
    long long cap_c_frame, cap_h_frame;   // records per frame slice
};

// Which pair kernel searches a frame: with a few B atoms per cell on average every B atom is a candidate
// of several A atoms, and loading it once per A CELL (warp per cell, lanes over B) beats loading it once
// per A ATOM (thread per atom).
__device__ __forceinline__ bool pc_large_is_dense(const PcLargeFrame &F, bool self, int mode = 0) {
    const unsigned nB = self ? F.nA_in : F.nB_in;
    if (mode) return F.ncell > 0 && mode == 2;
    return F.ncell > 0 && (unsigned long long)nB * 2ull >= 5ull * (unsigned long long)F.ncell;      // >= 2.5 per cell
}

// Hits are staged in per-WARP queues in shared memory and appended to the frame's slice with ONE global
// atomic per 32 work items (a CTA-wide queue cost a barrier per tile, 31 % of this kernel's stall samples on c3): in dense systems (a hit every ~6 candidates) a global atomic per
// hit would make every hit an L2 round trip inside the candidate loop.
#define PC_PAIR_Q 4096
#ifndef PC_PAIR_UNROLL
#define PC_PAIR_UNROLL 4
#endif
constexpr int pc_pair_unroll = PC_PAIR_UNROLL;
__global__ void __launch_bounds__(256) pc_large_pair_kernel(const PcScanArgs a, PcLargeFrame *fr, int f0,
                                                            const uint32_t *cellsA, const uint32_t *cellsB, int capCells,
                                                            const float4 *sortA, const float4 *sortB, int n1h_alloc,
                                                            int n2h_alloc, PcLargeOut o) {
    __shared__ PcLargeFrame F;
    __shared__ uint2 q[8][PC_PAIR_Q / 8];          // one queue per warp: no CTA barrier inside the tile loop
    __shared__ unsigned int qn[8];
    const int fb = blockIdx.y, f = f0 + fb;
    if (threadIdx.x == 0) F = fr[fb];
    __syncthreads();
    const bool self = a.self_mode != 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint2 *wq = q[warp];
    unsigned long long evals = 0;
    unsigned int status = 0;
    const float sqdist = a.sqdist, sqdist_hi = a.sqdist * 1.00000095367431640625f;   // 1 + 2^-20
    const uint32_t *cB = (self ? cellsA : cellsB) + (size_t)fb * (capCells + 1);
    const float4 *sA = sortA + (size_t)fb * n1h_alloc;
    const float4 *sB = self ? sA : sortB + (size_t)fb * n2h_alloc;
    pc_contact *slice = a.contacts + (long long)f * o.cap_c_frame;
    const int nx = F.nx, ny = F.ny, nz = F.nz;
    const long long items = 3ll * F.nA_in;
    if (pc_large_is_dense(F, self, a.dense_mode)) return;          // dense frames are searched by pc_large_pair_dense_kernel
    for (long long tile = (blockIdx.x * 8ll + warp) * 32; tile < items; tile += (long long)gridDim.x * 256) {
        if (lane == 0) qn[warp] = 0;
        __syncwarp();
        const long long w = tile + lane;
        if (w < items) {
            // layer-major items: a warp holds 32 neighbouring atoms (A is cell-sorted) on the same layer,
            // so its lanes walk the same or adjacent B runs (few distinct cache lines per load)
            const int layer = (int)(w / F.nA_in), ia = (int)(w - (long long)layer * F.nA_in);
            const float4 pa = __ldg(sA + ia);
            const int cx = (int)((pa.x - F.lo[0]) * F.inv), cy = (int)((pa.y - F.lo[1]) * F.inv),
                      cz = (int)((pa.z - F.lo[2]) * F.inv);
            const int z = cz + layer - 1;
            if (z >= 0 && z < nz) {
                const int xlo = max(cx - 1, 0), xhi = min(cx + 1, nx - 1);
                // the bounds of the layer's three B runs first (six independent loads in flight)
                int b0[3], b1[3];
#pragma unroll
                for (int dy = 0; dy < 3; dy++) {
                    const int y = cy + dy - 1;
                    if (y < 0 || y >= ny) { b0[dy] = 0; b1[dy] = 0; continue; }
                    const int row = (z * ny + y) * nx;
                    b0[dy] = (row + xlo) ? (int)__ldg(cB + row + xlo - 1) : 0;
                    b1[dy] = (int)__ldg(cB + row + xhi);
                }
                int ares = 0, aseg = 0;
                if (self) {
                    const int li = PC_W_INDEX(__float_as_uint(pa.w));
                    ares = __ldg(a.t.lres1 + li); aseg = __ldg(a.t.lseg1 + li);
                }
#pragma unroll
                for (int dy = 0; dy < 3; dy++) {
                    evals += (unsigned)(b1[dy] - b0[dy]);
                    // 32 candidates at a time into a hit mask: the candidate loop has no side effects (the
                    // compiler can batch its loads), and the queue is touched once per chunk, not per hit
                    for (int jb = b0[dy]; jb < b1[dy]; jb += 32) {
                        const int cnt = min(32, b1[dy] - jb);
                        unsigned m = 0;
#pragma unroll pc_pair_unroll
                        for (int k = 0; k < cnt; k++) {
                            const float4 pb = __ldg(sB + jb + k);
                            m |= (unsigned)pair_within(pa.x, pa.y, pa.z, pb.x, pb.y, pb.z, sqdist, sqdist_hi) << k;
                        }
                        if (self && m)
                            m = pc_self_filter(m, ares, aseg, a.t.lres1, a.t.lseg1,
                                               [&](int k) { return __float_as_uint(__ldg(&sB[jb + k].w)); });
                        if (!m) continue;
                        unsigned p = atomicAdd(&qn[warp], (unsigned)__popc(m));
                        while (m) {
                            const int j = jb + __ffs(m) - 1;
                            m &= m - 1;
                            if (p < PC_PAIR_Q / 8) {
                                wq[p] = make_uint2((unsigned)ia, (unsigned)j);
                            } else {
                                // queue full (very long cutoffs): this hit goes straight to the slice
                                const unsigned slot = atomicAdd(&fr[fb].ccursor, 1u);
                                if ((long long)slot < o.cap_c_frame) {
                                    pc_contact rec;
                                    rec.i = ia; rec.j = j; rec.distance = 0.f; rec.weight = 0.f;
                                    slice[slot] = rec;
                                } else {
                                    status |= (unsigned)PC_ST_OUT_OVERFLOW;
                                }
                            }
                            p++;
                        }
                    }
                }
            }
        }
        __syncwarp();
        const unsigned n = min(qn[warp], (unsigned)(PC_PAIR_Q / 8));
        unsigned qbase = 0;
        if (lane == 0 && n) qbase = atomicAdd(&fr[fb].ccursor, n);
        qbase = __shfl_sync(PC_FULL_MASK, qbase, 0);
        for (unsigned k = lane; k < n; k += 32) {
            const long long slot = (long long)qbase + k;
            if (slot < o.cap_c_frame) {
                pc_contact rec;
                rec.i = (int)wq[k].x; rec.j = (int)wq[k].y; rec.distance = 0.f; rec.weight = 0.f;
                slice[slot] = rec;
            } else {
                status |= (unsigned)PC_ST_OUT_OVERFLOW;
            }
        }
        __syncwarp();
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) evals += __shfl_xor_sync(PC_FULL_MASK, evals, s);
    if (lane == 0 && evals) atomicAdd(&a.counters[PC_CNT_EVALS], evals);
    status = __reduce_or_sync(PC_FULL_MASK, status);
    if (lane == 0 && status) atomicOr(&a.counters[PC_CNT_STATUS], (unsigned long long)status);
}

// Dense variant: one warp per (A cell, stencil z-layer).  The cell's A atoms sit in lane registers; the
// layer's three B runs are walked as ONE list, 32 atoms at a time (each B atom is loaded once per cell),
// and every lane tests its B atom against all A atoms of the cell (broadcast by shuffle) into a hit
// mask.  Hits go through a per-WARP queue in shared memory that the warp itself appends to the frame's
// slice (one global atomic per flush), so there is no CTA barrier and no per-hit global atomic however
// dense the system is.
#define PC_DENSE_WQ 512
__device__ __forceinline__ void pc_dense_flush(PcLargeFrame *frp, pc_contact *slice, const PcLargeOut &o, const uint2 *wq,
                                               unsigned n, int lane, unsigned int &status) {
    if (n == 0) return;
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(&frp->ccursor, n);
    base = __shfl_sync(PC_FULL_MASK, base, 0);
    for (unsigned k = lane; k < n; k += 32) {
        const long long slot = (long long)base + k;
        if (slot < o.cap_c_frame) {
            pc_contact rec;
            rec.i = (int)wq[k].x; rec.j = (int)wq[k].y; rec.distance = 0.f; rec.weight = 0.f;
            slice[slot] = rec;
        } else {
            status |= (unsigned)PC_ST_OUT_OVERFLOW;
        }
    }
    __syncwarp();
}

__global__ void __launch_bounds__(256) pc_large_pair_dense_kernel(const PcScanArgs a, PcLargeFrame *fr, int f0,
                                                                  const uint32_t *cellsA, const uint32_t *cellsB, int capCells,
                                                                  const float4 *sortA, const float4 *sortB, int n1h_alloc,
                                                                  int n2h_alloc, PcLargeOut o) {
    __shared__ PcLargeFrame F;
    __shared__ uint2 q[8][PC_DENSE_WQ];
    const int fb = blockIdx.y, f = f0 + fb;
    if (threadIdx.x == 0) F = fr[fb];
    __syncthreads();
    const bool self = a.self_mode != 0;
    if (!pc_large_is_dense(F, self, a.dense_mode)) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    uint2 *wq = q[warp];
    unsigned wn = 0;                      // entries in this warp's queue (same value in every lane)
    unsigned long long evals = 0;
    unsigned int status = 0;
    const float sqdist = a.sqdist, sqdist_hi = a.sqdist * 1.00000095367431640625f;   // 1 + 2^-20
    const uint32_t *cA = cellsA + (size_t)fb * (capCells + 1);
    const uint32_t *cB = (self ? cellsA : cellsB) + (size_t)fb * (capCells + 1);
    const float4 *sA = sortA + (size_t)fb * n1h_alloc;
    const float4 *sB = self ? sA : sortB + (size_t)fb * n2h_alloc;
    pc_contact *slice = a.contacts + (long long)f * o.cap_c_frame;
    const int nx = F.nx, ny = F.ny, nz = F.nz;
    const long long items = 3ll * F.ncell;
    for (long long w = (long long)blockIdx.x * nwarp + warp; w < items; w += (long long)gridDim.x * nwarp) {
        const int layer = (int)(w / F.ncell), c = (int)(w - (long long)layer * F.ncell);
        const int a0 = c ? (int)__ldg(cA + c - 1) : 0, a1 = (int)__ldg(cA + c);
        if (a0 == a1) continue;
        const int cx = c % nx, cy = (c / nx) % ny, cz = c / (nx * ny);
        const int z = cz + layer - 1;
        if (z < 0 || z >= nz) continue;
        const int xlo = max(cx - 1, 0), xhi = min(cx + 1, nx - 1);
        // the layer's three runs as one list: run r covers list positions [st[r], st[r+1])
        int rb[3], st[4];
        st[0] = 0;
#pragma unroll
        for (int dy = 0; dy < 3; dy++) {
            const int y = cy + dy - 1;
            int b0 = 0, b1 = 0;
            if (y >= 0 && y < ny) {
                const int row = (z * ny + y) * nx;
                b0 = (row + xlo) ? (int)__ldg(cB + row + xlo - 1) : 0;
                b1 = (int)__ldg(cB + row + xhi);
            }
            rb[dy] = b0; st[dy + 1] = st[dy] + (b1 - b0);
        }
        const int nb = st[3];
        if (nb == 0) continue;
        for (int ab = a0; ab < a1; ab += 32) {
            const int na = min(32, a1 - ab);
            float4 pa = make_float4(0.f, 0.f, 0.f, 0.f);
            int ares = 0, aseg = 0;
            if (lane < na) {
                pa = __ldg(sA + ab + lane);
                if (self) {
                    const int li = PC_W_INDEX(__float_as_uint(pa.w));
                    ares = __ldg(a.t.lres1 + li); aseg = __ldg(a.t.lseg1 + li);
                }
            }
            if (lane == 0) evals += (unsigned long long)na * (unsigned)nb;
            for (int t0 = 0; t0 < nb; t0 += 32) {
                const int t = t0 + lane;
                const bool valid = t < nb;
                const int r = (t >= st[1]) + (t >= st[2]);
                const int j = rb[r] + (t - st[r]);
                float4 pb = make_float4(0.f, 0.f, 0.f, 0.f);
                int bres = 0, bseg = 0;
                if (valid) {
                    pb = __ldg(sB + j);
                    if (self) {
                        const int lj = PC_W_INDEX(__float_as_uint(pb.w));
                        bres = __ldg(a.t.lres1 + lj); bseg = __ldg(a.t.lseg1 + lj);
                    }
                }
                unsigned hm = 0;          // bit ia: A atom ab + ia is within the cutoff of this lane's B atom
                for (int ia = 0; ia < na; ia++) {
                    const float px = __shfl_sync(PC_FULL_MASK, pa.x, ia);
                    const float py = __shfl_sync(PC_FULL_MASK, pa.y, ia);
                    const float pz = __shfl_sync(PC_FULL_MASK, pa.z, ia);
                    bool hit = valid && pair_within(px, py, pz, pb.x, pb.y, pb.z, sqdist, sqdist_hi);
                    if (self) {
                        // (resid1 - resid2) < 5 and same segment -> dropped (core/multi_trajectory.py:93-95)
                        const int r1 = __shfl_sync(PC_FULL_MASK, ares, ia), s1 = __shfl_sync(PC_FULL_MASK, aseg, ia);
                        if (hit && (r1 - bres) < 5 && s1 == bseg) hit = false;
                    }
                    hm |= (unsigned)hit << ia;
                }
                const unsigned cnt = (unsigned)__popc(hm);
                const unsigned incl = warp_incl_scan(cnt, lane);
                const unsigned total = __shfl_sync(PC_FULL_MASK, incl, 31);
                if (total == 0) continue;
                if (wn + total > PC_DENSE_WQ) { pc_dense_flush(&fr[fb], slice, o, wq, wn, lane, status); wn = 0; }
                if (total > PC_DENSE_WQ) {
                    // more hits in one chunk than the queue holds: straight to the slice
                    unsigned base = 0;
                    if (lane == 31) base = atomicAdd(&fr[fb].ccursor, total);
                    base = __shfl_sync(PC_FULL_MASK, base, 31);
                    long long p = (long long)base + incl - cnt;
                    while (hm) {
                        const int ia = __ffs(hm) - 1;
                        hm &= hm - 1;
                        if (p < o.cap_c_frame) {
                            pc_contact rec;
                            rec.i = ab + ia; rec.j = j; rec.distance = 0.f; rec.weight = 0.f;
                            slice[p] = rec;
                        } else {
                            status |= (unsigned)PC_ST_OUT_OVERFLOW;
                        }
                        p++;
                    }
                } else {
                    unsigned p = wn + incl - cnt;
                    while (hm) {
                        const int ia = __ffs(hm) - 1;
                        hm &= hm - 1;
                        wq[p++] = make_uint2((unsigned)(ab + ia), (unsigned)j);
                    }
                    wn += total;
                    __syncwarp();
                }
            }
        }
    }
    pc_dense_flush(&fr[fb], slice, o, wq, wn, lane, status);
    evals = __shfl_sync(PC_FULL_MASK, evals, 0);
    if (lane == 0 && evals) atomicAdd(&a.counters[PC_CNT_EVALS], evals);
    status = __reduce_or_sync(PC_FULL_MASK, status);
    if (lane == 0 && status) atomicOr(&a.counters[PC_CNT_STATUS], (unsigned long long)status);
}

// ---- L7: drain: hits -> contact records (in place) + H-bond records ------------------------------------
__global__ void __launch_bounds__(256) pc_large_drain_kernel(const PcScanArgs a, PcLargeFrame *fr, int f0, const float4 *sortA,
                                                             const float4 *sortB, int n1h_alloc, int n2h_alloc,
                                                             PcLargeOut o) {
    const int fb = blockIdx.y, f = f0 + fb;
    const bool self = a.self_mode != 0;
    const long long nhit = min((long long)fr[fb].ccursor, o.cap_c_frame);
    const float4 *sA = sortA + (size_t)fb * n1h_alloc;
    const float4 *sB = self ? sA : sortB + (size_t)fb * n2h_alloc;
    pc_contact *slice = a.contacts + (long long)f * o.cap_c_frame;
    pc_hbond *hslice = a.hbonds + (long long)f * o.cap_h_frame;
    const float *fr1 = a.xyz1 + (long long)f * a.pitch1;
    const float *fr2 = self ? fr1 : a.xyz2 + (long long)f * a.pitch2;
    unsigned int status = 0;
    for (long long h = blockIdx.x * (long long)blockDim.x + threadIdx.x; h < nhit; h += (long long)gridDim.x * blockDim.x) {
        pc_contact rec = slice[h];
        const float4 pa = __ldg(sA + rec.i), pb = __ldg(sB + rec.j);
        const unsigned wa = __float_as_uint(pa.w), wb = __float_as_uint(pb.w);
        const int li = PC_W_INDEX(wa), lj = PC_W_INDEX(wb);
        const double ax = pa.x, ay = pa.y, az = pa.z, bx = pb.x, by = pb.y, bz = pb.z;
        const double dx = ax - bx, dy = ay - by, dz = az - bz;
        const double dist = sqrt(dx * dx + dy * dy + dz * dz);
        rec.i = li; rec.j = lj; rec.distance = (float)dist; rec.weight = contact_weight_f32(dist);
        slice[h] = rec;
        if (wa & wb & PC_W_HBCLASS) {
            for (int side = 0; side < 2; side++) {
                const int *hptr = side == 0 ? a.t.lhptr1 : a.t.lhptr2;
                const int *hidx = side == 0 ? a.t.lhidx1 : a.t.lhidx2;
                const int heavy = side == 0 ? li : lj;
                const int p0 = __ldg(hptr + heavy), p1 = __ldg(hptr + heavy + 1);
                for (int p = p0; p < p1; p++) {
                    const int hl = __ldg(hidx + p);
                    if (hl < 0) { status |= (unsigned)PC_ST_MISSING_H; continue; }
                    float hxf, hyf, hzf;
                    load_xyz(side == 0 ? fr1 : fr2, hl, a.stride, hxf, hyf, hzf);
                    double d = 0, ang = 0;
                    const bool ok = side == 0
                        ? hbond_test(hxf, hyf, hzf, ax, ay, az, bx, by, bz, a.hb_cutoff, a.hb_angle, d, ang)
                        : hbond_test(hxf, hyf, hzf, bx, by, bz, ax, ay, az, a.hb_cutoff, a.hb_angle, d, ang);
                    if (ok) {
                        const unsigned hq = atomicAdd(&fr[fb].hcursor, 1u);
                        if ((long long)hq < o.cap_h_frame) {
                            pc_hbond r;
                            r.i = li; r.j = lj; r.hyd = (uint32_t)hl | (side ? 0x80000000u : 0u);
                            r.distance = (float)d; r.angle = (float)ang;
                            hslice[hq] = r;
                        } else {
                            status |= (unsigned)PC_ST_OUT_OVERFLOW;
                        }
                    }
                }
            }
        }
    }
    status = __reduce_or_sync(PC_FULL_MASK, status);
    if ((threadIdx.x & 31) == 0 && status) atomicOr(&a.counters[PC_CNT_STATUS], (unsigned long long)status);
}

// ---- L8: per-frame counts and totals -----------------------------------------------------------------
__global__ void pc_large_finish_kernel(const PcScanArgs a, PcLargeFrame *fr, int f0, int nfb, PcLargeOut o) {
    const int fb = blockIdx.x * blockDim.x + threadIdx.x;
    if (fb >= nfb) return;
    const int f = f0 + fb;
    const unsigned nc = (unsigned)min((long long)fr[fb].ccursor, o.cap_c_frame);
    const unsigned nh = (unsigned)min((long long)fr[fb].hcursor, o.cap_h_frame);
    a.frame_cstart[f] = (uint32_t)((long long)f * o.cap_c_frame); a.frame_ccount[f] = nc;
    a.frame_hstart[f] = (uint32_t)((long long)f * o.cap_h_frame); a.frame_hcount[f] = nh;
    // the totals report what WOULD be needed so the host can re-reserve after an overflow
    atomicAdd(&a.counters[PC_CNT_CONTACTS], (unsigned long long)fr[fb].ccursor);
    atomicAdd(&a.counters[PC_CNT_HBONDS], (unsigned long long)fr[fb].hcursor);
    atomicMax(&a.counters[6], (unsigned long long)fr[fb].ccursor);   // max contacts of any frame
    atomicMax(&a.counters[7], (unsigned long long)fr[fb].hcursor);
}

