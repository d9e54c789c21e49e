// pc_scan_large.cuh -- "large frame" path: frames whose heavy atoms do not fit one CTA's shared
// memory (1 M-atom membrane system, 3 M-atom capsid).
//
// Same semantics as pc_scan_small.cuh (loop_trajectory_grid, core/multi_trajectory.py:65-219) with
// the cell list in global memory and many CTAs per frame.  Per sub-batch of FB frames:
//   L1 bbox     bounding box of A's heavy atoms (and of B's when B is not much larger than A)
//   L2 grid     grid over (bboxA (+) c) n (bboxB (+) c), cell edge grown until it fits capCells
//   L3 bin      ONE pass over the raw coordinates: atoms inside the grid are compacted
//               (x, y, z, heavy ordinal) + cell id, cells counted with global atomics
//   L4 scan     exclusive scan of the cell counts (one CTA per frame and side)
//   L5 scatter  compacted atoms -> cell order (float4, coalesced for the pair kernel)
//   L6 pairs    persistent warps pull (frame, A-cell) work items; lanes hold B atoms of an x-run,
//               A atoms are broadcast by shuffle; exact float32 test; hits go through a per-warp
//               queue and are drained 32 at a time (float64 distance/weight, H-bond test, records)
//   L7 finish   per-frame counts, totals, overflow status
// Each frame owns a fixed slice of the batch's contact / H-bond buffers (capacity / nframes).
#pragma once
#include <string>
#include "pc_common.cuh"

struct PcLargeFrame {
    unsigned int bbox[12];            // order-preserving encoded floats: minA, maxA, minB, maxB
    float lo[3];
    float inv;
    int nx, ny, nz, ncell;
    unsigned int nA_in, nB_in;        // atoms inside the grid
    unsigned int ccursor, hcursor;    // records written into this frame's slices
    unsigned int item_base;           // first work item (A cell) of this frame in the sub-batch
    unsigned int pad;
};

struct PcLargeWork {
    int FB = 0, capCells = 0, n1h = 0, n2h = 0, self_mode = 0;
    PcLargeFrame *frames = nullptr;
    uint32_t *cellA = nullptr, *cellB = nullptr;      // FB x (capCells + 1)
    float4 *candA = nullptr, *candB = nullptr;        // FB x n_h   (compacted, unsorted)
    uint32_t *ccellA = nullptr, *ccellB = nullptr;    // FB x n_h   cell id of each compacted atom
    float4 *sortA = nullptr, *sortB = nullptr;        // FB x n_h   (cell order)
    unsigned int *work = nullptr;                     // [0] work counter, [1] total items
    void release() {
        void *ps[] = {frames, cellA, cellB, candA, candB, ccellA, ccellB, sortA, sortB, work};
        for (void *p : ps) if (p) cudaFree(p);
        frames = nullptr; cellA = cellB = nullptr; candA = candB = sortA = sortB = nullptr;
        ccellA = ccellB = nullptr; work = nullptr; FB = 0;
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
__global__ void pc_large_reset_kernel(PcLargeFrame *fr, int nfb, unsigned int *work) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f == 0) { work[0] = 0; work[1] = 0; }
    if (f >= nfb) return;
    for (int k = 0; k < 3; k++) {
        fr[f].bbox[k] = 0xffffffffu; fr[f].bbox[3 + k] = 0u;
        fr[f].bbox[6 + k] = 0xffffffffu; fr[f].bbox[9 + k] = 0u;
    }
    fr[f].nA_in = fr[f].nB_in = 0; fr[f].ccursor = fr[f].hcursor = 0; fr[f].ncell = 0;
}

// ---- L1: bounding boxes ----------------------------------------------------------------------
__global__ void __launch_bounds__(256) pc_large_bbox_kernel(const PcScanArgs a, PcLargeFrame *fr, int f0, int side,
                                                            int slot0) {
    const int fb = blockIdx.y;
    const int f = f0 + fb;
    const float *frame = side == 0 ? a.xyz1 + (long long)f * a.pitch1 : a.xyz2 + (long long)f * a.pitch2;
    const int *heavy = side == 0 ? a.t.heavy1 : a.t.heavy2;
    const int nh = side == 0 ? a.t.n1h : a.t.n2h;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nh; k += gridDim.x * blockDim.x) {
        float x, y, z;
        load_xyz(frame, __ldg(heavy + k), a.stride, x, y, z);
        mn[0] = fminf(mn[0], x); mn[1] = fminf(mn[1], y); mn[2] = fminf(mn[2], z);
        mx[0] = fmaxf(mx[0], x); mx[1] = fmaxf(mx[1], y); mx[2] = fmaxf(mx[2], z);
    }
    __shared__ float red[8][6];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 3; k++) { mn[k] = warp_min(mn[k]); mx[k] = warp_max(mx[k]); }
    if (lane == 0) for (int k = 0; k < 3; k++) { red[warp][k] = mn[k]; red[warp][3 + k] = mx[k]; }
    __syncthreads();
    if (threadIdx.x < 6) {
        const int k = threadIdx.x;
        float v = red[0][k];
        for (int w = 1; w < 8; w++) v = k < 3 ? fminf(v, red[w][k]) : fmaxf(v, red[w][k]);
        if (k < 3) atomicMin(&fr[fb].bbox[slot0 + k], pc_enc_float(v));
        else atomicMax(&fr[fb].bbox[slot0 + k], pc_enc_float(v));
    }
}

// ---- L2: grid geometry -----------------------------------------------------------------------
__global__ void pc_large_grid_kernel(const PcScanArgs a, PcLargeFrame *fr, int nfb, int capCells, int use_bboxB,
                                     unsigned long long *counters, unsigned int *work) {
    // one thread: the per-frame work-item prefix needs a serial pass anyway (nfb is small)
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    unsigned int items = 0;
    for (int f = 0; f < nfb; f++) {
        PcLargeFrame &F = fr[f];
        bool empty = false, bad = false;
        float lo[3], ext[3];
        for (int k = 0; k < 3; k++) {
            const float minA = pc_dec_float(F.bbox[k]), maxA = pc_dec_float(F.bbox[3 + k]);
            float l = minA - a.edge, h = maxA + a.edge;
            if (a.self_mode) { l = minA; h = maxA; }   // A == B: the joint box is the box itself
            else if (use_bboxB) {
                const float minB = pc_dec_float(F.bbox[6 + k]), maxB = pc_dec_float(F.bbox[9 + k]);
                l = fmaxf(minA, minB) - a.edge; h = fminf(maxA, maxB) + a.edge;
            }
            if (!(l <= h)) empty = true;
            if (!isfinite(l) || !isfinite(h)) bad = true;
            lo[k] = l; ext[k] = h - l;
        }
        if (bad && !empty) { atomicOr(&counters[PC_CNT_STATUS], PC_ST_BAD_COORDS); empty = true; }
        int nx = 0, ny = 0, nz = 0;
        float inv = 0.f;
        if (!empty) {
            float e = a.edge;
            for (;;) {
                inv = 1.0f / e;
                const long long lx = (long long)(ext[0] * inv) + 1, ly = (long long)(ext[1] * inv) + 1,
                                lz = (long long)(ext[2] * inv) + 1;
                if (lx * ly * lz <= (long long)capCells) { nx = (int)lx; ny = (int)ly; nz = (int)lz; break; }
                e *= 1.26f;
            }
        }
        F.lo[0] = lo[0]; F.lo[1] = lo[1]; F.lo[2] = lo[2]; F.inv = inv;
        F.nx = nx; F.ny = ny; F.nz = nz; F.ncell = nx * ny * nz;
        F.item_base = items;
        items += (unsigned)F.ncell;
    }
    work[0] = 0; work[1] = items;
}

__device__ __forceinline__ int pc_large_cell_of(const PcLargeFrame &F, float x, float y, float z) {
    const float ux = (x - F.lo[0]) * F.inv, uy = (y - F.lo[1]) * F.inv, uz = (z - F.lo[2]) * F.inv;
    if (!(ux >= 0.f && uy >= 0.f && uz >= 0.f)) return -1;
    const int cx = (int)ux, cy = (int)uy, cz = (int)uz;
    if (cx >= F.nx || cy >= F.ny || cz >= F.nz) return -1;
    return (cz * F.ny + cy) * F.nx + cx;
}

// ---- L3: bin (single pass over the raw coordinates) ----------------------------------------------
__global__ void __launch_bounds__(256) pc_large_bin_kernel(const PcScanArgs a, PcLargeFrame *fr, int f0, int side,
                                                           uint32_t *cells, int capCells, float4 *cand, uint32_t *ccell,
                                                           int nh_alloc) {
    const int fb = blockIdx.y;
    const int f = f0 + fb;
    __shared__ PcLargeFrame F;
    if (threadIdx.x == 0) F = fr[fb];
    __syncthreads();
    if (F.ncell == 0) return;
    const float *frame = side == 0 ? a.xyz1 + (long long)f * a.pitch1 : a.xyz2 + (long long)f * a.pitch2;
    const int *heavy = side == 0 ? a.t.heavy1 : a.t.heavy2;
    const int nh = side == 0 ? a.t.n1h : a.t.n2h;
    uint32_t *mycells = cells + (size_t)fb * (capCells + 1);
    float4 *mycand = cand + (size_t)fb * nh_alloc;
    uint32_t *myccell = ccell + (size_t)fb * nh_alloc;
    unsigned int *cursor = side == 0 ? &fr[fb].nA_in : &fr[fb].nB_in;
    const int lane = threadIdx.x & 31;
    const int stride = gridDim.x * blockDim.x;
    const int nround = (nh + stride - 1) / stride;
    for (int r = 0; r < nround; r++) {
        const int k = r * stride + blockIdx.x * blockDim.x + threadIdx.x;
        int c = -1;
        float x = 0, y = 0, z = 0;
        if (k < nh) {
            load_xyz(frame, __ldg(heavy + k), a.stride, x, y, z);
            c = pc_large_cell_of(F, x, y, z);
        }
        const unsigned m = __ballot_sync(PC_FULL_MASK, c >= 0);
        if (m) {
            unsigned base = 0;
            if (lane == __ffs(m) - 1) base = atomicAdd(cursor, (unsigned)__popc(m));
            base = __shfl_sync(PC_FULL_MASK, base, __ffs(m) - 1);
            if (c >= 0) {
                const unsigned p = base + __popc(m & ((1u << lane) - 1u));
                mycand[p] = make_float4(x, y, z, __int_as_float(k));
                myccell[p] = (uint32_t)c;
                atomicAdd(&mycells[c], 1u);
            }
        }
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
                                                               int nh_alloc) {
    const int fb = blockIdx.y;
    const unsigned n = side == 0 ? fr[fb].nA_in : fr[fb].nB_in;
    uint32_t *mycells = cells + (size_t)fb * (capCells + 1);
    const float4 *mycand = cand + (size_t)fb * nh_alloc;
    const uint32_t *myccell = ccell + (size_t)fb * nh_alloc;
    float4 *mysorted = sorted + (size_t)fb * nh_alloc;
    for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x)
        mysorted[atomicAdd(&mycells[myccell[p]], 1u)] = mycand[p];
}

// ---- L6: pair search + drain ---------------------------------------------------------------------
struct PcLargeOut {
    long long cap_c_frame, cap_h_frame;   // records per frame slice
};

#define PC_LQ 96    // per-warp hit queue entries (drained when >= 32)

__device__ __forceinline__ void pc_large_drain(const PcScanArgs &a, PcLargeFrame *frp, int f, int fb,
                                               const float4 *sA, const float4 *sB, const uint2 *q, int n,
                                               const PcLargeOut &o, unsigned int &status) {
    const int lane = threadIdx.x & 31;
    const bool self = a.self_mode != 0;
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(&frp->ccursor, (unsigned)n);
    base = __shfl_sync(PC_FULL_MASK, base, 0);
    if (lane >= n) return;
    const uint2 e = q[lane];
    const float4 pa = sA[e.x], pb = sB[e.y];
    const int ka = __float_as_int(pa.w), kb = __float_as_int(pb.w);
    const int li = __ldg(a.t.heavy1 + ka), lj = __ldg((self ? a.t.heavy1 : a.t.heavy2) + kb);
    const double ax = pa.x, ay = pa.y, az = pa.z, bx = pb.x, by = pb.y, bz = pb.z;
    const double dx = ax - bx, dy = ay - by, dz = az - bz;
    const double dist = sqrt(dx * dx + dy * dy + dz * dz);
    const long long slot = (long long)base + lane;
    if (slot < o.cap_c_frame) {
        pc_contact rec;
        rec.i = li; rec.j = lj; rec.distance = (float)dist; rec.weight = (float)contact_weight(dist);
        a.contacts[(long long)f * o.cap_c_frame + slot] = rec;
    } else {
        status |= (unsigned)PC_ST_OUT_OVERFLOW;
    }
    const uint8_t *hflagB = self ? a.t.hflag1 : a.t.hflag2;
    if ((__ldg(a.t.hflag1 + ka) & PC_ATOM_HBCLASS) && (__ldg(hflagB + kb) & PC_ATOM_HBCLASS)) {
        const float *fr1 = a.xyz1 + (long long)f * a.pitch1;
        const float *fr2 = self ? fr1 : a.xyz2 + (long long)f * a.pitch2;
        const int *hhptrB = self ? a.t.hhptr1 : a.t.hhptr2;
        const int *hidxB = self ? a.t.hidx1 : a.t.hidx2;
        for (int side = 0; side < 2; side++) {
            const int p0 = side == 0 ? __ldg(a.t.hhptr1 + ka) : __ldg(hhptrB + kb);
            const int p1 = side == 0 ? __ldg(a.t.hhptr1 + ka + 1) : __ldg(hhptrB + kb + 1);
            for (int p = p0; p < p1; p++) {
                const int hl = side == 0 ? __ldg(a.t.hidx1 + p) : __ldg(hidxB + p);
                if (hl < 0) { status |= (unsigned)PC_ST_MISSING_H; continue; }
                float hxf, hyf, hzf;
                load_xyz(side == 0 ? fr1 : fr2, hl, a.stride, hxf, hyf, hzf);
                double d = 0, ang = 0;
                const bool ok = side == 0
                    ? hbond_test(hxf, hyf, hzf, ax, ay, az, bx, by, bz, a.hb_cutoff, a.hb_angle, d, ang)
                    : hbond_test(hxf, hyf, hzf, bx, by, bz, ax, ay, az, a.hb_cutoff, a.hb_angle, d, ang);
                if (ok) {
                    const unsigned hq = atomicAdd(&frp->hcursor, 1u);
                    if ((long long)hq < o.cap_h_frame) {
                        pc_hbond r;
                        r.i = li; r.j = lj; r.hyd = (uint32_t)hl | (side ? 0x80000000u : 0u);
                        r.distance = (float)d; r.angle = (float)ang;
                        a.hbonds[(long long)f * o.cap_h_frame + hq] = r;
                    } else {
                        status |= (unsigned)PC_ST_OUT_OVERFLOW;
                    }
                }
            }
        }
    }
    (void)fb;
}

template <int NT>
__global__ void __launch_bounds__(NT) pc_large_pair_kernel(const PcScanArgs a, PcLargeFrame *fr, int f0, int nfb,
                                                           const uint32_t *cellsA, const uint32_t *cellsB, int capCells,
                                                           const float4 *sortA, const float4 *sortB, int n1h_alloc,
                                                           int n2h_alloc, unsigned int *work, PcLargeOut o) {
    __shared__ uint2 queue[NT / 32][PC_LQ];
    __shared__ PcLargeFrame sF[64];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < nfb; i += NT) sF[i] = fr[i];
    __syncthreads();
    const bool self = a.self_mode != 0;
    const unsigned total = work[1];
    unsigned long long evals = 0;
    unsigned int status = 0;
    uint2 *q = queue[warp];
    constexpr unsigned CHUNK = 4;      // A cells per work grab
    for (;;) {
        unsigned it0 = 0;
        if (lane == 0) it0 = atomicAdd(&work[0], CHUNK);
        it0 = __shfl_sync(PC_FULL_MASK, it0, 0);
        if (it0 >= total) break;
        for (unsigned it = it0; it < min(it0 + CHUNK, total); it++) {
            int fb = 0;
            while (fb + 1 < nfb && sF[fb + 1].item_base <= it) fb++;
            const PcLargeFrame &F = sF[fb];
            const int c = (int)(it - F.item_base);
            const uint32_t *cA = cellsA + (size_t)fb * (capCells + 1);
            const uint32_t *cB = self ? cA : cellsB + (size_t)fb * (capCells + 1);
            const int a0 = c ? (int)cA[c - 1] : 0, a1 = (int)cA[c];
            if (a0 == a1) continue;
            const float4 *sA = sortA + (size_t)fb * n1h_alloc;
            const float4 *sB = self ? sA : sortB + (size_t)fb * n2h_alloc;
            const int nx = F.nx, ny = F.ny, nz = F.nz;
            const int cx = c % nx, cy = (c / nx) % ny, cz = c / (nx * ny);
            const int xlo = max(cx - 1, 0), xhi = min(cx + 1, nx - 1);
            int nq = 0;
            for (int ab = a0; ab < a1; ab += 32) {
                const int na = min(32, a1 - ab);
                float4 pa = make_float4(0.f, 0.f, 0.f, 0.f);
                if (lane < na) pa = sA[ab + lane];
                int ares = 0, aseg = 0;
                if (self && lane < na) {
                    ares = __ldg(a.t.hres1 + __float_as_int(pa.w)); aseg = __ldg(a.t.hseg1 + __float_as_int(pa.w));
                }
                for (int dz = -1; dz <= 1; dz++) {
                    const int z = cz + dz;
                    if (z < 0 || z >= nz) continue;
                    for (int dy = -1; dy <= 1; dy++) {
                        const int y = cy + dy;
                        if (y < 0 || y >= ny) continue;
                        const int row = (z * ny + y) * nx;
                        const int b0 = (row + xlo) ? (int)cB[row + xlo - 1] : 0, b1 = (int)cB[row + xhi];
                        for (int base = b0; base < b1; base += 32) {
                            const int j = base + lane;
                            const bool valid = j < b1;
                            float4 pb = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (valid) pb = sB[j];
                            int bres = 0, bseg = 0;
                            if (self && valid) {
                                bres = __ldg(a.t.hres1 + __float_as_int(pb.w)); bseg = __ldg(a.t.hseg1 + __float_as_int(pb.w));
                            }
                            if (lane == 0) evals += (unsigned long long)na * (unsigned)min(32, b1 - base);
                            for (int ia = 0; ia < na; ia++) {
                                const float px = __shfl_sync(PC_FULL_MASK, pa.x, ia);
                                const float py = __shfl_sync(PC_FULL_MASK, pa.y, ia);
                                const float pz = __shfl_sync(PC_FULL_MASK, pa.z, ia);
                                bool hit = valid && !(dist2_exact(px, py, pz, pb.x, pb.y, pb.z) > a.sqdist);
                                if (self) {
                                    const int r1 = __shfl_sync(PC_FULL_MASK, ares, ia);
                                    const int s1 = __shfl_sync(PC_FULL_MASK, aseg, ia);
                                    if (hit && (r1 - bres) < 5 && s1 == bseg) hit = false;
                                }
                                const unsigned m = __ballot_sync(PC_FULL_MASK, hit);
                                if (m) {
                                    if (hit) q[nq + __popc(m & ((1u << lane) - 1u))] = make_uint2((unsigned)(ab + ia), (unsigned)j);
                                    nq += __popc(m);
                                    __syncwarp();
                                    if (nq >= 64) {      // keep room for one more full ballot (32)
                                        pc_large_drain(a, &fr[fb], f0 + fb, fb, sA, sB, q, 32, o, status);
                                        pc_large_drain(a, &fr[fb], f0 + fb, fb, sA, sB, q + 32, 32, o, status);
                                        __syncwarp();
                                        if (lane < nq - 64) q[lane] = q[64 + lane];
                                        nq -= 64;
                                        __syncwarp();
                                    }
                                }
                            }
                        }
                    }
                }
            }
            for (int d = 0; d < nq; d += 32) pc_large_drain(a, &fr[fb], f0 + fb, fb, sA, sB, q + d, min(32, nq - d), o, status);
            __syncwarp();
        }
    }
    evals = __shfl_sync(PC_FULL_MASK, evals, 0);
    if (lane == 0 && evals) atomicAdd(&a.counters[PC_CNT_EVALS], evals);
    status = __reduce_or_sync(PC_FULL_MASK, status);
    if (lane == 0 && status) atomicOr(&a.counters[PC_CNT_STATUS], (unsigned long long)status);
}

// ---- L7: per-frame counts and totals -----------------------------------------------------------------
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

static inline int pc_large_fail(std::string &err, const char *what, cudaError_t e) {
    err = std::string(what) + ": " + cudaGetErrorString(e);
    return PC_ERR_CUDA;
}
#define PCL_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return pc_large_fail(err, #expr, _e); } while (0)

static inline int pc_large_scan(PcLargeWork &W, const PcScanArgs &a, int num_sms, cudaStream_t s, std::string &err,
                                unsigned long long *launches) {
    const int n1h = a.t.n1h, n2h = a.self_mode ? a.t.n1h : a.t.n2h;
    const bool self = a.self_mode != 0;
    // cells per frame: a few per atom of the smaller side, bounded like the reference's MAXBOXES
    long long cc = 8ll * std::min(n1h, n2h);
    cc = std::max(4096ll, std::min(cc, 4000000ll));
    const int capCells = (int)cc;
    // frames in flight: bound the work memory to ~2 GB
    const size_t per_frame = (size_t)(capCells + 1) * 4 * (self ? 1 : 2) + (size_t)n1h * 36 + (self ? 0 : (size_t)n2h * 36);
    int FB = (int)std::max<size_t>(1, std::min<size_t>(64, (2ull << 30) / std::max<size_t>(per_frame, 1)));
    FB = std::min(FB, a.nframes);
    if (W.FB < FB || W.capCells != capCells || W.n1h != n1h || W.n2h != n2h || W.self_mode != a.self_mode) {
        W.release();
        PCL_TRY(cudaMalloc(&W.frames, sizeof(PcLargeFrame) * 64));
        PCL_TRY(cudaMalloc(&W.cellA, (size_t)FB * (capCells + 1) * 4));
        PCL_TRY(cudaMalloc(&W.candA, (size_t)FB * n1h * sizeof(float4)));
        PCL_TRY(cudaMalloc(&W.ccellA, (size_t)FB * n1h * 4));
        PCL_TRY(cudaMalloc(&W.sortA, (size_t)FB * n1h * sizeof(float4)));
        if (!self) {
            PCL_TRY(cudaMalloc(&W.cellB, (size_t)FB * (capCells + 1) * 4));
            PCL_TRY(cudaMalloc(&W.candB, (size_t)FB * n2h * sizeof(float4)));
            PCL_TRY(cudaMalloc(&W.ccellB, (size_t)FB * n2h * 4));
            PCL_TRY(cudaMalloc(&W.sortB, (size_t)FB * n2h * sizeof(float4)));
        }
        PCL_TRY(cudaMalloc(&W.work, 16));
        W.FB = FB; W.capCells = capCells; W.n1h = n1h; W.n2h = n2h; W.self_mode = a.self_mode;
    }
    FB = W.FB;
    PcLargeOut o;
    o.cap_c_frame = a.cap_contacts / a.nframes;
    o.cap_h_frame = a.cap_hbonds / a.nframes;
    if (o.cap_c_frame < 1 || o.cap_h_frame < 1) { err = "pc_reserve: capacities smaller than the number of frames"; return PC_ERR_INVALID; }
    const int use_bboxB = (!self && (long long)n2h <= 4ll * n1h) ? 1 : 0;
    const int blocksA = std::max(1, std::min((n1h + 255) / 256, 8 * num_sms));
    const int blocksB = std::max(1, std::min((n2h + 255) / 256, 8 * num_sms));
    for (int f0 = 0; f0 < a.nframes; f0 += FB) {
        const int nfb = std::min(FB, a.nframes - f0);
        pc_large_reset_kernel<<<1, 64, 0, s>>>(W.frames, nfb, W.work);
        PCL_TRY(cudaMemsetAsync(W.cellA, 0, (size_t)nfb * (capCells + 1) * 4, s));
        if (!self) PCL_TRY(cudaMemsetAsync(W.cellB, 0, (size_t)nfb * (capCells + 1) * 4, s));
        pc_large_bbox_kernel<<<dim3(blocksA, nfb), 256, 0, s>>>(a, W.frames, f0, 0, 0);
        if (use_bboxB) pc_large_bbox_kernel<<<dim3(blocksB, nfb), 256, 0, s>>>(a, W.frames, f0, 1, 6);
        pc_large_grid_kernel<<<1, 32, 0, s>>>(a, W.frames, nfb, capCells, use_bboxB, a.counters, W.work);
        pc_large_bin_kernel<<<dim3(blocksA, nfb), 256, 0, s>>>(a, W.frames, f0, 0, W.cellA, capCells, W.candA, W.ccellA, n1h);
        if (!self)
            pc_large_bin_kernel<<<dim3(blocksB, nfb), 256, 0, s>>>(a, W.frames, f0, 1, W.cellB, capCells, W.candB, W.ccellB, n2h);
        pc_large_scan_kernel<<<dim3(nfb, self ? 1 : 2), 1024, 0, s>>>(W.frames, W.cellA, W.cellB, capCells);
        pc_large_scatter_kernel<<<dim3(blocksA, nfb), 256, 0, s>>>(W.frames, 0, W.cellA, capCells, W.candA, W.ccellA, W.sortA, n1h);
        if (!self)
            pc_large_scatter_kernel<<<dim3(blocksB, nfb), 256, 0, s>>>(W.frames, 1, W.cellB, capCells, W.candB, W.ccellB, W.sortB, n2h);
        pc_large_pair_kernel<256><<<num_sms * 6, 256, 0, s>>>(a, W.frames, f0, nfb, W.cellA, W.cellB, capCells, W.sortA,
                                                            W.sortB, n1h, n2h, W.work, o);
        pc_large_finish_kernel<<<1, 64, 0, s>>>(a, W.frames, f0, nfb, o);
        PCL_TRY(cudaGetLastError());
        *launches += 7 + (use_bboxB ? 1 : 0) + (self ? 0 : 2);
    }
    return PC_OK;
}
