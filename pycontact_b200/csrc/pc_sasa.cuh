// pc_sasa.cuh -- solvent-accessible surface (Shrake-Rupley) and "within" queries on a per-frame cell list
// (sm_100a).  SURVEY.md row f-4: the two remaining routines of the reference's cy_modules,
//   sasa_grid    PyContact/cy_modules/src/gridsearch.C:430-529  (neighbour list: vmd_gridsearch1 :207-407)
//   find_within  PyContact/cy_modules/src/gridsearch.C:595-713  (+ find_within_routine :715-826)
//
// Both results are SETS decided by float32 comparisons, so the cell geometry is free as long as every
// pair the reference would test-and-accept is tested here with the same arithmetic:
//   * neighbour of atom i: 0.001 <= d2 <= pairdist^2, d2 = ((dx*dx + dy*dy) + dz*dz) non-fused (:48-55, :363-368)
//   * surface point: loc + rad*pt (multiply, then add), buried iff ((dx*dx + dy*dy) + dz*dz) <= radsq (:502-516)
//   * atom area: ((prefac*rad)*rad)*surfpts, total = running float32 sum in atom order (:522-523)
//   * within: ((dx*dx) + (dy*dy)) + (dz*dz) < r2, strict (:786-789)
// Frames are binned with a counting sort per frame in HBM (cell edge >= the search distance, grown x1.26
// like :244-256 while the grid exceeds its cell budget); a warp owns one atom: it gathers the atom's
// neighbour spheres from the 27 cells into a shared-memory list (ballot compaction), then its lanes own the
// sphere points and walk that list with the reference's early exit.
#pragma once
#include "pc_common.cuh"

struct PcGridGeom {
    float minx, miny, minz, inv;
    int nx, ny, nz, ncell;      // ncell == 0: no atom in the binned set
};

// order-preserving float <-> uint map for atomicMin / atomicMax on coordinates
__device__ __forceinline__ unsigned pcg_f2ord(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float pcg_ord2f(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// bbox[f][0..2] = 0xffffffff (min), bbox[f][3..5] = 0 (max)
__global__ void pcg_init_kernel(unsigned *bbox, int nframes) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nframes * 6) bbox[i] = (i % 6) < 3 ? 0xffffffffu : 0u;
}

// grid (chunks, frames).  mask == NULL: every atom is binned.
__global__ void __launch_bounds__(256) pcg_bbox_kernel(const float *xyz, long long pitch, int n, const int *mask,
                                                        unsigned *bbox) {
    const float *fr = xyz + (long long)blockIdx.y * pitch;
    float lx = INFINITY, ly = INFINITY, lz = INFINITY, hx = -INFINITY, hy = -INFINITY, hz = -INFINITY;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (mask && !__ldg(mask + i)) continue;
        const float x = __ldg(fr + 3ll * i), y = __ldg(fr + 3ll * i + 1), z = __ldg(fr + 3ll * i + 2);
        lx = fminf(lx, x); ly = fminf(ly, y); lz = fminf(lz, z);
        hx = fmaxf(hx, x); hy = fmaxf(hy, y); hz = fmaxf(hz, z);
    }
    lx = warp_min(lx); ly = warp_min(ly); lz = warp_min(lz);
    hx = warp_max(hx); hy = warp_max(hy); hz = warp_max(hz);
    if ((threadIdx.x & 31) == 0 && lx <= hx) {
        unsigned *b = bbox + 6 * blockIdx.y;
        atomicMin(b + 0, pcg_f2ord(lx)); atomicMin(b + 1, pcg_f2ord(ly)); atomicMin(b + 2, pcg_f2ord(lz));
        atomicMax(b + 3, pcg_f2ord(hx)); atomicMax(b + 4, pcg_f2ord(hy)); atomicMax(b + 5, pcg_f2ord(hz));
    }
}

// one thread per frame: cell edge `edge` (already a hair above the search distance), grown x1.26 until
// the grid fits cap_cells (gridsearch.C:244-256 does the same against its 4e6-box budget)
__global__ void pcg_geom_kernel(const unsigned *bbox, int nframes, float edge, int cap_cells, PcGridGeom *geom) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nframes) return;
    const unsigned *b = bbox + 6 * f;
    PcGridGeom g;
    g.minx = g.miny = g.minz = 0.f; g.inv = 0.f; g.nx = g.ny = g.nz = g.ncell = 0;
    if (b[0] != 0xffffffffu) {
        g.minx = pcg_ord2f(b[0]); g.miny = pcg_ord2f(b[1]); g.minz = pcg_ord2f(b[2]);
        const float rx = pcg_ord2f(b[3]) - g.minx, ry = pcg_ord2f(b[4]) - g.miny, rz = pcg_ord2f(b[5]) - g.minz;
        float e = edge;
        for (int it = 0; it < 200; it++) {
            g.inv = 1.0f / e;
            const float fx = rx * g.inv, fy = ry * g.inv, fz = rz * g.inv;
            const bool sane = fx < 2.0e6f && fy < 2.0e6f && fz < 2.0e6f;     // also false for inf / nan extents
            if (sane) {
                g.nx = (int)fx + 1; g.ny = (int)fy + 1; g.nz = (int)fz + 1;
                const long long tot = (long long)g.nx * g.ny * g.nz;
                if (tot <= cap_cells) { g.ncell = (int)tot; break; }
            }
            e *= 1.26f;
        }
        if (g.ncell == 0) { g.nx = g.ny = g.nz = 1; g.ncell = 1; g.inv = 0.f; }     // non-finite extents: one cell
    }
    geom[f] = g;
}

__device__ __forceinline__ void pcg_cell_of(const PcGridGeom &g, float x, float y, float z, int &cx, int &cy, int &cz) {
    // inside the binned set's box the products are in [0, n); queries from outside are clamped to [-2, n+1]
    const float fx = fminf(fmaxf(floorf((x - g.minx) * g.inv), -2.f), (float)(g.nx + 1));
    const float fy = fminf(fmaxf(floorf((y - g.miny) * g.inv), -2.f), (float)(g.ny + 1));
    const float fz = fminf(fmaxf(floorf((z - g.minz) * g.inv), -2.f), (float)(g.nz + 1));
    cx = (int)fx; cy = (int)fy; cz = (int)fz;                                  // NaN -> 0
}

// grid (chunks, frames): cell of every binned atom and its arrival rank inside the cell
__global__ void __launch_bounds__(256) pcg_count_kernel(const float *xyz, long long pitch, int n, const int *mask,
                                                         const PcGridGeom *geom, int cell_pitch, int *cellcnt,
                                                         int *cellid, int *rank) {
    const int f = blockIdx.y;
    const PcGridGeom g = geom[f];
    const float *fr = xyz + (long long)f * pitch;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int cell = -1, rk = 0;
        if (g.ncell && (!mask || __ldg(mask + i))) {
            int cx, cy, cz;
            pcg_cell_of(g, __ldg(fr + 3ll * i), __ldg(fr + 3ll * i + 1), __ldg(fr + 3ll * i + 2), cx, cy, cz);
            cx = min(max(cx, 0), g.nx - 1); cy = min(max(cy, 0), g.ny - 1); cz = min(max(cz, 0), g.nz - 1);
            cell = (cz * g.ny + cy) * g.nx + cx;
            rk = atomicAdd(cellcnt + (long long)f * cell_pitch + cell, 1);
        }
        cellid[(long long)f * n + i] = cell;
        rank[(long long)f * n + i] = rk;
    }
}

// one 1024-thread CTA per frame: counts -> exclusive starts, in place; cellcnt[f][ncell] = binned atoms
__global__ void __launch_bounds__(1024) pcg_scan_kernel(const PcGridGeom *geom, int cell_pitch, int *cellcnt) {
    __shared__ int s_warp[32];
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int ncell = geom[f].ncell;
    int *c = cellcnt + (long long)f * cell_pitch;
    const int per = (ncell + 1023) / 1024;
    const int b = min(tid * per, ncell), e = min(b + per, ncell);
    int sum = 0;
    for (int i = b; i < e; i++) sum += c[i];
    int incl = (int)warp_incl_scan((uint32_t)sum, lane);
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    if (w == 0) {
        const int v = s_warp[lane];
        const int s = (int)warp_incl_scan((uint32_t)v, lane);
        s_warp[lane] = s - v;
    }
    __syncthreads();
    int run = s_warp[w] + incl - sum;
    for (int i = b; i < e; i++) { const int v = c[i]; c[i] = run; run += v; }
    if (tid == 1023) c[ncell] = run;     // the last thread's inclusive prefix = atoms binned in this frame
}

// grid (chunks, frames): cell-sorted float4 (x, y, z, w[i]) and the original index of every slot
__global__ void __launch_bounds__(256) pcg_scatter_kernel(const float *xyz, long long pitch, int n, const float *w,
                                                           int cell_pitch, const int *cellstart, const int *cellid,
                                                           const int *rank, float4 *sorted, int *sidx) {
    const int f = blockIdx.y;
    const float *fr = xyz + (long long)f * pitch;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int cell = cellid[(long long)f * n + i];
        if (cell < 0) continue;
        const int pos = cellstart[(long long)f * cell_pitch + cell] + rank[(long long)f * n + i];
        sorted[(long long)f * n + pos] = make_float4(__ldg(fr + 3ll * i), __ldg(fr + 3ll * i + 1), __ldg(fr + 3ll * i + 2),
                                                     w ? __ldg(w + i) : 0.f);
        sidx[(long long)f * n + pos] = i;
    }
}

#define PCS_WARPS 8
#define PCS_LIST 192          // neighbour spheres staged per warp between two passes over the sphere points
#ifndef PCS_STEP
#define PCS_STEP 16           // neighbour spheres of phase 1 (lanes own points)
#endif
#ifndef PCS_SWITCH
#define PCS_SWITCH 12         // phase 1 runs only while more points than this are live
#endif
#define PCS_MAX_POINTS 1024

// rows of the 5 x 5 x 5 half-size-cell neighbourhood, the atom's own row first: near spheres bury most points, so the
// live list shrinks early
// index = (dz + 2) * 5 + (dy + 2), ordered by |dz| + |dy|
__constant__ signed char PCS_ROW25[25] = {12, 7, 11, 13, 17, 6, 8, 16, 18, 2, 10, 14, 22, 1, 3, 5, 9, 15, 19, 21, 23, 0, 4, 20, 24};

static inline size_t pc_sasa_smem(int npts) {       // dynamic shared memory of pc_sasa_atoms_kernel
    return 3 * (size_t)npts * 4 + (size_t)PCS_WARPS * (((size_t)npts + 31) & ~(size_t)31) * 2;
}

// grid (ceil(n / PCS_WARPS), frames); dynamic shared memory: the 3 * npts shared sphere points + one list of
// live (not yet buried) point indices per warp.  sorted[].w = radius + probe radius (float32, formed like
// gridsearch.C:497).  A warp owns one atom and keeps a compacted list of its live points (see flush below):
// the per-point early exit of gridsearch.C:506-516 without its divergence.
__global__ void __launch_bounds__(PCS_WARPS * 32) pc_sasa_atoms_kernel(
        const float4 *sorted, const int *sidx, const int *cellstart, const PcGridGeom *geom, int n, int cell_pitch,
        const float *pts, int npts, float sqdist, float reach, int restricted, const int *rlist, float prefac, float *area,
        int *exposed, unsigned long long *evals) {
    __shared__ float4 s_list[PCS_WARPS][PCS_LIST];
    __shared__ int s_rend[PCS_WARPS][32], s_rofs[PCS_WARPS][32];
    extern __shared__ float s_pts[];
    const int f = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int npad = (npts + 31) & ~31;
    unsigned short *s_live = reinterpret_cast<unsigned short *>(s_pts + 3 * npts) + warp * npad;
    for (int i = threadIdx.x; i < 3 * npts; i += blockDim.x) s_pts[i] = pts[i];
    for (int i = lane; i < npts; i += 32) s_live[i] = (unsigned short)i;
    __syncthreads();
    const PcGridGeom g = geom[f];
    if (!g.ncell) return;
    const int *cs = cellstart + (long long)f * cell_pitch;
    const int slot = blockIdx.x * PCS_WARPS + warp;
    if (slot >= cs[g.ncell]) return;
    const float4 *fs = sorted + (long long)f * n;
    const float4 me = fs[slot];
    const int orig = sidx[(long long)f * n + slot];
    if (restricted && !__ldg(rlist + orig)) {               // gridsearch.C:494: only atoms of the restriction contribute
        if (lane == 0) { area[(long long)f * n + orig] = 0.f; if (exposed) exposed[(long long)f * n + orig] = 0; }
        return;
    }
    const unsigned lt = (1u << lane) - 1u;
    unsigned long long my_evals = 0;
    int nlive = npts;                                       // same value in every lane

    // the live points against the staged neighbour spheres (gridsearch.C:500-521), in two phases:
    //  1. lanes own points, the first PCS_STEP spheres (the nearest rows come first): most buried points die here;
    //  2. the survivors one at a time, lanes own the remaining spheres 32 at a time (warp-uniform control flow;
    //     a point leaves at the first 32-sphere chunk that buries it) -- a handful of long-lived points no
    //     longer drags the whole warp down the list one sphere per iteration.
    auto flush = [&](int cnt) {
        __syncwarp();
        int k0 = 0;
        if (nlive > PCS_SWITCH) {
            k0 = min(cnt, PCS_STEP);
            int w = 0;
            for (int c = 0; c < nlive; c += 32) {
                const int idx = c + lane;
                bool live = false;
                int p = 0;
                if (idx < nlive) {
                    p = s_live[idx];
                    const float sx = __fadd_rn(me.x, __fmul_rn(me.w, s_pts[3 * p]));
                    const float sy = __fadd_rn(me.y, __fmul_rn(me.w, s_pts[3 * p + 1]));
                    const float sz = __fadd_rn(me.z, __fmul_rn(me.w, s_pts[3 * p + 2]));
                    live = true;
                    int k = 0;
                    for (; k < k0; k++) {
                        const float4 q = s_list[warp][k];
                        if (dist2_exact(sx, sy, sz, q.x, q.y, q.z) <= q.w) { live = false; break; }
                    }
                    my_evals += (unsigned long long)(k + (live ? 0 : 1));
                }
                // every entry of this chunk has been read by its lane before the ballot; the survivors go to
                // positions <= their old ones
                const unsigned m = __ballot_sync(PC_FULL_MASK, live);
                __syncwarp();
                if (live) s_live[w + __popc(m & lt)] = (unsigned short)p;
                w += __popc(m);
                __syncwarp();
            }
            nlive = w;
        }
        if (k0 < cnt && nlive) {
            int w = 0;
            for (int i = 0; i < nlive; i++) {
                const int p = s_live[i];
                const float sx = __fadd_rn(me.x, __fmul_rn(me.w, s_pts[3 * p]));
                const float sy = __fadd_rn(me.y, __fmul_rn(me.w, s_pts[3 * p + 1]));
                const float sz = __fadd_rn(me.z, __fmul_rn(me.w, s_pts[3 * p + 2]));
                bool bur = false;
                for (int kb = k0; kb < cnt; kb += 32) {
                    const int k = kb + lane;
                    bool hit = false;
                    if (k < cnt) {
                        const float4 q = s_list[warp][k];
                        hit = dist2_exact(sx, sy, sz, q.x, q.y, q.z) <= q.w;
                        my_evals++;
                    }
                    if (__any_sync(PC_FULL_MASK, hit)) { bur = true; break; }
                }
                if (!bur) {                                 // w <= i: that entry has been consumed
                    if (lane == 0) s_live[w] = (unsigned short)p;
                    w++;
                }
            }
            nlive = w;
        }
        __syncwarp();
    };

    // Candidates: the grid's cells are HALF the search distance wide, so the neighbourhood is 5 x 5 rows of up to five
    // cells.  Lane r < 25 owns one row (nearest rows first): it drops the row when its (y, z) gap to the atom already
    // exceeds `reach` (= pairdist plus a margin far above the rounding of these cell coordinates), trims the row's
    // x range to the chord of the search sphere, and fetches the bounds of what is left.  An inclusive scan turns the
    // rows into ONE candidate list (row table in shared memory, found by binary search) that the warp walks 32
    // atoms at a time: about twice the sphere's volume instead of the 27 full-size cells' 6.4 times.
    int cnt = 0;
    {
        const float ux = (me.x - g.minx) * g.inv, uy = (me.y - g.miny) * g.inv, uz = (me.z - g.minz) * g.inv;
        int rb = 0, rlen = 0;
        if (lane < 25) {
            const int rr = PCS_ROW25[lane];
            const int y = min(max((int)floorf(uy), 0), g.ny - 1) + rr % 5 - 2, z = min(max((int)floorf(uz), 0), g.nz - 1) + rr / 5 - 2;
            if (y >= 0 && y < g.ny && z >= 0 && z < g.nz) {
                int xlo = 0, xhi = g.nx - 1;
                bool keep = true;
                if (g.inv > 0.f) {
                    const float e = 1.0f / g.inv;
                    const float gy = fmaxf(0.f, fmaxf((float)y - uy, uy - (float)(y + 1))) * e;
                    const float gz = fmaxf(0.f, fmaxf((float)z - uz, uz - (float)(z + 1))) * e;
                    const float rem = reach * reach - gy * gy - gz * gz;
                    keep = rem >= 0.f;
                    if (keep) {
                        const float hw = sqrtf(rem) * g.inv;
                        xlo = max((int)floorf(ux - hw), 0);
                        xhi = min((int)floorf(ux + hw), g.nx - 1);
                        keep = xlo <= xhi;
                    }
                }
                if (keep) {
                    const int row = (z * g.ny + y) * g.nx;
                    rb = cs[row + xlo];
                    rlen = cs[row + xhi + 1] - rb;
                }
            }
        }
        const int incl = (int)warp_incl_scan((uint32_t)rlen, lane);
        const int total = __shfl_sync(PC_FULL_MASK, incl, 31);
        s_rend[warp][lane] = incl;                          // list positions [incl - rlen, incl) belong to this lane's row
        s_rofs[warp][lane] = rb - (incl - rlen);            // sorted index = list position + offset
        __syncwarp();
        for (int t0 = 0; t0 < total && nlive; t0 += 32) {
            const int t = t0 + lane;
            bool ok = false;
            float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t < total) {
                int r = 0;                                   // smallest r with t < s_rend[r]
#pragma unroll
                for (int step = 16; step > 0; step >>= 1)
                    if (s_rend[warp][r + step - 1] <= t) r += step;
                q = fs[t + s_rofs[warp][r]];
                const float d2 = dist2_exact(me.x, me.y, me.z, q.x, q.y, q.z);
                // the reference's neighbour list (:363-368).  A sphere further away than the two radii (plus a
                // margin 1000x the float32 rounding of the point test) cannot bury any point of this atom.
                const float rs = me.w + q.w;
                ok = d2 >= 0.001f && d2 <= sqdist && d2 <= rs * rs * 1.001f;
            }
            const unsigned m = __ballot_sync(PC_FULL_MASK, ok);
            const int add = __popc(m);
            if (cnt + add > PCS_LIST) { flush(cnt); cnt = 0; }
            if (ok) s_list[warp][cnt + __popc(m & lt)] = make_float4(q.x, q.y, q.z, __fmul_rn(q.w, q.w));
            cnt += add;
        }
    }
    if (nlive) flush(cnt);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my_evals += __shfl_xor_sync(PC_FULL_MASK, my_evals, o);
    if (lane == 0) {
        // nlive = surfpts of gridsearch.C:499-521
        area[(long long)f * n + orig] = __fmul_rn(__fmul_rn(__fmul_rn(prefac, me.w), me.w), (float)nlive);
        if (exposed) exposed[(long long)f * n + orig] = nlive;
        if (evals) atomicAdd(evals, my_evals);
    }
}

// one CTA per frame: the reference's running float32 sum in atom order (gridsearch.C:523) -- the order is part
// of the result, so one thread adds while the CTA streams the areas through shared memory
#define PCS_SUM_TILE 4096
__global__ void __launch_bounds__(256) pc_sasa_sum_kernel(const float *area, int n, float *total) {
    __shared__ float s_a[PCS_SUM_TILE];
    const float *a = area + (long long)blockIdx.x * n;
    float t = 0.f;
    for (int base = 0; base < n; base += PCS_SUM_TILE) {
        const int cnt = min(PCS_SUM_TILE, n - base);
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) s_a[i] = a[base + i];
        __syncthreads();
        if (threadIdx.x == 0) {
            int k = 0;
            for (; k + 4 <= cnt; k += 4) {
                const float4 v = *reinterpret_cast<const float4 *>(s_a + k);
                t = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(t, v.x), v.y), v.z), v.w);
            }
            for (; k < cnt; k++) t = __fadd_rn(t, s_a[k]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) total[blockIdx.x] = t;
}

// grid (chunks, frames): find_within -- the cell list holds the `others` atoms; a thread owns one flagged atom
__global__ void __launch_bounds__(256) pc_within_kernel(const float *xyz, long long pitch, int n, const int *flgs,
                                                         const float4 *sorted, const int *cellstart,
                                                         const PcGridGeom *geom, int cell_pitch, float r2, int *result) {
    const int f = blockIdx.y;
    const PcGridGeom g = geom[f];
    const float *fr = xyz + (long long)f * pitch;
    const int *cs = cellstart + (long long)f * cell_pitch;
    const float4 *fs = sorted + (long long)f * n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int res = 0;
        if (g.ncell && __ldg(flgs + i)) {
            const float x = __ldg(fr + 3ll * i), y = __ldg(fr + 3ll * i + 1), z = __ldg(fr + 3ll * i + 2);
            int cx, cy, cz;
            pcg_cell_of(g, x, y, z, cx, cy, cz);
            const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.nx - 1);
            if (x0 <= x1) {
                for (int dz = -1; dz <= 1 && !res; dz++) {
                    const int zz = cz + dz;
                    if (zz < 0 || zz >= g.nz) continue;
                    for (int dy = -1; dy <= 1 && !res; dy++) {
                        const int yy = cy + dy;
                        if (yy < 0 || yy >= g.ny) continue;
                        const int row = (zz * g.ny + yy) * g.nx;
                        const int b = cs[row + x0], e = cs[row + x1 + 1];
                        for (int j = b; j < e; j++) {
                            const float4 q = fs[j];
                            if (dist2_exact(x, y, z, q.x, q.y, q.z) < r2) { res = 1; break; }
                        }
                    }
                }
            }
        }
        result[(long long)f * n + i] = res;
    }
}
