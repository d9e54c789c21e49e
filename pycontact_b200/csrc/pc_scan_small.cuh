// pc_scan_small.cuh -- "small frame" path: one CTA analyses one whole frame out of shared memory.
//
// Replaces, for one frame, the body of loop_trajectory_grid (core/multi_trajectory.py:65-219):
//   find_contacts / vmd_gridsearch3 (cy_modules/src/gridsearch.C:408-427, 902-1162)
//   + heavy-atom filter (:89) + self-interaction filter (:93-95) + distance/weight (:98-107)
//   + H-bond test (:112-209).
//
// Phases (all in one CTA, frames are strided over the grid):
//   1  heavy atoms of both selections: global -> shared SoA (x[],y[],z[]), bounding boxes
//   2  grid over (bboxA (+) c) n (bboxB (+) c): only atoms inside it can have a partner
//   3  count atoms per cell (shared atomics)        4  exclusive scan (in place)
//   5  scatter heavy ordinals into cell order (perm arrays; the scan array becomes "cell end")
//   6  pair search: one warp per non-empty A cell; for each of the 9 x-runs of B cells, lanes hold
//      B atoms, A atoms are broadcast by shuffle; exact float32 test; hits are warp-compacted into a
//      shared hit queue
//   7  drain: one thread per hit -> float64 distance + weight, H-bond test on N/O/S pairs, records
//      written to one contiguous range of the batch output reserved with a single atomicAdd
//
// Hydrogens never enter the search (the reference discards every pair with a hydrogen at :89); they
// are read from global memory only inside the H-bond test.
#pragma once
#include "pc_common.cuh"

struct PcSmallSmem {
    // offsets (bytes) into dynamic shared memory, computed on the host
    int off_x, off_y, off_z, off_permA, off_permB, off_cellA, off_cellB, off_hits, off_hb, off_keys, total;
};

__host__ inline PcSmallSmem pc_small_layout(int n1h, int n2h, int self_mode, int capCells, int capHits,
                                            int capHb, int order_keys) {
    PcSmallSmem L;
    const int ntot = self_mode ? n1h : n1h + n2h;
    int o = 0;
    auto take = [&](int bytes) { int r = o; o += (bytes + 15) & ~15; return r; };
    L.off_x = take(4 * ntot); L.off_y = take(4 * ntot); L.off_z = take(4 * ntot);
    L.off_permA = take(2 * n1h);
    L.off_permB = self_mode ? L.off_permA : take(2 * n2h);
    L.off_cellA = take(4 * (capCells + 1));
    L.off_cellB = self_mode ? L.off_cellA : take(4 * (capCells + 1));
    L.off_hits = take(4 * capHits);
    L.off_hb = take((int)sizeof(pc_hbond) * capHb);
    L.off_keys = 0;
    (void)order_keys;
    L.total = o;
    return L;
}

struct PcGridSm {
    float lo[3];
    float inv;
    int nx, ny, nz, ncell;
    // reference grid (order keys only)
    float rmin[3];
    float rinv;
    int rnb[3];
};

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

template <int NT>
__global__ void __launch_bounds__(NT) pc_scan_small_kernel(const PcScanArgs a, const PcSmallSmem L) {
    extern __shared__ __align__(16) unsigned char smem[];
    float *sx = reinterpret_cast<float *>(smem + L.off_x);
    float *sy = reinterpret_cast<float *>(smem + L.off_y);
    float *sz = reinterpret_cast<float *>(smem + L.off_z);
    uint16_t *permA = reinterpret_cast<uint16_t *>(smem + L.off_permA);
    uint16_t *permB = reinterpret_cast<uint16_t *>(smem + L.off_permB);
    uint32_t *cellA = reinterpret_cast<uint32_t *>(smem + L.off_cellA);
    uint32_t *cellB = reinterpret_cast<uint32_t *>(smem + L.off_cellB);
    uint32_t *hits = reinterpret_cast<uint32_t *>(smem + L.off_hits);
    pc_hbond *hbq = reinterpret_cast<pc_hbond *>(smem + L.off_hb);

    __shared__ float red[NT / 32][12];
    __shared__ uint32_t wsum[32];
    __shared__ PcGridSm g;
    __shared__ int s_next;
    __shared__ unsigned int s_nhits, s_nhb;
    __shared__ long long s_cbase, s_hbase;
    __shared__ unsigned int s_status;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool self = a.self_mode != 0;
    const int n1h = a.t.n1h, n2h = self ? a.t.n1h : a.t.n2h;
    const int offB = self ? 0 : n1h;                 // B atoms live at sx[offB + k]
    const int *heavyB = self ? a.t.heavy1 : a.t.heavy2;
    const uint8_t *hflagB = self ? a.t.hflag1 : a.t.hflag2;
    const int *hhptrB = self ? a.t.hhptr1 : a.t.hhptr2;
    const int *hidxB = self ? a.t.hidx1 : a.t.hidx2;
    unsigned long long evals = 0;

    for (int f = blockIdx.x; f < a.nframes; f += gridDim.x) {
        const float *fr1 = a.xyz1 + (long long)f * a.pitch1;
        const float *fr2 = self ? fr1 : a.xyz2 + (long long)f * a.pitch2;

        // ---- phase 1: heavy atoms -> shared, bounding boxes ---------------------------------
        float mn[6], mx[6];
#pragma unroll
        for (int k = 0; k < 6; k++) { mn[k] = INFINITY; mx[k] = -INFINITY; }
        for (int k = tid; k < n1h; k += NT) {
            float x, y, z;
            load_xyz(fr1, __ldg(a.t.heavy1 + k), a.stride, x, y, z);
            sx[k] = x; sy[k] = y; sz[k] = z;
            mn[0] = fminf(mn[0], x); mn[1] = fminf(mn[1], y); mn[2] = fminf(mn[2], z);
            mx[0] = fmaxf(mx[0], x); mx[1] = fmaxf(mx[1], y); mx[2] = fmaxf(mx[2], z);
        }
        if (!self) {
            for (int k = tid; k < n2h; k += NT) {
                float x, y, z;
                load_xyz(fr2, __ldg(a.t.heavy2 + k), a.stride, x, y, z);
                sx[offB + k] = x; sy[offB + k] = y; sz[offB + k] = z;
                mn[3] = fminf(mn[3], x); mn[4] = fminf(mn[4], y); mn[5] = fminf(mn[5], z);
                mx[3] = fmaxf(mx[3], x); mx[4] = fmaxf(mx[4], y); mx[5] = fmaxf(mx[5], z);
            }
        }
#pragma unroll
        for (int k = 0; k < 6; k++) { mn[k] = warp_min(mn[k]); mx[k] = warp_max(mx[k]); }
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 6; k++) { red[warp][k] = mn[k]; red[warp][6 + k] = mx[k]; }
        }
        if (tid == 0) { s_nhits = 0; s_nhb = 0; s_next = 0; s_status = 0; }
        __syncthreads();

        // ---- phase 2: grid geometry (thread 0) --------------------------------------------------
        if (tid == 0) {
            float lo6[6], hi6[6];
            for (int k = 0; k < 6; k++) {
                float l = INFINITY, h = -INFINITY;
                for (int w = 0; w < NT / 32; w++) { l = fminf(l, red[w][k]); h = fmaxf(h, red[w][6 + k]); }
                lo6[k] = l; hi6[k] = h;
            }
            if (self) for (int k = 0; k < 3; k++) { lo6[3 + k] = lo6[k]; hi6[3 + k] = hi6[k]; }
            bool empty = (n1h == 0 || n2h == 0), bad = false;
            float lo[3], ext[3];
            for (int k = 0; k < 3; k++) {
                const float l = fmaxf(lo6[k], lo6[3 + k]) - a.edge;
                const float h = fminf(hi6[k], hi6[3 + k]) + a.edge;
                if (!(l <= h)) empty = true;
                if (!isfinite(l) || !isfinite(h)) bad = true;
                lo[k] = l; ext[k] = h - l;
            }
            if (bad && !empty) { atomicOr(&s_status, (unsigned)PC_ST_BAD_COORDS); empty = true; }
            int nx = 0, ny = 0, nz = 0, ncell = 0;
            float inv = 0.f;
            if (!empty) {
                float e = a.edge;
                for (;;) {
                    inv = 1.0f / e;
                    const long long lx = (long long)(ext[0] * inv) + 1, ly = (long long)(ext[1] * inv) + 1,
                                    lz = (long long)(ext[2] * inv) + 1;
                    if (lx * ly * lz <= (long long)a.capCells) { nx = (int)lx; ny = (int)ly; nz = (int)lz; break; }
                    e *= 1.26f;
                }
                ncell = nx * ny * nz;
            }
            g.lo[0] = lo[0]; g.lo[1] = lo[1]; g.lo[2] = lo[2];
            g.inv = inv; g.nx = nx; g.ny = ny; g.nz = nz; g.ncell = ncell;
        }
        __syncthreads();
        const int ncell = g.ncell;
        if (ncell > 0) {
            const int nx = g.nx, ny = g.ny, nz = g.nz;
            const float lox = g.lo[0], loy = g.lo[1], loz = g.lo[2], inv = g.inv;
            // upper corner of the grid box: atoms beyond it cannot have a partner
            const float hix = lox + (float)nx / inv, hiy = loy + (float)ny / inv, hiz = loz + (float)nz / inv;
            (void)hix; (void)hiy; (void)hiz;

            // ---- phase 3: count ---------------------------------------------------------------------
            for (int c = tid; c <= ncell; c += NT) { cellA[c] = 0; if (!self) cellB[c] = 0; }
            __syncthreads();
            auto cell_of = [&](float x, float y, float z) -> int {
                // inside test in cell units: u in [0, n) -- atoms outside are farther than one cell
                // edge (> cutoff) from the other selection's bounding box
                const float ux = (x - lox) * inv, uy = (y - loy) * inv, uz = (z - loz) * inv;
                if (!(ux >= 0.f && uy >= 0.f && uz >= 0.f)) return -1;
                const int cx = (int)ux, cy = (int)uy, cz = (int)uz;
                if (cx >= nx || cy >= ny || cz >= nz) return -1;
                return (cz * ny + cy) * nx + cx;
            };
            for (int k = tid; k < n1h; k += NT) {
                const int c = cell_of(sx[k], sy[k], sz[k]);
                if (c >= 0) atomicAdd(&cellA[c], 1u);
            }
            if (!self)
                for (int k = tid; k < n2h; k += NT) {
                    const int c = cell_of(sx[offB + k], sy[offB + k], sz[offB + k]);
                    if (c >= 0) atomicAdd(&cellB[c], 1u);
                }
            __syncthreads();
            // ---- phase 4: scans ---------------------------------------------------------------------
            block_excl_scan_inplace<NT>(cellA, ncell, wsum);
            if (!self) block_excl_scan_inplace<NT>(cellB, ncell, wsum);
            // ---- phase 5: scatter ordinals; afterwards cell[c] == END of cell c -----------------------
            for (int k = tid; k < n1h; k += NT) {
                const int c = cell_of(sx[k], sy[k], sz[k]);
                if (c >= 0) permA[atomicAdd(&cellA[c], 1u)] = (uint16_t)k;
            }
            if (!self)
                for (int k = tid; k < n2h; k += NT) {
                    const int c = cell_of(sx[offB + k], sy[offB + k], sz[offB + k]);
                    if (c >= 0) permB[atomicAdd(&cellB[c], 1u)] = (uint16_t)k;
                }
            __syncthreads();

            // ---- phase 6: pair search --------------------------------------------------------------
            for (;;) {
                int c = 0;
                if (lane == 0) c = atomicAdd(&s_next, 1);
                c = __shfl_sync(PC_FULL_MASK, c, 0);
                if (c >= ncell) break;
                const int a0 = c ? (int)cellA[c - 1] : 0, a1 = (int)cellA[c];
                if (a0 == a1) continue;
                const int cx = c % nx, cy = (c / nx) % ny, cz = c / (nx * ny);
                const int xlo = max(cx - 1, 0), xhi = min(cx + 1, nx - 1);
                for (int ab = a0; ab < a1; ab += 32) {
                    // this warp's chunk of A atoms (<= 32) lives in lane registers
                    const int na = min(32, a1 - ab);
                    int ka = 0;
                    float ax = 0.f, ay = 0.f, az = 0.f;
                    if (lane < na) { ka = permA[ab + lane]; ax = sx[ka]; ay = sy[ka]; az = sz[ka]; }
                    int ares = 0, aseg = 0;
                    if (self && lane < na) { ares = __ldg(a.t.hres1 + ka); aseg = __ldg(a.t.hseg1 + ka); }
                    for (int dz = -1; dz <= 1; dz++) {
                        const int z = cz + dz;
                        if (z < 0 || z >= nz) continue;
                        for (int dy = -1; dy <= 1; dy++) {
                            const int y = cy + dy;
                            if (y < 0 || y >= ny) continue;
                            const int row = (z * ny + y) * nx;
                            const int b0 = (row + xlo) ? (int)cellB[row + xlo - 1] : 0, b1 = (int)cellB[row + xhi];
                            for (int base = b0; base < b1; base += 32) {
                                const int j = base + lane;
                                const bool valid = j < b1;
                                int kb = 0;
                                float bx = 0.f, by = 0.f, bz = 0.f;
                                if (valid) { kb = permB[j]; bx = sx[offB + kb]; by = sy[offB + kb]; bz = sz[offB + kb]; }
                                int bres = 0, bseg = 0;
                                if (self && valid) { bres = __ldg(a.t.hres1 + kb); bseg = __ldg(a.t.hseg1 + kb); }
                                if (lane == 0) evals += (unsigned long long)na * (unsigned)min(32, b1 - base);
                                for (int ia = 0; ia < na; ia++) {
                                    const float px = __shfl_sync(PC_FULL_MASK, ax, ia);
                                    const float py = __shfl_sync(PC_FULL_MASK, ay, ia);
                                    const float pz = __shfl_sync(PC_FULL_MASK, az, ia);
                                    bool hit = valid && !(dist2_exact(px, py, pz, bx, by, bz) > a.sqdist);
                                    if (self) {
                                        // (resid1 - resid2) < 5 and same segment -> dropped (core/multi_trajectory.py:93-95)
                                        const int r1 = __shfl_sync(PC_FULL_MASK, ares, ia);
                                        const int s1 = __shfl_sync(PC_FULL_MASK, aseg, ia);
                                        if (hit && (r1 - bres) < 5 && s1 == bseg) hit = false;
                                    }
                                    const unsigned m = __ballot_sync(PC_FULL_MASK, hit);
                                    if (m) {
                                        unsigned pos = 0;
                                        if (lane == 0) pos = atomicAdd(&s_nhits, (unsigned)__popc(m));
                                        pos = __shfl_sync(PC_FULL_MASK, pos, 0);
                                        const int kA = __shfl_sync(PC_FULL_MASK, ka, ia);
                                        if (hit) {
                                            const unsigned p = pos + __popc(m & ((1u << lane) - 1u));
                                            if (p < (unsigned)a.capHits) hits[p] = ((unsigned)kA << 16) | (unsigned)kb;
                                        }
                                    }
                                }
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();

        // ---- phase 7: drain -----------------------------------------------------------------------
        unsigned nh = s_nhits;
        if (nh > (unsigned)a.capHits) { nh = a.capHits; if (tid == 0) atomicOr(&s_status, (unsigned)PC_ST_HITS_OVERFLOW); }
        if (tid == 0) {
            long long base = (long long)atomicAdd(&a.counters[PC_CNT_CONTACTS], (unsigned long long)nh);
            if (base + nh > a.cap_contacts) { atomicOr(&s_status, (unsigned)PC_ST_OUT_OVERFLOW); base = -1; }
            s_cbase = base;
        }
        __syncthreads();
        const long long cbase = s_cbase;
        for (unsigned h = tid; h < nh; h += NT) {
            const unsigned u = hits[h];
            const int ka = (int)(u >> 16), kb = (int)(u & 0xffffu);
            const int li = __ldg(a.t.heavy1 + ka), lj = __ldg(heavyB + kb);
            const double ax = sx[ka], ay = sy[ka], az = sz[ka];
            const double bx = sx[offB + kb], by = sy[offB + kb], bz = sz[offB + kb];
            const double dx = ax - bx, dy = ay - by, dz = az - bz;
            const double dist = sqrt(dx * dx + dy * dy + dz * dz);
            if (cbase >= 0) {
                pc_contact rec;
                rec.i = li; rec.j = lj; rec.distance = (float)dist; rec.weight = (float)contact_weight(dist);
                a.contacts[cbase + h] = rec;
            }
            // H-bond gate: both names start with O/N/S (core/multi_trajectory.py:112)
            if ((__ldg(a.t.hflag1 + ka) & PC_ATOM_HBCLASS) && (__ldg(hflagB + kb) & PC_ATOM_HBCLASS)) {
                for (int side = 0; side < 2; side++) {
                    const int p0 = side == 0 ? __ldg(a.t.hhptr1 + ka) : __ldg(hhptrB + kb);
                    const int p1 = side == 0 ? __ldg(a.t.hhptr1 + ka + 1) : __ldg(hhptrB + kb + 1);
                    for (int p = p0; p < p1; p++) {
                        const int hl = side == 0 ? __ldg(a.t.hidx1 + p) : __ldg(hidxB + p);
                        if (hl < 0) { atomicOr(&s_status, (unsigned)PC_ST_MISSING_H); continue; }
                        float hxf, hyf, hzf;
                        load_xyz(side == 0 ? fr1 : fr2, hl, a.stride, hxf, hyf, hzf);
                        double d = 0, ang = 0;
                        // side 0: donor = atom1, acceptor = atom2; side 1: donor = atom2, acceptor = atom1
                        const bool ok = side == 0
                            ? hbond_test(hxf, hyf, hzf, ax, ay, az, bx, by, bz, a.hb_cutoff, a.hb_angle, d, ang)
                            : hbond_test(hxf, hyf, hzf, bx, by, bz, ax, ay, az, a.hb_cutoff, a.hb_angle, d, ang);
                        if (ok) {
                            const unsigned q = atomicAdd(&s_nhb, 1u);
                            if (q < (unsigned)a.capHb) {
                                pc_hbond r;
                                r.i = li; r.j = lj; r.hyd = (uint32_t)hl | (side ? 0x80000000u : 0u);
                                r.distance = (float)d; r.angle = (float)ang;
                                hbq[q] = r;
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();
        unsigned nb = s_nhb;
        if (nb > (unsigned)a.capHb) { nb = a.capHb; if (tid == 0) atomicOr(&s_status, (unsigned)PC_ST_HB_OVERFLOW); }
        if (tid == 0) {
            long long base = (long long)atomicAdd(&a.counters[PC_CNT_HBONDS], (unsigned long long)nb);
            if (base + nb > a.cap_hbonds) { atomicOr(&s_status, (unsigned)PC_ST_OUT_OVERFLOW); base = -1; }
            s_hbase = base;
            a.frame_cstart[f] = (uint32_t)(cbase < 0 ? 0 : cbase);
            a.frame_ccount[f] = cbase < 0 ? 0u : nh;
            a.frame_hstart[f] = (uint32_t)(base < 0 ? 0 : base);
            a.frame_hcount[f] = base < 0 ? 0u : nb;
        }
        __syncthreads();
        if (s_hbase >= 0)
            for (unsigned q = tid; q < nb; q += NT) a.hbonds[s_hbase + q] = hbq[q];
        if (tid == 0 && s_status) atomicOr(&a.counters[PC_CNT_STATUS], (unsigned long long)s_status);
        __syncthreads();
    }
    // one evaluation counter per warp -> global
    evals = __shfl_sync(PC_FULL_MASK, evals, 0);
    if (lane == 0 && evals) atomicAdd(&a.counters[PC_CNT_EVALS], evals);
}
