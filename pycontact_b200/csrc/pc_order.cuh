// pc_order.cuh -- the reference's in-frame contact order as a sortable 64-bit key (optional pass).
//
// contactResults[f] lists a frame's contacts in the order the reference's nested loops emit them
// (core/multi_trajectory.py:83-84 over the rows cy_find_contacts returns): idx1 ascending, and inside
// one idx1 the partners in (neighbour-box rank, idx2 ascending) order, where the rank is the position
// of box(idx2) - box(idx1) in make_neighborlist_sym's table (cy_modules/src/gridsearch.C:858-885) on
// the reference's own grid: origin = joint minimum over ALL atoms of both selections (hydrogens too,
// gridsearch.C:940-949), edge = (float)cutoff grown x1.26 while there are more than 4e6 boxes
// (:969-985), box coordinate (int)((x - min) * inv) clamped to nb-1 (:1009-1016).
//
//     key = idx1 << 37 | rank << 32 | idx2            (idx1 < 2^27, rank < 27)
//
// Sorting a frame's contacts by this key reproduces the reference's list order, and the smallest key
// of a (frame, accumulation key) group is its first appearance (core/ContactAnalyzer.py:555-556).
// The scan kernels use their own, tighter grids; this pass only labels what they found.
#pragma once
#include "pc_common.cuh"

struct PcOrderGrid {
    float min[3];
    float inv;
    int nb[3];
};

template <int NT>
__global__ void __launch_bounds__(NT) pc_order_kernel(const PcScanArgs a) {
    __shared__ float red[NT / 32][6];
    __shared__ PcOrderGrid g;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool self = a.self_mode != 0;
    for (int f = blockIdx.x; f < a.nframes; f += gridDim.x) {
        const float *fr1 = a.xyz1 + (long long)f * a.pitch1;
        const float *fr2 = self ? fr1 : a.xyz2 + (long long)f * a.pitch2;
        float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int side = 0; side < (self ? 1 : 2); side++) {
            const float *fr = side ? fr2 : fr1;
            const int n = side ? a.t.n2 : a.t.n1;
            for (int i = tid; i < n; i += NT) {
                float x, y, z;
                load_xyz(fr, i, a.stride, x, y, z);
                mn[0] = fminf(mn[0], x); mn[1] = fminf(mn[1], y); mn[2] = fminf(mn[2], z);
                mx[0] = fmaxf(mx[0], x); mx[1] = fmaxf(mx[1], y); mx[2] = fmaxf(mx[2], z);
            }
        }
#pragma unroll
        for (int k = 0; k < 3; k++) { mn[k] = warp_min(mn[k]); mx[k] = warp_max(mx[k]); }
        if (lane == 0) for (int k = 0; k < 3; k++) { red[warp][k] = mn[k]; red[warp][3 + k] = mx[k]; }
        __syncthreads();
        if (tid == 0) {
            float lo[3], hi[3];
            for (int k = 0; k < 3; k++) {
                float l = red[0][k], h = red[0][3 + k];
                for (int w = 1; w < NT / 32; w++) { l = fminf(l, red[w][k]); h = fmaxf(h, red[w][3 + k]); }
                lo[k] = l; hi[k] = h;
            }
            // gridsearch.C:969-985, float32 throughout
            const float xr = __fsub_rn(hi[0], lo[0]), yr = __fsub_rn(hi[1], lo[1]), zr = __fsub_rn(hi[2], lo[2]);
            float pairdist = a.cutoff_f, next = a.cutoff_f, inv = 0.f;
            int xb, yb, zb;
            long long tot;
            int guard = 0;
            do {
                pairdist = next;
                inv = __fdiv_rn(1.0f, pairdist);
                xb = (int)__fmul_rn(xr, inv) + 1; yb = (int)__fmul_rn(yr, inv) + 1; zb = (int)__fmul_rn(zr, inv) + 1;
                tot = (long long)xb * yb * zb;
                next = __fmul_rn(pairdist, 1.26f);
            } while ((tot > 4000000ll || tot < 1) && ++guard < 256);
            g.min[0] = lo[0]; g.min[1] = lo[1]; g.min[2] = lo[2];
            g.inv = inv; g.nb[0] = xb; g.nb[1] = yb; g.nb[2] = zb;
        }
        __syncthreads();
        const uint32_t c0 = a.frame_cstart[f], cn = a.frame_ccount[f];
        for (uint32_t c = tid; c < cn; c += NT) {
            const pc_contact rec = a.contacts[c0 + c];
            float p[3], q[3];
            load_xyz(fr1, rec.i, a.stride, p[0], p[1], p[2]);
            load_xyz(fr2, rec.j, a.stride, q[0], q[1], q[2]);
            int d[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                int bi = (int)__fmul_rn(__fsub_rn(p[k], g.min[k]), g.inv);
                int bj = (int)__fmul_rn(__fsub_rn(q[k], g.min[k]), g.inv);
                if (bi >= g.nb[k]) bi = g.nb[k] - 1;
                if (bj >= g.nb[k]) bj = g.nb[k] - 1;
                d[k] = bj - bi;
            }
            // A pair the reference finds sits in adjacent boxes of ITS grid.  The scan's own cells are 2e-4 wider than
            // the cutoff, so it can in principle accept a pair whose reference boxes are two apart on one axis: that
            // needs |dx| within ~1e-6 (relative) of the cutoff, dy and dz ~ 0 and the two box coordinates rounding to
            // opposite sides of two box boundaries -- a pair the reference's stencil would not visit.  Never observed
            // (the parity tests compare pair SETS with the reference on every fixture and synthetic system); such a
            // pair gets rank 31, so it sorts behind its atom's real partners instead of aliasing another offset's rank.
            const bool adjacent = d[0] >= -1 && d[0] <= 1 && d[1] >= -1 && d[1] <= 1 && d[2] >= -1 && d[2] <= 1;
            const int si = (d[0] + 1) + 3 * (d[1] + 1) + 9 * (d[2] + 1);
            const unsigned rank = adjacent ? PC_STENCIL_RANK[si] : 31u;
            a.order_out[c0 + c] = ((unsigned long long)(unsigned)rec.i << 37) | ((unsigned long long)rank << 32) |
                                  (unsigned long long)(unsigned)rec.j;
        }
        __syncthreads();
    }
}
