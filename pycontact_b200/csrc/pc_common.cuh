// pc_common.cuh -- shared device helpers and argument blocks (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/pycontact_b200.h"

#define PC_FULL_MASK 0xffffffffu

// status bits written by kernels into counters[PC_CNT_STATUS]
#define PC_ST_HITS_OVERFLOW   1ull   // per-frame shared hit queue too small
#define PC_ST_HB_OVERFLOW     2ull   // per-frame shared H-bond staging too small
#define PC_ST_OUT_OVERFLOW    4ull   // batch-wide contact / hbond / row buffer too small
#define PC_ST_MISSING_H       8ull   // bonded hydrogen not in the selection (reference raises)
#define PC_ST_BAD_COORDS     16ull   // non-finite coordinates
#define PC_ST_TABLE_OVERFLOW 32ull   // accumulation hash table too small
#define PC_ST_STAGE_OVERFLOW 64ull   // small-frame path: more atoms inside the grid than the staging arrays hold
#define PC_ST_TAIL_OVERFLOW 128ull   // large-frame path: a frame's in-grid atoms do not fit the fused per-frame tail kernel
#define PC_ST_PRECISION     256ull   // a (frame, key) cell holds more contacts than a float32 sum keeps within the 1e-5 bar

// any of these: the batch's outputs are incomplete and the caller re-runs it with larger buffers / another path
#define PC_ST_RERUN_MASK (PC_ST_HITS_OVERFLOW | PC_ST_HB_OVERFLOW | PC_ST_OUT_OVERFLOW | PC_ST_TABLE_OVERFLOW | \
                          PC_ST_STAGE_OVERFLOW | PC_ST_TAIL_OVERFLOW | PC_ST_PRECISION)

// float32 per-cell sums (small-frame kernel's fused table, compact table): up to this many weights <= 1 per cell the
// rounding of the running sum stays ~sqrt(n) * 6e-8 relative (<= 2e-6); cells beyond it are redone in float64 /
// exact fixed point on the global-table path
#define PC_FLOAT_SUM_MAX 1024u

#ifdef PC_PHASE_CLOCKS
// profiling build only (tools/phase_clocks.py): cycles per phase of the small-frame kernel, summed over CTAs
enum { PC_CNT_CONTACTS = 0, PC_CNT_HBONDS = 1, PC_CNT_STATUS = 2, PC_CNT_EVALS = 3, PC_CNT_ROWS = 4,
       PC_CNT_WORK = 5, PC_CNT_PHASE = 8, PC_CNT_COUNT = 32 };
#define PC_PHASE_BEGIN() long long pc_t_prev = clock64()
#define PC_PHASE_MARK(i) do { if (threadIdx.x == 0) { const long long pc_t = clock64(); \
        atomicAdd(&a.counters[PC_CNT_PHASE + (i)], (unsigned long long)(pc_t - pc_t_prev)); pc_t_prev = pc_t; } } while (0)
#else
enum { PC_CNT_CONTACTS = 0, PC_CNT_HBONDS = 1, PC_CNT_STATUS = 2, PC_CNT_EVALS = 3, PC_CNT_ROWS = 4,
       PC_CNT_WORK = 5, PC_CNT_COUNT = 8 };
#define PC_PHASE_BEGIN()
#define PC_PHASE_MARK(i)
#endif

// cell edge = cutoff * PC_EDGE_SCALE.  A pair accepted by the reference's float32 test has a
// true separation <= cutoff*(1+3e-7); cell coordinates computed in float32 carry an error
// below 3e-5 cells for extents up to 1000 A, so with 2e-4 of slack two such atoms always
// land in the same or adjacent cells (DESIGN.md "grid exactness").
#define PC_EDGE_SCALE 1.0002f

// Cells per axis of a grid over a box of extent (ex, ey, ez): edge `edge`, grown x1.26 until the grid fits `cap`
// cells (the reference grows its boxes the same way, gridsearch.C:969-985).  Guard of the slack argument above: a
// cell coordinate u = (x - lo) * inv carries up to 2 * u * 2^-23 cells of float32 rounding, which the 2e-4 of
// PC_EDGE_SCALE covers up to ~800 cells per axis (4000 A at a 5 A cutoff); grids with a longer axis get cells
// wider by 5e-7 per cell of that axis, so accepted pairs still always sit in the same or adjacent cells.
__device__ __forceinline__ void pc_grid_dims(const float ex, const float ey, const float ez, const float edge, const long long cap,
                                             int &nx, int &ny, int &nz, float &inv) {
    float e = edge;
    bool widened = false;
    for (;;) {
        inv = 1.0f / e;
        const long long cx = (long long)(ex * inv) + 1, cy = (long long)(ey * inv) + 1, cz = (long long)(ez * inv) + 1;
        const long long longest = cx > cy ? (cx > cz ? cx : cz) : (cy > cz ? cy : cz);
        if (!widened && longest > 512) { e = e * (1.0f + 5e-7f * (float)longest); widened = true; continue; }
        if (cx * cy * cz <= cap) { nx = (int)cx; ny = (int)cy; nz = (int)cz; return; }
        e *= 1.26f;
    }
}

// Packed atom word carried in the .w of staged float4 coordinates and in PcTopoDev::hinfo:
//   bits 0..26 local index in the selection, bit 30 backbone, bit 31 name[0] in {O,N,S}
#define PC_W_INDEX(w)   ((int)((w) & 0x07ffffffu))
#define PC_W_BACKBONE   0x40000000u
#define PC_W_HBCLASS    0x80000000u

struct PcTopoDev {
    int n1, n2, n1h, n2h;
    const uint32_t *hinfo1, *hinfo2;   // per heavy ordinal: packed atom word (heavy atoms in index order)
    const uint32_t *hmask1, *hmask2;   // 1 bit per LOCAL atom: not a hydrogen (streaming kernels)
    const uint32_t *hpre1, *hpre2;     // per 32 LOCAL atoms: heavy atoms before atom 32*w (heavy ordinal of a range start)
    const uint32_t *linfo1, *linfo2;   // per LOCAL atom: packed atom word
    const int *lres1, *lseg1, *lres2, *lseg2;          // per LOCAL atom (self-interaction filter)
    const int *lhptr1, *lhidx1, *lhptr2, *lhidx2;      // bonded-H CSR per LOCAL atom -> local H index (-1: missing)
    const int *lgid1, *lgid2;          // per LOCAL atom group ids (maps)
    const uint8_t *lflag1, *lflag2;    // per LOCAL atom PC_ATOM_* flags
    unsigned long long ngroups2;
};

struct PcScanArgs {
    const float *xyz1, *xyz2;
    long long pitch1, pitch2;          // floats per frame
    int stride;                        // floats per atom (3 or 4)
    int nframes, self_mode, order_keys;
    PcTopoDev t;
    float edge, sqdist, cutoff_f;
    double hb_cutoff, hb_angle;
    int capCells, capHits, capHb;      // shared-memory capacities (small-frame path)
    int fuse_acc;                      // small-frame path: accumulate (frame, key) rows inside the scan kernel
    int dense_mode;                    // large-frame path: 0 = pair kernel chosen per frame, 1 = thread per atom, 2 = warp per cell
    pc_row *rows;
    long long cap_rows;
    uint32_t *frame_rstart, *frame_rcount;
    pc_contact *contacts;
    pc_hbond *hbonds;
    unsigned long long *order_out;
    unsigned long long *counters;
    long long cap_contacts, cap_hbonds;
    uint32_t *frame_cstart, *frame_ccount, *frame_hstart, *frame_hcount;
};

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(PC_FULL_MASK, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(PC_FULL_MASK, v, o));
    return v;
}
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t n = __shfl_up_sync(PC_FULL_MASK, v, o);
        if (lane >= o) v += n;
    }
    return v;
}

// The reference's pair test (gridsearch.C:48-55, :1126): float32, ((dx*dx + dy*dy) + dz*dz),
// every operation rounded separately -- the _rn intrinsics are never contracted into FMAs.
__device__ __forceinline__ float dist2_exact(float ax, float ay, float az, float bx, float by, float bz) {
    const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// weight_function (core/multi_trajectory.py:15-17), float64 like the reference
__device__ __forceinline__ double contact_weight(double d) { return 1.0 / (1.0 + exp(5.0 * (d - 4.0))); }

// Same function for the float32 records: the exponent argument is formed in float64 from the
// float64 distance and only then narrowed, so the relative error of the result stays below
// |t| * 2^-24 + 3 ulp (< 3e-6 up to a 12 A cutoff) -- inside the 1e-5 bar, at a fraction of the
// float64 exp's cost.
__device__ __forceinline__ float contact_weight_f32(double d) {
    const float t = (float)(5.0 * (d - 4.0));
    return __fdividef(1.0f, 1.0f + expf(t));
}

// exp(t) overflows float for t > 88 (d > 21.7 A): the weight is then < 1e-38 and rounds to 0
// either way.

// The reference's pair test with one fused evaluation in front of it: d2f (2 FFMA) differs from
// the reference's non-fused value by at most 4 ulp, so pairs with d2f > sqdist*(1+2^-20) are
// certainly outside; everything else (all hits, plus a sliver of near misses) is decided by
// dist2_exact.  Same set as the reference, ~30% fewer instructions per candidate.
__device__ __forceinline__ bool pair_within(float ax, float ay, float az, float bx, float by, float bz,
                                            float sqdist, float sqdist_hi);

__device__ __forceinline__ void load_xyz(const float *frame, int idx, int stride, float &x, float &y, float &z) {
    if (stride == 4) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(frame) + idx);
        x = v.x; y = v.y; z = v.z;
    } else {
        const float *p = frame + 3ll * idx;
        x = __ldg(p); y = __ldg(p + 1); z = __ldg(p + 2);
    }
}

// One donor-H...acceptor test (core/multi_trajectory.py:157-181 / :183-205), float64 on the
// exactly promoted float32 coordinates.  Returns true and fills dist/angle when it is an H-bond.
__device__ __forceinline__ bool hbond_test(double hx, double hy, double hz, double dx_, double dy_, double dz_,
                                           double ax, double ay, double az, double hb_cutoff, double hb_angle,
                                           double &dist, double &angle) {
    const double v1x = hx - ax, v1y = hy - ay, v1z = hz - az;   // H - acceptor
    const double n1 = sqrt(v1x * v1x + v1y * v1y + v1z * v1z);
    if (!(n1 <= hb_cutoff)) return false;
    const double v2x = hx - dx_, v2y = hy - dy_, v2z = hz - dz_; // H - donor
    const double n2 = sqrt(v2x * v2x + v2y * v2y + v2z * v2z);
    const double dot = v1x * v2x + v1y * v2y + v1z * v2z;
    const double ang = acos(dot / (n1 * n2)) * 57.29577951308232;   // np.degrees
    if (!(ang >= hb_angle)) return false;
    dist = n1; angle = ang;
    return true;
}

// The reference's pair test on Blackwell's packed float32 pipe: (bx-ax, by-ay) in one FADD2, both squares in
// one FMUL2.  Every operation is a separately rounded IEEE float32 one like dist2_exact's; b - a is the
// exact negative of a - b (round-to-nearest is symmetric), so the squares, the sums and the decision are
// bit-identical to gridsearch.C:48-55.  nxy = (-ax, -ay), naz = -az.
__device__ __forceinline__ bool pair_within_packed(const float2 nxy, const float naz, const float4 pb, const float sqdist) {
    const float2 dxy = __fadd2_rn(make_float2(pb.x, pb.y), nxy);
    const float dz = __fadd_rn(pb.z, naz);
    const float2 sxy = __fmul2_rn(dxy, dxy);
    const float sz = __fmul_rn(dz, dz);
    return !(__fadd_rn(__fadd_rn(sxy.x, sxy.y), sz) > sqdist);
}

// stencil rank of a box offset in make_neighborlist_sym order (gridsearch.C:858-885, App. A.3);
// index = (dx+1) + 3*(dy+1) + 9*(dz+1)
__constant__ unsigned char PC_STENCIL_RANK[27] = {
    /* dz=-1 */ 23, 19, 24,  18, 16, 21,  25, 22, 26,
    /* dz= 0 */ 17, 15,  7,  14,  0,  1,  20,  2,  4,
    /* dz=+1 */ 13,  9, 12,   8,  3,  5,  11,  6, 10};

__device__ __forceinline__ bool pair_within(float ax, float ay, float az, float bx, float by, float bz,
                                            float sqdist, float sqdist_hi) {
    const float dx = ax - bx, dy = ay - by, dz = az - bz;
    const float d2f = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    if (d2f > sqdist_hi) return false;
    return !(dist2_exact(ax, ay, az, bx, by, bz) > sqdist);
}

// Self-interaction filter on a hit mask (core/multi_trajectory.py:93-95: a pair is dropped when
// (resid1 - resid2) < 5, signed, and both atoms are in the same segment).  word(k) = packed atom word of
// candidate k.  Four hits per round with all their topology loads issued before the first test: written as
// `a && b` per hit, the segment load waited for the residue load, and each hit for the previous one.
template <typename WordOf>
__device__ __forceinline__ unsigned pc_self_filter(unsigned m, int ares, int aseg, const int *lres, const int *lseg, WordOf word) {
    unsigned t = m;
    while (t) {
        int k[4], r[4], g[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            k[u] = t ? __ffs(t) - 1 : k[0];
            t &= t - 1;                                    // 0 stays 0
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int lj = PC_W_INDEX(word(k[u]));
            r[u] = __ldg(lres + lj); g[u] = __ldg(lseg + lj);
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
            if ((ares - r[u]) < 5 && aseg == g[u]) m &= ~(1u << k[u]);
    }
    return m;
}

