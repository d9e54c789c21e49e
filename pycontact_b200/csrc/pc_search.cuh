// pc_search.cuh -- compatibility entry for cy_find_contacts (cy_modules/cy_gridsearch.pyx:31-35 ->
// find_contacts / vmd_gridsearch3, cy_modules/src/gridsearch.C:408-427, 902-1162): ALL atoms of both
// sets, no filters, CSR output in the reference's row order.
//
// Unlike the frame-batch paths this one uses the reference's own grid (joint bounding box origin,
// cell edge = (float)cutoff grown x1.26 while > 4e6 boxes, clamped box coordinates), because the
// order of a row's neighbours is defined by it: neighbour boxes in make_neighborlist_sym order
// (gridsearch.C:858-885), atoms of a box in index order.  B is binned with a counting sort whose
// boxes are then sorted by index (restoring the reference's append order); each A atom walks its
// <= 27 neighbour boxes in table order twice (count, then fill), so rows come out already ordered.
#pragma once
#include <string>
#include <vector>
#include "pc_common.cuh"

struct PcRefGrid {
    float min[3];
    float inv;
    int nb[3];
    int totb;
};

__constant__ signed char PC_STENCIL_OFF[27][3] = {
    {0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 1, 0}, {1, 0, 1}, {0, 1, 1}, {1, -1, 0}, {-1, 0, 1},
    {0, -1, 1}, {1, 1, 1}, {-1, 1, 1}, {1, -1, 1}, {-1, -1, 1}, {-1, 0, 0}, {0, -1, 0}, {0, 0, -1},
    {-1, -1, 0}, {-1, 0, -1}, {0, -1, -1}, {-1, 1, 0}, {1, 0, -1}, {0, 1, -1}, {-1, -1, -1}, {1, -1, -1},
    {-1, 1, -1}, {1, 1, -1}};

__device__ __forceinline__ void pc_ref_box(const PcRefGrid &g, float x, float y, float z, int &bx, int &by, int &bz) {
    // gridsearch.C:1009-1016: (int)((loc - min) * invpairdist), clamped to nb-1
    bx = (int)(__fmul_rn(__fsub_rn(x, g.min[0]), g.inv));
    by = (int)(__fmul_rn(__fsub_rn(y, g.min[1]), g.inv));
    bz = (int)(__fmul_rn(__fsub_rn(z, g.min[2]), g.inv));
    if (bx >= g.nb[0]) bx = g.nb[0] - 1;
    if (by >= g.nb[1]) by = g.nb[1] - 1;
    if (bz >= g.nb[2]) bz = g.nb[2] - 1;
}

__global__ void __launch_bounds__(256) pc_search_bbox_kernel(const float *xyz, int n, unsigned int *enc) {
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float x = xyz[3ll * i], y = xyz[3ll * i + 1], z = xyz[3ll * i + 2];
        mn[0] = fminf(mn[0], x); mn[1] = fminf(mn[1], y); mn[2] = fminf(mn[2], z);
        mx[0] = fmaxf(mx[0], x); mx[1] = fmaxf(mx[1], y); mx[2] = fmaxf(mx[2], z);
    }
#pragma unroll
    for (int k = 0; k < 3; k++) { mn[k] = warp_min(mn[k]); mx[k] = warp_max(mx[k]); }
    if ((threadIdx.x & 31) == 0)
        for (int k = 0; k < 3; k++) {
            auto encf = [](float f) { unsigned b = __float_as_uint(f); return (b & 0x80000000u) ? ~b : (b | 0x80000000u); };
            atomicMin(&enc[k], encf(mn[k]));
            atomicMax(&enc[3 + k], encf(mx[k]));
        }
}

__global__ void __launch_bounds__(256) pc_search_count_kernel(const float *xyz, int n, PcRefGrid g, uint32_t *cells,
                                                              uint32_t *box_of) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int bx, by, bz;
    pc_ref_box(g, xyz[3ll * i], xyz[3ll * i + 1], xyz[3ll * i + 2], bx, by, bz);
    const uint32_t b = (uint32_t)((bz * g.nb[1] + by) * g.nb[0] + bx);
    box_of[i] = b;
    atomicAdd(&cells[b], 1u);
}

// single-CTA exclusive scan, arr[n] = total
__global__ void __launch_bounds__(1024) pc_search_scan_kernel(uint32_t *arr, int n) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    constexpr int PER = 4;
    for (int base = 0; base < n; base += 1024 * PER) {
        uint32_t v[PER], s = 0;
#pragma unroll
        for (int q = 0; q < PER; q++) { const int i = base + tid * PER + q; v[q] = i < n ? arr[i] : 0; s += v[q]; }
        const uint32_t incl = warp_incl_scan(s, lane);
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) { const uint32_t w = wsum[lane]; const uint32_t wi = warp_incl_scan(w, lane); wsum[lane] = wi - w; }
        __syncthreads();
        uint32_t off = carry + wsum[warp] + incl - s;
#pragma unroll
        for (int q = 0; q < PER; q++) { const int i = base + tid * PER + q; if (i < n) arr[i] = off; off += v[q]; }
        __syncthreads();
        if (tid == 1023) carry = off;
        __syncthreads();
    }
    if (tid == 0) arr[n] = carry;
}

__global__ void __launch_bounds__(256) pc_search_scatter_kernel(int n, const uint32_t *box_of, uint32_t *cursor,
                                                                int32_t *items) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    items[atomicAdd(&cursor[box_of[i]], 1u)] = i;
}

// restore index order inside every box (the reference appends atoms in index order, gridsearch.C:1047-1074)
__global__ void __launch_bounds__(256) pc_search_sortbox_kernel(int totb, const uint32_t *start, int32_t *items) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= totb) return;
    const int lo = (int)start[b], hi = (int)start[b + 1];
    for (int i = lo + 1; i < hi; i++) {
        const int32_t v = items[i];
        int j = i - 1;
        while (j >= lo && items[j] > v) { items[j + 1] = items[j]; j--; }
        items[j + 1] = v;
    }
}

template <bool FILL>
__global__ void __launch_bounds__(128) pc_search_pairs_kernel(const float *xyzA, int nA, const float *xyzB, PcRefGrid g,
                                                              float sqdist, const uint32_t *startB, const int32_t *itemsB,
                                                              unsigned long long *row_ptr, int32_t *cols) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nA) return;
    const float ax = xyzA[3ll * i], ay = xyzA[3ll * i + 1], az = xyzA[3ll * i + 2];
    int bx, by, bz;
    pc_ref_box(g, ax, ay, az, bx, by, bz);
    unsigned long long out = FILL ? row_ptr[i] : 0ull;
    for (int s = 0; s < 27; s++) {
        const int x = bx + PC_STENCIL_OFF[s][0], y = by + PC_STENCIL_OFF[s][1], z = bz + PC_STENCIL_OFF[s][2];
        if (x < 0 || y < 0 || z < 0 || x >= g.nb[0] || y >= g.nb[1] || z >= g.nb[2]) continue;
        const int nb = (z * g.nb[1] + y) * g.nb[0] + x;
        const int lo = (int)startB[nb], hi = (int)startB[nb + 1];
        for (int p = lo; p < hi; p++) {
            const int j = itemsB[p];
            if (dist2_exact(ax, ay, az, xyzB[3ll * j], xyzB[3ll * j + 1], xyzB[3ll * j + 2]) > sqdist) continue;
            if (FILL) cols[out] = j;
            out++;
        }
    }
    if (!FILL) row_ptr[i] = out;      // counts; scanned on the host into offsets
}

#define PCS_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { err = std::string(#expr) + ": " + cudaGetErrorString(_e); rc = PC_ERR_CUDA; goto done; } } while (0)

static inline int pc_search_all(const float *h1, int n1, const float *h2, int n2, double cutoff, int64_t *h_row_ptr,
                                int32_t **h_cols, std::string &err) {
    int rc = PC_OK;
    float *d1 = nullptr, *d2 = nullptr;
    unsigned int *d_enc = nullptr;
    uint32_t *d_cells = nullptr, *d_start = nullptr, *d_box = nullptr;
    int32_t *d_items = nullptr, *d_cols = nullptr;
    unsigned long long *d_rows = nullptr;
    std::vector<unsigned long long> rows((size_t)n1 + 1);
    unsigned int enc[6];
    PcRefGrid g;
    const float pairdist0 = (float)cutoff;            // gridsearch.C:417 -> :904
    const float sqdist = pairdist0 * pairdist0;       // gridsearch.C:936
    unsigned long long total = 0;
    {
        PCS_TRY(cudaMalloc(&d1, (size_t)n1 * 12));
        PCS_TRY(cudaMalloc(&d2, (size_t)n2 * 12));
        PCS_TRY(cudaMemcpy(d1, h1, (size_t)n1 * 12, cudaMemcpyHostToDevice));
        PCS_TRY(cudaMemcpy(d2, h2, (size_t)n2 * 12, cudaMemcpyHostToDevice));
        PCS_TRY(cudaMalloc(&d_enc, 6 * sizeof(unsigned int)));
        const unsigned int init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
        PCS_TRY(cudaMemcpy(d_enc, init, sizeof init, cudaMemcpyHostToDevice));
        pc_search_bbox_kernel<<<std::min((n1 + 255) / 256, 1184), 256>>>(d1, n1, d_enc);
        pc_search_bbox_kernel<<<std::min((n2 + 255) / 256, 1184), 256>>>(d2, n2, d_enc);
        PCS_TRY(cudaMemcpy(enc, d_enc, sizeof enc, cudaMemcpyDeviceToHost));
        float lo[3], hi[3];
        for (int k = 0; k < 6; k++) {
            const unsigned int e = enc[k];
            const unsigned int b = (e & 0x80000000u) ? (e & 0x7fffffffu) : ~e;
            float f;
            memcpy(&f, &b, 4);
            (k < 3 ? lo[k] : hi[k - 3]) = f;
        }
        // grid sizing exactly as gridsearch.C:969-985 (scalar float arithmetic on the host)
        const int MAXBOXES = 4000000;
        volatile float pairdist = pairdist0, next = pairdist0;
        const float xr = hi[0] - lo[0], yr = hi[1] - lo[1], zr = hi[2] - lo[2];
        int xb, yb, zb, totb;
        do {
            pairdist = next;
            const float inv = 1.0f / pairdist;
            xb = ((int)(xr * inv)) + 1; yb = ((int)(yr * inv)) + 1; zb = ((int)(zr * inv)) + 1;
            totb = xb * yb * zb;
            next = pairdist * 1.26f;
        } while (totb > MAXBOXES || totb < 1);
        g.min[0] = lo[0]; g.min[1] = lo[1]; g.min[2] = lo[2];
        g.inv = 1.0f / pairdist; g.nb[0] = xb; g.nb[1] = yb; g.nb[2] = zb; g.totb = totb;

        PCS_TRY(cudaMalloc(&d_cells, ((size_t)totb + 1) * 4));
        PCS_TRY(cudaMalloc(&d_start, ((size_t)totb + 1) * 4));
        PCS_TRY(cudaMalloc(&d_box, (size_t)n2 * 4));
        PCS_TRY(cudaMalloc(&d_items, (size_t)n2 * 4));
        PCS_TRY(cudaMemset(d_cells, 0, ((size_t)totb + 1) * 4));
        pc_search_count_kernel<<<(n2 + 255) / 256, 256>>>(d2, n2, g, d_cells, d_box);
        pc_search_scan_kernel<<<1, 1024>>>(d_cells, totb);
        PCS_TRY(cudaMemcpy(d_start, d_cells, ((size_t)totb + 1) * 4, cudaMemcpyDeviceToDevice));
        pc_search_scatter_kernel<<<(n2 + 255) / 256, 256>>>(n2, d_box, d_cells, d_items);
        pc_search_sortbox_kernel<<<(totb + 255) / 256, 256>>>(totb, d_start, d_items);
        PCS_TRY(cudaMalloc(&d_rows, ((size_t)n1 + 1) * 8));
        pc_search_pairs_kernel<false><<<(n1 + 127) / 128, 128>>>(d1, n1, d2, g, sqdist, d_start, d_items, d_rows, nullptr);
        PCS_TRY(cudaGetLastError());
        PCS_TRY(cudaMemcpy(rows.data(), d_rows, (size_t)n1 * 8, cudaMemcpyDeviceToHost));
        for (int i = 0; i < n1; i++) { const unsigned long long c = rows[i]; rows[i] = total; h_row_ptr[i] = (int64_t)total; total += c; }
        rows[n1] = total; h_row_ptr[n1] = (int64_t)total;
        int32_t *cols = (int32_t *)malloc(std::max<size_t>(total, 1) * sizeof(int32_t));
        if (!cols) { err = "pc_find_contacts: out of host memory"; rc = PC_ERR_NOMEM; goto done; }
        *h_cols = cols;
        if (total) {
            PCS_TRY(cudaMemcpy(d_rows, rows.data(), ((size_t)n1 + 1) * 8, cudaMemcpyHostToDevice));
            PCS_TRY(cudaMalloc(&d_cols, total * 4));
            pc_search_pairs_kernel<true><<<(n1 + 127) / 128, 128>>>(d1, n1, d2, g, sqdist, d_start, d_items, d_rows, d_cols);
            PCS_TRY(cudaGetLastError());
            PCS_TRY(cudaMemcpy(cols, d_cols, total * 4, cudaMemcpyDeviceToHost));
        }
    }
done:
    {
        void *ps[] = {d1, d2, d_enc, d_cells, d_start, d_box, d_items, d_cols, d_rows};
        for (void *p : ps) if (p) cudaFree(p);
    }
    if (rc != PC_OK && *h_cols) { free(*h_cols); *h_cols = nullptr; }
    return rc;
}
