// pc_api.cu -- C ABI of libpycontact_b200.so (see include/pycontact_b200.h).
//
// Host-side glue only: context, device buffers, kernel selection and launches.  All
// numerical work is in the kernels (pc_scan_small.cuh, pc_scan_large.cuh, pc_accumulate.cuh).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>

#include "pc_common.cuh"
#include "pc_scan_small.cuh"
#include "pc_scan_large.cuh"
#include "pc_accumulate.cuh"
#include "pc_scan_group.cuh"
#include "pc_scan_tail.cuh"
#include "pc_order.cuh"
#include "pc_search.cuh"
#include "pc_synth.cuh"
#include "pc_stats.cuh"
#include "pc_sasa.cuh"
#include "pc_peak.cuh"
#include "pc_xtc.h"

static thread_local std::string g_err;
static unsigned long long g_launches = 0;      // kernels launched by this library (bench.py's gpu_launches)

static int fail(int code, const std::string &msg) { g_err = msg; return code; }

#define CUDA_TRY(expr)                                                                     \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess)                                                             \
            return fail(PC_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    cudaError_t ensure(size_t count) {
        if (count <= n && p) return cudaSuccess;
        release();
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    cudaError_t upload(const std::vector<T> &v, cudaStream_t s = 0) {
        cudaError_t e = ensure(v.size());
        if (e != cudaSuccess || v.empty()) return e;
        return cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
    }
};

struct pc_ctx {
    int device = 0;
    double cutoff = 0, hb_cutoff = 0, hb_angle = 0;
    int self_mode = 0;
    int num_sms = 148;
    size_t smem_optin = 0;
    bool have_topology = false, have_maps = false;
    int n1 = 0, n2 = 0, n1h = 0, n2h = 0;
    DevBuf<uint32_t> hinfo1, hinfo2, hmask1, hmask2, hpre1, hpre2, linfo1, linfo2;
    DevBuf<int> lres1, lseg1, lres2, lseg2, lhptr1, lhidx1, lhptr2, lhidx2, lgid1, lgid2;
    DevBuf<uint8_t> lflag1, lflag2;
    DevBuf<uint32_t> lgb1, lgb2;      // group id | backbone << 31 (pc_set_maps)
    unsigned long long ngroups1 = 0, ngroups2 = 0;
    // limits (0 = automatic)
    int lim_hits = 0, lim_hb = 0, lim_cells = 0, lim_table = 0;
    int force_path = 0;               // 0 = automatic, 1 = small-frame kernel, 2 = large-frame path
    int tail_mode = 3;                // large-frame path: 3 = fused tail kernel, selection A read from the frame (default);
                                      // 1 = fused tail kernel on the compacted atoms of both selections; 2 = walk kernel;
                                      // 0 = multi-kernel sequence
    int dense_mode = 0;               // multi-kernel sequence: pair kernel 0 = per frame on the device, 1 = thread per atom, 2 = warp per cell
    // batch outputs
    int max_frames = 0;
    long long cap_contacts = 0, cap_hbonds = 0, cap_rows = 0;
    DevBuf<pc_contact> contacts;
    DevBuf<pc_hbond> hbonds;
    DevBuf<pc_row> rows;
    DevBuf<pc_row_packed> rows_packed;
    DevBuf<unsigned long long> order_keys;
    DevBuf<unsigned long long> counters;
    DevBuf<uint32_t> frame_meta;      // 6 arrays of max_frames
    DevBuf<float> stage1, stage2;     // H2D staging for pc_scan_host
    PcLargeWork large;                // work buffers of the large-frame path
    DevBuf<PcGlobalEntry> gacc_tab;        // global accumulation table
    DevBuf<pc_fold_entry> fold_tab, fold_out;      // sparse time-aggregated maps (pc_fold_rows_sparse)
    DevBuf<unsigned long long> fold_meta;
    unsigned long long fold_cap = 0;
    pc_ctx *fold_owner = nullptr;          // fold into another context's table (contexts of one GPU share one)
    unsigned long long *h_fold_meta = nullptr;
    DevBuf<uint32_t> gacc_cursor, gacc_fresh;
    DevBuf<unsigned long long> gacc_nfresh;
    unsigned long long gacc_hint = 0;      // contacts of the previous accumulated batch (table sizing without a sync)
    unsigned long long gacc_hint_rows = 0; // and its rows
    int gacc_hint_frames = 0;
    int last_nframes = 0;
    bool last_order_keys = false;
    unsigned long long *h_counters = nullptr;   // pinned mirror
    int last_path = 0;                // 1 = small, 2 = large
    bool gacc_used = false, gacc_retry = false;
    bool packed_valid = false;        // rows_packed holds the last accumulated batch
    int pack_bytes = 24;              // transfer format of pc_pack_rows: 24 (pc_row_packed) or 12 (pc_row_slim)
    void *probe[2] = {nullptr, nullptr};   // optional events around the large path's dominant kernel (pc_set_probe)
    DevBuf<int> retry;                // [0] count, [4..] frames the fast small-frame plan handed to the full plan
    int small_ctas = 0;               // small-frame kernel plans: 0 = fast (2 CTAs/SM) + full (automatic), 1 = full only, 2 = fast only,
                                      // 3 / 4 = fast plan with 3 x 384 / 4 x 256 threads per SM + full
    int last_small_ctas = 0;
    bool want_fused_acc = true;       // accumulate inside the small-frame scan kernel when possible
    bool last_fused = false;
    std::vector<uint32_t> h_meta;     // host copy of the scan's 4 frame arrays (large path: sliced outputs)
    // group kernel of the large-frame path (pc_scan_group.cuh): units of whole accumulation groups of selection 1
    std::vector<uint32_t> h_hinfo1;   // host copy of selection 1's heavy-atom list
    DevBuf<uint32_t> group_units;     // n_units + 1 heavy ordinals
    int n_units = 0;                  // 0: the maps' groups are not runs of the selection's order (or larger than a unit)
    int group_mode = 1;               // 1 = use the group kernel when it applies (pc_set_group_mode)
    DevBuf<pc_row> row_slices;        // nframes x (cap_rows / nframes), gathered into `rows`
    bool last_group = false;
};

extern "C" const char *pc_last_error(void) { return g_err.c_str(); }
extern "C" int pc_version(void) { return 100; }
extern "C" int pc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int pc_create(int device, double cutoff, double hb_cutoff, double hb_angle_deg, int self_mode,
                         pc_ctx **out) {
    if (!out) return fail(PC_ERR_INVALID, "pc_create: out is NULL");
    if (!(cutoff > 0)) return fail(PC_ERR_INVALID, "pc_create: cutoff must be positive");
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(PC_ERR_INVALID, "pc_create: no such CUDA device");
    CUDA_TRY(cudaSetDevice(device));
    pc_ctx *c = new pc_ctx();
    c->device = device; c->cutoff = cutoff; c->hb_cutoff = hb_cutoff; c->hb_angle = hb_angle_deg;
    c->self_mode = self_mode ? 1 : 0;
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    c->num_sms = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    CUDA_TRY(c->counters.ensure(PC_CNT_COUNT));
    CUDA_TRY(cudaMemset(c->counters.p, 0, PC_CNT_COUNT * sizeof(unsigned long long)));
    CUDA_TRY(cudaHostAlloc(&c->h_counters, PC_CNT_COUNT * sizeof(unsigned long long), cudaHostAllocDefault));
    *out = c;
    return PC_OK;
}

extern "C" void pc_destroy(pc_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->h_counters) cudaFreeHost(c->h_counters);
    if (c->h_fold_meta) cudaFreeHost(c->h_fold_meta);
    delete c;
}

extern "C" int pc_set_path(pc_ctx *c, int path) {
    if (!c || path < 0 || path > 2) return fail(PC_ERR_INVALID, "pc_set_path: path must be 0, 1 or 2");
    c->force_path = path;
    return PC_OK;
}

extern "C" int pc_set_tail(pc_ctx *c, int on) {
    if (!c) return fail(PC_ERR_INVALID, "pc_set_tail: ctx is NULL");
    if (on != 0 && on != 1 && on != 3)
        return fail(PC_ERR_INVALID, "pc_set_tail: 0 (multi-kernel sequence), 1 (fused tail kernel on the compacted atoms) or 3 (tail kernel, A from the frame)");
    c->tail_mode = on;
    return PC_OK;
}

extern "C" int pc_set_dense_mode(pc_ctx *c, int mode) {
    if (!c || mode < 0 || mode > 2) return fail(PC_ERR_INVALID, "pc_set_dense_mode: 0 (per frame), 1 (thread per atom) or 2 (warp per cell)");
    c->dense_mode = mode;
    return PC_OK;
}

extern "C" int pc_set_group_mode(pc_ctx *c, int on) {
    if (!c) return fail(PC_ERR_INVALID, "pc_set_group_mode: ctx is NULL");
    c->group_mode = on ? 1 : 0;
    return PC_OK;
}

extern "C" int pc_last_group(const pc_ctx *c) { return c && c->last_group ? 1 : 0; }

extern "C" int pc_set_probe(pc_ctx *c, void *start_event, void *end_event) {
    if (!c) return fail(PC_ERR_INVALID, "pc_set_probe: ctx is NULL");
    c->probe[0] = start_event; c->probe[1] = end_event;
    return PC_OK;
}

extern "C" int pc_set_small_ctas(pc_ctx *c, int ctas) {
    if (!c || ctas < 0 || ctas > 4) return fail(PC_ERR_INVALID, "pc_set_small_ctas: 0 (automatic) .. 4");
    c->small_ctas = ctas;
    return PC_OK;
}

extern "C" int pc_set_limits(pc_ctx *c, int hits_per_frame, int hbonds_per_frame, int cells, int table_slots) {
    if (!c) return fail(PC_ERR_INVALID, "pc_set_limits: ctx is NULL");
    c->lim_hits = hits_per_frame; c->lim_hb = hbonds_per_frame; c->lim_cells = cells; c->lim_table = table_slots;
    return PC_OK;
}

// Device copies of one selection's static data: the packed atom word per heavy atom (heavy atoms in
// index order -- kernels never touch hydrogens except inside the H-bond test) and per local atom,
// a heavy-atom bitmask for the streaming kernels, and the per-atom arrays the rare paths read.
struct SideBufs {
    DevBuf<uint32_t> *hinfo, *hmask, *hpre, *linfo;
    DevBuf<int> *lres, *lseg, *lhptr, *lhidx;
    DevBuf<uint8_t> *lflag;
};

static cudaError_t upload_side(int n, const uint8_t *flags, const int32_t *resid, const int32_t *segid,
                               const int32_t *hptr, const int32_t *hidx, SideBufs b, int &nheavy,
                               std::vector<uint32_t> *hinfo_out = nullptr) {
    std::vector<uint32_t> hinfo, linfo((size_t)n), hmask(((size_t)n + 31) / 32 + 4, 0u);
    std::vector<int> res((size_t)n, 0), seg((size_t)n, 0), ptr((size_t)n + 1, 0), idx;
    for (int i = 0; i < n; i++) {
        uint32_t w = (uint32_t)i;
        if (flags[i] & PC_ATOM_BACKBONE) w |= PC_W_BACKBONE;
        if (flags[i] & PC_ATOM_HBCLASS) w |= PC_W_HBCLASS;
        linfo[i] = w;
        if (!(flags[i] & PC_ATOM_HYDROGEN)) { hinfo.push_back(w); hmask[i >> 5] |= 1u << (i & 31); }
        if (resid) res[i] = resid[i];
        if (segid) seg[i] = segid[i];
        if (hptr && (flags[i] & PC_ATOM_HBCLASS) && !(flags[i] & PC_ATOM_HYDROGEN))
            for (int p = hptr[i]; p < hptr[i + 1]; p++) idx.push_back(hidx[p]);
        ptr[i + 1] = (int)idx.size();
    }
    nheavy = (int)hinfo.size();
    std::vector<uint32_t> hpre(hmask.size() + 1, 0u);
    for (size_t w = 0; w < hmask.size(); w++) hpre[w + 1] = hpre[w] + (uint32_t)__builtin_popcount(hmask[w]);
    cudaError_t e;
    if ((e = b.hpre->upload(hpre)) != cudaSuccess) return e;
    if ((e = b.hinfo->upload(hinfo)) != cudaSuccess) return e;
    if ((e = b.hmask->upload(hmask)) != cudaSuccess) return e;
    if ((e = b.linfo->upload(linfo)) != cudaSuccess) return e;
    if ((e = b.lres->upload(res)) != cudaSuccess) return e;
    if ((e = b.lseg->upload(seg)) != cudaSuccess) return e;
    if ((e = b.lhptr->upload(ptr)) != cudaSuccess) return e;
    if ((e = b.lhidx->upload(idx)) != cudaSuccess) return e;
    const std::vector<uint8_t> fl(flags, flags + n);
    if ((e = b.lflag->upload(fl)) != cudaSuccess) return e;
    e = cudaDeviceSynchronize();         // every host vector above outlives this synchronisation
    if (hinfo_out) hinfo_out->swap(hinfo);
    return e;
}

extern "C" int pc_set_topology(pc_ctx *c, int32_t n1, int32_t n2, const uint8_t *f1, const uint8_t *f2,
                               const int32_t *resid1, const int32_t *segid1, const int32_t *resid2,
                               const int32_t *segid2, const int32_t *hptr1, const int32_t *hidx1,
                               const int32_t *hptr2, const int32_t *hidx2) {
    if (!c) return fail(PC_ERR_INVALID, "pc_set_topology: ctx is NULL");
    if (n1 <= 0 || !f1) return fail(PC_ERR_INVALID, "pc_set_topology: selection 1 is empty");
    if (!c->self_mode && (n2 <= 0 || !f2)) return fail(PC_ERR_INVALID, "pc_set_topology: selection 2 is empty");
    if (n1 >= (1 << 27) || n2 >= (1 << 27)) return fail(PC_ERR_INVALID, "pc_set_topology: at most 2^27-1 atoms per selection");
    CUDA_TRY(cudaSetDevice(c->device));
    c->n1 = n1;
    CUDA_TRY(upload_side(n1, f1, resid1, segid1, hptr1, hidx1,
                         SideBufs{&c->hinfo1, &c->hmask1, &c->hpre1, &c->linfo1, &c->lres1, &c->lseg1, &c->lhptr1, &c->lhidx1, &c->lflag1},
                         c->n1h, &c->h_hinfo1));
    if (c->self_mode) {
        c->n2 = n1; c->n2h = c->n1h;
    } else {
        c->n2 = n2;
        CUDA_TRY(upload_side(n2, f2, resid2, segid2, hptr2, hidx2,
                             SideBufs{&c->hinfo2, &c->hmask2, &c->hpre2, &c->linfo2, &c->lres2, &c->lseg2, &c->lhptr2, &c->lhidx2, &c->lflag2},
                             c->n2h));
    }
    c->have_topology = true;
    c->have_maps = false;
    return PC_OK;
}

extern "C" int pc_set_maps(pc_ctx *c, const int32_t *gid1, const int32_t *gid2, int64_t ngroups2) {
    if (!c || !c->have_topology) return fail(PC_ERR_INVALID, "pc_set_maps: set the topology first");
    if (!gid1 || !gid2 || ngroups2 <= 0) return fail(PC_ERR_INVALID, "pc_set_maps: bad arguments");
    CUDA_TRY(cudaSetDevice(c->device));
    const std::vector<int> g1v(gid1, gid1 + c->n1), g2v(gid2, gid2 + c->n2);      // alive until the synchronisation below
    CUDA_TRY(c->lgid1.upload(g1v));
    CUDA_TRY(c->lgid2.upload(g2v));
    CUDA_TRY(c->lgb1.ensure((size_t)c->n1));
    CUDA_TRY(c->lgb2.ensure((size_t)c->n2));
    if (c->n1) pc_pack_gid_flag_kernel<<<(c->n1 + 255) / 256, 256>>>(c->lgid1.p, c->lflag1.p, c->n1, c->lgb1.p);
    if (c->n2) pc_pack_gid_flag_kernel<<<(c->n2 + 255) / 256, 256>>>(c->lgid2.p, c->self_mode ? c->lflag1.p : c->lflag2.p, c->n2, c->lgb2.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaDeviceSynchronize());
    c->ngroups2 = (unsigned long long)ngroups2;
    int g1max = 0;
    for (int i = 0; i < c->n1; i++) g1max = std::max(g1max, gid1[i]);
    c->ngroups1 = (unsigned long long)g1max + 1;
    // Units of the group kernel: selection 1's heavy atoms in index order, cut into runs of whole groups of at most
    // PC_GROUP_NT atoms.  Only when every group is ONE run of consecutive heavy atoms (residue-level maps on any usual
    // topology) does a unit see all contacts of its keys.
    c->n_units = 0;
    {
        std::vector<uint32_t> units;
        std::vector<uint8_t> seen((size_t)g1max + 1, 0);
        bool ok = !c->h_hinfo1.empty() && (int)c->h_hinfo1.size() == c->n1h;
        uint32_t unit_first = 0, run_first = 0;
        units.push_back(0u);
        for (uint32_t h = 0; ok && h <= (uint32_t)c->n1h; h++) {
            const bool end = h == (uint32_t)c->n1h;
            const int g = end ? -1 : gid1[PC_W_INDEX(c->h_hinfo1[h])];
            if (!end && g < 0) { ok = false; break; }
            const int gprev = h ? gid1[PC_W_INDEX(c->h_hinfo1[h - 1])] : -2;
            if (h && g == gprev) continue;
            // a run [run_first, h) of group gprev ends here
            if (h) {
                if (h - run_first > (uint32_t)PC_GROUP_NT) { ok = false; break; }
                if (h - unit_first > (uint32_t)PC_GROUP_NT) { units.push_back(run_first); unit_first = run_first; }
            }
            if (!end) {
                if (seen[g]) { ok = false; break; }
                seen[g] = 1;
                run_first = h;
            }
        }
        if (ok) {
            units.push_back((uint32_t)c->n1h);
            CUDA_TRY(c->group_units.upload(units));
            CUDA_TRY(cudaDeviceSynchronize());
            c->n_units = (int)units.size() - 1;
        }
    }
    c->have_maps = true;
    return PC_OK;
}

extern "C" int pc_reserve(pc_ctx *c, int32_t max_frames, int64_t cap_contacts, int64_t cap_hbonds, int64_t cap_rows) {
    if (!c) return fail(PC_ERR_INVALID, "pc_reserve: ctx is NULL");
    if (max_frames <= 0 || cap_contacts <= 0 || cap_hbonds <= 0 || cap_rows <= 0)
        return fail(PC_ERR_INVALID, "pc_reserve: capacities must be positive");
    if (cap_contacts > 0xffffffffll || cap_hbonds > 0xffffffffll || cap_rows > 0xffffffffll)
        return fail(PC_ERR_INVALID, "pc_reserve: per-batch capacities are limited to 2^32-1 records");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(c->contacts.ensure((size_t)cap_contacts));
    CUDA_TRY(c->order_keys.ensure((size_t)cap_contacts));
    CUDA_TRY(c->hbonds.ensure((size_t)cap_hbonds));
    CUDA_TRY(c->rows.ensure((size_t)cap_rows));
    CUDA_TRY(c->frame_meta.ensure((size_t)max_frames * 6));
    c->max_frames = max_frames;
    c->cap_contacts = cap_contacts; c->cap_hbonds = cap_hbonds; c->cap_rows = cap_rows;
    return PC_OK;
}

static PcTopoDev topo_dev(const pc_ctx *c) {
    PcTopoDev t;
    const bool self = c->self_mode != 0;
    t.n1 = c->n1; t.n2 = c->n2; t.n1h = c->n1h; t.n2h = c->n2h;
    t.hinfo1 = c->hinfo1.p; t.hinfo2 = self ? c->hinfo1.p : c->hinfo2.p;
    t.hmask1 = c->hmask1.p; t.hmask2 = self ? c->hmask1.p : c->hmask2.p;
    t.hpre1 = c->hpre1.p; t.hpre2 = self ? c->hpre1.p : c->hpre2.p;
    t.linfo1 = c->linfo1.p; t.linfo2 = self ? c->linfo1.p : c->linfo2.p;
    t.lres1 = c->lres1.p; t.lseg1 = c->lseg1.p;
    t.lres2 = self ? c->lres1.p : c->lres2.p; t.lseg2 = self ? c->lseg1.p : c->lseg2.p;
    t.lhptr1 = c->lhptr1.p; t.lhidx1 = c->lhidx1.p;
    t.lhptr2 = self ? c->lhptr1.p : c->lhptr2.p; t.lhidx2 = self ? c->lhidx1.p : c->lhidx2.p;
    t.lgid1 = c->lgid1.p; t.lgid2 = c->lgid2.p;
    t.lflag1 = c->lflag1.p; t.lflag2 = self ? c->lflag1.p : c->lflag2.p;
    t.ngroups2 = c->ngroups2;
    return t;
}

static uint32_t *meta(pc_ctx *c, int which) { return c->frame_meta.p + (size_t)which * c->max_frames; }

// Shared-memory plans of the small-frame kernel.  "Fast": two 512-thread CTAs per SM, staging for a
// few thousand in-grid atoms (one frame's barrier and load latency hides behind the other CTA's
// work).  "Full": one 1024-thread CTA per SM with the largest staging that fits.  A batch runs the
// fast plan over all frames and the full plan over the frames the fast plan could not stage; when
// not even the full plan can hold every heavy atom of a frame, overflowing frames report
// PC_ST_STAGE_OVERFLOW and the caller switches to the large-frame path.
static bool plan_small_cfg(const pc_ctx *c, bool fuse_acc, int ctas, int &capCells, int &capHits, int &capHb, PcSmallSmem &L) {
    const int ntot = c->self_mode ? c->n1h : c->n1h + c->n2h;
    capHits = c->lim_hits > 0 ? c->lim_hits : (ctas >= 3 ? 3072 : 4096);
    capHb = c->lim_hb > 0 ? c->lim_hb : 256;
    int capTab = 0;
    if (fuse_acc) {
        capTab = 64;
        const int want = c->lim_table > 0 ? c->lim_table : (ctas >= 2 ? 512 : 1024);
        while (capTab < want) capTab <<= 1;
    }
    const size_t optin = c->smem_optin ? c->smem_optin : 232448;
    const size_t per_sm = 233472;
    const size_t budget = (ctas == 1 ? optin : std::min(optin, per_sm / ctas - 1024)) - 3072;   // minus static shared memory
    const int cell_arrays = c->self_mode ? 1 : 2;
    const size_t min_cells = (size_t)cell_arrays * 4 * (ctas == 1 ? 2049 : 1025);
    // the largest candidate pool that leaves room for the minimum grid
    int lo = 0, hi = std::max(1, std::min(ntot, 65535));
    while (lo < hi) {
        const int mid = (lo + hi + 1) / 2;
        PcSmallSmem t = pc_small_layout(c->self_mode, 0, capHits, capHb, mid, capTab);
        if ((size_t)t.total + min_cells <= budget) lo = mid; else hi = mid - 1;
    }
    const int capCand = lo;
    if (capCand < std::min(128, std::max(ntot, 1))) return false;
    // the fast plan only pays when a fair share of the atoms can be staged
    if (ctas > 1 && capCand < ntot && (long long)capCand * 6 < (long long)ntot) return false;
    PcSmallSmem base = pc_small_layout(c->self_mode, 0, capHits, capHb, capCand, capTab);
    int cells = (int)((budget - base.total) / (4 * cell_arrays)) - 8;
    cells = std::min(cells, 16384);
    if (c->lim_cells > 0) cells = std::min(cells, c->lim_cells);
    if (cells < 64) return false;
    capCells = cells;
    L = pc_small_layout(c->self_mode, capCells, capHits, capHb, capCand, capTab);
    return true;
}

static PcAccArgs acc_args(pc_ctx *c, bool order_keys, int nframes) {
    PcAccArgs a;
    memset(&a, 0, sizeof a);
    a.contacts = c->contacts.p; a.hbonds = c->hbonds.p;
    a.order_keys = order_keys ? c->order_keys.p : nullptr;
    a.frame_cstart = meta(c, 0); a.frame_ccount = meta(c, 1); a.frame_hstart = meta(c, 2); a.frame_hcount = meta(c, 3);
    a.lgid1 = c->lgid1.p; a.lgid2 = c->lgid2.p;
    a.lflag1 = c->lflag1.p; a.lflag2 = c->self_mode ? c->lflag1.p : c->lflag2.p;
    a.lgb1 = c->lgb1.p; a.lgb2 = c->lgb2.p;
    a.ngroups2 = c->ngroups2;
    a.nframes = nframes;
    a.rows = c->rows.p; a.cap_rows = c->cap_rows; a.counters = c->counters.p;
    a.frame_rstart = meta(c, 4); a.frame_rcount = meta(c, 5);
    return a;
}

extern "C" int pc_scan_device(pc_ctx *c, const float *d_xyz1, const float *d_xyz2, int32_t nframes,
                              int32_t layout, int64_t pitch1, int64_t pitch2, int32_t order_keys, void *stream) {
    if (!c || !c->have_topology) return fail(PC_ERR_INVALID, "pc_scan_device: set the topology first");
    if (nframes <= 0) return fail(PC_ERR_INVALID, "pc_scan_device: nframes must be positive");
    if (layout != PC_LAYOUT_XYZ && layout != PC_LAYOUT_XYZW) return fail(PC_ERR_INVALID, "pc_scan_device: bad layout");
    if (!d_xyz1 || (!c->self_mode && !d_xyz2)) return fail(PC_ERR_INVALID, "pc_scan_device: NULL coordinates");
    if (nframes > c->max_frames || c->cap_contacts <= 0)
        return fail(PC_ERR_INVALID, "pc_scan_device: call pc_reserve with max_frames >= nframes first");
    if (layout == PC_LAYOUT_XYZW && (((uintptr_t)d_xyz1 | (uintptr_t)d_xyz2) & 15 || (pitch1 & 3) || (pitch2 & 3)))
        return fail(PC_ERR_INVALID, "pc_scan_device: float4 layout needs 16-byte aligned frames");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;

    PcScanArgs a;
    memset(&a, 0, sizeof(a));
    a.xyz1 = d_xyz1; a.xyz2 = c->self_mode ? d_xyz1 : d_xyz2;
    a.pitch1 = pitch1; a.pitch2 = c->self_mode ? pitch1 : pitch2;
    a.stride = layout; a.nframes = nframes; a.self_mode = c->self_mode; a.order_keys = order_keys;
    a.t = topo_dev(c);
    a.cutoff_f = (float)c->cutoff;                       // gridsearch.C:417 -> :904 (double -> float)
    a.sqdist = a.cutoff_f * a.cutoff_f;                  // gridsearch.C:936
    a.edge = a.cutoff_f * PC_EDGE_SCALE;
    a.hb_cutoff = c->hb_cutoff; a.hb_angle = c->hb_angle;
    a.contacts = c->contacts.p; a.hbonds = c->hbonds.p; a.order_out = order_keys ? c->order_keys.p : nullptr;
    a.counters = c->counters.p;
    a.cap_contacts = c->cap_contacts; a.cap_hbonds = c->cap_hbonds;
    a.frame_cstart = meta(c, 0); a.frame_ccount = meta(c, 1); a.frame_hstart = meta(c, 2); a.frame_hcount = meta(c, 3);
    CUDA_TRY(cudaMemsetAsync(c->counters.p, 0, PC_CNT_COUNT * sizeof(unsigned long long), s));

    PcSmallSmem L;
    // accumulation inside the scan kernel when the maps are known and the reference's first-appearance
    // key (an order key) is not asked for; pc_accumulate then has nothing left to do for this batch
    const bool fuse_acc = c->want_fused_acc && c->have_maps && !order_keys && c->lim_table <= 2048;
    a.fuse_acc = 0;
    a.dense_mode = c->dense_mode;      // 0: pair kernel chosen per frame on the device (pc_large_is_dense)
    a.rows = c->rows.p; a.cap_rows = c->cap_rows; a.frame_rstart = meta(c, 4); a.frame_rcount = meta(c, 5);
    PcSmallSmem L2;
    int cells1 = 0, hits1 = 0, hb1 = 0, cells2 = 0, hits2 = 0, hb2 = 0;
    const bool small_ok = c->force_path != 2 && c->n1h <= 65535 && c->n2h <= 65535;
    const bool have_full = small_ok && c->small_ctas != 2 && plan_small_cfg(c, fuse_acc, 1, cells1, hits1, hb1, L);
    // CTAs per SM of the fast plan: two of 512 threads; selections of a few thousand heavy atoms (the reference's own fixture:
    // 2.2 k) run three of 384 -- their phases are short and latency-bound, a third frame in flight hides more of it (c1
    // kernel 0.216 -> 0.192 ms per 1600 frames, the 1920-atom synthetic system 0.386 -> 0.329 ms per 4096), while at C2's
    // 6000 heavy atoms the search and the drain are bound by the SM's shared-memory pipe and three CTAs are slower (0.49 vs 0.36)
    const int small_sel = (c->self_mode ? c->n1h : c->n1h + c->n2h) <= 3000;
    const int fast_ctas = c->small_ctas >= 3 ? c->small_ctas : (c->small_ctas == 0 && small_sel ? 3 : 2);
    const bool have_fast = small_ok && c->small_ctas != 1 && plan_small_cfg(c, fuse_acc, fast_ctas, cells2, hits2, hb2, L2);
    const bool fits_small = have_full || have_fast;
    c->last_fused = false;
    c->last_group = false;
    c->packed_valid = false;
    if (c->force_path == 1 && !fits_small)
        return fail(PC_ERR_INVALID, "pc_scan_device: the frame does not fit the small-frame kernel");
    if (fits_small) {
        a.fuse_acc = fuse_acc ? 1 : 0;
        CUDA_TRY(c->retry.ensure((size_t)c->max_frames + 4));
        unsigned int *retry_count = reinterpret_cast<unsigned int *>(c->retry.p);
        int *retry_list = c->retry.p + 4;
        CUDA_TRY(cudaMemsetAsync(retry_count, 0, sizeof(unsigned int), s));
        PcSmallFrames fl;
        if (have_fast) {
            auto kern = fast_ctas == 4 ? pc_scan_small_kernel<256> : fast_ctas == 3 ? pc_scan_small_kernel<384> : pc_scan_small_kernel<512>;
            const int nt = fast_ctas == 4 ? 256 : fast_ctas == 3 ? 384 : 512;
            CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L2.total));
            PcScanArgs a2 = a;
            a2.capCells = cells2; a2.capHits = hits2; a2.capHb = hb2;
            fl.frame_list = nullptr; fl.frame_count = nullptr;
            fl.retry_list = have_full ? retry_list : nullptr; fl.retry_count = have_full ? retry_count : nullptr;
            kern<<<std::min(nframes, fast_ctas * c->num_sms), nt, L2.total, s>>>(a2, L2, fl);
            CUDA_TRY(cudaGetLastError());
            g_launches += 1;
        }
        if (have_full) {
            auto kern = pc_scan_small_kernel<1024>;
            CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
            a.capCells = cells1; a.capHits = hits1; a.capHb = hb1;
            fl.frame_list = have_fast ? retry_list : nullptr; fl.frame_count = have_fast ? retry_count : nullptr;
            fl.retry_list = nullptr; fl.retry_count = nullptr;
            kern<<<std::min(nframes, c->num_sms), 1024, L.total, s>>>(a, L, fl);
            CUDA_TRY(cudaGetLastError());
            g_launches += 1;
        }
        c->last_small_ctas = have_fast ? fast_ctas : 1;
        c->last_fused = fuse_acc;
        c->last_path = 1;
    } else {
        // the fused tail kernel also accumulates when the maps are known, no order keys are asked for and the
        // frame's keys are expected to fit its shared-memory table (pc_set_limits' table_slots <= 8192)
        const int tail = c->self_mode ? 0 : c->tail_mode;
        const bool can_acc = c->want_fused_acc && c->have_maps && !order_keys && c->lim_table <= PC_COMPACT_SLOTS;
        const bool tail_acc = tail && can_acc;
        // frames the tail kernel does not take (self contacts, or handed over after PC_ST_TAIL_OVERFLOW): the group kernel
        // searches, records and accumulates unit by unit of whole groups (pc_scan_group.cuh) when the maps allow it
        const bool group = !tail && can_acc && c->group_mode && c->n_units > 0 && c->n2h < (1 << 24) &&
                           c->cap_rows / nframes >= 1;
        PcAccArgs acc = acc_args(c, false, nframes);
        PcGroupArgs ga;
        memset(&ga, 0, sizeof ga);
        if (group) {
            CUDA_TRY(c->row_slices.ensure((size_t)c->cap_rows));
            ga.unit_start = c->group_units.p; ga.nunits = c->n_units;
            ga.slices = c->row_slices.p; ga.cap_r_frame = c->cap_rows / nframes;
            CUDA_TRY(cudaMemsetAsync(acc.frame_rcount, 0, (size_t)nframes * sizeof(uint32_t), s));
        }
        int rc = pc_large_scan(c->large, a, c->num_sms, s, g_err, &g_launches, (cudaEvent_t)c->probe[0], (cudaEvent_t)c->probe[1],
                               tail, (tail_acc || group) ? &acc : nullptr, group ? &ga : nullptr);
        if (rc != PC_OK) return rc;
        if (group) {
            pc_group_offsets_kernel<<<1, 1024, 0, s>>>(acc, ga.cap_r_frame);
            const int gb = std::max(1, std::min(64, 4 * c->num_sms / nframes + 1));
            pc_group_gather_kernel<<<dim3(gb, nframes), 256, 0, s>>>(acc, ga.slices, ga.cap_r_frame);
            CUDA_TRY(cudaGetLastError());
            g_launches += 2;
        }
        c->last_path = 2;
        c->last_fused = tail_acc || group;
        c->last_group = group;
    }
    if (order_keys) {
        if (c->n1 >= (1 << 27)) return fail(PC_ERR_INVALID, "pc_scan_device: order keys need n1 < 2^27");
        pc_order_kernel<256><<<std::min(nframes, 8 * c->num_sms), 256, 0, s>>>(a);
        CUDA_TRY(cudaGetLastError());
        g_launches += 1;
    }
    c->last_nframes = nframes;
    c->last_order_keys = order_keys != 0;
    return PC_OK;
}

extern "C" int pc_set_fused_accumulate(pc_ctx *c, int on) {
    if (!c) return fail(PC_ERR_INVALID, "pc_set_fused_accumulate: ctx is NULL");
    c->want_fused_acc = on != 0;
    return PC_OK;
}

extern "C" int pc_accumulate(pc_ctx *c, void *stream) {
    if (!c || !c->have_maps) return fail(PC_ERR_INVALID, "pc_accumulate: set the maps first");
    if (c->last_nframes <= 0) return fail(PC_ERR_INVALID, "pc_accumulate: no scanned batch");
    if (c->last_fused) return PC_OK;            // the scan kernel already produced this batch's rows
    c->packed_valid = false;
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    PcAccArgs a = acc_args(c, c->last_order_keys, c->last_nframes);
    int cap = c->lim_table > 0 ? c->lim_table : 1024;
    int p2 = 64; while (p2 < cap) p2 <<= 1;
    a.cap = p2;
    const size_t smem = (size_t)a.cap * 48;
    const size_t smem_max = (c->smem_optin ? c->smem_optin : 232448) - 1024;
    // one CTA per frame with the table in shared memory while a frame's keys fit (<= 2048 slots);
    // otherwise -- and always after the large-frame scan -- one table in global memory for the batch
    if (c->last_path != 2 && a.cap <= 2048 && smem <= smem_max) {
        auto kern = pc_accumulate_kernel<256>;
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, 233472 / (smem + 2048)));
        const int grid = std::min(a.nframes, per_sm * c->num_sms);
        kern<<<grid, 256, smem, s>>>(a);
        CUDA_TRY(cudaGetLastError());
        g_launches += 1;
        return PC_OK;
    }
    // Thousands of keys per frame, no order keys: compact 28-byte slots, 8192 of them in one SM's shared memory
    if (!c->last_order_keys && c->lim_table <= 8192) {
        a.cap = PC_COMPACT_SLOTS;
        const size_t sm2 = PC_COMPACT_SMEM;
        auto kern = pc_accumulate_compact_kernel<1024>;
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
        kern<<<std::min(a.nframes, c->num_sms), 1024, sm2, s>>>(a);
        CUDA_TRY(cudaGetLastError());
        g_launches += 1;
        return PC_OK;
    }
    // The table is sized from the number of contacts: the previous batch's count (scaled by the frame
    // ratio, +50%) when there is one -- a too small table sets PC_ST_TABLE_OVERFLOW and the caller
    // retries -- else the counters are read back, which synchronises the stream once.
    unsigned long long ncontacts;
    const bool hinted = c->gacc_hint && c->gacc_hint_frames > 0 && !c->gacc_retry;
    if (hinted) {
        ncontacts = c->gacc_hint * (unsigned long long)a.nframes / c->gacc_hint_frames;
        ncontacts += ncontacts / 2 + 1024;
    } else {
        CUDA_TRY(cudaMemcpyAsync(c->h_counters, c->counters.p, PC_CNT_COUNT * sizeof(unsigned long long),
                                 cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        if (c->h_counters[PC_CNT_STATUS] & (PC_ST_OUT_OVERFLOW | PC_ST_HITS_OVERFLOW | PC_ST_HB_OVERFLOW | PC_ST_STAGE_OVERFLOW |
                                            PC_ST_TAIL_OVERFLOW))
            return PC_OK;                    // the scan overflowed; pc_batch_totals reports it and the caller rescans
        ncontacts = std::min<unsigned long long>(c->h_counters[PC_CNT_CONTACTS], c->cap_contacts);
        c->gacc_retry = false;
    }
    c->gacc_used = true;
    PcGlobalTable t;
    // distinct (frame, key) cells are what fills the table: the previous batch's row count when known
    unsigned long long cells = ncontacts;
    if (c->gacc_hint_rows && c->gacc_hint_frames > 0 && hinted) {
        cells = c->gacc_hint_rows * (unsigned long long)a.nframes / c->gacc_hint_frames;
        cells += cells / 2 + 1024;
    }
    t.cap = 1024;
    while (t.cap < 2 * cells + 16) t.cap <<= 1;
    t.nkeys = (unsigned long long)c->ngroups1 * c->ngroups2;
    if (t.nkeys == 0 || (unsigned long long)a.nframes > 0xffffffffffffffffull / std::max<unsigned long long>(t.nkeys, 1) - 1)
        return fail(PC_ERR_INVALID, "pc_accumulate: frame * key space exceeds 64 bits");
    CUDA_TRY(c->gacc_tab.ensure(t.cap));
    CUDA_TRY(c->gacc_cursor.ensure((size_t)c->max_frames));
    CUDA_TRY(c->gacc_fresh.ensure((size_t)c->cap_rows));
    CUDA_TRY(c->gacc_nfresh.ensure(1));
    t.e = c->gacc_tab.p;
    t.fresh = c->gacc_fresh.p; t.nfresh = c->gacc_nfresh.p;
    pc_gacc_init_kernel<<<4 * c->num_sms, 256, 0, s>>>(t);
    CUDA_TRY(cudaMemsetAsync(a.frame_rcount, 0, (size_t)a.nframes * sizeof(uint32_t), s));
    const int per_frame_blocks = std::max(1, 4 * c->num_sms / a.nframes);
    const unsigned long long avg = ncontacts / a.nframes + 1;
    const int blocks = (int)std::max<unsigned long long>(1, std::min<unsigned long long>((avg * 2 + 255) / 256, per_frame_blocks));
    CUDA_TRY(cudaFuncSetAttribute(pc_gacc_insert_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PC_GACC_SMEM));
    pc_gacc_insert_kernel<<<dim3(blocks, a.nframes), 256, PC_GACC_SMEM, s>>>(a, t);
    pc_gacc_hbond_kernel<<<dim3(std::max(1, blocks / 4), a.nframes), 256, 0, s>>>(a, t);
    pc_gacc_offsets_kernel<<<1, 1024, 0, s>>>(a, c->gacc_cursor.p);
    pc_gacc_emit_kernel<<<4 * c->num_sms, 256, 0, s>>>(a, t, c->gacc_cursor.p);
    CUDA_TRY(cudaGetLastError());
    g_launches += 5;
    return PC_OK;
}

extern "C" int pc_fold_rows(pc_ctx *c, double *d_totals, int64_t n_keys, void *stream) {
    if (!c || !c->have_maps || !d_totals || n_keys <= 0) return fail(PC_ERR_INVALID, "pc_fold_rows: bad arguments");
    CUDA_TRY(cudaSetDevice(c->device));
    pc_fold_rows_kernel<<<2 * c->num_sms, 256, 0, (cudaStream_t)stream>>>(c->rows.p, c->counters.p, c->cap_rows, d_totals,
                                                                         (unsigned long long)n_keys);
    CUDA_TRY(cudaGetLastError());
    g_launches += 1;
    return PC_OK;
}

static int key_stats_launch(int device, const double *d_scores, const uint32_t *d_hbonds, const long long *d_key_ptr,
                            const int *d_row_frame, int64_t n_keys, int32_t n_frames, double ns_per_frame, double threshold,
                            double *d_out, void *stream) {
    CUDA_TRY(cudaSetDevice(device));
    const int fpad = (n_frames + 1) & ~1;
    const size_t smem = sizeof(double) * (size_t)fpad + sizeof(unsigned int) * (size_t)(n_frames / 2 + 2);
    auto kern = d_key_ptr ? pc_key_stats_kernel<true> : pc_key_stats_kernel<false>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int sms = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int per_sm = std::max(1, std::min(4, (int)((size_t)(224 * 1024) / (smem + 4096))));
    const int grid = (int)std::min<int64_t>(n_keys, (int64_t)per_sm * sms);
    kern<<<grid, PC_STATS_NT, smem, (cudaStream_t)stream>>>(d_scores, d_hbonds, d_key_ptr, d_row_frame, (long long)n_keys, n_frames,
                                                           fpad, ns_per_frame, threshold, d_out);
    CUDA_TRY(cudaGetLastError());
    g_launches += 1;
    return PC_OK;
}

extern "C" int pc_key_stats(int device, const double *d_scores, const uint32_t *d_hbonds, int64_t n_keys, int32_t n_frames,
                            double ns_per_frame, double threshold, double *d_out, void *stream) {
    if (!d_scores || !d_out || n_keys < 0 || n_frames < 0) return fail(PC_ERR_INVALID, "pc_key_stats: bad arguments");
    if (n_frames > PC_STATS_MAX_FRAMES)
        return fail(PC_ERR_INVALID, "pc_key_stats: more than 16384 frames per call (split the frame range)");
    if (n_keys == 0) return PC_OK;
    return key_stats_launch(device, d_scores, d_hbonds, nullptr, nullptr, n_keys, n_frames, ns_per_frame, threshold, d_out, stream);
}

extern "C" int pc_key_stats_rows(int device, const int64_t *d_key_ptr, const int32_t *d_row_frame, const double *d_row_score,
                                 const uint32_t *d_row_hbonds, int64_t n_keys, int32_t n_frames, double ns_per_frame,
                                 double threshold, double *d_out, void *stream) {
    if (!d_key_ptr || !d_out || n_keys < 0 || n_frames < 0) return fail(PC_ERR_INVALID, "pc_key_stats_rows: bad arguments");
    if (n_frames > PC_STATS_MAX_FRAMES)
        return fail(PC_ERR_INVALID, "pc_key_stats_rows: more than 16384 frames per call (split the frame range)");
    if (n_keys == 0) return PC_OK;
    static_assert(sizeof(long long) == sizeof(int64_t), "key_ptr is int64");
    return key_stats_launch(device, d_row_score, d_row_hbonds, reinterpret_cast<const long long *>(d_key_ptr), d_row_frame, n_keys,
                            n_frames, ns_per_frame, threshold, d_out, stream);
}

extern "C" int pc_fold_sparse_init(pc_ctx *c, int64_t entries, void *stream) {
    if (!c || entries <= 0) return fail(PC_ERR_INVALID, "pc_fold_sparse_init: bad arguments");
    CUDA_TRY(cudaSetDevice(c->device));
    unsigned long long cap = 1024;
    while (cap < (unsigned long long)entries) cap <<= 1;
    CUDA_TRY(c->fold_tab.ensure(cap));
    CUDA_TRY(c->fold_meta.ensure(2));
    if (!c->h_fold_meta) CUDA_TRY(cudaHostAlloc(&c->h_fold_meta, 2 * sizeof(unsigned long long), cudaHostAllocDefault));
    c->fold_cap = cap;
    pc_fold_sparse_clear_kernel<<<4 * c->num_sms, 256, 0, (cudaStream_t)stream>>>(c->fold_tab.p, cap, c->fold_meta.p);
    CUDA_TRY(cudaGetLastError());
    g_launches += 1;
    return PC_OK;
}

extern "C" int pc_fold_sparse_share(pc_ctx *c, pc_ctx *owner) {
    if (!c || (owner && (owner->device != c->device || owner == c)))
        return fail(PC_ERR_INVALID, "pc_fold_sparse_share: owner must be another context of the same GPU");
    c->fold_owner = owner;
    return PC_OK;
}

extern "C" int pc_fold_rows_sparse(pc_ctx *c, void *stream) {
    pc_ctx *t = c && c->fold_owner ? c->fold_owner : c;
    if (!c || !c->have_maps || !t->fold_cap) return fail(PC_ERR_INVALID, "pc_fold_rows_sparse: call pc_fold_sparse_init first");
    CUDA_TRY(cudaSetDevice(c->device));
    pc_fold_sparse_kernel<<<2 * c->num_sms, 256, 0, (cudaStream_t)stream>>>(c->rows.p, c->counters.p, c->cap_rows, t->fold_tab.p,
                                                                           t->fold_cap, t->fold_meta.p);
    CUDA_TRY(cudaGetLastError());
    g_launches += 1;
    return PC_OK;
}

extern "C" int pc_fold_sparse_fetch(pc_ctx *c, void *stream, pc_fold_entry *h_entries, int64_t max_entries, int64_t *n_entries) {
    if (!c || !c->fold_cap || !h_entries || max_entries <= 0 || !n_entries)
        return fail(PC_ERR_INVALID, "pc_fold_sparse_fetch: bad arguments");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    CUDA_TRY(c->fold_out.ensure((size_t)max_entries));
    CUDA_TRY(cudaMemsetAsync(c->fold_meta.p + 1, 0, sizeof(unsigned long long), s));
    pc_fold_sparse_compact_kernel<<<4 * c->num_sms, 256, 0, s>>>(c->fold_tab.p, c->fold_cap, c->fold_out.p,
                                                                 (unsigned long long)max_entries, c->fold_meta.p);
    CUDA_TRY(cudaGetLastError());
    g_launches += 1;
    CUDA_TRY(cudaMemcpyAsync(c->h_fold_meta, c->fold_meta.p, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    if (c->h_fold_meta[0]) return fail(PC_ERR_CAPACITY, "pc_fold_sparse_fetch: the aggregation table overflowed; re-init it larger");
    if (c->h_fold_meta[1] > (unsigned long long)max_entries)
        return fail(PC_ERR_CAPACITY, "pc_fold_sparse_fetch: more entries than max_entries");
    *n_entries = (int64_t)c->h_fold_meta[1];
    if (*n_entries) {
        CUDA_TRY(cudaMemcpyAsync(h_entries, c->fold_out.p, (size_t)*n_entries * sizeof(pc_fold_entry), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
    }
    return PC_OK;
}

extern "C" int pc_fold_sparse_export(pc_ctx *c, void *stream, pc_fold_entry *d_out, int64_t max_entries) {
    if (!c || !c->fold_cap || !d_out || max_entries <= 0) return fail(PC_ERR_INVALID, "pc_fold_sparse_export: bad arguments");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    CUDA_TRY(cudaMemsetAsync(c->fold_meta.p + 1, 0, sizeof(unsigned long long), s));
    pc_fold_sparse_export_kernel<<<4 * c->num_sms, 256, 0, s>>>(c->fold_tab.p, c->fold_cap, d_out, (unsigned long long)max_entries,
                                                                c->fold_meta.p);
    pc_fold_sparse_pad_kernel<<<2 * c->num_sms, 256, 0, s>>>(d_out, (unsigned long long)max_entries, c->fold_meta.p);
    CUDA_TRY(cudaGetLastError());
    g_launches += 2;
    return PC_OK;
}

extern "C" int pc_fold_sparse_import(pc_ctx *c, void *stream, const pc_fold_entry *d_in, int64_t n_entries) {
    if (!c || !c->fold_cap || !d_in || n_entries < 0) return fail(PC_ERR_INVALID, "pc_fold_sparse_import: bad arguments");
    if (n_entries == 0) return PC_OK;
    CUDA_TRY(cudaSetDevice(c->device));
    pc_fold_sparse_import_kernel<<<2 * c->num_sms, 256, 0, (cudaStream_t)stream>>>(d_in, (unsigned long long)n_entries, c->fold_tab.p,
                                                                                   c->fold_cap, c->fold_meta.p);
    CUDA_TRY(cudaGetLastError());
    g_launches += 1;
    return PC_OK;
}

extern "C" int pc_fp32_peak(int device, int32_t iters, double *ffma_per_s) {
    if (!ffma_per_s || iters <= 0) return fail(PC_ERR_INVALID, "pc_fp32_peak: bad arguments");
    CUDA_TRY(cudaSetDevice(device));
    int sms = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    float *d = nullptr;
    CUDA_TRY(cudaMalloc(&d, 64));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    const int ctas = sms * 8;                      // 2048 threads per SM
    double best = 0.0;
    for (int rep = 0; rep < 4; rep++) {            // first pass warms up
        CUDA_TRY(cudaEventRecord(e0, 0));
        pc_ffma_chain_kernel<<<ctas, 256>>>(d, iters, 0.999f, 1e-3f);
        CUDA_TRY(cudaEventRecord(e1, 0));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        const double rate = (double)ctas * 256.0 * 64.0 * (double)iters / ((double)ms * 1e-3);
        if (rep > 0 && rate > best) best = rate;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    g_launches += 4;
    *ffma_per_s = best;
    return PC_OK;
}

extern "C" int pc_mem_info(int device, int64_t *free_bytes, int64_t *total_bytes) {
    CUDA_TRY(cudaSetDevice(device));
    size_t f = 0, t = 0;
    CUDA_TRY(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = (int64_t)f;
    if (total_bytes) *total_bytes = (int64_t)t;
    return PC_OK;
}

extern "C" int pc_load_records(pc_ctx *c, const pc_contact *h_contacts, const uint64_t *h_order_keys,
                               const pc_hbond *h_hbonds, int32_t nframes, const uint32_t *cs, const uint32_t *cc,
                               const uint32_t *hs, const uint32_t *hc, void *stream) {
    if (!c || !c->have_topology) return fail(PC_ERR_INVALID, "pc_load_records: set the topology first");
    if (nframes <= 0 || !cs || !cc || !hs || !hc) return fail(PC_ERR_INVALID, "pc_load_records: bad arguments");
    if (nframes > c->max_frames) return fail(PC_ERR_INVALID, "pc_load_records: call pc_reserve with max_frames >= nframes first");
    unsigned long long nc = 0, nh = 0;
    for (int f = 0; f < nframes; f++) {
        nc = std::max<unsigned long long>(nc, (unsigned long long)cs[f] + cc[f]);
        nh = std::max<unsigned long long>(nh, (unsigned long long)hs[f] + hc[f]);
    }
    if ((long long)nc > c->cap_contacts || (long long)nh > c->cap_hbonds)
        return fail(PC_ERR_CAPACITY, "pc_load_records: records exceed the reserved capacities");
    if ((nc && !h_contacts) || (nh && !h_hbonds)) return fail(PC_ERR_INVALID, "pc_load_records: NULL records");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    if (nc) CUDA_TRY(cudaMemcpyAsync(c->contacts.p, h_contacts, nc * sizeof(pc_contact), cudaMemcpyHostToDevice, s));
    if (nc && h_order_keys)
        CUDA_TRY(cudaMemcpyAsync(c->order_keys.p, h_order_keys, nc * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    if (nh) CUDA_TRY(cudaMemcpyAsync(c->hbonds.p, h_hbonds, nh * sizeof(pc_hbond), cudaMemcpyHostToDevice, s));
    const uint32_t *src[4] = {cs, cc, hs, hc};
    for (int k = 0; k < 4; k++)
        CUDA_TRY(cudaMemcpyAsync(meta(c, k), src[k], (size_t)nframes * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    unsigned long long cnt[PC_CNT_COUNT] = {0};
    cnt[PC_CNT_CONTACTS] = nc; cnt[PC_CNT_HBONDS] = nh;
    // the copies above read pageable or pinned caller memory: finish them before returning
    CUDA_TRY(cudaMemcpyAsync(c->counters.p, cnt, sizeof cnt, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    c->last_nframes = nframes;
    c->last_order_keys = h_order_keys != nullptr;
    c->last_path = 1;
    c->last_fused = false;
    return PC_OK;
}

extern "C" int pc_scan_host(pc_ctx *c, const float *h_xyz1, const float *h_xyz2, int32_t nframes,
                            int32_t layout, int32_t order_keys, int32_t accumulate, void *stream) {
    if (!c || !c->have_topology) return fail(PC_ERR_INVALID, "pc_scan_host: set the topology first");
    if (!h_xyz1 || (!c->self_mode && !h_xyz2)) return fail(PC_ERR_INVALID, "pc_scan_host: NULL coordinates");
    if (layout != PC_LAYOUT_XYZ && layout != PC_LAYOUT_XYZW) return fail(PC_ERR_INVALID, "pc_scan_host: bad layout");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    const size_t f1 = (size_t)c->n1 * layout, f2 = (size_t)c->n2 * layout;
    CUDA_TRY(c->stage1.ensure(f1 * nframes + 4));
    CUDA_TRY(cudaMemcpyAsync(c->stage1.p, h_xyz1, f1 * nframes * sizeof(float), cudaMemcpyHostToDevice, s));
    if (!c->self_mode) {
        CUDA_TRY(c->stage2.ensure(f2 * nframes + 4));
        CUDA_TRY(cudaMemcpyAsync(c->stage2.p, h_xyz2, f2 * nframes * sizeof(float), cudaMemcpyHostToDevice, s));
    }
    int rc = pc_scan_device(c, c->stage1.p, c->stage2.p, nframes, layout, (int64_t)f1, (int64_t)f2, order_keys, stream);
    if (rc != PC_OK) return rc;
    if (accumulate) return pc_accumulate(c, stream);
    return PC_OK;
}

extern "C" int pc_batch_totals(pc_ctx *c, void *stream, int64_t *n_contacts, int64_t *n_hbonds, int64_t *n_rows,
                               int64_t *n_pair_evals) {
    if (!c) return fail(PC_ERR_INVALID, "pc_batch_totals: ctx is NULL");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    CUDA_TRY(cudaMemcpyAsync(c->h_counters, c->counters.p, PC_CNT_COUNT * sizeof(unsigned long long),
                             cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    if (c->last_path == 2 && c->last_nframes > 0) {
        // large path: every frame owns a slice of the buffers; keep the slice table for pc_fetch
        c->h_meta.resize((size_t)4 * c->last_nframes);
        for (int k = 0; k < 4; k++)
            CUDA_TRY(cudaMemcpy(c->h_meta.data() + (size_t)k * c->last_nframes, meta(c, k),
                                (size_t)c->last_nframes * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    }
    const unsigned long long st = c->h_counters[PC_CNT_STATUS];
    if (c->gacc_used) {
        c->gacc_used = false;
        if (st & PC_ST_TABLE_OVERFLOW) c->gacc_retry = true;           // next pc_accumulate sizes from the real count
        else { c->gacc_hint_rows = c->h_counters[PC_CNT_ROWS];     // the true count, even beyond cap_rows
               c->gacc_hint = std::min<unsigned long long>(c->h_counters[PC_CNT_CONTACTS], c->cap_contacts);
               c->gacc_hint_frames = c->last_nframes; }
    }
    if (n_contacts) *n_contacts = (int64_t)std::min<unsigned long long>(c->h_counters[PC_CNT_CONTACTS], c->cap_contacts);
    if (n_hbonds) *n_hbonds = (int64_t)std::min<unsigned long long>(c->h_counters[PC_CNT_HBONDS], c->cap_hbonds);
    if (n_rows) *n_rows = (int64_t)std::min<unsigned long long>(c->h_counters[PC_CNT_ROWS], c->cap_rows);
    if (n_pair_evals) *n_pair_evals = (int64_t)c->h_counters[PC_CNT_EVALS];
    if (st & PC_ST_BAD_COORDS) return fail(PC_ERR_INVALID, "non-finite coordinates in a frame");
    if (st & PC_ST_MISSING_H)
        return fail(PC_ERR_MISSING_H, "a hydrogen bonded to a selected N/O/S atom is not part of that selection");
    if (st & PC_ST_RERUN_MASK) {
        char buf[320];
        snprintf(buf, sizeof buf,
                 "capacity exceeded (status 0x%llx: 1=hits/frame 2=hbonds/frame 4=batch buffers 32=accumulation table "
                 "64=small-frame staging 128=large-frame tail staging 256=float32 cell sums); "
                 "needed contacts=%llu hbonds=%llu rows=%llu detail=0x%llx",
                 st, c->h_counters[PC_CNT_CONTACTS], c->h_counters[PC_CNT_HBONDS], c->h_counters[PC_CNT_ROWS],
                 c->h_counters[PC_CNT_WORK]);
        g_err = buf;
        return PC_ERR_CAPACITY;
    }
    return PC_OK;
}

extern "C" int pc_last_status(pc_ctx *c, uint64_t *status, int64_t *need_contacts, int64_t *need_hbonds,
                              int64_t *need_rows, int32_t *path) {
    if (!c) return fail(PC_ERR_INVALID, "pc_last_status: ctx is NULL");
    if (status) *status = c->h_counters[PC_CNT_STATUS];
    if (need_contacts) *need_contacts = (int64_t)c->h_counters[PC_CNT_CONTACTS];
    if (need_hbonds) *need_hbonds = (int64_t)c->h_counters[PC_CNT_HBONDS];
    if (need_rows) *need_rows = (int64_t)c->h_counters[PC_CNT_ROWS];
    if (path) *path = c->last_path;
    return PC_OK;
}

#ifdef PC_PHASE_CLOCKS
// profiling build only: cycles per phase of the last scan (valid after pc_batch_totals)
extern "C" int pc_phase_clocks(pc_ctx *c, unsigned long long *out24) {
    for (int i = 0; i < 24; i++) out24[i] = c->h_counters[PC_CNT_PHASE + i];
    unsigned int retried = 0;                    // frames the fast plan handed to the full plan
    if (c->retry.p) CUDA_TRY(cudaMemcpy(&retried, c->retry.p, sizeof(retried), cudaMemcpyDeviceToHost));
    out24[14] = retried; out24[15] = (unsigned long long)c->last_small_ctas;
    return PC_OK;
}
#endif

extern "C" int pc_device_outputs(pc_ctx *c, pc_contact **d_contacts, pc_hbond **d_hbonds, pc_row **d_rows,
                                 uint64_t **d_order_keys, uint32_t **cs, uint32_t **cc, uint32_t **hs,
                                 uint32_t **hc, uint32_t **rs, uint32_t **rc) {
    if (!c) return fail(PC_ERR_INVALID, "pc_device_outputs: ctx is NULL");
    if (d_contacts) *d_contacts = c->contacts.p;
    if (d_hbonds) *d_hbonds = c->hbonds.p;
    if (d_rows) *d_rows = c->rows.p;
    if (d_order_keys) *d_order_keys = (uint64_t *)c->order_keys.p;
    uint32_t **outs[6] = {cs, cc, hs, hc, rs, rc};
    for (int k = 0; k < 6; k++) if (outs[k]) *outs[k] = meta(c, k);
    return PC_OK;
}

extern "C" int pc_fetch(pc_ctx *c, void *stream, pc_contact *h_contacts, pc_hbond *h_hbonds, pc_row *h_rows,
                        uint64_t *h_order_keys, uint32_t *cs, uint32_t *cc, uint32_t *hs, uint32_t *hc,
                        uint32_t *rs, uint32_t *rc) {
    if (!c) return fail(PC_ERR_INVALID, "pc_fetch: ctx is NULL");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    const size_t nc = std::min<unsigned long long>(c->h_counters[PC_CNT_CONTACTS], c->cap_contacts);
    const size_t nh = std::min<unsigned long long>(c->h_counters[PC_CNT_HBONDS], c->cap_hbonds);
    const size_t nr = std::min<unsigned long long>(c->h_counters[PC_CNT_ROWS], c->cap_rows);
    if (c->last_path == 2) {
        // sliced outputs -> dense host arrays; the start arrays handed back are the dense offsets
        const int nf = c->last_nframes;
        if ((int)c->h_meta.size() != 4 * nf) return fail(PC_ERR_INVALID, "pc_fetch: call pc_batch_totals first");
        const uint32_t *ms = c->h_meta.data();
        size_t oc = 0, oh = 0;
        for (int f = 0; f < nf; f++) {
            const uint32_t s0 = ms[f], n0 = ms[nf + f], s1 = ms[2 * nf + f], n1 = ms[3 * nf + f];
            if (h_contacts && n0)
                CUDA_TRY(cudaMemcpyAsync(h_contacts + oc, c->contacts.p + s0, n0 * sizeof(pc_contact), cudaMemcpyDeviceToHost, s));
            if (h_order_keys && n0 && c->last_order_keys)
                CUDA_TRY(cudaMemcpyAsync(h_order_keys + oc, c->order_keys.p + s0, n0 * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
            if (h_hbonds && n1)
                CUDA_TRY(cudaMemcpyAsync(h_hbonds + oh, c->hbonds.p + s1, n1 * sizeof(pc_hbond), cudaMemcpyDeviceToHost, s));
            if (cs) cs[f] = (uint32_t)oc;
            if (cc) cc[f] = n0;
            if (hs) hs[f] = (uint32_t)oh;
            if (hc) hc[f] = n1;
            oc += n0; oh += n1;
        }
        if (h_rows && nr)
            CUDA_TRY(cudaMemcpyAsync(h_rows, c->rows.p, nr * sizeof(pc_row), cudaMemcpyDeviceToHost, s));
        if (rs) CUDA_TRY(cudaMemcpyAsync(rs, meta(c, 4), (size_t)nf * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        if (rc) CUDA_TRY(cudaMemcpyAsync(rc, meta(c, 5), (size_t)nf * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        return PC_OK;
    }
    if (h_contacts && nc)
        CUDA_TRY(cudaMemcpyAsync(h_contacts, c->contacts.p, nc * sizeof(pc_contact), cudaMemcpyDeviceToHost, s));
    if (h_order_keys && nc && c->last_order_keys)
        CUDA_TRY(cudaMemcpyAsync(h_order_keys, c->order_keys.p, nc * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    if (h_hbonds && nh)
        CUDA_TRY(cudaMemcpyAsync(h_hbonds, c->hbonds.p, nh * sizeof(pc_hbond), cudaMemcpyDeviceToHost, s));
    if (h_rows && nr)
        CUDA_TRY(cudaMemcpyAsync(h_rows, c->rows.p, nr * sizeof(pc_row), cudaMemcpyDeviceToHost, s));
    uint32_t *outs[6] = {cs, cc, hs, hc, rs, rc};
    for (int k = 0; k < 6; k++)
        if (outs[k])
            CUDA_TRY(cudaMemcpyAsync(outs[k], meta(c, k), (size_t)c->last_nframes * sizeof(uint32_t),
                                     cudaMemcpyDeviceToHost, s));
    return PC_OK;
}

extern "C" int pc_pack_rows(pc_ctx *c, void *stream) {
    if (!c) return fail(PC_ERR_INVALID, "pc_pack_rows: ctx is NULL");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(c->rows_packed.ensure((size_t)c->cap_rows));
    if (c->pack_bytes == 12) {
        if (c->ngroups1 * c->ngroups2 > 0xffffffffull) return fail(PC_ERR_INVALID, "pc_pack_rows: key space exceeds 32 bits");
        pc_pack_rows_slim_kernel<<<2 * c->num_sms, 256, 0, (cudaStream_t)stream>>>(
            c->rows.p, c->counters.p, c->cap_rows, reinterpret_cast<pc_row_slim *>(c->rows_packed.p));
    } else {
        pc_pack_rows_kernel<<<2 * c->num_sms, 256, 0, (cudaStream_t)stream>>>(c->rows.p, c->counters.p, c->cap_rows, c->rows_packed.p);
    }
    CUDA_TRY(cudaGetLastError());
    g_launches += 1;
    c->packed_valid = true;
    return PC_OK;
}

extern "C" int pc_set_pack_format(pc_ctx *c, int bytes_per_row) {
    if (!c || (bytes_per_row != 24 && bytes_per_row != 12)) return fail(PC_ERR_INVALID, "pc_set_pack_format: 24 or 12");
    c->pack_bytes = bytes_per_row;
    c->packed_valid = false;
    return PC_OK;
}

extern "C" int pc_fetch_packed_rows(pc_ctx *c, void *stream, pc_row_packed *h_rows, uint32_t *rs, uint32_t *rc) {
    if (!c || !h_rows) return fail(PC_ERR_INVALID, "pc_fetch_packed_rows: bad arguments");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    const size_t nr = std::min<unsigned long long>(c->h_counters[PC_CNT_ROWS], c->cap_rows);
    if (!c->packed_valid) { int r = pc_pack_rows(c, stream); if (r != PC_OK) return r; }
    if (nr) CUDA_TRY(cudaMemcpyAsync(h_rows, c->rows_packed.p, nr * (size_t)c->pack_bytes, cudaMemcpyDeviceToHost, s));
    if (rs) CUDA_TRY(cudaMemcpyAsync(rs, meta(c, 4), (size_t)c->last_nframes * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    if (rc) CUDA_TRY(cudaMemcpyAsync(rc, meta(c, 5), (size_t)c->last_nframes * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    return PC_OK;
}

extern "C" int pc_find_contacts(int device, const float *h_xyz1, int32_t n1, const float *h_xyz2, int32_t n2,
                                double cutoff, int64_t *h_row_ptr, int32_t **h_cols) {
    if (!h_row_ptr || !h_cols) return fail(PC_ERR_INVALID, "pc_find_contacts: NULL output");
    *h_cols = nullptr;
    for (int i = 0; i <= std::max(n1, 0); i++) h_row_ptr[i] = 0;
    if (n1 <= 0 || n2 <= 0) return PC_OK;            // gridsearch.C:906
    if (!h_xyz1 || !h_xyz2) return fail(PC_ERR_INVALID, "pc_find_contacts: NULL coordinates");
    if (!(cutoff > 0)) return fail(PC_ERR_INVALID, "pc_find_contacts: cutoff must be positive");
    CUDA_TRY(cudaSetDevice(device));
    return pc_search_all(h_xyz1, n1, h_xyz2, n2, cutoff, h_row_ptr, h_cols, g_err);
}

extern "C" void pc_free_host(void *p) { free(p); }

extern "C" int pc_synth_frames(int device, const float *d_base, const int32_t *d_res, const float *d_amp,
                               const float *d_period, const float *d_phase, int32_t natoms, float sigma,
                               uint64_t seed, int32_t frame0, int32_t nframes, int32_t layout, int64_t pitch,
                               float *d_out, void *stream) {
    if (!d_base || !d_res || !d_amp || !d_period || !d_phase || !d_out || natoms <= 0 || nframes <= 0)
        return fail(PC_ERR_INVALID, "pc_synth_frames: bad arguments");
    if (layout != PC_LAYOUT_XYZ && layout != PC_LAYOUT_XYZW) return fail(PC_ERR_INVALID, "pc_synth_frames: bad layout");
    CUDA_TRY(cudaSetDevice(device));
    pc_synth_launch(d_base, d_res, d_amp, d_period, d_phase, natoms, sigma, seed, frame0, nframes, layout, pitch,
                    d_out, (cudaStream_t)stream);
    CUDA_TRY(cudaGetLastError());
    g_launches += (unsigned long long)((nframes + 32767) / 32768);
    return PC_OK;
}

extern "C" int pc_stream_create(int device, void **out) {
    if (!out) return fail(PC_ERR_INVALID, "pc_stream_create: out is NULL");
    CUDA_TRY(cudaSetDevice(device));
    cudaStream_t s;
    CUDA_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *out = (void *)s;
    return PC_OK;
}
extern "C" int pc_stream_sync(void *stream) {
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return PC_OK;
}
extern "C" void pc_stream_destroy(void *stream) { if (stream) cudaStreamDestroy((cudaStream_t)stream); }

extern "C" int pc_device_alloc(int device, int64_t bytes, void **out) {
    if (!out || bytes <= 0) return fail(PC_ERR_INVALID, "pc_device_alloc: bad arguments");
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaMalloc(out, (size_t)bytes));
    return PC_OK;
}
extern "C" void pc_device_free(int device, void *p) {
    if (!p) return;
    cudaSetDevice(device);
    cudaFree(p);
}
extern "C" int pc_memcpy(int device, void *dst, const void *src, int64_t bytes, int kind, void *stream) {
    if (!dst || !src || bytes < 0 || kind < 1 || kind > 3) return fail(PC_ERR_INVALID, "pc_memcpy: bad arguments");
    CUDA_TRY(cudaSetDevice(device));
    const cudaMemcpyKind k = kind == 1 ? cudaMemcpyHostToDevice : kind == 2 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    CUDA_TRY(cudaMemcpyAsync(dst, src, (size_t)bytes, k, (cudaStream_t)stream));
    return PC_OK;
}
extern "C" int pc_memset(int device, void *dst, int value, int64_t bytes, void *stream) {
    if (!dst || bytes < 0) return fail(PC_ERR_INVALID, "pc_memset: bad arguments");
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaMemsetAsync(dst, value, (size_t)bytes, (cudaStream_t)stream));
    return PC_OK;
}
extern "C" int pc_device_sync(int device) {
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaDeviceSynchronize());
    return PC_OK;
}

extern "C" int64_t pc_launch_count(void) { return (int64_t)g_launches; }

extern "C" int pc_event_create(int device, void **out) {
    if (!out) return fail(PC_ERR_INVALID, "pc_event_create: out is NULL");
    CUDA_TRY(cudaSetDevice(device));
    cudaEvent_t e;
    CUDA_TRY(cudaEventCreate(&e));
    *out = (void *)e;
    return PC_OK;
}
extern "C" int pc_event_record(void *event, void *stream) {
    CUDA_TRY(cudaEventRecord((cudaEvent_t)event, (cudaStream_t)stream));
    return PC_OK;
}
extern "C" int pc_event_elapsed_ms(void *start, void *end, float *ms) {
    if (!ms) return fail(PC_ERR_INVALID, "pc_event_elapsed_ms: ms is NULL");
    CUDA_TRY(cudaEventSynchronize((cudaEvent_t)end));
    CUDA_TRY(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)end));
    return PC_OK;
}
extern "C" void pc_event_destroy(void *event) { if (event) cudaEventDestroy((cudaEvent_t)event); }

extern "C" int pc_host_alloc(void **out, int64_t bytes) {
    if (!out || bytes <= 0) return fail(PC_ERR_INVALID, "pc_host_alloc: bad arguments");
    CUDA_TRY(cudaHostAlloc(out, (size_t)bytes, cudaHostAllocDefault));
    return PC_OK;
}
extern "C" void pc_host_free(void *p) { if (p) cudaFreeHost(p); }

// ---- SASA and "within" (kernels in pc_sasa.cuh) ----------------------------------------------

// glibc's random() after srandom(seed) (the additive-feedback TYPE_3 generator: r[i] = r[i-31] + r[i-3], 310
// outputs discarded, top 31 bits returned), written out so that the library neither depends on nor disturbs
// the process-wide generator the reference reseeds through vmd_srandom (gridsearch.C:32-45).
static void glibc_random_sequence(unsigned seed, int count, long *out) {
    std::vector<int32_t> r(344 + (size_t)count);
    r[0] = seed ? (int32_t)seed : 1;
    for (int i = 1; i < 31; i++) {
        const long long hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
        long long w = 16807 * lo - 2836 * hi;
        if (w < 0) w += 2147483647;
        r[i] = (int32_t)w;
    }
    for (int i = 31; i < 34; i++) r[i] = r[i - 31];
    for (size_t i = 34; i < r.size(); i++) r[i] = (int32_t)((uint32_t)r[i - 31] + (uint32_t)r[i - 3]);
    for (int k = 0; k < count; k++) out[k] = (long)((uint32_t)r[344 + k] >> 1);
}

// The sphere points every atom shares (gridsearch.C:453-484): pointstyle != 0 -> random points from the fixed
// seed 38572111, 0 -> the spiral (whose k-1 indexing makes the first point NaN: it never counts as buried).
extern "C" int pc_sasa_points(int32_t npts, int32_t pointstyle, float *h_pts) {
    if (npts < 1 || !h_pts) return fail(PC_ERR_INVALID, "pc_sasa_points: bad arguments");
    if (pointstyle) {
        const float rand_max_inv = 1.0f / (float)2147483647L;
        std::vector<long> rnd(2 * (size_t)npts);
        glibc_random_sequence(38572111u, 2 * npts, rnd.data());
        for (int i = 0; i < npts; i++) {
            const float u1 = (float)rnd[2 * i], u2 = (float)rnd[2 * i + 1];
            const float z = 2.0f * u1 * rand_max_inv - 1.0f;
            const float phi = (float)((double)2.0f * M_PI * (double)u2 * (double)rand_max_inv);
            const float R = sqrtf(1.0f - z * z);
            h_pts[3 * i] = R * cosf(phi);
            h_pts[3 * i + 1] = R * sinf(phi);
            h_pts[3 * i + 2] = z;
        }
    } else {
        for (int k = 0; k < npts; k++) {
            const float h = (float)(2.0 * (k - 1.0) / (npts - 1.0) - 1.0);
            const float theta = acosf(h);
            float phi = 0.f;
            if (!(k == 1 || k == npts))
                phi = (float)fmod((double)phi + 3.6 / (double)sqrtf((float)npts * (1.0f - h * h)), 2.0 * M_PI);
            h_pts[3 * k] = cosf(phi) * sinf(theta);
            h_pts[3 * k + 1] = sinf(phi) * sinf(theta);
            h_pts[3 * k + 2] = cosf(theta);
        }
    }
    return PC_OK;
}

namespace {
struct StreamBuf {            // stream-ordered scratch: released behind the work that uses it
    void *p = nullptr;
    cudaStream_t s = 0;
    ~StreamBuf() { if (p) cudaFreeAsync(p, s); }
    cudaError_t alloc(size_t bytes, cudaStream_t stream) {
        s = stream;
        return cudaMallocAsync(&p, std::max<size_t>(bytes, 16), stream);
    }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct GridWork {             // cell list of a sub-batch of frames
    StreamBuf bbox, geom, cellcnt, cellid, rank, sidx, sorted;
    int cap_cells = 0, cell_pitch = 0;
    static size_t bytes_per_frame(int n, int cap_cells) {
        return (size_t)n * (16 + 4 + 4 + 4) + ((size_t)cap_cells + 1) * 4 + 6 * 4 + sizeof(PcGridGeom);
    }
    cudaError_t alloc(int fb, int n, int cap, cudaStream_t s) {
        cap_cells = cap; cell_pitch = cap + 1;
        cudaError_t e;
        if ((e = bbox.alloc((size_t)fb * 6 * 4, s)) != cudaSuccess) return e;
        if ((e = geom.alloc((size_t)fb * sizeof(PcGridGeom), s)) != cudaSuccess) return e;
        if ((e = cellcnt.alloc((size_t)fb * cell_pitch * 4, s)) != cudaSuccess) return e;
        if ((e = cellid.alloc((size_t)fb * n * 4, s)) != cudaSuccess) return e;
        if ((e = rank.alloc((size_t)fb * n * 4, s)) != cudaSuccess) return e;
        if ((e = sidx.alloc((size_t)fb * n * 4, s)) != cudaSuccess) return e;
        return sorted.alloc((size_t)fb * n * 16, s);
    }
};

int grid_cap_cells(int n) { return std::min(std::max(4096, n), 1 << 22); }
int grid_chunks(int n) { return std::min(std::max((n + 1023) / 1024, 1), 1024); }
// cell edge a hair above the search distance: relative slack for the float32 cell coordinate, absolute slack
// for the rounding of large coordinates (0.01 A covers |x| up to ~5e4 A)
float grid_edge(float dist) { return std::max(dist * 1.001f, dist + 0.01f); }

// bbox -> geometry -> count -> scan -> scatter for frames [0, fb) at xyz
cudaError_t grid_build(GridWork &g, const float *xyz, long long pitch, int fb, int n, const int *d_mask, const float *d_w,
                       float edge, cudaStream_t s, unsigned long long &launches) {
    const dim3 cg(grid_chunks(n), fb);
    pcg_init_kernel<<<(fb * 6 + 255) / 256, 256, 0, s>>>(g.bbox.as<unsigned>(), fb);
    cudaError_t e = cudaMemsetAsync(g.cellcnt.p, 0, (size_t)fb * g.cell_pitch * 4, s);
    if (e != cudaSuccess) return e;
    pcg_bbox_kernel<<<cg, 256, 0, s>>>(xyz, pitch, n, d_mask, g.bbox.as<unsigned>());
    pcg_geom_kernel<<<(fb + 127) / 128, 128, 0, s>>>(g.bbox.as<unsigned>(), fb, edge, g.cap_cells,
                                                     g.geom.as<PcGridGeom>());
    pcg_count_kernel<<<cg, 256, 0, s>>>(xyz, pitch, n, d_mask, g.geom.as<PcGridGeom>(), g.cell_pitch, g.cellcnt.as<int>(),
                                        g.cellid.as<int>(), g.rank.as<int>());
    pcg_scan_kernel<<<fb, 1024, 0, s>>>(g.geom.as<PcGridGeom>(), g.cell_pitch, g.cellcnt.as<int>());
    pcg_scatter_kernel<<<cg, 256, 0, s>>>(xyz, pitch, n, d_w, g.cell_pitch, g.cellcnt.as<int>(), g.cellid.as<int>(),
                                          g.rank.as<int>(), g.sorted.as<float4>(), g.sidx.as<int>());
    launches += 6;
    return cudaGetLastError();
}

const size_t kGridBudget = 1536ull << 20;     // scratch per sub-batch of frames

// The stream-ordered scratch comes from the device's default memory pool; by default the pool hands its memory
// back to the driver at every synchronisation, so each call would pay for fresh allocations.  Keep it.
cudaError_t grid_pool_keep(int device) {
    static bool done[64] = {};
    if (device < 0 || device >= 64 || done[device]) return cudaSuccess;
    cudaMemPool_t pool;
    cudaError_t e = cudaDeviceGetDefaultMemPool(&pool, device);
    if (e != cudaSuccess) return e;
    uint64_t keep = ~0ull;
    e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    if (e == cudaSuccess) done[device] = true;
    return e;
}
}  // namespace

extern "C" int pc_sasa_device(int device, const float *d_xyz, int32_t nframes, int32_t natoms, int64_t pitch,
                              const float *h_radius, float pairdist, int32_t npts, double srad, int32_t pointstyle,
                              int32_t restricted, const int32_t *h_restricted, float *d_total, int32_t *d_exposed,
                              uint64_t *d_evals, void *stream) {
    if (nframes < 0 || natoms < 0 || !d_total || (natoms && (!d_xyz || !h_radius)) || pitch < 3ll * natoms)
        return fail(PC_ERR_INVALID, "pc_sasa_device: bad arguments");
    if (npts < 1 || npts > PCS_MAX_POINTS) return fail(PC_ERR_INVALID, "pc_sasa_device: 1 <= surface points <= 1024");
    if (!(pairdist > 0)) return fail(PC_ERR_INVALID, "pc_sasa_device: pairdist must be positive");
    if (restricted && !h_restricted) return fail(PC_ERR_INVALID, "pc_sasa_device: restricted without a restriction list");
    if (nframes == 0) return PC_OK;
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(grid_pool_keep(device));
    cudaStream_t s = (cudaStream_t)stream;
    if (natoms == 0) { CUDA_TRY(cudaMemsetAsync(d_total, 0, (size_t)nframes * 4, s)); return PC_OK; }

    std::vector<float> radp((size_t)natoms), pts(3 * (size_t)npts);
    for (int i = 0; i < natoms; i++) radp[i] = (float)((double)h_radius[i] + srad);     // float rad = radius[i]+srad (:497)
    int rc = pc_sasa_points(npts, pointstyle, pts.data());
    if (rc != PC_OK) return rc;
    const float prefac = (float)(4 * M_PI / npts);
    const float sqdist = pairdist * pairdist;
    // candidate rows are kept while their gap to the atom is within `reach`: pairdist plus slack for the float32 cell
    // coordinates (relative) and for the rounding of large coordinates (absolute); cells are half of it wide
    const float reach = pairdist * 1.001f + 0.02f;

    StreamBuf d_radp, d_pts, d_rl, d_area;
    CUDA_TRY(d_radp.alloc((size_t)natoms * 4, s));
    CUDA_TRY(d_pts.alloc(pts.size() * 4, s));
    CUDA_TRY(cudaMemcpyAsync(d_radp.p, radp.data(), (size_t)natoms * 4, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(d_pts.p, pts.data(), pts.size() * 4, cudaMemcpyHostToDevice, s));
    if (restricted) {
        CUDA_TRY(d_rl.alloc((size_t)natoms * 4, s));
        CUDA_TRY(cudaMemcpyAsync(d_rl.p, h_restricted, (size_t)natoms * 4, cudaMemcpyHostToDevice, s));
    }
    CUDA_TRY(cudaFuncSetAttribute(pc_sasa_atoms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pc_sasa_smem(PCS_MAX_POINTS)));
    const int cap = grid_cap_cells(natoms);
    const size_t per_frame = GridWork::bytes_per_frame(natoms, cap) + (size_t)natoms * 4;
    const int fb = (int)std::min<size_t>(std::min<size_t>(std::max<size_t>(kGridBudget / per_frame, 1), 32768), (size_t)nframes);
    GridWork g;
    CUDA_TRY(g.alloc(fb, natoms, cap, s));
    CUDA_TRY(d_area.alloc((size_t)fb * natoms * 4, s));
    for (int f0 = 0; f0 < nframes; f0 += fb) {
        const int nf = std::min(fb, nframes - f0);
        const float *xyz = d_xyz + (long long)f0 * pitch;
        CUDA_TRY(grid_build(g, xyz, pitch, nf, natoms, nullptr, d_radp.as<float>(), 0.5f * reach * 1.0001f, s, g_launches));
        pc_sasa_atoms_kernel<<<dim3((natoms + PCS_WARPS - 1) / PCS_WARPS, nf), PCS_WARPS * 32, pc_sasa_smem(npts), s>>>(
            g.sorted.as<float4>(), g.sidx.as<int>(), g.cellcnt.as<int>(), g.geom.as<PcGridGeom>(), natoms, g.cell_pitch,
            d_pts.as<float>(), npts, sqdist, reach, restricted, d_rl.as<int>(), prefac, d_area.as<float>(),
            d_exposed ? d_exposed + (long long)f0 * natoms : nullptr, (unsigned long long *)d_evals);
        pc_sasa_sum_kernel<<<nf, 256, 0, s>>>(d_area.as<float>(), natoms, d_total + f0);
        CUDA_TRY(cudaGetLastError());
        g_launches += 2;
    }
    return PC_OK;
}

extern "C" int pc_sasa(int device, const float *h_xyz, int32_t nframes, int32_t natoms, const float *h_radius,
                       float pairdist, int32_t npts, double srad, int32_t pointstyle, int32_t restricted,
                       const int32_t *h_restricted, float *h_total, int32_t *h_exposed) {
    if (nframes < 0 || natoms < 0 || !h_total || (nframes && natoms && !h_xyz))
        return fail(PC_ERR_INVALID, "pc_sasa: bad arguments");
    if (nframes == 0) return PC_OK;
    CUDA_TRY(cudaSetDevice(device));
    const size_t frame_bytes = (size_t)natoms * 12;
    const int fb = (int)std::min<size_t>(std::max<size_t>((1ull << 30) / std::max<size_t>(frame_bytes, 1), 1), (size_t)nframes);
    CUDA_TRY(grid_pool_keep(device));
    StreamBuf d_xyz, d_total, d_exp;                  // stream-ordered like the kernels' scratch (same pool, no cudaMalloc)
    CUDA_TRY(d_xyz.alloc((size_t)fb * natoms * 12, 0));
    CUDA_TRY(d_total.alloc((size_t)fb * 4, 0));
    if (h_exposed) CUDA_TRY(d_exp.alloc((size_t)fb * natoms * 4, 0));
    for (int f0 = 0; f0 < nframes; f0 += fb) {
        const int nf = std::min(fb, nframes - f0);
        if (natoms) CUDA_TRY(cudaMemcpyAsync(d_xyz.p, h_xyz + (size_t)f0 * natoms * 3, (size_t)nf * frame_bytes, cudaMemcpyHostToDevice, 0));
        const int rc = pc_sasa_device(device, d_xyz.as<float>(), nf, natoms, 3ll * natoms, h_radius, pairdist, npts, srad, pointstyle,
                                      restricted, h_restricted, d_total.as<float>(), h_exposed ? d_exp.as<int32_t>() : nullptr, nullptr,
                                      nullptr);
        if (rc != PC_OK) return rc;
        CUDA_TRY(cudaMemcpyAsync(h_total + f0, d_total.p, (size_t)nf * 4, cudaMemcpyDeviceToHost, 0));
        if (h_exposed && natoms)
            CUDA_TRY(cudaMemcpyAsync(h_exposed + (size_t)f0 * natoms, d_exp.p, (size_t)nf * natoms * 4, cudaMemcpyDeviceToHost, 0));
        CUDA_TRY(cudaStreamSynchronize(0));
    }
    return PC_OK;
}

extern "C" int pc_within_device(int device, const float *d_xyz, int32_t nframes, int32_t num, int64_t pitch,
                                const int32_t *d_flgs, const int32_t *d_others, float r, int32_t *d_result, void *stream) {
    if (nframes < 0 || num < 0 || (nframes && num && (!d_xyz || !d_flgs || !d_others || !d_result)) || pitch < 3ll * num)
        return fail(PC_ERR_INVALID, "pc_within_device: bad arguments");
    if (!(r > 0)) return fail(PC_ERR_INVALID, "pc_within_device: r must be positive");
    if (nframes == 0 || num == 0) return PC_OK;
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(grid_pool_keep(device));
    cudaStream_t s = (cudaStream_t)stream;
    const int cap = grid_cap_cells(num);
    const size_t per_frame = GridWork::bytes_per_frame(num, cap);
    const int fb = (int)std::min<size_t>(std::min<size_t>(std::max<size_t>(kGridBudget / per_frame, 1), 32768), (size_t)nframes);
    GridWork g;
    CUDA_TRY(g.alloc(fb, num, cap, s));
    const float r2 = r * r;                                   // const float r2 = (float)(r*r)  (gridsearch.C:677)
    for (int f0 = 0; f0 < nframes; f0 += fb) {
        const int nf = std::min(fb, nframes - f0);
        const float *xyz = d_xyz + (long long)f0 * pitch;
        CUDA_TRY(grid_build(g, xyz, pitch, nf, num, d_others, nullptr, grid_edge(r), s, g_launches));
        pc_within_kernel<<<dim3(grid_chunks(num), nf), 256, 0, s>>>(xyz, pitch, num, d_flgs, g.sorted.as<float4>(),
                                                                     g.cellcnt.as<int>(), g.geom.as<PcGridGeom>(), g.cell_pitch,
                                                                     r2, d_result + (long long)f0 * num);
        CUDA_TRY(cudaGetLastError());
        g_launches += 1;
    }
    return PC_OK;
}

extern "C" int pc_find_within(int device, const float *h_xyz, int32_t nframes, int32_t num, const int32_t *h_flgs,
                              const int32_t *h_others, float r, int32_t *h_result) {
    if (nframes < 0 || num < 0 || (nframes && num && (!h_xyz || !h_flgs || !h_others || !h_result)))
        return fail(PC_ERR_INVALID, "pc_find_within: bad arguments");
    if (nframes == 0 || num == 0) return PC_OK;
    CUDA_TRY(cudaSetDevice(device));
    const size_t frame_bytes = (size_t)num * 12;
    const int fb = (int)std::min<size_t>(std::max<size_t>((1ull << 30) / frame_bytes, 1), (size_t)nframes);
    CUDA_TRY(grid_pool_keep(device));
    StreamBuf d_xyz, d_f, d_o, d_res;
    CUDA_TRY(d_xyz.alloc((size_t)fb * num * 12, 0));
    CUDA_TRY(d_f.alloc((size_t)num * 4, 0));
    CUDA_TRY(d_o.alloc((size_t)num * 4, 0));
    CUDA_TRY(d_res.alloc((size_t)fb * num * 4, 0));
    CUDA_TRY(cudaMemcpyAsync(d_f.p, h_flgs, (size_t)num * 4, cudaMemcpyHostToDevice, 0));
    CUDA_TRY(cudaMemcpyAsync(d_o.p, h_others, (size_t)num * 4, cudaMemcpyHostToDevice, 0));
    for (int f0 = 0; f0 < nframes; f0 += fb) {
        const int nf = std::min(fb, nframes - f0);
        CUDA_TRY(cudaMemcpyAsync(d_xyz.p, h_xyz + (size_t)f0 * num * 3, (size_t)nf * frame_bytes, cudaMemcpyHostToDevice, 0));
        const int rc = pc_within_device(device, d_xyz.as<float>(), nf, num, 3ll * num, d_f.as<int32_t>(), d_o.as<int32_t>(), r,
                                        d_res.as<int32_t>(), nullptr);
        if (rc != PC_OK) return rc;
        CUDA_TRY(cudaMemcpyAsync(h_result + (size_t)f0 * num, d_res.p, (size_t)nf * num * 4, cudaMemcpyDeviceToHost, 0));
        CUDA_TRY(cudaStreamSynchronize(0));
    }
    return PC_OK;
}

// ---- XTC trajectory frames (decoder in pc_xtc.h, host code) -----------------------------------------
extern "C" int pc_xtc_open(const char *path, pc_xtc **out) {
    if (!path || !out) return fail(PC_ERR_INVALID, "pc_xtc_open: bad arguments");
    FILE *fh = fopen(path, "rb");
    if (!fh) return fail(PC_ERR_INVALID, std::string("pc_xtc_open: cannot open ") + path);
    pc_xtc *x = new pc_xtc();
    x->fh = fh;
    // index the frames: every header tells how long its frame is
    unsigned char hdr[92];
    long long pos = 0;
    for (;;) {
        if (fseeko(fh, (off_t)pos, SEEK_SET) != 0) break;
        memset(hdr, 0, sizeof hdr);
        const size_t got = fread(hdr, 1, sizeof hdr, fh);
        if (got < 56) break;
        int natoms = 0;
        const long long nb = pcxtc::frame_bytes(hdr, natoms);
        if (nb == 0 || (natoms > 9 && got < 92)) {
            if (x->offset.empty()) { fclose(fh); delete x; return fail(PC_ERR_INVALID, std::string("pc_xtc_open: not an XTC file: ") + path); }
            break;                                   // trailing bytes of an interrupted write
        }
        if (x->offset.empty()) x->natoms = natoms;
        else if (natoms != x->natoms) { fclose(fh); delete x; return fail(PC_ERR_INVALID, "pc_xtc_open: the number of atoms changes between frames"); }
        x->offset.push_back(pos);
        pos += nb;
    }
    fseeko(fh, 0, SEEK_END);
    const long long fsize = (long long)ftello(fh);
    if (!x->offset.empty() && pos > fsize) x->offset.pop_back();          // last frame cut short
    if (x->offset.empty()) { fclose(fh); delete x; return fail(PC_ERR_INVALID, std::string("pc_xtc_open: no complete frame in ") + path); }
    x->offset.push_back(std::min(pos, fsize));
    x->xyz.resize((size_t)x->natoms * 3);
    *out = x;
    return PC_OK;
}

extern "C" int pc_xtc_info(const pc_xtc *x, int32_t *natoms, int64_t *nframes) {
    if (!x) return fail(PC_ERR_INVALID, "pc_xtc_info: handle is NULL");
    if (natoms) *natoms = x->natoms;
    if (nframes) *nframes = (int64_t)x->offset.size() - 1;
    return PC_OK;
}

extern "C" int pc_xtc_read(pc_xtc *x, int64_t frame0, int32_t nframes, const int32_t *indices, int32_t nidx, float *h_xyz,
                           int32_t *h_step, float *h_time, float *h_box) {
    if (!x || !h_xyz || nframes < 0 || frame0 < 0 || frame0 + nframes > (int64_t)x->offset.size() - 1)
        return fail(PC_ERR_INVALID, "pc_xtc_read: bad frame range");
    if (indices && nidx < 0) return fail(PC_ERR_INVALID, "pc_xtc_read: bad selection");
    const int nout = indices ? nidx : x->natoms;
    if (indices)
        for (int k = 0; k < nidx; k++)
            if (indices[k] < 0 || indices[k] >= x->natoms) return fail(PC_ERR_INVALID, "pc_xtc_read: atom index out of range");
    for (int32_t f = 0; f < nframes; f++) {
        const long long o0 = x->offset[(size_t)(frame0 + f)], o1 = x->offset[(size_t)(frame0 + f) + 1];
        x->raw.resize((size_t)(o1 - o0));
        if (fseeko(x->fh, (off_t)o0, SEEK_SET) != 0 || fread(x->raw.data(), 1, x->raw.size(), x->fh) != x->raw.size())
            return fail(PC_ERR_INVALID, "pc_xtc_read: short read");
        float box[9];
        int step = 0;
        float time = 0.f;
        const std::string err = pcxtc::decode(x->raw.data(), x->raw.size(), x->natoms, x->xyz.data(), &step, &time, box);
        if (!err.empty()) return fail(PC_ERR_INVALID, "pc_xtc_read: frame " + std::to_string(frame0 + f) + ": " + err);
        // nm -> Angstrom in float32, like MDAnalysis' XTC reader hands positions to the reference
        float *dst = h_xyz + (size_t)f * nout * 3;
        const float *src = x->xyz.data();
        if (indices) {
            for (int k = 0; k < nidx; k++) {
                const float *p = src + 3 * (size_t)indices[k];
                dst[3 * k] = p[0] * 10.0f; dst[3 * k + 1] = p[1] * 10.0f; dst[3 * k + 2] = p[2] * 10.0f;
            }
        } else {
            for (size_t k = 0; k < (size_t)x->natoms * 3; k++) dst[k] = src[k] * 10.0f;
        }
        if (h_step) h_step[f] = step;
        if (h_time) h_time[f] = time;
        if (h_box) for (int k = 0; k < 9; k++) h_box[(size_t)f * 9 + k] = box[k] * 10.0f;
    }
    return PC_OK;
}

extern "C" void pc_xtc_close(pc_xtc *x) {
    if (!x) return;
    if (x->fh) fclose(x->fh);
    delete x;
}
