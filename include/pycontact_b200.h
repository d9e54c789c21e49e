/*
 * pycontact_b200.h -- C ABI of the B200 contact-analysis library
 * (libpycontact_b200.so, CUDA sm_100a).
 *
 * This is the drop-in boundary for the per-frame hot path of PyContact.  Each entry
 * point names the reference interface it replaces (paths relative to the
 * reference root).  Plain pointers and sizes only; no torch / numpy types.
 *
 * Conventions
 *   - every function returns 0 on success, a negative PC_ERR_* code otherwise;
 *     pc_last_error() gives a message for the calling thread's last failure;
 *   - one pc_ctx per GPU; a context is not thread-safe, different contexts are;
 *   - "local index" = position of an atom inside its selection (the idx1 / idx2
 *     that cy_find_contacts returns); the Python layer converts to global indices;
 *   - pointers named h_* are host memory, d_* device memory of the context's GPU;
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default
 *     stream); all work is enqueued on it and nothing synchronises unless stated.
 */
#ifndef PYCONTACT_B200_H
#define PYCONTACT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PC_OK                 0
#define PC_ERR_INVALID       -1   /* bad argument / call order */
#define PC_ERR_CUDA          -2   /* CUDA runtime error (message has the detail) */
#define PC_ERR_CAPACITY      -3   /* an output buffer was too small; see pc_scan_status */
#define PC_ERR_MISSING_H     -4   /* a hydrogen bonded to a selected N/O/S atom is not in that
                                     selection: the reference raises here too
                                     (core/ContactAnalyzer.py:432, core/multi_trajectory.py:158) */
#define PC_ERR_NOMEM         -5
#define PC_ST_STAGE_OVERFLOW_BIT 64 /* pc_last_status: small-frame staging too small -> use pc_set_path(ctx, 2) */
#define PC_ST_TAIL_OVERFLOW_BIT 128 /* large-frame path: a frame does not fit the fused tail kernel -> pc_set_tail(ctx, 0) */
#define PC_ST_PRECISION_BIT     256 /* a (frame, key) cell has > 1024 contacts: float32 cell sums would leave the 1e-5
                                       bar -> pc_set_limits(..., table_slots = 16384) selects the exact global table */

/* per-atom flag bits (pc_set_topology) */
#define PC_ATOM_HYDROGEN   1u   /* name starts with "H"   (core/ContactAnalyzer.py:360)      */
#define PC_ATOM_HBCLASS    2u   /* name[0] in {O,N,S}     (core/Biochemistry.py:32-33)       */
#define PC_ATOM_BACKBONE   4u   /* index in MDAnalysis "backbone" (core/ContactAnalyzer.py:287) */

/* coordinate layouts of a frame batch in device / host memory */
#define PC_LAYOUT_XYZ   3   /* packed float32 x,y,z   (numpy (F,n,3), what sel.positions gives) */
#define PC_LAYOUT_XYZW  4   /* float4 per atom, w ignored                                       */

/* One AtomContact (core/Biochemistry.py:299-312) without the Python object:
 * i, j are local indices into selection 1 / selection 2. */
typedef struct { int32_t i, j; float distance, weight; } pc_contact;

/* One HydrogenBond (core/Biochemistry.py:315-331).  (i, j) is the contact it belongs
 * to; `hyd` is the hydrogen's local index in the donor's selection; bit 31 of `hyd`
 * set => the donor is atom j (selection 2), clear => the donor is atom i. */
typedef struct { int32_t i, j; uint32_t hyd; float distance, angle; } pc_hbond;

/* One (frame, key) cell of the accumulation
 * (TempContactAccumulate, core/Biochemistry.py:283-296): key = gid1*G2 + gid2. */
typedef struct {
    uint64_t key;
    double   score;        /* fscore: sum of weights                                   */
    double   bb1, bb2;     /* backbone part of the score for atom 1 / atom 2 (sc = score - bb) */
    uint32_t ncontacts;    /* len(contributingAtomContacts)                            */
    uint32_t nhbonds;      /* sum of len(hbondinfo) -> AccumulatedContact.hbondFramesScan */
    uint64_t first;        /* smallest order key of a contributing contact (first appearance) */
} pc_row;

/* The same cell in 24 bytes for bulk transfer to the host: float32 sums (the 1e-5 bar leaves room),
 * counts saturating at 65535, no first-appearance key (callers that need the reference's key ORDER
 * fetch pc_row).  Everything AccumulatedContact keeps per frame -- scoreArray, hbondFrames -- and the
 * per-frame parts of its bb/sc totals are in here. */
typedef struct {
    uint64_t key;
    float    score, bb1, bb2;
    uint16_t ncontacts, nhbonds;
} pc_row_packed;

/* 12-byte variant for key spaces below 2^32 (residue-pair maps of all but the largest systems) when the
 * caller keeps the per-key bb sums elsewhere (pc_fold_rows' aggregated table): key, score, counts. */
typedef struct {
    uint32_t key;
    float    score;
    uint16_t ncontacts, nhbonds;
} pc_row_slim;

typedef struct pc_ctx pc_ctx;

const char *pc_last_error(void);
int pc_version(void);
int pc_device_count(void);   /* CUDA devices visible to this process (0 when there is none) */

/* Analyzer(psf, dcd, cutoff, hbondcutoff, hbondcutangle, sel1text, sel2text)
 * (core/ContactAnalyzer.py:23-46): the numeric part of the constructor.
 * self_mode != 0  <=>  sel2text == "self" (core/ContactAnalyzer.py:304-307). */
int pc_create(int device, double cutoff, double hb_cutoff, double hb_angle_deg, int self_mode,
              pc_ctx **out);
void pc_destroy(pc_ctx *ctx);

/* Static per-selection data that the reference re-derives per raw pair in its Python loop
 * (core/multi_trajectory.py:83-155): flags, residue ids / segment ids (self-interaction
 * filter, :93-95) and the bonded-hydrogen CSR (:115-155, local indices; -1 marks a hydrogen
 * outside the selection).  In self mode the *2 arrays are ignored.  Host pointers, copied. */
int pc_set_topology(pc_ctx *ctx, int32_t n1, int32_t n2,
                    const uint8_t *h_flags1, const uint8_t *h_flags2,
                    const int32_t *h_resid1, const int32_t *h_segid1,
                    const int32_t *h_resid2, const int32_t *h_segid2,
                    const int32_t *h_hptr1, const int32_t *h_hidx1,
                    const int32_t *h_hptr2, const int32_t *h_hidx2);

/* Accumulation maps (Analyzer.runContactAnalysis(map1, map2), core/ContactAnalyzer.py:64-70,
 * makeKeyArraysFromMaps :107-158): group id of every selection atom under map1 / map2. */
int pc_set_maps(pc_ctx *ctx, const int32_t *h_gid1, const int32_t *h_gid2, int64_t n_groups2);

/* Optional per-frame limits of the kernels' on-chip staging (0 = automatic): contacts and H-bonds
 * one frame may produce on the small-frame path, grid cells, accumulation-table slots. */
int pc_set_limits(pc_ctx *ctx, int32_t hits_per_frame, int32_t hbonds_per_frame, int32_t cells,
                  int32_t table_slots);

/* Kernel path: 0 = automatic (by selection size), 1 = one CTA per frame out of shared memory,
 * 2 = tiled large-frame path.  Both give the same records; tests use this to cover both. */
int pc_set_path(pc_ctx *ctx, int path);

/* Small-frame kernel: CTAs per SM.  0 = automatic (a "fast" plan of two 512-thread CTAs per SM -- three of 384 threads
 * for selections of at most 3000 heavy atoms together; frames whose
 * in-grid atoms or hits do not fit it are redone by the "full" plan, one 1024-thread CTA), 1 = full plan
 * only, 2 = fast plan only (PC_ST_STAGE_OVERFLOW when a frame does not fit), 3 / 4 = fast plan with
 * 3 x 384 / 4 x 256 threads per SM, then the full plan. */
int pc_set_small_ctas(pc_ctx *ctx, int ctas);

/* Large-frame path, what follows the streaming kernels.  mode 3 (default) and 1: ONE fused kernel, a CTA per frame
 * (cell sort, pair search, records, H-bond test and -- with maps, without order keys -- the accumulation, all out of
 * shared memory); 3 reads selection 1 straight from the frame (no bin pass for it), 1 works on
 * the compacted atoms of both selections.  0 = the multi-kernel sequence, which has no limit on a frame's in-grid
 * atoms.  A batch whose frames do not fit a fused kernel reports PC_ERR_CAPACITY with status bit 128; callers then
 * select 0 and re-run (ContactEngine does).  Same records either way (vmd_gridsearch3,
 * cy_modules/src/gridsearch.C:1111-1142). */
int pc_set_tail(pc_ctx *ctx, int mode);

/* Multi-kernel sequence only, for tests: which pair kernel searches a frame -- 0 = chosen per frame on the device
 * by B atoms per cell, 1 = thread per (A atom, layer), 2 = warp per (A cell, layer) with the B runs in lanes. */
int pc_set_dense_mode(pc_ctx *ctx, int mode);

/* Large-frame path, frames the fused tail kernel does not take (self contacts, frames handed over with status bit
 * 128): on = 1 (default) searches, records and accumulates in ONE kernel whose CTAs own whole accumulation groups of
 * selection 1 (runs of consecutive atoms with one group id, e.g. residues) and emit their (frame, key) rows from a
 * shared-memory table -- used when maps are set, no order keys are asked for and every group is one run of at most 256
 * heavy atoms; on = 0 keeps the multi-kernel sequence + global table (analyze_contactResultsWithMaps,
 * core/ContactAnalyzer.py:527-559, either way). */
int pc_set_group_mode(pc_ctx *ctx, int on);
/* 1 when the context's last scan ran the group kernel (tests). */
int pc_last_group(const pc_ctx *ctx);

/* Output capacities for one batch (totals over the batch, not per frame). */
int pc_reserve(pc_ctx *ctx, int32_t max_frames, int64_t cap_contacts, int64_t cap_hbonds, int64_t cap_rows);

/* The frame loop body for a batch of frames resident in HBM
 * (loop_trajectory_grid, core/multi_trajectory.py:50-222; serial twin
 * core/ContactAnalyzer.py:319-493): neighbour search + heavy-atom / self filters +
 * distance + weight + H-bond test for `nframes` frames.  d_xyz2 is ignored in self mode.
 * pitch* = floats between consecutive frames.  order_keys != 0 additionally computes the
 * reference's in-frame order key per contact. */
int pc_scan_device(pc_ctx *ctx, const float *d_xyz1, const float *d_xyz2, int32_t nframes,
                   int32_t layout, int64_t pitch1, int64_t pitch2, int32_t order_keys, void *stream);

/* analyze_contactResultsWithMaps (core/ContactAnalyzer.py:502-597), per-frame part:
 * groups the contacts of the last scan by key into pc_row's. */
int pc_accumulate(pc_ctx *ctx, void *stream);

/* Time-aggregated maps for the end-of-run reduction over ranks: adds the rows of the last
 * accumulated batch into d_totals[n_keys][4] = (sum score, sum bb1, sum bb2, frames with an
 * H-bond), key = gid1 * n_groups2 + gid2 < n_keys (the sums behind mean_score, bb/sc and
 * hbond_percentage, core/Biochemistry.py:150-186, core/ContactAnalyzer.py:588-591). */
int pc_fold_rows(pc_ctx *ctx, double *d_totals, int64_t n_keys, void *stream);

/* Per-key statistics over the (keys x frames) score matrix, all keys in one launch (one CTA per key):
 * d_out[n_keys][8] = mean score, median score, hbond_percentage, total_time, mean_life_time,
 * median_life_time, number of life-time entries, frames above the threshold -- AccumulatedContact.mean_score /
 * median_score / hbond_percentage / total_time / mean_life_time / median_life_time
 * (core/Biochemistry.py:150-213) for every contact at once, which is what the reference's filters and
 * sorts (core/ContactFilters.py:185-287) recompute per object.  d_scores is row-major float64
 * [n_keys][n_frames] (scoreArray of each key); d_hbonds (may be NULL) holds the H-bonds per (key, frame)
 * (hbondFramesScan).  n_frames <= 16384 per call. */
int pc_key_stats(int device, const double *d_scores, const uint32_t *d_hbonds, int64_t n_keys, int32_t n_frames,
                 double ns_per_frame, double threshold, double *d_out, void *stream);

/* The same statistics straight from the (frame, key) cells the accumulation emits, grouped by key (CSR): cells
 * d_key_ptr[k] .. d_key_ptr[k + 1] belong to key k, each with its frame position (0 <= frame < n_frames), its score
 * (AccumulatedContact.scoreArray[frame], core/ContactAnalyzer.py:561-577; frames without a cell score 0, :571-575) and its
 * H-bond count (may be NULL).  The (keys x frames) matrix is assembled per key in shared memory and never exists in HBM or
 * on the host.  Same output table and limits as pc_key_stats. */
int pc_key_stats_rows(int device, const int64_t *d_key_ptr, const int32_t *d_row_frame, const double *d_row_score,
                      const uint32_t *d_row_hbonds, int64_t n_keys, int32_t n_frames, double ns_per_frame, double threshold,
                      double *d_out, void *stream);

/* The same aggregation for key spaces too large for a dense table: a per-context open-addressing table
 * (key -> sum score, sum bb1, sum bb2, frames with an H-bond) that lives across batches.
 * pc_fold_sparse_init sizes it (entries, rounded up to a power of two) and clears it;
 * pc_fold_rows_sparse adds the last accumulated batch; pc_fold_sparse_fetch compacts the occupied entries
 * on the device, copies them to h_entries (room for max_entries) and reports their number -- it
 * synchronises the stream and fails with PC_ERR_CAPACITY when the table overflowed. */
typedef struct { uint64_t key; double score, bb1, bb2, hbframes; } pc_fold_entry;
int pc_fold_sparse_init(pc_ctx *ctx, int64_t entries, void *stream);
int pc_fold_rows_sparse(pc_ctx *ctx, void *stream);
/* Several contexts of one GPU (batches in flight on different streams) can fold into ONE table: ctx then
 * adds its rows to owner's table (owner = NULL undoes it).  Synchronise ctx's stream before fetching. */
int pc_fold_sparse_share(pc_ctx *ctx, pc_ctx *owner);
int pc_fold_sparse_fetch(pc_ctx *ctx, void *stream, pc_fold_entry *h_entries, int64_t max_entries, int64_t *n_entries);
/* Frame-sharded runs (one rank per GPU, core/multi_trajectory.py:423-451 with ranks for pool workers): the
 * ranks' sparse maps meet in ONE all-gather of fixed-size buffers.  export writes the table's occupied entries,
 * compacted, into d_out[0 .. max_entries) and pads the rest with key = 2^64-1 (more entries than that: the
 * overflow flag that pc_fold_sparse_fetch reports); import folds such a buffer (of another rank) into this
 * context's table.  Asynchronous on `stream`; device pointers. */
int pc_fold_sparse_export(pc_ctx *ctx, void *stream, pc_fold_entry *d_out, int64_t max_entries);
int pc_fold_sparse_import(pc_ctx *ctx, void *stream, const pc_fold_entry *d_in, int64_t n_entries);

/* Analyzer.runContactAnalysis(map1, map2) may be called again with other maps on the SAME
 * contactResults (core/ContactAnalyzer.py:64-70): loads previously fetched records of `nframes`
 * frames back into the context so that pc_accumulate can regroup them.  Synchronous. */
int pc_load_records(pc_ctx *ctx, const pc_contact *h_contacts, const uint64_t *h_order_keys,
                    const pc_hbond *h_hbonds, int32_t nframes,
                    const uint32_t *h_frame_cstart, const uint32_t *h_frame_ccount,
                    const uint32_t *h_frame_hstart, const uint32_t *h_frame_hcount, void *stream);

/* By default the small-frame scan kernel accumulates on the fly when the maps are set and no order
 * keys are requested (pc_accumulate is then a no-op for that batch); 0 switches that off so that the
 * stand-alone accumulation kernel runs (tests compare the two). */
int pc_set_fused_accumulate(pc_ctx *ctx, int on);

/* Same two steps for frames in HOST memory (pinned or pageable): H2D copy, scan,
 * accumulate if maps are set; all asynchronous on `stream`. */
int pc_scan_host(pc_ctx *ctx, const float *h_xyz1, const float *h_xyz2, int32_t nframes,
                 int32_t layout, int32_t order_keys, int32_t accumulate, void *stream);

/* Synchronises `stream` and reports the totals of the last batch and its status
 * (0, PC_ERR_CAPACITY or PC_ERR_MISSING_H). */
int pc_batch_totals(pc_ctx *ctx, void *stream, int64_t *n_contacts, int64_t *n_hbonds, int64_t *n_rows,
                    int64_t *n_pair_evals);

/* Raw status of the last pc_batch_totals: PC_ST_* bits, the record counts the batch NEEDED (for
 * re-reserving after PC_ERR_CAPACITY) and which kernel path ran (1 = one CTA per frame,
 * 2 = tiled large-frame path). */
int pc_last_status(pc_ctx *ctx, uint64_t *status, int64_t *need_contacts, int64_t *need_hbonds,
                   int64_t *need_rows, int32_t *path);

/* Device pointers of the last batch's outputs (valid until the next scan / reserve). */
int pc_device_outputs(pc_ctx *ctx, pc_contact **d_contacts, pc_hbond **d_hbonds, pc_row **d_rows,
                      uint64_t **d_order_keys,
                      uint32_t **d_frame_cstart, uint32_t **d_frame_ccount,
                      uint32_t **d_frame_hstart, uint32_t **d_frame_hcount,
                      uint32_t **d_frame_rstart, uint32_t **d_frame_rcount);

/* D2H of the last batch's outputs into caller buffers sized from pc_batch_totals
 * (any pointer may be NULL to skip it).  frame_* arrays have nframes entries.
 * Asynchronous on `stream`; synchronise before reading. */
int pc_fetch(pc_ctx *ctx, void *stream,
             pc_contact *h_contacts, pc_hbond *h_hbonds, pc_row *h_rows, uint64_t *h_order_keys,
             uint32_t *h_frame_cstart, uint32_t *h_frame_ccount,
             uint32_t *h_frame_hstart, uint32_t *h_frame_hcount,
             uint32_t *h_frame_rstart, uint32_t *h_frame_rcount);

/* Rows of the last accumulated batch as pc_row_packed (packed on the device, then copied): the
 * transfer format of throughput runs, half the bytes of pc_row.  h_rows needs room for the n_rows that
 * pc_batch_totals reported.  Asynchronous on `stream`. */
int pc_pack_rows(pc_ctx *ctx, void *stream);     /* the packing kernel alone (enqueue it right behind pc_accumulate) */
int pc_set_pack_format(pc_ctx *ctx, int bytes_per_row);   /* 24 = pc_row_packed (default), 12 = pc_row_slim */
int pc_fetch_packed_rows(pc_ctx *ctx, void *stream, pc_row_packed *h_rows,
                         uint32_t *h_frame_rstart, uint32_t *h_frame_rcount);

/* cy_find_contacts(xyz1, natoms1, xyz2, natoms2, r)  (cy_modules/cy_gridsearch.pyx:31-35 ->
 * find_contacts, cy_modules/src/gridsearch.C:408-427): ALL atoms, no filters.  Returns a CSR:
 * h_row_ptr[n1+1] (caller allocated); *h_cols is allocated by the library (pc_free_host) with
 * h_row_ptr[n1] entries, each row in the reference's order (stencil rank, then index).
 * Host pointers; synchronous. */
int pc_find_contacts(int device, const float *h_xyz1, int32_t n1, const float *h_xyz2, int32_t n2,
                     double cutoff, int64_t *h_row_ptr, int32_t **h_cols);
void pc_free_host(void *p);

/* ---- SASA and "within" (the other two routines of the reference's cy_modules) ---- */

/* The sphere points all atoms share (sasa_grid, cy_modules/src/gridsearch.C:453-484): pointstyle != 0 ->
 * random points from the reference's fixed seed (srandom(38572111) / random()), 0 -> its spiral.
 * h_pts receives 3 * npts floats.  Host-only (no GPU needed). */
int pc_sasa_points(int32_t npts, int32_t pointstyle, float *h_pts);

/* cy_sasa(npcoords, natoms, pairdist, 0, -1, nprad, surfacePoints, probeRadius, pointstyle, restricted,
 * restrictedList)  (cy_modules/cy_gridsearch.pyx:14-22 -> sasa_grid, gridsearch.C:430-529) for a BATCH of
 * frames of one selection -- the loop of calculate_sasa_parallel (gui/SasaWidgets.py:27-49) in one call.
 * h_xyz: nframes x natoms x 3 float32; h_radius[natoms]; h_restricted[natoms] is read only when
 * restricted != 0.  h_total[nframes] = the reference's float32 total per frame (same summation order);
 * h_exposed (may be NULL) = nframes x natoms exposed sphere points per atom (0 for atoms outside the
 * restriction).  Host pointers; synchronous. */
int pc_sasa(int device, const float *h_xyz, int32_t nframes, int32_t natoms, const float *h_radius,
            float pairdist, int32_t npts, double srad, int32_t pointstyle, int32_t restricted,
            const int32_t *h_restricted, float *h_total, int32_t *h_exposed);

/* The same for frames resident in HBM (pitch = floats between frames); outputs in device memory;
 * d_evals (may be NULL) is incremented by the number of point-sphere tests done.  Asynchronous on `stream`;
 * scratch memory is stream-ordered (cudaMallocAsync).  radius / restriction are host arrays. */
int pc_sasa_device(int device, const float *d_xyz, int32_t nframes, int32_t natoms, int64_t pitch,
                   const float *h_radius, float pairdist, int32_t npts, double srad, int32_t pointstyle,
                   int32_t restricted, const int32_t *h_restricted, float *d_total, int32_t *d_exposed,
                   uint64_t *d_evals, void *stream);

/* cy_find_within(xyz, flgs, others, num, r)  (cy_modules/cy_gridsearch.pyx:24-29 -> find_within,
 * gridsearch.C:595-713) for a batch of frames: h_result[f][i] = 1 iff flgs[i] and some atom j with others[j]
 * has ((dx*dx) + (dy*dy)) + (dz*dz) < r*r in float32 (strict) -- the geometric core of an "around"
 * selection (core/ContactAnalyzer.py:321-324 re-evaluates those per frame).  flgs / others are per-atom
 * 0/1 int32 arrays shared by all frames.  Host pointers; synchronous. */
int pc_find_within(int device, const float *h_xyz, int32_t nframes, int32_t num, const int32_t *h_flgs,
                   const int32_t *h_others, float r, int32_t *h_result);
int pc_within_device(int device, const float *d_xyz, int32_t nframes, int32_t num, int64_t pitch,
                     const int32_t *d_flgs, const int32_t *d_others, float r, int32_t *d_result, void *stream);

/* Synthetic trajectory generator used by bench.py: frame f =
 * base + amp[res]*sin(2*pi*f/period[res] + phase[res]) + N(0, sigma) jitter from a
 * counter-based generator keyed (seed, frame, atom).  Writes nframes frames in `layout`. */
int pc_synth_frames(int device, const float *d_base_xyz, const int32_t *d_residue_of,
                    const float *d_amp, const float *d_period, const float *d_phase,
                    int32_t natoms, float sigma, uint64_t seed, int32_t frame0, int32_t nframes,
                    int32_t layout, int64_t pitch, float *d_out, void *stream);

/* Streams for the frame-batch scheduler (the Pool of core/multi_trajectory.py:357, 434-445 becomes
 * batches in flight on CUDA streams): cudaStreamCreateWithFlags(non-blocking) / synchronize / destroy. */
int pc_stream_create(int device, void **out_stream);
int pc_stream_sync(void *stream);
void pc_stream_destroy(void *stream);

/* Device memory helpers for callers without a CUDA binding of their own (bench.py, tests):
 * cudaMalloc / cudaFree / cudaMemcpyAsync (kind 1 = H2D, 2 = D2H, 3 = D2D) / cudaMemsetAsync /
 * cudaDeviceSynchronize. */
int pc_device_alloc(int device, int64_t bytes, void **out);
void pc_device_free(int device, void *p);
int pc_memcpy(int device, void *dst, const void *src, int64_t bytes, int kind, void *stream);
int pc_memset(int device, void *dst, int value, int64_t bytes, void *stream);
int pc_device_sync(int device);

/* Measurement helpers (bench.py): kernels launched by this library so far, and CUDA events so that
 * the caller can time work on the stream it was enqueued on. */
int64_t pc_launch_count(void);
int pc_event_create(int device, void **out_event);
int pc_event_record(void *event, void *stream);
int pc_event_elapsed_ms(void *start_event, void *end_event, float *ms);   /* synchronises on end_event */
void pc_event_destroy(void *event);

/* Measurement: the large-frame path records these two events (pc_event_create) around its dominant
 * kernel -- the streaming bin of selection 2, or of selection 1 in self mode -- for the first
 * sub-batch of every pc_scan_device call.  NULL, NULL switches it off. */
int pc_set_probe(pc_ctx *ctx, void *start_event, void *end_event);

/* Measurement helpers of bench.py: the GPU's FP32 issue rate from chains of independent FFMAs (instructions per
 * second over all SMs; synchronous, default stream) and the device's free / total memory. */
int pc_fp32_peak(int device, int32_t iters, double *ffma_per_s);
int pc_mem_info(int device, int64_t *free_bytes, int64_t *total_bytes);

/* Pinned host memory helpers for the Python layer (cudaHostAlloc / cudaFreeHost). */
int pc_host_alloc(void **out, int64_t bytes);
void pc_host_free(void *p);

/* ---- XTC trajectories (ingest step in front of the path) --------------------------------------
 * The reference reads trajectories through MDAnalysis (core/ContactAnalyzer.py:282, core/multi_trajectory.py:362; its
 * tests open exampleData/md_noPBC.xtc, tests/test_basic.py:28-30).  Host code: GROMACS XTC frames (xdrfile's compressed
 * integer coordinates) -> float32 positions in Angstrom (nm * 10, like MDAnalysis), optionally gathered through a
 * selection's atom indices -- what sel.positions yields per frame (core/multi_trajectory.py:401-417). */
typedef struct pc_xtc pc_xtc;
int  pc_xtc_open(const char *path, pc_xtc **out);                       /* indexes the frames */
int  pc_xtc_info(const pc_xtc *x, int32_t *natoms, int64_t *nframes);
/* h_xyz: nframes x (nidx or natoms) x 3 floats; indices may be NULL (all atoms); h_step / h_time / h_box (nframes x 9,
 * Angstrom) may be NULL */
int  pc_xtc_read(pc_xtc *x, int64_t frame0, int32_t nframes, const int32_t *indices, int32_t nidx, float *h_xyz,
                 int32_t *h_step, float *h_time, float *h_box);
void pc_xtc_close(pc_xtc *x);

#ifdef __cplusplus
}
#endif
#endif /* PYCONTACT_B200_H */
