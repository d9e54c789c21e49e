/*
 * oracle/gridsearch_oracle.c  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the reference's native neighbour search
 *     find_contacts()      PyContact/cy_modules/src/gridsearch.C:408-427
 *     vmd_gridsearch3()    PyContact/cy_modules/src/gridsearch.C:902-1162
 *     find_minmax()        PyContact/cy_modules/src/gridsearch.C:63-116
 *     make_neighborlist_sym()  PyContact/cy_modules/src/gridsearch.C:845-899
 *     distance2()          PyContact/cy_modules/src/gridsearch.C:48-55
 *
 * It is written from the algorithm's description, not transliterated: bins are
 * a counting sort (the reference grows malloc'd arrays by 4), the 27-entry
 * stencil is a constant offset table evaluated per box (the reference mallocs
 * one list per box), and output is a CSR (row_ptr, cols) instead of a linked
 * list converted to vector<vector<int>>.  What IS preserved, because parity
 * depends on it:
 *   - float32 arithmetic with the reference's association order and no FMA
 *     (build with -ffp-contract=off), cutoff narrowed double->float, then
 *     squared in float32;
 *   - the joint bounding box over all atoms of both sets, the box count rule
 *     (int)(range * (1.0f/pairdist)) + 1, the "grow the edge x1.26 while more
 *     than 4e6 boxes" loop, and clamping of box coordinates;
 *   - the visiting order box-ascending -> stencil-table order -> i in A(box)
 *     -> j in B(nbr), which fixes the order of cols within each CSR row
 *     (SURVEY.md App. A.3).
 *
 * Pinning: checked against the compiled reference (oracle/_ref) on random
 * clouds and on the rpn11_ubq fixture by tests/test_oracle.py, and against
 * the golden vectors in tests/golden/ produced by the unmodified reference.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this file's library.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* stencil offsets in the order of make_neighborlist_sym (gridsearch.C:858-885) */
static const int8_t STENCIL[27][3] = {
    { 0, 0, 0}, { 1, 0, 0}, { 0, 1, 0}, { 0, 0, 1}, { 1, 1, 0}, { 1, 0, 1},
    { 0, 1, 1}, { 1,-1, 0}, {-1, 0, 1}, { 0,-1, 1}, { 1, 1, 1}, {-1, 1, 1},
    { 1,-1, 1}, {-1,-1, 1}, {-1, 0, 0}, { 0,-1, 0}, { 0, 0,-1}, {-1,-1, 0},
    {-1, 0,-1}, { 0,-1,-1}, {-1, 1, 0}, { 1, 0,-1}, { 0, 1,-1}, {-1,-1,-1},
    { 1,-1,-1}, {-1, 1,-1}, { 1, 1,-1}};

typedef struct {
    float min[3];
    float inv;      /* 1.0f / (possibly grown) cell edge */
    int nb[3];      /* boxes per axis */
} oracle_grid;

static void minmax3(const float *p, int n, float *lo, float *hi) {
    /* gridsearch.C:63-116 -- note the reference's "else if": max is only
     * tested when the value is not a new min.  Same result for n >= 1 except
     * that the first atom seeds both. */
    lo[0] = hi[0] = p[0]; lo[1] = hi[1] = p[1]; lo[2] = hi[2] = p[2];
    for (int i = 1; i < n; i++) {
        const float *q = p + 3L * i;
        for (int k = 0; k < 3; k++) {
            if (q[k] < lo[k]) lo[k] = q[k];
            else if (q[k] > hi[k]) hi[k] = q[k];
        }
    }
}

/* grid geometry exactly as gridsearch.C:940-985 */
void oracle_grid_geometry(const float *posA, int nA, const float *posB, int nB,
                          double cutoff, float *out_min, float *out_inv, int *out_nb) {
    float pairdist = (float)cutoff;
    float lo[3], hi[3], lo2[3], hi2[3];
    minmax3(posA, nA, lo, hi);
    minmax3(posB, nB, lo2, hi2);
    for (int k = 0; k < 3; k++) {
        if (lo2[k] < lo[k]) lo[k] = lo2[k];
        if (hi2[k] > hi[k]) hi[k] = hi2[k];
    }
    const int MAXBOXES = 4000000;
    float range[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
    float next = pairdist;
    int nb[3], tot;
    do {
        pairdist = next;
        const float inv = 1.0f / pairdist;
        nb[0] = ((int)(range[0] * inv)) + 1;
        nb[1] = ((int)(range[1] * inv)) + 1;
        nb[2] = ((int)(range[2] * inv)) + 1;
        tot = nb[0] * nb[1] * nb[2];
        next = pairdist * 1.26f;
    } while (tot > MAXBOXES || tot < 1);
    out_min[0] = lo[0]; out_min[1] = lo[1]; out_min[2] = lo[2];
    *out_inv = 1.0f / pairdist;
    out_nb[0] = nb[0]; out_nb[1] = nb[1]; out_nb[2] = nb[2];
}

static inline void box_of(const float *p, const oracle_grid *g, int *c) {
    for (int k = 0; k < 3; k++) {
        int v = (int)((p[k] - g->min[k]) * g->inv);
        if (v >= g->nb[k]) v = g->nb[k] - 1;
        c[k] = v;
    }
}

static inline float dist2_f32(const float *a, const float *b) {
    /* gridsearch.C:48-55: ((dx*dx + dy*dy) + dz*dz), every op rounded to f32 */
    float d = a[0] - b[0];
    float r2 = d * d;
    d = a[1] - b[1];
    r2 += d * d;
    d = a[2] - b[2];
    return r2 + d * d;
}

/* counting sort of atoms into boxes; stable, so atoms stay index-ascending
 * inside a box like the reference's append-in-order bins */
static int bin_atoms(const float *pos, int n, const oracle_grid *g, int totb,
                     int **start_out, int **items_out) {
    int *start = (int *)calloc((size_t)totb + 1, sizeof(int));
    int *items = (int *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int));
    int *box = (int *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int));
    if (!start || !items || !box) { free(start); free(items); free(box); return -1; }
    for (int i = 0; i < n; i++) {
        int c[3];
        box_of(pos + 3L * i, g, c);
        box[i] = (c[2] * g->nb[1] + c[1]) * g->nb[0] + c[0];
        start[box[i] + 1]++;
    }
    for (int b = 0; b < totb; b++) start[b + 1] += start[b];
    int *cur = (int *)malloc((size_t)totb * sizeof(int));
    if (!cur) { free(start); free(items); free(box); return -1; }
    memcpy(cur, start, (size_t)totb * sizeof(int));
    for (int i = 0; i < n; i++) items[cur[box[i]]++] = i;
    free(cur); free(box);
    *start_out = start; *items_out = items;
    return 0;
}

/*
 * oracle_find_contacts: CSR of set-B neighbours for every atom of set A.
 *   row_ptr : int64[nA+1], caller allocated
 *   cols    : *cols_out is malloc'd here (release with oracle_free), length row_ptr[nA]
 * returns 0, or -1 on allocation failure.  nA==0 or nB==0 -> all rows empty
 * (gridsearch.C:906).
 */
int oracle_find_contacts(const float *posA, int nA, const float *posB, int nB,
                         double cutoff, int64_t *row_ptr, int32_t **cols_out) {
    *cols_out = NULL;
    for (int i = 0; i <= nA; i++) row_ptr[i] = 0;
    if (nA <= 0 || nB <= 0) return 0;

    const float pairdist = (float)cutoff;
    const float sqdist = pairdist * pairdist;   /* gridsearch.C:936 */
    oracle_grid g;
    oracle_grid_geometry(posA, nA, posB, nB, cutoff, g.min, &g.inv, g.nb);
    const int totb = g.nb[0] * g.nb[1] * g.nb[2];

    int *startA, *itemsA, *startB, *itemsB;
    if (bin_atoms(posA, nA, &g, totb, &startA, &itemsA)) return -1;
    if (bin_atoms(posB, nB, &g, totb, &startB, &itemsB)) { free(startA); free(itemsA); return -1; }

    int32_t *cols = NULL;
    int64_t *fill = (int64_t *)malloc(((size_t)nA + 1) * sizeof(int64_t));
    if (!fill) return -1;

    for (int pass = 0; pass < 2; pass++) {
        if (pass == 1) {
            int64_t acc = 0;
            for (int i = 0; i < nA; i++) { int64_t c = row_ptr[i + 1]; row_ptr[i + 1] = acc + c; fill[i] = acc; acc += c; }
            row_ptr[0] = 0;
            cols = (int32_t *)malloc((size_t)(acc > 0 ? acc : 1) * sizeof(int32_t));
            if (!cols) { free(fill); return -1; }
        }
        for (int bz = 0; bz < g.nb[2]; bz++)
        for (int by = 0; by < g.nb[1]; by++)
        for (int bx = 0; bx < g.nb[0]; bx++) {
            const int b = (bz * g.nb[1] + by) * g.nb[0] + bx;
            if (startA[b] == startA[b + 1]) continue;
            for (int s = 0; s < 27; s++) {
                const int x = bx + STENCIL[s][0], y = by + STENCIL[s][1], z = bz + STENCIL[s][2];
                if (x < 0 || y < 0 || z < 0 || x >= g.nb[0] || y >= g.nb[1] || z >= g.nb[2]) continue;
                const int nb = (z * g.nb[1] + y) * g.nb[0] + x;
                for (int ia = startA[b]; ia < startA[b + 1]; ia++) {
                    const int i = itemsA[ia];
                    const float *pa = posA + 3L * i;
                    for (int jb = startB[nb]; jb < startB[nb + 1]; jb++) {
                        const int j = itemsB[jb];
                        if (dist2_f32(pa, posB + 3L * j) > sqdist) continue;   /* gridsearch.C:1126 */
                        if (pass == 0) row_ptr[i + 1]++;
                        else cols[fill[i]++] = j;
                    }
                }
            }
        }
    }
    free(fill); free(startA); free(itemsA); free(startB); free(itemsB);
    *cols_out = cols;
    return 0;
}

void oracle_free(void *p) { free(p); }

/* stencil rank of a box offset (dx,dy,dz in -1..1), -1 if not adjacent */
int oracle_stencil_rank(int dx, int dy, int dz) {
    for (int s = 0; s < 27; s++)
        if (STENCIL[s][0] == dx && STENCIL[s][1] == dy && STENCIL[s][2] == dz) return s;
    return -1;
}

/* box coordinates of n atoms under a given grid (for canonical-order checks) */
void oracle_box_coords(const float *pos, int n, const float *gmin, float inv,
                       const int *nb, int32_t *out_xyz) {
    oracle_grid g;
    g.min[0] = gmin[0]; g.min[1] = gmin[1]; g.min[2] = gmin[2];
    g.inv = inv; g.nb[0] = nb[0]; g.nb[1] = nb[1]; g.nb[2] = nb[2];
    for (int i = 0; i < n; i++) {
        int c[3];
        box_of(pos + 3L * i, &g, c);
        out_xyz[3L * i] = c[0]; out_xyz[3L * i + 1] = c[1]; out_xyz[3L * i + 2] = c[2];
    }
}

/* brute-force float32 search with the same arithmetic: an independent check of
 * the set (not the order).  counts only unless cols != NULL (then fills rows in
 * j-ascending order using row_ptr as produced by a counting call). */
int64_t oracle_brute_count(const float *posA, int nA, const float *posB, int nB,
                           double cutoff, int64_t *row_counts) {
    const float pairdist = (float)cutoff;
    const float sqdist = pairdist * pairdist;
    int64_t tot = 0;
    for (int i = 0; i < nA; i++) {
        int64_t c = 0;
        for (int j = 0; j < nB; j++)
            if (!(dist2_f32(posA + 3L * i, posB + 3L * j) > sqdist)) c++;
        if (row_counts) row_counts[i] = c;
        tot += c;
    }
    return tot;
}
