// oracle/ref_shim.cpp  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Thin C-ABI wrapper that compiles the UNMODIFIED reference search
// (PyContact/cy_modules/src/gridsearch.C) from where it lies under
// /root/reference and exposes find_contacts() (gridsearch.C:408-427) as a
// CSR-returning extern "C" function.  Built only by oracle/Makefile into
// oracle/_ref/ (git-ignored); no reference source is copied into the repo.
//
// The reference keeps two int flag arrays of nAtoms1 / nAtoms2 entries on the
// stack (gridsearch.C:413-416); callers with N1+N2 > ~2e6 must raise
// RLIMIT_STACK first (oracle/ref_native.py does).
#ifndef PC_REFERENCE_GRIDSEARCH
#error "define PC_REFERENCE_GRIDSEARCH to the path of the reference gridsearch.C"
#endif
#include PC_REFERENCE_GRIDSEARCH

#include <cstdint>
#include <cstdlib>

extern "C" int ref_find_contacts_csr(const float *pos1, int n1, const float *pos2, int n2,
                                     double cutoff, int64_t *row_ptr, int32_t **cols_out) {
    std::vector<std::vector<int>> rows = find_contacts(pos1, pos2, n1, n2, cutoff);
    int64_t total = 0;
    row_ptr[0] = 0;
    for (int i = 0; i < n1; i++) {
        total += (int64_t)rows[i].size();
        row_ptr[i + 1] = total;
    }
    int32_t *cols = (int32_t *)malloc((size_t)(total > 0 ? total : 1) * sizeof(int32_t));
    if (!cols) return -1;
    int64_t k = 0;
    for (int i = 0; i < n1; i++)
        for (size_t j = 0; j < rows[i].size(); j++) cols[k++] = rows[i][j];
    *cols_out = cols;
    // rows n1 .. n1+n2-1 are always empty in the reference (gridsearch.C:410,421-422)
    int64_t tail = 0;
    for (int i = n1; i < n1 + n2; i++) tail += (int64_t)rows[i].size();
    return tail == 0 ? 0 : -2;
}

extern "C" void ref_free(void *p) { free(p); }

// sasa_grid (gridsearch.C:430-529) and find_within (gridsearch.C:595-713), reached the way
// cy_gridsearch.pyx:14-31 reaches them (allow_double_counting = 0, maxpairs = -1).
extern "C" double ref_sasa_grid(const float *pos, int natoms, float pairdist, const float *radius, int npts,
                                double srad, int pointstyle, int restricted, const int *restrictedList) {
    // cy_sasa stores the double in a C float before returning it (cy_gridsearch.pyx:21-22)
    const float sasa = (float)sasa_grid(pos, natoms, pairdist, 0, -1, radius, npts, srad, pointstyle, restricted,
                                        restrictedList);
    return (double)sasa;
}

extern "C" void ref_find_within(const float *xyz, int *flgs, int *others, int num, float r, int *result_out) {
    int *res = find_within(xyz, flgs, others, num, r);
    for (int i = 0; i < num; i++) result_out[i] = res[i];
    delete[] res;
}
