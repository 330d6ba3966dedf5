/* oracle/miou_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * Plain-C restatement of the reference's only native module, src/helpers/miou_utils.pyx
 * (Cython, CPU).  Pinned by tests/test_oracle_golden.py against vectors produced by the real
 * (alias-patched) Cython module, see tests/golden/make_golden.py.
 */
#include <stdint.h>
#include <string.h>

/* miou_utils.pyx:7-30 -- cm[gt[i], preds[i]] += 1, rows = ground truth, int64 result. */
void oracle_fast_cm(const uint8_t *preds, const uint8_t *gt, int64_t n, int n_classes, int64_t *cm) {
    memset(cm, 0, sizeof(int64_t) * (size_t)n_classes * (size_t)n_classes);
    for (int64_t i = 0; i < n; ++i) cm[(int64_t)gt[i] * n_classes + preds[i]] += 1;
}

/* miou_utils.pyx:59-90 (and compute_iu :32-57 = the iu output alone).
 * Accumulators are C `unsigned int` in the reference, i.e. they wrap mod 2^32; default value 2
 * marks absent classes; true division to float64. */
void oracle_compute_ius_accs(const int64_t *cm, int n_classes, double *iu, int64_t *n_pixels, double *accs) {
    for (int i = 0; i < n_classes; ++i) {
        uint32_t pi = 0, gi = 0, ii, denom;
        for (int j = 0; j < n_classes; ++j) {
            pi += (uint32_t)cm[(int64_t)j * n_classes + i];
            gi += (uint32_t)cm[(int64_t)i * n_classes + j];
        }
        ii = (uint32_t)cm[(int64_t)i * n_classes + i];
        denom = pi + gi - ii;
        iu[i] = 2.0;
        accs[i] = 2.0;
        if (denom > 0) iu[i] = (double)ii / (double)denom;
        if (gi > 0) accs[i] = (double)ii / (double)gi;
        n_pixels[i] = (int64_t)gi;
    }
}
