/* CPU restatement (plain C) of the reference monotonic alignment search.
 *
 * TEST INFRASTRUCTURE ONLY -- never linked into the product library.
 * Restates seq2seq_vc/modules/alignments.py:63-93 (_monotonic_alignment_search) and the
 * per-utterance slicing / bincount of viterbi_decode (:301-305).  See oracle/mas_oracle.py for
 * the arithmetic contract.  Build: gcc -O2 -fno-fast-math -ffp-contract=off -shared -fPIC.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* one utterance: lp is (t_mel x t_inp) with row stride ld (floats). */
static int mas_one(const float *lp, long ld, int t_mel, int t_inp, int64_t *path, float *ds, int ds_len) {
    if (t_mel <= 0 || t_inp <= 0) return 0;
    double *Q = (double *)malloc(sizeof(double) * (size_t)t_mel * (size_t)t_inp);
    if (!Q) return -1;
    for (size_t n = 0; n < (size_t)t_mel * (size_t)t_inp; ++n) Q[n] = -INFINITY;
    /* alignments.py:72-73: row 0 = running sum in the array dtype (float32), widened on store */
    volatile float acc = 0.0f;
    for (int j = 0; j < t_mel; ++j) {
        acc = acc + lp[(long)j * ld];
        Q[j] = (double)acc; /* Q[0][j] */
    }
    /* alignments.py:76-78 */
    for (int j = 1; j < t_mel; ++j) {
        int imax = (j + 1 < t_inp) ? j + 1 : t_inp;
        for (int i = 1; i < imax; ++i) {
            double a = Q[(size_t)(i - 1) * t_mel + (j - 1)];
            double b = Q[(size_t)i * t_mel + (j - 1)];
            double m = (a > b) ? a : b;
            Q[(size_t)i * t_mel + j] = m + (double)lp[(long)j * ld + i];
        }
    }
    /* alignments.py:81-92 */
    path[t_mel - 1] = t_inp - 1;
    for (int j = t_mel - 2; j >= 0; --j) {
        int64_t ib = path[j + 1];
        int64_t r;
        if (ib == 0) r = 0;
        else if (Q[(size_t)(ib - 1) * t_mel + j] >= Q[(size_t)ib * t_mel + j]) r = ib - 1;
        else r = ib;
        path[j] = r;
    }
    for (int j = 0; j < t_mel; ++j)
        if (path[j] >= 0 && path[j] < ds_len) ds[path[j]] += 1.0f;
    free(Q);
    return 0;
}

int mas_oracle_batch(const float *log_p, int B, int t_feats, int t_text, const int64_t *text_lens,
                     const int64_t *feats_lens, int64_t *paths, float *ds, int threads) {
    int rc = 0;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel for schedule(dynamic) reduction(| : rc)
#endif
    for (int b = 0; b < B; ++b) {
        int fl = (int)feats_lens[b], tl = (int)text_lens[b];
        if (fl > t_feats || tl > t_text) { rc |= 1; continue; }
        rc |= mas_one(log_p + (size_t)b * t_feats * t_text, t_text, fl, tl,
                      paths + (size_t)b * t_feats, ds + (size_t)b * t_text, t_text) ? 2 : 0;
    }
    return rc;
}
