#ifndef PF_SHIM_GSL_LINALG_H
#define PF_SHIM_GSL_LINALG_H
#include <math.h>
#include "gsl_matrix.h"
inline int gsl_linalg_LU_decomp(gsl_matrix* A, gsl_permutation* p, int* signum) {
    const size_t N = A->size1;
    *signum = 1;
    for (size_t i = 0; i != N; ++i) p->data[i] = i;
    for (size_t j = 0; j + 1 < N; ++j) {
        double max = fabs(gsl_matrix_get(A, j, j));
        size_t i_pivot = j;
        for (size_t i = j + 1; i < N; ++i) {
            double aij = fabs(gsl_matrix_get(A, i, j));
            if (aij > max) { max = aij; i_pivot = i; }
        }
        if (i_pivot != j) {
            for (size_t k = 0; k != N; ++k) {
                double t = gsl_matrix_get(A, j, k);
                gsl_matrix_set(A, j, k, gsl_matrix_get(A, i_pivot, k));
                gsl_matrix_set(A, i_pivot, k, t);
            }
            size_t t = p->data[j]; p->data[j] = p->data[i_pivot]; p->data[i_pivot] = t;
            *signum = -*signum;
        }
        double ajj = gsl_matrix_get(A, j, j);
        if (ajj != 0.0) {
            for (size_t i = j + 1; i < N; ++i) {
                double aij = gsl_matrix_get(A, i, j) / ajj;
                gsl_matrix_set(A, i, j, aij);
                for (size_t k = j + 1; k < N; ++k)
                    gsl_matrix_set(A, i, k, gsl_matrix_get(A, i, k) - aij * gsl_matrix_get(A, j, k));
            }
        }
    }
    return 0;
}
inline int gsl_linalg_LU_invert(const gsl_matrix* LU, const gsl_permutation* p, gsl_matrix* inverse) {
    const size_t N = LU->size1;
    for (size_t c = 0; c != N; ++c) {
        double x[64];
        /* permuted unit vector, then L (unit lower) forward, then U backward */
        for (size_t i = 0; i != N; ++i) x[i] = (p->data[i] == c) ? 1.0 : 0.0;
        for (size_t i = 0; i != N; ++i)
            for (size_t k = 0; k < i; ++k) x[i] -= gsl_matrix_get(LU, i, k) * x[k];
        for (size_t ii = N; ii-- > 0;) {
            for (size_t k = ii + 1; k < N; ++k) x[ii] -= gsl_matrix_get(LU, ii, k) * x[k];
            x[ii] /= gsl_matrix_get(LU, ii, ii);
        }
        for (size_t i = 0; i != N; ++i) gsl_matrix_set(inverse, i, c, x[i]);
    }
    return 0;
}
inline double gsl_linalg_LU_det(gsl_matrix* LU, int signum) {
    double det = (double) signum;
    for (size_t i = 0; i != LU->size1; ++i) det *= gsl_matrix_get(LU, i, i);
    return det;
}
#endif
