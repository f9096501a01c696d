#ifndef PF_SHIM_GSL_BLAS_H
#define PF_SHIM_GSL_BLAS_H
#include "gsl_matrix.h"
enum CBLAS_UPLO { CblasUpper = 121, CblasLower = 122 };
/* y := alpha*A*x + beta*y, A symmetric, only the Uplo triangle referenced
 * (reference CBLAS row-major loop order). */
inline int gsl_blas_dsymv(CBLAS_UPLO uplo, double alpha, const gsl_matrix* A, const gsl_vector* X,
                          double beta, gsl_vector* Y) {
    const size_t N = A->size1;
    for (size_t i = 0; i != N; ++i) Y->data[i] = (beta == 0.0) ? 0.0 : beta * Y->data[i];
    if (uplo == CblasUpper) {
        for (size_t i = 0; i != N; ++i) {
            double temp1 = alpha * X->data[i], temp2 = 0.0;
            Y->data[i] += temp1 * gsl_matrix_get(A, i, i);
            for (size_t j = i + 1; j < N; ++j) {
                Y->data[j] += temp1 * gsl_matrix_get(A, i, j);
                temp2 += X->data[j] * gsl_matrix_get(A, i, j);
            }
            Y->data[i] += alpha * temp2;
        }
    } else {
        for (size_t ii = N; ii-- > 0;) {
            double temp1 = alpha * X->data[ii], temp2 = 0.0;
            Y->data[ii] += temp1 * gsl_matrix_get(A, ii, ii);
            for (size_t j = 0; j < ii; ++j) {
                Y->data[j] += temp1 * gsl_matrix_get(A, ii, j);
                temp2 += X->data[j] * gsl_matrix_get(A, ii, j);
            }
            Y->data[ii] += alpha * temp2;
        }
    }
    return 0;
}
inline int gsl_blas_ddot(const gsl_vector* X, const gsl_vector* Y, double* result) {
    double r = 0.0;
    for (size_t i = 0; i != X->size; ++i) r += X->data[i] * Y->data[i];
    *result = r;
    return 0;
}
#endif
