/* oracle/shim/gsl -- TEST INFRASTRUCTURE ONLY.
 * A small restatement of the dozen GSL entry points normalsampler.hh:19-21,47-51,59-63,
 * 92-112,138-150 calls, so that the reference's general-covariance branch runs inside
 * the oracle build.  GSL itself is an unvendored system dependency of the reference
 * (GNUmakefile:10).  The arithmetic order follows GSL 1.x: Gaussian elimination with
 * partial pivoting (linalg/lu.c), column-by-column inversion, and the reference-CBLAS
 * row-major/upper dsymv loop (cblas/source_symv.h). */
#ifndef PF_SHIM_GSL_MATRIX_H
#define PF_SHIM_GSL_MATRIX_H
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
struct gsl_vector { size_t size; double* data; };
struct gsl_matrix { size_t size1, size2; double* data; };
struct gsl_permutation { size_t size; size_t* data; };
struct gsl_vector_const_view { gsl_vector vector; };
inline gsl_vector* gsl_vector_alloc(size_t n) {
    gsl_vector* v = (gsl_vector*) malloc(sizeof(gsl_vector));
    v->size = n; v->data = (double*) calloc(n, sizeof(double)); return v;
}
inline gsl_matrix* gsl_matrix_alloc(size_t r, size_t c) {
    gsl_matrix* m = (gsl_matrix*) malloc(sizeof(gsl_matrix));
    m->size1 = r; m->size2 = c; m->data = (double*) calloc(r * c, sizeof(double)); return m;
}
inline void gsl_matrix_free(gsl_matrix* m) { free(m->data); free(m); }
inline void gsl_vector_free(gsl_vector* v) { free(v->data); free(v); }
inline void gsl_vector_set(gsl_vector* v, size_t i, double x) { v->data[i] = x; }
inline double gsl_vector_get(const gsl_vector* v, size_t i) { return v->data[i]; }
inline void gsl_matrix_set(gsl_matrix* m, size_t i, size_t j, double x) { m->data[i * m->size2 + j] = x; }
inline double gsl_matrix_get(const gsl_matrix* m, size_t i, size_t j) { return m->data[i * m->size2 + j]; }
inline gsl_permutation* gsl_permutation_alloc(size_t n) {
    gsl_permutation* p = (gsl_permutation*) malloc(sizeof(gsl_permutation));
    p->size = n; p->data = (size_t*) malloc(n * sizeof(size_t));
    for (size_t i = 0; i != n; ++i) p->data[i] = i;
    return p;
}
inline void gsl_permutation_free(gsl_permutation* p) { free(p->data); free(p); }
inline gsl_vector_const_view gsl_vector_const_view_array(const double* base, size_t n) {
    gsl_vector_const_view v; v.vector.size = n; v.vector.data = const_cast<double*>(base); return v;
}
inline int gsl_vector_memcpy(gsl_vector* dst, const gsl_vector* src) {
    memcpy(dst->data, src->data, sizeof(double) * src->size); return 0;
}
inline int gsl_vector_sub(gsl_vector* a, const gsl_vector* b) {
    for (size_t i = 0; i != a->size; ++i) a->data[i] -= b->data[i];
    return 0;
}
#endif
