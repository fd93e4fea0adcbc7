/*
 * orc_blas.c — restatement of the five dense/sparse helper loops of the reference
 * (/root/reference/src/dw_blas.c:52-122).  TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Each routine keeps the reference's summation order and operation order (multiply,
 * then add into the accumulator, left to right) so results agree bit-for-bit with the
 * reference object when both are compiled without FP contraction.
 */
#include "oracle.h"

/* dw_blas.c:52-59 — y[i] = alpha*x[i] + y[i] */
void orc_daxpy(int len, double alpha, const double *x, double *y)
{
    for (int t = 0; t < len; ++t) {
        double prod = alpha * x[t];
        y[t] = prod + y[t];
    }
}

/* dw_blas.c:65-75 — sequential dot, index order */
double orc_ddot(int len, const double *x, const double *y)
{
    double acc = 0.0;
    const double *px = x, *py = y, *end = x + (len > 0 ? len : 0);
    while (px != end) {
        acc += (*px++) * (*py++);
    }
    return acc;
}

/* dw_blas.c:82-93 — dot of a packed vector with a gathered dense vector */
double orc_ddoti(int nz, const double *x, const int *ind, const double *y)
{
    double acc = 0.0;
    for (int t = 0; t < nz; ++t)
        acc += x[t] * y[ind[t]];
    return acc;
}

/* dw_blas.c:96-103 */
void orc_dcopy(int len, const double *x, double *y)
{
    for (int t = 0; t < len; ++t)
        y[t] = x[t];
}

/* dw_blas.c:106-122 — c = A*b with A in coordinate form; c (m entries) is cleared
 * first and the triplets are applied in storage order. `k` (column count) is unused
 * by the reference as well. */
void orc_dcoogemv(int m, int k, const double *val, const int *indx, const int *jndx,
                  int nnz, const double *b, double *c)
{
    (void)k;
    for (int r = 0; r < m; ++r)
        c[r] = 0;
    for (int e = 0; e < nnz; ++e) {
        double prod = b[jndx[e]] * val[e];
        c[indx[e]] += prod;
    }
}
