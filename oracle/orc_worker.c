/*
 * orc_worker.c — restatement of the per-block worker of the reference
 * (/root/reference/src/dw_subprob.c).  TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 *   orc_block_new      <- dw_subprob.c:116-120 (GLP_LO columns clamped to [lb,3000]),
 *                         :228 (c_k), :230-253 (COO D_k, caller supplies reference order)
 *   orc_block_initial  <- dw_subprob.c:121-122 (cold solve, file objective), :262-266
 *   orc_block_iterate  <- dw_subprob.c:291-293 (y = D_k' pi), :302-304, :317 (d = c - y or -y),
 *                         :329 (warm solve), :336-340 (obj, status), :364-369
 *   prepare_column     <- dw_subprob.c:500-516 (y = D_k x, exact-nonzero compaction), :535
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

struct orc_block {
    int m, n, L, d_nnz;
    orc_lp *lp;
    double *c_file, *c_master;
    int *d_row, *d_col;
    double *d_val;
    double *y;      /* max(n, L) scratch, like the reference's y (dw_subprob.c:158) */
    double *dvec;   /* priced objective */
    double *zero;   /* col_zero_vector */
};

orc_block *orc_block_new(int m, int n, const int *rowptr, const int *colidx, const double *vals,
                         const double *row_lb, const double *row_ub,
                         const double *col_lb, const double *col_ub,
                         const double *c_file, const double *c_master,
                         int L, int d_nnz, const int *d_row, const int *d_col, const double *d_val)
{
    orc_block *b = calloc(1, sizeof(*b));
    if (!b) return NULL;
    b->m = m; b->n = n; b->L = L; b->d_nnz = d_nnz;
    double *ub = malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    for (int j = 0; j < n; ++j) {
        ub[j] = col_ub[j];
        /* GLP_LO column (finite lb, no ub) -> GLP_DB [lb, 3000] */
        if (fabs(col_lb[j]) < ORC_INF && !(fabs(col_ub[j]) < ORC_INF) && col_ub[j] > 0) ub[j] = 3000.0;
    }
    b->lp = orc_lp_new(m, n, rowptr, colidx, vals, row_lb, row_ub, col_lb, ub);
    free(ub);
    size_t nn = (size_t)(n > 0 ? n : 1), nz = (size_t)(d_nnz > 0 ? d_nnz : 1);
    size_t ny = (size_t)((n > L ? n : L) > 0 ? (n > L ? n : L) : 1);
    b->c_file = malloc(sizeof(double) * nn);
    b->c_master = malloc(sizeof(double) * nn);
    b->d_row = malloc(sizeof(int) * nz);
    b->d_col = malloc(sizeof(int) * nz);
    b->d_val = malloc(sizeof(double) * nz);
    b->y = calloc(ny, sizeof(double));
    b->dvec = calloc(nn, sizeof(double));
    b->zero = calloc(nn, sizeof(double));
    memcpy(b->c_file, c_file, sizeof(double) * (size_t)n);
    memcpy(b->c_master, c_master, sizeof(double) * (size_t)n);
    memcpy(b->d_row, d_row, sizeof(int) * (size_t)d_nnz);
    memcpy(b->d_col, d_col, sizeof(int) * (size_t)d_nnz);
    memcpy(b->d_val, d_val, sizeof(double) * (size_t)d_nnz);
    return b;
}

void orc_block_free(orc_block *b)
{
    if (!b) return;
    orc_lp_free(b->lp);
    free(b->c_file); free(b->c_master); free(b->d_row); free(b->d_col); free(b->d_val);
    free(b->y); free(b->dvec); free(b->zero); free(b);
}

void orc_block_counters(const orc_block *b, long *pivots, long *flips, long *refactors)
{
    orc_lp_counters(b->lp, pivots, flips, refactors);
}

/* dw_subprob.c:481-540 */
static void prepare_column(orc_block *b, orc_proposal *out)
{
    orc_dcoogemv(b->L, b->n, b->d_val, b->d_row, b->d_col, b->d_nnz, out->x, b->y);
    int count = 0;
    for (int i = 1; i <= b->L; ++i) {
        if (b->y[i - 1] != 0.0) {
            out->val[count] = b->y[i - 1];
            out->ind[count] = i;
            ++count;
        }
    }
    out->len = count;
    out->new_obj_coeff = orc_ddot(b->n, b->c_master, out->x);
}

int orc_block_initial(orc_block *b, orc_proposal *out)
{
    out->status = orc_lp_solve(b->lp, b->c_file, out->x, &out->obj);
    prepare_column(b, out);
    return 0;
}

int orc_block_iterate(orc_block *b, const double *pi, int phase_one, orc_proposal *out)
{
    /* y = D_k' pi: the transposed product is the same COO loop with the index arrays
     * swapped (dw_subprob.c:291-293) */
    orc_dcoogemv(b->n, b->L, b->d_val, b->d_col, b->d_row, b->d_nnz, pi, b->y);
    orc_dcopy(b->n, phase_one ? b->zero : b->c_master, b->dvec);
    orc_daxpy(b->n, -1.0, b->y, b->dvec);
    out->status = orc_lp_solve(b->lp, b->dvec, out->x, &out->obj);
    prepare_column(b, out);
    return 0;
}

/* ---- pool of pthreads over blocks -------------------------------------------------- */
typedef struct {
    orc_block **blocks; orc_proposal *out; int K; const double *pi; int phase_one, first;
    int next; pthread_mutex_t mu;
} batch_t;

static void *batch_worker(void *arg)
{
    batch_t *bt = arg;
    for (;;) {
        pthread_mutex_lock(&bt->mu);
        int k = bt->next++;
        pthread_mutex_unlock(&bt->mu);
        if (k >= bt->K) break;
        if (bt->first) orc_block_initial(bt->blocks[k], &bt->out[k]);
        else orc_block_iterate(bt->blocks[k], bt->pi, bt->phase_one, &bt->out[k]);
    }
    return NULL;
}

int orc_batch(orc_block **blocks, orc_proposal *out, int K, const double *pi, int phase_one,
              int first, int nthreads)
{
    batch_t bt = { blocks, out, K, pi, phase_one, first, 0, PTHREAD_MUTEX_INITIALIZER };
    if (nthreads < 1) nthreads = 1;
    if (nthreads > K) nthreads = K;
    if (nthreads == 1) { batch_worker(&bt); return 0; }
    pthread_t *th = malloc(sizeof(pthread_t) * (size_t)nthreads);
    for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, batch_worker, &bt);
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    free(th);
    return 0;
}
