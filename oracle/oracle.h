/*
 * oracle.h — CPU restatement of dwsolver's subproblem hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (dwsolver_b200/, include/) may
 * include, link or call this; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, as the checker or the CPU baseline.
 *
 * What it restates (all paths relative to /root/reference):
 *   orc_daxpy/ddot/ddoti/dcopy/dcoogemv  <- src/dw_blas.c:52-122 (same loops, same
 *        summation order; validated bit-for-bit against the reference object built
 *        into oracle/_ref/libdwblas_ref.so and against tests/test_blas.c:47-184)
 *   orc_block_initial                    <- src/dw_subprob.c:116-122, 262-266
 *   orc_block_iterate                    <- src/dw_subprob.c:291-321, 329, 336-340, 364-369
 *   orc_prepare_column (internal)        <- src/dw_subprob.c:481-540
 *   orc_lp_solve                         <- stands in for glp_simplex (system GLPK >= 4.65,
 *        configure.ac:52; NOT in the reference tree).  It is a plain dense
 *        bounded-variable primal simplex returning the exact LP optimum; pinned
 *        against scipy/HiGHS and the reference's published example optima
 *        (tests/dw-tests.sh), see tests/test_oracle_*.py.
 */
#ifndef DW_ORACLE_H_
#define DW_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

/* GLPK solution-status values (glpk.h): the reference stores glp_get_status() in
 * subprob_struct.unbounded (dw_subprob.c:340) and tests it against GLP_UNBND. */
#define ORC_UNDEF  1
#define ORC_FEAS   2
#define ORC_INFEAS 3
#define ORC_NOFEAS 4
#define ORC_OPT    5
#define ORC_UNBND  6

#define ORC_INF 1e30 /* |bound| >= ORC_INF means "no bound" */

/* ---- dw_blas restatement ---------------------------------------------------------- */
void   orc_daxpy(int len, double alpha, const double *x, double *y);
double orc_ddot(int len, const double *x, const double *y);
double orc_ddoti(int nz, const double *x, const int *ind, const double *y);
void   orc_dcopy(int len, const double *x, double *y);
void   orc_dcoogemv(int m, int k, const double *val, const int *indx, const int *jndx,
                    int nnz, const double *b, double *c);

/* ---- LP engine standing in for glp_simplex ---------------------------------------- */
typedef struct orc_lp orc_lp;

/* A in CSR (0-based).  Bounds use +-ORC_INF (or IEEE inf) for "none". */
orc_lp *orc_lp_new(int m, int n, const int *rowptr, const int *colidx, const double *vals,
                   const double *row_lb, const double *row_ub,
                   const double *col_lb, const double *col_ub);
void    orc_lp_free(orc_lp *lp);
/* Primal simplex from the stored basis (all-logical on the first call).  Returns a
 * status above; x (n) and *objval are written for OPT. */
int     orc_lp_solve(orc_lp *lp, const double *cost, double *x, double *objval);
void    orc_lp_counters(const orc_lp *lp, long *pivots, long *flips, long *refactors);
/* max scaled primal residual of x against rows and column bounds (north_star: <= 1e-9) */
double  orc_lp_residual(const orc_lp *lp, const double *x);

/* ---- block worker (the loop body of subproblem_thread) ----------------------------- */
typedef struct orc_block orc_block;

orc_block *orc_block_new(int m, int n, const int *rowptr, const int *colidx, const double *vals,
                         const double *row_lb, const double *row_ub,
                         const double *col_lb, const double *col_ub,
                         const double *c_file, const double *c_master,
                         int L, int d_nnz, const int *d_row, const int *d_col, const double *d_val);
void orc_block_free(orc_block *b);

/* proposal record written by initial/iterate (the live mailbox fields of
 * subprob_struct, dw_subprob.h:44-94) */
typedef struct {
    int     status;        /* glp_get_status analogue                     (unbounded) */
    double  obj;           /* optimum of the priced LP                    (obj)       */
    double  new_obj_coeff; /* c_k . x_k                                   (new_obj_coeff) */
    int     len;           /* nonzeros of D_k x_k                         (len)       */
    int    *ind;           /* 1-based master rows, capacity L             (ind[1..])  */
    double *val;           /*                                             (val[1..])  */
    double *x;             /* n values, 0-based                           (current_solution[1..]) */
} orc_proposal;

int orc_block_initial(orc_block *b, orc_proposal *out);
int orc_block_iterate(orc_block *b, const double *pi /* L */, int phase_one, orc_proposal *out);
void orc_block_counters(const orc_block *b, long *pivots, long *flips, long *refactors);

/* Solve a batch of blocks with a pool of nthreads pthreads (the reference runs one
 * pthread per block, dw_solver.c:154-184; a pool bounds the thread count at K=8192).
 * first = 1 runs orc_block_initial, else orc_block_iterate.  Returns 0. */
int orc_batch(orc_block **blocks, orc_proposal *out, int K, const double *pi, int phase_one,
              int first, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
