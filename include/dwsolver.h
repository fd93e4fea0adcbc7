/*
 * dwsolver.h — public API of the drop-in host library (libdwsolver_b200.so).
 *
 * Same symbols, struct layouts, defaults and status codes as the reference's public header
 * (/root/reference/src/dw_solver.h:88-183): a program written against libdwsolver links against
 * this library unchanged.  Behind dw_solve() the K subproblems are not K pthreads calling
 * glp_simplex (src/dw_solver.c:154-184, src/dw_subprob.c:63-409) but ONE batched engine on the GPU
 * (include/dwgpu.h); the restricted master stays on the host as in the reference
 * (src/dw_phases.c, src/dw_solver.c:259-694).  GLPK is not available in this build environment, so the
 * master LP is solved by the library's own bounded primal simplex (dwsolver_b200/host/master_lp.cc).
 */
#ifndef DWSOLVER_B200_H_
#define DWSOLVER_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define DW_OUTPUT_SILENT   (-1)
#define DW_OUTPUT_QUIET      0
#define DW_OUTPUT_NORMAL     5
#define DW_OUTPUT_VERBOSE   10
#define DW_OUTPUT_ALL       15

typedef enum {                       /* src/dw_solver.h:88-97 */
    DW_STATUS_OK               = 0,
    DW_STATUS_ERR_BAD_ARGS     = 1,
    DW_STATUS_ERR_FILE         = 2,
    DW_STATUS_ERR_INFEASIBLE   = 3,
    DW_STATUS_ERR_PHASE1_LIMIT = 4,
    DW_STATUS_ERR_PHASE2_LIMIT = 5,
    DW_STATUS_ERR_UNBOUNDED    = 6,
    DW_STATUS_ERR_INTERNAL     = 99
} dw_status_t;

typedef struct {                     /* src/dw_solver.h:105-118 — same 12 fields, same order */
    int    verbosity;               /* default 5 */
    double mip_gap;                 /* default 0.01 (unused: no MIP in this build) */
    int    max_phase1_iterations;   /* default 100 */
    int    max_phase2_iterations;   /* default 3000 */
    int    rounding_flag;           /* default 0 (not supported: ignored with a message) */
    int    integerize_flag;         /* default 0 (not supported) */
    int    enforce_sub_integrality; /* default 0 (not supported: SURVEY a7) */
    int    print_timing_data;       /* default 0 */
    int    print_final_master;      /* default 0 */
    int    print_relaxed_sol;       /* default 1 */
    int    perturb;                 /* default 0 */
    double shift;                   /* default 0.0 */
} dw_options_t;

typedef struct {                     /* src/dw_solver.h:126-131 */
    int     status;
    double  objective_value;
    int     num_vars;               /* number of final master columns (slacks + lambdas) */
    double *x;                      /* their primal values; release with dw_result_free() */
} dw_result_t;

void        dw_options_init(dw_options_t *opts);
int         dw_solve(const char *master_file, const char *const *subproblem_files, int num_subproblems,
                     const dw_options_t *opts, dw_result_t *result);
void        dw_result_free(dw_result_t *result);
const char *dw_version(void);

/* ---- additions of this build (not in the reference) ------------------------------------------- */
typedef struct {
    int    dw_iterations;        /* master rounds (Phase I + transition + Phase II) */
    int    phase1_iterations;
    int    master_columns;
    long long sub_pivots;        /* basis changes + bound flips in the GPU engine */
    double seconds_total;        /* inside dw_solve */
    double seconds_read;         /* LP parsing + block assembly */
    double seconds_subproblems;  /* dwgpu_* calls (device time + transfers) */
    double seconds_master;       /* restricted master solves */
} dw_stats_t;
/* statistics of the last dw_solve() / dw_solve_blocks() of this process */
void dw_last_stats(dw_stats_t *out);

/* dw_solve() on an in-memory instance: K blocks in the engine's descriptor format (include/dwgpu.h: CSR rows,
 * bounds, file and master costs, COO of D_k) and the L linking rows' bounds (+-DWGPU_INF = none; each row must be
 * <=, >= or =, like the reference's master, src/dw_solver.c:321-425).  Skips the LP text files — the reference
 * parses them under a global mutex (src/dw_subprob.c:105-111), which dominates at thousands of blocks
 * (SURVEY.md §8f-2).  relaxed_solution names the variables x_<k>_<j>. */
struct dwgpu_block_desc;
int dw_solve_blocks(int K, int L, const struct dwgpu_block_desc *blocks, const double *link_lb, const double *link_ub,
                    const dw_options_t *opts, dw_result_t *result);


/* Binary block-angular container (replaces, for repeated solves, the LP-text input path
 * src/dw_main.c:213-259 -> glp_read_lp at src/dw_subprob.c:105-111 and src/dw_solver.c:192-198, and the by-name
 * column mapping src/dw_subprob.c:203-258).  dw_write_container() reads the master and block LP files exactly like
 * dw_solve() and stores the assembled blocks (CSR rows, bounds with the engine's "no bound" value, file and master
 * costs, COO of D_k in the reference's order, linking-row bounds, column names) in one file; dw_solve_container()
 * loads it with one read, hands the arrays to the engine without a copy and then runs the same solve as dw_solve()
 * (same status codes, same relaxed_solution file with the original variable names; opts->shift == 0 means "use the
 * stored constant").
 * CLI: `dwsolver -g guidefile --write-container file` converts, `dwsolver --container file` solves. */
int dw_write_container(const char *master_file, const char *const *subproblem_files, int num_subproblems,
                       double shift /* the guidefile's objective constant, stored with the instance */,
                       const char *container_path);
int dw_solve_container(const char *container_path, const dw_options_t *opts, dw_result_t *result);

#ifdef __cplusplus
}
#endif
#endif
