/*
 * dwgpu.h — C ABI of the B200 batched subproblem engine.
 *
 * This is the seam that replaces the reference's pthread/semaphore hand-off between the
 * master thread and the K subproblem threads (all paths relative to /root/reference):
 *
 *   reference (what exists)                                   replacement entry point
 *   ---------------------------------------------------------------------------------------
 *   pthread_create(subproblem_thread, &sub_data[k])           dwgpu_create()
 *     src/dw_solver.c:154-184; block set-up
 *     src/dw_subprob.c:100-120 (bounds clamp), :203-258 (c_k, COO D_k)
 *   first proposal: cold glp_simplex with the file objective  dwgpu_initial_solve()
 *     src/dw_subprob.c:121-122, 262-269
 *   one round: md->row_duals / phase_one published, cond      dwgpu_iterate()  (host buffers)
 *     broadcast, K x { D_k'pi, d = c_k - y, glp_simplex,      dwgpu_iterate_device() (device
 *     x_k, D_k x_k compaction, c_k.x_k }, K x sem_post          buffers, caller's stream)
 *     src/dw_phases.c:526-548 (down-link), src/dw_subprob.c:291-373 (worker),
 *     src/dw_phases.c:338-350 (up-link)
 *   mailbox fields obj / unbounded / new_obj_coeff / len /    struct dwgpu_proposals
 *     ind / val / current_solution  src/dw_subprob.h:44-94
 *   COMMAND_STOP + pthread_join, free_sub_data                dwgpu_destroy()
 *     src/dw_solver.c:681-688, 757-772; src/dw_support.c:925-960
 *
 * Conventions: plain C, plain pointers and sizes.  Every function returns 0 on success and
 * a negative DWGPU_E_* code otherwise (the host library maps any failure to
 * DW_STATUS_ERR_INTERNAL = 99); nothing in this layer calls exit().  One host thread drives
 * one engine (like dw_solve(), the engine is not re-entrant).  There is no CPU fallback: if
 * no CUDA device is usable, dwgpu_create() fails with DWGPU_E_CUDA.
 */
#ifndef DWGPU_H_
#define DWGPU_H_

#ifdef __cplusplus
extern "C" {
#endif

#define DWGPU_INF 1e30 /* |bound| >= DWGPU_INF (or IEEE inf) means "no bound" */

/* per-block solution status: numerically equal to GLPK's glp_get_status() values, which is
 * what the reference stores in subprob_struct.unbounded (src/dw_subprob.c:340) */
#define DWGPU_UNDEF  1 /* iteration limit / numerical failure */
#define DWGPU_NOFEAS 4
#define DWGPU_OPT    5
#define DWGPU_UNBND  6

#define DWGPU_E_ARGS   (-1)
#define DWGPU_E_CUDA   (-2)
#define DWGPU_E_NOMEM  (-3)
#define DWGPU_E_STATE  (-4)

/* One block k, as the reference's worker assembles it.  Indices 0-based. */
typedef struct dwgpu_block_desc {
    int m, n;                /* rows / columns of A_k */
    const int    *a_rowptr;  /* CSR of A_k: m+1 */
    const int    *a_col;
    const double *a_val;
    const double *row_lb, *row_ub; /* m  (FX rows: lb == ub) */
    const double *col_lb, *col_ub; /* n */
    const double *c_file;    /* n: objective of the block's own LP file (cold solve) */
    const double *c_master;  /* n: master cost of each block column (c_k) */
    int d_nnz;               /* COO of D_k in the reference's append order */
    const int    *d_row;     /* master row   (0-based) */
    const int    *d_col;     /* local column (0-based) */
    const double *d_val;
} dwgpu_block_desc;

typedef struct dwgpu_options {
    int    device;        /* CUDA device ordinal (default 0) */
    int    clamp_lo_cols; /* 1: finite-lb/no-ub columns get ub = 3000 (dw_subprob.c:116-120) */
    int    force_path;    /* 0 auto; 1 shared-memory tableau; 2 legacy all-global kernel;
                             3 global-memory tableau + shared-memory vectors (one CTA per block);
                             4 thread-block cluster per block, tableau striped over the CTAs in HBM;
                             5 packed sparse tableau resident in shared memory (opt-in, also DWGPU_SPARSE=1: blocks of
                               the path-3 class with m <= 1024, n <= 512 and a sparse A_k; hands a block that outgrows
                               its heap over to path 3) */
    int    max_iter_mult; /* iteration limit = max_iter_mult*(m+n) + 1000 (default 50) */
    double tol_dj;        /* dual feasibility   (default 1e-9) */
    double tol_bnd;       /* primal feasibility, scaled by 1+|bound| (default 1e-9) */
    double tol_piv;       /* minimum |pivot|    (default 1e-9) */
    int    cluster_size;  /* path 4: CTAs per block (power of two <= 16); 0 = as many as keep every
                             block of the shard co-resident (148 SMs / blocks) */
    int    no_dual_start; /* path 4: 1 = never repair a primal-infeasible start with the dual simplex */
    int    reserved[6];
} dwgpu_options;

/* Host-side result buffers for one round; any pointer may be NULL (that field is skipped).
 * ind/val rows have stride L: block k's compacted column is ind[k*L .. k*L+len[k]).
 * x is the concatenation of all blocks' solutions (offsets = prefix sums of n_k). */
typedef struct dwgpu_proposals {
    double *obj;    /* K: optimum of the priced LP, d_k . x_k          (mailbox: obj) */
    double *cost;   /* K: c_k . x_k                                    (new_obj_coeff) */
    int    *status; /* K: DWGPU_OPT / NOFEAS / UNBND / UNDEF           (unbounded) */
    int    *len;    /* K: exact nonzeros of D_k x_k                    (len) */
    int    *ind;    /* K*L: 1-based master rows                        (ind[1..len]) */
    double *val;    /* K*L                                             (val[1..len]) */
    double *x;      /* sum n_k                                         (current_solution) */
} dwgpu_proposals;

/* One round's results in compact form, in engine-owned PINNED host memory that the GPU writes
 * directly (valid until the next call on this engine).  Block k's compacted column is
 * ind/val[colptr[k] .. colptr[k+1]) — the K mailboxes of src/dw_subprob.h:44-94 laid end to end. */
typedef struct dwgpu_packed {
    const double *obj;    /* K */
    const double *cost;   /* K */
    const int    *status; /* K */
    const int    *colptr; /* K+1 */
    const int    *ind;    /* nnz: 1-based master rows */
    const double *val;    /* nnz */
    long long     nnz;
} dwgpu_packed;

typedef struct dwgpu_counters {
    long long pivots;      /* basis changes (a rank-1 tableau update happened) */
    long long flips;       /* bound flips (no update) */
    long long rebuilds;    /* tableau rebuilt from A_k after a residual check */
    long long phase1_iter; /* iterations spent restoring primal feasibility */
    long long solves;      /* block solves */
    long long rows_swept;  /* tableau rows (incl. the reduced-cost row) touched by basis changes */
    long long cells_swept; /* tableau doubles rewritten by basis changes: only (active row) x (non-zero
                              item of the pivot row) cells move, so this is the algorithmic traffic / 16 B */
    long long price_cells; /* tableau doubles READ to (re)compute price rows: reduced costs at the start of a
                              solve (only the rows whose basic cost changed), phase-1 prices, beta rebuilds */
    long long mailbox_bytes; /* bytes the GPU stored into the host mailboxes (dwgpu_*_mailboxes) */
} dwgpu_counters;

typedef struct dwgpu_engine dwgpu_engine;

void dwgpu_options_init(dwgpu_options *opt);

/* Copies everything it needs; the descriptors may be freed after the call. */
int  dwgpu_create(int K, int L, const dwgpu_block_desc *blocks, const dwgpu_options *opt,
                  dwgpu_engine **out);
void dwgpu_destroy(dwgpu_engine *e);

/* Back to the state right after dwgpu_create(): all-logical basis, T = A_k (what a fresh
 * glp_read_lp gives the reference, src/dw_subprob.c:105-111).  Asynchronous on `stream`
 * (NULL = the engine's stream).  Counters are kept. */
int  dwgpu_reset(dwgpu_engine *e, void *stream);

/* Cold solve of every block with its file objective; fills *out (may be NULL). */
int  dwgpu_initial_solve(dwgpu_engine *e, dwgpu_proposals *out);

/* One round with HOST buffers: pi (L doubles) is copied to the device, every block is priced
 * and re-solved from its stored basis, and the proposals are copied back into *out.
 *
 * Duals as applied (all dwgpu_iterate* / dwgpu_round_comm calls, engines with blocks on the HBM-tableau path): the
 * engine remembers per linking row the dual it last applied.  A dual that differs from that value by no more than
 * 1e-13 (1 + |pi_r|) — the round-off level of the master's solve — is taken AT that value; the rows that moved further
 * are applied as given.  Rounds in which at most 16 rows moved ("delta rounds", csrc/delta_tg.cuh) look only at the
 * columns on those rows.  obj is therefore the block's optimum at a vector within 1e-13 relative of pi per row (its
 * effect on obj and on the reduced costs is three to four orders below the 1e-9 / 1e-7 tolerances of the path).  The
 * first round after a cold solve, dwgpu_reset() or a phase change applies every row as given; DWGPU_DELTA=0 in the
 * environment of dwgpu_create() switches the mechanism off. */
int  dwgpu_iterate(dwgpu_engine *e, const double *pi, int phase_one, dwgpu_proposals *out);

/* The same two calls with the proposals returned as ONE packed record (a single device-written
 * pinned buffer, one synchronisation, no stride-L arrays over PCIe).  r (K doubles, may be NULL)
 * are the convexity-row duals the master publishes next to pi (sub_data[k].r,
 * src/dw_phases.c:276-279, 531-534): when given, only the columns the master would accept,
 * r_k - obj_k > 1e-6 (src/dw_phases.c:108, 358; TOLERANCE src/dw.h:71), are shipped; the others
 * come back with an empty column.  obj/cost/status are always filled for every block. */
int  dwgpu_initial_solve_packed(dwgpu_engine *e, dwgpu_packed *out);
int  dwgpu_iterate_packed(dwgpu_engine *e, const double *pi, int phase_one, const double *r,
                          dwgpu_packed *out);

/* The reference's up-link as it is: K per-block mailboxes {obj, new_obj_coeff, status, len, ind[1..len], val[1..len]}
 * (/root/reference/src/dw_subprob.h:44-94), here six host arrays in pinned memory that the solve kernels' epilogues
 * store into directly while the round runs (ind/val with stride L, row k holds len[k] entries, 1-based master rows).
 * No pack kernels and no copy after the solve: the call returns after one stream synchronisation.  A block whose
 * state did not change in a round only rewrites its obj — the rest of its mailbox still holds.  The arrays belong to
 * the engine and are overwritten by the next *_mailboxes call. */
typedef struct dwgpu_mailboxes {
    const double *obj, *cost, *val;
    const int *status, *len, *ind;
    int K, L;
} dwgpu_mailboxes;
int  dwgpu_initial_solve_mailboxes(dwgpu_engine *e, dwgpu_mailboxes *out);
int  dwgpu_iterate_mailboxes(dwgpu_engine *e, const double *pi, int phase_one, dwgpu_mailboxes *out);

/* Multi-GPU up-link: pack K_total proposals that an all-gather left in DEVICE arrays (same layout as
 * dwgpu_device_results(), ind/val with stride L) into one pinned record on the calling rank, on the
 * caller's stream; synchronises that stream.  r_dev: K_total convexity duals on the device, or NULL. */
int  dwgpu_pack_gathered(dwgpu_engine *e, int K_total, const dwgpu_proposals *dev_arrays, const double *r_dev,
                         void *stream, dwgpu_packed *out);

/* The same for ONE all-gather per round: every rank's results live in one contiguous device record
 * (dwgpu_device_record: [obj K | cost K | val K*L] doubles, then [status K | len K | ind K*L] ints; all
 * ranks must hold the same K); `records` is the rank-major concatenation an all-gather of those
 * buffers produces. */
int  dwgpu_device_record(dwgpu_engine *e, void **dev_ptr, size_t *bytes);
int  dwgpu_pack_gathered_records(dwgpu_engine *e, int world, const void *records, const double *r_dev,
                                 void *stream, dwgpu_packed *out);

/* Cold solve on the caller's stream, results left on the device (see dwgpu_iterate_device). */
int  dwgpu_initial_solve_device(dwgpu_engine *e, void *stream);

/* One round with DEVICE buffers on the caller's stream (cudaStream_t passed as void*; NULL =
 * the engine's own stream).  Asynchronous; results stay on the device, see
 * dwgpu_device_results().  d_pi must stay valid until the stream reaches the end of the call. */
int  dwgpu_iterate_device(dwgpu_engine *e, const double *d_pi, int phase_one, void *stream);
/* Stream contract: a round or reset enqueued on a caller's stream is followed by an event; every entry point that reads
 * the engine's device arrays afterwards (fetch_*, history_*, get_*, *_packed) makes the engine's own stream wait for it,
 * so no synchronisation by the caller is needed between e.g. dwgpu_iterate_device(e, pi, phase, stream) and
 * dwgpu_fetch_results(e, &out). */

/* Device-resident result arrays of the last round (same layout as dwgpu_proposals). */
int  dwgpu_device_results(dwgpu_engine *e, dwgpu_proposals *dev_ptrs);
/* Copy the last round's results to host buffers (synchronises the engine stream). */
int  dwgpu_fetch_results(dwgpu_engine *e, dwgpu_proposals *out);
/* x_k of the last round for one block (n_k doubles). */
int  dwgpu_fetch_x(dwgpu_engine *e, int k, double *x);

/* Device-resident proposal history — what the reference keeps on the host as fg->x[k][iter], a copy of
 * current_solution per accepted column (src/dw_phases.c:173-179, 427-433), read back only by the final
 * reconstruction x_k = sum_t lambda_{k,t} x_k^t (src/dw_rounding.c:120-168).
 * dwgpu_history_push: keep the CURRENT x_k of the listed blocks (one launch, device to device); ids_out[i]
 *   identifies the kept copy of blocks[i].
 * dwgpu_history_combine: x_out (host, sum n_k doubles, block-major like dwgpu_proposals.x) =
 *   sum over the terms of weight[t] * kept copy ids[t] (added into the block it was pushed from; terms of a
 *   block are applied in the order given). */
int  dwgpu_history_push(dwgpu_engine *e, int count, const int *blocks, long long *ids_out);
int  dwgpu_history_combine(dwgpu_engine *e, int nterms, const long long *ids, const double *weights, double *x_out);

/* Several dual vectors per round (multi-column pricing; the reference prices one vector per round,
 * src/dw_phases.c:358-437, src/dw_subprob.c:291-329).  Pi holds P (1..32) dual vectors of L doubles each; for p = 0..P-1
 * every block is priced with Pi[p] and re-solved from the basis the previous vector left; out[p] is the packed record of
 * that pass (every column shipped; valid until the next call) and ids_out (NULL, or P*K entries) receives the history id
 * of every block's x_k of that pass (as dwgpu_history_push would return).  Blocks whose D_k is dense enough — at least
 * 1/8 of L x n, cluster path — get all P priced cost vectors from ONE FP64 tensor-core GEMM D_k' [pi_0 .. pi_P-1]
 * (dwgpu_dense_pricing_blocks says how many blocks that is). */
int  dwgpu_iterate_multi(dwgpu_engine *e, const double *Pi, int P, int phase_one, dwgpu_packed *out, long long *ids_out);
int  dwgpu_dense_pricing_blocks(dwgpu_engine *e);

int  dwgpu_get_counters(dwgpu_engine *e, dwgpu_counters *total);
/* per-block pivots+flips of the LAST round (K long longs) — load-balance diagnostics */
int  dwgpu_get_block_iterations(dwgpu_engine *e, long long *iters);
/* kernels launched by the last iterate / initial_solve call */
int  dwgpu_last_launch_count(dwgpu_engine *e);
/* per block: 1 = shared-memory tableau, 3 = global-memory tableau with shared-memory vectors,
 * 4 = cluster per block (HBM tableau striped over the CTAs), 2 = legacy kernel with everything in
 * global memory */
int  dwgpu_get_block_paths(dwgpu_engine *e, int *paths);
/* device time (CUDA events on the launching stream) of the last round's solve kernels:
 * ms[0] = shared-memory path launch, ms[1] = global-memory path launch (0 when not launched).
 * Synchronises on the end events. */
int  dwgpu_last_kernel_ms(dwgpu_engine *e, float ms[2]);
/* Shared-memory read+write bandwidth of the device in GB/s (FP64 read-modify-write of a
 * ~100 KB tile per CTA, 2 CTAs per SM, all SMs) — the roofline denominator for the
 * shared-memory path, which MEASURED_PEAKS.json does not carry. */
int  dwgpu_measure_smem_bandwidth(int device, double *gbps);
int  dwgpu_sync(dwgpu_engine *e);
/* Number of CUDA devices this process can use (0 when there is none or the driver is unusable); what
 * DWSOLVER_DEVICES=all expands to when dw_solve shards the blocks over one engine per device. */
int  dwgpu_device_count(void);
const char *dwgpu_last_error(void);
const char *dwgpu_version(void);

/* ---- multi-GPU rounds from the C layer --------------------------------------------------------------
 * One process (or host thread) per GPU, each with an engine over its contiguous block range.  Replaces the reference's
 * cond-broadcast of the duals and its semaphore / queue up-link (src/dw_phases.c:338-350, 526-548;
 * src/dw_subprob.c:417-427):
 *   dwgpu_comm_unique_id / dwgpu_comm_init   an NCCL communicator per engine (NCCL is bound at run time by soname);
 *   dwgpu_uplink_create (rank 0)             one mailbox per block of the WHOLE instance in rank 0's HBM + a 64-byte
 *                                            CUDA IPC handle for the other ranks; dev_out = the arrays (obj, cost,
 *                                            status, len: K_total; ind, val: K_total x L, block-major) for
 *                                            dwgpu_pack_gathered();
 *   dwgpu_uplink_attach (every rank)         maps the buffer (NULL handle on rank 0) and selects this rank's window;
 *   dwgpu_round_comm                         duals down from rank 0 -> solve kernels whose epilogues store the proposals
 *                                            into rank 0's buffer over NVLink peer memory (only what moved) -> close;
 *                                            all on one stream.  The two hand-shakes go through a control block at the
 *                                            end of the up-link buffer (sequence number + completion counter, system-scope
 *                                            loads / stores over NVLink; a wait of more than 10 s traps); with
 *                                            DWGPU_COMM=nccl in the environment they are an ncclBroadcast of pi and a
 *                                            one-int ncclAllReduce instead.  Every rank must make the same sequence of
 *                                            dwgpu_round_comm calls. */
int  dwgpu_comm_unique_id(void *id128);
int  dwgpu_comm_init(dwgpu_engine *e, const void *id128, int world, int rank);
void dwgpu_comm_release(dwgpu_engine *e);
int  dwgpu_uplink_create(dwgpu_engine *e, int K_total, void *ipc_handle64, dwgpu_proposals *dev_out);
int  dwgpu_uplink_attach(dwgpu_engine *e, const void *ipc_handle64, int K_total, int k_first);
int  dwgpu_round_comm(dwgpu_engine *e, double *pi_dev, int phase_one, int initial, void *stream);

/* ---- the restricted master on the device -------------------------------------------------------
 * Replaces the reference's glp_simplex(master_lp) calls (src/dw_phases.c:219, 450; src/dw_solver.c:561-562) for
 * large instances: a primal-dual interior-point solve (Mehrotra predictor-corrector) of
 *     min c'x   s.t.   A x = b,  x >= 0,
 * where rows 0..L-1 are the linking rows (all turned into equalities with explicit slack / artificial columns, as
 * src/dw_solver.c:321-419 does) and rows L..L+K-1 the convexity rows; a column has at most ONE entry in the
 * convexity rows, with coefficient 1 (its block).  The method works on an L x L system per iteration, formed with
 * FP64 tensor-core MMAs from all master columns (csrc/master_ipm.cu).  Columns are added in CSC form over all L+K
 * rows; set_costs / set_coef / del_cols mirror glp_set_obj_coef / glp_set_mat_col / glp_del_cols.
 * dwgpu_master_solve: *status = 0 optimal to `tol` (relative primal and dual infeasibility and duality gap),
 *   3 = not converged within max_iter iterations (a solve whose best iterate is within 100 tol counts as optimal: the
 *   FP64 normal equations lose accuracy in the last iterations; dwgpu_master_quality says what was reached).  dwgpu_master_get: objective c'x, dual objective b'y, the L+K row
 *   duals y (the pi and r_k of src/dw_support.c:142-152), the column values x, iterations and device seconds of the
 *   last solve (any pointer may be NULL). */
typedef struct dwgpu_master dwgpu_master;
int  dwgpu_master_create(int device, int L, int K, dwgpu_master **out);
void dwgpu_master_destroy(dwgpu_master *m);
int  dwgpu_master_set_rhs(dwgpu_master *m, const double *b /* L+K */);
int  dwgpu_master_add_cols(dwgpu_master *m, int ncols, const int *colptr, const int *rows, const double *vals, const double *cost);
int  dwgpu_master_set_costs(dwgpu_master *m, int first, int count, const double *cost);
int  dwgpu_master_set_coef(dwgpu_master *m, int col, int pos, double v);
int  dwgpu_master_del_cols(dwgpu_master *m, const char *flag /* one per column, non-zero = delete */);
int  dwgpu_master_cols(const dwgpu_master *m);
int  dwgpu_master_solve(dwgpu_master *m, double tol, int max_iter, int *status);
int  dwgpu_master_get(dwgpu_master *m, double *obj, double *dual_obj, double *y, double *x, int *iterations, double *seconds);
int  dwgpu_master_quality(dwgpu_master *m, double *pinf, double *dinf, double *gap);
const char *dwgpu_master_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* DWGPU_H_ */
