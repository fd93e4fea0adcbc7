/*
 * stub_dwgpu.c — TEST DOUBLE of the engine's C ABI (include/dwgpu.h).  TEST INFRASTRUCTURE ONLY.
 *
 * It exists so that the host side of the drop-in (dwsolver_b200/host: dw_solve, the Phase I / II loop, the stop
 * rules, block sharding over several engines, the container path, the command line) can be exercised by the CPU
 * test-suite on a machine without a GPU.  It implements exactly the ten entry points libdwsolver_b200.so imports,
 * on top of the CPU oracle (oracle/), and is built by tests/stub_engine/Makefile into tests/_build/lib/ — a private
 * directory into which the tests COPY the host library and the CLI, so that their `$ORIGIN` run-path finds this
 * file there instead of the CUDA engine.  Nothing in dwsolver_b200/ or include/ refers to it, it is never installed
 * next to the product, and the product has no way to select it: dwsolver_b200/lib/libdwgpu.so is the CUDA engine and
 * fails loudly without a device (tests/test_abi.py).  dwgpu_create() here refuses to work unless the test-suite
 * announces itself (DWSOLVER_TEST_DOUBLE=tests-only).  Results obtained through this file say nothing about the GPU
 * path; the parity tests proper are the `-m gpu` ones.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/dwgpu.h"
#include "../../oracle/oracle.h"

struct dwgpu_engine {
    int K, L;
    orc_block **blocks;
    orc_proposal *props;
    int *n;
    long long *xoff, ntot;
    /* packed record (what the CUDA engine keeps in pinned memory) */
    double *obj, *cost, *val;
    int *status, *colptr, *ind;
    int packed_rounds;              /* packed rounds served so far (failure injection) */
    /* records of a multi-vector round */
    int multi_P;
    double *m_obj, *m_cost, *m_val;
    int *m_status, *m_colptr, *m_ind;
    /* kept copies of x_k */
    double *hist;
    long long hist_used, hist_cap;
    long long *hist_off;
    int *hist_block;
    long long hist_n, hist_slots;
};

static __thread char g_err[256];
static int fail(int code, const char *msg) { snprintf(g_err, sizeof g_err, "%s", msg); return code; }

const char *dwgpu_last_error(void) { return g_err; }
int dwgpu_device_count(void) { return 1; }

void dwgpu_options_init(dwgpu_options *o)
{
    if (!o) return;
    memset(o, 0, sizeof *o);
    o->clamp_lo_cols = 1; o->max_iter_mult = 50; o->tol_dj = o->tol_bnd = o->tol_piv = 1e-9;
}

void dwgpu_destroy(dwgpu_engine *e)
{
    if (!e) return;
    for (int k = 0; k < e->K; ++k) {
        if (e->blocks && e->blocks[k]) orc_block_free(e->blocks[k]);
        if (e->props) { free(e->props[k].ind); free(e->props[k].val); free(e->props[k].x); }
    }
    free(e->blocks); free(e->props); free(e->n); free(e->xoff);
    free(e->obj); free(e->cost); free(e->val); free(e->status); free(e->colptr); free(e->ind);
    free(e->hist); free(e->hist_off); free(e->hist_block);
    free(e->m_obj); free(e->m_cost); free(e->m_val); free(e->m_status); free(e->m_colptr); free(e->m_ind);
    free(e);
}

int dwgpu_create(int K, int L, const dwgpu_block_desc *bd, const dwgpu_options *opt, dwgpu_engine **out)
{
    (void)opt;
    if (K < 1 || L < 0 || !bd || !out) return fail(DWGPU_E_ARGS, "bad arguments");
    /* refuses to work anywhere but under the test-suite, which announces itself */
    if (!getenv("DWSOLVER_TEST_DOUBLE") || strcmp(getenv("DWSOLVER_TEST_DOUBLE"), "tests-only") != 0)
        return fail(DWGPU_E_STATE, "this libdwgpu.so is the test double of tests/stub_engine, not the CUDA engine");
    if (getenv("DWSTUB_FAIL_CREATE")) return fail(DWGPU_E_CUDA, "stub engine: creation failure requested by the test");
    dwgpu_engine *e = calloc(1, sizeof *e);
    if (!e) return fail(DWGPU_E_NOMEM, "out of memory");
    e->K = K; e->L = L;
    e->blocks = calloc((size_t)K, sizeof *e->blocks);
    e->props = calloc((size_t)K, sizeof *e->props);
    e->n = calloc((size_t)K, sizeof *e->n);
    e->xoff = calloc((size_t)K + 1, sizeof *e->xoff);
    e->obj = calloc((size_t)K, sizeof(double)); e->cost = calloc((size_t)K, sizeof(double));
    e->status = calloc((size_t)K, sizeof(int)); e->colptr = calloc((size_t)K + 1, sizeof(int));
    e->ind = calloc((size_t)K * (size_t)(L > 0 ? L : 1), sizeof(int));
    e->val = calloc((size_t)K * (size_t)(L > 0 ? L : 1), sizeof(double));
    for (int k = 0; k < K; ++k) {
        const dwgpu_block_desc *b = &bd[k];
        e->n[k] = b->n;
        e->xoff[k + 1] = e->xoff[k] + b->n;
        e->blocks[k] = orc_block_new(b->m, b->n, b->a_rowptr, b->a_col, b->a_val, b->row_lb, b->row_ub, b->col_lb, b->col_ub,
                                     b->c_file, b->c_master, L, b->d_nnz, b->d_row, b->d_col, b->d_val);
        e->props[k].ind = calloc((size_t)(L > 0 ? L : 1) + 1, sizeof(int));
        e->props[k].val = calloc((size_t)(L > 0 ? L : 1) + 1, sizeof(double));
        e->props[k].x = calloc((size_t)(b->n > 0 ? b->n : 1), sizeof(double));
        if (!e->blocks[k] || !e->props[k].ind || !e->props[k].val || !e->props[k].x) { dwgpu_destroy(e); return fail(DWGPU_E_NOMEM, "out of memory"); }
    }
    e->ntot = e->xoff[K];
    *out = e;
    return 0;
}

/* src/dw_phases.c:108, 358 through the engine's packed record: with r given only the columns the master will accept
 * are shipped (dwsolver_b200/csrc/pack.cuh) */
/* worker threads of a round: 2 in the test-suite; bench.py's CPU arm sets DWSTUB_THREADS to the host's core count */
static int stub_threads(void)
{
    const char *t = getenv("DWSTUB_THREADS");
    const int n = t ? atoi(t) : 2;
    return n < 1 ? 1 : n;
}

static void pack(dwgpu_engine *e, const double *r, dwgpu_packed *out)
{
    int nnz = 0;
    for (int k = 0; k < e->K; ++k) {
        const orc_proposal *p = &e->props[k];
        e->obj[k] = p->obj; e->cost[k] = p->new_obj_coeff; e->status[k] = p->status;
        e->colptr[k] = nnz;
        /* failure injection: DWSTUB_STATUS="<round>:<block>:<status>" makes one block report that status in that
           packed round (0 = the initial solve) with an attractive objective, as a kernel failure exit would */
        if (getenv("DWSTUB_STATUS")) {
            int rd = -1, bk = -1, st = 0;
            if (sscanf(getenv("DWSTUB_STATUS"), "%d:%d:%d", &rd, &bk, &st) == 3 && rd == e->packed_rounds && bk == k) {
                e->status[k] = st;
                e->obj[k] = -1e9;
            }
        }
        if (r && (e->status[k] == DWGPU_UNDEF || e->status[k] == DWGPU_NOFEAS)) continue;
        if (r && !(r[k] - e->obj[k] > 1e-6)) continue;
        for (int t = 0; t < p->len; ++t) { e->ind[nnz] = p->ind[t]; e->val[nnz] = p->val[t]; ++nnz; }
    }
    e->colptr[e->K] = nnz;
    e->packed_rounds++;
    if (out) {
        out->obj = e->obj; out->cost = e->cost; out->status = e->status; out->colptr = e->colptr;
        out->ind = e->ind; out->val = e->val; out->nnz = nnz;
    }
}

int dwgpu_initial_solve_packed(dwgpu_engine *e, dwgpu_packed *out)
{
    if (!e) return fail(DWGPU_E_ARGS, "bad arguments");
    orc_batch(e->blocks, e->props, e->K, NULL, 0, 1, stub_threads());
    pack(e, NULL, out);
    return 0;
}

int dwgpu_iterate_packed(dwgpu_engine *e, const double *pi, int phase_one, const double *r, dwgpu_packed *out)
{
    if (!e || (e->L > 0 && !pi)) return fail(DWGPU_E_ARGS, "bad arguments");
    if (getenv("DWSTUB_FAIL_ITERATE")) return fail(DWGPU_E_CUDA, "stub engine: round failure requested by the test");
    orc_batch(e->blocks, e->props, e->K, pi, phase_one, 0, stub_threads());
    pack(e, r, out);
    return 0;
}

int dwgpu_history_push(dwgpu_engine *e, int count, const int *blocks, long long *ids_out)
{
    if (!e || count < 0 || (count > 0 && (!blocks || !ids_out))) return fail(DWGPU_E_ARGS, "bad arguments");
    for (int i = 0; i < count; ++i) {
        const int k = blocks[i];
        if (k < 0 || k >= e->K) return fail(DWGPU_E_ARGS, "block index out of range");
        if (e->hist_used + e->n[k] > e->hist_cap) {
            e->hist_cap = 2 * (e->hist_used + e->n[k]) + 2 * e->ntot;
            e->hist = realloc(e->hist, sizeof(double) * (size_t)e->hist_cap);
        }
        if (e->hist_n == e->hist_slots) {
            e->hist_slots = 2 * e->hist_slots + 64;
            e->hist_off = realloc(e->hist_off, sizeof(long long) * (size_t)e->hist_slots);
            e->hist_block = realloc(e->hist_block, sizeof(int) * (size_t)e->hist_slots);
        }
        if (!e->hist || !e->hist_off || !e->hist_block) return fail(DWGPU_E_NOMEM, "out of memory");
        memcpy(e->hist + e->hist_used, e->props[k].x, sizeof(double) * (size_t)e->n[k]);
        ids_out[i] = e->hist_n;
        e->hist_off[e->hist_n] = e->hist_used; e->hist_block[e->hist_n] = k; ++e->hist_n;
        e->hist_used += e->n[k];
    }
    return 0;
}

int dwgpu_history_combine(dwgpu_engine *e, int nterms, const long long *ids, const double *weights, double *x_out)
{
    if (!e || nterms < 0 || !x_out || (nterms > 0 && (!ids || !weights))) return fail(DWGPU_E_ARGS, "bad arguments");
    for (long long j = 0; j < e->ntot; ++j) x_out[j] = 0.0;
    for (int t = 0; t < nterms; ++t) {
        if (ids[t] < 0 || ids[t] >= e->hist_n) return fail(DWGPU_E_ARGS, "unknown history id");
        const int k = e->hist_block[ids[t]];
        const double *src = e->hist + e->hist_off[ids[t]];
        double *dst = x_out + e->xoff[k];
        for (int j = 0; j < e->n[k]; ++j) dst[j] = fma(weights[t], src[j], dst[j]);
    }
    return 0;
}

int dwgpu_get_counters(dwgpu_engine *e, dwgpu_counters *c)
{
    if (!e || !c) return fail(DWGPU_E_ARGS, "bad arguments");
    memset(c, 0, sizeof *c);
    for (int k = 0; k < e->K; ++k) {
        long p = 0, f = 0, rf = 0;
        orc_block_counters(e->blocks[k], &p, &f, &rf);
        c->pivots += p; c->flips += f; c->rebuilds += rf;
    }
    return 0;
}

/* several dual vectors per round (include/dwgpu.h: dwgpu_iterate_multi): P oracle passes, one packed record each
 * (every column shipped), every block's x_k of every pass kept in the history */
int dwgpu_iterate_multi(dwgpu_engine *e, const double *Pi, int P, int phase_one, dwgpu_packed *out, long long *ids_out)
{
    if (!e || P < 1 || P > 32 || !out || (e->L > 0 && !Pi)) return fail(DWGPU_E_ARGS, "bad arguments");
    const size_t K = (size_t)e->K, KL = K * (size_t)(e->L > 0 ? e->L : 1);
    if (P > e->multi_P) {
        e->m_obj = realloc(e->m_obj, sizeof(double) * K * (size_t)P); e->m_cost = realloc(e->m_cost, sizeof(double) * K * (size_t)P);
        e->m_status = realloc(e->m_status, sizeof(int) * K * (size_t)P); e->m_colptr = realloc(e->m_colptr, sizeof(int) * (K + 1) * (size_t)P);
        e->m_ind = realloc(e->m_ind, sizeof(int) * KL * (size_t)P); e->m_val = realloc(e->m_val, sizeof(double) * KL * (size_t)P);
        if (!e->m_obj || !e->m_cost || !e->m_status || !e->m_colptr || !e->m_ind || !e->m_val) return fail(DWGPU_E_NOMEM, "out of memory");
        e->multi_P = P;
    }
    for (int p = 0; p < P; ++p) {
        orc_batch(e->blocks, e->props, e->K, Pi + (size_t)p * (size_t)e->L, phase_one, 0, stub_threads());
        dwgpu_packed one;
        pack(e, NULL, &one);
        memcpy(e->m_obj + K * (size_t)p, e->obj, sizeof(double) * K); memcpy(e->m_cost + K * (size_t)p, e->cost, sizeof(double) * K);
        memcpy(e->m_status + K * (size_t)p, e->status, sizeof(int) * K); memcpy(e->m_colptr + (K + 1) * (size_t)p, e->colptr, sizeof(int) * (K + 1));
        memcpy(e->m_ind + KL * (size_t)p, e->ind, sizeof(int) * (size_t)one.nnz); memcpy(e->m_val + KL * (size_t)p, e->val, sizeof(double) * (size_t)one.nnz);
        out[p].obj = e->m_obj + K * (size_t)p; out[p].cost = e->m_cost + K * (size_t)p; out[p].status = e->m_status + K * (size_t)p;
        out[p].colptr = e->m_colptr + (K + 1) * (size_t)p; out[p].ind = e->m_ind + KL * (size_t)p; out[p].val = e->m_val + KL * (size_t)p;
        out[p].nnz = one.nnz;
        if (ids_out)
            for (int k = 0; k < e->K; ++k) {
                const int rc = dwgpu_history_push(e, 1, &k, &ids_out[(size_t)p * K + (size_t)k]);
                if (rc) return rc;
            }
    }
    return 0;
}
int dwgpu_dense_pricing_blocks(dwgpu_engine *e) { (void)e; return 0; }

/* ---- the device-resident master (dwgpu_master_*): not part of the test double — creation fails, so the host loop
 * keeps its simplex master (dwsolver_b200/host/ipm_master.cc) ---- */
int dwgpu_master_create(int device, int L, int K, dwgpu_master **out) { (void)device; (void)L; (void)K; if (out) *out = NULL; return DWGPU_E_CUDA; }
void dwgpu_master_destroy(dwgpu_master *m) { (void)m; }
int dwgpu_master_set_rhs(dwgpu_master *m, const double *b) { (void)m; (void)b; return DWGPU_E_STATE; }
int dwgpu_master_add_cols(dwgpu_master *m, int n, const int *cp, const int *r, const double *v, const double *c)
{ (void)m; (void)n; (void)cp; (void)r; (void)v; (void)c; return DWGPU_E_STATE; }
int dwgpu_master_set_costs(dwgpu_master *m, int f, int n, const double *c) { (void)m; (void)f; (void)n; (void)c; return DWGPU_E_STATE; }
int dwgpu_master_set_coef(dwgpu_master *m, int c, int p, double v) { (void)m; (void)c; (void)p; (void)v; return DWGPU_E_STATE; }
int dwgpu_master_del_cols(dwgpu_master *m, const char *f) { (void)m; (void)f; return DWGPU_E_STATE; }
int dwgpu_master_cols(const dwgpu_master *m) { (void)m; return 0; }
int dwgpu_master_solve(dwgpu_master *m, double tol, int it, int *st) { (void)m; (void)tol; (void)it; (void)st; return DWGPU_E_STATE; }
int dwgpu_master_get(dwgpu_master *m, double *o, double *d, double *y, double *x, int *it, double *s)
{ (void)m; (void)o; (void)d; (void)y; (void)x; (void)it; (void)s; return DWGPU_E_STATE; }
int dwgpu_master_quality(dwgpu_master *m, double *p, double *d, double *g) { (void)m; (void)p; (void)d; (void)g; return DWGPU_E_STATE; }
const char *dwgpu_master_last_error(void) { return "the engine test double has no device master"; }
