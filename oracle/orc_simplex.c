/*
 * orc_simplex.c — plain dense bounded-variable primal simplex.  TEST INFRASTRUCTURE ONLY.
 *
 * Stands in for the reference's glp_simplex() call per block
 * (/root/reference/src/dw_subprob.c:121 cold, :329 warm).  GLPK's source is not in the
 * reference tree, so this is not a restatement of GLPK's code: it is the textbook method
 * (Chvatal ch. 8; Maros ch. 9), kept deliberately simple and conservative so that it can
 * serve as the checker for the CUDA kernels:
 *   - condensed (Tucker) tableau  x_B = T x_N  over the m auxiliary (row) variables and n
 *     structural variables, GLPK-style: row i is  aux_i = sum_j a_ij x_j  with the row
 *     bounds carried by aux_i;
 *   - Dantzig pricing, textbook ratio test (exact minimum, ties -> largest |pivot|),
 *     Bland's rule after a run of degenerate steps;
 *   - composite phase 1 (minimise the sum of infeasibilities, re-priced every step);
 *   - warm start: the basis persists in the object between orc_lp_solve() calls, exactly
 *     like the glp_prob the reference keeps per thread;
 *   - at "optimal", the primal residual against the ORIGINAL rows is checked and the
 *     tableau is rebuilt from A if it drifted (then the solve continues).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

#define TOL_DJ   1e-9   /* dual feasibility */
#define TOL_BND  1e-9   /* primal feasibility (scaled by 1+|bound|) */
#define TOL_PIV  1e-9   /* smallest acceptable |pivot| */
#define TOL_TIE  1e-12  /* ratio ties */
#define TOL_RES  1e-10  /* residual that triggers a rebuild */

enum { AT_LO = 0, AT_UP = 1, AT_FREE = 2, AT_FIX = 3 };

struct orc_lp {
    int m, n;
    int *rp, *ci;      /* A, CSR */
    double *av;
    double *lo, *up;   /* bounds of all m+n variables: [0,m) aux, [m,m+n) structural */
    double *T;         /* m x n, row-major */
    double *beta;      /* values of basic variables */
    int *head;         /* variable basic in row i */
    int *nbv;          /* variable nonbasic in column j */
    signed char *stat; /* where nonbasic j sits */
    double *d;         /* reduced costs (phase 2) */
    double *d1;        /* phase-1 reduced costs (scratch) */
    double *colq, *rowp;
    long pivots, flips, refactors;
};

static int is_inf(double v) { return !(fabs(v) < ORC_INF); }

static double nb_value(const orc_lp *lp, int j)
{
    int v = lp->nbv[j];
    switch (lp->stat[j]) {
    case AT_LO: case AT_FIX: return lp->lo[v];
    case AT_UP: return lp->up[v];
    default: return 0.0;
    }
}

static signed char initial_stat(double lo, double up)
{
    if (!is_inf(lo) && !is_inf(up) && lo == up) return AT_FIX;
    if (!is_inf(lo)) return AT_LO;
    if (!is_inf(up)) return AT_UP;
    return AT_FREE;
}

static void load_initial_tableau(orc_lp *lp)
{
    int m = lp->m, n = lp->n;
    memset(lp->T, 0, sizeof(double) * (size_t)m * (size_t)n);
    for (int i = 0; i < m; ++i) {
        for (int p = lp->rp[i]; p < lp->rp[i + 1]; ++p)
            lp->T[(size_t)i * n + lp->ci[p]] += lp->av[p];
        lp->head[i] = i;
    }
    for (int j = 0; j < n; ++j)
        lp->nbv[j] = m + j;
}

static void recompute_beta(orc_lp *lp)
{
    int m = lp->m, n = lp->n;
    for (int i = 0; i < m; ++i) {
        double s = 0.0;
        const double *row = lp->T + (size_t)i * n;
        for (int j = 0; j < n; ++j) {
            double xv = nb_value(lp, j);
            if (xv != 0.0) s += row[j] * xv;
        }
        lp->beta[i] = s;
    }
}

orc_lp *orc_lp_new(int m, int n, const int *rowptr, const int *colidx, const double *vals,
                   const double *row_lb, const double *row_ub,
                   const double *col_lb, const double *col_ub)
{
    orc_lp *lp = calloc(1, sizeof(*lp));
    if (!lp) return NULL;
    int nnz = rowptr[m];
    lp->m = m; lp->n = n;
    lp->rp = malloc(sizeof(int) * (size_t)(m + 1));
    lp->ci = malloc(sizeof(int) * (size_t)(nnz > 0 ? nnz : 1));
    lp->av = malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
    lp->lo = malloc(sizeof(double) * (size_t)(m + n));
    lp->up = malloc(sizeof(double) * (size_t)(m + n));
    lp->T = malloc(sizeof(double) * (size_t)(m > 0 ? m : 1) * (size_t)(n > 0 ? n : 1));
    lp->beta = calloc((size_t)(m > 0 ? m : 1), sizeof(double));
    lp->head = malloc(sizeof(int) * (size_t)(m > 0 ? m : 1));
    lp->nbv = malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    lp->stat = malloc((size_t)(n > 0 ? n : 1));
    lp->d = calloc((size_t)(n > 0 ? n : 1), sizeof(double));
    lp->d1 = calloc((size_t)(n > 0 ? n : 1), sizeof(double));
    lp->colq = calloc((size_t)(m > 0 ? m : 1), sizeof(double));
    lp->rowp = calloc((size_t)(n > 0 ? n : 1), sizeof(double));
    memcpy(lp->rp, rowptr, sizeof(int) * (size_t)(m + 1));
    memcpy(lp->ci, colidx, sizeof(int) * (size_t)nnz);
    memcpy(lp->av, vals, sizeof(double) * (size_t)nnz);
    for (int i = 0; i < m; ++i) {
        lp->lo[i] = is_inf(row_lb[i]) ? -INFINITY : row_lb[i];
        lp->up[i] = is_inf(row_ub[i]) ? INFINITY : row_ub[i];
    }
    for (int j = 0; j < n; ++j) {
        lp->lo[m + j] = is_inf(col_lb[j]) ? -INFINITY : col_lb[j];
        lp->up[m + j] = is_inf(col_ub[j]) ? INFINITY : col_ub[j];
    }
    load_initial_tableau(lp);
    for (int j = 0; j < n; ++j)
        lp->stat[j] = initial_stat(lp->lo[m + j], lp->up[m + j]);
    recompute_beta(lp);
    return lp;
}

void orc_lp_free(orc_lp *lp)
{
    if (!lp) return;
    free(lp->rp); free(lp->ci); free(lp->av); free(lp->lo); free(lp->up); free(lp->T);
    free(lp->beta); free(lp->head); free(lp->nbv); free(lp->stat); free(lp->d); free(lp->d1);
    free(lp->colq); free(lp->rowp); free(lp);
}

void orc_lp_counters(const orc_lp *lp, long *pivots, long *flips, long *refactors)
{
    if (pivots) *pivots = lp->pivots;
    if (flips) *flips = lp->flips;
    if (refactors) *refactors = lp->refactors;
}

/* Tucker pivot on (p, q): x_B[p] <-> x_N[q].  `row_d` (may be NULL) is an extra row
 * (reduced costs) transformed with the body. */
static void pivot(orc_lp *lp, int p, int q, double *row_d)
{
    int m = lp->m, n = lp->n;
    double *T = lp->T;
    double piv = T[(size_t)p * n + q];
    double inv = 1.0 / piv;
    for (int i = 0; i < m; ++i) lp->colq[i] = T[(size_t)i * n + q];
    for (int j = 0; j < n; ++j) lp->rowp[j] = T[(size_t)p * n + j] * inv;
    for (int i = 0; i < m; ++i) {
        if (i == p) continue;
        double ci = lp->colq[i];
        double *row = T + (size_t)i * n;
        if (ci != 0.0) {
            for (int j = 0; j < n; ++j) row[j] -= ci * lp->rowp[j];
        }
        row[q] = ci * inv;
    }
    if (row_d) {
        double dq = row_d[q];
        if (dq != 0.0)
            for (int j = 0; j < n; ++j) row_d[j] -= dq * lp->rowp[j];
        row_d[q] = dq * inv;
    }
    double *prow = T + (size_t)p * n;
    for (int j = 0; j < n; ++j) prow[j] = -lp->rowp[j];
    prow[q] = inv;
    int tmp = lp->head[p]; lp->head[p] = lp->nbv[q]; lp->nbv[q] = tmp;
}

/* rebuild T from A for the current basis (set of basic variables) */
static int rebuild(orc_lp *lp)
{
    int m = lp->m, n = lp->n, tot = m + n;
    char *want_basic = calloc((size_t)tot, 1);
    signed char *vstat = malloc((size_t)tot);
    for (int i = 0; i < m; ++i) want_basic[lp->head[i]] = 1;
    for (int j = 0; j < n; ++j) vstat[lp->nbv[j]] = lp->stat[j];
    load_initial_tableau(lp);
    int ok = 1;
    for (int j = 0; j < n && ok; ++j) {
        int v = m + j;
        if (!want_basic[v]) continue;
        /* structural v sits in column j of the current tableau (columns never move) */
        int best = -1; double bmag = 0.0;
        for (int i = 0; i < m; ++i) {
            int h = lp->head[i];
            if (h >= m || want_basic[h]) continue;       /* row must hold a departing aux */
            double a = fabs(lp->T[(size_t)i * n + j]);
            if (a > bmag) { bmag = a; best = i; }
        }
        if (best < 0 || bmag < 1e-11) { ok = 0; break; }
        pivot(lp, best, j, NULL);
    }
    if (ok) {
        for (int j = 0; j < n; ++j) {
            int v = lp->nbv[j];
            lp->stat[j] = vstat[v];
        }
        recompute_beta(lp);
    }
    free(want_basic); free(vstat);
    lp->refactors++;
    return ok;
}

static void price_phase2(orc_lp *lp, const double *cost)
{
    int m = lp->m, n = lp->n;
    for (int j = 0; j < n; ++j) {
        int v = lp->nbv[j];
        lp->d[j] = v >= m ? cost[v - m] : 0.0;
    }
    for (int i = 0; i < m; ++i) {
        int v = lp->head[i];
        if (v < m) continue;
        double cb = cost[v - m];
        if (cb == 0.0) continue;
        const double *row = lp->T + (size_t)i * n;
        for (int j = 0; j < n; ++j) lp->d[j] += cb * row[j];
    }
}

static double bnd_tol(double b) { return TOL_BND * (1.0 + fabs(b)); }

static void extract_x(const orc_lp *lp, double *x)
{
    int m = lp->m, n = lp->n;
    for (int j = 0; j < n; ++j) {
        int v = lp->nbv[j];
        if (v >= m) x[v - m] = nb_value(lp, j);
    }
    for (int i = 0; i < m; ++i) {
        int v = lp->head[i];
        if (v >= m) x[v - m] = lp->beta[i];
    }
}

/* max |A x - aux| scaled, aux taken from the simplex state */
static double state_residual(const orc_lp *lp, const double *x)
{
    int m = lp->m, n = lp->n;
    double *aux = malloc(sizeof(double) * (size_t)(m > 0 ? m : 1));
    for (int j = 0; j < n; ++j) {
        int v = lp->nbv[j];
        if (v < m) aux[v] = nb_value(lp, j);
    }
    for (int i = 0; i < m; ++i) {
        int v = lp->head[i];
        if (v < m) aux[v] = lp->beta[i];
    }
    double worst = 0.0;
    for (int i = 0; i < m; ++i) {
        double s = 0.0;
        for (int p = lp->rp[i]; p < lp->rp[i + 1]; ++p) s += lp->av[p] * x[lp->ci[p]];
        double e = fabs(s - aux[i]) / (1.0 + fabs(s));
        if (e > worst) worst = e;
    }
    free(aux);
    return worst;
}

double orc_lp_residual(const orc_lp *lp, const double *x)
{
    int m = lp->m, n = lp->n;
    double worst = 0.0;
    for (int i = 0; i < m; ++i) {
        double s = 0.0;
        for (int p = lp->rp[i]; p < lp->rp[i + 1]; ++p) s += lp->av[p] * x[lp->ci[p]];
        if (!is_inf(lp->lo[i])) { double e = (lp->lo[i] - s) / (1.0 + fabs(lp->lo[i])); if (e > worst) worst = e; }
        if (!is_inf(lp->up[i])) { double e = (s - lp->up[i]) / (1.0 + fabs(lp->up[i])); if (e > worst) worst = e; }
    }
    for (int j = 0; j < n; ++j) {
        double lo = lp->lo[m + j], up = lp->up[m + j];
        if (!is_inf(lo)) { double e = (lo - x[j]) / (1.0 + fabs(lo)); if (e > worst) worst = e; }
        if (!is_inf(up)) { double e = (x[j] - up) / (1.0 + fabs(up)); if (e > worst) worst = e; }
    }
    return worst;
}

int orc_lp_solve(orc_lp *lp, const double *cost, double *x, double *objval)
{
    int m = lp->m, n = lp->n;
    int status = ORC_UNDEF;
    long it_limit = 200L * (m + n) + 20000L, it = 0;
    int degen_run = 0, bland = 0, rebuilds = 0;
    int d_valid = 0;

    for (;;) {
        if (++it > it_limit) { status = ORC_UNDEF; break; }
        /* ---- which phase? ---- */
        int infeasible = 0;
        for (int i = 0; i < m; ++i) {
            int v = lp->head[i];
            if (lp->beta[i] < lp->lo[v] - bnd_tol(lp->lo[v]) || lp->beta[i] > lp->up[v] + bnd_tol(lp->up[v])) {
                infeasible = 1; break;
            }
        }
        double *dd;
        if (infeasible) {
            memset(lp->d1, 0, sizeof(double) * (size_t)n);
            for (int i = 0; i < m; ++i) {
                int v = lp->head[i];
                double w = 0.0;
                if (lp->beta[i] < lp->lo[v] - bnd_tol(lp->lo[v])) w = -1.0;
                else if (lp->beta[i] > lp->up[v] + bnd_tol(lp->up[v])) w = 1.0;
                if (w == 0.0) continue;
                const double *row = lp->T + (size_t)i * n;
                for (int j = 0; j < n; ++j) lp->d1[j] += w * row[j];
            }
            dd = lp->d1;
            d_valid = 0;
        } else {
            if (!d_valid) { price_phase2(lp, cost); d_valid = 1; }
            dd = lp->d;
        }
        /* ---- entering variable ---- */
        int q = -1; double best = 0.0;
        for (int j = 0; j < n; ++j) {
            double dj = dd[j];
            int ok = 0;
            switch (lp->stat[j]) {
            case AT_LO: ok = dj < -TOL_DJ; break;
            case AT_UP: ok = dj > TOL_DJ; break;
            case AT_FREE: ok = fabs(dj) > TOL_DJ; break;
            default: ok = 0;
            }
            if (!ok) continue;
            if (bland) {
                if (q < 0 || lp->nbv[j] < lp->nbv[q]) q = j;
            } else if (fabs(dj) > best) { best = fabs(dj); q = j; }
        }
        if (q < 0) {
            if (infeasible) { status = ORC_NOFEAS; break; }
            /* claimed optimal: verify against the original rows */
            extract_x(lp, x);
            if (state_residual(lp, x) > TOL_RES && rebuilds < 5) {
                ++rebuilds;
                if (!rebuild(lp)) { status = ORC_UNDEF; break; }
                d_valid = 0;
                continue;
            }
            status = ORC_OPT;
            break;
        }
        double s = dd[q] < 0.0 ? 1.0 : -1.0;
        /* ---- ratio test ---- */
        int p = -1; double tmin = INFINITY, pmag = 0.0; int p_to_upper = 0;
        for (int i = 0; i < m; ++i) {
            double a = s * lp->T[(size_t)i * n + q];
            if (fabs(a) <= TOL_PIV) continue;
            int v = lp->head[i];
            double lo = lp->lo[v], up = lp->up[v], b = lp->beta[i];
            double t; int to_upper;
            int below = b < lo - bnd_tol(lo), above = b > up + bnd_tol(up);
            if (a > 0.0) {
                if (below) { t = (lo - b) / a; to_upper = 0; }          /* becomes feasible at lo */
                else if (above) continue;                                /* moves further away: priced, not blocked */
                else { if (is_inf(up)) continue; t = (up - b) / a; to_upper = 1; }
            } else {
                if (above) { t = (up - b) / a; to_upper = 1; }
                else if (below) continue;
                else { if (is_inf(lo)) continue; t = (lo - b) / a; to_upper = 0; }
            }
            if (t < 0.0) t = 0.0;
            int take = 0;
            if (t < tmin - TOL_TIE) take = 1;
            else if (t <= tmin + TOL_TIE) {
                if (bland) take = (p < 0 || v < lp->head[p]);
                else take = fabs(a) > pmag;
            }
            if (take) { if (t < tmin) tmin = t; p = i; pmag = fabs(a); p_to_upper = to_upper; }
        }
        int vq = lp->nbv[q];
        double range = (lp->stat[q] == AT_FREE) ? INFINITY : lp->up[vq] - lp->lo[vq];
        if (p >= 0) {
            /* the recorded row might not be the exact minimiser after tie handling */
            double a = s * lp->T[(size_t)p * n + q];
            int v = lp->head[p];
            double target = p_to_upper ? lp->up[v] : lp->lo[v];
            tmin = (target - lp->beta[p]) / a;
            if (tmin < 0.0) tmin = 0.0;
        }
        if (range <= tmin) {
            if (is_inf(range)) { status = infeasible ? ORC_UNDEF : ORC_UNBND; break; }
            /* bound flip */
            for (int i = 0; i < m; ++i) lp->beta[i] += s * range * lp->T[(size_t)i * n + q];
            lp->stat[q] = (lp->stat[q] == AT_LO) ? AT_UP : AT_LO;
            lp->flips++;
            degen_run = 0; bland = 0;
            continue;
        }
        if (p < 0) { status = infeasible ? ORC_UNDEF : ORC_UNBND; break; }
        /* ---- basis change ---- */
        double xq_new = nb_value(lp, q) + s * tmin;
        for (int i = 0; i < m; ++i)
            if (i != p) lp->beta[i] += s * tmin * lp->T[(size_t)i * n + q];
        int vleave = lp->head[p];
        pivot(lp, p, q, d_valid ? lp->d : NULL);
        lp->beta[p] = xq_new;
        if (lp->lo[vleave] == lp->up[vleave]) lp->stat[q] = AT_FIX;
        else lp->stat[q] = p_to_upper ? AT_UP : AT_LO;
        lp->pivots++;
        if (tmin <= TOL_TIE) { if (++degen_run > 60) bland = 1; }
        else { degen_run = 0; bland = 0; }
    }
    if (status == ORC_OPT) {
        extract_x(lp, x);
        double z = 0.0;
        for (int j = 0; j < n; ++j) z += cost[j] * x[j];
        if (objval) *objval = z;
    } else {
        extract_x(lp, x);
        if (objval) { double z = 0.0; for (int j = 0; j < n; ++j) z += cost[j] * x[j]; *objval = z; }
    }
    return status;
}
