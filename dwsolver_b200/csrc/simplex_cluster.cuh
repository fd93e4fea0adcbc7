// simplex_cluster.cuh — thread-block-CLUSTER-per-block simplex, tableau resident in HBM (config C class).
//
// Blocks whose working vectors alone fill a CTA's shared memory (2048 x 4096: T = 64 MB) are solved
// by one thread-block cluster of G CTAs (G = 2..16, one CTA per SM):
//   * T stays in HBM and is striped BY ROWS over the G CTAs; a basis change is a rank-1 update in
//     which every CTA streams only its own rows (LDG.128 -> FMA -> STG.128, .cg, 8 x 16 B per lane in
//     flight, rows whose pivot-column entry is exactly zero are skipped);
//   * everything else — beta, reduced-cost row d, head/nbv, bounds, column status — is REPLICATED in
//     every CTA's shared memory and updated redundantly with identical arithmetic, so pricing, the
//     Harris ratio test and all bookkeeping need no communication at all and every CTA takes the same
//     decision bit for bit;
//   * the two things a CTA cannot get locally cross the cluster through distributed shared memory:
//     the pivot column (each CTA reads its rows' entries and stores them into every peer's colq,
//     double-buffered) and column sums over all rows (partials in local smem, summed in rank order
//     by every CTA from its peers' smem);
//   * the pivot row is read by every CTA straight from HBM/L2 after cluster barrier #1 and is only
//     overwritten by its owner after barrier #2 (split arrive/wait: the sweep of the other rows runs
//     between the two halves), so a basis change costs two cluster barriers (~0.2 us each);
//   * phase 1 keeps its price row incrementally too (same rank-1 update with frozen weights, then
//     one row read per basic variable whose feasibility status changed) instead of re-reading every
//     infeasible row of T per iteration.
// Algorithm, tolerances and proposal epilogue are those of simplex_smem.cuh.  Reference being
// replaced: the warm-started glp_simplex call per block (/root/reference/src/dw_subprob.c:329) with
// the pricing / column code around it (:291-321, :453-540).
#pragma once
#include <cooperative_groups.h>

#include "simplex_smem.cuh"

namespace dwk {

namespace cg = cooperative_groups;

constexpr int NT_CL = 512;

__host__ __device__ inline size_t cl_smem_bytes(int m, int n)
{
    const size_t ld = tab_ld(n, true);
    return 2 * align16(sizeof(double) * ld)                  // drow, rowp
         + 2 * align16(sizeof(double) * (size_t)(m + 1))     // colq x 2
         + 4 * align16(sizeof(double) * (size_t)m)           // beta, w, rlo, rup
         + align16(sizeof(int) * (size_t)m) + align16(sizeof(int) * (size_t)n)   // head, nbv
         + align16((size_t)n)                                // cstat
         + align16(sizeof(unsigned short) * (size_t)(m + 1)) // arow
         + align16(sizeof(float) * (size_t)n);               // devex reference weights of the columns
}
// per block and rank: priced costs cs[n]
__host__ __device__ inline size_t cl_ws_doubles(int n, int G) { return (size_t)G * (size_t)((n + 1) & ~1); }

__device__ __forceinline__ void cl_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cl_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cl_sync() { cl_arrive(); cl_wait(); }

// progress markers (DWGPU_DEBUG): slot 8*rank + t of the block's 32-entry record, written by thread 0
#define CL_MARK(t, v) do { if (P.dbg && tid == 0) { volatile long long *d_ = P.dbg + 32 * (size_t)k + 8 * (rank & 3) + (t); *d_ = (long long)(v); } } while (0)

struct Cl {
    int m, n, ld, G, rank, r0, r1;
    double *T;
    double *drow, *rowp, *colq[2], *beta, *w, *rlo, *rup;
    int *head, *nbv;
    signed char *cstat;
    unsigned short *arow;
    float *dvx;
};

struct ClShared {
    Scratch<NT_CL> scr;
    int na;          // own active rows of the current column
    int nchg;        // rows whose phase-1 weight changed
    int flag;
};

// ---- column q of the tableau into every CTA's colq[buf]; own non-zero rows into arow -------------
__device__ __forceinline__ void cl_exchange_column(Cl &c, ClShared &sh, int q, int buf)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x;
    if (tid == 0) sh.na = 0;
    __syncthreads();
    for (int i = c.r0 + tid; i < c.r1; i += NT_CL) {
        const double v = __ldcg(c.T + (size_t)i * c.ld + q);
        for (int r = 0; r < c.G; ++r) cluster.map_shared_rank(c.colq[buf], r)[i] = v;
        if (v != 0.0) c.arow[atomicAdd(&sh.na, 1)] = (unsigned short)i;
    }
    cl_sync();
}

// ---- rowp <- T[p][:] * inv (every CTA, from HBM/L2) --------------------------------------------------
__device__ __forceinline__ void cl_stage_row(Cl &c, int p, double inv)
{
    const double2 *prow = reinterpret_cast<const double2 *>(c.T + (size_t)p * c.ld);
    double2 *rp2 = reinterpret_cast<double2 *>(c.rowp);
    const int n2 = c.ld >> 1;
    for (int j0 = 0; j0 < n2; j0 += NT_CL * 4) {
        double2 v[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int j = j0 + t * NT_CL + threadIdx.x;
            if (j < n2) v[t] = __ldcg(prow + j);
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int j = j0 + t * NT_CL + threadIdx.x;
            if (j < n2) { v[t].x *= inv; v[t].y *= inv; rp2[j] = v[t]; }
        }
    }
}

// ---- rank-1 update of the own active rows except p; one row per warp at a time ----------------------
__device__ __forceinline__ void cl_sweep(Cl &c, const double *colq, int na, int p, int q, double inv)
{
    constexpr int NW = NT_CL / 32, SB = 8;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n2 = c.ld >> 1, n2a = (c.n + 1) >> 1;
    const double2 *rowp2 = reinterpret_cast<const double2 *>(c.rowp);
    for (int a = warp; a < na; a += NW) {
        const int i = c.arow[a];
        if (i == p) continue;
        const double ci = colq[i];
        const double cq = ci * inv;
        double2 *row = reinterpret_cast<double2 *>(c.T) + (size_t)i * n2;
        for (int j0 = 0; j0 < n2a; j0 += 32 * SB) {
            double2 v[SB];
#pragma unroll
            for (int t = 0; t < SB; ++t) {
                const int j = j0 + t * 32 + lane;
                if (j < n2a) v[t] = __ldcg(row + j);
            }
#pragma unroll
            for (int t = 0; t < SB; ++t) {
                const int j = j0 + t * 32 + lane;
                if (j < n2a) {
                    const double2 ra = rowp2[j];
                    const int jq = q - 2 * j;
                    double2 r;
                    r.x = fma(-ci, ra.x, v[t].x);
                    r.y = fma(-ci, ra.y, v[t].y);
                    if (jq == 0) r.x = cq;
                    if (jq == 1) r.y = cq;
                    __stcg(row + j, r);
                }
            }
        }
    }
}

// ---- owner of row p: T[p][:] <- -rowp, T[p][q] <- inv ----------------------------------------------
__device__ __forceinline__ void cl_write_pivot_row(Cl &c, int p, int q, double inv)
{
    if (p < c.r0 || p >= c.r1) return;
    double2 *row = reinterpret_cast<double2 *>(c.T) + (size_t)p * (c.ld >> 1);
    const double2 *rowp2 = reinterpret_cast<const double2 *>(c.rowp);
    const int n2a = (c.n + 1) >> 1;
    for (int j = threadIdx.x; j < n2a; j += NT_CL) {
        const double2 ra = rowp2[j];
        const int jq = q - 2 * j;
        double2 r;
        r.x = -ra.x; r.y = -ra.y;
        if (jq == 0) r.x = inv;
        if (jq == 1) r.y = inv;
        __stcg(row + j, r);
    }
}

// ---- out[j] = base[j] + sum over ALL rows i of wt[i] * T[i][j]  (wt replicated, block-uniform) --------
// Partials over the own rows go to the local rowp; every CTA then adds its peers' partials in rank
// order (identical result everywhere).  Two cluster barriers.  base_of(j) supplies the c_N part.
template <typename BaseFn>
__device__ __forceinline__ void cl_colsum(Cl &c, ClShared &sh, const double *wt, BaseFn base_of)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x;
    const int n2 = c.ld >> 1;
    // own rows with a non-zero weight
    if (tid == 0) sh.na = 0;
    __syncthreads();
    for (int i = c.r0 + tid; i < c.r1; i += NT_CL)
        if (wt[i] != 0.0) c.arow[atomicAdd(&sh.na, 1)] = (unsigned short)i;
    __syncthreads();
    const int na = sh.na;
    const double2 *T2 = reinterpret_cast<const double2 *>(c.T);
    double2 *part = reinterpret_cast<double2 *>(c.rowp);
    for (int j0 = 0; j0 < n2; j0 += NT_CL * 2) {
        const int ja = j0 + tid, jb = j0 + NT_CL + tid;
        const bool oa = ja < n2, ob = jb < n2;
        double2 sa = make_double2(0.0, 0.0), sb = make_double2(0.0, 0.0);
        int a = 0;
        for (; a + 4 <= na; a += 4) {        // 4 rows x 2 double2 per thread in flight
            double2 va[4], vb[4];
            double wv[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int i = c.arow[a + t];
                wv[t] = wt[i];
                if (oa) va[t] = __ldcg(T2 + (size_t)i * n2 + ja);
                if (ob) vb[t] = __ldcg(T2 + (size_t)i * n2 + jb);
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                if (oa) { sa.x = fma(wv[t], va[t].x, sa.x); sa.y = fma(wv[t], va[t].y, sa.y); }
                if (ob) { sb.x = fma(wv[t], vb[t].x, sb.x); sb.y = fma(wv[t], vb[t].y, sb.y); }
            }
        }
        for (; a < na; ++a) {
            const int i = c.arow[a];
            const double wv = wt[i];
            if (oa) { const double2 v = __ldcg(T2 + (size_t)i * n2 + ja); sa.x = fma(wv, v.x, sa.x); sa.y = fma(wv, v.y, sa.y); }
            if (ob) { const double2 v = __ldcg(T2 + (size_t)i * n2 + jb); sb.x = fma(wv, v.x, sb.x); sb.y = fma(wv, v.y, sb.y); }
        }
        if (oa) part[ja] = sa;
        if (ob) part[jb] = sb;
    }
    cl_sync();                                   // every partial is complete
    for (int j = tid; j < c.ld; j += NT_CL) {
        double s = j < c.n ? base_of(j) : 0.0;
        for (int r = 0; r < c.G; ++r) s += cluster.map_shared_rank(c.rowp, r)[j];
        c.drow[j] = s;
    }
    cl_sync();                                   // nobody still reads a peer's rowp
}

// ---- beta = T x_N: every CTA does its own rows (one warp per row) and stores the result into every
// peer's beta through DSMEM.  rowp is used as scratch for x_N.  Ends with a cluster barrier.
__device__ __forceinline__ void cl_recompute_beta(Cl &c, const BlockDev &B)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int j = tid; j < c.n; j += NT_CL) {
        const int v = c.nbv[j];
        c.rowp[j] = nb_value(c.cstat[j], __ldg(B.lo + v), __ldg(B.up + v));
    }
    __syncthreads();
    const double2 *x2 = reinterpret_cast<const double2 *>(c.rowp);
    const int n2a = c.n >> 1;
    for (int i = c.r0 + warp; i < c.r1; i += NT_CL / 32) {
        const double2 *row = reinterpret_cast<const double2 *>(c.T + (size_t)i * c.ld);
        double s = 0.0;
        for (int j0 = 0; j0 < n2a; j0 += 32 * 8) {
            double2 v[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) { const int j = j0 + t * 32 + lane; if (j < n2a) v[t] = __ldcg(row + j); }
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const int j = j0 + t * 32 + lane;
                if (j < n2a) { const double2 x = x2[j]; s = fma(v[t].x, x.x, s); s = fma(v[t].y, x.y, s); }
            }
        }
        if ((c.n & 1) && lane == 0) s = fma(__ldcg(c.T + (size_t)i * c.ld + c.n - 1), c.rowp[c.n - 1], s);
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0)
            for (int r = 0; r < c.G; ++r) cluster.map_shared_rank(c.beta, r)[i] = s;
    }
    cl_sync();
}

// exact ratio of row i for a step of the entering variable in direction s; +inf when it does not block
template <bool INF>
__device__ __forceinline__ void cl_row_ratio(const Cl &c, int i, double s, double tiq, double tol_bnd, double tol_piv,
                                             double harris, bool bland, double &tex, double &tharris)
{
    tex = CUDART_INF; tharris = CUDART_INF;
    const double a = s * tiq;
    double lo = c.rlo[i], up = c.rup[i];
    const double b = c.beta[i];
    if (INF) {
        if (b < lo - tol_bnd * (1.0 + fabs(lo))) { up = lo; lo = -CUDART_INF; }
        else if (b > up + tol_bnd * (1.0 + fabs(up))) { lo = up; up = CUDART_INF; }
    }
    bool blocks = false; double slack = 0.0, target = 0.0;
    if (a > tol_piv) { if (up < BIG) { blocks = true; slack = up - b; target = up; } }
    else if (a < -tol_piv) { if (lo > -BIG) { blocks = true; slack = b - lo; target = lo; } }
    if (blocks) {
        const double r = __drcp_rn(fabs(a));
        tex = slack * r;
        const double delta = bland ? 0.0 : harris * (1.0 + fabs(target));
        tharris = fmax((slack + delta) * r, 0.0);
    }
}

__device__ __forceinline__ double cl_phase1_weight(double b, double lo, double up, double tol_bnd)
{
    if (b < lo - tol_bnd * (1.0 + fabs(lo))) return -1.0;
    if (b > up + tol_bnd * (1.0 + fabs(up))) return 1.0;
    return 0.0;
}

template <int NT>
__global__ void __launch_bounds__(NT, 1) solve_cluster_kernel(const BlockDev *blocks, const int *list, int count,
                                                              SolveParams P, int G)
{
    static_assert(NT == NT_CL, "the cluster kernel is written for NT_CL threads");
    extern __shared__ __align__(16) char smem_raw[];
    __shared__ ClShared sh;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int slot = (int)blockIdx.x / G;
    // every CTA of a cluster takes the same branch here, so no peer is left waiting at a barrier
    if (slot >= count) return;
    const int k = list[slot];
    const BlockDev &B = blocks[k];       // fields are fetched where they are used (read-only, L1-cached)
    const int m = B.m, n = B.n, L = B.L, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double tol_dj = P.tol_dj, tol_bnd = P.tol_bnd, tol_piv = P.tol_piv;
    const double harris = 0.5 * tol_bnd;

    Cl c;
    c.m = m; c.n = n; c.ld = B.ldg; c.G = G; c.rank = rank; c.T = B.T;
    const int ld = c.ld;
    {
        const int rpc = (m + G - 1) / G;
        c.r0 = min(m, rank * rpc); c.r1 = min(m, c.r0 + rpc);
    }
    char *sp = smem_raw;
    c.drow = (double *)sp;        sp += align16(sizeof(double) * (size_t)ld);
    c.rowp = (double *)sp;        sp += align16(sizeof(double) * (size_t)ld);
    c.colq[0] = (double *)sp;     sp += align16(sizeof(double) * (size_t)(m + 1));
    c.colq[1] = (double *)sp;     sp += align16(sizeof(double) * (size_t)(m + 1));
    c.beta = (double *)sp;        sp += align16(sizeof(double) * (size_t)m);
    c.w = (double *)sp;           sp += align16(sizeof(double) * (size_t)m);
    c.rlo = (double *)sp;         sp += align16(sizeof(double) * (size_t)m);
    c.rup = (double *)sp;         sp += align16(sizeof(double) * (size_t)m);
    c.head = (int *)sp;           sp += align16(sizeof(int) * (size_t)m);
    c.nbv = (int *)sp;            sp += align16(sizeof(int) * (size_t)n);
    c.cstat = (signed char *)sp;  sp += align16((size_t)n);
    c.arow = (unsigned short *)sp; sp += align16(sizeof(unsigned short) * (size_t)(m + 1));
    c.dvx = (float *)sp;
    double *cs = B.ws + (size_t)rank * (size_t)((n + 1) & ~1);     // this CTA's copy of the priced costs

    // ---- state in (replicated) + priced objective (same arithmetic as the reference, dw_subprob.c:291-321)
    for (int i = tid; i < m; i += NT) {
        const int v = B.head[i];
        c.head[i] = v; c.beta[i] = B.beta[i]; c.w[i] = 0.0;
        c.rlo[i] = __ldg(B.lo + v); c.rup[i] = __ldg(B.up + v);
    }
    for (int j = tid; j < n; j += NT) {
        const int v = B.nbv[j];
        c.nbv[j] = v;
        c.cstat[j] = B.vstat[v];
        c.dvx[j] = 1.0f;
        double cj;
        if (P.initial) cj = B.c_file[j];
        else if (P.cs_pre != nullptr && B.Dt != nullptr) cj = __ldg(P.cs_pre + B.xoff + j);     // priced by the multi-vector GEMM
        else {
            double y = 0.0;
            for (int e = B.dc_ptr[j]; e < B.dc_ptr[j + 1]; ++e)
                y = __dadd_rn(y, __dmul_rn(__ldg(P.pi + B.dc_row[e]), B.dc_val[e]));
            const double base = P.phase_one ? 0.0 : B.c_master[j];
            cj = __dadd_rn(__dmul_rn(-1.0, y), base);
        }
        cs[j] = cj;
    }
    cl_sync();          // all CTAs of the cluster are resident and initialised before any DSMEM traffic
    CL_MARK(0, 1);

    int status = 0;
    long long iter = 0;
    const long long max_iter = (long long)P.max_iter_mult * (m + n) + 1000;
    unsigned long long n_piv = 0, n_flip = 0, n_reb = 0, n_p1 = 0, n_rows = 0;
    unsigned long long n_price = 0;     // tableau doubles read for price rows / beta (replicated count)
    bool need_check = true, infeasible = false, d_valid = false, p1_valid = false, p1_fresh = false, bland = false;
    bool changed = false;
    int degen = 0, buf = 0;
    const int degen_limit = 100 + (m + n) / 2;
    int p1_leave_col = -1; double p1_leave_w = 0.0;      // column / frozen weight of the variable that just left

    // ---- dual simplex start ----------------------------------------------------------------------------
    // A primal-infeasible start (the cold solve: all-logical basis) is repaired by the DUAL simplex when
    // every nonbasic column can be put on the bound that makes its reduced cost sign-correct (boxed
    // columns, e.g. 0 <= x <= u): leaving row by dual devex (infeasibility^2 / reference weight), entering
    // column by a two-pass Harris dual ratio test on the pivot row, then the same rank-1 update as a
    // primal step.  On config C class blocks this needs an order of magnitude fewer basis changes than
    // primal phase 1 + phase 2 with Dantzig pricing.  Anything it cannot handle (free / one-sided
    // columns with the wrong sign, dual unboundedness, too many iterations) falls through to the primal
    // loop below, which also certifies the final basis (pricing, residual check, rebuild).
    if (m > 0 && !P.no_dual_start) {
        int viol0 = 0;
        for (int i = tid; i < m; i += NT)
            viol0 |= (cl_phase1_weight(c.beta[i], c.rlo[i], c.rup[i], tol_bnd) != 0.0);
        if (__syncthreads_or(viol0)) {
            // reduced costs of the current basis
            for (int i = tid; i < m; i += NT) {
                const int v = c.head[i];
                c.w[i] = v >= m ? cs[v - m] : 0.0;
            }
            __syncthreads();
            cl_colsum(c, sh, c.w, [&](int j) { const int v = c.nbv[j]; return v >= m ? cs[v - m] : 0.0; });
            // sign-correct bounds; count flips and impossible columns
            int bad = 0, nfl = 0;
            for (int j = tid; j < n; j += NT) {
                const double dj = c.drow[j];
                const signed char st = c.cstat[j];
                const int v = c.nbv[j];
                if (st == ST_LO && dj < -tol_dj) {
                    if (__ldg(B.up + v) < BIG) { c.cstat[j] = ST_UP; ++nfl; } else bad = 1;
                } else if (st == ST_UP && dj > tol_dj) {
                    if (__ldg(B.lo + v) > -BIG) { c.cstat[j] = ST_LO; ++nfl; } else bad = 1;
                } else if (st == ST_FREE && fabs(dj) > tol_dj) bad = 1;
            }
            const int any_bad = __syncthreads_or(bad);
            const int any_flip = __syncthreads_or(nfl);
            if (any_flip) {
                n_flip += (unsigned long long)block_sum<NT>((double)nfl, sh.scr);
                changed = true;
                cl_recompute_beta(c, B);
                n_price += (unsigned long long)m * (unsigned long long)n;
            }
            for (int i = tid; i < m; i += NT) c.w[i] = 1.0;       // dual devex reference weights
            __syncthreads();
            d_valid = true;
            const long long dual_limit = 20LL * (m + n) + 1000;
            long long dit = 0;
            CL_MARK(0, 2);
            while (!any_bad) {
                if (++dit > dual_limit) break;
                CL_MARK(1, dit);
                // ---- leaving row: largest infeasibility^2 / weight (replicated)
                double key = 0.0; int p = INT_MAX;
                for (int i = tid; i < m; i += NT) {
                    const double b = c.beta[i], lo = c.rlo[i], up = c.rup[i];
                    double v = 0.0;
                    if (b < lo - tol_bnd * (1.0 + fabs(lo))) v = lo - b;
                    else if (b > up + tol_bnd * (1.0 + fabs(up))) v = b - up;
                    if (v > 0.0) {
                        const double kk = v * v / c.w[i];
                        if (kk > key) { key = kk; p = i; }
                    }
                }
                if (key <= 0.0) p = INT_MAX;
                block_argmax<NT>(key, p, sh.scr);
                CL_MARK(2, p);
                if (p == INT_MAX) break;                          // primal feasible: done
                // everything read here is rewritten further down: read it before the barriers below
                const double bp = c.beta[p], lop = c.rlo[p], upp = c.rup[p], wp = c.w[p];
                const bool below = bp < lop;
                const double target = below ? lop : upp;
                const double sg = below ? 1.0 : -1.0;
                cl_sync();                                        // every sweep of the previous step has landed
                cl_stage_row(c, p, 1.0);                          // alpha = T[p][:], unscaled
                __syncthreads();
                // ---- entering column: dual ratio test, two passes (Harris)
                double rmin = CUDART_INF;
                for (int j = tid; j < n; j += NT) {
                    const double a = sg * c.rowp[j];
                    const signed char st = c.cstat[j];
                    const bool cand = (st == ST_LO && a > tol_piv) || (st == ST_UP && a < -tol_piv) ||
                                      (st == ST_FREE && fabs(a) > tol_piv);
                    if (cand) rmin = fmin(rmin, (fabs(c.drow[j]) + tol_dj) / fabs(a));
                }
                rmin = block_min<NT>(rmin, sh.scr);
                if (!(rmin < CUDART_INF)) break;                  // dual unbounded: let the primal loop decide
                double pk = -1.0; int q = INT_MAX;
                for (int j = tid; j < n; j += NT) {
                    const double a = sg * c.rowp[j];
                    const signed char st = c.cstat[j];
                    const bool cand = (st == ST_LO && a > tol_piv) || (st == ST_UP && a < -tol_piv) ||
                                      (st == ST_FREE && fabs(a) > tol_piv);
                    if (cand && fabs(c.drow[j]) / fabs(a) <= rmin && fabs(a) > pk) { pk = fabs(a); q = j; }
                }
                block_argmax<NT>(pk, q, sh.scr);
                CL_MARK(3, q);
                if (q == INT_MAX) break;
                // ---- the step: same update as a primal basis change on (p, q)
                const signed char stq = c.cstat[q];
                const int vq = c.nbv[q];
                const double loq = __ldg(B.lo + vq), upq = __ldg(B.up + vq);
                const double piv = c.rowp[q];
                const double inv = __drcp_rn(piv);
                const double dq = c.drow[q];
                const int vleave = c.head[p];
                cl_exchange_column(c, sh, q, buf);                // barrier: everybody has read row p as well
                const double *colq = c.colq[buf];
                const int na_own = sh.na;
                buf ^= 1;
                const double dxq = (target - bp) * inv;
                const double xq_new = nb_value(stq, loq, upq) + dxq;
                unsigned nz_local = 0;
                int wbig = 0;
                for (int i = tid; i < m; i += NT) {
                    const double ci = colq[i];
                    nz_local += (ci != 0.0);
                    if (i == p) { const double wn = fmax(wp * inv * inv, 1.0); c.beta[i] = xq_new; c.w[i] = wn; wbig |= wn > 1e12; }
                    else if (ci != 0.0) {
                        c.beta[i] = fma(dxq, ci, c.beta[i]);
                        const double g = ci * inv;
                        const double wn = fmax(c.w[i], g * g * wp);
                        c.w[i] = wn; wbig |= wn > 1e12;
                    }
                }
                if (__syncthreads_or(wbig))                       // devex: new reference framework
                    for (int i = tid; i < m; i += NT) c.w[i] = 1.0;
                {
                    double2 *rp2 = reinterpret_cast<double2 *>(c.rowp);
                    double2 *d2 = reinterpret_cast<double2 *>(c.drow);
                    const double cqd = dq * inv;
                    for (int j = tid; j < (ld >> 1); j += NT) {
                        double2 ra = rp2[j];
                        ra.x *= inv; ra.y *= inv;
                        rp2[j] = ra;
                        if (dq != 0.0 && j < ((n + 1) >> 1)) {
                            double2 r = d2[j];
                            const int jq = q - 2 * j;
                            r.x = fma(-dq, ra.x, r.x);
                            r.y = fma(-dq, ra.y, r.y);
                            if (jq == 0) r.x = cqd;
                            if (jq == 1) r.y = cqd;
                            d2[j] = r;
                        }
                    }
                }
                __syncthreads();
                if (tid == 0) {
                    c.rlo[p] = loq; c.rup[p] = upq;
                    c.cstat[q] = (lop == upp) ? ST_FIX : (below ? ST_LO : ST_UP);
                    c.head[p] = vq; c.nbv[q] = vleave;
                }
                cl_sweep(c, colq, na_own, p, q, inv);
                cl_write_pivot_row(c, p, q, inv);
                ++n_piv;
                n_rows += nz_local;
                changed = true;
                __syncthreads();
            }
            CL_MARK(0, 4);
            cl_sync();                                            // sweeps complete before the primal loop reads T
            for (int i = tid; i < m; i += NT) c.w[i] = 0.0;       // w is the phase-1 weight array again
            __syncthreads();
        }
    }

    for (;;) {
        if (++iter > max_iter) { status = 1; break; }
        CL_MARK(0, 5); CL_MARK(4, iter);
        // ---- A. feasibility scan (replicated); in phase 1 also the list of rows whose weight changed
        if (need_check || infeasible) {
            if (tid == 0) sh.nchg = 0;
            __syncthreads();
            int viol = 0;
            for (int i = tid; i < m; i += NT) {
                const double wi = cl_phase1_weight(c.beta[i], c.rlo[i], c.rup[i], tol_bnd);
                viol |= (wi != 0.0);
                if (p1_valid) {
                    const double dw = wi - c.w[i];
                    if (dw != 0.0) {
                        const int s_ = atomicAdd(&sh.nchg, 1);
                        c.arow[s_] = (unsigned short)i;
                        c.colq[buf ^ 1][s_] = dw;      // the idle column buffer doubles as the delta list
                    }
                }
                c.w[i] = wi;
            }
            const bool was = infeasible;
            infeasible = __syncthreads_or(viol) != 0;
            need_check = false;
            if (infeasible) d_valid = false;
            if (!infeasible && was) p1_valid = false;
        }
        if (infeasible) {
            if (!p1_valid) {
                // phase-1 prices from scratch: d1 = T' w
                cl_colsum(c, sh, c.w, [](int) { return 0.0; });
                {
                    unsigned cnt = 0;
                    for (int i = 0; i < m; ++i) cnt += c.w[i] != 0.0;
                    n_price += (unsigned long long)cnt * (unsigned long long)n;
                }
                p1_valid = true; p1_fresh = true;
            } else {
                // incremental: the rank-1 update already ran with frozen weights; take out the cost the
                // leaving variable carried, then add dw_i * T[i][:] for every row whose status changed
                const int nchg = sh.nchg;
                if (p1_leave_col >= 0 && tid == 0) c.drow[p1_leave_col] -= p1_leave_w;
                if (nchg > 0) {
                    cl_sync();                                   // the sweeps of every CTA are complete
                    // deterministic order: sort-free — each thread owns columns, loops rows in list order;
                    // the list order comes from atomics, so sort the (short) list by row first
                    if (tid == 0) {
                        for (int a = 1; a < nchg; ++a) {         // insertion sort, nchg is tiny
                            const unsigned short ri = c.arow[a]; const double dv = c.colq[buf ^ 1][a];
                            int b = a - 1;
                            while (b >= 0 && c.arow[b] > ri) { c.arow[b + 1] = c.arow[b]; c.colq[buf ^ 1][b + 1] = c.colq[buf ^ 1][b]; --b; }
                            c.arow[b + 1] = ri; c.colq[buf ^ 1][b + 1] = dv;
                        }
                    }
                    __syncthreads();
                    const int n2 = ld >> 1;
                    const double2 *T2 = reinterpret_cast<const double2 *>(c.T);
                    double2 *d2 = reinterpret_cast<double2 *>(c.drow);
                    for (int j = tid; j < n2; j += NT) {
                        double2 s = d2[j];
                        for (int a = 0; a < nchg; ++a) {
                            const double dv = c.colq[buf ^ 1][a];
                            const double2 v = __ldcg(T2 + (size_t)c.arow[a] * n2 + j);
                            s.x = fma(dv, v.x, s.x); s.y = fma(dv, v.y, s.y);
                        }
                        d2[j] = s;
                    }
                    p1_fresh = false;
                }
                __syncthreads();
            }
            p1_leave_col = -1;
            ++n_p1;
        } else if (!d_valid) {
            // reduced costs d = c_N + T' c_B
            for (int i = tid; i < m; i += NT) {
                const int v = c.head[i];
                c.w[i] = v >= m ? cs[v - m] : 0.0;
            }
            __syncthreads();
            cl_colsum(c, sh, c.w, [&](int j) { const int v = c.nbv[j]; return v >= m ? cs[v - m] : 0.0; });
            for (int i = tid; i < m; i += NT) c.w[i] = 0.0;     // w is the phase-1 weight array again
            d_valid = true;
            __syncthreads();
        }
        // ---- B. pricing (replicated): Dantzig, or Bland after a run of degenerate steps
        double key = -CUDART_INF; int q = INT_MAX;
        for (int j = tid; j < n; j += NT) {
            const double dj = c.drow[j];
            const signed char st = c.cstat[j];
            const bool ok = (st == ST_LO && dj < -tol_dj) || (st == ST_UP && dj > tol_dj) ||
                            (st == ST_FREE && fabs(dj) > tol_dj);
            if (ok) {
                // devex pricing: d_j^2 over the reference weight (Bland: lowest variable id)
                const double kj = bland ? -(double)c.nbv[j] : dj * dj / (double)c.dvx[j];
                if (kj > key) { key = kj; q = j; }
            }
        }
        block_argmax<NT>(key, q, sh.scr);
        if (q == INT_MAX) {
            if (infeasible) {
                if (!p1_fresh) { p1_valid = false; continue; }     // confirm with prices computed from scratch
                status = 4; break;
            }
            // ---- claimed optimal: x into rowp, auxiliary values into colq[buf]
            double *aux = c.colq[buf];
            for (int j = tid; j < n; j += NT) {
                const int v = c.nbv[j];
                const double xv = nb_value(c.cstat[j], __ldg(B.lo + v), __ldg(B.up + v));
                if (v >= m) c.rowp[v - m] = xv; else aux[v] = xv;
            }
            for (int i = tid; i < m; i += NT) {
                const int v = c.head[i];
                if (v >= m) c.rowp[v - m] = c.beta[i]; else aux[v] = c.beta[i];
            }
            __syncthreads();
            if (!changed && n_reb == 0 && !P.initial && !P.force_full) { status = 5; break; }
            double worst = 0.0;
            for (int i = tid; i < m; i += NT) {
                double s = 0.0;
                for (int e = B.a_rp[i]; e < B.a_rp[i + 1]; ++e) s = fma(B.a_v[e], c.rowp[B.a_ci[e]], s);
                worst = fmax(worst, fabs(s - aux[i]) / (1.0 + fabs(s)));
            }
            worst = block_max<NT>(worst, sh.scr);
            if (worst > 0.1 * tol_bnd && n_reb < 3) {
                // ---- rebuild T from A_k for the current basis (Gauss-Jordan with the cluster sweeps)
                ++n_reb;
                CL_MARK(0, 6); CL_MARK(5, n_reb);
                // target basis by variable id -> global vstat (rank 0), visible to all after the barrier
                if (rank == 0) {
                    for (int i = tid; i < m; i += NT) B.vstat[c.head[i]] = ST_BASIC;
                    for (int j = tid; j < n; j += NT) B.vstat[c.nbv[j]] = c.cstat[j];
                }
                // own rows of T <- A_k
                {
                    const int n2 = ld >> 1;
                    double2 *T2 = reinterpret_cast<double2 *>(c.T);
                    const double2 z = make_double2(0.0, 0.0);
                    for (size_t e = (size_t)c.r0 * n2 + tid; e < (size_t)c.r1 * n2; e += NT) __stcg(T2 + e, z);
                    __syncthreads();
                    for (int i = c.r0 + tid; i < c.r1; i += NT)
                        for (int e = B.a_rp[i]; e < B.a_rp[i + 1]; ++e) {
                            double *t = c.T + (size_t)i * ld + B.a_ci[e];
                            __stcg(t, __ldcg(t) + B.a_v[e]);
                        }
                }
                for (int i = tid; i < m; i += NT) c.head[i] = i;
                for (int j = tid; j < n; j += NT) c.nbv[j] = m + j;
                cl_sync();
                bool singular = false;
                for (int j = 0; j < n; ++j) {
                    if (__ldcg((const char *)B.vstat + m + j) != ST_BASIC) continue;     // cluster-uniform
                    cl_exchange_column(c, sh, j, buf);
                    const double *colq = c.colq[buf];
                    double pk = -1.0; int pr = INT_MAX;
                    for (int i = tid; i < m; i += NT) {
                        const int h = c.head[i];
                        if (h < m && __ldcg((const char *)B.vstat + h) != ST_BASIC) {
                            const double a = fabs(colq[i]);
                            if (a > pk) { pk = a; pr = i; }
                        }
                    }
                    block_argmax<NT>(pk, pr, sh.scr);
                    if (pr == INT_MAX || pk < 1e-11) { singular = true; break; }
                    const double inv = 1.0 / colq[pr];
                    cl_stage_row(c, pr, inv);
                    __syncthreads();
                    cl_arrive();
                    cl_sweep(c, colq, sh.na, pr, j, inv);
                    cl_wait();
                    cl_write_pivot_row(c, pr, j, inv);
                    if (tid == 0) { const int t = c.head[pr]; c.head[pr] = c.nbv[j]; c.nbv[j] = t; }
                    buf ^= 1;
                    __syncthreads();
                }
                if (singular) { status = 1; break; }
                cl_sync();
                // statuses, bounds, beta = T x_N (own rows, broadcast through DSMEM)
                for (int j = tid; j < n; j += NT) c.cstat[j] = __ldcg((const char *)B.vstat + c.nbv[j]);
                for (int i = tid; i < m; i += NT) {
                    const int v = c.head[i];
                    c.rlo[i] = __ldg(B.lo + v); c.rup[i] = __ldg(B.up + v);
                }
                __syncthreads();
                cl_recompute_beta(c, B);
                for (int i = tid; i < m; i += NT) c.w[i] = 0.0;
                need_check = true; d_valid = false; p1_valid = false;
                __syncthreads();
                continue;
            }
            status = 5;
            break;
        }
        // ---- C. ratio test (replicated) on the exchanged column
        const double dq = c.drow[q];
        const double s = dq < 0.0 ? 1.0 : -1.0;
        const signed char stq = c.cstat[q];
        const int vq = c.nbv[q];
        const double loq = __ldg(B.lo + vq), upq = __ldg(B.up + vq);
        cl_exchange_column(c, sh, q, buf);
        const double *colq = c.colq[buf];
        const int na_own = sh.na;
        buf ^= 1;
        double tmax = CUDART_INF;
        unsigned nz_local = 0;
        for (int i = tid; i < m; i += NT) {
            const double tiq = colq[i];
            nz_local += (tiq != 0.0);
            double tex, th;
            if (infeasible) cl_row_ratio<true>(c, i, s, tiq, tol_bnd, tol_piv, harris, bland, tex, th);
            else cl_row_ratio<false>(c, i, s, tiq, tol_bnd, tol_piv, harris, bland, tex, th);
            tmax = fmin(tmax, th);
        }
        tmax = block_min<NT>(tmax, sh.scr);
        double pk = -CUDART_INF; int p = INT_MAX;
        if (tmax < CUDART_INF) {
            const double lim = bland ? tmax + 1e-12 * (1.0 + tmax) : tmax;
            for (int i = tid; i < m; i += NT) {
                double tex, th;
                if (infeasible) cl_row_ratio<true>(c, i, s, colq[i], tol_bnd, tol_piv, harris, bland, tex, th);
                else cl_row_ratio<false>(c, i, s, colq[i], tol_bnd, tol_piv, harris, bland, tex, th);
                if (tex <= lim) {
                    const double kk = bland ? -(double)c.head[i] : fabs(colq[i]);
                    if (kk > pk) { pk = kk; p = i; }
                }
            }
            block_argmax<NT>(pk, p, sh.scr);
        }
        double theta = CUDART_INF;
        if (p != INT_MAX) {
            double tex, th;
            if (infeasible) cl_row_ratio<true>(c, p, s, colq[p], tol_bnd, tol_piv, harris, bland, tex, th);
            else cl_row_ratio<false>(c, p, s, colq[p], tol_bnd, tol_piv, harris, bland, tex, th);
            theta = fmax(0.0, tex);
        }
        const double range = (stq == ST_FREE) ? CUDART_INF : upq - loq;
        if (range <= theta) {
            if (!(range < BIG)) { status = infeasible ? 1 : 6; break; }
            // ---- bound flip (replicated, no tableau traffic)
            __syncthreads();
            for (int i = tid; i < m; i += NT) c.beta[i] = fma(s * range, colq[i], c.beta[i]);
            if (tid == 0) c.cstat[q] = (stq == ST_LO) ? ST_UP : ST_LO;
            ++n_flip;
            changed = true;
            degen = 0; bland = false;
            __syncthreads();
            continue;
        }
        if (p == INT_MAX) { status = infeasible ? 1 : 6; break; }
        // ---- D. basis change
        const double piv = colq[p];
        const double inv = __drcp_rn(piv);
        const double xq_new = nb_value(stq, loq, upq) + s * theta;
        const double lov = c.rlo[p], upv = c.rup[p];
        const double wp_old = c.w[p];
        const int vleave = c.head[p];
        double target;
        {
            double lo = lov, up = upv;
            if (infeasible) {
                const double bp = c.beta[p];
                if (bp < lo - tol_bnd * (1.0 + fabs(lo))) up = lo;
                else if (bp > up + tol_bnd * (1.0 + fabs(up))) lo = up;
            }
            target = (s * piv > 0.0) ? up : lo;
        }
        __syncthreads();                      // everyone has read beta[p] / rlo[p] / w[p] / head[p]
        cl_stage_row(c, p, inv);
        for (int i = tid; i < m; i += NT)
            c.beta[i] = (i == p) ? xq_new : fma(s * theta, colq[i], c.beta[i]);
        __syncthreads();                      // rowp complete
        cl_arrive();                          // barrier #2, first half: this CTA has read row p
        // the price row rides along (phase 2: reduced costs; phase 1: frozen-weight prices)
        {
            double2 *d2 = reinterpret_cast<double2 *>(c.drow);
            const double2 *rowp2 = reinterpret_cast<const double2 *>(c.rowp);
            float2 *w2 = reinterpret_cast<float2 *>(c.dvx);
            const double cqd = dq * inv;
            const double wq = (double)c.dvx[q];
            const float wq_new = (float)fmax(wq * inv * inv, 1.0);
            int wbig = 0;
            __syncthreads();                  // every thread holds wq before dvx[q] is rewritten
            for (int j = tid; j < ((n + 1) >> 1); j += NT) {
                const double2 ra = rowp2[j];
                const int jq = q - 2 * j;
                if (dq != 0.0) {
                    double2 r = d2[j];
                    r.x = fma(-dq, ra.x, r.x);
                    r.y = fma(-dq, ra.y, r.y);
                    if (jq == 0) r.x = cqd;
                    if (jq == 1) r.y = cqd;
                    d2[j] = r;
                }
                // devex: w_j = max(w_j, (alpha_pj / alpha_pq)^2 w_q); the leaving variable gets max(w_q / alpha_pq^2, 1)
                float2 w = w2[j];
                w.x = fmaxf(w.x, (float)(ra.x * ra.x * wq));
                w.y = fmaxf(w.y, (float)(ra.y * ra.y * wq));
                if (jq == 0) w.x = wq_new;
                if (jq == 1) w.y = wq_new;
                if (2 * j + 1 >= n) w.y = 1.0f;
                wbig |= (w.x > 1e12f) | (w.y > 1e12f);
                w2[j] = w;
            }
            if (__syncthreads_or(wbig))       // new reference framework
                for (int j = tid; j < n; j += NT) c.dvx[j] = 1.0f;
        }
        if (tid == 0) {
            const signed char ns = (lov == upv) ? ST_FIX : (target == lov ? ST_LO : ST_UP);
            c.rlo[p] = loq; c.rup[p] = upq;
            c.cstat[q] = ns;
            c.head[p] = vq; c.nbv[q] = vleave;
            c.w[p] = 0.0;                     // the entering variable is feasible: weight 0
        }
        cl_sweep(c, colq, na_own, p, q, inv);
        cl_wait();                            // barrier #2, second half: every CTA has read row p
        cl_write_pivot_row(c, p, q, inv);
        if (infeasible) { p1_leave_col = q; p1_leave_w = wp_old; p1_fresh = false; }
        ++n_piv;
        n_rows += nz_local;
        changed = true;
        if (theta <= 1e-11) { if (++degen > degen_limit) bland = true; }
        else { degen = 0; bland = false; }
        __syncthreads();
    }

    // ---- epilogue: rank 0 publishes the state and the proposal; T in HBM is already current
    __syncthreads();
    CL_MARK(0, 7); CL_MARK(6, status);
    if (status != 5) {
        for (int j = tid; j < n; j += NT) {
            const int v = c.nbv[j];
            if (v >= m) c.rowp[v - m] = nb_value(c.cstat[j], __ldg(B.lo + v), __ldg(B.up + v));
        }
        for (int i = tid; i < m; i += NT) {
            const int v = c.head[i];
            if (v >= m) c.rowp[v - m] = c.beta[i];
        }
        __syncthreads();
    }
    const bool moved = changed || n_reb != 0 || P.initial != 0 || P.force_full != 0 || status != 5;
    // swept rows: every thread counted the non-zeros of the columns it scanned (replicated), so rank 0's
    // block-wide sum is the cluster-wide number
    const double rows_total = block_sum<NT>((double)n_rows, sh.scr);
    if (rank == 0) {
        if (tid == 0) B.dflag[0] = 0;         // the incremental reduced-cost cache is not used on this path
        if (!moved) {
            double zo = 0.0;
            for (int j = tid; j < n; j += NT) zo = fma(cs[j], c.rowp[j], zo);
            zo = block_sum<NT>(zo, sh.scr);
            if (tid == 0) { P.obj[k] = zo; P.status[k] = 5; P.last_iters[k] = 0; }
        } else {
            for (int i = tid; i < m; i += NT) { B.beta[i] = c.beta[i]; B.head[i] = c.head[i]; B.vstat[c.head[i]] = ST_BASIC; }
            for (int j = tid; j < n; j += NT) { B.nbv[j] = c.nbv[j]; B.vstat[c.nbv[j]] = c.cstat[j]; }
            double zo = 0.0, zc = 0.0;
            for (int j = tid; j < n; j += NT) {
                const double xj = c.rowp[j];
                B.x[j] = xj;
                zo = fma(cs[j], xj, zo);
                zc = fma(B.c_master[j], xj, zc);
            }
            zo = block_sum<NT>(zo, sh.scr);
            zc = block_sum<NT>(zc, sh.scr);
            double *yv = P.val + (size_t)k * L;
            int *yi = P.ind + (size_t)k * L;
            for (int i = tid; i < L; i += NT) {
                double y = 0.0;   // same order and rounding as dw_dcoogemv (dw_blas.c:119-121)
                for (int e = B.dr_ptr[i]; e < B.dr_ptr[i + 1]; ++e)
                    y = __dadd_rn(y, __dmul_rn(c.rowp[B.dr_col[e]], B.dr_val[e]));
                yv[i] = y;
            }
            __syncthreads();
            if (warp == 0) {
                int base = 0;
                for (int i0 = 0; i0 < L; i0 += 32) {
                    const int i = i0 + lane;
                    const double y = i < L ? yv[i] : 0.0;
                    const bool nz = (i < L) && (y != 0.0);
                    const unsigned bal = __ballot_sync(0xffffffffu, nz);
                    __syncwarp();
                    if (nz) {
                        const int pos = base + __popc(bal & ((1u << lane) - 1u));
                        yv[pos] = y;
                        yi[pos] = i + 1;
                    }
                    base += __popc(bal);
                    __syncwarp();
                }
                if (lane == 0) {
                    P.len[k] = base;
                    P.obj[k] = zo;
                    P.cost[k] = zc;
                    P.status[k] = status;
                    P.counters[8 * (size_t)k + 0] += n_piv;
                    P.counters[8 * (size_t)k + 1] += n_flip;
                    P.counters[8 * (size_t)k + 2] += n_reb;
                    P.counters[8 * (size_t)k + 3] += n_p1;
                    P.counters[8 * (size_t)k + 4] += (unsigned long long)rows_total;
                    P.counters[8 * (size_t)k + 5] += (unsigned long long)rows_total * (unsigned long long)ld;
                    P.counters[8 * (size_t)k + 6] += n_price;
                    P.last_iters[k] = (long long)(n_piv + n_flip);
                }
            }
        }
    }
    cl_sync();          // no CTA leaves while a peer could still address its shared memory
    CL_MARK(0, 8);
}

}  // namespace dwk
