// simplex_cta.cuh — one-CTA-per-block FP64 bounded-variable primal simplex for sm_100a.
//
// Replaces, for a whole batch of blocks in one launch, the per-thread loop body of the
// reference worker (/root/reference/src/dw_subprob.c:291-373): pricing D_k' pi and
// d = c_k - y (:291-321), the warm-started LP solve (:329, GLPK's glp_simplex — external,
// not in the reference tree), x extraction (:453-475), the column D_k x_k with exact-zero
// compaction and c_k . x_k (:481-540).
//
// Representation: condensed (Tucker) tableau  x_B = T x_N  over m auxiliary (row) variables
// and n structural variables (GLPK's formulation: aux_i = sum_j a_ij x_j carries the row
// bounds).  T is m x n; exactly n variables are nonbasic, so no identity columns are stored.
// A basis change is a rank-1 update of T (and of the reduced-cost row d); its algorithmic
// traffic is 16*(m+1)*(n+1) bytes.
//
// Two instantiations of one code path:
//   TSMEM = true   T and all vectors live in shared memory for the whole solve; state is
//                  loaded from / stored to HBM once per launch (small blocks, config A).
//   TSMEM = false  T and the vectors stay in global memory (L2-resident for mid-size blocks).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <limits.h>

namespace dwk {

constexpr signed char ST_LO = 0, ST_UP = 1, ST_FREE = 2, ST_FIX = 3, ST_BASIC = 4;
constexpr double BIG = 1e30;

struct BlockDev {
    int m, n, L, ldg;
    int sparse, pad1;       // sparse: the block starts on path 5 (packed tableau image at T, simplex_sparse.cuh)
    double *T;              // m x ldg (global, persistent between rounds)
    double *beta;           // m
    int *head;              // m
    int *nbv;               // n
    signed char *vstat;     // m+n, by variable id
    const double *lo, *up;  // m+n, by variable id (+-inf for none)
    const int *a_rp, *a_ci; // CSR of A_k
    const double *a_v;
    const int *dc_ptr, *dc_row;  // D_k by local column, reference order inside a column
    const double *dc_val;
    const int *dr_ptr, *dr_col;  // D_k by master row, reference order inside a row
    const double *dr_val;
    const double *c_file, *c_master;
    double *x;              // n (output of the last round)
    double *ws;             // global workspace for the vectors (TSMEM=false)
    double *dsave;          // n: reduced costs by column position at the end of the last optimal solve
    double *cssave;         // n: priced structural costs they were computed with
    int *dflag;             // solves since dsave was last recomputed in full (0 = dsave invalid)
    unsigned *rmask;        // path 3: per row, one bit per 32-byte item of T that may be non-zero (nullptr: none)
    int rwd;                // 32-bit words per row of rmask
    int dnnz;               // entries of D_k
    long long xoff;         // offset of this block's columns in the engine-wide column arrays (x, cs_pre)
    const double *Dt;       // dense D_k transposed, n x L row-major (blocks whose D_k is dense enough for the pricing GEMM; else nullptr)
    int *pos;               // m+n, by variable id: tableau row of a basic variable / column slot of a nonbasic one (blocks that
                            // take part in delta rounds, delta_tg.cuh; nullptr: not kept)
    int *spflag;            // path 5: 0 = T holds the packed image, 1 = the block moved to the dense layout of path 3 (nullptr: not a path-5 block)
};

struct SolveParams {
    const double *pi;   // L duals on the device (unused when initial)
    int phase_one;      // price with c == 0 (dw_subprob.c:302-303)
    int initial;        // cold solve with the block file's own objective (dw_subprob.c:121)
    int max_iter_mult;
    int no_dual_start;  // cluster path: skip the dual-simplex repair of an infeasible start
    const double *cs_pre; // priced costs d = (phase_one ? 0 : c_k) - D_k' pi of ALL columns (engine-wide, BlockDev::xoff), computed by
                        // the multi-vector pricing GEMM (pack.cuh: price_multi_dmma_kernel) for blocks with a dense D_k; nullptr: none
    int force_full;     // no proposal of this engine state exists yet (first round after create / reset without an
                        // initial solve): a block that needs no pivot must still write its whole proposal
    double tol_dj, tol_bnd, tol_piv;
    double *obj, *cost; // K
    int *status, *len;  // K
    int *ind;           // K*L
    double *val;        // K*L
    unsigned long long *counters;  // K*8: pivots, flips, rebuilds, phase-1 iterations, rows swept, 3 spare
    long long *last_iters;         // K
    long long *dbg;                // K*32 progress markers in mapped host memory (nullptr: off)
    // host mailboxes (mapped pinned memory, written by the kernels' epilogues; nullptr: off): same layout as the
    // device arrays above.  A block whose state did not change in a round only rewrites its obj.
    double *m_obj, *m_cost, *m_val;
    int *m_status, *m_len, *m_ind;
};

// working-set view (shared or global)
struct Ctx {
    int m, n, ld;
    double *T, *d, *d1, *rowp, *cs, *beta, *colq, *w;
    const double *lo, *up;
    int *nbv, *head;
    signed char *vstat;
};

__host__ __device__ inline size_t align8(size_t v) { return (v + 7) & ~size_t(7); }

// bytes of the vector part of the working set
__host__ __device__ inline size_t vec_bytes(int m, int n)
{
    return sizeof(double) * (size_t)(4 * n + 3 * m) + align8(sizeof(int) * (size_t)(n + m)) + align8((size_t)(m + n));
}
__host__ __device__ inline int smem_ld(int n) { return (n | 1); }
__host__ __device__ inline size_t smem_bytes(int m, int n)
{
    return sizeof(double) * (size_t)m * smem_ld(n) + vec_bytes(m, n) + sizeof(double) * 2 * (size_t)(m + n);
}

template <int NT>
struct Scratch {
    double v[NT / 32];
    int i[NT / 32];
};

__device__ __forceinline__ void warp_argmax(double &v, int &i)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, v, o);
        int oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
}

// argmax with smallest-index tie-break; every thread gets the result
template <int NT>
__device__ __forceinline__ void block_argmax(double &v, int &i, Scratch<NT> &s)
{
    constexpr int NW = NT / 32;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    warp_argmax(v, i);
    if (lane == 0) { s.v[w] = v; s.i[w] = i; }
    __syncthreads();
    v = lane < NW ? s.v[lane] : -CUDART_INF;
    i = lane < NW ? s.i[lane] : INT_MAX;
    warp_argmax(v, i);
    __syncthreads();
}

template <int NT>
__device__ __forceinline__ double block_min(double v, Scratch<NT> &s)
{
    constexpr int NW = NT / 32;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) s.v[w] = v;
    __syncthreads();
    v = lane < NW ? s.v[lane] : CUDART_INF;
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    return v;
}

template <int NT>
__device__ __forceinline__ double block_max(double v, Scratch<NT> &s)
{
    return -block_min<NT>(-v, s);
}

// deterministic sum (fixed tree for a fixed NT)
template <int NT>
__device__ __forceinline__ double block_sum(double v, Scratch<NT> &s)
{
    constexpr int NW = NT / 32;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s.v[w] = v;
    __syncthreads();
    v = lane < NW ? s.v[lane] : 0.0;
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    return v;
}

__device__ __forceinline__ double nb_value(signed char st, double lo, double up)
{
    return st == ST_UP ? up : (st == ST_FREE ? 0.0 : lo);
}

__device__ __forceinline__ signed char initial_stat(double lo, double up)
{
    const bool flo = lo > -BIG, fup = up < BIG;
    if (flo && fup && lo == up) return ST_FIX;
    if (flo) return ST_LO;
    if (fup) return ST_UP;
    return ST_FREE;
}

// T <- A (dense scatter), all auxiliaries basic
template <int NT>
__device__ void reset_tableau(const BlockDev &B, Ctx &c)
{
    const int m = c.m, n = c.n, ld = c.ld;
    {
        // 16-byte streaming stores: the reset of config B zeroes 8.6 GB, it should run at HBM speed
        const size_t tot = (size_t)m * ld;
        const bool vec = (tot & 1) == 0 && (reinterpret_cast<size_t>(c.T) & 15) == 0;
        if (vec) {
            double2 *t2 = reinterpret_cast<double2 *>(c.T);
            const double2 z = make_double2(0.0, 0.0);
            for (size_t e = threadIdx.x; e < tot / 2; e += NT) __stcs(t2 + e, z);
        } else {
            for (size_t e = threadIdx.x; e < tot; e += NT) c.T[e] = 0.0;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += NT) {
        for (int p = B.a_rp[i]; p < B.a_rp[i + 1]; ++p) c.T[(size_t)i * ld + B.a_ci[p]] += B.a_v[p];
        c.head[i] = i;
    }
    for (int j = threadIdx.x; j < n; j += NT) c.nbv[j] = m + j;
    __syncthreads();
}

// beta_i = sum_j T_ij xN_j ; one warp per row, fixed summation tree
template <int NT>
__device__ void recompute_beta(Ctx &c)
{
    const int m = c.m, n = c.n, ld = c.ld;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    constexpr int NW = NT / 32;
    for (int i = w; i < m; i += NW) {
        double s = 0.0;
        for (int j = lane; j < n; j += 32) {
            const int v = c.nbv[j];
            const double xv = nb_value(c.vstat[v], c.lo[v], c.up[v]);
            s += c.T[(size_t)i * ld + j] * xv;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) c.beta[i] = s;
    }
    __syncthreads();
}

// Tucker pivot on (p,q) of the tableau body; optionally carries the reduced-cost row d.
// Preconditions: colq[i] = T[i][q] for all i (staged by the caller), barrier reached.
template <int NT>
__device__ __forceinline__ void pivot_update(Ctx &c, int p, int q, bool with_d)
{
    const int m = c.m, n = c.n, ld = c.ld;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    constexpr int NW = NT / 32;
    const double inv = 1.0 / c.colq[p];
    const double dq = with_d ? c.d[q] : 0.0;
    for (int j = threadIdx.x; j < n; j += NT) c.rowp[j] = c.T[(size_t)p * ld + j] * inv;
    __syncthreads();
    const int rows = with_d ? m + 1 : m;
    for (int i = w; i < rows; i += NW) {
        double *row = i < m ? c.T + (size_t)i * ld : c.d;
        const double ci = i < m ? c.colq[i] : dq;
        if (i == p) {
            for (int j = lane; j < n; j += 32) row[j] = (j == q) ? inv : -c.rowp[j];
        } else if (ci != 0.0) {
            const double cq = ci * inv;
            for (int j = lane; j < n; j += 32) {
                const double t = row[j];
                row[j] = (j == q) ? cq : fma(-ci, c.rowp[j], t);
            }
        }
    }
}

// limits of basic row i for a step of the entering variable in direction s
struct RowLim { double a; double target; bool blocks; };
__device__ __forceinline__ RowLim row_limit(const Ctx &c, int i, double s, double tiq, bool infeasible,
                                            double tol_bnd, double tol_piv)
{
    RowLim r; r.a = s * tiq; r.blocks = false; r.target = 0.0;
    const int v = c.head[i];
    double lo = c.lo[v], up = c.up[v];
    if (infeasible) {
        const double b = c.beta[i];
        if (b < lo - tol_bnd * (1.0 + fabs(lo))) { up = lo; lo = -CUDART_INF; }       // below: may rise to lo
        else if (b > up + tol_bnd * (1.0 + fabs(up))) { lo = up; up = CUDART_INF; }  // above: may fall to up
    }
    if (r.a > tol_piv) { if (up < BIG) { r.blocks = true; r.target = up; } }
    else if (r.a < -tol_piv) { if (lo > -BIG) { r.blocks = true; r.target = lo; } }
    return r;
}

template <bool TSMEM, int NT>
__device__ void solve_block(const BlockDev &B, const SolveParams &P, int k, Ctx &c, Scratch<NT> &scr,
                            int *s_misc)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NT / 32;
    const int m = c.m, n = c.n, ld = c.ld, L = B.L;
    const double tol_dj = P.tol_dj, tol_bnd = P.tol_bnd, tol_piv = P.tol_piv;
    const double harris = 0.5 * tol_bnd;

    // ---- prologue: priced objective cs = (phase_one ? 0 : c_k) - D_k' pi   (dw_subprob.c:291-321)
    // multiply-then-add in the reference's order, no FMA contraction, so cs is bit-identical
    for (int j = tid; j < n; j += NT) {
        double cj;
        if (P.initial) cj = B.c_file[j];
        else {
            double y = 0.0;
            for (int e = B.dc_ptr[j]; e < B.dc_ptr[j + 1]; ++e)
                y = __dadd_rn(y, __dmul_rn(__ldg(P.pi + B.dc_row[e]), B.dc_val[e]));
            const double base = P.phase_one ? 0.0 : B.c_master[j];
            cj = __dadd_rn(__dmul_rn(-1.0, y), base);
        }
        c.cs[j] = cj;
    }
    __syncthreads();

    int status = 0;
    long long iter = 0;
    const long long max_iter = (long long)P.max_iter_mult * (m + n) + 1000;
    unsigned long long n_piv = 0, n_flip = 0, n_reb = 0, n_p1 = 0;
    bool need_check = true, infeasible = false, d_valid = false, bland = false;
    int degen = 0;
    const int degen_limit = 100 + (m + n) / 2;

    for (;;) {
        if (++iter > max_iter) { status = 1; break; }
        // ---- primal feasibility scan (solve start, after a rebuild, every phase-1 step)
        if (need_check || infeasible) {
            int viol = 0;
            for (int i = tid; i < m; i += NT) {
                const int v = c.head[i];
                const double b = c.beta[i], lo = c.lo[v], up = c.up[v];
                double wi = 0.0;
                if (b < lo - tol_bnd * (1.0 + fabs(lo))) wi = -1.0;
                else if (b > up + tol_bnd * (1.0 + fabs(up))) wi = 1.0;
                c.w[i] = wi;
                viol |= (wi != 0.0);
            }
            infeasible = __syncthreads_or(viol) != 0;
            need_check = false;
            if (infeasible) d_valid = false;
        }
        const double *dd;
        if (infeasible) {
            // phase-1 prices: d1 = T' w   (w = -1 / +1 on rows below / above their bounds)
            for (int j = tid; j < n; j += NT) {
                double s = 0.0;
                for (int i = 0; i < m; ++i) {
                    const double wi = c.w[i];
                    if (wi != 0.0) s += wi * c.T[(size_t)i * ld + j];
                }
                c.d1[j] = s;
            }
            dd = c.d1;
            ++n_p1;
            __syncthreads();
        } else {
            if (!d_valid) {
                // d = c_N + T' c_B
                for (int i = tid; i < m; i += NT) {
                    const int v = c.head[i];
                    c.colq[i] = v >= m ? c.cs[v - m] : 0.0;
                }
                __syncthreads();
                for (int j = tid; j < n; j += NT) {
                    const int v = c.nbv[j];
                    double s = v >= m ? c.cs[v - m] : 0.0;
                    for (int i = 0; i < m; ++i) {
                        const double cb = c.colq[i];
                        if (cb != 0.0) s = fma(cb, c.T[(size_t)i * ld + j], s);
                    }
                    c.d[j] = s;
                }
                d_valid = true;
                __syncthreads();
            }
            dd = c.d;
        }
        // ---- entering variable: Dantzig (largest |d_j|), or Bland (smallest variable id)
        double key = -CUDART_INF; int q = INT_MAX;
        for (int j = tid; j < n; j += NT) {
            const double dj = dd[j];
            const int v = c.nbv[j];
            const signed char st = c.vstat[v];
            const bool ok = (st == ST_LO && dj < -tol_dj) || (st == ST_UP && dj > tol_dj) ||
                            (st == ST_FREE && fabs(dj) > tol_dj);
            if (ok) {
                const double kj = bland ? -(double)v : fabs(dj);
                if (kj > key) { key = kj; q = j; }
            }
        }
        block_argmax<NT>(key, q, scr);
        if (q == INT_MAX) {
            if (infeasible) { status = 4; break; }
            // ---- claimed optimal: extract x, check the residual against the original rows
            for (int j = tid; j < n; j += NT) {
                const int v = c.nbv[j];
                const double xv = nb_value(c.vstat[v], c.lo[v], c.up[v]);
                if (v >= m) c.rowp[v - m] = xv; else c.colq[v] = xv;
            }
            for (int i = tid; i < m; i += NT) {
                const int v = c.head[i];
                if (v >= m) c.rowp[v - m] = c.beta[i]; else c.colq[v] = c.beta[i];
            }
            __syncthreads();
            double worst = 0.0;
            for (int i = tid; i < m; i += NT) {
                double s = 0.0;
                for (int p = B.a_rp[i]; p < B.a_rp[i + 1]; ++p) s = fma(B.a_v[p], c.rowp[B.a_ci[p]], s);
                worst = fmax(worst, fabs(s - c.colq[i]) / (1.0 + fabs(s)));
            }
            worst = block_max<NT>(worst, scr);
            if (worst > 0.1 * tol_bnd && n_reb < 3) {
                // ---- rebuild T from A_k for the current basis, then keep iterating
                ++n_reb;
                reset_tableau<NT>(B, c);
                bool singular = false;
                for (int j = 0; j < n; ++j) {
                    if (c.vstat[m + j] != ST_BASIC) continue;
                    double pk = -1.0; int pr = INT_MAX;
                    for (int i = tid; i < m; i += NT) {
                        const int h = c.head[i];
                        const double tq = c.T[(size_t)i * ld + j];
                        c.colq[i] = tq;
                        if (h < m && c.vstat[h] != ST_BASIC) {
                            const double a = fabs(tq);
                            if (a > pk) { pk = a; pr = i; }
                        }
                    }
                    block_argmax<NT>(pk, pr, scr);
                    if (pr == INT_MAX || pk < 1e-11) { singular = true; break; }
                    pivot_update<NT>(c, pr, j, false);
                    if (tid == 0) { const int t = c.head[pr]; c.head[pr] = c.nbv[j]; c.nbv[j] = t; }
                    __syncthreads();
                }
                if (singular) { status = 1; break; }
                recompute_beta<NT>(c);
                need_check = true; d_valid = false;
                continue;
            }
            status = 5;
            break;
        }
        const double dq = dd[q];
        const double s = dq < 0.0 ? 1.0 : -1.0;
        const int vq = c.nbv[q];
        const signed char stq = c.vstat[vq];
        // ---- ratio test, pass 1 (Harris: bounds relaxed by delta)  [Bland: exact]
        double tmax = CUDART_INF;
        for (int i = tid; i < m; i += NT) {
            const double tiq = c.T[(size_t)i * ld + q];
            c.colq[i] = tiq;
            const RowLim r = row_limit(c, i, s, tiq, infeasible, tol_bnd, tol_piv);
            if (r.blocks) {
                const double delta = bland ? 0.0 : harris * (1.0 + fabs(r.target));
                double t = (r.target - c.beta[i]) / r.a + delta / fabs(r.a);
                t = fmax(t, 0.0);
                tmax = fmin(tmax, t);
            }
        }
        tmax = block_min<NT>(tmax, scr);     // also publishes colq
        // ---- pass 2: among rows whose exact ratio is within tmax take the largest |pivot|
        double pk = -CUDART_INF; int p = INT_MAX;
        if (tmax < CUDART_INF) {
            const double slack = bland ? 1e-12 * (1.0 + tmax) : 0.0;
            for (int i = tid; i < m; i += NT) {
                const RowLim r = row_limit(c, i, s, c.colq[i], infeasible, tol_bnd, tol_piv);
                if (r.blocks) {
                    const double t = (r.target - c.beta[i]) / r.a;
                    if (t <= tmax + slack) {
                        const double kk = bland ? -(double)c.head[i] : fabs(r.a);
                        if (kk > pk) { pk = kk; p = i; }
                    }
                }
            }
            block_argmax<NT>(pk, p, scr);
        }
        double theta = CUDART_INF, target = 0.0;
        if (p != INT_MAX) {
            const RowLim r = row_limit(c, p, s, c.colq[p], infeasible, tol_bnd, tol_piv);
            target = r.target;
            theta = fmax(0.0, (r.target - c.beta[p]) / r.a);
        }
        const double range = (stq == ST_FREE) ? CUDART_INF : c.up[vq] - c.lo[vq];
        if (range <= theta) {
            if (!(range < BIG)) { status = infeasible ? 1 : 6; break; }
            // ---- bound flip: no basis change
            for (int i = tid; i < m; i += NT) c.beta[i] = fma(s * range, c.colq[i], c.beta[i]);
            if (tid == 0) c.vstat[vq] = (stq == ST_LO) ? ST_UP : ST_LO;
            ++n_flip;
            degen = 0; bland = false;
            __syncthreads();
            continue;
        }
        // ---- basis change
        const double xq_new = nb_value(stq, c.lo[vq], c.up[vq]) + s * theta;
        const int vleave = c.head[p];
        __syncthreads();   // every thread has read beta[p]/head[p] before they change
        for (int i = tid; i < m; i += NT)
            c.beta[i] = (i == p) ? xq_new : fma(s * theta, c.colq[i], c.beta[i]);
        pivot_update<NT>(c, p, q, d_valid && !infeasible);
        if (tid == 0) {
            const double lov = c.lo[vleave], upv = c.up[vleave];
            c.vstat[vleave] = (lov == upv) ? ST_FIX : (target == lov ? ST_LO : ST_UP);
            c.vstat[vq] = ST_BASIC;
            c.head[p] = vq;
            c.nbv[q] = vleave;
        }
        ++n_piv;
        if (theta <= 1e-11) { if (++degen > degen_limit) bland = true; }
        else { degen = 0; bland = false; }
        __syncthreads();
    }

    // ---- epilogue: x_k, obj = d.x, cost = c_k.x, column D_k x_k (exact zeros dropped)
    if (status != 5) {
        // still publish the current point (the reference reads glp_get_col_prim regardless)
        for (int j = tid; j < n; j += NT) {
            const int v = c.nbv[j];
            if (v >= m) c.rowp[v - m] = nb_value(c.vstat[v], c.lo[v], c.up[v]);
        }
        for (int i = tid; i < m; i += NT) {
            const int v = c.head[i];
            if (v >= m) c.rowp[v - m] = c.beta[i];
        }
        __syncthreads();
    }
    double zo = 0.0, zc = 0.0;
    for (int j = tid; j < n; j += NT) {
        const double xj = c.rowp[j];
        B.x[j] = xj;
        zo = fma(c.cs[j], xj, zo);
        zc = fma(B.c_master[j], xj, zc);
    }
    zo = block_sum<NT>(zo, scr);
    zc = block_sum<NT>(zc, scr);
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
        // order-preserving in-place compaction of y != 0.0  (dw_subprob.c:506-516)
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
            P.counters[8 * (size_t)k + 4] += n_piv * (unsigned long long)(m + 1);   // dense sweep bound
            P.counters[8 * (size_t)k + 5] += n_piv * (unsigned long long)(m + 1) * (unsigned long long)n;
            P.last_iters[k] = (long long)(n_piv + n_flip);
        }
    }
    (void)s_misc; (void)NW;
}

// carve the working set out of a flat buffer (shared or global)
__device__ __forceinline__ void carve(Ctx &c, char *p, int m, int n)
{
    c.d = (double *)p;      p += sizeof(double) * n;
    c.d1 = (double *)p;     p += sizeof(double) * n;
    c.rowp = (double *)p;   p += sizeof(double) * n;
    c.cs = (double *)p;     p += sizeof(double) * n;
    c.beta = (double *)p;   p += sizeof(double) * m;
    c.colq = (double *)p;   p += sizeof(double) * m;
    c.w = (double *)p;      p += sizeof(double) * m;
    c.nbv = (int *)p;       p += sizeof(int) * n;
    c.head = (int *)p;      p += align8(sizeof(int) * (size_t)(n + m)) - sizeof(int) * n;
    c.vstat = (signed char *)p;
}

template <bool TSMEM, int NT>
__global__ void __launch_bounds__(NT) solve_kernel(const BlockDev *blocks, const int *list, int count,
                                                   SolveParams P)
{
    extern __shared__ __align__(16) char smem_raw[];
    __shared__ Scratch<NT> scr;
    __shared__ int s_misc[4];
    if ((int)blockIdx.x >= count) return;
    const int k = list[blockIdx.x];
    const BlockDev B = blocks[k];
    const int m = B.m, n = B.n, tid = threadIdx.x;
    Ctx c;
    c.m = m; c.n = n;
    if (TSMEM) {
        c.ld = smem_ld(n);
        c.T = (double *)smem_raw;
        char *p = smem_raw + sizeof(double) * (size_t)m * c.ld;
        double *slo = (double *)p;  p += sizeof(double) * (m + n);
        double *sup = (double *)p;  p += sizeof(double) * (m + n);
        carve(c, p, m, n);
        for (int v = tid; v < m + n; v += NT) { slo[v] = B.lo[v]; sup[v] = B.up[v]; }
        c.lo = slo; c.up = sup;
        for (int i = 0; i < m; ++i)
            for (int j = tid; j < n; j += NT) c.T[(size_t)i * c.ld + j] = B.T[(size_t)i * B.ldg + j];
    } else {
        c.ld = B.ldg;
        c.T = B.T;
        carve(c, (char *)B.ws, m, n);
        c.lo = B.lo; c.up = B.up;
    }
    for (int i = tid; i < m; i += NT) { c.beta[i] = B.beta[i]; c.head[i] = B.head[i]; }
    for (int j = tid; j < n; j += NT) c.nbv[j] = B.nbv[j];
    for (int v = tid; v < m + n; v += NT) c.vstat[v] = B.vstat[v];
    __syncthreads();

    solve_block<TSMEM, NT>(B, P, k, c, scr, s_misc);

    __syncthreads();
    if (TSMEM) {
        for (int i = 0; i < m; ++i)
            for (int j = tid; j < n; j += NT) B.T[(size_t)i * B.ldg + j] = c.T[(size_t)i * c.ld + j];
    }
    for (int i = tid; i < m; i += NT) { B.beta[i] = c.beta[i]; B.head[i] = c.head[i]; }
    for (int j = tid; j < n; j += NT) B.nbv[j] = c.nbv[j];
    for (int v = tid; v < m + n; v += NT) B.vstat[v] = c.vstat[v];
}

template <int NT> __device__ void sp_init_image(const BlockDev &B);     // simplex_sparse.cuh

// initial state: all-logical basis, structurals at a finite bound (GLPK's "standard" basis
// after glp_read_lp), T = A_k, beta = A_k x_N
template <int NT>
__global__ void __launch_bounds__(NT) init_kernel(const BlockDev *blocks, int K)
{
    const int k = blockIdx.x;
    if (k >= K) return;
    const BlockDev B = blocks[k];
    Ctx c;
    c.m = B.m; c.n = B.n; c.ld = B.ldg; c.T = B.T;
    c.beta = B.beta; c.head = B.head; c.nbv = B.nbv; c.vstat = B.vstat; c.lo = B.lo; c.up = B.up;
    c.d = c.d1 = c.rowp = c.cs = c.colq = c.w = nullptr;
    for (int v = threadIdx.x; v < B.m + B.n; v += NT)
        B.vstat[v] = v < B.m ? ST_BASIC : initial_stat(B.lo[v], B.up[v]);
    if (threadIdx.x == 0) B.dflag[0] = 0;
    __syncthreads();
    if (B.sparse) {
        sp_init_image<NT>(B);        // path 5: packed image (simplex_sparse.cuh)
    } else if (B.rmask != nullptr) {
        // Path 3 reads a tableau cell only through its row's item mask, so only the 32-byte items that hold an
        // entry of A_k are written (whole items: their other cells are real zeros) and the rest of the 1 MB
        // row-major image is left as it is — the reset of config B writes 0.7 GB instead of zeroing 8.6 GB.
        for (int e = threadIdx.x; e < B.m * B.rwd; e += NT) B.rmask[e] = 0u;
        __syncthreads();
        for (int i = threadIdx.x; i < B.m; i += NT) {            // one thread per row: no race
            double *row = c.T + (size_t)i * c.ld;
            for (int p = B.a_rp[i]; p < B.a_rp[i + 1]; ++p) {
                const int t = B.a_ci[p] >> 2;
                B.rmask[(size_t)i * B.rwd + (t >> 5)] |= 1u << (t & 31);
                reinterpret_cast<double2 *>(row)[2 * t] = make_double2(0.0, 0.0);
                reinterpret_cast<double2 *>(row)[2 * t + 1] = make_double2(0.0, 0.0);
            }
            for (int p = B.a_rp[i]; p < B.a_rp[i + 1]; ++p) row[B.a_ci[p]] += B.a_v[p];
            c.head[i] = i;
        }
        for (int j = threadIdx.x; j < B.n; j += NT) c.nbv[j] = B.m + j;
        __syncthreads();
    } else
    reset_tableau<NT>(B, c);
    if (B.pos != nullptr) {          // all-logical basis: variable i sits in row i, structural j in slot j
        for (int i = threadIdx.x; i < B.m; i += NT) B.pos[i] = i;
        for (int j = threadIdx.x; j < B.n; j += NT) B.pos[B.m + j] = j;
    }
    // beta = A_k x_N straight from the CSR rows (all structurals are nonbasic): no second pass over T
    for (int i = threadIdx.x; i < B.m; i += NT) {
        double s = 0.0;
        for (int p = B.a_rp[i]; p < B.a_rp[i + 1]; ++p) {
            const int v = B.m + B.a_ci[p];
            s = fma(B.a_v[p], nb_value(B.vstat[v], B.lo[v], B.up[v]), s);
        }
        B.beta[i] = s;
    }
    // x_k of the initial state (the no-pivot fast path of a later round reads it back)
    for (int j = threadIdx.x; j < B.n; j += NT) {
        const int v = B.m + j;
        B.x[j] = nb_value(B.vstat[v], B.lo[v], B.up[v]);
    }
}

}  // namespace dwk
