// simplex_smem.cuh — shared-memory-resident CTA-per-block simplex (config A class blocks).
//
// Same algorithm and tolerances as simplex_cta.cuh (the generic kernel), specialised for blocks
// whose whole working set fits one CTA's shared memory:
//   * the block's state image [T | beta head nbv vstat] and its bounds are moved HBM<->SMEM with
//     TMA bulk copies (cp.async.bulk + mbarrier): one elected thread issues three copies, nobody
//     spins on LDG latency, and T is written back only if the basis changed;
//   * the reduced-cost row d is row m of the shared tableau, so one rank-1 sweep updates both;
//   * pricing (argmax |d_j|) and the two-pass Harris ratio test run inside warp 0 with shuffle
//     reductions only; bound flips loop inside warp 0 without any block barrier; a basis change
//     costs two __syncthreads (decision -> update -> next decision);
//   * the rank-1 update is row-per-warp, 16-byte (double2) shared accesses, four rows in flight
//     per warp, and rows whose pivot-column entry is exactly zero are skipped (network LPs have
//     very sparse tableau columns).
// Reference being replaced: one warm-started glp_simplex call per block and round
// (/root/reference/src/dw_subprob.c:329) plus the pricing / column code around it (:291-321,
// :453-540).
#pragma once
#include <stdint.h>

#include "simplex_cta.cuh"

namespace dwk {

// ---- TMA bulk copy / mbarrier helpers (sm_90+ PTX, emitted as UBLKCP / SYNCS on sm_100a) -------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase)
{
    uint32_t ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    const char *s = (const char *)src;
    char *d = (char *)dst;
    while (bytes) {   // chunks of <= 32 KiB
        const uint32_t c = bytes > 32768u ? 32768u : bytes;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(d)), "l"(s), "r"(c), "r"(smem_u32(bar)) : "memory");
        s += c; d += c; bytes -= c;
    }
}
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, uint32_t bytes)
{
    char *d = (char *)dst;
    const char *s = (const char *)src;
    while (bytes) {
        const uint32_t c = bytes > 32768u ? 32768u : bytes;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(d), "r"(smem_u32(s)), "r"(c) : "memory");
        s += c; d += c; bytes -= c;
    }
}

// ---- layouts ---------------------------------------------------------------------------------
__host__ __device__ inline size_t align16(size_t v) { return (v + 15) & ~size_t(15); }
__host__ __device__ inline int smem2_ld(int n) { return (n + 2) & ~1; }                 // even, > n
// persistent small state image: beta[m] | head[m] | nbv[n] | vstat[m+n], padded to 16 bytes
__host__ __device__ inline size_t pst_bytes(int m, int n)
{
    return align16(sizeof(double) * (size_t)m + sizeof(int) * (size_t)(m + n) + (size_t)(m + n));
}
__host__ __device__ inline size_t smem2_bytes(int m, int n)
{
    const size_t ld = smem2_ld(n);
    return align16(sizeof(double) * (size_t)(m + 1) * ld)     // T and the d row
         + align16(sizeof(double) * 2 * (size_t)(m + n))      // lo | up
         + pst_bytes(m, n)
         + align16(sizeof(double) * ld)                       // rowp
         + align16(sizeof(double) * (size_t)n) * 2            // cs, d1
         + align16(sizeof(double) * (size_t)(m + 1))          // colq
         + align16(sizeof(double) * (size_t)m);               // w
}

enum { K_NONE = 0, K_PIVOT = 1, K_FAIL = 2, K_REDO = 3 };

struct Decision {
    int kind, p, q, fail_status;
    double inv;
};

struct Smem2 {
    int m, n, ld;
    double *T;       // (m+1) x ld, row m = reduced costs d
    double *lo, *up;
    double *beta;
    int *head, *nbv;
    signed char *vstat;
    double *rowp, *cs, *d1, *colq, *w;
};

__device__ __forceinline__ double warp_min(double v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ RowLim row_limit2(const Smem2 &c, int i, double s, double tiq, bool infeasible,
                                             double tol_bnd, double tol_piv)
{
    RowLim r; r.a = s * tiq; r.blocks = false; r.target = 0.0;
    const int v = c.head[i];
    double lo = c.lo[v], up = c.up[v];
    if (infeasible) {
        const double b = c.beta[i];
        if (b < lo - tol_bnd * (1.0 + fabs(lo))) { up = lo; lo = -CUDART_INF; }
        else if (b > up + tol_bnd * (1.0 + fabs(up))) { lo = up; up = CUDART_INF; }
    }
    if (r.a > tol_piv) { if (up < BIG) { r.blocks = true; r.target = up; } }
    else if (r.a < -tol_piv) { if (lo > -BIG) { r.blocks = true; r.target = lo; } }
    return r;
}

// Executed by warp 0 only.  Chooses the entering column and the leaving row, applies bound flips
// on the spot, and for a basis change stages everything the rank-1 sweep needs (colq, rowp, beta,
// head/nbv/vstat).  dd = reduced costs to price with (row m of T in phase 2, d1 in phase 1).
__device__ int decide_warp0(Smem2 &c, const double *dd, bool infeasible, bool d_live, bool &bland, int &degen,
                            int degen_limit, double tol_dj, double tol_bnd, double tol_piv,
                            unsigned long long &n_piv, unsigned long long &n_flip, Decision *dec)
{
    const int lane = threadIdx.x & 31;
    const int m = c.m, n = c.n, ld = c.ld;
    const double harris = 0.5 * tol_bnd;
    for (;;) {
        double key = -CUDART_INF; int q = INT_MAX;
        for (int j = lane; j < n; j += 32) {
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
        warp_argmax(key, q);
        if (q == INT_MAX) return K_NONE;
        const double dq = dd[q];
        const double s = dq < 0.0 ? 1.0 : -1.0;
        const int vq = c.nbv[q];
        const signed char stq = c.vstat[vq];
        // pass 1: Harris bound on the step
        double tmax = CUDART_INF;
        for (int i = lane; i < m; i += 32) {
            const double tiq = c.T[(size_t)i * ld + q];
            c.colq[i] = tiq;
            const RowLim r = row_limit2(c, i, s, tiq, infeasible, tol_bnd, tol_piv);
            if (r.blocks) {
                const double delta = bland ? 0.0 : harris * (1.0 + fabs(r.target));
                const double t = fmax((r.target - c.beta[i]) / r.a + delta / fabs(r.a), 0.0);
                tmax = fmin(tmax, t);
            }
        }
        tmax = warp_min(tmax);
        // pass 2: largest |pivot| among rows whose exact ratio is within the bound
        double pk = -CUDART_INF; int p = INT_MAX;
        if (tmax < CUDART_INF) {
            const double slack = bland ? 1e-12 * (1.0 + tmax) : 0.0;
            for (int i = lane; i < m; i += 32) {
                const RowLim r = row_limit2(c, i, s, c.colq[i], infeasible, tol_bnd, tol_piv);
                if (r.blocks) {
                    const double t = (r.target - c.beta[i]) / r.a;
                    if (t <= tmax + slack) {
                        const double kk = bland ? -(double)c.head[i] : fabs(r.a);
                        if (kk > pk) { pk = kk; p = i; }
                    }
                }
            }
            warp_argmax(pk, p);
        }
        __syncwarp();
        double theta = CUDART_INF, target = 0.0;
        if (p != INT_MAX) {
            const RowLim r = row_limit2(c, p, s, c.colq[p], infeasible, tol_bnd, tol_piv);
            target = r.target;
            theta = fmax(0.0, (r.target - c.beta[p]) / r.a);
        }
        const double range = (stq == ST_FREE) ? CUDART_INF : c.up[vq] - c.lo[vq];
        if (range <= theta) {
            if (!(range < BIG)) { if (lane == 0) dec->fail_status = infeasible ? 1 : 6; return K_FAIL; }
            for (int i = lane; i < m; i += 32) c.beta[i] = fma(s * range, c.colq[i], c.beta[i]);
            if (lane == 0) c.vstat[vq] = (stq == ST_LO) ? ST_UP : ST_LO;
            ++n_flip;
            degen = 0; bland = false;
            __syncwarp();
            if (infeasible) return K_REDO;     // phase-1 prices depend on beta
            continue;
        }
        if (p == INT_MAX) { if (lane == 0) dec->fail_status = infeasible ? 1 : 6; return K_FAIL; }
        // ---- basis change: stage the sweep
        const double xq_new = nb_value(stq, c.lo[vq], c.up[vq]) + s * theta;
        const int vleave = c.head[p];
        const double inv = 1.0 / c.colq[p];
        __syncwarp();
        for (int i = lane; i < m; i += 32)
            c.beta[i] = (i == p) ? xq_new : fma(s * theta, c.colq[i], c.beta[i]);
        for (int j = lane; j < ld; j += 32) c.rowp[j] = c.T[(size_t)p * ld + j] * inv;
        if (lane == 0) {
            c.colq[m] = d_live ? dq : 0.0;
            const double lov = c.lo[vleave], upv = c.up[vleave];
            c.vstat[vleave] = (lov == upv) ? ST_FIX : (target == lov ? ST_LO : ST_UP);
            c.vstat[vq] = ST_BASIC;
            c.head[p] = vq;
            c.nbv[q] = vleave;
            dec->p = p; dec->q = q; dec->inv = inv;
        }
        ++n_piv;
        if (theta <= 1e-11) { if (++degen > degen_limit) bland = true; }
        else { degen = 0; bland = false; }
        return K_PIVOT;
    }
}

// rank-1 sweep over rows [0, rows) of T (row m = d): row-per-warp, double2, 4 rows in flight,
// rows with a zero multiplier are skipped
template <int NT>
__device__ __forceinline__ void sweep(Smem2 &c, int p, int q, double inv, int rows)
{
    constexpr int NW = NT / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nc2 = c.ld >> 1;
    double2 *T2 = reinterpret_cast<double2 *>(c.T);
    const double2 *rowp2 = reinterpret_cast<const double2 *>(c.rowp);
    for (int c0 = 0; c0 < nc2; c0 += 32) {
        const int j2 = c0 + lane;
        const bool act = j2 < nc2;
        const double2 rp = act ? rowp2[j2] : make_double2(0.0, 0.0);
        const int jq = q - 2 * j2;               // 0: .x is column q, 1: .y is column q
        for (int i0 = warp; i0 < rows; i0 += 4 * NW) {
            double ci[4]; double2 v[4]; bool on[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * NW;
                const bool in = i < rows;
                ci[u] = in ? c.colq[i] : 0.0;
                on[u] = in && act && (ci[u] != 0.0 || i == p);
                if (on[u]) v[u] = T2[(size_t)i * nc2 + j2];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * NW;
                if (on[u]) {
                    double2 r;
                    if (i == p) {
                        r.x = jq == 0 ? inv : -rp.x;
                        r.y = jq == 1 ? inv : -rp.y;
                    } else {
                        const double cq = ci[u] * inv;
                        r.x = jq == 0 ? cq : fma(-ci[u], rp.x, v[u].x);
                        r.y = jq == 1 ? cq : fma(-ci[u], rp.y, v[u].y);
                    }
                    T2[(size_t)i * nc2 + j2] = r;
                }
            }
        }
    }
}

template <int NT>
__global__ void __launch_bounds__(NT, 3) solve_smem_kernel(const BlockDev *blocks, const int *list, int count,
                                                          SolveParams P)
{
    extern __shared__ __align__(16) char smem_raw[];
    __shared__ Scratch<NT> scr;
    __shared__ Decision dec;
    __shared__ __align__(8) uint64_t bar;
    __shared__ int s_kind;
    if ((int)blockIdx.x >= count) return;
    const int k = list[blockIdx.x];
    const BlockDev B = blocks[k];
    const int m = B.m, n = B.n, L = B.L, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double tol_dj = P.tol_dj, tol_bnd = P.tol_bnd, tol_piv = P.tol_piv;

    Smem2 c;
    c.m = m; c.n = n; c.ld = smem2_ld(n);
    const int ld = c.ld;
    char *sp = smem_raw;
    c.T = (double *)sp;            sp += align16(sizeof(double) * (size_t)(m + 1) * ld);
    c.lo = (double *)sp; c.up = c.lo + (m + n);  sp += align16(sizeof(double) * 2 * (size_t)(m + n));
    char *pst = sp;                sp += pst_bytes(m, n);
    c.beta = (double *)pst;
    c.head = (int *)(pst + sizeof(double) * (size_t)m);
    c.nbv = c.head + m;
    c.vstat = (signed char *)(c.nbv + n);
    c.rowp = (double *)sp;         sp += align16(sizeof(double) * ld);
    c.cs = (double *)sp;           sp += align16(sizeof(double) * (size_t)n);
    c.d1 = (double *)sp;           sp += align16(sizeof(double) * (size_t)n);
    c.colq = (double *)sp;         sp += align16(sizeof(double) * (size_t)(m + 1));
    c.w = (double *)sp;
    double *drow = c.T + (size_t)m * ld;

    // ---- state in: three bulk copies, one mbarrier
    const uint32_t t_bytes = (uint32_t)(sizeof(double) * (size_t)m * ld);
    const uint32_t b_bytes = (uint32_t)(sizeof(double) * 2 * (size_t)(m + n));
    const uint32_t p_bytes = (uint32_t)pst_bytes(m, n);
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_expect_tx(&bar, t_bytes + b_bytes + p_bytes);
        if (t_bytes) bulk_g2s(c.T, B.T, t_bytes, &bar);
        if (b_bytes) bulk_g2s(c.lo, B.lo, b_bytes, &bar);
        bulk_g2s(pst, B.beta, p_bytes, &bar);
    }
    // priced objective while the copies fly (same arithmetic as the generic kernel / reference)
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
    __syncthreads();            // mbarrier init + cs visible
    mbar_wait(&bar, 0);

    int status = 0;
    long long iter = 0;
    const long long max_iter = (long long)P.max_iter_mult * (m + n) + 1000;
    unsigned long long n_piv = 0, n_flip = 0, n_reb = 0, n_p1 = 0;   // live in warp 0, lane-uniform
    bool need_check = true, infeasible = false, d_valid = false, bland = false;
    int degen = 0;
    const int degen_limit = 100 + (m + n) / 2;

    for (;;) {
        if (++iter > max_iter) { status = 1; break; }
        // ---- A. cooperative parts: feasibility scan, phase-1 prices, fresh reduced costs
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
        if (infeasible) {
            for (int j = tid; j < n; j += NT) {
                double s = 0.0;
                for (int i = 0; i < m; ++i) {
                    const double wi = c.w[i];
                    if (wi != 0.0) s += wi * c.T[(size_t)i * ld + j];
                }
                c.d1[j] = s;
            }
            ++n_p1;
            __syncthreads();
        } else if (!d_valid) {
            for (int i = tid; i < m; i += NT) {
                const int v = c.head[i];
                c.colq[i] = v >= m ? c.cs[v - m] : 0.0;
            }
            __syncthreads();
            for (int j = tid; j < ld; j += NT) {
                double s = 0.0;
                if (j < n) {
                    const int v = c.nbv[j];
                    s = v >= m ? c.cs[v - m] : 0.0;
                    for (int i = 0; i < m; ++i) {
                        const double cb = c.colq[i];
                        if (cb != 0.0) s = fma(cb, c.T[(size_t)i * ld + j], s);
                    }
                }
                drow[j] = s;
            }
            d_valid = true;
            __syncthreads();
        }
        // ---- B. decision (warp 0)
        if (warp == 0) {
            const int kind = decide_warp0(c, infeasible ? c.d1 : drow, infeasible, d_valid && !infeasible, bland, degen,
                                          degen_limit, tol_dj, tol_bnd, tol_piv, n_piv, n_flip, &dec);
            if (lane == 0) s_kind = kind;
        }
        __syncthreads();
        const int kind = s_kind;
        if (kind == K_PIVOT) {
            sweep<NT>(c, dec.p, dec.q, dec.inv, m + 1);
            __syncthreads();
            continue;
        }
        if (kind == K_REDO) continue;
        if (kind == K_FAIL) { status = dec.fail_status; break; }
        // ---- K_NONE
        if (infeasible) { status = 4; break; }
        // claimed optimal: extract x, check the residual against the original rows
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
        if (worst > 0.1 * tol_bnd && n_reb < 3) {          // n_reb is block-uniform
            ++n_reb;
            // rebuild T from A_k for the current basis
            for (int e = tid; e < m * ld; e += NT) c.T[e] = 0.0;
            __syncthreads();
            for (int i = tid; i < m; i += NT) {
                for (int p = B.a_rp[i]; p < B.a_rp[i + 1]; ++p) c.T[(size_t)i * ld + B.a_ci[p]] += B.a_v[p];
                c.head[i] = i;
            }
            for (int j = tid; j < n; j += NT) c.nbv[j] = m + j;
            __syncthreads();
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
                const double inv = 1.0 / c.colq[pr];
                for (int jj = tid; jj < ld; jj += NT) c.rowp[jj] = c.T[(size_t)pr * ld + jj] * inv;
                __syncthreads();
                sweep<NT>(c, pr, j, inv, m);
                if (tid == 0) { const int t = c.head[pr]; c.head[pr] = c.nbv[j]; c.nbv[j] = t; }
                __syncthreads();
            }
            if (singular) { status = 1; break; }
            {   // beta = T x_N, one warp per row
                constexpr int NW = NT / 32;
                for (int i = warp; i < m; i += NW) {
                    double s = 0.0;
                    for (int j = lane; j < n; j += 32) {
                        const int v = c.nbv[j];
                        s += c.T[(size_t)i * ld + j] * nb_value(c.vstat[v], c.lo[v], c.up[v]);
                    }
#pragma unroll
                    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                    if (lane == 0) c.beta[i] = s;
                }
                __syncthreads();
            }
            need_check = true; d_valid = false;
            continue;
        }
        status = 5;
        break;
    }

    // ---- epilogue
    if (status != 5) {
        __syncthreads();
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
    // state out first (async), then the proposal while the copies fly
    __syncthreads();
    if (tid == 0) {
        const bool t_dirty = (n_piv + n_reb) != 0;        // thread 0 sits in warp 0, which owns n_piv
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (t_dirty && t_bytes) bulk_s2g(B.T, c.T, t_bytes);
        bulk_s2g(B.beta, pst, p_bytes);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
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
        double y = 0.0;
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
            P.counters[4 * (size_t)k + 0] += n_piv;
            P.counters[4 * (size_t)k + 1] += n_flip;
            P.counters[4 * (size_t)k + 2] += n_reb;
            P.counters[4 * (size_t)k + 3] += n_p1;
            P.last_iters[k] = (long long)(n_piv + n_flip);
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    }
}

}  // namespace dwk
