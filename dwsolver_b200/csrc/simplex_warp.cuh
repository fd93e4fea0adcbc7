// simplex_warp.cuh — one WARP per block: bounded-variable primal simplex on a tableau that lives in HBM
// behind row-item masks (the state format of path 3, simplex_smem.cuh with TG = true).
//
// Why: the CTA-per-block kernel of path 3 spends a basis change as a ~14 k-cycle dependent chain with seven
// CTA barriers, at 16 warps per SM (4 CTAs) and < 30 % of the issue slots (profiles/r2d_ncu_B.md: barrier 26 %
// of the stall samples, warps active 24 %).  A network block's basis change touches ~13 rows x ~23 32-byte items —
// work for one warp, not four.  Here a block is ONE warp: no CTA barrier anywhere (only __syncwarp), every list is
// built with ballots, and the shared state of a 256 x 512 block is 20 KB instead of 29 KB, so 10-11 blocks are
// resident per SM instead of 4 and the dependent chains of 2.5x as many blocks overlap.
//   * pricing: 16 columns per lane, redux.sync arg-max with lowest-column tie-break;
//   * pivot column: the rows whose mask holds the column's item are listed by ballot from shared memory alone; then
//     the column entry, both bounds of the row's basic variable and beta are fetched per listed row in one batch
//     (two chunks of 32 rows in flight) and the Harris pass 1 runs on registers;
//   * pivot row: its item list is its mask; the items are fetched, scaled into shared memory and written back;
//   * rank-1 sweep over (listed rows) x (listed items): lanes are laid out as (row sub-group, item), DWK_W_U
//     rows in flight per lane before the first FMA, masks kept exact by shared-memory atomics;
//   * bounds are read by variable id from HBM/L2 (no per-row copies in shared memory).
// Anything rare and bulky (the Gauss-Jordan rebuild after a failed residual check) is NOT done here: the warp writes
// its consistent state back, appends the block to a redo list and the CTA kernel of path 3 continues from that basis
// in a second launch of the same round (solve_smem_kernel with a Continuation).
// Reference being replaced: one warm-started glp_simplex call per block and round
// (/root/reference/src/dw_subprob.c:329) plus the pricing / column code around it (:291-321, :453-540).
#pragma once
#include <stdint.h>

#include "simplex_smem.cuh"

#ifndef DWK_W_OCC
#define DWK_W_OCC 10        // warps (= blocks of the LP) per SM the warp kernel is compiled for
#endif
#ifndef DWK_W_U
#define DWK_W_U 8           // tableau rows in flight per lane in the sweep
#endif

namespace dwk {

// redo list of a round: blocks the warp kernel handed over to the CTA kernel
struct WarpExtra {
    int *redo_list;             // K
    int *redo_count;            // 1 (zeroed before the launch)
    unsigned char *moved;       // K: the basis or a bound status changed before the hand-over
    int bail_after;             // test hook: hand over after this many basis changes (0: never)
};

__host__ __device__ inline size_t warp_smem_bytes(int m, int n)
{
    const size_t ld = tab_ld(n, true);
    const size_t scr = sizeof(double) * (ld > (size_t)(m + 1) ? ld : (size_t)(m + 1));
    return align16(sizeof(double) * ld)                           // d (reduced costs / phase-1 prices)
         + align16(scr)                                           // scr: scaled pivot-row items | cached ratios | x staging
         + align16(sizeof(double) * (size_t)m)                    // beta
         + align16(sizeof(double) * (size_t)(m + 1))              // colv: pivot-column entries by list position | weights by row
         + rmask_bytes(m, n)                                      // row-item masks
         + align16(sizeof(unsigned short) * (size_t)m)            // head
         + align16(sizeof(unsigned short) * (size_t)n)            // nbv
         + align16((size_t)(m + n))                               // vstat
         + align16((size_t)n)                                     // cstat
         + align16(sizeof(unsigned short) * (size_t)(m + 1))      // arow
         + align16(sizeof(unsigned short) * (ld / 4 + 1));        // acol
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// (key, idx) -> idx of the largest non-negative key among valid lanes, ties to the smallest idx; -1 if none
__device__ __forceinline__ int warp_argmax_lowidx(double key, int idx, bool valid)
{
    const unsigned hi = valid ? (unsigned)__double2hiint(key) : 0u;
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const bool c1 = valid && hi == mh;
    const unsigned lo = c1 ? (unsigned)__double2loint(key) : 0u;
    const unsigned ml = __reduce_max_sync(0xffffffffu, lo);
    const bool c2 = c1 && lo == ml;
    const unsigned r = __reduce_min_sync(0xffffffffu, c2 ? (unsigned)idx : 0xffffffffu);
    return r == 0xffffffffu ? -1 : (int)r;
}

struct WarpCtx {
    int m, n, ld, nitems, rwd;
    double *T;
    double *d, *scr, *beta, *colv;
    unsigned *rmask;
    unsigned short *head, *nbv, *arow, *acol;
    signed char *vstat, *cstat;
    const double *glo, *gup;
    double *cs;
};

// out[4t .. 4t+3] += sum over the listed rows (ascending) of wt[row] * T[row][4t .. 4t+3] for every item t, visiting
// only the (row, item) pairs the row masks know.  Lane l owns the items t = l (mod 32): no two lanes touch the same
// cell of out, the order of the sum is the row order, and a lane keeps up to 8 items (32 B each) in flight.
__device__ __forceinline__ unsigned long long warp_weighted_itemsum(const WarpCtx &c, const double *wt, int nl, double *out)
{
    const int lane = threadIdx.x;
    const int n2 = c.ld >> 1, rwd = c.rwd;
    const double2 *T2 = reinterpret_cast<const double2 *>(c.T);
    double2 *o2 = reinterpret_cast<double2 *>(out);
    unsigned cells = 0;
    for (int w0 = 0; w0 < rwd; w0 += 4) {                  // four mask words = 128 items = 4 per lane
        for (int b0 = 0; b0 < nl; b0 += 32) {
            const int lim = min(32, nl - b0);
            unsigned h0 = 0, h1 = 0, h2 = 0, h3 = 0;       // bit u: listed row b0+u carries my item of word w0+k
            for (int u = 0; u < lim; ++u) {
                const unsigned *mw = c.rmask + (size_t)c.arow[b0 + u] * rwd + w0;
                h0 |= ((mw[0] >> lane) & 1u) << u;
                if (w0 + 1 < rwd) h1 |= ((mw[1] >> lane) & 1u) << u;
                if (w0 + 2 < rwd) h2 |= ((mw[2] >> lane) & 1u) << u;
                if (w0 + 3 < rwd) h3 |= ((mw[3] >> lane) & 1u) << u;
            }
            cells += __popc(h0) + __popc(h1) + __popc(h2) + __popc(h3);
            while (h0 | h1 | h2 | h3) {                    // per-lane loop: no warp collectives inside
                double2 v0[8], v1[8];
                double wv[8];
                int it[8];
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    it[s] = -1;
                    if (h0 | h1 | h2 | h3) {
                        const int kk = h0 ? 0 : (h1 ? 1 : (h2 ? 2 : 3));
                        const unsigned hh = h0 ? h0 : (h1 ? h1 : (h2 ? h2 : h3));
                        const int u = __ffs(hh) - 1;
                        if (kk == 0) h0 &= h0 - 1; else if (kk == 1) h1 &= h1 - 1; else if (kk == 2) h2 &= h2 - 1; else h3 &= h3 - 1;
                        const int i = c.arow[b0 + u];
                        const int t = (w0 + kk) * 32 + lane;
                        it[s] = t;
                        wv[s] = wt[i];
                        v0[s] = __ldcg(T2 + (size_t)i * n2 + 2 * t);
                        v1[s] = __ldcg(T2 + (size_t)i * n2 + 2 * t + 1);
                    }
                }
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    if (it[s] >= 0) {
                        double2 a0 = o2[2 * it[s]], a1 = o2[2 * it[s] + 1];
                        a0.x = fma(wv[s], v0[s].x, a0.x); a0.y = fma(wv[s], v0[s].y, a0.y);
                        a1.x = fma(wv[s], v1[s].x, a1.x); a1.y = fma(wv[s], v1[s].y, a1.y);
                        o2[2 * it[s]] = a0; o2[2 * it[s] + 1] = a1;
                    }
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) cells += __shfl_xor_sync(0xffffffffu, cells, o);
    return 4ull * cells;
}

// rows with a non-zero weight, ascending, into arow; count returned to every lane
__device__ __forceinline__ int warp_list_rows(const WarpCtx &c, const double *wt)
{
    const int lane = threadIdx.x;
    int nl = 0;
    for (int i0 = 0; i0 < c.m; i0 += 32) {
        const int i = i0 + lane;
        const bool nz = i < c.m && wt[i] != 0.0;
        const unsigned bal = __ballot_sync(0xffffffffu, nz);
        if (nz) c.arow[nl + __popc(bal & ((1u << lane) - 1u))] = (unsigned short)i;
        nl += __popc(bal);
    }
    __syncwarp();
    return nl;
}

template <int MC, int NC>
__global__ void __launch_bounds__(32, DWK_W_OCC) solve_warp_kernel(const BlockDev *blocks, const int *list, int count,
                                                                   SolveParams P, WarpExtra X)
{
    extern __shared__ __align__(16) char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    if ((int)blockIdx.x >= count) return;
    const int k = list[blockIdx.x];
    const BlockDev &B = blocks[k];
    const int m = MC ? MC : B.m, n = NC ? NC : B.n, L = B.L, lane = threadIdx.x;
    const unsigned lt_mask = (1u << lane) - 1u;
    const double tol_dj = P.tol_dj, tol_bnd = P.tol_bnd, tol_piv = P.tol_piv;
    const double harris = 0.5 * tol_bnd;

    WarpCtx c;
    c.m = m; c.n = n; c.ld = tab_ld(n, true); c.nitems = c.ld >> 2; c.rwd = rmask_words(n);
    const int ld = c.ld, rwd = c.rwd, n2 = ld >> 1;
    c.T = B.T; c.glo = B.lo; c.gup = B.up; c.cs = B.ws;
    {
        char *sp = smem_raw;
        const size_t scrb = sizeof(double) * (size_t)(ld > m + 1 ? ld : m + 1);
        c.d = (double *)sp;               sp += align16(sizeof(double) * (size_t)ld);
        c.scr = (double *)sp;             sp += align16(scrb);
        c.beta = (double *)sp;            sp += align16(sizeof(double) * (size_t)m);
        c.colv = (double *)sp;            sp += align16(sizeof(double) * (size_t)(m + 1));
        c.rmask = (unsigned *)sp;         sp += rmask_bytes(m, n);
        c.head = (unsigned short *)sp;    sp += align16(sizeof(unsigned short) * (size_t)m);
        c.nbv = (unsigned short *)sp;     sp += align16(sizeof(unsigned short) * (size_t)n);
        c.vstat = (signed char *)sp;      sp += align16((size_t)(m + n));
        c.cstat = (signed char *)sp;      sp += align16((size_t)n);
        c.arow = (unsigned short *)sp;    sp += align16(sizeof(unsigned short) * (size_t)(m + 1));
        c.acol = (unsigned short *)sp;
    }
    double2 *T2 = reinterpret_cast<double2 *>(c.T);

    // ---- state in: masks and beta by bulk copy (one mbarrier), the integer lists narrowed to 16 bits by the lanes
    const uint32_t k_bytes = (uint32_t)rmask_bytes(m, n);
    const uint32_t b_bytes = (uint32_t)(sizeof(double) * (size_t)m);
    const bool bulk_beta = (b_bytes & 15u) == 0u && b_bytes != 0u;
    if (lane == 0) {
        mbar_init(&bar, 1);
        mbar_expect_tx(&bar, k_bytes + (bulk_beta ? b_bytes : 0u));
        bulk_g2s(c.rmask, B.rmask, k_bytes, &bar);
        if (bulk_beta) bulk_g2s(c.beta, B.beta, b_bytes, &bar);
    }
    {
        const int *gh = B.head, *gn = B.nbv;
        const signed char *gv = B.vstat;
        for (int i = lane; i < m; i += 32) c.head[i] = (unsigned short)__ldcg(gh + i);
        for (int j = lane; j < n; j += 32) c.nbv[j] = (unsigned short)__ldcg(gn + j);
        for (int v = lane; v < m + n; v += 32) c.vstat[v] = __ldcg(gv + v);
        if (!bulk_beta) { const double *gb = B.beta; for (int i = lane; i < m; i += 32) c.beta[i] = __ldcg(gb + i); }
    }
    // priced objective while the copies fly (same arithmetic as the CTA kernels / reference: dw_subprob.c:291-321)
    if (P.initial) {
        const double *cf = B.c_file;
        for (int j = lane; j < n; j += 32) c.cs[j] = __ldg(cf + j);
    } else {
        const int *dcp = B.dc_ptr, *dcr = B.dc_row;
        const double *dcv = B.dc_val, *cm = B.c_master;
        constexpr int U = 8;
        for (int j0 = lane; j0 < n; j0 += 32 * U) {
            int pb[U], pe[U], r0[U];
            double v0[U], p0[U], base[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int j = j0 + u * 32;
                pb[u] = pe[u] = 0;
                if (j < n) { pb[u] = __ldg(dcp + j); pe[u] = __ldg(dcp + j + 1); }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int j = j0 + u * 32;
                base[u] = (j < n && !P.phase_one) ? __ldg(cm + j) : 0.0;
                r0[u] = 0; v0[u] = 0.0;
                if (pb[u] < pe[u]) { r0[u] = __ldg(dcr + pb[u]); v0[u] = __ldg(dcv + pb[u]); }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) p0[u] = pb[u] < pe[u] ? __ldg(P.pi + r0[u]) : 0.0;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int j = j0 + u * 32;
                if (j >= n) continue;
                double y = 0.0;
                if (pb[u] < pe[u]) {
                    y = __dadd_rn(y, __dmul_rn(p0[u], v0[u]));
                    for (int e = pb[u] + 1; e < pe[u]; ++e)
                        y = __dadd_rn(y, __dmul_rn(__ldg(P.pi + __ldg(dcr + e)), __ldg(dcv + e)));
                }
                c.cs[j] = __dadd_rn(__dmul_rn(-1.0, y), base[u]);
            }
        }
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    for (int j = lane; j < n; j += 32) c.cstat[j] = c.vstat[c.nbv[j]];
    __syncwarp();

    int status = 0;
    long long iter = 0;
    const long long max_iter = (long long)P.max_iter_mult * (m + n) + 1000;
    unsigned long long n_piv = 0, n_flip = 0, n_p1 = 0, n_rows = 0, n_cells = 0, n_price = 0;   // warp-uniform
    bool need_check = true, infeasible = false, d_valid = false, bland = false, d_full = false, x_in_scr = true;
    bool bail = false;
    int degen = 0;
    const int degen_limit = 100 + (m + n) / 2;
    const int dflag_in = B.dflag[0];
    const bool use_saved = dflag_in > 0 && dflag_in < 32;

    for (;;) {
        if (++iter > max_iter) { status = 1; break; }
        // ---- A. feasibility scan, phase-1 prices, fresh reduced costs
        if (need_check || infeasible) {
            bool viol = false;
            constexpr int U = 8;
            for (int i0 = lane; i0 < m; i0 += 32 * U) {
                double lo[U], up[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int i = i0 + u * 32;
                    if (i < m) { const int v = c.head[i]; lo[u] = __ldg(c.glo + v); up[u] = __ldg(c.gup + v); }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int i = i0 + u * 32;
                    if (i < m) {
                        const double b = c.beta[i];
                        double wi = 0.0;
                        if (b < lo[u] - tol_bnd * (1.0 + fabs(lo[u]))) wi = -1.0;
                        else if (b > up[u] + tol_bnd * (1.0 + fabs(up[u]))) wi = 1.0;
                        c.colv[i] = wi;
                        viol |= (wi != 0.0);
                    }
                }
            }
            infeasible = __any_sync(0xffffffffu, viol);
            need_check = false;
            if (infeasible) d_valid = false;
            __syncwarp();
        }
        if (infeasible) {
            // phase-1 prices T'w into d (consumed by the pricing below; d is recomputed once the block is feasible)
            const int nl = warp_list_rows(c, c.colv);
            for (int j = lane; j < ld; j += 32) c.d[j] = 0.0;
            __syncwarp();
            n_price += warp_weighted_itemsum(c, c.colv, nl, c.d);
            ++n_p1;
            __syncwarp();
        } else if (!d_valid) {
            // reduced costs.  If the basis has not moved since the last optimal solve of this block, update the saved
            // row incrementally: d = d_saved + dc_N + T' dc_B with dc = cs - cs_saved, which only touches the rows of T
            // whose basic cost changed; otherwise (and every 32nd solve, to stop drift) recompute d = c_N + T' c_B.
            const bool inc = use_saved && (n_piv + n_flip) == 0;
            const double *cssave = B.cssave, *dsave = B.dsave;
            constexpr int U = 8;
            for (int i0 = lane; i0 < m; i0 += 32 * U) {
                double a[U], b[U];
                bool st[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int i = i0 + u * 32;
                    st[u] = false; a[u] = 0.0; b[u] = 0.0;
                    if (i < m) {
                        const int v = c.head[i];
                        if (v >= m) { st[u] = true; a[u] = __ldcg(c.cs + v - m); if (inc) b[u] = __ldg(cssave + v - m); }
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int i = i0 + u * 32;
                    if (i < m) c.colv[i] = st[u] ? (inc ? cost_delta(a[u], b[u]) : a[u]) : 0.0;
                }
            }
            __syncwarp();
            const int nl = warp_list_rows(c, c.colv);
            for (int j0 = lane; j0 < ld; j0 += 32 * U) {
                double a[U], b[U], dsv[U];
                bool st[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int j = j0 + u * 32;
                    st[u] = false; a[u] = 0.0; b[u] = 0.0; dsv[u] = 0.0;
                    if (j < n) {
                        const int v = c.nbv[j];
                        if (v >= m) { st[u] = true; a[u] = __ldcg(c.cs + v - m); if (inc) b[u] = __ldg(cssave + v - m); }
                        if (inc) dsv[u] = __ldg(dsave + j);
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int j = j0 + u * 32;
                    if (j >= ld) continue;
                    double s = 0.0;
                    if (j < n) {
                        if (st[u]) s = inc ? cost_delta(a[u], b[u]) : a[u];
                        if (inc) s += dsv[u];
                    }
                    c.d[j] = s;
                }
            }
            __syncwarp();
            n_price += warp_weighted_itemsum(c, c.colv, nl, c.d);
            d_valid = true;
            d_full = !inc;
            __syncwarp();
        }

        // ---- B. decision + sweep, repeated while the prices stay valid
        int kind = K_NONE;
        for (;;) {
            // pricing
            int q;
            {
                double key = 0.0; int qj = -1;
                for (int j = lane; j < n; j += 32) {
                    const double dj = c.d[j];
                    const signed char st = c.cstat[j];
                    const bool ok = (st == ST_LO && dj < -tol_dj) || (st == ST_UP && dj > tol_dj) ||
                                    (st == ST_FREE && fabs(dj) > tol_dj);
                    const double kj = bland ? 2147483648.0 - (double)c.nbv[j] : fabs(dj);
                    if (ok && kj > key) { key = kj; qj = j; }
                }
                q = warp_argmax_lowidx(key, qj, qj >= 0);
            }
            if (q < 0) { kind = K_NONE; break; }
            const double dq = c.d[q];
            const double s = dq < 0.0 ? 1.0 : -1.0;
            const signed char stq = c.cstat[q];
            const int vq = c.nbv[q];
            const double loq = __ldg(c.glo + vq), upq = __ldg(c.gup + vq);
            // rows whose mask knows the item of column q (a superset of the column's non-zeros): shared memory only
            int nf = 0;
            {
                const int tq = q >> 2;
                const unsigned *mq = c.rmask + (tq >> 5);
                const int bq = tq & 31;
                for (int i0 = 0; i0 < m; i0 += 32) {
                    const int i = i0 + lane;
                    const bool f = i < m && ((mq[(size_t)i * rwd] >> bq) & 1u);
                    const unsigned bal = __ballot_sync(0xffffffffu, f);
                    if (f) c.arow[nf + __popc(bal & lt_mask)] = (unsigned short)i;
                    nf += __popc(bal);
                }
                __syncwarp();
            }
            // column entries, bounds and beta of the listed rows, two chunks of 32 in flight; Harris pass 1 on
            // registers; rows with an exact zero leave the list (in place: a write never overtakes a pending read)
            double tmax = CUDART_INF;
            int na = 0;
            for (int c0 = 0; c0 < nf; c0 += 64) {
                double tc[2], lo[2], up[2], bb[2];
                int ii[2];
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int e = c0 + 32 * r + lane;
                    ii[r] = -1; tc[r] = 0.0; lo[r] = 0.0; up[r] = 0.0; bb[r] = 0.0;
                    if (e < nf) {
                        const int i = c.arow[e];
                        ii[r] = i;
                        tc[r] = __ldcg(c.T + (size_t)i * ld + q);
                        const int hv = c.head[i];
                        lo[r] = __ldg(c.glo + hv); up[r] = __ldg(c.gup + hv);
                        bb[r] = c.beta[i];
                    }
                }
                __syncwarp();
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    if (c0 + 32 * r < nf) {                     // warp-uniform
                    const bool nz = ii[r] >= 0 && tc[r] != 0.0;
                    double tex = CUDART_INF;
                    if (nz) {
                        const double a = s * tc[r];
                        double l = lo[r], u = up[r];
                        const double b = bb[r];
                        if (infeasible) {
                            if (b < l - tol_bnd * (1.0 + fabs(l))) { u = l; l = -CUDART_INF; }
                            else if (b > u + tol_bnd * (1.0 + fabs(u))) { l = u; u = CUDART_INF; }
                        }
                        bool blocks = false; double slack = 0.0, target = 0.0;
                        if (a > tol_piv) { if (u < BIG) { blocks = true; slack = u - b; target = u; } }
                        else if (a < -tol_piv) { if (l > -BIG) { blocks = true; slack = b - l; target = l; } }
                        if (blocks) {
                            const double rr = __drcp_rn(fabs(a));
                            tex = slack * rr;
                            const double delta = bland ? 0.0 : harris * (1.0 + fabs(target));
                            tmax = fmin(tmax, fmax((slack + delta) * rr, 0.0));
                        }
                    }
                    const unsigned bal = __ballot_sync(0xffffffffu, nz);
                    if (nz) {
                        const int pos = na + __popc(bal & lt_mask);
                        c.arow[pos] = (unsigned short)ii[r];
                        c.colv[pos] = tc[r];
                        c.scr[pos] = tex;                      // cached by list position
                    }
                    na += __popc(bal);
                    }
                }
                __syncwarp();
            }
            tmax = warp_min_nonneg(tmax);
            // pass 2: largest |pivot| among the rows whose exact ratio is within the bound
            int p = -1, pe = -1;
            if (tmax < CUDART_INF) {
                const double lim = bland ? tmax + 1e-12 * (1.0 + tmax) : tmax;
                double pk = 0.0; int be = -1;
                for (int e = lane; e < na; e += 32) {
                    if (c.scr[e] <= lim) {
                        const double kk = bland ? 2147483648.0 - (double)c.head[c.arow[e]] : fabs(c.colv[e]);
                        if (kk > pk) { pk = kk; be = e; }
                    }
                }
                pe = warp_argmax_lowidx(pk, be, be >= 0);
                if (pe >= 0) p = c.arow[pe];
            }
            const double theta = p >= 0 ? fmax(0.0, c.scr[pe]) : CUDART_INF;
            const double range = (stq == ST_FREE) ? CUDART_INF : upq - loq;
            if (range <= theta) {
                if (!(range < BIG)) { status = infeasible ? 1 : 6; kind = K_FAIL; break; }
                for (int e = lane; e < na; e += 32) {
                    const int i = c.arow[e];
                    c.beta[i] = fma(s * range, c.colv[e], c.beta[i]);
                }
                if (lane == 0) {
                    const signed char ns = (stq == ST_LO) ? ST_UP : ST_LO;
                    c.cstat[q] = ns;
                    c.vstat[vq] = ns;
                }
                ++n_flip;
                degen = 0; bland = false;
                __syncwarp();
                if (infeasible) { kind = K_REDO; break; }      // phase-1 prices depend on beta
                continue;
            }
            if (p < 0) { status = infeasible ? 1 : 6; kind = K_FAIL; break; }
            // ---- basis change
            const double piv = c.colv[pe];
            const double inv = __drcp_rn(piv);
            const double xq_new = nb_value(stq, loq, upq) + s * theta;
            const int vleave = c.head[p];
            const double lov = __ldg(c.glo + vleave), upv = __ldg(c.gup + vleave);
            double target;
            {
                double lo = lov, up = upv;
                if (infeasible) {
                    const double bp = c.beta[p];
                    if (bp < lo - tol_bnd * (1.0 + fabs(lo))) up = lo;          // rises to its lower bound
                    else if (bp > up + tol_bnd * (1.0 + fabs(up))) lo = up;     // falls to its upper bound
                }
                target = (s * piv > 0.0) ? up : lo;
            }
            __syncwarp();
            for (int e = lane; e < na; e += 32) {
                const int i = c.arow[e];
                c.beta[i] = (i == p) ? xq_new : fma(s * theta, c.colv[e], c.beta[i]);
            }
            // item list of the pivot row = its mask
            int nc = 0;
            {
                const unsigned *mp = c.rmask + (size_t)p * rwd;
                for (int w_ = 0; w_ < rwd; ++w_) {
                    const unsigned bits = mp[w_];
                    if ((bits >> lane) & 1u) c.acol[nc + __popc(bits & lt_mask)] = (unsigned short)(w_ * 32 + lane);
                    nc += __popc(bits);
                }
            }
            if (lane == 0) {
                const signed char ns = (lov == upv) ? ST_FIX : (target == lov ? ST_LO : ST_UP);
                c.cstat[q] = ns; c.vstat[vleave] = ns; c.vstat[vq] = ST_BASIC;
                c.head[p] = (unsigned short)vq; c.nbv[q] = (unsigned short)vleave;
            }
            const bool with_d = !infeasible && dq != 0.0;
            ++n_piv;
            n_rows += (unsigned long long)na;
            n_cells += (unsigned long long)(na + (with_d ? 1 : 0)) * (unsigned long long)nc * 4ull;
            if (theta <= 1e-11) { if (++degen > degen_limit) bland = true; }
            else { degen = 0; bland = false; }
            __syncwarp();
            // pivot row: fetch its items, scale them into scr (by list position), write the new row back
            const int tq = q >> 2, cq_pos = q & 3;
            for (int e0 = 0; e0 < nc; e0 += 32 * 2) {
                double2 v0[2], v1[2];
                int it[2];
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int e = e0 + 32 * r + lane;
                    it[r] = -1;
                    if (e < nc) {
                        it[r] = c.acol[e];
                        v0[r] = __ldcg(T2 + (size_t)p * n2 + 2 * it[r]);
                        v1[r] = __ldcg(T2 + (size_t)p * n2 + 2 * it[r] + 1);
                    }
                }
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int e = e0 + 32 * r + lane;
                    if (it[r] >= 0) {
                        double2 a0 = make_double2(v0[r].x * inv, v0[r].y * inv);
                        double2 a1 = make_double2(v1[r].x * inv, v1[r].y * inv);
                        reinterpret_cast<double2 *>(c.scr)[2 * e] = a0;
                        reinterpret_cast<double2 *>(c.scr)[2 * e + 1] = a1;
                        double2 r0 = make_double2(-a0.x, -a0.y), r1 = make_double2(-a1.x, -a1.y);
                        if (it[r] == tq) {
                            if (cq_pos == 0) r0.x = inv; else if (cq_pos == 1) r0.y = inv; else if (cq_pos == 2) r1.x = inv; else r1.y = inv;
                        }
                        __stcg(T2 + (size_t)p * n2 + 2 * it[r], r0);
                        __stcg(T2 + (size_t)p * n2 + 2 * it[r] + 1, r1);
                    }
                }
            }
            __syncwarp();
            // rank-1 sweep over (listed rows except p) x (listed items); lanes = (row sub-group, item)
            {
                int lg = 5;
                if (nc <= 16) { lg = 4; if (nc <= 8) lg = 3; if (nc <= 4) lg = 2; }
                const int ncp = 1 << lg, G = 32 >> lg;
                const int esub = lane & (ncp - 1), sub = lane >> lg;
                for (int e0 = 0; e0 < nc; e0 += ncp) {
                    const int e = e0 + esub;
                    const bool ve = e < nc;
                    const int item = ve ? (int)c.acol[e] : 0;
                    double2 rv0 = make_double2(0.0, 0.0), rv1 = rv0;
                    if (ve) { rv0 = reinterpret_cast<const double2 *>(c.scr)[2 * e]; rv1 = reinterpret_cast<const double2 *>(c.scr)[2 * e + 1]; }
                    const int mword = item >> 5;
                    const unsigned mbit = 1u << (item & 31);
                    const bool isq = item == tq;
                    for (int a0 = 0; a0 < na; a0 += G * DWK_W_U) {
                        double2 v0[DWK_W_U], v1[DWK_W_U];
                        int ri[DWK_W_U];                       // row index; -1: nothing to do
                        bool had[DWK_W_U];
                        double ci[DWK_W_U];
#pragma unroll
                        for (int u = 0; u < DWK_W_U; ++u) {
                            const int a = a0 + u * G + sub;
                            ri[u] = -1; had[u] = false; ci[u] = 0.0;
                            v0[u] = make_double2(0.0, 0.0); v1[u] = v0[u];
                            if (ve && a < na) {
                                const int i = c.arow[a];
                                if (i != p) {
                                    ri[u] = i;
                                    ci[u] = c.colv[a];
                                    had[u] = (c.rmask[(size_t)i * rwd + mword] & mbit) != 0u;
                                    if (had[u]) {
                                        v0[u] = __ldcg(T2 + (size_t)i * n2 + 2 * item);
                                        v1[u] = __ldcg(T2 + (size_t)i * n2 + 2 * item + 1);
                                    }
                                }
                            }
                        }
#pragma unroll
                        for (int u = 0; u < DWK_W_U; ++u) {
                            if (ri[u] >= 0) {
                                const int i = ri[u];
                                double2 r0, r1;
                                r0.x = fma(-ci[u], rv0.x, v0[u].x); r0.y = fma(-ci[u], rv0.y, v0[u].y);
                                r1.x = fma(-ci[u], rv1.x, v1[u].x); r1.y = fma(-ci[u], rv1.y, v1[u].y);
                                if (isq) {
                                    const double cq = ci[u] * inv;
                                    if (cq_pos == 0) r0.x = cq; else if (cq_pos == 1) r0.y = cq; else if (cq_pos == 2) r1.x = cq; else r1.y = cq;
                                }
                                const bool all_zero = r0.x == 0.0 && r0.y == 0.0 && r1.x == 0.0 && r1.y == 0.0;
                                if (had[u] || !all_zero) {
                                    __stcg(T2 + (size_t)i * n2 + 2 * item, r0);
                                    __stcg(T2 + (size_t)i * n2 + 2 * item + 1, r1);
                                }
                                // the row's mask stays exact for every item the sweep visits
                                unsigned *mw = c.rmask + (size_t)i * rwd + mword;
                                if (all_zero) { if (had[u]) atomicAnd(mw, ~mbit); }
                                else if (!had[u]) atomicOr(mw, mbit);
                            }
                        }
                    }
                }
                if (with_d) {
                    double2 *d2 = reinterpret_cast<double2 *>(c.d);
                    const double cq = dq * inv;
                    for (int e = lane; e < nc; e += 32) {
                        const int item = c.acol[e];
                        const double2 rv0 = reinterpret_cast<const double2 *>(c.scr)[2 * e];
                        const double2 rv1 = reinterpret_cast<const double2 *>(c.scr)[2 * e + 1];
                        double2 r0 = d2[2 * item], r1 = d2[2 * item + 1];
                        r0.x = fma(-dq, rv0.x, r0.x); r0.y = fma(-dq, rv0.y, r0.y);
                        r1.x = fma(-dq, rv1.x, r1.x); r1.y = fma(-dq, rv1.y, r1.y);
                        if (item == tq) {
                            if (cq_pos == 0) r0.x = cq; else if (cq_pos == 1) r0.y = cq; else if (cq_pos == 2) r1.x = cq; else r1.y = cq;
                        }
                        d2[2 * item] = r0; d2[2 * item + 1] = r1;
                    }
                }
                __syncwarp();
            }
            if (X.bail_after > 0 && n_piv >= (unsigned long long)X.bail_after) { bail = true; kind = K_FAIL; break; }
            if (infeasible) { kind = K_PIVOT; break; }         // phase 1: re-scan and re-price
            if (++iter > max_iter) { status = 1; kind = K_FAIL; break; }
        }
        if (kind == K_PIVOT || kind == K_REDO) continue;
        if (kind == K_FAIL) break;
        // ---- K_NONE
        if (infeasible) { status = 4; break; }
        // nothing moved since the last optimal solve of this block: x_k is still what B.x holds
        if ((n_piv + n_flip) == 0 && !P.initial && !P.force_full) { status = 5; x_in_scr = false; break; }
        // claimed optimal: x into scr, auxiliary values into colv (by variable id)
        for (int j = lane; j < n; j += 32) {
            const int v = c.nbv[j];
            const signed char st = c.cstat[j];
            const double xv = st == ST_FREE ? 0.0 : (st == ST_UP ? __ldg(c.gup + v) : __ldg(c.glo + v));
            if (v >= m) c.scr[v - m] = xv; else c.colv[v] = xv;
        }
        for (int i = lane; i < m; i += 32) {
            const int v = c.head[i];
            if (v >= m) c.scr[v - m] = c.beta[i]; else c.colv[v] = c.beta[i];
        }
        __syncwarp();
        // residual of the state against the original rows
        double worst = 0.0;
        {
            const int *arp = B.a_rp, *aci = B.a_ci;
            const double *av = B.a_v;
            for (int i = lane; i < m; i += 32) {
                double sres = 0.0;
                const int e1 = __ldg(arp + i + 1);
                for (int e = __ldg(arp + i); e < e1; ++e) sres = fma(__ldg(av + e), c.scr[__ldg(aci + e)], sres);
                worst = fmax(worst, fabs(sres - c.colv[i]) / (1.0 + fabs(sres)));
            }
        }
        worst = warp_max(worst);
        if (worst > 0.1 * tol_bnd) { bail = true; break; }     // the CTA kernel rebuilds the tableau from A_k
        status = 5;
        break;
    }

    // ---- epilogue
    __syncwarp();
    const bool moved = (n_piv + n_flip) != 0;
    if (bail) {
        // hand the block over: consistent state image back to HBM, saved reduced costs invalidated
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            if (n_piv != 0) bulk_s2g(B.rmask, c.rmask, k_bytes);
            if (bulk_beta) bulk_s2g(B.beta, c.beta, b_bytes);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (moved) {
            int *gh = B.head, *gn = B.nbv;
            signed char *gv = B.vstat;
            for (int i = lane; i < m; i += 32) gh[i] = c.head[i];
            for (int j = lane; j < n; j += 32) gn[j] = c.nbv[j];
            for (int v = lane; v < m + n; v += 32) gv[v] = c.vstat[v];
            if (!bulk_beta) { double *gb = B.beta; for (int i = lane; i < m; i += 32) gb[i] = c.beta[i]; }
        }
        if (lane == 0) {
            B.dflag[0] = 0;
            X.moved[k] = moved ? 1 : 0;
            X.redo_list[atomicAdd(X.redo_count, 1)] = k;
            P.counters[8 * (size_t)k + 0] += n_piv;
            P.counters[8 * (size_t)k + 1] += n_flip;
            P.counters[8 * (size_t)k + 3] += n_p1;
            P.counters[8 * (size_t)k + 4] += n_rows;
            P.counters[8 * (size_t)k + 5] += n_cells;
            P.counters[8 * (size_t)k + 6] += n_price;
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
        return;
    }
    const bool changed = moved || P.initial != 0 || P.force_full != 0 || status != 5;
    if (status != 5) {
        for (int j = lane; j < n; j += 32) {
            const int v = c.nbv[j];
            if (v >= m) {
                const signed char st = c.cstat[j];
                c.scr[v - m] = st == ST_FREE ? 0.0 : (st == ST_UP ? __ldg(c.gup + v) : __ldg(c.glo + v));
            }
        }
        for (int i = lane; i < m; i += 32) {
            const int v = c.head[i];
            if (v >= m) c.scr[v - m] = c.beta[i];
        }
        __syncwarp();
    }
    if (!changed) {
        // the block was already optimal under the new prices: x_k, D_k x_k, c_k.x_k and the whole state in HBM are
        // still those of the previous round; only the priced optimum moves
        double zo = 0.0;
        {
            const double *xk = B.x;
            double *dsave = B.dsave, *cssave = B.cssave;
            constexpr int U = 8;
            for (int j0 = lane; j0 < n; j0 += 32 * U) {
                double cv[U], xv[U], sv[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int j = j0 + u * 32;
                    if (j < n) { cv[u] = __ldcg(c.cs + j); xv[u] = x_in_scr ? c.scr[j] : __ldcg(xk + j); sv[u] = __ldcg(cssave + j); }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int j = j0 + u * 32;
                    if (j < n) {
                        zo = fma(cv[u], xv[u], zo);
                        dsave[j] = c.d[j];
                        if (d_full || cost_delta(cv[u], sv[u]) != 0.0) cssave[j] = cv[u];
                    }
                }
            }
        }
        zo = warp_sum(zo);
        if (lane == 0) {
            P.obj[k] = zo; P.status[k] = 5; P.last_iters[k] = 0;
            if (P.m_obj != nullptr) { P.m_obj[k] = zo; P.m_status[k] = 5; P.counters[8 * (size_t)k + 7] += 12; }   // the rest of the mailbox still holds
            P.counters[8 * (size_t)k + 6] += n_price;
            B.dflag[0] = d_full ? 1 : dflag_in + 1;
        }
        return;
    }
    if (status == 5 && d_valid) {
        double *dsave = B.dsave, *cssave = B.cssave;
        for (int j = lane; j < n; j += 32) {
            dsave[j] = c.d[j];
            const double cv = __ldcg(c.cs + j);
            if (d_full || cost_delta(cv, cssave[j]) != 0.0) cssave[j] = cv;
        }
        if (lane == 0) B.dflag[0] = d_full ? 1 : dflag_in + 1;
    } else if (lane == 0) B.dflag[0] = 0;
    // state out first (async for the bulk part), then the proposal while the copies fly
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
        if (n_piv != 0) bulk_s2g(B.rmask, c.rmask, k_bytes);
        if (bulk_beta) bulk_s2g(B.beta, c.beta, b_bytes);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    {
        int *gh = B.head, *gn = B.nbv;
        signed char *gv = B.vstat;
        if (n_piv != 0) {
            for (int i = lane; i < m; i += 32) gh[i] = c.head[i];
            for (int j = lane; j < n; j += 32) gn[j] = c.nbv[j];
        }
        for (int v = lane; v < m + n; v += 32) gv[v] = c.vstat[v];
        if (!bulk_beta) { double *gb = B.beta; for (int i = lane; i < m; i += 32) gb[i] = c.beta[i]; }
    }
    double zo = 0.0, zc = 0.0;
    {
        double *xk = B.x;
        const double *cm = B.c_master;
        for (int j = lane; j < n; j += 32) {
            const double xj = c.scr[j];
            xk[j] = xj;
            zo = fma(__ldcg(c.cs + j), xj, zo);
            zc = fma(__ldg(cm + j), xj, zc);
        }
    }
    zo = warp_sum(zo);
    zc = warp_sum(zc);
    double *yv = P.val + (size_t)k * L;
    int *yi = P.ind + (size_t)k * L;
    {
        const int *drp = B.dr_ptr, *drc = B.dr_col;
        const double *drv = B.dr_val;
        for (int i = lane; i < L; i += 32) {
            double y = 0.0;
            const int e1 = __ldg(drp + i + 1);
            for (int e = __ldg(drp + i); e < e1; ++e)
                y = __dadd_rn(y, __dmul_rn(c.scr[__ldg(drc + e)], __ldg(drv + e)));
            yv[i] = y;
        }
    }
    __syncwarp();
    int base = 0;
    for (int i0 = 0; i0 < L; i0 += 32) {
        const int i = i0 + lane;
        const double y = i < L ? yv[i] : 0.0;
        const bool nz = (i < L) && (y != 0.0);
        const unsigned bal = __ballot_sync(0xffffffffu, nz);
        __syncwarp();
        if (nz) {
            const int pos = base + __popc(bal & lt_mask);
            yv[pos] = y;
            yi[pos] = i + 1;
            if (P.m_val != nullptr) { P.m_val[(size_t)k * L + pos] = y; P.m_ind[(size_t)k * L + pos] = i + 1; }
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
        P.counters[8 * (size_t)k + 3] += n_p1;
        P.counters[8 * (size_t)k + 4] += n_rows;
        P.counters[8 * (size_t)k + 5] += n_cells;
        P.counters[8 * (size_t)k + 6] += n_price;
        P.last_iters[k] = (long long)(n_piv + n_flip);
        if (P.m_obj != nullptr) {
            P.m_obj[k] = zo; P.m_cost[k] = zc; P.m_status[k] = status; P.m_len[k] = base;
            P.counters[8 * (size_t)k + 7] += 24ull + 12ull * (unsigned long long)base;
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

}  // namespace dwk
