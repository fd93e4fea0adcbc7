// simplex_smem.cuh — shared-memory-resident CTA-per-block simplex (config A class blocks).
//
// Same algorithm and tolerances as simplex_cta.cuh (the generic kernel), specialised for blocks
// whose whole working set fits one CTA's shared memory (three CTAs per SM at 64 x 128):
//   * the block's state image [T | beta head nbv vstat] moves HBM<->SMEM with TMA bulk copies
//     (cp.async.bulk + mbarrier): one elected thread issues them, nobody spins on LDG latency; T
//     is written back only if the basis changed, and nothing at all is written back (not even the
//     proposal) when the block was already optimal under the new prices;
//   * the reduced-cost row d is row m of the shared tableau, so one rank-1 sweep updates both;
//   * bounds are kept per ROW (basic variable of row i) and per COLUMN (nonbasic variable of
//     column j) and swapped at a basis change, so pricing and the ratio test have no dependent
//     gathers;
//   * pricing (argmax |d_j|) and the two-pass Harris ratio test run inside warp 0: one reciprocal
//     per blocking row, exact ratios cached for pass 2, `redux.sync` max/min on the two 32-bit
//     halves of the (non-negative) keys instead of 64-bit shuffle trees; bound flips loop inside
//     warp 0 without any block barrier; a basis change costs two __syncthreads;
//   * while it reads the pivot column, warp 0 builds the list of rows whose entry is non-zero
//     (ballot + popc); the rank-1 sweep only visits those rows (network LPs: ~10 % of the rows),
//     one row per warp at a time, 16-byte (double2) shared accesses.
// Reference being replaced: one warm-started glp_simplex call per block and round
// (/root/reference/src/dw_subprob.c:329) plus the pricing / column code around it (:291-321,
// :453-540).
#pragma once
#include <stdint.h>

#include "simplex_cta.cuh"

#ifndef DWK_TG_SB
#define DWK_TG_SB 4         // 32-byte items per thread in flight in the sparse sweep (TG)
#endif
#ifndef DWK_TG_RB
#define DWK_TG_RB 4         // double2 per thread in flight while staging the pivot row (TG)
#endif
#ifndef DWK_TG_FUSED
#define DWK_TG_FUSED 1      // 1: sparse pivots fetch the pivot row inside the sweep's first batch (one HBM round trip less)
#endif
#ifndef DWK_DEAD_COLS
#define DWK_DEAD_COLS 1     // 1: a fixed variable that leaves the basis can never come back (pricing skips ST_FIX), so the column
                            // slot it moves into is written as exact zeros instead of the transformed column; nothing ever reads
                            // it for a decision, and it drops out of the row masks, the item lists and all later sweeps
                            // (equality-constrained blocks: every auxiliary variable the cold solve drives out, half the slots)
#endif
#ifndef DWK_TGW_OCC
#define DWK_TGW_OCC 4       // CTAs per SM the 128- and 256-thread global-tableau kernels are compiled for
#endif
#ifndef DWK_TG_OCC
#define DWK_TG_OCC 6        // CTAs per SM the 64-thread global-tableau kernel is compiled for (measured best of 3..8: no spills at 167 registers)
#endif

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
// leading dimension of the tableau: shared (TG = false) or global-memory resident (TG = true)
__host__ __device__ inline int tab_ld(int n, bool tg) { return tg ? ((n + 3) & ~3) : smem2_ld(n); }
// row-item mask: one bit per 32-byte item (4 doubles) of a global-memory tableau row
__host__ __device__ inline int rmask_words(int n) { return (((n + 3) & ~3) / 4 + 31) / 32; }
__host__ __device__ inline size_t rmask_bytes(int m, int n) { return align16(sizeof(unsigned) * (size_t)m * rmask_words(n)); }

__host__ __device__ inline size_t smem2_bytes(int m, int n, bool tg = false)
{
    const size_t ld = tab_ld(n, tg);
    return align16(sizeof(double) * (size_t)(tg ? 1 : m + 1) * ld)     // (T and) the d row
         + pst_bytes(m, n)
         // rlo rup (clo cup): column bounds and the priced costs cs are only touched a few times per
         // pivot / solve; the global-tableau kernel reads them from HBM/L2 instead (12 KB less per CTA
         // at 256 x 512: 9 CTAs per SM instead of 6)
         + align16(sizeof(double) * 2 * (size_t)(tg ? m : m + n))
         + align16(sizeof(double) * ld)                       // rowp (also phase-1 prices, x staging)
         + (tg ? 0 : align16(sizeof(double) * (size_t)n))     // cs
         + align16(sizeof(double) * (size_t)(m + 1))          // colq
         + align16(sizeof(double) * (size_t)m)                // w (phase-1 weights / cached ratios)
         + align16((size_t)n)                                 // cstat
         + align16(sizeof(unsigned short) * (size_t)(m + 1))  // active rows
         + align16(sizeof(unsigned short) * (ld / 2 + 1))     // active column items of the pivot row
         + (tg ? align16(sizeof(unsigned short) * (size_t)(m + 32 * 8)) : 0)    // per-warp column segments
         + (tg ? rmask_bytes(m, n) : 0);                                        // row-item masks
}

enum { K_NONE = 0, K_PIVOT = 1, K_FAIL = 2, K_REDO = 3 };

// Second launch of a round over the blocks the warp kernel (simplex_warp.cuh) handed over: the list length is read
// on the device, and a block that had already moved must not take the "nothing changed" exit.
struct Continuation {
    const int *dyn_count;           // nullptr: the launch covers `count` blocks
    const unsigned char *moved;     // by block id; nullptr: none
    const unsigned char *reb;       // by block id: the tableau in HBM is not valid, rebuild it from A_k for the basis of the
                                    // state image before anything else (blocks handed over by path 5); nullptr: none
    const int *full_flag;           // delta rounds (delta_tg.cuh): non-zero on the device = this launch covers ALL `count`
    const int *full_list;           // blocks of full_list instead of the first *dyn_count entries of `list`; nullptr: n/a
};

// tableau accessors: global-memory tableaux (TG) bypass L1 (.cg) — rows are streamed once per basis
// change and the strided column read touches one sector per row, so L1 only adds thrash
template <bool TG> __device__ __forceinline__ double ld_t(const double *p) { return TG ? __ldcg(p) : *p; }
template <bool TG> __device__ __forceinline__ double2 ld_t2(const double2 *p) { return TG ? __ldcg(p) : *p; }
template <bool TG> __device__ __forceinline__ void st_t2(double2 *p, double2 v) { if (TG) __stcg(p, v); else *p = v; }


struct Decision {
    int p, q, na, fail_status;
    int nc;          // active column items of the pivot row; -1: the row is dense, sweep whole rows
    int fused;       // 1: rowp has NOT been staged — the sparse sweep fetches row p (first in the list) with its first batch
    int dead;        // 1: the leaving variable is fixed (lo == up), so it can never be priced in again: the column it moves
                     //    into is written as zeros and drops out of the row masks, the item lists and every later sweep
    double inv;
};

// Phase timers (profiling builds only, -DDWK_TIMERS=1): thread 0 of every CTA charges the cycles between two
// marks to the phase that was current; the totals are read with dwgpu_debug_phases().
#ifndef DWK_TIMERS
#define DWK_TIMERS 0
#endif
#if DWK_TIMERS
__device__ unsigned long long g_phase[16];
struct PhaseClock { long long acc[16]; long long last; int cur; };
__device__ __forceinline__ PhaseClock &phase_clock() { __shared__ PhaseClock pc; return pc; }
__device__ __forceinline__ void phase_init()
{
    if (threadIdx.x == 0) { PhaseClock &pc = phase_clock(); for (int i = 0; i < 16; ++i) pc.acc[i] = 0; pc.cur = 0; pc.last = clock64(); }
}
__device__ __forceinline__ void phase(int k)
{
    if (threadIdx.x == 0) { PhaseClock &pc = phase_clock(); const long long t = clock64(); pc.acc[pc.cur] += t - pc.last; pc.last = t; pc.cur = k; }
}
__device__ __forceinline__ void phase_flush()
{
    if (threadIdx.x == 0) {
        phase(15);
        PhaseClock &pc = phase_clock();
        for (int i = 0; i < 15; ++i) if (pc.acc[i]) atomicAdd(&g_phase[i], (unsigned long long)pc.acc[i]);
        atomicAdd(&g_phase[15], 1ull);
    }
}
#else
__device__ __forceinline__ void phase_init() {}
__device__ __forceinline__ void phase(int) {}
__device__ __forceinline__ void phase_flush() {}
#endif

// A column ITEM is the unit the sparse sweep moves: one 32-byte sector (4 doubles) of a global-memory
// tableau row, one double2 of a shared-memory one.
template <bool TG> struct ItemW { static constexpr int D2 = TG ? 2 : 1; };

struct Sm3 {
    int m, n, ld;
    double *T, *drow;
    double *beta;
    int *head, *nbv;
    signed char *vstat;
    double *rlo, *rup, *clo, *cup;       // per row / per column bounds (clo/cup: shared-memory tableau only)
    const double *glo, *gup;             // bounds by variable id in HBM (column bounds when TG)
    double *rowp, *cs, *colq, *w;
    signed char *cstat;
    unsigned short *arow;
    unsigned short *acol;
    unsigned short *arow2;               // per-warp column segments before compaction (TG)
    unsigned *rmask;                     // row-item masks (TG), rwd words per row
    int rwd;
};

// lane holding the maximum of non-negative keys (ties: lowest lane); -1 if no lane is valid
__device__ __forceinline__ int warp_argmax_nonneg(double key, bool valid)
{
    const unsigned hi = valid ? (unsigned)__double2hiint(key) : 0u;
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const bool c1 = valid && hi == mh;
    const unsigned lo = c1 ? (unsigned)__double2loint(key) : 0u;
    const unsigned ml = __reduce_max_sync(0xffffffffu, lo);
    const unsigned b = __ballot_sync(0xffffffffu, c1 && lo == ml);
    return b ? (__ffs(b) - 1) : -1;
}
// minimum of non-negative doubles (+inf allowed)
__device__ __forceinline__ double warp_min_nonneg(double v)
{
    const unsigned hi = (unsigned)__double2hiint(v);
    const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
    const unsigned lo = hi == mh ? (unsigned)__double2loint(v) : 0xffffffffu;
    const unsigned ml = __reduce_min_sync(0xffffffffu, lo);
    return __hiloint2double((int)mh, (int)ml);
}

// Executed by warp 0 only.  Chooses the entering column and the leaving row, applies bound flips
// on the spot, and for a basis change stages everything the rank-1 sweep needs (colq, rowp,
// active rows, beta, head/nbv/vstat/bounds).  dd = prices (row m of T in phase 2; phase-1 prices
// live in rowp and are consumed before rowp is overwritten).
template <bool INF, int MC, int NC, bool TG>
__device__ __forceinline__ int decide_warp0(Sm3 &c, const double *dd, bool d_live, bool &bland, int &degen, int degen_limit,
                            double tol_dj, double tol_bnd, double tol_piv,
                            unsigned long long &n_piv, unsigned long long &n_flip, unsigned long long &n_rows,
                            unsigned long long &n_cells, Decision *dec)
{
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int m = MC ? MC : c.m, n = NC ? NC : c.n, ld = NC ? tab_ld(NC, TG) : c.ld;
    const double harris = 0.5 * tol_bnd;
    for (;;) {
        // ---- pricing
        double key = 0.0; int q = -1;
        for (int j = lane; j < n; j += 32) {
            const double dj = dd[j];
            const signed char st = c.cstat[j];
            const bool ok = (st == ST_LO && dj < -tol_dj) || (st == ST_UP && dj > tol_dj) ||
                            (st == ST_FREE && fabs(dj) > tol_dj);
            const double kj = bland ? 2147483648.0 - (double)c.nbv[j] : fabs(dj);
            if (ok && kj > key) { key = kj; q = j; }
        }
        {
            const int src = warp_argmax_nonneg(key, q >= 0);
            if (src < 0) return K_NONE;
            q = __shfl_sync(0xffffffffu, q, src);
        }
        const double dq = dd[q];
        const double s = dq < 0.0 ? 1.0 : -1.0;
        const signed char stq = c.cstat[q];
        const double loq = TG ? __ldg(c.glo + c.nbv[q]) : c.clo[q];
        const double upq = TG ? __ldg(c.gup + c.nbv[q]) : c.cup[q];
        // ---- ratio test, pass 1: Harris bound; caches exact ratios in w, builds the active-row list
        double tmax = CUDART_INF;
        int na = 0;
        // the column is fetched in batches of 8 rows per lane: all loads of a batch are in flight
        // together (one memory round trip per 256 rows instead of one per 32)
        constexpr int CB = MC ? 8 : 2;
        for (int i00 = 0; i00 < m; i00 += 32 * CB) {
          double tcol[CB];
#pragma unroll
          for (int r = 0; r < CB; ++r) {
              const int i = i00 + r * 32 + lane;
              tcol[r] = i < m ? ld_t<TG>(c.T + (size_t)i * ld + q) : 0.0;
          }
#pragma unroll
          for (int r = 0; r < CB; ++r) {
            const int i0 = i00 + r * 32;
            if (i0 >= m) break;
            const int i = i0 + lane;
            const bool in = i < m;
            double tiq = 0.0, tex = CUDART_INF;
            if (in) {
                tiq = tcol[r];
                c.colq[i] = tiq;
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
                    tmax = fmin(tmax, fmax((slack + delta) * r, 0.0));
                }
                c.w[i] = tex;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, in && tiq != 0.0);
            if (in && tiq != 0.0) c.arow[na + __popc(bal & lt_mask)] = (unsigned short)i;
            na += __popc(bal);
          }
        }
        tmax = warp_min_nonneg(tmax);
        // ---- pass 2: largest |pivot| among rows whose exact ratio is within the bound
        int p = -1;
        if (tmax < CUDART_INF) {
            const double lim = bland ? tmax + 1e-12 * (1.0 + tmax) : tmax;
            double pk = 0.0;
            __syncwarp();
            for (int i = lane; i < m; i += 32) {
                if (c.w[i] <= lim) {
                    const double kk = bland ? 2147483648.0 - (double)c.head[i] : fabs(c.colq[i]);
                    if (kk > pk) { pk = kk; p = i; }
                }
            }
            const int src = warp_argmax_nonneg(pk, p >= 0);
            p = src >= 0 ? __shfl_sync(0xffffffffu, p, src) : -1;
        }
        __syncwarp();
        const double theta = p >= 0 ? fmax(0.0, c.w[p]) : CUDART_INF;
        const double range = (stq == ST_FREE) ? CUDART_INF : upq - loq;
        if (range <= theta) {
            if (!(range < BIG)) { if (lane == 0) dec->fail_status = INF ? 1 : 6; return K_FAIL; }
            for (int i = lane; i < m; i += 32) c.beta[i] = fma(s * range, c.colq[i], c.beta[i]);
            if (lane == 0) {
                const signed char ns = (stq == ST_LO) ? ST_UP : ST_LO;
                c.cstat[q] = ns;
                c.vstat[c.nbv[q]] = ns;
            }
            ++n_flip;
            degen = 0; bland = false;
            __syncwarp();
            if (INF) return K_REDO;     // phase-1 prices depend on beta
            continue;
        }
        if (p < 0) { if (lane == 0) dec->fail_status = INF ? 1 : 6; return K_FAIL; }
        // ---- basis change: stage the sweep
        const double piv = c.colq[p];
        const double inv = __drcp_rn(piv);
        const double xq_new = nb_value(stq, loq, upq) + s * theta;
        const double lov = c.rlo[p], upv = c.rup[p];
        double target;      // the bound the leaving variable was heading for
        {
            double lo = lov, up = upv;
            if (INF) {
                const double bp = c.beta[p];
                if (bp < lo - tol_bnd * (1.0 + fabs(lo))) up = lo;          // rises to its lower bound
                else if (bp > up + tol_bnd * (1.0 + fabs(up))) lo = up;     // falls to its upper bound
            }
            target = (s * piv > 0.0) ? up : lo;
        }
        __syncwarp();
        for (int i = lane; i < m; i += 32)
            c.beta[i] = (i == p) ? xq_new : fma(s * theta, c.colq[i], c.beta[i]);
        {
            const double2 *prow = reinterpret_cast<const double2 *>(c.T + (size_t)p * ld);
            double2 *rp2 = reinterpret_cast<double2 *>(c.rowp);
            constexpr int RB = NC ? 8 : 2;        // 8 x 16 B per lane in flight (a 512-column row in one trip)
            for (int j0 = 0; j0 < (ld >> 1); j0 += 32 * RB) {
                double2 v[RB];
#pragma unroll
                for (int t = 0; t < RB; ++t) {
                    const int j2 = j0 + t * 32 + lane;
                    if (j2 < (ld >> 1)) v[t] = ld_t2<TG>(prow + j2);
                }
#pragma unroll
                for (int t = 0; t < RB; ++t) {
                    const int j2 = j0 + t * 32 + lane;
                    if (j2 < (ld >> 1)) { v[t].x *= inv; v[t].y *= inv; rp2[j2] = v[t]; }
                }
            }
        }
        // non-zero items of the scaled pivot row: T'[i][j] = T[i][j] - colq[i] * rowp[j] leaves every
        // column with rowp[j] == 0 untouched, so only (active row) x (active item) cells are rewritten
        int nc = 0;
        {
            constexpr int IW = ItemW<TG>::D2;
            const int nitems = (ld >> 1) / IW;
            const double2 *rp2 = reinterpret_cast<const double2 *>(c.rowp);
            __syncwarp();
            for (int t0 = 0; t0 < nitems; t0 += 32) {
                const int t = t0 + lane;
                bool nz = false;
                if (t < nitems) {
#pragma unroll
                    for (int u = 0; u < IW; ++u) { const double2 v = rp2[t * IW + u]; nz |= (v.x != 0.0) | (v.y != 0.0); }
                }
                const unsigned bal = __ballot_sync(0xffffffffu, nz);
                if (nz) c.acol[nc + __popc(bal & lt_mask)] = (unsigned short)t;
                nc += __popc(bal);
            }
            if (2 * nc > nitems) nc = -1;       // dense pivot row: whole-row sweeps are cheaper
        }
        if (lane == 0) {
            const signed char ns = (lov == upv) ? ST_FIX : (target == lov ? ST_LO : ST_UP);
            const int vleave = c.head[p], vq = c.nbv[q];
            c.rlo[p] = loq; c.rup[p] = upq;
            if (!TG) { c.clo[q] = lov; c.cup[q] = upv; }
            c.cstat[q] = ns; c.vstat[vleave] = ns; c.vstat[vq] = ST_BASIC;
            c.head[p] = vq; c.nbv[q] = vleave;
            const bool with_d = d_live && dq != 0.0;
            c.colq[m] = with_d ? dq : 0.0;
            if (with_d) c.arow[na] = (unsigned short)m;      // the reduced-cost row rides along
            dec->p = p; dec->q = q; dec->inv = inv; dec->na = na + (with_d ? 1 : 0); dec->nc = nc; dec->fused = 0;
            dec->dead = (DWK_DEAD_COLS && lov == upv) ? 1 : 0;
        }
        ++n_piv;
        n_rows += (unsigned long long)na;
        n_cells += nc < 0 ? (unsigned long long)(na + (d_live && dq != 0.0 ? 1 : 0)) * (unsigned long long)ld
                          : (unsigned long long)(na + (d_live && dq != 0.0 ? 1 : 0)) * (unsigned long long)nc * (2 * ItemW<TG>::D2);
        if (theta <= 1e-11) { if (++degen > degen_limit) bland = true; }
        else { degen = 0; bland = false; }
        return K_PIVOT;
    }
}

// ---- block-cooperative decision for the global-memory tableau (TG) kernel ----------------------------
// With the tableau in HBM a basis change is a chain of dependent memory round trips (column, pivot
// row, sweep); what sits between them must be as short as possible.  Here every thread of the CTA
// takes part: pricing is split over all threads, each warp fetches a contiguous quarter/half of the
// pivot column in ONE batch of loads and compacts its non-zero rows, and the ratio test, the beta
// update and the bookkeeping then run over that short list only (network blocks: ~13 of 256 rows)
// instead of all m rows.  Same pivoting rules as decide_warp0.
struct DecShared {
    double key[8];
    int q[8];
    int cnt[8];
    int kind, bland;
};
enum { K_FLIPPED = 4 };

// Entering column: largest |d_j| among the attractive nonbasic columns (Bland: lowest variable id); all threads take
// part and all get the answer (-1: none).  One block barrier.
template <int NT, int NC>
__device__ __forceinline__ int price_block_tg(const Sm3 &c, const double *dd, bool bland, double tol_dj, DecShared *ds)
{
    constexpr int NW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = NC ? NC : c.n;
    double key = 0.0; int q = -1;
#pragma unroll 8
    for (int j = tid; j < n; j += NT) {
        const double dj = dd[j];
        const signed char st = c.cstat[j];
        const bool ok = (st == ST_LO && dj < -tol_dj) || (st == ST_UP && dj > tol_dj) ||
                        (st == ST_FREE && fabs(dj) > tol_dj);
        const double kj = bland ? 2147483648.0 - (double)c.nbv[j] : fabs(dj);
        if (ok && kj > key) { key = kj; q = j; }
    }
    {
        const int src = warp_argmax_nonneg(key, q >= 0);
        const int sl = src < 0 ? 0 : src;
        const double kw = __shfl_sync(0xffffffffu, key, sl);
        const int qw = __shfl_sync(0xffffffffu, q, sl);
        if (lane == 0) { ds->key[warp] = kw; ds->q[warp] = src < 0 ? -1 : qw; }
    }
    __syncthreads();
    q = -1; key = -1.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        const double kw = ds->key[w];
        const int qw = ds->q[w];
        if (qw >= 0 && kw > key) { key = kw; q = qw; }
    }
    return q;
}

template <bool INF, int NT, int MC, int NC>
__device__ __forceinline__ int decide_block_tg(Sm3 &c, const double *dd, bool d_live, bool &bland, int &degen, int degen_limit,
                            double tol_dj, double tol_bnd, double tol_piv,
                            unsigned long long &n_piv, unsigned long long &n_flip, unsigned long long &n_rows,
                            unsigned long long &n_cells, Decision *dec, DecShared *ds)
{
    constexpr int NW = NT / 32;
    static_assert(NW <= 8, "DecShared holds 8 warp slots");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int m = MC ? MC : c.m, ld = NC ? tab_ld(NC, true) : c.ld;
    const double harris = 0.5 * tol_bnd;
    const int RW = ((m + NW - 1) / NW + 31) & ~31;          // rows of the column each warp fetches
    for (;;) {
        phase(3);
        // ---- pricing, all threads
        const int q = price_block_tg<NT, NC>(c, dd, bland, tol_dj, ds);
        if (q < 0) return K_NONE;                                // block-uniform
        phase(4);
        const double dq = dd[q];
        const double s = dq < 0.0 ? 1.0 : -1.0;
        const signed char stq = c.cstat[q];
        const int vq = c.nbv[q];
        const double loq = __ldg(c.glo + vq), upq = __ldg(c.gup + vq);
        // ---- pivot column: warp w fetches rows [w RW, (w+1) RW) in one batch and lists its non-zeros
        int cnt = 0;
        {
            const int rbeg = warp * RW, rend = min(m, rbeg + RW);
            constexpr int CB = MC ? (MC / NW + 31) / 32 : 4;
            // a row whose mask has no bit for q's item holds an exact zero there: its sector is not fetched
            // (a strided column costs one 32-byte sector per row; a network column has ~5 % non-zeros)
            const unsigned *mq = c.rmask + (q >> 7);
            const unsigned qbit = 1u << ((q >> 2) & 31);
            const int rwd = c.rwd;
            for (int i00 = rbeg; i00 < rend; i00 += 32 * CB) {
                double tcol[CB];
#pragma unroll
                for (int r = 0; r < CB; ++r) {
                    const int i = i00 + r * 32 + lane;
                    tcol[r] = (i < rend && (mq[(size_t)i * rwd] & qbit)) ? __ldcg(c.T + (size_t)i * ld + q) : 0.0;
                }
#pragma unroll
                for (int r = 0; r < CB; ++r) {
                    const int i = i00 + r * 32 + lane;
                    const bool nz = i < rend && tcol[r] != 0.0;
                    const unsigned bal = __ballot_sync(0xffffffffu, nz);
                    if (nz) {
                        c.colq[i] = tcol[r];
                        c.arow2[rbeg + cnt + __popc(bal & lt_mask)] = (unsigned short)i;
                    }
                    cnt += __popc(bal);
                }
            }
            if (lane == 0) ds->cnt[warp] = cnt;
        }
        __syncthreads();
        int na = 0;
        {
            int my_off = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) { if (w == warp) my_off = na; na += ds->cnt[w]; }
            for (int e = lane; e < cnt; e += 32) c.arow[my_off + e] = c.arow2[warp * RW + e];
        }
        __syncthreads();
        phase(5);
        // ---- ratio test, flips and the staging of a basis change over the list: warp 0
        if (warp == 0) {
            int kind = K_PIVOT;
            double tmax = CUDART_INF;
            for (int e = lane; e < na; e += 32) {
                const int i = c.arow[e];
                const double tiq = c.colq[i];
                const double a = s * tiq;
                double lo = c.rlo[i], up = c.rup[i];
                const double b = c.beta[i];
                if (INF) {
                    if (b < lo - tol_bnd * (1.0 + fabs(lo))) { up = lo; lo = -CUDART_INF; }
                    else if (b > up + tol_bnd * (1.0 + fabs(up))) { lo = up; up = CUDART_INF; }
                }
                bool blocks = false; double slack = 0.0, target = 0.0, tex = CUDART_INF;
                if (a > tol_piv) { if (up < BIG) { blocks = true; slack = up - b; target = up; } }
                else if (a < -tol_piv) { if (lo > -BIG) { blocks = true; slack = b - lo; target = lo; } }
                if (blocks) {
                    const double r = __drcp_rn(fabs(a));
                    tex = slack * r;
                    const double delta = bland ? 0.0 : harris * (1.0 + fabs(target));
                    tmax = fmin(tmax, fmax((slack + delta) * r, 0.0));
                }
                c.w[e] = tex;                                    // cached by list position
            }
            tmax = warp_min_nonneg(tmax);
            int p = -1, pe = -1;
            if (tmax < CUDART_INF) {
                const double lim = bland ? tmax + 1e-12 * (1.0 + tmax) : tmax;
                double pk = 0.0;
                __syncwarp();
                for (int e = lane; e < na; e += 32) {
                    if (c.w[e] <= lim) {
                        const int i = c.arow[e];
                        const double kk = bland ? 2147483648.0 - (double)c.head[i] : fabs(c.colq[i]);
                        if (kk > pk) { pk = kk; p = i; pe = e; }
                    }
                }
                const int src = warp_argmax_nonneg(pk, p >= 0);
                const int sl = src < 0 ? 0 : src;
                p = __shfl_sync(0xffffffffu, p, sl);
                pe = __shfl_sync(0xffffffffu, pe, sl);
                if (src < 0) { p = -1; pe = -1; }
            }
            __syncwarp();
            const double theta = p >= 0 ? fmax(0.0, c.w[pe]) : CUDART_INF;
            const double range = (stq == ST_FREE) ? CUDART_INF : upq - loq;
            if (range <= theta) {
                if (!(range < BIG)) { if (lane == 0) dec->fail_status = INF ? 1 : 6; kind = K_FAIL; }
                else {
                    for (int e = lane; e < na; e += 32) {
                        const int i = c.arow[e];
                        c.beta[i] = fma(s * range, c.colq[i], c.beta[i]);
                    }
                    if (lane == 0) {
                        const signed char ns = (stq == ST_LO) ? ST_UP : ST_LO;
                        c.cstat[q] = ns;
                        c.vstat[vq] = ns;
                    }
                    ++n_flip;
                    degen = 0; bland = false;
                    kind = INF ? K_REDO : K_FLIPPED;             // phase-1 prices depend on beta
                }
            } else if (p < 0) {
                if (lane == 0) dec->fail_status = INF ? 1 : 6;
                kind = K_FAIL;
            } else {
                const double piv = c.colq[p];
                const double inv = __drcp_rn(piv);
                const double xq_new = nb_value(stq, loq, upq) + s * theta;
                const double lov = c.rlo[p], upv = c.rup[p];
                double target;
                {
                    double lo = lov, up = upv;
                    if (INF) {
                        const double bp = c.beta[p];
                        if (bp < lo - tol_bnd * (1.0 + fabs(lo))) up = lo;
                        else if (bp > up + tol_bnd * (1.0 + fabs(up))) lo = up;
                    }
                    target = (s * piv > 0.0) ? up : lo;
                }
                __syncwarp();
                for (int e = lane; e < na; e += 32) {
                    const int i = c.arow[e];
                    c.beta[i] = (i == p) ? xq_new : fma(s * theta, c.colq[i], c.beta[i]);
                }
                __syncwarp();
                if (lane == 0) {
                    c.arow[pe] = c.arow[0]; c.arow[0] = (unsigned short)p;      // the sweep wants the pivot row first
                    const signed char ns = (lov == upv) ? ST_FIX : (target == lov ? ST_LO : ST_UP);
                    const int vleave = c.head[p];
                    c.rlo[p] = loq; c.rup[p] = upq;
                    c.cstat[q] = ns; c.vstat[vleave] = ns; c.vstat[vq] = ST_BASIC;
                    c.head[p] = vq; c.nbv[q] = vleave;
                    const bool with_d = d_live && dq != 0.0;
                    c.colq[m] = with_d ? dq : 0.0;
                    if (with_d) c.arow[na] = (unsigned short)m;      // the reduced-cost row rides along
                    dec->p = p; dec->q = q; dec->inv = inv; dec->na = na + (with_d ? 1 : 0);
                    dec->dead = (DWK_DEAD_COLS && lov == upv) ? 1 : 0;
                }
                ++n_piv;
                n_rows += (unsigned long long)na;
                if (theta <= 1e-11) { if (++degen > degen_limit) bland = true; }
                else { degen = 0; bland = false; }
            }
            if (lane == 0) { ds->kind = kind; ds->bland = bland ? 1 : 0; }
        }
        __syncthreads();
        const int kind = ds->kind;
        bland = ds->bland != 0;
        if (kind == K_FLIPPED) continue;
        if (kind != K_PIVOT) return kind;
        phase(6);
        // ---- sparse pivot row (the usual case): its item list is its mask — exact whenever the row was last
        // written by a sparse sweep — and the row itself is fetched by the sweep's first batch together with the
        // rows it updates, so a pivot is two dependent HBM round trips (column, sweep) instead of three
        if (DWK_TG_FUSED) {
            constexpr int IW = ItemW<true>::D2;
            const int p = dec->p, rwd = c.rwd;
            const int nitems = (ld >> 1) / IW;
            const unsigned *mp = c.rmask + (size_t)p * rwd;
            int ncm = 0;
            for (int w_ = 0; w_ < rwd; ++w_) ncm += __popc(mp[w_]);
            if (2 * ncm <= nitems && ncm <= NT * DWK_TG_SB) {        // block-uniform; the row fits the sweep's first batch
                if (warp == 0) {
                    int nc = 0;
                    for (int w_ = 0; w_ < rwd; ++w_) {
                        const unsigned bits = mp[w_];
                        if ((bits >> lane) & 1u) c.acol[nc + __popc(bits & lt_mask)] = (unsigned short)(w_ * 32 + lane);
                        nc += __popc(bits);
                    }
                    n_cells += (unsigned long long)dec->na * (unsigned long long)nc * (2 * IW);
                    if (lane == 0) { dec->nc = nc; dec->fused = 1; }
                }
                __syncthreads();
                return K_PIVOT;
            }
        }
        // ---- pivot row -> rowp (scaled): only the 32-byte items whose mask bit is set are fetched (a network
        // row has ~25 of 128), everything else is written as zero; warp 0 turns the mask into the item list
        {
            constexpr int IW = ItemW<true>::D2;
            const int p = dec->p;
            const double inv = dec->inv;
            const int nitems = (ld >> 1) / IW;
            const unsigned *mp = c.rmask + (size_t)p * c.rwd;
            const double2 *prow = reinterpret_cast<const double2 *>(c.T + (size_t)p * ld);
            double2 *rp2 = reinterpret_cast<double2 *>(c.rowp);
            constexpr int RB = NC ? ((NC / 4 + NT - 1) / NT < DWK_TG_RB ? (NC / 4 + NT - 1) / NT : DWK_TG_RB) : 2;
            for (int t0 = 0; t0 < nitems; t0 += NT * RB) {
                double2 v[RB][IW];
                bool on[RB];
#pragma unroll
                for (int u = 0; u < RB; ++u) {
                    const int t = t0 + u * NT + tid;
                    on[u] = t < nitems && ((mp[t >> 5] >> (t & 31)) & 1u);
                    if (on[u]) {
#pragma unroll
                        for (int x = 0; x < IW; ++x) v[u][x] = __ldcg(prow + t * IW + x);
                    }
                }
#pragma unroll
                for (int u = 0; u < RB; ++u) {
                    const int t = t0 + u * NT + tid;
                    if (t < nitems) {
#pragma unroll
                        for (int x = 0; x < IW; ++x) {
                            double2 r = make_double2(0.0, 0.0);
                            if (on[u]) { r.x = v[u][x].x * inv; r.y = v[u][x].y * inv; }
                            rp2[t * IW + x] = r;
                        }
                    }
                }
            }
        }
        __syncthreads();
        if (warp == 0) {
            // exact non-zero items of the scaled pivot row -> acol; the row's mask is refreshed to exactly
            // that pattern (the sweeps keep the masks of the rows they touch exact / a superset themselves)
            constexpr int IW = ItemW<true>::D2;
            const int p = dec->p;
            const int nitems = (ld >> 1) / IW, rwd = c.rwd;
            const double2 *rp2 = reinterpret_cast<const double2 *>(c.rowp);
            unsigned *mp = c.rmask + (size_t)p * rwd;
            int nc = 0;
            for (int t0 = 0; t0 < nitems; t0 += 32) {
                const int t = t0 + lane;
                bool nz = false;
                if (t < nitems) {
#pragma unroll
                    for (int u = 0; u < IW; ++u) { const double2 v = rp2[t * IW + u]; nz |= (v.x != 0.0) | (v.y != 0.0); }
                }
                const unsigned bal = __ballot_sync(0xffffffffu, nz);
                if (nz) c.acol[nc + __popc(bal & lt_mask)] = (unsigned short)t;
                nc += __popc(bal);
                if (lane == 0) mp[t0 >> 5] = bal;
            }
            __syncwarp();
            if (2 * nc > nitems) nc = -1;       // dense pivot row: whole-row sweeps are cheaper
            const int na_all = dec->na;
            n_cells += nc < 0 ? (unsigned long long)na_all * (unsigned long long)ld
                              : (unsigned long long)na_all * (unsigned long long)nc * (2 * IW);
            if (lane == 0) { dec->nc = nc; dec->fused = 0; }
        }
        __syncthreads();
        return K_PIVOT;
    }
}

// Change of a priced cost that the saved reduced costs still have to absorb.  Differences at the
// noise level of the master's duals (<= 1e-11 relative) are NOT applied and the saved cost is left
// as it is, so they keep accumulating until they matter; the reduced costs used for pricing are
// then off by at most ~1e-11 |T|_1, two orders below tol_dj.
__device__ __forceinline__ double cost_delta(double cs_new, double cs_saved)
{
    const double d = cs_new - cs_saved;
    return fabs(d) > 1e-11 * (1.0 + fabs(cs_new)) ? d : 0.0;
}

// the reduced-cost row's share of a sparse rank-1 sweep (it lives in shared memory): d'[j] = d[j] - d_q rowp[j] over the
// pivot row's items, d'[q] = -d_q / pivot
template <int NT, bool TG>
__device__ __forceinline__ void sweep_drow(Sm3 &c, int q, double inv, int nc, bool dead)
{
    constexpr int IW = ItemW<TG>::D2;
    const double2 *rowp2 = reinterpret_cast<const double2 *>(c.rowp);
    double2 *d2 = reinterpret_cast<double2 *>(c.drow);
    const double dq = c.colq[c.m];
    const double cq = dead ? 0.0 : dq * inv;
    for (int t = threadIdx.x; t < nc; t += NT) {
        const int j = (int)c.acol[t] * IW;
#pragma unroll
        for (int u = 0; u < IW; ++u) {
            const double2 ra = rowp2[j + u];
            const int jq = q - 2 * (j + u);
            double2 r = d2[j + u];
            r.x = fma(-dq, ra.x, r.x);
            r.y = fma(-dq, ra.y, r.y);
            if (jq == 0) r.x = cq;
            if (jq == 1) r.y = cq;
            d2[j + u] = r;
        }
    }
}

// rank-1 sweep over (active rows) x (active column items): every load of a batch is issued before the
// first dependent instruction, so a typical network pivot (13 rows x 30 items) is ONE memory round trip.
// The reduced-cost row (last list entry when it rides along) lives in shared memory and is done apart,
// so the tableau loop is pure global (TG) or pure shared traffic with 32-bit cell offsets.
template <int NT, int NC, bool TG>
__device__ __forceinline__ void sweep_sparse(Sm3 &c, int p, int q, double inv, int na, int nc, bool fused, bool dead)
{
    const double invq = dead ? 0.0 : inv;                  // what column q receives: nothing when the leaving variable is fixed
    constexpr int IW = ItemW<TG>::D2;
    constexpr int SB = TG ? DWK_TG_SB : 2;                 // items per thread in flight
    const int nc2 = (NC ? tab_ld(NC, TG) : c.ld) >> 1;    // row pitch in double2
    double2 *T2 = reinterpret_cast<double2 *>(c.T);
    const double2 *rowp2 = reinterpret_cast<const double2 *>(c.rowp);
    const bool with_d = na > 0 && c.arow[na - 1] == c.m;
    const int nat = with_d ? na - 1 : na;
    const int total = nat * nc;
    // e / nc by multiplication: ceil(2^32 / nc) is exact for e * nc < 2^32 (a generic 32-bit division costs ~25 instructions
    // per element, a tenth of what a warp executes per basis change)
    const unsigned nc_magic = nc > 1 ? 0xffffffffu / (unsigned)nc + 1u : 0u;
    for (int e0 = 0; e0 < total; e0 += NT * SB) {
        double2 v[SB][IW];
        int off[SB];                                       // double2 offset of the item; -1: none
        bool had[SB];                                      // TG: the row's mask knew the item (else it is a zero that was never stored)
#pragma unroll
        for (int t = 0; t < SB; ++t) {
            const int e = e0 + t * NT + threadIdx.x;
            off[t] = -1; had[t] = true;
            if (e < total) {
                const int a = nc > 1 ? (int)__umulhi((unsigned)e, nc_magic) : e;
                const int i = c.arow[a], item = c.acol[e - a * nc];
                off[t] = i * nc2 + item * IW;
                if (TG) had[t] = (c.rmask[(size_t)i * c.rwd + (item >> 5)] >> (item & 31)) & 1u;
#pragma unroll
                for (int u = 0; u < IW; ++u) v[t][u] = had[t] ? ld_t2<TG>(T2 + off[t] + u) : make_double2(0.0, 0.0);
            }
        }
        if (TG && fused && e0 == 0) {
            // the list starts with the pivot row: elements [0, nc) are its items, still unscaled in v — scale
            // them into rowp for everybody (block-uniform branch)
            double2 *rw2 = reinterpret_cast<double2 *>(c.rowp);
#pragma unroll
            for (int t = 0; t < SB; ++t) {
                const int e = t * NT + threadIdx.x;
                if (e < nc) {
                    const int j = (int)c.acol[e] * IW;
#pragma unroll
                    for (int u = 0; u < IW; ++u) rw2[j + u] = make_double2(v[t][u].x * inv, v[t][u].y * inv);
#if DWK_TIMERS
                    {   // packing statistics of the pivot row: non-zero doubles per 32-byte item the sweep moves
                        int nzd = 0;
#pragma unroll
                        for (int u = 0; u < IW; ++u) nzd += (v[t][u].x != 0.0) + (v[t][u].y != 0.0);
                        atomicAdd(&g_phase[12], (unsigned long long)nzd);
                        atomicAdd(&g_phase[13], 1ull);
                    }
#endif
                }
            }
            __syncthreads();
        }
#pragma unroll
        for (int t = 0; t < SB; ++t) {
            if (off[t] >= 0) {
                const int i = off[t] / nc2;
                const int j = off[t] - i * nc2;
                const double ci = c.colq[i];
                const double cq = dead ? 0.0 : ci * inv;
                const bool is_p = i == p;
                bool all_zero = true;
                double2 res[IW];
#pragma unroll
                for (int u = 0; u < IW; ++u) {
                    const double2 ra = rowp2[j + u];
                    const int jq = q - 2 * (j + u);
                    double2 r;
                    r.x = is_p ? -ra.x : fma(-ci, ra.x, v[t][u].x);
                    r.y = is_p ? -ra.y : fma(-ci, ra.y, v[t][u].y);
                    if (jq == 0) r.x = is_p ? invq : cq;
                    if (jq == 1) r.y = is_p ? invq : cq;
                    all_zero = all_zero && r.x == 0.0 && r.y == 0.0;
                    res[u] = r;
                }
                if (!TG || had[t] || !all_zero) {
#pragma unroll
                    for (int u = 0; u < IW; ++u) st_t2<TG>(T2 + off[t] + u, res[u]);
                }
                if (TG) {
                    // the row's mask stays exact for every item the sweep visits: an item that cancelled to zero
                    // (network tableaux do that a lot) leaves it, a fill-in enters it
                    const int item = j / IW;
                    unsigned *mw = c.rmask + (size_t)i * c.rwd + (item >> 5);
                    if (all_zero) { if (had[t]) atomicAnd(mw, ~(1u << (item & 31))); }
                    else if (!had[t]) atomicOr(mw, 1u << (item & 31));
                }
            }
        }
    }
    if (with_d) sweep_drow<NT, TG>(c, q, inv, nc, dead);
}

// rank-1 sweep over the active rows only (row m = reduced costs when it is in the list)
template <int NT, int NC, bool TG>
__device__ __forceinline__ void sweep_active(Sm3 &c, int p, int q, double inv, int na, bool dead)
{
    const double invq = dead ? 0.0 : inv;
    constexpr int NW = NT / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nc2 = (NC ? tab_ld(NC, TG) : c.ld) >> 1;    // row pitch in double2
    const int nc2a = TG ? nc2 : ((NC ? NC : c.n) + 1) >> 1;   // double2 columns that can be non-zero (TG: whole items, padding included)
    double2 *T2 = reinterpret_cast<double2 *>(c.T);
    double2 *d2 = reinterpret_cast<double2 *>(c.drow);
    const double2 *rowp2 = reinterpret_cast<const double2 *>(c.rowp);
    for (int a = warp; a < na; a += NW) {
        const int i = c.arow[a];
        const double ci = c.colq[i];
        const double cq = dead ? 0.0 : ci * inv;
        double2 *row = (i == c.m) ? d2 : T2 + (size_t)i * nc2;
        const bool is_p = i == p;
        if (TG) {
            // global-memory tableau: issue every load of (up to) 256 double2 of the row before the
            // first dependent instruction — the sweep is a chain of DRAM round trips otherwise
            // A cell of a global row exists only if the row's mask has its item (2 double2): everything else
            // reads as zero without a load, and a fill-in adds the item to the mask (one warp owns the row).
            constexpr int SB = NC ? 8 : 2;
            const bool glob = i != c.m;
            unsigned *mi = glob && c.rmask != nullptr ? c.rmask + (size_t)i * c.rwd : nullptr;
            for (int j0 = 0; j0 < nc2a; j0 += 32 * SB) {
                double2 v[SB];
                bool had[SB], nzv[SB];
#pragma unroll
                for (int t = 0; t < SB; ++t) {
                    const int j = j0 + t * 32 + lane;
                    had[t] = true; nzv[t] = false;
                    if (j < nc2a) {
                        if (mi != nullptr) had[t] = (mi[j >> 6] >> ((j >> 1) & 31)) & 1u;
                        v[t] = !had[t] ? make_double2(0.0, 0.0) : (glob ? __ldcg(row + j) : row[j]);
                    }
                }
                __syncwarp();                              // all mask reads of the batch before its updates
#pragma unroll
                for (int t = 0; t < SB; ++t) {
                    const int j = j0 + t * 32 + lane;
                    if (j < nc2a) {
                        const double2 ra = rowp2[j];
                        const int jq = q - 2 * j;
                        double2 r;
                        r.x = is_p ? -ra.x : fma(-ci, ra.x, v[t].x);
                        r.y = is_p ? -ra.y : fma(-ci, ra.y, v[t].y);
                        if (jq == 0) r.x = is_p ? invq : cq;
                        if (jq == 1) r.y = is_p ? invq : cq;
                        nzv[t] = r.x != 0.0 || r.y != 0.0;
                        v[t] = r;
                    }
                }
#pragma unroll
                for (int t = 0; t < SB; ++t) {
                    const int j = j0 + t * 32 + lane;
                    // a new item: both halves must be written (the partner half may hold a real zero)
                    const bool partner_nz = __shfl_xor_sync(0xffffffffu, nzv[t] ? 1 : 0, 1) != 0;
                    if (j < nc2a) {
                        if (glob) { if (had[t] || nzv[t] || partner_nz) __stcg(row + j, v[t]); } else row[j] = v[t];
                        if (mi != nullptr && !had[t] && nzv[t]) atomicOr(mi + (j >> 6), 1u << ((j >> 1) & 31));
                    }
                }
            }
        } else
        for (int j0 = 0; j0 < nc2a; j0 += 64) {
            const int ja = j0 + lane, jb = j0 + 32 + lane;
            const bool oa = ja < nc2a, ob = jb < nc2a;
            double2 va, vb, ra, rb;
            if (oa) { va = row[ja]; ra = rowp2[ja]; }
            if (ob) { vb = row[jb]; rb = rowp2[jb]; }
            if (oa) {
                const int jq = q - 2 * ja;
                double2 r;
                r.x = is_p ? -ra.x : fma(-ci, ra.x, va.x);
                r.y = is_p ? -ra.y : fma(-ci, ra.y, va.y);
                if (jq == 0) r.x = is_p ? invq : cq;
                if (jq == 1) r.y = is_p ? invq : cq;
                row[ja] = r;
            }
            if (ob) {
                const int jq = q - 2 * jb;
                double2 r;
                r.x = is_p ? -rb.x : fma(-ci, rb.x, vb.x);
                r.y = is_p ? -rb.y : fma(-ci, rb.y, vb.y);
                if (jq == 0) r.x = is_p ? invq : cq;
                if (jq == 1) r.y = is_p ? invq : cq;
                row[jb] = r;
            }
        }
    }
}

// per-row / per-column bounds and column status from the by-variable arrays in HBM
template <int NT, bool TG>
__device__ __forceinline__ void gather_bounds(const BlockDev &B, Sm3 &c)
{
    const double *glo = B.lo, *gup = B.up;
    constexpr int U = 4;
    for (int i0 = threadIdx.x; i0 < c.m; i0 += NT * U) {
        double lo[U], up[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = i0 + u * NT;
            if (i < c.m) { const int v = c.head[i]; lo[u] = __ldg(glo + v); up[u] = __ldg(gup + v); }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = i0 + u * NT;
            if (i < c.m) { c.rlo[i] = lo[u]; c.rup[i] = up[u]; }
        }
    }
    for (int j = threadIdx.x; j < c.n; j += NT) {
        const int v = c.nbv[j];
        if (!TG) { c.clo[j] = __ldg(glo + v); c.cup[j] = __ldg(gup + v); }
        c.cstat[j] = c.vstat[v];
    }
}

// value of the nonbasic variable in column j
template <bool TG>
__device__ __forceinline__ double nb_col_value(const Sm3 &c, int j)
{
    const signed char st = c.cstat[j];
    if (st == ST_FREE) return 0.0;
    if (TG) { const int v = c.nbv[j]; return st == ST_UP ? __ldg(c.gup + v) : __ldg(c.glo + v); }
    return st == ST_UP ? c.cup[j] : c.clo[j];
}

// rows with a non-zero weight, ascending, into arow (warp 0); returns the count to warp 0's lanes
__device__ __forceinline__ int list_weighted_rows(Sm3 &c, const double *wt, int m)
{
    const int lane = threadIdx.x & 31;
    int nl = 0;
    for (int i0 = 0; i0 < m; i0 += 32) {
        const int i = i0 + lane;
        const bool nz = i < m && wt[i] != 0.0;
        const unsigned bal = __ballot_sync(0xffffffffu, nz);
        if (nz) c.arow[nl + __popc(bal & ((1u << lane) - 1u))] = (unsigned short)i;
        nl += __popc(bal);
    }
    return nl;
}

// s + sum over the listed rows (ascending) of wt[i] * T[i][j]; 8 loads in flight per thread
template <bool TG>
__device__ __forceinline__ double weighted_colsum(const Sm3 &c, const double *wt, int nl, int ld, int j, double s)
{
    int a = 0;
    for (; a + 8 <= nl; a += 8) {
        double v[8], wv[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int i = c.arow[a + t];
            wv[t] = wt[i];
            v[t] = ld_t<TG>(c.T + (size_t)i * ld + j);
        }
#pragma unroll
        for (int t = 0; t < 8; ++t) s = fma(wv[t], v[t], s);
    }
    for (; a < nl; ++a) {
        const int i = c.arow[a];
        s = fma(wt[i], ld_t<TG>(c.T + (size_t)i * ld + j), s);
    }
    return s;
}

// out[4t .. 4t+3] += sum over the listed rows (ascending) of wt[i] * T[i][4t .. 4t+3], visiting only the rows whose
// mask says item t may be non-zero (global-memory tableau).  Deterministic order.  The thread that owns item t first
// collects, 32 listed rows at a time, WHICH of them carry the item (shared-memory reads only), then fetches those —
// four 32-byte items in flight, every one of them useful (a network row carries ~20 % of the items, so a plain
// "4 rows at a time, predicated" loop spends most of its HBM round trips on one load or none).
template <int NT>
__device__ __forceinline__ void weighted_itemsum_tg(const Sm3 &c, const double *wt, int nl, int ld, double *out)
{
    const int nitems = ld >> 2, rwd = c.rwd, n2 = ld >> 1;
    const double2 *T2 = reinterpret_cast<const double2 *>(c.T);
    double2 *o2 = reinterpret_cast<double2 *>(out);
    for (int t = threadIdx.x; t < nitems; t += NT) {
        double2 a0 = o2[2 * t], a1 = o2[2 * t + 1];
        const unsigned *mcol = c.rmask + (t >> 5);
        const int sh = t & 31;
        for (int b0 = 0; b0 < nl; b0 += 32) {
            unsigned have = 0;
            const int lim = min(32, nl - b0);
#pragma unroll 8
            for (int u = 0; u < lim; ++u) have |= ((mcol[(size_t)c.arow[b0 + u] * rwd] >> sh) & 1u) << u;
            while (have) {
                double2 v0[4], v1[4];
                double wv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    wv[u] = 0.0; v0[u] = make_double2(0.0, 0.0); v1[u] = v0[u];      // w = 0 contributes exactly nothing
                    if (have) {
                        const int a = __ffs(have) - 1;
                        have &= have - 1;
                        const int i = c.arow[b0 + a];
                        wv[u] = wt[i];
                        v0[u] = __ldcg(T2 + (size_t)i * n2 + 2 * t);
                        v1[u] = __ldcg(T2 + (size_t)i * n2 + 2 * t + 1);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    a0.x = fma(wv[u], v0[u].x, a0.x); a0.y = fma(wv[u], v0[u].y, a0.y);
                    a1.x = fma(wv[u], v1[u].x, a1.x); a1.y = fma(wv[u], v1[u].y, a1.y);
                }
            }
        }
        o2[2 * t] = a0; o2[2 * t + 1] = a1;
    }
}

// tableau doubles the masked price pass reads for the listed rows (warp 0; result in every lane)
__device__ __forceinline__ unsigned masked_cells(const Sm3 &c, int nl)
{
    const int lane = threadIdx.x & 31;
    unsigned cnt = 0;
    for (int a = lane; a < nl; a += 32) {
        const unsigned *mi = c.rmask + (size_t)c.arow[a] * c.rwd;
        for (int w_ = 0; w_ < c.rwd; ++w_) cnt += __popc(mi[w_]);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    return cnt * 4u;
}

// Priced optimum of a block that did not move, from its standing proposal in the result arrays (one warp; every lane
// gets the value): (phase_one ? 0 : c_k.x_k) - sum_e pi[ind_e] val_e over the len non-zeros of D_k x_k.
__device__ __forceinline__ double standing_optimum(const SolveParams &P, int k, int L, int lane)
{
    const int len = P.len[k];
    const double *yv = P.val + (size_t)k * L;
    const int *yi = P.ind + (size_t)k * L;
    double z = 0.0;
    for (int t = lane; t < len; t += 32) z = fma(__ldg(P.pi + (__ldcg(yi + t) - 1)), __ldcg(yv + t), z);
#pragma unroll
    for (int o = 16; o; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
    return (P.phase_one ? 0.0 : __ldcg(P.cost + k)) - z;
}

// MC/NC != 0: block shape known at compile time (all shared-memory offsets and trip counts are
// constants); MC = NC = 0: any shape that fits.
// TG = true: the tableau stays in global memory (L2/HBM) and only the vectors live in shared memory
// (blocks too large for one CTA's shared memory, config B class); the decision/sweep code is shared.
template <int NT, int MC, int NC, bool TG>
__global__ void __launch_bounds__(NT, TG ? (NT <= 64 ? DWK_TG_OCC : DWK_TGW_OCC) : 3) solve_smem_kernel(const BlockDev *blocks, const int *list, int count,
                                                          SolveParams P, Continuation CN)
{
    extern __shared__ __align__(16) char smem_raw[];
    __shared__ Scratch<NT> scr;
    __shared__ Decision dec;
    __shared__ DecShared dsh;
    __shared__ __align__(8) uint64_t bar;
    __shared__ int s_kind, s_changed;
    if ((int)blockIdx.x >= count) return;
    const bool whole = CN.full_flag != nullptr && *CN.full_flag != 0;
    if (!whole && CN.dyn_count != nullptr && (int)blockIdx.x >= *CN.dyn_count) return;
    phase_init();
    const int k = (whole ? CN.full_list : list)[blockIdx.x];
    const BlockDev &B = blocks[k];      // fields are fetched where they are used (read-only, cached)
    const int m = MC ? MC : B.m, n = NC ? NC : B.n, L = B.L, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double tol_dj = P.tol_dj, tol_bnd = P.tol_bnd, tol_piv = P.tol_piv;

    Sm3 c;
    c.m = m; c.n = n; c.ld = tab_ld(n, TG);
    const int ld = c.ld;
    char *sp = smem_raw;
    if (TG) {
        c.T = B.T;
        c.drow = (double *)sp;     sp += align16(sizeof(double) * (size_t)ld);
    } else {
        c.T = (double *)sp;        sp += align16(sizeof(double) * (size_t)(m + 1) * ld);
        c.drow = c.T + (size_t)m * ld;
    }
    char *pst = sp;                sp += pst_bytes(m, n);
    c.beta = (double *)pst;
    c.head = (int *)(pst + sizeof(double) * (size_t)m);
    c.nbv = c.head + m;
    c.vstat = (signed char *)(c.nbv + n);
    c.rlo = (double *)sp; c.rup = c.rlo + m; c.clo = c.rup + m; c.cup = c.clo + n;
    sp += align16(sizeof(double) * 2 * (size_t)(TG ? m : m + n));
    c.glo = B.lo; c.gup = B.up;
    c.rowp = (double *)sp;         sp += align16(sizeof(double) * ld);
    if (TG) c.cs = B.ws;           // priced costs in HBM (written and read by this CTA only)
    else { c.cs = (double *)sp;    sp += align16(sizeof(double) * (size_t)n); }
    c.colq = (double *)sp;         sp += align16(sizeof(double) * (size_t)(m + 1));
    c.w = (double *)sp;            sp += align16(sizeof(double) * (size_t)m);
    c.cstat = (signed char *)sp;   sp += align16((size_t)n);
    c.arow = (unsigned short *)sp; sp += align16(sizeof(unsigned short) * (size_t)(m + 1));
    c.acol = (unsigned short *)sp; sp += align16(sizeof(unsigned short) * (size_t)(ld / 2 + 1));
    c.arow2 = (unsigned short *)sp; sp += TG ? align16(sizeof(unsigned short) * (size_t)(m + 32 * 8)) : 0;
    c.rmask = TG ? (unsigned *)sp : nullptr;
    c.rwd = rmask_words(n);
    const int dflag_in = B.dflag[0];        // (loaded here: its latency hides behind the bulk copies)
    // ---- state in: bulk copies (tableau or row masks, state image), one mbarrier
    const uint32_t t_bytes = TG ? 0u : (uint32_t)(sizeof(double) * (size_t)m * ld);
    const uint32_t p_bytes = (uint32_t)pst_bytes(m, n);
    const uint32_t k_bytes = TG ? (uint32_t)rmask_bytes(m, n) : 0u;
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_expect_tx(&bar, t_bytes + p_bytes + k_bytes);
        if (t_bytes) bulk_g2s(c.T, B.T, t_bytes, &bar);
        if (k_bytes) bulk_g2s(c.rmask, B.rmask, k_bytes, &bar);
        bulk_g2s(pst, B.beta, p_bytes, &bar);
        s_changed = (CN.moved != nullptr && CN.moved[k]) ? 1 : 0;
    }
    // priced objective while the copies fly (same arithmetic as the generic kernel / reference)
    // (four columns per thread at a time, every level of the CSC walk issued as one batch of read-only
    //  loads: a chain of 3 dependent loads per column otherwise, which dominated the no-pivot rounds)
    if (P.initial) {
        const double *cf = B.c_file;
        for (int j = tid; j < n; j += NT) c.cs[j] = __ldg(cf + j);
    } else {
        const int *dcp = B.dc_ptr, *dcr = B.dc_row;
        const double *dcv = B.dc_val, *cm = B.c_master;
        constexpr int U = (NC && NC / NT >= 8) ? 8 : 4;      // 64-thread CTAs on 512 columns: the whole walk is one batch per level
        for (int j0 = tid; j0 < n; j0 += NT * U) {
            int pb[U], pe[U], r0[U];
            double v0[U], p0[U], base[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int j = j0 + u * NT;
                pb[u] = pe[u] = 0;
                if (j < n) { pb[u] = __ldg(dcp + j); pe[u] = __ldg(dcp + j + 1); }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int j = j0 + u * NT;
                base[u] = (j < n && !P.phase_one) ? __ldg(cm + j) : 0.0;
                r0[u] = 0; v0[u] = 0.0;
                if (pb[u] < pe[u]) { r0[u] = __ldg(dcr + pb[u]); v0[u] = __ldg(dcv + pb[u]); }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) p0[u] = pb[u] < pe[u] ? __ldg(P.pi + r0[u]) : 0.0;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int j = j0 + u * NT;
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
    __syncthreads();            // mbarrier init + cs visible
    mbar_wait(&bar, 0);
    gather_bounds<NT, TG>(B, c);
    __syncthreads();

    int status = 0;
    long long iter = 0;
    const long long max_iter = (long long)P.max_iter_mult * (m + n) + 1000;
    unsigned long long n_piv = 0, n_flip = 0, n_reb = 0, n_p1 = 0, n_rows = 0, n_cells = 0;   // n_piv/n_flip/n_rows/n_cells live in warp 0
    unsigned long long n_price = 0;     // tableau doubles read to (re)compute price rows (block-uniform)
    bool need_check = true, infeasible = false, d_valid = false, bland = false, d_full = false, x_in_rowp = true;
    int degen = 0;
    const int degen_limit = 100 + (m + n) / 2;
    const bool use_saved = dflag_in > 0 && dflag_in < 32;
    bool reb_now = TG && CN.reb != nullptr && CN.reb[k] != 0;      // block-uniform

    for (;;) {
        if (++iter > max_iter) { status = 1; break; }
        phase(1);
        if (!reb_now) {
        // ---- A. cooperative parts: feasibility scan, phase-1 prices, fresh reduced costs
        if (need_check || infeasible) {
            int viol = 0;
            for (int i = tid; i < m; i += NT) {
                const double b = c.beta[i], lo = c.rlo[i], up = c.rup[i];
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
            // phase-1 prices T'w into rowp (consumed by the pricing loop before rowp is restaged)
            if (warp == 0) {
                const int nl = list_weighted_rows(c, c.w, m);
                if (lane == 0) dec.na = nl;
                if (TG) { __syncwarp(); n_price += masked_cells(c, nl); } else n_price += (unsigned long long)nl * (unsigned long long)n;
            }
            if (TG) for (int j = tid; j < ld; j += NT) c.rowp[j] = 0.0;
            __syncthreads();
            {
                const int nl = dec.na;
                if (TG) weighted_itemsum_tg<NT>(c, c.w, nl, ld, c.rowp);
                else for (int j = tid; j < n; j += NT) c.rowp[j] = weighted_colsum<TG>(c, c.w, nl, ld, j, 0.0);
            }
            ++n_p1;
            __syncthreads();
        } else if (!d_valid) {
            phase(2);
            // reduced costs.  If the basis has not moved since the last optimal solve of this block,
            // update the saved row incrementally: d = d_saved + dc_N + T' dc_B with dc = cs - cs_saved,
            // which only touches the rows of T whose basic cost changed (late rounds: a handful);
            // otherwise (and every 32nd solve, to stop drift) recompute d = c_N + T' c_B in full.
            const bool inc = use_saved && !s_changed && n_reb == 0;
            const double *cssave = B.cssave, *dsave = B.dsave;     // only written in the epilogue: read-only here
            constexpr int U = (NC && NC / NT >= 8) ? 8 : 4;
            for (int i0 = tid; i0 < m; i0 += NT * U) {
                double a[U], b[U];
                bool st[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int i = i0 + u * NT;
                    st[u] = false; a[u] = 0.0; b[u] = 0.0;
                    if (i < m) {
                        const int v = c.head[i];
                        if (v >= m) { st[u] = true; a[u] = c.cs[v - m]; if (inc) b[u] = __ldg(cssave + v - m); }
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int i = i0 + u * NT;
                    if (i < m) c.colq[i] = st[u] ? (inc ? cost_delta(a[u], b[u]) : a[u]) : 0.0;
                }
            }
            __syncthreads();
            if (warp == 0) {
                const int nl = list_weighted_rows(c, c.colq, m);
                if (lane == 0) dec.na = nl;
                if (TG) { __syncwarp(); n_price += masked_cells(c, nl); } else n_price += (unsigned long long)nl * (unsigned long long)n;
            }
            __syncthreads();
            const int nl = dec.na;
            for (int j0 = tid; j0 < ld; j0 += NT * U) {
                double a[U], b[U], dsv[U];
                bool st[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int j = j0 + u * NT;
                    st[u] = false; a[u] = 0.0; b[u] = 0.0; dsv[u] = 0.0;
                    if (j < n) {
                        const int v = c.nbv[j];
                        if (v >= m) { st[u] = true; a[u] = c.cs[v - m]; if (inc) b[u] = __ldg(cssave + v - m); }
                        if (inc) dsv[u] = __ldg(dsave + j);
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int j = j0 + u * NT;
                    if (j >= ld) continue;
                    double s = 0.0;
                    if (j < n) {
                        if (st[u]) s = inc ? cost_delta(a[u], b[u]) : a[u];
                        if (inc) s += dsv[u];
                        if (!TG) s = weighted_colsum<TG>(c, c.colq, nl, ld, j, s);
                    }
                    c.drow[j] = s;
                }
            }
            if (TG) {           // base row written above; now the masked row contributions, item by item
                __syncthreads();
                weighted_itemsum_tg<NT>(c, c.colq, nl, ld, c.drow);
            }
            d_valid = true;
            d_full = !inc;
            __syncthreads();
        }
        // ---- B. decision (TG: the whole CTA; shared-memory tableau: warp 0)
        if (TG) {       // (the row-item masks are maintained by decide_block_tg only)
            int kind;
            if (infeasible)
                kind = decide_block_tg<true, NT, MC, NC>(c, c.rowp, false, bland, degen, degen_limit, tol_dj, tol_bnd, tol_piv,
                                                         n_piv, n_flip, n_rows, n_cells, &dec, &dsh);
            else
                kind = decide_block_tg<false, NT, MC, NC>(c, c.drow, true, bland, degen, degen_limit, tol_dj, tol_bnd, tol_piv,
                                                          n_piv, n_flip, n_rows, n_cells, &dec, &dsh);
            if (tid == 0) { s_kind = kind; if (n_piv + n_flip) s_changed = 1; }
        } else if (warp == 0) {
            int kind;
            if (infeasible)
                kind = decide_warp0<true, MC, NC, TG>(c, c.rowp, false, bland, degen, degen_limit, tol_dj, tol_bnd, tol_piv,
                                          n_piv, n_flip, n_rows, n_cells, &dec);
            else
                kind = decide_warp0<false, MC, NC, TG>(c, c.drow, true, bland, degen, degen_limit, tol_dj, tol_bnd, tol_piv,
                                           n_piv, n_flip, n_rows, n_cells, &dec);
            if (lane == 0) { s_kind = kind; if (n_piv + n_flip) s_changed = 1; }
        }
        __syncthreads();
        const int kind = s_kind;
        if (kind == K_PIVOT) {
            phase(dec.nc >= 0 ? 7 : 11);
            if (dec.nc >= 0) sweep_sparse<NT, NC, TG>(c, dec.p, dec.q, dec.inv, dec.na, dec.nc, TG && dec.fused != 0, dec.dead != 0);
            else sweep_active<NT, NC, TG>(c, dec.p, dec.q, dec.inv, dec.na, dec.dead != 0);
            __syncthreads();
            continue;
        }
        if (kind == K_REDO) continue;
        if (kind == K_FAIL) { status = dec.fail_status; break; }
        // ---- K_NONE
        if (infeasible) { status = 4; break; }
        // nothing moved since the last optimal solve of this block: x_k is still what B.x holds
        if (!s_changed && n_reb == 0 && !P.initial && !P.force_full) { status = 5; x_in_rowp = false; break; }
        }   // !reb_now
        phase(8);
        // claimed optimal: x into rowp, auxiliary values into colq
        for (int j = tid; j < n; j += NT) {
            const int v = c.nbv[j];
            const double xv = nb_col_value<TG>(c, j);
            if (v >= m) c.rowp[v - m] = xv; else c.colq[v] = xv;
        }
        for (int i = tid; i < m; i += NT) {
            const int v = c.head[i];
            if (v >= m) c.rowp[v - m] = c.beta[i]; else c.colq[v] = c.beta[i];
        }
        __syncthreads();
        // residual of the state against the original rows
        double worst = 0.0;
        for (int i = tid; i < m; i += NT) {
            double s = 0.0;
            for (int e = B.a_rp[i]; e < B.a_rp[i + 1]; ++e) s = fma(B.a_v[e], c.rowp[B.a_ci[e]], s);
            worst = fmax(worst, fabs(s - c.colq[i]) / (1.0 + fabs(s)));
        }
        worst = block_max<NT>(worst, scr);
        if ((worst > 0.1 * tol_bnd || reb_now) && n_reb < 3) {          // n_reb is block-uniform
            ++n_reb;
            reb_now = false;
            phase(9);
            // rebuild T from A_k for the current basis
            for (int e = tid; e < m * ld; e += NT) c.T[e] = 0.0;
            if (TG) {       // the whole image is rewritten here and the rebuilt pattern is unknown: every item exists
                const int nitems = ld >> 2, tail = nitems & 31;
                for (int e = tid; e < m * c.rwd; e += NT)
                    c.rmask[e] = (tail && (e % c.rwd) == c.rwd - 1) ? (1u << tail) - 1u : 0xffffffffu;
            }
            __syncthreads();
            for (int i = tid; i < m; i += NT) {
                for (int e = B.a_rp[i]; e < B.a_rp[i + 1]; ++e) c.T[(size_t)i * ld + B.a_ci[e]] += B.a_v[e];
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
                    const double tq = ld_t<TG>(c.T + (size_t)i * ld + j);
                    c.colq[i] = tq;
                    if (h < m && c.vstat[h] != ST_BASIC) {
                        const double a = fabs(tq);
                        if (a > pk) { pk = a; pr = i; }
                    }
                }
                block_argmax<NT>(pk, pr, scr);
                if (pr == INT_MAX || pk < 1e-11) { singular = true; break; }
                const double inv = 1.0 / c.colq[pr];
                for (int jj = tid; jj < ld; jj += NT) c.rowp[jj] = ld_t<TG>(c.T + (size_t)pr * ld + jj) * inv;
                for (int i = tid; i < m; i += NT) c.arow[i] = (unsigned short)i;
                __syncthreads();
                sweep_active<NT, NC, TG>(c, pr, j, inv, m, false);
                if (tid == 0) { const int t = c.head[pr]; c.head[pr] = c.nbv[j]; c.nbv[j] = t; }
                __syncthreads();
            }
            if (singular) { status = 1; break; }
            gather_bounds<NT, TG>(B, c);
            __syncthreads();
            {   // beta = T x_N, one warp per row
                constexpr int NW = NT / 32;
                for (int i = warp; i < m; i += NW) {
                    double s = 0.0;
                    for (int j = lane; j < n; j += 32)
                        s += ld_t<TG>(c.T + (size_t)i * ld + j) * nb_col_value<TG>(c, j);
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
    __syncthreads();
    phase(10);
    const bool changed = s_changed != 0 || n_reb != 0 || P.initial != 0 || P.force_full != 0 || status != 5;
    if (status != 5) {
        for (int j = tid; j < n; j += NT) {
            const int v = c.nbv[j];
            if (v >= m) c.rowp[v - m] = nb_col_value<TG>(c, j);
        }
        for (int i = tid; i < m; i += NT) {
            const int v = c.head[i];
            if (v >= m) c.rowp[v - m] = c.beta[i];
        }
        __syncthreads();
    }
    if (!changed) {
        // the block was already optimal under the new prices: x_k, D_k x_k, c_k.x_k and the whole
        // state in HBM are still those of the previous round; only the priced optimum moves
        {
            const double *cssv = B.cssave;
            double *dsave = B.dsave, *cssave = B.cssave;
            constexpr int U = (NC && NC / NT >= 8) ? 8 : 4;
            for (int j0 = tid; j0 < n; j0 += NT * U) {
                double cv[U], sv[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int j = j0 + u * NT;
                    if (j < n) { cv[u] = c.cs[j]; sv[u] = __ldcg(cssv + j); }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int j = j0 + u * NT;
                    if (j < n) {
                        dsave[j] = c.drow[j];
                        if (d_full || cost_delta(cv[u], sv[u]) != 0.0) cssave[j] = cv[u];
                    }
                }
            }
        }
        // priced optimum of the kept x_k from the block's standing proposal: c_k.x_k - pi.(D_k x_k) (= sum cs_j x_j); the
        // same warp code as the delta rounds (delta_tg.cuh), so that both routes of an unmoved block agree bit for bit
        double zo = 0.0;
        if (warp == 0) zo = standing_optimum(P, k, L, lane);
        if (tid == 0) {
            P.obj[k] = zo; P.status[k] = 5; P.last_iters[k] = 0;
            if (P.m_obj != nullptr) { P.m_obj[k] = zo; P.m_status[k] = 5; P.counters[8 * (size_t)k + 7] += 12; }   // the rest of the mailbox still holds
            P.counters[8 * (size_t)k + 6] += n_price;
            B.dflag[0] = d_full ? 1 : dflag_in + 1;
        }
        phase_flush();
        return;
    }
    if (status == 5 && d_valid) {
        for (int j = tid; j < n; j += NT) {
            B.dsave[j] = c.drow[j];
            if (d_full || cost_delta(c.cs[j], B.cssave[j]) != 0.0) B.cssave[j] = c.cs[j];
        }
        if (tid == 0) B.dflag[0] = d_full ? 1 : dflag_in + 1;
    } else if (tid == 0) B.dflag[0] = 0;
    if (TG && B.pos != nullptr) {       // where every variable sits now (delta rounds look it up instead of searching)
        int *pos = B.pos;
        for (int i = tid; i < m; i += NT) pos[c.head[i]] = i;
        for (int j = tid; j < n; j += NT) pos[c.nbv[j]] = j;
    }
    // state out first (async), then the proposal while the copies fly
    if (tid == 0) {
        const bool t_dirty = (n_piv + n_reb) != 0;        // thread 0 sits in warp 0, which owns n_piv
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (t_dirty && t_bytes) bulk_s2g(B.T, c.T, t_bytes);
        if (t_dirty && k_bytes) bulk_s2g(B.rmask, c.rmask, k_bytes);
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
            P.counters[8 * (size_t)k + 2] += n_reb;
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
    phase_flush();
}

}  // namespace dwk
