// simplex_sparse.cuh — path 5: CTA-per-block simplex on a PACKED SPARSE tableau that lives in shared memory.
//
// Why: the condensed tableau of a network-type block (config B: 256 x 512) holds ~3 % non-zeros (measured along the
// recorded Dantzig-Wolfe trajectory: 3.8 k of 131 k cells on average, 12 k at most).  Path 3 keeps the dense 1 MB image
// in HBM and pays two dependent HBM round trips plus ~900 scattered 32-byte sector operations per basis change.  Packed,
// the same tableau is ~30 KB of values and 17 KB of bit masks: it fits one CTA's shared memory, two CTAs per SM, and a
// basis change never leaves the SM.
//
// Representation (identical in HBM and in shared memory, so one TMA bulk copy moves it each way):
//   hdr[4]            ints: [0] top = doubles of the heap in use, [1] live = sum of the row capacities
//   mask[m][WS]       one bit per column: EXACT non-zero pattern of the row (WS = odd word stride: conflict-free)
//   off[m] (u32), cap[m], len[m] (u16)    row directory
//   heap[]            the rows' non-zero values in column order; a row owns cap >= len consecutive doubles
// A rank-1 update rewrites every active row through a two-pass warp routine (values staged per warp, exclusive scans
// of the per-lane popcounts give the packed positions); fill-ins grow a row inside its slack or move it to the top of
// the heap, cancellations to exact zero shrink it.  When the heap fills up the rows are re-packed through the HBM
// image; a block that outgrows the small heap (2 CTAs per SM) is handed to a second launch with one CTA per SM, and one
// that outgrows that to the dense path 3, which rebuilds its tableau from A_k for the current basis.
//
// Rounds in which a block needs no pivot (most blocks of most rounds) never load the tableau: a PROBE launch with
// vector-only shared memory (8 CTAs per SM) prices the block, updates the saved reduced costs from the few packed
// rows whose basic cost moved (read straight from the HBM image) and either finishes the block or appends it to the
// list of the solve launch.
//
// Same pivoting rules, tolerances, arithmetic (operation order, FMA use) and tie-breaks as path 3
// (simplex_smem.cuh, 128-thread instantiation): the two paths visit the same vertices bit for bit.
// Reference being replaced: one warm-started glp_simplex call per block and round
// (/root/reference/src/dw_subprob.c:329) plus the pricing / column code around it (:291-321, :453-540).
#pragma once
#include "simplex_smem.cuh"

namespace dwk {

constexpr int SP_STG = 128;          // doubles of per-warp staging (rows with more candidates are rewritten out of place)
constexpr int SP_PROBE_ROWS = 32;    // packed rows a probe stages for the incremental reduced costs
constexpr int SP_MAX_N = 512;        // one 16-column chunk per lane
constexpr int SP_MAX_M = 1024;

__host__ __device__ inline int sp_words(int n) { return (n + 31) >> 5; }
__host__ __device__ inline int sp_stride(int n) { return sp_words(n) | 1; }
__host__ __device__ inline size_t sp_mask_bytes(int m, int n) { return align16(sizeof(unsigned) * (size_t)m * sp_stride(n)); }
__host__ __device__ inline size_t sp_off_pos(int m, int n) { return 16 + sp_mask_bytes(m, n); }
__host__ __device__ inline size_t sp_cap_pos(int m, int n) { return sp_off_pos(m, n) + align16(sizeof(unsigned) * (size_t)m); }
__host__ __device__ inline size_t sp_len_pos(int m, int n) { return sp_cap_pos(m, n) + align16(sizeof(unsigned short) * (size_t)m); }
__host__ __device__ inline size_t sp_heap_pos(int m, int n) { return sp_len_pos(m, n) + align16(sizeof(unsigned short) * (size_t)m); }
// capacity given to a row of t non-zeros
__host__ __device__ inline int sp_newcap(int t, int n)
{
    const int c = t + (t >> 2 > 4 ? t >> 2 : 4);
    return c < n ? c : n;
}
// vectors every kernel of the path keeps in shared memory
__host__ __device__ inline size_t sp_vec_bytes(int m, int n)
{
    return align16(sizeof(double) * (size_t)n)                 // drow (reduced costs / phase-1 prices)
         + pst_bytes(m, n)                                     // beta head nbv vstat
         + align16(sizeof(double) * 2 * (size_t)m)             // rlo rup
         + align16(sizeof(double) * (size_t)n)                 // rowp (scaled pivot row; x staging)
         + align16(sizeof(double) * (size_t)(m + 1))           // colq
         + align16(sizeof(double) * (size_t)m)                 // w
         + align16((size_t)n)                                  // cstat
         + align16(sizeof(unsigned short) * (size_t)(m + 1))   // arow
         + align16(sizeof(unsigned short) * (size_t)(m + 32 * 8))   // per-warp column segments
         + align16(sizeof(unsigned) * (size_t)sp_stride(n));   // pivot-row pattern
}
__host__ __device__ inline size_t sp_probe_bytes(int m, int n)
{
    return sp_vec_bytes(m, n) + align16(sizeof(unsigned) * (size_t)SP_PROBE_ROWS * (sp_stride(n) + 1));
}
// shared memory of the solve kernel with a heap of H doubles
__host__ __device__ inline size_t sp_solve_bytes(int m, int n, int nt, int H)
{
    return sp_vec_bytes(m, n) + sp_heap_pos(m, n) + sizeof(double) * (size_t)H + sizeof(double) * (size_t)(nt / 32) * SP_STG;
}

// hand-over lists of a round (device side)
struct SparseRound {
    const int *in_count;            // nullptr: the launch covers `count` blocks; else the device-side length of the list
    int use_moved;                  // the blocks of this launch come from a launch that may have moved them
    int *next_list, *next_count;    // blocks that outgrow this launch's heap (probe: blocks that need a pivot)
    int *dense_list, *dense_count;  // blocks for the dense path-3 kernel
    unsigned char *moved, *reb;     // by block id: the state changed before the hand-over / the dense kernel must rebuild T
    unsigned long long *stats;      // [0] small -> large heap, [1] -> dense (heap), [2] -> dense (residual), [3] re-packs,
                                    // [4] batched sweeps, [5] union-size checks, [6] blocks the probe finished, [7] probe hand-overs
};

struct Sp {
    int m, n, W, WS, H;
    int *hdr;
    unsigned *mask, *off;
    unsigned short *cap, *len;
    double *heap;
    double *drow, *rowp, *beta, *rlo, *rup, *colq, *w, *cs, *stg;
    int *head, *nbv;
    signed char *vstat, *cstat;
    const double *glo, *gup;
    unsigned short *arow, *arow2;
    unsigned *pmask;
};

struct SpDec {
    int p, q, na, pe, kind, need, bland, fail_status, lenp;
    double inv, dq;
};
enum { K_TIGHT = 5 };

// non-zeros of a row before column j
__device__ __forceinline__ int sp_rank(const unsigned *mrow, int j)
{
    const int wq = j >> 5;
    int r = __popc(mrow[wq] & ((1u << (j & 31)) - 1u));
    for (int w_ = 0; w_ < wq; ++w_) r += __popc(mrow[w_]);
    return r;
}
// the lane's 16 columns of a row pattern
__device__ __forceinline__ unsigned sp_chunk(const unsigned *mrow, int lane, int W)
{
    const int w_ = lane >> 1;
    return w_ < W ? (mrow[w_] >> ((lane & 1) * 16)) & 0xffffu : 0u;
}
__device__ __forceinline__ int warp_excl_scan(int v, int lane, int &total)
{
    int s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += t;
    }
    total = __shfl_sync(0xffffffffu, s, 31);
    return s - v;
}
// pricing order of the 128-thread kernels of path 3: among equal keys the smaller (j mod 128, j div 128) wins
__device__ __forceinline__ int sp_tie(int j) { return ((j & 127) << 16) | (j >> 7); }

// ---- T <- A_k as a packed image in HBM (init / reset): one thread per row sorts its entries by column ----------
template <int NT>
__device__ void sp_init_image(const BlockDev &B)
{
    const int m = B.m, n = B.n, WS = sp_stride(n);
    char *img = reinterpret_cast<char *>(B.T);
    int *hdr = reinterpret_cast<int *>(img);
    unsigned *mask = reinterpret_cast<unsigned *>(img + 16);
    unsigned *off = reinterpret_cast<unsigned *>(img + sp_off_pos(m, n));
    unsigned short *cap = reinterpret_cast<unsigned short *>(img + sp_cap_pos(m, n));
    unsigned short *len = reinterpret_cast<unsigned short *>(img + sp_len_pos(m, n));
    double *heap = reinterpret_cast<double *>(img + sp_heap_pos(m, n));
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int i = 0; i < m; ++i) {
            const int c = sp_newcap(B.a_rp[i + 1] - B.a_rp[i], n);
            off[i] = t; cap[i] = (unsigned short)c; t += (unsigned)c;
        }
        hdr[0] = (int)t; hdr[1] = (int)t; hdr[2] = 0; hdr[3] = 0;
    }
    for (int e = threadIdx.x; e < m * WS; e += NT) mask[e] = 0u;
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += NT) {
        double *row = heap + off[i];
        unsigned *mr = mask + (size_t)i * WS;
        int l = 0;
        for (int p = B.a_rp[i]; p < B.a_rp[i + 1]; ++p) {
            const int j = B.a_ci[p];
            const double v = B.a_v[p];
            const int r = sp_rank(mr, j);
            if ((mr[j >> 5] >> (j & 31)) & 1u) { row[r] += v; continue; }     // duplicate entry: summed like the dense scatter
            for (int t = l; t > r; --t) row[t] = row[t - 1];
            row[r] = v; ++l;
            mr[j >> 5] |= 1u << (j & 31);
        }
        // entries that are (or summed to) exact zero leave the pattern
        int o = 0, t = 0;
        for (int w_ = 0; w_ < sp_words(n); ++w_) {
            unsigned bits = mr[w_], keep = bits;
            while (bits) {
                const int b = __ffs(bits) - 1; bits &= bits - 1;
                const double v = row[t++];
                if (v != 0.0) row[o++] = v; else keep &= ~(1u << b);
            }
            mr[w_] = keep;
        }
        len[i] = (unsigned short)o;
        B.head[i] = i;
    }
    for (int j = threadIdx.x; j < n; j += NT) B.nbv[j] = m + j;
    if (threadIdx.x == 0) B.spflag[0] = 0;
    __syncthreads();
}

// ---- pieces shared by the probe and the solve kernel -----------------------------------------------------------
// priced objective cs = (phase_one ? 0 : c_master) - D_k' pi (same arithmetic as path 3 / the reference), or the block
// file's own objective for the cold solve
template <int NT>
__device__ __forceinline__ void sp_priced_costs(const BlockDev &B, const SolveParams &P, double *cs, int n)
{
    const int tid = threadIdx.x;
    if (P.initial) {
        const double *cf = B.c_file;
        for (int j = tid; j < n; j += NT) cs[j] = __ldg(cf + j);
        return;
    }
    const int *dcp = B.dc_ptr, *dcr = B.dc_row;
    const double *dcv = B.dc_val, *cm = B.c_master;
    constexpr int U = 4;
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
            cs[j] = __dadd_rn(__dmul_rn(-1.0, y), base[u]);
        }
    }
}

template <int NT>
__device__ __forceinline__ void sp_gather_bounds(const BlockDev &B, Sp &c)
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
    for (int j = threadIdx.x; j < c.n; j += NT) c.cstat[j] = c.vstat[c.nbv[j]];
}

__device__ __forceinline__ double sp_nb_col_value(const Sp &c, int j)
{
    const signed char st = c.cstat[j];
    if (st == ST_FREE) return 0.0;
    const int v = c.nbv[j];
    return st == ST_UP ? __ldg(c.gup + v) : __ldg(c.glo + v);
}

// rows with a non-zero weight, ascending, into arow (warp 0); returns the count to warp 0's lanes
__device__ __forceinline__ int sp_list_rows(Sp &c, const double *wt, int m)
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

// out[j] += sum over the listed rows (ascending) of wt[i] * T[i][j]: the thread that owns a group of GC columns walks the
// list in order (deterministic), tests the group's bits in the row pattern and fetches only the values that exist.
// STAGED: the listed rows' patterns and offsets were copied to shared memory (smask, soff; list position a), the values
// are read from the HBM image; otherwise everything is in shared memory.
template <int NT, bool STAGED>
__device__ __forceinline__ void sp_weighted_sum(const Sp &c, const double *wt, int nl, double *out,
                                                const unsigned *smask, const unsigned *soff)
{
    constexpr int GC = NT >= 256 ? 2 : 4;
    const int n = c.n, WS = c.WS;
    const int groups = (n + GC - 1) / GC;
    for (int g = threadIdx.x; g < groups; g += NT) {
        const int j0 = g * GC, wq = j0 >> 5, sh = j0 & 31;
        const unsigned gm = ((1u << GC) - 1u) << sh;
        double acc[GC];
#pragma unroll
        for (int u = 0; u < GC; ++u) acc[u] = j0 + u < n ? out[j0 + u] : 0.0;
        for (int a = 0; a < nl; ++a) {
            const int i = c.arow[a];
            const unsigned *mr = STAGED ? smask + (size_t)a * WS : c.mask + (size_t)i * WS;
            const unsigned word = mr[wq];
            if (!(word & gm)) continue;
            int r = __popc(word & ((1u << sh) - 1u));
            for (int w_ = 0; w_ < wq; ++w_) r += __popc(mr[w_]);
            const unsigned base = (STAGED ? soff[a] : c.off[i]) + (unsigned)r;
            const double wv = wt[i];
            int t = 0;
#pragma unroll
            for (int u = 0; u < GC; ++u) {
                if ((word >> (sh + u)) & 1u) {
                    const double v = STAGED ? __ldcg(c.heap + base + t) : c.heap[base + t];
                    acc[u] = fma(wv, v, acc[u]);
                    ++t;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < GC; ++u) if (j0 + u < n) out[j0 + u] = acc[u];
    }
}

// non-zeros of the listed rows (what a price pass reads), warp 0
__device__ __forceinline__ unsigned sp_listed_cells(const Sp &c, int nl, const unsigned short *len)
{
    const int lane = threadIdx.x & 31;
    unsigned cnt = 0;
    for (int a = lane; a < nl; a += 32) cnt += len[c.arow[a]];
#pragma unroll
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    return cnt;
}

// pricing over all threads: entering column (block-uniform), -1 if none
template <int NT>
__device__ __forceinline__ int sp_price(const Sp &c, const double *dd, bool bland, double tol_dj, DecShared *ds)
{
    constexpr int NW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n = c.n;
    double key = 0.0; int q = -1, tq = 0x7fffffff;
    for (int j = tid; j < n; j += NT) {
        const double dj = dd[j];
        const signed char st = c.cstat[j];
        const bool ok = (st == ST_LO && dj < -tol_dj) || (st == ST_UP && dj > tol_dj) || (st == ST_FREE && fabs(dj) > tol_dj);
        const double kj = bland ? 2147483648.0 - (double)c.nbv[j] : fabs(dj);
        if (ok && kj > 0.0) {
            const int tj = sp_tie(j);
            if (kj > key || (kj == key && tj < tq)) { key = kj; q = j; tq = tj; }
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const double ok_ = __shfl_xor_sync(0xffffffffu, key, o);
        const int oq = __shfl_xor_sync(0xffffffffu, q, o), ot = __shfl_xor_sync(0xffffffffu, tq, o);
        if (oq >= 0 && (q < 0 || ok_ > key || (ok_ == key && ot < tq))) { key = ok_; q = oq; tq = ot; }
    }
    if (lane == 0) { ds->key[warp] = key; ds->q[warp] = q; }
    __syncthreads();
    q = -1; key = 0.0; tq = 0x7fffffff;
#pragma unroll
    for (int w_ = 0; w_ < NW; ++w_) {
        const double kw = ds->key[w_];
        const int qw = ds->q[w_];
        if (qw >= 0) {
            const int tw = sp_tie(qw);
            if (q < 0 || kw > key || (kw == key && tw < tq)) { key = kw; q = qw; tq = tw; }
        }
    }
    return q;
}

// feasibility of the basic values: weights -1 / 0 / +1 into w; block-uniform result
template <int NT>
__device__ __forceinline__ bool sp_feasibility(Sp &c, double tol_bnd)
{
    int viol = 0;
    for (int i = threadIdx.x; i < c.m; i += NT) {
        const double b = c.beta[i], lo = c.rlo[i], up = c.rup[i];
        double wi = 0.0;
        if (b < lo - tol_bnd * (1.0 + fabs(lo))) wi = -1.0;
        else if (b > up + tol_bnd * (1.0 + fabs(up))) wi = 1.0;
        c.w[i] = wi;
        viol |= (wi != 0.0);
    }
    return __syncthreads_or(viol) != 0;
}

// reduced costs, step 1: cost (or cost change) of the basic variables into colq; step 3 (after the row list is known):
// base row into drow.  inc: incremental update of the saved row (d = d_saved + dc_N + T' dc_B).
template <int NT>
__device__ __forceinline__ void sp_basic_costs(const BlockDev &B, Sp &c, bool inc)
{
    const double *cssave = B.cssave;
    constexpr int U = 4;
    const int m = c.m;
    for (int i0 = threadIdx.x; i0 < m; i0 += NT * U) {
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
}
template <int NT>
__device__ __forceinline__ void sp_base_row(const BlockDev &B, Sp &c, bool inc)
{
    const double *cssave = B.cssave, *dsave = B.dsave;
    constexpr int U = 4;
    const int m = c.m, n = c.n;
    for (int j0 = threadIdx.x; j0 < n; j0 += NT * U) {
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
            if (j >= n) continue;
            double s = 0.0;
            if (st[u]) s = inc ? cost_delta(a[u], b[u]) : a[u];
            if (inc) s += dsv[u];
            c.drow[j] = s;
        }
    }
}

// the block was already optimal under the new prices: only the priced optimum and the saved costs move
template <int NT>
__device__ __forceinline__ void sp_unchanged_epilogue(const BlockDev &B, const SolveParams &P, Sp &c, int k, bool d_full, int dflag_in,
                                                      unsigned long long n_price, Scratch<NT> &scr)
{
    const int tid = threadIdx.x, n = c.n;
    double zo = 0.0;
    {
        const double *xk = B.x;
        double *dsave = B.dsave, *cssave = B.cssave;
        constexpr int U = 4;
        for (int j0 = tid; j0 < n; j0 += NT * U) {
            double cv[U], xv[U], sv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int j = j0 + u * NT;
                if (j < n) { cv[u] = c.cs[j]; xv[u] = __ldcg(xk + j); sv[u] = __ldcg(cssave + j); }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int j = j0 + u * NT;
                if (j < n) {
                    zo = fma(cv[u], xv[u], zo);
                    dsave[j] = c.drow[j];
                    if (d_full || cost_delta(cv[u], sv[u]) != 0.0) cssave[j] = cv[u];
                }
            }
        }
    }
    zo = block_sum<NT>(zo, scr);
    if (tid == 0) {
        P.obj[k] = zo; P.status[k] = 5; P.last_iters[k] = 0;
        if (P.m_obj != nullptr) { P.m_obj[k] = zo; P.m_status[k] = 5; P.counters[8 * (size_t)k + 7] += 12; }
        P.counters[8 * (size_t)k + 6] += n_price;
        B.dflag[0] = d_full ? 1 : dflag_in + 1;
    }
}

// carve the vector part of the shared-memory layout; returns the first free byte
__device__ __forceinline__ char *sp_carve(Sp &c, char *sp, char *&pst, const BlockDev &B)
{
    const int m = c.m, n = c.n;
    c.drow = (double *)sp;          sp += align16(sizeof(double) * (size_t)n);
    pst = sp;                       sp += pst_bytes(m, n);
    c.beta = (double *)pst;
    c.head = (int *)(pst + sizeof(double) * (size_t)m);
    c.nbv = c.head + m;
    c.vstat = (signed char *)(c.nbv + n);
    c.rlo = (double *)sp; c.rup = c.rlo + m; sp += align16(sizeof(double) * 2 * (size_t)m);
    c.rowp = (double *)sp;          sp += align16(sizeof(double) * (size_t)n);
    c.colq = (double *)sp;          sp += align16(sizeof(double) * (size_t)(m + 1));
    c.w = (double *)sp;             sp += align16(sizeof(double) * (size_t)m);
    c.cstat = (signed char *)sp;    sp += align16((size_t)n);
    c.arow = (unsigned short *)sp;  sp += align16(sizeof(unsigned short) * (size_t)(m + 1));
    c.arow2 = (unsigned short *)sp; sp += align16(sizeof(unsigned short) * (size_t)(m + 32 * 8));
    c.pmask = (unsigned *)sp;       sp += align16(sizeof(unsigned) * (size_t)c.WS);
    c.glo = B.lo; c.gup = B.up;
    c.cs = B.ws;
    return sp;
}
__device__ __forceinline__ void sp_image_pointers(Sp &c, char *img)
{
    c.hdr = reinterpret_cast<int *>(img);
    c.mask = reinterpret_cast<unsigned *>(img + 16);
    c.off = reinterpret_cast<unsigned *>(img + sp_off_pos(c.m, c.n));
    c.cap = reinterpret_cast<unsigned short *>(img + sp_cap_pos(c.m, c.n));
    c.len = reinterpret_cast<unsigned short *>(img + sp_len_pos(c.m, c.n));
    c.heap = reinterpret_cast<double *>(img + sp_heap_pos(c.m, c.n));
}

// ---- probe: blocks that need no pivot this round are finished here, the others go to the solve launch ----------
template <int NT>
__global__ void __launch_bounds__(NT, 1024 / NT) probe_sparse_kernel(const BlockDev *blocks, const int *list, int count,
                                                                                      SolveParams P, SparseRound R)
{
    extern __shared__ __align__(16) char smem_raw[];
    __shared__ Scratch<NT> scr;
    __shared__ DecShared dsh;
    __shared__ __align__(8) uint64_t bar;
    __shared__ int s_nl;
    if ((int)blockIdx.x >= count) return;
    const int k = list[blockIdx.x];
    const BlockDev &B = blocks[k];
    const int m = B.m, n = B.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (__ldcg(B.spflag) != 0) {        // the block lives in the dense layout now
        if (tid == 0) { R.moved[k] = 0; R.reb[k] = 0; R.dense_list[atomicAdd(R.dense_count, 1)] = k; }
        return;
    }
    Sp c;
    c.m = m; c.n = n; c.W = sp_words(n); c.WS = sp_stride(n); c.H = 0;
    char *pst;
    char *sp = sp_carve(c, smem_raw, pst, B);
    unsigned *smask = (unsigned *)sp;
    unsigned *soff = smask + (size_t)SP_PROBE_ROWS * c.WS;
    sp_image_pointers(c, reinterpret_cast<char *>(B.T));       // the HBM image
    c.stg = nullptr;

    const uint32_t p_bytes = (uint32_t)pst_bytes(m, n);
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_expect_tx(&bar, p_bytes);
        bulk_g2s(pst, B.beta, p_bytes, &bar);
    }
    sp_priced_costs<NT>(B, P, c.cs, n);
    __syncthreads();
    mbar_wait(&bar, 0);
    sp_gather_bounds<NT>(B, c);
    __syncthreads();

    const int dflag_in = B.dflag[0];
    const bool inc = dflag_in > 0 && dflag_in < 32;
    bool handoff = sp_feasibility<NT>(c, P.tol_bnd);       // an infeasible start needs phase 1
    unsigned long long n_price = 0;
    if (!handoff) {
        sp_basic_costs<NT>(B, c, inc);
        __syncthreads();
        if (warp == 0) {
            const int nl = sp_list_rows(c, c.colq, m);
            if (lane == 0) s_nl = nl;
        }
        __syncthreads();
        const int nl = s_nl;
        if (nl > SP_PROBE_ROWS) handoff = true;            // (a full recomputation reads about half the rows)
        else {
            for (int e = tid; e < nl * c.WS; e += NT) {
                const int a = e / c.WS;
                smask[e] = __ldcg(c.mask + (size_t)c.arow[a] * c.WS + (e - a * c.WS));
            }
            for (int a = tid; a < nl; a += NT) soff[a] = __ldcg(c.off + c.arow[a]);
            sp_base_row<NT>(B, c, inc);
            __syncthreads();
            if (warp == 0) n_price += sp_listed_cells(c, nl, c.len);    // (len read from the image)
            sp_weighted_sum<NT, true>(c, c.colq, nl, c.drow, smask, soff);
            __syncthreads();
            handoff = sp_price<NT>(c, c.drow, false, P.tol_dj, &dsh) >= 0;
        }
    }
    if (handoff) {
        if (tid == 0) { R.moved[k] = 0; R.reb[k] = 0; R.next_list[atomicAdd(R.next_count, 1)] = k; atomicAdd(&R.stats[7], 1ull); }
        return;
    }
    if (tid == 0) atomicAdd(&R.stats[6], 1ull);
    n_price = __shfl_sync(0xffffffffu, n_price, 0);
    sp_unchanged_epilogue<NT>(B, P, c, k, !inc, dflag_in, warp == 0 ? n_price : 0ull, scr);
}

// ---- rank-1 update of one packed row by one warp ---------------------------------------------------------------
// new[j] = old[j] - ci * rowp[j] over the union of the row's pattern and the pivot row's; column q becomes ci * inv.
// Pass 1 evaluates every candidate (values staged per warp when they fit), pass 2 writes the survivors packed.
__device__ __forceinline__ int sp_update_row(Sp &c, int i, double ci, double inv, int q, int lane, double *stg)
{
    unsigned *mr = c.mask + (size_t)i * c.WS;
    const unsigned mi = sp_chunk(mr, lane, c.W), mp = sp_chunk(c.pmask, lane, c.W);
    const unsigned u = mi | mp;
    int tot;
    const int pre = warp_excl_scan(__popc(mi) | (__popc(u) << 16), lane, tot);
    const int pre_i = pre & 0xffff, pre_u = pre >> 16, tot_u = tot >> 16;
    const unsigned oi = c.off[i];
    const int jb = lane * 16;
    const bool staged = tot_u <= SP_STG;
    const double cq = ci * inv;
    unsigned nz = 0;
    {
        unsigned bits = u;
        int ru = 0;
        while (bits) {
            const int b = __ffs(bits) - 1; bits &= bits - 1;
            const int j = jb + b;
            const double old = ((mi >> b) & 1u) ? c.heap[oi + pre_i + __popc(mi & ((1u << b) - 1u))] : 0.0;
            const double rp = ((mp >> b) & 1u) ? c.rowp[j] : 0.0;
            const double v = (j == q) ? cq : fma(-ci, rp, old);
            if (v != 0.0) nz |= 1u << b;
            if (staged) stg[pre_u + ru] = v;
            ++ru;
        }
    }
    int tot_n;
    const int pre_n = warp_excl_scan(__popc(nz), lane, tot_n);
    __syncwarp();
    unsigned dst = oi;
    if (tot_n > (int)c.cap[i] || !staged) {
        // (the caller has checked that the heap holds a fresh copy of every active row)
        const int nc = sp_newcap(tot_n, c.n);
        if (lane == 0) {
            dst = (unsigned)atomicAdd(&c.hdr[0], nc);
            atomicAdd(&c.hdr[1], nc - (int)c.cap[i]);
            c.off[i] = dst; c.cap[i] = (unsigned short)nc;
        }
        dst = __shfl_sync(0xffffffffu, dst, 0);
    }
    {
        unsigned bits = u;
        int ru = 0, rn = 0;
        while (bits) {
            const int b = __ffs(bits) - 1; bits &= bits - 1;
            if ((nz >> b) & 1u) {
                double v;
                if (staged) v = stg[pre_u + ru];
                else {
                    const int j = jb + b;
                    const double old = ((mi >> b) & 1u) ? c.heap[oi + pre_i + __popc(mi & ((1u << b) - 1u))] : 0.0;
                    const double rp = ((mp >> b) & 1u) ? c.rowp[j] : 0.0;
                    v = (j == q) ? cq : fma(-ci, rp, old);
                }
                c.heap[dst + pre_n + rn] = v;
                ++rn;
            }
            ++ru;
        }
    }
    const unsigned other = __shfl_xor_sync(0xffffffffu, nz, 1);
    if (!(lane & 1) && (lane >> 1) < c.W) mr[lane >> 1] = nz | (other << 16);
    if (lane == 0) c.len[i] = (unsigned short)tot_n;
    __syncwarp();
    return tot_n;
}

// Re-pack the heap through the HBM image (rows in index order, fresh slack or none).  All threads.
template <int NT>
__device__ void sp_compact(Sp &c, double *gheap, bool slack)
{
    constexpr int NW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, m = c.m;
    unsigned *noff = reinterpret_cast<unsigned *>(c.stg);    // m words of scratch: nobody stages a row during a re-pack
    __shared__ unsigned s_newtop;
    if (warp == 0) {
        // chunked exclusive scan of the new capacities
        const int per = (m + 31) / 32;
        int loc = 0;
        for (int t = 0; t < per; ++t) { const int i = lane * per + t; if (i < m) loc += slack ? sp_newcap(c.len[i], c.n) : (int)c.len[i]; }
        int tot;
        int run = warp_excl_scan(loc, lane, tot);
        for (int t = 0; t < per; ++t) {
            const int i = lane * per + t;
            if (i < m) { noff[i] = (unsigned)run; run += slack ? sp_newcap(c.len[i], c.n) : (int)c.len[i]; }
        }
        if (lane == 0) s_newtop = (unsigned)tot;
    }
    __syncthreads();
    for (int i = warp; i < m; i += NW) {
        const unsigned o = c.off[i], d = noff[i];
        const int l = c.len[i];
        for (int t = lane; t < l; t += 32) gheap[d + t] = c.heap[o + t];
    }
    __syncthreads();
    const int newtop = (int)s_newtop;
    for (int i = tid; i < m; i += NT) {
        const int l = c.len[i];
        c.off[i] = noff[i];
        c.cap[i] = (unsigned short)(slack ? sp_newcap(l, c.n) : l);
    }
    {   // copy back: 16-byte loads, four per thread in flight (the image is L2-resident: it was written a moment ago)
        const double2 *g2 = reinterpret_cast<const double2 *>(gheap);
        double2 *h2 = reinterpret_cast<double2 *>(c.heap);
        const int n2 = (newtop + 1) >> 1;
        for (int e0 = tid; e0 < n2; e0 += NT * 4) {
            double2 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) if (e0 + u * NT < n2) v[u] = __ldcg(g2 + e0 + u * NT);
#pragma unroll
            for (int u = 0; u < 4; ++u) if (e0 + u * NT < n2) h2[e0 + u * NT] = v[u];
        }
    }
    if (tid == 0) { c.hdr[0] = newtop; c.hdr[1] = newtop; }
    __syncthreads();
}

// block-wide sums of three ints and the maximum of a fourth through shared atomics (integers: no order dependence)
__device__ __forceinline__ void sp_block_reduce4(int &a, int &b, int &c3, int &mx, int *s4)
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o);
        c3 += __shfl_xor_sync(0xffffffffu, c3, o); mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (threadIdx.x == 0) { s4[0] = 0; s4[1] = 0; s4[2] = 0; s4[3] = 0; }
    __syncthreads();
    if (lane == 0) { atomicAdd(&s4[0], a); atomicAdd(&s4[1], b); atomicAdd(&s4[2], c3); atomicMax(&s4[3], mx); }
    __syncthreads();
    a = s4[0]; b = s4[1]; c3 = s4[2]; mx = s4[3];
    __syncthreads();
}

// non-zeros of the union of two row patterns
__device__ __forceinline__ int sp_union_size(const unsigned *ma, const unsigned *mb, int W)
{
    int u = 0;
    for (int w_ = 0; w_ < W; ++w_) u += __popc(ma[w_] | mb[w_]);
    return u;
}

// Hand-over to the dense path: every row into the row-major image of path 3.  Only the 32-byte items that hold a
// non-zero are written (whole items), the item masks say which — exactly what path 3 expects of its tableau.
template <int NT>
__device__ void sp_unpack_dense(const Sp &c, const BlockDev &B)
{
    constexpr int NW = NT / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ld = B.ldg, rwd = B.rwd;
    double2 *T2 = reinterpret_cast<double2 *>(B.T);
    for (int i = warp; i < c.m; i += NW) {
        const unsigned ch = sp_chunk(c.mask + (size_t)i * c.WS, lane, c.W);
        int tot;
        const int pre = warp_excl_scan(__popc(ch), lane, tot);
        const double *row = c.heap + c.off[i] + pre;
        int r = 0;
        unsigned nib = 0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned b4 = (ch >> (4 * u)) & 0xfu;
            if (b4) {
                double v[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) { v[t] = 0.0; if ((b4 >> t) & 1u) v[t] = row[r++]; }
                double2 *dst = T2 + ((size_t)i * ld) / 2 + (size_t)(lane * 4 + u) * 2;
                __stcg(dst, make_double2(v[0], v[1]));
                __stcg(dst + 1, make_double2(v[2], v[3]));
                nib |= 1u << u;
            }
        }
        unsigned word = nib << (4 * (lane & 7));
        word |= __shfl_xor_sync(0xffffffffu, word, 1);
        word |= __shfl_xor_sync(0xffffffffu, word, 2);
        word |= __shfl_xor_sync(0xffffffffu, word, 4);
        if ((lane & 7) == 0 && (lane >> 3) < rwd) B.rmask[(size_t)i * rwd + (lane >> 3)] = word;
    }
}

// ---- solve kernel ------------------------------------------------------------------------------------------------
// level 0: small heap (two CTAs per SM); level 1: the large heap (one CTA per SM).  H = doubles of heap in this launch.
template <int NT>
__global__ void __launch_bounds__(NT, 2) solve_sparse_kernel(const BlockDev *blocks, const int *list, int count,
                                                             SolveParams P, SparseRound R, int level, int H)
{
    constexpr int NW = NT / 32;
    extern __shared__ __align__(16) char smem_raw[];
    __shared__ Scratch<NT> scr;
    __shared__ SpDec dec;
    __shared__ DecShared dsh;
    __shared__ __align__(8) uint64_t bar;
    __shared__ int s_changed, s_nl, s_red[4];
    if ((int)blockIdx.x >= count) return;
    if (R.in_count != nullptr && (int)blockIdx.x >= *R.in_count) return;
    const int k = list[blockIdx.x];
    const BlockDev &B = blocks[k];
    const int m = B.m, n = B.n, L = B.L, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const double tol_dj = P.tol_dj, tol_bnd = P.tol_bnd, tol_piv = P.tol_piv;
    const double harris = 0.5 * tol_bnd;
    char *gimg = reinterpret_cast<char *>(B.T);
    const bool moved_in = R.use_moved && R.moved[k] != 0;
    if (__ldcg(B.spflag) != 0) {
        if (tid == 0) { R.moved[k] = moved_in ? 1 : 0; R.reb[k] = 0; R.dense_list[atomicAdd(R.dense_count, 1)] = k; }
        return;
    }
    const int top_in = __ldcg(reinterpret_cast<const int *>(gimg));
    if (top_in > H) {           // the image was left by a launch with a larger heap
        if (tid == 0) {
            R.moved[k] = moved_in ? 1 : 0; R.reb[k] = level ? 1 : 0;
            if (level) { B.spflag[0] = 1; B.dflag[0] = 0; R.dense_list[atomicAdd(R.dense_count, 1)] = k; }
            else R.next_list[atomicAdd(R.next_count, 1)] = k;
        }
        return;
    }

    phase_init();
    Sp c;
    c.m = m; c.n = n; c.W = sp_words(n); c.WS = sp_stride(n); c.H = H;
    char *pst;
    char *sp = sp_carve(c, smem_raw, pst, B);
    c.stg = (double *)sp;           sp += sizeof(double) * (size_t)NW * SP_STG;
    char *simg = sp;
    sp_image_pointers(c, simg);
    double *gheap = reinterpret_cast<double *>(gimg + sp_heap_pos(m, n));
    double *stg = c.stg + (size_t)warp * SP_STG;

    // ---- state in: one mbarrier, bulk copies of the state image and of the packed tableau
    const uint32_t p_bytes = (uint32_t)pst_bytes(m, n);
    const uint32_t head_bytes = (uint32_t)sp_heap_pos(m, n);
    if (tid == 0) {
        const uint32_t i_bytes = head_bytes + (uint32_t)align16(sizeof(double) * (size_t)top_in);
        mbar_init(&bar, 1);
        mbar_expect_tx(&bar, p_bytes + i_bytes);
        bulk_g2s(simg, gimg, i_bytes, &bar);
        bulk_g2s(pst, B.beta, p_bytes, &bar);
        s_changed = moved_in ? 1 : 0;
    }
    sp_priced_costs<NT>(B, P, c.cs, n);
    __syncthreads();
    mbar_wait(&bar, 0);
    sp_gather_bounds<NT>(B, c);
    __syncthreads();

    int status = 0;
    long long iter = 0;
    const long long max_iter = (long long)P.max_iter_mult * (m + n) + 1000;
    unsigned long long n_piv = 0, n_flip = 0, n_p1 = 0, n_rows = 0;     // warp 0
    unsigned long long n_cells = 0;                                     // per warp: doubles rewritten by this warp's row updates
    unsigned long long n_price = 0;                                     // warp 0
    bool need_check = true, infeasible = false, d_valid = false, bland = false, d_full = false;
    bool img_dirty = false;
    int degen = 0, exit_to = 0;
    bool batched = false, bad = false, need_reb = false;         // exit_to: 1 = next level (heap too small), 2 = dense path with a rebuild
    const int degen_limit = 100 + (m + n) / 2;
    const int dflag_in = B.dflag[0];
    const bool use_saved = dflag_in > 0 && dflag_in < 32;

    for (;;) {
        if (++iter > max_iter) { status = 1; break; }
        // ---- A. feasibility scan, phase-1 prices, fresh reduced costs
        phase(1);
        if (need_check || infeasible) {
            infeasible = sp_feasibility<NT>(c, tol_bnd);
            need_check = false;
            if (infeasible) d_valid = false;
        }
        if (infeasible) {
            // phase-1 prices T'w (into drow: the reduced costs are not live while the block is infeasible)
            if (warp == 0) {
                const int nl = sp_list_rows(c, c.w, m);
                if (lane == 0) s_nl = nl;
                __syncwarp();
                n_price += sp_listed_cells(c, nl, c.len);
            }
            for (int j = tid; j < n; j += NT) c.drow[j] = 0.0;
            __syncthreads();
            sp_weighted_sum<NT, false>(c, c.w, s_nl, c.drow, nullptr, nullptr);
            if (warp == 0) ++n_p1;
            __syncthreads();
        } else if (!d_valid) {
            phase(2);
            const bool inc = use_saved && !s_changed;
            sp_basic_costs<NT>(B, c, inc);
            __syncthreads();
            if (warp == 0) {
                const int nl = sp_list_rows(c, c.colq, m);
                if (lane == 0) s_nl = nl;
                __syncwarp();
                n_price += sp_listed_cells(c, nl, c.len);
            }
            sp_base_row<NT>(B, c, inc);
            __syncthreads();
            sp_weighted_sum<NT, false>(c, c.colq, s_nl, c.drow, nullptr, nullptr);
            d_valid = true;
            d_full = !inc;
            __syncthreads();
        }
        // ---- B. decision
        const bool INF = infeasible;
        const bool d_live = !infeasible;
        int kind;
        for (;;) {
            phase(3);
            const int q = sp_price<NT>(c, c.drow, bland, tol_dj, &dsh);
            if (q < 0) { kind = K_NONE; break; }
            phase(4);
            const double dq = c.drow[q];
            const double s = dq < 0.0 ? 1.0 : -1.0;
            const signed char stq = c.cstat[q];
            const int vq = c.nbv[q];
            const double loq = __ldg(c.glo + vq), upq = __ldg(c.gup + vq);
            // pivot column: one thread per row tests the row's bit for q and fetches the value by its rank
            const int RW = ((m + NW - 1) / NW + 31) & ~31;
            int cnt = 0;
            {
                const int rbeg = warp * RW, rend = min(m, rbeg + RW);
                for (int i0 = rbeg; i0 < rend; i0 += 32) {
                    const int i = i0 + lane;
                    double v = 0.0;
                    if (i < rend) {
                        const unsigned *mr = c.mask + (size_t)i * c.WS;
                        if ((mr[q >> 5] >> (q & 31)) & 1u) v = c.heap[c.off[i] + sp_rank(mr, q)];
                    }
                    const bool nz = v != 0.0;
                    const unsigned bal = __ballot_sync(0xffffffffu, nz);
                    if (nz) {
                        c.colq[i] = v;
                        c.arow2[rbeg + cnt + __popc(bal & lt_mask)] = (unsigned short)i;
                    }
                    cnt += __popc(bal);
                }
                if (lane == 0) dsh.cnt[warp] = cnt;
            }
            __syncthreads();
            int na = 0;
            {
                int my_off = 0;
#pragma unroll
                for (int w_ = 0; w_ < NW; ++w_) { if (w_ == warp) my_off = na; na += dsh.cnt[w_]; }
                for (int e = lane; e < cnt; e += 32) c.arow[my_off + e] = c.arow2[warp * RW + e];
            }
            __syncthreads();
            // ratio test and flips over the list: warp 0 (same rules as path 3); a basis change is only CHOSEN here
            phase(5);
            double theta = 0.0;
            int p = -1, pe = -1;
            if (warp == 0) {
                int kd = K_PIVOT;
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
                    c.w[e] = tex;
                }
                tmax = warp_min_nonneg(tmax);
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
                theta = p >= 0 ? fmax(0.0, c.w[pe]) : CUDART_INF;
                const double range = (stq == ST_FREE) ? CUDART_INF : upq - loq;
                if (range <= theta) {
                    if (!(range < BIG)) { if (lane == 0) dec.fail_status = INF ? 1 : 6; kd = K_FAIL; }
                    else {
                        for (int e = lane; e < na; e += 32) {
                            const int i = c.arow[e];
                            c.beta[i] = fma(s * range, c.colq[i], c.beta[i]);
                        }
                        if (lane == 0) {
                            const signed char ns = (stq == ST_LO) ? ST_UP : ST_LO;
                            c.cstat[q] = ns;
                            c.vstat[vq] = ns;
                            s_changed = 1;
                        }
                        ++n_flip;
                        degen = 0; bland = false;
                        kd = INF ? K_REDO : K_FLIPPED;
                    }
                } else if (p < 0) {
                    if (lane == 0) dec.fail_status = INF ? 1 : 6;
                    kd = K_FAIL;
                } else {
                    // room for a fresh copy of every active row (an upper bound of what the sweep can allocate)
                    const int lenp = c.len[p];
                    int need = 0;
                    for (int e = lane; e < na; e += 32) {
                        const int i = c.arow[e];
                        if (i != p) need += sp_newcap(min(n, (int)c.len[i] + lenp - 1), n);
                    }
#pragma unroll
                    for (int o = 16; o; o >>= 1) need += __shfl_xor_sync(0xffffffffu, need, o);
                    if (lane == 0) { dec.need = need; dec.p = p; dec.pe = pe; dec.q = q; dec.na = na; dec.lenp = lenp; }
                }
                if (lane == 0) { dec.kind = kd; dec.bland = bland ? 1 : 0; }
            }
            __syncthreads();
            kind = dec.kind;
            bland = dec.bland != 0;
            if (kind == K_FLIPPED) continue;
            if (kind != K_PIVOT) break;
            batched = false;
            phase(6);
            if (c.hdr[0] + dec.need > H) {      // block-uniform; nothing has been modified yet
                // The cheap bound (len_i + len_p per active row) does not fit: count the union patterns.
                //   need_a   what the sweep can allocate with the rows where they are (a union that fits its slot stays)
                //   need_all the same if every active row has to move (after a re-pack)
                //   grow     by how much the tableau can grow in total, maxu = the largest single new row
                int need_a = 0, need_all = 0, grow = 0, maxu = 0;
                const unsigned *mpp = c.mask + (size_t)dec.p * c.WS;
                for (int e = tid; e < dec.na; e += NT) {
                    const int i = c.arow[e];
                    if (i == dec.p) continue;
                    const int u = sp_union_size(c.mask + (size_t)i * c.WS, mpp, c.W);
                    const int nc = sp_newcap(u, n);
                    need_all += nc;
                    if (u > (int)c.cap[i] || u > SP_STG) need_a += nc;
                    grow += max(0, u - (int)c.len[i]);
                    maxu = max(maxu, nc);
                }
                sp_block_reduce4(need_a, need_all, grow, maxu, s_red);
                if (tid == 0) atomicAdd(&R.stats[5], 1ull);
                if (c.hdr[0] + need_a > H) {
                    // re-pack the heap through the HBM image: with fresh slack if a copy of every active row still fits
                    // then, without otherwise; if not even that, sweep in batches with a re-pack between them (needs the
                    // final tableau plus one row); and if the tableau itself outgrows the heap the block moves on
                    int live_s = 0, live_0 = 0, z0 = 0, z1 = 0;
                    for (int i = tid; i < m; i += NT) { const int l = c.len[i]; live_s += sp_newcap(l, n); live_0 += l; }
                    sp_block_reduce4(live_s, live_0, z0, z1, s_red);
                    if (live_s + need_all <= H) sp_compact<NT>(c, gheap, true);
                    else if (live_0 + need_all <= H) sp_compact<NT>(c, gheap, false);
                    else if (live_0 + grow + maxu <= H) { sp_compact<NT>(c, gheap, false); batched = true; }
                    else { kind = K_TIGHT; break; }
                    img_dirty = true;
                    if (tid == 0) { atomicAdd(&R.stats[3], 1ull); if (batched) atomicAdd(&R.stats[4], 1ull); }
                }
            }
            // ---- the basis change: warp 0 does the bookkeeping, warp 1 (or 0) expands the scaled pivot row
            phase(11);
            const int pp = dec.p;
            const double piv = c.colq[pp];
            const double inv = __drcp_rn(piv);
            if (warp == 0) {
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
                    c.arow[pe] = c.arow[0]; c.arow[0] = (unsigned short)p;
                    const signed char ns = (lov == upv) ? ST_FIX : (target == lov ? ST_LO : ST_UP);
                    const int vleave = c.head[p];
                    c.rlo[p] = loq; c.rup[p] = upq;
                    c.cstat[q] = ns; c.vstat[vleave] = ns; c.vstat[vq] = ST_BASIC;
                    c.head[p] = vq; c.nbv[q] = vleave;
                    dec.inv = inv; dec.dq = (d_live && dq != 0.0) ? dq : 0.0;
                    s_changed = 1;
                }
                ++n_piv;
                n_rows += (unsigned long long)na;
                if (theta <= 1e-11) { if (++degen > degen_limit) bland = true; }
                else { degen = 0; bland = false; }
                if (lane == 0) dec.bland = bland ? 1 : 0;
            }
            if (warp == (NW > 1 ? 1 : 0)) {
                // row p: rowp[j] = T[p][j] / piv for the pivot row's pattern, and in place T'[p][j] = -rowp[j], T'[p][q] = 1 / piv
                const unsigned *mr = c.mask + (size_t)pp * c.WS;
                const unsigned mp = sp_chunk(mr, lane, c.W);
                int tot;
                const int pre = warp_excl_scan(__popc(mp), lane, tot);
                double *row = c.heap + c.off[pp] + pre;
                unsigned bits = mp;
                int r = 0;
                while (bits) {
                    const int b = __ffs(bits) - 1; bits &= bits - 1;
                    const int j = lane * 16 + b;
                    const double sv = row[r] * inv;
                    c.rowp[j] = sv;
                    row[r] = (j == q) ? inv : -sv;
                    ++r;
                }
                if (lane < c.W) c.pmask[lane] = mr[lane];
                n_cells += (unsigned long long)tot;
            }
            __syncthreads();
            bland = dec.bland != 0;
            break;
        }
        if (kind == K_PIVOT) {
            phase(batched ? 9 : 7);
            const int na = dec.na, q = dec.q;
            const double inv = dec.inv;
            if (warp == NW - 1 && dec.dq != 0.0) {
                // the reduced-cost row rides along: d[j] -= d_q rowp[j] over the pivot row's pattern, d[q] = d_q / piv
                const double dq = dec.dq;
                unsigned bits = sp_chunk(c.pmask, lane, c.W);
                while (bits) {
                    const int b = __ffs(bits) - 1; bits &= bits - 1;
                    const int j = lane * 16 + b;
                    c.drow[j] = (j == q) ? dq * inv : fma(-dq, c.rowp[j], c.drow[j]);
                }
            }
            if (!batched) {
                for (int e = 1 + warp; e < na; e += NW) {
                    const int i = c.arow[e];
                    n_cells += (unsigned long long)sp_update_row(c, i, c.colq[i], inv, q, lane, stg);
                }
            } else {
                // not enough room for a copy of every active row at once: as many rows as fit, re-pack, repeat
                unsigned short *usz = c.arow2;
                for (int e = 1 + tid; e < na; e += NT)
                    usz[e] = (unsigned short)sp_union_size(c.mask + (size_t)c.arow[e] * c.WS, c.pmask, c.W);
                __syncthreads();
                int e0 = 1;
                while (e0 < na) {
                    if (tid == 0) {
                        int room = H - c.hdr[0], e1 = e0;
                        while (e1 < na) {
                            const int nc = sp_newcap(usz[e1], n);
                            if (nc > room) break;
                            room -= nc; ++e1;
                        }
                        s_nl = e1;
                    }
                    __syncthreads();
                    const int e1 = s_nl;
                    if (e1 == e0) { bad = true; break; }      // cannot happen (the check above reserved the largest row)
                    for (int e = e0 + warp; e < e1; e += NW) {
                        const int i = c.arow[e];
                        n_cells += (unsigned long long)sp_update_row(c, i, c.colq[i], inv, q, lane, stg);
                    }
                    __syncthreads();
                    e0 = e1;
                    if (e0 < na) sp_compact<NT>(c, gheap, false);
                }
            }
            img_dirty = true;
            __syncthreads();
            if (bad) { status = 1; break; }
            continue;
        }
        if (kind == K_TIGHT) {         // the tableau outgrows this launch's heap (nothing of the chosen pivot has been applied)
            exit_to = level ? 2 : 1;
            break;
        }
        if (kind == K_REDO) continue;
        if (kind == K_FAIL) { status = dec.fail_status; break; }
        // ---- K_NONE
        if (infeasible) { status = 4; break; }
        if (!s_changed && !P.initial && !P.force_full) { status = 5; break; }
        // claimed optimal: x into rowp, auxiliary values into colq
        phase(8);
        for (int j = tid; j < n; j += NT) {
            const int v = c.nbv[j];
            const double xv = sp_nb_col_value(c, j);
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
            for (int e = B.a_rp[i]; e < B.a_rp[i + 1]; ++e) s = fma(B.a_v[e], c.rowp[B.a_ci[e]], s);
            worst = fmax(worst, fabs(s - c.colq[i]) / (1.0 + fabs(s)));
        }
        worst = block_max<NT>(worst, scr);
        if (worst > 0.1 * tol_bnd) { exit_to = 2; need_reb = true; break; }     // the dense kernel rebuilds the tableau from A_k
        status = 5;
        break;
    }

    // ---- epilogue
    __syncthreads();
    phase(10);
    n_piv = __shfl_sync(0xffffffffu, n_piv, 0); n_flip = __shfl_sync(0xffffffffu, n_flip, 0);
    if (exit_to != 0) {
        // hand the block over with a consistent state image in HBM
        const bool moved = s_changed != 0;
        if (exit_to == 2 && !need_reb) sp_unpack_dense<NT>(c, B);
        if (tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (img_dirty && exit_to == 1)
                bulk_s2g(gimg, simg, head_bytes + (uint32_t)align16(sizeof(double) * (size_t)c.hdr[0]));
            if (moved) bulk_s2g(B.beta, pst, p_bytes);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            B.dflag[0] = 0;
            R.moved[k] = moved ? 1 : 0;
            R.reb[k] = need_reb ? 1 : 0;
            if (exit_to == 2) { B.spflag[0] = 1; R.dense_list[atomicAdd(R.dense_count, 1)] = k; atomicAdd(&R.stats[need_reb ? 2 : 1], 1ull); }
            else { R.next_list[atomicAdd(R.next_count, 1)] = k; atomicAdd(&R.stats[0], 1ull); }
        }
        // counters: every warp adds its share
        if (lane == 0) atomicAdd(&P.counters[8 * (size_t)k + 5], n_cells);
        if (tid == 0) {
            atomicAdd(&P.counters[8 * (size_t)k + 0], n_piv);
            atomicAdd(&P.counters[8 * (size_t)k + 1], n_flip);
            atomicAdd(&P.counters[8 * (size_t)k + 3], n_p1);
            atomicAdd(&P.counters[8 * (size_t)k + 4], n_rows);
            atomicAdd(&P.counters[8 * (size_t)k + 6], n_price);
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
        phase_flush();
        return;
    }
    const bool changed = s_changed != 0 || P.initial != 0 || P.force_full != 0 || status != 5;
    if (status != 5) {
        for (int j = tid; j < n; j += NT) {
            const int v = c.nbv[j];
            if (v >= m) c.rowp[v - m] = sp_nb_col_value(c, j);
        }
        for (int i = tid; i < m; i += NT) {
            const int v = c.head[i];
            if (v >= m) c.rowp[v - m] = c.beta[i];
        }
        __syncthreads();
    }
    if (!changed) {
        sp_unchanged_epilogue<NT>(B, P, c, k, d_full, dflag_in, warp == 0 ? n_price : 0ull, scr);
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
    if (tid == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (img_dirty) bulk_s2g(gimg, simg, head_bytes + (uint32_t)align16(sizeof(double) * (size_t)c.hdr[0]));
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
    if (lane == 0) atomicAdd(&P.counters[8 * (size_t)k + 5], n_cells);
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
            atomicAdd(&P.counters[8 * (size_t)k + 0], n_piv);
            atomicAdd(&P.counters[8 * (size_t)k + 1], n_flip);
            atomicAdd(&P.counters[8 * (size_t)k + 3], n_p1);
            atomicAdd(&P.counters[8 * (size_t)k + 4], n_rows);
            atomicAdd(&P.counters[8 * (size_t)k + 6], n_price);
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
