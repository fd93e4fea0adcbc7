// master_ipm.cu — the restricted master LP solved ON THE DEVICE by a primal-dual interior-point method.
//
// SURVEY.md §8f-1 / VERDICT r1 item 4: with the subproblems at 3 ms a round, the host simplex master (~250 k
// iterations of an L x L working basis, 5 s at config B) is what decides DW time-to-optimal.  The reference solves
// its master with glp_simplex (/root/reference/src/dw_phases.c:219, 450); this file replaces that call for large
// instances by Mehrotra's predictor-corrector method, whose work per iteration is massively parallel:
//
//   min c'x  s.t.  A x = b, x >= 0,   A = [A_L ; E]   (L linking rows; E = block indicator: K convexity rows)
//
// The convexity rows make the (L+K) x (L+K) normal equations A D A' block-structured (GUB): the K x K part is the
// diagonal delta_k = sum_{j in k} d_j, and eliminating it leaves the L x L Schur complement
//   S = sum_k sum_{j in k} d_j (a_j - abar_k)(a_j - abar_k)' + sum_{j in no block} d_j a_j a_j',  abar_k = sum d_j a_j / delta_k
// i.e. S = Bt' Bt for the n x L matrix Bt of centred, scaled columns: a genuine FP64 contraction over the n master
// columns (30 k - 100 k at config B), done here with FP64 tensor-core MMAs (mma.sync m8n8k4 f64 -> DMMA in SASS),
// followed by an in-shared-memory panel Cholesky of the 256 x 256 system.  Everything else is O(nnz) vector work.
// All sums run in a fixed order (no floating-point atomics): a solve is bit-reproducible.
#include <cuda_runtime.h>
#include <math_constants.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/dwgpu.h"

namespace {

thread_local std::string g_merr;
int mfail(int code, const std::string &msg) { g_merr = msg; return code; }
#define MCU(call)                                                                              \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) return mfail(DWGPU_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

constexpr int RED_CTAS = 256, RED_NT = 256;      // fixed reduction geometry: deterministic sums

// ---------------------------------------------------------------------------------------------------------------
// reductions: per-CTA partials over a fixed chunking, then one CTA folds the partials in a fixed tree
__device__ __forceinline__ double cta_sum(double v, double *sh)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sh[w] = v;
    __syncthreads();
    v = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
    if (w == 0) {
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    }
    __syncthreads();
    return v;           // valid in warp 0
}
__device__ __forceinline__ double cta_max(double v, double *sh)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) sh[w] = v;
    __syncthreads();
    v = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : -CUDART_INF;
    if (w == 0) {
#pragma unroll
        for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    }
    __syncthreads();
    return v;
}

// scalars of an iteration (device array, copied to pinned host memory once per iteration)
enum { SC_XS = 0, SC_PINF_L, SC_PINF_K, SC_DINF, SC_POBJ, SC_DOBJ, SC_AP, SC_AD, SC_MUAFF, SC_MINX, SC_MINS, SC_SUMX, SC_SUMS,
       SC_XS2, SC_COUNT = 16 };

// ---- per-column vector kernels ------------------------------------------------------------------------------
__global__ void k_dvec(int n, const double *x, const double *s, double *d)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) d[j] = x[j] / s[j];
}

// one CTA per block k: delta_k and abar_k = sum_j d_j a_j / delta_k over the block's columns, in column order
__global__ void k_block_prep(int L, const int *blkptr, const int *blkcols, const int *colptr, const int *rowidx, const double *val,
                             const double *d, double *delta, double *abar)
{
    extern __shared__ double g[];
    const int k = blockIdx.x;
    for (int i = threadIdx.x; i < L; i += blockDim.x) g[i] = 0.0;
    __syncthreads();
    double dl = 0.0;
    for (int e = blkptr[k]; e < blkptr[k + 1]; ++e) {
        const int j = blkcols[e];
        const double dj = d[j];
        dl += dj;
        for (int t = colptr[j] + threadIdx.x; t < colptr[j + 1]; t += blockDim.x) g[rowidx[t]] += dj * val[t];
        __syncthreads();
    }
    if (threadIdx.x == 0) delta[k] = dl;
    const double inv = dl > 0.0 ? 1.0 / dl : 0.0;
    for (int i = threadIdx.x; i < L; i += blockDim.x) abar[(size_t)k * L + i] = g[i] * inv;
}

// Bt[j][:] = sqrt(d_j) (a_j - abar_blk(j)); one CTA per column
__global__ void k_build_bt(int L, const int *colptr, const int *rowidx, const double *val, const int *blk, const double *d,
                           const double *abar, double *Bt)
{
    const int j = blockIdx.x;
    const double sd = sqrt(d[j]);
    const int k = blk[j];
    double *row = Bt + (size_t)j * L;
    for (int i = threadIdx.x; i < L; i += blockDim.x) row[i] = k >= 0 ? -sd * abar[(size_t)k * L + i] : 0.0;
    __syncthreads();
    for (int t = colptr[j] + threadIdx.x; t < colptr[j + 1]; t += blockDim.x) row[rowidx[t]] += sd * val[t];
}

// ---- S = Bt' Bt with FP64 tensor-core MMAs ----------------------------------------------------------------------
// grid (tile pairs of the lower triangle in 64 x 64 blocks, column chunks); each CTA (8 warps) accumulates its 64 x 64
// block over its chunk of columns and writes the partial; syrk_reduce sums the chunks in order.
__device__ __forceinline__ void dmma_884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
constexpr int SY_T = 64, SY_KS = 32, SY_PITCH = SY_T + 4;
__global__ void __launch_bounds__(256) k_syrk_dmma(int L, int n, int chunk, const double *Bt, double *part)
{
    __shared__ double sa[SY_KS][SY_PITCH], sb[SY_KS][SY_PITCH];
    // tile pair (bi >= bj) from the linear index
    int p = blockIdx.x, bi = 0;
    while ((bi + 1) * (bi + 2) / 2 <= p) ++bi;
    const int bj = p - bi * (bi + 1) / 2;
    const int I0 = bi * SY_T, J0 = bj * SY_T;
    const int j_begin = blockIdx.y * chunk, j_end = min(n, j_begin + chunk);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ar = lane >> 2, ak = lane & 3;
    double acc[8][2];
#pragma unroll
    for (int c = 0; c < 8; ++c) { acc[c][0] = 0.0; acc[c][1] = 0.0; }
    for (int j0 = j_begin; j0 < j_end; j0 += SY_KS) {
        // stage SY_KS columns: rows I0.. and J0.. of each (zero beyond L / beyond the chunk)
        for (int e = threadIdx.x; e < SY_KS * SY_T; e += 256) {
            const int kk = e / SY_T, i = e - kk * SY_T;
            const int j = j0 + kk;
            double va = 0.0, vb = 0.0;
            if (j < j_end) {
                if (I0 + i < L) va = Bt[(size_t)j * L + I0 + i];
                if (J0 + i < L) vb = Bt[(size_t)j * L + J0 + i];
            }
            sa[kk][i] = va; sb[kk][i] = vb;
        }
        __syncthreads();
#pragma unroll
        for (int k4 = 0; k4 < SY_KS; k4 += 4) {
            const double a = sa[k4 + ak][warp * 8 + ar];
#pragma unroll
            for (int c = 0; c < 8; ++c) dmma_884(acc[c][0], acc[c][1], a, sb[k4 + ak][c * 8 + ar]);
        }
        __syncthreads();
    }
    double *out = part + ((size_t)blockIdx.y * gridDim.x + p) * (SY_T * SY_T);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int r = warp * 8 + ar, cc = c * 8 + 2 * ak;
        out[r * SY_T + cc] = acc[c][0];
        out[r * SY_T + cc + 1] = acc[c][1];
    }
}
__global__ void k_syrk_reduce(int L, int npairs, int nchunks, const double *part, double *S)
{
    const int p = blockIdx.x;
    int bi = 0;
    while ((bi + 1) * (bi + 2) / 2 <= p) ++bi;
    const int bj = p - bi * (bi + 1) / 2;
    for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < SY_T * SY_T; e += gridDim.y * blockDim.x) {
        double s = 0.0;
        for (int c = 0; c < nchunks; ++c) s += part[((size_t)c * npairs + p) * (SY_T * SY_T) + e];
        const int i = bi * SY_T + e / SY_T, j = bj * SY_T + e % SY_T;
        if (bi == bj && i < j) continue;          // diagonal blocks: the lower half decides both halves
        if (i < L && j < L) { S[(size_t)i * L + j] = s; S[(size_t)j * L + i] = s; }
    }
}

// ---- Cholesky of S (lower triangle, in place), one CTA, panels of 32 columns factored in shared memory ------------
constexpr int CH_P = 32, CH_LD = CH_P + 1;        // panel width and its shared-memory pitch (odd: no bank conflicts)
__global__ void __launch_bounds__(1024) k_cholesky(int L, double *S, double reg_rel)
{
    extern __shared__ double pan[];          // (L - p0) x CH_P, row-major with pitch CH_LD
    __shared__ double red[32];
    __shared__ double s_reg;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    {
        double tr = 0.0;
        for (int i = tid; i < L; i += nt) tr += S[(size_t)i * L + i];
        tr = cta_sum(tr, red);
        if (tid == 0) s_reg = reg_rel * fmax(tr / L, 1e-300);
        __syncthreads();
        for (int i = tid; i < L; i += nt) S[(size_t)i * L + i] += s_reg;
        __syncthreads();
    }
    for (int p0 = 0; p0 < L; p0 += CH_P) {
        const int pw = min(CH_P, L - p0), rows = L - p0;
        for (int e = tid; e < rows * pw; e += nt) { const int r = e / pw, c = e - r * pw; pan[r * CH_LD + c] = S[(size_t)(p0 + r) * L + p0 + c]; }
        __syncthreads();
        // (a) the pw x pw diagonal block: one warp, lane = row, warp-synchronous
        //     (measured: register-resident variants of (a) and (b) need > 64 registers, i.e. 256-thread CTAs, and lose more
        //      in the trailing update than they gain here: 0.58 ms against 0.36 ms per factorisation)
        if (tid < 32) {
            for (int k = 0; k < pw; ++k) {
                const double dkk = pan[k * CH_LD + k];
                const double piv = dkk > 0.0 ? sqrt(dkk) : 1e150;       // a non-positive pivot: the direction is dropped
                __syncwarp();
                double lrk = 0.0;
                if (lane >= k && lane < pw) { lrk = (lane == k) ? piv : pan[lane * CH_LD + k] / piv; pan[lane * CH_LD + k] = lrk; }
                __syncwarp();
                if (lane > k && lane < pw)
                    for (int c = k + 1; c <= lane; ++c) pan[lane * CH_LD + c] -= lrk * pan[c * CH_LD + k];
                __syncwarp();
            }
        }
        __syncthreads();
        // (b) the rows below it: L21 = A21 L11^-T, one thread per row (forward substitution along the row)
        for (int r = pw + tid; r < rows; r += nt) {
            double *row = pan + r * CH_LD;
            for (int c = 0; c < pw; ++c) {
                double v = row[c];
                const double *lc = pan + c * CH_LD;
                for (int k = 0; k < c; ++k) v -= row[k] * lc[k];
                row[c] = v / lc[c];
            }
        }
        __syncthreads();
        for (int e = tid; e < rows * pw; e += nt) { const int r = e / pw, c = e - r * pw; if (r >= c) S[(size_t)(p0 + r) * L + p0 + c] = pan[r * CH_LD + c]; }
        // (c) trailing update: S[i][j] -= sum_k pan[i][k] pan[j][k], i >= j >= p0 + pw.  4 x 4 register tiles (8 shared loads
        // per 16 FMAs: the plain dot-product form is bound by shared-memory bandwidth); a tile's rows are TT apart so that
        // neighbouring threads read neighbouring panel rows (pitch 33: conflict-free)
        const int t0 = pw, tn = rows - pw;
        const int TT = (tn + 3) / 4;
        for (int e = tid; e < TT * TT; e += nt) {
            const int ti = e / TT, tj = e - ti * TT;
            double acc[4][4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) acc[u][v] = 0.0;
            const double *pa[4], *pb[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                pa[u] = pan + (t0 + min(ti + u * TT, tn - 1)) * CH_LD;
                pb[u] = pan + (t0 + min(tj + u * TT, tn - 1)) * CH_LD;
            }
            for (int k = 0; k < pw; ++k) {
                double a[4], b[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { a[u] = pa[u][k]; b[u] = pb[u][k]; }
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int v = 0; v < 4; ++v) acc[u][v] += a[u] * b[v];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const int ri = ti + u * TT, rj = tj + v * TT;
                    if (ri < tn && rj < tn && ri >= rj) S[(size_t)(p0 + t0 + ri) * L + p0 + t0 + rj] -= acc[u][v];
                }
        }
        __syncthreads();
    }
}

// x = (L L')^-1 b for the factor stored in the lower triangle of S; one CTA; b and x in global, work in shared.
// Each 32 x 32 diagonal block is staged in shared memory before its 32 dependent steps (one warp); the off-diagonal
// part of a panel is applied by all threads.
__global__ void __launch_bounds__(1024) k_chol_solve(int L, const double *S, const double *b, double *x)
{
    extern __shared__ double y[];            // L
    __shared__ double dg[CH_P][CH_LD];
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < L; i += nt) y[i] = b[i];
    // forward: L y = b, panel by panel
    for (int p0 = 0; p0 < L; p0 += CH_P) {
        const int pw = min(CH_P, L - p0);
        for (int e = tid; e < pw * pw; e += nt) { const int r = e / pw, c = e - r * pw; dg[r][c] = S[(size_t)(p0 + r) * L + p0 + c]; }
        __syncthreads();
        if (tid < 32) {
            for (int k = 0; k < pw; ++k) {
                const double yk = y[p0 + k] / dg[k][k];
                __syncwarp();
                if (tid == k) y[p0 + k] = yk;
                if (tid > k && tid < pw) y[p0 + tid] -= dg[tid][k] * yk;
                __syncwarp();
            }
        }
        __syncthreads();
        for (int i = p0 + pw + tid; i < L; i += nt) {
            const double *row = S + (size_t)i * L + p0;
            double s = 0.0;
            for (int k = 0; k < pw; ++k) s += row[k] * y[p0 + k];
            y[i] -= s;
        }
        __syncthreads();
    }
    // backward: L' x = y
    for (int p1 = ((L - 1) / CH_P) * CH_P; p1 >= 0; p1 -= CH_P) {
        const int pw = min(CH_P, L - p1);
        for (int e = tid; e < pw * pw; e += nt) { const int r = e / pw, c = e - r * pw; dg[r][c] = S[(size_t)(p1 + r) * L + p1 + c]; }
        __syncthreads();
        if (tid < 32) {
            for (int k = pw - 1; k >= 0; --k) {
                const double xk = y[p1 + k] / dg[k][k];
                __syncwarp();
                if (tid == k) y[p1 + k] = xk;
                if (tid < k) y[p1 + tid] -= dg[k][tid] * xk;
                __syncwarp();
            }
        }
        __syncthreads();
        for (int i = tid; i < p1; i += nt) {
            double s = 0.0;
            for (int k = 0; k < pw; ++k) s += S[(size_t)(p1 + k) * L + i] * y[p1 + k];
            y[i] -= s;
        }
        __syncthreads();
    }
    for (int i = tid; i < L; i += nt) x[i] = y[i];
}

// ---- Newton step pieces --------------------------------------------------------------------------------------
// mode 0: rc = -x s (affine);  mode 1: rc = sigma mu - x s - dxa dsa (corrector);  mode 2: rc given in `rcin`
// h = d rd - rc / s ; rcs = rc / s
__global__ void k_h(int n, int mode, const double *sigmu, const double *x, const double *s, const double *d, const double *rd,
                    const double *dxa, const double *dsa, const double *rcin, double *h, double *rcs)
{
    const double sm = mode == 1 ? sigmu[0] : 0.0;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        double rc;
        if (mode == 0) rc = -x[j] * s[j];
        else if (mode == 1) rc = sm - x[j] * s[j] - dxa[j] * dsa[j];
        else rc = rcin[j];
        const double q = rc / s[j];
        rcs[j] = q;
        h[j] = d[j] * rd[j] - q;
    }
}
// per block: u_k = rpK_k + sum_{j in k} h_j ; t_k = u_k / delta_k
__global__ void k_block_u(int K, const int *blkptr, const int *blkcols, const double *h, const double *rpK, const double *delta,
                          double *u, double *t)
{
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < K; k += gridDim.x * blockDim.x) {
        double sK = 0.0;
        for (int e = blkptr[k]; e < blkptr[k + 1]; ++e) sK += h[blkcols[e]];
        const double uk = rpK[k] + sK;
        u[k] = uk;
        t[k] = delta[k] > 0.0 ? uk / delta[k] : 0.0;
    }
}
// w_j = h_j - d_j t_blk(j)
__global__ void k_w(int n, const int *blk, const double *h, const double *d, const double *t, double *w)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const int k = blk[j];
        w[j] = k >= 0 ? h[j] - d[j] * t[k] : h[j];
    }
}
// out_i = base_i + sign * sum over row i of A_L of val * w[col]; one CTA per row, fixed reduction tree
__global__ void k_row_reduce(const int *rowptr, const int *colidx, const double *rval, const double *w, const double *base, double sign,
                             double *out)
{
    __shared__ double sh[32];
    const int i = blockIdx.x;
    double s = 0.0;
    for (int t = rowptr[i] + threadIdx.x; t < rowptr[i + 1]; t += blockDim.x) s += rval[t] * w[colidx[t]];
    s = cta_sum(s, sh);
    if (threadIdx.x == 0) out[i] = (base ? base[i] : 0.0) + sign * s;
}
// v_j = a_j' yL
__global__ void k_col_dot(int n, const int *colptr, const int *rowidx, const double *val, const double *yL, double *v)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int t = colptr[j]; t < colptr[j + 1]; ++t) s += val[t] * yL[rowidx[t]];
        v[j] = s;
    }
}
// dyK_k = (u_k - sum_{j in k} d_j v_j) / delta_k
__global__ void k_block_dy(int K, const int *blkptr, const int *blkcols, const double *d, const double *v, const double *u,
                           const double *delta, double *dyK)
{
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < K; k += gridDim.x * blockDim.x) {
        double sK = 0.0;
        for (int e = blkptr[k]; e < blkptr[k + 1]; ++e) { const int j = blkcols[e]; sK += d[j] * v[j]; }
        dyK[k] = delta[k] > 0.0 ? (u[k] - sK) / delta[k] : 0.0;
    }
}
// ds = rd - v - dyK[blk] ; dx = rcs - d ds ; and the largest steps keeping x + a dx > 0, s + a ds > 0 (per-CTA minima)
__global__ void k_dxds(int n, const int *blk, const double *rd, const double *v, const double *dyK, const double *rcs, const double *d,
                       const double *x, const double *s, double *dx, double *ds, double *part /* 2 x gridDim */)
{
    __shared__ double sh[32];
    double ap = CUDART_INF, ad = CUDART_INF;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const int k = blk[j];
        const double dsj = rd[j] - v[j] - (k >= 0 ? dyK[k] : 0.0);
        const double dxj = rcs[j] - d[j] * dsj;
        dx[j] = dxj; ds[j] = dsj;
        if (dxj < 0.0) ap = fmin(ap, -x[j] / dxj);
        if (dsj < 0.0) ad = fmin(ad, -s[j] / dsj);
    }
    const double mp = -cta_max(-ap, sh);
    const double md = -cta_max(-ad, sh);
    if (threadIdx.x == 0) { part[blockIdx.x] = mp; part[gridDim.x + blockIdx.x] = md; }
}
__global__ void k_fold_min2(int parts, const double *part, double *scal, double eta)
{
    __shared__ double sh[32];
    double a = CUDART_INF, b = CUDART_INF;
    for (int i = threadIdx.x; i < parts; i += blockDim.x) { a = fmin(a, part[i]); b = fmin(b, part[parts + i]); }
    a = -cta_max(-a, sh);
    b = -cta_max(-b, sh);
    if (threadIdx.x == 0) { scal[SC_AP] = fmin(1.0, eta * a); scal[SC_AD] = fmin(1.0, eta * b); }
}
// sum (x + ap dx)(s + ad ds)
__global__ void k_muaff(int n, const double *x, const double *s, const double *dx, const double *ds, const double *scal, double *part)
{
    __shared__ double sh[32];
    const double ap = scal[SC_AP], ad = scal[SC_AD];
    const int per = (n + gridDim.x - 1) / gridDim.x;
    const int j0 = blockIdx.x * per, j1 = min(n, j0 + per);
    double a = 0.0;
    for (int j = j0 + threadIdx.x; j < j1; j += blockDim.x) a += (x[j] + ap * dx[j]) * (s[j] + ad * ds[j]);
    a = cta_sum(a, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = a;
}
__global__ void k_fold_sum(int parts, const double *part, double *dst)
{
    __shared__ double sh[32];
    double a = 0.0;
    for (int i = threadIdx.x; i < parts; i += blockDim.x) a += part[i];
    a = cta_sum(a, sh);
    if (threadIdx.x == 0) dst[0] = a;
}
// sigma mu from mu_aff: sigmu[0] = (mu_aff / mu)^3 mu
__global__ void k_sigma(int n, double *scal, double *sigmu)
{
    const double mu = scal[SC_XS] / n, mua = scal[SC_MUAFF] / n;
    const double r = mu > 0.0 ? mua / mu : 0.0;
    sigmu[0] = r * r * r * mu;
}
__global__ void k_update(int n, int L, int K, const double *scal, double *x, double *s, const double *dx, const double *ds,
                         double *yL, const double *dyL, double *yK, const double *dyK)
{
    const double ap = scal[SC_AP], ad = scal[SC_AD];
    const int tot = n + L + K;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += gridDim.x * blockDim.x) {
        if (e < n) { x[e] += ap * dx[e]; s[e] += ad * ds[e]; }
        else if (e < n + L) yL[e - n] += ad * dyL[e - n];
        else yK[e - n - L] += ad * dyK[e - n - L];
    }
}
// residuals that live per column: rd = c - a_j'yL - yK[blk] - s ; partial sums of x's, c'x ; max |rd|
__global__ void k_resid_cols(int n, const int *colptr, const int *rowidx, const double *val, const int *blk, const double *c,
                             const double *yL, const double *yK, const double *x, const double *s, double *rd, double *part /* 3 x grid */)
{
    __shared__ double sh[32];
    const int per = (n + gridDim.x - 1) / gridDim.x;
    const int j0 = blockIdx.x * per, j1 = min(n, j0 + per);
    double xs = 0.0, cx = 0.0, mx = 0.0;
    for (int j = j0 + threadIdx.x; j < j1; j += blockDim.x) {
        double a = 0.0;
        for (int t = colptr[j]; t < colptr[j + 1]; ++t) a += val[t] * yL[rowidx[t]];
        const int k = blk[j];
        const double r = c[j] - a - (k >= 0 ? yK[k] : 0.0) - s[j];
        rd[j] = r;
        xs += x[j] * s[j];
        cx += c[j] * x[j];
        mx = fmax(mx, fabs(r));
    }
    xs = cta_sum(xs, sh);
    cx = cta_sum(cx, sh);
    mx = cta_max(mx, sh);
    if (threadIdx.x == 0) { part[blockIdx.x] = xs; part[gridDim.x + blockIdx.x] = cx; part[2 * gridDim.x + blockIdx.x] = mx; }
}
// rpK_k = bK_k - sum_{j in k} x_j
__global__ void k_resid_blocks(int K, const int *blkptr, const int *blkcols, const double *x, const double *bK, double *rpK)
{
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < K; k += gridDim.x * blockDim.x) {
        double sK = 0.0;
        for (int e = blkptr[k]; e < blkptr[k + 1]; ++e) sK += x[blkcols[e]];
        rpK[k] = bK[k] - sK;
    }
}
// one CTA: fold the column partials, the row / block residual norms and b'y into the scalar array
__global__ void k_fold_resid(int parts, const double *part, int L, int K, const double *rpL, const double *rpK, const double *bL,
                             const double *bK, const double *yL, const double *yK, double *scal)
{
    __shared__ double sh[32];
    double xs = 0.0, cx = 0.0, mx = 0.0;
    for (int i = threadIdx.x; i < parts; i += blockDim.x) { xs += part[i]; cx += part[parts + i]; mx = fmax(mx, part[2 * parts + i]); }
    xs = cta_sum(xs, sh); cx = cta_sum(cx, sh); mx = cta_max(mx, sh);
    double pl = 0.0, pk = 0.0, by = 0.0;
    for (int i = threadIdx.x; i < L; i += blockDim.x) { pl = fmax(pl, fabs(rpL[i])); by += bL[i] * yL[i]; }
    double byk = 0.0;
    for (int k = threadIdx.x; k < K; k += blockDim.x) { pk = fmax(pk, fabs(rpK[k])); byk += bK[k] * yK[k]; }
    pl = cta_max(pl, sh); pk = cta_max(pk, sh);
    by = cta_sum(by, sh); byk = cta_sum(byk, sh);
    if (threadIdx.x == 0) {
        scal[SC_XS] = xs; scal[SC_POBJ] = cx; scal[SC_DINF] = mx; scal[SC_PINF_L] = pl; scal[SC_PINF_K] = pk; scal[SC_DOBJ] = by + byk;
    }
}
// starting point helpers: min and sum of x and s, and x's (per-CTA partials: 5 x grid)
__global__ void k_start_stats(int n, const double *x, const double *s, double *part)
{
    __shared__ double sh[32];
    const int per = (n + gridDim.x - 1) / gridDim.x;
    const int j0 = blockIdx.x * per, j1 = min(n, j0 + per);
    double mnx = CUDART_INF, mns = CUDART_INF, sx = 0.0, ss = 0.0, xs = 0.0;
    for (int j = j0 + threadIdx.x; j < j1; j += blockDim.x) {
        mnx = fmin(mnx, x[j]); mns = fmin(mns, s[j]); sx += x[j]; ss += s[j]; xs += x[j] * s[j];
    }
    mnx = -cta_max(-mnx, sh); mns = -cta_max(-mns, sh);
    sx = cta_sum(sx, sh); ss = cta_sum(ss, sh); xs = cta_sum(xs, sh);
    if (threadIdx.x == 0) {
        part[blockIdx.x] = mnx; part[gridDim.x + blockIdx.x] = mns; part[2 * gridDim.x + blockIdx.x] = sx;
        part[3 * gridDim.x + blockIdx.x] = ss; part[4 * gridDim.x + blockIdx.x] = xs;
    }
}
__global__ void k_fold_start(int parts, const double *part, double *scal)
{
    __shared__ double sh[32];
    double mnx = CUDART_INF, mns = CUDART_INF, sx = 0.0, ss = 0.0, xs = 0.0;
    for (int i = threadIdx.x; i < parts; i += blockDim.x) {
        mnx = fmin(mnx, part[i]); mns = fmin(mns, part[parts + i]); sx += part[2 * parts + i]; ss += part[3 * parts + i]; xs += part[4 * parts + i];
    }
    mnx = -cta_max(-mnx, sh); mns = -cta_max(-mns, sh);
    sx = cta_sum(sx, sh); ss = cta_sum(ss, sh); xs = cta_sum(xs, sh);
    if (threadIdx.x == 0) { scal[SC_MINX] = mnx; scal[SC_MINS] = mns; scal[SC_SUMX] = sx; scal[SC_SUMS] = ss; scal[SC_XS2] = xs; }
}
__global__ void k_shift(int n, double *x, double ax, double *s, double as)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) { x[j] += ax; s[j] += as; }
}
// x_j = a_j'yL + yK[blk] (A'y) or s_j = c_j - that
__global__ void k_aty(int n, const int *colptr, const int *rowidx, const double *val, const int *blk, const double *yL, const double *yK,
                      const double *c, double *out)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        double a = 0.0;
        for (int t = colptr[j]; t < colptr[j + 1]; ++t) a += val[t] * yL[rowidx[t]];
        const int k = blk[j];
        a += k >= 0 ? yK[k] : 0.0;
        out[j] = c ? c[j] - a : a;
    }
}
__global__ void k_fill(int n, double *p, double v)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) p[j] = v;
}

template <typename T>
struct DBuf {
    T *p = nullptr;
    size_t cap = 0;
    cudaError_t need(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = n + n / 2 + 64;
        return cudaMalloc((void **)&p, sizeof(T) * cap);
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace

struct dwgpu_master {
    int device = 0, L = 0, K = 0;
    cudaStream_t st = nullptr;
    // host model (CSC over the linking rows; the convexity entry of a column is its block id)
    std::vector<int> colptr{0}, rowidx, blk;
    std::vector<double> val, cost, rhs;
    bool dirty = true;
    int n = 0;
    // device model
    DBuf<int> d_colptr, d_rowidx, d_blk, d_rowptr, d_colidx, d_blkptr, d_blkcols;
    DBuf<double> d_val, d_cost, d_rval, d_b;
    // iterates and work vectors
    DBuf<double> x, s, rd, dx, ds, dxa, dsa, h, rcs, v, w, d;
    DBuf<double> yL, dyL, rpL, tL, yK, dyK, rpK, u, t, delta, abar, Bt, S, Spart, part, scal, sigmu;
    DBuf<double> bx, bs, byL, byK;      // best iterate so far (returned when the tail of a solve loses accuracy)
    double q_pinf = 0.0, q_dinf = 0.0, q_gap = 0.0;
    double *h_scal = nullptr;
    // last solution (host)
    std::vector<double> sol_x, sol_y;
    double sol_obj = 0.0, sol_dobj = 0.0;
    long long iters = 0;
    int last_iters = 0;
    double last_seconds = 0.0;
};

extern "C" {

const char *dwgpu_master_last_error(void) { return g_merr.c_str(); }

int dwgpu_master_create(int device, int L, int K, dwgpu_master **out)
{
    if (!out || L < 1 || K < 1) return mfail(DWGPU_E_ARGS, "bad arguments");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return mfail(DWGPU_E_CUDA, "no such CUDA device");
    MCU(cudaSetDevice(device));
    dwgpu_master *m = new dwgpu_master();
    m->device = device; m->L = L; m->K = K;
    m->rhs.assign((size_t)L + K, 0.0);
    cudaError_t e = cudaStreamCreateWithFlags(&m->st, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMallocHost((void **)&m->h_scal, sizeof(double) * SC_COUNT);
    if (e != cudaSuccess) { delete m; return mfail(DWGPU_E_CUDA, cudaGetErrorString(e)); }
    // shared-memory needs of the one-CTA kernels
    const int pan_bytes = (int)(sizeof(double) * (size_t)L * CH_LD);
    if (pan_bytes > 48 * 1024) {
        e = cudaFuncSetAttribute(k_cholesky, cudaFuncAttributeMaxDynamicSharedMemorySize, pan_bytes);
        if (e != cudaSuccess) { cudaFreeHost(m->h_scal); cudaStreamDestroy(m->st); delete m; return mfail(DWGPU_E_ARGS, "too many linking rows for the device master"); }
    }
    *out = m;
    return 0;
}

void dwgpu_master_destroy(dwgpu_master *m)
{
    if (!m) return;
    cudaSetDevice(m->device);
    for (DBuf<int> *b : {&m->d_colptr, &m->d_rowidx, &m->d_blk, &m->d_rowptr, &m->d_colidx, &m->d_blkptr, &m->d_blkcols}) b->release();
    for (DBuf<double> *b : {&m->d_val, &m->d_cost, &m->d_rval, &m->d_b, &m->x, &m->s, &m->rd, &m->dx, &m->ds, &m->dxa, &m->dsa, &m->h, &m->rcs,
                            &m->v, &m->w, &m->d, &m->yL, &m->dyL, &m->rpL, &m->tL, &m->yK, &m->dyK, &m->rpK, &m->u, &m->t, &m->delta, &m->abar,
                            &m->Bt, &m->S, &m->Spart, &m->part, &m->scal, &m->sigmu, &m->bx, &m->bs, &m->byL, &m->byK}) b->release();
    if (m->h_scal) cudaFreeHost(m->h_scal);
    if (m->st) cudaStreamDestroy(m->st);
    delete m;
}

int dwgpu_master_set_rhs(dwgpu_master *m, const double *b)
{
    if (!m || !b) return mfail(DWGPU_E_ARGS, "bad arguments");
    std::copy(b, b + m->L + m->K, m->rhs.begin());
    m->dirty = true;
    return 0;
}

// columns in CSC form over ALL rows (row < L: linking coefficient; row >= L: the column's convexity row, coefficient 1)
int dwgpu_master_add_cols(dwgpu_master *m, int ncols, const int *colptr, const int *rows, const double *vals, const double *cost)
{
    if (!m || ncols < 0 || (ncols > 0 && (!colptr || !cost))) return mfail(DWGPU_E_ARGS, "bad arguments");
    for (int j = 0; j < ncols; ++j) {
        int k = -1;
        for (int t = colptr[j]; t < colptr[j + 1]; ++t) {
            const int r = rows[t];
            if (r < 0 || r >= m->L + m->K) return mfail(DWGPU_E_ARGS, "row index out of range");
            if (r >= m->L) { k = r - m->L; continue; }
            if (vals[t] == 0.0) continue;
            m->rowidx.push_back(r); m->val.push_back(vals[t]);
        }
        m->colptr.push_back((int)m->rowidx.size());
        m->blk.push_back(k);
        m->cost.push_back(cost[j]);
    }
    m->n = (int)m->cost.size();
    m->dirty = true;
    return 0;
}

int dwgpu_master_set_costs(dwgpu_master *m, int first, int count, const double *cost)
{
    if (!m || first < 0 || count < 0 || first + count > m->n || (count > 0 && !cost)) return mfail(DWGPU_E_ARGS, "bad arguments");
    std::copy(cost, cost + count, m->cost.begin() + first);
    m->dirty = true;
    return 0;
}

int dwgpu_master_set_coef(dwgpu_master *m, int col, int pos, double v)
{
    if (!m || col < 0 || col >= m->n || pos < 0 || m->colptr[col] + pos >= m->colptr[col + 1]) return mfail(DWGPU_E_ARGS, "bad arguments");
    m->val[(size_t)m->colptr[col] + pos] = v;
    m->dirty = true;
    return 0;
}

int dwgpu_master_del_cols(dwgpu_master *m, const char *flag)
{
    if (!m || (m->n > 0 && !flag)) return mfail(DWGPU_E_ARGS, "bad arguments");
    std::vector<int> cp{0}, ri, bk;
    std::vector<double> vl, cs;
    for (int j = 0; j < m->n; ++j) {
        if (flag[j]) continue;
        for (int t = m->colptr[j]; t < m->colptr[j + 1]; ++t) { ri.push_back(m->rowidx[t]); vl.push_back(m->val[t]); }
        cp.push_back((int)ri.size());
        bk.push_back(m->blk[j]);
        cs.push_back(m->cost[j]);
    }
    m->colptr.swap(cp); m->rowidx.swap(ri); m->blk.swap(bk); m->val.swap(vl); m->cost.swap(cs);
    m->n = (int)m->cost.size();
    m->dirty = true;
    return 0;
}

int dwgpu_master_cols(const dwgpu_master *m) { return m ? m->n : 0; }

static int upload_model(dwgpu_master *m)
{
    const int n = m->n, L = m->L, K = m->K;
    const size_t nnz = m->rowidx.size();
    // CSR of A_L and the per-block column lists (columns in index order: every sum has a fixed order)
    std::vector<int> rowptr((size_t)L + 1, 0), colidx(nnz), blkptr((size_t)K + 1, 0), blkcols;
    std::vector<double> rval(nnz);
    for (size_t t = 0; t < nnz; ++t) rowptr[(size_t)m->rowidx[t] + 1]++;
    for (int i = 0; i < L; ++i) rowptr[(size_t)i + 1] += rowptr[i];
    {
        std::vector<int> pos(rowptr.begin(), rowptr.end() - 1);
        for (int j = 0; j < n; ++j)
            for (int t = m->colptr[j]; t < m->colptr[j + 1]; ++t) { const int p = pos[m->rowidx[t]]++; colidx[p] = j; rval[p] = m->val[t]; }
    }
    for (int j = 0; j < n; ++j) if (m->blk[j] >= 0) blkptr[(size_t)m->blk[j] + 1]++;
    for (int k = 0; k < K; ++k) {
        if (blkptr[(size_t)k + 1] == 0) return mfail(DWGPU_E_STATE, "a block has no master column");
        blkptr[(size_t)k + 1] += blkptr[k];
    }
    blkcols.resize((size_t)blkptr[K]);
    {
        std::vector<int> pos(blkptr.begin(), blkptr.end() - 1);
        for (int j = 0; j < n; ++j) if (m->blk[j] >= 0) blkcols[(size_t)pos[m->blk[j]]++] = j;
    }
    auto up_i = [&](DBuf<int> &b, const std::vector<int> &h) -> cudaError_t {
        cudaError_t e = b.need(std::max<size_t>(h.size(), 1));
        if (e != cudaSuccess || h.empty()) return e;
        return cudaMemcpyAsync(b.p, h.data(), sizeof(int) * h.size(), cudaMemcpyHostToDevice, m->st);
    };
    auto up_d = [&](DBuf<double> &b, const std::vector<double> &h) -> cudaError_t {
        cudaError_t e = b.need(std::max<size_t>(h.size(), 1));
        if (e != cudaSuccess || h.empty()) return e;
        return cudaMemcpyAsync(b.p, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice, m->st);
    };
    MCU(up_i(m->d_colptr, m->colptr)); MCU(up_i(m->d_rowidx, m->rowidx)); MCU(up_i(m->d_blk, m->blk));
    MCU(up_i(m->d_rowptr, rowptr)); MCU(up_i(m->d_colidx, colidx)); MCU(up_i(m->d_blkptr, blkptr)); MCU(up_i(m->d_blkcols, blkcols));
    MCU(up_d(m->d_val, m->val)); MCU(up_d(m->d_cost, m->cost)); MCU(up_d(m->d_rval, rval)); MCU(up_d(m->d_b, m->rhs));
    for (DBuf<double> *b : {&m->x, &m->s, &m->rd, &m->dx, &m->ds, &m->dxa, &m->dsa, &m->h, &m->rcs, &m->v, &m->w, &m->d, &m->bx, &m->bs}) MCU(b->need((size_t)n));
    MCU(m->byL.need((size_t)L)); MCU(m->byK.need((size_t)K));
    for (DBuf<double> *b : {&m->yL, &m->dyL, &m->rpL, &m->tL}) MCU(b->need((size_t)L));
    for (DBuf<double> *b : {&m->yK, &m->dyK, &m->rpK, &m->u, &m->t, &m->delta}) MCU(b->need((size_t)K));
    MCU(m->abar.need((size_t)K * L)); MCU(m->Bt.need((size_t)n * L)); MCU(m->S.need((size_t)L * L));
    MCU(m->part.need(8 * (size_t)RED_CTAS)); MCU(m->scal.need(SC_COUNT)); MCU(m->sigmu.need(1));
    MCU(cudaStreamSynchronize(m->st));       // the host vectors above go out of scope
    m->dirty = false;
    return 0;
}

// tol: relative primal / dual infeasibility and duality gap.  status: 0 optimal, 3 not converged.
int dwgpu_master_solve(dwgpu_master *m, double tol, int max_iter, int *status_out)
{
    if (!m || !status_out) return mfail(DWGPU_E_ARGS, "bad arguments");
    MCU(cudaSetDevice(m->device));
    if (m->n < 1) return mfail(DWGPU_E_STATE, "the master has no columns");
    cudaEvent_t ev0, ev1;
    MCU(cudaEventCreate(&ev0)); MCU(cudaEventCreate(&ev1));
    if (m->dirty) { const int rc = upload_model(m); if (rc) return rc; }
    const int n = m->n, L = m->L, K = m->K;
    cudaStream_t st = m->st;
    MCU(cudaEventRecord(ev0, st));
    const int *colptr = m->d_colptr.p, *rowidx = m->d_rowidx.p, *blk = m->d_blk.p, *rowptr = m->d_rowptr.p, *colidx = m->d_colidx.p,
              *blkptr = m->d_blkptr.p, *blkcols = m->d_blkcols.p;
    const double *val = m->d_val.p, *c = m->d_cost.p, *rval = m->d_rval.p, *bL = m->d_b.p, *bK = m->d_b.p + L;
    const int GV = std::min(1024, (n + 255) / 256), GK = std::min(1024, (K + 255) / 256);
    const int nb = (L + SY_T - 1) / SY_T, npairs = nb * (nb + 1) / 2;
    const int chunk = std::max(SY_KS, ((n + 147) / 148 + SY_KS - 1) / SY_KS * SY_KS);       // ~one wave of CTAs per tile pair row
    const int nchunks = (n + chunk - 1) / chunk;
    MCU(m->Spart.need((size_t)nchunks * npairs * SY_T * SY_T));
    const size_t pan_bytes = sizeof(double) * (size_t)L * CH_LD, y_bytes = sizeof(double) * (size_t)L;
    double *scal = m->scal.p, *part = m->part.p;

    // per-phase device times (DWGPU_MASTER_DEBUG=2): events around the pieces of one factorisation / solve
    const bool ptime = std::getenv("DWGPU_MASTER_DEBUG") && std::atoi(std::getenv("DWGPU_MASTER_DEBUG")) >= 2;
    cudaEvent_t pe[8];
    float pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (ptime) for (auto &e_ : pe) cudaEventCreate(&e_);
    auto mark = [&](int i) { if (ptime) cudaEventRecord(pe[i], st); };
    auto collect = [&](int n_marks) {
        if (!ptime) return;
        cudaEventSynchronize(pe[n_marks - 1]);
        for (int i = 0; i + 1 < n_marks; ++i) { float ms_ = 0.f; cudaEventElapsedTime(&ms_, pe[i], pe[i + 1]); pacc[i] += ms_; }
    };
    auto factor = [&]() -> cudaError_t {     // d -> delta, abar, Bt, S = Bt'Bt, Cholesky
        mark(0);
        k_block_prep<<<K, 128, sizeof(double) * L, st>>>(L, blkptr, blkcols, colptr, rowidx, val, m->d.p, m->delta.p, m->abar.p);
        k_build_bt<<<n, 64, 0, st>>>(L, colptr, rowidx, val, blk, m->d.p, m->abar.p, m->Bt.p);
        mark(1);
        k_syrk_dmma<<<dim3(npairs, nchunks), 256, 0, st>>>(L, n, chunk, m->Bt.p, m->Spart.p);
        mark(2);
        k_syrk_reduce<<<dim3(npairs, 16), 256, 0, st>>>(L, npairs, nchunks, m->Spart.p, m->S.p);
        mark(3);
        k_cholesky<<<1, 1024, pan_bytes, st>>>(L, m->S.p, 1e-14);
        mark(4);
        collect(5);
        return cudaGetLastError();
    };
    // solves (A D A') [dyL; dyK] = [rL + A_L h ; rK + E h] given h (and rcs for the primal step when wanted)
    auto normal_solve = [&](const double *rL, const double *rK, const double *hvec, double *dyL, double *dyK) -> cudaError_t {
        k_block_u<<<GK, 256, 0, st>>>(K, blkptr, blkcols, hvec, rK, m->delta.p, m->u.p, m->t.p);
        k_w<<<GV, 256, 0, st>>>(n, blk, hvec, m->d.p, m->t.p, m->w.p);
        k_row_reduce<<<L, 256, 0, st>>>(rowptr, colidx, rval, m->w.p, rL, 1.0, m->tL.p);
        k_chol_solve<<<1, 1024, y_bytes, st>>>(L, m->S.p, m->tL.p, dyL);
        k_col_dot<<<GV, 256, 0, st>>>(n, colptr, rowidx, val, dyL, m->v.p);
        k_block_dy<<<GK, 256, 0, st>>>(K, blkptr, blkcols, m->d.p, m->v.p, m->u.p, m->delta.p, dyK);
        return cudaGetLastError();
    };
    auto read_scal = [&]() -> cudaError_t {
        cudaError_t e = cudaMemcpyAsync(m->h_scal, scal, sizeof(double) * SC_COUNT, cudaMemcpyDeviceToHost, st);
        return e != cudaSuccess ? e : cudaStreamSynchronize(st);
    };

    // ---- Mehrotra's starting point: x = A'(AA')^-1 b, y = (AA')^-1 A c, s = c - A'y, shifted into the positive orthant
    k_fill<<<GV, 256, 0, st>>>(n, m->d.p, 1.0);
    MCU(factor());
    k_fill<<<GV, 256, 0, st>>>(n, m->h.p, 0.0);
    MCU(normal_solve(bL, bK, m->h.p, m->yL.p, m->yK.p));
    k_aty<<<GV, 256, 0, st>>>(n, colptr, rowidx, val, blk, m->yL.p, m->yK.p, nullptr, m->x.p);
    // y = (AA')^-1 A c: with d = 1 and h = c, rL = rK = 0 the same solve returns it
    k_fill<<<1, 256, 0, st>>>(L, m->rpL.p, 0.0);
    k_fill<<<GK, 256, 0, st>>>(K, m->rpK.p, 0.0);
    MCU(normal_solve(m->rpL.p, m->rpK.p, c, m->yL.p, m->yK.p));
    k_aty<<<GV, 256, 0, st>>>(n, colptr, rowidx, val, blk, m->yL.p, m->yK.p, c, m->s.p);
    k_start_stats<<<RED_CTAS, RED_NT, 0, st>>>(n, m->x.p, m->s.p, part);
    k_fold_start<<<1, 256, 0, st>>>(RED_CTAS, part, scal);
    MCU(read_scal());
    {
        const double dx0 = std::max(-1.5 * m->h_scal[SC_MINX], 0.0), ds0 = std::max(-1.5 * m->h_scal[SC_MINS], 0.0);
        // sums after the first shift, analytically
        double xs = m->h_scal[SC_XS2] + dx0 * m->h_scal[SC_SUMS] + ds0 * m->h_scal[SC_SUMX] + (double)n * dx0 * ds0;
        double sx = m->h_scal[SC_SUMX] + n * dx0, ss = m->h_scal[SC_SUMS] + n * ds0;
        double ax = dx0, as = ds0;
        if (!(xs > 0.0) || !(sx > 0.0) || !(ss > 0.0)) { ax += 1.0; as += 1.0; xs = xs + sx + ss + n; sx += n; ss += n; }
        ax += 0.5 * xs / ss;
        as += 0.5 * xs / sx;
        k_shift<<<GV, 256, 0, st>>>(n, m->x.p, ax, m->s.p, as);
    }
    double nb_ = 1.0, nc_ = 1.0;
    for (double v : m->rhs) nb_ = std::max(nb_, 1.0 + std::fabs(v));
    for (double v : m->cost) nc_ = std::max(nc_, 1.0 + std::fabs(v));

    int status = 3, it = 0, worse = 0;
    double best_merit = INFINITY, best_pobj = 0.0, best_dobj = 0.0;
    double best_q[3] = {INFINITY, INFINITY, INFINITY};
    auto keep_best = [&]() -> cudaError_t {
        cudaError_t e = cudaMemcpyAsync(m->bx.p, m->x.p, sizeof(double) * n, cudaMemcpyDeviceToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(m->bs.p, m->s.p, sizeof(double) * n, cudaMemcpyDeviceToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(m->byL.p, m->yL.p, sizeof(double) * L, cudaMemcpyDeviceToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(m->byK.p, m->yK.p, sizeof(double) * K, cudaMemcpyDeviceToDevice, st);
        return e;
    };
    for (it = 1; it <= max_iter; ++it) {
        // residuals and the scalars of this iterate
        k_row_reduce<<<L, 256, 0, st>>>(rowptr, colidx, rval, m->x.p, bL, -1.0, m->rpL.p);
        k_resid_blocks<<<GK, 256, 0, st>>>(K, blkptr, blkcols, m->x.p, bK, m->rpK.p);
        k_resid_cols<<<RED_CTAS, RED_NT, 0, st>>>(n, colptr, rowidx, val, blk, c, m->yL.p, m->yK.p, m->x.p, m->s.p, m->rd.p, part);
        k_fold_resid<<<1, 256, 0, st>>>(RED_CTAS, part, L, K, m->rpL.p, m->rpK.p, bL, bK, m->yL.p, m->yK.p, scal);
        MCU(read_scal());
        const double pobj = m->h_scal[SC_POBJ], dobj = m->h_scal[SC_DOBJ];
        const double pinf = std::max(m->h_scal[SC_PINF_L], m->h_scal[SC_PINF_K]) / nb_, dinf = m->h_scal[SC_DINF] / nc_;
        const double gap = std::fabs(pobj - dobj) / (1.0 + std::fabs(pobj));
        if (std::getenv("DWGPU_MASTER_DEBUG"))
            std::fprintf(stderr, "    ipm %3d pobj %.10e dobj %.10e pinf %.1e dinf %.1e gap %.1e mu %.1e\n", it, pobj, dobj, pinf, dinf, gap,
                         m->h_scal[SC_XS] / n);
        const double merit = std::max(pinf, std::max(dinf, gap));
        if (!std::isfinite(merit) || !std::isfinite(pobj) || !std::isfinite(dobj) || !std::isfinite(m->h_scal[SC_XS])) break;   // the last step lost it: the best iterate is returned
        if (merit < best_merit) {
            best_merit = merit; best_pobj = pobj; best_dobj = dobj; best_q[0] = pinf; best_q[1] = dinf; best_q[2] = gap;
            MCU(keep_best());
            worse = 0;
        } else if (++worse >= 4 && best_merit <= 1e3 * tol) break;       // past its best accuracy (ill-conditioned tail)
        if (merit <= tol) { status = 0; break; }
        const double mu = m->h_scal[SC_XS] / n;
        if (!(mu > 0.0)) break;
        k_dvec<<<GV, 256, 0, st>>>(n, m->x.p, m->s.p, m->d.p);
        MCU(factor());
        // predictor
        k_h<<<GV, 256, 0, st>>>(n, 0, nullptr, m->x.p, m->s.p, m->d.p, m->rd.p, nullptr, nullptr, nullptr, m->h.p, m->rcs.p);
        MCU(normal_solve(m->rpL.p, m->rpK.p, m->h.p, m->dyL.p, m->dyK.p));
        k_dxds<<<RED_CTAS, RED_NT, 0, st>>>(n, blk, m->rd.p, m->v.p, m->dyK.p, m->rcs.p, m->d.p, m->x.p, m->s.p, m->dxa.p, m->dsa.p, part);
        k_fold_min2<<<1, 256, 0, st>>>(RED_CTAS, part, scal, 1.0);
        k_muaff<<<RED_CTAS, RED_NT, 0, st>>>(n, m->x.p, m->s.p, m->dxa.p, m->dsa.p, scal, part);
        k_fold_sum<<<1, 256, 0, st>>>(RED_CTAS, part, scal + SC_MUAFF);
        k_sigma<<<1, 1, 0, st>>>(n, scal, m->sigmu.p);
        // corrector
        k_h<<<GV, 256, 0, st>>>(n, 1, m->sigmu.p, m->x.p, m->s.p, m->d.p, m->rd.p, m->dxa.p, m->dsa.p, nullptr, m->h.p, m->rcs.p);
        MCU(normal_solve(m->rpL.p, m->rpK.p, m->h.p, m->dyL.p, m->dyK.p));
        k_dxds<<<RED_CTAS, RED_NT, 0, st>>>(n, blk, m->rd.p, m->v.p, m->dyK.p, m->rcs.p, m->d.p, m->x.p, m->s.p, m->dx.p, m->ds.p, part);
        const double eta = std::max(0.995, 1.0 - mu);
        k_fold_min2<<<1, 256, 0, st>>>(RED_CTAS, part, scal, eta);
        k_update<<<std::min(1024, (n + L + K + 255) / 256), 256, 0, st>>>(n, L, K, scal, m->x.p, m->s.p, m->dx.p, m->ds.p, m->yL.p, m->dyL.p,
                                                                          m->yK.p, m->dyK.p);
        MCU(cudaGetLastError());
    }
    if (ptime) {
        std::fprintf(stderr, "    ipm phases (ms over %d factorisations): prep+build %.2f, syrk %.2f, reduce %.2f, cholesky %.2f\n", it, pacc[0], pacc[1], pacc[2], pacc[3]);
        for (auto &e_ : pe) cudaEventDestroy(e_);
    }
    if (!std::isfinite(best_merit)) return mfail(DWGPU_E_STATE, "the interior-point master produced no finite iterate");
    if (status != 0 && best_merit <= 1e2 * tol) status = 0;      // converged as far as FP64 normal equations go: accepted,
                                                                 // the achieved quality is reported by dwgpu_master_quality
    m->q_pinf = best_q[0]; m->q_dinf = best_q[1]; m->q_gap = best_q[2];
    if (it > max_iter) it = max_iter;
    // results
    m->sol_x.resize((size_t)n); m->sol_y.resize((size_t)L + K);
    MCU(cudaMemcpyAsync(m->sol_x.data(), m->bx.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    MCU(cudaMemcpyAsync(m->sol_y.data(), m->byL.p, sizeof(double) * L, cudaMemcpyDeviceToHost, st));
    MCU(cudaMemcpyAsync(m->sol_y.data() + L, m->byK.p, sizeof(double) * K, cudaMemcpyDeviceToHost, st));
    MCU(cudaEventRecord(ev1, st));
    MCU(cudaStreamSynchronize(st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev0, ev1);
    cudaEventDestroy(ev0); cudaEventDestroy(ev1);
    m->last_seconds = ms * 1e-3;
    m->sol_obj = best_pobj; m->sol_dobj = best_dobj;
    m->last_iters = it;
    m->iters += it;
    *status_out = status;
    return 0;
}

int dwgpu_master_get(dwgpu_master *m, double *obj, double *dual_obj, double *y, double *x, int *iterations, double *seconds)
{
    if (!m || m->sol_x.size() != (size_t)m->n) return mfail(DWGPU_E_STATE, "no solution");
    if (obj) *obj = m->sol_obj;
    if (dual_obj) *dual_obj = m->sol_dobj;
    if (y) std::copy(m->sol_y.begin(), m->sol_y.end(), y);
    if (x) std::copy(m->sol_x.begin(), m->sol_x.end(), x);
    if (iterations) *iterations = m->last_iters;
    if (seconds) *seconds = m->last_seconds;
    return 0;
}

// what the last solve achieved: relative primal infeasibility, dual infeasibility and duality gap of the returned iterate
int dwgpu_master_quality(dwgpu_master *m, double *pinf, double *dinf, double *gap)
{
    if (!m) return mfail(DWGPU_E_ARGS, "bad arguments");
    if (pinf) *pinf = m->q_pinf;
    if (dinf) *dinf = m->q_dinf;
    if (gap) *gap = m->q_gap;
    return 0;
}

}  // extern "C"
