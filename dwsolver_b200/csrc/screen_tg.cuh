// screen_tg.cuh — first launch of a priced round on path 3 (tableau in HBM): ONE WARP per block decides whether the block
// needs a pivot at all under the new duals.
//
// Most rounds of a Dantzig-Wolfe run move few blocks (config B: 29 of the 40 rounds of the recorded trajectory move
// fewer than 100 of 8192).  For a block that stays optimal the CTA kernel (simplex_smem.cuh) still pays its whole
// prologue and epilogue — a chain of ~11 dependent HBM/L2 round trips with 4 blocks resident per SM, 0.17-0.19 ms per
// launch.  Here a block is one warp with ~10 KB of shared memory, 20 blocks resident per SM: it prices the columns
// (y = D_k' pi, same order and rounding as src/dw_blas.c:106-122 / the CTA kernel), updates the saved reduced costs
// incrementally (d = d_saved + dc_N + T' dc_B over the masked items of the rows whose basic cost moved — the identical
// fma sequence per item as weighted_itemsum_tg), and prices.  No attractive column: the block is finished here exactly
// like the CTA kernel's "nothing moved" exit (obj, status, saved costs).  Otherwise nothing is written and the block id
// goes to the list of the solve launch together with the number of attractive columns, which orders that launch
// (longest expected work first).  Reference path replaced: the warm-started glp_simplex call of a block that turns out
// to need zero iterations (src/dw_subprob.c:291-340).
#pragma once
#include "simplex_smem.cuh"

namespace dwk {

struct ScreenOut {
    int *next_list;       // blocks left for the solve launch (arrival order; sorted by order_screened_kernel)
    int *next_count;
    unsigned *key;        // by block id: attractive columns found (0xffffffff: not screened — no valid saved row)
};

__host__ __device__ inline size_t screen_warp_bytes(int m, int n)
{
    return align16(sizeof(double) * (size_t)n)                       // dl: cost changes by structural column
         + align16(sizeof(unsigned) * (size_t)m * rmask_words(n))    // row-item masks
         + align16(sizeof(double) * (size_t)m)                       // weights by list position
         + align16(sizeof(unsigned short) * (size_t)(m + 2));        // listed rows
}

// VNT: columns priced per batch of loads (the CTA kernel's thread count).  Handles n <= 512 (<= 4 items per lane).
template <int NW, int VNT>
__global__ void __launch_bounds__(NW * 32) screen_tg_kernel(const BlockDev *blocks, const int *list, int count, SolveParams P,
                                                            ScreenOut S, int warp_bytes)
{
    extern __shared__ __align__(16) char smem_raw[];
    constexpr int VW = VNT / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int e = blockIdx.x * NW + warp;
    if (e >= count) return;
    const int k = list[e];
    const BlockDev &B = blocks[k];
    const int m = B.m, n = B.n, ld = tab_ld(n, true), rwd = B.rwd;
    const int nitems = ld >> 2;
    char *sp = smem_raw + (size_t)warp * warp_bytes;
    double *dl = (double *)sp;                 sp += align16(sizeof(double) * (size_t)n);
    unsigned *smask = (unsigned *)sp;          sp += align16(sizeof(unsigned) * (size_t)m * rwd);
    double *wts = (double *)sp;                sp += align16(sizeof(double) * (size_t)m);
    unsigned short *rows = (unsigned short *)sp;
    const unsigned lt_mask = (1u << lane) - 1u;
    const double tol_dj = P.tol_dj, tol_bnd = P.tol_bnd;

    const int dflag_in = B.dflag[0];
    if (!(dflag_in > 0 && dflag_in < 32)) {    // no valid saved row, or the 32nd solve (full recomputation): the CTA kernel's job
        if (lane == 0) { S.key[k] = 0xffffffffu; S.next_list[atomicAdd(S.next_count, 1)] = k; }
        return;
    }
    // row masks -> shared (one coalesced pass; read again only through shared memory)
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(B.rmask);
        uint4 *dst = reinterpret_cast<uint4 *>(smask);
        const int nv = (m * rwd) >> 2, tail = (m * rwd) & 3;
        for (int t = lane; t < nv; t += 32) dst[t] = __ldcg(src + t);
        for (int t = lane; t < tail; t += 32) smask[4 * nv + t] = __ldcg(B.rmask + 4 * nv + t);
    }
    // ---- feasibility of the kept basis (the CTA kernel's scan; a violated bound means phase 1)
    const int *head = B.head, *nbv = B.nbv;
    const signed char *vstat = B.vstat;
    {
        bool viol = false;
        constexpr int U = 4;
        for (int i0 = lane; i0 < m; i0 += 32 * U) {
            double b[U], lo[U], up[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + 32 * u;
                b[u] = 0.0; lo[u] = -CUDART_INF; up[u] = CUDART_INF;
                if (i < m) { const int v = __ldcg(head + i); b[u] = __ldcg(B.beta + i); lo[u] = __ldg(B.lo + v); up[u] = __ldg(B.up + v); }
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
                viol |= (b[u] < lo[u] - tol_bnd * (1.0 + fabs(lo[u]))) || (b[u] > up[u] + tol_bnd * (1.0 + fabs(up[u])));
        }
        if (__any_sync(0xffffffffu, viol)) {
            if (lane == 0) { S.key[k] = 0xffffffffu; S.next_list[atomicAdd(S.next_count, 1)] = k; }
            return;
        }
    }
    // ---- priced costs cs = (phase_one ? 0 : c_master) - D_k' pi  -> B.ws (as the CTA kernel), cost changes -> dl,
    {
        const int *dcp = B.dc_ptr, *dcr = B.dc_row;
        const double *dcv = B.dc_val, *cm = B.c_master, *cssave = B.cssave;
        double *csw = B.ws;
        for (int base = 0; base < n; base += VNT) {
            int pb[VW], pe[VW], r0[VW];
            double v0[VW], p0[VW], cb[VW], sv[VW];
#pragma unroll
            for (int w = 0; w < VW; ++w) {
                const int j = base + 32 * w + lane;
                pb[w] = pe[w] = 0;
                if (j < n) { pb[w] = __ldg(dcp + j); pe[w] = __ldg(dcp + j + 1); }
            }
#pragma unroll
            for (int w = 0; w < VW; ++w) {
                const int j = base + 32 * w + lane;
                cb[w] = (j < n && !P.phase_one) ? __ldg(cm + j) : 0.0;
                sv[w] = 0.0;
                if (j < n) sv[w] = __ldcg(cssave + j);
                r0[w] = 0; v0[w] = 0.0;
                if (pb[w] < pe[w]) { r0[w] = __ldg(dcr + pb[w]); v0[w] = __ldg(dcv + pb[w]); }
            }
#pragma unroll
            for (int w = 0; w < VW; ++w) p0[w] = pb[w] < pe[w] ? __ldg(P.pi + r0[w]) : 0.0;
#pragma unroll
            for (int w = 0; w < VW; ++w) {
                const int j = base + 32 * w + lane;
                if (j >= n) continue;
                double y = 0.0;
                if (pb[w] < pe[w]) {
                    y = __dadd_rn(y, __dmul_rn(p0[w], v0[w]));
                    for (int t = pb[w] + 1; t < pe[w]; ++t)
                        y = __dadd_rn(y, __dmul_rn(__ldg(P.pi + __ldg(dcr + t)), __ldg(dcv + t)));
                }
                const double c = __dadd_rn(__dmul_rn(-1.0, y), cb[w]);
                csw[j] = c;
                dl[j] = cost_delta(c, sv[w]);
            }
        }
    }
    __syncwarp();
    // ---- rows whose basic cost moved, ascending, with their weights
    int nl = 0;
    unsigned cells = 0;
    for (int i0 = 0; i0 < m; i0 += 32) {
        const int i = i0 + lane;
        double w = 0.0;
        if (i < m) { const int v = __ldcg(head + i); if (v >= m) w = dl[v - m]; }
        const bool nz = w != 0.0;
        const unsigned bal = __ballot_sync(0xffffffffu, nz);
        if (nz) {
            const int pos = nl + __popc(bal & lt_mask);
            rows[pos] = (unsigned short)i; wts[pos] = w;
            for (int w_ = 0; w_ < rwd; ++w_) cells += __popc(smask[(size_t)i * rwd + w_]);
        }
        nl += __popc(bal);
    }
    __syncwarp();
    // ---- reduced costs of this lane's items: base row, then the masked row contributions in list order
    double2 a0[4], a1[4];
    int4 nb4[4];
    bool attractive = false;
    unsigned n_attr = 0;
    const double2 *T2 = reinterpret_cast<const double2 *>(B.T);
    const int n2 = ld >> 1;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int t = lane + 32 * u;
        a0[u] = make_double2(0.0, 0.0); a1[u] = a0[u]; nb4[u] = make_int4(0, 0, 0, 0);
        if (t < nitems) {
            const int j = 4 * t;
            double dsv[4];
            int vv[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                vv[x] = -1; dsv[x] = 0.0;
                if (j + x < n) { vv[x] = __ldcg(nbv + j + x); dsv[x] = __ldcg(B.dsave + j + x); }
            }
            double s[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                s[x] = 0.0;
                if (j + x < n) { if (vv[x] >= m) s[x] = dl[vv[x] - m]; s[x] += dsv[x]; }
            }
            a0[u] = make_double2(s[0], s[1]); a1[u] = make_double2(s[2], s[3]);
            nb4[u] = make_int4(vv[0], vv[1], vv[2], vv[3]);
        }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int t = lane + 32 * u;
        if (t >= nitems) continue;
        const unsigned *mcol = smask + (t >> 5);
        const int sh = t & 31;
        for (int b0 = 0; b0 < nl; b0 += 32) {
            unsigned have = 0;
            const int lim = min(32, nl - b0);
            for (int r = 0; r < lim; ++r) have |= ((mcol[(size_t)rows[b0 + r] * rwd] >> sh) & 1u) << r;
            while (have) {
                double2 v0[4], v1[4];
                double wv[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    wv[r] = 0.0; v0[r] = make_double2(0.0, 0.0); v1[r] = v0[r];
                    if (have) {
                        const int a = __ffs(have) - 1;
                        have &= have - 1;
                        const int i = rows[b0 + a];
                        wv[r] = wts[b0 + a];
                        v0[r] = __ldcg(T2 + (size_t)i * n2 + 2 * t);
                        v1[r] = __ldcg(T2 + (size_t)i * n2 + 2 * t + 1);
                    }
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    a0[u].x = fma(wv[r], v0[r].x, a0[u].x); a0[u].y = fma(wv[r], v0[r].y, a0[u].y);
                    a1[u].x = fma(wv[r], v1[r].x, a1[u].x); a1[u].y = fma(wv[r], v1[r].y, a1[u].y);
                }
            }
        }
        // pricing over the item's four slots
        const double dj[4] = {a0[u].x, a0[u].y, a1[u].x, a1[u].y};
        const int vv[4] = {nb4[u].x, nb4[u].y, nb4[u].z, nb4[u].w};
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            if (4 * t + x < n) {
                const signed char st = vstat[vv[x]];
                const bool ok = (st == ST_LO && dj[x] < -tol_dj) || (st == ST_UP && dj[x] > tol_dj) ||
                                (st == ST_FREE && fabs(dj[x]) > tol_dj);
                if (ok) { attractive = true; ++n_attr; }
            }
        }
    }
    if (__any_sync(0xffffffffu, attractive)) {
#pragma unroll
        for (int o = 16; o; o >>= 1) n_attr += __shfl_xor_sync(0xffffffffu, n_attr, o);
        if (lane == 0) { S.key[k] = n_attr; S.next_list[atomicAdd(S.next_count, 1)] = k; }
        return;
    }
    // ---- the block stays where it is: what the CTA kernel's "nothing moved" exit writes
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int t = lane + 32 * u;
        if (t >= nitems) continue;
        const int j = 4 * t;
        if (j + 3 < n) {
            double2 *ds2 = reinterpret_cast<double2 *>(B.dsave + j);
            ds2[0] = a0[u]; ds2[1] = a1[u];
        } else {
            const double dj[4] = {a0[u].x, a0[u].y, a1[u].x, a1[u].y};
            for (int x = 0; x < 4; ++x) if (j + x < n) B.dsave[j + x] = dj[x];
        }
    }
    __syncwarp();
    for (int j = lane; j < n; j += 32)
        if (dl[j] != 0.0) B.cssave[j] = B.ws[j];          // (written by this lane above)
    const double zo = standing_optimum(P, k, B.L, lane);
#pragma unroll
    for (int o = 16; o; o >>= 1) cells += __shfl_xor_sync(0xffffffffu, cells, o);
    if (lane == 0) {
        P.obj[k] = zo; P.status[k] = 5; P.last_iters[k] = 0;
        if (P.m_obj != nullptr) { P.m_obj[k] = zo; P.m_status[k] = 5; P.counters[8 * (size_t)k + 7] += 12; }
        P.counters[8 * (size_t)k + 6] += (unsigned long long)cells * 4ull;
        B.dflag[0] = dflag_in + 1;
    }
}

// Sorts the first *dyn_count entries of `list` by key (descending; ties: lower block id first), one CTA.  The list was
// appended to with atomics, so this also makes its order — and with it the launch order of the solve kernel — the same
// in every run.
template <int NT>
__global__ void __launch_bounds__(NT) order_screened_kernel(int *list, const int *dyn_count, const unsigned *key_by_block, int cap)
{
    extern __shared__ unsigned long long skey[];
    const int count = min(*dyn_count, cap);
    if (count <= 1) return;
    int P2 = 2;
    while (P2 < count) P2 <<= 1;
    for (int e = threadIdx.x; e < P2; e += NT) {
        unsigned long long kv = 0ull;                       // padding sorts last
        if (e < count) {
            const int k = list[e];
            unsigned long long it = key_by_block[k];
            kv = ((it + 1ull) << 32) | (unsigned long long)(0xffffffffu - (unsigned)k);
        }
        skey[e] = kv;
    }
    __syncthreads();
    for (int size = 2; size <= P2; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int e = threadIdx.x; e < P2 / 2; e += NT) {
                const int lo = 2 * e - (e & (stride - 1));
                const int hi = lo + stride;
                const bool desc = (lo & size) == 0;
                const unsigned long long a = skey[lo], b = skey[hi];
                if ((a < b) == desc) { skey[lo] = b; skey[hi] = a; }
            }
            __syncthreads();
        }
    for (int e = threadIdx.x; e < count; e += NT) list[e] = (int)(0xffffffffu - (unsigned)(skey[e] & 0xffffffffull));
}

}  // namespace dwk
