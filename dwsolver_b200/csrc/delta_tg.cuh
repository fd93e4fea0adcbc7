// delta_tg.cuh — rounds in which only a few master duals moved (path 3, tableau in HBM).
//
// Late Dantzig-Wolfe rounds change a handful of the L duals (config B: 3-14 of 256 from round 10 on; the rest moves at
// the 1e-14 level of the master's round-off), and then almost no block needs a pivot.  The CTA kernel still re-prices
// all n columns of every block and walks its whole prologue/epilogue: ~1 260 instructions per warp, 45 KB of reads and
// 0.17-0.19 ms per launch of 8192 blocks, whatever moved.  Here the engine keeps the duals it last APPLIED per linking
// row (`pi_ref`); `delta_rows_kernel` lists the rows whose dual left that value by more than 1e-13 relative and snaps
// the others to it, so that every kernel of the round prices with the same vector pi' (|pi' - pi| <= 1e-13 (1 + |pi|)
// per row: three orders below what the reduced-cost tolerance 1e-7 can see).  `delta_round_kernel` then gives every
// block ONE WARP, which only looks at the columns of D_k that sit on a listed row:
//   * priced cost of such a column, exactly as the CTA kernel computes it (order and rounding of src/dw_blas.c:106-122),
//     and its change against the saved cost (cost_delta);
//   * nonbasic column: its reduced cost moves by that change; basic column: the change times its tableau row (masked
//     items only) is added to the saved reduced costs — the same fma sequence per item, rows ascending, as the CTA
//     kernel's incremental update — and only the touched items are priced;
//   * nothing attractive: the touched reduced costs and saved costs are written back and the priced optimum becomes
//     c_k.x_k - pi'.(D_k x_k) from the block's standing proposal; the block is done (~6 KB of traffic).  Otherwise
//     nothing is written and the block joins the list the CTA kernel solves in the same round.
// More than DELTA_MAX_ROWS listed rows, a phase change, the first round after a cold solve / reset: the flag `full` is
// set on the device and the CTA kernel covers its whole static list as before (the delta launch exits at once).
// Reference path replaced: the warm-started glp_simplex call that needs zero iterations (src/dw_subprob.c:291-340).
#pragma once
#include "simplex_smem.cuh"

namespace dwk {

constexpr int DELTA_MAX_ROWS = 16;      // listed linking rows a delta round handles
constexpr int DELTA_MAX_COLS = 32;      // columns of one block on those rows (one per lane)

struct DeltaMeta { int nrows, full, count_b, count_g; };

struct DeltaIn {
    const int *rows;            // listed linking rows (DELTA_MAX_ROWS)
    const DeltaMeta *meta;
    int *next_list;             // blocks left for the CTA kernel
    int *next_count;
};

// (delta_rows_kernel / round_prep: pack.cuh — the dual screening shares a launch with the launch-order kernel and, in
//  multi-GPU rounds, with the duals' hand-shake)

__host__ __device__ inline size_t delta_warp_bytes()
{
    return sizeof(double) * 32 * 2          // weights of the listed tableau rows, cost changes of the nonbasic columns
         + sizeof(unsigned) * 32 * 4        // row-item masks of the listed rows (n <= 512: 4 words)
         + sizeof(int) * 32 * 2;            // listed tableau rows, slots of the nonbasic columns
}

template <int NW>
__global__ void __launch_bounds__(NW * 32, 8) delta_round_kernel(const BlockDev *blocks, const int *list, int count, SolveParams P, DeltaIn D)
{
    __shared__ __align__(16) char smem_raw[NW * (32 * 2 * 8 + 32 * 4 * 4 + 32 * 2 * 4)];
    __shared__ double2 stash_all[NW][4 * 32 * 2];           // new reduced costs of the touched items until the block is decided
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int e = blockIdx.x * NW + warp;
    if (e >= count) return;
    if (D.meta->full) return;                               // the CTA kernel takes the whole list this round
    const int k = list[e];
    const BlockDev &B = blocks[k];
    const int m = B.m, n = B.n, ld = tab_ld(n, true), rwd = B.rwd, nitems = ld >> 2;
    char *sp = smem_raw + (size_t)warp * delta_warp_bytes();
    double *wts = (double *)sp;                 sp += sizeof(double) * 32;
    double *ndl = (double *)sp;                 sp += sizeof(double) * 32;
    unsigned *smask = (unsigned *)sp;           sp += sizeof(unsigned) * 32 * 4;
    int *rws = (int *)sp;                       sp += sizeof(int) * 32;
    int *nsl = (int *)sp;
    const double tol_dj = P.tol_dj;
    const unsigned FULL = 0xffffffffu;

    auto handoff = [&]() { if (lane == 0) D.next_list[atomicAdd(D.next_count, 1)] = k; };

    // (independent loads are issued side by side: the warp's time is its chain of dependent round trips)
    const int dflag_in = B.dflag[0];
    const int nR = D.meta->nrows;
    int rb = 0, re = 0;
    if (lane < nR) { const int r = D.rows[lane]; rb = __ldg(B.dr_ptr + r); re = __ldg(B.dr_ptr + r + 1); }
    const double zo = standing_optimum(P, k, B.L, lane);   // needed only if the block stays; three round trips hidden here
    if (!(dflag_in > 0 && dflag_in < 32)) { handoff(); return; }     // no valid saved row, or the periodic full recomputation
    // ---- columns of D_k on the listed rows, one per lane
    int cnt = re - rb, incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += t; }
    const int total = __shfl_sync(FULL, incl, 31);
    if (total > DELTA_MAX_COLS) { handoff(); return; }
    int j = -1 - lane;                                      // (distinct dummies for the lanes without a column)
    {
        int src = -1, off = 0;
        for (int t = 0; t < nR; ++t) {
            const int it = __shfl_sync(FULL, incl, t), ct = __shfl_sync(FULL, cnt, t), bt = __shfl_sync(FULL, rb, t);
            if (src < 0 && lane < it && lane >= it - ct) { src = t; off = bt + (lane - (it - ct)); }
        }
        if (src >= 0) j = __ldg(B.dr_col + off);
    }
    {   // a column that sits on two listed rows is priced once
        const unsigned peers = __match_any_sync(FULL, j);
        if (j >= 0 && (__ffs(peers) - 1) != lane) j = -1 - lane;
    }
    // ---- its priced cost under the new duals and the change against the saved one
    double dl = 0.0, cs_new = 0.0;
    bool basic = false;
    int var = -1;
    int pos = -1;               // where the variable sits: tableau row of a basic one, column slot of a nonbasic one
    if (j >= 0) {
        var = m + j;
        const int pb = __ldg(B.dc_ptr + j), pe = __ldg(B.dc_ptr + j + 1);
        const double base = P.phase_one ? 0.0 : __ldg(B.c_master + j);
        const double sv = __ldcg(B.cssave + j);
        const signed char st = __ldcg(B.vstat + var);
        const int ps = __ldcg(B.pos + var);
        double y = 0.0;
        for (int t = pb; t < pe; ++t) y = __dadd_rn(y, __dmul_rn(__ldg(P.pi + __ldg(B.dc_row + t)), __ldg(B.dc_val + t)));
        basic = st == ST_BASIC;
        const bool in_range = ps >= 0 && ps < (basic ? m : n);
        const int owner = in_range ? __ldcg((basic ? B.head : B.nbv) + ps) : -1;
        cs_new = __dadd_rn(__dmul_rn(-1.0, y), base);
        dl = cost_delta(cs_new, sv);
        pos = owner == var ? ps : -1;
    }
    const bool act = j >= 0 && dl != 0.0;
    const unsigned actB = __ballot_sync(FULL, act && basic), actN = __ballot_sync(FULL, act && !basic);
    if (__any_sync(FULL, act && pos < 0)) { handoff(); return; }     // (inconsistent state image: let the CTA kernel look)
    // listed tableau rows in ascending order with their weights and masks; slots of the nonbasic columns
    const int nl = __popc(actB), nn = __popc(actN);
    if (act && basic) {
        int rank = 0;
        unsigned todo = actB;
        while (todo) { const int t = __ffs(todo) - 1; todo &= todo - 1; rank += __shfl_sync(actB, pos, t) < pos ? 1 : 0; }
        rws[rank] = pos; wts[rank] = dl;
        const unsigned *mi = B.rmask + (size_t)pos * rwd;
        for (int w_ = 0; w_ < 4; ++w_) smask[rank * 4 + w_] = w_ < rwd ? __ldcg(mi + w_) : 0u;
    }
    if (act && !basic) { const int idx = __popc(actN & ((1u << lane) - 1u)); nsl[idx] = pos; ndl[idx] = dl; }
    __syncwarp();
    // ---- touched items of this lane (item t = lane + 32 u): saved reduced costs + changes, then pricing
    double2 *stash = stash_all[warp];
    unsigned touched = 0;
    bool attractive = false;
    unsigned cells = 0;
    const double2 *T2 = reinterpret_cast<const double2 *>(B.T);
    const int n2 = ld >> 1;
    for (int a = lane; a < nl; a += 32)
        for (int w_ = 0; w_ < 4; ++w_) cells += __popc(smask[a * 4 + w_]);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int t = lane + 32 * u;
        if (t >= nitems) continue;
        unsigned have = 0;
        for (int a = 0; a < nl; ++a) have |= ((smask[a * 4 + u] >> lane) & 1u) << a;
        double add0 = 0.0, add1 = 0.0, add2 = 0.0, add3 = 0.0;
        bool nbhit = false;
        for (int a = 0; a < nn; ++a) {
            const int s = nsl[a];
            if ((s >> 2) == t) {
                const double v = ndl[a];
                const int x = s & 3;
                if (x == 0) add0 = v; else if (x == 1) add1 = v; else if (x == 2) add2 = v; else add3 = v;
                nbhit = true;
            }
        }
        if (!have && !nbhit) continue;
        touched |= 1u << u;
        const int s0 = 4 * t;
        double dsv[4];
        int vv[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            dsv[x] = 0.0; vv[x] = -1;
            if (s0 + x < n) { dsv[x] = __ldcg(B.dsave + s0 + x); vv[x] = __ldcg(B.nbv + s0 + x); }
        }
        // base as the CTA kernel forms it: (change of a nonbasic column, else 0) + saved reduced cost
        double2 a0 = make_double2(add0 + dsv[0], add1 + dsv[1]);
        double2 a1 = make_double2(add2 + dsv[2], add3 + dsv[3]);
        while (have) {
            double2 v0[4], v1[4];
            double wv[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                wv[r] = 0.0; v0[r] = make_double2(0.0, 0.0); v1[r] = v0[r];
                if (have) {
                    const int a = __ffs(have) - 1;
                    have &= have - 1;
                    const int i = rws[a];
                    wv[r] = wts[a];
                    v0[r] = __ldcg(T2 + (size_t)i * n2 + 2 * t);
                    v1[r] = __ldcg(T2 + (size_t)i * n2 + 2 * t + 1);
                }
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                a0.x = fma(wv[r], v0[r].x, a0.x); a0.y = fma(wv[r], v0[r].y, a0.y);
                a1.x = fma(wv[r], v1[r].x, a1.x); a1.y = fma(wv[r], v1[r].y, a1.y);
            }
        }
        stash[(u * 32 + lane) * 2] = a0; stash[(u * 32 + lane) * 2 + 1] = a1;
        const double dj[4] = {a0.x, a0.y, a1.x, a1.y};
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            if (s0 + x < n) {
                const signed char st = __ldcg(B.vstat + vv[x]);
                attractive |= (st == ST_LO && dj[x] < -tol_dj) || (st == ST_UP && dj[x] > tol_dj) ||
                              (st == ST_FREE && fabs(dj[x]) > tol_dj);
            }
        }
    }
    if (__any_sync(FULL, attractive)) { handoff(); return; }
    // ---- the block stays optimal: write back what moved
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        if (!((touched >> u) & 1u)) continue;
        const int s0 = 4 * (lane + 32 * u);
        const double2 a0 = stash[(u * 32 + lane) * 2], a1 = stash[(u * 32 + lane) * 2 + 1];
        const double dj[4] = {a0.x, a0.y, a1.x, a1.y};
#pragma unroll
        for (int x = 0; x < 4; ++x) if (s0 + x < n) B.dsave[s0 + x] = dj[x];
    }
    if (act) B.cssave[j] = cs_new;
    // priced optimum of the kept x_k from the standing proposal: c_k.x_k - pi'.(D_k x_k)
#pragma unroll
    for (int o = 16; o; o >>= 1) cells += __shfl_xor_sync(FULL, cells, o);
    if (lane == 0) {
        P.obj[k] = zo; P.status[k] = 5; P.last_iters[k] = 0;
        if (P.m_obj != nullptr) { P.m_obj[k] = zo; P.m_status[k] = 5; P.counters[8 * (size_t)k + 7] += 12; }
        P.counters[8 * (size_t)k + 6] += (unsigned long long)cells * 4ull;
        B.dflag[0] = dflag_in + 1;
    }
}

}  // namespace dwk
