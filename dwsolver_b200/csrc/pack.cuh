// pack.cuh — one round's proposals compacted into ONE contiguous host-visible record.
//
// The reference's up-link is K mailboxes {obj, unbounded, new_obj_coeff, len, ind[1..len],
// val[1..len]} (/root/reference/src/dw_subprob.h:44-94) that the master walks one by one
// (src/dw_phases.c:338-437).  The dense device arrays of the engine (stride-L ind/val) would cost
// 12*K*L bytes of D2H per round (25 MB at config B); the master only needs the len[k] real entries,
// and — when it passes the convexity duals r_k down — only the columns it is going to accept
// (r_k - obj_k > 1e-6, src/dw_phases.c:108,358 with TOLERANCE from src/dw.h:71).
//
// Two tiny kernels after the solve kernels of a round:
//   pack_scan_kernel   one CTA: exclusive scan of the shipped lengths -> colptr[K+1]
//   pack_copy_kernel   one warp per block: scalars + its column into the mapped pinned mirror
// The mirror is written by the GPU directly (zero-copy stores), so a round ends with ONE stream
// synchronisation and no cudaMemcpy whose size would have to be known on the host first.
#pragma once
#include <cuda_runtime.h>

namespace dwk {

struct PackDst {          // device-visible aliases of the mapped pinned host mirror
    double *obj, *cost, *val;
    int *status, *colptr, *ind;
    long long *nnz;
};

struct PackSrc {
    const double *obj, *cost, *val;   // K, K, K*L
    const int *status, *len, *ind;    // K, K, K*L
    const double *r;                  // K convexity duals (nullptr: ship every column)
    int K, L;
    double accept_tol;
    // segmented source (multi-GPU up-link): `recs` holds one engine record per rank, back to back, each
    // [obj seg_K | cost seg_K | val seg_K*L] doubles then [status seg_K | len seg_K | ind seg_K*L] ints
    const char *recs;
    size_t rec_bytes;
    int seg_K;
};

struct PackBlk {                      // where block k's fields live
    const double *obj, *cost, *val;
    const int *status, *len, *ind;
};

__device__ __forceinline__ PackBlk pack_block(const PackSrc &s, int k)
{
    PackBlk b;
    if (s.recs == nullptr) {
        b.obj = s.obj + k; b.cost = s.cost + k; b.val = s.val + (size_t)k * s.L;
        b.status = s.status + k; b.len = s.len + k; b.ind = s.ind + (size_t)k * s.L;
    } else {
        const int rk = k / s.seg_K, kk = k - rk * s.seg_K;
        const double *dp = reinterpret_cast<const double *>(s.recs + (size_t)rk * s.rec_bytes);
        const int *ip = reinterpret_cast<const int *>(dp + 2 * (size_t)s.seg_K + (size_t)s.seg_K * s.L);
        b.obj = dp + kk; b.cost = dp + s.seg_K + kk; b.val = dp + 2 * (size_t)s.seg_K + (size_t)kk * s.L;
        b.status = ip + kk; b.len = ip + s.seg_K + kk; b.ind = ip + 2 * (size_t)s.seg_K + (size_t)kk * s.L;
    }
    return b;
}

__device__ __forceinline__ int shipped_len(const PackSrc &s, int k)
{
    const PackBlk b = pack_block(s, k);
    int ln = b.len[0];
    if (s.r != nullptr && !(s.r[k] - b.obj[0] > s.accept_tol)) ln = 0;
    // a block that did not reach an optimum (1: iteration limit / numerical failure, 4: infeasible) proposes nothing;
    // an unbounded one (6) keeps its record — the host reports it (dw_solve.cc, consume)
    if (s.r != nullptr && (b.status[0] == 1 || b.status[0] == 4)) ln = 0;
    return ln;
}

template <int NT>
__global__ void __launch_bounds__(NT) pack_scan_kernel(PackSrc s, int *colptr_dev, PackDst d)
{
    __shared__ int warp_tot[NT / 32];
    __shared__ int carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < s.K; base += NT) {
        const int k = base + tid;
        const int v = k < s.K ? shipped_len(s, k) : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < warp; ++w) woff += warp_tot[w];
        const int c0 = carry;
        if (k < s.K) {
            const int ex = c0 + woff + incl - v;
            colptr_dev[k] = ex;
            d.colptr[k] = ex;
        }
        __syncthreads();
        if (tid == NT - 1) carry = c0 + woff + incl;
        __syncthreads();
    }
    if (tid == 0) {
        colptr_dev[s.K] = carry;
        d.colptr[s.K] = carry;
        d.nnz[0] = carry;
    }
}

template <int NT>
__global__ void __launch_bounds__(NT) pack_copy_kernel(PackSrc s, const int *colptr_dev, PackDst d)
{
    const int lane = threadIdx.x & 31;
    const int k = blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
    if (k >= s.K) return;
    const PackBlk blk = pack_block(s, k);
    if (lane == 0) {
        d.obj[k] = blk.obj[0];
        d.cost[k] = blk.cost[0];
        d.status[k] = blk.status[0];
    }
    const int b = colptr_dev[k], e = colptr_dev[k + 1];
    const int *si = blk.ind;
    const double *sv = blk.val;
    for (int t = lane; t < e - b; t += 32) {
        d.ind[b + t] = si[t];
        d.val[b + t] = sv[t];
    }
}

// ---- host mailboxes: blocks whose kernel does not write its mailbox itself (legacy and cluster paths), or every block
// when the mailboxes were not written in the previous round: one warp per listed block (list == nullptr: blocks 0..count-1)
struct MailDst { double *obj, *cost, *val; int *status, *len, *ind; };

template <int NT>
__global__ void __launch_bounds__(NT) mailbox_copy_kernel(PackSrc s, const int *list, int count, MailDst d, unsigned long long *counters)
{
    const int lane = threadIdx.x & 31;
    const int e = blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
    if (e >= count) return;
    const int k = list ? list[e] : e;
    const PackBlk blk = pack_block(s, k);
    const int ln = blk.len[0];
    if (lane == 0) {
        d.obj[k] = blk.obj[0]; d.cost[k] = blk.cost[0]; d.status[k] = blk.status[0]; d.len[k] = ln;
        if (counters) counters[8 * (size_t)k + 7] += 24ull + 12ull * (unsigned long long)ln;
    }
    for (int t = lane; t < ln; t += 32) {
        d.ind[(size_t)k * s.L + t] = blk.ind[t];
        d.val[(size_t)k * s.L + t] = blk.val[t];
    }
}

// ---- multi-GPU round control over peer memory ---------------------------------------------------------------------
// The up-link buffer in rank 0's HBM (every rank maps it: CUDA IPC over NVLink) ends with a control block.  A round's two
// hand-shakes — duals down, proposals complete — are a sequence number and a counter in that block instead of two NCCL
// collectives (a 2 KB broadcast and a one-int all-reduce cost 20-30 us each on 8 GPUs; a flag costs one NVLink round trip):
//   rank 0   prep_job (comm_mode 1): copies pi into the block, fence, seq = s
//   others   prep_job (comm_mode 2): waits until seq >= s (system-scope acquire loads), copies pi into its own buffer
//   all      comm_signal_kernel after the solve kernels: fence, done += 1 (their proposal stores into rank 0's HBM are
//            ordered before it)
//   rank 0   comm_signal_wait_kernel: counts itself in and waits until done >= world x rounds; what follows on its stream
//            may read the up-link buffer
// A wait that lasts longer than 10 s traps (a peer died): the process fails instead of hanging.
struct UplinkCtl {
    unsigned long long seq, done, pad0, pad1;
    double pi[1];                      // L doubles
};
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void spin_until(const unsigned long long *flag, unsigned long long want)
{
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(flag) < want) {
        if (global_ns() - t0 > 10000000000ull) __trap();
        __nanosleep(100);
    }
}
__global__ void comm_signal_kernel(UplinkCtl *ctl)
{
    __threadfence_system();
    atomicAdd_system(&ctl->done, 1ull);
}
__global__ void comm_signal_wait_kernel(UplinkCtl *ctl, unsigned long long want)       // rank 0: both in one launch
{
    __threadfence_system();
    atomicAdd_system(&ctl->done, 1ull);
    spin_until(&ctl->done, want);
}

// ---- start of a priced round: one CTA's worth of work that rides along with another small launch ----------------------
// (a) multi-GPU: the duals' hand-shake (publish on rank 0, fetch elsewhere), (b) delta rounds (delta_tg.cuh): which duals
// moved since they were last applied — pi_ref is updated in place and is the vector the round prices with.  Launched as
// its own kernel (delta_rows_kernel) or as the second CTA of the launch-order kernel below: every launch less is ~4 us of a
// round that may only last 40.
struct PrepJob {
    const double *pi;           // this round's duals on the device (multi-GPU fetch: written first, see comm_mode)
    double *pi_ref;             // nullptr: no dual screening
    int L, force_full;
    int *rows;
    DeltaMeta *meta;
    UplinkCtl *ctl;             // multi-GPU control block; comm_mode 0: none, 1: publish pi, 2: fetch pi
    unsigned long long seq;
    int comm_mode;
};
__device__ __forceinline__ void prep_job(const PrepJob &J)
{
    __shared__ int s_n;
    const int nt = blockDim.x;
    if (J.comm_mode == 1) {
        for (int r = threadIdx.x; r < J.L; r += nt) J.ctl->pi[r] = J.pi[r];
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(&J.ctl->seq), "l"(J.seq) : "memory");
        }
    } else if (J.comm_mode == 2) {
        if (threadIdx.x == 0) spin_until(&J.ctl->seq, J.seq);
        __syncthreads();
        double *dst = const_cast<double *>(J.pi);
        for (int r = threadIdx.x; r < J.L; r += nt) {
            double v;
            asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(J.ctl->pi + r) : "memory");
            dst[r] = v;
        }
        __syncthreads();        // (each thread screens below exactly the rows it has just written)
    }
    if (J.pi_ref == nullptr) return;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    for (int r = threadIdx.x; r < J.L; r += nt) {
        const double p = J.pi[r], ref = J.pi_ref[r];
        const bool moved = J.force_full || !(fabs(p - ref) <= 1e-13 * (1.0 + fabs(p)));
        if (moved) {
            J.pi_ref[r] = p;
            const int pos = atomicAdd(&s_n, 1);
            if (pos < DELTA_MAX_ROWS) J.rows[pos] = r;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        J.meta->nrows = s_n;
        J.meta->full = (J.force_full || s_n > DELTA_MAX_ROWS) ? 1 : 0;
        J.meta->count_b = 0; J.meta->count_g = 0;
    }
}
__global__ void __launch_bounds__(256) delta_rows_kernel(PrepJob J) { prep_job(J); }

// ---- longest-processing-time-first launch order -----------------------------------------------------
// A round's kernel time is max(total work / resident CTAs, longest block).  Blocks differ by 100x in
// pivots (most need none), and the blocks that pivoted a lot in the previous round tend to do so
// again; dispatching them first keeps the long poles off the tail of the launch.  One CTA orders the
// (<= 8192-entry) block list by last round's iteration count, descending (counting sort in shared memory).
template <int NT, int CAP>
__global__ void __launch_bounds__(NT) order_by_last_iterations_kernel(int *list, int count, const long long *last_iters, PrepJob J)
{
    if (blockIdx.x == 1) { prep_job(J); return; }          // (grid of 2: the round's preparation rides along)
    // Counting sort by min(iterations, NB - 1), descending (the bitonic sort of 8192 64-bit keys this replaces took 80 us
    // in each of the rounds that pivot — 0.9 ms of a 30 ms step at config B).  Blocks with the same count keep no
    // particular order: the order of a launch list only decides when a block starts, never what it computes.
    constexpr int NB = 1024;
    static_assert(NT == NB, "one thread per bucket");
    extern __shared__ unsigned long long key[];          // CAP x 8 bytes (dynamic, opt-in): ids | buckets
    int *ids = reinterpret_cast<int *>(key);
    unsigned short *bk = reinterpret_cast<unsigned short *>(ids + CAP);
    __shared__ int hist[NB], offs[NB], wsum[NB / 32];
    // nothing to gain when the previous round was light everywhere (late DW rounds): keep the order
    {
        int heavy = 0;
        for (int e = threadIdx.x; e < count; e += NT) heavy |= last_iters[list[e]] >= 16;
        if (!__syncthreads_or(heavy)) return;
    }
    hist[threadIdx.x] = 0;
    __syncthreads();
    for (int e = threadIdx.x; e < count; e += NT) {
        const int k = list[e];
        long long it = last_iters[k];
        if (it > NB - 1) it = NB - 1;
        if (it < 0) it = 0;
        const int b = NB - 1 - (int)it;                    // bucket 0 = longest
        ids[e] = k; bk[e] = (unsigned short)b;
        atomicAdd(&hist[b], 1);
    }
    __syncthreads();
    {   // exclusive prefix sum over the buckets: thread t owns bucket t
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        const int v = hist[threadIdx.x];
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) wsum[w] = inc;
        __syncthreads();
        if (w == 0) {
            int s = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
            wsum[lane] = s;
        }
        __syncthreads();
        offs[threadIdx.x] = inc - v + (w ? wsum[w - 1] : 0);
        hist[threadIdx.x] = 0;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < count; e += NT) {
        const int b = bk[e];
        list[offs[b] + atomicAdd(&hist[b], 1)] = ids[e];
    }
}

// ---- multi-vector pricing: cs_p = (phase_one ? 0 : c_k) - D_k' pi_p for P dual vectors at once ---------
// For a block whose D_k is dense enough (config C: 25 % of 256 x 4096) this is a real FP64 contraction
// Y (n x P) = Dt (n x L) . Pi' (L x P), done with FP64 tensor-core MMAs (mma.sync m8n8k4 f64 -> DMMA).  The reference
// prices one vector at a time with a COO loop (src/dw_subprob.c:291-293 -> src/dw_blas.c:106-122).
__global__ void densify_dt_kernel(const BlockDev *blocks, const int *list, int count)
{
    const int e = blockIdx.x;
    if (e >= count) return;
    const BlockDev &B = blocks[list[e]];
    double *Dt = const_cast<double *>(B.Dt);
    const int n = B.n, L = B.L;
    for (size_t t = threadIdx.x; t < (size_t)n * L; t += blockDim.x) Dt[t] = 0.0;
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += blockDim.x)
        for (int t = B.dc_ptr[j]; t < B.dc_ptr[j + 1]; ++t) Dt[(size_t)j * L + B.dc_row[t]] += B.dc_val[t];
}

// grid (dense blocks, ceil(n / 64)); 8 warps, warp w prices rows [64 bx + 8 w, +8) for all P vectors (P <= 32, padded to
// a multiple of 8 with zero vectors in shared memory).  out[p * ntot + xoff + j] = base_j - sum_i Dt[j][i] Pi[p][i].
__global__ void __launch_bounds__(256) price_multi_dmma_kernel(const BlockDev *blocks, const int *list, const double *Pi, int P,
                                                               int phase_one, long long ntot, double *out)
{
    extern __shared__ double spi[];              // [32][L + 4]: Pi rows, zero beyond P
    const BlockDev &B = blocks[list[blockIdx.x]];
    const int n = B.n, L = B.L, ldp = L + 4;
    for (int e = threadIdx.x; e < 32 * L; e += 256) { const int p = e / L, i = e - p * L; spi[p * ldp + i] = p < P ? Pi[(size_t)p * L + i] : 0.0; }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, ar = lane >> 2, ak = lane & 3;
    const int j0 = blockIdx.y * 64 + warp * 8;
    if (j0 >= n) return;
    const int jr = min(j0 + ar, n - 1);
    const double *drow = B.Dt + (size_t)jr * L;
    double acc[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
    const int PT = (P + 7) / 8;
    for (int k0 = 0; k0 < L; k0 += 4) {
        const double a = (k0 + ak < L) ? __ldg(drow + k0 + ak) : 0.0;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            if (t < PT) {
                const double b = (k0 + ak < L) ? spi[(t * 8 + ar) * ldp + k0 + ak] : 0.0;
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(acc[t][0]), "+d"(acc[t][1]) : "d"(a), "d"(b));
            }
        }
    }
    // C[row = ar][col = 2 ak + {0,1}] of p-tile t  ->  vector p = 8 t + 2 ak + {0,1}, column j0 + ar
    const int j = j0 + ar;
    if (j < n) {
        const double base = phase_one ? 0.0 : __ldg(B.c_master + j);
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int p = t * 8 + 2 * ak + h;
                if (t < PT && p < P) out[(size_t)p * ntot + B.xoff + j] = base - acc[t][h];
            }
    }
}

// ---- device-resident proposal history (see dwgpu_history_push / _combine in include/dwgpu.h) -------------
// one CTA per listed block: hist[dst[i] ..) <- x[src[i] ..)
__global__ void __launch_bounds__(256) history_push_kernel(const double *x, double *hist, const long long *src,
                                                           const long long *dst, const int *len, int count)
{
    const int i = blockIdx.x;
    if (i >= count) return;
    const double *s = x + src[i];
    double *d = hist + dst[i];
    for (int j = threadIdx.x; j < len[i]; j += 256) d[j] = s[j];
}

// one CTA per block that has terms: x_out[xoff ..) = sum_t w[t] * hist[id[t] ..), terms applied in the given order
__global__ void __launch_bounds__(256) history_combine_kernel(const double *hist, double *x_out, const long long *xoff,
                                                              const int *len, const int *term_ptr, const long long *term_id,
                                                              const double *term_w, int nblocks)
{
    const int b = blockIdx.x;
    if (b >= nblocks) return;
    for (int j = threadIdx.x; j < len[b]; j += 256) {
        double s = 0.0;
        for (int t = term_ptr[b]; t < term_ptr[b + 1]; ++t) s = fma(term_w[t], hist[term_id[t] + j], s);
        x_out[xoff[b] + j] = s;
    }
}

}  // namespace dwk
