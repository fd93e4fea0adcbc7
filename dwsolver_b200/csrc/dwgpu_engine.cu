// dwgpu_engine.cu — host side of the B200 batched subproblem engine behind include/dwgpu.h.
//
// Owns every device allocation (one pooled arena per kind, per-block offsets), builds the
// per-block descriptors, picks the kernel path per block and launches the rounds.  The
// reference's equivalent state is the per-thread glp_prob plus the subprob_struct mailbox
// (/root/reference/src/dw_subprob.h:44-94); here it is the (T, beta, head, nbv, vstat)
// arrays that stay resident in HBM between rounds.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/dwgpu.h"
#include "simplex_cta.cuh"
#include "simplex_smem.cuh"
#include "simplex_warp.cuh"
#include "screen_tg.cuh"
#include "delta_tg.cuh"
#include "simplex_sparse.cuh"
#include "pack.cuh"
#include "simplex_cluster.cuh"

namespace {

thread_local std::string g_err;

// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to the (device, kernel) pair, not to an engine: a second engine
// with smaller blocks must not lower the limit under a first one that is still launching (several engines live in one
// process when dw_solve shards the blocks, DWSOLVER_DEVICES).  The limit only ever grows; launches ask for what they need.
std::mutex g_attr_mu;
std::map<std::pair<int, const void *>, int> g_attr_max;
cudaError_t allow_dynamic_smem(int device, const void *kernel, int bytes)
{
    std::lock_guard<std::mutex> lock(g_attr_mu);
    int &cur = g_attr_max[std::make_pair(device, kernel)];
    if (bytes <= cur) return cudaSuccess;
    const cudaError_t rc = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (rc == cudaSuccess) cur = bytes;
    return rc;
}

int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}

#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(DWGPU_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));    \
    } while (0)

constexpr int NT_SMEM = 256;
constexpr int NT_GMEM = 1024;
// global-tableau kernel: 64-thread CTAs at 6 per SM give the best throughput when a shard has many
// waves of blocks; 128-thread CTAs at 4 per SM finish a single block ~25 % sooner, which is what bounds a
// launch when the shard is only a wave or two deep (multi-GPU shards of config B: 1024 blocks at N = 8)
constexpr int NT_TG = 64, NT_TG_WIDE = 128;
constexpr int SCREEN_NW = 4;        // blocks (warps) per CTA of the screening launch

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    cudaError_t alloc(size_t count)
    {
        n = count;
        return cudaMalloc((void **)&p, sizeof(T) * std::max<size_t>(count, 1));
    }
    cudaError_t upload(const std::vector<T> &h)
    {
        cudaError_t e = alloc(h.size());
        if (e != cudaSuccess || h.empty()) return e;
        return cudaMemcpy(p, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice);
    }
    template <typename H>
    cudaError_t upload_buf(const H &h)
    {
        cudaError_t e = alloc(h.size());
        if (e != cudaSuccess || h.empty()) return e;
        return cudaMemcpy(p, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice);
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
    }
};

// plain uninitialised host array (malloc), vector-like enough for the staging code and DevBuf::upload
template <typename T>
struct HostBuf {
    T *p = nullptr;
    size_t n = 0;
    explicit HostBuf(size_t count) : n(count) { p = (T *)std::malloc(sizeof(T) * std::max<size_t>(count, 1)); }
    ~HostBuf() { std::free(p); }
    HostBuf(const HostBuf &) = delete;
    HostBuf &operator=(const HostBuf &) = delete;
    bool ok() const { return p != nullptr; }
    T *data() { return p; }
    const T *data() const { return p; }
    size_t size() const { return n; }
    bool empty() const { return n == 0; }
    T &operator[](size_t i) { return p[i]; }
    const T &operator[](size_t i) const { return p[i]; }
};

inline double norm_bound(double v, bool upper)
{
    if (std::isnan(v)) return upper ? INFINITY : -INFINITY;
    if (v >= DWGPU_INF) return INFINITY;
    if (v <= -DWGPU_INF) return -INFINITY;
    return v;
}

}  // namespace

extern "C" void dwgpu_comm_release(struct dwgpu_engine *e);

struct dwgpu_engine {
    int K = 0, L = 0, device = 0;
    dwgpu_options opt{};
    cudaStream_t stream = nullptr;
    std::vector<int> m, n;
    std::vector<long long> xoff;   // prefix sums of n
    long long ntot = 0;
    std::vector<int> path;         // 1 smem, 2 gmem
    // kernel lists: smem-generic, legacy global, specialised 64x128 (A), global-T generic (g), global-T 256x512 (B)
    std::vector<int> list_smem, list_gmem, list_a, list_g, list_b, list_c;
    int cluster_size = 1;              // CTAs per block on the cluster path (list_c)
    size_t smem_dyn_c = 0;
    size_t smem_dyn_g = 0;
    int tg_threads = 0;                // 0: 128-thread CTAs; DWGPU_TG_THREADS=32 (one warp per block)|64|128|256 forces one
    size_t smem_dyn_w = 0;             // warp kernel, generic shapes (list_g)
    bool warp_ok_g = false;            // every block of list_g fits the warp kernel
    DevBuf<int> redo_list, redo_count; // blocks the warp kernel handed over to the CTA kernel this round
    DevBuf<unsigned char> redo_moved;
    // path 5 (simplex_sparse.cuh): packed tableau in shared memory.  Hand-over lists of a round: [0] probe -> solve launch,
    // [1] small heap -> large heap, [2] -> dense path-3 kernel
    std::vector<int> list_s;
    DevBuf<int> d_list_s, sp_lists, sp_counts, spflag;
    DevBuf<unsigned char> sp_moved, sp_reb;
    DevBuf<unsigned long long> sp_stats;   // hand-over / re-pack counters (dwgpu_debug_sparse_stats)
    int sp_nt = 256;                   // threads of the solve kernel (DWGPU_SPARSE_THREADS = 128 | 256)
    int sp_H0 = 0, sp_H1 = 0;          // heap doubles of the two solve launches
    size_t sp_smem0 = 0, sp_smem1 = 0, sp_smem_probe = 0;
    bool sp_probe = true;              // DWGPU_SPARSE_PROBE=0: every round goes straight to the solve launch
    // screening launch of a priced round on path 3 (screen_tg.cuh): one warp per block finishes the blocks that need no
    // pivot; the solve launch then covers the device-side list of the others, most attractive columns first
    bool screen = false;               // DWGPU_SCREEN=1 switches it on.  Measured on config B (profiles/r2o_*): the screening
                                       // launch costs what a no-pivot round of the CTA kernel costs (0.19 ms: both are bound by the
                                       // ~55 KB of per-block state a round reads, not by the 4 resident CTAs), so it only adds
                                       // to the rounds that pivot; ordering the solve launch by attractive columns does not
                                       // shorten its tail either
    bool screen_sort = true;           // DWGPU_SCREEN_SORT=0: the solve launch keeps the (atomic) arrival order
    bool screen_ok_b = false, screen_ok_g = false;
    int screen_wb_b = 0, screen_wb_g = 0;      // shared memory per warp
    DevBuf<int> scr_list, scr_count;   // 2 x K ids, 2 counters (list_b, list_g)
    DevBuf<unsigned> scr_key;          // by block id
    // delta rounds (delta_tg.cuh): rounds in which few duals moved are screened by one warp per block against the duals
    // last applied per linking row; DWGPU_DELTA=0 switches it off (then the duals are used exactly as given)
    bool delta = true;
    int delta_min_blocks = 1536;       // lists shorter than this go through the CTA kernel alone (same duals, same results: below
                                       // ~2.5 waves of CTAs one launch is shorter than delta launch + CTA launch — measured with
                                       // profiles/scan_delta.py: 1024 blocks 37 vs 47 us, 2048 blocks 58 vs 49 us); DWGPU_DELTA_MIN_BLOCKS
    bool delta_ok_b = false, delta_ok_g = false;
    DevBuf<double> pi_ref;             // duals last applied, by linking row: the vector a round prices with
    DevBuf<int> dl_rows, dl_meta;      // listed rows; DeltaMeta
    DevBuf<int> dl_list;               // 2 x K: blocks left for the CTA kernel (list_b, list_g)
    DevBuf<int> dl_pos;                // position of every variable of the blocks above (BlockDev::pos)
    bool pi_ref_valid = false;         // false: the next round lists every row (after create, cold solve, reset)
    int last_phase = -1;
    int warp_bail_after = 0;           // test hook (DWGPU_WARP_BAIL_AFTER): hand over after this many basis changes
    bool reorder = true;
    bool have_proposals = false;       // a round has written every block's proposal since create / reset
    // multi-GPU: the up-link lives in rank 0's HBM (one mailbox per block of the WHOLE instance); every rank's kernels
    // store their blocks' proposals straight into it over NVLink peer memory (CUDA IPC mapping), NCCL carries the duals
    // down and closes the round
    char *uplink_base = nullptr;       // rank 0: owned allocation; other ranks: IPC mapping
    bool uplink_owned = false;
    int uplink_K = 0, uplink_first = 0;
    dwk::MailDst uplink{};             // this rank's window (pointers already offset by its first block)
    dwk::MailDst uplink_all{};         // the whole buffer (rank 0 reads it)
    bool uplink_valid = false;         // the previous round wrote the window
    // several dual vectors per round (dwgpu_iterate_multi)
    DevBuf<double> Dt;                 // dense D_k' (n x L) of the blocks whose D_k is dense enough for the pricing GEMM
    std::vector<int> list_dense;       // those blocks
    DevBuf<int> d_list_dense;
    DevBuf<double> multi_pi, multi_cs; // P x L dual vectors, P x ntot priced costs
    int multi_P = 0;                   // capacity of the above and of the pinned records below
    const double *cs_pre_now = nullptr; // priced costs of the vector the next launch_round solves for (nullptr: kernels price)
    char *h_mpack = nullptr, *d_mpack = nullptr;
    std::vector<dwk::PackDst> mpk_dev, mpk_host;
    DevBuf<long long> hp_src, hp_dst;  // history push: persistent index staging (no cudaMalloc per round)
    DevBuf<int> hp_len;
    size_t hp_cap = 0;
    ncclComm_t comm = nullptr;
    int comm_world = 0, comm_rank = 0;
    // round hand-shakes over the control block at the end of the up-link buffer (pack.cuh: UplinkCtl) instead of NCCL
    // collectives; DWGPU_COMM=nccl restores the broadcast + all-reduce
    bool comm_p2p = true;
    dwk::UplinkCtl *uplink_ctl = nullptr;
    unsigned long long comm_seq = 0, comm_rounds = 0;
    int prep_comm_mode = 0;            // set by dwgpu_round_comm for the launch_round that follows: 1 publish, 2 fetch the duals
    unsigned long long prep_seq = 0;
    DevBuf<int> comm_flag;
    size_t smem_dyn = 0;
    int max_smem_optin = 0;
    // device arenas
    DevBuf<double> T, bnd, dsave, cssave, a_v, dc_val, dr_val, c_file, c_master, x, ws;
    DevBuf<int> a_rp, a_ci, dc_ptr, dc_row, dr_ptr, dr_col;
    DevBuf<char> pst;
    DevBuf<int> dflag;
    DevBuf<unsigned> rmask;            // path 3: row-item masks
    DevBuf<dwk::BlockDev> blocks;
    DevBuf<int> d_list_smem, d_list_gmem, d_list_a, d_list_g, d_list_b, d_list_c;
    // round inputs / outputs
    DevBuf<double> pi;
    DevBuf<char> rec;                  // [obj K | cost K | val K*L] doubles, then [status K | len K | ind K*L] ints
    size_t rec_bytes = 0;
    struct View { double *p = nullptr; } obj, cost, val;
    struct IView { int *p = nullptr; } status, len, ind;
    DevBuf<int> colptr;                // K+1, device scratch of the pack kernel
    // mapped pinned host mirror written by the pack kernel: header + packed columns
    char *h_pack = nullptr, *d_pack = nullptr;
    size_t pack_bytes = 0;
    dwk::PackDst pk_dev{}, pk_host{};  // the same mirror through its device alias / its host address
    DevBuf<double> r_dev;              // K convexity duals of the packed call
    double *h_r = nullptr;
    // second mirror for dwgpu_pack_gathered (K_total blocks of all ranks), allocated on first use
    char *h_gpack = nullptr, *d_gpack = nullptr;
    int gpack_K = 0;
    dwk::PackDst gpk_dev{}, gpk_host{};
    // host mailboxes (dwgpu_*_mailboxes): mapped pinned arrays, created by the first call
    char *h_mail = nullptr, *d_mail = nullptr;
    dwk::MailDst mail_dev{}, mail_host{};
    bool mail_valid = false;           // the previous round wrote the mailboxes (unchanged blocks may keep theirs)
    DevBuf<int> gcolptr;
    // proposal history: one growing device arena; per kept copy its offset, block and length
    DevBuf<double> hist;
    long long hist_used = 0, hist_cap = 0;
    std::vector<long long> hist_off;
    std::vector<int> hist_block;
    long long *h_dbg = nullptr, *d_dbg = nullptr;   // DWGPU_DEBUG=1: progress markers the kernels leave behind
    DevBuf<unsigned long long> counters;
    DevBuf<long long> last_iters;
    // pinned staging
    double *h_pi = nullptr;
    long long solves = 0;
    int last_launches = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // gmem start/end, smem start/end
    // A round (or a reset) may run on a caller's stream; everything that READS the engine's device arrays afterwards runs on
    // e->stream.  ev_done is recorded behind the last launch on whichever stream it used and the readers wait for it, so a
    // fetch right after dwgpu_iterate_device(..., other_stream) is ordered without the caller synchronising.
    cudaEvent_t ev_done = nullptr;
    bool ev_done_pending = false;
    bool ev_used[2] = {false, false};
    bool initialised = false;
};

extern "C" {

const char *dwgpu_last_error(void) { return g_err.c_str(); }
const char *dwgpu_version(void) { return "dwgpu 0.2 (sm_100a)"; }


static int launch_round(dwgpu_engine *e, const double *d_pi, int phase_one, int initial, cudaStream_t st, bool mail, const dwk::MailDst *direct);

// ---- multi-GPU: NCCL from the C layer (replaces the reference's cond-broadcast of the duals and its
// semaphore/queue up-link, src/dw_phases.c:338-350, 526-548 and src/dw_subprob.c:417-427) --------------------------
namespace {
struct NcclApi {
    void *lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;
// NCCL is bound at run time: a single-GPU user of libdwgpu.so needs no NCCL at all, and a process that already
// carries one (PyTorch's) shares it — dlopen by soname returns the loaded copy
int nccl_api()
{
    std::lock_guard<std::mutex> lock(g_nccl_mu);
    if (g_nccl.lib) return 0;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail(DWGPU_E_STATE, std::string("NCCL is not available: ") + dlerror());
#define NSYM(field, name) g_nccl.field = (decltype(g_nccl.field))dlsym(h, name); if (!g_nccl.field) return fail(DWGPU_E_STATE, "NCCL symbol missing: " name);
    NSYM(GetUniqueId, "ncclGetUniqueId") NSYM(CommInitRank, "ncclCommInitRank") NSYM(CommDestroy, "ncclCommDestroy")
    NSYM(Broadcast, "ncclBroadcast") NSYM(AllReduce, "ncclAllReduce") NSYM(GetErrorString, "ncclGetErrorString")
#undef NSYM
    g_nccl.lib = h;
    return 0;
}
#define NC(call)                                                                                  \
    do {                                                                                          \
        ncclResult_t r_ = (call);                                                                 \
        if (r_ != ncclSuccess) return fail(DWGPU_E_CUDA, std::string(#call) + ": " + g_nccl.GetErrorString(r_)); \
    } while (0)
}  // namespace

void dwgpu_comm_release(dwgpu_engine *e)
{
    if (e && e->comm && g_nccl.CommDestroy) { g_nccl.CommDestroy(e->comm); e->comm = nullptr; }
}

int dwgpu_comm_unique_id(void *id128)
{
    if (!id128) return fail(DWGPU_E_ARGS, "id buffer is NULL");
    if (int rc = nccl_api()) return rc;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    NC(g_nccl.GetUniqueId((ncclUniqueId *)id128));
    return 0;
}

int dwgpu_comm_init(dwgpu_engine *e, const void *id128, int world, int rank)
{
    if (!e || !id128 || world < 1 || rank < 0 || rank >= world) return fail(DWGPU_E_ARGS, "bad arguments");
    if (int rc = nccl_api()) return rc;
    CU(cudaSetDevice(e->device));
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    NC(g_nccl.CommInitRank(&e->comm, world, id, rank));
    e->comm_world = world; e->comm_rank = rank;
    CU(e->comm_flag.alloc(2));
    CU(cudaMemset(e->comm_flag.p, 0, 2 * sizeof(int)));
    return 0;
}

static size_t uplink_ctl_offset(size_t K, size_t KL)
{
    return (sizeof(double) * (2 * K + KL) + sizeof(int) * (2 * K + KL) + 64 + 255) & ~(size_t)255;
}

// Rank 0 allocates one mailbox per block of the whole instance in its HBM and publishes a CUDA IPC handle (64 bytes,
// sent to the other ranks by whatever the host uses for its rendezvous).
int dwgpu_uplink_create(dwgpu_engine *e, int K_total, void *ipc_handle64, dwgpu_proposals *dev_out)
{
    if (!e || K_total < e->K || !ipc_handle64) return fail(DWGPU_E_ARGS, "bad arguments");
    CU(cudaSetDevice(e->device));
    const size_t K = (size_t)K_total, KL = K * (size_t)e->L;
    const size_t ctl_off = uplink_ctl_offset(K, KL);
    const size_t bytes = ctl_off + sizeof(dwk::UplinkCtl) + sizeof(double) * (size_t)std::max(e->L, 1);
    CU(cudaMalloc((void **)&e->uplink_base, bytes));
    CU(cudaMemset(e->uplink_base, 0, bytes));
    e->uplink_owned = true;
    e->uplink_K = K_total;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, e->uplink_base));
    std::memcpy(ipc_handle64, &h, sizeof(h));
    double *dd = (double *)e->uplink_base;
    int *ii = (int *)(dd + 2 * K + KL);
    e->uplink_all = dwk::MailDst{dd, dd + K, dd + 2 * K, ii, ii + K, ii + 2 * K};
    if (dev_out) {
        dev_out->obj = e->uplink_all.obj; dev_out->cost = e->uplink_all.cost; dev_out->val = e->uplink_all.val;
        dev_out->status = e->uplink_all.status; dev_out->len = e->uplink_all.len; dev_out->ind = e->uplink_all.ind; dev_out->x = nullptr;
    }
    return 0;
}

// Every rank (rank 0 too, with ipc_handle64 = NULL): this engine's blocks are blocks k_first .. k_first + K - 1 of the
// instance; from now on dwgpu_round_comm() makes the kernels store their proposals into that window of rank 0's buffer.
int dwgpu_uplink_attach(dwgpu_engine *e, const void *ipc_handle64, int K_total, int k_first)
{
    if (!e || K_total < 1 || k_first < 0 || k_first + e->K > K_total) return fail(DWGPU_E_ARGS, "bad arguments");
    CU(cudaSetDevice(e->device));
    if (ipc_handle64) {
        cudaIpcMemHandle_t h;
        std::memcpy(&h, ipc_handle64, sizeof(h));
        void *p = nullptr;
        CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        e->uplink_base = (char *)p;
        e->uplink_owned = false;
        e->uplink_K = K_total;
    } else if (!e->uplink_base || e->uplink_K != K_total) return fail(DWGPU_E_STATE, "rank 0 attaches after dwgpu_uplink_create");
    const size_t K = (size_t)K_total, KL = K * (size_t)e->L;
    double *dd = (double *)e->uplink_base;
    int *ii = (int *)(dd + 2 * K + KL);
    e->uplink_all = dwk::MailDst{dd, dd + K, dd + 2 * K, ii, ii + K, ii + 2 * K};
    const size_t f = (size_t)k_first;
    e->uplink = dwk::MailDst{dd + f, dd + K + f, dd + 2 * K + f * e->L, ii + f, ii + K + f, ii + 2 * K + f * e->L};
    e->uplink_first = k_first;
    e->uplink_valid = false;
    e->uplink_ctl = (dwk::UplinkCtl *)(e->uplink_base + uplink_ctl_offset(K, KL));
    if (const char *cv = std::getenv("DWGPU_COMM")) e->comm_p2p = std::strcmp(cv, "nccl") != 0;
    return 0;
}

// One DW round of a sharded instance, everything enqueued on ONE stream with no host step in between:
//   duals down   ncclBroadcast of pi (L doubles, root = rank 0)                       [skipped for the initial solve]
//   solve        this rank's blocks; the kernels' epilogues store each block's proposal into rank 0's up-link buffer
//                over NVLink (a block that did not move rewrites 12 bytes)
//   close        a one-int ncclAllReduce: when it completes on rank 0, every rank's stores have landed
int dwgpu_round_comm(dwgpu_engine *e, double *pi_dev, int phase_one, int initial, void *stream)
{
    if (!e || !e->initialised) return fail(DWGPU_E_STATE, "engine not initialised");
    if (!e->comm || !e->uplink_base) return fail(DWGPU_E_STATE, "dwgpu_comm_init / dwgpu_uplink_attach first");
    if (!initial && e->L > 0 && !pi_dev) return fail(DWGPU_E_ARGS, "pi_dev is NULL");
    CU(cudaSetDevice(e->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
    const bool p2p = e->comm_p2p && e->uplink_ctl != nullptr;
    if (!initial && e->L > 0) {
        if (p2p) {          // duals down through the control block in rank 0's HBM: part of the round's preparation launch
            ++e->comm_seq;
            e->prep_comm_mode = e->comm_rank == 0 ? 1 : 2;
            e->prep_seq = e->comm_seq;
        } else NC(g_nccl.Broadcast(pi_dev, pi_dev, (size_t)e->L, ncclDouble, 0, e->comm, st));
    }
    const bool direct = e->uplink_valid;
    int rc = launch_round(e, pi_dev, phase_one ? 1 : 0, initial ? 1 : 0, st, false, &e->uplink);
    if (rc) return rc;
    dwk::PackSrc s;
    std::memset(&s, 0, sizeof(s));
    s.obj = e->obj.p; s.cost = e->cost.p; s.val = e->val.p; s.status = e->status.p; s.len = e->len.p; s.ind = e->ind.p;
    s.K = e->K; s.L = e->L;
    auto copy = [&](const int *list, size_t count) -> cudaError_t {
        if (!count) return cudaSuccess;
        dwk::mailbox_copy_kernel<256><<<(unsigned)((count + 7) / 8), 256, 0, st>>>(s, list, (int)count, e->uplink, e->counters.p);
        e->last_launches++;
        return cudaGetLastError();
    };
    // kernels that do not write mailboxes themselves (legacy, cluster) and the first round after a reset: copied
    if (!direct) CU(copy(nullptr, (size_t)e->K));
    else { CU(copy(e->d_list_gmem.p, e->list_gmem.size())); CU(copy(e->d_list_c.p, e->list_c.size())); }
    e->uplink_valid = true;
    if (p2p) {              // close: every rank counts itself in; rank 0 waits for all of them
        ++e->comm_rounds;
        if (e->comm_rank == 0) dwk::comm_signal_wait_kernel<<<1, 1, 0, st>>>(e->uplink_ctl, e->comm_rounds * (unsigned long long)e->comm_world);
        else dwk::comm_signal_kernel<<<1, 1, 0, st>>>(e->uplink_ctl);
        CU(cudaGetLastError());
    } else NC(g_nccl.AllReduce(e->comm_flag.p, e->comm_flag.p + 1, 1, ncclInt, ncclSum, e->comm, st));
    return 0;
}

int dwgpu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

#if DWK_TIMERS
// profiling builds only (not part of include/dwgpu.h): cycles per kernel phase summed over all CTAs since the last reset
int dwgpu_debug_phases(unsigned long long *out16, int reset)
{
    cudaDeviceSynchronize();
    if (out16 && cudaMemcpyFromSymbol(out16, dwk::g_phase, 16 * sizeof(unsigned long long)) != cudaSuccess) return 1;
    if (reset) { unsigned long long z[16] = {0}; if (cudaMemcpyToSymbol(dwk::g_phase, z, sizeof z) != cudaSuccess) return 1; }
    return 0;
}
#endif

// path 5 diagnostics (not part of include/dwgpu.h): [0] blocks moved small -> large heap, [1] -> dense path (heap), [2] -> dense
// path (residual check), [3] heap re-packs, [4] batched sweeps, [5] union-size checks, [6] blocks finished by the probe launch,
// [7] blocks the probe passed on; [8] heap doubles of the small launch, [9] of the large one
extern "C" int dwgpu_debug_sparse_stats(dwgpu_engine *e, unsigned long long *out10, int reset)
{
    if (!e || !out10) return 1;
    for (int i = 0; i < 10; ++i) out10[i] = 0;
    if (e->list_s.empty()) return 0;
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    if (cudaMemcpy(out10, e->sp_stats.p, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
    out10[8] = (unsigned long long)e->sp_H0; out10[9] = (unsigned long long)e->sp_H1;
    if (reset && cudaMemset(e->sp_stats.p, 0, 8 * sizeof(unsigned long long)) != cudaSuccess) return 1;
    return 0;
}

void dwgpu_options_init(dwgpu_options *o)
{
    if (!o) return;
    std::memset(o, 0, sizeof(*o));
    o->device = 0;
    o->clamp_lo_cols = 1;
    o->force_path = 0;
    o->max_iter_mult = 50;
    o->tol_dj = 1e-9;
    o->tol_bnd = 1e-9;
    o->tol_piv = 1e-9;
}

void dwgpu_destroy(dwgpu_engine *e)
{
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    e->T.release(); e->bnd.release(); e->pst.release(); e->a_v.release(); e->dsave.release(); e->cssave.release(); e->dflag.release(); e->rmask.release();
    e->dc_val.release(); e->dr_val.release(); e->c_file.release(); e->c_master.release();
    e->x.release(); e->ws.release(); e->a_rp.release();
    e->a_ci.release(); e->dc_ptr.release(); e->dc_row.release(); e->dr_ptr.release();
    e->dr_col.release(); e->blocks.release(); e->d_list_smem.release();
    e->d_list_gmem.release(); e->d_list_a.release(); e->d_list_g.release(); e->d_list_b.release(); e->d_list_c.release(); e->pi.release(); e->rec.release(); e->colptr.release();
    if (e->uplink_base) { if (e->uplink_owned) cudaFree(e->uplink_base); else cudaIpcCloseMemHandle(e->uplink_base); }
    e->comm_flag.release();
    e->Dt.release(); e->d_list_dense.release(); e->multi_pi.release(); e->multi_cs.release();
    e->hp_src.release(); e->hp_dst.release(); e->hp_len.release();
    if (e->h_mpack) cudaFreeHost(e->h_mpack);
    if (e->comm) dwgpu_comm_release(e);
    if (e->ev_done) cudaEventDestroy(e->ev_done);
    if (e->h_pack) cudaFreeHost(e->h_pack);
    if (e->h_mail) cudaFreeHost(e->h_mail);
    if (e->h_r) cudaFreeHost(e->h_r);
    if (e->h_dbg) cudaFreeHost(e->h_dbg);
    if (e->h_gpack) cudaFreeHost(e->h_gpack);
    e->hist.release();
    e->gcolptr.release();
    e->r_dev.release();
    e->counters.release(); e->last_iters.release();
    e->redo_list.release(); e->redo_count.release(); e->redo_moved.release();
    e->scr_list.release(); e->scr_count.release(); e->scr_key.release();
    e->pi_ref.release(); e->dl_rows.release(); e->dl_meta.release(); e->dl_list.release(); e->dl_pos.release();
    e->d_list_s.release(); e->sp_lists.release(); e->sp_counts.release(); e->spflag.release(); e->sp_moved.release(); e->sp_reb.release(); e->sp_stats.release();
    if (e->h_pi) cudaFreeHost(e->h_pi);
    for (auto &ev : e->ev) if (ev) cudaEventDestroy(ev);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

int dwgpu_create(int K, int L, const dwgpu_block_desc *bd, const dwgpu_options *opt_in, dwgpu_engine **out)
{
    if (!out) return fail(DWGPU_E_ARGS, "out is NULL");
    *out = nullptr;
    if (K < 1 || L < 0 || !bd) return fail(DWGPU_E_ARGS, "bad K/L/blocks");
    dwgpu_options opt;
    if (opt_in) opt = *opt_in; else dwgpu_options_init(&opt);
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev < 1)
        return fail(DWGPU_E_CUDA, std::string("no usable CUDA device: ") + cudaGetErrorString(ce));
    if (opt.device < 0 || opt.device >= ndev) return fail(DWGPU_E_ARGS, "device ordinal out of range");
    CU(cudaSetDevice(opt.device));

    dwgpu_engine *e = new dwgpu_engine();
    e->K = K; e->L = L; e->device = opt.device; e->opt = opt;
    auto bail = [&](int code) { dwgpu_destroy(e); return code; };
#define CUE(call)                                                                             \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            fail(DWGPU_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));           \
            return bail(DWGPU_E_CUDA);                                                        \
        }                                                                                     \
    } while (0)

    const bool timing = std::getenv("DWGPU_DEBUG_TIMING") != nullptr;
    auto stamp = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double ts[8]; int nts = 0;
    ts[nts++] = stamp();
    CUE(cudaDeviceGetAttribute(&e->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, opt.device));
    CUE(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    for (auto &ev : e->ev) CUE(cudaEventCreate(&ev));
    CUE(cudaEventCreateWithFlags(&e->ev_done, cudaEventDisableTiming));

    // ---- sizes and offsets
    e->m.resize(K); e->n.resize(K); e->xoff.resize(K + 1); e->path.resize(K);
    std::vector<long long> t_off(K + 1, 0), r_off(K + 1, 0), v_off(K + 1, 0), a_off(K + 1, 0),
        arp_off(K + 1, 0), d_off(K + 1, 0), dcp_off(K + 1, 0), drp_off(K + 1, 0), ws_off(K + 1, 0), pst_off(K + 1, 0);
    std::vector<int> ldg(K);
    std::vector<long long> rm_off(K + 1, 0);
    e->xoff[0] = 0;
    int n_cluster_blocks = 0;
    const size_t smem_cap = (size_t)e->max_smem_optin - 2048;   // static scratch + reserve
    // path 5 is opt-in (DWGPU_SPARSE=1 or force_path = 5): measured slower than path 3 on config B
    bool sparse_on = false;
    if (const char *sv = std::getenv("DWGPU_SPARSE")) sparse_on = std::atoi(sv) != 0;
    if (opt.force_path == 5) sparse_on = true;
    if (const char *sv = std::getenv("DWGPU_SPARSE_THREADS")) e->sp_nt = std::atoi(sv) == 128 ? 128 : 256;
    if (const char *sv = std::getenv("DWGPU_SPARSE_PROBE")) e->sp_probe = std::atoi(sv) != 0;
    for (int k = 0; k < K; ++k) {
        const dwgpu_block_desc &b = bd[k];
        if (b.m < 0 || b.n < 0 || b.d_nnz < 0 || (b.m > 0 && !b.a_rowptr))
            { fail(DWGPU_E_ARGS, "bad block descriptor"); return bail(DWGPU_E_ARGS); }
        e->m[k] = b.m; e->n[k] = b.n;
        const bool fits = dwk::smem2_bytes(b.m, b.n) <= smem_cap;
        const bool fits_g = dwk::smem2_bytes(b.m, b.n, true) <= smem_cap / 2;   // at least 2 CTAs per SM
        const bool fits_c = dwk::cl_smem_bytes(b.m, b.n) <= smem_cap && b.m >= 1 && b.m < 65535 && b.n < 65535;
        int path = fits ? 1 : (fits_g ? 3 : (fits_c ? 4 : 2));
        // path 5: blocks of the path-3 class whose constraint matrix is sparse enough for the packed tableau to live in
        // shared memory (its fill is watched at run time; a block that outgrows the heaps moves to path 3 for good)
        const int nnz_a = b.m > 0 ? b.a_rowptr[b.m] : 0;
        // (the initial image, rows with their slack, must fill at most half of the small heap)
        const long long sp_fixed = (long long)dwk::sp_solve_bytes(b.m, b.n, e->sp_nt, 0);
        const long long sp_h0 = (((long long)smem_cap - 4096) / 2 - sp_fixed) / (long long)sizeof(double);
        const bool fits_s = fits_g && sparse_on && b.m >= 1 && b.m <= dwk::SP_MAX_M && b.n >= 1 && b.n <= dwk::SP_MAX_N &&
                            2 * (nnz_a + nnz_a / 4 + 4LL * b.m) <= sp_h0;
        if (path == 3 && fits_s) path = 5;
        if (opt.force_path == 5 && fits_s && !fits) path = 5;
        if (opt.force_path == 2) path = 2;
        if (opt.force_path == 3 && fits_g) path = 3;
        if (opt.force_path == 4 && fits_c) path = 4;
        if (opt.force_path == 1 && !fits)
            { fail(DWGPU_E_ARGS, "force_path=1 but a block does not fit shared memory"); return bail(DWGPU_E_ARGS); }
        e->path[k] = path;
        // shared-memory path: the HBM image of T has the shared layout (even ld), one bulk copy
        ldg[k] = path == 1 ? dwk::smem2_ld(b.n) : ((b.n + 3) & ~3);
        if (path == 4) n_cluster_blocks++;
        const int nnz = b.m > 0 ? b.a_rowptr[b.m] : 0;
        long long t_cells = (long long)b.m * ldg[k];
        if (path == 5) {       // the slot holds the dense image (path-3 fall-back) or the packed one with the largest heap
            const long long img = (long long)((dwk::sp_heap_pos(b.m, b.n) + 15) / 8) + (long long)(smem_cap / sizeof(double)) + 2;
            t_cells = std::max(t_cells, img);
            t_cells = (t_cells + 1) & ~1LL;
        }
        t_off[k + 1] = t_off[k] + t_cells;
        pst_off[k + 1] = pst_off[k] + (long long)dwk::pst_bytes(b.m, b.n);
        rm_off[k + 1] = rm_off[k] + ((path == 3 || path == 5) ? (long long)(dwk::rmask_bytes(b.m, b.n) / sizeof(unsigned)) : 0);
        r_off[k + 1] = r_off[k] + b.m;
        e->xoff[k + 1] = e->xoff[k] + b.n;
        v_off[k + 1] = v_off[k] + b.m + b.n;
        a_off[k + 1] = a_off[k] + nnz;
        arp_off[k + 1] = arp_off[k] + b.m + 1;
        // (every block's slices start on a 16-byte boundary and are padded to one, so that they can be moved by bulk copies
        //  or 16-byte loads)
        d_off[k + 1] = d_off[k] + ((b.d_nnz + 3) & ~3);
        dcp_off[k + 1] = dcp_off[k] + ((b.n + 1 + 3) & ~3);
        drp_off[k + 1] = drp_off[k] + L + 1;
        if (path == 1 && b.m == 64 && b.n == 128) e->list_a.push_back(k);
        else if (path == 1) { e->list_smem.push_back(k); e->smem_dyn = std::max(e->smem_dyn, dwk::smem2_bytes(b.m, b.n)); }
        else if (path == 3 && b.m == 256 && b.n == 512) e->list_b.push_back(k);
        else if (path == 3) {
            e->list_g.push_back(k);
            e->smem_dyn_g = std::max(e->smem_dyn_g, dwk::smem2_bytes(b.m, b.n, true));
            e->smem_dyn_w = std::max(e->smem_dyn_w, dwk::warp_smem_bytes(b.m, b.n));
        }
        else if (path == 4) { e->list_c.push_back(k); e->smem_dyn_c = std::max(e->smem_dyn_c, dwk::cl_smem_bytes(b.m, b.n)); }
        else if (path == 5) {
            e->list_s.push_back(k);
            e->smem_dyn_g = std::max(e->smem_dyn_g, dwk::smem2_bytes(b.m, b.n, true));     // the dense fall-back runs the generic path-3 kernel
            e->sp_smem_probe = std::max(e->sp_smem_probe, dwk::sp_probe_bytes(b.m, b.n));
            e->sp_smem0 = std::max(e->sp_smem0, dwk::sp_solve_bytes(b.m, b.n, e->sp_nt, 0));   // without the heap for now
        }
        else e->list_gmem.push_back(k);
    }
    if (!e->list_s.empty()) {
        // heap sizes: the small launch keeps two CTAs per SM, the large one takes all the shared memory of an SM
        const size_t fixed = e->sp_smem0;
        const size_t cap0 = (smem_cap - 4096) / 2, cap1 = smem_cap;
        e->sp_H0 = (int)((cap0 - fixed) / sizeof(double)) & ~1;
        e->sp_H1 = (int)((cap1 - fixed) / sizeof(double)) & ~1;
        // test hooks: smaller heaps force the hand-overs (small heap -> large heap -> dense path 3)
        if (const char *hv = std::getenv("DWGPU_SPARSE_H0")) e->sp_H0 = std::max(64, std::min(e->sp_H0, std::atoi(hv))) & ~1;
        if (const char *hv = std::getenv("DWGPU_SPARSE_H1")) e->sp_H1 = std::max(64, std::min(e->sp_H1, std::atoi(hv))) & ~1;
        e->sp_smem0 = fixed + sizeof(double) * (size_t)e->sp_H0;
        e->sp_smem1 = fixed + sizeof(double) * (size_t)e->sp_H1;
    }
    // cluster path: G CTAs (one per SM) per block, the largest power of two that keeps every block of
    // the shard co-resident (148 SMs / blocks), 16 at most (non-portable size); reserved[0] overrides
    if (n_cluster_blocks > 0) {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, opt.device);
        int G = 1;
        while (G * 2 <= 16 && G * 2 * n_cluster_blocks <= sms) G *= 2;
        if (opt.cluster_size > 0) { G = 1; while (G * 2 <= 16 && G * 2 <= opt.cluster_size) G *= 2; }
        e->cluster_size = G;
    }
    for (int k = 0; k < K; ++k) {
        const dwgpu_block_desc &b = bd[k];
        long long w = 0;
        if (e->path[k] == 2) w = (long long)((dwk::vec_bytes(b.m, b.n) + 15) / 16 * 2);
        if (e->path[k] == 3 || e->path[k] == 5) w = (long long)((b.n + 1) & ~1);      // priced costs cs[n]
        if (e->path[k] == 4) w = (long long)dwk::cl_ws_doubles(b.n, e->cluster_size);
        ws_off[k + 1] = ws_off[k] + w;
    }
    e->ntot = e->xoff[K];

    ts[nts++] = stamp();
    // ---- host staging of the static data
    // (uninitialised buffers: a zero-filling std::vector would touch every page from this one thread first — 100 MB at
    //  config B — before the staging team gets to write it)
    HostBuf<double> h_bnd(2 * v_off[K]), h_av(a_off[K]), h_dcv(d_off[K]), h_drv(d_off[K]), h_cf(e->ntot), h_cm(e->ntot);
    HostBuf<int> h_arp(arp_off[K]), h_aci(a_off[K]), h_dcp(dcp_off[K]), h_dcr(d_off[K]), h_drp(drp_off[K]), h_drc(d_off[K]);
    if (!h_bnd.ok() || !h_av.ok() || !h_dcv.ok() || !h_drv.ok() || !h_cf.ok() || !h_cm.ok() || !h_arp.ok() || !h_aci.ok() ||
        !h_dcp.ok() || !h_dcr.ok() || !h_drp.ok() || !h_drc.ok()) { fail(DWGPU_E_NOMEM, "host staging buffers"); return bail(DWGPU_E_NOMEM); }
    // (one block's slices are written by one thread: the K blocks are staged by a small team of host threads — at config B
    //  this loop was 0.16 s of a 0.22 s engine creation, itself 40 % of a 0.56 s DW solve)
    std::atomic<int> stage_err{0};
    auto stage_block = [&](int k) {
        const dwgpu_block_desc &b = bd[k];
        const int m = b.m, n = b.n;
        double *lo = h_bnd.data() + 2 * v_off[k], *up = lo + (m + n);
        for (int i = 0; i < m; ++i) { lo[i] = norm_bound(b.row_lb[i], false); up[i] = norm_bound(b.row_ub[i], true); }
        for (int j = 0; j < n; ++j) {
            double l = norm_bound(b.col_lb[j], false), u = norm_bound(b.col_ub[j], true);
            // dw_subprob.c:116-120 — GLP_LO column (finite lb, no ub) becomes GLP_DB [lb, 3000]
            if (opt.clamp_lo_cols && std::isfinite(l) && std::isinf(u) && u > 0) u = 3000.0;
            lo[m + j] = l; up[m + j] = u;
            h_cf[e->xoff[k] + j] = b.c_file ? b.c_file[j] : 0.0;
            h_cm[e->xoff[k] + j] = b.c_master ? b.c_master[j] : 0.0;
        }
        const int nnz = m > 0 ? b.a_rowptr[m] : 0;
        for (int i = 0; i <= m; ++i) h_arp[arp_off[k] + i] = m > 0 ? b.a_rowptr[i] : 0;
        for (int p = 0; p < nnz; ++p) {
            if (b.a_col[p] < 0 || b.a_col[p] >= n) { stage_err = 1; return; }
            h_aci[a_off[k] + p] = b.a_col[p];
            h_av[a_off[k] + p] = b.a_val[p];
        }
        // D_k by column and by row, both keeping the reference's COO order inside a bucket
        int *cp = h_dcp.data() + dcp_off[k], *rp = h_drp.data() + drp_off[k];
        std::fill(cp, cp + n + 1, 0); std::fill(rp, rp + L + 1, 0);
        for (int t = 0; t < b.d_nnz; ++t) {
            if (b.d_row[t] < 0 || b.d_row[t] >= L || b.d_col[t] < 0 || b.d_col[t] >= n) { stage_err = 2; return; }
            cp[b.d_col[t] + 1]++; rp[b.d_row[t] + 1]++;
        }
        for (int j = 0; j < n; ++j) cp[j + 1] += cp[j];
        for (int i = 0; i < L; ++i) rp[i + 1] += rp[i];
        std::vector<int> cpos(cp, cp + n), rpos(rp, rp + L);
        for (int t = 0; t < b.d_nnz; ++t) {
            int a = cpos[b.d_col[t]]++, r = rpos[b.d_row[t]]++;
            h_dcr[d_off[k] + a] = b.d_row[t]; h_dcv[d_off[k] + a] = b.d_val[t];
            h_drc[d_off[k] + r] = b.d_col[t]; h_drv[d_off[k] + r] = b.d_val[t];
        }
    };
    {
        const int nthreads = (int)std::max(1u, std::min(16u, std::min(std::thread::hardware_concurrency(), (unsigned)((K + 255) / 256))));
        std::atomic<int> next{0};
        auto worker = [&]() { for (int k; (k = next.fetch_add(1)) < K && !stage_err;) stage_block(k); };
        std::vector<std::thread> team;
        try { for (int t = 1; t < nthreads; ++t) team.emplace_back(worker); } catch (...) {}
        worker();
        for (auto &th : team) th.join();
    }
    if (stage_err) { fail(DWGPU_E_ARGS, stage_err == 1 ? "A column index out of range" : "D index out of range"); return bail(DWGPU_E_ARGS); }

    ts[nts++] = stamp();
    // ---- device arenas
    CUE(e->T.alloc(t_off[K]));
    // Test hook: fill the tableau arena with NaN bit patterns before the first initialisation.  Path 3 never zeroes its
    // tableaux (cells outside a row's item mask are never read); a kernel that did read one would now produce NaNs.
    if (std::getenv("DWGPU_POISON_TABLEAU")) CUE(cudaMemset(e->T.p, 0xFF, sizeof(double) * (size_t)t_off[K]));
    CUE(e->pst.alloc(pst_off[K]));
    CUE(cudaMemset(e->pst.p, 0, std::max<long long>(pst_off[K], 1)));
    CUE(e->bnd.upload_buf(h_bnd));
    CUE(e->a_rp.upload_buf(h_arp)); CUE(e->a_ci.upload_buf(h_aci)); CUE(e->a_v.upload_buf(h_av));
    CUE(e->dc_ptr.upload_buf(h_dcp)); CUE(e->dc_row.upload_buf(h_dcr)); CUE(e->dc_val.upload_buf(h_dcv));
    CUE(e->dr_ptr.upload_buf(h_drp)); CUE(e->dr_col.upload_buf(h_drc)); CUE(e->dr_val.upload_buf(h_drv));
    CUE(e->c_file.upload_buf(h_cf)); CUE(e->c_master.upload_buf(h_cm));
    CUE(e->x.alloc(e->ntot));
    CUE(e->dsave.alloc(e->ntot)); CUE(e->cssave.alloc(e->ntot)); CUE(e->dflag.alloc(K));
    CUE(cudaMemset(e->dflag.p, 0, sizeof(int) * (size_t)K));
    CUE(e->ws.alloc(ws_off[K]));
    CUE(e->rmask.alloc((size_t)rm_off[K]));
    CUE(cudaMemset(e->rmask.p, 0, sizeof(unsigned) * std::max<size_t>((size_t)rm_off[K], 1)));
    if (!e->list_s.empty()) {
        CUE(e->spflag.alloc(K)); CUE(cudaMemset(e->spflag.p, 0, sizeof(int) * (size_t)K));
        CUE(e->sp_lists.alloc(3 * (size_t)K)); CUE(e->sp_counts.alloc(4));
        CUE(cudaMemset(e->sp_counts.p, 0, sizeof(int) * 4));
        CUE(e->sp_moved.alloc(K)); CUE(e->sp_reb.alloc(K));
        CUE(cudaMemset(e->sp_moved.p, 0, (size_t)K)); CUE(cudaMemset(e->sp_reb.p, 0, (size_t)K));
        CUE(e->sp_stats.alloc(8)); CUE(cudaMemset(e->sp_stats.p, 0, sizeof(unsigned long long) * 8));
    }
    CUE(e->pi.alloc(L));
    ts[nts++] = stamp();
    {
        const size_t KL = (size_t)K * L;
        const size_t nd = 2 * (size_t)K + KL, ni = 2 * (size_t)K + KL;
        e->rec_bytes = (sizeof(double) * nd + sizeof(int) * ni + 15) & ~(size_t)15;   // records concatenate 16-byte aligned
        CUE(e->rec.alloc(e->rec_bytes));
        CUE(cudaMemset(e->rec.p, 0, e->rec_bytes));
        double *dp = (double *)e->rec.p;
        e->obj.p = dp; e->cost.p = dp + K; e->val.p = dp + 2 * (size_t)K;
        int *ip = (int *)(dp + nd);
        e->status.p = ip; e->len.p = ip + K; e->ind.p = ip + 2 * (size_t)K;
        CUE(e->colptr.alloc((size_t)K + 1));
        // host mirror: obj K, cost K, val nnz<=KL (doubles) | status K, len K, colptr K+1, ind nnz<=KL (ints)
        // host mirror: [nnz (8 B) | obj K | cost K | val KL] doubles, then [status K | colptr K+1 | ind KL] ints
        e->pack_bytes = sizeof(double) * (1 + 2 * (size_t)K + KL) + sizeof(int) * (2 * (size_t)K + 1 + KL) + 64;
        CUE(cudaHostAlloc((void **)&e->h_pack, e->pack_bytes, cudaHostAllocMapped));
        CUE(cudaHostGetDevicePointer((void **)&e->d_pack, e->h_pack, 0));
        std::memset(e->h_pack, 0, e->pack_bytes);
        auto carve_pack = [&](char *base, dwk::PackDst &d) {
            d.nnz = (long long *)base;
            double *dd = (double *)(base + sizeof(double));
            d.obj = dd; d.cost = dd + K; d.val = dd + 2 * (size_t)K;
            int *ii = (int *)(dd + 2 * (size_t)K + KL);
            d.status = ii; d.colptr = ii + K; d.ind = ii + 2 * (size_t)K + 1;
        };
        carve_pack(e->d_pack, e->pk_dev);
        carve_pack(e->h_pack, e->pk_host);
        if (std::getenv("DWGPU_DEBUG")) {
            CUE(cudaHostAlloc((void **)&e->h_dbg, sizeof(long long) * 32 * (size_t)K, cudaHostAllocMapped));
            CUE(cudaHostGetDevicePointer((void **)&e->d_dbg, e->h_dbg, 0));
            std::memset(e->h_dbg, 0, sizeof(long long) * 32 * (size_t)K);
        }
        CUE(e->r_dev.alloc(K));
        CUE(cudaMallocHost((void **)&e->h_r, sizeof(double) * (size_t)K));
    }
    CUE(e->counters.alloc((size_t)K * 8)); CUE(e->last_iters.alloc(K));
    CUE(cudaMemset(e->counters.p, 0, sizeof(unsigned long long) * (size_t)K * 8));
    CUE(cudaMemset(e->x.p, 0, sizeof(double) * std::max<long long>(e->ntot, 1)));
    CUE(cudaMallocHost((void **)&e->h_pi, sizeof(double) * std::max(L, 1)));

    // blocks whose D_k is dense enough (>= 1/8 of L x n) to price several dual vectors with one FP64 tensor-core GEMM;
    // only the cluster path (large blocks) uses it — a network block's D_k has a handful of entries per column
    std::vector<long long> dt_off(K, -1);
    {
        long long dt_tot = 0;
        for (int k = 0; k < K; ++k)
            if (e->path[k] == 4 && L > 0 && (long long)bd[k].d_nnz * 8 >= (long long)L * bd[k].n) {
                dt_off[k] = dt_tot; dt_tot += (long long)L * bd[k].n; e->list_dense.push_back(k);
            }
        if (dt_tot > 0) { CUE(e->Dt.alloc((size_t)dt_tot)); CUE(e->d_list_dense.upload(e->list_dense)); }
    }
    ts[nts++] = stamp();
    // delta rounds (delta_tg.cuh): which path-3 lists take part, and the variable-position map of their blocks
    std::vector<long long> pos_off(K, -1);
    {
        if (const char *sv = std::getenv("DWGPU_DELTA")) e->delta = std::atoi(sv) != 0;
        if (const char *sv = std::getenv("DWGPU_DELTA_MIN_BLOCKS")) e->delta_min_blocks = std::max(0, std::atoi(sv));
        if (const char *tg = std::getenv("DWGPU_TG_THREADS")) if (std::atoi(tg) == 32) e->delta = false;   // the warp kernel keeps no map
        auto fits = [&](const std::vector<int> &l) {
            if (l.empty()) return false;
            for (int k : l) if (e->n[k] > 512 || e->m[k] < 1) return false;       // 4 items per lane, 4 mask words per row
            return true;
        };
        if (e->delta && L > 0) { e->delta_ok_b = fits(e->list_b); e->delta_ok_g = fits(e->list_g); }
        long long tot = 0;
        if (e->delta_ok_b) for (int k : e->list_b) { pos_off[k] = tot; tot += e->m[k] + e->n[k]; }
        if (e->delta_ok_g) for (int k : e->list_g) { pos_off[k] = tot; tot += e->m[k] + e->n[k]; }
        if (tot > 0) {
            CUE(e->dl_pos.alloc((size_t)tot));
            CUE(e->pi_ref.alloc(L)); CUE(cudaMemset(e->pi_ref.p, 0, sizeof(double) * (size_t)L));
            CUE(e->dl_rows.alloc(dwk::DELTA_MAX_ROWS)); CUE(e->dl_meta.alloc(4)); CUE(e->dl_list.alloc(2 * (size_t)K));
            CUE(cudaMemset(e->dl_rows.p, 0, sizeof(int) * dwk::DELTA_MAX_ROWS));
            CUE(cudaMemset(e->dl_meta.p, 0, sizeof(int) * 4));
        }
    }
    std::vector<dwk::BlockDev> hb(K);
    for (int k = 0; k < K; ++k) {
        dwk::BlockDev &B = hb[k];
        std::memset(&B, 0, sizeof(B));
        B.m = e->m[k]; B.n = e->n[k]; B.L = L; B.ldg = ldg[k];
        B.T = e->T.p + t_off[k];
        char *ps = e->pst.p + pst_off[k];        // beta[m] | head[m] | nbv[n] | vstat[m+n]
        B.beta = (double *)ps; B.head = (int *)(ps + sizeof(double) * (size_t)B.m);
        B.nbv = B.head + B.m; B.vstat = (signed char *)(B.nbv + B.n);
        B.lo = e->bnd.p + 2 * v_off[k]; B.up = B.lo + (B.m + B.n);
        B.a_rp = e->a_rp.p + arp_off[k]; B.a_ci = e->a_ci.p + a_off[k]; B.a_v = e->a_v.p + a_off[k];
        B.dc_ptr = e->dc_ptr.p + dcp_off[k]; B.dc_row = e->dc_row.p + d_off[k]; B.dc_val = e->dc_val.p + d_off[k];
        B.dr_ptr = e->dr_ptr.p + drp_off[k]; B.dr_col = e->dr_col.p + d_off[k]; B.dr_val = e->dr_val.p + d_off[k];
        B.c_file = e->c_file.p + e->xoff[k]; B.c_master = e->c_master.p + e->xoff[k];
        B.x = e->x.p + e->xoff[k];
        B.dsave = e->dsave.p + e->xoff[k]; B.cssave = e->cssave.p + e->xoff[k]; B.dflag = e->dflag.p + k;
        B.ws = e->ws.p + ws_off[k];
        B.rmask = (e->path[k] == 3 || e->path[k] == 5) ? e->rmask.p + rm_off[k] : nullptr;
        B.sparse = e->path[k] == 5 ? 1 : 0;
        B.spflag = e->path[k] == 5 ? e->spflag.p + k : nullptr;
        B.rwd = dwk::rmask_words(B.n);
        B.dnnz = bd[k].d_nnz;
        B.pos = pos_off[k] >= 0 ? e->dl_pos.p + pos_off[k] : nullptr;
        B.xoff = e->xoff[k];
        B.Dt = dt_off[k] >= 0 ? e->Dt.p + dt_off[k] : nullptr;
    }
    CUE(e->blocks.upload(hb));
    CUE(e->d_list_smem.upload(e->list_smem));
    CUE(e->d_list_gmem.upload(e->list_gmem));
    CUE(e->d_list_a.upload(e->list_a));
    CUE(e->d_list_g.upload(e->list_g));
    CUE(e->d_list_b.upload(e->list_b));
    CUE(e->d_list_c.upload(e->list_c));
    CUE(e->d_list_s.upload(e->list_s));

    if (!e->list_s.empty()) {
        CUE(allow_dynamic_smem(e->device, (const void *)dwk::probe_sparse_kernel<128>, (int)e->sp_smem_probe));
        CUE(allow_dynamic_smem(e->device, (const void *)dwk::solve_sparse_kernel<128>, (int)e->sp_smem1));
        CUE(allow_dynamic_smem(e->device, (const void *)dwk::solve_sparse_kernel<256>, (int)e->sp_smem1));
    }
    if (!e->list_smem.empty())
        CUE(allow_dynamic_smem(e->device, (const void *)dwk::solve_smem_kernel<NT_SMEM, 0, 0, false>, (int)e->smem_dyn));
    if (!e->list_a.empty())
        CUE(allow_dynamic_smem(e->device, (const void *)dwk::solve_smem_kernel<NT_SMEM, 64, 128, false>, (int)dwk::smem2_bytes(64, 128)));
    if (!e->list_c.empty()) {
        CUE(allow_dynamic_smem(e->device, (const void *)dwk::solve_cluster_kernel<dwk::NT_CL>, (int)e->smem_dyn_c));
        if (e->cluster_size > 8)
            CUE(cudaFuncSetAttribute(dwk::solve_cluster_kernel<dwk::NT_CL>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        // the hardware may not be able to co-schedule a cluster this large with this much shared memory
        // (GPC size): halve until at least one fits
        while (e->cluster_size > 1) {
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3((unsigned)(e->list_c.size() * e->cluster_size)); cfg.blockDim = dim3(dwk::NT_CL);
            cfg.dynamicSmemBytes = e->smem_dyn_c;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = (unsigned)e->cluster_size; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int ncl = 0;
            cudaError_t oe = cudaOccupancyMaxActiveClusters(&ncl, dwk::solve_cluster_kernel<dwk::NT_CL>, &cfg);
            if (oe == cudaSuccess && ncl >= 1) break;
            cudaGetLastError();
            e->cluster_size /= 2;
        }
    }
    CUE(allow_dynamic_smem(e->device, (const void *)dwk::order_by_last_iterations_kernel<1024, 8192>, (int)(8192 * sizeof(unsigned long long))));
    if (const char *nr = std::getenv("DWGPU_NO_REORDER")) e->reorder = std::atoi(nr) == 0;
    if (const char *tg = std::getenv("DWGPU_TG_THREADS")) {
        const int v = std::atoi(tg);
        e->tg_threads = (v == 256 || v == NT_TG_WIDE || v == NT_TG || v == 32) ? v : 0;
    }
    if (const char *ba = std::getenv("DWGPU_WARP_BAIL_AFTER")) e->warp_bail_after = std::max(0, std::atoi(ba));
    if (const char *sv = std::getenv("DWGPU_SCREEN")) e->screen = std::atoi(sv) != 0;
    if (const char *sv = std::getenv("DWGPU_SCREEN_SORT")) e->screen_sort = std::atoi(sv) != 0;
    if (e->screen && (!e->list_g.empty() || !e->list_b.empty())) {
        // the screening warp keeps a block's reduced costs in registers (4 items per lane: n <= 512) and 16-bit row ids
        auto fits = [&](const std::vector<int> &l, int &wb) {
            if (l.empty()) return false;
            size_t mx = 0;
            for (int k : l) {
                if (e->n[k] > 512 || e->m[k] >= 65535) return false;
                if ((e->n[k] & 3) || (e->xoff[k] & 1)) return false;       // 16-byte stores into the saved row
                mx = std::max(mx, dwk::screen_warp_bytes(e->m[k], e->n[k]));
            }
            wb = (int)mx;
            return mx * SCREEN_NW <= (size_t)smem_cap / 2;
        };
        e->screen_ok_b = fits(e->list_b, e->screen_wb_b);
        e->screen_ok_g = fits(e->list_g, e->screen_wb_g);
        if (e->screen_ok_b || e->screen_ok_g) {
            CUE(e->scr_list.alloc(2 * (size_t)K)); CUE(e->scr_count.alloc(2)); CUE(e->scr_key.alloc(K));
            CUE(cudaMemset(e->scr_count.p, 0, 2 * sizeof(int)));
            CUE(cudaMemset(e->scr_key.p, 0, sizeof(unsigned) * (size_t)K));
            const int wb = std::max(e->screen_wb_b, e->screen_wb_g) * SCREEN_NW;
            CUE(allow_dynamic_smem(e->device, (const void *)dwk::screen_tg_kernel<SCREEN_NW, NT_TG>, wb));
            CUE(allow_dynamic_smem(e->device, (const void *)dwk::screen_tg_kernel<SCREEN_NW, NT_TG_WIDE>, wb));
            CUE(allow_dynamic_smem(e->device, (const void *)dwk::screen_tg_kernel<SCREEN_NW, NT_SMEM>, wb));
            CUE(allow_dynamic_smem(e->device, (const void *)dwk::order_screened_kernel<1024>, (int)(8192 * sizeof(unsigned long long))));
        }
    }
    // warp kernel: 16-bit variable ids, at least 4 blocks per SM
    e->warp_ok_g = !e->list_g.empty() && e->smem_dyn_w <= smem_cap / 4;
    for (int k : e->list_g) if (e->m[k] + e->n[k] >= 65535) e->warp_ok_g = false;
    if (!e->list_g.empty() || !e->list_b.empty()) {
        CUE(e->redo_list.alloc(K)); CUE(e->redo_count.alloc(1)); CUE(e->redo_moved.alloc(K));
        CUE(cudaMemset(e->redo_count.p, 0, sizeof(int)));
        CUE(cudaMemset(e->redo_moved.p, 0, (size_t)K));
        if (e->warp_ok_g)
            CUE(allow_dynamic_smem(e->device, (const void *)dwk::solve_warp_kernel<0, 0>, (int)e->smem_dyn_w));
        if (!e->list_b.empty())
            CUE(allow_dynamic_smem(e->device, (const void *)dwk::solve_warp_kernel<256, 512>, (int)dwk::warp_smem_bytes(256, 512)));
    }
    if (!e->list_g.empty() || !e->list_s.empty()) {
        CUE(allow_dynamic_smem(e->device, (const void *)dwk::solve_smem_kernel<NT_SMEM, 0, 0, true>, (int)e->smem_dyn_g));
        CUE(allow_dynamic_smem(e->device, (const void *)dwk::solve_smem_kernel<NT_TG, 0, 0, true>, (int)e->smem_dyn_g));
        CUE(allow_dynamic_smem(e->device, (const void *)dwk::solve_smem_kernel<NT_TG_WIDE, 0, 0, true>, (int)e->smem_dyn_g));
    }
    if (!e->list_b.empty()) {
        CUE(allow_dynamic_smem(e->device, (const void *)dwk::solve_smem_kernel<NT_SMEM, 256, 512, true>, (int)dwk::smem2_bytes(256, 512, true)));
        CUE(allow_dynamic_smem(e->device, (const void *)dwk::solve_smem_kernel<NT_TG, 256, 512, true>, (int)dwk::smem2_bytes(256, 512, true)));
        CUE(allow_dynamic_smem(e->device, (const void *)dwk::solve_smem_kernel<NT_TG_WIDE, 256, 512, true>, (int)dwk::smem2_bytes(256, 512, true)));
    }

    ts[nts++] = stamp();
    if (!e->list_dense.empty()) {
        dwk::densify_dt_kernel<<<(unsigned)e->list_dense.size(), 1024, 0, e->stream>>>(e->blocks.p, e->d_list_dense.p, (int)e->list_dense.size());
        CUE(cudaGetLastError());
    }
    dwk::init_kernel<256><<<K, 256, 0, e->stream>>>(e->blocks.p, K);
    CUE(cudaGetLastError());
    CUE(cudaStreamSynchronize(e->stream));
    ts[nts++] = stamp();
    if (timing)
        std::fprintf(stderr, "dwgpu_create: sizes %.4f s, host staging %.4f s, arenas + uploads %.4f s, pinned records %.4f s, descriptors + attributes %.4f s, init kernel %.4f s\n",
                     ts[1] - ts[0], ts[2] - ts[1], ts[3] - ts[2], ts[4] - ts[3], ts[5] - ts[4], ts[6] - ts[5]);
    e->initialised = true;
    *out = e;
    return 0;
#undef CUE
}

static int launch_round(dwgpu_engine *e, const double *d_pi, int phase_one, int initial, cudaStream_t st, bool mail = false,
                        const dwk::MailDst *direct = nullptr)
{
    dwk::SolveParams P;
    if (direct) {         // the kernels' epilogues store into this window (multi-GPU up-link in rank 0's HBM)
        P.m_obj = direct->obj; P.m_cost = direct->cost; P.m_val = direct->val;
        P.m_status = direct->status; P.m_len = direct->len; P.m_ind = direct->ind;
        e->mail_valid = false;
    } else if (mail && e->mail_valid) {
        P.m_obj = e->mail_dev.obj; P.m_cost = e->mail_dev.cost; P.m_val = e->mail_dev.val;
        P.m_status = e->mail_dev.status; P.m_len = e->mail_dev.len; P.m_ind = e->mail_dev.ind;
    } else {
        P.m_obj = P.m_cost = P.m_val = nullptr; P.m_status = P.m_len = P.m_ind = nullptr;
        if (!mail) e->mail_valid = false;
    }
    P.pi = d_pi; P.phase_one = phase_one; P.initial = initial;
    bool delta_round = false;
    P.max_iter_mult = e->opt.max_iter_mult;
    P.no_dual_start = e->opt.no_dual_start;
    // the "nothing moved" exit of the kernels rewrites only obj/status and relies on the previous proposal of
    // the block still standing in the result arrays; after create / reset there is none until a round has written it
    P.cs_pre = e->cs_pre_now;
    P.force_full = (!initial && !e->have_proposals) ? 1 : 0;
    e->have_proposals = true;
    P.dbg = e->d_dbg;
    P.tol_dj = e->opt.tol_dj; P.tol_bnd = e->opt.tol_bnd; P.tol_piv = e->opt.tol_piv;
    P.obj = e->obj.p; P.cost = e->cost.p; P.status = e->status.p; P.len = e->len.p;
    P.ind = e->ind.p; P.val = e->val.p; P.counters = e->counters.p; P.last_iters = e->last_iters.p;
    e->last_launches = 0;
    // ---- the round's preparation (pack.cuh: prep_job): multi-GPU hand-shake of the duals + dual screening for the delta
    // rounds.  It rides along as the second CTA of the first launch-order kernel of the round, or is its own small launch.
    dwk::PrepJob J;
    std::memset(&J, 0, sizeof(J));
    J.pi = d_pi; J.L = e->L;
    J.ctl = e->uplink_ctl; J.seq = e->prep_seq; J.comm_mode = (initial || d_pi == nullptr) ? 0 : e->prep_comm_mode;
    e->prep_comm_mode = 0;
    if (initial) e->pi_ref_valid = false;
    else if ((e->delta_ok_b || e->delta_ok_g) && e->tg_threads != 32 && d_pi != nullptr) {
        if (P.cs_pre != nullptr) e->pi_ref_valid = false;       // (priced by the multi-vector GEMM with the duals as given)
        else {
            // duals down: list the rows that moved since they were last applied, snap the others; the round prices with pi_ref
            J.force_full = (!e->pi_ref_valid || e->last_phase != phase_one || P.force_full) ? 1 : 0;
            J.pi_ref = e->pi_ref.p; J.rows = e->dl_rows.p; J.meta = (dwk::DeltaMeta *)e->dl_meta.p;
            e->pi_ref_valid = true; e->last_phase = phase_one;
            P.pi = e->pi_ref.p;
            delta_round = true;
        }
    }
    bool prep_pending = J.pi_ref != nullptr || J.comm_mode != 0;
    // longest-processing-time-first inside each list (DWGPU_NO_REORDER=1 keeps the natural order)
    if (!initial && e->reorder) {
        auto order = [&](DevBuf<int> &d, size_t n) -> cudaError_t {
            // pays off whenever a launch is more than one wave of CTAs deep (592 resident 128-thread CTAs): multi-GPU
            // shards of 1024 blocks are 1.7 waves, where the long poles of an unordered launch end up in the tail
            if (n < 600 || n > 8192) return cudaSuccess;
            const unsigned grid = prep_pending ? 2u : 1u;
            if (n <= 1024)
                dwk::order_by_last_iterations_kernel<1024, 1024><<<grid, 1024, 1024 * sizeof(unsigned long long), st>>>(d.p, (int)n, e->last_iters.p, J);
            else if (n <= 2048)
                dwk::order_by_last_iterations_kernel<1024, 2048><<<grid, 1024, 2048 * sizeof(unsigned long long), st>>>(d.p, (int)n, e->last_iters.p, J);
            else if (n <= 4096)
                dwk::order_by_last_iterations_kernel<1024, 4096><<<grid, 1024, 4096 * sizeof(unsigned long long), st>>>(d.p, (int)n, e->last_iters.p, J);
            else
                dwk::order_by_last_iterations_kernel<1024, 8192><<<grid, 1024, 8192 * sizeof(unsigned long long), st>>>(d.p, (int)n, e->last_iters.p, J);
            prep_pending = false;
            e->last_launches++;
            return cudaGetLastError();
        };
        // (lists that go through the screening launch are ordered by it)
        const bool scr_any = !P.force_full && P.cs_pre == nullptr && e->tg_threads != 32;
        if (!(scr_any && e->screen_ok_b && e->list_b.size() <= 8192)) CU(order(e->d_list_b, e->list_b.size()));
        if (!(scr_any && e->screen_ok_g && e->list_g.size() <= 8192)) CU(order(e->d_list_g, e->list_g.size()));
        CU(order(e->d_list_s, e->list_s.size()));
        CU(order(e->d_list_a, e->list_a.size())); CU(order(e->d_list_smem, e->list_smem.size()));
    }
    if (prep_pending) {
        dwk::delta_rows_kernel<<<1, 256, 0, st>>>(J);
        CU(cudaGetLastError());
        e->last_launches++;
    }
    // large blocks first: they are the long poles
    e->ev_used[0] = e->ev_used[1] = false;
    if (!e->list_gmem.empty()) {
        CU(cudaEventRecord(e->ev[0], st));
        dwk::solve_kernel<false, NT_GMEM><<<(unsigned)e->list_gmem.size(), NT_GMEM, 0, st>>>(
            e->blocks.p, e->d_list_gmem.p, (int)e->list_gmem.size(), P);
        CU(cudaGetLastError());
        CU(cudaEventRecord(e->ev[1], st));
        e->ev_used[1] = true;
        e->last_launches++;
    }
    if (!e->list_c.empty()) {
        const bool own_events = !e->ev_used[1];
        if (own_events) CU(cudaEventRecord(e->ev[0], st));
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)(e->list_c.size() * e->cluster_size)); cfg.blockDim = dim3(dwk::NT_CL);
        cfg.dynamicSmemBytes = e->smem_dyn_c; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)e->cluster_size; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CU(cudaLaunchKernelEx(&cfg, dwk::solve_cluster_kernel<dwk::NT_CL>, (const dwk::BlockDev *)e->blocks.p,
                              (const int *)e->d_list_c.p, (int)e->list_c.size(), P, e->cluster_size));
        CU(cudaEventRecord(e->ev[1], st));
        e->ev_used[1] = true;
        e->last_launches++;
    }
    if (!e->list_s.empty()) {
        // path 5: [probe ->] solve (small heap, 2 CTAs per SM) -> solve (large heap) -> dense path-3 kernel; every
        // launch after the first reads the length of its list on the device (normally 0 for the last two)
        const bool own_events = !e->ev_used[1];
        if (own_events) CU(cudaEventRecord(e->ev[0], st));
        const unsigned nb = (unsigned)e->list_s.size();
        const int K_ = e->K;
        int *l0 = e->sp_lists.p, *l1 = l0 + K_, *l2 = l1 + K_;
        int *c0 = e->sp_counts.p, *c1 = c0 + 1, *c2 = c0 + 2;
        CU(cudaMemsetAsync(e->sp_counts.p, 0, sizeof(int) * 4, st));
        const bool probe = e->sp_probe && !initial && !P.force_full;
        const int *solve_list = e->d_list_s.p;
        const int *solve_count = nullptr;
        if (probe) {
            dwk::SparseRound R{nullptr, 0, l0, c0, l2, c2, e->sp_moved.p, e->sp_reb.p, e->sp_stats.p};
            dwk::probe_sparse_kernel<128><<<nb, 128, e->sp_smem_probe, st>>>(e->blocks.p, e->d_list_s.p, (int)nb, P, R);
            CU(cudaGetLastError());
            e->last_launches++;
            solve_list = l0; solve_count = c0;
        }
        auto solve = [&](const int *list, const int *cnt, int use_moved, int *nl, int *nc, int level, int H, size_t sm) {
            dwk::SparseRound R{cnt, use_moved, nl, nc, l2, c2, e->sp_moved.p, e->sp_reb.p, e->sp_stats.p};
            if (e->sp_nt == 128) dwk::solve_sparse_kernel<128><<<nb, 128, sm, st>>>(e->blocks.p, list, (int)nb, P, R, level, H);
            else dwk::solve_sparse_kernel<256><<<nb, 256, sm, st>>>(e->blocks.p, list, (int)nb, P, R, level, H);
            e->last_launches++;
            return cudaGetLastError();
        };
        CU(solve(solve_list, solve_count, 0, l1, c1, 0, e->sp_H0, e->sp_smem0));
        CU(solve(l1, c1, 1, l2, c2, 1, e->sp_H1, e->sp_smem1));
        {
            const dwk::Continuation cn{c2, e->sp_moved.p, e->sp_reb.p, nullptr, nullptr};
            dwk::solve_smem_kernel<NT_TG_WIDE, 0, 0, true><<<nb, NT_TG_WIDE, e->smem_dyn_g, st>>>(e->blocks.p, l2, (int)nb, P, cn);
            CU(cudaGetLastError());
            e->last_launches++;
        }
        CU(cudaEventRecord(e->ev[1], st));
        e->ev_used[1] = true;
    }
    if (!e->list_g.empty() || !e->list_b.empty()) {
        const bool own_events = !e->ev_used[1];
        if (own_events) CU(cudaEventRecord(e->ev[0], st));
        // Kernel choice for the HBM-tableau blocks.  Default: the CTA kernel with 128 threads, 4 per SM.
        // DWGPU_TG_THREADS = 64 | 256 picks another CTA width; DWGPU_TG_THREADS = 32 runs ONE WARP per block
        // (simplex_warp.cuh, 10 blocks per SM resident) followed by the CTA kernel over the blocks the warps handed over
        // (normally none: the second launch reads the list length on the device and its CTAs exit at once).  Measured on
        // config B (profiles/r2a_warp_kernel.md): the warp kernel needs 30 % fewer instructions per basis change but a
        // single warp's dependent chain is 3x longer, so 10 resident blocks do not beat 4 CTAs (40.1 vs 33.1 ms a step).
        const dwk::Continuation none{nullptr, nullptr, nullptr, nullptr, nullptr};
        auto cta_launch = [&](bool spec_b, const int *list, unsigned nb, int nt, const dwk::Continuation &cn) {
            const size_t sm = spec_b ? dwk::smem2_bytes(256, 512, true) : e->smem_dyn_g;
            if (spec_b) {
                if (nt == 256) dwk::solve_smem_kernel<NT_SMEM, 256, 512, true><<<nb, NT_SMEM, sm, st>>>(e->blocks.p, list, (int)nb, P, cn);
                else if (nt == NT_TG) dwk::solve_smem_kernel<NT_TG, 256, 512, true><<<nb, NT_TG, sm, st>>>(e->blocks.p, list, (int)nb, P, cn);
                else dwk::solve_smem_kernel<NT_TG_WIDE, 256, 512, true><<<nb, NT_TG_WIDE, sm, st>>>(e->blocks.p, list, (int)nb, P, cn);
            } else {
                if (nt == 256) dwk::solve_smem_kernel<NT_SMEM, 0, 0, true><<<nb, NT_SMEM, sm, st>>>(e->blocks.p, list, (int)nb, P, cn);
                else if (nt == NT_TG) dwk::solve_smem_kernel<NT_TG, 0, 0, true><<<nb, NT_TG, sm, st>>>(e->blocks.p, list, (int)nb, P, cn);
                else dwk::solve_smem_kernel<NT_TG_WIDE, 0, 0, true><<<nb, NT_TG_WIDE, sm, st>>>(e->blocks.p, list, (int)nb, P, cn);
            }
            e->last_launches++;
            return cudaGetLastError();
        };
        auto run_list = [&](bool spec_b, const int *list, unsigned nb, bool warp_fits) -> cudaError_t {
            const bool warp = warp_fits && e->tg_threads == 32;
            const bool scr = !warp && e->tg_threads != 32 && (spec_b ? e->screen_ok_b : e->screen_ok_g) && !initial &&
                             !P.force_full && P.cs_pre == nullptr && nb <= 8192;
            if (delta_round && !warp && (spec_b ? e->delta_ok_b : e->delta_ok_g) && (int)nb >= e->delta_min_blocks) {
                // one warp per block finishes the blocks no listed dual can move; the CTA kernel takes the rest (or, when the
                // device-side flag says so, its whole static list)
                const dwk::DeltaMeta *meta = (const dwk::DeltaMeta *)e->dl_meta.p;
                int *nl = e->dl_list.p + (spec_b ? 0 : e->K);
                int *nc = e->dl_meta.p + (spec_b ? 2 : 3);
                const dwk::DeltaIn D{e->dl_rows.p, meta, nl, nc};
                constexpr int DNW = 4;
                dwk::delta_round_kernel<DNW><<<(nb + DNW - 1) / DNW, DNW * 32, 0, st>>>(e->blocks.p, list, (int)nb, P, D);
                e->last_launches++;
                cudaError_t rc = cudaGetLastError();
                if (rc != cudaSuccess) return rc;
                const dwk::Continuation cn{nc, nullptr, nullptr, e->dl_meta.p + 1, list};
                return cta_launch(spec_b, nl, nb, e->tg_threads, cn);
            }
            if (scr) {
                // screening launch -> (sort) -> solve launch over the device-side list of the blocks that still need pivots
                int *nl = e->scr_list.p + (spec_b ? 0 : e->K), *nc = e->scr_count.p + (spec_b ? 0 : 1);
                cudaError_t rc = cudaMemsetAsync(nc, 0, sizeof(int), st);
                if (rc != cudaSuccess) return rc;
                const dwk::ScreenOut S{nl, nc, e->scr_key.p};
                const int wb = spec_b ? e->screen_wb_b : e->screen_wb_g;
                const unsigned gb = (nb + SCREEN_NW - 1) / SCREEN_NW;
                const int nt = e->tg_threads ? e->tg_threads : NT_TG_WIDE;
                if (nt == NT_SMEM) dwk::screen_tg_kernel<SCREEN_NW, NT_SMEM><<<gb, SCREEN_NW * 32, (size_t)wb * SCREEN_NW, st>>>(e->blocks.p, list, (int)nb, P, S, wb);
                else if (nt == NT_TG) dwk::screen_tg_kernel<SCREEN_NW, NT_TG><<<gb, SCREEN_NW * 32, (size_t)wb * SCREEN_NW, st>>>(e->blocks.p, list, (int)nb, P, S, wb);
                else dwk::screen_tg_kernel<SCREEN_NW, NT_TG_WIDE><<<gb, SCREEN_NW * 32, (size_t)wb * SCREEN_NW, st>>>(e->blocks.p, list, (int)nb, P, S, wb);
                e->last_launches++;
                rc = cudaGetLastError();
                if (rc != cudaSuccess) return rc;
                if (e->screen_sort) {
                    dwk::order_screened_kernel<1024><<<1, 1024, 8192 * sizeof(unsigned long long), st>>>(nl, nc, e->scr_key.p, 8192);
                    e->last_launches++;
                    rc = cudaGetLastError();
                    if (rc != cudaSuccess) return rc;
                }
                const dwk::Continuation cn{nc, nullptr, nullptr, nullptr, nullptr};
                return cta_launch(spec_b, nl, nb, e->tg_threads, cn);
            }
            if (!warp) return cta_launch(spec_b, list, nb, e->tg_threads == 32 ? 0 : e->tg_threads, none);
            cudaError_t rc = cudaMemsetAsync(e->redo_count.p, 0, sizeof(int), st);
            if (rc != cudaSuccess) return rc;
            dwk::WarpExtra X{e->redo_list.p, e->redo_count.p, e->redo_moved.p, e->warp_bail_after};
            if (spec_b)
                dwk::solve_warp_kernel<256, 512><<<nb, 32, dwk::warp_smem_bytes(256, 512), st>>>(e->blocks.p, list, (int)nb, P, X);
            else
                dwk::solve_warp_kernel<0, 0><<<nb, 32, e->smem_dyn_w, st>>>(e->blocks.p, list, (int)nb, P, X);
            e->last_launches++;
            rc = cudaGetLastError();
            if (rc != cudaSuccess) return rc;
            const dwk::Continuation cn{e->redo_count.p, e->redo_moved.p, nullptr, nullptr, nullptr};
            return cta_launch(spec_b, e->redo_list.p, nb, 0, cn);
        };
        if (!e->list_b.empty()) CU(run_list(true, e->d_list_b.p, (unsigned)e->list_b.size(), true));
        if (!e->list_g.empty()) CU(run_list(false, e->d_list_g.p, (unsigned)e->list_g.size(), e->warp_ok_g));
        CU(cudaEventRecord(e->ev[1], st));
        e->ev_used[1] = true;
    }
    if (!e->list_smem.empty() || !e->list_a.empty()) {
        CU(cudaEventRecord(e->ev[2], st));
        if (!e->list_a.empty()) {
            dwk::solve_smem_kernel<NT_SMEM, 64, 128, false><<<(unsigned)e->list_a.size(), NT_SMEM, dwk::smem2_bytes(64, 128), st>>>(
                e->blocks.p, e->d_list_a.p, (int)e->list_a.size(), P, dwk::Continuation{nullptr, nullptr, nullptr, nullptr, nullptr});
            CU(cudaGetLastError());
            e->last_launches++;
        }
        if (!e->list_smem.empty()) {
            dwk::solve_smem_kernel<NT_SMEM, 0, 0, false><<<(unsigned)e->list_smem.size(), NT_SMEM, e->smem_dyn, st>>>(
                e->blocks.p, e->d_list_smem.p, (int)e->list_smem.size(), P, dwk::Continuation{nullptr, nullptr, nullptr, nullptr, nullptr});
            CU(cudaGetLastError());
            e->last_launches++;
        }
        CU(cudaEventRecord(e->ev[3], st));
        e->ev_used[0] = true;
    }
    e->solves += e->K;
    if (st != e->stream && e->ev_done) { CU(cudaEventRecord(e->ev_done, st)); e->ev_done_pending = true; }
    return 0;
}

// orders e->stream behind the last round / reset that ran on another stream
static cudaError_t after_last_round(dwgpu_engine *e)
{
    if (!e->ev_done_pending) return cudaSuccess;
    e->ev_done_pending = false;
    return cudaStreamWaitEvent(e->stream, e->ev_done, 0);
}

int dwgpu_fetch_results(dwgpu_engine *e, dwgpu_proposals *o)
{
    if (!e) return fail(DWGPU_E_ARGS, "engine is NULL");
    CU(cudaSetDevice(e->device));
    CU(after_last_round(e));
    cudaStream_t st = e->stream;
    if (e->h_dbg) {
        cudaError_t se = cudaStreamSynchronize(st);
        if (se != cudaSuccess) {
            for (int k = 0; k < e->K && k < 16; ++k) {
                std::fprintf(stderr, "dwgpu debug block %d:", k);
                for (int t = 0; t < 32; ++t) std::fprintf(stderr, " %lld", e->h_dbg[32 * (size_t)k + t]);
                std::fprintf(stderr, "\n");
            }
        }
        CU(se);
    }
    if (o) {
        const size_t K = e->K, KL = (size_t)e->K * e->L;
        if (o->obj) CU(cudaMemcpyAsync(o->obj, e->obj.p, sizeof(double) * K, cudaMemcpyDeviceToHost, st));
        if (o->cost) CU(cudaMemcpyAsync(o->cost, e->cost.p, sizeof(double) * K, cudaMemcpyDeviceToHost, st));
        if (o->status) CU(cudaMemcpyAsync(o->status, e->status.p, sizeof(int) * K, cudaMemcpyDeviceToHost, st));
        if (o->len) CU(cudaMemcpyAsync(o->len, e->len.p, sizeof(int) * K, cudaMemcpyDeviceToHost, st));
        if (o->ind && KL) CU(cudaMemcpyAsync(o->ind, e->ind.p, sizeof(int) * KL, cudaMemcpyDeviceToHost, st));
        if (o->val && KL) CU(cudaMemcpyAsync(o->val, e->val.p, sizeof(double) * KL, cudaMemcpyDeviceToHost, st));
        if (o->x && e->ntot) CU(cudaMemcpyAsync(o->x, e->x.p, sizeof(double) * e->ntot, cudaMemcpyDeviceToHost, st));
    }
    CU(cudaStreamSynchronize(st));
    return 0;
}

// scan + copy of the last round's proposals into the mapped pinned mirror, then one synchronisation
static int pack_round(dwgpu_engine *e, const double *r, dwgpu_packed *out)
{
    if (!out) return fail(DWGPU_E_ARGS, "out is NULL");
    cudaStream_t st = e->stream;
    CU(after_last_round(e));
    dwk::PackSrc s;
    s.obj = e->obj.p; s.cost = e->cost.p; s.val = e->val.p;
    s.status = e->status.p; s.len = e->len.p; s.ind = e->ind.p;
    s.r = nullptr; s.K = e->K; s.L = e->L; s.accept_tol = 1e-6;   // TOLERANCE, src/dw.h:71
    s.recs = nullptr; s.rec_bytes = 0; s.seg_K = 0;
    if (r) {
        std::memcpy(e->h_r, r, sizeof(double) * (size_t)e->K);
        CU(cudaMemcpyAsync(e->r_dev.p, e->h_r, sizeof(double) * (size_t)e->K, cudaMemcpyHostToDevice, st));
        s.r = e->r_dev.p;
    }
    dwk::pack_scan_kernel<1024><<<1, 1024, 0, st>>>(s, e->colptr.p, e->pk_dev);
    CU(cudaGetLastError());
    dwk::pack_copy_kernel<256><<<(unsigned)((e->K + 7) / 8), 256, 0, st>>>(s, e->colptr.p, e->pk_dev);
    CU(cudaGetLastError());
    e->last_launches += 2;
    CU(cudaStreamSynchronize(st));
    out->obj = e->pk_host.obj; out->cost = e->pk_host.cost; out->status = e->pk_host.status;
    out->colptr = e->pk_host.colptr; out->ind = e->pk_host.ind; out->val = e->pk_host.val;
    out->nnz = e->pk_host.nnz[0];
    return 0;
}

static int pack_gathered_impl(dwgpu_engine *e, int K_total, const dwgpu_proposals *dev, const void *recs, int world,
                              const double *r_dev, void *stream, dwgpu_packed *out);

int dwgpu_pack_gathered(dwgpu_engine *e, int K_total, const dwgpu_proposals *dev, const double *r_dev, void *stream,
                        dwgpu_packed *out)
{
    if (!e || !dev || !out || K_total < 1) return fail(DWGPU_E_ARGS, "bad arguments");
    if (!dev->obj || !dev->cost || !dev->status || !dev->len || (e->L > 0 && (!dev->ind || !dev->val)))
        return fail(DWGPU_E_ARGS, "dev_arrays must hold obj, cost, status, len, ind, val");
    return pack_gathered_impl(e, K_total, dev, nullptr, 0, r_dev, stream, out);
}

int dwgpu_device_record(dwgpu_engine *e, void **dev_ptr, size_t *bytes)
{
    if (!e || !dev_ptr || !bytes) return fail(DWGPU_E_ARGS, "NULL argument");
    *dev_ptr = e->rec.p; *bytes = e->rec_bytes;
    return 0;
}

int dwgpu_pack_gathered_records(dwgpu_engine *e, int world, const void *records, const double *r_dev, void *stream,
                                dwgpu_packed *out)
{
    if (!e || !records || !out || world < 1) return fail(DWGPU_E_ARGS, "bad arguments");
    return pack_gathered_impl(e, world * e->K, nullptr, records, world, r_dev, stream, out);
}

static int pack_gathered_impl(dwgpu_engine *e, int K_total, const dwgpu_proposals *dev, const void *recs, int world,
                              const double *r_dev, void *stream, dwgpu_packed *out)
{
    CU(cudaSetDevice(e->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
    const size_t K = (size_t)K_total, KL = K * (size_t)e->L;
    if (e->gpack_K < K_total) {
        if (e->h_gpack) { CU(cudaStreamSynchronize(st)); cudaFreeHost(e->h_gpack); e->h_gpack = nullptr; }
        e->gcolptr.release();
        const size_t bytes = sizeof(double) * (1 + 2 * K + KL) + sizeof(int) * (2 * K + 1 + KL) + 64;
        CU(cudaHostAlloc((void **)&e->h_gpack, bytes, cudaHostAllocMapped));
        CU(cudaHostGetDevicePointer((void **)&e->d_gpack, e->h_gpack, 0));
        std::memset(e->h_gpack, 0, bytes);
        auto carve = [&](char *base, dwk::PackDst &d) {
            d.nnz = (long long *)base;
            double *dd = (double *)(base + sizeof(double));
            d.obj = dd; d.cost = dd + K; d.val = dd + 2 * K;
            int *ii = (int *)(dd + 2 * K + KL);
            d.status = ii; d.colptr = ii + K; d.ind = ii + 2 * K + 1;
        };
        carve(e->d_gpack, e->gpk_dev);
        carve(e->h_gpack, e->gpk_host);
        CU(e->gcolptr.alloc(K + 1));
        e->gpack_K = K_total;
    }
    dwk::PackSrc s;
    std::memset(&s, 0, sizeof(s));
    if (dev) { s.obj = dev->obj; s.cost = dev->cost; s.val = dev->val; s.status = dev->status; s.len = dev->len; s.ind = dev->ind; }
    else { s.recs = (const char *)recs; s.rec_bytes = e->rec_bytes; s.seg_K = e->K; (void)world; }
    s.r = r_dev; s.K = K_total; s.L = e->L; s.accept_tol = 1e-6;
    dwk::pack_scan_kernel<1024><<<1, 1024, 0, st>>>(s, e->gcolptr.p, e->gpk_dev);
    CU(cudaGetLastError());
    dwk::pack_copy_kernel<256><<<(unsigned)((K_total + 7) / 8), 256, 0, st>>>(s, e->gcolptr.p, e->gpk_dev);
    CU(cudaGetLastError());
    e->last_launches += 2;
    CU(cudaStreamSynchronize(st));
    out->obj = e->gpk_host.obj; out->cost = e->gpk_host.cost; out->status = e->gpk_host.status;
    out->colptr = e->gpk_host.colptr; out->ind = e->gpk_host.ind; out->val = e->gpk_host.val;
    out->nnz = e->gpk_host.nnz[0];
    return 0;
}

// One round whose results land in the host mailboxes.  When the previous round wrote them too, the shared-memory /
// global-tableau kernels store into them from their epilogues and only the blocks of the legacy and cluster paths
// are copied afterwards; otherwise (first call, or another API was used in between) every block is copied once.
static int mailbox_round(dwgpu_engine *e, const double *d_pi, int phase_one, int initial, dwgpu_mailboxes *out)
{
    if (!out) return fail(DWGPU_E_ARGS, "out is NULL");
    cudaStream_t st = e->stream;
    const size_t K = (size_t)e->K, KL = K * (size_t)e->L;
    if (!e->h_mail) {
        const size_t bytes = sizeof(double) * (2 * K + KL) + sizeof(int) * (2 * K + KL) + 64;
        CU(cudaHostAlloc((void **)&e->h_mail, bytes, cudaHostAllocMapped));
        CU(cudaHostGetDevicePointer((void **)&e->d_mail, e->h_mail, 0));
        std::memset(e->h_mail, 0, bytes);
        auto carve = [&](char *base, dwk::MailDst &d) {
            double *dd = (double *)base;
            d.obj = dd; d.cost = dd + K; d.val = dd + 2 * K;
            int *ii = (int *)(dd + 2 * K + KL);
            d.status = ii; d.len = ii + K; d.ind = ii + 2 * K;
        };
        carve(e->d_mail, e->mail_dev);
        carve(e->h_mail, e->mail_host);
        e->mail_valid = false;
    }
    const bool direct = e->mail_valid;
    int rc = launch_round(e, d_pi, phase_one, initial, st, true);
    if (rc) return rc;
    dwk::PackSrc s;
    std::memset(&s, 0, sizeof(s));
    s.obj = e->obj.p; s.cost = e->cost.p; s.val = e->val.p; s.status = e->status.p; s.len = e->len.p; s.ind = e->ind.p;
    s.K = e->K; s.L = e->L;
    auto copy = [&](const int *list, size_t count) -> cudaError_t {
        if (!count) return cudaSuccess;
        dwk::mailbox_copy_kernel<256><<<(unsigned)((count + 7) / 8), 256, 0, st>>>(s, list, (int)count, e->mail_dev, e->counters.p);
        e->last_launches++;
        return cudaGetLastError();
    };
    if (!direct) CU(copy(nullptr, K));
    else { CU(copy(e->d_list_gmem.p, e->list_gmem.size())); CU(copy(e->d_list_c.p, e->list_c.size())); }
    e->mail_valid = true;
    CU(cudaStreamSynchronize(st));
    out->obj = e->mail_host.obj; out->cost = e->mail_host.cost; out->val = e->mail_host.val;
    out->status = e->mail_host.status; out->len = e->mail_host.len; out->ind = e->mail_host.ind;
    out->K = e->K; out->L = e->L;
    return 0;
}

int dwgpu_initial_solve_mailboxes(dwgpu_engine *e, dwgpu_mailboxes *out)
{
    if (!e || !e->initialised) return fail(DWGPU_E_STATE, "engine not initialised");
    CU(cudaSetDevice(e->device));
    return mailbox_round(e, nullptr, 0, 1, out);
}

int dwgpu_iterate_mailboxes(dwgpu_engine *e, const double *pi, int phase_one, dwgpu_mailboxes *out)
{
    if (!e || !e->initialised) return fail(DWGPU_E_STATE, "engine not initialised");
    if (!pi && e->L > 0) return fail(DWGPU_E_ARGS, "pi is NULL");
    CU(cudaSetDevice(e->device));
    if (e->L > 0) {
        std::memcpy(e->h_pi, pi, sizeof(double) * e->L);
        CU(cudaMemcpyAsync(e->pi.p, e->h_pi, sizeof(double) * e->L, cudaMemcpyHostToDevice, e->stream));
    }
    return mailbox_round(e, e->pi.p, phase_one ? 1 : 0, 0, out);
}

int dwgpu_initial_solve_packed(dwgpu_engine *e, dwgpu_packed *out)
{
    if (!e || !e->initialised) return fail(DWGPU_E_STATE, "engine not initialised");
    CU(cudaSetDevice(e->device));
    int rc = launch_round(e, nullptr, 0, 1, e->stream);
    if (rc) return rc;
    return pack_round(e, nullptr, out);
}

int dwgpu_iterate_packed(dwgpu_engine *e, const double *pi, int phase_one, const double *r, dwgpu_packed *out)
{
    if (!e || !e->initialised) return fail(DWGPU_E_STATE, "engine not initialised");
    if (!pi && e->L > 0) return fail(DWGPU_E_ARGS, "pi is NULL");
    CU(cudaSetDevice(e->device));
    if (e->L > 0) {
        std::memcpy(e->h_pi, pi, sizeof(double) * e->L);
        CU(cudaMemcpyAsync(e->pi.p, e->h_pi, sizeof(double) * e->L, cudaMemcpyHostToDevice, e->stream));
    }
    int rc = launch_round(e, e->pi.p, phase_one ? 1 : 0, 0, e->stream);
    if (rc) return rc;
    return pack_round(e, r, out);
}

int dwgpu_reset(dwgpu_engine *e, void *stream)
{
    if (!e || !e->initialised) return fail(DWGPU_E_STATE, "engine not initialised");
    CU(cudaSetDevice(e->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
    dwk::init_kernel<256><<<e->K, 256, 0, st>>>(e->blocks.p, e->K);
    CU(cudaGetLastError());
    e->have_proposals = false;
    e->mail_valid = false;
    e->uplink_valid = false;
    e->pi_ref_valid = false;
    if (st != e->stream && e->ev_done) { CU(cudaEventRecord(e->ev_done, st)); e->ev_done_pending = true; }
    return 0;
}

int dwgpu_initial_solve_device(dwgpu_engine *e, void *stream)
{
    if (!e || !e->initialised) return fail(DWGPU_E_STATE, "engine not initialised");
    CU(cudaSetDevice(e->device));
    return launch_round(e, nullptr, 0, 1, stream ? (cudaStream_t)stream : e->stream);
}

int dwgpu_initial_solve(dwgpu_engine *e, dwgpu_proposals *out)
{
    if (!e || !e->initialised) return fail(DWGPU_E_STATE, "engine not initialised");
    CU(cudaSetDevice(e->device));
    int rc = launch_round(e, nullptr, 0, 1, e->stream);
    if (rc) return rc;
    return dwgpu_fetch_results(e, out);
}

int dwgpu_iterate(dwgpu_engine *e, const double *pi, int phase_one, dwgpu_proposals *out)
{
    if (!e || !e->initialised) return fail(DWGPU_E_STATE, "engine not initialised");
    if (!pi && e->L > 0) return fail(DWGPU_E_ARGS, "pi is NULL");
    CU(cudaSetDevice(e->device));
    if (e->L > 0) {
        std::memcpy(e->h_pi, pi, sizeof(double) * e->L);
        CU(cudaMemcpyAsync(e->pi.p, e->h_pi, sizeof(double) * e->L, cudaMemcpyHostToDevice, e->stream));
    }
    int rc = launch_round(e, e->pi.p, phase_one ? 1 : 0, 0, e->stream);
    if (rc) return rc;
    return dwgpu_fetch_results(e, out);
}

int dwgpu_iterate_device(dwgpu_engine *e, const double *d_pi, int phase_one, void *stream)
{
    if (!e || !e->initialised) return fail(DWGPU_E_STATE, "engine not initialised");
    if (!d_pi && e->L > 0) return fail(DWGPU_E_ARGS, "d_pi is NULL");
    CU(cudaSetDevice(e->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
    return launch_round(e, d_pi, phase_one ? 1 : 0, 0, st);
}

int dwgpu_device_results(dwgpu_engine *e, dwgpu_proposals *d)
{
    if (!e || !d) return fail(DWGPU_E_ARGS, "NULL argument");
    d->obj = e->obj.p; d->cost = e->cost.p; d->status = e->status.p; d->len = e->len.p;
    d->ind = e->ind.p; d->val = e->val.p; d->x = e->x.p;
    return 0;
}

int dwgpu_fetch_x(dwgpu_engine *e, int k, double *x)
{
    if (!e || !x || k < 0 || k >= e->K) return fail(DWGPU_E_ARGS, "bad arguments");
    CU(cudaSetDevice(e->device));
    CU(after_last_round(e));
    CU(cudaStreamSynchronize(e->stream));
    if (e->n[k] > 0)
        CU(cudaMemcpy(x, e->x.p + e->xoff[k], sizeof(double) * e->n[k], cudaMemcpyDeviceToHost));
    return 0;
}

int dwgpu_history_push(dwgpu_engine *e, int count, const int *blocks, long long *ids_out)
{
    if (!e || count < 0 || (count > 0 && (!blocks || !ids_out))) return fail(DWGPU_E_ARGS, "bad arguments");
    if (count == 0) return 0;
    CU(cudaSetDevice(e->device));
    CU(after_last_round(e));
    cudaStream_t st = e->stream;
    long long need = 0;
    for (int i = 0; i < count; ++i) {
        if (blocks[i] < 0 || blocks[i] >= e->K) return fail(DWGPU_E_ARGS, "block index out of range");
        need += e->n[blocks[i]];
    }
    if (e->hist_used + need > e->hist_cap) {                 // grow geometrically, keep what is there
        long long cap = std::max<long long>(2 * e->hist_cap, e->hist_used + need);
        cap = std::max<long long>(cap, 2 * e->ntot);
        DevBuf<double> bigger;
        CU(bigger.alloc((size_t)cap));
        if (e->hist_used > 0)
            CU(cudaMemcpyAsync(bigger.p, e->hist.p, sizeof(double) * (size_t)e->hist_used, cudaMemcpyDeviceToDevice, st));
        CU(cudaStreamSynchronize(st));
        e->hist.release();
        e->hist = bigger;
        e->hist_cap = cap;
    }
    std::vector<long long> src(count), dst(count);
    std::vector<int> len(count);
    for (int i = 0; i < count; ++i) {
        const int k = blocks[i];
        src[i] = e->xoff[k]; dst[i] = e->hist_used; len[i] = e->n[k];
        ids_out[i] = (long long)e->hist_off.size();
        e->hist_off.push_back(e->hist_used);
        e->hist_block.push_back(k);
        e->hist_used += e->n[k];
    }
    // index lists through persistent device staging (grow-only): no cudaMalloc / cudaFree per round
    if ((size_t)count > e->hp_cap) {
        e->hp_src.release(); e->hp_dst.release(); e->hp_len.release();
        e->hp_cap = (size_t)std::max(count * 2, e->K);
        CU(e->hp_src.alloc(e->hp_cap)); CU(e->hp_dst.alloc(e->hp_cap)); CU(e->hp_len.alloc(e->hp_cap));
    }
    CU(cudaMemcpyAsync(e->hp_src.p, src.data(), sizeof(long long) * (size_t)count, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(e->hp_dst.p, dst.data(), sizeof(long long) * (size_t)count, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(e->hp_len.p, len.data(), sizeof(int) * (size_t)count, cudaMemcpyHostToDevice, st));
    dwk::history_push_kernel<<<(unsigned)count, 256, 0, st>>>(e->x.p, e->hist.p, e->hp_src.p, e->hp_dst.p, e->hp_len.p, count);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));          // the pageable host lists above are read by the copies
    return 0;
}

// Several dual vectors per round (the reference prices ONE vector per round, src/dw_phases.c:358-437): for
// p = 0 .. P-1 every block is priced with Pi[p] and re-solved from the basis the previous vector left, its proposal packed
// into record p and its x_k kept in the device history (ids_out[p * K + k]); one synchronisation at the end.  Blocks with
// a dense D_k (cluster path) get their P priced cost vectors from one FP64 tensor-core GEMM up front.
int dwgpu_iterate_multi(dwgpu_engine *e, const double *Pi, int P, int phase_one, dwgpu_packed *out, long long *ids_out)
{
    if (!e || !e->initialised) return fail(DWGPU_E_STATE, "engine not initialised");
    if (P < 1 || P > 32 || !out || (e->L > 0 && !Pi)) return fail(DWGPU_E_ARGS, "1 <= P <= 32 dual vectors, out = P records");
    CU(cudaSetDevice(e->device));
    cudaStream_t st = e->stream;
    const size_t K = (size_t)e->K, KL = K * (size_t)e->L, L = (size_t)e->L;
    const size_t rec_stride = (e->pack_bytes + 63) & ~(size_t)63;        // records start 64-byte aligned
    if (P > e->multi_P) {
        CU(cudaStreamSynchronize(st));
        e->multi_pi.release(); e->multi_cs.release();
        if (e->h_mpack) { cudaFreeHost(e->h_mpack); e->h_mpack = nullptr; }
        CU(e->multi_pi.alloc((size_t)P * std::max<size_t>(L, 1)));
        if (!e->list_dense.empty()) CU(e->multi_cs.alloc((size_t)P * (size_t)e->ntot));
        CU(cudaHostAlloc((void **)&e->h_mpack, rec_stride * (size_t)P, cudaHostAllocMapped));
        CU(cudaHostGetDevicePointer((void **)&e->d_mpack, e->h_mpack, 0));
        std::memset(e->h_mpack, 0, rec_stride * (size_t)P);
        e->mpk_dev.assign((size_t)P, dwk::PackDst{}); e->mpk_host.assign((size_t)P, dwk::PackDst{});
        auto carve = [&](char *base, dwk::PackDst &d) {
            d.nnz = (long long *)base;
            double *dd = (double *)(base + sizeof(double));
            d.obj = dd; d.cost = dd + K; d.val = dd + 2 * K;
            int *ii = (int *)(dd + 2 * K + KL);
            d.status = ii; d.colptr = ii + K; d.ind = ii + 2 * K + 1;
        };
        for (int p = 0; p < P; ++p) {
            carve(e->d_mpack + rec_stride * (size_t)p, e->mpk_dev[(size_t)p]);
            carve(e->h_mpack + rec_stride * (size_t)p, e->mpk_host[(size_t)p]);
        }
        e->multi_P = P;
    }
    if (L > 0) CU(cudaMemcpyAsync(e->multi_pi.p, Pi, sizeof(double) * (size_t)P * L, cudaMemcpyHostToDevice, st));
    if (!e->list_dense.empty() && L > 0) {
        int nmax = 0;
        for (int k : e->list_dense) nmax = std::max(nmax, e->n[k]);
        const size_t sm = sizeof(double) * 32 * (L + 4);
        CU(allow_dynamic_smem(e->device, (const void *)dwk::price_multi_dmma_kernel, (int)sm));
        dwk::price_multi_dmma_kernel<<<dim3((unsigned)e->list_dense.size(), (unsigned)((nmax + 63) / 64)), 256, sm, st>>>(
            e->blocks.p, e->d_list_dense.p, e->multi_pi.p, P, phase_one ? 1 : 0, e->ntot, e->multi_cs.p);
        CU(cudaGetLastError());
    }
    // history room for P copies of every block's x
    if (ids_out) {
        const long long need = (long long)P * e->ntot;
        if (e->hist_used + need > e->hist_cap) {
            long long cap = std::max<long long>(2 * e->hist_cap, e->hist_used + need);
            DevBuf<double> bigger;
            CU(bigger.alloc((size_t)cap));
            if (e->hist_used > 0)
                CU(cudaMemcpyAsync(bigger.p, e->hist.p, sizeof(double) * (size_t)e->hist_used, cudaMemcpyDeviceToDevice, st));
            CU(cudaStreamSynchronize(st));
            e->hist.release();
            e->hist = bigger;
            e->hist_cap = cap;
        }
    }
    int launches = 0;
    for (int p = 0; p < P; ++p) {
        e->cs_pre_now = e->list_dense.empty() ? nullptr : e->multi_cs.p + (size_t)p * (size_t)e->ntot;
        const int rc = launch_round(e, e->multi_pi.p + (size_t)p * L, phase_one ? 1 : 0, 0, st);
        e->cs_pre_now = nullptr;
        if (rc) return rc;
        launches += e->last_launches;
        dwk::PackSrc s;
        std::memset(&s, 0, sizeof(s));
        s.obj = e->obj.p; s.cost = e->cost.p; s.val = e->val.p; s.status = e->status.p; s.len = e->len.p; s.ind = e->ind.p;
        s.K = e->K; s.L = e->L; s.accept_tol = 1e-6;
        dwk::pack_scan_kernel<1024><<<1, 1024, 0, st>>>(s, e->colptr.p, e->mpk_dev[(size_t)p]);
        CU(cudaGetLastError());
        dwk::pack_copy_kernel<256><<<(unsigned)((e->K + 7) / 8), 256, 0, st>>>(s, e->colptr.p, e->mpk_dev[(size_t)p]);
        CU(cudaGetLastError());
        launches += 2;
        if (ids_out) {           // x of all blocks is one contiguous array: one device-to-device copy keeps it
            CU(cudaMemcpyAsync(e->hist.p + e->hist_used, e->x.p, sizeof(double) * (size_t)e->ntot, cudaMemcpyDeviceToDevice, st));
            for (int k = 0; k < e->K; ++k) {
                ids_out[(size_t)p * K + k] = (long long)e->hist_off.size();
                e->hist_off.push_back(e->hist_used + e->xoff[k]);
                e->hist_block.push_back(k);
            }
            e->hist_used += e->ntot;
        }
    }
    e->last_launches = launches;
    CU(cudaStreamSynchronize(st));
    for (int p = 0; p < P; ++p) {
        const dwk::PackDst &h = e->mpk_host[(size_t)p];
        out[p].obj = h.obj; out[p].cost = h.cost; out[p].status = h.status; out[p].colptr = h.colptr;
        out[p].ind = h.ind; out[p].val = h.val; out[p].nnz = h.nnz[0];
    }
    return 0;
}

/* number of blocks whose pricing runs through the multi-vector FP64 tensor-core GEMM */
int dwgpu_dense_pricing_blocks(dwgpu_engine *e) { return e ? (int)e->list_dense.size() : 0; }

int dwgpu_history_combine(dwgpu_engine *e, int nterms, const long long *ids, const double *weights, double *x_out)
{
    if (!e || nterms < 0 || !x_out || (nterms > 0 && (!ids || !weights))) return fail(DWGPU_E_ARGS, "bad arguments");
    CU(cudaSetDevice(e->device));
    cudaStream_t st = e->stream;
    // group the terms by block, keeping the given order inside a block
    std::vector<int> cnt(e->K + 1, 0);
    for (int t = 0; t < nterms; ++t) {
        if (ids[t] < 0 || ids[t] >= (long long)e->hist_off.size()) return fail(DWGPU_E_ARGS, "unknown history id");
        cnt[e->hist_block[(size_t)ids[t]] + 1]++;
    }
    for (int k = 0; k < e->K; ++k) cnt[k + 1] += cnt[k];
    std::vector<long long> tid((size_t)std::max(nterms, 1));
    std::vector<double> tw((size_t)std::max(nterms, 1));
    {
        std::vector<int> pos(cnt.begin(), cnt.end() - 1);
        for (int t = 0; t < nterms; ++t) {
            const int k = e->hist_block[(size_t)ids[t]];
            tid[(size_t)pos[k]] = e->hist_off[(size_t)ids[t]];
            tw[(size_t)pos[k]] = weights[t];
            pos[k]++;
        }
    }
    std::vector<long long> xoff(e->xoff.begin(), e->xoff.end() - 1);
    DevBuf<long long> d_xoff, d_tid;
    DevBuf<int> d_len, d_ptr;
    DevBuf<double> d_tw, d_out;
    CU(d_xoff.upload(xoff)); CU(d_len.upload(e->n)); CU(d_ptr.upload(cnt)); CU(d_tid.upload(tid)); CU(d_tw.upload(tw));
    CU(d_out.alloc((size_t)std::max<long long>(e->ntot, 1)));
    dwk::history_combine_kernel<<<(unsigned)e->K, 256, 0, st>>>(e->hist.p, d_out.p, d_xoff.p, d_len.p, d_ptr.p, d_tid.p, d_tw.p, e->K);
    CU(cudaGetLastError());
    if (e->ntot > 0) CU(cudaMemcpyAsync(x_out, d_out.p, sizeof(double) * (size_t)e->ntot, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    d_xoff.release(); d_tid.release(); d_len.release(); d_ptr.release(); d_tw.release(); d_out.release();
    return 0;
}

int dwgpu_get_counters(dwgpu_engine *e, dwgpu_counters *t)
{
    if (!e || !t) return fail(DWGPU_E_ARGS, "NULL argument");
    CU(cudaSetDevice(e->device));
    CU(after_last_round(e));
    CU(cudaStreamSynchronize(e->stream));
    std::vector<unsigned long long> h((size_t)e->K * 8);
    CU(cudaMemcpy(h.data(), e->counters.p, sizeof(unsigned long long) * h.size(), cudaMemcpyDeviceToHost));
    std::memset(t, 0, sizeof(*t));
    for (int k = 0; k < e->K; ++k) {
        t->pivots += (long long)h[8 * (size_t)k]; t->flips += (long long)h[8 * (size_t)k + 1];
        t->rebuilds += (long long)h[8 * (size_t)k + 2]; t->phase1_iter += (long long)h[8 * (size_t)k + 3];
        t->rows_swept += (long long)h[8 * (size_t)k + 4];
        t->cells_swept += (long long)h[8 * (size_t)k + 5];
        t->price_cells += (long long)h[8 * (size_t)k + 6];
        t->mailbox_bytes += (long long)h[8 * (size_t)k + 7];
    }
    t->solves = e->solves;
    return 0;
}

int dwgpu_get_block_iterations(dwgpu_engine *e, long long *iters)
{
    if (!e || !iters) return fail(DWGPU_E_ARGS, "NULL argument");
    CU(cudaSetDevice(e->device));
    CU(after_last_round(e));
    CU(cudaStreamSynchronize(e->stream));
    CU(cudaMemcpy(iters, e->last_iters.p, sizeof(long long) * e->K, cudaMemcpyDeviceToHost));
    return 0;
}

int dwgpu_last_launch_count(dwgpu_engine *e) { return e ? e->last_launches : 0; }

int dwgpu_get_block_paths(dwgpu_engine *e, int *paths)
{
    if (!e || !paths) return fail(DWGPU_E_ARGS, "NULL argument");
    for (int k = 0; k < e->K; ++k) paths[k] = e->path[k];
    return 0;
}

int dwgpu_last_kernel_ms(dwgpu_engine *e, float ms[2])
{
    if (!e || !ms) return fail(DWGPU_E_ARGS, "NULL argument");
    CU(cudaSetDevice(e->device));
    ms[0] = ms[1] = 0.f;
    if (e->ev_used[0]) { CU(cudaEventSynchronize(e->ev[3])); CU(cudaEventElapsedTime(&ms[0], e->ev[2], e->ev[3])); }
    if (e->ev_used[1]) { CU(cudaEventSynchronize(e->ev[1])); CU(cudaEventElapsedTime(&ms[1], e->ev[0], e->ev[1])); }
    return 0;
}

namespace {
// FP64 read-modify-write sweep over a shared-memory tile: the access pattern of the pivot update
__global__ void __launch_bounds__(256) smem_rw_kernel(double *sink, int n_elems, int iters)
{
    extern __shared__ __align__(16) double tile[];
    for (int e = threadIdx.x; e < n_elems; e += 256) tile[e] = (double)e;
    __syncthreads();
    const double a = 1.0000001, b = 1e-9;
    for (int it = 0; it < iters; ++it) {
        for (int e = threadIdx.x; e < n_elems; e += 256) tile[e] = fma(tile[e], a, b);
        __syncthreads();
    }
    if (sink) sink[blockIdx.x * 256 + threadIdx.x] = tile[threadIdx.x];
}
}  // namespace

int dwgpu_measure_smem_bandwidth(int device, double *gbps)
{
    if (!gbps) return fail(DWGPU_E_ARGS, "NULL argument");
    CU(cudaSetDevice(device));
    int sms = 0;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int n_elems = 12800, iters = 2000;             // 100 KB tile, 2 CTAs per SM
    const size_t bytes = sizeof(double) * n_elems;
    CU(cudaFuncSetAttribute(smem_rw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    double *sink = nullptr;
    CU(cudaMalloc((void **)&sink, sizeof(double) * 256 * (size_t)sms * 2));
    cudaEvent_t a, b;
    CU(cudaEventCreate(&a)); CU(cudaEventCreate(&b));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CU(cudaEventRecord(a, 0));
        smem_rw_kernel<<<sms * 2, 256, bytes, 0>>>(sink, n_elems, iters);
        CU(cudaEventRecord(b, 0));
        CU(cudaEventSynchronize(b));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, a, b));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(sink);
    *gbps = 16.0 * n_elems * (double)iters * sms * 2 / (best * 1e-3) / 1e9;
    return 0;
}

int dwgpu_sync(dwgpu_engine *e)
{
    if (!e) return fail(DWGPU_E_ARGS, "engine is NULL");
    CU(cudaSetDevice(e->device));
    CU(cudaStreamSynchronize(e->stream));
    return 0;
}

}  // extern "C"
