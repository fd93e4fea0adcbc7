// instance.cc — see instance.h.
#include "instance.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <thread>

namespace dwh {

static double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

double to_engine_bound(double v) { return v >= 1e300 ? DWGPU_INF : (v <= -1e300 ? -DWGPU_INF : v); }

// CSR of the block's rows (duplicate entries of a row are summed, like glp_read_lp's coefficient merge)
static void build_block(const LpModel &sub, const LpModel &master, const std::vector<std::vector<std::pair<int, double>>> &mcols,
                        BlockHost &b)
{
    b.m = sub.m(); b.n = sub.n();
    std::vector<std::map<int, double>> rows(b.m);
    for (size_t t = 0; t < sub.a_val.size(); ++t) rows[sub.a_row[t]][sub.a_col[t]] += sub.a_val[t];
    b.a_rowptr.assign(b.m + 1, 0);
    for (int i = 0; i < b.m; ++i) {
        for (auto &e : rows[i]) { b.a_col.push_back(e.first); b.a_val.push_back(e.second); }
        b.a_rowptr[i + 1] = (int)b.a_col.size();
    }
    b.row_lb.resize(b.m); b.row_ub.resize(b.m);
    for (int i = 0; i < b.m; ++i) { b.row_lb[i] = to_engine_bound(sub.row_lb[i]); b.row_ub[i] = to_engine_bound(sub.row_ub[i]); }
    b.col_lb.resize(b.n); b.col_ub.resize(b.n); b.c_file = sub.obj; b.c_master.assign(b.n, 0.0);
    b.col_names = sub.col_names;
    for (int j = 0; j < b.n; ++j) {
        b.col_lb[j] = to_engine_bound(sub.col_lb[j]); b.col_ub[j] = to_engine_bound(sub.col_ub[j]);
        auto it = master.col_index.find(sub.col_names[j]);
        if (it == master.col_index.end())
            throw std::runtime_error("block column '" + sub.col_names[j] + "' not found in the master problem");
        const int mj = it->second;
        b.c_master[j] = master.obj[mj];                              // src/dw_subprob.c:228
        for (auto &e : mcols[mj]) {                                  // :230-253, raw coefficients
            b.d_row.push_back(e.first); b.d_col.push_back(j); b.d_val.push_back(e.second);
        }
    }
}

// Master and block files -> in-memory blocks (src/dw_solver.c:192-198, src/dw_subprob.c:105-111, 203-258).
// The reference reads the K block files under two process-wide mutexes (glpk_mutex, master_mutex), i.e. one after
// the other.  Block files are independent and the master is read-only once parsed, so `readers` host threads
// (0 = one per core, at most 16; DWSOLVER_READ_THREADS overrides) take them in an interleaved order.
// The result does not depend on the number of readers.  Throws std::runtime_error on an unreadable file.
void read_problem(const char *master_file, const char *const *subproblem_files, int K, int readers, LpModel &master,
                  std::vector<BlockHost> &blocks)
{
    const bool dbg = std::getenv("DWSOLVER_READ_DEBUG") != nullptr;
    const double t0 = now_s();
    master = read_lp_file(master_file);
    // Objective direction: the reference leaves a file's Maximize in its glp_prob and glp_simplex honours it — for a block
    // that would MAXIMISE the reduced cost the master asks it to minimise (src/dw_subprob.c:320-329).  Neither is a
    // Dantzig-Wolfe run; this library says so instead of silently minimising.
    if (master.maximize) throw std::runtime_error(std::string("the master file ") + master_file + " has a Maximize objective: only minimisation is supported");
    const double t1 = now_s();
    blocks.assign((size_t)K, BlockHost());
    std::vector<std::vector<std::pair<int, double>>> mcols(master.n());
    for (size_t t = 0; t < master.a_val.size(); ++t) mcols[master.a_col[t]].emplace_back(master.a_row[t], master.a_val[t]);
    if (readers <= 0) {
        readers = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
        if (const char *e = std::getenv("DWSOLVER_READ_THREADS")) readers = std::max(1, std::atoi(e));
    }
    readers = std::max(1, std::min(readers, K));
    std::vector<std::string> errors((size_t)readers);
    auto read_some = [&](int t) {
        try {
            for (int k = t; k < K; k += readers) {
                LpModel sub = read_lp_file(subproblem_files[k]);
                if (sub.maximize) throw std::runtime_error(std::string("the subproblem file ") + subproblem_files[k] + " has a Maximize objective: only minimisation is supported");
                build_block(sub, master, mcols, blocks[k]);
            }
        } catch (const std::exception &e) {
            errors[(size_t)t] = e.what()[0] ? e.what() : "cannot read a subproblem file";
        }
    };
    const double t2 = now_s();
    std::vector<std::thread> pool;
    for (int t = 1; t < readers; ++t) pool.emplace_back(read_some, t);
    read_some(0);
    for (auto &th : pool) th.join();
    if (dbg) std::fprintf(stderr, "dwsolver-b200: read master %.3f s, column index %.3f s, %d block files %.3f s with %d reader(s)\n", t1 - t0, t2 - t1, K, now_s() - t2, readers);
    for (auto &msg : errors) if (!msg.empty()) throw std::runtime_error(msg);
}

// ---- binary block-angular container ---------------------------------------------------------------
// One file holding exactly what read_problem() produces, so that a second solve of the same instance skips the LP
// text: little-endian, every array 8-byte aligned, read with one fread and handed to the engine without a copy.
//   header   : char magic[8] = "DWB200\0\1"; int32 K, L, has_names, 0; double shift (the guidefile's objective
//              constant); double link_lb[L], link_ub[L]
//   table    : K x { int32 m, n, a_nnz, d_nnz }
//   block k  : int32 a_rowptr[m+1], a_col[a_nnz], d_row[d_nnz], d_col[d_nnz], (pad to 8);
//              double a_val[a_nnz], row_lb[m], row_ub[m], col_lb[n], col_ub[n], c_file[n], c_master[n], d_val[d_nnz]
//   names    : if has_names, per block n NUL-terminated column names, in column order
// Bounds are stored the way the engine takes them (no bound = +-DWGPU_INF).
static constexpr char kContainerMagic[8] = {'D', 'W', 'B', '2', '0', '0', 0, 1};

struct ContainerWriter {
    std::FILE *f;
    size_t written = 0;
    bool ok = true;
    void raw(const void *p, size_t bytes)
    {
        if (bytes && std::fwrite(p, 1, bytes, f) != bytes) ok = false;
        written += bytes;
    }
    template <class T> void vec(const std::vector<T> &v) { raw(v.data(), v.size() * sizeof(T)); }
    void pad8() { static const char zero[8] = {0}; if (written % 8) raw(zero, 8 - written % 8); }
};


// throws std::runtime_error on a missing, truncated or foreign file
void load_container(const char *path, ContainerView &v)
{
    std::FILE *f = std::fopen(path, "rb");
    if (!f) throw std::runtime_error(std::string("cannot open container ") + path);
    std::fseek(f, 0, SEEK_END);
    const long size = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    v.image.resize(size > 0 ? (size_t)size : 0);
    const size_t got = v.image.empty() ? 0 : std::fread(v.image.data(), 1, v.image.size(), f);
    std::fclose(f);
    if (got != v.image.size()) throw std::runtime_error(std::string("cannot read container ") + path);
    const char *base = v.image.data();
    size_t pos = 0;
    const size_t end = v.image.size();
    auto take = [&](size_t bytes) -> const char * {
        if (bytes > end - pos) throw std::runtime_error(std::string(path) + ": truncated container");
        const char *p = base + pos;
        pos += bytes;
        return p;
    };
    if (end < 32 || std::memcmp(take(8), kContainerMagic, 8) != 0)
        throw std::runtime_error(std::string(path) + ": not a dwsolver-b200 container (or another format version)");
    const int *hdr = reinterpret_cast<const int *>(take(16));
    v.K = hdr[0]; v.L = hdr[1]; v.has_names = hdr[2] != 0;
    if (v.K < 1 || v.L < 0) throw std::runtime_error(std::string(path) + ": bad container header");
    std::memcpy(&v.shift, take(8), 8);
    v.link_lb = reinterpret_cast<const double *>(take(sizeof(double) * (size_t)v.L));
    v.link_ub = reinterpret_cast<const double *>(take(sizeof(double) * (size_t)v.L));
    const int *table = reinterpret_cast<const int *>(take(16 * (size_t)v.K));
    v.desc.assign((size_t)v.K, dwgpu_block_desc());
    for (int k = 0; k < v.K; ++k) {
        const int m = table[4 * k], n = table[4 * k + 1], an = table[4 * k + 2], dn = table[4 * k + 3];
        if (m < 0 || n < 0 || an < 0 || dn < 0) throw std::runtime_error(std::string(path) + ": bad block sizes in the container");
        dwgpu_block_desc &d = v.desc[(size_t)k];
        d.m = m; d.n = n; d.d_nnz = dn;
        d.a_rowptr = reinterpret_cast<const int *>(take(4 * ((size_t)m + 1)));
        d.a_col = reinterpret_cast<const int *>(take(4 * (size_t)an));
        d.d_row = reinterpret_cast<const int *>(take(4 * (size_t)dn));
        d.d_col = reinterpret_cast<const int *>(take(4 * (size_t)dn));
        if (pos % 8) take(8 - pos % 8);
        auto dbl = [&](size_t cnt) { return reinterpret_cast<const double *>(take(8 * cnt)); };
        d.a_val = dbl((size_t)an);
        d.row_lb = dbl((size_t)m); d.row_ub = dbl((size_t)m);
        d.col_lb = dbl((size_t)n); d.col_ub = dbl((size_t)n);
        d.c_file = dbl((size_t)n); d.c_master = dbl((size_t)n);
        d.d_val = dbl((size_t)dn);
        // the structure the engine indexes with must be sane before it reaches the device
        if (d.a_rowptr[0] != 0 || d.a_rowptr[m] != an) throw std::runtime_error(std::string(path) + ": corrupt row pointers in the container");
        for (int i = 0; i < m; ++i) if (d.a_rowptr[i] > d.a_rowptr[i + 1]) throw std::runtime_error(std::string(path) + ": corrupt row pointers in the container");
        for (int t = 0; t < an; ++t) if (d.a_col[t] < 0 || d.a_col[t] >= n) throw std::runtime_error(std::string(path) + ": column index out of range in the container");
        for (int t = 0; t < dn; ++t)
            if (d.d_row[t] < 0 || d.d_row[t] >= v.L || d.d_col[t] < 0 || d.d_col[t] >= n)
                throw std::runtime_error(std::string(path) + ": linking entry out of range in the container");
    }
    if (v.has_names) {
        v.names.assign((size_t)v.K, {});
        for (int k = 0; k < v.K; ++k) {
            v.names[(size_t)k].reserve((size_t)v.desc[(size_t)k].n);
            for (int j = 0; j < v.desc[(size_t)k].n; ++j) {
                const void *z = pos < end ? std::memchr(base + pos, 0, end - pos) : nullptr;
                if (!z) throw std::runtime_error(std::string(path) + ": truncated names in the container");
                const size_t len = (size_t)(static_cast<const char *>(z) - (base + pos));
                v.names[(size_t)k].emplace_back(base + pos, len);
                pos += len + 1;
            }
        }
    }
}

// order-sensitive digest shared by the two test hooks below (text path and container path must agree bit for bit)
void digest_blocks(int K, int L, const dwgpu_block_desc *desc, double *out)
{
    for (int i = 0; i < 8; ++i) out[i] = 0.0;
    out[0] = L;
    auto fin = [](double v) { return std::fabs(v) >= 1e29 ? (v > 0 ? 7.0 : -7.0) : v; };
    for (int k = 0; k < K; ++k) {
        const dwgpu_block_desc &b = desc[k];
        const int an = b.a_rowptr[b.m];
        out[1] += b.m; out[2] += b.n; out[3] += an; out[4] += b.d_nnz;
        const double w = 1.0 + (k % 97);
        for (int i = 0; i < b.m; ++i)
            for (int t = b.a_rowptr[i]; t < b.a_rowptr[i + 1]; ++t) out[5] += w * b.a_val[t] * (1 + (i % 13)) * (1 + (b.a_col[t] % 17));
        for (int t = 0; t < b.d_nnz; ++t) out[6] += w * b.d_val[t] * (1 + (b.d_row[t] % 13)) * (1 + (b.d_col[t] % 17));
        for (int i = 0; i < b.m; ++i) out[7] += w * (fin(b.row_lb[i]) + 2 * fin(b.row_ub[i]));
        for (int j = 0; j < b.n; ++j) out[7] += w * (fin(b.col_lb[j]) + 2 * fin(b.col_ub[j]) + 3 * b.c_file[j] + 5 * b.c_master[j]) * (1 + (j % 11));
    }
}

void describe(std::vector<BlockHost> &blocks, std::vector<dwgpu_block_desc> &desc)
{
    desc.assign(blocks.size(), dwgpu_block_desc());
    for (size_t k = 0; k < blocks.size(); ++k) {
        BlockHost &b = blocks[k];
        dwgpu_block_desc &d = desc[k];
        d.m = b.m; d.n = b.n;
        d.a_rowptr = b.a_rowptr.data(); d.a_col = b.a_col.data(); d.a_val = b.a_val.data();
        d.row_lb = b.row_lb.data(); d.row_ub = b.row_ub.data(); d.col_lb = b.col_lb.data(); d.col_ub = b.col_ub.data();
        d.c_file = b.c_file.data(); d.c_master = b.c_master.data();
        d.d_nnz = (int)b.d_val.size(); d.d_row = b.d_row.data(); d.d_col = b.d_col.data(); d.d_val = b.d_val.data();
    }
}


bool write_container(const char *path, const LpModel &master, const std::vector<BlockHost> &blocks, double shift)
{
    const int K = (int)blocks.size(), L = master.m();
    std::FILE *f = std::fopen(path, "wb");
    if (!f) return false;
    ContainerWriter w{f};
    w.raw(kContainerMagic, 8);
    const int hdr[4] = {K, L, 1, 0};
    w.raw(hdr, sizeof hdr);
    w.raw(&shift, sizeof shift);
    std::vector<double> llb(L), lub(L);
    for (int i = 0; i < L; ++i) { llb[i] = to_engine_bound(master.row_lb[i]); lub[i] = to_engine_bound(master.row_ub[i]); }
    w.vec(llb); w.vec(lub);
    for (int k = 0; k < K; ++k) {
        const BlockHost &b = blocks[k];
        const int t[4] = {b.m, b.n, (int)b.a_val.size(), (int)b.d_val.size()};
        w.raw(t, sizeof t);
    }
    for (int k = 0; k < K; ++k) {
        const BlockHost &b = blocks[k];
        w.vec(b.a_rowptr); w.vec(b.a_col); w.vec(b.d_row); w.vec(b.d_col);
        w.pad8();
        w.vec(b.a_val); w.vec(b.row_lb); w.vec(b.row_ub); w.vec(b.col_lb); w.vec(b.col_ub); w.vec(b.c_file); w.vec(b.c_master);
        w.vec(b.d_val);
    }
    for (int k = 0; k < K; ++k)
        for (const std::string &nm : blocks[k].col_names) w.raw(nm.c_str(), nm.size() + 1);
    const bool closed = std::fclose(f) == 0;
    return w.ok && closed;
}

}  // namespace dwh
