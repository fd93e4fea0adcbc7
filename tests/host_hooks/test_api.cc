// test_api.cc — TEST-ONLY library (libdwhost_testhooks.so, never linked into the product): C entry points that let the
// (Python) test-suite exercise the host pieces in isolation:
// the LP reader against the package's own reader, the master LP solver against HiGHS.
#include <cstring>
#include <stdexcept>
#include <vector>

#include "../../dwsolver_b200/host/lp_model.h"
#include "../../dwsolver_b200/host/gub_master.h"
#include "../../dwsolver_b200/host/master_lp.h"
#include "../../dwsolver_b200/host/instance.h"
#include <cstdio>
#include <string>

extern "C" {

// dims[0..2] = rows, columns, non-zeros; returns 0, or 1 when the file cannot be read / parsed
int dwhost_lp_dims(const char *path, int *dims)
{
    try {
        const dwh::LpModel lp = dwh::read_lp_file(path);
        dims[0] = lp.m(); dims[1] = lp.n(); dims[2] = (int)lp.a_val.size();
        return 0;
    } catch (const std::exception &) { return 1; }
}

// fills caller-allocated arrays (sizes from dwhost_lp_dims); names are written as '\n'-joined text
int dwhost_lp_read(const char *path, double *obj, double *col_lb, double *col_ub, double *row_lb, double *row_ub,
                   int *a_row, int *a_col, double *a_val, char *col_names, int col_names_cap, int *maximize)
{
    try {
        const dwh::LpModel lp = dwh::read_lp_file(path);
        for (int j = 0; j < lp.n(); ++j) { obj[j] = lp.obj[j]; col_lb[j] = lp.col_lb[j]; col_ub[j] = lp.col_ub[j]; }
        for (int i = 0; i < lp.m(); ++i) { row_lb[i] = lp.row_lb[i]; row_ub[i] = lp.row_ub[i]; }
        for (size_t t = 0; t < lp.a_val.size(); ++t) { a_row[t] = lp.a_row[t]; a_col[t] = lp.a_col[t]; a_val[t] = lp.a_val[t]; }
        std::string names;
        for (auto &s : lp.col_names) { names += s; names += '\n'; }
        if ((int)names.size() + 1 > col_names_cap) return 2;
        std::memcpy(col_names, names.c_str(), names.size() + 1);
        *maximize = lp.maximize ? 1 : 0;
        return 0;
    } catch (const std::exception &) { return 1; }
}

// min cost'x  s.t.  A x = rhs, 0 <= x <= ub (ub >= 1e300: none); A by columns (CSC).
// Optionally solved in two stages (first n_first columns, then all) to exercise the warm start.
int dwhost_master_solve(int m, int n, const int *colptr, const int *idx, const double *val, const double *cost,
                        const double *ub, const double *rhs, int n_first, double *x, double *y, double *obj, long long *iters)
{
    dwh::MasterLP M(m);
    for (int i = 0; i < m; ++i) M.set_rhs(i, rhs[i]);
    int rc = 0;
    auto add = [&](int j) {
        std::vector<int> ix(idx + colptr[j], idx + colptr[j + 1]);
        std::vector<double> vv(val + colptr[j], val + colptr[j + 1]);
        M.add_col("c" + std::to_string(j), cost[j], ub[j] >= 1e300 ? dwh::kInf : ub[j], ix, vv);
    };
    for (int j = 0; j < n_first && j < n; ++j) add(j);
    if (n_first > 0 && n_first < n) rc = M.solve();
    for (int j = (n_first > 0 ? n_first : 0); j < n; ++j) add(j);
    rc = M.solve();
    for (int j = 0; j < n; ++j) x[j] = M.col_prim(j);
    for (int i = 0; i < m; ++i) y[i] = M.row_dual(i);
    *obj = M.obj_val();
    *iters = M.it_cnt();
    return rc;
}

// The same through the GUB master: rows 0..L-1 are linking rows, rows L..L+K-1 convexity rows (every column has at
// most one entry there, coefficient 1).  A is given by columns over all L+K rows.
int dwhost_gub_solve(int L, int K, int n, const int *colptr, const int *idx, const double *val, const double *cost,
                     const double *rhs, int n_first, double *x, double *y, double *obj, long long *iters)
{
    dwh::GubMaster M(L, K);
    for (int i = 0; i < L + K; ++i) M.set_rhs(i, rhs[i]);
    auto add = [&](int j) {
        std::vector<int> ix(idx + colptr[j], idx + colptr[j + 1]);
        std::vector<double> vv(val + colptr[j], val + colptr[j + 1]);
        M.add_col("c" + std::to_string(j), cost[j], dwh::kInf, ix, vv);
    };
    int rc = 0;
    for (int j = 0; j < n_first && j < n; ++j) add(j);
    if (n_first > 0 && n_first < n) rc = M.solve();
    for (int j = (n_first > 0 ? n_first : 0); j < n; ++j) add(j);
    rc = M.solve();
    for (int j = 0; j < n; ++j) x[j] = M.col_prim(j);
    for (int i = 0; i < L + K; ++i) y[i] = M.row_dual(i);
    *obj = M.obj_val();
    *iters = M.it_cnt();
    return rc;
}

}  // extern "C"

// ---- incremental handle API over IMaster: lets the Python DW harness (tests/dw_loop.py) drive the product's master
// LP with the CPU oracle as the subproblem engine (no GPU needed), e.g. to profile the master at config B scale
extern "C" {

void *dwhost_im_new(int L, int K, int gub)
{
    if (gub) return new dwh::GubMaster(L, K);
    return new dwh::MasterLP(L + K);
}
void dwhost_im_free(void *h) { delete static_cast<dwh::IMaster *>(h); }
void dwhost_im_set_rhs(void *h, int i, double b) { static_cast<dwh::IMaster *>(h)->set_rhs(i, b); }
int dwhost_im_add_col(void *h, double cost, double ub, int nnz, const int *idx, const double *val)
{
    return static_cast<dwh::IMaster *>(h)->add_col("c", cost, ub >= 1e300 ? dwh::kInf : ub, std::vector<int>(idx, idx + nnz),
                                                   std::vector<double>(val, val + nnz));
}
void dwhost_im_set_cost(void *h, int j, double c) { static_cast<dwh::IMaster *>(h)->set_cost(j, c); }
void dwhost_im_set_coef(void *h, int j, int pos, double v) { static_cast<dwh::IMaster *>(h)->set_coef(j, pos, v); }
void dwhost_im_del_cols(void *h, int n, const char *flag)
{
    static_cast<dwh::IMaster *>(h)->del_cols(std::vector<char>(flag, flag + n));
}
int dwhost_im_solve(void *h, double *obj, double *duals, long long *iters)
{
    dwh::IMaster *M = static_cast<dwh::IMaster *>(h);
    const int rc = M->solve();
    *obj = M->obj_val();
    for (int i = 0; i < M->rows(); ++i) duals[i] = M->row_dual(i);
    *iters = M->it_cnt();
    return rc;
}
void dwhost_im_primal(void *h, double *x)
{
    dwh::IMaster *M = static_cast<dwh::IMaster *>(h);
    for (int j = 0; j < M->cols(); ++j) x[j] = M->col_prim(j);
}

// Test hooks (tests/test_host_library.py, no GPU needed): sizes and order-sensitive checksums of everything the engine
// would be given — out[0..7] = L, sum m, sum n, nnz(A), nnz(D), checksum(A), checksum(D), checksum(bounds and costs) —
// for an instance read from LP text with `readers` threads, and for one loaded from a container.
int dwhost_read_problem_digest(const char *master_file, const char *const *subproblem_files, int K, int readers, double *out)
{
    try {
        dwh::LpModel master;
        std::vector<dwh::BlockHost> blocks;
        dwh::read_problem(master_file, subproblem_files, K, readers, master, blocks);
        std::vector<dwgpu_block_desc> desc;
        dwh::describe(blocks, desc);
        dwh::digest_blocks(K, master.m(), desc.data(), out);
        return 0;
    } catch (const std::exception &) { return 1; }
}

int dwhost_container_digest(const char *container_path, double *out, int *K_out, char *first_name, int first_name_cap)
{
    try {
        dwh::ContainerView v;
        dwh::load_container(container_path, v);
        dwh::digest_blocks(v.K, v.L, v.desc.data(), out);
        if (K_out) *K_out = v.K;
        if (first_name && first_name_cap > 0) {
            const std::string nm = v.has_names && !v.names[0].empty() ? v.names[0][0] : std::string();
            std::snprintf(first_name, (size_t)first_name_cap, "%s", nm.c_str());
        }
        return 0;
    } catch (const std::exception &) { return 1; }
}

}  // extern "C"
