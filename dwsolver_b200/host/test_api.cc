// test_api.cc — C entry points that let the (Python) test-suite exercise the host pieces in isolation:
// the LP reader against the package's own reader, the master LP solver against HiGHS.
#include <cstring>
#include <stdexcept>
#include <vector>

#include "lp_model.h"
#include "master_lp.h"

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

}  // extern "C"
