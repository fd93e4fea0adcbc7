// ipm_master.cc — see ipm_master.h.
#include "ipm_master.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <limits>

namespace dwh {

IpmMaster *IpmMaster::create(int device, int linking_rows, int blocks)
{
    dwgpu_master *h = nullptr;
    if (linking_rows < 1 || blocks < 1) return nullptr;
    if (dwgpu_master_create(device, linking_rows, blocks, &h) != 0 || !h) return nullptr;
    return new IpmMaster(h, linking_rows, blocks);
}

IpmMaster::IpmMaster(dwgpu_master *h, int L, int K) : L_(L), K_(K), h_(h), b_((size_t)L + K, 0.0), y_((size_t)L + K, 0.0) {}

IpmMaster::~IpmMaster()
{
    if (h_) dwgpu_master_destroy(h_);
}

void IpmMaster::set_rhs(int i, double b)
{
    b_[i] = b;
    rebuild_ = true; changed_ = true;
    if (fallback_) fallback_->set_rhs(i, b);
}

int IpmMaster::add_col(const std::string &name, double cost, double ub, const std::vector<int> &idx, const std::vector<double> &val)
{
    name_.push_back(name); cost_.push_back(cost); idx_.push_back(idx); val_.push_back(val);
    x_.push_back(0.0);
    changed_ = true;
    if (fallback_) fallback_->add_col(name, cost, ub, idx, val);
    return cols() - 1;
}

void IpmMaster::set_cost(int j, double c)
{
    if (cost_[j] != c) { cost_[j] = c; costs_dirty_ = true; changed_ = true; }
    if (fallback_) fallback_->set_cost(j, c);
}

void IpmMaster::set_coef(int j, int pos, double v)
{
    val_[j][pos] = v;
    rebuild_ = true; changed_ = true;
    if (fallback_) fallback_->set_coef(j, pos, v);
}

void IpmMaster::del_cols(const std::vector<char> &flag)
{
    const int n = cols();
    int w = 0;
    for (int j = 0; j < n; ++j) {
        if (flag[j]) continue;
        if (w != j) { name_[w] = std::move(name_[j]); cost_[w] = cost_[j]; idx_[w] = std::move(idx_[j]); val_[w] = std::move(val_[j]); x_[w] = x_[j]; }
        ++w;
    }
    name_.resize(w); cost_.resize(w); idx_.resize(w); val_.resize(w); x_.resize(w);
    rebuild_ = true; changed_ = true;
    if (fallback_) fallback_->del_cols(flag);
}

int IpmMaster::push_columns(int first)
{
    const int n = cols();
    if (first >= n) return 0;
    std::vector<int> colptr{0}, rows;
    std::vector<double> vals;
    for (int j = first; j < n; ++j) {
        rows.insert(rows.end(), idx_[j].begin(), idx_[j].end());
        vals.insert(vals.end(), val_[j].begin(), val_[j].end());
        colptr.push_back((int)rows.size());
    }
    return dwgpu_master_add_cols(h_, n - first, colptr.data(), rows.data(), vals.data(), cost_.data() + first);
}

int IpmMaster::solve_fallback()
{
    if (!fallback_) {
        fallback_.reset(new GubMaster(L_, K_));
        for (int i = 0; i < L_ + K_; ++i) fallback_->set_rhs(i, b_[i]);
        for (int j = 0; j < cols(); ++j) fallback_->add_col(name_[j], cost_[j], std::numeric_limits<double>::infinity(), idx_[j], val_[j]);
    }
    fallback_->set_obj_const(c0_);
    const int rc = fallback_->solve();
    obj_ = fallback_->obj_val();
    for (int i = 0; i < L_ + K_; ++i) y_[i] = fallback_->row_dual(i);
    for (int j = 0; j < cols(); ++j) x_[j] = fallback_->col_prim(j);
    iters_ += 1;            // the DW loop's progress test only needs the counter to move
    changed_ = false;
    last_status_ = rc;
    return rc;
}

int IpmMaster::solve()
{
    if (!changed_) return last_status_;            // same LP as last time: no work, the iteration counter stands still
    if (fallback_) return solve_fallback();
    const int n = cols();
    int rc = 0;
    auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();
    if (rebuild_) {
        std::vector<char> all((size_t)std::max(dwgpu_master_cols(h_), 1), 1);
        if (dwgpu_master_cols(h_) > 0) rc = dwgpu_master_del_cols(h_, all.data());
        if (!rc) rc = dwgpu_master_set_rhs(h_, b_.data());
        if (!rc) rc = push_columns(0);
        costs_dirty_ = false;
    } else {
        if (costs_dirty_ && synced_ > 0) rc = dwgpu_master_set_costs(h_, 0, synced_, cost_.data());
        if (!rc) rc = push_columns(synced_);
        costs_dirty_ = false;
    }
    if (rc) { std::fprintf(stderr, "dwsolver-b200: device master: %s; the simplex master takes over\n", dwgpu_master_last_error()); return solve_fallback(); }
    synced_ = n; rebuild_ = false;
    const double t_push = now();
    const char *te = std::getenv("DWSOLVER_IPM_TOL");
    const double tol = te ? std::atof(te) : 1e-9;
    int st = 3;
    rc = dwgpu_master_solve(h_, tol, 100, &st);
    double pinf = 1.0, dinf = 1.0, gap = 1.0;
    if (!rc) dwgpu_master_quality(h_, &pinf, &dinf, &gap);
    // accepted when the returned iterate is a good LP solution even if the last digits were out of reach: the duals are
    // made exactly dual feasible below and the DW loop bounds what inexact prices can cost (they only delay columns)
    const bool usable = !rc && (st == 0 || std::max(pinf, std::max(dinf, gap)) <= 1e-6);
    if (!usable) {
        std::fprintf(stderr, "dwsolver-b200: device master did not converge (%s); the simplex master takes over\n",
                     rc ? dwgpu_master_last_error() : "iteration limit");
        return solve_fallback();
    }
    const double t_solved = now();
    double pobj = 0.0, dobj = 0.0, secs = 0.0;
    int its = 0;
    x_.assign((size_t)n, 0.0);
    dwgpu_master_get(h_, &pobj, &dobj, y_.data(), x_.data(), &its, &secs);
    dev_seconds_ += secs;
    iters_ += std::max(its, 1);
    for (int j = 0; j < n; ++j) if (x_[j] < 1e-11) x_[j] = 0.0;          // the interior solution's dust
    // pi within the signs the block-less singleton columns (slacks, artificials) allow: c_j - a pi_i >= 0
    for (int j = 0; j < n; ++j) {
        if (idx_[j].size() != 1 || idx_[j][0] >= L_) continue;
        const int i = idx_[j][0];
        const double a = val_[j][0];
        if (a > 0.0) y_[i] = std::min(y_[i], cost_[j] / a);
        else if (a < 0.0) y_[i] = std::max(y_[i], cost_[j] / a);
    }
    // r_k = min over the block's columns of the reduced cost at pi: (pi, r) exactly dual feasible
    std::vector<double> r((size_t)K_, std::numeric_limits<double>::infinity());
    for (int j = 0; j < n; ++j) {
        int k = -1;
        double d = cost_[j];
        for (size_t t = 0; t < idx_[j].size(); ++t) {
            if (idx_[j][t] >= L_) k = idx_[j][t] - L_;
            else d -= y_[idx_[j][t]] * val_[j][t];
        }
        if (k >= 0 && d < r[(size_t)k]) r[(size_t)k] = d;
    }
    for (int k = 0; k < K_; ++k) if (std::isfinite(r[(size_t)k])) y_[(size_t)L_ + k] = r[(size_t)k];
    obj_ = pobj + c0_;
    changed_ = false;
    last_status_ = 0;
    if (std::getenv("DWSOLVER_DEBUG_TIMING"))
        std::fprintf(stderr, "dwsolver-b200: device master: %d columns, %d iterations, quality %.1e/%.1e/%.1e; push %.4f s, solve call %.4f s (device %.4f s), duals %.4f s\n",
                     n, its, pinf, dinf, gap, t_push - t_begin, t_solved - t_push, secs, now() - t_solved);
    return 0;
}

}  // namespace dwh
