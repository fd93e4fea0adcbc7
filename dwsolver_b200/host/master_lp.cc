// master_lp.cc — bounded-variable revised primal simplex for the restricted master (see master_lp.h).
#include "master_lp.h"

#include <algorithm>
#include <cmath>
#include <limits>

namespace dwh {
namespace {
constexpr double kTolDj = 1e-9, kTolBnd = 1e-9, kTolPiv = 1e-9, kTolTie = 1e-12;
constexpr double kInfD = std::numeric_limits<double>::infinity();
}  // namespace

MasterLP::MasterLP(int rows) : m_(rows), b_(rows, 0.0), head_(rows), logical_basic_(rows, 1), y_(rows, 0.0)
{
    for (int i = 0; i < rows; ++i) head_[i] = -(i + 1);
}

int MasterLP::add_col(const std::string &name, double cost, double ub, const std::vector<int> &idx,
                      const std::vector<double> &val)
{
    name_.push_back(name);
    cost_.push_back(cost);
    ub_.push_back(ub);
    idx_.push_back(idx);
    val_.push_back(val);
    stat_.push_back(AT_LB);
    x_.push_back(0.0);
    return cols() - 1;
}

void MasterLP::del_cols(const std::vector<char> &flag)
{
    const int n = cols();
    std::vector<int> remap(n, -1);
    int w = 0;
    for (int j = 0; j < n; ++j) {
        if (flag[j]) continue;
        remap[j] = w;
        if (w != j) {
            name_[w] = std::move(name_[j]); cost_[w] = cost_[j]; ub_[w] = ub_[j];
            idx_[w] = std::move(idx_[j]); val_[w] = std::move(val_[j]); stat_[w] = stat_[j]; x_[w] = x_[j];
        }
        ++w;
    }
    name_.resize(w); cost_.resize(w); ub_.resize(w); idx_.resize(w); val_.resize(w); stat_.resize(w); x_.resize(w);
    // the basic set now lacks the deleted basic columns: refactor() fills the holes with logicals
    // (the reference rebuilds the basis with glp_adv_basis after deleting the artificials,
    //  /root/reference/src/dw_solver.c:535-562)
    factor_valid_ = false;
    xb_valid_ = false;
}

void MasterLP::refactor()
{
    const int m = m_, n = cols();
    binv_.assign((size_t)m * m, 0.0);
    for (int i = 0; i < m; ++i) { binv_[(size_t)i * m + i] = 1.0; head_[i] = -(i + 1); }
    std::vector<char> keep_logical = logical_basic_;          // logicals that are meant to stay basic
    std::vector<double> alpha(m);
    int placed = 0, wanted = 0;
    for (int j = 0; j < n; ++j) {
        if (stat_[j] != BASIC) continue;
        ++wanted;
        std::fill(alpha.begin(), alpha.end(), 0.0);
        for (size_t t = 0; t < idx_[j].size(); ++t) {
            const int k = idx_[j][t];
            const double v = val_[j][t];
            if (v == 0.0) continue;
            for (int i = 0; i < m; ++i) alpha[i] += binv_[(size_t)i * m + k] * v;
        }
        int p = -1;
        double best = 1e-9;
        for (int i = 0; i < m; ++i) {
            if (head_[i] >= 0) continue;                       // already holds a structural
            const int r = -head_[i] - 1;
            if (keep_logical[r]) continue;                     // this logical is meant to stay
            if (std::fabs(alpha[i]) > best) { best = std::fabs(alpha[i]); p = i; }
        }
        if (p < 0) {                                           // second chance: displace a kept logical
            for (int i = 0; i < m; ++i) {
                if (head_[i] >= 0) continue;
                if (std::fabs(alpha[i]) > best) { best = std::fabs(alpha[i]); p = i; }
            }
            if (p >= 0) keep_logical[-head_[p] - 1] = 0;
        }
        if (p < 0) { stat_[j] = AT_LB; continue; }            // dependent column: drop it from the basis
        const double inv = 1.0 / alpha[p];
        double *rp = &binv_[(size_t)p * m];
        for (int k = 0; k < m; ++k) rp[k] *= inv;
        for (int i = 0; i < m; ++i) {
            if (i == p || alpha[i] == 0.0) continue;
            const double a = alpha[i];
            double *ri = &binv_[(size_t)i * m];
            for (int k = 0; k < m; ++k) ri[k] -= a * rp[k];
        }
        head_[p] = j;
        ++placed;
    }
    (void)wanted; (void)placed;
    std::fill(logical_basic_.begin(), logical_basic_.end(), 0);
    for (int i = 0; i < m; ++i) if (head_[i] < 0) logical_basic_[-head_[i] - 1] = 1;
    factor_valid_ = true;
    xb_valid_ = false;
}

void MasterLP::compute_xb()
{
    const int m = m_, n = cols();
    std::vector<double> rhs = b_;
    for (int j = 0; j < n; ++j)
        if (stat_[j] == AT_UB && ub_[j] < kInfD)
            for (size_t t = 0; t < idx_[j].size(); ++t) rhs[idx_[j][t]] -= val_[j][t] * ub_[j];
    xb_.assign(m, 0.0);
    for (int i = 0; i < m; ++i) {
        const double *ri = &binv_[(size_t)i * m];
        double s = 0.0;
        for (int k = 0; k < m; ++k) s += ri[k] * rhs[k];
        xb_[i] = s;
    }
    xb_valid_ = true;
}

int MasterLP::solve()
{
    const int m = m_;
    if (!factor_valid_) refactor();
    if (!xb_valid_) compute_xb();
    std::vector<double> y(m), alpha(m), w(m), cb(m), cb_used(m, 0.0);
    bool y_valid = false, y_fresh = false;
    int degen_run = 0, refactors = 0;
    bool bland = false;
    const long long it_limit = iters_ + 200LL * (m + cols()) + 20000LL;
    int status = 3;
    for (;;) {
        if (iters_ > it_limit) { status = 3; break; }
        const int n = cols();
        // ---- phase: weights of the infeasible basic variables
        bool phase1 = false;
        for (int i = 0; i < m; ++i) {
            const double up = var_ub(head_[i]);
            double wi = 0.0;
            if (xb_[i] < -kTolBnd) wi = -1.0;
            else if (xb_[i] > up + kTolBnd * (1.0 + std::fabs(up))) wi = 1.0;
            w[i] = wi;
            phase1 |= wi != 0.0;
        }
        // ---- duals y = c_B' B^-1  (phase 1: c_B = w).  A basis change updates y in O(m) (below); the
        // O(m^2) product is only redone when some other basic cost changed (phase switch, a phase-1
        // weight flipping) or after a refactorisation.
        bool same = y_valid;
        for (int i = 0; i < m; ++i) {
            cb[i] = phase1 ? w[i] : var_cost(head_[i]);
            same = same && cb[i] == cb_used[i];
        }
        if (!same) {
            std::fill(y.begin(), y.end(), 0.0);
            for (int i = 0; i < m; ++i) {
                if (cb[i] == 0.0) continue;
                const double *ri = &binv_[(size_t)i * m];
                const double c = cb[i];
                for (int k = 0; k < m; ++k) y[k] += c * ri[k];
            }
            cb_used = cb;
            y_valid = true;
            y_fresh = true;
        }
        // ---- pricing
        int q = -1;
        double best = 0.0, dq = 0.0;
        for (int j = 0; j < n; ++j) {
            const char st = stat_[j];
            if (st == BASIC || !(ub_[j] > 0.0)) continue;
            double d = phase1 ? 0.0 : cost_[j];
            const std::vector<int> &ix = idx_[j];
            const std::vector<double> &vv = val_[j];
            for (size_t t = 0; t < ix.size(); ++t) d -= y[ix[t]] * vv[t];
            const bool ok = (st == AT_LB && d < -kTolDj) || (st == AT_UB && d > kTolDj);
            if (!ok) continue;
            if (bland) { q = j; dq = d; break; }                 // lowest index
            if (std::fabs(d) > best) { best = std::fabs(d); q = j; dq = d; }
        }
        if (q < 0) {
            if (!y_fresh) { y_valid = false; continue; }      // confirm with duals computed from scratch
            if (phase1) { status = 1; break; }
            // optimal: verify A x = b, refactor once or twice if the inverse has drifted
            std::vector<double> act(m, 0.0);
            for (int j = 0; j < n; ++j) x_[j] = stat_[j] == AT_UB ? ub_[j] : 0.0;
            for (int i = 0; i < m; ++i) if (head_[i] >= 0) x_[head_[i]] = xb_[i];
            for (int j = 0; j < n; ++j)
                if (x_[j] != 0.0) for (size_t t = 0; t < idx_[j].size(); ++t) act[idx_[j][t]] += val_[j][t] * x_[j];
            double worst = 0.0;
            for (int r = 0; r < m; ++r) if (head_[r] < 0) act[-head_[r] - 1] += xb_[r];       // basic logicals
            for (int i = 0; i < m; ++i) worst = std::max(worst, std::fabs(act[i] - b_[i]) / (1.0 + std::fabs(b_[i])));
            if (worst > 1e-9 && refactors < 3) { ++refactors; refactor(); compute_xb(); y_valid = false; continue; }
            status = 0;
            break;
        }
        const double s = stat_[q] == AT_LB ? 1.0 : -1.0;          // entering variable increases / decreases
        // ---- FTRAN
        std::fill(alpha.begin(), alpha.end(), 0.0);
        for (size_t t = 0; t < idx_[q].size(); ++t) {
            const int k = idx_[q][t];
            const double v = val_[q][t];
            if (v == 0.0) continue;
            for (int i = 0; i < m; ++i) alpha[i] += binv_[(size_t)i * m + k] * v;
        }
        // ---- ratio test (composite phase 1: a violated bound is the breakpoint of an infeasible variable)
        int p = -1;
        double tmin = kInfD, pmag = 0.0;
        bool p_to_upper = false;
        for (int i = 0; i < m; ++i) {
            const double rate = -s * alpha[i];                    // d x_Bi / dt
            if (std::fabs(rate) <= kTolPiv) continue;
            const double up = var_ub(head_[i]), xi = xb_[i];
            const bool below = xi < -kTolBnd, above = xi > up + kTolBnd * (1.0 + std::fabs(up));
            double t;
            bool to_upper;
            if (rate > 0.0) {
                if (below) { t = -xi / rate; to_upper = false; }
                else if (above) continue;
                else { if (!(up < kInfD)) continue; t = (up - xi) / rate; to_upper = true; }
            } else {
                if (above) { t = (up - xi) / rate; to_upper = true; }
                else if (below) continue;
                else { t = -xi / rate; to_upper = false; }
            }
            if (t < 0.0) t = 0.0;
            bool take = false;
            if (t < tmin - kTolTie) take = true;
            else if (t <= tmin + kTolTie) take = bland ? (p < 0 || head_[i] < head_[p]) : std::fabs(rate) > pmag;
            if (take) { if (t < tmin) tmin = t; p = i; pmag = std::fabs(rate); p_to_upper = to_upper; }
        }
        const double range = ub_[q];
        if (range <= tmin) {
            if (!(range < kInfD)) { status = phase1 ? 3 : 2; break; }
            for (int i = 0; i < m; ++i) xb_[i] -= s * range * alpha[i];
            stat_[q] = stat_[q] == AT_LB ? AT_UB : AT_LB;
            ++iters_;
            degen_run = 0; bland = false;
            continue;
        }
        if (p < 0) { status = phase1 ? 3 : 2; break; }
        // ---- basis change
        const double xq_new = (stat_[q] == AT_LB ? 0.0 : ub_[q]) + s * tmin;
        for (int i = 0; i < m; ++i) xb_[i] -= s * tmin * alpha[i];
        xb_[p] = xq_new;
        const int vleave = head_[p];
        if (vleave >= 0) stat_[vleave] = p_to_upper ? AT_UB : AT_LB;
        else logical_basic_[-vleave - 1] = 0;
        stat_[q] = BASIC;
        head_[p] = q;
        {
            const double inv = 1.0 / alpha[p];
            double *rp = &binv_[(size_t)p * m];
            for (int k = 0; k < m; ++k) rp[k] *= inv;
            for (int i = 0; i < m; ++i) {
                if (i == p || alpha[i] == 0.0) continue;
                const double a = alpha[i];
                double *ri = &binv_[(size_t)i * m];
                for (int k = 0; k < m; ++k) ri[k] -= a * rp[k];
            }
            // y' = y + d_q * (new row p of B^-1); the entering variable's basic cost is what it was priced with
            for (int k = 0; k < m; ++k) y[k] += dq * rp[k];
            y_fresh = false;
            cb_used[p] = phase1 ? 0.0 : cost_[q];
        }
        ++iters_;
        if (tmin <= 1e-11) { if (++degen_run > 200) bland = true; }
        else { degen_run = 0; bland = false; }
    }
    // ---- publish
    const int n = cols();
    for (int j = 0; j < n; ++j) x_[j] = stat_[j] == AT_UB ? ub_[j] : 0.0;
    for (int i = 0; i < m; ++i) if (head_[i] >= 0) x_[head_[i]] = xb_[i];
    obj_ = c0_;
    for (int j = 0; j < n; ++j) obj_ += cost_[j] * x_[j];
    std::fill(y_.begin(), y_.end(), 0.0);
    for (int i = 0; i < m; ++i) {
        const double cb = var_cost(head_[i]);
        if (cb == 0.0) continue;
        const double *ri = &binv_[(size_t)i * m];
        for (int k = 0; k < m; ++k) y_[k] += cb * ri[k];
    }
    return status;
}

}  // namespace dwh
