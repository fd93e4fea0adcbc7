// master_lp.h — the restricted master problem and its LP solver.
//
// The reference keeps the master in a glp_prob and calls glp_simplex (primal, warm start) after every
// batch of new columns (/root/reference/src/dw_phases.c:219, 450; src/dw_solver.c:561-562).  GLPK is
// not available in this build, so this class provides the handful of operations the master loop uses
// (SURVEY.md §8c lists the glp_* calls): add / delete columns, change costs, solve from the current
// basis, read row duals, column values, the objective and the iteration counter.
//
// Problem form:  min c'x + c0  s.t.  A x = b,  0 <= x_j <= u_j  (all rows are equalities: the
// reference turns every linking row into GLP_FX, src/dw_solver.c:321-419, and convexity rows are = 1).
// Solver: bounded-variable revised primal simplex, explicit dense basis inverse, one fixed logical
// column per row (GLPK's auxiliary variables), composite phase 1, Dantzig pricing with Bland fallback.
#pragma once
#include <string>
#include <vector>

namespace dwh {

class MasterLP {
public:
    explicit MasterLP(int rows);
    int rows() const { return m_; }
    int cols() const { return (int)cost_.size(); }
    void set_rhs(int i, double b) { b_[i] = b; xb_valid_ = false; }
    double rhs(int i) const { return b_[i]; }
    // returns the column index; the new column is nonbasic at 0
    int add_col(const std::string &name, double cost, double ub, const std::vector<int> &idx, const std::vector<double> &val);
    void set_cost(int j, double c) { cost_[j] = c; }
    double cost(int j) const { return cost_[j]; }
    void set_coef(int j, int pos, double v) { val_[j][pos] = v; factor_valid_ = false; }
    const std::string &name(int j) const { return name_[j]; }
    const std::vector<int> &col_idx(int j) const { return idx_[j]; }
    const std::vector<double> &col_val(int j) const { return val_[j]; }
    double col_ub(int j) const { return ub_[j]; }
    // delete the columns whose flag is set; the basis is repaired with logicals (glp_del_cols + glp_adv_basis)
    void del_cols(const std::vector<char> &flag);
    void set_obj_const(double c0) { c0_ = c0; }
    // 0 = optimal, 1 = infeasible, 2 = unbounded, 3 = iteration limit / numerical failure
    int solve();
    double obj_val() const { return obj_; }
    double row_dual(int i) const { return y_[i]; }
    double col_prim(int j) const { return x_[j]; }
    long long it_cnt() const { return iters_; }        // cumulative, like glp_get_it_cnt

private:
    enum { AT_LB = 0, AT_UB = 1, BASIC = 2 };
    int m_;
    std::vector<double> b_;
    std::vector<std::string> name_;
    std::vector<double> cost_, ub_;
    std::vector<std::vector<int>> idx_;
    std::vector<std::vector<double>> val_;
    std::vector<char> stat_;            // per structural column
    // basis: head_[i] = structural j (>= 0) or logical of row r encoded as -(r+1)
    std::vector<int> head_;
    std::vector<char> logical_basic_;   // per row
    std::vector<double> binv_;          // m x m, row-major
    std::vector<double> xb_, x_, y_;
    bool factor_valid_ = false, xb_valid_ = false;
    double c0_ = 0.0, obj_ = 0.0;
    long long iters_ = 0;

    void refactor();                    // B^-1 from scratch for the current basic set (repairing it if singular)
    void compute_xb();
    double var_ub(int v) const { return v >= 0 ? ub_[v] : 0.0; }   // logicals are fixed at 0
    double var_cost(int v) const { return v >= 0 ? cost_[v] : 0.0; }
};

}  // namespace dwh
