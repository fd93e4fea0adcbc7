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

// what the master loop needs from an LP object (the glp_* calls of the reference)
class IMaster {
public:
    virtual ~IMaster() {}
    virtual int rows() const = 0;
    virtual int cols() const = 0;
    virtual void set_rhs(int i, double b) = 0;
    virtual double rhs(int i) const = 0;
    // returns the column index; the new column is nonbasic at 0
    virtual int add_col(const std::string &name, double cost, double ub, const std::vector<int> &idx,
                        const std::vector<double> &val) = 0;
    virtual void set_cost(int j, double c) = 0;
    virtual double cost(int j) const = 0;
    virtual void set_coef(int j, int pos, double v) = 0;
    virtual const std::string &name(int j) const = 0;
    virtual const std::vector<int> &col_idx(int j) const = 0;
    virtual const std::vector<double> &col_val(int j) const = 0;
    // delete the columns whose flag is set; the basis is repaired with logicals (glp_del_cols + glp_adv_basis)
    virtual void del_cols(const std::vector<char> &flag) = 0;
    virtual void set_obj_const(double c0) = 0;
    // 0 = optimal, 1 = infeasible, 2 = unbounded, 3 = iteration limit / numerical failure
    virtual int solve() = 0;
    virtual double obj_val() const = 0;
    virtual double row_dual(int i) const = 0;
    virtual double col_prim(int j) const = 0;
    virtual long long it_cnt() const = 0;              // cumulative, like glp_get_it_cnt
};

class MasterLP : public IMaster {
public:
    explicit MasterLP(int rows);
    int rows() const override { return m_; }
    int cols() const override { return (int)cost_.size(); }
    void set_rhs(int i, double b) override { b_[i] = b; xb_valid_ = false; }
    double rhs(int i) const override { return b_[i]; }
    int add_col(const std::string &name, double cost, double ub, const std::vector<int> &idx,
                const std::vector<double> &val) override;
    void set_cost(int j, double c) override { cost_[j] = c; }
    double cost(int j) const override { return cost_[j]; }
    void set_coef(int j, int pos, double v) override { val_[j][pos] = v; factor_valid_ = false; }
    const std::string &name(int j) const override { return name_[j]; }
    const std::vector<int> &col_idx(int j) const override { return idx_[j]; }
    const std::vector<double> &col_val(int j) const override { return val_[j]; }
    double col_ub(int j) const { return ub_[j]; }
    void del_cols(const std::vector<char> &flag) override;
    void set_obj_const(double c0) override { c0_ = c0; }
    int solve() override;
    double obj_val() const override { return obj_; }
    double row_dual(int i) const override { return y_[i]; }
    double col_prim(int j) const override { return x_[j]; }
    long long it_cnt() const override { return iters_; }

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
