// gub_master.h — restricted master with generalised upper bounding (Dantzig & Van Slyke).
//
// The DW master has L linking rows and K convexity rows `sum_{j in block k} lambda_j = 1`; each lambda
// column appears in exactly one convexity row with coefficient 1.  A basis of the (L+K)-row problem always
// holds at least one column of every block; declaring one of them the block's KEY column and subtracting it
// from the block's other basic columns leaves an L x L WORKING basis, so one simplex iteration costs
// O(L^2 + pricing) instead of O((L+K)^2): 256^2 instead of 8448^2 at config B.  Same role and interface as
// MasterLP (master_lp.h), i.e. the glp_* operations the reference's master loop uses
// (/root/reference/src/dw_phases.c:219, 450; src/dw_solver.c:259-432, 535-562).
//
// Problem form: min c'x + c0, linking rows A x = b (equalities), convexity rows with right-hand side g_k,
// x >= 0 (the reference's bound lambda <= 1 is implied by the convexity row and is dropped), one fixed
// logical per linking row as a basis placeholder.  Primal simplex with composite phase 1, multiple pricing
// (a candidate list re-priced between full passes), periodic refactorisation of the working inverse.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "master_lp.h"
#include "thread_team.h"

namespace dwh {

class GubMaster : public IMaster {
public:
    GubMaster(int linking_rows, int blocks);
    int rows() const override { return L_ + K_; }
    int cols() const override { return (int)cost_.size(); }
    void set_rhs(int i, double b) override { b_[i] = b; vals_valid_ = false; }
    double rhs(int i) const override { return b_[i]; }
    int add_col(const std::string &name, double cost, double ub, const std::vector<int> &idx,
                const std::vector<double> &val) override;
    void set_cost(int j, double c) override { cost_[j] = c; }
    double cost(int j) const override { return cost_[j]; }
    void set_coef(int j, int pos, double v) override { val_[j][pos] = v; factor_valid_ = false; flat_valid_ = false; }
    const std::string &name(int j) const override { return name_[j]; }
    const std::vector<int> &col_idx(int j) const override { return idx_[j]; }
    const std::vector<double> &col_val(int j) const override { return val_[j]; }
    void del_cols(const std::vector<char> &flag) override;
    void set_obj_const(double c0) override { c0_ = c0; }
    int solve() override;
    double obj_val() const override { return obj_; }
    double row_dual(int i) const override { return y_[i]; }
    double col_prim(int j) const override { return x_[j]; }
    long long it_cnt() const override { return iters_; }

private:
    enum { NONBASIC = 0, NONKEY = 1, KEY = 2 };
    int L_, K_;
    std::vector<double> b_;
    std::vector<std::string> name_;
    std::vector<double> cost_;
    std::vector<std::vector<int>> idx_;
    std::vector<std::vector<double>> val_;
    // the linking-row entries of all columns, end to end in column order (what pricing walks: one contiguous stream
    // instead of two heap vectors per column); entry order inside a column = order in idx_/val_
    std::vector<int> lptr_, lrow_;
    std::vector<double> lval_;
    bool flat_valid_ = false;
    void rebuild_flat();
    std::vector<int> set_;              // block of the column, -1: none (slack / artificial)
    std::vector<char> stat_;
    std::vector<int> key_;              // per block: key column, -1: none yet
    std::vector<int> pos_var_;          // working-basis position -> column (>= 0) or -(row+1) for a logical
    std::vector<double> winv_;          // L x L
    // pending rank-1 updates of winv_ (blocked updates, DWSOLVER_MASTER_BLOCK): true inverse = winv_ - sum_j u_j v_j'
    std::vector<double> pend_u_, pend_v_;
    int pend_n_ = 0;
    std::vector<double> xn_, xkey_, x_, y_;
    bool factor_valid_ = false, vals_valid_ = false, crashed_ = false;
    double c0_ = 0.0, obj_ = 0.0;
    long long iters_ = 0;
    int scan_pos_ = 0;                  // where the next (cyclic, partial) pricing pass starts
    std::unique_ptr<ThreadTeam> team_;  // row-slice owners for the dense L x L work (created by the first large solve)

    void link_minus_key(int j, std::vector<double> &w) const;     // dense L-vector a_j - a_key(set(j))
    void add_link(int j, double scale, std::vector<double> &w) const;
    bool build_basis();                  // keys, crash, working inverse; false: some block has no column
    void refactor();
    void compute_values();
    void publish(bool with_duals);
};

}  // namespace dwh
