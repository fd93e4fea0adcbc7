// ipm_master.h — restricted master solved on the device by the engine library's interior-point solver
// (include/dwgpu.h, dwgpu_master_*; csrc/master_ipm.cu) behind the same IMaster interface as the simplex masters.
//
// Why: at config B (K = 8192 blocks, L = 256 linking rows) the host simplex master needs ~250 k iterations over the
// ~12 master solves of a DW run (5 s); the subproblems need 30 ms.  An interior-point solve costs ~25 iterations of
// massively parallel work per master, whatever the number of new columns.  The reference's call being replaced:
// glp_simplex(master_lp) at /root/reference/src/dw_phases.c:219, 450 and src/dw_solver.c:561-562.
//
// What the DW loop gets back:
//   * col_prim / obj_val: the interior-point primal solution (optimal to the solver's tolerance; not a vertex);
//   * row duals: pi = the interior-point duals of the linking rows, clipped to the sign the slack columns allow, and
//     r_k RECOMPUTED as min over the block's columns of (c_j - pi'a_j).  (pi, r) is then exactly dual feasible for the
//     restricted master, so the acceptance test r_k - obj_k > 1e-6 (src/dw_phases.c:108, 358) accepts a proposal iff
//     it beats every column the block already has at these prices — independent of the solver's last digits.
// If the device solver is unavailable or fails to converge, the GUB simplex takes over (same columns, its own basis).
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "../../include/dwgpu.h"
#include "gub_master.h"
#include "master_lp.h"

namespace dwh {

class IpmMaster : public IMaster {
public:
    // returns nullptr when the engine library has no usable device master
    static IpmMaster *create(int device, int linking_rows, int blocks);
    ~IpmMaster() override;
    int rows() const override { return L_ + K_; }
    int cols() const override { return (int)cost_.size(); }
    void set_rhs(int i, double b) override;
    double rhs(int i) const override { return b_[i]; }
    int add_col(const std::string &name, double cost, double ub, const std::vector<int> &idx,
                const std::vector<double> &val) override;
    void set_cost(int j, double c) override;
    double cost(int j) const override { return cost_[j]; }
    void set_coef(int j, int pos, double v) override;
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
    double device_seconds() const { return dev_seconds_; }

private:
    IpmMaster(dwgpu_master *h, int L, int K);
    int L_, K_;
    dwgpu_master *h_;
    std::vector<double> b_, cost_, x_, y_;
    std::vector<std::string> name_;
    std::vector<std::vector<int>> idx_;
    std::vector<std::vector<double>> val_;
    int synced_ = 0;                 // columns the device model holds
    bool rebuild_ = true;            // device model must be rebuilt from scratch (rhs or a coefficient changed)
    bool costs_dirty_ = false, changed_ = true;
    int last_status_ = 3;
    double c0_ = 0.0, obj_ = 0.0, dev_seconds_ = 0.0;
    long long iters_ = 0;
    std::unique_ptr<GubMaster> fallback_;        // takes over after a device-solver failure
    int solve_fallback();
    int push_columns(int first);
};

}  // namespace dwh
