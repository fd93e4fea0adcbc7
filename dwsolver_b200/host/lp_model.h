// lp_model.h — in-memory LP as read from a CPLEX-LP text file, and the block-angular assembly.
#pragma once
#include <limits>
#include <string>
#include <unordered_map>
#include <vector>

namespace dwh {

constexpr double kInf = std::numeric_limits<double>::infinity();

struct LpModel {
    std::vector<std::string> col_names, row_names;
    std::unordered_map<std::string, int> col_index;
    std::vector<double> obj;            // n
    double obj_const = 0.0;
    bool maximize = false;
    std::vector<double> col_lb, col_ub; // n  (default [0, +inf))
    std::vector<double> row_lb, row_ub; // m
    // coefficients in file order
    std::vector<int> a_row, a_col;
    std::vector<double> a_val;
    int n() const { return (int)col_names.size(); }
    int m() const { return (int)row_names.size(); }
};

// Reads the CPLEX-LP subset glp_read_lp accepts that the reference's fixtures use:
// objective (min/max), constraints (<=, >=, =, also =<, =>, <, >), bounds (l <= x <= u, x >= l, x <= u,
// x = v, x free, +-inf), general/integer/binary sections (kinds are ignored, binary sets [0,1]), end.
// Throws std::runtime_error with a message on syntax errors / unreadable files.
LpModel read_lp_file(const std::string &path);
LpModel parse_lp_text(const std::string &text);

}  // namespace dwh
