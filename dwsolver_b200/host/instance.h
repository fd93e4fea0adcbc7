// instance.h — a block-angular instance in host memory, and the two ways it gets there: the reference's LP text
// (master + K block files, /root/reference/src/dw_solver.c:192-198, src/dw_subprob.c:105-111, 203-258) and this
// build's binary container.  Both end in the engine's block descriptors (include/dwgpu.h).
#pragma once
#include <string>
#include <vector>

#include "../../include/dwgpu.h"
#include "lp_model.h"

namespace dwh {

struct BlockHost {
    int m = 0, n = 0;
    std::vector<int> a_rowptr, a_col;
    std::vector<double> a_val, row_lb, row_ub, col_lb, col_ub, c_file, c_master;
    std::vector<int> d_row, d_col;
    std::vector<double> d_val;
    std::vector<std::string> col_names;
};

// LP-file infinity (>= 1e300) -> the engine's "no bound" value
double to_engine_bound(double v);

// Master and block files -> in-memory blocks.  `readers` host threads (0 = one per core, at most 16;
// DWSOLVER_READ_THREADS overrides) parse the block files; the result does not depend on their number.
// Throws std::runtime_error on an unreadable or malformed file.
void read_problem(const char *master_file, const char *const *subproblem_files, int K, int readers, LpModel &master,
                  std::vector<BlockHost> &blocks);

// engine descriptors pointing into `blocks`
void describe(std::vector<BlockHost> &blocks, std::vector<dwgpu_block_desc> &desc);

struct ContainerView {               // a loaded container: descriptors pointing into `image`
    std::vector<char> image;
    int K = 0, L = 0;
    double shift = 0.0;
    const double *link_lb = nullptr, *link_ub = nullptr;
    std::vector<dwgpu_block_desc> desc;
    std::vector<std::vector<std::string>> names;
    bool has_names = false;
};

// throws std::runtime_error on a missing, truncated or foreign file
void load_container(const char *path, ContainerView &v);
// false when the file cannot be written
bool write_container(const char *path, const LpModel &master, const std::vector<BlockHost> &blocks, double shift);

// order-sensitive digest of everything the engine would be given (test hooks: the text path and the container path
// must agree bit for bit): out[0..7] = L, sum m, sum n, nnz(A), nnz(D), checksum(A), checksum(D), checksum(bounds, costs)
void digest_blocks(int K, int L, const dwgpu_block_desc *desc, double *out);

}  // namespace dwh
