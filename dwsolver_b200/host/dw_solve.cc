// dw_solve.cc — the Dantzig-Wolfe driver behind the reference's public API (include/dwsolver.h).
//
// Follows the reference's control flow (all paths relative to /root/reference):
//   set-up            src/dw_solver.c:918-1016 (argument checks, option defaults), :154-184 (workers),
//                     src/dw_subprob.c:100-269 (block set-up: 3000 clamp, c_k, COO D_k by column name)
//   restricted master src/dw_solver.c:247-440 (rows made FX, su_/sl_/y_ columns, convexity rows = 1)
//   Phase I           src/dw_solver.c:461-628 + src/dw_phases.c:54-304 (prices with c = 0, first-round
//                     artificial sign fix :165-171, 201-217, stop on |obj| < 1e-6 or no column)
//   Phase II          src/dw_solver.c:649-694 + src/dw_phases.c:321-554 (accept r_k - obj_k > 1e-6,
//                     stop rule on master simplex progress :462-484)
//   reconstruction    src/dw_rounding.c:115-168, 561-597 (x_k = sum lambda_{k,t} x_k^t, relaxed_solution)
// What differs by design: the K subproblem threads, their mailboxes and the semaphore/cond-var round
// (src/dw_subprob.c:415-448, src/dw_phases.c:338-350, 526-548) are ONE synchronous call into the GPU
// engine per round (include/dwgpu.h); blocks are visited in index order, so runs are deterministic.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <utility>
#include <string>
#include <thread>
#include <vector>

#include "../../include/dwgpu.h"
#include "../../include/dwsolver.h"
#include "instance.h"
#include "lp_model.h"
#include "gub_master.h"
#include "ipm_master.h"
#include "master_lp.h"

namespace {

using namespace dwh;

constexpr double kTolerance = 1e-6;     // TOLERANCE, src/dw.h:71
constexpr double kDwInfinity = 1e6;     // DW_INFINITY, src/dw.h:72 (initial r_k)

dw_stats_t g_stats;

double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct Proposal {            // X[k][iter] of the reference (src/dw_phases.c:173-179, 427-433): kept on the device
    int k, iter;
    long long id;            // dwgpu_history_push id of the copy of x_k
};

}  // namespace

// The final restricted master as CPLEX-LP text, `done.cpxlp` (src/dw_solver.c:749-753: glp_write_lp of master_lp after
// Phase II — "solving this on its own later will provide the optimal value").  Rows: the linking rows under their
// original names, all equalities by now (src/dw_solver.c:321-419), then sub<k>_convexity = 1 (:262-265); columns: the
// slack / surplus columns and every lambda ever accepted, lambdas bounded by 1 and of integer kind like the
// reference's (src/dw_phases.c:393-400).
static bool write_final_master(const char *path, const IMaster &M, int L, int K, double obj_const,
                               const std::vector<std::string> *link_names)
{
    std::FILE *f = std::fopen(path, "w");
    if (!f) return false;
    auto row_name = [&](int i) {
        if (i >= L) return "sub" + std::to_string(i - L + 1) + "_convexity";
        if (link_names && i < (int)link_names->size() && !(*link_names)[(size_t)i].empty()) return (*link_names)[(size_t)i];
        return "r_" + std::to_string(i + 1);
    };
    const int n = M.cols();
    std::fprintf(f, "\\* Problem: master *\\\n\n");
    if (obj_const != 0.0) std::fprintf(f, "\\* Objective constant (not part of the LP text): %.15g *\\\n\n", obj_const);
    std::fprintf(f, "Minimize\n obj:");
    int on_line = 0;
    bool any = false;
    for (int j = 0; j < n; ++j) {
        const double c = M.cost(j);
        if (c == 0.0) continue;
        std::fprintf(f, " %c %.15g %s", c < 0 ? '-' : '+', std::fabs(c), M.name(j).c_str());
        any = true;
        if (++on_line % 4 == 0) std::fprintf(f, "\n");
    }
    if (!any && n > 0) std::fprintf(f, " + 0 %s", M.name(0).c_str());
    std::fprintf(f, "\n\nSubject To\n");
    // by rows: gather every column's entries
    std::vector<std::vector<std::pair<int, double>>> rows((size_t)(L + K));
    for (int j = 0; j < n; ++j) {
        const std::vector<int> &ix = M.col_idx(j);
        const std::vector<double> &vv = M.col_val(j);
        for (size_t t = 0; t < ix.size(); ++t) if (vv[t] != 0.0) rows[(size_t)ix[t]].emplace_back(j, vv[t]);
    }
    for (int i = 0; i < L + K; ++i) {
        std::fprintf(f, " %s:", row_name(i).c_str());
        on_line = 0;
        if (rows[(size_t)i].empty() && n > 0) std::fprintf(f, " + 0 %s", M.name(0).c_str());
        for (auto &e : rows[(size_t)i]) {
            std::fprintf(f, " %c %.15g %s", e.second < 0 ? '-' : '+', std::fabs(e.second), M.name(e.first).c_str());
            if (++on_line % 4 == 0) std::fprintf(f, "\n");
        }
        std::fprintf(f, " = %.15g\n", M.rhs(i));
    }
    std::fprintf(f, "\nBounds\n");
    std::vector<int> lambdas;
    for (int j = 0; j < n; ++j)
        if (M.name(j).compare(0, 7, "lambda_") == 0) { std::fprintf(f, " 0 <= %s <= 1\n", M.name(j).c_str()); lambdas.push_back(j); }
    if (!lambdas.empty()) {
        std::fprintf(f, "\nGenerals\n");
        for (int j : lambdas) std::fprintf(f, " %s\n", M.name(j).c_str());
    }
    std::fprintf(f, "\nEnd\n");
    return std::fclose(f) == 0;
}

// The solve proper, on in-memory blocks (what both dw_solve and dw_solve_blocks end up calling).
static int dw_core(int K, int L, const dwgpu_block_desc *desc_in, const double *link_lb, const double *link_ub,
                   const std::vector<std::vector<std::string>> *names, const dw_options_t &opt, double t_begin,
                   dw_result_t *result, const std::vector<std::string> *link_names = nullptr)
{
    auto fail = [&](int code) { result->status = code; return code; };
    const int verb = opt.verbosity;
    std::vector<dwgpu_block_desc> desc(desc_in, desc_in + K);
    // ---- GPU engines: replace pthread_create(subproblem_thread) x K (src/dw_solver.c:154-184) ----------
    // One engine per device; blocks are sharded statically and contiguously over them: shard g owns
    // blocks [K g / G, K (g+1) / G).  DWSOLVER_DEVICES = comma-separated CUDA ordinals (a device may be listed more than
    // once) or "all"; default: one engine on DWSOLVER_DEVICE (or 0).  Per round every shard is driven by its own
    // host thread; the duals go down to each device and the packed proposals come up from each, no peer traffic.
    std::vector<int> devices;
    if (const char *list = std::getenv("DWSOLVER_DEVICES")) {
        if (std::strcmp(list, "all") == 0) {
            const int n_dev = dwgpu_device_count();
            for (int d = 0; d < n_dev; ++d) devices.push_back(d);
        } else {
            for (const char *p = list; *p;) {
                char *endp = nullptr;
                const long d = std::strtol(p, &endp, 10);
                if (endp == p) break;
                devices.push_back((int)d);
                p = *endp == ',' ? endp + 1 : endp;
                if (*endp && *endp != ',') break;
            }
        }
    }
    if (devices.empty()) {
        const char *dev = std::getenv("DWSOLVER_DEVICE");
        devices.push_back(dev ? std::atoi(dev) : 0);
    }
    if ((int)devices.size() > K) devices.resize((size_t)K);
    const int G = (int)devices.size();
    struct Shard {
        dwgpu_engine *eng = nullptr;
        int k0 = 0, k1 = 0;          // block range
        long long x0 = 0, xn = 0;    // its slice of the block-major x vector
        dwgpu_packed pk;
        int rc = 0;
        std::string err;
        std::vector<int> accepted;   // local block ids accepted this round
        std::vector<long long> kept_ids;
        std::vector<dwgpu_packed> mpk;       // several dual vectors per round: one record per vector
        std::vector<long long> mids;         // and the history ids of every block's x_k of every pass
    };
    std::vector<Shard> shards((size_t)G);
    struct EngineGuard {
        std::vector<Shard> &s;
        ~EngineGuard() { for (auto &sh : s) if (sh.eng) dwgpu_destroy(sh.eng); }
    } guard{shards};
    {
        long long xoff = 0;
        for (int g = 0; g < G; ++g) {
            Shard &sh = shards[(size_t)g];
            sh.k0 = (int)((long long)K * g / G); sh.k1 = (int)((long long)K * (g + 1) / G);
            sh.x0 = xoff;
            for (int k = sh.k0; k < sh.k1; ++k) xoff += desc[k].n;
            sh.xn = xoff - sh.x0;
        }
    }
    // run f(shard) on every shard at once (the calls into the engine block until the shard's round is done)
    auto on_all_shards = [&](auto &&f) {
        auto body = [&](int g) {
            Shard &sh = shards[(size_t)g];
            sh.rc = f(sh);
            if (sh.rc != 0) sh.err = dwgpu_last_error();       // the error text is per thread
        };
        std::vector<std::thread> pool;
        for (int g = 1; g < G; ++g) pool.emplace_back(body, g);
        body(0);
        for (auto &th : pool) th.join();
        for (auto &sh : shards) if (sh.rc != 0) { std::fprintf(stderr, "dwsolver-b200: %s\n", sh.err.c_str()); return sh.rc; }
        return 0;
    };
    const double t_create0 = now_s();
    if (on_all_shards([&](Shard &sh) {
            dwgpu_options gopt;
            dwgpu_options_init(&gopt);
            gopt.device = devices[(size_t)(&sh - shards.data())];
            return dwgpu_create(sh.k1 - sh.k0, L, desc.data() + sh.k0, &gopt, &sh.eng);
        }) != 0) {
        std::fprintf(stderr, "dwsolver-b200: cannot create the GPU engine\n");
        return fail(DW_STATUS_ERR_INTERNAL);
    }
    const double t_create = now_s() - t_create0;
    double t_consume = 0.0, t_history = 0.0;
    auto shard_of = [&](int k) -> Shard & {
        int g = (int)(((long long)k * G) / K);                  // close; fix up at the range borders
        while (k < shards[(size_t)g].k0) --g;
        while (k >= shards[(size_t)g].k1) ++g;
        return shards[(size_t)g];
    };

    // ---- restricted master (src/dw_solver.c:247-440) ---------------------------------------------------
    // master LP: GUB working basis (L x L) by default; DWSOLVER_MASTER=dense selects the (L+K)^2 dense inverse
    // Large instances (K L >= 2^18, e.g. config B's 8192 x 256): the interior-point master on the device
    // (ipm_master.h; DWSOLVER_MASTER=ipm forces it, =gub the simplex) — ~25 parallel iterations per master instead of
    // tens of thousands of simplex pivots.
    const char *mk = std::getenv("DWSOLVER_MASTER");
    std::unique_ptr<IMaster> master_lp;
    const bool want_ipm = mk ? std::strcmp(mk, "ipm") == 0 : ((long long)K * L >= (1LL << 18));
    if (want_ipm && L > 0) master_lp.reset(IpmMaster::create(devices[0], L, K));
    const bool ipm_master = (bool)master_lp;
    if (!master_lp) {
        if (mk && std::strcmp(mk, "dense") == 0) master_lp.reset(new MasterLP(L + K));
        else master_lp.reset(new GubMaster(L, K));
    }
    if (std::getenv("DWSOLVER_DEBUG_TIMING")) std::fprintf(stderr, "dwsolver-b200: master LP solver: %s\n", ipm_master ? "interior point on the device" : "simplex on the host");
    IMaster &M = *master_lp;
    std::vector<int> art_col(L, -1);
    std::vector<char> is_art;
    auto add = [&](const std::string &name, double cost, double ub, std::vector<int> idx, std::vector<double> val, bool art) {
        const int j = M.add_col(name, cost, ub, idx, val);
        if ((int)is_art.size() <= j) is_art.resize(j + 1, 0);
        is_art[j] = art;
        return j;
    };
    for (int i = 0; i < L; ++i) {
        const double lb = link_lb[i], ub = link_ub[i];
        const std::string id = std::to_string(i + 1);
        if (lb == ub) {                                          // FX: artificial only (:321-336)
            M.set_rhs(i, ub);
            art_col[i] = add("y_" + id, 1.0, kInf, {i}, {ub < 0.0 ? -1.0 : 1.0}, true);
        } else if (lb <= -DWGPU_INF && ub < DWGPU_INF) {                 // UP: su_ (+1) and y_ (-1) (:338-375)
            M.set_rhs(i, ub);
            add("su_" + id, 0.0, kInf, {i}, {1.0}, false);
            art_col[i] = add("y_" + id, 1.0, kInf, {i}, {-1.0}, true);
        } else if (ub >= DWGPU_INF && lb > -DWGPU_INF) {                 // LO: sl_ (-1) and y_ (+1) (:381-419)
            M.set_rhs(i, lb);
            add("sl_" + id, 0.0, kInf, {i}, {-1.0}, false);
            art_col[i] = add("y_" + id, 1.0, kInf, {i}, {1.0}, true);
        } else {
            std::fprintf(stderr, "dwsolver-b200: linking row %d is free or ranged, which the reference does not support (src/dw_solver.c:420-425)\n", i + 1);
            return fail(DW_STATUS_ERR_INTERNAL);
        }
    }
    for (int k = 0; k < K; ++k) M.set_rhs(L + k, 1.0);
    // --perturb: every linking right-hand side is moved by 1e-8 * (1-based row index) against degeneracy
    // (src/dw_solver.c:427-431) and moved back after Phase II (:701-707)
    if (opt.perturb) for (int i = 0; i < L; ++i) M.set_rhs(i, M.rhs(i) + 1e-8 * (i + 1));
    const bool need_phase1 = L > 0;

    // ---- per-round plumbing ---------------------------------------------------------------------------
    std::vector<double> r(K, kDwInfinity), pi(std::max(L, 1), 0.0);
    std::vector<Proposal> X;
    std::vector<std::pair<int, double>> remembered;              // Phase I: (column, true cost) (:143-144, 163)
    std::vector<double> xbuf;
    int iter = 0;
    bool warned_nofeas = false;
    // DWSOLVER_MULTI=P (2..8): Phase II rounds price every block with P dual vectors — the master's duals and P-1 points
    // between them and the previous round's duals — and offer all the proposals to the master (the reference
    // prices one vector per round, src/dw_phases.c:358-437).  multi_now says whether the round just run was such a round.
    int multiP = 1;
    if (const char *mp = std::getenv("DWSOLVER_MULTI")) multiP = std::max(1, std::min(8, std::atoi(mp)));
    std::vector<double> pi_prev;
    bool multi_now = false;
    auto consume = [&](bool phase_one, int *n_added) -> int {
        const double t_c0 = now_s();
        int added = 0;
        const size_t first_new = X.size();
        const int passes = multi_now ? multiP : 1;
        for (Shard &sh : shards) sh.accepted.clear();
        for (int pass = 0; pass < passes; ++pass)
        for (Shard &sh : shards) {
            const dwgpu_packed &pk = multi_now ? sh.mpk[(size_t)pass] : sh.pk;
            for (int k = sh.k0; k < sh.k1; ++k) {
                const int kl = k - sh.k0;
                if (pk.status[kl] == DWGPU_UNBND) {
                    std::fprintf(stderr, "dwsolver-b200: subproblem %d is unbounded; unbounded subproblems are not supported (src/dw_phases.c:111-115)\n", k);
                    return DW_STATUS_ERR_UNBOUNDED;
                }
                // The reference never looks at glp_simplex's return code (src/dw_subprob.c:329) because GLPK practically
                // never fails there; this engine's kernels have real failure exits (iteration limit, singular rebuild).
                // A block that did not reach an optimum must not become a lambda column: its x_k need not even be feasible.
                if (pk.status[kl] == DWGPU_UNDEF) {
                    std::fprintf(stderr, "dwsolver-b200: subproblem %d did not reach an optimum (engine status %d)\n", k, pk.status[kl]);
                    return DW_STATUS_ERR_INTERNAL;
                }
                if (pk.status[kl] == DWGPU_NOFEAS) {
                    if (!warned_nofeas && verb >= DW_OUTPUT_QUIET) std::printf("Subproblem %d is infeasible.\n", k);
                    warned_nofeas = true;
                    continue;                                          // no point of the block exists: nothing to propose
                }
                // the acceptance test compares with the reduced cost AT THE MASTER'S DUALS: for the pass priced with them
                // that is the subproblem optimum itself (src/dw_phases.c:108, 358), for the other passes it is recomputed
                double dred = pk.obj[kl];
                if (pass > 0) {
                    dred = pk.cost[kl];
                    for (int t = pk.colptr[kl]; t < pk.colptr[kl + 1]; ++t) dred -= pi[pk.ind[t] - 1] * pk.val[t];
                }
                if (!(r[k] - dred > kTolerance)) continue;
                std::vector<int> idx;
                std::vector<double> val;
                for (int t = pk.colptr[kl]; t < pk.colptr[kl + 1]; ++t) { idx.push_back(pk.ind[t] - 1); val.push_back(pk.val[t]); }
                idx.push_back(L + k); val.push_back(1.0);
                const int tag = iter + 100000 * pass;                   // unique per (round, pass)
                const std::string name = "lambda_" + std::to_string(k) + "_" + std::to_string(tag);
                const int j = add(name, phase_one ? 0.0 : pk.cost[kl], 1.0, idx, val, false);
                if (phase_one) remembered.emplace_back(j, pk.cost[kl]);
                Proposal p; p.k = k; p.iter = tag; p.id = multi_now ? sh.mids[(size_t)pass * (size_t)(sh.k1 - sh.k0) + kl] : -1;
                X.push_back(p);
                if (!multi_now) sh.accepted.push_back(kl);
                ++added;
            }
        }
        // one device-to-device launch per shard keeps the x_k of every accepted block (no per-column host copies);
        // a multi-vector round kept every pass's x_k on the device already
        const double t_h0 = now_s();
        if (!multi_now) {
            if (on_all_shards([&](Shard &sh) {
                    sh.kept_ids.resize(sh.accepted.size());
                    return dwgpu_history_push(sh.eng, (int)sh.accepted.size(), sh.accepted.data(), sh.kept_ids.data());
                }) != 0) return DW_STATUS_ERR_INTERNAL;
            size_t w = first_new;
            for (Shard &sh : shards) for (size_t t = 0; t < sh.accepted.size(); ++t) X[w++].id = sh.kept_ids[t];
        }
        t_history += now_s() - t_h0;
        t_consume += now_s() - t_c0;
        *n_added = added;
        return 0;
    };
    auto solve_master = [&]() {
        const double t0 = now_s();
        const int rc = M.solve();
        g_stats.seconds_master += now_s() - t0;
        for (int i = 0; i < L; ++i) pi[i] = M.row_dual(i);
        for (int k = 0; k < K; ++k) r[k] = M.row_dual(L + k);
        return rc;
    };
    auto banner = [&]() {
        if (verb >= DW_OUTPUT_NORMAL) {
            std::printf("################################################\n");
            std::printf("####  Master objective value = %e \n", M.obj_val());
            std::printf("################################################\n");
        }
    };
    auto round = [&](bool phase_one) -> int {
        const double t0 = now_s();
        int rc;
        multi_now = multiP > 1 && !phase_one && L > 0 && (int)pi_prev.size() == L;
        if (multi_now) {
            // pass 0: the master's duals; pass p: a point p / P of the way back to the previous round's duals
            std::vector<double> Pi((size_t)multiP * L);
            for (int p = 0; p < multiP; ++p) {
                const double a = (double)p / multiP;
                for (int i = 0; i < L; ++i) Pi[(size_t)p * L + i] = (1.0 - a) * pi[i] + a * pi_prev[i];
            }
            rc = on_all_shards([&](Shard &sh) {
                sh.mpk.resize((size_t)multiP);
                sh.mids.resize((size_t)multiP * (size_t)(sh.k1 - sh.k0));
                return dwgpu_iterate_multi(sh.eng, Pi.data(), multiP, 0, sh.mpk.data(), sh.mids.data());
            });
        } else {
            // with r given the engine ships only the columns the acceptance test above will take
            rc = on_all_shards([&](Shard &sh) {
                return dwgpu_iterate_packed(sh.eng, pi.data(), phase_one ? 1 : 0, r.data() + sh.k0, &sh.pk);
            });
        }
        if (!phase_one && L > 0) pi_prev.assign(pi.begin(), pi.begin() + L);
        g_stats.seconds_subproblems += now_s() - t0;
        return rc;
    };

    // first proposals: cold solve with the file objective, always accepted (r = 1e6) (src/dw_subprob.c:121-122, 184)
    {
        const double t0 = now_s();
        if (on_all_shards([&](Shard &sh) { return dwgpu_initial_solve_packed(sh.eng, &sh.pk); }) != 0)
            return fail(DW_STATUS_ERR_INTERNAL);
        g_stats.seconds_subproblems += now_s() - t0;
    }
    for (Shard &sh : shards)
        for (int k = sh.k0; k < sh.k1; ++k)
            if (sh.pk.status[k - sh.k0] == DWGPU_NOFEAS) {
                if (verb >= DW_OUTPUT_QUIET) std::printf("Subproblem %d is infeasible.\n", k);
                warned_nofeas = true;
            }

    int status = DW_STATUS_OK;
    int added = 0;
    // ---- Phase I (src/dw_solver.c:461-628) ---------------------------------------------------------------
    if (need_phase1) {
        M.set_obj_const(0.0);                                    // shift zeroed during Phase I (:463)
        bool first = true, feasible = false;
        int p1 = 0;
        for (; p1 < opt.max_phase1_iterations; ++p1) {
            const int n_before = M.cols();
            int rc = consume(true, &added);
            if (rc) return fail(rc);
            if (first) {                                         // src/dw_phases.c:165-171, 201-217
                std::vector<double> acc(L, 0.0);
                for (int j = n_before; j < M.cols(); ++j) {
                    const auto &ix = M.col_idx(j);
                    const auto &vv = M.col_val(j);
                    for (size_t t = 0; t + 1 < ix.size(); ++t) acc[ix[t]] += vv[t];
                }
                for (int i = 0; i < L; ++i)
                    if (M.rhs(i) - acc[i] < 0.0 && art_col[i] >= 0) M.set_coef(art_col[i], 0, -1.0);
                first = false;
            }
            const int mrc = solve_master();
            ++iter;
            ++g_stats.phase1_iterations;
            if (verb >= DW_OUTPUT_NORMAL) {
                std::printf("################################################\n");
                std::printf("####  Master objective value = %.8e \n", M.obj_val());      // src/dw_phases.c:250
                std::printf("################################################\n");
            }
            if (mrc != 0) { std::fprintf(stderr, "dwsolver-b200: Phase I master LP failed (code %d)\n", mrc); return fail(DW_STATUS_ERR_INTERNAL); }
            if (std::fabs(M.obj_val()) < kTolerance) { feasible = true; break; }     // :509-514
            if (added == 0) break;                                                   // :516-521
            if (round(true) != 0) return fail(DW_STATUS_ERR_INTERNAL);
        }
        if (!feasible) {
            if (p1 >= opt.max_phase1_iterations) {
                if (verb >= DW_OUTPUT_QUIET) std::printf("Phase I iteration limit reached.\n");
                status = DW_STATUS_ERR_PHASE1_LIMIT;
            } else {
                if (verb >= DW_OUTPUT_QUIET) std::printf("Problem is infeasible.\n");
                status = DW_STATUS_ERR_INFEASIBLE;
            }
        }
        // leave Phase I: drop the artificials, install the remembered costs, re-solve (:535-562)
        {
            std::vector<char> flag(M.cols(), 0);
            std::vector<int> remap(M.cols(), -1);
            int w = 0;
            for (int j = 0; j < M.cols(); ++j) { flag[j] = is_art[j]; if (!flag[j]) remap[j] = w++; }
            M.del_cols(flag);
            std::vector<char> na(w, 0);
            is_art.swap(na);
            for (auto &e : remembered) if (remap[e.first] >= 0) M.set_cost(remap[e.first], e.second);
            remembered.clear();
        }
        M.set_obj_const(opt.shift);                              // :568
        const int mrc = solve_master();
        ++iter;                                                  // :618 (the drained round is simply not launched)
        if (mrc == 1 && status == DW_STATUS_OK) status = DW_STATUS_ERR_INFEASIBLE;
        if (mrc == 2 && status == DW_STATUS_OK) status = DW_STATUS_ERR_UNBOUNDED;
        if (mrc == 3 && status == DW_STATUS_OK) { std::fprintf(stderr, "dwsolver-b200: master LP failed after Phase I (code %d)\n", mrc); return fail(DW_STATUS_ERR_INTERNAL); }
    } else {
        M.set_obj_const(opt.shift);
        int rc = consume(false, &added);
        if (rc) return fail(rc);
        const int mrc0 = solve_master();
        ++iter;
        if (mrc0 == 1) status = DW_STATUS_ERR_INFEASIBLE;
        else if (mrc0 == 2) status = DW_STATUS_ERR_UNBOUNDED;
        else if (mrc0 != 0) { std::fprintf(stderr, "dwsolver-b200: master LP failed (code %d)\n", mrc0); return fail(DW_STATUS_ERR_INTERNAL); }
    }

    // ---- Phase II (src/dw_solver.c:649-694, stop rule src/dw_phases.c:462-484) --------------------------------
    if (status == DW_STATUS_OK) {
        long long simplex_iterations = 0;                        // per solve (the reference's is function-static)
        int p2 = 0;
        bool done = false;
        for (; p2 < opt.max_phase2_iterations && !done; ++p2) {
            if (round(false) != 0) return fail(DW_STATUS_ERR_INTERNAL);
            int rc = consume(false, &added);
            if (rc) return fail(rc);
            const int mrc = solve_master();
            ++iter;
            if (mrc == 2) { status = DW_STATUS_ERR_UNBOUNDED; break; }
            if (mrc != 0) { std::fprintf(stderr, "dwsolver-b200: master LP failed (code %d)\n", mrc); return fail(DW_STATUS_ERR_INTERNAL); }
            const bool prog = simplex_iterations < M.it_cnt();
            if (prog) {
                simplex_iterations = M.it_cnt();
                if (added == 0 && verb >= DW_OUTPUT_NORMAL) std::printf("No columns added, but simplex made progress.\n");
            } else if (added > 0) {
                if (p2 != 0) {
                    if (verb >= DW_OUTPUT_NORMAL)
                        std::printf("There was no simplex progress made despite columns being added. Finishing now.\n");
                    done = true;
                }
            } else {
                if (verb >= DW_OUTPUT_NORMAL) std::printf("I think we've converged on an optimum.\n");
                done = true;
            }
            banner();
        }
        if (!done && status == DW_STATUS_OK && p2 >= opt.max_phase2_iterations) status = DW_STATUS_ERR_PHASE2_LIMIT;
    }
    if (opt.perturb) {                                           // src/dw_solver.c:701-724
        for (int i = 0; i < L; ++i) M.set_rhs(i, M.rhs(i) - 1e-8 * (i + 1));
        const double t0 = now_s();
        const int mrc = M.solve();
        g_stats.seconds_master += now_s() - t0;
        if (mrc != 0 && status == DW_STATUS_OK) {
            std::fprintf(stderr, "dwsolver-b200: master LP failed after the perturbation was removed (code %d)\n", mrc);
            return fail(DW_STATUS_ERR_INTERNAL);
        }
    }
    if (verb >= DW_OUTPUT_NORMAL) {                              // src/dw_solver.c:721
        std::printf("################################################\n");
        std::printf("####  Master objective value = %e \n", M.obj_val());
        std::printf("################################################\n");
    }

    if (opt.print_final_master) {                            // src/dw_solver.c:749-753
        if (verb >= DW_OUTPUT_NORMAL) std::printf("%-40s ", "Printing final master to done.cpxlp...");
        const bool ok = write_final_master("done.cpxlp", M, L, K, opt.shift, link_names);
        if (verb >= DW_OUTPUT_NORMAL) std::printf("%s\n", ok ? "DONE!" : "FAILED");
        if (!ok && verb >= DW_OUTPUT_QUIET) std::fprintf(stderr, "dwsolver-b200: cannot write done.cpxlp\n");
    }

    // ---- results (src/dw_solver.c:877-889) ---------------------------------------------------------------------
    result->objective_value = M.obj_val();
    result->num_vars = M.cols();
    result->x = (double *)std::malloc(sizeof(double) * (size_t)std::max(M.cols(), 1));
    if (!result->x) return fail(DW_STATUS_ERR_INTERNAL);
    for (int j = 0; j < M.cols(); ++j) result->x[j] = M.col_prim(j);

    // ---- relaxed_solution: x_k = sum_t lambda_{k,t} x_k^t (src/dw_rounding.c:115-168, 561-597) -------------------
    if (opt.print_relaxed_sol) {
        std::map<std::pair<int, int>, const Proposal *> by_id;
        for (auto &p : X) by_id[{p.k, p.iter}] = &p;
        struct Terms { std::vector<long long> id; std::vector<double> w; };
        std::vector<Terms> terms((size_t)G);                     // per shard, in master-column order
        for (int j = 0; j < M.cols(); ++j) {                     // every column is scanned (the reference skips the first L)
            const double v = M.col_prim(j);
            if (v == 0.0) continue;
            int k = -1, t = -1;
            if (std::sscanf(M.name(j).c_str(), "lambda_%d_%d", &k, &t) != 2) continue;
            auto it = by_id.find({k, t});
            if (it == by_id.end()) continue;
            Terms &tm = terms[(size_t)(&shard_of(k) - shards.data())];
            tm.id.push_back(it->second->id);
            tm.w.push_back(v);
        }
        long long ntot = 0;
        for (int k = 0; k < K; ++k) ntot += desc[k].n;
        std::vector<double> xall((size_t)std::max<long long>(ntot, 1), 0.0);
        // contiguous shards: every engine's block-major x lands in its own slice of the global one
        if (on_all_shards([&](Shard &sh) {
                Terms &tm = terms[(size_t)(&sh - shards.data())];
                return dwgpu_history_combine(sh.eng, (int)tm.id.size(), tm.id.data(), tm.w.data(), xall.data() + sh.x0);
            }) != 0)
            return fail(DW_STATUS_ERR_INTERNAL);
        std::vector<std::vector<double>> xs(K);
        {
            long long off = 0;
            for (int k = 0; k < K; ++k) { xs[k].assign(xall.begin() + off, xall.begin() + off + desc[k].n); off += desc[k].n; }
        }
        FILE *f = std::fopen("relaxed_solution", "w");
        if (f) {
            for (int k = 0; k < K; ++k)
                for (int c = 0; c < desc[k].n; ++c) {
                    if (names) std::fprintf(f, "%s\t%f\n", (*names)[k][c].c_str(), xs[k][c]);
                    else std::fprintf(f, "x_%d_%d\t%f\n", k, c, xs[k][c]);
                }
            std::fclose(f);
        } else if (verb >= DW_OUTPUT_QUIET) {
            std::fprintf(stderr, "dwsolver-b200: cannot write relaxed_solution\n");
        }
    }

    g_stats.sub_pivots = 0;
    for (Shard &sh : shards) {
        dwgpu_counters cnt;
        if (dwgpu_get_counters(sh.eng, &cnt) == 0) g_stats.sub_pivots += cnt.pivots + cnt.flips;
    }
    g_stats.dw_iterations = iter;
    g_stats.master_columns = M.cols();
    g_stats.seconds_total = now_s() - t_begin;
    if (std::getenv("DWSOLVER_DEBUG_TIMING"))
        std::fprintf(stderr, "dwsolver-b200: %d engine(s) created in %.4f s; columns consumed in %.4f s (of which history push %.4f s); "
                             "subproblem calls %.4f s, master %.4f s, total %.4f s\n",
                     G, t_create, t_consume, t_history, g_stats.seconds_subproblems, g_stats.seconds_master, g_stats.seconds_total);
    if (opt.print_timing_data && verb >= DW_OUTPUT_SILENT) {
        std::printf("Total wall time: %.6f s (read %.6f, subproblems on the GPU %.6f, master %.6f); %d DW iterations, %lld subproblem pivots\n",
                    g_stats.seconds_total, g_stats.seconds_read, g_stats.seconds_subproblems, g_stats.seconds_master,
                    g_stats.dw_iterations, g_stats.sub_pivots);
    }
    return fail(status);
}

// dw_core behind the C boundary: nothing in this layer may exit or throw into the caller (a thread that cannot be
// created, an allocation that fails) — it comes back as DW_STATUS_ERR_INTERNAL like any engine failure
template <typename... A>
static int dw_core_guarded(dw_result_t *result, A &&...args)
{
    try {
        return dw_core(std::forward<A>(args)...);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "dwsolver-b200: %s\n", e.what());
    } catch (...) {
        std::fprintf(stderr, "dwsolver-b200: unexpected failure in the solve loop\n");
    }
    if (result) { std::free(result->x); result->x = nullptr; result->num_vars = 0; result->status = DW_STATUS_ERR_INTERNAL; }
    return DW_STATUS_ERR_INTERNAL;
}

extern "C" {

void dw_options_init(dw_options_t *o)
{
    if (!o) return;
    o->verbosity = DW_OUTPUT_NORMAL;       // src/dw_solver.c:60-75
    o->mip_gap = 0.01;
    o->max_phase1_iterations = 100;
    o->max_phase2_iterations = 3000;
    o->rounding_flag = 0;
    o->integerize_flag = 0;
    o->enforce_sub_integrality = 0;
    o->print_timing_data = 0;
    o->print_final_master = 0;
    o->print_relaxed_sol = 1;
    o->perturb = 0;
    o->shift = 0.0;
}

const char *dw_version(void) { return "dwsolver-b200 0.1 (API of dwsolver 1.2.1; subproblems on sm_100a)"; }

void dw_result_free(dw_result_t *r)
{
    if (!r) return;
    std::free(r->x);
    std::memset(r, 0, sizeof(*r));
}

void dw_last_stats(dw_stats_t *out) { if (out) *out = g_stats; }

int dw_solve(const char *master_file, const char *const *subproblem_files, int num_subproblems, const dw_options_t *opts_in,
             dw_result_t *result)
{
    if (!result) return DW_STATUS_ERR_BAD_ARGS;
    std::memset(result, 0, sizeof(*result));
    auto fail = [&](int code) { result->status = code; return code; };
    if (!master_file || !subproblem_files || num_subproblems < 1) return fail(DW_STATUS_ERR_BAD_ARGS);
    for (int k = 0; k < num_subproblems; ++k) if (!subproblem_files[k]) return fail(DW_STATUS_ERR_BAD_ARGS);
    dw_options_t opt;
    if (opts_in) opt = *opts_in; else dw_options_init(&opt);
    const int verb = opt.verbosity;
    std::memset(&g_stats, 0, sizeof(g_stats));
    const double t_begin = now_s();
    if (opt.enforce_sub_integrality || opt.integerize_flag || opt.rounding_flag) {
        if (verb >= DW_OUTPUT_QUIET)
            std::fprintf(stderr, "dwsolver-b200: integer options (-e, -i, -r) are not supported by the GPU path; solving the LP relaxation.\n");
    }

    // ---- read the problem -------------------------------------------------------------------------
    const int K = num_subproblems;
    LpModel master;
    std::vector<BlockHost> blocks;
    try {
        read_problem(master_file, subproblem_files, K, 0, master, blocks);
    } catch (const std::exception &e) {
        if (verb >= DW_OUTPUT_QUIET) std::fprintf(stderr, "dwsolver-b200: %s\n", e.what());
        return fail(DW_STATUS_ERR_FILE);
    }
    const int L = master.m();
    g_stats.seconds_read = now_s() - t_begin;
    std::vector<dwgpu_block_desc> desc;
    describe(blocks, desc);
    std::vector<std::vector<std::string>> names(K);
    for (int k = 0; k < K; ++k) names[k] = blocks[k].col_names;
    std::vector<double> llb(L), lub(L);
    for (int i = 0; i < L; ++i) { llb[i] = to_engine_bound(master.row_lb[i]); lub[i] = to_engine_bound(master.row_ub[i]); }
    return dw_core_guarded(result, K, L, desc.data(), llb.data(), lub.data(), &names, opt, t_begin, result, &master.row_names);
}

int dw_solve_blocks(int K, int L, const dwgpu_block_desc *blocks, const double *link_lb, const double *link_ub,
                    const dw_options_t *opts_in, dw_result_t *result)
{
    if (!result) return DW_STATUS_ERR_BAD_ARGS;
    std::memset(result, 0, sizeof(*result));
    if (K < 1 || L < 0 || !blocks || (L > 0 && (!link_lb || !link_ub))) { result->status = DW_STATUS_ERR_BAD_ARGS; return DW_STATUS_ERR_BAD_ARGS; }
    dw_options_t opt;
    if (opts_in) opt = *opts_in; else dw_options_init(&opt);
    std::memset(&g_stats, 0, sizeof(g_stats));
    return dw_core_guarded(result, K, L, blocks, link_lb, link_ub, nullptr, opt, now_s(), result);
}

int dw_write_container(const char *master_file, const char *const *subproblem_files, int num_subproblems, double shift,
                       const char *container_path)
{
    if (!master_file || !subproblem_files || num_subproblems < 1 || !container_path) return DW_STATUS_ERR_BAD_ARGS;
    for (int k = 0; k < num_subproblems; ++k) if (!subproblem_files[k]) return DW_STATUS_ERR_BAD_ARGS;
    const int K = num_subproblems;
    LpModel master;
    std::vector<BlockHost> blocks;
    try {
        read_problem(master_file, subproblem_files, K, 0, master, blocks);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "dwsolver-b200: %s\n", e.what());
        return DW_STATUS_ERR_FILE;
    }
    if (!write_container(container_path, master, blocks, shift)) {
        std::fprintf(stderr, "dwsolver-b200: cannot write %s\n", container_path);
        return DW_STATUS_ERR_FILE;
    }
    return DW_STATUS_OK;
}

int dw_solve_container(const char *container_path, const dw_options_t *opts_in, dw_result_t *result)
{
    if (!result) return DW_STATUS_ERR_BAD_ARGS;
    std::memset(result, 0, sizeof(*result));
    if (!container_path) { result->status = DW_STATUS_ERR_BAD_ARGS; return DW_STATUS_ERR_BAD_ARGS; }
    dw_options_t opt;
    if (opts_in) opt = *opts_in; else dw_options_init(&opt);
    std::memset(&g_stats, 0, sizeof(g_stats));
    const double t_begin = now_s();
    ContainerView v;
    try {
        load_container(container_path, v);
    } catch (const std::exception &e) {
        if (opt.verbosity >= DW_OUTPUT_QUIET) std::fprintf(stderr, "dwsolver-b200: %s\n", e.what());
        result->status = DW_STATUS_ERR_FILE;
        return DW_STATUS_ERR_FILE;
    }
    g_stats.seconds_read = now_s() - t_begin;
    if (opt.shift == 0.0) opt.shift = v.shift;          // an explicit option wins over the stored constant
    return dw_core_guarded(result, v.K, v.L, v.desc.data(), v.link_lb, v.link_ub, v.has_names ? &v.names : nullptr, opt, t_begin, result);
}

}  // extern "C"

