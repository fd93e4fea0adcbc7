// gub_master.cc — GUB (key-column) revised primal simplex for the restricted master; see gub_master.h.
#include "gub_master.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <limits>
#include <thread>

namespace dwh {
namespace {
// the two loops the master spends its time in; cloned for wider vector units and picked at load time
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define DWH_CLONES __attribute__((target_clones("avx512f", "avx2", "default")))
#else
#define DWH_CLONES
#endif
DWH_CLONES void axpy(int n, double a, const double *x, double *y)
{
    for (int i = 0; i < n; ++i) y[i] += a * x[i];
}
DWH_CLONES void scale(int n, double a, double *x)
{
    for (int i = 0; i < n; ++i) x[i] *= a;
}
// row -= sum_j coef[j] * V[j][:], j = 0..k-1 in this order for every element (what k successive axpy calls compute), in
// one pass over the row: a chunk of the row stays in registers while the k pending vectors stream through it
DWH_CLONES void apply_pending(int n, int k, const double *coef, const double *V, int ldv, double *row)
{
    constexpr int kChunk = 32;
    int i = 0;
    for (; i + kChunk <= n; i += kChunk) {
        double acc[kChunk];
        for (int t = 0; t < kChunk; ++t) acc[t] = row[i + t];
        for (int j = 0; j < k; ++j) {
            const double a = coef[j];
            const double *v = V + (size_t)j * ldv + i;
            for (int t = 0; t < kChunk; ++t) acc[t] -= a * v[t];
        }
        for (int t = 0; t < kChunk; ++t) row[i + t] = acc[t];
    }
    for (; i < n; ++i) {
        double s = row[i];
        for (int j = 0; j < k; ++j) s -= coef[j] * V[(size_t)j * ldv + i];
        row[i] = s;
    }
}
constexpr double kTolDj = 1e-9, kTolBnd = 1e-9, kTolPiv = 1e-9, kTolTie = 1e-12;
constexpr double kInfD = std::numeric_limits<double>::infinity();
constexpr int kRefactorEvery = 2000;     // counted in iterations that changed the working inverse (a refactorisation
                                          // costs ~2 L^3 flops: ~150 ordinary updates at L = 256)
constexpr size_t kMaxCandidates = 2048, kCandidateChunk = 128, kEnoughCandidates = 192;
}  // namespace

GubMaster::GubMaster(int linking_rows, int blocks)
    : L_(linking_rows), K_(blocks), b_(linking_rows + blocks, 0.0), key_(blocks, -1), pos_var_(linking_rows),
      y_(linking_rows + blocks, 0.0)
{
    for (int r = 0; r < L_; ++r) pos_var_[r] = -(r + 1);
}

int GubMaster::add_col(const std::string &name, double cost, double /*ub*/, const std::vector<int> &idx,
                       const std::vector<double> &val)
{
    name_.push_back(name);
    cost_.push_back(cost);
    idx_.push_back(idx);
    val_.push_back(val);
    int k = -1;
    for (size_t t = 0; t < idx.size(); ++t) if (idx[t] >= L_) k = idx[t] - L_;     // convexity entry (coefficient 1)
    set_.push_back(k);
    stat_.push_back(NONBASIC);
    x_.push_back(0.0);
    if (flat_valid_) {                       // keep the flat mirror in step (columns only ever arrive at the end)
        for (size_t t = 0; t < idx.size(); ++t) if (idx[t] < L_) { lrow_.push_back(idx[t]); lval_.push_back(val[t]); }
        lptr_.push_back((int)lrow_.size());
    }
    return cols() - 1;
}

void GubMaster::rebuild_flat()
{
    const int n = cols();
    lptr_.assign(1, 0); lrow_.clear(); lval_.clear();
    lptr_.reserve((size_t)n + 1);
    for (int j = 0; j < n; ++j) {
        for (size_t t = 0; t < idx_[j].size(); ++t) if (idx_[j][t] < L_) { lrow_.push_back(idx_[j][t]); lval_.push_back(val_[j][t]); }
        lptr_.push_back((int)lrow_.size());
    }
    flat_valid_ = true;
}

void GubMaster::del_cols(const std::vector<char> &flag)
{
    const int n = cols();
    std::vector<int> remap(n, -1);
    int w = 0;
    for (int j = 0; j < n; ++j) {
        if (flag[j]) continue;
        remap[j] = w;
        if (w != j) {
            name_[w] = std::move(name_[j]); cost_[w] = cost_[j]; idx_[w] = std::move(idx_[j]); val_[w] = std::move(val_[j]);
            set_[w] = set_[j]; stat_[w] = stat_[j]; x_[w] = x_[j];
        }
        ++w;
    }
    name_.resize(w); cost_.resize(w); idx_.resize(w); val_.resize(w); set_.resize(w); stat_.resize(w); x_.resize(w);
    for (int k = 0; k < K_; ++k) key_[k] = key_[k] >= 0 ? remap[key_[k]] : -1;
    factor_valid_ = false;       // build_basis() re-keys blocks whose key went away and fills holes with logicals
    vals_valid_ = false;
    flat_valid_ = false;
}

void GubMaster::add_link(int j, double scale, std::vector<double> &w) const
{
    if (flat_valid_) {
        for (int t = lptr_[j]; t < lptr_[j + 1]; ++t) w[lrow_[t]] += scale * lval_[t];
        return;
    }
    const std::vector<int> &ix = idx_[j];
    const std::vector<double> &vv = val_[j];
    for (size_t t = 0; t < ix.size(); ++t) if (ix[t] < L_) w[ix[t]] += scale * vv[t];
}

void GubMaster::link_minus_key(int j, std::vector<double> &w) const
{
    std::fill(w.begin(), w.end(), 0.0);
    add_link(j, 1.0, w);
    if (set_[j] >= 0 && key_[set_[j]] >= 0 && key_[set_[j]] != j) add_link(key_[set_[j]], -1.0, w);
}

void GubMaster::refactor()
{
    const int L = L_, n = cols();
    pend_n_ = 0;                      // a fresh inverse of the current basis supersedes every pending update
    winv_.assign((size_t)L * L, 0.0);
    for (int r = 0; r < L; ++r) { winv_[(size_t)r * L + r] = 1.0; pos_var_[r] = -(r + 1); }
    std::vector<double> w(L), alpha(L);
    std::vector<int> nz;
    for (int j = 0; j < n; ++j) {
        if (stat_[j] != NONKEY) continue;
        link_minus_key(j, w);
        nz.clear();
        for (int i = 0; i < L; ++i) if (w[i] != 0.0) nz.push_back(i);
        for (int r = 0; r < L; ++r) {
            const double *row = &winv_[(size_t)r * L];
            double s = 0.0;
            for (size_t t = 0; t < nz.size(); ++t) s += row[nz[t]] * w[nz[t]];
            alpha[r] = s;
        }
        int p = -1;
        double best = 1e-9;
        for (int r = 0; r < L; ++r)
            if (pos_var_[r] < 0 && std::fabs(alpha[r]) > best) { best = std::fabs(alpha[r]); p = r; }
        if (p < 0) { stat_[j] = NONBASIC; continue; }            // dependent on the columns already placed
        double *rp = &winv_[(size_t)p * L];
        scale(L, 1.0 / alpha[p], rp);
        for (int r = 0; r < L; ++r) {
            if (r == p || alpha[r] == 0.0) continue;
            axpy(L, -alpha[r], rp, &winv_[(size_t)r * L]);
        }
        pos_var_[p] = j;
    }
    factor_valid_ = true;
    vals_valid_ = false;
}

bool GubMaster::build_basis()
{
    const int n = cols();
    // every block needs a key column
    std::vector<int> first(K_, -1);
    for (int j = 0; j < n; ++j) if (set_[j] >= 0 && first[set_[j]] < 0) first[set_[j]] = j;
    for (int k = 0; k < K_; ++k) {
        if (key_[k] >= 0 && stat_[key_[k]] == KEY) continue;
        int cand = -1;
        for (int j = 0; j < n && cand < 0; ++j) if (set_[j] == k && stat_[j] == NONKEY) cand = j;     // keep a basic one basic
        if (cand < 0) cand = first[k];
        if (cand < 0) return false;
        key_[k] = cand;
        stat_[cand] = KEY;
    }
    if (!crashed_) {
        // first basis: for every linking row a singleton column (slack / surplus / artificial) whose sign matches
        // the residual left by the key columns — primal feasible without an LP phase 1 whenever the master was
        // built the way the reference builds it (src/dw_solver.c:321-419, artificial sign fix src/dw_phases.c:201-217)
        std::vector<double> res(b_.begin(), b_.begin() + L_);
        for (int k = 0; k < K_; ++k) add_link(key_[k], -b_[L_ + k], res);
        std::vector<int> pick(L_, -1);
        for (int j = 0; j < n; ++j) {
            if (set_[j] >= 0 || stat_[j] != NONBASIC) continue;
            int row = -1, cnt = 0;
            double coef = 0.0;
            for (size_t t = 0; t < idx_[j].size(); ++t) if (val_[j][t] != 0.0) { row = idx_[j][t]; coef = val_[j][t]; ++cnt; }
            if (cnt != 1 || row < 0 || row >= L_) continue;
            const bool sign_ok = res[row] == 0.0 || (res[row] > 0.0) == (coef > 0.0);
            if (!sign_ok) continue;
            if (pick[row] < 0 || cost_[j] < cost_[pick[row]]) pick[row] = j;       // prefer the free slack over the artificial
        }
        for (int i = 0; i < L_; ++i) if (pick[i] >= 0) stat_[pick[i]] = NONKEY;
        crashed_ = true;
    }
    refactor();
    return true;
}

void GubMaster::compute_values()
{
    const int L = L_;
    std::vector<double> rhs(b_.begin(), b_.begin() + L);
    for (int k = 0; k < K_; ++k) add_link(key_[k], -b_[L + k], rhs);
    xn_.assign(L, 0.0);
    for (int r = 0; r < L; ++r) {
        const double *row = &winv_[(size_t)r * L];
        double s = 0.0;
        for (int i = 0; i < L; ++i) s += row[i] * rhs[i];
        xn_[r] = s;
    }
    xkey_.assign(K_, 0.0);
    for (int k = 0; k < K_; ++k) xkey_[k] = b_[L + k];
    for (int r = 0; r < L; ++r) {
        const int v = pos_var_[r];
        if (v >= 0 && set_[v] >= 0) xkey_[set_[v]] -= xn_[r];
    }
    vals_valid_ = true;
}

void GubMaster::publish(bool with_duals)
{
    const int n = cols(), L = L_;
    std::fill(x_.begin(), x_.end(), 0.0);
    for (int r = 0; r < L; ++r) if (pos_var_[r] >= 0) x_[pos_var_[r]] = xn_[r];
    for (int k = 0; k < K_; ++k) if (key_[k] >= 0) x_[key_[k]] = xkey_[k];
    obj_ = c0_;
    for (int j = 0; j < n; ++j) obj_ += cost_[j] * x_[j];
    if (!with_duals) return;
    // pi = W^-T ctilde, mu_k = c_key - pi . a_key   (true costs)
    std::vector<double> ct(L, 0.0);
    for (int r = 0; r < L; ++r) {
        const int v = pos_var_[r];
        if (v < 0) continue;
        ct[r] = cost_[v] - (set_[v] >= 0 ? cost_[key_[set_[v]]] : 0.0);
    }
    std::fill(y_.begin(), y_.end(), 0.0);
    for (int r = 0; r < L; ++r) {
        if (ct[r] == 0.0) continue;
        const double *row = &winv_[(size_t)r * L];
        for (int i = 0; i < L; ++i) y_[i] += ct[r] * row[i];
    }
    for (int k = 0; k < K_; ++k) {
        const int j = key_[k];
        double s = cost_[j];
        for (size_t t = 0; t < idx_[j].size(); ++t) if (idx_[j][t] < L) s -= y_[idx_[j][t]] * val_[j][t];
        y_[L + k] = s;
    }
}

int GubMaster::solve()
{
    const int L = L_, K = K_;
    if (!flat_valid_) rebuild_flat();
    if (!factor_valid_) { if (!build_basis()) return 1; }
    if (!vals_valid_) compute_values();
    // the dense L x L work is split by rows over a persistent team (thread_team.h); small masters stay serial
    if (!team_) {
        // measured on the 16-core B200 host at L = 256: 4 members take the master of config B from 8.3 s to 6.3 s,
        // 8 are no better (the serial ratio test and candidate pricing remain); on small or shared hosts the hand-offs
        // cost more than the slices save
        int want = 1;
        if (L >= 128 && std::thread::hardware_concurrency() >= 12) want = 4;
        if (const char *e = std::getenv("DWSOLVER_MASTER_THREADS")) want = std::max(1, std::min(64, std::atoi(e)));
        team_.reset(new ThreadTeam(std::min(want, std::max(1, L / 16))));
    }
    ThreadTeam &team = *team_;
    const int T = team.size();
    // Blocked updates of the working inverse (k = 8 once the L x L array has outgrown the L1 cache, L >= 128;
    // DWSOLVER_MASTER_BLOCK = k, 2..64, overrides and 0 switches them off): a column
    // replacement W^-1 -= u v' is kept as a pending pair instead of being swept over the L x L array (that sweep is
    // bound by L2 bandwidth: 1 MB per iteration at L = 256), FTRAN, the duals and the pivot row fold the pending pairs
    // in at O(kL), and every k-th replacement applies them all in ONE pass over the array.  Same mathematics, other
    // rounding: the pivot sequence differs from the unblocked solver's (config B: 249 088 iterations against 246 554).
    int block_k = L >= 128 ? 8 : 0;      // measured at L = 256: k = 8 beats 4, 6, 12 and 16 (profiles/README.md)
    if (const char *e = std::getenv("DWSOLVER_MASTER_BLOCK")) block_k = std::max(0, std::min(64, std::atoi(e)));
    if (block_k == 1) block_k = 0;
    if (block_k > 0 && (int)pend_u_.size() < block_k * L) { pend_u_.assign((size_t)block_k * L, 0.0); pend_v_.assign((size_t)block_k * L, 0.0); }
    auto flush = [&]() {
        if (pend_n_ == 0) return;
        team.run([&](int member, int members) {
            const int r0 = L * member / members, r1 = L * (member + 1) / members;
            double coef[64];
            for (int r = r0; r < r1; ++r) {
                bool any = false;
                for (int j = 0; j < pend_n_; ++j) { coef[j] = pend_u_[(size_t)j * L + r]; any |= coef[j] != 0.0; }
                if (any) apply_pending(L, pend_n_, coef, pend_v_.data(), L, &winv_[(size_t)r * L]);
            }
        });
        pend_n_ = 0;
    };
    std::vector<double> pi(L), ct(L), y(L), wq(L), wn(L), u(L), rpc(L);
    std::vector<double> wk(K, 0.0), rate(K, 0.0);
    std::vector<long long> rate_stamp(K, -1), best_stamp(K, -1);
    // mu_k memo, one per team member (member 0 = this thread), so that a parallel pricing pass shares nothing mutable
    std::vector<std::vector<double>> mu_m(T, std::vector<double>(K, 0.0));
    std::vector<std::vector<long long>> mu_stamp_m(T, std::vector<long long>(K, -1));
    std::vector<double> dbuf;
    std::vector<int> best_slot(K, 0);
    std::vector<int> touched, cand;
    long long stamp = 0;
    int degen_run = 0, since_refactor = 0, verify_rounds = 0;
    long long n_degen = 0, n_trivial = 0, n_full = 0, n_keyswap = 0, it_begin = iters_;
    const bool prof = std::getenv("DWSOLVER_MASTER_DEBUG") != nullptr;
    double t_dual = 0, t_cand = 0, t_pass = 0, t_ftran = 0, t_ratio = 0, t_update = 0;
    auto tick = [&]() { return prof ? std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count() : 0.0; };
    bool bland = false, pi_valid = false, last_phase1 = false, scan_keys = true;
    int n_key_infeasible = 0;
    std::vector<int> wq_nz;
    const long long it_limit = iters_ + 200LL * (L + K + cols()) + 20000LL;
    int status = 3;

    auto is_logical = [&](int r) { return pos_var_[r] < 0; };
    for (;;) {
        if (iters_ > it_limit) { status = 3; break; }
        if (since_refactor >= kRefactorEvery) { refactor(); compute_values(); since_refactor = 0; pi_valid = false; scan_keys = true; }
        const int n = cols();
        ++stamp;
        // ---- phase: infeasibility weights (all variables >= 0; logicals are fixed at 0)
        bool phase1 = false;
        for (int r = 0; r < L; ++r) {
            double w = 0.0;
            if (xn_[r] < -kTolBnd) w = -1.0;
            else if (is_logical(r) && xn_[r] > kTolBnd) w = 1.0;
            wn[r] = w;
            phase1 |= w != 0.0;
        }
        // key variables: a full scan only after the values were recomputed; otherwise only the blocks the last
        // step touched can have changed sign (K = 8192 at config B: this loop would cost more than the pivot)
        if (scan_keys) {
            n_key_infeasible = 0;
            for (int k = 0; k < K; ++k) { wk[k] = xkey_[k] < -kTolBnd ? -1.0 : 0.0; n_key_infeasible += wk[k] != 0.0; }
            scan_keys = false;
        }
        phase1 |= n_key_infeasible > 0;
        auto cost_of_key = [&](int k) { return phase1 ? wk[k] : cost_[key_[k]]; };
        double t0 = tick();
        // ---- duals of the working basis.  pi only depends on the working basis and on the costs of the blocks that
        // have non-key basic columns; the most common iteration of a DW master — a block's only basic column (its key)
        // is replaced by a better column of the same block — changes neither, so pi is kept across it.
        if (phase1 || last_phase1 || !pi_valid) {
            for (int r = 0; r < L; ++r) {
                const int v = pos_var_[r];
                double c = phase1 ? wn[r] : (v >= 0 ? cost_[v] : 0.0);
                if (v >= 0 && set_[v] >= 0) c -= cost_of_key(set_[v]);
                ct[r] = c;
            }
            // pi = W^-T ct.  Every member accumulates its own slice of COLUMNS over all rows in row order — the same
            // additions in the same order as a serial loop, so the result does not depend on the team size
            std::fill(pi.begin(), pi.end(), 0.0);
            team.run([&](int member, int members) {
                const int c0 = L * member / members, c1 = L * (member + 1) / members;
                for (int r = 0; r < L; ++r) {
                    if (ct[r] == 0.0) continue;
                    axpy(c1 - c0, ct[r], &winv_[(size_t)r * L + c0], pi.data() + c0);
                }
            });
            for (int j = 0; j < pend_n_; ++j) {              // true inverse = winv_ - sum_j u_j v_j'
                double sj = 0.0;
                for (int r = 0; r < L; ++r) sj += pend_u_[(size_t)j * L + r] * ct[r];
                if (sj != 0.0) axpy(L, -sj, &pend_v_[(size_t)j * L], pi.data());
            }
            pi_valid = true;
        }
        last_phase1 = phase1;
        auto mu_of = [&](int k, int member) {
            std::vector<double> &mu = mu_m[member];
            std::vector<long long> &mu_stamp = mu_stamp_m[member];
            if (mu_stamp[k] != stamp) {
                const int j = key_[k];
                double s = cost_of_key(k);
                for (int t = lptr_[j]; t < lptr_[j + 1]; ++t) s -= pi[lrow_[t]] * lval_[t];
                mu[k] = s; mu_stamp[k] = stamp;
            }
            return mu[k];
        };
        auto reduced_cost_m = [&](int j, int member) {
            double d = phase1 ? 0.0 : cost_[j];
            for (int t = lptr_[j]; t < lptr_[j + 1]; ++t) d -= pi[lrow_[t]] * lval_[t];
            if (set_[j] >= 0) d -= mu_of(set_[j], member);
            return d;
        };
        auto reduced_cost = [&](int j) { return reduced_cost_m(j, 0); };
        t_dual += tick() - t0; t0 = tick();
        // ---- pricing: candidates of the last full pass first, a full pass when they are exhausted
        int q = -1;
        double dq = 0.0;
        if (!bland) {
            // re-price the candidates of the last full pass in chunks (they are ordered by the reduced cost they had
            // then); stop at the first chunk that still holds an attractive column, drop the ones that no longer are
            double best = -kTolDj;
            size_t keep = 0, t = 0;
            while (t < cand.size()) {
                const size_t chunk_end = std::min(cand.size(), t + kCandidateChunk);
                for (; t < chunk_end; ++t) {
                    const int j = cand[t];
                    if (j >= n || stat_[j] != NONBASIC) continue;
                    const double d = reduced_cost(j);
                    if (d < -kTolDj) cand[keep++] = j;
                    if (d < best) { best = d; q = j; dq = d; }
                }
                if (q >= 0) break;
            }
            for (; t < cand.size(); ++t) cand[keep++] = cand[t];     // not looked at this time
            cand.resize(keep);
        }
        t_cand += tick() - t0; t0 = tick();
        if (q < 0) {
            ++n_full;
            cand.clear();
            // full pass.  Columns of one block compete with each other (once one of them has entered, its siblings
            // stop being attractive), so the candidate list takes only the BEST column of every block: it then lasts
            // for hundreds of pivots instead of a handful.
            std::vector<std::pair<double, int>> neg;
            auto offer = [&](double d, int j) {
                const int k = set_[j];
                if (k < 0) { neg.emplace_back(d, j); return; }
                if (best_stamp[k] != stamp) { best_stamp[k] = stamp; best_slot[k] = (int)neg.size(); neg.emplace_back(d, j); }
                else if (d < neg[(size_t)best_slot[k]].first) neg[(size_t)best_slot[k]] = std::make_pair(d, j);
            };
            // cyclic partial pass: start where the last one stopped and stop once enough blocks have offered a
            // column (at least an eighth of the columns is always looked at); only a complete cycle without a
            // find proves optimality.  Bland's rule needs the lowest index: always a complete pass from 0.
            const int start = bland ? 0 : (scan_pos_ < n ? scan_pos_ : 0);
            int scanned = 0;
            if (T > 1 && !bland) {
                // the team prices a chunk of columns (cyclic positions [scanned, scanned + len)) into dbuf, then this
                // thread walks the chunk exactly as the serial loop below would: same candidates, same stop
                const int chunk = std::max(1024, (n / 8 + 1) / 2);
                dbuf.resize((size_t)chunk);
                bool stop_pass = false;
                while (scanned < n && !stop_pass) {
                    const int len = std::min(chunk, n - scanned);
                    team.run([&](int member, int members) {
                        const int e0 = (int)((long long)len * member / members), e1 = (int)((long long)len * (member + 1) / members);
                        for (int e = e0; e < e1; ++e) {
                            int j = start + scanned + e;
                            if (j >= n) j -= n;
                            dbuf[(size_t)e] = stat_[j] == NONBASIC ? reduced_cost_m(j, member) : 0.0;
                        }
                    });
                    for (int e = 0; e < len; ++e, ++scanned) {
                        const double d = dbuf[(size_t)e];
                        if (d < -kTolDj) {
                            int j = start + scanned;
                            if (j >= n) j -= n;
                            offer(d, j);
                            if (neg.size() >= kEnoughCandidates && scanned >= n / 8) { ++scanned; stop_pass = true; break; }
                        }
                    }
                }
            } else
            for (; scanned < n; ++scanned) {
                int j = start + scanned;
                if (j >= n) j -= n;
                if (stat_[j] != NONBASIC) continue;
                const double d = reduced_cost(j);
                if (d < -kTolDj) {
                    if (bland) { q = j; dq = d; break; }
                    offer(d, j);
                    if (neg.size() >= kEnoughCandidates && scanned >= n / 8) { ++scanned; break; }
                }
            }
            if (!bland) { scan_pos_ = start + scanned; if (scan_pos_ >= n) scan_pos_ -= n; }
            if (!bland && !neg.empty()) {
                const size_t keep = std::min(neg.size(), kMaxCandidates);
                std::partial_sort(neg.begin(), neg.begin() + keep, neg.end());
                q = neg[0].second; dq = neg[0].first;
                for (size_t t = 0; t < keep; ++t) cand.push_back(neg[t].second);
            }
        }
        if (q < 0) {
            if (phase1) { status = 1; break; }
            // optimal for the current inverse: verify with a fresh factorisation once
            if (verify_rounds < 3 && since_refactor > 0) {
                ++verify_rounds;
                refactor(); compute_values(); since_refactor = 0; pi_valid = false; scan_keys = true;
                continue;
            }
            status = 0;
            break;
        }
        t_pass += tick() - t0; t0 = tick();
        // ---- FTRAN on the working basis
        const int sq = set_[q];
        link_minus_key(q, wq);
        wq_nz.clear();
        for (int i = 0; i < L; ++i) if (wq[i] != 0.0) wq_nz.push_back(i);
        team.run([&](int member, int members) {
            const int r0 = L * member / members, r1 = L * (member + 1) / members;
            for (int r = r0; r < r1; ++r) {
                const double *row = &winv_[(size_t)r * L];
                double s = 0.0;
                for (size_t t = 0; t < wq_nz.size(); ++t) s += row[wq_nz[t]] * wq[wq_nz[t]];
                y[r] = s;
            }
        });
        for (int j = 0; j < pend_n_; ++j) {
            double sj = 0.0;
            for (size_t t = 0; t < wq_nz.size(); ++t) sj += pend_v_[(size_t)j * L + wq_nz[t]] * wq[wq_nz[t]];
            if (sj != 0.0) axpy(L, -sj, &pend_u_[(size_t)j * L], y.data());
        }
        t_ftran += tick() - t0; t0 = tick();
        // d x_key(s) / d theta = sum_{r in s} y_r - [s == set(q)]
        touched.clear();
        auto touch = [&](int k) { if (rate_stamp[k] != stamp) { rate_stamp[k] = stamp; rate[k] = 0.0; touched.push_back(k); } };
        if (sq >= 0) { touch(sq); rate[sq] -= 1.0; }
        for (int r = 0; r < L; ++r) {
            if (y[r] == 0.0) continue;
            const int v = pos_var_[r];
            if (v >= 0 && set_[v] >= 0) { touch(set_[v]); rate[set_[v]] += y[r]; }
        }
        // ---- ratio test (composite phase 1: an infeasible variable blocks where it becomes feasible)
        int leave_pos = -1, leave_key = -1;
        double tmin = kInfD, pmag = 0.0;
        auto consider = [&](double v, double rho, bool logical, int pos, int keyset, int varid) {
            if (std::fabs(rho) <= kTolPiv) return;
            const bool below = v < -kTolBnd, above = logical && v > kTolBnd;
            double t;
            if (rho < 0.0) {
                if (above) t = (0.0 - v) / rho;
                else if (below) return;
                else t = -v / rho;
            } else {
                if (below) t = -v / rho;
                else if (above) return;
                else { if (!logical) return; t = 0.0; }
            }
            if (t < 0.0) t = 0.0;
            bool take = false;
            if (t < tmin - kTolTie) take = true;
            else if (t <= tmin + kTolTie) {
                if (bland) {
                    const int cur = leave_pos >= 0 ? pos_var_[leave_pos] : (leave_key >= 0 ? key_[leave_key] : 1 << 30);
                    take = varid < cur;
                } else take = std::fabs(rho) > pmag;
            }
            if (take) { if (t < tmin) tmin = t; leave_pos = pos; leave_key = keyset; pmag = std::fabs(rho); }
        };
        for (int r = 0; r < L; ++r) if (y[r] != 0.0) consider(xn_[r], -y[r], is_logical(r), r, -1, pos_var_[r]);
        for (size_t t = 0; t < touched.size(); ++t) {
            const int k = touched[t];
            consider(xkey_[k], rate[k], false, -1, k, key_[k]);
        }
        if (leave_pos < 0 && leave_key < 0) { status = phase1 ? 3 : 2; break; }
        const double theta = tmin;
        // ---- values
        for (int r = 0; r < L; ++r) if (y[r] != 0.0) xn_[r] -= theta * y[r];
        for (size_t t = 0; t < touched.size(); ++t) {
            const int k = touched[t];
            xkey_[k] += theta * rate[k];
            const double w = xkey_[k] < -kTolBnd ? -1.0 : 0.0;
            n_key_infeasible += (w != 0.0) - (wk[k] != 0.0);
            wk[k] = w;
        }
        t_ratio += tick() - t0; t0 = tick();
        // ---- basis change
        if (leave_key >= 0) flush();          // the two key-changing updates below work on the rows themselves
        if (leave_key >= 0 && leave_key == sq) {
            // the key of the entering column's own block leaves: the entering column becomes the key.  The block's
            // other basic columns change from a_i - a_old to a_i - a_q:  W' = W (I - y 1_R'),  Sherman-Morrison.
            const int s = sq;
            double sigma = 0.0;
            std::fill(u.begin(), u.end(), 0.0);
            bool any = false;
            for (int r = 0; r < L; ++r) {
                const int v = pos_var_[r];
                if (v >= 0 && set_[v] == s) {
                    any = true; sigma += y[r];
                    axpy(L, 1.0, &winv_[(size_t)r * L], u.data());
                }
            }
            if (any) {
                pi_valid = false;
                ++since_refactor;
                const double f = 1.0 / (1.0 - sigma);
                team.run([&](int member, int members) {
                    const int r0 = L * member / members, r1 = L * (member + 1) / members;
                    for (int r = r0; r < r1; ++r) {
                        if (y[r] == 0.0) continue;
                        axpy(L, y[r] * f, u.data(), &winv_[(size_t)r * L]);
                    }
                });
            }
            if (!any) ++n_trivial;
            stat_[key_[s]] = NONBASIC;
            key_[s] = q;
            stat_[q] = KEY;
            xkey_[s] = theta;
            n_key_infeasible -= wk[s] != 0.0; wk[s] = 0.0;
        } else {
            int r = leave_pos;
            if (leave_key >= 0) {
                // the key of another block leaves: first hand the key role to the block's basic column t with the
                // largest |y_t| (W' = W M, M an involution touching row t only), then the old key leaves from position t
                const int s = leave_key;
                int t = -1;
                double best = 0.0;
                for (int rr = 0; rr < L; ++rr) {
                    const int v = pos_var_[rr];
                    if (v >= 0 && set_[v] == s && std::fabs(y[rr]) > best) { best = std::fabs(y[rr]); t = rr; }
                }
                if (t < 0) { status = 3; break; }
                ++n_keyswap;
                double *rt = &winv_[(size_t)t * L];
                double ysum = 0.0;
                for (int i = 0; i < L; ++i) rt[i] = -rt[i];
                for (int rr = 0; rr < L; ++rr) {
                    const int v = pos_var_[rr];
                    if (v < 0 || set_[v] != s) continue;
                    ysum += y[rr];
                    if (rr == t) continue;
                    axpy(L, -1.0, &winv_[(size_t)rr * L], rt);
                }
                const int newkey = pos_var_[t], oldkey = key_[s];
                pos_var_[t] = oldkey; stat_[oldkey] = NONKEY;
                key_[s] = newkey; stat_[newkey] = KEY;
                std::swap(xn_[t], xkey_[s]);
                {
                    const double w = xkey_[s] < -kTolBnd ? -1.0 : 0.0;
                    n_key_infeasible += (w != 0.0) - (wk[s] != 0.0);
                    wk[s] = w;
                }
                y[t] = -ysum;                                    // y' = M y
                r = t;
            }
            // column replacement at position r
            const int vleave = pos_var_[r];
            const double piv = y[r];
            if (std::fabs(piv) < 1e-12) { refactor(); compute_values(); since_refactor = 0; pi_valid = false; scan_keys = true; ++iters_; continue; }
            const double inv = 1.0 / piv;
            double *rp = &winv_[(size_t)r * L];
            if (block_k > 0) {
                // pending form: v = (true row r) / piv, u = y - e_r; the true row r afterwards equals v
                for (int i = 0; i < L; ++i) rpc[i] = rp[i];
                for (int j = 0; j < pend_n_; ++j) {
                    const double a = pend_u_[(size_t)j * L + r];
                    if (a != 0.0) axpy(L, -a, &pend_v_[(size_t)j * L], rpc.data());
                }
                scale(L, inv, rpc.data());
                double *pu = &pend_u_[(size_t)pend_n_ * L], *pv = &pend_v_[(size_t)pend_n_ * L];
                for (int i = 0; i < L; ++i) { pu[i] = y[i]; pv[i] = rpc[i]; }
                pu[r] = piv - 1.0;
                ++pend_n_;
                if (vleave >= 0) stat_[vleave] = NONBASIC;
                pos_var_[r] = q;
                stat_[q] = NONKEY;
                xn_[r] = theta;
                if (leave_key < 0 && !phase1 && pi_valid) axpy(L, dq, rpc.data(), pi.data());
                else pi_valid = false;
                ++since_refactor;
                if (pend_n_ >= block_k) flush();
            } else {
            for (int i = 0; i < L; ++i) rpc[i] = rp[i] * inv;      // the scaled pivot row, handed to every row owner
            team.run([&](int member, int members) {
                const int r0 = L * member / members, r1 = L * (member + 1) / members;
                for (int rr = r0; rr < r1; ++rr) {
                    double *row = &winv_[(size_t)rr * L];
                    if (rr == r) { for (int i = 0; i < L; ++i) row[i] = rpc[i]; continue; }
                    if (y[rr] == 0.0) continue;
                    axpy(L, -y[rr], rpc.data(), row);
                }
            });
            if (vleave >= 0) stat_[vleave] = NONBASIC;
            pos_var_[r] = q;
            stat_[q] = NONKEY;
            xn_[r] = theta;
            // pi' = pi + d_q * (new row r of W^-1): O(L) instead of the O(L^2) product — valid when only this column
            // changed and the prices were the phase-2 ones (a key hand-over re-keys other columns: recompute then)
            if (leave_key < 0 && !phase1 && pi_valid) axpy(L, dq, rp, pi.data());
            else pi_valid = false;
            ++since_refactor;
            }
        }
        t_update += tick() - t0;
        ++iters_;
        verify_rounds = 0;
        if (theta <= 1e-11) ++n_degen;
        if (theta <= 1e-11) { if (++degen_run > 500) bland = true; }
        else { degen_run = 0; bland = false; }
    }
    flush();
    if (std::getenv("DWSOLVER_MASTER_DEBUG"))
        std::fprintf(stderr, "gub master: status %d, %lld iterations (%lld degenerate, %lld trivial key swaps, %lld key hand-overs), %lld full pricing passes, %d columns\n",
                     status, iters_ - it_begin, n_degen, n_trivial, n_keyswap, n_full, cols());
    if (prof && iters_ - it_begin > 1000)
        std::fprintf(stderr, "  seconds: duals %.3f, candidate pricing %.3f, pricing passes %.3f, ftran %.3f, ratio+values %.3f, basis update %.3f\n",
                     t_dual, t_cand, t_pass, t_ftran, t_ratio, t_update);
    publish(true);
    return status;
}

}  // namespace dwh
