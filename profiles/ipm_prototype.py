#!/usr/bin/env python
"""Prototype (numpy) of the interior-point restricted master that csrc/master_ipm.cuh implements on the GPU: Mehrotra
predictor-corrector on  min c'x, A x = b, x >= 0  with the GUB structure eliminated — the K convexity rows form a
diagonal block of A D A', so the normal equations reduce to an L x L system
    S = sum_k sum_{j in k} d_j (a_j - abar_k)(a_j - abar_k)'  +  sum_{slack j} d_j a_j a_j',   abar_k = sum d_j a_j / sum d_j.
Runs the whole DW loop of a config on the CPU (oracle subproblems) and prints per round: columns, IPM iterations,
master objective, Lagrangian bound.  Usage: python profiles/ipm_prototype.py [A|B] [tol]"""
import os
import sys
import time

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from dwsolver_b200 import problem as P  # noqa: E402
from dw_loop import HighsMaster, TOLERANCE, DW_INFINITY  # noqa: E402
from oracle import oracle as O  # noqa: E402


def ipm(AL, blk, c, b, L, K, tol=1e-9, max_it=100, verbose=False):
    """AL: L x n csc; blk[j] in [0,K) or -1; c (n); b (L+K).  Returns dict(x, y, s, obj, iters, status)."""
    n = AL.shape[1]
    inb = blk >= 0
    ALd = AL.toarray()
    bL, bK = b[:L], b[L:]
    blk_safe = np.where(inb, blk, 0)

    class F:
        pass

    def factor(d):
        f = F()
        f.delta = np.bincount(blk[inb], weights=d[inb], minlength=K)
        G = np.zeros((L, K))
        W = ALd[:, inb] * d[inb]
        np.add.at(G.T, blk[inb], W.T)
        f.G = G
        f.abar = G / f.delta
        Bt = ALd.copy()
        Bt[:, inb] -= f.abar[:, blk[inb]]
        Bt *= np.sqrt(d)
        S = Bt @ Bt.T
        reg = 1e-14 * max(np.trace(S) / L, 1e-300)
        for _ in range(8):
            try:
                f.chol = sla.cho_factor(S + reg * np.eye(L), lower=True, check_finite=False)
                break
            except np.linalg.LinAlgError:
                reg *= 100.0
        return f

    def solve(f, rL, rK):
        t = rK / f.delta
        yL = sla.cho_solve(f.chol, rL - f.G @ t, check_finite=False)
        yK = (rK - f.G.T @ yL) / f.delta
        return yL, yK

    def At(yL, yK):
        return ALd.T @ yL + np.where(inb, yK[blk_safe], 0.0)

    def Ax(h):
        return ALd @ h, np.bincount(blk[inb], weights=h[inb], minlength=K)

    # Mehrotra's starting point
    f1 = factor(np.ones(n))
    yL, yK = solve(f1, bL, bK)
    x = At(yL, yK)
    AcL, AcK = Ax(c)
    yL, yK = solve(f1, AcL, AcK)
    s = c - At(yL, yK)
    dx = max(-1.5 * x.min(), 0.0)
    ds = max(-1.5 * s.min(), 0.0)
    x = x + dx
    s = s + ds
    xs = float(x @ s)
    if xs <= 0:
        x += 1.0; s += 1.0; xs = float(x @ s)
    x = x + 0.5 * xs / s.sum()
    s = s + 0.5 * xs / x.sum()
    nb, nc = 1.0 + np.abs(b).max(), 1.0 + np.abs(c).max()
    status = 3
    it = 0
    for it in range(1, max_it + 1):
        AxL, AxK = Ax(x)
        rpL, rpK = bL - AxL, bK - AxK
        rd = c - At(yL, yK) - s
        mu = float(x @ s) / n
        pobj, dobj = float(c @ x), float(bL @ yL + bK @ yK)
        pinf = max(np.abs(rpL).max(), np.abs(rpK).max()) / nb
        dinf = np.abs(rd).max() / nc
        gap = abs(pobj - dobj) / (1.0 + abs(pobj))
        if verbose:
            print(f"    ipm {it:3d} pobj {pobj:.10e} dobj {dobj:.10e} pinf {pinf:.1e} dinf {dinf:.1e} gap {gap:.1e} mu {mu:.1e}")
        if pinf <= tol and dinf <= tol and gap <= tol:
            status = 0
            break
        d = x / s
        f = factor(d)

        def newton(rc):
            h = d * rd - rc / s
            hL, hK = Ax(h)
            dyL, dyK = solve(f, rpL + hL, rpK + hK)
            dsv = rd - At(dyL, dyK)
            dxv = rc / s - d * dsv
            return dxv, dyL, dyK, dsv

        def steplen(v, dv):
            neg = dv < 0
            return min(1.0, float((-v[neg] / dv[neg]).min())) if neg.any() else 1.0

        dxa, _, _, dsa = newton(-x * s)
        ap, ad = steplen(x, dxa), steplen(s, dsa)
        mu_aff = float((x + ap * dxa) @ (s + ad * dsa)) / n
        sigma = (mu_aff / mu) ** 3
        dxv, dyL_, dyK_, dsv = newton(sigma * mu - x * s - dxa * dsa)
        eta = max(0.995, 1.0 - mu)
        ap, ad = min(1.0, eta * steplen(x, dxv) if (dxv < 0).any() else 1.0), min(1.0, eta * steplen(s, dsv) if (dsv < 0).any() else 1.0)
        x = x + ap * dxv
        s = s + ad * dsv
        yL = yL + ad * dyL_
        yK = yK + ad * dyK_
    return dict(x=x, yL=yL, yK=yK, s=s, obj=float(c @ x), dobj=float(bL @ yL + bK @ yK), iters=it, status=status)


class IpmMaster(HighsMaster):
    def __init__(self, prob, tol=1e-9, verbose=False):
        super().__init__(prob)
        self.tol, self.verbose = tol, verbose
        self.seconds, self.iterations = 0.0, 0

    def solve(self):
        L, K = self.L, self.K
        n = len(self.cols)
        rows, cj, vals, blk = [], [], [], np.full(n, -1, np.int64)
        for j, c in enumerate(self.cols):
            for r_, v_ in zip(c["rows"], c["vals"]):
                if r_ < L:
                    rows.append(r_); cj.append(j); vals.append(v_)
                else:
                    blk[j] = r_ - L
        AL = sp.csc_matrix((vals, (rows, cj)), shape=(L, n))
        cost = np.array([c["cost"] for c in self.cols], float)
        t0 = time.perf_counter()
        res = ipm(AL, blk, cost, self.rhs, L, K, tol=self.tol, verbose=self.verbose)
        self.seconds += time.perf_counter() - t0
        self.iterations += res["iters"]
        self.last = res
        if res["status"] != 0:
            print("  !! ipm did not converge", res["iters"])
            import pickle
            pickle.dump(dict(AL=AL, blk=blk, cost=cost, rhs=self.rhs, L=L, K=K), open("/tmp/ipm_fail.pkl", "wb"))
        return dict(obj=res["obj"], x=res["x"], pi=res["yL"].copy(), r=res["yK"].copy())


def run(prob, tol):
    K, L = prob.K, prob.L
    blocks = O.OracleBlocks(prob)
    threads = os.cpu_count()
    master = IpmMaster(prob, tol=tol)
    r = np.full(K, DW_INFINITY)
    it = 0

    def consume(props, phase_one):
        added = 0
        for k, p in enumerate(props):
            if r[k] - p["obj"] > TOLERANCE:
                master.add_lambda(k, it, p["cost"], p["ind"], p["val"], phase_one)
                added += 1
        return added

    props = blocks.initial(threads)
    first = True
    for _ in range(100):
        n_before = len(master.cols)
        added = consume(props, True)
        if first:
            acc = np.zeros(L)
            for c in master.cols[n_before:]:
                np.add.at(acc, c["rows"][:-1], c["vals"][:-1])
            for i in range(L):
                if master.rhs[i] - acc[i] < 0.0:
                    for c in master.cols:
                        if c["art"] and c["name"] == f"y_{i+1}":
                            c["vals"] = np.array([-1.0])
            first = False
        sol = master.solve()
        r = sol["r"]; it += 1
        print(f"P1 it={it} added={added} cols={len(master.cols)} ipm_it={master.last['iters']} obj={sol['obj']:.6e}", flush=True)
        if abs(sol["obj"]) < TOLERANCE or added == 0:
            break
        props = blocks.iterate(sol["pi"], True, threads)
    master.cols = [c for c in master.cols if not c["art"]]
    for c in master.cols:
        c["cost"] = c["true_cost"]
    sol = master.solve()
    r = sol["r"]; it += 1
    print(f"P1->2 cols={len(master.cols)} ipm_it={master.last['iters']} obj={sol['obj']:.10e}", flush=True)
    for _ in range(200):
        props = blocks.iterate(sol["pi"], False, threads)
        lb = float(master.rhs[:L] @ sol["pi"]) + sum(p["obj"] for p in props)
        added = consume(props, False)
        worst = max(r[k] - p["obj"] for k, p in enumerate(props))
        print(f"P2 it={it} added={added} cols={len(master.cols)} UB={sol['obj']:.10e} LB={lb:.10e} relgap={(sol['obj']-lb)/abs(sol['obj']):.2e} max(r-obj)={worst:.2e}", flush=True)
        if added == 0:
            break
        sol = master.solve()
        r = sol["r"]; it += 1
        print(f"      master: ipm_it={master.last['iters']} obj={sol['obj']:.10e} dobj={master.last['dobj']:.10e}", flush=True)
    print(f"done: objective {sol['obj'] + prob.shift:.10e}, DW iterations {it}, master seconds {master.seconds:.1f}, IPM iterations {master.iterations}")


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "A"
    tol = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-9
    run(P.make_config(tag), tol)
