"""Block-angular LP container, CPLEX-LP subset reader/writer and synthetic generators.

This is the host-side data model the rest of the repo shares.  It mirrors what the
reference assembles at start-up:

* the master LP file supplies the L linking rows, the master costs ``c`` and the
  coupling coefficients ``D`` (reference: ``src/dw_solver.c:192-238``);
* every block LP file supplies ``A_k``, row/column bounds and the block's *file*
  objective used for the cold solve (reference: ``src/dw_subprob.c:105-122``);
* the name-based mapping block column -> master column builds ``c_k`` and the COO
  triplets of ``D_k`` in *local column order* (reference: ``src/dw_subprob.c:203-258``).

Nothing here is on the timed hot path: it feeds the C-ABI engine (``include/dwgpu.h``),
the CPU oracle (``oracle/``) and the test/bench harnesses with identical arrays.
"""
from __future__ import annotations

import math
import os
import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

INF = float("inf")
#: value used by the reference to clamp lower-bounded block columns (dw_subprob.c:116-120)
SUB_UB_CLAMP = 3000.0


# --------------------------------------------------------------------------------------
# Plain LP model (one file)
# --------------------------------------------------------------------------------------
@dataclass
class LPModel:
    """One LP file: min c'x  s.t. row_lb <= A x <= row_ub, col_lb <= x <= col_ub."""

    col_names: List[str]
    row_names: List[str]
    obj: np.ndarray                      # (n,)
    obj_const: float
    a_rows: np.ndarray                   # COO, entry order = file order
    a_cols: np.ndarray
    a_vals: np.ndarray
    row_lb: np.ndarray
    row_ub: np.ndarray
    col_lb: np.ndarray
    col_ub: np.ndarray
    maximize: bool = False

    @property
    def n(self) -> int:
        return len(self.col_names)

    @property
    def m(self) -> int:
        return len(self.row_names)

    def csr(self):
        """CSR (rowptr, colidx, vals) with duplicates summed, columns sorted per row."""
        return coo_to_csr(self.m, self.n, self.a_rows, self.a_cols, self.a_vals)

    def columns_coo(self):
        """Per-column entry lists in row order — GLPK stores one entry per (row, col)."""
        rowptr, colidx, vals = self.csr()
        cols: List[List[Tuple[int, float]]] = [[] for _ in range(self.n)]
        for i in range(self.m):
            for p in range(rowptr[i], rowptr[i + 1]):
                cols[colidx[p]].append((i, vals[p]))
        return cols


def coo_to_csr(m: int, n: int, rows, cols, vals):
    rows = np.asarray(rows, dtype=np.int64)
    cols = np.asarray(cols, dtype=np.int64)
    vals = np.asarray(vals, dtype=np.float64)
    if rows.size == 0:
        return np.zeros(m + 1, np.int32), np.zeros(0, np.int32), np.zeros(0, np.float64)
    key = rows * max(n, 1) + cols
    order = np.argsort(key, kind="stable")
    key_s = key[order]
    vals_s = vals[order]
    uniq, start = np.unique(key_s, return_index=True)
    summed = np.add.reduceat(vals_s, start)
    r = (uniq // max(n, 1)).astype(np.int64)
    c = (uniq % max(n, 1)).astype(np.int32)
    rowptr = np.zeros(m + 1, np.int64)
    np.add.at(rowptr, r + 1, 1)
    rowptr = np.cumsum(rowptr).astype(np.int32)
    return rowptr, c, summed.astype(np.float64)


# --------------------------------------------------------------------------------------
# CPLEX-LP subset reader (the subset glp_read_lp accepts that the fixtures use;
# SURVEY.md Appendix B / §4a)
# --------------------------------------------------------------------------------------
_NAME_START = r"A-Za-z!\"#$%&()/,;?@_`'{}|~"
_TOKEN_RE = re.compile(
    r"\s*(?:"
    r"(?P<num>(?:\d+\.?\d*|\.\d+)(?:[eE][+-]?\d+)?)"
    r"|(?P<name>[" + _NAME_START + r"][" + _NAME_START + r"0-9.]*)"
    r"|(?P<op><=|>=|=<|=>|<|>|=|\+|-|:)"
    r")"
)

_SECTION_WORDS = {
    "minimize": "obj", "minimise": "obj", "minimum": "obj", "min": "obj",
    "maximize": "objmax", "maximise": "objmax", "maximum": "objmax", "max": "objmax",
    "st": "rows", "s.t.": "rows", "st.": "rows",
    "bounds": "bounds", "bound": "bounds",
    "general": "ints", "generals": "ints", "gen": "ints",
    "integer": "ints", "integers": "ints", "int": "ints",
    "binary": "bins", "binaries": "bins", "bin": "bins",
    "end": "end",
}


def _tokenize(text: str):
    toks = []
    for raw in text.splitlines():
        line = raw.split("\\", 1)[0]          # backslash starts a comment
        pos = 0
        while pos < len(line):
            mobj = _TOKEN_RE.match(line, pos)
            if not mobj:
                if line[pos:].strip() == "":
                    break
                raise ValueError(f"LP syntax error near: {line[pos:pos+30]!r}")
            pos = mobj.end()
            if mobj.lastgroup == "num":
                toks.append(("num", float(mobj.group("num"))))
            elif mobj.lastgroup == "name":
                toks.append(("name", mobj.group("name")))
            else:
                toks.append(("op", mobj.group("op")))
    return toks


def _sectionize(toks):
    """Split the token stream into (section, tokens) pieces, recognising the two-word
    headers ``subject to`` / ``such that``."""
    out = []
    cur_kind, cur = None, []
    i = 0
    while i < len(toks):
        kind, val = toks[i]
        sec = None
        adv = 1
        if kind == "name":
            low = val.lower()
            # a section keyword followed by ':' is a label, not a header
            is_label = i + 1 < len(toks) and toks[i + 1] == ("op", ":")
            if not is_label:
                if low in ("subject", "such") and i + 1 < len(toks) and toks[i + 1][0] == "name" \
                        and toks[i + 1][1].lower() in ("to", "that"):
                    sec, adv = "rows", 2
                elif low in _SECTION_WORDS:
                    sec = _SECTION_WORDS[low]
        if sec is not None:
            if cur_kind is not None:
                out.append((cur_kind, cur))
            cur_kind, cur = sec, []
            i += adv
            if sec == "end":
                break
            continue
        cur.append(toks[i])
        i += 1
    if cur_kind is not None:
        out.append((cur_kind, cur))
    return out


def _parse_inf(tok) -> Optional[float]:
    if tok[0] == "name" and tok[1].lower() in ("inf", "infinity"):
        return INF
    return None


class _Cols:
    def __init__(self):
        self.names: List[str] = []
        self.index: Dict[str, int] = {}

    def get(self, name: str) -> int:
        j = self.index.get(name)
        if j is None:
            j = len(self.names)
            self.index[name] = j
            self.names.append(name)
        return j


def _parse_linear(toks, pos, cols: _Cols, stop_ops=("<=", ">=", "=<", "=>", "<", ">", "=")):
    """Parse ``[+|-] [coef] name ...`` until a sense operator / end / next label.
    Returns (terms, pos)."""
    terms: List[Tuple[int, float]] = []
    sign = 1.0
    coef: Optional[float] = None
    n = len(toks)
    while pos < n:
        kind, val = toks[pos]
        if kind == "op" and val in stop_ops:
            break
        if kind == "op" and val == "+":
            pos += 1
            continue
        if kind == "op" and val == "-":
            sign = -sign
            pos += 1
            continue
        if kind == "num":
            coef = val if coef is None else coef * val
            pos += 1
            continue
        if kind == "name":
            # label of the next row?  (name followed by ':')
            if pos + 1 < n and toks[pos + 1] == ("op", ":"):
                break
            terms.append((cols.get(val), sign * (1.0 if coef is None else coef)))
            sign, coef = 1.0, None
            pos += 1
            continue
        raise ValueError(f"unexpected token {toks[pos]} in linear expression")
    const = 0.0
    if coef is not None:            # trailing constant term
        const = sign * coef
    return terms, const, pos


def parse_lp_text(text: str) -> LPModel:
    toks = _tokenize(text)
    sections = _sectionize(toks)
    cols = _Cols()
    obj_terms: List[Tuple[int, float]] = []
    obj_const = 0.0
    maximize = False
    row_names: List[str] = []
    row_lb: List[float] = []
    row_ub: List[float] = []
    a_r: List[int] = []
    a_c: List[int] = []
    a_v: List[float] = []
    lb: Dict[int, float] = {}
    ub: Dict[int, float] = {}

    for kind, st in sections:
        if kind in ("obj", "objmax"):
            maximize = kind == "objmax"
            pos = 0
            if len(st) >= 2 and st[0][0] == "name" and st[1] == ("op", ":"):
                pos = 2
            obj_terms, obj_const, pos = _parse_linear(st, pos, cols)
        elif kind == "rows":
            pos = 0
            while pos < len(st):
                name = None
                if st[pos][0] == "name" and pos + 1 < len(st) and st[pos + 1] == ("op", ":"):
                    name = st[pos][1]
                    pos += 2
                terms, const, pos = _parse_linear(st, pos, cols)
                if pos >= len(st) or st[pos][0] != "op":
                    raise ValueError("constraint without sense")
                sense = st[pos][1]
                pos += 1
                sgn = 1.0
                while pos < len(st) and st[pos][0] == "op" and st[pos][1] in "+-":
                    if st[pos][1] == "-":
                        sgn = -sgn
                    pos += 1
                if pos >= len(st) or st[pos][0] != "num":
                    raise ValueError("constraint without numeric right-hand side")
                rhs = sgn * st[pos][1] - const
                pos += 1
                i = len(row_names)
                row_names.append(name if name is not None else f"r.{i + 1}")
                if sense in ("<=", "=<", "<"):
                    row_lb.append(-INF); row_ub.append(rhs)
                elif sense in (">=", "=>", ">"):
                    row_lb.append(rhs); row_ub.append(INF)
                else:
                    row_lb.append(rhs); row_ub.append(rhs)
                for j, v in terms:
                    a_r.append(i); a_c.append(j); a_v.append(v)
        elif kind == "bounds":
            pos = 0

            def read_value(p):
                sgn = 1.0
                while p < len(st) and st[p][0] == "op" and st[p][1] in "+-":
                    if st[p][1] == "-":
                        sgn = -sgn
                    p += 1
                if p < len(st) and st[p][0] == "num":
                    return sgn * st[p][1], p + 1
                if p < len(st) and _parse_inf(st[p]) is not None:
                    return sgn * INF, p + 1
                return None, p

            while pos < len(st):
                v1, p1 = read_value(pos)
                if v1 is not None:                      # "l <= x [<= u]"  or  "u >= x"
                    op1 = st[p1][1]
                    j = cols.get(st[p1 + 1][1])
                    pos = p1 + 2
                    if op1 in ("<=", "=<", "<"):
                        lb[j] = v1
                    elif op1 in (">=", "=>", ">"):
                        ub[j] = v1
                    if pos < len(st) and st[pos][0] == "op" and st[pos][1] in ("<=", "=<", "<", ">=", "=>", ">"):
                        v2, p2 = read_value(pos + 1)
                        if v2 is not None:
                            if st[pos][1] in ("<=", "=<", "<"):
                                ub[j] = v2
                            else:
                                lb[j] = v2
                            pos = p2
                    continue
                if st[pos][0] != "name":
                    raise ValueError(f"bad bounds entry at {st[pos]}")
                j = cols.get(st[pos][1])
                pos += 1
                if pos < len(st) and st[pos][0] == "name" and st[pos][1].lower() == "free":
                    lb[j], ub[j] = -INF, INF
                    pos += 1
                    continue
                op = st[pos][1]
                v, pos = read_value(pos + 1)
                if op in ("<=", "=<", "<"):
                    ub[j] = v
                    # glpk: an upper bound < 0 with default lb 0 resets lb to -inf? keep lb.
                elif op in (">=", "=>", ">"):
                    lb[j] = v
                else:
                    lb[j] = ub[j] = v
        elif kind in ("ints", "bins"):
            for tk in st:
                if tk[0] == "name":
                    j = cols.get(tk[1])
                    if kind == "bins":
                        lb.setdefault(j, 0.0)
                        ub[j] = 1.0
        elif kind == "end":
            break

    n = len(cols.names)
    obj = np.zeros(n)
    for j, v in obj_terms:
        obj[j] += v
    col_lb = np.zeros(n)
    col_ub = np.full(n, INF)
    for j, v in lb.items():
        col_lb[j] = v
    for j, v in ub.items():
        col_ub[j] = v
    return LPModel(
        col_names=cols.names, row_names=row_names, obj=obj, obj_const=obj_const,
        a_rows=np.asarray(a_r, np.int32), a_cols=np.asarray(a_c, np.int32),
        a_vals=np.asarray(a_v, np.float64),
        row_lb=np.asarray(row_lb, np.float64), row_ub=np.asarray(row_ub, np.float64),
        col_lb=col_lb, col_ub=col_ub, maximize=maximize)


def read_lp(path: str) -> LPModel:
    with open(path, "r") as fh:
        return parse_lp_text(fh.read())


def _fmt(v: float) -> str:
    return repr(float(v))


def write_lp(model: LPModel, path: str, name: str = "obj", obj_all: bool = False) -> None:
    """Emit an LP file in the subset above (our own layout: one term per line chunk).
    ``obj_all`` lists every column in the objective (zero coefficients included) so that a column
    without any other occurrence still exists for a reader."""
    rowptr, colidx, vals = model.csr()
    with open(path, "w") as fh:
        fh.write("Maximize\n" if model.maximize else "Minimize\n")
        terms = [f"{'+' if model.obj[j] >= 0 else '-'} {_fmt(abs(model.obj[j]))} {model.col_names[j]}"
                 for j in range(model.n) if obj_all or model.obj[j] != 0.0]
        if not terms and model.n:
            terms = [f"+ 0 {model.col_names[0]}"]
        fh.write(f" {name}:")
        for t0 in range(0, len(terms), 6):
            fh.write(" " + " ".join(terms[t0:t0 + 6]) + "\n")
        if not terms:
            fh.write("\n")
        fh.write("Subject To\n")
        for i in range(model.m):
            parts = [f"{'+' if vals[p] >= 0 else '-'} {_fmt(abs(vals[p]))} {model.col_names[colidx[p]]}"
                     for p in range(rowptr[i], rowptr[i + 1])]
            lo, hi = model.row_lb[i], model.row_ub[i]
            if lo == hi:
                sense, rhs = "=", lo
            elif math.isinf(lo):
                sense, rhs = "<=", hi
            elif math.isinf(hi):
                sense, rhs = ">=", lo
            else:
                raise ValueError("ranged rows are not representable in this LP subset")
            fh.write(f" {model.row_names[i]}:")
            for t0 in range(0, max(len(parts), 1), 6):
                fh.write(" " + " ".join(parts[t0:t0 + 6]) + ("\n" if t0 + 6 < len(parts) else ""))
            fh.write(f" {sense} {_fmt(rhs)}\n")
        bl = []
        for j in range(model.n):
            lo, hi = model.col_lb[j], model.col_ub[j]
            nm = model.col_names[j]
            if lo == 0.0 and math.isinf(hi) and hi > 0:
                continue
            if math.isinf(lo) and math.isinf(hi):
                bl.append(f" {nm} free")
            elif lo == hi:
                bl.append(f" {nm} = {_fmt(lo)}")
            elif math.isinf(hi):
                bl.append(f" {nm} >= {_fmt(lo)}")
            elif math.isinf(lo):
                bl.append(f" -inf <= {nm} <= {_fmt(hi)}")
            else:
                bl.append(f" {_fmt(lo)} <= {nm} <= {_fmt(hi)}")
        if bl:
            fh.write("Bounds\n" + "\n".join(bl) + "\n")
        fh.write("End\n")


# --------------------------------------------------------------------------------------
# Block-angular container
# --------------------------------------------------------------------------------------
@dataclass
class Block:
    """One subproblem.  All index arrays 0-based int32, values float64.

    ``d_row/d_col/d_val`` are the COO triplets of D_k in the order the reference
    appends them (by local column, then master-column storage order;
    dw_subprob.c:230-253) — summation order in the pricing kernels follows it.
    """
    m: int
    n: int
    a_rowptr: np.ndarray
    a_col: np.ndarray
    a_val: np.ndarray
    row_lb: np.ndarray
    row_ub: np.ndarray
    col_lb: np.ndarray
    col_ub: np.ndarray
    c_file: np.ndarray
    c_master: np.ndarray
    d_row: np.ndarray
    d_col: np.ndarray
    d_val: np.ndarray
    col_names: Optional[List[str]] = None
    master_cols: Optional[np.ndarray] = None     # index of each local column in the master file

    def dense_a(self) -> np.ndarray:
        a = np.zeros((self.m, self.n))
        for i in range(self.m):
            s, e = self.a_rowptr[i], self.a_rowptr[i + 1]
            a[i, self.a_col[s:e]] += self.a_val[s:e]
        return a

    def dense_d(self, L: int) -> np.ndarray:
        d = np.zeros((L, self.n))
        np.add.at(d, (self.d_row, self.d_col), self.d_val)
        return d


@dataclass
class BlockAngularLP:
    K: int
    L: int
    blocks: List[Block]
    link_lb: np.ndarray
    link_ub: np.ndarray
    shift: float = 0.0
    link_names: Optional[List[str]] = None
    name: str = ""
    meta: dict = field(default_factory=dict)

    # ---- monolithic form (for independent cross-checks with HiGHS) -------------------
    def monolithic(self):
        """Return (c, A_csr(scipy), row_lb, row_ub, col_lb, col_ub) of the full LP with the
        reference's 3000 clamp applied to block columns (so both sides solve the same LP)."""
        import scipy.sparse as sp
        offs = np.cumsum([0] + [b.n for b in self.blocks])
        ntot = int(offs[-1])
        rows, cols, vals = [], [], []
        for k, b in enumerate(self.blocks):
            rows.append(b.d_row.astype(np.int64)); cols.append(b.d_col.astype(np.int64) + offs[k]); vals.append(b.d_val)
        r0 = self.L
        rlb = [self.link_lb]; rub = [self.link_ub]
        for k, b in enumerate(self.blocks):
            rr = np.repeat(np.arange(b.m, dtype=np.int64), np.diff(b.a_rowptr))
            rows.append(rr + r0); cols.append(b.a_col.astype(np.int64) + offs[k]); vals.append(b.a_val)
            rlb.append(b.row_lb); rub.append(b.row_ub)
            r0 += b.m
        A = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(r0, ntot))
        c = np.concatenate([b.c_master for b in self.blocks])
        clb = np.concatenate([b.col_lb for b in self.blocks])
        cub = np.concatenate([effective_col_ub(b) for b in self.blocks])
        return c, A, np.concatenate(rlb), np.concatenate(rub), clb, cub


def effective_col_ub(b: Block) -> np.ndarray:
    """dw_subprob.c:116-120 — columns with a finite lower and no upper bound get ub=3000."""
    ub = b.col_ub.copy()
    mask = np.isinf(ub) & (ub > 0) & np.isfinite(b.col_lb)
    ub[mask] = SUB_UB_CLAMP
    return ub


def assemble(master: LPModel, subs: Sequence[LPModel], shift: float = 0.0, name: str = "") -> BlockAngularLP:
    """Name-based mapping of block columns into the master (dw_subprob.c:203-258)."""
    L = master.m
    mcols = master.columns_coo()
    index = {nm: j for j, nm in enumerate(master.col_names)}
    blocks = []
    for sub in subs:
        rowptr, colidx, vals = sub.csr()
        c_master = np.zeros(sub.n)
        dr, dc, dv = [], [], []
        mc = np.zeros(sub.n, np.int32)
        for i, nm in enumerate(sub.col_names):
            j = index.get(nm)
            if j is None:
                raise KeyError(f"block column {nm!r} not found in the master problem")
            mc[i] = j
            c_master[i] = master.obj[j]
            for (r, v) in mcols[j]:
                dr.append(r); dc.append(i); dv.append(v)
        blocks.append(Block(
            m=sub.m, n=sub.n, a_rowptr=rowptr, a_col=colidx, a_val=vals,
            row_lb=sub.row_lb.copy(), row_ub=sub.row_ub.copy(),
            col_lb=sub.col_lb.copy(), col_ub=sub.col_ub.copy(),
            c_file=sub.obj.copy(), c_master=c_master,
            d_row=np.asarray(dr, np.int32), d_col=np.asarray(dc, np.int32), d_val=np.asarray(dv, np.float64),
            col_names=list(sub.col_names), master_cols=mc))
    return BlockAngularLP(K=len(blocks), L=L, blocks=blocks, link_lb=master.row_lb.copy(),
                          link_ub=master.row_ub.copy(), shift=shift,
                          link_names=list(master.row_names), name=name)


def read_guidefile(path: str):
    """Guidefile: n, n block files, master file, (ignored) monolithic file, optional shift
    (dw_main.c:213-259).  Paths are relative to the *current directory* in the reference;
    here they are resolved relative to the guidefile for convenience."""
    base = os.path.dirname(os.path.abspath(path))
    with open(path) as fh:
        lines = [ln.rstrip("\r\n") for ln in fh.readlines()]
    n = int(lines[0].strip())
    subs = [os.path.join(base, lines[1 + i].strip()) for i in range(n)]
    master = os.path.join(base, lines[1 + n].strip())
    shift = 0.0
    if len(lines) > 3 + n and lines[3 + n].strip():
        try:
            shift = float(lines[3 + n].strip())
        except ValueError:
            shift = 0.0
    return master, subs, shift


def load_guidefile(path: str) -> BlockAngularLP:
    master, subs, shift = read_guidefile(path)
    return assemble(read_lp(master), [read_lp(s) for s in subs], shift=shift,
                    name=os.path.basename(os.path.dirname(os.path.abspath(path))))


def write_block_angular(p: BlockAngularLP, directory: str) -> dict:
    """The instance in the reference's input format (dw_main.c:213-259): one CPLEX-LP file per block
    (its own rows, bounds and file objective), the master LP file (linking rows over all block columns,
    costs c_k) and the guidefile listing them.  Returns the paths.  Column names are the blocks'
    ``col_names`` or ``x_<k>_<j>``."""
    os.makedirs(directory, exist_ok=True)
    names = [list(b.col_names) if b.col_names is not None else [f"x_{k}_{j}" for j in range(b.n)]
             for k, b in enumerate(p.blocks)]
    subs = []
    for k, b in enumerate(p.blocks):
        rows = np.repeat(np.arange(b.m, dtype=np.int32), np.diff(b.a_rowptr))
        sub = LPModel(col_names=names[k], row_names=[f"n_{k}_{i}" for i in range(b.m)], obj=np.asarray(b.c_file, float),
                      obj_const=0.0, a_rows=rows, a_cols=np.asarray(b.a_col, np.int32), a_vals=np.asarray(b.a_val, float),
                      row_lb=np.asarray(b.row_lb, float), row_ub=np.asarray(b.row_ub, float),
                      col_lb=np.asarray(b.col_lb, float), col_ub=np.asarray(b.col_ub, float), maximize=False)
        path = os.path.join(directory, f"sub{k + 1}.lp")
        write_lp(sub, path, obj_all=True)
        subs.append(path)
    off = np.concatenate([[0], np.cumsum([b.n for b in p.blocks])])
    m_rows = np.concatenate([np.asarray(b.d_row, np.int32) for b in p.blocks]) if p.blocks else np.zeros(0, np.int32)
    m_cols = np.concatenate([np.asarray(b.d_col, np.int32) + off[k] for k, b in enumerate(p.blocks)]).astype(np.int32)
    m_vals = np.concatenate([np.asarray(b.d_val, float) for b in p.blocks])
    link_names = list(p.link_names) if getattr(p, "link_names", None) else [f"cap_{i}" for i in range(p.L)]
    master = LPModel(col_names=[nm for ns in names for nm in ns], row_names=link_names,
                     obj=np.concatenate([np.asarray(b.c_master, float) for b in p.blocks]), obj_const=0.0,
                     a_rows=m_rows, a_cols=m_cols, a_vals=m_vals, row_lb=np.asarray(p.link_lb, float),
                     row_ub=np.asarray(p.link_ub, float), col_lb=np.zeros(int(off[-1])), col_ub=np.full(int(off[-1]), INF),
                     maximize=False)
    mpath = os.path.join(directory, "master.lp")
    write_lp(master, mpath, obj_all=True)
    gpath = os.path.join(directory, "guidefile")
    with open(gpath, "w") as fh:
        fh.write(f"{p.K}\n")
        for s_ in subs:
            fh.write(os.path.basename(s_) + "\n")
        fh.write("master.lp\nmonolithic.lp\n" + repr(float(p.shift)) + "\n")
    return {"master": mpath, "subs": subs, "guidefile": gpath}


# --------------------------------------------------------------------------------------
# npz (de)serialisation — the committed fixtures under tests/golden/
# --------------------------------------------------------------------------------------
_BLOCK_FIELDS = ("a_rowptr", "a_col", "a_val", "row_lb", "row_ub", "col_lb", "col_ub",
                 "c_file", "c_master", "d_row", "d_col", "d_val")


def save_npz(p: BlockAngularLP, path: str) -> None:
    d = {"K": np.int64(p.K), "L": np.int64(p.L), "shift": np.float64(p.shift),
         "link_lb": p.link_lb, "link_ub": p.link_ub}
    for k, b in enumerate(p.blocks):
        d[f"b{k}_mn"] = np.asarray([b.m, b.n], np.int64)
        for f in _BLOCK_FIELDS:
            d[f"b{k}_{f}"] = getattr(b, f)
        if b.col_names is not None:
            d[f"b{k}_names"] = np.asarray(b.col_names)
    np.savez_compressed(path, **d)


def load_npz(path: str) -> BlockAngularLP:
    z = np.load(path, allow_pickle=False)
    K, L = int(z["K"]), int(z["L"])
    blocks = []
    for k in range(K):
        m, n = (int(v) for v in z[f"b{k}_mn"])
        kw = {f: z[f"b{k}_{f}"] for f in _BLOCK_FIELDS}
        names = [str(s) for s in z[f"b{k}_names"]] if f"b{k}_names" in z.files else None
        blocks.append(Block(m=m, n=n, col_names=names, **kw))
    return BlockAngularLP(K=K, L=L, blocks=blocks, link_lb=z["link_lb"], link_ub=z["link_ub"],
                          shift=float(z["shift"]), name=os.path.splitext(os.path.basename(path))[0])


# --------------------------------------------------------------------------------------
# Synthetic generators
# --------------------------------------------------------------------------------------
def gen_mcf(K: int, nodes: int, cols: int, L: int, seed: int, name: str = "mcf",
            block_range: Optional[range] = None) -> BlockAngularLP:
    """Multicommodity flow: one shared digraph (ring + random chords, ``cols-1`` arcs) plus one
    per-commodity overflow arc; block k = commodity k with flow-conservation rows (FX) and
    per-commodity arc capacities; linking rows = joint capacity on the first L (ring) arcs,
    set to 0.6 x the load of independent shortest-path routing so that they bind.
    All data are small integers (linking capacities are multiples of 0.5), so every basic
    solution is exactly representable.  ``block_range`` builds only a shard of the blocks (the
    instance itself — graph, costs, capacities, linking rhs — never depends on the shard)."""
    rng = np.random.default_rng(seed)
    n_arcs = cols - 1
    if n_arcs < nodes or L > nodes:
        raise ValueError("need cols-1 >= nodes and L <= nodes")
    tail = np.empty(n_arcs, np.int64)
    head = np.empty(n_arcs, np.int64)
    tail[:nodes] = np.arange(nodes)
    head[:nodes] = (np.arange(nodes) + 1) % nodes
    for a in range(nodes, n_arcs):
        while True:
            u, v = rng.integers(0, nodes, 2)
            if u != v and (v - u) % nodes != 1:
                break
        tail[a], head[a] = u, v
    src = rng.integers(0, nodes, K)
    dst = (src + 1 + rng.integers(0, nodes - 1, K)) % nodes
    dem = rng.integers(1, 11, K).astype(np.float64)
    cap = rng.integers(5, 16, (K, n_arcs)).astype(np.float64)
    cost = rng.integers(1, 21, (K, n_arcs)).astype(np.float64)

    # unconstrained load: per-commodity shortest path by cost (Bellman-Ford, all commodities at once)
    dist = np.full((K, nodes), np.inf)
    dist[np.arange(K), src] = 0.0
    for _ in range(nodes):
        nd = dist[:, tail] + cost                      # K x arcs
        best = np.full((K, nodes), np.inf)
        np.minimum.at(best, (np.arange(K)[:, None], head[None, :]), nd)
        if not (best < dist).any():
            break
        dist = np.minimum(dist, best)
    tight = (dist[:, tail] + cost) == dist[:, head]    # K x arcs; costs are integers: exact
    load = np.zeros(n_arcs)
    cur = dst.copy()
    alive = cur != src
    ks = np.arange(K)
    for _ in range(nodes):
        if not alive.any():
            break
        cand = tight & (head[None, :] == cur[:, None]) & alive[:, None]
        arc = np.argmax(cand, axis=1)                  # first tight incoming arc
        ok = cand[ks, arc]
        np.add.at(load, arc[ok], dem[ok])
        cur = np.where(ok, tail[arc], cur)
        alive = ok & (cur != src)
    link_ub = np.maximum(1.0, np.floor(0.6 * load[:L] * 2.0) / 2.0)

    # shared incidence CSR (+1 leaves, -1 enters), columns sorted inside each row
    rows_b = np.concatenate([tail, head])
    cols_b = np.concatenate([np.arange(n_arcs), np.arange(n_arcs)])
    vals_b = np.concatenate([np.ones(n_arcs), -np.ones(n_arcs)])
    order = np.lexsort((cols_b, rows_b))
    rows_b, cols_b, vals_b = rows_b[order], cols_b[order].astype(np.int32), vals_b[order]
    rp_b = np.zeros(nodes + 1, np.int64)
    np.add.at(rp_b, rows_b + 1, 1)
    rp_b = np.cumsum(rp_b)
    d_row = np.arange(L, dtype=np.int32)
    d_val = np.ones(L)
    blocks = []
    vidx = np.arange(nodes + 1)
    for k in (range(K) if block_range is None else block_range):
        s, t = int(src[k]), int(dst[k])
        # the overflow arc (last column) is appended to the rows of its two end nodes
        ins = np.array([rp_b[s + 1], rp_b[t + 1]]) if s < t else np.array([rp_b[t + 1], rp_b[s + 1]])
        insv = np.array([1.0, -1.0]) if s < t else np.array([-1.0, 1.0])
        ci = np.insert(cols_b, ins, n_arcs).astype(np.int32)
        cv = np.insert(vals_b, ins, insv)
        rowptr = (rp_b + (vidx > s) + (vidx > t)).astype(np.int32)
        b = np.zeros(nodes)
        b[s] = dem[k]
        b[t] = -dem[k]
        c = np.concatenate([cost[k], [1000.0]])
        ub = np.concatenate([cap[k], [dem[k]]])
        blocks.append(Block(
            m=nodes, n=cols, a_rowptr=rowptr, a_col=ci, a_val=cv, row_lb=b, row_ub=b.copy(),
            col_lb=np.zeros(cols), col_ub=ub, c_file=c, c_master=c.copy(),
            d_row=d_row, d_col=d_row, d_val=d_val))
    return BlockAngularLP(K=len(blocks), L=L, blocks=blocks, link_lb=np.full(L, -INF), link_ub=link_ub,
                          name=name, meta={"kind": "mcf", "seed": seed, "nodes": nodes, "cols": cols, "K_total": K})


def gen_dense(K: int, m: int, n: int, L: int, seed: int, dens_a: float = 0.02, dens_d: float = 0.25,
              name: str = "dense") -> BlockAngularLP:
    """Random feasible/bounded LP blocks with dense-ish coupling (config C in SURVEY.md §8d)."""
    rng = np.random.default_rng(seed)
    blocks = []
    link_act = np.zeros(L)
    for _ in range(K):
        x0 = rng.random(n)
        nnz_row = max(1, int(round(dens_a * n)))
        a_col = np.empty((m, nnz_row), np.int32)
        for i in range(m):
            a_col[i] = np.sort(rng.choice(n, nnz_row, replace=False))
        a_val = rng.standard_normal((m, nnz_row))
        rowptr = (np.arange(m + 1) * nnz_row).astype(np.int32)
        act = (a_val * x0[a_col]).sum(axis=1)
        row_ub = act + rng.uniform(0.1, 1.0, m)
        nd = max(1, int(round(dens_d * L)))
        d_row = np.empty((n, nd), np.int32)
        for j in range(n):
            d_row[j] = np.sort(rng.choice(L, nd, replace=False))
        d_val = rng.random((n, nd))
        d_col = np.repeat(np.arange(n, dtype=np.int32), nd)
        np.add.at(link_act, d_row.ravel(), (d_val * x0[:, None]).ravel())
        c = rng.standard_normal(n)
        blocks.append(Block(
            m=m, n=n, a_rowptr=rowptr, a_col=a_col.ravel(), a_val=a_val.ravel(),
            row_lb=np.full(m, -INF), row_ub=row_ub, col_lb=np.zeros(n), col_ub=np.ones(n),
            c_file=c.copy(), c_master=c.copy(), d_row=d_row.ravel(), d_col=d_col, d_val=d_val.ravel()))
    return BlockAngularLP(K=K, L=L, blocks=blocks, link_lb=np.full(L, -INF), link_ub=link_act + 1.0,
                          name=name, meta={"kind": "dense", "seed": seed})


#: the named synthetic configurations of BASELINE.json (shapes) with SURVEY.md §8d seeds
CONFIGS = {
    "A": dict(kind="mcf", K=1024, m=64, n=128, L=32, seed=1024001),
    "B": dict(kind="mcf", K=8192, m=256, n=512, L=256, seed=8192001),
    "C": dict(kind="dense", K=64, m=2048, n=4096, L=256, seed=64001),
}


def make_config(tag: str, K: Optional[int] = None, block_range: Optional[range] = None) -> BlockAngularLP:
    """The named synthetic instance; ``K`` overrides the block count (a different, smaller
    instance), ``block_range`` builds only that shard of the SAME instance."""
    cfg = dict(CONFIGS[tag])
    if K is not None:
        cfg["K"] = K
    if cfg["kind"] == "mcf":
        return gen_mcf(cfg["K"], cfg["m"], cfg["n"], cfg["L"], cfg["seed"], name=f"mcf{tag}", block_range=block_range)
    p = gen_dense(cfg["K"], cfg["m"], cfg["n"], cfg["L"], cfg["seed"], name=f"dense{tag}")
    if block_range is not None:
        p.blocks = [p.blocks[k] for k in block_range]
        p.K = len(p.blocks)
    return p


def shard_range(K: int, world: int, rank: int) -> range:
    """Static contiguous block range of a rank: rank g of G owns blocks [g K / G, (g + 1) K / G).
    K must divide evenly (it does for every BASELINE.json configuration at 1, 2, 4 and 8 GPUs)."""
    if K % world:
        raise ValueError(f"{K} blocks do not divide over {world} ranks")
    per = K // world
    return range(rank * per, (rank + 1) * per)
