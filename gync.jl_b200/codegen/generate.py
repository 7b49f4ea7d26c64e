#!/usr/bin/env python
"""Code generator: symbolic GynCycle model -> straight-line FP64 code for sm_100a.

Emits gync.jl_b200/csrc/gen/gync_model_gen.h containing (all `GYNC_HD static`,
usable from CUDA device code and, for the CPU-side unit tests, from plain C++):

  gync_pack_params(p114, q)        derived per-sample parameter vector q[GYNC_NQ]
                                   (used parameters and reciprocals of divisor parameters)
  gync_rhs(y, q, f)                f = dy/dt                       (gyncycle.jl:50-244)
  gync_rhs_w(y, q, fac, f, A)      f and W = fac*I - J written straight into the sparse LU
                                   storage A[GYNC_NLU] (fill-in positions zeroed)
  gync_lu(A)                       in-place sparse LU, static Markowitz order, diagonal pivots,
                                   inverse pivots stored on the diagonal (KPP-style)
  gync_lusolve(A, b)               b <- W^{-1} b
  tables: GYNC_LU_ROW/COL (pattern), op counts as macros (for the roofline bookkeeping)

Run:  python gync.jl_b200/codegen/generate.py      (the output is committed)
"""
from __future__ import annotations

import os
import re
import sys
from collections import OrderedDict

import sympy as sp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from model import NY, Model, Rcp, RFrac, RPow  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.environ.get("GYNC_GEN_OUT") or os.path.join(HERE, "..", "csrc", "gen", "gync_model_gen.h")


# ----------------------------------------------------------------------------- printing
class Counter:
    def __init__(self):
        self.add = self.mul = self.rcp = self.pow = 0

    def flops(self):
        return self.add + self.mul

    def __repr__(self):
        return f"add={self.add} mul={self.mul} rcp={self.rcp} pow={self.pow}"


def cprint(e, cnt: Counter) -> str:
    """Print a sympy expression as C, counting FP64 operations (a*b+c counts as mul+add;
    the compiler contracts to DFMA)."""
    if isinstance(e, sp.Symbol):
        return NAMES.get(e, e.name)
    if isinstance(e, sp.Integer):
        return f"{int(e)}.0"
    if isinstance(e, sp.Rational):
        return f"({int(e.p)}.0/{int(e.q)}.0)"
    if isinstance(e, sp.Float):
        return repr(float(e))
    if isinstance(e, Rcp):
        cnt.rcp += 1
        return f"GYNC_RCP({cprint(e.args[0], cnt)})"
    if isinstance(e, RFrac):
        cnt.pow += 1
        return f"GYNC_POW0689({cprint(e.args[0], cnt)})"
    if isinstance(e, RPow):
        raise AssertionError("RPow must be lowered before printing")
    if isinstance(e, sp.Pow):
        b, n = e.args
        assert isinstance(n, sp.Integer) and int(n) >= 2, e
        n = int(n)
        cnt.mul += _POWMULS[n]
        return f"gync_ipow{n}({cprint(b, cnt)})"
    if isinstance(e, sp.Mul):
        c, rest = e.as_coeff_Mul()
        terms = []
        sign = ""
        if c == -1:
            sign = "-"
        elif c != 1:
            terms.append(cprint(c, cnt))
        for a in (rest.args if isinstance(rest, sp.Mul) else [rest]):
            s = cprint(a, cnt)
            terms.append(f"({s})" if isinstance(a, sp.Add) else s)
        cnt.mul += len(terms) - 1
        return sign + "*".join(terms)
    if isinstance(e, sp.Add):
        out = ""
        args = sorted(e.args, key=lambda a: a.could_extract_minus_sign())
        for k, a in enumerate(args):
            if a.could_extract_minus_sign() and k > 0:
                out += " - " + _wrap(-a, cnt)
            else:
                out += (" + " if k else "") + _wrap(a, cnt)
        cnt.add += len(args) - 1
        return out
    raise NotImplementedError(type(e))


def _wrap(a, cnt):
    s = cprint(a, cnt)
    return f"({s})" if isinstance(a, sp.Add) else s


_POWMULS = {2: 1, 3: 2, 4: 2, 5: 3, 6: 3, 8: 3, 9: 4, 10: 4}
_IPOW_SRC = """
GYNC_HD static double gync_ipow2(double x){return x*x;}
GYNC_HD static double gync_ipow3(double x){return x*x*x;}
GYNC_HD static double gync_ipow4(double x){double a=x*x;return a*a;}
GYNC_HD static double gync_ipow5(double x){double a=x*x;return a*a*x;}
GYNC_HD static double gync_ipow6(double x){double a=x*x;return a*a*a;}
GYNC_HD static double gync_ipow8(double x){double a=x*x;a=a*a;return a*a;}
GYNC_HD static double gync_ipow9(double x){double a=x*x;a=a*a;return a*a*x;}
GYNC_HD static double gync_ipow10(double x){double a=x*x;double b=a*a;return b*b*a;}
"""

NAMES: dict = {}


# ----------------------------------------------------------------------------- scheduling
def lower_rpow(e):
    """RPow(r) -> r^3 * RFrac(r)  (one shared pow for value and derivative)."""
    return e.replace(lambda x: isinstance(x, RPow), lambda x: x.args[0] ** 3 * RFrac(x.args[0]))


def schedule(assigns, outputs, tmp_prefix):
    """assigns: ordered [(Symbol, expr)] intermediates; outputs: [(lhs_str, expr)].
    Runs a joint CSE and returns topologically ordered [(lhs_str, expr, is_decl)]."""
    exprs = [lower_rpow(e) for _, e in assigns] + [lower_rpow(e) for _, e in outputs]
    syms = sp.numbered_symbols(tmp_prefix, real=True)
    repl, red = sp.cse(exprs, symbols=syms, order="none")
    todo = [(s.name, e, True) for s, e in repl]
    todo += [(s.name, e, True) for (s, _), e in zip(assigns, red[:len(assigns)])]
    todo += [(l, e, False) for (l, _), e in zip(outputs, red[len(assigns):])]
    defined = set()
    lhs_syms = {t[0] for t in todo if t[2]}
    order = []
    pending = todo
    while pending:
        nxt = []
        progressed = False
        for l, e, d in pending:
            need = {s.name for s in e.free_symbols if s.name in lhs_syms}
            if need <= defined:
                order.append((l, e, d))
                if d:
                    defined.add(l)
                progressed = True
            else:
                nxt.append((l, e, d))
        assert progressed, "cyclic dependency"
        pending = nxt
    # drop intermediates nobody uses
    used = set()
    keep = []
    for l, e, d in reversed(order):
        if d and l not in used:
            continue
        keep.append((l, e, d))
        used |= {s.name for s in e.free_symbols}
    return list(reversed(keep))


Q_LOOKAHEAD = int(os.environ.get("GYNC_GEN_Q_LOOKAHEAD", "6"))     # statements of lookahead for parameter loads (software prefetch distance)


def emit(stmts, cnt):
    """Print the scheduled statements.  Every parameter q[k] is loaded ONCE into a named local, and
    the loads are emitted a bounded number of statements ahead of their first use: the arrays may
    live in shared memory behind volatile reads, which the compiler keeps glued to their use
    (load-use distance 1 => the full shared-memory latency exposed per operation)."""
    qsyms = {sym: name for sym, name in NAMES.items() if name.startswith("q[")}
    local = {sym: sp.Symbol("q" + name[2:-1], real=True) for sym, name in qsyms.items()}
    need = [[sym for sym in e.free_symbols if sym in qsyms] for _, e, _ in stmts]
    loaded = set()
    lines = []
    for i, (l, e, d) in enumerate(stmts):
        for j in range(i, min(i + Q_LOOKAHEAD, len(stmts))):
            for sym in sorted(need[j], key=lambda v: int(qsyms[v][2:-1])):
                if sym not in loaded:
                    loaded.add(sym)
                    lines.append(f"  const double {local[sym].name} = {qsyms[sym]};")
        rhs = cprint(e.xreplace(local), cnt)
        lines.append(f"  {'const double ' if d else ''}{l} = {rhs};")
    return "\n".join(lines)


# ----------------------------------------------------------------------------- sparse LU
def markowitz(pattern, n):
    """Greedy Markowitz ordering with diagonal pivots on a structurally-nonsymmetric pattern.
    Returns (order, filled_pattern)."""
    rows = {i: {j for (a, j) in pattern if a == i} | {i} for i in range(n)}
    cols = {j: {i for (i, b) in pattern if b == j} | {j} for j in range(n)}
    full = {(i, j) for i in range(n) for j in rows[i]}
    remaining = set(range(n))
    order = []
    while remaining:
        best = None
        for k in sorted(remaining):
            r = len(rows[k] & remaining) - 1
            c = len(cols[k] & remaining) - 1
            fill = sum(1 for i in (cols[k] & remaining) - {k} for j in (rows[k] & remaining) - {k}
                       if j not in rows[i])
            key = (fill, r * c, k)
            if best is None or key < best[0]:
                best = (key, k)
        k = best[1]
        order.append(k)
        remaining.discard(k)
        for i in cols[k] & remaining:
            for j in rows[k] & remaining:
                if j not in rows[i]:
                    rows[i].add(j)
                    cols[j].add(i)
                    full.add((i, j))
    return order, full


def main():
    m = Model()
    J, ddefs = m.jacobian()
    y = m.y
    pattern = set(J.keys())
    assert all((i, i) in pattern for i in range(NY))
    order, full = markowitz(pattern, NY)
    rank = {k: r for r, k in enumerate(order)}
    rows = {i: sorted([j for (a, j) in full if a == i], key=lambda j: rank[j]) for i in range(NY)}
    cols = {j: sorted([i for (i, b) in full if b == j], key=lambda i: rank[i]) for j in range(NY)}
    touched = set()
    for k in order:
        for i in [i for i in cols[k] if rank[i] > rank[k]]:
            touched.add((i, k))
            for j in [j for j in rows[k] if rank[j] > rank[k]]:
                touched.add((i, j))
    # storage order: by pivot rank of the row, then pivot rank of the column (row-major in pivot order);
    # entries the factorisation never touches (unmodified U off-diagonals) go LAST (the "tail"), in
    # backward-substitution use order, so that a layout may keep them in a second memory (TMEM on the device)
    tail = [(k, j) for k in reversed(order) for j in rows[k] if rank[j] > rank[k] and (k, j) not in touched]
    tail = tail[:(len(tail) // 8) * 8]
    body = sorted([ij for ij in full if ij not in set(tail)], key=lambda ij: (rank[ij[0]], rank[ij[1]]))
    lu_entries = body + tail
    pos = {ij: n for n, ij in enumerate(lu_entries)}
    nlu = len(lu_entries)
    nbody = len(body)
    assert all(ij in J for ij in tail)
    # parameter slots
    psyms = list(m.psyms.values())
    for k, s in enumerate(psyms):
        NAMES[s] = f"q[{k}]"
    for i in range(NY):
        NAMES[y[i]] = f"y[{i}]"
    nq = len(psyms)

    defs = list(m.defs.items())
    fac = sp.Symbol("fac", real=True)

    # ---- rhs only
    c_rhs = Counter()
    st = schedule(defs, [(f"f[{i}]", m.f[i]) for i in range(NY)], "c")
    src_rhs = emit(st, c_rhs)

    # ---- rhs + W
    c_w = Counter()
    outs = [(f"f[{i}]", m.f[i]) for i in range(NY)]
    for ij in lu_entries:
        i, j = ij
        if ij in J:
            lhs = f"A[{pos[ij]}]" if pos[ij] < nbody else f"const double at{pos[ij] - nbody}"
            outs.append((lhs, (fac if i == j else 0) - J[ij]))
    st = schedule(defs + list(ddefs.items()), outs, "w")
    src_w = emit(st, c_w)
    # each tail chunk is stored as soon as its 8 values exist (keeps the live range of the locals short)
    wl = src_w.split("\n")
    where = {}
    for li, ln in enumerate(wl):
        m_ = re.match(r"\s*const double at(\d+) = ", ln)
        if m_:
            where[int(m_.group(1))] = li
    inserts = sorted(((max(where[8 * c8 + n] for n in range(8)), c8) for c8 in range(len(tail) // 8)), reverse=True)
    for li, c8 in inserts:
        wl.insert(li + 1, "  gync_tail_store8(A, %d, %s);" % (c8, ", ".join(f"at{8 * c8 + n}" for n in range(8))))
    src_w = "\n".join(wl)
    zero_fill = "\n".join(f"  A[{pos[ij]}] = 0.0;" for ij in lu_entries if ij not in J)

    # ---- LU factorisation (right-looking, pivot order `order`)
    c_lu = Counter()
    lu_lines = []
    rows = {i: sorted([j for (a, j) in full if a == i], key=lambda j: rank[j]) for i in range(NY)}
    cols = {j: sorted([i for (i, b) in full if b == j], key=lambda i: rank[i]) for j in range(NY)}
    LU_BATCH = int(os.environ.get("GYNC_GEN_LU_BATCH", "8"))
    lu_chunks = set()
    for k in order:
        right = [j for j in rows[k] if rank[j] > rank[k]]
        below = [i for i in cols[k] if rank[i] > rank[k]]
        if below:      # pivot-row entries kept in the tail are read (never written) here: fetch their chunks first
            for j in right:
                if pos[(k, j)] >= nbody:
                    c8 = (pos[(k, j)] - nbody) // 8
                    if c8 not in lu_chunks:
                        lu_chunks.add(c8)
                        lu_lines.append(f"  double tl{c8}[8]; gync_tail_load8(A, {c8}, tl{c8});")
        lu_lines.append(f"  {{ const double piv = GYNC_RCP(A[{pos[(k, k)]}]); A[{pos[(k, k)]}] = piv;")
        c_lu.rcp += 1
        if below:
            for j in right:
                p_ = pos[(k, j)]
                src_ = f"A[{p_}]" if p_ < nbody else f"tl{(p_ - nbody) // 8}[{(p_ - nbody) % 8}]"
                lu_lines.append(f"    const double u{j} = {src_};")
            for i in below:
                lu_lines.append(f"    const double l{i} = A[{pos[(i, k)]}] * piv;")
                c_lu.mul += 1
            upd = [(i, j) for i in below for j in right]
            for b0 in range(0, len(upd), LU_BATCH):       # loads of a batch first, then the FMAs, then the stores
                batch = upd[b0:b0 + LU_BATCH]
                for (i, j) in batch:
                    lu_lines.append(f"    double t{i}_{j} = A[{pos[(i, j)]}];")
                for (i, j) in batch:
                    lu_lines.append(f"    t{i}_{j} -= l{i} * u{j};")
                    c_lu.mul += 1
                    c_lu.add += 1
                for (i, j) in batch:
                    lu_lines.append(f"    A[{pos[(i, j)]}] = t{i}_{j};")
            for i in below:
                lu_lines.append(f"    A[{pos[(i, k)]}] = l{i};")
        lu_lines.append("  }")
    src_lu = "\n".join(lu_lines)

    # ---- triangular solves
    c_sol = Counter()
    # factors are loaded SOL_BATCH operations ahead of their use.  ptxas keeps roughly the distance it is given, and the
    # distance decides how much of the 29-cycle shared-memory latency each of the 133 loads exposes: measured on B200
    # (evaluations/s, bit-identical results) 8: 207.4 k, 16: 206.1 k, 20: 214.8 k, 24: 218.5 k, 28: 218.6 k, 32: 217.9 k,
    # 48: 187.0 k, 64: 170.1 k (spills) — profiles/r2_ab_solbatch*.log.  LU_BATCH and Q_LOOKAHEAD measured flat.
    SOL_BATCH = int(os.environ.get("GYNC_GEN_SOL_BATCH", "24"))
    ops = []      # (position in A, statement using the local name `a<n>`)
    for k in order:   # forward: L has unit diagonal, multipliers stored below the pivot
        for i in [i for i in cols[k] if rank[i] > rank[k]]:
            ops.append((pos[(i, k)], f"b[{i}] -= {{a}} * b[{k}];"))
            c_sol.mul += 1
            c_sol.add += 1
    for k in reversed(order):
        right = [j for j in rows[k] if rank[j] > rank[k]]
        for j in right:
            ops.append((pos[(k, j)], f"b[{k}] -= {{a}} * b[{j}];"))
        ops.append((pos[(k, k)], f"b[{k}] *= {{a}};"))
        c_sol.mul += len(right) + 1
        c_sol.add += len(right)
    sol_lines = []
    batches = [ops[b0:b0 + SOL_BATCH] for b0 in range(0, len(ops), SOL_BATCH)]
    def loads(bi):
        return [f"  const double a{bi}_{n} = A[{p}];" for n, (p, _) in enumerate(batches[bi]) if p < nbody]
    sol_lines += loads(0)
    chunks_loaded = set()
    for bi, batch in enumerate(batches):
        if bi + 1 < len(batches):
            sol_lines += loads(bi + 1)          # next batch's factors are in flight while this batch computes
        for n, (p, stmt) in enumerate(batch):
            if p < nbody:
                sol_lines.append("  " + stmt.format(a=f"a{bi}_{n}"))
            else:
                c8, off = divmod(p - nbody, 8)
                if c8 not in chunks_loaded:
                    chunks_loaded.add(c8)
                    sol_lines.append(f"  double tl{c8}[8]; gync_tail_load8(A, {c8}, tl{c8});")
                sol_lines.append("  " + stmt.format(a=f"tl{c8}[{off}]"))
    src_sol = "\n".join(sol_lines)

    # ---- pack params
    pack = []
    for k, s in enumerate(psyms):
        idx = int(s.name.lstrip("IP")) - 1
        pack.append(f"  q[{k}] = {'1.0 / ' if s.name.startswith('IP') else ''}p[{idx}];")
    used_p = sorted({int(s.name.lstrip("IP")) for s in psyms})
    import json
    C = json.load(open(os.path.join(HERE, "..", "data", "reference_constants.json")))
    hill = [4, 6, 10, 18, 20, 22, 26, 33, 36, 39, 43, 47, 49, 52, 55, 59, 65, 95, 98, 101, 103]  # model.jl:3
    sampled = [i for i in range(1, 104) if i not in hill]                                           # model.jl:4
    assert len(sampled) == 82 and not (set(hill) & set(used_p)), "Hill-exponent slots must be unused by the RHS"

    hdr = f"""// GENERATED by gync.jl_b200/codegen/generate.py — do not edit.
// GynCycle RHS / Jacobian / sparse LU as straight-line FP64 code (model: reference
// src/gync/gyncycle.jl:50-244; see codegen/model.py for the form chosen).
#pragma once
#ifndef GYNC_HD
#  ifdef __CUDACC__
#    define GYNC_HD __host__ __device__ __forceinline__
#  else
#    define GYNC_HD inline
#  endif
#endif
#ifndef GYNC_RCP
#  define GYNC_RCP(x) (1.0 / (x))
#endif
#ifndef GYNC_POW0689
#  define GYNC_POW0689(x) pow((x), 0.689)
#endif
#define GYNC_NY   {NY}
#define GYNC_NQ   {nq}      /* derived per-sample parameters */
#define GYNC_NNZJ {len(J)}     /* structural nonzeros of the Jacobian */
#define GYNC_NLU  {nlu}     /* nonzeros of L+U after fill (Markowitz order, diagonal pivots) */
#define GYNC_NLU_BODY {nbody}   /* entries [0, BODY) are touched by the factorisation; [BODY, NLU) ("tail") only by gync_rhs_w and gync_lusolve */
/* FP64 operation counts of the generated code (adds, muls, reciprocals, pows) */
#define GYNC_OPS_RHS_FLOP  {c_rhs.flops()}
#define GYNC_OPS_RHS_RCP   {c_rhs.rcp}
#define GYNC_OPS_RHS_POW   {c_rhs.pow}
#define GYNC_OPS_RHSW_FLOP {c_w.flops()}
#define GYNC_OPS_RHSW_RCP  {c_w.rcp}
#define GYNC_OPS_RHSW_POW  {c_w.pow}
#define GYNC_OPS_LU_FLOP   {c_lu.flops()}
#define GYNC_OPS_LU_RCP    {c_lu.rcp}
#define GYNC_OPS_SOLVE_FLOP {c_sol.flops()}
{_IPOW_SRC}
/* pivot order (species indices, 0-based) */
static const int GYNC_PIVOT_ORDER[{NY}] = {{{", ".join(map(str, order))}}};
/* pattern of L+U in storage order */
static const unsigned char GYNC_LU_ROW[{nlu}] = {{{", ".join(str(i) for i, _ in lu_entries)}}};
static const unsigned char GYNC_LU_COL[{nlu}] = {{{", ".join(str(j) for _, j in lu_entries)}}};
/* position of the diagonal entries */
static const unsigned short GYNC_LU_DIAG[{NY}] = {{{", ".join(str(pos[(i, i)]) for i in range(NY))}}};
/* reference data (src/gync/data/refparms.jl, refy0.jl) and index sets (model.jl:1-4,10), as initialiser lists */
#define GYNC_REFALLPARMS_LIST {", ".join(repr(float(v)) for v in C["refallparms"])}
#define GYNC_REFY0_LIST {", ".join(repr(float(v)) for v in C["refy0"])}
#define GYNC_SAMPLEDINDS0_LIST {", ".join(str(i - 1) for i in sampled)}   /* 0-based */
#define GYNC_MEASINDS0_LIST 1, 6, 23, 24                                /* 0-based [2,7,24,25] */
#define GYNC_MEASMAG_LIST 120.0, 10.0, 400.0, 15.0
/* 1-based reference parameter indices that the RHS reads ({len(used_p)} of 114) */
static const unsigned char GYNC_USED_P[{len(used_p)}] = {{{", ".join(map(str, used_p))}}};

/* tail of the LU storage in chunks of 8 (generic layouts keep it in the same array; the TMEM layout overloads these) */
template <class AT>
GYNC_HD static void gync_tail_store8(AT& A, const int c, const double v0, const double v1, const double v2, const double v3,
                                     const double v4, const double v5, const double v6, const double v7) {{
  A[GYNC_NLU_BODY + 8 * c + 0] = v0; A[GYNC_NLU_BODY + 8 * c + 1] = v1; A[GYNC_NLU_BODY + 8 * c + 2] = v2; A[GYNC_NLU_BODY + 8 * c + 3] = v3;
  A[GYNC_NLU_BODY + 8 * c + 4] = v4; A[GYNC_NLU_BODY + 8 * c + 5] = v5; A[GYNC_NLU_BODY + 8 * c + 6] = v6; A[GYNC_NLU_BODY + 8 * c + 7] = v7;
}}
template <class AT>
GYNC_HD static void gync_tail_load8(const AT& A, const int c, double (&v)[8]) {{
  for (int i = 0; i < 8; ++i) v[i] = A[GYNC_NLU_BODY + 8 * c + i];
}}

/* q <- derived parameters from the 114 reference-ordered parameters */
template <class PT, class QT>
GYNC_HD static void gync_pack_params(const PT& p, QT& q) {{
{chr(10).join(pack)}
}}

/* f = dy/dt */
template <class YT, class QT, class FT>
GYNC_HD static void gync_rhs(const YT& y, const QT& q, FT& f) {{
{src_rhs}
}}

/* f = dy/dt and A = fac*I - df/dy in LU storage */
template <class YT, class QT, class FT, class AT>
GYNC_HD static void gync_rhs_w(const YT& y, const QT& q, const double fac, FT& f, AT& A) {{
{src_w}
{zero_fill}
}}

/* in-place LU, inverse pivots on the diagonal */
template <class AT>
GYNC_HD static void gync_lu(AT& A) {{
{src_lu}
}}

/* b <- (LU)^-1 b */
template <class AT, class BT>
GYNC_HD static void gync_lusolve(const AT& A, BT& b) {{
{src_sol}
}}
"""
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as fh:
        fh.write(hdr)
    print(f"nq={nq} nnzJ={len(J)} nLU={nlu} order={order}")
    print("rhs :", c_rhs)
    print("rhsW:", c_w)
    print("lu  :", c_lu)
    print("sol :", c_sol)


if __name__ == "__main__":
    main()
