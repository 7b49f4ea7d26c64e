#!/usr/bin/env python
"""Generate tests/golden/rhs_golden.npz from the REFERENCE'S OWN SOURCE TEXT.

Build-container only (reads /root/reference).  The body of gyncycle_rhs!
(src/gync/gyncycle.jl:50-244) is transliterated MECHANICALLY at run time (regex: `^`->`**`,
1-based `y[..]`/`p[..]`/`f[..]` through index-shifting proxies, hplus/hminus from :41-46)
and executed on seeded random inputs; no reference source is written into the repo.
The outputs pin oracle/gync_oracle.py::gyncycle_rhs, oracle/gync_oracle.c and the CUDA RHS.
"""
import os, re
import numpy as np

SRC = "/root/reference/src/gync/gyncycle.jl"
HERE = os.path.dirname(os.path.abspath(__file__))


class One:                       # 1-based view
    def __init__(s, a): s.a = a
    def __getitem__(s, i): return s.a[i - 1]
    def __setitem__(s, i, v): s.a[i - 1] = v


def load_rhs():
    txt = open(SRC).read()
    body = txt[txt.index("function gyncycle_rhs!"):]
    body = body[:body.index("return f")]
    lines = []
    for ln in body.splitlines()[1:]:
        ln = ln.split("#")[0].rstrip()
        if not ln.strip() or ln.strip() in ("@inbounds begin", "end"):
            continue
        lines.append(ln.strip())
    # join continuation lines (a statement continues while the previous line ends with an operator)
    stmts, cur = [], ""
    for ln in lines:
        cur = (cur + " " + ln).strip()
        if cur[-1] in "+-*/(":
            continue
        stmts.append(cur)
        cur = ""
    assert not cur
    code = "\n".join(s.replace("^", "**") for s in stmts)
    hp = lambda s, t, n: ((s / t) ** n) / (1 + (s / t) ** n)      # gyncycle.jl:41-44
    hm = lambda s, t, n: 1 / (1 + (s / t) ** n)                   # gyncycle.jl:46
    compiled = compile(code, "gyncycle_rhs", "exec")

    def rhs(y, p):
        f = np.zeros(33)
        exec(compiled, {"hplus": hp, "hminus": hm}, {"y": One(y), "p": One(p), "f": One(f)})
        return f
    return rhs, len(stmts)


if __name__ == "__main__":
    import json
    rhs, n = load_rhs()
    C = json.load(open(os.path.join(HERE, "..", "..", "gync.jl_b200", "data", "reference_constants.json")))
    y0, p0 = np.array(C["refy0"]), np.array(C["refallparms"])
    rng = np.random.Generator(np.random.PCG64(20160911))
    Y, P = [y0], [p0]
    for k in range(63):
        s = 0.05 if k < 32 else 0.7
        Y.append(y0 * np.exp(s * rng.standard_normal(33)))
        P.append(p0 * np.exp(s * rng.standard_normal(114)))
    Y, P = np.array(Y), np.array(P)
    F = np.array([rhs(y, p) for y, p in zip(Y, P)])
    assert np.all(np.isfinite(F))
    np.savez(os.path.join(HERE, "rhs_golden.npz"), Y=Y, P=P, F=F)
    print("statements:", n, "f(refy0)[[0,5,1,6]] =", F[0, [0, 5, 1, 6]])
