#!/usr/bin/env python
"""Extract the reference's DATA (not code) into neutral files.

Run in the build container only (needs /root/reference); the outputs are
committed, so nothing at test/bench time reads /root/reference.

  gync.jl_b200/data/reference_constants.json
      refy0 (33), refallparms (114)      <- src/gync/data/refy0.jl, refparms.jl
      species / parameter names          <- src/gync/data/speciesnames.jl, parameternames.jl
  gync.jl_b200/data/patients.npz
      lausanne (45,31,4), pfizer (13,31,4), NaN = missing
      following the loaders' semantics   <- src/gync/data.jl:10-36, 41-58
"""
import json, os, re, sys
import numpy as np

REF = "/root/reference/src/gync/data"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gync.jl_b200", "data")


def numbers(fn):
    return [float(x) for x in re.findall(r"[-+]?[0-9][-+0-9.eE]*", open(os.path.join(REF, fn)).read())]


def strings(fn):
    return re.findall(r'"([^"]*)"', open(os.path.join(REF, fn)).read())


def lausanne():
    # data.jl:10-36 — day d (1..31) is CSV column d+2; numeric cells only; E2 scaled by 1/3.671
    files = ["lh.csv", "fsh.csv", "oestr.csv", "prog.csv"]
    scale = [1.0, 1.0, 1.0 / 3.671, 1.0]
    out = np.full((45, 31, 4), np.nan)
    for h, fn in enumerate(files):
        rows = [ln.rstrip("\r\n").split(",") for ln in open(os.path.join(REF, "lausaunne", fn))]
        for case in range(1, 46):
            row = next((r for r in rows if _num(r[0]) == case), None)
            if row is None:
                continue
            for day in range(1, 32):
                v = _num(row[day + 1]) if day + 1 < len(row) else None
                if v is not None:
                    out[case - 1, day - 1, h] = v * scale[h]
    return out


def _num(s):
    try:
        return float(s)
    except ValueError:
        return None


def pfizer():
    # data.jl:41-58 — readtable(header=true) consumes the first line as a header (file has none);
    # groupby column 6 in order of first appearance; day -> (d+30)%31+1 ; columns 2..5 = LH FSH E2 P4
    lines = [ln.rstrip("\r\n").split("\t") for ln in open(os.path.join(REF, "pfizer_normal.txt"))][1:]
    order, groups = [], {}
    for r in lines:
        if len(r) < 6:
            continue
        sid = r[5]
        if sid not in groups:
            groups[sid] = []
            order.append(sid)
        groups[sid].append(r)
    out = np.full((len(order), 31, 4), np.nan)
    for gi, sid in enumerate(order):
        for r in groups[sid]:
            day = (int(float(r[0])) + 30) % 31 + 1
            for i in range(4):
                v = _num(r[i + 1])
                out[gi, day - 1, i] = np.nan if v is None else v
    return out, order


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    y0, p = numbers("refy0.jl"), numbers("refparms.jl")
    assert len(y0) == 33 and len(p) == 114
    json.dump({"refy0": y0, "refallparms": p,
               "speciesnames": strings("speciesnames.jl"),
               "parameternames": strings("parameternames.jl")},
              open(os.path.join(OUT, "reference_constants.json"), "w"), indent=0)
    la = lausanne()
    pf, order = pfizer()
    np.savez_compressed(os.path.join(OUT, "patients.npz"), lausanne=la, pfizer=pf)
    print("lausanne", la.shape, "non-missing per patient:", (~np.isnan(la)).sum((1, 2))[:10])
    print("pfizer", pf.shape, order)
