"""Symbolic definition of the GynCycle right-hand side for code generation.

This is the PRODUCT's statement of the model (the oracle has its own, independent one in
oracle/); tests compare the two and both against tests/golden/rhs_golden.npz.

Model source: reference src/gync/gyncycle.jl:50-244 (equations), :41-46 (Hill functions),
:60-92 (species order).  Differences in *form* (not in mathematics), chosen for the FP64 pipe:
  * every threshold / divisor parameter enters as a precomputed reciprocal IPk = 1/p[k]
    (computed once per sample), so a Hill term costs one division instead of two;
  * Hill terms are built from a shared power chain u^(n-1), u^n and one reciprocal
    r = 1/(1+u^n); their derivatives reuse the chain (no division by the state);
  * the real power (LH_R/p42)^3.689 is r^3 * r^0.689 so that its derivative
    3.689 r^2 r^0.689 / p42 needs no second pow and no division by the state.
"""
from __future__ import annotations

from collections import OrderedDict

import sympy as sp

NY, NP = 33, 114

# species (gyncycle.jl:60-92), 0-based
(LH_pit, LH_bl, R_LH, LH_R, R_LH_des, FSH_pit, FSH_bl, R_FSH, FSH_R, R_FSH_des, SENS, AF1, AF2, AF3, AF4,
 PrF, OvF, Sc1, Sc2, Lut1, Lut2, Lut3, Lut4, E2, P4, IhA, IhB, IhA_e, G, R_G_a, R_G_i, G_R_a, G_R_i) = range(33)


class Rcp(sp.Function):
    """Rcp(a) = 1/a, kept opaque so that d/dx Rcp(a) = -Rcp(a)^2 a' reuses the reciprocal."""
    nargs = 1

    def fdiff(self, argindex=1):
        return -Rcp(self.args[0]) ** 2


class RPow(sp.Function):
    """RPow(r) = r^3.689 (gyncycle.jl:144).  d/dr = 3.689 r^2 RFrac(r), RFrac(r) = r^0.689."""
    nargs = 1

    def fdiff(self, argindex=1):
        r = self.args[0]
        return sp.Float("3.689") * r ** 2 * RFrac(r)


class RFrac(sp.Function):
    nargs = 1


class Model:
    """Builds f (33 expressions) and the sparse Jacobian as expressions over
    y0..y32, parameter symbols P<k>/IP<k> (1-based k as in the reference) and
    ordered intermediates."""

    def __init__(self):
        self.y = sp.symbols("y0:33", real=True)
        self.psyms: "OrderedDict[str, sp.Symbol]" = OrderedDict()
        self.defs: "OrderedDict[sp.Symbol, sp.Expr]" = OrderedDict()
        self._dmemo = {}
        self._n = 0
        self.f = [None] * NY
        self._build()

    # ---- symbols
    def P(self, k):
        return self.psyms.setdefault(f"P{k}", sp.Symbol(f"P{k}", real=True))

    def IP(self, k):
        return self.psyms.setdefault(f"IP{k}", sp.Symbol(f"IP{k}", real=True))

    def let(self, expr, hint="t"):
        """Name an intermediate (evaluated once, in order)."""
        s = sp.Symbol(f"{hint}_{self._n}", real=True)
        self._n += 1
        self.defs[s] = expr
        return s

    # ---- Hill terms: returns (hplus, hminus)
    def hill(self, s, ipk, n, hint="h"):
        u = self.let(s * ipk, hint + "u") if not isinstance(s * ipk, sp.Symbol) else s * ipk
        um = self.let(u ** (n - 1), hint + "m") if n > 2 else (u if n == 2 else sp.Integer(1))
        q = self.let(um * u, hint + "q") if n > 1 else u
        r = self.let(Rcp(1 + q), hint + "r")
        return self.let(q * r, hint + "p"), r

    # ---- the equations
    def _build(self):
        y, P, IP, let, hill = self.y, self.P, self.IP, self.let, self.hill
        f = self.f
        # eq 29a/b (gyncycle.jl:99-100)
        hpE2_97, _ = hill(y[E2], IP(97), 10, "a")
        _, hmP4_94 = hill(y[P4], IP(94), 2, "b")
        y_freq = let(P(93) * hmP4_94 * (1 + P(96) * hpE2_97), "yfreq")
        hpE2_100, _ = hill(y[E2], IP(100), 2, "c")
        _, hmE2_102 = hill(y[E2], IP(102), 1, "d")
        y_mass = let(P(99) * (hpE2_100 + hmE2_102), "ymass")
        # eq 1, 2 (:103-108)
        hpE2_3, _ = hill(y[E2], IP(3), 10, "e")
        _, hmP4_5 = hill(y[P4], IP(5), 1, "g")
        syn_lh = let((P(1) + P(2) * hpE2_3) * hmP4_5, "synlh")
        hpG_9, _ = hill(y[G_R_a], IP(9), 5, "i")
        rel_lh = let(P(7) + P(8) * hpG_9, "rellh")
        f[LH_pit] = syn_lh - rel_lh * y[LH_pit]
        f[LH_bl] = rel_lh * y[LH_pit] * IP(11) - (P(12) * y[R_LH] + P(13)) * y[LH_bl]
        # eq 3-5 (:111-117)
        f[R_LH] = P(14) * y[R_LH_des] - P(12) * y[LH_bl] * y[R_LH]
        f[LH_R] = P(12) * y[LH_bl] * y[R_LH] - P(15) * y[LH_R]
        f[R_LH_des] = P(15) * y[LH_R] - P(14) * y[R_LH_des]
        # eq 6, 7 (:120-125)
        ua = let(y[IhA_e] * IP(17), "ua")
        ub = let(y[IhB] * IP(19), "ub")
        ua4 = let(ua ** 4, "ua4")
        ub2 = let(ub ** 2, "ub2")
        rinh = let(Rcp(1 + ua4 * ua + ub2 * ub), "rinh")
        _, hmF_21 = hill(y_freq, IP(21), 3, "j")
        syn_fsh = let(P(16) * rinh * hmF_21, "synfsh")
        hpG_25, _ = hill(y[G_R_a], IP(25), 3, "k")
        rel_fsh = let(P(23) + P(24) * hpG_25, "relfsh")
        f[FSH_pit] = syn_fsh - rel_fsh * y[FSH_pit]
        f[FSH_bl] = rel_fsh * y[FSH_pit] * IP(11) - (P(27) * y[R_FSH] + P(28)) * y[FSH_bl]
        # eq 8-10 (:128-134)
        f[R_FSH] = P(29) * y[R_FSH_des] - P(27) * y[FSH_bl] * y[R_FSH]
        f[FSH_R] = P(27) * y[FSH_bl] * y[R_FSH] - P(30) * y[FSH_R]
        f[R_FSH_des] = P(30) * y[FSH_R] - P(29) * y[R_FSH_des]
        # eq 11, 12 (:137-140)
        hpF_32, _ = hill(y[FSH_bl], IP(32), 5, "l")
        hpP4_35, _ = hill(y[P4], IP(35), 5, "m")
        f[SENS] = P(31) * hpF_32 - P(34) * hpP4_35 * y[SENS]
        hpFR_38, _ = hill(y[FSH_R], IP(38), 3, "n")
        f[AF1] = P(37) * hpFR_38 - P(40) * y[FSH_R] * y[AF1]
        # eq 13-17 (:144-167)
        r = let(y[LH_R] * IP(42), "rr")
        rpow = let(RPow(r), "rpow")
        r2 = let(r ** 2, "rr2")
        r5 = let(r2 * r2 * r, "rr5")
        r6 = let(r5 * r, "rr6")
        t13 = P(41) * rpow * y[SENS] * y[AF2]
        t14 = P(46) * r5 * y[SENS] * y[AF3]
        t15 = P(50) * r * y[SENS] * y[AF4]
        t16 = P(51) * r6 * y[SENS] * y[PrF]
        f[AF2] = P(40) * y[FSH_R] * y[AF1] - t13
        f[AF3] = t13 + P(44) * y[FSH_R] * y[AF3] * (1 - y[AF3] * IP(45)) - t14
        f[AF4] = t14 + P(48) * r2 * y[AF4] * (1 - y[AF4] * IP(45)) - t15
        f[PrF] = t15 - t16
        hpPrF_54, _ = hill(y[PrF], IP(54), 10, "o")
        f[OvF] = P(53) * r6 * y[SENS] * hpPrF_54 - P(56) * y[OvF]
        # eq 18-23 (:170-185)
        hpOvF_58, _ = hill(y[OvF], IP(58), 10, "q")
        f[Sc1] = P(57) * hpOvF_58 - P(60) * y[Sc1]
        f[Sc2] = P(60) * y[Sc1] - P(61) * y[Sc2]
        hpG_64, _ = hill(y[G_R_a], IP(64), 5, "s")
        lut = let(1 + P(63) * hpG_64, "lut")
        f[Lut1] = P(61) * y[Sc2] - P(62) * lut * y[Lut1]
        f[Lut2] = P(62) * y[Lut1] - P(66) * lut * y[Lut2]
        f[Lut3] = P(66) * y[Lut2] - P(67) * lut * y[Lut3]
        f[Lut4] = P(67) * y[Lut3] - P(68) * lut * y[Lut4]
        # eq 24-28 (:188-215)
        f[E2] = (P(69) + P(70) * y[AF2] + P(71) * y[LH_bl] * y[AF3] + P(72) * y[AF4]
                 + P(73) * y[LH_bl] * y[PrF] + P(74) * y[Lut1] + P(75) * y[Lut4] - P(76) * y[E2])
        f[P4] = P(77) + P(78) * y[Lut4] - P(79) * y[P4]
        f[IhA] = (P(80) + P(81) * y[PrF] + P(82) * y[Sc1] + P(83) * y[Lut1] + P(84) * y[Lut2]
                  + P(85) * y[Lut3] + P(86) * y[Lut4] - P(87) * y[IhA])
        f[IhB] = P(88) + P(89) * y[AF2] + P(90) * y[Sc2] - P(91) * y[IhB]
        f[IhA_e] = P(87) * y[IhA] - P(92) * y[IhA_e]
        # eq 29-33 (:218-240)
        bind = P(104) * y[G] * y[R_G_a]
        f[G] = y_mass * y_freq - bind + P(105) * y[G_R_a] - P(106) * y[G]
        f[R_G_a] = P(105) * y[G_R_a] - bind - P(107) * y[R_G_a] + P(108) * y[R_G_i]
        f[R_G_i] = P(109) + P(107) * y[R_G_a] - P(108) * y[R_G_i] + P(110) * y[G_R_i] - P(111) * y[R_G_i]
        f[G_R_a] = bind - P(105) * y[G_R_a] - P(112) * y[G_R_a] + P(113) * y[G_R_i]
        f[G_R_i] = P(112) * y[G_R_a] - P(113) * y[G_R_i] - P(110) * y[G_R_i] - P(114) * y[G_R_i]

    # ---- total derivatives through the intermediates (chain rule, memoised)
    def total_diff(self, expr, yj):
        res = sp.diff(expr, yj)
        for s in sorted(expr.free_symbols & set(self.defs), key=lambda v: v.name):
            ds = self._dsym(s, yj)
            if ds != 0:
                res += sp.diff(expr, s) * ds
        return res

    def _dsym(self, s, yj):
        key = (s, yj)
        if key in self._dmemo:
            return self._dmemo[key]
        d = self.total_diff(self.defs[s], yj)
        # fold opaque functions of already-named arguments back onto their symbols
        d = self._fold(d)
        if d != 0:
            if not isinstance(d, sp.Symbol):
                ds = sp.Symbol(f"d{s.name}_{yj.name}", real=True)
                self.ddefs[ds] = d
                d = ds
        self._dmemo[key] = d
        return d

    def _fold(self, e):
        rep = {v: k for k, v in self.defs.items() if isinstance(v, (Rcp, RPow))}
        return e.xreplace(rep)

    def jacobian(self):
        """Returns ({(i,j): expr}, ddefs) — nonzero entries df_i/dy_j and the derivative intermediates
        (ordered so that each one only uses earlier ones)."""
        self.ddefs: "OrderedDict[sp.Symbol, sp.Expr]" = OrderedDict()
        J = {}
        for i in range(NY):
            for j in range(NY):
                d = self._fold(self.total_diff(self.f[i], self.y[j]))
                if d != 0:
                    J[(i, j)] = d
        return J, self.ddefs
