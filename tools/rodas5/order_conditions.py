"""Rosenbrock order conditions by rooted trees (Hairer & Wanner II, Thm IV.7.4 in the form with the full
beta = alpha + Gamma INCLUDING the diagonal, which makes every right-hand side 1/gamma(tree)):
   sum_j b_j Phi_j(tree) = 1 / gamma(tree)
   Phi_j(leaf) = 1;  a vertex with ONE son s:  Phi_j = sum_k beta_jk Phi_k(s);  with m>=2 sons: prod_i sum_k alpha_jk Phi_k(s_i).
Works from the implementation form (gamma, A = (a_ij), C = (c_ij)) used by the solver:
   Gamma = (I/gamma - C)^-1,  alpha = A Gamma,  b = m Gamma  with m the weights of the k_i of the implementation form."""
import itertools
import numpy as np


def trees(order):
    """All rooted trees with `order` vertices as nested sorted tuples."""
    if order == 1:
        return [()]
    out = set()
    def parts(n, maxp):
        if n == 0:
            yield ()
            return
        for p in range(min(n, maxp), 0, -1):
            for rest in parts(n - p, p):
                yield (p,) + rest
    for part in parts(order - 1, order - 1):
        for combo in itertools.product(*[trees(p) for p in part]):
            out.add(tuple(sorted(combo)))
    return sorted(out)


def order_of(t):
    return 1 + sum(order_of(s) for s in t)


def density(t):
    g = order_of(t)
    for s in t:
        g *= density(s)
    return g


def phi(t, al, be):
    n = al.shape[0]
    if len(t) == 0:
        return np.ones(n)
    if len(t) == 1:
        return be @ phi(t[0], al, be)
    r = np.ones(n)
    for s in t:
        r = r * (al @ phi(s, al, be))
    return r


def from_impl(g, A, Cm, m):
    s = A.shape[0]
    Gam = np.linalg.inv(np.eye(s) / g - Cm)
    al = A @ Gam
    return al, al + Gam, m @ Gam, Gam


def residuals(g, A, Cm, m, order):
    al, be, b, _ = from_impl(g, A, Cm, m)
    res = []
    for p in range(1, order + 1):
        for t in trees(p):
            res.append((p, t, b @ phi(t, al, be) - 1.0 / density(t)))
    return res
