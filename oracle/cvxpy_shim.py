"""A stand-in for the `cvxpy` module, just large enough to EXECUTE the reference's three scripts unmodified.
TEST INFRASTRUCTURE (oracle/): never imported by the product.

Why: the reference has no library API and no recorded outputs -- its results exist only as what
`prob.solve()` leaves in `prob.value`, `psi.value`, `deltas[i].value`, `lambdas[i].value`
(/root/reference/arbitrage.py:81-84, liquidation.py:84-87, two-asset.py:90-100).  cvxpy and every conic solver are
absent from this image (no wheels, no network), so the scripts cannot run as they are.  With this module installed as
`sys.modules["cvxpy"]`, `tests/golden/make_golden_from_reference.py` runs the reference FILES THEMSELVES
(`runpy.run_path` on /root/reference/*.py: their data literals, their A_i construction, their expression graph, their
constraint list), and only the numerical back end differs: the modelled program is handed to scipy SLSQP with exact
first derivatives instead of to ECOS/Clarabel.  The problem is convex, so any correct solver returns the same optimum;
the fixtures written from these runs pin the oracle and the CUDA path to the reference's own model text.

API subset (exactly what the scripts touch):
    Variable(n, nonneg=True)            arbitrage.py:51-52
    ndarray @ expr, list @ expr, expr +/- expr/array, scalar * expr, expr[i]      arbitrage.py:54,57,60; liquidation.py:57
    sum(list of exprs) | sum(expr) | sum(ndarray)                                 arbitrage.py:54,73
    geo_mean(expr | ndarray, p=None)    arbitrage.py:65,68-70 (weights p normalised to p/sum(p), as cvxpy does)
    expr >= c, expr == c, geo_mean(...) >= c                                      arbitrage.py:63-77; liquidation.py:77-80
    Maximize(expr), Problem(obj, cons).solve(), .value, .status                   arbitrage.py:57,81-84
"""
from __future__ import annotations

import itertools

import numpy as np
from scipy import optimize

__version__ = "shim-0 (scipy SLSQP back end; oracle/cvxpy_shim.py)"

_ids = itertools.count()


class Expression:
    """Affine expression sum_v C_v x_v + b over Variables; shape (r,) or scalar."""
    __array_ufunc__ = None          # numpy defers to our reflected operators (ndarray @ expr, ndarray + expr, ...)
    __hash__ = object.__hash__

    def __init__(self, terms, const, scalar):
        self.terms = terms          # {variable id: (Variable, ndarray (r, n_v))}
        self.const = np.atleast_1d(np.asarray(const, float))
        self.scalar = bool(scalar)

    # ---- structure
    @property
    def size(self):
        return len(self.const)

    @property
    def shape(self):
        return () if self.scalar else (self.size,)

    def _rows(self, r):
        """broadcast a scalar expression to r rows"""
        if self.size == r:
            return self
        if self.size != 1:
            raise ValueError(f"shape mismatch: {self.size} vs {r}")
        return Expression({k: (v, np.repeat(C, r, 0)) for k, (v, C) in self.terms.items()}, np.repeat(self.const, r), False)

    @staticmethod
    def _lift(x):
        if isinstance(x, Expression):
            return x
        a = np.asarray(x, float)
        if a.ndim > 1:
            raise ValueError("only scalars and vectors")
        return Expression({}, a.reshape(-1), a.ndim == 0)

    # ---- arithmetic
    def __add__(self, other):
        o = Expression._lift(other)
        r = max(self.size, o.size)
        a, b = self._rows(r), o._rows(r)
        terms = dict(a.terms)
        for k, (v, C) in b.terms.items():
            terms[k] = (v, terms[k][1] + C) if k in terms else (v, C)
        return Expression(terms, a.const + b.const, self.scalar and o.scalar)

    __radd__ = __add__

    def __neg__(self):
        return Expression({k: (v, -C) for k, (v, C) in self.terms.items()}, -self.const, self.scalar)

    def __sub__(self, other):
        return self + (-Expression._lift(other))

    def __rsub__(self, other):
        return Expression._lift(other) + (-self)

    def __mul__(self, other):
        if isinstance(other, Expression):
            raise TypeError("product of two expressions is not affine")
        a = np.asarray(other, float)
        if a.ndim == 0:
            return Expression({k: (v, float(a) * C) for k, (v, C) in self.terms.items()}, float(a) * self.const, self.scalar)
        if a.ndim != 1:
            raise ValueError("elementwise product needs a vector")
        e = self._rows(len(a))
        return Expression({k: (v, a[:, None] * C) for k, (v, C) in e.terms.items()}, a * e.const, False)

    __rmul__ = __mul__

    def __truediv__(self, other):
        return self * (1.0 / np.asarray(other, float))

    def __rmatmul__(self, left):
        A = np.asarray(left, float)
        if A.ndim == 1:                                     # vector @ expr -> scalar
            if len(A) != self.size:
                raise ValueError("shape mismatch in @")
            return Expression({k: (v, A[None, :] @ C) for k, (v, C) in self.terms.items()}, [float(A @ self.const)], True)
        if A.ndim == 2:
            if A.shape[1] != self.size:
                raise ValueError("shape mismatch in @")
            return Expression({k: (v, A @ C) for k, (v, C) in self.terms.items()}, A @ self.const, False)
        raise ValueError("@ needs a vector or a matrix on the left")

    def __getitem__(self, idx):
        if self.scalar:
            raise IndexError("scalar expression")
        if isinstance(idx, (int, np.integer)):
            return Expression({k: (v, C[idx:idx + 1] if idx != -1 else C[-1:]) for k, (v, C) in self.terms.items()},
                              [self.const[idx]], True)
        return Expression({k: (v, C[idx]) for k, (v, C) in self.terms.items()}, self.const[idx], False)

    # ---- constraints
    def __ge__(self, other):
        return Constraint("ge", self - other)

    def __le__(self, other):
        return Constraint("ge", Expression._lift(other) - self)

    def __eq__(self, other):          # noqa: PLW1641  (hash is identity; term dictionaries are keyed by integer ids)
        return Constraint("eq", self - other)

    # ---- numbers
    def _eval(self, get):
        out = self.const.copy()
        for _, (v, C) in self.terms.items():
            out = out + C @ get(v)
        return out

    @property
    def value(self):
        if any(v._value is None for v, _ in self.terms.values()):
            return None
        out = self._eval(lambda v: v._value)
        return float(out[0]) if self.scalar else out

    def variables(self):
        return [v for v, _ in self.terms.values()]


class Variable(Expression):
    def __init__(self, shape=(), nonneg=False, name=None, **kw):
        if kw:
            raise NotImplementedError(f"Variable attributes {sorted(kw)} are outside the shim's subset")
        n = 1 if shape == () else int(shape if np.isscalar(shape) else shape[0])
        self.id = next(_ids)
        self.n = n
        self.nonneg = bool(nonneg)
        self.name = name or f"var{self.id}"
        self._value = None
        super().__init__({self.id: (self, np.eye(n))}, np.zeros(n), shape == ())


class GeoMean:
    """geo_mean(x, p): prod_j x_j^(p_j / sum p) of an affine x -- concave; only `>= constant` is supported."""
    __array_ufunc__ = None

    def __init__(self, expr, w):
        self.expr, self.w = expr, w

    def __ge__(self, other):
        c = float(np.asarray(other, float))
        if not c > 0:
            raise ValueError("geo_mean(x) >= c needs c > 0")
        return Constraint("geo", self.expr, w=self.w, rhs=c)

    @property
    def value(self):
        x = self.expr.value
        return None if x is None else float(np.exp(np.dot(self.w, np.log(x))))


class Constraint:
    def __init__(self, kind, expr, w=None, rhs=None):
        self.kind, self.expr, self.w, self.rhs = kind, expr, w, rhs


def _weights(n, p):
    w = np.ones(n) if p is None else np.asarray(p, float)
    if len(w) != n or np.any(w <= 0):
        raise ValueError("geo_mean weights must be positive, one per entry")
    return w / w.sum()


def geo_mean(x, p=None):
    if isinstance(x, Expression):
        return GeoMean(x, _weights(x.size, p))
    a = np.asarray(x, float).reshape(-1)
    return float(np.exp(np.dot(_weights(len(a), p), np.log(a))))


_builtin_sum = sum


def sum(x):          # noqa: A001  (cvxpy's own name)
    if isinstance(x, Expression):
        return np.ones(x.size) @ x
    if isinstance(x, (list, tuple)) and any(isinstance(e, Expression) for e in x):
        return _builtin_sum(x[1:], x[0])
    return np.sum(x)


class Maximize:
    sign = -1.0

    def __init__(self, expr):
        e = Expression._lift(expr)
        if e.size != 1:
            raise ValueError("objective must be scalar")
        self.expr = e

    @property
    def value(self):
        return self.expr.value


class Minimize(Maximize):
    sign = 1.0


class SolverError(Exception):
    pass


class Problem:
    def __init__(self, objective, constraints=()):
        self.objective, self.constraints = objective, list(constraints)
        self.value = None
        self.status = None

    def variables(self):
        seen = {}
        for e in [self.objective.expr] + [c.expr for c in self.constraints]:
            for v in e.variables():
                seen[v.id] = v
        return [seen[k] for k in sorted(seen)]

    def solve(self, **kw):
        vs = self.variables()
        off, N = {}, 0
        for v in vs:
            off[v.id] = N
            N += v.n

        def dense(e):
            M = np.zeros((e.size, N))
            for k, (v, C) in e.terms.items():
                M[:, off[k]:off[k] + v.n] += C
            return M, e.const

        c_row, c0 = dense(self.objective.expr)
        sgn = self.objective.sign
        cons = []
        checks = []
        for c in self.constraints:
            M, b = dense(c.expr)
            if c.kind == "geo":
                lr = float(np.log(c.rhs))

                def f(x, M=M, b=b, w=c.w, lr=lr):
                    return np.array([np.dot(w, np.log(np.maximum(M @ x + b, 1e-300))) - lr])

                def jac(x, M=M, b=b, w=c.w):
                    return ((w / np.maximum(M @ x + b, 1e-300)) @ M)[None, :]
                cons.append(dict(type="ineq", fun=f, jac=jac))
                checks.append(("ineq", f))
            else:
                f = (lambda x, M=M, b=b: M @ x + b)
                cons.append(dict(type="ineq" if c.kind == "ge" else "eq", fun=f, jac=(lambda x, M=M: M)))
                checks.append(("ineq" if c.kind == "ge" else "eq", f))
        bounds = []
        for v in vs:
            bounds += [(0.0, None) if v.nonneg else (None, None)] * v.n
        fun = lambda x: sgn * float(c_row[0] @ x + c0[0])
        jac = lambda x: sgn * c_row[0]
        best = None
        rng = np.random.default_rng(0)
        for z0 in [np.zeros(N)] + [0.1 * rng.random(N) for _ in range(3)]:
            res = optimize.minimize(fun, z0, jac=jac, method="SLSQP", bounds=bounds, constraints=cons,
                                    options=dict(maxiter=int(kw.get("max_iters", 3000)), ftol=1e-16))
            viol = 0.0
            for kind, f in checks:
                r = np.atleast_1d(f(res.x))
                viol = max(viol, float(np.max(-r)) if kind == "ineq" else float(np.max(np.abs(r))))
            if viol <= 1e-9 and (best is None or res.fun < best.fun):
                best = res
        if best is None:
            self.status = "infeasible"
            self.value = -np.inf if sgn < 0 else np.inf
            return self.value
        for v in vs:
            v._value = np.maximum(best.x[off[v.id]:off[v.id] + v.n], 0.0) if v.nonneg else best.x[off[v.id]:off[v.id] + v.n].copy()
        self.status = "optimal"
        self.value = float(self.objective.expr.value)
        return self.value
