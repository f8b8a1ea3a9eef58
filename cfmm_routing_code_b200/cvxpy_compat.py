"""The reference's cvxpy call site, served by the B200 path:  ``import cfmm_routing_code_b200.cvxpy_compat as cp``.

The reference scripts have no API of their own: they wire cvxpy objects (arbitrage.py:50-78) and call
``prob.solve()`` (arbitrage.py:81-82, liquidation.py:84-85, two-asset.py:90-91).  This module offers the modelling
subset those scripts touch, so that changing their ``import cvxpy as cp`` line is the whole port:

    Variable(n, nonneg=True)                                                     arbitrage.py:51-52
    ndarray @ expr, expr +/- expr|array, scalar * expr, expr[i]                  arbitrage.py:54,57,60; liquidation.py:57
    sum(list of exprs) | sum(expr) | sum(ndarray)                                arbitrage.py:54,73
    geo_mean(expr | ndarray, p=None)      (p normalised to p / sum p)            arbitrage.py:65,68-70
    expr >= c, expr == c, geo_mean(expr) >= level                                arbitrage.py:63-77; liquidation.py:77-80
    Maximize(expr), Problem(obj, cons).solve(), .value, .status                  arbitrage.py:57,81-84

It is NOT a general convex modelling layer.  ``Problem.solve()`` *recognises* the optimal-routing program

    maximise  c' psi + const                     psi = sum_i A_i (Lambda_i - Delta_i)
    s.t.      phi_i(R_i + gamma_i Delta_i - Lambda_i) >= phi_i(R_i)      (geo_mean / 2-token product / sum with x >= 0 /
                                                                         product on virtual reserves with real x >= 0)
              psi_j + a_j >= 0  |  psi_j + a_j == 0  |  psi_j free       (token by token)
              Delta_i, Lambda_i >= 0

in the expression graph -- the pools (reserves, fee, trading function, weights), which local slot is which token, and
the linear + box utility -- hands those literals to ``api.solve`` (CUDA; there is no CPU path), and writes the optimal
trades back into the Variables, so every ``.value`` the scripts read afterwards (prob.value, psi.value, deltas[i].value,
lambdas[i].value, obj.value) is an evaluation of their own expressions.  A model outside that family raises
``NotRoutingProblem`` naming the construct it could not place; nothing is approximated silently.
"""
from __future__ import annotations

import builtins
import dataclasses
from typing import Callable, List, Optional

import numpy as np

__version__ = "cfmm_routing_code_b200.cvxpy_compat"

OPTIMAL, OPTIMAL_INACCURATE, INFEASIBLE, USER_LIMIT = "optimal", "optimal_inaccurate", "infeasible", "user_limit"


class NotRoutingProblem(NotImplementedError):
    """The model is not an optimal-routing program of the reference's form (or uses cvxpy features outside the subset)."""


# ----------------------------------------------------------------------------------------------------------------
# affine expressions over Variables
# ----------------------------------------------------------------------------------------------------------------
class Expression:
    """value = const + sum_v coef[v] @ v ;  a vector of `rows` entries, or a scalar (rows == 1, scalar=True)"""
    __array_ufunc__ = None               # ndarray (op) Expression defers to the reflected operators below
    __hash__ = object.__hash__

    def __init__(self, coef, const, scalar):
        self.coef = coef                 # {Variable: ndarray (rows, v.size)}
        self.const = np.atleast_1d(np.asarray(const, float)).copy()
        self.scalar = bool(scalar)

    @property
    def rows(self):
        return len(self.const)

    @property
    def shape(self):
        return () if self.scalar else (self.rows,)

    @property
    def size(self):
        return self.rows

    def __len__(self):
        if self.scalar:
            raise TypeError("len() of a scalar expression")
        return self.rows

    def variables(self):
        return list(self.coef)

    def _stretched(self, rows):
        if self.rows == rows:
            return self
        if not self.scalar:
            raise ValueError(f"cannot combine expressions of {self.rows} and {rows} entries")
        return Expression({v: np.repeat(C, rows, 0) for v, C in self.coef.items()}, np.repeat(self.const, rows), False)

    # ---- arithmetic
    def __add__(self, other):
        o = _expr(other)
        rows = max(self.rows, o.rows)
        a, b = self._stretched(rows), o._stretched(rows)
        coef = {v: C.copy() for v, C in a.coef.items()}
        for v, C in b.coef.items():
            coef[v] = coef[v] + C if v in coef else C.copy()
        return Expression(coef, a.const + b.const, self.scalar and o.scalar)

    __radd__ = __add__

    def __neg__(self):
        return Expression({v: -C for v, C in self.coef.items()}, -self.const, self.scalar)

    def __sub__(self, other):
        return self + (-_expr(other))

    def __rsub__(self, other):
        return _expr(other) + (-self)

    def __mul__(self, other):
        if isinstance(other, Expression):
            if other.coef and self.coef:
                raise NotRoutingProblem("product of two expressions that both contain variables")
            if other.coef:
                return other * self
            other = other.const[0] if other.scalar else other.const
        k = np.asarray(other, float)
        if k.ndim == 0:
            return Expression({v: C * float(k) for v, C in self.coef.items()}, self.const * float(k), self.scalar)
        if k.ndim != 1:
            raise NotRoutingProblem("elementwise product with a matrix")
        e = self._stretched(len(k))
        return Expression({v: C * k[:, None] for v, C in e.coef.items()}, e.const * k, False)

    __rmul__ = __mul__

    def __truediv__(self, other):
        return self * (1.0 / np.asarray(other, float))

    def __rmatmul__(self, mat):          # ndarray @ expression
        M = np.asarray(mat, float)
        if self.scalar or M.ndim not in (1, 2) or M.shape[-1] != self.rows:
            raise ValueError(f"matmul shapes {M.shape} @ {self.shape}")
        M2 = np.atleast_2d(M)
        return Expression({v: M2 @ C for v, C in self.coef.items()}, M2 @ self.const, M.ndim == 1)

    def __matmul__(self, vec):           # expression @ ndarray (a dot product)
        w = np.asarray(vec, float)
        if w.ndim != 1:
            raise NotRoutingProblem("expression @ matrix")
        return w @ self

    def __getitem__(self, key):
        if self.scalar:
            raise IndexError("scalar expression")
        idx = np.arange(self.rows)[key]
        one = np.ndim(idx) == 0
        idx = np.atleast_1d(idx)
        return Expression({v: C[idx] for v, C in self.coef.items()}, self.const[idx], one)

    # ---- relations
    def __ge__(self, other):
        if isinstance(other, _GeoMean):
            return other <= self
        return Constraint(self - other, ">=")

    def __le__(self, other):
        if isinstance(other, _GeoMean):
            return other >= self
        return Constraint(_expr(other) - self, ">=")

    def __eq__(self, other):             # noqa: cvxpy semantics: builds a constraint
        return Constraint(self - other, "==")

    # ---- evaluation
    @property
    def value(self):
        out = self.const.copy()
        for v, C in self.coef.items():
            if v._value is None:
                return None
            out = out + C @ v._value
        return float(out[0]) if self.scalar else out


class Variable(Expression):
    def __init__(self, shape=(), nonneg=False, name=None, **unsupported):
        if unsupported:
            raise NotRoutingProblem(f"Variable attributes {sorted(unsupported)} are outside the supported subset")
        if isinstance(shape, (tuple, list)):
            if len(shape) > 1:
                raise NotRoutingProblem("matrix variables")
            n, scalar = (int(shape[0]), False) if len(shape) else (1, True)
        else:
            n, scalar = int(shape), False
        self.nonneg = bool(nonneg)
        self.name = name
        self._value = None
        Expression.__init__(self, {self: np.eye(n)}, np.zeros(n), scalar)

    __hash__ = object.__hash__

    @property
    def value(self):
        if self._value is None:
            return None
        return float(self._value[0]) if self.scalar else self._value

    @value.setter
    def value(self, v):
        self._value = None if v is None else np.atleast_1d(np.asarray(v, float)).copy()


def _expr(x) -> Expression:
    if isinstance(x, Expression):
        return x
    if isinstance(x, _GeoMean):
        raise NotRoutingProblem("geo_mean(...) may only be compared with a level (geo_mean(x) >= number)")
    a = np.asarray(x, float)
    if a.ndim > 1:
        raise NotRoutingProblem("matrix constants in an expression")
    return Expression({}, a, a.ndim == 0)


@dataclasses.dataclass(eq=False)
class Constraint:
    expr: Expression            # expr >= 0   or   expr == 0   (row by row)
    op: str


class _GeoMean:
    """geo_mean(x, p) of an affine vector x: only ever compared with a level"""
    __array_ufunc__ = None

    def __init__(self, x: Expression, w: np.ndarray):
        self.x, self.w = x, w

    def _level(self, other):
        if isinstance(other, (Expression, _GeoMean)):
            if isinstance(other, Expression) and not other.coef and other.scalar:
                return float(other.const[0])
            raise NotRoutingProblem("geo_mean(new_reserves) must be compared with a constant level")
        return float(other)

    def __ge__(self, other):
        return PoolConstraint(self.x, self.w, self._level(other))

    def __le__(self, other):
        raise NotRoutingProblem("geo_mean(x) <= level is not convex")

    @property
    def value(self):
        x = self.x.value
        return None if x is None else float(np.prod(np.asarray(x, float) ** self.w))


@dataclasses.dataclass(eq=False)
class PoolConstraint:
    x: Expression               # geo_mean(x, w) >= level
    w: np.ndarray
    level: float


def _weights(p, k):
    if p is None:
        return np.full(k, 1.0 / k)
    w = np.asarray(p, float)
    if w.shape != (k,) or np.any(w <= 0):
        raise ValueError("geo_mean: p needs one positive weight per entry")
    return w / w.sum()


def geo_mean(x, p=None):
    """prod x_i^(p_i / sum p)   (arbitrage.py:65: the weighted pool; :68-70: p=None on two tokens = constant product)"""
    if isinstance(x, Expression) and x.coef:
        if x.scalar:
            raise ValueError("geo_mean of a scalar")
        return _GeoMean(x, _weights(p, x.rows))
    a = np.asarray(x.const if isinstance(x, Expression) else x, float)
    return float(np.prod(a ** _weights(p, len(a))))


def sum(x, axis=None):                   # noqa: A001  (the cvxpy name)
    """cvxpy semantics: a python list is added up entry by entry (arbitrage.py:54); an expression is totalled (:73)"""
    if isinstance(x, (list, tuple)) and builtins.any(isinstance(e, Expression) for e in x):
        return builtins.sum(x[1:], x[0])
    if isinstance(x, Expression):
        if axis is not None:
            raise NotRoutingProblem("sum(axis=...)")
        return x if x.scalar else np.ones(x.rows) @ x
    return float(np.sum(np.asarray(x, float)))


class Maximize:
    sign = 1.0

    def __init__(self, expr):
        e = _expr(expr)
        if not e.scalar:
            raise ValueError("the objective must be a scalar expression")
        self.expr = e

    @property
    def value(self):                     # two-asset.py:100 reads obj.value
        return self.expr.value


class Minimize(Maximize):
    sign = -1.0


# ----------------------------------------------------------------------------------------------------------------
# recognising the routing program
# ----------------------------------------------------------------------------------------------------------------
@dataclasses.dataclass
class RoutingModel:
    """the literals of `api.solve` (arbitrage.py:5-36) as found in the expression graph"""
    n_tokens: int
    local_indices: List[List[int]]
    reserves: List[np.ndarray]
    fees: List[float]
    kinds: List[str]
    weights: List[Optional[np.ndarray]]
    c: np.ndarray
    a: np.ndarray
    eq: np.ndarray
    pinned: np.ndarray
    obj_const: float
    trades: list                # [(Delta_i Variable, Lambda_i Variable)] in pool order


_RTOL = 1e-9


def _reserves_after_trade(x: Expression, allow_zero: bool = False):
    """x == R + gamma * Delta - Lambda  (arbitrage.py:60) -> (Delta, Lambda, R, gamma), else None"""
    if x.scalar or len(x.coef) != 2:
        return None
    k = x.rows
    found = {}
    for v, C in x.coef.items():
        if not isinstance(v, Variable) or v.scalar or v.rows != k or C.shape != (k, k):
            return None
        d = np.diag(C)
        if np.any(C != np.diag(d)) or np.any(d != d[0]):
            return None
        found["lam" if d[0] == -1.0 else "del"] = (v, float(d[0]))
    if set(found) != {"lam", "del"}:
        return None
    gamma = found["del"][1]
    if not (0.0 < gamma <= 1.0) or np.any(x.const < 0) or (not allow_zero and np.any(x.const == 0)):
        return None
    return found["del"][0], found["lam"][0], x.const.copy(), gamma


def recognise(objective: Maximize, constraints) -> RoutingModel:
    # NB: Expression.__eq__ builds a Constraint (cvxpy semantics), so Variables are only ever compared by identity here
    pools = {}                     # (id(Delta), id(Lambda)) -> dict(D, L, R, gamma, kind, w, nonneg)
    order = []
    token_rows = []                # (Expression, op)
    sum_levels = []                # scalar affine constraints that look like a trading-function level

    def pool_of(pos, D, L, R, gamma):
        key = (id(D), id(L))
        if key not in pools:
            for k2 in pools:
                if set(k2) & set(key):
                    raise NotRoutingProblem("a trade variable is used by two different pools")
            if not (D.nonneg and L.nonneg):
                raise NotRoutingProblem("tendered / received baskets must be Variable(..., nonneg=True)   (arbitrage.py:51-52)")
            pools[key] = dict(D=D, L=L, R=R, gamma=gamma, kind=None, w=None, nonneg=False, pos=pos)
            order.append(key)
        p = pools[key]
        if np.any(p["R"] != R) or p["gamma"] != gamma:
            raise NotRoutingProblem("two constraints describe the reserves of one pool differently")
        return p

    affine = []
    for pos, con in enumerate(constraints):          # trading functions first: they define which variables are pools
        if isinstance(con, PoolConstraint):
            nr = _reserves_after_trade(con.x)
            if nr is None:
                raise NotRoutingProblem("geo_mean(x) >= level: x is not of the form R + gamma * Delta - Lambda   (arbitrage.py:60)")
            p = pool_of(pos, *nr)
            here = float(np.prod(nr[2] ** con.w))
            if abs(con.level - here) > _RTOL * abs(here):
                raise NotRoutingProblem(f"geo_mean level {con.level!r} differs from the trading function at the current reserves "
                                        f"({here!r}): only phi(new reserves) >= phi(reserves) is supported")
            if p["kind"] is not None:
                raise NotRoutingProblem("two trading-function constraints on one pool")
            equal = bool(np.all(con.w == con.w[0]))
            p["kind"], p["w"] = ("product", None) if (len(con.w) == 2 and equal) else ("geomean", con.w.copy())
        elif isinstance(con, Constraint):
            affine.append((pos, con))
        elif isinstance(con, (bool, np.bool_)):
            if not con:
                raise NotRoutingProblem("a constraint between constants is False")
        else:
            raise NotRoutingProblem(f"unsupported constraint object {type(con).__name__}")
    for pos, con in affine:
        nr = _reserves_after_trade(con.expr) if con.op == ">=" else None
        if nr is not None and (id(nr[1]), id(nr[0])) in pools:
            nr = None                                # psi + a >= 0 of a one-pool problem: the known pool's net flow, not reserves
        real = _reserves_after_trade(con.expr, allow_zero=True) if con.op == ">=" else None
        p = pools.get((id(real[0]), id(real[1]))) if real is not None else None
        if p is not None and p["kind"] == "product" and p["gamma"] == real[3] and np.all(real[2] <= p["R"]) and np.any(real[2] < p["R"]):
            # geo_mean(R + o + gamma D - L) >= geo_mean(R + o) together with R + gamma D - L >= 0: a constant product on
            # VIRTUAL reserves whose real reserves stay non-negative (one tick range of a concentrated-liquidity pool)
            p["kind"], p["w"], p["R"], p["nonneg"] = "bounded_product", p["R"] - real[2], real[2], True
        elif nr is not None:                         # new_reserves >= 0   (arbitrage.py:74)
            pool_of(pos, *nr)["nonneg"] = True
        elif con.op == ">=" and con.expr.scalar and _looks_like_sum_level(con.expr):
            sum_levels.append(con.expr)
        else:
            token_rows.append((con.expr, con.op))

    for e in sum_levels:                             # sum(new_reserves) >= sum(reserves)   (arbitrage.py:73)
        (va, Ca), (vb, Cb) = e.coef.items()
        D, L = (va, vb) if Ca[0, 0] > 0 else (vb, va)
        if (id(D), id(L)) not in pools:
            raise NotRoutingProblem("sum(new_reserves) >= level without new_reserves >= 0 on the same pool: the constant-sum "
                                    "pool needs both   (arbitrage.py:73-74)")
        p = pools[(id(D), id(L))]
        gam = e.coef[D][0]
        if p["kind"] is not None or np.any(gam != p["gamma"]) or np.any(e.coef[L][0] != -1.0):
            raise NotRoutingProblem("sum(...) >= level does not match the pool its variables belong to")
        if abs(e.const[0]) > _RTOL * p["R"].sum():   # e = sum(R) - level + gamma sum(D) - sum(L)
            raise NotRoutingProblem("sum level differs from the current reserves' sum: only phi(new) >= phi(current) is supported")
        if not p["nonneg"]:
            raise NotRoutingProblem("constant-sum pool without new_reserves >= 0")
        p["kind"] = "sum"
    order.sort(key=lambda k: pools[k]["pos"])        # pools in the order the model states them
    for key in order:
        if pools[key]["kind"] is None:
            raise NotRoutingProblem("a pool has reserves (new_reserves >= 0) but no trading-function constraint")
    if not order:
        raise NotRoutingProblem("no pool (trading-function constraint) in the problem")

    # ---- tokens: local slot (pool i, position s) <-> global token.  Every remaining row is u' psi + const with
    # psi = sum_i A_i (Lambda_i - Delta_i); two slots are the same token iff they carry the same coefficient in every row.
    rows = [(e._stretched(e.rows), op) for e, op in token_rows] + [(objective.expr * objective.sign, "obj")]
    nrow = builtins.sum(e.rows for e, _ in rows)
    slot_cols = []
    for key in order:
        D, L = pools[key]["D"], pools[key]["L"]
        k = D.rows
        col = np.zeros((nrow, k))
        r0 = 0
        for e, _ in rows:
            CL = e.coef.get(L, np.zeros((e.rows, k)))
            CD = e.coef.get(D, np.zeros((e.rows, k)))
            if np.any(CL != -CD):
                raise NotRoutingProblem("a constraint or the objective uses Delta and Lambda other than through the net flow "
                                        "Lambda - Delta   (arbitrage.py:54)")
            col[r0:r0 + e.rows] = CL
            r0 += e.rows
        slot_cols.append(col)
    known = {i for key in order for i in key}
    for e, _ in rows:
        for v in e.coef:
            if id(v) not in known:
                raise NotRoutingProblem("a variable appears that is not the trade of any pool")
    token_cols, local_indices = [], []
    for col in slot_cols:
        li = []
        for s in range(col.shape[1]):
            c_ = col[:, s]
            if not np.any(c_):
                raise NotRoutingProblem("a pool trades a token that neither the objective nor any constraint mentions: its net "
                                        "flow would be free and worthless (unbounded)")
            for j, t in enumerate(token_cols):
                if np.array_equal(t, c_):
                    li.append(j)
                    break
            else:
                token_cols.append(c_.copy())
                li.append(len(token_cols) - 1)
        if len(set(li)) != len(li):
            raise NotRoutingProblem("a pool lists the same token twice")
        local_indices.append(li)
    n = len(token_cols)
    U = np.stack(token_cols, 1)                      # [row, token]: the row reads  U[row] @ psi + const
    c = np.zeros(n); a = np.full(n, np.inf); eq = np.zeros(n, bool); has = np.zeros(n, bool)
    obj_const = 0.0
    r0 = 0
    for e, op in rows:
        for r in range(e.rows):
            u, k0 = U[r0 + r], float(e.const[r])
            if op == "obj":
                c, obj_const = u.copy(), k0
                continue
            nz = np.nonzero(u)[0]
            if len(nz) == 0:
                if (op == ">=" and k0 < 0) or (op == "==" and k0 != 0):
                    raise NotRoutingProblem("a constant constraint is violated (infeasible as stated)")
                continue
            if len(nz) > 1:
                raise NotRoutingProblem("a constraint couples the net flows of several tokens; supported: psi_j + a_j >= 0 | == 0 "
                                        "token by token   (arbitrage.py:77, liquidation.py:77-80)")
            j = int(nz[0])
            if op == ">=" and u[j] < 0:
                raise NotRoutingProblem("upper bound on a net flow (psi_j <= b)")
            aj = k0 / u[j]
            if op == "==":
                if eq[j] and a[j] != aj:
                    raise NotRoutingProblem("two different equalities on one token")
                if has[j] and not eq[j] and aj > a[j]:
                    raise NotRoutingProblem("an equality and a tighter inequality on one token")
                eq[j], a[j] = True, aj
            elif eq[j]:
                if a[j] > aj:                       # psi_j = -a_j must also satisfy psi_j >= -aj
                    raise NotRoutingProblem("an equality and a tighter inequality on one token")
            else:
                a[j] = min(a[j], aj)
            has[j] = True
        r0 += e.rows
    pinned = ~has
    a = np.where(has, a, 0.0)
    if np.any(c < 0):
        raise NotRoutingProblem("negative objective weight on a net flow")
    if np.any(pinned & (c <= 0)):
        raise NotRoutingProblem("a token is neither constrained nor valued by the objective (unbounded)")
    return RoutingModel(n, local_indices, [pools[k]["R"] for k in order], [pools[k]["gamma"] for k in order],
                        [pools[k]["kind"] for k in order], [pools[k]["w"] for k in order], c, a, eq, pinned, obj_const,
                        [(pools[k]["D"], pools[k]["L"]) for k in order])


def _looks_like_sum_level(e: Expression) -> bool:
    """const + gamma * sum(Delta) - sum(Lambda)  with 0 < gamma <= 1: the total of a pool's reserves after the trade"""
    if len(e.coef) != 2:
        return False
    (va, Ca), (vb, Cb) = e.coef.items()
    if Ca.shape != Cb.shape or Ca.shape[0] != 1:
        return False
    lo, hi = (Ca[0], Cb[0]) if Ca[0, 0] < 0 else (Cb[0], Ca[0])
    return bool(np.all(lo == -1.0) and np.all(hi == hi[0]) and 0.0 < hi[0] <= 1.0 and (hi[0] < 1.0 or e.const[0] == 0.0))


# ----------------------------------------------------------------------------------------------------------------
_backend: Optional[Callable] = None      # None = api.solve (CUDA).  Tests install a checker here; the product never does.


class Problem:
    def __init__(self, objective, constraints=()):
        if not isinstance(objective, Maximize):
            raise NotRoutingProblem("Problem(objective, ...): objective must be Maximize(...) or Minimize(...)")
        self.objective = objective
        self.constraints = list(constraints)
        self.value = None
        self.status = None
        self.result = None               # the api.Result of the last solve (nu, gap, iteration counts, ...)
        self.model: Optional[RoutingModel] = None

    def variables(self):
        seen = {}
        exprs = [self.objective.expr] + [con.x if isinstance(con, PoolConstraint) else getattr(con, "expr", None)
                                          for con in self.constraints]
        for e in exprs:
            for v in (e.variables() if e is not None else []):
                seen.setdefault(id(v), v)
        return list(seen.values())

    def solve(self, tol: float = 1e-9, verbose: bool = False, **kw):
        """arbitrage.py:82.  Recognise the routing program, solve it on the GPU, write the trades back."""
        for ignored in ("solver", "warm_start", "ignore_dpp", "gp", "qcp", "requires_grad", "enforce_dpp"):
            kw.pop(ignored, None)        # cvxpy arguments that select or tune ITS back ends: meaningless here
        for alias in ("max_iters", "max_iter"):
            if alias in kw:
                kw["max_iter"] = int(kw.pop(alias))
        m = self.model = recognise(self.objective, self.constraints)
        from .api import LinearUtility, solve as gpu_solve           # torch + the CUDA library load here, not at import
        util = LinearUtility(m.c, m.a, m.eq, m.pinned)
        run = gpu_solve if _backend is None else _backend
        r = self.result = run(m.local_indices, m.reserves, m.fees, m.kinds, m.weights, utility=util, n_tokens=m.n_tokens,
                              tol=tol, verbose=verbose, **kw)
        for (D, L), d, l in zip(m.trades, r.deltas, r.lambdas):
            D.value, L.value = np.asarray(d, float), np.asarray(l, float)
        self.status = {"optimal": OPTIMAL, "infeasible": INFEASIBLE}.get(r.status, USER_LIMIT)
        if self.status == INFEASIBLE:
            for D, L in m.trades:
                D.value = L.value = None
            self.value = -np.inf * self.objective.sign
            return self.value
        self.value = float(self.objective.expr.value)
        return self.value
