"""Rate-expression compiler: PyREMOT ``reaction-rates`` dicts -> CUDA device functions.

Replaces the interpreted executor `PyREMOT/docs/rmtReaction.py:11-61` (``reactionRateExe``): there,
every RHS evaluation walks the ``VARS`` dict in insertion order calling ~33 Python lambdas and then the
``RATES`` lambdas.  Here the same dicts are *traced once* into an expression DAG (hash-consed, constant
folded) and emitted as two straight-line CUDA functions that the solver kernels inline through NVRTC:

    pbr_rates    (T, P, x[NC], C[NC], par[NP]) -> r[NR]
    pbr_rates_jac(...)  -> r[NR], dr/dT, dr/dP, dr/dx[NR][NC], dr/dC[NR][NC]   (symbolic forward mode)

Semantics kept from the reference executor: evaluation order = dict insertion order; a ``VARS`` entry
that is a number is a literal; a forward reference raises ``KeyError``; seeds are ``R_CONST, T, P,
MoFri, SpCoi`` (`rmtReaction.py:29-35`).  Scalar ``VARS`` entries named in ``params`` become
per-instance runtime parameters ``par[k]`` (how sweeps / parameter estimation feed kinetic constants
without recompiling); all other scalars are folded as constants.

Front-ends: python lambdas (traced by calling them with a recording dict and a recording ``math``),
or the dashboard text form (`README.md:88-120, 160-164`: ``"name": <expr>;`` lines) which is parsed
with ``ast`` into the same lambdas.  Grammar (census in SURVEY.md Appendix B): ``+ - * / **``, unary
minus, ``math.exp/log/log10/sqrt/pow``, literals, ``x['NAME']``, ``x['MoFri'][i]``, ``x['SpCoi'][i]``.
"""
from __future__ import annotations

import ast
import hashlib
import math
import os
import types

R_CONST = 8.314472  # PyREMOT/core/constants.py:8


class RateCompileError(Exception):
    pass


# ----------------------------------------------------------------------------- DAG
class Graph:
    """Hash-consed expression DAG. Node = (op, args...) ; ids are topologically ordered."""

    def __init__(self):
        self.nodes = []
        self.index = {}

    def _mk(self, key):
        i = self.index.get(key)
        if i is None:
            i = len(self.nodes)
            self.nodes.append(key)
            self.index[key] = i
        return i

    def const(self, v):
        v = float(v)
        if v == 0.0:
            v = 0.0  # fold -0.0
        return self._mk(("const", v))

    def inp(self, kind, idx=0):
        return self._mk(("in", kind, idx))

    def is_const(self, i, v=None):
        n = self.nodes[i]
        return n[0] == "const" and (v is None or n[1] == v)

    def cval(self, i):
        return self.nodes[i][1]

    # --- arithmetic with folding / identities
    def add(self, a, b):
        if self.is_const(a) and self.is_const(b):
            return self.const(self.cval(a) + self.cval(b))
        if self.is_const(a, 0.0):
            return b
        if self.is_const(b, 0.0):
            return a
        return self._mk(("add", a, b))

    def sub(self, a, b):
        if self.is_const(a) and self.is_const(b):
            return self.const(self.cval(a) - self.cval(b))
        if self.is_const(b, 0.0):
            return a
        if self.is_const(a, 0.0):
            return self.neg(b)
        return self._mk(("sub", a, b))

    def mul(self, a, b):
        if self.is_const(a) and self.is_const(b):
            return self.const(self.cval(a)*self.cval(b))
        if self.is_const(a, 0.0) or self.is_const(b, 0.0):
            return self.const(0.0)
        if self.is_const(a, 1.0):
            return b
        if self.is_const(b, 1.0):
            return a
        if self.is_const(a, -1.0):
            return self.neg(b)
        if self.is_const(b, -1.0):
            return self.neg(a)
        return self._mk(("mul", a, b))

    def div(self, a, b):
        """a/b is emitted as a * rcp(b): one FP64 division per distinct denominator instead of one per use
        (an FP64 divide is ~12 instructions on the device); the result differs from a true divide by <= 1 ulp."""
        if self.is_const(a) and self.is_const(b):
            return self.const(self.cval(a)/self.cval(b))
        if self.is_const(a, 0.0):
            return self.const(0.0)
        if self.is_const(b, 1.0):
            return a
        if self.is_const(b):
            return self.mul(a, self.const(1.0/self.cval(b)))
        return self.mul(a, self._mk(("rcp", b)))

    def neg(self, a):
        if self.is_const(a):
            return self.const(-self.cval(a))
        if self.nodes[a][0] == "neg":
            return self.nodes[a][1]
        return self._mk(("neg", a))

    def fn(self, name, a):
        if self.is_const(a):
            return self.const(getattr(math, name)(self.cval(a)))
        return self._mk((name, a))

    def pow(self, a, b):
        if self.is_const(a) and self.is_const(b):
            return self.const(math.pow(self.cval(a), self.cval(b)))
        if self.is_const(b):
            e = self.cval(b)
            if e == 0.0:
                return self.const(1.0)
            if e == 1.0:
                return a
            if e == int(e) and abs(e) <= 16:
                return self._mk(("powi", a, int(e)))
            if e == 0.5:
                return self.fn("sqrt", a)
        return self._mk(("pow", a, b))


class Expr:
    """Recording scalar handed to the user's lambdas."""
    __slots__ = ("g", "i")

    def __init__(self, g, i):
        self.g, self.i = g, i

    def _w(self, o):
        if isinstance(o, Expr):
            return o.i
        if isinstance(o, (int, float)) or hasattr(o, "__float__"):
            return self.g.const(float(o))
        raise RateCompileError(f"unsupported operand in rate expression: {o!r}")

    def __add__(self, o): return Expr(self.g, self.g.add(self.i, self._w(o)))
    def __radd__(self, o): return Expr(self.g, self.g.add(self._w(o), self.i))
    def __sub__(self, o): return Expr(self.g, self.g.sub(self.i, self._w(o)))
    def __rsub__(self, o): return Expr(self.g, self.g.sub(self._w(o), self.i))
    def __mul__(self, o): return Expr(self.g, self.g.mul(self.i, self._w(o)))
    def __rmul__(self, o): return Expr(self.g, self.g.mul(self._w(o), self.i))
    def __truediv__(self, o): return Expr(self.g, self.g.div(self.i, self._w(o)))
    def __rtruediv__(self, o): return Expr(self.g, self.g.div(self._w(o), self.i))
    def __pow__(self, o): return Expr(self.g, self.g.pow(self.i, self._w(o)))
    def __rpow__(self, o): return Expr(self.g, self.g.pow(self._w(o), self.i))
    def __neg__(self): return Expr(self.g, self.g.neg(self.i))
    def __pos__(self): return self

    def _nope(self, *a):
        raise RateCompileError("rate expressions must be straight-line arithmetic: comparisons / branches / "
                               "conversions on T, P, MoFri, SpCoi cannot be compiled")
    __bool__ = __float__ = __int__ = __lt__ = __le__ = __gt__ = __ge__ = _nope


class _TraceMath:
    """Stands in for ``math`` (and ``np``) inside the traced lambdas."""
    pi, e = math.pi, math.e

    def __init__(self, g):
        self._g = g

    def _f(self, name, a):
        if isinstance(a, Expr):
            return Expr(self._g, self._g.fn(name, a.i))
        return getattr(math, name)(a)

    def exp(self, a): return self._f("exp", a)
    def log(self, a, *base):
        if base:
            return self._f("log", a)/self._f("log", base[0])
        return self._f("log", a)
    def log10(self, a): return self._f("log10", a)
    def sqrt(self, a): return self._f("sqrt", a)

    def pow(self, a, b):
        if isinstance(a, Expr):
            return a**b
        if isinstance(b, Expr):
            return b.__rpow__(a)
        return math.pow(a, b)
    power = pow


class _TraceDict(dict):
    def __missing__(self, k):
        raise KeyError(k)  # same failure as the reference executor on a forward reference


def _rebind(f, tm):
    g = dict(f.__globals__)
    for name in ("math", "np", "numpy"):
        g[name] = tm
    return types.FunctionType(f.__code__, g, f.__name__, f.__defaults__, f.__closure__)


# ----------------------------------------------------------------------------- dashboard text front-end
_ALLOWED_NODES = (ast.Expression, ast.BinOp, ast.UnaryOp, ast.Add, ast.Sub, ast.Mult, ast.Div, ast.Pow,
                  ast.USub, ast.UAdd, ast.Constant, ast.Subscript, ast.Name, ast.Load, ast.Call,
                  ast.Attribute)


def parse_text(text, namespace=None):
    """Dashboard form ``"name": <expr>;`` (one per ';') -> ordered dict of numbers / lambdas.

    Bare identifiers (``"CaBeDe" : CaBeDe``) are looked up in ``namespace``. Only the arithmetic grammar
    above is accepted (validated on the AST before anything is evaluated).
    """
    out = {}
    namespace = dict(namespace or {})
    for item in text.split(";"):
        item = item.strip()
        if not item:
            continue
        key, sep, expr = item.partition(":")
        if not sep:
            raise RateCompileError(f"cannot parse rate line {item!r}")
        key = key.strip().strip("\"'")
        tree = ast.parse(expr.strip(), mode="eval")
        uses_x = False
        for nd in ast.walk(tree):
            if not isinstance(nd, _ALLOWED_NODES):
                raise RateCompileError(f"unsupported syntax {type(nd).__name__} in {key!r}")
            if isinstance(nd, ast.Name):
                if nd.id == "x":
                    uses_x = True
                elif nd.id not in ("math", "np") and nd.id not in namespace:
                    raise RateCompileError(f"unknown name {nd.id!r} in {key!r}")
            if isinstance(nd, ast.Attribute) and not (isinstance(nd.value, ast.Name) and nd.value.id in ("math", "np")
                                                      and nd.attr in ("exp", "log", "log10", "sqrt", "pow", "power", "pi", "e")):
                raise RateCompileError(f"unsupported attribute in {key!r}")
        env = {"math": math, "__builtins__": {}, **namespace}
        if uses_x:
            lam = ast.fix_missing_locations(ast.Expression(body=ast.Lambda(
                args=ast.arguments(posonlyargs=[], args=[ast.arg(arg="x")], kwonlyargs=[], kw_defaults=[], defaults=[]),
                body=tree.body)))
            out[key] = eval(compile(lam, f"<rate:{key}>", "eval"), env)
        else:
            out[key] = eval(compile(tree, f"<rate:{key}>", "eval"), env)
    return out


# ----------------------------------------------------------------------------- compiler
class CompiledRates:
    """Result of :func:`compile_rates`: the DAG, output node ids, parameter names and CUDA source."""

    def __init__(self, g, outs, n_comp, params, param_defaults, var_nodes):
        self.g, self.outs, self.n_comp = g, outs, n_comp
        self.params, self.param_defaults = list(params), list(param_defaults)
        self.var_nodes = var_nodes
        self.n_rxn = len(outs)
        self._src = None

    # ---- numeric evaluation of the DAG (host; used by tests to check the tracer, not by the product path)
    def evaluate(self, T, P, x, C, par=()):
        vals = _eval_nodes(self.g, {"T": [T], "P": [P], "X": list(x), "C": list(C), "PAR": list(par)})
        return [vals[o] for o in self.outs]

    def op_counts(self):
        """(ops in pbr_rates, ops in pbr_rates_jac): +,-,*,/ and each exp/log/log10/pow/sqrt call count 1,
        x**k counts its multiplications.  Used for the algorithmic-flop figure of the roofline report."""
        g = self.g
        n0 = len(g.nodes)
        derivs = _derivatives(g, self.outs, self.n_comp)

        def count(needed):
            live, stack = set(), list(needed)
            while stack:
                i = stack.pop()
                if i not in live:
                    live.add(i)
                    stack.extend(_children(g.nodes[i]))
            c = 0
            for i in live:
                op = g.nodes[i][0]
                if op in ("const", "in"):
                    continue
                c += max(1, abs(g.nodes[i][2]).bit_length()) if op == "powi" else 1
            return c
        need = list(self.outs)
        for dj in derivs:
            need.extend(v for k, v in dj.items() if k[0] != "PAR")
        return count(self.outs), count(need)

    @property
    def cuda_source(self):
        if self._src is None:
            self._src = _emit(self)
        return self._src

    @property
    def key(self):
        return hashlib.sha256(self.cuda_source.encode()).hexdigest()[:16]


def compile_rates(rate_expr, n_comp, params=(), optimize=None):
    """Trace ``{"VARS": {...}, "RATES": {...}}`` (lambdas, numbers, or dashboard text) into a CompiledRates."""
    VARS, RATES = rate_expr["VARS"], rate_expr["RATES"]
    if isinstance(VARS, str):
        VARS = parse_text(VARS)
    if isinstance(RATES, str):
        RATES = parse_text(RATES)
    g = Graph()
    tm = _TraceMath(g)
    d = _TraceDict()
    d["R_CONST"] = R_CONST
    d["T"] = Expr(g, g.inp("T"))
    d["P"] = Expr(g, g.inp("P"))
    d["MoFri"] = [Expr(g, g.inp("X", i)) for i in range(n_comp)]
    d["SpCoi"] = [Expr(g, g.inp("C", i)) for i in range(n_comp)]
    params = list(params)
    defaults = []
    for p in params:
        if p not in VARS or isinstance(VARS[p], types.FunctionType):
            raise RateCompileError(f"sweep parameter {p!r} is not a scalar entry of VARS")
    var_nodes = {}
    for k, v in VARS.items():
        if k in params:
            d[k] = Expr(g, g.inp("PAR", params.index(k)))
        elif isinstance(v, types.FunctionType):
            d[k] = _rebind(v, tm)(d)
        else:
            d[k] = v
        if isinstance(d[k], Expr):
            var_nodes[k] = d[k].i
    for p in params:
        defaults.append(float(VARS[p]))
    outs = []
    for k, f in RATES.items():
        v = _rebind(f, tm)(d) if isinstance(f, types.FunctionType) else f
        outs.append(v.i if isinstance(v, Expr) else g.const(float(v)))
    if optimize is None:
        optimize = os.environ.get("PBR_RATEC_OPT", "1") != "0"      # A/B switch for measurements
    if optimize:
        g, outs, var_nodes = _optimize(g, outs, var_nodes)
    return CompiledRates(g, outs, n_comp, params, defaults, var_nodes)


def _optimize(g, outs, var_nodes):
    """Strength-reducing rewrites on the traced DAG (values change by <= a few ulp, the op count drops):

    * ``log10(u) -> log(u) * (1/ln 10)``: shares the logarithm with ``log(u)`` when both occur (KP2 / KP1, KP3 of the README kinetics);
    * ``1/exp(u) -> exp(-u)`` and ``1/10**u -> 10**(-u)`` when the exponential is used in no other way (equilibrium constants
      that only ever divide): one reciprocal less per site, and its partials lose the ``-(1/u)^2`` factor;
    * ``1/(c*u) -> (1/c) * (1/u)`` and ``1/u**k -> (1/u)**k``: reciprocals of ``R*T``, ``T**3`` ... collapse onto the one of ``T``.
    Returns (new graph, new outs, new var_nodes)."""
    live, stack = set(), list(outs)
    while stack:
        i = stack.pop()
        if i not in live:
            live.add(i)
            stack.extend(_children(g.nodes[i]))
    uses = {}
    for i in sorted(live):
        for c in _children(g.nodes[i]):
            uses.setdefault(c, []).append(i)
    for o in outs:
        uses.setdefault(o, []).append(-1)

    def only_divides(i):
        return all(u >= 0 and g.nodes[u][0] == "rcp" for u in uses.get(i, [-1]))

    ng = Graph()
    m = {}
    inv_ln10 = 1.0/math.log(10.0)
    for i, n in enumerate(g.nodes):
        if i not in live and i not in var_nodes.values():
            continue
        op = n[0]
        if op == "const":
            m[i] = ng.const(n[1])
        elif op == "in":
            m[i] = ng.inp(n[1], n[2])
        elif op == "add":
            m[i] = ng.add(m[n[1]], m[n[2]])
        elif op == "sub":
            m[i] = ng.sub(m[n[1]], m[n[2]])
        elif op == "mul":
            m[i] = ng.mul(m[n[1]], m[n[2]])
        elif op == "neg":
            m[i] = ng.neg(m[n[1]])
        elif op == "powi":
            m[i] = ng.pow(m[n[1]], ng.const(float(n[2])))
        elif op == "pow":
            m[i] = ng.pow(m[n[1]], m[n[2]])
        elif op == "log10":
            m[i] = ng.mul(ng.fn("log", m[n[1]]), ng.const(inv_ln10))
        elif op == "rcp":
            a = n[1]
            an = g.nodes[a]
            if an[0] == "exp" and only_divides(a):
                m[i] = ng.fn("exp", ng.neg(m[an[1]]))
            elif an[0] == "pow" and g.is_const(an[1], 10.0) and only_divides(a):
                m[i] = ng.pow(ng.const(10.0), ng.neg(m[an[2]]))
            elif an[0] == "mul" and (g.is_const(an[1]) or g.is_const(an[2])):
                cst, u = (an[1], an[2]) if g.is_const(an[1]) else (an[2], an[1])
                m[i] = ng.mul(ng.const(1.0/g.cval(cst)), ng.div(ng.const(1.0), m[u]))
            elif an[0] == "powi" and an[2] > 0:
                m[i] = ng.pow(ng.div(ng.const(1.0), m[an[1]]), ng.const(float(an[2])))
            else:
                m[i] = ng.div(ng.const(1.0), m[a])
        else:
            m[i] = ng.fn(op, m[n[1]])
    return ng, [m[o] for o in outs], {k: m[v] for k, v in var_nodes.items() if v in m}


def _eval_nodes(g, inputs):
    vals = [0.0]*len(g.nodes)
    for i, n in enumerate(g.nodes):
        op = n[0]
        if op == "const":
            vals[i] = n[1]
        elif op == "in":
            vals[i] = inputs[n[1]][n[2]]
        elif op == "add":
            vals[i] = vals[n[1]] + vals[n[2]]
        elif op == "sub":
            vals[i] = vals[n[1]] - vals[n[2]]
        elif op == "mul":
            vals[i] = vals[n[1]]*vals[n[2]]
        elif op == "neg":
            vals[i] = -vals[n[1]]
        elif op == "rcp":
            vals[i] = 1.0/vals[n[1]]
        elif op == "powi":
            vals[i] = math.pow(vals[n[1]], n[2])
        elif op == "pow":
            vals[i] = math.pow(vals[n[1]], vals[n[2]])
        else:
            vals[i] = getattr(math, op)(vals[n[1]])
    return vals


# ----------------------------------------------------------------------------- symbolic forward-mode derivatives
def _union(da, db):
    """Keys of both partial maps in a fixed order.  (A set of tuples that contain strings iterates in an order that depends
    on PYTHONHASHSEED; node ids, the emitted text and therefore the cubin-cache key would differ between processes.)"""
    return sorted(set(da) | set(db))


def _derivatives(g, outs, n_comp):
    """For every output, partials w.r.t. T, P, X_i, C_i and the runtime parameters PAR_k as node ids in ``g`` (sparse: missing = 0)."""
    LN10 = math.log(10.0)
    dmap = {}  # node -> {input key: node}
    n_primal = len(g.nodes)
    for i in range(n_primal):
        n = g.nodes[i]
        op = n[0]
        d = {}
        if op == "in":
            if n[1] in ("T", "P"):
                d[(n[1], 0)] = g.const(1.0)
            elif n[1] in ("X", "C", "PAR"):
                d[(n[1], n[2])] = g.const(1.0)
        elif op in ("add", "sub"):
            da, db = dmap[n[1]], dmap[n[2]]
            for k in _union(da, db):
                a = da.get(k, g.const(0.0))
                b = db.get(k, g.const(0.0))
                d[k] = g.add(a, b) if op == "add" else g.sub(a, b)
        elif op == "mul":
            da, db = dmap[n[1]], dmap[n[2]]
            for k in _union(da, db):
                t = g.const(0.0)
                if k in da:
                    t = g.add(t, g.mul(da[k], n[2]))
                if k in db:
                    t = g.add(t, g.mul(n[1], db[k]))
                d[k] = t
        elif op == "rcp":      # (1/u)' = -(1/u)^2 u'
            if dmap[n[1]]:
                m = g.neg(g.mul(i, i))
                d = {k: g.mul(m, v) for k, v in dmap[n[1]].items()}
        elif op == "neg":
            d = {k: g.neg(v) for k, v in dmap[n[1]].items()}
        elif op == "exp":
            d = {k: g.mul(i, v) for k, v in dmap[n[1]].items()}
        elif op == "log":
            if dmap[n[1]]:
                r = g.div(g.const(1.0), n[1])
                d = {k: g.mul(v, r) for k, v in dmap[n[1]].items()}
        elif op == "log10":
            if dmap[n[1]]:
                r = g.div(g.const(1.0/LN10), n[1])
                d = {k: g.mul(v, r) for k, v in dmap[n[1]].items()}
        elif op == "sqrt":
            if dmap[n[1]]:
                r = g.div(g.const(0.5), i)
                d = {k: g.mul(v, r) for k, v in dmap[n[1]].items()}
        elif op == "powi":
            e = n[2]
            if dmap[n[1]]:
                pm1 = g.pow(n[1], g.const(float(e - 1)))
                d = {k: g.mul(g.mul(g.const(float(e)), pm1), v) for k, v in dmap[n[1]].items()}
        elif op == "pow":
            da, db = dmap[n[1]], dmap[n[2]]
            for k in _union(da, db):
                t = g.const(0.0)
                if k in da:   # v * u^(v-1) * u'
                    t = g.add(t, g.mul(g.mul(n[2], g.pow(n[1], g.sub(n[2], g.const(1.0)))), da[k]))
                if k in db:   # u^v * ln(u) * v'
                    t = g.add(t, g.mul(g.mul(i, g.fn("log", n[1])), db[k]))
                d[k] = t
        dmap[i] = {k: v for k, v in d.items() if not g.is_const(v, 0.0)}
    return [dmap[o] for o in outs]


# ----------------------------------------------------------------------------- CUDA emission
def _lit(v):
    if v != v or v in (float("inf"), float("-inf")):
        raise RateCompileError("non-finite constant in rate expression")
    s = repr(float(v))
    return s if ("e" in s or "." in s) else s + ".0"


class _ConstTable:
    """Literals of the rate expressions go to a __constant__ table (PBR_K[i]) instead of the instruction stream:
    a c[bank][offset] operand is free, a 64-bit immediate costs two extra move instructions on sm_100a."""

    def __init__(self):
        self.vals, self.idx = [], {}

    def ref(self, v):
        v = float(v)
        if v in (0.0, 1.0, 2.0, -1.0, 0.5):          # exactly representable in the 32-bit immediate forms
            return "(" + _lit(v) + ")" if v < 0 else _lit(v)
        i = self.idx.get(v)
        if i is None:
            i = self.idx[v] = len(self.vals)
            self.vals.append(v)
        return f"PBR_K[{i}]"

    def emit(self):
        body = ", ".join(_lit(v) for v in self.vals) if self.vals else "0.0"
        return f"PBR_CTABLE double PBR_K[{max(1, len(self.vals))}] = {{{body}}};"


def _children(n):
    op = n[0]
    if op in ("const", "in"):
        return ()
    if op in ("add", "sub", "mul", "pow"):
        return (n[1], n[2])
    return (n[1],)   # neg, powi (n[2] is the integer exponent), exp, log, log10, sqrt


def _emit_body(g, needed, lines, names, ktab=None):
    """Emit SSA statements for all nodes reachable from ``needed`` (ids), in topological (= id) order."""
    live = set()
    stack = list(needed)
    while stack:
        i = stack.pop()
        if i in live:
            continue
        live.add(i)
        stack.extend(_children(g.nodes[i]))
    for i in sorted(live):
        n = g.nodes[i]
        op = n[0]
        if i in names:
            continue
        if op == "const":
            names[i] = ktab.ref(n[1]) if ktab is not None else ("(" + _lit(n[1]) + ")" if n[1] < 0 else _lit(n[1]))
            continue
        if op == "in":
            names[i] = {"T": "T", "P": "P"}.get(n[1]) or {"X": "x", "C": "C", "PAR": "par"}[n[1]] + f"[{n[2]}]"
            continue
        a = names[n[1]]
        if op in ("add", "sub", "mul"):
            b = names[n[2]]
            rhs = f"{a} {dict(add='+', sub='-', mul='*')[op]} {b}"
        elif op == "neg":
            rhs = f"-{a}"
        elif op == "rcp":
            rhs = f"pbr_rcp({a})"
        elif op == "powi":
            e = n[2]
            rhs = _powi(a, abs(e))
            if e < 0:
                rhs = f"1.0 / ({rhs})"
        elif op == "pow":
            if g.is_const(n[1], 10.0):
                rhs = f"pbr_exp10({names[n[2]]})"
            else:
                rhs = f"pow({a}, {names[n[2]]})"
        elif op in ("exp", "log", "log10", "sqrt"):
            rhs = f"pbr_{op}({a})"
        else:
            rhs = f"{op}({a})"
        names[i] = f"t{i}"
        lines.append(f"    const double t{i} = {rhs};")


class _NoFloatForm(Exception):
    pass


def _flit(v):
    if v != 0.0 and not (1.2e-38 < abs(float(v)) < 3.4e38):
        raise _NoFloatForm(v)                  # not a normal single-precision number: the model keeps double partials
    s = repr(float(v))
    return (s if ("e" in s or "." in s or "inf" in s or "nan" in s) else s + ".0") + "f"


def _emit_body_mixed(g, primal, needed, lines, names, ktab):
    """Like _emit_body, but only the nodes in ``primal`` (what the rate VALUES need) are evaluated in double; every other node --
    the partial derivatives -- is single precision (``const float f<i>``), reading the double values it needs through one cast
    each.  The partials only build the Newton matrix I - cJ (they steer the iteration, the residual stays FP64), so 24 bits are
    enough (the contraction of the inexact Newton iteration is ~cond(I - cJ)*6e-8), they halve the registers the
    Jacobian holds and they run on the FP32 pipe next to the FP64 one."""
    live, stack = set(), list(needed)
    while stack:
        i = stack.pop()
        if i in live:
            continue
        live.add(i)
        stack.extend(_children(g.nodes[i]))
    fnames = {}          # node -> name of its float value

    def fl(j):
        """float operand for node j"""
        if j in fnames:
            return fnames[j]
        n = g.nodes[j]
        if n[0] == "const":
            fnames[j] = _flit(n[1]) if n[1] >= 0 else "(" + _flit(n[1]) + ")"
        else:                                   # a double value (primal node or input): cast once
            lines.append(f"    const float g{j} = (float)({names[j]});")
            fnames[j] = f"g{j}"
        return fnames[j]

    _emit_body(g, [i for i in live if i in primal], lines, names, ktab)
    for i in sorted(live):
        if i in primal or i in names:
            continue
        n = g.nodes[i]
        op = n[0]
        if op == "const":
            fnames[i] = _flit(n[1]) if n[1] >= 0 else "(" + _flit(n[1]) + ")"
            continue
        if op == "in":
            names[i] = {"T": "T", "P": "P"}.get(n[1]) or {"X": "x", "C": "C", "PAR": "par"}[n[1]] + f"[{n[2]}]"
            continue
        a = fl(n[1])
        if op in ("add", "sub", "mul"):
            rhs = f"{a} {dict(add='+', sub='-', mul='*')[op]} {fl(n[2])}"
        elif op == "neg":
            rhs = f"-{a}"
        elif op == "rcp":
            rhs = f"pbr_rcpf({a})"
        elif op == "powi":
            e = n[2]
            rhs = _powi(a, abs(e))
            if e < 0:
                rhs = f"pbr_rcpf({rhs})"
        elif op == "pow":
            rhs = f"pbr_powf({a}, {fl(n[2])})"
        elif op in ("exp", "log", "log10", "sqrt"):
            rhs = f"pbr_{op}f({a})"
        else:
            raise RateCompileError(f"no single-precision form of {op}")
        fnames[i] = f"f{i}"
        lines.append(f"    const float f{i} = {rhs};")
    return fnames, fl


def _powi(a, e):
    # square-and-multiply, fully expanded (e <= 16)
    if e == 1:
        return a
    if e == 2:
        return f"{a} * {a}"
    if e == 3:
        return f"{a} * {a} * {a}"
    h = _powi(a, e//2)
    sq = f"({h}) * ({h})"
    return sq if e % 2 == 0 else f"{sq} * {a}"


def _emit(cr):
    g, outs, nc, nr = cr.g, cr.outs, cr.n_comp, cr.n_rxn
    n0 = len(g.nodes)
    derivs = _derivatives(g, outs, nc)
    src = []
    src.append("// generated by pyremot_b200.ratec — do not edit")
    src.append(f"// {n0} primal nodes, {len(g.nodes)} with derivatives; params: {cr.params}")
    src.append(f"#define PBR_NP {len(cr.params)}")
    sig = "double T, double P, const double* x, const double* C, const double* par"
    ktab = _ConstTable()
    head = len(src)
    # values only
    lines, names = [], {}
    _emit_body(g, outs, lines, names, ktab)
    src.append(f"PBR_DEV void pbr_rates({sig}, double* r) {{")
    src.extend(lines)
    for j, o in enumerate(outs):
        src.append(f"    r[{j}] = {names[o]};")
    src.append("}")
    # values + partials
    lines, names = [], {}
    need = list(outs)
    for dj in derivs:
        need.extend(v for k, v in dj.items() if k[0] != "PAR")
    _emit_body(g, need, lines, names, ktab)
    src.append(f"PBR_DEV void pbr_rates_jac({sig}, double* r, double* drT, double* drP, "
               f"double (*drx)[PBR_NC], double (*drC)[PBR_NC]) {{")
    src.extend(lines)
    for j, o in enumerate(outs):
        src.append(f"    r[{j}] = {names[o]};")
        dj = derivs[j]
        src.append(f"    drT[{j}] = {names[dj[('T', 0)]] if ('T', 0) in dj else '0.0'};")
        src.append(f"    drP[{j}] = {names[dj[('P', 0)]] if ('P', 0) in dj else '0.0'};")
        for i in range(nc):
            src.append(f"    drx[{j}][{i}] = {names[dj[('X', i)]] if ('X', i) in dj else '0.0'};")
            src.append(f"    drC[{j}][{i}] = {names[dj[('C', i)]] if ('C', i) in dj else '0.0'};")
    src.append("}")
    # values in double, partials in single precision (PBR_JAC_F32)
    primal, stack = set(), list(outs)
    while stack:
        i = stack.pop()
        if i not in primal:
            primal.add(i)
            stack.extend(_children(g.nodes[i]))
    try:
        lines, names = [], {}
        fnames, fl = _emit_body_mixed(g, primal, need, lines, names, ktab)
        tail = []

        def fout(node):
            if node is None:
                return "0.0f"
            return fl(node)
        for j, o in enumerate(outs):
            tail.append(f"    r[{j}] = {names[o]};")
            dj = derivs[j]
            tail.append(f"    drT[{j}] = {fout(dj.get(('T', 0)))};")
            tail.append(f"    drP[{j}] = {fout(dj.get(('P', 0)))};")
            for i in range(nc):
                tail.append(f"    drx[{j}][{i}] = {fout(dj.get(('X', i)))};")
                tail.append(f"    drC[{j}][{i}] = {fout(dj.get(('C', i)))};")
        src.append("#define PBR_HAVE_JACF 1")
        src.append(f"PBR_DEV void pbr_rates_jacf({sig}, double* r, float* drT, float* drP, "
                   f"float (*drx)[PBR_NC], float (*drC)[PBR_NC]) {{")
        src.extend(lines)
        src.extend(tail)
        src.append("}")
    except _NoFloatForm:
        src.append("// no single-precision form of the partials (a constant is outside the float range): PBR_JAC_F32 falls back to double")
    # values + state partials + partials with respect to the runtime parameters (forward sensitivities, config 5)
    if cr.params:
        lines, names = [], {}
        needp = list(need)
        for dj in derivs:
            needp.extend(v for k, v in dj.items() if k[0] == "PAR")
        _emit_body(g, needp, lines, names, ktab)
        src.append(f"PBR_DEV void pbr_rates_jac_dpar({sig}, double* r, double* drT, double* drP, "
                   f"double (*drx)[PBR_NC], double (*drC)[PBR_NC], double (*drp)[PBR_NP]) {{")
        src.extend(lines)
        for j, o in enumerate(outs):
            src.append(f"    r[{j}] = {names[o]};")
            dj = derivs[j]
            src.append(f"    drT[{j}] = {names[dj[('T', 0)]] if ('T', 0) in dj else '0.0'};")
            src.append(f"    drP[{j}] = {names[dj[('P', 0)]] if ('P', 0) in dj else '0.0'};")
            for i in range(nc):
                src.append(f"    drx[{j}][{i}] = {names[dj[('X', i)]] if ('X', i) in dj else '0.0'};")
                src.append(f"    drC[{j}][{i}] = {names[dj[('C', i)]] if ('C', i) in dj else '0.0'};")
            for k in range(len(cr.params)):
                src.append(f"    drp[{j}][{k}] = {names[dj[('PAR', k)]] if ('PAR', k) in dj else '0.0'};")
        src.append("}")
    src.insert(head, ktab.emit())
    return "\n".join(src) + "\n"
