"""Pin the ALGORITHMIC operation count of the reference's RHS formulas: evaluate a scalar, line-by-line
restatement of modelEquationN1 / one node of modelEquationN2 with an op-counting number type, check the values against
the vectorised oracle, and print the counts.  TEST/REPORT TOOL: imports oracle/, never used by the product.

    python tools/count_flops.py

Every +, -, *, /, ** and every exp / log / log10 / sqrt / pow call counts 1; comparisons, max(), indexing and the
multiplications by literal +-1 that the reference spells out (`-1*x`) count as written.
"""
import os
import sys
import types

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pbr_oracle as O  # noqa: E402
from pyremot_b200 import cases  # noqa: E402


class Cnt:
    """float that counts the operations applied to it (shared counter dict)."""
    ops = {}

    def __init__(self, v):
        self.v = float(v.v if isinstance(v, Cnt) else v)

    @staticmethod
    def _tick(k):
        Cnt.ops[k] = Cnt.ops.get(k, 0) + 1

    @staticmethod
    def _val(o):
        return o.v if isinstance(o, Cnt) else float(o)

    def _bin(self, o, k, f):
        Cnt._tick(k)
        return Cnt(f(self.v, Cnt._val(o)))

    def __add__(self, o): return self._bin(o, "add", lambda a, b: a + b)
    def __radd__(self, o): return self._bin(o, "add", lambda a, b: b + a)
    def __sub__(self, o): return self._bin(o, "add", lambda a, b: a - b)
    def __rsub__(self, o): return self._bin(o, "add", lambda a, b: b - a)
    def __mul__(self, o): return self._bin(o, "mul", lambda a, b: a*b)
    def __rmul__(self, o): return self._bin(o, "mul", lambda a, b: b*a)
    def __truediv__(self, o): return self._bin(o, "div", lambda a, b: a/b)
    def __rtruediv__(self, o): return self._bin(o, "div", lambda a, b: b/a)
    def __pow__(self, o): return self._bin(o, "pow", lambda a, b: a**b)
    def __rpow__(self, o): return self._bin(o, "pow", lambda a, b: b**a)
    def __neg__(self): Cnt._tick("add"); return Cnt(-self.v)
    def __float__(self): return self.v
    def __lt__(self, o): return self.v < Cnt._val(o)
    def __gt__(self, o): return self.v > Cnt._val(o)
    def __eq__(self, o): return self.v == Cnt._val(o)
    __hash__ = None


class _CntMath:
    pi, e = np.pi, np.e

    @staticmethod
    def _f(k, f):
        def g(*a):
            Cnt._tick(k)
            return Cnt(f(*[Cnt._val(x) for x in a]))
        return staticmethod(g)


for _k, _f in (("exp", np.exp), ("log", np.log), ("log10", np.log10), ("sqrt", np.sqrt), ("pow", np.power)):
    setattr(_CntMath, _k, _CntMath._f(_k, _f))


def _rebind(f):
    g = dict(f.__globals__)
    g["math"] = _CntMath
    return types.FunctionType(f.__code__, g, f.__name__, f.__defaults__, f.__closure__)


def rates(rate_expr, T, P, x, C):
    """reactionRateExe, rmtReaction.py:11-61"""
    d = {"R_CONST": O.R_CONST, "T": T, "P": P, "MoFri": x, "SpCoi": C}
    for k, v in rate_expr["VARS"].items():
        d[k] = _rebind(v)(d) if isinstance(v, types.FunctionType) else v
    return [_rebind(f)(d) for f in rate_expr["RATES"].values()]


def cp(sym, T):
    """rmtThermo.py:16-49: eval of the Cp string a + b*T + c*T**2 + d*T**3"""
    a, b, c, d = O._DB[sym][1]
    v = a + b*T + c*(T**2)
    if d != 0.0:
        v = v + d*(T**3)
    return v


def cp_mean(sym, T):
    """rmtThermo.py:52-75; Cp(Tref) is re-evaluated on every call by the reference (a constant: not counted)"""
    return (float(cp(sym, Cnt(O.TREF)).v) + cp(sym, T))*0.50


def thermo(e, x, T, r):
    """Cp_mix (:3246-3249) and sum_j r_j dH_j(T) (:3261-3266, rmtThermo.py:258-312)"""
    saved = dict(Cnt.ops)
    _ = [cp(s, Cnt(O.TREF)) for s in e.comps]            # warm-up of nothing; keep counters exact
    Cnt.ops.clear(); Cnt.ops.update(saved)
    Cpm = [cp_mean(s, T) for s in e.comps]
    Cpmix = sum((x[i]*Cpm[i] for i in range(1, e.n)), x[0]*Cpm[0])
    HeReT = []
    for j, (rea, pro) in enumerate(e.sides):
        l3 = sum((cp_mean(s, T)*c for s, c in rea[1:]), cp_mean(rea[0][0], T)*rea[0][1])
        l7 = sum((cp_mean(s, T)*c for s, c in pro[1:]), cp_mean(pro[0][0], T)*pro[0][1])
        HeReT.append((l7 + l3)*(T - O.TREF) + float(e.dH25[j]))
    Ov = sum((r[j]*HeReT[j] for j in range(1, len(r))), r[0]*HeReT[0])
    return Cpmix, Ov


def rhs_n1(e, y):
    """modelEquationN1, pbHomoReactor.py:3133-3302, scalar (instance 0 of the ensemble)."""
    n = e.n
    g = lambda a: float(a[0])                                   # noqa: E731  per-instance constant
    y = [Cnt(v) for v in y]
    C = [y[i]*g(e.Cmax) for i in range(n)]                                      # :3161
    Ct = sum(C[1:], C[0])                                                       # :3167
    T = (0.0 if e.iso else y[n + 1])*g(e.T0) + g(e.T0)                          # :3170
    P = y[n]*g(e.P0)                                                            # :3173
    x = [C[i]/Ct for i in range(n)]                                             # :3176
    vi = g(e.vi0)*(Ct/g(e.Ct0))*(g(e.P0)/P)                                     # :3180
    vi_s = vi/g(e.vi0)                                                          # :3182
    vs = vi*g(e.BeVoFr)                                                         # :3184
    vs_s = vs/(g(e.vi0)*g(e.BeVoFr))                                            # :3186
    M = sum((x[i]*float(e.MW[i]) for i in range(1, n)), x[0]*float(e.MW[0]))*1e-3   # :3202
    rho = P/((O.R_CONST/M)*T)                                                   # :3206
    rho_s = rho/g(e.rho0)                                                       # :3208
    eps, dp = g(e.BeVoFr), g(e.PaDi)
    ergA = 150*g(e.mu)*vs/(dp**2)                                               # :3214 (dp**2, eps terms: constants)
    ergB = ((1 - eps)**2)/(eps**3)
    ergC = 1.75*rho*(vs**2)/dp
    ergD = (1 - eps)/(eps**3)
    dp_dz = -1*(ergA*ergB + ergC*ergD)/(g(e.P0)/g(e.ReLe))                      # :3219-3220
    r = rates(e.rate_expr, T, P, x, C)                                          # :3226-3229
    S = np.asarray(e.S)
    ri = []
    for i in range(n):                                                          # rmtReaction.py:64-97
        acc = None
        for j in range(len(r)):
            if S[j, i] != 0.0:
                t = r[j]*float(S[j, i])
                acc = t if acc is None else acc + t
        ri.append(acc if acc is not None else Cnt(0.0))
    out = [(1/vs_s)*(ri[i]/g(e.G)) for i in range(n)]                           # :3283, :3288
    out.append(dp_dz)
    if not e.iso:
        Cpmix, Ov = thermo(e, x, T, r)
        Cp_s = Cpmix/g(e.Cpmix0)                                                # :3252
        Qm = 0.0 if g(e.Tm) == 0 else g(e.U)*g(e.a)*(g(e.Tm) - T)               # rmtUtility.py:424-452
        constT1 = 1/(rho_s*(Cp_s*eps)*vi_s)                                     # :3257
        out.append(constT1*((-Ov + Qm)/g(e.GH)))                                # :3298
    return [float(v.v) if isinstance(v, Cnt) else float(v) for v in out]


def rhs_n2_node(e, yc, yb, first, Pz, zNo):
    """one node of modelEquationN2, pbHomoReactor.py:3892-4128 (the pressure march adds 8 ops per node, counted here)."""
    n = e.n
    g = lambda a: float(a[0])                                   # noqa: E731
    dz = 1/(zNo - 1)
    yc = [Cnt(v) for v in yc]
    yb = [Cnt(v) for v in yb]
    c = [v if v > O.EPS_CONST else Cnt(O.EPS_CONST) for v in yc[:n]]           # :3899
    C = [c[i]*g(e.Cmax) for i in range(n)]
    Ct = sum(C[1:], C[0])
    T = (0.0 if e.iso else yc[n])*g(e.T0) + g(e.T0)                             # :3914
    x = [C[i]/Ct for i in range(n)]
    M = sum((x[i]*float(e.MW[i]) for i in range(1, n)), x[0]*float(e.MW[0]))*1e-3
    P = Cnt(Pz)
    rho = P/((O.R_CONST/M)*T)                                                   # :3964
    vs = g(e.vi0)*g(e.BeVoFr)
    eps, dp = g(e.BeVoFr), g(e.PaDi)
    ergA = 150*g(e.mu)*vs/(dp**2)
    ergB = ((1 - eps)**2)/(eps**3)
    ergC = 1.75*rho*(vs**2)/dp
    ergD = (1 - eps)/(eps**3)
    dPdz = -1*(ergA*ergB + ergC*ergD)                                           # :3974
    Pnext = dPdz*dz + P                                                         # :3979
    rho_s = rho/g(e.rho0)
    r = rates(e.rate_expr, T, P, x, C)
    S = np.asarray(e.S)
    ri = []
    for i in range(n):
        acc = None
        for j in range(len(r)):
            if S[j, i] != 0.0:
                t = r[j]*float(S[j, i])
                acc = t if acc is None else acc + t
        ri.append(acc if acc is not None else Cnt(0.0))
    F1 = 1/(eps*(g(e.ReLe)/g(e.vs0)))
    out = []
    for i in range(n):
        cb = float(e.c0[0][i]) if first else (yb[i] if yb[i] > O.EPS_CONST else Cnt(O.EPS_CONST))   # :4088-4093
        out.append(F1*(-1*((yc[i] - cb)/dz) + ri[i]/g(e.G)))                    # :4095-4098
    if not e.iso:
        Cpmix, Ov = thermo(e, x, T, r)
        Cp_s = Cpmix/g(e.Cpmix0)
        Qm = 0.0 if g(e.Tm) == 0 else g(e.U)*g(e.a)*(g(e.Tm) - T)
        thb = 0.0 if first else yb[n]
        conv = -1*rho_s*(Cp_s*eps)*((yc[n] - thb)/dz)                           # :4116
        T2 = 1/(rho_s*(Cp_s*eps)*(g(e.ReLe)/g(e.vs0)))                          # :4077 (vi_s == 1)
        out.append(T2*(conv + (-Ov + Qm)/g(e.GH)))                              # :4118-4125
    return [float(v.v) if isinstance(v, Cnt) else float(v) for v in out], float(Pnext.v)


def main():
    res = {}
    for tag, mi in (("N1 non-iso", cases.readme_case("N1")), ("N1 iso", cases.readme_case("N1", "iso-thermal"))):
        e = O.Ensemble(mi)
        y0 = e.iv_n1()[0]*np.array([0.9, 1.0, 30.0, 1.1, 40.0, 50.0, 0.999] + ([] if e.iso else [0.05]))
        Cnt.ops.clear()
        f = rhs_n1(e, y0)
        ref = e.rhs_n1(y0[None])[0]
        err = np.max(np.abs(np.array(f) - ref)/np.abs(ref))
        assert err < 1e-12, err
        res[tag] = dict(Cnt.ops)
        print(f"{tag:12s} total {sum(Cnt.ops.values()):4d}  {dict(sorted(Cnt.ops.items()))}  (max rel diff vs oracle {err:.1e})")
    mi = cases.readme_case("N2")
    e = O.Ensemble(mi)
    zNo = 12
    Y = e.iv_n2(zNo)[0].copy()
    Y[:e.n] *= np.linspace(0.9, 1.1, zNo)[None, :]
    Y[e.n] = np.linspace(0, 0.1, zNo)
    ref = e.rhs_n2(Y[None])[0]
    Cnt.ops.clear()
    Pz, worst = float(e.P0[0]), 0.0
    for z in range(zNo):
        f, Pz = rhs_n2_node(e, Y[:, z], Y[:, z - 1] if z else Y[:, 0]*0, z == 0, Pz, zNo)
        worst = max(worst, np.max(np.abs(np.array(f) - ref[:, z])/np.abs(ref[:, z])))
    assert worst < 1e-11, worst
    per = {k: v/zNo for k, v in Cnt.ops.items()}
    print(f"{'N2 node':12s} total {sum(per.values()):6.1f}  { {k: round(v, 1) for k, v in sorted(per.items())} }  (max rel diff vs oracle {worst:.1e})")
    return res


if __name__ == "__main__":
    main()
