"""CPU oracle: NumPy restatement of PyREMOT's pseudo-homogeneous packed-bed path (models N1, N2).

TEST INFRASTRUCTURE ONLY.  Nothing under ``pyremot_b200/`` may import this module; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs do.  It restates, vectorised over reactor instances, what the reference computes one
reactor at a time, and integrates with the same third-party integrator the reference calls
(SciPy ``solve_ivp``; un-pinned dependency of the reference, `setup.py:26-27`; 1.18.1 here).

Parity status: PINNED.  ``oracle/make_golden.py`` runs the *unmodified reference* from
``/root/reference`` (in the build container) and stores its outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every function below against those vectors
(RHS values ≤1e-12 relative, converged solves ≤1e-8), and against the known-answer values
the reference's own property scripts print (`PyREMOT/tests/test_heatCapacity.py`,
`PyREMOT/tests/test_viscosity.py`).

Reference lines followed (all under ``/root/reference/PyREMOT``):
  prolog            docs/pbHomoReactor.py:2744-2862 (N1), :3370-3497 (N2)
  N1 RHS            docs/pbHomoReactor.py:3133-3302
  N2 RHS            docs/pbHomoReactor.py:3827-4131
  rates             docs/rmtReaction.py:11-61 ; formation rates :64-97
  Cp / dH           docs/rmtThermo.py:16-101, 129-198, 258-312 ; ideal gas :315-369
  utilities         docs/rmtUtility.py:57-93, 172-249, 266-277, 405-496
  viscosity/Wilke   docs/gasTransPor.py:137-274 ; data/dataGasViscosity.py:8-141
  property data     data/componentData.py:11-86, 109-424
  constants         core/constants.py:8-23 ; docs/modelSetting.py:11-23
  un-scaling        solvers/solResultAnalysis.py:191-301
  integrator calls  docs/pbHomoReactor.py:2931 (N1), :3589-3685 (N2 interval loop)
"""
from __future__ import annotations

import math
import re
import types

import numpy as np

R_CONST = 8.314472      # core/constants.py:8
EPS_CONST = 1e-30       # core/constants.py:11
TREF = 273.15 + 25.00   # core/constants.py:19-23
PI_CONST = math.pi      # core/constants.py:14

# symbol: (MW g/mol, Cp poly [a,b,c,d] J/mol/K, dHf25 kJ/mol, viscosity spec)
# data/componentData.py:11-22 (MW), :72-86 (dHf25), Cp strings :119.. ; data/dataGasViscosity.py:8-141
# viscosity spec: ("eq1", [A,B,C,D]) or ("eq2", callable) for the DME expression string (:133).
_DB = {
    "CO2":   (44.01, [22.243, 5.98E-02, -3.50E-05, 7.46E-09], -393.51, ("eq1", [4.719875, 0.373279, 512.686300, -6119.961])),
    "H2":    (2.0,   [26.879, 4.35E-03, -3.30E-07, 0.0], 0.0, ("eq1", [0.169104, 0.692485, -7.634394, 467.120])),
    "CH3OH": (32.04, [19.038, 9.15E-02, -1.22E-05, -8.03E-09], -200.7, ("eq1", [0.477915, 0.641076, 284.838034, -3230.713])),
    "H2O":   (18.01, [29.163, 1.45E-02, -2.02E-06, 0.0], -241.820, ("eq1", [0.501246, 0.709247, 869.465599, -90063.891])),
    "CO":    (28.01, [27.113, 6.55E-03, -1.00E-06, 0.0], -110.53, ("eq1", [0.734306, 0.588574, 52.318660, 1018.822])),
    "DME":   (46.07, [19.8, 0.17, -5.66e-5, 0.0], -184.1, ("eq2", lambda T: 2.68e-7*(T**0.3975)/(1+(534/T)))),
    "N2":    (28,    [28.883, -1.57E-03, 8.08E-06, -2.87E-09], 0, ("eq1", [0.847662, 0.574033, 75.437536, 56.771])),
    "CH4":   (16.04, [19.875, 5.021E-02, 1.268E-05, -11.004E-09], -74.90, ("eq1", [1.119178, 0.493234, 214.627200, -3952.087])),
    "C2H4":  (28.05, [3.950, 15.628E-02, -8.339E-05, 17.657E-09], 52.32, ("eq1", [1.503552, 0.456140, 288.342422, 73.362])),
    "C3H6":  (42.08, [3.151, 23.812E-02, -12.176E-05, 24.603E-09], 20.4, ("eq1", [0.876767, 0.520871, 293.618650, -182.857])),
    "C3H8":  (44.1,  [-4.042, 30.456E-02, -15.711E-05, 31.716E-09], -103.9, ("eq1", [0.173966, 0.734798, 143.207060, -7147.859])),
    "C4H10": (58.12, [-7.908, 41.573E-02, -22.992E-05, 49.875E-09], -126.2, ("eq1", [0.075828, 0.837082, 67618677, -2141.762])),
}
COMPONENT_SYMBOLS = tuple(_DB)  # data/componentData.py:435 (order of the payload list)


# ----------------------------------------------------------------------------- property helpers
def cp(comps, T):
    """Cp_i(T) = a + b T + c T^2 + d T^3, J/mol/K. rmtThermo.py:16-49 (eval of componentData Cp strings)."""
    T = np.asarray(T, dtype=np.float64)
    out = []
    for s in comps:
        a, b, c, d = _DB[s][1]
        v = a + b*T + c*(T**2)
        if d != 0.0:
            v = v + d*(T**3)
        out.append(v)
    return np.stack(out, axis=-1)


def cp_mean(comps, T):
    """Endpoint mean 0.5*(Cp(Tref)+Cp(T)). rmtThermo.py:52-75."""
    return (cp(comps, TREF) + cp(comps, T))*0.50


def viscosity(comps, T):
    """mu_i(T) Pa.s. gasTransPor.py:137-223."""
    T = np.asarray(T, dtype=np.float64)
    out = []
    for s in comps:
        kind, spec = _DB[s][3]
        if kind == "eq1":
            A, B, C, D = spec
            out.append(A*1e-6*(T**B)/(1+C*(1/T)+D*(T**-2.0)))
        else:
            out.append(spec(T))
    return np.stack(out, axis=-1)


def wilke(mu, x, MW):
    """Wilke mixing rule. gasTransPor.py:229-274. mu, x: (...,n); MW: (n,)."""
    mu = np.asarray(mu, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    n = len(MW)
    phi = [[None]*n for _ in range(n)]
    for i in range(n):
        for j in range(n):
            if i == j:
                phi[i][j] = np.ones(mu.shape[:-1])
            elif i < j:
                A = 1 + np.sqrt(mu[..., i]/mu[..., j])*((MW[j]/MW[i])**(1/4))
                B = 8*(1+(MW[i]/MW[j]))
                phi[i][j] = A**2/math.sqrt(B)
    for i in range(n):
        for j in range(n):
            if i > j:
                phi[i][j] = (mu[..., i]/mu[..., j])*(MW[j]/MW[i])*phi[j][i]
    tot = 0.0
    for i in range(n):
        den = 0.0
        for j in range(n):
            den = den + x[..., j]*phi[i][j]
        tot = tot + mu[..., i]*x[..., i]/den
    return tot


_TERM = re.compile(r"([0-9.]*)([a-zA-Z0-9.]+)")


def parse_reactions(reactions, comps):
    """Reaction strings -> (S [R,n], dH25 [R] J/mol, sides).

    rmtUtility.py:172-249 (coefficients: reactants negative), rmtReaction.py:64-97 (species not
    in ``comps`` are dropped from S), rmtThermo.py:129-198 (dH25 = 1000*(sum prod - sum react)).
    ``sides`` keeps the (symbol, coeff) lists per side for the dH(T) correction (rmtThermo.py:258-312).
    """
    S, dH25, sides = [], [], []
    for expr in reactions.values():
        lhs, rhs = expr.replace("<", "").replace(">", "").replace(" ", "").split("=")
        rea = [(s, -1.0*float(c) if c else -1.0) for c, s in _TERM.findall(lhs)]
        pro = [(s, float(c) if c else 1.0) for c, s in _TERM.findall(rhs)]
        row = np.zeros(len(comps))
        for s, c in rea + pro:
            if s in comps:
                row[list(comps).index(s)] += c
        S.append(row)
        hp = sum(_DB[s][2]*abs(c) for s, c in pro)
        hr = sum(_DB[s][2]*abs(c) for s, c in rea)
        dH25.append((hp - hr)*1000.00)
        sides.append((rea, pro))
    return np.array(S), np.array(dH25), sides


def dH_correction(sides, T):
    """(sum_k nu_k Cpmean_k(T))*(T-Tref) per reaction, J/mol. rmtThermo.py:258-312."""
    T = np.asarray(T, dtype=np.float64)
    out = []
    for rea, pro in sides:
        l3 = np.dot(cp_mean([s for s, _ in rea], T), np.array([c for _, c in rea]))
        l7 = np.dot(cp_mean([s for s, _ in pro], T), np.array([c for _, c in pro]))
        out.append((l7 + l3)*(T - TREF))
    return np.stack(out, axis=-1)


# ----------------------------------------------------------------------------- rate expressions
class _NpMath:
    """``math`` look-alike over ndarrays, bound into the user's lambdas for vectorised evaluation."""
    exp = staticmethod(np.exp)
    log = staticmethod(np.log)
    log10 = staticmethod(np.log10)
    sqrt = staticmethod(np.sqrt)
    pow = staticmethod(np.power)
    pi = math.pi
    e = math.e


def _rebind(f):
    g = dict(f.__globals__)
    g["math"] = _NpMath
    return types.FunctionType(f.__code__, g, f.__name__, f.__defaults__, f.__closure__)


def rates(rate_expr, T, P, x, C, overrides=None, vectorised=True):
    """r_j(T,P,x,C), rmtReaction.py:11-61: VARS in insertion order (callable -> f(dict), else literal), then RATES.

    T, P: (...,) ; x, C: (...,n).  ``overrides`` maps scalar VARS names to per-instance arrays.
    ``vectorised=False`` loops over instances calling the untouched lambdas with ``math`` (slow, exact).
    """
    T = np.asarray(T, dtype=np.float64)
    P = np.asarray(P, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    C = np.asarray(C, dtype=np.float64)
    VARS, RATES = rate_expr["VARS"], rate_expr["RATES"]
    overrides = overrides or {}
    if not vectorised:
        shp = T.shape
        Tf, Pf = T.reshape(-1), P.reshape(-1)
        xf, Cf = x.reshape(-1, x.shape[-1]), C.reshape(-1, C.shape[-1])
        out = np.zeros((Tf.size, len(RATES)))
        for i in range(Tf.size):
            d = {"R_CONST": R_CONST, "T": float(Tf[i]), "P": float(Pf[i]), "MoFri": xf[i], "SpCoi": Cf[i]}
            for k, v in VARS.items():
                if k in overrides:
                    d[k] = float(np.asarray(overrides[k]).reshape(-1)[i])
                else:
                    d[k] = v(d) if isinstance(v, types.FunctionType) else v
            out[i] = [f(d) for f in RATES.values()]
        return out.reshape(shp + (len(RATES),))
    d = {"R_CONST": R_CONST, "T": T, "P": P,
         "MoFri": [x[..., i] for i in range(x.shape[-1])],
         "SpCoi": [C[..., i] for i in range(C.shape[-1])]}
    for k, v in VARS.items():
        if k in overrides:
            d[k] = np.asarray(overrides[k], dtype=np.float64)
        else:
            d[k] = _rebind(v)(d) if isinstance(v, types.FunctionType) else v
    out = [np.broadcast_to(_rebind(f)(d), T.shape) for f in RATES.values()]
    return np.stack(out, axis=-1)


# ----------------------------------------------------------------------------- ensemble + prolog
SWEEP_KEYS = ("T", "P", "VoFlRa", "conc", "ReInDi", "ReLe", "PaDi", "BeVoFr", "U", "Tm")


class Ensemble:
    """N reactor instances sharing one ``modelInput`` (species, reactions, kinetics, process type).

    ``sweep``: optional dict of per-instance arrays overriding the base inputs: keys of SWEEP_KEYS
    (``conc`` is [N,n]) and ``"VARS:<name>"`` for scalar entries of the VARS dict.
    """

    def __init__(self, modelInput, sweep=None, N=None):
        mi = modelInput
        sweep = dict(sweep or {})
        if N is None:
            N = len(next(iter(sweep.values()))) if sweep else 1
        self.N = N
        self.mi = mi
        self.model = mi["model"]
        self.comps = list(mi["feed"]["components"]["shell"])
        self.n = len(self.comps)
        self.iso = mi["operating-conditions"]["process-type"] == "iso-thermal"  # modelSetting.py:21-23
        self.rate_expr = mi["reaction-rates"]
        self.S, self.dH25, self.sides = parse_reactions(mi["reactions"], self.comps)
        self.R = self.S.shape[0]
        self.MW = np.array([_DB[s][0] for s in self.comps], dtype=np.float64)

        def col(key, base):
            v = sweep.get(key, base)
            return np.broadcast_to(np.asarray(v, dtype=np.float64), (N,)).copy()

        oc, fe, re_, eh = mi["operating-conditions"], mi["feed"], mi["reactor"], mi["external-heat"]
        self.T0 = col("T", oc["temperature"])
        self.P0 = col("P", oc["pressure"])
        self.VoFlRa = col("VoFlRa", fe["volumetric-flowrate"])
        self.C0 = np.broadcast_to(np.asarray(sweep.get("conc", fe["concentration"]), dtype=np.float64),
                                  (N, self.n)).copy()
        self.ReInDi = col("ReInDi", re_["ReInDi"])
        self.ReLe = col("ReLe", re_["ReLe"])
        self.PaDi = col("PaDi", re_["PaDi"])
        self.BeVoFr = col("BeVoFr", re_["BeVoFr"])
        self.U = col("U", eh["OvHeTrCo"])
        self.Tm = col("Tm", eh["MeTe"])
        self.period = oc.get("period")
        # dimensional model M7 only (pbReactor.py:1238, :1545): mixture viscosity and heat-exchange area are INPUTS there
        self.mu_in = col("GaMiVi", fe.get("mixture-viscosity", float("nan")))
        self.a_in = col("EfHeTrAr", eh.get("EfHeTrAr", float("nan")))
        self.v_in = col("SuGaVe0", fe.get("superficial-velocity", float("nan")))      # M8: the inlet velocity is an input (pbReactor.py:1626)
        self.var_overrides = {k[5:]: np.broadcast_to(np.asarray(v, dtype=np.float64), (N,)).copy()
                              for k, v in sweep.items() if k.startswith("VARS:")}
        self.prolog()

    def prolog(self):
        """Derived per-reactor constants. pbHomoReactor.py:2744-2862 / :3370-3497."""
        self.A = PI_CONST*(self.ReInDi**2)/4                    # :2751
        self.Ct0 = np.sum(self.C0, axis=1)                      # :2761
        self.vs0 = self.VoFlRa/self.A                           # :2763
        self.x0 = self.C0/self.Ct0[:, None]                     # :2766
        self.a = 4/self.ReInDi                                  # :2778 (input EfHeTrAr ignored)
        self.mu_i0 = viscosity(self.comps, self.T0)             # :2782
        self.mu = wilke(self.mu_i0, self.x0, self.MW)           # :2783 frozen at feed
        self.Cpm0 = cp_mean(self.comps, self.T0)                # :2787
        self.Cpmix0 = np.sum(self.x0*self.Cpm0, axis=1)         # :2789
        self.M0 = np.sum(self.x0*self.MW, axis=1)*1e-3          # :2793
        self.rho0 = self.M0*self.Ct0                            # :2796
        self.Cmax = np.max(self.C0, axis=1)                     # :2819-2820 ("MAX" setting)
        self.G = (self.vs0/self.ReLe)*self.Cmax                 # :2821 GaMaCoTe0 (same for all i)
        self.GH = self.rho0*self.vs0*self.T0*(self.Cpmix0/self.M0)/self.ReLe  # :2823
        self.vi0 = self.VoFlRa/(self.A*self.BeVoFr)             # :3137
        self.c0 = self.C0/self.Cmax[:, None]                    # :2833

    # ---- initial values
    def iv_n1(self):
        """[C0/Cmax (n), 1, 0]; iso-thermal drops theta. pbHomoReactor.py:2831-2842."""
        cols = [self.c0, np.ones((self.N, 1))]
        if not self.iso:
            cols.append(np.zeros((self.N, 1)))
        return np.concatenate(cols, axis=1)

    def iv_n2(self, zNo):
        """IV2D[(n+1), zNo] = c0 everywhere, theta 0. pbHomoReactor.py:3483-3497. Returns [N, nv, zNo]."""
        nv = self.n if self.iso else self.n + 1
        Y = np.zeros((self.N, nv, zNo))
        Y[:, :self.n, :] = self.c0[:, :, None]
        return Y

    # ---- shared per-state physics (state given as real C [.., n], T, P)
    def _thermo_terms(self, x, T, r):
        Cpm = cp_mean(self.comps, T)                                   # :3246
        Cpmix = np.sum(x*Cpm, axis=-1)                                 # :3249
        HeReT = dH_correction(self.sides, T) + self.dH25               # :3261-3264
        OvHeReT = np.sum(r*HeReT, axis=-1)                             # :3266
        return Cpmix, OvHeReT

    def rhs_n1(self, y):
        """modelEquationN1, pbHomoReactor.py:3133-3302. y: [N, nv] -> [N, nv]."""
        n = self.n
        y = np.asarray(y, dtype=np.float64)
        c = y[:, :n]
        p = y[:, n]
        th = 0.0*p if self.iso else y[:, n+1]                          # :3155
        C = c*self.Cmax[:, None]                                       # :3161
        Ct = np.sum(C, axis=1)                                         # :3167
        T = th*self.T0 + self.T0                                       # :3170
        P = p*self.P0                                                  # :3173
        x = C/Ct[:, None]                                              # :3176
        vi = self.vi0*(Ct/self.Ct0)*(self.P0/P)                        # :3180, rmtUtility.py:405-421
        vi_s = vi/self.vi0                                             # :3182
        vs = vi*self.BeVoFr                                            # :3184
        vs_s = vs/(self.vi0*self.BeVoFr)                               # :3139, :3186
        M = np.sum(x*self.MW, axis=1)*1e-3                             # :3202
        rho = P/((R_CONST/M)*T)                                        # :3206, rmtThermo.py:353-369
        rho_s = rho/self.rho0                                          # :3208
        eps, dp = self.BeVoFr, self.PaDi
        ergA = 150*self.mu*vs/(dp**2)                                  # :3214
        ergB = ((1-eps)**2)/(eps**3)
        ergC = 1.75*rho*(vs**2)/dp
        ergD = (1-eps)/(eps**3)
        dp_dz = -1*(ergA*ergB + ergC*ergD)/(self.P0/self.ReLe)         # :3219-3220
        r = rates(self.rate_expr, T, P, x, C, self.var_overrides)      # :3226-3229
        ri = r @ self.S                                                # :3237, rmtReaction.py:64-97
        out = np.zeros_like(y)
        out[:, :n] = (1/vs_s)[:, None]*(ri/self.G[:, None])            # :3283, :3288
        out[:, n] = dp_dz                                              # :3292
        if not self.iso:
            Cpmix, OvHeReT = self._thermo_terms(x, T, r)
            Cp_s = Cpmix/self.Cpmix0                                   # :3252
            Qm = np.where(self.Tm == 0, 0.0, self.U*self.a*(self.Tm - T))  # rmtUtility.py:424-452
            constT1 = 1/(rho_s*(Cp_s*eps)*vi_s)                        # :3257, :3284
            out[:, n+1] = constT1*((-OvHeReT + Qm)/self.GH)            # :3298
        return out

    # ---- M7: dimensional steady-state model (rmtCore.py:300-309 -> pbReactor.py runM3 :1170-1369)
    def iv_m7(self):
        """[C0 (n), T, P], pbReactor.py:1243-1247."""
        return np.concatenate([self.C0, self.T0[:, None], self.P0[:, None]], axis=1)

    def rhs_m7(self, y):
        """modelEquationM3, pbReactor.py:1371-1577. y = [C_i (n), T, P] dimensional, independent variable z in [0, ReLe] m."""
        n = self.n
        y = np.asarray(y, dtype=np.float64)
        C, T, P = y[:, :n], y[:, n], y[:, n+1]                         # :1448-1452
        Ct = np.sum(C, axis=1)                                         # :1455
        x = C/Ct[:, None]                                              # :1458
        vi = self.vi0*(Ct/self.Ct0)*(self.P0/P)                        # :1462, rmtUtility.py:405-421
        vs = vi*self.BeVoFr                                            # :1464
        MoFl = Ct*vs                                                   # :1468, :1474 (area cancels)
        M = np.sum(x*self.MW, axis=1)*1e-3                             # :1480
        GaDe = M*Ct                                                    # :1483, rmtThermo.py:338-348 (not the EOS density)
        eps, dp = self.BeVoFr, self.PaDi
        ergA = 150*self.mu_in*vs/(dp**2)                               # :1490 (viscosity is an input, :1238)
        ergB = ((1-eps)**2)/(eps**3)
        ergC = 1.75*GaDe*(vs**2)/dp
        ergD = (1-eps)/(eps**3)
        r = rates(self.rate_expr, T, P, x, C, self.var_overrides)      # :1505-1508
        ri = r @ self.S                                                # :1517
        Cpmix, OvHeReT = self._thermo_terms(x, T, r)                   # :1525-1539
        Qm = self.U*self.a_in*(self.Tm - T)                            # :1543-1551 (no Tm == 0 switch, area from the input)
        out = np.zeros_like(y)
        out[:, :n] = ri/vs[:, None]                                    # :1557, :1562
        out[:, n] = (-OvHeReT + Qm)/(MoFl*Cpmix)                       # :1558, :1566
        out[:, n+1] = -1*(ergA*ergB + ergC*ergD)                       # :1494, :1570
        return out

    def unscale_m7(self, y):
        """runM3 result assembly, pbReactor.py:1305-1322: rows x_i (n), T -- the pressure row is dropped. y: [N, n+2, K]."""
        C = y[:, :self.n]
        return np.concatenate([C/np.sum(C, axis=1, keepdims=True), y[:, self.n:self.n+1]], axis=1)

    # ---- M8: dimensional steady-state model with the velocity as an unknown (rmtCore.py:311-320 -> pbReactor.py runM4 :1580-1776)
    def iv_m8(self):
        """[C0 (n), T, P, v0], pbReactor.py:1651-1656 (v0 = feed['superficial-velocity'])."""
        return np.concatenate([self.C0, self.T0[:, None], self.P0[:, None], self.v_in[:, None]], axis=1)

    def rhs_m8(self, y):
        """modelEquationM4, pbReactor.py:1778-1995. y = [C_i (n), T, P, v] dimensional.  The kinetics are NOT the user's rate
        expressions: the reference calls its built-in methanol/DME rates `1000*modelReactions(P, T, MoFri, CaBeDe)` (:1897-1899,
        :8850-8930), which are the README kinetics; `rates_m8` below restates them."""
        n = self.n
        y = np.asarray(y, dtype=np.float64)
        C, T, P, v = y[:, :n], y[:, n], y[:, n+1], y[:, n+2]           # :1857-1864
        Ct = np.sum(C, axis=1)                                         # :1867
        x = C/Ct[:, None]                                              # :1870
        MoFl = Ct*v                                                    # :1873, :1879
        M = np.sum(x*self.MW, axis=1)*1e-3                             # :1885
        GaDe = M*Ct                                                    # :1888
        eps, dp = self.BeVoFr, self.PaDi
        dPdz = -1*(150*self.mu_in*v/(dp**2)*((1-eps)**2)/(eps**3) + 1.75*GaDe*(v**2)/dp*(1-eps)/(eps**3))   # :1893-1897
        r = 1000*self.rates_m8(T, P, x)                                # :1903-1905
        ri = r @ self.S                                                # :1909-1916
        OvR = np.sum(ri, axis=1)                                       # :1919
        Cpmix, OvHeReT = self._thermo_terms(x, T, r)                   # :1923-1936
        Qm = self.U*self.a_in*(self.Tm - T)                            # :1940-1948
        dTdz = (-OvHeReT + Qm)/(MoFl*Cpmix)                            # :1956, :1960
        dvdz = (1/Ct)*((-v/R_CONST)*((1/T)*dPdz - (P/T**2)*dTdz) + OvR)   # :1964-1965
        out = np.zeros_like(y)
        out[:, :n] = (-C*dvdz[:, None] + ri)/v[:, None]                # :1969-1971
        out[:, n] = dTdz
        out[:, n+1] = dPdz
        out[:, n+2] = dvdz
        return out

    def rates_m8(self, T, P, x):
        """PackedBedReactorClass.modelReactions, pbReactor.py:8850-8930 [kmol/m3/s]: the methanol/DME kinetics with CaBeDe from the
        reactor spec; x = mole fractions of (H2, CO2, H2O, CO, CH3OH, DME)."""
        CaBeDe = float(self.mi["reactor"]["CaBeDe"])
        RT = R_CONST*T
        K1 = 35.45*np.exp(-1.7069e4/RT); K2 = 7.3976*np.exp(-2.0436e4/RT); K3 = 8.2894e4*np.exp(-5.2940e4/RT)
        KH2 = 0.249*np.exp(3.4394e4/RT); KCO2 = 1.02e-7*np.exp(6.74e4/RT); KCO = 7.99e-7*np.exp(5.81e4/RT)
        KP1 = np.exp(4213/T - 5.752*np.log(T) - 1.707e-3*T + 2.682e-6*T**2 - 7.232e-10*T**3 + 17.6)
        KP2 = np.power(10.0, 2167/T - 0.5194*np.log10(T) + 1.037e-3*T - 2.331e-7*T**2 - 1.2777)
        KP3 = np.exp(4019/T + 3.707*np.log(T) - 2.783e-3*T + 3.8e-7*T**2 - 6.56e-4/T**3 - 26.64)
        PH2, PCO2, PH2O, PCO, PCH3OH, PDME = [P*x[:, i]*1e-5 for i in range(6)]
        ra2 = 1 + KCO2*PCO2 + KCO*PCO + np.sqrt(KH2*PH2)
        r1 = K1*((PCO2*PH2)/ra2**3)*(1 - (1/KP1)*((PH2O*PCH3OH)/(PCO2*PH2**3)))*CaBeDe
        r2 = K2*(1/ra2)*(PH2O - (1/KP2)*((PCO2*PH2)/PCO))*CaBeDe
        r3 = K3*(PCH3OH**2/PH2O - PDME/KP3)*CaBeDe
        return np.stack([r1, r2, r3], axis=1)

    def unscale_m8(self, y):
        """runM4 result assembly, pbReactor.py:1706-1722: dataYs rows x_i (n), T; the plotted series add P and v."""
        C = y[:, :self.n]
        return np.concatenate([C/np.sum(C, axis=1, keepdims=True), y[:, self.n:]], axis=1)       # [x_i ; T ; P ; v]

    # ---- M9: dynamic dimensional model with marched pressure AND velocity (rmtCore.py:322-331 -> pbReactor.py runM5 :1997-2294)
    def iv_m9(self, zNo):
        """IV2D[(n+1), zNo] = C0 everywhere, T0 (dimensional), pbReactor.py:2063-2075. Returns [N, n+1, zNo]."""
        Y = np.zeros((self.N, self.n + 1, zNo))
        Y[:, :self.n, :] = self.C0[:, :, None]
        Y[:, self.n, :] = self.T0[:, None]
        return Y

    def rhs_m9(self, Y):
        """modelEquationM5, pbReactor.py:2296-2666. Y: [N, n+1, zNo] dimensional [C_i ; T] per node (var-major), t in seconds.
        Pressure and velocity are marched node by node inside the evaluation (:2459, :2535), the rate evaluation of node z sits in
        that chain (:2466-2470); the velocity slope carries the reference's 1000 factors (:2531-2533) and the heat exchange its
        kJ conversion (rmtUtility.py:424-452 with unit 'kJ/m^3.s'); catalyst heat capacity (1-eps) CaDe CaSpHeCa in the energy balance."""
        n = self.n
        Y = np.asarray(Y, dtype=np.float64)
        N, nv, zNo = Y.shape
        eps, dp = self.BeVoFr, self.PaDi
        CaDe = float(self.mi["reactor"]["CaDe"]); CaSpHeCa = float(self.mi["reactor"]["CaSpHeCa"])
        dz = self.ReLe/(zNo - 1)                                       # :2058
        Pz = self.P0.copy()                                            # :2397
        vz = self.v_in.copy()                                          # :2400
        out = np.zeros_like(Y)
        for z in range(zNo):
            Cc = Y[:, :n, z]
            C = np.maximum(Cc, EPS_CONST)                              # :2424
            Ct = np.sum(C, axis=1)
            T = Y[:, n, z]
            x = C/Ct[:, None]
            MoFl = Ct*vz                                               # :2441, :2447
            M = np.sum(x*self.MW, axis=1)*1e-3
            GaDe = M*Ct                                                # :2453
            dPdz = -1*(150*self.mu_in*vz/(dp**2)*((1-eps)**2)/(eps**3) + 1.75*GaDe*(vz**2)/dp*(1-eps)/(eps**3))   # :2456-2461
            r = rates(self.rate_expr, T, Pz, x, C, self.var_overrides)  # :2466-2470
            ri = r @ self.S
            OvR = np.sum(ri, axis=1)                                   # :2480
            Cpmix, OvHeReT = self._thermo_terms(x, T, r)               # :2484-2497
            Qm = self.U*self.a_in*(self.Tm - T)
            Qm = np.where(self.Tm == 0, 0.0, Qm)*1e-3                  # :2508-2509, rmtUtility.py:424-452 ('kJ/m^3.s')
            Tb = self.T0 if z == 0 else Y[:, n, z-1]                   # :2513-2516
            dTdz = (T - Tb)/dz                                         # :2518
            dvdz = (1/(Ct*1000))*((-vz/R_CONST)*((1/T)*dPdz - (Pz/T**2)*dTdz) + OvR*1000)   # :2531-2533
            cT1 = MoFl*Cpmix                                           # :2539
            cT2 = 1/(Ct*Cpmix*eps + (1-eps)*CaDe*CaSpHeCa)             # :2540
            Cb = self.C0 if z == 0 else np.maximum(Y[:, :n, z-1], EPS_CONST)   # :2548-2552
            out[:, :n, z] = (1/eps)[:, None]*(-vz[:, None]*(Cc - Cb)/dz[:, None] - Cc*dvdz[:, None] + ri)   # :2554-2556
            out[:, n, z] = cT2*(-cT1*dTdz + (-OvHeReT + Qm))           # :2571
            vz_next = dvdz*dz + vz                                     # :2535
            Pz = dPdz*dz + Pz                                          # :2462
            vz = vz_next
        return out

    def rhs_n2(self, Y):
        """modelEquationN2, pbHomoReactor.py:3827-4131. Y: [N, nv, zNo] -> same (var-major per instance).

        Vectorised over nodes; only the pressure march P_z[z+1] = P_z[z] + (dP/dz)_z*dz (:3979) is a loop."""
        n = self.n
        Y = np.asarray(Y, dtype=np.float64)
        N, nv, zNo = Y.shape
        dz = 1/(zNo-1)                                                 # :3439 (DoLe = 1)
        eps, dp = self.BeVoFr[:, None], self.PaDi[:, None]
        vs = (self.vi0*self.BeVoFr)[:, None]                           # :3833, v_z constant :3937, :4066
        vs_s = vs/self.vs0[:, None]                                    # v/vf :3923 (== 1 up to rounding)
        vi_s = (vs/eps)/self.vi0[:, None]                              # :3941-3943
        const_F1 = 1/(eps*(self.ReLe/self.vs0)[:, None])               # :4075
        c = np.maximum(Y[:, :n, :], EPS_CONST)                         # :3899       [N, n, zNo]
        C = np.moveaxis(c*self.Cmax[:, None, None], 1, 2)              # :3903       [N, zNo, n]
        Ct = np.sum(C, axis=2)
        th = np.zeros((N, zNo)) if self.iso else Y[:, n, :]            # :3883-3884
        T = th*self.T0[:, None] + self.T0[:, None]                     # :3914
        x = C/Ct[:, :, None]                                           # :3927
        M = np.sum(x*self.MW, axis=2)*1e-3                             # :3960
        ergA = 150*self.mu[:, None]*vs/(dp**2)                         # :3970
        ergB = ((1-eps)**2)/(eps**3)
        ergD = (1-eps)/(eps**3)
        P = np.zeros((N, zNo))
        Pz = self.P0.copy()                                            # :3848
        for z in range(zNo):
            P[:, z] = Pz
            rho_z = Pz/((R_CONST/M[:, z])*T[:, z])                     # :3964
            ergC = 1.75*rho_z*(vs[:, 0]**2)/dp[:, 0]
            dPdz = -1*(ergA[:, 0]*ergB[:, 0] + ergC*ergD[:, 0])        # :3974 (dimensional slope)
            Pz = dPdz*dz + Pz                                          # :3979 (dimensionless dz on a Pa/m slope)
        rho = P/((R_CONST/M)*T)
        rho_s = rho/self.rho0[:, None]
        ov = {k: v[:, None] for k, v in self.var_overrides.items()}
        r = rates(self.rate_expr, T, P, x, C, ov)                      # :3989       [N, zNo, R]
        ri = r @ self.S                                                #             [N, zNo, n]
        cb = np.concatenate([self.c0[:, :, None], c[:, :, :-1]], axis=2)   # :4088-4093
        dCdz = (Y[:, :n, :] - cb)/dz                                   # :4095 (unclamped centre value)
        out = np.zeros_like(Y)
        out[:, :n, :] = const_F1[:, None, :]*(-vs_s[:, None, :]*dCdz + np.moveaxis(ri, 2, 1)/self.G[:, None, None])  # :4098
        if not self.iso:
            Cpm = cp_mean(self.comps, T)
            Cpmix = np.sum(x*Cpm, axis=-1)
            OvHeReT = np.sum(r*(dH_correction(self.sides, T) + self.dH25), axis=-1)
            Cp_s = Cpmix/self.Cpmix0[:, None]
            Tm, U, a = self.Tm[:, None], self.U[:, None], self.a[:, None]
            Qm = np.where(Tm == 0, 0.0, U*a*(Tm - T))
            GH = self.GH[:, None]
            const_T2 = 1/(rho_s*Cp_s*eps*(self.ReLe/self.vs0)[:, None])     # :4077
            thb = np.concatenate([np.zeros((N, 1)), Y[:, n, :-1]], axis=1)   # :4108, :4111
            dTdz = (Y[:, n, :] - thb)/dz
            conv = -1*vi_s*rho_s*(Cp_s*eps)*dTdz                       # :4116
            out[:, n, :] = const_T2*(conv + (1/GH)*(-OvHeReT) + (1/GH)*Qm)  # :4125
        return out

    # ---- un-scaling (solvers/solResultAnalysis.py:191-301; pbHomoReactor.py:2973-2983, :3655-3661)
    def unscale_n1(self, y):
        """y: [N, nv, K] -> dataYs [N, nv, K] = [x (n rows); P; T] (iso: no T row)."""
        n = self.n
        C = y[:, :n]*self.Cmax[:, None, None]
        x = C/np.sum(C, axis=1, keepdims=True)
        rows = [x, (y[:, n]*self.P0[:, None])[:, None, :]]
        if not self.iso:
            rows.append((y[:, n+1]*self.T0[:, None] + self.T0[:, None])[:, None, :])
        return np.concatenate(rows, axis=1)

    def unscale_n2(self, Y):
        """Y: [N, nv, zNo] -> dataYs [N, n+1, zNo] = [x; T] (iso: T row = Tf)."""
        n = self.n
        C = Y[:, :n]*self.Cmax[:, None, None]
        x = C/np.sum(C, axis=1, keepdims=True)
        th = np.zeros((self.N, Y.shape[2])) if self.iso else Y[:, n]
        T = th*self.T0[:, None] + self.T0[:, None]
        return np.concatenate([x, T[:, None, :]], axis=1)

    def sub(self, i):
        """Single-instance view (for per-instance SciPy integration)."""
        return _select(self, [i])


# ----------------------------------------------------------------------------- integration (SciPy)
def solve_n1(ens, method="LSODA", rtol=1e-3, atol=1e-6, zNo=100, instances=None):
    """Per-instance solve_ivp over z in [0,1] with t_eval = linspace(0,1,zNo+1). pbHomoReactor.py:2856-2862, :2931.

    Defaults are SciPy's (what the reference gets: it passes neither rtol nor atol). Returns
    (y [M, nv, zNo+1] scaled state, dataYs [M, nv, zNo+1] un-scaled, nfev [M]).
    """
    from scipy.integrate import solve_ivp
    idx = range(ens.N) if instances is None else instances
    ts = np.linspace(0, 1, zNo+1)
    ys, nfev = [], []
    for i in idx:
        e = ens.sub(i)
        y0 = e.iv_n1()[0]
        sol = solve_ivp(lambda t, y: e.rhs_n1(y[None, :])[0], [0, 1], y0, method=method, t_eval=ts,
                        rtol=rtol, atol=atol)
        if not sol.success:
            raise RuntimeError(f"oracle N1 solve failed for instance {i}: {sol.message}")
        ys.append(sol.y)
        nfev.append(sol.nfev)
    y = np.stack(ys)
    sel = ens if instances is None else _select(ens, list(idx))
    return y, sel.unscale_n1(y), np.array(nfev)


def solve_m7(ens, method="LSODA", rtol=1e-3, atol=1e-6, zNo=30, instances=None):
    """Per-instance solve_ivp over z in [0, ReLe] with t_eval = linspace(0, ReLe, zNo) (zNo = solverSetting['M9']['zNo'] = 30),
    pbReactor.py:1283-1298.  Returns (y [M, n+2, zNo] dimensional state, dataYs [M, n+1, zNo] = [x_i ; T], nfev [M])."""
    from scipy.integrate import solve_ivp
    idx = range(ens.N) if instances is None else instances
    ys, nfev = [], []
    for i in idx:
        e = ens.sub(i)
        L = float(e.ReLe[0])
        sol = solve_ivp(lambda t, y: e.rhs_m7(y[None, :])[0], [0, L], e.iv_m7()[0], method=method, t_eval=np.linspace(0, L, zNo),
                        rtol=rtol, atol=atol)
        if not sol.success:
            raise RuntimeError(f"oracle M7 solve failed for instance {i}: {sol.message}")
        ys.append(sol.y)
        nfev.append(sol.nfev)
    y = np.stack(ys)
    sel = ens if instances is None else _select(ens, list(idx))
    return y, sel.unscale_m7(y), np.array(nfev)


def solve_m8(ens, rtol=1e-3, atol=1e-6, zNo=25, instances=None):
    """runM4: LSODA over z in [0, ReLe], t_eval = linspace(0, ReLe, solverSetting['S3']['timesNo'] = 25), pbReactor.py:1690-1701.
    Returns (y [M, n+3, zNo] dimensional, series [M, n+3, zNo] = [x_i ; T ; P ; v], nfev [M])."""
    from scipy.integrate import solve_ivp
    idx = range(ens.N) if instances is None else instances
    ys, nfev = [], []
    for i in idx:
        e = ens.sub(i)
        L = float(e.ReLe[0])
        sol = solve_ivp(lambda t, y: e.rhs_m8(y[None, :])[0], [0, L], e.iv_m8()[0], method="LSODA", t_eval=np.linspace(0, L, zNo),
                        rtol=rtol, atol=atol)
        if not sol.success:
            raise RuntimeError(f"oracle M8 solve failed for instance {i}: {sol.message}")
        ys.append(sol.y)
        nfev.append(sol.nfev)
    y = np.stack(ys)
    sel = ens if instances is None else _select(ens, list(idx))
    return y, sel.unscale_m8(y), np.array(nfev)


def solve_m9(ens, zNo=100, tNo=10, timesNo=5, method="LSODA", rtol=1e-3, atol=1e-6, instances=None, sparse=False):
    """runM5's interval loop with solver restart, pbReactor.py:2108-2148 (zNo, tNo, timesNo from solverSetting['S2']).
    Returns (Y [M, tNo, n+1, zNo] dimensional snapshots at the interval ends, T [M, tNo, zNo] -- what rmtExe("M9") reports --, nfev)."""
    from scipy.integrate import solve_ivp
    idx = range(ens.N) if instances is None else instances
    out, nfev = [], []
    for i in idx:
        e = ens.sub(i)
        IV = e.iv_m9(zNo)[0].reshape(-1)
        nv = IV.size//zNo
        span = np.linspace(0, float(e.period), tNo + 1)
        snaps, nf = [], 0
        for k in range(tNo):
            t = [span[k], span[k+1]]
            kw = {"jac_sparsity": n2_jac_sparsity(nv, zNo)} if (sparse and method in ("BDF", "Radau")) else {}   # grouping only: the
            # upstream (P, v) coupling missing from the pattern makes the FD Jacobian inexact, which costs Newton iterations, not accuracy
            sol = solve_ivp(lambda t, y: e.rhs_m9(y.reshape(1, nv, zNo))[0].reshape(-1), t, IV, method=method,
                            t_eval=np.linspace(t[0], t[1], timesNo), rtol=rtol, atol=atol, **kw)
            if not sol.success:
                raise RuntimeError(f"oracle M9 solve failed for instance {i}: {sol.message}")
            IV = sol.y[:, -1]
            nf += sol.nfev
            snaps.append(IV.reshape(nv, zNo).copy())
        out.append(np.stack(snaps))
        nfev.append(nf)
    Y = np.stack(out)
    return Y, Y[:, :, -1, :], np.array(nfev)


def n2_jac_sparsity(nv, zNo):
    """Block lower-bidiagonal pattern of dF/dY in the var-major flatten (the ~1e-3 pressure coupling to all upstream
    nodes is left out: with SciPy's BDF a slightly inexact Jacobian only costs Newton iterations, not accuracy)."""
    from scipy.sparse import lil_matrix
    Sp = lil_matrix((nv*zNo, nv*zNo), dtype=np.int8)
    for z in range(zNo):
        for i in range(nv):
            for k in range(nv):
                Sp[i*zNo + z, k*zNo + z] = 1
            if z > 0:
                Sp[i*zNo + z, i*zNo + z - 1] = 1
    return Sp.tocsc()


def solve_n2(ens, zNo=20, tNo=5, timesNo=5, method="LSODA", rtol=1e-3, atol=1e-6, instances=None, sparse=False):
    """Interval loop with solver restart, pbHomoReactor.py:3557-3558, :3589-3685.

    Returns (Y [M, tNo, nv, zNo] scaled snapshots at interval ends, dataYs [M, tNo, n+1, zNo], nfev [M]).
    """
    from scipy.integrate import solve_ivp
    idx = range(ens.N) if instances is None else instances
    out, nfev = [], []
    for i in idx:
        e = ens.sub(i)
        IV = e.iv_n2(zNo)[0].reshape(-1)
        nv = IV.size//zNo
        span = np.linspace(0, float(e.period), tNo+1)
        snaps, nf = [], 0
        for k in range(tNo):
            t = [span[k], span[k+1]]
            kw = {"jac_sparsity": n2_jac_sparsity(nv, zNo)} if sparse else {}
            sol = solve_ivp(lambda t, y: e.rhs_n2(y.reshape(1, nv, zNo))[0].reshape(-1), t, IV, method=method,
                            t_eval=np.linspace(t[0], t[1], timesNo), rtol=rtol, atol=atol, **kw)
            if not sol.success:
                raise RuntimeError(f"oracle N2 solve failed for instance {i}: {sol.message}")
            IV = sol.y[:, -1]
            nf += sol.nfev
            snaps.append(IV.reshape(nv, zNo).copy())
        out.append(np.stack(snaps))
        nfev.append(nf)
    Y = np.stack(out)
    sel = ens if instances is None else _select(ens, list(idx))
    M = Y.shape[0]
    data = np.stack([sel.unscale_n2(Y[:, k]) for k in range(tNo)], axis=1)
    return Y, data, np.array(nfev)


def _select(ens, idx):
    e = object.__new__(Ensemble)
    e.__dict__.update(ens.__dict__)
    e.N = len(idx)
    for k, v in ens.__dict__.items():
        if isinstance(v, np.ndarray) and v.shape[:1] == (ens.N,) and k not in ("S", "dH25", "MW"):
            setattr(e, k, v[idx])
    e.var_overrides = {k: v[idx] for k, v in ens.var_overrides.items()}
    return e


def fd_jacobian(f, y, rel=1e-7):
    """Central-difference Jacobian of f at y (1-D). Test helper for the CUDA analytic Jacobians."""
    y = np.asarray(y, dtype=np.float64)
    f0 = np.asarray(f(y))
    J = np.zeros((f0.size, y.size))
    for k in range(y.size):
        h = rel*max(abs(y[k]), 1e-6)
        yp, ym = y.copy(), y.copy()
        yp[k] += h
        ym[k] -= h
        J[:, k] = (np.asarray(f(yp)) - np.asarray(f(ym)))/(2*h)
    return J
