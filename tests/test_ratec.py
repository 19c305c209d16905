"""Rate-expression compiler (pyremot_b200/ratec.py) against the reference executor's semantics
(`PyREMOT/docs/rmtReaction.py:11-61`, restated in oracle.rates with the untouched lambdas)."""
import math
import os

import numpy as np
import pytest

from oracle import pbr_oracle as O
from pyremot_b200 import cases, ratec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _state(seed=0, n=6):
    rng = np.random.default_rng(seed)
    x = rng.random(n) + 0.01
    x /= x.sum()
    T, P = 480 + 200*rng.random(), 1e6 + 6e6*rng.random()
    return T, P, x, x*P/(8.314472*T)


def test_traced_dag_equals_lambda_execution():
    mi = cases.readme_case("N1")
    cr = ratec.compile_rates(mi["reaction-rates"], 6)
    assert cr.n_rxn == 3 and cr.params == []
    for seed in range(20):
        T, P, x, C = _state(seed)
        ref = O.rates(mi["reaction-rates"], T, P, x, C, vectorised=False)
        got = np.array(cr.evaluate(T, P, x, C))
        assert np.max(np.abs(got - ref)/np.abs(ref)) < 1e-13     # a/b -> a*rcp(b): <= 1 ulp per division


def test_dashboard_text_front_end_gives_the_same_program():
    lam = ratec.compile_rates(cases.readme_case("N1")["reaction-rates"], 6)
    txt = ratec.compile_rates({"VARS": cases.MEOH_DME_VARS_TEXT, "RATES": cases.MEOH_DME_RATES_TEXT}, 6)
    assert txt.cuda_source == lam.cuda_source and txt.key == lam.key


def test_symbolic_partials_match_finite_differences():
    mi = cases.readme_case("N1")
    cr = ratec.compile_rates(mi["reaction-rates"], 6)
    g = cr.g
    derivs = ratec._derivatives(g, cr.outs, 6)
    T, P, x, C = _state(3)
    vals = ratec._eval_nodes(g, {"T": [T], "P": [P], "X": list(x), "C": list(C), "PAR": []})

    def f(T_, P_, x_):
        return np.array(cr.evaluate(T_, P_, x_, C))
    for j in range(3):
        dT = (f(T*(1+1e-6), P, x)[j] - f(T*(1-1e-6), P, x)[j])/(2e-6*T)
        assert abs(vals[derivs[j][("T", 0)]] - dT) < 1e-6*abs(dT)
        if ("P", 0) in derivs[j]:
            dP = (f(T, P*(1+1e-6), x)[j] - f(T, P*(1-1e-6), x)[j])/(2e-6*P)
            assert abs(vals[derivs[j][("P", 0)]] - dP) < 1e-6*abs(dP)
        for i in range(6):
            if ("X", i) in derivs[j]:
                xp, xm = x.copy(), x.copy()
                xp[i] *= 1+1e-6
                xm[i] *= 1-1e-6
                dx = (f(T, P, xp)[j] - f(T, P, xm)[j])/(2e-6*x[i])
                assert abs(vals[derivs[j][("X", i)]] - dx) < 1e-5*abs(dx) + 1e-12


def test_scalar_vars_become_runtime_parameters():
    mi = cases.readme_case("N1", theta_as_vars=True)
    names = list(cases.MEOH_DME_THETA)
    cr = ratec.compile_rates(mi["reaction-rates"], 6, params=names)
    assert cr.params == names and cr.param_defaults == list(cases.MEOH_DME_THETA.values())
    assert "#define PBR_NP 6" in cr.cuda_source and "par[5]" in cr.cuda_source
    T, P, x, C = _state(1)
    base = np.array(ratec.compile_rates(cases.readme_case("N1")["reaction-rates"], 6).evaluate(T, P, x, C))
    assert np.allclose(cr.evaluate(T, P, x, C, cr.param_defaults), base, rtol=1e-14)
    with pytest.raises(ratec.RateCompileError):
        ratec.compile_rates(mi["reaction-rates"], 6, params=["RT"])       # a lambda, not a scalar


def test_reference_executor_semantics():
    # forward reference -> KeyError (rmtReaction.py:44-51); literal VARS; SpCoi and ** (test_rmt_N2_CH4.py)
    with pytest.raises(KeyError):
        ratec.compile_rates({"VARS": {"a": lambda x: x["b"]*2, "b": 3.0}, "RATES": {"r": lambda x: x["a"]}}, 2)
    cr = ratec.compile_rates(cases.ch4_n2_case()["reaction-rates"], 3)
    assert np.isclose(cr.evaluate(973.0, 3e5, [0.9, 0.05, 0.05], [33.0, 1.8, 1.8])[0], 0.00072*33.0**2, rtol=1e-15)
    # data-dependent branching cannot be compiled: loud error, no silent fallback
    with pytest.raises(ratec.RateCompileError):
        ratec.compile_rates({"VARS": {}, "RATES": {"r": lambda x: 1.0 if x["T"] > 500 else 2.0}}, 2)
    # only the arithmetic grammar is accepted in the text form
    with pytest.raises(ratec.RateCompileError):
        ratec.parse_text('"a": __import__("os").system("true")')
    # general pow, log with base, math constants
    cr = ratec.compile_rates({"VARS": {"n": 1.7}, "RATES": {"r": lambda x: math.pow(x["T"], x["n"]) + math.pi*x["P"]**-0.5 + 2**x["MoFri"][0]}}, 2)
    assert np.isclose(cr.evaluate(500.0, 2e5, [0.3, 0.7], [1, 1])[0], 500.0**1.7 + math.pi*(2e5)**-0.5 + 2**0.3, rtol=1e-14)


def test_emitted_program_does_not_depend_on_the_hash_seed():
    """The cubin cache is keyed by the program text: it must be the same in every process (PYTHONHASHSEED)."""
    import subprocess
    import sys
    code = ("from pyremot_b200 import ratec, cases\n"
            "cr = ratec.compile_rates(cases.readme_case('N1', theta_as_vars=True)['reaction-rates'], 6, params=tuple(cases.MEOH_DME_THETA))\n"
            "print(cr.key)")
    keys = set()
    for seed in ("1", "2", "77"):
        env = dict(os.environ, PYTHONHASHSEED=seed, PYTHONPATH=ROOT)
        keys.add(subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True).stdout.strip())
    assert len(keys) == 1, keys
