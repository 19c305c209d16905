"""The algorithmic operation counts quoted in DESIGN.md §4 / bench.py's roofline are pinned here: a scalar
restatement of the reference RHS formulas evaluated with an op-counting number type, value-checked against the oracle."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


def test_reference_formula_op_counts_are_pinned(capsys):
    import count_flops
    res = count_flops.main()                      # asserts value agreement with the oracle internally
    out = capsys.readouterr().out
    assert sum(res["N1 non-iso"].values()) == 491 and sum(res["N1 iso"].values()) == 181
    assert res["N1 non-iso"]["exp"] == 8 and res["N1 non-iso"]["log"] == 2 and res["N1 non-iso"]["sqrt"] == 1
    assert "N2 node      total  505.0" in out


def test_compiled_rate_dag_is_smaller_than_the_formulas_as_written():
    """The device evaluates the same function after common-subexpression elimination: bench.py's roofline counts THAT
    (conservative), 295 ops per N1 RHS evaluation instead of the 491 the reference spells out.  (Round 2: the DAG rewrites
    log10 -> log * const, 1/exp(u) -> exp(-u), 1/(c u) -> (1/c)(1/u) trade reciprocals for multiplications: 117 / 384 ops, of which 5
    reciprocals instead of 11; without them (PBR_RATEC_OPT=0) 116 / 393.)"""
    from pyremot_b200 import cases, engine
    m = engine.Model(cases.readme_case("N1"), device=None)
    ops_r, ops_rj = m.rates.op_counts()
    assert (ops_r, ops_rj) == (117, 384)
    assert ops_r + 16*m.n + 14*m.n_rxn + 40 == 295 < 491
    from pyremot_b200 import ratec
    raw = ratec.compile_rates(cases.readme_case("N1")["reaction-rates"], 6, optimize=False)
    assert raw.op_counts() == (116, 393)
