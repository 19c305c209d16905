"""CPU oracle for the PyREMOT N1/N2 path — TEST INFRASTRUCTURE, never imported by the product.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package (see ``oracle/pbr_oracle.py`` for the parity status and the reference
lines each function restates).
"""
