"""CPU oracle for the MCMC hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker (or
as the timed CPU baseline), never as the thing shipped.  The product path
(``mcmc-for-nested-data_b200/``) never imports this package and fails loudly
when its CUDA library is missing.

Parity pin: the restatement is checked byte-for-byte against sample CSVs and
to <=1e-12 against diagnostics produced by the UNMODIFIED reference run in
the build container (``tests/golden/make_golden.py`` is the generating
script; fixtures live in ``tests/golden/``).  The reference itself ships no
tests or golden vectors.
"""
