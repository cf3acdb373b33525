"""CPU oracle for the meta-training hot path -- TEST INFRASTRUCTURE ONLY.

This package is a CPU (torch fp32/fp64, eager) restatement of the reference
algorithm for the path named in BASELINE.json's north_star.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it; the product package never does.

Parity status: PINNED.  ``oracle/gen_golden.py`` imports the unmodified
reference model and loss from ``/root/reference`` (in the build container),
injects the same dropout masks, checks this restatement against it and writes
the reference's own outputs to ``tests/golden/*.npz``; ``tests/test_oracle.py``
re-checks the restatement against those committed fixtures everywhere.
The ``higher`` inner-loop semantics (third-party, absent, unpinned) are
restated from SURVEY.md Appendix C and are marked "parity unpinned" in
``oracle/meta.py``.
"""
