"""CPU oracle for the temporal-head hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and there only as the checker / the CPU arm.  The
product path (``computervision_codes_b200``) never imports this package and raises if
its CUDA library is missing.

Parity status: PINNED.  The reference ships no tests, golden vectors or
checkpoints, so the restatement here is pinned against outputs
of the reference's own modules executed in the build container
(``oracle/gen_golden.py`` imports ``/root/reference/...`` and writes
``tests/golden/*.npz``); ``tests/test_oracle_golden.py`` checks every oracle
function against those fixtures.  Exception, stated where it applies: the 7-way
"phase" softmax-CE head has no reference implementation at all (the word "phase"
does not occur in the reference) -> parity unpinned, oracle = the textbook
definition.
"""
