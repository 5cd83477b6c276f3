"""CPU oracle for the MinGraph-UNet graph block.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it, and only as the checker or the timed CPU baseline.  The
product path (``mingraph_unet_b200``) never imports this package and raises if
its CUDA library is missing.

Parity status: the reference ships no tests or golden vectors for this path
("parity unpinned" by the reference itself).  The restatement
in ``oracle/restate.py`` is therefore pinned against outputs of the *untouched
reference classes* run in the build container: ``tests/golden/make_golden.py``
imports them from ``/root/reference`` and writes ``tests/golden/*.npz``; the
CPU test-suite checks the restatement against those fixtures (and against the
live reference whenever ``/root/reference`` is mounted).

Two independently written restatements: ``restate.py`` (numpy / torch CPU ops, everything) and
``restate_int.c`` (plain C, the integer / index arithmetic: grid and region edge lists, stable CSR,
argmax labels, nearest un-pool indices; loaded through ``cint.py``); ``tests/test_oracle_c.py``
checks them against each other and against the fixtures.
"""
