"""CPU oracle for the short-run Langevin hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or as the
timed CPU baseline -- never as the thing shipped.  The product path
(``lsnf_b200``) never imports this package and fails loudly when its CUDA
library is missing.

Parity status: the reference repository ships no tests, golden vectors or
fixtures, so parity is pinned by running the reference's
own ``model.py`` (imported unmodified from ``/root/reference`` in the build
container) on seeded inputs: ``oracle/make_golden.py`` checks this restatement
against it and writes ``tests/golden/*.npz``; ``tests/test_oracle_*.py`` replay
those fixtures wherever the reference is absent (the GPU box).
"""
