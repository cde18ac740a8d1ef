"""CPU oracle for the MUVO geometric sensor-to-grid hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``muvo_b200/`` may import this package:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and only as the checker or as
the timed CPU baseline -- never as a fallback for the CUDA path.

Parity status: the reference ships no tests, golden vectors or fixtures for
this path, so the oracle is pinned by running the
reference's own functions in the build container on seeded inputs
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``, generated with
numpy 2.3.5 / torch 2.11.0) plus the known-answer vectors of SURVEY.md
Appendix A.  ``tests/test_oracle_vs_reference.py`` re-checks the restatement
against the live reference whenever ``/root/reference`` is present.
"""
from .muvo_oracle import *  # noqa: F401,F403
