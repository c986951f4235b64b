"""TEST INFRASTRUCTURE ONLY — CPU/PyTorch-fp32 oracle for the ViNet/AViNet hot path.

Nothing under ``oracle/`` is part of the product path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker (or as the timed CPU baseline).

Parity pinning: the reference ships no tests, golden vectors or fixtures for this path
(SURVEY.md §8c).  The oracle is therefore pinned against *outputs of the reference itself*:
``oracle/make_golden.py`` imports the unmodified reference from ``/root/reference`` (only in
the build container), checks this restatement against it, and writes the vectors under
``tests/golden/``; ``tests/test_oracle_golden.py`` re-checks the restatement against those
committed vectors everywhere else (the GPU box has no ``/root/reference``).
"""
