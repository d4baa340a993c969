"""CPU oracle for the efficient_kws scoring path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``enhance-cb-whisper_b200/`` imports
this package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and there only
as the checker or as the timed CPU baseline -- never on the product path.

Parity pin: the reference ships no tests, golden vectors or known answers for
this path, so the restatement in ``kws_oracle.py`` is
pinned against outputs of the unmodified reference ``KWSModel.forward``
executed in the build container (``oracle/ref_stub.py`` imports
``/root/reference/src/efficient_kws/model.py`` under three stub modules);
``oracle/make_golden.py`` commits those outputs as ``tests/golden/*.npz``.
"""
