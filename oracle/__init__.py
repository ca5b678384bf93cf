"""ORACLE — test infrastructure, not product code.

CPU restatements of the reference's algorithms for the FFWM flow-warping hot
path.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package; nothing
under ``ffwm_b200/`` does (tests/test_layout.py enforces that).
"""
