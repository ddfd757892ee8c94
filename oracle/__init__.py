"""CPU oracle for the ICNN inner-loop hot path -- TEST INFRASTRUCTURE ONLY.

Nothing in ``icnn_b200`` (the product) may import from here.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs use this package, and only as the checker / the timed CPU baseline.
"""
