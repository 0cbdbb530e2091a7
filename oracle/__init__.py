"""CPU oracle for the GP-transport posterior path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import
anything from this package.  The product (``gaussian_process_transportation_b200``) never does.
"""
