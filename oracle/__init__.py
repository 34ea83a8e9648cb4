"""CPU oracle for the B200 engine — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package; ``giql_b200`` never does.
"""
