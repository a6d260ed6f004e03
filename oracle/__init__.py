"""TEST INFRASTRUCTURE -- CPU oracles (restatements of the reference algorithm).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import
this package.  The product (``metagym_b200``) never does; it fails loudly when its CUDA library is missing.
"""
