"""CPU oracle for the path-simulation + LSM hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product
(``options_model_b200``) never imports it and has no CPU fallback.
"""
