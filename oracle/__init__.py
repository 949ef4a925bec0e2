"""TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may
import it, and only as the checker.  The product (``sdformerflow_b200``) never imports
this package and has no CPU fallback.
"""
