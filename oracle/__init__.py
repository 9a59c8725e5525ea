"""CPU oracle for the EgoEgo stage-2 sampling hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or the
timed CPU baseline.  The product path (``egoego_release_b200``) never imports
this package and fails loudly when its CUDA library is missing.
"""
