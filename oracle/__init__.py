"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference algorithm.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import anything from here; the product (``chunkflow_b200``) never does.
"""
