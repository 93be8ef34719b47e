"""CPU oracle for the VelesDB hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs
may import this package.  The product package ``velesdb_b200`` must never do so.
"""
from .oracle import *  # noqa: F401,F403
