"""ORACLE — test infrastructure only (CPU restatement of the reference's SpMV path + the reference itself).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this package. The product (``spmv_acc_b200``) never does.
"""
from .oracle import *  # noqa: F401,F403
