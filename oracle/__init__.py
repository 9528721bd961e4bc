"""ORACLE package — test infrastructure only.

CPU restatements of the reference's hot path (block forward/backward, RMSprop, EMA, L2 decay).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package, and only as the checker / CPU baseline — never the product path.
"""
