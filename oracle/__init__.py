"""CPU float64 oracle -- TEST INFRASTRUCTURE ONLY (see oracle/dm_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package.  PARITY UNPINNED.
"""
