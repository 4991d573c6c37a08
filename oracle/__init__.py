"""CPU parity oracle for the B200 `mom_step!` library — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package (see the header of oracle/wl_oracle.cpp).
"""
from .oracle import OracleSim, lib, build  # noqa: F401
