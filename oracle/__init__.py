"""CPU oracle (test infrastructure only) - see oracle/sdz_oracle.h.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package.  The product (sd-zlib_b200/) never does.
"""
from .oracle import *  # noqa: F401,F403
