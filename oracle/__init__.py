"""CPU oracle (TEST INFRASTRUCTURE ONLY -- see oracle/oracle.h). Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this package."""
from .binding import *  # noqa: F401,F403
