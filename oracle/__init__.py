"""CPU oracle package -- TEST INFRASTRUCTURE ONLY (see oracle/skyjo_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package.  The product (skyjo_rl_b200) never does.
"""
from .oracle import OracleGame, build_oracle, lib, rollout  # noqa: F401
