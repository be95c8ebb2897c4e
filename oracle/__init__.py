"""CPU oracle (TEST INFRASTRUCTURE — see oracle/oracle.cpp header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package. It mirrors the reference's RenderBackend surface
(util/render_backend.h:12-32) so parity tests read like backend-vs-backend comparisons.
"""
from .oracle import OracleBackend, build_oracle, load_oracle_lib  # noqa: F401
