"""CPU oracle (TEST INFRASTRUCTURE).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package; the product never does."""
from .oracle import OracleSolver, batch_run, build, debug_set_iter_max, load_fixture, solve_extended_precision, solve_trace, FIXTURE_DIR  # noqa: F401
