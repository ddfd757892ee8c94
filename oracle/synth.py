"""Seeded synthetic workloads -- re-exported from icnn_b200/workloads.py (a pure-numpy data
generator shared with bench.py; it contains no algorithm).  TEST INFRASTRUCTURE ONLY."""
from icnn_b200.workloads import CONFIGS, WY_SCALE, make_inputs  # noqa: F401
