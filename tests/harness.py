"""Test-side names: the ctypes binding and the workload generators live in the package
(conex_b200/binding.py, conex_b200/workloads.py), the oracle loader under oracle/."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

from conex_b200.binding import *  # noqa: F401,F403,E402
from conex_b200.binding import PRODUCT_SO, c_double_p, c_int_p  # noqa: F401,E402
from conex_b200.workloads import *  # noqa: F401,F403,E402
from oracle_loader import ORACLE_SO, build_oracle, oracle  # noqa: F401,E402
