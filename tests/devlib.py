"""Test-side alias of the cxb_* device-layer binding (conex_b200/binding.py)."""
import harness  # noqa: F401  (sys.path)
from conex_b200.binding import (dzeros, from_dev, init_communicator, izeros, product, ptr,  # noqa: F401
                                shard_range, to_dev, vp)
