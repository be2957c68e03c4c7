#!/usr/bin/env python
"""CPU rate of the snapshot feature rows: the oracle port (one core) on the golden state corpus.  For scale: the
unmodified Python reference needed 36.6 s for 144 + 240 boards in the build container (tests/golden/generate_golden.py
snapshots), i.e. ~10.5 boards/s."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import orc  # noqa: E402

states = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "random_states.npz"))["states"].view(orc.STATE_DTYPE).reshape(-1)
t0 = time.perf_counter(); n = 0
while time.perf_counter() - t0 < 5.0:
    for s in states:
        orc.snapshot_features(s)
    n += len(states)
print({"oracle_snapshot_boards_per_sec_1core": n / (time.perf_counter() - t0)})
