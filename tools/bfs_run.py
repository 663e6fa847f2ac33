#!/usr/bin/env python
"""BFS closure from solved to a given depth on the device (BASELINE configs[4]): layer counts, wall time."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rl_rubiks_b200 import frontier  # noqa: E402

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 7
is2024 = (sys.argv[2] != "686") if len(sys.argv) > 2 else True
from rl_rubiks_b200 import _native as N  # noqa: E402
for rep in range(3):
	torch.cuda.synchronize()
	l0 = N.lib.rb_launch_count()
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	t0 = time.perf_counter()
	e0.record()
	counts, hs = frontier.bfs_layers(depth, is2024=is2024, capacity=int(os.environ.get("BFS_CAP", str(1 << 25 if is2024 else 1 << 22))))
	e1.record()
	torch.cuda.synchronize()
	dt = time.perf_counter() - t0
	print(f"  launches {N.lib.rb_launch_count() - l0}, device span {e0.elapsed_time(e1):.2f} ms, per layer ms {[round(x, 3) for x in frontier.LAST_LAYER_MS]}")
	print(f"depth {depth} rep {'2024' if is2024 else '686'}: counts {counts} unique {sum(counts)} children {12 * sum(counts[:-1])} "
		  f"{dt * 1e3:.2f} ms  {12 * sum(counts[:-1]) / dt / 1e9:.2f} G children/s")
	del hs
