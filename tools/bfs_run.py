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
for rep in range(2):
	torch.cuda.synchronize()
	t0 = time.perf_counter()
	counts, hs = frontier.bfs_layers(depth, is2024=is2024, capacity=1 << 25 if is2024 else 1 << 22)
	torch.cuda.synchronize()
	dt = time.perf_counter() - t0
	print(f"depth {depth} rep {'2024' if is2024 else '686'}: counts {counts} unique {sum(counts)} children {12 * sum(counts[:-1])} "
		  f"{dt * 1e3:.2f} ms  {12 * sum(counts[:-1]) / dt / 1e9:.2f} G children/s")
	del hs
