#!/usr/bin/env python
"""Times rb_scramble (2^24 cubes x 100 moves, BASELINE configs[1]) for the current RB_SCRAMBLE_* environment and checks a
subsample against the oracle.  One line: threads, median ms, fraction of the HBM roofline, parity."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rl_rubiks_b200 import _native as N  # noqa: E402
from oracle import cube_oracle as O  # noqa: E402

n, depth = 1 << 24, int(os.environ.get("DEPTH", "100"))
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(3)
acts = torch.randint(0, 12, (n, depth), dtype=torch.uint8, device=dev, generator=g)
out = torch.empty(n, 20, dtype=torch.int8, device=dev)
sh = N.stream_handle()
n = int(os.environ.get("CUBES", str(n)))
acts, out = acts[:n], out[:n]
if os.environ.get("LAYOUT", "cube") == "move":          # the reference's draw shape (depth, n): move-major
	acts_t = acts.t().contiguous()
	run = lambda: N.check(N.lib.rb_scramble(N.REP_2024, N.ptr(acts_t), 1, n, None, N.ptr(out), n, depth, sh))
else:
	run = lambda: N.check(N.lib.rb_scramble(N.REP_2024, N.ptr(acts), depth, 1, None, N.ptr(out), n, depth, sh))
for _ in range(3):
	run()
ms = []
for _ in range(10):
	a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	a.record(); run(); b.record(); torch.cuda.synchronize()
	ms.append(a.elapsed_time(b))
sub = torch.arange(0, n, max(1, n // 4096), device=dev)
f, d = O.indices_to_actions(acts[sub].cpu().numpy())
ok = bool((out[sub].cpu().numpy() == O.scramble_many(f, d, True)).all())
med = float(np.median(ms))
print("layout", os.environ.get("LAYOUT", "cube"), "cubes", n, "threads", os.environ.get("RB_SCRAMBLE_THREADS", "default"), "depth", depth, "ms", round(med, 4), "frac", round(n * (depth + 20) / med / 1e6 / 6499.0, 4), "parity", ok)
