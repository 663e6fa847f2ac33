#!/usr/bin/env python
"""How fast do N ranks of one box get results back into host memory?  Every rank runs rbh_scramble_seeded (device-side draw, 20 B
per cube come back) into (a) a torch pinned buffer, (b) a registered 4 KB-paged mapping, (c) a registered mapping advised to
transparent huge pages (rbh_host_alloc).  Launch with torchrun; prints the aggregate rate per buffer kind on rank 0."""
import ctypes as C
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rl_rubiks_b200 import _native as N  # noqa: E402

rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
	dist.init_process_group("nccl", device_id=dev)
n, depth, reps = 1 << 24, 100, 5
nbytes = n * 20


def barrier():
	if world > 1:
		dist.barrier()
	torch.cuda.synchronize()


def run(ptr, label):
	N.check(N.lib.rbh_scramble_seeded(N.REP_2024, 1, rank * n, ptr, n, depth))
	barrier()
	t0 = time.perf_counter()
	for _ in range(reps):
		N.check(N.lib.rbh_scramble_seeded(N.REP_2024, 1, rank * n, ptr, n, depth))
	barrier()
	t = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64, device=dev)
	if world > 1:
		dist.all_reduce(t, op=dist.ReduceOp.MAX)
	if rank == 0:
		s = float(t.item())
		print(f"N={world} {label:28s} {s * 1e3:7.2f} ms/step  {world * nbytes / s / 1e9:6.1f} GB/s D2H aggregate  {world * n * depth / s:.3e} moves/s", flush=True)


pinned = torch.empty(n, 20, dtype=torch.int8, pin_memory=True)
run(C.c_void_p(pinned.data_ptr()), "torch pin_memory")
for huge in (0, 1):
	p = N.lib.rbh_host_alloc(nbytes, huge)
	assert p, N.lib.rb_last_error()
	run(C.c_void_p(p), f"rbh_host_alloc(huge={huge})")
	# how much of it really sits on huge pages
	if rank == 0 and huge:
		try:
			thp = [l for l in open("/proc/self/smaps_rollup") if "AnonHugePages" in l]
			print("   ", thp[0].strip(), flush=True)
		except OSError:
			pass
	N.check(N.lib.rbh_host_free(C.c_void_p(p), nbytes))
N.check(N.lib.rbh_release())
if world > 1:
	dist.destroy_process_group()
