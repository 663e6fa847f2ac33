#!/usr/bin/env python
"""Multi-GPU check of the batched evaluator (searches shard by rank, results gathered at the end): under torchrun
every rank runs `Evaluator.eval_batched` on its share of the cubes; rank 0 compares the gathered results with a single-rank run
of all cubes.  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/eval_sharded_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rl_rubiks_b200.evaluation import Evaluator  # noqa: E402
from rl_rubiks_b200.frontier import AStarBatch  # noqa: E402


class FakeNet(torch.nn.Module):
	def __init__(self, w):
		super().__init__()
		self.w = torch.from_numpy(w).cuda()

	def forward(self, x, policy=True, value=True):
		return torch.floor((x @ self.w) / 4.0).unsqueeze(1)


def main():
	rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
	torch.cuda.set_device(local)
	if world > 1:
		if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
			os.environ["NCCL_DEBUG"] = "WARN"
		dist.init_process_group("nccl", device_id=torch.device("cuda", local))
	w = np.random.RandomState(11).randint(-6, 7, 480).astype(np.float32)
	ev = Evaluator(n_games=5, scrambling_depths=[1, 2, 3, 4, 6], max_states=4000)
	np.random.seed(9)                                                  # same draws on every rank
	res, states, times = ev.eval_batched(AStarBatch(FakeNet(w), 0.2, 20), shard=True)
	np.random.seed(9)
	res1, states1, _ = ev.eval_batched(AStarBatch(FakeNet(w), 0.2, 20), shard=False)
	ok = bool((res == res1).all() and (states == states1).all())
	if rank == 0:
		print(f"world {world}: sharded evaluation equals the single-rank run: {ok}; solved {(res != -1).sum()}/{res.size}, "
			  f"states {int(states.sum())}, {times.flat[0] * res.size * 1e3:.1f} ms")
	if world > 1:
		dist.destroy_process_group()
	sys.exit(0 if ok else 1)


if __name__ == "__main__":
	main()
