#!/usr/bin/env python
"""ADI generation feeding a random-init value/policy net (BASELINE configs[4], second half): rollouts of the
device-resident training loop (rl_rubiks_b200.train.Train) at the configs/main_train.ini shape -- 7500 games x depth 30 per
rollout, fc_small (480 -> 4096 -> 2048 -> {512 -> 12, 512 -> 1}, ELU + BatchNorm1d, model.py:117-161) -- reporting ADI
samples/s for (a) batch generation alone (kernels + value-net forward over the 2.7 M children) and (b) generation + the
minibatch SGD pass, per GPU and aggregated over ranks (each rank generates and trains on its own share of the games;
gradients are averaged with one all-reduce per minibatch, nothing else crosses GPUs).

  python tools/train_bench.py [--games 7500 --depth 30 --rollouts 3 --batch 1000 --tf32]
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/train_bench.py ...
"""
import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rl_rubiks_b200 import adi, sharding  # noqa: E402
from rl_rubiks_b200.train import Train  # noqa: E402


class FcSmall(torch.nn.Module):
	def __init__(self, shared=(480, 4096, 2048), part=(2048, 512)):
		super().__init__()

		def fc(sizes, final):
			layers = []
			for i in range(len(sizes) - 1):
				layers.append(torch.nn.Linear(sizes[i], sizes[i + 1]))
				torch.nn.init.xavier_uniform_(layers[-1].weight)
				if not (final and i == len(sizes) - 2):
					layers += [torch.nn.ELU(), torch.nn.BatchNorm1d(sizes[i + 1])]
			return torch.nn.Sequential(*layers)
		self.shared_net = fc(list(shared), False)
		self.policy_net = fc(list(part) + [12], True)
		self.value_net = fc(list(part) + [1], True)

	def forward(self, x, policy=True, value=True):
		x = self.shared_net(x)
		out = ([self.policy_net(x)] if policy else []) + ([self.value_net(x)] if value else [])
		return out if len(out) > 1 else out[0]


def main():
	ap = argparse.ArgumentParser()
	ap.add_argument("--games", type=int, default=7500)
	ap.add_argument("--depth", type=int, default=30)
	ap.add_argument("--rollouts", type=int, default=3)
	ap.add_argument("--batch", type=int, default=1000)
	ap.add_argument("--ff-batches", type=int, default=4, help="slices of the value-net forward over the children (train.py:249-254)")
	ap.add_argument("--tf32", action="store_true")
	ap.add_argument("--bf16", action="store_true", help="bf16 one-hot batches + bf16 autocast forward (opt-in, not the reference's dtype)")
	args = ap.parse_args()
	oh_dtype = torch.bfloat16 if args.bf16 else torch.float32
	rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
	torch.cuda.set_device(local)
	dev = torch.device("cuda", local)
	if world > 1:
		dist.init_process_group("nccl", device_id=dev)
	torch.backends.cuda.matmul.allow_tf32 = args.tf32
	torch.manual_seed(0)                                                  # same replica on every rank
	np.random.seed(0)                                                     # one global numpy stream; Train derives a per-rank draw stream from it
	net = FcSmall().to(dev)
	lo, hi = sharding.shard_bounds(args.games, world, rank)
	games = hi - lo

	# (a) generation alone: kernels + value-net forward + targets
	g = adi.ADIGenerator(games, args.depth, "lapanfix", oh_dtype=oh_dtype)
	rng = np.random.RandomState(sharding.rank_seed(0, rank))              # own scrambles on every rank
	cast = torch.autocast("cuda", dtype=torch.bfloat16, enabled=args.bf16)
	for _ in range(2):
		with cast:
			adi.adi_traindata(net, games, args.depth, "lapanfix", 0.5, ff_batches=args.ff_batches, generator=g, rng=rng)
	torch.cuda.synchronize()
	t0 = time.perf_counter()
	for _ in range(args.rollouts):
		with cast:
			adi.adi_traindata(net, games, args.depth, "lapanfix", 0.5, ff_batches=args.ff_batches, generator=g, rng=rng)
	torch.cuda.synchronize()
	gen_s = (time.perf_counter() - t0) / args.rollouts
	# kernels only (no net): generate + targets on stale values
	vals = torch.randn(12 * games * args.depth, device=dev)
	a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	a.record()
	for _ in range(args.rollouts):
		g.generate(); g.targets(vals, 0.5)
	b.record(); torch.cuda.synchronize()
	ker_s = a.elapsed_time(b) * 1e-3 / args.rollouts
	del g, vals

	# (b) the whole loop: generation + SGD over the rollout's minibatches
	# Train shards `rollout_games` over the ranks itself when data_parallel is set
	t = Train(rollouts=args.rollouts + 1, batch_size=args.batch, rollout_games=args.games, rollout_depth=args.depth, optim_fn=torch.optim.Adam,
			  alpha_update=0.5, lr=1e-5, gamma=1, update_interval=1, tau=1, reward_method="lapanfix", data_parallel=world > 1, oh_dtype=oh_dtype)
	t.adi_ff_batches = args.ff_batches
	marks = []
	t.log = lambda *_: (torch.cuda.synchronize(), marks.append(time.perf_counter()))
	if world > 1:
		dist.barrier()
	torch.cuda.synchronize()
	t.train(net)
	loop_s = (marks[-1] - marks[0]) / args.rollouts                        # first rollout = warm-up (allocations, cuBLAS heuristics)
	stats = sharding.reduce_stats({"gen": gen_s, "ker": ker_s, "loop": loop_s}, op="max", device=dev)
	if rank == 0:
		n = args.games * args.depth
		print(f"gpus {world} games {args.games} depth {args.depth} ({n} samples, {12 * n} children per rollout) batch {args.batch} "
			  f"{'bf16 one-hot + autocast' if args.bf16 else ('tf32' if args.tf32 else 'fp32')} net fc_small:")
		print(f"  ADI kernels only        : {stats['ker'] * 1e3:9.3f} ms / rollout = {n / stats['ker'] / 1e6:9.2f} M samples/s")
		print(f"  ADI incl. value forward : {stats['gen'] * 1e3:9.3f} ms / rollout = {n / stats['gen'] / 1e6:9.2f} M samples/s")
		print(f"  ADI + SGD (Train.train) : {stats['loop'] * 1e3:9.3f} ms / rollout = {n / stats['loop'] / 1e6:9.2f} M samples/s  "
			  f"final loss {t.train_losses[-1]:.4f}")
	if world > 1:
		dist.destroy_process_group()


if __name__ == "__main__":
	main()
