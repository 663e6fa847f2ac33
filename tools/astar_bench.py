#!/usr/bin/env python
"""BASELINE configs[3]: batched weighted A* on K cubes scrambled `depth` moves deep, lambda = 0.16, N = 700
(configs/main_eval.ini:8-9), max_states = 175000 (runeval.py:43), random-init value MLP of the reference's fc_small shape
(model.py:143-161: 480 -> 4096 -> 2048 -> 512 -> 1, ELU + BatchNorm1d) as a plain torch module.
Reports states/s with the net included and the frontier machinery alone (--cheap-net: a 480 -> 32 -> 1 net; --null-net: the
values are zeros and no GEMM runs at all), the library's kernel launches per step and the CUDA-graph launches per step."""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rl_rubiks_b200 import cube, frontier  # noqa: E402


class ValueMLP(torch.nn.Module):
	def __init__(self, cheap=False):
		super().__init__()
		dims = [480, 32, 1] if cheap else [480, 4096, 2048, 512, 1]
		layers = []
		for a, b in zip(dims[:-2], dims[1:-1]):
			layers += [torch.nn.Linear(a, b), torch.nn.ELU(), torch.nn.BatchNorm1d(b)]
		layers.append(torch.nn.Linear(dims[-2], dims[-1]))
		self.net = torch.nn.Sequential(*layers)

	def forward(self, x, policy=False, value=True):
		out = []
		for i in range(0, x.shape[0], 1 << 20):                  # bounded activation memory
			out.append(self.net(x[i:i + (1 << 20)]))
		return torch.cat(out) if out else x.new_zeros(0, 1)


def main():
	ap = argparse.ArgumentParser()
	ap.add_argument("--cubes", type=int, default=256)
	ap.add_argument("--depth", type=int, default=1000)
	ap.add_argument("--expansions", type=int, default=700)
	ap.add_argument("--max-states", type=int, default=175000)
	ap.add_argument("--lam", type=float, default=0.16)
	ap.add_argument("--cheap-net", action="store_true")
	ap.add_argument("--null-net", action="store_true", help="value net replaced by zeros: one-hot rows + frontier kernels only")
	ap.add_argument("--no-graphs", action="store_true", help="plain launches instead of the two CUDA graphs per step")
	ap.add_argument("--tf32", action="store_true")
	ap.add_argument("--bf16", action="store_true", help="bf16 one-hot rows + bf16 autocast value forward (opt-in)")
	args = ap.parse_args()
	torch.manual_seed(0)
	torch.backends.cuda.matmul.allow_tf32 = args.tf32
	dev = torch.device("cuda", 0)
	net = ValueMLP(args.cheap_net).to(dev).eval()
	if args.null_net:
		class Null(torch.nn.Module):
			def forward(self, x, policy=False, value=True):
				return x.new_zeros(x.shape[0], 1, dtype=torch.float32)
		net = Null()
	g = torch.Generator(device=dev); g.manual_seed(0)
	acts = torch.randint(0, 12, (args.cubes, args.depth), dtype=torch.uint8, device=dev, generator=g)
	starts = cube.scramble_batch(acts)
	agent = frontier.AStarBatch(net, args.lam, args.expansions, oh_dtype=torch.bfloat16 if args.bf16 else torch.float32,
								use_graphs=not args.no_graphs)
	agent.search_many(starts, args.max_states, max_steps=1)           # warm-up: buffer allocation (~25 GB at 1000 cubes), cuBLAS heuristics
	torch.cuda.synchronize()
	t0 = time.perf_counter()
	won, queues, count = agent.search_many(starts, args.max_states)
	torch.cuda.synchronize()
	dt = time.perf_counter() - t0
	per_step = agent.launches / max(agent.steps, 1)
	print(f"[kernels of the library per step {per_step:.1f}; host launches per step: {'2 graphs + one-hot' if not args.no_graphs else f'{per_step:.0f} kernels'} + the net] "
		  f"cubes {args.cubes} depth {args.depth} N {args.expansions} max_states {args.max_states} net {'null' if args.null_net else 'cheap' if args.cheap_net else 'fc_small'}"
		  f"{' tf32' if args.tf32 else ''}{' bf16' if args.bf16 else ''}: steps {agent.steps} solved {int(won.sum())} states {int(count.sum())} in {dt:.3f} s = "
		  f"{count.sum() / dt / 1e6:.2f} M states/s ({count.sum() / max(agent.steps, 1) / args.cubes:.0f} new states/step/cube)")


if __name__ == "__main__":
	main()
