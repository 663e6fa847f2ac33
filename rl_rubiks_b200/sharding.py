"""
Multi-GPU sharding of the cube hot path: one process per GPU, every rank owns a contiguous slice of the
independent units (cubes, games, searches) and runs the single-GPU kernels on it.  There is NO collective on the data
path; `torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is used only after the kernels, for the optional
gather of results to every rank and for reducing counters / timings.

The reference has no distributed code at all (single process, `dev/hpc_job.sh:3` asks for one GPU); the unit
boundaries below are the reference's own batch axes: cubes of `multi_rotate`/`scramble` (cube.py:49-52, 206-216),
games of `sequence_scrambler` / `Train.ADI_traindata` (cube.py:218-234, train.py:277), cubes of `Evaluator.eval`
(evaluation.py:45-94).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def world() -> tuple[int, int]:
	"""(rank, world_size); (0, 1) when no process group is initialised."""
	if dist.is_available() and dist.is_initialized():
		return dist.get_rank(), dist.get_world_size()
	return 0, 1


def shard_bounds(n: int, world_size: int, rank: int) -> tuple[int, int]:
	"""Contiguous slice [lo, hi) of n units for `rank`: sizes differ by at most one, lower ranks take the remainder."""
	if not 0 <= rank < world_size:
		raise IndexError(f"rank {rank} outside world of {world_size}")
	base, rem = divmod(int(n), world_size)
	lo = rank * base + min(rank, rem)
	return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(n: int, world_size: int) -> list[int]:
	return [hi - lo for lo, hi in (shard_bounds(n, world_size, r) for r in range(world_size))]


def rank_seed(seed: int, rank: int) -> int:
	"""Per-rank seed for on-rank random draws (distinct streams, reproducible for a given world size)."""
	return (int(seed) * 1_000_003 + 7919 * (rank + 1)) % (2 ** 31 - 1)


def take_shard(x, axis: int = 0):
	"""This rank's slice of a host/device array along `axis` (the unit axis)."""
	rank, ws = world()
	lo, hi = shard_bounds(x.shape[axis], ws, rank)
	idx = [slice(None)] * x.ndim
	idx[axis] = slice(lo, hi)
	return x[tuple(idx)]


def gather_rows(local: torch.Tensor, total: int | None = None) -> torch.Tensor:
	"""Concatenation over ranks of per-rank row blocks (rank order = unit order), on every rank.  Shards may differ in
	size by one row (shard_bounds), so blocks are padded to the largest and trimmed after one all_gather."""
	rank, ws = world()
	if ws == 1:
		return local
	n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
	sizes = [torch.zeros_like(n_local) for _ in range(ws)]
	dist.all_gather(sizes, n_local)
	sizes = [int(s.item()) for s in sizes]
	if total is not None and sum(sizes) != total:
		raise RuntimeError(f"shards hold {sum(sizes)} rows, expected {total}")
	m = max(sizes)
	pad = local
	if local.shape[0] < m:
		pad = torch.cat([local, local.new_zeros((m - local.shape[0],) + tuple(local.shape[1:]))])
	out = local.new_empty((ws * m,) + tuple(local.shape[1:]))
	dist.all_gather_into_tensor(out, pad.contiguous())
	if all(s == m for s in sizes):
		return out
	return torch.cat([out[r * m:r * m + s] for r, s in enumerate(sizes)])


def reduce_stats(stats: dict, op: str = "sum", device=None) -> dict:
	"""All-reduce a flat dict of numbers (counters: 'sum'; timings: 'max') in one collective."""
	rank, ws = world()
	if ws == 1:
		return dict(stats)
	keys = sorted(stats)
	t = torch.tensor([float(stats[k]) for k in keys], dtype=torch.float64, device=device)
	dist.all_reduce(t, op={"sum": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN}[op])
	return {k: float(v) for k, v in zip(keys, t.tolist())}


def allreduce_mean_(tensors) -> None:
	"""In-place mean over ranks of a list of same-dtype tensors (the gradients of one minibatch in data-parallel training,
	rl_rubiks_b200.train) with ONE collective: flatten, all-reduce(SUM), scale, scatter back.  No-op without a process group."""
	rank, ws = world()
	tensors = [t for t in tensors if t is not None]
	if ws == 1 or not tensors:
		return
	flat = torch.cat([t.reshape(-1) for t in tensors])
	dist.all_reduce(flat, op=dist.ReduceOp.SUM)
	flat.div_(ws)
	o = 0
	for t in tensors:
		n = t.numel()
		t.copy_(flat[o:o + n].view_as(t))
		o += n


def sharded_apply(fn, *arrays, axis: int = 0, gather: bool = True, device=None):
	"""Runs `fn` on this rank's slice of every array (sliced along `axis`) and, if `gather`, returns the rank-ordered
	concatenation of the per-rank results on every rank.  `fn` returns one array/tensor whose rows are units.
	Example (8 GPUs):  states = sharded_apply(cube.scramble_batch, actions)   # actions (n, depth) on every rank"""
	parts = [take_shard(a, axis) for a in arrays]
	res = fn(*parts)
	if not gather:
		return res
	was_np = isinstance(res, np.ndarray)
	t = torch.from_numpy(np.ascontiguousarray(res)) if was_np else res
	if device is not None:
		t = t.to(device)
	total = arrays[0].shape[axis] if t.shape[0] == parts[0].shape[axis] else None
	out = gather_rows(t, total)
	return out.cpu().numpy() if was_np else out


def sharded_scramble(actions, gather: bool = True):
	"""`cube.scramble_batch` over all ranks: actions (n, depth) replicated or host-resident on every rank; rank r
	scrambles cubes [lo_r, hi_r).  With gather=False each rank keeps its own (hi-lo, *shape) states on its GPU."""
	from . import cube
	dev = torch.device("cuda", torch.cuda.current_device())
	return sharded_apply(cube.scramble_batch, actions, gather=gather, device=dev)


def dp_minibatch_bounds(own_states: int, games: int, depth: int, batch_size: int, world_size: int) -> list:
	"""Minibatch slices of one rank's share of a rollout in data-parallel training.  Every rank must run the same number of
	minibatches (one gradient all-reduce each) although shards may differ by one game: the count comes from the largest
	shard (ceil(games / world) * depth states at `batch_size` per minibatch), the rank's own states are cut into that many
	near-equal contiguous slices (possibly empty ones on a rank with fewer states than minibatches)."""
	largest = -(-int(games) // world_size) * int(depth)
	nb = max(1, -(-largest // int(batch_size)))
	edges = [own_states * k // nb for k in range(nb + 1)]
	return [slice(edges[k], edges[k + 1]) for k in range(nb)]


def seeded_shard(n_total: int) -> tuple[int, int]:
	"""(first_cube, count) of this rank's share of a device-seeded scramble of `n_total` cubes.  The cube id is the Philox
	subsequence (`rb_scramble_seeded`), so the union of the ranks' results is the single-GPU result whatever the world size."""
	rank, ws = world()
	lo, hi = shard_bounds(n_total, ws, rank)
	return lo, hi - lo


def sharded_scramble_seeded(n_total: int, depth: int, seed: int, gather: bool = False, scramble=None):
	"""`cube.scramble_seeded` over all ranks: rank r draws and scrambles cubes [lo_r, hi_r) of ONE stream; no collective unless
	`gather`.  `scramble(count, depth, seed, first_cube)` defaults to the CUDA path (the CPU tests pass a stand-in)."""
	if scramble is None:
		from . import cube
		scramble = cube.scramble_seeded
	first, count = seeded_shard(n_total)
	local = scramble(count, depth, seed, first)
	if not gather:
		return local
	was_np = isinstance(local, np.ndarray)
	t = torch.from_numpy(np.ascontiguousarray(local)) if was_np else local
	out = gather_rows(t, n_total)
	return out.cpu().numpy() if was_np else out


def sharded_adi_generator(games: int, depth: int, reward_method: str = "lapanfix", **kw):
	"""ADIGenerator for this rank's share of `games` (each rank then draws its own actions with rank_seed, runs its own
	net replica and keeps its batch shard for data-parallel training)."""
	from . import adi
	rank, ws = world()
	lo, hi = shard_bounds(games, ws, rank)
	return adi.ADIGenerator(hi - lo, depth, reward_method, **kw)
