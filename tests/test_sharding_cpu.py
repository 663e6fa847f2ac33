"""N>1 host logic on CPU: world_size-2 gloo processes shard the units, run the work on their slice (the
oracle stands in for the kernels here: no GPU in this container) and gather / reduce.  The GPU kernels themselves are
covered by the -m gpu tests; bench.py runs the same sharding under NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
	sys.path.insert(0, ROOT)

from rl_rubiks_b200 import sharding as S  # noqa: E402


def test_shard_bounds_partition_every_unit_once():
	for n in (0, 1, 7, 8, 9, 1000, 1 << 24):
		for ws in (1, 2, 3, 4, 8):
			b = [S.shard_bounds(n, ws, r) for r in range(ws)]
			assert b[0][0] == 0 and b[-1][1] == n
			assert all(b[r][1] == b[r + 1][0] for r in range(ws - 1))
			sizes = S.shard_sizes(n, ws)
			assert sum(sizes) == n and max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
	with pytest.raises(IndexError):
		S.shard_bounds(10, 2, 2)
	assert len({S.rank_seed(0, r) for r in range(8)}) == 8


def test_dp_minibatch_bounds_agree_across_ranks():
	"""Data-parallel training: every rank runs the same number of minibatches whatever its shard size, and its slices cover its
	states exactly once, in order."""
	for games, depth, bsize, ws in ((7501, 30, 1000, 2), (7, 30, 100, 2), (3, 5, 4, 8), (1000, 25, 25000, 4), (9, 2, 1, 4)):
		counts = set()
		for r in range(ws):
			lo, hi = S.shard_bounds(games, ws, r)
			own = (hi - lo) * depth
			b = S.dp_minibatch_bounds(own, games, depth, bsize, ws)
			counts.add(len(b))
			assert b[0].start == 0 and b[-1].stop == own and all(b[k].stop == b[k + 1].start for k in range(len(b) - 1))
			assert max(x.stop - x.start for x in b) <= bsize
		assert len(counts) == 1


def test_single_process_is_identity():
	x = torch.arange(12).reshape(6, 2)
	assert S.world() == (0, 1)
	assert S.gather_rows(x) is x
	assert S.reduce_stats({"a": 2.0}) == {"a": 2.0}
	assert (S.take_shard(x) == x).all()


def _free_port():
	with socket.socket() as s:
		s.bind(("127.0.0.1", 0))
		return s.getsockname()[1]


def _worker(rank, ws, port, n, depth, ret):
	os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
	dist.init_process_group("gloo", rank=rank, world_size=ws)
	try:
		from oracle import cube_oracle as O
		g = np.random.RandomState(7)                                   # same draw on every rank = replicated input
		acts = g.randint(0, 12, (n, depth)).astype(np.uint8)

		def scramble(a):
			f, d = O.indices_to_actions(a)
			return O.scramble_many(f, d, True)

		full = S.sharded_apply(scramble, acts)                          # gathered on every rank
		lo, hi = S.shard_bounds(n, ws, rank)
		local = S.sharded_apply(scramble, acts, gather=False)
		stats = S.reduce_stats({"cubes": hi - lo, "moves": (hi - lo) * depth})
		tmax = S.reduce_stats({"ms": 10.0 + rank}, op="max")
		# game axis of the (depth, games) ADI draw shards along axis 1
		draw = g.randint(0, 12, (depth, n)).astype(np.uint8)
		mine = S.take_shard(draw, axis=1)
		ret[rank] = dict(full=full, local=local, lo=lo, hi=hi, stats=stats, tmax=tmax, mine_shape=mine.shape,
						 mine_ok=bool((mine == draw[:, lo:hi]).all()), expect=scramble(acts))
	finally:
		dist.destroy_process_group()


@pytest.mark.parametrize("n", [64, 65])
def test_two_rank_gloo_shard_gather_reduce(n):
	ws, depth = 2, 9
	mgr = mp.Manager()
	ret = mgr.dict()
	mp.spawn(_worker, args=(ws, _free_port(), n, depth, ret), nprocs=ws, join=True)
	assert sorted(ret.keys()) == [0, 1]
	for r in range(ws):
		out = ret[r]
		assert out["full"].shape == (n, 20) and (out["full"] == out["expect"]).all()      # rank order == unit order
		assert (out["local"] == out["expect"][out["lo"]:out["hi"]]).all()
		assert out["stats"] == {"cubes": float(n), "moves": float(n * depth)}
		assert out["tmax"] == {"ms": 11.0}
		assert out["mine_shape"] == (depth, out["hi"] - out["lo"]) and out["mine_ok"]


def _grad_worker(rank, ws, port, ret):
	os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
	dist.init_process_group("gloo", rank=rank, world_size=ws)
	try:
		torch.manual_seed(0)                                            # same replica and same full batch on every rank
		net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ELU(), torch.nn.Linear(5, 1))
		x, y = torch.randn(8, 6), torch.randn(8, 1)
		lo, hi = S.shard_bounds(8, ws, rank)
		torch.nn.functional.mse_loss(net(x[lo:hi]), y[lo:hi]).backward()            # mean over this rank's half
		grads = [p.grad for p in net.parameters()]
		S.allreduce_mean_(grads)
		ret[rank] = [g.clone().numpy() for g in grads]
	finally:
		dist.destroy_process_group()


def test_two_rank_gradient_mean_equals_full_batch_gradient():
	"""Data-parallel training (rl_rubiks_b200.train): the rank-mean of per-shard mean-loss gradients is the full-batch gradient."""
	mgr = mp.Manager()
	ret = mgr.dict()
	mp.spawn(_grad_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
	torch.manual_seed(0)
	net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ELU(), torch.nn.Linear(5, 1))
	x, y = torch.randn(8, 6), torch.randn(8, 1)
	torch.nn.functional.mse_loss(net(x), y).backward()
	want = [p.grad.numpy() for p in net.parameters()]
	for r in range(2):
		for a, b in zip(ret[r], want):
			np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-7)
	S.allreduce_mean_([torch.ones(3)])                                  # no process group: no-op


def _seeded_worker(rank, ws, port, n, depth, seed, ret):
	os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
	dist.init_process_group("gloo", rank=rank, world_size=ws)
	try:
		from oracle import cube_oracle as O

		def scramble(count, depth, seed, first):                         # the oracle's restatement of the seeded kernel's stream
			f, d = O.indices_to_actions(O.seeded_actions(seed, first, count, depth))
			return O.scramble_many(f, d, True) if count else np.zeros((0, 20), np.int8)

		ret[rank] = dict(shard=S.seeded_shard(n), full=S.sharded_scramble_seeded(n, depth, seed, gather=True, scramble=scramble),
						 local=S.sharded_scramble_seeded(n, depth, seed, scramble=scramble))
	finally:
		dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 33])
def test_two_rank_seeded_scramble_is_one_stream(n):
	"""Device-seeded scramble under sharding: cube id = Philox subsequence, so two ranks together produce exactly what one rank
	produces for the whole range (host logic; the stream comes from the oracle's restatement here, from the kernel on the GPU)."""
	from oracle import cube_oracle as O
	depth, seed, ws = 17, 4242, 2
	mgr = mp.Manager()
	ret = mgr.dict()
	mp.spawn(_seeded_worker, args=(ws, _free_port(), n, depth, seed, ret), nprocs=ws, join=True)
	f, d = O.indices_to_actions(O.seeded_actions(seed, 0, n, depth))
	want = O.scramble_many(f, d, True)
	firsts = [ret[r]["shard"] for r in range(ws)]
	assert firsts[0][0] == 0 and firsts[1][0] == firsts[0][1] and firsts[0][1] + firsts[1][1] == n
	for r in range(ws):
		assert (ret[r]["full"] == want).all()
		lo, cnt = firsts[r]
		assert (ret[r]["local"] == want[lo:lo + cnt]).all()
