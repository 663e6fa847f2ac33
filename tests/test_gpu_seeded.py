"""GPU parity tests of the device-seeded and packed-action scrambles: the moves the seeded kernel draws (Philox4x32-10 on the
device, written out by rb_seeded_actions) equal the numpy restatement of the stream, and the states it produces equal the CPU
oracle replaying exactly those moves (cube.py:206-211: sequential rotate from solved)."""
import numpy as np
import pytest
import torch

from oracle import cube_oracle as O

pytestmark = pytest.mark.gpu

REPS = [pytest.param(True, id="2024"), pytest.param(False, id="686")]


@pytest.fixture(autouse=True)
def _repr_guard():
	from rl_rubiks_b200 import cube
	cube.set_is2024(True)
	yield
	cube.set_is2024(True)


@pytest.mark.parametrize("depth", [0, 1, 2, 3, 4, 5, 11, 12, 13, 20, 23, 24, 25, 26, 47, 48, 49, 100, 333, 1000])
def test_seeded_actions_equal_numpy_philox(depth):
	from rl_rubiks_b200 import cube
	for n, seed, first in ((1, 0, 0), (77, 1234, 5), (1025, 2 ** 63 + 12345, 2 ** 32 - 7), (33, 2 ** 64 - 1, 2 ** 40 + 3)):
		got = cube.seeded_actions(n, depth, seed, first).cpu().numpy()
		assert got.shape == (n, depth) and (got == O.seeded_actions(seed, first, n, depth)).all(), (depth, n)


@pytest.mark.parametrize("is2024", REPS)
@pytest.mark.parametrize("n", [1, 33, 777, 4097])
def test_seeded_scramble_equals_oracle_on_the_dumped_actions(is2024, n):
	from rl_rubiks_b200 import cube
	cube.set_is2024(is2024)
	depths = [0, 1, 2, 3, 4, 5, 6, 7, 11, 12, 13, 14, 23, 24, 25, 26, 35, 36, 37, 38, 47, 48, 49, 50, 97, 98, 99, 100, 101, 250, 1000]
	for depth in depths if is2024 or n <= 777 else depths[::3]:
		seed, first = 1000 * depth + n, 3 * n
		acts = cube.seeded_actions(n, depth, seed, first).cpu().numpy()
		f, d = O.indices_to_actions(acts)
		want = O.scramble_many(f, d, is2024) if depth else np.repeat(O.solved(is2024)[None], n, 0)
		got = cube.scramble_seeded(n, depth, seed, first)
		assert got.is_cuda and (got.cpu().numpy() == want).all(), depth
		# the same cubes through the ordinary host-supplied path
		if depth:
			assert (cube.scramble_batch(acts) == want).all()


def test_seeded_scramble_is_sharding_invariant_and_prefix_stable():
	from rl_rubiks_b200 import cube
	n, depth, seed = 5000, 100, 42
	whole = cube.scramble_seeded(n, depth, seed, 10)
	parts = torch.cat([cube.scramble_seeded(1234, depth, seed, 10), cube.scramble_seeded(n - 1234, depth, seed, 10 + 1234)])
	assert torch.equal(whole, parts)
	# a shorter scramble is a prefix of a longer one with the same seed
	a100, a37 = cube.seeded_actions(64, 100, seed, 0), cube.seeded_actions(64, 37, seed, 0)
	assert torch.equal(a100[:, :37], a37)
	# start states: the sequence is applied to them
	start = cube.scramble_seeded(n, 30, 7)
	both = cube.scramble_seeded(n, depth, seed, 10, start=start)
	acts = cube.seeded_actions(n, depth, seed, 10)
	assert torch.equal(both, cube.scramble_batch(acts, start=start))


def test_seeded_moves_are_uniform():
	"""Each move uniform over the 12 actions, moves independent (cube.py:208-209 draws faces and directions independently):
	chi-square of single moves and of adjacent pairs over 2^20 x 24 draws."""
	from rl_rubiks_b200 import cube
	a = cube.seeded_actions(1 << 20, 24, 99).long()
	counts = torch.bincount(a.reshape(-1), minlength=12).double()
	exp = a.numel() / 12
	assert float(((counts - exp) ** 2 / exp).sum()) < 40            # 11 dof: p(chi2 > 40) ~ 4e-5
	for lag in (1, 2, 3, 12):
		pair = (a[:, :-lag] * 12 + a[:, lag:]).reshape(-1)
		c2 = torch.bincount(pair, minlength=144).double()
		e2 = pair.numel() / 144
		assert float(((c2 - e2) ** 2 / e2).sum()) < 230             # 143 dof: p(chi2 > 230) ~ 6e-6


def test_host_buffer_seeded_and_packed_entry_points():
	from rl_rubiks_b200 import _native as N, cube
	for is2024 in (True, False):
		cube.set_is2024(is2024)
		n, depth, seed = 6001, 100, 5
		want = cube.scramble_seeded(n, depth, seed, 17).cpu().numpy()
		assert (cube.scramble_seeded(n, depth, seed, 17, host_out=True) == want).all()
		for d in (100, 37, 1, 20):
			acts = cube.seeded_actions(n, d, seed, 17).cpu().numpy()
			packed = cube.pack_actions(acts)
			assert packed.shape == (n, (d + 1) // 2) and (O.pack_actions(acts) == packed).all()
			assert (O.unpack_actions(packed, d) == acts).all()
			out = torch.empty(n, d, dtype=torch.uint8, device="cuda")
			N.check(N.lib.rb_unpack_actions(N.ptr(torch.from_numpy(packed).cuda()), N.ptr(out), n, d, N.stream_handle()))
			assert (out.cpu().numpy() == acts).all()
			f, dd = O.indices_to_actions(acts)
			assert (cube.scramble_batch_packed(packed, d) == O.scramble_many(f, dd, is2024)).all()
	with pytest.raises(IndexError):
		cube.scramble_batch_packed(np.full((4, 5), 12 + 13 * 3, np.uint8), 10)
	with pytest.raises(IndexError):
		cube.scramble_batch_packed(np.full((4, 5), 0, np.uint8), 9)          # odd depth: last byte must carry the "no move" digit
	N.check(N.lib.rbh_release())


def test_seeded_full_size_inverse_round_trip():
	"""2^22 cubes x 100 seeded moves: applying the reversed inverse of the dumped moves returns every cube to solved."""
	from rl_rubiks_b200 import cube
	n, depth = 1 << 22, 100
	scr = cube.scramble_seeded(n, depth, 2024)
	acts = cube.seeded_actions(n, depth, 2024)
	back = cube.scramble_batch((acts ^ 1).flip(1).contiguous(), start=scr)
	assert bool(cube.multi_is_solved(back).all())
	assert float(cube.multi_is_solved(scr).float().mean()) < 1e-3
