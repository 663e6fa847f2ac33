"""CPU emulation of the slot-major macro-move scramble (rl_rubiks_b200/csrc/rb_scramble_macro.cuh) from the table the
library exports: the same row decoding, PRMT gathers, twist accumulation and final cubie-major rebuild as the kernel, in
numpy, against the oracle.  Validates the table and the algorithm without a GPU; the kernel itself is checked by the
`-m gpu` scramble tests."""
import numpy as np
import pytest

from oracle import cube_oracle as O


def _table():
	from rl_rubiks_b200 import _native as N
	rows = np.empty((13 ** 3, 6), dtype=np.uint32)
	N.check(N.lib.rb_get_macro_table(rows.ctypes.data))
	return rows


def _prmt(a, b, sel):
	"""Byte gather from the 8 bytes {a (0-3), b (4-7)} with 4 selector nibbles; a, b: (n, 4) uint8, sel: (n,) uint32."""
	src = np.concatenate([a, b], axis=1)
	idx = np.stack([(sel >> (4 * i)) & 7 for i in range(4)], axis=1).astype(np.int64)
	return np.take_along_axis(src, idx, axis=1)


def _emulate(actions, rows):
	n, depth = actions.shape
	pad = (-depth) % 3
	a = np.concatenate([actions, np.full((n, pad), 12, np.uint8)], axis=1).astype(np.int64)
	C = np.tile(np.arange(8, dtype=np.uint8), (n, 1))
	W = np.zeros((n, 8), dtype=np.int64)
	E = np.tile(np.arange(12, dtype=np.uint8), (n, 1))
	for m in range(0, depth + pad, 3):
		r = rows[a[:, m] + 13 * a[:, m + 1] + 169 * a[:, m + 2]]
		sc, tw, fl = r[:, 0], r[:, 4], r[:, 5]
		C0, C1 = _prmt(C[:, :4], C[:, 4:], sc), _prmt(C[:, :4], C[:, 4:], sc >> 16)
		Wb = W.astype(np.uint8)
		assert (W < 256).all()
		W0, W1 = _prmt(Wb[:, :4], Wb[:, 4:], sc).astype(np.int64), _prmt(Wb[:, :4], Wb[:, 4:], sc >> 16).astype(np.int64)
		W0 += np.stack([(tw >> (8 * i)) & 15 for i in range(4)], axis=1)
		W1 += np.stack([(tw >> (8 * i + 4)) & 15 for i in range(4)], axis=1)
		C, W = np.concatenate([C0, C1], 1), np.concatenate([W0, W1], 1)
		Es = []
		for d in range(3):
			s = r[:, 1 + d]
			x = _prmt(E[:, :4], E[:, 4:8], s)
			e = _prmt(x, E[:, 8:], s >> 16)
			e = e ^ (np.stack([(fl >> (8 * i + 4 + d)) & 1 for i in range(4)], axis=1).astype(np.uint8) << 4)
			Es.append(e)
		E = np.concatenate(Es, 1)
	out = np.zeros((n, 20), dtype=np.int8)
	rows_i = np.arange(n)
	for q in range(8):
		t = W[:, q] % 3
		ori = np.where(np.isin(q, (0, 2, 5, 7)), (3 - t) % 3, t)
		out[rows_i, C[:, q] & 7] = 3 * q + ori
	for q in range(12):
		out[rows_i, 8 + (E[:, q] & 15)] = 2 * q + ((E[:, q] >> 4) & 1)
	return out


@pytest.mark.parametrize("depth", [1, 2, 3, 4, 11, 12, 13, 25, 100, 250])
def test_macro_scramble_emulation_matches_oracle(depth):
	rows = _table()
	g = np.random.RandomState(depth)
	actions = g.randint(0, 12, (300, depth)).astype(np.uint8)
	f, d = O.indices_to_actions(actions)
	assert (_emulate(actions, rows) == O.scramble_many(f, d, True)).all()


def test_macro_table_rows_are_permutations():
	rows = _table()
	sc = rows[:, 0]
	src = np.stack([(sc >> (4 * i)) & 15 for i in range(8)], axis=1)
	assert (np.sort(src, axis=1) == np.arange(8)).all()
	ident = rows[12 + 13 * 12 + 169 * 12]
	assert ident[0] == 0x76543210 and ident[4] == 0 and ident[5] == 0
