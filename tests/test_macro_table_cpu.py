"""CPU emulation of the slot-major macro-move scramble (rl_rubiks_b200/csrc/rb_scramble_macro.cuh) from the table the
library exports: the same row decoding, PRMT gathers, 5-bit twist accumulation with periodic folding and final cubie-major rebuild as the kernel, in
numpy, against the oracle.  Validates the table and the algorithm without a GPU; the kernel itself is checked by the
`-m gpu` scramble tests."""
import numpy as np
import pytest

from oracle import cube_oracle as O


def _table():
	from rl_rubiks_b200 import _native as N
	rows = np.empty((13 ** 2, 5), dtype=np.uint32)
	N.check(N.lib.rb_get_macro_table(rows.ctypes.data))
	return rows


def _table3():
	from rl_rubiks_b200 import _native as N
	rows = np.empty((12 ** 3, 5), dtype=np.uint32)
	N.check(N.lib.rb_get_macro3_table(rows.ctypes.data))
	return rows


def _prmt(a, b, sel):
	"""Byte gather from the 8 bytes {a (0-3), b (4-7)} with 4 selector nibbles; a, b: (n, 4) uint8, sel: (n,) uint32."""
	src = np.concatenate([a, b], axis=1)
	idx = np.stack([(sel >> (4 * i)) & 7 for i in range(4)], axis=1).astype(np.int64)
	return np.take_along_axis(src, idx, axis=1)


def _bytes_of(word):
	return np.stack([(word >> (8 * i)) & 0xff for i in range(4)], axis=1).astype(np.uint8)


def _fold(c):
	"""fold_twists(): keeps the 5-bit twist accumulators <= 10 without changing them mod 3."""
	t = (c & 31).astype(np.int64)
	assert (t <= 31).all()
	t = (t & 3) + ((t >> 2) & 7)
	assert (t <= 10).all()
	return ((c & 0xe0) | t.astype(np.uint8)).astype(np.uint8)


def _row_stream(actions, rows, rows3):
	"""The table rows a cube's action sequence selects, in order: 2-move rows (identity padded), or 3-move rows with the
	trailing depth % 3 moves taken from the 2-move table (k_scramble_macro3)."""
	n, depth = actions.shape
	a = actions.astype(np.int64)
	if rows3 is None:
		pad = (-depth) % 2
		a = np.concatenate([a, np.full((n, pad), 12, np.int64)], axis=1)
		for m in range(0, depth + pad, 2):
			yield rows[a[:, m] + 13 * a[:, m + 1]]
		return
	m = 0
	for m in range(0, depth - depth % 3, 3):
		yield rows3[a[:, m] + 12 * a[:, m + 1] + 144 * a[:, m + 2]]
	m = depth - depth % 3
	if m < depth:
		a1 = a[:, m + 1] if m + 1 < depth else np.full(n, 12, np.int64)
		yield rows[a[:, m] + 13 * a1]


def _emulate(actions, rows, rows3=None, fold_every=10):
	n, depth = actions.shape
	C = np.tile((np.arange(8, dtype=np.uint8) << 5).astype(np.uint8), (n, 1))      # twist accumulator | id << 5
	E = np.tile(np.arange(12, dtype=np.uint8), (n, 1))                             # id | flip << 4
	steps = 0
	for r in _row_stream(actions, rows, rows3):
		sc, tf = r[:, 0], r[:, 4]
		c0 = _prmt(C[:, :4], C[:, 4:], sc).astype(np.int64) + _bytes_of(tf & 0x03030303)
		c1 = _prmt(C[:, :4], C[:, 4:], sc >> 16).astype(np.int64) + _bytes_of((tf >> 2) & 0x03030303)
		assert ((c0 & 31) >= (_bytes_of(tf & 0x03030303))).all() and (c0 < 256).all() and (c1 < 256).all()   # no carry into the id bits
		C = np.concatenate([c0, c1], 1).astype(np.uint8)
		Es = []
		for d in range(3):
			s = r[:, 1 + d]
			x = _prmt(E[:, :4], E[:, 4:8], s)
			e = _prmt(x, E[:, 8:], s >> 16)
			Es.append(e ^ _bytes_of(tf & (0x10101010 << d)))                    # partial flip bit 4 + d
		E = np.concatenate(Es, 1)
		steps += 1
		if steps == fold_every:                                                    # kReduceEvery (2-move rows) / 8 (3-move rows)
			steps, C = 0, _fold(C)
	out = np.zeros((n, 20), dtype=np.int8)
	rows_i = np.arange(n)
	for q in range(8):
		t = (C[:, q] & 31).astype(np.int64) % 3
		ori = np.where(np.isin(q, (0, 2, 5, 7)), (3 - t) % 3, t)
		out[rows_i, C[:, q] >> 5] = 3 * q + ori
	for q in range(12):
		out[rows_i, 8 + (E[:, q] & 15)] = 2 * q + (((E[:, q] >> 4) ^ (E[:, q] >> 5) ^ (E[:, q] >> 6)) & 1)
	return out


@pytest.mark.parametrize("depth", [1, 2, 3, 4, 11, 12, 13, 25, 100, 250])
def test_macro_scramble_emulation_matches_oracle(depth):
	rows = _table()
	g = np.random.RandomState(depth)
	actions = g.randint(0, 12, (300, depth)).astype(np.uint8)
	f, d = O.indices_to_actions(actions)
	assert (_emulate(actions, rows) == O.scramble_many(f, d, True)).all()


@pytest.mark.parametrize("depth", [1, 2, 3, 4, 5, 11, 12, 13, 24, 25, 26, 100, 250])
def test_macro3_scramble_emulation_matches_oracle(depth):
	"""3-move rows (12^3) with the depth % 3 tail from the 2-move table; the kernel folds after every 8 rows (24 moves)."""
	rows, rows3 = _table(), _table3()
	g = np.random.RandomState(1000 + depth)
	actions = g.randint(0, 12, (300, depth)).astype(np.uint8)
	f, d = O.indices_to_actions(actions)
	assert (_emulate(actions, rows, rows3, fold_every=8) == O.scramble_many(f, d, True)).all()


def _emulate_inverse(actions, rows, rows3):
	"""k_scramble_macro3 as shipped: the kernel multiplies the INVERSE moves in REVERSE order, so its slot-major result is the
	inverse group element -- byte q then holds the position (and minus the twist) of CUBIE q, i.e. the reference's cubie-major
	state up to a per-byte formula, with no scatter.  Row order as in the kernel: the depth % 24 remainder at the end of the
	sequence first, then the 24-move groups from the last to the first; folds after the remainder and after every group."""
	n, depth = actions.shape
	inv = (actions.astype(np.int64) ^ 1)
	C = np.tile((np.arange(8, dtype=np.uint8) << 5).astype(np.uint8), (n, 1))
	E = np.tile(np.arange(12, dtype=np.uint8), (n, 1))

	def apply(r, C, E):
		sc, tf = r[:, 0], r[:, 4]
		c0 = _prmt(C[:, :4], C[:, 4:], sc).astype(np.int64) + _bytes_of(tf & 0x03030303)
		c1 = _prmt(C[:, :4], C[:, 4:], sc >> 16).astype(np.int64) + _bytes_of((tf >> 2) & 0x03030303)
		assert (c0 < 256).all() and (c1 < 256).all() and ((c0 & 31) >= _bytes_of(tf & 0x03030303)).all()
		C = np.concatenate([c0, c1], 1).astype(np.uint8)
		Es = []
		for d in range(3):
			sel = r[:, 1 + d]
			x = _prmt(E[:, :4], E[:, 4:8], sel)
			Es.append(_prmt(x, E[:, 8:], sel >> 16) ^ _bytes_of(tf & (0x10101010 << d)))
		return C, np.concatenate(Es, 1)

	M = depth - depth % 24
	pos = depth
	while pos - 3 >= M:
		C, E = apply(rows3[inv[:, pos - 1] + 12 * inv[:, pos - 2] + 144 * inv[:, pos - 3]], C, E)
		pos -= 3
	if pos > M:
		a1 = inv[:, pos - 2] if pos - 2 >= M else np.full(n, 12, np.int64)
		C, E = apply(rows[inv[:, pos - 1] + 13 * a1], C, E)
	C = _fold(C)
	for g in range(depth // 24 - 1, -1, -1):
		for pos in range(24 * g + 24, 24 * g, -3):
			C, E = apply(rows3[inv[:, pos - 1] + 12 * inv[:, pos - 2] + 144 * inv[:, pos - 3]], C, E)
		C = _fold(C)
	# per-byte conversion: cubie q sits at position id; its twist is minus the accumulated one
	out = np.zeros((n, 20), dtype=np.int8)
	p = (C >> 5).astype(np.int64)
	t = (C & 31).astype(np.int64) % 3
	neg = np.isin(p, (0, 2, 5, 7))
	out[:, :8] = 3 * p + np.where(neg, t, (3 - t) % 3)
	pe = (E & 15).astype(np.int64)
	out[:, 8:] = 2 * pe + (((E >> 4) ^ (E >> 5) ^ (E >> 6)) & 1)
	return out


@pytest.mark.parametrize("depth", [1, 2, 3, 4, 5, 11, 12, 13, 23, 24, 25, 26, 47, 48, 50, 99, 100, 101, 250])
def test_macro3_inverse_sequence_emulation_matches_oracle(depth):
	rows, rows3 = _table(), _table3()
	g = np.random.RandomState(2000 + depth)
	actions = g.randint(0, 12, (300, depth)).astype(np.uint8)
	f, d = O.indices_to_actions(actions)
	assert (_emulate_inverse(actions, rows, rows3) == O.scramble_many(f, d, True)).all()


# ---- the kernels' other row / fold schedules -------------------------------------------------------------------------------
# A schedule is a list of ("r3", p): the 3-move row of moves p-3, p-2, p-1; ("r2", p, k): the 2-move row of the k (1 or 2) moves
# ending at p; ("fold",).  The twist accumulators have 5 bits: a schedule with more than 10 rows between two folds could carry
# into the id bits -- `apply` below asserts that it never happens.
def _run_schedule(actions, rows, rows3, schedule):
	n, depth = actions.shape
	inv = actions.astype(np.int64) ^ 1
	C = np.tile((np.arange(8, dtype=np.uint8) << 5).astype(np.uint8), (n, 1))
	E = np.tile(np.arange(12, dtype=np.uint8), (n, 1))
	covered = []
	for op in schedule:
		if op[0] == "fold":
			C = _fold(C)
			continue
		if op[0] == "r3":
			p = op[1]
			r = rows3[inv[:, p - 1] + 12 * inv[:, p - 2] + 144 * inv[:, p - 3]]
			covered += [p - 1, p - 2, p - 3]
		else:
			_, p, k = op
			r = rows[inv[:, p - 1] + 13 * (inv[:, p - 2] if k == 2 else np.full(n, 12, np.int64))]
			covered += [p - 1] + ([p - 2] if k == 2 else [])
		sc, tf = r[:, 0], r[:, 4]
		c0 = _prmt(C[:, :4], C[:, 4:], sc).astype(np.int64) + _bytes_of(tf & 0x03030303)
		c1 = _prmt(C[:, :4], C[:, 4:], sc >> 16).astype(np.int64) + _bytes_of((tf >> 2) & 0x03030303)
		assert ((c0 & 31) >= _bytes_of(tf & 0x03030303)).all() and ((c1 & 31) >= _bytes_of((tf >> 2) & 0x03030303)).all()    # no carry out of the 5 bits
		C = np.concatenate([c0, c1], 1).astype(np.uint8)
		Es = []
		for d in range(3):
			sel = r[:, 1 + d]
			x = _prmt(E[:, :4], E[:, 4:8], sel)
			Es.append(_prmt(x, E[:, 8:], sel >> 16) ^ _bytes_of(tf & (0x10101010 << d)))
		E = np.concatenate(Es, 1)
	assert covered == list(range(depth - 1, -1, -1))            # every move exactly once, last move first
	out = np.zeros((n, 20), dtype=np.int8)
	pc, t = (C >> 5).astype(np.int64), (C & 31).astype(np.int64) % 3
	out[:, :8] = 3 * pc + np.where(np.isin(pc, (0, 2, 5, 7)), t, (3 - t) % 3)
	out[:, 8:] = 2 * (E & 15).astype(np.int64) + (((E >> 4) ^ (E >> 5) ^ (E >> 6)) & 1)
	return out


def _schedule_a16(depth):
	"""k_scramble_macro3 mode 2 / 3 (depth % 16 == 0): pairs of 24-move groups anchored at multiples of 48 from the row start; what
	lies above them -- an odd group and / or a remainder of 8 or 16 moves -- goes first (rb_scramble_macro.cuh, `if (a16)`)."""
	assert depth % 16 == 0
	M = depth - depth % 24
	rem, top, odd = depth - M, (M // 48) * 48, (M // 24) & 1
	words12 = lambda p: [("r3", p), ("r3", p - 3), ("r3", p - 6), ("r3", p - 9)]                     # apply_words: 12 moves ending at p
	rem8 = lambda p: [("r3", p), ("r3", p - 3), ("r2", p - 6, 2)]
	rem4 = lambda p: [("r3", p), ("r2", p - 3, 1)]
	sch = []
	base = top + (24 if odd else 0)                                                                   # the remainder starts here
	if rem == 8:
		sch += rem8(base + 8) + [("fold",)]
	elif rem == 16:
		sch += words12(base + 16) + rem4(base + 4) + [("fold",)]
	if odd:
		sch += words12(top + 24) + words12(top + 12) + [("fold",)]
	for m in range(top, 0, -48):
		sch += words12(m) + words12(m - 12) + [("fold",)] + words12(m - 24) + words12(m - 36) + [("fold",)]
	return sch


def _schedule_seeded(depth):
	"""k_scramble_seeded: the trailing depth % 3 moves first, then the triples from the last to the first, four per Philox block;
	folds after the top block plus one, then after every second block, and at the end."""
	Q, rem = depth // 3, depth % 3
	n_words = Q + (1 if rem else 0)
	n_blocks = (n_words + 3) // 4
	sch, since = [], 0
	for j in range(n_blocks - 1, -1, -1):
		for k in (3, 2, 1, 0):
			q = 4 * j + k
			if q > Q or (q == Q and rem == 0):
				continue
			sch.append(("r2", depth, rem) if q == Q else ("r3", 3 * q + 3))
		since += 1
		if since == 2:
			sch.append(("fold",)); since = 0
	return sch + [("fold",)]


@pytest.mark.parametrize("depth", [16, 32, 48, 64, 80, 96, 112, 128, 144, 160, 256, 384, 400])
def test_macro3_a16_schedule_matches_oracle(depth):
	rows, rows3 = _table(), _table3()
	g = np.random.RandomState(3000 + depth)
	actions = g.randint(0, 12, (200, depth)).astype(np.uint8)
	f, d = O.indices_to_actions(actions)
	assert (_run_schedule(actions, rows, rows3, _schedule_a16(depth)) == O.scramble_many(f, d, True)).all()


@pytest.mark.parametrize("depth", [1, 2, 3, 4, 5, 11, 12, 13, 14, 23, 24, 25, 26, 47, 48, 49, 100, 101, 250])
def test_seeded_schedule_on_the_seeded_stream_matches_oracle(depth):
	"""The device-seeded kernel's row order and fold schedule, on the action stream the oracle's Philox restatement draws."""
	rows, rows3 = _table(), _table3()
	actions = O.seeded_actions(99 + depth, 5, 200, depth)
	f, d = O.indices_to_actions(actions)
	assert (_run_schedule(actions, rows, rows3, _schedule_seeded(depth)) == O.scramble_many(f, d, True)).all()
	# worst case for the accumulators: every row adds 2 to some twist
	worst = np.tile(np.array([0, 2, 4, 6, 8, 10], dtype=np.uint8), (4, depth // 6 + 1))[:, :depth]
	f, d = O.indices_to_actions(worst)
	assert (_run_schedule(worst, rows, rows3, _schedule_seeded(depth)) == O.scramble_many(f, d, True)).all()


def test_macro_table_rows_are_permutations():
	rows = _table()
	sc = rows[:, 0]
	src = np.stack([(sc >> (4 * i)) & 15 for i in range(8)], axis=1)
	assert (np.sort(src, axis=1) == np.arange(8)).all()
	ident = rows[12 + 13 * 12]
	assert ident[0] == 0x76543210 and ident[4] == 0
	assert ((rows[:, 4] & 0x80808080) == 0).all()                                   # bit 7 of every twist|flip byte unused


def test_sticker_tables_render_686_from_2024():
	"""The fast 6x8x6 scramble renders the 6x8x6 state from the 20x24 state of the same move sequence through the sticker
	tables the library derives (csrc/rb_tables.cuh build_stickers).  Emulated here in numpy against the 6x8x6 oracle, from
	solved and from arbitrary start states."""
	from rl_rubiks_b200 import _native as N
	ch, cd = np.empty((8, 3), np.uint8), np.empty((8, 24, 3), np.uint8)
	eh, ed = np.empty((12, 2), np.uint8), np.empty((12, 24, 2), np.uint8)
	N.check(N.lib.rb_get_stickers686(ch.ctypes.data, cd.ctypes.data, eh.ctypes.data, ed.ctypes.data))
	assert sorted(ch.ravel().tolist() + eh.ravel().tolist()) == list(range(48))          # every sticker slot exactly once
	g = np.random.RandomState(3)
	n, depth = 200, 37
	faces, dirs = g.randint(0, 6, (n, depth)), g.randint(0, 2, (n, depth))
	s2024 = O.scramble_many(faces, dirs, True)
	start = O.scramble_many(g.randint(0, 6, (n, 11)), g.randint(0, 2, (n, 11)), False).reshape(n, 48, 6)
	start[0] = g.randint(-5, 6, (48, 6))                                                # arbitrary int8 content is carried too
	want = start.copy()
	for m in range(depth):
		want = O.multi_rotate_686(want.reshape(n, 6, 8, 6), faces[:, m], dirs[:, m]).reshape(n, 48, 6)
	out = np.zeros_like(start)
	rows = np.arange(n)
	for c in range(8):
		for k in range(3):
			out[rows, cd[c, s2024[:, c], k]] = start[rows, ch[c, k]]
	for e in range(12):
		for k in range(2):
			out[rows, ed[e, s2024[:, 8 + e], k]] = start[rows, eh[e, k]]
	assert (out == want).all()
	assert (O.scramble_many(faces, dirs, False).reshape(n, 48, 6) == _render_solved(ch, cd, eh, ed, s2024)).all()


def _render_solved(ch, cd, eh, ed, s2024):
	n = len(s2024)
	out = np.zeros((n, 48, 6), np.int8)
	rows = np.arange(n)
	eye = np.eye(6, dtype=np.int8)
	for c in range(8):
		for k in range(3):
			out[rows, cd[c, s2024[:, c], k]] = eye[ch[c, k] // 8]
	for e in range(12):
		for k in range(2):
			out[rows, ed[e, s2024[:, 8 + e], k]] = eye[eh[e, k] // 8]
	return out


def _prmt_word(a: int, b: int, sel: int) -> int:
	"""prmt.b32 in its default mode on scalar words: nibble k of `sel` (low 16 bits) picks byte (nibble & 7) of {a, b}; bit 3 of
	the nibble replicates that byte's sign bit instead."""
	src = [(a >> (8 * i)) & 0xff for i in range(4)] + [(b >> (8 * i)) & 0xff for i in range(4)]
	out = 0
	for k in range(4):
		nib = (sel >> (4 * k)) & 0xf
		byte = src[nib & 7]
		if nib & 8:
			byte = 0xff if byte & 0x80 else 0x00
		out |= byte << (8 * k)
	return out


def test_register_lut_formulation_equals_the_table():
	"""rb_cube2024.cuh lut_sel / lut24: selector nibbles s & 15 with PRMT's sign-replicate bit doing the range masking, one byte mask
	for bit 4.  Emulated on scalar words for every action row, both kinds, and every 4-tuple pattern of values 0..23."""
	from oracle import cube_oracle as O
	lut = O.build_lut2024()                                     # (12, 2, 24): new value of a cubie with value s under action a
	rng = np.random.RandomState(0)
	words = [tuple(rng.randint(0, 24, 4)) for _ in range(400)] + [(s, s, s, s) for s in range(24)] + [(0, 7, 8, 15), (16, 23, 15, 8), (23, 0, 16, 7)]
	for a in range(12):
		for kind in range(2):
			row = [int.from_bytes(bytes(int(x) for x in lut[a, kind, 4 * k:4 * k + 4]), "little") for k in range(6)]
			for vals in words:
				w = int.from_bytes(bytes(int(v) for v in vals), "little")
				y = (w & 0x0f0f0f0f) | ((w >> 4) & 0xf0f0f0f0)
				sel = _prmt_word(y, y, 0x3320)
				selx = sel ^ 0x8888
				m2 = _prmt_word((w * 8) & 0xffffffff, 0, 0xba98)
				c0, c1, c2 = _prmt_word(row[0], row[1], sel), _prmt_word(row[2], row[3], selx), _prmt_word(row[4], row[5], sel)
				got = (c2 & m2) | ((c0 | c1) & ~m2 & 0xffffffff)
				want = int.from_bytes(bytes(int(lut[a, kind, v]) for v in vals), "little")
				assert got == want, (a, kind, vals)
