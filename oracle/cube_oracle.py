"""
CPU oracle for the rl-rubiks cube-dynamics hot path.  TEST INFRASTRUCTURE ONLY.

This module is a numpy restatement of the algorithms in the reference
(`peleiden/rl-rubiks`, `/root/reference`) that the CUDA kernels in
`rl_rubiks_b200/csrc` replace.  Nothing in the product path
(`rl_rubiks_b200/*`) may import it; only `tests/`, `__graft_entry__.smoke()`
and the CPU legs of `bench.py` do, and only as the checker / the CPU baseline.

Parity status: PINNED.  `tests/golden/make_golden.py` imports the reference
itself (in the build container, where `/root/reference` exists) and writes
`tests/golden/*.npz`; `tests/test_oracle_golden.py` checks every function below
against those fixtures and against the literal vectors of the reference's own
`tests/test_cube.py`.

Every function cites the reference file:line it restates.  The arithmetic is
all integer table lookups except one-hot emission (f32 0/1) and the ADI target
assembly (f32 add / compare, f64 loss weights).
"""
from __future__ import annotations

import numpy as np

# ---------------------------------------------------------------------------
# Actions.  Reference: librubiks/cube/cube.py:30-35 (action_space), :179-200.
# Action index a in [0,12) <-> (face = a // 2, direction = 1 - a % 2).
# ---------------------------------------------------------------------------
N_ACTIONS = 12
FACE_OF_ACTION = np.repeat(np.arange(6), 2).astype(np.uint8)        # 0 0 1 1 ...
DIR_OF_ACTION = np.tile(np.array([1, 0]), 6).astype(np.uint8)       # 1 0 1 0 ...


def action_index(face, direction):
	"""Inverse of cube.py:186-192 `indices_to_actions`: (face, dir) -> index."""
	return np.asarray(face).astype(np.int64) * 2 + (1 - np.asarray(direction).astype(np.int64))


def iter_actions(n: int = 1) -> np.ndarray:
	"""cube.py:179-184: uint8 (2, 12 n), faces row then directions row, tiled n times."""
	return np.stack([np.tile(FACE_OF_ACTION, n), np.tile(DIR_OF_ACTION, n)]).astype(np.uint8)


def indices_to_actions(idx: np.ndarray):
	"""cube.py:186-192."""
	idx = np.asarray(idx)
	return idx // 2, 1 - idx % 2


def rev_action(a: int) -> int:
	"""cube.py:194-195: the inverse move is the other direction of the same face."""
	return a ^ 1


def rev_actions(a: np.ndarray) -> np.ndarray:
	"""cube.py:197-200."""
	return np.asarray(a) ^ 1


# ---------------------------------------------------------------------------
# 20x24 representation tables.  Reference: librubiks/cube/maps.py:74-145.
# A state is int8[20]: 8 corners (value = 3*pos + orientation) then 12 edges
# (value = 2*pos + orientation), maps.py:101-105.
# ---------------------------------------------------------------------------
# Positive-direction 4-cycles of corner and edge positions per face and the
# orientation rule (maps.py:74-98).  Order F, B, T, D, L, R.
_CORNER_CYCLE = ((0, 1, 2, 3), (4, 7, 6, 5), (0, 3, 7, 4), (1, 5, 6, 2), (0, 4, 5, 1), (7, 3, 2, 6))
_EDGE_CYCLE = ((0, 1, 2, 3), (8, 11, 10, 9), (0, 7, 8, 4), (2, 5, 10, 6), (1, 4, 9, 5), (3, 6, 11, 7))
_CORNER_STATIC = (0, 0, 1, 1, 2, 2)       # orientation kept; the other two swap (maps.py:128)
_EDGE_FLIPS = (False, False, True, True, False, False)  # maps.py:135

KIND = np.array([0] * 8 + [1] * 12, dtype=np.int64)    # cube.py:240 corner_side_idcs


def build_delta_maps() -> np.ndarray:
	"""maps.py:107-145 `get_tensor_map`: int8 (2 [neg,pos], 6 faces, 2 [corner,edge], 24);
	new_value = value + maps[dir, face, kind, value]."""
	maps = np.zeros((2, 6, 2, 24), dtype=np.int8)
	for f in range(6):
		cc, ec, st, flip = _CORNER_CYCLE[f], _EDGE_CYCLE[f], _CORNER_STATIC[f], _EDGE_FLIPS[f]
		for j in range(4):
			for k in range(3):
				nk = k if k == st else 3 - st - k
				src, dst = 3 * cc[j] + k, 3 * cc[(j + 1) % 4] + nk
				maps[1, f, 0, src] = dst - src
				maps[0, f, 0, dst] = src - dst
			for k in range(2):
				nk = (1 - k) if flip else k
				src, dst = 2 * ec[j] + k, 2 * ec[(j + 1) % 4] + nk
				maps[1, f, 1, src] = dst - src
				maps[0, f, 1, dst] = src - dst
	return maps


DELTA_MAPS = build_delta_maps()


def build_lut2024() -> np.ndarray:
	"""Direct form of the delta tables: LUT[a, kind, s] = s + maps[dir(a), face(a), kind, s]
	(uint8 (12, 2, 24)); each LUT[a, kind] is a permutation of 0..23."""
	lut = np.empty((12, 2, 24), dtype=np.uint8)
	s = np.arange(24)
	for a in range(12):
		lut[a] = s + DELTA_MAPS[DIR_OF_ACTION[a], FACE_OF_ACTION[a]]
	return lut


LUT2024 = build_lut2024()


def solved_2024() -> np.ndarray:
	"""cube.py:58-65: cubie j sits at position j with orientation 0."""
	return np.concatenate([3 * np.arange(8), 2 * np.arange(12)]).astype(np.int8)


def solved_686() -> np.ndarray:
	"""cube.py:67-71: sticker colour one-hot, face f all colour f."""
	s = np.zeros((6, 8, 6), dtype=np.int8)
	for f in range(6):
		s[f, :, f] = 1
	return s


# ---------------------------------------------------------------------------
# 20x24 dynamics.  Reference: cube.py:244-263.
# ---------------------------------------------------------------------------
def rotate_2024(state: np.ndarray, face: int, direction: int) -> np.ndarray:
	"""cube.py:245-254."""
	d = DELTA_MAPS[int(direction), int(face)]
	return (state + d[KIND, state]).astype(state.dtype)


def multi_rotate_2024(states: np.ndarray, faces: np.ndarray, directions: np.ndarray) -> np.ndarray:
	"""cube.py:257-263: action (faces[i], directions[i]) on states[i]; returns a new array."""
	states = np.asarray(states)
	d = DELTA_MAPS[np.asarray(directions).astype(np.int64), np.asarray(faces).astype(np.int64)]  # (n,2,24)
	rows = np.arange(len(states))[:, None]
	return (states + d[rows, KIND[None, :], states.astype(np.int64)]).astype(states.dtype)


def multi_act_2024(states: np.ndarray, actions: np.ndarray) -> np.ndarray:
	"""Same transition through the direct LUT, indexed by action index."""
	states = np.asarray(states)
	a = np.asarray(actions).astype(np.int64)[:, None]
	return LUT2024[a, KIND[None, :], states.astype(np.int64)].astype(states.dtype)


def as_oh_2024(states: np.ndarray) -> np.ndarray:
	"""cube.py:265-277: f32 (n, 480), oh[i, 24 j + states[i, j]] = 1.  1-D input -> (1, 480)."""
	states = np.atleast_2d(np.asarray(states))
	n = len(states)
	oh = np.zeros((n, 480), dtype=np.float32)
	cols = 24 * np.arange(20)[None, :] + states.astype(np.int64)
	oh[np.arange(n)[:, None], cols] = 1
	return oh


# ---------------------------------------------------------------------------
# 6x8x6 representation.  Reference: cube.py:311-361, maps.py:149-156.
# state[f, p, c] = 1 iff sticker p (clockwise ring index) of face f has colour c.
# ---------------------------------------------------------------------------
NEIGHBORS_686 = np.array([
	[4, 3, 5, 2], [3, 4, 2, 5], [0, 5, 1, 4], [5, 0, 4, 1], [2, 1, 3, 0], [1, 2, 0, 3],
])  # maps.py:149-156, neighbours of each face in positive direction
_ADJ = np.array([6, 7, 0, 2, 3, 4, 4, 5, 6, 0, 1, 2])              # cube.py:316
_ADJ_ROLLED = np.roll(_ADJ, 3)                                      # cube.py:317
_BLOCK_03 = np.repeat(np.arange(4), 3)                              # cube.py:311
_BLOCK_N13 = _BLOCK_03 - 1                                          # cube.py:312


def rotate_686(state: np.ndarray, face: int, direction: int) -> np.ndarray:
	"""cube.py:330-347.  The turned face's 8-ring shifts by two; the 12 adjacent
	stickers move one neighbour face along."""
	out = state.copy()
	nb = NEIGHBORS_686[face]
	ring = state[nb]                       # (4, 8, ...) the four neighbour faces
	if direction:
		out[face] = state[face, (np.arange(8) - 2) % 8]
		out[nb[_BLOCK_03], _ADJ] = ring[_BLOCK_N13, _ADJ_ROLLED]
	else:
		out[face] = state[face, (np.arange(8) + 2) % 8]
		out[nb[_BLOCK_N13], _ADJ_ROLLED] = ring[_BLOCK_03, _ADJ]
	return out


def build_perm686() -> np.ndarray:
	"""Gather form of cube.py:330-347: PERM[a, slot] = source sticker slot (f*8+p) whose
	content lands in `slot` under action a (uint8 (12, 48)).  Obtained by pushing slot
	labels through `rotate_686`."""
	labels = np.arange(48).reshape(6, 8, 1)
	perm = np.empty((12, 48), dtype=np.uint8)
	for a in range(12):
		perm[a] = rotate_686(labels, int(FACE_OF_ACTION[a]), int(DIR_OF_ACTION[a])).reshape(48)
	return perm


PERM686 = build_perm686()


def multi_rotate_686(states: np.ndarray, faces: np.ndarray, directions: np.ndarray) -> np.ndarray:
	"""cube.py:349-361, vectorised through PERM686 (the reference loops per state)."""
	states = np.asarray(states)
	n = len(states)
	a = action_index(faces, directions)
	flat = states.reshape(n, 48, 6)
	return flat[np.arange(n)[:, None], PERM686[a].astype(np.int64)].reshape(n, 6, 8, 6)


def as_oh_686(states: np.ndarray) -> np.ndarray:
	"""cube.py:363-369: already one-hot; ravel to (n, 288) and widen to f32."""
	states = np.asarray(states)
	if states.ndim == 3:
		states = states[None]
	return states.reshape(len(states), 288).astype(np.float32)


def as_correct_686(oh: np.ndarray) -> np.ndarray:
	"""cube.py:371-380: (n, 288) -> f32 (n, 6, 8): +1 where the sticker equals the
	solved sticker in all 6 colour channels, else -1."""
	t = np.asarray(oh).reshape(len(oh), 6, 8, 6)
	ok = (t == solved_686()[None]).all(axis=3)
	return np.where(ok, 1.0, -1.0).astype(np.float32)


# ---------------------------------------------------------------------------
# Representation-polymorphic helpers (cube.py:41-52, 77-89, 127-147).
# ---------------------------------------------------------------------------
def solved(is2024: bool) -> np.ndarray:
	return solved_2024() if is2024 else solved_686()


def rotate(state, face, direction, is2024: bool):
	return rotate_2024(state, face, direction) if is2024 else rotate_686(state, face, direction)


def multi_rotate(states, faces, directions, is2024: bool):
	return (multi_rotate_2024 if is2024 else multi_rotate_686)(states, faces, directions)


def multi_is_solved(states: np.ndarray, is2024: bool) -> np.ndarray:
	"""cube.py:88-89."""
	s = solved(is2024)
	return (np.asarray(states) == s).all(axis=tuple(range(1, s.ndim + 1)))


def is_solved(state: np.ndarray, is2024: bool) -> bool:
	"""cube.py:85-86."""
	return bool((np.asarray(state) == solved(is2024)).all())


def as_oh(states, is2024: bool) -> np.ndarray:
	return as_oh_2024(states) if is2024 else as_oh_686(states)


def expand12(states: np.ndarray, is2024: bool) -> np.ndarray:
	"""The 12-neighbour idiom of train.py:285 / agents.py:277-281:
	multi_rotate(np.repeat(S, 12, 0), *iter_actions(len(S))); child i*12+a = action a on state i."""
	states = np.asarray(states)
	rep = np.repeat(states, 12, axis=0)
	f, d = iter_actions(len(states))
	return multi_rotate(rep, f, d, is2024)


# ---------------------------------------------------------------------------
# Scramblers.  Reference: cube.py:206-234.
# ---------------------------------------------------------------------------
def scramble(faces: np.ndarray, directions: np.ndarray, is2024: bool) -> np.ndarray:
	"""cube.py:206-211 with the random draw supplied by the caller: sequential moves from solved."""
	s = solved(is2024)
	for f, d in zip(faces, directions):
		s = rotate(s, int(f), int(d), is2024)
	return s


def scramble_many(faces: np.ndarray, directions: np.ndarray, is2024: bool) -> np.ndarray:
	"""n independent scrambles; faces/directions are (n, depth).  Final states only."""
	faces, directions = np.asarray(faces), np.asarray(directions)
	n, depth = faces.shape
	s = np.repeat(solved(is2024)[None], n, axis=0)
	for d in range(depth):
		s = multi_rotate(s, faces[:, d], directions[:, d], is2024)
	return s


def draw_sequence_actions(games: int, depth: int, rng=np.random):
	"""cube.py:226-227 draw order: faces (depth, games) first, then directions."""
	faces = rng.randint(0, 6, (depth, games))
	dirs = rng.randint(0, 2, (depth, games))
	return faces, dirs


def sequence_scrambler(faces: np.ndarray, directions: np.ndarray, with_solved: bool, is2024: bool):
	"""cube.py:218-234 with host-supplied draws of shape (depth, games).  Returns
	(states (games*depth, *shape) game-major / depth-minor, one-hot f32).  With
	`with_solved` the solved state is emitted first and only rows 0..depth-2 are applied."""
	faces, directions = np.asarray(faces), np.asarray(directions)
	depth, games = faces.shape
	cur = np.repeat(solved(is2024)[None], games, axis=0)
	seq = [cur] if with_solved else []
	for d in range(depth - int(with_solved)):
		cur = multi_rotate(cur, faces[d], directions[d], is2024)
		seq.append(cur)
	states = np.stack(seq, axis=1).reshape(games * depth, *cur.shape[1:])
	return states, as_oh(states, is2024)


# ---------------------------------------------------------------------------
# Device-seeded scramble stream (no reference analogue beyond the distribution of cube.py:208-209: every move uniform over
# 6 faces x 2 directions).  The generator is Philox4x32-10 (Salmon et al., SC'11; Random123 / cuRAND), restated here from the
# paper and pinned on the Random123 known-answer vectors (tests/test_oracle_golden.py); the word -> moves rule is the one
# csrc/rb_scramble_seeded.cuh documents.
# ---------------------------------------------------------------------------
_PHILOX_M0, _PHILOX_M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_PHILOX_W0, _PHILOX_W1 = 0x9E3779B9, 0xBB67AE85


def philox4x32_10(counter: np.ndarray, key) -> np.ndarray:
	"""counter: uint32 (n, 4); key: (k0, k1).  Returns uint32 (n, 4).  Ten rounds of
	(c0, c1, c2, c3) <- (hi(M1*c2) ^ c1 ^ k0, lo(M1*c2), hi(M0*c0) ^ c3 ^ k1, lo(M0*c0)), the key bumped by the Weyl constants
	between rounds."""
	c = np.asarray(counter, dtype=np.uint64).copy()
	k0, k1 = int(key[0]) & 0xffffffff, int(key[1]) & 0xffffffff
	mask = np.uint64(0xffffffff)
	for _ in range(10):
		p0, p1 = _PHILOX_M0 * c[:, 0], _PHILOX_M1 * c[:, 2]
		n0 = (p1 >> np.uint64(32)) ^ c[:, 1] ^ np.uint64(k0)
		n2 = (p0 >> np.uint64(32)) ^ c[:, 3] ^ np.uint64(k1)
		c = np.stack([n0, p1 & mask, n2, p0 & mask], axis=1)
		k0, k1 = (k0 + _PHILOX_W0) & 0xffffffff, (k1 + _PHILOX_W1) & 0xffffffff
	return c.astype(np.uint32)


def seeded_actions(seed: int, first_cube: int, n: int, depth: int) -> np.ndarray:
	"""Action indices uint8 (n, depth) of the device-seeded stream: cube i = Philox subsequence first_cube + i (counter words
	2, 3), block j = counter word 0; output word k of block j is triple q = 4 j + k; t = (word * 1728) >> 32;
	moves 3q, 3q+1, 3q+2 = t // 144, t // 12 % 12, t % 12."""
	n_triples = (depth + 2) // 3
	n_blocks = (n_triples + 3) // 4
	cube = (np.arange(n, dtype=np.uint64) + np.uint64(first_cube & 0xffffffffffffffff))
	out = np.empty((n, n_blocks * 12), dtype=np.uint8)
	for j in range(n_blocks):
		ctr = np.stack([np.full(n, j, np.uint64), np.zeros(n, np.uint64), cube & np.uint64(0xffffffff), cube >> np.uint64(32)], axis=1)
		words = philox4x32_10(ctr, (seed & 0xffffffff, (seed >> 32) & 0xffffffff)).astype(np.uint64)
		t = (words * np.uint64(1728)) >> np.uint64(32)
		blk = np.stack([t // 144, t // 12 % 12, t % 12], axis=2).reshape(n, 12)
		out[:, 12 * j:12 * j + 12] = blk
	return np.ascontiguousarray(out[:, :depth])


def pack_actions(actions: np.ndarray) -> np.ndarray:
	"""uint8 (n, depth) action indices -> packed uint8 (n, (depth+1)//2): p = a(2k) + 13 * a(2k+1), 12 = no second move."""
	a = np.asarray(actions, dtype=np.uint8)
	n, depth = a.shape
	if depth % 2:
		a = np.concatenate([a, np.full((n, 1), 12, np.uint8)], axis=1)
	return (a[:, 0::2] + 13 * a[:, 1::2]).astype(np.uint8)


def unpack_actions(packed: np.ndarray, depth: int) -> np.ndarray:
	p = np.asarray(packed, dtype=np.uint8)
	a = np.stack([p % 13, p // 13], axis=2).reshape(len(p), -1)
	return np.ascontiguousarray(a[:, :depth])


# ---------------------------------------------------------------------------
# ADI training-batch assembly.  Reference: librubiks/train.py:256-339.
# ---------------------------------------------------------------------------
REWARD_METHODS = ("paper", "lapanfix", "schultzfix", "reward0")


def adi_rewards(solved_children: np.ndarray, reward_method: str) -> np.ndarray:
	"""train.py:292-296: +1 (0 with reward0) for a solved child, -1 otherwise; f32."""
	r = np.full(solved_children.shape, 0.0 if reward_method == "reward0" else 1.0, dtype=np.float32)
	r[~solved_children] = -1
	return r


def adi_targets(values: np.ndarray, solved_children: np.ndarray, solved_states: np.ndarray,
				reward_method: str, depth: int):
	"""train.py:313-325.  values: f32 (12 n,) net outputs for the children.
	Returns (policy_targets int64 (n,), value_targets f32 (n,)).  argmax takes the first
	maximum (torch.argmax behaviour, SURVEY 8c iv); NaN counts as the maximum, as in torch."""
	v = (np.asarray(values, dtype=np.float32) + adi_rewards(solved_children, reward_method)).reshape(-1, 12)
	nan = np.isnan(v)
	policy = np.where(nan.any(axis=1), nan.argmax(axis=1), np.argmax(np.where(nan, -np.inf, v), axis=1)).astype(np.int64)
	value = v[np.arange(len(v)), policy].copy()
	if reward_method == "lapanfix":
		value[np.asarray(solved_states, dtype=bool)] = 0
	elif reward_method == "schultzfix":
		value[np.arange(0, len(v), depth)] = 0
	return policy, value


def adi_loss_weights(games: int, depth: int, alpha: float) -> np.ndarray:
	"""train.py:329-333, computed in float64 and cast to f32 like the reference."""
	weighted = np.tile(1 / np.arange(1, depth + 1), games)
	unweighted = np.ones_like(weighted)
	ws, us = weighted.sum(), len(unweighted)
	return (((1 - alpha) * weighted / ws + alpha * unweighted / us) * (ws + us)).astype(np.float32)


def adi_traindata(faces, directions, value_fn, reward_method: str, alpha: float, is2024: bool = True):
	"""train.py:256-339 with host-supplied draws (depth, games) and a callable
	`value_fn(children_oh f32 (12n, W)) -> f32 (12n,)` standing in for the net.
	Returns dict with every intermediate the CUDA path is checked on."""
	depth, games = np.asarray(faces).shape
	states, oh_states = sequence_scrambler(faces, directions, reward_method == "lapanfix", is2024)
	solved_states = multi_is_solved(states, is2024)
	children = expand12(states, is2024)
	children_oh = as_oh(children, is2024)
	solved_children = multi_is_solved(children, is2024)
	values = np.asarray(value_fn(children_oh), dtype=np.float32).reshape(-1)
	policy, value = adi_targets(values, solved_children, solved_states, reward_method, depth)
	return dict(states=states, oh_states=oh_states, solved_states=solved_states, children=children,
				children_oh=children_oh, solved_children=solved_children, values=values,
				policy_targets=policy, value_targets=value,
				loss_weights=adi_loss_weights(games, depth, alpha))


# ---------------------------------------------------------------------------
# Search-frontier bookkeeping.  Reference: librubiks/solving/agents.py.
# ---------------------------------------------------------------------------
class SeenSet:
	"""The dict keyed on state bytes used by AStar / BFS / MCTS (agents.py:103, 233, 286-303,
	466, 517-526).  `insert_unique` restates the frontier part of AStar.expand_batch
	(agents.py:286-306): for a batch of states in order, report which were seen before the
	batch, which are first occurrences within the batch, and the index of every state, new
	states getting len(self)+1, len(self)+2, ... in batch order (index 0 is never used)."""

	def __init__(self):
		self.index = {}

	def __len__(self):
		return len(self.index)

	def insert_unique(self, states: np.ndarray):
		keys = [np.ascontiguousarray(s).tobytes() for s in states]
		seen = np.array([k in self.index for k in keys], dtype=bool)
		first = np.zeros(len(keys), dtype=bool)
		mark = set()
		for i, k in enumerate(keys):
			if k not in mark:
				mark.add(k)
				first[i] = True
		base = len(self.index)
		k_new = 0
		for i, k in enumerate(keys):
			if first[i] and not seen[i]:
				k_new += 1
				self.index[k] = base + k_new
		idx = np.array([self.index[k] for k in keys], dtype=np.int64)
		return seen, first, idx

	def lookup(self, states: np.ndarray) -> np.ndarray:
		"""agents.py:606-607 (`_complete_graph`): index or 0 when absent."""
		return np.array([self.index.get(np.ascontiguousarray(s).tobytes(), 0) for s in states], dtype=np.int64)


def bfs_layers(max_depth: int, is2024: bool = True, start: np.ndarray | None = None):
	"""Layer-synchronous restatement of BFS.search (agents.py:96-123) without the
	early exit: returns the per-depth counts of newly discovered states and the
	SeenSet.  Parents are visited in discovery order and their children in action
	order, which is exactly the FIFO order of the reference."""
	seen = SeenSet()
	frontier = (solved(is2024) if start is None else np.asarray(start))[None]
	seen.insert_unique(frontier)
	counts = [1]
	for _ in range(max_depth):
		children = expand12(frontier, is2024)
		was_seen, first, _ = seen.insert_unique(children)
		new = first & ~was_seen
		frontier = children[new]
		counts.append(int(new.sum()))
	return counts, seen


def bfs_search(start: np.ndarray, max_states: int, is2024: bool = True):
	"""BFS.search (agents.py:96-123) restated literally: FIFO queue, dict keyed on the state bytes, the budget test
	`len(states) < max_states` before EVERY parent pop (agents.py:104), children in action order, a solved child ends the
	search before it is recorded.  Returns (found, len(agent), action queue)."""
	start = np.asarray(start)
	if is_solved(start, is2024):
		return True, 0, []
	states = {start.tobytes(): (None, None)}
	queue = [start]
	head = 0
	while len(states) < max_states:
		state = queue[head]; head += 1
		tstate = state.tobytes()
		for a in range(12):
			child = rotate(state, a // 2, 1 - a % 2, is2024)
			key = child.tobytes()
			if key in states:
				continue
			if is_solved(child, is2024):
				actions = [a]
				while states[tstate][0] is not None:
					actions.insert(0, states[tstate][1])
					tstate = states[tstate][0]
				return True, len(states), actions
			states[key] = (tstate, a)
			queue.append(child)
	return False, len(states), []


def bfs_layer_counts_packed(max_depth: int) -> list:
	"""Same counts for the 20x24 representation at scale (depth 7 = 9.2 M states) using
	sorted 128-bit packed keys instead of a Python dict; used to pin the KAT
	1, 12, 114, 1068, 10011, 93840, 878880, 8221632."""
	def pack(s):
		s = s.astype(np.uint64)
		lo = np.zeros(len(s), dtype=np.uint64)
		hi = np.zeros(len(s), dtype=np.uint64)
		for j in range(12):
			lo |= s[:, j] << np.uint64(5 * j)
		for j in range(12, 20):
			hi |= s[:, j] << np.uint64(5 * (j - 12))
		k = np.empty(len(s), dtype=[("hi", np.uint64), ("lo", np.uint64)])
		k["hi"], k["lo"] = hi, lo
		return k
	frontier = solved_2024()[None]
	seen = pack(frontier)
	counts = [1]
	for _ in range(max_depth):
		children = expand12(frontier, True)
		keys = pack(children)
		uniq, first_idx = np.unique(keys, return_index=True)
		fresh = ~np.isin(uniq, seen)
		first_idx = np.sort(first_idx[fresh])
		frontier = children[first_idx]
		seen = np.union1d(seen, uniq[fresh])
		counts.append(len(first_idx))
	return counts


class AStarFrontier:
	"""Restatement of the bookkeeping in AStar (agents.py:221-402) around a callable
	`h_fn(states) -> f32 (n,)` standing in for minus the net's value (agents.py:380-381).
	The open list is a heapq of (cost, index) tuples as in the reference; G is float64,
	index 0 is unused and the root is index 1 (agents.py:233-234)."""

	def __init__(self, lambda_: float, expansions: int, h_fn, is2024: bool = True):
		self.lambda_, self.expansions, self.h_fn, self.is2024 = lambda_, expansions, h_fn, is2024

	def reset(self, state: np.ndarray, capacity: int = 1 << 16):
		self.seen = SeenSet()
		self.seen.insert_unique(state[None])
		self.states = np.zeros((capacity, *state.shape), dtype=state.dtype)
		self.G = np.zeros(capacity)
		self.parents = np.zeros(capacity, dtype=np.int64)
		self.parent_actions = np.zeros(capacity, dtype=np.int64)
		self.states[1] = state
		self.open = [(0, 1)]

	def __len__(self):
		return len(self.seen)

	def _grow(self, need: int):
		while need >= len(self.G):                       # agents.py:396-402
			self.states = np.concatenate([self.states, np.zeros_like(self.states)])
			self.G = np.concatenate([self.G, np.zeros_like(self.G)])
			self.parents = np.concatenate([self.parents, np.zeros_like(self.parents)])
			self.parent_actions = np.concatenate([self.parent_actions, np.zeros_like(self.parent_actions)])

	def pop_batch(self) -> np.ndarray:
		"""agents.py:238-239."""
		import heapq
		n = min(len(self.open), self.expansions)
		return np.array([heapq.heappop(self.open)[1] for _ in range(n)], dtype=np.int64)

	def expand_batch(self, expand_idcs: np.ndarray):
		"""agents.py:254-331.  Returns (won, trace of the dedup step)."""
		import heapq
		self._grow(len(self) + 12 * len(expand_idcs) + 1)
		parents = np.repeat(expand_idcs, 12)
		acts = np.tile(np.arange(12), len(expand_idcs))
		children = expand12(self.states[expand_idcs], self.is2024)
		seen, first, idx = self.seen.insert_unique(children)
		new = first & ~seen
		new_idx = idx[new]
		self.states[new_idx] = children[new]
		self.G[new_idx] = self.G[parents[new]] + 1
		self.parent_actions[new_idx] = acts[new]
		self.parents[new_idx] = parents[new]
		trace = dict(seen=seen, first=first, idx=idx, new_idx=new_idx)
		if len(new_idx):
			H = np.asarray(self.h_fn(children[new]), dtype=np.float32).reshape(-1)
			cost = self.lambda_ * self.G[new_idx] + H          # agents.py:383 (f64 + f32 -> f64)
			for c, i in zip(cost, new_idx):
				heapq.heappush(self.open, (c, int(i)))
			if multi_is_solved(children[new], self.is2024).any():
				return True, trace
		old = first & seen                                      # agents.py:294, 327-328
		self.relax(idx[old], parents[old], acts[old])
		return False, trace

	def relax(self, state_idcs, parent_idcs, actions):
		"""agents.py:333-367; numpy fancy assignment, so the last duplicate wins."""
		G = self.G
		nw = G[parent_idcs] + 1 < G[state_idcs]
		G[state_idcs[nw]] = G[parent_idcs[nw]] + 1
		self.parent_actions[state_idcs[nw]] = actions[nw]
		self.parents[state_idcs[nw]] = parent_idcs[nw]
		sc = G[state_idcs] + 1 < G[parent_idcs]
		G[parent_idcs[sc]] = G[state_idcs[sc]] + 1
		self.parent_actions[parent_idcs[sc]] = rev_actions(actions[sc])
		self.parents[parent_idcs[sc]] = state_idcs[sc]


class MCTSOracle:
	"""Restatement of MCTS (agents.py:415-645) around `net_fn(oh f32 (n, W)) -> (policy logits (n, 12), values (n,))`.
	Node arrays are float64 / int as in the reference (agents.py:438-446); the policy goes through an f32 softmax
	(agents.py:472, 552).  `max_states` is the only stopping rule here (no wall-clock limit)."""

	def __init__(self, net_fn, c: float, search_graph: bool, is2024: bool = True, nu: float = 100):
		self.net_fn, self.c, self.search_graph, self.is2024, self.nu = net_fn, c, search_graph, is2024, nu

	def __len__(self):
		return len(self.seen)

	@staticmethod
	def _softmax32(x):
		x = np.asarray(x, dtype=np.float32)
		e = np.exp(x - x.max(axis=1, keepdims=True))
		return (e / e.sum(axis=1, keepdims=True, dtype=np.float32)).astype(np.float32)

	def _reset(self, n: int, shape):
		self.seen = SeenSet()
		self.states = np.empty((n, *shape), dtype=np.int8)
		self.neighbors = np.zeros((n, 12), dtype=int)
		self.leaves = np.ones(n, dtype=bool)
		self.P, self.V = np.empty((n, 12)), np.empty(n)
		self.N, self.W, self.L = np.zeros((n, 12), dtype=int), np.zeros((n, 12)), np.zeros((n, 12))
		self.action_queue = []

	def _grow(self):                                      # agents.py:449-458
		k = len(self.states)
		self.states = np.concatenate([self.states, np.empty_like(self.states)])
		self.neighbors = np.concatenate([self.neighbors, np.zeros((k, 12), dtype=int)])
		self.leaves = np.concatenate([self.leaves, np.ones(k, dtype=bool)])
		self.P, self.V = np.concatenate([self.P, np.empty((k, 12))]), np.concatenate([self.V, np.empty(k)])
		self.N = np.concatenate([self.N, np.zeros((k, 12), dtype=int)])
		self.W, self.L = np.concatenate([self.W, np.zeros((k, 12))]), np.concatenate([self.L, np.zeros((k, 12))])

	def search(self, state: np.ndarray, max_states: int) -> bool:
		"""agents.py:461-494."""
		self._reset(1000, state.shape)
		self.seen.insert_unique(state[None])
		self.states[1] = state
		if is_solved(state, self.is2024):
			return True
		p, v = self.net_fn(as_oh(state[None], self.is2024))
		self.P[1], self.V[1] = self._softmax32(p)[0], np.asarray(v).reshape(-1)[0]
		visited, actions = [1], []
		while len(self) + 12 <= max_states:
			leaf, act = self.expand_leaf(visited, actions)
			if leaf != -1:
				self.action_queue = list(actions) + [int(act)]
				if self.search_graph:
					self.complete_graph()
					self.shorten_action_queue(leaf)
				return True
			visited, actions = self.find_leaf()
		self.action_queue = list(actions)
		return False

	def expand_leaf(self, visited: list, actions: list):
		"""agents.py:496-573."""
		if len(self) + 12 > len(self.states):
			self._grow()
		leaf = visited[-1]
		sub = expand12(self.states[leaf][None], self.is2024)
		seen, _, idx = self.seen.insert_unique(sub)
		new_idx, new_states = idx[~seen], sub[~seen]
		self.states[new_idx] = new_states
		acts = np.arange(12)
		self.neighbors[leaf, acts] = idx
		self.neighbors[idx, rev_actions(acts)] = leaf
		self.leaves[leaf] = False
		solve_leaf = solve_action = -1
		hit = np.where(multi_is_solved(sub, self.is2024))[0]
		if hit.size:
			solve_leaf, solve_action = int(idx[hit[0]]), int(hit[0])
		p, v = self.net_fn(as_oh(new_states, self.is2024))
		v = np.asarray(v, dtype=np.float32).reshape(-1)
		self.P[new_idx] = self._softmax32(p)
		self.V[new_idx] = v
		self.W[leaf] = self.V[self.neighbors[leaf]]
		self.W[new_idx] = np.tile(v, (12, 1)).T
		if len(v):                                        # the reference raises on an empty batch (v.max() of nothing)
			self.W[visited[:-1], actions] = np.maximum(self.W[visited[:-1], actions], v.max())
		if actions:
			self.N[visited[:-1], actions] += 1
			self.L[visited[:-1], actions] = 0
			self.L[visited[1:], rev_actions(np.array(actions))] = 0
		return solve_leaf, solve_action

	def find_leaf(self):
		"""agents.py:575-595."""
		cur, visited, actions = 1, [1], []
		while not self.leaves[cur]:
			sqrtN = np.sqrt(self.N[cur].sum())
			U = self.c * self.P[cur] * sqrtN / (1 + self.N[cur])
			a = int((U + self.W[cur] - self.L[cur]).argmax())
			self.L[cur, a] += self.nu
			cur = int(self.neighbors[cur, a])
			self.L[cur, rev_action(a)] += self.nu
			visited.append(cur)
			actions.append(a)
		return visited, actions

	def complete_graph(self):
		"""agents.py:597-611."""
		leaves = np.where(self.leaves[:len(self) + 1])[0][1:]
		acts = np.tile(np.arange(12), len(leaves))
		rep = np.repeat(leaves, 12)
		idx = self.seen.lookup(expand12(self.states[leaves], self.is2024))
		self.neighbors[rep, acts] = idx
		self.neighbors[idx, rev_actions(acts)] = rep
		self.neighbors[0] = 0

	def shorten_action_queue(self, solved_index: int):
		"""agents.py:613-633: BFS over the neighbour table from the root to the solved node."""
		if solved_index == 1:
			return
		from collections import deque
		self.action_queue = []
		visited = {1: (None, None)}
		q = deque([1])
		while q:
			v = q.popleft()
			for i, n in enumerate(self.neighbors[v]):
				n = int(n)
				if not n or n in visited:
					continue
				if n == solved_index:
					self.action_queue.insert(0, i)
					while visited[v][0] is not None:
						self.action_queue.insert(0, visited[v][1])
						v = visited[v][0]
					return
				visited[n] = (v, i)
				q.append(n)


# ---------------------------------------------------------------------------
# 6x3x3 sticker view, used only to pin the oracle against the literal layouts in
# the reference's tests/test_cube.py:33-92.  Reference: cube.py:149-173, 279-307,
# 382-388, maps.py:26-51.
# ---------------------------------------------------------------------------
_F, _B, _T, _D, _L, _R = range(6)
_CORNER_633 = (
	((_F, 0, 0), (_L, 0, 2), (_T, 2, 0)), ((_F, 2, 0), (_D, 0, 0), (_L, 2, 2)),
	((_F, 2, 2), (_R, 2, 0), (_D, 0, 2)), ((_F, 0, 2), (_T, 2, 2), (_R, 0, 0)),
	((_B, 0, 2), (_T, 0, 0), (_L, 0, 0)), ((_B, 2, 2), (_L, 2, 0), (_D, 2, 0)),
	((_B, 2, 0), (_D, 2, 2), (_R, 2, 2)), ((_B, 0, 0), (_R, 0, 2), (_T, 0, 2)),
)
_EDGE_633 = (
	((_F, 0, 1), (_T, 2, 1)), ((_F, 1, 0), (_L, 1, 2)), ((_F, 2, 1), (_D, 0, 1)), ((_F, 1, 2), (_R, 1, 0)),
	((_T, 1, 0), (_L, 0, 1)), ((_D, 1, 0), (_L, 2, 1)), ((_D, 1, 2), (_R, 2, 1)), ((_T, 1, 2), (_R, 0, 1)),
	((_B, 0, 1), (_T, 0, 1)), ((_B, 1, 2), (_L, 1, 0)), ((_B, 2, 1), (_D, 2, 1)), ((_B, 1, 0), (_R, 1, 2)),
)


def as633_2024(state: np.ndarray) -> np.ndarray:
	"""cube.py:279-307."""
	out = np.repeat(np.arange(6), 9).reshape(6, 3, 3)
	for i in range(8):
		pos, ori = int(state[i]) // 3, int(state[i]) % 3
		if pos in (0, 2, 5, 7):
			ori = -ori
		colours = np.roll([c[0] for c in _CORNER_633[i]], ori)
		for slot, col in zip(_CORNER_633[pos], colours):
			out[slot] = col
	for i in range(12):
		pos, ori = int(state[i + 8]) // 2, int(state[i + 8]) % 2
		colours = np.roll([c[0] for c in _EDGE_633[i]], ori)
		for slot, col in zip(_EDGE_633[pos], colours):
			out[slot] = col
	return out


_RING_TO_33 = np.array([0, 3, 6, 7, 8, 5, 2, 1])   # cube.py:324
_RING_SHIFT = np.array([0, 6, 6, 4, 2, 4])         # cube.py:326


def as633_686(state: np.ndarray) -> np.ndarray:
	"""cube.py:382-388."""
	colours = np.where(state == 1)[2].reshape(6, 8)
	out = np.repeat(np.arange(6), 9).reshape(6, 9)
	for f in range(6):
		out[f, _RING_TO_33] = np.roll(colours[f], -_RING_SHIFT[f])
	return out.reshape(6, 3, 3)


def as633(state, is2024: bool):
	return as633_2024(state) if is2024 else as633_686(state)


def stringify(state, is2024: bool) -> str:
	"""cube.py:160-173: 9x12 unfolded-cross text layout."""
	s = as633(state, is2024)
	canvas = np.full((9, 12), " ", dtype="<U1")
	place = {_T: (0, 1), _L: (1, 0), _F: (1, 1), _R: (1, 2), _B: (1, 3), _D: (2, 1)}
	for f, (r, c) in place.items():
		canvas[3 * r:3 * r + 3, 3 * c:3 * c + 3] = s[f].astype(str)
	return "\n".join(" ".join(row) for row in canvas)
