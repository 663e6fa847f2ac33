"""
Drop-in for `librubiks.cube` (reference: librubiks/cube/cube.py) backed by librubiks_b200.so.

Same module-level function API, same names, same argument meaning: two representations selected by the
module-global flag (`set_is2024` / `get_is2024` / `store_repr` / `restore_repr` / `with_used_repr`,
cube.py:96-124), purely functional calls that return new arrays, numpy int8 states on the host and a
torch f32 one-hot on the GPU (cube.py:130-133).  Every array-valued function additionally accepts torch CUDA
tensors and then keeps the data on the device (tensor in -> tensor out), which is how the ADI generator
and the search frontier use it.  There is no CPU path: all arithmetic runs in the sm_100a kernels.

Additions with no reference analogue (derivable from the reference API): `expand12`,
`scramble_batch`, `scramble_seeded` / `seeded_actions` (random draw made on the device), `pack_actions` /
`scramble_batch_packed` (two moves per byte over PCIe), `sequence_scrambler_from`.
"""
from __future__ import annotations

import functools

import numpy as np
import torch

from . import _native as N

# ---- action constants (cube.py:30-35) ------------------------------------------------------------
F, B, T, D, L, R = 0, 1, 2, 3, 4, 5
action_names = ('F', 'B', 'T', 'D', 'L', 'R')
action_space = [(f, d) for f in range(6) for d in (1, 0)]
action_dim = len(action_space)
dtype = np.int8

# ---- representation flag (cube.py:96-124) --------------------------------------------------------
_is2024 = True
_stored_repr = True


def set_is2024(is2024: bool):
	global _is2024
	assert type(is2024) is bool
	_is2024 = is2024


def get_is2024() -> bool:
	return _is2024


def store_repr():
	global _stored_repr
	_stored_repr = _is2024


def restore_repr():
	global _is2024
	_is2024 = _stored_repr


def with_used_repr(fun):
	"""Method decorator: runs the method with the representation set to `self.is2024` (cube.py:115-124)."""
	@functools.wraps(fun)
	def wrapper(self, *args, **kwargs):
		store_repr()
		set_is2024(self.is2024)
		res = fun(self, *args, **kwargs)
		restore_repr()
		return res
	return wrapper


def _rep() -> int:
	return N.REP_2024 if _is2024 else N.REP_686


def shape() -> tuple:
	return (20,) if _is2024 else (6, 8, 6)


def get_oh_shape() -> int:
	return 480 if _is2024 else 288


# ---- host tables: produced by the library, never retyped ------------------------------------------
def _solved_host(rep: int) -> np.ndarray:
	buf = np.empty(20 if rep == N.REP_2024 else 288, dtype=np.int8)
	N.check(N.lib.rb_get_solved(rep, buf.ctypes.data))
	return buf if rep == N.REP_2024 else buf.reshape(6, 8, 6)


_solved = {N.REP_2024: _solved_host(N.REP_2024), N.REP_686: _solved_host(N.REP_686)}
for _s in _solved.values():
	_s.setflags(write=False)


def get_solved_instance() -> np.ndarray:
	"""The shared read-only instance (cube.py:77-80)."""
	return _solved[_rep()]


def get_solved() -> np.ndarray:
	return get_solved_instance().copy()


def tables() -> dict:
	"""The move tables the kernels stage in shared memory, as numpy arrays (for inspection and tests)."""
	delta = np.empty((2, 6, 2, 24), dtype=np.int8)
	lut = np.empty((12, 2, 24), dtype=np.uint8)
	perm = np.empty((12, 48), dtype=np.uint8)
	N.check(N.lib.rb_get_delta_maps(delta.ctypes.data))
	N.check(N.lib.rb_get_lut2024(lut.ctypes.data))
	N.check(N.lib.rb_get_perm686(perm.ctypes.data))
	return dict(delta_maps=delta, lut2024=lut, perm686=perm)


# ---- marshalling ---------------------------------------------------------------------------------
def _dev():
	N.require_cuda()
	return torch.device("cuda", torch.cuda.current_device())


def _states_in(states):
	"""-> (contiguous int8 CUDA tensor shaped (n, *shape()), was_numpy, had_batch_dim)."""
	nd = len(shape())
	if isinstance(states, torch.Tensor):
		t, was_np = states, False
		if not t.is_cuda:
			t = t.to(_dev())
	else:
		t, was_np = torch.from_numpy(np.ascontiguousarray(states)).to(_dev()), True
	if t.dtype != torch.int8:
		t = t.to(torch.int8)
	batched = t.dim() == nd + 1
	if not batched:
		if t.dim() != nd:
			raise IndexError(f"state array of shape {tuple(t.shape)} does not match representation shape {shape()}")
		t = t.unsqueeze(0)
	if tuple(t.shape[1:]) != shape():
		raise IndexError(f"state array of shape {tuple(t.shape)} does not match representation shape {shape()}")
	return t.contiguous(), was_np, batched


def _u8_in(a, n=None):
	"""faces / directions / actions -> contiguous uint8 CUDA tensor (values are range-checked by the caller)."""
	if isinstance(a, torch.Tensor):
		t = a.to(device=_dev(), dtype=torch.uint8)
	else:
		arr = np.asarray(a)
		if arr.dtype == np.bool_:
			arr = arr.astype(np.uint8)
		if arr.size and (arr.min() < 0 or arr.max() > 255):
			raise IndexError("face / direction index out of range")
		t = torch.from_numpy(np.ascontiguousarray(arr.astype(np.uint8))).to(_dev())
	if n is not None and t.numel() != n:
		raise IndexError(f"expected {n} actions, got {t.numel()}")
	return t.contiguous()


def _out(t: torch.Tensor, was_np: bool):
	return t.cpu().numpy() if was_np else t


_range_scratch = {}


def _check_range(rep, states, faces, dirs):
	"""Reproduces the reference's IndexError on out-of-range input (numpy fancy indexing, cube.py:259-262)."""
	dev = torch.cuda.current_device()
	if dev not in _range_scratch:
		_range_scratch[dev] = torch.zeros(1, dtype=torch.int32, device=_dev())
	N.check(N.lib.rb_check_range(rep, N.ptr(states), 0 if states is None else states.shape[0], N.ptr(faces), N.ptr(dirs),
								 0 if faces is None else faces.numel(), N.ptr(_range_scratch[dev]), N.stream_handle()))


# ---- rotate logic (cube.py:41-52) ----------------------------------------------------------------
def multi_rotate(states, faces, directions):
	"""Performs action (faces[i], directions[i]) on states[i]; returns a new array (cube.py:49-52)."""
	s, was_np, _ = _states_in(states)
	n = s.shape[0]
	f, d = _u8_in(faces, n), _u8_in(directions, n)
	if was_np:      # host path: reproduce the reference's IndexError; device tensors are trusted (no sync)
		_check_range(_rep(), s, f, d)
	out = torch.empty_like(s)
	N.check(N.lib.rb_multi_rotate(_rep(), N.ptr(s), N.ptr(f), N.ptr(d), N.ptr(out), n, N.stream_handle()))
	return _out(out, was_np)


def rotate(state, face: int, direction: int):
	"""One move on one cube: side 0-5, direction 0 (negative) or 1 (positive) (cube.py:41-47)."""
	s, was_np, batched = _states_in(state)
	if batched:
		raise IndexError("rotate takes a single state; use multi_rotate for batches")
	f = torch.tensor([int(face)], dtype=torch.uint8, device=s.device) if 0 <= int(face) < 256 else None
	if f is None:
		raise IndexError("face index out of range")
	d = torch.tensor([int(direction) & 0xff], dtype=torch.uint8, device=s.device)
	if was_np:
		_check_range(_rep(), s, f, d)
	out = torch.empty_like(s)
	N.check(N.lib.rb_multi_rotate(_rep(), N.ptr(s), N.ptr(f), N.ptr(d), N.ptr(out), 1, N.stream_handle()))
	return _out(out[0], was_np)


# ---- solving logic (cube.py:85-89) ---------------------------------------------------------------
def multi_is_solved(states):
	s, was_np, _ = _states_in(states)
	flags = torch.empty(s.shape[0], dtype=torch.uint8, device=s.device)
	N.check(N.lib.rb_multi_is_solved(_rep(), N.ptr(s), N.ptr(flags), s.shape[0], N.stream_handle()))
	flags = flags.bool()
	return flags.cpu().numpy() if was_np else flags


def is_solved(state) -> bool:
	s, _, batched = _states_in(state)
	if batched:
		raise IndexError("is_solved takes a single state")
	return bool(multi_is_solved(s)[0].item())


# ---- one-hot (cube.py:130-140, 265-277, 363-380) ---------------------------------------------------
def _oh_fn(name: str, dtype):
	"""The f32 entry point (the reference's one-hot dtype) or its bfloat16 variant (opt-in: same 0/1 values, half the bytes)."""
	if dtype == torch.float32:
		return getattr(N.lib, name)
	if dtype == torch.bfloat16:
		return getattr(N.lib, name + "_bf16")
	raise TypeError(f"one-hot dtype must be torch.float32 or torch.bfloat16, got {dtype}")


def as_oh(states, dtype=torch.float32) -> torch.Tensor:
	"""n states -> f32 (n, 480 | 288) on the GPU; a single state gives a leading dimension of 1.  `dtype=torch.bfloat16`
	(not in the reference) emits the same rows as bf16 for a bf16 forward of the net."""
	s, _, _ = _states_in(states)
	oh = torch.empty(s.shape[0], get_oh_shape(), dtype=dtype, device=s.device)
	N.check(_oh_fn("rb_as_oh", dtype)(_rep(), N.ptr(s), N.ptr(oh), s.shape[0], N.stream_handle()))
	return oh


def as_correct(t: torch.Tensor) -> torch.Tensor:
	assert not get_is2024(), "Correctness representation is only implemented for 6x8x6 representation"
	x = t.to(device=_dev(), dtype=torch.float32).reshape(len(t), 288).contiguous()
	out = torch.empty(len(t), 6, 8, dtype=torch.float32, device=x.device)
	N.check(N.lib.rb_as_correct_686(N.ptr(x), N.ptr(out), len(t), N.stream_handle()))
	return out


def to_686(states2024):
	"""The same cubes in the 6x8x6 representation: (n, 20) -> (n, 6, 8, 6).  (The reference converts both to the 6x3x3 sticker
	view, cube.py:149-173; the library goes from one to the other directly, on the device.)"""
	was_np = not isinstance(states2024, torch.Tensor)
	s = (torch.from_numpy(np.ascontiguousarray(states2024, dtype=np.int8)) if was_np else states2024).to(device=_dev(), dtype=torch.int8).reshape(-1, 20).contiguous()
	out = torch.empty(s.shape[0], 6, 8, 6, dtype=torch.int8, device=s.device)
	N.check(N.lib.rb_as686(N.ptr(s), N.ptr(out), s.shape[0], N.stream_handle()))
	return _out(out, was_np)


def to_2024(states686):
	"""(n, 6, 8, 6) -> (n, 20); IndexError if a state is not a reachable cube."""
	was_np = not isinstance(states686, torch.Tensor)
	s = (torch.from_numpy(np.ascontiguousarray(states686, dtype=np.int8)) if was_np else states686).to(device=_dev(), dtype=torch.int8).reshape(-1, 6, 8, 6).contiguous()
	out = torch.empty(s.shape[0], 20, dtype=torch.int8, device=s.device)
	ok = torch.empty(s.shape[0], dtype=torch.uint8, device=s.device)
	N.check(N.lib.rb_as2024(N.ptr(s), N.ptr(out), N.ptr(ok), s.shape[0], N.stream_handle()))
	if not bool(ok.all().item()):
		raise IndexError("a 6x8x6 state is not a reachable cube")
	return _out(out, was_np)


# ---- action logic (cube.py:142-147, 179-200) -------------------------------------------------------
def repeat_state(state: np.ndarray, n: int = action_dim) -> np.ndarray:
	"""n copies of one state, shape (n, *shape()) (cube.py:142-147)."""
	if isinstance(state, torch.Tensor):
		return state.unsqueeze(0).repeat(n, *[1] * len(shape()))
	return np.tile(state, [n, *[1] * len(shape())])


_ITER_ROW = np.array([[f for f, _ in action_space], [d for _, d in action_space]], dtype=np.uint8)


def iter_actions(n: int = 1) -> np.ndarray:
	"""uint8 (2, 12 n): tiled faces row and directions row (cube.py:179-184)."""
	return np.tile(_ITER_ROW, (1, n))


def indices_to_actions(indices):
	"""Action indices [0, 12) -> (faces, dirs) (cube.py:186-192)."""
	faces = indices // 2
	dirs = ~(indices % 2) + 2
	return faces, dirs


def rev_action(action: int) -> int:
	return action + 1 if action % 2 == 0 else action - 1


def rev_actions(actions):
	rev = actions - 1
	rev[actions % 2 == 0] += 2
	return rev


def expand12(states, with_oh: bool = False, with_solved: bool = False, oh_dtype=torch.float32):
	"""The 12-neighbour expansion `multi_rotate(np.repeat(S, 12, 0), *iter_actions(len(S)))` (train.py:285,
	agents.py:277-281) as one fused kernel: child i*12+a = action a on state i.  Optionally also returns the
	children's one-hot (f32 CUDA tensor) and solved flags from the same launch."""
	s, was_np, _ = _states_in(states)
	n = s.shape[0]
	children = torch.empty(12 * n, *shape(), dtype=torch.int8, device=s.device)
	oh = torch.empty(12 * n, get_oh_shape(), dtype=oh_dtype, device=s.device) if with_oh else None
	flags = torch.empty(12 * n, dtype=torch.uint8, device=s.device) if with_solved else None
	N.check(_oh_fn("rb_expand12", oh_dtype)(_rep(), N.ptr(s), N.ptr(children), N.ptr(oh), N.ptr(flags), n, N.stream_handle()))
	res = [_out(children, was_np)]
	if with_oh:
		res.append(oh)
	if with_solved:
		res.append(flags.bool().cpu().numpy() if was_np else flags.bool())
	return res[0] if len(res) == 1 else tuple(res)


# ---- scramble logic (cube.py:206-234) ---------------------------------------------------------------
def _actions_u8(faces, dirs):
	"""(faces, dirs) numpy draws -> uint8 action indices a = 2*face + 1 - dir, range-checked on the host."""
	faces, dirs = np.asarray(faces), np.asarray(dirs)
	if faces.size and (faces.min() < 0 or faces.max() > 5 or dirs.min() < 0 or dirs.max() > 1):
		raise IndexError("face / direction index out of range")
	return (faces * 2 + (1 - dirs)).astype(np.uint8)


def scramble_batch(actions, start=None):
	"""`depth` moves on each of n cubes, final states only.  actions: (n, depth) action indices (numpy or CUDA
	tensor); start: optional (n, *shape()) states, default solved.  Batched form of cube.py:206-211."""
	a = _u8_in(actions)
	if a.dim() != 2:
		raise IndexError("actions must be (n, depth)")
	n, depth = a.shape
	was_np = not isinstance(actions, torch.Tensor)
	if was_np:
		_check_range(_rep(), None, a, None)
	st = None
	if start is not None:
		st, _, _ = _states_in(start)
	out = torch.empty(n, *shape(), dtype=torch.int8, device=a.device)
	N.check(N.lib.rb_scramble(_rep(), N.ptr(a), depth, 1, N.ptr(st), N.ptr(out), n, depth, N.stream_handle()))
	return _out(out, was_np)


def scramble_seeded(n: int, depth: int, seed: int, first_cube: int = 0, start=None, host_out: bool = False):
	"""`cube.scramble(depth)` for n cubes with the random draw made on the device: cube i takes `depth` moves, each uniform
	over the 12 actions, from subsequence `first_cube + i` of the Philox4x32-10 stream keyed by `seed` (the result for a cube
	does not depend on how cubes are split over calls or GPUs).  Returns the (n, *shape()) final states as a CUDA tensor, or
	written straight into a numpy array through the host-buffer C-ABI call when `host_out` (nothing but the seed is sent to the
	device).  `seeded_actions` returns the moves that were applied."""
	n, depth, seed, first_cube = int(n), int(depth), int(seed) & (2 ** 64 - 1), int(first_cube) & (2 ** 64 - 1)
	if host_out:
		if start is not None:
			raise ValueError("host_out does not take start states")
		N.require_cuda()
		out = np.empty((n, *shape()), dtype=np.int8)
		N.check(N.lib.rbh_scramble_seeded(_rep(), seed, first_cube, out.ctypes.data, n, depth))
		return out
	st = None
	if start is not None:
		st, _, _ = _states_in(start)
		if st.shape[0] != n:
			raise IndexError(f"expected {n} start states, got {st.shape[0]}")
	out = torch.empty(n, *shape(), dtype=torch.int8, device=_dev())
	N.check(N.lib.rb_scramble_seeded(_rep(), seed, first_cube, N.ptr(st), N.ptr(out), n, depth, N.stream_handle()))
	return out


def seeded_actions(n: int, depth: int, seed: int, first_cube: int = 0) -> torch.Tensor:
	"""The action indices `scramble_seeded` applies: uint8 (n, depth) CUDA tensor."""
	a = torch.empty(int(n), int(depth), dtype=torch.uint8, device=_dev())
	N.check(N.lib.rb_seeded_actions(int(seed) & (2 ** 64 - 1), int(first_cube) & (2 ** 64 - 1), N.ptr(a), int(n), int(depth), N.stream_handle()))
	return a


def pack_actions(actions: np.ndarray) -> np.ndarray:
	"""(n, depth) action indices -> (n, (depth + 1) // 2) packed bytes, two moves each: p = a[2k] + 13 * a[2k+1] (12 = no second
	move) -- the format of `scramble_batch_packed` / `rbh_scramble_packed` (half the bytes over PCIe)."""
	a = np.asarray(actions)
	if a.ndim != 2:
		raise IndexError("actions must be (n, depth)")
	if a.size and (a.min() < 0 or a.max() > 11):
		raise IndexError("action index out of range")
	a = a.astype(np.uint8)
	if a.shape[1] % 2:
		a = np.concatenate([a, np.full((len(a), 1), 12, np.uint8)], axis=1)
	return np.ascontiguousarray(a[:, 0::2] + 13 * a[:, 1::2])


def scramble_batch_packed(packed: np.ndarray, depth: int) -> np.ndarray:
	"""`scramble_batch` on packed host actions (see `pack_actions`); numpy in, numpy out through the host-buffer C-ABI call."""
	N.require_cuda()
	p = np.ascontiguousarray(packed, dtype=np.uint8)
	if p.ndim != 2 or p.shape[1] != (depth + 1) // 2:
		raise IndexError(f"packed actions must be (n, {(depth + 1) // 2})")
	if p.size:
		hi, lo = p // 13, p % 13
		ok = (lo < 12).all() and (hi[:, :depth // 2] < 12).all() and (hi <= 12).all() and (depth % 2 == 0 or (hi[:, -1] == 12).all())
		if not ok:
			raise IndexError("packed action byte out of range")
	out = np.empty((len(p), *shape()), dtype=np.int8)
	N.check(N.lib.rbh_scramble_packed(_rep(), p.ctypes.data, out.ctypes.data, len(p), int(depth)))
	return out


def scramble(depth: int, force_not_solved=False):
	"""Random scramble of one cube; same draw order as the reference (faces then dirs, cube.py:206-216)."""
	faces = np.random.randint(6, size=(depth,))
	dirs = np.random.randint(2, size=(depth,))
	state = scramble_batch(_actions_u8(faces, dirs)[None])[0] if depth else get_solved()
	if force_not_solved and is_solved(state) and depth != 0:
		return scramble(depth, True)
	return state, faces, dirs


def sequence_scrambler_from(faces, dirs, with_solved: bool, device_out: bool = False, with_oh: bool = True,
							with_flags: bool = False, oh_dtype=torch.float32):
	"""`sequence_scrambler` with the (depth, games) draws supplied by the caller.  `with_oh=False` skips the one-hot
	(states-only fast kernel); `with_flags` also returns the solved flag of every emitted state."""
	faces, dirs = np.asarray(faces), np.asarray(dirs)
	depth, games = faces.shape
	a = torch.from_numpy(_actions_u8(faces, dirs)).to(_dev())
	n = games * depth
	states = torch.empty(n, *shape(), dtype=torch.int8, device=a.device)
	oh = torch.empty(n, get_oh_shape(), dtype=oh_dtype, device=a.device) if with_oh else None
	flags = torch.empty(n, dtype=torch.uint8, device=a.device) if with_flags else None
	N.check(_oh_fn("rb_sequence_scramble", oh_dtype)(_rep(), N.ptr(a), None, games, depth, int(bool(with_solved)), N.ptr(states), N.ptr(oh),
									   N.ptr(flags), N.stream_handle()))
	res = [states if device_out else states.cpu().numpy(), oh]
	if with_flags:
		res.append(flags.bool() if device_out else flags.bool().cpu().numpy())
	return tuple(res)


def sequence_scrambler(games: int, depth: int, with_solved: bool):
	"""Out-of-place scrambler returning every state of every game, game-major, and their one-hot
	(cube.py:218-234).  Draw order as the reference: faces (depth, games), then dirs."""
	faces = np.random.randint(0, 6, (depth, games))
	dirs = np.random.randint(0, 2, (depth, games))
	return sequence_scrambler_from(faces, dirs, with_solved)


# ---- 6x3x3 views (cube.py:149-173, 279-307, 382-388): presentation only, host side ------------------
_CORNER_633 = (
	((F, 0, 0), (L, 0, 2), (T, 2, 0)), ((F, 2, 0), (D, 0, 0), (L, 2, 2)), ((F, 2, 2), (R, 2, 0), (D, 0, 2)),
	((F, 0, 2), (T, 2, 2), (R, 0, 0)), ((B, 0, 2), (T, 0, 0), (L, 0, 0)), ((B, 2, 2), (L, 2, 0), (D, 2, 0)),
	((B, 2, 0), (D, 2, 2), (R, 2, 2)), ((B, 0, 0), (R, 0, 2), (T, 0, 2)),
)
_EDGE_633 = (
	((F, 0, 1), (T, 2, 1)), ((F, 1, 0), (L, 1, 2)), ((F, 2, 1), (D, 0, 1)), ((F, 1, 2), (R, 1, 0)),
	((T, 1, 0), (L, 0, 1)), ((D, 1, 0), (L, 2, 1)), ((D, 1, 2), (R, 2, 1)), ((T, 1, 2), (R, 0, 1)),
	((B, 0, 1), (T, 0, 1)), ((B, 1, 2), (L, 1, 0)), ((B, 2, 1), (D, 2, 1)), ((B, 1, 0), (R, 1, 2)),
)
_RING_CELLS = np.array([0, 3, 6, 7, 8, 5, 2, 1])
_RING_SHIFT = np.array([0, 6, 6, 4, 2, 4])


def as633(state) -> np.ndarray:
	"""6x3x3 sticker colours, face order F, B, T, D, L, R (cube.py:149-154)."""
	if isinstance(state, torch.Tensor):
		state = state.cpu().numpy()
	out = np.repeat(np.arange(6), 9).reshape(6, 3, 3)
	if _is2024:
		for i in range(8):
			pos, ori = int(state[i]) // 3, int(state[i]) % 3
			ori = -ori if pos in (0, 2, 5, 7) else ori
			for cell, col in zip(_CORNER_633[pos], np.roll([c[0] for c in _CORNER_633[i]], ori)):
				out[cell] = col
		for i in range(12):
			pos, ori = int(state[i + 8]) // 2, int(state[i + 8]) % 2
			for cell, col in zip(_EDGE_633[pos], np.roll([c[0] for c in _EDGE_633[i]], ori)):
				out[cell] = col
		return out
	colours = np.where(state == 1)[2].reshape(6, 8)
	flat = out.reshape(6, 9)
	for f in range(6):
		flat[f, _RING_CELLS] = np.roll(colours[f], -_RING_SHIFT[f])
	return flat.reshape(6, 3, 3)


def as69(state) -> np.ndarray:
	return as633(state).reshape((6, 9))


def stringify(state) -> str:
	s = as633(state)
	canvas = np.full((9, 12), " ", dtype="<U1")
	for f, (r, c) in {T: (0, 1), L: (1, 0), F: (1, 1), R: (1, 2), B: (1, 3), D: (2, 1)}.items():
		canvas[3 * r:3 * r + 3, 3 * c:3 * c + 3] = s[f].astype(str)
	return "\n".join(" ".join(row) for row in canvas)
