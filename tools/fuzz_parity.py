#!/usr/bin/env python
"""Randomised parity soak: random sizes, depths, layouts, alignments and representations through the C ABI / the Python mirror,
every result compared bit for bit with the oracle.  Complements tests/ (fixed cases) -- run it for as long as you like:

  python tools/fuzz_parity.py [seconds] [seed]        # default 120 s; prints one line per case family and a summary

Exits 1 on the first mismatch (the case's parameters are printed, and the seed reproduces it)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cube_oracle as O  # noqa: E402
from rl_rubiks_b200 import _native as N  # noqa: E402
from rl_rubiks_b200 import adi, cube  # noqa: E402
from rl_rubiks_b200.frontier import StateHashSet  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else int(time.time()) & 0xffff
rng = np.random.RandomState(seed)
dev = torch.device("cuda", 0)
counts, t_end = {}, time.time() + budget


def fail(family, **kw):
	print(f"MISMATCH in {family}: seed {seed} params {kw}", flush=True)
	sys.exit(1)


def rand_n():
	return int(rng.choice([0, 1, 2, 31, 32, 33, 63, 64, 65, rng.randint(1, 300), rng.randint(300, 5000), rng.randint(5000, 40000)]))


def rand_states(n, is2024, depth=None):
	depth = rng.randint(0, 40) if depth is None else depth
	f, d = rng.randint(0, 6, (n, depth)), rng.randint(0, 2, (n, depth))
	return O.scramble_many(f, d, is2024) if n else np.zeros((0,) + tuple(O.solved(is2024).shape), np.int8)


def misaligned(t: torch.Tensor, off: int) -> torch.Tensor:
	"""Copy of t (contiguous) living `off` bytes past a 256-byte aligned allocation."""
	raw = torch.empty(t.numel() * t.element_size() + 64, dtype=torch.uint8, device=t.device)
	v = raw[off:off + t.numel() * t.element_size()].view(t.dtype).view(t.shape)
	v.copy_(t)
	return v


def case_scramble():
	is2024 = rng.rand() < 0.7
	cube.set_is2024(bool(is2024))
	n = min(rand_n(), 6000 if is2024 else 600)
	depth = int(rng.choice([0, 1, 2, 3, 19, 20, 21, 23, 24, 25, 47, 48, 64, 96, 100, 128, 256, rng.randint(1, 420), rng.randint(400, 1000)]))
	if depth > 420:
		n = min(n, 700)
	acts = rng.randint(0, 12, (n, depth)).astype(np.uint8)
	f, d = O.indices_to_actions(acts)
	start = rand_states(n, is2024, 5) if rng.rand() < 0.3 else None
	want = O.scramble_many(f, d, is2024) if start is None else None
	if start is not None:
		want = start.copy()
		for m in range(depth):
			want = O.multi_rotate(want, f[:, m], d[:, m], is2024)
	layout = rng.choice(["cube", "move", "strided"])
	a = torch.from_numpy(acts).to(dev)
	out = torch.empty(n, *cube.shape(), dtype=torch.int8, device=dev)
	st = torch.from_numpy(start).to(dev) if start is not None else None
	if layout == "cube":
		off = int(rng.choice([0, 0, 1, 4, 16]))
		a = misaligned(a, off) if n * depth else a
		sc, sm, buf = depth, 1, a
	elif layout == "move":
		buf = a.t().contiguous()
		sc, sm = 1, n
	else:
		pad = int(rng.randint(1, 9))
		buf = torch.zeros(n, depth + pad, dtype=torch.uint8, device=dev)
		buf[:, :depth] = a
		sc, sm = depth + pad, 1
	N.check(N.lib.rb_scramble(cube._rep(), N.ptr(buf), sc, sm, N.ptr(st), N.ptr(out), n, depth, N.stream_handle()))
	if not (out.cpu().numpy() == want).all():
		fail("scramble", is2024=is2024, n=n, depth=depth, layout=layout, start=start is not None)
	if is2024 and start is None and n:
		p = cube.pack_actions(acts)
		got = cube.scramble_batch_packed(p, depth)
		if not (got == want).all():
			fail("scramble_packed", n=n, depth=depth)


def case_seeded():
	cube.set_is2024(bool(rng.rand() < 0.8))
	n, depth = min(rand_n(), 3000), int(rng.choice([0, 1, 2, 3, 4, 5, 24, 25, 26, 99, 100, 101, rng.randint(1, 300), rng.randint(300, 1000)]))
	if depth > 300:
		n = min(n, 700)
	sd, first = int(rng.randint(0, 2 ** 31)), int(rng.choice([0, 1, 2 ** 32 - 5, rng.randint(0, 2 ** 40)]))
	got = cube.scramble_seeded(n, depth, sd, first)
	acts = O.seeded_actions(sd, first, n, depth)
	f, d = O.indices_to_actions(acts)
	want = O.scramble_many(f, d, cube.get_is2024())
	if not (got.cpu().numpy() == want).all():
		fail("seeded", is2024=cube.get_is2024(), n=n, depth=depth, seed=sd, first=first)
	if n and not (cube.seeded_actions(n, depth, sd, first).cpu().numpy() == acts).all():
		fail("seeded_actions", n=n, depth=depth, seed=sd, first=first)


def case_rotate_solved_oh():
	is2024 = bool(rng.rand() < 0.6)
	cube.set_is2024(is2024)
	n = min(rand_n(), 20000 if is2024 else 1500)
	s = rand_states(n, is2024)
	f, d = rng.randint(0, 6, n), rng.randint(0, 2, n)
	off = int(rng.choice([0, 0, 4, 1, 3])) if is2024 else 0
	st = misaligned(torch.from_numpy(s).to(dev), off) if n else torch.from_numpy(s).to(dev)
	got = cube.multi_rotate(st, torch.from_numpy(f.astype(np.uint8)).to(dev), torch.from_numpy(d.astype(np.uint8)).to(dev))
	if not (got.cpu().numpy() == O.multi_rotate(s, f, d, is2024)).all():
		fail("multi_rotate", is2024=is2024, n=n, off=off)
	if n:
		s[rng.randint(0, n, max(1, n // 7))] = O.solved(is2024)
		st = misaligned(torch.from_numpy(s).to(dev), off)
	if not (cube.multi_is_solved(st).cpu().numpy() == O.multi_is_solved(s, is2024)).all():
		fail("multi_is_solved", is2024=is2024, n=n, off=off)
	m = min(n, 3000)
	dt = torch.float32 if rng.rand() < 0.7 else torch.bfloat16
	oh = cube.as_oh(st[:m], dtype=dt)
	if m and not (oh.float().cpu().numpy() == O.as_oh(s[:m], is2024)).all():
		fail("as_oh", is2024=is2024, n=m, off=off, dtype=str(dt))
	ch, coh, fl = cube.expand12(st[:m], with_oh=True, with_solved=True, oh_dtype=dt)
	wch = O.expand12(s[:m], is2024)
	if not ((ch.cpu().numpy() == wch).all() and (coh.float().cpu().numpy() == O.as_oh(wch, is2024)).all()
			and (fl.cpu().numpy() == O.multi_is_solved(wch, is2024)).all()):
		fail("expand12", is2024=is2024, n=m, off=off, dtype=str(dt))
	ch2 = cube.expand12(st[:m])
	if not (ch2.cpu().numpy() == wch).all():
		fail("expand12_states", is2024=is2024, n=m, off=off)


def case_sequence_adi():
	is2024 = bool(rng.rand() < 0.7)
	cube.set_is2024(is2024)
	games, depth = int(rng.choice([1, 2, 31, 33, rng.randint(1, 200)])), int(rng.choice([1, 2, 7, 8, 9, 25, 30, rng.randint(1, 60)]))
	ws = bool(rng.randint(0, 2))
	f, d = rng.randint(0, 6, (depth, games)), rng.randint(0, 2, (depth, games))
	want_s, want_oh = O.sequence_scrambler(f, d, ws, is2024)
	with_oh = bool(rng.randint(0, 2))
	res = cube.sequence_scrambler_from(f, d, ws, with_oh=with_oh, with_flags=True)
	if not ((res[0] == want_s).all() and (res[2] == O.multi_is_solved(want_s, is2024)).all()):
		fail("sequence", is2024=is2024, games=games, depth=depth, ws=ws, with_oh=with_oh)
	if with_oh and not (res[1].cpu().numpy() == want_oh).all():
		fail("sequence_oh", is2024=is2024, games=games, depth=depth, ws=ws)
	method = str(rng.choice(["paper", "lapanfix", "schultzfix", "reward0"]))
	alpha = float(rng.choice([0.0, 0.3, 1.0]))
	w = rng.randint(-3, 4, cube.get_oh_shape()).astype(np.float32)
	quant = float(rng.choice([1.0, 4.0, 16.0]))

	def vf_np(x):
		return np.floor((x @ w) / quant).astype(np.float32)
	wt = torch.from_numpy(w).to(dev)

	class Net(torch.nn.Module):
		def forward(self, x, policy=True, value=True):
			return torch.floor((x.float() @ wt) / quant).unsqueeze(1)
	want = O.adi_traindata(f, d, vf_np, method, alpha, is2024)
	got = adi.adi_traindata(Net(), games, depth, method, alpha, faces=f, dirs=d)
	for a, k in zip(got, ("oh_states", "policy_targets", "value_targets", "loss_weights")):
		if not np.array_equal(a.cpu().numpy(), np.asarray(want[k])):
			fail("adi_traindata", is2024=is2024, games=games, depth=depth, method=method, alpha=alpha, output=k)


def case_hashset():
	is2024 = bool(rng.rand() < 0.7)
	cube.set_is2024(is2024)
	hs, ref = StateHashSet(1 << int(rng.randint(4, 12)), is2024), O.SeenSet()
	pool = rand_states(int(rng.randint(1, 800)), is2024, int(rng.randint(1, 6)))
	for _ in range(int(rng.randint(1, 5))):
		batch = pool[rng.randint(0, len(pool), int(rng.randint(0, 2500 if is2024 else 400)))]
		got, want = hs.insert_unique(batch), ref.insert_unique(batch)
		if not all(np.array_equal(a, b) for a, b in zip(got, want)) or len(hs) != len(ref):
			fail("hashset_insert", is2024=is2024, batch=len(batch), size=len(ref))
		q = pool[rng.randint(0, len(pool), 100)]
		if not np.array_equal(hs.lookup(q), ref.lookup(q)):
			fail("hashset_lookup", is2024=is2024)
	fr = pool[rng.randint(0, len(pool), int(rng.randint(1, 300 if is2024 else 40)))]
	out = hs.expand(torch.from_numpy(fr).to(dev), flags=True, index=True)
	ch = O.expand12(fr, is2024)
	seen, first, idx = ref.insert_unique(ch)
	n_new = int(out["n_new"].item())
	ok = n_new == int(first.sum()) and np.array_equal(out["seen"].cpu().numpy().astype(bool), seen) and \
		np.array_equal(out["first"].cpu().numpy().astype(bool), first) and np.array_equal(out["index"].cpu().numpy(), idx) and \
		np.array_equal(out["next"][:n_new].cpu().numpy(), ch[first]) and \
		np.array_equal(out["parent"][:n_new].cpu().numpy(), np.nonzero(first)[0] // 12) and \
		np.array_equal(out["action"][:n_new].cpu().numpy(), np.nonzero(first)[0] % 12) and \
		np.array_equal(out["solved"][:n_new].cpu().numpy().astype(bool), O.multi_is_solved(ch[first], is2024))
	if not ok:
		fail("frontier_expand", is2024=is2024, frontier=len(fr), size=len(ref))


def case_convert():
	cube.set_is2024(True)
	n = rand_n() % 3000
	s = rand_states(n, True)
	s686 = cube.to_686(torch.from_numpy(s).to(dev))
	acts = rng.randint(0, 12, n)
	a_f, a_d = O.indices_to_actions(acts)
	# the two representations commute with a move: render(rotate(s)) == rotate(render(s))
	cube.set_is2024(False)
	lhs = cube.multi_rotate(s686, torch.from_numpy(a_f.astype(np.uint8)).to(dev), torch.from_numpy(a_d.astype(np.uint8)).to(dev))
	cube.set_is2024(True)
	rhs = cube.to_686(torch.from_numpy(O.multi_rotate(s, a_f, a_d, True)).to(dev))
	back = cube.to_2024(s686)
	if not (torch.equal(lhs, rhs) and (back.cpu().numpy() == s).all()):
		fail("convert", n=n)


FAMILIES = [case_scramble, case_seeded, case_rotate_solved_oh, case_sequence_adi, case_hashset, case_convert]
while time.time() < t_end:
	fn = FAMILIES[rng.randint(0, len(FAMILIES))]
	fn()
	counts[fn.__name__] = counts.get(fn.__name__, 0) + 1
cube.set_is2024(True)
print(f"fuzz_parity: seed {seed}, {sum(counts.values())} cases in {budget:.0f} s, no mismatch: {counts}")
