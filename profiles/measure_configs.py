#!/usr/bin/env python
"""Times every kernel of the hot path at the BASELINE.json config sizes with CUDA events and prints
one JSON line per kernel: algorithmic bytes per launch, median ms, achieved GB/s and the fraction of the measured HBM
copy peak.  Not the bench (bench.py is); this is the per-kernel table DESIGN.md quotes.

  python profiles/measure_configs.py [--quick] [--only NAME_SUBSTR] > gpurun_out/configs.jsonl
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from rl_rubiks_b200 import _native as N, adi, cube, frontier  # noqa: E402

PEAK = 6499.0
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
	PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


ONCE = False       # --once: a single launch per kernel (for ncu captures)


def timeit(fn, reps=7, warm=2):
	if ONCE:
		reps, warm = 1, 0
	for _ in range(warm):
		fn()
	ms = []
	for _ in range(reps):
		FLUSH.zero_()
		a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		a.record(); fn(); b.record()
		torch.cuda.synchronize()
		ms.append(a.elapsed_time(b))
	return float(np.median(ms)), float(np.min(ms))


def report(name, nbytes, units, unit_name, fn, **extra):
	med, best = timeit(fn)
	gbs = nbytes / (med * 1e-3) / 1e9
	row = {"kernel": name, "bytes": int(nbytes), "ms": round(med, 4), "ms_best": round(best, 4), "GBps": round(gbs, 1),
		   "frac_hbm": round(gbs / PEAK, 4), unit_name + "_per_s": units / (med * 1e-3), **extra}
	print(json.dumps(row), flush=True)
	return row


def scrambled_2024(n, depth=30, seed=0):
	g = torch.Generator(device=dev); g.manual_seed(seed)
	acts = torch.randint(0, 12, (n, depth), dtype=torch.uint8, device=dev, generator=g)
	out = torch.empty(n, 20, dtype=torch.int8, device=dev)
	N.check(N.lib.rb_scramble(N.REP_2024, N.ptr(acts), depth, 1, None, N.ptr(out), n, depth, N.stream_handle()))
	return out


def scrambled_686(n, depth=30, seed=0):
	g = torch.Generator(device=dev); g.manual_seed(seed)
	acts = torch.randint(0, 12, (n, depth), dtype=torch.uint8, device=dev, generator=g)
	out = torch.empty(n, 6, 8, 6, dtype=torch.int8, device=dev)
	N.check(N.lib.rb_scramble(N.REP_686, N.ptr(acts), depth, 1, None, N.ptr(out), n, depth, N.stream_handle()))
	return out


def main():
	ap = argparse.ArgumentParser()
	ap.add_argument("--quick", action="store_true")
	ap.add_argument("--only", default="")
	ap.add_argument("--once", action="store_true", help="one launch per kernel, no warm-up (use under ncu)")
	args = ap.parse_args()
	global ONCE
	ONCE = args.once
	sh = N.stream_handle()
	q = 4 if args.quick else 1
	want = lambda name: args.only in name
	g = torch.Generator(device=dev); g.manual_seed(1)

	# ---- calibration: what a pure streaming write / copy reaches on this GPU (torch library kernels, not ours) ----
	if want("calibration"):
		buf = torch.empty(2 << 30, dtype=torch.float32, device=dev)
		report("calibration: torch fill_ 8 GiB (write only)", buf.numel() * 4, buf.numel() * 4, "bytes", lambda: buf.fill_(1.0))
		half = buf.numel() // 2
		report("calibration: torch copy_ 4 GiB -> 4 GiB (read + write)", buf.numel() * 4, buf.numel() * 4, "bytes",
			   lambda: buf[:half].copy_(buf[half:]))
		del buf

	# ---- C2 raw scramble + companions (20x24) ----
	if want("scramble_2024"):
		n, depth = (1 << 24) // q, 100
		acts = torch.randint(0, 12, (n, depth), dtype=torch.uint8, device=dev, generator=g)
		out = torch.empty(n, 20, dtype=torch.int8, device=dev)
		report("scramble_2024 (C2: n x 100 moves, final state)", n * 120, n * depth, "moves",
			   lambda: N.check(N.lib.rb_scramble(N.REP_2024, N.ptr(acts), depth, 1, None, N.ptr(out), n, depth, sh)), n=n)
		acts_t = acts.t().contiguous()
		report("scramble_2024 move-major actions [100][n] (the reference's draw shape)", n * 120, n * depth, "moves",
			   lambda: N.check(N.lib.rb_scramble(N.REP_2024, N.ptr(acts_t), 1, n, None, N.ptr(out), n, depth, sh)), n=n)
		del acts_t
		report("scramble_seeded_2024 (moves drawn in the kernel: 20 B per cube)", n * 20, n * depth, "moves",
			   lambda: N.check(N.lib.rb_scramble_seeded(N.REP_2024, 7, 0, None, N.ptr(out), n, depth, sh)), n=n)
		packed = (acts[:, 0::2] + 13 * acts[:, 1::2]).contiguous()
		report("unpack_actions (two moves per byte -> action bytes)", n * 150, n * depth, "moves",
			   lambda: N.check(N.lib.rb_unpack_actions(N.ptr(packed), N.ptr(acts), n, depth, sh)), n=n)
		del acts, out, packed
	if want("multi_rotate_2024"):
		n = (1 << 24) // q
		s = scrambled_2024(n)
		a = torch.randint(0, 12, (n,), dtype=torch.uint8, device=dev, generator=g)
		out = torch.empty_like(s)
		report("multi_rotate_2024", n * 41, n, "states",
			   lambda: N.check(N.lib.rb_multi_rotate(N.REP_2024, N.ptr(s), N.ptr(a), None, N.ptr(out), n, sh)), n=n)
		fl = torch.empty(n, dtype=torch.uint8, device=dev)
		report("multi_is_solved_2024", n * 21, n, "states",
			   lambda: N.check(N.lib.rb_multi_is_solved(N.REP_2024, N.ptr(s), N.ptr(fl), n, sh)), n=n)
		del s, a, out, fl
	if want("as_oh_2024"):
		n = (1 << 22) // q
		s = scrambled_2024(n)
		oh = torch.empty(n, 480, dtype=torch.float32, device=dev)
		report("as_oh_2024", n * 1940, n, "states", lambda: N.check(N.lib.rb_as_oh(N.REP_2024, N.ptr(s), N.ptr(oh), n, sh)), n=n)
		ohb = torch.empty(n, 480, dtype=torch.bfloat16, device=dev)
		report("as_oh_2024 bf16 rows (opt-in)", n * 980, n, "states", lambda: N.check(N.lib.rb_as_oh_bf16(N.REP_2024, N.ptr(s), N.ptr(ohb), n, sh)), n=n)
		del s, oh, ohb
	if want("expand12_2024"):
		n = (1 << 18) // q
		s = scrambled_2024(n)
		ch = torch.empty(12 * n, 20, dtype=torch.int8, device=dev)
		oh = torch.empty(12 * n, 480, dtype=torch.float32, device=dev)
		fl = torch.empty(12 * n, dtype=torch.uint8, device=dev)
		report("expand12_2024 states+oh+solved", n * (20 + 12 * 1941), 12 * n, "children",
			   lambda: N.check(N.lib.rb_expand12(N.REP_2024, N.ptr(s), N.ptr(ch), N.ptr(oh), N.ptr(fl), n, sh)), n=n)
		del oh
		n2 = (1 << 22) // q
		s2 = scrambled_2024(n2)
		ch2 = torch.empty(12 * n2, 20, dtype=torch.int8, device=dev)
		report("expand12_2024 states only", n2 * (20 + 240), 12 * n2, "children",
			   lambda: N.check(N.lib.rb_expand12(N.REP_2024, N.ptr(s2), N.ptr(ch2), None, None, n2, sh)), n=n2)
		del s, ch, fl, s2, ch2
	if want("sequence_2024"):
		games, depth = 7500, 30
		a = torch.randint(0, 12, (depth, games), dtype=torch.uint8, device=dev, generator=g)
		n = games * depth
		st = torch.empty(n, 20, dtype=torch.int8, device=dev)
		oh = torch.empty(n, 480, dtype=torch.float32, device=dev)
		report("sequence_2024 states+oh (7500 x 30)", n * (1 + 20 + 1920), n, "states",
			   lambda: N.check(N.lib.rb_sequence_scramble(N.REP_2024, N.ptr(a), None, games, depth, 1, N.ptr(st), N.ptr(oh), None, sh)), n=n)
		gs, ds = 1000, 25
		a_s = torch.randint(0, 12, (ds, gs), dtype=torch.uint8, device=dev, generator=g)
		report("sequence_2024 states+oh (1000 x 25)", gs * ds * (1 + 20 + 1920), gs * ds, "states",
			   lambda: N.check(N.lib.rb_sequence_scramble(N.REP_2024, N.ptr(a_s), None, gs, ds, 1, N.ptr(st), N.ptr(oh), None, sh)), n=gs * ds)
		games2, depth2 = (1 << 20) // q, 100
		a2 = torch.randint(0, 12, (depth2, games2), dtype=torch.uint8, device=dev, generator=g)
		st2 = torch.empty(games2 * depth2, 20, dtype=torch.int8, device=dev)
		report("sequence_2024 states only (2^20 games x 100)", games2 * depth2 * 21, games2 * depth2, "moves",
			   lambda: N.check(N.lib.rb_sequence_scramble(N.REP_2024, N.ptr(a2), None, games2, depth2, 0, N.ptr(st2), None, None, sh)), n=games2)
		del a, st, oh, a2, st2

	# ---- C1 ADI ----
	for games, depth in ((1000, 25), (7500, 30)):
		if not want("adi"):
			break
		gen = adi.ADIGenerator(games, depth, "lapanfix", keep_states=True)
		gen.set_actions(torch.randint(0, 12, (depth, games), dtype=torch.uint8, device=dev, generator=g))
		nst = games * depth
		values = torch.randn(12 * nst, device=dev)
		gen_bytes = 1920 * 13 * nst + 20 * nst + 13 * nst + nst
		tgt_bytes = 4 * 12 * nst + 12 * nst + nst + 12 * nst + 4 * nst
		report(f"adi_generate_2024 ({games} x {depth})", gen_bytes, nst, "samples", gen.generate, n=nst)
		report(f"adi_targets+loss_weights ({games} x {depth})", tgt_bytes, nst, "samples", lambda: gen.targets(values, 0.3), n=nst)
		del gen
		genb = adi.ADIGenerator(games, depth, "lapanfix", keep_states=True, oh_dtype=torch.bfloat16)
		genb.set_actions(torch.randint(0, 12, (depth, games), dtype=torch.uint8, device=dev, generator=g))
		report(f"adi_generate_2024 bf16 rows (opt-in) ({games} x {depth})", 960 * 13 * nst + 20 * nst + 13 * nst + nst, nst, "samples", genb.generate, n=nst)
		del genb, values

	# ---- C3 6x8x6 ----
	if want("686"):
		n = (1 << 22) // q
		s = scrambled_686(n)
		a = torch.randint(0, 12, (n,), dtype=torch.uint8, device=dev, generator=g)
		out = torch.empty_like(s)
		report("multi_rotate_686 (C3)", n * 577, n, "states",
			   lambda: N.check(N.lib.rb_multi_rotate(N.REP_686, N.ptr(s), N.ptr(a), None, N.ptr(out), n, sh)), n=n)
		del out
		oh = torch.empty(n, 288, dtype=torch.float32, device=dev)
		report("as_oh_686 (C3)", n * 1440, n, "states", lambda: N.check(N.lib.rb_as_oh(N.REP_686, N.ptr(s), N.ptr(oh), n, sh)), n=n)
		del oh
		fl = torch.empty(n, dtype=torch.uint8, device=dev)
		report("multi_is_solved_686", n * 289, n, "states", lambda: N.check(N.lib.rb_multi_is_solved(N.REP_686, N.ptr(s), N.ptr(fl), n, sh)), n=n)
		ch = torch.empty(12 * n, 6, 8, 6, dtype=torch.int8, device=dev)
		report("expand12_686 states only (C3)", n * 3744, 12 * n, "children",
			   lambda: N.check(N.lib.rb_expand12(N.REP_686, N.ptr(s), N.ptr(ch), None, None, n, sh)), n=n)
		del ch, fl
		n4 = (1 << 17) // q
		s4 = s[:n4].contiguous()
		ch4 = torch.empty(12 * n4, 6, 8, 6, dtype=torch.int8, device=dev)
		oh4 = torch.empty(12 * n4, 288, dtype=torch.float32, device=dev)
		fl4 = torch.empty(12 * n4, dtype=torch.uint8, device=dev)
		report("expand12_686 states+oh+solved", n4 * (288 + 12 * (288 + 1152 + 1)), 12 * n4, "children",
			   lambda: N.check(N.lib.rb_expand12(N.REP_686, N.ptr(s4), N.ptr(ch4), N.ptr(oh4), N.ptr(fl4), n4, sh)), n=n4)
		del s4, ch4, oh4, fl4
		games, gd = 1000, 25
		ga = torch.randint(0, 12, (gd, games), dtype=torch.uint8, device=dev, generator=g)
		nst = games * gd
		st = torch.empty(nst, 6, 8, 6, dtype=torch.int8, device=dev)
		oh = torch.empty(nst, 288, dtype=torch.float32, device=dev)
		ss = torch.empty(nst, dtype=torch.uint8, device=dev)
		ch = torch.empty(12 * nst, 6, 8, 6, dtype=torch.int8, device=dev)
		coh = torch.empty(12 * nst, 288, dtype=torch.float32, device=dev)
		sc = torch.empty(12 * nst, dtype=torch.uint8, device=dev)
		report("sequence_686 states+oh (1000 x 25)", nst * (1 + 288 + 1152), nst, "states",
			   lambda: N.check(N.lib.rb_sequence_scramble(N.REP_686, N.ptr(ga), None, games, gd, 1, N.ptr(st), N.ptr(oh), None, sh)), n=nst)
		report("adi_generate_686 (1000 x 25)", nst * (1 + 288 + 1152 + 1 + 12 * (288 + 1152 + 1)), nst, "samples",
			   lambda: N.check(N.lib.rb_adi_generate(N.REP_686, N.ptr(ga), None, games, gd, 1, N.ptr(st), N.ptr(oh), N.ptr(ch), N.ptr(coh), N.ptr(ss), N.ptr(sc), sh)), n=nst)
		del ga, st, oh, ss, ch, coh, sc
		n3, depth = (1 << 20) // q, 100
		acts = torch.randint(0, 12, (n3, depth), dtype=torch.uint8, device=dev, generator=g)
		o3 = torch.empty(n3, 6, 8, 6, dtype=torch.int8, device=dev)
		report("scramble_686 (n x 100 moves)", n3 * (100 + 288), n3 * depth, "moves",
			   lambda: N.check(N.lib.rb_scramble(N.REP_686, N.ptr(acts), depth, 1, None, N.ptr(o3), n3, depth, sh)), n=n3)
		del s, a, acts, o3

	# ---- C4/C5 frontier ----
	if want("frontier"):
		for is2024 in (True, False):
			tag = "2024" if is2024 else "686"
			n = (1 << 22) // q if is2024 else (1 << 20) // q
			s = scrambled_2024(n, 40) if is2024 else scrambled_686(n, 40)
			hs = frontier.StateHashSet(4 * n, is2024)

			def ins():
				hs.clear()
				hs.insert_unique(s)
			sb = 20 if is2024 else 288
			report(f"hashset_insert_unique_{tag} (n fresh states, load 1/4)", n * (sb + 32 + 6 + 4), n, "states", ins, n=n)
			del hs, s
		import time
		for depth in ((6, 7) if not args.quick else (6,)):
			torch.cuda.synchronize()
			t0 = time.perf_counter()
			counts, hs = frontier.bfs_layers(depth, is2024=True, capacity=1 << 25)
			torch.cuda.synchronize()
			dt = time.perf_counter() - t0
			gen_children = 12 * sum(counts[:-1])
			print(json.dumps({"kernel": f"bfs_layers_2024 depth {depth} (C5, wall clock incl. allocations)", "ms": dt * 1e3, "counts": counts,
							  "children_per_s": gen_children / dt, "unique": sum(counts)}), flush=True)
			del hs
		# A* shaped step: 700 parents -> 8400 children against a table holding ~170k states
		hs = frontier.StateHashSet(1 << 19, True)
		hs.insert_unique(scrambled_2024(170000, 30, 5))
		par = scrambled_2024(700, 30, 6)
		med, best = timeit(lambda: hs.expand(par, flags=True, index=True), reps=20)
		print(json.dumps({"kernel": "frontier_expand_2024 A* step (700 parents, 8400 children, incl. torch allocs)", "ms": med, "ms_best": best,
						  "children_per_s": 8400 / (med * 1e-3)}), flush=True)


if __name__ == "__main__":
	main()
