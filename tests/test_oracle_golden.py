"""Pins the CPU oracle (oracle/cube_oracle.py) against fixtures produced by running the
reference itself (tests/golden/make_golden.py) and against the literal vectors of the
reference's own tests (/root/reference/tests/test_cube.py, cited per test)."""
import heapq

import numpy as np
import pytest

from oracle import cube_oracle as O


# ---- tables --------------------------------------------------------
def test_delta_maps_equal_reference_and_frontend_json(golden):
	g = golden("tables")
	assert (O.DELTA_MAPS == g["delta_maps"]).all()
	# frontend/src/assets/maps.json holds the same tables as literals
	assert (O.DELTA_MAPS[0] == g["json_map_neg"]).all() and (O.DELTA_MAPS[1] == g["json_map_pos"]).all()
	for a in range(12):
		for k in range(2):
			assert sorted(O.LUT2024[a, k].tolist()) == list(range(24))


def test_solved_perm_and_action_helpers(golden):
	g = golden("tables")
	assert (O.solved_2024() == g["solved2024"]).all() and (O.solved_686() == g["solved686"]).all()
	assert (O.PERM686 == g["perm686"]).all()
	assert (O.iter_actions(2) == g["iter_actions2"]).all() and O.iter_actions(2).dtype == np.uint8
	f, d = O.indices_to_actions(np.arange(12))
	assert (f == g["idx2act_faces"]).all() and (d == g["idx2act_dirs"]).all()
	assert (O.rev_actions(np.arange(12)) == g["rev_actions"]).all()
	assert [O.rev_action(a) for a in range(12)] == g["rev_actions"].tolist()
	assert (np.stack([O.FACE_OF_ACTION, O.DIR_OF_ACTION], 1) == g["action_space"]).all()
	# literals of reference tests/test_cube.py:116-127
	assert O.iter_actions(2).tolist() == [[0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5] * 2, [1, 0] * 12]


# ---- dynamics (rows a2-a7) ------------------------------------------------------------
@pytest.mark.parametrize("tag,is2024", [("2024", True), ("686", False)])
def test_multi_rotate_expand_oh_solved(golden, tag, is2024):
	g = golden("dynamics")
	faces, dirs, steps = g[f"faces{tag}"], g[f"dirs{tag}"], g[f"steps{tag}"]
	n = faces.shape[1]
	s = np.repeat(O.solved(is2024)[None], n, 0)
	for d in range(len(faces)):
		s = O.multi_rotate(s, faces[d], dirs[d], is2024)
		assert s.dtype == np.int8 and (s == steps[d]).all()
	assert (O.scramble_many(faces.T, dirs.T, is2024) == steps[-1]).all()
	assert (O.scramble(faces[:, 3], dirs[:, 3], is2024) == steps[-1][3]).all()
	assert (O.rotate(steps[-2][5], faces[-1][5], dirs[-1][5], is2024) == steps[-1][5]).all()
	assert (O.expand12(s, is2024) == g[f"children{tag}"]).all()
	flags = O.multi_is_solved(np.concatenate([s, O.solved(is2024)[None]]), is2024)
	assert (flags == g[f"solved_flags{tag}"]).all() and flags[-1]
	oh = O.as_oh(s, is2024)
	assert oh.dtype == np.float32 and (oh == g[f"oh{tag}"]).all()
	assert (O.as_oh(s[0], is2024) == g[f"oh_single{tag}"]).all()
	assert (np.stack([O.as633(x, is2024) for x in s[:8]]) == g[f"as633_{tag}"]).all()
	assert [O.stringify(x, is2024) for x in s[:8]] == g[f"strings{tag}"].tolist()
	if is2024:
		assert (O.multi_act_2024(steps[-2], O.action_index(faces[-1], dirs[-1])) == steps[-1]).all()
	else:
		assert (O.as_correct_686(oh) == g["correct686"]).all()


SOLVED_STR = "\n".join([
	"      2 2 2            ", "      2 2 2            ", "      2 2 2            ",
	"4 4 4 0 0 0 5 5 5 1 1 1", "4 4 4 0 0 0 5 5 5 1 1 1", "4 4 4 0 0 0 5 5 5 1 1 1",
	"      3 3 3            ", "      3 3 3            ", "      3 3 3            "])
F_POS_STR = "\n".join([
	"      2 2 2            ", "      2 2 2            ", "      5 5 5            ",
	"4 4 2 0 0 0 3 5 5 1 1 1", "4 4 2 0 0 0 3 5 5 1 1 1", "4 4 2 0 0 0 3 5 5 1 1 1",
	"      4 4 4            ", "      3 3 3            ", "      3 3 3            "])
ALL12_STR = "\n".join([
	"      2 0 2            ", "      5 2 4            ", "      2 1 2            ",
	"4 2 4 0 2 0 5 2 5 1 2 1", "4 4 4 0 0 0 5 5 5 1 1 1", "4 3 4 0 3 0 5 3 5 1 3 1",
	"      3 1 3            ", "      5 3 4            ", "      3 0 3            "])


@pytest.mark.parametrize("is2024", [True, False])
def test_reference_test_cube_literals(is2024):
	"""/root/reference/tests/test_cube.py:26-92 (`_rotation_tests`), both representations."""
	s = O.solved(is2024)
	assert O.stringify(s, is2024) == SOLVED_STR
	for m, a in zip(((0, 1), (0, 0), (0, 1), (1, 1), (2, 0), (3, 0)), (False, True, False, False, False, False)):
		s = O.rotate(s, *m, is2024)
		assert O.is_solved(s, is2024) == a
	for m, a in zip(((3, 1), (2, 1), (1, 0), (0, 0)), (False, False, False, True)):
		s = O.rotate(s, *m, is2024)
		assert O.is_solved(s, is2024) == a
	assert O.stringify(O.rotate(O.solved(is2024), 0, 1, is2024), is2024) == F_POS_STR
	s = O.solved(is2024)
	for m in [(f, 0) for f in range(6)] + [(f, 1) for f in range(6)]:
		s = O.rotate(s, *m, is2024)
		assert not O.is_solved(s, is2024)
	assert O.stringify(s, is2024) == ALL12_STR


def test_reference_as_oh_and_as_correct_literals():
	"""tests/test_cube.py:129-139 (one-hot of solved) and :149-166 (as_correct literal)."""
	oh = O.as_oh_2024(O.solved_2024())
	want = np.zeros((20, 24), dtype=np.float32)
	want[np.arange(8), 3 * np.arange(8)] = 1
	want[np.arange(8, 20), 2 * np.arange(12)] = 1
	assert oh.shape == (1, 480) and (oh[0] == want.ravel()).all()
	s = O.rotate_686(O.rotate_686(O.solved_686(), 0, 1), 5, 0)
	lit = np.array([[1, 1, 1, 1, -1, -1, -1, 1], [-1, 1, 1, 1, 1, 1, -1, -1], [-1, -1, -1, -1, -1, 1, 1, 1],
					[-1, -1, -1, -1, -1, 1, 1, 1], [-1, 1, 1, 1, 1, 1, -1, -1], [1, 1, -1, -1, -1, 1, 1, 1]], dtype=np.float32)
	assert (O.as_correct_686(O.as_oh_686(s))[0] == lit).all()


def test_2024_and_686_agree_through_as633():
	"""SURVEY 8c KAT (ii): the two representations describe the same cube."""
	g = np.random.RandomState(0)
	a, b = O.solved_2024(), O.solved_686()
	for _ in range(200):
		f, d = int(g.randint(6)), int(g.randint(2))
		a, b = O.rotate_2024(a, f, d), O.rotate_686(b, f, d)
		assert (O.as633_2024(a) == O.as633_686(b)).all()


# ---- scramblers (rows a8, a9) ----------------------------------------------------------
@pytest.mark.parametrize("tag,is2024", [("2024", True), ("686", False)])
def test_sequence_scrambler_and_scramble(golden, tag, is2024):
	g = golden("scramblers")
	for games, depth, ws in ((3, 4, True), (5, 7, False), (4, 1, True)):
		key = f"{tag}_{games}_{depth}_{int(ws)}"
		np.random.seed(0)
		faces, dirs = O.draw_sequence_actions(games, depth)
		assert (faces == g[f"seq_faces_{key}"]).all() and (dirs == g[f"seq_dirs_{key}"]).all()
		states, oh = O.sequence_scrambler(faces, dirs, ws, is2024)
		assert states.dtype == np.int8 and (states == g[f"seq_states_{key}"]).all()
		assert (oh == g[f"seq_oh_{key}"]).all()
	if is2024:   # SURVEY 8c KAT (iii)
		assert g["seq_states_2024_3_4_1"][1].tolist() == [4, 16, 6, 9, 1, 13, 18, 21, 0, 10, 4, 6, 2, 18, 12, 14, 16, 8, 20, 22]
	s = O.scramble(g[f"scr_faces_{tag}"], g[f"scr_dirs_{tag}"], is2024)
	assert (s == g[f"scr_state_{tag}"]).all() and not O.is_solved(s, is2024)
	# tests/test_cube.py:103-114: undoing the scramble with reversed inverse moves solves it
	for f, d in zip(g[f"scr_faces_{tag}"][::-1], g[f"scr_dirs_{tag}"][::-1]):
		s = O.rotate(s, int(f), int(1 - d), is2024)
	assert O.is_solved(s, is2024)


# ---- ADI (row a10) ---------------------------------------------------------------------
def _fake_value_fn(w, quant=4.0):
	return lambda oh: np.floor((oh @ w) / np.float32(quant)).astype(np.float32)


@pytest.mark.parametrize("tag,is2024", [("2024", True), ("686", False)])
@pytest.mark.parametrize("method", O.REWARD_METHODS)
def test_adi_traindata(golden, tag, is2024, method):
	g = golden("adi")
	for ai in ((0, 1, 2) if is2024 else (1,)):
		key = f"{tag}_{method}_{ai}"
		r = O.adi_traindata(g[f"faces_{key}"], g[f"dirs_{key}"], _fake_value_fn(g[f"w_{key}"]), method, float(g[f"alpha_{key}"]), is2024)
		assert (r["values"] == g[f"values_{key}"]).all()
		assert (r["oh_states"] == g[f"oh_states_{key}"]).all()
		assert r["policy_targets"].dtype == np.int64 and (r["policy_targets"] == g[f"policy_{key}"]).all()
		assert r["value_targets"].dtype == np.float32 and (r["value_targets"] == g[f"value_{key}"]).all()
		assert r["loss_weights"].dtype == np.float32 and (r["loss_weights"] == g[f"lw_{key}"]).all()


def test_adi_all_tie_and_big_loss_weights(golden):
	g = golden("adi")
	for method in ("paper", "lapanfix"):
		r = O.adi_traindata(g[f"tie_faces_{method}"], g[f"tie_dirs_{method}"], lambda oh: np.zeros(len(oh), np.float32), method, 0.5)
		assert (r["policy_targets"] == g[f"tie_policy_{method}"]).all()
		assert (r["value_targets"] == g[f"tie_value_{method}"]).all()
		assert (r["loss_weights"] == g[f"tie_lw_{method}"]).all()
	for games, depth, alpha in ((1000, 25, 0.3), (7500, 30, 0.7), (17, 999, 0.05)):
		assert (O.adi_loss_weights(games, depth, alpha) == g[f"lwbig_{games}_{depth}"]).all()


def test_adi_targets_nan_counts_as_max():
	import torch
	v = np.array([0, 1, np.nan, 3, 2, np.nan, 0, 0, 0, 0, 0, 0] + list(range(12)), dtype=np.float32)
	p, val = O.adi_targets(v, np.zeros(24, bool), np.zeros(2, bool), "paper", 1)
	tp = torch.argmax(torch.from_numpy(v - 1).reshape(-1, 12), dim=1).numpy()
	assert (p == tp).all() and np.isnan(val[0]) and val[1] == 10


# ---- search frontier (rows a11-a13) ----------------------------------------------------
def test_bfs_layer_counts(golden):
	g = golden("search")
	counts, seen = O.bfs_layers(5)
	assert counts == g["bfs_layer_counts"].tolist() == [1, 12, 114, 1068, 10011, 93840]
	assert len(seen) == sum(counts)
	assert O.bfs_layer_counts_packed(5) == counts


def test_bfs_agent_len(golden):
	"""BFS.search (agents.py:96-123) stops at the first solved child; len = dict size then."""
	g = golden("search")
	start = g["bfs_start"]
	seen = O.SeenSet()
	seen.insert_unique(start[None])
	frontier, found, queue = start[None], False, None
	parent_of = {}
	states = {1: start}
	while not found:
		nxt = []
		for s in frontier:
			pi = seen.lookup(s[None])[0]
			for a in range(12):
				c = O.rotate_2024(s, a // 2, 1 - a % 2)
				if seen.lookup(c[None])[0]:
					continue
				if O.is_solved(c, True):
					queue = [a]
					while pi in parent_of:
						queue.insert(0, parent_of[pi][1]); pi = parent_of[pi][0]
					found = True
					break
				_, _, idx = seen.insert_unique(c[None])
				parent_of[int(idx[0])] = (int(pi), a)
				nxt.append(c)
			if found:
				break
		frontier = np.array(nxt)
	assert len(seen) == int(g["bfs_len"]) and queue == g["bfs_queue"].tolist()


def test_bfs_search_budget_limited(golden):
	"""agents.py:104: the budget is tested before every parent pop -- found flag, len(agent) and the action queue recorded
	from the reference under binding budgets (tests/golden/make_golden.py gen_bfs_budget), both representations."""
	g = golden("bfs_budget")
	n = int(g["n_cases"])
	for c in range(0, n, 1):
		start, max_states = g[f"start_{c}"], int(g[f"max_{c}"])
		if max_states > 6000 and c % 3:                       # the slow Python loop: every third of the large budgets
			continue
		found, length, queue = O.bfs_search(start, max_states, start.shape == (20,))
		assert found == bool(g[f"found_{c}"]) and length == int(g[f"len_{c}"]) and queue == g[f"queue_{c}"].tolist(), c


@pytest.mark.parametrize("tag,is2024", [("2024", True), ("686", False)])
def test_astar_expand_batch_trace(golden, tag, is2024):
	g = golden("search")
	w = g[f"astar_w_{tag}"]
	h_fn = lambda s: -np.floor((O.as_oh(s, is2024) @ w) / np.float32(4.0)).astype(np.float32)
	a = O.AStarFrontier(0.16, 7, h_fn, is2024)
	a.reset(g[f"astar_start_{tag}"])
	won = False
	for step, want in enumerate(g[f"astar_batches_{tag}"]):
		idcs = a.pop_batch()
		assert idcs.tolist() == want[want >= 0].tolist()
		won, _ = a.expand_batch(idcs)
		assert len(a) == g[f"astar_lens_{tag}"][step]
		if won:
			break
	assert won == bool(g[f"astar_won_{tag}"])
	L = len(a)
	assert (a.states[1:L + 1] == g[f"astar_states_{tag}"]).all()
	assert (a.G[1:L + 1] == g[f"astar_G_{tag}"]).all()
	assert (a.parents[2:L + 1] == g[f"astar_parents_{tag}"]).all()
	assert (a.parent_actions[2:L + 1] == g[f"astar_pact_{tag}"]).all()
	assert np.allclose(np.array(sorted(a.open)), g[f"astar_open_{tag}"], rtol=0, atol=0)


def test_astar_full_search(golden):
	g = golden("search")
	h_fn = lambda s: np.zeros(len(s), np.float32)
	a = O.AStarFrontier(1.0, 5, h_fn)
	start = g["astar_full_start"]
	a.reset(start)
	won = False
	while len(a) + 5 * 12 <= 20000 and not won:
		won, _ = a.expand_batch(a.pop_batch())
	assert won and bool(g["astar_full_ok"]) and len(a) == int(g["astar_full_len"])
	i = int(a.seen.lookup(O.solved_2024()[None])[0])
	queue = []
	while i != 1:
		queue.insert(0, int(a.parent_actions[i])); i = int(a.parents[i])
	assert queue == g["astar_full_queue"].tolist()
	s = start
	for act in queue:
		s = O.rotate_2024(s, act // 2, 1 - act % 2)
	assert O.is_solved(s, True)


def test_mcts_indices_and_neighbors(golden):
	"""MCTS.expand_leaf (agents.py:511-544): stored states get contiguous indices from 1 in
	discovery order and the neighbour table is consistent with the transitions."""
	g = golden("search")
	states, nb, leaves = g["mcts_states"], g["mcts_neighbors"], g["mcts_leaves"]
	seen = O.SeenSet()
	_, first, idx = seen.insert_unique(states)
	assert first.all() and idx.tolist() == list(range(1, len(states) + 1))
	for i in np.where(~leaves[1:])[0] + 1:
		ch_idx = seen.lookup(O.expand12(states[i - 1][None], True))
		assert (ch_idx == nb[i]).all()
		assert (nb[nb[i], O.rev_actions(np.arange(12))] == i).all()


def _count_net(oh):
	"""The golden MCTS runs' fake net: uniform policy logits, value = number of cubies in their solved place."""
	w = O.as_oh_2024(O.solved_2024()[None])[0]
	return np.zeros((len(oh), 12), np.float32), np.floor(oh @ w)


@pytest.mark.parametrize("tag,c,graph,max_states", [("a", 5.0, True, 3000), ("b", 0.6, False, 400)])
def test_mcts_full_trace(golden, tag, c, graph, max_states):
	"""MCTS.search end to end (agents.py:461-633): node numbering, neighbour table, leaf flags, visit counts, W, V and
	the action queue (after graph completion + shortening in case a) equal the reference's."""
	g = golden("search")
	m = O.MCTSOracle(_count_net, c=c, search_graph=graph)
	assert m.search(g[f"mcts{tag}_start"], max_states) == bool(g[f"mcts{tag}_ok"])
	L = len(m)
	assert L == int(g[f"mcts{tag}_len"]) and m.action_queue == g[f"mcts{tag}_queue"].tolist()
	assert (m.states[1:L + 1] == g[f"mcts{tag}_states"]).all()
	assert (m.neighbors[:L + 1] == g[f"mcts{tag}_neighbors"]).all() and (m.leaves[:L + 1] == g[f"mcts{tag}_leaves"]).all()
	assert (m.N[:L + 1] == g[f"mcts{tag}_N"]).all() and (m.W[1:L + 1] == g[f"mcts{tag}_W"]).all() and (m.V[1:L + 1] == g[f"mcts{tag}_V"]).all()


# ---- device-seeded stream: the generator restated in the oracle is Philox4x32-10 ---------------------------------------
def test_philox4x32_10_known_answers():
	"""Random123 kat_vectors (philox4x32 10): counter, key -> output."""
	kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
		   ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
		   ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
	for ctr, key, want in kat:
		assert O.philox4x32_10(np.array([ctr], dtype=np.uint32), key)[0].tolist() == list(want)
	a = O.seeded_actions(1234, 5, 64, 100)
	assert a.shape == (64, 100) and a.max() == 11 and a.min() == 0
	assert (O.seeded_actions(1234, 5, 64, 37) == a[:, :37]).all()               # prefix property
	assert (O.seeded_actions(1234, 25, 10, 100) == a[20:30]).all()              # cube id = subsequence
	for depth in (100, 37):
		assert (O.unpack_actions(O.pack_actions(a[:, :depth]), depth) == a[:, :depth]).all()
