"""GPU parity tests for the search frontier: hash-set dedup with the reference's batch-order
index semantics, BFS layer counts (known answers), the BFS and A* agents against traces recorded from the reference."""
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


def _states(n, is2024, seed, depth):
	g = np.random.RandomState(seed)
	return O.scramble_many(g.randint(0, 6, (n, depth)), g.randint(0, 2, (n, depth)), is2024)


@pytest.mark.parametrize("is2024", REPS)
def test_insert_unique_matches_dict_semantics(is2024):
	"""agents.py:286-306: seen / first-occurrence flags and batch-order indices, with heavy in-batch duplication, across
	several batches and through table growth."""
	from rl_rubiks_b200.frontier import StateHashSet
	hs = StateHashSet(16, is2024)
	ref = O.SeenSet()
	for b, (n, depth) in enumerate(((1, 0), (500, 2), (3000, 3), (257, 1), (5000, 4))):
		s = _states(n, is2024, seed=b, depth=depth) if depth else O.solved(is2024)[None]
		seen, first, idx = hs.insert_unique(s)
		w_seen, w_first, w_idx = ref.insert_unique(s)
		assert (seen == w_seen).all() and (first == w_first).all() and (idx == w_idx).all()
		assert len(hs) == len(ref)
	probe = np.concatenate([_states(100, is2024, seed=99, depth=6), _states(100, is2024, seed=1, depth=2)])
	assert (hs.lookup(probe) == ref.lookup(probe)).all()


@pytest.mark.parametrize("is2024,depth", [pytest.param(True, 6, id="2024"), pytest.param(False, 4, id="686")])
def test_bfs_layer_counts_known_answers(is2024, depth):
	"""SURVEY 8c KAT (i): 1, 12, 114, 1068, 10011, 93840, 878880 new states per depth (computed with the reference)."""
	from rl_rubiks_b200.frontier import bfs_layers
	counts, hs = bfs_layers(depth, is2024=is2024)
	assert counts == [1, 12, 114, 1068, 10011, 93840, 878880][:depth + 1]
	assert len(hs) == sum(counts)


def test_frontier_expand_vs_oracle_step_by_step():
	from rl_rubiks_b200.frontier import StateHashSet
	for is2024 in (True, False):
		hs, ref = StateHashSet(1 << 10, is2024), O.SeenSet()
		frontier = _states(40, is2024, seed=7, depth=3)
		_, first, _ = ref.insert_unique(frontier)
		hs.insert_unique(frontier)
		frontier = frontier[first]
		for _ in range(2):
			out = hs.expand(torch.from_numpy(frontier).cuda(), flags=True, index=True)
			children = O.expand12(frontier, is2024)
			seen, first, idx = ref.insert_unique(children)
			new = first & ~seen
			n_new = int(out["n_new"].item())
			assert n_new == new.sum()
			assert (out["next"][:n_new].cpu().numpy() == children[new]).all()
			assert (out["parent"][:n_new].cpu().numpy() == np.repeat(np.arange(len(frontier)), 12)[new]).all()
			assert (out["action"][:n_new].cpu().numpy() == np.tile(np.arange(12), len(frontier))[new]).all()
			assert (out["solved"][:n_new].cpu().numpy().astype(bool) == O.multi_is_solved(children[new], is2024)).all()
			assert (out["seen"].cpu().numpy().astype(bool) == seen).all() and (out["first"].cpu().numpy().astype(bool) == first).all()
			assert (out["index"].cpu().numpy() == idx).all()
			frontier = children[new]


def test_bfs_agent_matches_reference(golden):
	"""BFS.search (agents.py:96-123): found flag, states explored and action queue recorded from the reference."""
	from rl_rubiks_b200 import cube
	from rl_rubiks_b200.frontier import BFS
	g = golden("search")
	agent = BFS()
	assert agent.search(g["bfs_start"], None, 10 ** 6) == bool(g["bfs_found"])
	assert len(agent) == int(g["bfs_len"])
	assert list(agent.action_queue) == g["bfs_queue"].tolist()
	s = g["bfs_start"]
	for a in agent.action_queue:
		s = cube.rotate(s, *cube.action_space[a])
	assert cube.is_solved(s)
	assert BFS().search(cube.get_solved(), None, 10) is True


def test_bfs_agent_budget_limited_matches_reference(golden):
	"""agents.py:104 tests the budget before every parent pop: found flag, len(agent) and the action queue recorded from the
	reference under binding budgets (1 ... 20000 states on depth 3-6 scrambles, both representations)."""
	from rl_rubiks_b200 import cube
	from rl_rubiks_b200.frontier import BFS
	g = golden("bfs_budget")
	for slice_ in (4096, 7):                                   # 7: every layer is expanded in many slices
		for c in range(int(g["n_cases"])):
			start, max_states = g[f"start_{c}"], int(g[f"max_{c}"])
			if slice_ == 7 and max_states > 2000:
				continue
			cube.set_is2024(start.shape == (20,))
			agent = BFS()
			agent._min_slice = slice_
			found = agent.search(start, None, max_states)
			assert found == bool(g[f"found_{c}"]), (c, slice_)
			assert len(agent) == int(g[f"len_{c}"]), (c, slice_, len(agent), int(g[f"len_{c}"]))
			assert list(agent.action_queue) == g[f"queue_{c}"].tolist(), (c, slice_)
	cube.set_is2024(True)


def test_bfs_depth7_known_answer():
	"""BASELINE configs[4] / SURVEY 8c KAT (i): the depth-7 closure from solved has 8 221 632 new states in its last layer
	(OEIS A080583) and 9 205 558 in total; every one of the 11 807 112 generated children is accounted for."""
	from rl_rubiks_b200.frontier import bfs_layers
	counts, hs = bfs_layers(7, is2024=True, capacity=1 << 25)
	assert counts == [1, 12, 114, 1068, 10011, 93840, 878880, 8221632]
	assert len(hs) == 9205558
	# the closure is closed under lookups: the 12 neighbours of the depth-6 states are all in the set
	from rl_rubiks_b200 import cube
	acts = np.random.RandomState(0).randint(0, 12, (5000, 6)).astype(np.uint8)
	inner = cube.scramble_batch(acts)
	assert (hs.lookup(cube.expand12(inner)) > 0).all()


def test_full_table_is_reported_not_silent():
	"""A batch that finds the table full sets the device counters to -1 (no sync inside the call); the mirror raises on the
	next read.  Dropped items get index -1.  Through the C ABI directly, below the mirror's load-factor management."""
	from rl_rubiks_b200 import _native as N
	from rl_rubiks_b200.frontier import StateHashSet, read_count
	hs = StateHashSet(64, True)
	assert hs.capacity == 64
	states = torch.from_numpy(_states(500, True, seed=3, depth=8)).cuda()
	n = states.shape[0]
	seen = torch.empty(n, dtype=torch.uint8, device="cuda"); first = torch.empty_like(seen)
	index = torch.empty(n, dtype=torch.int32, device="cuda")
	scratch = torch.empty(N.lib.rb_hashset_scratch_bytes(n), dtype=torch.uint8, device="cuda")
	N.check(N.lib.rb_hashset_insert_unique(hs.rep, N.ptr(hs.table), hs.capacity, N.ptr(states), n, N.ptr(hs.count), N.ptr(seen), N.ptr(first),
										   N.ptr(index), N.ptr(scratch), N.stream_handle()))
	assert int(hs.count.item()) == -1 and int((index == -1).sum().item()) >= n - 64
	with pytest.raises(N.RubiksError):
		len(hs)
	# the error is sticky on that counter
	N.check(N.lib.rb_hashset_insert_unique(hs.rep, N.ptr(hs.table), hs.capacity, N.ptr(states[:4]), 4, N.ptr(hs.count), None, None, None,
										   N.ptr(scratch), N.stream_handle()))
	with pytest.raises(N.RubiksError):
		read_count(hs.count)
	# the mirror itself never gets there: it grows the table first
	hs2 = StateHashSet(16, True)
	_, first2, idx2 = hs2.insert_unique(states)
	assert len(hs2) == int(first2.sum().item()) and int(idx2.min().item()) >= 1


class _FakeNet(torch.nn.Module):
	def __init__(self, w, quant=4.0):
		super().__init__()
		self.w, self.quant = torch.from_numpy(np.asarray(w, dtype=np.float32)).cuda(), quant

	def forward(self, x, policy=True, value=True):
		return torch.floor((x @ self.w) / self.quant).unsqueeze(1)


@pytest.mark.parametrize("is2024", REPS)
def test_astar_expand_batch_trace_matches_reference(golden, is2024):
	"""AStar.expand_batch (agents.py:254-331) step by step: popped batches, set sizes, stored states, G, parents and
	parent actions equal the trace recorded from the reference with a tie-heavy integer net."""
	import heapq
	from rl_rubiks_b200 import cube
	from tests.agent_harness import AStar
	cube.set_is2024(is2024)
	tag = "2024" if is2024 else "686"
	g = golden("search")
	a = AStar(_FakeNet(g[f"astar_w_{tag}"]), lambda_=0.16, expansions=7)
	a.reset(None, 10 ** 6)
	root, _ = a.hs._states(g[f"astar_start_{tag}"])
	a.hs.insert_unique(root)
	a.states[1], a.G[1], a.n_states = root[0], 0, 1
	a.open_queue = [(0, 1)]
	won = False
	for step, want in enumerate(g[f"astar_batches_{tag}"]):
		n = min(len(a.open_queue), a.expansions)
		idcs = np.array([heapq.heappop(a.open_queue)[1] for _ in range(n)], dtype=int)
		assert idcs.tolist() == want[want >= 0].tolist()
		won = a.expand_batch(idcs)
		assert len(a) == g[f"astar_lens_{tag}"][step] == len(a.hs)
		if won:
			break
	assert won == bool(g[f"astar_won_{tag}"])
	L = len(a)
	assert (a.states[1:L + 1].cpu().numpy() == g[f"astar_states_{tag}"]).all()
	assert (a.G[1:L + 1] == g[f"astar_G_{tag}"]).all()
	assert (a.parents[2:L + 1] == g[f"astar_parents_{tag}"]).all()
	assert (a.parent_actions[2:L + 1] == g[f"astar_pact_{tag}"]).all()
	assert np.array_equal(np.array(sorted(a.open_queue)), g[f"astar_open_{tag}"])


def test_astar_full_search_matches_reference(golden):
	from rl_rubiks_b200 import cube
	from tests.agent_harness import AStar
	g = golden("search")
	a = AStar(_FakeNet(np.zeros(480)), lambda_=1.0, expansions=5)
	assert a.search(g["astar_full_start"], None, 20000) == bool(g["astar_full_ok"])
	assert len(a) == int(g["astar_full_len"]) and list(a.action_queue) == g["astar_full_queue"].tolist()
	s = g["astar_full_start"]
	for act in a.action_queue:
		s = cube.rotate(s, *cube.action_space[act])
	assert cube.is_solved(s)
	# reference tests/test_agents.py:122-134: after a search the 12 children of the root have G == 1 and parent == 1
	assert (a.G[2:14] == 1).all() and (a.parents[2:14] == 1).all()


def test_mcts_style_lookup(golden):
	"""MCTS bookkeeping (agents.py:511-544, 597-611): contiguous indices in discovery order, neighbour lookups."""
	from rl_rubiks_b200.frontier import StateHashSet
	g = golden("search")
	states, nb, leaves = g["mcts_states"], g["mcts_neighbors"], g["mcts_leaves"]
	hs = StateHashSet(1 << 12, True)
	_, first, idx = hs.insert_unique(states)
	assert first.all() and idx.tolist() == list(range(1, len(states) + 1))
	inner = np.where(~leaves[1:])[0] + 1
	ch_idx = hs.lookup(O.expand12(states[inner - 1], True)).reshape(-1, 12)
	assert (ch_idx == nb[inner]).all()


class _CountNet(torch.nn.Module):
	"""The golden MCTS runs' fake net: uniform policy logits, value = number of cubies in their solved place."""

	def __init__(self):
		super().__init__()
		self.w = torch.from_numpy(O.as_oh_2024(O.solved_2024()[None])[0]).cuda()

	def forward(self, x, policy=True, value=True):
		return torch.zeros(x.shape[0], 12, device=x.device), torch.floor(x @ self.w).unsqueeze(1)


@pytest.mark.parametrize("tag,c,graph,max_states", [("a", 5.0, True, 3000), ("b", 0.6, False, 400)])
def test_mcts_full_trace_matches_reference(golden, tag, c, graph, max_states):
	"""MCTS.search (agents.py:461-633) with the child expansion / dedup / graph completion on the device: node numbering,
	neighbour table, leaf flags, visit counts, W, V and the action queue equal the trace recorded from the reference."""
	from rl_rubiks_b200 import cube
	from tests.agent_harness import MCTS
	g = golden("search")
	m = MCTS(_CountNet(), c=c, search_graph=graph)
	assert m.search(g[f"mcts{tag}_start"], None, max_states) == bool(g[f"mcts{tag}_ok"])
	L = len(m)
	assert L == int(g[f"mcts{tag}_len"]) == len(m.hs) and list(m.action_queue) == g[f"mcts{tag}_queue"].tolist()
	assert (m.states[1:L + 1].cpu().numpy() == g[f"mcts{tag}_states"]).all()
	assert (m.neighbors[:L + 1] == g[f"mcts{tag}_neighbors"]).all() and (m.leaves[:L + 1] == g[f"mcts{tag}_leaves"]).all()
	assert (m.N[:L + 1] == g[f"mcts{tag}_N"]).all() and (m.W[1:L + 1] == g[f"mcts{tag}_W"]).all() and (m.V[1:L + 1] == g[f"mcts{tag}_V"]).all()
	if bool(g[f"mcts{tag}_ok"]):
		s = g[f"mcts{tag}_start"]
		for act in m.action_queue:
			s = cube.rotate(s, *cube.action_space[act])
		assert cube.is_solved(s)


def test_mcts_686_solves_shallow_scramble():
	"""Same agent on the 6x8x6 representation (no golden trace: the action queue must solve the cube)."""
	from rl_rubiks_b200 import cube
	from tests.agent_harness import MCTS
	cube.set_is2024(False)

	class Net(torch.nn.Module):
		def forward(self, x, policy=True, value=True):
			w = torch.from_numpy(O.as_oh_686(O.solved_686()[None])[0]).to(x.device)
			return torch.zeros(x.shape[0], 12, device=x.device), torch.floor(x @ w).unsqueeze(1)

	s = O.solved_686()
	for a in (3, 8):
		s = O.rotate_686(s, a // 2, 1 - a % 2)
	m = MCTS(Net(), c=0.6, search_graph=True)
	assert m.search(s, None, 2000)
	for act in m.action_queue:
		s = cube.rotate(s, *cube.action_space[act])
	assert cube.is_solved(s) and len(m) == len(m.hs)


def _oracle_astar(start, lam, n_exp, w, max_states, quant=4.0):
	"""AStar.search (agents.py:221-252) on the oracle's bookkeeping with the integer fake net."""
	h = lambda states: -np.floor((O.as_oh_2024(states) @ w) / np.float32(quant)).astype(np.float32)
	a = O.AStarFrontier(lam, n_exp, h)
	a.reset(start, capacity=max_states + 12 * n_exp + 2)
	if O.is_solved(start, True):
		return True, a
	while len(a) + 12 * n_exp <= max_states:
		won, _ = a.expand_batch(a.pop_batch())
		if won:
			return True, a
	return False, a


@pytest.mark.parametrize("n_exp,max_states,lam", [(7, 700, 0.16), (64, 5000, 0.5), (700, 30000, 0.16)])
def test_astar_batch_matches_single_search_traces(golden, n_exp, max_states, lam):
	"""AStarBatch (device open list, pops, dedup, relaxation for K cubes in lockstep) against the single-search bookkeeping
	of the reference, search by search: solved flag, len, stored states, G, parents, parent actions, open list, action queue."""
	from rl_rubiks_b200 import cube
	from rl_rubiks_b200.frontier import AStarBatch
	g = golden("search")
	w = g["astar_w_2024"]
	rng = np.random.RandomState(n_exp)
	starts = [g["astar_start_2024"], O.solved_2024()]
	for depth in (1, 2, 3, 4, 5, 6, 8, 11, 30):
		starts.append(O.scramble(rng.randint(0, 6, depth), rng.randint(0, 2, depth), True))
	starts = np.stack(starts)
	agent = AStarBatch(_FakeNet(w), lambda_=lam, expansions=n_exp)
	won, queues, count = agent.search_many(starts, max_states)
	for s, start in enumerate(starts):
		ok, a = _oracle_astar(start, lam, n_exp, w, max_states)
		L = len(a)
		assert bool(won[s]) == ok and count[s] == L, (s, won[s], ok, count[s], L)
		assert (agent.states[s, 1:L + 1].cpu().numpy() == a.states[1:L + 1]).all()
		assert (agent.G[s, 1:L + 1].cpu().numpy() == a.G[1:L + 1]).all()
		assert (agent.parents[s, 2:L + 1].cpu().numpy() == a.parents[2:L + 1]).all()
		assert (agent.parent_actions[s, 2:L + 1].cpu().numpy() == a.parent_actions[2:L + 1]).all()
		if not ok or L == 1:
			in_open = agent.in_open[s, :L + 1].bool().cpu().numpy()
			cost = agent.cost[s, :L + 1].cpu().numpy()
			mine = sorted((float(cost[i]), int(i)) for i in np.where(in_open)[0])
			assert mine == sorted((float(c), int(i)) for c, i in a.open) or L == 1
		if ok:
			state = start
			for act in queues[s]:
				state = cube.rotate(state, *cube.action_space[act])
			assert cube.is_solved(state)
			i, q = int(a.seen.lookup(O.solved_2024()[None])[0]), []
			while i != 1:
				q.insert(0, int(a.parent_actions[i])); i = int(a.parents[i])
			assert q == queues[s]
	# a second call on the same agent reuses every buffer and must reproduce the run
	won2, queues2, count2 = agent.search_many(starts, max_states)
	assert (won2 == won).all() and (count2 == count).all() and queues2 == queues


def test_representation_conversion_round_trip():
	"""rb_as686 / rb_as2024: the same move sequence applied in either representation gives states that convert into each other."""
	from rl_rubiks_b200 import cube
	g = np.random.RandomState(0)
	f, d = g.randint(0, 6, (3000, 25)), g.randint(0, 2, (3000, 25))
	s2024, s686 = O.scramble_many(f, d, True), O.scramble_many(f, d, False)
	s2024[0], s686[0] = O.solved(True), O.solved(False)
	assert (cube.to_686(s2024) == s686).all()
	assert (cube.to_2024(s686) == s2024).all()
	bad = s686[:3].copy()
	bad[1, 0, 0] = bad[1, 1, 0]                               # two stickers of one corner show colours no corner has
	with pytest.raises(IndexError):
		cube.to_2024(bad)


@pytest.mark.parametrize("use_graphs", [True, False])
def test_astar_batch_686_matches_harness_traces(golden, use_graphs):
	"""AStarBatch under the 6x8x6 representation (searches run on the 20-byte states, 288-wide one-hot rows for the net) against
	the single-search harness on the 6x8x6 states themselves and against the reference's recorded 6x8x6 trace."""
	from rl_rubiks_b200 import cube
	from rl_rubiks_b200.frontier import AStarBatch
	from tests.agent_harness import AStar
	g = golden("search")
	w = g["astar_w_686"]
	rng = np.random.RandomState(4)
	starts = [g["astar_start_686"], O.solved_686()]
	for depth in (1, 2, 3, 4, 5, 7):
		starts.append(O.scramble(rng.randint(0, 6, depth), rng.randint(0, 2, depth), False))
	starts = np.stack(starts)
	cube.set_is2024(False)
	batch = AStarBatch(_FakeNet(w), lambda_=0.16, expansions=7, use_graphs=use_graphs)
	assert not batch.is2024
	won, queues, count = batch.search_many(starts, 2000)
	for s, start in enumerate(starts):
		single = AStar(_FakeNet(w), lambda_=0.16, expansions=7)
		ok = single.search(start, None, 2000)
		assert bool(won[s]) == ok and (count[s] == len(single) or (ok and len(single) == 0)), (s, won[s], ok, count[s], len(single))
		assert queues[s] == list(single.action_queue)
		L = len(single)
		if L:
			assert (cube.to_686(batch.states[s, 1:L + 1]).cpu().numpy() == single.states[1:L + 1].cpu().numpy()).all()
			assert (batch.G[s, 1:L + 1].cpu().numpy() == single.G[1:L + 1]).all()
			assert (batch.parents[s, 2:L + 1].cpu().numpy() == single.parents[2:L + 1]).all()
	# the reference's own recorded trace of search 0 (12 expansion steps)
	L = int(g["astar_lens_686"][-1])
	if not bool(g["astar_won_686"]) and count[0] >= L:
		assert (cube.to_686(batch.states[0, 1:L + 1]).cpu().numpy() == g["astar_states_686"]).all()
	cube.set_is2024(True)
	# the one-cube agent interface against the reference's recorded full search (agents.py:221-252)
	one = AStarBatch(_FakeNet(np.zeros(480)), lambda_=1.0, expansions=5, use_graphs=use_graphs)
	assert one.search(g["astar_full_start"], None, 20000) == bool(g["astar_full_ok"])
	assert len(one) == int(g["astar_full_len"]) and list(one.action_queue) == g["astar_full_queue"].tolist()
	assert one.search(O.solved_2024(), None, 100) and len(one) == 0 and not one.action_queue


def test_chained_layers_equal_layer_by_layer_expansion():
	"""`rb_frontier_expand_chain` (layer sizes left on the device, grids sized for the upper bound 12^d) against one
	`rb_frontier_expand` per layer with a host read in between: same counts, same numbering of every state."""
	from rl_rubiks_b200 import cube
	from rl_rubiks_b200.frontier import bfs_layers
	start = O.scramble(np.array([0, 3, 5]), np.array([1, 0, 1]), True)
	for depth in (1, 2, 5):
		c_chain, hs_chain = bfs_layers(depth, start=start, is2024=True)
		c_plain, hs_plain = bfs_layers(depth, start=start, is2024=True, chain_items=0)
		assert c_chain == c_plain and len(hs_chain) == len(hs_plain) == sum(c_plain)
		probe = cube.scramble_batch(np.random.RandomState(depth).randint(0, 12, (3000, depth)).astype(np.uint8), start=np.repeat(start[None], 3000, 0))
		idx = hs_plain.lookup(probe)
		assert (idx > 0).all() and (hs_chain.lookup(probe) == idx).all()


def test_hash_set_batches_are_deterministic_under_contention():
	"""Stand-in for racecheck on the hash kernels: a batch with heavy in-batch duplication (every state ~16 times, claimed and raced
	for by many threads at once) must give the same seen / first / index arrays on every run and equal the dict semantics."""
	from rl_rubiks_b200.frontier import StateHashSet
	base = _states(4000, True, seed=1, depth=12)
	rng = np.random.RandomState(2)
	batch = base[rng.randint(0, len(base), 64000)]
	want = O.SeenSet().insert_unique(batch)
	for _ in range(6):
		hs = StateHashSet(1 << 18, True)
		seen, first, idx = hs.insert_unique(batch)
		assert (seen == want[0]).all() and (first == want[1]).all() and (idx == want[2]).all()
		seen2, first2, idx2 = hs.insert_unique(batch[::-1].copy())
		assert seen2.all() and (idx2 == idx[::-1]).all() and len(hs) == int(want[1].sum())
