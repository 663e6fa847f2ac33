"""
Search frontier on the device: the seen-set, child expansion and dedup that `BFS`, `AStar` and `MCTS` do with a
Python dict keyed on `state.tostring()` (reference: librubiks/solving/agents.py:96-123, 254-331, 511-544, 597-611).

`StateHashSet` is an open-addressing hash set on the packed state in caller-owned (torch) device memory with the
reference's index semantics: states are numbered 1, 2, ... in insertion order, a batch is numbered in batch order,
in-batch duplicates resolve to their first occurrence (agents.py:286-306).  `BFS` and `AStar` mirror the reference
agents' interface (`search(state, time_limit, max_states) -> bool`, `action_queue`, `len(agent)`).
"""
from __future__ import annotations

import heapq
from collections import deque
from time import perf_counter

import numpy as np
import torch

from . import _native as N
from . import cube


def _pow2_at_least(x: int) -> int:
	return 1 << max(4, int(x - 1).bit_length())


def read_count(t: torch.Tensor) -> int:
	"""Host read of a device counter of the frontier kernels (`count`, `n_new`): -1 means a batch found the hash table full
	(include/rubiks_b200.h, RB_ERR_CAPACITY semantics without a sync inside the call)."""
	n = int(t.item())
	if n < 0:
		raise N.RubiksError(f"librubiks_b200 error {N.RB_ERR_CAPACITY}: the state hash set overflowed (table full); its contents are undefined")
	return n


class StateHashSet:
	def __init__(self, capacity: int = 1 << 16, is2024: bool | None = None):
		N.require_cuda()
		self.is2024 = cube.get_is2024() if is2024 is None else is2024
		self.rep = N.REP_2024 if self.is2024 else N.REP_686
		self.shape = (20,) if self.is2024 else (6, 8, 6)
		self.dev = torch.device("cuda", torch.cuda.current_device())
		self.capacity = _pow2_at_least(capacity)
		self.table = torch.empty(N.lib.rb_hashset_bytes(self.capacity), dtype=torch.uint8, device=self.dev)
		self.count = torch.zeros(1, dtype=torch.int32, device=self.dev)
		self.n_new = torch.zeros(1, dtype=torch.int32, device=self.dev)
		self._upper = 0                      # host-side upper bound of the set size (no sync needed)
		self._scratch = None
		self.clear()

	def clear(self):
		N.check(N.lib.rb_hashset_clear(N.ptr(self.table), self.capacity, N.stream_handle()))
		self.count.zero_()
		self._upper = 0

	def __len__(self) -> int:
		n = read_count(self.count)
		self._upper = n
		return n

	def _reserve(self, incoming: int):
		"""Keeps the load factor <= 1/2 using the host-side upper bound (exact size read back only when needed)."""
		if 2 * (self._upper + incoming) <= self.capacity:
			return
		len(self)
		if 2 * (self._upper + incoming) <= self.capacity:
			return
		new_cap = _pow2_at_least(4 * (self._upper + incoming))
		new_table = torch.empty(N.lib.rb_hashset_bytes(new_cap), dtype=torch.uint8, device=self.dev)
		N.check(N.lib.rb_hashset_rehash(N.ptr(self.table), self.capacity, N.ptr(new_table), new_cap, N.stream_handle()))
		self.table, self.capacity = new_table, new_cap

	def _scratch_for(self, nbytes: int) -> torch.Tensor:
		if self._scratch is None or self._scratch.numel() < nbytes:
			self._scratch = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=self.dev)
		return self._scratch

	def _states(self, states):
		if isinstance(states, torch.Tensor):
			t, was_np = states.to(device=self.dev, dtype=torch.int8), False
		else:
			t, was_np = torch.from_numpy(np.ascontiguousarray(states, dtype=np.int8)).to(self.dev), True
		if t.dim() == len(self.shape):
			t = t.unsqueeze(0)
		if tuple(t.shape[1:]) != self.shape:
			raise IndexError(f"states of shape {tuple(t.shape)} do not match {self.shape}")
		return t.contiguous(), was_np

	def insert_unique(self, states):
		"""-> (seen bool (n,), first bool (n,), index int32 (n,)) with the semantics of agents.py:286-306."""
		s, was_np = self._states(states)
		n = s.shape[0]
		self._reserve(n)
		seen = torch.empty(n, dtype=torch.uint8, device=self.dev)
		first = torch.empty(n, dtype=torch.uint8, device=self.dev)
		index = torch.empty(n, dtype=torch.int32, device=self.dev)
		scratch = self._scratch_for(N.lib.rb_hashset_scratch_bytes(n))
		N.check(N.lib.rb_hashset_insert_unique(self.rep, N.ptr(self.table), self.capacity, N.ptr(s), n, N.ptr(self.count), N.ptr(seen),
											   N.ptr(first), N.ptr(index), N.ptr(scratch), N.stream_handle()))
		self._upper += n
		out = (seen.bool(), first.bool(), index)
		return tuple(x.cpu().numpy() for x in out) if was_np else out

	def lookup(self, states):
		"""Index of every state, 0 when absent (agents.py:606-607)."""
		s, was_np = self._states(states)
		n = s.shape[0]
		index = torch.empty(n, dtype=torch.int32, device=self.dev)
		scratch = self._scratch_for(N.lib.rb_hashset_scratch_bytes(n))
		N.check(N.lib.rb_hashset_lookup(self.rep, N.ptr(self.table), self.capacity, N.ptr(s), n, N.ptr(index), N.ptr(scratch), N.stream_handle()))
		return index.cpu().numpy() if was_np else index

	def expand(self, frontier: torch.Tensor, flags: bool = False, parents: bool = True, solved: bool = True, index: bool = False):
		"""One frontier step: 12 children of every frontier state, dedup against the set and within the batch, new states
		compacted in batch order.  Returns a dict of device tensors sized for the worst case (12 n rows); the first
		`n_new` rows are valid.  `n_new` itself stays on the device (`out['n_new']`) until the caller reads it."""
		f, _ = self._states(frontier)
		n = f.shape[0]
		m = 12 * n
		self._reserve(m)
		out = {"next": torch.empty(m, *self.shape, dtype=torch.int8, device=self.dev)}
		out["parent"] = torch.empty(m, dtype=torch.int32, device=self.dev) if parents else None
		out["action"] = torch.empty(m, dtype=torch.uint8, device=self.dev) if parents else None
		out["solved"] = torch.empty(m, dtype=torch.uint8, device=self.dev) if solved else None
		out["seen"] = torch.empty(m, dtype=torch.uint8, device=self.dev) if flags else None
		out["first"] = torch.empty(m, dtype=torch.uint8, device=self.dev) if flags else None
		out["index"] = torch.empty(m, dtype=torch.int32, device=self.dev) if index else None
		scratch = self._scratch_for(N.lib.rb_frontier_scratch_bytes(self.rep, n))
		N.check(N.lib.rb_frontier_expand(self.rep, N.ptr(self.table), self.capacity, N.ptr(f), n, N.ptr(self.count), N.ptr(out["next"]),
										 N.ptr(out["parent"]), N.ptr(out["action"]), N.ptr(out["solved"]), N.ptr(out["seen"]),
										 N.ptr(out["first"]), N.ptr(out["index"]), N.ptr(self.n_new), N.ptr(scratch), N.stream_handle()))
		self._upper += m
		out["n_new"] = self.n_new
		return out


LAST_LAYER_MS = []


def bfs_layers(max_depth: int, start=None, is2024: bool | None = None, capacity: int | None = None):
	"""Layer-synchronous BFS closure from `start` (default solved): per-depth counts of newly discovered states
	(1, 12, 114, 1068, ... for the 20x24 cube) and the StateHashSet.  Everything but one int per layer stays on the device."""
	is2024 = cube.get_is2024() if is2024 is None else is2024
	hs = StateHashSet(capacity or (1 << 16), is2024)
	if start is None:
		start = cube._solved[hs.rep]
	frontier, _ = hs._states(start)
	hs.insert_unique(frontier)
	counts = [1]
	events = [torch.cuda.Event(enable_timing=True) for _ in range(max_depth + 1)]
	events[0].record()
	for d in range(max_depth):
		out = hs.expand(frontier, parents=False, solved=False)
		events[d + 1].record()
		n_new = read_count(out["n_new"])
		frontier = out["next"][:n_new]
		counts.append(n_new)
	global LAST_LAYER_MS
	LAST_LAYER_MS = [events[d].elapsed_time(events[d + 1]) for d in range(max_depth)]      # device time per layer (diagnostics)
	return counts, hs


class Agent:
	"""Interface of the reference agents (agents.py:14-64)."""

	def __init__(self):
		self.action_queue = deque()
		self._explored_states = 0

	def reset(self, time_limit, max_states):
		self._explored_states = 0
		self.action_queue = deque()
		if hasattr(self, "net"):
			self.net.eval()
		assert time_limit or max_states
		return time_limit or 1e10, max_states or int(1e10)

	def __len__(self):
		return self._explored_states


class BFS(Agent):
	"""Breadth-first search (agents.py:92-129) as layer-synchronous kernel sequences instead of one dict probe per child.

	The reference pops ONE parent at a time and tests its budget `len(self) < max_states` before every pop (agents.py:104),
	so a search that runs out of budget stops in the middle of a layer.  Parents are expanded here a slice of the FIFO
	frontier at a time; new children come back compacted in FIFO order with their parent position, so the parent at which
	the reference would have stopped is found by position: the last admitted parent is the one that records the
	(budget)-th new state, every later parent's children are dropped, and a solved child only counts under an admitted
	parent.  `len(agent)`, the found flag and the action queue are the reference's for any budget
	(tests/golden/bfs_budget.npz).  The time limit is tested once per slice, not once per parent."""

	_min_slice = 4096                              # parents per expansion when the budget is small

	def __init__(self, is2024: bool | None = None):
		super().__init__()
		self.is2024 = is2024

	def search(self, state, time_limit: float = None, max_states: int = None) -> bool:
		time_limit, max_states = self.reset(time_limit, max_states)
		t0 = perf_counter()
		is2024 = cube.get_is2024() if self.is2024 is None else self.is2024
		hs = StateHashSet(1 << 16, is2024)
		frontier, _ = hs._states(state)
		if bool((frontier[0].cpu().numpy() == cube._solved[hs.rep]).all()):
			return True
		hs.insert_unique(frontier)
		total = self._explored_states = 1
		layers = []                                   # per layer: (position of the parent in the previous layer, action)
		while frontier.shape[0] and total < max_states:
			nxt, par, act, done = [], [], [], 0
			while done < frontier.shape[0] and total < max_states and perf_counter() - t0 < time_limit:
				# never more parents than the remaining budget could admit if each recorded a single new state ...
				take = min(frontier.shape[0] - done, max(self._min_slice, max_states - total))
				out = hs.expand(frontier[done:done + take])
				n_new = read_count(out["n_new"])
				parent = out["parent"][:n_new]
				n_adm = n_new
				if total + n_new >= max_states:
					# ... and the parent that records new state number (max_states - total) is the last one popped
					last = parent[max_states - total - 1]
					n_adm = int(torch.searchsorted(parent, last, right=True).item())
				hit = torch.nonzero(out["solved"][:n_adm])
				par.append(parent[:n_adm] + done); act.append(out["action"][:n_adm])
				if hit.numel():
					k = int(hit[0].item())                # first solved new child in FIFO order
					# the reference returns before recording it: the dict holds the states discovered before it
					self._explored_states = total + k
					layers.append((torch.cat(par), torch.cat(act)))
					pos = sum(p.shape[0] for p in par[:-1]) + k
					for parent, action in reversed(layers):
						self.action_queue.appendleft(int(action[pos].item()))
						pos = int(parent[pos].item())
					return True
				total += n_adm
				self._explored_states = total
				nxt.append(out["next"][:n_adm])
				done += take
			if done < frontier.shape[0]:
				return False                              # budget or time ran out inside the layer
			layers.append((torch.cat(par), torch.cat(act)))
			frontier = torch.cat(nxt)
		return False

	def __str__(self):
		return "Breadth-first search"


class AStar(Agent):
	"""Batched weighted A* (agents.py:171-413).  Device: child expansion, seen-set dedup with batch-order index
	assignment, compaction of the new states, one-hot, value net, solved test.  Host (as in the reference; the device
	open list is SURVEY 8f row N2): the heapq open list of (cost, index) and the G / parent relaxation."""

	_stack_expand = 1000

	def __init__(self, net, lambda_: float, expansions: int, is2024: bool | None = None):
		super().__init__()
		self.net, self.lambda_, self.expansions, self.is2024 = net, lambda_, expansions, is2024

	def reset(self, time_limit, max_states):
		time_limit, max_states = super().reset(time_limit, max_states)
		is2024 = cube.get_is2024() if self.is2024 is None else self.is2024
		self.hs = StateHashSet(1 << 16, is2024)
		self.open_queue = []
		self.states = torch.empty(self._stack_expand, *self.hs.shape, dtype=torch.int8, device=self.hs.dev)
		self.parents = np.empty(self._stack_expand, dtype=int)
		self.parent_actions = np.zeros(self._stack_expand, dtype=int)
		self.G = np.empty(self._stack_expand)
		self.n_states = 0
		return time_limit, max_states

	def increase_stack_size(self):
		self.states = torch.cat([self.states, torch.empty_like(self.states)])
		self.parents = np.concatenate([self.parents, np.zeros_like(self.parents)])
		self.parent_actions = np.concatenate([self.parent_actions, np.zeros_like(self.parent_actions)])
		self.G = np.concatenate([self.G, np.empty_like(self.G)])

	def __len__(self):
		return self.n_states

	@torch.no_grad()
	def search(self, state, time_limit: float = None, max_states: int = None) -> bool:
		t0 = perf_counter()
		time_limit, max_states = self.reset(time_limit, max_states)
		root, _ = self.hs._states(state)
		if bool((root[0].cpu().numpy() == cube._solved[self.hs.rep]).all()):
			return True
		self.hs.insert_unique(root)
		self.states[1], self.G[1], self.n_states = root[0], 0, 1
		heapq.heappush(self.open_queue, (0, 1))
		while perf_counter() - t0 < time_limit and len(self) + self.expansions * 12 <= max_states:
			n_remove = min(len(self.open_queue), self.expansions)
			expand_idcs = np.array([heapq.heappop(self.open_queue)[1] for _ in range(n_remove)], dtype=int)
			if self.expand_batch(expand_idcs):
				i = self.solved_index
				while i != 1:
					self.action_queue.appendleft(int(self.parent_actions[i]))
					i = int(self.parents[i])
				return True
		return False

	def expand_batch(self, expand_idcs: np.ndarray) -> bool:
		"""agents.py:254-331."""
		expand_size = len(expand_idcs)
		while len(self) + expand_size * 12 >= len(self.states):
			self.increase_stack_size()
		idcs_dev = torch.from_numpy(expand_idcs).to(self.hs.dev)
		out = self.hs.expand(self.states[idcs_dev], flags=True, index=True)
		n_new = read_count(out["n_new"])
		base = self.n_states
		new_states = out["next"][:n_new]
		self.states[base + 1:base + 1 + n_new] = new_states
		self.n_states = base + n_new
		new_idcs = base + np.arange(n_new) + 1
		parent_pos = out["parent"][:n_new].cpu().numpy()
		new_parent_idcs = expand_idcs[parent_pos]
		self.G[new_idcs] = self.G[new_parent_idcs] + 1
		self.parent_actions[new_idcs] = out["action"][:n_new].cpu().numpy()
		self.parents[new_idcs] = new_parent_idcs
		if n_new:
			costs = self.cost(new_states, new_idcs)
			for c, i in zip(costs, new_idcs):
				heapq.heappush(self.open_queue, (c, int(i)))
			solved = out["solved"][:n_new]
			if bool(solved.any().item()):
				self.solved_index = int(new_idcs[int(torch.nonzero(solved)[0].item())])
				return True
		seen = out["seen"].bool().cpu().numpy()
		first = out["first"].bool().cpu().numpy()
		index = out["index"].cpu().numpy().astype(int)
		old = first & seen
		parent_idcs = np.repeat(expand_idcs, 12)
		actions_taken = np.tile(np.arange(12), expand_size)
		self.relax_seen_states(index[old], parent_idcs[old], actions_taken[old])
		return False

	def relax_seen_states(self, state_idcs, parent_idcs, actions_taken):
		"""agents.py:333-367."""
		new_ways = self.G[parent_idcs] + 1 < self.G[state_idcs]
		nw_states, nw_parents = state_idcs[new_ways], parent_idcs[new_ways]
		self.G[nw_states] = self.G[nw_parents] + 1
		self.parent_actions[nw_states] = actions_taken[new_ways]
		self.parents[nw_states] = nw_parents
		shortcuts = self.G[state_idcs] + 1 < self.G[parent_idcs]
		sc_states, sc_parents = state_idcs[shortcuts], parent_idcs[shortcuts]
		self.G[sc_parents] = self.G[sc_states] + 1
		self.parent_actions[sc_parents] = cube.rev_actions(actions_taken[shortcuts])
		self.parents[sc_parents] = sc_states

	@torch.no_grad()
	def cost(self, states: torch.Tensor, indeces: np.ndarray) -> np.ndarray:
		"""agents.py:369-383: lambda * G + H with H = -value net; one-hot born on the device."""
		oh = torch.empty(states.shape[0], 480 if self.hs.is2024 else 288, dtype=torch.float32, device=self.hs.dev)
		N.check(N.lib.rb_as_oh(self.hs.rep, N.ptr(states.contiguous()), N.ptr(oh), states.shape[0], N.stream_handle()))
		H = -self.net(oh, value=True, policy=False)
		H = H.cpu().squeeze(-1).detach().numpy() if H.dim() > 1 else H.cpu().detach().numpy()
		return self.lambda_ * self.G[indeces] + H

	def __str__(self):
		return f'AStar (lambda={self.lambda_}, N={self.expansions})'


class MCTS(Agent):
	"""Monte Carlo tree search (agents.py:415-645).  Device: the child part of `expand_leaf` (12 children, seen-set
	dedup with batch-order numbering, compaction of the new states, solved test, one-hot; agents.py:511-544) and of
	`_complete_graph` (children of every leaf + index lookup; agents.py:597-611).  Host, as in the reference: the node
	arrays P, V, N, W, L, the neighbour table and the sequential UCB walk `find_leaf` (out of scope)."""

	def __init__(self, net, c: float, search_graph: bool, is2024: bool | None = None):
		super().__init__()
		self.net, self.c, self.search_graph, self.is2024 = net, c, search_graph, is2024
		self.nu = 100
		self.expand_nodes = 1000

	def reset(self, time_limit, max_states):
		time_limit, max_states = super().reset(time_limit, max_states)
		is2024 = cube.get_is2024() if self.is2024 is None else self.is2024
		self.hs = StateHashSet(1 << 14, is2024)
		n = self.expand_nodes
		self.states = torch.empty(n, *self.hs.shape, dtype=torch.int8, device=self.hs.dev)
		self.neighbors = np.zeros((n, 12), dtype=int)
		self.leaves = np.ones(n, dtype=bool)
		self.P, self.V = np.empty((n, 12)), np.empty(n)
		self.N, self.W, self.L = np.zeros((n, 12), dtype=int), np.zeros((n, 12)), np.zeros((n, 12))
		self.n_states = 0
		return time_limit, max_states

	def increase_stack_size(self):
		k = len(self.states)
		self.states = torch.cat([self.states, torch.empty_like(self.states)])
		self.neighbors = np.concatenate([self.neighbors, np.zeros((k, 12), dtype=int)])
		self.leaves = np.concatenate([self.leaves, np.ones(k, dtype=bool)])
		self.P, self.V = np.concatenate([self.P, np.empty((k, 12))]), np.concatenate([self.V, np.empty(k)])
		self.N = np.concatenate([self.N, np.zeros((k, 12), dtype=int)])
		self.W, self.L = np.concatenate([self.W, np.zeros((k, 12))]), np.concatenate([self.L, np.zeros((k, 12))])

	def __len__(self):
		return self.n_states

	def _oh(self, states: torch.Tensor) -> torch.Tensor:
		oh = torch.empty(states.shape[0], 480 if self.hs.is2024 else 288, dtype=torch.float32, device=self.hs.dev)
		if states.shape[0]:
			N.check(N.lib.rb_as_oh(self.hs.rep, N.ptr(states.contiguous()), N.ptr(oh), states.shape[0], N.stream_handle()))
		return oh

	@torch.no_grad()
	def search(self, state, time_limit: float = None, max_states: int = None) -> bool:
		t0 = perf_counter()
		time_limit, max_states = self.reset(time_limit, max_states)
		root, _ = self.hs._states(state)
		self.hs.insert_unique(root)
		self.states[1], self.n_states = root[0], 1
		if bool((root[0].cpu().numpy() == cube._solved[self.hs.rep]).all()):
			return True
		p, v = self.net(self._oh(root))
		self.P[1] = p.softmax(dim=1).cpu().numpy()
		self.V[1] = v.cpu().numpy().reshape(-1)[0]
		indices_visited, actions_taken = [1], []
		while perf_counter() - t0 < time_limit and len(self) + 12 <= max_states:
			solve_leaf_index, solve_action = self.expand_leaf(indices_visited, actions_taken)
			if solve_leaf_index != -1:
				self.action_queue = deque(actions_taken) + deque([solve_action])
				if self.search_graph:
					self._complete_graph()
					self._shorten_action_queue(solve_leaf_index)
				return True
			indices_visited, actions_taken = self.find_leaf(time_limit - (perf_counter() - t0))
		self.action_queue = deque(actions_taken)
		return False

	def expand_leaf(self, visited_states_idcs: list, actions_taken: list):
		"""agents.py:496-573; one `rb_frontier_expand` call replaces the 12 `tostring()` dict probes."""
		if len(self) + 12 > len(self.states):
			self.increase_stack_size()
		leaf_index = visited_states_idcs[-1]
		out = self.hs.expand(self.states[leaf_index:leaf_index + 1], flags=True, index=True, parents=False)
		substate_idcs = out["index"].cpu().numpy().astype(int)
		n_new = read_count(out["n_new"])
		new_substate_idcs = self.n_states + np.arange(n_new) + 1
		new_substates = out["next"][:n_new]
		self.states[self.n_states + 1:self.n_states + 1 + n_new] = new_substates
		self.n_states += n_new
		actions = np.arange(12)
		self.neighbors[leaf_index, actions] = substate_idcs
		self.neighbors[substate_idcs, cube.rev_actions(actions)] = leaf_index
		self.leaves[leaf_index] = False
		solve_leaf, solve_action = -1, -1
		# a solved child is always a new one (the search stops when it is first generated): test the compacted new states
		solved_new = torch.nonzero(out["solved"][:n_new]).reshape(-1)
		if solved_new.numel():
			solve_leaf = int(new_substate_idcs[int(solved_new[0].item())])
			solve_action = int(np.where(substate_idcs == solve_leaf)[0][0])
		p, v = self.net(self._oh(new_substates))
		p, v = p.softmax(dim=1).cpu().numpy(), v.cpu().numpy().reshape(-1)
		self.P[new_substate_idcs] = p
		self.V[new_substate_idcs] = v
		self.W[leaf_index] = self.V[self.neighbors[leaf_index]]
		self.W[new_substate_idcs] = np.tile(v, (12, 1)).T
		if n_new:        # the reference calls v.max() on an empty batch here and raises; all-seen children change nothing
			self.W[visited_states_idcs[:-1], actions_taken] = np.maximum(self.W[visited_states_idcs[:-1], actions_taken], v.max())
		if actions_taken:
			self.N[visited_states_idcs[:-1], actions_taken] += 1
			self.L[visited_states_idcs[:-1], actions_taken] = 0
			self.L[visited_states_idcs[1:], cube.rev_actions(np.array(actions_taken))] = 0
		return solve_leaf, solve_action

	def find_leaf(self, time_limit: float):
		"""agents.py:575-595 (sequential UCB walk on the host arrays)."""
		t0 = perf_counter()
		current_index, indices_visited, actions_taken = 1, [1], []
		while not self.leaves[current_index] and perf_counter() - t0 < time_limit:
			sqrtN = np.sqrt(self.N[current_index].sum())
			U = self.c * self.P[current_index] * sqrtN / (1 + self.N[current_index])
			Q = self.W[current_index] - self.L[current_index]
			action = int((U + Q).argmax())
			self.L[current_index, action] += self.nu
			current_index = int(self.neighbors[current_index, action])
			self.L[current_index, cube.rev_action(action)] += self.nu
			indices_visited.append(current_index)
			actions_taken.append(action)
		return indices_visited, actions_taken

	def _complete_graph(self):
		"""agents.py:597-611: children of every leaf in one expand12 + one hash lookup."""
		leaves_idcs = np.where(self.leaves[:len(self) + 1])[0][1:]
		if not len(leaves_idcs):
			return
		actions_taken = np.tile(np.arange(12), len(leaves_idcs))
		repeated_leaves_idcs = np.repeat(leaves_idcs, 12)
		leaf_states = self.states[torch.from_numpy(leaves_idcs).to(self.hs.dev)].contiguous()
		substates = torch.empty(12 * len(leaves_idcs), *self.hs.shape, dtype=torch.int8, device=self.hs.dev)
		N.check(N.lib.rb_expand12(self.hs.rep, N.ptr(leaf_states), N.ptr(substates), None, None, len(leaves_idcs), N.stream_handle()))
		substate_idcs = self.hs.lookup(substates).cpu().numpy().astype(int)
		self.neighbors[repeated_leaves_idcs, actions_taken] = substate_idcs
		self.neighbors[substate_idcs, cube.rev_actions(actions_taken)] = repeated_leaves_idcs
		self.neighbors[0] = 0

	def _shorten_action_queue(self, solved_index: int):
		"""agents.py:613-633."""
		if solved_index == 1:
			return
		self.action_queue = deque()
		visited = {1: (None, None)}
		q = deque([1])
		while q:
			v = q.popleft()
			for i, n in enumerate(self.neighbors[v]):
				n = int(n)
				if not n or n in visited:
					continue
				elif n == solved_index:
					self.action_queue.appendleft(i)
					while visited[v][0] is not None:
						self.action_queue.appendleft(visited[v][1])
						v = visited[v][0]
					return
				else:
					visited[n] = (v, i)
					q.append(n)

	def __str__(self):
		return ("BFS" if self.search_graph else "Naive") + f" MCTS (c={self.c})"


class AStarBatch:
	"""Batched weighted A* (agents.py:171-413) for K cubes at once, entirely on the device (SURVEY 8f rows N2 + N3): open
	list, seen-set, G / parent relaxation and the pop of the N cheapest states per search are kernels (csrc/rb_astar.cuh);
	the host only runs the value net on the contiguous batch of new states of all searches and reads two counters per step.
	Every search follows the reference's trace exactly (same pops, same state numbering, same G / parents / action queue).
	20x24 representation."""

	def __init__(self, net, lambda_: float, expansions: int, oh_dtype=torch.float32):
		N.require_cuda()
		if not 0 < expansions <= 1024:
			raise ValueError("expansions must be in 1..1024")
		self.net, self.lambda_, self.expansions = net, float(lambda_), int(expansions)
		# torch.bfloat16 (opt-in, not the reference's dtype): bf16 one-hot rows and a bf16-autocast forward of the value net
		self.oh_dtype = oh_dtype
		self._as_oh = cube._oh_fn("rb_as_oh", oh_dtype)
		self.dev = torch.device("cuda", torch.cuda.current_device())

	def _alloc(self, K: int, max_states: int):
		dev, Nx = self.dev, self.expansions
		M = max_states + 1
		if getattr(self, "_shape", None) == (K, M, Nx):           # buffers of the previous search_many are reused (everything that
			self.parents.zero_(); self.parent_actions.zero_()     # matters is re-initialised by rb_astar_init / rb_hashset_clear)
			self.in_open.zero_()
			return
		self._shape = (K, M, Nx)
		self.K, self.M = K, M
		self.states = torch.empty(K, M, 20, dtype=torch.int8, device=dev)
		self.G = torch.empty(K, M, dtype=torch.float64, device=dev)
		self.parents = torch.zeros(K, M, dtype=torch.int32, device=dev)
		self.parent_actions = torch.zeros(K, M, dtype=torch.uint8, device=dev)
		self.cost = torch.empty(K, M, dtype=torch.float64, device=dev)
		self.in_open = torch.zeros(K, M, dtype=torch.uint8, device=dev)
		self.count = torch.zeros(K, dtype=torch.int32, device=dev)
		self.n_sel = torch.zeros(K, dtype=torch.int32, device=dev)
		self.sel = torch.zeros(K, Nx, dtype=torch.int32, device=dev)
		self.won = torch.zeros(K, dtype=torch.uint8, device=dev)
		self.solved_index = torch.zeros(K, dtype=torch.int32, device=dev)
		self.capacity = _pow2_at_least(2 * K * M)
		self.table = torch.empty(N.lib.rb_hashset_bytes(self.capacity), dtype=torch.uint8, device=dev)
		self.scratch = torch.empty(N.lib.rb_astar_scratch_bytes(K, Nx), dtype=torch.uint8, device=dev)
		P = 12 * Nx
		self.new_states = torch.empty(K * P, 20, dtype=torch.int8, device=dev)
		self.new_search = torch.empty(K * P, dtype=torch.int32, device=dev)
		self.new_index = torch.empty(K * P, dtype=torch.int32, device=dev)
		self.counters = torch.zeros(2, dtype=torch.int32, device=dev)          # n_new_total, n_active
		self.oh = torch.empty(K * P, 480, dtype=self.oh_dtype, device=dev)
		self.view = N.AStarView(K, M, Nx, *(N.ptr(t) for t in (self.states, self.G, self.parents, self.parent_actions, self.cost, self.in_open,
																 self.count, self.n_sel, self.sel, self.won, self.solved_index, self.table)),
								self.capacity, N.ptr(self.scratch))

	@torch.no_grad()
	def search_many(self, states, max_states: int, max_steps: int | None = None):
		"""states: (K, 20) int8 (numpy or CUDA tensor).  Returns (solved bool (K,), action queues: list of K lists,
		len per search int (K,)), the three things `AStar.search` leaves behind for one cube."""
		import ctypes as C
		s = states if isinstance(states, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(states, dtype=np.int8))
		roots = s.to(device=self.dev, dtype=torch.int8).reshape(-1, 20).contiguous()
		self._alloc(roots.shape[0], int(max_states))
		self.net.eval()
		sh = N.stream_handle()
		v = C.byref(self.view)
		N.check(N.lib.rb_hashset_clear(N.ptr(self.table), self.capacity, sh))
		N.check(N.lib.rb_astar_init(v, N.ptr(roots), sh))
		n_total_ptr, n_active_ptr = C.c_void_p(self.counters.data_ptr()), C.c_void_p(self.counters.data_ptr() + 4)
		self.steps = 0
		while max_steps is None or self.steps < max_steps:
			N.check(N.lib.rb_astar_expand(v, int(max_states), N.ptr(self.new_states), N.ptr(self.new_search), N.ptr(self.new_index),
										  n_total_ptr, n_active_ptr, sh))
			n_total, n_active = (int(x) for x in self.counters.tolist())         # the one host sync of a step
			if n_active == 0:
				break
			values = self._values(n_total)
			N.check(N.lib.rb_astar_commit(v, N.ptr(values), self.lambda_, N.ptr(self.new_search), N.ptr(self.new_index), n_total_ptr, sh))
			self.steps += 1
		return self._results()

	def _values(self, n: int) -> torch.Tensor:
		"""agents.py:379-381: one-hot born on the device, value head only; f32 (n,)."""
		if n == 0:
			return torch.zeros(1, dtype=torch.float32, device=self.dev)
		oh = self.oh[:n]
		N.check(self._as_oh(N.REP_2024, N.ptr(self.new_states), N.ptr(oh), n, N.stream_handle()))
		with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.oh_dtype == torch.bfloat16):
			val = self.net(oh, value=True, policy=False)
		return val.reshape(-1).float().contiguous()

	def _results(self):
		won = self.won.bool().cpu().numpy()
		count = self.count.cpu().numpy().astype(int)
		solved_index = self.solved_index.cpu().numpy()
		queues = []
		for s in range(self.K):
			q = []
			if won[s] and solved_index[s] > 1:
				n = count[s] + 1
				parents, actions = self.parents[s, :n].cpu().numpy(), self.parent_actions[s, :n].cpu().numpy()
				i = int(solved_index[s])
				while i != 1:                                            # agents.py:245-251
					q.insert(0, int(actions[i]))
					i = int(parents[i])
			queues.append(q)
		return won, queues, count
