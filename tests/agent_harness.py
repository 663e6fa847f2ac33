"""Trace harnesses (test infrastructure, not product): the HOST side of the reference's single-search A* and MCTS agents
(librubiks/solving/agents.py:171-413, 415-645 -- heapq open list, numpy G / parent relaxation, the N / W / L / P / V node
arrays and the sequential UCB walk) replayed around the device primitives of rl_rubiks_b200.frontier (`StateHashSet.expand`,
`insert_unique`, `lookup`, `rb_expand12`, `rb_as_oh`).  They exist so that those primitives can be checked against traces
recorded from the reference itself (tests/golden/search.npz): same node numbering, same neighbour tables, same action
queues.  The host bookkeeping below follows the reference statement by statement on purpose; the product's own agents are
`frontier.BFS` and `frontier.AStarBatch`."""
from __future__ import annotations

import heapq
from collections import deque
from time import perf_counter

import numpy as np
import torch

from rl_rubiks_b200 import _native as N
from rl_rubiks_b200 import cube
from rl_rubiks_b200.frontier import Agent, StateHashSet, read_count


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
