"""
Generates the golden fixtures in this directory by RUNNING THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference, which is read-only and does not
travel to the GPU box):

    python tests/golden/make_golden.py

It imports `librubiks` straight from /root/reference (nothing is copied into the repo),
with two in-memory shims the reference needs under numpy 2.3 / no matplotlib
(SURVEY.md 8c): `ndarray.tostring` -> `tobytes` in agents.py (text replaced before
exec, in memory) and MagicMock stand-ins for matplotlib.  Outputs: small .npz files
holding inputs (host-supplied action draws, fake-net outputs) and the reference's
outputs for every row of SURVEY.md 8(a).
"""
import importlib.util
import json
import os
import sys
import types
from unittest.mock import MagicMock

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)

mpl = MagicMock()
mpl.colors.BASE_COLORS = {}
mpl.colors.TABLEAU_COLORS = {}
for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.animation"):
	sys.modules[name] = mpl

from librubiks import cube  # noqa: E402
from librubiks.cube import maps as ref_maps  # noqa: E402


def load_patched_agents():
	"""agents.py with `.tostring()` -> `.tobytes()` (same bytes; numpy >= 2.3 dropped tostring)."""
	src = open(f"{REF}/librubiks/solving/agents.py").read().replace(".tostring()", ".tobytes()")
	mod = types.ModuleType("librubiks.solving.agents")
	mod.__file__ = f"{REF}/librubiks/solving/agents.py"
	sys.modules["librubiks.solving.agents"] = mod
	exec(compile(src, mod.__file__, "exec"), mod.__dict__)
	return mod


agents = load_patched_agents()


def save(name, **arrays):
	path = os.path.join(HERE, name + ".npz")
	np.savez_compressed(path, **arrays)
	print(f"{name}.npz  {os.path.getsize(path) / 1024:.1f} KiB  keys={len(arrays)}")


class FakeNet(torch.nn.Module):
	"""Deterministic stand-in for the value/policy net: integer weights so the f32 sums are
	exact whatever the summation order, quantised so that ties are frequent.  Records every
	value batch it returns."""

	def __init__(self, width, seed, quant=4.0):
		super().__init__()
		g = np.random.RandomState(seed)
		self.w = torch.from_numpy(g.randint(-6, 7, size=(width,)).astype(np.float32))
		self.wp = torch.from_numpy(g.randint(-3, 4, size=(width, 12)).astype(np.float32))
		self.quant = quant
		self.value_log = []

	def forward(self, x, policy=True, value=True):
		x = x.float().cpu()
		out = []
		if policy:
			out.append(x @ self.wp)
		if value:
			v = (torch.floor((x @ self.w) / self.quant)).unsqueeze(1)
			self.value_log.append(v.squeeze(1).numpy().copy())
			out.append(v)
		return out if len(out) > 1 else out[0]


def gen_tables():
	m = ref_maps.get_tensor_map(np.int8)
	js = json.load(open(f"{REF}/frontend/src/assets/maps.json"))
	cube.set_is2024(True)
	s2024 = cube.get_solved()
	cube.set_is2024(False)
	s686 = cube.get_solved()
	perm = np.stack([cube.rotate(np.arange(48).reshape(6, 8, 1), *cube.action_space[a]).reshape(48) for a in range(12)])
	cube.set_is2024(True)
	save("tables", delta_maps=m, json_map_neg=np.array(js["map_neg"], dtype=np.int8), json_map_pos=np.array(js["map_pos"], dtype=np.int8),
		 solved2024=s2024, solved686=s686, perm686=perm.astype(np.uint8),
		 iter_actions2=cube.iter_actions(2), idx2act_faces=cube.indices_to_actions(np.arange(12))[0],
		 idx2act_dirs=cube.indices_to_actions(np.arange(12))[1], rev_actions=cube.rev_actions(np.arange(12)),
		 action_space=np.array(cube.action_space))


def gen_dynamics():
	out = {}
	for is2024, tag, n, depth in ((True, "2024", 256, 40), (False, "686", 48, 30)):
		cube.set_is2024(is2024)
		g = np.random.RandomState(7 if is2024 else 8)
		faces = g.randint(0, 6, (depth, n))
		dirs = g.randint(0, 2, (depth, n))
		states = np.array([cube.get_solved()] * n)
		per_step = []
		for d in range(depth):
			states = cube.multi_rotate(states, faces[d], dirs[d])
			per_step.append(states)
		# multi_rotate == rotate state by state (positive directions included)
		single = np.array([cube.rotate(s, f, d) for s, f, d in zip(per_step[-2], faces[-1], dirs[-1])])
		assert (single == per_step[-1]).all()
		final = per_step[-1]
		f12, d12 = cube.iter_actions(len(final))
		children = cube.multi_rotate(np.repeat(final, 12, axis=0), f12, d12)
		out.update({
			f"faces{tag}": faces.astype(np.uint8), f"dirs{tag}": dirs.astype(np.uint8),
			f"steps{tag}": np.stack(per_step).astype(np.int8), f"children{tag}": children.astype(np.int8),
			f"solved_flags{tag}": cube.multi_is_solved(np.concatenate([final, np.array([cube.get_solved()])])),
			f"oh{tag}": cube.as_oh(final).numpy().astype(np.uint8),
			f"oh_single{tag}": cube.as_oh(final[0]).numpy().astype(np.uint8),
			f"strings{tag}": np.array([cube.stringify(s) for s in final[:8]]),
			f"as633_{tag}": np.stack([cube.as633(s) for s in final[:8]]),
		})
		if not is2024:
			out["correct686"] = cube.as_correct(cube.as_oh(final)).numpy()
	cube.set_is2024(True)
	save("dynamics", **out)


def gen_scramblers():
	out = {}
	for is2024, tag in ((True, "2024"), (False, "686")):
		cube.set_is2024(is2024)
		for games, depth, with_solved in ((3, 4, True), (5, 7, False), (4, 1, True)):
			key = f"{tag}_{games}_{depth}_{int(with_solved)}"
			np.random.seed(0)
			faces = np.random.randint(0, 6, (depth, games))
			dirs = np.random.randint(0, 2, (depth, games))
			np.random.seed(0)
			states, oh = cube.sequence_scrambler(games, depth, with_solved)
			out[f"seq_faces_{key}"], out[f"seq_dirs_{key}"] = faces.astype(np.uint8), dirs.astype(np.uint8)
			out[f"seq_states_{key}"], out[f"seq_oh_{key}"] = states.astype(np.int8), oh.numpy().astype(np.uint8)
		np.random.seed(42)
		faces = np.random.randint(6, size=(20,))
		dirs = np.random.randint(2, size=(20,))
		np.random.seed(42)
		state, f2, d2 = cube.scramble(20)
		assert (faces == f2).all() and (dirs == d2).all()
		out[f"scr_faces_{tag}"], out[f"scr_dirs_{tag}"], out[f"scr_state_{tag}"] = faces.astype(np.uint8), dirs.astype(np.uint8), state
	cube.set_is2024(True)
	save("scramblers", **out)


def gen_adi():
	from librubiks.train import Train
	from librubiks.utils import TickTock
	out = {}
	for is2024, tag, width in ((True, "2024", 480), (False, "686", 288)):
		cube.set_is2024(is2024)
		for method in ("paper", "lapanfix", "schultzfix", "reward0"):
			for ai, alpha in enumerate((0.0, 0.3, 1.0)):
				if not is2024 and ai != 1:
					continue
				games, depth = (6, 5) if is2024 else (3, 4)
				t = object.__new__(Train)
				t.rollout_games, t.rollout_depth, t.reward_method = games, depth, method
				t.adi_ff_batches, t.with_analysis, t.tt = 2, False, TickTock()
				net = FakeNet(width, seed=11 + ai)
				seed = 100 + ai
				np.random.seed(seed)
				faces = np.random.randint(0, 6, (depth, games))
				dirs = np.random.randint(0, 2, (depth, games))
				np.random.seed(seed)
				oh_states, policy, value, lw = t.ADI_traindata(net, alpha)
				key = f"{tag}_{method}_{ai}"
				out[f"faces_{key}"], out[f"dirs_{key}"] = faces.astype(np.uint8), dirs.astype(np.uint8)
				out[f"values_{key}"] = np.concatenate(net.value_log).astype(np.float32)
				out[f"oh_states_{key}"] = oh_states.cpu().numpy().astype(np.uint8)
				out[f"policy_{key}"], out[f"value_{key}"] = policy.numpy(), value.numpy()
				out[f"lw_{key}"], out[f"alpha_{key}"] = lw.numpy(), np.float64(alpha)
				out[f"w_{key}"] = net.w.numpy()
	# an all-tie case (constant net, `nn_init` number in runtrain.py:88-92): every argmax row is a tie
	cube.set_is2024(True)
	for method in ("paper", "lapanfix"):
		t = object.__new__(Train)
		t.rollout_games, t.rollout_depth, t.reward_method = 4, 3, method
		t.adi_ff_batches, t.with_analysis, t.tt = 1, False, TickTock()
		net = FakeNet(480, seed=1)
		net.w[:] = 0
		np.random.seed(5)
		faces = np.random.randint(0, 6, (3, 4)); dirs = np.random.randint(0, 2, (3, 4))
		np.random.seed(5)
		_, policy, value, lw = t.ADI_traindata(net, 0.5)
		out[f"tie_faces_{method}"], out[f"tie_dirs_{method}"] = faces.astype(np.uint8), dirs.astype(np.uint8)
		out[f"tie_policy_{method}"], out[f"tie_value_{method}"], out[f"tie_lw_{method}"] = policy.numpy(), value.numpy(), lw.numpy()
	# loss weights at the benchmark shapes (f64 -> f32 rounding pinned)
	for games, depth, alpha in ((1000, 25, 0.3), (7500, 30, 0.7), (17, 999, 0.05)):
		weighted = np.tile(1 / np.arange(1, depth + 1), games)
		ws, us = weighted.sum(), len(weighted)
		out[f"lwbig_{games}_{depth}"] = torch.from_numpy(((1 - alpha) * weighted / ws + alpha * np.ones_like(weighted) / us) * (ws + us)).float().numpy()
		out[f"lwbig_ws_{games}_{depth}"] = np.float64(ws)
	save("adi", **out)


def gen_search():
	out = {}
	cube.set_is2024(True)
	# BFS layer sizes with the reference's multi_rotate + a Python set (SURVEY 8c KAT i)
	frontier = np.array([cube.get_solved()])
	seen = {frontier[0].tobytes()}
	counts = [1]
	for _ in range(5):
		ch = cube.multi_rotate(np.repeat(frontier, 12, axis=0), *cube.iter_actions(len(frontier)))
		nxt = []
		for s in ch:
			k = s.tobytes()
			if k not in seen:
				seen.add(k); nxt.append(s)
		frontier = np.array(nxt); counts.append(len(nxt))
	out["bfs_layer_counts"] = np.array(counts)
	# BFS agent on a depth-4 scramble: states explored + action queue
	np.random.seed(3)
	state, _, _ = cube.scramble(4, True)
	bfs = agents.BFS()
	ok = bfs.search(state, None, 10 ** 6)
	out["bfs_start"], out["bfs_found"], out["bfs_len"], out["bfs_queue"] = state, np.bool_(ok), np.int64(len(bfs)), np.array(bfs.action_queue)
	# A*: trace of expand_batch with an integer fake net (ties are frequent -> index tie-breaks matter)
	for is2024, tag, width in ((True, "2024", 480), (False, "686", 288)):
		cube.set_is2024(is2024)
		np.random.seed(9)
		state, _, _ = cube.scramble(9 if is2024 else 6, True)
		net = FakeNet(width, seed=21)
		a = agents.AStar(net, lambda_=0.16, expansions=7)
		a.reset(None, 10 ** 6)
		a.indices[state.tobytes()], a.states[1], a.G[1] = 1, state, 0
		a.open_queue = [(0, 1)]
		import heapq
		lens, batches = [], []
		won = False
		for step in range(12):
			n = min(len(a.open_queue), a.expansions)
			idcs = np.array([heapq.heappop(a.open_queue)[1] for _ in range(n)], dtype=int)
			batches.append(np.pad(idcs, (0, 7 - len(idcs)), constant_values=-1))
			won = a.expand_batch(idcs)
			lens.append(len(a))
			if won:
				break
		L = len(a)
		out[f"astar_start_{tag}"], out[f"astar_w_{tag}"] = state, net.w.numpy()
		out[f"astar_batches_{tag}"], out[f"astar_lens_{tag}"], out[f"astar_won_{tag}"] = np.array(batches), np.array(lens), np.bool_(won)
		out[f"astar_states_{tag}"], out[f"astar_G_{tag}"] = a.states[1:L + 1].copy(), a.G[1:L + 1].copy()
		out[f"astar_parents_{tag}"], out[f"astar_pact_{tag}"] = a.parents[2:L + 1].copy(), a.parent_actions[2:L + 1].copy()
		out[f"astar_open_{tag}"] = np.array(sorted(a.open_queue))
	# A* full search on shallow scrambles: solved flag + action queue
	cube.set_is2024(True)
	np.random.seed(1)
	state, _, _ = cube.scramble(4, True)
	zero_net = FakeNet(480, seed=21)
	zero_net.w[:] = 0       # h == 0: weighted A* degenerates to uniform-cost search, ties broken by index
	a = agents.AStar(zero_net, lambda_=1.0, expansions=5)
	ok = a.search(state, None, 20000)
	assert ok
	out["astar_full_start"], out["astar_full_ok"], out["astar_full_queue"], out["astar_full_len"] = state, np.bool_(ok), np.array(a.action_queue), np.int64(len(a))
	# MCTS: expand_leaf bookkeeping (indices, neighbours) after a short search
	np.random.seed(2)
	state, _, _ = cube.scramble(6, True)
	m = agents.MCTS(FakeNet(480, seed=4), c=0.6, search_graph=False)
	m.search(state, None, 200)
	L = len(m)
	out["mcts_start"], out["mcts_len"] = state, np.int64(L)
	out["mcts_states"], out["mcts_neighbors"], out["mcts_leaves"] = m.states[1:L + 1].copy(), m.neighbors[:L + 1].copy(), m.leaves[:L + 1].copy()
	# MCTS full traces with a uniform policy (logits 0 -> softmax exactly 1/12 on any machine) and an integer value net
	# that counts the cubies in their solved place (exact in f32, tie-heavy):
	# (a) graph search that finds the solution after ~2300 states (_complete_graph over ~2100 leaves, then
	#     _shorten_action_queue), (b) tree search cut off by max_states
	for tag, depth, seed, c, graph, max_states in (("a", 3, 2, 5.0, True, 3000), ("b", 4, 0, 0.6, False, 400)):
		np.random.seed(seed)
		state, _, _ = cube.scramble(depth, True)
		net = FakeNet(480, seed=11, quant=1.0)
		net.wp[:] = 0
		net.w[:] = cube.as_oh(cube.get_solved()).cpu().reshape(-1)
		m = agents.MCTS(net, c=c, search_graph=graph)
		ok = m.search(state, None, max_states)
		L = len(m)
		out[f"mcts{tag}_start"], out[f"mcts{tag}_ok"], out[f"mcts{tag}_len"] = state, np.bool_(ok), np.int64(L)
		out[f"mcts{tag}_queue"] = np.array(m.action_queue, dtype=np.int64)
		out[f"mcts{tag}_states"], out[f"mcts{tag}_neighbors"] = m.states[1:L + 1].copy(), m.neighbors[:L + 1].copy()
		out[f"mcts{tag}_leaves"], out[f"mcts{tag}_N"] = m.leaves[:L + 1].copy(), m.N[:L + 1].copy()
		out[f"mcts{tag}_W"], out[f"mcts{tag}_V"] = m.W[1:L + 1].copy(), m.V[1:L + 1].copy()
	save("search", **out)


def gen_bfs_budget():
	"""BFS.search (agents.py:96-123) with a BINDING state budget: the reference tests `len(self) < max_states` before every
	parent pop (agents.py:104), so a search that runs out of budget stops in the middle of a layer.  Records found flag,
	len(agent) and the action queue for scrambles of depth 3-6 under budgets that end the search at every stage (first layer,
	mid layer, exactly on a parent boundary, solved child just inside / just outside the admitted parents)."""
	cube.set_is2024(True)
	out, case = {}, 0
	for depth in (3, 4, 5, 6):
		for seed in (0, 1, 2):
			np.random.seed(1000 * depth + seed)
			state, _, _ = cube.scramble(depth, True)
			for max_states in (1, 2, 13, 14, 50, 100, 127, 1000, 1195, 2000, 5000, 20000):
				bfs = agents.BFS()
				ok = bfs.search(state, None, max_states)
				out[f"start_{case}"], out[f"max_{case}"] = state, np.int64(max_states)
				out[f"found_{case}"], out[f"len_{case}"] = np.bool_(ok), np.int64(len(bfs))
				out[f"queue_{case}"] = np.array(bfs.action_queue, dtype=np.int64)
				case += 1
	# 6x8x6: same loop through cube.* under the other representation
	cube.set_is2024(False)
	for seed, depth in ((0, 3), (1, 4)):
		np.random.seed(77 + seed)
		state, _, _ = cube.scramble(depth, True)
		for max_states in (14, 100, 1000, 3000):
			bfs = agents.BFS()
			ok = bfs.search(state, None, max_states)
			out[f"start_{case}"], out[f"max_{case}"] = state, np.int64(max_states)
			out[f"found_{case}"], out[f"len_{case}"] = np.bool_(ok), np.int64(len(bfs))
			out[f"queue_{case}"] = np.array(bfs.action_queue, dtype=np.int64)
			case += 1
	cube.set_is2024(True)
	out["n_cases"] = np.int64(case)
	save("bfs_budget", **out)


def gen_train():
	"""Train.train (train.py:111-247) run end to end on the CPU with a narrowed fc net (same layer pattern as fc_small,
	model.py:117-161, sizes 480 -> 64 -> 32 -> {16 -> 12, 16 -> 1} so that the initial weights fit a fixture): the per-rollout
	losses, the lr / alpha schedule it implies and the final parameters.  tau < 1 exercises the generator-net mix, the 30
	states of a rollout split into minibatches of 16 + 14."""
	from librubiks.model import Model, ModelConfig
	from librubiks.train import Train
	cube.set_is2024(True)
	ModelConfig._fc_small_arch = {"shared_sizes": [64, 32], "part_sizes": [16]}
	out = {}
	for tag, kw in (("a", dict(rollouts=4, batch_size=16, rollout_games=6, rollout_depth=5, alpha_update=0.25, lr=5e-3, gamma=0.5,
							   update_interval=1, tau=0.5, reward_method="lapanfix")),
					("b", dict(rollouts=3, batch_size=20, rollout_games=5, rollout_depth=4, alpha_update=1, lr=1e-2, gamma=1,
							   update_interval=2, tau=1, reward_method="paper"))):
		torch.manual_seed(7)
		net = Model.create(ModelConfig())
		init = {k: v.clone().numpy() for k, v in net.state_dict().items()}

		class _Agent:
			net = None
		t = Train(optim_fn=torch.optim.Adam, agent=_Agent(), evaluator=None, evaluation_interval=0, with_analysis=False, **kw)
		# (batch_size=0, "one batch per rollout", crashes in the reference: train.py:58 reads states_per_rollout before
		# train.py:125 sets it -- case b passes the full rollout size instead)
		np.random.seed(42)
		trained, _ = t.train(net)
		for k, v in init.items():
			out[f"{tag}_init_{k}"] = v
		for k, v in trained.state_dict().items():
			out[f"{tag}_final_{k}"] = v.numpy()
		out[f"{tag}_policy_losses"], out[f"{tag}_value_losses"], out[f"{tag}_train_losses"] = t.policy_losses, t.value_losses, t.train_losses
		out[f"{tag}_kw"] = np.array(json.dumps(kw))
	save("train", **out)


def gen_eval():
	"""Evaluator.eval (evaluation.py:54-94) with the A* agent and the integer fake value net: results (turns or -1) and
	states explored per game, for fixed depths and for the deep (depth ~ U[100, 999]) mode."""
	from librubiks.solving.evaluation import Evaluator
	cube.set_is2024(True)
	out = {}
	for tag, depths, n_games, max_states in (("fixed", [1, 3, 5], 3, 3000), ("deep", range(0), 2, 500)):
		net = FakeNet(480, seed=11)
		agent = agents.AStar(net, lambda_=0.2, expansions=20)
		ev = Evaluator(n_games=n_games, scrambling_depths=depths, max_time=None, max_states=max_states)
		np.random.seed(9)
		res, states, _ = ev.eval(agent)
		out[f"{tag}_res"], out[f"{tag}_states"], out[f"{tag}_w"] = res, states, net.w.numpy()
		out[f"{tag}_max_states"], out[f"{tag}_n_games"] = np.int64(max_states), np.int64(n_games)
		out[f"{tag}_depths"] = np.array(list(depths), dtype=np.int64)
	save("evaluation", **out)


if __name__ == "__main__":
	torch.manual_seed(0)
	which = sys.argv[1:] or ["tables", "dynamics", "scramblers", "adi", "search", "bfs_budget", "train", "eval"]
	for name in which:
		globals()["gen_" + name]()
