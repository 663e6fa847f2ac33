"""
Device-resident Autodidactic-Iteration training loop: mirror of `Train.train` (reference: librubiks/train.py:111-247)
around the fused ADI generator.

What changes against the reference: the ADI batch is born on the GPU (`adi.ADIGenerator`, buffers allocated once), so the
four `.to(gpu)` copies of train.py:156-159 and the `.cpu()` of train.py:304 are gone, and the per-minibatch
`.detach().cpu().numpy().mean()` (train.py:178-179, one host sync per minibatch) becomes one device accumulation read back
once per rollout.  What stays the reference's, because identical seeds must give identical losses: the order of the rollout (generator-net mix
train.py:341-352, ADI batch, minibatch pass, schedules, evaluation), the consumption of the global numpy stream (the draws of
cube.py:226-227 and the unused shuffle of train.py:405), loss = mean((CE + MSE) * loss_weights), the lr / alpha schedule
(train.py:191-202), evaluation rollouts (train.py:63-73) and best-net bookkeeping (train.py:211-227).  The network forward /
backward stays torch (dense GEMMs).

Data-parallel use (one process per GPU): every rank generates its own share of the games from its own random stream and
gradients are averaged with one flat NCCL all-reduce per minibatch (`sharding.allreduce_mean_`); there is no collective on
the ADI path.  See `Train.train`.
"""
from __future__ import annotations

import copy

import numpy as np
import torch

from . import _native as N
from . import adi, sharding


def _clone(net):
	"""Model.clone() (model.py:163-169) when the net has it, else a deep copy."""
	return net.clone() if hasattr(net, "clone") else copy.deepcopy(net)


class Train:
	"""Same constructor arguments as the reference's `Train` (train.py:28-46) minus logging / analysis; `agent` and
	`evaluator` are optional and duck-typed (`evaluator.eval(agent) -> (results, states, times)`, `agent.net`)."""

	def __init__(self, rollouts: int, batch_size: int, rollout_games: int, rollout_depth: int, optim_fn, alpha_update: float,
				 lr: float, gamma: float, update_interval: int, tau: float, reward_method: str, agent=None, evaluator=None,
				 evaluation_interval: int = 0, policy_criterion=torch.nn.CrossEntropyLoss, value_criterion=torch.nn.MSELoss,
				 data_parallel: bool = False, log=None, oh_dtype=torch.float32):
		N.require_cuda()
		self.rollouts = int(rollouts)
		self.train_rollouts = np.arange(self.rollouts)
		self.rollout_games, self.rollout_depth = int(rollout_games), int(rollout_depth)
		self.states_per_rollout = self.rollout_depth * self.rollout_games
		self.batch_size = self.states_per_rollout if not batch_size else int(batch_size)      # train.py:58
		self.adi_ff_batches = 1
		self.reward_method = reward_method
		self.evaluation_rollouts = self.evaluation_schedule(self.rollouts, evaluation_interval)
		self.agent, self.evaluator = agent, evaluator
		self.tau, self.alpha_update, self.lr, self.gamma, self.update_interval = tau, alpha_update, lr, gamma, update_interval
		self.optim = optim_fn
		self.policy_criterion = policy_criterion(reduction="none")
		self.value_criterion = value_criterion(reduction="none")
		self.data_parallel = bool(data_parallel)
		# torch.bfloat16 (not in the reference): one-hot batches are emitted as bf16 and every forward runs under bf16 autocast
		self.oh_dtype = oh_dtype
		self.log = log or (lambda *a, **k: None)
		self.alphas, self.lrs = [], []

	@staticmethod
	def evaluation_schedule(rollouts: int, evaluation_interval: int) -> np.ndarray:
		"""train.py:63-73: evaluate every `evaluation_interval` rollouts and after the last one."""
		if not evaluation_interval:
			return np.array([])
		ev = np.arange(0, rollouts, evaluation_interval) - 1
		if evaluation_interval == 1:
			ev = ev[1:]
		else:
			ev[0] = 0
		if not len(ev) or rollouts - 1 != ev[-1]:
			ev = np.append(ev, rollouts - 1)
		return ev

	@staticmethod
	def _get_batches(size: int, bsize: int):
		"""Minibatch bounds of one rollout, in order: ceil(size / bsize) slices, the last one short (train.py:400-410).  The
		reference also shuffles an index array here that it never uses for indexing; the call is kept because it advances the
		global numpy stream that the next rollout's scramble draws come from (identical seeds must give identical scrambles)."""
		np.random.shuffle(np.arange(size))
		return [slice(lo, min(lo + bsize, size)) for lo in range(0, size, bsize)]

	def _advance_alpha(self, alpha):
		"""One step of the loss-weighting schedule (train.py:196-202): alpha grows by `alpha_update` until it reaches 1; a step
		that would overshoot lands on 1, one that lands within float tolerance of 1 keeps its value."""
		if not self.alpha_update:
			return alpha
		nxt = alpha + self.alpha_update
		if nxt <= 1 or np.isclose(nxt, 1):
			return nxt
		return 1 if alpha < 1 else alpha

	@torch.no_grad()
	def _update_gen_net(self, generator_net, net):
		"""train.py:341-352: generator <- tau * net + (1 - tau) * generator, over the whole state dict (buffers included)."""
		gen, cur = generator_net.state_dict(), net.state_dict()
		for name, p in cur.items():
			gen[name].data.copy_(self.tau * p.data + (1 - self.tau) * gen[name].data)
		generator_net.load_state_dict(gen)
		return generator_net

	def ADI_traindata(self, net, alpha: float):
		"""train.py:256-339 on the device (rl_rubiks_b200.adi): (oh_states, policy_targets, value_targets, loss_weights)."""
		with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.oh_dtype == torch.bfloat16):
			return adi.adi_traindata(net, self._generator.games, self.rollout_depth, self.reward_method, alpha,
									 ff_batches=self.adi_ff_batches, generator=self._generator, rng=self._draw_rng)

	def _sgd_pass(self, net, optimizer, params, batch, acc):
		"""One pass over a rollout's batch in minibatches (train.py:165-179): loss = mean((CE + MSE) * loss_weights); the two
		per-minibatch loss means are accumulated in f64 on the device (one host read per rollout instead of one per minibatch)."""
		oh, policy_t, value_t, weights = batch
		rank, ws = sharding.world() if self.data_parallel else (0, 1)
		if ws == 1:
			bounds = self._get_batches(oh.shape[0], self.batch_size)
		else:
			# every rank must run the same number of minibatches (one gradient all-reduce each) although shards may differ by a
			# game: the count is taken from the largest shard, the rank's own states are cut into that many near-equal slices; the
			# unused shuffle consumes the global numpy stream for the WHOLE rollout's size, identically on every rank
			np.random.shuffle(np.arange(self.rollout_games * self.rollout_depth))
			bounds = sharding.dp_minibatch_bounds(oh.shape[0], self.rollout_games, self.rollout_depth, self.batch_size, ws)
		acc.zero_()
		net.train()
		for sl in bounds:
			optimizer.zero_grad()
			if sl.stop > sl.start:
				with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.oh_dtype == torch.bfloat16):
					policy_pred, value_pred = net(oh[sl], policy=True, value=True)
				w = weights[sl]
				policy_loss = self.policy_criterion(policy_pred.float(), policy_t[sl]) * w
				value_loss = self.value_criterion(value_pred.float().squeeze(), value_t[sl]) * w
				torch.mean(policy_loss + value_loss).backward()
				acc[0] += policy_loss.detach().mean().double() / len(bounds)
				acc[1] += value_loss.detach().mean().double() / len(bounds)
			else:                                            # fewer states than minibatches on this rank: it joins the exchange with zeros
				for p in params:
					p.grad = torch.zeros_like(p)
			if self.data_parallel:
				sharding.allreduce_mean_([p.grad for p in params if p.grad is not None])
			optimizer.step()
		return acc.tolist()

	def train(self, net):
		"""Returns (net after the last rollout, net with the best evaluation score), as train.py:111-247.

		Data-parallel (`data_parallel=True`, one process per GPU): this rank generates `shard_bounds(rollout_games)` of the games
		from its OWN random stream (`sharding.rank_seed` of one number taken from the global numpy stream, so the global
		stream -- which `Evaluator.eval_batched` needs identical on every rank -- stays in step), trains on its share, averages
		the gradients of every minibatch with one all-reduce and the floating-point module buffers once per rollout."""
		dev = torch.device("cuda", torch.cuda.current_device())
		rank, ws = sharding.world() if self.data_parallel else (0, 1)
		lo, hi = sharding.shard_bounds(self.rollout_games, ws, rank)
		self._generator = adi.ADIGenerator(hi - lo, self.rollout_depth, self.reward_method, oh_dtype=self.oh_dtype)
		self._draw_rng = np.random.RandomState(sharding.rank_seed(int(np.random.randint(2 ** 31 - 1)), rank)) if ws > 1 else None
		best_solve, best_net = 0, _clone(net)
		if self.agent is not None:
			self.agent.net = net
		generator_net = _clone(net)
		alpha = 1 if self.alpha_update == 1 else 0
		optimizer = self.optim(net.parameters(), lr=self.lr)
		lr_scheduler = torch.optim.lr_scheduler.StepLR(optimizer, 1, self.gamma)
		self.policy_losses, self.value_losses, self.train_losses = np.zeros(self.rollouts), np.zeros(self.rollouts), np.empty(self.rollouts)
		self.sol_percents, self.alphas, self.lrs = [], [], []
		params = [p for p in net.parameters() if p.requires_grad]
		acc = torch.zeros(2, dtype=torch.float64, device=dev)

		for rollout in range(self.rollouts):
			if self.tau != 1:
				generator_net = self._update_gen_net(generator_net, net)
			self.alphas.append(float(alpha))
			self.lrs.append(float(optimizer.param_groups[0]["lr"]))
			batch = self.ADI_traindata(generator_net if self.tau != 1 else net, alpha)
			self.policy_losses[rollout], self.value_losses[rollout] = self._sgd_pass(net, optimizer, params, batch, acc)
			self.train_losses[rollout] = self.policy_losses[rollout] + self.value_losses[rollout]
			if ws > 1:
				sharding.allreduce_mean_([b for b in net.buffers() if b.is_floating_point()])

			if rollout and self.update_interval and rollout % self.update_interval == 0:       # train.py:191-202
				if self.gamma != 1:
					lr_scheduler.step()
				alpha = self._advance_alpha(alpha)
			self.log(f"Rollout {rollout} completed with mean loss {self.train_losses[rollout]}")

			if rollout in self.evaluation_rollouts and self.evaluator is not None and self.agent is not None:   # train.py:211-227
				net.eval()
				self.agent.net = net
				solved = np.asarray(self.evaluator.eval(self.agent)[0]) != -1
				self.sol_percents.append(solved.mean())
				if self.sol_percents[-1] > best_solve:
					best_solve, best_net = self.sol_percents[-1], _clone(net)
		return net, best_net
