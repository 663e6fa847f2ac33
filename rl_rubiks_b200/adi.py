"""
ADI training-batch generation on the device: drop-in for the data path of `Train.ADI_traindata`
(reference: librubiks/train.py:256-339).

One fused kernel produces the scrambled states, their 12 children, both one-hot tensors and the solved flags
(train.py:277-296); the value net forward stays a torch call (dense GEMM); a second kernel assembles the
policy / value targets (train.py:313-325) and a third the loss weights (train.py:329-333).  Nothing leaves the
device and there is no per-state Python loop.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _native as N
from . import cube


class ADIGenerator:
	"""Pre-allocates every buffer of one ADI rollout (games x depth states) on the current device."""

	def __init__(self, games: int, depth: int, reward_method: str = "lapanfix", keep_states: bool = False,
				 keep_children: bool = False, is2024: bool | None = None, oh_dtype=torch.float32):
		N.require_cuda()
		if reward_method not in N.REWARD_METHODS:
			raise KeyError(f"reward_method must be one of {list(N.REWARD_METHODS)}, got {reward_method!r}")
		self.games, self.depth, self.reward_method = int(games), int(depth), reward_method
		self.is2024 = cube.get_is2024() if is2024 is None else is2024
		self.rep = N.REP_2024 if self.is2024 else N.REP_686
		self.width = 480 if self.is2024 else 288
		shape = (20,) if self.is2024 else (6, 8, 6)
		n = self.n = self.games * self.depth
		dev = torch.device("cuda", torch.cuda.current_device())
		self.actions = torch.empty(self.depth, self.games, dtype=torch.uint8, device=dev)
		self.oh_dtype = oh_dtype                 # torch.float32 (the reference's) or torch.bfloat16 (opt-in, same 0/1 rows)
		self._generate_fn = cube._oh_fn("rb_adi_generate", oh_dtype)
		self.oh_states = torch.empty(n, self.width, dtype=oh_dtype, device=dev)
		self.children_oh = torch.empty(12 * n, self.width, dtype=oh_dtype, device=dev)
		self.solved_states = torch.empty(n, dtype=torch.uint8, device=dev)
		self.solved_children = torch.empty(12 * n, dtype=torch.uint8, device=dev)
		self.states = torch.empty(n, *shape, dtype=torch.int8, device=dev) if keep_states else None
		self.children = torch.empty(12 * n, *shape, dtype=torch.int8, device=dev) if keep_children else None
		self.policy_targets = torch.empty(n, dtype=torch.int64, device=dev)
		self.value_targets = torch.empty(n, dtype=torch.float32, device=dev)
		self.loss_weights = torch.empty(n, dtype=torch.float32, device=dev)
		# f64 sum of the 1/d weights, obtained the reference's way (numpy pairwise sum, train.py:330-332)
		self.weight_sum = float(np.tile(1 / np.arange(1, self.depth + 1), self.games).sum())

	@property
	def with_solved(self) -> bool:
		return self.reward_method == "lapanfix"      # train.py:277

	def set_actions(self, faces, dirs=None):
		"""Host- or device-supplied draws of shape (depth, games): (faces, dirs) or action indices."""
		if dirs is None:
			a = faces if isinstance(faces, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(faces).astype(np.uint8)))
		else:
			a = torch.from_numpy(cube._actions_u8(faces, dirs))
		if tuple(a.shape) != (self.depth, self.games):
			raise IndexError(f"actions must have shape (depth, games) = {(self.depth, self.games)}, got {tuple(a.shape)}")
		self.actions.copy_(a.to(torch.uint8), non_blocking=True)

	def draw_actions(self, rng=None):
		"""The reference's draw (cube.py:226-227): faces (depth, games) then dirs, from the global numpy stream or from `rng`
		(a RandomState: the per-rank stream of data-parallel training)."""
		rng = rng or np.random
		faces = rng.randint(0, 6, (self.depth, self.games))
		dirs = rng.randint(0, 2, (self.depth, self.games))
		self.set_actions(faces, dirs)

	def generate(self):
		"""Kernel A (train.py:277-296).  Returns (oh_states, children_oh)."""
		N.check(self._generate_fn(self.rep, N.ptr(self.actions), None, self.games, self.depth, int(self.with_solved),
									  N.ptr(self.states), N.ptr(self.oh_states), N.ptr(self.children), N.ptr(self.children_oh),
									  N.ptr(self.solved_states), N.ptr(self.solved_children), N.stream_handle()))
		return self.oh_states, self.children_oh

	def targets(self, values: torch.Tensor, alpha: float):
		"""Kernel B (train.py:313-333: targets and loss weights in one launch) from the net's values for the children, f32 (12 n,)."""
		v = values.reshape(-1)
		if v.dtype != torch.float32 or not v.is_cuda or v.numel() != 12 * self.n:
			raise IndexError("values must be a float32 CUDA tensor with 12 * games * depth elements")
		v = v.contiguous()
		s = N.stream_handle()
		N.check(N.lib.rb_adi_targets_weights(N.ptr(v), N.ptr(self.solved_children), N.ptr(self.solved_states), self.games, self.depth,
											 N.REWARD_METHODS[self.reward_method], float(alpha), self.weight_sum,
											 N.ptr(self.policy_targets), N.ptr(self.value_targets), N.ptr(self.loss_weights), s))
		return self.policy_targets, self.value_targets, self.loss_weights


def _value_forward(net, x: torch.Tensor, ff_batches: int) -> torch.Tensor:
	"""train.py:249-254, 301-311: the value head over the children in `ff_batches` slices."""
	slice_size = x.shape[0] // ff_batches + 1
	parts = [net(x[i * slice_size:(i + 1) * slice_size], policy=False, value=True).squeeze() for i in range(ff_batches)]
	return torch.cat([p.reshape(-1) for p in parts]).float()


@torch.no_grad()
def adi_traindata(net, games: int, depth: int, reward_method: str, alpha: float, faces=None, dirs=None,
				  ff_batches: int = 1, generator: ADIGenerator | None = None, rng=None):
	"""Drop-in for `Train.ADI_traindata(net, alpha)` (train.py:256-339): returns
	(oh_states f32 (n, W), policy_targets i64 (n,), value_targets f32 (n,), loss_weights f32 (n,)), all on the GPU.
	`faces`/`dirs` (depth, games) override the random draw (identical host-supplied actions give bit-identical
	batches to the reference's numpy path for identical net outputs)."""
	g = generator or ADIGenerator(games, depth, reward_method)
	net.eval()
	if faces is None:
		g.draw_actions(rng)
	else:
		g.set_actions(faces, dirs)
	oh_states, children_oh = g.generate()
	values = _value_forward(net, children_oh, ff_batches)
	policy, value, weights = g.targets(values, alpha)
	return oh_states, policy, value, weights
