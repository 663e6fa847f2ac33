"""Host logic of the training-loop mirror (rl_rubiks_b200/train.py) that needs no GPU: the evaluation schedule and the
minibatch slicing of the reference (train.py:63-73, 400-410), including the global-numpy-stream side effect."""
import numpy as np

from rl_rubiks_b200.train import Train


def test_evaluation_schedule_matches_reference_rule():
	assert Train.evaluation_schedule(10, 0).tolist() == []
	assert Train.evaluation_schedule(10, 1).tolist() == list(range(10))               # train.py:66-67
	assert Train.evaluation_schedule(10, 4).tolist() == [0, 3, 7, 9]                  # train.py:68-71
	assert Train.evaluation_schedule(9, 4).tolist() == [0, 3, 7, 8]
	assert Train.evaluation_schedule(3, 2).tolist() == [0, 1, 2]


def test_get_batches_slices_and_rng_side_effect():
	np.random.seed(3)
	b = Train._get_batches(30, 16)
	after = np.random.randint(0, 1 << 30)
	assert b == [slice(0, 16), slice(16, 30)]
	np.random.seed(3)
	np.random.shuffle(np.arange(30))                                                  # what train.py:405-406 consumes
	assert after == np.random.randint(0, 1 << 30)
	assert Train._get_batches(32, 16) == [slice(0, 16), slice(16, 32)]
	assert Train._get_batches(5, 16) == [slice(0, 5)]
