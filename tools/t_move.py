import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rl_rubiks_b200 import _native as N
from oracle import cube_oracle as O
for n, depth in [(int(a), int(b)) for a, b in (x.split("x") for x in sys.argv[1:])]:
	g = np.random.RandomState(depth * 7 + n)
	acts = g.randint(0, 12, (depth, n)).astype(np.uint8)
	f, d = O.indices_to_actions(acts.T)
	want = O.scramble_many(f, d, True)
	a_t = torch.from_numpy(acts).cuda()
	out = torch.empty(n, 20, dtype=torch.int8, device="cuda")
	N.check(N.lib.rb_scramble(0, N.ptr(a_t), 1, n, None, N.ptr(out), n, depth, N.stream_handle()))
	torch.cuda.synchronize()
	print(n, depth, bool((out.cpu().numpy() == want).all()), flush=True)
