#!/usr/bin/env python
"""Puts the UNMODIFIED reference implementation of the hot path under baseline/_ref/ (git-ignored, but it travels to the GPU box
with the rest of the working tree), so that `bench.py --impl reference` and the `cpu_baseline` leg can time the reference's own
`librubiks.cube` functions on the box's host cores (SURVEY 8d "CPU baseline").

The reference is pure Python without a setup.py / pyproject, so `pip install --target baseline/_ref /root/reference` has
nothing to build; this script copies the four files the path needs, byte for byte, and records their SHA-256 beside them.
Run in the build container (needs /root/reference).  Nothing under baseline/_ref is imported by the product or the tests.
"""
import hashlib
import os
import shutil
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["librubiks/__init__.py", "librubiks/cube/__init__.py", "librubiks/cube/cube.py", "librubiks/cube/maps.py"]


def main() -> int:
	if not os.path.isdir(REF):
		print(f"{REF} is not here: baseline/_ref is left as it is", file=sys.stderr)
		return 0 if os.path.isdir(DST) else 1
	lines = []
	for rel in FILES:
		src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
		os.makedirs(os.path.dirname(dst), exist_ok=True)
		shutil.copyfile(src, dst)
		lines.append(f"{hashlib.sha256(open(dst, 'rb').read()).hexdigest()}  {rel}")
	with open(os.path.join(DST, "PROVENANCE.txt"), "w") as f:
		f.write("peleiden/rl-rubiks, copied unmodified from /root/reference by baseline/make_ref.py\n" + "\n".join(lines) + "\n")
	print(f"baseline/_ref: {len(FILES)} files")
	return 0


if __name__ == "__main__":
	sys.exit(main())
