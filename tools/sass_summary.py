#!/usr/bin/env python
"""Per-kernel SASS evidence of librubiks_b200.so (cuobjdump -sass): counts of the mnemonics that show how each kernel moves and
permutes its data -- bulk / tensor async copies (UBLKCP / UTMALDG: TMA engine), mbarrier traffic (SYNCS), 128-bit CAS
(ATOMG.E.CAS.128), byte permutes (PRMT), byte dot products (IDP.4A), 16-byte shared-memory loads (LDS.128), streaming stores.

  python tools/sass_summary.py > profiles/r2_sass_summary.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "rl_rubiks_b200", "librubiks_b200.so")
PATTERNS = [("UBLKCP", r"\bUBLKCP"), ("UTMALDG", r"\bUTMALDG"), ("SYNCS", r"\bSYNCS"), ("ATOMG.CAS.128", r"ATOMG\.E\.CAS\.128"), ("ATOMG", r"\bATOMG"),
			("REDG", r"\bREDG"), ("PRMT", r"\bPRMT"), ("IDP.4A", r"\bIDP\.4A"), ("IDP.2A", r"\bIDP\.2A"), ("LOP3", r"\bLOP3"), ("IMAD.HI", r"\bIMAD\.HI"), ("LDS.128", r"\bLDS\.128"),
			("LDS", r"\bLDS\b"), ("STG.128", r"\bSTG\.E\.(EF\.)?128"), ("LDG.128", r"\bLDG\.E\.(\w+\.)*128"), ("SHFL", r"\bSHFL"), ("BAR", r"\bBAR\.")]


def main():
	sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
	demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
	counts, order, cur = collections.defaultdict(collections.Counter), [], None
	for line in sass.splitlines():
		m = re.search(r"Function : (\S+)", line)
		if m:
			cur = m.group(1)
			order.append(cur)
			continue
		if cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
			counts[cur]["instructions"] += 1
			for name, pat in PATTERNS:
				if re.search(pat, line):
					counts[cur][name] += 1
	names = ["instructions"] + [n for n, _ in PATTERNS]
	print("# cuobjdump -sass rl_rubiks_b200/librubiks_b200.so (sm_100a), mnemonic counts per kernel; regenerate with tools/sass_summary.py")
	print("kernel | " + " | ".join(names))
	for fn in order:
		d = demangle(fn).replace("void ", "")
		d = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", d).replace("(int)", "").replace("(bool)", "")      # drop the parameter list
		print(d + " | " + " | ".join(str(counts[fn][n]) for n in names))
	total = collections.Counter()
	for fn in order:
		total.update(counts[fn])
	print("TOTAL | " + " | ".join(str(total[n]) for n in names))


if __name__ == "__main__":
	main()
