#!/usr/bin/env python
"""profiles/<tag>_configs.jsonl (output of measure_configs.py on the B200) -> markdown table for DESIGN.md 3.7."""
import json
import sys

rows = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")]
print("| kernel (config) | algorithmic bytes / launch | ms (median) | GB/s | frac of HBM peak | throughput |")
print("|---|---:|---:|---:|---:|---|")
for r in rows:
	thr = next((f"{v:.3g} {k[:-6]}/s" for k, v in r.items() if k.endswith("_per_s")), "")
	print(f"| {r['kernel']} | {r.get('bytes', '')} | {r['ms']:.4g} | {r.get('GBps', '')} | {r.get('frac_hbm', '')} | {thr} |")
