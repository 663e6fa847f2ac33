#!/usr/bin/env python
"""Turns ncu output brought back in gpurun_out/ into the small text summaries committed under profiles/.

  python profiles/summarize_ncu.py launches gpurun_out/launches.csv            > profiles/<round>_launches.txt
  python profiles/summarize_ncu.py kernels  gpurun_out/prof.ncu-rep [regex]    > profiles/<round>_<kernel>_ncu.txt
"""
import csv
import re
import subprocess
import sys

KEYS = [
	"gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
	"gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
	"sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
	"lts__throughput.avg.pct_of_peak_sustained_elapsed",
	"l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
	"l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
	"smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_lsu.sum",
	"sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
	"sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
	"smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
	"launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
	"launch__shared_mem_per_block_static", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
	"sm__cycles_elapsed.max", "smsp__cycles_active.avg",
	"smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
	"smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
	"smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
	"smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
	"smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
	"smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
	"smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
	"smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def launches(path):
	lines = [l for l in open(path).read().splitlines() if l.startswith('"')]
	rows = list(csv.reader(lines))
	hdr = rows[0]
	ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
	agg = {}
	for r in rows[1:]:
		name = re.sub(r"\(.*", "", r[ki])
		v = float(r[vi].replace(",", ""))
		if r[ui] in ("us", "usecond"):
			v *= 1e3
		elif r[ui] in ("ms", "msecond"):
			v *= 1e6
		a = agg.setdefault(name, [0, 0.0])
		a[0] += 1
		a[1] += v
	total = sum(a[1] for a in agg.values())
	print(f"# ncu --metrics gpu__time_duration.sum --clock-control none : {len(rows) - 1} launches, {total / 1e6:.3f} ms total (cold-cache, serialised)")
	print(f"{'kernel':70s} {'launches':>8s} {'total_ms':>10s} {'avg_us':>10s} {'share':>7s}")
	for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
		print(f"{name[:70]:70s} {n:8d} {t / 1e6:10.3f} {t / n / 1e3:10.2f} {t / total:7.3f}")


def kernels(path, pattern=None):
	out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
	rows = list(csv.reader([l for l in out.splitlines() if l.startswith('"')]))
	hdr, units = rows[0], rows[1]
	ki = hdr.index("Kernel Name")
	for r in rows[2:]:
		if pattern and not re.search(pattern, r[ki]):
			continue
		print(f"## {r[ki][:100]}")
		for k in KEYS:
			if k in hdr:
				i = hdr.index(k)
				print(f"{k:90s} {r[i]:>18s} {units[i]}")
		print()


if __name__ == "__main__":
	if sys.argv[1] == "launches":
		launches(sys.argv[2])
	else:
		kernels(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
