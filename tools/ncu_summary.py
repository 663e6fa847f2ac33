#!/usr/bin/env python
"""Turns an ncu report (.ncu-rep, read here with `ncu -i ... --page raw --csv`) into the compact JSON that bench.py reads at run
time for `roofline.traffic` and `roofline.limiter` (so those fields are measured values of a committed capture, not typed in).

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-substring] > profiles/r2_scramble_macro3_ncu.json
"""
import csv
import json
import subprocess
import sys

KEEP = [
	"gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
	"l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
	"sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
	"l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
	"sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
	"smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
	"launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
	"smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
	"smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
	"smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
	"smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def main():
	rep, want = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
	text = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
	rows = list(csv.reader(text.splitlines()))
	hdr, units = rows[0], rows[1]
	idx = {h: i for i, h in enumerate(hdr)}
	out = []
	for r in rows[2:]:
		name = r[idx["Kernel Name"]]
		if want not in name:
			continue
		m = {}
		for k in KEEP:
			if k in idx and r[idx[k]] not in ("", "n/a"):
				m[k] = float(r[idx[k]].replace(",", "")) * SCALE.get(units[idx[k]], 1.0)
		out.append({"kernel": name.split("(")[0], "units": "bytes, microseconds, percent, counts", "metrics": m})
	json.dump({"report": rep.split("/")[-1], "launches": out}, sys.stdout, indent=1)
	print()


if __name__ == "__main__":
	main()
