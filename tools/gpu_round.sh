#!/bin/bash
# One GPU-box visit: parity tests, both bench arms, per-kernel config table, ncu launch list of the bench.
# Usage (from the repo root, on the GPU box):  bash tools/gpu_round.sh <tag>
tag=${1:-r1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_$tag.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>> gpurun_out/bench_$tag.err
python profiles/measure_configs.py > gpurun_out/configs_$tag.jsonl 2> gpurun_out/configs_$tag.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 5 --warmup 3 > gpurun_out/ncu_launch_$tag.log 2>&1
tail -3 gpurun_out/pytest_$tag.log
cat gpurun_out/bench_$tag.json | cut -c1-600
cat gpurun_out/configs_$tag.jsonl
tail -5 gpurun_out/configs_$tag.err
