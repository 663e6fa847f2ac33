python -m pytest tests/test_gpu_parity.py -q -x -k "scramble" 2>&1 | tail -2
for d in 21 24 50 99 100 101 250; do echo "depth $d: $(DEPTH=$d python tools/scramble_sweep.py 2>&1 | tail -1)"; done
