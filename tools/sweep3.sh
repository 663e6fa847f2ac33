python -m pytest tests/test_gpu_parity.py -q -x -k "scramble" 2>&1 | tail -1
for d in 100 100 32 48 64 96 128 192 256; do echo "depth $d: $(DEPTH=$d python tools/scramble_sweep.py 2>&1 | tail -1)"; done
