#!/bin/bash
# Scramble-kernel sweep over depths (parity on a subsample + median ms per depth): bash tools/sweep3.sh on the GPU box.
# RB_SCRAMBLE_MOVES_PER_ROW=2|3, RB_SCRAMBLE_R2=1|2|4 and RB_SCRAMBLE_THREADS select kernel variants (rb_scramble_macro.cuh).
python -m pytest tests/test_gpu_parity.py -q -x -k "scramble" 2>&1 | tail -1
for d in ${DEPTHS:-20 24 32 48 50 64 96 99 100 101 128 192 250 256}; do echo "depth $d: $(DEPTH=$d python tools/scramble_sweep.py 2>&1 | tail -1)"; done
