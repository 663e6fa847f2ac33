for cfg in "3 1 1024" "2 1 1024"; do
  set -- $cfg
  echo "rows=$1 r2=$2 thr=$3: $(RB_SCRAMBLE_MOVES_PER_ROW=$1 RB_SCRAMBLE_R2=$2 RB_SCRAMBLE_THREADS=$3 python tools/scramble_sweep.py 2>&1 | tail -1)"
done
for d in 21 24 50 99 101 250; do echo "depth $d: $(DEPTH=$d python tools/scramble_sweep.py 2>&1 | tail -1)"; done
python -m pytest tests/test_gpu_parity.py -q -x -k "scramble" 2>&1 | tail -3
