# ncu --set full of every kernel of the library at the quick sizes of measure_configs.py; only the text summary travels back
mkdir -p gpurun_out
rm -f /tmp/p_all.ncu-rep
timeout 540 ncu --set full --clock-control none -k regex:"^k_" -c 80 -o /tmp/p_all python profiles/measure_configs.py --once --quick > gpurun_out/ncu_all.log 2>&1
python profiles/summarize_ncu.py kernels /tmp/p_all.ncu-rep > gpurun_out/r1q_kernels_ncu.txt 2>&1
grep -c "^## " gpurun_out/r1q_kernels_ncu.txt
