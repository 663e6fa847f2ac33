# targeted ncu --set full captures of kernels added after the r1k all-kernel summary; only the text summaries travel back
set -x
mkdir -p gpurun_out
rm -f /tmp/p_*.ncu-rep
ncu --set full --clock-control none -k regex:k_select -s 12 -c 2 -o /tmp/p_select python tools/astar_bench.py --cubes 64 --depth 1000 --cheap-net > /dev/null 2>&1
ncu --set full --clock-control none -k regex:k_render_from2024 -c 2 -o /tmp/p_render python profiles/measure_configs.py --only 686 --once > /dev/null 2>&1
ncu --set full --clock-control none -k regex:k_as_oh -c 2 -o /tmp/p_bf16 python profiles/measure_configs.py --only as_oh_2024 --once > /dev/null 2>&1
for f in select render bf16; do python profiles/summarize_ncu.py kernels /tmp/p_$f.ncu-rep > gpurun_out/r1o_ncu_$f.txt 2>&1; done
wc -l gpurun_out/r1o_ncu_*.txt
