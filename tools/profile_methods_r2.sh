#!/bin/bash
# ncu --set full summaries of the other methods' kernels (one scan each) -> gpurun_out/r2_methods_ncu.txt
set -x
out=gpurun_out/r2_methods_ncu.txt
: > $out
timeout 300 ncu --set full --clock-control none -k regex:k_gp_ --launch-skip 48 -c 7 -o gpurun_out/r2_gp python tools/gp_small.py > /dev/null 2>&1
echo "== GPOctoMap, sim_unstructured scan (tools/gp_small.py): k_gp_*" >> $out
python tools/ncu_summary.py gpurun_out/r2_gp.ncu-rep >> $out 2>&1
timeout 300 ncu --set full --clock-control none -k regex:k_bgkl_ --launch-skip 60 -c 8 -o gpurun_out/r2_bgkl python tests/perf/big_methods.py bgkl 65536 > /dev/null 2>&1
echo "== BGKLOctoMap, synthetic 64 k-point scan (tests/perf/big_methods.py bgkl 65536): k_bgkl_*" >> $out
python tools/ncu_summary.py gpurun_out/r2_bgkl.ncu-rep >> $out 2>&1
timeout 300 ncu --set full --clock-control none -k regex:k_lv_ --launch-skip 26 -c 10 -o gpurun_out/r2_lv python tests/perf/big_methods.py bgklv 16384 > /dev/null 2>&1
echo "== BGKLVOctoMap, synthetic 16 k-point scan (tests/perf/big_methods.py bgklv 16384): k_lv_*" >> $out
python tools/ncu_summary.py gpurun_out/r2_lv.ncu-rep >> $out 2>&1
grep -c "^kernel:" $out
