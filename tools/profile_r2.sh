#!/bin/bash
# round-2 evidence on one B200: launch list, full ncu capture of the predict kernel and of the fused front-end, bench line,
# per-method lines.  Outputs under gpurun_out/ (copy the summaries to profiles/).
set -x
tag=${1:-r2}
# 1. launch list of the timed region (cold-cache, serialised: read the SHARE column)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 70 -c 140 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 12 --warmup 3 --no-cpu-baseline --config4-scans 0 > gpurun_out/${tag}_launches.log 2>&1
python tools/launch_summary.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launches.txt
# 2. ncu --set full of one predict launch (scan 8 of the sequence) and of the three fused kernels of the same scan
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_predict_bgk_flat --launch-skip 24 -c 1 \
    -o gpurun_out/${tag}_flat python bench.py --steps 7 --warmup 3 --no-cpu-baseline --config4-scans 0 > gpurun_out/${tag}_ncu_flat.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused_ --launch-skip 32 -c 4 \
    -o gpurun_out/${tag}_fused python bench.py --steps 7 --warmup 3 --no-cpu-baseline --config4-scans 0 > gpurun_out/${tag}_ncu_fused.log 2>&1
python tools/ncu_summary.py gpurun_out/${tag}_flat.ncu-rep > gpurun_out/${tag}_predict_bgk_flat_ncu.txt 2>&1
python tools/ncu_summary.py gpurun_out/${tag}_fused.ncu-rep > gpurun_out/${tag}_fused_ncu.txt 2>&1
ncu -i gpurun_out/${tag}_flat.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${tag}_flat_src.csv 2>/dev/null
python tools/ncu_lines.py gpurun_out/${tag}_flat_src.csv 40 > gpurun_out/${tag}_predict_bgk_flat_hot_lines.txt 2>&1
rm -f gpurun_out/${tag}_flat_src.csv
# 3. the bench line (with the CPU baseline) and the reference arm
timeout 900 python bench.py > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_reference_arm.err
# 4. per-method lines; the legacy front-end next to them
python tests/perf/bench_methods.py > gpurun_out/${tag}_methods.jsonl 2> gpurun_out/${tag}_methods.err
LA3DM_LEGACY_FRONTEND=1 python tests/perf/bench_methods.py > gpurun_out/${tag}_methods_legacy_frontend.jsonl 2>/dev/null
LA3DM_LEGACY_FRONTEND=1 python bench.py --no-cpu-baseline --config4-scans 0 > gpurun_out/${tag}_bench_1gpu_legacy_frontend.json 2>/dev/null
LA3DM_FUSED_TRACE=1 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --config4-scans 0 2>&1 | grep "^\[fused" | tail -3 > gpurun_out/${tag}_fused_phase_trace.txt
