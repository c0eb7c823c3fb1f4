#!/bin/bash
# round-2 GPU check: BGK parity (oracle + compiled reference), bench A/B of the predict kernels, launch list
set -x
python -m pytest tests/test_gpu_bgk.py tests/test_gpu_parity_ref.py -x -q -m gpu -k "not config4" > gpurun_out/r2a_pytest.log 2>&1
tail -5 gpurun_out/r2a_pytest.log
LA3DM_BENCH_VERBOSE=1 python bench.py --no-cpu-baseline > gpurun_out/r2a_bench_flat.json 2> gpurun_out/r2a_bench_flat.err
LA3DM_PREDICT_OCT=1 LA3DM_BENCH_VERBOSE=1 python bench.py --no-cpu-baseline > gpurun_out/r2a_bench_oct.json 2> gpurun_out/r2a_bench_oct.err
python - <<'PY'
import json
for v in ("flat", "oct"):
    try:
        d = json.load(open("gpurun_out/r2a_bench_%s.json" % v))
        print(v, "step %.4f ms" % d["ms_per_step"], "predict %.4f ms" % d["roofline"]["kernel_ms"], "e2e %.4f ms" % d["e2e"]["ms_per_step"], d.get("units_match_oracle_fixture"), d["units_per_step"])
    except Exception as e:
        print(v, "FAILED", e)
PY
grep "per-scan" gpurun_out/r2a_bench_flat.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_l.log 2>&1
python tools/launch_summary.py gpurun_out/r2a_launches.csv > gpurun_out/r2a_launches.txt 2>&1; head -30 gpurun_out/r2a_launches.txt
