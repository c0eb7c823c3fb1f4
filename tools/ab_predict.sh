#!/bin/bash
# A/B of the BGK predict kernel variants on the headline workload (run on the GPU box): tools/ab_predict.sh tag v1 oct2 oct3 oct4
tag=$1; shift
for v in "$@"; do
  unset LA3DM_PREDICT_V1 LA3DM_OCT_CTAS LA3DM_OCT_NO_HEAVY
  case $v in
    v1) export LA3DM_PREDICT_V1=1;;
    oct2) export LA3DM_OCT_CTAS=2;;
    oct3) export LA3DM_OCT_CTAS=3;;
    oct4) export LA3DM_OCT_CTAS=4;;
    oct1) export LA3DM_OCT_CTAS=1;;
    oct2nh) export LA3DM_OCT_CTAS=2 LA3DM_OCT_NO_HEAVY=1;;
  esac
  LA3DM_BENCH_VERBOSE=1 python bench.py --no-cpu-baseline > gpurun_out/${tag}_bench_$v.json 2> gpurun_out/${tag}_bench_$v.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_bench_$v.json"))
    print("$v", "step %.4f ms" % d["ms_per_step"], "predict %.4f ms" % d["roofline"]["kernel_ms"], "e2e %.4f ms" % d["e2e"]["ms_per_step"], d.get("units_match_oracle_fixture"))
except Exception as e:
    print("$v", "FAILED", e)
PY
  grep "per-scan" gpurun_out/${tag}_bench_$v.err
done
