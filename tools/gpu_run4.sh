#!/bin/bash
tag=${1:-r2f}
timeout 600 python -m pytest tests/test_gpu_bgk.py tests/test_gpu_query.py -x -q -m gpu -k "not config4" > gpurun_out/${tag}_pytest.log 2>&1
tail -4 gpurun_out/${tag}_pytest.log
LA3DM_BENCH_VERBOSE=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_bench.json"))
    print("step %.4f ms" % d["ms_per_step"], "predict %.4f ms" % d["roofline"]["kernel_ms"], "e2e %.4f ms" % d["e2e"]["ms_per_step"], d.get("units_match_oracle_fixture"), "frac", d["roofline"]["frac"], d["roofline"]["bound"])
except Exception as e:
    print("FAILED", e)
PY
