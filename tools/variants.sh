set -e
cd la3dm_b200/csrc
for v in 5 6; do
  rm -f build/predict_bgk.o
  make EXTRA="-DLA3DM_PREDICT_MIN_CTAS=$v" > /dev/null 2>&1
  cd ../..
  python -m pytest tests/test_gpu_bgk.py -m gpu -x -q 2>&1 | tail -2
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_var$v.json 2>/dev/null
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_var$v.json'))
print("MIN_CTAS=$v", "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), "predict", round(d["roofline"]["kernel_ms"],4))
PY
  cd la3dm_b200/csrc
done
