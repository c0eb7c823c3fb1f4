#!/bin/bash
# multi-GPU bench: tools/gpu_run_n.sh N tag
N=$1; tag=$2
NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_${N}gpu.json 2> gpurun_out/${tag}_bench_${N}gpu.err
tail -3 gpurun_out/${tag}_bench_${N}gpu.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_bench_${N}gpu.json"))
    print("N=$N step %.4f ms" % d["ms_per_step"], "predict %.4f ms" % d["roofline"]["kernel_ms"], "e2e %.4f ms" % d["e2e"]["ms_per_step"], d["final_leaves"], "config4:", d["config4"] if not d["config4"] or "error" in d["config4"] else (d["config4"]["ms_per_step"], d["config4"]["predict_ms"], d["config4"]["value"]))
except Exception as e:
    print("FAILED", e)
PY
