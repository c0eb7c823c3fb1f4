#!/bin/bash
# same-box A/B of environment switches: tools/ab.sh "VAR=1" ...
for v in "" "$@"; do
  env $v LA3DM_BENCH_VERBOSE=1 python bench.py --no-cpu-baseline --config4-scans 0 2> gpurun_out/ab.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('[$v]', 'step %.4f predict %.4f e2e %.4f' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step']))"
done
