#!/bin/bash
# same-box A/B of library builds: tools/ab_lib.sh a.so b.so ...   (the first run is the library in place)
cp la3dm_b200/lib/libla3dm_b200.so /tmp/lib_orig.so
for v in "" "$@"; do
  if [ -n "$v" ]; then cp "$v" la3dm_b200/lib/libla3dm_b200.so; else cp /tmp/lib_orig.so la3dm_b200/lib/libla3dm_b200.so; fi
  python bench.py --no-cpu-baseline --config4-scans 0 2> gpurun_out/ab.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('[$v]', 'step %.4f predict %.4f e2e %.4f' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step']))"
done
cp /tmp/lib_orig.so la3dm_b200/lib/libla3dm_b200.so
