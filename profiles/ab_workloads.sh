#!/bin/bash
# every build_variants/*.so through configs 4 (fp64 and fp32 grid), 3 and 5: gpurun_out/ab_workloads.txt
for so in build_variants/*.so; do
  name=$(basename $so .so)
  for w in "era5" "era5 --dtype f32" "planar" "regional"; do
    DABRY_B200_LIBRARY=$PWD/$so python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary --workload $w 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
p = d['config']['phase_ms_rank0']
print('%-8s %-18s %.4g steps/s  %.2f ms  free arcs %.2f  obstacle arcs %.2f' % ('$name', '$w', d['value'], d['ms_per_step'], p['ms_integrate'], p['ms_obstacle']))"
  done
done
