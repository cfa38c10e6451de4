#!/bin/bash
for so in build_variants/g*.so; do
  name=$(basename $so .so)
  for w in era5 planar; do
    DABRY_B200_LIBRARY=$PWD/$so python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary --workload $w 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$name $w', 'resident %.2f ms' % d['ms_per_step'], 'e2e %.2f ms' % d['e2e']['ms_per_step'], d['e2e']['ms_each_step_rank0'])"
  done
done
