#!/bin/bash
# A/B of kernel variants on one box: every build_variants/*.so through the same bench command (DABRY_B200_LIBRARY),
# plus one ncu pass for the lane-utilisation counters of the hot kernel.  Output: gpurun_out/ab_<name>.log
mkdir -p gpurun_out
for so in build_variants/*.so; do
  name=$(basename $so .so)
  DABRY_B200_LIBRARY=$PWD/$so python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/ab_$name.log 2>&1
  tail -1 gpurun_out/ab_$name.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$name', '%.4g steps/s' % d['value'], d['ms_per_step'], d['config']['phase_ms_rank0'], 'cost %.17g' % d['config']['arrival_cost_global'], 'steps/solve %d' % d['roofline']['rk_steps_free_per_solve'])"
  if [ -n "$AB_NCU" ]; then
    DABRY_B200_LIBRARY=$PWD/$so ncu --metrics l1tex__data_pipe_lsu_wavefronts.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,launch__registers_per_thread,gpu__time_duration.sum \
      --clock-control none --kernel-name-base demangled -k regex:FreeArc -c 1 --csv --log-file gpurun_out/ab_ncu_$name.csv \
      python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e "$@" > /dev/null 2>&1
    grep -v "^==" gpurun_out/ab_ncu_$name.csv | python -c "import csv,sys; [print(\"   \", r[-3], r[-1]) for r in list(csv.reader(sys.stdin))[1:]]"
  fi
done
