#!/bin/bash
# round-2 evidence on one GPU: writes into gpurun_out/
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-e2e --no-secondary"
python profiles/small_case_latency.py > gpurun_out/r02_small_case_latency.txt 2>&1
bash profiles/sweep_workloads.sh --no-secondary > gpurun_out/r02_sweep.txt 2>&1
for w in era5 planar regional; do
  $B --workload $w --steps 2 --warmup 1 > /dev/null 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --kernel-name-base demangled --log-file gpurun_out/r02_launches_$w.csv $B --workload $w --steps 2 --warmup 1 > /dev/null 2>&1
done
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:FreeArc -c 1 -o gpurun_out/r02_freearc -f $B --steps 1 --warmup 0 > gpurun_out/r02_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:ObstacleArc -c 1 -o gpurun_out/r02_obstacle -f $B --steps 1 --warmup 0 >> gpurun_out/r02_ncu_full.log 2>&1
python profiles/capture_hot_kernel.py > gpurun_out/r02_capture.log 2>&1
cp profiles/r02_hot_kernel_counters.json gpurun_out/
ls -la gpurun_out
