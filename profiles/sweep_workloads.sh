#!/bin/bash
# One bench line per BASELINE config on one GPU (plus the fp32-grid and RK4 variants of config 4): gpurun_out/sweep_<name>.json
mkdir -p gpurun_out
run() { name=$1; shift; python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" 2> gpurun_out/sweep_$name.err | tail -1 > gpurun_out/sweep_$name.json
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.load(open('gpurun_out/sweep_%s.json' % name))
    e = d.get('e2e') or {}
    print('%-14s %.4g steps/s  %.2f ms/step  sites %s  e2e %.4g steps/s (%.1f ms)  phases %s' % (
        name, d['value'], d['ms_per_step'], d['config']['sites_rank0'], e.get('value', float('nan')),
        e.get('ms_per_step', float('nan')), d['config']['phase_ms_rank0']))
except Exception as exc:
    print(name, 'FAILED', exc)
PY
}
run era5_f64
run era5_f32 --dtype f32
run era5_rk4 --integrator rk4
run gyre_rk4 --workload gyre --integrator rk4
run gyre_dopri5 --workload gyre
run planar --workload planar
run regional --workload regional
