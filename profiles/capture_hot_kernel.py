"""ncu counters of the hot kernel (FreeArcPhase) at a bench workload -> profiles/r02_hot_kernel_counters.json.

    python profiles/capture_hot_kernel.py [--workload era5] [bench flags...]        (on a GPU box, one GPU)

Runs ONE solve of the workload under `ncu --clock-control none` restricted to the free-arc kernel, sums the counters over
the free-arc launches of that solve, and divides the thread-level fp64 instruction counts by the RK steps the same solve
took inside that kernel (dabry_stats.rk_steps_free, read from a second, unprofiled solve of the same inputs: the solver is
deterministic).  bench.py reads the file: `roofline.fp64_flop_per_rk_step`, `roofline.traffic` and
`roofline.ncu_at_this_workload` are printed only for the workload signature they were captured on.
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, 'profiles', 'r02_hot_kernel_counters.json')

SUM_METRICS = [
    'gpu__time_duration.sum',
    'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum',
    'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum',
    'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum',
    'smsp__inst_executed.sum',
    'smsp__thread_inst_executed.sum',
    'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
    'l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum',
    'l1tex__t_requests_pipe_lsu_mem_local_op_st.sum',
    'l1tex__data_pipe_lsu_wavefronts.sum',
]
AVG_METRICS = [  # weighted by launch duration
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
    'l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct',
    'lts__t_sector_hit_rate.pct',
    'launch__registers_per_thread',
]


def main():
    argv = sys.argv[1:]
    import bench
    ap_workload, roots, dtype, integ, small = 'era5', 0, 'f64', 'dopri5', False
    for i, a in enumerate(argv):
        if a == '--workload':
            ap_workload = argv[i + 1]
        if a == '--roots-per-gpu':
            roots = int(argv[i + 1])
        if a == '--dtype':
            dtype = argv[i + 1]
        if a == '--integrator':
            integ = argv[i + 1]
        if a == '--small':
            small = True
    roots = roots or bench.default_roots(ap_workload)
    sig = bench.workload_signature(ap_workload, roots, dtype, integ, small)
    base = [sys.executable, os.path.join(ROOT, 'bench.py'), '--no-cpu-baseline', '--no-e2e', '--no-secondary'] + argv
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    log = os.path.join(ROOT, 'gpurun_out', 'capture_%s.csv' % sig.replace('|', '_'))
    # one solve under ncu (steps 1, warm-up 0): every FreeArc launch of it
    subprocess.run(['ncu', '--metrics', ','.join(SUM_METRICS + AVG_METRICS), '--clock-control', 'none',
                    '--kernel-name-base', 'demangled', '-k', 'regex:FreeArc', '-c', '200', '--csv', '--log-file', log]
                   + base + ['--steps', '1', '--warmup', '0'], check=True, stdout=subprocess.DEVNULL)
    # the same solve unprofiled: RK steps inside the free-arc kernel
    out = subprocess.run(base + ['--steps', '1', '--warmup', '0'], check=True, capture_output=True, text=True).stdout
    line = json.loads(out.strip().splitlines()[-1])
    rows = [r for r in csv.reader(l for l in open(log) if not l.startswith('=='))]
    hdr, rows = rows[0], rows[1:]
    iid, imetric, ival, iname = hdr.index('ID'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Kernel Name')
    launches = {}
    for r in rows:
        launches.setdefault(int(r[iid]), {'name': r[iname]})[r[imetric]] = float(r[ival].replace(',', ''))
    tot = {m: sum(v.get(m, 0.) for v in launches.values()) for m in SUM_METRICS}
    dur = tot['gpu__time_duration.sum']
    avg = {m: sum(v.get(m, 0.) * v.get('gpu__time_duration.sum', 0.) for v in launches.values()) / max(dur, 1.) for m in AVG_METRICS}
    steps_free = float(line['roofline'].get('rk_steps_free_per_solve') or 0.)
    if steps_free <= 0:
        raise SystemExit('bench line carries no rk_steps_free_per_solve')
    dfma, dmul, dadd = (tot['smsp__sass_thread_inst_executed_op_%s_pred_on.sum' % o] for o in ('dfma', 'dmul', 'dadd'))
    entry = {
        'signature': sig, 'kernel': sorted({v['name'] for v in launches.values()})[0], 'launches_per_solve': len(launches),
        'rk_steps_free_per_solve': steps_free,
        'fp64_thread_inst_per_rk_step': (dfma + dmul + dadd) / steps_free,
        'fp64_flop_per_rk_step': (2 * dfma + dmul + dadd) / steps_free,
        'dram_bytes_per_launch': (tot['dram__bytes_read.sum'] + tot['dram__bytes_write.sum']) / len(launches),
        'kernel_ms_under_ncu': dur / 1e6,
        'counters': {
            'fp64_pipe_active_pct': round(avg[AVG_METRICS[0]], 2), 'issue_active_pct': round(avg[AVG_METRICS[1]], 2),
            'warps_active_pct': round(avg[AVG_METRICS[2]], 2), 'l1_data_pipe_lsu_wavefronts_pct': round(avg[AVG_METRICS[3]], 2),
            'l1_hit_pct_global_loads': round(avg[AVG_METRICS[4]], 2), 'l2_hit_pct': round(avg[AVG_METRICS[5]], 2),
            'registers_per_thread': round(avg[AVG_METRICS[6]]),
            'active_lanes_per_warp_instruction': round(tot['smsp__thread_inst_executed.sum'] / tot['smsp__inst_executed.sum'], 2),
            'warp_instructions': tot['smsp__inst_executed.sum'],
            'thread_dfma': dfma, 'thread_dmul': dmul, 'thread_dadd': dadd,
            'local_load_requests': tot['l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum'],
            'local_store_requests': tot['l1tex__t_requests_pipe_lsu_mem_local_op_st.sum'],
            'global_load_requests': tot['l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum'],
            'l1_data_pipe_wavefronts': tot['l1tex__data_pipe_lsu_wavefronts.sum'],
            'dram_read_bytes': tot['dram__bytes_read.sum'], 'dram_write_bytes': tot['dram__bytes_write.sum'],
        },
    }
    data = {}
    if os.path.exists(OUT):
        with open(OUT) as f:
            data = json.load(f)
    data[sig] = entry
    with open(OUT, 'w') as f:
        json.dump(data, f, indent=1, sort_keys=True)
    print(json.dumps(entry, indent=1))


if __name__ == '__main__':
    main()
