"""Turn an ncu report (.ncu-rep, `--set full --import-source on`) into the text summary committed under profiles/:
key raw metrics per profiled kernel, warp-stall breakdown, opcode mix and the hottest SASS instructions.
    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/rNN_<name>.txt"""
import collections
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum',
        # the L1 data stage: every 128 B delivered to the register file is one wavefront, broadcast or not
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_local_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio']


def main(rep):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    head, units = rows[0], rows[1]
    for r in rows[2:]:
        print('=' * 100)
        print('kernel:', r[head.index('Kernel Name')])
        for name in WANT:
            if name in head:
                i = head.index(name)
                print('  %-72s %-10s %s' % (name, units[i], r[i]))
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    blocks, cur = [], None
    for r in csv.reader(src.splitlines()):
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'rows': []}
            blocks.append(cur)
        elif cur is not None:
            cur['rows'].append(r)
    seen = set()
    for b in blocks:
        if b['name'] in seen or not b['rows']:
            continue
        seen.add(b['name'])
        h = b['rows'][0]
        idx = {n: i for i, n in enumerate(h)}
        stalls = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
        tot, ops, n_inst, hot = collections.Counter(), collections.Counter(), 0, []
        for r in b['rows'][1:]:
            if len(r) < len(h):
                continue
            for s in stalls:
                tot[s] += int(r[idx[s]] or 0)
            op = [o for o in r[idx['Source']].split() if not o.startswith('@')]
            op = op[0].split('.')[0] if op else '?'
            n = int(r[idx['Instructions Executed']] or 0)
            ops[op] += n
            n_inst += n
            hot.append((int(r[idx['# Samples']] or 0), r[idx['Source']].strip()[:80]))
        total = max(1, sum(tot.values()))
        print('-' * 100)
        print('source page:', b['name'][:90])
        print('  SASS instructions: %d   warp instructions executed: %d' % (len(b['rows']) - 1, n_inst))
        print('  warp stall samples (%):', ', '.join('%s %.1f' % (k[6:], 100 * v / total) for k, v in tot.most_common(9)))
        print('  opcode mix (%):', ', '.join('%s %.1f' % (k, 100 * v / max(1, n_inst)) for k, v in ops.most_common(16)))
        print('  hottest instructions (samples):')
        for s, text in sorted(hot, reverse=True)[:10]:
            print('    %7d  %s' % (s, text))


if __name__ == '__main__':
    main(sys.argv[1])
