"""Turn an .ncu-rep (ncu --set full --import-source on) into the text summary committed under profiles/.

    python profiles/summarize.py gpurun_out/prof.ncu-rep <units per launch> [kernel index] > profiles/r01_xxx.txt
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__waves_per_multiprocessor',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct']


def main():
    rep, units = sys.argv[1], float(sys.argv[2])
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, unit_row, data = rows[0], rows[1], rows[2:]
    d = data[which]
    name = d[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'
    print(f'kernel: {name}')
    print(f'units per launch (joints / maps / joint-frames): {units:.6g}')
    vals = {}
    for k in KEYS:
        if k in hdr:
            vals[k] = d[hdr.index(k)]
            print(f'{k} = {d[hdr.index(k)]} {unit_row[hdr.index(k)]}')
    for h in hdr:
        if 'issue_stalled' in h and h.endswith('_per_issue_active.ratio'):
            try:
                v = float(d[hdr.index(h)])
            except ValueError:
                continue
            if v > 0.2:
                print(f'stall {h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")} = {v:.2f} warps per issue')
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    sections, cur = [], None
    for r in rows:
        if 'Source' in r and 'Instructions Executed' in r:      # one header row per profiled kernel
            cur = {'hdr': r, 'rows': []}
            sections.append(cur)
        elif cur is not None:
            cur['rows'].append(r)
    if which < len(sections):
        hdr2, body = sections[which]['hdr'], sections[which]['rows']
        si, ei = hdr2.index('Source'), hdr2.index('Instructions Executed')
        cnt, tot = collections.Counter(), 0
        for r in body:
            if len(r) <= max(si, ei):
                continue
            try:
                n = int(float(r[ei]))
            except ValueError:
                continue
            t = re.sub(r'^@!?U?P\w+\s+', '', r[si].strip())
            op = t.split()[0].split('.')[0] if t else '?'
            cnt[op] += n
            tot += n
        per = units / 32.0
        print(f'warp instructions executed: {tot}  = {tot / per:.1f} per unit (thread-level)')
        print('opcode mix per unit: ' + ' '.join(f'{op}:{n / per:.1f}' for op, n in cnt.most_common(24)))
        sass = ' '.join(cnt)
        print('TMA / Blackwell evidence in executed SASS: ' + ', '.join(op for op in ('UBLKCP', 'SYNCS', 'FFMA2', 'FMUL2', 'UTMALDG') if op in cnt))


if __name__ == '__main__':
    main()
