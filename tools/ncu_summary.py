#!/usr/bin/env python
"""
Summarise ncu captures into the tracked files under profiles/.

    python tools/ncu_summary.py --tag r1 gpurun_out/prof_img.ncu-rep gpurun_out/prof_gather.ncu-rep ...

For every kernel launch in the reports it extracts duration, registers, occupancy, pipe
utilisation, issue activity, the dominant stall reasons, L1 / L2 / DRAM traffic, and
writes
    profiles/<tag>_ncu_kernels.csv     one row per profiled launch
    profiles/ncu_summary.json          keyed entries bench.py attaches to its roofline
It only reads the .ncu-rep files (through `ncu --page raw --csv`).
"""
import argparse
import csv
import io
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

COLS = {
    'duration_us': 'gpu__time_duration.sum',
    'registers': 'launch__registers_per_thread',
    'grid': 'launch__grid_size',
    'block': 'launch__block_size',
    'warps_active_pct': 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'issue_active_pct': 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'fp64_pipe_pct': 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
    'dmma_pipe_pct': 'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
    'lsu_data_pipe_pct': 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
    'l1_ld_hit_pct': 'l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct',
    'lts_throughput_pct': 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'dram_read': 'dram__bytes_read.sum',
    'dram_write': 'dram__bytes_write.sum',
    'dram_write_pct': 'dram__bytes_write.sum.pct_of_peak_sustained_elapsed',
    'warp_inst': 'smsp__inst_executed.sum',
    'icc_hit_pct': 'sm__icc_request_hit_rate.pct',
    'stall_wait': 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'stall_long_sb': 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'stall_short_sb': 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'stall_math_throttle': 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'stall_not_selected': 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
    'stall_no_inst': 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
    'stall_barrier': 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'stall_lg_throttle': 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
}
UNIT_SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3,
              's': 1e6}


def short_name(name):
    n = name.split('(')[0].replace('void ', '').replace('pm::', '')
    return n.strip()


def read_report(path):
    txt = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        rec = {'report': os.path.basename(path), 'kernel': short_name(r[hdr.index('Kernel Name')])}
        for key, col in COLS.items():
            if col not in hdr:
                rec[key] = None
                continue
            i = hdr.index(col)
            try:
                v = float(r[i].replace(',', ''))
            except ValueError:
                rec[key] = None
                continue
            u = units[i]
            if key in ('dram_read', 'dram_write') or key == 'duration_us':
                v *= UNIT_SCALE.get(u, 1.0)
            rec[key] = v
        out.append(rec)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--tag', default='r1')
    ap.add_argument('--key', action='append', default=[],
                    help='summary key = kernel-substring[:nth] (e.g. backplanes_img_c2=backplanes_img_kernel<false>:0)')
    ap.add_argument('reports', nargs='+')
    args = ap.parse_args()
    recs = []
    for p in args.reports:
        recs += read_report(p)
    os.makedirs(os.path.join(ROOT, 'profiles'), exist_ok=True)
    path = os.path.join(ROOT, 'profiles', f'{args.tag}_ncu_kernels.csv')
    with open(path, 'w', newline='') as f:
        w = csv.DictWriter(f, fieldnames=['report', 'kernel'] + list(COLS))
        w.writeheader()
        for r in recs:
            w.writerow({k: (f'{v:.6g}' if isinstance(v, float) else v) for k, v in r.items()})
    print('wrote', path, len(recs), 'launches')
    spath = os.path.join(ROOT, 'profiles', 'ncu_summary.json')
    summary = json.load(open(spath)) if os.path.exists(spath) else {}
    for spec in args.key:
        key, sel = spec.split('=', 1)
        sub, _, nth = sel.partition(':')
        match = [r for r in recs if sub in r['kernel']]
        if not match:
            print('no launch matches', sub)
            continue
        r = match[int(nth or 0)]
        summary[key] = {
            'kernel': r['kernel'], 'report': r['report'], 'duration_ms_under_ncu': r['duration_us'] / 1e3,
            'dram_bytes_per_launch': (r['dram_read'] or 0) + (r['dram_write'] or 0),
            'fp64_pipe_util': None if r['fp64_pipe_pct'] is None else r['fp64_pipe_pct'] / 100,
            'dmma_pipe_util': None if r['dmma_pipe_pct'] is None else r['dmma_pipe_pct'] / 100,
            'issue_active': None if r['issue_active_pct'] is None else r['issue_active_pct'] / 100,
            'lsu_data_pipe_util': None if r['lsu_data_pipe_pct'] is None else r['lsu_data_pipe_pct'] / 100,
            'registers': r['registers'], 'warp_instructions': r['warp_inst'],
        }
    with open(spath, 'w') as f:
        json.dump(summary, f, indent=1)
    print('wrote', spath, sorted(summary))


if __name__ == '__main__':
    main()
