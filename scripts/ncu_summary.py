"""Summarises an ``ncu --set full`` report (``.ncu-rep``) as a small markdown table for ``profiles/``.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [--title "..."] > profiles/rNN_ncu_<what>.md

Needs only the ``ncu`` CLI (no GPU): it reads ``ncu -i <rep> --page raw --csv``.
"""
import argparse
import csv
import io
import subprocess
import sys

METRICS = [
    ('gpu__time_duration.sum', 'duration (cold cache, serialised by ncu)'),
    ('sm__cycles_elapsed.avg', 'SM cycles elapsed'),
    ('sm__cycles_elapsed.avg.per_second', 'SM clock during the capture'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('launch__registers_per_thread', 'registers / thread'),
    ('launch__shared_mem_per_block_dynamic', 'dynamic shared memory / CTA'),
    ('dram__bytes_read.sum', 'DRAM bytes read'),
    ('dram__bytes_write.sum', 'DRAM bytes written'),
    ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput, % of peak'),
    ('lts__t_bytes.sum', 'L2 bytes'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput, % of peak'),
    ('sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'tensor pipe cycles active (realtime), % of elapsed'),
    ('sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg', 'tensor HMMA sub-pipe cycles active (sum of the 4 sub-partitions of an SM)'),
    ('sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tensor-memory (TMEM) cycles active, % of elapsed'),
    ('sm__inst_executed.sum', 'warp instructions executed'),
    ('smsp__inst_executed.sum', 'warp instructions executed (SMSP)'),
    ('sm__inst_executed.avg.per_cycle_elapsed', 'IPC (elapsed)'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy, %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy, %'),
    ('l1tex__t_sector_hit_rate.pct', 'L1/TEX hit rate, %'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'shared-memory bank conflicts'),
    ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall: long scoreboard / issue'),
    ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'stall: short scoreboard / issue'),
    ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'stall: wait / issue'),
    ('smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'stall: branch resolving / issue'),
    ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'stall: math pipe throttle / issue'),
    ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'stall: barrier / issue'),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('report')
    ap.add_argument('--title', default=None)
    ap.add_argument('--note', default=None)
    args = ap.parse_args()
    raw = subprocess.run(['ncu', '-i', args.report, '--page', 'raw', '--csv'], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    (hdr, units) = (rows[0], rows[1])
    col = {h: i for (i, h) in enumerate(hdr)}
    for (i, h) in enumerate(hdr):      # 'TPC.TriageCompute.sm__...' -> 'sm__...'
        col.setdefault(h.split('.Triage')[-1].split('.', 1)[-1] if '.Triage' in h else h, i)
    print('# {}'.format(args.title or args.report))
    print()
    print('Source: `ncu --set full --clock-control none --import-source on` on one B200 (`{}`), read with '
          '`ncu -i ... --page raw --csv`. Durations under ncu are cold-cache and serialised; the bench line is '
          'the timing.'.format(args.report.split('/')[-1]))
    if args.note:
        print()
        print(args.note)
    for r in rows[2:]:
        name = r[col['Kernel Name']]
        print()
        print('## `{}`'.format(name.split('(')[0].replace('unnamed>::', '')))
        print()
        print('| metric | value |')
        print('|---|---|')
        vals = {}
        for (m, label) in METRICS:
            if m in col and r[col[m]] != '':
                vals[m] = r[col[m]]
                print('| {} (`{}`) | {} {} |'.format(label, m, r[col[m]], units[col[m]]))
        try:
            scale = {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1., 'Gbyte': 1e3}
            rd = float(vals['dram__bytes_read.sum'].replace(',', ''))*scale[units[col['dram__bytes_read.sum']]]
            wr = float(vals['dram__bytes_write.sum'].replace(',', ''))*scale[units[col['dram__bytes_write.sum']]]
            print('| **DRAM traffic (read + write)** | {:.3f} Mbyte |'.format(rd + wr))
        except (KeyError, ValueError):
            pass
        try:
            h = float(vals['sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg'].replace(',', ''))
            c = float(vals['sm__cycles_elapsed.avg'].replace(',', ''))
            print('| **tensor sub-pipe active / (4 x SM cycles elapsed)** | {:.1f} % |'.format(100.*h/(4.*c)))
        except (KeyError, ValueError):
            pass


if __name__ == '__main__':
    sys.exit(main())
