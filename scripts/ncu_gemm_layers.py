"""Per-layer table of the tap-list GEMM launches of one bench step from an ``ncu --set full`` capture.

    ncu -i gpurun_out/prof_rXX_gemm.ncu-rep --page raw --csv > raw.csv
    python scripts/ncu_gemm_layers.py raw.csv "title" [images] [math] [traffic.json] > profiles/rXX_ncu_full_gemm_layers.md

The capture may start anywhere in a step (take -c 26): the table starts at the first launch after a last-layer launch.
traffic.json (optional) receives the DRAM bytes of the step, which bench.py reads for `roofline.traffic`.

math = tf32x3 (3 MMAs per product everywhere) or mixed (synthesis side: one MMA per product, fused IGDN norms too).
"""
import csv
import sys

NAMES = ['conv1 k9 s4 (uint8 patches in-kernel) + GDN1', 'conv2 k5 s2 + GDN2', 'conv3 k5 s2 + GDN3', 'IGDN4 (standalone)',
         'tconv1 phase (0,0), 4 taps + IGDN5', 'tconv1 phase (0,1), 6 taps + IGDN5', 'tconv1 phase (1,0), 6 taps + IGDN5',
         'tconv1 phase (1,1), 9 taps + IGDN5', 'tconv2 phase (0,0), 4 taps + IGDN6', 'tconv2 phase (0,1), 6 taps + IGDN6',
         'tconv2 phase (1,0), 6 taps + IGDN6', 'tconv2 phase (1,1), 9 taps + IGDN6',
         'tconv3 k9 s4 + col2im gather + BT.601 cast (persistent)']
# algorithmic GFLOP per image and MMA passes actually issued
ALG = [0.5096 + 0.8053, 5.0332 + 0.2013, 1.2583 + 0.0503, 0.0503] + [1.2583*t/25 + 0.2013/4 for t in (4, 6, 6, 9)] + \
      [5.0332*t/25 + 0.8053/4 for t in (4, 6, 6, 9)] + [0.5096]
EXE = {'tf32x3': [0.5096*2 + 0.8053*3, (5.0332 + 0.2013)*3, (1.2583 + 0.0503)*3, 0.0503*3] +
                 [(1.2583*t/25 + 0.2013/4)*3 for t in (4, 6, 6, 9)] + [(5.0332*t/25 + 0.8053/4)*3 for t in (4, 6, 6, 9)] +
                 [0.5096*3],
       'mixed': [0.5096*2 + 0.8053*3, (5.0332 + 0.2013)*3, (1.2583 + 0.0503)*3, 0.0503*3] +
                [(1.2583*t/25 + 0.2013/4) for t in (4, 6, 6, 9)] + [(5.0332*t/25 + 0.8053/4) for t in (4, 6, 6, 9)] +
                [0.5096]}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    title = sys.argv[2] if len(sys.argv) > 2 else 'tap-list GEMM, every layer of one 24-image step'
    images = int(sys.argv[3]) if len(sys.argv) > 3 else 24
    exe = EXE[sys.argv[4] if len(sys.argv) > 4 else 'tf32x3']
    hdr = rows[0]

    def col(name):
        for (i, h) in enumerate(hdr):
            if h == name or h.endswith('.' + name):
                return i
        return None

    c = {k: col(k) for k in ['Kernel Name', 'launch__grid_size', 'gpu__time_duration.sum', 'dram__bytes_read.sum',
                             'dram__bytes_write.sum', 'sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg',
                             'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second', 'launch__registers_per_thread',
                             'launch__shared_mem_per_block_dynamic']}
    print('# ' + title + '\n')
    print('Source: `ncu --set full --clock-control none --import-source on -k "regex:gemm_umma|tconv9s4" -c 14` around '
          '`python bench.py --steps 1 --warmup 1 --depth 1 --no-cpu-baseline --no-parity` on one B200, read with `ncu -i ... --page raw --csv` '
          '(`scripts/ncu_gemm_layers.py`). Durations under ncu are cold-cache and serialised (SM clock {} GHz in the capture); the '
          'bench line is the timing. One launch = one layer (or one output phase of a transposed convolution) over {} images of '
          '512 x 768.\n'.format(rows[2][c['sm__cycles_elapsed.avg.per_second']][:5], images))
    body = rows[2:]
    start = 0
    stitched = False
    for (k, r) in enumerate(body):          # a step starts right after the persistent last-layer kernel of the previous one
        if 'umma7' in r[c['Kernel Name']]:
            if k + 1 + len(NAMES) <= len(body):
                start = k + 1
            elif len(body) >= len(NAMES):
                # the capture straddles two consecutive (identical) steps: the launches after the last layer, then the
                # ones before it that complete the step
                missing = len(NAMES) - (len(body) - (k + 1))
                body = body[k + 1:] + body[k + 1 - missing:k + 1]
                stitched = True
            break
    if stitched:
        print('(The 14 captured launches straddle two consecutive steps of identical work: the first '
              '{} rows of the table are from the later step, the rest from the one before it.)\n'.format(len(NAMES) - missing))
    print('| launch | kernel | grid | duration us | DRAM read MB | DRAM write MB | tensor sub-pipe active / (4 x SM cycles) | '
          'algorithmic TFLOP/s | executed-MMA TFLOP/s |')
    print('|---|---|---|---|---|---|---|---|---|')
    (tt, tr, tw, ta, te) = (0., 0., 0., 0., 0.)
    for (k, r) in enumerate(body[start:start + len(NAMES)]):
        d = float(r[c['gpu__time_duration.sum']])
        rd = float(r[c['dram__bytes_read.sum']])
        wr = float(r[c['dram__bytes_write.sum']])
        try:
            h = float(r[c['sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg']])
            cy = float(r[c['sm__cycles_elapsed.avg']])
            active = '{:.1f} %'.format(100.*h/(4.*cy))
        except ValueError:
            active = 'n/a (counter not collected in this pass)'
        name = r[c['Kernel Name']]
        kern = 'v7' if 'umma7' in name else ('v6' if 'umma6' in name else ('v4' if 'umma4' in name else ('v5' if 'umma5' in name else 'v3')))
        print('| {} | {} | {} | {:.1f} | {:.1f} | {:.1f} | {} | {:.0f} | {:.0f} |'.format(
            NAMES[k], kern, r[c['launch__grid_size']], d, rd, wr, active, ALG[k]*images/d*1e3, exe[k]*images/d*1e3))
        tt += d; tr += rd; tw += wr; ta += ALG[k]*images; te += exe[k]*images
    print('| **all 13 launches** | | | **{:.1f}** | **{:.1f}** | **{:.1f}** | | **{:.0f}** | **{:.0f}** |'.format(
        tt, tr, tw, ta/tt*1e3, te/tt*1e3))
    print()
    print('DRAM traffic of the 13 launches: {:.0f} MB per step = {:.1f} MB per launch on average.'.format(tr + tw, (tr + tw)/13.))
    if len(sys.argv) > 5:
        import json
        with open(sys.argv[5], 'w') as f:
            json.dump({'math': sys.argv[4], 'images': images, 'launches_per_step': len(NAMES), 'dram_bytes_per_step': (tr + tw)*1e6,
                       'source': 'ncu --set full capture summarised by scripts/ncu_gemm_layers.py (dram__bytes_read.sum + dram__bytes_write.sum of the 13 tap-list GEMM launches of one step)'}, f)
            f.write('\n')


if __name__ == '__main__':
    main()
