"""Summarises ncu captures of bench.py into the tracked files under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv  profiles/r1_launches_summary.md
    python tools/ncu_summary.py full     gpurun_out/prof.ncu-rep  profiles/r1_conv_ncu.md profiles/r1_traffic.json

`launches`: CSV of `ncu --metrics gpu__time_duration.sum --clock-control none --csv` -> per-kernel totals and shares.
`full`: .ncu-rep of `ncu --set full -k regex:"modconv|upconv|up_finish"` over exactly one step -> per-launch table
(duration, grid, tensor-pipe %, L2 %, DRAM %, DRAM bytes) and the DRAM traffic of the conv GEMMs per step (bench.py's
roofline.traffic)."""
import collections
import csv
import json
import re
import subprocess
import sys


def launches(src, dst, cmd_note='', what='python bench.py --steps 2 --warmup 3 --cpu-baseline 0'):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.OrderedDict()
    body = rows[1:]
    if 'Process ID' in hdr and body:           # bench.py re-runs itself once with SGR_FUSE_FIR=0: keep the parent process only
        pi = hdr.index('Process ID')
        body = [r for r in body if r[pi] == body[0][pi]]
    for r in body:
        name = r[ki]
        t = float(r[vi].replace(',', '')) / 1000.0            # ns -> us
        n, tot = agg.get(name, (0, 0.0))
        agg[name] = (n + 1, tot + t)
    total = sum(t for _, t in agg.values())
    with open(dst, 'w') as f:
        f.write('# ncu launch list of `%s`\n\n' % what)
        f.write('Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c <N> --csv --log-file '
                'gpurun_out/launches.csv %s` (B200; cold-cache, '
                'serialised launches: compare shares, not absolutes). %s\n\n' % (what, cmd_note))
        f.write('| kernel | launches | total us | share |\n|---|---|---|---|\n')
        for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('| `%s` | %d | %.1f | %.1f%% |\n' % (name[:90], n, t, 100 * t / total))
    print('wrote', dst)


def full(src, dst_md, dst_json):
    raw = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, data = rows[0], rows[2:]
    cols = [('Kernel Name', 'kernel'), ('gpu__time_duration.sum', 'us'), ('launch__grid_size', 'grid'),
            ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe %'),
            ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM %'),
            ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 %'),
            ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM %'),
            ('dram__bytes_read.sum', 'DRAM rd MB'), ('dram__bytes_write.sum', 'DRAM wr MB')]
    idx = [hdr.index(c) for c, _ in cols]
    conv_bytes = 0.0
    fin_bytes = 0.0
    lines = []
    for r in data:
        vals = [r[i] for i in idx]
        name = re.sub(r'\((int|bool|unsigned int)\)', '', vals[0]).replace('void sgr::', '').replace('sgr::', '').split('(')[0]
        rd, wr = float(vals[7].replace(',', '')), float(vals[8].replace(',', ''))
        if 'up_finish' in name:
            fin_bytes += (rd + wr) * 1e6
        elif 'splitk_finish' in name:
            conv_bytes += (rd + wr) * 1e6        # part of its convolution
        else:
            conv_bytes += (rd + wr) * 1e6
        lines.append('| `%s` | %s | %s | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f |' % (
            name, vals[1], vals[2], float(vals[3]), float(vals[4]), float(vals[5]), float(vals[6]), rd, wr))
    with open(dst_md, 'w') as f:
        f.write('# `ncu --set full` of the conv GEMM + FIR-pass launches of one step (B=32, 256^2, cm=1)\n\n')
        f.write('Command: `ncu --set full --clock-control none --import-source on -k regex:"modconv|upconv|up_finish|splitk_finish" '
                '-s <launches of 3 warm-up steps> -c <launches of one step> -o gpurun_out/prof python bench.py --steps 2 '
                '--warmup 3 --cpu-baseline 0` (the .ncu-rep stays out of git; this table is `ncu -i ... --page raw --csv` '
                'filtered by tools/ncu_summary.py).  Durations under ncu are cold-cache and serialised.\n\n')
        f.write('| ' + ' | '.join(n for _, n in cols) + ' |\n|' + '---|' * len(cols) + '\n')
        f.write('\n'.join(lines) + '\n\n')
        f.write('DRAM traffic per step: conv GEMMs %.1f MB, FIR pass %.1f MB.\n' % (conv_bytes / 1e6, fin_bytes / 1e6))
    with open(dst_json, 'w') as f:
        json.dump({'conv_dram_bytes_per_step': conv_bytes, 'up_finish_dram_bytes_per_step': fin_bytes,
                   'source': 'ncu --set full, one step of bench.py (B=32, 256^2, cm=1); see ' + dst_md}, f, indent=1)
    print('wrote', dst_md, dst_json)


def hbm(src, dst_md, peak_gbs=None):
    """Per-launch table of the HBM-bound kernels: duration, DRAM bytes, achieved GB/s against the measured copy peak,
    sectors per request of the global loads / stores."""
    import os
    if peak_gbs is None:
        try:
            peak_gbs = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         'MEASURED_PEAKS.json')))['hbm_gbs'])
        except Exception:
            peak_gbs = 6650.0
    raw = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    keys = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'dram__bytes_read.sum',
            'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
            'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
            'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
            'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread']
    ci = {k: (hdr.index(k) if k in hdr else None) for k in keys}

    def num(r, k):
        i = ci[k]
        if i is None or r[i] in ('', 'n/a'):
            return float('nan')
        return float(r[i].replace(',', ''))

    def scale(k, table):
        return table.get(units[ci[k]].lower(), 1) if ci[k] is not None else 1
    byte_u = {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}
    time_u = {'ns': 1e-3, 'nsecond': 1e-3, 'us': 1, 'usecond': 1, 'ms': 1e3, 'msecond': 1e3, 'second': 1e6}
    with open(dst_md, 'w') as f:
        f.write('# `ncu --set full` of the HBM-bound kernels (tools/gpu_hbm_kernels.py)\n\n')
        f.write('Command: `ncu --set full --clock-control none -k regex:"upfirdn2d_kernel|torgb_tail|bwd_act|up_bwd_prepare|'
                'frames_to_uint8" -o gpurun_out/prof_hbm python tools/gpu_hbm_kernels.py`; table = `ncu -i ... --page raw '
                '--csv` through tools/ncu_summary.py hbm.  GB/s = (dram read + write bytes) / duration; peak = %.0f GB/s '
                '(MEASURED_PEAKS.json copy bandwidth).  Durations under ncu are cold-cache.\n\n' % peak_gbs)
        f.write('| kernel | grid | us | DRAM rd MB | DRAM wr MB | GB/s | of peak | DRAM % (ncu) | SM % | sectors/req ld | sectors/req st | regs |\n')
        f.write('|---|---|---|---|---|---|---|---|---|---|---|---|\n')
        for r in data:
            name = re.sub(r'\((int|bool|unsigned int)\)', '', r[ci['Kernel Name']]).replace('void sgr::', '').replace('sgr::', '').split('(')[0]
            us = num(r, 'gpu__time_duration.sum') * scale('gpu__time_duration.sum', time_u)
            rd = num(r, 'dram__bytes_read.sum') * scale('dram__bytes_read.sum', byte_u)
            wr = num(r, 'dram__bytes_write.sum') * scale('dram__bytes_write.sum', byte_u)
            gbs = (rd + wr) / (us * 1e-6) / 1e9 if us > 0 else float('nan')
            f.write('| `%s` | %d | %.1f | %.2f | %.2f | %.0f | %.2f | %.1f | %.1f | %.1f | %.1f | %d |\n' % (
                name, num(r, 'launch__grid_size'), us, rd / 1e6, wr / 1e6, gbs, gbs / peak_gbs,
                num(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
                num(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'),
                num(r, 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum') / max(num(r, 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum'), 1),
                num(r, 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum') / max(num(r, 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum'), 1),
                num(r, 'launch__registers_per_thread')))
    print('wrote', dst_md)


if __name__ == '__main__':
    if sys.argv[1] == 'hbm':
        hbm(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == 'launches':
        launches(sys.argv[2], sys.argv[3], *([''] + sys.argv[4:5] if len(sys.argv) > 4 else []))
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4])
