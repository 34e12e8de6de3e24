"""Summarises ncu output into small text files for profiles/.
   python tools/ncu_summary.py launches <launches.csv>          -> per-kernel totals / share of the run
   python tools/ncu_summary.py full <report.ncu-rep>            -> key metrics + top stall PCs per kernel"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict


def launches(path):
    rows = [r for r in csv.reader(open(path, errors='replace')) if len(r) > 5]
    hdr = rows[0]
    ik, iv, im = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name')
    tot = OrderedDict()
    for r in rows[1:]:
        if r[im] != 'gpu__time_duration.sum':
            continue
        name = r[ik].split('(')[0]
        ns = float(r[iv].replace(',', ''))
        t = tot.setdefault(name, [0, 0.0])
        t[0] += 1
        t[1] += ns
    total = sum(v[1] for v in tot.values())
    print('# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches)')
    print('%-60s %8s %12s %10s %7s' % ('kernel', 'launches', 'total_us', 'avg_us', 'share'))
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print('%-60s %8d %12.1f %10.2f %6.1f%%' % (k[:60], v[0], v[1] / 1e3, v[1] / v[0] / 1e3, 100 * v[1] / total))
    print('%-60s %8d %12.1f' % ('TOTAL', sum(v[0] for v in tot.values()), total / 1e3))


KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform.sum', 'sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tc.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__shared_mem_per_block_dynamic', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum']


def full(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for n, r in enumerate(rows[2:]):
        print('=== launch %d: %s' % (n, r[hdr.index('Kernel Name')][:90]))
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print('  %-62s %16s %s' % (k, r[i], units[i]))
        st = []
        for i, h in enumerate(hdr):
            if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('_not_issued'):
                try:
                    st.append((float(r[i]), h[len('smsp__pcsamp_warps_issue_stalled_'):]))
                except ValueError:
                    pass
        tot = sum(s[0] for s in st) or 1.0
        print('  stall samples: ' + ', '.join('%s %.0f%%' % (nm, 100 * v / tot) for v, nm in sorted(st, reverse=True)[:7]))
    src = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    blocks, cur = [], None
    for r in rows:
        if len(r) >= 2 and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'hdr': None, 'rows': []}
            blocks.append(cur)
        elif cur is not None and cur['hdr'] is None and r and r[0] == 'Address':
            cur['hdr'] = r
        elif cur is not None and cur['hdr'] is not None and len(r) == len(cur['hdr']):
            cur['rows'].append(r)
    for n, b in enumerate(blocks):
        h = b['hdr']
        if not h:
            continue
        isrc, isamp = h.index('Source'), h.index('# Samples')
        data = []
        for r in b['rows']:
            try:
                data.append((int(r[isamp]), r[isrc]))
            except ValueError:
                data.append((0, r[isrc]))
        tot = sum(d[0] for d in data) or 1
        print('=== launch %d top stall PCs (%d samples): %s' % (n, tot, b['name'][:70]))
        top = sorted(range(len(data)), key=lambda i: -data[i][0])[:14]
        for i in sorted(top):
            prev = data[i - 1][1].strip()[:64] if i else ''
            print('  %5d %5.1f%%  %-70s <- %s' % (i, 100.0 * data[i][0] / tot, data[i][1].strip()[:70], prev))


def traffic(paths):
    """JSON: kernel tag -> mean (dram__bytes_read + dram__bytes_write) per launch, from `ncu --set full` reports."""
    import json
    import re
    tags = [('block_bwd_fused_reduce', 'block_wgrad'), ('post2_xent', 'softmax_xent'), ('block_bwd_chain', 'block_bwd_chain'), ('block_wgrad_h_all', 'block_wgrad'), ('generator_pipe', 'generator_pipe'), ('block_fwd_chain', 'block_fwd'), ('block_fwd_h', 'block_fwd'), ('block_fwd_umma', 'block_fwd'), ('block_bwd_pre_umma', 'block_bwd_pre'), ('block_bwd_dx_umma', 'block_bwd_dx'),
            ('block_wgrad_all', 'block_wgrad'), ('block_wgrad_umma', 'block_wgrad'), ('gemm_umma_kernel', 'gemm_umma'), ('generator_lat', 'generator_lat'),
            ('generator_kernel', 'generator')]
    acc = {}
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    for path in paths:
        raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        ir, iw, ik, it = (hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum'), hdr.index('Kernel Name'),
                          hdr.index('gpu__time_duration.sum'))
        for r in rows[2:]:
            name = r[ik]
            tag = next((t for k, t in tags if k in name), None)
            if tag is None:
                continue
            b = float(r[ir].replace(',', '')) * scale[units[ir]] + float(r[iw].replace(',', '')) * scale[units[iw]]
            a = acc.setdefault(tag, {'launches': 0, 'dram_bytes': 0.0, 'us': 0.0})
            a['launches'] += 1
            a['dram_bytes'] += b
            a['us'] += float(r[it].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'usecond': 1.0, 'nsecond': 1e-3, 'msecond': 1e3}.get(units[it], 1.0)
    out = {k: {'dram_bytes_per_launch': v['dram_bytes'] / v['launches'], 'ncu_us_per_launch': v['us'] / v['launches'],
               'launches_profiled': v['launches']} for k, v in acc.items()}
    print(json.dumps(out, indent=1, sort_keys=True))


if __name__ == '__main__':
    if sys.argv[1] == 'traffic':
        traffic(sys.argv[2:])
    else:
        (launches if sys.argv[1] == 'launches' else full)(sys.argv[2])
