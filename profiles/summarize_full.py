#!/usr/bin/env python
"""Summarise an `ncu --set full` capture:  ncu -i X.ncu-rep --page raw --csv > raw.csv ; summarize_full.py raw.csv [out.json]

Prints one line per profiled launch (duration, DRAM read/write bytes, DRAM %, tensor-pipe %, occupancy, registers,
grid) and, with a second argument, writes the per-kernel mean DRAM traffic as JSON (bench.py reads
profiles/traffic_<workload>.json to fill `roofline.traffic`)."""
import collections
import csv
import json
import sys

COLS = [('Kernel Name', 'kernel', 44), ('gpu__time_duration.sum', 'us', 8), ('dram__bytes_read.sum', 'rd_MB', 8),
        ('dram__bytes_write.sum', 'wr_MB', 8), ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%', 7),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor%', 8),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%', 6), ('launch__registers_per_thread', 'regs', 5),
        ('launch__grid_size', 'grid', 6), ('launch__block_size', 'block', 6)]


def to_bytes(v, unit):
    v = float(v.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)


def to_us(v, unit):
    v = float(v.replace(',', ''))
    return v * {'ns': 1e-3, 'us': 1, 'ms': 1e3, 's': 1e6}.get(unit, 1)


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, body = rows[0], rows[1], rows[2:]
    ix = {name: hdr.index(name) for name, _, _ in COLS if name in hdr}
    print(' '.join('%-*s' % (w, short) for _, short, w in COLS))
    per = collections.defaultdict(lambda: collections.defaultdict(list))
    for r in body:
        out = []
        for name, short, w in COLS:
            if name not in ix:
                out.append('%-*s' % (w, '-'))
                continue
            v, u = r[ix[name]], units[ix[name]]
            if short == 'kernel':
                v = v.replace('psm::', '')[:w]
            elif short == 'us':
                t = to_us(v, u)
                per[r[ix['Kernel Name']].split('(')[0]]['us'].append(t)
                v = '%.2f' % t
            elif short in ('rd_MB', 'wr_MB'):
                b = to_bytes(v, u)
                per[r[ix['Kernel Name']].split('(')[0]][short].append(b)
                v = '%.3f' % (b / 1e6)
            elif short.endswith('%'):
                v = '%.1f' % float(v.replace(',', ''))
            out.append('%-*s' % (w, v))
        print(' '.join(out))
    summary = {}
    for k, d in per.items():
        n = len(d['us'])
        summary[k.replace('void ', '').replace('psm::', '')] = {
            'launches_profiled': n, 'mean_us': sum(d['us']) / n,
            'dram_read_bytes': sum(d['rd_MB']) / n, 'dram_write_bytes': sum(d['wr_MB']) / n,
            'dram_traffic_bytes': (sum(d['rd_MB']) + sum(d['wr_MB'])) / n}
    print()
    for k, v in summary.items():
        print('%-40s n=%2d  mean %.2f us  traffic %.3f MB/launch' % (k, v['launches_profiled'], v['mean_us'], v['dram_traffic_bytes'] / 1e6))
    if len(sys.argv) > 2:
        json.dump(summary, open(sys.argv[2], 'w'), indent=1, sort_keys=True)


if __name__ == '__main__':
    main()
