"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv --log-file` launch list by kernel:
launch count, total / mean duration and share of the profiled time. usage: launch_list.py in.csv out.json [note]"""
import csv
import json
import re
import sys


def main(path, out, note=''):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = {}
    for r in rows[1:]:
        name = re.sub(r'\(.*', '', r[ki]).replace('void ', '').replace('<unnamed>::', '')
        name = re.sub(r'<.*', '', name) if name.startswith('at::') else name
        v = float(r[vi].replace(',', ''))
        us = v / 1e3 if r[ui] in ('ns', 'nsecond') else (v if r[ui] in ('us', 'usecond') else v * 1e3)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += us
    tot = sum(a[1] for a in agg.values())
    ks = [dict(kernel=k, launches=a[0], total_us=round(a[1], 1), mean_us=round(a[1] / a[0], 2), share=round(a[1] / tot, 4))
          for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])]
    json.dump(dict(note=note, profiled_us=round(tot, 1), kernels=ks), open(out, 'w'), indent=1)
    for k in ks[:14]:
        print('%-60s n=%5d  %9.1f us  %5.1f%%' % (k['kernel'][:60], k['launches'], k['total_us'], 100 * k['share']))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else '')
