#!/usr/bin/env python3
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections, csv, re, sys
for path in sys.argv[1:]:
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = re.sub(r'\(.*', '', row['Kernel Name'])[:60]
        try:
            v = float(row['Metric Value'].replace(',', ''))
        except ValueError:
            continue
        u = row['Metric Unit']
        v *= {'ns': 1e-3, 'nsecond': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3, 's': 1e6, 'second': 1e6}.get(u, 1.0)
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(path)
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('  %-62s n=%4d total=%10.1f us  avg=%9.1f us  share=%5.1f%%' % (k, n, t, t / n, 100 * t / tot))
