"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares."""
import csv, collections, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
rows = list(csv.DictReader(lines))
agg = collections.OrderedDict(); tot = 0.0
seq = []
for row in rows:
    name = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '').replace('wg::', '')
    t = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit']
    t = t / 1000 if u == 'ns' else (t * 1000 if u == 'ms' else t)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t; tot += t
    seq.append((name, row['Grid Size'], t))
print(f"{len(rows)} launches, {tot:.1f} us total (cold-cache, serialised: compare shares)")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:10.1f} us {n:4d} launches {t / tot * 100:5.1f}%  {k}")
if len(sys.argv) > 2:
    for i, (n, g, t) in enumerate(seq[:int(sys.argv[2])]):
        print(i, n[:50], g, f"{t:.1f}")
