"""Aggregate an ncu launch list (gpu__time_duration.sum, --csv) by kernel + grid: python profiles/launch_summary.py file.csv [N]"""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
agg = collections.OrderedDict()
tot = 0.0
for x in csv.DictReader(lines):
    if x['Metric Name'] != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(mmnn.*', '', x['Kernel Name']).replace('void ', '').replace('mmnn::', '')
    name = re.sub(r'\((int|bool)\)', '', name)[:48]
    v = float(x['Metric Value'].replace(',', ''))
    v = v / 1000 if x['Metric Unit'] == 'ns' else (v * 1000 if x['Metric Unit'] == 'ms' else v)
    a = agg.setdefault((name, x['Grid Size']), [0, 0.0]); a[0] += 1; a[1] += v; tot += v
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print(f"total {tot/1000:.3f} ms over {sum(a[0] for a in agg.values())} launches")
for (k, g), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:n]:
    print(f"{t:8.1f} us {c:4d}x {t/c:7.1f} us  {k:48s} {g}")
