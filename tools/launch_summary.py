"""Dev tool: aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import csv, collections, sys
path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
r = list(csv.reader(lines))
h = r[0]; ix = {n: i for i, n in enumerate(h)}
agg = collections.OrderedDict()
for x in r[1:]:
    if len(x) < len(h) or x[ix['Metric Name']] != 'gpu__time_duration.sum':
        continue
    v = float(x[ix['Metric Value']].replace(',', ''))
    u = x[ix['Metric Unit']]
    v = v / 1e3 if u == 'ns' else v * 1e3 if u == 'ms' else v
    a = agg.setdefault(x[ix['Kernel Name']], [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print(f'# {path}: total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches (cold-cache, serialised per-launch times)')
print('# count   total_us  avg_us  share  kernel')
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{a[0]:5d} {a[1]:10.1f} {a[1]/a[0]:8.1f} {a[1]/tot*100:5.1f}% {n[:140]}')
