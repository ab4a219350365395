import csv, collections, sys
path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    k = row['Kernel Name'][:64]
    v = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit']
    v *= {'ns': 1, 'us': 1e3, 'ms': 1e6, 'nsecond': 1, 'usecond': 1e3, 'msecond': 1e6}.get(u, 1)
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':66s} {'n':>5s} {'total_us':>10s} {'avg_us':>9s} {'share':>6s}")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:66s} {v[0]:5d} {v[1]/1e3:10.1f} {v[1]/v[0]/1e3:9.1f} {v[1]/tot*100:5.1f}%")
