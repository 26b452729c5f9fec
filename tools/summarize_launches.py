"""Summarises an `ncu --metrics gpu__time_duration.sum,... --csv` launch list by kernel name."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        hdr, start = r, i + 1
        break
iK, iM, iV = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value')
iU = hdr.index('Metric Unit')
t, n, by = collections.defaultdict(float), collections.Counter(), collections.defaultdict(float)
scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1, 'us': 1e3, 'ms': 1e6}
for r in rows[start:]:
    if len(r) <= iV:
        continue
    k = re.sub(r'\(.*', '', r[iK]).replace('void <unnamed>::', '').replace('<unnamed>::', '')
    v = float(r[iV].replace(',', '')) * scale.get(r[iU], 1)
    if r[iM] == 'gpu__time_duration.sum':
        t[k] += v; n[k] += 1
    elif r[iM].startswith('dram__bytes'):
        by[k] += v
tot = sum(t.values())
print(f"{sum(n.values())} launches, {tot/1e6:.3f} ms GPU time (cold-cache, serialised: compare shares)")
print(f"{'kernel':44s} {'n':>5s} {'ms':>9s} {'share':>6s} {'avg us':>9s} {'dram GB/s':>9s}")
for k, v in sorted(t.items(), key=lambda x: -x[1]):
    print(f"{k:44s} {n[k]:5d} {v/1e6:9.3f} {100*v/tot:5.1f}% {v/n[k]/1e3:9.1f} {by[k]/v:9.0f}")
