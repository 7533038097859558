"""aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if 'Kernel Name' in r:
        hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            v = float(d['Metric Value'].replace(',', ''))
        except Exception:
            continue
        u = d.get('Metric Unit', 'ns')
        if u in ('us', 'usecond'): v *= 1e3
        elif u in ('ms', 'msecond'): v *= 1e6
        elif u in ('s', 'second'): v *= 1e9
        k = d['Kernel Name'][:70]
        agg[k][0] += 1; agg[k][1] += v
tot = sum(v[1] for v in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print("%-72s n=%5d  %9.3f ms  %5.1f%%" % (k, n, t / 1e6, 100 * t / tot))
print("total %.3f ms over %d launches" % (tot / 1e6, sum(v[0] for v in agg.values())))
