"""dynamic SASS opcode histogram (warp instructions per MX block of a warp) from `ncu --page source --csv --print-source sass`
usage: ncu_dyn.py src.csv n_elements block_size"""
import csv, re, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
nelem, bs = int(sys.argv[2]), int(sys.argv[3])
ks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {'name': r[1], 'hdr': None, 'data': []}; ks.append(cur); continue
    if cur is None: continue
    if cur['hdr'] is None: cur['hdr'] = r; continue
    cur['data'].append(r)
for n, kk in enumerate(ks):
    hdr, data = kk['hdr'], kk['data']
    iS = hdr.index("Source"); iE = hdr.index("Instructions Executed"); iSmp = hdr.index("Warp Stall Sampling (All Samples)")
    W = nelem / bs / 32
    tot = sum(int(r[iE]) for r in data)
    print(kk['name'], "warp instructions per block-warp %.1f = %.2f per element" % (tot / W, tot / W / bs))
    c = collections.Counter(); s = collections.Counter()
    for r in data:
        op = re.sub(r'^@!?U?P\d+\s+', '', r[iS].strip()).split()[0].split('.')[0]
        c[op] += int(r[iE]) / W; s[op] += int(r[iSmp])
    ts = sum(s.values())
    for k, v in c.most_common(34):
        print("   %-10s %7.1f  stall%% %5.1f" % (k, v, 100 * s[k] / ts))
    with open('/tmp/dyn_%d.txt' % n, 'w') as f:
        for r in data:
            f.write("%6.2f %5s  %s\n" % (int(r[iE]) / W, r[iSmp], r[iS]))
