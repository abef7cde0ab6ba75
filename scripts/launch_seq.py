"""Run-length compressed kernel sequence of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = None; seq = []
for r in rows:
    if r[0] == 'ID': hdr = r; continue
    if hdr is None: continue
    d = dict(zip(hdr, r))
    if d.get('Metric Name') != 'gpu__time_duration.sum': continue
    v = float(d['Metric Value'].replace(',', ''))
    u = d['Metric Unit']
    if u == 'us': v *= 1e3
    elif u == 'ms': v *= 1e6
    name = re.sub(r'void |lr::<unnamed>::|\(.*', '', d['Kernel Name'])[:50]
    seq.append((name, v / 1e6, d.get('Grid Size', '')))
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10**9
out = []
for n, v, g in seq:
    if out and out[-1][0] == n: out[-1][1] += v; out[-1][2] += 1
    else: out.append([n, v, 1, g])
for i, (n, v, c, g) in enumerate(out):
    if lo <= i < hi: print(f"{i:4d} {v:9.3f} ms x{c:3d} {n} {g}")
