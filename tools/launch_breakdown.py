"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (second half = timed step)."""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}
seq = [(r[ki][:70], float(r[vi].replace(',', '')) * scale.get(r[ui], 1.0)) for r in rows[1:]]
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
half = seq[int(len(seq) * frac):]
agg = collections.OrderedDict()
for k, v in half:
    a = agg.setdefault(k, [0.0, 0]); a[0] += v; a[1] += 1
tot = sum(v for v, _ in agg.values())
for k, (v, c) in sorted(agg.items(), key=lambda x: -x[1][0])[:int(sys.argv[3]) if len(sys.argv) > 3 else 12]:
    print(f"{v:10.1f} us {c:4d}  {100 * v / tot:5.1f}%  {k}")
print(f"total {tot:.1f} us over {len(half)} launches")
