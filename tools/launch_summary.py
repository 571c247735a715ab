"""Summarise an ncu launch list (gpu__time_duration.sum CSV): per-kernel totals over the last N launches and the
sequence of the final forward."""
import csv, collections, re, sys
path = sys.argv[1]; last_n = int(sys.argv[2]) if len(sys.argv) > 2 else 0; seq_n = int(sys.argv[3]) if len(sys.argv) > 3 else 0
lines = [l for l in open(path) if l.startswith('"')]
r = csv.reader(lines); hdr = next(r)
ki, vi, gi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size')
def short(k):
    k = re.sub(r'\(.*', '', k).replace('void ', '').replace('dpot::<unnamed>::', '')
    return k
data = [(short(row[ki]), float(row[vi].replace(',', '')), row[gi]) for row in r]
sel = data[-last_n:] if last_n else data
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v, g in sel:
    agg[k][0] += 1; agg[k][1] += v
tot = sum(v for _, v in agg.values())
print(f"# {path}: {len(sel)} launches, {tot/1e3:.1f} us total (ncu per-launch times: cold-cache, serialised)")
print(f"{'kernel':42s} {'n':>5s} {'total us':>10s} {'avg us':>9s} {'share':>7s}")
for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:42s} {n:5d} {v/1e3:10.1f} {v/n/1e3:9.1f} {v/tot*100:6.1f}%")
if seq_n:
    print("\n# last", seq_n, "launches in order")
    for k, v, g in data[-seq_n:]:
        print(f"{k:42s} {g:16s} {v/1e3:8.1f}")
