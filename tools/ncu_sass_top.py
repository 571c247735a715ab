"""Top stalled SASS instructions of one captured launch.
    python tools/ncu_sass_top.py rep.ncu-rep <launch index in report> [top N]"""
import csv, io, subprocess, sys
rep, idx = sys.argv[1], int(sys.argv[2]); top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(idx),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
lines = out.splitlines()
print(lines[0][:150])
end = next((i for i in range(1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
rows = list(csv.reader(io.StringIO("\n".join(lines[1:end]))))
h = rows[0]; data = [r for r in rows[1:] if len(r) == len(h)]
si = h.index("# Samples"); src = h.index("Source"); ex = h.index("Instructions Executed")
stalls = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
tot = sum(int(r[si] or 0) for r in data)
print(f"total samples {tot}, instructions {len(data)}")
order = sorted(range(len(data)), key=lambda i: -int(data[i][si] or 0))[:top]
for i in sorted(order):
    r = data[i]
    st = sorted(((int(r[j] or 0), h[j][6:]) for j in stalls), reverse=True)[:2]
    print(f"{i:5d} {int(r[si]):7d} {100*int(r[si])/max(tot,1):5.1f}%  ex={r[ex]:>8s}  {r[src].strip()[:70]:70s} {st}")
